// radix_sort.cuh — stable LSD radix sort of 16-byte records on the device (hand-written, no CUB).
//
// Used by the finalize stage (K5): the reference writes pair_contact.csv in dict insertion order,
// i.e. by (first contact step, a1, a2) (contacts.py:135-183; pairs of one step are visited row-major,
// close_contacts_calculator.py:104-124); the contact log and the coordinate records are ordered the
// same way.  All three are 16-byte records on the device, so one sorter serves them.
//
// One pass over a digit of <= 8 bits:
//   k_radix_hist     per tile of kSortTile records a 256-bin histogram (shared-memory atomics),
//                    stored bin-major [bin][tile] so that ONE exclusive scan over the whole array
//                    yields, for every (bin, tile), the first output index of that tile's records
//   scan_u32         three-phase exclusive scan of that array (tile scan, scan of totals, add)
//   k_radix_scatter  every warp owns a contiguous piece of the tile and walks it in index order:
//                    __match_any_sync ranks the lanes that share a digit, a per-warp counter row in
//                    shared memory carries the running count, so the rank of a record among the
//                    tile's records with the same digit needs no atomics and the pass is stable
// HBM-bound: per pass 16 B read (hist) + 16 B read + 16 B written per record.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace rsort {

struct Pass {
  int word;   // which 32-bit key word (KeyFn argument), least significant word first
  int shift;  // digit = (word >> shift) & ((1 << bits) - 1)
  int bits;   // 1..8
};

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kItems = 16;                    // records per thread
constexpr int kTile = kThreads * kItems;      // 4096 records per CTA
constexpr int kBins = 256;

template <typename KeyFn>
__global__ void __launch_bounds__(kThreads) k_radix_hist(const uint4* __restrict__ in, unsigned long long n, Pass ps,
                                                         uint32_t* __restrict__ tile_hist, unsigned n_tiles, KeyFn key) {
  __shared__ uint32_t h[kBins];
  h[threadIdx.x] = 0;
  __syncthreads();
  const unsigned long long base = (unsigned long long)blockIdx.x * kTile;
  const uint32_t mask = (1u << ps.bits) - 1u;
#pragma unroll 4
  for (int i = 0; i < kItems; ++i) {
    unsigned long long idx = base + (unsigned long long)i * kThreads + threadIdx.x;
    if (idx < n) atomicAdd(&h[(key(in[idx], ps.word) >> ps.shift) & mask], 1u);
  }
  __syncthreads();
  tile_hist[(size_t)threadIdx.x * n_tiles + blockIdx.x] = h[threadIdx.x];
}

template <typename KeyFn>
__global__ void __launch_bounds__(kThreads) k_radix_scatter(const uint4* __restrict__ in, uint4* __restrict__ out,
                                                            unsigned long long n, Pass ps,
                                                            const uint32_t* __restrict__ tile_ofs, unsigned n_tiles,
                                                            KeyFn key) {
  __shared__ uint32_t wcount[kWarps][kBins];
  for (int i = threadIdx.x; i < kWarps * kBins; i += kThreads) (&wcount[0][0])[i] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const unsigned long long wbase = (unsigned long long)blockIdx.x * kTile + (unsigned long long)w * (32 * kItems);
  const uint32_t mask = (1u << ps.bits) - 1u;
  const unsigned lt = (1u << lane) - 1u;
  uint4 r[kItems];
  uint32_t pos[kItems];  // digit | rank among the warp's earlier records with that digit << 8
#pragma unroll
  for (int i = 0; i < kItems; ++i) {
    const unsigned long long idx = wbase + (unsigned long long)i * 32 + lane;
    const bool valid = idx < n;
    uint32_t d = 0x100u + lane;  // lanes past the end match nobody
    if (valid) {
      r[i] = in[idx];
      d = (key(r[i], ps.word) >> ps.shift) & mask;
    }
    const unsigned peers = __match_any_sync(0xffffffffu, d);
    const int leader = __ffs(peers) - 1;
    uint32_t prev = 0;
    if (lane == leader && valid) {
      prev = wcount[w][d];
      wcount[w][d] = prev + __popc(peers);
    }
    prev = __shfl_sync(0xffffffffu, prev, leader);
    pos[i] = valid ? (d | ((prev + __popc(peers & lt)) << 8)) : 0xffffffffu;
    __syncwarp();
  }
  __syncthreads();
  {  // this tile's first output index per digit, then the warps in order
    const int d = threadIdx.x;
    uint32_t run = tile_ofs[(size_t)d * n_tiles + blockIdx.x];
#pragma unroll
    for (int ww = 0; ww < kWarps; ++ww) {
      uint32_t c = wcount[ww][d];
      wcount[ww][d] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < kItems; ++i)
    if (pos[i] != 0xffffffffu) out[wcount[w][pos[i] & 0xffu] + (pos[i] >> 8)] = r[i];
}

// ---- exclusive scan of a long uint32 array (total < 2^32) -------------------------------------
constexpr int kScanBlock = 1024;
constexpr int kScanPer = 4;  // elements per thread: 4096 per CTA

__global__ void __launch_bounds__(kScanBlock) k_scan_tiles(uint32_t* data, unsigned long long n, uint32_t* sums) {
  __shared__ uint32_t warp_tot[32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const unsigned long long i0 = ((unsigned long long)blockIdx.x * kScanBlock + threadIdx.x) * kScanPer;
  uint32_t v[kScanPer];
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < kScanPer; ++k) {
    v[k] = (i0 + k < n) ? data[i0 + k] : 0u;
    s += v[k];
  }
  uint32_t inc = s;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += o;
  }
  if (lane == 31) warp_tot[w] = inc;
  __syncthreads();
  if (w == 0) {
    uint32_t t = warp_tot[lane], ti = t;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t o = __shfl_up_sync(0xffffffffu, ti, d);
      if (lane >= d) ti += o;
    }
    warp_tot[lane] = ti - t;
    if (lane == 31) sums[blockIdx.x] = ti;
  }
  __syncthreads();
  uint32_t ex = warp_tot[w] + inc - s;
#pragma unroll
  for (int k = 0; k < kScanPer; ++k) {
    if (i0 + k < n) data[i0 + k] = ex;
    ex += v[k];
  }
}

__global__ void __launch_bounds__(kScanBlock) k_scan_add(uint32_t* data, unsigned long long n, const uint32_t* sums) {
  const uint32_t off = sums[blockIdx.x];
  const unsigned long long i0 = ((unsigned long long)blockIdx.x * kScanBlock + threadIdx.x) * kScanPer;
#pragma unroll
  for (int k = 0; k < kScanPer; ++k)
    if (i0 + k < n) data[i0 + k] += off;
}

// Exclusive scan in place.  `scratch` holds the tile totals of every level: at least
// scan_scratch_elems(n) uint32.  Recursion depth is 3 for any n < 2^36.
inline unsigned long long scan_scratch_elems(unsigned long long n) {
  unsigned long long tot = 0;
  while (n > 1) {
    n = (n + kScanBlock * kScanPer - 1) / (kScanBlock * kScanPer);
    tot += n;
  }
  return tot + 1;
}

inline cudaError_t scan_u32(uint32_t* data, unsigned long long n, uint32_t* scratch, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  const unsigned long long tiles = (n + kScanBlock * kScanPer - 1) / (kScanBlock * kScanPer);
  k_scan_tiles<<<(unsigned)tiles, kScanBlock, 0, s>>>(data, n, scratch);
  if (tiles > 1) {
    cudaError_t e = scan_u32(scratch, tiles, scratch + tiles, s);
    if (e != cudaSuccess) return e;
    k_scan_add<<<(unsigned)tiles, kScanBlock, 0, s>>>(data, n, scratch);
  }
  return cudaGetLastError();
}

// Split a key word with `bits` significant bits into digits of at most 8 bits, evenly.
inline int plan_word(Pass* out, int word, int bits) {
  if (bits <= 0) return 0;
  const int np = (bits + 7) / 8;
  int shift = 0;
  for (int i = 0; i < np; ++i) {
    int b = (bits - shift + (np - i) - 1) / (np - i);
    out[i] = Pass{word, shift, b};
    shift += b;
  }
  return np;
}

inline size_t workspace_bytes(unsigned long long n) {
  const unsigned long long tiles = (n + kTile - 1) / kTile;
  return sizeof(uint32_t) * ((size_t)tiles * kBins + (size_t)scan_scratch_elems(tiles * kBins));
}

// Sort `n` records by the key words of `passes` (least significant first).  `a` holds the input,
// `b` is a buffer of the same size; returns (through *result) which of the two holds the output.
template <typename KeyFn>
inline cudaError_t sort(uint4* a, uint4* b, unsigned long long n, const Pass* passes, int n_passes, uint32_t* workspace,
                        cudaStream_t s, KeyFn key, uint4** result, unsigned long long* launches = nullptr) {
  *result = a;
  if (n < 2) return cudaSuccess;
  const unsigned long long tiles = (n + kTile - 1) / kTile;
  uint32_t* hist = workspace;
  uint32_t* scratch = workspace + tiles * kBins;
  uint4 *src = a, *dst = b;
  for (int p = 0; p < n_passes; ++p) {
    k_radix_hist<<<(unsigned)tiles, kThreads, 0, s>>>(src, n, passes[p], hist, (unsigned)tiles, key);
    cudaError_t e = scan_u32(hist, tiles * kBins, scratch, s);
    if (e != cudaSuccess) return e;
    k_radix_scatter<<<(unsigned)tiles, kThreads, 0, s>>>(src, dst, n, passes[p], hist, (unsigned)tiles, key);
    if (launches) *launches += 4;
    uint4* t = src; src = dst; dst = t;
  }
  *result = src;
  return cudaGetLastError();
}

}  // namespace rsort
