// citam_b200.cu — B200 (sm_100a) implementation of CITAM's per-timestep loop.
//
// What it replaces in the reference (paths relative to the upstream repo):
//   Agent.step + Schedule.get_next_position   citam/engine/core/agent.py:67-109,
//                                             citam/engine/schedulers/office_schedule.py:632-638
//   Floorplan.identify_this_location          citam/engine/map/floorplan.py:215-241 (+ space.py, geometry.py)
//   CloseContactsCalculator.run               citam/engine/calculators/close_contacts_calculator.py:64-231
//   ContactEvents.add_contact / statistics    citam/engine/calculators/contacts.py:92-294
//
// Pipeline per chunk of S <= 256 time steps (one launch per stage covers all steps of the chunk;
// every accumulator is commutative, so steps, chunks and time blocks are independent):
//   k_expand_bin      itinerary segments -> per-step 16-byte records {x, y, state word, rank}; one
//                     atomic per located agent returns its rank in its uniform-grid cell
//   k_scan(+_add)     per-step exclusive scan of the cell histogram (rows x pieces CTAs)
//   k_scatter         counting-sort scatter into cell order (16-byte records)
//   k_select_changed  per step, the agents whose state may differ from the step before, in cell order
//   k_pairs           first step of the chunk: every located agent against its half neighbourhood
//                     (own cell tail, right cell, 3 cells above), candidate window in shared memory
//   k_pairs_changed   other steps: only changed agents, 3x3 neighbourhood, tests flattened over a
//                     256-agent batch; a contact of two unchanged agents is inside a run that an
//                     earlier step added in one piece (stationary-run aggregation)
//   k_accumulate      (side stream, overlaps the next chunk) one contact-run record per thread:
//                     16-byte pair-table slot probe + one CAS, one 64-bit reduction into the dense
//                     per-coordinate histogram, two reductions into the L2-resident per-agent counters
// The distance predicate is the reference's float64 sqrt(dx^2+dy^2) < D as an exact integer
// threshold (find_s_max).  All of this is HBM/L2-bound integer work; there is no GEMM in the path.

#include "../../include/citam_b200.h"

#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace cg = cooperative_groups;

// --------------------------------------------------------------------------
// error plumbing
// --------------------------------------------------------------------------
static thread_local std::string g_err;

static int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CK(call)                                                                          \
  do {                                                                                    \
    cudaError_t _e = (call);                                                              \
    if (_e != cudaSuccess)                                                                \
      return fail(CITAM_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), \
                  __FILE__, __LINE__);                                                    \
  } while (0)

// --------------------------------------------------------------------------
// device-side data
// --------------------------------------------------------------------------
struct FloorDev {
  int32_t minx, miny, W, H;  // lattice of the location table
  int32_t ncx, ncy;          // uniform grid over the same lattice
  int32_t cell_base;         // first cell id of this floor
  int32_t n_spaces;
  const int16_t* lut;   // [H][W] space index, -1 = none; NULL = floor not loaded
  const uint32_t* adj;  // n_spaces*n_spaces bits, NULL = no hallway graph
  long long hist_base;  // first bucket of this floor in the dense coordinate histogram
  int32_t hist_w, hist_h;  // (2W-1) x (2H-1) half-unit buckets: contact positions are midpoints
};

// One slot of the per-pair table: 16 bytes, two per 32-byte sector.  `acc` packs everything the
// reference keeps per pair (contacts.py:155-183) so that one compare-and-swap on the sector the
// probe already fetched applies a whole update:
//   bits  0-23  contact-steps (sum of event durations)        < 2^24 = the step-range limit
//   bits 24-39  number of events                              < 2^16
//   bits 40-63  0xFFFFFF - first contact step (max-reduced; 0 = none yet)
struct __align__(16) PairEntry {
  unsigned long long key;  // (a1 << 32) | a2 with a1 < a2, 0 = empty
  unsigned long long acc;
};
static_assert(sizeof(PairEntry) == 16, "two pair entries per 32-byte sector");
#define ACC_DUR(a) ((uint32_t)((a) & 0xffffffULL))
#define ACC_NEV(a) ((uint32_t)(((a) >> 24) & 0xffffULL))
#define ACC_FIRST_INV(a) ((uint32_t)((a) >> 40))
#define NO_STEP 0xFFFFFFFFu

struct __align__(16) CoordEntry {
  unsigned long long key;  // bit63 | floor<<56 | (sx+2^27)<<28 | (sy+2^27); 0 = empty
  unsigned long long count;
};

// One detected contact run, appended by k_pairs and applied by k_accumulate.
struct __align__(16) ContactRec {
  uint32_t a1, a2;   // a1 < a2
  uint32_t t_run;    // step (24 bits) | (run - 1) << 24
  uint32_t bucket;   // dense-histogram bucket (30 bits) | bit 31: in contact at t-1 | bit 30: evaluate that at accumulation
};

constexpr int kSubBuffers = 64;    // the contact buffer is split so that appends do not all hit one counter
constexpr int kCounterStride = 16;  // 128 bytes between sub-buffer counters

struct DevCounters {
  unsigned long long pair_tests;
  unsigned long long contact_steps;
  unsigned long long pair_used;
  unsigned long long coord_used;
  unsigned long long log_count;
  unsigned long long overflow;
  unsigned long long export_cursor;
  unsigned long long stats[6];
  unsigned long long contact_runs;
  unsigned long long pad[2];
  unsigned long long cbuf_count[2][kSubBuffers * kCounterStride];  // one set per contact buffer
  unsigned long long pair_used_sub[kSubBuffers * kCounterStride];  // spread: inserts are frequent
};

struct Params {
  int32_t N, F;
  int32_t cell;
  int32_t total_cells;
  int32_t cells_stride;  // uint32 per step row of the histogram (>= total_cells+1, multiple of 4)
  double D;
  long long s_max;  // largest integer s with sqrt_rn((double)s) < D  (-1: none)
  const FloorDev* floors;
  const int4* segs;  // {t0 | floor<<24, x0, y0, dx&0xffff | dy<<16}
  const long long* seg_off;
  const int2* window;  // per agent {first step in the facility, first step of a final stay outside every space}
                       // (INT_MAX when there is none); NULL = not available
  int4* rec;        // [(S+1)][N] {x, y, state word, rank in cell}; row 0 = halo step t_begin-1
  uint32_t* cells;  // [S][cells_stride]
  int4* sorted;     // [S][N] {x, y, agent, state word} in cell order
  PairEntry* pairs;
  unsigned long long pair_mask;
  CoordEntry* coords;     // hashed histogram (lattices too large for the dense one)
  unsigned long long coord_mask;
  unsigned long long* hist;  // dense histogram, NULL = use the hash
  int32_t range_begin;       // first step of the current citam_run range
  int32_t chunk_begin;       // first step of the chunk being processed
  citam_contact* log;
  unsigned long long log_cap;
  ContactRec* cbuf;  // per-chunk contact buffer, kSubBuffers equal parts (NULL: accumulate inline)
  unsigned long long cbuf_sub_cap;
  unsigned long long* cbuf_count;  // its kSubBuffers counters (stride kCounterStride)
  unsigned long long* agent_dur;   // [N] contact-steps per agent
  uint32_t* chg_list;   // [S][N] located agents whose state may differ from the step before
  uint32_t* chg_count;  // [S] x 32 (one counter per 128 bytes); NULL = lists not kept
  DevCounters* ctr;
};

#define OVF_PAIR 1ULL
#define OVF_COORD 2ULL
#define OVF_LOG 4ULL
#define OVF_GRID 8ULL
#define OVF_FIELD 16ULL

// state word: bits 0-14 space index (0x7fff = none); bit 15 "changed": the agent is NOT known to
// be exactly where it was one step earlier (it moves, or this is the first step of its segment);
// bits 16-23: for a still agent "stay" = for how many following steps of this chunk the state is
// guaranteed identical (same zero-velocity segment); for a moving agent (bit 30) the step it
// just made, (dx+8) | (dy+8) << 4, or 0 when that is not known here (first step of a segment, a
// stride beyond +-7); bits 24-29 floor; bit 31 = not in the facility.  changed/stay come from the
// segment store alone, so they agree with each other: changed(t+k) == 0 for k = 1..stay(t) and
// changed(t+stay+1) == 1 (or the chunk ends).
constexpr int kMaxChunk = 256;  // stay must fit 8 bits
constexpr int kMaxFloors = 63;
__device__ __forceinline__ int32_t pack_fl(int floor, int loc, int changed = 1, int stay = 0, int moving = 0) {
  if (floor < 0) return (int32_t)0x8000ffffu;
  return (int32_t)(((uint32_t)(moving & 1) << 30) | ((uint32_t)(floor & 0x3f) << 24) | ((uint32_t)(stay & 0xff) << 16) |
                   ((uint32_t)(changed & 1) << 15) | ((uint32_t)loc & 0x7fffu));
}
__device__ __forceinline__ int fl_floor(int32_t fl) { return fl < 0 ? -1 : (fl >> 24) & 0x3f; }
__device__ __forceinline__ int fl_loc(int32_t fl) { int v = fl & 0x7fff; return v == 0x7fff ? -1 : v; }
__device__ __forceinline__ bool fl_changed(int32_t fl) { return (fl >> 15) & 1; }
__device__ __forceinline__ bool fl_moving(int32_t fl) { return (fl >> 30) & 1; }
__device__ __forceinline__ int fl_stay(int32_t fl) { return fl_moving(fl) ? 0 : (fl >> 16) & 0xff; }
__device__ __forceinline__ int fl_step_code(int32_t fl) { return fl_moving(fl) ? (fl >> 16) & 0xff : 0; }

__device__ __forceinline__ unsigned long long mix64(unsigned long long k) {
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdULL;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ULL;
  k ^= k >> 33;
  return k;
}

__device__ __forceinline__ int lut_lookup(const FloorDev& fd, int x, int y) {
  if (fd.lut == nullptr) return -1;
  int fx = x - fd.minx, fy = y - fd.miny;
  if ((unsigned)fx >= (unsigned)fd.W || (unsigned)fy >= (unsigned)fd.H) return -1;
  return (int)__ldg(fd.lut + (size_t)fy * fd.W + fx);
}

__device__ __forceinline__ int cell_of(const FloorDev& fd, int cell, int x, int y) {
  int cx = (x - fd.minx) / cell, cy = (y - fd.miny) / cell;
  return fd.cell_base + cy * fd.ncx + cx;
}

// Per agent, the steps at which it can matter to the loop: before `x` it is not in the facility;
// from `y` on it sits for ever at a point outside every space (the sticky tail of a finished
// itinerary).  Outside [x - 1, y) the step kernels neither write nor read the agent's records.
__global__ void k_agent_windows(Params p, int2* out) {
  int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= p.N) return;
  long long sb = p.seg_off[a], se = p.seg_off[a + 1];
  int2 w = make_int2(INT_MAX, INT_MAX);
  if (se > sb) {
    w.x = p.segs[sb].x & 0xffffff;
    int4 last = p.segs[se - 1];
    int f = (last.x >> 24) & 0xff;
    if (last.w == 0 && (f >= p.F || lut_lookup(p.floors[f], last.y, last.z) < 0)) w.y = last.x & 0xffffff;
  }
  out[a] = w;
}

// --------------------------------------------------------------------------
// K1a: itinerary expansion + grid histogram.
// One thread walks one agent through kRowsPerThread consecutive steps: the segment cursor
// advances monotonically, the location table is touched only when the position changes, and
// every 16-byte record store is coalesced over the 32 agents of a warp.
// --------------------------------------------------------------------------

__global__ void __launch_bounds__(128) k_expand_bin(Params p, int t_first, int rows, int do_bin, int kRowsPerThread) {
  int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= p.N) return;
  int r0 = blockIdx.y * kRowsPerThread;
  int2 win = make_int2(INT_MIN, INT_MAX);
  if (do_bin && p.window != nullptr) {
    win = p.window[a];
    // nothing of this thread's steps lies in [enter - 1, exit): no record is written (none is read)
    if (win.x > t_first + min(r0 + kRowsPerThread, rows) || win.y <= t_first + r0) return;
  }
  long long sb = p.seg_off[a], se = p.seg_off[a + 1];
  int t = t_first + r0;
  long long lo = sb, hi = se;  // first segment with t0 > t
  while (lo < hi) {
    long long mid = (lo + hi) >> 1;
    int t0 = __ldg(&p.segs[mid].x) & 0xffffff;
    if (t0 <= t) lo = mid + 1; else hi = mid;
  }
  long long c = lo - 1;
  int4 s = make_int4(0, 0, 0, 0);
  int next_t0 = INT_MAX;
  if (c >= sb) s = __ldg(&p.segs[c]);
  if (c + 1 < se) next_t0 = __ldg(&p.segs[c + 1].x) & 0xffffff;
  int lastx = INT_MIN, lasty = INT_MIN, lastf = -1, loc = -1;
#pragma unroll 1
  for (int r = 0; r < kRowsPerThread; ++r) {
    int row = r0 + r;
    if (row >= rows) break;
    t = t_first + row;
    while (t >= next_t0) {
      ++c;
      s = __ldg(&p.segs[c]);
      next_t0 = (c + 1 < se) ? (__ldg(&p.segs[c + 1].x) & 0xffffff) : INT_MAX;
    }
    size_t o = (size_t)row * p.N + a;
    if (t + 1 < win.x || t >= win.y) continue;  // outside the agent's window (see k_agent_windows)
    if (c < sb || t < 0) {
      p.rec[o] = make_int4(0, 0, pack_fl(-1, -1), 0);
      continue;
    }
    int dt = t - (s.x & 0xffffff);
    int f = (s.x >> 24) & 0xff;
    int x = s.y + (int)(int16_t)(s.w & 0xffff) * dt;
    int y = s.z + (int)(int16_t)((uint32_t)s.w >> 16) * dt;
    if (x != lastx || y != lasty || f != lastf) {
      loc = (f < p.F) ? lut_lookup(p.floors[f], x, y) : -1;
      lastx = x; lasty = y; lastf = f;
    }
    // stationary inside one segment: same state as the step before / for `stay` more steps
    const bool still = s.w == 0;
    int changed = !(still && dt >= 1);
    int stay = still ? min(min(next_t0 - 1 - t, rows - 1 - row), kMaxChunk - 1) : 0;
    if (!still) {  // the step just made, when it is this segment's stride
      int vx = (int)(int16_t)(s.w & 0xffff), vy = (int)(int16_t)((uint32_t)s.w >> 16);
      stay = (dt >= 1 && vx >= -7 && vx <= 7 && vy >= -7 && vy <= 7) ? ((vx + 8) | ((vy + 8) << 4)) : 0;
    }
    uint32_t rk = 0;
    if (do_bin && row >= 1 && loc >= 0) {
      int key = cell_of(p.floors[f], p.cell, x, y);
      rk = atomicAdd(&p.cells[(size_t)(row - 1) * p.cells_stride + key], 1u);
    }
    p.rec[o] = make_int4(x, y, pack_fl(f, loc, changed, stay, still ? 0 : 1), (int)rk);
  }
}

// K1b: histogram from caller-supplied state rows (the Calculator.run entry).
__global__ void k_bin_states(Params p, int rows) {
  int a = blockIdx.x * blockDim.x + threadIdx.x;
  int row = blockIdx.y + 1;
  if (a >= p.N || row >= rows) return;
  size_t o = (size_t)row * p.N + a;
  int4 r = p.rec[o];
  int f = fl_floor(r.z), loc = fl_loc(r.z);
  if (f < 0 || loc < 0) return;
  bool ok = f < p.F;
  if (ok) {
    const FloorDev& fd = p.floors[f];
    int fx = r.x - fd.minx, fy = r.y - fd.miny;
    ok = fd.lut != nullptr && (unsigned)fx < (unsigned)fd.W && (unsigned)fy < (unsigned)fd.H;
  }
  if (!ok) {  // caller says "located" but the point is outside the floor's grid
    atomicOr(&p.ctr->overflow, OVF_GRID);
    p.rec[o].z = pack_fl(f, -1);
    return;
  }
  int key = cell_of(p.floors[f], p.cell, r.x, r.y);
  p.rec[o].w = (int)atomicAdd(&p.cells[(size_t)(row - 1) * p.cells_stride + key], 1u);
}

__global__ void k_states_to_rec(const citam_agent_state* st, int N, int4* rec) {
  int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= N) return;
  citam_agent_state s = st[a];
  int f = s.floor < 0 ? -1 : s.floor;
  int l = (f < 0 || s.loc < 0) ? -1 : s.loc;
  rec[a] = make_int4(s.x, s.y, pack_fl(f, l), 0);
}

__global__ void k_rec_to_states(const int4* rec, long long N, citam_agent_state* st) {
  long long a = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= N) return;
  int4 r = rec[a];
  citam_agent_state s;
  s.x = r.x; s.y = r.y; s.floor = (int16_t)fl_floor(r.z); s.loc = (int16_t)fl_loc(r.z);
  st[a] = s;
}

__global__ void k_fill_inactive(int4* rec, int N) {
  int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= N) return;
  rec[a] = make_int4(0, 0, pack_fl(-1, -1), 0);
}

// --------------------------------------------------------------------------
// K2: per-step exclusive scan of the cell histogram, in place.
// cells[row][0..total_cells) counts -> starts; cells[row][total_cells] = located agents.
// A row is cut into `parts` pieces so that rows x parts CTAs fill the machine: k_scan scans its
// piece (warp shuffles over 4 cells per thread, warp totals through shared memory) and leaves
// the piece total in part_sums; k_scan_add adds the totals of the pieces before it.
// --------------------------------------------------------------------------
constexpr int kScanThreads = 1024;

__global__ void __launch_bounds__(kScanThreads) k_scan(uint32_t* cells, int stride, int n /* total_cells+1 */,
                                                       int part_len4, uint32_t* part_sums) {
  __shared__ uint32_t warp_off[32];
  __shared__ uint32_t s_tot;
  uint4* c4 = reinterpret_cast<uint4*>(cells + (size_t)blockIdx.y * stride);
  const int n4 = (n + 3) >> 2;  // stride is a multiple of 4 and the padding is zero
  const int begin = blockIdx.x * part_len4, end = min(n4, begin + part_len4);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint32_t carry = 0;
  for (int base = begin; base < end; base += kScanThreads) {
    int i = base + threadIdx.x;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (i < end) v = c4[i];
    uint32_t sum = v.x + v.y + v.z + v.w;
    uint32_t inc = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += o;
    }
    if (lane == 31) warp_off[w] = inc;
    __syncthreads();
    if (w == 0) {
      uint32_t wt = warp_off[lane];
      uint32_t winc = wt;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        uint32_t o = __shfl_up_sync(0xffffffffu, winc, d);
        if (lane >= d) winc += o;
      }
      warp_off[lane] = winc - wt;
      if (lane == 31) s_tot = winc;
    }
    __syncthreads();
    uint32_t ex = carry + warp_off[w] + (inc - sum);
    if (i < end) {
      uint4 o;
      o.x = ex;
      o.y = ex + v.x;
      o.z = o.y + v.y;
      o.w = o.z + v.z;
      c4[i] = o;
    }
    carry += s_tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) part_sums[blockIdx.y * gridDim.x + blockIdx.x] = carry;
}

__global__ void __launch_bounds__(256) k_scan_add(uint32_t* cells, int stride, int n, int part_len4,
                                                  const uint32_t* part_sums) {
  const int part = blockIdx.x + 1;  // piece 0 needs no offset
  uint32_t off = 0;
  for (int q = 0; q < part; ++q) off += part_sums[blockIdx.y * (gridDim.x + 1) + q];
  if (off == 0) return;
  uint4* c4 = reinterpret_cast<uint4*>(cells + (size_t)blockIdx.y * stride);
  const int n4 = (n + 3) >> 2;
  const int begin = part * part_len4, end = min(n4, begin + part_len4);
  for (int i = begin + threadIdx.x; i < end; i += blockDim.x) {
    uint4 v = c4[i];
    v.x += off; v.y += off; v.z += off; v.w += off;
    c4[i] = v;
  }
}

// --------------------------------------------------------------------------
// K3: counting-sort scatter into cell order.
// --------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_scatter(Params p, int rows /* S */, int t_begin) {
  int a = blockIdx.x * blockDim.x + threadIdx.x;
  int row = blockIdx.y;
  if (a >= p.N || row >= rows) return;
  if (p.window != nullptr) {  // 8 bytes instead of the 16-byte record for agents that are not around
    const int2 win = p.window[a];
    const int t = t_begin + row;
    if (t < win.x || t >= win.y) return;
  }
  int4 r = p.rec[(size_t)(row + 1) * p.N + a];
  int f = fl_floor(r.z), loc = fl_loc(r.z);
  if (f < 0 || loc < 0) return;
  int key = cell_of(p.floors[f], p.cell, r.x, r.y);
  uint32_t dest = p.cells[(size_t)row * p.cells_stride + key] + (uint32_t)r.w;
  p.sorted[(size_t)row * p.N + dest] = make_int4(r.x, r.y, a, r.z);
}

// K3b: per step, the sorted positions of the agents whose state may have changed, compacted in
// CELL ORDER (each warp appends its 32 consecutive sorted positions' hits in one piece), so that
// neighbouring threads of k_pairs_changed work on neighbouring agents: their candidate ranges
// overlap (L1 hits) and their histogram buckets share sectors.
__global__ void __launch_bounds__(256) k_select_changed(Params p, int rows /* S */) {
  const int row = blockIdx.y + 1;  // the first step of a chunk is handled by k_pairs in full
  if (row >= rows) return;
  const int total = (int)p.cells[(size_t)row * p.cells_stride + p.total_cells];
  const int lane = threadIdx.x & 31;
  for (int q0 = blockIdx.x * blockDim.x; q0 < total; q0 += gridDim.x * blockDim.x) {
    const int q = q0 + threadIdx.x;
    bool sel = false;
    if (q < total) sel = fl_changed(p.sorted[(size_t)row * p.N + q].w);
    unsigned m = __ballot_sync(0xffffffffu, sel);
    if (m == 0) continue;
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(&p.chg_count[row * 32], (uint32_t)__popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (sel) p.chg_list[(size_t)row * p.N + base + __popc(m & ((1u << lane) - 1))] = (uint32_t)q;
  }
}

// --------------------------------------------------------------------------
// K4: pair test + accumulation.
// --------------------------------------------------------------------------
__device__ __forceinline__ bool loc_rule(const FloorDev& fd, int l1, int l2) {
  // close_contacts_calculator.py:111-124,173-203: same space, or both hallways with a graph edge
  if (l1 == l2) return true;
  if (fd.adj == nullptr) return false;
  unsigned bit = (unsigned)l1 * (unsigned)fd.n_spaces + (unsigned)l2;
  return (__ldg(fd.adj + (bit >> 5)) >> (bit & 31)) & 1u;
}

// has_edge(space_of_lower_id, space_of_higher_id): argument order as in the reference, which
// matters only if a directed hallway graph is ever supplied
__device__ __forceinline__ bool loc_rule_ordered(const FloorDev& fd, int ida, int la, int idb, int lb) {
  return ida < idb ? loc_rule(fd, la, lb) : loc_rule(fd, lb, la);
}

// scipy pdist (euclidean) on integer coordinates: sqrt(dx*dx+dy*dy) in float64, then the strict
// `< contact_distance` of np.argwhere (close_contacts_calculator.py:169-171).  dx, dy are exact
// integers, so s = dx^2+dy^2 is exact, and s -> sqrt_rn((double)s) is monotone: the predicate
// holds exactly for s <= s_max, where the host found s_max with the same IEEE operations
// (find_s_max) and citam_create checked it on the device against __dsqrt_rn (k_check_s_max).
__device__ __forceinline__ bool within(int dx, int dy, const Params& p) {
  long long s = (long long)dx * dx + (long long)dy * dy;
  return s <= p.s_max;
}

// Linear probing is bounded: a table that needs more probes than this is reported as full
// (CITAM_ERR_CAPACITY) instead of degrading into a scan of the whole table.
constexpr int kMaxProbes = 256;

__device__ __forceinline__ ulonglong2 load_entry(const PairEntry* e) {
  ulonglong2 v;
  asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(e));
  return v;
}

// New pairs are counted per sub-counter and per converged group of lanes: with transient
// contacts most updates insert, and a single counter would serialise them at the L2.
__device__ __forceinline__ void count_insert(const Params& p) {
  cg::coalesced_group g = cg::coalesced_threads();
  if (g.thread_rank() == 0)
    atomicAdd(&p.ctr->pair_used_sub[((blockIdx.x + blockIdx.y) & (kSubBuffers - 1)) * kCounterStride],
              (unsigned long long)g.size());
}

// Find or insert the pair's slot.  Returns the slot index (or ~0 when the table is full) and the
// accumulator word last seen there, fetched with the key in one 16-byte load.
__device__ __forceinline__ unsigned long long pair_slot(const Params& p, unsigned long long key,
                                                        unsigned long long* acc_seen) {
  unsigned long long h = mix64(key) & p.pair_mask;
  for (int probe = 0; probe < kMaxProbes; ++probe) {
    ulonglong2 e = load_entry(&p.pairs[h]);
    if (e.x == key) { *acc_seen = e.y; return h; }
    if (e.x == 0ULL) {
      unsigned long long old = atomicCAS(&p.pairs[h].key, 0ULL, key);
      if (old == 0ULL) { count_insert(p); *acc_seen = 0ULL; return h; }
      if (old == key) { *acc_seen = *(volatile unsigned long long*)&p.pairs[h].acc; return h; }
    }
    h = (h + 1) & p.pair_mask;
  }
  atomicOr(&p.ctr->overflow, OVF_PAIR);
  return ~0ULL;
}

__device__ __forceinline__ void pair_apply(const Params& p, unsigned long long h, unsigned long long old, uint32_t first,
                                           unsigned long long add_dur, unsigned long long add_nev, uint32_t a1, uint32_t a2);
// add_dur contact-steps and add_nev events; `first` (a step, or NO_STEP) lowers the pair's
// first-contact step.  The first contact of a pair is always an event start, so callers pass a
// step only for event starts and for the first step of a run range.
__device__ __forceinline__ void pair_upsert(const Params& p, uint32_t a1, uint32_t a2, uint32_t first,
                                            unsigned long long add_dur, unsigned long long add_nev) {
  unsigned long long key = ((unsigned long long)a1 << 32) | a2;
  unsigned long long old;
  unsigned long long h = pair_slot(p, key, &old);
  if (h == ~0ULL) return;
  pair_apply(p, h, old, first, add_dur, add_nev, a1, a2);
}

__device__ __forceinline__ void coord_upsert(const Params& p, int floor, int sx, int sy, unsigned long long add) {
  if (p.hist != nullptr) {  // dense: one reduction, no probe
    const FloorDev& fd = p.floors[floor];
    long long ix = (long long)sx - 2LL * fd.minx, iy = (long long)sy - 2LL * fd.miny;
    atomicAdd(p.hist + fd.hist_base + iy * fd.hist_w + ix, add);
    return;
  }
  unsigned long long key = (1ULL << 63) | ((unsigned long long)(floor & 0x7f) << 56) |
                           ((unsigned long long)((uint32_t)(sx + (1 << 27)) & 0xfffffffu) << 28) |
                           (unsigned long long)((uint32_t)(sy + (1 << 27)) & 0xfffffffu);
  unsigned long long h = mix64(key) & p.coord_mask;
  bool found = false;
  for (int probe = 0; probe < kMaxProbes; ++probe) {
    unsigned long long k = *(volatile unsigned long long*)&p.coords[h].key;
    if (k == key) { found = true; break; }
    if (k == 0ULL) {
      unsigned long long old = atomicCAS(&p.coords[h].key, 0ULL, key);
      if (old == 0ULL) { atomicAdd(&p.ctr->coord_used, 1ULL); found = true; break; }
      if (old == key) { found = true; break; }
    }
    h = (h + 1) & p.coord_mask;
  }
  if (!found) { atomicOr(&p.ctr->overflow, OVF_COORD); return; }
  atomicAdd(&p.coords[h].count, add);
}

// Was the pair in contact at the previous step?  Full predicate on the halo/previous row
// (contacts.py:131-133: an event continues iff last.start_step + last.duration == step).
// One 16-byte record per agent: two sector reads.
__device__ __forceinline__ bool contact_eval(const Params& p, const int4& r1, const int4& r2) {
  int fl1 = fl_floor(r1.z), fl2 = fl_floor(r2.z);
  int l1 = fl_loc(r1.z), l2 = fl_loc(r2.z);
  if (fl1 < 0 || fl1 != fl2 || l1 < 0 || l2 < 0) return false;
  long long dx = (long long)r2.x - r1.x, dy = (long long)r2.y - r1.y;
  long long lim = (long long)p.D + 1;
  if (dx > lim || dx < -lim || dy > lim || dy < -lim) return false;
  if (!within((int)dx, (int)dy, p)) return false;
  return loc_rule(p.floors[fl1], l1, l2);
}
__device__ __forceinline__ bool contact_prev(const Params& p, size_t prow_off, uint32_t a1, uint32_t a2) {
  return contact_eval(p, p.rec[prow_off + a1], p.rec[prow_off + a2]);
}

// pair_upsert split in two so that a caller can put independent loads between probe and update.
// Also adds the contact-steps to both agents' counters (close_contacts_calculator.py:221-224): the
// counter array is N x 8 bytes, L2-resident, so these two reductions cost no DRAM access.
__device__ __forceinline__ void pair_apply(const Params& p, unsigned long long h, unsigned long long old, uint32_t first,
                                           unsigned long long add_dur, unsigned long long add_nev, uint32_t a1, uint32_t a2) {
  atomicAdd(&p.agent_dur[a1], add_dur);
  atomicAdd(&p.agent_dur[a2], add_dur);
  unsigned long long finv = first == NO_STEP ? 0ULL : (unsigned long long)(0xFFFFFFu - (first & 0xFFFFFFu));
  if (finv <= ACC_FIRST_INV(old)) {
    // the first-contact field cannot move (it only grows, and ours is not larger): the update is a
    // plain addition on the packed word -- one reduction, no round trip, no retry
    if (ACC_DUR(old) + add_dur > 0xffffffULL || ACC_NEV(old) + add_nev > 0xffffULL) { atomicOr(&p.ctr->overflow, OVF_FIELD); return; }
    atomicAdd(&p.pairs[h].acc, add_dur | (add_nev << 24));
    return;
  }
  for (;;) {
    unsigned long long d = ACC_DUR(old) + add_dur, n = ACC_NEV(old) + add_nev;
    if (d > 0xffffffULL || n > 0xffffULL) { atomicOr(&p.ctr->overflow, OVF_FIELD); return; }
    unsigned long long f = max((unsigned long long)ACC_FIRST_INV(old), finv);
    unsigned long long seen = atomicCAS(&p.pairs[h].acc, old, d | (n << 24) | (f << 40));
    if (seen == old) return;
    old = seen;
  }
}

// A contact found at step t that lasts `run` steps (both agents provably stay put for run-1
// more steps of this chunk): one table update for the whole run.  `was`: the pair was already in
// contact at t-1 (the event continues, contacts.py:131-133); `late`: not decided yet -- evaluate the
// predicate on the previous record row when the record is applied.
__device__ __forceinline__ void on_contact(const Params& p, uint32_t t, const int4& me, const int4& ot, uint32_t run,
                                           bool was, int sub, bool late = false) {
  uint32_t a1 = (uint32_t)min(me.z, ot.z), a2 = (uint32_t)max(me.z, ot.z);
  const int floor = fl_floor(me.w);
  if (p.cbuf != nullptr) {
    // warp-aggregated append into the chunk's contact buffer; k_accumulate applies it
    const FloorDev& fd = p.floors[floor];
    uint32_t ix = (uint32_t)((me.x + ot.x) - 2 * fd.minx), iy = (uint32_t)((me.y + ot.y) - 2 * fd.miny);
    uint32_t bucket = ((uint32_t)fd.hist_base + iy * (uint32_t)fd.hist_w + ix) | (was ? 0x80000000u : 0u);
    cg::coalesced_group g = cg::coalesced_threads();
    unsigned long long base = 0;
    if (g.thread_rank() == 0)
      base = atomicAdd(&p.cbuf_count[sub * kCounterStride], (unsigned long long)g.size());
    base = g.shfl(base, 0) + g.thread_rank();
    if (base < p.cbuf_sub_cap) {
      ContactRec r;
      r.a1 = a1; r.a2 = a2; r.t_run = (t & 0xffffffu) | ((run - 1u) << 24);
      r.bucket = late ? (bucket & 0x3fffffffu) | 0x40000000u : bucket;  // bit 30: evaluate `was` in k_accumulate
      p.cbuf[(unsigned long long)sub * p.cbuf_sub_cap + base] = r;
      return;
    }
    // sub-buffer full: fall through and accumulate inline (k_accumulate clamps the count)
  }
  if (late) was = contact_prev(p, (size_t)((int)t - p.chunk_begin) * p.N, a1, a2);
  uint32_t first = (!was || (int32_t)t == p.range_begin) ? t : NO_STEP;
  pair_upsert(p, a1, a2, first, run, was ? 0u : 1u);
  coord_upsert(p, floor, me.x + ot.x, me.y + ot.y, (unsigned long long)run);
  if (p.log_cap) {
    // warp-aggregated append: one atomic per converged group of lanes
    cg::coalesced_group g = cg::coalesced_threads();
    unsigned long long base = 0;
    if (g.thread_rank() == 0) base = atomicAdd(&p.ctr->log_count, (unsigned long long)g.size());
    base = g.shfl(base, 0) + g.thread_rank();
    if (base < p.log_cap) {
      citam_contact c; c.step = t; c.a1 = a1; c.a2 = a2;
      p.log[base] = c;
    } else {
      atomicOr(&p.ctr->overflow, OVF_LOG);
    }
  }
}

// One thread per located agent of one step ("me", in cell order) walks its half neighbourhood:
// the rest of its cell, the cell to the right and the three cells of the row above.  In cell
// order all of a block's candidates form ONE contiguous window of the sorted array, which the
// block stages in shared memory with coalesced 16-byte loads, so the per-candidate loop runs at
// shared-memory latency (windows beyond kWindow records are read from global memory instead).
//
// Stationary-run aggregation: a pair whose two agents are both provably where they were one step
// earlier (changed == 0) was handled at the step its run started -- by the owner step's thread,
// which added the whole run (1 + min(stay_me, stay_ot) steps, clipped to the chunk) in one
// update -- so later steps of the run skip it without a distance evaluation.  The first step of
// a chunk owns everything it finds.  With the contact log on, every contact-step is evaluated
// and logged on its own (aggregate == 0).
constexpr int kPairThreads = 256;
constexpr int kWindow = 2560;  // records (40 KB)

__global__ void __launch_bounds__(kPairThreads) k_pairs(Params p, int t_begin, int rows /* S */, int aggregate) {
  __shared__ int4 win[kWindow];
  __shared__ int s_w1;
  const int row = blockIdx.y;
  const uint32_t* cs = p.cells + (size_t)row * p.cells_stride;
  const int total = (int)cs[p.total_cells];
  const int w0 = blockIdx.x * kPairThreads;
  if (w0 >= total) return;  // whole block
  if (*(volatile unsigned long long*)&p.ctr->overflow & (OVF_PAIR | OVF_COORD | OVF_FIELD)) return;  // results are void already
  const int q0 = w0 + threadIdx.x;
  const int4* srt = p.sorted + (size_t)row * p.N;
  if (threadIdx.x == 0) s_w1 = w0;
  __syncthreads();
  int4 me = make_int4(0, 0, 0, 0);
  int qa = 0, qa_end = 0, qb = 0, qb_end = 0, f = 0, loc = -1;
  if (q0 < total) {
    me = srt[q0];
    f = fl_floor(me.w); loc = fl_loc(me.w);
    const FloorDev& fd = p.floors[f];
    int cx = (me.x - fd.minx) / p.cell, cy = (me.y - fd.miny) / p.cell;
    int rowbase = fd.cell_base + cy * fd.ncx;
    qa = q0 + 1;
    qa_end = (int)cs[rowbase + min(cx + 1, fd.ncx - 1) + 1];
    int hi = qa_end;
    if (cy + 1 < fd.ncy) {
      int rb2 = rowbase + fd.ncx;
      qb = (int)cs[rb2 + max(cx - 1, 0)];
      qb_end = (int)cs[rb2 + min(cx + 1, fd.ncx - 1) + 1];
      hi = max(hi, qb_end);
    }
    atomicMax(&s_w1, hi);
  }
  __syncthreads();
  const int wlen = s_w1 - w0;
  const bool staged = wlen <= kWindow;
  if (staged) {
    for (int i = threadIdx.x; i < wlen; i += kPairThreads) win[i] = srt[w0 + i];
    __syncthreads();
  }
  unsigned long long tests = 0, hits = 0, runs = 0;
  if (q0 < total) {
    const FloorDev& fd = p.floors[f];
    const uint32_t t = (uint32_t)(t_begin + row);
    const bool me_still = aggregate && row > 0 && !fl_changed(me.w);
    const int me_stay = aggregate ? fl_stay(me.w) : 0;
    const int sub = (blockIdx.x + blockIdx.y) & (kSubBuffers - 1);
#pragma unroll 1
    for (int rg = 0; rg < 2; ++rg) {
      const int qs = rg ? qb : qa, qe = rg ? qb_end : qa_end;
      for (int q = qs; q < qe; ++q) {
        int4 ot = staged ? win[q - w0] : srt[q];
        if (me_still && !fl_changed(ot.w)) continue;  // inside a run owned by an earlier step
        ++tests;
        if (within(ot.x - me.x, ot.y - me.y, p) && loc_rule_ordered(fd, me.z, loc, ot.z, fl_loc(ot.w))) {
          uint32_t run = 1u + (uint32_t)min(me_stay, aggregate ? fl_stay(ot.w) : 0);
          hits += run;
          ++runs;
          // both agents exactly where (and what) they were a step ago: in contact then too;
          // otherwise the predicate on the previous record row is evaluated at accumulation
          const bool known = !fl_changed(me.w) && !fl_changed(ot.w);
          on_contact(p, t, me, ot, run, known, sub, !known);
        }
      }
    }
  }
  // one counter update per warp
  for (int d = 16; d > 0; d >>= 1) {
    tests += __shfl_down_sync(0xffffffffu, tests, d);
    hits += __shfl_down_sync(0xffffffffu, hits, d);
    runs += __shfl_down_sync(0xffffffffu, runs, d);
  }
  if ((threadIdx.x & 31) == 0 && (tests | hits)) {
    atomicAdd(&p.ctr->pair_tests, tests);
    if (hits) { atomicAdd(&p.ctr->contact_steps, hits); atomicAdd(&p.ctr->contact_runs, runs); }
  }
}

// Steps after the first of a chunk: only agents whose state may have changed can start, end or
// move a contact (everything else is inside a run that an earlier step already added).  A block
// takes a batch of 256 changed agents (in cell order, from the per-step list of k_select_changed);
// every agent's 3x3 neighbourhood is three contiguous ranges of the sorted array.  The tests of the
// whole batch are flattened into one index space (block scan of the candidate counts, binary search
// in shared memory), so all 32 lanes of a warp test a candidate each, whatever the neighbour counts
// are, and consecutive lanes read consecutive sorted records.  A pair of two changed agents is
// owned by the lower id.  Hits are parked in shared memory and appended to the contact buffer once
// per warp and batch (one atomic for the warp, contiguous 16-byte stores).  Whether the pair was in
// contact one step earlier is left to k_accumulate (bit 30 of the record), where the two record
// reads overlap the table probe.  (Measured alternatives, config 3, ms per 12 blocks
// pairs/accumulate: one thread per agent 47/64; deciding the previous-step predicate here from the
// packed stride 68/59; four candidates per iteration 92/70; eight lanes per agent 126/60.)
constexpr int kHitSlots = 6;
constexpr int kPC = 256;

__global__ void __launch_bounds__(kPC) k_pairs_changed(Params p, int t_begin, int rows /* S */) {
  __shared__ ContactRec hitbuf[kPC * kHitSlots];
  __shared__ int4 s_me[kPC];
  __shared__ int s_qs[3][kPC];
  __shared__ int s_l0[kPC], s_l01[kPC];
  __shared__ uint32_t s_hb[kPC];
  __shared__ int s_off[kPC + 1];
  __shared__ int s_warp[kPC / 32];
  const int row = blockIdx.y + 1;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const uint32_t n = p.chg_count[row * 32];
  if (blockIdx.x * kPC >= n) return;
  if (*(volatile unsigned long long*)&p.ctr->overflow & (OVF_PAIR | OVF_COORD | OVF_FIELD)) return;
  const uint32_t* cs = p.cells + (size_t)row * p.cells_stride;
  const int4* srt = p.sorted + (size_t)row * p.N;
  const uint32_t t = (uint32_t)(t_begin + row);
  const int sub = (blockIdx.x + blockIdx.y) & (kSubBuffers - 1);
  const bool buffered = p.cbuf != nullptr;
  ContactRec* mine = hitbuf + threadIdx.x * kHitSlots;
  unsigned long long tests = 0, hits = 0, runs = 0;
  for (uint32_t base = blockIdx.x * kPC; base < n; base += gridDim.x * kPC) {
    // ---- phase 1: one thread per agent describes its candidate ranges
    const uint32_t idx = base + threadIdx.x;
    int total = 0;
    if (idx < n) {
      const int4 me = srt[p.chg_list[(size_t)row * p.N + idx]];
      const FloorDev& fd = p.floors[fl_floor(me.w)];
      const int cx = (me.x - fd.minx) / p.cell, cy = (me.y - fd.miny) / p.cell;
      const int x0 = max(cx - 1, 0), x1 = min(cx + 1, fd.ncx - 1);
      const int rb = fd.cell_base + cy * fd.ncx;
      const bool up = cy + 1 < fd.ncy, dn = cy > 0;
      const int rbu = up ? rb + fd.ncx : rb, rbd = dn ? rb - fd.ncx : rb;
      const int q0 = (int)cs[rbd + x0], l0 = dn ? (int)cs[rbd + x1 + 1] - q0 : 0;
      const int q1 = (int)cs[rb + x0], l1 = (int)cs[rb + x1 + 1] - q1;
      const int q2 = (int)cs[rbu + x0], l2 = up ? (int)cs[rbu + x1 + 1] - q2 : 0;
      s_me[threadIdx.x] = me;
      s_qs[0][threadIdx.x] = q0; s_qs[1][threadIdx.x] = q1; s_qs[2][threadIdx.x] = q2;
      s_l0[threadIdx.x] = l0; s_l01[threadIdx.x] = l0 + l1;
      s_hb[threadIdx.x] = (uint32_t)fd.hist_base - (uint32_t)(2 * fd.miny) * (uint32_t)fd.hist_w - (uint32_t)(2 * fd.minx);
      total = l0 + l1 + l2;
    }
    // block exclusive scan of the candidate counts
    int inc = total;
    for (int d = 1; d < 32; d <<= 1) {
      int v = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += v;
    }
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < wid; ++w) woff += s_warp[w];
    s_off[threadIdx.x] = woff + inc - total;
    if (threadIdx.x == kPC - 1) s_off[kPC] = woff + inc;
    __syncthreads();
    const int P = s_off[kPC];
    // ---- phase 2: one candidate test per thread and iteration
    int nh = 0;
    for (int pp = threadIdx.x; pp < P; pp += kPC) {
      int lo = 0, hi = kPC;  // largest i with s_off[i] <= pp
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        int mid = (lo + hi) >> 1;
        if (s_off[mid] <= pp) lo = mid; else hi = mid;
      }
      const int i = lo, k = pp - s_off[i];
      const int l0 = s_l0[i], l01 = s_l01[i];
      const int q = k < l0 ? s_qs[0][i] + k : (k < l01 ? s_qs[1][i] + (k - l0) : s_qs[2][i] + (k - l01));
      const int4 ot = srt[q];
      const int4 me = s_me[i];
      const int a = me.z;
      if (ot.z == a) continue;
      if (fl_changed(ot.w) && ot.z < a) continue;  // that agent owns the pair
      ++tests;
      if (!within(ot.x - me.x, ot.y - me.y, p)) continue;
      const FloorDev& fd = p.floors[fl_floor(me.w)];
      if (!loc_rule_ordered(fd, a, fl_loc(me.w), ot.z, fl_loc(ot.w))) continue;
      const uint32_t run = 1u + (uint32_t)min(fl_stay(me.w), fl_stay(ot.w));
      hits += run;
      ++runs;
      if (buffered && nh < kHitSlots) {
        ContactRec r;
        r.a1 = (uint32_t)min(a, ot.z); r.a2 = (uint32_t)max(a, ot.z);
        r.t_run = (t & 0xffffffu) | ((run - 1u) << 24);
        r.bucket = ((s_hb[i] + (uint32_t)(me.y + ot.y) * (uint32_t)fd.hist_w + (uint32_t)(me.x + ot.x)) & 0x3fffffffu) | 0x40000000u;
        mine[nh++] = r;
      } else {
        on_contact(p, t, me, ot, run, false, sub, true);
      }
    }
    if (buffered) {
      // one reservation for the whole warp
      int off = nh;
      for (int d = 1; d < 32; d <<= 1) {
        int v = __shfl_up_sync(0xffffffffu, off, d);
        if (lane >= d) off += v;
      }
      const int wtotal = __shfl_sync(0xffffffffu, off, 31);
      if (wtotal > 0) {
        unsigned long long wbase = 0;
        if (lane == 31) wbase = atomicAdd(&p.cbuf_count[sub * kCounterStride], (unsigned long long)wtotal);
        wbase = __shfl_sync(0xffffffffu, wbase, 31) + (unsigned long long)(off - nh);
        for (int j = 0; j < nh; ++j) {
          if (wbase + j < p.cbuf_sub_cap) {
            p.cbuf[(unsigned long long)sub * p.cbuf_sub_cap + wbase + j] = mine[j];
          } else {  // sub-buffer full: apply inline (k_accumulate clamps the count)
            const ContactRec r = mine[j];
            const uint32_t rt = r.t_run & 0xffffffu, rrun = (r.t_run >> 24) + 1u;
            bool was = contact_prev(p, (size_t)((int)rt - p.chunk_begin) * p.N, r.a1, r.a2);
            uint32_t first = (!was || (int32_t)rt == p.range_begin) ? rt : NO_STEP;
            pair_upsert(p, r.a1, r.a2, first, rrun, was ? 0u : 1u);
            atomicAdd(p.hist + (r.bucket & 0x3fffffffu), (unsigned long long)rrun);
          }
        }
      }
    }
    __syncthreads();  // the batch's shared arrays are reused
  }
  for (int d = 16; d > 0; d >>= 1) {
    tests += __shfl_down_sync(0xffffffffu, tests, d);
    hits += __shfl_down_sync(0xffffffffu, hits, d);
    runs += __shfl_down_sync(0xffffffffu, runs, d);
  }
  if (lane == 0 && (tests | hits)) {
    atomicAdd(&p.ctr->pair_tests, tests);
    if (hits) { atomicAdd(&p.ctr->contact_steps, hits); atomicAdd(&p.ctr->contact_runs, runs); }
  }
}

// K4b: apply the chunk's contact buffer to the tables.  One record per thread: every thread's
// table accesses are independent, so the kernel runs at the memory system's random-access rate
// instead of at the pace of k_pairs' longest neighbour list.  Per record: one 16-byte slot probe
// + one 64-bit compare-and-swap on the same sector (pair table) and one 64-bit reduction
// (coordinate histogram); for run starts that are not known continuations also two 16-byte
// record reads for the previous-step predicate, issued before the probe so they overlap it.
// many short-lived blocks rather than a persistent grid: the kernel shares the GPU with the next
// chunk's build on the other stream, and short blocks let that stream's blocks in (measured on config 3:
// 24 blocks per sub-buffer 2.08e10 agent-steps/s, 96: 2.27e10, 256: 2.37e10)
constexpr int kAccBlocksPerSub = 256;

__global__ void __launch_bounds__(256) k_accumulate(Params p, int t_begin) {
  const int sub = blockIdx.x % kSubBuffers, part = blockIdx.x / kSubBuffers;
  const unsigned long long n = min(p.cbuf_count[sub * kCounterStride], p.cbuf_sub_cap);
  const ContactRec* buf = p.cbuf + (unsigned long long)sub * p.cbuf_sub_cap;
  const unsigned long long stride = (unsigned long long)(gridDim.x / kSubBuffers) * blockDim.x;
  for (unsigned long long i = (unsigned long long)part * blockDim.x + threadIdx.x; i < n; i += stride) {
    ContactRec r = buf[i];
    uint32_t t = r.t_run & 0xffffffu, run = (r.t_run >> 24) + 1u;
    bool was = r.bucket >> 31;
    int4 r1 = make_int4(0, 0, 0, 0), r2 = r1;
    const bool late = (r.bucket >> 30) & 1;
    if (late) {
      size_t prow = (size_t)((int)t - t_begin) * p.N;  // rec row holding step t-1
      r1 = p.rec[prow + r.a1];
      r2 = p.rec[prow + r.a2];
    }
    // the histogram reduction and the table probe are independent: issue both, then update
    atomicAdd(p.hist + (r.bucket & 0x3fffffffu), (unsigned long long)run);
    unsigned long long acc;
    unsigned long long h = pair_slot(p, ((unsigned long long)r.a1 << 32) | r.a2, &acc);
    if (h == ~0ULL) continue;
    if (late) was = contact_eval(p, r1, r2);
    uint32_t first = (!was || (int32_t)t == p.range_begin) ? t : NO_STEP;
    pair_apply(p, h, acc, first, run, was ? 0u : 1u, r.a1, r.a2);
  }
}

__global__ void k_reset_cbuf(unsigned long long* cbuf_count) {
  if (threadIdx.x < kSubBuffers) cbuf_count[threadIdx.x * kCounterStride] = 0ULL;
}

// The integer distance threshold must agree with the float64 predicate on THIS device.
__global__ void k_check_s_max(double D, long long s_max, int* bad) {
  bool ok = (s_max < 0) ? !(__dsqrt_rn(0.0) < D)
                        : ((__dsqrt_rn((double)s_max) < D) && !(__dsqrt_rn((double)(s_max + 1)) < D));
  *bad = ok ? 0 : 1;
}

// --------------------------------------------------------------------------
// K0: location table.  One thread per lattice point evaluates the reference's rule
// (floorplan.py:215-241 first match; space.py:194-280; geometry.py:31-127,558-588) in
// float64 with explicitly rounded operations (no FMA contraction).
// --------------------------------------------------------------------------
__device__ __forceinline__ bool in_box(double px, double py, double qx, double qy, double rx, double ry) {
  return qx <= fmax(px, rx) && qx >= fmin(px, rx) && qy <= fmax(py, ry) && qy >= fmin(py, ry);
}

__device__ __forceinline__ int orient(double px, double py, double qx, double qy, double rx, double ry) {
  double v = __dsub_rn(__dmul_rn(__dsub_rn(qy, py), __dsub_rn(rx, qx)),
                       __dmul_rn(__dsub_rn(qx, px), __dsub_rn(ry, qy)));
  return v == 0.0 ? 0 : (v > 0.0 ? 1 : 2);
}

__device__ bool segments_cross(double p1x, double p1y, double q1x, double q1y, double p2x, double p2y,
                               double q2x, double q2y) {
  int o1 = orient(p1x, p1y, q1x, q1y, p2x, p2y);
  int o2 = orient(p1x, p1y, q1x, q1y, q2x, q2y);
  int o3 = orient(p2x, p2y, q2x, q2y, p1x, p1y);
  int o4 = orient(p2x, p2y, q2x, q2y, q1x, q1y);
  if (o1 != o2 && o3 != o4) return true;
  if (o1 == 0 && in_box(p1x, p1y, p2x, p2y, q1x, q1y)) return true;
  if (o2 == 0 && in_box(p1x, p1y, q2x, q2y, q1x, q1y)) return true;
  if (o3 == 0 && in_box(p2x, p2y, p1x, p1y, q2x, q2y)) return true;
  if (o4 == 0 && in_box(p2x, p2y, q1x, q1y, q2x, q2y)) return true;
  return false;
}

__global__ void __launch_bounds__(128) k_build_lut(int minx, int miny, int W, int H, int n_spaces,
                                                   const citam_space* __restrict__ spaces,
                                                   const citam_boundary* __restrict__ bnd,
                                                   const citam_door* __restrict__ doors, int16_t* __restrict__ lut) {
  int ix = blockIdx.x * blockDim.x + threadIdx.x;
  int iy = blockIdx.y;
  if (ix >= W || iy >= H) return;
  int xi = minx + ix, yi = miny + iy;
  double x = (double)xi, y = (double)yi;
  const double INF = 1e10;
  int result = -1;
  for (int s = 0; s < n_spaces && result < 0; ++s) {
    citam_space sp = spaces[s];
    if (xi < sp.bb_xmin || xi > sp.bb_xmax || yi < sp.bb_ymin || yi > sp.bb_ymax) continue;
    bool inside = false;
    for (int d = sp.door_begin; d < sp.door_end && !inside; ++d) {
      citam_door dr = doors[d];
      inside = in_box(dr.sx, dr.sy, x, y, dr.ex, dr.ey);
    }
    for (int b = sp.seg_begin; b < sp.seg_end && !inside; ++b) {
      citam_boundary sg = bnd[b];
      if (!sg.long_enough) continue;
      if ((x == sg.sx && y == sg.sy) || (x == sg.ex && y == sg.ey)) { inside = true; break; }
      if (in_box(sg.sx, sg.sy, x, y, sg.ex, sg.ey)) {
        double ddx = __dsub_rn(x, sg.sx), ddy = __dsub_rn(y, sg.sy);
        double h = hypot(ddx, ddy);
        // normal of start->point: -1j * (d / |d|)  ==  (d.imag/|d|, -d.real/|d|)
        double n2x = __ddiv_rn(ddy, h), n2y = -__ddiv_rn(ddx, h);
        if (fabs(__dsub_rn(sg.nx, n2x)) < 1e-3 && fabs(__dsub_rn(sg.ny, n2y)) < 1e-3) inside = true;
      }
    }
    if (!inside) {
      double ex[8] = {INF, __dsub_rn(x, 5.0), -INF, __dsub_rn(x, 5.0), INF, INF, -INF, -INF};
      double ey[8] = {__dadd_rn(y, 5.0), INF, __dadd_rn(y, 5.0), -INF, INF, -INF, INF, -INF};
#pragma unroll 1
      for (int r = 0; r < 8 && !inside; ++r) {
        int count = 0;
        for (int b = sp.seg_begin; b < sp.seg_end; ++b) {
          citam_boundary sg = bnd[b];
          count += segments_cross(sg.sx, sg.sy, sg.ex, sg.ey, x, y, ex[r], ey[r]) ? 1 : 0;
        }
        inside = (count & 1) != 0;
      }
    }
    if (inside) result = s;
  }
  lut[(size_t)iy * W + ix] = (int16_t)result;
}

// --------------------------------------------------------------------------
// K5: finalize — compaction of the tables and the integer sums of the statistics.
// --------------------------------------------------------------------------
__global__ void k_compact_pairs(const PairEntry* tab, unsigned long long cap, citam_pair* out,
                                unsigned long long out_cap, DevCounters* ctr) {
  unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cap) return;
  PairEntry e = tab[i];
  if (e.key == 0ULL) return;
  cg::coalesced_group g = cg::coalesced_threads();
  unsigned long long base = 0;
  if (g.thread_rank() == 0) base = atomicAdd(&ctr->export_cursor, (unsigned long long)g.size());
  base = g.shfl(base, 0) + g.thread_rank();
  if (base >= out_cap) return;
  citam_pair r;
  r.a1 = (uint32_t)(e.key >> 32);
  r.a2 = (uint32_t)(e.key & 0xffffffffu);
  uint32_t finv = ACC_FIRST_INV(e.acc);
  r.first_step = finv ? 0xFFFFFFu - finv : NO_STEP;
  r.n_events = ACC_NEV(e.acc);
  r.duration = ACC_DUR(e.acc);
  out[base] = r;
}

__global__ void k_compact_coords(const CoordEntry* tab, unsigned long long cap, citam_coord* out,
                                 unsigned long long out_cap, DevCounters* ctr) {
  unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cap) return;
  CoordEntry e = tab[i];
  if (e.key == 0ULL) return;
  cg::coalesced_group g = cg::coalesced_threads();
  unsigned long long base = 0;
  if (g.thread_rank() == 0) base = atomicAdd(&ctr->export_cursor, (unsigned long long)g.size());
  base = g.shfl(base, 0) + g.thread_rank();
  if (base >= out_cap) return;
  citam_coord r;
  r.floor = (int32_t)((e.key >> 56) & 0x7f);
  r.sx = (int32_t)((e.key >> 28) & 0xfffffffu) - (1 << 27);
  r.sy = (int32_t)(e.key & 0xfffffffu) - (1 << 27);
  r.reserved = 0;
  r.count = e.count;
  out[base] = r;
}

// dense histogram: count / compact the non-zero buckets of one floor
__global__ void k_hist_count(const unsigned long long* hist, long long n, DevCounters* ctr) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long nz = (i < n && hist[i] != 0ULL) ? 1ULL : 0ULL;
  for (int d = 16; d > 0; d >>= 1) nz += __shfl_down_sync(0xffffffffu, nz, d);
  if ((threadIdx.x & 31) == 0 && nz) atomicAdd(&ctr->export_cursor, nz);
}

__global__ void k_hist_compact(const unsigned long long* hist, long long n, int floor, int hist_w, int minx2, int miny2,
                               citam_coord* out, unsigned long long out_cap, DevCounters* ctr) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned long long c = hist[i];
  if (c == 0ULL) return;
  cg::coalesced_group g = cg::coalesced_threads();
  unsigned long long base = 0;
  if (g.thread_rank() == 0) base = atomicAdd(&ctr->export_cursor, (unsigned long long)g.size());
  base = g.shfl(base, 0) + g.thread_rank();
  if (base >= out_cap) return;
  citam_coord r;
  r.floor = floor;
  r.sx = (int32_t)(i % hist_w) + minx2;
  r.sy = (int32_t)(i / hist_w) + miny2;
  r.reserved = 0;
  r.count = c;
  out[base] = r;
}

__global__ void k_rehash_pairs(Params p, const PairEntry* old_tab, unsigned long long old_cap) {
  unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= old_cap) return;
  PairEntry e = old_tab[i];
  if (e.key == 0ULL) return;
  unsigned long long seen;
  unsigned long long h = pair_slot(p, e.key, &seen);
  if (h == ~0ULL) return;
  p.pairs[h].acc = e.acc;  // keys are unique in the old table: no other thread touches this slot
}

// pass 1: per-agent event sums and "has a pair" flags; pass 2: reduce over agents
__global__ void k_stats_pairs(const PairEntry* tab, unsigned long long cap, uint32_t* agent_nev,
                              uint32_t* agent_has, DevCounters* ctr) {
  unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long dur = 0, np = 0;
  if (i < cap) {
    PairEntry e = tab[i];
    if (e.key != 0ULL) {
      uint32_t a1 = (uint32_t)(e.key >> 32), a2 = (uint32_t)(e.key & 0xffffffffu);
      uint32_t nev = ACC_NEV(e.acc);
      atomicAdd(&agent_nev[a1], nev);
      atomicAdd(&agent_nev[a2], nev);
      agent_has[a1] = 1;
      agent_has[a2] = 1;
      dur = ACC_DUR(e.acc);
      np = 1;
    }
  }
  for (int d = 16; d > 0; d >>= 1) {
    dur += __shfl_down_sync(0xffffffffu, dur, d);
    np += __shfl_down_sync(0xffffffffu, np, d);
  }
  if ((threadIdx.x & 31) == 0 && np) {
    atomicAdd(&ctr->stats[0], dur);
    atomicAdd(&ctr->stats[1], np);
  }
}

__global__ void k_stats_agents(const uint32_t* agent_nev, const uint32_t* agent_has, int N, DevCounters* ctr) {
  int a = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long has = 0, nev = 0, mx = 0;
  if (a < N) { has = agent_has[a]; nev = agent_nev[a]; mx = nev; }
  for (int d = 16; d > 0; d >>= 1) {
    has += __shfl_down_sync(0xffffffffu, has, d);
    nev += __shfl_down_sync(0xffffffffu, nev, d);
    mx = max(mx, __shfl_down_sync(0xffffffffu, mx, d));
  }
  if ((threadIdx.x & 31) == 0 && (has | nev)) {
    atomicAdd(&ctr->stats[2], has);
    atomicAdd(&ctr->stats[3], nev);
    atomicMax(&ctr->stats[4], mx);
  }
}

__global__ void k_merge_pairs(Params p, const citam_pair* in, unsigned long long n) {
  unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  citam_pair r = in[i];
  pair_upsert(p, r.a1, r.a2, r.first_step, r.duration, r.n_events);
}

__global__ void k_merge_coords(Params p, const citam_coord* in, unsigned long long n) {
  unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  citam_coord r = in[i];
  coord_upsert(p, r.floor, r.sx, r.sy, r.count);
}

__global__ void k_repack_segments(const citam_segment* in, long long n, int4* out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  citam_segment s = in[i];
  out[i] = make_int4((s.t0 & 0xffffff) | ((int)(s.floor & 0xff) << 24), s.x0, s.y0,
                     (int)(((uint32_t)(uint16_t)s.dx) | ((uint32_t)(uint16_t)s.dy << 16)));
}

// --------------------------------------------------------------------------
// host side
// --------------------------------------------------------------------------
struct citam_engine {
  citam_config cfg{};
  int device = 0;
  cudaStream_t stream = nullptr;
  int N = 0, F = 0, cell = 1;
  double D = 6.0;
  long long s_max = 35;
  std::vector<FloorDev> floors_h;
  std::vector<int16_t*> lut_d;
  std::vector<uint32_t*> adj_d;
  FloorDev* floors_d = nullptr;
  bool floors_dirty = true;
  int total_cells = 0, cells_stride = 4;

  int4* segs_d = nullptr;
  citam_segment* raw_d = nullptr;  // staging copy of the caller's records
  int64_t segs_cap = 0;
  long long* seg_off_d = nullptr;
  long long n_segs = 0;
  int2* window_d = nullptr;  // k_agent_windows
  bool window_dirty = true;
  int acc_blocks = kAccBlocksPerSub;  // k_accumulate blocks per sub-buffer (CITAM_ACC_BLOCKS overrides)
  bool explicit_states = false;  // citam_step_states in progress: states come from the caller, no windows
  bool have_itin = false;

  int S = 0;  // chunk capacity in steps
  int4* rec = nullptr;  // the current one of rec_buf
  int4* rec_buf[2] = {nullptr, nullptr};
  int cur = 0;  // buffer set of the chunk being built (records + contact buffer are double-buffered)
  cudaStream_t stream2 = nullptr;  // k_accumulate of chunk i runs here while chunk i+1 is built on `stream`
  cudaEvent_t ev_built[2] = {nullptr, nullptr}, ev_applied[2] = {nullptr, nullptr};
  bool applied_pending[2] = {false, false};
  uint32_t* cells = nullptr;
  int4* sorted = nullptr;

  PairEntry* pairs = nullptr;
  unsigned long long pair_cap = 0;
  bool pair_auto = false;  // capacity chosen by the library: grow when half full
  unsigned long long* hist = nullptr;  // dense coordinate histogram (all floors), NULL = hashed
  long long hist_total = 0;
  int32_t range_begin = -1;
  CoordEntry* coords = nullptr;
  unsigned long long coord_cap = 0;
  citam_contact* log = nullptr;
  unsigned long long log_cap = 0;
  ContactRec* cbuf = nullptr;  // the current one of cbuf_buf
  ContactRec* cbuf_buf[2] = {nullptr, nullptr};
  unsigned long long cbuf_cap = 0;
  uint32_t *chg_list = nullptr, *chg_count = nullptr, *part_sums = nullptr;
  unsigned long long* agent_dur = nullptr;
  uint32_t *agent_nev = nullptr, *agent_has = nullptr;
  DevCounters* ctr_d = nullptr;

  // Calculator.run entry: previous explicit state row
  int4* prev = nullptr;
  long long prev_step = LLONG_MIN;
  citam_agent_state* states_d = nullptr;

  unsigned long long agent_steps = 0, launches = 0;
  bool timing = false;
  struct Ev { cudaEvent_t a, b; int kind; };
  std::vector<Ev> evs;
  double t_ms[CITAM_N_KERNELS] = {0};
  unsigned long long t_n[CITAM_N_KERNELS] = {0};
};

constexpr int kMaxScanParts = 32;
constexpr long long kDenseHistMax = (1LL << 30) - 1;  // buckets (8 GiB); bucket ids must fit 30 bits

// Largest integer s with sqrt((double)s) < D, by bisection on the monotone predicate, using the
// host's correctly rounded IEEE sqrt (the same function the device's __dsqrt_rn computes).
static long long find_s_max(double D) {
  auto ok = [D](long long s) { return std::sqrt((double)s) < D; };
  if (!ok(0)) return -1;
  long long lo = 0, hi = 1;
  while (ok(hi)) { lo = hi; hi *= 2; if (hi > (1LL << 60)) break; }
  while (hi - lo > 1) {
    long long mid = lo + (hi - lo) / 2;
    if (ok(mid)) lo = mid; else hi = mid;
  }
  return lo;
}

static unsigned long long next_pow2(unsigned long long v) {
  unsigned long long p = 1;
  while (p < v) p <<= 1;
  return p;
}

struct Timed {
  citam_engine* e;
  int kind;
  cudaStream_t s;
  cudaEvent_t a = nullptr, b = nullptr;
  Timed(citam_engine* e_, int k, cudaStream_t s_ = nullptr, bool use_s = false) : e(e_), kind(k), s(use_s ? s_ : e_->stream) {
    e->launches++;
    if (e->timing) {
      cudaEventCreate(&a);
      cudaEventCreate(&b);
      cudaEventRecord(a, s);
    }
  }
  ~Timed() {
    if (e->timing) {
      cudaEventRecord(b, s);
      e->evs.push_back({a, b, kind});
    }
  }
};

// Start building a chunk: take the other buffer set, once the accumulation that last read it is done.
static int begin_chunk(citam_engine* e) {
  e->cur ^= 1;
  e->rec = e->rec_buf[e->cur];
  e->cbuf = e->cbuf_buf[e->cur];
  if (e->applied_pending[e->cur]) {
    CK(cudaStreamWaitEvent(e->stream, e->ev_applied[e->cur], 0));
    e->applied_pending[e->cur] = false;
  }
  return CITAM_OK;
}

// Everything enqueued on the side stream happens before whatever follows on the engine's stream.
static int join_side_stream(citam_engine* e) {
  for (int b = 0; b < 2; ++b)
    if (e->applied_pending[b]) {
      CK(cudaStreamWaitEvent(e->stream, e->ev_applied[b], 0));
      e->applied_pending[b] = false;
    }
  return CITAM_OK;
}

static void drain_timings(citam_engine* e) {
  for (auto& v : e->evs) {
    cudaEventSynchronize(v.b);
    float ms = 0;
    cudaEventElapsedTime(&ms, v.a, v.b);
    e->t_ms[v.kind] += ms;
    e->t_n[v.kind]++;
    cudaEventDestroy(v.a);
    cudaEventDestroy(v.b);
  }
  e->evs.clear();
}

static unsigned long long pairs_used(const DevCounters& c) {
  unsigned long long n = c.pair_used;
  for (int i = 0; i < kSubBuffers; ++i) n += c.pair_used_sub[i * kCounterStride];
  return n;
}

static Params make_params(citam_engine* e) {
  Params p{};
  p.N = e->N; p.F = e->F; p.cell = e->cell; p.total_cells = e->total_cells; p.cells_stride = e->cells_stride;
  p.D = e->D; p.s_max = e->s_max; p.floors = e->floors_d; p.segs = e->segs_d; p.seg_off = e->seg_off_d;
  p.window = (e->have_itin && !e->window_dirty && !e->explicit_states) ? e->window_d : nullptr;
  p.rec = e->rec; p.cells = e->cells;
  p.sorted = e->sorted; p.pairs = e->pairs; p.pair_mask = e->pair_cap - 1;
  p.coords = e->coords; p.coord_mask = e->coord_cap - 1; p.hist = e->hist; p.range_begin = e->range_begin;
  p.log = e->log; p.log_cap = e->log_cap;
  p.cbuf = (e->log_cap || !e->hist) ? nullptr : e->cbuf; p.cbuf_sub_cap = e->cbuf_cap / kSubBuffers;
  p.cbuf_count = e->ctr_d->cbuf_count[e->cur];
  p.agent_dur = e->agent_dur;
  const bool agg = e->log_cap == 0;  // run aggregation (off when every contact-step is logged)
  p.chg_list = agg ? e->chg_list : nullptr; p.chg_count = agg ? e->chg_count : nullptr;
  p.ctr = e->ctr_d;
  return p;
}

static void free_chunk(citam_engine* e) {
  cudaStreamSynchronize(e->stream);
  if (e->stream2) cudaStreamSynchronize(e->stream2);
  e->applied_pending[0] = e->applied_pending[1] = false;
  for (int b = 0; b < 2; ++b) { cudaFree(e->rec_buf[b]); cudaFree(e->cbuf_buf[b]); e->rec_buf[b] = nullptr; e->cbuf_buf[b] = nullptr; }
  cudaFree(e->cells);
  cudaFree(e->sorted); cudaFree(e->chg_list); cudaFree(e->chg_count); cudaFree(e->part_sums);
  e->chg_list = e->chg_count = e->part_sums = nullptr;
  e->rec = nullptr; e->cells = nullptr; e->sorted = nullptr;
  e->cbuf = nullptr; e->cbuf_cap = 0;
  e->S = 0;
}

// Upload floor descriptors and (re)size the chunk buffers for `want_steps` steps.
static int ensure_ready(citam_engine* e, int want_steps) {
  CK(cudaSetDevice(e->device));
  if (e->floors_dirty) {
    int base = 0;
    for (auto& f : e->floors_h) {
      f.ncx = f.lut ? (f.W + e->cell - 1) / e->cell : 0;
      f.ncy = f.lut ? (f.H + e->cell - 1) / e->cell : 0;
      f.cell_base = base;
      base += f.ncx * f.ncy;
    }
    e->total_cells = base;
    e->cells_stride = ((base + 1) + 3) & ~3;
    // dense per-coordinate histogram over every floor's half-unit lattice (contact positions are
    // midpoints of integer points); lattices beyond kDenseHistMax buckets use the hashed table.
    // Changing a floor's lattice drops the histogram's contents (documented in the header).
    long long tot = 0;
    bool same = true;
    for (auto& f : e->floors_h) {
      int hw = f.lut ? 2 * f.W - 1 : 0, hh = f.lut ? 2 * f.H - 1 : 0;
      if (f.hist_base != tot || f.hist_w != hw || f.hist_h != hh) same = false;
      f.hist_base = tot; f.hist_w = hw; f.hist_h = hh;
      tot += (long long)hw * hh;
    }
    if (!same || tot != e->hist_total) {
      cudaFree(e->hist);
      e->hist = nullptr;
      e->hist_total = 0;
      if (tot > 0 && tot <= kDenseHistMax && !(e->cfg.flags & CITAM_FLAG_HASHED_COORDS)) {
        CK(cudaMalloc(&e->hist, sizeof(unsigned long long) * (size_t)tot));
        CK(cudaMemsetAsync(e->hist, 0, sizeof(unsigned long long) * (size_t)tot, e->stream));
        e->hist_total = tot;
      }
    }
    CK(cudaMemcpyAsync(e->floors_d, e->floors_h.data(), sizeof(FloorDev) * e->F, cudaMemcpyHostToDevice, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    e->floors_dirty = false;
    e->window_dirty = true;  // the final-stay test reads the location tables
    free_chunk(e);
  }
  if (e->S >= want_steps && e->rec) return CITAM_OK;
  free_chunk(e);
  int S = want_steps;
  size_t n = (size_t)e->N;
  for (int b = 0; b < 2; ++b) CK(cudaMalloc(&e->rec_buf[b], sizeof(int4) * n * (S + 1)));
  e->rec = e->rec_buf[e->cur];
  CK(cudaMalloc(&e->cells, sizeof(uint32_t) * (size_t)e->cells_stride * S));
  CK(cudaMalloc(&e->sorted, sizeof(int4) * n * S));
  // contact buffer: one record per detected contact run of a chunk (overflow falls back to
  // inline accumulation, so the size is a performance knob, not a limit)
  e->cbuf_cap = std::min<unsigned long long>(128ULL << 20, std::max<unsigned long long>(1ULL << 16, 2ULL * n * S));
  e->cbuf_cap = (e->cbuf_cap + kSubBuffers - 1) / kSubBuffers * kSubBuffers;
  for (int b = 0; b < 2; ++b) CK(cudaMalloc(&e->cbuf_buf[b], sizeof(ContactRec) * e->cbuf_cap));
  e->cbuf = e->cbuf_buf[e->cur];
  CK(cudaMalloc(&e->chg_list, sizeof(uint32_t) * n * S));
  CK(cudaMalloc(&e->chg_count, sizeof(uint32_t) * 32 * S));
  CK(cudaMalloc(&e->part_sums, sizeof(uint32_t) * kMaxScanParts * S));
  e->S = S;
  return CITAM_OK;
}

static int pick_chunk(citam_engine* e, int steps) {
  if (e->cfg.steps_per_chunk > 0) return std::min(std::min(steps, e->cfg.steps_per_chunk), kMaxChunk);
  // ~64M agent-steps per launch group (2 GiB of per-step records), bounded by 1 GiB of histogram
  // rows and by the 8-bit stationary-run counter
  long long by_agents = std::max(1LL, (64LL << 20) / std::max(1, e->N));
  long long by_cells = std::max(1LL, (1LL << 28) / std::max(4, e->cells_stride));
  long long s = std::min(by_agents, by_cells);
  s = std::min<long long>(s, kMaxChunk);
  s = std::max<long long>(s, 1);
  return (int)std::min<long long>(s, steps);
}

// bin (already done by caller) -> scan -> scatter -> pairs for `rows` steps starting at t_begin
static int run_sorted_stages(citam_engine* e, int t_begin, int rows) {
  Params p = make_params(e);
  p.chunk_begin = t_begin;
  {
    Timed tm(e, CITAM_K_SCAN);
    const int n4 = (e->total_cells + 1 + 3) >> 2;
    int parts = std::max(1, std::min(kMaxScanParts, std::min((296 + rows - 1) / rows, (n4 + 4095) / 4096)));
    int part_len4 = (n4 + parts - 1) / parts;
    parts = (n4 + part_len4 - 1) / part_len4;
    k_scan<<<dim3(parts, rows), kScanThreads, 0, e->stream>>>(e->cells, e->cells_stride, e->total_cells + 1, part_len4,
                                                              e->part_sums);
    if (parts > 1) e->launches++;
    if (parts > 1)
      k_scan_add<<<dim3(parts - 1, rows), 256, 0, e->stream>>>(e->cells, e->cells_stride, e->total_cells + 1, part_len4,
                                                               e->part_sums);
  }
  {
    Timed tm(e, CITAM_K_SCATTER);
    dim3 g((e->N + 255) / 256, rows);
    k_scatter<<<g, 256, 0, e->stream>>>(p, rows, t_begin);
    if (p.chg_count != nullptr && rows > 1) {
      e->launches++;
      dim3 g3((unsigned)std::min(296, std::max(1, (e->N + 255) / 256)), rows - 1);
      k_select_changed<<<g3, 256, 0, e->stream>>>(p, rows);
    }
  }
  {
    Timed tm(e, CITAM_K_PAIRS);
    const bool agg = p.chg_count != nullptr;
    // first step of the chunk (or every step when the log is on): all located agents
    dim3 g((e->N + 255) / 256, agg ? 1 : rows);
    k_pairs<<<g, 256, 0, e->stream>>>(p, t_begin, rows, agg ? 1 : 0);
    if (agg && rows > 1) {
      e->launches++;
      dim3 g2((unsigned)std::min(592, std::max(1, (e->N + 255) / 256)), rows - 1);
      k_pairs_changed<<<g2, 256, 0, e->stream>>>(p, t_begin, rows);
    }
  }
  if (p.cbuf != nullptr) {
    // the contact buffer is applied on the side stream, overlapping the next chunk's build
    CK(cudaEventRecord(e->ev_built[e->cur], e->stream));
    CK(cudaStreamWaitEvent(e->stream2, e->ev_built[e->cur], 0));
    {
      Timed tm(e, CITAM_K_ACCUMULATE, e->stream2, true);
      k_accumulate<<<kSubBuffers * e->acc_blocks, 256, 0, e->stream2>>>(p, t_begin);
      k_reset_cbuf<<<1, kSubBuffers, 0, e->stream2>>>(p.cbuf_count);
      e->launches++;
    }
    CK(cudaEventRecord(e->ev_applied[e->cur], e->stream2));
    e->applied_pending[e->cur] = true;
  }
  CK(cudaGetLastError());
  return CITAM_OK;
}

static int check_overflow(citam_engine* e) {
  DevCounters c;
  CK(cudaMemcpyAsync(&c, e->ctr_d, sizeof c, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  if (c.overflow & OVF_PAIR) return fail(CITAM_ERR_CAPACITY, "pair table full (%llu slots): raise pair_capacity", e->pair_cap);
  if (c.overflow & OVF_COORD) return fail(CITAM_ERR_CAPACITY, "coordinate table full (%llu slots): raise coord_capacity", e->coord_cap);
  if (c.overflow & OVF_LOG) return fail(CITAM_ERR_CAPACITY, "contact log full (%llu records): raise log_capacity", e->log_cap);
  if (c.overflow & OVF_FIELD) return fail(CITAM_ERR_CAPACITY, "a pair exceeded 2^24-1 contact-steps or 65535 events");
  if (c.overflow & OVF_GRID) return fail(CITAM_ERR_INVALID, "an agent was given a location but lies outside its floor's lattice");
  return CITAM_OK;
}

extern "C" {

int citam_abi_version(void) { return CITAM_ABI_VERSION; }
const char* citam_last_error(void) { return g_err.c_str(); }

int citam_create(const citam_config* cfg, citam_engine** out) {
  if (!cfg || !out) return fail(CITAM_ERR_INVALID, "null argument");
  if (cfg->n_agents <= 0 || cfg->n_floors <= 0 || cfg->n_floors > kMaxFloors)
    return fail(CITAM_ERR_INVALID, "n_agents must be > 0 and n_floors in 1..63");
  if (!(cfg->contact_distance >= 0.0) || cfg->contact_distance > 1e6)
    return fail(CITAM_ERR_INVALID, "contact_distance out of range");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(CITAM_ERR_NO_DEVICE, "no CUDA device: this library has no CPU fallback");
  if (cfg->device < 0 || cfg->device >= ndev) return fail(CITAM_ERR_INVALID, "device %d out of range", cfg->device);
  CK(cudaSetDevice(cfg->device));
  citam_engine* e = new citam_engine();
  if (const char* g = getenv("CITAM_ACC_BLOCKS")) e->acc_blocks = std::max(1, atoi(g));
  e->cfg = *cfg;
  e->device = cfg->device;
  e->N = cfg->n_agents;
  e->F = cfg->n_floors;
  e->D = cfg->contact_distance;
  e->s_max = find_s_max(e->D);
  int mincell = (int)std::ceil(cfg->contact_distance);
  if (mincell < 1) mincell = 1;
  e->cell = cfg->cell_size > 0 ? std::max(cfg->cell_size, mincell) : mincell;
  e->timing = (cfg->flags & CITAM_FLAG_TIMING) != 0;
  e->floors_h.assign(e->F, FloorDev{});
  e->lut_d.assign(e->F, nullptr);
  e->adj_d.assign(e->F, nullptr);
  e->pair_auto = cfg->pair_capacity == 0;
  e->pair_cap = next_pow2(cfg->pair_capacity ? cfg->pair_capacity : std::min<unsigned long long>(1ULL << 28, std::max<unsigned long long>(1ULL << 22, 256ULL * e->N)));
  e->coord_cap = next_pow2(cfg->coord_capacity ? cfg->coord_capacity : std::min<unsigned long long>(1ULL << 27, std::max<unsigned long long>(1ULL << 20, 64ULL * e->N)));
  e->log_cap = cfg->log_capacity;
  *out = e;
  size_t n = (size_t)e->N;
  CK(cudaMalloc(&e->floors_d, sizeof(FloorDev) * e->F));
  CK(cudaMalloc(&e->pairs, sizeof(PairEntry) * e->pair_cap));
  CK(cudaMalloc(&e->coords, sizeof(CoordEntry) * e->coord_cap));
  if (e->log_cap) CK(cudaMalloc(&e->log, sizeof(citam_contact) * e->log_cap));
  CK(cudaMalloc(&e->agent_dur, sizeof(unsigned long long) * n));
  CK(cudaMalloc(&e->agent_nev, sizeof(uint32_t) * n));
  CK(cudaMalloc(&e->agent_has, sizeof(uint32_t) * n));
  CK(cudaMalloc(&e->ctr_d, sizeof(DevCounters)));
  CK(cudaMalloc(&e->prev, sizeof(int4) * n));
  CK(cudaMalloc(&e->states_d, sizeof(citam_agent_state) * n));
  CK(cudaStreamCreateWithFlags(&e->stream2, cudaStreamNonBlocking));
  for (int b = 0; b < 2; ++b) {
    CK(cudaEventCreateWithFlags(&e->ev_built[b], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&e->ev_applied[b], cudaEventDisableTiming));
  }
  {
    int* bad_d = nullptr;
    int bad = 1;
    CK(cudaMalloc(&bad_d, sizeof(int)));
    k_check_s_max<<<1, 1, 0, e->stream>>>(e->D, e->s_max, bad_d);
    CK(cudaMemcpyAsync(&bad, bad_d, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    cudaFree(bad_d);
    if (bad) return fail(CITAM_ERR_CUDA, "integer distance threshold %lld disagrees with the float64 predicate", e->s_max);
  }
  return citam_reset(e);
}

void citam_destroy(citam_engine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  cudaStreamSynchronize(e->stream);
  if (e->stream2) cudaStreamSynchronize(e->stream2);
  drain_timings(e);
  free_chunk(e);
  if (e->stream2) cudaStreamDestroy(e->stream2);
  for (int b = 0; b < 2; ++b) { if (e->ev_built[b]) cudaEventDestroy(e->ev_built[b]); if (e->ev_applied[b]) cudaEventDestroy(e->ev_applied[b]); }
  for (auto p : e->lut_d) cudaFree(p);
  for (auto p : e->adj_d) cudaFree(p);
  cudaFree(e->floors_d); cudaFree(e->segs_d); cudaFree(e->raw_d); cudaFree(e->seg_off_d); cudaFree(e->window_d); cudaFree(e->pairs);
  cudaFree(e->coords); cudaFree(e->hist);
  cudaFree(e->log); cudaFree(e->agent_dur); cudaFree(e->agent_nev); cudaFree(e->agent_has); cudaFree(e->ctr_d);
  cudaFree(e->prev); cudaFree(e->states_d);
  delete e;
}

int citam_set_stream(citam_engine* e, void* s) {
  if (!e) return fail(CITAM_ERR_INVALID, "null engine");
  e->stream = (cudaStream_t)s;
  return CITAM_OK;
}

int citam_reset(citam_engine* e) {
  if (!e) return fail(CITAM_ERR_INVALID, "null engine");
  CK(cudaSetDevice(e->device));
  CK(cudaMemsetAsync(e->pairs, 0, sizeof(PairEntry) * e->pair_cap, e->stream));
  CK(cudaMemsetAsync(e->coords, 0, sizeof(CoordEntry) * e->coord_cap, e->stream));
  if (e->hist) CK(cudaMemsetAsync(e->hist, 0, sizeof(unsigned long long) * (size_t)e->hist_total, e->stream));
  CK(cudaMemsetAsync(e->agent_dur, 0, sizeof(unsigned long long) * (size_t)e->N, e->stream));
  CK(cudaMemsetAsync(e->ctr_d, 0, sizeof(DevCounters), e->stream));
  e->prev_step = LLONG_MIN;
  e->agent_steps = 0;
  return CITAM_OK;
}

static int install_lut(citam_engine* e, int floor, int minx, int miny, int W, int H, int n_spaces) {
  if (floor < 0 || floor >= e->F) return fail(CITAM_ERR_INVALID, "floor %d out of range", floor);
  if (W <= 0 || H <= 0 || (long long)W * H > (1LL << 31)) return fail(CITAM_ERR_INVALID, "bad lattice %dx%d", W, H);
  if (n_spaces < 0 || n_spaces > 32767) return fail(CITAM_ERR_INVALID, "n_spaces must be <= 32767");
  CK(cudaSetDevice(e->device));
  cudaFree(e->lut_d[floor]);
  e->lut_d[floor] = nullptr;
  CK(cudaMalloc(&e->lut_d[floor], sizeof(int16_t) * (size_t)W * H));
  FloorDev& f = e->floors_h[floor];
  f.minx = minx; f.miny = miny; f.W = W; f.H = H; f.n_spaces = n_spaces; f.lut = e->lut_d[floor];
  e->floors_dirty = true;
  return CITAM_OK;
}

int citam_set_floor_lut(citam_engine* e, int32_t floor, int32_t minx, int32_t miny, int32_t W, int32_t H,
                        int32_t n_spaces, const int16_t* lut) {
  if (!e || !lut) return fail(CITAM_ERR_INVALID, "null argument");
  int rc = install_lut(e, floor, minx, miny, W, H, n_spaces);
  if (rc) return rc;
  CK(cudaMemcpyAsync(e->lut_d[floor], lut, sizeof(int16_t) * (size_t)W * H, cudaMemcpyHostToDevice, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  return CITAM_OK;
}

int citam_build_floor_lut(citam_engine* e, int32_t floor, int32_t minx, int32_t miny, int32_t W, int32_t H,
                          int32_t n_spaces, const citam_space* spaces, int32_t n_boundaries,
                          const citam_boundary* boundaries, int32_t n_doors, const citam_door* doors) {
  if (!e || (n_spaces > 0 && !spaces)) return fail(CITAM_ERR_INVALID, "null argument");
  for (int s = 0; s < n_spaces; ++s) {
    const citam_space& sp = spaces[s];
    if (sp.seg_begin < 0 || sp.seg_end > n_boundaries || sp.seg_begin > sp.seg_end || sp.door_begin < 0 ||
        sp.door_end > n_doors || sp.door_begin > sp.door_end)
      return fail(CITAM_ERR_INVALID, "space %d has a bad segment/door range", s);
  }
  int rc = install_lut(e, floor, minx, miny, W, H, n_spaces);
  if (rc) return rc;
  citam_space* sp_d = nullptr;
  citam_boundary* b_d = nullptr;
  citam_door* d_d = nullptr;
  CK(cudaMalloc(&sp_d, sizeof(citam_space) * std::max(1, n_spaces)));
  CK(cudaMalloc(&b_d, sizeof(citam_boundary) * std::max(1, n_boundaries)));
  CK(cudaMalloc(&d_d, sizeof(citam_door) * std::max(1, n_doors)));
  if (n_spaces) CK(cudaMemcpyAsync(sp_d, spaces, sizeof(citam_space) * n_spaces, cudaMemcpyHostToDevice, e->stream));
  if (n_boundaries) CK(cudaMemcpyAsync(b_d, boundaries, sizeof(citam_boundary) * n_boundaries, cudaMemcpyHostToDevice, e->stream));
  if (n_doors) CK(cudaMemcpyAsync(d_d, doors, sizeof(citam_door) * n_doors, cudaMemcpyHostToDevice, e->stream));
  {
    Timed tm(e, CITAM_K_LUT);
    dim3 g((W + 127) / 128, H);
    k_build_lut<<<g, 128, 0, e->stream>>>(minx, miny, W, H, n_spaces, sp_d, b_d, d_d, e->lut_d[floor]);
  }
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(e->stream));
  cudaFree(sp_d); cudaFree(b_d); cudaFree(d_d);
  return CITAM_OK;
}

int citam_get_floor_lut(citam_engine* e, int32_t floor, int16_t* out) {
  if (!e || !out || floor < 0 || floor >= e->F || !e->lut_d[floor]) return fail(CITAM_ERR_INVALID, "floor has no table");
  CK(cudaSetDevice(e->device));
  const FloorDev& f = e->floors_h[floor];
  CK(cudaMemcpyAsync(out, e->lut_d[floor], sizeof(int16_t) * (size_t)f.W * f.H, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  return CITAM_OK;
}

int citam_set_floor_adjacency(citam_engine* e, int32_t floor, int32_t n_spaces, const uint8_t* adj) {
  if (!e || floor < 0 || floor >= e->F) return fail(CITAM_ERR_INVALID, "floor out of range");
  CK(cudaSetDevice(e->device));
  cudaFree(e->adj_d[floor]);
  e->adj_d[floor] = nullptr;
  e->floors_h[floor].adj = nullptr;
  e->floors_dirty = true;
  if (!adj) return CITAM_OK;
  if (n_spaces != e->floors_h[floor].n_spaces)
    return fail(CITAM_ERR_INVALID, "adjacency is %d spaces, floor table has %d (load the table first)", n_spaces,
                e->floors_h[floor].n_spaces);
  size_t bits = (size_t)n_spaces * n_spaces, words = (bits + 31) / 32;
  std::vector<uint32_t> w(std::max<size_t>(1, words), 0u);
  for (size_t i = 0; i < bits; ++i)
    if (adj[i]) w[i >> 5] |= 1u << (i & 31);
  CK(cudaMalloc(&e->adj_d[floor], sizeof(uint32_t) * w.size()));
  CK(cudaMemcpyAsync(e->adj_d[floor], w.data(), sizeof(uint32_t) * w.size(), cudaMemcpyHostToDevice, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  e->floors_h[floor].adj = e->adj_d[floor];
  return CITAM_OK;
}

int citam_load_itineraries(citam_engine* e, const int64_t* seg_off, const citam_segment* segs, int64_t n_segs) {
  if (!e || !seg_off || (n_segs > 0 && !segs)) return fail(CITAM_ERR_INVALID, "null argument");
  if (seg_off[0] != 0 || seg_off[e->N] != n_segs) return fail(CITAM_ERR_INVALID, "seg_off must run from 0 to n_segs");
  CK(cudaSetDevice(e->device));
  e->have_itin = false;
  // buffers are kept between calls (time-block streaming reloads a slice per block)
  if (n_segs > e->segs_cap || !e->segs_d) {
    cudaFree(e->segs_d); cudaFree(e->raw_d);
    e->segs_d = nullptr; e->raw_d = nullptr; e->segs_cap = 0;
    int64_t cap = std::max<int64_t>(1024, n_segs + n_segs / 4);
    CK(cudaMalloc(&e->segs_d, sizeof(int4) * cap));
    CK(cudaMalloc(&e->raw_d, sizeof(citam_segment) * cap));
    e->segs_cap = cap;
  }
  if (!e->seg_off_d) CK(cudaMalloc(&e->seg_off_d, sizeof(long long) * ((size_t)e->N + 1)));
  citam_segment* raw = e->raw_d;
  CK(cudaMemcpyAsync(e->seg_off_d, seg_off, sizeof(long long) * ((size_t)e->N + 1), cudaMemcpyHostToDevice, e->stream));
  if (n_segs) {
    CK(cudaMemcpyAsync(raw, segs, sizeof(citam_segment) * n_segs, cudaMemcpyHostToDevice, e->stream));
    Timed tm(e, CITAM_K_MISC);
    k_repack_segments<<<(unsigned)((n_segs + 255) / 256), 256, 0, e->stream>>>(raw, n_segs, e->segs_d);
  }
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(e->stream));  // the caller's host buffers are free to change again
  e->n_segs = n_segs;
  e->have_itin = true;
  e->window_dirty = true;
  return CITAM_OK;
}

// Library-chosen pair capacity: double the table once it is half full (checked between chunks).
static int maybe_grow_pairs(citam_engine* e) {
  if (!e->pair_auto) return CITAM_OK;
  {  // the table is about to be read (and perhaps replaced): no accumulation may be in flight
    int rcj = join_side_stream(e);
    if (rcj) return rcj;
  }
  DevCounters c;
  CK(cudaMemcpyAsync(&c, e->ctr_d, sizeof c, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  while (pairs_used(c) * 2 > e->pair_cap && !(c.overflow & OVF_PAIR)) {
    unsigned long long ncap = e->pair_cap * 2;
    PairEntry* np = nullptr;
    if (cudaMalloc(&np, sizeof(PairEntry) * ncap) != cudaSuccess) {
      cudaGetLastError();
      e->pair_auto = false;  // out of memory: keep the table, overflow is reported if it comes
      return CITAM_OK;
    }
    CK(cudaMemsetAsync(np, 0, sizeof(PairEntry) * ncap, e->stream));
    PairEntry* old = e->pairs;
    unsigned long long ocap = e->pair_cap;
    e->pairs = np; e->pair_cap = ncap;
    CK(cudaMemsetAsync(&e->ctr_d->pair_used, 0, sizeof(unsigned long long), e->stream));
    CK(cudaMemsetAsync(e->ctr_d->pair_used_sub, 0, sizeof(c.pair_used_sub), e->stream));
    Params p = make_params(e);
    {
      Timed tm(e, CITAM_K_MISC);
      k_rehash_pairs<<<(unsigned)((ocap + 255) / 256), 256, 0, e->stream>>>(p, old, ocap);
    }
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(e->stream));
    cudaFree(old);
  }
  return CITAM_OK;
}

int citam_run(citam_engine* e, int32_t t_begin, int32_t t_end) {
  if (!e) return fail(CITAM_ERR_INVALID, "null engine");
  if (!e->have_itin) return fail(CITAM_ERR_STATE, "citam_run before citam_load_itineraries");
  if (t_begin < 0 || t_end < t_begin || t_end > (1 << 24)) return fail(CITAM_ERR_INVALID, "bad step range");
  if (t_end == t_begin) return CITAM_OK;
  CK(cudaSetDevice(e->device));
  int rc = ensure_ready(e, 1);  // uploads floor descriptors so pick_chunk sees the grid
  if (rc) return rc;
  int S = pick_chunk(e, t_end - t_begin);
  rc = ensure_ready(e, S);
  if (rc) return rc;
  if (e->window_dirty) {
    if (!e->window_d) CK(cudaMalloc(&e->window_d, sizeof(int2) * (size_t)e->N));
    Params p0 = make_params(e);
    {
      Timed tm(e, CITAM_K_MISC);
      k_agent_windows<<<(e->N + 255) / 256, 256, 0, e->stream>>>(p0, e->window_d);
    }
    CK(cudaGetLastError());
    e->window_dirty = false;
  }
  e->range_begin = t_begin;
  for (int tb = t_begin; tb < t_end; tb += S) {
    rc = begin_chunk(e);
    if (rc) return rc;
    Params p = make_params(e);
    int steps = std::min(S, t_end - tb);
    int rows = steps + 1;  // + halo row for step tb-1
    CK(cudaMemsetAsync(e->cells, 0, sizeof(uint32_t) * (size_t)e->cells_stride * steps, e->stream));
    CK(cudaMemsetAsync(e->chg_count, 0, sizeof(uint32_t) * 32 * steps, e->stream));
    {
      Timed tm(e, CITAM_K_EXPAND_BIN);
      // one thread walks an agent through `rpt` consecutive steps: as many as possible (one
      // segment search per agent) while the launch still fills the machine
      int rpt = 16;
      while (rpt < rows && (long long)e->N * ((rows + 2 * rpt - 1) / (2 * rpt)) >= 600000) rpt *= 2;
      dim3 g((e->N + 127) / 128, (rows + rpt - 1) / rpt);
      k_expand_bin<<<g, 128, 0, e->stream>>>(p, tb - 1, rows, 1, rpt);
    }
    CK(cudaGetLastError());
    rc = run_sorted_stages(e, tb, steps);
    if (rc) return rc;
    e->agent_steps += (unsigned long long)steps * e->N;
    rc = maybe_grow_pairs(e);
    if (rc) return rc;
  }
  e->prev_step = LLONG_MIN;
  return join_side_stream(e);
}

int citam_step_states(citam_engine* e, int32_t step, const citam_agent_state* states) {
  if (!e || !states) return fail(CITAM_ERR_INVALID, "null argument");
  if (step < 0) return fail(CITAM_ERR_INVALID, "negative step");
  CK(cudaSetDevice(e->device));
  int rc = ensure_ready(e, 1);
  if (rc) return rc;
  size_t n = (size_t)e->N;
  int g = (e->N + 255) / 256;
  rc = begin_chunk(e);
  if (rc) return rc;
  struct Flag { bool& f; Flag(bool& x) : f(x) { f = true; } ~Flag() { f = false; } } flag(e->explicit_states);
  CK(cudaMemcpyAsync(e->states_d, states, sizeof(citam_agent_state) * n, cudaMemcpyHostToDevice, e->stream));
  if (e->prev_step == (long long)step - 1) {
    CK(cudaMemcpyAsync(e->rec, e->prev, sizeof(int4) * n, cudaMemcpyDeviceToDevice, e->stream));
  } else {
    k_fill_inactive<<<g, 256, 0, e->stream>>>(e->rec, e->N);
    e->launches++;
  }
  k_states_to_rec<<<g, 256, 0, e->stream>>>(e->states_d, e->N, e->rec + n);
  e->launches++;
  if (e->prev_step != (long long)step - 1) e->range_begin = step;
  CK(cudaMemsetAsync(e->cells, 0, sizeof(uint32_t) * (size_t)e->cells_stride, e->stream));
  Params p = make_params(e);
  {
    Timed tm(e, CITAM_K_EXPAND_BIN);
    k_bin_states<<<dim3(g, 1), 256, 0, e->stream>>>(p, 2);
  }
  CK(cudaGetLastError());
  rc = run_sorted_stages(e, step, 1);
  if (rc) return rc;
  CK(cudaMemcpyAsync(e->prev, e->rec + n, sizeof(int4) * n, cudaMemcpyDeviceToDevice, e->stream));
  e->prev_step = step;
  e->agent_steps += e->N;
  rc = join_side_stream(e);
  if (rc) return rc;
  rc = maybe_grow_pairs(e);
  if (rc) return rc;
  return check_overflow(e);
}

int citam_states_at(citam_engine* e, int32_t step, citam_agent_state* out) {
  if (!e || !out) return fail(CITAM_ERR_INVALID, "null argument");
  if (!e->have_itin) return fail(CITAM_ERR_STATE, "no itineraries loaded");
  CK(cudaSetDevice(e->device));
  int rc = ensure_ready(e, 1);
  if (rc) return rc;
  Params p = make_params(e);
  {
    Timed tm(e, CITAM_K_MISC);
    dim3 g((e->N + 127) / 128, 1);
    k_expand_bin<<<g, 128, 0, e->stream>>>(p, step, 1, 0, 16);
    k_rec_to_states<<<(e->N + 255) / 256, 256, 0, e->stream>>>(e->rec, e->N, e->states_d);
  }
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, e->states_d, sizeof(citam_agent_state) * (size_t)e->N, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  e->prev_step = LLONG_MIN;
  return CITAM_OK;
}

int citam_states_range(citam_engine* e, int32_t t_begin, int32_t t_end, citam_agent_state* out) {
  if (!e || !out) return fail(CITAM_ERR_INVALID, "null argument");
  if (!e->have_itin) return fail(CITAM_ERR_STATE, "no itineraries loaded");
  if (t_begin < 0 || t_end < t_begin) return fail(CITAM_ERR_INVALID, "bad step range");
  if (t_end == t_begin) return CITAM_OK;
  CK(cudaSetDevice(e->device));
  int rc = ensure_ready(e, 1);
  if (rc) return rc;
  const int rows_max = std::max(1, std::min(e->S + 1, t_end - t_begin));
  size_t n = (size_t)e->N;
  citam_agent_state* tmp = nullptr;
  CK(cudaMalloc(&tmp, sizeof(citam_agent_state) * n * rows_max));
  Params p = make_params(e);
  for (int tb = t_begin; tb < t_end; tb += rows_max) {
    int rows = std::min(rows_max, t_end - tb);
    {
      Timed tm(e, CITAM_K_MISC);
      const int kRowsPerThread = 16;
      dim3 g((e->N + 127) / 128, (rows + kRowsPerThread - 1) / kRowsPerThread);
      k_expand_bin<<<g, 128, 0, e->stream>>>(p, tb, rows, 0, kRowsPerThread);
      size_t tot = n * rows;
      k_rec_to_states<<<(unsigned)((tot + 255) / 256), 256, 0, e->stream>>>(e->rec, (long long)tot, tmp);
    }
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out + (size_t)(tb - t_begin) * n, tmp, sizeof(citam_agent_state) * n * rows, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
  }
  cudaFree(tmp);
  e->prev_step = LLONG_MIN;
  return CITAM_OK;
}

int citam_get_counters(citam_engine* e, citam_counters* out) {
  if (!e || !out) return fail(CITAM_ERR_INVALID, "null argument");
  CK(cudaSetDevice(e->device));
  DevCounters c;
  CK(cudaMemcpyAsync(&c, e->ctr_d, sizeof c, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  out->agent_steps = e->agent_steps;
  out->pair_tests = c.pair_tests;
  out->contact_steps = c.contact_steps;
  out->kernel_launches = e->launches;
  out->pair_slots_used = pairs_used(c);
  out->coord_slots_used = c.coord_used;
  out->log_records = std::min<unsigned long long>(c.log_count, e->log_cap);
  out->overflow = c.overflow;
  out->contact_runs = c.contact_runs;
  return CITAM_OK;
}

int citam_pair_count(citam_engine* e, uint64_t* n) {
  citam_counters c;
  int rc = citam_get_counters(e, &c);
  if (rc) return rc;
  rc = check_overflow(e);
  if (rc) return rc;
  *n = c.pair_slots_used;
  return CITAM_OK;
}

static int dense_coord_count(citam_engine* e, uint64_t* n) {
  CK(cudaMemsetAsync(&e->ctr_d->export_cursor, 0, sizeof(unsigned long long), e->stream));
  if (e->hist_total > 0) {
    Timed tm(e, CITAM_K_FINALIZE);
    k_hist_count<<<(unsigned)((e->hist_total + 255) / 256), 256, 0, e->stream>>>(e->hist, e->hist_total, e->ctr_d);
  }
  CK(cudaGetLastError());
  unsigned long long v = 0;
  CK(cudaMemcpyAsync(&v, &e->ctr_d->export_cursor, sizeof v, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  *n = v;
  return CITAM_OK;
}

int citam_coord_count(citam_engine* e, uint64_t* n) {
  if (!e || !n) return fail(CITAM_ERR_INVALID, "null argument");
  CK(cudaSetDevice(e->device));
  if (e->hist) return dense_coord_count(e, n);
  citam_counters c;
  int rc = citam_get_counters(e, &c);
  if (rc) return rc;
  *n = c.coord_slots_used;
  return CITAM_OK;
}

int citam_export_pairs(citam_engine* e, citam_pair* out, uint64_t capacity, uint64_t* n) {
  if (!e || !n || (capacity && !out)) return fail(CITAM_ERR_INVALID, "null argument");
  int rc = check_overflow(e);
  if (rc) return rc;
  uint64_t have = 0;
  rc = citam_pair_count(e, &have);
  if (rc) return rc;
  *n = have;
  if (have == 0) return CITAM_OK;
  if (capacity < have) return fail(CITAM_ERR_CAPACITY, "output holds %llu pairs, %llu needed", (unsigned long long)capacity, (unsigned long long)have);
  citam_pair* tmp = nullptr;
  CK(cudaMalloc(&tmp, sizeof(citam_pair) * have));
  CK(cudaMemsetAsync(&e->ctr_d->export_cursor, 0, sizeof(unsigned long long), e->stream));
  {
    Timed tm(e, CITAM_K_FINALIZE);
    k_compact_pairs<<<(unsigned)((e->pair_cap + 255) / 256), 256, 0, e->stream>>>(e->pairs, e->pair_cap, tmp, have, e->ctr_d);
  }
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, tmp, sizeof(citam_pair) * have, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  cudaFree(tmp);
  // first-contact order = dict insertion order of the reference (contacts.py:135-142;
  // pairs of one step are visited row-major, close_contacts_calculator.py:104-124)
  std::sort(out, out + have, [](const citam_pair& a, const citam_pair& b) {
    if (a.first_step != b.first_step) return a.first_step < b.first_step;
    if (a.a1 != b.a1) return a.a1 < b.a1;
    return a.a2 < b.a2;
  });
  return CITAM_OK;
}

int citam_export_coords(citam_engine* e, citam_coord* out, uint64_t capacity, uint64_t* n) {
  if (!e || !n || (capacity && !out)) return fail(CITAM_ERR_INVALID, "null argument");
  int rc = check_overflow(e);
  if (rc) return rc;
  uint64_t have = 0;
  rc = citam_coord_count(e, &have);
  if (rc) return rc;
  *n = have;
  if (have == 0) return CITAM_OK;
  if (capacity < have) return fail(CITAM_ERR_CAPACITY, "output holds %llu coords, %llu needed", (unsigned long long)capacity, (unsigned long long)have);
  citam_coord* tmp = nullptr;
  CK(cudaMalloc(&tmp, sizeof(citam_coord) * have));
  CK(cudaMemsetAsync(&e->ctr_d->export_cursor, 0, sizeof(unsigned long long), e->stream));
  if (e->hist) {
    Timed tm(e, CITAM_K_FINALIZE);
    for (int f = 0; f < e->F; ++f) {
      const FloorDev& fd = e->floors_h[f];
      long long nb = (long long)fd.hist_w * fd.hist_h;
      if (nb == 0) continue;
      k_hist_compact<<<(unsigned)((nb + 255) / 256), 256, 0, e->stream>>>(e->hist + fd.hist_base, nb, f, fd.hist_w,
                                                                          2 * fd.minx, 2 * fd.miny, tmp, have, e->ctr_d);
    }
  } else {
    Timed tm(e, CITAM_K_FINALIZE);
    k_compact_coords<<<(unsigned)((e->coord_cap + 255) / 256), 256, 0, e->stream>>>(e->coords, e->coord_cap, tmp, have, e->ctr_d);
  }
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, tmp, sizeof(citam_coord) * have, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  cudaFree(tmp);
  std::sort(out, out + have, [](const citam_coord& a, const citam_coord& b) {
    if (a.floor != b.floor) return a.floor < b.floor;
    if (a.sx != b.sx) return a.sx < b.sx;
    return a.sy < b.sy;
  });
  return CITAM_OK;
}

int citam_export_agent_durations(citam_engine* e, uint64_t* out) {
  if (!e || !out) return fail(CITAM_ERR_INVALID, "null argument");
  CK(cudaSetDevice(e->device));
  int rc = check_overflow(e);
  if (rc) return rc;
  CK(cudaMemcpyAsync(out, e->agent_dur, sizeof(uint64_t) * (size_t)e->N, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  return CITAM_OK;
}

int citam_export_contact_log(citam_engine* e, citam_contact* out, uint64_t capacity, uint64_t* n) {
  if (!e || !n || (capacity && !out)) return fail(CITAM_ERR_INVALID, "null argument");
  int rc = check_overflow(e);
  if (rc) return rc;
  citam_counters c;
  rc = citam_get_counters(e, &c);
  if (rc) return rc;
  *n = c.log_records;
  if (c.log_records == 0) return CITAM_OK;
  if (capacity < c.log_records) return fail(CITAM_ERR_CAPACITY, "output holds %llu records, %llu needed", (unsigned long long)capacity, (unsigned long long)c.log_records);
  CK(cudaMemcpyAsync(out, e->log, sizeof(citam_contact) * c.log_records, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  // deterministic order: (step, a1, a2) == the order the reference visits contacts
  std::sort(out, out + c.log_records, [](const citam_contact& a, const citam_contact& b) {
    if (a.step != b.step) return a.step < b.step;
    if (a.a1 != b.a1) return a.a1 < b.a1;
    return a.a2 < b.a2;
  });
  return CITAM_OK;
}

int citam_clear_contact_log(citam_engine* e) {
  if (!e) return fail(CITAM_ERR_INVALID, "null engine");
  CK(cudaSetDevice(e->device));
  CK(cudaMemsetAsync(&e->ctr_d->log_count, 0, sizeof(unsigned long long), e->stream));
  return CITAM_OK;
}

int citam_statistics(citam_engine* e, citam_stats* out) {
  if (!e || !out) return fail(CITAM_ERR_INVALID, "null argument");
  int rc = check_overflow(e);
  if (rc) return rc;
  size_t n = (size_t)e->N;
  CK(cudaMemsetAsync(e->agent_nev, 0, 4 * n, e->stream));
  CK(cudaMemsetAsync(e->agent_has, 0, 4 * n, e->stream));
  CK(cudaMemsetAsync(e->ctr_d->stats, 0, sizeof(unsigned long long) * 6, e->stream));
  {
    Timed tm(e, CITAM_K_FINALIZE);
    k_stats_pairs<<<(unsigned)((e->pair_cap + 255) / 256), 256, 0, e->stream>>>(e->pairs, e->pair_cap, e->agent_nev, e->agent_has, e->ctr_d);
    k_stats_agents<<<(e->N + 255) / 256, 256, 0, e->stream>>>(e->agent_nev, e->agent_has, e->N, e->ctr_d);
  }
  CK(cudaGetLastError());
  DevCounters c;
  CK(cudaMemcpyAsync(&c, e->ctr_d, sizeof c, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  out->total_duration = c.stats[0];
  out->n_pairs = c.stats[1];
  out->n_agents_with_contacts = c.stats[2];
  out->sum_events_per_agent = c.stats[3];
  out->max_events_per_agent = c.stats[4];
  out->reserved = 0;
  return CITAM_OK;
}

int citam_get_timings(citam_engine* e, citam_timings* out) {
  if (!e || !out) return fail(CITAM_ERR_INVALID, "null argument");
  CK(cudaSetDevice(e->device));
  drain_timings(e);
  for (int k = 0; k < CITAM_N_KERNELS; ++k) { out->ms[k] = e->t_ms[k]; out->launches[k] = e->t_n[k]; }
  return CITAM_OK;
}

int citam_agent_durations_device(citam_engine* e, void** ptr, uint64_t* n_elements) {
  if (!e || !ptr || !n_elements) return fail(CITAM_ERR_INVALID, "null argument");
  CK(cudaSetDevice(e->device));
  CK(cudaStreamSynchronize(e->stream));
  *ptr = e->agent_dur;
  *n_elements = (uint64_t)e->N;
  return CITAM_OK;
}

int citam_merge_pairs(citam_engine* e, const citam_pair* pairs, uint64_t n) {
  if (!e || (n && !pairs)) return fail(CITAM_ERR_INVALID, "null argument");
  if (!n) return CITAM_OK;
  CK(cudaSetDevice(e->device));
  for (uint64_t i = 0; i < n; ++i)
    if (pairs[i].a1 >= pairs[i].a2 || pairs[i].a2 >= (uint32_t)e->N) return fail(CITAM_ERR_INVALID, "bad pair record %llu", (unsigned long long)i);
  citam_pair* tmp = nullptr;
  CK(cudaMalloc(&tmp, sizeof(citam_pair) * n));
  CK(cudaMemcpyAsync(tmp, pairs, sizeof(citam_pair) * n, cudaMemcpyHostToDevice, e->stream));
  Params p = make_params(e);
  {
    Timed tm(e, CITAM_K_FINALIZE);
    k_merge_pairs<<<(unsigned)((n + 255) / 256), 256, 0, e->stream>>>(p, tmp, n);
  }
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(e->stream));
  cudaFree(tmp);
  return check_overflow(e);
}

int citam_merge_coords(citam_engine* e, const citam_coord* coords, uint64_t n) {
  if (!e || (n && !coords)) return fail(CITAM_ERR_INVALID, "null argument");
  if (!n) return CITAM_OK;
  CK(cudaSetDevice(e->device));
  int rc0 = ensure_ready(e, 1);
  if (rc0) return rc0;
  for (uint64_t i = 0; i < n; ++i) {
    const citam_coord& c = coords[i];
    if (c.floor < 0 || c.floor >= e->F) return fail(CITAM_ERR_INVALID, "bad coordinate record %llu", (unsigned long long)i);
    const FloorDev& fd = e->floors_h[c.floor];
    if (e->hist && (c.sx < 2 * fd.minx || c.sx - 2 * fd.minx >= fd.hist_w || c.sy < 2 * fd.miny || c.sy - 2 * fd.miny >= fd.hist_h))
      return fail(CITAM_ERR_INVALID, "coordinate record %llu lies outside floor %d", (unsigned long long)i, c.floor);
  }
  citam_coord* tmp = nullptr;
  CK(cudaMalloc(&tmp, sizeof(citam_coord) * n));
  CK(cudaMemcpyAsync(tmp, coords, sizeof(citam_coord) * n, cudaMemcpyHostToDevice, e->stream));
  Params p = make_params(e);
  {
    Timed tm(e, CITAM_K_FINALIZE);
    k_merge_coords<<<(unsigned)((n + 255) / 256), 256, 0, e->stream>>>(p, tmp, n);
  }
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(e->stream));
  cudaFree(tmp);
  return check_overflow(e);
}

}  // extern "C"
