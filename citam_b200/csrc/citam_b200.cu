// citam_b200.cu — B200 (sm_100a) implementation of CITAM's per-timestep loop.
//
// What it replaces in the reference (paths relative to the upstream repo):
//   Agent.step + Schedule.get_next_position   citam/engine/core/agent.py:67-109,
//                                             citam/engine/schedulers/office_schedule.py:632-638
//   Floorplan.identify_this_location          citam/engine/map/floorplan.py:215-241 (+ space.py, geometry.py)
//   CloseContactsCalculator.run               citam/engine/calculators/close_contacts_calculator.py:64-231
//   ContactEvents.add_contact / statistics    citam/engine/calculators/contacts.py:92-294
//
// Two step pipelines; every accumulator is commutative, so steps, chunks and time ranges are independent.
//
// citam_run without the contact log -- the SPLIT pipeline, per chunk of S <= 495 steps (one launch per
// stage covers all steps of the chunk):
//   k_classify        per agent: out / static (stands in one zero-velocity segment from the halo step to
//                     the chunk's end: binned ONCE, fine grid) / dynamic (dense index)
//   k_classify_dyn    per dynamic agent: its stays of >= 16 steps inside the chunk become entries of the
//                     same static array, valid for [since, until); the 8-row groups that still hold a
//                     binned step go into per-group work lists
//   k_scan(+_add)     exclusive scans of the static cell histogram and of the per-step dynamic ones
//   k_scatter_static  counting-sort scatter of the static array
//   k_expand_dyn      work-list entries -> per-step 16-byte records {xy, until, state word, rank} of the
//                     steps no stay covers; one atomic per binned record returns its rank in its coarse cell
//   k_scatter_dyn, k_select_changed_dyn   per-step counting-sort scatter; the changed agents of each step
//   k_pairs           first step of a citam_run RANGE only: static-static pairs, half neighbourhood,
//                     candidate window in shared memory
//   k_pairs_changed_split   every step: only changed agents, against 3x3 fine cells of the static array and
//                     3x3 coarse cells of the step's dynamic array, tests flattened over a 128-agent batch;
//                     a contact of two unchanged agents is inside a run that an earlier step added in one
//                     piece (stationary-run aggregation: ONE table update per contact run, however many
//                     chunks the run crosses)
//   k_accumulate      (side stream) contact-run records -> 16-byte pair-table slot probe + one reduction,
//                     one 64-bit reduction into the dense per-coordinate histogram, two reductions into
//                     the L2-resident per-agent words (contact seconds | contact events << 40)
// With the contact log (full-fidelity writers), for citam_step_states and under CITAM_NO_SPLIT=1 every
// agent is expanded and binned at every step: k_expand_bin, k_scan, k_scatter, k_select_changed,
// k_pairs / k_pairs_changed, k_accumulate.
// Finalize (K5): compaction of the tables + device radix sort (radix_sort.cuh) into the reference's
// output order; partition by key hash for the multi-GPU all-to-all; merge of received records;
// statistics from the per-agent words (O(n_agents), no pass over the pair table).
// The distance predicate is the reference's float64 sqrt(dx^2+dy^2) < D as an exact integer
// threshold (find_s_max).  All of this is HBM/L2-bound integer work; there is no GEMM in the path.

#include "../../include/citam_b200.h"
#include "radix_sort.cuh"

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
  int32_t dncx, dncy, dcell_base;  // the coarse grid (cells of kDynCellMul x cell) the walking agents are binned in
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
  uint32_t t_run;    // step - chunk_begin (12 bits) | (run - 1) << 12 (run <= 2^20 steps)
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
  unsigned long long merged_max_step;  // largest first-contact step of a merged record
  unsigned long long binned;           // (agent, step) records binned per step
  unsigned long long cbuf_count[2][kSubBuffers * kCounterStride];  // one set per contact buffer
  unsigned long long pair_used_sub[kSubBuffers * kCounterStride];  // spread: inserts are frequent
};

struct Params {
  int32_t N, F;
  int32_t cell;
  int32_t total_cells;
  int32_t cells_stride;  // uint32 per step row of the histogram (>= total_cells+1, multiple of 4)
  uint32_t cell_magic;  // floor(2^32 / cell) + 1: x / cell == __umulhi(x, cell_magic) for x < 2^16 (cell > 1)
  uint32_t d_lim;       // floor(sqrt(s_max)): |dx|, |dy| beyond it cannot be in contact
  double D;
  unsigned long long s_lim;  // s_max + 1, where s_max is the largest integer s with sqrt_rn((double)s) < D (0: none)
  const FloorDev* floors;
  const int4* segs;  // {t0 | floor<<24, x0, y0, dx&0xffff | dy<<16}
  const long long* seg_off;
  const int2* window;  // per agent {first step in the facility, first step of a final stay outside every space}
                       // (INT_MAX when there is none); NULL = not available
  int4* rec;        // [(S+1)][N] row 0 = halo step t_begin-1.  Located agent: {xy, until, state word, rank in
                    // cell}; in the facility but in no space: {x, y, state word, 0}; else {0, 0, state word, 0}
  uint32_t* cells;  // [S][cells_stride]
  int4* sorted;     // [S][N] {xy, agent, state word, until} in cell order
  PairEntry* pairs;
  unsigned long long pair_mask;
  CoordEntry* coords;     // hashed histogram (lattices too large for the dense one)
  unsigned long long coord_mask;
  unsigned long long* hist;  // dense histogram, NULL = use the hash
  int32_t range_begin;       // first step of the current citam_run range
  int32_t range_end;         // its end: contact runs are clipped here
  int32_t chunk_begin;       // first step of the chunk being processed
  uint4* log;                // {step, a1, a2, 0}
  unsigned long long log_cap;
  ContactRec* cbuf;  // per-chunk contact buffer, kSubBuffers equal parts (NULL: accumulate inline)
  unsigned long long cbuf_sub_cap;
  unsigned long long* cbuf_count;  // its kSubBuffers counters (stride kCounterStride)
  unsigned long long* agent_dur;   // [N] contact-steps (40 bits) | contact events << 40 per agent
  uint32_t* chg_list;   // [S][N] located agents whose state may differ from the step before
  uint32_t* chg_count;  // [S] x 32 (one counter per 128 bytes); NULL = lists not kept
  // split pipeline (citam_run without the contact log): per chunk, an agent that stands in ONE
  // zero-velocity segment from the halo step to the chunk's end is "static" -- binned once, in
  // `scells`/`ssorted` (fine grid); everything else that is around is "dynamic": expanded, binned
  // (coarse grid, rows of `cells`) and sorted per step (`rec`/`sorted` rows are then indexed by the
  // dynamic index, not by the agent).  cls == NULL: every agent is expanded at every step.
  int32_t dcell;            // coarse cell edge
  uint32_t dcell_magic;
  int32_t dtotal_cells, dcells_stride;
  uint32_t* cls;            // [N] dynamic index | kClsStatic | kClsOut
  int4* srec;               // [N] a static agent's record {xy, until, state word, rank in cell}
  int4* ssorted;            // static agents in cell order {xy, agent, state word, until}
  uint32_t* scells;         // [cells_stride] static cell counts -> starts
  uint32_t* dyn_agent;      // [N] agent of dynamic index j
  uint32_t* dyn_count;      // dynamic agents of this chunk
  uint32_t* seg_cur;        // [N] per agent: 1 + index (inside its segments) of the segment k_classify last found
  // still intervals of DYNAMIC agents: a stay of >= kStillMin steps inside the chunk is one more entry
  // of the static array, valid from `since` (state word bits 16-23: since - (chunk_begin - 1)) until
  // `until`; the agent is then binned per step only at the steps no such entry covers
  int4* ient;               // [(S / kStillMin + 1) * N] {xy, agent, state word | since << 16, until}
  uint32_t* ient_rank;      // rank of the entry in its fine cell
  uint32_t* ient_count;
  uint8_t* dmask;           // [S][N] by dynamic index: 1 = the agent is binned at this step
  uint32_t* work;           // [kMaxGroups][N] per group of 8 rows: the dynamic agents (by dynamic index) with a step in it
                            // at which they are binned, or the record before one (k_classify_dyn); every other
                            // row group of an agent is skipped by every kernel
  uint32_t* work_count;     // [kMaxGroups]
  int32_t chunk_steps;      // steps of the chunk being processed
  int32_t still_min;        // shortest stay that becomes an entry (>= kStillMin; larger than a chunk: none does)
  DevCounters* ctr;
};
constexpr uint32_t kClsOut = 0xFFFFFFFFu, kClsStatic = 0xFFFFFFFEu;
constexpr int kStillMin = 16;       // shortest stay (steps inside the chunk) that becomes a static-array entry
constexpr int kSplitMaxChunk = 495;  // `since` has 9 bits; 62 row groups of 8 rows hold the halo row + 495 steps
constexpr int kMaxGroups = 64;
constexpr int kSplitDefaultChunk = 480;
// a static-array entry {xy, agent, state word | since << 16, until} is a candidate at row `row` (step t) iff
// (`since` = first row of the stay, 9 bits: state-word bits 16-23 and bit 30, which static entries do not use otherwise)
__device__ __forceinline__ int entry_since(int32_t z) { return ((z >> 16) & 0xff) | (((z >> 30) & 1) << 8); }
__device__ __forceinline__ int32_t pack_since(int since) { return ((since & 0xff) << 16) | ((since >> 8) << 30); }
__device__ __forceinline__ bool entry_valid(const int4& e, int row, int t) { return entry_since(e.z) <= row + 1 && t < e.w; }
__device__ __forceinline__ unsigned long long bits_range(int a, int b) {  // bits a..b (0 <= a <= b <= 63)
  return (b >= 63 ? ~0ULL : ((1ULL << (b + 1)) - 1ULL)) & ~((1ULL << a) - 1ULL);
}
// steps of [tb, tb + S) at which an agent standing in the zero-velocity segment [t0, next_t0) is unchanged
__device__ __forceinline__ int still_len(int t0, int next_t0, int tb, int S) { return min(next_t0, tb + S) - max(t0 + 1, tb); }
constexpr int kDynCellMul = 4;

#define OVF_PAIR 1ULL
#define OVF_COORD 2ULL
#define OVF_LOG 4ULL
#define OVF_GRID 8ULL
#define OVF_FIELD 16ULL

// state word: bits 0-14 space index (0x7fff = none); bit 15 "changed": the agent is NOT known to
// be exactly where (and what) it was one step earlier (it moves, or this is the first step of its
// segment); bits 24-29 floor; bit 30 moving; bit 31 = not in the facility.
// xy: lattice-relative position of a LOCATED agent, (x - minx) | (y - miny) << 16 (an agent has a
// space only inside its floor's location table, whose sides are <= 65536).
// until: the first step at which the agent's state may differ from the current one -- the start of
// its next segment for an agent standing in a zero-velocity segment, t + 1 otherwise.  changed and
// until come from the segment store alone, so they agree with each other: for a still agent
// changed(t') == 0 for t < t' < until(t), whatever chunk t' falls into.
constexpr int kMaxChunk = 1024;     // steps per launch group (12 bits in ContactRec.t_run: <= 4096)
constexpr int kMaxRunLog2 = 20;     // a contact run is clipped to 2^20 steps: longer citam_run ranges are split
constexpr int kMaxFloors = 63;
constexpr int kAgentDurBits = 40;   // per-agent counter word: contact-steps | events << 40
constexpr unsigned long long kAgentDurMask = (1ULL << kAgentDurBits) - 1ULL;
__device__ __forceinline__ int32_t pack_fl(int floor, int loc, int changed = 1, int moving = 0) {
  if (floor < 0) return (int32_t)0x8000ffffu;
  return (int32_t)(((uint32_t)(moving & 1) << 30) | ((uint32_t)(floor & 0x3f) << 24) |
                   ((uint32_t)(changed & 1) << 15) | ((uint32_t)loc & 0x7fffu));
}
__device__ __forceinline__ int fl_floor(int32_t fl) { return fl < 0 ? -1 : (fl >> 24) & 0x3f; }
__device__ __forceinline__ int fl_loc(int32_t fl) { int v = fl & 0x7fff; return v == 0x7fff ? -1 : v; }
__device__ __forceinline__ bool fl_changed(int32_t fl) { return (fl >> 15) & 1; }
__device__ __forceinline__ bool fl_located(int32_t fl) { return fl >= 0 && (fl & 0x7fff) != 0x7fff; }
__device__ __forceinline__ int xy_x(int32_t xy) { return xy & 0xffff; }
__device__ __forceinline__ int xy_y(int32_t xy) { return (int)((uint32_t)xy >> 16); }

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

// v / cell for a lattice-relative coordinate v < 2^16 (exact: cell_magic = floor(2^32/cell) + 1)
__device__ __forceinline__ int div_cell(const Params& p, int v) {
  return p.cell == 1 ? v : (int)__umulhi((uint32_t)v, p.cell_magic);
}
__device__ __forceinline__ int cell_of_xy(const Params& p, const FloorDev& fd, int32_t xy) {
  return fd.cell_base + div_cell(p, xy_y(xy)) * fd.ncx + div_cell(p, xy_x(xy));
}
__device__ __forceinline__ int div_dcell(const Params& p, int v) { return (int)__umulhi((uint32_t)v, p.dcell_magic); }
__device__ __forceinline__ int dcell_of_xy(const Params& p, const FloorDev& fd, int32_t xy) {
  return fd.dcell_base + div_dcell(p, xy_y(xy)) * fd.dncx + div_dcell(p, xy_x(xy));
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

// Checks the limits the header documents on a freshly loaded segment store (one thread per agent):
// seg_off non-decreasing inside [0, n_segs]; per agent t0 strictly increasing in [0, 2^24 - 2];
// floor in 0..62; the last segment holds still (the sticky tail).  Any violation sets *bad.
__global__ void k_validate_segments(const long long* seg_off, const citam_segment* raw, long long n_segs, int N, int* bad) {
  int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= N) return;
  long long sb = seg_off[a], se = seg_off[a + 1];
  if (sb < 0 || se < sb || se > n_segs) { atomicOr(bad, 1); return; }
  int prev = -1;
  for (long long i = sb; i < se; ++i) {
    citam_segment g = raw[i];
    if (g.t0 <= prev || g.t0 > 0xFFFFFD) { atomicOr(bad, 2); return; }
    if (g.floor < 0 || g.floor >= kMaxFloors) { atomicOr(bad, 4); return; }
    if (g.x0 <= -(1 << 26) || g.x0 >= (1 << 26) || g.y0 <= -(1 << 26) || g.y0 >= (1 << 26)) { atomicOr(bad, 16); return; }
    prev = g.t0;
    if (i == se - 1 && (g.dx != 0 || g.dy != 0)) { atomicOr(bad, 8); return; }
  }
}

// --------------------------------------------------------------------------
// K1a: itinerary expansion + grid histogram.
// One thread walks one agent through kRowsPerThread consecutive steps: the segment cursor
// advances monotonically, the location table is touched only when the position changes, and
// every 16-byte record store is coalesced over the 32 agents of a warp.
// --------------------------------------------------------------------------

// `slot`: the agent's index inside a record row (the agent itself, or its dynamic index);
// `coarse`: bin into the coarse per-step grid of the split pipeline.
__device__ __forceinline__ void expand_rows(const Params& p, int a, size_t slot, int t_first, int rows, int r0, int do_bin,
                                            int kRowsPerThread, bool coarse) {
  int2 win = make_int2(INT_MIN, INT_MAX);
  if (do_bin && p.window != nullptr) {
    win = p.window[a];
    // nothing of this thread's steps lies in [enter - 1, exit): no record is written (none is read);
    // the split pipeline's rows say "not around" instead, so that its scatter needs no window
    if (win.x > t_first + min(r0 + kRowsPerThread, rows) || win.y <= t_first + r0) {
      if (coarse)
        for (int row = r0; row < min(r0 + kRowsPerThread, rows); ++row) {
          p.rec[(size_t)row * p.N + slot] = make_int4(0, 0, pack_fl(-1, -1), 0);
          if (row >= 1) p.dmask[(size_t)(row - 1) * p.N + slot] = 0;
        }
      return;
    }
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
  int4 s = make_int4(0, 0, 0, 0), nxs = make_int4(0, 0, 0, 0);  // the current segment and the one after it
  int next_t0 = INT_MAX;
  if (c >= sb) s = __ldg(&p.segs[c]);
  if (c + 1 < se) { nxs = __ldg(&p.segs[c + 1]); next_t0 = nxs.x & 0xffffff; }
  int lastx = INT_MIN, lasty = INT_MIN, lastf = -1, loc = -1;
  const int stride = coarse ? p.dcells_stride : p.cells_stride;
  // rows in groups of kGroup: the group's records are computed first, then its rank atomics are issued
  // together and its stores follow -- a thread waits once per group for an atomic's round trip, not
  // once per row (the kernel is bound by that latency)
  constexpr int kGroup = 4;
  const int row_end = min(r0 + kRowsPerThread, rows);
#pragma unroll 1
  for (int rg = r0; rg < row_end; rg += kGroup) {
    int4 out[kGroup];
    int cellv[kGroup];  // >= 0: cell to take a rank in; -1: store as is; -2: no store
    bool live[kGroup];  // a row of the chunk
#pragma unroll
    for (int k = 0; k < kGroup; ++k) {
      const int row = rg + k;
      cellv[k] = -2;
      live[k] = row < row_end;
      if (!live[k]) continue;
      t = t_first + row;
      while (t >= next_t0) {  // one 16-byte load per segment, a segment ahead of its use
        ++c;
        s = nxs;
        next_t0 = INT_MAX;
        if (c + 1 < se) { nxs = __ldg(&p.segs[c + 1]); next_t0 = nxs.x & 0xffffff; }
      }
      out[k] = make_int4(0, 0, pack_fl(-1, -1), 0);
      if (t + 1 < win.x || t >= win.y) {  // outside the agent's window (see k_agent_windows)
        if (coarse) cellv[k] = -1;
        continue;
      }
      cellv[k] = -1;
      if (c < sb || t < 0) continue;
      const int dt = t - (s.x & 0xffffff);
      const int f = (s.x >> 24) & 0xff;
      const int x = s.y + (int)(int16_t)(s.w & 0xffff) * dt;
      const int y = s.z + (int)(int16_t)((uint32_t)s.w >> 16) * dt;
      if (x != lastx || y != lasty || f != lastf) {
        loc = (f < p.F) ? lut_lookup(p.floors[f], x, y) : -1;
        lastx = x; lasty = y; lastf = f;
      }
      // stationary inside one segment: same state as the step before, and until the next segment starts
      const bool still = s.w == 0;
      const int changed = !(still && dt >= 1);
      if (loc < 0) {  // in the facility, in no space: never binned, never in contact
        out[k] = make_int4(x, y, pack_fl(f, -1, changed, still ? 0 : 1), 0);
        continue;
      }
      const FloorDev& fd = p.floors[f];
      const int32_t xy = (x - fd.minx) | ((y - fd.miny) << 16);
      out[k] = make_int4(xy, still ? next_t0 : t + 1, pack_fl(f, loc, changed, still ? 0 : 1), 0);
      if (do_bin && row >= 1) {
        if (!coarse) {
          cellv[k] = cell_of_xy(p, fd, xy);
        } else if (!changed && still_len(s.x & 0xffffff, next_t0, p.chunk_begin, p.chunk_steps) >= p.still_min) {
          // inside a stay that k_classify entered into the static array: not binned at this step; the
          // record itself is needed only as the previous-step record of the stay's last step + 1
          cellv[k] = (t + 1 >= next_t0) ? -1 : -3;
        } else {
          cellv[k] = dcell_of_xy(p, fd, xy);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < kGroup; ++k)
      if (cellv[k] >= 0) out[k].w = (int)atomicAdd(&p.cells[(size_t)(rg + k - 1) * stride + cellv[k]], 1u);
#pragma unroll
    for (int k = 0; k < kGroup; ++k) {
      if (cellv[k] > -2) p.rec[(size_t)(rg + k) * p.N + slot] = out[k];
      if (coarse && live[k] && rg + k >= 1) p.dmask[(size_t)(rg + k - 1) * p.N + slot] = cellv[k] >= 0;
    }
  }
}

__global__ void __launch_bounds__(128) k_expand_bin(Params p, int t_first, int rows, int do_bin, int kRowsPerThread) {
  int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= p.N) return;
  expand_rows(p, a, (size_t)a, t_first, rows, blockIdx.y * kRowsPerThread, do_bin, kRowsPerThread, false);
}

// ---- split pipeline -------------------------------------------------------------------------
// Class of every agent for the chunk of S steps starting at tb (records cover tb-1 .. tb+S-1):
//   static   one zero-velocity segment holds from tb-1 to the chunk's end, inside a space: its state
//            is the same at every step (changed == 0 throughout): ONE record, one grid cell rank
//   dynamic  anything else whose window meets the chunk: gets the next dynamic index
//   out      not around (or standing outside every space for the whole chunk)
__global__ void __launch_bounds__(256) k_classify(Params p, int tb, int S) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t c = kClsOut;
  bool dyn = false;
  if (a < p.N) {
    const int2 win = p.window[a];
    if (!(win.x > tb + S || win.y <= tb - 1)) {
      const int t = tb - 1;
      const long long sb = p.seg_off[a], se = p.seg_off[a + 1];
      // the segment holding step t: chunks mostly follow one another in time, so the segment found for
      // the previous chunk (seg_cur: its index + 1, 0 = before the first) or one of the next few is
      // nearly always the one; otherwise a binary search
      long long lo = sb + (long long)p.seg_cur[a];  // candidate for "first segment with t0 > t"
      int4 s = make_int4(0, 0, 0, 0);
      int next_t0 = INT_MAX;
      bool found = false;
      if (lo <= se) {
        bool below = true;
        if (lo > sb) { s = __ldg(&p.segs[lo - 1]); below = (s.x & 0xffffff) <= t; }
        if (below) {
          for (int k = 0; k < 4; ++k) {
            if (lo >= se) { next_t0 = INT_MAX; found = true; break; }
            const int4 nx = __ldg(&p.segs[lo]);
            if ((nx.x & 0xffffff) > t) { next_t0 = nx.x & 0xffffff; found = true; break; }
            s = nx;
            ++lo;
          }
        }
      }
      if (!found) {
        long long hi = se;
        lo = sb;
        while (lo < hi) {
          long long mid = (lo + hi) >> 1;
          int t0 = __ldg(&p.segs[mid].x) & 0xffffff;
          if (t0 <= t) lo = mid + 1; else hi = mid;
        }
        if (lo > sb) s = __ldg(&p.segs[lo - 1]);
        next_t0 = (lo < se) ? (__ldg(&p.segs[lo].x) & 0xffffff) : INT_MAX;
      }
      p.seg_cur[a] = (uint32_t)(lo - sb);
      long long ci = lo - 1;
      dyn = true;
      if (ci >= sb && t >= 0 && s.w == 0 && next_t0 >= tb + S) {
        dyn = false;
        const int f = (s.x >> 24) & 0xff;
        const int loc = (f < p.F) ? lut_lookup(p.floors[f], s.y, s.z) : -1;
        if (loc >= 0) {
          const FloorDev& fd = p.floors[f];
          const int32_t xy = (s.y - fd.minx) | ((s.z - fd.miny) << 16);
          const uint32_t rk = atomicAdd(&p.scells[cell_of_xy(p, fd, xy)], 1u);
          p.srec[a] = make_int4(xy, next_t0, pack_fl(f, loc, 0, 0), (int)rk);
          c = kClsStatic;
        }
      }
    }
  }
  const unsigned m = __ballot_sync(0xffffffffu, dyn);
  if (m) {
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(p.dyn_count, (uint32_t)__popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (dyn) {
      c = base + __popc(m & ((1u << lane) - 1u));
      p.dyn_agent[c] = (uint32_t)a;
    }
  }
  if (a < p.N) p.cls[a] = c;
}

// Second half of the classification, one thread per DYNAMIC agent (dense): its stays of >= still_min
// steps inside the chunk become entries of the static array (k_expand_dyn applies the same rule,
// still_len, to leave those steps out of the per-step arrays); the row groups (8 rows; row = step -
// (tb - 1)) that hold a step at which the agent is binned, or the record before one, go into `need`
// and into the per-group work lists that the per-step kernels walk -- a group that lies inside a stay
// or outside the agent's window is touched by no kernel.
__global__ void __launch_bounds__(256) k_classify_dyn(Params p, int tb, int S) {
  const uint32_t Dn = *p.dyn_count;
  constexpr int gs = 3, G = 8;
  const int groups = (S + 1 + G - 1) >> gs;
  const int lane = threadIdx.x & 31;
  for (uint32_t j0 = blockIdx.x * blockDim.x; j0 < Dn; j0 += gridDim.x * blockDim.x) {  // whole warps iterate together
    const uint32_t j = j0 + threadIdx.x;
    unsigned long long need = 0;
    if (j < Dn) {
      const int a = (int)p.dyn_agent[j];
      const int2 win = p.window[a];
      const long long sb = p.seg_off[a], se = p.seg_off[a + 1];
      need = groups >= 64 ? ~0ULL : (1ULL << groups) - 1ULL;
      {
        // rows with step + 1 < win.x or step >= win.y carry nothing: rows [r_lo, r_hi) matter
        const long long r_lo = (long long)win.x - 1 - (tb - 1), r_hi = (win.y == INT_MAX) ? (1LL << 40) : (long long)win.y - (tb - 1);
        // groups holding a row of [r_lo, r_hi): floor(r_lo / 8) .. floor((r_hi - 1) / 8)
        const long long ga = r_lo <= 0 ? 0 : (r_lo >> gs), gb = r_hi > 8LL * 64 ? 63 : ((r_hi - 1) >> gs);
        need &= (r_hi <= 0 || ga > 63 || gb < ga) ? 0ULL : bits_range((int)ga, (int)gb);
      }
      long long ci = sb + (long long)p.seg_cur[a] - 1;  // the segment holding step tb - 1 (k_classify found it)
      // the walk keeps the next segment in registers: one 16-byte load per segment, issued a segment ahead
      int4 sg = make_int4(0, 0, 0, 0), nx = make_int4(INT_MAX, 0, 0, 0);
      if (ci < sb) ci = sb;
      if (ci < se) {
        sg = __ldg(&p.segs[ci]);
        if (ci + 1 < se) nx = __ldg(&p.segs[ci + 1]);
      }
      while (ci < se) {
        const int t0 = sg.x & 0xffffff;
        const int next_t0 = (ci + 1 < se) ? (nx.x & 0xffffff) : INT_MAX;
        if (t0 >= tb + S) break;
        if (sg.w == 0 && still_len(t0, next_t0, tb, S) >= p.still_min) {
          const int f = (sg.x >> 24) & 0xff;
          const int loc = (f < p.F) ? lut_lookup(p.floors[f], sg.y, sg.z) : -1;
          if (loc >= 0) {
            const FloorDev& fd = p.floors[f];
            const int32_t xy = (sg.y - fd.minx) | ((sg.z - fd.miny) << 16);
            const uint32_t rk = atomicAdd(&p.scells[cell_of_xy(p, fd, xy)], 1u);
            uint32_t e = 0;
            {  // one reservation for the lanes that are here together
              cg::coalesced_group cgp = cg::coalesced_threads();
              if (cgp.thread_rank() == 0) e = atomicAdd(p.ient_count, (uint32_t)cgp.size());
              e = cgp.shfl(e, 0) + cgp.thread_rank();
            }
            const int since = max(t0 + 1, tb) - (tb - 1);
            p.ient[e] = make_int4(xy, a, pack_fl(f, loc, 0, 0) | pack_since(since), next_t0);
            p.ient_rank[e] = rk;
            // covered rows [since, rb); the last one is kept when a step of the chunk follows it (it is
            // that step's previous record)
            const int rb = min(next_t0, tb + S) - (tb - 1);
            const int keep_from = (next_t0 < tb + S) ? rb - 1 : rb;
            // groups with every row in [since, keep_from): ceil(since / 8) .. floor((keep_from - 8) / 8)
            const int ca = (since + G - 1) >> gs, cb = (keep_from - G) >> gs;
            if (keep_from >= G && cb >= ca) need &= ~bits_range(ca, min(cb, 63));
          }
        }
        if (next_t0 >= tb + S) break;
        ++ci;
        if (ci >= se) break;
        sg = nx;
        if (ci + 1 < se) nx = __ldg(&p.segs[ci + 1]);
      }
    }
    // per-group work lists: one reservation per warp and group, entries in lane (= dynamic index) order.
    // Lane l reserves for groups l and l + 32, so a warp waits for ONE round of atomics, not for one per
    // group (62 dependent round trips were half of this kernel's time).
    uint32_t cnt0 = 0, cnt1 = 0;
    for (int g = 0; g < groups; ++g) {
      const unsigned m = __ballot_sync(0xffffffffu, (need >> g) & 1ULL);
      if (lane == (g & 31)) { if (g < 32) cnt0 = __popc(m); else cnt1 = __popc(m); }
    }
    const uint32_t base0 = cnt0 ? atomicAdd(&p.work_count[lane], cnt0) : 0u;
    const uint32_t base1 = cnt1 ? atomicAdd(&p.work_count[lane + 32], cnt1) : 0u;
    for (int g = 0; g < groups; ++g) {
      const unsigned m = __ballot_sync(0xffffffffu, (need >> g) & 1ULL);
      const uint32_t base = __shfl_sync(0xffffffffu, g < 32 ? base0 : base1, g & 31);
      if ((need >> g) & 1ULL) p.work[(size_t)g * p.N + base + __popc(m & ((1u << lane) - 1u))] = j;
    }
  }
}

__global__ void __launch_bounds__(256) k_scatter_static(Params p) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a < p.N && p.cls[a] == kClsStatic) {
    const int4 r = p.srec[a];
    const uint32_t dest = p.scells[cell_of_xy(p, p.floors[fl_floor(r.z)], r.x)] + (uint32_t)r.w;
    p.ssorted[dest] = make_int4(r.x, a, r.z, r.y);
  }
  // the dynamic agents' stays
  const uint32_t ne = *p.ient_count;
  for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < ne; e += gridDim.x * blockDim.x) {
    const int4 r = p.ient[e];
    const uint32_t dest = p.scells[cell_of_xy(p, p.floors[fl_floor(r.z)], r.x)] + p.ient_rank[e];
    p.ssorted[dest] = r;
  }
}

// expansion of the dynamic agents: work item = an entry of a row group's work list (blockIdx.y = the
// group of 8 rows), so every lane has rows to expand and neighbouring lanes hold neighbouring agents
__global__ void __launch_bounds__(128) k_expand_dyn(Params p, int t_first, int rows) {
  const int g = blockIdx.y;
  const uint32_t n = p.work_count[g];
  const uint32_t* list = p.work + (size_t)g * p.N;
  for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < n; w += gridDim.x * blockDim.x) {
    const uint32_t j = list[w];
    expand_rows(p, (int)p.dyn_agent[j], (size_t)j, t_first, rows, g * 8, 1, 8, true);
  }
}

__global__ void __launch_bounds__(256) k_scatter_dyn(Params p, int rows /* S */) {
  const int row = blockIdx.y;
  if (row >= rows) return;
  const int g = (row + 1) >> 3;  // the row group of record row `row + 1`
  const uint32_t n = p.work_count[g];
  const uint32_t* list = p.work + (size_t)g * p.N;
  const int4* rec = p.rec + (size_t)(row + 1) * p.N;
  const uint32_t* cs = p.cells + (size_t)row * p.dcells_stride;
  const uint8_t* mask = p.dmask + (size_t)row * p.N;
  int4* out = p.sorted + (size_t)row * p.N;
  for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < n; w += gridDim.x * blockDim.x) {
    const uint32_t j = list[w];
    if (!mask[j]) continue;  // not around, in no space, or inside a stay that is in the static array
    const int4 r = rec[j];
    const uint32_t dest = cs[dcell_of_xy(p, p.floors[fl_floor(r.z)], r.x)] + (uint32_t)r.w;
    out[dest] = make_int4(r.x, (int)p.dyn_agent[j], r.z, r.y);
  }
}

// per step the dynamic agents whose state may have changed, in (coarse) cell order; on `all_row`
// (the first step of a range) every located dynamic agent
__global__ void __launch_bounds__(256) k_select_changed_dyn(Params p, int rows /* S */, int all_row) {
  const int row = blockIdx.y;
  if (row >= rows) return;
  const int total = (int)p.cells[(size_t)row * p.dcells_stride + p.dtotal_cells];
  const int lane = threadIdx.x & 31;
  if (blockIdx.x == 0 && threadIdx.x == 0 && total) atomicAdd(&p.ctr->binned, (unsigned long long)total);
  for (int q0 = blockIdx.x * blockDim.x; q0 < total; q0 += gridDim.x * blockDim.x) {
    const int q = q0 + threadIdx.x;
    bool sel = false;
    if (q < total) sel = row == all_row || fl_changed(p.sorted[(size_t)row * p.N + q].z);
    unsigned m = __ballot_sync(0xffffffffu, sel);
    if (m == 0) continue;
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(&p.chg_count[row * 32], (uint32_t)__popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (sel) p.chg_list[(size_t)row * p.N + base + __popc(m & ((1u << lane) - 1))] = (uint32_t)q;
  }
}

// K1b: the Calculator.run entry: caller-supplied states -> record row (+ histogram).
__global__ void k_states_to_rec(Params p, const citam_agent_state* st, int4* rec, int step, int do_bin) {
  int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= p.N) return;
  citam_agent_state s = st[a];
  int f = s.floor < 0 ? -1 : s.floor;
  int l = (f < 0 || s.loc < 0) ? -1 : s.loc;
  if (f < 0) { rec[a] = make_int4(0, 0, pack_fl(-1, -1), 0); return; }
  if (l >= 0) {
    bool ok = f < p.F;
    if (ok) {
      const FloorDev& fd = p.floors[f];
      int fx = s.x - fd.minx, fy = s.y - fd.miny;
      ok = fd.lut != nullptr && (unsigned)fx < (unsigned)fd.W && (unsigned)fy < (unsigned)fd.H && l < fd.n_spaces;
    }
    if (!ok) {  // caller says "located" but the point is outside the floor's grid
      atomicOr(&p.ctr->overflow, OVF_GRID);
      l = -1;
    }
  }
  if (l < 0) { rec[a] = make_int4(s.x, s.y, pack_fl(f & 0x3f, -1), 0); return; }
  const FloorDev& fd = p.floors[f];
  const int32_t xy = (s.x - fd.minx) | ((s.y - fd.miny) << 16);
  uint32_t rk = 0;
  if (do_bin) rk = atomicAdd(&p.cells[cell_of_xy(p, fd, xy)], 1u);
  rec[a] = make_int4(xy, step + 1, pack_fl(f, l), (int)rk);
}

__global__ void k_rec_to_states(Params p, const int4* rec, long long N, citam_agent_state* st) {
  long long a = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= N) return;
  int4 r = rec[a];
  citam_agent_state s;
  s.floor = (int16_t)fl_floor(r.z); s.loc = (int16_t)fl_loc(r.z);
  if (fl_located(r.z)) {
    const FloorDev& fd = p.floors[s.floor];
    s.x = xy_x(r.x) + fd.minx; s.y = xy_y(r.x) + fd.miny;
  } else {
    s.x = r.x; s.y = r.y;
  }
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
  __shared__ uint32_t s_off;
  if (threadIdx.x < 32) {  // at most kMaxScanParts = 32 pieces: one load per lane, shuffle sum
    uint32_t v = (int)threadIdx.x < part ? part_sums[blockIdx.y * (gridDim.x + 1) + threadIdx.x] : 0u;
    v = __reduce_add_sync(0xffffffffu, v);
    if (threadIdx.x == 0) s_off = v;
  }
  __syncthreads();
  const uint32_t off = s_off;
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
  if (a == 0) atomicAdd(&p.ctr->binned, (unsigned long long)p.cells[(size_t)row * p.cells_stride + p.total_cells]);
  if (p.window != nullptr) {  // 8 bytes instead of the 16-byte record for agents that are not around
    const int2 win = p.window[a];
    const int t = t_begin + row;
    if (t < win.x || t >= win.y) return;
  }
  int4 r = p.rec[(size_t)(row + 1) * p.N + a];
  if (!fl_located(r.z)) return;
  int key = cell_of_xy(p, p.floors[fl_floor(r.z)], r.x);
  uint32_t dest = p.cells[(size_t)row * p.cells_stride + key] + (uint32_t)r.w;
  p.sorted[(size_t)row * p.N + dest] = make_int4(r.x, a, r.z, r.y);
}

// K3b: per step, the sorted positions of the agents whose state may have changed, compacted in
// CELL ORDER (each warp appends its 32 consecutive sorted positions' hits in one piece), so that
// neighbouring threads of k_pairs_changed work on neighbouring agents: their candidate ranges
// overlap (L1 hits) and their histogram buckets share sectors.
__global__ void __launch_bounds__(256) k_select_changed(Params p, int row0, int rows /* S */) {
  const int row = blockIdx.y + row0;  // a range's first step is handled by k_pairs in full
  if (row >= rows) return;
  const int total = (int)p.cells[(size_t)row * p.cells_stride + p.total_cells];
  const int lane = threadIdx.x & 31;
  for (int q0 = blockIdx.x * blockDim.x; q0 < total; q0 += gridDim.x * blockDim.x) {
    const int q = q0 + threadIdx.x;
    bool sel = false;
    if (q < total) sel = fl_changed(p.sorted[(size_t)row * p.N + q].z);
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
// |dx|, |dy| < 2^16 (lattice-relative), so each square fits 32 bits.
__device__ __forceinline__ bool within(int dx, int dy, const Params& p) {
  const uint32_t ax = (uint32_t)abs(dx), ay = (uint32_t)abs(dy);
  if (max(ax, ay) > p.d_lim) return false;
  return (unsigned long long)(ax * ax) + (unsigned long long)(ay * ay) < p.s_lim;
}
__device__ __forceinline__ bool within_xy(int32_t xy1, int32_t xy2, const Params& p) {
  return within(xy_x(xy2) - xy_x(xy1), xy_y(xy2) - xy_y(xy1), p);
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

// The same when the caller has already fetched the entry at the key's home slot `h`.
__device__ __forceinline__ unsigned long long pair_slot_from(const Params& p, unsigned long long key, unsigned long long h,
                                                             ulonglong2 e, unsigned long long* acc_seen) {
  for (int probe = 0; probe < kMaxProbes; ++probe) {
    if (probe) e = load_entry(&p.pairs[h]);
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

// Apply add_dur contact-steps and add_nev events to the slot; `first` (a step, or NO_STEP) lowers
// the pair's first-contact step.  The first contact of a pair is always an event start, so callers
// pass a step only for event starts and for the first step of a run range.
// `old` is the accumulator word the probe saw.  While it is far from the field limits (less than
// half way: no set of concurrent updates of one pair can cover the other half, a pair gets at most
// one record per step) and the first-contact field cannot move, the update is a plain addition on
// the packed word -- one reduction, no round trip, no retry.  Otherwise a compare-and-swap loop
// that checks the limits on the live value.
__device__ __forceinline__ void pair_apply_acc(const Params& p, unsigned long long h, unsigned long long old, uint32_t first,
                                               unsigned long long add_dur, unsigned long long add_nev) {
  unsigned long long finv = first == NO_STEP ? 0ULL : (unsigned long long)(0xFFFFFFu - (first & 0xFFFFFFu));
  if (finv <= ACC_FIRST_INV(old) && ACC_DUR(old) + add_dur <= 0x7fffffULL && ACC_NEV(old) + add_nev <= 0x7fffULL) {
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

// The same plus both agents' counters (close_contacts_calculator.py:221-224): the counter array is
// N x 8 bytes, L2-resident, so these two reductions cost no DRAM access.  One word per agent carries
// its contact-steps (low 40 bits) and its number of contact events (high 24 bits), so the run's
// statistics (contacts.py:185-294: events per agent, agents with contacts) need no pass over the pair
// table.  citam_run keeps (n_agents - 1) x steps since the reset below 2^40 (the most contact-steps an
// agent can collect); the event field is exact whenever the agent's contact-steps are < 2^24 (an event
// has at least one step), which the statistics kernel checks before trusting it.
__device__ __forceinline__ void pair_apply(const Params& p, unsigned long long h, unsigned long long old, uint32_t first,
                                           unsigned long long add_dur, unsigned long long add_nev, uint32_t a1, uint32_t a2) {
  const unsigned long long w = add_dur | (add_nev << kAgentDurBits);
  atomicAdd(&p.agent_dur[a1], w);
  atomicAdd(&p.agent_dur[a2], w);
  pair_apply_acc(p, h, old, first, add_dur, add_nev);
}

__device__ __forceinline__ void pair_upsert(const Params& p, uint32_t a1, uint32_t a2, uint32_t first,
                                            unsigned long long add_dur, unsigned long long add_nev) {
  unsigned long long key = ((unsigned long long)a1 << 32) | a2;
  unsigned long long old;
  unsigned long long h = pair_slot(p, key, &old);
  if (h == ~0ULL) return;
  pair_apply(p, h, old, first, add_dur, add_nev, a1, a2);
}

// ix, iy: lattice-relative contact position x 2 (the sum of the two agents' relative coordinates)
__device__ __forceinline__ void coord_upsert_rel(const Params& p, int floor, int ix, int iy, unsigned long long add) {
  const FloorDev& fd = p.floors[floor];
  if (p.hist != nullptr) {  // dense: one reduction, no probe
    atomicAdd(p.hist + fd.hist_base + (long long)iy * fd.hist_w + ix, add);
    return;
  }
  const int sx = ix + 2 * fd.minx, sy = iy + 2 * fd.miny;
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
  if (!fl_located(r1.z) || !fl_located(r2.z)) return false;
  const int fl1 = fl_floor(r1.z);
  if (fl1 != fl_floor(r2.z)) return false;
  if (!within_xy(r1.x, r2.x, p)) return false;
  return loc_rule(p.floors[fl1], fl_loc(r1.z), fl_loc(r2.z));
}
// An agent's record at the step before row `t_rel + 1` of the chunk (row t_rel; row 0 = halo step)
__device__ __forceinline__ int4 prev_rec(const Params& p, uint32_t t_rel, uint32_t a) {
  if (p.cls == nullptr) return p.rec[(size_t)t_rel * p.N + a];
  const uint32_t c = p.cls[a];
  if (c == kClsStatic) return p.srec[a];
  if (c == kClsOut) return make_int4(0, 0, (int32_t)0x8000ffffu, 0);
  return p.rec[(size_t)t_rel * p.N + c];
}
__device__ __forceinline__ bool contact_prev(const Params& p, uint32_t t_rel, uint32_t a1, uint32_t a2) {
  return contact_eval(p, prev_rec(p, t_rel, a1), prev_rec(p, t_rel, a2));
}

// A contact found at step t that lasts `run` steps (both agents provably stay put for run-1
// more steps of the range): one table update for the whole run.  `was`: the pair was already in
// contact at t-1 (the event continues, contacts.py:131-133); `late`: not decided yet -- evaluate the
// predicate on the previous record row when the record is applied.
// me / ot are sorted records {xy, agent, state word, until}.
__device__ __forceinline__ void on_contact(const Params& p, uint32_t t, const int4& me, const int4& ot, uint32_t run,
                                           bool was, int sub, bool late = false) {
  uint32_t a1 = (uint32_t)min(me.y, ot.y), a2 = (uint32_t)max(me.y, ot.y);
  const int floor = fl_floor(me.z);
  const int ix = xy_x(me.x) + xy_x(ot.x), iy = xy_y(me.x) + xy_y(ot.x);
  if (p.cbuf != nullptr) {
    // warp-aggregated append into the chunk's contact buffer; k_accumulate applies it
    const FloorDev& fd = p.floors[floor];
    uint32_t bucket = ((uint32_t)fd.hist_base + (uint32_t)iy * (uint32_t)fd.hist_w + (uint32_t)ix) | (was ? 0x80000000u : 0u);
    cg::coalesced_group g = cg::coalesced_threads();
    unsigned long long base = 0;
    if (g.thread_rank() == 0)
      base = atomicAdd(&p.cbuf_count[sub * kCounterStride], (unsigned long long)g.size());
    base = g.shfl(base, 0) + g.thread_rank();
    if (base < p.cbuf_sub_cap) {
      ContactRec r;
      r.a1 = a1; r.a2 = a2; r.t_run = (t - (uint32_t)p.chunk_begin) | ((run - 1u) << 12);
      r.bucket = late ? (bucket & 0x3fffffffu) | 0x40000000u : bucket;  // bit 30: evaluate `was` in k_accumulate
      p.cbuf[(unsigned long long)sub * p.cbuf_sub_cap + base] = r;
      return;
    }
    // sub-buffer full: fall through and accumulate inline (k_accumulate clamps the count)
  }
  if (late) was = contact_prev(p, t - (uint32_t)p.chunk_begin, a1, a2);
  uint32_t first = (!was || (int32_t)t == p.range_begin) ? t : NO_STEP;
  pair_upsert(p, a1, a2, first, run, was ? 0u : 1u);
  coord_upsert_rel(p, floor, ix, iy, (unsigned long long)run);
  if (p.log_cap) {
    // warp-aggregated append: one atomic per converged group of lanes
    cg::coalesced_group g = cg::coalesced_threads();
    unsigned long long base = 0;
    if (g.thread_rank() == 0) base = atomicAdd(&p.ctr->log_count, (unsigned long long)g.size());
    base = g.shfl(base, 0) + g.thread_rank();
    if (base < p.log_cap) {
      p.log[base] = make_uint4(t, a1, a2, 0u);
    } else {
      atomicOr(&p.ctr->overflow, OVF_LOG);
    }
  }
}

// steps a contact found at t lasts: until either agent's state may change, or the range ends
__device__ __forceinline__ uint32_t run_length(const Params& p, int until_a, int until_b, uint32_t t) {
  return (uint32_t)(min(min(until_a, until_b), p.range_end) - (int)t);
}

// One thread per located agent of one step ("me", in cell order) walks its half neighbourhood:
// the rest of its cell, the cell to the right and the three cells of the row above.  In cell
// order all of a block's candidates form ONE contiguous window of the sorted array, which the
// block stages in shared memory with coalesced 16-byte loads, so the per-candidate loop runs at
// shared-memory latency (windows beyond kWindow records are read from global memory instead).
//
// Run for the FIRST step of a citam_run range, which owns every contact it finds and adds the
// whole run (until either agent's state may change, clipped to the range) in one update; later
// steps only look at agents whose state may have changed (k_pairs_changed).  With the contact log
// on, every step is evaluated here in full and logged contact-step by contact-step (aggregate == 0).
constexpr int kPairThreads = 256;
constexpr int kWindow = 2560;  // records (40 KB)

__global__ void __launch_bounds__(kPairThreads) k_pairs(Params p, int t_begin, int rows /* S */, int aggregate, int statics) {
  __shared__ int4 win[kWindow];
  __shared__ int s_w1;
  __shared__ int s_void;
  const int row = blockIdx.y;
  // statics: the chunk's static agents (split pipeline; one row, the range's first step)
  const uint32_t* cs = statics ? p.scells : p.cells + (size_t)row * p.cells_stride;
  const int total = (int)cs[p.total_cells];
  const int w0 = blockIdx.x * kPairThreads;
  if (w0 >= total) return;  // whole block
  const int q0 = w0 + threadIdx.x;
  const int4* srt = statics ? p.ssorted : p.sorted + (size_t)row * p.N;
  if (threadIdx.x == 0) {
    s_w1 = w0;
    // results are void already when a table has overflowed: one read per block, the same answer for all its threads
    s_void = (*(volatile unsigned long long*)&p.ctr->overflow & (OVF_PAIR | OVF_COORD | OVF_FIELD)) != 0;
  }
  __syncthreads();
  if (s_void) return;
  int4 me = make_int4(0, 0, 0, 0);
  int qa = 0, qa_end = 0, qb = 0, qb_end = 0, f = 0, loc = -1;
  bool me_ok = q0 < total;
  if (q0 < total) {
    me = srt[q0];
    // the static array also holds stays that begin later in the chunk: only those holding at this step
    if (statics) me_ok = entry_valid(me, row, t_begin + row);
    f = fl_floor(me.z); loc = fl_loc(me.z);
    const FloorDev& fd = p.floors[f];
    int cx = div_cell(p, xy_x(me.x)), cy = div_cell(p, xy_y(me.x));
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
  if (me_ok) {
    const FloorDev& fd = p.floors[f];
    const uint32_t t = (uint32_t)(t_begin + row);
    const int sub = (blockIdx.x + blockIdx.y) & (kSubBuffers - 1);
#pragma unroll 1
    for (int rg = 0; rg < 2; ++rg) {
      const int qs = rg ? qb : qa, qe = rg ? qb_end : qa_end;
      for (int q = qs; q < qe; ++q) {
        int4 ot = staged ? win[q - w0] : srt[q];
        if (statics && !entry_valid(ot, row, (int)t)) continue;
        ++tests;
        if (within_xy(me.x, ot.x, p) && loc_rule_ordered(fd, me.y, loc, ot.y, fl_loc(ot.z))) {
          uint32_t run = aggregate ? run_length(p, me.w, ot.w, t) : 1u;
          hits += run;
          ++runs;
          // both agents exactly where (and what) they were a step ago: in contact then too;
          // otherwise the predicate on the previous record row is evaluated at accumulation
          const bool known = !fl_changed(me.z) && !fl_changed(ot.z);
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

// Every step but a range's first: only agents whose state may have changed can start, end or
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

__global__ void __launch_bounds__(kPC) k_pairs_changed(Params p, int t_begin, int row0, int rows /* S */) {
  __shared__ ContactRec hitbuf[kPC * kHitSlots];
  __shared__ int4 s_me[kPC];
  __shared__ int s_qs[3][kPC];
  __shared__ int s_l0[kPC], s_l01[kPC];
  __shared__ int s_off[kPC + 1];
  __shared__ int s_warp[kPC / 32];
  __shared__ int s_void;
  const int row = blockIdx.y + row0;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const uint32_t n = p.chg_count[row * 32];
  if (blockIdx.x * kPC >= n) return;
  if (threadIdx.x == 0)
    s_void = (*(volatile unsigned long long*)&p.ctr->overflow & (OVF_PAIR | OVF_COORD | OVF_FIELD)) != 0;
  __syncthreads();
  if (s_void) return;
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
      const FloorDev& fd = p.floors[fl_floor(me.z)];
      const int cx = div_cell(p, xy_x(me.x)), cy = div_cell(p, xy_y(me.x));
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
      const int a = me.y;
      if (ot.y == a) continue;
      if (fl_changed(ot.z) && ot.y < a) continue;  // that agent owns the pair
      ++tests;
      if (!within_xy(me.x, ot.x, p)) continue;
      const FloorDev& fd = p.floors[fl_floor(me.z)];
      if (!loc_rule_ordered(fd, a, fl_loc(me.z), ot.y, fl_loc(ot.z))) continue;
      const uint32_t run = run_length(p, me.w, ot.w, t);
      hits += run;
      ++runs;
      if (buffered && nh < kHitSlots) {
        ContactRec r;
        r.a1 = (uint32_t)min(a, ot.y); r.a2 = (uint32_t)max(a, ot.y);
        r.t_run = (t - (uint32_t)p.chunk_begin) | ((run - 1u) << 12);
        r.bucket = (((uint32_t)fd.hist_base + (uint32_t)(xy_y(me.x) + xy_y(ot.x)) * (uint32_t)fd.hist_w +
                     (uint32_t)(xy_x(me.x) + xy_x(ot.x))) & 0x3fffffffu) | 0x40000000u;
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
            const uint32_t rt = (uint32_t)p.chunk_begin + (r.t_run & 0xfffu), rrun = (r.t_run >> 12) + 1u;
            bool was = contact_prev(p, r.t_run & 0xfffu, r.a1, r.a2);
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

// The same for the split pipeline: the changed agents of a step are all dynamic; each is tested
// against the chunk's STATIC agents (3 x 3 fine cells of `scells` / `ssorted`: three ranges) and
// against the step's DYNAMIC agents (3 x 3 coarse cells of the step's row of `cells` / `sorted`:
// three more ranges) -- six contiguous ranges per agent.  The batch's candidates are flattened into
// one index space; a shared-memory tile maps a flat index to its (agent, range) -- filled by the
// agents' threads, read back with one load per candidate (the binary search this replaces was a
// quarter of the kernel's instructions) -- and `s_delta` turns the flat index into the candidate's
// position in the sorted array.  A static agent never owns a pair with a changed one; two dynamic
// agents: the lower id owns the pair if the other one is changed too (on `all_row`, the first step of
// a range, every dynamic agent counts as changed: that step owns every contact it finds,
// static-static pairs are k_pairs').
constexpr int kTile = 4096;  // candidates per tile of the flat index space

// Candidates that passed the distance test wait in a per-warp queue and are finished 32 at a time by
// all lanes of the warp (space rule, run length, record, ONE reservation in the contact buffer per 32
// records): one candidate in six passes, so finishing each where it was found left 27 of 32 lanes idle
// through a quarter of the kernel's instructions.
struct HitQueue {
  int4 ot[64];       // the candidate's sorted record
  uint16_t who[64];  // the changed agent's index in the batch
};

__device__ __forceinline__ void drain_hits(const Params& p, HitQueue& q, const int4* s_me, const int4* s_prev, int count,
                                           uint32_t t, int sub, unsigned long long& hits, unsigned long long& runs) {
  const int lane = threadIdx.x & 31;
  bool good = false;
  ContactRec rec;
  int4 me = make_int4(0, 0, 0, 0), ot = me;
  uint32_t run = 0;
  if (lane < count) {
    ot = q.ot[lane];
    me = s_me[q.who[lane]];
    const FloorDev& fd = p.floors[fl_floor(me.z)];  // cells are per floor: the candidate is on the same one
    good = loc_rule_ordered(fd, me.y, fl_loc(me.z), ot.y, fl_loc(ot.z));
    if (good) {
      run = run_length(p, me.w, ot.w, t);
      hits += run;
      ++runs;
      rec.a1 = (uint32_t)min(me.y, ot.y); rec.a2 = (uint32_t)max(me.y, ot.y);
      rec.t_run = (t - (uint32_t)p.chunk_begin) | ((run - 1u) << 12);
      rec.bucket = ((uint32_t)fd.hist_base + (uint32_t)(xy_y(me.x) + xy_y(ot.x)) * (uint32_t)fd.hist_w +
                    (uint32_t)(xy_x(me.x) + xy_x(ot.x))) & 0x3fffffffu;
      // Was the pair in contact one step earlier (does the event continue)?  A candidate that is
      // exactly where and what it was a step ago (static, or standing: `changed` clear) IS its own
      // previous record, and the changed agent's previous record was fetched once in phase 1: decided
      // here, k_accumulate reads no records for it.  Two changed agents: left to k_accumulate (bit 30).
      if (!fl_changed(ot.z)) {
        const int4 mp = s_prev[q.who[lane]];
        const bool was = me.y < ot.y ? contact_eval(p, mp, ot) : contact_eval(p, ot, mp);  // (lower id, higher id), as k_accumulate does
        rec.bucket |= was ? 0x80000000u : 0u;
      }
      else rec.bucket |= 0x40000000u;
    }
  }
  const unsigned m = __ballot_sync(0xffffffffu, good);
  if (m == 0) return;
  bool stored = false;
  if (p.cbuf != nullptr) {
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(&p.cbuf_count[sub * kCounterStride], (unsigned long long)__popc(m));
    base = __shfl_sync(0xffffffffu, base, 0) + __popc(m & ((1u << lane) - 1u));
    if (good && base < p.cbuf_sub_cap) {
      p.cbuf[(unsigned long long)sub * p.cbuf_sub_cap + base] = rec;
      stored = true;
    }
  }
  if (good && !stored) {  // no contact buffer, or it is full (k_accumulate clamps the count): apply here
    const bool was = (rec.bucket & 0x40000000u) ? contact_prev(p, rec.t_run & 0xfffu, rec.a1, rec.a2) : (rec.bucket >> 31) != 0;
    const uint32_t first = (!was || (int32_t)t == p.range_begin) ? t : NO_STEP;
    pair_upsert(p, rec.a1, rec.a2, first, run, was ? 0u : 1u);
    coord_upsert_rel(p, fl_floor(me.z), xy_x(me.x) + xy_x(ot.x), xy_y(me.x) + xy_y(ot.x), (unsigned long long)run);
  }
}

template <int kPS>
__global__ void __launch_bounds__(kPS, 1024 / kPS) k_pairs_changed_split(Params p, int t_begin, int rows /* S */, int all_row) {
  __shared__ int4 s_me[kPS];
  __shared__ int4 s_prev[kPS];  // the batch's agents one step earlier
  __shared__ HitQueue s_queue[kPS / 32];
  __shared__ int s_delta[kPS * 6];        // candidate position = flat index + s_delta[agent * 6 + range]
  __shared__ uint16_t s_owner[kTile];     // agent * 6 + range of a flat index (relative to the tile)
  __shared__ int s_warp[kPS / 32];
  __shared__ int s_total;
  __shared__ int s_void;
  const int row = blockIdx.y;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const uint32_t n = p.chg_count[row * 32];
  if (blockIdx.x * kPS >= n) return;
  if (threadIdx.x == 0)
    s_void = (*(volatile unsigned long long*)&p.ctr->overflow & (OVF_PAIR | OVF_COORD | OVF_FIELD)) != 0;
  __syncthreads();
  if (s_void) return;
  const uint32_t* cs = p.scells;                                      // static starts (fine grid)
  const uint32_t* cd = p.cells + (size_t)row * p.dcells_stride;       // this step's dynamic starts (coarse grid)
  const int4* srt = p.sorted + (size_t)row * p.N;
  const uint32_t t = (uint32_t)(t_begin + row);
  const int sub = (blockIdx.x + blockIdx.y) & (kSubBuffers - 1);
  const bool all = row == all_row;
  HitQueue& queue = s_queue[wid];
  int qn = 0;  // queued hits of this warp (the same in all its lanes)
  unsigned long long tests = 0, hits = 0, runs = 0;
  for (uint32_t base = blockIdx.x * kPS; base < n; base += gridDim.x * kPS) {
    // ---- phase 1: one thread per agent describes its six candidate ranges
    const uint32_t idx = base + threadIdx.x;
    int total = 0;
    int qs[6], len[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) { qs[r] = 0; len[r] = 0; }
    if (idx < n) {
      const int4 me = srt[p.chg_list[(size_t)row * p.N + idx]];
      const FloorDev& fd = p.floors[fl_floor(me.z)];
      {
        const int cx = div_cell(p, xy_x(me.x)), cy = div_cell(p, xy_y(me.x));
        const int x0 = max(cx - 1, 0), x1 = min(cx + 1, fd.ncx - 1);
        const int rb = fd.cell_base + cy * fd.ncx;
        const bool up = cy + 1 < fd.ncy, dn = cy > 0;
        const int rbu = up ? rb + fd.ncx : rb, rbd = dn ? rb - fd.ncx : rb;
        qs[0] = (int)cs[rbd + x0]; qs[1] = (int)cs[rb + x0]; qs[2] = (int)cs[rbu + x0];
        len[0] = dn ? (int)cs[rbd + x1 + 1] - qs[0] : 0;
        len[1] = (int)cs[rb + x1 + 1] - qs[1];
        len[2] = up ? (int)cs[rbu + x1 + 1] - qs[2] : 0;
      }
      {
        const int cx = div_dcell(p, xy_x(me.x)), cy = div_dcell(p, xy_y(me.x));
        const int x0 = max(cx - 1, 0), x1 = min(cx + 1, fd.dncx - 1);
        const int rb = fd.dcell_base + cy * fd.dncx;
        const bool up = cy + 1 < fd.dncy, dn = cy > 0;
        const int rbu = up ? rb + fd.dncx : rb, rbd = dn ? rb - fd.dncx : rb;
        qs[3] = (int)cd[rbd + x0]; qs[4] = (int)cd[rb + x0]; qs[5] = (int)cd[rbu + x0];
        len[3] = dn ? (int)cd[rbd + x1 + 1] - qs[3] : 0;
        len[4] = (int)cd[rb + x1 + 1] - qs[4];
        len[5] = up ? (int)cd[rbu + x1 + 1] - qs[5] : 0;
      }
      s_me[threadIdx.x] = me;
      s_prev[threadIdx.x] = p.rec[(size_t)row * p.N + p.cls[me.y]];  // row `row` holds step t-1 (row 0: the halo step)
#pragma unroll
      for (int r = 0; r < 6; ++r) total += len[r];
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
    int start[6];  // first flat index of each range
    {
      int o = woff + inc - total;
#pragma unroll
      for (int r = 0; r < 6; ++r) {
        start[r] = o;
        s_delta[threadIdx.x * 6 + r] = qs[r] - o;
        o += len[r];
      }
    }
    if (threadIdx.x == kPS - 1) s_total = woff + inc;
    __syncthreads();
    const int P = s_total;
    for (int tile = 0; tile < P; tile += kTile) {
      // ---- the tile's flat index -> (agent, range) map
#pragma unroll
      for (int r = 0; r < 6; ++r) {
        const int lo = max(start[r], tile) - tile, hi = min(start[r] + len[r], tile + kTile) - tile;
        const uint16_t o = (uint16_t)(threadIdx.x * 6 + r);
        for (int x = lo; x < hi; ++x) s_owner[x] = o;
      }
      __syncthreads();
      // ---- phase 2: one candidate test per thread and iteration (the same trip count for the whole block)
      const int tn = min(kTile, P - tile);
      for (int pp0 = 0; pp0 < tn; pp0 += kPS) {
        const int pp = pp0 + threadIdx.x;
        bool hit = false;
        int4 ot = make_int4(0, 0, 0, 0);
        int i = 0;
        if (pp < tn) {
          const int o = s_owner[pp];
          i = (o * 10923) >> 16;  // o / 6 for o < 1536
          const bool dyn_cand = o - i * 6 >= 3;
          ot = __ldg((dyn_cand ? srt : p.ssorted) + (tile + pp + s_delta[o]));
          const int4 me = s_me[i];
          // a dynamic candidate: not the agent itself, and not a pair the other agent owns;
          // a static-array candidate: a stay that holds at this step
          const bool skip = dyn_cand ? (ot.y == me.y || ((all || fl_changed(ot.z)) && ot.y < me.y))
                                     : !entry_valid(ot, row, (int)t);
          if (!skip) {
            ++tests;
            hit = within_xy(me.x, ot.x, p);
          }
        }
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (m) {
          if (hit) {
            const int pos = qn + __popc(m & ((1u << lane) - 1u));
            queue.ot[pos] = ot;
            queue.who[pos] = (uint16_t)i;
          }
          qn += __popc(m);
          __syncwarp();
          if (qn >= 32) {
            drain_hits(p, queue, s_me, s_prev, 32, t, sub, hits, runs);
            __syncwarp();
            int4 mv = make_int4(0, 0, 0, 0);
            uint16_t mw = 0;
            if (32 + lane < qn) { mv = queue.ot[32 + lane]; mw = queue.who[32 + lane]; }
            __syncwarp();
            if (32 + lane < qn) { queue.ot[lane] = mv; queue.who[lane] = mw; }
            qn -= 32;
            __syncwarp();
          }
        }
      }
      __syncthreads();  // the tile map is rewritten
    }
    // the batch's agents (s_me) are replaced next: finish what is queued
    if (qn) {
      drain_hits(p, queue, s_me, s_prev, qn, t, sub, hits, runs);
      qn = 0;
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
// + one 64-bit reduction on the same sector (pair table) and one 64-bit reduction
// (coordinate histogram); for run starts that are not known continuations also two 16-byte
// record reads for the previous-step predicate, issued before the probe so they overlap it.
// many short-lived blocks rather than a persistent grid: the kernel shares the GPU with the next
// chunk's build on the other stream, and short blocks let that stream's blocks in (measured on config 3:
// 24 blocks per sub-buffer 2.08e10 agent-steps/s, 96: 2.27e10, 256: 2.37e10)
constexpr int kAccBlocksPerSub = 256;

// kAccIlp records per thread and iteration: the loads of all of them (records, previous-step records,
// first slot probes) are issued before any is consumed -- the kernel is bound by memory latency at full
// occupancy, so the independent chains in flight per thread are what sets its rate.
template <int K>
__global__ void __launch_bounds__(256) k_accumulate(Params p, int t_begin) {
  const int sub = blockIdx.x % kSubBuffers, part = blockIdx.x / kSubBuffers;
  const unsigned long long n = min(p.cbuf_count[sub * kCounterStride], p.cbuf_sub_cap);
  const ContactRec* buf = p.cbuf + (unsigned long long)sub * p.cbuf_sub_cap;
  const unsigned long long stride = (unsigned long long)(gridDim.x / kSubBuffers) * blockDim.x;
  for (unsigned long long i0 = (unsigned long long)part * blockDim.x + threadIdx.x; i0 < n; i0 += stride * K) {
    ContactRec r[K];
    bool live[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      live[k] = i0 + k * stride < n;
      if (live[k]) r[k] = buf[i0 + k * stride];
    }
    int4 r1[K], r2[K];
    unsigned long long h[K];
    ulonglong2 e0[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      if (!live[k]) continue;
      // the histogram reduction, the previous-step records and the table probe are independent: issue all
      atomicAdd(p.hist + (r[k].bucket & 0x3fffffffu), (unsigned long long)((r[k].t_run >> 12) + 1u));
      if ((r[k].bucket >> 30) & 1) {  // both agents' records of step t-1
        r1[k] = prev_rec(p, r[k].t_run & 0xfffu, r[k].a1);
        r2[k] = prev_rec(p, r[k].t_run & 0xfffu, r[k].a2);
      }
      h[k] = mix64(((unsigned long long)r[k].a1 << 32) | r[k].a2) & p.pair_mask;
      e0[k] = load_entry(&p.pairs[h[k]]);
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
      if (!live[k]) continue;
      const uint32_t t_rel = r[k].t_run & 0xfffu, run = (r[k].t_run >> 12) + 1u;
      const uint32_t t = (uint32_t)t_begin + t_rel;
      bool was = r[k].bucket >> 31;
      unsigned long long acc;
      const unsigned long long slot = pair_slot_from(p, ((unsigned long long)r[k].a1 << 32) | r[k].a2, h[k], e0[k], &acc);
      if (slot == ~0ULL) continue;
      if ((r[k].bucket >> 30) & 1) was = contact_eval(p, r1[k], r2[k]);
      uint32_t first = (!was || (int32_t)t == p.range_begin) ? t : NO_STEP;
      pair_apply(p, slot, acc, first, run, was ? 0u : 1u, r[k].a1, r[k].a2);
    }
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
// K5: finalize — compaction of the tables, device sort into the reference's output order, and the
// integer sums of the statistics.
// --------------------------------------------------------------------------
__device__ __forceinline__ citam_pair entry_to_pair(unsigned long long key, unsigned long long acc) {
  citam_pair r;
  r.a1 = (uint32_t)(key >> 32);
  r.a2 = (uint32_t)(key & 0xffffffffu);
  uint32_t finv = ACC_FIRST_INV(acc);
  r.first_step = finv ? 0xFFFFFFu - finv : NO_STEP;
  r.n_events = ACC_NEV(acc);
  r.duration = ACC_DUR(acc);
  return r;
}

// the occupied slots of the pair table, packed (any order)
__global__ void k_entries_to_pairs(const PairEntry* in, unsigned long long n, citam_pair* out) {
  unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  PairEntry e = in[i];
  out[i] = entry_to_pair(e.key, e.acc);
}

// Sort keys (radix_sort.cuh; word 0 is the least significant).  A PairEntry as uint4 is
// {a2, a1, acc low, acc high}: pair_contact.csv order is (first contact step, a1, a2).
struct PairKey {
  __device__ __forceinline__ uint32_t operator()(const uint4& r, int word) const {
    return word == 0 ? r.x : (word == 1 ? r.y : (0xFFFFFFu - (r.w >> 8)) & 0xFFFFFFu);
  }
};
struct LogKey {  // {step, a1, a2, 0}: the reference's visiting order (step, a1, a2)
  __device__ __forceinline__ uint32_t operator()(const uint4& r, int word) const {
    return word == 0 ? r.z : (word == 1 ? r.y : r.x);
  }
};
struct CoordKey {  // CoordEntry {key, count}: the key is monotone in (floor, sx, sy)
  __device__ __forceinline__ uint32_t operator()(const uint4& r, int word) const { return word == 0 ? r.x : r.y; }
};

__global__ void k_log_to_contacts(const uint4* in, unsigned long long n, citam_contact* out) {
  unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint4 r = in[i];
  citam_contact c; c.step = r.x; c.a1 = r.y; c.a2 = r.z;
  out[i] = c;
}

__global__ void k_compact_coord_entries(const CoordEntry* tab, unsigned long long cap, CoordEntry* out,
                                        unsigned long long out_cap, DevCounters* ctr) {
  unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cap) return;
  CoordEntry e = tab[i];
  if (e.key == 0ULL) return;
  cg::coalesced_group g = cg::coalesced_threads();
  unsigned long long base = 0;
  if (g.thread_rank() == 0) base = atomicAdd(&ctr->export_cursor, (unsigned long long)g.size());
  base = g.shfl(base, 0) + g.thread_rank();
  if (base < out_cap) out[base] = e;
}

__global__ void k_coord_entries_to_records(const CoordEntry* in, unsigned long long n, citam_coord* out) {
  unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  CoordEntry e = in[i];
  citam_coord r;
  r.floor = (int32_t)((e.key >> 56) & 0x7f);
  r.sx = (int32_t)((e.key >> 28) & 0xfffffffu) - (1 << 27);
  r.sy = (int32_t)(e.key & 0xfffffffu) - (1 << 27);
  r.reserved = 0;
  r.count = e.count;
  out[i] = r;
}

// dense histogram: count the non-zero buckets
__global__ void k_hist_count(const unsigned long long* hist, long long n, DevCounters* ctr) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long nz = (i < n && hist[i] != 0ULL) ? 1ULL : 0ULL;
  for (int d = 16; d > 0; d >>= 1) nz += __shfl_down_sync(0xffffffffu, nz, d);
  if ((threadIdx.x & 31) == 0 && nz) atomicAdd(&ctr->export_cursor, nz);
}

// Dense histogram -> records in (floor, sx, sy) order without a sort: thread j of a floor looks at
// bucket (ix, iy) = (j / hist_h, j % hist_h), i.e. the floor is walked column by column; pass 1
// counts the non-zero buckets per CTA, an exclusive scan over all CTAs of all floors gives each CTA
// its first output index, pass 2 writes (ballot ranks inside the CTA).
constexpr int kHistThreads = 256;
__global__ void __launch_bounds__(kHistThreads) k_hist_ordered(const unsigned long long* hist, long long n, int hist_w,
                                                               int hist_h, int floor, int minx2, int miny2,
                                                               uint32_t* block_counts, citam_coord* out,
                                                               unsigned long long out_cap) {
  __shared__ uint32_t wsum[kHistThreads / 32];
  const long long j = (long long)blockIdx.x * kHistThreads + threadIdx.x;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  unsigned long long c = 0;
  int ix = 0, iy = 0;
  if (j < n) {
    ix = (int)(j / hist_h); iy = (int)(j % hist_h);
    c = hist[(long long)iy * hist_w + ix];
  }
  const unsigned m = __ballot_sync(0xffffffffu, c != 0ULL);
  if (lane == 0) wsum[w] = __popc(m);
  __syncthreads();
  if (out == nullptr) {  // pass 1
    if (threadIdx.x == 0) {
      uint32_t t = 0;
      for (int k = 0; k < kHistThreads / 32; ++k) t += wsum[k];
      block_counts[blockIdx.x] = t;
    }
    return;
  }
  if (c == 0ULL) return;
  unsigned long long pos = block_counts[blockIdx.x];  // scanned: first output index of this CTA
  for (int k = 0; k < w; ++k) pos += wsum[k];
  pos += __popc(m & ((1u << lane) - 1u));
  if (pos >= out_cap) return;
  citam_coord r;
  r.floor = floor; r.sx = ix + minx2; r.sy = iy + miny2; r.reserved = 0; r.count = c;
  out[pos] = r;
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

// ---- multi-GPU exchange: the table's records grouped by destination rank, hash(key) mod n_parts ----
__device__ __forceinline__ uint32_t part_of(unsigned long long key, uint32_t n_parts) {
  return (uint32_t)(mix64(key) >> 40) % n_parts;
}
constexpr int kMaxParts = 64;

// Both passes walk the table in tiles of 2048 slots per block; a warp owns 256 consecutive slots
// (8 coalesced rounds of 32).  Lanes that share a destination are ranked with __match_any_sync and a
// per-warp counter row in shared memory (no atomics); a block makes ONE global atomic per destination
// and tile (it was one per warp and round on two addresses: 33 ms for a 2-way split of 1.25e8 pairs).
constexpr int kPartRounds = 8;
constexpr uint32_t kBallotParts = 8;
constexpr int kPartTile = 256 * kPartRounds;

__global__ void __launch_bounds__(256) k_partition_count(const PairEntry* tab, unsigned long long cap, uint32_t n_parts,
                                                         unsigned long long* counts) {
  __shared__ uint32_t wcount[8][kMaxParts];
  for (int i = threadIdx.x; i < 8 * kMaxParts; i += 256) (&wcount[0][0])[i] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (unsigned long long tile = blockIdx.x; tile * kPartTile < cap; tile += gridDim.x) {
    const unsigned long long wbase = tile * kPartTile + (unsigned long long)w * (32 * kPartRounds);
#pragma unroll
    for (int i = 0; i < kPartRounds; ++i) {
      const unsigned long long idx = wbase + (unsigned long long)i * 32 + lane;
      const unsigned long long key = idx < cap ? tab[idx].key : 0ULL;
      const uint32_t part = key != 0ULL ? part_of(key, n_parts) : 0x100u + lane;  // empty slots match nobody
      if (n_parts <= kBallotParts) {  // a ballot per destination is cheaper than match.any for a few destinations
        for (uint32_t d = 0; d < n_parts; ++d) {
          const unsigned m = __ballot_sync(0xffffffffu, part == d);
          if (lane == 0 && m) wcount[w][d] += __popc(m);
        }
      } else {
        const unsigned peers = __match_any_sync(0xffffffffu, part);
        if (key != 0ULL && lane == __ffs(peers) - 1) wcount[w][part] += __popc(peers);
      }
      __syncwarp();
    }
  }
  __syncthreads();
  if (threadIdx.x < n_parts) {
    unsigned long long t = 0;
    for (int ww = 0; ww < 8; ++ww) t += wcount[ww][threadIdx.x];
    if (t) atomicAdd(&counts[threadIdx.x], t);
  }
}

__device__ __forceinline__ void store_entry(citam_pair* out, unsigned long long o, const PairEntry& e) { out[o] = entry_to_pair(e.key, e.acc); }
__device__ __forceinline__ void store_entry(PairEntry* out, unsigned long long o, const PairEntry& e) { out[o] = e; }

// cursors[part] starts at the part's first output index.  Out = citam_pair: the send buffer of the
// all-to-all; Out = PairEntry with n_parts == 1: the compaction in front of the ordered export.
template <typename Out>
__global__ void __launch_bounds__(256) k_partition_scatter(const PairEntry* tab, unsigned long long cap, uint32_t n_parts,
                                                           unsigned long long* cursors, Out* out,
                                                           unsigned long long out_cap) {
  __shared__ uint32_t wcount[8][kMaxParts];
  __shared__ unsigned long long s_base[kMaxParts];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const unsigned lt = (1u << lane) - 1u;
  for (unsigned long long tile = blockIdx.x; tile * kPartTile < cap; tile += gridDim.x) {
    for (int i = threadIdx.x; i < 8 * kMaxParts; i += 256) (&wcount[0][0])[i] = 0;
    __syncthreads();
    const unsigned long long wbase = tile * kPartTile + (unsigned long long)w * (32 * kPartRounds);
    PairEntry ent[kPartRounds];
    uint32_t pos[kPartRounds];  // part | rank among the warp's earlier entries of that part << 8
#pragma unroll
    for (int i = 0; i < kPartRounds; ++i) {
      const unsigned long long idx = wbase + (unsigned long long)i * 32 + lane;
      ent[i].key = 0ULL; ent[i].acc = 0ULL;
      if (idx < cap) ent[i] = tab[idx];
      const bool valid = ent[i].key != 0ULL;
      const uint32_t part = valid ? part_of(ent[i].key, n_parts) : 0x100u + lane;
      unsigned peers = 0;
      if (n_parts <= kBallotParts) {
        for (uint32_t d = 0; d < n_parts; ++d) {
          const unsigned m = __ballot_sync(0xffffffffu, part == d);
          if (part == d) peers = m;
        }
      } else {
        peers = __match_any_sync(0xffffffffu, part);
      }
      const int leader = valid ? __ffs(peers) - 1 : lane;
      uint32_t prev = 0;
      if (valid && lane == leader) {
        prev = wcount[w][part];
        wcount[w][part] = prev + __popc(peers);
      }
      prev = __shfl_sync(0xffffffffu, prev, leader);
      pos[i] = valid ? (part | ((prev + __popc(peers & lt)) << 8)) : 0xffffffffu;
      __syncwarp();
    }
    __syncthreads();
    if (threadIdx.x < n_parts) {  // one reservation per destination and tile, then the warps in order
      const int d = threadIdx.x;
      uint32_t tot = 0;
      for (int ww = 0; ww < 8; ++ww) { const uint32_t c = wcount[ww][d]; wcount[ww][d] = tot; tot += c; }
      s_base[d] = tot ? atomicAdd(&cursors[d], (unsigned long long)tot) : 0ULL;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kPartRounds; ++i) {
      if (pos[i] == 0xffffffffu) continue;
      const uint32_t part = pos[i] & 0xffu;
      const unsigned long long o = s_base[part] + wcount[w][part] + (pos[i] >> 8);
      if (o < out_cap) store_entry(out, o, ent[i]);
    }
    __syncthreads();
  }
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

// The same sums from the per-agent counter words alone (no pass over the table): events per agent
// and has-contact flags into the two arrays, sums into stats[0] (contact-steps over agents = 2 x over
// pairs), [2], [3], [4]; stats[5] = the largest per-agent contact-steps (the event field's guard).
__global__ void k_stats_counters(const unsigned long long* agent_dur, uint32_t* agent_nev, uint32_t* agent_has, int N,
                                 DevCounters* ctr) {
  int a = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long dur = 0, has = 0, nev = 0, mx = 0, mxd = 0;
  if (a < N) {
    const unsigned long long w = agent_dur[a];
    dur = w & kAgentDurMask; nev = w >> kAgentDurBits; has = nev != 0; mx = nev; mxd = dur;
    agent_nev[a] = (uint32_t)nev; agent_has[a] = (uint32_t)has;
  }
  for (int d = 16; d > 0; d >>= 1) {
    dur += __shfl_down_sync(0xffffffffu, dur, d);
    has += __shfl_down_sync(0xffffffffu, has, d);
    nev += __shfl_down_sync(0xffffffffu, nev, d);
    mx = max(mx, __shfl_down_sync(0xffffffffu, mx, d));
    mxd = max(mxd, __shfl_down_sync(0xffffffffu, mxd, d));
  }
  if ((threadIdx.x & 31) == 0 && dur) {
    atomicAdd(&ctr->stats[0], dur);
    atomicAdd(&ctr->stats[2], has);
    atomicAdd(&ctr->stats[3], nev);
    atomicMax(&ctr->stats[4], mx);
    atomicMax(&ctr->stats[5], mxd);
  }
}

__global__ void k_unpack_durations(const unsigned long long* agent_dur, int N, unsigned long long* out) {
  int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a < N) out[a] = agent_dur[a] & kAgentDurMask;
}

// Measurement aid (citam_measure_random_updates): n random read-modify-writes over a table, nothing
// else -- the rate the memory system sustains for k_accumulate's access pattern.
__global__ void __launch_bounds__(256) k_random_updates(unsigned long long* tab, unsigned long long mask, unsigned long long n,
                                                        int probe_first) {
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i0 = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += 4 * stride) {
    unsigned long long h[4], add[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {  // four independent updates in flight per thread
      h[k] = mix64(i0 + k * stride + 1) & mask;
      add[k] = 1ULL;
      if (probe_first && i0 + k * stride < n) {  // 16-byte load of the slot pair, then the reduction on the same sector
        ulonglong2 v;
        asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(tab + (h[k] & ~1ULL)));
        add[k] += (v.x == 0xdeadbeefULL);
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (i0 + k * stride < n) atomicAdd(tab + h[k], add[k]);
  }
}

// Records of another engine (rank) into this table.  Only the pair table is touched: per-agent
// contact seconds are reduced separately (all-reduce of the counters), see the header.
// kMergeIlp records per thread, first probes of all of them in flight together (latency-bound random inserts).
constexpr int kMergeIlp = 4;
__global__ void __launch_bounds__(256) k_merge_pairs(Params p, const citam_pair* in, unsigned long long n, int* bad) {
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  const unsigned long long i0 = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  citam_pair r[kMergeIlp];
  bool live[kMergeIlp];
  unsigned long long h[kMergeIlp];
  ulonglong2 e0[kMergeIlp];
  uint32_t mx = 0;
#pragma unroll
  for (int k = 0; k < kMergeIlp; ++k) {
    live[k] = i0 + k * stride < n;
    if (!live[k]) continue;
    r[k] = in[i0 + k * stride];
    if (r[k].a1 >= r[k].a2 || r[k].a2 >= (uint32_t)p.N || r[k].duration > 0xffffffULL || r[k].n_events > 0xffffu) {
      atomicOr(bad, 1);
      live[k] = false;
      continue;
    }
    h[k] = mix64(((unsigned long long)r[k].a1 << 32) | r[k].a2) & p.pair_mask;
    e0[k] = load_entry(&p.pairs[h[k]]);
  }
#pragma unroll
  for (int k = 0; k < kMergeIlp; ++k) {
    if (!live[k]) continue;
    unsigned long long old;
    const unsigned long long slot = pair_slot_from(p, ((unsigned long long)r[k].a1 << 32) | r[k].a2, h[k], e0[k], &old);
    if (slot == ~0ULL) continue;
    const uint32_t first = r[k].first_step > 0xFFFFFFu ? NO_STEP : r[k].first_step;
    if (first != NO_STEP) mx = max(mx, first);
    pair_apply_acc(p, slot, old, first, r[k].duration, r[k].n_events);
  }
  // the largest first-contact step merged so far (sort key width of the ordered export)
  mx = __reduce_max_sync(0xffffffffu, mx);
  if ((threadIdx.x & 31) == 0 && mx) atomicMax(&p.ctr->merged_max_step, (unsigned long long)mx);
}

__global__ void k_merge_coords(Params p, const citam_coord* in, unsigned long long n) {
  unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  citam_coord r = in[i];
  const FloorDev& fd = p.floors[r.floor];
  coord_upsert_rel(p, r.floor, r.sx - 2 * fd.minx, r.sy - 2 * fd.miny, r.count);
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
  int pair_tpb = 128;  // threads per block of k_pairs_changed_split (CITAM_PAIR_TPB: 128 or 256)
  int acc_ilp = 2;  // records per thread and iteration in k_accumulate (CITAM_ACC_ILP: 1, 2 or 4)
  bool explicit_states = false;  // citam_step_states in progress: states come from the caller, no windows
  bool have_itin = false;

  int S = 0;  // chunk capacity in steps
  int4* rec = nullptr;  // the current one of rec_buf
  int4* rec_buf[2] = {nullptr, nullptr};
  int cur = 0;  // buffer set of the chunk being built (records + contact buffer are double-buffered)
  cudaStream_t stream2 = nullptr;  // k_accumulate of chunk i runs here while chunk i+1 is built on `stream`
  cudaEvent_t ev_built[2] = {nullptr, nullptr}, ev_applied[2] = {nullptr, nullptr};
  bool applied_pending[2] = {false, false};
  // experiment (profiles/README.md; CITAM_DEFER_ACC=1): hold the accumulation of chunk i back until the
  // pair kernel of chunk i + 1 is launched, so that the DRAM-bound kernel runs beside the
  // instruction-bound one instead of beside the latency-bound binning kernels.  Measured: no gain
  // (107.4 vs 107.3 ms per config-3 day) -- the step is bound by the sum of its kernels' work.
  bool defer_acc = false;
  bool acc_held = false;
  int acc_held_cur = 0, acc_held_tb = 0;
  Params acc_held_p{};
  cudaEvent_t ev_go = nullptr;
  uint32_t* cells = nullptr;
  int4* sorted = nullptr;
  // split pipeline: static agents' records and classes are read by k_accumulate too -> double-buffered
  int4* srec_buf[2] = {nullptr, nullptr};
  uint32_t* cls_buf[2] = {nullptr, nullptr};
  int4* ssorted = nullptr;
  uint32_t *scells = nullptr, *dyn_agent = nullptr, *dyn_count = nullptr, *seg_cur = nullptr;
  int4* ient = nullptr;
  uint32_t *ient_rank = nullptr, *ient_count = nullptr;
  uint8_t* dmask = nullptr;
  uint32_t *work = nullptr, *work_count = nullptr;
  int chunk_steps = 0;
  int still_min = kStillMin;  // CITAM_STILL_MIN (>= kStillMin; e.g. 100000 keeps every dynamic agent in the per-step arrays)
  int dtotal_cells = 0, dcells_stride = 4;
  bool split = false;  // the chunk being processed uses the split pipeline
  bool no_fast_stats = false;  // CITAM_NO_FAST_STATS=1: statistics always from the pass over the pair table (tests)
  bool no_split = false;  // CITAM_NO_SPLIT=1: every agent is expanded and binned at every step (A/B measurements, tests)

  PairEntry* pairs = nullptr;
  unsigned long long pair_cap = 0;
  bool pair_auto = false;  // capacity chosen by the library: grown when a quarter full
  DevCounters* snap_h = nullptr;  // pinned: the counters as of the previous chunk (occupancy check without a stall)
  cudaEvent_t ev_snap = nullptr;
  bool snap_pending = false;
  unsigned long long* hist = nullptr;  // dense coordinate histogram (all floors), NULL = hashed
  long long hist_total = 0;
  int32_t range_begin = -1, range_end = INT_MAX;
  uint32_t max_step = 0;  // largest step any accumulated record can carry (sort key width)
  CoordEntry* coords = nullptr;  // allocated only when the dense histogram is not in use
  unsigned long long coord_cap = 0;
  uint4* log = nullptr;
  unsigned long long log_cap = 0;
  ContactRec* cbuf = nullptr;  // the current one of cbuf_buf
  ContactRec* cbuf_buf[2] = {nullptr, nullptr};
  unsigned long long cbuf_cap = 0;
  uint32_t *chg_list = nullptr, *chg_count = nullptr, *part_sums = nullptr;
  unsigned long long* agent_dur = nullptr;
  uint32_t *agent_nev = nullptr, *agent_has = nullptr;
  unsigned long long* dur_out = nullptr;  // unpacked contact seconds (export scratch)
  unsigned long long* part_cnt_d = nullptr;
  uint32_t* sort_ws = nullptr;
  size_t sort_ws_bytes = 0;
  void* sort_buf[2] = {nullptr, nullptr};  // ping-pong buffers of the ordered pair export, kept between calls
  size_t sort_bytes = 0;                   // (a first cudaMalloc of several GB costs hundreds of ms)
  bool counters_match = true;  // the per-agent counter words describe exactly the pairs in the table
  unsigned long long steps_since_reset = 0;
  DevCounters* ctr_d = nullptr;
  int* flag_d = nullptr;  // scratch word for device-side validation

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

static int bits_of(unsigned long long v) {
  int b = 0;
  while (v) { ++b; v >>= 1; }
  return b;
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

static int release_accumulate(citam_engine* e);

// Start building a chunk: take the other buffer set, once the accumulation that last read it is done.
static int begin_chunk(citam_engine* e) {
  if (e->acc_held && e->acc_held_cur == (e->cur ^ 1)) {  // the set about to be reused is still waiting to be applied
    int rc = release_accumulate(e);
    if (rc) return rc;
  }
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
  int rc0 = release_accumulate(e);
  if (rc0) return rc0;
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
  p.cell_magic = e->cell > 1 ? (uint32_t)((1ULL << 32) / (unsigned)e->cell) + 1u : 0u;
  p.D = e->D;
  p.s_lim = (unsigned long long)(e->s_max + 1);
  p.d_lim = e->s_max < 0 ? 0u : (uint32_t)std::min<long long>(65535, (long long)std::floor(std::sqrt((double)e->s_max)) + 1);
  p.floors = e->floors_d; p.segs = e->segs_d; p.seg_off = e->seg_off_d;
  p.window = (e->have_itin && !e->window_dirty && !e->explicit_states) ? e->window_d : nullptr;
  p.rec = e->rec; p.cells = e->cells;
  p.sorted = e->sorted; p.pairs = e->pairs; p.pair_mask = e->pair_cap - 1;
  p.coords = e->coords; p.coord_mask = e->coord_cap - 1; p.hist = e->hist;
  p.range_begin = e->range_begin; p.range_end = e->range_end;
  p.log = e->log; p.log_cap = e->log_cap;
  p.cbuf = (e->log_cap || !e->hist) ? nullptr : e->cbuf; p.cbuf_sub_cap = e->cbuf_cap / kSubBuffers;
  p.cbuf_count = e->ctr_d->cbuf_count[e->cur];
  p.agent_dur = e->agent_dur;
  const bool agg = e->log_cap == 0;  // run aggregation (off when every contact-step is logged)
  p.chg_list = agg ? e->chg_list : nullptr; p.chg_count = agg ? e->chg_count : nullptr;
  p.dcell = e->cell * kDynCellMul;
  p.dcell_magic = (uint32_t)((1ULL << 32) / (unsigned)p.dcell) + 1u;
  p.dtotal_cells = e->dtotal_cells; p.dcells_stride = e->dcells_stride;
  if (e->split) {
    p.cls = e->cls_buf[e->cur]; p.srec = e->srec_buf[e->cur];
    p.ssorted = e->ssorted; p.scells = e->scells; p.dyn_agent = e->dyn_agent; p.dyn_count = e->dyn_count; p.seg_cur = e->seg_cur;
    p.ient = e->ient; p.ient_rank = e->ient_rank; p.ient_count = e->ient_count; p.dmask = e->dmask; p.work = e->work; p.work_count = e->work_count;
    p.chunk_steps = e->chunk_steps; p.still_min = e->still_min;
  }
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
  for (int b = 0; b < 2; ++b) { cudaFree(e->srec_buf[b]); cudaFree(e->cls_buf[b]); e->srec_buf[b] = nullptr; e->cls_buf[b] = nullptr; }
  cudaFree(e->ssorted); cudaFree(e->scells); cudaFree(e->dyn_agent); cudaFree(e->dyn_count); cudaFree(e->seg_cur);
  cudaFree(e->ient); cudaFree(e->ient_rank); cudaFree(e->ient_count); cudaFree(e->dmask); cudaFree(e->work); cudaFree(e->work_count);
  e->ssorted = nullptr; e->scells = e->dyn_agent = e->dyn_count = e->seg_cur = nullptr;
  e->ient = nullptr; e->ient_rank = e->ient_count = e->work = e->work_count = nullptr; e->dmask = nullptr;
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
    int dbase = 0;
    const int dcell = e->cell * kDynCellMul;
    for (auto& f : e->floors_h) {
      f.dncx = f.lut ? (f.W + dcell - 1) / dcell : 0;
      f.dncy = f.lut ? (f.H + dcell - 1) / dcell : 0;
      f.dcell_base = dbase;
      dbase += f.dncx * f.dncy;
    }
    e->dtotal_cells = dbase;
    e->dcells_stride = ((dbase + 1) + 3) & ~3;
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
    if (!e->hist && !e->coords) {  // the hashed histogram exists only when the dense one does not
      CK(cudaMalloc(&e->coords, sizeof(CoordEntry) * e->coord_cap));
      CK(cudaMemsetAsync(e->coords, 0, sizeof(CoordEntry) * e->coord_cap, e->stream));
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
  for (int b = 0; b < 2; ++b) {
    CK(cudaMalloc(&e->srec_buf[b], sizeof(int4) * n));
    CK(cudaMalloc(&e->cls_buf[b], sizeof(uint32_t) * n));
  }
  // the static array: one entry per agent standing through the chunk, plus the dynamic agents' stays
  // (disjoint, >= kStillMin steps each: at most S / kStillMin per agent)
  const size_t n_stays = (e->log_cap == 0 && !e->no_split) ? n * (size_t)(std::min(S, kSplitMaxChunk) / kStillMin + 1) : 1;
  CK(cudaMalloc(&e->ssorted, sizeof(int4) * std::max(n, n_stays)));
  CK(cudaMalloc(&e->ient, sizeof(int4) * n_stays));
  CK(cudaMalloc(&e->ient_rank, sizeof(uint32_t) * n_stays));
  CK(cudaMalloc(&e->ient_count, sizeof(uint32_t)));
  CK(cudaMalloc(&e->work, (e->log_cap == 0 && !e->no_split) ? sizeof(uint32_t) * n * (size_t)std::min(kMaxGroups, (std::min(S, kSplitMaxChunk) + 8) / 8) : 4));
  CK(cudaMalloc(&e->work_count, sizeof(uint32_t) * kMaxGroups));
  CK(cudaMalloc(&e->dmask, (e->log_cap == 0 && !e->no_split) ? n * (size_t)S : 1));
  CK(cudaMalloc(&e->scells, sizeof(uint32_t) * (size_t)e->cells_stride));
  CK(cudaMalloc(&e->dyn_agent, sizeof(uint32_t) * n));
  CK(cudaMalloc(&e->dyn_count, sizeof(uint32_t)));
  CK(cudaMalloc(&e->seg_cur, sizeof(uint32_t) * n));
  CK(cudaMemsetAsync(e->seg_cur, 0, sizeof(uint32_t) * n, e->stream));
  e->S = S;
  return CITAM_OK;
}

static int pick_chunk(citam_engine* e, int steps) {
  const bool split_mode = e->log_cap == 0 && !e->no_split;
  if (e->cfg.steps_per_chunk > 0) return std::min(std::min(steps, e->cfg.steps_per_chunk), split_mode ? kSplitMaxChunk : kMaxChunk);
  // ~64M agent-steps per launch group (3.3 GiB of per-step records), bounded by 1 GiB of histogram rows
  const bool split = e->log_cap == 0 && !e->no_split;
  long long by_agents = std::max(1LL, ((split ? 512LL : 64LL) << 20) / std::max(1, e->N));
  long long by_cells = std::max(1LL, (1LL << 28) / std::max(4, e->cells_stride));
  long long s = std::min(by_agents, by_cells);
  // without the contact log (split pipeline) an agent is binned once per chunk if it stands still for
  // the whole chunk: short chunks keep most agents in that class
  s = std::min<long long>(s, e->log_cap == 0 ? kSplitDefaultChunk : kMaxChunk);
  s = std::max<long long>(s, 1);
  return (int)std::min<long long>(s, steps);
}

// per-row exclusive scan of `rows` histogram rows of n = cells + 1 words (stride words apart)
static void launch_scan(citam_engine* e, uint32_t* cells, int stride, int n, int rows) {
  Timed tm(e, CITAM_K_SCAN);
  const int n4 = (n + 3) >> 2;
  int parts = std::max(1, std::min(kMaxScanParts, std::min((296 + rows - 1) / rows, (n4 + 4095) / 4096)));
  if (rows >= 64 && n4 <= 16384) parts = 1;  // enough rows to fill the machine: one block per row, no second pass
  int part_len4 = (n4 + parts - 1) / parts;
  parts = (n4 + part_len4 - 1) / part_len4;
  k_scan<<<dim3(parts, rows), kScanThreads, 0, e->stream>>>(cells, stride, n, part_len4, e->part_sums);
  if (parts > 1) {
    e->launches++;
    k_scan_add<<<dim3(parts - 1, rows), 256, 0, e->stream>>>(cells, stride, n, part_len4, e->part_sums);
  }
}

// the chunk's contact buffer is applied on the side stream, overlapping the next chunk's build
static int issue_accumulate(citam_engine* e, const Params& p, int t_begin, int cur, cudaEvent_t after) {
  CK(cudaStreamWaitEvent(e->stream2, after, 0));
  {
    Timed tm(e, CITAM_K_ACCUMULATE, e->stream2, true);
    const int nb = kSubBuffers * e->acc_blocks;
    if (e->acc_ilp >= 4) k_accumulate<4><<<nb, 256, 0, e->stream2>>>(p, t_begin);
    else if (e->acc_ilp >= 2) k_accumulate<2><<<nb, 256, 0, e->stream2>>>(p, t_begin);
    else k_accumulate<1><<<nb, 256, 0, e->stream2>>>(p, t_begin);
    k_reset_cbuf<<<1, kSubBuffers, 0, e->stream2>>>(p.cbuf_count);
    e->launches++;
  }
  CK(cudaEventRecord(e->ev_applied[cur], e->stream2));
  e->applied_pending[cur] = true;
  return CITAM_OK;
}

// Launch the accumulation that is being held back (see citam_engine::defer_acc): it starts once
// everything enqueued so far on the engine's stream is done.
static int release_accumulate(citam_engine* e) {
  if (!e->acc_held) return CITAM_OK;
  e->acc_held = false;
  CK(cudaEventRecord(e->ev_go, e->stream));
  return issue_accumulate(e, e->acc_held_p, e->acc_held_tb, e->acc_held_cur, e->ev_go);
}

static int launch_accumulate(citam_engine* e, const Params& p, int t_begin) {
  if (p.cbuf == nullptr) return CITAM_OK;
  if (e->split && e->defer_acc) {
    int rc = release_accumulate(e);  // at most one is held
    if (rc) return rc;
    e->acc_held = true; e->acc_held_cur = e->cur; e->acc_held_tb = t_begin; e->acc_held_p = p;
    return CITAM_OK;
  }
  CK(cudaEventRecord(e->ev_built[e->cur], e->stream));
  return issue_accumulate(e, p, t_begin, e->cur, e->ev_built[e->cur]);
}

// bin (already done by caller) -> scan -> scatter -> pairs for `rows` steps starting at t_begin
static int run_sorted_stages(citam_engine* e, int t_begin, int rows) {
  Params p = make_params(e);
  p.chunk_begin = t_begin;
  launch_scan(e, e->cells, e->cells_stride, e->total_cells + 1, rows);
  const bool agg = p.chg_count != nullptr;
  // the first step of a range is evaluated in full (k_pairs); every other step through its changed agents
  const int row0 = (t_begin == e->range_begin) ? 1 : 0;
  {
    Timed tm(e, CITAM_K_SCATTER);
    dim3 g((e->N + 255) / 256, rows);
    k_scatter<<<g, 256, 0, e->stream>>>(p, rows, t_begin);
    if (agg && rows > row0) {
      e->launches++;
      dim3 g3((unsigned)std::min(296, std::max(1, (e->N + 255) / 256)), rows - row0);
      k_select_changed<<<g3, 256, 0, e->stream>>>(p, row0, rows);
    }
  }
  {
    Timed tm(e, CITAM_K_PAIRS);
    bool launched = false;
    if (!agg || row0 == 1) {
      dim3 g((e->N + 255) / 256, agg ? 1 : rows);
      k_pairs<<<g, 256, 0, e->stream>>>(p, t_begin, rows, agg ? 1 : 0, 0);
      launched = true;
    }
    if (agg && rows > row0) {
      if (launched) e->launches++;
      dim3 g2((unsigned)std::min(592, std::max(1, (e->N + 255) / 256)), rows - row0);
      k_pairs_changed<<<g2, 256, 0, e->stream>>>(p, t_begin, row0, rows);
    }
  }
  int rc = launch_accumulate(e, p, t_begin);
  if (rc) return rc;
  CK(cudaGetLastError());
  return CITAM_OK;
}

// One chunk of the split pipeline (citam_run without the contact log): agents that stand still from
// the halo step to the chunk's end are binned ONCE (fine grid); the others -- walking, arriving,
// leaving, or changing segment inside the chunk -- are expanded, binned (coarse grid) and sorted per
// step.  Only a range's first step looks at static-static pairs (k_pairs on the static array): at every
// other step such a pair is inside a run that an earlier step added in one piece.
static int run_split_chunk(citam_engine* e, int tb, int steps) {
  e->chunk_steps = steps;
  Params p = make_params(e);
  p.chunk_begin = tb;
  const int rows = steps + 1;  // + halo row for step tb-1
  CK(cudaMemsetAsync(e->ient_count, 0, sizeof(uint32_t), e->stream));
  CK(cudaMemsetAsync(e->work_count, 0, sizeof(uint32_t) * kMaxGroups, e->stream));
  CK(cudaMemsetAsync(e->scells, 0, sizeof(uint32_t) * (size_t)e->cells_stride, e->stream));
  CK(cudaMemsetAsync(e->cells, 0, sizeof(uint32_t) * (size_t)e->dcells_stride * steps, e->stream));
  CK(cudaMemsetAsync(e->chg_count, 0, sizeof(uint32_t) * 32 * steps, e->stream));
  CK(cudaMemsetAsync(e->dyn_count, 0, sizeof(uint32_t), e->stream));
  const int gN = (e->N + 255) / 256;
  const int wide = 148 * 8;  // grid-stride kernels over the (device-side) dynamic count
  const int groups = (rows + 7) >> 3;
  {
    Timed tm(e, CITAM_K_CLASSIFY);
    k_classify<<<gN, 256, 0, e->stream>>>(p, tb, steps);
    k_classify_dyn<<<wide / 2, 256, 0, e->stream>>>(p, tb, steps);
    e->launches++;
  }
  {
    Timed tm(e, CITAM_K_EXPAND_BIN);
    k_expand_dyn<<<dim3(96, groups), 128, 0, e->stream>>>(p, tb - 1, rows);
  }
  launch_scan(e, e->scells, e->cells_stride, e->total_cells + 1, 1);
  launch_scan(e, e->cells, e->dcells_stride, e->dtotal_cells + 1, steps);
  const bool first = tb == e->range_begin;
  {
    Timed tm(e, CITAM_K_SCATTER);
    k_scatter_static<<<gN, 256, 0, e->stream>>>(p);
    k_scatter_dyn<<<dim3(48, steps), 256, 0, e->stream>>>(p, steps);
    k_select_changed_dyn<<<dim3(64, steps), 256, 0, e->stream>>>(p, steps, first ? 0 : -1);
    e->launches += 2;
  }
  {
    int rc = release_accumulate(e);  // the previous chunk's accumulation runs beside this chunk's pair kernel
    if (rc) return rc;
  }
  {
    Timed tm(e, CITAM_K_PAIRS);
    if (first) {
      // the static array holds up to steps / kStillMin + 1 entries per agent (blocks past its end return at once)
      const unsigned gS = (unsigned)(((size_t)e->N * (size_t)(steps / kStillMin + 1) + 255) / 256);
      k_pairs<<<dim3(gS, 1), 256, 0, e->stream>>>(p, tb, 1, 1, 1);
      e->launches++;
    }
    if (e->pair_tpb == 128) k_pairs_changed_split<128><<<dim3(128, steps), 128, 0, e->stream>>>(p, tb, steps, first ? 0 : -1);
    else k_pairs_changed_split<256><<<dim3(64, steps), 256, 0, e->stream>>>(p, tb, steps, first ? 0 : -1);
  }
  int rc = launch_accumulate(e, p, tb);
  if (rc) return rc;
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

// Replace the pair table by one of `ncap` slots holding the same entries.  The caller has joined the side stream.
static int rehash_pairs(citam_engine* e, unsigned long long ncap) {
  PairEntry* np = nullptr;
  if (cudaMalloc(&np, sizeof(PairEntry) * ncap) != cudaSuccess) {
    cudaGetLastError();
    return CITAM_ERR_CAPACITY;  // not fatal for the callers that grow opportunistically
  }
  CK(cudaMemsetAsync(np, 0, sizeof(PairEntry) * ncap, e->stream));
  PairEntry* old = e->pairs;
  unsigned long long ocap = e->pair_cap;
  e->pairs = np; e->pair_cap = ncap;
  CK(cudaMemsetAsync(&e->ctr_d->pair_used, 0, sizeof(unsigned long long), e->stream));
  CK(cudaMemsetAsync(e->ctr_d->pair_used_sub, 0, sizeof(unsigned long long) * kSubBuffers * kCounterStride, e->stream));
  Params p = make_params(e);
  {
    Timed tm(e, CITAM_K_MISC);
    k_rehash_pairs<<<(unsigned)((ocap + 255) / 256), 256, 0, e->stream>>>(p, old, ocap);
  }
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(e->stream));
  cudaFree(old);
  return CITAM_OK;
}

// Library-chosen pair capacity: the table is doubled until it is at most a quarter full.  Between
// chunks the occupancy comes from an asynchronous snapshot of the counters taken one chunk earlier
// (no stall: k_accumulate keeps overlapping the next chunk's build); `now` forces a fresh read.
static int maybe_grow_pairs(citam_engine* e, bool now) {
  if (!e->pair_auto) return CITAM_OK;
  unsigned long long used = 0;
  bool have = false, overflowed = false;
  if (now) {
    int rcj = join_side_stream(e);
    if (rcj) return rcj;
    if (e->snap_pending) CK(cudaEventSynchronize(e->ev_snap));  // an older snapshot must not land after this one
    CK(cudaMemcpyAsync(e->snap_h, e->ctr_d, sizeof(DevCounters), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    e->snap_pending = false;
    have = true;
  } else if (e->snap_pending && cudaEventQuery(e->ev_snap) == cudaSuccess) {
    e->snap_pending = false;
    have = true;
  }
  if (have) {
    used = pairs_used(*e->snap_h);
    overflowed = (e->snap_h->overflow & OVF_PAIR) != 0;
    if (used * 4 > e->pair_cap && !overflowed) {
      int rcj = join_side_stream(e);
      if (rcj) return rcj;
      unsigned long long ncap = e->pair_cap;
      while (used * 4 > ncap) ncap *= 2;
      int rc = rehash_pairs(e, ncap);
      if (rc == CITAM_ERR_CAPACITY) e->pair_auto = false;  // out of memory: keep the table, overflow is reported if it comes
      else if (rc) return rc;
    }
  }
  if (!now && !e->snap_pending) {
    // counters as they are once everything enqueued so far on both streams is done
    cudaStream_t s = e->applied_pending[e->cur] ? e->stream2 : e->stream;
    CK(cudaMemcpyAsync(e->snap_h, e->ctr_d, sizeof(DevCounters), cudaMemcpyDeviceToHost, s));
    CK(cudaEventRecord(e->ev_snap, s));
    e->snap_pending = true;
  }
  return CITAM_OK;
}

static int create_impl(citam_engine* e, const citam_config* cfg) {
  if (const char* g = getenv("CITAM_ACC_BLOCKS")) e->acc_blocks = std::max(1, atoi(g));
  if (const char* g = getenv("CITAM_ACC_ILP")) e->acc_ilp = atoi(g);
  if (const char* g = getenv("CITAM_PAIR_TPB")) e->pair_tpb = atoi(g);
  if (const char* g = getenv("CITAM_STILL_MIN")) e->still_min = std::max(kStillMin, atoi(g));
  if (const char* g = getenv("CITAM_NO_SPLIT")) e->no_split = atoi(g) != 0;
  if (const char* g = getenv("CITAM_NO_FAST_STATS")) e->no_fast_stats = atoi(g) != 0;
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
  size_t n = (size_t)e->N;
  CK(cudaMalloc(&e->floors_d, sizeof(FloorDev) * e->F));
  CK(cudaMalloc(&e->pairs, sizeof(PairEntry) * e->pair_cap));
  if (e->log_cap) CK(cudaMalloc(&e->log, sizeof(uint4) * e->log_cap));
  CK(cudaMalloc(&e->agent_dur, sizeof(unsigned long long) * n));
  CK(cudaMalloc(&e->agent_nev, sizeof(uint32_t) * n));
  CK(cudaMalloc(&e->agent_has, sizeof(uint32_t) * n));
  CK(cudaMalloc(&e->ctr_d, sizeof(DevCounters)));
  CK(cudaMalloc(&e->flag_d, sizeof(int)));
  CK(cudaMalloc(&e->prev, sizeof(int4) * n));
  CK(cudaMalloc(&e->states_d, sizeof(citam_agent_state) * n));
  CK(cudaMallocHost(&e->snap_h, sizeof(DevCounters)));
  CK(cudaEventCreateWithFlags(&e->ev_snap, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&e->ev_go, cudaEventDisableTiming));
  if (const char* g = getenv("CITAM_DEFER_ACC")) e->defer_acc = atoi(g) != 0;
  CK(cudaStreamCreateWithFlags(&e->stream2, cudaStreamNonBlocking));
  for (int b = 0; b < 2; ++b) {
    CK(cudaEventCreateWithFlags(&e->ev_built[b], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&e->ev_applied[b], cudaEventDisableTiming));
  }
  if (const char* g = getenv("CITAM_L2_PERSIST")) {
    // experiment (profiles/README.md): pin the per-agent counter words (2 reductions per applied record)
    // in the persisting part of L2 for the accumulation stream
    if (atoi(g) != 0) {
      cudaDeviceProp prop;
      CK(cudaGetDeviceProperties(&prop, e->device));
      size_t bytes = std::min<size_t>(sizeof(unsigned long long) * n, (size_t)prop.accessPolicyMaxWindowSize);
      CK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, std::min<size_t>(bytes, (size_t)prop.persistingL2CacheMaxSize)));
      cudaStreamAttrValue v{};
      v.accessPolicyWindow.base_ptr = e->agent_dur;
      v.accessPolicyWindow.num_bytes = bytes;
      v.accessPolicyWindow.hitRatio = 1.0f;
      v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
      v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
      CK(cudaStreamSetAttribute(e->stream2, cudaStreamAttributeAccessPolicyWindow, &v));
    }
  }
  {
    int bad = 1;
    k_check_s_max<<<1, 1, 0, e->stream>>>(e->D, e->s_max, e->flag_d);
    CK(cudaMemcpyAsync(&bad, e->flag_d, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    if (bad) return fail(CITAM_ERR_CUDA, "integer distance threshold %lld disagrees with the float64 predicate", e->s_max);
  }
  return citam_reset(e);
}

// Device sort of `n` 16-byte records in `a` (scratch `b`); *res points at the sorted copy.
template <typename KeyFn>
static int device_sort(citam_engine* e, uint4* a, uint4* b, unsigned long long n, const rsort::Pass* passes, int np,
                       KeyFn key, uint4** res) {
  *res = a;
  if (n < 2 || np == 0) return CITAM_OK;
  if (n >= (1ULL << 32)) return fail(CITAM_ERR_CAPACITY, "device sort holds fewer than 2^32 records");
  const size_t ws_bytes = rsort::workspace_bytes(n);
  if (e->sort_ws_bytes < ws_bytes) {  // kept between calls
    cudaFree(e->sort_ws);
    e->sort_ws = nullptr; e->sort_ws_bytes = 0;
    CK(cudaMalloc(&e->sort_ws, ws_bytes + ws_bytes / 8));
    e->sort_ws_bytes = ws_bytes + ws_bytes / 8;
  }
  uint32_t* ws = e->sort_ws;
  cudaError_t err;
  {
    Timed tm(e, CITAM_K_FINALIZE);
    err = rsort::sort(a, b, n, passes, np, ws, e->stream, key, res, &e->launches);
  }
  if (err == cudaSuccess) err = cudaStreamSynchronize(e->stream);
  if (err != cudaSuccess) return fail(CITAM_ERR_CUDA, "device sort failed: %s", cudaGetErrorString(err));
  return CITAM_OK;
}

// One flag word, set by a device-side check: returns its value (and clears it).
static int read_flag(citam_engine* e, int* v) {
  CK(cudaMemcpyAsync(v, e->flag_d, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  return CITAM_OK;
}

extern "C" {

int citam_abi_version(void) { return CITAM_ABI_VERSION; }
const char* citam_last_error(void) { return g_err.c_str(); }

int citam_create(const citam_config* cfg, citam_engine** out) {
  if (!cfg || !out) return fail(CITAM_ERR_INVALID, "null argument");
  *out = nullptr;
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
  int rc = create_impl(e, cfg);
  if (rc) {  // nothing of a half-built engine survives
    std::string msg = g_err;
    citam_destroy(e);
    g_err = msg;
    return rc;
  }
  *out = e;
  return CITAM_OK;
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
  if (e->ev_snap) cudaEventDestroy(e->ev_snap);
  if (e->ev_go) cudaEventDestroy(e->ev_go);
  if (e->snap_h) cudaFreeHost(e->snap_h);
  for (auto p : e->lut_d) cudaFree(p);
  for (auto p : e->adj_d) cudaFree(p);
  cudaFree(e->floors_d); cudaFree(e->segs_d); cudaFree(e->raw_d); cudaFree(e->seg_off_d); cudaFree(e->window_d); cudaFree(e->pairs);
  cudaFree(e->coords); cudaFree(e->hist);
  cudaFree(e->log); cudaFree(e->agent_dur); cudaFree(e->agent_nev); cudaFree(e->agent_has); cudaFree(e->dur_out); cudaFree(e->sort_buf[0]); cudaFree(e->sort_buf[1]); cudaFree(e->part_cnt_d); cudaFree(e->sort_ws); cudaFree(e->ctr_d);
  cudaFree(e->flag_d); cudaFree(e->prev); cudaFree(e->states_d);
  cudaGetLastError();
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
  int rc = join_side_stream(e);
  if (rc) return rc;
  CK(cudaMemsetAsync(e->pairs, 0, sizeof(PairEntry) * e->pair_cap, e->stream));
  if (e->coords) CK(cudaMemsetAsync(e->coords, 0, sizeof(CoordEntry) * e->coord_cap, e->stream));
  if (e->hist) CK(cudaMemsetAsync(e->hist, 0, sizeof(unsigned long long) * (size_t)e->hist_total, e->stream));
  CK(cudaMemsetAsync(e->agent_dur, 0, sizeof(unsigned long long) * (size_t)e->N, e->stream));
  CK(cudaMemsetAsync(e->ctr_d, 0, sizeof(DevCounters), e->stream));
  e->snap_pending = false;
  e->prev_step = LLONG_MIN;
  e->agent_steps = 0;
  e->max_step = 0;
  e->counters_match = true;
  e->steps_since_reset = 0;
  return CITAM_OK;
}

int citam_clear_pairs(citam_engine* e) {
  if (!e) return fail(CITAM_ERR_INVALID, "null engine");
  CK(cudaSetDevice(e->device));
  int rc = join_side_stream(e);
  if (rc) return rc;
  CK(cudaMemsetAsync(e->pairs, 0, sizeof(PairEntry) * e->pair_cap, e->stream));
  CK(cudaMemsetAsync(&e->ctr_d->pair_used, 0, sizeof(unsigned long long), e->stream));
  CK(cudaMemsetAsync(e->ctr_d->pair_used_sub, 0, sizeof(unsigned long long) * kSubBuffers * kCounterStride, e->stream));
  e->snap_pending = false;
  e->counters_match = false;
  return CITAM_OK;
}

static int install_lut(citam_engine* e, int floor, int minx, int miny, int W, int H, int n_spaces) {
  if (floor < 0 || floor >= e->F) return fail(CITAM_ERR_INVALID, "floor %d out of range", floor);
  if (W <= 0 || H <= 0 || W > 65536 || H > 65536 || (long long)W * H > (1LL << 31))
    return fail(CITAM_ERR_INVALID, "bad lattice %dx%d (sides <= 65536, at most 2^31 points)", W, H);
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
  if (n_segs < 0 || seg_off[0] != 0 || seg_off[e->N] != n_segs) return fail(CITAM_ERR_INVALID, "seg_off must run from 0 to n_segs");
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
  CK(cudaMemsetAsync(e->flag_d, 0, sizeof(int), e->stream));
  CK(cudaMemcpyAsync(e->seg_off_d, seg_off, sizeof(long long) * ((size_t)e->N + 1), cudaMemcpyHostToDevice, e->stream));
  // `segs` may be a host or a DEVICE pointer (unified addressing): a multi-GPU run uploads 1/G of the store per
  // rank and all-gathers it over NVLink (distributed.ShardedRun.load_store)
  if (n_segs) CK(cudaMemcpyAsync(raw, segs, sizeof(citam_segment) * n_segs, cudaMemcpyDefault, e->stream));
  {
    // the documented limits, checked on the device copy (no host pass over 10^7-10^8 records)
    Timed tm(e, CITAM_K_MISC);
    k_validate_segments<<<(e->N + 255) / 256, 256, 0, e->stream>>>(e->seg_off_d, raw, n_segs, e->N, e->flag_d);
    if (n_segs) k_repack_segments<<<(unsigned)((n_segs + 255) / 256), 256, 0, e->stream>>>(raw, n_segs, e->segs_d);
  }
  CK(cudaGetLastError());
  int bad = 0;
  int rc = read_flag(e, &bad);  // also: the caller's host buffers are free to change again
  if (rc) return rc;
  if (bad)
    return fail(CITAM_ERR_INVALID, "itinerary store violates the documented limits:%s%s%s%s%s",
                (bad & 1) ? " seg_off is not non-decreasing inside [0, n_segs];" : "",
                (bad & 2) ? " t0 must increase strictly per agent and stay in [0, 2^24 - 3];" : "",
                (bad & 4) ? " floor must be in 0..62;" : "",
                (bad & 16) ? " |coordinates| must be below 2^26;" : "",
                (bad & 8) ? " an agent's last segment must have zero velocity (the sticky tail);" : "");
  e->n_segs = n_segs;
  e->have_itin = true;
  e->window_dirty = true;
  return CITAM_OK;
}

int citam_run(citam_engine* e, int32_t t_begin, int32_t t_end) {
  if (!e) return fail(CITAM_ERR_INVALID, "null engine");
  if (!e->have_itin) return fail(CITAM_ERR_STATE, "citam_run before citam_load_itineraries");
  if (t_begin < 0 || t_end < t_begin || t_end > (1 << 24) - 1) return fail(CITAM_ERR_INVALID, "bad step range (steps < 2^24 - 1)");
  if (t_end == t_begin) return CITAM_OK;
  if ((double)(e->N - 1) * (double)(e->steps_since_reset + (unsigned long long)(t_end - t_begin)) >= 1099511627776.0)
    return fail(CITAM_ERR_CAPACITY, "(n_agents - 1) x steps since the last reset must stay below 2^40 (per-agent counter width)");
  e->steps_since_reset += (unsigned long long)(t_end - t_begin);
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
  e->max_step = std::max<uint32_t>(e->max_step, (uint32_t)t_end);
  // static / dynamic split of the agents per chunk: whenever runs are aggregated (no contact log)
  struct Split { bool& f; Split(bool& x, bool v) : f(x) { f = v; } ~Split() { f = false; } } split_scope(e->split, e->log_cap == 0 && !e->no_split);
  // a contact run is one table update however many chunks it crosses; runs are clipped to the range
  // (and to 2^20 steps, the width of the run field: longer ranges are processed as several)
  for (int rb = t_begin; rb < t_end; rb += (1 << kMaxRunLog2)) {
    e->range_begin = rb;
    e->range_end = std::min(t_end, rb + (1 << kMaxRunLog2));
    for (int tb = rb; tb < e->range_end; tb += S) {
      rc = begin_chunk(e);
      if (rc) return rc;
      int steps = std::min(S, e->range_end - tb);
      if (e->split) {
        rc = run_split_chunk(e, tb, steps);
        if (rc) return rc;
        e->agent_steps += (unsigned long long)steps * e->N;
        rc = maybe_grow_pairs(e, false);
        if (rc) return rc;
        continue;
      }
      Params p = make_params(e);
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
      rc = maybe_grow_pairs(e, false);
      if (rc) return rc;
    }
  }
  e->prev_step = LLONG_MIN;
  rc = join_side_stream(e);
  if (rc) return rc;
  return maybe_grow_pairs(e, true);
}

int citam_step_states(citam_engine* e, int32_t step, const citam_agent_state* states) {
  if (!e || !states) return fail(CITAM_ERR_INVALID, "null argument");
  if (step < 0 || step > (1 << 24) - 2) return fail(CITAM_ERR_INVALID, "step out of range");
  if ((double)(e->N - 1) * (double)(e->steps_since_reset + 1) >= 1099511627776.0)
    return fail(CITAM_ERR_CAPACITY, "(n_agents - 1) x steps since the last reset must stay below 2^40 (per-agent counter width)");
  e->steps_since_reset += 1;
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
    e->range_begin = step;
  }
  e->range_end = INT_MAX;  // explicit states: every agent's `until` is step + 1
  e->max_step = std::max<uint32_t>(e->max_step, (uint32_t)step + 1);
  CK(cudaMemsetAsync(e->cells, 0, sizeof(uint32_t) * (size_t)e->cells_stride, e->stream));
  CK(cudaMemsetAsync(e->chg_count, 0, sizeof(uint32_t) * 32, e->stream));
  Params p = make_params(e);
  {
    Timed tm(e, CITAM_K_EXPAND_BIN);
    k_states_to_rec<<<g, 256, 0, e->stream>>>(p, e->states_d, e->rec + n, step, 1);
  }
  CK(cudaGetLastError());
  rc = run_sorted_stages(e, step, 1);
  if (rc) return rc;
  CK(cudaMemcpyAsync(e->prev, e->rec + n, sizeof(int4) * n, cudaMemcpyDeviceToDevice, e->stream));
  e->prev_step = step;
  e->agent_steps += e->N;
  rc = join_side_stream(e);
  if (rc) return rc;
  rc = maybe_grow_pairs(e, true);
  if (rc) return rc;
  return check_overflow(e);
}

int citam_states_at(citam_engine* e, int32_t step, citam_agent_state* out) {
  if (!e || !out) return fail(CITAM_ERR_INVALID, "null argument");
  if (!e->have_itin) return fail(CITAM_ERR_STATE, "no itineraries loaded");
  CK(cudaSetDevice(e->device));
  int rc = ensure_ready(e, 1);
  if (rc) return rc;
  rc = join_side_stream(e);
  if (rc) return rc;
  Params p = make_params(e);
  {
    Timed tm(e, CITAM_K_MISC);
    dim3 g((e->N + 127) / 128, 1);
    k_expand_bin<<<g, 128, 0, e->stream>>>(p, step, 1, 0, 16);
    k_rec_to_states<<<(e->N + 255) / 256, 256, 0, e->stream>>>(p, e->rec, e->N, e->states_d);
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
  rc = join_side_stream(e);
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
      k_rec_to_states<<<(unsigned)((tot + 255) / 256), 256, 0, e->stream>>>(p, e->rec, (long long)tot, tmp);
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
  out->binned_agent_steps = c.binned;
  return CITAM_OK;
}

int citam_pair_count(citam_engine* e, uint64_t* n) {
  if (!n) return fail(CITAM_ERR_INVALID, "null argument");
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
  int rc = ensure_ready(e, 1);
  if (rc) return rc;
  if (e->hist) return dense_coord_count(e, n);
  citam_counters c;
  rc = citam_get_counters(e, &c);
  if (rc) return rc;
  *n = c.coord_slots_used;
  return CITAM_OK;
}

// The pair table in the reference's row order -- first contact step, then (a1, a2): dict insertion
// order of contacts.py:135-142 with pairs of one step visited row-major
// (close_contacts_calculator.py:104-124) -- into a DEVICE buffer.
int citam_export_pairs_device(citam_engine* e, citam_pair* out_device, uint64_t capacity, uint64_t* n) {
  if (!e || !n || (capacity && !out_device)) return fail(CITAM_ERR_INVALID, "null argument");
  CK(cudaSetDevice(e->device));
  uint64_t have = 0;
  int rc = citam_pair_count(e, &have);
  if (rc) return rc;
  *n = have;
  if (have == 0) return CITAM_OK;
  if (capacity < have) return fail(CITAM_ERR_CAPACITY, "output holds %llu pairs, %llu needed", (unsigned long long)capacity, (unsigned long long)have);
  if (e->sort_bytes < sizeof(PairEntry) * have) {
    cudaFree(e->sort_buf[0]); cudaFree(e->sort_buf[1]);
    e->sort_buf[0] = e->sort_buf[1] = nullptr;
    e->sort_bytes = 0;
    const size_t want = sizeof(PairEntry) * (have + have / 8);
    if (cudaMalloc(&e->sort_buf[0], want) != cudaSuccess || cudaMalloc(&e->sort_buf[1], want) != cudaSuccess) {
      cudaGetLastError();
      cudaFree(e->sort_buf[0]);
      e->sort_buf[0] = e->sort_buf[1] = nullptr;
      return fail(CITAM_ERR_CUDA, "out of device memory for the export of %llu pairs", (unsigned long long)have);
    }
    e->sort_bytes = want;
  }
  PairEntry *a = static_cast<PairEntry*>(e->sort_buf[0]), *b = static_cast<PairEntry*>(e->sort_buf[1]);
  CK(cudaMemsetAsync(&e->ctr_d->export_cursor, 0, sizeof(unsigned long long), e->stream));
  {
    Timed tm(e, CITAM_K_FINALIZE);
    const unsigned blocks = (unsigned)std::min<unsigned long long>(148 * 8, (e->pair_cap + kPartTile - 1) / kPartTile);
    k_partition_scatter<PairEntry><<<blocks, 256, 0, e->stream>>>(e->pairs, e->pair_cap, 1u, &e->ctr_d->export_cursor, a, have);
  }
  CK(cudaGetLastError());
  rsort::Pass passes[12];
  int np = 0;
  const int ab = bits_of((unsigned long long)e->N - 1);
  np += rsort::plan_word(passes + np, 0, ab);
  np += rsort::plan_word(passes + np, 1, ab);
  np += rsort::plan_word(passes + np, 2, std::min(24, bits_of(e->max_step)));
  uint4* res = nullptr;
  rc = device_sort(e, reinterpret_cast<uint4*>(a), reinterpret_cast<uint4*>(b), have, passes, np, PairKey(), &res);
  if (rc == CITAM_OK) {
    Timed tm(e, CITAM_K_FINALIZE);
    k_entries_to_pairs<<<(unsigned)((have + 255) / 256), 256, 0, e->stream>>>(reinterpret_cast<PairEntry*>(res), have, out_device);
    cudaError_t err = cudaStreamSynchronize(e->stream);
    if (err != cudaSuccess) rc = fail(CITAM_ERR_CUDA, "pair export failed: %s", cudaGetErrorString(err));
  }
  return rc;
}

int citam_export_pairs(citam_engine* e, citam_pair* out, uint64_t capacity, uint64_t* n) {
  if (!e || !n || (capacity && !out)) return fail(CITAM_ERR_INVALID, "null argument");
  CK(cudaSetDevice(e->device));
  uint64_t have = 0;
  int rc = citam_pair_count(e, &have);
  if (rc) return rc;
  *n = have;
  if (have == 0) return CITAM_OK;
  if (capacity < have) return fail(CITAM_ERR_CAPACITY, "output holds %llu pairs, %llu needed", (unsigned long long)capacity, (unsigned long long)have);
  citam_pair* tmp = nullptr;
  CK(cudaMalloc(&tmp, sizeof(citam_pair) * have));
  rc = citam_export_pairs_device(e, tmp, have, n);
  if (rc == CITAM_OK) {
    cudaError_t err = cudaMemcpyAsync(out, tmp, sizeof(citam_pair) * have, cudaMemcpyDeviceToHost, e->stream);
    if (err == cudaSuccess) err = cudaStreamSynchronize(e->stream);
    if (err != cudaSuccess) rc = fail(CITAM_ERR_CUDA, "pair export failed: %s", cudaGetErrorString(err));
  }
  cudaFree(tmp);
  return rc;
}

// Coordinate records ordered by (floor, sx, sy), into a DEVICE buffer.
int citam_export_coords_device(citam_engine* e, citam_coord* out_device, uint64_t capacity, uint64_t* n) {
  if (!e || !n || (capacity && !out_device)) return fail(CITAM_ERR_INVALID, "null argument");
  int rc = check_overflow(e);
  if (rc) return rc;
  uint64_t have = 0;
  rc = citam_coord_count(e, &have);
  if (rc) return rc;
  *n = have;
  if (have == 0) return CITAM_OK;
  if (capacity < have) return fail(CITAM_ERR_CAPACITY, "output holds %llu coords, %llu needed", (unsigned long long)capacity, (unsigned long long)have);
  if (e->hist) {
    // ordered compaction: per-CTA counts of every floor, one scan, then the writes
    std::vector<unsigned long long> first(e->F + 1, 0);
    for (int f = 0; f < e->F; ++f) {
      long long nb = (long long)e->floors_h[f].hist_w * e->floors_h[f].hist_h;
      first[f + 1] = first[f] + (unsigned long long)((nb + kHistThreads - 1) / kHistThreads);
    }
    const unsigned long long nblk = first[e->F];
    uint32_t *counts = nullptr, *scratch = nullptr;
    CK(cudaMalloc(&counts, sizeof(uint32_t) * std::max<unsigned long long>(1, nblk)));
    CK(cudaMalloc(&scratch, sizeof(uint32_t) * rsort::scan_scratch_elems(nblk)));
    cudaError_t err = cudaSuccess;
    {
      Timed tm(e, CITAM_K_FINALIZE);
      for (int pass = 0; pass < 2; ++pass) {
        for (int f = 0; f < e->F; ++f) {
          const FloorDev& fd = e->floors_h[f];
          long long nb = (long long)fd.hist_w * fd.hist_h;
          if (nb == 0) continue;
          k_hist_ordered<<<(unsigned)(first[f + 1] - first[f]), kHistThreads, 0, e->stream>>>(
              e->hist + fd.hist_base, nb, fd.hist_w, fd.hist_h, f, 2 * fd.minx, 2 * fd.miny, counts + first[f],
              pass ? out_device : nullptr, have);
          e->launches++;
        }
        if (pass == 0) err = rsort::scan_u32(counts, nblk, scratch, e->stream);
        if (err != cudaSuccess) break;
      }
    }
    if (err == cudaSuccess) err = cudaStreamSynchronize(e->stream);
    cudaFree(counts); cudaFree(scratch);
    if (err != cudaSuccess) return fail(CITAM_ERR_CUDA, "coordinate export failed: %s", cudaGetErrorString(err));
    return CITAM_OK;
  }
  CoordEntry *a = nullptr, *b = nullptr;
  CK(cudaMalloc(&a, sizeof(CoordEntry) * have));
  CK(cudaMalloc(&b, sizeof(CoordEntry) * have));
  CK(cudaMemsetAsync(&e->ctr_d->export_cursor, 0, sizeof(unsigned long long), e->stream));
  {
    Timed tm(e, CITAM_K_FINALIZE);
    k_compact_coord_entries<<<(unsigned)((e->coord_cap + 255) / 256), 256, 0, e->stream>>>(e->coords, e->coord_cap, a, have, e->ctr_d);
  }
  CK(cudaGetLastError());
  rsort::Pass passes[8];
  int np = rsort::plan_word(passes, 0, 32);
  np += rsort::plan_word(passes + np, 1, 32);
  uint4* res = nullptr;
  rc = device_sort(e, reinterpret_cast<uint4*>(a), reinterpret_cast<uint4*>(b), have, passes, np, CoordKey(), &res);
  if (rc == CITAM_OK) {
    Timed tm(e, CITAM_K_FINALIZE);
    k_coord_entries_to_records<<<(unsigned)((have + 255) / 256), 256, 0, e->stream>>>(reinterpret_cast<CoordEntry*>(res), have, out_device);
    cudaError_t err = cudaStreamSynchronize(e->stream);
    if (err != cudaSuccess) rc = fail(CITAM_ERR_CUDA, "coordinate export failed: %s", cudaGetErrorString(err));
  }
  cudaFree(a); cudaFree(b);
  return rc;
}

int citam_export_coords(citam_engine* e, citam_coord* out, uint64_t capacity, uint64_t* n) {
  if (!e || !n || (capacity && !out)) return fail(CITAM_ERR_INVALID, "null argument");
  uint64_t have = 0;
  int rc = citam_coord_count(e, &have);
  if (rc) return rc;
  *n = have;
  if (have == 0) return check_overflow(e);
  if (capacity < have) return fail(CITAM_ERR_CAPACITY, "output holds %llu coords, %llu needed", (unsigned long long)capacity, (unsigned long long)have);
  citam_coord* tmp = nullptr;
  CK(cudaMalloc(&tmp, sizeof(citam_coord) * have));
  rc = citam_export_coords_device(e, tmp, have, n);
  if (rc == CITAM_OK) {
    cudaError_t err = cudaMemcpyAsync(out, tmp, sizeof(citam_coord) * have, cudaMemcpyDeviceToHost, e->stream);
    if (err == cudaSuccess) err = cudaStreamSynchronize(e->stream);
    if (err != cudaSuccess) rc = fail(CITAM_ERR_CUDA, "coordinate export failed: %s", cudaGetErrorString(err));
  }
  cudaFree(tmp);
  return rc;
}

int citam_export_agent_durations(citam_engine* e, uint64_t* out) {
  if (!e || !out) return fail(CITAM_ERR_INVALID, "null argument");
  CK(cudaSetDevice(e->device));
  int rc = check_overflow(e);
  if (rc) return rc;
  if (!e->dur_out) CK(cudaMalloc(&e->dur_out, sizeof(unsigned long long) * (size_t)e->N));
  {
    Timed tm(e, CITAM_K_FINALIZE);
    k_unpack_durations<<<(e->N + 255) / 256, 256, 0, e->stream>>>(e->agent_dur, e->N, e->dur_out);
  }
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, e->dur_out, sizeof(uint64_t) * (size_t)e->N, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  return CITAM_OK;
}

// The contact log in the order the reference visits contacts: (step, a1, a2), sorted on the device.
int citam_export_contact_log(citam_engine* e, citam_contact* out, uint64_t capacity, uint64_t* n) {
  if (!e || !n || (capacity && !out)) return fail(CITAM_ERR_INVALID, "null argument");
  int rc = check_overflow(e);
  if (rc) return rc;
  citam_counters c;
  rc = citam_get_counters(e, &c);
  if (rc) return rc;
  const uint64_t have = c.log_records;
  *n = have;
  if (have == 0) return CITAM_OK;
  if (capacity < have) return fail(CITAM_ERR_CAPACITY, "output holds %llu records, %llu needed", (unsigned long long)capacity, (unsigned long long)have);
  uint4 *a = nullptr, *b = nullptr;
  citam_contact* packed = nullptr;
  CK(cudaMalloc(&a, sizeof(uint4) * have));
  CK(cudaMalloc(&b, sizeof(uint4) * have));
  CK(cudaMalloc(&packed, sizeof(citam_contact) * have));
  CK(cudaMemcpyAsync(a, e->log, sizeof(uint4) * have, cudaMemcpyDeviceToDevice, e->stream));
  rsort::Pass passes[12];
  int np = 0;
  const int ab = bits_of((unsigned long long)e->N - 1);
  np += rsort::plan_word(passes + np, 0, ab);
  np += rsort::plan_word(passes + np, 1, ab);
  np += rsort::plan_word(passes + np, 2, std::min(24, bits_of(e->max_step)));
  uint4* res = nullptr;
  rc = device_sort(e, a, b, have, passes, np, LogKey(), &res);
  if (rc == CITAM_OK) {
    Timed tm(e, CITAM_K_FINALIZE);
    k_log_to_contacts<<<(unsigned)((have + 255) / 256), 256, 0, e->stream>>>(res, have, packed);
    cudaError_t err = cudaMemcpyAsync(out, packed, sizeof(citam_contact) * have, cudaMemcpyDeviceToHost, e->stream);
    if (err == cudaSuccess) err = cudaStreamSynchronize(e->stream);
    if (err != cudaSuccess) rc = fail(CITAM_ERR_CUDA, "contact log export failed: %s", cudaGetErrorString(err));
  }
  cudaFree(a); cudaFree(b); cudaFree(packed);
  return rc;
}

int citam_clear_contact_log(citam_engine* e) {
  if (!e) return fail(CITAM_ERR_INVALID, "null engine");
  CK(cudaSetDevice(e->device));
  CK(cudaMemsetAsync(&e->ctr_d->log_count, 0, sizeof(unsigned long long), e->stream));
  return CITAM_OK;
}

// Statistics in two halves, so that ranks holding disjoint parts of one run's pair table can
// all-reduce the per-agent arrays in between (citam_stats_arrays_device).
int citam_statistics_begin(citam_engine* e) {
  if (!e) return fail(CITAM_ERR_INVALID, "null engine");
  int rc = check_overflow(e);
  if (rc) return rc;
  size_t n = (size_t)e->N;
  CK(cudaMemsetAsync(e->agent_nev, 0, 4 * n, e->stream));
  CK(cudaMemsetAsync(e->agent_has, 0, 4 * n, e->stream));
  CK(cudaMemsetAsync(e->ctr_d->stats, 0, sizeof(unsigned long long) * 6, e->stream));
  {
    Timed tm(e, CITAM_K_FINALIZE);
    k_stats_pairs<<<(unsigned)((e->pair_cap + 255) / 256), 256, 0, e->stream>>>(e->pairs, e->pair_cap, e->agent_nev, e->agent_has, e->ctr_d);
  }
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(e->stream));
  return CITAM_OK;
}

int citam_stats_arrays_device(citam_engine* e, void** events_per_agent, void** has_contact, uint64_t* n_elements) {
  if (!e || !events_per_agent || !has_contact || !n_elements) return fail(CITAM_ERR_INVALID, "null argument");
  *events_per_agent = e->agent_nev;
  *has_contact = e->agent_has;
  *n_elements = (uint64_t)e->N;
  return CITAM_OK;
}

int citam_statistics_end(citam_engine* e, citam_stats* out) {
  if (!e || !out) return fail(CITAM_ERR_INVALID, "null argument");
  CK(cudaSetDevice(e->device));
  {
    Timed tm(e, CITAM_K_FINALIZE);
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

// Statistics from the per-agent counter words and the table's pair count alone: O(n_agents).  Valid
// when the words describe the pairs of the run -- this engine's own (nothing merged or cleared), or
// the whole run's after the all-reduce of citam_agent_durations_device (n_pairs is then this engine's
// share: add it over the ranks).  CITAM_ERR_CAPACITY when some agent has >= 2^24 contact-steps (its
// event field may have wrapped): use the table pass (citam_statistics_begin/_end) then.
int citam_statistics_counters(citam_engine* e, citam_stats* out) {
  if (!e || !out) return fail(CITAM_ERR_INVALID, "null argument");
  int rc = check_overflow(e);
  if (rc) return rc;
  CK(cudaMemsetAsync(e->ctr_d->stats, 0, sizeof(unsigned long long) * 6, e->stream));
  {
    Timed tm(e, CITAM_K_FINALIZE);
    k_stats_counters<<<(e->N + 255) / 256, 256, 0, e->stream>>>(e->agent_dur, e->agent_nev, e->agent_has, e->N, e->ctr_d);
  }
  CK(cudaGetLastError());
  DevCounters c;
  CK(cudaMemcpyAsync(&c, e->ctr_d, sizeof c, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  if (c.stats[5] >= (1ULL << (64 - kAgentDurBits)))
    return fail(CITAM_ERR_CAPACITY, "an agent has %llu contact-steps: its event count needs the pass over the pair table", c.stats[5]);
  out->total_duration = c.stats[0] / 2;
  out->n_pairs = pairs_used(c);
  out->n_agents_with_contacts = c.stats[2];
  out->sum_events_per_agent = c.stats[3];
  out->max_events_per_agent = c.stats[4];
  out->reserved = 0;
  return CITAM_OK;
}

int citam_statistics(citam_engine* e, citam_stats* out) {
  if (!e || !out) return fail(CITAM_ERR_INVALID, "null argument");
  if (e->counters_match && !e->no_fast_stats) {
    int rc = citam_statistics_counters(e, out);
    if (rc != CITAM_ERR_CAPACITY) return rc;
    rc = check_overflow(e);  // a table overflow is final; a wide agent counter is not
    if (rc) return rc;
  }
  int rc = citam_statistics_begin(e);
  if (rc) return rc;
  return citam_statistics_end(e, out);
}

int citam_measure_random_updates(int32_t device, uint64_t table_bytes, uint64_t n_updates, int32_t probe_first, double* ms) {
  if (!ms || table_bytes < 16 || n_updates == 0) return fail(CITAM_ERR_INVALID, "bad argument");
  CK(cudaSetDevice(device));
  unsigned long long words = 2;
  while (words * 2 * 8 <= table_bytes) words *= 2;
  unsigned long long* tab = nullptr;
  CK(cudaMalloc(&tab, words * 8));
  CK(cudaMemset(tab, 0, words * 8));
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  k_random_updates<<<148 * 16, 256>>>(tab, words - 1, std::min<uint64_t>(n_updates, 1 << 20), probe_first);  // warm-up
  CK(cudaEventRecord(a));
  k_random_updates<<<148 * 16, 256>>>(tab, words - 1, n_updates, probe_first);
  CK(cudaEventRecord(b));
  CK(cudaEventSynchronize(b));
  float t = 0;
  CK(cudaEventElapsedTime(&t, a, b));
  *ms = t;
  cudaEventDestroy(a); cudaEventDestroy(b);
  cudaFree(tab);
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
  e->counters_match = false;  // the caller is about to reduce other ranks' words into the array
  *ptr = e->agent_dur;
  *n_elements = (uint64_t)e->N;
  return CITAM_OK;
}

int citam_hist_device(citam_engine* e, void** ptr, uint64_t* n_elements) {
  if (!e || !ptr || !n_elements) return fail(CITAM_ERR_INVALID, "null argument");
  CK(cudaSetDevice(e->device));
  int rc = ensure_ready(e, 1);
  if (rc) return rc;
  CK(cudaStreamSynchronize(e->stream));
  *ptr = e->hist;
  *n_elements = (uint64_t)e->hist_total;
  return CITAM_OK;
}

// The table's records grouped by destination, hash(a1, a2) mod n_parts, into a DEVICE buffer:
// part k occupies out[sum(counts[0..k)) .. + counts[k]).  The send side of the all-to-all.
int citam_partition_pairs_device(citam_engine* e, citam_pair* out_device, uint64_t capacity, int32_t n_parts, uint64_t* counts) {
  if (!e || !counts || (capacity && !out_device)) return fail(CITAM_ERR_INVALID, "null argument");
  if (n_parts < 1 || n_parts > kMaxParts) return fail(CITAM_ERR_INVALID, "n_parts must be in 1..%d", kMaxParts);
  CK(cudaSetDevice(e->device));
  int rc = check_overflow(e);
  if (rc) return rc;
  // per-destination counters / cursors: kept with the engine (a cudaMalloc + cudaFree pair per call cost more
  // than both passes over the table)
  if (!e->part_cnt_d) CK(cudaMalloc(&e->part_cnt_d, sizeof(unsigned long long) * kMaxParts));
  unsigned long long* cnt_d = e->part_cnt_d;
  CK(cudaMemsetAsync(cnt_d, 0, sizeof(unsigned long long) * kMaxParts, e->stream));
  const unsigned blocks = (unsigned)std::min<unsigned long long>(148 * 8, (e->pair_cap + kPartTile - 1) / kPartTile);
  {
    Timed tm(e, CITAM_K_FINALIZE);
    k_partition_count<<<blocks, 256, 0, e->stream>>>(e->pairs, e->pair_cap, (uint32_t)n_parts, cnt_d);
  }
  unsigned long long h[kMaxParts], cur[kMaxParts];
  CK(cudaMemcpyAsync(h, cnt_d, sizeof h, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  unsigned long long tot = 0;
  for (int k = 0; k < n_parts; ++k) { counts[k] = h[k]; cur[k] = tot; tot += h[k]; }
  for (int k = n_parts; k < kMaxParts; ++k) cur[k] = tot;
  if (tot > capacity) {
    return fail(CITAM_ERR_CAPACITY, "output holds %llu pairs, %llu needed", (unsigned long long)capacity, tot);
  }
  if (tot) {
    CK(cudaMemcpyAsync(cnt_d, cur, sizeof cur, cudaMemcpyHostToDevice, e->stream));
    Timed tm(e, CITAM_K_FINALIZE);
    k_partition_scatter<citam_pair><<<blocks, 256, 0, e->stream>>>(e->pairs, e->pair_cap, (uint32_t)n_parts, cnt_d, out_device, capacity);
  }
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(e->stream));
  return CITAM_OK;
}

// Merge pair records exported by another engine (DEVICE buffer): steps and events add, the first
// contact step is the minimum.  Only the pair table is touched (see the header).
int citam_merge_pairs_device(citam_engine* e, const citam_pair* pairs_device, uint64_t n) {
  if (!e || (n && !pairs_device)) return fail(CITAM_ERR_INVALID, "null argument");
  if (!n) return CITAM_OK;
  CK(cudaSetDevice(e->device));
  int rc = join_side_stream(e);
  if (rc) return rc;
  if (e->pair_auto) {  // room for every record before the first one is applied
    citam_counters c;
    rc = citam_get_counters(e, &c);
    if (rc) return rc;
    unsigned long long ncap = e->pair_cap;
    while ((c.pair_slots_used + n) * 2 > ncap) ncap *= 2;
    if (ncap != e->pair_cap) {
      rc = rehash_pairs(e, ncap);
      if (rc == CITAM_ERR_CAPACITY) return fail(CITAM_ERR_CAPACITY, "out of device memory growing the pair table to %llu slots", ncap);
      if (rc) return rc;
    }
  }
  CK(cudaMemsetAsync(e->flag_d, 0, sizeof(int), e->stream));
  Params p = make_params(e);
  {
    Timed tm(e, CITAM_K_FINALIZE);
    e->counters_match = false;
    const unsigned long long per_block = 256ULL * kMergeIlp;
    k_merge_pairs<<<(unsigned)((n + per_block - 1) / per_block), 256, 0, e->stream>>>(p, pairs_device, n, e->flag_d);
  }
  CK(cudaGetLastError());
  int bad = 0;
  rc = read_flag(e, &bad);
  if (rc) return rc;
  if (bad) return fail(CITAM_ERR_INVALID, "a merged pair record is malformed (need a1 < a2 < n_agents, duration < 2^24, events < 2^16)");
  // merged first-contact steps are not bounded by what this engine ran: the kernel kept their maximum
  DevCounters c;
  CK(cudaMemcpyAsync(&c, e->ctr_d, sizeof c, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  e->max_step = std::max<uint32_t>(e->max_step, (uint32_t)std::min<unsigned long long>(0xFFFFFEu, c.merged_max_step) + 1u);
  return check_overflow(e);
}

int citam_merge_pairs(citam_engine* e, const citam_pair* pairs, uint64_t n) {
  if (!e || (n && !pairs)) return fail(CITAM_ERR_INVALID, "null argument");
  if (!n) return CITAM_OK;
  CK(cudaSetDevice(e->device));
  citam_pair* tmp = nullptr;
  CK(cudaMalloc(&tmp, sizeof(citam_pair) * n));
  cudaError_t err = cudaMemcpyAsync(tmp, pairs, sizeof(citam_pair) * n, cudaMemcpyHostToDevice, e->stream);
  int rc = err == cudaSuccess ? citam_merge_pairs_device(e, tmp, n) : fail(CITAM_ERR_CUDA, "copy failed: %s", cudaGetErrorString(err));
  cudaStreamSynchronize(e->stream);
  cudaFree(tmp);
  return rc;
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
    if (c.sx < 2 * fd.minx || c.sx - 2 * fd.minx >= std::max(1, 2 * fd.W - 1) || c.sy < 2 * fd.miny || c.sy - 2 * fd.miny >= std::max(1, 2 * fd.H - 1))
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
