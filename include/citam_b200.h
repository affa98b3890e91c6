/*
 * citam_b200.h — C ABI of the B200-native CITAM step loop.
 *
 * The upstream reference (corning-incorporated/citam) is pure Python and has no
 * FFI: its hot path is reached through the `Calculator` plugin ABC
 * (citam/engine/calculators/calculator.py:9-33) and `Simulation.run_serial`
 * (citam/engine/core/simulation.py:215-275).  This header is the boundary a
 * maintainer would bind from that Python code with ctypes;
 * every entry point names the reference code it replaces.
 *
 * Conventions: plain C, opaque handle, every call returns 0 on success or a
 * negative citam_status; `citam_last_error()` gives the message of the last
 * failing call on this thread.  All pointers are caller-owned HOST buffers
 * unless a name ends in `_device`.  One engine is bound to one GPU; calls on
 * one engine are not thread-safe.  Work is enqueued on the engine's stream
 * (`citam_set_stream`, default: the legacy default stream) and calls that
 * return data synchronise that stream.
 *
 * There is no CPU fallback: `citam_create` fails when no sm_100 device is
 * usable.
 *
 * Limits (checked; violations return CITAM_ERR_INVALID or CITAM_ERR_CAPACITY):
 *   floors <= 63, spaces per floor <= 32767, location-table sides <= 65536,
 *   steps < 2^24 - 1, |coordinates| < 2^26, per-step stride of a segment within
 *   int16, contact-steps per pair < 2^24 and events per pair < 2^16,
 *   (n_agents - 1) x steps run since the last reset < 2^40 (the most contact-
 *   steps one agent can collect), steps_per_chunk <= 1024, fewer than 2^32
 *   records per export.  Itinerary
 *   coordinates are integers (the reference's navigation lattice); contact
 *   positions are their midpoints, reported as (x1+x2, y1+y2).
 *   citam_load_itineraries checks the segment store on the device: seg_off
 *   non-decreasing inside [0, n_segs], t0 strictly increasing per agent, floors
 *   0..62, a zero-velocity last segment per agent.
 */
#ifndef CITAM_B200_H
#define CITAM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CITAM_ABI_VERSION 3

typedef enum {
  CITAM_OK = 0,
  CITAM_ERR_INVALID = -1,   /* bad argument                                  */
  CITAM_ERR_CUDA = -2,      /* CUDA runtime error (message has the detail)   */
  CITAM_ERR_CAPACITY = -3,  /* a device table overflowed; enlarge and re-run */
  CITAM_ERR_STATE = -4,     /* call order (e.g. run before load)             */
  CITAM_ERR_NO_DEVICE = -5  /* no usable CUDA device                         */
} citam_status;

typedef struct citam_engine citam_engine;

/* ---- configuration ---------------------------------------------------- */
typedef struct {
  int32_t n_agents;         /* agents 0..n-1 in schedule order (simulation.py:277-289) */
  int32_t n_floors;         /* 1..63 */
  double contact_distance;  /* CloseContactsCalculator.contact_distance, strict `<`
                               (close_contacts_calculator.py:21-28,171)                */
  int32_t cell_size;        /* uniform-grid cell edge in drawing units; 0 = ceil(D)    */
  int32_t device;           /* CUDA device ordinal                                     */
  uint64_t pair_capacity;   /* slots of the per-pair table (rounded up to 2^k); 0=auto */
  uint64_t coord_capacity;  /* slots of the per-coordinate histogram; 0 = auto         */
  uint64_t log_capacity;    /* contact-log records kept for the full-fidelity writers
                               (contacts.txt, raw_contact_data.ccd); 0 = log off       */
  int32_t steps_per_chunk;  /* time steps processed per launch group; 0 = auto         */
  int32_t flags;            /* CITAM_FLAG_*                                            */
} citam_config;

#define CITAM_FLAG_TIMING 1        /* record CUDA events around every kernel (citam_get_timings) */
#define CITAM_FLAG_HASHED_COORDS 2 /* keep the per-coordinate histogram in a hash table even when the
                                      floors' lattices are small enough for the dense one          */

/* One constant-velocity run of an itinerary.  Agent a's position for
 * t in [t0, next.t0) is (x0 + dx*(t-t0), y0 + dy*(t-t0)) on `floor`; before the
 * first segment the agent is not in the facility; the last segment lasts for
 * ever (the reference's sticky tail: agent.py:77-107 with
 * office_schedule.py:632-638).  Produced by citam_b200.itinerary.encode(). */
typedef struct {
  int32_t t0;
  int32_t x0, y0;
  int16_t dx, dy;
  int16_t floor;
  int16_t reserved;
} citam_segment;

/* Per-agent state at one step; what Agent.pos / current_floor /
 * current_location hold (agent.py:63-65).  floor < 0: not in the facility
 * (pos None); loc < 0: inside no space (current_location None). */
typedef struct {
  int32_t x, y;
  int16_t floor;
  int16_t loc;
} citam_agent_state;

/* One space of a floor for the location-table builder (Space, space.py:25-117). */
typedef struct {
  int32_t bb_xmin, bb_xmax, bb_ymin, bb_ymax; /* round() of space.path.bbox() (space.py:209-216) */
  int32_t seg_begin, seg_end;                 /* range in the floor's boundary-segment array      */
  int32_t door_begin, door_end;               /* range in the floor's door array; empty when the
                                                 door shortcut does not apply (space.py:218-224)  */
} citam_space;

/* One boundary segment: endpoints, the unit normal the reference compares
 * against (svgpathtools Line.normal, geometry.py:580) and whether
 * length() > 1.0 (space.py:189). */
typedef struct {
  double sx, sy, ex, ey;
  double nx, ny;
  int32_t long_enough;
  int32_t reserved;
} citam_boundary;

typedef struct {
  double sx, sy, ex, ey;
} citam_door;

typedef struct {
  uint32_t a1, a2;       /* a1 < a2                                               */
  uint32_t first_step;   /* step of the first contact: row order of pair_contact.csv */
  uint32_t n_events;     /* len(contact_data[key])            (contacts.py:170)   */
  uint64_t duration;     /* sum of event durations            (contacts.py:171-172) */
} citam_pair;

typedef struct {
  int32_t floor;
  int32_t sx, sy;        /* x1+x2, y1+y2: twice the contact position (close_contacts_calculator.py:218-220) */
  int32_t reserved;
  uint64_t count;
} citam_coord;

typedef struct {
  uint32_t step, a1, a2; /* one contact-step; a1 < a2 */
} citam_contact;

/* Integer sums behind ContactEvents.extract_statistics (contacts.py:185-294);
 * the host does the final float division/rounding exactly as Python does. */
typedef struct {
  uint64_t total_duration;      /* sum over pairs of duration                          */
  uint64_t n_pairs;             /* == sum(n_others.values())                           */
  uint64_t n_agents_with_contacts;
  uint64_t sum_events_per_agent;/* sum over agents of their n_events (== 2*sum n_events) */
  uint64_t max_events_per_agent;
  uint64_t reserved;
} citam_stats;

typedef struct {
  uint64_t agent_steps;   /* agent-steps advanced (N * steps)                     */
  uint64_t pair_tests;    /* distance evaluations actually executed               */
  uint64_t contact_steps; /* contacts found                                       */
  uint64_t kernel_launches;
  uint64_t pair_slots_used, coord_slots_used, log_records;
  uint64_t overflow;      /* bit0 pair table, bit1 coord table, bit2 log          */
  uint64_t contact_runs;  /* table updates: a run of contact-steps of a stationary pair is one,
                             however many chunks of the citam_run range it crosses        */
  uint64_t binned_agent_steps; /* (agent, step) records binned and sorted per step: every located
                             agent-step with the contact log on; without it only the steps at which
                             an agent is not covered by an entry of the per-chunk static array */
} citam_counters;

#define CITAM_N_KERNELS 9
typedef struct {
  double ms[CITAM_N_KERNELS];      /* summed CUDA-event time per kernel kind */
  uint64_t launches[CITAM_N_KERNELS];
} citam_timings;
/* kernel kinds (index into citam_timings) */
#define CITAM_K_EXPAND_BIN 0
#define CITAM_K_SCAN 1
#define CITAM_K_SCATTER 2
#define CITAM_K_PAIRS 3
#define CITAM_K_LUT 4
#define CITAM_K_FINALIZE 5
#define CITAM_K_MISC 6
#define CITAM_K_ACCUMULATE 7
#define CITAM_K_CLASSIFY 8 /* per-chunk classification of the agents (split pipeline) */

/* ---- life cycle -------------------------------------------------------- */
int citam_abi_version(void);
const char* citam_last_error(void);
int citam_create(const citam_config* cfg, citam_engine** out);
void citam_destroy(citam_engine* e);
int citam_set_stream(citam_engine* e, void* cuda_stream);

/* ---- facility: replaces Floorplan.identify_this_location + hallway test -- */
/* Build floor `f`'s (x,y)->space table on the device with the reference's rule
 * (floorplan.py:215-241, space.py:194-280, geometry.py:31-127,558-588) over the
 * lattice [minx, minx+W) x [miny, miny+H). */
int citam_build_floor_lut(citam_engine* e, int32_t floor, int32_t minx, int32_t miny,
                          int32_t W, int32_t H, int32_t n_spaces, const citam_space* spaces,
                          int32_t n_boundaries, const citam_boundary* boundaries,
                          int32_t n_doors, const citam_door* doors);
/* Or install a table computed elsewhere (int16 [H][W], -1 = no space). */
int citam_set_floor_lut(citam_engine* e, int32_t floor, int32_t minx, int32_t miny,
                        int32_t W, int32_t H, int32_t n_spaces, const int16_t* lut);
int citam_get_floor_lut(citam_engine* e, int32_t floor, int16_t* out /* [H*W] */);
/* adj[a*n_spaces+b] != 0  <=>  both spaces are hallways and the floor's hallway
 * graph has the edge (close_contacts_calculator.py:173-203); NULL = no graph. */
int citam_set_floor_adjacency(citam_engine* e, int32_t floor, int32_t n_spaces, const uint8_t* adj);

/* ---- itineraries: replaces Schedule.itinerary / get_next_position ------- */
/* seg_off is a host array; segs may be a host or a device pointer (unified addressing). */
int citam_load_itineraries(citam_engine* e, const int64_t* seg_off /* [n_agents+1] */,
                           const citam_segment* segs, int64_t n_segs);

/* ---- the step loop: replaces Simulation.step x (t_end - t_begin) --------- */
/* Advance every agent through steps [t_begin, t_end) and accumulate contacts
 * (simulation.py:291-361 + close_contacts_calculator.py:64-137).  Accumulators
 * are commutative, so ranges may be run in any order / on different engines. */
int citam_run(citam_engine* e, int32_t t_begin, int32_t t_end);
/* One step on caller-supplied states (the Calculator.run(current_step, agents)
 * entry: calculator.py:24-25).  `states` has n_agents records. */
int citam_step_states(citam_engine* e, int32_t step, const citam_agent_state* states);
/* Agent states at one step from the loaded itineraries (trajectory writer,
 * simulation.py:312-334). */
int citam_states_at(citam_engine* e, int32_t step, citam_agent_state* out /* [n_agents] */);
/* Same for every step of [t_begin, t_end): out[(t - t_begin) * n_agents + a]. */
int citam_states_range(citam_engine* e, int32_t t_begin, int32_t t_end, citam_agent_state* out);
int citam_reset(citam_engine* e); /* clear all accumulators, keep facility + itineraries */

/* ---- results: replaces ContactEvents / extract_* ------------------------ */
/* Rows come in the reference's own order, produced by a device radix sort (K5):
 * pairs by (first contact step, a1, a2) = dict insertion order of
 * ContactEvents.contact_data (contacts.py:135-183), coordinates by (floor, sx, sy),
 * the contact log by (step, a1, a2) = the visiting order of
 * close_contacts_calculator.py:104-124. */
int citam_pair_count(citam_engine* e, uint64_t* n);
int citam_export_pairs(citam_engine* e, citam_pair* out, uint64_t capacity, uint64_t* n);
int citam_export_agent_durations(citam_engine* e, uint64_t* out /* [n_agents] */);
int citam_coord_count(citam_engine* e, uint64_t* n);
int citam_export_coords(citam_engine* e, citam_coord* out, uint64_t capacity, uint64_t* n);
int citam_export_contact_log(citam_engine* e, citam_contact* out, uint64_t capacity, uint64_t* n);
int citam_clear_contact_log(citam_engine* e); /* drop exported records (block-wise full-fidelity writers) */
int citam_statistics(citam_engine* e, citam_stats* out);
int citam_get_counters(citam_engine* e, citam_counters* out);
int citam_get_timings(citam_engine* e, citam_timings* out);
/* Measurement aid (no engine): time n_updates 8-byte reductions at hashed addresses of a zeroed
 * table of table_bytes (rounded down to a power of two) on `device`; probe_first != 0 adds the
 * 16-byte slot load that precedes a pair-table update.  The rate is the memory system's ceiling for
 * the accumulation kernel's access pattern (bench.py prints it beside the roofline line). */
int citam_measure_random_updates(int32_t device, uint64_t table_bytes, uint64_t n_updates,
                                 int32_t probe_first, double* ms);

/* ---- multi-GPU: one scenario sharded by time ranges over G engines -------- */
/* (the "timestep blocks" parallel runner the reference's docstring asks for,
 *  simulation.py:222-227).  Every accumulator is commutative, so the ranks meet
 *  once, at the end; the collectives themselves (NCCL through torch.distributed)
 *  run on the DEVICE buffers these calls expose or fill:
 *    per-agent contact seconds   all-reduce(sum) on citam_agent_durations_device
 *    dense coordinate histogram  all-reduce(sum) on citam_hist_device
 *    pair table                  citam_partition_pairs_device -> all-to-all by
 *                                hash(a1,a2) mod G -> citam_clear_pairs +
 *                                citam_merge_pairs_device: every rank ends with
 *                                the complete records of its share of the pairs
 *    statistics                  citam_statistics_begin, all-reduce(sum) on the two
 *                                citam_stats_arrays_device arrays, citam_statistics_end
 *                                (total_duration and n_pairs are then per-share sums:
 *                                add them over the ranks)
 * Device views are the live arrays: uint64 [n_agents] per-agent counter words
 * (contact seconds in bits 0-39, contact events in bits 40-63; sums of words are
 * sums of both fields; citam_export_agent_durations returns the seconds), uint64
 * [n] histogram buckets (NULL/0 when the hashed histogram is in use), uint32
 * [n_agents] events per agent and has-contact flags. */
int citam_agent_durations_device(citam_engine* e, void** ptr, uint64_t* n_elements);
int citam_hist_device(citam_engine* e, void** ptr, uint64_t* n_elements);
int citam_export_pairs_device(citam_engine* e, citam_pair* out_device, uint64_t capacity, uint64_t* n);
int citam_export_coords_device(citam_engine* e, citam_coord* out_device, uint64_t capacity, uint64_t* n);
/* All pairs of the table grouped by destination = hash(a1,a2) mod n_parts (n_parts <= 64):
 * part k is out_device[sum(counts[0..k)) ... + counts[k]); counts is a HOST array. */
int citam_partition_pairs_device(citam_engine* e, citam_pair* out_device, uint64_t capacity,
                                 int32_t n_parts, uint64_t* counts);
int citam_clear_pairs(citam_engine* e); /* empty the pair table only */
/* Merge pair / coordinate records exported by another engine (other rank):
 * steps and events add, first contact step = min.  Only the pair table /
 * histogram are touched -- per-agent contact seconds are NOT updated (reduce
 * them with the all-reduce above).  A library-sized pair table is grown before
 * the first record is applied. */
int citam_merge_pairs_device(citam_engine* e, const citam_pair* pairs_device, uint64_t n);
int citam_merge_pairs(citam_engine* e, const citam_pair* pairs, uint64_t n);
int citam_merge_coords(citam_engine* e, const citam_coord* coords, uint64_t n);
/* Statistics from the per-agent counter words and the pair count alone, O(n_agents), no
 * pass over the pair table: valid when the words describe the run's pairs -- this engine's
 * own (citam_statistics takes this path by itself), or the whole run's after the all-reduce
 * of citam_agent_durations_device (n_pairs is then this engine's share: add it over the
 * ranks).  CITAM_ERR_CAPACITY if an agent has >= 2^24 contact-steps (its event count is then
 * not provably exact): use citam_statistics_begin/_end. */
int citam_statistics_counters(citam_engine* e, citam_stats* out);
int citam_statistics_begin(citam_engine* e);
int citam_stats_arrays_device(citam_engine* e, void** events_per_agent, void** has_contact,
                              uint64_t* n_elements);
int citam_statistics_end(citam_engine* e, citam_stats* out);

#ifdef __cplusplus
}
#endif
#endif /* CITAM_B200_H */
