/*
 * grid_oracle.c — plain-C CPU restatement of CITAM's per-timestep loop, sized for
 * inputs the reference's O(N^2) formulation cannot reach.
 *
 * TEST INFRASTRUCTURE ONLY.  Built by oracle/Makefile into oracle/_build/; only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference`
 * legs may load it.  The product library (citam_b200/csrc) never links or calls it.
 *
 * Restates (paths relative to the upstream reference):
 *   Agent.step + get_next_position   citam/engine/core/agent.py:67-109,
 *                                    citam/engine/schedulers/office_schedule.py:632-638
 *   Floorplan.identify_this_location citam/engine/map/floorplan.py:215-241,
 *                                    citam/engine/map/space.py:194-280,
 *                                    citam/engine/map/geometry.py:31-127,558-588
 *   CloseContactsCalculator.run      citam/engine/calculators/close_contacts_calculator.py:64-231
 *   ContactEvents.add_contact        citam/engine/calculators/contacts.py:92-142
 *   end-of-run tables                citam/engine/calculators/contacts.py:155-346
 *
 * The reference finds close pairs with an N x N scipy distance matrix
 * (close_contacts_calculator.py:152-171); here candidates come from a uniform grid, the
 * predicate on each candidate is the reference's own: float64 sqrt(dx^2+dy^2) < D
 * (strict), same floor, same space or adjacent hallways.  The event rule is the
 * reference's sequential one (extend the last event iff last.start+last.duration == step),
 * which is a different formulation from the CUDA path's order-independent rule — that is
 * what makes this a useful cross-check.  Parity pinning: tests/test_oracle_golden.py checks
 * this file against the faithful Python port (oracle/port.py) and against output files of
 * the real reference on every golden case.
 *
 * Threads (go_run with threads > 1, for the CPU baseline): time blocks are independent
 * once the block's first step knows which pairs were in contact one step earlier; each
 * thread runs a block with private tables and the tables are merged (sum, sum, min).
 * Compile with -ffp-contract=off.
 */
#include "../include/citam_b200.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
  int32_t minx, miny, W, H, n_spaces;
  const int16_t* lut; /* [H][W], -1 = none */
  const uint8_t* adj; /* [n_spaces][n_spaces] or NULL */
} go_floor;

/* ---- open-addressing tables ------------------------------------------- */
typedef struct {
  uint64_t key; /* (a1<<32)|a2, a1<a2; UINT64_MAX = empty */
  uint64_t dur;
  uint32_t nev, first, last; /* last = last contact step */
  uint32_t pad;
} pair_ent;

typedef struct {
  pair_ent* e;
  uint64_t cap, used;
} pair_tab;

typedef struct {
  uint64_t key; /* floor<<56 | (sx+2^27)<<28 | (sy+2^27); UINT64_MAX = empty */
  uint64_t count;
} coord_ent;

typedef struct {
  coord_ent* e;
  uint64_t cap, used;
} coord_tab;

#define EMPTY UINT64_MAX

static uint64_t mix(uint64_t k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
  return k;
}

static void pt_init(pair_tab* t, uint64_t cap) {
  t->cap = cap; t->used = 0;
  t->e = (pair_ent*)malloc(sizeof(pair_ent) * cap);
  for (uint64_t i = 0; i < cap; ++i) t->e[i].key = EMPTY;
}
static pair_ent* pt_slot(pair_tab* t, uint64_t key);
static void pt_grow(pair_tab* t) {
  pair_tab n; pt_init(&n, t->cap * 2);
  for (uint64_t i = 0; i < t->cap; ++i)
    if (t->e[i].key != EMPTY) { pair_ent* s = pt_slot(&n, t->e[i].key); *s = t->e[i]; }
  n.used = t->used;
  free(t->e); *t = n;
}
static pair_ent* pt_slot(pair_tab* t, uint64_t key) {
  uint64_t m = t->cap - 1, h = mix(key) & m;
  while (t->e[h].key != EMPTY && t->e[h].key != key) h = (h + 1) & m;
  return &t->e[h];
}
static pair_ent* pt_get(pair_tab* t, uint64_t key) { /* find or insert (zeroed) */
  if ((t->used + 1) * 10 > t->cap * 7) pt_grow(t);
  pair_ent* s = pt_slot(t, key);
  if (s->key == EMPTY) { s->key = key; s->dur = 0; s->nev = 0; s->first = UINT32_MAX; s->last = UINT32_MAX; s->pad = 0; t->used++; }
  return s;
}

static void ct_init(coord_tab* t, uint64_t cap) {
  t->cap = cap; t->used = 0;
  t->e = (coord_ent*)malloc(sizeof(coord_ent) * cap);
  for (uint64_t i = 0; i < cap; ++i) t->e[i].key = EMPTY;
}
static coord_ent* ct_slot(coord_tab* t, uint64_t key) {
  uint64_t m = t->cap - 1, h = mix(key) & m;
  while (t->e[h].key != EMPTY && t->e[h].key != key) h = (h + 1) & m;
  return &t->e[h];
}
static void ct_add(coord_tab* t, uint64_t key, uint64_t n) {
  if ((t->used + 1) * 10 > t->cap * 7) {
    coord_tab g; ct_init(&g, t->cap * 2);
    for (uint64_t i = 0; i < t->cap; ++i)
      if (t->e[i].key != EMPTY) { coord_ent* s = ct_slot(&g, t->e[i].key); *s = t->e[i]; }
    g.used = t->used; free(t->e); *t = g;
  }
  coord_ent* s = ct_slot(t, key);
  if (s->key == EMPTY) { s->key = key; s->count = 0; t->used++; }
  s->count += n;
}
static uint64_t coord_key(int floor, int sx, int sy) {
  return ((uint64_t)(floor & 0x7f) << 56) | ((uint64_t)((uint32_t)(sx + (1 << 27)) & 0xfffffffu) << 28) |
         (uint64_t)((uint32_t)(sy + (1 << 27)) & 0xfffffffu);
}

/* ---- the engine --------------------------------------------------------- */
typedef struct {
  int32_t N, F, cell;
  double D;
  go_floor* floors;
  int32_t *ncx, *ncy, *cell_base, total_cells;
  const int64_t* seg_off;
  const citam_segment* segs;
  pair_tab pairs;
  coord_tab coords;
  uint64_t* agent_dur;
  uint64_t pair_tests, contact_steps;
  struct go_work_s* ws; /* per-thread private accumulators, merged lazily (go_finalize) */
  int n_ws;
} go_engine;

typedef struct go_work_s { /* per-thread scratch + private accumulators */
  int32_t *x, *y, *fl, *loc; /* state of every agent at the current step */
  int64_t* cur;              /* segment cursor per agent (-1 before the first) */
  int32_t *cell_start, *cell_fill, *order, *key;
  pair_tab pairs;
  coord_tab coords;
  uint64_t* agent_dur;
  uint64_t pair_tests, contact_steps;
} go_work;

static int lut_at(const go_floor* f, int x, int y) {
  if (!f->lut) return -1;
  int fx = x - f->minx, fy = y - f->miny;
  if (fx < 0 || fx >= f->W || fy < 0 || fy >= f->H) return -1;
  return f->lut[(size_t)fy * f->W + fx];
}

void* go_create(int32_t N, int32_t F, double D, int32_t cell_size, const go_floor* floors, const int64_t* seg_off,
                const citam_segment* segs) {
  go_engine* e = (go_engine*)calloc(1, sizeof *e);
  e->N = N; e->F = F; e->D = D;
  int c = (int)ceil(D); if (c < 1) c = 1;
  e->cell = cell_size > c ? cell_size : c;
  e->floors = (go_floor*)malloc(sizeof(go_floor) * F);
  memcpy(e->floors, floors, sizeof(go_floor) * F);
  e->ncx = (int32_t*)malloc(4 * F); e->ncy = (int32_t*)malloc(4 * F); e->cell_base = (int32_t*)malloc(4 * F);
  int base = 0;
  for (int f = 0; f < F; ++f) {
    e->ncx[f] = floors[f].lut ? (floors[f].W + e->cell - 1) / e->cell : 0;
    e->ncy[f] = floors[f].lut ? (floors[f].H + e->cell - 1) / e->cell : 0;
    e->cell_base[f] = base; base += e->ncx[f] * e->ncy[f];
  }
  e->total_cells = base;
  e->seg_off = seg_off; e->segs = segs;
  pt_init(&e->pairs, 1024); ct_init(&e->coords, 1024);
  e->agent_dur = (uint64_t*)calloc((size_t)N, 8);
  return e;
}

static void work_free(struct go_work_s* w);
void go_destroy(void* h) {
  go_engine* e = (go_engine*)h;
  if (!e) return;
  for (int b = 0; b < e->n_ws; ++b) work_free(&e->ws[b]);
  free(e->ws);
  free(e->floors); free(e->ncx); free(e->ncy); free(e->cell_base); free(e->pairs.e); free(e->coords.e);
  free(e->agent_dur); free(e);
}

static void work_init(go_engine* e, go_work* w) {
  size_t n = (size_t)e->N;
  w->x = (int32_t*)malloc(4 * n); w->y = (int32_t*)malloc(4 * n); w->fl = (int32_t*)malloc(4 * n);
  w->loc = (int32_t*)malloc(4 * n); w->cur = (int64_t*)malloc(8 * n);
  w->cell_start = (int32_t*)malloc(4 * ((size_t)e->total_cells + 1));
  w->cell_fill = (int32_t*)malloc(4 * ((size_t)e->total_cells + 1));
  w->order = (int32_t*)malloc(4 * n); w->key = (int32_t*)malloc(4 * n);
  pt_init(&w->pairs, 1024); ct_init(&w->coords, 1024);
  w->agent_dur = (uint64_t*)calloc(n, 8);
  w->pair_tests = w->contact_steps = 0;
}
static void work_free(go_work* w) {
  free(w->x); free(w->y); free(w->fl); free(w->loc); free(w->cur); free(w->cell_start); free(w->cell_fill);
  free(w->order); free(w->key); free(w->pairs.e); free(w->coords.e); free(w->agent_dur);
}

/* agent.py:67-109 with office_schedule.py:632-638: entry min(t, len-1); None is sticky.
 * In segment form: the last segment with t0 <= t.  `first` positions the cursors by search. */
static void advance(go_engine* e, go_work* w, int t, int first) {
  for (int a = 0; a < e->N; ++a) {
    int64_t sb = e->seg_off[a], se = e->seg_off[a + 1];
    int64_t c;
    if (first) {
      int64_t lo = sb, hi = se;
      while (lo < hi) { int64_t m = (lo + hi) >> 1; if (e->segs[m].t0 <= t) lo = m + 1; else hi = m; }
      c = lo - 1;
    } else {
      c = w->cur[a];
      if (c < sb) c = sb - 1;
      while (c + 1 < se && e->segs[c + 1].t0 <= t) ++c;
    }
    w->cur[a] = c;
    if (c < sb || t < 0) { w->fl[a] = -1; w->loc[a] = -1; w->x[a] = 0; w->y[a] = 0; continue; }
    const citam_segment* s = &e->segs[c];
    int dt = t - s->t0;
    int x = s->x0 + s->dx * dt, y = s->y0 + s->dy * dt, f = s->floor;
    w->x[a] = x; w->y[a] = y; w->fl[a] = f;
    w->loc[a] = (f >= 0 && f < e->F) ? lut_at(&e->floors[f], x, y) : -1;
  }
}

/* close_contacts_calculator.py:104-124,173-203 on one candidate pair, a1 < a2 */
static int is_contact(go_engine* e, go_work* w, int a1, int a2) {
  double dx = (double)(w->x[a1] - w->x[a2]), dy = (double)(w->y[a1] - w->y[a2]);
  double d = sqrt(dx * dx + dy * dy); /* scipy pdist euclidean, float64 */
  if (!(d < e->D)) return 0;
  if (w->fl[a1] != w->fl[a2]) return 0;
  int l1 = w->loc[a1], l2 = w->loc[a2];
  if (l1 == l2) return 1;
  const go_floor* f = &e->floors[w->fl[a1]];
  if (!f->adj) return 0;
  return f->adj[(size_t)l1 * f->n_spaces + l2] != 0; /* has_edge(name(a1.loc), name(a2.loc)) */
}

/* one step; mark_only = 1 records "in contact at this step" without accumulating (block halo) */
static void step(go_engine* e, go_work* w, int t, int mark_only) {
  int N = e->N, cs = e->cell;
  /* close_contacts_calculator.py:91-99: fewer than two active agents -> nothing */
  int active = 0;
  for (int a = 0; a < N; ++a) active += w->fl[a] >= 0;
  if (active <= 1) return;
  memset(w->cell_start, 0, 4 * ((size_t)e->total_cells + 1));
  int located = 0;
  for (int a = 0; a < N; ++a) {
    if (w->fl[a] < 0 || w->loc[a] < 0) { w->key[a] = -1; continue; }
    int f = w->fl[a];
    int cx = (w->x[a] - e->floors[f].minx) / cs, cy = (w->y[a] - e->floors[f].miny) / cs;
    int k = e->cell_base[f] + cy * e->ncx[f] + cx;
    w->key[a] = k; w->cell_start[k + 1]++; located++;
  }
  if (located == 0) return;
  for (int k = 0; k < e->total_cells; ++k) w->cell_start[k + 1] += w->cell_start[k];
  memcpy(w->cell_fill, w->cell_start, 4 * ((size_t)e->total_cells + 1));
  for (int a = 0; a < N; ++a) if (w->key[a] >= 0) w->order[w->cell_fill[w->key[a]]++] = a; /* id order inside a cell */
  for (int a = 0; a < N; ++a) {
    int k = w->key[a];
    if (k < 0) continue;
    int f = w->fl[a];
    int rel = k - e->cell_base[f], ncx = e->ncx[f], ncy = e->ncy[f];
    int cx = rel % ncx, cy = rel / ncx;
    for (int oy = -1; oy <= 1; ++oy) {
      int yy = cy + oy; if (yy < 0 || yy >= ncy) continue;
      for (int ox = -1; ox <= 1; ++ox) {
        int xx = cx + ox; if (xx < 0 || xx >= ncx) continue;
        int kk = e->cell_base[f] + yy * ncx + xx;
        for (int q = w->cell_start[kk]; q < w->cell_start[kk + 1]; ++q) {
          int b = w->order[q];
          if (b <= a) continue; /* the reference skips i >= j (close_contacts_calculator.py:106) */
          w->pair_tests++;
          if (!is_contact(e, w, a, b)) continue;
          /* add_contact_event (close_contacts_calculator.py:205-231) */
          pair_ent* p = pt_get(&w->pairs, ((uint64_t)(uint32_t)a << 32) | (uint32_t)b);
          if (mark_only) { p->last = (uint32_t)t; continue; }
          w->contact_steps++;
          w->agent_dur[a]++; w->agent_dur[b]++;
          /* contacts.py:131-142: extend the last event iff it ended at t-1, else a new event */
          if (!(p->last != UINT32_MAX && p->last + 1 == (uint32_t)t)) p->nev++;
          if (p->first == UINT32_MAX) p->first = (uint32_t)t;
          p->last = (uint32_t)t; p->dur++;
          /* contact position = midpoint; key by its double (exact for integers) */
          ct_add(&w->coords, coord_key(f, w->x[a] + w->x[b], w->y[a] + w->y[b]), 1);
        }
      }
    }
  }
}

static void run_block(go_engine* e, go_work* w, int t0, int t1) {
  if (t0 > 0) { advance(e, w, t0 - 1, 1); step(e, w, t0 - 1, 1); advance(e, w, t0, 0); }
  else advance(e, w, t0, 1);
  for (int t = t0; t < t1; ++t) {
    if (t > t0) advance(e, w, t, 0);
    step(e, w, t, 0);
  }
}

static void merge(go_engine* e, go_work* w) {
  for (uint64_t i = 0; i < w->pairs.cap; ++i) {
    pair_ent* s = &w->pairs.e[i];
    if (s->key == EMPTY || s->dur == 0) continue; /* halo-only marks carry nothing */
    pair_ent* d = pt_get(&e->pairs, s->key);
    d->dur += s->dur; d->nev += s->nev;
    if (s->first < d->first) d->first = s->first;
  }
  for (uint64_t i = 0; i < w->coords.cap; ++i)
    if (w->coords.e[i].key != EMPTY) ct_add(&e->coords, w->coords.e[i].key, w->coords.e[i].count);
  for (int a = 0; a < e->N; ++a) e->agent_dur[a] += w->agent_dur[a];
}

/* fold every thread's private tables into the engine's (called by the exporters) */
static void go_finalize(go_engine* e) {
  for (int b = 0; b < e->n_ws; ++b) { merge(e, &e->ws[b]); work_free(&e->ws[b]); }
  free(e->ws); e->ws = NULL; e->n_ws = 0;
}

int go_run(void* h, int32_t t_begin, int32_t t_end, int32_t threads) {
  go_engine* e = (go_engine*)h;
  if (t_end <= t_begin) return 0;
  int nb = threads < 1 ? 1 : threads;
  if (nb > t_end - t_begin) nb = t_end - t_begin;
  if (nb > e->n_ws) {
    e->ws = (go_work*)realloc(e->ws, (size_t)nb * sizeof(go_work));
    for (int b = e->n_ws; b < nb; ++b) work_init(e, &e->ws[b]);
    e->n_ws = nb;
  }
  go_work* ws = e->ws;
#ifdef _OPENMP
#pragma omp parallel for num_threads(nb) schedule(static, 1)
#endif
  for (int b = 0; b < nb; ++b) {
    int64_t span = t_end - t_begin;
    int t0 = t_begin + (int)(span * b / nb), t1 = t_begin + (int)(span * (b + 1) / nb);
    run_block(e, &ws[b], t0, t1);
  }
  uint64_t pt = 0, cs = 0;
  for (int b = 0; b < e->n_ws; ++b) { pt += ws[b].pair_tests; cs += ws[b].contact_steps; }
  e->pair_tests = pt; e->contact_steps = cs;
  return 0;
}

uint64_t go_pair_count(void* h) { go_finalize((go_engine*)h); return ((go_engine*)h)->pairs.used; }
uint64_t go_coord_count(void* h) { go_finalize((go_engine*)h); return ((go_engine*)h)->coords.used; }
uint64_t go_pair_tests(void* h) { return ((go_engine*)h)->pair_tests; }
uint64_t go_contact_steps(void* h) { return ((go_engine*)h)->contact_steps; }

static int cmp_pair(const void* a, const void* b) {
  const citam_pair *p = (const citam_pair*)a, *q = (const citam_pair*)b;
  if (p->first_step != q->first_step) return p->first_step < q->first_step ? -1 : 1;
  if (p->a1 != q->a1) return p->a1 < q->a1 ? -1 : 1;
  return p->a2 < q->a2 ? -1 : (p->a2 > q->a2);
}
static int cmp_coord(const void* a, const void* b) {
  const citam_coord *p = (const citam_coord*)a, *q = (const citam_coord*)b;
  if (p->floor != q->floor) return p->floor < q->floor ? -1 : 1;
  if (p->sx != q->sx) return p->sx < q->sx ? -1 : 1;
  return p->sy < q->sy ? -1 : (p->sy > q->sy);
}

/* rows in first-contact order = dict insertion order of the reference (contacts.py:135-142) */
void go_export_pairs(void* h, citam_pair* out) {
  go_engine* e = (go_engine*)h;
  go_finalize(e);
  uint64_t n = 0;
  for (uint64_t i = 0; i < e->pairs.cap; ++i) {
    pair_ent* s = &e->pairs.e[i];
    if (s->key == EMPTY) continue;
    out[n].a1 = (uint32_t)(s->key >> 32); out[n].a2 = (uint32_t)s->key; out[n].first_step = s->first;
    out[n].n_events = s->nev; out[n].duration = s->dur; ++n;
  }
  qsort(out, n, sizeof *out, cmp_pair);
}
void go_export_coords(void* h, citam_coord* out) {
  go_engine* e = (go_engine*)h;
  go_finalize(e);
  uint64_t n = 0;
  for (uint64_t i = 0; i < e->coords.cap; ++i) {
    coord_ent* s = &e->coords.e[i];
    if (s->key == EMPTY) continue;
    out[n].floor = (int32_t)((s->key >> 56) & 0x7f);
    out[n].sx = (int32_t)((s->key >> 28) & 0xfffffffu) - (1 << 27);
    out[n].sy = (int32_t)(s->key & 0xfffffffu) - (1 << 27);
    out[n].reserved = 0; out[n].count = s->count; ++n;
  }
  qsort(out, n, sizeof *out, cmp_coord);
}
void go_export_durations(void* h, uint64_t* out) {
  go_engine* e = (go_engine*)h;
  go_finalize(e);
  memcpy(out, e->agent_dur, 8 * (size_t)e->N);
}

/* ---- the location rule (floorplan.py:215-241 first match; space.py:194-280) -------- */
static int in_box(double px, double py, double qx, double qy, double rx, double ry) { /* geometry.py:31-53 */
  return qx <= fmax(px, rx) && qx >= fmin(px, rx) && qy <= fmax(py, ry) && qy >= fmin(py, ry);
}
static int orient(double px, double py, double qx, double qy, double rx, double ry) { /* geometry.py:56-82 */
  double v = (qy - py) * (rx - qx) - (qx - px) * (ry - qy);
  return v == 0.0 ? 0 : (v > 0.0 ? 1 : 2);
}
static int crosses(double p1x, double p1y, double q1x, double q1y, double p2x, double p2y, double q2x, double q2y) {
  int o1 = orient(p1x, p1y, q1x, q1y, p2x, p2y), o2 = orient(p1x, p1y, q1x, q1y, q2x, q2y);
  int o3 = orient(p2x, p2y, q2x, q2y, p1x, p1y), o4 = orient(p2x, p2y, q2x, q2y, q1x, q1y);
  if (o1 != o2 && o3 != o4) return 1;
  if (o1 == 0 && in_box(p1x, p1y, p2x, p2y, q1x, q1y)) return 1;
  if (o2 == 0 && in_box(p1x, p1y, q2x, q2y, q1x, q1y)) return 1;
  if (o3 == 0 && in_box(p2x, p2y, p1x, p1y, q2x, q2y)) return 1;
  if (o4 == 0 && in_box(p2x, p2y, q1x, q1y, q2x, q2y)) return 1;
  return 0;
}

static int point_in_space(const citam_space* sp, const citam_boundary* bnd, const citam_door* doors, int xi, int yi) {
  if (xi < sp->bb_xmin || xi > sp->bb_xmax || yi < sp->bb_ymin || yi > sp->bb_ymax) return 0; /* space.py:209-216 */
  double x = xi, y = yi;
  for (int d = sp->door_begin; d < sp->door_end; ++d) /* space.py:218-224 */
    if (in_box(doors[d].sx, doors[d].sy, x, y, doors[d].ex, doors[d].ey)) return 1;
  for (int b = sp->seg_begin; b < sp->seg_end; ++b) { /* space.py:179-192 + geometry.py:558-588 */
    const citam_boundary* s = &bnd[b];
    if (!s->long_enough) continue;
    if ((x == s->sx && y == s->sy) || (x == s->ex && y == s->ey)) return 1;
    if (in_box(s->sx, s->sy, x, y, s->ex, s->ey)) {
      double ddx = x - s->sx, ddy = y - s->sy, hh = hypot(ddx, ddy);
      double n2x = ddy / hh, n2y = -(ddx / hh);
      if (fabs(s->nx - n2x) < 1e-3 && fabs(s->ny - n2y) < 1e-3) return 1;
    }
  }
  const double INF = 1e10; /* space.py:236-280 */
  double ex[8] = {INF, x - 5.0, -INF, x - 5.0, INF, INF, -INF, -INF};
  double ey[8] = {y + 5.0, INF, y + 5.0, -INF, INF, -INF, INF, -INF};
  for (int r = 0; r < 8; ++r) {
    int count = 0;
    for (int b = sp->seg_begin; b < sp->seg_end; ++b)
      count += crosses(bnd[b].sx, bnd[b].sy, bnd[b].ex, bnd[b].ey, x, y, ex[r], ey[r]);
    if (count & 1) return 1;
  }
  return 0;
}

int go_build_lut(int32_t minx, int32_t miny, int32_t W, int32_t H, int32_t n_spaces, const citam_space* spaces,
                 const citam_boundary* bnd, const citam_door* doors, int16_t* out, int32_t threads) {
#ifdef _OPENMP
#pragma omp parallel for num_threads(threads < 1 ? 1 : threads) schedule(dynamic, 8)
#endif
  for (int iy = 0; iy < H; ++iy)
    for (int ix = 0; ix < W; ++ix) {
      int r = -1;
      for (int s = 0; s < n_spaces; ++s)
        if (point_in_space(&spaces[s], bnd, doors, minx + ix, miny + iy)) { r = s; break; }
      out[(size_t)iy * W + ix] = (int16_t)r;
    }
  return 0;
}

int go_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
