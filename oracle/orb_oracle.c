/* TEST INFRASTRUCTURE ONLY — see orb_oracle.h.  CPU restatement of the
 * reference's o=0 path; every function cites the reference lines it follows.
 * Compile with -fno-fast-math -ffp-contract=off (oracle/Makefile): the float
 * bisection arithmetic must not be reassociated or fused.
 */
#include "orb_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <pthread.h>

/* ------------------------------------------------------------------ generator
 * init.cu:11-25.  `unsigned long` is 64-bit on the reference's platform (LP64).
 * Return expression: (float) z / ULONG_MAX - 0.5
 *   - (float) z                : uint64 -> float, round to nearest
 *   - / numeric_limits<unsigned long>::max() : the divisor is converted to float
 *     (= 2^64 exactly after rounding), division in float
 *   - - 0.5                    : 0.5 is a double literal, so the subtraction is in double
 *   - return                   : rounded to float
 */
void orb_oracle_xorshf96_init(orb_xorshf96_state *s) {
    s->x = 123456789ULL; s->y = 362436069ULL; s->z = 521288629ULL;
}

float orb_oracle_xorshf96(orb_xorshf96_state *s) {
    uint64_t t;
    s->x ^= s->x << 16;
    s->x ^= s->x >> 5;
    s->x ^= s->x << 1;
    t = s->x;
    s->x = s->y;
    s->y = s->z;
    s->z = t ^ s->x ^ s->y;
    float q = (float)s->z / (float)UINT64_MAX;
    return (float)((double)q - 0.5);
}

/* init.cu:47-53 */
void orb_oracle_generate_uniform(orb_xorshf96_state *s, float *x, float *y, float *z, uint64_t n) {
    for (uint64_t i = 0; i < n; ++i) {
        x[i] = orb_oracle_xorshf96(s);
        y[i] = orb_oracle_xorshf96(s);
        z[i] = orb_oracle_xorshf96(s);
    }
}

/* ------------------------------------------------------------------ Cell (cell.h) */

/* cell.h:19-45 */
void orb_oracle_cell_init(orb_oracle_cell *c, int id, int nLeafCells, const float *lower, const float *upper) {
    memset(c, 0, sizeof(*c));
    c->id = id;
    c->nLeafCells = nLeafCells;
    c->foundCut = 0;
    c->prevCutAxis = -1;
    c->cutAxis = -1;
    c->cutMarginLeft = 0.0f;
    c->cutMarginRight = 0.0f;
    for (int k = 0; k < 3; ++k) { c->lower[k] = lower[k]; c->upper[k] = upper[k]; }
}

/* cell.h:65-67: ceil(log2(nLeafCells)) evaluated in double, truncated to int */
int orb_oracle_n_levels(int nLeafCells) { return (int)ceil(log2((double)nLeafCells)); }

/* cell.h:69-72 */
int orb_oracle_n_cells_on_last_level(int nLeafCells) {
    int depth = orb_oracle_n_levels(nLeafCells);
    return (int)(2 * nLeafCells - pow(2, depth));
}

/* cell.h:74-76: float add, then /2.0 in double, rounded to float on return */
float orb_oracle_cell_get_cut(const orb_oracle_cell *c) {
    float sum = c->cutMarginRight + c->cutMarginLeft;
    return (float)((double)sum / 2.0);
}

/* cell.h:78-100 */
void orb_oracle_cell_cut(const orb_oracle_cell *c, orb_oracle_cell *left, orb_oracle_cell *right) {
    int nCellsLeft = (int)ceil(c->nLeafCells / 2.0);
    int nCellsRight = c->nLeafCells - nCellsLeft;
    float cut = orb_oracle_cell_get_cut(c);
    orb_oracle_cell l, r;
    orb_oracle_cell_init(&l, (c->id + 1) * 2 - 1, nCellsLeft, c->lower, c->upper);
    l.upper[c->cutAxis] = cut;
    l.prevCutAxis = c->cutAxis;
    orb_oracle_cell_init(&r, (c->id + 1) * 2, nCellsRight, c->lower, c->upper);
    r.lower[c->cutAxis] = cut;
    r.prevCutAxis = c->cutAxis;
    *left = l;
    *right = r;
}

/* cell.h:102-121: strict '>' from maxSize = 0 => lowest axis wins ties, all-zero extents => -1 */
void orb_oracle_cell_set_cut_axis(orb_oracle_cell *c) {
    int maxD = -1;
    float maxSize = 0.0f;
    for (int d = 0; d < 3; ++d) {
        float size = c->upper[d] - c->lower[d];
        if (size > maxSize) { maxSize = size; maxD = d; }
    }
    c->cutAxis = maxD;
}

/* cell.h:123-126 */
void orb_oracle_cell_set_cut_margin(orb_oracle_cell *c) {
    c->cutMarginLeft = c->lower[c->cutAxis];
    c->cutMarginRight = c->upper[c->cutAxis];
}

/* ------------------------------------------------------------------ services */

/* countLeft.cpp:31-36 */
uint32_t orb_oracle_count_left(const float *col, int64_t begin, int64_t end, float cut) {
    int nLeft = 0;
    for (const float *p = col + begin; p < col + end; ++p) nLeft += *p < cut;
    return (uint32_t)nLeft;
}

/* orbit.cpp:204-229 (FAST_MEDIAN is never enabled, orbit.cpp:54,59,65).
 *   float ratio = ceil(nLeafCells / 2.0) / nLeafCells;      double math, rounded to float
 *   int difference = oCountsLeft[i] - oCounts[i] * ratio;   unsigned*float -> float; unsigned-float -> float; trunc
 */
int orb_oracle_bisect_step(orb_oracle_cell *c, uint32_t countLeft, uint32_t count) {
    float ratio = (float)(ceil(c->nLeafCells / 2.0) / c->nLeafCells);
    float prod = (float)count * ratio;
    float fdiff = (float)countLeft - prod;
    int difference = (int)fdiff;
    if (abs(difference) < 3) {
        c->foundCut = 1;
        return 1;
    } else if (difference > 0) {
        c->cutMarginRight = orb_oracle_cell_get_cut(c);
    } else {
        c->cutMarginLeft = orb_oracle_cell_get_cut(c);
    }
    return 0;
}

static inline void swap3(float *x, float *y, float *z, int64_t a, int64_t b) {
    float t;
    t = x[a]; x[a] = x[b]; x[b] = t;
    t = y[a]; y[a] = y[b]; y[b] = t;
    t = z[a]; z[a] = z[b]; z[b] = t;
}

/* Canonical tie mode: stable split with predicate col[p] < cut. */
int64_t orb_oracle_partition_canonical(float *x, float *y, float *z, int64_t begin, int64_t end, int axis, float cut) {
    int64_t n = end - begin;
    if (n <= 0) return begin;
    float *col = axis == 0 ? x : (axis == 1 ? y : z);
    float *tmp = (float *)malloc((size_t)n * 3 * sizeof(float));
    int64_t nl = 0;
    for (int64_t p = begin; p < end; ++p) nl += col[p] < cut;
    int64_t il = 0, ir = nl;
    for (int64_t p = begin; p < end; ++p) {
        int64_t dst = (col[p] < cut) ? il++ : ir++;
        tmp[dst] = x[p]; tmp[n + dst] = y[p]; tmp[2 * n + dst] = z[p];
    }
    memcpy(x + begin, tmp, (size_t)n * sizeof(float));
    memcpy(y + begin, tmp + n, (size_t)n * sizeof(float));
    memcpy(z + begin, tmp + 2 * n, (size_t)n * sizeof(float));
    free(tmp);
    return begin + nl;
}

/* partition.cpp:30-60, verbatim, on the reference's storage layout: one block
 * P of (N,3) floats, column-major (init.cu:32-45), so P(i,d) = P[i + d*N] and a
 * read one past a column lands in the next column exactly as in the reference.
 * `P` has one guard float before and after the 3N payload for the reads the
 * reference performs outside the array (i == end == N on the z column,
 * j == begin-1 == -1 on the x column); the guards hold +inf / are never decisive.
 * Returns i (partition.cpp:56-58). */
static int64_t partition_hoare_block(float *P, int64_t N, int64_t beginInd, int64_t endInd, int axis, float cut) {
#define PP(i, d) P[(i) + (int64_t)(d) * N]
    int64_t i = beginInd - 1, j = endInd;
    for (;;) {
        do { i++; } while (PP(i, axis) < cut && i <= endInd);
        do { j--; } while (PP(j, axis) > cut && j >= beginInd);
        if (i >= j) break;
        for (int d = 0; d < 3; ++d) { float t = PP(i, d); PP(i, d) = PP(j, d); PP(j, d) = t; }
    }
    /* partition.cpp:52: swap(particles, i, endInd - 1) */
    for (int d = 0; d < 3; ++d) { float t = PP(i, d); PP(i, d) = PP(endInd - 1, d); PP(endInd - 1, d) = t; }
#undef PP
    return i;
}

void orb_oracle_bbox(const float *x, const float *y, const float *z, int64_t begin, int64_t end, float *out6) {
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int64_t p = begin; p < end; ++p) {
        if (x[p] < mn[0]) mn[0] = x[p];
        if (x[p] > mx[0]) mx[0] = x[p];
        if (y[p] < mn[1]) mn[1] = y[p];
        if (y[p] > mx[1]) mx[1] = y[p];
        if (z[p] < mn[2]) mn[2] = z[p];
        if (z[p] > mx[2]) mx[2] = z[p];
    }
    for (int k = 0; k < 3; ++k) { out6[k] = mn[k]; out6[3 + k] = mx[k]; }
}

/* ------------------------------------------------------------------ hashes (same as oracle/ref_shim/ref_tap.cpp) */
static inline uint64_t mix64(uint64_t v) {
    v ^= v >> 30; v *= 0xbf58476d1ce4e5b9ULL;
    v ^= v >> 27; v *= 0x94d049bb133111ebULL;
    v ^= v >> 31;
    return v;
}
static inline uint32_t fbits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

uint64_t orb_oracle_particle_hash(float x, float y, float z) {
    uint64_t a = ((uint64_t)fbits(x) << 32) | fbits(y);
    uint64_t b = (uint64_t)fbits(z) | 0x9e3779b900000000ULL;
    return mix64(a ^ mix64(b));
}

uint64_t orb_oracle_fnv1a(const void *bytes, size_t n) {
    const unsigned char *b = (const unsigned char *)bytes;
    uint64_t h = 1469598103934665603ULL;
    for (size_t i = 0; i < n; ++i) h = (h ^ b[i]) * 1099511628211ULL;
    return h;
}

void orb_oracle_range_hashes(const float *x, const float *y, const float *z, int64_t begin, int64_t end,
                             uint64_t *set_hash, uint64_t *ordered_hash) {
    uint64_t set = 0, ord = 1469598103934665603ULL;
    for (int64_t i = begin; i < end; ++i) {
        uint64_t ph = orb_oracle_particle_hash(x[i], y[i], z[i]);
        set += ph;
        ord = (ord ^ ph) * 1099511628211ULL;
    }
    *set_hash = set;
    *ordered_hash = ord;
}

/* ------------------------------------------------------------------ trace (format of ref_tap.cpp) */
enum { SID_INIT = 2, SID_COUNTLEFT = 8, SID_PARTITION = 9, SID_COUNT = 11, REC_PARTICLES = 1000 }; /* pst.h:66-81 */

static void rec_header(FILE *f, uint32_t kind, uint32_t nCells, uint64_t bytes) {
    fwrite(&kind, 4, 1, f); fwrite(&nCells, 4, 1, f); fwrite(&bytes, 8, 1, f);
}
static void rec_particles(FILE *f, const float *x, const float *y, const float *z, uint64_t n) {
    rec_header(f, REC_PARTICLES, 0, 4 + n * 12);
    uint32_t n32 = (uint32_t)n;
    fwrite(&n32, 4, 1, f);
    fwrite(x, 4, n, f); fwrite(y, 4, n, f); fwrite(z, 4, n, f);
}

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* ------------------------------------------------------------------ shard-parallel helper
 * The reference runs one mdl thread per shard (orbit.cpp:83, TraversePST.cpp:28-44).
 * Here shards are dealt round-robin to n_threads pthreads per phase. */
typedef void (*shard_fn)(void *ctx, int s);
typedef struct par_job { shard_fn fn; void *ctx; int t, nThreads, nShards; } par_job;
static void *par_entry(void *v) {
    par_job *j = (par_job *)v;
    for (int s = j->t; s < j->nShards; s += j->nThreads) j->fn(j->ctx, s);
    return NULL;
}
static void par_shards(int nThreads, int nShards, shard_fn fn, void *ctx) {
    if (nThreads > nShards) nThreads = nShards;
    if (nThreads <= 1) { for (int s = 0; s < nShards; ++s) fn(ctx, s); return; }
    pthread_t *th = (pthread_t *)malloc((size_t)nThreads * sizeof(pthread_t));
    par_job *jb = (par_job *)malloc((size_t)nThreads * sizeof(par_job));
    for (int t = 0; t < nThreads; ++t) {
        jb[t].fn = fn; jb[t].ctx = ctx; jb[t].t = t; jb[t].nThreads = nThreads; jb[t].nShards = nShards;
        if (t) pthread_create(&th[t], NULL, par_entry, &jb[t]);
    }
    par_entry(&jb[0]);
    for (int t = 1; t < nThreads; ++t) pthread_join(th[t], NULL);
    free(th); free(jb);
}

/* ------------------------------------------------------------------ whole build: orbit.cpp:68-275 */
typedef struct shard {
    int64_t n;          /* particles in this shard (orbit.cpp:83: N / Threads) */
    float *x, *y, *z;   /* caller's columns for this shard */
    float *P;           /* hoare mode: guarded column-major block, payload at P[0..3n) */
    float *T;           /* particlesT (init.cu:70), filled by MakeAxis */
    uint32_t *range;    /* cellToRangeMap: (2d-1) x 2 */
} shard;

typedef struct build_ctx {
    const orb_oracle_params *p;
    shard *sh;
    orb_oracle_cell *cells;   /* the level's slice of the heap */
    int nCells, d;
    uint32_t *part;           /* per-shard partial counts: n_shards x (d+1) */
    uint64_t *ties;           /* per-shard tie counters */
} build_ctx;

static inline float *shard_col(const build_ctx *b, int s, int a) {
    shard *h = &b->sh[s];
    if (b->p->ties == ORB_TIES_HOARE) return h->P + (int64_t)a * h->n;
    return a == 0 ? h->x : (a == 1 ? h->y : h->z);
}

/* MakeAxis: makeAxis.cpp:20-28 (copies end-begin+1 floats: blitz Range is inclusive) */
static void phase_make_axis(void *v, int s) {
    build_ctx *b = (build_ctx *)v;
    shard *h = &b->sh[s];
    for (int c = 0; c < b->nCells; ++c) {
        int64_t bg = h->range[2 * b->cells[c].id], en = h->range[2 * b->cells[c].id + 1];
        int64_t cnt = en - bg + 1;
        if (bg + cnt > h->n) cnt = h->n - bg;   /* the overrun float is never read back */
        if (cnt > 0) memcpy(h->T + bg, shard_col(b, s, b->cells[c].cutAxis) + bg, (size_t)cnt * sizeof(float));
    }
}

/* CountLeft service body: countLeft.cpp:16-39 */
static void phase_count_left(void *v, int s) {
    build_ctx *b = (build_ctx *)v;
    shard *h = &b->sh[s];
    uint32_t *out = b->part + (size_t)s * ((size_t)b->d + 1);
    for (int c = 0; c < b->nCells; ++c) {
        if (b->cells[c].foundCut) continue;                              /* countLeft.cpp:19-21 */
        int64_t bg = h->range[2 * b->cells[c].id], en = h->range[2 * b->cells[c].id + 1];
        out[c] = orb_oracle_count_left(h->T, bg, en, orb_oracle_cell_get_cut(&b->cells[c]));
    }
}

/* statistics only (untimed): particles sitting exactly on the cut */
static void phase_count_ties(void *v, int s) {
    build_ctx *b = (build_ctx *)v;
    shard *h = &b->sh[s];
    uint64_t ties = 0;
    for (int c = 0; c < b->nCells; ++c) {
        int64_t bg = h->range[2 * b->cells[c].id], en = h->range[2 * b->cells[c].id + 1];
        float cut = orb_oracle_cell_get_cut(&b->cells[c]);
        const float *col = shard_col(b, s, b->cells[c].cutAxis);
        for (int64_t q = bg; q < en && q < h->n; ++q) ties += col[q] == cut;
    }
    b->ties[s] = ties;
}

/* Partition service body: partition.cpp:18-65 */
static void phase_partition(void *v, int s) {
    build_ctx *b = (build_ctx *)v;
    shard *h = &b->sh[s];
    for (int c = 0; c < b->nCells; ++c) {
        int id = b->cells[c].id, axis = b->cells[c].cutAxis;
        int64_t bg = h->range[2 * id], en = h->range[2 * id + 1];
        float cut = orb_oracle_cell_get_cut(&b->cells[c]);
        int64_t i;
        if (b->p->ties == ORB_TIES_HOARE) i = partition_hoare_block(h->P, h->n, bg, en, axis, cut);
        else i = orb_oracle_partition_canonical(h->x, h->y, h->z, bg, en, axis, cut);
        int lid = (id + 1) * 2 - 1, rid = (id + 1) * 2;                   /* partition.cpp:54-60 */
        h->range[2 * lid] = (uint32_t)bg; h->range[2 * lid + 1] = (uint32_t)i;
        h->range[2 * rid] = (uint32_t)i;  h->range[2 * rid + 1] = (uint32_t)en;
    }
}

int orb_oracle_build(const orb_oracle_params *p, float *x, float *y, float *z, const uint64_t *shard_off,
                     orb_oracle_cell *heap, uint32_t *ranges, orb_oracle_stats *stats) {
    const int d = p->d;
    if (d < 1 || (d & (d - 1)) != 0) return -1;           /* the CLI only produces powers of two */
    const int nShards = p->n_shards > 0 ? p->n_shards : 1;
    const int maxIter = p->max_iter > 0 ? p->max_iter : 32;
    const int nHeap = 2 * d - 1;
    int nThreads = p->n_threads > 0 ? p->n_threads : 1;
    if (p->trace_path && nShards != 1) return -2;

    orb_oracle_stats st;
    memset(&st, 0, sizeof(st));

    shard *sh = (shard *)calloc((size_t)nShards, sizeof(shard));
    for (int s = 0; s < nShards; ++s) {
        sh[s].n = (int64_t)(shard_off[s + 1] - shard_off[s]);
        sh[s].x = x + shard_off[s]; sh[s].y = y + shard_off[s]; sh[s].z = z + shard_off[s];
        sh[s].T = (float *)malloc(((size_t)sh[s].n + 2) * sizeof(float));
        sh[s].range = ranges + (size_t)s * nHeap * 2;
        memset(sh[s].range, 0, (size_t)nHeap * 2 * sizeof(uint32_t));
        sh[s].range[0] = 0;                       /* init.cu:64-65 */
        sh[s].range[1] = (uint32_t)sh[s].n;
        if (p->ties == ORB_TIES_HOARE) {
            int64_t n = sh[s].n;
            float *blk = (float *)malloc(((size_t)3 * n + 5) * sizeof(float));
            blk[0] = INFINITY;
            for (int g = 1; g <= 4; ++g) blk[3 * n + g] = INFINITY;
            sh[s].P = blk + 1;
            memcpy(sh[s].P, sh[s].x, (size_t)n * 4);
            memcpy(sh[s].P + n, sh[s].y, (size_t)n * 4);
            memcpy(sh[s].P + 2 * n, sh[s].z, (size_t)n * 4);
        }
    }
#define COL(s, a) (p->ties == ORB_TIES_HOARE ? sh[s].P + (int64_t)(a) * sh[s].n : ((a) == 0 ? sh[s].x : ((a) == 1 ? sh[s].y : sh[s].z)))

    FILE *tf = NULL;
    if (p->trace_path) {
        tf = fopen(p->trace_path, "wb");
        if (!tf) return -3;
        const char magic[8] = {'O', 'R', 'B', 'T', 'R', 'A', 'C', 'E'};
        uint32_t ver = 1, cb = (uint32_t)sizeof(orb_oracle_cell);
        fwrite(magic, 1, 8, tf); fwrite(&ver, 4, 1, tf); fwrite(&cb, 4, 1, tf);
        rec_header(tf, SID_INIT, 0, 4);
        uint32_t n32 = (uint32_t)sh[0].n;
        fwrite(&n32, 4, 1, tf);
        if (p->trace_particles) rec_particles(tf, COL(0, 0), COL(0, 1), COL(0, 2), (uint64_t)sh[0].n);
    }

    /* orbit.cpp:45-46,74-81 */
    float lower[3] = {-0.5f, -0.5f, -0.5f}, upper[3] = {0.5f, 0.5f, 0.5f};
    orb_oracle_cell root;
    orb_oracle_cell_init(&root, 0, d, lower, upper);
    if (p->tight_box) {
        float bb[6], g[6] = {INFINITY, INFINITY, INFINITY, -INFINITY, -INFINITY, -INFINITY};
        for (int s = 0; s < nShards; ++s) {
            orb_oracle_bbox(COL(s, 0), COL(s, 1), COL(s, 2), 0, sh[s].n, bb);
            for (int k = 0; k < 3; ++k) { if (bb[k] < g[k]) g[k] = bb[k]; if (bb[3 + k] > g[3 + k]) g[3 + k] = bb[3 + k]; }
        }
        if (g[0] <= g[3]) for (int k = 0; k < 3; ++k) { root.lower[k] = g[k]; root.upper[k] = g[3 + k]; }
        orb_oracle_cell_set_cut_axis(&root);
    } else {
        root.cutAxis = 0;
    }
    orb_oracle_cell_set_cut_margin(&root);
    memset(heap, 0, (size_t)nHeap * sizeof(orb_oracle_cell));
    heap[0] = root;

    /* orbit.cpp:99-100: these live across levels and iterations; found cells keep stale values */
    uint32_t *oCounts = (uint32_t *)calloc((size_t)d + 1, sizeof(uint32_t));
    uint32_t *oCountsLeft = (uint32_t *)calloc((size_t)d + 1, sizeof(uint32_t));
    uint32_t *part = (uint32_t *)malloc((size_t)nShards * ((size_t)d + 1) * sizeof(uint32_t));
    uint64_t *tieCnt = (uint64_t *)calloc((size_t)nShards, sizeof(uint64_t));
    build_ctx bc;
    bc.p = p; bc.sh = sh; bc.cells = NULL; bc.nCells = 0; bc.d = d; bc.part = part; bc.ties = tieCnt;

    const int nLevels = orb_oracle_n_levels(d);
    const int lEnd = p->full_levels ? nLevels + 1 : nLevels;     /* orbit.cpp:102 */
    double t0 = now_s();

    for (int l = 1; l < lEnd; ++l) {
        int a = (int)pow(2, l - 1) - 1;                                           /* orbit.cpp:104 */
        int lastLevel = orb_oracle_n_cells_on_last_level(d);
        int b2 = (int)pow(2, l);
        int b = (lastLevel < b2 ? lastLevel : b2) - 2;                            /* orbit.cpp:105-107 */
        int nCells = b - a + 1;
        orb_oracle_cell *cells = heap + a;                                        /* orbit.cpp:111 aliases the heap */

        bc.cells = cells; bc.nCells = nCells;
        double tm0 = now_s();
        par_shards(nThreads, nShards, phase_make_axis, &bc);
        st.t_makeaxis_s += now_s() - tm0;

        /* Count: count.cpp:16,28 */
        for (int c = 0; c < nCells; ++c) {
            uint32_t sum = 0;
            for (int s = 0; s < nShards; ++s) sum += sh[s].range[2 * cells[c].id + 1] - sh[s].range[2 * cells[c].id];
            oCounts[c] = sum;
        }
        if (tf) {
            rec_header(tf, SID_COUNT, (uint32_t)nCells, (uint64_t)nCells * (sizeof(orb_oracle_cell) + 4));
            fwrite(cells, sizeof(orb_oracle_cell), (size_t)nCells, tf);
            fwrite(oCounts, 4, (size_t)nCells, tf);
        }

        /* bisection loop: orbit.cpp:146-232 */
        int foundAll = 0, j = 0;
        while (!foundAll && j < maxIter) {
            j++;
            foundAll = 1;
            double tc0 = now_s();
            par_shards(nThreads, nShards, phase_count_left, &bc);
            for (int c = 0; c < nCells; ++c) {                                    /* Combine: countLeft.cpp:44-53 */
                if (cells[c].foundCut) continue;
                uint32_t sum = 0;
                for (int s = 0; s < nShards; ++s) {
                    sum += part[(size_t)s * ((size_t)d + 1) + c];
                    st.active_passes += sh[s].range[2 * cells[c].id + 1] - sh[s].range[2 * cells[c].id];
                }
                oCountsLeft[c] = sum;
            }
            st.t_count_s += now_s() - tc0;
            if (tf) {
                rec_header(tf, SID_COUNTLEFT, (uint32_t)nCells, (uint64_t)nCells * (sizeof(orb_oracle_cell) + 4));
                fwrite(cells, sizeof(orb_oracle_cell), (size_t)nCells, tf);
                fwrite(oCountsLeft, 4, (size_t)nCells, tf);
            }
            for (int c = 0; c < nCells; ++c) {                                    /* orbit.cpp:191-231 */
                if (cells[c].foundCut) continue;
                if (!orb_oracle_bisect_step(&cells[c], oCountsLeft[c], oCounts[c])) foundAll = 0;
            }
        }
        if (l < 64) {
            st.iters[l - 1] = j;
            int nf = 0;
            for (int c = 0; c < nCells; ++c) nf += !cells[c].foundCut;
            st.not_found[l - 1] = nf;
        }

        /* split: orbit.cpp:235-250 */
        for (int c = 0; c < nCells; ++c) {
            orb_oracle_cell L, R;
            orb_oracle_cell_cut(&cells[c], &L, &R);
            orb_oracle_cell_set_cut_axis(&R); orb_oracle_cell_set_cut_margin(&R);
            orb_oracle_cell_set_cut_axis(&L); orb_oracle_cell_set_cut_margin(&L);
            heap[L.id] = L;
            heap[R.id] = R;
        }

        par_shards(nThreads, nShards, phase_count_ties, &bc);
        for (int s = 0; s < nShards; ++s) st.tie_particles += bc.ties[s];

        double tp0 = now_s();
        par_shards(nThreads, nShards, phase_partition, &bc);
        st.t_partition_s += now_s() - tp0;

        /* north-star extension (not in the reference): tight boxes for the children */
        if (p->tight_box) {
            for (int c = 0; c < nCells; ++c) {
                for (int k = 0; k < 2; ++k) {
                    int cid = (cells[c].id + 1) * 2 - 1 + k;
                    float g[6] = {INFINITY, INFINITY, INFINITY, -INFINITY, -INFINITY, -INFINITY}, bb[6];
                    for (int s = 0; s < nShards; ++s) {
                        orb_oracle_bbox(COL(s, 0), COL(s, 1), COL(s, 2), sh[s].range[2 * cid], sh[s].range[2 * cid + 1], bb);
                        for (int q = 0; q < 3; ++q) { if (bb[q] < g[q]) g[q] = bb[q]; if (bb[3 + q] > g[3 + q]) g[3 + q] = bb[3 + q]; }
                    }
                    if (g[0] <= g[3]) {
                        for (int q = 0; q < 3; ++q) { heap[cid].lower[q] = g[q]; heap[cid].upper[q] = g[3 + q]; }
                        orb_oracle_cell_set_cut_axis(&heap[cid]);
                        if (heap[cid].cutAxis < 0) heap[cid].cutAxis = 0;   /* degenerate box: keep a valid axis */
                        orb_oracle_cell_set_cut_margin(&heap[cid]);
                    }
                }
            }
        }

        if (tf) {
            rec_header(tf, SID_PARTITION, (uint32_t)nCells, (uint64_t)nCells * (sizeof(orb_oracle_cell) + 16 + 32));
            fwrite(cells, sizeof(orb_oracle_cell), (size_t)nCells, tf);
            for (int c = 0; c < nCells; ++c) {
                int ids[2] = {(cells[c].id + 1) * 2 - 1, (cells[c].id + 1) * 2};
                uint32_t r[4]; uint64_t h[4];
                for (int k = 0; k < 2; ++k) {
                    int64_t bg = sh[0].range[2 * ids[k]], en = sh[0].range[2 * ids[k] + 1];
                    r[2 * k] = (uint32_t)bg; r[2 * k + 1] = (uint32_t)en;
                    if (en > sh[0].n) en = sh[0].n;
                    orb_oracle_range_hashes(COL(0, 0), COL(0, 1), COL(0, 2), bg, en, &h[2 * k], &h[2 * k + 1]);
                }
                fwrite(r, 4, 4, tf); fwrite(h, 8, 4, tf);
            }
            if (p->trace_particles) rec_particles(tf, COL(0, 0), COL(0, 1), COL(0, 2), (uint64_t)sh[0].n);
        }
        st.n_levels = l;
    }
    st.t_total_s = now_s() - t0;

    if (tf) fclose(tf);
    for (int s = 0; s < nShards; ++s) {
        if (p->ties == ORB_TIES_HOARE) {
            int64_t n = sh[s].n;
            memcpy(sh[s].x, sh[s].P, (size_t)n * 4);
            memcpy(sh[s].y, sh[s].P + n, (size_t)n * 4);
            memcpy(sh[s].z, sh[s].P + 2 * n, (size_t)n * 4);
            free(sh[s].P - 1);
        }
        free(sh[s].T);
    }
#undef COL
    free(sh); free(oCounts); free(oCountsLeft); free(part); free(tieCnt);
    if (stats) *stats = st;
    return 0;
}
