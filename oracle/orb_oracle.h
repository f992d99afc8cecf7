/* TEST INFRASTRUCTURE ONLY — CPU oracle for the ORB hot path.
 *
 * A plain-C restatement of the reference's CPU-only mode (`orbit <x> <y> 0`).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this; the product (gpu-load-balance_b200/)
 * never does.
 *
 * PARITY IS PINNED: oracle/_ref/orbit_ref is the UNMODIFIED reference compiled
 * from /root/reference/src (oracle/Makefile); with ORB_REF_TRACE it dumps every
 * service call.  orb_oracle_build() writes the same trace format and
 * tests/test_oracle_vs_reference.py requires the two files to be byte-identical
 * (cells, margins, foundCut flags, counts per bisection iteration, child
 * ranges, ordered particle hashes).  Golden traces made that way are committed
 * under tests/golden/ (tests/golden/make_golden.py).
 *
 * Each function cites the reference file:line it follows.
 */
#ifndef ORB_ORACLE_H
#define ORB_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Layout-identical to the reference's `struct Cell` (cell.h:9-17): 52 bytes,
 * `bool foundCut` at offset 16 followed by 3 padding bytes. */
typedef struct orb_oracle_cell {
    int32_t id;
    int32_t nLeafCells;
    int32_t prevCutAxis;
    int32_t cutAxis;
    uint8_t foundCut;
    uint8_t pad_[3];
    float cutMarginLeft;
    float cutMarginRight;
    float lower[3];
    float upper[3];
} orb_oracle_cell;

enum { ORB_TIES_HOARE = 0, ORB_TIES_CANONICAL = 1 };

typedef struct orb_oracle_params {
    int32_t d;            /* number of leaf cells (the CLI passes 1<<y, orbit.cpp:42) */
    int32_t full_levels;  /* 0: reference loop bound l < getNLevels() (orbit.cpp:102, one level short);
                             1: l <= getNLevels() (the 2^y leaves the README promises) */
    int32_t ties;         /* ORB_TIES_HOARE: verbatim partition.cpp:30-60; ORB_TIES_CANONICAL: stable x<cut split */
    int32_t n_shards;     /* the reference's mdl threads: static particle shards (orbit.cpp:83) */
    int32_t n_threads;    /* host threads used to process the shards */
    int32_t max_iter;     /* 32 (orbit.cpp:149) */
    int32_t tight_box;    /* 0: reference (geometric boxes, cell.h:102-126); 1: north-star extension —
                             child axis/margins from the particle bounding box */
    int32_t trace_particles;
    const char *trace_path; /* optional; only with n_shards == 1 */
} orb_oracle_params;

typedef struct orb_oracle_stats {
    int32_t n_levels;
    int32_t iters[64];          /* bisection iterations per level */
    int32_t not_found[64];      /* cells still unfound when the level loop ended */
    uint64_t active_passes;     /* sum over iterations of particles in unfound cells */
    uint64_t tie_particles;     /* particles with coord == cut at partition time */
    double t_count_s, t_partition_s, t_makeaxis_s, t_total_s;
} orb_oracle_stats;

/* init.cu:11-25 — the reference's generator, state exposed so shards can continue one stream. */
typedef struct orb_xorshf96_state { uint64_t x, y, z; } orb_xorshf96_state;
void orb_oracle_xorshf96_init(orb_xorshf96_state *s);
float orb_oracle_xorshf96(orb_xorshf96_state *s);
/* init.cu:47-53 — for i: for d: particles(i,d) = xorshf96(); columns are x,y,z */
void orb_oracle_generate_uniform(orb_xorshf96_state *s, float *x, float *y, float *z, uint64_t n);

/* cell.h:19-45,74-126 */
void orb_oracle_cell_init(orb_oracle_cell *c, int id, int nLeafCells, const float *lower, const float *upper);
float orb_oracle_cell_get_cut(const orb_oracle_cell *c);
void orb_oracle_cell_cut(const orb_oracle_cell *c, orb_oracle_cell *left, orb_oracle_cell *right);
void orb_oracle_cell_set_cut_axis(orb_oracle_cell *c);
void orb_oracle_cell_set_cut_margin(orb_oracle_cell *c);
int orb_oracle_n_levels(int nLeafCells);
int orb_oracle_n_cells_on_last_level(int nLeafCells);

/* countLeft.cpp:31-36 */
uint32_t orb_oracle_count_left(const float *col, int64_t begin, int64_t end, float cut);
/* orbit.cpp:204-229 — one bisection decision; returns 1 if the cell is now found */
int orb_oracle_bisect_step(orb_oracle_cell *c, uint32_t countLeft, uint32_t count);
/* canonical stable split: returns begin + #{p : col_axis[p] < cut} */
int64_t orb_oracle_partition_canonical(float *x, float *y, float *z, int64_t begin, int64_t end, int axis, float cut);
/* per-cell particle bounding box (north-star extension): out = min[3], max[3] */
void orb_oracle_bbox(const float *x, const float *y, const float *z, int64_t begin, int64_t end, float *out6);

/* Whole build.  x,y,z hold all shards back to back: shard s owns [shard_off[s], shard_off[s+1]).
 * heap: (2d-1) cells.  ranges: n_shards * (2d-1) * 2 shard-local indices (cellToRangeMap, init.cu:63-66). */
int orb_oracle_build(const orb_oracle_params *p, float *x, float *y, float *z, const uint64_t *shard_off,
                     orb_oracle_cell *heap, uint32_t *ranges, orb_oracle_stats *stats);

/* hashes used in traces and tests */
uint64_t orb_oracle_particle_hash(float x, float y, float z);
uint64_t orb_oracle_fnv1a(const void *bytes, size_t n);   /* FNV-1a 64 over a byte string (heapHash of the CLI) */
void orb_oracle_range_hashes(const float *x, const float *y, const float *z, int64_t begin, int64_t end,
                             uint64_t *set_hash, uint64_t *ordered_hash);

#ifdef __cplusplus
}
#endif
#endif
