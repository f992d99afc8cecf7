/* orb_b200.h — C ABI of the B200-native ORB hot path (liborb_b200.so).
 *
 * Drop-in boundary for the cut-search / partition / next-axis path of
 * andrinr/gpu-load-balance.  The reference has no FFI of its own: its host
 * (src/orbit.cpp) reaches the GPU through mdl2 services whose wire types are
 * `Cell[nCells]` in, `unsigned int[nCells]` out (headers under src/services).  Each entry
 * point below names the reference service / lines it replaces; the C++ service
 * wrappers in gpu-load-balance_b200/host/services call exactly these.
 *
 * Conventions: plain pointers and sizes, no C++/torch types; every function
 * returns 0 (ORB_OK) or a negative ORB_ERR_* code and never calls exit();
 * orb_last_error() gives the text.  One context per GPU rank; a context is
 * thread-compatible (one caller at a time), not thread-safe.
 * There is no CPU fallback: without a CUDA device orb_create() fails.
 */
#ifndef ORB_B200_H
#define ORB_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORB_OK 0
#define ORB_ERR_CUDA (-1)      /* a CUDA runtime call failed (text in orb_last_error) */
#define ORB_ERR_NCCL (-2)      /* NCCL missing or a collective failed */
#define ORB_ERR_ARG (-3)       /* bad argument */
#define ORB_ERR_RANGE (-4)     /* cells of a level do not tile [0, n_local) in order */
#define ORB_ERR_STATE (-5)     /* call out of sequence (e.g. no particles uploaded) */

/* Layout-identical to the reference's `struct Cell` (src/cell.h:9-17): 52 bytes,
 * `bool foundCut` at offset 16 + 3 padding bytes.  It is the wire format of
 * every service. */
typedef struct orb_cell {
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
} orb_cell;

typedef struct orb_ctx orb_ctx;

/* orb_build flags */
#define ORB_FULL_LEVELS 1u   /* split log2(d) levels (2^y leaves) instead of the reference's log2(d)-1 (orbit.cpp:102) */
#define ORB_TIGHT_BOX 2u     /* north-star extension: child axis/margins from the particle bounding box
                                (not in the reference, which uses geometric boxes, cell.h:102-126) */

typedef struct orb_build_stats {
    int32_t n_levels;            /* split levels executed */
    int32_t iters[64];           /* bisection iterations per level (the reference's j, orbit.cpp:149) */
    int32_t passes[64];          /* HBM count passes per level (each evaluates up to 2^m-1 trial cuts) */
    int32_t not_found[64];       /* cells that hit the 32-iteration cap */
    uint64_t active_passes;      /* sum over HBM count passes of local particles in active cells (4 B each streamed) */
    uint64_t iter_particle_passes; /* the same weighted by bisection iterations: the reference-equivalent
                                      "particle-passes" (what orbit.cpp's loop would have streamed) */
    uint64_t count_launches;     /* kernel launches by category (this rank) */
    uint64_t update_launches;
    uint64_t partition_launches;
    uint64_t other_launches;
    float ms_count;              /* CUDA-event time inside count kernels (only if ORB_PROFILE env or profile flag set) */
    float ms_partition;
    float ms_total;              /* CUDA-event time of the whole build on the context's stream */
    uint32_t search_fallback_cells; /* cells the selection-based cut search left to the iterative bisection (massive
                                       ties, candidates beyond shared memory); the same number on every rank */
} orb_build_stats;

/* ---- lifetime (replaces ServiceInit / ServiceFinalize allocation: init.cu:85-141, finalize.cu:15-45) ---- */
int orb_create(orb_ctx **ctx, int device, uint64_t n_local, uint32_t n_leaf_cells);
int orb_destroy(orb_ctx *ctx);
const char *orb_last_error(void);
int orb_version(void);

/* tuning: trial-cut depth m per count pass (1..3: 2^m-1 cuts per cell per HBM pass); 0 = default */
int orb_set_trial_depth(orb_ctx *ctx, int m);
/* tie handling of the partition: 0 = canonical (stable, x < cut goes left; default),
 * 1 = Hoare-exact: reproduces partition.cpp:30-60 including which particles with coord == cut go to which side and
 * the particle order (in place; also selected by env ORB_TIES=hoare).  Cells without any particle >= cut make the
 * reference touch rows outside the cell; they are reported as ORB_ERR_RANGE instead of emulated. */
#define ORB_TIES_CANONICAL 0
#define ORB_TIES_HOARE 1
int orb_set_tie_mode(orb_ctx *ctx, int mode);
/* profiling: when on, every count / partition kernel launch is bracketed by CUDA events on the context's
 * stream and orb_build_stats.ms_count / ms_partition are filled (also enabled by env ORB_PROFILE=1) */
int orb_set_profile(orb_ctx *ctx, int on);

/* ---- diagnostic: how a level would be searched (pure host arithmetic, needs no device) ----
 * A context of n_local particles on one of n_ranks ranks (n_global particles in all, smallest shard n_local_min), level
 * of n_cells cells, prev_cells_prefuse = whether the previous level's partition may build this level's histogram rows
 * (ORB_PREFUSE semantics: -1 automatic, 0 never, 1 always).  Everything that shapes a multi-rank exchange depends on
 * (n_cells, n_global, n_local_min, n_ranks) only - never on n_local - so that all ranks take the same branch. */
typedef struct orb_level_plan {
    int32_t search;        /* 0 iterative count kernels, 1 selection: streaming passes, 2 selection: one block per cell,
                              3 selection over several ranks (two exchanges) */
    int32_t hist_bins;     /* bins per cell of the histogram rows (0: none) */
    uint32_t cand_cap;     /* candidates one finish block stages */
    uint32_t slot_words;   /* several ranks: words of one rank's candidate slot per cell (last word = count) */
    uint64_t hist_words;   /* words of histogram rows the level needs in global memory */
    int32_t prefuse_bins;  /* != 0: the previous level's partition builds this level's rows with this many bins */
    int32_t sample_stride; /* > 1: the rows come from a sample (every sample_stride-th tile / piece of a cell), the gathering
                            * pass proves the bracket with exact counts (one rank, >= 2^25 particles per GPU) */
} orb_level_plan;
int orb_plan_level(uint64_t n_local, uint64_t n_global, uint64_t n_local_min, int n_ranks, uint32_t n_leaf_cells,
                   uint32_t n_cells, int prefuse_mode, orb_level_plan *out);

/* ---- multi-GPU (replaces the mdl2 reduce tree, TraversePST.cpp:28-44 + Combine in countLeft.cpp:44-53) ----
 * One rank per GPU.  Either let the library own the communicator (rank 0 makes an id, the caller
 * broadcasts the 128 bytes by any means, every rank calls orb_comm_init), or attach an existing
 * ncclComm_t (thread-per-GPU hosts using ncclCommInitAll). */
int orb_comm_unique_id(void *id128);
int orb_comm_init(orb_ctx *ctx, const void *id128, int rank, int n_ranks);
int orb_comm_attach(orb_ctx *ctx, void *nccl_comm, int rank, int n_ranks);

/* Optional: fuse the count combine into the bisection-update kernel.  Each rank exports a descriptor of its receive
 * rows (cudaMalloc memory: CUDA IPC handle for other processes, raw pointer for rank threads of the same process);
 * the caller gathers all descriptors and imports the table on every rank (after orb_comm_init / orb_comm_attach).
 * From then on, levels of up to 8192 cells combine their per-cell counts inside the update kernel: every block pushes
 * its count rows to all peers over NVLink, flags them, waits for the peers' rows and decides - no collective call
 * between count and update; larger levels keep NCCL. */
typedef struct orb_peer_info {
    uint8_t ipc_cnt[64];
    uint8_t ipc_flag[64];
    uint64_t ptr_cnt;
    uint64_t ptr_flag;
    int64_t pid;
    int32_t device;
    int32_t reserved_;
    uint8_t ipc_xchg[64];   /* exchange arena of the selection search: flags | slot fill levels | candidate slots | histogram rows */
    uint64_t ptr_xchg;
} orb_peer_info;
int orb_peer_export(orb_ctx *ctx, orb_peer_info *out);
int orb_peer_import(orb_ctx *ctx, const orb_peer_info *all, int n_ranks);

/* ---- particles (replaces ServiceCopyParticles o=2: copyParticles.cu:29-57) ---- */
int orb_upload_xyz(orb_ctx *ctx, const float *x, const float *y, const float *z);     /* host -> device, sets range[0]=[0,n_local) */
int orb_load_device_xyz(orb_ctx *ctx, const float *dx, const float *dy, const float *dz); /* device -> device */
int orb_download_xyz(orb_ctx *ctx, float *x, float *y, float *z);                     /* device -> host (current order) */
int orb_device_xyz(orb_ctx *ctx, const float **dx, const float **dy, const float **dz);

/* ---- service-granular entry points (same granularity as the reference's RunService calls) ----
 * `cells` is the level's Cell array exactly as master() passes it (orbit.cpp:111-112). */
/* ServiceCount (count.cpp:8-30): out[c] = particles in cell c, summed over ranks */
int orb_count(orb_ctx *ctx, const orb_cell *cells, uint32_t n_cells, uint32_t *out);
/* ServiceCountLeft / CountLeftGPU / CountLeftGPUAxis (countLeft.cpp:9-53, countLeftGPUAxis.cu:188-270):
 * out[c] = #{p in cell c : coord_axis(c)[p] < getCut(c)} summed over ranks; entries of cells with
 * foundCut set are left untouched (countLeft.cpp:19-21). */
int orb_count_left(orb_ctx *ctx, const orb_cell *cells, uint32_t n_cells, uint32_t *out);
/* ServicePartition / PartitionGPU (partition.cpp:18-65): split every cell at getCut() (stable, x<cut left),
 * children get ranges [begin,begin+nLeft) and [begin+nLeft,end) (partition.cpp:54-60). */
int orb_partition(orb_ctx *ctx, const orb_cell *cells, uint32_t n_cells);
/* new service (north star): per-cell particle bounding box, out[6c..6c+5] = min xyz, max xyz over all ranks */
int orb_bbox(orb_ctx *ctx, const orb_cell *cells, uint32_t n_cells, float *out6);
/* cellToRangeMap read-back (init.cu:63-66): out[2i],out[2i+1] = local [begin,end) of cell id first_id+i */
int orb_get_ranges(orb_ctx *ctx, uint32_t first_id, uint32_t n, uint32_t *out);

/* ---- fused entry points (no host round trip per bisection iteration) ----
 * Whole bisection loop of one level (orbit.cpp:146-232) on the device: on return cells[] carry the final
 * margins / foundCut exactly as master() would have left them; iters/passes optional. */
int orb_find_cuts(orb_ctx *ctx, orb_cell *cells, uint32_t n_cells, int32_t *iters, int32_t *passes);
/* Whole build (orbit.cpp:74-275): root cell -> heap of 2d-1 cells, all levels, count+bisect+split+partition.
 * heap_out (host, 2d-1 cells) may be NULL. */
int orb_build(orb_ctx *ctx, uint32_t flags, orb_cell *heap_out, orb_build_stats *stats);

/* ---- host-side helpers shared by every driver (deterministic inputs; no device work) ----
 * Reference generator (init.cu:11-25,47-53): one xorshf96 stream, x0,y0,z0,x1,...; `skip` particles are
 * drawn and discarded first so rank r can take its contiguous slice of the single stream. */
void orb_generate_uniform(uint64_t skip, uint64_t n, float *x, float *y, float *z);
/* Clustered inputs (not in the reference): kind 0 = Gaussian clumps, 1 = Plummer spheres */
void orb_generate_clustered(int kind, uint64_t skip, uint64_t n, float *x, float *y, float *z);

#ifdef __cplusplus
}
#endif
#endif
