// orb_kernels.cuh — hand-written sm_100a kernels of the ORB hot path.
//
// Data layout in HBM (per rank): particles are SoA, three contiguous float[n_local]
// columns x,y,z (the reference's column-major blitz (N,3) array, init.cu:32-35),
// plus a second x,y,z set used as the partition's ping-pong target.  The cells of
// one tree level tile [0, n_local) in id order, so a level is described by one
// monotone boundary array bnd[nCells+1]; per-cell state is SoA (axis, margins,
// trial cuts, counters).
//
// Kernels (all HBM-bound integer/compare work; no tensor cores on this path):
//   k_count<NC>    all active cells of a level in ONE launch; NC = 2^m-1 trial cuts per cell
//                  per pass (the cuts of the next m bisection steps) so one HBM read serves
//                  m iterations of orbit.cpp:149-232.            4 B / active particle / pass
//   k_update<M>    one thread per cell: replays the reference's float decision rule
//                  (orbit.cpp:204-229) over the counted trial cuts; emits the next cuts.
//   k_partition_*  stable split of every cell in one launch: cooperative reduce-then-scan over
//                  contiguous tile ranges (or one block per small cell), scatter staged through
//                  shared memory.                              24 B / particle / level
//   k_bbox         per-cell min/max of x,y,z (north-star extension).  12 B / particle
//   k_level_setup, k_tile_map, k_split, k_finalize_*  O(nCells) bookkeeping.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/orb_b200.h"

namespace orb {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kMapTile = 2048;     // granularity of the tile -> first-cell map
constexpr int kCountTile = 4096;   // 256 threads x 4 float4
constexpr int kPartTile = 2048;    // 256 threads x 2 float4 per column
constexpr int kCS = 8;             // stride (in words) of per-cell cut / counter rows
constexpr int kMaxIter = 32;       // orbit.cpp:149
constexpr int kCountCellsSmem = 64;   // per-tile shared accumulators in the fragmented path
constexpr int kPartCells = 256;       // per-tile cell table in the partition (one entry per thread)

// ---- per-level device state (SoA over the cells of the level) ----
struct LevelState {
    uint32_t *bnd;        // [nCells+1] local particle index where cell c begins; bnd[nCells] = end of last
    int32_t *axis;        // [nCells]
    float *mL, *mR;       // [nCells] cutMarginLeft / cutMarginRight (live bisection bracket)
    uint32_t *total;      // [nCells] particles in the cell summed over ranks (ServiceCount)
    int32_t *nleaf;       // [nCells]
    uint32_t *active;     // [nCells] 1 while the cell still bisects
    uint32_t *found;      // [nCells]
    int32_t *iter;        // [nCells] bisection iterations consumed
    uint32_t *nleft_g;    // [nCells] global count left of the final cut
    uint32_t *nleft_l;    // [nCells] local count left of the final cut (partition offset)
    float *cuts;          // [nCells][kCS] trial cuts of the next pass, heap order of the bisection tree
    uint32_t *cnt_l;      // [nCells][kCS] local counters (k_count output)
    uint32_t *cnt_g;      // [nCells][kCS] counters summed over ranks (== cnt_l on one rank)
    // byte-reducing search: once the first pass of a level has fixed the bracket [compL, compR),
    // the second pass compacts the particles inside it; later passes only read those candidates
    float *compL, *compR; // [nCells] bracket at compaction time (-inf / +inf if that side never moved)
    uint32_t *base_l;     // [nCells] local particles below compL in compacted tiles (added to every later count)
    uint32_t *tile_ncand; // [nCountTiles][kWarps] candidates stored per (tile, warp)
};

// Peer view for the fused combine + bisection update (multi-GPU).  Every rank maps every rank's receive rows
// and flags (NVLink peer memory).  k_update block b pushes the 32-byte count rows of its cells to all peers with
// vector stores, raises flag[parity][self][b] on every peer, waits for the peers' block b, sums the rows and runs the
// reference's decision rule: the collective (Combine of countLeft.cpp:44-53) and the compute are one kernel, there is
// no NCCL call between count and update for levels of up to kPeerMaxCells cells.  Rows/flags are double-buffered by
// pass parity (a rank can be at most one pass ahead of a peer).
constexpr int kMaxPeers = 8;
constexpr uint32_t kPeerMaxCells = 8192;
constexpr uint32_t kPeerMaxBlocks = kPeerMaxCells / kThreads;
struct PeerSet {
    int n;                         // 0: disabled (single rank, or NCCL path)
    int self;
    uint32_t seq;                  // pass sequence number, identical on all ranks; parity = seq & 1
    uint32_t *recv[kMaxPeers];     // rank r's receive rows: [2][kMaxPeers (source)][kPeerMaxCells][kCS]
    uint32_t *flag[kMaxPeers];     // rank r's flags:        [2][kMaxPeers (source)][kPeerMaxBlocks]
};
__device__ __forceinline__ uint32_t *peer_rows(const PeerSet &ps, int r, int src) {
    return ps.recv[r] + ((size_t)(ps.seq & 1u) * kMaxPeers + src) * kPeerMaxCells * kCS;
}
__device__ __forceinline__ uint32_t *peer_flag(const PeerSet &ps, int r, int src, uint32_t block) {
    return ps.flag[r] + ((size_t)(ps.seq & 1u) * kMaxPeers + src) * kPeerMaxBlocks + block;
}

// Programmatic dependent launch (sm_90+): a kernel launched with the programmatic-stream-serialization attribute may
// start while its predecessor in the stream drains; pdl_wait() blocks until the predecessor has completed and its
// writes are visible, pdl_trigger() lets the successor start launching.  Both are no-ops for ordinary launches.
// Every kernel that is launched this way calls pdl_enter() before touching global memory.
__device__ __forceinline__ void pdl_enter() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

// float midpoint exactly as Cell::getCut (cell.h:74-76): float add, halve, round to float.
// (R+L)/2.0 in double then cast == correctly rounded half of the float sum == __fmul_rn(sum,0.5f).
__device__ __forceinline__ float mid_cut(float L, float R) { return __fmul_rn(__fadd_rn(R, L), 0.5f); }

// ---- bin function of the selection-based cut search (orb_select.cuh) ----
// scale of the bin function over [L, R): nb / (R - L), or 0 (everything in bin 0) for an empty or non-finite box
__device__ __forceinline__ float sel_scale(float L, float R, int nb) {
    const float inf = __int_as_float(0x7f800000);
    const float w = __fsub_rn(R, L);
    float s = (w > 0.f && w < inf) ? __fdiv_rn((float)nb, w) : 0.f;
    if (!(s < inf)) s = 0.f;
    return s;
}
// Bin of a coordinate.  The ONLY property the method needs is that this is monotone non-decreasing in x for every
// float (sub, mul by a non-negative scale, clamp and truncation all are); NaN products (inf * 0) go to bin 0.
__device__ __forceinline__ int sel_bin(float x, float lo, float scale, int nb) {
    float t = __fmul_rn(__fsub_rn(x, lo), scale);
    t = fminf(fmaxf(t, 0.f), (float)(nb - 1));
    return __float2int_rz(t);
}

__device__ __forceinline__ const float *pick_col(int a, const float *x, const float *y, const float *z) {
    return a == 0 ? x : (a == 1 ? y : z);
}

// =====================================================================================
// Level setup: flatten Cell[] (the reference's wire format) into the SoA level state.
// Mirrors what ServiceCopyCells prepares per level (copyCells.cu:29-61) without block descriptors.
// =====================================================================================
// trial cuts of a cell's first pass: the implicit bisection tree below (L,R), heap order; counters cleared
__device__ __forceinline__ void init_first_cuts(const LevelState &lv, uint32_t c, float L, float R, int nc) {
    float cv[kCS];
    float lo[kCS], hi[kCS];
    lo[0] = L; hi[0] = R;
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        if (k < nc) {
            cv[k] = mid_cut(lo[k], hi[k]);
            if (2 * k + 2 < 7) { lo[2 * k + 1] = lo[k]; hi[2 * k + 1] = cv[k]; lo[2 * k + 2] = cv[k]; hi[2 * k + 2] = hi[k]; }
        } else cv[k] = 0.f;
    }
    cv[7] = 0.f;
#pragma unroll
    for (int k = 0; k < kCS; ++k) { lv.cuts[c * kCS + k] = cv[k]; lv.cnt_l[c * kCS + k] = 0u; }
}

__global__ void k_level_setup(const orb_cell *__restrict__ cells, uint32_t nCells, const uint32_t *__restrict__ range,
                              const uint32_t *__restrict__ total_by_id, LevelState lv, uint32_t nLocal, int nc,
                              int *__restrict__ err, uint32_t *__restrict__ n_active0) {
    pdl_enter();
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c == 0 && n_active0) *n_active0 = 1u;      // any non-zero value opens the gate of the level's first count pass
    if (c >= nCells) return;
    orb_cell cell = cells[c];
    uint32_t b = range[2 * cell.id], e = range[2 * cell.id + 1];
    lv.bnd[c] = b;
    if (c == 0 && b != 0) atomicExch(err, ORB_ERR_RANGE);
    if (c + 1 == nCells) {
        lv.bnd[nCells] = e;
        if (e != nLocal) atomicExch(err, ORB_ERR_RANGE);
    } else {
        uint32_t nb = range[2 * cells[c + 1].id];
        if (nb != e) atomicExch(err, ORB_ERR_RANGE);
    }
    if (e < b) atomicExch(err, ORB_ERR_RANGE);
    if (cell.cutAxis < 0 || cell.cutAxis > 2) atomicExch(err, ORB_ERR_ARG);
    lv.axis[c] = cell.cutAxis < 0 ? 0 : (cell.cutAxis > 2 ? 2 : cell.cutAxis);
    float L = cell.cutMarginLeft, R = cell.cutMarginRight;
    lv.mL[c] = L;
    lv.mR[c] = R;
    lv.total[c] = total_by_id ? total_by_id[cell.id] : 0u;
    lv.nleaf[c] = cell.nLeafCells;
    uint32_t fnd = cell.foundCut ? 1u : 0u;
    lv.found[c] = fnd;
    lv.active[c] = fnd ? 0u : 1u;
    lv.iter[c] = 0;
    lv.nleft_g[c] = 0;
    lv.nleft_l[c] = 0;
    init_first_cuts(lv, c, L, R, nc);
}

// ServiceCount on the device (count.cpp:16): local size of every cell of the level from the range map
__global__ void k_cell_sizes(const orb_cell *__restrict__ cells, uint32_t nCells, const uint32_t *__restrict__ range,
                             uint32_t *__restrict__ out) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < nCells) out[c] = range[2 * cells[c].id + 1] - range[2 * cells[c].id];
}
// ... and the sizes summed over ranks remembered per cell id (the totals the bisection's target derives from)
__global__ void k_store_totals(const orb_cell *__restrict__ cells, uint32_t nCells, const uint32_t *__restrict__ tot,
                               uint32_t *__restrict__ total_by_id) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < nCells) total_by_id[cells[c].id] = tot[c];
}

// first cell whose range extends beyond the start of each map tile
__global__ void k_tile_map(const uint32_t *__restrict__ bnd, uint32_t nCells, uint32_t nMapTiles,
                           uint32_t *__restrict__ tile_first, uint32_t *__restrict__ zero, size_t nZero) {
    pdl_enter();
    // also clears the selection search's histogram rows of the level (saves a memset between two kernels)
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nZero; i += (size_t)gridDim.x * blockDim.x) zero[i] = 0u;
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nMapTiles) return;
    uint32_t start = t * (uint32_t)kMapTile;
    // smallest c with bnd[c+1] > start
    uint32_t lo = 0, hi = nCells - 1;
    while (lo < hi) {
        uint32_t m = (lo + hi) >> 1;
        if (bnd[m + 1] > start) hi = m; else lo = m + 1;
    }
    tile_first[t] = lo;
}

// =====================================================================================
// Count-left: replaces reduce3/reduce (countLeftGPUAxis.cu:133-186, countLeftGPU.cu:21-78)
// with the CPU comparison `x < cut` (countLeft.cpp:35).
//
// Two kernels, chosen per level by the average cell size (see launch_count in orb_capi.cu):
//   k_count_stream  cells much larger than a tile: persistent blocks stream 16 KB tiles, counters persist in
//                   registers across tiles of one cell, warp REDUX -> shared -> ONE global atomic per block per
//                   (cell, cut);
//   k_count_cells   small cells: one block or one warp owns a cell and stores its counts (no atomics).
//
// NC = 7 (three bisection steps per pass): the seven cuts form a sorted binary tree, so each particle
// is binned with a 3-compare descent (3 FSETP + 4 FSEL) and the bin is added to two packed 4x8-bit
// accumulators; 12 instructions per particle instead of 7 compare+add pairs.  Bins are turned into the
// per-cut cumulative counts `#{x < cut_k}` at flush time.  Sorted order of the heap-ordered cuts:
// c3 <= c1 <= c4 <= c0 <= c5 <= c2 <= c6 (float midpoints are monotone, so the order is non-strict but
// never violated; equal cuts only leave bins empty).
// =====================================================================================
// heap node -> number of sorted bins at or below it (cumulative index): node k gets bins [0 .. kSortedRank[k]]
__device__ __constant__ int kSortedRank7[7] = {3, 1, 5, 0, 2, 4, 6};

// 0xFFFFFFFF if a < b else 0, computed as the sign of (a - b): one FADD (FMA pipe) + one arithmetic shift,
// instead of FSETP + SEL (two ALU-pipe ops; the ALU pipe is the count kernel's bottleneck).
// Exact for all finite / infinite a, b with a != -0.0: IEEE subtraction never rounds a non-zero difference to a zero
// of the wrong sign, and a == b gives +0.  Callers canonicalise -0.0 to +0.0 first (canon0).  NaN coordinates are
// not supported (the CPU `<` yields false, the sign of a NaN difference is unspecified).
__device__ __forceinline__ int lt_mask(float a, float b) {
    return __float_as_int(__fsub_rn(a, b)) >> 31;
}
// -0.0 -> +0.0, everything else unchanged (x + (+0) in round-to-nearest)
__device__ __forceinline__ float canon0(float x) { return __fadd_rn(x, 0.0f); }
// m ? a : b for m in {-1, 0}, on the FMA pipe: b + m * (b - a); `bma` = b - a precomputed per cell
__device__ __forceinline__ int sel_mask(int m, int bma, int b) {
    int r;
    asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(r) : "r"(m), "r"(bma), "r"(b));
    return r;
}

// Cut set of one cell prepared for the 3-step descent: bit patterns and select deltas (heap order c0..c6).
struct Cuts7 {
    float c0;
    int c2b, d21;        // level 2: m0 ? c1 : c2
    int c4b, d43;        // level 3 (left subtree):  m1 ? c3 : c4
    int c6b, d65;        // level 3 (right subtree): m1 ? c5 : c6
    __device__ __forceinline__ void set(const float *cv) {
        c0 = cv[0];
        const int c1b = __float_as_int(cv[1]);
        c2b = __float_as_int(cv[2]);
        const int c3b = __float_as_int(cv[3]);
        c4b = __float_as_int(cv[4]);
        const int c5b = __float_as_int(cv[5]);
        c6b = __float_as_int(cv[6]);
        d21 = c2b - c1b; d43 = c4b - c3b; d65 = c6b - c5b;
    }
};

// One particle through the sorted cut tree: 3 float compares (FSET masks), selects as integer multiply-adds on the
// bit patterns (exact), packed 8-bit bin increment.  Sorted bin s = 7 + 4*m0 + 2*m1 + m2 (masks are -1 / 0).
__device__ __forceinline__ void bin7(float xin, const Cuts7 &k, unsigned &lo, unsigned &hi) {
    const float x = canon0(xin);
    const int m0 = lt_mask(x, k.c0);
    const int ab = sel_mask(m0, k.d21, k.c2b);
    const int m1 = lt_mask(x, __int_as_float(ab));
    const int tb = sel_mask(m1, k.d43, k.c4b);
    const int ub = sel_mask(m1, k.d65, k.c6b);
    const int bb = sel_mask(m0, ub - tb, ub);
    const int m2 = lt_mask(x, __int_as_float(bb));
    const int sh = sel_mask(m1, 16, sel_mask(m2, 8, 24));   // 8 * (s & 3) = 24 + 16*m1 + 8*m2
    const unsigned inc = 1u << sh;
    lo += inc & (unsigned)m0;
    hi += inc & ~(unsigned)m0;
}

template <int NC>
__device__ __forceinline__ void count_vals(const float (&v)[4], const bool (&in)[4], const float (&cv)[NC], unsigned (&cnt)[NC],
                                           unsigned &lo, unsigned &hi) {
    if constexpr (NC == 7) {
        Cuts7 k;
        k.set(cv);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            unsigned l2 = 0u, h2 = 0u;
            bin7(v[j], k, l2, h2);
            lo += in[j] ? l2 : 0u;
            hi += in[j] ? h2 : 0u;
        }
    } else {
#pragma unroll
        for (int k = 0; k < NC; ++k)
#pragma unroll
            for (int j = 0; j < 4; ++j) cnt[k] += (in[j] && v[j] < cv[k]);
    }
}

template <int NC>
__device__ __forceinline__ void count_f4(const float4 q, const float (&cv)[NC], unsigned (&cnt)[NC], unsigned &lo, unsigned &hi) {
    const float v[4] = {q.x, q.y, q.z, q.w};
    const bool in[4] = {true, true, true, true};
    count_vals<NC>(v, in, cv, cnt, lo, hi);
}

// NC==7: fold the packed 8-bit bins into the eight 32-bit bin counters (call at least every 255 particles per thread)
__device__ __forceinline__ void unpack_bins(unsigned &lo, unsigned &hi, unsigned (&bin)[8]) {
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        bin[s] += (lo >> (8 * s)) & 0xffu;
        bin[4 + s] += (hi >> (8 * s)) & 0xffu;
    }
    lo = 0u;
    hi = 0u;
}

struct PassCtl {
    uint32_t *n_active;        // device: [maxPasses+2] active cells after pass p (index p+1); [0] = before first pass
    uint32_t *done;            // device: [maxPasses+2] block tickets
    volatile uint32_t *h_status;   // pinned host: [maxPasses+2] n_active+1 after pass p (0 = not yet known)
    unsigned long long *active_particles;   // device: [0] local particles streamed (cells active in a pass, per HBM pass)
                                            //         [1] the same weighted by bisection iterations consumed (reference-equivalent)
    int32_t *level_iters;      // device: max iterations over cells (the reference's j)
};


// One cell's bisection decisions for the pass just counted (orbit.cpp:191-231): up to M steps over the counted
// trial-cut tree, literal float arithmetic of the reference.  g0/g1 = the eight global counters of the cell.
// Returns 1 if the cell stays active.  npart / nipart / it feed the pass statistics.
// baseMode: 0 = counters restart from zero; 1 = this was the first pass of a level that will compact: record the
// bracket, clear the base; 2 = candidates are in place: counters restart from the cell's base.
template <int M>
__device__ __forceinline__ uint32_t update_cell(const LevelState &lv, uint32_t c, uint4 g0, uint4 g1,
                                                unsigned long long &npart, unsigned long long &nipart, int &it,
                                                int baseMode = 0) {
    constexpr int NC = (1 << M) - 1;
    const uint4 l0 = __ldcg(reinterpret_cast<const uint4 *>(lv.cnt_l + c * kCS)), l1 = __ldcg(reinterpret_cast<const uint4 *>(lv.cnt_l + c * kCS + 4));
    const float4 q0 = *reinterpret_cast<const float4 *>(lv.cuts + c * kCS), q1 = *reinterpret_cast<const float4 *>(lv.cuts + c * kCS + 4);
    const uint32_t cg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const uint32_t cl8[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
    const float cu[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
    float L = lv.mL[c], R = lv.mR[c];
    it = lv.iter[c];
    const int it0 = it;
    const uint32_t total = lv.total[c];
    const int nleaf = lv.nleaf[c];
    const float ratio = (float)(ceil(nleaf / 2.0) / nleaf);          // orbit.cpp:204
    const float prod = __fmul_rn(__uint2float_rn(total), ratio);      // oCounts[i] * ratio
    npart = (unsigned long long)(lv.bnd[c + 1] - lv.bnd[c]);
    int node = 0;
    bool fnd = false, movedL = false, movedR = false;
#pragma unroll
    for (int s = 0; s < M; ++s) {
        float cut = 0.f;
        uint32_t cnt = 0, cntl = 0;
#pragma unroll
        for (int k = 0; k < NC; ++k)
            if (k == node) { cut = cu[k]; cnt = cg[k]; cntl = cl8[k]; }
        const int diff = __float2int_rz(__fsub_rn(__uint2float_rn(cnt), prod));   // orbit.cpp:205
        ++it;
        if (abs(diff) < 3) {                                                      // orbit.cpp:208
            fnd = true;
            lv.nleft_g[c] = cnt;
            lv.nleft_l[c] = cntl;
            break;
        } else if (diff > 0) { R = cut; movedR = true; node = 2 * node + 1; }     // orbit.cpp:219
        else { L = cut; movedL = true; node = 2 * node + 2; }                     // orbit.cpp:227
        if (it >= kMaxIter) break;                                                // orbit.cpp:149
    }
    lv.mL[c] = L; lv.mR[c] = R; lv.iter[c] = it;
    if (baseMode == 1) {
        lv.compL[c] = movedL ? L : __int_as_float(0xff800000);
        lv.compR[c] = movedR ? R : __int_as_float(0x7f800000);
        lv.base_l[c] = 0u;
    }
    nipart = npart * (unsigned long long)(it - it0);
    uint32_t still = 0;
    if (fnd) { lv.found[c] = 1u; lv.active[c] = 0u; }
    else if (it >= kMaxIter) { lv.active[c] = 0u; }
    else {
        still = 1;
        float cv[8], lo[7], hi[7];
        lo[0] = L; hi[0] = R;
#pragma unroll
        for (int k = 0; k < 8; ++k) cv[k] = 0.f;
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            cv[k] = mid_cut(lo[k], hi[k]);
            if (2 * k + 2 < NC) { lo[2 * k + 1] = lo[k]; hi[2 * k + 1] = cv[k]; lo[2 * k + 2] = cv[k]; hi[2 * k + 2] = hi[k]; }
        }
        *reinterpret_cast<float4 *>(lv.cuts + c * kCS) = make_float4(cv[0], cv[1], cv[2], cv[3]);
        *reinterpret_cast<float4 *>(lv.cuts + c * kCS + 4) = make_float4(cv[4], cv[5], cv[6], cv[7]);
    }
    const uint32_t b0 = (baseMode == 2) ? __ldcg(lv.base_l + c) : 0u;
    *reinterpret_cast<uint4 *>(lv.cnt_l + c * kCS) = make_uint4(b0, b0, b0, b0);
    *reinterpret_cast<uint4 *>(lv.cnt_l + c * kCS + 4) = make_uint4(b0, b0, b0, b0);
    return still;
}

// Single-rank fusion of the update into the count kernel: the block that finishes last (ticket) runs the
// decisions for every cell of the level, so a pass is ONE launch.  Used for levels of up to kFuseMaxCells cells.
constexpr uint32_t kFuseMaxCells = 4096;
struct FuseCtl {
    int enabled;        // 0: separate k_update launch
    int M;              // bisection steps per pass
    int pass;
    int baseMode;       // see update_cell
    uint32_t *tickets;  // [pass slots] blocks finished (zeroed per build)
    PassCtl ctl;
};

template <int M>
__device__ __forceinline__ void fused_update_tail(const LevelState &lv, uint32_t nCells, const FuseCtl &fc) {
    __shared__ int s_last;
    __shared__ uint32_t s_n;
    __shared__ unsigned long long s_p, s_q;
    __shared__ int s_it;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = (atomicAdd(&fc.tickets[fc.pass], 1u) == gridDim.x - 1) ? 1 : 0;
        s_n = 0; s_p = 0ull; s_q = 0ull; s_it = 0;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    uint32_t still = 0;
    unsigned long long npS = 0, nipS = 0;
    int itMax = 0;
    for (uint32_t c = threadIdx.x; c < nCells; c += blockDim.x) {
        if (!lv.active[c]) continue;
        const uint4 g0 = __ldcg(reinterpret_cast<const uint4 *>(lv.cnt_l + c * kCS)), g1 = __ldcg(reinterpret_cast<const uint4 *>(lv.cnt_l + c * kCS + 4));
        unsigned long long np = 0, nip = 0;
        int it = 0;
        still += update_cell<M>(lv, c, g0, g1, np, nip, it, fc.baseMode);
        npS += np; nipS += nip; itMax = max(itMax, it);
    }
    still = __reduce_add_sync(0xffffffffu, still);
    itMax = __reduce_max_sync(0xffffffffu, itMax);
    for (int o = 16; o; o >>= 1) {
        npS += __shfl_xor_sync(0xffffffffu, npS, o);
        nipS += __shfl_xor_sync(0xffffffffu, nipS, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (still) atomicAdd(&s_n, still);
        if (npS) atomicAdd(&s_p, npS);
        if (nipS) atomicAdd(&s_q, nipS);
        if (itMax) atomicMax(&s_it, itMax);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_p) atomicAdd(fc.ctl.active_particles, s_p);
        if (s_q) atomicAdd(fc.ctl.active_particles + 1, s_q);
        if (s_it) atomicMax(fc.ctl.level_iters, s_it);
        fc.ctl.n_active[fc.pass + 1] = s_n;
        fc.ctl.h_status[fc.pass] = s_n + 1u;
    }
}

__device__ __forceinline__ void fused_update_dispatch(const LevelState &lv, uint32_t nCells, const FuseCtl &fc) {
    if (!fc.enabled) return;
    if (fc.M == 3) fused_update_tail<3>(lv, nCells, fc);
    else if (fc.M == 2) fused_update_tail<2>(lv, nCells, fc);
    else fused_update_tail<1>(lv, nCells, fc);
}

// ---- asynchronous tile loads (LDGSTS): global -> shared without staging in registers ----
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// ---- shared pieces of the count kernels -------------------------------------------------------------
// per-thread accumulators: NC < 7 -> one counter per cut; NC == 7 -> eight sorted bins (+ packed lo/hi)
template <int NC>
struct Acc {
    static constexpr int NB = (NC == 7) ? 8 : NC;
    unsigned a[NB];
    unsigned lo, hi;
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int k = 0; k < NB; ++k) a[k] = 0u;
        lo = hi = 0u;
    }
    __device__ __forceinline__ void add_f4(const float4 q, const float (&cv)[NC]) {
        const float v[4] = {q.x, q.y, q.z, q.w};
        const bool in[4] = {true, true, true, true};
        unsigned dummy[NC];
        if constexpr (NC == 7) count_vals<NC>(v, in, cv, dummy, lo, hi);
        else count_vals<NC>(v, in, cv, a, lo, hi);
    }
    // NC == 7 fast path with the cut set already prepared (no per-call setup)
    __device__ __forceinline__ void add_f4(const float4 q, const Cuts7 &k) {
        bin7(q.x, k, lo, hi); bin7(q.y, k, lo, hi); bin7(q.z, k, lo, hi); bin7(q.w, k, lo, hi);
    }
    __device__ __forceinline__ void add_masked(const float (&v)[4], const bool (&in)[4], const float (&cv)[NC]) {
        unsigned dummy[NC];
        if constexpr (NC == 7) count_vals<NC>(v, in, cv, dummy, lo, hi);
        else count_vals<NC>(v, in, cv, a, lo, hi);
    }
    // NC == 7: call at least every 255 particles per thread
    __device__ __forceinline__ void fold() {
        if constexpr (NC == 7) unpack_bins(lo, hi, a);
    }
};

// cumulative count `#{x < cut_k}` for heap node k from the eight sorted bins
__device__ __forceinline__ unsigned cum_from_bins(const unsigned *bins, int k) {
    unsigned v = 0u;
    const int r = kSortedRank7[k];
    for (int i = 0; i <= r; ++i) v += bins[i];
    return v;
}

// Fragmented tile [t0,t1): several cells.  The tile's cell table (bounds, axis, active flag, trial cuts of up to
// kCountCellsSmem cells) is staged in shared memory once - one memory latency for the whole tile instead of a chain of
// dependent loads per warp segment - then every warp bins its 128-particle segments against the table and adds into
// shared per-cell accumulators; one global atomic per (cell, cut) per tile.  Block-uniform call (contains barriers).
// Tiles with more cells than the table are processed in several rounds.
template <int NC>
struct FragSmem {
    uint32_t beg[kCountCellsSmem + 1];
    uint32_t act[kCountCellsSmem];
    int ax[kCountCellsSmem];
    float cuts[kCountCellsSmem][kCS];
};

template <int NC>
__device__ __forceinline__ void count_fragmented_tile(const float *__restrict__ x, const float *__restrict__ y,
                                                      const float *__restrict__ z, const LevelState &lv, uint32_t nCells,
                                                      uint32_t cT, uint32_t t0, uint32_t t1, uint32_t *s_cell) {
    __shared__ FragSmem<NC> fs;
    __shared__ uint32_t s_nc;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t cFirst = cT, r0 = t0;               // round: cells cFirst.., particles from r0
    while (r0 < t1) {
        // ---- stage the table: thread i owns cell cFirst + i ----
        {
            const uint32_t cc = cFirst + tid;
            bool valid = false;
            uint32_t b = 0;
            if (tid <= kCountCellsSmem && cc <= nCells) {
                b = lv.bnd[min(cc, nCells)];
                valid = (tid == 0) || (b < t1);
            }
            if (tid <= kCountCellsSmem) fs.beg[tid] = valid ? b : 0xffffffffu;   // entry n = end of cell n-1; beyond the tile: +inf
            if (tid < kCountCellsSmem && valid && cc < nCells) {
                fs.act[tid] = __ldcg(&lv.active[cc]);
                fs.ax[tid] = lv.axis[cc];
                const float4 *cp = reinterpret_cast<const float4 *>(lv.cuts + cc * kCS);
                *reinterpret_cast<float4 *>(&fs.cuts[tid][0]) = __ldcg(cp);
                *reinterpret_cast<float4 *>(&fs.cuts[tid][4]) = __ldcg(cp + 1);
            }
            const int nv = __syncthreads_count(tid < kCountCellsSmem && valid && cc < nCells);
            if (tid == 0) s_nc = (uint32_t)nv;
        }
        for (int i = tid; i < kCountCellsSmem * NC; i += kThreads) s_cell[i] = 0u;
        __syncthreads();
        const uint32_t nc = s_nc;                                  // cells in the table (>= 1)
        // the table covers particles up to the begin of the first cell that did not fit (or the tile end)
        uint32_t r1 = t1;
        if (nc == (uint32_t)kCountCellsSmem) r1 = min(t1, fs.beg[kCountCellsSmem] == 0xffffffffu ? t1 : fs.beg[kCountCellsSmem]);
        // segments stay 128-aligned relative to the tile (vector loads need 16-byte alignment); particles before r0
        // belong to cells of an earlier round and are masked out by the cell bounds
        const uint32_t segBase = t0 + ((r0 - t0) / 128u) * 128u;
        for (uint32_t seg = segBase + warp * 128u; seg < r1; seg += kWarps * 128u) {
            const uint32_t segEnd = min(seg + 128u, r1);
            uint32_t j = 0;
            while (j + 1 < nc && fs.beg[j + 1] <= seg) ++j;
            const uint32_t e0 = seg + lane * 4u;
            for (; j < nc; ++j) {
                const uint32_t b = fs.beg[j], e = min(fs.beg[j + 1], r1);
                if (b >= segEnd) break;
                const uint32_t lo_e = max(b, seg), hi_e = min(e, segEnd);
                if (hi_e > lo_e && fs.act[j]) {
                    const float *cl = pick_col(fs.ax[j], x, y, z);
                    float v[4];
                    bool in[4];
                    if (e0 >= lo_e && e0 + 4u <= hi_e) {
                        const float4 q = __ldg(reinterpret_cast<const float4 *>(cl + e0));
                        v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
                        in[0] = in[1] = in[2] = in[3] = true;
                    } else {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint32_t ee = e0 + k;
                            in[k] = (ee >= lo_e && ee < hi_e);
                            v[k] = in[k] ? __ldg(cl + ee) : 0.f;
                        }
                    }
                    float ccv[NC];
#pragma unroll
                    for (int k = 0; k < NC; ++k) ccv[k] = fs.cuts[j][k];
                    unsigned n[NC];
#pragma unroll
                    for (int k = 0; k < NC; ++k) n[k] = 0u;
                    if constexpr (NC == 7) {
                        // <=128 particles per warp segment: packed 8-bit bins survive the warp sum
                        unsigned plo = 0u, phi = 0u;
                        count_vals<NC>(v, in, ccv, n, plo, phi);
                        plo = __reduce_add_sync(0xffffffffu, plo);
                        phi = __reduce_add_sync(0xffffffffu, phi);
                        unsigned run = 0u, cum[8];
#pragma unroll
                        for (int sIdx = 0; sIdx < 4; ++sIdx) { run += (plo >> (8 * sIdx)) & 0xffu; cum[sIdx] = run; }
#pragma unroll
                        for (int sIdx = 0; sIdx < 4; ++sIdx) { run += (phi >> (8 * sIdx)) & 0xffu; cum[4 + sIdx] = run; }
                        n[0] = cum[3]; n[1] = cum[1]; n[2] = cum[5]; n[3] = cum[0]; n[4] = cum[2]; n[5] = cum[4]; n[6] = cum[6];
                    } else {
                        unsigned dlo = 0u, dhi = 0u;
                        count_vals<NC>(v, in, ccv, n, dlo, dhi);
#pragma unroll
                        for (int k = 0; k < NC; ++k) n[k] = __reduce_add_sync(0xffffffffu, n[k]);
                    }
                    if (lane == 0) {
#pragma unroll
                        for (int k = 0; k < NC; ++k)
                            if (n[k]) atomicAdd(&s_cell[j * NC + k], n[k]);
                    }
                }
                if (e > segEnd) break;
            }
        }
        __syncthreads();
        for (int i = tid; i < (int)nc * NC; i += kThreads) {
            const unsigned v = s_cell[i];
            if (v) atomicAdd(&lv.cnt_l[(cFirst + i / NC) * kCS + (i % NC)], v);
        }
        __syncthreads();
        r0 = r1;
        cFirst += nc;
    }
}

// ---- regime A: cells much larger than a tile.  Persistent blocks, block b owns a contiguous range of tiles. ----
// Phase 1: every thread classifies one of the block's tiles (cell, stream / skip / fragmented) and caches the cell's
// cuts in shared memory - two memory latencies for the whole block instead of a dependent chain per tile.
// Phase 2: the streamable tiles run through a per-thread cp.async ring; counters persist across tiles of one cell.
//
// MODE (byte-reducing search; only with NC == 7):
//   kCountFull     every pass reads the whole cut-axis column of the active cells;
//   kCountCompact  second pass of a level: particles inside the bracket [compL, compR) fixed by the first pass are
//                  binned AND written, per warp, to the warp's 512-slot region of the idle ping-pong column
//                  (`cand`), the number kept per (tile, warp) goes to tile_ncand; particles below compL are only
//                  counted (they are left of every later cut) and summed into base_l;
//   kCountCand     later passes read just those candidates: counts = base_l + #{cand < cut}.
// Tiles that contain a cell boundary are never compacted: they are recounted in full every pass.
constexpr int kMaxUnits = 64;    // tiles classified per round
constexpr int kCountStages = 3;  // depth of the per-thread cp.async ring (tiles in flight per block: kCountStages - 1)
constexpr int kCountFull = 0, kCountCompact = 1, kCountCand = 2;
constexpr int kWarpSlots = kCountTile / kWarps;   // candidate slots per (tile, warp)

__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// static shared memory of one streaming count pass (one instance per kernel)
template <int NC>
struct CountSmem {
    uint32_t acc[Acc<NC>::NB + 1];                        // [NB] = particles below compL (compact mode)
    uint32_t cell[kCountCellsSmem * NC];
    uint32_t uTile[kMaxUnits], uCell[kMaxUnits];          // compacted streamable tiles
    int uAx[kMaxUnits];
    alignas(16) float uCuts[kMaxUnits][kCS];              // their cells' trial cuts (no global load on a cell change)
    float uBr[kMaxUnits][2];                              // compact mode: bracket of the unit's cell
    alignas(16) uint32_t uN[kMaxUnits][kWarps];           // cand mode: candidates per warp region
    uint32_t fTile[kMaxUnits], fCell[kMaxUnits];          // fragmented tiles
    uint32_t wS[kWarps], wF[kWarps];
};

// One streaming count pass over this block's tiles (block-uniform call; ends with all counters flushed).
template <int NC, int MODE>
__device__ __forceinline__ void stream_count_pass(const float *__restrict__ x, const float *__restrict__ y,
                                                  const float *__restrict__ z, float *__restrict__ cand,
                                                  const LevelState &lv, const uint32_t *__restrict__ tile_first,
                                                  uint32_t nCells, uint32_t nLocal, uint32_t nTiles, float4 *ring,
                                                  CountSmem<NC> &sm, unsigned long long *stamps = nullptr) {
    constexpr int NB = Acc<NC>::NB;
    uint32_t *s_acc = sm.acc, *s_cell = sm.cell, *s_uTile = sm.uTile, *s_uCell = sm.uCell, *s_fTile = sm.fTile, *s_fCell = sm.fCell;
    uint32_t *s_wS = sm.wS, *s_wF = sm.wF;
    int *s_uAx = sm.uAx;
    float(*s_uCuts)[kCS] = sm.uCuts;
    float(*s_uBr)[2] = sm.uBr;
    uint32_t(*s_uN)[kWarps] = sm.uN;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __syncthreads();
    if (tid <= NB) s_acc[tid] = 0u;

    Acc<NC> acc;
    acc.clear();
    unsigned below = 0u;
    float cv[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) cv[k] = 0.f;
    Cuts7 k7;
    k7.c0 = 0.f; k7.c2b = k7.d21 = k7.c4b = k7.d43 = k7.c6b = k7.d65 = 0;
    float brL = 0.f, brR = 0.f;
    int cur = -1;

    // One global atomic per block per (cell, cut): warp REDUX -> shared -> global.  Atomics to the few counter
    // rows of a level all land in one or two L2 slices (~0.65 ns each, serialised), so their number must stay small:
    // a per-warp flush cost 232K atomics per pass and 150 us at 8 cells (profiles/r01_count_notes.txt).
    auto flush = [&]() {   // block-uniform
        if (cur < 0) return;
        acc.fold();
#pragma unroll
        for (int k = 0; k < NB; ++k) {
            const unsigned v = __reduce_add_sync(0xffffffffu, acc.a[k]);
            if (lane == 0 && v) atomicAdd(&s_acc[k], v);
            acc.a[k] = 0u;
        }
        if (MODE == kCountCompact) {
            const unsigned v = __reduce_add_sync(0xffffffffu, below);
            if (lane == 0 && v) atomicAdd(&s_acc[NB], v);
            below = 0u;
        }
        __syncthreads();
        if (tid < NC) {
            unsigned v;
            if constexpr (NC == 7) v = cum_from_bins(s_acc, tid);
            else v = s_acc[tid];
            if (MODE == kCountCompact) v += s_acc[NB];            // particles below the bracket are left of every cut
            if (v) atomicAdd(&lv.cnt_l[(uint32_t)cur * kCS + tid], v);
        }
        if (MODE == kCountCompact && tid == NC && s_acc[NB]) atomicAdd(&lv.base_l[cur], s_acc[NB]);
        __syncthreads();
        if (tid <= NB) s_acc[tid] = 0u;
        __syncthreads();
    };

    // Block b owns the contiguous tiles [tb0, tb1): its tiles share one or two cells, so counters are flushed
    // once or twice per pass instead of once per tile.
    const uint32_t tilesPerBlock = (nTiles + gridDim.x - 1) / gridDim.x;
    const uint32_t tb0 = min(blockIdx.x * tilesPerBlock, nTiles), tb1 = min(tb0 + tilesPerBlock, nTiles);

    for (uint32_t base = tb0; base < tb1; base += (uint32_t)kMaxUnits) {
        // ---- phase 1: classify up to kMaxUnits tiles of this block ----
        const uint32_t t = base + (uint32_t)tid;
        int kind = 0;   // 0 none/skip, 1 stream, 2 fragmented
        uint32_t c = 0;
        int ax = 0;
        if (tid < kMaxUnits && t < tb1) {
            const uint32_t t0 = t * (uint32_t)kCountTile, t1 = min(t0 + (uint32_t)kCountTile, nLocal);
            c = tile_first[t * (kCountTile / kMapTile)];
            const uint32_t cb = lv.bnd[c], ce = lv.bnd[c + 1];
            ax = lv.axis[c];
            if (cb <= t0 && ce >= t1 && (t1 - t0) == (uint32_t)kCountTile) kind = __ldcg(&lv.active[c]) ? 1 : 0;
            else kind = 2;
        }
        const unsigned mS = __ballot_sync(0xffffffffu, kind == 1), mF = __ballot_sync(0xffffffffu, kind == 2);
        if (lane == 0) { s_wS[warp] = __popc(mS); s_wF[warp] = __popc(mF); }
        __syncthreads();
        uint32_t offS = 0, offF = 0, nS = 0, nF = 0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
            if (w < warp) { offS += s_wS[w]; offF += s_wF[w]; }
            nS += s_wS[w]; nF += s_wF[w];
        }
        const unsigned ltMask = (1u << lane) - 1u;
        if (kind == 1) {
            const uint32_t r = offS + __popc(mS & ltMask);
            s_uTile[r] = t; s_uCell[r] = c; s_uAx[r] = ax;
            const float4 *cp = reinterpret_cast<const float4 *>(lv.cuts + c * kCS);
            // per-pass mutable state is read past L1 (the persistent level kernel re-reads it every pass)
            *reinterpret_cast<float4 *>(&s_uCuts[r][0]) = __ldcg(cp);
            *reinterpret_cast<float4 *>(&s_uCuts[r][4]) = __ldcg(cp + 1);
            if (MODE == kCountCompact) { s_uBr[r][0] = __ldcg(&lv.compL[c]); s_uBr[r][1] = __ldcg(&lv.compR[c]); }
            if (MODE == kCountCand) {
                const uint4 *np = reinterpret_cast<const uint4 *>(lv.tile_ncand + (size_t)t * kWarps);
                *reinterpret_cast<uint4 *>(&s_uN[r][0]) = __ldcg(np);
                *reinterpret_cast<uint4 *>(&s_uN[r][4]) = __ldcg(np + 1);
            }
        }
        if (kind == 2) { const uint32_t r = offF + __popc(mF & ltMask); s_fTile[r] = t; s_fCell[r] = c; }
        __syncthreads();
        if (stamps && tid == 0 && base == tb0) stamps[1] = gtimer();   // classified

        auto enter_cell = [&](uint32_t k) {   // block-uniform: unit k starts a new cell
            const uint32_t cK = s_uCell[k];
            if ((int)cK == cur) return;
            flush();
            cur = (int)cK;
#pragma unroll
            for (int j = 0; j < NC; ++j) cv[j] = s_uCuts[k][j];
            if constexpr (NC == 7) k7.set(cv);
            if (MODE == kCountCompact) { brL = s_uBr[k][0]; brR = s_uBr[k][1]; }
        };

        if (nS && MODE != kCountCand) {
            // ---- phase 2a: streamable tiles through a per-thread cp.async ring: every thread copies the four 16-byte
            //      pieces it will count itself into its own shared-memory slots, kCountStages-1 tiles ahead, so the
            //      wait is per thread (cp.async.wait_group) and needs no barrier. ----
            auto issue = [&](uint32_t k) {
                const float4 *p = reinterpret_cast<const float4 *>(pick_col(s_uAx[k], x, y, z) + s_uTile[k] * (uint32_t)kCountTile) + tid;
                float4 *dst = ring + (k % kCountStages) * (4 * kThreads) + tid;
#pragma unroll
                for (int j = 0; j < 4; ++j) cp_async16(dst + j * kThreads, p + j * kThreads);
            };
#pragma unroll
            for (int pre = 0; pre < kCountStages - 1; ++pre) {
                if ((uint32_t)pre < nS) issue(pre);
                cp_async_commit();
            }
            for (uint32_t k = 0; k < nS; ++k) {
                if (k + kCountStages - 1 < nS) issue(k + kCountStages - 1);
                cp_async_commit();
                cp_async_wait<kCountStages - 1>();     // tile k has landed in my slots
                enter_cell(k);
                const float4 *src = ring + (k % kCountStages) * (4 * kThreads) + tid;
                const float4 q0 = src[0], q1 = src[kThreads], q2 = src[2 * kThreads], q3 = src[3 * kThreads];
                if constexpr (MODE == kCountCompact && NC == 7) {
                    // bin + keep the particles inside [brL, brR); count the ones below brL
                    const float v[16] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, q3.x, q3.y, q3.z, q3.w};
                    unsigned keep = 0u;
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float xv = canon0(v[j]);
                        const int mLo = lt_mask(xv, brL), mHi = lt_mask(xv, brR);
                        const unsigned in = (unsigned)(mHi & ~mLo);
                        below += (unsigned)mLo & 1u;
                        unsigned l2 = 0u, h2 = 0u;
                        bin7(v[j], k7, l2, h2);
                        acc.lo += l2 & in;
                        acc.hi += h2 & in;
                        keep |= (in & 1u) << j;
                    }
                    // warp-private compaction: exclusive scan of the per-thread keep counts, then each kept particle
                    // goes to the warp's region of this tile in `cand` (order is irrelevant for counting)
                    const unsigned mine = __popc(keep);
                    unsigned incl = mine;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const unsigned up = __shfl_up_sync(0xffffffffu, incl, o);
                        if (lane >= o) incl += up;
                    }
                    const unsigned wtot = __shfl_sync(0xffffffffu, incl, 31);
                    float *dstc = cand + (size_t)s_uTile[k] * kCountTile + warp * kWarpSlots + (incl - mine);
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (keep & (1u << j)) dstc[__popc(keep & ((1u << j) - 1u))] = v[j];
                    if (lane == 0) lv.tile_ncand[(size_t)s_uTile[k] * kWarps + warp] = wtot;
                } else if constexpr (NC == 7) {
                    acc.add_f4(q0, k7); acc.add_f4(q1, k7); acc.add_f4(q2, k7); acc.add_f4(q3, k7);
                } else {
                    acc.add_f4(q0, cv); acc.add_f4(q1, cv); acc.add_f4(q2, cv); acc.add_f4(q3, cv);
                }
                if ((k & 7u) == 7u) acc.fold();   // 16 particles per tile per thread: fold before 255
            }
            cp_async_wait<0>();
            acc.fold();
        }
        if (nS && MODE == kCountCand) {
            // ---- phase 2a': candidates only.  Warp w reads the slots it filled in the compaction pass; loads of up to
            //      eight tiles are issued together so the pass pays one memory latency per eight tiles. ----
            if constexpr (NC == 7) {
                for (uint32_t k0 = 0; k0 < nS; k0 += 8u) {
                    float v0[8], v1[8];
                    unsigned n[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        n[u] = (k0 + u < nS) ? s_uN[k0 + u][warp] : 0u;
                        const float *src = cand + (size_t)((k0 + u < nS) ? s_uTile[k0 + u] : 0u) * kCountTile + warp * kWarpSlots;
                        v0[u] = ((unsigned)lane < n[u]) ? __ldcg(src + lane) : 0.f;
                        v1[u] = ((unsigned)lane + 32u < n[u]) ? __ldcg(src + lane + 32) : 0.f;
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        if (k0 + u >= nS) break;               // block-uniform
                        enter_cell(k0 + u);
                        unsigned l2 = 0u, h2 = 0u;
                        bin7(v0[u], k7, l2, h2);
                        if ((unsigned)lane < n[u]) { acc.lo += l2; acc.hi += h2; }
                        l2 = h2 = 0u;
                        bin7(v1[u], k7, l2, h2);
                        if ((unsigned)lane + 32u < n[u]) { acc.lo += l2; acc.hi += h2; }
                        if (n[u] > 64u) {                      // warp-uniform: rare long tail (clustered data)
                            const float *src = cand + (size_t)s_uTile[k0 + u] * kCountTile + warp * kWarpSlots;
                            for (unsigned i = 64u + lane; i < n[u]; i += 32u) {
                                l2 = h2 = 0u;
                                bin7(__ldcg(src + i), k7, l2, h2);
                                acc.lo += l2; acc.hi += h2;
                            }
                        }
                        acc.fold();                            // <= 2 + 14 particles per thread per tile
                    }
                }
            }
        }
        // ---- phase 2b: fragmented tiles (cell boundaries, array tail) are always counted in full ----
        if (nF) {
            flush();
            cur = -1;
            for (uint32_t k = 0; k < nF; ++k) {
                const uint32_t tt = s_fTile[k];
                const uint32_t t0 = tt * (uint32_t)kCountTile, t1 = min(t0 + (uint32_t)kCountTile, nLocal);
                count_fragmented_tile<NC>(x, y, z, lv, nCells, s_fCell[k], t0, t1, s_cell);
            }
        }
        __syncthreads();
    }
    if (stamps && tid == 0) stamps[2] = gtimer();   // streamed, before the last flush
    flush();
}


template <int NC, int MODE>
__global__ void __launch_bounds__(kThreads, 4) k_count_stream(const float *__restrict__ x, const float *__restrict__ y,
                                                              const float *__restrict__ z, float *__restrict__ cand,
                                                              LevelState lv, const uint32_t *__restrict__ tile_first,
                                                              uint32_t nCells, uint32_t nLocal, uint32_t nTiles,
                                                              const uint32_t *__restrict__ gate, FuseCtl fc) {
    if (gate && *gate == 0u) {   // speculative pass after convergence: nothing to count
        if (fc.enabled && blockIdx.x == 0 && threadIdx.x == 0) { fc.ctl.n_active[fc.pass + 1] = 0u; fc.ctl.h_status[fc.pass] = 1u; }
        return;
    }
    extern __shared__ __align__(16) unsigned char count_smem[];       // kCountStages x 16 KB ring
    __shared__ CountSmem<NC> sm;
    stream_count_pass<NC, MODE>(x, y, z, cand, lv, tile_first, nCells, nLocal, nTiles, reinterpret_cast<float4 *>(count_smem), sm);
    fused_update_dispatch(lv, nCells, fc);
}

// ---- N2: host-free level loop.  ONE cooperative launch runs the whole bisection of a level (orbit.cpp:146-232):
// every pass = streaming count (full / compact / candidates) -> grid barrier -> bisection update spread over the
// blocks -> grid barrier; the loop ends on the device when no cell is active (or after 32 iterations).  Cells that hit
// the iteration cap get their extra count at the final cut in the same launch.  No kernel launch, no host polling and
// no speculative pass per bisection pass.  Single rank, levels of up to kPersistMaxCells cells in the streaming regime.
constexpr uint32_t kPersistMaxCells = 8192;

// Lightweight grid barrier for co-resident blocks (cooperative launch): one monotonic counter, barrier number `gen`
// is complete when the counter reaches (gen + 1) * gridDim.x.  One atomic and one polling thread per block - cheaper
// than cooperative_groups::grid_group::sync() for two barriers per bisection pass.
__device__ __forceinline__ void grid_barrier(unsigned int *counter, unsigned int &gen) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        const unsigned int target = (gen + 1u) * gridDim.x;
        while (*((volatile unsigned int *)counter) < target) {}
        __threadfence();
    }
    ++gen;
    __syncthreads();
}
struct LevelCtl {
    uint32_t *n_active;                     // [maxPasses + 2] cells active after pass p at [p + 1]; zeroed per level
    unsigned long long *active_particles;   // [2] statistics (see PassCtl)
    int32_t *level_iters;
    int32_t *passes_out;                    // passes executed in this level
    uint32_t *n_unfound_out;                // cells that hit the iteration cap
    unsigned int *barrier;                  // grid barrier counter, zeroed before the launch
    int compaction;                         // 1: full, compact, then candidate passes
    unsigned long long *dbg;                // optional (ORB_DEBUG_TIMES): globaltimer stamps of block 0, 5 per pass
    unsigned long long *dbg_blocks;         // optional (ORB_DEBUG_TIMES=2): [pass][block][4] start, classified, streamed, flushed
    const uint32_t *gate;                   // optional: the launch does nothing when *gate == 0 (fallback of the selection search)
    const uint32_t *only;                   // optional [nCells]: with `gate`, the cells this launch is responsible for
};

template <int M>
__global__ void __launch_bounds__(kThreads, 4) k_level_persistent(const float *__restrict__ x, const float *__restrict__ y,
                                                                  const float *__restrict__ z, float *__restrict__ cand,
                                                                  LevelState lv, const uint32_t *__restrict__ tile_first,
                                                                  uint32_t nCells, uint32_t nLocal, uint32_t nTiles,
                                                                  LevelCtl lc) {
    constexpr int NC = (1 << M) - 1;
    constexpr int maxPasses = (kMaxIter + M - 1) / M;
    extern __shared__ __align__(16) unsigned char count_smem[];
    __shared__ CountSmem<NC> sm;
    __shared__ uint32_t s_n;
    __shared__ unsigned long long s_p, s_q;
    __shared__ int s_it;
    float4 *ring = reinterpret_cast<float4 *>(count_smem);
    unsigned int gen = 0;
    const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x, gsize = gridDim.x * blockDim.x;
    if (lc.gate && __ldcg(lc.gate) == 0u) return;   // grid-uniform

    int pass = 0;
    for (; pass < maxPasses; ++pass) {
        const int mode = !lc.compaction ? kCountFull : (pass == 0 ? kCountFull : (pass == 1 ? kCountCompact : kCountCand));
        const int baseMode = !lc.compaction ? 0 : (pass == 0 ? 1 : 2);
        const bool stamp = lc.dbg && blockIdx.x == 0 && threadIdx.x == 0;
        if (stamp) lc.dbg[pass * 5 + 0] = gtimer();
        unsigned long long *bs = lc.dbg_blocks ? lc.dbg_blocks + ((size_t)pass * gridDim.x + blockIdx.x) * 4 : nullptr;
        if (bs && threadIdx.x == 0) bs[0] = gtimer();
        if constexpr (NC == 7) {
            if (mode == kCountCompact) stream_count_pass<NC, kCountCompact>(x, y, z, cand, lv, tile_first, nCells, nLocal, nTiles, ring, sm, bs);
            else if (mode == kCountCand) stream_count_pass<NC, kCountCand>(x, y, z, cand, lv, tile_first, nCells, nLocal, nTiles, ring, sm, bs);
            else stream_count_pass<NC, kCountFull>(x, y, z, cand, lv, tile_first, nCells, nLocal, nTiles, ring, sm, bs);
        } else {
            stream_count_pass<NC, kCountFull>(x, y, z, cand, lv, tile_first, nCells, nLocal, nTiles, ring, sm, bs);
        }
        if (bs && threadIdx.x == 0) bs[3] = gtimer();
        if (stamp) lc.dbg[pass * 5 + 1] = gtimer();
        grid_barrier(lc.barrier, gen);
        if (stamp) lc.dbg[pass * 5 + 2] = gtimer();
        // ---- bisection update, one thread per cell across the whole grid ----
        if (threadIdx.x == 0) { s_n = 0; s_p = 0ull; s_q = 0ull; s_it = 0; }
        __syncthreads();
        uint32_t still = 0;
        unsigned long long npS = 0, nipS = 0;
        int itMax = 0;
        for (uint32_t c = gtid; c < nCells; c += gsize) {
            if (!__ldcg(&lv.active[c])) continue;
            const uint4 g0 = __ldcg(reinterpret_cast<const uint4 *>(lv.cnt_l + c * kCS)), g1 = __ldcg(reinterpret_cast<const uint4 *>(lv.cnt_l + c * kCS + 4));
            unsigned long long np = 0, nip = 0;
            int it = 0;
            still += update_cell<M>(lv, c, g0, g1, np, nip, it, baseMode);
            npS += np; nipS += nip; itMax = max(itMax, it);
        }
        still = __reduce_add_sync(0xffffffffu, still);
        itMax = __reduce_max_sync(0xffffffffu, itMax);
        for (int o = 16; o; o >>= 1) {
            npS += __shfl_xor_sync(0xffffffffu, npS, o);
            nipS += __shfl_xor_sync(0xffffffffu, nipS, o);
        }
        if ((threadIdx.x & 31) == 0) {
            if (still) atomicAdd(&s_n, still);
            if (npS) atomicAdd(&s_p, npS);
            if (nipS) atomicAdd(&s_q, nipS);
            if (itMax) atomicMax(&s_it, itMax);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            if (s_n) atomicAdd(&lc.n_active[pass + 1], s_n);
            if (s_p) atomicAdd(lc.active_particles, s_p);
            if (s_q) atomicAdd(lc.active_particles + 1, s_q);
            if (s_it) atomicMax(lc.level_iters, s_it);
        }
        if (stamp) lc.dbg[pass * 5 + 3] = gtimer();
        grid_barrier(lc.barrier, gen);
        if (stamp) lc.dbg[pass * 5 + 4] = gtimer();
        if (__ldcg(&lc.n_active[pass + 1]) == 0u) { ++pass; break; }
    }
    // ---- cells that hit the 32-iteration cap: one count at getCut() of their last margins (never counted before) ----
    if (pass >= maxPasses) {
        uint32_t need = 0;
        for (uint32_t c = gtid; c < nCells; c += gsize) {
            if (lc.only && !lc.only[c]) continue;   // finished by the selection search (its capped cells are counted)
            const uint32_t nf = lv.found[c] ? 0u : 1u;
            lv.active[c] = nf;
            if (nf) {
                lv.cuts[c * kCS] = mid_cut(lv.mL[c], lv.mR[c]);
                *reinterpret_cast<uint4 *>(lv.cnt_l + c * kCS) = make_uint4(0u, 0u, 0u, 0u);
                *reinterpret_cast<uint4 *>(lv.cnt_l + c * kCS + 4) = make_uint4(0u, 0u, 0u, 0u);
                ++need;
            }
        }
        need = __reduce_add_sync(0xffffffffu, need);
        if ((threadIdx.x & 31) == 0 && need) atomicAdd(lc.n_unfound_out, need);
        grid_barrier(lc.barrier, gen);
        if (__ldcg(lc.n_unfound_out) != 0u) {
            stream_count_pass<NC, kCountFull>(x, y, z, cand, lv, tile_first, nCells, nLocal, nTiles, ring, sm);
            grid_barrier(lc.barrier, gen);
            for (uint32_t c = gtid; c < nCells; c += gsize)
                if (__ldcg(&lv.active[c])) {
                    const uint32_t v = __ldcg(lv.cnt_l + c * kCS);
                    lv.nleft_g[c] = v;
                    lv.nleft_l[c] = v;
                    lv.active[c] = 0u;
                }
        }
    }
    if (gtid == 0) atomicAdd(lc.passes_out, pass);
}

// ---- regime B: cells of at most a few tiles.  One group of G threads per cell (G = 256: block, G = 32: warp);
// the group owns the cell, so the result is a plain store (no atomics) and inactive cells cost one load. ----
template <int NC, int G>
__global__ void __launch_bounds__(kThreads, 4) k_count_cells(const float *__restrict__ x, const float *__restrict__ y,
                                                             const float *__restrict__ z, LevelState lv, uint32_t nCells,
                                                             const uint32_t *__restrict__ gate, FuseCtl fc) {
    if (gate && *gate == 0u) {
        if (fc.enabled && blockIdx.x == 0 && threadIdx.x == 0) { fc.ctl.n_active[fc.pass + 1] = 0u; fc.ctl.h_status[fc.pass] = 1u; }
        return;
    }
    constexpr int NB = Acc<NC>::NB;
    constexpr int GPB = kThreads / G;                 // groups per block
    __shared__ uint32_t s_acc[GPB][kWarps][NB];       // only used when G == 256
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gtid = tid % G;                         // thread index inside the group
    const uint32_t group = blockIdx.x * GPB + tid / G, nGroups = gridDim.x * GPB;

    for (uint32_t c = group; c < nCells; c += nGroups) {
        const uint32_t b = lv.bnd[c], e = lv.bnd[c + 1];
        const uint32_t act = lv.active[c];
        const int ax = lv.axis[c];
        float cv[NC];
#pragma unroll
        for (int k = 0; k < NC; ++k) cv[k] = lv.cuts[c * kCS + k];
        if (!act) continue;                            // group-uniform
        Acc<NC> acc;
        acc.clear();
        Cuts7 k7;
        if constexpr (NC == 7) k7.set(cv);
        else { k7.c0 = 0.f; k7.c2b = k7.d21 = k7.c4b = k7.d43 = k7.c6b = k7.d65 = 0; }
        if (e > b) {
            const float *col = pick_col(ax, x, y, z);
            const uint32_t a0 = min((b + 3u) & ~3u, e);    // first 16-byte aligned particle
            const uint32_t a1 = max(a0, e & ~3u);          // end of the aligned body
            // head [b,a0) and tail [a1,e): at most 3 + 3 particles
            if (gtid < 8) {
                float v[4] = {0.f, 0.f, 0.f, 0.f};
                bool in[4] = {false, false, false, false};
                const uint32_t ee = (gtid < 4) ? b + gtid : a1 + (gtid - 4);
                const bool ok = (gtid < 4) ? (ee < a0) : (ee < e);
                in[0] = ok;
                v[0] = ok ? __ldg(col + ee) : 0.f;
                acc.add_masked(v, in, cv);
            }
            // body: G threads stride over float4; 4 loads in flight per thread
            const float4 *p = reinterpret_cast<const float4 *>(col + a0);
            const uint32_t n4 = (a1 - a0) >> 2;
            uint32_t i = gtid;
            int sinceFold = 1;
            for (; i + 3u * G < n4; i += 4u * G) {
                const float4 q0 = __ldg(p + i), q1 = __ldg(p + i + G), q2 = __ldg(p + i + 2 * G), q3 = __ldg(p + i + 3 * G);
                if constexpr (NC == 7) { acc.add_f4(q0, k7); acc.add_f4(q1, k7); acc.add_f4(q2, k7); acc.add_f4(q3, k7); }
                else { acc.add_f4(q0, cv); acc.add_f4(q1, cv); acc.add_f4(q2, cv); acc.add_f4(q3, cv); }
                sinceFold += 16;
                if (sinceFold > 224) { acc.fold(); sinceFold = 0; }
            }
            for (; i < n4; i += G) {
                if constexpr (NC == 7) acc.add_f4(__ldg(p + i), k7);
                else acc.add_f4(__ldg(p + i), cv);
                sinceFold += 4;
                if (sinceFold > 224) { acc.fold(); sinceFold = 0; }
            }
        }
        acc.fold();
        // reduce over the group and store
        unsigned r[NB];
#pragma unroll
        for (int k = 0; k < NB; ++k) r[k] = __reduce_add_sync(0xffffffffu, acc.a[k]);
        if constexpr (G == 32) {
            if (lane < NC) {
                unsigned v;
                if constexpr (NC == 7) v = cum_from_bins(r, lane);
                else { v = 0u;
#pragma unroll
                    for (int k = 0; k < NC; ++k) if (k == lane) v = r[k]; }
                lv.cnt_l[c * kCS + lane] = v;
            }
        } else {
            __syncthreads();   // previous cell's readers are done with s_acc
            if (lane == 0) {
#pragma unroll
                for (int k = 0; k < NB; ++k) s_acc[0][warp][k] = r[k];
            }
            __syncthreads();
            if (tid < NC) {
                unsigned bins[NB];
#pragma unroll
                for (int k = 0; k < NB; ++k) {
                    unsigned v = 0u;
#pragma unroll
                    for (int w = 0; w < kWarps; ++w) v += s_acc[0][w][k];
                    bins[k] = v;
                }
                unsigned v;
                if constexpr (NC == 7) v = cum_from_bins(bins, tid);
                else { v = 0u;
#pragma unroll
                    for (int k = 0; k < NC; ++k) if (k == tid) v = bins[k]; }
                lv.cnt_l[c * kCS + tid] = v;
            }
        }
    }
    fused_update_dispatch(lv, nCells, fc);
}

// =====================================================================================
// Bisection update: orbit.cpp:191-231 replayed on the device, one thread per cell.
// M steps per pass over the counted trial-cut tree; literal float arithmetic of the reference:
//   float ratio = ceil(nLeafCells/2.0)/nLeafCells;  int difference = countLeft - count*ratio;
// =====================================================================================
template <int M>
__global__ void __launch_bounds__(kThreads) k_update(LevelState lv, uint32_t nCells, int pass, PassCtl ctl, PeerSet ps, int baseMode) {
    constexpr int NC = (1 << M) - 1;
    const uint32_t gate = ctl.n_active[pass];
    const uint32_t cT = blockIdx.x * blockDim.x + threadIdx.x;   // this thread's cell
    if (ps.n && gate != 0u) {
        // ---- phase A: push this block's local count rows to every peer, then raise this block's flag there ----
        if (cT < nCells) {
            const uint4 a0 = *reinterpret_cast<const uint4 *>(lv.cnt_l + cT * kCS), a1 = *reinterpret_cast<const uint4 *>(lv.cnt_l + cT * kCS + 4);
            for (int r = 0; r < ps.n; ++r) {
                if (r == ps.self) continue;
                uint4 *dst = reinterpret_cast<uint4 *>(peer_rows(ps, r, ps.self) + cT * kCS);
                dst[0] = a0;
                dst[1] = a1;
            }
        }
        __threadfence_system();
        __syncthreads();
        if ((int)threadIdx.x < ps.n && (int)threadIdx.x != ps.self)
            *((volatile uint32_t *)peer_flag(ps, threadIdx.x, ps.self, blockIdx.x)) = ps.seq;
        // ---- phase B: wait for block `blockIdx.x` of every peer ----
        if ((int)threadIdx.x < ps.n && (int)threadIdx.x != ps.self) {
            volatile uint32_t *f = peer_flag(ps, ps.self, threadIdx.x, blockIdx.x);
            while ((int32_t)(*f - ps.seq) < 0) {}
            __threadfence_system();
        }
        __syncthreads();
    }
    __shared__ uint32_t s_n;
    __shared__ unsigned long long s_p, s_q;
    __shared__ int s_it;
    if (threadIdx.x == 0) { s_n = 0; s_p = 0ull; s_q = 0ull; s_it = 0; }
    __syncthreads();
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t still = 0;
    unsigned long long npart = 0, nipart = 0;
    int it = 0;
    if (gate != 0u && c < nCells && lv.active[c]) {
        uint4 g0 = *reinterpret_cast<const uint4 *>(lv.cnt_g + c * kCS), g1 = *reinterpret_cast<const uint4 *>(lv.cnt_g + c * kCS + 4);
        if (ps.n) {   // ---- phase C: sum over ranks (own row from cnt_l, the peers' rows have arrived in recv) ----
            g0 = *reinterpret_cast<const uint4 *>(lv.cnt_l + c * kCS);
            g1 = *reinterpret_cast<const uint4 *>(lv.cnt_l + c * kCS + 4);
            for (int r = 0; r < ps.n; ++r) {
                if (r == ps.self) continue;
                const uint4 *src = reinterpret_cast<const uint4 *>(peer_rows(ps, ps.self, r) + c * kCS);
                const uint4 b0 = __ldcg(src), b1 = __ldcg(src + 1);
                g0.x += b0.x; g0.y += b0.y; g0.z += b0.z; g0.w += b0.w;
                g1.x += b1.x; g1.y += b1.y; g1.z += b1.z; g1.w += b1.w;
            }
        }
        still = update_cell<M>(lv, c, g0, g1, npart, nipart, it, baseMode);
    }
    // block -> grid reduction of (cells still active, particles streamed this pass, max iterations)
    uint32_t wn = __reduce_add_sync(0xffffffffu, still);
    int wit = __reduce_max_sync(0xffffffffu, it);
    for (int o = 16; o; o >>= 1) {
        npart += __shfl_xor_sync(0xffffffffu, npart, o);
        nipart += __shfl_xor_sync(0xffffffffu, nipart, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (wn) atomicAdd(&s_n, wn);
        if (npart) atomicAdd(&s_p, npart);
        if (nipart) atomicAdd(&s_q, nipart);
        if (wit) atomicMax(&s_it, wit);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_p) atomicAdd(ctl.active_particles, s_p);
        if (s_q) atomicAdd(ctl.active_particles + 1, s_q);
        if (s_it) atomicMax(ctl.level_iters, s_it);
        uint32_t n = s_n;
        bool last = true;
        if (gridDim.x > 1) {
            if (n) atomicAdd(&ctl.n_active[pass + 1], n);
            __threadfence();
            last = atomicAdd(&ctl.done[pass], 1u) == gridDim.x - 1;
            if (last) {
                __threadfence();
                n = *((volatile uint32_t *)&ctl.n_active[pass + 1]);
            }
        } else {
            ctl.n_active[pass + 1] = n;
        }
        // mapped pinned memory: the host polls this word, no stream sync; the store drains at the latest when the
        // kernel retires, so no system-scope fence is spent on it
        if (last) ctl.h_status[pass] = n + 1u;
    }
}

// Cells that hit the 32-iteration cap are cut at getCut() of their last margins — a position that was
// never counted.  One extra single-cut pass gives the partition its offsets.
__global__ void k_finalize_prepare(LevelState lv, uint32_t nCells, uint32_t *__restrict__ n_unfound) {
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nCells) return;
    uint32_t need = lv.found[c] ? 0u : 1u;
    lv.active[c] = need;
    if (need) {
        lv.cuts[c * kCS] = mid_cut(lv.mL[c], lv.mR[c]);
        lv.cnt_l[c * kCS] = 0u;
        atomicAdd(n_unfound, 1u);
    }
}
__global__ void k_finalize_apply(LevelState lv, uint32_t nCells) {
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nCells) return;
    if (lv.active[c]) {
        lv.nleft_g[c] = lv.cnt_g[c * kCS];
        lv.nleft_l[c] = lv.cnt_l[c * kCS];
        lv.active[c] = 0u;
    }
}

// write the level's bisection result back into the Cell array (what master() holds after orbit.cpp:232)
__global__ void k_writeback_cells(orb_cell *__restrict__ cells, LevelState lv, uint32_t nCells) {
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nCells) return;
    cells[c].cutMarginLeft = lv.mL[c];
    cells[c].cutMarginRight = lv.mR[c];
    cells[c].foundCut = lv.found[c] ? 1 : 0;
}

// =====================================================================================
// Split: orbit.cpp:235-250 + cell.h:78-126 (children boxes, longest geometric side, margins = box faces)
// plus the child ranges the partition will produce (partition.cpp:54-60) and child totals.
// =====================================================================================
__device__ __forceinline__ void child_axis_margins(orb_cell &ch) {
    int maxD = -1;
    float maxSize = 0.0f;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        float size = __fsub_rn(ch.upper[d], ch.lower[d]);
        if (size > maxSize) { maxSize = size; maxD = d; }    // strict '>' : lowest axis wins ties (cell.h:115)
    }
    ch.cutAxis = maxD;
    int a = maxD < 0 ? 0 : maxD;   // all-zero extents (cell.h would index [-1]); keep a defined value
    ch.cutMarginLeft = ch.lower[a];
    ch.cutMarginRight = ch.upper[a];
}

// `gate` (may be null): the kernel was enqueued before the host knew whether the level's search left cells to the
// iterative loop; a non-zero word means it did, and the launch does nothing (the host enqueues it again afterwards).
// `nx.enabled`: the same launch also prepares the NEXT level - what k_level_setup and k_tile_map would do from the
// children it has just made: their SoA level state (in the other LevelState buffer, the partition that follows still
// reads this level's), the tile -> first-cell map of the children's ranges, cleared histogram rows.  Saves two launches
// per level.
struct NextLevel {
    int enabled;
    LevelState lv;                // the other buffer
    int nc;                       // trial cuts per pass of the iterative search
    int *err;
    uint32_t *n_active0;          // gate of the next level's first count pass
    uint32_t nMapTiles;
    uint32_t *tile_first;         // the other buffer
    const uint32_t *tile_first_cur;   // this level's map: the parent of a tile's first particle without a search
    uint32_t *zero;               // histogram rows of the next level
    size_t nZero;
};
__device__ __forceinline__ void setup_child(const NextLevel &nx, uint32_t idx, const orb_cell &ch, uint32_t begin, uint32_t total) {
    nx.lv.bnd[idx] = begin;
    if (ch.cutAxis < 0 || ch.cutAxis > 2) atomicExch(nx.err, ORB_ERR_ARG);
    nx.lv.axis[idx] = ch.cutAxis < 0 ? 0 : (ch.cutAxis > 2 ? 2 : ch.cutAxis);
    nx.lv.mL[idx] = ch.cutMarginLeft;
    nx.lv.mR[idx] = ch.cutMarginRight;
    nx.lv.total[idx] = total;
    nx.lv.nleaf[idx] = ch.nLeafCells;
    nx.lv.found[idx] = 0u;
    nx.lv.active[idx] = 1u;
    nx.lv.iter[idx] = 0;
    nx.lv.nleft_g[idx] = 0u;
    nx.lv.nleft_l[idx] = 0u;
    init_first_cuts(nx.lv, idx, ch.cutMarginLeft, ch.cutMarginRight, nx.nc);
}
__global__ void k_split(orb_cell *__restrict__ heap, uint32_t first, uint32_t nCells, LevelState lv,
                        uint32_t *__restrict__ range, uint32_t *__restrict__ total_by_id, float *__restrict__ final_cut,
                        const uint32_t *__restrict__ gate, NextLevel nx) {
    pdl_enter();
    if (gate && *((volatile const uint32_t *)gate) != 0u) return;
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (nx.enabled) {
        for (size_t i = c; i < nx.nZero; i += (size_t)gridDim.x * blockDim.x) nx.zero[i] = 0u;
        if (c == 0 && nx.n_active0) *nx.n_active0 = 1u;
        if (c < nx.nMapTiles) {
            // smallest child whose range extends beyond the tile's first particle: its parent is this level's entry of
            // the same tile (one load instead of a binary search over the boundaries: log2(nCells) dependent loads
            // were most of this kernel's time at the deep levels), then left or right of the split position
            const uint32_t start = c * (uint32_t)kMapTile;
            uint32_t lo;
            if (nx.tile_first_cur) lo = nx.tile_first_cur[c];
            else {
                lo = 0;
                uint32_t hi = nCells - 1;
                while (lo < hi) {
                    const uint32_t m = (lo + hi) >> 1;
                    if (lv.bnd[m + 1] > start) hi = m; else lo = m + 1;
                }
            }
            nx.tile_first[c] = 2u * lo + ((lv.bnd[lo] + lv.nleft_l[lo] > start) ? 0u : 1u);
        }
    }
    if (c >= nCells) return;
    orb_cell p = heap[first + c];
    p.cutMarginLeft = lv.mL[c];
    p.cutMarginRight = lv.mR[c];
    p.foundCut = lv.found[c] ? 1 : 0;
    heap[first + c] = p;
    const float cut = mid_cut(p.cutMarginLeft, p.cutMarginRight);
    final_cut[c] = cut;
    const int nL = (int)ceil(p.nLeafCells / 2.0), nR = p.nLeafCells - nL;   // cell.h:79-80
    orb_cell l, r;
    l.id = (p.id + 1) * 2 - 1; r.id = (p.id + 1) * 2;
    l.nLeafCells = nL; r.nLeafCells = nR;
    l.prevCutAxis = r.prevCutAxis = p.cutAxis;
    l.foundCut = r.foundCut = 0;
    l.pad_[0] = l.pad_[1] = l.pad_[2] = r.pad_[0] = r.pad_[1] = r.pad_[2] = 0;
#pragma unroll
    for (int d = 0; d < 3; ++d) { l.lower[d] = r.lower[d] = p.lower[d]; l.upper[d] = r.upper[d] = p.upper[d]; }
    const int ax = lv.axis[c];
    l.upper[ax] = cut;
    r.lower[ax] = cut;
    child_axis_margins(l);
    child_axis_margins(r);
    heap[l.id] = l;
    heap[r.id] = r;
    const uint32_t b = lv.bnd[c], e = lv.bnd[c + 1], m = b + lv.nleft_l[c];
    range[2 * l.id] = b; range[2 * l.id + 1] = m;
    range[2 * r.id] = m; range[2 * r.id + 1] = e;
    total_by_id[l.id] = lv.nleft_g[c];
    total_by_id[r.id] = lv.total[c] - lv.nleft_g[c];
    if (nx.enabled) {
        setup_child(nx, 2u * c, l, b, lv.nleft_g[c]);
        setup_child(nx, 2u * c + 1u, r, m, lv.total[c] - lv.nleft_g[c]);
        if (c + 1u == nCells) nx.lv.bnd[2u * nCells] = e;
    }
}

// service-granular partition: only ranges + final cut (the host owns the Cell heap)
__global__ void k_ranges_from_level(const orb_cell *__restrict__ cells, uint32_t nCells, LevelState lv,
                                    uint32_t *__restrict__ range, float *__restrict__ final_cut) {
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nCells) return;
    const int id = cells[c].id;
    final_cut[c] = mid_cut(lv.mL[c], lv.mR[c]);
    const uint32_t b = lv.bnd[c], e = lv.bnd[c + 1], m = b + lv.nleft_l[c];
    const int lid = (id + 1) * 2 - 1, rid = (id + 1) * 2;
    range[2 * lid] = b; range[2 * lid + 1] = m;
    range[2 * rid] = m; range[2 * rid + 1] = e;
}

// =====================================================================================
// Partition: replaces partition<256>+permute<256> (partitionGPU.cu:58-280) and the CPU Hoare loop
// (partition.cpp:30-60) with a STABLE split by `x < cut` (canonical tie mode).
// One launch for all cells.  Tiles of kPartTile particles, staged in shared memory and written out in destination
// order; a tile needs a carry-in (number of left particles of its first cell in earlier tiles) only if that cell began
// before the tile.  k_partition_coop gets the carries by reduce-then-scan over contiguous per-block tile ranges,
// k_partition_cells (small cells) walks each cell with one block.
// =====================================================================================
// Left counts of the partition's per-block trailing segments, known BEFORE the partition runs (no phase-1 read of the
// cut-axis column): the search's last pass over the column (COMPACT) uses the same contiguous chunk per block as the
// partition, counts the chunk's particles below the candidate bins and remembers where the block's candidates went;
// once the cut is final, lefts = below + #{those candidates < cut}.  Blocks of k_sel_percell, which own a whole cell,
// record final numbers for the chunks that end inside their cell.  A record that does not carry the level's tag or
// names another cell is ignored (the block then counts the column itself, as before).
struct PreLeft {
    uint32_t tag;        // level tag (stale records of earlier levels never match)
    uint32_t cell;       // level-local index of the chunk's trailing cell
    uint32_t below;      // kind 0 / 2: particles of the segment below the candidate bins; kind 1: its left particles
    uint32_t gbase, tot; // kind 0: the segment's candidates are list[gbase .. gbase + tot) of the cell's list; kind 2: of the whole column
    uint32_t kind;
    uint32_t pad_[2];
};

// Histogram rows of the NEXT level's selection search, produced while the particles pass through the partition
// anyway: every particle is binned on its child's cut axis with the child's bin function (sel_bin over the child's
// margins) - the next level's HIST pass, 4 B per particle, is not needed.  k_split has prepared the children's level
// state (axis, margins) and cleared the rows before the partition starts.
struct NextHist {
    int enabled;
    uint32_t *hist;           // [2 * nCells][nb]
    int nb;                   // 512 or 1024 (2 * nb words alias PartSmem::sd, which only boundary tiles use)
    const float *mL, *mR;     // children's margins / axes (the next level's LevelState)
    const int32_t *axis;
};

// dynamic shared memory of the partition kernels
struct PartSmem {
    float raw[2][3][kPartTile];       // double-buffered x,y,z tile (48 KB)
    uint32_t sd[kPartTile];           // destination index of the particle staged at slot p
    uint16_t sperm[kPartTile];        // tile offset of the particle that goes to slot p (inverse permutation)
    uint32_t cbeg[kPartCells], cend[kPartCells], nleft[kPartCells], B[kPartCells], Bend[kPartCells];
    float cut[kPartCells];
    int axis[kPartCells];
    uint32_t warpTot[kWarps];
    uint32_t lbsum[kWarps];
    float chLo[2], chScale[2];        // NextHist: bin function of the running cell's left / right child
    int chAx[2];
    alignas(8) unsigned long long mbar[2];   // bulk-copy tile loads: one transaction barrier per buffer
};

// NextHist helpers (block-uniform calls).  The block histogram of the running cell's two children lives in sm.sd.
__device__ __forceinline__ void part_hist_zero(PartSmem &sm, const NextHist &nh) {
    for (int b = threadIdx.x; b < 2 * nh.nb; b += kThreads) sm.sd[b] = 0u;
}
__device__ __forceinline__ void part_hist_enter(PartSmem &sm, const NextHist &nh, uint32_t c) {
    if (threadIdx.x < 2) {
        const uint32_t ch = 2u * c + threadIdx.x;
        const float L = nh.mL[ch], R = nh.mR[ch];
        sm.chAx[threadIdx.x] = nh.axis[ch];
        sm.chLo[threadIdx.x] = L;
        sm.chScale[threadIdx.x] = sel_scale(L, R, nh.nb);
    }
    __syncthreads();
}
__device__ __forceinline__ void part_hist_flush(PartSmem &sm, const NextHist &nh, uint32_t c) {
    __syncthreads();
    for (int b = threadIdx.x; b < 2 * nh.nb; b += kThreads) {
        const uint32_t v = sm.sd[b];
        if (v) {
            atomicAdd(&nh.hist[(size_t)(2u * c + (b >= nh.nb ? 1u : 0u)) * nh.nb + (b & (nh.nb - 1))], v);
            sm.sd[b] = 0u;
        }
    }
    __syncthreads();
}

// issue the asynchronous copy of tile [T, T+kPartTile) of x,y,z into buffer `bsel` (6 x 16 B per thread)
__device__ __forceinline__ void part_prefetch(PartSmem &sm, int bsel, const float *__restrict__ x, const float *__restrict__ y,
                                              const float *__restrict__ z, uint32_t T) {
    const int tid = threadIdx.x;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const uint32_t o = (uint32_t)(h * kThreads + tid) * 4u;
        cp_async16(&sm.raw[bsel][0][o], x + T + o);
        cp_async16(&sm.raw[bsel][1][o], y + T + o);
        cp_async16(&sm.raw[bsel][2][o], z + T + o);
    }
}

// ---- the same tile load as ONE bulk asynchronous copy per column (cp.async.bulk, the TMA engine's linear mode):
//      a single thread arms the buffer's transaction barrier with the tile's byte count and issues three 8 KB
//      copies; every thread then waits on the barrier's phase.  No thread spends issue slots on LDGSTS. ----
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(a), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(a), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(void *smem, const void *gmem, unsigned bytes, unsigned long long *bar) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem), ba = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(sa), "l"(gmem), "r"(bytes), "r"(ba) : "memory");
}
__device__ __forceinline__ void part_prefetch_bulk(PartSmem &sm, int bsel, const float *__restrict__ x, const float *__restrict__ y,
                                                   const float *__restrict__ z, uint32_t T) {
    if (threadIdx.x == 0) {
        constexpr unsigned bytes = kPartTile * sizeof(float);
        mbar_expect_tx(&sm.mbar[bsel], 3u * bytes);
        bulk_load(&sm.raw[bsel][0][0], x + T, bytes, &sm.mbar[bsel]);
        bulk_load(&sm.raw[bsel][1][0], y + T, bytes, &sm.mbar[bsel]);
        bulk_load(&sm.raw[bsel][2][0], z + T, bytes, &sm.mbar[bsel]);
    }
}

// Table-driven tile body: particles of tile [T, T+kPartTile) that lie in [s0,s1) are split per cell, stable.
// The cell table (ncell entries, first entry = cell containing s0) is in shared memory; `carryIn` = left particles
// of the first cell that precede s0.  Returns the number of left particles of the segment that reaches s1.
__device__ __forceinline__ uint32_t part_tile_body(PartSmem &sm, int bsel, uint32_t T, uint32_t s0, uint32_t s1, int ncell,
                                                   uint32_t carryIn, float *__restrict__ x2, float *__restrict__ y2,
                                                   float *__restrict__ z2, const NextHist &nh, uint32_t cellBase) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // ---- cell of each of my 8 particles (two groups of 4 consecutive), left flag ----
    bool ok[8], fl[8];
    uint32_t jj[8];
#pragma unroll
    for (int g = 0; g < 2; ++g) {
        const uint32_t o = warp * 256u + g * 128u + lane * 4u;
        const uint32_t e = T + o;
        if (ncell == 1) {
            const float4 q = *reinterpret_cast<const float4 *>(&sm.raw[bsel][sm.axis[0]][o]);
            const float v[4] = {q.x, q.y, q.z, q.w};
            const float cutv = sm.cut[0];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                jj[4 * g + k] = 0u;
                ok[4 * g + k] = (e + k) >= s0 && (e + k) < s1;
                fl[4 * g + k] = ok[4 * g + k] && (v[k] < cutv);
            }
        } else {
            uint32_t j = 0;
            {   // largest j with cbeg[j] <= e
                uint32_t lo = 0, hi = (uint32_t)ncell - 1;
                while (lo < hi) {
                    const uint32_t m = (lo + hi + 1) >> 1;
                    if (sm.cbeg[m] <= e) lo = m; else hi = m - 1;
                }
                j = lo;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t ee = e + k;
                while (j + 1 < (uint32_t)ncell && sm.cbeg[j + 1] <= ee) ++j;
                jj[4 * g + k] = j;
                ok[4 * g + k] = ee >= s0 && ee < s1;
                const float v = sm.raw[bsel][sm.axis[j]][o + k];
                fl[4 * g + k] = ok[4 * g + k] && (v < sm.cut[j]);
            }
        }
    }
    if (nh.enabled) {   // boundary tiles are rare: the children's rows take these particles with global atomics
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            if (!ok[q]) continue;
            const uint32_t o = warp * 256u + (q >> 2) * 128u + lane * 4u + (q & 3);
            const uint32_t ch = 2u * (cellBase + jj[q]) + (fl[q] ? 0u : 1u);
            const float L = __ldg(nh.mL + ch), R = __ldg(nh.mR + ch);
            const float v = sm.raw[bsel][__ldg(nh.axis + ch)][o];
            atomicAdd(&nh.hist[(size_t)ch * nh.nb + sel_bin(v, L, sel_scale(L, R, nh.nb), nh.nb)], 1u);
        }
    }
    // ---- block exclusive scan of the left flags (order: warp region, group, lane, k) ----
    const uint32_t c0n = (uint32_t)fl[0] + fl[1] + fl[2] + fl[3];
    const uint32_t c1n = (uint32_t)fl[4] + fl[5] + fl[6] + fl[7];
    uint32_t i0 = c0n, i1 = c1n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t a = __shfl_up_sync(0xffffffffu, i0, o);
        const uint32_t b = __shfl_up_sync(0xffffffffu, i1, o);
        if (lane >= o) { i0 += a; i1 += b; }
    }
    const uint32_t tot0 = __shfl_sync(0xffffffffu, i0, 31), tot1 = __shfl_sync(0xffffffffu, i1, 31);
    if (lane == 0) sm.warpTot[warp] = tot0 + tot1;
    __syncthreads();
    uint32_t woff = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
        const uint32_t v = sm.warpTot[w];
        if (w < warp) woff += v;
        total += v;
    }
    uint32_t LE[8];
    {
        uint32_t r0 = woff + (i0 - c0n), r1 = woff + tot0 + (i1 - c1n);
#pragma unroll
        for (int k = 0; k < 4; ++k) { LE[k] = r0; r0 += fl[k]; LE[4 + k] = r1; r1 += fl[4 + k]; }
    }
    // ---- first / last particle of every segment record the running left count ----
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        if (!ok[q]) continue;
        const uint32_t ee = T + warp * 256u + (q >> 2) * 128u + lane * 4u + (q & 3);
        const uint32_t j = jj[q];
        if (ee == max(sm.cbeg[j], s0)) sm.B[j] = LE[q];
        if (ee + 1u == min(sm.cend[j], s1)) sm.Bend[j] = LE[q] + fl[q];
    }
    __syncthreads();

    const uint32_t lastSeg = total - sm.B[ncell - 1];   // left particles of the segment that reaches s1

    // ---- slot p of every particle (per segment: lefts, then rights) and its destination without carry ----
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        if (!ok[q]) continue;
        const uint32_t o = warp * 256u + (q >> 2) * 128u + lane * 4u + (q & 3);
        const uint32_t ee = T + o;
        const uint32_t j = jj[q];
        const uint32_t cbj = sm.cbeg[j];
        const uint32_t segStart = max(cbj, s0);
        const uint32_t lb = LE[q] - sm.B[j];                 // lefts of this cell before me, inside [s0,s1)
        uint32_t p, d;
        if (fl[q]) {
            p = (segStart - T) + lb;
            d = cbj + lb;
        } else {
            const uint32_t nLt = sm.Bend[j] - sm.B[j];
            const uint32_t rb = (ee - segStart) - lb;        // rights of this cell before me, inside [s0,s1)
            p = (segStart - T) + nLt + rb;
            d = cbj + sm.nleft[j] + ((ee - cbj) - lb);
        }
        sm.sperm[p] = (uint16_t)o;
        sm.sd[p] = d;
    }
    __syncthreads();

    const uint32_t carry0 = carryIn;
    // ---- coalesced runs out to the ping-pong columns; the first cell's destinations shift by the carry ----
    const uint32_t seg0len = min(sm.cend[0], s1) - s0;
    const uint32_t nLt0 = seg0len ? (sm.Bend[0] - sm.B[0]) : 0u;
    const uint32_t pBase = s0 - T;
    for (uint32_t p = pBase + tid; p < (s1 - T); p += kThreads) {
        uint32_t d = sm.sd[p];
        const uint32_t rel = p - pBase;
        if (rel < seg0len) d = (rel < nLt0) ? d + carry0 : d - carry0;
        const uint32_t o = sm.sperm[p];
        x2[d] = sm.raw[bsel][0][o];
        y2[d] = sm.raw[bsel][1][o];
        z2[d] = sm.raw[bsel][2][o];
    }
    __syncthreads();
    return lastSeg;
}

// build the cell table of a sub-range starting at s0 with first cell `cfirst`; returns ncell, sets s1
__device__ __forceinline__ int part_build_table(PartSmem &sm, const LevelState &lv, const float *__restrict__ final_cut,
                                                uint32_t nCells, uint32_t cfirst, uint32_t tEnd, uint32_t &s1) {
    const int tid = threadIdx.x;
    const uint32_t cidx = cfirst + tid;
    bool valid = false;
    uint32_t cb = 0, ce = 0;
    if (cidx < nCells) {
        cb = lv.bnd[cidx];
        valid = (tid == 0) || (cb < tEnd);
        if (valid) ce = lv.bnd[cidx + 1];
    }
    const int ncell = __syncthreads_count(valid);   // cells are consecutive, so valid is a prefix of the threads
    if (valid) {
        sm.cbeg[tid] = cb; sm.cend[tid] = ce;
        sm.nleft[tid] = lv.nleft_l[cidx];
        sm.cut[tid] = final_cut[cidx];
        sm.axis[tid] = lv.axis[cidx];
    }
    __syncthreads();
    s1 = tEnd;
    if (ncell == kPartCells) s1 = min(sm.cend[kPartCells - 1], tEnd);   // table full: rest in another sub-range
    return ncell;
}

// Lean body for a tile [T, tEnd) that lies inside ONE cell (the common case when cells are larger than a
// tile): all cell parameters are block-uniform registers, destinations are affine in the slot index, so only
// the inverse permutation goes through shared memory.  Returns the number of left particles in the tile.
__device__ __forceinline__ uint32_t part_tile_single(PartSmem &sm, int bsel, uint32_t oLo, uint32_t oHi, int axis, float cutv,
                                                     uint32_t baseL, uint32_t baseR, float *__restrict__ x2,
                                                     float *__restrict__ y2, float *__restrict__ z2, const NextHist &nh) {
    // valid particles are the tile offsets [oLo, oHi)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t nValid = oHi - oLo;
    bool fl[8];
    uint32_t o8[2];
#pragma unroll
    for (int g = 0; g < 2; ++g) {
        const uint32_t o = warp * 256u + g * 128u + lane * 4u;
        o8[g] = o;
        const float4 q = *reinterpret_cast<const float4 *>(&sm.raw[bsel][axis][o]);
        fl[4 * g + 0] = (o + 0u >= oLo) && (o + 0u < oHi) && (q.x < cutv);
        fl[4 * g + 1] = (o + 1u >= oLo) && (o + 1u < oHi) && (q.y < cutv);
        fl[4 * g + 2] = (o + 2u >= oLo) && (o + 2u < oHi) && (q.z < cutv);
        fl[4 * g + 3] = (o + 3u >= oLo) && (o + 3u < oHi) && (q.w < cutv);
    }
    unsigned flm = 0u;
#pragma unroll
    for (int q = 0; q < 8; ++q) flm |= (unsigned)fl[q] << q;
    if (nh.enabled) {   // bin every valid particle on its child's axis into the block histogram of the running cell's
                        // two children (measured: here, ahead of the scan, costs ~2 us per level less than after the stores)
        const int axL = sm.chAx[0], axR = sm.chAx[1];
        const float loL = sm.chLo[0], scL = sm.chScale[0], loR = sm.chLo[1], scR = sm.chScale[1];
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            const float4 a = *reinterpret_cast<const float4 *>(&sm.raw[bsel][axL][o8[g]]);
            const float4 b = *reinterpret_cast<const float4 *>(&sm.raw[bsel][axR][o8[g]]);
            const float va[4] = {a.x, a.y, a.z, a.w}, vb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t o = o8[g] + k;
                if (o >= oLo && o < oHi) {
                    const bool left = (flm >> (4 * g + k)) & 1u;
                    const int bin = sel_bin(left ? va[k] : vb[k], left ? loL : loR, left ? scL : scR, nh.nb);
                    atomicAdd(&sm.sd[(left ? 0 : nh.nb) + bin], 1u);
                }
            }
        }
    }
    const uint32_t c0n = (uint32_t)fl[0] + fl[1] + fl[2] + fl[3];
    const uint32_t c1n = (uint32_t)fl[4] + fl[5] + fl[6] + fl[7];
    uint32_t i0 = c0n, i1 = c1n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t a = __shfl_up_sync(0xffffffffu, i0, o);
        const uint32_t b = __shfl_up_sync(0xffffffffu, i1, o);
        if (lane >= o) { i0 += a; i1 += b; }
    }
    const uint32_t tot0 = __shfl_sync(0xffffffffu, i0, 31), tot1 = __shfl_sync(0xffffffffu, i1, 31);
    if (lane == 0) sm.warpTot[warp] = tot0 + tot1;
    __syncthreads();
    uint32_t woff = 0, nLt = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
        const uint32_t v = sm.warpTot[w];
        if (w < warp) woff += v;
        nLt += v;
    }
    // slot of each particle: lefts keep their order in [0,nLt), rights in [nLt,nValid)
    {
        uint32_t r0 = woff + (i0 - c0n), r1 = woff + tot0 + (i1 - c1n);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t oa = o8[0] + k, ob = o8[1] + k;
            if (oa >= oLo && oa < oHi) sm.sperm[fl[k] ? r0 : (nLt + (oa - oLo) - r0)] = (uint16_t)oa;
            r0 += fl[k];
            if (ob >= oLo && ob < oHi) sm.sperm[fl[4 + k] ? r1 : (nLt + (ob - oLo) - r1)] = (uint16_t)ob;
            r1 += fl[4 + k];
        }
    }
    __syncthreads();
    for (uint32_t p = tid; p < nValid; p += kThreads) {
        const uint32_t o = sm.sperm[p];
        const uint32_t d = (p < nLt) ? (baseL + p) : (baseR + (p - nLt));
        x2[d] = sm.raw[bsel][0][o];
        y2[d] = sm.raw[bsel][1][o];
        z2[d] = sm.raw[bsel][2][o];
    }
    __syncthreads();
    return nLt;
}

// Regime A: cells larger than a tile.  Reduce-then-scan in ONE cooperative launch.  Block b owns the contiguous
// tiles [tb0,tb1).  Phase 1 counts the left particles of the block's trailing segment (the cell that continues into
// the next block) - reads at most the cut-axis column of the block's range.  After one grid-wide barrier every block
// derives the carry-in of its first cell from the (at most gridDim) per-block records and then streams its tiles in
// order with the carry in a register: no tile ever waits on another block.
// (The first implementation was a single-pass decoupled look-back; ncu showed 35-42% of the stall samples on the
//  look-back wait and ~170 instructions per particle, see profiles/r01_partition_lookback_notes.txt.)
__global__ void __launch_bounds__(kThreads, 3) k_partition_coop(const float *__restrict__ x, const float *__restrict__ y,
                                                               const float *__restrict__ z, float *__restrict__ x2,
                                                               float *__restrict__ y2, float *__restrict__ z2,
                                                               LevelState lv, const float *__restrict__ final_cut,
                                                               const uint32_t *__restrict__ tile_first, uint32_t nCells,
                                                               uint32_t nLocal, uint32_t nTiles, uint32_t *blkLeft,
                                                               uint32_t *blkRestart, const uint32_t *__restrict__ gate, NextHist nh,
                                                               const PreLeft *__restrict__ pre, uint32_t preTag, uint32_t tilesPerBlockIn,
                                                               const float *__restrict__ preList, uint32_t preListStride,
                                                               int bulk /* tile loads by cp.async.bulk + mbarrier instead of per-thread cp.async */) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PartSmem &sm = *reinterpret_cast<PartSmem *>(smem_raw);
    if (gate && *((volatile const uint32_t *)gate) != 0u) return;     // see k_split; uniform over the grid
    unsigned mphase[2] = {0u, 0u};
    if (bulk) {
        if (threadIdx.x == 0) { mbar_init(&sm.mbar[0], 1u); mbar_init(&sm.mbar[1], 1u); mbar_fence_init(); }
        __syncthreads();
    }
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // (tilesPerBlockIn: the chunking the search's last pass used, so that `pre` describes this block's chunk)
    const uint32_t tilesPerBlock = tilesPerBlockIn ? tilesPerBlockIn : (nTiles + gridDim.x - 1) / gridDim.x;
    const uint32_t tb0 = min(blockIdx.x * tilesPerBlock, nTiles), tb1 = min(tb0 + tilesPerBlock, nTiles);
    const uint32_t chunkStart = tb0 * (uint32_t)kPartTile, chunkEnd = min(tb1 * (uint32_t)kPartTile, nLocal);

    // start the first tile's copy right away; it lands while phase 1 runs
    if (tb0 < tb1) {
        if (bulk) part_prefetch_bulk(sm, 0, x, y, z, chunkStart);
        else part_prefetch(sm, 0, x, y, z, chunkStart);
    }
    cp_async_commit();

    // ---------------- phase 1: left particles of the trailing segment ----------------
    {
        uint32_t cnt = 0, restart = 0;
        if (tb0 < tb1) {
            uint32_t c = tile_first[tb1 - 1];
            while (lv.bnd[c + 1] < chunkEnd) ++c;            // cell that contains particle chunkEnd-1
            const uint32_t cb = lv.bnd[c];
            restart = cb >= chunkStart ? 1u : 0u;
            const uint32_t b = max(cb, chunkStart);
            uint32_t e = chunkEnd;
            const float *col = pick_col(lv.axis[c], x, y, z);
            const float cutv = final_cut[c];
            if (pre) {      // block-uniform: the search already knows this segment's left count (or all but its candidates)
                const PreLeft P = pre[blockIdx.x];
                if (P.tag == preTag && P.cell == c) {
                    if (tid == 0) cnt = P.below;
                    if (P.kind == 0u || P.kind == 2u) {      // kind 2: gbase is an absolute index (the search block's private region)
                        const float *list = P.kind == 2u ? preList + P.gbase : preList + (preListStride ? (size_t)c * preListStride : (size_t)cb) + P.gbase;
                        uint32_t i = tid;
                        for (; i + 3u * kThreads < P.tot; i += 4u * kThreads) {      // four loads in flight per thread
                            const float a0 = __ldcg(list + i), a1 = __ldcg(list + i + kThreads), a2 = __ldcg(list + i + 2 * kThreads), a3 = __ldcg(list + i + 3 * kThreads);
                            cnt += (uint32_t)(a0 < cutv) + (uint32_t)(a1 < cutv) + (uint32_t)(a2 < cutv) + (uint32_t)(a3 < cutv);
                        }
                        for (; i < P.tot; i += kThreads) cnt += (__ldcg(list + i) < cutv) ? 1u : 0u;
                    }
                    e = b;      // nothing left to read
                }
            }
            const uint32_t a0 = min((b + 3u) & ~3u, e), a1 = max(a0, e & ~3u);
            if (tid < 8) {
                const uint32_t ee = (tid < 4) ? b + tid : a1 + (tid - 4);
                const bool ok = (tid < 4) ? (ee < a0) : (ee < e);
                if (ok) cnt += (__ldg(col + ee) < cutv);
            }
            const float4 *p = reinterpret_cast<const float4 *>(col + a0);
            const uint32_t n4 = (a1 - a0) >> 2;
            uint32_t i = tid;
            for (; i + 3u * kThreads < n4; i += 4u * kThreads) {
                const float4 q0 = __ldg(p + i), q1 = __ldg(p + i + kThreads), q2 = __ldg(p + i + 2 * kThreads), q3 = __ldg(p + i + 3 * kThreads);
                cnt += (q0.x < cutv) + (q0.y < cutv) + (q0.z < cutv) + (q0.w < cutv);
                cnt += (q1.x < cutv) + (q1.y < cutv) + (q1.z < cutv) + (q1.w < cutv);
                cnt += (q2.x < cutv) + (q2.y < cutv) + (q2.z < cutv) + (q2.w < cutv);
                cnt += (q3.x < cutv) + (q3.y < cutv) + (q3.z < cutv) + (q3.w < cutv);
            }
            for (; i < n4; i += kThreads) {
                const float4 q = __ldg(p + i);
                cnt += (q.x < cutv) + (q.y < cutv) + (q.z < cutv) + (q.w < cutv);
            }
        }
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        if (lane == 0) sm.warpTot[warp] = cnt;
        __syncthreads();
        if (tid == 0) {
            uint32_t tot = 0;
#pragma unroll
            for (int w = 0; w < kWarps; ++w) tot += sm.warpTot[w];
            blkLeft[blockIdx.x] = tot;
            blkRestart[blockIdx.x] = restart;
        }
    }
    grid.sync();
    if (tb0 >= tb1) return;

    // ---------------- phase 2: carry-in of the first cell, then stream the tiles ----------------
    uint32_t c = tile_first[tb0];
    uint32_t cb = lv.bnd[c], ce = lv.bnd[c + 1];
    uint32_t carry = 0;
    if (cb < chunkStart) {
        if (warp == 0) {   // walk the predecessors back to the block in which this cell began
            uint32_t acc = 0;
            int pos = (int)blockIdx.x - 1;
            for (;;) {
                const int idx = pos - lane;
                const uint32_t v = idx >= 0 ? blkLeft[idx] : 0u;
                const uint32_t r = idx >= 0 ? blkRestart[idx] : 1u;
                const unsigned m = __ballot_sync(0xffffffffu, r != 0u);
                const int fp = m ? (__ffs(m) - 1) : 32;
                acc += __reduce_add_sync(0xffffffffu, lane <= fp ? v : 0u);
                if (fp < 32) break;
                pos -= 32;
            }
            if (lane == 0) sm.lbsum[0] = acc;
        }
        __syncthreads();
        carry = sm.lbsum[0];
    }
    uint32_t nleft = lv.nleft_l[c];
    float cutv = final_cut[c];
    int axis = lv.axis[c];
    if (nh.enabled) {
        part_hist_zero(sm, nh);
        part_hist_enter(sm, nh, c);
    }

    int it = 0;
    for (uint32_t t = tb0; t < tb1; ++t, ++it) {
        const int bsel = it & 1;
        const uint32_t T = t * (uint32_t)kPartTile, tEnd = min(T + (uint32_t)kPartTile, nLocal);
        if (bulk) {
            // (the buffer being refilled was last read before the barrier that ended the previous iteration)
            if (t + 1 < tb1) part_prefetch_bulk(sm, bsel ^ 1, x, y, z, T + (uint32_t)kPartTile);
            mbar_wait(&sm.mbar[bsel], mphase[bsel]);
            mphase[bsel] ^= 1u;
        } else {
            if (t + 1 < tb1) {
                part_prefetch(sm, bsel ^ 1, x, y, z, T + (uint32_t)kPartTile);
                cp_async_commit();
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            __syncthreads();   // tile t is in sm.raw[bsel]
        }
        if (ce >= tEnd) {
            // ---- tile inside the running cell ----
            const uint32_t baseL = cb + carry;
            const uint32_t baseR = cb + nleft + ((T - cb) - carry);
            carry += part_tile_single(sm, bsel, 0u, tEnd - T, axis, cutv, baseL, baseR, x2, y2, z2, nh);
        } else {
            // ---- a cell boundary inside the tile: table-driven body, then re-seed the running cell ----
            if (nh.enabled) part_hist_flush(sm, nh, c);      // the table-driven body uses sm.sd
            uint32_t s0 = T, cfirst = c;
            bool firstSub = true;
            uint32_t lastSeg = 0;
            int ncell = 1;
            while (s0 < tEnd) {
                uint32_t s1;
                ncell = part_build_table(sm, lv, final_cut, nCells, cfirst, tEnd, s1);
                lastSeg = part_tile_body(sm, bsel, T, s0, s1, ncell, firstSub ? carry : 0u, x2, y2, z2, nh, cfirst);
                s0 = s1;
                cfirst += (uint32_t)ncell;
                firstSub = false;
            }
            c = cfirst - 1;            // cell that reaches the end of the tile
            cb = lv.bnd[c];
            ce = lv.bnd[c + 1];
            nleft = lv.nleft_l[c];
            cutv = final_cut[c];
            axis = lv.axis[c];
            carry = lastSeg;           // it began inside this tile, so this is all of its lefts so far
            if (nh.enabled) {
                part_hist_zero(sm, nh);
                part_hist_enter(sm, nh, c);
            }
        }
    }
    if (nh.enabled) part_hist_flush(sm, nh, c);
}

// Regime B: cells of at most a few tiles.  One block per cell; the block walks the cell's tiles in order with the
// carry in a register, so there is no inter-block dependency at all.
__global__ void __launch_bounds__(kThreads, 3) k_partition_cells(const float *__restrict__ x, const float *__restrict__ y,
                                                                const float *__restrict__ z, float *__restrict__ x2,
                                                                float *__restrict__ y2, float *__restrict__ z2,
                                                                LevelState lv, const float *__restrict__ final_cut,
                                                                uint32_t nCells, uint32_t nLocal, const uint32_t *__restrict__ gate,
                                                                NextHist nh) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PartSmem &sm = *reinterpret_cast<PartSmem *>(smem_raw);
    if (gate && *((volatile const uint32_t *)gate) != 0u) return;     // see k_split
    if (nh.enabled) part_hist_zero(sm, nh);
    for (uint32_t c = blockIdx.x; c < nCells; c += gridDim.x) {
        const uint32_t b = lv.bnd[c], e = lv.bnd[c + 1];
        if (e <= b) continue;   // block-uniform
        if (nh.enabled) part_hist_enter(sm, nh, c);
        const uint32_t nleft = lv.nleft_l[c];
        const float cutv = final_cut[c];
        const int axis = lv.axis[c];
        const uint32_t Tfirst = (b / (uint32_t)kPartTile) * (uint32_t)kPartTile;
        part_prefetch(sm, 0, x, y, z, Tfirst);
        cp_async_commit();
        uint32_t carry = 0u;
        int it = 0;
        for (uint32_t T = Tfirst; T < e; T += (uint32_t)kPartTile, ++it) {
            const int bsel = it & 1;
            if (T + (uint32_t)kPartTile < e) {
                part_prefetch(sm, bsel ^ 1, x, y, z, T + (uint32_t)kPartTile);
                cp_async_commit();
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            __syncthreads();
            const uint32_t s0 = max(b, T), s1 = min(e, T + (uint32_t)kPartTile);
            const uint32_t baseL = b + carry;
            const uint32_t baseR = b + nleft + ((s0 - b) - carry);
            carry += part_tile_single(sm, bsel, s0 - T, s1 - T, axis, cutv, baseL, baseR, x2, y2, z2, nh);
        }
        if (nh.enabled) part_hist_flush(sm, nh, c);
    }
    (void)nLocal;
}

// Regime C: cells of at most a couple of thousand particles (the deepest levels of a 2^20-leaf tree: 512 particles
// per cell and rank).  One WARP per cell, no shared memory, no barrier: lane i takes particle b + 32 k + i, the split
// position of each particle is a ballot prefix, lefts and rights go out as two contiguous runs.  A block-per-cell tile
// (regime B) would read a whole 2048-particle tile for a 512-particle cell.
__global__ void __launch_bounds__(kThreads) k_partition_warp(const float *__restrict__ x, const float *__restrict__ y,
                                                              const float *__restrict__ z, float *__restrict__ x2,
                                                              float *__restrict__ y2, float *__restrict__ z2, LevelState lv,
                                                              const float *__restrict__ final_cut, uint32_t nCells,
                                                              const uint32_t *__restrict__ gate) {
    if (gate && *((volatile const uint32_t *)gate) != 0u) return;     // see k_split
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const uint32_t warpId = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nWarps = (gridDim.x * blockDim.x) >> 5;
    // the next cell's scalars are fetched while the current cell streams (a cell is a few memory latencies, not bandwidth)
    uint32_t c = warpId;
    uint32_t nb_ = 0, ne_ = 0, nnl_ = 0;
    float ncut_ = 0.f;
    int nax_ = 0;
    if (c < nCells) { nb_ = lv.bnd[c]; ne_ = lv.bnd[c + 1]; nnl_ = lv.nleft_l[c]; ncut_ = final_cut[c]; nax_ = lv.axis[c]; }
    for (; c < nCells; c += nWarps) {
        const uint32_t b = nb_, e = ne_, nl = nnl_;
        const float cutv = ncut_;
        const int ax = nax_;
        const uint32_t cn = c + nWarps;
        if (cn < nCells) { nb_ = lv.bnd[cn]; ne_ = lv.bnd[cn + 1]; nnl_ = lv.nleft_l[cn]; ncut_ = final_cut[cn]; nax_ = lv.axis[cn]; }
        if (e <= b) continue;
        const float *col = pick_col(ax, x, y, z);
        uint32_t posL = b, posR = b + nl;
        constexpr int U = 4;        // iterations in flight
        for (uint32_t i0 = b; i0 < e; i0 += 32u * U) {
            float vx[U], vy[U], vz[U], vc[U];
            bool in[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint32_t i = i0 + (uint32_t)u * 32u + (uint32_t)lane;
                in[u] = i < e;
                vx[u] = in[u] ? __ldg(x + i) : 0.f;
                vy[u] = in[u] ? __ldg(y + i) : 0.f;
                vz[u] = in[u] ? __ldg(z + i) : 0.f;
                vc[u] = in[u] ? __ldg(col + i) : 0.f;      // (one of the three lines just requested)
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const bool left = in[u] && vc[u] < cutv;
                const unsigned mIn = __ballot_sync(0xffffffffu, in[u]);
                const unsigned mL = __ballot_sync(0xffffffffu, left);
                const unsigned mR = mIn & ~mL;
                if (in[u]) {
                    const uint32_t d = left ? posL + (uint32_t)__popc(mL & lt) : posR + (uint32_t)__popc(mR & lt);
                    x2[d] = vx[u]; y2[d] = vy[u]; z2[d] = vz[u];
                }
                posL += (uint32_t)__popc(mL);
                posR += (uint32_t)__popc(mR);
            }
        }
    }
}

// =====================================================================================
// Hoare-exact partition: reproduces partition.cpp:30-60 bit for bit, including WHICH particles
// with coord == cut land on which side and the resulting particle ORDER (later tie decisions depend on it).
//
// The sequential two-pointer loop is equivalent to (checked against a verbatim loop, tests/test_hoare_model.py):
//   S_i = positions with v >= cut, ascending ("i-stoppers"); S_j = positions with v <= cut, descending;
//   swap k happens iff S_i[k] < S_j[k]; K = number of swaps; i_stop = min(S_i[K], S_j[K-1]);
//   then rows i_stop and end-1 are exchanged (partition.cpp:52); left child = [begin, i_stop).
// Kernels: k_hoare_scan ranks the stoppers of every cell (exclusive prefix counts inside the cell) and writes their
// positions into two lists (stored in the idle ping-pong columns); k_hoare_swap exchanges the K pairs in place;
// k_hoare_finish finds K and i_stop per cell, does the final exchange and records the local left count.
// A cell without any particle >= cut makes the reference read (and swap) rows outside the cell; such degenerate
// cells are reported (ORB_ERR_RANGE) instead of emulated.
// =====================================================================================
struct HoareLists {
    uint32_t *posI;   // [n_local] cell c, rank k (from the left) of its v >= cut particles  -> position, at bnd[c] + k
    uint32_t *posJ;   // [n_local] cell c, rank r (from the left) of its v <= cut particles  -> position, at bnd[c] + r
    uint32_t *nGE;    // [nCells]
    uint32_t *nLE;    // [nCells]
};

// ranks + list writes for up to 2048 consecutive particles [p0, p0+cnt) of one cell: eight rows of 256 (row k, thread
// t <-> particle p0 + 256 k + t), one barrier pair for all rows.  Block-uniform call.
constexpr int kHoareRows = 8;
__device__ __forceinline__ void hoare_rows(const float *__restrict__ col, float cutv, uint32_t p0, uint32_t cnt, uint32_t listBase,
                                           uint32_t &carryGE, uint32_t &carryLE, const HoareLists &hl, uint32_t *s_w) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    bool ge[kHoareRows], le[kHoareRows];
    unsigned mg[kHoareRows], ml[kHoareRows];
#pragma unroll
    for (int k = 0; k < kHoareRows; ++k) {
        const uint32_t i = (uint32_t)k * kThreads + tid;
        ge[k] = le[k] = false;
        if (i < cnt) {
            const float v = __ldg(col + p0 + i);
            ge[k] = v >= cutv;
            le[k] = v <= cutv;
        }
        mg[k] = __ballot_sync(0xffffffffu, ge[k]);
        ml[k] = __ballot_sync(0xffffffffu, le[k]);
        if (lane == 0) { s_w[(k * kWarps + warp) * 2] = __popc(mg[k]); s_w[(k * kWarps + warp) * 2 + 1] = __popc(ml[k]); }
    }
    __syncthreads();
    const unsigned lt = (1u << lane) - 1u;
    uint32_t baseG = carryGE, baseL = carryLE;
#pragma unroll
    for (int k = 0; k < kHoareRows; ++k) {
        uint32_t preG = 0, preL = 0, totG = 0, totL = 0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
            const uint32_t a = s_w[(k * kWarps + w) * 2], b = s_w[(k * kWarps + w) * 2 + 1];
            if (w < warp) { preG += a; preL += b; }
            totG += a; totL += b;
        }
        const uint32_t pos = p0 + (uint32_t)k * kThreads + tid;
        if (ge[k]) hl.posI[listBase + baseG + preG + __popc(mg[k] & lt)] = pos;
        if (le[k]) hl.posJ[listBase + baseL + preL + __popc(ml[k] & lt)] = pos;
        baseG += totG;
        baseL += totL;
    }
    carryGE = baseG;
    carryLE = baseL;
    __syncthreads();
}

__global__ void __launch_bounds__(kThreads, 4) k_hoare_scan(const float *__restrict__ x, const float *__restrict__ y,
                                                           const float *__restrict__ z, LevelState lv,
                                                           const float *__restrict__ final_cut,
                                                           const uint32_t *__restrict__ tile_first, uint32_t nCells,
                                                           uint32_t nLocal, uint32_t nTiles, HoareLists hl,
                                                           uint32_t *blkGE, uint32_t *blkLE, uint32_t *blkRestart) {
    __shared__ uint32_t s_w[2 * kWarps * kHoareRows];
    __shared__ uint32_t s_c[2];
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t tilesPerBlock = (nTiles + gridDim.x - 1) / gridDim.x;
    const uint32_t tb0 = min(blockIdx.x * tilesPerBlock, nTiles), tb1 = min(tb0 + tilesPerBlock, nTiles);
    const uint32_t chunkStart = tb0 * (uint32_t)kPartTile, chunkEnd = min(tb1 * (uint32_t)kPartTile, nLocal);

    // ---- phase 1: stoppers of the trailing segment (the cell that continues into the next block) ----
    {
        uint32_t cg = 0, cl = 0, restart = 0;
        if (tb0 < tb1) {
            uint32_t c = tile_first[tb1 - 1];
            while (lv.bnd[c + 1] < chunkEnd) ++c;
            const uint32_t cb = lv.bnd[c];
            restart = cb >= chunkStart ? 1u : 0u;
            const float *col = pick_col(lv.axis[c], x, y, z);
            const float cutv = final_cut[c];
            for (uint32_t p = max(cb, chunkStart) + tid; p < chunkEnd; p += kThreads) {
                const float v = __ldg(col + p);
                cg += (v >= cutv);
                cl += (v <= cutv);
            }
        }
        cg = __reduce_add_sync(0xffffffffu, cg);
        cl = __reduce_add_sync(0xffffffffu, cl);
        if (lane == 0) { s_w[warp] = cg; s_w[kWarps + warp] = cl; }
        __syncthreads();
        if (tid == 0) {
            uint32_t a = 0, b = 0;
#pragma unroll
            for (int w = 0; w < kWarps; ++w) { a += s_w[w]; b += s_w[kWarps + w]; }
            blkGE[blockIdx.x] = a;
            blkLE[blockIdx.x] = b;
            blkRestart[blockIdx.x] = restart;
        }
        __syncthreads();
    }
    grid.sync();
    if (tb0 >= tb1) return;

    // ---- phase 2: carries of the first cell, then rank every particle of the chunk cell by cell ----
    uint32_t c = tile_first[tb0];
    uint32_t carryGE = 0, carryLE = 0;
    if (lv.bnd[c] < chunkStart) {
        if (warp == 0) {
            uint32_t ag = 0, al = 0;
            int pos = (int)blockIdx.x - 1;
            for (;;) {
                const int idx = pos - lane;
                const uint32_t vg = idx >= 0 ? blkGE[idx] : 0u, vl = idx >= 0 ? blkLE[idx] : 0u;
                const uint32_t r = idx >= 0 ? blkRestart[idx] : 1u;
                const unsigned m = __ballot_sync(0xffffffffu, r != 0u);
                const int fp = m ? (__ffs(m) - 1) : 32;
                ag += __reduce_add_sync(0xffffffffu, lane <= fp ? vg : 0u);
                al += __reduce_add_sync(0xffffffffu, lane <= fp ? vl : 0u);
                if (fp < 32) break;
                pos -= 32;
            }
            if (lane == 0) { s_c[0] = ag; s_c[1] = al; }
        }
        __syncthreads();
        carryGE = s_c[0];
        carryLE = s_c[1];
    }
    uint32_t p = chunkStart;
    while (p < chunkEnd) {
        const uint32_t cb = lv.bnd[c], ce = lv.bnd[c + 1];
        const uint32_t segEnd = min(ce, chunkEnd);
        if (segEnd > p) {
            const float *col = pick_col(lv.axis[c], x, y, z);
            const float cutv = final_cut[c];
            for (uint32_t q = p; q < segEnd; q += kThreads * kHoareRows)
                hoare_rows(col, cutv, q, min((uint32_t)(kThreads * kHoareRows), segEnd - q), cb, carryGE, carryLE, hl, s_w);
            p = segEnd;
        }
        if (ce <= chunkEnd) {   // the cell ends inside this chunk: its totals are complete
            if (tid == 0) { hl.nGE[c] = carryGE; hl.nLE[c] = carryLE; }
            carryGE = carryLE = 0;
            ++c;
            if (c >= nCells) break;
        }
    }
}

// final cut of every cell of the level = getCut() of its last margins (cell.h:74-76)
__global__ void k_final_cut(LevelState lv, uint32_t nCells, float *__restrict__ final_cut) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < nCells) final_cut[c] = mid_cut(lv.mL[c], lv.mR[c]);
}
__global__ void k_copy_u32(const uint32_t *__restrict__ src, uint32_t *__restrict__ dst, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i];
}

// exchange the K swap pairs of every cell, in place (all three columns)
__global__ void __launch_bounds__(kThreads) k_hoare_swap(float *__restrict__ x, float *__restrict__ y, float *__restrict__ z,
                                                        const uint32_t *__restrict__ bnd, const uint32_t *__restrict__ tile_first,
                                                        uint32_t nCells, uint32_t nLocal, HoareLists hl) {
    for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < nLocal; g += gridDim.x * blockDim.x) {
        uint32_t c = tile_first[g / (uint32_t)kMapTile];
        while (bnd[c + 1] <= g) ++c;
        const uint32_t b = bnd[c], k = g - b;
        const uint32_t nG = hl.nGE[c], nL = hl.nLE[c];
        if (k >= min(nG, nL)) continue;
        const uint32_t a = hl.posI[b + k], q = hl.posJ[b + nL - 1u - k];
        if (a < q) {
            float t;
            t = x[a]; x[a] = x[q]; x[q] = t;
            t = y[a]; y[a] = y[q]; y[q] = t;
            t = z[a]; z[a] = z[q]; z[q] = t;
        }
    }
}

// per cell: number of swaps K (the swap condition is monotone in k), i_stop, the final exchange with row end-1,
// local left count
__global__ void k_hoare_finish(float *__restrict__ x, float *__restrict__ y, float *__restrict__ z, LevelState lv,
                               uint32_t nCells, HoareLists hl, int *__restrict__ err) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nCells) return;
    const uint32_t b = lv.bnd[c], e = lv.bnd[c + 1];
    if (e <= b) { lv.nleft_l[c] = 0u; return; }
    const uint32_t nG = hl.nGE[c], nL = hl.nLE[c];
    if (nG == 0u) {   // the reference would scan past the end of the cell (partition.cpp:38): not emulated
        atomicExch(err, ORB_ERR_RANGE);
        lv.nleft_l[c] = e - b;
        return;
    }
    uint32_t lo = 0, hi = min(nG, nL);   // K in [lo, hi]: first k with posI[k] >= posJ[nL-1-k]
    while (lo < hi) {
        const uint32_t m = (lo + hi) >> 1;
        if (hl.posI[b + m] < hl.posJ[b + nL - 1u - m]) lo = m + 1; else hi = m;
    }
    const uint32_t K = lo;
    uint32_t iStop = 0xffffffffu;
    if (K < nG) iStop = hl.posI[b + K];
    if (K > 0u) iStop = min(iStop, hl.posJ[b + nL - K]);
    if (iStop < e - 1u) {   // partition.cpp:52: swap(particles, i, endInd - 1)
        float t;
        t = x[iStop]; x[iStop] = x[e - 1u]; x[e - 1u] = t;
        t = y[iStop]; y[iStop] = y[e - 1u]; y[e - 1u] = t;
        t = z[iStop]; z[iStop] = z[e - 1u]; z[e - 1u] = t;
    }
    lv.nleft_l[c] = iStop - b;
}

// =====================================================================================
// Bounding boxes (north-star extension): per-cell min/max of x,y,z.
// Floats are mapped to order-preserving uint32 so warp REDUX and global atomicMin/Max apply.
// =====================================================================================
__device__ __forceinline__ uint32_t f2ord(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__global__ void k_bbox_init(uint32_t *__restrict__ bb, uint32_t nCells) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nCells * 8u) return;
    bb[i] = ((i & 7u) < 3u) ? 0xffffffffu : 0u;   // mins start at +max, maxes at 0
}

__global__ void __launch_bounds__(kThreads) k_bbox(const float *__restrict__ x, const float *__restrict__ y,
                                                   const float *__restrict__ z, const uint32_t *__restrict__ bnd,
                                                   const uint32_t *__restrict__ tile_first, uint32_t nCells,
                                                   uint32_t nLocal, uint32_t nTiles, uint32_t *__restrict__ bb) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // Every lane keeps the running min / max of the cell its warp is in; the warp reduces and touches the cell's six
    // words only when it moves on to another cell or runs out of tiles.  (Reducing per 128-particle piece - the first
    // version - is 6 global atomics per piece on the SAME six words while a level has few cells: 4.8 ms for one cell
    // of 2^27 particles, profiles/r02v_bbox.txt.)
    uint32_t accCell = 0xffffffffu;
    uint32_t amn[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, amx[3] = {0u, 0u, 0u};
    auto flush = [&]() {      // warp-uniform
        if (accCell == 0xffffffffu) return;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const uint32_t wmn = __reduce_min_sync(0xffffffffu, amn[a]);
            const uint32_t wmx = __reduce_max_sync(0xffffffffu, amx[a]);
            if (lane == 0) {
                if (wmn != 0xffffffffu) atomicMin(&bb[accCell * 8u + a], wmn);
                if (wmx != 0u) atomicMax(&bb[accCell * 8u + 3 + a], wmx);
            }
            amn[a] = 0xffffffffu; amx[a] = 0u;
        }
    };
    for (uint32_t t = blockIdx.x; t < nTiles; t += gridDim.x) {
        const uint32_t t0 = t * (uint32_t)kMapTile, t1 = min(t0 + (uint32_t)kMapTile, nLocal);
        const uint32_t c = tile_first[t];
        for (uint32_t chunk = t0 + warp * 128u; chunk < t1; chunk += kWarps * 128u) {
            const uint32_t cend = min(chunk + 128u, t1);
            uint32_t cc = c;
            while (bnd[cc + 1] <= chunk) ++cc;
            const uint32_t e0 = chunk + lane * 4u;
            while (cc < nCells) {
                const uint32_t b = bnd[cc], e = bnd[cc + 1];
                if (b >= cend) break;
                const uint32_t lo = max(b, chunk), hi = min(e, cend);
                if (hi > lo) {
                    if (cc != accCell) { flush(); accCell = cc; }
                    if (e0 >= lo && e0 + 4u <= hi) {
                        const float4 q[3] = {__ldg(reinterpret_cast<const float4 *>(x + e0)),
                                             __ldg(reinterpret_cast<const float4 *>(y + e0)),
                                             __ldg(reinterpret_cast<const float4 *>(z + e0))};
#pragma unroll
                        for (int a = 0; a < 3; ++a) {
                            const float lo4 = fminf(fminf(q[a].x, q[a].y), fminf(q[a].z, q[a].w));
                            const float hi4 = fmaxf(fmaxf(q[a].x, q[a].y), fmaxf(q[a].z, q[a].w));
                            amn[a] = min(amn[a], f2ord(lo4)); amx[a] = max(amx[a], f2ord(hi4));
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const uint32_t ee = e0 + j;
                            if (ee >= lo && ee < hi) {
                                const uint32_t ux = f2ord(__ldg(x + ee)), uy = f2ord(__ldg(y + ee)), uz = f2ord(__ldg(z + ee));
                                amn[0] = min(amn[0], ux); amx[0] = max(amx[0], ux);
                                amn[1] = min(amn[1], uy); amx[1] = max(amx[1], uy);
                                amn[2] = min(amn[2], uz); amx[2] = max(amx[2], uz);
                            }
                        }
                    }
                }
                if (e > cend) break;
                ++cc;
            }
        }
    }
    flush();
}

__global__ void k_bbox_decode(const uint32_t *__restrict__ bb, uint32_t nCells, float *__restrict__ out6) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nCells * 6u) return;
    const uint32_t c = i / 6u, k = i % 6u;
    const uint32_t u = bb[c * 8u + k];
    // empty cell: +inf / -inf like a CPU loop that never executes
    float v;
    if (k < 3u) v = (u == 0xffffffffu) ? __int_as_float(0x7f800000) : ord2f(u);
    else v = (u == 0u) ? __int_as_float(0xff800000) : ord2f(u);
    out6[i] = v;
}

// tight-box mode: children take the particle bounding box as their box (axis = longest side, margins = faces)
__global__ void k_apply_bbox(orb_cell *__restrict__ cells, uint32_t nCells, const float *__restrict__ bb6,
                             const uint32_t *__restrict__ total_by_id) {
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nCells) return;
    orb_cell ch = cells[c];
    if (bb6[c * 6] <= bb6[c * 6 + 3]) {   // has particles
#pragma unroll
        for (int d = 0; d < 3; ++d) { ch.lower[d] = bb6[c * 6 + d]; ch.upper[d] = bb6[c * 6 + 3 + d]; }
        child_axis_margins(ch);
        if (ch.cutAxis < 0) ch.cutAxis = 0;
        cells[c] = ch;
    }
    (void)total_by_id;
}

}  // namespace orb
