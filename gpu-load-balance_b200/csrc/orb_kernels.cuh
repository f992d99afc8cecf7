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
//   k_partition    stable split of every cell in one launch: block scan + decoupled
//                  look-back across tiles with restarts at cell boundaries, scatter staged
//                  through shared memory.                      24 B / particle / level
//   k_bbox         per-cell min/max of x,y,z (north-star extension).  12 B / particle
//   k_level_setup, k_tile_map, k_split, k_finalize_*  O(nCells) bookkeeping.
#pragma once
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
};

// float midpoint exactly as Cell::getCut (cell.h:74-76): float add, halve, round to float.
// (R+L)/2.0 in double then cast == correctly rounded half of the float sum == __fmul_rn(sum,0.5f).
__device__ __forceinline__ float mid_cut(float L, float R) { return __fmul_rn(__fadd_rn(R, L), 0.5f); }

__device__ __forceinline__ const float *pick_col(int a, const float *x, const float *y, const float *z) {
    return a == 0 ? x : (a == 1 ? y : z);
}

// =====================================================================================
// Level setup: flatten Cell[] (the reference's wire format) into the SoA level state.
// Mirrors what ServiceCopyCells prepares per level (copyCells.cu:29-61) without block descriptors.
// =====================================================================================
__global__ void k_level_setup(const orb_cell *__restrict__ cells, uint32_t nCells, const uint32_t *__restrict__ range,
                              const uint32_t *__restrict__ total_by_id, LevelState lv, uint32_t nLocal, int nc,
                              int *__restrict__ err) {
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
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
    // trial cuts of the first pass: the implicit bisection tree below (L,R), heap order
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

// first cell whose range extends beyond the start of each map tile
__global__ void k_tile_map(const uint32_t *__restrict__ bnd, uint32_t nCells, uint32_t nMapTiles,
                           uint32_t *__restrict__ tile_first) {
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
// =====================================================================================
template <int NC>
__device__ __forceinline__ void count4(const float4 v, const float (&cv)[NC], unsigned (&cnt)[NC]) {
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        cnt[k] += (v.x < cv[k]);
        cnt[k] += (v.y < cv[k]);
        cnt[k] += (v.z < cv[k]);
        cnt[k] += (v.w < cv[k]);
    }
}

template <int NC>
__global__ void __launch_bounds__(kThreads, 4) k_count(const float *__restrict__ x, const float *__restrict__ y,
                                                    const float *__restrict__ z, LevelState lv,
                                                    const uint32_t *__restrict__ tile_first, uint32_t nCells,
                                                    uint32_t nLocal, uint32_t nTiles,
                                                    const uint32_t *__restrict__ gate) {
    if (gate && *gate == 0u) return;   // speculative pass after convergence: nothing to do
    __shared__ uint32_t s_acc[NC];
    __shared__ uint32_t s_cell[kCountCellsSmem * NC];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < NC) s_acc[tid] = 0u;
    __syncthreads();

    unsigned cnt[NC];
    float cv[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) { cnt[k] = 0u; cv[k] = 0.f; }
    int cur = -1;
    const float *col = x;

    // one global atomic per block per (cell, cut): warp REDUX -> shared -> global
    auto flush = [&]() {
        if (cur < 0) return;
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            unsigned v = __reduce_add_sync(0xffffffffu, cnt[k]);
            if (lane == 0 && v) atomicAdd(&s_acc[k], v);
            cnt[k] = 0u;
        }
        __syncthreads();
        if (tid < NC) {
            unsigned v = s_acc[tid];
            if (v) atomicAdd(&lv.cnt_l[(uint32_t)cur * kCS + tid], v);
            s_acc[tid] = 0u;
        }
        __syncthreads();
    };

    for (uint32_t t = blockIdx.x; t < nTiles; t += gridDim.x) {
        const uint32_t t0 = t * (uint32_t)kCountTile;
        const uint32_t t1 = min(t0 + (uint32_t)kCountTile, nLocal);
        const uint32_t c = tile_first[t * (kCountTile / kMapTile)];
        const uint32_t cb = lv.bnd[c], ce = lv.bnd[c + 1];
        if (cb <= t0 && ce >= t1) {
            // ---- tile inside one cell (block-uniform branch) ----
            if (!lv.active[c]) continue;
            if ((int)c != cur) {
                flush();
                cur = (int)c;
#pragma unroll
                for (int k = 0; k < NC; ++k) cv[k] = lv.cuts[c * kCS + k];
                col = pick_col(lv.axis[c], x, y, z);
            }
            if (t1 - t0 == (uint32_t)kCountTile) {
                const float4 *p = reinterpret_cast<const float4 *>(col + t0) + tid;
                float4 v0 = __ldg(p), v1 = __ldg(p + kThreads), v2 = __ldg(p + 2 * kThreads), v3 = __ldg(p + 3 * kThreads);
                count4<NC>(v0, cv, cnt);
                count4<NC>(v1, cv, cnt);
                count4<NC>(v2, cv, cnt);
                count4<NC>(v3, cv, cnt);
            } else {
                for (uint32_t e = t0 + tid; e < t1; e += kThreads) {
                    float v = __ldg(col + e);
#pragma unroll
                    for (int k = 0; k < NC; ++k) cnt[k] += (v < cv[k]);
                }
            }
        } else {
            // ---- fragmented tile: several cells; per-warp 128-element chunks ----
            flush();
            cur = -1;
            for (int i = tid; i < kCountCellsSmem * NC; i += kThreads) s_cell[i] = 0u;
            __syncthreads();
            for (uint32_t chunk = t0 + warp * 128u; chunk < t1; chunk += kWarps * 128u) {
                const uint32_t cend = min(chunk + 128u, t1);
                uint32_t cc = c;
                while (lv.bnd[cc + 1] <= chunk) ++cc;
                const uint32_t e0 = chunk + lane * 4u;
                while (cc < nCells) {
                    const uint32_t b = lv.bnd[cc], e = lv.bnd[cc + 1];
                    if (b >= cend) break;
                    const uint32_t lo = max(b, chunk), hi = min(e, cend);
                    if (hi > lo && lv.active[cc]) {
                        const float *cl = pick_col(lv.axis[cc], x, y, z);
                        float v[4];
                        bool in[4];
                        if (e0 >= lo && e0 + 4u <= hi) {
                            float4 q = __ldg(reinterpret_cast<const float4 *>(cl + e0));
                            v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
                            in[0] = in[1] = in[2] = in[3] = true;
                        } else {
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                uint32_t ee = e0 + j;
                                in[j] = (ee >= lo && ee < hi);
                                v[j] = in[j] ? __ldg(cl + ee) : 0.f;
                            }
                        }
#pragma unroll
                        for (int k = 0; k < NC; ++k) {
                            float ck = lv.cuts[cc * kCS + k];
                            unsigned n = 0;
#pragma unroll
                            for (int j = 0; j < 4; ++j) n += (in[j] && v[j] < ck);
                            n = __reduce_add_sync(0xffffffffu, n);
                            if (lane == 0 && n) {
                                uint32_t j = cc - c;
                                if (j < (uint32_t)kCountCellsSmem) atomicAdd(&s_cell[j * NC + k], n);
                                else atomicAdd(&lv.cnt_l[cc * kCS + k], n);
                            }
                        }
                    }
                    if (e > cend) break;
                    ++cc;
                }
            }
            __syncthreads();
            for (int i = tid; i < kCountCellsSmem * NC; i += kThreads) {
                unsigned v = s_cell[i];
                if (v) atomicAdd(&lv.cnt_l[(c + i / NC) * kCS + (i % NC)], v);
            }
            __syncthreads();
        }
    }
    flush();
}

// =====================================================================================
// Bisection update: orbit.cpp:191-231 replayed on the device, one thread per cell.
// M steps per pass over the counted trial-cut tree; literal float arithmetic of the reference:
//   float ratio = ceil(nLeafCells/2.0)/nLeafCells;  int difference = countLeft - count*ratio;
// =====================================================================================
struct PassCtl {
    uint32_t *n_active;        // device: [maxPasses+2] active cells after pass p (index p+1); [0] = before first pass
    uint32_t *done;            // device: [maxPasses+2] block tickets
    volatile uint32_t *h_status;   // pinned host: [maxPasses+2] n_active+1 after pass p (0 = not yet known)
    unsigned long long *active_particles;   // device: [0] local particles streamed (cells active in a pass, per HBM pass)
                                            //         [1] the same weighted by bisection iterations consumed (reference-equivalent)
    int32_t *level_iters;      // device: max iterations over cells (the reference's j)
};

template <int M>
__global__ void __launch_bounds__(kThreads) k_update(LevelState lv, uint32_t nCells, int pass, PassCtl ctl) {
    constexpr int NC = (1 << M) - 1;
    const uint32_t gate = ctl.n_active[pass];
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
        float L = lv.mL[c], R = lv.mR[c];
        it = lv.iter[c];
        const int it0 = it;
        const uint32_t total = lv.total[c];
        const int nleaf = lv.nleaf[c];
        const float ratio = (float)(ceil(nleaf / 2.0) / nleaf);          // orbit.cpp:204
        const float prod = __fmul_rn(__uint2float_rn(total), ratio);      // oCounts[i] * ratio
        npart = (unsigned long long)(lv.bnd[c + 1] - lv.bnd[c]);
        int node = 0;
        bool fnd = false;
#pragma unroll
        for (int s = 0; s < M; ++s) {
            const float cut = lv.cuts[c * kCS + node];
            const uint32_t cl = lv.cnt_g[c * kCS + node];
            const int diff = __float2int_rz(__fsub_rn(__uint2float_rn(cl), prod));   // orbit.cpp:205
            ++it;
            if (abs(diff) < 3) {                                                      // orbit.cpp:208
                fnd = true;
                lv.nleft_g[c] = cl;
                lv.nleft_l[c] = lv.cnt_l[c * kCS + node];
                break;
            } else if (diff > 0) { R = cut; node = 2 * node + 1; }                    // orbit.cpp:219
            else { L = cut; node = 2 * node + 2; }                                    // orbit.cpp:227
            if (it >= kMaxIter) break;                                                // orbit.cpp:149
        }
        lv.mL[c] = L; lv.mR[c] = R; lv.iter[c] = it;
        nipart = npart * (unsigned long long)(it - it0);
        if (fnd) { lv.found[c] = 1u; lv.active[c] = 0u; }
        else if (it >= kMaxIter) { lv.active[c] = 0u; }
        else {
            still = 1;
            float cv[7], lo[7], hi[7];
            lo[0] = L; hi[0] = R;
#pragma unroll
            for (int k = 0; k < NC; ++k) {
                cv[k] = mid_cut(lo[k], hi[k]);
                if (2 * k + 2 < NC) { lo[2 * k + 1] = lo[k]; hi[2 * k + 1] = cv[k]; lo[2 * k + 2] = cv[k]; hi[2 * k + 2] = hi[k]; }
            }
#pragma unroll
            for (int k = 0; k < NC; ++k) lv.cuts[c * kCS + k] = cv[k];
        }
#pragma unroll
        for (int k = 0; k < NC; ++k) lv.cnt_l[c * kCS + k] = 0u;
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
        if (s_n) atomicAdd(&ctl.n_active[pass + 1], s_n);
        if (s_p) atomicAdd(ctl.active_particles, s_p);
        if (s_q) atomicAdd(ctl.active_particles + 1, s_q);
        if (s_it) atomicMax(ctl.level_iters, s_it);
        __threadfence();
        uint32_t ticket = atomicAdd(&ctl.done[pass], 1u);
        if (ticket == gridDim.x - 1) {
            __threadfence();
            uint32_t n = *((volatile uint32_t *)&ctl.n_active[pass + 1]);
            ctl.h_status[pass] = n + 1u;   // mapped pinned memory: the host polls this, no stream sync
            __threadfence_system();
        }
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

__global__ void k_split(orb_cell *__restrict__ heap, uint32_t first, uint32_t nCells, LevelState lv,
                        uint32_t *__restrict__ range, uint32_t *__restrict__ total_by_id, float *__restrict__ final_cut) {
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
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
// One launch for all cells.  Tiles of kPartTile particles; tile ids come from an atomic ticket so
// look-back only ever waits on tiles that already run.  A tile needs a carry-in (number of left
// particles of its first cell in earlier tiles) only if that cell began before the tile; tiles that
// contain a cell boundary publish an inclusive prefix at once, so look-back chains restart at every
// cell boundary.
// =====================================================================================
constexpr uint64_t kStAgg = 1ull, kStPrefix = 2ull;
__device__ __forceinline__ uint64_t pack_state(uint32_t epoch, uint64_t st, uint32_t v) {
    return ((uint64_t)epoch << 34) | (st << 32) | (uint64_t)v;
}

__global__ void __launch_bounds__(kThreads, 3) k_partition(const float *__restrict__ x, const float *__restrict__ y,
                                                        const float *__restrict__ z, float *__restrict__ x2,
                                                        float *__restrict__ y2, float *__restrict__ z2,
                                                        LevelState lv, const float *__restrict__ final_cut,
                                                        const uint32_t *__restrict__ tile_first, uint32_t nCells,
                                                        uint32_t nLocal, uint32_t nTiles, uint64_t *tile_state,
                                                        uint32_t epoch, uint32_t *ticket) {
    __shared__ float sx[kPartTile], sy[kPartTile], sz[kPartTile];
    __shared__ uint32_t sd[kPartTile];
    __shared__ uint32_t s_cbeg[kPartCells], s_cend[kPartCells], s_nleft[kPartCells], s_B[kPartCells], s_Bend[kPartCells];
    __shared__ float s_cut[kPartCells];
    __shared__ int s_axis[kPartCells];
    __shared__ uint32_t s_warp[kWarps];
    __shared__ uint32_t s_tile, s_carry;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t t = s_tile;
    if (t >= nTiles) return;
    const uint32_t t0 = t * (uint32_t)kPartTile, t1 = min(t0 + (uint32_t)kPartTile, nLocal);
    const uint32_t c0 = tile_first[t];
    const bool needCarry = lv.bnd[c0] < t0;

    uint32_t s0 = t0, cfirst = c0;
    bool firstSub = true;
    while (s0 < t1) {
        // ---- cell table of this sub-range: cells cfirst+i that begin before t1 ----
        uint32_t cidx = cfirst + tid;
        bool valid = false;
        uint32_t cb = 0, ce = 0;
        if (cidx < nCells) {
            cb = lv.bnd[cidx];
            valid = (tid == 0) || (cb < t1);
            if (valid) ce = lv.bnd[cidx + 1];
        }
        const int ncell = __syncthreads_count(valid);   // cells are consecutive, so valid is a prefix of the threads
        uint32_t s1 = t1;
        if (valid) {
            s_cbeg[tid] = cb; s_cend[tid] = ce;
            s_nleft[tid] = lv.nleft_l[cidx];
            s_cut[tid] = final_cut[cidx];
            s_axis[tid] = lv.axis[cidx];
        }
        __syncthreads();
        if (ncell == kPartCells) s1 = min(s_cend[kPartCells - 1], t1);   // table full: finish the rest in another sub-range
        const bool lastSub = (s1 == t1);

        // ---- load 8 particles per thread: two float4 groups, warp-striped ----
        float vx[8], vy[8], vz[8];
        bool ok[8], fl[8];
        uint32_t jj[8];
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            const uint32_t o = warp * 256u + g * 128u + lane * 4u;   // offset in tile
            const uint32_t e = t0 + o;
            if (e >= s0 && e + 4u <= s1) {
                float4 a = __ldg(reinterpret_cast<const float4 *>(x + e));
                float4 b = __ldg(reinterpret_cast<const float4 *>(y + e));
                float4 cq = __ldg(reinterpret_cast<const float4 *>(z + e));
                vx[4 * g] = a.x; vx[4 * g + 1] = a.y; vx[4 * g + 2] = a.z; vx[4 * g + 3] = a.w;
                vy[4 * g] = b.x; vy[4 * g + 1] = b.y; vy[4 * g + 2] = b.z; vy[4 * g + 3] = b.w;
                vz[4 * g] = cq.x; vz[4 * g + 1] = cq.y; vz[4 * g + 2] = cq.z; vz[4 * g + 3] = cq.w;
#pragma unroll
                for (int k = 0; k < 4; ++k) ok[4 * g + k] = true;
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t ee = e + k;
                    const bool in = (ee >= s0 && ee < s1);
                    ok[4 * g + k] = in;
                    vx[4 * g + k] = in ? __ldg(x + ee) : 0.f;
                    vy[4 * g + k] = in ? __ldg(y + ee) : 0.f;
                    vz[4 * g + k] = in ? __ldg(z + ee) : 0.f;
                }
            }
        }
        // ---- cell of each particle, left flag ----
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            const uint32_t e = t0 + warp * 256u + g * 128u + lane * 4u;
            uint32_t j = 0;
            if (ncell > 1) {   // largest j with s_cbeg[j] <= e (j >= 1 entries are sorted begins inside the tile)
                uint32_t lo = 0, hi = (uint32_t)ncell - 1;
                while (lo < hi) {
                    uint32_t m = (lo + hi + 1) >> 1;
                    if (s_cbeg[m] <= e) lo = m; else hi = m - 1;
                }
                j = lo;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t ee = e + k;
                while (j + 1 < (uint32_t)ncell && s_cbeg[j + 1] <= ee) ++j;
                jj[4 * g + k] = j;
                const int a = s_axis[j];
                const float v = a == 0 ? vx[4 * g + k] : (a == 1 ? vy[4 * g + k] : vz[4 * g + k]);
                fl[4 * g + k] = ok[4 * g + k] && (v < s_cut[j]);
            }
        }
        // ---- block exclusive scan of the left flags (order: warp region, group, lane, k) ----
        uint32_t c0n = (uint32_t)fl[0] + fl[1] + fl[2] + fl[3];
        uint32_t c1n = (uint32_t)fl[4] + fl[5] + fl[6] + fl[7];
        uint32_t i0 = c0n, i1 = c1n;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t a = __shfl_up_sync(0xffffffffu, i0, o);
            uint32_t b = __shfl_up_sync(0xffffffffu, i1, o);
            if (lane >= o) { i0 += a; i1 += b; }
        }
        const uint32_t tot0 = __shfl_sync(0xffffffffu, i0, 31), tot1 = __shfl_sync(0xffffffffu, i1, 31);
        if (lane == 0) s_warp[warp] = tot0 + tot1;
        __syncthreads();
        uint32_t woff = 0, total = 0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
            const uint32_t v = s_warp[w];
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
        for (int g = 0; g < 2; ++g)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int q = 4 * g + k;
                if (!ok[q]) continue;
                const uint32_t ee = t0 + warp * 256u + g * 128u + lane * 4u + k;
                const uint32_t j = jj[q];
                if (ee == max(s_cbeg[j], s0)) s_B[j] = LE[q];
                if (ee + 1u == min(s_cend[j], s1)) s_Bend[j] = LE[q] + fl[q];
            }
        __syncthreads();

        // ---- publish this tile's state, fetch the carry-in ----
        if (firstSub && tid == 0) s_carry = 0u;
        if (lastSub && tid == 0) {
            const uint32_t lastSeg = total - s_B[ncell - 1];   // left particles of the segment that reaches t1
            const bool restart = !(firstSub && ncell == 1 && needCarry);
            ((volatile uint64_t *)tile_state)[t] = pack_state(epoch, restart ? kStPrefix : kStAgg, lastSeg);
        }
        if (firstSub && needCarry && warp == 0) {
            uint32_t carry = 0;
            int pos = (int)t - 1;
            for (;;) {
                const int idx = pos - lane;
                uint64_t st = pack_state(epoch, kStPrefix, 0u);
                if (idx >= 0) st = ((volatile uint64_t *)tile_state)[idx];
                const bool okst = ((uint32_t)(st >> 34) == epoch) && (((st >> 32) & 3ull) != 0ull);
                const bool isP = okst && (((st >> 32) & 3ull) == kStPrefix);
                const unsigned inval = __ballot_sync(0xffffffffu, !okst);
                const unsigned pref = __ballot_sync(0xffffffffu, isP);
                const int fp = pref ? (__ffs(pref) - 1) : 32;
                const unsigned need = (fp >= 31) ? 0xffffffffu : ((2u << fp) - 1u);
                if (inval & need) continue;   // a predecessor has not published yet: poll again
                const uint32_t contrib = (lane <= fp) ? (uint32_t)(st & 0xffffffffull) : 0u;
                carry += __reduce_add_sync(0xffffffffu, contrib);
                if (fp < 32) break;
                pos -= 32;
            }
            if (lane == 0) {
                s_carry = carry;
                if (lastSub && ncell == 1)
                    ((volatile uint64_t *)tile_state)[t] = pack_state(epoch, kStPrefix, carry + total);
            }
        }
        __syncthreads();
        const uint32_t carry0 = firstSub ? s_carry : 0u;

        // ---- stage in shared memory in destination order (per segment: lefts, then rights) ----
#pragma unroll
        for (int g = 0; g < 2; ++g)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int q = 4 * g + k;
                if (!ok[q]) continue;
                const uint32_t ee = t0 + warp * 256u + g * 128u + lane * 4u + k;
                const uint32_t j = jj[q];
                const uint32_t cbj = s_cbeg[j];
                const uint32_t segStart = max(cbj, s0);
                const uint32_t lb = LE[q] - s_B[j];                 // lefts of this cell before me, inside this sub-range
                const uint32_t cj = (j == 0) ? carry0 : 0u;         // ... and in earlier tiles
                uint32_t p, d;
                if (fl[q]) {
                    p = (segStart - t0) + lb;
                    d = cbj + cj + lb;
                } else {
                    const uint32_t nLt = s_Bend[j] - s_B[j];
                    const uint32_t rb = (ee - segStart) - lb;       // rights of this cell before me, inside this sub-range
                    p = (segStart - t0) + nLt + rb;
                    d = cbj + s_nleft[j] + ((ee - cbj) - (cj + lb));
                }
                sx[p] = vx[q]; sy[p] = vy[q]; sz[p] = vz[q]; sd[p] = d;
            }
        __syncthreads();
        // ---- coalesced runs out to the ping-pong columns ----
        for (uint32_t o = (s0 - t0) + tid; o < (s1 - t0); o += kThreads) {
            const uint32_t d = sd[o];
            x2[d] = sx[o]; y2[d] = sy[o]; z2[d] = sz[o];
        }
        __syncthreads();
        s0 = s1;
        cfirst += (uint32_t)ncell;
        firstSub = false;
    }
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
                    uint32_t mn[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, mx[3] = {0u, 0u, 0u};
                    if (e0 >= lo && e0 + 4u <= hi) {
                        const float4 q[3] = {__ldg(reinterpret_cast<const float4 *>(x + e0)),
                                             __ldg(reinterpret_cast<const float4 *>(y + e0)),
                                             __ldg(reinterpret_cast<const float4 *>(z + e0))};
#pragma unroll
                        for (int a = 0; a < 3; ++a) {
                            const float lo4 = fminf(fminf(q[a].x, q[a].y), fminf(q[a].z, q[a].w));
                            const float hi4 = fmaxf(fmaxf(q[a].x, q[a].y), fmaxf(q[a].z, q[a].w));
                            mn[a] = f2ord(lo4); mx[a] = f2ord(hi4);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const uint32_t ee = e0 + j;
                            if (ee >= lo && ee < hi) {
                                const uint32_t ux = f2ord(__ldg(x + ee)), uy = f2ord(__ldg(y + ee)), uz = f2ord(__ldg(z + ee));
                                mn[0] = min(mn[0], ux); mx[0] = max(mx[0], ux);
                                mn[1] = min(mn[1], uy); mx[1] = max(mx[1], uy);
                                mn[2] = min(mn[2], uz); mx[2] = max(mx[2], uz);
                            }
                        }
                    }
#pragma unroll
                    for (int a = 0; a < 3; ++a) {
                        const uint32_t wmn = __reduce_min_sync(0xffffffffu, mn[a]);
                        const uint32_t wmx = __reduce_max_sync(0xffffffffu, mx[a]);
                        if (lane == 0) {
                            atomicMin(&bb[cc * 8u + a], wmn);
                            atomicMax(&bb[cc * 8u + 3 + a], wmx);
                        }
                    }
                }
                if (e > cend) break;
                ++cc;
            }
        }
    }
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
