// orb_exchange.cuh — the selection-based cut search over several ranks, any level size (replaces the
// per-iteration Combine of countLeft.cpp:44-53 inside the loop orbit.cpp:149-189 by FOUR flag exchanges per level,
// none of them a collective call).
//
// Every rank maps every rank's exchange arena (NVLink peer memory, orb_peer_import).  Per level:
//
//   HIST      local histogram rows of every cell (k_sel_stream<HIST> / NextHist rows of the partition / k_xd_hist)
//   REDUCE    k_xr_reduce: flat all-reduce of the rows inside one kernel - rank r pulls its 1/R share of the words
//             from every rank (16-byte remote loads), sums, and pushes the sums into every rank's hist_g (16-byte
//             remote stores): every rank then holds the global rows                                    [barrier 1]
//   COMPACT   every rank resolves every cell from the global rows (identical data -> identical candidate bins) and
//             gathers its own candidates; they are pushed into the slot of the cell's OWNER rank (cell c is owned by
//             rank c mod R): streaming levels by k_xc_push after k_sel_stream<COMPACT>, small cells directly by
//             k_xd_compact                                                                              [barrier 2]
//   FINISH    k_xf_finish: the owner searches the candidates of all ranks (the reference's decisions replayed on exact
//             counts) and pushes a 32-byte result record to every rank: margins, iterations, found, global left count
//             and THAT rank's number of candidates left of the final cut                                [barrier 3]
//   APPLY     k_xa_apply: every rank copies the records into its level state; nleft_l = own particles below the
//             candidate bins + own candidates left of the cut; reports 1 + cells flagged to the host    [barrier 4]
//
// Work per rank is 1/R of the cells in REDUCE and FINISH (the expensive exchanges); the wire carries each row word
// and each candidate once per direction.  All inputs of a decision are exchanged data, so margins / iterations /
// counts are bit-identical on every rank, and so is the set of flagged cells (too many candidates, a slot that
// overflowed on some rank, a capped cell whose final cut leaves the candidate bins): they fall back together to the
// host-driven iterative loop.
//
// Barrier n of a level: block 0 of the kernel FOLLOWING the producer stores the exchange's sequence number into every
// peer's flag word (after griddepcontrol.wait: the producer has completed, a system fence orders its remote stores
// before the flag); every block spins on its own rank's flag words.  One buffer of each kind suffices: a rank
// overwrites a buffer of level l only after a barrier that every peer reaches after it has finished reading level l.
#pragma once
#include "orb_select.cuh"

namespace orb {

__device__ __forceinline__ uint32_t x_owned_stride(uint32_t nCells, int n) { return (nCells + (uint32_t)n - 1u) / (uint32_t)n; }

// ---- REDUCE: hist_g (everywhere) = sum over ranks of hist_l; n4 = words / 4 ----
__global__ void __launch_bounds__(kThreads) k_xr_reduce(XArena xa, uint32_t n4) {
    pdl_enter();
    x_barrier(xa);
    const uint32_t per = (((n4 + (uint32_t)xa.n - 1u) / (uint32_t)xa.n) + 7u) & ~7u;
    const uint32_t i0 = min((uint32_t)xa.self * per, n4), i1 = min(i0 + per, n4);
    for (uint32_t i = i0 + blockIdx.x * blockDim.x + threadIdx.x; i < i1; i += gridDim.x * blockDim.x) {
        uint4 b[kMaxPeers];
#pragma unroll
        for (int r = 0; r < kMaxPeers; ++r) {
            b[r] = make_uint4(0u, 0u, 0u, 0u);
            if (r < xa.n) b[r] = __ldcg(reinterpret_cast<const uint4 *>(xa.arena[r] + xa.offHistL) + i);
        }
        uint4 a = b[0];
#pragma unroll
        for (int r = 1; r < kMaxPeers; ++r) { a.x += b[r].x; a.y += b[r].y; a.z += b[r].z; a.w += b[r].w; }
#pragma unroll
        for (int r = 0; r < kMaxPeers; ++r)
            if (r < xa.n) reinterpret_cast<uint4 *>(xa.arena[r] + xa.offHistG)[i] = a;
    }
}

// ---- streaming levels, after COMPACT: publish every cell's resolve (also of cells without local particles), the
//      local particles below the candidate bins, and push the own candidates + their number to the cell's owner ----
__global__ void __launch_bounds__(kThreads) k_xc_push(LevelState lv, SelState sg /* hist = global rows */, SelMrState mr, XArena xa,
                                                       XArena xbar /* n > 1: this kernel makes barrier 2 (no COMPACT pass ran) */,
                                                       uint32_t nCells, int nb1, uint32_t candCap) {
    extern __shared__ __align__(16) unsigned char sel_smem[];
    uint32_t *hbuf = reinterpret_cast<uint32_t *>(sel_smem);
    __shared__ SelResolveSmem rs;
    __shared__ uint32_t s_red[kWarps];
    pdl_enter();
    x_barrier(xbar);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t stride = x_owned_stride(nCells, xa.n);
    for (uint32_t c = blockIdx.x; c < nCells; c += gridDim.x) {
        const int owner = (int)(c % (uint32_t)xa.n);
        const uint32_t slot = (uint32_t)xa.self * stride + c / (uint32_t)xa.n;
        uint32_t *cntDst = xa.arena[owner] + xa.offRecvCnt + slot;
        if (!lv.active[c]) {          // block-uniform
            if (tid == 0) { sg.flag[c] = 0u; *cntDst = 0u; mr.loc_base[c] = 0u; sg.cursor[c] = 0u; }
            continue;
        }
        uint32_t bf, bl;
        sel_resolve_cell(lv, sg, c, nb1, candCap, true, hbuf, rs, bf, bl, false, false);
        uint32_t s = 0;
        if (bf <= bl) for (uint32_t i = tid; i < bf; i += kThreads) s += __ldcg(mr.hist_l + (size_t)c * nb1 + i);
        s = __reduce_add_sync(0xffffffffu, s);
        if (lane == 0) s_red[warp] = s;
        const uint32_t n = __ldcg(&sg.cursor[c]);
        const uint32_t n4 = (min(n, mr.slotWords - 1u) + 3u) >> 2;
        const uint4 *src = reinterpret_cast<const uint4 *>(mr.slots_l + (size_t)c * mr.slotWords);
        uint4 *dst = reinterpret_cast<uint4 *>(xa.arena[owner] + xa.offRecv + (size_t)slot * mr.slotWords);
        for (uint32_t i = tid; i < n4; i += kThreads) dst[i] = __ldcg(src + i);
        __syncthreads();
        if (tid == 0) {
            uint32_t t = 0;
#pragma unroll
            for (int w = 0; w < kWarps; ++w) t += s_red[w];
            mr.loc_base[c] = t;
            *cntDst = n;                 // > slotWords - 1: the slot overflowed, the owner flags the cell
            sg.cursor[c] = 0u;           // zero between levels (COMPACT counts into it)
        }
        __syncthreads();
    }
}

// ---- small cells: one group of G threads (warp or block) per cell ----
// apply f(value) to the K values at src (4-byte aligned): 16-byte loads over the aligned body, U in flight per thread
template <int G, int U, typename F>
__device__ __forceinline__ void grp_for_each(const float *__restrict__ src, uint32_t K, int gtid, F f) {
    const uint32_t mis = (uint32_t)((reinterpret_cast<uintptr_t>(src) >> 2) & 3u);
    const uint32_t head = mis ? min(4u - mis, K) : 0u;
    const uint32_t body4 = (K - head) / 4u;
    const float4 *g4 = reinterpret_cast<const float4 *>(src + head);
    uint32_t i = (uint32_t)gtid;
    for (; i + (uint32_t)(U - 1) * G < body4; i += (uint32_t)U * G) {
        float4 q[U];
#pragma unroll
        for (int u = 0; u < U; ++u) q[u] = __ldg(g4 + i + (uint32_t)u * G);
#pragma unroll
        for (int u = 0; u < U; ++u) { f(q[u].x); f(q[u].y); f(q[u].z); f(q[u].w); }
    }
    for (; i < body4; i += G) { const float4 a = __ldg(g4 + i); f(a.x); f(a.y); f(a.z); f(a.w); }
    if ((uint32_t)gtid < head) f(__ldg(src + gtid));
    const uint32_t tail0 = head + body4 * 4u;
    if (tail0 + (uint32_t)gtid < K) f(__ldg(src + tail0 + gtid));
}
template <int G>
__device__ __forceinline__ void grp_sync() {
    if (G == 32) __syncwarp(); else __syncthreads();
}

// HIST of small cells: the group bins its cell into shared memory and stores the row (no atomics, no clearing)
template <int G>
__global__ void __launch_bounds__(kThreads) k_xd_hist(const float *__restrict__ x, const float *__restrict__ y,
                                                      const float *__restrict__ z, LevelState lv, uint32_t *__restrict__ hist_l,
                                                      uint32_t nCells, int nb) {
    extern __shared__ __align__(16) unsigned char sel_smem[];
    constexpr int GPB = kThreads / G;
    pdl_enter();
    const int tid = threadIdx.x, gtid = tid % G, grp = tid / G;
    uint32_t *h = reinterpret_cast<uint32_t *>(sel_smem) + (size_t)grp * nb;
    const float nbm1 = (float)(nb - 1);
    for (uint32_t c = blockIdx.x * GPB + grp; c < nCells; c += gridDim.x * GPB) {      // (G == 256: c is block-uniform)
        if (!lv.active[c]) continue;
        const uint32_t b = lv.bnd[c], K = lv.bnd[c + 1] - b;
        const float lo = lv.mL[c], scale = sel_scale(lo, lv.mR[c], nb);
        for (int i = gtid; i < nb; i += G) h[i] = 0u;
        grp_sync<G>();
        grp_for_each<G, 4>(pick_col(lv.axis[c], x, y, z) + b, K, gtid, [&](float v) {
            const float t = fminf(fmaxf(__fmul_rn(__fsub_rn(v, lo), scale), 0.f), nbm1);      // == sel_bin(v, lo, scale, nb)
            atomicAdd(&h[__float2int_rz(t)], 1u);
        });
        grp_sync<G>();
        for (int i = gtid; i < nb; i += G) hist_l[(size_t)c * nb + i] = h[i];
        grp_sync<G>();
    }
}

// group-wide exclusive scan + total of one value per thread (G == 32: shuffles; G == 256: sel_block_scan)
template <int G>
__device__ __forceinline__ uint32_t grp_scan(uint32_t v, uint32_t *s_w, uint32_t &total) {
    if (G == 32) {
        const int lane = threadIdx.x & 31;
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += up;
        }
        total = __shfl_sync(0xffffffffu, incl, 31);
        return incl - v;
    } else {
        return sel_block_scan(v, s_w, total);
    }
}

struct XdSmem {               // per group
    int first, last;
    uint32_t base, end, cnt;
};

// RESOLVE + COMPACT of small cells: candidate bins from the global row, own candidates straight into the owner's slot
template <int G>
__global__ void __launch_bounds__(kThreads) k_xd_compact(const float *__restrict__ x, const float *__restrict__ y,
                                                         const float *__restrict__ z, LevelState lv, SelState ss, SelMrState mr,
                                                         XArena xa, uint32_t nCells, int nb, uint32_t candCap) {
    extern __shared__ __align__(16) unsigned char sel_smem[];
    constexpr int GPB = kThreads / G;
    __shared__ XdSmem s_x[GPB];
    __shared__ uint32_t s_w[32];
    pdl_enter();
    x_barrier(xa);
    const int tid = threadIdx.x, gtid = tid % G, grp = tid / G, lane = tid & 31;
    uint32_t *h = reinterpret_cast<uint32_t *>(sel_smem) + (size_t)grp * nb;
    XdSmem &sx = s_x[grp];
    const uint32_t stride = x_owned_stride(nCells, xa.n);
    const uint32_t *histG = xa.arena[xa.self] + xa.offHistG;
    const int per = (nb + G - 1) / G;          // bins per thread (thread t: bins [t per, (t+1) per), those < nb)
    for (uint32_t c = blockIdx.x * GPB + grp; c < nCells; c += gridDim.x * GPB) {
        const int owner = (int)(c % (uint32_t)xa.n);
        const uint32_t slot = (uint32_t)xa.self * stride + c / (uint32_t)xa.n;
        uint32_t *cntDst = xa.arena[owner] + xa.offRecvCnt + slot;
        if (!lv.active[c]) {
            if (gtid == 0) { ss.flag[c] = 0u; *cntDst = 0u; mr.loc_base[c] = 0u; }
            continue;
        }
        grp_sync<G>();
        for (int i = gtid; i < nb; i += G) h[i] = __ldcg(histG + (size_t)c * nb + i);
        if (gtid == 0) { sx.first = nb; sx.last = -1; sx.base = 0u; sx.end = 0u; sx.cnt = 0u; }
        grp_sync<G>();
        SelTarget tg;
        tg.init(lv.total[c], lv.nleaf[c]);
        uint32_t sum = 0;
        for (int j = 0; j < per; ++j) { const int bb = gtid * per + j; if (bb < nb) sum += h[bb]; }
        uint32_t total;
        const uint32_t excl = grp_scan<G>(sum, s_w, total);
        int myFirst = nb, myLast = -1;
        {
            uint32_t p = excl;
            for (int j = 0; j < per; ++j) {
                const int bb = gtid * per + j;
                if (bb < nb) {
                    const uint32_t pn = p + h[bb];
                    if (tg.diff(pn) > -3) myFirst = min(myFirst, bb);
                    if (tg.diff(p) < 3) myLast = max(myLast, bb);
                    p = pn;
                }
            }
        }
        if (G == 32) {
            myFirst = __reduce_min_sync(0xffffffffu, myFirst);
            myLast = __reduce_max_sync(0xffffffffu, myLast);
        } else {
            if (myFirst < nb) atomicMin(&sx.first, myFirst);
            if (myLast >= 0) atomicMax(&sx.last, myLast);
            __syncthreads();
            myFirst = sx.first; myLast = sx.last;
        }
        const int first = myFirst, last = myLast;
        {
            uint32_t p = excl;
            for (int j = 0; j < per; ++j) {
                const int bb = gtid * per + j;
                if (bb < nb) {
                    if (bb == first) sx.base = p;
                    p += h[bb];
                    if (bb == last) sx.end = p;
                }
            }
        }
        grp_sync<G>();
        const uint32_t base = sx.base, K2 = sx.end - sx.base;
        const bool ok = first <= last && first < nb && last >= 0 && K2 <= candCap;
        // local particles below the candidate bins, from this rank's own row
        uint32_t lb = 0;
        if (ok) for (int i = gtid; i < first; i += G) lb += __ldcg(mr.hist_l + (size_t)c * nb + i);
        lb = __reduce_add_sync(0xffffffffu, lb);
        if (G != 32) {
            __syncthreads();                 // s_w is free again (grp_scan's readers are done)
            if (lane == 0) s_w[tid >> 5] = lb;
            __syncthreads();
            lb = 0;
            for (int w = 0; w < kWarps; ++w) lb += s_w[w];
        }
        if (gtid == 0) {
            ss.bfirst[c] = ok ? (uint32_t)first : 1u; ss.blast[c] = ok ? (uint32_t)last : 0u;
            ss.base[c] = ok ? base : 0u; ss.ncand[c] = ok ? K2 : 0u; ss.flag[c] = ok ? 0u : 1u;
            mr.loc_base[c] = lb;
        }
        if (ok) {
            const uint32_t b = lv.bnd[c], K = lv.bnd[c + 1] - b;
            const float lo = lv.mL[c], scale = sel_scale(lo, lv.mR[c], nb);
            float fLo, fHi;
            sel_bin_bounds((uint32_t)first, (uint32_t)last, nb, fLo, fHi);
            float *dst = reinterpret_cast<float *>(xa.arena[owner] + xa.offRecv + (size_t)slot * mr.slotWords);
            const uint32_t lim = mr.slotWords - 1u;
            // all threads of a warp run the same number of calls (grp_for_each strides uniformly except for the ragged
            // tails, where the ballot below just sees fewer lanes): positions come from a warp-aggregated counter
            grp_for_each<G, 4>(pick_col(lv.axis[c], x, y, z) + b, K, gtid, [&](float v) {
                const float t = fmaxf(__fmul_rn(__fsub_rn(v, lo), scale), 0.f);
                const bool keep = t >= fLo && t < fHi;
                const unsigned act = __activemask();
                const unsigned m = __ballot_sync(act, keep);
                if (m) {
                    const int leader = __ffs(m) - 1;
                    uint32_t p0 = 0;
                    if (lane == leader) p0 = atomicAdd(&sx.cnt, (uint32_t)__popc(m));
                    p0 = __shfl_sync(act, p0, leader);
                    if (keep) {
                        const uint32_t p = p0 + (uint32_t)__popc(m & ((1u << lane) - 1u));
                        if (p < lim) dst[p] = v;
                    }
                }
            });
        }
        grp_sync<G>();
        if (gtid == 0) *cntDst = ok ? sx.cnt : 0u;
    }
}

// ---- warp-per-cell versions for the deepest levels (a few hundred to a couple of thousand particles per cell and
//      rank).  At that size a cell costs a handful of dependent memory latencies, not bandwidth: the per-cell scalars of
//      the NEXT cell are fetched while the current one is processed, and everything the current cell needs (its rows,
//      its particles) is requested at once - the particles land in the warp's shared-memory staging area. ----
struct CellMeta {
    uint32_t b, e, total, active;
    int nleaf, axis;
    float L, R;
};
__device__ __forceinline__ CellMeta cell_meta(const LevelState &lv, uint32_t c, uint32_t nCells) {
    CellMeta m;
    m.b = m.e = m.total = m.active = 0u; m.nleaf = 1; m.axis = 0; m.L = m.R = 0.f;
    if (c < nCells) {
        m.b = lv.bnd[c]; m.e = lv.bnd[c + 1]; m.total = lv.total[c]; m.active = lv.active[c];
        m.nleaf = lv.nleaf[c]; m.axis = lv.axis[c]; m.L = lv.mL[c]; m.R = lv.mR[c];
    }
    return m;
}
__device__ __forceinline__ void cp_async4(void *smem, const void *gmem) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(sa), "l"(gmem) : "memory");
}
constexpr uint32_t kXdStage = 2048;       // particles of one cell a warp stages (larger cells are read in place)

// HIST, one warp per cell.  dynamic shared memory: kWarps x nb words
__global__ void __launch_bounds__(kThreads) k_xd_hist_warp(const float *__restrict__ x, const float *__restrict__ y,
                                                           const float *__restrict__ z, LevelState lv, uint32_t *__restrict__ hist_l,
                                                           uint32_t nCells, int nb) {
    extern __shared__ __align__(16) unsigned char sel_smem[];
    pdl_enter();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *h = reinterpret_cast<uint32_t *>(sel_smem) + (size_t)warp * nb;
    const float nbm1 = (float)(nb - 1);
    const uint32_t stride = gridDim.x * kWarps;
    uint32_t c = blockIdx.x * kWarps + warp;
    CellMeta cur = cell_meta(lv, c, nCells);
    while (c < nCells) {
        const CellMeta nxt = cell_meta(lv, c + stride, nCells);
        if (cur.active) {       // warp-uniform
            const float lo = cur.L, scale = sel_scale(lo, cur.R, nb);
            for (int i = lane; i < nb; i += 32) h[i] = 0u;
            __syncwarp();
            grp_for_each<32, 4>(pick_col(cur.axis, x, y, z) + cur.b, cur.e - cur.b, lane, [&](float v) {
                const float t = fminf(fmaxf(__fmul_rn(__fsub_rn(v, lo), scale), 0.f), nbm1);      // == sel_bin(v, lo, scale, nb)
                atomicAdd(&h[__float2int_rz(t)], 1u);
            });
            __syncwarp();
            for (int i = lane; i < nb; i += 32) hist_l[(size_t)c * nb + i] = h[i];
            __syncwarp();
        }
        c += stride;
        cur = nxt;
    }
}

// RESOLVE + COMPACT, one warp per cell.  dynamic shared memory per warp: rowG[nb] | rowL[nb] | vals[kXdStage]
__host__ __device__ inline size_t xd_compact_warp_smem(int nb) { return (size_t)kWarps * (2u * (size_t)nb + kXdStage) * 4u; }
__global__ void __launch_bounds__(kThreads) k_xd_compact_warp(const float *__restrict__ x, const float *__restrict__ y,
                                                              const float *__restrict__ z, LevelState lv, SelState ss, SelMrState mr,
                                                              XArena xa, uint32_t nCells, int nb, uint32_t candCap) {
    extern __shared__ __align__(16) unsigned char sel_smem[];
    __shared__ uint32_t s_cnt[kWarps], s_base[kWarps], s_end[kWarps];
    pdl_enter();
    x_barrier(xa);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *rowG = reinterpret_cast<uint32_t *>(sel_smem) + (size_t)warp * (2u * (size_t)nb + kXdStage);
    uint32_t *rowL = rowG + nb;
    float *vals = reinterpret_cast<float *>(rowL + nb);
    const uint32_t ostride = x_owned_stride(nCells, xa.n);
    const uint32_t *histG = xa.arena[xa.self] + xa.offHistG;
    const int per = (nb + 31) / 32;
    const uint32_t stride = gridDim.x * kWarps;
    uint32_t c = blockIdx.x * kWarps + warp;
    CellMeta cur = cell_meta(lv, c, nCells);
    while (c < nCells) {
        const CellMeta nxt = cell_meta(lv, c + stride, nCells);
        const int owner = (int)(c % (uint32_t)xa.n);
        const uint32_t slot = (uint32_t)xa.self * ostride + c / (uint32_t)xa.n;
        uint32_t *cntDst = xa.arena[owner] + xa.offRecvCnt + slot;
        if (!cur.active) {
            if (lane == 0) { ss.flag[c] = 0u; *cntDst = 0u; mr.loc_base[c] = 0u; }
        } else {
            const uint32_t K = cur.e - cur.b;
            const float *col = pick_col(cur.axis, x, y, z) + cur.b;
            const bool staged = K + 4u <= kXdStage;
            // everything this cell needs, requested at once (rows: nb is a multiple of 32 words, 16-byte pieces; the
            // particles start at any 4-byte address: the staged copy begins `mis` floats into the buffer so that shared
            // and global addresses agree mod 16, head and tail go in 4-byte pieces)
            for (int i = lane; i < nb / 4; i += 32) {
                cp_async16(reinterpret_cast<uint4 *>(rowG) + i, reinterpret_cast<const uint4 *>(histG + (size_t)c * nb) + i);
                cp_async16(reinterpret_cast<uint4 *>(rowL) + i, reinterpret_cast<const uint4 *>(mr.hist_l + (size_t)c * nb) + i);
            }
            const uint32_t mis = (uint32_t)((reinterpret_cast<uintptr_t>(col) >> 2) & 3u);
            float *sv = vals + mis;         // sv[i] <-> col[i]
            if (staged) {
                const uint32_t head = mis ? min(4u - mis, K) : 0u, body4 = (K - head) / 4u, tail0 = head + body4 * 4u;
                for (uint32_t i = lane; i < body4; i += 32u)
                    cp_async16(reinterpret_cast<float4 *>(sv + head) + i, reinterpret_cast<const float4 *>(col + head) + i);
                if ((uint32_t)lane < head) cp_async4(sv + lane, col + lane);
                if (tail0 + (uint32_t)lane < K) cp_async4(sv + tail0 + lane, col + tail0 + lane);
            }
            cp_async_commit();
            if (lane == 0) s_cnt[warp] = 0u;
            cp_async_wait<0>();
            __syncwarp();
            // ---- resolve (same arithmetic on every rank: the row is the global one) ----
            SelTarget tg;
            tg.init(cur.total, cur.nleaf);
            uint32_t sum = 0;
            for (int j = 0; j < per; ++j) { const int bb = lane * per + j; if (bb < nb) sum += rowG[bb]; }
            uint32_t total;
            const uint32_t excl = grp_scan<32>(sum, nullptr, total);
            int myFirst = nb, myLast = -1;
            {
                uint32_t p = excl;
                for (int j = 0; j < per; ++j) {
                    const int bb = lane * per + j;
                    if (bb < nb) {
                        const uint32_t pn = p + rowG[bb];
                        if (tg.diff(pn) > -3) myFirst = min(myFirst, bb);
                        if (tg.diff(p) < 3) myLast = max(myLast, bb);
                        p = pn;
                    }
                }
            }
            const int first = __reduce_min_sync(0xffffffffu, myFirst), last = __reduce_max_sync(0xffffffffu, myLast);
            {
                uint32_t p = excl;
                for (int j = 0; j < per; ++j) {
                    const int bb = lane * per + j;
                    if (bb < nb) {
                        if (bb == first) s_base[warp] = p;
                        p += rowG[bb];
                        if (bb == last) s_end[warp] = p;
                    }
                }
            }
            __syncwarp();
            const bool okBins = first <= last && first < nb && last >= 0;
            const uint32_t base = okBins ? s_base[warp] : 0u, K2 = okBins ? s_end[warp] - s_base[warp] : 0u;
            const bool ok = okBins && K2 <= candCap;
            uint32_t lb = 0;
            if (ok) for (int i = lane; i < first; i += 32) lb += rowL[i];
            lb = __reduce_add_sync(0xffffffffu, lb);
            if (lane == 0) {
                ss.bfirst[c] = ok ? (uint32_t)first : 1u; ss.blast[c] = ok ? (uint32_t)last : 0u;
                ss.base[c] = ok ? base : 0u; ss.ncand[c] = ok ? K2 : 0u; ss.flag[c] = ok ? 0u : 1u;
                mr.loc_base[c] = lb;
            }
            if (ok) {
                const float lo = cur.L, scale = sel_scale(lo, cur.R, nb);
                float fLo, fHi;
                sel_bin_bounds((uint32_t)first, (uint32_t)last, nb, fLo, fHi);
                float *dst = reinterpret_cast<float *>(xa.arena[owner] + xa.offRecv + (size_t)slot * mr.slotWords);
                const uint32_t lim = mr.slotWords - 1u;
                // all lanes take part in every round (ballot), out-of-range lanes just keep nothing
                const uint32_t rounds = (K + 31u) / 32u;
                uint32_t p0 = 0;
                for (uint32_t r = 0; r < rounds; ++r) {
                    const uint32_t i = r * 32u + (uint32_t)lane;
                    float v = 0.f;
                    bool keep = false;
                    if (i < K) {
                        v = staged ? sv[i] : __ldg(col + i);
                        const float t = fmaxf(__fmul_rn(__fsub_rn(v, lo), scale), 0.f);
                        keep = t >= fLo && t < fHi;
                    }
                    const unsigned m = __ballot_sync(0xffffffffu, keep);
                    if (keep) {
                        const uint32_t p = p0 + (uint32_t)__popc(m & ((1u << lane) - 1u));
                        if (p < lim) dst[p] = v;
                    }
                    p0 += (uint32_t)__popc(m);
                }
                if (lane == 0) s_cnt[warp] = p0;
            }
            __syncwarp();
            if (lane == 0) *cntDst = ok ? s_cnt[warp] : 0u;
            __syncwarp();
        }
        c += stride;
        cur = nxt;
    }
}

// ---- FINISH by the owner ----
// result record pushed to every rank (8 words): mL, mR, meta, nleft_g, nloc (of the receiving rank), 0, 0, 0
constexpr uint32_t kXMetaFound = 1u << 8, kXMetaFlag = 1u << 16;

__device__ __forceinline__ void x_push_record(const XArena &xa, uint32_t c, int target, float L, float R, uint32_t meta,
                                              uint32_t nleft, uint32_t nloc) {
    uint4 *dst = reinterpret_cast<uint4 *>(xa.arena[target] + xa.offRes + (size_t)c * 8u);
    dst[0] = make_uint4(__float_as_uint(L), __float_as_uint(R), meta, nleft);
    dst[1] = make_uint4(nloc, 0u, 0u, 0u);
}

// block per owned cell (candidates beyond what a warp stages): the block search of orb_select.cuh
// (capBig > cap with `scratch`: a cell whose candidates do not fit shared memory - the top levels of a 2^30-particle
//  build, 131072 particles per bin - is searched in a global scratch array instead: the block search only needs a
//  pointer it can read a few times, and a few hundred KB stay in L2)
__global__ void __launch_bounds__(1024) k_xf_finish_block(LevelState lv, SelState ss, SelMrState mr, XArena xa, uint32_t nCells,
                                                          int nb1, uint32_t cap, int *__restrict__ err, float *__restrict__ scratch,
                                                          uint32_t capBig) {
    extern __shared__ __align__(16) unsigned char sel_smem[];
    float *sbuf = reinterpret_cast<float *>(sel_smem);
    uint32_t *hist2 = reinterpret_cast<uint32_t *>(sbuf + cap + 4);
    float *amb = reinterpret_cast<float *>(hist2 + kSelBins2);
    __shared__ SelSearchSmem sm;
    __shared__ uint32_t s_cnt[kMaxPeers], s_nl[kMaxPeers];
    pdl_enter();
    x_barrier(xa);
    const int tid = threadIdx.x, lane = tid & 31, nThreads = (int)blockDim.x;
    const uint32_t stride = x_owned_stride(nCells, xa.n);
    const uint32_t nOwned = nCells > (uint32_t)xa.self ? (nCells - (uint32_t)xa.self + (uint32_t)xa.n - 1u) / (uint32_t)xa.n : 0u;
    const uint32_t *recvCnt = xa.arena[xa.self] + xa.offRecvCnt;
    const float *recv = reinterpret_cast<const float *>(xa.arena[xa.self] + xa.offRecv);
    for (uint32_t oi = blockIdx.x; oi < nOwned; oi += gridDim.x) {
        const uint32_t c = oi * (uint32_t)xa.n + (uint32_t)xa.self;
        __syncthreads();
        const uint32_t act = lv.active[c];
        const uint32_t flg = __ldcg(&ss.flag[c]), K = __ldcg(&ss.ncand[c]), base = __ldcg(&ss.base[c]);
        const uint32_t bf = __ldcg(&ss.bfirst[c]), bl = __ldcg(&ss.blast[c]);
        const float L = lv.mL[c], R = lv.mR[c];
        if (!act) continue;                          // (apply skips inactive cells without looking at their record)
        if (tid < xa.n) { s_cnt[tid] = __ldcg(recvCnt + (size_t)tid * stride + oi); s_nl[tid] = 0u; }
        __syncthreads();
        uint32_t sum = 0;
        bool over = false;
        for (int r = 0; r < xa.n; ++r) { over |= s_cnt[r] > mr.slotWords - 1u; sum += s_cnt[r]; }
        bool flagged = flg != 0u || over;
        const bool big = K > cap;
        if (!flagged && (sum != K || (big && (!scratch || K > capBig)))) {     // cannot happen: all ranks bin with the same function and resolve the same rows
            if (tid == 0) atomicExch(err, ORB_ERR_STATE);
            flagged = true;
        }
        float *vals = big ? scratch + (size_t)oi * capBig : sbuf;
        if (!flagged) {
            uint32_t off = 0;
            for (int r = 0; r < xa.n; ++r) {
                const uint32_t n = s_cnt[r];
                const float4 *s4 = reinterpret_cast<const float4 *>(recv + ((size_t)r * stride + oi) * mr.slotWords);
                const uint32_t n4 = (n + 3u) >> 2;
                for (uint32_t i = tid; i < n4; i += nThreads) {
                    const float4 q = __ldcg(s4 + i);
                    const uint32_t k = 4u * i;
                    float *d = vals + off + k;
                    d[0] = q.x;
                    if (k + 1u < n) d[1] = q.y;
                    if (k + 2u < n) d[2] = q.z;
                    if (k + 3u < n) d[3] = q.w;
                }
                off += n;
            }
            __syncthreads();
            flagged = !sel_block_search_core(vals, K, base, 1, L, sel_scale(L, R, nb1), nb1, (int)bf, (int)bl, hist2, amb, lv, c, sm);
            __syncthreads();
        }
        if (!flagged) {
            // every rank's candidates left of the final cut (getCut() of the final margins: the found cut as well as
            // the capped cell's cut)
            const float cutf = mid_cut(sm.resL, sm.resR);
            uint32_t off = 0;
            for (int r = 0; r < xa.n; ++r) {
                const uint32_t n = s_cnt[r];
                uint32_t m = 0;
                for (uint32_t i = tid; i < n; i += nThreads) m += (vals[off + i] < cutf) ? 1u : 0u;
                m = __reduce_add_sync(0xffffffffu, m);
                if (lane == 0 && m) atomicAdd(&s_nl[r], m);
                off += n;
            }
            __syncthreads();
        }
        if (tid < xa.n) {
            if (flagged) x_push_record(xa, c, tid, L, R, kXMetaFlag, 0u, 0u);
            else x_push_record(xa, c, tid, sm.resL, sm.resR, (uint32_t)sm.resIt | (sm.resFnd ? kXMetaFound : 0u), sm.resNleft, s_nl[tid]);
        }
    }
}

// warp per owned cell (deep levels: a few hundred candidates): the replay counts over the staged candidates directly
constexpr uint32_t kXWarpCap = 2048;     // candidates one warp stages
__global__ void __launch_bounds__(kThreads) k_xf_finish_warp(LevelState lv, SelState ss, SelMrState mr, XArena xa, uint32_t nCells,
                                                             int nb1, uint32_t cap /* <= kXWarpCap */, int *__restrict__ err) {
    extern __shared__ __align__(16) unsigned char sel_smem[];
    pdl_enter();
    x_barrier(xa);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float *sbuf = reinterpret_cast<float *>(sel_smem) + (size_t)warp * (cap + 4u);
    const uint32_t stride = x_owned_stride(nCells, xa.n);
    const uint32_t nOwned = nCells > (uint32_t)xa.self ? (nCells - (uint32_t)xa.self + (uint32_t)xa.n - 1u) / (uint32_t)xa.n : 0u;
    const uint32_t *recvCnt = xa.arena[xa.self] + xa.offRecvCnt;
    const float *recv = reinterpret_cast<const float *>(xa.arena[xa.self] + xa.offRecv);
    for (uint32_t oi = blockIdx.x * kWarps + warp; oi < nOwned; oi += gridDim.x * kWarps) {
        const uint32_t c = oi * (uint32_t)xa.n + (uint32_t)xa.self;
        if (!lv.active[c]) continue;                 // warp-uniform
        const uint32_t flg = __ldcg(&ss.flag[c]), K = __ldcg(&ss.ncand[c]), base = __ldcg(&ss.base[c]);
        const int bf = (int)__ldcg(&ss.bfirst[c]), bl = (int)__ldcg(&ss.blast[c]);
        const float L0 = lv.mL[c], R0 = lv.mR[c];
        const uint32_t myCnt = lane < xa.n ? __ldcg(recvCnt + (size_t)lane * stride + oi) : 0u;
        const bool over = __any_sync(0xffffffffu, myCnt > mr.slotWords - 1u);
        const uint32_t sum = __reduce_add_sync(0xffffffffu, myCnt);
        bool flagged = flg != 0u || over;
        if (!flagged && (sum != K || K > cap)) {
            if (lane == 0) atomicExch(err, ORB_ERR_STATE);
            flagged = true;
        }
        float L = L0, R = R0;
        int it = 0;
        bool fnd = false;
        uint32_t nleft = 0, nloc = 0;
        if (!flagged) {
            __syncwarp();
            uint32_t off = 0;
            for (int r = 0; r < xa.n; ++r) {
                const uint32_t n = __shfl_sync(0xffffffffu, myCnt, r);
                const float *src = recv + ((size_t)r * stride + oi) * mr.slotWords;
                for (uint32_t i = lane; i < n; i += 32u) sbuf[off + i] = __ldcg(src + i);
                off += n;
            }
            __syncwarp();
            SelTarget tg;
            tg.init(lv.total[c], lv.nleaf[c]);
            const float lo1 = L0, scale1 = sel_scale(L0, R0, nb1);
            auto count_below = [&](float cut) {
                uint32_t n = 0;
                for (uint32_t i = lane; i < K; i += 32u) n += (sbuf[i] < cut) ? 1u : 0u;
                return __reduce_add_sync(0xffffffffu, n);
            };
            // replay of orbit.cpp:149-232 (every lane computes the same scalars)
            while (it < kMaxIter) {
                const float cut = mid_cut(L, R);
                const int b1 = sel_bin(cut, lo1, scale1, nb1);
                int dec = b1 < bf ? -1 : (b1 > bl ? 1 : 0);
                ++it;
                if (dec == 0) {
                    const uint32_t cnt = base + count_below(cut);
                    const int d = tg.diff(cnt);
                    if (abs(d) < 3) { fnd = true; nleft = cnt; break; }       // orbit.cpp:208
                    dec = d > 0 ? 1 : -1;
                }
                if (dec > 0) R = cut; else L = cut;                            // orbit.cpp:219,227
            }
            const float cutf = mid_cut(L, R);
            if (!fnd) {     // capped cell: the count at getCut() of the last margins (never counted by the loop)
                const int b1 = sel_bin(cutf, lo1, scale1, nb1);
                if (b1 >= bf && b1 <= bl) nleft = base + count_below(cutf);
                else flagged = true;        // outside the candidate bins: its exact count is not known here
            }
            if (!flagged) {
                uint32_t off2 = 0;
                for (int r = 0; r < xa.n; ++r) {
                    const uint32_t n = __shfl_sync(0xffffffffu, myCnt, r);
                    uint32_t m = 0;
                    for (uint32_t i = lane; i < n; i += 32u) m += (sbuf[off2 + i] < cutf) ? 1u : 0u;
                    m = __reduce_add_sync(0xffffffffu, m);
                    if (lane == r) nloc = m;
                    off2 += n;
                }
            }
        }
        if (lane < xa.n) {
            if (flagged) x_push_record(xa, c, lane, L0, R0, kXMetaFlag, 0u, 0u);
            else x_push_record(xa, c, lane, L, R, (uint32_t)it | (fnd ? kXMetaFound : 0u), nleft, nloc);
        }
        __syncwarp();
    }
}

// ---- APPLY: the owners' records into this rank's level state; statistics; 1 + cells flagged to the host ----
__global__ void __launch_bounds__(kThreads) k_xa_apply(LevelState lv, SelState ss, SelCtl sc, SelMrState mr, XArena xa, uint32_t nCells,
                                                        int hbmPasses) {
    __shared__ uint32_t s_flag, s_unf;
    __shared__ unsigned long long s_p, s_q;
    __shared__ int s_it;
    pdl_enter();
    x_barrier(xa);
    const int tid = threadIdx.x, lane = tid & 31;
    if (tid == 0) { s_flag = 0u; s_unf = 0u; s_p = 0ull; s_q = 0ull; s_it = 0; }
    __syncthreads();
    if (blockIdx.x == 0 && tid == 0) atomicAdd(sc.passes_out, hbmPasses);
    const uint32_t c = blockIdx.x * blockDim.x + tid;
    uint32_t nflag = 0, nunf = 0;
    unsigned long long np1 = 0, np2 = 0;
    int itMax = 0;
    if (c < nCells && lv.active[c]) {
        const uint4 *rec = reinterpret_cast<const uint4 *>(xa.arena[xa.self] + xa.offRes + (size_t)c * 8u);
        const uint4 a = __ldcg(rec), b = __ldcg(rec + 1);
        if (a.z & kXMetaFlag) {
            ss.flag[c] = 1u;
            nflag = 1;
        } else {
            const int it = (int)(a.z & 0xffu);
            const uint32_t fnd = (a.z & kXMetaFound) ? 1u : 0u;
            lv.mL[c] = __uint_as_float(a.x); lv.mR[c] = __uint_as_float(a.y); lv.iter[c] = it;
            lv.found[c] = fnd;
            lv.active[c] = 0u;
            lv.nleft_g[c] = a.w;
            lv.nleft_l[c] = mr.loc_base[c] + b.x;
            ss.flag[c] = 0u;
            const unsigned long long np = (unsigned long long)(lv.bnd[c + 1] - lv.bnd[c]);
            np1 = np * (unsigned long long)hbmPasses;
            np2 = np * (unsigned long long)it;
            itMax = it;
            nunf = fnd ? 0u : 1u;
        }
    }
    nflag = __reduce_add_sync(0xffffffffu, nflag);
    nunf = __reduce_add_sync(0xffffffffu, nunf);
    itMax = __reduce_max_sync(0xffffffffu, itMax);
    for (int o = 16; o; o >>= 1) {
        np1 += __shfl_xor_sync(0xffffffffu, np1, o);
        np2 += __shfl_xor_sync(0xffffffffu, np2, o);
    }
    if (lane == 0) {
        if (nflag) atomicAdd(&s_flag, nflag);
        if (nunf) atomicAdd(&s_unf, nunf);
        if (np1) atomicAdd(&s_p, np1);
        if (np2) atomicAdd(&s_q, np2);
        if (itMax) atomicMax(&s_it, itMax);
    }
    __syncthreads();
    if (tid == 0) {
        if (s_flag) atomicAdd(ss.n_flagged, s_flag);
        if (s_unf) atomicAdd(sc.n_unfound_out, s_unf);
        if (s_p) atomicAdd(sc.active_particles, s_p);
        if (s_q) atomicAdd(sc.active_particles + 1, s_q);
        if (s_it) atomicMax(sc.level_iters, s_it);
        __threadfence();
        if (atomicAdd(mr.done, 1u) == gridDim.x - 1u) {
            __threadfence();
            *mr.h_status = *((volatile uint32_t *)ss.n_flagged) + 1u;
        }
    }
}

}  // namespace orb
