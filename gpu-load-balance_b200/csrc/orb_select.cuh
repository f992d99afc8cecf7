// orb_select.cuh — selection-based cut search (byte-reducing search that keeps the cut sequence).
//
// The reference finds each cell's cut with up to 32 bisection steps, every step one count `#{x < cut}` over the cell
// (orbit.cpp:146-232, countLeft.cpp:9-42).  A step uses the count only through
//     diff = (int)((float)cnt - (float)n * ratio)          (orbit.cpp:204-205)
// which is monotone in cnt.  So the whole loop can be replayed without re-reading the particles:
//   1. HIST     one pass over the cut-axis column bins every particle with a monotone function of its coordinate
//               (`sel_bin`); the prefix sums bracket `#{x < cut}` for ANY cut by the bin the cut itself falls in;
//   2. RESOLVE  (inside the COMPACT pass, by every block as it enters a cell) per cell, the bins whose count range can still contain |diff| < 3 or flip the sign of diff are the
//               candidate bins [first, last] (normally one bin); every cut outside them is decided by the prefix sums;
//   3. COMPACT  a second pass gathers the particles of the candidate bins into a dense per-cell list (in the idle
//               ping-pong column);
//   4. FINISH   one block per cell stages the candidates in shared memory, refines once more with a 2048-bin
//               histogram, keeps the few still-ambiguous values and replays the reference's float decisions for all
//               (up to 32) steps with exact counts  base + #{ambiguous < cut}.
// Cells that fit in shared memory (deep levels) skip 1-3: one block reads the cell once and runs step 4 on it.
// Result (margins, iterations, foundCut, nLeft) is bit-identical to the literal loop; tests/test_select_model.py
// restates the algorithm in numpy and checks it against the literal loop, tests/test_gpu_parity.py checks this file
// against the oracle.  Cells the method cannot finish (candidates beyond shared memory because of massive ties, a
// capped cell whose final cut leaves the candidate bins) are flagged and left untouched for the iterative path
// (k_level_persistent / k_count_*), which then runs for them alone.  With several ranks the histogram rows are
// all-reduced and the candidates all-gathered (two exchanges per level, see k_selmr_* at the end of this file).
#pragma once
#include "orb_kernels.cuh"

namespace orb {

constexpr int kSelBinsMin = 512;      // bins per cell of the HIST pass (power of two, chosen per level from the cell size)
constexpr int kSelBinsMax = 8192;
constexpr int kSelBins2 = 2048;       // bins of the in-block refinement
constexpr int kSelAmbCap = 2048;      // ambiguous values kept for the replay
constexpr int kSelWarpStage = 640;    // COMPACT: staging slots per warp (a tile adds at most 512 per warp)
constexpr int kSelHist = 0, kSelCompact = 1;

struct SelState {
    uint32_t *hist;       // [nCells][nb1] bin counts of the HIST pass (zeroed per level)
    uint32_t *bfirst;     // [nCells] candidate bins [bfirst, blast]; bfirst > blast: nothing to gather
    uint32_t *blast;
    uint32_t *base;       // [nCells] particles in bins below bfirst
    uint32_t *ncand;      // [nCells] particles in the candidate bins
    uint32_t *cursor;     // [nCells] COMPACT: fill level of the cell's candidate list; FINISH leaves it zero
    uint32_t *flag;       // [nCells] 1: left to the iterative path
    uint32_t *n_flagged;  // this level's count of flagged cells (gate of the iterative fallback)
    float *vlo, *vhi;     // [nCells] value bounds of the candidate bins (sel_value_bounds), published by k_sel_resolve; may be null
};

struct SelCtl {           // statistics of the level (same meaning as PassCtl / LevelCtl)
    unsigned long long *active_particles;
    int32_t *level_iters;
    int32_t *passes_out;
    uint32_t *n_unfound_out;
};

// (sel_scale / sel_bin, the bin function, live in orb_kernels.cuh: the partition bins with them as well)

// the reference's decision value for a count (orbit.cpp:204-205), literal float arithmetic
struct SelTarget {
    float prod;
    // ratio = (float)(ceil(nLeafCells / 2.0) / nLeafCells) (orbit.cpp:204: double division, then rounded to float).
    // Evaluated as ONE correctly rounded float division, which gives the same bits: numerator a = (n+1)/2 and
    // denominator n are integers below 2^21, so a/n is either exactly a float rounding midpoint or at least
    // 2^-46 (relative) away from one - the intermediate rounding to double (2^-53) can never move it across.
    // (Every thread of a block evaluates this; FP64 division would serialise on the narrow FP64 pipe.)
    __device__ __forceinline__ void init(uint32_t total, int nleaf) {
        const float ratio = __fdiv_rn((float)((nleaf + 1) >> 1), (float)nleaf);
        prod = __fmul_rn(__uint2float_rn(total), ratio);
    }
    __device__ __forceinline__ int diff(uint32_t cnt) const { return __float2int_rz(__fsub_rn(__uint2float_rn(cnt), prod)); }
};

// ---- sampled rows (single rank) ----
// The HIST pass may bin only a SAMPLE of a cell (every S-th tile, or every S-th 512-byte piece of a cell that one block
// searches): the rows then only ESTIMATE the prefix counts.  RESOLVE widens the candidate bins by z standard deviations
// of the estimate; the pass that gathers the candidates (which reads every particle anyway) counts the particles below
// the candidate bins EXACTLY, and the search is entered only if the exact numbers prove the bracket:
//     (first == 0 || diff(base) <= -3)  &&  (last == nb - 1 || diff(base + K) >= 3)
// - the same two facts the exact rows establish (sel_resolve_cell).  The result therefore never depends on the sample;
// a bracket that fails is searched again with exact rows (k_sel_percell) or left to the iterative search (k_sel_finish).
struct SelSampleEst {
    float ns, sc, z, nTot;
    __device__ __forceinline__ void init(uint32_t sampleTotal, uint32_t cellTotal, float zIn) {
        ns = (float)sampleTotal; nTot = (float)cellTotal; z = zIn;
        sc = sampleTotal ? __fdiv_rn(nTot, ns) : 0.f;
    }
    // bound on the cell's exact prefix count, given the sample's prefix count p
    __device__ __forceinline__ uint32_t bound(uint32_t p, bool upper) const {
        if (!(ns > 0.f)) return upper ? (uint32_t)nTot : 0u;
        const float pf = (float)p;
        const float sd = sc * sqrtf(fmaxf(pf * (ns - pf), 0.f) / ns);
        const float m = z * sd + 2.f * sc + 8.f;
        const float v = fminf(fmaxf(upper ? pf * sc + m : pf * sc - m, 0.f), nTot);
        return (uint32_t)v;
    }
};

// smallest x in (lo, hi] with pred(x), for a monotone predicate with pred(hi) true and pred(lo) false (or lo outside the
// domain); called by a whole warp, 32 probes per round
template <typename P>
__device__ __forceinline__ uint32_t sel_warp_first_true(uint32_t lo, uint32_t hi, P pred) {
    const uint32_t lane = threadIdx.x & 31u;
    while (hi - lo > 1u) {
        const uint32_t span = hi - lo, step = (span + 31u) / 32u;
        const uint64_t xx = (uint64_t)lo + (uint64_t)step * (lane + 1u);
        const uint32_t x = xx >= (uint64_t)hi ? hi : (uint32_t)xx;
        const unsigned m = __ballot_sync(0xffffffffu, x == hi || pred(x));
        const int j = __ffs(m) - 1;                           // (lane 31 probes hi: m != 0)
        const uint32_t hj = __shfl_sync(0xffffffffu, x, j);
        const uint32_t lj = j ? __shfl_sync(0xffffffffu, x, j - 1) : lo;
        hi = hj; lo = lj;
    }
    return hi;
}
// The sampled RESOLVE without a square root and a division per bin: the two bounds are monotone in the sample prefix
// where it matters (away from the very ends), so "upper bound reaches the target" / "lower bound still below it" turn
// into two critical sample prefixes pA, pB, found once per cell by a warp.  A bin then costs two integer compares.
// (Where the bounds are not monotone the candidate bins may come out differently - the bracket is proven afterwards
// with exact counts whatever they are.)
__device__ __forceinline__ void sel_sample_crit(const SelSampleEst &se, const SelTarget &tg, uint32_t sampleTotal, uint32_t &pA, uint32_t &pB) {
    // pA = smallest p in [0, ns] with diff(upper(p)) > -3   (search over y = p + 1 so that "none below" is y = 0)
    pA = sel_warp_first_true(0u, sampleTotal + 1u, [&](uint32_t y) { return tg.diff(se.bound(y - 1u, true)) > -3; }) - 1u;
    // pB = largest p in [0, ns] with diff(lower(p)) < 3 = (smallest p in (0, ns + 1] where it fails) - 1
    pB = sel_warp_first_true(0u, sampleTotal + 1u, [&](uint32_t x) { return !(tg.diff(se.bound(x, false)) < 3); }) - 1u;
}

// COMPACT with private regions (single rank): block b appends the candidates it meets to ITS OWN region of the idle
// column (the particles [chunkStart, chunkEnd) it streams - always room, no global atomic, contiguous per visited cell)
// and leaves one record per visited cell.  FINISH gathers a cell's pieces from the blocks whose chunks overlap it; the
// exact count below the candidate bins is the sum of the records' `below`.
constexpr int kSelMaxVisit = 4;
constexpr int kSelMaxPieces = 1024;    // chunks one cell may overlap (pieces FINISH gathers)
struct SelVisit { uint32_t cell, off, cnt, below; };       // off: absolute index of the piece in the idle column
struct SelVisitRec { uint32_t n, pad_[3]; SelVisit v[kSelMaxVisit]; };
struct SelPriv {
    int sampleS;            // HIST: bin every sampleS-th tile only; COMPACT / FINISH: the rows are sampled (> 1)
    float z;
    SelVisitRec *visits;    // != null: private regions + visit records (one per COMPACT block)
    uint32_t chunk;         // particles per COMPACT block (FINISH: to find the blocks that overlap a cell)
    int useBounds;          // COMPACT with preResolved: ss.vlo / ss.vhi hold the cells' value bounds (k_sel_resolve ran)
    // parallel FINISH (k_sel_fine / k_sel_fin_a / k_sel_gather / k_sel_fin_b): the cells' FINE histograms (kSelBins2 bins
    // over the candidate range, bin function lo2 / sc2 published by k_sel_resolve)
    uint32_t *fine;         // [nCells][kSelBins2], null: one-block FINISH (k_sel_finish)
    float *lo2, *sc2;       // [nCells]
};

// block-wide exclusive scan of one value per thread (any block size up to 1024); total in `total`
__device__ __forceinline__ uint32_t sel_block_scan(uint32_t v, uint32_t *s_w /*[32]*/, uint32_t &total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nWarps = (int)(blockDim.x >> 5);
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
    }
    __syncthreads();                       // s_w may still be read from a previous call
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    uint32_t off = 0, tot = 0;
    for (int w = 0; w < nWarps; ++w) {
        const uint32_t x = s_w[w];
        if (w < warp) off += x;
        tot += x;
    }
    total = tot;
    return off + incl - v;
}

// append a thread's kept values (bit j of keep <-> v[j]) to its warp's staging region; *wN = fill level of the region
template <int NV>
__device__ __forceinline__ void sel_warp_append(const float (&v)[NV], unsigned keep, float *stageWarp, uint32_t *wN) {
    const int lane = threadIdx.x & 31;
    const unsigned any = __ballot_sync(0xffffffffu, keep != 0u);
    if (!any) return;
    const unsigned mine = __popc(keep);
    unsigned incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
    }
    const unsigned wtot = __shfl_sync(0xffffffffu, incl, 31);
    const uint32_t start = *wN;
    float *dst = stageWarp + start + (incl - mine);
#pragma unroll
    for (int j = 0; j < NV; ++j)
        if (keep & (1u << j)) dst[__popc(keep & ((1u << j) - 1u))] = v[j];
    __syncwarp();
    if (lane == 0) *wN = start + wtot;
    __syncwarp();
}

// =====================================================================================
// RESOLVE of one cell by one block of kThreads threads: prefix sums of the cell's histogram -> candidate bins
// [first, last] (bins b with NOT(diff(P[b+1]) <= -3) ... diff(P[b]) < 3), base = P[first], ncand = P[last+1] - base.
// Block-uniform call from the COMPACT pass when a block enters a cell (every block that touches the cell computes the
// same result; the block holding the cell's first particle publishes it for FINISH).  `hbuf`: nb1 words of shared
// memory.  Returns the candidate bins (first > last: nothing to gather, the cell is left to the iterative search).
// =====================================================================================
struct SelResolveSmem {
    uint32_t w[32];
    int first, last;
    uint32_t base, end;
    uint32_t pA, pB;          // sampled rows: critical sample prefixes (sel_sample_crit)
};
__device__ __forceinline__ void sel_resolve_cell(const LevelState &lv, const SelState &ss, uint32_t c, int nb1, uint32_t candCap,
                                                 bool publish, uint32_t *hbuf, SelResolveSmem &rs, uint32_t &bfOut, uint32_t &blOut,
                                                 bool preloaded = false /* hbuf already holds the row (barrier done by the caller) */,
                                                 bool countFlag = true /* false: the caller counts flagged cells elsewhere */,
                                                 int sampleS = 1 /* > 1: the row is a sample (see SelSampleEst) */, float sampleZ = 0.f) {
    const int tid = threadIdx.x;
    const int per = nb1 / kThreads;          // 2 .. 32
    if (!preloaded) __syncthreads();
    if (tid == 0) { rs.first = nb1; rs.last = -1; rs.base = 0u; rs.end = 0u; }
    SelTarget tg;
    tg.init(lv.total[c], lv.nleaf[c]);
    if (!preloaded) {
        const uint32_t *g = ss.hist + (size_t)c * nb1;
        for (int i = tid; i < nb1; i += kThreads) hbuf[i] = __ldcg(g + i);      // coalesced, one latency
    }
    __syncthreads();
    const uint32_t *h = hbuf + tid * per;
    uint32_t sum = 0;
    for (int j = 0; j < per; ++j) sum += h[j];
    uint32_t total;
    const uint32_t excl = sel_block_scan(sum, rs.w, total);
    int myFirst = nb1, myLast = -1;
    uint32_t p = excl;
    const bool sampled = sampleS > 1;
    if (sampled) {      // block-uniform
        if (tid < 32) {
            SelSampleEst se;
            se.init(total, lv.total[c], sampleZ);
            uint32_t pA, pB;
            sel_sample_crit(se, tg, total, pA, pB);
            if (tid == 0) { rs.pA = pA; rs.pB = pB; }
        }
        __syncthreads();
        const uint32_t pA = rs.pA, pB = rs.pB;
        for (int j = 0; j < per; ++j) {
            const uint32_t pn = p + h[j];
            const int b = tid * per + j;
            if (pn >= pA) myFirst = min(myFirst, b);
            if (p <= pB) myLast = max(myLast, b);
            p = pn;
        }
    } else {
        for (int j = 0; j < per; ++j) {
            const uint32_t pn = p + h[j];
            const int b = tid * per + j;
            if (tg.diff(pn) > -3) myFirst = min(myFirst, b);
            if (tg.diff(p) < 3) myLast = max(myLast, b);
            p = pn;
        }
    }
    if (myFirst < nb1) atomicMin(&rs.first, myFirst);
    if (myLast >= 0) atomicMax(&rs.last, myLast);
    __syncthreads();
    const int bf = rs.first, bl = rs.last;
    p = excl;
    for (int j = 0; j < per; ++j) {
        const int b = tid * per + j;
        if (b == bf) rs.base = p;
        p += h[j];
        if (b == bl) rs.end = p;
    }
    __syncthreads();
    // bf <= bl always (tests/test_select_model.py::ambiguous_range); guarded anyway.  More candidates than one block
    // can stage: the cell is left to the iterative search.
    // (sampled rows: base / ncand published below are sample counts - FINISH takes the exact ones from the visit records)
    const bool ok = bf <= bl && bf < nb1 && bl >= 0 && (sampled || (rs.end - rs.base) <= candCap);
    bfOut = ok ? (uint32_t)bf : 1u;
    blOut = ok ? (uint32_t)bl : 0u;
    if (publish && tid == 0) {
        ss.bfirst[c] = bfOut; ss.blast[c] = blOut;
        ss.base[c] = ok ? rs.base : 0u;
        ss.ncand[c] = ok ? rs.end - rs.base : 0u;
        ss.flag[c] = ok ? 0u : 1u;
        if (!ok && countFlag) atomicAdd(ss.n_flagged, 1u);
    }
    __syncthreads();
}

// =====================================================================================
// HIST / COMPACT: one streaming pass over the tiles of the level (same tile ownership, classification and cp.async
// ring as stream_count_pass).  HIST keeps `rep` copies of the block's histogram (copy = lane mod rep) to thin out
// shared-memory bank conflicts; COMPACT tests bin membership on the un-truncated bin coordinate.
// =====================================================================================
struct SelStreamSmem {
    uint32_t uTile[kMaxUnits], uCell[kMaxUnits];
    int uAx[kMaxUnits];
    float uLo[kMaxUnits], uScale[kMaxUnits];
    uint32_t fTile[kMaxUnits], fCell[kMaxUnits];
    uint32_t wS[kWarps], wF[kWarps];
    uint32_t wN[kWarps];                          // COMPACT: staged candidates per warp
    uint32_t gbase;
    uint32_t below, spilled;                      // COMPACT: PreLeft bookkeeping of the running cell
    uint32_t blkCur, visitStart, nVis;            // COMPACT with private regions: fill level of the block's region, start of the running visit
    SelVisit vis[kSelMaxVisit];
    float vLo, vHi;                               // COMPACT: value bounds of the running cell's candidate bins
    SelResolveSmem rs;
};

// bounds on t = max((x - lo) * scale, 0) equivalent to first <= sel_bin(x) <= last (bins are integers <= nb - 1)
__device__ __forceinline__ void sel_bin_bounds(uint32_t first, uint32_t last, int nb, float &fLo, float &fHi) {
    fLo = (float)first;
    fHi = (last + 1u >= (uint32_t)nb) ? __int_as_float(0x7f800000) : (float)(last + 1u);
    if (first > last) { fLo = 1.f; fHi = 0.f; }     // empty
}

// The same membership test on the VALUE instead of the bin coordinate: sel_bin is monotone, so
//     first <= sel_bin(v) <= last   <=>   vLo <= v < vHi
// for vLo = the smallest float whose bin is >= first and vHi = the smallest float whose bin is > last, found by bisection
// over the floats in their total order (32 evaluations of sel_bin, once per cell).  Two compares per particle instead
// of subtract, multiply, clamp and three compares.  vHi = NaN stands for "no upper limit" (the tests below are written
// so that NaN passes everything); an empty range gives vLo = +inf, vHi = -inf.  (NaN coordinates are not supported
// by the search: they would be binned low here and never counted left by the reference.)
__device__ __forceinline__ uint32_t sel_f2key(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float sel_key2f(uint32_t k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }
// smallest float v in [-inf, +inf] with sel_bin(v, lo, scale, nb) >= b (b >= 1; +inf if even +inf falls short)
__device__ __forceinline__ float sel_bin_threshold(int b, float lo, float scale, int nb) {
    const float inf = __int_as_float(0x7f800000);
    uint32_t a = sel_f2key(-inf), e = sel_f2key(inf);        // answer in [a, e]
    if (sel_bin(-inf, lo, scale, nb) >= b) return -inf;
    // invariant: bin(key a) < b; bin(key e) >= b or e = key(+inf): if even +inf falls short the loop ends at e.
    // The bin edge lo + b / scale is within a few ulps of the answer: bracket it first (18 evaluations instead of 32)
    if (scale > 0.f) {
        const float g = __fadd_rn(lo, __fdiv_rn((float)b, scale));
        if (fabsf(g) < inf) {
            const uint32_t kg = sel_f2key(g);
            if (kg - a > 256u) { const uint32_t a2 = kg - 256u; if (sel_bin(sel_key2f(a2), lo, scale, nb) < b) a = a2; }
            if (e - kg > 256u) { const uint32_t e2 = kg + 256u; if (sel_bin(sel_key2f(e2), lo, scale, nb) >= b) e = e2; }
        }
    }
    while (e - a > 1u) {
        const uint32_t m = a + ((e - a) >> 1);
        if (sel_bin(sel_key2f(m), lo, scale, nb) >= b) e = m; else a = m;
    }
    return sel_key2f(e);
}
// The same by a whole warp (all 32 lanes call it with the same arguments and get the same result): 32 keys of the
// bracket are tried at once, so the bisection takes 2 dependent evaluations near the bin edge (7 from the full range)
// instead of 18 (32) - it sits on the critical path of every cell that one block searches.
__device__ __forceinline__ float sel_bin_threshold_warp(int b, float lo, float scale, int nb) {
    const float inf = __int_as_float(0x7f800000);
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t a = sel_f2key(-inf), e = sel_f2key(inf);
    if (sel_bin(-inf, lo, scale, nb) >= b) return -inf;
    if (scale > 0.f) {
        const float g = __fadd_rn(lo, __fdiv_rn((float)b, scale));
        if (fabsf(g) < inf) {
            const uint32_t kg = sel_f2key(g);
            // lane 0 tries kg - 256 as the new lower end, lane 1 kg + 256 as the new upper end
            const bool canA = kg - a > 256u, canE = e - kg > 256u;
            const uint32_t kt = lane == 0u ? kg - 256u : kg + 256u;
            const bool ge = sel_bin(sel_key2f(kt), lo, scale, nb) >= b;
            const unsigned m = __ballot_sync(0xffffffffu, ge);
            if (canA && !(m & 1u)) a = kg - 256u;
            if (canE && (m & 2u)) e = kg + 256u;
        }
    }
    // invariant: bin(key a) < b; bin(key e) >= b, or e = key(+inf)
    while (e - a > 1u) {
        const uint32_t span = e - a;                                   // >= 2
        const uint32_t step = (span + 31u) / 32u;                      // lanes try a + step, a + 2 step, ... (clamped to e)
        const uint64_t kk = (uint64_t)a + (uint64_t)step * (lane + 1u);
        const uint32_t kt = kk >= (uint64_t)e ? e : (uint32_t)kk;
        const bool ge = sel_bin(sel_key2f(kt), lo, scale, nb) >= b;
        const unsigned m = __ballot_sync(0xffffffffu, ge);
        if (!m) return sel_key2f(e);                                   // even e falls short (e = key(+inf)): +inf
        const int j = __ffs(m) - 1;                                    // first lane at or above the threshold
        const uint32_t ej = __shfl_sync(0xffffffffu, kt, j);
        const uint32_t aj = j ? __shfl_sync(0xffffffffu, kt, j - 1) : a;
        if (ej - aj >= span) return sel_key2f(e);                      // (cannot happen: the bracket always shrinks)
        e = ej; a = aj;
    }
    return sel_key2f(e);
}
// value bounds by one warp (block-uniform results must be broadcast by the caller)
__device__ __forceinline__ void sel_value_bounds_warp(uint32_t first, uint32_t last, float lo, float scale, int nb, float &vLo, float &vHi) {
    const float inf = __int_as_float(0x7f800000);
    if (first > last) { vLo = inf; vHi = -inf; return; }
    vLo = first == 0u ? -inf : sel_bin_threshold_warp((int)first, lo, scale, nb);
    vHi = (last + 1u >= (uint32_t)nb) ? __int_as_float(0x7fc00000) : sel_bin_threshold_warp((int)last + 1, lo, scale, nb);
}
__device__ __forceinline__ void sel_value_bounds(uint32_t first, uint32_t last, float lo, float scale, int nb, float &vLo, float &vHi) {
    const float inf = __int_as_float(0x7f800000);
    if (first > last) { vLo = inf; vHi = -inf; return; }
    vLo = first == 0u ? -inf : sel_bin_threshold((int)first, lo, scale, nb);
    vHi = (last + 1u >= (uint32_t)nb) ? __int_as_float(0x7fc00000) : sel_bin_threshold((int)last + 1, lo, scale, nb);
}
// membership tests (vHi may be NaN = no upper limit)
__device__ __forceinline__ bool sel_is_low(float v, float vLo) { return v < vLo; }
__device__ __forceinline__ bool sel_is_cand(float v, float vLo, float vHi) { return v >= vLo && !(v >= vHi); }

// Exchange arenas of all ranks (protocol of orb_exchange.cuh) and the cross-rank barrier at the start of a kernel:
// block 0 stores the exchange's sequence number into every peer's flag word (this rank's preceding kernel has
// completed: call after pdl_enter), every block spins on its own rank's flag words.  n <= 1: nothing to wait for.
struct XArena {
    int n, self;
    uint32_t seq;                      // sequence number of this kernel's barrier (monotone, identical on all ranks)
    uint32_t *arena[kMaxPeers];        // flags[64] | ... (word offsets below, identical on all ranks)
    // result records [nCells][8] | candidate counts [R][ownedStride] | candidate slots [R][ownedStride][slotWords] |
    // local rows | global rows
    uint32_t offRes, offRecvCnt, offRecv, offHistL, offHistG;
};
__device__ __forceinline__ uint32_t ld_sys_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void x_barrier(const XArena &xa) {
    const int t = threadIdx.x;
    if (xa.n <= 1) return;
    if (blockIdx.x == 0 && t < xa.n && t != xa.self) {
        __threadfence_system();
        asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(xa.arena[t] + xa.self), "r"(xa.seq) : "memory");
    }
    if (t < xa.n && t != xa.self) {
        const uint32_t *f = xa.arena[xa.self] + t;
        while ((int32_t)(ld_sys_u32(f) - xa.seq) < 0) {}
        __threadfence_system();
    }
    __syncthreads();
}

// (PRIV: COMPACT with private candidate regions + cells resolved by k_sel_resolve - a separate instantiation, so that the
//  plain pass does not carry its code: these kernels run for 20-30 us on 2^24 particles and an instruction-cache
//  miss at every phase change shows)
template <int MODE, bool PRIV = false>
__global__ void __launch_bounds__(kThreads, MODE == kSelHist ? 4 : 3)
k_sel_stream(const float *__restrict__ x, const float *__restrict__ y, const float *__restrict__ z, float *__restrict__ cand,
             LevelState lv, SelState ss, const uint32_t *__restrict__ tile_first, uint32_t nCells, uint32_t nLocal,
             uint32_t nTiles, int nb1, int rep, uint32_t candCap, unsigned long long *dbg, float *__restrict__ slots,
             uint32_t slotWords, int preResolved, XArena xa /* n > 1: cross-rank barrier before the pass (the rows it
             resolves from are the all-reduced ones) */,
             PreLeft *__restrict__ pre /* COMPACT: records for the partition's phase 1 (may be null) */, uint32_t preTag,
             uint32_t tilesPerBlockIn /* != 0: chunk per block in count tiles, shared with the partition */,
             SelPriv sp /* single rank: sampled rows, private candidate regions (zero: neither) */) {
    extern __shared__ __align__(16) unsigned char sel_smem[];
    pdl_enter();
    x_barrier(xa);
    unsigned long long *bs = dbg ? dbg + (size_t)blockIdx.x * 4 : nullptr;      // ORB_DEBUG_TIMES=2: per-block phase stamps
    if (bs && threadIdx.x == 0) bs[0] = gtimer();
    // dynamic: ring (kCountStages x 16 KB) | HIST: hist[nb1][rep]  /  COMPACT: stage[kWarps][kSelWarpStage] | hbuf[nb1]
    float4 *ring = reinterpret_cast<float4 *>(sel_smem);
    uint32_t *s_hist = reinterpret_cast<uint32_t *>(sel_smem + (size_t)kCountStages * kCountTile * sizeof(float));
    float *s_stage = reinterpret_cast<float *>(s_hist);
    uint32_t *s_hbuf = s_hist + kWarps * kSelWarpStage;
    __shared__ SelStreamSmem sm;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int repSel = lane & (rep - 1);

    if (MODE == kSelHist) for (int i = tid; i < nb1 * rep; i += kThreads) s_hist[i] = 0u;
    if (MODE == kSelCompact && tid < kWarps) sm.wN[tid] = 0u;
    if (tid == 0) { sm.below = 0u; sm.spilled = 0u; sm.blkCur = 0u; sm.visitStart = 0u; sm.nVis = 0u; }
    __syncthreads();

    // block b owns the contiguous tiles [tb0, tb1) - of the tiles the pass visits: a sampled HIST pass (sampled rows)
    // visits the tiles 0, S, 2S, ... only, and counts in those
    const uint32_t tileStep = (MODE == kSelHist && sp.sampleS > 1) ? (uint32_t)sp.sampleS : 1u;
    const uint32_t nVisit = (nTiles + tileStep - 1u) / tileStep;
    const uint32_t tilesPerBlock = tilesPerBlockIn ? tilesPerBlockIn : (nVisit + gridDim.x - 1) / gridDim.x;
    const uint32_t tb0 = min(blockIdx.x * tilesPerBlock, nVisit), tb1 = min(tb0 + tilesPerBlock, nVisit);
    constexpr bool priv = MODE == kSelCompact && PRIV;
    const uint32_t chunkStart = tb0 * (uint32_t)kCountTile;
    int cur = -1;
    float lo = 0.f, scale = 0.f;
    float vLo = __int_as_float(0x7f800000), vHi = __int_as_float(0xff800000);      // empty
    const float nbm1 = (float)(nb1 - 1);
    unsigned below = 0u;          // COMPACT: this thread's particles of the running cell below the candidate bins
    bool curOk = false;           // ... the running cell has candidate bins (not flagged by the resolve)

    // ---- flush of the per-cell block state (block-uniform) ----
    auto flush = [&]() {
        if (cur < 0) return;
        __syncthreads();
        if (MODE == kSelHist) {
            for (int b = tid; b < nb1; b += kThreads) {
                uint32_t v = 0;
                for (int r = 0; r < rep; ++r) { v += s_hist[b * rep + r]; s_hist[b * rep + r] = 0u; }
                if (v) atomicAdd(&ss.hist[(size_t)cur * nb1 + b], v);
            }
        } else {
            uint32_t tot = 0, off = 0;
#pragma unroll
            for (int w = 0; w < kWarps; ++w) { const uint32_t n = sm.wN[w]; if (w < warp) off += n; tot += n; }
            if (pre || priv) {      // block-uniform
                const unsigned wb = __reduce_add_sync(0xffffffffu, below);
                if (lane == 0 && wb) atomicAdd(&sm.below, wb);
                below = 0u;
            }
            if (priv) {
                // the visit's candidates are [visitStart, blkCur) of the block's region, whatever the warps spilled before
                if (tid == 0) { sm.gbase = sm.blkCur; sm.blkCur += tot; }
                __syncthreads();
                if (tid == 0) {
                    const uint32_t v0 = sm.visitStart, cnt = sm.blkCur - v0;
                    if (sm.nVis < (uint32_t)kSelMaxVisit) {
                        SelVisit V;
                        V.cell = (uint32_t)cur; V.off = chunkStart + v0; V.cnt = cnt; V.below = sm.below;
                        sm.vis[sm.nVis] = V;
                    }
                    sm.nVis++;          // beyond kSelMaxVisit: FINISH sees the overflow and flags the cells this block touched
                    if (pre) {
                        PreLeft P;
                        P.tag = curOk ? preTag : ~preTag;
                        P.cell = (uint32_t)cur; P.below = sm.below; P.gbase = chunkStart + v0; P.tot = cnt; P.kind = 2u; P.pad_[0] = P.pad_[1] = 0u;
                        pre[blockIdx.x] = P;
                    }
                    sm.visitStart = sm.blkCur;
                    sm.below = 0u;
                }
                if (tot) {
                    float *dst = cand + chunkStart + sm.gbase + off;
                    const uint32_t n = sm.wN[warp];
                    for (uint32_t i = lane; i < n; i += 32u) dst[i] = s_stage[warp * kSelWarpStage + i];
                }
                __syncthreads();
                if (tid < kWarps) sm.wN[tid] = 0u;
                __syncthreads();
                return;
            }
            if (tid == 0) sm.gbase = tot ? atomicAdd(&ss.cursor[cur], tot) : 0u;
            __syncthreads();
            if (pre && tid == 0) {
                // The last record a block writes describes its trailing segment (cells are entered in order).  A warp
                // that spilled on its own has scattered the segment's candidates: no record for it.
                const uint32_t lim2 = slotWords ? slotWords - 1u : 0xffffffffu;
                PreLeft P;
                P.tag = (curOk && !sm.spilled && sm.gbase + tot <= lim2) ? preTag : ~preTag;
                P.cell = (uint32_t)cur; P.below = sm.below; P.gbase = sm.gbase; P.tot = tot; P.kind = 0u; P.pad_[0] = P.pad_[1] = 0u;
                pre[blockIdx.x] = P;
                sm.below = 0u; sm.spilled = 0u;
            }
            if (tot) {
                // one rank: the cell's list in the idle column.  Several ranks: the cell's fixed-size slot of the
                // all-gather buffer (values beyond the slot are dropped; the count word tells every rank)
                float *dst = slotWords ? slots + (size_t)cur * slotWords : cand + lv.bnd[cur];
                const uint32_t lim = slotWords ? slotWords - 1u : 0xffffffffu;
                const uint32_t o0 = sm.gbase + off;
                const uint32_t n = sm.wN[warp];
                for (uint32_t i = lane; i < n; i += 32u)
                    if (o0 + i < lim) dst[o0 + i] = s_stage[warp * kSelWarpStage + i];
            }
            __syncthreads();
            if (tid < kWarps) sm.wN[tid] = 0u;
        }
        __syncthreads();
    };
    // a warp's staging region is nearly full (massive ties): write it out on its own
    auto warp_spill = [&]() {
        const uint32_t n = sm.wN[warp];
        if (n + 512u <= (uint32_t)kSelWarpStage) return;
        uint32_t g = 0;
        if (lane == 0) {
            if (priv) g = atomicAdd(&sm.blkCur, n);
            else { g = atomicAdd(&ss.cursor[cur], n); sm.spilled = 1u; }
        }
        g = __shfl_sync(0xffffffffu, g, 0);
        float *dst = priv ? cand + chunkStart : (slotWords ? slots + (size_t)cur * slotWords : cand + lv.bnd[cur]);
        const uint32_t lim = slotWords ? slotWords - 1u : 0xffffffffu;
        for (uint32_t i = lane; i < n; i += 32u)
            if (g + i < lim) dst[g + i] = s_stage[warp * kSelWarpStage + i];
        __syncwarp();
        if (lane == 0) sm.wN[warp] = 0u;
        __syncwarp();
    };
    auto enter_cell = [&](uint32_t c, float cLo, float cScale) {   // block-uniform
        if ((int)c == cur) return;
        flush();
        cur = (int)c; lo = cLo; scale = cScale;
        if (MODE == kSelCompact) {
            // resolve the cell's candidate bins from its (complete) histogram; the block that holds the cell's first
            // particle publishes the result for FINISH
            const uint32_t cb = lv.bnd[c];
            // (several ranks: k_selmr_prep resolves and publishes every cell, also those without local particles)
            const bool publish = slotWords == 0u && cb >= tb0 * (uint32_t)kCountTile && cb < tb1 * (uint32_t)kCountTile;
            uint32_t bf, bl;
            if (PRIV || preResolved) { bf = __ldcg(&ss.bfirst[c]); bl = __ldcg(&ss.blast[c]); }     // k_selx_resolve / k_sel_resolve did it for the level
            else sel_resolve_cell(lv, ss, c, nb1, candCap, publish, s_hbuf, sm.rs, bf, bl);
            if (PRIV) {
                vLo = __ldcg(&ss.vlo[c]); vHi = __ldcg(&ss.vhi[c]);
            }
            else {
                __syncthreads();
                if (warp == 0) {
                    float a, b;
                    sel_value_bounds_warp(bf, bl, cLo, cScale, nb1, a, b);
                    if (lane == 0) { sm.vLo = a; sm.vHi = b; }
                }
                __syncthreads();
                vLo = sm.vLo; vHi = sm.vHi;
            }
            curOk = bf <= bl;
        }
    };
    // bin coordinate: max((x - lo) * scale, 0); NaN (inf * 0) -> 0
    auto coord = [&](float v) { return fmaxf(__fmul_rn(__fsub_rn(v, lo), scale), 0.f); };


    for (uint32_t base = tb0; base < tb1; base += (uint32_t)kMaxUnits) {
        // ---- phase 1: classify up to kMaxUnits tiles ----
        const uint32_t tv = base + (uint32_t)tid;
        const uint32_t t = tv * tileStep;
        int kind = 0;   // 0 skip, 1 stream, 2 fragmented
        uint32_t c = 0;
        if (tid < kMaxUnits && tv < tb1) {
            const uint32_t t0 = t * (uint32_t)kCountTile, t1 = min(t0 + (uint32_t)kCountTile, nLocal);
            c = tile_first[t * (kCountTile / kMapTile)];
            const uint32_t cb = lv.bnd[c], ce = lv.bnd[c + 1];
            if (cb <= t0 && ce >= t1 && (t1 - t0) == (uint32_t)kCountTile) {
                kind = lv.active[c] ? 1 : 0;
            } else kind = 2;
        }
        const unsigned mS = __ballot_sync(0xffffffffu, kind == 1), mF = __ballot_sync(0xffffffffu, kind == 2);
        __syncthreads();
        if (lane == 0) { sm.wS[warp] = __popc(mS); sm.wF[warp] = __popc(mF); }
        __syncthreads();
        uint32_t offS = 0, offF = 0, nS = 0, nF = 0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
            if (w < warp) { offS += sm.wS[w]; offF += sm.wF[w]; }
            nS += sm.wS[w]; nF += sm.wF[w];
        }
        const unsigned ltMask = (1u << lane) - 1u;
        if (kind == 1) {
            const uint32_t r = offS + __popc(mS & ltMask);
            sm.uTile[r] = t; sm.uCell[r] = c; sm.uAx[r] = lv.axis[c];
            const float L = lv.mL[c], R = lv.mR[c];
            sm.uLo[r] = L; sm.uScale[r] = sel_scale(L, R, nb1);
        }
        if (kind == 2) { const uint32_t r = offF + __popc(mF & ltMask); sm.fTile[r] = t; sm.fCell[r] = c; }
        __syncthreads();
        if (bs && tid == 0 && base == tb0) bs[1] = gtimer();

        // tiles holding cell boundaries (or the array tail): per cell segment, plain loads.  Processed in tile order
        // between the streamed tiles, so the block changes cell (and flushes) once per cell it touches.
        auto process_frag = [&](uint32_t k) {
            const uint32_t tt = sm.fTile[k];
            const uint32_t t0 = tt * (uint32_t)kCountTile, t1 = min(t0 + (uint32_t)kCountTile, nLocal);
            for (uint32_t cc = sm.fCell[k]; cc < nCells; ++cc) {
                const uint32_t cb = lv.bnd[cc], ce = lv.bnd[cc + 1];
                if (cb >= t1) break;
                const uint32_t s0 = max(cb, t0), s1 = min(ce, t1);
                const bool use = s1 > s0 && lv.active[cc] != 0u;
                if (use) {
                    const float L = lv.mL[cc], R = lv.mR[cc];
                    enter_cell(cc, L, sel_scale(L, R, nb1));
                    // a segment is at most one tile: every thread loads its (up to) 16 elements before using any
                    const float *col = pick_col(lv.axis[cc], x, y, z);
                    float v[16];
                    unsigned inMask = 0u;
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const uint32_t i = s0 + (uint32_t)tid + (uint32_t)j * kThreads;
                        const bool in = i < s1;
                        v[j] = in ? __ldg(col + i) : 0.f;
                        inMask |= (unsigned)in << j;
                    }
                    if (MODE == kSelHist) {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (inMask & (1u << j)) atomicAdd(&s_hist[__float2int_rz(fminf(coord(v[j]), nbm1)) * rep + repSel], 1u);
                    } else {
                        unsigned keep = 0u, lows = 0u;
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            keep |= (unsigned)sel_is_cand(v[j], vLo, vHi) << j;
                            lows |= (unsigned)sel_is_low(v[j], vLo) << j;
                        }
                        keep &= inMask;
                        below += __popc(lows & inMask);
                        warp_spill();
                        sel_warp_append<16>(v, keep, s_stage + warp * kSelWarpStage, &sm.wN[warp]);

                    }
                }
                if (ce >= t1) break;
            }
        };
        uint32_t fi = 0;
        // ---- phase 2a: whole tiles through the per-thread cp.async ring ----
        if (nS) {
            auto issue = [&](uint32_t k) {
                const float4 *p = reinterpret_cast<const float4 *>(pick_col(sm.uAx[k], x, y, z) + sm.uTile[k] * (uint32_t)kCountTile) + tid;
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
                cp_async_wait<kCountStages - 1>();
                for (; fi < nF && sm.fTile[fi] < sm.uTile[k]; ++fi) process_frag(fi);
                enter_cell(sm.uCell[k], sm.uLo[k], sm.uScale[k]);
                const float4 *src = ring + (k % kCountStages) * (4 * kThreads) + tid;
                const float4 q0 = src[0], q1 = src[kThreads], q2 = src[2 * kThreads], q3 = src[3 * kThreads];
                const float v[16] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, q3.x, q3.y, q3.z, q3.w};
                if (MODE == kSelHist) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int b = __float2int_rz(fminf(coord(v[j]), nbm1));      // == sel_bin(v[j], lo, scale, nb1)
                        atomicAdd(&s_hist[b * rep + repSel], 1u);
                    }
                } else {
                    // two compares and two predicated adds per particle: #{v >= vLo} and #{v >= vHi}; the thread holds a
                    // candidate iff the two differ (vHi = NaN, "no upper limit", compares false)
                    unsigned nGeLo = 0u, nGeHi = 0u;
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        nGeLo += (v[j] >= vLo) ? 1u : 0u;
                        nGeHi += (v[j] >= vHi) ? 1u : 0u;
                    }
                    below += 16u - nGeLo;
                    const bool has = nGeLo != nGeHi;
                    // Candidates are sparse (a thread rarely has one): only warps that hold any take this path, only the
                    // lanes that hold one work in it - a shared-memory atomic on the warp's fill level, the values
                    // re-read from the tile in shared memory (no register array indexed by a run-time bit position).
                    if (__any_sync(0xffffffffu, has)) {
                        warp_spill();
                        unsigned keep = 0u;
                        if (has) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) keep |= (unsigned)sel_is_cand(v[j], vLo, vHi) << j;
                        }
                        if (keep) {
                            uint32_t pos = atomicAdd(&sm.wN[warp], (uint32_t)__popc(keep));
                            float *dstw = s_stage + warp * kSelWarpStage;
                            while (keep) {
                                const int j = __ffs(keep) - 1;
                                keep &= keep - 1u;
                                dstw[pos++] = reinterpret_cast<const float *>(src + (j >> 2) * kThreads)[j & 3];
                            }
                        }
                        __syncwarp();
                    }
                }
            }
            cp_async_wait<0>();
        }
        for (; fi < nF; ++fi) process_frag(fi);
        __syncthreads();
    }
    if (bs && tid == 0) bs[2] = gtimer();
    flush();
    if (priv) {
        __syncthreads();
        if (tid == 0) sp.visits[blockIdx.x].n = sm.nVis;
        if (tid < kSelMaxVisit && (uint32_t)tid < sm.nVis) sp.visits[blockIdx.x].v[tid] = sm.vis[tid];
    }
    if (bs && tid == 0) bs[3] = gtimer();
}

// State of the parallel FINISH (sampled streaming levels).  One block finishing a cell walks all its candidates twice
// (refinement histogram, then the few ambiguous values): with sampled rows that is 60 000 - 160 000 values on ONE SM,
// 60 - 135 us.  Instead, four small kernels (binning inside COMPACT's append path was measured too: +10...35 us per pass):
//   k_sel_fine    every COMPACT block's region again (L2), by as many blocks: the candidates binned into their cell's
//                 fine histogram (shared-memory histogram per visit, non-empty bins added to the global row);
//   k_sel_fin_a   one block per cell: exact `below` / candidate count from the visit records, proof of the bracket,
//                 scan of the fine histogram -> ambiguous fine bins [f2, l2], their value bounds, count below them;
//   k_sel_gather  every COMPACT block's region again (L2), by as many blocks: values inside the ambiguous fine bins go
//                 to the cell's short list (a global cursor; a few hundred values per cell);
//   k_sel_fin_b   one block per cell: the block search on that list, with the fine bins as its outer histogram.
struct SelPar {
    uint32_t *fine;            // [maxCells][kSelBins2] fine histograms (cleared by k_sel_resolve)
    float *lo2, *sc2;          // [maxCells] fine bin function
    int32_t *f2, *l2;          // [maxCells] ambiguous fine bins
    uint32_t *base2, *k2;      // [maxCells] particles below fine bin f2; values in [f2, l2]
    float *v2lo, *v2hi;        // [maxCells] value bounds of [f2, l2]
    uint32_t *ambCnt;          // [maxCells] fill level of the cell's list
    float *amb;                // [maxCells][kSelAmbCap]
};

// RESOLVE of every cell of a streaming level as a kernel of its own (single rank, between HIST and COMPACT): one block
// per cell publishes the candidate bins and their value bounds.  COMPACT then enters a cell with four loads instead of
// a 32 KB row, a scan and a bisection per block and cell, and needs no row buffer in shared memory.
__global__ void __launch_bounds__(kThreads) k_sel_resolve(LevelState lv, SelState ss, uint32_t nCells, int nb1, uint32_t candCap,
                                                           int sampleS, float sampleZ, SelPar pr /* fine == null: no parallel FINISH */) {
    extern __shared__ __align__(16) unsigned char sel_smem[];
    uint32_t *hbuf = reinterpret_cast<uint32_t *>(sel_smem);
    __shared__ SelResolveSmem rs;
    pdl_enter();
    for (uint32_t c = blockIdx.x; c < nCells; c += gridDim.x) {
        if (!lv.active[c]) continue;          // block-uniform
        uint32_t bf, bl;
        sel_resolve_cell(lv, ss, c, nb1, candCap, true, hbuf, rs, bf, bl, false, true, sampleS, sampleZ);
        if (threadIdx.x < 32) {
            const float L = lv.mL[c];
            float vLo, vHi;
            sel_value_bounds_warp(bf, bl, L, sel_scale(L, lv.mR[c], nb1), nb1, vLo, vHi);
            if (threadIdx.x == 0) { ss.vlo[c] = vLo; ss.vhi[c] = vHi; }
        }
        if (pr.fine) {      // fine bin function over the candidate bins' interval (the range k_sel_finish refines over), cleared row
            for (int i = threadIdx.x; i < kSelBins2; i += kThreads) pr.fine[(size_t)c * kSelBins2 + i] = 0u;
            if (threadIdx.x == 0) {
                const float L = lv.mL[c], s1 = sel_scale(L, lv.mR[c], nb1);
                float a = 0.f, b = 0.f;
                if (s1 > 0.f && bf <= bl) {
                    a = __fadd_rn(L, __fdiv_rn((float)bf, s1));
                    b = __fadd_rn(L, __fdiv_rn((float)(bl + 1u), s1));
                }
                pr.lo2[c] = a; pr.sc2[c] = sel_scale(a, b, kSelBins2);
                pr.ambCnt[c] = 0u;
            }
        }
    }
}

// =====================================================================================
// Block-level search on values staged in shared memory (step 4 above).  Block-uniform call, any block size that is a
// multiple of 32 and divides 2048.
// vals[K]: the candidates; `base`: particles of the cell known to be smaller than every candidate; `outer*`: the
// HIST pass' bin function and candidate bins (hasOuter = 0 when vals is the whole cell).
// Writes the cell's result (margins, iter, found, nleft) and returns true, or flags it and returns false (block-
// uniform); nothing else is written for a flagged cell.
// =====================================================================================
// f(v) for every staged value; `glob`: vals lies in global memory (a cell with more candidates than shared memory holds)
// - eight loads in flight per thread instead of one dependent load per iteration
template <typename F>
__device__ __forceinline__ void sel_vals_for_each(const float *vals, uint32_t K, bool glob, F f) {
    const uint32_t nT = blockDim.x;
    uint32_t i = threadIdx.x;
    if (glob) {
        for (; i + 7u * nT < K; i += 8u * nT) {
            float q[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) q[u] = __ldcg(vals + i + (uint32_t)u * nT);
#pragma unroll
            for (int u = 0; u < 8; ++u) f(q[u]);
        }
        for (; i < K; i += nT) f(__ldcg(vals + i));
    } else {
        for (; i < K; i += nT) f(vals[i]);
    }
}

// where the block search reads its values from: an array (shared or global memory), or the pieces the COMPACT blocks
// left in their private regions (piece i: cand[pOff[i] .. pOff[i] + pCnt[i]), pPos[i] = values in the pieces before it)
struct SelSrcArray {
    const float *vals;
    uint32_t K;
    bool glob;
    template <typename F>
    __device__ __forceinline__ void for_each(F f) const { sel_vals_for_each(vals, K, glob, f); }
    __device__ __forceinline__ const float *shared_ptr() const { return glob ? nullptr : vals; }      // the values, if they lie in shared memory
};
struct SelSrcPieces {
    const float *cand;
    const uint32_t *pOff, *pCnt, *pPos;      // shared memory
    uint32_t K;
    __device__ __forceinline__ const float *shared_ptr() const { return nullptr; }
    // thread t takes the values t, t + nT, ... of the concatenation; its piece index only ever advances.  kFly loads
    // in flight per thread (the pieces lie in L2: the pass is a few latencies, not bandwidth).
    static constexpr int kFly = 16;
    template <typename F>
    __device__ __forceinline__ void for_each(F f) const {
        const uint32_t nT = blockDim.x;
        uint32_t p = 0;
        for (uint32_t i = threadIdx.x; i < K; i += (uint32_t)kFly * nT) {
            float q[kFly];
            bool in[kFly];
#pragma unroll
            for (int u = 0; u < kFly; ++u) {
                const uint32_t ii = i + (uint32_t)u * nT;
                in[u] = ii < K;
                q[u] = 0.f;
                if (in[u]) {
                    while (ii >= pPos[p] + pCnt[p]) ++p;       // (ii < K: ends at the last non-empty piece at the latest)
                    q[u] = __ldcg(cand + pOff[p] + (ii - pPos[p]));
                }
            }
#pragma unroll
            for (int u = 0; u < kFly; ++u)
                if (in[u]) f(q[u]);
        }
    }
};

struct SelSearchSmem {
    float wmin[32], wmax[32];
    uint32_t w[32];
    int first, last;
    uint32_t base2, end2, namb, finalCnt;
    float lo2, scale2, cutf;
    int needFinal;
    float resL, resR;             // result of sel_block_search_core
    int resIt, resFnd;
    uint32_t resNleft;
};

// Core of the search: computes the cell's result into sm.res* (valid for every thread on return) and returns true, or
// returns false when the cell has to be left to the iterative path (block-uniform).  Writes no global memory.
template <typename SRC>
__device__ __forceinline__ bool sel_block_search_core_src(const SRC &src, uint32_t K, uint32_t base, int hasOuter, float lo1,
                                                          float scale1, int nb1, int bfirst, int blast, uint32_t *hist2, float *amb,
                                                          const LevelState &lv, uint32_t c, SelSearchSmem &sm,
                                                          unsigned long long *dbg = nullptr) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nThreads = (int)blockDim.x, nWarps = nThreads >> 5;
    const float inf = __int_as_float(0x7f800000);
    SelTarget tg;
    tg.init(lv.total[c], lv.nleaf[c]);
    const float L0 = lv.mL[c], R0 = lv.mR[c];
    // Few values in shared memory (the candidates of a small cell searched with exact rows): no refinement - every value
    // takes part in the replay's counts.  Two barriers instead of a dozen, no second histogram.
    constexpr uint32_t kTiny = 256;
    const int nb2 = K > 4096u ? kSelBins2 : max(256, nThreads);
    const int per = nb2 / nThreads;          // 1, 2 or 8
    const float *ambp = amb;
    float lo2 = 0.f, scale2 = 0.f;
    int first = 0, last = -1;
    uint32_t base2 = base, K2 = K;
    const bool tiny = K <= kTiny && src.shared_ptr() != nullptr;      // block-uniform
    if (tiny) {
        ambp = src.shared_ptr();
        __syncthreads();                    // (sm.* may still be read from the previous cell)
        if (tid == 0) { sm.finalCnt = 0u; sm.needFinal = 0; sm.cutf = 0.f; }
    } else {

    // ---- range of the staged values: the candidate bins' interval when there is an outer histogram, else min / max
    //      (any range gives a valid monotone bin function; a tight one just resolves better) ----
    const bool fromBins = hasOuter && scale1 > 0.f;
    float mn = inf, mx = -inf;
    if (!fromBins) {
        src.for_each([&](float v) { mn = fminf(mn, v); mx = fmaxf(mx, v); });
#pragma unroll
        for (int o = 16; o; o >>= 1) { mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o)); mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
    }
    __syncthreads();
    if (lane == 0) { sm.wmin[warp] = mn; sm.wmax[warp] = mx; }
    for (int i = tid; i < nb2; i += nThreads) hist2[i] = 0u;
    __syncthreads();
    if (tid == 0) {
        float a = inf, b = -inf;
        if (fromBins) {
            a = __fadd_rn(lo1, __fdiv_rn((float)bfirst, scale1));
            b = __fadd_rn(lo1, __fdiv_rn((float)(blast + 1), scale1));
        } else {
            for (int w = 0; w < nWarps; ++w) { a = fminf(a, sm.wmin[w]); b = fmaxf(b, sm.wmax[w]); }
        }
        if (K == 0u) { a = 0.f; b = 0.f; }
        sm.lo2 = a; sm.scale2 = sel_scale(a, b, nb2);
        sm.first = nb2; sm.last = -1; sm.base2 = 0u; sm.end2 = 0u; sm.namb = 0u; sm.finalCnt = 0u; sm.needFinal = 0; sm.cutf = 0.f;
    }
    __syncthreads();
    lo2 = sm.lo2; scale2 = sm.scale2;
    if (dbg && tid == 0) dbg[2] = gtimer();
    src.for_each([&](float v) { atomicAdd(&hist2[sel_bin(v, lo2, scale2, nb2)], 1u); });
    __syncthreads();
    if (dbg && tid == 0) dbg[3] = gtimer();

    // ---- prefix sums, ambiguous bins [first, last] ----
    uint32_t h[8];
    uint32_t sum = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) { h[j] = (j < per) ? hist2[tid * per + j] : 0u; sum += h[j]; }
    uint32_t total;
    const uint32_t excl = sel_block_scan(sum, sm.w, total);
    int myFirst = nb2, myLast = -1;
    {
        uint32_t p = base + excl;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (j < per) {
                const uint32_t pn = p + h[j];
                const int b = tid * per + j;
                if (tg.diff(pn) > -3) myFirst = min(myFirst, b);
                if (tg.diff(p) < 3) myLast = max(myLast, b);
                p = pn;
            }
        }
    }
    if (myFirst < nb2) atomicMin(&sm.first, myFirst);
    if (myLast >= 0) atomicMax(&sm.last, myLast);
    __syncthreads();
    first = sm.first; last = sm.last;
    {
        uint32_t p = base + excl;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (j < per) {
                const int b = tid * per + j;
                if (b == first) sm.base2 = p;
                p += h[j];
                if (b == last) sm.end2 = p;
            }
        }
    }
    __syncthreads();
    if (dbg && tid == 0) dbg[4] = gtimer();
    base2 = sm.base2;
    K2 = sm.end2 - base2;
    const bool tooMany = !(first <= last) || K2 > (uint32_t)kSelAmbCap;
    if (tooMany) return false;   // massive ties: leave the cell to the iterative path
    src.for_each([&](float v) {
        const int b = sel_bin(v, lo2, scale2, nb2);
        if (b >= first && b <= last) amb[atomicAdd(&sm.namb, 1u)] = v;
    });
    __syncthreads();

    }
    if (dbg && tid == 0) dbg[5] = gtimer();
    // ---- replay of orbit.cpp:149-232 by warp 0 (every lane computes the same scalars) ----
    float L = L0, R = R0;
    int it = 0;
    bool fnd = false;
    uint32_t nleft = 0;
    if (warp == 0) {
        while (it < kMaxIter) {
            const float cut = mid_cut(L, R);
            int dec = 0;
            if (hasOuter) {
                const int b1 = sel_bin(cut, lo1, scale1, nb1);
                dec = b1 < bfirst ? -1 : (b1 > blast ? 1 : 0);
            }
            if (dec == 0 && !tiny) {
                const int b2 = sel_bin(cut, lo2, scale2, nb2);
                dec = b2 < first ? -1 : (b2 > last ? 1 : 0);
            }
            ++it;
            if (dec == 0) {
                uint32_t n = 0;
                for (uint32_t i = lane; i < K2; i += 32u) n += (ambp[i] < cut) ? 1u : 0u;
                n = __reduce_add_sync(0xffffffffu, n);
                const uint32_t cnt = base2 + n;
                const int d = tg.diff(cnt);
                if (abs(d) < 3) { fnd = true; nleft = cnt; break; }       // orbit.cpp:208
                dec = d > 0 ? 1 : -1;
            }
            if (dec > 0) R = cut; else L = cut;                            // orbit.cpp:219,227
        }
        if (!fnd && lane == 0) {
            // capped cell: the count at getCut() of the last margins is needed for the partition
            const float cutf = mid_cut(L, R);
            int ok = 1;
            if (hasOuter) { const int b1 = sel_bin(cutf, lo1, scale1, nb1); ok = (b1 >= bfirst && b1 <= blast) ? 1 : 0; }
            sm.cutf = cutf;
            sm.needFinal = ok ? 1 : 2;
        }
    }
    __syncthreads();
    if (dbg && tid == 0) dbg[6] = gtimer();
    const int needFinal = sm.needFinal;
    if (needFinal == 2) return false;   // final cut outside the candidate bins: its exact count is not known here
    if (needFinal == 1) {
        const float cutf = sm.cutf;
        uint32_t n = 0;
        src.for_each([&](float v) { n += (v < cutf) ? 1u : 0u; });
        n = __reduce_add_sync(0xffffffffu, n);
        if (lane == 0 && n) atomicAdd(&sm.finalCnt, n);
        __syncthreads();
        nleft = base + sm.finalCnt;
    }
    if (tid == 0) { sm.resL = L; sm.resR = R; sm.resIt = it; sm.resFnd = fnd ? 1 : 0; sm.resNleft = nleft; }
    __syncthreads();
    return true;
}

__device__ __forceinline__ bool sel_block_search_core(const float *vals, uint32_t K, uint32_t base, int hasOuter, float lo1,
                                                      float scale1, int nb1, int bfirst, int blast, uint32_t *hist2, float *amb,
                                                      const LevelState &lv, uint32_t c, SelSearchSmem &sm,
                                                      unsigned long long *dbg = nullptr, bool glob = false /* vals in global memory */) {
    SelSrcArray src;
    src.vals = vals; src.K = K; src.glob = glob;
    return sel_block_search_core_src(src, K, base, hasOuter, lo1, scale1, nb1, bfirst, blast, hist2, amb, lv, c, sm, dbg);
}

// Search + commit: writes the cell's result (margins, iter, found, nleft) and the level statistics and returns true, or
// flags the cell and returns false (block-uniform); nothing else is written for a flagged cell.
template <typename SRC>
__device__ __forceinline__ bool sel_block_search_src(const SRC &src, uint32_t K, uint32_t base, int hasOuter, float lo1,
                                                     float scale1, int nb1, int bfirst, int blast, uint32_t *hist2, float *amb,
                                                     const LevelState &lv, const SelState &ss, const SelCtl &sc, uint32_t c,
                                                     int hbmPasses, SelSearchSmem &sm, unsigned long long *dbg = nullptr) {
    const bool ok = sel_block_search_core_src(src, K, base, hasOuter, lo1, scale1, nb1, bfirst, blast, hist2, amb, lv, c, sm, dbg);
    if (threadIdx.x == 0) {
        if (!ok) { ss.flag[c] = 1u; atomicAdd(ss.n_flagged, 1u); }
        else {
            const int it = sm.resIt;
            lv.mL[c] = sm.resL; lv.mR[c] = sm.resR; lv.iter[c] = it;
            lv.found[c] = sm.resFnd ? 1u : 0u;
            lv.active[c] = 0u;
            lv.nleft_g[c] = sm.resNleft;
            lv.nleft_l[c] = sm.resNleft;
            const unsigned long long np = (unsigned long long)(lv.bnd[c + 1] - lv.bnd[c]);
            if (np) {
                atomicAdd(sc.active_particles, np * (unsigned long long)hbmPasses);
                atomicAdd(sc.active_particles + 1, np * (unsigned long long)it);
            }
            atomicMax(sc.level_iters, it);
            if (!sm.resFnd) atomicAdd(sc.n_unfound_out, 1u);
        }
    }
    return ok;
}
__device__ __forceinline__ bool sel_block_search(const float *vals, uint32_t K, uint32_t base, int hasOuter, float lo1,
                                                 float scale1, int nb1, int bfirst, int blast, uint32_t *hist2, float *amb,
                                                 const LevelState &lv, const SelState &ss, const SelCtl &sc, uint32_t c,
                                                 int hbmPasses, SelSearchSmem &sm, unsigned long long *dbg = nullptr, bool glob = false) {
    SelSrcArray src;
    src.vals = vals; src.K = K; src.glob = glob;
    return sel_block_search_src(src, K, base, hasOuter, lo1, scale1, nb1, bfirst, blast, hist2, amb, lv, ss, sc, c, hbmPasses, sm, dbg);
}

// dynamic shared memory of the two search kernels: vals[cap + 4] | hist2[kSelBins2] | amb[kSelAmbCap]
__host__ __device__ inline size_t sel_search_smem_bytes(uint32_t cap) { return ((size_t)cap + 4u + kSelBins2 + kSelAmbCap) * 4u; }

// Stage src[0..K) in shared memory with 16-byte asynchronous copies (every thread keeps all its copies in flight, so
// one block can pull a whole cell at memory speed).  src is only 4-byte aligned: the staged array starts `mis` floats
// into sbuf so that shared and global addresses are congruent mod 16; head and tail go through scalar loads.
// Returns the staged array (vals[i] <-> src[i]).  Ends with a block barrier.
__device__ __forceinline__ float *sel_stage_vals(float *sbuf, const float *__restrict__ src, uint32_t K) {
    const uint32_t mis = (uint32_t)((reinterpret_cast<uintptr_t>(src) >> 2) & 3u);
    float *vals = sbuf + mis;
    const uint32_t head = mis ? min(4u - mis, K) : 0u;
    const uint32_t body4 = (K - head) / 4u;
    const float4 *g4 = reinterpret_cast<const float4 *>(src + head);
    float4 *s4 = reinterpret_cast<float4 *>(vals + head);
    for (uint32_t i = threadIdx.x; i < body4; i += blockDim.x) cp_async16(s4 + i, g4 + i);
    cp_async_commit();
    if (threadIdx.x < head) vals[threadIdx.x] = __ldcg(src + threadIdx.x);
    const uint32_t tail0 = head + body4 * 4u;
    if (tail0 + threadIdx.x < K) vals[tail0 + threadIdx.x] = __ldcg(src + tail0 + threadIdx.x);
    cp_async_wait<0>();
    __syncthreads();
    return vals;
}

// level-wide "all blocks done" report of the single-rank search kernels: the last block writes 1 + (cells flagged) into
// mapped pinned memory, which is all the host ever waits for (see orb_build)
struct SelDone {
    uint32_t *done;                 // blocks-finished counter (zeroed per build)
    volatile uint32_t *h_status;
};
__device__ __forceinline__ void sel_report_done(const SelDone &dn, const uint32_t *n_flagged) {
    __syncthreads();
    if (threadIdx.x == 0 && dn.done) {
        __threadfence();
        if (atomicAdd(dn.done, 1u) == gridDim.x - 1u) {
            __threadfence();
            *dn.h_status = *((volatile const uint32_t *)n_flagged) + 1u;
        }
    }
}

// FINISH (streaming regime): one block per cell, candidates from the dense list written by COMPACT - or, with private
// regions (sp.visits: cap = 0, dynamic shared memory hist2 | amb | piece tables), read from the pieces the COMPACT blocks
// left in their own regions
template <bool PRIV>
__global__ void __launch_bounds__(1024) k_sel_finish(const float *__restrict__ cand, LevelState lv, SelState ss, SelCtl sc,
                                                     uint32_t nCells, int nb1, uint32_t cap, int *__restrict__ err,
                                                     unsigned long long *dbg, int hbmPasses /* reads of the column: 2, or 1 when the partition built the rows */,
                                                     int allowGlobal /* more candidates than `cap`: search them where they lie instead of failing */,
                                                     SelDone dn, SelPriv sp) {
    extern __shared__ __align__(16) unsigned char sel_smem[];
    float *sbuf = reinterpret_cast<float *>(sel_smem);
    uint32_t *hist2 = reinterpret_cast<uint32_t *>(sbuf + cap + 4);
    float *amb = reinterpret_cast<float *>(hist2 + kSelBins2);
    __shared__ SelSearchSmem sm;
    __shared__ uint32_t s_below, s_bad;
    pdl_enter();
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(sc.passes_out, hbmPasses);
    unsigned long long *bs = (dbg && blockIdx.x == 0) ? dbg : nullptr;    // ORB_DEBUG_TIMES=2: phases of block 0's first cell
    if (bs && threadIdx.x == 0) bs[0] = gtimer();
    const int tid = threadIdx.x, lane = tid & 31, nThreads = (int)blockDim.x;
    for (uint32_t c = blockIdx.x; c < nCells; c += gridDim.x) {
        __syncthreads();
        // independent loads of the cell's parameters, issued together
        const uint32_t act = lv.active[c], b0 = lv.bnd[c], b1 = lv.bnd[c + 1];
        const uint32_t flg = __ldcg(&ss.flag[c]), cur = __ldcg(&ss.cursor[c]), bs_ = __ldcg(&ss.base[c]);
        uint32_t K = __ldcg(&ss.ncand[c]);
        const uint32_t bf = __ldcg(&ss.bfirst[c]), bl = __ldcg(&ss.blast[c]);
        const float L = lv.mL[c], R = lv.mR[c];
        if (!act) continue;
        if (b1 == b0) {     // empty cell: no block ever resolved it; the search on nothing finds the cut at once
            if (threadIdx.x == 0) ss.flag[c] = 0u;
            sel_block_search(sbuf, 0u, 0u, 0, 0.f, 0.f, 1, 0, 0, hist2, amb, lv, ss, sc, c, hbmPasses, sm);
            continue;
        }
        if (flg) continue;
        if (PRIV) {
            // ---- the cell's pieces: chunks k0..k1 overlap it; each chunk's block left at most one record for the cell.
            //      The search reads the pieces where they lie (L2), nothing is staged. ----
            uint32_t *pOff = reinterpret_cast<uint32_t *>(amb + kSelAmbCap), *pCnt = pOff + kSelMaxPieces, *pPos = pCnt + kSelMaxPieces;
            const uint32_t k0 = b0 / sp.chunk, k1 = (b1 - 1u) / sp.chunk, nP = k1 - k0 + 1u;
            if (tid == 0) { s_below = 0u; s_bad = nP > (uint32_t)kSelMaxPieces ? 1u : 0u; }
            __syncthreads();
            uint32_t myBelow = 0;
            if (!s_bad) {
                for (uint32_t i = tid; i < nP; i += nThreads) {
                    const SelVisitRec *r = sp.visits + (k0 + i);
                    const uint32_t n = __ldcg(&r->n);
                    uint32_t off = 0, cnt = 0;
                    if (n > (uint32_t)kSelMaxVisit) s_bad = 1u;      // the block dropped a record: this cell's may be the one
                    for (uint32_t j = 0; j < min(n, (uint32_t)kSelMaxVisit); ++j) {
                        if (__ldcg(&r->v[j].cell) == c) { off = __ldcg(&r->v[j].off); cnt = __ldcg(&r->v[j].cnt); myBelow += __ldcg(&r->v[j].below); }
                    }
                    pOff[i] = off; pCnt[i] = cnt;
                }
            }
            myBelow = __reduce_add_sync(0xffffffffu, myBelow);
            if (lane == 0 && myBelow) atomicAdd(&s_below, myBelow);
            __syncthreads();
            if (s_bad) {        // block-uniform
                if (tid == 0) { ss.flag[c] = 1u; atomicAdd(ss.n_flagged, 1u); }
                continue;
            }
            uint32_t carry = 0;
            for (uint32_t r0 = 0; r0 < nP; r0 += (uint32_t)nThreads) {       // exclusive scan of the piece sizes
                const uint32_t i = r0 + (uint32_t)tid;
                const uint32_t v = i < nP ? pCnt[i] : 0u;
                uint32_t tot;
                const uint32_t ex = sel_block_scan(v, sm.w, tot);
                if (i < nP) pPos[i] = carry + ex;
                carry += tot;
            }
            __syncthreads();
            K = carry;
            const uint32_t base = s_below;
            // the exact numbers must prove the bracket the (sampled) rows suggested - see SelSampleEst
            SelTarget tg;
            tg.init(lv.total[c], lv.nleaf[c]);
            const bool proven = (bf == 0u || tg.diff(base) <= -3) && (bl + 1u >= (uint32_t)nb1 || tg.diff(base + K) >= 3);
            if (!proven || bf > bl) {
                if (tid == 0) { ss.flag[c] = 1u; atomicAdd(ss.n_flagged, 1u); }
                continue;
            }
            SelSrcPieces src;
            src.cand = cand; src.pOff = pOff; src.pCnt = pCnt; src.pPos = pPos; src.K = K;
            if (bs && threadIdx.x == 0) bs[1] = gtimer();
            sel_block_search_src(src, K, base, 1, L, sel_scale(L, R, nb1), nb1, (int)bf, (int)bl, hist2, amb, lv, ss, sc, c, hbmPasses, sm,
                                 c == blockIdx.x ? bs : nullptr);
            if (bs && threadIdx.x == 0 && c == blockIdx.x) bs[7] = gtimer();
            continue;
        }
        if (cur != K || (K > cap && !allowGlobal)) {   // cannot happen: HIST and COMPACT use the same bin function
            if (threadIdx.x == 0) atomicExch(err, ORB_ERR_STATE);
            continue;
        }
        // Normally the candidates are staged in shared memory.  With allowGlobal (experimental, ORB_SELECT_BIG_FINISH=1)
        // COMPACT gathers any number of them (the list has room for the whole cell) and a dense bin's candidates are
        // searched in place: the block search only needs a pointer it can read a few times.
        const float *vals = K <= cap ? sel_stage_vals(sbuf, cand + b0, K) : cand + b0;
        if (threadIdx.x == 0) ss.cursor[c] = 0u;      // zero between levels (HIST counts into it)
        if (bs && threadIdx.x == 0) bs[1] = gtimer();
        sel_block_search(vals, K, bs_, 1, L, sel_scale(L, R, nb1), nb1, (int)bf, (int)bl, hist2, amb, lv, ss, sc, c, hbmPasses, sm,
                         c == blockIdx.x ? bs : nullptr, K > cap);
        if (bs && threadIdx.x == 0 && c == blockIdx.x) bs[7] = gtimer();
    }
    sel_report_done(dn, ss.n_flagged);
}

// ---- parallel FINISH: fine histograms of the candidates, one block per COMPACT block ----
__global__ void __launch_bounds__(kThreads) k_sel_fine(const float *__restrict__ cand, SelPriv sp, SelPar pr, uint32_t nBlocks) {
    __shared__ uint32_t s_h[kSelBins2];
    pdl_enter();
    for (uint32_t b = blockIdx.x; b < nBlocks; b += gridDim.x) {
        const SelVisitRec *r = sp.visits + b;
        const uint32_t n = min(__ldcg(&r->n), (uint32_t)kSelMaxVisit);
        for (uint32_t j = 0; j < n; ++j) {
            const uint32_t c = __ldcg(&r->v[j].cell), off = __ldcg(&r->v[j].off), cnt = __ldcg(&r->v[j].cnt);
            if (!cnt) continue;         // block-uniform
            const float lo2 = __ldcg(&pr.lo2[c]), sc2 = __ldcg(&pr.sc2[c]);
            __syncthreads();
            for (int i = threadIdx.x; i < kSelBins2; i += kThreads) s_h[i] = 0u;
            __syncthreads();
            uint32_t i = threadIdx.x;
            for (; i + 3u * kThreads < cnt; i += 4u * kThreads) {      // four loads in flight per thread
                float q[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) q[u] = __ldcg(cand + off + i + (uint32_t)u * kThreads);
#pragma unroll
                for (int u = 0; u < 4; ++u) atomicAdd(&s_h[sel_bin(q[u], lo2, sc2, kSelBins2)], 1u);
            }
            for (; i < cnt; i += kThreads) atomicAdd(&s_h[sel_bin(__ldcg(cand + off + i), lo2, sc2, kSelBins2)], 1u);
            __syncthreads();
            for (int k = threadIdx.x; k < kSelBins2; k += kThreads) {
                const uint32_t v = s_h[k];
                if (v) atomicAdd(&pr.fine[(size_t)c * kSelBins2 + k], v);
            }
        }
    }
}

// ---- parallel FINISH, step A ----
__global__ void __launch_bounds__(kThreads) k_sel_fin_a(LevelState lv, SelState ss, SelPriv sp, SelPar pr, uint32_t nCells, int nb1) {
    __shared__ uint32_t s_w[32];
    __shared__ uint32_t s_below, s_cnt, s_bad, s_base2, s_end2;
    __shared__ int s_first, s_last;
    pdl_enter();
    const int tid = threadIdx.x, lane = tid & 31;
    constexpr int per = kSelBins2 / kThreads;        // 8
    for (uint32_t c = blockIdx.x; c < nCells; c += gridDim.x) {
        __syncthreads();
        const uint32_t act = lv.active[c], b0 = lv.bnd[c], b1 = lv.bnd[c + 1];
        const uint32_t flg = __ldcg(&ss.flag[c]);
        const uint32_t bf = __ldcg(&ss.bfirst[c]), bl = __ldcg(&ss.blast[c]);
        if (!act || flg || b1 == b0) continue;        // (empty cells: k_sel_fin_b searches on nothing)
        const uint32_t k0 = b0 / sp.chunk, k1 = (b1 - 1u) / sp.chunk, nP = k1 - k0 + 1u;
        if (tid == 0) { s_below = 0u; s_cnt = 0u; s_bad = 0u; s_first = kSelBins2; s_last = -1; s_base2 = 0u; s_end2 = 0u; }
        __syncthreads();
        uint32_t myBelow = 0, myCnt = 0;
        for (uint32_t i = tid; i < nP; i += kThreads) {
            const SelVisitRec *r = sp.visits + (k0 + i);
            const uint32_t n = __ldcg(&r->n);
            if (n > (uint32_t)kSelMaxVisit) s_bad = 1u;
            for (uint32_t j = 0; j < min(n, (uint32_t)kSelMaxVisit); ++j)
                if (__ldcg(&r->v[j].cell) == c) { myCnt += __ldcg(&r->v[j].cnt); myBelow += __ldcg(&r->v[j].below); }
        }
        myBelow = __reduce_add_sync(0xffffffffu, myBelow);
        myCnt = __reduce_add_sync(0xffffffffu, myCnt);
        if (lane == 0) { if (myBelow) atomicAdd(&s_below, myBelow); if (myCnt) atomicAdd(&s_cnt, myCnt); }
        __syncthreads();
        const uint32_t base = s_below, K = s_cnt;
        SelTarget tg;
        tg.init(lv.total[c], lv.nleaf[c]);
        const bool proven = (bf == 0u || tg.diff(base) <= -3) && (bl + 1u >= (uint32_t)nb1 || tg.diff(base + K) >= 3);
        // ---- scan of the fine histogram (it holds exactly the K candidates) ----
        uint32_t h[per], sum = 0;
#pragma unroll
        for (int j = 0; j < per; ++j) { h[j] = __ldcg(&pr.fine[(size_t)c * kSelBins2 + tid * per + j]); sum += h[j]; }
        uint32_t total;
        const uint32_t excl = sel_block_scan(sum, s_w, total);
        int myFirst = kSelBins2, myLast = -1;
        {
            uint32_t p = base + excl;
#pragma unroll
            for (int j = 0; j < per; ++j) {
                const uint32_t pn = p + h[j];
                const int b = tid * per + j;
                if (tg.diff(pn) > -3) myFirst = min(myFirst, b);
                if (tg.diff(p) < 3) myLast = max(myLast, b);
                p = pn;
            }
        }
        if (myFirst < kSelBins2) atomicMin(&s_first, myFirst);
        if (myLast >= 0) atomicMax(&s_last, myLast);
        __syncthreads();
        const int f2 = s_first, l2 = s_last;
        {
            uint32_t p = base + excl;
#pragma unroll
            for (int j = 0; j < per; ++j) {
                const int b = tid * per + j;
                if (b == f2) s_base2 = p;
                p += h[j];
                if (b == l2) s_end2 = p;
            }
        }
        __syncthreads();
        const uint32_t base2 = s_base2, K2 = s_end2 - s_base2;
        // not proven / a dropped visit record / fine histogram incomplete / massive ties: the iterative search takes the cell
        const bool ok = proven && !s_bad && bf <= bl && total == K && f2 <= l2 && K2 <= (uint32_t)kSelAmbCap;
        if (tid < 32) {
            float a = 0.f, b = 0.f;
            if (ok) sel_value_bounds_warp((uint32_t)f2, (uint32_t)l2, __ldcg(&pr.lo2[c]), __ldcg(&pr.sc2[c]), kSelBins2, a, b);
            if (tid == 0) {
                if (!ok) { ss.flag[c] = 1u; atomicAdd(ss.n_flagged, 1u); }
                pr.f2[c] = f2; pr.l2[c] = l2; pr.base2[c] = base2; pr.k2[c] = K2;
                pr.v2lo[c] = ok ? a : __int_as_float(0x7f800000);       // (flagged: empty range, k_sel_gather appends nothing)
                pr.v2hi[c] = ok ? b : __int_as_float(0xff800000);
            }
        }
    }
}

// ---- step B: every COMPACT block's pieces again; values inside the ambiguous fine bins of their cell go to its list ----
__global__ void __launch_bounds__(kThreads) k_sel_gather(const float *__restrict__ cand, SelPriv sp, SelPar pr, uint32_t nBlocks) {
    pdl_enter();
    for (uint32_t b = blockIdx.x; b < nBlocks; b += gridDim.x) {
        const SelVisitRec *r = sp.visits + b;
        const uint32_t n = min(__ldcg(&r->n), (uint32_t)kSelMaxVisit);
        for (uint32_t j = 0; j < n; ++j) {
            const uint32_t c = __ldcg(&r->v[j].cell), off = __ldcg(&r->v[j].off), cnt = __ldcg(&r->v[j].cnt);
            const float vLo = __ldcg(&pr.v2lo[c]), vHi = __ldcg(&pr.v2hi[c]);
            float *dst = pr.amb + (size_t)c * kSelAmbCap;
            uint32_t i = threadIdx.x;
            for (; i + 3u * kThreads < cnt; i += 4u * kThreads) {      // four loads in flight per thread
                float q[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) q[u] = __ldcg(cand + off + i + (uint32_t)u * kThreads);
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (sel_is_cand(q[u], vLo, vHi)) { const uint32_t p = atomicAdd(&pr.ambCnt[c], 1u); if (p < (uint32_t)kSelAmbCap) dst[p] = q[u]; }
            }
            for (; i < cnt; i += kThreads) {
                const float q = __ldcg(cand + off + i);
                if (sel_is_cand(q, vLo, vHi)) { const uint32_t p = atomicAdd(&pr.ambCnt[c], 1u); if (p < (uint32_t)kSelAmbCap) dst[p] = q; }
            }
        }
    }
}

// ---- step C: the block search of every cell on its short list; the fine bins are the search's outer histogram ----
__global__ void __launch_bounds__(kThreads) k_sel_fin_b(LevelState lv, SelState ss, SelCtl sc, SelPar pr, uint32_t nCells, int *__restrict__ err,
                                                         int hbmPasses, SelDone dn) {
    __shared__ float s_vals[kSelAmbCap + 4];
    __shared__ uint32_t s_hist2[kSelBins2];
    __shared__ float s_amb[kSelAmbCap];
    __shared__ SelSearchSmem sm;
    pdl_enter();
    const int tid = threadIdx.x;
    if (blockIdx.x == 0 && tid == 0) atomicAdd(sc.passes_out, hbmPasses);
    for (uint32_t c = blockIdx.x; c < nCells; c += gridDim.x) {
        __syncthreads();
        const uint32_t act = lv.active[c], b0 = lv.bnd[c], b1 = lv.bnd[c + 1];
        if (!act) continue;
        if (b1 == b0) {     // empty cell: the search on nothing finds the cut at once
            if (tid == 0) ss.flag[c] = 0u;
            sel_block_search(s_vals, 0u, 0u, 0, 0.f, 0.f, 1, 0, 0, s_hist2, s_amb, lv, ss, sc, c, hbmPasses, sm);
            continue;
        }
        if (__ldcg(&ss.flag[c])) continue;
        const uint32_t K2 = __ldcg(&pr.k2[c]), got = __ldcg(&pr.ambCnt[c]);
        if (got != K2 || K2 > (uint32_t)kSelAmbCap) {     // cannot happen: the fine histogram and the gather use one bin function
            if (tid == 0) atomicExch(err, ORB_ERR_STATE);
            continue;
        }
        for (uint32_t i = tid; i < K2; i += kThreads) s_vals[i] = __ldcg(pr.amb + (size_t)c * kSelAmbCap + i);
        __syncthreads();
        sel_block_search(s_vals, K2, __ldcg(&pr.base2[c]), 1, __ldcg(&pr.lo2[c]), __ldcg(&pr.sc2[c]), kSelBins2, __ldcg(&pr.f2[c]), __ldcg(&pr.l2[c]),
                         s_hist2, s_amb, lv, ss, sc, c, hbmPasses, sm);
    }
    sel_report_done(dn, ss.n_flagged);
}

// =====================================================================================
// Whole search of a cell by ONE block (levels with many cells): HIST of the cell into a shared-memory histogram,
// RESOLVE, second read of the cell (it is still in L2) that keeps the candidates in shared memory, FINISH.  Nothing
// but the result goes back to global memory; several blocks per SM overlap each other's memory phases.
// dynamic shared memory: hist[kSelBins2] (reused as the refinement histogram) | list[candCap] | amb[kSelAmbCap]
// =====================================================================================
__host__ __device__ inline size_t sel_percell_smem_bytes(uint32_t candCap) { return ((size_t)kSelBins2 + candCap + kSelAmbCap) * 4u; }

constexpr int kSelMaxSeg = 8;     // chunk segments of one cell for which k_sel_percell records left counts
struct SelPerCellSmem {
    SelSearchSmem search;
    uint32_t w[32];
    int first, last;
    uint32_t base, end, nlist, lowTot, pA, pB;
    float vLo, vHi;
    uint32_t segBelow[kSelMaxSeg], segListEnd[kSelMaxSeg], segLeft[kSelMaxSeg];
};

// apply f(value) to every element of src[0..K): 16-byte loads where src is aligned, U loads in flight per thread
// (a block that has an SM to itself needs U = 8 to keep enough bytes in flight for its share of the HBM bandwidth)
template <int U, typename F>
__device__ __forceinline__ void sel_for_each(const float *__restrict__ src, uint32_t K, F f) {
    const uint32_t mis = (uint32_t)((reinterpret_cast<uintptr_t>(src) >> 2) & 3u);
    const uint32_t head = mis ? min(4u - mis, K) : 0u;
    const uint32_t body4 = (K - head) / 4u;
    const float4 *g4 = reinterpret_cast<const float4 *>(src + head);
    const uint32_t nT = blockDim.x;
    uint32_t i = threadIdx.x;
    for (; i + (uint32_t)(U - 1) * nT < body4; i += (uint32_t)U * nT) {
        float4 q[U];
#pragma unroll
        for (int u = 0; u < U; ++u) q[u] = __ldg(g4 + i + (uint32_t)u * nT);
#pragma unroll
        for (int u = 0; u < U; ++u) { f(q[u].x); f(q[u].y); f(q[u].z); f(q[u].w); }
    }
    for (; i < body4; i += nT) { const float4 a = __ldg(g4 + i); f(a.x); f(a.y); f(a.z); f(a.w); }
    if (threadIdx.x < head) f(__ldg(src + threadIdx.x));
    const uint32_t tail0 = head + body4 * 4u;
    if (tail0 + threadIdx.x < K) f(__ldg(src + tail0 + threadIdx.x));
}

// COMPACT of a cell one block searches: counts the values below vLo (returned per thread) and appends those in
// [vLo, vHi) to `list` (fill level *nlist; values beyond cap are counted but not stored).  One shared-memory atomic per
// float4 that holds candidates, not per candidate.
template <int U>
__device__ __forceinline__ unsigned sel_compact_pass(const float *__restrict__ src, uint32_t K, float vLo, float vHi, float *list,
                                                     uint32_t *nlist, uint32_t cap) {
    unsigned lows = 0u;
    auto one = [&](float v) {
        lows += sel_is_low(v, vLo) ? 1u : 0u;
        if (sel_is_cand(v, vLo, vHi)) {
            const uint32_t idx = atomicAdd(nlist, 1u);
            if (idx < cap) list[idx] = v;
        }
    };
    auto four = [&](const float4 q) {
        const float v[4] = {q.x, q.y, q.z, q.w};
        // #{v >= vLo} and #{v >= vHi}: they differ iff the float4 holds a candidate (vHi = NaN compares false)
        unsigned nGeLo = 0u, nGeHi = 0u;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            nGeLo += (v[k] >= vLo) ? 1u : 0u;
            nGeHi += (v[k] >= vHi) ? 1u : 0u;
        }
        lows += 4u - nGeLo;
        if (nGeLo != nGeHi) {
            unsigned keep = 0u;
#pragma unroll
            for (int k = 0; k < 4; ++k) keep |= (unsigned)sel_is_cand(v[k], vLo, vHi) << k;
            uint32_t idx = atomicAdd(nlist, (uint32_t)__popc(keep));
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (keep & (1u << k)) { if (idx < cap) list[idx] = v[k]; ++idx; }
        }
    };
    const uint32_t mis = (uint32_t)((reinterpret_cast<uintptr_t>(src) >> 2) & 3u);
    const uint32_t head = mis ? min(4u - mis, K) : 0u;
    const uint32_t body4 = (K - head) / 4u;
    const float4 *g4 = reinterpret_cast<const float4 *>(src + head);
    const uint32_t nT = blockDim.x;
    uint32_t i = threadIdx.x;
    for (; i + (uint32_t)(U - 1) * nT < body4; i += (uint32_t)U * nT) {
        float4 q[U];
#pragma unroll
        for (int u = 0; u < U; ++u) q[u] = __ldg(g4 + i + (uint32_t)u * nT);
#pragma unroll
        for (int u = 0; u < U; ++u) four(q[u]);
    }
    for (; i < body4; i += nT) four(__ldg(g4 + i));
    if (threadIdx.x < head) one(__ldg(src + threadIdx.x));
    const uint32_t tail0 = head + body4 * 4u;
    if (tail0 + threadIdx.x < K) one(__ldg(src + tail0 + threadIdx.x));
    return lows;
}

// sampled variant of sel_for_each: every S-th 512-byte piece (32 float4) of the 16-byte aligned body of src[0..K)
template <int U, typename F>
__device__ __forceinline__ void sel_for_each_sampled(const float *__restrict__ src, uint32_t K, uint32_t S, F f) {
    const uint32_t mis = (uint32_t)((reinterpret_cast<uintptr_t>(src) >> 2) & 3u);
    const uint32_t head = mis ? min(4u - mis, K) : 0u;
    const uint32_t nPieces = ((K - head) / 4u) >> 5;                // whole pieces
    const uint32_t nSamp = (nPieces + S - 1u) / S;                  // pieces 0, S, 2S, ...
    const float4 *g4 = reinterpret_cast<const float4 *>(src + head);
    const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5, nW = blockDim.x >> 5;
    for (uint32_t j = wid; j < nSamp; j += nW * (uint32_t)U) {
        float4 q[U];
        bool in[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t jj = j + (uint32_t)u * nW;
            in[u] = jj < nSamp;
            q[u] = in[u] ? __ldg(g4 + (size_t)jj * S * 32u + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (in[u]) { f(q[u].x); f(q[u].y); f(q[u].z); f(q[u].w); }
    }
}

// Smaller cells are searched with exact rows: the sample's margin would make a tenth or more of such a cell a
// candidate, and its second read comes from L2 anyway.
constexpr uint32_t kSelSampleMinCell = 65536;      // (default of ORB_SAMPLE_MIN_CELL)
// bins of a block's own histogram of a cell of K particles: about 64 particles per bin, at least one bin per thread
__device__ __forceinline__ int sel_percell_bins(uint32_t K, int nThreads) {
    const int nb = K > 131072u ? kSelBins2 : (K > 32768u ? 1024 : (K > 8192u ? 512 : 256));
    return max(nb, nThreads);
}

template <int THREADS, int MINBLOCKS, int U = 4, bool SMP = false /* first attempt with sampled rows compiled in */>
__global__ void __launch_bounds__(THREADS, MINBLOCKS) k_sel_percell(const float *__restrict__ x, const float *__restrict__ y,
                                                                    const float *__restrict__ z, LevelState lv, SelState ss,
                                                                    SelCtl sc, uint32_t nCells, uint32_t candCap, int preNb, SelDone dn,
                                                                    PreLeft *__restrict__ pre, uint32_t preTag,
                                                                    uint32_t chunk /* particles per partition block; 0: no records */,
                                                                    uint32_t nLocal, int sampleS /* > 1: first attempt with sampled rows */,
                                                                    float sampleZ, uint32_t sampleMinCell) {
    extern __shared__ __align__(16) unsigned char sel_smem[];
    uint32_t *hist = reinterpret_cast<uint32_t *>(sel_smem);
    float *list = reinterpret_cast<float *>(hist + kSelBins2);
    float *amb = list + candCap;
    __shared__ SelPerCellSmem sm;
    pdl_enter();
    const int tid = threadIdx.x, nThreads = (int)blockDim.x;
    if (blockIdx.x == 0 && tid == 0) atomicAdd(sc.passes_out, preNb ? 1 : (sampleS > 1 ? 1 : 2));
    for (uint32_t c = blockIdx.x; c < nCells; c += gridDim.x) {
        __syncthreads();
        const uint32_t act = lv.active[c], b = lv.bnd[c], K = lv.bnd[c + 1] - b;
        const int ax = lv.axis[c];
        const float L = lv.mL[c], R = lv.mR[c];
        if (tid == 0) ss.flag[c] = 0u;
        if (!act) continue;
        const float *col = pick_col(ax, x, y, z) + b;
        SelTarget tg;
        tg.init(lv.total[c], lv.nleaf[c]);
        // Attempt 0 (sampleS > 1): rows from a sample of the cell - one read of 1/S of it; the gathering read then counts
        // the particles below the candidate bins exactly and the bracket must be proven by those numbers (SelSampleEst).
        // Attempt 1: exact rows, as many zoom rounds as it takes.
        for (int attempt = (SMP && sampleS > 1 && !preNb && K >= sampleMinCell) ? 0 : 1; attempt < 2; ++attempt) {
            const bool smp = SMP && attempt == 0;
            // Rounds: HIST (bins over [lo, lo + nb/scale)), RESOLVE; if the candidate bins hold more particles than the
            // block stages (a dense clump inside a wide cell), zoom the bin function onto them and go again - any monotone
            // bin function over the whole cell is valid, so a round needs nothing from the previous one but the interval.
            // Round 0 with preNb != 0: the cell's row (preNb bins over the margins) was built by the partition of the
            // previous level (NextHist) and is loaded instead of read from the particles.  The scan runs over
            // nbScan >= nb bins; bins beyond nb are empty (an empty trailing bin never becomes the first candidate bin;
            // as the last one it only means "up to the end", which sel_bin_bounds and the search treat the same way).
            int nb = preNb ? preNb : sel_percell_bins(K, nThreads);
            float lo = L, scale = sel_scale(L, R, nb);
            int first = 0, last = -1;
            uint32_t base = 0, K2 = 0;
            bool ok = false;
            int reads = 1;                                        // of the cell: HIST rounds + COMPACT
            for (int round = 0; round < (smp ? 1 : 3); ++round) {
                const int nbScan = max(nb, nThreads);
                const int per = nbScan / nThreads;                // 1, 2 or 8
                const float nbm1 = (float)(nb - 1);
                __syncthreads();
                if (preNb && round == 0) for (int i = tid; i < nbScan; i += nThreads) hist[i] = i < nb ? __ldcg(ss.hist + (size_t)c * nb + i) : 0u;
                else for (int i = tid; i < nbScan; i += nThreads) hist[i] = 0u;
                if (tid == 0) { sm.first = nbScan; sm.last = -1; sm.base = 0u; sm.end = 0u; sm.nlist = 0u; sm.lowTot = 0u; }
                __syncthreads();
                // ---- HIST ----
                if (!(preNb && round == 0)) {
                    auto bin = [&](float v) {
                        const float t = fminf(fmaxf(__fmul_rn(__fsub_rn(v, lo), scale), 0.f), nbm1);      // == sel_bin(v, lo, scale, nb)
                        atomicAdd(&hist[__float2int_rz(t)], 1u);
                    };
                    if (smp) sel_for_each_sampled<U>(col, K, (uint32_t)sampleS, bin);
                    else { ++reads; sel_for_each<U>(col, K, bin); }
                    __syncthreads();
                }
                // ---- RESOLVE ----
                uint32_t h[8];
                uint32_t sum = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) { h[j] = (j < per) ? hist[tid * per + j] : 0u; sum += h[j]; }
                uint32_t total;
                const uint32_t excl = sel_block_scan(sum, sm.w, total);
                SelSampleEst se;
                se.init(total, K, sampleZ);
                uint32_t pA = 0u, pB = 0u;
                if (smp) {      // block-uniform: critical sample prefixes by one warp
                    if (tid < 32) {
                        sel_sample_crit(se, tg, total, pA, pB);
                        if (tid == 0) { sm.pA = pA; sm.pB = pB; }
                    }
                    __syncthreads();
                    pA = sm.pA; pB = sm.pB;
                }
                int myFirst = nbScan, myLast = -1;
                {
                    uint32_t p = excl;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (j < per) {
                            const uint32_t pn = p + h[j];
                            const int bb = tid * per + j;
                            if (smp ? (pn >= pA) : (tg.diff(pn) > -3)) myFirst = min(myFirst, bb);
                            if (smp ? (p <= pB) : (tg.diff(p) < 3)) myLast = max(myLast, bb);
                            p = pn;
                        }
                    }
                }
                if (myFirst < nbScan) atomicMin(&sm.first, myFirst);
                if (myLast >= 0) atomicMax(&sm.last, myLast);
                __syncthreads();
                first = sm.first; last = sm.last;
                {
                    uint32_t p = excl;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (j < per) {
                            const int bb = tid * per + j;
                            if (bb == first) sm.base = p;
                            p += h[j];
                            if (bb == last) sm.end = p;
                        }
                    }
                }
                __syncthreads();
                base = sm.base; K2 = sm.end - base;
                if (!(first <= last)) break;                    // cannot happen (tests/test_select_model.py::ambiguous_range)
                if (smp) {      // estimated number of candidates, with headroom: the list must not overflow
                    const float est = (float)K2 * se.sc;
                    ok = est * 1.125f + 64.f <= (float)candCap;
                    break;
                }
                if (K2 <= candCap) { ok = true; break; }
                // ---- zoom onto the candidate bins (widened by 1/16 of their width on either side) ----
                const float a = __fadd_rn(lo, __fdiv_rn((float)first, scale));
                const float e = __fadd_rn(lo, __fdiv_rn((float)min(last + 1, nb), scale));
                const float w = __fsub_rn(e, a);
                const int nbNew = max(kSelBins2, nThreads);
                const float loNew = __fsub_rn(a, w * 0.0625f);
                const float scNew = sel_scale(loNew, __fadd_rn(e, w * 0.0625f), nbNew);
                if (!(scale > 0.f) || !(w > 0.f) || !(scNew > scale)) break;     // degenerate box / no resolution left: ties
                lo = loNew; scale = scNew; nb = nbNew;
            }
            if (!ok) {
                if (smp) continue;       // the sample's candidate bins are too wide for the block: exact rows
                // too many candidates for one block even after zooming (massive ties): the iterative search takes the cell
                if (tid == 0) { ss.flag[c] = 1u; atomicAdd(ss.n_flagged, 1u); }
                break;
            }
            // ---- COMPACT: another read (L2), candidates into shared memory ----
            if (tid < 32) {
                float a, bb;
                sel_value_bounds_warp((uint32_t)first, (uint32_t)last, lo, scale, nb, a, bb);
                if (tid == 0) { sm.vLo = a; sm.vHi = bb; }
            }
            __syncthreads();
            const float vLo = sm.vLo, vHi = sm.vHi;
            // With `pre`: the cell is read segment by segment (pieces between the partition's chunk boundaries), so that the
            // left count of every chunk that ends inside this cell is known afterwards: particles below the candidate bins
            // + the segment's candidates left of the cut.
            const uint32_t e = b + K;
            int nSeg = 1;
            if (pre && chunk) {
                const uint32_t k0 = b / chunk, k1 = (e - 1u) / chunk;      // chunks of the first / last particle (K > 0 here)
                nSeg = (int)(k1 - k0 + 1u);
            }
            const bool segs = pre && chunk && nSeg <= kSelMaxSeg && K > 0u;
            if (!segs) nSeg = 1;
            if (segs) {
                if (tid < kSelMaxSeg) sm.segBelow[tid] = 0u;
                __syncthreads();
            }
            for (int sgi = 0; sgi < nSeg; ++sgi) {
                uint32_t s0 = b, s1 = e;
                if (segs) {
                    const uint32_t k = b / chunk + (uint32_t)sgi;
                    s0 = max(b, k * chunk);
                    s1 = min(e, (k + 1u) * chunk);
                }
                // (exact rows never overfill the list; a sampled estimate may)
                unsigned lows = sel_compact_pass<U>(col + (s0 - b), s1 - s0, vLo, vHi, list, &sm.nlist, candCap);
                if (segs || smp) {
                    lows = __reduce_add_sync(0xffffffffu, lows);
                    if (smp && (tid & 31) == 0 && lows) atomicAdd(&sm.lowTot, lows);
                }
                if (segs) {
                    __syncthreads();                                  // the segment's candidates are all in the list
                    if ((tid & 31) == 0 && lows) atomicAdd(&sm.segBelow[sgi], lows);
                    if (tid == 0) sm.segListEnd[sgi] = sm.nlist;
                    __syncthreads();                                  // ... before anyone appends the next segment's
                }
            }
            __syncthreads();
            if (smp) {
                // exact numbers of the gathering read: they must fit the list and prove the bracket, else exact rows
                base = sm.lowTot; K2 = sm.nlist;
                const bool proven = (first == 0 || tg.diff(base) <= -3) && (last + 1 >= nb || tg.diff(base + K2) >= 3);
                if (K2 > candCap || !proven) continue;
            }
            // ---- FINISH ----
            const bool done = sel_block_search(list, K2, base, 1, lo, scale, nb, first, last, hist, amb, lv, ss, sc, c, reads, sm.search);
            if (segs && done) {       // block-uniform
                const float cutf = mid_cut(sm.search.resL, sm.search.resR);
                if (tid < kSelMaxSeg) sm.segLeft[tid] = 0u;
                __syncthreads();
                for (int sgi = 0; sgi < nSeg; ++sgi) {
                    const uint32_t l0 = sgi ? sm.segListEnd[sgi - 1] : 0u, l1 = sm.segListEnd[sgi];
                    uint32_t m = 0;
                    for (uint32_t i = l0 + tid; i < l1; i += nThreads) m += (list[i] < cutf) ? 1u : 0u;
                    m = __reduce_add_sync(0xffffffffu, m);
                    if ((tid & 31) == 0 && m) atomicAdd(&sm.segLeft[sgi], m);
                }
                __syncthreads();
                if (tid < nSeg) {
                    const uint32_t k = b / chunk + (uint32_t)tid;
                    const uint32_t segEnd = min(e, (k + 1u) * chunk);
                    // the segment is chunk k's trailing segment iff it reaches the chunk's end (the last chunk ends at nLocal)
                    if (segEnd == (k + 1u) * chunk || segEnd == nLocal) {
                        PreLeft P;
                        P.tag = preTag; P.cell = c; P.below = sm.segBelow[tid] + sm.segLeft[tid]; P.gbase = 0u; P.tot = 0u; P.kind = 1u;
                        P.pad_[0] = P.pad_[1] = 0u;
                        pre[k] = P;
                    }
                }
            }
            break;
        }
    }
    sel_report_done(dn, ss.n_flagged);
}

// =====================================================================================
// Several ranks (particles sharded).  The search needs only TWO exchanges per level (the iterative
// search needs one per bisection pass): the histogram rows are summed over ranks, the candidates are gathered, and
// every rank runs the same block search over the candidates of all ranks - it replays the reference's decisions on
// identical data and ends with bit-identical margins / iterations / global counts.  The local left count (the
// partition's split offset) is #{own particles below the candidate bins} + #{own candidates < final cut}.
// Everything that decides whether a cell is flagged derives from exchanged data, so all ranks flag the same cells.
//
// Two transports:
//  * NVLink peer memory (SelPeers; default when the peer table is imported): no collective call at all.  Every rank
//    keeps its histogram rows, candidate slots and slot fill levels in an exchange arena that all ranks map.
//      HIST -> k_selx_resolve [signal + wait: all ranks' HIST done; block per cell sums the rows of all ranks with
//      remote loads, resolves, publishes] -> COMPACT (reads the published bins, own candidates into the cell's slot)
//      -> k_selmr_finish<true> [signal + wait: all ranks' COMPACT done; block per cell pulls the candidates of all
//      ranks straight out of their slots, searches].
//    The signal is a remote store of the exchange's sequence number into every peer's flag word, issued by the first
//    block of the kernel that FOLLOWS the producing kernel in the stream (so the producer has completed); the wait
//    spins on the rank's own flag words.  One buffer of each kind suffices: a rank overwrites rows / slots / fill
//    levels of level l only after an exchange that every peer reaches after it has finished reading level l.
//  * NCCL (peer table not imported): HIST -> allreduce(rows) -> COMPACT (+ per-block resolve) -> k_selmr_prep
//    (publishes every cell's resolve, the slot's count word) -> all-gather(slots) -> k_selmr_finish<false>.
// =====================================================================================
struct SelPeers {
    int n, self;                  // n == 0: NCCL transport
    uint32_t seq;                 // sequence number of the kernel's exchange (monotone, identical on all ranks)
    uint32_t *arena[kMaxPeers];   // every rank's exchange arena: flags[kMaxPeers] | ... (word offsets below, rank-invariant)
    uint32_t offCursor, offSlots, offHist;
};
struct SelMrState {
    const uint32_t *hist_l;   // [nCells][nb1] this rank's histogram rows
    uint32_t *loc_base;       // [nCells] local particles in bins below the candidate bins
    float *slots_l;           // [nCells][slotWords] own candidates (NCCL: word slotWords-1 = their count, uint32 bits)
    const float *slots_g;     // NCCL: [nRanks][nCells][slotWords] all ranks' slots after the all-gather
    uint32_t slotWords;
    int nRanks, self;
    uint32_t *done;           // finish: blocks-finished counter of the level (zeroed per build)
    volatile uint32_t *h_status;   // finish: mapped pinned word, receives 1 + cells flagged at this level
};

// Cross-rank barrier at the start of a kernel (after pdl_enter: this rank's preceding kernel has completed).
__device__ __forceinline__ void selx_barrier(const SelPeers &px) {
    const int t = threadIdx.x;
    if (blockIdx.x == 0 && t < px.n && t != px.self) {
        __threadfence_system();
        asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(px.arena[t] + px.self), "r"(px.seq) : "memory");
    }
    if (t < px.n && t != px.self) {
        const uint32_t *f = px.arena[px.self] + t;
        while ((int32_t)(ld_sys_u32(f) - px.seq) < 0) {}
        __threadfence_system();
    }
    __syncthreads();
}

// peer transport: rows of all ranks -> resolve of every cell of the level
__global__ void __launch_bounds__(kThreads) k_selx_resolve(LevelState lv, SelState ss, SelMrState mr, SelPeers px, uint32_t nCells,
                                                            int nb1, uint32_t candCap) {
    extern __shared__ __align__(16) unsigned char sel_smem[];
    uint32_t *hbuf = reinterpret_cast<uint32_t *>(sel_smem);
    __shared__ SelResolveSmem rs;
    __shared__ uint32_t s_red[kWarps];
    pdl_enter();
    selx_barrier(px);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (uint32_t c = blockIdx.x; c < nCells; c += gridDim.x) {
        __syncthreads();
        // every peer is past its finish of the previous level (it has signalled this exchange): the fill level may go
        if (tid == 0) ss.cursor[c] = 0u;
        if (!lv.active[c]) {          // block-uniform
            if (tid == 0) { ss.flag[c] = 0u; mr.loc_base[c] = 0u; }
            continue;
        }
        const uint32_t *own = mr.hist_l + (size_t)c * nb1;
        // (peer memory is not cached by this GPU's L2, .cg skips its L1: plain loads the compiler may batch - all
        //  ranks' rows of a chunk are in flight together, one NVLink round trip per chunk instead of one per rank)
        for (int i4 = tid; i4 < nb1 / 4; i4 += kThreads) {
            uint4 b[kMaxPeers];
#pragma unroll
            for (int r = 0; r < kMaxPeers; ++r) {
                b[r] = make_uint4(0u, 0u, 0u, 0u);
                if (r < px.n) {
                    const uint32_t *row = (r == px.self) ? own : px.arena[r] + px.offHist + (size_t)c * nb1;
                    b[r] = __ldcg(reinterpret_cast<const uint4 *>(row) + i4);
                }
            }
            uint4 a = b[0];
#pragma unroll
            for (int r = 1; r < kMaxPeers; ++r) { a.x += b[r].x; a.y += b[r].y; a.z += b[r].z; a.w += b[r].w; }
            reinterpret_cast<uint4 *>(hbuf)[i4] = a;
        }
        __syncthreads();
        uint32_t bf, bl;
        sel_resolve_cell(lv, ss, c, nb1, candCap, true, hbuf, rs, bf, bl, true);
        uint32_t s = 0;
        if (bf <= bl) for (uint32_t i = tid; i < bf; i += kThreads) s += __ldcg(own + i);
        s = __reduce_add_sync(0xffffffffu, s);
        if (lane == 0) s_red[warp] = s;
        __syncthreads();
        if (tid == 0) {
            uint32_t t = 0;
#pragma unroll
            for (int w = 0; w < kWarps; ++w) t += s_red[w];
            mr.loc_base[c] = t;
        }
    }
}

// NCCL transport: after COMPACT (which resolved per block from the all-reduced rows) publish every cell's resolve,
// also for cells without local particles, and put the fill level into the slot's count word
__global__ void __launch_bounds__(kThreads) k_selmr_prep(LevelState lv, SelState ss /* hist = global rows */, SelMrState mr,
                                                          uint32_t nCells, int nb1, uint32_t candCap) {
    extern __shared__ __align__(16) unsigned char sel_smem[];
    uint32_t *hbuf = reinterpret_cast<uint32_t *>(sel_smem);
    __shared__ SelResolveSmem rs;
    __shared__ uint32_t s_red[kWarps];
    pdl_enter();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (uint32_t c = blockIdx.x; c < nCells; c += gridDim.x) {
        uint32_t *cntWord = reinterpret_cast<uint32_t *>(mr.slots_l + (size_t)c * mr.slotWords + (mr.slotWords - 1u));
        if (!lv.active[c]) {          // block-uniform
            if (tid == 0) { ss.flag[c] = 0u; *cntWord = 0u; mr.loc_base[c] = 0u; }
            continue;
        }
        uint32_t bf, bl;
        sel_resolve_cell(lv, ss, c, nb1, candCap, true, hbuf, rs, bf, bl);
        uint32_t s = 0;
        if (bf <= bl) for (uint32_t i = tid; i < bf; i += kThreads) s += __ldcg(mr.hist_l + (size_t)c * nb1 + i);
        s = __reduce_add_sync(0xffffffffu, s);
        if (lane == 0) s_red[warp] = s;
        __syncthreads();
        if (tid == 0) {
            uint32_t t = 0;
#pragma unroll
            for (int w = 0; w < kWarps; ++w) t += s_red[w];
            mr.loc_base[c] = t;
            *cntWord = __ldcg(&ss.cursor[c]);
        }
        __syncthreads();
    }
}

// block search over the candidates of all ranks.  PEER: pulled from the ranks' slots over NVLink; else from the
// all-gathered copy.  The last block to finish reports 1 + (cells flagged at this level) to the host.
template <bool PEER>
__global__ void __launch_bounds__(1024) k_selmr_finish(LevelState lv, SelState ss, SelCtl sc, SelMrState mr, SelPeers px,
                                                       uint32_t nCells, int nb1, uint32_t cap, int *__restrict__ err, int hbmPasses) {
    extern __shared__ __align__(16) unsigned char sel_smem[];
    float *sbuf = reinterpret_cast<float *>(sel_smem);
    uint32_t *hist2 = reinterpret_cast<uint32_t *>(sbuf + cap + 4);
    float *amb = reinterpret_cast<float *>(hist2 + kSelBins2);
    __shared__ SelSearchSmem sm;
    __shared__ uint32_t s_cnt[kMaxPeers];
    __shared__ uint32_t s_nl;
    pdl_enter();
    if (PEER) selx_barrier(px);
    const int tid = threadIdx.x, lane = tid & 31, nThreads = (int)blockDim.x;
    if (blockIdx.x == 0 && tid == 0) atomicAdd(sc.passes_out, hbmPasses);
    for (uint32_t c = blockIdx.x; c < nCells; c += gridDim.x) {
        __syncthreads();
        const uint32_t act = lv.active[c];
        const uint32_t flg = __ldcg(&ss.flag[c]), K = __ldcg(&ss.ncand[c]), base = __ldcg(&ss.base[c]);
        const uint32_t bf = __ldcg(&ss.bfirst[c]), bl = __ldcg(&ss.blast[c]);
        const float L = lv.mL[c], R = lv.mR[c];
        if (tid == 0) {
            s_nl = 0u;
            if (!PEER) ss.cursor[c] = 0u;      // NCCL: the count word was taken by k_selmr_prep.  PEER: peers still read it
        }
        if (!act || flg) continue;              // flagged by the resolve: already counted
        if (tid < mr.nRanks) {
            if (PEER) s_cnt[tid] = tid == mr.self ? __ldcg(&ss.cursor[c]) : ld_sys_u32(px.arena[tid] + px.offCursor + c);
            else s_cnt[tid] = __float_as_uint(__ldcg(mr.slots_g + ((size_t)tid * nCells + c) * mr.slotWords + (mr.slotWords - 1u)));
        }
        __syncthreads();
        uint32_t sum = 0;
        bool over = false;
        for (int r = 0; r < mr.nRanks; ++r) { over |= s_cnt[r] > mr.slotWords - 1u; sum += s_cnt[r]; }
        if (over) {        // some rank's candidates did not fit its slot: every rank sees it, the iterative search takes the cell
            if (tid == 0) { ss.flag[c] = 1u; atomicAdd(ss.n_flagged, 1u); }
            continue;
        }
        if (sum != K || K > cap) {   // cannot happen: all ranks bin with the same function and resolve the same rows
            if (tid == 0) atomicExch(err, ORB_ERR_STATE);
            continue;
        }
        const float *mine = PEER ? mr.slots_l + (size_t)c * mr.slotWords : mr.slots_g + ((size_t)mr.self * nCells + c) * mr.slotWords;
        // slots are 128-byte aligned: 16-byte loads, four in flight per thread and rank (remote ones cross NVLink)
        uint32_t off = 0;
        for (int r = 0; r < mr.nRanks; ++r) {
            const uint32_t n = s_cnt[r];
            const float *src = PEER ? (r == mr.self ? mine : reinterpret_cast<const float *>(px.arena[r] + px.offSlots) + (size_t)c * mr.slotWords)
                                    : mr.slots_g + ((size_t)r * nCells + c) * mr.slotWords;
            const float4 *s4 = reinterpret_cast<const float4 *>(src);
            const uint32_t n4 = n >> 2;
            uint32_t i = tid;
            for (; i + 3u * nThreads < n4; i += 4u * nThreads) {
                const float4 q0 = __ldcg(s4 + i), q1 = __ldcg(s4 + i + nThreads), q2 = __ldcg(s4 + i + 2u * nThreads), q3 = __ldcg(s4 + i + 3u * nThreads);
                float *d = sbuf + off + 4u * i;
                d[0] = q0.x; d[1] = q0.y; d[2] = q0.z; d[3] = q0.w;
                d += 4u * nThreads; d[0] = q1.x; d[1] = q1.y; d[2] = q1.z; d[3] = q1.w;
                d += 4u * nThreads; d[0] = q2.x; d[1] = q2.y; d[2] = q2.z; d[3] = q2.w;
                d += 4u * nThreads; d[0] = q3.x; d[1] = q3.y; d[2] = q3.z; d[3] = q3.w;
            }
            for (; i < n4; i += nThreads) {
                const float4 q = __ldcg(s4 + i);
                float *d = sbuf + off + 4u * i;
                d[0] = q.x; d[1] = q.y; d[2] = q.z; d[3] = q.w;
            }
            if (4u * n4 + tid < n) sbuf[off + 4u * n4 + tid] = __ldcg(src + 4u * n4 + tid);
            off += n;
        }
        __syncthreads();
        const bool done = sel_block_search(sbuf, K, base, 1, L, sel_scale(L, R, nb1), nb1, (int)bf, (int)bl, hist2, amb, lv, ss, sc, c, hbmPasses, sm);
        __syncthreads();
        if (!done) continue;
        // local left count at the final cut: getCut() of the final margins is the found cut as well as the capped cell's cut
        const float cutf = mid_cut(lv.mL[c], lv.mR[c]);
        uint32_t n = 0;
        for (uint32_t i = tid; i < s_cnt[mr.self]; i += nThreads) n += (__ldcg(mine + i) < cutf) ? 1u : 0u;
        n = __reduce_add_sync(0xffffffffu, n);
        if (lane == 0 && n) atomicAdd(&s_nl, n);
        __syncthreads();
        if (tid == 0) lv.nleft_l[c] = mr.loc_base[c] + s_nl;
    }
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        if (atomicAdd(mr.done, 1u) == gridDim.x - 1u) {
            __threadfence();
            *mr.h_status = *((volatile uint32_t *)ss.n_flagged) + 1u;
        }
    }
}

}  // namespace orb
