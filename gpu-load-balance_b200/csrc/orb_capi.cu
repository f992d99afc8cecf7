// orb_capi.cu — context, per-level orchestration and the extern "C" surface of liborb_b200.so
// (declared in include/orb_b200.h).  The level loop of orbit.cpp:102-275 runs here with no host
// round trip per bisection iteration: count -> (NCCL allreduce) -> device-side decision, with the
// host only polling a mapped status word a bounded number of passes behind the GPU.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <unistd.h>
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <memory>
#include <vector>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include "orb_kernels.cuh"
#include "orb_select.cuh"
#include "orb_exchange.cuh"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CK(call)                                                                                           \
    do {                                                                                                   \
        cudaError_t rc_ = (call);                                                                          \
        if (rc_ != cudaSuccess)                                                                            \
            return fail(ORB_ERR_CUDA, "%s error %d in %s(%d)\n%s", #call, (int)rc_, __FILE__, __LINE__,    \
                        cudaGetErrorString(rc_));                                                          \
    } while (0)

// ---- NCCL, resolved at run time so a single-GPU user needs no NCCL at all ----
struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool load() {
        if (lib) return true;
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (lib) break;
        }
        if (!lib) return false;
        GetUniqueId = (decltype(GetUniqueId))dlsym(lib, "ncclGetUniqueId");
        CommInitRank = (decltype(CommInitRank))dlsym(lib, "ncclCommInitRank");
        CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
        AllReduce = (decltype(AllReduce))dlsym(lib, "ncclAllReduce");
        AllGather = (decltype(AllGather))dlsym(lib, "ncclAllGather");
        GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
        return GetUniqueId && CommInitRank && CommDestroy && AllReduce && AllGather && GetErrorString;
    }
} g_nccl;

#define NK(call)                                                                                     \
    do {                                                                                             \
        ncclResult_t rc_ = (call);                                                                   \
        if (rc_ != ncclSuccess)                                                                      \
            return fail(ORB_ERR_NCCL, "%s failed: %s", #call, g_nccl.GetErrorString(rc_));          \
    } while (0)

constexpr int kMaxLevels = 40;
constexpr uint32_t kSelCellCapMax = 40960;   // candidates k_sel_percell keeps in shared memory at most (176 KB with its histogram)
constexpr uint32_t kParMaxCells = 256;    // cells of a level the parallel FINISH holds fine histograms / lists for (streaming levels have fewer)
constexpr uint32_t kVisitRecs = 4096;     // COMPACT blocks that may leave visit records (>= resident blocks of any supported device)
constexpr uint32_t kSelValsCap = 49152;   // values one block of the selection search stages in shared memory (192 KB)
constexpr uint32_t kSelMrMaxCells = 2048; // several ranks: levels up to this many cells use the selection search
constexpr size_t kSelSlotWordsTotal = (size_t)1 << 20;   // candidate slots of one rank and level (4 MB; v2 arena: at least this)
constexpr uint32_t kSelSlotWordsMax = 65536;             // largest slot of one rank and cell
constexpr size_t kXScratchWords = (size_t)1 << 22;       // owner's scratch for cells whose candidates exceed shared memory (16 MB)
constexpr int kDbgPasses = 12;      // ORB_DEBUG_TIMES=2: passes and blocks recorded per level
constexpr uint32_t kDbgBlocks = 1024;
constexpr int kPassSlots = 40;   // >= 32 passes + slack, per level

}  // namespace

struct orb_ctx {
    int device = 0;
    int nSM = 148;
    uint64_t nLocal = 0;
    uint32_t d = 0;           // leaf cells
    uint32_t nHeap = 0;       // 2d-1
    uint32_t maxLevelCells = 0;
    cudaStream_t stream = nullptr;

    float *x[2] = {nullptr, nullptr}, *y[2] = {nullptr, nullptr}, *z[2] = {nullptr, nullptr};
    int cur = 0;
    bool haveParticles = false;

    orb_cell *d_heap = nullptr;       // [nHeap]
    orb_cell *d_cells = nullptr;      // [maxLevelCells] staging for service-granular calls
    uint32_t *d_range = nullptr;      // [nHeap*2]
    uint32_t *d_total = nullptr;      // [nHeap]
    orb::LevelState lv{};
    orb::LevelState lvAlt{};          // second set of the per-cell arrays: k_split prepares the next level in it while the
    uint32_t *d_tile_first_alt = nullptr;   // partition still reads this level's (orb_build swaps the two per level)
    bool fuseNextLevel = true;        // ORB_FUSE_NEXT=0: separate k_level_setup / k_tile_map launches per level
    int prefuseHist = -1;             // ORB_PREFUSE: 1 the partition builds the next level's histogram rows, 0 never, unset: where it pays
    uint32_t *d_cnt_g_buf = nullptr;  // separate allreduce target (multi-rank only)
    float *d_final_cut = nullptr;     // [maxLevelCells]
    uint32_t *d_tile_first = nullptr; // [nMapTiles]
    uint32_t *d_blk_left = nullptr, *d_blk_restart = nullptr;   // [resident blocks] partition phase-1 records
    uint32_t *d_tickets = nullptr;    // [kMaxLevels]
    uint32_t *d_nactive = nullptr;    // [kMaxLevels*kPassSlots]
    uint32_t *d_done = nullptr;       // [kMaxLevels*kPassSlots]
    uint32_t *d_misc = nullptr;       // [0]=n_unfound
    unsigned long long *d_active_particles = nullptr;
    int32_t *d_level_iters = nullptr; // [kMaxLevels]
    int *d_err = nullptr;
    uint32_t *d_bb = nullptr;         // [maxLevelCells*2*8] encoded boxes
    float *d_bb6 = nullptr;           // [maxLevelCells*2*6]

    volatile uint32_t *h_status = nullptr;   // pinned+mapped [kMaxLevels*kPassSlots]
    uint32_t *h_status_dev = nullptr;        // device alias of h_status
    uint32_t *h_scratch = nullptr;           // pinned scratch for small read-backs
    size_t h_scratch_bytes = 0;

    int occPartStream = 1, occPartCells = 1;   // resident blocks per SM of the partition kernels
    int trialDepth = 3;
    int runAhead = 1;
    bool fuseUpdate = true;
    int tieMode = 0;               // 0 canonical (stable x<cut), 1 Hoare-exact (partition.cpp:30-60)
    int occHoare = 1;
    uint32_t *d_blk_le = nullptr, *d_nGE = nullptr, *d_nLE = nullptr;
    int streamMinTiles = 16;       // average cell size (in 16 KB tiles) from which the tile-streaming count kernel is used
    bool compaction = true;
    bool persist = true;           // host-free level loop (k_level_persistent) where it applies
    int occPersist[4] = {0, 0, 0, 0}; // resident blocks per SM of k_level_persistent<M>
    int32_t *d_lvl_passes = nullptr;  // [kMaxLevels]
    unsigned long long *d_dbg = nullptr; // ORB_DEBUG_TIMES: [kMaxLevels][64] globaltimer stamps
    unsigned long long *d_dbg_blocks = nullptr;   // ORB_DEBUG_TIMES=2: [kMaxLevels][kDbgPasses][kDbgBlocks][4] per-block stamps
    uint32_t dbgGrid[kMaxLevels] = {};
    uint32_t *d_lvl_unfound = nullptr; // [kMaxLevels]
    // selection-based cut search (orb_select.cuh), default trial depth; several ranks: see selectMr below
    bool select = true;
    int selPerCellMinCells = 64;   // levels with at least this many cells: one block searches a whole cell
    bool selBigBlocks = true;      // ORB_SELECT_BIG_BLOCKS=0: no 1024 x 1 / 512 x 2 variants of k_sel_percell
    bool selBigFinish = false;     // ORB_SELECT_BIG_FINISH=1 (experimental, unmeasured): streaming levels gather any number of
                                   // candidates and the finish kernel searches an over-full list in global memory
    int selBinAvg = 8192;          // ORB_SELECT_BIN_AVG: HIST bins per cell are doubled (512..8192) until a bin holds at most this many particles on average
    int selT512MinAvg = 32768;     // ORB_SELECT_T512_MIN: cells of at least this many particles get 512-thread blocks
    // Sampled rows (single rank, orb_select.cuh SelSampleEst): HIST bins every sampleS-th tile (piece) only, RESOLVE widens the
    // candidate bins by sampleZ standard deviations, the gathering pass proves the bracket with exact counts.
    int sampleS = 8;               // ORB_SAMPLE_STRIDE (1: exact rows everywhere)
    float sampleZ = 5.f;           // ORB_SAMPLE_Z
    // Where it pays (measured, profiles/r02k_*): the HIST pass must be bandwidth-bound for a sample to save anything - from
    // 2^25 particles per GPU, like the partition-built rows it replaces - and the candidates of a cell must not swamp the
    // one block that finishes it: streaming cells of at most 2^25 particles (the margin of a 2^27-particle cell is
    // 160 000 candidates), block-searched cells of at least 2^16 (kSelSampleMinCell).
    uint64_t sampleMinLocal = 1ull << 25;   // ORB_SAMPLE_MIN_LOCAL
    uint64_t sampleMaxAvg = 1ull << 25;     // ORB_SAMPLE_MAX_AVG
    bool sampleMaxAvgDefault = true;        // (not set by the environment)
    uint32_t sampleMinCell = orb::kSelSampleMinCell;   // ORB_SAMPLE_MIN_CELL: block-searched cells of fewer particles use exact rows
    // parallel FINISH of the sampled streaming levels (orb_select.cuh SelPar): fine histogram filled by COMPACT, then
    // k_sel_fin_a / k_sel_gather / k_sel_fin_b instead of one block walking all candidates of its cell twice
    bool parFinish = true;         // ORB_PAR_FINISH=0: k_sel_finish<true>
    orb::SelPar par{};
    int chunkOcc = 2;              // ORB_CHUNK_OCC: blocks per SM of the chunking the search's last pass and the cooperative partition share
                                   // (measured: the partition streams 10 % faster with 2 x 148 chunks than with 3 x 148)
    bool selLowOcc = false;        // ORB_SELECT_LOW_OCC=1: k_sel_percell with 64 registers per thread (4 x 256 / 2 x 512 threads per SM, no spills)
    bool levelSampled = false;     // the level being searched uses sampled rows
    bool sampleOff = false;        // this build: a level's brackets failed (particle order correlates with position) - exact rows from there on
    orb::SelVisitRec *d_visits = nullptr;   // [kVisitRecs] COMPACT with private candidate regions: one record per block
    bool pdl = true;               // programmatic dependent launch between the small kernels of a level
    // partition without its phase-1 read (PreLeft, orb_kernels.cuh): the search's last pass and the partition share one
    // chunk per block; ORB_PRELEFT=0 disables
    int partBulk = 0;              // ORB_PART_BULK=1: the cooperative partition loads its tiles with cp.async.bulk + mbarrier instead of
                                   // per-thread cp.async (measured 2.8 % slower at C3 with two tiles of shared memory: profiles/r02h_partition_bulk_ab.txt)
    int partWarpMax = 2048;        // ORB_PART_WARP_MAX: average local cell size up to which the partition runs one warp per cell
    bool preLeft = true;
    orb::PreLeft *d_pre = nullptr; // [64 * nSM]
    uint32_t preTagSeq = 0;        // tag of the current level's records
    uint32_t chunkTiles = 0;       // count tiles (4096 particles) per block of the current level's shared chunking; 0: none
    uint32_t preListStride = 0;    // several ranks: words of one cell's candidate slot (the records index into d_slots_l)
    bool preValid = false;         // this level's search wrote records the partition may use
    orb::SelState sel{};
    size_t selHistWords = 0;
    uint32_t *d_sel_nflag = nullptr;   // [kMaxLevels] cells left to the iterative path per level
    int occSelStream[3] = {1, 1, 1};   // HIST, COMPACT, COMPACT without a row buffer (cells resolved by k_sel_resolve)
    bool profile = false;
    struct LabelledEvent { const char *label; int level; cudaEvent_t e0, e1; };
    std::vector<LabelledEvent> evAux;      // profile mode: per-kernel times of the selection search
    size_t evAuxUsed = 0;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> evCount, evPart;
    size_t evCountUsed = 0, evPartUsed = 0;

    // multi-GPU
    ncclComm_t comm = nullptr;
    bool ownComm = false;
    int rank = 0, nRanks = 1;
    // selection search over several ranks (k_selmr_* in orb_select.cuh): global histogram rows, candidate slots
    bool selectMr = true;              // ORB_SELECT_MR=0: iterative search on every multi-rank level
    uint64_t nLocalMin = 0, nGlobal = 0;   // over ranks; every decision that shapes a collective uses these only
    uint32_t *d_sel_hist_g = nullptr;  // [selHistWords] NCCL transport: rows summed over ranks
    uint32_t *d_sel_locbase = nullptr; // [maxLevelCells]
    float *d_slots_l = nullptr;        // [kSelSlotWordsTotal] (inside the exchange arena)
    float *d_slots_g = nullptr;        // [nRanks][kSelSlotWordsTotal] NCCL transport: all-gathered slots
    // exchange arena mapped by every rank (peer transport): flags[64] | cursor[L] | slots | histogram rows
    uint32_t *d_xchg = nullptr;
    uint32_t xOffCursor = 0, xOffSlots = 0, xOffHist = 0;   // word offsets, identical on all ranks
    uint32_t *peerX[orb::kMaxPeers] = {nullptr};
    uint32_t xSeq = 0;                 // exchanges issued so far (same on all ranks)
    // protocol v2 (orb_exchange.cuh): owner-computed search over peer memory for levels of any size.  Arena regions
    // beyond the v1 ones: result records | candidate counts | received candidate slots | global rows
    bool mrV2 = true;                  // ORB_MR_V1=1: the two-exchange kernels of orb_select.cuh (levels of <= 2048 cells)
    bool mrSelf = false;               // ORB_MR_SELF=1 (testing): one rank runs the multi-rank protocol against itself
    uint32_t xOffRes = 0, xOffRecvCnt = 0, xOffRecv = 0, xOffHistG = 0;
    size_t slotTotal = 0;              // words of candidate slots per rank and level (v2 arena)
    uint64_t nLocalMax = 0;
    float *d_xscratch = nullptr;       // [kXScratchWords]
    uint32_t xCandCap = kSelValsCap;   // ORB_X_CAND_CAP (testing): candidates an owner block stages in shared memory
    std::vector<int> extraPasses;      // passes of the iterative fallback per level (host-driven on several ranks)

    // fused combine+update over NVLink peer memory (optional; see PeerSet in orb_kernels.cuh)
    bool peerEnabled = false;
    uint32_t peerSeq = 0;
    uint32_t *d_peer_cnt = nullptr;    // receive rows [2][kMaxPeers][kPeerMaxCells][kCS]
    uint32_t *d_peer_flag = nullptr;   // flags [2][kMaxPeers][kPeerMaxBlocks]
    uint32_t *d_cdone = nullptr;       // [kMaxLevels*kPassSlots] per-pass "blocks finished" counters
    uint32_t *peerCnt[orb::kMaxPeers] = {nullptr};
    uint32_t *peerFlag[orb::kMaxPeers] = {nullptr};
    bool peerIpc[orb::kMaxPeers] = {false};

    // launch accounting
    uint64_t nCountLaunch = 0, nUpdateLaunch = 0, nPartLaunch = 0, nOtherLaunch = 0;
};

namespace {

inline uint32_t ceil_div(uint64_t a, uint32_t b) { return (uint32_t)((a + b - 1) / b); }

// Launch with programmatic dependent launch allowed (see pdl_enter in orb_kernels.cuh): the kernel may begin while
// the previous kernel of the stream drains.  ORB_PDL=0 launches it as an ordinary kernel.
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(const orb_ctx *c, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = c->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = c->pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

int ensure_scratch(orb_ctx *c, size_t bytes) {
    if (c->h_scratch_bytes >= bytes) return ORB_OK;
    if (c->h_scratch) cudaFreeHost(c->h_scratch);
    c->h_scratch = nullptr;
    c->h_scratch_bytes = 0;
    CK(cudaMallocHost((void **)&c->h_scratch, bytes));
    c->h_scratch_bytes = bytes;
    return ORB_OK;
}

int check_device_err(orb_ctx *c) {
    int e = 0;
    CK(cudaMemcpyAsync(&e, c->d_err, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (e != 0) {
        CK(cudaMemsetAsync(c->d_err, 0, sizeof(int), c->stream));
        if (e == ORB_ERR_RANGE) return fail(ORB_ERR_RANGE, "cells of the level do not tile [0,n_local) in id order");
        return fail(e, "invalid cell (cutAxis outside 0..2)");
    }
    return ORB_OK;
}

// ---- level preparation: SoA state + tile map (ServiceCopyCells' role, copyCells.cu:29-61) ----
int level_prepare(orb_ctx *c, const orb_cell *d_cells, uint32_t nCells, int nc, uint32_t *n_active0, uint32_t *zero = nullptr,
                  size_t nZero = 0) {
    using namespace orb;
    if (nCells == 0 || nCells > c->maxLevelCells) return fail(ORB_ERR_ARG, "n_cells %u out of range (max %u)", nCells, c->maxLevelCells);
    const uint32_t blocks = ceil_div(nCells, 256);
    CK(launch_pdl(c, k_level_setup, dim3(blocks), dim3(256), 0, d_cells, nCells, (const uint32_t *)c->d_range, (const uint32_t *)c->d_total, c->lv,
                  (uint32_t)c->nLocal, nc, c->d_err, n_active0));
    c->nOtherLaunch++;
    const uint32_t nMap = ceil_div(c->nLocal, kMapTile);
    if (nMap) {
        CK(launch_pdl(c, k_tile_map, dim3(ceil_div(nMap, 256)), dim3(256), 0, (const uint32_t *)c->lv.bnd, nCells, nMap, c->d_tile_first,
                      zero, nZero));
        c->nOtherLaunch++;
    }
    CK(cudaGetLastError());
    return ORB_OK;
}

// cells of at least 16 tiles on average: the tile-streaming kernel (and the byte-reducing search) apply
inline bool count_streams(const orb_ctx *c, uint32_t nCells) { return c->nLocal / nCells >= (uint64_t)c->streamMinTiles * orb::kCountTile; }

// Count pass over all active cells of the level.  Kernel choice by average local cell size:
//   >= 16 tiles   : k_count_stream (persistent, tile streaming, one atomic per block per (cell,cut))
//   >= 1024       : k_count_cells, one block per cell
//   otherwise     : k_count_cells, one warp per cell
template <int NC>
int launch_count_nc(orb_ctx *c, uint32_t nCells, const uint32_t *gate, const orb::FuseCtl &fc, int mode) {
    using namespace orb;
    const float *x = c->x[c->cur], *y = c->y[c->cur], *z = c->z[c->cur];
    float *cand = c->x[c->cur ^ 1];   // the idle ping-pong column holds the candidates of the byte-reducing search
    if (count_streams(c, nCells)) {
        const uint32_t nTiles = ceil_div(c->nLocal, kCountTile);
        const uint32_t grid = std::min<uint32_t>(nTiles, (uint32_t)c->nSM * 4u);
        const size_t ringBytes = (size_t)kCountStages * kCountTile * sizeof(float);
        if (NC == 7 && mode == kCountCompact)
            k_count_stream<7, kCountCompact><<<grid, kThreads, ringBytes, c->stream>>>(x, y, z, cand, c->lv, c->d_tile_first, nCells, (uint32_t)c->nLocal, nTiles, gate, fc);
        else if (NC == 7 && mode == kCountCand)
            k_count_stream<7, kCountCand><<<grid, kThreads, ringBytes, c->stream>>>(x, y, z, cand, c->lv, c->d_tile_first, nCells, (uint32_t)c->nLocal, nTiles, gate, fc);
        else
            k_count_stream<NC, kCountFull><<<grid, kThreads, ringBytes, c->stream>>>(x, y, z, cand, c->lv, c->d_tile_first, nCells, (uint32_t)c->nLocal, nTiles, gate, fc);
    } else if (c->nLocal / nCells >= 1024) {
        const uint32_t grid = std::min<uint32_t>(nCells, (uint32_t)c->nSM * 4u);
        k_count_cells<NC, 256><<<grid, kThreads, 0, c->stream>>>(x, y, z, c->lv, nCells, gate, fc);
    } else {
        const uint32_t grid = std::min<uint32_t>(ceil_div(nCells, kWarps), (uint32_t)c->nSM * 4u);
        k_count_cells<NC, 32><<<grid, kThreads, 0, c->stream>>>(x, y, z, c->lv, nCells, gate, fc);
    }
    return ORB_OK;
}

orb::PeerSet no_peers() {
    orb::PeerSet ps;
    memset(&ps, 0, sizeof(ps));
    return ps;
}

orb::XArena no_arena() {
    orb::XArena xa;
    memset(&xa, 0, sizeof(xa));
    return xa;
}

orb::SelPriv no_priv() {
    orb::SelPriv sp;
    memset(&sp, 0, sizeof(sp));
    return sp;
}
orb::FuseCtl no_fuse() {
    orb::FuseCtl fc;
    memset(&fc, 0, sizeof(fc));
    return fc;
}

int launch_count(orb_ctx *c, uint32_t nCells, int nc, const uint32_t *gate, const orb::FuseCtl &fc = no_fuse(), int mode = 0) {
    using namespace orb;
    if (!c->nLocal && !fc.enabled) return ORB_OK;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (c->profile) {
        if (c->evCountUsed == c->evCount.size()) {
            cudaEvent_t a, b;
            CK(cudaEventCreate(&a));
            CK(cudaEventCreate(&b));
            c->evCount.emplace_back(a, b);
        }
        e0 = c->evCount[c->evCountUsed].first;
        e1 = c->evCount[c->evCountUsed].second;
        c->evCountUsed++;
        CK(cudaEventRecord(e0, c->stream));
    }
    switch (nc) {
    case 1: launch_count_nc<1>(c, nCells, gate, fc, mode); break;
    case 3: launch_count_nc<3>(c, nCells, gate, fc, mode); break;
    case 7: launch_count_nc<7>(c, nCells, gate, fc, mode); break;
    default: return fail(ORB_ERR_ARG, "unsupported trial count %d", nc);
    }
    if (c->profile) CK(cudaEventRecord(e1, c->stream));
    c->nCountLaunch++;
    CK(cudaGetLastError());
    return ORB_OK;
}

// sum the per-cell counters over ranks (Combine: countLeft.cpp:44-53), in-stream
int allreduce_counts(orb_ctx *c, uint32_t nCells, int nc) {
    if (c->nRanks <= 1) return ORB_OK;
    const size_t n = (size_t)nCells * orb::kCS;
    (void)nc;
    NK(g_nccl.AllReduce(c->lv.cnt_l, c->lv.cnt_g, n, ncclUint32, ncclSum, c->comm, c->stream));
    return ORB_OK;
}

int launch_update(orb_ctx *c, uint32_t nCells, int M, int passSlot, orb::PassCtl ctl, const orb::PeerSet &ps, int baseMode) {
    using namespace orb;
    const uint32_t blocks = ceil_div(nCells, kThreads);
    switch (M) {
    case 1: k_update<1><<<blocks, kThreads, 0, c->stream>>>(c->lv, nCells, passSlot, ctl, ps, baseMode); break;
    case 2: k_update<2><<<blocks, kThreads, 0, c->stream>>>(c->lv, nCells, passSlot, ctl, ps, baseMode); break;
    case 3: k_update<3><<<blocks, kThreads, 0, c->stream>>>(c->lv, nCells, passSlot, ctl, ps, baseMode); break;
    default: return fail(ORB_ERR_ARG, "unsupported trial depth %d", M);
    }
    c->nUpdateLaunch++;
    CK(cudaGetLastError());
    return ORB_OK;
}

inline void cpu_relax() {
#if defined(__x86_64__)
    _mm_pause();
#endif
}

// Wait for a mapped status word the GPU writes (no stream synchronisation in the common case).  Every few thousand
// polls the stream is queried, so that a kernel fault surfaces as ORB_ERR_CUDA instead of a hang; a rank that waits
// longer than ORB_WAIT_TIMEOUT_S (default 120 s: a peer died or diverged) gives up with ORB_ERR_STATE.
int wait_status(orb_ctx *c, volatile uint32_t *w, uint32_t *out) {
    static const double limit = [] { const char *e = getenv("ORB_WAIT_TIMEOUT_S"); return e && atof(e) > 0 ? atof(e) : 120.0; }();
    uint32_t s;
    uint64_t polls = 0;
    timespec t0{};
    while ((s = *w) == 0u) {
        cpu_relax();
        if ((++polls & 0x3fffu) == 0) {
            const cudaError_t q = cudaStreamQuery(c->stream);
            if (q != cudaSuccess && q != cudaErrorNotReady)
                return fail(ORB_ERR_CUDA, "stream failed while waiting for the device: %s", cudaGetErrorString(q));
            timespec t1;
            clock_gettime(CLOCK_MONOTONIC, &t1);
            if (t0.tv_sec == 0 && t0.tv_nsec == 0) t0 = t1;
            if ((double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec) > limit)
                return fail(ORB_ERR_STATE, "no answer from the device after %.0f s (a peer rank stopped or took another branch?)", limit);
            if (q == cudaSuccess && (s = *w) == 0u)    // everything enqueued has run and the word is still unset
                return fail(ORB_ERR_STATE, "the stream drained without the status word being written");
        }
    }
    *out = s;
    return ORB_OK;
}

// Bisection loop of one level (orbit.cpp:146-232).  `slotBase` selects this level's private region of
// the pass-control arrays (zeroed when the build / call starts).  Returns passes launched.
int run_bisection(orb_ctx *c, uint32_t nCells, int M, int slotBase, int levelIdx, int *passesOut) {
    using namespace orb;
    const int nc = (1 << M) - 1;
    const int maxPasses = (kMaxIter + M - 1) / M;
    PassCtl ctl;
    ctl.n_active = c->d_nactive + slotBase;
    ctl.done = c->d_done + slotBase;
    ctl.h_status = (volatile uint32_t *)(c->h_status_dev + slotBase);
    ctl.active_particles = c->d_active_particles;
    ctl.level_iters = c->d_level_iters + levelIdx;
    volatile uint32_t *hs = c->h_status + slotBase;
    // multi-rank combine: fused into the update kernel over peer memory when the level is small enough,
    // otherwise an in-stream ncclAllReduce between count and update
    const bool fused = c->nRanks > 1 && c->peerEnabled && nCells <= kPeerMaxCells;
    int launched = 0;
    for (;;) {
        PeerSet ps = no_peers();
        if (fused) {
            ps.n = c->nRanks;
            ps.self = c->rank;
            ps.seq = ++c->peerSeq;
            for (int r = 0; r < c->nRanks; ++r) { ps.recv[r] = c->peerCnt[r]; ps.flag[r] = c->peerFlag[r]; }
        }
        // byte-reducing search: pass 0 reads everything and fixes the bracket, pass 1 compacts, later passes read candidates
        const bool compacting = c->compaction && M == 3 && c->nLocal > 0 && count_streams(c, nCells);
        const int mode = !compacting ? kCountFull : (launched == 0 ? kCountFull : (launched == 1 ? kCountCompact : kCountCand));
        const int baseMode = !compacting ? 0 : (launched == 0 ? 1 : 2);
        // single rank, small level: the count kernel's last block runs the update itself (one launch per pass)
        FuseCtl fc = no_fuse();
        if (c->nRanks == 1 && nCells <= kFuseMaxCells && c->nLocal > 0 && c->fuseUpdate) {
            fc.enabled = 1; fc.M = M; fc.pass = launched; fc.baseMode = baseMode; fc.tickets = c->d_cdone + slotBase; fc.ctl = ctl;
        }
        int rc = launch_count(c, nCells, nc, ctl.n_active + launched, fc, mode);
        if (rc) return rc;
        if (!fc.enabled) {
            if (!fused) {
                rc = allreduce_counts(c, nCells, nc);
                if (rc) return rc;
            }
            rc = launch_update(c, nCells, M, launched, ctl, ps, baseMode);
            if (rc) return rc;
        }
        launched++;
        if (launched >= maxPasses) break;
        // Decide deterministically (identically on every rank): look at the status of pass
        // launched-1-runAhead; stop once a pass reported "no active cells".
        const int k = launched - 1 - c->runAhead;
        if (k >= 0) {
            uint32_t s = 0;
            rc = wait_status(c, hs + k, &s);
            if (rc) return rc;
            if (s == 1u) break;
        }
    }
    if (passesOut) *passesOut = launched;
    return ORB_OK;
}

// Whole bisection loop of a level in ONE cooperative launch (k_level_persistent): single rank, streaming regime.
bool level_can_persist(const orb_ctx *c, uint32_t nCells, int M) {
    return c->persist && c->nRanks == 1 && c->nLocal > 0 && nCells <= orb::kPersistMaxCells && count_streams(c, nCells) &&
           M >= 1 && M <= 3 && c->occPersist[M] >= 1;
}

int launch_level_persistent(orb_ctx *c, uint32_t nCells, int M, int slotBase, int levelIdx, const uint32_t *gate = nullptr,
                            const uint32_t *only = nullptr) {
    using namespace orb;
    const uint32_t nTiles = ceil_div(c->nLocal, kCountTile);
    const uint32_t grid = std::min<uint32_t>(nTiles, (uint32_t)c->nSM * (uint32_t)std::min(c->occPersist[M], 4));
    const float *x = c->x[c->cur], *y = c->y[c->cur], *z = c->z[c->cur];
    float *cand = c->x[c->cur ^ 1];
    uint32_t nC = nCells, nL = (uint32_t)c->nLocal, nT = nTiles;
    LevelCtl lc;
    lc.n_active = c->d_nactive + slotBase;
    lc.active_particles = c->d_active_particles;
    lc.level_iters = c->d_level_iters + levelIdx;
    lc.passes_out = c->d_lvl_passes + levelIdx;
    lc.n_unfound_out = c->d_lvl_unfound + levelIdx;
    lc.compaction = (c->compaction && M == 3) ? 1 : 0;
    lc.barrier = c->d_cdone + slotBase;   // per-level counter, zeroed with the pass-control arrays
    lc.dbg = (c->d_dbg && !gate) ? c->d_dbg + (size_t)levelIdx * 64 : nullptr;
    lc.gate = gate;
    lc.only = only;
    lc.dbg_blocks = (c->d_dbg_blocks && grid <= kDbgBlocks && !gate) ? c->d_dbg_blocks + (size_t)levelIdx * kDbgPasses * kDbgBlocks * 4 : nullptr;
    if (lc.dbg_blocks) c->dbgGrid[levelIdx] = grid;
    void *args[] = {(void *)&x, (void *)&y, (void *)&z, (void *)&cand, (void *)&c->lv, (void *)&c->d_tile_first,
                    (void *)&nC, (void *)&nL, (void *)&nT, (void *)&lc};
    const size_t ringBytes = (size_t)kCountStages * kCountTile * sizeof(float);
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (c->profile) {
        if (c->evCountUsed == c->evCount.size()) {
            cudaEvent_t a, b;
            CK(cudaEventCreate(&a));
            CK(cudaEventCreate(&b));
            c->evCount.emplace_back(a, b);
        }
        e0 = c->evCount[c->evCountUsed].first;
        e1 = c->evCount[c->evCountUsed].second;
        c->evCountUsed++;
        CK(cudaEventRecord(e0, c->stream));
    }
    const void *fn = M == 3 ? (const void *)k_level_persistent<3> : (M == 2 ? (const void *)k_level_persistent<2> : (const void *)k_level_persistent<1>);
    CK(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(kThreads), args, ringBytes, c->stream));
    if (c->profile) CK(cudaEventRecord(e1, c->stream));
    c->nCountLaunch++;
    CK(cudaGetLastError());
    return ORB_OK;
}

// profile mode: time one kernel of the selection search under a label
int aux_begin(orb_ctx *c, const char *label, int level) {
    if (!c->profile) return ORB_OK;
    if (c->evAuxUsed == c->evAux.size()) {
        orb_ctx::LabelledEvent ev{label, level, nullptr, nullptr};
        CK(cudaEventCreate(&ev.e0));
        CK(cudaEventCreate(&ev.e1));
        c->evAux.push_back(ev);
    }
    c->evAux[c->evAuxUsed].label = label;
    c->evAux[c->evAuxUsed].level = level;
    CK(cudaEventRecord(c->evAux[c->evAuxUsed].e0, c->stream));
    return ORB_OK;
}
int aux_end(orb_ctx *c) {
    if (!c->profile) return ORB_OK;
    CK(cudaEventRecord(c->evAux[c->evAuxUsed].e1, c->stream));
    c->evAuxUsed++;
    return ORB_OK;
}

// Selection-based cut search of a level (orb_select.cuh): two streaming passes + a per-cell finish where cells are
// large, one read per cell where they fit in shared memory; then the iterative search (gated on the device, no host
// round trip) for the cells it flagged.  This is the single-rank flow (launch_level_select_mr: several ranks).
bool level_can_select(const orb_ctx *c, uint32_t nCells, int M) {
    return c->select && c->nRanks == 1 && c->nLocal > 0 && M == 3 && c->persist && c->occPersist[3] >= 1 && nCells >= 1;
}

// event pair of one particle-streaming kernel (counted in ms_count: these are the HBM-bound kernels of the search)
int count_event_begin(orb_ctx *c) {
    if (!c->profile) return ORB_OK;
    if (c->evCountUsed == c->evCount.size()) {
        cudaEvent_t a, b;
        CK(cudaEventCreate(&a));
        CK(cudaEventCreate(&b));
        c->evCount.emplace_back(a, b);
    }
    CK(cudaEventRecord(c->evCount[c->evCountUsed].first, c->stream));
    return ORB_OK;
}
int count_event_end(orb_ctx *c) {
    if (!c->profile) return ORB_OK;
    CK(cudaEventRecord(c->evCount[c->evCountUsed].second, c->stream));
    c->evCountUsed++;
    return ORB_OK;
}

// Shape of the selection search at a level: cells staged whole in shared memory, or HIST / COMPACT / FINISH with nb1
// bins (rep copies per block) and room for candCap candidates per cell.
struct SelPlan {
    bool cellsInSmem;     // one block per cell (k_sel_percell)
    uint32_t cellCap;     // candidates a block keeps in shared memory
    int threads;          // block size of k_sel_percell
    int variant;          // 0: <512,3> / <256,6>, 1: <512,2,8>, 2: <1024,1,8>
    int nb1, rep;
    size_t histWords;     // nCells * nb1 (cleared by k_tile_map during level preparation)
    uint32_t candCap;
    int sampleS;          // > 1: rows from a sample (streaming: private candidate regions + visit records as well)
};
// single rank, default trial depth: the search may work from sampled rows
// (not with ORB_PREFUSE=1, which asks for the partition-built rows)
inline bool sampling_on(const orb_ctx *c) {
    return c->sampleS > 1 && !c->sampleOff && c->nRanks == 1 && !c->mrSelf && c->d_visits != nullptr && c->prefuseHist != 1 &&
           c->nLocal >= c->sampleMinLocal;
}
// candidates a cell of `avg` particles is expected to have with sampled rows of nb bins: the z-sigma margin of the
// sample's rank estimate on either side plus the two bins that hold the bracket's ends
inline uint64_t sampled_cand_estimate(const orb_ctx *c, uint64_t avg, int nb) {
    return (uint64_t)(c->sampleZ * std::sqrt((double)c->sampleS * (double)avg)) + 2 * (avg / (uint64_t)nb) + 64;
}
SelPlan sel_plan(const orb_ctx *c, uint32_t nCells, int forcedNb = 0 /* rows already built by the partition with this many bins */) {
    SelPlan p{};
    const uint64_t avg = c->nLocal / nCells;
    // one block per cell once there are enough cells to fill the GPU; fewer, larger cells are streamed by all blocks
    p.cellsInSmem = nCells >= (uint32_t)c->selPerCellMinCells && avg <= (1u << 20);
    // block shape: few big cells -> a block per SM (or two) with 8 vector loads in flight per thread, so that the
    // blocks that exist can pull the whole HBM bandwidth; otherwise 3 x 512 or 6 x 256 threads per SM
    p.threads = avg >= (uint64_t)c->selT512MinAvg ? 512 : 256;
    p.variant = 0;
    if (c->selBigBlocks && nCells <= (uint32_t)c->nSM) { p.variant = 2; p.threads = 1024; }
    else if (c->selBigBlocks && nCells <= 2u * (uint32_t)c->nSM) { p.variant = 1; p.threads = 512; }
    p.cellCap = avg >= 65536 ? 8192u : 4096u;
    p.nb1 = orb::kSelBinsMin;
    while (p.nb1 < orb::kSelBinsMax && avg / (uint64_t)p.nb1 > (uint64_t)c->selBinAvg) p.nb1 <<= 1;
    if (forcedNb) p.nb1 = forcedNb;
    // (cells beyond sampleMaxAvg: their candidates would swamp a one-block FINISH; the parallel FINISH has no such limit)
    p.sampleS = (sampling_on(c) && !forcedNb && (p.cellsInSmem || avg <= c->sampleMaxAvg || (c->parFinish && c->sampleMaxAvgDefault && nCells <= kParMaxCells)))
                    ? c->sampleS : 1;
    if (p.sampleS > 1 && !p.cellsInSmem) {
        // sampled rows: the margin dominates the candidates, finer bins cost nothing (few atomics) - bins of <= 1024 particles
        while (p.nb1 < orb::kSelBinsMax && avg / (uint64_t)p.nb1 > 1024) p.nb1 <<= 1;
        if ((size_t)nCells * (size_t)p.nb1 > c->selHistWords) p.sampleS = 1;
    }
    p.rep = p.nb1 <= 512 ? 4 : (p.nb1 <= 1024 ? 2 : 1);
    p.histWords = (p.cellsInSmem && !forcedNb) ? 0 : (size_t)nCells * (size_t)p.nb1;
    // candidates one block will stage per cell: a few bins' worth; cells beyond it go to the iterative search
    p.candCap = (uint32_t)std::min<uint64_t>(kSelValsCap, 4 * (avg / (uint64_t)p.nb1) + 4096);
    if (p.sampleS > 1) {
        if (p.cellsInSmem) {
            const uint64_t want = (sampled_cand_estimate(c, avg, avg > 4096 ? orb::kSelBins2 : 256) * 5 / 4 + 1023) & ~(uint64_t)1023;
            p.cellCap = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(p.cellCap, want), kSelCellCapMax);
        } else {
            // (more than this is searched in global memory by k_sel_finish, not flagged)
            p.candCap = (uint32_t)std::min<uint64_t>(kSelValsCap, (sampled_cand_estimate(c, avg, p.nb1) * 5 / 4 + 4095) & ~(uint64_t)1023);
        }
    }
    return p;
}

// does the partition of a level of nCells cells run as the cooperative reduce-then-scan kernel (see launch_partition)?
inline bool partition_is_coop(const orb_ctx *c, uint32_t nCells) {
    const uint64_t avg = c->nLocal / nCells;
    return !(avg <= 16ull * orb::kPartTile && nCells >= 2u * (uint32_t)c->nSM);
}
// Chunk per block (in count tiles) that the search's last pass and the cooperative partition share at this level, so
// that the search can hand the partition the left counts of its blocks' trailing segments (PreLeft); 0: not at this level.
uint32_t level_chunk_tiles(const orb_ctx *c, uint32_t nCells) {
    if (!c->preLeft || c->tieMode == 1 || !c->nLocal || !partition_is_coop(c, nCells)) return 0u;
    const uint32_t nTiles = ceil_div(c->nLocal, orb::kCountTile);
    // (a sampled level's COMPACT carries no row buffer: k_sel_resolve has resolved the cells)
    const SelPlan pl = sel_plan(c, nCells);
    const int occCompact = (pl.sampleS > 1 && !pl.cellsInSmem) ? c->occSelStream[2] : c->occSelStream[1];
    const uint32_t G = (uint32_t)c->nSM * (uint32_t)std::max(1, std::min(std::min(c->occPartStream, occCompact), c->chunkOcc));
    return std::max<uint32_t>(1u, ceil_div(nTiles, G));
}

int launch_level_select(orb_ctx *c, uint32_t nCells, int slotBase, int levelIdx, int preNb = 0) {
    using namespace orb;
    const float *x = c->x[c->cur], *y = c->y[c->cur], *z = c->z[c->cur];
    float *cand = c->x[c->cur ^ 1];
    SelState ss = c->sel;
    ss.n_flagged = c->d_sel_nflag + levelIdx;
    SelCtl sc;
    sc.active_particles = c->d_active_particles;
    sc.level_iters = c->d_level_iters + levelIdx;
    sc.passes_out = c->d_lvl_passes + levelIdx;
    sc.n_unfound_out = c->d_lvl_unfound + levelIdx;
    int rc;
    const SelPlan pl = sel_plan(c, nCells, preNb);
    SelDone dn;
    dn.done = c->d_cdone + slotBase + kPassSlots - 1;
    dn.h_status = (volatile uint32_t *)(c->h_status_dev + slotBase + kPassSlots - 1);
    c->chunkTiles = level_chunk_tiles(c, nCells);
    c->preValid = c->chunkTiles != 0u;
    PreLeft *pre = c->preValid ? c->d_pre : nullptr;
    const uint32_t preTag = ++c->preTagSeq;
    // block size of the per-cell search kernels: big blocks when shared memory allows one block per SM anyway
    auto search_threads = [](size_t smem) { return smem > 112 * 1024 ? 1024 : (smem > 56 * 1024 ? 512 : 256); };
    c->levelSampled = pl.sampleS > 1;
    if (pl.cellsInSmem) {
        // ---- many cells: one block runs the whole search of a cell ----
        const size_t smem = sel_percell_smem_bytes(pl.cellCap);
        int occ = 1;
        // (the sampled first attempt is a separate instantiation: cells below kSelSampleMinCell and builds without sampling
        //  run the plain one)
        const int percellS = (c->nLocal / nCells >= (uint64_t)c->sampleMinCell) ? pl.sampleS : 1;
        auto kern = percellS > 1
                        ? (pl.variant == 2 ? k_sel_percell<1024, 1, 8, true> : (pl.variant == 1 ? k_sel_percell<512, 2, 8, true>
                           : (pl.threads == 512 ? k_sel_percell<512, 3, 4, true> : k_sel_percell<256, 6, 4, true>)))
                        : (pl.variant == 2 ? k_sel_percell<1024, 1, 8> : (pl.variant == 1 ? k_sel_percell<512, 2, 8>
                           : (pl.threads == 512 ? (c->selLowOcc ? k_sel_percell<512, 2> : k_sel_percell<512, 3>)
                                                : (c->selLowOcc ? k_sel_percell<256, 4> : k_sel_percell<256, 6>))));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, pl.threads, smem));
        const uint32_t grid = std::min<uint32_t>(nCells, (uint32_t)c->nSM * (uint32_t)std::max(occ, 1));
        if ((rc = count_event_begin(c))) return rc;
        CK(launch_pdl(c, kern, dim3(grid), dim3(pl.threads), smem, x, y, z, c->lv, ss, sc, nCells, pl.cellCap, preNb, dn, pre, preTag,
                      c->chunkTiles * (uint32_t)kCountTile, (uint32_t)c->nLocal,
                      percellS, c->sampleZ, c->sampleMinCell));
        if ((rc = count_event_end(c))) return rc;
        c->nCountLaunch++;
    } else {
        // ---- HIST, COMPACT (+ RESOLVE on cell entry), FINISH; the histogram rows were cleared by level_prepare ----
        const int nb1 = pl.nb1, rep = pl.rep;
        const uint32_t candCap = pl.candCap;
        if (pl.histWords > c->selHistWords) return fail(ORB_ERR_STATE, "selection histogram of %zu words exceeds the buffer", pl.histWords);
        unsigned long long *dbgBase = c->d_dbg_blocks ? c->d_dbg_blocks + (size_t)levelIdx * kDbgPasses * kDbgBlocks * 4 : nullptr;
        if (dbgBase) c->dbgGrid[levelIdx] = kDbgBlocks;
        const uint32_t nTiles = ceil_div(c->nLocal, kCountTile);
        const size_t ringBytes = (size_t)kCountStages * kCountTile * sizeof(float);
        const uint32_t nL = (uint32_t)c->nLocal;
        // sampled rows: COMPACT appends to block-private regions and leaves visit records, FINISH gathers (orb_select.cuh SelPriv)
        SelPriv sp;
        memset(&sp, 0, sizeof(sp));
        uint32_t compactTiles = c->chunkTiles;      // count tiles per COMPACT block
        if (pl.sampleS > 1) {
            if (!compactTiles) compactTiles = std::max<uint32_t>(1u, ceil_div(nTiles, (uint32_t)c->nSM * (uint32_t)std::max(1, std::min(c->occSelStream[2], 3))));
            if (ceil_div(nTiles, compactTiles) <= kVisitRecs) {
                sp.sampleS = pl.sampleS;
                sp.z = c->sampleZ;
                sp.visits = c->d_visits;
                sp.chunk = compactTiles * (uint32_t)kCountTile;
                sp.useBounds = 1;
                if (c->parFinish && nCells <= kParMaxCells) { sp.fine = c->par.fine; sp.lo2 = c->par.lo2; sp.sc2 = c->par.sc2; }
            }
        }
        c->levelSampled = sp.sampleS > 1;
        if (!preNb) {      // (preNb: the rows were built by the previous level's partition)
            const size_t smem = ringBytes + (size_t)nb1 * rep * 4;
            int occ = 1;
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_sel_stream<kSelHist>, kThreads, smem));
            const uint32_t nVisit = sp.sampleS > 1 ? ceil_div(nTiles, (uint32_t)sp.sampleS) : nTiles;      // tiles the pass reads
            const uint32_t grid = std::min<uint32_t>(nVisit, (uint32_t)c->nSM * (uint32_t)std::min(std::max(occ, 1), 4));
            if ((rc = count_event_begin(c))) return rc;
            CK(launch_pdl(c, k_sel_stream<kSelHist>, dim3(grid), dim3(kThreads), smem, x, y, z, cand, c->lv, ss, (const uint32_t *)c->d_tile_first,
                          nCells, nL, nTiles, nb1, rep, candCap, dbgBase, (float *)nullptr, 0u, 0, no_arena(), (PreLeft *)nullptr, 0u, 0u, sp));
            if ((rc = count_event_end(c))) return rc;
            c->nCountLaunch++;
        }
        if (sp.useBounds) {     // RESOLVE as its own kernel: candidate bins and value bounds of every cell
            if ((rc = aux_begin(c, "resolve", levelIdx))) return rc;
            SelPar prr = c->par;
            if (!sp.fine) prr.fine = nullptr;
            CK(launch_pdl(c, k_sel_resolve, dim3(std::min<uint32_t>(nCells, 4u * (uint32_t)c->nSM)), dim3(kThreads), (size_t)nb1 * 4, c->lv, ss, nCells, nb1,
                          0x7fffffffu, sp.sampleS, sp.z, prr));
            if ((rc = aux_end(c))) return rc;
            c->nOtherLaunch++;
        }
        {
            const size_t smem = ringBytes + (size_t)kWarps * kSelWarpStage * 4 + (sp.useBounds ? 0 : (size_t)nb1 * 4);
            int occ = 1;
            auto kcomp = sp.visits ? k_sel_stream<kSelCompact, true> : k_sel_stream<kSelCompact, false>;
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kcomp, kThreads, smem));
            const uint32_t grid = compactTiles ? ceil_div(nTiles, compactTiles)
                                               : std::min<uint32_t>(nTiles, (uint32_t)c->nSM * (uint32_t)std::min(std::max(occ, 1), 3));
            if ((rc = count_event_begin(c))) return rc;
            CK(launch_pdl(c, kcomp, dim3(grid), dim3(kThreads), smem, x, y, z, cand, c->lv, ss, (const uint32_t *)c->d_tile_first,
                          nCells, nL, nTiles, nb1, 1, (c->selBigFinish || sp.visits) ? 0x7fffffffu : candCap,
                          dbgBase ? dbgBase + (size_t)kDbgBlocks * 4 : (unsigned long long *)nullptr,
                          (float *)nullptr, 0u, sp.useBounds ? 1 : 0, no_arena(), pre, preTag, compactTiles, sp));
            if ((rc = count_event_end(c))) return rc;
        }
        if (sp.fine) {
            // ---- parallel FINISH: proof + scan of the fine histogram, gather of the ambiguous values, search on the short lists ----
            const uint32_t nCompact = ceil_div(nTiles, compactTiles);
            if ((rc = aux_begin(c, "finish", levelIdx))) return rc;
            CK(launch_pdl(c, k_sel_fine, dim3(nCompact), dim3(kThreads), 0, (const float *)cand, sp, c->par, nCompact));
            CK(launch_pdl(c, k_sel_fin_a, dim3(nCells), dim3(kThreads), 0, c->lv, ss, sp, c->par, nCells, nb1));
            CK(launch_pdl(c, k_sel_gather, dim3(nCompact), dim3(kThreads), 0, (const float *)cand, sp, c->par, nCompact));
            CK(launch_pdl(c, k_sel_fin_b, dim3(nCells), dim3(kThreads), 0, c->lv, ss, sc, c->par, nCells, c->d_err, 1, dn));
            if ((rc = aux_end(c))) return rc;
            c->nOtherLaunch += 3;
        } else {
            // (private regions: the search reads the pieces in place - no staging area, three piece tables instead)
            const uint32_t finishCap = sp.visits ? 0u : candCap;
            const size_t smem = sel_search_smem_bytes(finishCap) + (sp.visits ? (size_t)3 * kSelMaxPieces * 4 : 0);
            const int threads = (nCells <= 2u * (uint32_t)c->nSM || sp.visits) ? 1024 : search_threads(smem);
            int occ = 1;
            auto kfin = sp.visits ? k_sel_finish<true> : k_sel_finish<false>;
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kfin, threads, smem));
            const uint32_t grid = std::min<uint32_t>(nCells, (uint32_t)c->nSM * (uint32_t)std::max(occ, 1));
            if ((rc = aux_begin(c, "finish", levelIdx))) return rc;
            CK(launch_pdl(c, kfin, dim3(grid), dim3(threads), smem, (const float *)cand, c->lv, ss, sc, nCells, nb1, finishCap, c->d_err,
                          dbgBase ? dbgBase + (size_t)2 * kDbgBlocks * 4 : (unsigned long long *)nullptr, (preNb || sp.sampleS > 1) ? 1 : 2,
                          c->selBigFinish ? 1 : 0, dn, sp));
            if ((rc = aux_end(c))) return rc;
        }
        c->nCountLaunch += 1;
        c->nUpdateLaunch += 1;    // the finish kernel takes the place of the per-pass update kernels
    }
    CK(cudaGetLastError());
    // Cells the search could not finish are reported, not handled here: the last block of the search writes 1 + their
    // number into mapped memory; orb_build enqueues split + partition gated on that count and runs the iterative search
    // (select_fallback) only if the host then reads a non-zero count - no launch at all in the common case.
    return ORB_OK;
}

// iterative search for the cells the single-rank selection search flagged at this level (cooperative, host-free)
int select_fallback(orb_ctx *c, uint32_t nCells, int slotBase, int levelIdx) {
    int rc;
    if ((rc = aux_begin(c, "fallback", levelIdx))) return rc;
    const bool prof = c->profile;
    c->profile = false;            // its time belongs to the aux group, not to ms_count
    rc = launch_level_persistent(c, nCells, 3, slotBase, levelIdx, nullptr, c->sel.flag);
    c->profile = prof;
    if (rc) return rc;
    return aux_end(c);
}

// after the loop: cells that hit the iteration cap need one extra count at their final cut
int finalize_unfound(orb_ctx *c, uint32_t nCells, uint32_t *nUnfoundOut) {
    using namespace orb;
    CK(cudaMemsetAsync(c->d_misc, 0, 4, c->stream));
    k_finalize_prepare<<<ceil_div(nCells, 256), 256, 0, c->stream>>>(c->lv, nCells, c->d_misc);
    c->nOtherLaunch++;
    uint32_t n = 0;
    CK(cudaMemcpyAsync(&n, c->d_misc, 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (nUnfoundOut) *nUnfoundOut = n;
    // every rank sees the same n (found flags derive from global counts), so the collective below matches
    if (n) {
        int rc = launch_count(c, nCells, 1, nullptr);
        if (rc) return rc;
        rc = allreduce_counts(c, nCells, 1);
        if (rc) return rc;
        k_finalize_apply<<<ceil_div(nCells, 256), 256, 0, c->stream>>>(c->lv, nCells);
        c->nOtherLaunch++;
    }
    CK(cudaGetLastError());
    return ORB_OK;
}

// ---- selection search over several ranks: HIST -> allreduce(rows) -> COMPACT into slots -> prep -> all-gather(slots) ->
//      finish on the candidates of all ranks; cells it flags go to the host-driven iterative loop (run_bisection).
//      Every number that shapes a collective (nb1, slot size, whether the level qualifies) derives from nCells,
//      nLocalMin and nGlobal, which are identical on all ranks. ----
struct SelMrPlan {
    bool ok;
    bool v2;              // protocol of orb_exchange.cuh (peer memory); else the two-exchange kernels of orb_select.cuh
    int regime;           // v2: 0 streaming passes (k_sel_stream), 1 one block per cell, 2 one warp per cell (k_xd_*)
    bool warpFinish;      // v2: the owner searches a cell's candidates with one warp
    uint32_t candCapBig;  // v2 streaming levels: candidates a cell may have when the owner searches them in global scratch (0: candCap)
    int nb1, rep;
    uint32_t candCap, slotWords;
    size_t histWords;     // rows the level exchanges
    size_t zeroWords;     // rows that must be cleared before the level starts (built with atomics)
};
SelMrPlan sel_plan_mr(const orb_ctx *c, uint32_t nCells, int M, int forcedNb = 0) {
    SelMrPlan p{};
    p.ok = false;
    const bool multi = c->nRanks > 1 || c->mrSelf;
    if (!(c->select && c->selectMr && multi && c->nRanks <= orb::kMaxPeers && M == 3 && nCells >= 1 && c->nLocalMin > 0)) return p;
    p.v2 = c->mrV2 && c->peerEnabled;
    if (!p.v2 && !(c->d_slots_g && nCells <= kSelMrMaxCells)) return p;
    const size_t slotTotal = p.v2 ? c->slotTotal : kSelSlotWordsTotal;
    const uint64_t gavg = c->nGlobal / nCells, lavg = std::max<uint64_t>(c->nGlobal / c->nRanks, c->nLocalMin) / nCells;
    // <= every rank's selHistWords
    const size_t histFit = p.v2 ? (size_t)c->nLocalMin / 8 + 2 * (size_t)orb::kSelBinsMax : (size_t)c->nLocalMin / 16 + 2 * (size_t)orb::kSelBinsMax;
    if (p.v2 && ((nCells >= 512 && lavg <= (1u << 20)) || lavg <= 8192) && !forcedNb) {
        // ---- small cells: a group of threads per cell bins / gathers it; about 64 particles per bin over all ranks ----
        p.regime = lavg >= 4096 ? 1 : 2;        // (2: one warp per cell; its rows must fit the warp's shared memory)
        // Bins per cell: the level exchanges its rows (nCells * nb words per rank, reduced and broadcast) and its
        // candidates (about 1.5 bins' worth of particles per cell, once); the two balance at about 3 sqrt(cell size)
        // particles per bin.  (64 per bin - rows of 64 MB per level at 2^30 particles - cost 200 us per level in the
        // row exchange alone, profiles/r02d_c5_8gpu_levels.txt.)
        uint64_t perBin = 64;
        while (perBin * perBin < 4 * gavg) perBin <<= 1;          // power of two >= 2 sqrt(gavg)
        int nb = 32;
        while (nb < orb::kSelBins2 && gavg / (uint64_t)nb > perBin) nb <<= 1;
        while (nb > 32 && (size_t)nCells * (size_t)nb > histFit) nb >>= 1;
        // ... and a rank's slot (a few bins' worth of its own particles) must fit its share of the slot area: more bins
        // where it would not (a slot that overflows sends its cell to the iterative search)
        while (nb < orb::kSelBins2 && (size_t)nCells * (size_t)(2 * nb) <= histFit) {
            uint64_t w = 4 * (lavg / (uint64_t)nb) + 32 + 1, sw2 = 32;
            while (sw2 < w) sw2 <<= 1;
            if (sw2 * nCells <= slotTotal) break;
            nb <<= 1;
        }
        if (nb > 512) p.regime = 1;
        p.nb1 = nb;
        p.rep = 1;
        p.histWords = (size_t)nCells * (size_t)nb;
        p.zeroWords = 0;
        p.candCap = (uint32_t)std::min<uint64_t>(kSelValsCap, 4 * std::max<uint64_t>(1, gavg / (uint64_t)nb) + 128);
        p.warpFinish = p.candCap <= orb::kXWarpCap;
        const uint64_t want = std::min<uint64_t>(p.candCap, 4 * (lavg / (uint64_t)nb) + 32) + 1;
        uint32_t sw = 32;
        while (sw < want) sw <<= 1;
        while (sw > 32 && (size_t)sw * nCells > slotTotal) sw >>= 1;
        p.slotWords = sw;
        p.ok = p.histWords <= histFit && (size_t)sw * nCells <= slotTotal;
        return p;
    }
    p.regime = 0;
    p.nb1 = orb::kSelBinsMin;
    while (p.nb1 < orb::kSelBinsMax && gavg / (uint64_t)p.nb1 > (uint64_t)c->selBinAvg) p.nb1 <<= 1;
    if (forcedNb) p.nb1 = forcedNb;
    p.rep = p.nb1 <= 512 ? 4 : (p.nb1 <= 1024 ? 2 : 1);
    p.histWords = (size_t)nCells * (size_t)p.nb1;
    p.zeroWords = p.histWords;
    p.candCap = (uint32_t)std::min<uint64_t>(p.v2 ? c->xCandCap : kSelValsCap, 4 * (gavg / (uint64_t)p.nb1) + 4096);
    p.warpFinish = false;
    // slot of one rank and cell: a few bins' worth of its own particles + the count word, a power of two (v2: not bounded
    // by what a block stages in shared memory - the owner may search a huge cell in global scratch)
    uint64_t want = std::min<uint64_t>(p.v2 ? kSelSlotWordsMax - 1 : p.candCap, 4 * (lavg / (uint64_t)p.nb1) + 96) + 1;
    uint32_t sw = 128;
    while (sw < want) sw <<= 1;
    while (sw > 32 && (size_t)sw * nCells > slotTotal) sw >>= 1;
    p.slotWords = sw;
    p.ok = p.histWords <= histFit && (size_t)sw * nCells <= slotTotal;
    // cells too large for a block's shared memory even with the most bins (more than 2^26 particles): global scratch
    p.candCapBig = 0;
    if (p.v2 && 4 * (gavg / (uint64_t)p.nb1) + 4096 > c->xCandCap) {
        const uint64_t nOwnedMax = (nCells + c->nRanks - 1) / c->nRanks;
        const uint64_t big = std::min<uint64_t>((uint64_t)c->nRanks * (sw - 1u), kXScratchWords / nOwnedMax);
        if (big > p.candCap) p.candCapBig = (uint32_t)big;
    }
    return p;
}

// Bins of the histogram rows the partition of the previous level builds for a level of nNext cells (NextHist), or 0:
// the level must use the selection search, a bin must not hold more candidates than one block stages, the rows must
// fit.  Decided from rank-invariant numbers only (it shapes the multi-rank exchanges).
int prefuse_nb(const orb_ctx *c, uint32_t nNext, int M) {
    if (c->prefuseHist == 0) return 0;
    // sampled rows (single rank) replace the HIST pass for a fraction of its cost without touching the partition
    if (c->prefuseHist < 0 && sampling_on(c)) return 0;
    // Measured (profiles/r01x_*, r01z_*): binning in the partition costs ~10 us per level at 2^24 particles - as much as
    // the HIST pass it replaces, whose column is still half in L2 at that size - but pays from ~2^25 particles per GPU
    // (HIST is then a full HBM pass: -9 % build time at 2^27) and whenever ranks have to agree on the rows anyway.
    if (c->prefuseHist < 0 && c->nRanks == 1 && c->nLocal < (1ull << 25)) return 0;
    // tiny cells are partitioned by one warp each (k_partition_warp), which does not bin
    if (nNext >= 2 && !partition_is_coop(c, nNext / 2) && c->nLocal / (nNext / 2) <= (uint64_t)c->partWarpMax) return 0;
    // The partition's block histogram holds at most 1024 bins per child.  Where the level streams (HIST / COMPACT /
    // FINISH) the rows must have the bins the level would choose itself - coarser rows mean more candidates per bin
    // and, on clustered inputs, cells that overflow the finish kernel's staging; where one block searches a whole
    // cell it can zoom on its own (k_sel_percell), so coarser rows only cost that cell another read.
    if (c->nRanks > 1 || c->mrSelf) {
        const SelMrPlan p = sel_plan_mr(c, nNext, M);
        if (!p.ok || p.regime != 0 || p.nb1 > 1024) return 0;
        return p.nb1;
    }
    if (!level_can_select(c, nNext, M)) return 0;
    const SelPlan p = sel_plan(c, nNext);
    if (!p.cellsInSmem && p.nb1 > 1024) return 0;
    const int nb = std::min(p.nb1, 1024);
    if ((size_t)nNext * (size_t)nb > c->selHistWords) return 0;
    return nb;
}

// Enqueues the level's search up to the finish kernel; no host synchronisation.  The finish kernel's last block
// reports 1 + (cells flagged) in h_status[slotBase + kPassSlots - 1] (see select_mr_flagged).
int launch_level_select_mr(orb_ctx *c, uint32_t nCells, const SelMrPlan &pl, int slotBase, int levelIdx, int preNb = 0) {
    using namespace orb;
    c->chunkTiles = 0;
    c->preValid = false;
    const float *x = c->x[c->cur], *y = c->y[c->cur], *z = c->z[c->cur];
    float *cand = c->x[c->cur ^ 1];
    const bool peer = c->peerEnabled && c->peerX[c->rank] != nullptr;
    SelState ss = c->sel;                       // hist = this rank's rows
    ss.n_flagged = c->d_sel_nflag + levelIdx;
    SelState sg = ss;
    sg.hist = c->d_sel_hist_g;                  // NCCL transport: rows summed over ranks
    SelCtl sc;
    sc.active_particles = c->d_active_particles;
    sc.level_iters = c->d_level_iters + levelIdx;
    sc.passes_out = c->d_lvl_passes + levelIdx;
    sc.n_unfound_out = c->d_lvl_unfound + levelIdx;
    SelMrState mr;
    mr.hist_l = c->sel.hist;
    mr.loc_base = c->d_sel_locbase;
    mr.slots_l = c->d_slots_l;
    mr.slots_g = c->d_slots_g;
    mr.slotWords = pl.slotWords;
    mr.nRanks = c->nRanks;
    mr.self = c->rank;
    mr.done = c->d_cdone + slotBase + kPassSlots - 1;
    mr.h_status = (volatile uint32_t *)(c->h_status_dev + slotBase + kPassSlots - 1);
    SelPeers px;
    memset(&px, 0, sizeof(px));
    if (peer) {
        px.n = c->nRanks;
        px.self = c->rank;
        for (int r = 0; r < c->nRanks; ++r) px.arena[r] = c->peerX[r];
        px.offCursor = c->xOffCursor; px.offSlots = c->xOffSlots; px.offHist = c->xOffHist;
    }
    int rc;
    const int nb1 = pl.nb1;
    const uint32_t nTiles = ceil_div(c->nLocal, kCountTile);
    const size_t ringBytes = (size_t)kCountStages * kCountTile * sizeof(float);
    const uint32_t nL = (uint32_t)c->nLocal;
    // ---- HIST: this rank's rows (cleared by level_prepare); not needed when the previous level's partition built them ----
    if (nTiles && !preNb) {
        const size_t smem = ringBytes + (size_t)nb1 * pl.rep * 4;
        int occ = 1;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_sel_stream<kSelHist>, kThreads, smem));
        const uint32_t grid = std::min<uint32_t>(nTiles, (uint32_t)c->nSM * (uint32_t)std::min(std::max(occ, 1), 4));
        if ((rc = count_event_begin(c))) return rc;
        CK(launch_pdl(c, k_sel_stream<kSelHist>, dim3(grid), dim3(kThreads), smem, x, y, z, cand, c->lv, ss, (const uint32_t *)c->d_tile_first,
                      nCells, nL, nTiles, nb1, pl.rep, pl.candCap, (unsigned long long *)nullptr, (float *)nullptr, 0u, 0, no_arena(), (PreLeft *)nullptr, 0u, 0u, no_priv()));
        if ((rc = count_event_end(c))) return rc;
        c->nCountLaunch++;
    }
    // ---- exchange 1: rows over ranks, resolve ----
    if (peer) {
        px.seq = ++c->xSeq;
        const uint32_t grid = std::min<uint32_t>(nCells, (uint32_t)c->nSM * 4u);
        if ((rc = aux_begin(c, "mr_resolve", levelIdx))) return rc;
        CK(launch_pdl(c, k_selx_resolve, dim3(grid), dim3(kThreads), (size_t)nb1 * 4, c->lv, ss, mr, px, nCells, nb1, pl.candCap));
        if ((rc = aux_end(c))) return rc;
        c->nOtherLaunch++;
    } else {
        NK(g_nccl.AllReduce(c->sel.hist, c->d_sel_hist_g, pl.histWords, ncclUint32, ncclSum, c->comm, c->stream));
    }
    // ---- COMPACT: own candidates into the cells' slots ----
    if (nTiles) {
        const size_t smem = ringBytes + (size_t)kWarps * kSelWarpStage * 4 + (size_t)nb1 * 4;
        int occ = 1;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_sel_stream<kSelCompact>, kThreads, smem));
        const uint32_t grid = std::min<uint32_t>(nTiles, (uint32_t)c->nSM * (uint32_t)std::min(std::max(occ, 1), 3));
        if ((rc = count_event_begin(c))) return rc;
        CK(launch_pdl(c, k_sel_stream<kSelCompact>, dim3(grid), dim3(kThreads), smem, x, y, z, cand, c->lv, peer ? ss : sg,
                      (const uint32_t *)c->d_tile_first, nCells, nL, nTiles, nb1, 1, pl.candCap, (unsigned long long *)nullptr,
                      c->d_slots_l, pl.slotWords, peer ? 1 : 0, no_arena(), (PreLeft *)nullptr, 0u, 0u, no_priv()));
        if ((rc = count_event_end(c))) return rc;
        c->nCountLaunch++;
    }
    // ---- exchange 2: candidates over ranks, block search ----
    if (!peer) {
        const uint32_t grid = std::min<uint32_t>(nCells, (uint32_t)c->nSM * 8u);
        if ((rc = aux_begin(c, "mr_prep", levelIdx))) return rc;
        CK(launch_pdl(c, k_selmr_prep, dim3(grid), dim3(kThreads), (size_t)nb1 * 4, c->lv, sg, mr, nCells, nb1, pl.candCap));
        if ((rc = aux_end(c))) return rc;
        c->nOtherLaunch++;
        NK(g_nccl.AllGather(c->d_slots_l, c->d_slots_g, (size_t)nCells * pl.slotWords, ncclUint32, c->comm, c->stream));
    }
    {
        const size_t smem = sel_search_smem_bytes(pl.candCap);
        const int threads = nCells <= 2u * (uint32_t)c->nSM ? 1024 : (smem > 112 * 1024 ? 1024 : (smem > 56 * 1024 ? 512 : 256));
        auto kern = peer ? k_selmr_finish<true> : k_selmr_finish<false>;
        int occ = 1;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
        const uint32_t grid = std::min<uint32_t>(nCells, (uint32_t)c->nSM * (uint32_t)std::max(occ, 1));
        if (peer) px.seq = ++c->xSeq;
        if ((rc = aux_begin(c, "mr_finish", levelIdx))) return rc;
        CK(launch_pdl(c, kern, dim3(grid), dim3(threads), smem, c->lv, peer ? ss : sg, sc, mr, px, nCells, nb1, pl.candCap, c->d_err, preNb ? 1 : 2));
        if ((rc = aux_end(c))) return rc;
        c->nUpdateLaunch++;
    }
    CK(cudaGetLastError());
    return ORB_OK;
}

// Protocol v2 (orb_exchange.cuh): HIST -> REDUCE -> COMPACT (-> PUSH) -> FINISH by the owner -> APPLY.  No collective
// call, no host synchronisation; the apply kernel's last block reports 1 + (cells flagged) like the v1 finish kernel.
orb::XArena make_arena(orb_ctx *c) {
    orb::XArena xa;
    memset(&xa, 0, sizeof(xa));
    xa.n = c->nRanks;
    xa.self = c->rank;
    for (int r = 0; r < c->nRanks; ++r) xa.arena[r] = c->peerX[r];
    xa.offRes = c->xOffRes; xa.offRecvCnt = c->xOffRecvCnt; xa.offRecv = c->xOffRecv;
    xa.offHistL = c->xOffHist; xa.offHistG = c->xOffHistG;
    return xa;
}

int launch_level_select_mr2(orb_ctx *c, uint32_t nCells, const SelMrPlan &pl, int slotBase, int levelIdx, int preNb = 0) {
    using namespace orb;
    const float *x = c->x[c->cur], *y = c->y[c->cur], *z = c->z[c->cur];
    float *cand = c->x[c->cur ^ 1];
    SelState ss = c->sel;                       // hist = this rank's rows
    ss.n_flagged = c->d_sel_nflag + levelIdx;
    SelState sg = ss;
    sg.hist = c->d_xchg + c->xOffHistG;         // rows summed over ranks (this rank's copy)
    SelCtl sc;
    sc.active_particles = c->d_active_particles;
    sc.level_iters = c->d_level_iters + levelIdx;
    sc.passes_out = c->d_lvl_passes + levelIdx;
    sc.n_unfound_out = c->d_lvl_unfound + levelIdx;
    SelMrState mr;
    memset(&mr, 0, sizeof(mr));
    mr.hist_l = c->sel.hist;
    mr.loc_base = c->d_sel_locbase;
    mr.slots_l = c->d_slots_l;
    mr.slotWords = pl.slotWords;
    mr.nRanks = c->nRanks;
    mr.self = c->rank;
    mr.done = c->d_cdone + slotBase + kPassSlots - 1;
    mr.h_status = (volatile uint32_t *)(c->h_status_dev + slotBase + kPassSlots - 1);
    XArena xa = make_arena(c);
    int rc;
    const int nb1 = pl.nb1;
    const uint32_t nTiles = ceil_div(c->nLocal, kCountTile);
    const size_t ringBytes = (size_t)kCountStages * kCountTile * sizeof(float);
    const uint32_t nL = (uint32_t)c->nLocal;
    const uint32_t nSM = (uint32_t)c->nSM;
    // streaming levels: the COMPACT pass records the partition's per-block left counts (candidates in this rank's slots)
    c->chunkTiles = pl.regime == 0 ? level_chunk_tiles(c, nCells) : 0u;
    c->preValid = c->chunkTiles != 0u;
    c->preListStride = pl.slotWords;
    ++c->preTagSeq;
    const uint32_t nOwnedMax = ceil_div(nCells, (uint32_t)c->nRanks);
    // ---- HIST: this rank's rows ----
    if (pl.regime == 0) {
        if (nTiles && !preNb) {       // (preNb: the previous level's partition built them)
            const size_t smem = ringBytes + (size_t)nb1 * pl.rep * 4;
            int occ = 1;
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_sel_stream<kSelHist>, kThreads, smem));
            const uint32_t grid = std::min<uint32_t>(nTiles, nSM * (uint32_t)std::min(std::max(occ, 1), 4));
            if ((rc = count_event_begin(c))) return rc;
            CK(launch_pdl(c, k_sel_stream<kSelHist>, dim3(grid), dim3(kThreads), smem, x, y, z, cand, c->lv, ss, (const uint32_t *)c->d_tile_first,
                          nCells, nL, nTiles, nb1, pl.rep, pl.candCap, (unsigned long long *)nullptr, (float *)nullptr, 0u, 0, no_arena(), (PreLeft *)nullptr, 0u, 0u, no_priv()));
            if ((rc = count_event_end(c))) return rc;
            c->nCountLaunch++;
        }
    } else {
        if ((rc = count_event_begin(c))) return rc;
        if (pl.regime == 1) {
            const uint32_t grid = std::min<uint32_t>(nCells, nSM * 8u);
            CK(launch_pdl(c, k_xd_hist<256>, dim3(grid), dim3(kThreads), (size_t)nb1 * 4, x, y, z, c->lv, c->sel.hist, nCells, nb1));
        } else {
            const uint32_t grid = std::min<uint32_t>(ceil_div(nCells, kWarps), nSM * 8u);
            CK(launch_pdl(c, k_xd_hist_warp, dim3(grid), dim3(kThreads), (size_t)nb1 * 4 * kWarps, x, y, z, c->lv, c->sel.hist, nCells, nb1));
        }
        if ((rc = count_event_end(c))) return rc;
        c->nCountLaunch++;
    }
    // ---- REDUCE: rows of all ranks -> global rows on every rank ----
    {
        const uint32_t n4 = (uint32_t)(pl.histWords / 4);
        xa.seq = ++c->xSeq;
        const uint32_t grid = std::max<uint32_t>(1u, std::min<uint32_t>(ceil_div(ceil_div(n4, (uint32_t)c->nRanks), kThreads), 2u * nSM));
        if ((rc = aux_begin(c, "x_reduce", levelIdx))) return rc;
        CK(launch_pdl(c, k_xr_reduce, dim3(grid), dim3(kThreads), 0, xa, n4));
        if ((rc = aux_end(c))) return rc;
        c->nOtherLaunch++;
    }
    // ---- COMPACT: resolve from the global rows, own candidates to the owners ----
    const uint32_t resolveCap = std::max(pl.candCap, pl.candCapBig);     // candidates a cell may have without being flagged
    xa.seq = ++c->xSeq;
    if (pl.regime == 0) {
        if (nTiles) {
            const size_t smem = ringBytes + (size_t)kWarps * kSelWarpStage * 4 + (size_t)nb1 * 4;
            int occ = 1;
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_sel_stream<kSelCompact>, kThreads, smem));
            const uint32_t grid = c->chunkTiles ? ceil_div(nTiles, c->chunkTiles) : std::min<uint32_t>(nTiles, nSM * (uint32_t)std::min(std::max(occ, 1), 3));
            if ((rc = count_event_begin(c))) return rc;
            CK(launch_pdl(c, k_sel_stream<kSelCompact>, dim3(grid), dim3(kThreads), smem, x, y, z, cand, c->lv, sg,
                          (const uint32_t *)c->d_tile_first, nCells, nL, nTiles, nb1, 1, resolveCap, (unsigned long long *)nullptr,
                          c->d_slots_l, pl.slotWords, 0, xa, c->preValid ? c->d_pre : (PreLeft *)nullptr, c->preTagSeq, c->chunkTiles, no_priv()));
            if ((rc = count_event_end(c))) return rc;
            c->nCountLaunch++;
        }
        {
            // (no particles on this rank: the barrier the COMPACT pass would have made is made here)
            XArena xb = xa;
            if (nTiles) xb.n = 0;
            const uint32_t grid = std::min<uint32_t>(nCells, nSM * 8u);
            if ((rc = aux_begin(c, "x_push", levelIdx))) return rc;
            CK(launch_pdl(c, k_xc_push, dim3(grid), dim3(kThreads), (size_t)nb1 * 4, c->lv, sg, mr, xa, xb, nCells, nb1, resolveCap));
            if ((rc = aux_end(c))) return rc;
            c->nOtherLaunch++;
        }
    } else {
        if ((rc = count_event_begin(c))) return rc;
        if (pl.regime == 1) {
            const uint32_t grid = std::min<uint32_t>(nCells, nSM * 8u);
            CK(launch_pdl(c, k_xd_compact<256>, dim3(grid), dim3(kThreads), (size_t)nb1 * 4, x, y, z, c->lv, ss, mr, xa, nCells, nb1, pl.candCap));
        } else {
            const uint32_t grid = std::min<uint32_t>(ceil_div(nCells, kWarps), nSM * 8u);
            CK(launch_pdl(c, k_xd_compact_warp, dim3(grid), dim3(kThreads), xd_compact_warp_smem(nb1), x, y, z, c->lv, ss, mr, xa, nCells, nb1, pl.candCap));
        }
        if ((rc = count_event_end(c))) return rc;
        c->nCountLaunch++;
    }
    // ---- FINISH: the owner searches its cells' candidates, results to every rank ----
    xa.seq = ++c->xSeq;
    if ((rc = aux_begin(c, "x_finish", levelIdx))) return rc;
    if (pl.warpFinish) {
        const size_t smem = (size_t)kWarps * (pl.candCap + 4u) * 4u;
        const uint32_t grid = std::max<uint32_t>(1u, std::min<uint32_t>(ceil_div(nOwnedMax, kWarps), nSM * 6u));
        CK(launch_pdl(c, k_xf_finish_warp, dim3(grid), dim3(kThreads), smem, c->lv, ss, mr, xa, nCells, nb1, pl.candCap, c->d_err));
    } else {
        const size_t smem = sel_search_smem_bytes(pl.candCap);
        const int threads = nOwnedMax <= 2u * nSM ? 1024 : (smem > 112 * 1024 ? 1024 : (smem > 56 * 1024 ? 512 : 256));
        int occ = 1;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_xf_finish_block, threads, smem));
        const uint32_t grid = std::max<uint32_t>(1u, std::min<uint32_t>(nOwnedMax, nSM * (uint32_t)std::max(occ, 1)));
        CK(launch_pdl(c, k_xf_finish_block, dim3(grid), dim3(threads), smem, c->lv, ss, mr, xa, nCells, nb1, pl.candCap, c->d_err,
                      pl.candCapBig ? c->d_xscratch : (float *)nullptr, pl.candCapBig));
    }
    if ((rc = aux_end(c))) return rc;
    c->nUpdateLaunch++;
    // ---- APPLY ----
    xa.seq = ++c->xSeq;
    if ((rc = aux_begin(c, "x_apply", levelIdx))) return rc;
    CK(launch_pdl(c, k_xa_apply, dim3(ceil_div(nCells, kThreads)), dim3(kThreads), 0, c->lv, ss, sc, mr, xa, nCells,
                  pl.regime == 0 ? (preNb ? 1 : 2) : 2));
    if ((rc = aux_end(c))) return rc;
    c->nOtherLaunch++;
    CK(cudaGetLastError());
    return ORB_OK;
}

// Cells the multi-rank selection search left over at this level (the same number on every rank).  Spins on the
// mapped status word the finish kernel's last block writes - no stream synchronisation.
int select_mr_flagged(orb_ctx *c, int slotBase, uint32_t *nFlagged) {
    uint32_t s = 0;
    const int rc = wait_status(c, c->h_status + slotBase + kPassSlots - 1, &s);
    if (rc) return rc;
    *nFlagged = s - 1u;
    return ORB_OK;
}

// iterative loop for the cells the multi-rank selection search flagged (every rank runs it: the flags agree)
int select_mr_fallback(orb_ctx *c, uint32_t nCells, int slotBase, int levelIdx) {
    int np = 0, rc;
    if ((rc = run_bisection(c, nCells, 3, slotBase, levelIdx, &np))) return rc;
    if (np >= (orb::kMaxIter + 2) / 3) {
        uint32_t nu = 0;
        if ((rc = finalize_unfound(c, nCells, &nu))) return rc;
        // nu counts every cell of the level without a found cut (also those the selection search itself capped):
        // it replaces the device-side statistic of the level
        if ((rc = ensure_scratch(c, 64))) return rc;
        c->h_scratch[1] = nu;
        CK(cudaMemcpyAsync(c->d_lvl_unfound + levelIdx, c->h_scratch + 1, 4, cudaMemcpyHostToDevice, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    if ((size_t)levelIdx < c->extraPasses.size()) c->extraPasses[levelIdx] = np;
    return ORB_OK;
}

// Stable split of every cell of the level (canonical tie mode).  Kernel choice by average local cell size:
// cells of at most 16 tiles (and enough of them to fill the GPU) -> one block per cell with a running carry;
// otherwise cooperative reduce-then-scan over contiguous tile ranges (k_partition_coop: all blocks co-resident).
// `nh`: also build the next level's histogram rows (NextHist).
orb::NextHist no_next_hist() {
    orb::NextHist nh;
    memset(&nh, 0, sizeof(nh));
    return nh;
}

int launch_partition(orb_ctx *c, uint32_t nCells, uint32_t *ticket, const uint32_t *gate = nullptr, orb::NextHist nh = no_next_hist(),
                     bool usePre = false /* this level's search left PreLeft records for the chunking c->chunkTiles */) {
    using namespace orb;
    (void)ticket;
    const uint32_t nTiles = ceil_div(c->nLocal, kPartTile);
    if (!nTiles) return ORB_OK;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (c->profile) {
        if (c->evPartUsed == c->evPart.size()) {
            cudaEvent_t a, b;
            CK(cudaEventCreate(&a));
            CK(cudaEventCreate(&b));
            c->evPart.emplace_back(a, b);
        }
        e0 = c->evPart[c->evPartUsed].first;
        e1 = c->evPart[c->evPartUsed].second;
        c->evPartUsed++;
        CK(cudaEventRecord(e0, c->stream));
    }
    const int o = c->cur ^ 1;
    const float *x = c->x[c->cur], *y = c->y[c->cur], *z = c->z[c->cur];
    float *x2 = c->x[o], *y2 = c->y[o], *z2 = c->z[o];
    const uint64_t avg = c->nLocal / nCells;
    const size_t smem = sizeof(PartSmem);
    if (!partition_is_coop(c, nCells) && avg <= (uint64_t)c->partWarpMax && !nh.enabled) {
        // tiny cells: one warp per cell
        const uint32_t grid = std::min<uint32_t>(ceil_div(nCells, kWarps), (uint32_t)c->nSM * 8u);
        k_partition_warp<<<grid, kThreads, 0, c->stream>>>(x, y, z, x2, y2, z2, c->lv, c->d_final_cut, nCells, gate);
    } else if (!partition_is_coop(c, nCells)) {
        const uint32_t grid = std::min<uint32_t>(nCells, (uint32_t)c->nSM * (uint32_t)c->occPartCells);
        k_partition_cells<<<grid, kThreads, smem, c->stream>>>(x, y, z, x2, y2, z2, c->lv, c->d_final_cut, nCells, (uint32_t)c->nLocal, gate, nh);
    } else {
        uint32_t grid = std::min<uint32_t>(nTiles, (uint32_t)c->nSM * (uint32_t)c->occPartStream);
        // PreLeft: same chunk per block as the search's last pass (count tiles of 4096 = two partition tiles)
        const PreLeft *pre = nullptr;
        uint32_t preTag = 0, tpb = 0, listStride = 0;
        const float *preList = nullptr;
        if (usePre && c->preValid && c->chunkTiles) {
            tpb = 2u * c->chunkTiles;
            grid = ceil_div(ceil_div(c->nLocal, kCountTile), c->chunkTiles);
            pre = c->d_pre;
            preTag = c->preTagSeq;
            if (c->nRanks > 1 || c->mrSelf) { preList = c->d_slots_l; listStride = c->preListStride; }
            else preList = x2;      // the cells' candidate lists lie in the idle column, which this launch overwrites - after phase 1
        }
        uint32_t nLocal32 = (uint32_t)c->nLocal, nT = nTiles, nC = nCells;
        void *args[] = {(void *)&x, (void *)&y, (void *)&z, (void *)&x2, (void *)&y2, (void *)&z2, (void *)&c->lv,
                        (void *)&c->d_final_cut, (void *)&c->d_tile_first, (void *)&nC, (void *)&nLocal32, (void *)&nT,
                        (void *)&c->d_blk_left, (void *)&c->d_blk_restart, (void *)&gate, (void *)&nh,
                        (void *)&pre, (void *)&preTag, (void *)&tpb, (void *)&preList, (void *)&listStride, (void *)&c->partBulk};
        CK(cudaLaunchCooperativeKernel((const void *)k_partition_coop, dim3(grid), dim3(kThreads), args, smem, c->stream));
    }
    if (c->profile) CK(cudaEventRecord(e1, c->stream));
    c->nPartLaunch++;
    CK(cudaGetLastError());
    c->cur = o;
    return ORB_OK;
}

// Hoare-exact partition of every cell of the level, in place (no ping-pong flip): rank scan -> pair swaps -> finish.
// Leaves the local left counts in lv.nleft_l and their sum over ranks in lv.nleft_g.
int launch_partition_hoare(orb_ctx *c, uint32_t nCells) {
    using namespace orb;
    const uint32_t nTiles = ceil_div(c->nLocal, kPartTile);
    const uint32_t cb = ceil_div(nCells, 256);
    k_final_cut<<<cb, 256, 0, c->stream>>>(c->lv, nCells, c->d_final_cut);
    c->nOtherLaunch++;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (c->profile) {
        if (c->evPartUsed == c->evPart.size()) {
            cudaEvent_t a, b;
            CK(cudaEventCreate(&a));
            CK(cudaEventCreate(&b));
            c->evPart.emplace_back(a, b);
        }
        e0 = c->evPart[c->evPartUsed].first;
        e1 = c->evPart[c->evPartUsed].second;
        c->evPartUsed++;
        CK(cudaEventRecord(e0, c->stream));
    }
    if (nTiles) {
        float *x = c->x[c->cur], *y = c->y[c->cur], *z = c->z[c->cur];
        const float *cx = x, *cy = y, *cz = z;
        HoareLists hl;
        hl.posI = reinterpret_cast<uint32_t *>(c->y[c->cur ^ 1]);   // the idle ping-pong columns hold the stopper lists
        hl.posJ = reinterpret_cast<uint32_t *>(c->z[c->cur ^ 1]);
        hl.nGE = c->d_nGE;
        hl.nLE = c->d_nLE;
        const uint32_t grid = std::min<uint32_t>(nTiles, (uint32_t)c->nSM * (uint32_t)std::min(c->occHoare, 4));
        uint32_t nLocal32 = (uint32_t)c->nLocal, nT = nTiles, nC = nCells;
        void *args[] = {(void *)&cx, (void *)&cy, (void *)&cz, (void *)&c->lv, (void *)&c->d_final_cut, (void *)&c->d_tile_first,
                        (void *)&nC, (void *)&nLocal32, (void *)&nT, (void *)&hl, (void *)&c->d_blk_left, (void *)&c->d_blk_le,
                        (void *)&c->d_blk_restart};
        CK(cudaLaunchCooperativeKernel((const void *)k_hoare_scan, dim3(grid), dim3(kThreads), args, 0, c->stream));
        const uint32_t sgrid = std::min<uint32_t>(ceil_div(c->nLocal, kThreads), (uint32_t)c->nSM * 8u);
        k_hoare_swap<<<sgrid, kThreads, 0, c->stream>>>(x, y, z, c->lv.bnd, c->d_tile_first, nCells, nLocal32, hl);
        k_hoare_finish<<<cb, 256, 0, c->stream>>>(x, y, z, c->lv, nCells, hl, c->d_err);
        c->nPartLaunch += 3;
    }
    if (c->profile) CK(cudaEventRecord(e1, c->stream));
    if (c->nRanks > 1) NK(g_nccl.AllReduce(c->lv.nleft_l, c->lv.nleft_g, nCells, ncclUint32, ncclSum, c->comm, c->stream));
    else k_copy_u32<<<cb, 256, 0, c->stream>>>(c->lv.nleft_l, c->lv.nleft_g, nCells);
    c->nOtherLaunch++;
    CK(cudaGetLastError());
    return ORB_OK;
}

// Exchange arena: flags[64] | slot fill levels | own candidate slots | (v2: result records | candidate counts | received
// slots) | own histogram rows | (v2: global rows).  One allocation, one IPC handle; with several ranks every offset
// derives from numbers all ranks share (d, the largest shard).
int alloc_arena(orb_ctx *c, bool multi) {
    const size_t L = c->maxLevelCells;
    const size_t Lr = (L + 63) & ~(size_t)63;
    if (c->d_xchg) { CK(cudaFree(c->d_xchg)); c->d_xchg = nullptr; }
    c->slotTotal = kSelSlotWordsTotal;
    size_t off = 64;
    c->xOffCursor = (uint32_t)off; off += Lr;
    if (multi) {
        c->selHistWords = (size_t)c->nLocalMax / 8 + 2 * (size_t)orb::kSelBinsMax;
        c->selHistWords = (c->selHistWords + 63) & ~(size_t)63;
        c->slotTotal = std::max<size_t>(std::max<size_t>(kSelSlotWordsTotal, 64 * L), (((size_t)c->nLocalMax / 8) + 63) & ~(size_t)63);
    }
    c->xOffSlots = (uint32_t)off; off += c->slotTotal;
    if (multi) {
        c->xOffRes = (uint32_t)off; off += 8 * Lr;
        c->xOffRecvCnt = (uint32_t)off; off += Lr + 64;
        c->xOffRecv = (uint32_t)off; off += c->slotTotal + (size_t)orb::kMaxPeers * kSelSlotWordsMax;
    }
    c->xOffHist = (uint32_t)off; off += c->selHistWords;
    if (multi) { c->xOffHistG = (uint32_t)off; off += c->selHistWords; }
    if (off >= ((size_t)1 << 32)) return fail(ORB_ERR_ARG, "exchange arena of %zu words exceeds 32-bit word offsets", off);
    if (multi && !c->d_xscratch) CK(cudaMalloc(&c->d_xscratch, kXScratchWords * 4));
    CK(cudaMalloc(&c->d_xchg, off * 4));
    CK(cudaMemset(c->d_xchg, 0, off * 4));
    c->sel.cursor = c->d_xchg + c->xOffCursor;
    c->d_slots_l = reinterpret_cast<float *>(c->d_xchg + c->xOffSlots);
    c->sel.hist = c->d_xchg + c->xOffHist;
    return ORB_OK;
}

int reset_pass_ctl(orb_ctx *c) {
    CK(cudaMemsetAsync(c->d_nactive, 0, sizeof(uint32_t) * kMaxLevels * kPassSlots, c->stream));
    CK(cudaMemsetAsync(c->d_done, 0, sizeof(uint32_t) * kMaxLevels * kPassSlots, c->stream));
    CK(cudaMemsetAsync(c->d_tickets, 0, sizeof(uint32_t) * kMaxLevels, c->stream));
    CK(cudaMemsetAsync(c->d_cdone, 0, sizeof(uint32_t) * kMaxLevels * kPassSlots, c->stream));
    CK(cudaMemsetAsync(c->d_lvl_passes, 0, sizeof(int32_t) * kMaxLevels, c->stream));
    CK(cudaMemsetAsync(c->d_lvl_unfound, 0, sizeof(uint32_t) * kMaxLevels, c->stream));
    CK(cudaMemsetAsync(c->d_sel_nflag, 0, sizeof(uint32_t) * kMaxLevels, c->stream));
    CK(cudaMemsetAsync(c->d_level_iters, 0, sizeof(int32_t) * kMaxLevels, c->stream));
    CK(cudaMemsetAsync(c->d_active_particles, 0, 2 * sizeof(unsigned long long), c->stream));
    // the previous call's speculative passes may still be writing status words: drain first
    CK(cudaStreamSynchronize(c->stream));
    memset((void *)c->h_status, 0, sizeof(uint32_t) * kMaxLevels * kPassSlots);
    return ORB_OK;
}

int upload_cells(orb_ctx *c, const orb_cell *cells, uint32_t nCells) {
    if (!cells || nCells == 0 || nCells > c->maxLevelCells) return fail(ORB_ERR_ARG, "bad cell array (n=%u, max %u)", nCells, c->maxLevelCells);
    for (uint32_t i = 0; i < nCells; ++i)
        if (cells[i].id < 0 || (uint32_t)cells[i].id >= c->nHeap) return fail(ORB_ERR_ARG, "cell id %d outside heap of %u", cells[i].id, c->nHeap);
    CK(cudaMemcpyAsync(c->d_cells, cells, (size_t)nCells * sizeof(orb_cell), cudaMemcpyHostToDevice, c->stream));
    return ORB_OK;
}

float sum_events(std::vector<std::pair<cudaEvent_t, cudaEvent_t>> &ev, size_t used) {
    float tot = 0.f;
    for (size_t i = 0; i < used; ++i) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ev[i].first, ev[i].second) == cudaSuccess) tot += ms;
    }
    return tot;
}

}  // namespace

// =====================================================================================
extern "C" {

const char *orb_last_error(void) { return g_err; }
int orb_version(void) { return 100; }

static int create_impl(orb_ctx **out, int device, uint64_t n_local, uint32_t n_leaf_cells, orb_ctx **partial);

int orb_create(orb_ctx **out, int device, uint64_t n_local, uint32_t n_leaf_cells) {
    if (!out) return fail(ORB_ERR_ARG, "null ctx pointer");
    *out = nullptr;
    orb_ctx *partial = nullptr;
    const int rc = create_impl(out, device, n_local, n_leaf_cells, &partial);
    if (rc != ORB_OK && partial) {      // e.g. cudaMalloc ran out of memory half-way: give everything back
        char keep[sizeof(g_err)];
        memcpy(keep, g_err, sizeof(keep));
        orb_destroy(partial);           // tolerates null members
        memcpy(g_err, keep, sizeof(keep));
        cudaGetLastError();
        *out = nullptr;
    }
    return rc;
}

static int create_impl(orb_ctx **out, int device, uint64_t n_local, uint32_t n_leaf_cells, orb_ctx **partial) {
    if (n_leaf_cells < 1 || (n_leaf_cells & (n_leaf_cells - 1))) return fail(ORB_ERR_ARG, "n_leaf_cells must be a power of two (orbit.cpp:42)");
    if (n_local >= (1ull << 32) - orb::kCountTile) return fail(ORB_ERR_ARG, "n_local must fit 32-bit particle indices");
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(ORB_ERR_CUDA, "no CUDA device %d (found %d); there is no CPU fallback", device, ndev);
    CK(cudaSetDevice(device));
    orb_ctx *c = new orb_ctx();
    *partial = c;
    c->device = device;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    c->nSM = prop.multiProcessorCount;
    c->nLocal = n_local;
    c->d = n_leaf_cells;
    c->nHeap = 2 * n_leaf_cells - 1;
    c->maxLevelCells = std::max<uint32_t>(1u, n_leaf_cells);   // deepest level we may be handed (children of the last split level)
    CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    const size_t nPad = (size_t)n_local + orb::kCountTile;   // tail tiles may be read with vector loads only when fully inside; pad anyway
    for (int b = 0; b < 2; ++b) {
        CK(cudaMalloc(&c->x[b], nPad * 4));
        CK(cudaMalloc(&c->y[b], nPad * 4));
        CK(cudaMalloc(&c->z[b], nPad * 4));
    }
    const size_t L = c->maxLevelCells;
    CK(cudaMalloc(&c->d_heap, (size_t)c->nHeap * sizeof(orb_cell)));
    CK(cudaMalloc(&c->d_cells, L * sizeof(orb_cell)));
    CK(cudaMalloc(&c->d_range, (size_t)c->nHeap * 8));
    CK(cudaMalloc(&c->d_total, (size_t)c->nHeap * 4));
    CK(cudaMemset(c->d_range, 0, (size_t)c->nHeap * 8));
    CK(cudaMemset(c->d_total, 0, (size_t)c->nHeap * 4));
    CK(cudaMalloc(&c->lv.bnd, (L + 1) * 4));
    CK(cudaMalloc(&c->lv.axis, L * 4));
    CK(cudaMalloc(&c->lv.mL, L * 4));
    CK(cudaMalloc(&c->lv.mR, L * 4));
    CK(cudaMalloc(&c->lv.total, L * 4));
    CK(cudaMalloc(&c->lv.nleaf, L * 4));
    CK(cudaMalloc(&c->lv.active, L * 4));
    CK(cudaMalloc(&c->lv.found, L * 4));
    CK(cudaMalloc(&c->lv.iter, L * 4));
    CK(cudaMalloc(&c->lv.nleft_g, L * 4));
    CK(cudaMalloc(&c->lv.nleft_l, L * 4));
    CK(cudaMalloc(&c->lv.cuts, L * orb::kCS * 4));
    CK(cudaMalloc(&c->lv.cnt_l, L * orb::kCS * 4));
    c->lv.cnt_g = c->lv.cnt_l;
    CK(cudaMalloc(&c->lv.compL, L * 4));
    CK(cudaMalloc(&c->lv.compR, L * 4));
    CK(cudaMalloc(&c->lv.base_l, L * 4));
    CK(cudaMemset(c->lv.base_l, 0, L * 4));
    {
        const size_t nCountTiles = ceil_div(n_local, orb::kCountTile) + 1;
        CK(cudaMalloc(&c->lv.tile_ncand, nCountTiles * orb::kWarps * 4));
        CK(cudaMemset(c->lv.tile_ncand, 0, nCountTiles * orb::kWarps * 4));
    }
    c->lvAlt = c->lv;                 // search scratch (cuts, counters, compaction state) is shared
    c->lvAlt.bnd = nullptr; c->lvAlt.axis = nullptr; c->lvAlt.mL = c->lvAlt.mR = nullptr; c->lvAlt.total = nullptr;
    c->lvAlt.nleaf = nullptr; c->lvAlt.active = c->lvAlt.found = nullptr; c->lvAlt.iter = nullptr;
    c->lvAlt.nleft_g = c->lvAlt.nleft_l = nullptr;      // (own arrays: allocated below; orb_destroy frees exactly these)
    CK(cudaMalloc(&c->lvAlt.bnd, (L + 1) * 4));
    CK(cudaMalloc(&c->lvAlt.axis, L * 4));
    CK(cudaMalloc(&c->lvAlt.mL, L * 4));
    CK(cudaMalloc(&c->lvAlt.mR, L * 4));
    CK(cudaMalloc(&c->lvAlt.total, L * 4));
    CK(cudaMalloc(&c->lvAlt.nleaf, L * 4));
    CK(cudaMalloc(&c->lvAlt.active, L * 4));
    CK(cudaMalloc(&c->lvAlt.found, L * 4));
    CK(cudaMalloc(&c->lvAlt.iter, L * 4));
    CK(cudaMalloc(&c->lvAlt.nleft_g, L * 4));
    CK(cudaMalloc(&c->lvAlt.nleft_l, L * 4));
    CK(cudaMalloc(&c->d_final_cut, L * 4));
    const size_t nMap = ceil_div(n_local, orb::kMapTile) + 1;
    CK(cudaMalloc(&c->d_tile_first, nMap * 4));
    CK(cudaMalloc(&c->d_tile_first_alt, nMap * 4));
    CK(cudaMalloc(&c->d_blk_left, sizeof(uint32_t) * 64 * (size_t)c->nSM));
    CK(cudaMalloc(&c->d_blk_restart, sizeof(uint32_t) * 64 * (size_t)c->nSM));
    CK(cudaMalloc(&c->d_blk_le, sizeof(uint32_t) * 64 * (size_t)c->nSM));
    {
        orb::SelPar &pr = c->par;
        CK(cudaMalloc(&pr.fine, sizeof(uint32_t) * kParMaxCells * orb::kSelBins2));
        CK(cudaMalloc(&pr.amb, sizeof(float) * kParMaxCells * orb::kSelAmbCap));
        CK(cudaMalloc(&pr.lo2, 4 * kParMaxCells)); CK(cudaMalloc(&pr.sc2, 4 * kParMaxCells));
        CK(cudaMalloc(&pr.f2, 4 * kParMaxCells)); CK(cudaMalloc(&pr.l2, 4 * kParMaxCells));
        CK(cudaMalloc(&pr.base2, 4 * kParMaxCells)); CK(cudaMalloc(&pr.k2, 4 * kParMaxCells));
        CK(cudaMalloc(&pr.v2lo, 4 * kParMaxCells)); CK(cudaMalloc(&pr.v2hi, 4 * kParMaxCells));
        CK(cudaMalloc(&pr.ambCnt, 4 * kParMaxCells));
        CK(cudaMemset(pr.ambCnt, 0, 4 * kParMaxCells));
    }
    CK(cudaMalloc(&c->d_visits, sizeof(orb::SelVisitRec) * kVisitRecs));
    CK(cudaMemset(c->d_visits, 0, sizeof(orb::SelVisitRec) * kVisitRecs));
    CK(cudaMalloc(&c->d_pre, sizeof(orb::PreLeft) * 64 * (size_t)c->nSM));
    CK(cudaMemset(c->d_pre, 0, sizeof(orb::PreLeft) * 64 * (size_t)c->nSM));
    CK(cudaMalloc(&c->d_nGE, (size_t)std::max<uint32_t>(1u, n_leaf_cells) * 4));
    CK(cudaMalloc(&c->d_nLE, (size_t)std::max<uint32_t>(1u, n_leaf_cells) * 4));
    CK(cudaMalloc(&c->d_tickets, sizeof(uint32_t) * kMaxLevels));
    CK(cudaMalloc(&c->d_nactive, sizeof(uint32_t) * kMaxLevels * kPassSlots));
    CK(cudaMalloc(&c->d_done, sizeof(uint32_t) * kMaxLevels * kPassSlots));
    CK(cudaMalloc(&c->d_cdone, sizeof(uint32_t) * kMaxLevels * kPassSlots));
    CK(cudaMalloc(&c->d_peer_cnt, sizeof(uint32_t) * 2 * orb::kMaxPeers * orb::kPeerMaxCells * orb::kCS));
    CK(cudaMemset(c->d_peer_cnt, 0, sizeof(uint32_t) * 2 * orb::kMaxPeers * orb::kPeerMaxCells * orb::kCS));
    CK(cudaMalloc(&c->d_peer_flag, sizeof(uint32_t) * 2 * orb::kMaxPeers * orb::kPeerMaxBlocks));
    CK(cudaMemset(c->d_peer_flag, 0, sizeof(uint32_t) * 2 * orb::kMaxPeers * orb::kPeerMaxBlocks));
    CK(cudaMalloc(&c->d_lvl_passes, sizeof(int32_t) * kMaxLevels));
    CK(cudaMalloc(&c->d_lvl_unfound, sizeof(uint32_t) * kMaxLevels));
    {
        c->selHistWords = (size_t)n_local / 16 + 2 * (size_t)orb::kSelBinsMax;
        {   // exchange arena: one allocation so that one IPC handle maps all of it
            int rcA = alloc_arena(c, false);
            if (rcA) return rcA;
        }
        CK(cudaMalloc(&c->sel.bfirst, L * 4));
        CK(cudaMalloc(&c->sel.blast, L * 4));
        CK(cudaMalloc(&c->sel.base, L * 4));
        CK(cudaMalloc(&c->sel.ncand, L * 4));
        CK(cudaMalloc(&c->sel.flag, L * 4));
        CK(cudaMemset(c->sel.flag, 0, L * 4));
        CK(cudaMalloc(&c->sel.vlo, L * 4));
        CK(cudaMalloc(&c->sel.vhi, L * 4));
        CK(cudaMalloc(&c->d_sel_nflag, sizeof(uint32_t) * kMaxLevels));
        CK(cudaMemset(c->d_sel_nflag, 0, sizeof(uint32_t) * kMaxLevels));
        const int ringBytes = orb::kCountStages * orb::kCountTile * (int)sizeof(float);
        const int histBytes = ringBytes + orb::kSelBinsMax * 4 /* >= nb1 * rep * 4 for every level */, compBytes = ringBytes + orb::kWarps * orb::kSelWarpStage * 4 + orb::kSelBinsMax * 4;
        CK(cudaFuncSetAttribute(orb::k_sel_stream<orb::kSelHist>, cudaFuncAttributeMaxDynamicSharedMemorySize, histBytes));
        CK(cudaFuncSetAttribute(orb::k_sel_stream<orb::kSelCompact>, cudaFuncAttributeMaxDynamicSharedMemorySize, compBytes));
        CK(cudaFuncSetAttribute(orb::k_sel_stream<orb::kSelCompact, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, compBytes));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c->occSelStream[0], orb::k_sel_stream<orb::kSelHist>, orb::kThreads, histBytes));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c->occSelStream[1], orb::k_sel_stream<orb::kSelCompact>, orb::kThreads, compBytes));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c->occSelStream[2], orb::k_sel_stream<orb::kSelCompact, true>, orb::kThreads,
                                                         ringBytes + orb::kWarps * orb::kSelWarpStage * 4));
        CK(cudaFuncSetAttribute(orb::k_sel_resolve, cudaFuncAttributeMaxDynamicSharedMemorySize, orb::kSelBinsMax * 4));
        const int searchBytes = (int)orb::sel_search_smem_bytes(kSelValsCap);
        CK(cudaFuncSetAttribute(orb::k_sel_finish<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, searchBytes));
        CK(cudaFuncSetAttribute(orb::k_sel_finish<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, searchBytes));
        CK(cudaFuncSetAttribute(orb::k_selmr_finish<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, searchBytes));
        CK(cudaFuncSetAttribute(orb::k_selmr_finish<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, searchBytes));
        CK(cudaFuncSetAttribute(orb::k_sel_percell<512, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)orb::sel_percell_smem_bytes(kSelCellCapMax)));
        CK(cudaFuncSetAttribute(orb::k_sel_percell<256, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)orb::sel_percell_smem_bytes(kSelCellCapMax)));
        CK(cudaFuncSetAttribute(orb::k_sel_percell<256, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)orb::sel_percell_smem_bytes(kSelCellCapMax)));
        CK(cudaFuncSetAttribute(orb::k_sel_percell<1024, 1, 8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)orb::sel_percell_smem_bytes(kSelCellCapMax)));
        CK(cudaFuncSetAttribute(orb::k_sel_percell<512, 2, 8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)orb::sel_percell_smem_bytes(kSelCellCapMax)));
        CK(cudaFuncSetAttribute(orb::k_sel_percell<512, 3, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)orb::sel_percell_smem_bytes(kSelCellCapMax)));
        CK(cudaFuncSetAttribute(orb::k_sel_percell<256, 6, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)orb::sel_percell_smem_bytes(kSelCellCapMax)));
        CK(cudaFuncSetAttribute(orb::k_sel_percell<512, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)orb::sel_percell_smem_bytes(kSelCellCapMax)));
        CK(cudaFuncSetAttribute(orb::k_sel_percell<512, 2, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)orb::sel_percell_smem_bytes(kSelCellCapMax)));
        CK(cudaFuncSetAttribute(orb::k_sel_percell<1024, 1, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)orb::sel_percell_smem_bytes(kSelCellCapMax)));
    }
    CK(cudaFuncSetAttribute(orb::k_xf_finish_block, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)orb::sel_search_smem_bytes(kSelValsCap)));
    CK(cudaFuncSetAttribute(orb::k_xd_compact_warp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)orb::xd_compact_warp_smem(512)));
    CK(cudaFuncSetAttribute(orb::k_xf_finish_warp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(orb::kWarps * (orb::kXWarpCap + 4u) * 4u)));
    CK(cudaMalloc(&c->d_misc, 64));
    CK(cudaMalloc(&c->d_active_particles, 16));
    CK(cudaMalloc(&c->d_level_iters, sizeof(int32_t) * kMaxLevels));
    CK(cudaMalloc(&c->d_err, sizeof(int)));
    CK(cudaMemset(c->d_err, 0, sizeof(int)));
    CK(cudaMalloc(&c->d_bb, L * 2 * 8 * 4));
    CK(cudaMalloc(&c->d_bb6, L * 2 * 6 * 4));
    CK(cudaHostAlloc((void **)&c->h_status, sizeof(uint32_t) * kMaxLevels * kPassSlots, cudaHostAllocMapped));
    memset((void *)c->h_status, 0, sizeof(uint32_t) * kMaxLevels * kPassSlots);
    CK(cudaHostGetDevicePointer((void **)&c->h_status_dev, (void *)c->h_status, 0));
    {
        const int ringBytes = orb::kCountStages * orb::kCountTile * (int)sizeof(float);
        CK(cudaFuncSetAttribute(orb::k_count_stream<1, orb::kCountFull>, cudaFuncAttributeMaxDynamicSharedMemorySize, ringBytes));
        CK(cudaFuncSetAttribute(orb::k_count_stream<3, orb::kCountFull>, cudaFuncAttributeMaxDynamicSharedMemorySize, ringBytes));
        CK(cudaFuncSetAttribute(orb::k_count_stream<7, orb::kCountFull>, cudaFuncAttributeMaxDynamicSharedMemorySize, ringBytes));
        CK(cudaFuncSetAttribute(orb::k_count_stream<7, orb::kCountCompact>, cudaFuncAttributeMaxDynamicSharedMemorySize, ringBytes));
        CK(cudaFuncSetAttribute(orb::k_count_stream<7, orb::kCountCand>, cudaFuncAttributeMaxDynamicSharedMemorySize, ringBytes));
    }
    {
        const int ringBytes = orb::kCountStages * orb::kCountTile * (int)sizeof(float);
        CK(cudaFuncSetAttribute(orb::k_level_persistent<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, ringBytes));
        CK(cudaFuncSetAttribute(orb::k_level_persistent<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, ringBytes));
        CK(cudaFuncSetAttribute(orb::k_level_persistent<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, ringBytes));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c->occPersist[1], orb::k_level_persistent<1>, orb::kThreads, ringBytes));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c->occPersist[2], orb::k_level_persistent<2>, orb::kThreads, ringBytes));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c->occPersist[3], orb::k_level_persistent<3>, orb::kThreads, ringBytes));
    }
    CK(cudaFuncSetAttribute(orb::k_partition_coop, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(orb::PartSmem)));
    CK(cudaFuncSetAttribute(orb::k_partition_cells, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(orb::PartSmem)));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c->occPartStream, orb::k_partition_coop, orb::kThreads, sizeof(orb::PartSmem)));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c->occPartCells, orb::k_partition_cells, orb::kThreads, sizeof(orb::PartSmem)));
    if (c->occPartStream < 1 || c->occPartCells < 1) return fail(ORB_ERR_CUDA, "partition kernels do not fit on this device");
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c->occHoare, orb::k_hoare_scan, orb::kThreads, 0));
    if (c->occHoare < 1) c->occHoare = 1;
    const char *sml = getenv("ORB_SAMPLE_MIN_LOCAL");
    if (sml && atoll(sml) >= 0) c->sampleMinLocal = (uint64_t)atoll(sml);
    const char *sma = getenv("ORB_SAMPLE_MAX_AVG");
    if (sma && atoll(sma) >= 1) { c->sampleMaxAvg = (uint64_t)atoll(sma); c->sampleMaxAvgDefault = false; }
    const char *pfn = getenv("ORB_PAR_FINISH");
    if (pfn) c->parFinish = atoi(pfn) != 0;
    const char *smc = getenv("ORB_SAMPLE_MIN_CELL");
    if (smc && atoll(smc) >= 4096) c->sampleMinCell = (uint32_t)atoll(smc);
    const char *cko = getenv("ORB_CHUNK_OCC");
    if (cko && atoi(cko) >= 1 && atoi(cko) <= 3) c->chunkOcc = atoi(cko);
    const char *slo = getenv("ORB_SELECT_LOW_OCC");
    if (slo) c->selLowOcc = atoi(slo) != 0;
    const char *sst = getenv("ORB_SAMPLE_STRIDE");
    if (sst && atoi(sst) >= 1 && atoi(sst) <= 64) c->sampleS = atoi(sst);
    const char *ssz = getenv("ORB_SAMPLE_Z");
    if (ssz && atof(ssz) >= 0.0) c->sampleZ = (float)atof(ssz);
    const char *tm = getenv("ORB_TIES");
    if (tm && std::string(tm) == "hoare") c->tieMode = 1;
    const char *p = getenv("ORB_PROFILE");
    c->profile = p && atoi(p) != 0;
    const char *td = getenv("ORB_TRIAL_DEPTH");
    if (td && atoi(td) >= 1 && atoi(td) <= 3) c->trialDepth = atoi(td);
    if (getenv("ORB_DEBUG_TIMES")) {
        CK(cudaMalloc(&c->d_dbg, sizeof(unsigned long long) * 64 * kMaxLevels));
        CK(cudaMemset(c->d_dbg, 0, sizeof(unsigned long long) * 64 * kMaxLevels));
        if (atoi(getenv("ORB_DEBUG_TIMES")) >= 2) {
            const size_t bytes = sizeof(unsigned long long) * (size_t)kMaxLevels * kDbgPasses * kDbgBlocks * 4;
            CK(cudaMalloc(&c->d_dbg_blocks, bytes));
            CK(cudaMemset(c->d_dbg_blocks, 0, bytes));
        }
    }
    const char *pbk = getenv("ORB_PART_BULK");
    if (pbk) c->partBulk = atoi(pbk) != 0 ? 1 : 0;
    const char *pwm = getenv("ORB_PART_WARP_MAX");
    if (pwm) c->partWarpMax = atoi(pwm);
    const char *plf = getenv("ORB_PRELEFT");
    if (plf) c->preLeft = atoi(plf) != 0;
    const char *pd = getenv("ORB_PDL");
    if (pd) c->pdl = atoi(pd) != 0;
    const char *spc = getenv("ORB_SELECT_PERCELL_MIN");
    if (spc && atoi(spc) >= 1) c->selPerCellMinCells = atoi(spc);
    const char *sbb = getenv("ORB_SELECT_BIG_BLOCKS");
    if (sbb) c->selBigBlocks = atoi(sbb) != 0;
    const char *sbf = getenv("ORB_SELECT_BIG_FINISH");
    if (sbf) c->selBigFinish = atoi(sbf) != 0;
    const char *sba = getenv("ORB_SELECT_BIN_AVG");
    if (sba && atoi(sba) >= 64) c->selBinAvg = atoi(sba);
    const char *st5 = getenv("ORB_SELECT_T512_MIN");
    if (st5 && atoi(st5) >= 1) c->selT512MinAvg = atoi(st5);
    const char *pfh = getenv("ORB_PREFUSE");
    if (pfh) c->prefuseHist = atoi(pfh) != 0 ? 1 : 0;
    const char *fnl = getenv("ORB_FUSE_NEXT");
    if (fnl) c->fuseNextLevel = atoi(fnl) != 0;
    const char *se = getenv("ORB_SELECT");
    if (se) c->select = atoi(se) != 0;
    const char *sem = getenv("ORB_SELECT_MR");
    if (sem) c->selectMr = atoi(sem) != 0;
    const char *mv1 = getenv("ORB_MR_V1");
    if (mv1) c->mrV2 = atoi(mv1) == 0;
    const char *xcc = getenv("ORB_X_CAND_CAP");
    if (xcc && atoi(xcc) >= 256 && (uint32_t)atoi(xcc) <= kSelValsCap) c->xCandCap = (uint32_t)atoi(xcc);
    const char *msf = getenv("ORB_MR_SELF");
    if (msf && atoi(msf) != 0 && c->mrV2) {
        // testing aid: the multi-rank protocol of orb_exchange.cuh with this rank as its only peer
        c->mrSelf = true;
        c->nLocalMin = c->nGlobal = c->nLocalMax = c->nLocal;
        int rcA = alloc_arena(c, true);
        if (rcA) return rcA;
        CK(cudaMalloc(&c->d_sel_locbase, (size_t)c->maxLevelCells * 4));
        c->peerX[0] = c->d_xchg;
        c->peerEnabled = true;
    }
    const char *smt = getenv("ORB_STREAM_MIN_TILES");
    if (smt && atoi(smt) >= 1) c->streamMinTiles = atoi(smt);
    const char *pe = getenv("ORB_PERSIST");
    if (pe) c->persist = atoi(pe) != 0;
    const char *cp = getenv("ORB_COMPACT");
    if (cp) c->compaction = atoi(cp) != 0;
    const char *fu = getenv("ORB_FUSE_UPDATE");
    if (fu) c->fuseUpdate = atoi(fu) != 0;
    const char *ra = getenv("ORB_RUN_AHEAD");
    if (ra && atoi(ra) >= 0 && atoi(ra) <= 8) c->runAhead = atoi(ra);
    CK(cudaDeviceSynchronize());
    *out = c;
    return ORB_OK;
}

int orb_destroy(orb_ctx *c) {
    if (!c) return ORB_OK;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->comm && c->ownComm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
    for (int b = 0; b < 2; ++b) { cudaFree(c->x[b]); cudaFree(c->y[b]); cudaFree(c->z[b]); }
    cudaFree(c->d_heap); cudaFree(c->d_cells); cudaFree(c->d_range); cudaFree(c->d_total);
    cudaFree(c->lv.bnd); cudaFree(c->lv.axis); cudaFree(c->lv.mL); cudaFree(c->lv.mR); cudaFree(c->lv.total);
    cudaFree(c->lv.nleaf); cudaFree(c->lv.active); cudaFree(c->lv.found); cudaFree(c->lv.iter);
    cudaFree(c->lv.nleft_g); cudaFree(c->lv.nleft_l); cudaFree(c->lv.cuts); cudaFree(c->lv.cnt_l);
    cudaFree(c->lv.compL); cudaFree(c->lv.compR); cudaFree(c->lv.base_l); cudaFree(c->lv.tile_ncand);
    cudaFree(c->lvAlt.bnd); cudaFree(c->lvAlt.axis); cudaFree(c->lvAlt.mL); cudaFree(c->lvAlt.mR); cudaFree(c->lvAlt.total);
    cudaFree(c->lvAlt.nleaf); cudaFree(c->lvAlt.active); cudaFree(c->lvAlt.found); cudaFree(c->lvAlt.iter);
    cudaFree(c->lvAlt.nleft_g); cudaFree(c->lvAlt.nleft_l); cudaFree(c->d_tile_first_alt);
    if (c->d_cnt_g_buf) cudaFree(c->d_cnt_g_buf);
    cudaFree(c->d_dbg); cudaFree(c->d_dbg_blocks);
    cudaFree(c->sel.bfirst); cudaFree(c->sel.blast); cudaFree(c->sel.base); cudaFree(c->sel.ncand);
    cudaFree(c->sel.flag); cudaFree(c->d_sel_nflag); cudaFree(c->d_visits); cudaFree(c->sel.vlo); cudaFree(c->sel.vhi);
    cudaFree(c->par.fine); cudaFree(c->par.amb); cudaFree(c->par.lo2); cudaFree(c->par.sc2); cudaFree(c->par.f2); cudaFree(c->par.l2);
    cudaFree(c->par.base2); cudaFree(c->par.k2); cudaFree(c->par.v2lo); cudaFree(c->par.v2hi); cudaFree(c->par.ambCnt);
    cudaFree(c->d_sel_hist_g); cudaFree(c->d_sel_locbase); cudaFree(c->d_slots_g);
    for (int r = 0; r < orb::kMaxPeers; ++r)
        if (c->peerIpc[r] && c->peerX[r]) cudaIpcCloseMemHandle(c->peerX[r]);
    cudaFree(c->d_xchg);
    cudaFree(c->d_xscratch);
    for (int r = 0; r < orb::kMaxPeers; ++r)
        if (c->peerIpc[r]) { cudaIpcCloseMemHandle(c->peerCnt[r]); cudaIpcCloseMemHandle(c->peerFlag[r]); }
    cudaFree(c->d_lvl_passes); cudaFree(c->d_lvl_unfound); cudaFree(c->d_cdone); cudaFree(c->d_peer_cnt); cudaFree(c->d_peer_flag);
    cudaFree(c->d_final_cut); cudaFree(c->d_tile_first); cudaFree(c->d_blk_left); cudaFree(c->d_blk_restart); cudaFree(c->d_blk_le); cudaFree(c->d_pre); cudaFree(c->d_nGE); cudaFree(c->d_nLE); cudaFree(c->d_tickets);
    cudaFree(c->d_nactive); cudaFree(c->d_done); cudaFree(c->d_misc); cudaFree(c->d_active_particles);
    cudaFree(c->d_level_iters); cudaFree(c->d_err); cudaFree(c->d_bb); cudaFree(c->d_bb6);
    if (c->h_status) cudaFreeHost((void *)c->h_status);
    if (c->h_scratch) cudaFreeHost(c->h_scratch);
    for (auto &e : c->evCount) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
    for (auto &e : c->evPart) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
    for (auto &e : c->evAux) { if (e.e0) cudaEventDestroy(e.e0); if (e.e1) cudaEventDestroy(e.e1); }
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return ORB_OK;
}

int orb_set_trial_depth(orb_ctx *c, int m) {
    if (!c) return fail(ORB_ERR_ARG, "null ctx");
    if (m == 0) m = 3;
    if (m < 1 || m > 3) return fail(ORB_ERR_ARG, "trial depth must be 1..3");
    c->trialDepth = m;
    return ORB_OK;
}

int orb_set_tie_mode(orb_ctx *c, int mode) {
    if (!c) return fail(ORB_ERR_ARG, "null ctx");
    if (mode != 0 && mode != 1) return fail(ORB_ERR_ARG, "tie mode must be 0 (canonical) or 1 (hoare)");
    c->tieMode = mode;
    return ORB_OK;
}

int orb_plan_level(uint64_t n_local, uint64_t n_global, uint64_t n_local_min, int n_ranks, uint32_t n_leaf_cells,
                   uint32_t n_cells, int prefuse_mode, orb_level_plan *out) {
    if (!out || n_cells == 0 || n_ranks < 1) return fail(ORB_ERR_ARG, "bad plan arguments");
    // a context that never touches a device: only the fields the planning functions read
    std::unique_ptr<orb_ctx> c(new orb_ctx());
    c->nLocal = n_local;
    c->nGlobal = n_ranks > 1 ? n_global : n_local;
    c->nLocalMin = n_ranks > 1 ? n_local_min : n_local;
    c->nRanks = n_ranks;
    c->d = n_leaf_cells;
    c->maxLevelCells = std::max<uint32_t>(1u, n_leaf_cells);
    c->selHistWords = (size_t)n_local / 16 + 2 * (size_t)orb::kSelBinsMax;
    c->occPersist[3] = 1;
    c->prefuseHist = prefuse_mode < 0 ? -1 : (prefuse_mode ? 1 : 0);
    c->d_slots_g = reinterpret_cast<float *>(uintptr_t(16));     // "allocated" (sel_plan_mr only tests the pointer)
    c->d_visits = reinterpret_cast<orb::SelVisitRec *>(uintptr_t(16));      // (likewise sampling_on)
    const char *sst = getenv("ORB_SAMPLE_STRIDE");
    if (sst && atoi(sst) >= 1 && atoi(sst) <= 64) c->sampleS = atoi(sst);
    c->nLocalMax = c->nLocalMin;      // (the plan never reads it: buffers are sized from it, decisions use the minimum)
    c->peerEnabled = n_ranks > 1;     // plan of the peer-memory protocol (orb_exchange.cuh)
    c->slotTotal = std::max<size_t>(std::max<size_t>(kSelSlotWordsTotal, 64 * (size_t)c->maxLevelCells), (((size_t)n_local_min / 8) + 63) & ~(size_t)63);
    const char *mv1 = getenv("ORB_MR_V1");
    if (mv1) c->mrV2 = atoi(mv1) == 0;
    memset(out, 0, sizeof(*out));
    const int M = 3;
    const int pre = n_cells >= 2 ? prefuse_nb(c.get(), n_cells, M) : 0;
    out->prefuse_bins = pre;
    const SelMrPlan mr = sel_plan_mr(c.get(), n_cells, M, pre);
    if (mr.ok) {
        out->search = !mr.v2 ? 3 : (mr.regime == 0 ? 3 : (mr.regime == 1 ? 4 : 5));
        out->hist_bins = mr.nb1;
        out->cand_cap = mr.candCap;
        out->slot_words = mr.slotWords;
        out->hist_words = mr.histWords;
    } else if (level_can_select(c.get(), n_cells, M)) {
        const SelPlan p = sel_plan(c.get(), n_cells, pre);
        out->sample_stride = (p.sampleS > 1 && (!p.cellsInSmem || n_local / n_cells >= (uint64_t)c->sampleMinCell)) ? p.sampleS : 1;
        out->search = p.cellsInSmem ? 2 : 1;
        out->hist_bins = (p.cellsInSmem && !pre) ? 0 : p.nb1;
        out->cand_cap = p.cellsInSmem ? p.cellCap : p.candCap;
        out->hist_words = p.histWords;
    }
    c->d_slots_g = nullptr;
    c->d_visits = nullptr;
    return ORB_OK;
}

int orb_set_profile(orb_ctx *c, int on) {
    if (!c) return fail(ORB_ERR_ARG, "null ctx");
    c->profile = on != 0;
    return ORB_OK;
}

// ---------------------------------------------------------------- multi-GPU
int orb_comm_unique_id(void *id128) {
    if (!id128) return fail(ORB_ERR_ARG, "null id buffer");
    if (!g_nccl.load()) return fail(ORB_ERR_NCCL, "libnccl.so.2 not found: %s", dlerror());
    ncclUniqueId id;
    NK(g_nccl.GetUniqueId(&id));
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    memcpy(id128, &id, 128);
    return ORB_OK;
}

static int setup_multi(orb_ctx *c) {
    if (c->nRanks > 1 && !c->d_cnt_g_buf) {
        CK(cudaMalloc(&c->d_cnt_g_buf, (size_t)c->maxLevelCells * orb::kCS * 4));
        c->lv.cnt_g = c->d_cnt_g_buf;
        c->lvAlt.cnt_g = c->d_cnt_g_buf;
    }
    c->nLocalMin = c->nGlobal = c->nLocalMax = c->nLocal;
    if (c->nRanks > 1) {
        // shard sizes over ranks (min, sum, max): whatever shapes a collective or an exchange must be decided from
        // rank-invariant numbers
        unsigned long long h[3] = {~(unsigned long long)c->nLocal, (unsigned long long)c->nLocal, (unsigned long long)c->nLocal}, *d = nullptr;
        CK(cudaMalloc(&d, 24));
        CK(cudaMemcpyAsync(d, h, 24, cudaMemcpyHostToDevice, c->stream));
        NK(g_nccl.AllReduce(d, d, 1, ncclUint64, ncclMax, c->comm, c->stream));          // max of ~n = ~min
        NK(g_nccl.AllReduce(d + 1, d + 1, 1, ncclUint64, ncclSum, c->comm, c->stream));
        NK(g_nccl.AllReduce(d + 2, d + 2, 1, ncclUint64, ncclMax, c->comm, c->stream));
        CK(cudaMemcpyAsync(h, d, 24, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        CK(cudaFree(d));
        c->nLocalMin = ~h[0];
        c->nGlobal = h[1];
        c->nLocalMax = h[2];
        if (c->mrV2) {     // the arena of the multi-rank protocol (must precede orb_peer_export)
            int rcA = alloc_arena(c, true);
            if (rcA) return rcA;
        }
        if (!c->d_sel_hist_g) {
            CK(cudaMalloc(&c->d_sel_hist_g, c->selHistWords * 4));
            CK(cudaMalloc(&c->d_sel_locbase, (size_t)c->maxLevelCells * 4));
            CK(cudaMalloc(&c->d_slots_g, kSelSlotWordsTotal * 4 * (size_t)c->nRanks));
        }
    }
    return ORB_OK;
}

int orb_comm_init(orb_ctx *c, const void *id128, int rank, int n_ranks) {
    if (!c || !id128 || n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(ORB_ERR_ARG, "bad communicator arguments");
    if (!g_nccl.load()) return fail(ORB_ERR_NCCL, "libnccl.so.2 not found: %s", dlerror());
    CK(cudaSetDevice(c->device));
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    NK(g_nccl.CommInitRank(&c->comm, n_ranks, id, rank));
    c->ownComm = true;
    c->rank = rank;
    c->nRanks = n_ranks;
    return setup_multi(c);
}

int orb_comm_attach(orb_ctx *c, void *nccl_comm, int rank, int n_ranks) {
    if (!c || !nccl_comm || n_ranks < 1) return fail(ORB_ERR_ARG, "bad communicator arguments");
    if (!g_nccl.load()) return fail(ORB_ERR_NCCL, "libnccl.so.2 not found: %s", dlerror());
    CK(cudaSetDevice(c->device));
    c->comm = (ncclComm_t)nccl_comm;
    c->ownComm = false;
    c->rank = rank;
    c->nRanks = n_ranks;
    return setup_multi(c);
}

// Peer memory for the fused count+combine.  Every rank exports a descriptor of its counter rows and flags, the caller
// gathers the descriptors of all ranks (any transport) and hands the table to every rank.
int orb_peer_export(orb_ctx *c, orb_peer_info *out) {
    if (!c || !out) return fail(ORB_ERR_ARG, "null argument");
    CK(cudaSetDevice(c->device));
    memset(out, 0, sizeof(*out));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, c->d_peer_cnt));
    memcpy(out->ipc_cnt, &h, 64);
    CK(cudaIpcGetMemHandle(&h, c->d_peer_flag));
    memcpy(out->ipc_flag, &h, 64);
    CK(cudaIpcGetMemHandle(&h, c->d_xchg));
    memcpy(out->ipc_xchg, &h, 64);
    out->ptr_cnt = (uint64_t)(uintptr_t)c->d_peer_cnt;
    out->ptr_flag = (uint64_t)(uintptr_t)c->d_peer_flag;
    out->ptr_xchg = (uint64_t)(uintptr_t)c->d_xchg;
    out->pid = (int64_t)getpid();
    out->device = c->device;
    return ORB_OK;
}

int orb_peer_import(orb_ctx *c, const orb_peer_info *all, int n_ranks) {
    if (!c || !all) return fail(ORB_ERR_ARG, "null argument");
    if (n_ranks != c->nRanks || n_ranks > orb::kMaxPeers) return fail(ORB_ERR_ARG, "peer table must have one entry per rank (<= %d)", orb::kMaxPeers);
    CK(cudaSetDevice(c->device));
    for (int r = 0; r < n_ranks; ++r) {
        if (r == c->rank) {
            c->peerCnt[r] = c->d_peer_cnt;
            c->peerFlag[r] = c->d_peer_flag;
            c->peerX[r] = c->d_xchg;
        } else if (all[r].pid == (int64_t)getpid()) {
            // same process (thread-per-GPU host): plain pointers + peer access
            int can = 0;
            CK(cudaDeviceCanAccessPeer(&can, c->device, all[r].device));
            if (!can) return fail(ORB_ERR_CUDA, "device %d cannot access peer %d", c->device, all[r].device);
            cudaError_t e = cudaDeviceEnablePeerAccess(all[r].device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CK(e);
            cudaGetLastError();
            c->peerCnt[r] = (uint32_t *)(uintptr_t)all[r].ptr_cnt;
            c->peerFlag[r] = (uint32_t *)(uintptr_t)all[r].ptr_flag;
            c->peerX[r] = (uint32_t *)(uintptr_t)all[r].ptr_xchg;
        } else {
            cudaIpcMemHandle_t h;
            void *p = nullptr;
            memcpy(&h, all[r].ipc_cnt, 64);
            CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
            c->peerCnt[r] = (uint32_t *)p;
            memcpy(&h, all[r].ipc_flag, 64);
            CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
            c->peerFlag[r] = (uint32_t *)p;
            memcpy(&h, all[r].ipc_xchg, 64);
            CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
            c->peerX[r] = (uint32_t *)p;
            c->peerIpc[r] = true;
        }
    }
    c->peerEnabled = true;
    return ORB_OK;
}

// ---------------------------------------------------------------- particles
static int set_root_range(orb_ctx *c) {
    uint32_t r[2] = {0u, (uint32_t)c->nLocal};
    CK(cudaMemcpyAsync(c->d_range, r, 8, cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->haveParticles = true;
    return ORB_OK;
}

int orb_upload_xyz(orb_ctx *c, const float *x, const float *y, const float *z) {
    if (!c || !x || !y || !z) return fail(ORB_ERR_ARG, "null argument");
    CK(cudaSetDevice(c->device));
    c->cur = 0;
    CK(cudaMemcpyAsync(c->x[0], x, c->nLocal * 4, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->y[0], y, c->nLocal * 4, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->z[0], z, c->nLocal * 4, cudaMemcpyHostToDevice, c->stream));
    return set_root_range(c);
}

int orb_load_device_xyz(orb_ctx *c, const float *dx, const float *dy, const float *dz) {
    if (!c || !dx || !dy || !dz) return fail(ORB_ERR_ARG, "null argument");
    CK(cudaSetDevice(c->device));
    c->cur = 0;
    CK(cudaMemcpyAsync(c->x[0], dx, c->nLocal * 4, cudaMemcpyDeviceToDevice, c->stream));
    CK(cudaMemcpyAsync(c->y[0], dy, c->nLocal * 4, cudaMemcpyDeviceToDevice, c->stream));
    CK(cudaMemcpyAsync(c->z[0], dz, c->nLocal * 4, cudaMemcpyDeviceToDevice, c->stream));
    return set_root_range(c);
}

int orb_download_xyz(orb_ctx *c, float *x, float *y, float *z) {
    if (!c || !x || !y || !z) return fail(ORB_ERR_ARG, "null argument");
    if (!c->haveParticles) return fail(ORB_ERR_STATE, "no particles uploaded");
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpyAsync(x, c->x[c->cur], c->nLocal * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(y, c->y[c->cur], c->nLocal * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(z, c->z[c->cur], c->nLocal * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return ORB_OK;
}

int orb_device_xyz(orb_ctx *c, const float **dx, const float **dy, const float **dz) {
    if (!c) return fail(ORB_ERR_ARG, "null ctx");
    if (dx) *dx = c->x[c->cur];
    if (dy) *dy = c->y[c->cur];
    if (dz) *dz = c->z[c->cur];
    return ORB_OK;
}

int orb_get_ranges(orb_ctx *c, uint32_t first_id, uint32_t n, uint32_t *out) {
    if (!c || !out || (uint64_t)first_id + n > c->nHeap) return fail(ORB_ERR_ARG, "range read outside heap");
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpyAsync(out, c->d_range + 2 * (size_t)first_id, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return ORB_OK;
}

// ---------------------------------------------------------------- service-granular calls
// particles per cell of the level (already uploaded to d_cells), summed over ranks, into lv.cnt_g[0..n) and into
// d_total[id]: the device side of ServiceCount (count.cpp:8-30)
static int level_totals(orb_ctx *c, uint32_t n_cells) {
    const uint32_t blocks = ceil_div(n_cells, 256);
    orb::k_cell_sizes<<<blocks, 256, 0, c->stream>>>(c->d_cells, n_cells, c->d_range, c->lv.cnt_l);
    if (c->nRanks > 1) NK(g_nccl.AllReduce(c->lv.cnt_l, c->lv.cnt_g, n_cells, ncclUint32, ncclSum, c->comm, c->stream));
    orb::k_store_totals<<<blocks, 256, 0, c->stream>>>(c->d_cells, n_cells, c->lv.cnt_g, c->d_total);
    c->nOtherLaunch += 2;
    CK(cudaGetLastError());
    return ORB_OK;
}

int orb_count(orb_ctx *c, const orb_cell *cells, uint32_t n_cells, uint32_t *out) {
    if (!c || !out) return fail(ORB_ERR_ARG, "null argument");
    if (!c->haveParticles) return fail(ORB_ERR_STATE, "no particles uploaded");
    CK(cudaSetDevice(c->device));
    int rc = upload_cells(c, cells, n_cells);
    if (rc) return rc;
    // local sizes come straight from the range map (count.cpp:16), summed over ranks like Combine (count.cpp:21-30);
    // only the level's n_cells words travel back
    rc = level_totals(c, n_cells);
    if (rc) return rc;
    CK(cudaMemcpyAsync(out, c->lv.cnt_g, (size_t)n_cells * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return ORB_OK;
}

int orb_count_left(orb_ctx *c, const orb_cell *cells, uint32_t n_cells, uint32_t *out) {
    if (!c || !out) return fail(ORB_ERR_ARG, "null argument");
    if (!c->haveParticles) return fail(ORB_ERR_STATE, "no particles uploaded");
    CK(cudaSetDevice(c->device));
    int rc = upload_cells(c, cells, n_cells);
    if (rc) return rc;
    rc = level_prepare(c, c->d_cells, n_cells, 1, nullptr);   // cuts[c] = getCut(c) (cell.h:74-76)
    if (rc) return rc;
    rc = launch_count(c, n_cells, 1, nullptr);
    if (rc) return rc;
    rc = allreduce_counts(c, n_cells, 1);
    if (rc) return rc;
    rc = ensure_scratch(c, (size_t)n_cells * orb::kCS * 4);
    if (rc) return rc;
    CK(cudaMemcpyAsync(c->h_scratch, c->lv.cnt_g, (size_t)n_cells * orb::kCS * 4, cudaMemcpyDeviceToHost, c->stream));
    rc = check_device_err(c);
    if (rc) return rc;
    for (uint32_t i = 0; i < n_cells; ++i)
        if (!cells[i].foundCut) out[i] = c->h_scratch[(size_t)i * orb::kCS];   // found cells keep their old entry (countLeft.cpp:19-21)
    return ORB_OK;
}

int orb_partition(orb_ctx *c, const orb_cell *cells, uint32_t n_cells) {
    if (!c) return fail(ORB_ERR_ARG, "null ctx");
    if (!c->haveParticles) return fail(ORB_ERR_STATE, "no particles uploaded");
    CK(cudaSetDevice(c->device));
    int rc = upload_cells(c, cells, n_cells);
    if (rc) return rc;
    // the children's ranges are written at ids 2 id + 1 and 2 id + 2 (partition.cpp:54-60): leaves have none
    for (uint32_t i = 0; i < n_cells; ++i)
        if (2ull * (uint64_t)cells[i].id + 2ull >= (uint64_t)c->nHeap)
            return fail(ORB_ERR_ARG, "cell id %d has no children in a heap of %u cells", cells[i].id, c->nHeap);
    // count at the final cut of EVERY cell (found or not) to get the split offsets, then scatter
    std::vector<orb_cell> tmp(cells, cells + n_cells);
    for (auto &t : tmp) t.foundCut = 0;
    CK(cudaMemcpyAsync(c->d_cells, tmp.data(), (size_t)n_cells * sizeof(orb_cell), cudaMemcpyHostToDevice, c->stream));
    rc = level_prepare(c, c->d_cells, n_cells, 1, nullptr);
    if (rc) return rc;
    rc = launch_count(c, n_cells, 1, nullptr);
    if (rc) return rc;
    rc = allreduce_counts(c, n_cells, 1);
    if (rc) return rc;
    orb::k_finalize_apply<<<ceil_div(n_cells, 256), 256, 0, c->stream>>>(c->lv, n_cells);
    c->nOtherLaunch++;
    if (c->tieMode == 1) {
        rc = launch_partition_hoare(c, n_cells);
        if (rc) return rc;
    }
    orb::k_ranges_from_level<<<ceil_div(n_cells, 256), 256, 0, c->stream>>>(c->d_cells, n_cells, c->lv, c->d_range, c->d_final_cut);
    c->nOtherLaunch++;
    if (c->tieMode != 1) {
        CK(cudaMemsetAsync(c->d_tickets, 0, 4, c->stream));
        rc = launch_partition(c, n_cells, c->d_tickets);
        if (rc) return rc;
    }
    return check_device_err(c);
}

static int bbox_level(orb_ctx *c, const uint32_t *bnd_unused, uint32_t nCells, float *d_out6) {
    using namespace orb;
    (void)bnd_unused;
    k_bbox_init<<<ceil_div((uint64_t)nCells * 8, 256), 256, 0, c->stream>>>(c->d_bb, nCells);
    const uint32_t nTiles = ceil_div(c->nLocal, kMapTile);
    if (nTiles) {
        const uint32_t grid = std::min<uint32_t>(nTiles, (uint32_t)c->nSM * 8u);
        k_bbox<<<grid, kThreads, 0, c->stream>>>(c->x[c->cur], c->y[c->cur], c->z[c->cur], c->lv.bnd, c->d_tile_first, nCells,
                                                 (uint32_t)c->nLocal, nTiles, c->d_bb);
    }
    c->nOtherLaunch += 2;
    if (c->nRanks > 1) {
        // encoded uint32 boxes: unsigned min / max over ranks (one call each over the interleaved rows would mix
        // mins and maxes, so reduce twice and let the decode pick the right half)
        uint32_t *tmp = c->d_bb + (size_t)c->maxLevelCells * 8;   // second half of d_bb (allocated 2x)
        NK(g_nccl.AllReduce(c->d_bb, tmp, (size_t)nCells * 8, ncclUint32, ncclMax, c->comm, c->stream));
        NK(g_nccl.AllReduce(c->d_bb, c->d_bb, (size_t)nCells * 8, ncclUint32, ncclMin, c->comm, c->stream));
        // merge: mins from the Min result (in place), maxes from the Max result
        CK(cudaMemcpy2DAsync(c->d_bb + 3, 32, tmp + 3, 32, 12, nCells, cudaMemcpyDeviceToDevice, c->stream));
    }
    k_bbox_decode<<<ceil_div((uint64_t)nCells * 6, 256), 256, 0, c->stream>>>(c->d_bb, nCells, d_out6);
    c->nOtherLaunch++;
    CK(cudaGetLastError());
    return ORB_OK;
}

int orb_bbox(orb_ctx *c, const orb_cell *cells, uint32_t n_cells, float *out6) {
    if (!c || !out6) return fail(ORB_ERR_ARG, "null argument");
    if (!c->haveParticles) return fail(ORB_ERR_STATE, "no particles uploaded");
    CK(cudaSetDevice(c->device));
    if (n_cells > c->maxLevelCells) return fail(ORB_ERR_ARG, "too many cells");
    int rc = upload_cells(c, cells, n_cells);
    if (rc) return rc;
    rc = level_prepare(c, c->d_cells, n_cells, 1, nullptr);
    if (rc) return rc;
    rc = bbox_level(c, nullptr, n_cells, c->d_bb6);
    if (rc) return rc;
    CK(cudaMemcpyAsync(out6, c->d_bb6, (size_t)n_cells * 24, cudaMemcpyDeviceToHost, c->stream));
    return check_device_err(c);
}

// ---------------------------------------------------------------- fused level
int orb_find_cuts(orb_ctx *c, orb_cell *cells, uint32_t n_cells, int32_t *iters, int32_t *passes) {
    if (!c) return fail(ORB_ERR_ARG, "null ctx");
    if (!c->haveParticles) return fail(ORB_ERR_STATE, "no particles uploaded");
    CK(cudaSetDevice(c->device));
    int rc = upload_cells(c, cells, n_cells);
    if (rc) return rc;
    rc = reset_pass_ctl(c);
    if (rc) return rc;
    // the bisection's target is half of each cell's particle count over all ranks (orbit.cpp:204-205): derived here
    // from the range map, so the call does not depend on an earlier orb_count
    rc = level_totals(c, n_cells);
    if (rc) return rc;
    const int M = c->trialDepth;
    rc = level_prepare(c, c->d_cells, n_cells, (1 << M) - 1, c->d_nactive);
    if (rc) return rc;
    int np = 0;
    rc = run_bisection(c, n_cells, M, 0, 0, &np);
    if (rc) return rc;
    orb::k_writeback_cells<<<ceil_div(n_cells, 256), 256, 0, c->stream>>>(c->d_cells, c->lv, n_cells);
    c->nOtherLaunch++;
    CK(cudaMemcpyAsync(cells, c->d_cells, (size_t)n_cells * sizeof(orb_cell), cudaMemcpyDeviceToHost, c->stream));
    int32_t it = 0;
    CK(cudaMemcpyAsync(&it, c->d_level_iters, 4, cudaMemcpyDeviceToHost, c->stream));
    rc = check_device_err(c);
    if (rc) return rc;
    if (iters) *iters = it;
    if (passes) *passes = np;
    return ORB_OK;
}

// ---------------------------------------------------------------- whole build (orbit.cpp:74-275)
int orb_build(orb_ctx *c, uint32_t flags, orb_cell *heap_out, orb_build_stats *stats) {
    using namespace orb;
    if (!c) return fail(ORB_ERR_ARG, "null ctx");
    if (!c->haveParticles) return fail(ORB_ERR_STATE, "no particles uploaded");
    CK(cudaSetDevice(c->device));
    const bool tight = (flags & ORB_TIGHT_BOX) != 0;
    const int nLevelsRef = (int)std::ceil(std::log2((double)c->d));            // Cell::getNLevels (cell.h:65-67)
    const int lEnd = (flags & ORB_FULL_LEVELS) ? nLevelsRef + 1 : nLevelsRef;  // orbit.cpp:102
    if (lEnd - 1 > kMaxLevels) return fail(ORB_ERR_ARG, "too many levels");
    struct EventPair {      // destroyed on every return path
        cudaEvent_t a = nullptr, b = nullptr;
        ~EventPair() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
    } evp;
    CK(cudaEventCreate(&evp.a));
    CK(cudaEventCreate(&evp.b));
    const cudaEvent_t evA = evp.a, evB = evp.b;
    CK(cudaEventRecord(evA, c->stream));   // the build's clock starts before any of its bookkeeping
    int rc = reset_pass_ctl(c);
    if (rc) return rc;
    c->evCountUsed = c->evPartUsed = 0;
    c->evAuxUsed = 0;
    const uint64_t l0 = c->nCountLaunch, l1 = c->nUpdateLaunch, l2 = c->nPartLaunch, l3 = c->nOtherLaunch;

    // root cell: orbit.cpp:45-46,74-76
    orb_cell root;
    memset(&root, 0, sizeof(root));
    root.id = 0; root.nLeafCells = (int)c->d; root.prevCutAxis = -1; root.cutAxis = 0; root.foundCut = 0;
    for (int k = 0; k < 3; ++k) { root.lower[k] = -0.5f; root.upper[k] = 0.5f; }
    root.cutMarginLeft = root.lower[0]; root.cutMarginRight = root.upper[0];
    CK(cudaMemsetAsync(c->d_heap, 0, (size_t)c->nHeap * sizeof(orb_cell), c->stream));
    CK(cudaMemcpyAsync(c->d_heap, &root, sizeof(root), cudaMemcpyHostToDevice, c->stream));
    {   // global particle count of the root (ServiceCount at level 1, count.cpp:16,28)
        uint32_t n = (uint32_t)c->nLocal;
        CK(cudaMemcpyAsync(c->d_total, &n, 4, cudaMemcpyHostToDevice, c->stream));
        if (c->nRanks > 1) NK(g_nccl.AllReduce(c->d_total, c->d_total, 1, ncclUint32, ncclSum, c->comm, c->stream));
        uint32_t r[2] = {0u, n};
        CK(cudaMemcpyAsync(c->d_range, r, 8, cudaMemcpyHostToDevice, c->stream));
    }
    if (tight) {   // extension: root box = particle bounding box
        rc = level_prepare(c, c->d_heap, 1, 1, nullptr);
        if (rc) return rc;
        rc = bbox_level(c, nullptr, 1, c->d_bb6);
        if (rc) return rc;
        k_apply_bbox<<<1, 32, 0, c->stream>>>(c->d_heap, 1, c->d_bb6, c->d_total);
        c->nOtherLaunch++;
    }

    const int M = c->trialDepth;
    c->sampleOff = false;
    c->extraPasses.assign(kMaxLevels, 0);
    std::vector<int> passes;
    std::vector<uint32_t> unfound;
    int nDone = 0;
    bool prepared = false;      // this level's SoA state, tile map and cleared histogram rows came from the previous level's k_split
    int preNb = 0;              // != 0: the previous level's partition built this level's histogram rows with this many bins
    for (int l = 1; l < lEnd; ++l) {
        const uint32_t first = (1u << (l - 1)) - 1u;   // a = 2^(l-1)-1 (orbit.cpp:104); nCells = 2^(l-1) for d = 2^y
        const uint32_t nCells = 1u << (l - 1);
        const int slot = (l - 1) * kPassSlots;
        const bool useSelect = level_can_select(c, nCells, M);
        const SelMrPlan mrPlan = sel_plan_mr(c, nCells, M, preNb);
        const size_t histWords = mrPlan.ok ? mrPlan.zeroWords : (useSelect ? sel_plan(c, nCells, preNb).histWords : 0);
        if (!prepared) {
            rc = level_prepare(c, c->d_heap + first, nCells, (1 << M) - 1, c->d_nactive + slot, c->sel.hist, std::min(histWords, c->selHistWords));
            if (rc) return rc;
        }
        prepared = false;
        c->chunkTiles = 0;
        c->preValid = false;
        int np = 0;
        uint32_t nu = 0;
        bool speculate = false;     // split + partition enqueued behind the search before its flag count is known
        if (mrPlan.ok) {
            rc = mrPlan.v2 ? launch_level_select_mr2(c, nCells, mrPlan, slot, l - 1, preNb) : launch_level_select_mr(c, nCells, mrPlan, slot, l - 1, preNb);
            if (rc) return rc;
            np = -1;
            speculate = c->tieMode != 1;
            if (!speculate) {
                uint32_t nf = 0;
                if ((rc = select_mr_flagged(c, slot, &nf))) return rc;
                if (nf && (rc = select_mr_fallback(c, nCells, slot, l - 1))) return rc;
            }
        } else if (useSelect) {
            rc = launch_level_select(c, nCells, slot, l - 1, preNb);
            if (rc) return rc;
            np = -1;
            speculate = c->tieMode != 1;
            if (!speculate) {
                uint32_t nf = 0;
                if ((rc = select_mr_flagged(c, slot, &nf))) return rc;
                if (nf && (rc = select_fallback(c, nCells, slot, l - 1))) return rc;
            }
        } else if (level_can_persist(c, nCells, M)) {
            // host-free: the whole loop (and the extra count of capped cells) is one cooperative launch;
            // passes / unfound are read back with the other statistics after the build
            rc = launch_level_persistent(c, nCells, M, slot, l - 1);
            if (rc) return rc;
            np = -1;
        } else {
            rc = run_bisection(c, nCells, M, slot, l - 1, &np);
            if (rc) return rc;
            // only a level that used every pass can have unfound cells
            if (np >= (kMaxIter + M - 1) / M) {
                rc = finalize_unfound(c, nCells, &nu);
                if (rc) return rc;
            }
        }
        passes.push_back(np);
        unfound.push_back(nu);
        if (c->tieMode == 1) {   // reference-exact ties: partition first (it decides the child sizes), then split
            rc = launch_partition_hoare(c, nCells);
            if (rc) return rc;
        }
        // multi-rank selection search: split and partition go out gated on the level's flag count, the host looks at
        // the count afterwards; only if cells were flagged it runs the iterative loop and enqueues both again
        const uint32_t *gate = speculate ? c->d_sel_nflag + (l - 1) : nullptr;
        // the same launch prepares the next level (not in the modes that build the next level's state elsewhere)
        NextLevel nx;
        memset(&nx, 0, sizeof(nx));
        NextHist nh = no_next_hist();
        int preNext = 0;
        uint32_t splitBlocks = ceil_div(nCells, 256);
        if (c->fuseNextLevel && l + 1 < lEnd && !tight && c->tieMode != 1 && 2u * nCells <= c->maxLevelCells && c->nLocal > 0) {
            const uint32_t nNext = 2u * nCells;
            preNext = prefuse_nb(c, nNext, M);
            const SelMrPlan mrNext = sel_plan_mr(c, nNext, M, preNext);
            const size_t hwNext = mrNext.ok ? mrNext.zeroWords : (level_can_select(c, nNext, M) ? sel_plan(c, nNext, preNext).histWords : 0);
            if (preNext) {
                nh.enabled = 1;
                nh.hist = c->sel.hist;
                nh.nb = preNext;
                nh.mL = c->lvAlt.mL; nh.mR = c->lvAlt.mR; nh.axis = c->lvAlt.axis;
            }
            nx.enabled = 1;
            nx.lv = c->lvAlt;
            nx.nc = (1 << M) - 1;
            nx.err = c->d_err;
            nx.n_active0 = c->d_nactive + l * kPassSlots;
            nx.nMapTiles = ceil_div(c->nLocal, kMapTile);
            nx.tile_first = c->d_tile_first_alt;
            nx.tile_first_cur = c->d_tile_first;
            nx.zero = c->sel.hist;
            nx.nZero = std::min(hwNext, c->selHistWords);
            splitBlocks = std::max(splitBlocks, std::max(ceil_div(nx.nMapTiles, 256), std::min<uint32_t>(ceil_div(nx.nZero, 2048), 4u * (uint32_t)c->nSM)));
        }
        for (int attempt = 0; attempt < 2; ++attempt) {
            CK(launch_pdl(c, k_split, dim3(splitBlocks), dim3(256), 0, c->d_heap, first, nCells, c->lv, c->d_range, c->d_total, c->d_final_cut, gate, nx));
            c->nOtherLaunch++;
            if (c->tieMode != 1) {
                // (a second attempt follows the iterative fallback: the flagged cells have no PreLeft records)
                rc = launch_partition(c, nCells, c->d_tickets + (l - 1), gate, nh, attempt == 0);
                if (rc) return rc;
            }
            if (!gate) break;
            uint32_t nf = 0;
            if ((rc = select_mr_flagged(c, slot, &nf))) return rc;
            if (nf == 0u) break;
            // sampled rows whose brackets fail on more than a stray cell: the particle order is not independent of the
            // coordinates (a pre-sorted snapshot) - exact rows for the rest of this build
            if (c->levelSampled && nf > std::max<uint32_t>(1u, nCells / 64u)) c->sampleOff = true;
            c->cur ^= 1;                  // the gated partition did nothing: undo the ping-pong flip
            rc = mrPlan.ok ? select_mr_fallback(c, nCells, slot, l - 1) : select_fallback(c, nCells, slot, l - 1);
            if (rc) return rc;
            gate = nullptr;
        }
        if (nx.enabled) {       // the partition has been enqueued with this level's state: switch to the prepared one
            std::swap(c->lv, c->lvAlt);
            std::swap(c->d_tile_first, c->d_tile_first_alt);
            prepared = true;
        }
        preNb = preNext;
        if (tight) {   // children boxes from their particles; needs the children's ranges as a level
            const uint32_t cf = (1u << l) - 1u, cn = 1u << l;
            if (cn <= c->maxLevelCells) {
                rc = level_prepare(c, c->d_heap + cf, cn, 1, nullptr);
                if (rc) return rc;
                rc = bbox_level(c, nullptr, cn, c->d_bb6);
                if (rc) return rc;
                k_apply_bbox<<<ceil_div(cn, 256), 256, 0, c->stream>>>(c->d_heap + cf, cn, c->d_bb6, c->d_total);
                c->nOtherLaunch++;
            }
        }
        nDone = l;
    }
    CK(cudaEventRecord(evB, c->stream));
    if (heap_out) CK(cudaMemcpyAsync(heap_out, c->d_heap, (size_t)c->nHeap * sizeof(orb_cell), cudaMemcpyDeviceToHost, c->stream));
    int32_t iters[kMaxLevels], lvlPasses[kMaxLevels];
    uint32_t lvlUnfound[kMaxLevels];
    CK(cudaMemcpyAsync(lvlPasses, c->d_lvl_passes, sizeof(lvlPasses), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(lvlUnfound, c->d_lvl_unfound, sizeof(lvlUnfound), cudaMemcpyDeviceToHost, c->stream));
    unsigned long long ap[2] = {0, 0};
    CK(cudaMemcpyAsync(iters, c->d_level_iters, sizeof(iters), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(ap, c->d_active_particles, 16, cudaMemcpyDeviceToHost, c->stream));
    rc = check_device_err(c);
    if (rc) return rc;
    if (c->d_dbg) {
        std::vector<unsigned long long> t(64 * kMaxLevels);
        cudaMemcpy(t.data(), c->d_dbg, t.size() * 8, cudaMemcpyDeviceToHost);
        for (int l = 0; l < nDone; ++l) {
            fprintf(stderr, "dbg level %d:", l + 1);
            for (int p = 0; p < 11 && t[l * 64 + p * 5]; ++p) {
                const unsigned long long *q = &t[l * 64 + p * 5];
                fprintf(stderr, " [count %.1f bar %.1f upd %.1f bar %.1f]", (q[1] - q[0]) * 1e-3, (q[2] - q[1]) * 1e-3, (q[3] - q[2]) * 1e-3, (q[4] - q[3]) * 1e-3);
            }
            fprintf(stderr, "\n");
        }
        cudaMemset(c->d_dbg, 0, sizeof(unsigned long long) * 64 * kMaxLevels);
    }
    uint32_t selFlagged[kMaxLevels];
    CK(cudaMemcpy(selFlagged, c->d_sel_nflag, sizeof(selFlagged), cudaMemcpyDeviceToHost));
    if (getenv("ORB_DEBUG_SELECT") && c->profile) {
        for (size_t i = 0; i < c->evCountUsed; ++i) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, c->evCount[i].first, c->evCount[i].second);
            fprintf(stderr, "count[%zu] %.1f us\n", i, ms * 1e3);
        }
        for (size_t i = 0; i < c->evAuxUsed; ++i) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, c->evAux[i].e0, c->evAux[i].e1);
            fprintf(stderr, "aux L%d %s %.1f us\n", c->evAux[i].level + 1, c->evAux[i].label, ms * 1e3);
        }
        for (size_t i = 0; i < c->evPartUsed; ++i) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, c->evPart[i].first, c->evPart[i].second);
            fprintf(stderr, "part[%zu] %.1f us\n", i, ms * 1e3);
        }
    }
    if (getenv("ORB_DEBUG_SELECT")) {
        uint32_t nf[kMaxLevels];
        cudaMemcpy(nf, c->d_sel_nflag, sizeof(nf), cudaMemcpyDeviceToHost);
        fprintf(stderr, "select: cells left to the iterative search per level:");
        for (int l = 0; l < nDone; ++l) fprintf(stderr, " %u", nf[l]);
        fprintf(stderr, "\n");
    }
    if (c->d_dbg_blocks) {
        // raw dump for offline analysis: per level u32 level, u32 grid, then [kDbgPasses][grid][4] u64 stamps
        const char *path = getenv("ORB_DEBUG_TIMES_FILE");
        FILE *f = fopen(path ? path : "orb_block_times.bin", "wb");
        if (f) {
            std::vector<unsigned long long> t((size_t)kDbgPasses * kDbgBlocks * 4);
            for (int l = 0; l < nDone; ++l) {
                const uint32_t g = c->dbgGrid[l];
                if (!g) continue;
                cudaMemcpy(t.data(), c->d_dbg_blocks + (size_t)l * kDbgPasses * kDbgBlocks * 4, (size_t)kDbgPasses * g * 4 * 8, cudaMemcpyDeviceToHost);
                const uint32_t hdr[2] = {(uint32_t)l + 1u, g};
                fwrite(hdr, 4, 2, f);
                fwrite(t.data(), 8, (size_t)kDbgPasses * g * 4, f);
            }
            fclose(f);
        }
        cudaMemset(c->d_dbg_blocks, 0, sizeof(unsigned long long) * (size_t)kMaxLevels * kDbgPasses * kDbgBlocks * 4);
        memset(c->dbgGrid, 0, sizeof(c->dbgGrid));
    }
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        stats->n_levels = nDone;
        for (int l = 0; l < nDone && l < 64; ++l) {
            stats->iters[l] = iters[l];
            stats->passes[l] = passes[l] >= 0 ? passes[l] : lvlPasses[l] + c->extraPasses[l];
            stats->not_found[l] = passes[l] >= 0 ? (int32_t)unfound[l] : (int32_t)lvlUnfound[l];
        }
        stats->active_passes = ap[0];
        stats->iter_particle_passes = ap[1];
        stats->count_launches = c->nCountLaunch - l0;
        stats->update_launches = c->nUpdateLaunch - l1;
        stats->partition_launches = c->nPartLaunch - l2;
        stats->other_launches = c->nOtherLaunch - l3;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, evA, evB);
        stats->ms_total = ms;
        for (int l = 0; l < nDone; ++l) stats->search_fallback_cells += selFlagged[l];
        if (c->profile) {
            stats->ms_count = sum_events(c->evCount, c->evCountUsed);
            stats->ms_partition = sum_events(c->evPart, c->evPartUsed);
        }
    }
    return ORB_OK;
}

}  // extern "C"
