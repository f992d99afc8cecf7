// TEST INFRASTRUCTURE ONLY (oracle/): observer linked into oracle/_ref/orbit_ref,
// the binary built from the UNMODIFIED reference sources.  The reference never
// writes a result (its tree dies on master()'s stack, orbit.cpp:79,284-286), so
// this file hooks the mdl runtime's RunService tap and dumps, per service call,
// what the reference computed: the Cell array it passed (orbit.cpp:111-112
// aliases the heap, so margins/foundCut are live), the counts it got back
// (count.cpp:16, countLeft.cpp:31-38), and after each partition
// (partition.cpp:54-60) the child ranges plus hashes of the particles in them.
//
// Enabled by ORB_REF_TRACE=<file>; ORB_REF_TRACE_PARTICLES=1 additionally dumps
// the x,y,z columns after Init and after every partition (small N only).
// Only meaningful with ORB_MDL_THREADS=1 (thread 0's LocalData is dumped; the
// reference's generator is racy with more threads, init.cu:11).
//
// Record layout (little endian): u32 kind, u32 nCells, u64 payloadBytes, payload.
//   kind = pst_service id for service records; 1000 = particle dump.
#include <chrono>
#include <csignal>
#include <execinfo.h>
#include <unistd.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda_runtime.h>
#include "mdl.h"
#include "cell.h"           // /root/reference/src/cell.h
#include "services/pst.h"   // /root/reference/src/services/pst.h (LocalData, pst_service)

namespace {

FILE *g_trace = nullptr;
bool g_dumpParticles = false;
int g_nParticles = 0;
std::chrono::steady_clock::time_point g_tInitDone, g_tLast;
bool g_haveInit = false;
std::vector<unsigned> g_levelCounts;          // last ServiceCount output (global particles per cell)
unsigned long long g_particlePasses = 0;      // sum over count-left calls of particles in unfound cells

inline uint64_t mix64(uint64_t v) {   // splitmix64 finaliser
    v ^= v >> 30; v *= 0xbf58476d1ce4e5b9ULL;
    v ^= v >> 27; v *= 0x94d049bb133111ebULL;
    v ^= v >> 31;
    return v;
}

inline uint32_t fbits(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }

// per-particle hash of the (x,y,z) bit patterns
inline uint64_t particleHash(float x, float y, float z) {
    uint64_t a = ((uint64_t)fbits(x) << 32) | fbits(y);
    uint64_t b = (uint64_t)fbits(z) | 0x9e3779b900000000ULL;
    return mix64(a ^ mix64(b));
}

void put(const void *p, size_t n) { std::fwrite(p, 1, n, g_trace); }

void header(uint32_t kind, uint32_t nCells, uint64_t bytes) {
    put(&kind, 4); put(&nCells, 4); put(&bytes, 8);
}

void dumpParticles(LocalData *lcl) {
    // init.cu:32-45: column-major (N,3) => three contiguous float[N] columns
    uint64_t bytes = 4 + (uint64_t)g_nParticles * 12;
    header(1000, 0, bytes);
    uint32_t n = (uint32_t)g_nParticles;
    put(&n, 4);
    for (int d = 0; d < 3; ++d)
        for (int i = 0; i < g_nParticles; ++i) { float v = lcl->particles(i, d); put(&v, 4); }
}

void tap(mdl::mdlClass *mdl, int sid, int nIn, void *pIn, void *pOut, int /*nOut*/) {
    auto now = std::chrono::steady_clock::now();
    g_tLast = now;
    PST pst = static_cast<PST>(mdl->worker_ctx);
    LocalData *lcl = pst->lcl;
    if (sid == PST_INIT) {
        g_tInitDone = now;
        g_haveInit = true;
        g_nParticles = *static_cast<int *>(pIn);   // ServiceInit::input.nParticles is the first field (init.h:7-12)
        if (g_trace) {
            header(PST_INIT, 0, 4);
            uint32_t n = (uint32_t)g_nParticles;
            put(&n, 4);
            if (g_dumpParticles) dumpParticles(lcl);
        }
        return;
    }
    uint32_t nCells = (uint32_t)(nIn / sizeof(Cell));
    const Cell *cells = static_cast<const Cell *>(pIn);
    // cheap bookkeeping for the CPU-baseline throughput figure (no trace file needed)
    if (sid == PST_COUNT) {
        g_levelCounts.assign(static_cast<unsigned *>(pOut), static_cast<unsigned *>(pOut) + nCells);
    } else if (sid == PST_COUNTLEFT || sid == PST_COUNTLEFTGPU || sid == PST_COUNTLEFTAXISGPU) {
        for (uint32_t c = 0; c < nCells && c < g_levelCounts.size(); ++c)
            if (!cells[c].foundCut) g_particlePasses += g_levelCounts[c];
    }
    if (!g_trace) return;
    switch (sid) {
    case PST_COUNT:
    case PST_COUNTLEFT:
    case PST_COUNTLEFTGPU:
    case PST_COUNTLEFTAXISGPU: {
        header((uint32_t)sid, nCells, (uint64_t)nCells * (sizeof(Cell) + 4));
        put(cells, nCells * sizeof(Cell));
        put(pOut, nCells * 4);
        break;
    }
    case PST_PARTITION:
    case PST_PARTITIONGPU: {
        // per cell: child ranges (4 x u32) + multiset hash and ordered hash of each child (4 x u64)
        header((uint32_t)sid, nCells, (uint64_t)nCells * (sizeof(Cell) + 16 + 32));
        put(cells, nCells * sizeof(Cell));
        for (uint32_t c = 0; c < nCells; ++c) {
            int ids[2] = {cells[c].getLeftChildId(), cells[c].getRightChildId()};
            uint32_t r[4];
            uint64_t h[4];
            for (int k = 0; k < 2; ++k) {
                uint32_t b = lcl->cellToRangeMap(ids[k], 0), e = lcl->cellToRangeMap(ids[k], 1);
                r[2 * k] = b; r[2 * k + 1] = e;
                uint64_t set = 0, ord = 1469598103934665603ULL;
                for (uint32_t i = b; i < e && (int)i < g_nParticles; ++i) {
                    uint64_t ph = particleHash(lcl->particles(i, 0), lcl->particles(i, 1), lcl->particles(i, 2));
                    set += ph;
                    ord = (ord ^ ph) * 1099511628211ULL;
                }
                h[2 * k] = set; h[2 * k + 1] = ord;
            }
            put(r, 16);
            put(h, 32);
        }
        if (g_dumpParticles) dumpParticles(lcl);
        break;
    }
    default:
        break;
    }
}

void onSegv(int sig) {
    void *frames[48];
    int n = backtrace(frames, 48);
    const char msg[] = "orbit_ref: fatal signal, backtrace:\n";
    if (write(2, msg, sizeof(msg) - 1) < 0) {}
    backtrace_symbols_fd(frames, n, 2);
    _exit(128 + sig);
}

struct Installer {
    Installer() {
        std::signal(SIGSEGV, onSegv);
        std::signal(SIGBUS, onSegv);
        const char *path = std::getenv("ORB_REF_TRACE");
        if (path && *path) {
            g_trace = std::fopen(path, "wb");
            if (!g_trace) { std::perror("ORB_REF_TRACE"); std::exit(2); }
            const char magic[8] = {'O', 'R', 'B', 'T', 'R', 'A', 'C', 'E'};
            uint32_t ver = 1, cellBytes = (uint32_t)sizeof(Cell);
            put(magic, 8); put(&ver, 4); put(&cellBytes, 4);
        }
        const char *dp = std::getenv("ORB_REF_TRACE_PARTICLES");
        g_dumpParticles = dp && std::atoi(dp) != 0;
        mdl::setRunServiceTap(tap);
    }
    ~Installer() {
        if (g_trace) std::fclose(g_trace);
        if (g_haveInit) {
            // wall time of the ORB build proper: end of Init -> last service return
            // (the reference starts its own clocks after Init too, orbit.cpp:85-87)
            long long us = std::chrono::duration_cast<std::chrono::microseconds>(g_tLast - g_tInitDone).count();
            std::fprintf(stderr, "RefBuildWall-us, %lld\n", us);
            std::fprintf(stderr, "RefParticlePasses, %llu\n", g_particlePasses);
        }
    }
} g_installer;

}  // namespace

// The reference allocates its particle array with cudaMallocHost even in the
// CPU-only mode (init.cu:39).  On a box without a GPU that call fails and
// CUDA_CHECK exits (constants.h:19-23).  This definition in the executable
// pre-empts libcudart.so's: it forwards to the real function and, only if that
// fails, falls back to ordinary aligned host memory so the o=0 path can run.
#include <dlfcn.h>
extern "C" cudaError_t CUDARTAPI cudaMallocHost(void **ptr, size_t size) {
    typedef cudaError_t (*fn_t)(void **, size_t);
    static fn_t real = (fn_t)dlsym(RTLD_NEXT, "cudaMallocHost");
    // + one page: MakeAxis copies end-begin+1 floats (blitz Range is inclusive, makeAxis.cpp:23-28), i.e. it reads
    // one float past the last column; with exact page-granular pinned allocations that read faults
    size += 4096;
    if (real) {
        cudaError_t rc = real(ptr, size);
        if (rc == cudaSuccess) return rc;
        typedef cudaError_t (*gle_t)(void);
        static gle_t gle = (gle_t)dlsym(RTLD_NEXT, "cudaGetLastError");
        if (gle) gle();
    }
    void *p = nullptr;
    if (posix_memalign(&p, 4096, size ? size : 1) != 0) return cudaErrorMemoryAllocation;
    *ptr = p;
    return cudaSuccess;
}
