// orbit.cpp — `orbit <x> <y> [o]`: ORB domain decomposition of 2^x particles into 2^y leaf cells.
//
// Keeps the reference's driver surface (andrinr/gpu-load-balance src/orbit.cpp): same positional
// arguments (orbit.cpp:26-66), the same service-per-step structure on an mdl2-style runtime, the same
// level loop bounds including the reference's one-level-short quirk (orbit.cpp:102, opt out with
// ORB_FULL_LEVELS=1), the same float bisection rule (orbit.cpp:204-229) and the same three stdout lines
// that the sweep scripts parse (orbit.cpp:284-286).  All particle work runs on the GPU through the
// C ABI of liborb_b200.so; there is no CPU path.
//
//   o = 0 / absent : fused — the whole build runs device-side (PST_BUILD), no host round trip per iteration
//   o = 1          : the reference's control flow — master() bisects on the host, one PST_COUNTLEFTAXISGPU
//                    call per iteration (orbit.cpp:166-177), PST_PARTITIONGPU per level
//   o = 2          : per level one PST_FINDCUTS (device-side bisection loop) + PST_PARTITIONGPU
// Environment: ORB_MDL_THREADS=<ranks = GPUs>, ORB_FULL_LEVELS=1, ORB_TIGHT_BOX=1 (o=0 only),
//              ORB_DIST=uniform|gaussian|plummer, ORB_TIPSY=<snapshot> (positions from a tipsy file; x = 0 takes
//              all its bodies), ORB_DUMP=<prefix> (heap + per-rank ranges/particles).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "cell.h"
#include "constants.h"
#include "mdl.h"
#include "services/pst.h"
#include "services/services.h"
#include "services/setadd.h"
#include "tipsy/tipsy.h"

namespace {

typedef std::chrono::high_resolution_clock hr_clock;
inline long long usSince(hr_clock::time_point t0) {
    return std::chrono::duration_cast<std::chrono::microseconds>(hr_clock::now() - t0).count();
}

bool envFlag(const char *name) {
    const char *v = std::getenv(name);
    return v && std::atoi(v) != 0;
}

// One bisection decision of master() (orbit.cpp:204-229, FAST_MEDIAN branch is dead: orbit.cpp:54,59,65).
// Types and order are the reference's: the ratio is computed in double and rounded to float, the
// difference is evaluated in float and truncated to int.
bool bisectStep(Cell &cell, unsigned int countLeft, unsigned int count) {
    float ratio = std::ceil(cell.nLeafCells / 2.0) / cell.nLeafCells;
    int difference = countLeft - count * ratio;
    if (std::abs(difference) < 3) {
        cell.foundCut = true;
        return true;
    }
    if (difference > 0) cell.cutMarginRight = cell.getCut();
    else cell.cutMarginLeft = cell.getCut();
    return false;
}

// orbit.cpp:235-250: children of every cell of the level go to their heap slots
void splitLevel(std::vector<Cell> &heap, int first, int nCells) {
    for (int i = 0; i < nCells; ++i) {
        Cell left, right;
        std::tie(left, right) = heap[first + i].cut();
        right.setCutAxis();
        right.setCutMargin();
        left.setCutAxis();
        left.setCutMargin();
        heap[left.id] = left;
        heap[right.id] = right;
    }
}

void dumpHeap(const std::string &prefix, const std::vector<Cell> &heap) {
    FILE *f = std::fopen((prefix + ".heap").c_str(), "wb");
    if (!f) { std::perror(prefix.c_str()); std::exit(1); }
    std::fwrite(heap.data(), sizeof(Cell), heap.size(), f);
    std::fclose(f);
}

}  // namespace

int master(MDL vmdl, void *vpst) {
    auto mdl = static_cast<mdl::mdlClass *>(vmdl);
    (void)vpst;
    if (mdl->argc < 3) {
        std::printf("Usage: %s <x> <y> [o]   (2^x particles, 2^y leaf cells)\n", mdl->argv[0]);
        return 1;
    }
    ServiceSetAdd::input inAdd(mdl->Threads());
    mdl->RunService(PST_SETADD, sizeof(inAdd), &inAdd);

    const int pN = (int)std::strtol(mdl->argv[1], nullptr, 0);
    const int pd = (int)std::strtol(mdl->argv[2], nullptr, 0);
    int N = 1 << pN;
    const int d = 1 << pd;
    // ORB_TIPSY=<file>: positions come from a tipsy snapshot (the reference's dead `generate == false` branch,
    // init.cu:54-59); x = 0 takes every body of the file, otherwise the first 2^x
    const char *tipsyPath = std::getenv("ORB_TIPSY");
    if (tipsyPath && *tipsyPath) {
        TipsyIO io;
        if (!io.open(tipsyPath)) {
            std::fprintf(stderr, "orbit: %s\n", io.error().c_str());
            return 1;
        }
        std::fprintf(stderr, "Count %llu n %d\n", (unsigned long long)io.count(), pN ? N : (int)io.count());   // init.cu:57
        if (pN == 0) N = (int)io.count();
        if ((unsigned long long)N > io.count() || N < mdl->Threads()) {
            std::fprintf(stderr, "orbit: %s holds %llu bodies, fewer than the 2^%d requested\n", tipsyPath, (unsigned long long)io.count(), pN);
            return 1;
        }
    } else {
        tipsyPath = nullptr;
    }
    const int mode = mdl->argc > 3 ? (int)std::strtol(mdl->argv[3], nullptr, 0) : 0;
    if (mode < 0 || mode > 2) {
        std::fprintf(stderr, "orbit: o must be 0, 1 or 2\n");
        return 1;
    }
    META_PARAMS params;
    params.GPU_COUNT = true;
    params.GPU_PARTITION = true;
    params.FAST_MEDIAN = false;
    const bool fullLevels = envFlag("ORB_FULL_LEVELS");
    const bool tightBox = envFlag("ORB_TIGHT_BOX");

    long long tPartitions = 0, tCountCopy = 0, tMakeAxis = 0;

    // root cell and heap (orbit.cpp:45-46,74-81)
    float lower[3] = {-0.5f, -0.5f, -0.5f}, upper[3] = {0.5f, 0.5f, 0.5f};
    Cell root(0, d, lower, upper);
    root.cutAxis = 0;
    root.setCutMargin();
    std::vector<Cell> heap((size_t)root.getTotalNumberOfCells());
    std::memset(heap.data(), 0, heap.size() * sizeof(Cell));
    heap[0] = root;

    ServiceInit::input iInit{N / mdl->Threads(), d, tipsyPath == nullptr, params, N};
    ServiceInit::output oInit[1];
    mdl->RunService(PST_INIT, sizeof(iInit), &iInit, oInit);

    // the clocks start after Init, like the reference's (orbit.cpp:85-87); the one upload is charged to CountCopy
    {
        auto t0 = hr_clock::now();
        ServiceCopyParticles::input iCopy{params};
        ServiceCopyParticles::output oCopy[1];
        mdl->RunService(PST_COPYPARTICLES, sizeof(iCopy), &iCopy, oCopy);
        tCountCopy += usSince(t0);
    }

    if (mode == 0) {
        // ---- fused: the level loop of orbit.cpp:102-275 runs on the device ----
        ServiceBuild::input iBuild{(fullLevels ? ORB_FULL_LEVELS : 0u) | (tightBox ? ORB_TIGHT_BOX : 0u), heap.data()};
        ServiceBuild::output oBuild[1];
        auto t0 = hr_clock::now();
        mdl->RunService(PST_BUILD, sizeof(iBuild), &iBuild, oBuild);
        const long long us = usSince(t0);
        // split of the wall time between the two reference timers, from the library's own event clock
        const double fracPart = (oBuild[0].ms_total > 0 && oBuild[0].ms_partition > 0) ? oBuild[0].ms_partition / oBuild[0].ms_total : 0.0;
        tPartitions += (long long)(us * fracPart);
        tCountCopy += us - (long long)(us * fracPart);
    } else {
        const int nLevels = root.getNLevels();
        const int lEnd = fullLevels ? nLevels + 1 : nLevels;
        std::vector<unsigned int> oCounts((size_t)d), oCountsLeft((size_t)d);     // orbit.cpp:99-100, off the stack
        std::vector<Cell> found((size_t)d);
        for (int l = 1; l < lEnd; ++l) {                                           // orbit.cpp:102
            const int a = (1 << (l - 1)) - 1;                                      // orbit.cpp:104
            const int b = std::min(root.getNCellsOnLastLevel(), 1 << l) - 2;       // orbit.cpp:105-107
            const int nCells = b - a + 1;
            Cell *cells = heap.data() + a;                                         // aliases the heap (orbit.cpp:111)

            mdl->RunService(PST_COUNT, nCells * sizeof(Cell), cells, oCounts.data());       // orbit.cpp:124
            {
                ServiceCopyCells::output oCopy[1];
                mdl->RunService(PST_COPYCELLS, nCells * sizeof(Cell), cells, oCopy);        // orbit.cpp:140-143
            }
            if (mode == 1) {
                bool foundAll = false;
                int j = 0;
                while (!foundAll && j < 32) {                                      // orbit.cpp:149
                    ++j;
                    foundAll = true;
                    auto t0 = hr_clock::now();
                    mdl->RunService(PST_COUNTLEFTAXISGPU, nCells * sizeof(Cell), cells, oCountsLeft.data());
                    tCountCopy += usSince(t0);
                    for (int i = 0; i < nCells; ++i) {
                        if (cells[i].foundCut) continue;
                        if (!bisectStep(cells[i], oCountsLeft[i], oCounts[i])) foundAll = false;
                    }
                }
            } else {
                auto t0 = hr_clock::now();
                mdl->RunService(PST_FINDCUTS, nCells * sizeof(Cell), cells, found.data());
                std::memcpy(cells, found.data(), (size_t)nCells * sizeof(Cell));
                tCountCopy += usSince(t0);
            }
            splitLevel(heap, a, nCells);
            auto t0 = hr_clock::now();
            ServicePartitionGPU::output oPart[1];
            mdl->RunService(PST_PARTITIONGPU, nCells * sizeof(Cell), cells, oPart);         // orbit.cpp:252-262
            tPartitions += usSince(t0);
        }
    }

    if (const char *prefix = std::getenv("ORB_DUMP")) {
        dumpHeap(prefix, heap);
        ServiceDump::input iDump;
        std::memset(&iDump, 0, sizeof(iDump));
        std::strncpy(iDump.path, prefix, sizeof(iDump.path) - 1);
        ServiceDump::output oDump[1];
        mdl->RunService(PST_DUMP, sizeof(iDump), &iDump, oDump);
    }

    ServiceFinalize::input iFree{params};
    ServiceFinalize::output oFree[1];
    mdl->RunService(PST_FINALIZE, sizeof(iFree), &iFree, oFree);

    // orbit.cpp:284-286 — same three lines (CountCopy / Partition in microseconds, MakeAxis in milliseconds;
    // MakeAxis is always 0: the axis-column copy is designed out)
    std::printf("CountCopy-%u-%u, %u \n", pN, pd, (unsigned)tCountCopy);
    std::printf("Partition-%u-%u, %u \n", pN, pd, (unsigned)tPartitions);
    std::printf("MakeAxis-%u-%u, %u \n", pN, pd, (unsigned)tMakeAxis);
    return 0;
}

void *worker_init(MDL vmdl) {
    auto mdl = static_cast<mdl::mdlClass *>(vmdl);
    auto pst = new pstNode(mdl);
    pst->lcl = new LocalData();
    mdl->AddService(std::make_unique<ServiceSetAdd>(pst));
    mdl->AddService(std::make_unique<ServiceInit>(pst));
    mdl->AddService(std::make_unique<ServiceCount>(pst));
    mdl->AddService(std::make_unique<ServiceCopyParticles>(pst));
    mdl->AddService(std::make_unique<ServiceCopyCells>(pst));
    mdl->AddService(std::make_unique<ServiceCountLeftGPU>(pst));
    mdl->AddService(std::make_unique<ServiceCountLeftGPUAxis>(pst));
    mdl->AddService(std::make_unique<ServicePartitionGPU>(pst));
    // the ids of the reference's CPU services (orbit.cpp:303,306,309): registered so that its o=0 / o=1 control flow
    // dispatches here too - count and partition run on the device, MakeAxis has nothing to do
    mdl->AddService(std::make_unique<ServiceCountLeft>(pst));
    mdl->AddService(std::make_unique<ServicePartition>(pst));
    mdl->AddService(std::make_unique<ServiceMakeAxis>(pst));
    mdl->AddService(std::make_unique<ServiceFinalize>(pst));
    mdl->AddService(std::make_unique<ServiceBBox>(pst));
    mdl->AddService(std::make_unique<ServiceFindCuts>(pst));
    mdl->AddService(std::make_unique<ServiceBuild>(pst));
    mdl->AddService(std::make_unique<ServiceDump>(pst));
    return pst;
}

void worker_done(MDL, void *ctx) {
    auto pst = static_cast<PST>(ctx);
    delete pst;
}

int main(int argc, char **argv) { return mdlLaunch(argc, argv, master, worker_init, worker_done); }
