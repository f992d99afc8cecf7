// services.cpp — Service()/Combine() bodies: thin calls into the C ABI (include/orb_b200.h).
#include "services.h"
#include "../tipsy/tipsy.h"

#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <type_traits>
#include <vector>

static_assert(std::is_trivial<ServiceInit::input>() && std::is_trivial<ServiceCount::input>() &&
                  std::is_trivial<ServiceBBox::output>() && std::is_trivial<ServiceBuild::input>() &&
                  std::is_trivial<ServiceDump::input>(),
              "service wire types travel by memcpy");
static_assert(sizeof(Cell) == sizeof(orb_cell), "Cell and orb_cell are the same 52-byte record");

namespace {
inline const orb_cell *asOrb(const void *cells) { return static_cast<const orb_cell *>(cells); }
inline orb_cell *asOrb(void *cells) { return static_cast<orb_cell *>(cells); }

// one NCCL id per launch, made by whichever rank thread arrives first
std::once_flag g_idOnce;
unsigned char g_ncclId[128];

// peer descriptors of all rank threads (fused count+combine over NVLink); a counting barrier separates
// "everyone exported" from "everyone imports"
orb_peer_info g_peers[8];
std::mutex g_peerMutex;
std::condition_variable g_peerCv;
int g_peerArrived = 0;
void peerBarrier(int nRanks) {
    std::unique_lock<std::mutex> lk(g_peerMutex);
    const int target = ((g_peerArrived / nRanks) + 1) * nRanks;
    ++g_peerArrived;
    g_peerCv.notify_all();
    g_peerCv.wait(lk, [&] { return g_peerArrived >= target; });
}
}  // namespace

// ------------------------------------------------------------------ Init (init.cu:27-144)
int ServiceInit::Service(PST pst, void *vin, int, void *, int) {
    LocalData *lcl = pst->lcl;
    const input in = *static_cast<input *>(vin);
    lcl->rank = mdlSelf(pst->mdl);
    lcl->nRanks = mdlThreads(pst->mdl);
    lcl->nLeafCells = in.d;
    if (in.generate) {
        lcl->nParticles = in.nParticles;
        lcl->firstParticle = (unsigned long long)lcl->rank * (unsigned long long)in.nParticles;
    } else {
        // contiguous slice [r*T/R, (r+1)*T/R) of the first T bodies of the file
        const unsigned long long T = (unsigned long long)in.nTotal, R = (unsigned long long)lcl->nRanks;
        lcl->firstParticle = (unsigned long long)lcl->rank * T / R;
        lcl->nParticles = (int)(((unsigned long long)lcl->rank + 1) * T / R - lcl->firstParticle);
    }
    lcl->x.resize(lcl->nParticles);
    lcl->y.resize(lcl->nParticles);
    lcl->z.resize(lcl->nParticles);
    // deterministic slice of the single xorshf96 stream (the reference's generator state is a racy
    // file-scope static, init.cu:11, so its multi-thread runs are not reproducible; this is)
    const char *dist = std::getenv("ORB_DIST");
    if (!in.generate) {
        // init.cu:54-59: `TipsyIO io; io.open(...); io.load(particles)` — every rank reads only its slice
        const char *path = std::getenv("ORB_TIPSY");
        TipsyIO io;
        if (!path || !io.open(path) || !io.load(lcl->firstParticle, (uint64_t)lcl->nParticles, lcl->x.data(), lcl->y.data(), lcl->z.data())) {
            std::fprintf(stderr, "orbit: tipsy input: %s\n", path ? io.error().c_str() : "ORB_TIPSY is not set");
            std::exit(1);
        }
    } else if (dist && std::string(dist) == "gaussian") orb_generate_clustered(0, lcl->firstParticle, in.nParticles, lcl->x.data(), lcl->y.data(), lcl->z.data());
    else if (dist && std::string(dist) == "plummer") orb_generate_clustered(1, lcl->firstParticle, in.nParticles, lcl->x.data(), lcl->y.data(), lcl->z.data());
    else orb_generate_uniform(lcl->firstParticle, in.nParticles, lcl->x.data(), lcl->y.data(), lcl->z.data());
    ORB_CHECK(orb_create(&lcl->ctx, lcl->rank, (uint64_t)lcl->nParticles, (uint32_t)in.d));
    if (lcl->nRanks > 1) {
        std::call_once(g_idOnce, [] { ORB_CHECK(orb_comm_unique_id(g_ncclId)); });
        ORB_CHECK(orb_comm_init(lcl->ctx, g_ncclId, lcl->rank, lcl->nRanks));
        const char *noPeer = std::getenv("ORB_NO_PEER");
        if (lcl->nRanks <= 8 && !(noPeer && std::atoi(noPeer) != 0)) {
            ORB_CHECK(orb_peer_export(lcl->ctx, &g_peers[lcl->rank]));
            peerBarrier(lcl->nRanks);
            ORB_CHECK(orb_peer_import(lcl->ctx, g_peers, lcl->nRanks));
            peerBarrier(lcl->nRanks);
        }
    }
    return 0;
}
int ServiceInit::Combine(void *, void *, int, int, int) { return 0; }

// ------------------------------------------------------------------ CopyParticles (copyParticles.cu:29-57)
int ServiceCopyParticles::Service(PST pst, void *, int, void *, int) {
    LocalData *lcl = pst->lcl;
    ORB_CHECK(orb_upload_xyz(lcl->ctx, lcl->x.data(), lcl->y.data(), lcl->z.data()));
    return 0;
}
int ServiceCopyParticles::Combine(void *, void *, int, int, int) { return 0; }

// ------------------------------------------------------------------ CopyCells (copyCells.cu:10-64): nothing to stage
int ServiceCopyCells::Service(PST, void *, int, void *, int) { return 0; }
int ServiceCopyCells::Combine(void *, void *, int, int, int) { return 0; }

// ------------------------------------------------------------------ Count (count.cpp:8-30)
int ServiceCount::Service(PST pst, void *vin, int nIn, void *vout, int) {
    const unsigned nCells = nIn / sizeof(input);
    ORB_CHECK(orb_count(pst->lcl->ctx, asOrb(vin), nCells, static_cast<output *>(vout)));
    return nCells * sizeof(output);
}
int ServiceCount::Combine(void *, void *, int nIn, int, int) { return (nIn / sizeof(input)) * sizeof(output); }

// ------------------------------------------------------------------ CountLeft on the GPU (countLeftGPU.cu:80-162, countLeftGPUAxis.cu:188-259)
int ServiceCountLeftGPU::Service(PST pst, void *vin, int nIn, void *vout, int) {
    const unsigned nCells = nIn / sizeof(input);
    ORB_CHECK(orb_count_left(pst->lcl->ctx, asOrb(vin), nCells, static_cast<output *>(vout)));
    return nCells * sizeof(output);
}
int ServiceCountLeftGPU::Combine(void *, void *, int nIn, int, int) { return (nIn / sizeof(input)) * sizeof(output); }

int ServiceCountLeftGPUAxis::Service(PST pst, void *vin, int nIn, void *vout, int) {
    const unsigned nCells = nIn / sizeof(input);
    ORB_CHECK(orb_count_left(pst->lcl->ctx, asOrb(vin), nCells, static_cast<output *>(vout)));
    return nCells * sizeof(output);
}
int ServiceCountLeftGPUAxis::Combine(void *, void *, int nIn, int, int) { return (nIn / sizeof(input)) * sizeof(output); }

// the CPU service id (countLeft.cpp:9-53): same device count
int ServiceCountLeft::Service(PST pst, void *vin, int nIn, void *vout, int) {
    const unsigned nCells = nIn / sizeof(input);
    ORB_CHECK(orb_count_left(pst->lcl->ctx, asOrb(vin), nCells, static_cast<output *>(vout)));
    return nCells * sizeof(output);
}
int ServiceCountLeft::Combine(void *, void *, int nIn, int, int) { return (nIn / sizeof(input)) * sizeof(output); }

// ------------------------------------------------------------------ MakeAxis (makeAxis.cpp:9-33): nothing to gather
int ServiceMakeAxis::Service(PST, void *, int, void *, int) { return 0; }
int ServiceMakeAxis::Combine(void *, void *, int, int, int) { return 0; }

// ------------------------------------------------------------------ Partition under the CPU service id (partition.cpp:18-65)
int ServicePartition::Service(PST pst, void *vin, int nIn, void *, int) {
    ORB_CHECK(orb_partition(pst->lcl->ctx, asOrb(vin), nIn / sizeof(input)));
    return 0;
}
int ServicePartition::Combine(void *, void *, int, int, int) { return 0; }

// ------------------------------------------------------------------ Partition on the GPU (partitionGPU.cu:283-527)
int ServicePartitionGPU::Service(PST pst, void *vin, int nIn, void *, int) {
    ORB_CHECK(orb_partition(pst->lcl->ctx, asOrb(vin), nIn / sizeof(input)));
    return 0;
}
int ServicePartitionGPU::Combine(void *, void *, int, int, int) { return 0; }

// ------------------------------------------------------------------ Finalize (finalize.cu:15-45)
int ServiceFinalize::Service(PST pst, void *, int, void *, int) {
    LocalData *lcl = pst->lcl;
    if (lcl->ctx) {
        // leave the partitioned particles in host memory, where the reference's CPU partition leaves them
        ORB_CHECK(orb_download_xyz(lcl->ctx, lcl->x.data(), lcl->y.data(), lcl->z.data()));
        ORB_CHECK(orb_destroy(lcl->ctx));
        lcl->ctx = nullptr;
    }
    return 0;
}
int ServiceFinalize::Combine(void *, void *, int, int, int) { return 0; }

// ------------------------------------------------------------------ BBox (new)
int ServiceBBox::Service(PST pst, void *vin, int nIn, void *vout, int) {
    const unsigned nCells = nIn / sizeof(input);
    ORB_CHECK(orb_bbox(pst->lcl->ctx, asOrb(vin), nCells, static_cast<float *>(vout)));
    return nCells * sizeof(output);
}
int ServiceBBox::Combine(void *, void *, int nIn, int, int) { return (nIn / sizeof(input)) * sizeof(output); }

// ------------------------------------------------------------------ FindCuts (new)
int ServiceFindCuts::Service(PST pst, void *vin, int nIn, void *vout, int) {
    const unsigned nCells = nIn / sizeof(input);
    std::memcpy(vout, vin, (size_t)nCells * sizeof(Cell));
    ORB_CHECK(orb_find_cuts(pst->lcl->ctx, asOrb(vout), nCells, nullptr, nullptr));
    return nCells * sizeof(output);
}
int ServiceFindCuts::Combine(void *, void *, int nIn, int, int) { return (nIn / sizeof(input)) * sizeof(output); }

// ------------------------------------------------------------------ Build (new)
int ServiceBuild::Service(PST pst, void *vin, int, void *vout, int) {
    const input in = *static_cast<input *>(vin);
    LocalData *lcl = pst->lcl;
    // every rank computes the identical heap (replicated tree); only rank 0 copies it out
    orb_cell *heap = (lcl->rank == 0) ? reinterpret_cast<orb_cell *>(in.heapOut) : nullptr;
    ORB_CHECK(orb_build(lcl->ctx, in.flags, heap, static_cast<output *>(vout)));
    return sizeof(output);
}
int ServiceBuild::Combine(void *, void *, int, int, int) { return sizeof(output); }

// ------------------------------------------------------------------ Dump (new; parity harness)
// file "<path>.<rank>": u32 nHeap, u32 nParticles, ranges[nHeap][2], x[n], y[n], z[n] (device order)
int ServiceDump::Service(PST pst, void *vin, int, void *, int) {
    LocalData *lcl = pst->lcl;
    const input *in = static_cast<input *>(vin);
    const unsigned nHeap = 2u * (unsigned)lcl->nLeafCells - 1u, n = (unsigned)lcl->nParticles;
    std::vector<uint32_t> rng((size_t)nHeap * 2);
    std::vector<float> x(n), y(n), z(n);
    ORB_CHECK(orb_get_ranges(lcl->ctx, 0, nHeap, rng.data()));
    ORB_CHECK(orb_download_xyz(lcl->ctx, x.data(), y.data(), z.data()));
    const std::string path = std::string(in->path) + "." + std::to_string(lcl->rank);
    FILE *f = std::fopen(path.c_str(), "wb");
    if (!f) { std::perror(path.c_str()); std::exit(1); }
    std::fwrite(&nHeap, 4, 1, f);
    std::fwrite(&n, 4, 1, f);
    std::fwrite(rng.data(), 4, rng.size(), f);
    std::fwrite(x.data(), 4, n, f);
    std::fwrite(y.data(), 4, n, f);
    std::fwrite(z.data(), 4, n, f);
    std::fclose(f);
    return 0;
}
int ServiceDump::Combine(void *, void *, int, int, int) { return 0; }
