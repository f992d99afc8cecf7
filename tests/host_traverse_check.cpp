// CPU-only check of the host runtime the C++ driver is built on: ServiceSetAdd organises the rank threads into a
// tree, TraverseCombinePST services fan out over it, every rank runs Service() exactly once on its own LocalData and
// the replies are folded with Combine().  Built and run by tests/test_host_runtime.py (no GPU, no liborb_b200.so).
#include <cstdio>
#include <cstring>
#include <memory>
#include <vector>

#include "services/TraversePST.h"
#include "services/setadd.h"

enum { SVC_TALLY = 100, SVC_SCALE = 101, SVC_NOTIFY = 102 };

// fixed-size output: who ran, how often
class ServiceTally : public TraverseCombinePST {
public:
    struct input { int offset; };
    struct output { long long sum; int calls; unsigned long long mask; };
    explicit ServiceTally(PST pst) : TraverseCombinePST(pst, SVC_TALLY, sizeof(input), sizeof(output), "Tally") {}
protected:
    int Service(PST pst, void *vin, int nIn, void *vout, int nOut) override {
        mdlassert(pst->mdl, nIn == (int)sizeof(input) && nOut == (int)sizeof(output));
        mdlassert(pst->mdl, pst->AmCore() && pst->lcl && pst->lcl->rank == mdlSelf(pst->mdl));
        auto *o = static_cast<output *>(vout);
        o->sum = pst->lcl->rank + static_cast<input *>(vin)->offset;
        o->calls = 1;
        o->mask = 1ull << pst->lcl->rank;
        return sizeof(output);
    }
    int Combine(void *vout, void *vout2, int, int, int nOut2) override {
        auto *a = static_cast<output *>(vout);
        auto *b = static_cast<output *>(vout2);
        if (nOut2 != (int)sizeof(output)) std::abort();
        a->sum += b->sum; a->calls += b->calls; a->mask |= b->mask;
        return sizeof(output);
    }
};

// output size follows the input size (the shape of ServiceCount / ServiceCountLeft: one word per cell)
class ServiceScale : public TraverseCombinePST {
public:
    typedef int input;
    typedef long long output;
    explicit ServiceScale(PST pst) : TraverseCombinePST(pst, SVC_SCALE, 4096 * sizeof(input), 4096 * sizeof(output), "Scale") {}
protected:
    int Service(PST pst, void *vin, int nIn, void *vout, int) override {
        const int n = nIn / (int)sizeof(input);
        for (int i = 0; i < n; ++i) static_cast<output *>(vout)[i] = (long long)static_cast<input *>(vin)[i] * (pst->lcl->rank + 1);
        return n * (int)sizeof(output);
    }
    int Combine(void *vout, void *vout2, int nIn, int, int nOut2) override {
        const int n = nIn / (int)sizeof(input);
        if (nOut2 != n * (int)sizeof(output)) std::abort();
        for (int i = 0; i < n; ++i) static_cast<output *>(vout)[i] += static_cast<output *>(vout2)[i];
        return n * (int)sizeof(output);
    }
};

static int master(MDL vmdl, void *) {
    auto *mdl = static_cast<mdl::mdlClass *>(vmdl);
    const int T = mdl->Threads();
    ServiceSetAdd::input all(T);
    mdl->RunService(PST_SETADD, sizeof(all), &all);

    for (int rep = 0; rep < 3; ++rep) {
        ServiceTally::input in{10 * rep};
        ServiceTally::output out;
        std::memset(&out, 0, sizeof(out));
        const int nOut = mdl->RunService(SVC_TALLY, sizeof(in), &in, &out);
        const long long want = (long long)T * (T - 1) / 2 + (long long)T * in.offset;
        const unsigned long long full = T >= 64 ? ~0ull : ((1ull << T) - 1ull);
        if (nOut != (int)sizeof(out) || out.sum != want || out.calls != T || out.mask != full) {
            std::printf("tally mismatch: nOut %d sum %lld (want %lld) calls %d mask %llx\n", nOut, out.sum, want, out.calls, out.mask);
            return 1;
        }
    }
    for (int n : {1, 7, 4096}) {
        std::vector<int> in(n);
        std::vector<long long> out(4096, -1);
        for (int i = 0; i < n; ++i) in[i] = 3 * i - 5;
        const int nOut = mdl->RunService(SVC_SCALE, n * (int)sizeof(int), in.data(), out.data());
        if (nOut != n * (int)sizeof(long long)) { std::printf("scale: nOut %d for n %d\n", nOut, n); return 1; }
        for (int i = 0; i < n; ++i)
            if (out[i] != (long long)in[i] * T * (T + 1) / 2) { std::printf("scale mismatch at %d\n", i); return 1; }
    }
    std::printf("host runtime ok: %d ranks\n", T);
    return 0;
}

static void *worker_init(MDL vmdl) {
    auto *mdl = static_cast<mdl::mdlClass *>(vmdl);
    auto *pst = new pstNode(mdl);
    pst->lcl = new LocalData();
    pst->lcl->rank = mdl->Self();
    pst->lcl->nRanks = mdl->Threads();
    mdl->AddService(std::make_unique<ServiceSetAdd>(pst));
    mdl->AddService(std::make_unique<ServiceTally>(pst));
    mdl->AddService(std::make_unique<ServiceScale>(pst));
    return pst;
}

static void worker_done(MDL, void *ctx) {
    auto *pst = static_cast<PST>(ctx);
    delete pst->lcl;
    delete pst;
}

int main(int argc, char **argv) { return mdlLaunch(argc, argv, master, worker_init, worker_done); }
