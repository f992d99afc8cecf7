// TraversePST.h — service base classes (reference: src/services/TraversePST.h:11-60).
// A service is called on thread 0; it walks the PST (binary tree over ranks): send the request to the
// upper half, recurse into the lower half, fetch the reply, fold it with Combine.  At a leaf (one rank)
// Service() runs.  Same class names and virtual signatures as the reference so service code ports 1:1.
#ifndef ORB_HOST_TRAVERSEPST_H
#define ORB_HOST_TRAVERSEPST_H
#include <cstdint>

#include "pst.h"

class TraversePST : public mdl::BasicService {
    PST node_pst;
public:
    explicit TraversePST(PST pst, int service_id, int nInBytes, int nOutBytes, const char *service_name = "")
        : BasicService(service_id, nInBytes, nOutBytes, service_name), node_pst(pst) {}
    explicit TraversePST(PST pst, int service_id, int nInBytes, const char *service_name = "")
        : BasicService(service_id, nInBytes, 0, service_name), node_pst(pst) {}
    explicit TraversePST(PST pst, int service_id, const char *service_name = "")
        : BasicService(service_id, 0, 0, service_name), node_pst(pst) {}
    virtual ~TraversePST() = default;

protected:
    virtual int operator()(int nIn, void *pIn, void *pOut) final;
    virtual int Traverse(PST pst, void *vin, int nIn, void *vout, int nOut);
    virtual int OffNode(PST pst, void *vin, int nIn, void *vout, int nOut) { return Recurse(pst, vin, nIn, vout, nOut); }
    virtual int AtNode(PST pst, void *vin, int nIn, void *vout, int nOut) { return Recurse(pst, vin, nIn, vout, nOut); }
    virtual int Recurse(PST pst, void *vin, int nIn, void *vout, int nOut);
    virtual int Service(PST pst, void *vin, int nIn, void *vout, int nOut) = 0;
    static int Traverse(unsigned sid, PST pst, void *vin, int nIn, void *vout, int nOut);
};

// Same input on every rank, fixed-size outputs folded pairwise by Combine().
class TraverseCombinePST : public TraversePST {
public:
    explicit TraverseCombinePST(PST pst, int service_id, int nInBytes = 0, int nOutBytes = 0, const char *service_name = "")
        : TraversePST(pst, service_id, nInBytes, nOutBytes, service_name) {}
    virtual ~TraverseCombinePST() = default;

protected:
    virtual int Recurse(PST pst, void *vin, int nIn, void *vout, int nOut) final;
    virtual int Combine(void *vout, void *vout2, int nIn, int nOut1, int nOut2) = 0;
};

// Services that return one 64-bit count, summed over ranks.
class TraverseCountN : public TraverseCombinePST {
public:
    typedef uint64_t output;
    explicit TraverseCountN(PST pst, int service_id, int nInBytes, const char *service_name = "")
        : TraverseCombinePST(pst, service_id, nInBytes, sizeof(output), service_name) {}
    explicit TraverseCountN(PST pst, int service_id, const char *service_name = "")
        : TraverseCombinePST(pst, service_id, 0, sizeof(output), service_name) {}

protected:
    virtual int Combine(void *vout, void *vout2, int nIn, int nOut1, int nOut2) final;
    virtual int Service(PST pst, void *vin, int nIn, void *vout, int nOut) override = 0;
};
#endif
