// TraversePST.h — how one service call spreads over the rank threads.
//
// Replaces src/services/TraversePST.{h,cpp} of the reference.  What service code sees is unchanged — it derives
// from TraverseCombinePST(pst, id, maxInBytes, maxOutBytes, name) and overrides Service() and Combine() with the
// reference's signatures (TraversePST.h:37-46 there) — but the machinery underneath is a single loop rather than the
// reference's mutually recursive Traverse / Recurse / OffNode / AtNode hooks: this host is one process with one
// thread per GPU, so a node of the rank tree (pst.h) is either a leaf or it is not.
#pragma once
#include "pst.h"

// Entry point the mdl runtime calls on the rank that received the request.  The call walks this rank's chain of
// lower halves; on the way down every upper half gets the same request, at the bottom Service() does this rank's
// work, on the way back up the upper halves' replies are folded in, innermost first.
class TraversePST : public mdl::BasicService {
public:
    TraversePST(PST pst, int service_id, int nInBytes = 0, int nOutBytes = 0, const char *service_name = "")
        : BasicService(service_id, nInBytes, nOutBytes, service_name), tree_(pst) {}
    ~TraversePST() override = default;

protected:
    int operator()(int nIn, void *pIn, void *pOut) final;

    // this rank's share; returns the bytes written to vout (capacity nOut = the service's maxOutBytes)
    virtual int Service(PST pst, void *vin, int nIn, void *vout, int nOut) = 0;

    // fold the reply of an upper half (nOut2 bytes at vout2) into vout (capacity nOut1); returns the size of the result.
    // Default: the replies only signal completion.
    virtual int Combine(void *vout, void *vout2, int nIn, int nOut1, int nOut2) {
        (void)vout; (void)vout2; (void)nIn; (void)nOut2;
        return nOut1;
    }

    PST tree() const { return tree_; }

private:
    PST tree_;     // this rank's root of the rank tree (ServiceSetAdd hangs the lower halves below it)
};

// Same input on every rank, fixed-size outputs folded pairwise: the only kind of service the ORB path has.
class TraverseCombinePST : public TraversePST {
public:
    explicit TraverseCombinePST(PST pst, int service_id, int nInBytes = 0, int nOutBytes = 0, const char *service_name = "")
        : TraversePST(pst, service_id, nInBytes, nOutBytes, service_name) {}

protected:
    int Combine(void *vout, void *vout2, int nIn, int nOut1, int nOut2) override = 0;
};
