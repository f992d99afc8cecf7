// TraversePST.cpp — PST walk (reference: src/services/TraversePST.cpp:3-57).  The sibling's reply is
// received into a heap buffer (the reference uses alloca(nOut), which cannot hold the count vectors of
// 2^20-cell levels).
#include "TraversePST.h"

#include <vector>

int TraversePST::operator()(int nIn, void *pIn, void *pOut) { return Traverse(node_pst, pIn, nIn, pOut, getMaxBytesOut()); }

int TraversePST::Traverse(PST pst, void *vin, int nIn, void *vout, int nOut) {
    if (pst->AmCore()) return Service(pst, vin, nIn, vout, nOut);
    if (pst->OffNode()) return OffNode(pst, vin, nIn, vout, nOut);
    if (pst->AmNode()) return AtNode(pst, vin, nIn, vout, nOut);
    return Recurse(pst, vin, nIn, vout, nOut);
}

// run another service over the same subtree
int TraversePST::Traverse(unsigned sid, PST pst, void *vin, int nIn, void *vout, int nOut) {
    auto *m = static_cast<mdl::mdlClass *>(pst->mdl);
    auto *svc = dynamic_cast<TraversePST *>(m->GetService(sid));
    mdlassert(pst->mdl, svc != nullptr);
    return svc->Traverse(pst, vin, nIn, vout, nOut);
}

int TraversePST::Recurse(PST pst, void *vin, int nIn, void *vout, int nOut) {
    auto *m = static_cast<mdl::mdlClass *>(pst->mdl);
    const int request = m->ReqService(pst->idUpper, getServiceID(), vin, nIn);
    Traverse(pst->pstLower, vin, nIn, vout, nOut);
    return m->GetReply(request, vout);
}

int TraverseCombinePST::Recurse(PST pst, void *vin, int nIn, void *vout, int nOut) {
    auto *m = static_cast<mdl::mdlClass *>(pst->mdl);
    const int request = m->ReqService(pst->idUpper, getServiceID(), vin, nIn);
    Traverse(pst->pstLower, vin, nIn, vout, nOut);
    std::vector<char> sibling((size_t)(nOut > 0 ? nOut : 1));
    const int nOut2 = m->GetReply(request, sibling.data());
    return Combine(vout, sibling.data(), nIn, nOut, nOut2);
}

int TraverseCountN::Combine(void *vout, void *vout2, int, int, int) {
    *static_cast<output *>(vout) += *static_cast<output *>(vout2);
    return sizeof(output);
}
