// setadd.cpp — split the rank interval [idLower, idUpper) in halves until every PST leaf is one rank.
// Single process, so no "keep a process together" adjustment is needed (reference setadd.cpp:24-30).
#include "setadd.h"

#include <type_traits>
static_assert(std::is_trivial<ServiceSetAdd::input>(), "service inputs travel by memcpy");

int ServiceSetAdd::operator()(int nIn, void *pIn, void *) {
    mdlassert(node_pst->mdl, nIn == (int)sizeof(input));
    SetAdd(node_pst, static_cast<input *>(pIn));
    return 0;
}

void ServiceSetAdd::SetAdd(PST pst, input *in) {
    mdlassert(pst->mdl, pst->nLeaves == 1);
    mdlassert(pst->mdl, in->idLower == mdlSelf(pst->mdl));
    const int span = in->idUpper - in->idLower;
    if (span <= 1) return;
    const int middle = (in->idUpper + in->idLower) / 2;
    pst->nLeaves += span - 1;
    pst->nLower = middle - in->idLower;
    pst->nUpper = in->idUpper - middle;
    pst->idUpper = middle;
    input upper(middle, in->idUpper);          // the upper half builds its own subtree ...
    const int request = mdlReqService(pst->mdl, pst->idUpper, getServiceID(), &upper, sizeof(upper));
    input lower(mdlSelf(pst->mdl), middle);    // ... while this thread descends into the lower half
    pst->pstLower = new pstNode(pst->mdl);
    pst->pstLower->lcl = pst->lcl;
    SetAdd(pst->pstLower, &lower);
    mdlGetReply(pst->mdl, request, nullptr, nullptr);
}
