// setadd.cpp — halve the rank interval until this rank is alone.
// Every halving hands the upper half to its first rank (a request for this same service, so that rank organises its
// own subtree concurrently) and hangs a new node for the lower half below the current one.  One process, so the
// reference's adjustment that keeps a process's threads in one subtree (setadd.cpp:24-30 there) has nothing to do.
#include "setadd.h"

#include <type_traits>
#include <vector>

static_assert(std::is_trivial<ServiceSetAdd::input>::value, "service inputs travel by memcpy");

int ServiceSetAdd::operator()(int nIn, void *pIn, void *) {
    MDL mdl = root_->mdl;
    mdlassert(mdl, nIn == (int)sizeof(input));
    const input all = *static_cast<const input *>(pIn);
    mdlassert(mdl, all.idLower == mdlSelf(mdl));

    std::vector<int> outstanding;
    PST node = root_;
    for (int first = all.idLower, end = all.idUpper; end - first > 1;) {
        mdlassert(mdl, node->nLeaves == 1);               // not organised yet
        const int middle = (first + end) / 2;
        node->nLeaves = end - first;
        node->nLower = middle - first;
        node->nUpper = end - middle;
        node->idUpper = middle;
        input upperHalf(middle, end);
        outstanding.push_back(mdlReqService(mdl, middle, getServiceID(), &upperHalf, sizeof(upperHalf)));
        node->pstLower = new pstNode(mdl);
        node->pstLower->lcl = node->lcl;
        node = node->pstLower;
        end = middle;
    }
    for (auto it = outstanding.rbegin(); it != outstanding.rend(); ++it) mdlGetReply(mdl, *it, nullptr, nullptr);
    return 0;
}
