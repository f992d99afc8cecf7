// TraversePST.cpp — the walk over the rank tree (see TraversePST.h).
// The replies of the upper halves are received into one heap buffer of maxOutBytes (the reference receives each
// into alloca(nOut), which cannot hold the per-cell vectors of levels with 2^20 cells).
#include "TraversePST.h"

#include <algorithm>
#include <memory>
#include <vector>

int TraversePST::operator()(int nIn, void *pIn, void *pOut) {
    auto *rt = static_cast<mdl::mdlClass *>(tree_->mdl);
    const int capacity = getMaxBytesOut();

    // down: leave the request with the first rank of every upper half, keep to the lower halves
    std::vector<int> outstanding;
    PST node = tree_;
    while (node->NotCore()) {
        outstanding.push_back(rt->ReqService(node->idUpper, getServiceID(), pIn, nIn));
        node = node->pstLower;
    }

    // bottom: this rank's own work
    int nOut = Service(node, pIn, nIn, pOut, capacity);

    // up: the most recently asked upper half is the sibling of the node we just finished.  (The receive buffer is only
    // made when there is something to receive: a single rank never pays for maxOutBytes, which is tens of MB for the
    // per-cell services at 2^20 cells.)
    if (!outstanding.empty()) {
        std::unique_ptr<char[]> theirs(new char[(size_t)std::max(capacity, 1)]);
        while (!outstanding.empty()) {
            const int nTheirs = rt->GetReply(outstanding.back(), theirs.get());
            outstanding.pop_back();
            nOut = Combine(pOut, theirs.get(), nIn, capacity, nTheirs);
        }
    }
    return nOut;
}
