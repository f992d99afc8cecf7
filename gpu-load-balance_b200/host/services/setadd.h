// setadd.h — ServiceSetAdd: organises the rank threads into the tree every other service walks.
//
// Replaces src/services/setadd.{h,cpp} of the reference.  master() calls it once with the interval of all ranks,
// the way the reference does (orbit.cpp:30-31 there):
//
//     ServiceSetAdd::input all(mdl->Threads());
//     mdl->RunService(PST_SETADD, sizeof(all), &all);
//
// The rank that receives [first, end) is `first`.  It halves the interval until it is alone: every upper half is
// handed to its own first rank (a request for this same service, so the subtrees are organised concurrently), every
// lower half becomes a new pstNode below the current one.  Afterwards node->idUpper / nLower / nUpper / pstLower
// describe, on every rank, the chain TraversePST::operator() follows.
#pragma once
#include "pst.h"

class ServiceSetAdd : public mdl::BasicService {
    PST root_;                              // this rank's (still unorganised) tree node, made in worker_init
    int operator()(int nBytesIn, void *in, void *out) override;

public:
    using output = void;
    struct input {                          // wire format: two ints, trivially copyable
        int idLower;                        // first rank of the interval == the rank that gets the request
        int idUpper;                        // one past the last rank
        input() = default;
        explicit input(int nRanks) : idLower(0), idUpper(nRanks) {}
        input(int first, int end) : idLower(first), idUpper(end) {}
    };

    explicit ServiceSetAdd(PST root) : BasicService(PST_SETADD, (int)sizeof(input), "SetAdd"), root_(root) {}
};
