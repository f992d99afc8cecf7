// constants.h — compile-time knobs and the flag struct shared by the services
// (reference: src/constants.h:8-23).  MAX_CELLS is no longer a hard cap on the tree (the reference
// silently overflows for y >= 13): it only sizes the fixed service buffers and is large
// enough for 2^20 leaf cells; device storage is sized from the actual d at Init.
#ifndef ORB_HOST_CONSTANTS_H
#define ORB_HOST_CONSTANTS_H

#include <cstdio>
#include <cstdlib>

#include "../../include/orb_b200.h"

static const int MAX_CELLS = (1 << 21);     // ids run to 2d-2; the widest level handed to a service has d cells
static const int N_STREAMS = 1;

struct META_PARAMS {
    bool GPU_COUNT;
    bool GPU_PARTITION;
    bool FAST_MEDIAN;     // never enabled by the reference (orbit.cpp:54,59,65); kept for layout
};

// The C ABI returns status codes and never exits; the services keep the reference's
// abort-on-error behaviour (CUDA_CHECK in constants.h:19-23: message to stderr, exit(1)).
#define ORB_CHECK(call)                                                                                     \
    do {                                                                                                    \
        int orb_rc_ = (call);                                                                               \
        if (orb_rc_ != 0) {                                                                                 \
            std::fprintf(stderr, "%s error %d in %s(%d)\n%s\n", #call, orb_rc_, __FILE__, __LINE__,        \
                         orb_last_error());                                                                 \
            std::exit(1);                                                                                   \
        }                                                                                                   \
    } while (0)

#endif
