// TEST INFRASTRUCTURE ONLY (oracle/): stand-in for the reference's missing
// src/tipsy/tipsy.h (init.cu:5 includes it; the sources are absent from the
// checkout, CMakeLists.txt:74).  The branch that uses it (init.cu:54-59,
// generate == false) is dead: master always passes generate=true
// (orbit.cpp:83).  This stub only lets init.cu compile unmodified.
#ifndef ORB_REF_SHIM_TIPSY_H
#define ORB_REF_SHIM_TIPSY_H
#include <cstdio>
#include <cstdlib>
#include <blitz/array.h>
class TipsyIO {
public:
    void open(const char *path) {
        std::fprintf(stderr, "TipsyIO stub: cannot open %s (tipsy sources are not part of the reference checkout)\n", path);
        std::abort();
    }
    int count() const { return 0; }
    void load(blitz::Array<float, 2> &) {}
};
#endif
