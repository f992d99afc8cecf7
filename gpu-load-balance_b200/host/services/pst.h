// pst.h — per-rank state and the service id table (reference: src/services/pst.h).
// One mdl thread = one rank = one GPU.  LocalData shrinks to the host copy of the rank's particle
// slice plus an opaque device context: every device buffer the reference listed here
// (pst.h:13-38) now lives inside liborb_b200.so.
#ifndef ORB_HOST_PST_H
#define ORB_HOST_PST_H

#include <vector>

#include "mdl.h"
#include "../../../include/orb_b200.h"

class LocalData {
public:
    // host SoA columns of this rank's slice (the reference's column-major blitz (N,3) array, init.cu:32-45)
    std::vector<float> x, y, z;
    unsigned long long firstParticle = 0;   // offset of the slice in the single generator stream
    int nParticles = 0;
    int nLeafCells = 0;
    orb_ctx *ctx = nullptr;                 // device side of this rank (created in ServiceInit)
    int rank = 0, nRanks = 1;
    LocalData() = default;
};

class pstNode {
public:
    pstNode *pstLower;
    LocalData *lcl;
    MDL mdl;
    int idSelf;
    int idUpper;
    int nLeaves;
    int nLower;
    int nUpper;
    explicit pstNode(MDL mdl_)
        : pstLower(nullptr), lcl(nullptr), mdl(mdl_), idSelf(mdlSelf(mdl_)), idUpper(0), nLeaves(1), nLower(0), nUpper(0) {}
    // a subtree either fits the cores of this process or it does not; with one process all of it is "on node"
    bool OffNode() const { return nLeaves > mdlCores(mdl); }
    bool OnNode() const { return !OffNode(); }
    bool AmNode() const { return nLeaves == mdlCores(mdl); }
    bool NotNode() const { return !AmNode(); }
    bool AmCore() const { return nLeaves == 1; }
    bool NotCore() const { return !AmCore(); }
};
typedef pstNode *PST;

// Same ids as the reference (pst.h:66-81) so service numbers in logs/scripts keep their meaning;
// new services are appended.
enum pst_service {
    PST_SRV_STOP = 0,
    PST_SETADD,
    PST_INIT,
    PST_INITGPU,
    PST_COPYPARTICLES,
    PST_COPYCELLS,
    PST_COUNTLEFTGPU,
    PST_COUNTLEFTAXISGPU,
    PST_COUNTLEFT,
    PST_PARTITION,
    PST_PARTITIONGPU,
    PST_COUNT,
    PST_FINALIZE,
    PST_MAKEAXIS,
    // --- additions of the B200 host ---
    PST_BBOX,        // per-cell particle bounding boxes (north-star extension)
    PST_FINDCUTS,    // whole bisection loop of a level on the device
    PST_BUILD,       // whole ORB build on the device
    PST_DUMP         // parity dump: ranges + particle hashes of this rank
};
#endif
