// services.h — the ORB services of the B200 host.  Class names, service ids, wire types
// (`Cell[nCells]` in, `unsigned int[nCells]` out) and the Service()/Combine() split follow the
// reference's src/services/*.h one to one; each class cites the header it stands in for.  Every
// Service() body is a thin call into the C ABI (include/orb_b200.h) on this rank's device context.
//
// Combine(): the reference sums per-thread outputs on the way up the PST (countLeft.cpp:44-53,
// count.cpp:21-30).  Here the device already holds the sum over ranks (NCCL allreduce in-stream,
// inside the C ABI), so every rank returns the same global vector and Combine() just keeps one copy.
#ifndef ORB_HOST_SERVICES_H
#define ORB_HOST_SERVICES_H

#include "../cell.h"
#include "../constants.h"
#include "TraversePST.h"

// ---- init.h:5-20 ----
class ServiceInit : public TraverseCombinePST {
public:
    struct input {
        int nParticles;     // per rank (N / Threads, orbit.cpp:83)
        int d;
        bool generate;      // false: read positions from the tipsy file named by ORB_TIPSY (init.cu:54-59)
        META_PARAMS params;
        int nTotal;         // generate == false: bodies taken from the file, split in contiguous slices over the ranks
    };
    typedef int output;
    explicit ServiceInit(PST pst) : TraverseCombinePST(pst, PST_INIT, sizeof(input), sizeof(output), "Init") {}
protected:
    virtual int Service(PST pst, void *vin, int nIn, void *vout, int nOut);
    virtual int Combine(void *vout, void *vout2, int nIn, int nOut1, int nOut2);
};

// ---- copyParticles.h:5-18 : host -> device upload of this rank's x,y,z (once) ----
class ServiceCopyParticles : public TraverseCombinePST {
public:
    struct input { META_PARAMS params; };
    typedef int output;
    explicit ServiceCopyParticles(PST pst)
        : TraverseCombinePST(pst, PST_COPYPARTICLES, sizeof(input), sizeof(output), "CopyToDevice") {}
protected:
    virtual int Service(PST pst, void *vin, int nIn, void *vout, int nOut);
    virtual int Combine(void *vout, void *vout2, int nIn, int nOut1, int nOut2);
};

// ---- copyCells.h:5-16 : kept as a service id; the cell table now rides along with each call ----
class ServiceCopyCells : public TraverseCombinePST {
public:
    typedef struct Cell input;
    typedef int output;
    explicit ServiceCopyCells(PST pst)
        : TraverseCombinePST(pst, PST_COPYCELLS, sizeof(input) * MAX_CELLS, sizeof(output), "CopyCells") {}
protected:
    virtual int Service(PST pst, void *vin, int nIn, void *vout, int nOut);
    virtual int Combine(void *vout, void *vout2, int nIn, int nOut1, int nOut2);
};

// ---- count.h:5-16 ----
class ServiceCount : public TraverseCombinePST {
public:
    typedef struct Cell input;
    typedef unsigned int output;
    explicit ServiceCount(PST pst)
        : TraverseCombinePST(pst, PST_COUNT, MAX_CELLS * sizeof(input), MAX_CELLS * sizeof(output), "Count") {}
protected:
    virtual int Service(PST pst, void *vin, int nIn, void *vout, int nOut);
    virtual int Combine(void *vout, void *vout2, int nIn, int nOut1, int nOut2);
};

// ---- countLefGPU.h:5-16 (o=2 flavour) ----
class ServiceCountLeftGPU : public TraverseCombinePST {
public:
    typedef struct Cell input;
    typedef unsigned int output;
    explicit ServiceCountLeftGPU(PST pst)
        : TraverseCombinePST(pst, PST_COUNTLEFTGPU, MAX_CELLS * sizeof(input), MAX_CELLS * sizeof(output), "CountLeftGPU") {}
protected:
    virtual int Service(PST pst, void *vin, int nIn, void *vout, int nOut);
    virtual int Combine(void *vout, void *vout2, int nIn, int nOut1, int nOut2);
};

// ---- countLeftGPUAxis.h:4-14 (o=1 flavour; same kernel now, no MakeAxis copy needed) ----
class ServiceCountLeftGPUAxis : public TraverseCombinePST {
public:
    typedef struct Cell input;
    typedef unsigned int output;
    explicit ServiceCountLeftGPUAxis(PST pst)
        : TraverseCombinePST(pst, PST_COUNTLEFTAXISGPU, MAX_CELLS * sizeof(input), MAX_CELLS * sizeof(output), "CountLeftAxis") {}
protected:
    virtual int Service(PST pst, void *vin, int nIn, void *vout, int nOut);
    virtual int Combine(void *vout, void *vout2, int nIn, int nOut1, int nOut2);
};

// ---- countLeft.h:5-15 (the CPU service id of o=0): the unmodified control flow of orbit.cpp:154-189 keeps working, the
//      count itself runs on the device like the two GPU flavours ----
class ServiceCountLeft : public TraverseCombinePST {
public:
    typedef struct Cell input;
    typedef unsigned int output;
    explicit ServiceCountLeft(PST pst)
        : TraverseCombinePST(pst, PST_COUNTLEFT, MAX_CELLS * sizeof(input), MAX_CELLS * sizeof(output), "CountLeft") {}
protected:
    virtual int Service(PST pst, void *vin, int nIn, void *vout, int nOut);
    virtual int Combine(void *vout, void *vout2, int nIn, int nOut1, int nOut2);
};

// ---- partition.h:5-15 (the CPU service id of o=0 / o=1, orbit.cpp:263-273) ----
class ServicePartition : public TraverseCombinePST {
public:
    typedef struct Cell input;
    typedef int output;
    explicit ServicePartition(PST pst)
        : TraverseCombinePST(pst, PST_PARTITION, MAX_CELLS * sizeof(input), sizeof(output), "Reshuffle") {}
protected:
    virtual int Service(PST pst, void *vin, int nIn, void *vout, int nOut);
    virtual int Combine(void *vout, void *vout2, int nIn, int nOut1, int nOut2);
};

// ---- makeAxis.h:5-15: the axis-column gather is designed out (the kernels read x/y/z by per-cell axis); the id stays
//      registered so that orbit.cpp:113-121 can call it ----
class ServiceMakeAxis : public TraverseCombinePST {
public:
    typedef struct Cell input;
    typedef int output;
    explicit ServiceMakeAxis(PST pst)
        : TraverseCombinePST(pst, PST_MAKEAXIS, MAX_CELLS * sizeof(input), sizeof(output), "MakeAxis") {}
protected:
    virtual int Service(PST pst, void *vin, int nIn, void *vout, int nOut);
    virtual int Combine(void *vout, void *vout2, int nIn, int nOut1, int nOut2);
};

// ---- partitionGPU.h:4-14 ----
class ServicePartitionGPU : public TraverseCombinePST {
public:
    typedef struct Cell input;
    typedef int output;
    explicit ServicePartitionGPU(PST pst)
        : TraverseCombinePST(pst, PST_PARTITIONGPU, MAX_CELLS * sizeof(input), sizeof(output), "PartitionGPU") {}
protected:
    virtual int Service(PST pst, void *vin, int nIn, void *vout, int nOut);
    virtual int Combine(void *vout, void *vout2, int nIn, int nOut1, int nOut2);
};

// ---- finalize.h:5-18 : particles back to the host copy, device context released ----
class ServiceFinalize : public TraverseCombinePST {
public:
    struct input { META_PARAMS params; };
    typedef int output;
    explicit ServiceFinalize(PST pst) : TraverseCombinePST(pst, PST_FINALIZE, sizeof(input), sizeof(output), "Finalize") {}
protected:
    virtual int Service(PST pst, void *vin, int nIn, void *vout, int nOut);
    virtual int Combine(void *vout, void *vout2, int nIn, int nOut1, int nOut2);
};

// ---- new: bounding boxes (north star; the reference only has a commented-out Bound in pst.h:43,54) ----
class ServiceBBox : public TraverseCombinePST {
public:
    typedef struct Cell input;
    struct output { float lower[3], upper[3]; };
    explicit ServiceBBox(PST pst)
        : TraverseCombinePST(pst, PST_BBOX, MAX_CELLS * sizeof(input), MAX_CELLS * sizeof(output), "BBox") {}
protected:
    virtual int Service(PST pst, void *vin, int nIn, void *vout, int nOut);
    virtual int Combine(void *vout, void *vout2, int nIn, int nOut1, int nOut2);
};

// ---- new: the whole bisection loop of a level (orbit.cpp:146-232) without host round trips ----
class ServiceFindCuts : public TraverseCombinePST {
public:
    typedef struct Cell input;
    typedef struct Cell output;   // same cells with final margins / foundCut
    explicit ServiceFindCuts(PST pst)
        : TraverseCombinePST(pst, PST_FINDCUTS, MAX_CELLS * sizeof(input), MAX_CELLS * sizeof(output), "FindCuts") {}
protected:
    virtual int Service(PST pst, void *vin, int nIn, void *vout, int nOut);
    virtual int Combine(void *vout, void *vout2, int nIn, int nOut1, int nOut2);
};

// ---- new: the whole build (orbit.cpp:74-275) on the device; the heap lands in rank 0's memory ----
class ServiceBuild : public TraverseCombinePST {
public:
    struct input { unsigned int flags; Cell *heapOut; };   // heapOut: thread 0's buffer of 2d-1 cells (same address space)
    typedef orb_build_stats output;
    explicit ServiceBuild(PST pst) : TraverseCombinePST(pst, PST_BUILD, sizeof(input), sizeof(output), "Build") {}
protected:
    virtual int Service(PST pst, void *vin, int nIn, void *vout, int nOut);
    virtual int Combine(void *vout, void *vout2, int nIn, int nOut1, int nOut2);
};

// ---- new: parity dump (the reference never writes a result) ----
class ServiceDump : public TraverseCombinePST {
public:
    struct input { char path[480]; };
    typedef int output;
    explicit ServiceDump(PST pst) : TraverseCombinePST(pst, PST_DUMP, sizeof(input), sizeof(output), "Dump") {}
protected:
    virtual int Service(PST pst, void *vin, int nIn, void *vout, int nOut);
    virtual int Combine(void *vout, void *vout2, int nIn, int nOut1, int nOut2);
};

#endif
