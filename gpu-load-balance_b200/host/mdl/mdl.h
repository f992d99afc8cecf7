// mdl.h — a small, self-contained runtime exposing exactly the slice of the
// pkdgrav3 "mdl2" API that the ORB driver and its services use.
//
// The reference links the un-vendored mdl2 library for thread launch and
// request/reply transport (reference call sites: src/orbit.cpp:24-35,83-85,
// 291-324; src/services/TraversePST.cpp:3-44; src/services/setadd.cpp:19-43;
// src/services/pst.h:44-63).  No arithmetic of the hot path lives in mdl2, so
// this runtime replaces it with: one std::thread per rank (rank r drives GPU r
// in the B200 host), a per-thread service registry, and a mailbox per thread
// for ReqService/GetReply.  Thread 0 runs `master`.
//
// The number of threads comes from the environment variable ORB_MDL_THREADS
// (default 1), mirroring mdl2 where it is a runtime, not a compile-time, choice.
#ifndef ORB_MDL_H
#define ORB_MDL_H

#include <cassert>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

typedef void *MDL;

namespace mdl {

class mdlClass;

// Base class of every service: fixed id, maximum input/output sizes, and a
// call operator invoked by the runtime on the thread that owns the service.
class BasicService {
    friend class mdlClass;
    int service_id_;
    int max_in_bytes_;
    int max_out_bytes_;
    std::string name_;
public:
    explicit BasicService(int service_id, int nInBytes, int nOutBytes, const char *service_name = "")
        : service_id_(service_id), max_in_bytes_(nInBytes), max_out_bytes_(nOutBytes), name_(service_name) {}
    explicit BasicService(int service_id, int nInBytes, const char *service_name = "")
        : service_id_(service_id), max_in_bytes_(nInBytes), max_out_bytes_(0), name_(service_name) {}
    explicit BasicService(int service_id, const char *service_name = "")
        : service_id_(service_id), max_in_bytes_(0), max_out_bytes_(0), name_(service_name) {}
    virtual ~BasicService() = default;
    int getServiceID() const { return service_id_; }
    int getMaxBytesIn() const { return max_in_bytes_; }
    int getMaxBytesOut() const { return max_out_bytes_; }
    const std::string &getName() const { return name_; }
protected:
    virtual int operator()(int nIn, void *pIn, void *pOut) = 0;
};

struct Runtime;   // shared state of one mdlLaunch (mailboxes, thread table)

class mdlClass {
    friend struct Runtime;
    Runtime *rt_;
    int self_;
    std::map<int, std::unique_ptr<BasicService>> services_;
public:
    int argc;
    char **argv;
    void *worker_ctx;   // what worker_init returned for this thread

    mdlClass(Runtime *rt, int self, int argc_, char **argv_)
        : rt_(rt), self_(self), argc(argc_), argv(argv_), worker_ctx(nullptr) {}

    int Self() const { return self_; }
    int Threads() const;
    int Cores() const { return Threads(); }   // single process: every thread is a local core
    int Procs() const { return 1; }

    void AddService(std::unique_ptr<BasicService> &&service);
    BasicService *GetService(unsigned sid);

    // Run a service on this thread (the service itself fans out over the PST).
    int RunService(int sid, int nIn, void *pIn, void *pOut = nullptr);
    int RunService(int sid, void *pOut = nullptr) { return RunService(sid, 0, nullptr, pOut); }
    // Asynchronous request to another thread; returns a request id for GetReply.
    int ReqService(int target, int sid, void *pIn = nullptr, int nIn = 0);
    // Wait for the reply of a previous ReqService; copies the output, returns its size in bytes.
    int GetReply(int rID, void *pOut = nullptr);

    // used by the worker loop
    int Dispatch(int sid, int nIn, void *pIn, void *pOut);
};

// Optional observer, called on thread 0 after every top-level RunService
// returns.  The parity harness uses it to dump what the reference never
// writes out ("no result is ever written").
typedef std::function<void(mdlClass *mdl, int sid, int nIn, void *pIn, void *pOut, int nOut)> RunServiceTap;
void setRunServiceTap(RunServiceTap tap);

}  // namespace mdl

// ---- C-style calls used by pst.h / setadd.cpp of the reference ----
int mdlSelf(MDL mdl);
int mdlThreads(MDL mdl);
int mdlCores(MDL mdl);
int mdlThreadToProc(MDL mdl, int iThread);
int mdlProcToThread(MDL mdl, int iProc);
int mdlReqService(MDL mdl, int id, int sid, void *vin, int nInBytes);
void mdlGetReply(MDL mdl, int rID, void *vout, int *pnOut);
#define mdlassert(mdl, expr)                                                                   \
    do {                                                                                       \
        if (!(expr)) {                                                                         \
            std::fprintf(stderr, "mdlassert failed: %s at %s:%d\n", #expr, __FILE__, __LINE__); \
            std::abort();                                                                      \
        }                                                                                      \
    } while (0)

// Launches ORB_MDL_THREADS threads; `worker_init` runs on every thread and
// returns that thread's context, `master` runs on thread 0 with its context,
// `worker_done` runs on every thread afterwards.  Returns master's result.
int mdlLaunch(int argc, char **argv, int (*master)(MDL, void *), void *(*worker_init)(MDL),
              void (*worker_done)(MDL, void *));

#endif
