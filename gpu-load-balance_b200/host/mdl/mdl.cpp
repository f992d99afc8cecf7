// mdl.cpp — thread-per-rank runtime behind mdl.h (see the header for scope and
// the reference call sites it stands in for).
#include "mdl.h"

#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>

namespace mdl {

namespace {
// function-local static: safe to set from another translation unit's static initialiser
RunServiceTap &tapRef() {
    static RunServiceTap tap;
    return tap;
}
}

void setRunServiceTap(RunServiceTap tap) { tapRef() = std::move(tap); }

// One mailbox per thread: at most one request is outstanding per target,
// because a thread only ever receives requests from its parent in the PST.
struct Mailbox {
    std::mutex m;
    std::condition_variable cv;
    enum { IDLE, REQUEST, REPLY } state = IDLE;
    int sid = 0;
    int nIn = 0;
    int nOut = 0;
    bool stop = false;
    std::vector<char> in, out;
};

struct Runtime {
    int nThreads = 1;
    std::vector<std::unique_ptr<Mailbox>> box;
    std::vector<std::unique_ptr<mdlClass>> mdl;

    // startup barrier: master must not issue requests before every thread registered its services
    std::mutex m;
    std::condition_variable cv;
    int nReady = 0;

    void workerLoop(int self);
};

int mdlClass::Threads() const { return rt_->nThreads; }

void mdlClass::AddService(std::unique_ptr<BasicService> &&service) {
    int sid = service->getServiceID();
    services_[sid] = std::move(service);
}

BasicService *mdlClass::GetService(unsigned sid) {
    auto it = services_.find((int)sid);
    return it == services_.end() ? nullptr : it->second.get();
}

int mdlClass::Dispatch(int sid, int nIn, void *pIn, void *pOut) {
    BasicService *svc = GetService((unsigned)sid);
    if (!svc) {
        std::fprintf(stderr, "mdl: thread %d has no service %d\n", self_, sid);
        std::abort();
    }
    return (*svc)(nIn, pIn, pOut);
}

int mdlClass::RunService(int sid, int nIn, void *pIn, void *pOut) {
    int nOut = Dispatch(sid, nIn, pIn, pOut);
    if (self_ == 0 && tapRef()) tapRef()(this, sid, nIn, pIn, pOut, nOut);
    return nOut;
}

int mdlClass::ReqService(int target, int sid, void *pIn, int nIn) {
    Mailbox &b = *rt_->box[target];
    std::unique_lock<std::mutex> lk(b.m);
    b.cv.wait(lk, [&] { return b.state == Mailbox::IDLE; });
    b.sid = sid;
    b.nIn = nIn;
    b.in.resize((size_t)nIn);
    if (nIn) std::memcpy(b.in.data(), pIn, (size_t)nIn);
    b.state = Mailbox::REQUEST;
    b.cv.notify_all();
    return target;
}

int mdlClass::GetReply(int rID, void *pOut) {
    Mailbox &b = *rt_->box[rID];
    std::unique_lock<std::mutex> lk(b.m);
    b.cv.wait(lk, [&] { return b.state == Mailbox::REPLY; });
    int nOut = b.nOut;
    if (pOut && nOut > 0) std::memcpy(pOut, b.out.data(), (size_t)nOut);
    b.state = Mailbox::IDLE;
    b.cv.notify_all();
    return nOut;
}

void Runtime::workerLoop(int self) {
    Mailbox &b = *box[self];
    mdlClass *me = mdl[self].get();
    for (;;) {
        int sid, nIn;
        {
            std::unique_lock<std::mutex> lk(b.m);
            b.cv.wait(lk, [&] { return b.state == Mailbox::REQUEST || b.stop; });
            if (b.stop && b.state != Mailbox::REQUEST) return;
            sid = b.sid;
            nIn = b.nIn;
        }
        BasicService *svc = me->GetService((unsigned)sid);
        size_t cap = svc ? (size_t)svc->getMaxBytesOut() : 0;
        if (b.out.size() < cap) b.out.resize(cap);
        // The request buffer is only touched by this thread until REPLY is posted.
        int nOut = me->Dispatch(sid, nIn, b.in.data(), b.out.data());
        {
            std::unique_lock<std::mutex> lk(b.m);
            b.nOut = nOut;
            b.state = Mailbox::REPLY;
            b.cv.notify_all();
        }
    }
}

}  // namespace mdl

using mdl::mdlClass;

int mdlSelf(MDL m) { return static_cast<mdlClass *>(m)->Self(); }
int mdlThreads(MDL m) { return static_cast<mdlClass *>(m)->Threads(); }
int mdlCores(MDL m) { return static_cast<mdlClass *>(m)->Cores(); }
int mdlThreadToProc(MDL, int) { return 0; }
int mdlProcToThread(MDL, int) { return 0; }
int mdlReqService(MDL m, int id, int sid, void *vin, int nInBytes) {
    return static_cast<mdlClass *>(m)->ReqService(id, sid, vin, nInBytes);
}
void mdlGetReply(MDL m, int rID, void *vout, int *pnOut) {
    int n = static_cast<mdlClass *>(m)->GetReply(rID, vout);
    if (pnOut) *pnOut = n;
}

int mdlLaunch(int argc, char **argv, int (*master)(MDL, void *), void *(*worker_init)(MDL),
              void (*worker_done)(MDL, void *)) {
    mdl::Runtime rt;
    const char *env = std::getenv("ORB_MDL_THREADS");
    rt.nThreads = env ? std::atoi(env) : 1;
    if (rt.nThreads < 1) rt.nThreads = 1;
    for (int t = 0; t < rt.nThreads; ++t) {
        rt.box.emplace_back(new mdl::Mailbox());
        rt.mdl.emplace_back(new mdlClass(&rt, t, argc, argv));
    }

    std::vector<std::thread> workers;
    for (int t = 1; t < rt.nThreads; ++t) {
        workers.emplace_back([&rt, t, worker_init, worker_done] {
            mdlClass *me = rt.mdl[t].get();
            me->worker_ctx = worker_init(me);
            {
                std::unique_lock<std::mutex> lk(rt.m);
                rt.nReady++;
                rt.cv.notify_all();
            }
            rt.workerLoop(t);
            worker_done(me, me->worker_ctx);
        });
    }

    mdlClass *me0 = rt.mdl[0].get();
    me0->worker_ctx = worker_init(me0);
    {
        std::unique_lock<std::mutex> lk(rt.m);
        rt.cv.wait(lk, [&] { return rt.nReady == rt.nThreads - 1; });
    }
    int rc = master(me0, me0->worker_ctx);

    for (int t = 1; t < rt.nThreads; ++t) {
        mdl::Mailbox &b = *rt.box[t];
        std::unique_lock<std::mutex> lk(b.m);
        b.stop = true;
        b.cv.notify_all();
    }
    for (auto &w : workers) w.join();
    worker_done(me0, me0->worker_ctx);
    return rc;
}
