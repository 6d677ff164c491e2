// runtime.cu -- context, handle registry, error strings, host<->device field staging and
// the NCCL all-reduce used for the grid moments (loaded lazily with dlopen so that the
// single-GPU path has no NCCL dependency and a host process that already carries a
// libnccl.so.2 -- e.g. torch's -- shares it).
#include <dlfcn.h>
#include <nccl.h>

#include <condition_variable>
#include <algorithm>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <thread>

#include "common.cuh"

namespace gempic {

static thread_local std::string g_error;

void set_error(const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_error = buf;
}

void fail(int code, const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_error = buf;
    throw Fail{code};
}

// ---- ranks ------------------------------------------------------------------------------
// All mutable library state lives in a Rank: the context (device, stream, communicator), the handle registry, the
// pinned staging cursor and the profiling slots.  A process normally has ONE rank (gempic_init).  After
// gempic_init_devices(n, ids) it has n, one per device, each driven by its own worker thread (md.cu); every API call
// is then executed by all of them -- the in-process form of "one process per GPU".  `t_rank` is the rank the calling
// thread works for.
struct ProfSlot {
    std::string tag;
    double ms = 0.0;
    int64_t launches = 0;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;
};
struct Rank {
    Context ctx;
    std::mutex mu;
    std::unordered_map<gempic_handle, std::unique_ptr<Object>> objects;
    std::vector<std::unique_ptr<Object>> zombies;   // destroyed handles that a splitting object still points to
    gempic_handle next = 0x1000;
    size_t pinned_off = 0;
    bool profile = false;
    std::vector<ProfSlot> prof;
    std::vector<cudaEvent_t> event_pool;
    std::unordered_map<const void *, size_t> func_smem;
    bool host_out = true;
    ~Rank();   // ordered teardown at process exit when gempic_finalize was not called
};
static Rank g_rank0;
static thread_local Rank *t_rank = &g_rank0;
static Rank &rk() { return *t_rank; }
#define g_ctx (rk().ctx)
#define g_mu (rk().mu)
#define g_objects (rk().objects)
#define g_next (rk().next)
#define g_zombies (rk().zombies)
#define g_pinned_off (rk().pinned_off)
#define g_profile (rk().profile)
#define g_prof (rk().prof)
#define g_event_pool (rk().event_pool)

Context &ctx() { return g_ctx; }
bool host_out_enabled() { return rk().host_out; }
namespace md { static bool g_in_process_peers = false; }   // exchange buffers are plain peer pointers (gempic_init_devices)

void require_init()
{
    GP_REQUIRE(g_ctx.ready, GEMPIC_ENOTINIT, "gempic_init() has not been called (or no CUDA device)");
}

void ensure_func_smem(const void *func, size_t bytes)
{
    size_t &have = rk().func_smem[func];
    if (bytes > have) {
        GP_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        have = bytes;
    }
}

// ---- registry -------------------------------------------------------------------------

gempic_handle register_object(std::unique_ptr<Object> obj)
{
    std::lock_guard<std::mutex> lk(g_mu);
    gempic_handle h = g_next++;
    g_objects[h] = std::move(obj);
    return h;
}

Object *lookup(gempic_handle h, Kind kind, const char *what)
{
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_objects.find(h);
    if (it == g_objects.end() || it->second->kind != kind)
        fail(GEMPIC_EHANDLE, "invalid %s handle 0x%llx", what, (unsigned long long)h);
    return it->second.get();
}

void retain(Object *o)
{
    std::lock_guard<std::mutex> lk(g_mu);
    o->users++;
}

void release(Object *o)
{
    std::unique_ptr<Object> victim;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        if (--o->users > 0) return;
        for (auto it = g_zombies.begin(); it != g_zombies.end(); ++it)
            if (it->get() == o) {
                victim = std::move(*it);
                g_zombies.erase(it);
                break;
            }
    }
}

void destroy(gempic_handle h, Kind kind, const char *what)
{
    std::unique_ptr<Object> victim;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_objects.find(h);
        if (it == g_objects.end() || it->second->kind != kind)
            fail(GEMPIC_EHANDLE, "invalid %s handle 0x%llx", what, (unsigned long long)h);
        victim = std::move(it->second);
        g_objects.erase(it);
        if (victim->users > 0) g_zombies.push_back(std::move(victim));
    }
    if (g_ctx.ready) cudaStreamSynchronize(g_ctx.stream);
}   // `victim` (and, through its destructor, the releases of what it retained) goes here, outside the lock

// A process may exit without gempic_finalize (a script that simply ends).  The registry then goes down with the static
// Rank: the splitting objects must go first (their destructors release what they point to, which walks this rank's
// zombie list), and release() must find THIS rank whatever thread runs the destructor.
Rank::~Rank()
{
    Rank *saved = t_rank;
    t_rank = this;
    try {
        destroy_all();
    } catch (...) {
    }
    t_rank = saved == this ? &g_rank0 : saved;
}

void destroy_all()
{
    // the splitting objects first: their destructors release the objects they point to
    std::vector<std::unique_ptr<Object>> doomed;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        for (auto it = g_objects.begin(); it != g_objects.end();) {
            const Kind k = it->second->kind;
            if (k == Kind::Splitting || k == Kind::Boris || k == Kind::Splitting2D) {
                doomed.push_back(std::move(it->second));
                it = g_objects.erase(it);
            } else {
                ++it;
            }
        }
    }
    doomed.clear();
    {
        std::lock_guard<std::mutex> lk(g_mu);
        for (auto &kv : g_objects) doomed.push_back(std::move(kv.second));
        g_objects.clear();
        for (auto &z : g_zombies) doomed.push_back(std::move(z));
        g_zombies.clear();
    }
    doomed.clear();
}

// ---- staging ---------------------------------------------------------------------------
// Small host vectors (field dofs) go through a pinned bump buffer so that H2D copies are
// truly asynchronous and the caller's buffer can be reused as soon as the call returns.
static const size_t kPinnedBytes = 1 << 20;

static void ensure_pinned()
{
    Context &c = ctx();
    if (!c.pinned) {
        GP_CUDA(cudaMallocHost(&c.pinned, kPinnedBytes));
        c.pinned_bytes = kPinnedBytes;
        g_pinned_off = 0;
    }
}

void h2d(double *dev, const double *host, size_t n)
{
    Context &c = ctx();
    const size_t bytes = n * sizeof(double);
    if (bytes <= kPinnedBytes / 4) {
        ensure_pinned();
        if (g_pinned_off + bytes > kPinnedBytes) {
            GP_CUDA(cudaStreamSynchronize(c.stream));
            g_pinned_off = 0;
        }
        char *p = reinterpret_cast<char *>(c.pinned) + g_pinned_off;
        std::memcpy(p, host, bytes);
        GP_CUDA(cudaMemcpyAsync(dev, p, bytes, cudaMemcpyHostToDevice, c.stream));
        g_pinned_off += (bytes + 255) & ~(size_t)255;
    } else {
        GP_CUDA(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, c.stream));
        GP_CUDA(cudaStreamSynchronize(c.stream));
        g_pinned_off = 0;
    }
}

// k vectors of n doubles that sit back to back on the device <-> k separate host arrays (null = skip), in ONE transfer
// through the pinned staging buffer: the aliased e_dofs / b_dofs of a drop-in call (3 x 256 B) would otherwise be three
// pageable copies of ~10 us each way.
void h2d_vectors(double *dev, const double *const *host, int k, size_t n)
{
    Context &c = ctx();
    const size_t bytes = (size_t)k * n * sizeof(double);
    bool all = true;
    for (int i = 0; i < k; ++i) all = all && host[i];
    if (!all || bytes > kPinnedBytes / 4) {
        for (int i = 0; i < k; ++i)
            if (host[i]) h2d(dev + (size_t)i * n, host[i], n);
        return;
    }
    ensure_pinned();
    if (g_pinned_off + bytes > kPinnedBytes) {
        GP_CUDA(cudaStreamSynchronize(c.stream));
        g_pinned_off = 0;
    }
    char *p = reinterpret_cast<char *>(c.pinned) + g_pinned_off;
    for (int i = 0; i < k; ++i) std::memcpy(p + (size_t)i * n * sizeof(double), host[i], n * sizeof(double));
    GP_CUDA(cudaMemcpyAsync(dev, p, bytes, cudaMemcpyHostToDevice, c.stream));
    g_pinned_off += (bytes + 255) & ~(size_t)255;
}

void d2h_vectors(double *const *host, const double *dev, int k, size_t n)
{
    Context &c = ctx();
    int last = -1;
    for (int i = 0; i < k; ++i)
        if (host[i]) last = i;
    if (last < 0) {
        GP_CUDA(cudaStreamSynchronize(c.stream));
        return;
    }
    if (!host_out_enabled()) {   // replicated result: rank 0 writes the caller's arrays
        GP_CUDA(cudaStreamSynchronize(c.stream));
        return;
    }
    const size_t bytes = (size_t)(last + 1) * n * sizeof(double);
    if (bytes > kPinnedBytes / 4) {
        for (int i = 0; i <= last; ++i)
            if (host[i]) GP_CUDA(cudaMemcpyAsync(host[i], dev + (size_t)i * n, n * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
        GP_CUDA(cudaStreamSynchronize(c.stream));
        return;
    }
    ensure_pinned();
    // the staging buffer is shared with in-flight uploads: drain the stream by reading into a fresh region and waiting
    if (g_pinned_off + bytes > kPinnedBytes) {
        GP_CUDA(cudaStreamSynchronize(c.stream));
        g_pinned_off = 0;
    }
    char *p = reinterpret_cast<char *>(c.pinned) + g_pinned_off;
    GP_CUDA(cudaMemcpyAsync(p, dev, bytes, cudaMemcpyDeviceToHost, c.stream));
    GP_CUDA(cudaStreamSynchronize(c.stream));
    for (int i = 0; i <= last; ++i)
        if (host[i]) std::memcpy(host[i], p + (size_t)i * n * sizeof(double), n * sizeof(double));
    g_pinned_off = 0;   // everything before the synchronise has been consumed
}

void d2h(double *host, const double *dev, size_t n)
{
    Context &c = ctx();
    if (host_out_enabled()) GP_CUDA(cudaMemcpyAsync(host, dev, n * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    GP_CUDA(cudaStreamSynchronize(c.stream));
    g_pinned_off = 0;
}

// ---- per-kernel profiling ---------------------------------------------------------------
static cudaEvent_t take_event()
{
    if (!g_event_pool.empty()) {
        cudaEvent_t e = g_event_pool.back();
        g_event_pool.pop_back();
        return e;
    }
    cudaEvent_t e;
    GP_CUDA(cudaEventCreate(&e));
    return e;
}
static ProfSlot &prof_slot(const char *tag)
{
    for (auto &s : g_prof)
        if (s.tag == tag) return s;
    g_prof.emplace_back();
    g_prof.back().tag = tag;
    return g_prof.back();
}
void profile_begin(const char *tag)
{
    if (!g_profile) return;
    ProfSlot &s = prof_slot(tag);
    cudaEvent_t a = take_event(), b = take_event();
    GP_CUDA(cudaEventRecord(a, ctx().stream));
    s.pending.emplace_back(a, b);
}
void profile_end(const char *tag)
{
    if (!g_profile) return;
    ProfSlot &s = prof_slot(tag);
    GP_CUDA(cudaEventRecord(s.pending.back().second, ctx().stream));
}
static void profile_collect()
{
    for (auto &s : g_prof) {
        for (auto &pr : s.pending) {
            float ms = 0.f;
            GP_CUDA(cudaEventSynchronize(pr.second));
            GP_CUDA(cudaEventElapsedTime(&ms, pr.first, pr.second));
            s.ms += ms;
            s.launches++;
            g_event_pool.push_back(pr.first);
            g_event_pool.push_back(pr.second);
        }
        s.pending.clear();
    }
}

// ---- NCCL (dlopen) ---------------------------------------------------------------------
struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;

static std::mutex g_nccl_mu;
static NcclApi &nccl()
{
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    if (g_nccl.lib && g_nccl.GetErrorString) return g_nccl;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
        g_nccl.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.lib) break;
    }
    GP_REQUIRE(g_nccl.lib, GEMPIC_ENCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
#define GP_SYM(field, name)                                                           \
    g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(g_nccl.lib, name)); \
    GP_REQUIRE(g_nccl.field, GEMPIC_ENCCL, "libnccl lacks %s", name)
    GP_SYM(GetUniqueId, "ncclGetUniqueId");
    GP_SYM(CommInitRank, "ncclCommInitRank");
    GP_SYM(CommInitAll, "ncclCommInitAll");
    GP_SYM(CommDestroy, "ncclCommDestroy");
    GP_SYM(AllReduce, "ncclAllReduce");
    GP_SYM(AllGather, "ncclAllGather");
    GP_SYM(GetErrorString, "ncclGetErrorString");
#undef GP_SYM
    return g_nccl;
}

#define GP_NCCL(expr)                                                                       \
    do {                                                                                    \
        ncclResult_t _r = (expr);                                                           \
        if (_r != ncclSuccess)                                                              \
            ::gempic::fail(GEMPIC_ENCCL, "%s failed: %s", #expr, nccl().GetErrorString(_r)); \
    } while (0)

// ---- peer-memory exchange (xchg.cuh) -------------------------------------------------------------------------
// stand-alone form: block b owns a slice of the vector and flag b
__global__ void __launch_bounds__(256) k_xchg_allreduce(XchgDev X, double *__restrict__ v, int n, int per_block)
{
    const int lo = blockIdx.x * per_block;
    const int cnt = min(per_block, n - lo);
    if (cnt > 0) xchg_allreduce_block(X, v + lo, cnt, blockIdx.x, lo);
}

XchgDev xchg_next()
{
    Context &c = ctx();
    XchgDev X{};
    X.n_ranks = 1;
    if (!c.xchg_ready || c.n_ranks <= 1) return X;
    X.n_ranks = c.n_ranks;
    X.rank = c.rank;
    X.seq = ++c.xchg_seq;
    for (int r = 0; r < c.n_ranks; ++r) X.buf[r] = c.xchg_buf[r];
    return X;
}

void allreduce_sum(double *dev, int64_t n)
{
    Context &c = ctx();
    if (c.n_ranks <= 1 || n <= 0) return;
    if (c.xchg_ready && n <= kXchgSlot) {
        int blocks = (int)std::min<int64_t>(kXchgFlags, (n + 511) / 512);
        const int per_block = (int)((n + blocks - 1) / blocks);
        blocks = (int)((n + per_block - 1) / per_block);
        k_xchg_allreduce<<<blocks, 256, 0, c.stream>>>(xchg_next(), dev, (int)n, per_block);
        GP_CUDA(cudaGetLastError());
        count_launch();
        return;
    }
    GP_NCCL(nccl().AllReduce(dev, dev, (size_t)n, ncclDouble, ncclSum, (ncclComm_t)c.nccl_comm, c.stream));
    count_launch();
}

static bool xchg_wanted()
{
    const char *e = getenv("GEMPIC_NO_XCHG");
    return !(e && e[0] && e[0] != '0');
}

// every rank agrees (min over ranks) on whether the exchange buffers are usable; a single rank that failed to map a
// peer would otherwise wait for exchanges the others route through NCCL
static void xchg_agree(bool mine_ok)
{
    Context &c = ctx();
    DevBuf<double> flag(1);
    const double v = mine_ok ? 0.0 : 1.0;
    GP_CUDA(cudaMemcpyAsync(flag.p, &v, sizeof(double), cudaMemcpyHostToDevice, c.stream));
    GP_NCCL(nccl().AllReduce(flag.p, flag.p, 1, ncclDouble, ncclSum, (ncclComm_t)c.nccl_comm, c.stream));
    double sum = 1.0;
    GP_CUDA(cudaMemcpyAsync(&sum, flag.p, sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    GP_CUDA(cudaStreamSynchronize(c.stream));
    c.xchg_ready = sum == 0.0;
}

// one process per GPU: allocate the buffer, all-gather the cudaIpc handles through NCCL, map the peers
static void xchg_setup_ipc()
{
    Context &c = ctx();
    c.xchg_ready = false;
    if (c.n_ranks <= 1 || c.n_ranks > kXchgMaxRanks || !c.nccl_comm) return;
    bool ok = xchg_wanted();
    double *mine = nullptr;
    cudaIpcMemHandle_t hmine;
    std::memset(&hmine, 0, sizeof(hmine));
    if (ok) ok = cudaMalloc(&mine, kXchgBufDoubles * sizeof(double)) == cudaSuccess;
    if (ok) ok = cudaMemset(mine, 0, kXchgBufDoubles * sizeof(double)) == cudaSuccess;
    if (ok) ok = cudaIpcGetMemHandle(&hmine, mine) == cudaSuccess;
    cudaGetLastError();
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    DevBuf<double> send(8), recv((size_t)8 * c.n_ranks);
    GP_CUDA(cudaMemcpyAsync(send.p, &hmine, 64, cudaMemcpyHostToDevice, c.stream));
    GP_NCCL(nccl().AllGather(send.p, recv.p, 8, ncclDouble, (ncclComm_t)c.nccl_comm, c.stream));
    std::vector<cudaIpcMemHandle_t> all(c.n_ranks);
    GP_CUDA(cudaMemcpyAsync(all.data(), recv.p, (size_t)64 * c.n_ranks, cudaMemcpyDeviceToHost, c.stream));
    GP_CUDA(cudaStreamSynchronize(c.stream));
    xchg_agree(ok);              // everybody has a buffer and a handle?
    if (!c.xchg_ready) {
        if (mine) cudaFree(mine);
        return;
    }
    c.xchg_buf[c.rank] = mine;
    for (int r = 0; r < c.n_ranks && ok; ++r) {
        if (r == c.rank) continue;
        void *p = nullptr;
        if (cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) ok = false;
        c.xchg_buf[r] = (double *)p;
    }
    cudaGetLastError();
    xchg_agree(ok);              // everybody has mapped everybody?
    c.xchg_seq = 0;
}

static void xchg_teardown()
{
    Context &c = ctx();
    for (int r = 0; r < kXchgMaxRanks; ++r) {
        if (!c.xchg_buf[r]) continue;
        if (r == c.rank) cudaFree(c.xchg_buf[r]);
        else if (!md::g_in_process_peers) cudaIpcCloseMemHandle(c.xchg_buf[r]);
        c.xchg_buf[r] = nullptr;
    }
    c.xchg_ready = false;
    cudaGetLastError();
}

// ---- in-process multi-device mode ----------------------------------------------------------------------------
namespace md {

struct Worker {
    Rank rank;
    int device = 0, index = 0;
    std::thread th;
    std::mutex m;
    std::condition_variable cv;
    const std::function<int(int)> *job = nullptr;
    bool done = true, quit = false;
    int rc = 0;
    std::string err;
};
static void stop_workers();
struct WorkerList : std::vector<std::unique_ptr<Worker>> {
    ~WorkerList() { stop_workers(); }   // exit without gempic_finalize: the threads must be joined before they are destroyed
};
static WorkerList g_workers;
static thread_local bool t_in_worker = false;

static void worker_main(Worker *w)
{
    cudaSetDevice(w->device);
    t_rank = &w->rank;
    t_in_worker = true;
    std::unique_lock<std::mutex> lk(w->m);
    for (;;) {
        w->cv.wait(lk, [&] { return w->job || w->quit; });
        if (w->quit) return;
        const std::function<int(int)> *job = w->job;
        lk.unlock();
        int rc;
        try {
            rc = (*job)(w->index);
        } catch (...) {
            set_error("internal error in the worker of device %d", w->device);
            rc = GEMPIC_ECUDA;
        }
        lk.lock();
        w->rc = rc;
        w->err = rc ? g_error : std::string();
        w->job = nullptr;
        w->done = true;
        w->cv.notify_all();
    }
}

bool dispatching() { return !g_workers.empty() && !t_in_worker; }
int n_ranks() { return g_workers.empty() ? 1 : (int)g_workers.size(); }

static int run_on(const std::function<int(int)> &fn, int first, int last)
{
    for (int r = first; r < last; ++r) {
        Worker &w = *g_workers[r];
        std::lock_guard<std::mutex> lk(w.m);
        w.job = &fn;
        w.done = false;
        w.cv.notify_all();
    }
    int rc = 0;
    for (int r = first; r < last; ++r) {
        Worker &w = *g_workers[r];
        std::unique_lock<std::mutex> lk(w.m);
        w.cv.wait(lk, [&] { return w.done; });
        if (w.rc && !rc) {
            rc = w.rc;
            g_error = w.err;
        }
    }
    return rc;
}
int run_all(const std::function<int(int)> &fn) { return run_on(fn, 0, (int)g_workers.size()); }
int run_rank0(const std::function<int(int)> &fn) { return run_on(fn, 0, 1); }
int run_on_caller(const std::function<int(int)> &fn)
{
    // the workers are idle between dispatches, so the caller may borrow rank 0's state (host-only work: the set-up
    // quadratures that call back into the caller's language run here, never on a foreign thread)
    Rank *saved = t_rank;
    t_rank = &g_workers[0]->rank;
    int rc;
    try {
        rc = fn(0);
    } catch (...) {
        t_rank = saved;
        throw;
    }
    t_rank = saved;
    return rc;
}

void shard(int64_t n_global, int rank, int64_t &first, int64_t &count)
{
    const int64_t R = n_ranks();
    int64_t per = (n_global + R - 1) / R;
    per = (per + 31) / 32 * 32;   // multiples of 32: symmetric sampling needs multiples of 8, rows stay 256 B aligned
    first = std::min<int64_t>((int64_t)rank * per, n_global);
    count = std::min<int64_t>(per, n_global - first);
}

void set_host_out(bool on) { rk().host_out = on; }

static void stop_workers()
{
    for (auto &w : g_workers) {
        {
            std::lock_guard<std::mutex> lk(w->m);
            w->quit = true;
            w->cv.notify_all();
        }
        if (w->th.joinable()) w->th.join();
    }
    g_workers.clear();
}

}  // namespace md
}  // namespace gempic

using namespace gempic;

static int init_this_rank(int device);

extern "C" {

int gempic_device_count(void) { return md::n_ranks(); }

int gempic_init_devices(int n_devices, const int *device_ids)
{
    GP_API_BEGIN
    GP_REQUIRE(!g_rank0.ctx.ready && md::g_workers.empty(), GEMPIC_EINVAL, "the library is already initialised");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        fail(GEMPIC_ENOTINIT, "no CUDA device available (%s); libgempic_b200 has no CPU path",
             e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    GP_REQUIRE(n_devices >= 1 && n_devices <= count, GEMPIC_EINVAL, "%d devices requested, %d visible", n_devices, count);
    std::vector<int> ids(n_devices);
    for (int r = 0; r < n_devices; ++r) {
        ids[r] = device_ids ? device_ids[r] : r;
        GP_REQUIRE(ids[r] >= 0 && ids[r] < count, GEMPIC_EINVAL, "device %d out of range [0,%d)", ids[r], count);
        for (int q = 0; q < r; ++q) GP_REQUIRE(ids[q] != ids[r], GEMPIC_EINVAL, "device %d listed twice", ids[r]);
    }
    for (int r = 0; r < n_devices; ++r) {
        auto w = std::make_unique<md::Worker>();
        w->device = ids[r];
        w->index = r;
        w->rank.host_out = r == 0;
        w->th = std::thread(md::worker_main, w.get());
        md::g_workers.push_back(std::move(w));
    }
    int rc = md::run_all([&](int r) { return init_this_rank(ids[r]); });
    if (rc == GEMPIC_OK && n_devices > 1) {
        std::vector<ncclComm_t> comms(n_devices);
        ncclResult_t nr = nccl().CommInitAll(comms.data(), n_devices, ids.data());
        if (nr != ncclSuccess) {
            set_error("ncclCommInitAll failed: %s", nccl().GetErrorString(nr));
            rc = GEMPIC_ENCCL;
        } else {
            for (int r = 0; r < n_devices; ++r) {
                Context &c = md::g_workers[r]->rank.ctx;
                c.nccl_comm = comms[r];
                c.n_ranks = n_devices;
                c.rank = r;
            }
            // exchange buffers over plain peer access (same process: no IPC handles needed)
            if (n_devices <= kXchgMaxRanks && xchg_wanted()) {
                std::vector<double *> bufs(n_devices, nullptr);
                std::vector<int> okv(n_devices, 0);
                md::run_all([&](int r) {
                    bool ok = true;
                    for (int q = 0; q < n_devices && ok; ++q) {
                        if (q == r) continue;
                        int can = 0;
                        ok = cudaDeviceCanAccessPeer(&can, ids[r], ids[q]) == cudaSuccess && can;
                        if (ok) {
                            const cudaError_t pe = cudaDeviceEnablePeerAccess(ids[q], 0);
                            ok = pe == cudaSuccess || pe == cudaErrorPeerAccessAlreadyEnabled;
                        }
                    }
                    cudaGetLastError();
                    if (ok) ok = cudaMalloc(&bufs[r], kXchgBufDoubles * sizeof(double)) == cudaSuccess &&
                                 cudaMemset(bufs[r], 0, kXchgBufDoubles * sizeof(double)) == cudaSuccess;
                    cudaDeviceSynchronize();
                    okv[r] = ok ? 1 : 0;
                    return 0;
                });
                bool all_ok = true;
                for (int r = 0; r < n_devices; ++r) all_ok = all_ok && okv[r];
                md::g_in_process_peers = true;
                for (int r = 0; r < n_devices; ++r) {
                    Context &c = md::g_workers[r]->rank.ctx;
                    if (all_ok) {
                        for (int q = 0; q < n_devices; ++q) c.xchg_buf[q] = bufs[q];
                        c.xchg_ready = true;
                        c.xchg_seq = 0;
                    } else {
                        c.xchg_buf[r] = bufs[r];   // freed by xchg_teardown
                    }
                }
            }
        }
    }
    if (rc != GEMPIC_OK) {
        md::run_all([&](int) { return gempic_finalize(); });   // (renamed: the per-rank implementation)
        md::stop_workers();
        return rc;
    }
    GP_API_END
}

const char *gempic_last_error(void) { return g_error.c_str(); }
int gempic_version(void) { return 100; }

}  // extern "C"

static int init_this_rank(int device) { return gempic_init(device); }

// called by md.cu after the workers have finalised their ranks
namespace gempic { namespace md { void shutdown() { stop_workers(); } } }

extern "C" {

int gempic_init(int device)
{
    GP_API_BEGIN
    Context &c = ctx();
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        fail(GEMPIC_ENOTINIT, "no CUDA device available (%s); libgempic_b200 has no CPU path",
             e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    GP_REQUIRE(device >= 0 && device < count, GEMPIC_EINVAL, "device %d out of range [0,%d)", device, count);
    if (c.ready && c.device == device) return GEMPIC_OK;
    GP_REQUIRE(!c.ready, GEMPIC_EINVAL, "already initialised on device %d", c.device);
    GP_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    GP_CUDA(cudaGetDeviceProperties(&prop, device));
    GP_REQUIRE(prop.major == 10, GEMPIC_ENOTINIT,
               "device %d is sm_%d%d; this library ships sm_100a code only (B200)", device, prop.major, prop.minor);
    c.device = device;
    c.sm_count = prop.multiProcessorCount;
    c.smem_optin = prop.sharedMemPerBlockOptin;
    GP_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    c.ready = true;
    GP_API_END
}

int gempic_finalize(void)
{
    GP_API_BEGIN
    Context &c = ctx();
    if (!c.ready) return GEMPIC_OK;
    cudaStreamSynchronize(c.stream);
    destroy_all();
    if (c.nccl_comm) {
        xchg_teardown();
        nccl().CommDestroy((ncclComm_t)c.nccl_comm);
        c.nccl_comm = nullptr;
        c.n_ranks = 1;
        c.rank = 0;
    }
    if (c.pinned) cudaFreeHost(c.pinned);
    c.pinned = nullptr;
    c.pinned_bytes = 0;
    if (c.stage) cudaFree(c.stage);
    c.stage = nullptr;
    c.stage_n = 0;
    c.stage_busy = false;
    cudaStreamDestroy(c.stream);
    c.stream = nullptr;
    c.ready = false;
    GP_API_END
}

int gempic_synchronize(void)
{
    GP_API_BEGIN
    require_init();
    GP_CUDA(cudaStreamSynchronize(ctx().stream));
    GP_API_END
}

void *gempic_stream(void) { return (void *)ctx().stream; }

int gempic_device_info(int *sm_count, int64_t *free_bytes, int64_t *total_bytes)
{
    GP_API_BEGIN
    require_init();
    size_t f = 0, t = 0;
    GP_CUDA(cudaMemGetInfo(&f, &t));
    if (sm_count) *sm_count = ctx().sm_count;
    if (free_bytes) *free_bytes = (int64_t)f;
    if (total_bytes) *total_bytes = (int64_t)t;
    GP_API_END
}

int64_t gempic_launch_count(int reset)
{
    int64_t v = ctx().launches;
    if (reset) ctx().launches = 0;
    return v;
}

int gempic_set_option(const char *name, int64_t value)
{
    GP_API_BEGIN
    GP_REQUIRE(name, GEMPIC_EINVAL, "null option name");
    (void)value;
    fail(GEMPIC_EINVAL, "unknown option '%s'", name);
    GP_API_END
}

int gempic_profile_enable(int on)
{
    GP_API_BEGIN
    require_init();
    profile_collect();
    g_profile = on != 0;
    if (on) g_prof.clear();
    GP_API_END
}

int gempic_profile_read(int slot, char *tag, int tag_len, double *ms, int64_t *launches)
{
    GP_API_BEGIN
    require_init();
    profile_collect();
    GP_REQUIRE(slot >= 0, GEMPIC_EINVAL, "negative slot");
    if (slot >= (int)g_prof.size()) return 100;  // end of list
    const ProfSlot &s = g_prof[slot];
    if (tag && tag_len > 0) {
        std::strncpy(tag, s.tag.c_str(), tag_len - 1);
        tag[tag_len - 1] = 0;
    }
    if (ms) *ms = s.ms;
    if (launches) *launches = s.launches;
    GP_API_END
}

int gempic_comm_unique_id(void *id128)
{
    GP_API_BEGIN
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    GP_REQUIRE(id128, GEMPIC_EINVAL, "null id buffer");
    ncclUniqueId id;
    GP_NCCL(nccl().GetUniqueId(&id));
    std::memcpy(id128, &id, sizeof(id));
    GP_API_END
}

int gempic_comm_init(int n_ranks, int rank, const void *id128)
{
    GP_API_BEGIN
    require_init();
    Context &c = ctx();
    GP_REQUIRE(n_ranks >= 1 && rank >= 0 && rank < n_ranks, GEMPIC_EINVAL, "bad rank %d of %d", rank, n_ranks);
    GP_REQUIRE(!c.nccl_comm, GEMPIC_EINVAL, "communicator already initialised");
    if (n_ranks == 1) {
        c.n_ranks = 1;
        c.rank = 0;
        return GEMPIC_OK;
    }
    GP_REQUIRE(id128, GEMPIC_EINVAL, "null id buffer");
    ncclUniqueId id;
    std::memcpy(&id, id128, sizeof(id));
    ncclComm_t comm;
    GP_NCCL(nccl().CommInitRank(&comm, n_ranks, id, rank));
    c.nccl_comm = comm;
    c.n_ranks = n_ranks;
    c.rank = rank;
    xchg_setup_ipc();
    GP_API_END
}

int gempic_comm_finalize(void)
{
    GP_API_BEGIN
    Context &c = ctx();
    if (c.nccl_comm) {
        cudaStreamSynchronize(c.stream);
        xchg_teardown();
        GP_NCCL(nccl().CommDestroy((ncclComm_t)c.nccl_comm));
        c.nccl_comm = nullptr;
    }
    c.n_ranks = 1;
    c.suspended_ranks = 0;
    c.rank = 0;
    GP_API_END
}

int gempic_comm_size(void) { return ctx().n_ranks; }

/* suspend = 1: the communicator stays alive but this rank works on its own (every all-reduce is skipped, the single-GPU
 * launch sequences are used) until suspend = 0.  For un-sharded reference runs next to a sharded one (bench.py's
 * sharded_parity, tests/dist_worker.py); every rank must resume before the next collective. */
int gempic_comm_suspend(int suspend)
{
    GP_API_BEGIN
    require_init();
    Context &c = ctx();
    if (suspend && !c.suspended_ranks && c.nccl_comm) {
        GP_CUDA(cudaStreamSynchronize(c.stream));
        c.suspended_ranks = c.n_ranks;
        c.n_ranks = 1;
    } else if (!suspend && c.suspended_ranks) {
        GP_CUDA(cudaStreamSynchronize(c.stream));
        c.n_ranks = c.suspended_ranks;
        c.suspended_ranks = 0;
    }
    GP_API_END
}

int gempic_comm_allreduce(double *host_inout, int64_t n)
{
    GP_API_BEGIN
    require_init();
    GP_REQUIRE(host_inout && n > 0, GEMPIC_EINVAL, "bad buffer");
    DevBuf<double> d((size_t)n);
    h2d(d.p, host_inout, (size_t)n);
    allreduce_sum(d.p, n);
    d2h(host_inout, d.p, (size_t)n);
    GP_API_END
}

}  // extern "C"
