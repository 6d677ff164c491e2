// common.cuh -- runtime context, handle registry and error plumbing of libgempic_b200.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <memory>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include <functional>

#include "../../include/gempic_b200.h"
#include "xchg.cuh"
#ifndef GEMPIC_NO_RENAME
#include "md_rename.inc"   // gempic_X -> gempic_impl_X: the .cu files define the per-rank implementations (md.cu wraps them)
#endif

namespace gempic {

// ---- errors ---------------------------------------------------------------------------
void set_error(const char *fmt, ...);
struct Fail {
    int code;
};  // thrown inside the library, converted to a status at the C boundary

[[noreturn]] void fail(int code, const char *fmt, ...);

#define GP_CUDA(expr)                                                                        \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess)                                                               \
            ::gempic::fail(GEMPIC_ECUDA, "%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, \
                           cudaGetErrorString(_e));                                          \
    } while (0)

#define GP_REQUIRE(cond, code, ...)                  \
    do {                                             \
        if (!(cond)) ::gempic::fail(code, __VA_ARGS__); \
    } while (0)

// ---- context --------------------------------------------------------------------------
struct Context {
    bool ready = false;
    int device = -1;
    int sm_count = 0;
    size_t smem_optin = 0;
    cudaStream_t stream = nullptr;
    int64_t launches = 0;
    uint64_t particle_epoch = 1;   // bumped by every launch that writes particle rows (invalidates ParticleGroup::tail)
    // multi-GPU
    void *nccl_comm = nullptr;
    int n_ranks = 1, rank = 0;
    int suspended_ranks = 0;   // gempic_comm_suspend: n_ranks while this rank works on its own
    // NVLink peer-memory exchange (xchg.cuh): buffers of all ranks as mapped here; ready once every rank has them
    bool xchg_ready = false;
    double *xchg_buf[kXchgMaxRanks] = {nullptr};
    unsigned long long xchg_seq = 0;
    double *pinned = nullptr;  // small pinned staging buffer for field I/O
    size_t pinned_bytes = 0;
    double *stage = nullptr;   // grow-only device scratch of the host-buffer entry points (hostutil.hpp Stage)
    size_t stage_n = 0;
    bool stage_busy = false;
};
Context &ctx();   // of the calling thread's rank (one per process, or one per device after gempic_init_devices)
void require_init();
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (rank, kernel, size): function attributes are per device
void ensure_func_smem(const void *func, size_t bytes);
// false on the ranks > 0 of an in-process multi-device run: replicated host outputs are written by rank 0 only
bool host_out_enabled();
inline void count_launch(int n = 1) { ctx().launches += n; }
inline void particles_changed() { ctx().particle_epoch++; }
// optional per-kernel device timing (CUDA events around each tagged launch, no host sync);
// read back with gempic_profile_read after a synchronize
void profile_begin(const char *tag);
void profile_end(const char *tag);
// sum-allreduce `n` doubles in place on the library stream (no-op for one rank): one k_xchg_allreduce launch over peer
// memory when the exchange buffers are mapped and the vector fits, ncclAllReduce otherwise
void allreduce_sum(double *dev, int64_t n);
// descriptor of the NEXT in-kernel exchange (advances the sequence number); n_ranks = 1 when there is nothing to exchange
XchgDev xchg_next();
inline bool xchg_active() { return ctx().xchg_ready && ctx().n_ranks > 1; }

// ---- in-process multi-device mode (gempic_init_devices; runtime.cu, md.cu) --------------------------------------
namespace md {
bool dispatching();   // multi-device mode is on and the caller is not one of the worker threads
int n_ranks();
// fn(rank) on every worker thread / on rank 0's thread / on the calling thread with rank 0's state (C callbacks);
// returns the first non-zero status and leaves that rank's message in the caller's gempic_last_error()
int run_all(const std::function<int(int)> &fn);
int run_rank0(const std::function<int(int)> &fn);
int run_on_caller(const std::function<int(int)> &fn);
void shard(int64_t n_global, int rank, int64_t &first, int64_t &count);   // index range of a rank (multiples of 32)
void set_host_out(bool on);   // calling worker: write host outputs (per-particle arrays) although it is not rank 0
}  // namespace md

// ---- device buffers -------------------------------------------------------------------
template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    explicit DevBuf(size_t count) { alloc(count); }
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    void alloc(size_t count)
    {
        release();
        n = count;
        if (count) GP_CUDA(cudaMalloc(&p, count * sizeof(T)));
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    void zero(cudaStream_t s) { if (n) GP_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s)); }
};

void h2d(double *dev, const double *host, size_t n);  // synchronous w.r.t. the host buffer
void d2h(double *host, const double *dev, size_t n);  // synchronises the library stream
// k back-to-back device vectors of n doubles <-> k host arrays (null entries skipped), one transfer through pinned memory
void h2d_vectors(double *dev, const double *const *host, int k, size_t n);
void d2h_vectors(double *const *host, const double *dev, int k, size_t n);   // synchronises the library stream

// ---- objects behind handles -----------------------------------------------------------
enum class Kind : uint32_t { ParticleGroup = 1, Pmc1D, Pmc2D, Maxwell1D, Splitting, Boris, Maxwell2D, Splitting2D };

struct Object {
    Kind kind;
    int users = 0;   // splitting objects that hold a raw pointer to this one (retain / release)
    explicit Object(Kind k) : kind(k) {}
    virtual ~Object() = default;
};
// Lifetime of the objects a splitting points to (its particle group, smoothers, Maxwell solver): destroying a handle
// that is still retained only removes it from the registry; the memory goes when the last user releases it, so a
// garbage collector may finalise the host-side wrappers in any order.
void retain(Object *o);
void release(Object *o);

gempic_handle register_object(std::unique_ptr<Object> obj);
Object *lookup(gempic_handle h, Kind kind, const char *what);
void destroy(gempic_handle h, Kind kind, const char *what);
void destroy_all();

template <typename T>
T *get(gempic_handle h, const char *what)
{
    return static_cast<T *>(lookup(h, T::kKind, what));
}

}  // namespace gempic

// Converts exceptions into status codes at every extern "C" entry point.
#define GP_API_BEGIN try {
#define GP_API_END                                              \
    return GEMPIC_OK;                                           \
    }                                                           \
    catch (const ::gempic::Fail &f) { return f.code; }          \
    catch (const std::exception &e)                             \
    {                                                           \
        ::gempic::set_error("internal error: %s", e.what());    \
        return GEMPIC_ECUDA;                                    \
    }
