// xchg.cuh -- one-shot all-reduce of small vectors over NVLink peer memory, callable from inside a kernel.
//
// The grid moments of a depositing sub-step are 64 doubles (1D) to 8192 doubles (2D): far too small for a ring, and the
// step they sit in is `particle pass -> sum of the per-block partials -> sum over ranks -> field solve`.  With NCCL that
// is three stream-ordered launches (k_reduce_partials, ncclAllReduce, the field kernel) whose launch gaps and straggler
// waits are the whole scaling loss (round 1: 0.9725 at 8 GPUs).  Here every rank owns an exchange buffer that all its
// peers have mapped (cudaIpc handles between processes, plain peer access inside one process):
//     publish   copy the local vector into the own buffer, __threadfence_system, store the sequence number into the flag
//     gather    poll every rank's flag (acquire, system scope), then add the R vectors in RANK ORDER
// so the result is bitwise identical on all ranks (replicas cannot drift) and run-to-run deterministic.  The buffers
// are double-buffered by the parity of the sequence number: a rank can only publish exchange s+2 after every rank has
// published s+1, i.e. after every rank has finished reading s.
// xchg_allreduce_block is called by ONE block from inside the fused field kernels (k_strang_fields, k_boris_fields):
// reduction of the partials, exchange and field solve are one launch.  k_xchg_allreduce is the stand-alone form for the
// 2D grids (one flag per block).
#pragma once
#include <cuda_runtime.h>

namespace gempic {

constexpr int kXchgMaxRanks = 16;
constexpr int kXchgFlags = 64;                       // flags (= blocks of the stand-alone kernel) per parity
constexpr int kXchgFlagStride = 16;                  // doubles between two flags (128 B)
constexpr int kXchgHeader = kXchgFlags * kXchgFlagStride;
constexpr int kXchgSlot = 16384;                     // doubles of payload per parity (128 KB)
constexpr size_t kXchgBufDoubles = 2 * (size_t)(kXchgHeader + kXchgSlot);

struct XchgDev {
    double *buf[kXchgMaxRanks];   // every rank's exchange buffer as mapped into this rank
    int n_ranks, rank;            // n_ranks <= 1: no exchange
    unsigned long long seq;       // sequence number of this exchange (the same on every rank)
};

__device__ __forceinline__ unsigned long long xchg_ld_flag(const double *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void xchg_st_flag(double *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ double xchg_ld(const double *p)
{
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long xchg_now()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// v[0 .. n) (shared or global memory of the calling block) <- sum over ranks, in rank order.  Called by all threads of a
// block; `flag` selects the flag slot (block index of the stand-alone kernel, 0 inside the fused field kernels);
// `offset` is the position of v inside the exchanged vector.  A peer that does not show up within 5 s poisons the
// result with NaN instead of hanging the GPU.
__device__ __forceinline__ void xchg_allreduce_block(const XchgDev &X, double *v, int n, int flag, int offset)
{
    const size_t base = (size_t)(X.seq & 1ull) * (kXchgHeader + kXchgSlot);
    double *mine = X.buf[X.rank] + base;
    for (int i = threadIdx.x; i < n; i += blockDim.x) mine[kXchgHeader + offset + i] = v[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) xchg_st_flag(mine + flag * kXchgFlagStride, X.seq);
    __shared__ int xchg_ok;
    if (threadIdx.x == 0) xchg_ok = 1;
    __syncthreads();
    if (threadIdx.x < X.n_ranks) {
        const double *f = X.buf[threadIdx.x] + base + flag * kXchgFlagStride;
        const unsigned long long t0 = xchg_now();
        while (xchg_ld_flag(f) < X.seq) {
            if (xchg_now() - t0 > 5000000000ull) {
                xchg_ok = 0;
                break;
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double s = 0.0;
        for (int r = 0; r < X.n_ranks; ++r) s += xchg_ld(X.buf[r] + base + kXchgHeader + offset + i);
        v[i] = xchg_ok ? s : __longlong_as_double(0x7ff8000000000000ll);
    }
    __syncthreads();
}

}  // namespace gempic
