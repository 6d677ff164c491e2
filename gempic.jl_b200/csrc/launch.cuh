// launch.cuh -- host-side launch planning for k_pass instantiations.
#pragma once
#include <cstdlib>

#include "ops1d.cuh"

namespace gempic {

// Scratch for per-block partial sums, grown on demand, owned by the context user.
struct PartialScratch {
    DevBuf<double> buf;
    double *ensure(size_t n)
    {
        if (buf.n < n) buf.alloc(n);
        return buf.p;
    }
};

// LANE mode is used while one block's private copies stay below this many bytes, which
// keeps >= 2 blocks (8 warps) resident per SM.
constexpr size_t kLanePrivateMaxBytes = 110 * 1024;
constexpr size_t kSmemBudget = 200 * 1024;
constexpr size_t kSmemMaxOptin = 232448;   // 227 KB: largest dynamic shared memory of one sm_100 block

struct LaunchInfo {
    int grid = 0;
    size_t smem = 0;
    bool lane_private = false;
    int copies = 0;
};

template <class Op>
inline LaunchInfo plan_pass(const PassParams<Op> &P)
{
    LaunchInfo L;
    constexpr int kNW = op_threads<Op>::value / 32;
    const size_t field_bytes = (size_t)Op::NF * (P.m.n + op_field_halo<Op>::value) * op_field_copies<Op>::value * sizeof(double);
    if (!Op::DEPOSIT) {
        L.lane_private = true;
        L.smem = field_bytes;
        return L;
    }
    const size_t one = (size_t)acc_slots<Op>(P.m.n) * sizeof(double);
    const size_t lp = one * 32 * kNW;
    // 128-thread ops keep two blocks per SM; a 256-thread op is planned as one block per SM
    const size_t lp_max = op_threads<Op>::value > 128 ? kSmemMaxOptin : kLanePrivateMaxBytes;
    if (field_bytes + lp <= lp_max) {
        L.lane_private = true;
        L.smem = field_bytes + lp;
    } else {
        GP_REQUIRE(field_bytes + one <= kSmemBudget, GEMPIC_EINVAL,
                   "deposit grid of %d dofs does not fit in shared memory", acc_outputs<Op>(P.m.n));
        size_t copies = (kSmemBudget / 2 - field_bytes) / one;
        if (copies < 1) copies = 1;
        if (copies > (size_t)kNW) copies = kNW;
        L.lane_private = false;
        L.copies = (int)copies;
        L.smem = field_bytes + copies * one;
    }
    return L;
}

// GEMPIC_K1_TMA=1: the 256-thread deposit passes (fused Strang pass, Boris step) stream their rows with cp.async.bulk
// instead of the register prefetch (round-2 experiment; slower, see DESIGN.md section 6, so off by default)
inline bool pass_tma_wanted()
{
    static const bool on = [] { const char *e = getenv("GEMPIC_K1_TMA"); return e && e[0] == '1'; }();
    return on;
}

template <class Op, bool LP, bool TMA = false>
inline int configure_kernel(size_t smem)
{
    auto kern = k_pass<Op, LP, TMA>;
    if (smem > 48 * 1024) ensure_func_smem((const void *)kern, smem);
    int per_sm = 0;
    GP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, op_threads<Op>::value, smem));
    return per_sm;
}

// Launch one pass. For deposit ops the reduced accumulators (n_acc doubles) land in `out`
// (device) through `scratch`.  `out` may be null for non-deposit ops.  With `defer` the per-block partials are left
// un-reduced for the caller's next kernel (k_strang_fields sums them itself): defer->partials / n_blocks describe them
// (n_blocks = 0: no particles, the sums are zero).
struct DeferredReduce {
    const double *partials = nullptr;
    int n_blocks = 0;
};
template <class Op>
inline void launch_pass(PassParams<Op> P, PartialScratch *scratch, double *out, const char *tag = nullptr,
                        DeferredReduce *defer = nullptr)
{
    Context &c = ctx();
    const int n_out = acc_outputs<Op>(P.m.n);
    if (P.n_particles <= 0) {
        if (Op::DEPOSIT && out && !defer) GP_CUDA(cudaMemsetAsync(out, 0, sizeof(double) * n_out, c.stream));
        return;
    }
    LaunchInfo L = plan_pass(P);
    bool tma = false;
    if constexpr (Op::DEPOSIT && op_threads<Op>::value == 256)
        tma = pass_tma_wanted() && L.lane_private && L.smem + pass_tma_bytes<Op>() <= kSmemMaxOptin;
    if (tma) L.smem += pass_tma_bytes<Op>();
    int per_sm;
    if constexpr (Op::DEPOSIT && op_threads<Op>::value == 256) {
        per_sm = tma ? configure_kernel<Op, true, true>(L.smem)
                     : (L.lane_private ? configure_kernel<Op, true>(L.smem) : configure_kernel<Op, false>(L.smem));
    } else {
        per_sm = L.lane_private ? configure_kernel<Op, true>(L.smem) : configure_kernel<Op, false>(L.smem);
    }
    GP_REQUIRE(per_sm >= 1, GEMPIC_EINVAL, "pass does not fit on an SM (smem %zu B)", L.smem);
    int grid = c.sm_count * per_sm;
    const int64_t pairs = (P.n_particles + 1) / 2;
    constexpr int kThreads = op_threads<Op>::value;
    const int64_t need = (pairs + kThreads - 1) / kThreads;
    if (need < grid) grid = (int)need;
    P.copies = L.copies;
    if (Op::DEPOSIT) P.partials = scratch->ensure((size_t)grid * n_out);
    if (tag) profile_begin(tag);
    if constexpr (Op::DEPOSIT && op_threads<Op>::value == 256) {
        if (tma) k_pass<Op, true, true><<<grid, kThreads, L.smem, c.stream>>>(P);
        else if (L.lane_private) k_pass<Op, true><<<grid, kThreads, L.smem, c.stream>>>(P);
        else k_pass<Op, false><<<grid, kThreads, L.smem, c.stream>>>(P);
    } else {
        if (L.lane_private) k_pass<Op, true><<<grid, kThreads, L.smem, c.stream>>>(P);
        else k_pass<Op, false><<<grid, kThreads, L.smem, c.stream>>>(P);
    }
    GP_CUDA(cudaGetLastError());
    if (tag) profile_end(tag);
    count_launch();
    if (Op::WRITE != 0) particles_changed();
    if (Op::DEPOSIT && defer) {
        defer->partials = P.partials;
        defer->n_blocks = grid;
    } else if (Op::DEPOSIT) {
        const int warps_per_block = 4;
        const int blocks = (n_out + warps_per_block - 1) / warps_per_block;
        k_reduce_partials<<<blocks, warps_per_block * 32, 0, c.stream>>>(P.partials, grid, n_out, out);
        GP_CUDA(cudaGetLastError());
        count_launch();
    }
}

// Degree dispatch: D0 in 1..3, D1 in {D0-1, D0}  (the reference's Maxwell1DFEM supports
// degree 1..3 with s_deg_1 = degree-1; test_vm_1d2v.jl uses equal degrees)
#define GP_DISPATCH_DEGREES(D0v, D1v, ...)                                                    \
    do {                                                                                       \
        const int _k = (D0v) * 10 + (D1v);                                                     \
        switch (_k) {                                                                          \
        case 10: { constexpr int D0 = 1, D1 = 0; __VA_ARGS__; } break;                                \
        case 11: { constexpr int D0 = 1, D1 = 1; __VA_ARGS__; } break;                                \
        case 21: { constexpr int D0 = 2, D1 = 1; __VA_ARGS__; } break;                                \
        case 22: { constexpr int D0 = 2, D1 = 2; __VA_ARGS__; } break;                                \
        case 32: { constexpr int D0 = 3, D1 = 2; __VA_ARGS__; } break;                                \
        case 33: { constexpr int D0 = 3, D1 = 3; __VA_ARGS__; } break;                                \
        default:                                                                               \
            ::gempic::fail(GEMPIC_EINVAL, "unsupported spline degree pair (%d, %d)", (D0v), (D1v)); \
        }                                                                                      \
    } while (0)

#define GP_DISPATCH_DEGREE(Dv, ...)                                                    \
    do {                                                                               \
        switch (Dv) {                                                                  \
        case 0: { constexpr int D = 0; __VA_ARGS__; } break;                                  \
        case 1: { constexpr int D = 1; __VA_ARGS__; } break;                                  \
        case 2: { constexpr int D = 2; __VA_ARGS__; } break;                                  \
        case 3: { constexpr int D = 3; __VA_ARGS__; } break;                                  \
        default: ::gempic::fail(GEMPIC_EINVAL, "unsupported spline degree %d", (Dv));  \
        }                                                                              \
    } while (0)

}  // namespace gempic
