// pass.cuh -- the one streaming skeleton every particle kernel is built from.
//
// A "pass" streams the SoA particle rows once (128-bit coalesced loads/stores, two
// particles per thread per load, two independent load batches in flight), applies an Op
// per particle, and -- if the Op deposits -- accumulates grid moments in block-private
// shared memory, reduces them hierarchically and writes one partial vector per block.
// A tiny second kernel (reduce_partials) sums the per-block partials in a fixed order, so
// the deposit is run-to-run deterministic in LANE mode.
//
// Shared-memory layout
//   fields        Op::NF vectors of n + kHalo doubles; the first kHalo (= 4) dofs are repeated
//                 after the last one so that a gather reads dofs g0 .. g0+D with ONE
//                 periodic wrap per particle instead of one per dof.
//   accumulators  Op::NG grids of n + kHalo slots (same halo trick; the halo is folded back
//                 when the block reduces) followed by Op::NS scalar slots.
//
// Deposit accumulators (SURVEY section 7 "hard parts": fp64 shared atomics are CAS loops):
//   LANE mode  every lane of every warp owns a private copy of the (tiny) periodic grid:
//              slot(s) = warp_base[s*32 + lane]. Plain LDS/DADD/STS, no atomics, and bank
//              conflict free for any cell pattern (a half-warp's 16 lanes x 8 B always
//              cover the 32 banks once).  Used when the copies fit in shared memory.
//   ATOM mode  `copies` block-shared copies, warp w adds into copy w % copies with
//              atomicAdd(double) on shared memory. Fallback for larger grids.
#pragma once
#include <type_traits>

#include "common.cuh"
#include "splines.cuh"
#include "tma.cuh"

namespace gempic {

constexpr int kBlock = 128;  // threads per block (4 warps)
constexpr int kWarps = kBlock / 32;
constexpr int kHalo = kMaxDegree + 1;  // a degree-D stencil window shifted by one cell still fits (fused Hp1)
constexpr int kMaxFields = 5;

enum RowBits : int { ROW_X = 1, ROW_V1 = 2, ROW_V2 = 4, ROW_W = 8 };

struct Rows {
    double *__restrict__ x;
    double *__restrict__ v1;
    double *__restrict__ v2;
    double *__restrict__ w;
};

struct Particle {
    double x, v1, v2, w;
};

// ---- optional per-Op launch traits (defaults in parentheses) -------------------------------
//   Op::THREADS       threads per block (kBlock)
//   Op::HALO          periodic halo dofs of the shared field / accumulator vectors (kHalo)
//   Op::FIELD_COPIES  lane-interleaved copies of every staged field dof (1).  With 16 copies a
//                     half-warp's gather is bank-conflict free for ANY cell pattern: element
//                     (dof i, copy c) sits at (i*16 + c) doubles and lane l reads copy l & 15.
//   Op::stage(P, sfield, tid)   custom staging of the Op::NF shared field vectors
template <class Op, class = void>
struct op_threads { static constexpr int value = 128; };
template <class Op>
struct op_threads<Op, std::void_t<decltype(Op::THREADS)>> { static constexpr int value = Op::THREADS; };
template <class Op, class = void>
struct op_halo { static constexpr int value = kMaxDegree + 1; };
template <class Op>
struct op_halo<Op, std::void_t<decltype(Op::HALO)>> { static constexpr int value = Op::HALO; };
template <class Op, class = void>
struct op_field_halo { static constexpr int value = op_halo<Op>::value; };   // halo of the staged FIELD vectors
template <class Op>
struct op_field_halo<Op, std::void_t<decltype(Op::FIELD_HALO)>> { static constexpr int value = Op::FIELD_HALO; };
template <class Op, class = void>
struct op_field_copies { static constexpr int value = 1; };
template <class Op>
struct op_field_copies<Op, std::void_t<decltype(Op::FIELD_COPIES)>> { static constexpr int value = Op::FIELD_COPIES; };
template <class Op, class = void>
struct op_has_stage : std::false_type {};
template <class Op>
struct op_has_stage<Op, std::void_t<decltype(Op::CUSTOM_STAGE)>> : std::true_type {};

constexpr int kMaxScalars = 5;
template <bool LP>
struct Acc {
    double *p;  // LP: warp base + lane ; ATOM: base of this warp's copy
    // scalar sums of an Op (Op::NS <= kMaxScalars) live in registers of the thread for the whole pass and enter the
    // accumulator slots once, before the block reduction (ncu r2e: as per-particle read-modify-writes of shared memory
    // they were a quarter of the wavefronts of the diagnostics passes)
    mutable double s[kMaxScalars] = {0.0, 0.0, 0.0, 0.0, 0.0};
    __device__ __forceinline__ void add(int slot, double v) const
    {
        if (LP) p[slot * 32] += v;
        else atomicAdd(p + slot, v);
    }
    __device__ __forceinline__ void sum(int k, double v) const { s[k] += v; }
};

template <class Op>
struct PassParams {
    Rows r;
    int64_t n_particles;
    Mesh1D m;
    const double *fields[kMaxFields];  // Op::NF device field vectors of m.n doubles staged in smem
    double *partials;                  // [gridDim.x][Op::NG * m.n + Op::NS] (deposit ops only)
    int copies;                        // ATOM mode: accumulator copies per block
    typename Op::Params op;
};

// slots of one accumulator copy / doubles of one reduced output vector
template <class Op>
__host__ __device__ constexpr int acc_slots(int n) { return Op::NG * (n + op_halo<Op>::value) + Op::NS; }
template <class Op>
__host__ __device__ constexpr int acc_outputs(int n) { return Op::NG * n + Op::NS; }

template <class Op>
__device__ __forceinline__ void load_pair(const Rows &r, int64_t pair, Particle &a, Particle &b)
{
    if (Op::READ & ROW_X) { double2 t = reinterpret_cast<const double2 *>(r.x)[pair]; a.x = t.x; b.x = t.y; }
    if (Op::READ & ROW_V1) { double2 t = reinterpret_cast<const double2 *>(r.v1)[pair]; a.v1 = t.x; b.v1 = t.y; }
    if (Op::READ & ROW_V2) { double2 t = reinterpret_cast<const double2 *>(r.v2)[pair]; a.v2 = t.x; b.v2 = t.y; }
    if (Op::READ & ROW_W) { double2 t = reinterpret_cast<const double2 *>(r.w)[pair]; a.w = t.x; b.w = t.y; }
}
template <class Op>
__device__ __forceinline__ void store_pair(const Rows &r, int64_t pair, const Particle &a, const Particle &b)
{
    if (Op::WRITE & ROW_X) reinterpret_cast<double2 *>(r.x)[pair] = make_double2(a.x, b.x);
    if (Op::WRITE & ROW_V1) reinterpret_cast<double2 *>(r.v1)[pair] = make_double2(a.v1, b.v1);
    if (Op::WRITE & ROW_V2) reinterpret_cast<double2 *>(r.v2)[pair] = make_double2(a.v2, b.v2);
}
template <class Op>
__device__ __forceinline__ void load_one(const Rows &r, int64_t i, Particle &a)
{
    if (Op::READ & ROW_X) a.x = r.x[i];
    if (Op::READ & ROW_V1) a.v1 = r.v1[i];
    if (Op::READ & ROW_V2) a.v2 = r.v2[i];
    if (Op::READ & ROW_W) a.w = r.w[i];
}
template <class Op>
__device__ __forceinline__ void store_one(const Rows &r, int64_t i, const Particle &a)
{
    if (Op::WRITE & ROW_X) r.x[i] = a.x;
    if (Op::WRITE & ROW_V1) r.v1[i] = a.v1;
    if (Op::WRITE & ROW_V2) r.v2[i] = a.v2;
}

// Ops may process the two particles of a pair jointly (Op::PAIRWISE + Op::apply_pair) so that the
// arithmetic of both interleaves in one straight-line block; everything else goes one by one.
template <class Op, class = void>
struct is_pairwise : std::false_type {};
template <class Op>
struct is_pairwise<Op, std::enable_if_t<Op::PAIRWISE>> : std::true_type {};

template <class Op, bool LP>
__device__ __forceinline__ void apply2(Particle &a, Particle &b, const PassParams<Op> &P, const double *sf, const Acc<LP> &acc)
{
    if constexpr (is_pairwise<Op>::value && LP) {
        Op::apply_pair(a, b, P, sf, acc);
    } else {
        Op::apply(a, P, sf, acc);
        Op::apply(b, P, sf, acc);
    }
}

template <class Op, bool LP>
__device__ __forceinline__ void apply4(Particle &a0, Particle &a1, Particle &b0, Particle &b1, const PassParams<Op> &P,
                                       const double *sf, const Acc<LP> &acc)
{
    if constexpr (is_pairwise<Op>::value && LP) {
        Op::apply_quad(a0, a1, b0, b1, P, sf, acc);
    } else {
        Op::apply(a0, P, sf, acc);
        Op::apply(a1, P, sf, acc);
        Op::apply(b0, P, sf, acc);
        Op::apply(b1, P, sf, acc);
    }
}

// shared memory of the TMA row stream of k_pass (TMA = true): per warp one tile of 2 batches x 4 rows x 64 doubles
// (the pairs of both batches of an iteration) and one mbarrier
constexpr int kPassTileDoubles = 2 * 4 * 64;
template <class Op>
__host__ __device__ constexpr size_t pass_tma_bytes() { return (size_t)(op_threads<Op>::value / 32) * (kPassTileDoubles * sizeof(double) + 8); }

// TMA (deposit ops only, experiment of round 2): instead of prefetching the next iteration's rows into registers, every
// warp streams them into its shared-memory tile with cp.async.bulk (one 512 B copy per row and batch) one iteration ahead.
template <class Op, bool LP, bool TMA = false>
__global__ void __launch_bounds__(op_threads<Op>::value) k_pass(const __grid_constant__ PassParams<Op> P)
{
    constexpr int kThreads = op_threads<Op>::value, kNW = kThreads / 32, kH = op_halo<Op>::value,
                  kFH = op_field_halo<Op>::value, kFC = op_field_copies<Op>::value;
    extern __shared__ double smem[];
    const int tid = threadIdx.x;
    const int n = P.m.n;
    const int nh = n + kH, nhf = n + kFH;
    // ---- stage the field dofs this op gathers from (with periodic halo) -------------------
    double *sfield = smem;
    if constexpr (op_has_stage<Op>::value) {
        Op::stage(P, sfield, tid);
    } else {
#pragma unroll
        for (int f = 0; f < Op::NF; ++f)
            for (int i = tid; i < nhf; i += kThreads) {
                const double v = P.fields[f][i < n ? i : i - n];
#pragma unroll
                for (int c = 0; c < kFC; ++c) sfield[(size_t)(f * nhf + i) * kFC + c] = v;
            }
    }
    // ---- zero the block-private accumulators --------------------------------------------
    double *sacc = smem + (size_t)Op::NF * nhf * kFC;
    const int slots = acc_slots<Op>(n);
    const int acc_words = Op::DEPOSIT ? (LP ? slots * 32 * kNW : slots * P.copies) : 0;
    for (int i = tid; i < acc_words; i += kThreads) sacc[i] = 0.0;
    __syncthreads();

    Acc<LP> acc;
    {
        const int warp = tid >> 5, lane = tid & 31;
        acc.p = LP ? sacc + (size_t)warp * slots * 32 + lane : sacc + (size_t)(warp % (P.copies > 0 ? P.copies : 1)) * slots;
        if (kFC > 1) sfield += lane & (kFC - 1);
    }

    // ---- stream the particles: pairs, two batches per iteration -------------------------
    // Deposit ops are limited to a few warps per SM by their shared-memory accumulators, so
    // registers are plentiful: the next iteration's rows are loaded before the current one is
    // processed (software prefetch), which hides the HBM latency those few warps cannot.
    const int64_t n_pairs = P.n_particles >> 1;
    const int64_t T = (int64_t)gridDim.x * kThreads;
    int64_t p = (int64_t)blockIdx.x * kThreads + tid;
    if constexpr (Op::DEPOSIT && TMA) {
        const int warp = tid >> 5, lane = tid & 31;
        double *tile = sacc + acc_words + (size_t)warp * kPassTileDoubles;
        unsigned long long *bar = reinterpret_cast<unsigned long long *>(sacc + acc_words + (size_t)kNW * kPassTileDoubles) + warp;
        if (lane == 0) {
            mbar_init(bar, 1);
            mbar_fence_init();
        }
        __syncwarp();
        const double *rows[4] = {P.r.x, P.r.v1, P.r.v2, P.r.w};
        constexpr int bits[4] = {ROW_X, ROW_V1, ROW_V2, ROW_W};
        constexpr int n_rows = ((Op::READ & ROW_X) != 0) + ((Op::READ & ROW_V1) != 0) + ((Op::READ & ROW_V2) != 0) + ((Op::READ & ROW_W) != 0);
        auto issue = [&](int64_t pw) {   // lane 0: the 32 pairs from pw and the 32 pairs from pw + T, every row the op reads
            fence_proxy_async();
            mbar_expect_tx(bar, n_rows * 2 * 512);
#pragma unroll
            for (int bt = 0; bt < 2; ++bt)
#pragma unroll
                for (int r = 0; r < 4; ++r)
                    if (Op::READ & bits[r]) bulk_g2s(tile + (bt * 4 + r) * 64, rows[r] + 2 * (pw + bt * T), 512, bar);
        };
        auto take = [&](int bt, Particle &u, Particle &v) {
            const double2 *t = reinterpret_cast<const double2 *>(tile + bt * 4 * 64) + lane;
            if (Op::READ & ROW_X) { const double2 q = t[0]; u.x = q.x; v.x = q.y; }
            if (Op::READ & ROW_V1) { const double2 q = t[32]; u.v1 = q.x; v.v1 = q.y; }
            if (Op::READ & ROW_V2) { const double2 q = t[64]; u.v2 = q.x; v.v2 = q.y; }
            if (Op::READ & ROW_W) { const double2 q = t[96]; u.w = q.x; v.w = q.y; }
        };
        int64_t pw = p - lane;   // first pair of the warp's batch
        unsigned phase = 0;
        bool have = pw + 31 + T < n_pairs;
        if (have && lane == 0) issue(pw);
        while (have) {
            Particle a0, a1, b0, b1;
            mbar_wait(bar, phase);
            phase ^= 1u;
            take(0, a0, a1);
            take(1, b0, b1);
            __syncwarp();
            const int64_t q = pw + 2 * T;
            const bool have_next = q + 31 + T < n_pairs;
            if (have_next && lane == 0) issue(q);
            apply4<Op, LP>(a0, a1, b0, b1, P, sfield, acc);
            store_pair<Op>(P.r, pw + lane, a0, a1);
            store_pair<Op>(P.r, pw + lane + T, b0, b1);
            pw = q;
            have = have_next;
        }
        p = pw + lane;
    } else if (Op::DEPOSIT) {
        Particle a0, a1, b0, b1;
        bool have = p + T < n_pairs;
        if (have) {
            load_pair<Op>(P.r, p, a0, a1);
            load_pair<Op>(P.r, p + T, b0, b1);
        }
        while (have) {
            const int64_t q = p + 2 * T;
            const bool have_next = q + T < n_pairs;
            Particle c0, c1, d0, d1;
            if (have_next) {
                load_pair<Op>(P.r, q, c0, c1);
                load_pair<Op>(P.r, q + T, d0, d1);
            }
            apply4<Op, LP>(a0, a1, b0, b1, P, sfield, acc);
            store_pair<Op>(P.r, p, a0, a1);
            store_pair<Op>(P.r, p + T, b0, b1);
            a0 = c0; a1 = c1; b0 = d0; b1 = d1;
            p = q;
            have = have_next;
        }
    } else {
        for (; p + T < n_pairs; p += 2 * T) {
            Particle a0, a1, b0, b1;
            load_pair<Op>(P.r, p, a0, a1);
            load_pair<Op>(P.r, p + T, b0, b1);
            apply4<Op, LP>(a0, a1, b0, b1, P, sfield, acc);
            store_pair<Op>(P.r, p, a0, a1);
            store_pair<Op>(P.r, p + T, b0, b1);
        }
    }
    for (; p < n_pairs; p += T) {
        Particle a0, a1;
        load_pair<Op>(P.r, p, a0, a1);
        apply2<Op, LP>(a0, a1, P, sfield, acc);
        store_pair<Op>(P.r, p, a0, a1);
    }
    if ((P.n_particles & 1) && blockIdx.x == 0 && tid == 0) {
        Particle a;
        load_one<Op>(P.r, P.n_particles - 1, a);
        Op::apply(a, P, sfield, acc);
        store_one<Op>(P.r, P.n_particles - 1, a);
    }

    // ---- hierarchical reduce: lanes -> warps -> one partial vector per block ------------
    if (Op::DEPOSIT) {
        if constexpr (Op::NS > 0) {
            static_assert(Op::NS <= kMaxScalars, "scalar sums");
#pragma unroll
            for (int k = 0; k < Op::NS; ++k) acc.add(Op::NG * nh + k, acc.s[k]);
        }
        __syncthreads();
        const int n_out = acc_outputs<Op>(n);
        double *out = P.partials + (size_t)blockIdx.x * n_out;
        for (int o = tid; o < n_out; o += kThreads) {
            // output o -> first slot s0 (+ its halo image s1, or -1)
            int s0, s1 = -1;
            if (o < Op::NG * n) {
                const int k = o / n, g = o - k * n;
                s0 = k * nh + g;
                if (g < kH) s1 = s0 + n;
            } else {
                s0 = Op::NG * nh + (o - Op::NG * n);
            }
            double s = 0.0;
            if (LP) {
                for (int w = 0; w < kNW; ++w) {
                    const double *base = sacc + (size_t)w * slots * 32;
#pragma unroll 8
                    for (int l = 0; l < 32; ++l) s += base[(size_t)s0 * 32 + ((l + o) & 31)];  // rotated: conflict free
                    if (s1 >= 0)
#pragma unroll 8
                        for (int l = 0; l < 32; ++l) s += base[(size_t)s1 * 32 + ((l + o) & 31)];
                }
            } else {
                for (int c = 0; c < P.copies; ++c) {
                    s += sacc[(size_t)c * slots + s0];
                    if (s1 >= 0) s += sacc[(size_t)c * slots + s1];
                }
            }
            out[o] = s;
        }
    }
}

// out[g] = sum_b partials[b][g], b ascending; one warp per dof, fixed tree -> deterministic.
__global__ void k_reduce_partials(const double *__restrict__ partials, int n_blocks, int n_acc,
                                  double *__restrict__ out);

}  // namespace gempic
