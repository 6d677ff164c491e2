// ops1d.cuh -- per-particle operators of the 1D (1d1v / 1d2v) paths, plugged into k_pass.
//
// Template parameters D0 / D1 are the spline degrees of kernel_smoother_0 / kernel_smoother_1
// (src/hamiltonian_splitting.jl:23-24): e2, j2 and rho live on degree D0, e1, b and j1 on
// degree D1 (src/hamiltonian_splitting_1d2v.jl:70,102,151,163,207-208).
//
// Shared-memory fields and accumulator grids carry a periodic halo of kHalo dofs (pass.cuh),
// so a particle in cell c touches the contiguous slots g0 .. g0+D with g0 = (c-D) mod n.
#pragma once
#include "pass.cuh"

namespace gempic {

// Position of a particle on the mesh: cell, in-cell offset and the first dof of a degree-D
// stencil.  first<D>() wraps once; first<D-1>() follows from it with one compare.
struct Pos {
    int c;
    double t;
};
__device__ __forceinline__ Pos locate(double x, const Mesh1D &m)
{
    Pos p;
    cell_offset(x, m, p.c, p.t);
    return p;
}
template <int D>
__device__ __forceinline__ int first_dof(const Pos &p, const Mesh1D &m)
{
    return wrap_index(p.c - D, m);
}

// sum_k field[g0+k] * b[k], accumulated from 0.0 in k order (evaluate, pmc1d.jl:446-450)
template <int D>
__device__ __forceinline__ double gather_h(const double *__restrict__ field, int g0, const double (&b)[D + 1])
{
    double v = 0.0;
#pragma unroll
    for (int k = 0; k <= D; ++k) v = fma(field[g0 + k], b[k], v);
    return v;
}

// the same gather on a field staged in S lane-interleaved copies (Op::FIELD_COPIES = S, pass.cuh): element g of the
// calling lane's copy sits S doubles after element g-1, so a half-warp's gather is conflict free for any cell pattern
template <int D, int S>
__device__ __forceinline__ double gather_hs(const double *__restrict__ field, int g0, const double (&b)[D + 1])
{
    double v = 0.0;
#pragma unroll
    for (int k = 0; k <= D; ++k) v = fma(field[(g0 + k) * S], b[k], v);
    return v;
}

// grid[g0+k] += ws * N_k   with ws = marker charge * scaling   (add_charge!, pmc1d.jl:261-280)
template <int D, bool LP>
__device__ __forceinline__ void deposit_h(const Acc<LP> &acc, int slot0, const double (&b)[D + 1], double ws)
{
#pragma unroll
    for (int k = 0; k <= D; ++k) acc.add(slot0 + k, ws * b[k]);
}

// add_current_update_v! (pmc1d.jl:296-376, pp arithmetic of :131-230).
// Line integral of the degree-D splines along x_old -> x_new (un-wrapped), split at cell
// boundaries.  With P = prim_pp the per-cell weights are
//     same cell         cell_old:  P(r_new) - P(r_old)
//     forward crossing  cell_old:  P(1) - P(r_old)      cell_new:  P(r_new)        interior: P(1)
//     backward crossing cell_old:      - P(r_old)       cell_new:  P(r_new) - P(1) interior: -P(1)
// all times dx; j[dof] += w*scaling*weight and v -= q/m * weight * B[dof]  (:416-422).
// P(r_old), P(r_new) are evaluated once by every lane; only the extra deposits of crossing
// particles diverge.  `pn` is the position of x_new (trunc, or floor for the 1d1v form :487).
template <int D, bool LP, bool WITH_B>
__device__ __forceinline__ double current_update_v(const Acc<LP> &acc, int slot_base, const double *__restrict__ bfield,
                                                   const Pos &po, const Pos &pn, double ws_dx, double qm_dx, double vi,
                                                   const Mesh1D &m)
{
    double A[D + 1], B[D + 1];
    prim_pp<D>(po.t, A);
    prim_pp<D>(pn.t, B);
    const bool same = po.c == pn.c, fwd = po.c < pn.c;
    const int g_old = first_dof<D>(po, m);
    double bsum = 0.0;
#pragma unroll
    for (int k = 0; k <= D; ++k) {
        const double e = same ? B[k] : (fwd ? prim_full<D>(k) : 0.0);
        const double s = e - A[k];
        acc.add(slot_base + g_old + k, ws_dx * s);
        if (WITH_B) bsum = fma(s, bfield[g_old + k], bsum);
    }
    if (!same) {
        const int g_new = first_dof<D>(pn, m);
#pragma unroll
        for (int k = 0; k <= D; ++k) {
            const double s = fwd ? B[k] : B[k] - prim_full<D>(k);
            acc.add(slot_base + g_new + k, ws_dx * s);
            if (WITH_B) bsum = fma(s, bfield[g_new + k], bsum);
        }
        // whole cells strictly between (rare: |dt v| > dx)
        const int lo = fwd ? po.c : pn.c, hi = fwd ? pn.c : po.c;
        for (int c = lo + 1; c < hi; ++c) {
            const int g = wrap_index(c - D, m);
#pragma unroll
            for (int k = 0; k <= D; ++k) {
                const double s = fwd ? prim_full<D>(k) : -prim_full<D>(k);
                acc.add(slot_base + g + k, ws_dx * s);
                if (WITH_B) bsum = fma(s, bfield[g + k], bsum);
            }
        }
    }
    if (WITH_B) vi = fma(-qm_dx, bsum, vi);
    return vi;
}

// ---- building blocks shared by the per-operator and the fused passes --------------------
// operatorHE kick at a located particle (hamiltonian_splitting_1d2v.jl:198-215)
template <int D0, int D1>
__device__ __forceinline__ void kick_e(Particle &p, int g0, int g1, const double (&b0)[D0 + 1], const double (&b1)[D1 + 1],
                                       const double *se1, const double *se2, double dtqm)
{
    p.v1 = fma(dtqm, gather_h<D1>(se1, g1, b1), p.v1);
    p.v2 = fma(dtqm, gather_h<D0>(se2, g0, b0), p.v2);
}
// operatorHp2 at a located particle (:141-167): v1 += dt q/m v2 B(x); j2 += w v2 N(x)
template <int D0, int D1, bool LP>
__device__ __forceinline__ void push_p2(Particle &p, int g0, int g1, const double (&b0)[D0 + 1], const double (&b1)[D1 + 1],
                                        const double *sb, double dtqm, double ws0, const Acc<LP> &acc, int slot_base)
{
    const double bf = gather_h<D1>(sb, g1, b1);
    p.v1 = fma(dtqm * p.v2, bf, p.v1);
    deposit_h<D0, LP>(acc, slot_base + g0, b0, ws0 * p.v2);
}
// first dof of the degree-(D0) and degree-(D1) stencils of one position (D1 in {D0-1, D0})
template <int D0, int D1>
__device__ __forceinline__ void first_dofs(const Pos &ps, const Mesh1D &m, int &g0, int &g1)
{
    g0 = first_dof<D0>(ps, m);
    g1 = (D1 == D0) ? g0 : wrap_next(g0 + (D0 - D1), m.n);
}

// ---- operatorHE{1,2}  (hamiltonian_splitting_1d2v.jl:198-215), also Boris push_v_epart!
//      (hamiltonian_splitting_boris.jl:189-204).  fields: [e1, e2]
template <int D0, int D1>
struct OpHE {
    static constexpr int READ = ROW_X | ROW_V1 | ROW_V2, WRITE = ROW_V1 | ROW_V2;
    static constexpr int NF = 2, NG = 0, NS = 0;
    static constexpr bool DEPOSIT = false;
    struct Params { double dtqm; };
    template <bool LP>
    static __device__ __forceinline__ void apply(Particle &p, const PassParams<OpHE> &P, const double *sf, const Acc<LP> &)
    {
        const int nh = P.m.n + kHalo;
        const Pos ps = locate(p.x, P.m);
        int g0, g1;
        first_dofs<D0, D1>(ps, P.m, g0, g1);
        double b1[D1 + 1], b0[D0 + 1];
        basis_pp<D1>(ps.t, b1);
        basis_pp<D0>(ps.t, b0);
        kick_e<D0, D1>(p, g0, g1, b0, b1, sf, sf + nh, P.op.dtqm);
    }
};

// ---- operatorHp2{1,2}  (hamiltonian_splitting_1d2v.jl:141-167).  fields: [b]; deposit j2 (D0)
template <int D0, int D1>
struct OpHp2 {
    static constexpr int READ = ROW_X | ROW_V1 | ROW_V2 | ROW_W, WRITE = ROW_V1;
    static constexpr int NF = 1, NG = 1, NS = 0;
    static constexpr bool DEPOSIT = true;
    struct Params { double dtqm, wscale0; };   // wscale0 = charge * common_weight * scaling_0
    template <bool LP>
    static __device__ __forceinline__ void apply(Particle &p, const PassParams<OpHp2> &P, const double *sf, const Acc<LP> &acc)
    {
        const Pos ps = locate(p.x, P.m);
        int g0, g1;
        first_dofs<D0, D1>(ps, P.m, g0, g1);
        double b1[D1 + 1], b0[D0 + 1];
        basis_pp<D1>(ps.t, b1);
        basis_pp<D0>(ps.t, b0);
        push_p2<D0, D1, LP>(p, g0, g1, b0, b1, sf, P.op.dtqm, p.w * P.op.wscale0, acc, 0);
    }
};

// ---- operatorHp1{1,2}  (hamiltonian_splitting_1d2v.jl:48-106).  fields: [b];
//      deposit j1 (D1) into grid 0 and, when RHO, rho (D0) into grid 1.
//      RHO=false is used inside strang_splitting!, where that rho is dead data.
template <int D0, int D1, bool RHO>
struct OpHp1 {
    static constexpr int READ = ROW_X | ROW_V1 | ROW_V2 | ROW_W, WRITE = ROW_X | ROW_V2;
    static constexpr int NF = 1, NG = RHO ? 2 : 1, NS = 0;
    static constexpr bool DEPOSIT = true;
    struct Params { double dt, qm_dx, wscale0, wscale1_dx; };   // wscale1_dx = charge*cw*scaling_1*dx
    template <bool LP>
    static __device__ __forceinline__ void apply(Particle &p, const PassParams<OpHp1> &P, const double *sf, const Acc<LP> &acc)
    {
        const double x_new = fma(P.op.dt, p.v1, p.x);
        const Pos po = locate(p.x, P.m);
        Pos pn = locate(x_new, P.m);
        p.v2 = current_update_v<D1, LP, true>(acc, 0, sf, po, pn, p.w * P.op.wscale1_dx, P.op.qm_dx, p.v2, P.m);
        p.x = mod_julia(x_new, P.m.Lx);
        if (RHO) {
            if (p.x != x_new) pn = locate(p.x, P.m);
            double b0[D0 + 1];
            basis_pp<D0>(pn.t, b0);
            deposit_h<D0, LP>(acc, (P.m.n + kHalo) + first_dof<D0>(pn, P.m), b0, p.w * P.op.wscale0);
        }
    }
};

// ---- fused [HE x NHE, Hp2, Hp1, Hp2] of one Strang step (hamiltonian_splitting.jl:98-108).
// Inside strang_splitting! the operators HE(dt/2), Hp2(dt/2), Hp1(dt), Hp2(dt/2) only exchange
// data through the particles: the HE kick reads e1, e2, the three pushes read b (which
// compute_b_from_e! updated from the *old* e2 before the pass), and their deposits j2, j1, j2
// feed field solves whose results are first read by the trailing HE.  One pass therefore
// performs all four: 32 B read + 24 B written per particle instead of 4 passes (168 B).
// NHE = 2 additionally folds in the trailing HE of the previous step (fields e1T, e2T), which
// is separated from this step's leading HE only by field-only updates.
//   source fields: [e1, e2] x NHE, b      grids: j2 (both Hp2 half steps), j1
//
// What bounds this pass is the shared-memory pipe (ncu r01c/r01d/r01m: l1tex data-pipe wavefronts
// 87-96 % of peak), so it is laid out to touch as few shared-memory words per particle as possible:
//   * the fields are staged in pp form (the reference's b_to_pp / evaluate_pp, splinepp.jl:241-285,
//     pmc1d.jl:242-250): per cell the D+1 polynomial coefficients of  E1 = sum_h dt_h q/m e1^(h),
//     E2 likewise, and of the cell-cumulative ANTIDERIVATIVE G of b.  One set of D1+1 coefficients of b
//     then serves the Hp2 gather at the old position, the line integral  int b dx = G(new) - G(old)  of
//     Hp1 and the Hp2 gather at the new position: 3 loads instead of 10 (the lanes whose particle
//     changes cell load the new cell's set and the two cell constants with predicated LDS), and the
//     kicks of all NHE electric fields cost one Horner each;
//   * every staged coefficient exists in 16 lane-interleaved copies (Op::FIELD_COPIES): a gather is
//     conflict free for any cell pattern;
//   * deposits are read-modify-writes of lane-private grids (pass.cuh), one contiguous window
//     per grid and particle, issued after the branch-free arithmetic of a whole quad:
//       j2           D0+1 slots at the old cell, holding BOTH Hp2 deposits when the particle stays in
//                    its cell; only the lanes that changed cell touch a second window at the new one.
//                    Both half steps add into ONE grid: the two compute_e_from_j!(e2, ., 2) solves are
//                    linear and nothing reads e2 in between, so e2 -= M0^-1 (dt/2)(j2a + j2b)/dx.
//                    (j_dofs[2] as the reference leaves it -- dt/2 * j2b -- is rebuilt on demand from
//                    the particles, hs1d.cu materialise_j2.)
//       j1           D1+2 slots starting at min(cell_old, cell_new): the old-cell and new-cell
//                    segments of add_current_update_v! merged (the last one predicated: it is zero
//                    unless the cell changed);
//   * one block of 8 warps per SM: 8 x 2 lane-private grids + the coefficient copies = 191 KB.
// Particles that move more than one cell (or sit outside one period) take the general
// per-particle code (apply) instead, which reads the dof-form fields from global memory.
template <int D0, int D1>
struct FusedWork {
    double x, v1, v2;
    int g0, gw, g0n;
    bool slow, crossed, moved;   // crossed: x -> x_new left the cell (un-wrapped); moved: the wrapped new cell differs
    double d2a[D0 + 1], dj1[D1 + 2];
    double t2b, tn;   // second Hp2 deposit of a particle that changed cell: w v2 and the new offset (rebuilt in commit)
};

// (c - D) mod n for c - D in [-n, 2n)
__device__ __forceinline__ int wrap_near(int g, int n)
{
    g = g < 0 ? g + n : g;
    return g >= n ? g - n : g;
}

// coefficient of t^j of piece k of the degree-D uniform B-spline (columns of SplinePP.poly_coeffs,
// splinepp.jl:39-69, lowest power first)
template <int D>
__host__ __device__ constexpr double pp_coef(int k, int j)
{
    if (D == 0) return 1.0;
    if (D == 1) return k == 0 ? (j == 0 ? 1.0 : -1.0) : (j == 0 ? 0.0 : 1.0);
    if (D == 2) {
        constexpr double c[3][3] = {{0.5, -1.0, 0.5}, {0.5, 1.0, -1.0}, {0.0, 0.0, 0.5}};
        return c[k][j];
    }
    constexpr double c[4][4] = {{1.0 / 6.0, -0.5, 0.5, -1.0 / 6.0}, {4.0 / 6.0, 0.0, -1.0, 0.5},
                                {1.0 / 6.0, 0.5, 0.5, -0.5}, {0.0, 0.0, 0.0, 1.0 / 6.0}};
    return c[k][j];
}

// p(t) = sum_j c[j*S] t^j
template <int D, int S>
__device__ __forceinline__ double horner_s(const double *__restrict__ c, double t)
{
    double v = c[D * S];
#pragma unroll
    for (int j = D - 1; j >= 0; --j) v = fma(v, t, c[j * S]);
    return v;
}
// polynomial of a degree-D spline field in cell c (b_to_pp, splinepp.jl:241-261): co[j] = sum_k dof((c-D+k) mod n) pp_coef(k, j)
template <int D, class F>
__device__ __forceinline__ void cell_poly(int c, int n, F dof, double *co)
{
#pragma unroll
    for (int j = 0; j <= D; ++j) co[j] = 0.0;
#pragma unroll
    for (int k = 0; k <= D; ++k) {
        int g = c - D + k;
        g = g < 0 ? g + n : g;
        g = g >= n ? g - n : g;
        const double d = dof(g);
#pragma unroll
        for (int j = 0; j <= D; ++j) co[j] = fma(d, pp_coef<D>(k, j), co[j]);
    }
}

// h = coefficients of p with Q(t) = t p(t) the primitive of a field: returns Q(t)
template <int D>
__device__ __forceinline__ double prim_eval(const double (&h)[D + 1], double t)
{
    double p = h[D];
#pragma unroll
    for (int j = D - 1; j >= 0; --j) p = fma(p, t, h[j]);
    return p * t;
}
// the field itself, Q'(t) = p(t) + t p'(t)
template <int D>
__device__ __forceinline__ double prim_deriv(const double (&h)[D + 1], double t)
{
    double p = h[D], dp = 0.0;
#pragma unroll
    for (int j = D - 1; j >= 0; --j) {
        dp = (j == D - 1) ? p : fma(dp, t, p);
        p = fma(p, t, h[j]);
    }
    return D == 0 ? p : fma(dp, t, p);
}
// both at the same t
template <int D>
__device__ __forceinline__ void prim_both(const double (&h)[D + 1], double t, double &Q, double &f)
{
    double p = h[D], dp = 0.0;
#pragma unroll
    for (int j = D - 1; j >= 0; --j) {
        dp = (j == D - 1) ? p : fma(dp, t, p);
        p = fma(p, t, h[j]);
    }
    Q = p * t;
    f = D == 0 ? p : fma(dp, t, p);
}

// v = *p if cond (p in shared memory).  A predicated LDS: no branch (the four particle streams of a quad stay
// interleaved) and no shared-memory wavefront for the lanes that keep their value.
__device__ __forceinline__ void lds_if(double &v, const double *p, bool cond)
{
    asm("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q ld.shared.f64 %0, [%1];\n\t}"
        : "+d"(v)
        : "r"((unsigned)__cvta_generic_to_shared(p)), "r"((int)cond));
}
// *p += v if cond: predicated read-modify-write of a lane-private slot, ordered after all earlier shared-memory stores
__device__ __forceinline__ void rmw_if(double *p, double v, bool cond)
{
    asm volatile("{\n\t.reg .pred q;\n\t.reg .f64 t;\n\tsetp.ne.b32 q, %2, 0;\n\t@q ld.shared.f64 t, [%0];\n\t"
                 "add.f64 t, t, %1;\n\t@q st.shared.f64 [%0], t;\n\t}"
                 :
                 : "r"((unsigned)__cvta_generic_to_shared(p)), "d"(v), "r"((int)cond)
                 : "memory");
}

template <int D0, int D1, int NHE>
struct OpStrangFused {
    static constexpr int READ = ROW_X | ROW_V1 | ROW_V2 | ROW_W, WRITE = ROW_X | ROW_V1 | ROW_V2;
    static constexpr int NSRC = 2 * NHE + 1;   // global field vectors in PassParams::fields
    static constexpr int NC1 = D1 + 1, NC0 = D0 + 1;
    static constexpr int NF = 2 * NC1 + NC0 + 1;   // staged coefficients per cell: E1 (NC1), E2 (NC0), antiderivative of b (1 + NC1)
    static constexpr int NG = 2, NS = 0;       // grids: j2 (first + second Hp2), j1
    static constexpr bool DEPOSIT = true;
    static constexpr bool PAIRWISE = true;
    static constexpr int THREADS = 256;
    static constexpr int FIELD_COPIES = 16;
    static constexpr int HALO = (D0 > D1 + 1) ? D0 : D1 + 1;   // widest window minus one
    static constexpr int FIELD_HALO = 2;                       // pp tables cover the cells -1 .. n
    static constexpr bool CUSTOM_STAGE = true;
    static constexpr int FC = FIELD_COPIES;
    static constexpr int OFF_E2 = NC1 * FC, OFF_C = (NC1 + NC0) * FC, OFF_B = OFF_C + FC, CELL = NF * FC;   // doubles
    struct Params { double dtqm_e[2], dtqm_p2, dt, qm_dx, wscale0, wscale1_dx; };
    using PP = PassParams<OpStrangFused>;

    // combined electric dofs  sum_h dtqm_e[h] e^(h)[g]  (the kicks are linear in the dofs)
    static __device__ __forceinline__ double dof_e1(const PP &P, int g)
    {
        double c = P.op.dtqm_e[0] * P.fields[0][g];
        if (NHE == 2) c = fma(P.op.dtqm_e[1], P.fields[2][g], c);
        return c;
    }
    static __device__ __forceinline__ double dof_e2(const PP &P, int g)
    {
        double c = P.op.dtqm_e[0] * P.fields[1][g];
        if (NHE == 2) c = fma(P.op.dtqm_e[1], P.fields[3][g], c);
        return c;
    }
    static __device__ __forceinline__ double dof_b(const PP &P, int g) { return P.fields[2 * NHE][g]; }

    // staged tables, cell i = c + 1 for c = -1 .. n (periodic), coefficient q, copy k:
    //   sfield[(i * NF + q) * FC + k]
    //   q in [0, NC1): E1 pp     [NC1, NC1+NC0): E2 pp     then the antiderivative of b,  G_c(t) = C_c + t sum_j h_j t^j
    //   with h_j = c_j / (j+1) and C_c = int of b from the left end of cell 0 to the left end of cell c: the line
    //   integral of b along ANY path is G(new) - G(old), whatever the number of cell boundaries crossed.
    static __device__ __forceinline__ void stage(const PP &P, double *sfield, int tid)
    {
        const int n = P.m.n;
        for (int i = tid; i < n + FIELD_HALO; i += THREADS) {
            const int c = i - 1;
            double co[NF];
            cell_poly<D1>(c, n, [&](int g) { return dof_e1(P, g); }, co);
            cell_poly<D0>(c, n, [&](int g) { return dof_e2(P, g); }, co + NC1);
            cell_poly<D1>(c, n, [&](int g) { return dof_b(P, g); }, co + NC1 + NC0 + 1);
#pragma unroll
            for (int j = 1; j <= D1; ++j) co[NC1 + NC0 + 1 + j] *= 1.0 / (double)(j + 1);
            {   // C_c: whole-cell integrals of the cells 0 .. c-1 (cell -1: minus that of cell n-1), in a fixed order
                double cum = 0.0;
                const int c_hi = c < 0 ? n : c, c_lo = c < 0 ? n - 1 : 0;
                for (int cc = c_lo; cc < c_hi; ++cc) {
                    double q1 = 0.0;
#pragma unroll
                    for (int k = 0; k <= D1; ++k) {
                        int g = cc - D1 + k;
                        g = g < 0 ? g + n : g;
                        q1 = fma(dof_b(P, g), prim_full<D1>(k), q1);
                    }
                    cum += q1;
                }
                co[NC1 + NC0] = c < 0 ? -cum : cum;
            }
#pragma unroll
            for (int q = 0; q < NF; ++q)
#pragma unroll
                for (int k = 0; k < FC; ++k) sfield[((size_t)i * NF + q) * FC + k] = co[q];
        }
    }

    // general per-particle form (any displacement): reads the dof-form fields from global memory
    template <bool LP>
    static __device__ __forceinline__ void apply(Particle &p, const PP &P, const double *, const Acc<LP> &acc)
    {
        const int n = P.m.n, nh = n + HALO;
        Pos ps = locate(p.x, P.m);
        int g0, g1;
        first_dofs<D0, D1>(ps, P.m, g0, g1);
        double b1[D1 + 1], b0[D0 + 1];
        basis_pp<D1>(ps.t, b1);
        basis_pp<D0>(ps.t, b0);
        {
            double s1 = 0.0, s2 = 0.0, sb = 0.0;
            int g = g1;
#pragma unroll
            for (int k = 0; k <= D1; ++k) {
                s1 = fma(dof_e1(P, g), b1[k], s1);
                sb = fma(dof_b(P, g), b1[k], sb);
                g = wrap_next(g + 1, n);
            }
            g = g0;
#pragma unroll
            for (int k = 0; k <= D0; ++k) {
                s2 = fma(dof_e2(P, g), b0[k], s2);
                g = wrap_next(g + 1, n);
            }
            p.v1 += s1;
            p.v2 += s2;
            p.v1 = fma(P.op.dtqm_p2 * p.v2, sb, p.v1);
        }
        const double ws0 = p.w * P.op.wscale0;
        deposit_h<D0, LP>(acc, g0, b0, ws0 * p.v2);
        // Hp1: add_current_update_v! cell by cell
        const double x_new = fma(P.op.dt, p.v1, p.x);
        Pos pn = locate(x_new, P.m);
        {
            double A[D1 + 1], B[D1 + 1];
            prim_pp<D1>(ps.t, A);
            prim_pp<D1>(pn.t, B);
            const bool same = ps.c == pn.c, fwd = ps.c < pn.c;
            const double ws_dx = p.w * P.op.wscale1_dx;
            double bsum = 0.0;
            int g = first_dof<D1>(ps, P.m);
#pragma unroll
            for (int k = 0; k <= D1; ++k) {
                const double s = (same ? B[k] : (fwd ? prim_full<D1>(k) : 0.0)) - A[k];
                acc.add(nh + g + k, ws_dx * s);
                bsum = fma(s, dof_b(P, wrap_next(g + k, n)), bsum);
            }
            if (!same) {
                g = first_dof<D1>(pn, P.m);
#pragma unroll
                for (int k = 0; k <= D1; ++k) {
                    const double s = fwd ? B[k] : B[k] - prim_full<D1>(k);
                    acc.add(nh + g + k, ws_dx * s);
                    bsum = fma(s, dof_b(P, wrap_next(g + k, n)), bsum);
                }
                const int lo = fwd ? ps.c : pn.c, hi = fwd ? pn.c : ps.c;
                for (int c = lo + 1; c < hi; ++c) {
                    g = wrap_index(c - D1, P.m);
#pragma unroll
                    for (int k = 0; k <= D1; ++k) {
                        const double s = fwd ? prim_full<D1>(k) : -prim_full<D1>(k);
                        acc.add(nh + g + k, ws_dx * s);
                        bsum = fma(s, dof_b(P, wrap_next(g + k, n)), bsum);
                    }
                }
            }
            p.v2 = fma(-P.op.qm_dx, bsum, p.v2);
        }
        p.x = mod_julia(x_new, P.m.Lx);
        if (p.x != x_new) pn = locate(p.x, P.m);
        // second Hp2 at the new position
        first_dofs<D0, D1>(pn, P.m, g0, g1);
        basis_pp<D1>(pn.t, b1);
        basis_pp<D0>(pn.t, b0);
        {
            double sb = 0.0;
            int g = g1;
#pragma unroll
            for (int k = 0; k <= D1; ++k) {
                sb = fma(dof_b(P, g), b1[k], sb);
                g = wrap_next(g + 1, n);
            }
            p.v1 = fma(P.op.dtqm_p2 * p.v2, sb, p.v1);
        }
        deposit_h<D0, LP>(acc, g0, b0, ws0 * p.v2);
    }

    // out of line and by value, so that the rare call does not pin the particles of the fast
    // path to local memory
    struct XV { double x, v1, v2; };
    template <bool LP>
    static __device__ __noinline__ XV apply_general(double x, double v1, double v2, double w, const PP &P, double *accp)
    {
        Particle p{x, v1, v2, w};
        Acc<LP> acc{accp};
        apply<LP>(p, P, nullptr, acc);
        return XV{p.x, p.v1, p.v2};
    }

    // branch-free arithmetic of one particle; no shared-memory writes
    static __device__ __forceinline__ FusedWork<D0, D1> work(const Particle &p, const PP &P, const double *sf)
    {
        FusedWork<D0, D1> W;
        const int n = P.m.n;
        double v1 = p.v1, v2 = p.v2;
        // ---- HE kick(s) and first Hp2 at the old position: cell polynomials of E1, E2 and of b
        const Pos po = locate(p.x, P.m);
        const int io = min((unsigned)(po.c + 1), (unsigned)(n + 1));   // table cell (clamped: far-out particles are `slow`)
        const double *co = sf + (size_t)io * CELL;
        const int g0 = wrap_near(po.c - D0, n);
        double A[D1 + 1], b0[D0 + 1];
        prim_pp<D1>(po.t, A);
        if constexpr (D1 == D0 - 1) basis_from_prim<D0>(A, b0);
        else basis_pp<D0>(po.t, b0);
        v1 += horner_s<D1, FC>(co, po.t);
        v2 += horner_s<D0, FC>(co + OFF_E2, po.t);
        double ho[D1 + 1];
#pragma unroll
        for (int j = 0; j <= D1; ++j) ho[j] = co[OFF_B + j * FC];
        double Qo, bo;
        prim_both<D1>(ho, po.t, Qo, bo);
        v1 = fma(P.op.dtqm_p2 * v2, bo, v1);
        const double ws0 = p.w * P.op.wscale0;
        {
            const double t2 = ws0 * v2;
#pragma unroll
            for (int k = 0; k <= D0; ++k) W.d2a[k] = t2 * b0[k];
        }
        W.g0 = g0;
        // ---- Hp1: line integral over x -> x_new.  Window of D1+2 dofs from cmin = min(cells);
        // with Phi_m(s, t) = int of spline m from the window start to (cell cmin+s, offset t):
        //   Phi_m(0, t) = P_m(t) [m <= D1]          Phi_m(1, t) = P_m(1) [m <= D1] + P_{m-1}(t) [m >= 1]
        // the weight of dof m is Phi_m(new) - Phi_m(old)   (= the per-cell sums of pmc1d.jl:316-373)
        const double x_new = fma(P.op.dt, v1, p.x);
        const Pos pn = locate(x_new, P.m);
        double B[D1 + 1];
        prim_pp<D1>(pn.t, B);
        const int cmin = min(po.c, pn.c);
        const bool o1 = po.c != cmin, n1 = pn.c != cmin;
        const int gw = wrap_near(cmin - D1, n);
        const double ws1 = p.w * P.op.wscale1_dx;
#pragma unroll
        for (int m = 0; m <= D1 + 1; ++m) {
            const double F = m <= D1 ? prim_full<D1>(m <= D1 ? m : 0) : 0.0;
            const double n0 = m <= D1 ? B[m <= D1 ? m : 0] : 0.0, o0 = m <= D1 ? A[m <= D1 ? m : 0] : 0.0;
            const double nn = m == 0 ? F : (m <= D1 ? F + B[m >= 1 ? m - 1 : 0] : B[D1]);
            const double oo = m == 0 ? F : (m <= D1 ? F + A[m >= 1 ? m - 1 : 0] : A[D1]);
            W.dj1[m] = ws1 * ((n1 ? nn : n0) - (o1 ? oo : o0));
        }
        W.gw = gw;
        W.crossed = o1 | n1;   // otherwise dj1[D1+1] = 0
        // v2 -= q/m int b dx along the same path (:416-422) = G(new) - G(old); the cell constants cancel unless a
        // boundary was crossed, so they (and the new cell's coefficients) are only loaded by the lanes that crossed
        const int in = min((unsigned)(pn.c + 1), (unsigned)(n + 1));
        const bool crossed = in != io;
        const double *cn = sf + (size_t)in * CELL;
        double hn[D1 + 1], Co = 0.0, Cn = 0.0;
#pragma unroll
        for (int j = 0; j <= D1; ++j) {
            hn[j] = ho[j];
            lds_if(hn[j], cn + OFF_B + j * FC, crossed);
        }
        lds_if(Co, co + OFF_C, crossed);
        lds_if(Cn, cn + OFF_C, crossed);
        v2 = fma(-P.op.qm_dx, (prim_eval<D1>(hn, pn.t) - Qo) + (Cn - Co), v2);
        // x = mod(x_new, Lx) for x_new within one period of the domain (:79)
        const double L = P.m.Lx;
        const double x = x_new < 0.0 ? x_new + L : (x_new >= L ? x_new - L : x_new);
        // ---- second Hp2 at the wrapped new position
        const Pos p2 = locate(x, P.m);
        const int g0n = wrap_near(p2.c - D0, n);
        const int i2 = min((unsigned)(p2.c + 1), (unsigned)(n + 1));
        {   // another polynomial only after the trunc quirk at x_new < 0: cell 0 with a negative offset
            // wraps into cell n-1.  (x_new >= L: cell n of the table is cell 0.)
            const bool quirk = i2 != in && !(in == n + 1 && i2 == 1);
            const double *c2 = sf + (size_t)i2 * CELL + OFF_B;
#pragma unroll
            for (int j = 0; j <= D1; ++j) lds_if(hn[j], c2 + j * FC, quirk);
        }
        v1 = fma(P.op.dtqm_p2 * v2, prim_deriv<D1>(hn, p2.t), v1);
        basis_pp<D0>(p2.t, b0);
        {   // both Hp2 deposits go to the same grid: a particle that stays in its cell adds them through ONE window
            const double t2 = ws0 * v2;
            W.moved = g0n != g0;
            double t2s = W.moved ? 0.0 : t2;
            asm("" : "+d"(t2s));   // keep the select on the scalar (the compiler would otherwise select every product)
#pragma unroll
            for (int k = 0; k <= D0; ++k) {
                W.d2a[k] = fma(t2s, b0[k], W.d2a[k]);
            }
            W.t2b = t2;
            W.tn = p2.t;
        }
        W.g0n = g0n;
        W.x = x;
        W.v1 = v1;
        W.v2 = v2;
        // fast-path domain: both cells within one cell of the grid, at most one boundary crossed
        const unsigned span = (unsigned)(n + 2);
        const int dc = pn.c - po.c;
        W.slow = (unsigned)(po.c + 1) >= span || (unsigned)(pn.c + 1) >= span || dc > 1 || dc < -1 || n < 8;
        return W;
    }

    // the read-modify-write windows of one particle (lane-private slots: s -> s*32): D0+1 slots of j2 at the old cell
    // (both Hp2 deposits unless the particle changed cell), D1+1 slots of j1, and -- only for the lanes that left their
    // cell, so without shared-memory wavefronts for the others -- the last j1 slot and the j2 window at the new cell.
    static __device__ __forceinline__ void commit(const FusedWork<D0, D1> &W, double *acc, int nh)
    {
        double *q0 = acc + (size_t)W.g0 * 32, *q1 = acc + (size_t)(nh + W.gw) * 32, *q2 = acc + (size_t)W.g0n * 32;
        double r0[D0 + 1], r1[D1 + 1];
#pragma unroll
        for (int k = 0; k <= D0; ++k) r0[k] = q0[k * 32];
#pragma unroll
        for (int m = 0; m <= D1; ++m) r1[m] = q1[m * 32];
#pragma unroll
        for (int k = 0; k <= D0; ++k) q0[k * 32] = r0[k] + W.d2a[k];
#pragma unroll
        for (int m = 0; m <= D1; ++m) q1[m * 32] = r1[m] + W.dj1[m];
        rmw_if(q1 + (D1 + 1) * 32, W.dj1[D1 + 1], W.crossed);
        if (W.moved) {   // keeping these four products live for all lanes spills (254 registers); the few lanes rebuild them
            double b0[D0 + 1];
            basis_pp<D0>(W.tn, b0);
#pragma unroll
            for (int k = 0; k <= D0; ++k) q2[k * 32] += W.t2b * b0[k];
        }
    }

    template <bool LP>
    static __device__ __forceinline__ void general(Particle &a, const PP &P, const Acc<LP> &acc)
    {
        const XV r = apply_general<LP>(a.x, a.v1, a.v2, a.w, P, acc.p);
        a.x = r.x; a.v1 = r.v1; a.v2 = r.v2;
    }

    template <bool LP>
    static __device__ __forceinline__ void apply_pair(Particle &a, Particle &b, const PP &P, const double *sf, const Acc<LP> &acc)
    {
        const FusedWork<D0, D1> Wa = work(a, P, sf);
        const FusedWork<D0, D1> Wb = work(b, P, sf);
        if (__builtin_expect(Wa.slow | Wb.slow, 0)) {
            general<LP>(a, P, acc);
            general<LP>(b, P, acc);
            return;
        }
        const int nh = P.m.n + HALO;
        commit(Wa, acc.p, nh);
        commit(Wb, acc.p, nh);
        a.x = Wa.x; a.v1 = Wa.v1; a.v2 = Wa.v2;
        b.x = Wb.x; b.v1 = Wb.v1; b.v2 = Wb.v2;
    }

    // both batches of an iteration: four independent instruction streams (work_quad), then the deposits (commit_quad)
    struct Quad { FusedWork<D0, D1> a, b, c, d; };
    static __device__ __forceinline__ void work_quad(Quad &W, const Particle &a, const Particle &b, const Particle &c,
                                                     const Particle &d, const PP &P, const double *sf)
    {
        W.a = work(a, P, sf);
        W.b = work(b, P, sf);
        W.c = work(c, P, sf);
        W.d = work(d, P, sf);
    }
    template <bool LP>
    static __device__ __forceinline__ void commit_quad(const Quad &W, Particle &a, Particle &b, Particle &c, Particle &d,
                                                       const PP &P, const Acc<LP> &acc)
    {
        if (__builtin_expect(W.a.slow | W.b.slow | W.c.slow | W.d.slow, 0)) {
            general<LP>(a, P, acc);
            general<LP>(b, P, acc);
            general<LP>(c, P, acc);
            general<LP>(d, P, acc);
            return;
        }
        const int nh = P.m.n + HALO;
        commit(W.a, acc.p, nh);
        commit(W.b, acc.p, nh);
        commit(W.c, acc.p, nh);
        commit(W.d, acc.p, nh);
        a.x = W.a.x; a.v1 = W.a.v1; a.v2 = W.a.v2;
        b.x = W.b.x; b.v1 = W.b.v1; b.v2 = W.b.v2;
        c.x = W.c.x; c.v1 = W.c.v1; c.v2 = W.c.v2;
        d.x = W.d.x; d.v1 = W.d.v1; d.v2 = W.d.v2;
    }
    template <bool LP>
    static __device__ __forceinline__ void apply_quad(Particle &a, Particle &b, Particle &c, Particle &d, const PP &P,
                                                      const double *sf, const Acc<LP> &acc)
    {
        Quad W;
        work_quad(W, a, b, c, d, P, sf);
        commit_quad<LP>(W, a, b, c, d, P, acc);
    }
};

// ---- j2 = sum w v2 N(x) at the current particle state (D0): what the second operatorHp2 of a Strang step
//      leaves in j_dofs[2] before the dt scaling (hamiltonian_splitting_1d2v.jl:141-175).  The fused pass adds both
//      half-step currents into one grid; this rebuilds the reference's j_dofs[2] when somebody asks for it.
template <int D0>
struct OpJ2 {
    static constexpr int READ = ROW_X | ROW_V2 | ROW_W, WRITE = 0;
    static constexpr int NF = 0, NG = 1, NS = 0;
    static constexpr bool DEPOSIT = true;
    struct Params { double wscale0; };
    template <bool LP>
    static __device__ __forceinline__ void apply(Particle &p, const PassParams<OpJ2> &P, const double *, const Acc<LP> &acc)
    {
        const Pos ps = locate(p.x, P.m);
        double b0[D0 + 1];
        basis_pp<D0>(ps.t, b0);
        deposit_h<D0, LP>(acc, first_dof<D0>(ps, P.m), b0, (p.w * P.op.wscale0) * p.v2);
    }
};

// ---- the same deposit fused with the deferred operatorHE kick that follows it (pg_sync, hs1d.cu).  fields: [e1, e2]
template <int D0, int D1>
struct OpHEJ2 {
    static constexpr int READ = ROW_X | ROW_V1 | ROW_V2 | ROW_W, WRITE = ROW_V1 | ROW_V2;
    static constexpr int NF = 2, NG = 1, NS = 0;
    static constexpr bool DEPOSIT = true;
    struct Params { double dtqm, wscale0; };
    template <bool LP>
    static __device__ __forceinline__ void apply(Particle &p, const PassParams<OpHEJ2> &P, const double *sf, const Acc<LP> &acc)
    {
        const int nh = P.m.n + kHalo;
        const Pos ps = locate(p.x, P.m);
        int g0, g1;
        first_dofs<D0, D1>(ps, P.m, g0, g1);
        double b1[D1 + 1], b0[D0 + 1];
        basis_pp<D1>(ps.t, b1);
        basis_pp<D0>(ps.t, b0);
        deposit_h<D0, LP>(acc, g0, b0, (p.w * P.op.wscale0) * p.v2);
        kick_e<D0, D1>(p, g0, g1, b0, b1, sf, sf + nh, P.op.dtqm);
    }
};

// ---- the particle part of a diagnostics loop body after a fused strang_splitting! in ONE pass (48 B/particle):
//      examples/strong_landau_damping_1d2v.jl:46-59 calls strang_splitting!; solve_poisson!; write_step! every step.
//      The deferred trailing operatorHE kick (with the j_dofs[2] deposit that must precede it, OpHEJ2), the rho deposit
//      of solve_poisson! (diagnostics.jl:24-28) and the five particle sums of write_step! (:45-92,197-211, taken with
//      the kicked velocities) read the same rows.  fields: [e1T, e2T (kick), e1, e2, b (sums)]; grids: j2, rho;
//      scalars: KE, P1, P2, transfer, vvb.
template <int D0, int D1>
struct OpLoopTail {
    static constexpr int READ = ROW_X | ROW_V1 | ROW_V2 | ROW_W, WRITE = ROW_V1 | ROW_V2;
    static constexpr int NF = 5, NG = 2, NS = 5;
    static constexpr bool DEPOSIT = true;
    // 16 lane-interleaved copies of the five gathered fields: with one copy, 29 % of the shared-memory wavefronts of
    // this pass were bank conflicts of the gathers (ncu r2e; 32 lanes reading random cells of 36-double vectors)
    static constexpr int FIELD_COPIES = 16;
    static constexpr int FC = FIELD_COPIES;
    struct Params { double dtqm, wscale0, charge, mass, cw; };
    template <bool LP>
    static __device__ __forceinline__ void apply(Particle &p, const PassParams<OpLoopTail> &P, const double *sf, const Acc<LP> &acc)
    {
        const int nh = P.m.n + kHalo;
        const Pos ps = locate(p.x, P.m);
        int g0, g1;
        first_dofs<D0, D1>(ps, P.m, g0, g1);
        double b1[D1 + 1], b0[D0 + 1];
        basis_pp<D1>(ps.t, b1);
        basis_pp<D0>(ps.t, b0);
        const double ws = p.w * P.op.wscale0;
        deposit_h<D0, LP>(acc, g0, b0, ws * p.v2);                       // j_dofs[2] before the kick (OpHEJ2)
        p.v1 = fma(P.op.dtqm, gather_hs<D1, FC>(sf, g1, b1), p.v1);       // operatorHE kick with the snapshot fields
        p.v2 = fma(P.op.dtqm, gather_hs<D0, FC>(sf + (size_t)nh * FC, g0, b0), p.v2);
        deposit_h<D0, LP>(acc, nh + g0, b0, ws);                          // add_charge! (OpCharge)
        double wm = p.w * P.op.mass;                                      // OpDiag, operation for operation
        wm *= P.op.cw;
        acc.sum(0, (p.v1 * p.v1 + p.v2 * p.v2) * wm);
        acc.sum(1, p.v1 * wm);
        acc.sum(2, p.v2 * wm);
        const double e1 = gather_hs<D1, FC>(sf + (size_t)2 * nh * FC, g1, b1);
        const double e2 = gather_hs<D0, FC>(sf + (size_t)3 * nh * FC, g0, b0);
        const double bf = gather_hs<D1, FC>(sf + (size_t)4 * nh * FC, g1, b1);
        const double wq = P.op.charge * p.w * P.op.cw;
        acc.sum(3, (p.v1 * e1 + p.v2 * e2) * wq);
        acc.sum(4, wq * p.v1 * p.v2 * bf);
    }
};

// ---- 1d1v operatorHB = E-kick without q/m (hamiltonian_splitting_1d1v.jl:113-125). fields: [e1]
template <int D1>
struct OpHB11 {
    static constexpr int READ = ROW_X | ROW_V1, WRITE = ROW_V1;
    static constexpr int NF = 1, NG = 0, NS = 0;
    static constexpr bool DEPOSIT = false;
    struct Params { double dt; };
    template <bool LP>
    static __device__ __forceinline__ void apply(Particle &p, const PassParams<OpHB11> &P, const double *sf, const Acc<LP> &)
    {
        const Pos ps = locate(p.x, P.m);
        double b1[D1 + 1];
        basis_pp<D1>(ps.t, b1);
        p.v1 = fma(P.op.dt, gather_h<D1>(sf, first_dof<D1>(ps, P.m), b1), p.v1);
    }
};

// ---- 1d1v operatorHp1 (hamiltonian_splitting_1d1v.jl:70-94): j1 only, v unchanged;
//      new index by floor (pmc1d.jl:487)
template <int D1>
struct OpHp111 {
    static constexpr int READ = ROW_X | ROW_V1 | ROW_W, WRITE = ROW_X;
    static constexpr int NF = 0, NG = 1, NS = 0;
    static constexpr bool DEPOSIT = true;
    struct Params { double dt, wscale1_dx; };
    template <bool LP>
    static __device__ __forceinline__ void apply(Particle &p, const PassParams<OpHp111> &P, const double *, const Acc<LP> &acc)
    {
        const double x_new = fma(P.op.dt, p.v1, p.x);
        const Pos po = locate(p.x, P.m);
        Pos pn;
        cell_offset_floor(x_new, P.m, pn.c, pn.t);
        current_update_v<D1, LP, false>(acc, 0, nullptr, po, pn, p.w * P.op.wscale1_dx, 0.0, 0.0, P.m);
        p.x = mod_julia(x_new, P.m.Lx);
    }
};

// ---- Boris push_v_bpart! (hamiltonian_splitting_boris.jl:211-233). fields: [b_mid]
__device__ __forceinline__ void boris_rotate(Particle &p, double bfield, double qmdt)
{
    bfield = qmdt * bfield;
    double M11 = 1.0 / (1.0 + bfield * bfield);
    const double M12 = M11 * bfield * 2.0;
    M11 = M11 * (1 - bfield * bfield);
    const double v1 = M11 * p.v1 + M12 * p.v2;
    const double v2 = -M12 * p.v1 + M11 * p.v2;
    p.v1 = v1;
    p.v2 = v2;
}
template <int D1>
struct OpBorisB {
    static constexpr int READ = ROW_X | ROW_V1 | ROW_V2, WRITE = ROW_V1 | ROW_V2;
    static constexpr int NF = 1, NG = 0, NS = 0;
    static constexpr bool DEPOSIT = false;
    struct Params { double qmdt; };
    template <bool LP>
    static __device__ __forceinline__ void apply(Particle &p, const PassParams<OpBorisB> &P, const double *sf, const Acc<LP> &)
    {
        const Pos ps = locate(p.x, P.m);
        double b1[D1 + 1];
        basis_pp<D1>(ps.t, b1);
        boris_rotate(p, gather_h<D1>(sf, first_dof<D1>(ps, P.m), b1), P.op.qmdt);
    }
};

// ---- Boris push_x_accumulate_j! (hamiltonian_splitting_boris.jl:250-288):
//      deposits w*v1 (D1) -> grid 0 and w*v2 (D0) -> grid 1 at the un-wrapped midpoint
template <int D0, int D1, bool LP>
__device__ __forceinline__ void boris_push_x(Particle &p, double dt, double wscale0, double wscale1, const Acc<LP> &acc,
                                             const Mesh1D &m)
{
    const double x_new = fma(dt, p.v1, p.x);
    const Pos pm = locate((p.x + x_new) * 0.5, m);
    int g0, g1;
    first_dofs<D0, D1>(pm, m, g0, g1);
    double b1[D1 + 1], b0[D0 + 1];
    basis_pp<D1>(pm.t, b1);
    basis_pp<D0>(pm.t, b0);
    deposit_h<D1, LP>(acc, g1, b1, (p.w * wscale1) * p.v1);
    deposit_h<D0, LP>(acc, (m.n + kHalo) + g0, b0, (p.w * wscale0) * p.v2);
    p.x = mod_julia(x_new, m.Lx);
}
template <int D0, int D1>
struct OpBorisX {
    static constexpr int READ = ROW_X | ROW_V1 | ROW_V2 | ROW_W, WRITE = ROW_X;
    static constexpr int NF = 0, NG = 2, NS = 0;
    static constexpr bool DEPOSIT = true;
    struct Params { double dt, wscale0, wscale1; };
    template <bool LP>
    static __device__ __forceinline__ void apply(Particle &p, const PassParams<OpBorisX> &P, const double *, const Acc<LP> &acc)
    {
        boris_push_x<D0, D1, LP>(p, P.op.dt, P.op.wscale0, P.op.wscale1, acc, P.m);
    }
};

// ---- one whole Boris step per particle: epart(dt/2), bpart(dt), epart(dt/2), push_x(dt)
//      (hamiltonian_splitting_boris.jl:146-155; no field changes in between, so the four
//      reference loops collapse into one pass: 56 B/particle instead of 160).
//      fields: [e1_mid, e2_mid, b_mid]
// Laid out like the fused Strang pass (shared-memory pipe and latency bound): the three fields are staged in pp form
// (evaluate_pp, pmc1d.jl:242-250) with the kick factors folded in -- all three are gathered at the same point, so the
// B-spline bases are only needed for the deposits at the midpoint; 16 lane-interleaved copies; four particles per lane
// run their branch-free arithmetic before the lane-private read-modify-writes; one block of 8 warps per SM.
template <int D0, int D1>
struct BorisWork {
    double x, v1, v2;
    int g0, g1;
    bool slow;
    double d1[D1 + 1], d0[D0 + 1];
};
template <int D0, int D1>
struct OpBorisStep {
    static constexpr int READ = ROW_X | ROW_V1 | ROW_V2 | ROW_W, WRITE = ROW_X | ROW_V1 | ROW_V2;
    static constexpr int NC1 = D1 + 1, NC0 = D0 + 1;
    static constexpr int NF = 2 * NC1 + NC0, NG = 2, NS = 0;   // staged per cell: half_dtqm e1 (NC1), half_dtqm e2 (NC0), qmdt b (NC1)
    static constexpr bool DEPOSIT = true;
    static constexpr bool PAIRWISE = true;
    static constexpr int THREADS = 256;   // (10 warps fit -- 222 KB, 168 registers -- but run 11 % slower, r01r)
    static constexpr int FIELD_COPIES = 16;
    static constexpr int FIELD_HALO = 2;   // pp tables cover the cells -1 .. n
    static constexpr bool CUSTOM_STAGE = true;
    static constexpr int FC = FIELD_COPIES;
    static constexpr int OFF_E2 = NC1 * FC, OFF_B = (NC1 + NC0) * FC, CELL = NF * FC;
    struct Params { double dt, half_dtqm, qmdt, wscale0, wscale1; };
    using PP = PassParams<OpBorisStep>;

    static __device__ __forceinline__ void stage(const PP &P, double *sfield, int tid)
    {
        const int n = P.m.n;
        for (int i = tid; i < n + FIELD_HALO; i += THREADS) {
            double co[NF];
            cell_poly<D1>(i - 1, n, [&](int g) { return P.op.half_dtqm * P.fields[0][g]; }, co);
            cell_poly<D0>(i - 1, n, [&](int g) { return P.op.half_dtqm * P.fields[1][g]; }, co + NC1);
            cell_poly<D1>(i - 1, n, [&](int g) { return P.op.qmdt * P.fields[2][g]; }, co + NC1 + NC0);
#pragma unroll
            for (int q = 0; q < NF; ++q)
#pragma unroll
                for (int k = 0; k < FC; ++k) sfield[((size_t)i * NF + q) * FC + k] = co[q];
        }
    }

    // general per-particle form (any position / displacement): dof-form fields from global memory
    template <bool LP>
    static __device__ __forceinline__ void apply(Particle &p, const PP &P, const double *, const Acc<LP> &acc)
    {
        const int n = P.m.n;
        const Pos ps = locate(p.x, P.m);
        int g0, g1;
        first_dofs<D0, D1>(ps, P.m, g0, g1);
        double b1[D1 + 1], b0[D0 + 1];
        basis_pp<D1>(ps.t, b1);
        basis_pp<D0>(ps.t, b0);
        double e1 = 0.0, e2 = 0.0, bf = 0.0;
        int g = g1;
#pragma unroll
        for (int k = 0; k <= D1; ++k) {
            e1 = fma(P.fields[0][g], b1[k], e1);
            bf = fma(P.fields[2][g], b1[k], bf);
            g = wrap_next(g + 1, n);
        }
        g = g0;
#pragma unroll
        for (int k = 0; k <= D0; ++k) {
            e2 = fma(P.fields[1][g], b0[k], e2);
            g = wrap_next(g + 1, n);
        }
        p.v1 = fma(P.op.half_dtqm, e1, p.v1);
        p.v2 = fma(P.op.half_dtqm, e2, p.v2);
        boris_rotate(p, bf, P.op.qmdt);
        p.v1 = fma(P.op.half_dtqm, e1, p.v1);
        p.v2 = fma(P.op.half_dtqm, e2, p.v2);
        boris_push_x<D0, D1, LP>(p, P.op.dt, P.op.wscale0, P.op.wscale1, acc, P.m);
    }
    struct XV { double x, v1, v2; };
    template <bool LP>
    static __device__ __noinline__ XV apply_general(double x, double v1, double v2, double w, const PP &P, double *accp)
    {
        Particle p{x, v1, v2, w};
        Acc<LP> acc{accp};
        apply<LP>(p, P, nullptr, acc);
        return XV{p.x, p.v1, p.v2};
    }
    template <bool LP>
    static __device__ __forceinline__ void general(Particle &a, const PP &P, const Acc<LP> &acc)
    {
        const XV r = apply_general<LP>(a.x, a.v1, a.v2, a.w, P, acc.p);
        a.x = r.x; a.v1 = r.v1; a.v2 = r.v2;
    }

    // branch-free arithmetic of one particle; no shared-memory writes
    static __device__ __forceinline__ BorisWork<D0, D1> work(const Particle &p, const PP &P, const double *sf)
    {
        BorisWork<D0, D1> W;
        const int n = P.m.n;
        const Pos po = locate(p.x, P.m);
        const double *co = sf + (size_t)min((unsigned)(po.c + 1), (unsigned)(n + 1)) * CELL;
        const double e1 = horner_s<D1, FC>(co, po.t);            // half_dtqm * E1(x)
        const double e2 = horner_s<D0, FC>(co + OFF_E2, po.t);   // half_dtqm * E2(x)
        const double beta = horner_s<D1, FC>(co + OFF_B, po.t);  // q/m dt/2 * B(x)
        double v1 = p.v1 + e1, v2 = p.v2 + e2;
        {   // push_v_bpart! :211-233
            double M11 = 1.0 / fma(beta, beta, 1.0);
            const double M12 = (M11 * beta) * 2.0;
            M11 = M11 * fma(-beta, beta, 1.0);
            const double r1 = fma(M12, v2, M11 * v1);
            const double r2 = fma(M11, v2, -(M12 * v1));
            v1 = r1 + e1;
            v2 = r2 + e2;
        }
        // push_x_accumulate_j! :250-288: deposits at the un-wrapped midpoint
        const double x_new = fma(P.op.dt, v1, p.x);
        const Pos pm = locate((p.x + x_new) * 0.5, P.m);
        const int g0 = wrap_near(pm.c - D0, n);
        const int g1 = (D1 == D0) ? g0 : wrap_next(g0 + (D0 - D1), n);
        double b1[D1 + 1], b0[D0 + 1];
        basis_pp<D1>(pm.t, b1);
        basis_pp<D0>(pm.t, b0);
        const double w1 = (p.w * P.op.wscale1) * v1, w0 = (p.w * P.op.wscale0) * v2;
#pragma unroll
        for (int k = 0; k <= D1; ++k) W.d1[k] = w1 * b1[k];
#pragma unroll
        for (int k = 0; k <= D0; ++k) W.d0[k] = w0 * b0[k];
        W.g0 = g0;
        W.g1 = g1;
        const double L = P.m.Lx;
        W.x = x_new < 0.0 ? x_new + L : (x_new >= L ? x_new - L : x_new);
        W.v1 = v1;
        W.v2 = v2;
        const unsigned span = (unsigned)(n + 2);
        W.slow = (unsigned)(po.c + 1) >= span || (unsigned)(pm.c + 1) >= span || !(x_new >= -L && x_new < 2.0 * L) || n < 8;
        return W;
    }
    static __device__ __forceinline__ void commit(const BorisWork<D0, D1> &W, double *acc, int nh)
    {
        double *q1 = acc + (size_t)W.g1 * 32, *q0 = acc + (size_t)(nh + W.g0) * 32;
        double r1[D1 + 1], r0[D0 + 1];
#pragma unroll
        for (int k = 0; k <= D1; ++k) r1[k] = q1[k * 32];
#pragma unroll
        for (int k = 0; k <= D0; ++k) r0[k] = q0[k * 32];
#pragma unroll
        for (int k = 0; k <= D1; ++k) q1[k * 32] = r1[k] + W.d1[k];
#pragma unroll
        for (int k = 0; k <= D0; ++k) q0[k * 32] = r0[k] + W.d0[k];
    }
    template <bool LP>
    static __device__ __forceinline__ void apply_pair(Particle &a, Particle &b, const PP &P, const double *sf, const Acc<LP> &acc)
    {
        const BorisWork<D0, D1> Wa = work(a, P, sf), Wb = work(b, P, sf);
        if (__builtin_expect(Wa.slow | Wb.slow, 0)) {
            general<LP>(a, P, acc);
            general<LP>(b, P, acc);
            return;
        }
        const int nh = P.m.n + kHalo;
        commit(Wa, acc.p, nh);
        commit(Wb, acc.p, nh);
        a.x = Wa.x; a.v1 = Wa.v1; a.v2 = Wa.v2;
        b.x = Wb.x; b.v1 = Wb.v1; b.v2 = Wb.v2;
    }
    template <bool LP>
    static __device__ __forceinline__ void apply_quad(Particle &a, Particle &b, Particle &c, Particle &d, const PP &P,
                                                      const double *sf, const Acc<LP> &acc)
    {
        const BorisWork<D0, D1> Wa = work(a, P, sf), Wb = work(b, P, sf), Wc = work(c, P, sf), Wd = work(d, P, sf);
        if (__builtin_expect(Wa.slow | Wb.slow | Wc.slow | Wd.slow, 0)) {
            general<LP>(a, P, acc);
            general<LP>(b, P, acc);
            general<LP>(c, P, acc);
            general<LP>(d, P, acc);
            return;
        }
        const int nh = P.m.n + kHalo;
        commit(Wa, acc.p, nh);
        commit(Wb, acc.p, nh);
        commit(Wc, acc.p, nh);
        commit(Wd, acc.p, nh);
        a.x = Wa.x; a.v1 = Wa.v1; a.v2 = Wa.v2;
        b.x = Wb.x; b.v1 = Wb.v1; b.v2 = Wb.v2;
        c.x = Wc.x; c.v1 = Wc.v1; c.v2 = Wc.v2;
        d.x = Wd.x; d.v1 = Wd.v1; d.v2 = Wd.v2;
    }
};

// ---- add_charge! over a ParticleGroup with marker charge get_charge (diagnostics.jl:24-28)
template <int D>
struct OpCharge {
    static constexpr int READ = ROW_X | ROW_W, WRITE = 0;
    static constexpr int NF = 0, NG = 1, NS = 0;
    static constexpr bool DEPOSIT = true;
    struct Params { double wscale; };   // charge * common_weight * scaling
    template <bool LP>
    static __device__ __forceinline__ void apply(Particle &p, const PassParams<OpCharge> &P, const double *, const Acc<LP> &acc)
    {
        const Pos ps = locate(p.x, P.m);
        double b[D + 1];
        basis_pp<D>(ps.t, b);
        deposit_h<D, LP>(acc, first_dof<D>(ps, P.m), b, p.w * P.op.wscale);
    }
};

// ---- write_step! particle sums (diagnostics.jl:45-92,197-211): scalars [KE, P1, P2, transfer, vvb]
//      fields: [e1, e2, b]
template <int D0, int D1>
struct OpDiag {
    static constexpr int READ = ROW_X | ROW_V1 | ROW_V2 | ROW_W, WRITE = 0;
    static constexpr int NF = 3, NG = 0, NS = 5;
    static constexpr bool DEPOSIT = true;
    struct Params { double charge, mass, cw; };
    template <bool LP>
    static __device__ __forceinline__ void apply(Particle &p, const PassParams<OpDiag> &P, const double *sf, const Acc<LP> &acc)
    {
        const int nh = P.m.n + kHalo;
        double wm = p.w * P.op.mass;
        wm *= P.op.cw;
        acc.sum(0, (p.v1 * p.v1 + p.v2 * p.v2) * wm);
        acc.sum(1, p.v1 * wm);
        acc.sum(2, p.v2 * wm);
        const Pos ps = locate(p.x, P.m);
        int g0, g1;
        first_dofs<D0, D1>(ps, P.m, g0, g1);
        double b1[D1 + 1], b0[D0 + 1];
        basis_pp<D1>(ps.t, b1);
        basis_pp<D0>(ps.t, b0);
        const double e1 = gather_h<D1>(sf, g1, b1);
        const double e2 = gather_h<D0>(sf + nh, g0, b0);
        const double bf = gather_h<D1>(sf + 2 * nh, g1, b1);
        const double wq = P.op.charge * p.w * P.op.cw;  // get_charge
        acc.sum(3, (p.v1 * e1 + p.v2 * e2) * wq);
        acc.sum(4, wq * p.v1 * p.v2 * bf);
    }
};

}  // namespace gempic
