// splines.cuh -- device-side uniform B-spline arithmetic shared by all particle kernels.
//
// Follows the arithmetic (operation order) of the reference so that per-particle results
// agree to the last few ulps (the only systematic difference is FMA contraction):
//   uniform_bsplines_eval_basis!          src/low_level_bsplines.jl:63-80
//   index / offset split (trunc)          src/particle_mesh_coupling_1d.jl:267-270, 306-314, 439-442
//   periodic dof index mod1(c-d+i, n)     src/particle_mesh_coupling_1d.jl:276-279, 416-418
#pragma once
#include <cuda_runtime.h>

namespace gempic {

constexpr int kMaxDegree = 3;  // Maxwell1DFEM supports degree 1..3 (src/maxwell_1d_fem.jl:60-92)

// de Boor recurrence for the D+1 non-zero uniform B-splines at `offset`; fully unrolled so
// that b[] lives in registers.
template <int D>
__device__ __forceinline__ void bspline_basis(double offset, double (&b)[D + 1])
{
    b[0] = 1.0;
#pragma unroll
    for (int j = 1; j <= D; ++j) {
        double xx = -offset;
        const double j_real = (double)j;
        const double inv_j = 1.0 / j_real;
        double saved = 0.0;
#pragma unroll
        for (int r = 0; r < j; ++r) {
            xx = xx + 1.0;
            const double temp = b[r] * inv_j;
            b[r] = saved + xx * temp;
            saved = (j_real - xx) * temp;
        }
        b[j] = saved;
    }
}

// ---- pp form (src/splinepp.jl) ----------------------------------------------------------
// The particle kernels evaluate the D+1 spline pieces as polynomials in Horner form with the
// coefficient columns of SplinePP.poly_coeffs (splinepp.jl:39-69) -- the arithmetic of the
// reference's evaluate_pp / add_charge_pp! (pmc1d.jl:106-122, 242-250) -- instead of the
// de Boor recurrence of evaluate / add_charge! (low_level_bsplines.jl:63-80).  Both are
// held to the same 1e-15 goldens by the reference's tests
// (test_particle_mesh_coupling_spline_1d.jl:57-65, 137-144); the pp form needs less than
// half the fp64 instructions.
template <int D>
__device__ __forceinline__ void basis_pp(double t, double (&b)[D + 1])
{
    constexpr double i2 = 0.5, i6 = 1.0 / 6.0;
    if constexpr (D == 0) {
        b[0] = 1.0;
    } else if constexpr (D == 1) {
        b[0] = fma(-1.0, t, 1.0);
        b[1] = t;
    } else if constexpr (D == 2) {
        b[0] = fma(fma(i2, t, -1.0), t, i2);
        b[1] = fma(fma(-1.0, t, 1.0), t, i2);
        b[2] = (i2 * t) * t;
    } else {
        b[0] = fma(fma(fma(-i6, t, i2), t, -i2), t, i6);
        b[1] = fma((fma(i2, t, -1.0) * t), t, 4.0 * i6);
        b[2] = fma(fma(fma(-i2, t, i2), t, i2), t, i6);
        b[3] = ((i6 * t) * t) * t;
    }
}

// primitives P_k(t) = int_0^t N_k: columns of SplinePP.poly_coeffs_fp (splinepp.jl:41,47-49,
// 71-88) evaluated like horner_primitive_1d (:281-285).  The line integral of a spline over
// an in-cell segment [lo, up] is dx (P_k(up) - P_k(lo)) -- update_jv_pp! (pmc1d.jl:198-230),
// equal to the Gauss-Legendre sum of update_jv! (:385-425) to rounding
// (test_particle_mesh_coupling_spline_1d.jl:82-124 pins both to 1e-15).
template <int D>
__device__ __forceinline__ void prim_pp(double t, double (&P)[D + 1])
{
    constexpr double i2 = 0.5, i3 = 1.0 / 3.0, i4 = 0.25, i6 = 1.0 / 6.0, i8 = 0.125, i24 = 1.0 / 24.0;
    if constexpr (D == 0) {
        P[0] = t;
    } else if constexpr (D == 1) {
        P[0] = fma(-i2, t, 1.0) * t;
        P[1] = (i2 * t) * t;
    } else if constexpr (D == 2) {
        P[0] = fma(fma(i6, t, -i2), t, i2) * t;
        P[1] = fma(fma(-i3, t, i2), t, i2) * t;
        P[2] = ((i6 * t) * t) * t;
    } else {
        P[0] = fma(fma(fma(-i24, t, i6), t, -i4), t, i6) * t;
        P[1] = fma((fma(i8, t, -i3) * t), t, 4.0 * i6) * t;
        P[2] = fma(fma(fma(-i8, t, i6), t, i4), t, i6) * t;
        P[3] = (((i24 * t) * t) * t) * t;
    }
}
// Degree-D pieces from the degree-(D-1) primitives at the same offset: d/dt N^D_k = N^{D-1}_{k-1} - N^{D-1}_k
// (pieces numbered as in basis_pp / prim_pp), so  N^D_k(t) = N^D_k(0) + P_{k-1}(t) - P_k(t).
// A pass that needs both prim_pp<D-1>(t) and basis_pp<D>(t) gets the latter for D DADDs.
template <int D>
__device__ __forceinline__ void basis_from_prim(const double (&P)[D], double (&b)[D + 1])
{
    static_assert(D >= 1, "degree");
    // N^D_k(0): D=1 (1,0)  D=2 (1/2,1/2,0)  D=3 (1/6,4/6,1/6,0)
    constexpr double at0[4][4] = {{1, 0, 0, 0}, {1, 0, 0, 0}, {0.5, 0.5, 0, 0}, {1.0 / 6.0, 4.0 / 6.0, 1.0 / 6.0, 0}};
    b[0] = at0[D][0] - P[0];
#pragma unroll
    for (int k = 1; k < D; ++k) b[k] = (at0[D][k] + P[k - 1]) - P[k];
    b[D] = P[D - 1];
}

// P_k(1): integral of piece k over its whole cell
template <int D>
__device__ __forceinline__ double prim_full(int k)
{
    if (D == 0) return 1.0;
    if (D == 1) return 0.5;
    if (D == 2) return (k == 1) ? 2.0 / 3.0 : 1.0 / 6.0;
    return (k == 1 || k == 2) ? 11.0 / 24.0 : 1.0 / 24.0;
}

// 1D mesh as the kernels see it
struct Mesh1D {
    double xmin;
    double dx;
    double inv_dx;  // RN(1/dx), used by div_dx
    double Lx;  // domain length used by x = mod(x_new, Lx)  (hamiltonian_splitting_1d2v.jl:79)
    int n;
    int pow2;   // n is a power of two -> mask instead of %
};

// floored modulo into [0, n).  Particles sit within one period of the domain, so g is in
// [-n, 2n) and one conditional add/subtract suffices; anything further out (|dt v| > L)
// takes the out-of-line integer division.
static __device__ __noinline__ int wrap_index_far(int g, int n)
{
    g %= n;
    return g < 0 ? g + n : g;
}
__device__ __forceinline__ int wrap_index(int g, const Mesh1D &m)
{
    const int n = m.n;
    if (__builtin_expect((unsigned)(g + n) >= (unsigned)(3 * n), 0)) return wrap_index_far(g, n);
    g = g < 0 ? g + n : g;
    return g >= n ? g - n : g;
}
// g is known to be in [0, 2n): one conditional subtraction
__device__ __forceinline__ int wrap_next(int g, int n) { return g >= n ? g - n : g; }

// a / dx without the ~20-instruction fp64 division sequence: one Newton correction of
// a * RN(1/dx) with an exact FMA residual (Markstein).  The result is the correctly rounded
// quotient for all but a vanishing set of operands (where it is off by one ulp), i.e. the
// value the reference's `(x - xmin) / delta_x` produces (pmc1d.jl:267, 306, 312, 439).
__device__ __forceinline__ double div_dx(double a, const Mesh1D &m)
{
    const double q = a * m.inv_dx;
    const double r = fma(-q, m.dx, a);
    return fma(r, m.inv_dx, q);
}

// cell = trunc(xi), offset = xi - cell    (NOT floor: SURVEY appendix A.2 Q1)
__device__ __forceinline__ void cell_offset(double x, const Mesh1D &m, int &cell, double &offset)
{
    const double xi = div_dx(x - m.xmin, m);
    cell = __double2int_rz(xi);
    offset = xi - (double)cell;
}
// cell = floor(xi): new index of the 1d1v add_current_update_v! (pmc1d.jl:487)
__device__ __forceinline__ void cell_offset_floor(double x, const Mesh1D &m, int &cell, double &offset)
{
    const double xi = div_dx(x - m.xmin, m);
    cell = __double2int_rd(xi);
    offset = xi - (double)cell;
}

// Julia's mod(x, L) for L > 0 (rem, then shift; a zero result is +0.0)
__device__ __forceinline__ double mod_julia(double x, double L)
{
    if (x >= 0.0 && x < L) return x;
    double r = fmod(x, L);
    if (r < 0.0) r += L;
    else if (r == 0.0) r = 0.0;
    return r;
}

}  // namespace gempic
