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

// 1D mesh as the kernels see it
struct Mesh1D {
    double xmin;
    double dx;
    double Lx;  // domain length used by x = mod(x_new, Lx)  (hamiltonian_splitting_1d2v.jl:79)
    int n;
    int pow2;   // n is a power of two -> mask instead of %
};

// floored modulo into [0, n)
__device__ __forceinline__ int wrap_index(int g, const Mesh1D &m)
{
    if (m.pow2) return g & (m.n - 1);
    g %= m.n;
    return g < 0 ? g + m.n : g;
}
// g is known to be in [0, 2n): one conditional subtraction
__device__ __forceinline__ int wrap_next(int g, int n) { return g >= n ? g - n : g; }

// cell = trunc(xi), offset = xi - cell    (NOT floor: SURVEY appendix A.2 Q1)
__device__ __forceinline__ void cell_offset(double x, const Mesh1D &m, int &cell, double &offset)
{
    const double xi = (x - m.xmin) / m.dx;
    cell = __double2int_rz(xi);
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

// sum_k field[(cell-D+k) mod n] * b[k], accumulated from 0.0 in k order (evaluate, :446-450)
template <int D>
__device__ __forceinline__ double gather(const double *__restrict__ field, int cell, const double (&b)[D + 1],
                                         const Mesh1D &m)
{
    int g = wrap_index(cell - D, m);
    double v = 0.0;
#pragma unroll
    for (int k = 0; k <= D; ++k) {
        v += field[g] * b[k];
        g = wrap_next(g + 1, m.n);
    }
    return v;
}

// Gauss-Legendre nodes on [-1,1] for n_quad = (degree+2)/2 (src/particle_mesh_coupling_1d.jl:67,72)
template <int D>
struct Quad {
    static constexpr int n = (D + 2) / 2;
};
__device__ __forceinline__ double quad_x(int nq, int q)
{
    // n=1: {0}; n=2: {-1/sqrt(3), +1/sqrt(3)}
    return nq == 1 ? 0.0 : (q == 0 ? -0.57735026918962576451 : 0.57735026918962576451);
}
__device__ __forceinline__ double quad_w(int nq, int /*q*/) { return nq == 1 ? 2.0 : 1.0; }

// Line-integral weights of one in-cell segment [lower, upper] (cell units):
//   s[k] = sign * dx * sum_q w_q c1 N_k(c1 x_q + c2)      (update_jv!, :399-414)
template <int D>
__device__ __forceinline__ void segment_weights(double lower, double upper, double sign_dx, double (&s)[D + 1])
{
    constexpr int NQ = Quad<D>::n;
    const double c1 = 0.5 * (upper - lower);
    const double c2 = 0.5 * (upper + lower);
    bspline_basis<D>(c1 * quad_x(NQ, 0) + c2, s);
    const double f = quad_w(NQ, 0) * c1;
#pragma unroll
    for (int k = 0; k <= D; ++k) s[k] *= f;
    if (NQ > 1) {
        double more[D + 1];
        bspline_basis<D>(c1 * quad_x(NQ, 1) + c2, more);
#pragma unroll
        for (int k = 0; k <= D; ++k) s[k] += more[k] * quad_w(NQ, 1) * c1;
    }
#pragma unroll
    for (int k = 0; k <= D; ++k) s[k] *= sign_dx;
}

}  // namespace gempic
