// hs2d.cu -- HamiltonianSplitting{2,3}: the 2d3v Vlasov-Maxwell splitting on ParticleGroup{2,3} and
// TwoDMaxwell (BASELINE config 5).
//
// GEMPIC.jl ships no 2d3v integrator (src/hamiltonian_splitting_2d3v.jl is an empty file); these
// operators are the 2D extension of src/hamiltonian_splitting_1d2v.jl:41-236 on the reference's 2D
// building blocks (tensor-product splines of particle_mesh_coupling_2d.jl, the mixed-degree spaces of
// maxwell_2d_fem.jl), following the GEMPIC paper cited in CITATION.bib:
//   HE   v += dt q/m E(x);  b -= dt curl e                      (1d2v.jl:191-219)
//   HB   e += dt M1^-1 curl^T M2 b                              (:234-236)
//   Hp1  x1 += dt v1;  j1 = line integral;  v2 -= q/m int B3 dx1;  v3 += q/m int B2 dx1   (:41-112)
//   Hp2  x2 += dt v2;  j2 = line integral;  v1 += q/m int B3 dx2;  v3 -= q/m int B1 dx2
//   Hp3  v1 -= dt q/m v3 B2;  v2 += dt q/m v3 B1;  j3 = w v3 N(x) dt                      (:129-176)
//   strang_splitting!: HB HE Hp3 Hp2 Hp1 Hp2 Hp3 HE HB  (hamiltonian_splitting.jl:98-108 with the third
//   velocity component added symmetrically)
// Spaces (p = degree): E1 S^{p-1}xS^p, E2 S^pxS^{p-1}, E3 S^pxS^p; B1 S^pxS^{p-1}, B2 S^{p-1}xS^p,
// B3 S^{p-1}xS^{p-1} (test/test_maxwell_2d_fem.jl:54-56,88-90).  Cells are floor((x-xmin)/dx).
//
// Kernel design (k2_pass).  A 64x64 grid is 32 KB per component: it cannot be privatised per lane like
// the 1D grids.  Particles are kept cell-sorted (pg_sort_2d) and every warp streams a contiguous chunk
// of them, so the dofs it touches lie in a small window around the cell of the chunk's first particle:
//   * deposits go to a LANE-PRIVATE window tile in shared memory ((2R+1+p)^2 dofs, R = 2 cells of drift
//     tolerance; slot s of lane l at (s*32 + l)*8 B): plain LDS/DADD/STS, no atomics, conflict free.  At
//     the end of the chunk the 32 copies are summed and added to the global grid with one fp64 RED per
//     dof.  Particles outside the window (rare while the sort is fresh) fall back to global REDs;
//   * the field dofs of the same window are staged once per chunk in a small warp-shared tile, so a gather
//     is a run of LDS with compile-time offsets from one base (the lanes of a warp share (nearly) the
//     same stencil: broadcast, conflict free); outside the window a particle reads global memory;
//   * the next two particles of every lane are loaded before the current two are processed.
#include <algorithm>
#include <cstdlib>

#include "hostutil.hpp"
#include "objects.cuh"
#include "tma.cuh"

namespace gempic {

#ifndef K2_PASS_MINB
#define K2_PASS_MINB 3   // resident blocks per SM of the tile passes: their 70 KB of tiles per block allow no more
#endif
constexpr int kR2 = 2;          // window radius (cells) around the chunk's base cell
constexpr int kWarps2 = 4;      // warps per block
constexpr int kThreads2 = kWarps2 * 32;

struct Mesh2 {
    double xmin[2], d[2], inv_d[2], L[2], xmax[2];
    int n[2];
};

struct Rows2 {
    double *__restrict__ x[2];
    double *__restrict__ v[3];
    double *__restrict__ w;
};

struct Part2 {
    double x[2], v[3], w;
};

__device__ __forceinline__ int wrapi(int g, int n)
{
    if (__builtin_expect((unsigned)(g + n) >= (unsigned)(3 * n), 0)) {
        g %= n;
        return g < 0 ? g + n : g;
    }
    g = g < 0 ? g + n : g;
    return g >= n ? g - n : g;
}

template <int AX>
__device__ __forceinline__ void locate2(double x, const Mesh2 &m, int &c, double &t)
{
    const double a = x - m.xmin[AX];
    const double q = a * m.inv_d[AX];
    const double r = fma(-q, m.d[AX], a);
    const double xi = fma(r, m.inv_d[AX], q);   // correctly rounded a / d (see div_dx, splines.cuh)
    c = __double2int_rd(xi);
    t = xi - (double)c;
}

// relative cell offset of c (unwrapped, within one period of the box) to the wrapped base cell
__device__ __forceinline__ int rel_cell(int c, int base, int n)
{
    int r = c - base;
    const int h = n >> 1;
    if (r > h) r -= n;
    else if (r < -h) r += n;
    return r;
}

// periodic dof indices of the degree-D0 stencil: idx[a] = (c - D0 + a) mod n.  The degree-(D0-1) stencil
// of the same cell is idx[1..D0].
template <int D0>
__device__ __forceinline__ void stencil(int c, int n, int (&idx)[D0 + 1])
{
    int g = wrapi(c - D0, n);
#pragma unroll
    for (int a = 0; a <= D0; ++a) {
        idx[a] = g;
        g = g + 1 == n ? 0 : g + 1;
    }
}

// sum_{a<=DX, b<=DY} f[ix[OX+a] + iy[OY+b]*nx] bx[a] by[b]
template <int DX, int DY, int OX, int OY, int D0>
__device__ __forceinline__ double eval2(const double *__restrict__ f, int nx, const int (&ix)[D0 + 1], const int (&iy)[D0 + 1],
                                        const double (&bx)[DX + 1], const double (&by)[DY + 1])
{
    double v = 0.0;
#pragma unroll
    for (int b = 0; b <= DY; ++b) {
        const double *row = f + (size_t)iy[OY + b] * nx;
        double s = 0.0;
#pragma unroll
        for (int a = 0; a <= DX; ++a) s = fma(__ldg(row + ix[OX + a]), bx[a], s);
        v = fma(s, by[b], v);
    }
    return v;
}

template <class Op>
struct P2 {
    Rows2 r;
    int64_t n;
    int64_t chunk;
    Mesh2 m;
    const double *f[8];
    double *grid;   // deposit target (nx*ny), zeroed by the launcher
    Rows2 dst;      // Op::SCATTER: the rows of the other particle buffer
    int *cursor;    // Op::SCATTER: next free slot of every cell (exclusive scan of the histogram of the cells after the push)
    typename Op::Params op;
};

// Cell-sort key of a position, and the position an operatorHp2 push of dt leaves a particle at.  The histogram kernel
// and the scattering pass must agree on these bit for bit (a particle counted in one cell and placed in another would
// overrun a neighbour's range), so both call exactly these functions.
__device__ __forceinline__ int sort_key2(double x, double y, const Mesh2 &m)
{
    int cx = __double2int_rd((x - m.xmin[0]) * m.inv_d[0]), cy = __double2int_rd((y - m.xmin[1]) * m.inv_d[1]);
    cx = min(max(cx, 0), m.n[0] - 1);
    cy = min(max(cy, 0), m.n[1] - 1);
    return cx + cy * m.n[0];
}
template <int DIR>
__device__ __forceinline__ double pushed(double x_old, double v, double dt, const Mesh2 &m)
{
    double xw = fma(dt, v, x_old);
    while (xw < m.xmin[DIR]) xw += m.L[DIR];
    while (xw >= m.xmax[DIR]) xw -= m.L[DIR];
    return xw;
}

// Window of a warp: cells base-R .. base+R in both directions; dofs base-R-D0 .. base+R, i.e. W x W with
//   local dof index = (cell - base) + R + (D0 - degree) + k,  k = 0..degree.
template <int D0>
struct Tile {
    static constexpr int W = 2 * kR2 + 1 + D0;
    static constexpr int SLOTS = W * W;
    // Row stride of the warp-shared FIELD tiles.  After some drift the lanes of a half-warp sit in cells that differ by
    // up to two in either direction, so two lanes read words (dy * WS + dx) doubles apart; with WS = W = 8 the common
    // pair dy = +-2, dx = 0 falls into the same bank (16 doubles = 32 banks): a 2-way conflict on almost every gather
    // (ncu r01h: 35 % of the shared wavefronts of Op2HE).  WS is the smallest stride >= W for which no offset with
    // |dx|, |dy| <= 2 is a multiple of 16 doubles.
    static constexpr int field_stride()
    {
        for (int S = W;; ++S) {
            bool ok = true;
            for (int dy = 1; dy <= 2; ++dy)
                for (int dx = -2; dx <= 2; ++dx)
                    if ((dy * S + dx) % 16 == 0) ok = false;
            if (ok) return S;
        }
    }
    static constexpr int WS = field_stride();
    static constexpr int FS = WS * W;   // doubles of one field tile
};

// sum_{b<=DY} (sum_{a<=DX} q[b*W + a] bx[a]) by[b] on a shared field tile (q = first dof of the stencil)
template <int DX, int DY, int W>
__device__ __forceinline__ double eval_tile(const double *__restrict__ q, const double (&bx)[DX + 1], const double (&by)[DY + 1])
{
    double v = 0.0;
#pragma unroll
    for (int b = 0; b <= DY; ++b) {
        double s = q[b * W] * bx[0];
#pragma unroll
        for (int a = 1; a <= DX; ++a) s = fma(q[b * W + a], bx[a], s);
        v = fma(s, by[b], v);
    }
    return v;
}

struct V3 { double v0, v1, v2; };

// ---- operatorHE: v += dt q/m E(x).  fields: E1 (D1,D0), E2 (D0,D1), E3 (D0,D0) ------------------------
template <int D0>
struct Op2HE {
    static __device__ __forceinline__ double stage(const P2<Op2HE> &P, int f, size_t g) { return P.f[f][g]; }
    static constexpr bool DEPOSIT = false, WRITE_X = false, WRITE_V = true, SCATTER = false;
    static constexpr int D = D0, NF = 3;
    struct Params { double dtqm; };
    using PP = P2<Op2HE>;

    static __device__ __noinline__ V3 slow(double x, double y, const PP &P)
    {
        constexpr int D1 = D0 - 1;
        int cx, cy, ix[D0 + 1], iy[D0 + 1];
        double tx, ty, bx0[D0 + 1], bx1[D1 + 1], by0[D0 + 1], by1[D1 + 1];
        locate2<0>(x, P.m, cx, tx);
        locate2<1>(y, P.m, cy, ty);
        stencil<D0>(cx, P.m.n[0], ix);
        stencil<D0>(cy, P.m.n[1], iy);
        basis_pp<D0>(tx, bx0); basis_pp<D1>(tx, bx1);
        basis_pp<D0>(ty, by0); basis_pp<D1>(ty, by1);
        const int nx = P.m.n[0];
        return V3{eval2<D1, D0, 1, 0, D0>(P.f[0], nx, ix, iy, bx1, by0), eval2<D0, D1, 0, 1, D0>(P.f[1], nx, ix, iy, bx0, by1),
                  eval2<D0, D0, 0, 0, D0>(P.f[2], nx, ix, iy, bx0, by0)};
    }

    static __device__ __forceinline__ void apply(Part2 &p, const PP &P, const double *ft, double *, int bx, int by)
    {
        constexpr int D1 = D0 - 1, WS = Tile<D0>::WS, FS = Tile<D0>::FS;
        int cx, cy;
        double tx, ty;
        locate2<0>(p.x[0], P.m, cx, tx);
        locate2<1>(p.x[1], P.m, cy, ty);
        const int rx = rel_cell(cx, bx, P.m.n[0]), ry = rel_cell(cy, by, P.m.n[1]);
        V3 e;
        if (__builtin_expect(rx >= -kR2 && rx <= kR2 && ry >= -kR2 && ry <= kR2, 1)) {
            double bx0[D0 + 1], bx1[D1 + 1], by0[D0 + 1], by1[D1 + 1];
            basis_pp<D0>(tx, bx0); basis_pp<D1>(tx, bx1);
            basis_pp<D0>(ty, by0); basis_pp<D1>(ty, by1);
            const double *q = ft + (ry + kR2) * WS + rx + kR2;
            e.v0 = eval_tile<D1, D0, WS>(q + 1, bx1, by0);
            e.v1 = eval_tile<D0, D1, WS>(q + FS + WS, bx0, by1);
            e.v2 = eval_tile<D0, D0, WS>(q + 2 * FS, bx0, by0);
        } else {
            e = slow(p.x[0], p.x[1], P);
        }
        p.v[0] = fma(P.op.dtqm, e.v0, p.v[0]);
        p.v[1] = fma(P.op.dtqm, e.v1, p.v[1]);
        p.v[2] = fma(P.op.dtqm, e.v2, p.v[2]);
    }
};

// (D0+1)^2 deposit of wv * wx (x) wy into the lane-private tile at local dof (lx, ly)
template <int D0>
__device__ __forceinline__ void deposit_tile(double *tile, int lx, int ly, const double (&wx)[D0 + 1], const double (&wy)[D0 + 1], double wv)
{
    constexpr int W = Tile<D0>::W;
    double *q = tile + (size_t)(ly * W + lx) * 32;
#pragma unroll
    for (int b = 0; b <= D0; ++b) {
        const double wb = wv * wy[b];
        double r[D0 + 1];
#pragma unroll
        for (int a = 0; a <= D0; ++a) r[a] = q[(b * W + a) * 32];
#pragma unroll
        for (int a = 0; a <= D0; ++a) q[(b * W + a) * 32] = fma(wb, wx[a], r[a]);
    }
}
// the same deposit straight to the global grid (particles outside the warp's window)
template <int D0>
__device__ __noinline__ void deposit_global(double *__restrict__ grid, int nx, int ny, int cx, int cy, double tx, double ty, double wv)
{
    int ix[D0 + 1], iy[D0 + 1];
    double wx[D0 + 1], wy[D0 + 1];
    stencil<D0>(cx, nx, ix);
    stencil<D0>(cy, ny, iy);
    basis_pp<D0>(tx, wx);
    basis_pp<D0>(ty, wy);
    for (int b = 0; b <= D0; ++b)
        for (int a = 0; a <= D0; ++a) atomicAdd(grid + ix[a] + (size_t)iy[b] * nx, (wv * wy[b]) * wx[a]);
}

// ---- add_charge!: rho += q w N^p(x1) N^p(x2) ----------------------------------------------------
template <int D0>
struct Op2Charge {
    static __device__ __forceinline__ double stage(const P2<Op2Charge> &P, int f, size_t g) { return P.f[f][g]; }
    static constexpr bool DEPOSIT = true, WRITE_X = false, WRITE_V = false, SCATTER = false;
    static constexpr int D = D0, NF = 0;
    struct Params { double wscale; };
    static __device__ __forceinline__ void apply(Part2 &p, const P2<Op2Charge> &P, const double *, double *tile, int bx, int by)
    {
        int cx, cy;
        double tx, ty;
        locate2<0>(p.x[0], P.m, cx, tx);
        locate2<1>(p.x[1], P.m, cy, ty);
        const int rx = rel_cell(cx, bx, P.m.n[0]), ry = rel_cell(cy, by, P.m.n[1]);
        const double wv = p.w * P.op.wscale;
        if (__builtin_expect(rx >= -kR2 && rx <= kR2 && ry >= -kR2 && ry <= kR2, 1)) {
            double bx0[D0 + 1], by0[D0 + 1];
            basis_pp<D0>(tx, bx0);
            basis_pp<D0>(ty, by0);
            deposit_tile<D0>(tile, rx + kR2, ry + kR2, bx0, by0, wv);
        } else {
            deposit_global<D0>(P.grid, P.m.n[0], P.m.n[1], cx, cy, tx, ty, wv);
        }
    }
};

// ---- operatorHp3: v1 -= dt q/m v3 B2, v2 += dt q/m v3 B1, j3 += q w v3 dt N N.  fields: B1 (D0,D1), B2 (D1,D0)
template <int D0>
struct Op2Hp3 {
    static __device__ __forceinline__ double stage(const P2<Op2Hp3> &P, int f, size_t g) { return P.f[f][g]; }
    static constexpr bool DEPOSIT = true, WRITE_X = false, WRITE_V = true, SCATTER = false;
    static constexpr int D = D0, NF = 2;
    struct Params { double dtqm, wscale_dt; };
    using PP = P2<Op2Hp3>;

    static __device__ __noinline__ V3 slow(double x, double y, const PP &P)
    {
        constexpr int D1 = D0 - 1;
        int cx, cy, ix[D0 + 1], iy[D0 + 1];
        double tx, ty, bx0[D0 + 1], bx1[D1 + 1], by0[D0 + 1], by1[D1 + 1];
        locate2<0>(x, P.m, cx, tx);
        locate2<1>(y, P.m, cy, ty);
        stencil<D0>(cx, P.m.n[0], ix);
        stencil<D0>(cy, P.m.n[1], iy);
        basis_pp<D0>(tx, bx0); basis_pp<D1>(tx, bx1);
        basis_pp<D0>(ty, by0); basis_pp<D1>(ty, by1);
        const int nx = P.m.n[0];
        return V3{eval2<D0, D1, 0, 1, D0>(P.f[0], nx, ix, iy, bx0, by1), eval2<D1, D0, 1, 0, D0>(P.f[1], nx, ix, iy, bx1, by0), 0.0};
    }

    static __device__ __forceinline__ void apply(Part2 &p, const PP &P, const double *ft, double *tile, int bx, int by)
    {
        constexpr int D1 = D0 - 1, W = Tile<D0>::W, WS = Tile<D0>::WS, FS = Tile<D0>::FS;
        int cx, cy;
        double tx, ty;
        locate2<0>(p.x[0], P.m, cx, tx);
        locate2<1>(p.x[1], P.m, cy, ty);
        const int rx = rel_cell(cx, bx, P.m.n[0]), ry = rel_cell(cy, by, P.m.n[1]);
        const double v3 = p.v[2];
        const double wv = (p.w * P.op.wscale_dt) * v3;
        double B1, B2;
        if (__builtin_expect(rx >= -kR2 && rx <= kR2 && ry >= -kR2 && ry <= kR2, 1)) {
            double bx0[D0 + 1], bx1[D1 + 1], by0[D0 + 1], by1[D1 + 1];
            basis_pp<D0>(tx, bx0); basis_pp<D1>(tx, bx1);
            basis_pp<D0>(ty, by0); basis_pp<D1>(ty, by1);
            const int lx = rx + kR2, ly = ry + kR2;
            const double *q = ft + ly * WS + lx;
            B1 = eval_tile<D0, D1, WS>(q + WS, bx0, by1);
            B2 = eval_tile<D1, D0, WS>(q + FS + 1, bx1, by0);
            deposit_tile<D0>(tile, lx, ly, bx0, by0, wv);
        } else {
            const V3 bb = slow(p.x[0], p.x[1], P);
            B1 = bb.v0;
            B2 = bb.v1;
            deposit_global<D0>(P.grid, P.m.n[0], P.m.n[1], cx, cy, tx, ty, wv);
        }
        p.v[0] = fma(-(P.op.dtqm * v3), B2, p.v[0]);
        p.v[1] = fma(P.op.dtqm * v3, B1, p.v[1]);
    }
};

// ---- fused [HE x NHE, Hp3]: the leading HE and Hp3 of a Strang step act at the same position, and for NHE = 2 the
// trailing HE of the previous step is folded in as well (it is separated from the leading HE by field-only updates).
// The NHE kicks are linear in the dofs, so the staged tiles hold E_c = sum_h dt_h q/m e_c^(h): one gather per component.
//   source fields: f[0..2] = e (current), f[3..5] = e snapshot before the trailing HE (NHE = 2), f[6] = B1, f[7] = B2
//   staged tiles : E1 E2 E3 B1 B2
template <int D0, int NHE>
struct Op2HEHp3 {
    static constexpr bool DEPOSIT = true, WRITE_X = false, WRITE_V = true, SCATTER = false;
    static constexpr int D = D0, NF = 5;
    struct Params { double dtqm_e[2], dtqm, wscale_dt; };
    using PP = P2<Op2HEHp3>;
    static __device__ __forceinline__ double stage(const PP &P, int f, size_t g)
    {
        if (f >= 3) return P.f[3 + f][g];
        double c = P.op.dtqm_e[0] * P.f[f][g];
        if (NHE == 2) c = fma(P.op.dtqm_e[1], P.f[3 + f][g], c);
        return c;
    }

    struct Out { double e0, e1, e2; };
    // E kicks from global memory (particle outside the window); B and the deposit follow in slow_b
    static __device__ __noinline__ Out slow_e(double x, double y, const PP &P)
    {
        constexpr int D1 = D0 - 1;
        int cx, cy, ix[D0 + 1], iy[D0 + 1];
        double tx, ty, bx0[D0 + 1], bx1[D1 + 1], by0[D0 + 1], by1[D1 + 1];
        locate2<0>(x, P.m, cx, tx);
        locate2<1>(y, P.m, cy, ty);
        stencil<D0>(cx, P.m.n[0], ix);
        stencil<D0>(cy, P.m.n[1], iy);
        basis_pp<D0>(tx, bx0); basis_pp<D1>(tx, bx1);
        basis_pp<D0>(ty, by0); basis_pp<D1>(ty, by1);
        const int nx = P.m.n[0];
        Out o{0.0, 0.0, 0.0};
        for (int h = 0; h < NHE; ++h) {
            o.e0 = fma(P.op.dtqm_e[h], eval2<D1, D0, 1, 0, D0>(P.f[3 * h + 0], nx, ix, iy, bx1, by0), o.e0);
            o.e1 = fma(P.op.dtqm_e[h], eval2<D0, D1, 0, 1, D0>(P.f[3 * h + 1], nx, ix, iy, bx0, by1), o.e1);
            o.e2 = fma(P.op.dtqm_e[h], eval2<D0, D0, 0, 0, D0>(P.f[3 * h + 2], nx, ix, iy, bx0, by0), o.e2);
        }
        return o;
    }
    static __device__ __noinline__ Out slow_b(double x, double y, const PP &P)
    {
        constexpr int D1 = D0 - 1;
        int cx, cy, ix[D0 + 1], iy[D0 + 1];
        double tx, ty, bx0[D0 + 1], bx1[D1 + 1], by0[D0 + 1], by1[D1 + 1];
        locate2<0>(x, P.m, cx, tx);
        locate2<1>(y, P.m, cy, ty);
        stencil<D0>(cx, P.m.n[0], ix);
        stencil<D0>(cy, P.m.n[1], iy);
        basis_pp<D0>(tx, bx0); basis_pp<D1>(tx, bx1);
        basis_pp<D0>(ty, by0); basis_pp<D1>(ty, by1);
        const int nx = P.m.n[0];
        return Out{eval2<D0, D1, 0, 1, D0>(P.f[6], nx, ix, iy, bx0, by1), eval2<D1, D0, 1, 0, D0>(P.f[7], nx, ix, iy, bx1, by0), 0.0};
    }

    static __device__ __forceinline__ void apply(Part2 &p, const PP &P, const double *ft, double *tile, int bx, int by)
    {
        constexpr int D1 = D0 - 1, W = Tile<D0>::W, WS = Tile<D0>::WS, FS = Tile<D0>::FS;
        int cx, cy;
        double tx, ty;
        locate2<0>(p.x[0], P.m, cx, tx);
        locate2<1>(p.x[1], P.m, cy, ty);
        const int rx = rel_cell(cx, bx, P.m.n[0]), ry = rel_cell(cy, by, P.m.n[1]);
        if (__builtin_expect(rx >= -kR2 && rx <= kR2 && ry >= -kR2 && ry <= kR2, 1)) {
            double bx0[D0 + 1], bx1[D1 + 1], by0[D0 + 1], by1[D1 + 1];
            basis_pp<D0>(tx, bx0); basis_pp<D1>(tx, bx1);
            basis_pp<D0>(ty, by0); basis_pp<D1>(ty, by1);
            const int lx = rx + kR2, ly = ry + kR2;
            const double *q = ft + ly * WS + lx;
            p.v[0] += eval_tile<D1, D0, WS>(q + 1, bx1, by0);
            p.v[1] += eval_tile<D0, D1, WS>(q + FS + WS, bx0, by1);
            p.v[2] += eval_tile<D0, D0, WS>(q + 2 * FS, bx0, by0);
            const double v3 = p.v[2];
            const double B1 = eval_tile<D0, D1, WS>(q + 3 * FS + WS, bx0, by1);
            const double B2 = eval_tile<D1, D0, WS>(q + 4 * FS + 1, bx1, by0);
            p.v[0] = fma(-(P.op.dtqm * v3), B2, p.v[0]);
            p.v[1] = fma(P.op.dtqm * v3, B1, p.v[1]);
            deposit_tile<D0>(tile, lx, ly, bx0, by0, (p.w * P.op.wscale_dt) * v3);
        } else {
            const Out e = slow_e(p.x[0], p.x[1], P);
            p.v[0] += e.e0;
            p.v[1] += e.e1;
            p.v[2] += e.e2;
            const double v3 = p.v[2];
            const Out bb = slow_b(p.x[0], p.x[1], P);
            p.v[0] = fma(-(P.op.dtqm * v3), bb.e1, p.v[0]);
            p.v[1] = fma(P.op.dtqm * v3, bb.e0, p.v[1]);
            deposit_global<D0>(P.grid, P.m.n[0], P.m.n[1], cx, cy, tx, ty, (p.w * P.op.wscale_dt) * v3);
        }
    }
};

// ---- operatorHp1 (DIR = 0) / operatorHp2 (DIR = 1) ------------------------------------------------
// f[0] = B3, f[1] = B2 (DIR 0) or B1 (DIR 1).  Along DIR the degree-(p-1) splines are integrated over the
// straight path x_old -> x_new (primitive form, splines.cuh prim_pp; window of p+1 dofs from the lower cell
// as in OpStrangFused); across DIR the degree-p (j, B2/B1) and degree-(p-1) (B3) splines are evaluated.
// SORT: the pass writes every particle (all rows) to its cell-sorted place in the other buffer instead of updating it in
// place -- the periodic cell sort rides in the last push of a Strang step (k2_pass, sorting_hp2).
// HIST (operatorHp1 only): the pass also histograms the cell-sort keys the particles will have after the operatorHp2 push
// that follows (dt_next) -- every input of that push is final once Hp1 is done -- which saves the separate 24 B/particle
// histogram pass of the riding sort.
template <int D0, int DIR, bool SORT = false, bool HIST_ = false>
struct Op2Hp12 {
    static __device__ __forceinline__ double stage(const P2<Op2Hp12> &P, int f, size_t g) { return P.f[f][g]; }
    static constexpr bool DEPOSIT = true, WRITE_X = true, WRITE_V = true, SCATTER = SORT, HIST = HIST_;
    // cell-sort key of the particle after the NEXT operatorHp2 push (HIST)
    static __device__ __forceinline__ int key_next(const Part2 &p, const P2<Op2Hp12> &P)
    {
        return sort_key2(p.x[0], pushed<1>(p.x[1], p.v[1], P.op.dt_next, P.m), P.m);
    }
    // cell-sort key of the particle after this push (SCATTER)
    static __device__ __forceinline__ int key_after(const Part2 &p, const P2<Op2Hp12> &P)
    {
        const double xn = pushed<DIR>(p.x[DIR], p.v[DIR], P.op.dt, P.m);
        return DIR == 0 ? sort_key2(xn, p.x[1], P.m) : sort_key2(p.x[0], xn, P.m);
    }
    static constexpr int D = D0, NF = 2;
    struct Params { double dt, qm_h, wscale_h, dt_next; };   // h = d[DIR]
    using PP = P2<Op2Hp12>;

    // general path: any displacement, global loads and REDs (rare: more than one cell per step or outside the window)
    struct Sums { double z, o; };
    static __device__ __noinline__ Sums general(double x_old, double x_new, double x_t, const PP &P, double wq)
    {
        constexpr int D1 = D0 - 1, O = 1 - DIR;
        const Mesh2 &m = P.m;
        double sum_z = 0.0, sum_o = 0.0;
        int co, cn, ct;
        double to, tn, tt, bt0[D0 + 1], bt1[D1 + 1];
        locate2<DIR>(x_old, m, co, to);
        locate2<DIR>(x_new, m, cn, tn);
        locate2<O>(x_t, m, ct, tt);
        basis_pp<D0>(tt, bt0);
        basis_pp<D1>(tt, bt1);
        const int step = cn >= co ? 1 : -1;
        for (int c = co;; c += step) {
            double lower, upper, sgn;
            if (co == cn) { lower = to; upper = tn; sgn = 1.0; }
            else if (step > 0) { lower = (c == co) ? to : 0.0; upper = (c == cn) ? tn : 1.0; sgn = 1.0; }
            else { lower = (c == cn) ? tn : 0.0; upper = (c == co) ? to : 1.0; sgn = -1.0; }
            double Pl[D1 + 1], Pu[D1 + 1];
            prim_pp<D1>(lower, Pl);
            prim_pp<D1>(upper, Pu);
            for (int k = 0; k <= D1; ++k) {
                const double s = sgn * (Pu[k] - Pl[k]);
                const int ia = wrapi(c - D1 + k, m.n[DIR]);
                for (int b = 0; b <= D0; ++b) {
                    const int it = wrapi(ct - D0 + b, m.n[O]);
                    const size_t idx = DIR == 0 ? (size_t)ia + (size_t)it * m.n[0] : (size_t)it + (size_t)ia * m.n[0];
                    atomicAdd(P.grid + idx, wq * s * bt0[b]);
                    sum_o += P.f[1][idx] * s * bt0[b];
                }
                for (int b = 0; b <= D1; ++b) {
                    const int it = wrapi(ct - D1 + b, m.n[O]);
                    const size_t idx = DIR == 0 ? (size_t)ia + (size_t)it * m.n[0] : (size_t)it + (size_t)ia * m.n[0];
                    sum_z += P.f[0][idx] * s * bt1[b];
                }
            }
            if (c == cn) break;
        }
        return Sums{sum_z, sum_o};
    }

    static __device__ __forceinline__ void apply(Part2 &p, const PP &P, const double *ft, double *tile, int bx, int by)
    {
        constexpr int D1 = D0 - 1, O = 1 - DIR, W = Tile<D0>::W, WS = Tile<D0>::WS, FS = Tile<D0>::FS;
        const Mesh2 &m = P.m;
        const double x_old = p.x[DIR], x_new = fma(P.op.dt, p.v[DIR], x_old);
        int co, cn, ct;
        double to, tn, tt;
        locate2<DIR>(x_old, m, co, to);
        locate2<DIR>(x_new, m, cn, tn);
        locate2<O>(p.x[O], m, ct, tt);
        const double wq = p.w * P.op.wscale_h;
        double sum_z, sum_o;
        const int base_d = DIR == 0 ? bx : by, base_t = DIR == 0 ? by : bx;
        const int ro = rel_cell(co, base_d, m.n[DIR]), rn = rel_cell(cn, base_d, m.n[DIR]), rt = rel_cell(ct, base_t, m.n[O]);
        const int dc = cn - co;
        // the window of D1+2 dofs starts at local index rmin + kR2 + 1 and must end inside the tile: rmin <= kR2 - 1
        const int rmin = min(ro, rn);
        const bool fast = dc >= -1 && dc <= 1 && rmin >= -kR2 && rmin < kR2 && rt >= -kR2 && rt <= kR2;
        if (__builtin_expect(fast, 1)) {
            // window of D1+2 dofs from cmin (see OpStrangFused::work): weight_m = Phi_m(new) - Phi_m(old)
            double A[D1 + 1], B[D1 + 1], win[D1 + 2], bt0[D0 + 1], bt1[D1 + 1];
            prim_pp<D1>(to, A);
            prim_pp<D1>(tn, B);
            basis_pp<D0>(tt, bt0);
            basis_pp<D1>(tt, bt1);
            const bool o1 = co > cn, n1 = cn > co;   // which endpoint sits in the window's second cell
#pragma unroll
            for (int k = 0; k <= D1 + 1; ++k) {
                const double F = k <= D1 ? prim_full<D1>(k <= D1 ? k : 0) : 0.0;
                const double n0 = k <= D1 ? B[k <= D1 ? k : 0] : 0.0, o0 = k <= D1 ? A[k <= D1 ? k : 0] : 0.0;
                const double nn = k == 0 ? F : (k <= D1 ? F + B[k >= 1 ? k - 1 : 0] : B[D1]);
                const double oo = k == 0 ? F : (k <= D1 ? F + A[k >= 1 ? k - 1 : 0] : A[D1]);
                win[k] = (n1 ? nn : n0) - (o1 ? oo : o0);
            }
            const int lw = rmin + kR2 + (D0 - D1), lt = rt + kR2;
            // strides along / across DIR of the deposit tile (sd, st) and of the field tiles (fd, ft_)
            constexpr int sd = DIR == 0 ? 1 : W, st = DIR == 0 ? W : 1;
            constexpr int fd = DIR == 0 ? 1 : WS, ft_ = DIR == 0 ? WS : 1;
            const double *q3 = ft + lw * fd + (lt + 1) * ft_;        // B3: transverse degree D1
            const double *qo = ft + FS + lw * fd + lt * ft_;          // B2 / B1: transverse degree D0
            sum_z = 0.0;
            sum_o = 0.0;
#pragma unroll
            for (int k = 0; k <= D1 + 1; ++k) {
                double sz = q3[k * fd] * bt1[0], so = qo[k * fd] * bt0[0];
#pragma unroll
                for (int b = 1; b <= D1; ++b) sz = fma(q3[k * fd + b * ft_], bt1[b], sz);
#pragma unroll
                for (int b = 1; b <= D0; ++b) so = fma(qo[k * fd + b * ft_], bt0[b], so);
                sum_z = fma(sz, win[k], sum_z);
                sum_o = fma(so, win[k], sum_o);
            }
            double *q = tile + (size_t)(lw * sd + lt * st) * 32;
#pragma unroll
            for (int b = 0; b <= D0; ++b) {
                const double wb = wq * bt0[b];
                double r[D1 + 2];
#pragma unroll
                for (int k = 0; k <= D1 + 1; ++k) r[k] = q[(b * st + k * sd) * 32];
#pragma unroll
                for (int k = 0; k <= D1 + 1; ++k) q[(b * st + k * sd) * 32] = fma(wb, win[k], r[k]);
            }
        } else {
            const Sums g = general(x_old, x_new, p.x[O], P, wq);
            sum_z = g.z;
            sum_o = g.o;
        }
        if (DIR == 0) {
            p.v[1] = fma(-P.op.qm_h, sum_z, p.v[1]);
            p.v[2] = fma(P.op.qm_h, sum_o, p.v[2]);
        } else {
            p.v[0] = fma(P.op.qm_h, sum_z, p.v[0]);
            p.v[2] = fma(-P.op.qm_h, sum_o, p.v[2]);
        }
        p.x[DIR] = pushed<DIR>(x_old, p.v[DIR], P.op.dt, m);   // v[DIR] is not changed by this operator
    }
};

template <class Op>
__device__ __forceinline__ void load2(const Rows2 &r, int64_t i, Part2 &p)
{
    p.x[0] = r.x[0][i];
    p.x[1] = r.x[1][i];
    p.v[0] = r.v[0][i];
    p.v[1] = r.v[1][i];
    p.v[2] = r.v[2][i];
    p.w = r.w[i];
}
template <class Op>
__device__ __forceinline__ void store2(const Rows2 &r, int64_t i, const Part2 &p)
{
    if (Op::WRITE_X) {
        r.x[0][i] = p.x[0];
        r.x[1][i] = p.x[1];
    }
    if (Op::WRITE_V) {
        r.v[0][i] = p.v[0];
        r.v[1][i] = p.v[1];
        r.v[2][i] = p.v[2];
    }
}

// all rows of a particle to slot d of the other buffer (Op::SCATTER)
__device__ __forceinline__ void store2_at(const Rows2 &r, int64_t d, const Part2 &p)
{
    r.x[0][d] = p.x[0];
    r.x[1][d] = p.x[1];
    r.v[0][d] = p.v[0];
    r.v[1][d] = p.v[1];
    r.v[2][d] = p.v[2];
    r.w[d] = p.w;
}
// Slots for the particles of a warp: one cursor reservation per distinct key (key < 0: lane without a particle), in
// two halves so that the round trip of the atomic (1-2 us under load) overlaps the particle arithmetic: reserve_begin
// issues it (the leader keeps the un-consumed result), reserve_end distributes it.  Both are called by all 32 lanes.
struct Reservation {
    unsigned peers;
    int start;
};
__device__ __forceinline__ Reservation reserve_begin(int *__restrict__ cursor, int key, int lane)
{
    Reservation r;
    r.peers = __match_any_sync(0xffffffffu, key);
    r.start = 0;
    if (lane == __ffs(r.peers) - 1 && key >= 0) r.start = atomicAdd(cursor + key, __popc(r.peers));
    return r;
}
__device__ __forceinline__ int64_t reserve_end(const Reservation &r, int lane)
{
    const int start = __shfl_sync(0xffffffffu, r.start, __ffs(r.peers) - 1);
    return (int64_t)start + __popc(r.peers & ((1u << lane) - 1u));
}

// histogram of the cell-sort keys the particles will have after an operatorHp2 push of dt (the scattering pass reserves
// its slots from the exclusive scan of it).  Warp-aggregated shared counters per block, non-zero bins added to `hist`.
__global__ void __launch_bounds__(256) k2_hist_after_hp2(Rows2 r, int64_t n, double dt, Mesh2 m, int ncell, int *__restrict__ hist)
{
    extern __shared__ int shist[];
    for (int i = threadIdx.x; i < ncell; i += 256) shist[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t per_block = ((n + gridDim.x - 1) / gridDim.x + 255) / 256 * 256;
    const int64_t lo = (int64_t)blockIdx.x * per_block, hi = min(n, lo + per_block);
    for (int64_t base = lo + (threadIdx.x & ~31); base < hi; base += 512) {
        // two batches per iteration: both sets of loads are issued before the first key is needed
        const int64_t i0 = base + lane, i1 = base + 256 + lane;
        const bool a0 = i0 < hi, a1 = i1 < hi;
        double x0 = 0, y0 = 0, v0 = 0, x1 = 0, y1 = 0, v1 = 0;
        if (a0) { x0 = r.x[0][i0]; y0 = r.x[1][i0]; v0 = r.v[1][i0]; }
        if (a1) { x1 = r.x[0][i1]; y1 = r.x[1][i1]; v1 = r.v[1][i1]; }
        const int k0 = a0 ? sort_key2(x0, pushed<1>(y0, v0, dt, m), m) : -1;
        const int k1 = a1 ? sort_key2(x1, pushed<1>(y1, v1, dt, m), m) : -1;
        const unsigned p0 = __match_any_sync(0xffffffffu, k0), p1 = __match_any_sync(0xffffffffu, k1);
        if (a0 && (p0 & ((1u << lane) - 1u)) == 0) atomicAdd(&shist[k0], __popc(p0));
        if (a1 && (p1 & ((1u << lane) - 1u)) == 0) atomicAdd(&shist[k1], __popc(p1));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < ncell; i += 256) {
        const int c = shist[i];
        if (c) atomicAdd(&hist[i], c);
    }
}

template <class Op, class = void>
struct op_hist : std::false_type {};
template <class Op>
struct op_hist<Op, std::enable_if_t<Op::HIST>> : std::true_type {};
// hist[key] += 1 for the converged lanes, one atomic per distinct key
__device__ __forceinline__ void hist_add(int *__restrict__ hist, int key, int lane)
{
    const unsigned act = __activemask();
    const unsigned peers = __match_any_sync(act, key);
    if (lane == __ffs(peers) - 1) atomicAdd(hist + key, __popc(peers));
}

// shared memory of one warp: Op::NF field tiles (W*W doubles each, shared by the lanes) followed by the
// lane-private deposit tile (W*W*32 doubles)
template <class Op>
constexpr size_t warp_smem_doubles()
{
    return (size_t)Op::NF * Tile<Op::D>::FS + (Op::DEPOSIT ? (size_t)Tile<Op::D>::SLOTS * 32 : 0);
}

template <class Op>
__global__ void __launch_bounds__(kThreads2, K2_PASS_MINB) k2_pass(const __grid_constant__ P2<Op> P)
{
    extern __shared__ double smem[];
    constexpr int SLOTS = Tile<Op::D>::SLOTS, W = Tile<Op::D>::W, WS = Tile<Op::D>::WS, FS = Tile<Op::D>::FS;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *ftile = smem + (size_t)warp * warp_smem_doubles<Op>();
    double *wtile = ftile + (size_t)Op::NF * FS;
    double *tile = wtile + lane;
    const int nx = P.m.n[0], ny = P.m.n[1];
    const int64_t n_chunks = (P.n + P.chunk - 1) / P.chunk;
    for (int64_t ch = (int64_t)blockIdx.x * kWarps2 + warp; ch < n_chunks; ch += (int64_t)gridDim.x * kWarps2) {
        const int64_t lo = ch * P.chunk, hi = min(P.n, lo + P.chunk);
        // window base = mean cell of 32 particles spread evenly over the chunk (a sorted chunk spans one or two
        // cells plus the drift since the sort, so the mean centres the window); then stage the field tiles
        int bx, by;
        {
            const int64_t is = lo + (int64_t)lane * ((hi - lo + 31) / 32);
            const int64_t ic = is < hi ? is : lo;
            double t;
            int cx, cy;
            locate2<0>(P.r.x[0][ic], P.m, cx, t);
            locate2<1>(P.r.x[1][ic], P.m, cy, t);
            const int cx0 = wrapi(__shfl_sync(0xffffffffu, cx, 0), nx), cy0 = wrapi(__shfl_sync(0xffffffffu, cy, 0), ny);
            int sx = rel_cell(wrapi(cx, nx), cx0, nx), sy = rel_cell(wrapi(cy, ny), cy0, ny);
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                sx += __shfl_xor_sync(0xffffffffu, sx, off);
                sy += __shfl_xor_sync(0xffffffffu, sy, off);
            }
            bx = wrapi(cx0 + __float2int_rn((float)sx * (1.0f / 32.0f)), nx);
            by = wrapi(cy0 + __float2int_rn((float)sy * (1.0f / 32.0f)), ny);
        }
        const int ox = bx - kR2 - Op::D, oy = by - kR2 - Op::D;
        for (int s = lane; s < SLOTS; s += 32) {
            const int ly = s / W, lx = s - ly * W;
            const size_t g = (size_t)wrapi(ox + lx, nx) + (size_t)wrapi(oy + ly, ny) * nx;
#pragma unroll
            for (int f = 0; f < Op::NF; ++f) ftile[f * FS + ly * WS + lx] = Op::stage(P, f, g);
        }
        if (Op::DEPOSIT)
            for (int s = 0; s < SLOTS; ++s) tile[s * 32] = 0.0;
        __syncwarp();
        int64_t i = lo + lane;
        Part2 a, b;
        bool ha = i < hi, hb = i + 32 < hi;
        if (ha) load2<Op>(P.r, i, a);
        if (hb) load2<Op>(P.r, i + 32, b);
        if constexpr (Op::SCATTER) {
            // same stream, but every particle goes to its cell-sorted place in the other buffer.  The trip count is warp
            // uniform (reserve_slot synchronises the warp); the slots are reserved before the arithmetic of the pair, so
            // the round trip of the cursor atomics is hidden behind it.
            for (int64_t base = lo; base < hi; base += 64) {
                const int64_t ni = i + 64;
                const bool hc = ni < hi, hd = ni + 32 < hi;
                Part2 c, d;
                if (hc) load2<Op>(P.r, ni, c);
                if (hd) load2<Op>(P.r, ni + 32, d);
                const Reservation ra = reserve_begin(P.cursor, ha ? Op::key_after(a, P) : -1, lane);
                const Reservation rb = reserve_begin(P.cursor, hb ? Op::key_after(b, P) : -1, lane);
                if (ha) Op::apply(a, P, ftile, tile, bx, by);
                if (hb) Op::apply(b, P, ftile, tile, bx, by);
                const int64_t da = reserve_end(ra, lane), db = reserve_end(rb, lane);
                if (ha) store2_at(P.dst, da, a);
                if (hb) store2_at(P.dst, db, b);
                a = c; b = d;
                ha = hc; hb = hd;
                i = ni;
            }
        } else
        while (ha) {
            const int64_t ni = i + 64;
            const bool hc = ni < hi, hd = ni + 32 < hi;
            Part2 c, d;
            if (hc) load2<Op>(P.r, ni, c);
            if (hd) load2<Op>(P.r, ni + 32, d);
            Op::apply(a, P, ftile, tile, bx, by);
            store2<Op>(P.r, i, a);
            if constexpr (op_hist<Op>::value) hist_add(P.cursor, Op::key_next(a, P), lane);
            if (hb) {
                Op::apply(b, P, ftile, tile, bx, by);
                store2<Op>(P.r, i + 32, b);
                if constexpr (op_hist<Op>::value) hist_add(P.cursor, Op::key_next(b, P), lane);
            }
            a = c; b = d;
            ha = hc; hb = hd;
            i = ni;
        }
        __syncwarp();
        if (Op::DEPOSIT) {
            // sum the 32 lane copies of every slot (rotated: conflict free) and add to the global grid
            for (int s = lane; s < SLOTS; s += 32) {
                double sum = 0.0;
#pragma unroll 8
                for (int l = 0; l < 32; ++l) sum += wtile[(size_t)s * 32 + ((l + lane) & 31)];
                if (sum != 0.0) {
                    const int ly = s / W, lx = s - ly * W;
                    atomicAdd(P.grid + wrapi(ox + lx, nx) + (size_t)wrapi(oy + ly, ny) * nx, sum);
                }
            }
            __syncwarp();
        }
    }
}

// =====================================================================================================
// Sorted fast path (k2_sorted): the operators that START from the cell-sorted particle order.
//
// What bounds k2_pass is the shared-memory pipe (ncu r01z: l1tex data-pipe wavefronts 78-85 % of peak): every gathered
// dof is an LDS.64 (2 wavefronts per warp even when all lanes read the same word) and every deposited dof a
// lane-private LDS + STS (4 wavefronts).  Right after the riding sort, however, the 32 lanes of a warp -- and both
// particles every lane holds -- sit in the SAME cell for hundreds of iterations (64x64 cells, 1.25e8 particles: 30 000
// per cell).  So for the operators that act before the particles drift apart again --
//     [HE x NHE, Hp3, Hp2]  the head of a Strang step (trailing HE of the previous step folded in), and
//     [Hp3]                 its tail (Hp3 follows the sorting Hp2 and does not move anybody) --
//   * the fields come from a per-cell table in GLOBAL memory (k2_build_celltab: the cell polynomials of E1, E2, E3
//     (kick factors folded in), B1, B2 -- b_to_pp_2d, splinepp.jl:329-392 -- and the dof stencils of B3, B1 that the
//     y-push can reach), read with warp-uniform 128-bit loads: one L1 wavefront per two coefficients for 64 particles;
//   * the deposits accumulate in REGISTERS across the whole cell run (j3: (D0+1)^2 dofs of the cell; j2: the D0+1 x-dofs
//     times the D1+3 y-dofs a push of at most one cell can touch, the per-particle window shifted by selects) and are
//     flushed -- warp-reduced, one RED per dof -- when the warp's cell changes;
//   * no shared memory at all.
// An iteration whose 64 particles are not all in one cell (cell-run boundaries, chunk tails) or move more than one cell
// takes the general per-particle path (global loads and REDs), which handles anything.
// =====================================================================================================
template <int D0>
struct CellTab {
    static constexpr int D1 = D0 - 1, NX0 = D0 + 1, NX1 = D1 + 1, ROWS = D1 + 3;
    static constexpr int even(int n) { return (n + 1) & ~1; }
    static constexpr int O_E1 = 0;                                // x-degree D1, y-degree D0: c[j * NX1 + i]
    static constexpr int O_E2 = O_E1 + even(NX1 * NX0);           // x-degree D0, y-degree D1: c[j * NX0 + i]
    static constexpr int O_E3 = O_E2 + even(NX0 * NX1);           // D0 x D0
    static constexpr int O_B1 = O_E3 + even(NX0 * NX0);           // like E2
    static constexpr int O_B2 = O_B1 + even(NX0 * NX1);           // like E1
    static constexpr int O_S3 = O_B2 + even(NX1 * NX0);           // B3 dofs: rows cy-1-D1 .. cy+1, cols cx-D1 .. cx
    static constexpr int O_S1 = O_S3 + even(ROWS * NX1);          // B1 dofs: same rows, cols cx-D0 .. cx
    static constexpr int N = O_S1 + even(ROWS * NX0);             // doubles per cell (even: 16-byte aligned sections)
};

struct CellTabParams {
    const double *e[3], *eT[3], *b[3];
    double dtqm_e[2];
    int nx, ny, nhe;
    double *out;
};

// cell polynomial of a tensor-product spline field: c[j * (DX+1) + i] = coefficient of tx^i ty^j in cell (cx, cy)
template <int DX, int DY, class F>
__device__ __forceinline__ void cell_poly2(int cx, int cy, int nx, int ny, F dof, double *c)
{
#pragma unroll
    for (int q = 0; q < (DX + 1) * (DY + 1); ++q) c[q] = 0.0;
#pragma unroll
    for (int b = 0; b <= DY; ++b) {
        const int gy = wrapi(cy - DY + b, ny);
#pragma unroll
        for (int a = 0; a <= DX; ++a) {
            const double d = dof((size_t)wrapi(cx - DX + a, nx) + (size_t)gy * nx);
#pragma unroll
            for (int j = 0; j <= DY; ++j)
#pragma unroll
                for (int i = 0; i <= DX; ++i) c[j * (DX + 1) + i] = fma(d, pp_coef<DX>(a, i) * pp_coef<DY>(b, j), c[j * (DX + 1) + i]);
        }
    }
}

template <int D0>
__global__ void __launch_bounds__(128) k2_build_celltab(const CellTabParams P)
{
    using T = CellTab<D0>;
    constexpr int D1 = T::D1;
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= P.nx * P.ny) return;
    const int cy = cell / P.nx, cx = cell - cy * P.nx;
    double *o = P.out + (size_t)cell * T::N;
    double c[(D0 + 1) * (D0 + 1)];
    if (P.nhe > 0) {
        auto ef = [&](int k) {
            return [&, k](size_t g) {
                double v = P.dtqm_e[0] * P.e[k][g];
                if (P.nhe == 2) v = fma(P.dtqm_e[1], P.eT[k][g], v);
                return v;
            };
        };
        cell_poly2<D1, D0>(cx, cy, P.nx, P.ny, ef(0), c);
        for (int q = 0; q < T::NX1 * T::NX0; ++q) o[T::O_E1 + q] = c[q];
        cell_poly2<D0, D1>(cx, cy, P.nx, P.ny, ef(1), c);
        for (int q = 0; q < T::NX0 * T::NX1; ++q) o[T::O_E2 + q] = c[q];
        cell_poly2<D0, D0>(cx, cy, P.nx, P.ny, ef(2), c);
        for (int q = 0; q < T::NX0 * T::NX0; ++q) o[T::O_E3 + q] = c[q];
    }
    cell_poly2<D0, D1>(cx, cy, P.nx, P.ny, [&](size_t g) { return P.b[0][g]; }, c);
    for (int q = 0; q < T::NX0 * T::NX1; ++q) o[T::O_B1 + q] = c[q];
    cell_poly2<D1, D0>(cx, cy, P.nx, P.ny, [&](size_t g) { return P.b[1][g]; }, c);
    for (int q = 0; q < T::NX1 * T::NX0; ++q) o[T::O_B2 + q] = c[q];
    for (int r = 0; r < T::ROWS; ++r) {
        const size_t gy = (size_t)wrapi(cy - 1 - D1 + r, P.ny) * P.nx;
        for (int a = 0; a <= D1; ++a) o[T::O_S3 + r * T::NX1 + a] = P.b[2][wrapi(cx - D1 + a, P.nx) + gy];
        for (int a = 0; a <= D0; ++a) o[T::O_S1 + r * T::NX0 + a] = P.b[0][wrapi(cx - D0 + a, P.nx) + gy];
    }
}

// N doubles (N even, 16-byte aligned, warp-uniform address) into registers
template <int N>
__device__ __forceinline__ void ldg_uniform(const double *__restrict__ p, double (&c)[N])
{
    static_assert(N % 2 == 0, "even");
#pragma unroll
    for (int q = 0; q < N; q += 2) {
        const double2 t = __ldg(reinterpret_cast<const double2 *>(p + q));
        c[q] = t.x;
        c[q + 1] = t.y;
    }
}

// value of the cell polynomial c[j * (DX+1) + i] at (tx, ty), for two particles at once
template <int DX, int DY, int NC>
__device__ __forceinline__ void horner2x2(const double (&c)[NC], double txa, double tya, double txb, double tyb, double &va, double &vb)
{
#pragma unroll
    for (int j = DY; j >= 0; --j) {
        double ra = c[j * (DX + 1) + DX], rb = ra;
#pragma unroll
        for (int i = DX - 1; i >= 0; --i) {
            ra = fma(ra, txa, c[j * (DX + 1) + i]);
            rb = fma(rb, txb, c[j * (DX + 1) + i]);
        }
        va = (j == DY) ? ra : fma(va, tya, ra);
        vb = (j == DY) ? rb : fma(vb, tyb, rb);
    }
}

template <int D0>
struct FastParams {
    Rows2 r;
    int64_t n, chunk;
    Mesh2 m;
    const double *tab;                  // CellTab<D0>::N doubles per cell
    const double *e[3], *eT[3], *b[3];  // dof arrays (general path)
    double *j2, *j3;                    // deposit grids (global, zeroed by the launcher)
    double dtqm_e[2];                   // kick factors (current e, snapshot eT)
    int nhe;                            // number of electric kicks folded into the table (general path: 1 or 2)
    double dtqm3, wscale3;              // Hp3: dt q/m, charge * common_weight * dt
    double dt2, qm_h2, wscale_h2;       // Hp2: dt, q/m dy, charge * common_weight * dy
};

// ---- general per-particle path (any cell, any displacement): global loads and REDs --------------------------
// Out of line and BY VALUE: a by-reference particle would pin the particles of the fast path to local memory (ncu r2e:
// 47 % of the stall samples of the Hp3 pass sat on the STL that parked a freshly loaded row on the stack).
struct SortedOut { double x1, v0, v1, v2; };
template <int D0, int NHE, bool HP3, bool HP2>
__device__ __noinline__ SortedOut sorted_general_v(double px0, double px1, double pv0, double pv1, double pv2, double pw,
                                                   const FastParams<D0> &P)
{
    Part2 p;
    p.x[0] = px0; p.x[1] = px1; p.v[0] = pv0; p.v[1] = pv1; p.v[2] = pv2; p.w = pw;
    constexpr int D1 = D0 - 1;
    const int nx = P.m.n[0], ny = P.m.n[1];
    if (NHE > 0 || HP3) {
        int cx, cy, ix[D0 + 1], iy[D0 + 1];
        double tx, ty, bx0[D0 + 1], bx1[D1 + 1], by0[D0 + 1], by1[D1 + 1];
        locate2<0>(p.x[0], P.m, cx, tx);
        locate2<1>(p.x[1], P.m, cy, ty);
        stencil<D0>(cx, nx, ix);
        stencil<D0>(cy, ny, iy);
        basis_pp<D0>(tx, bx0); basis_pp<D1>(tx, bx1);
        basis_pp<D0>(ty, by0); basis_pp<D1>(ty, by1);
        if (NHE > 0) {
            double k0 = 0.0, k1 = 0.0, k2 = 0.0;
            for (int h = 0; h < P.nhe; ++h) {
                const double *const *f = h == 0 ? P.e : P.eT;
                k0 = fma(P.dtqm_e[h], eval2<D1, D0, 1, 0, D0>(f[0], nx, ix, iy, bx1, by0), k0);
                k1 = fma(P.dtqm_e[h], eval2<D0, D1, 0, 1, D0>(f[1], nx, ix, iy, bx0, by1), k1);
                k2 = fma(P.dtqm_e[h], eval2<D0, D0, 0, 0, D0>(f[2], nx, ix, iy, bx0, by0), k2);
            }
            p.v[0] += k0; p.v[1] += k1; p.v[2] += k2;
        }
        if (HP3) {
            const double v3 = p.v[2];
            const double B1 = eval2<D0, D1, 0, 1, D0>(P.b[0], nx, ix, iy, bx0, by1);
            const double B2 = eval2<D1, D0, 1, 0, D0>(P.b[1], nx, ix, iy, bx1, by0);
            p.v[0] = fma(-(P.dtqm3 * v3), B2, p.v[0]);
            p.v[1] = fma(P.dtqm3 * v3, B1, p.v[1]);
            const double wv = (p.w * P.wscale3) * v3;
            for (int b = 0; b <= D0; ++b)
                for (int a = 0; a <= D0; ++a) atomicAdd(P.j3 + ix[a] + (size_t)iy[b] * nx, (wv * by0[b]) * bx0[a]);
        }
    }
    if (HP2) {
        P2<Op2Hp12<D0, 1>> Q;
        Q.m = P.m;
        Q.f[0] = P.b[2];
        Q.f[1] = P.b[0];
        Q.grid = P.j2;
        const double x_old = p.x[1], x_new = fma(P.dt2, p.v[1], x_old);
        const typename Op2Hp12<D0, 1>::Sums g = Op2Hp12<D0, 1>::general(x_old, x_new, p.x[0], Q, p.w * P.wscale_h2);
        p.v[0] = fma(P.qm_h2, g.z, p.v[0]);
        p.v[2] = fma(-P.qm_h2, g.o, p.v[2]);
        p.x[1] = pushed<1>(x_old, p.v[1], P.dt2, P.m);
    }
    return SortedOut{p.x[1], p.v[0], p.v[1], p.v[2]};
}
template <int D0, int NHE, bool HP3, bool HP2>
__device__ __forceinline__ void sorted_general(Part2 &p, const FastParams<D0> &P)
{
    const SortedOut o = sorted_general_v<D0, NHE, HP3, HP2>(p.x[0], p.x[1], p.v[0], p.v[1], p.v[2], p.w, P);
    p.x[1] = o.x1; p.v[0] = o.v0; p.v[1] = o.v1; p.v[2] = o.v2;
}

// registers of one lane that persist across a cell run
template <int D0, bool HP3, bool HP2>
struct FastAcc {
    static constexpr int N3 = HP3 ? (D0 + 1) * (D0 + 1) : 1, N2 = HP2 ? (D0 + 2) * (D0 + 1) : 1;
    double a3[N3];   // j3[b * (D0+1) + a] at dof (cx-D0+a, cy-D0+b)
    double a2[N2];   // j2[r * (D0+1) + a] at dof (cx-D0+a, cy-1-D1+r), r = 0 .. D1+2
};

template <int D0, bool HP3, bool HP2>
__device__ __forceinline__ void fast_flush(FastAcc<D0, HP3, HP2> &A, int bcx, int bcy, const FastParams<D0> &P, int lane)
{
    constexpr int D1 = D0 - 1;
    const int nx = P.m.n[0], ny = P.m.n[1];
    if (HP3) {
#pragma unroll
        for (int q = 0; q < (D0 + 1) * (D0 + 1); ++q) {
            double s = A.a3[q];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
            if (lane == 0 && s != 0.0) {
                const int b = q / (D0 + 1), a = q - b * (D0 + 1);
                atomicAdd(P.j3 + wrapi(bcx - D0 + a, nx) + (size_t)wrapi(bcy - D0 + b, ny) * nx, s);
            }
            A.a3[q] = 0.0;
        }
    }
    if (HP2) {
#pragma unroll
        for (int q = 0; q < (D0 + 2) * (D0 + 1); ++q) {
            double s = A.a2[q];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
            if (lane == 0 && s != 0.0) {
                const int r = q / (D0 + 1), a = q - r * (D0 + 1);
                atomicAdd(P.j2 + wrapi(bcx - D0 + a, nx) + (size_t)wrapi(bcy - 1 - D1 + r, ny) * nx, s);
            }
            A.a2[q] = 0.0;
        }
    }
}

// ---- TMA row stream (PF == 2): every warp runs its own kTmaStages-deep pipeline of bulk copies (cp.async.bulk, one per
// particle row and stage, 64 particles each) into shared memory, completion tracked by one mbarrier per stage.  The loads
// of the next kTmaStages iterations are in flight whatever the register budget; lane 0 issues, all lanes wait.
constexpr int kTmaStages = 4, kTmaRows = 6, kTmaTile = 64;
constexpr size_t kTmaWarpBytes = (size_t)kTmaStages * kTmaRows * kTmaTile * sizeof(double) + kTmaStages * sizeof(unsigned long long);
// lane 0: arm stage `st` and start the six row copies of the 64-particle tile that begins at particle i0
__device__ __forceinline__ void tma_issue(const Rows2 &r, double *buf, unsigned long long *bars, int st, int64_t i0, int64_t hi)
{
    const int cnt = (int)min((int64_t)kTmaTile, hi - i0);
    const unsigned bytes = (unsigned)(((cnt + 1) & ~1) * sizeof(double));   // 16-byte granules (rows are padded to 32 doubles)
    fence_proxy_async();                                                    // the tile's previous contents have been read
    mbar_expect_tx(bars + st, kTmaRows * bytes);
    double *dst = buf + (size_t)st * kTmaRows * kTmaTile;
    const double *src[kTmaRows] = {r.x[0], r.x[1], r.v[0], r.v[1], r.v[2], r.w};
#pragma unroll
    for (int q = 0; q < kTmaRows; ++q) bulk_g2s(dst + q * kTmaTile, src[q] + i0, bytes, bars + st);
}

// MINB: resident blocks per SM the register allocation aims at; PF: 0 plain loads, 1 the next iteration's rows are loaded
// into registers before the current one is processed (software prefetch: 24 more live registers), 2 TMA row stream
template <int D0, int NHE, bool HP3, bool HP2, int MINB, int PF>
__global__ void __launch_bounds__(kThreads2, MINB) k2_sorted(const __grid_constant__ FastParams<D0> P)
{
    using T = CellTab<D0>;
    constexpr int D1 = D0 - 1, NX0 = D0 + 1, NX1 = D1 + 1, ROWS = D1 + 3;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nx = P.m.n[0];
    const int64_t n_chunks = (P.n + P.chunk - 1) / P.chunk;
    FastAcc<D0, HP3, HP2> A;
#pragma unroll
    for (int q = 0; q < A.N3; ++q) A.a3[q] = 0.0;
#pragma unroll
    for (int q = 0; q < A.N2; ++q) A.a2[q] = 0.0;
    // TMA row stream: this warp's stage buffers and barriers
    extern __shared__ __align__(128) unsigned char k2s_smem[];
    double *tbuf = nullptr;
    unsigned long long *tbar = nullptr;
    unsigned tcount = 0;   // tiles consumed so far by this warp: stage = tcount % kTmaStages, parity = (tcount / kTmaStages) & 1
    if constexpr (PF == 2) {
        tbuf = reinterpret_cast<double *>(k2s_smem + (size_t)warp * kTmaWarpBytes);
        tbar = reinterpret_cast<unsigned long long *>(tbuf + (size_t)kTmaStages * kTmaRows * kTmaTile);
        if (lane == 0) {
#pragma unroll
            for (int st = 0; st < kTmaStages; ++st) mbar_init(tbar + st, 1);
            mbar_fence_init();
        }
        __syncwarp();
    }
    for (int64_t ch = (int64_t)blockIdx.x * kWarps2 + warp; ch < n_chunks; ch += (int64_t)gridDim.x * kWarps2) {
        const int64_t lo = ch * P.chunk, hi = min(P.n, lo + P.chunk);
        int bcx = -1, bcy = -1;   // cell the register accumulators belong to
        int64_t i = lo + lane;
        Part2 a, b;
        bool ha = i < hi, hb = i + 32 < hi;
        if constexpr (PF == 2) {
            if (lane == 0) {
#pragma unroll
                for (int st = 0; st < kTmaStages; ++st)
                    if (lo + (int64_t)st * kTmaTile < hi) tma_issue(P.r, tbuf, tbar, (tcount + st) % kTmaStages, lo + (int64_t)st * kTmaTile, hi);
            }
        } else {
            if (ha) load2<void>(P.r, i, a);
            if (hb) load2<void>(P.r, i + 32, b);
        }
        while (__any_sync(0xffffffffu, ha)) {
            const int64_t ni = i + 64;
            const bool hc = ni < hi, hd = ni + 32 < hi;
            Part2 c, d;
            if constexpr (PF == 2) {
                // wait for the tile of this iteration, take the two particles of the lane, hand the stage back to the copies
                const int st = tcount % kTmaStages;
                const unsigned parity = (tcount / kTmaStages) & 1u;
                mbar_wait(tbar + st, parity);
                const double *t = tbuf + (size_t)st * kTmaRows * kTmaTile;
                if (ha) { a.x[0] = t[lane]; a.x[1] = t[kTmaTile + lane]; a.v[0] = t[2 * kTmaTile + lane]; a.v[1] = t[3 * kTmaTile + lane]; a.v[2] = t[4 * kTmaTile + lane]; a.w = t[5 * kTmaTile + lane]; }
                if (hb) { b.x[0] = t[32 + lane]; b.x[1] = t[kTmaTile + 32 + lane]; b.v[0] = t[2 * kTmaTile + 32 + lane]; b.v[1] = t[3 * kTmaTile + 32 + lane]; b.v[2] = t[4 * kTmaTile + 32 + lane]; b.w = t[5 * kTmaTile + 32 + lane]; }
                __syncwarp();
                const int64_t i_next = (i - lane) + (int64_t)kTmaStages * kTmaTile;
                if (lane == 0 && i_next < hi) tma_issue(P.r, tbuf, tbar, st, i_next, hi);
                ++tcount;
            }
            if constexpr (PF == 1) {
                if (hc) load2<void>(P.r, ni, c);
                if (hd) load2<void>(P.r, ni + 32, d);
            }
            // ---- all 64 particles in one cell?
            int cxa, cya, cxb, cyb;
            double txa, tya, txb, tyb;
            locate2<0>(ha ? a.x[0] : 0.0, P.m, cxa, txa);
            locate2<1>(ha ? a.x[1] : 0.0, P.m, cya, tya);
            locate2<0>(hb ? b.x[0] : 0.0, P.m, cxb, txb);
            locate2<1>(hb ? b.x[1] : 0.0, P.m, cyb, tyb);
            const int kx = __shfl_sync(0xffffffffu, cxa, 0), ky = __shfl_sync(0xffffffffu, cya, 0);
            const bool inside = (unsigned)kx < (unsigned)nx && (unsigned)ky < (unsigned)P.m.n[1];
            const bool uni = __all_sync(0xffffffffu, ha && hb && cxa == kx && cxb == kx && cya == ky && cyb == ky) && inside;
            if (uni) {
                if (kx != bcx || ky != bcy) {
                    if (bcx >= 0) fast_flush<D0, HP3, HP2>(A, bcx, bcy, P, lane);
                    bcx = kx;
                    bcy = ky;
                }
                const double *tab = P.tab + ((size_t)kx + (size_t)ky * nx) * T::N;
                if (NHE > 0) {
                    double ka, kb;
                    {
                        double c1[T::even(NX1 * NX0)];
                        ldg_uniform(tab + T::O_E1, c1);
                        horner2x2<D1, D0>(c1, txa, tya, txb, tyb, ka, kb);
                        a.v[0] += ka; b.v[0] += kb;
                    }
                    {
                        double c2[T::even(NX0 * NX1)];
                        ldg_uniform(tab + T::O_E2, c2);
                        horner2x2<D0, D1>(c2, txa, tya, txb, tyb, ka, kb);
                        a.v[1] += ka; b.v[1] += kb;
                    }
                    {
                        double c3[T::even(NX0 * NX0)];
                        ldg_uniform(tab + T::O_E3, c3);
                        horner2x2<D0, D0>(c3, txa, tya, txb, tyb, ka, kb);
                        a.v[2] += ka; b.v[2] += kb;
                    }
                }
                double bxa[NX0], bxb[NX0];   // degree-D0 x basis: j3 and j2 deposits, B1 line integral
                basis_pp<D0>(txa, bxa);
                basis_pp<D0>(txb, bxb);
                if (HP3) {
                    double B1a, B1b, B2a, B2b;
                    {
                        double c1[T::even(NX0 * NX1)];
                        ldg_uniform(tab + T::O_B1, c1);
                        horner2x2<D0, D1>(c1, txa, tya, txb, tyb, B1a, B1b);
                    }
                    {
                        double c2[T::even(NX1 * NX0)];
                        ldg_uniform(tab + T::O_B2, c2);
                        horner2x2<D1, D0>(c2, txa, tya, txb, tyb, B2a, B2b);
                    }
                    const double v3a = a.v[2], v3b = b.v[2];
                    a.v[0] = fma(-(P.dtqm3 * v3a), B2a, a.v[0]);
                    a.v[1] = fma(P.dtqm3 * v3a, B1a, a.v[1]);
                    b.v[0] = fma(-(P.dtqm3 * v3b), B2b, b.v[0]);
                    b.v[1] = fma(P.dtqm3 * v3b, B1b, b.v[1]);
                    double bya[NX0], byb[NX0];
                    basis_pp<D0>(tya, bya);
                    basis_pp<D0>(tyb, byb);
                    const double wa = (a.w * P.wscale3) * v3a, wb = (b.w * P.wscale3) * v3b;
#pragma unroll
                    for (int q = 0; q <= D0; ++q) {
                        const double ya = wa * bya[q], yb = wb * byb[q];
#pragma unroll
                        for (int k = 0; k <= D0; ++k) A.a3[q * NX0 + k] = fma(yb, bxb[k], fma(ya, bxa[k], A.a3[q * NX0 + k]));
                    }
                }
                if (HP2) {
                    // operatorHp2 (Op2Hp12<D0, 1>::apply) from the common cell (kx, ky)
                    const double yoa = a.x[1], yob = b.x[1];
                    const double yna = fma(P.dt2, a.v[1], yoa), ynb = fma(P.dt2, b.v[1], yob);
                    int cna, cnb;
                    double tna, tnb;
                    locate2<1>(yna, P.m, cna, tna);
                    locate2<1>(ynb, P.m, cnb, tnb);
                    const int dca = cna - ky, dcb = cnb - ky;
                    if (__all_sync(0xffffffffu, dca >= -1 && dca <= 1 && dcb >= -1 && dcb <= 1)) {
                        // window weights of the D1+2 dofs from min(cell_old, cell_new) (see OpStrangFused::work), placed
                        // in the D1+3 rows cy-1-D1 .. cy+1 a push of at most one cell can reach
                        double w5a[ROWS], w5b[ROWS];
                        auto window = [&](double to, double tn, int dc, double (&w5)[ROWS]) {
                            double Ao[D1 + 1], Bn[D1 + 1], win[D1 + 2];
                            prim_pp<D1>(to, Ao);
                            prim_pp<D1>(tn, Bn);
                            const bool o1 = dc < 0, n1 = dc > 0;
#pragma unroll
                            for (int k = 0; k <= D1 + 1; ++k) {
                                const double F = k <= D1 ? prim_full<D1>(k <= D1 ? k : 0) : 0.0;
                                const double n0 = k <= D1 ? Bn[k <= D1 ? k : 0] : 0.0, o0 = k <= D1 ? Ao[k <= D1 ? k : 0] : 0.0;
                                const double nn = k == 0 ? F : (k <= D1 ? F + Bn[k >= 1 ? k - 1 : 0] : Bn[D1]);
                                const double oo = k == 0 ? F : (k <= D1 ? F + Ao[k >= 1 ? k - 1 : 0] : Ao[D1]);
                                win[k] = (n1 ? nn : n0) - (o1 ? oo : o0);
                            }
                            // the window starts at row 0 when the particle moves down (cmin = cy - 1), at row 1 otherwise
#pragma unroll
                            for (int r = 0; r < ROWS; ++r) {
                                const double lo_ = r <= D1 + 1 ? win[r <= D1 + 1 ? r : 0] : 0.0;
                                const double hi_ = r >= 1 ? win[r >= 1 ? r - 1 : 0] : 0.0;
                                w5[r] = o1 ? lo_ : hi_;
                            }
                        };
                        window(tya, tna, dca, w5a);
                        window(tyb, tnb, dcb, w5b);
                        double bta[NX1], btb[NX1];   // degree-D1 x basis (B3 is D1 x D1)
                        basis_pp<D1>(txa, bta);
                        basis_pp<D1>(txb, btb);
                        double sza = 0.0, szb = 0.0, soa = 0.0, sob = 0.0;
                        {
                            double s3[T::even(ROWS * NX1)];
                            ldg_uniform(tab + T::O_S3, s3);
#pragma unroll
                            for (int r = 0; r < ROWS; ++r) {
                                double ra = s3[r * NX1] * bta[0], rb = s3[r * NX1] * btb[0];
#pragma unroll
                                for (int k = 1; k <= D1; ++k) {
                                    ra = fma(s3[r * NX1 + k], bta[k], ra);
                                    rb = fma(s3[r * NX1 + k], btb[k], rb);
                                }
                                sza = fma(ra, w5a[r], sza);
                                szb = fma(rb, w5b[r], szb);
                            }
                        }
                        {
                            double s1[T::even(ROWS * NX0)];
                            ldg_uniform(tab + T::O_S1, s1);
#pragma unroll
                            for (int r = 0; r < ROWS; ++r) {
                                double ra = s1[r * NX0] * bxa[0], rb = s1[r * NX0] * bxb[0];
#pragma unroll
                                for (int k = 1; k <= D0; ++k) {
                                    ra = fma(s1[r * NX0 + k], bxa[k], ra);
                                    rb = fma(s1[r * NX0 + k], bxb[k], rb);
                                }
                                soa = fma(ra, w5a[r], soa);
                                sob = fma(rb, w5b[r], sob);
                            }
                        }
                        const double wqa = a.w * P.wscale_h2, wqb = b.w * P.wscale_h2;
#pragma unroll
                        for (int r = 0; r < ROWS; ++r) {
                            const double ya = wqa * w5a[r], yb = wqb * w5b[r];
#pragma unroll
                            for (int k = 0; k <= D0; ++k) A.a2[r * NX0 + k] = fma(yb, bxb[k], fma(ya, bxa[k], A.a2[r * NX0 + k]));
                        }
                        a.v[0] = fma(P.qm_h2, sza, a.v[0]);
                        a.v[2] = fma(-P.qm_h2, soa, a.v[2]);
                        b.v[0] = fma(P.qm_h2, szb, b.v[0]);
                        b.v[2] = fma(-P.qm_h2, sob, b.v[2]);
                        a.x[1] = pushed<1>(yoa, a.v[1], P.dt2, P.m);
                        b.x[1] = pushed<1>(yob, b.v[1], P.dt2, P.m);
                    } else {
                        sorted_general<D0, 0, false, true>(a, P);
                        sorted_general<D0, 0, false, true>(b, P);
                    }
                }
            } else {
                if (ha) sorted_general<D0, NHE, HP3, HP2>(a, P);
                if (hb) sorted_general<D0, NHE, HP3, HP2>(b, P);
            }
            if (ha) {
                if (HP2) P.r.x[1][i] = a.x[1];
                P.r.v[0][i] = a.v[0];
                P.r.v[1][i] = a.v[1];
                if (NHE > 0 || HP2) P.r.v[2][i] = a.v[2];
            }
            if (hb) {
                if (HP2) P.r.x[1][i + 32] = b.x[1];
                P.r.v[0][i + 32] = b.v[0];
                P.r.v[1][i + 32] = b.v[1];
                if (NHE > 0 || HP2) P.r.v[2][i + 32] = b.v[2];
            }
            if constexpr (PF == 0) {
                if (hc) load2<void>(P.r, ni, c);
                if (hd) load2<void>(P.r, ni + 32, d);
            }
            if constexpr (PF != 2) { a = c; b = d; }
            ha = hc; hb = hd;
            i = ni;
        }
        if (bcx >= 0) fast_flush<D0, HP3, HP2>(A, bcx, bcy, P, lane);
    }
}

// sum_p w v.v, sum_p w v_k: one block-reduced partial per block, finished by k_reduce_partials
__global__ void __launch_bounds__(256) k2_moments(Rows2 r, int64_t n, double *__restrict__ partials)
{
    __shared__ double red[4][256];
    double s[4] = {0, 0, 0, 0};
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const double w = r.w[i], v1 = r.v[0][i], v2 = r.v[1][i], v3 = r.v[2][i];
        s[0] += w * (v1 * v1 + v2 * v2 + v3 * v3);
        s[1] += w * v1;
        s[2] += w * v2;
        s[3] += w * v3;
    }
    for (int k = 0; k < 4; ++k) red[k][threadIdx.x] = s[k];
    __syncthreads();
    for (int h = 128; h > 0; h >>= 1) {
        if (threadIdx.x < h)
            for (int k = 0; k < 4; ++k) red[k][threadIdx.x] += red[k][threadIdx.x + h];
        __syncthreads();
    }
    if (threadIdx.x < 4) partials[(size_t)blockIdx.x * 4 + threadIdx.x] = red[threadIdx.x][0];
}

// ---- host side -----------------------------------------------------------------------------------
static Mesh2 mesh2(const Maxwell2D &mx)
{
    Mesh2 m;
    m.xmin[0] = mx.xmin; m.xmin[1] = mx.ymin;
    m.d[0] = mx.dx; m.d[1] = mx.dy;
    m.inv_d[0] = 1.0 / mx.dx; m.inv_d[1] = 1.0 / mx.dy;
    m.L[0] = mx.Lx; m.L[1] = mx.Ly;
    m.xmax[0] = mx.xmin + mx.Lx; m.xmax[1] = mx.ymin + mx.Ly;
    m.n[0] = mx.nx; m.n[1] = mx.ny;
    return m;
}

static Rows2 rows2(ParticleGroup &pg)
{
    Rows2 r;
    r.x[0] = pg.row(0); r.x[1] = pg.row(1);
    r.v[0] = pg.row(2); r.v[1] = pg.row(3); r.v[2] = pg.row(4);
    r.w = pg.row(5);
    return r;
}

template <class Op>
static void launch2(Splitting2D &h, P2<Op> P, const char *tag)
{
    Context &c = ctx();
    P.r = rows2(*h.pg);
    P.n = h.pg->n;
    P.m = mesh2(*h.maxwell);
    if (P.n <= 0) return;
    const size_t smem = (size_t)kWarps2 * warp_smem_doubles<Op>() * sizeof(double);
    if (smem > 48 * 1024) ensure_func_smem((const void *)k2_pass<Op>, smem);
    int per_sm = 0;
    GP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k2_pass<Op>, kThreads2, smem));
    GP_REQUIRE(per_sm >= 1, GEMPIC_EINVAL, "2D pass does not fit on an SM (smem %zu B)", smem);
    // chunk: a multiple of 64 particles per warp visit, ~16 visits per warp so that the tail stays balanced
    const int64_t warps = (int64_t)c.sm_count * per_sm * kWarps2;
    int64_t chunk = (P.n + warps * 16 - 1) / (warps * 16);
    chunk = std::max<int64_t>(512, std::min<int64_t>((chunk + 63) / 64 * 64, 8192));
    P.chunk = chunk;
    const int64_t n_chunks = (P.n + chunk - 1) / chunk;
    const int grid = (int)std::min<int64_t>((n_chunks + kWarps2 - 1) / kWarps2, (int64_t)c.sm_count * per_sm);
    if (tag) profile_begin(tag);
    k2_pass<Op><<<grid, kThreads2, smem, c.stream>>>(P);
    GP_CUDA(cudaGetLastError());
    if (tag) profile_end(tag);
    count_launch();
    if (Op::WRITE_X || Op::WRITE_V) particles_changed();
}

#define GP_DISPATCH_D0(Dv, ...)                                                           \
    do {                                                                                  \
        switch (Dv) {                                                                     \
        case 1: { constexpr int D0 = 1; __VA_ARGS__; } break;                             \
        case 2: { constexpr int D0 = 2; __VA_ARGS__; } break;                             \
        case 3: { constexpr int D0 = 3; __VA_ARGS__; } break;                             \
        default: ::gempic::fail(GEMPIC_EINVAL, "unsupported spline degree %d", (Dv));      \
        }                                                                                 \
    } while (0)

static void zero_grid(double *g, size_t n) { GP_CUDA(cudaMemsetAsync(g, 0, n * sizeof(double), ctx().stream)); }

void hs2d_charge(Splitting2D &h, double *rho_dev)
{
    zero_grid(rho_dev, h.nd);
    GP_DISPATCH_D0(h.maxwell->s_deg_0, {
        P2<Op2Charge<D0>> P{};
        P.grid = rho_dev;
        P.op.wscale = h.pg->charge * h.pg->common_weight;
        launch2(h, P, "add_charge2d");
    });
    allreduce_sum(rho_dev, h.nd);
}

// the particle kick of operatorHE with the given field vectors
static void op2_HE_particles(Splitting2D &h, double dt, const double *e1, const double *e2, const double *e3)
{
    GP_DISPATCH_D0(h.maxwell->s_deg_0, {
        P2<Op2HE<D0>> P{};
        P.f[0] = e1; P.f[1] = e2; P.f[2] = e3;
        P.op.dtqm = dt * h.pg->q_over_m;
        launch2(h, P, "operatorHE{2,3}");
    });
}

static void op2_HE(Splitting2D &h, double dt)
{
    op2_HE_particles(h, dt, h.e(0), h.e(1), h.e(2));
    double *b[3] = {h.b(0), h.b(1), h.b(2)};
    const double *e[3] = {h.e(0), h.e(1), h.e(2)};
    m2d_b_from_e(*h.maxwell, b, dt, e);
}

static void op2_HB(Splitting2D &h, double dt)
{
    double *e[3] = {h.e(0), h.e(1), h.e(2)};
    const double *b[3] = {h.b(0), h.b(1), h.b(2)};
    m2d_e_from_b(*h.maxwell, e, dt, b);
}

static void op2_Hp3(Splitting2D &h, double dt)
{
    zero_grid(h.j(2), h.nd);
    GP_DISPATCH_D0(h.maxwell->s_deg_0, {
        P2<Op2Hp3<D0>> P{};
        P.f[0] = h.b(0); P.f[1] = h.b(1);
        P.grid = h.j(2);
        P.op.dtqm = dt * h.pg->q_over_m;
        P.op.wscale_dt = h.pg->charge * h.pg->common_weight * dt;
        launch2(h, P, "operatorHp3{2,3}");
    });
    allreduce_sum(h.j(2), h.nd);
    m2d_e_from_j(*h.maxwell, h.e(2), h.j(2), 3);
}

template <int DIR>
static void op2_Hp12(Splitting2D &h, double dt)
{
    zero_grid(h.j(DIR), h.nd);
    const double hd = DIR == 0 ? h.maxwell->dx : h.maxwell->dy;
    GP_DISPATCH_D0(h.maxwell->s_deg_0, {
        P2<Op2Hp12<D0, DIR>> P{};
        P.f[0] = h.b(2);
        P.f[1] = DIR == 0 ? h.b(1) : h.b(0);
        P.grid = h.j(DIR);
        P.op.dt = dt;
        P.op.qm_h = h.pg->q_over_m * hd;
        P.op.wscale_h = h.pg->charge * h.pg->common_weight * hd;
        launch2(h, P, DIR == 0 ? "operatorHp1{2,3}" : "operatorHp2{2,3}");
    });
    allreduce_sum(h.j(DIR), h.nd);
    m2d_e_from_j(*h.maxwell, h.e(DIR), h.j(DIR), DIR + 1);
}

// operatorHp1 that also histograms the sort keys after the operatorHp2(dt_next) push that follows (Op2Hp12 HIST)
static void op2_Hp1_with_hist(Splitting2D &h, double dt, double dt_next)
{
    Context &c = ctx();
    ParticleGroup &pg = *h.pg;
    const int cells = (int)h.nd;
    if (pg.sort_keys.n < (size_t)cells) pg.sort_keys.alloc(cells);
    GP_CUDA(cudaMemsetAsync(pg.sort_keys.p, 0, sizeof(int) * cells, c.stream));
    zero_grid(h.j(0), h.nd);
    GP_DISPATCH_D0(h.maxwell->s_deg_0, {
        P2<Op2Hp12<D0, 0, false, true>> P{};
        P.f[0] = h.b(2);
        P.f[1] = h.b(1);
        P.grid = h.j(0);
        P.cursor = pg.sort_keys.p;
        P.op.dt = dt;
        P.op.qm_h = pg.q_over_m * h.maxwell->dx;
        P.op.wscale_h = pg.charge * pg.common_weight * h.maxwell->dx;
        P.op.dt_next = dt_next;
        launch2(h, P, "operatorHp1{2,3}+hist");
    });
    pg.sorted2d = false;
    allreduce_sum(h.j(0), h.nd);
    m2d_e_from_j(*h.maxwell, h.e(0), h.j(0), 1);
}

// operatorHp2 that leaves the particles cell sorted: histogram of the cells after the push (taken by the preceding
// operatorHp1 pass when have_hist, else by k2_hist_after_hp2) -> scan -> the push pass places every particle (all rows)
// at its sorted position in the other buffer -> swap.  96 (+ 24) B/particle instead of 72 (push in place) + 112
// (stand-alone sort).
static void sorting_hp2(Splitting2D &h, double dt, bool have_hist = false)
{
    Context &c = ctx();
    ParticleGroup &pg = *h.pg;
    const int cells = (int)h.nd;
    zero_grid(h.j(1), h.nd);
    if (pg.sort_keys.n < (size_t)cells) pg.sort_keys.alloc(cells);
    if (pg.sort_tmp.n < pg.data.n) pg.sort_tmp.alloc(pg.data.n);
    int *hist = pg.sort_keys.p;
    const Mesh2 m = mesh2(*h.maxwell);
    if (!have_hist) {
        GP_CUDA(cudaMemsetAsync(hist, 0, sizeof(int) * cells, c.stream));
        const size_t hsmem = (size_t)cells * sizeof(int);
        GP_REQUIRE(hsmem <= kSmemMaxOptin, GEMPIC_EINVAL, "riding sort: %d cells exceed the shared-memory histogram", cells);
        if (hsmem > 48 * 1024) ensure_func_smem((const void *)k2_hist_after_hp2, hsmem);
        profile_begin("cell histogram after Hp2");
        const int hgrid = (int)std::min<int64_t>((int64_t)c.sm_count * 8, (pg.n + 255) / 256);
        k2_hist_after_hp2<<<hgrid, 256, hsmem, c.stream>>>(rows2(pg), pg.n, dt, m, cells, hist);
        GP_CUDA(cudaGetLastError());
        profile_end("cell histogram after Hp2");
        count_launch();
    }
    sort_scan(hist, cells);
    // rows of the other buffer
    Rows2 dst;
    {
        double *base = pg.sort_tmp.p;
        dst.x[0] = base; dst.x[1] = base + pg.stride;
        dst.v[0] = base + 2 * pg.stride; dst.v[1] = base + 3 * pg.stride; dst.v[2] = base + 4 * pg.stride;
        dst.w = base + 5 * pg.stride;
    }
    GP_DISPATCH_D0(h.maxwell->s_deg_0, {
        P2<Op2Hp12<D0, 1, true>> P{};
        P.f[0] = h.b(2);
        P.f[1] = h.b(0);
        P.grid = h.j(1);
        P.dst = dst;
        P.cursor = hist;
        P.op.dt = dt;
        P.op.qm_h = h.pg->q_over_m * h.maxwell->dy;
        P.op.wscale_h = h.pg->charge * h.pg->common_weight * h.maxwell->dy;
        launch2(h, P, "operatorHp2{2,3}+sort");
    });
    std::swap(pg.data.p, pg.sort_tmp.p);
    std::swap(pg.data.n, pg.sort_tmp.n);
    pg.generation++;
    pg.sorted2d = true;
    allreduce_sum(h.j(1), h.nd);
    m2d_e_from_j(*h.maxwell, h.e(1), h.j(1), 2);
}

void hs2d_operator(Splitting2D &h, int op, double dt)
{
    if (op == GEMPIC_OP_HP1 || op == GEMPIC_OP_HP2) h.pg->sorted2d = false;
    switch (op) {
    case GEMPIC_OP_HP1: op2_Hp12<0>(h, dt); break;
    case GEMPIC_OP_HP2: op2_Hp12<1>(h, dt); break;
    case GEMPIC_OP_HP3: op2_Hp3(h, dt); break;
    case GEMPIC_OP_HE: op2_HE(h, dt); break;
    case GEMPIC_OP_HB: op2_HB(h, dt); break;
    default: fail(GEMPIC_EINVAL, "unknown operator %d", op);
    }
}

// [HE x n_he, Hp3](dt/2) in one pass + the e3 solve; n_he = 2 reads the snapshot eT of the trailing HE's field
static void fused_he_hp3(Splitting2D &h, double dt, int n_he, double dt_T)
{
    zero_grid(h.j(2), h.nd);
    const double qm = h.pg->q_over_m;
    GP_DISPATCH_D0(h.maxwell->s_deg_0, {
        if (n_he == 1) {
            P2<Op2HEHp3<D0, 1>> P{};
            for (int c = 0; c < 3; ++c) P.f[c] = h.e(c);
            P.f[6] = h.b(0); P.f[7] = h.b(1);
            P.grid = h.j(2);
            P.op.dtqm_e[0] = P.op.dtqm_e[1] = 0.5 * dt * qm;
            P.op.dtqm = 0.5 * dt * qm;
            P.op.wscale_dt = h.pg->charge * h.pg->common_weight * 0.5 * dt;
            launch2(h, P, "fused[HE,Hp3]{2,3}");
        } else {
            P2<Op2HEHp3<D0, 2>> P{};
            for (int c = 0; c < 3; ++c) { P.f[c] = h.e(c); P.f[3 + c] = h.eT(c); }
            P.f[6] = h.b(0); P.f[7] = h.b(1);
            P.grid = h.j(2);
            P.op.dtqm_e[0] = 0.5 * dt * qm;
            P.op.dtqm_e[1] = 0.5 * dt_T * qm;   // trailing HE of the previous step: snapshot fields eT, that step's dt
            P.op.dtqm = 0.5 * dt * qm;
            P.op.wscale_dt = h.pg->charge * h.pg->common_weight * 0.5 * dt;
            launch2(h, P, "fused[HE,HE,Hp3]{2,3}");
        }
    });
    allreduce_sum(h.j(2), h.nd);
    m2d_e_from_j(*h.maxwell, h.e(2), h.j(2), 3);
}

// ---- sorted fast path: host side ---------------------------------------------------------------------------
template <int D0>
static void build_celltab(Splitting2D &h, int nhe, double dtqm_e0, double dtqm_e1)
{
    using T = CellTab<D0>;
    const size_t need = h.nd * (size_t)T::N;
    if (h.celltab.n < need) h.celltab.alloc(need);
    CellTabParams P{};
    for (int c = 0; c < 3; ++c) { P.e[c] = h.e(c); P.eT[c] = h.eT(c); P.b[c] = h.b(c); }
    P.dtqm_e[0] = dtqm_e0; P.dtqm_e[1] = dtqm_e1;
    P.nx = h.maxwell->nx; P.ny = h.maxwell->ny; P.nhe = nhe;
    P.out = h.celltab.p;
    k2_build_celltab<D0><<<(int)((h.nd + 127) / 128), 128, 0, ctx().stream>>>(P);
    GP_CUDA(cudaGetLastError());
    count_launch();
}

template <int D0, int NHE, bool HP3, bool HP2, int MINB, int PF>
static void launch_sorted_v(Splitting2D &h, FastParams<D0> P, const char *tag);

// kernel variant: GEMPIC_K2_HEAD / GEMPIC_K2_TAIL = "<blocks per SM><n|p|t>" (n: plain loads, p: register prefetch,
// t: TMA row stream), for tuning runs; code = blocks * 3 + fetch mode
static int sorted_variant(bool head)
{
    auto parse = [](const char *name, int dflt) {
        const char *e = getenv(name);
        if (!e || !e[0] || !e[1]) return dflt;
        return (e[0] - '0') * 3 + (e[1] == 'p' ? 1 : e[1] == 't' ? 2 : 0);
    };
    static const int vh = parse("GEMPIC_K2_HEAD", 2 * 3 + 1);
    static const int vt = parse("GEMPIC_K2_TAIL", 3 * 3 + 2);
    return head ? vh : vt;
}

template <int D0, int NHE, bool HP3, bool HP2>
static void launch_sorted(Splitting2D &h, FastParams<D0> P, const char *tag)
{
    if constexpr (D0 == 3) {
        switch (sorted_variant(HP2)) {
        case 2 * 3 + 0: return launch_sorted_v<D0, NHE, HP3, HP2, 2, 0>(h, P, tag);
        case 2 * 3 + 2: return launch_sorted_v<D0, NHE, HP3, HP2, 2, 2>(h, P, tag);
        case 3 * 3 + 0: return launch_sorted_v<D0, NHE, HP3, HP2, 3, 0>(h, P, tag);
        case 3 * 3 + 1: return launch_sorted_v<D0, NHE, HP3, HP2, 3, 1>(h, P, tag);
        case 3 * 3 + 2: return launch_sorted_v<D0, NHE, HP3, HP2, 3, 2>(h, P, tag);
        case 4 * 3 + 2: return launch_sorted_v<D0, NHE, HP3, HP2, 4, 2>(h, P, tag);
        default: break;
        }
    }
    launch_sorted_v<D0, NHE, HP3, HP2, 2, 1>(h, P, tag);
}

template <int D0, int NHE, bool HP3, bool HP2, int MINB, int PF>
static void launch_sorted_v(Splitting2D &h, FastParams<D0> P, const char *tag)
{
    Context &c = ctx();
    const size_t smem = PF == 2 ? (size_t)kWarps2 * kTmaWarpBytes : 0;
    if (smem > 48 * 1024) ensure_func_smem((const void *)k2_sorted<D0, NHE, HP3, HP2, MINB, PF>, smem);
    P.r = rows2(*h.pg);
    P.n = h.pg->n;
    P.m = mesh2(*h.maxwell);
    P.tab = h.celltab.p;
    for (int k = 0; k < 3; ++k) { P.e[k] = h.e(k); P.eT[k] = h.eT(k); P.b[k] = h.b(k); }
    P.j2 = h.j(1);
    P.j3 = h.j(2);
    if (P.n <= 0) return;
    int per_sm = 0;
    GP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k2_sorted<D0, NHE, HP3, HP2, MINB, PF>, kThreads2, smem));
    GP_REQUIRE(per_sm >= 1, GEMPIC_EINVAL, "sorted pass does not fit on an SM");
    const int64_t warps = (int64_t)c.sm_count * per_sm * kWarps2;
    int64_t chunk = (P.n + warps * 16 - 1) / (warps * 16);
    chunk = std::max<int64_t>(512, std::min<int64_t>((chunk + 63) / 64 * 64, 16384));
    P.chunk = chunk;
    const int64_t n_chunks = (P.n + chunk - 1) / chunk;
    const int grid = (int)std::min<int64_t>((n_chunks + kWarps2 - 1) / kWarps2, (int64_t)c.sm_count * per_sm);
    if (tag) profile_begin(tag);
    k2_sorted<D0, NHE, HP3, HP2, MINB, PF><<<grid, kThreads2, smem, c.stream>>>(P);
    GP_CUDA(cudaGetLastError());
    if (tag) profile_end(tag);
    count_launch();
    particles_changed();
}

// head of a Strang step from the cell-sorted order: [HE x n_he, Hp3, Hp2](dt/2) in one register-resident pass + the
// e3 and e2 solves (nothing between them reads e)
static void sorted_head(Splitting2D &h, double dt, int n_he, double dt_T)
{
    const double qm = h.pg->q_over_m, cq = h.pg->charge * h.pg->common_weight;
    GP_CUDA(cudaMemsetAsync(h.j(1), 0, 2 * h.nd * sizeof(double), ctx().stream));   // j2 | j3 adjacent
    GP_DISPATCH_D0(h.maxwell->s_deg_0, {
        build_celltab<D0>(h, n_he, 0.5 * dt * qm, 0.5 * dt_T * qm);
        FastParams<D0> P{};
        P.dtqm_e[0] = 0.5 * dt * qm;
        P.dtqm_e[1] = 0.5 * dt_T * qm;
        P.nhe = n_he;
        P.dtqm3 = 0.5 * dt * qm;
        P.wscale3 = cq * 0.5 * dt;
        P.dt2 = 0.5 * dt;
        P.qm_h2 = qm * h.maxwell->dy;
        P.wscale_h2 = cq * h.maxwell->dy;
        // (the kernel is the same for one or two kicks: their factors are folded into the table)
        launch_sorted<D0, 1, true, true>(h, P, n_he == 1 ? "fused[HE,Hp3,Hp2]{2,3}" : "fused[HE,HE,Hp3,Hp2]{2,3}");
    });
    h.pg->sorted2d = false;
    allreduce_sum(h.j(1), 2 * h.nd);
    m2d_e_from_j(*h.maxwell, h.e(2), h.j(2), 3);
    m2d_e_from_j(*h.maxwell, h.e(1), h.j(1), 2);
}

// tail of a Strang step: Hp3(dt/2) right after the sorting Hp2 (same b, so the table of sorted_head is still valid)
static void sorted_hp3(Splitting2D &h, double dt)
{
    zero_grid(h.j(2), h.nd);
    GP_DISPATCH_D0(h.maxwell->s_deg_0, {
        FastParams<D0> P{};
        P.dtqm3 = dt * h.pg->q_over_m;
        P.wscale3 = h.pg->charge * h.pg->common_weight * dt;
        launch_sorted<D0, 0, true, false>(h, P, "operatorHp3{2,3} sorted");
    });
    allreduce_sum(h.j(2), h.nd);
    m2d_e_from_j(*h.maxwell, h.e(2), h.j(2), 3);
}

// strang_splitting! with the point-wise operators fused (same trajectory up to rounding):
//   first step      HB ; [b -= dt/2 curl e] ; fused{HE,Hp3} ; Hp2 Hp1 Hp2 Hp3
//   between steps   eT = e ; b -= dt/2 curl e ; HB ; HB ; b -= dt/2 curl e ; fused{HE(eT),HE,Hp3} ; Hp2 Hp1 Hp2 Hp3
//   after the last  HE ; HB
// (compute_b_from_e! of a HE only reads e, which the kick reads too, and writes b, which the kick does not read)
static void strang2d_fused(Splitting2D &h, double dt, int64_t steps)
{
    const Maxwell2D &m = *h.maxwell;
    double *b[3] = {h.b(0), h.b(1), h.b(2)};
    const double *e[3] = {h.e(0), h.e(1), h.e(2)};
    // sort_interval == 1: the cell sort rides in the last push of every step (sorting_hp2); a stand-alone sort is only
    // needed when somebody else has moved the particles since
    // grids beyond the cell sort's shared-memory histogram (51200 cells) run un-sorted: the tile passes then take their
    // general path more often, but nothing fails
    const bool sortable = h.nd * sizeof(int) <= 200 * 1024;
    const bool ride = sortable && h.sort_interval == 1 && h.pg->W == 1 && h.pg->n >= 2;
    // fuse level 2: the operators that start from the sorted order run on the register-resident fast path (k2_sorted)
    const bool fast = ride && h.fuse >= 2;
    // a trailing HE kick this splitting left pending in its previous call (fields already advanced, snapshot in eT)
    // rides in the first pass of this one; anybody else's is applied first
    ParticleGroup &pg = *h.pg;
    if (pg.pending || (pg.pending2d && pg.pending2d != &h)) pg_sync(pg);
    const bool pending = pg.pending2d == &h;
    pg.pending2d = nullptr;
    for (int64_t s = 0; s < steps; ++s) {
        if (ride) {
            if (!h.pg->sorted2d) pg_sort_2d(*h.pg, *h.maxwell);
        } else if (sortable && h.sort_interval > 0 && h.steps_done % h.sort_interval == 0) {
            pg_sort_2d(*h.pg, *h.maxwell);
        }
        if (s == 0) {
            hs2d_operator(h, GEMPIC_OP_HB, 0.5 * dt);
            m2d_b_from_e(m, b, 0.5 * dt, e);
            if (fast) sorted_head(h, dt, pending ? 2 : 1, pending ? h.pending_dt : dt);
            else fused_he_hp3(h, dt, pending ? 2 : 1, pending ? h.pending_dt : dt);
        } else {
            GP_CUDA(cudaMemcpyAsync(h.eT(0), h.e(0), 3 * h.nd * sizeof(double), cudaMemcpyDeviceToDevice, ctx().stream));
            m2d_b_from_e(m, b, 0.5 * dt, e);              // trailing HE of step s-1, field part
            hs2d_operator(h, GEMPIC_OP_HB, 0.5 * dt);     // trailing HB of step s-1
            hs2d_operator(h, GEMPIC_OP_HB, 0.5 * dt);     // leading HB of step s
            m2d_b_from_e(m, b, 0.5 * dt, e);              // leading HE of step s, field part
            if (fast) sorted_head(h, dt, 2, dt);
            else fused_he_hp3(h, dt, 2, dt);
        }
        if (!fast) hs2d_operator(h, GEMPIC_OP_HP2, 0.5 * dt);
        if (ride) {   // the Hp1 pass also takes the histogram of the sort keys after the Hp2 push that follows
            op2_Hp1_with_hist(h, dt, 0.5 * dt);
            sorting_hp2(h, 0.5 * dt, true);
        } else {
            hs2d_operator(h, GEMPIC_OP_HP1, dt);
            hs2d_operator(h, GEMPIC_OP_HP2, 0.5 * dt);
        }
        if (fast) sorted_hp3(h, 0.5 * dt);
        else hs2d_operator(h, GEMPIC_OP_HP3, 0.5 * dt);
        h.steps_done++;
    }
    // trailing HE + HB: the fields are advanced now; the particle kick, which only reads the snapshot eT, is deferred to
    // the next call of this splitting or applied by pg_sync() as soon as anybody else touches the particles
    GP_CUDA(cudaMemcpyAsync(h.eT(0), h.e(0), 3 * h.nd * sizeof(double), cudaMemcpyDeviceToDevice, ctx().stream));
    m2d_b_from_e(m, b, 0.5 * dt, e);
    hs2d_operator(h, GEMPIC_OP_HB, 0.5 * dt);
    h.pending_dt = dt;
    pg.pending2d = &h;
}

void hs2d_apply_pending(ParticleGroup &pg)
{
    Splitting2D *h = pg.pending2d;
    if (!h) return;
    pg.pending2d = nullptr;
    op2_HE_particles(*h, 0.5 * h->pending_dt, h->eT(0), h->eT(1), h->eT(2));
}

void hs2d_strang(Splitting2D &h, double dt, int64_t steps)
{
    if (steps <= 0) return;
    if (h.fuse) {
        strang2d_fused(h, dt, steps);
        return;
    }
    pg_sync(*h.pg);
    for (int64_t s = 0; s < steps; ++s) {
        if (h.nd * sizeof(int) <= 200 * 1024 && h.sort_interval > 0 && h.steps_done % h.sort_interval == 0) pg_sort_2d(*h.pg, *h.maxwell);
        hs2d_operator(h, GEMPIC_OP_HB, 0.5 * dt);
        hs2d_operator(h, GEMPIC_OP_HE, 0.5 * dt);
        hs2d_operator(h, GEMPIC_OP_HP3, 0.5 * dt);
        hs2d_operator(h, GEMPIC_OP_HP2, 0.5 * dt);
        hs2d_operator(h, GEMPIC_OP_HP1, dt);
        hs2d_operator(h, GEMPIC_OP_HP2, 0.5 * dt);
        hs2d_operator(h, GEMPIC_OP_HP3, 0.5 * dt);
        hs2d_operator(h, GEMPIC_OP_HE, 0.5 * dt);
        hs2d_operator(h, GEMPIC_OP_HB, 0.5 * dt);
        h.steps_done++;
    }
}

Splitting2D::~Splitting2D()
{
    if (pg && pg->pending2d == this) pg->pending2d = nullptr;
    if (maxwell) release(maxwell);
    if (pg) release(pg);
}

void hs2d_moments(Splitting2D &h, double *out4_dev)
{
    Context &c = ctx();
    const int grid = c.sm_count * 4;
    double *part = h.scratch.ensure((size_t)grid * 4);
    k2_moments<<<grid, 256, 0, c.stream>>>(rows2(*h.pg), h.pg->n, part);
    GP_CUDA(cudaGetLastError());
    k_reduce_partials<<<1, 128, 0, c.stream>>>(part, grid, 4, out4_dev);
    GP_CUDA(cudaGetLastError());
    count_launch(2);
    allreduce_sum(out4_dev, 4);
}

}  // namespace gempic

// =============================== C ABI ==========================================================
using namespace gempic;

extern "C" {

int gempic_hs2d_create(gempic_handle maxwell2d, gempic_handle pgh, gempic_handle *out)
{
    GP_API_BEGIN
    require_init();
    GP_REQUIRE(out, GEMPIC_EINVAL, "null output handle");
    Maxwell2D *mx = get<Maxwell2D>(maxwell2d, "TwoDMaxwell");
    ParticleGroup *pg = get_pg(pgh);
    GP_REQUIRE(pg->D == 2 && pg->V == 3, GEMPIC_EASSERT, "dims == (2, 3) (hamiltonian_splitting.jl:47)");
    GP_REQUIRE(pg->W >= 1, GEMPIC_EASSERT, "particle group needs a weight row");
    GP_REQUIRE(mx->s_deg_0 >= 1, GEMPIC_EINVAL, "degree");
    GP_REQUIRE(mx->nx >= 2 * kR2 + 2 && mx->ny >= 2 * kR2 + 2, GEMPIC_EINVAL,
               "HamiltonianSplitting{2,3} needs at least %d cells per direction", 2 * kR2 + 2);
    auto h = std::make_unique<Splitting2D>();
    h->maxwell = mx; h->pg = pg;
    retain(mx); retain(pg);
    h->nd = (size_t)mx->nx * mx->ny;
    h->fields.alloc(13 * h->nd + 16);
    h->fields.zero(ctx().stream);
    *out = register_object(std::move(h));
    GP_API_END
}

int gempic_hs2d_destroy(gempic_handle hs)
{
    GP_API_BEGIN
    Splitting2D *h = get<Splitting2D>(hs, "HamiltonianSplitting{2,3}");
    if (h->pg->pending2d == h) pg_sync(*h->pg);   // a deferred HE kick must not be lost
    destroy(hs, Kind::Splitting2D, "HamiltonianSplitting{2,3}");
    GP_API_END
}

int gempic_hs2d_set_fields(gempic_handle hs, const double *e1, const double *e2, const double *e3, const double *b1,
                           const double *b2, const double *b3)
{
    GP_API_BEGIN
    require_init();
    Splitting2D *h = get<Splitting2D>(hs, "HamiltonianSplitting{2,3}");
    const double *src[6] = {e1, e2, e3, b1, b2, b3};
    for (int k = 0; k < 6; ++k) {
        GP_REQUIRE(src[k], GEMPIC_EINVAL, "null field buffer");
        GP_CUDA(cudaMemcpyAsync(h->fields.p + k * h->nd, src[k], h->nd * sizeof(double), cudaMemcpyHostToDevice, ctx().stream));
    }
    GP_CUDA(cudaStreamSynchronize(ctx().stream));
    GP_API_END
}

/* any pointer may be NULL (skipped); j1, j2, j3 are the currents of the last Hp1, Hp2, Hp3 */
int gempic_hs2d_get_fields(gempic_handle hs, double *e1, double *e2, double *e3, double *b1, double *b2, double *b3,
                           double *j1, double *j2, double *j3)
{
    GP_API_BEGIN
    require_init();
    Splitting2D *h = get<Splitting2D>(hs, "HamiltonianSplitting{2,3}");
    double *dst[9] = {e1, e2, e3, b1, b2, b3, j1, j2, j3};
    for (int k = 0; k < 9; ++k)
        if (dst[k] && host_out_enabled())
            GP_CUDA(cudaMemcpyAsync(dst[k], h->fields.p + k * h->nd, h->nd * sizeof(double), cudaMemcpyDeviceToHost, ctx().stream));
    GP_CUDA(cudaStreamSynchronize(ctx().stream));
    GP_API_END
}

int gempic_hs2d_operator(gempic_handle hs, int op, double dt)
{
    GP_API_BEGIN
    require_init();
    Splitting2D *h = get<Splitting2D>(hs, "HamiltonianSplitting{2,3}");
    pg_sync(*h->pg);
    hs2d_operator(*h, op, dt);
    GP_API_END
}

int gempic_hs2d_strang_splitting(gempic_handle hs, double dt, int64_t number_steps)
{
    GP_API_BEGIN
    require_init();
    hs2d_strang(*get<Splitting2D>(hs, "HamiltonianSplitting{2,3}"), dt, number_steps);
    GP_API_END
}

/* drop-in form with HOST field buffers (the aliased e_dofs / b_dofs of the reference struct):
 * H2D(e, b) -> number_steps Strang steps -> D2H(e, b), synchronous */
int gempic_hs2d_strang_splitting_host(gempic_handle hs, double dt, int64_t number_steps, double *e1, double *e2, double *e3,
                                      double *b1, double *b2, double *b3)
{
    int rc = gempic_hs2d_set_fields(hs, e1, e2, e3, b1, b2, b3);
    if (rc) return rc;
    rc = gempic_hs2d_strang_splitting(hs, dt, number_steps);
    if (rc) return rc;
    return gempic_hs2d_get_fields(hs, e1, e2, e3, b1, b2, b3, nullptr, nullptr, nullptr);
}

int gempic_hs2d_operator_host(gempic_handle hs, int op, double dt, double *e1, double *e2, double *e3, double *b1,
                              double *b2, double *b3)
{
    int rc = gempic_hs2d_set_fields(hs, e1, e2, e3, b1, b2, b3);
    if (rc) return rc;
    rc = gempic_hs2d_operator(hs, op, dt);
    if (rc) return rc;
    return gempic_hs2d_get_fields(hs, e1, e2, e3, b1, b2, b3, nullptr, nullptr, nullptr);
}

/* cell-sort the particles every `interval` Strang steps (0: never; default 1) */
int gempic_hs2d_set_sort_interval(gempic_handle hs, int interval)
{
    GP_API_BEGIN
    Splitting2D *h = get<Splitting2D>(hs, "HamiltonianSplitting{2,3}");
    GP_REQUIRE(interval >= 0, GEMPIC_EINVAL, "interval");
    h->sort_interval = interval;
    GP_API_END
}

/* 2 (default): fused particle passes inside strang_splitting, the operators that start from the cell-sorted order on
 * the register-resident fast path (k2_sorted); 1: fused [HE,Hp3] tile pass and the cross-step HE fold only;
 * 0: one pass per operator */
int gempic_hs2d_set_fusion(gempic_handle hs, int fuse)
{
    GP_API_BEGIN
    GP_REQUIRE(fuse >= 0 && fuse <= 2, GEMPIC_EINVAL, "fuse must be 0, 1 or 2");
    get<Splitting2D>(hs, "HamiltonianSplitting{2,3}")->fuse = fuse;
    GP_API_END
}

/* rho[nx*ny] = add_charge! of all particles (degree p x p, get_charge weights), all-reduced over ranks */
int gempic_hs2d_charge_density(gempic_handle hs, double *rho)
{
    GP_API_BEGIN
    require_init();
    Splitting2D *h = get<Splitting2D>(hs, "HamiltonianSplitting{2,3}");
    GP_REQUIRE(rho, GEMPIC_EINVAL, "null buffer");
    pg_sync(*h->pg);
    double *dr = h->fields.p + 9 * h->nd;
    hs2d_charge(*h, dr);
    d2h(rho, dr, h->nd);
    GP_API_END
}

/* out[4] = sum_p w |v|^2, sum_p w v1, sum_p w v2, sum_p w v3 (all ranks) */
int gempic_hs2d_moments(gempic_handle hs, double *out4)
{
    GP_API_BEGIN
    require_init();
    Splitting2D *h = get<Splitting2D>(hs, "HamiltonianSplitting{2,3}");
    GP_REQUIRE(out4, GEMPIC_EINVAL, "null buffer");
    pg_sync(*h->pg);
    double *d = h->fields.p + 13 * h->nd;
    hs2d_moments(*h, d);
    d2h(out4, d, 4);
    GP_API_END
}

int gempic_pg_sort2d(gempic_handle pgh, gempic_handle maxwell2d)
{
    GP_API_BEGIN
    require_init();
    ParticleGroup *pg = get_pg(pgh);
    GP_REQUIRE(pg->D == 2, GEMPIC_EINVAL, "2D cell sort needs a D = 2 particle group");
    pg_sort_2d(*pg, *get<Maxwell2D>(maxwell2d, "TwoDMaxwell"));
    GP_API_END
}

}  // extern "C"
