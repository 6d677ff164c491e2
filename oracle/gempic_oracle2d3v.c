/* CPU ORACLE (part 3): particle loops of a 2d3v Hamiltonian splitting on ParticleGroup{2,3}.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/oracle.py).  PARITY UNPINNED at the integrator level: the
 * reference ships no 2d3v integrator (src/hamiltonian_splitting_2d3v.jl is an empty file),
 * so there is nothing in GEMPIC.jl to pin these loops to.  They are the 2D extension of the reference's
 * 1d2v operators, written in the reference's own style so that each piece can be checked against the
 * pinned 1D oracle (tests/test_oracle_2d3v.py: x2-independent data reproduces the 1d2v operators) and
 * against the invariants of the scheme (discrete Gauss law, energy):
 *   operatorHE   v += dt q/m E(x)                               like hamiltonian_splitting_1d2v.jl:198-215
 *   operatorHp1  x1 += dt v1, line-integrated j1, v2 -= q/m int B3 dx1, v3 += q/m int B2 dx1
 *                                                                like :48-82 + pmc1d.jl:296-425
 *   operatorHp2  x2 += dt v2, line-integrated j2, v1 += q/m int B3 dx2, v3 -= q/m int B1 dx2
 *   operatorHp3  v1 -= dt q/m v3 B2(x), v2 += dt q/m v3 B1(x), j3 += w v3 N(x)     like :141-167
 * (Kraus, Kormann, Morrison, Sonnendruecker 2017, "GEMPIC", the paper cited in CITATION.bib).
 *
 * Discrete spaces (test/test_maxwell_2d_fem.jl:54-56,88-90), p = degree:
 *   E1 in S^{p-1} x S^p,  E2 in S^p x S^{p-1},  E3 in S^p x S^p          (1-form)
 *   B1 in S^p x S^{p-1},  B2 in S^{p-1} x S^p,  B3 in S^{p-1} x S^{p-1}  (2-form)
 * Basis: uniform_bsplines_eval_basis! (src/low_level_bsplines.jl:63-80); a particle in cell c with
 * offset t touches dofs (c-d+k) mod n, k = 0..d (src/particle_mesh_coupling_1d.jl:8).  Cells are
 * floor((x-xmin)/dx) (the 2D coupling's ceil convention, pmc2d.jl:57-58, gives the same dofs except
 * exactly on a grid line).  Line integrals use the reference's Gauss-Legendre rule per cell segment
 * (update_jv!, pmc1d.jl:385-425) -- deliberately NOT the primitive form the CUDA kernels use.
 * Dofs: flat nx*ny, x fastest.  Particles: column-major 6 x N records (x1,x2,v1,v2,v3,w).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

typedef struct {
    double xmin[2], L[2], d[2];
    int n[2];
    int deg;
} orc2_mesh;

static void basis(int degree, double offset, double *bspl)   /* low_level_bsplines.jl:63-80 */
{
    bspl[0] = 1.0;
    for (int j = 1; j <= degree; ++j) {
        double xx = -offset, saved = 0.0;
        const double jr = (double)j, inv_j = 1.0 / jr;
        for (int r = 0; r < j; ++r) {
            xx = xx + 1.0;
            const double temp = bspl[r] * inv_j;
            bspl[r] = saved + xx * temp;
            saved = (jr - xx) * temp;
        }
        bspl[j] = saved;
    }
}

static inline long imod(long a, long n) { long r = a % n; return r < 0 ? r + n : r; }

static inline void locate(const orc2_mesh *m, int ax, double x, long *cell, double *t)
{
    const double xi = (x - m->xmin[ax]) / m->d[ax];
    const double c = floor(xi);
    *cell = (long)c;
    *t = xi - c;
}

/* sum_{a,b} f[(cx-dx+a) mod nx, (cy-dy+b) mod ny] Nx_a Ny_b */
static double eval2(const orc2_mesh *m, const double *f, int dgx, int dgy, long cx, const double *bx, long cy, const double *by)
{
    double v = 0.0;
    for (int b = 0; b <= dgy; ++b) {
        const long iy = imod(cy - dgy + b, m->n[1]);
        for (int a = 0; a <= dgx; ++a) {
            const long ix = imod(cx - dgx + a, m->n[0]);
            v += f[ix + iy * m->n[0]] * bx[a] * by[b];
        }
    }
    return v;
}

void orc2_he(const orc2_mesh *m, double *pa, int64_t n, double dtqm, const double *e1, const double *e2, const double *e3)
{
    const int p = m->deg;
    for (int64_t i = 0; i < n; ++i) {
        double *r = pa + 6 * i;
        long cx, cy;
        double tx, ty, bx0[4], bx1[4], by0[4], by1[4];
        locate(m, 0, r[0], &cx, &tx);
        locate(m, 1, r[1], &cy, &ty);
        basis(p, tx, bx0); basis(p - 1, tx, bx1);
        basis(p, ty, by0); basis(p - 1, ty, by1);
        r[2] += dtqm * eval2(m, e1, p - 1, p, cx, bx1, cy, by0);
        r[3] += dtqm * eval2(m, e2, p, p - 1, cx, bx0, cy, by1);
        r[4] += dtqm * eval2(m, e3, p, p, cx, bx0, cy, by0);
    }
}

/* deposit wscale * w * N^p(x1) N^p(x2) (add_charge!, pmc2d.jl:85-105 with galerkin scaling 1) */
void orc2_charge(const orc2_mesh *m, const double *pa, int64_t n, double wscale, double *rho)
{
    const int p = m->deg;
    for (int64_t i = 0; i < n; ++i) {
        const double *r = pa + 6 * i;
        long cx, cy;
        double tx, ty, bx0[4], by0[4];
        locate(m, 0, r[0], &cx, &tx);
        locate(m, 1, r[1], &cy, &ty);
        basis(p, tx, bx0);
        basis(p, ty, by0);
        const double w = r[5] * wscale;
        for (int b = 0; b <= p; ++b)
            for (int a = 0; a <= p; ++a)
                rho[imod(cx - p + a, m->n[0]) + imod(cy - p + b, m->n[1]) * m->n[0]] += w * bx0[a] * by0[b];
    }
}

void orc2_hp3(const orc2_mesh *m, double *pa, int64_t n, double dt, double qm, double wscale, const double *b1,
              const double *b2, double *j3)
{
    const int p = m->deg;
    for (int64_t i = 0; i < n; ++i) {
        double *r = pa + 6 * i;
        long cx, cy;
        double tx, ty, bx0[4], bx1[4], by0[4], by1[4];
        locate(m, 0, r[0], &cx, &tx);
        locate(m, 1, r[1], &cy, &ty);
        basis(p, tx, bx0); basis(p - 1, tx, bx1);
        basis(p, ty, by0); basis(p - 1, ty, by1);
        const double B1 = eval2(m, b1, p, p - 1, cx, bx0, cy, by1);
        const double B2 = eval2(m, b2, p - 1, p, cx, bx1, cy, by0);
        const double v3 = r[4];
        r[2] -= dt * qm * v3 * B2;
        r[3] += dt * qm * v3 * B1;
        const double w = r[5] * wscale * v3;
        for (int b = 0; b <= p; ++b)
            for (int a = 0; a <= p; ++a)
                j3[imod(cx - p + a, m->n[0]) + imod(cy - p + b, m->n[1]) * m->n[0]] += w * bx0[a] * by0[b];
    }
}

/* line-integral weights of the degree-dg splines over one in-cell segment [lower, upper] (cell units),
 * update_jv! (pmc1d.jl:399-414): s_k = sign * h * sum_q w_q c1 N_k(c1 x_q + c2) */
static void segment(int dg, double lower, double upper, double sign_h, double *s)
{
    const int nq = (dg + 2) / 2;
    const double qx[2][2] = {{0.0, 0.0}, {-0.57735026918962576451, 0.57735026918962576451}};
    const double qw[2][2] = {{2.0, 0.0}, {1.0, 1.0}};
    const double c1 = (upper - lower) * 0.5, c2 = (upper + lower) * 0.5;
    double val[4];
    for (int k = 0; k <= dg; ++k) s[k] = 0.0;
    for (int q = 0; q < nq; ++q) {
        basis(dg, c1 * qx[nq - 1][q] + c2, val);
        for (int k = 0; k <= dg; ++k) s[k] += val[k] * qw[nq - 1][q] * c1;
    }
    for (int k = 0; k <= dg; ++k) s[k] *= sign_h;
}

/* dir = 0: operatorHp1 (push x1; bz = B3, bo = B2, j = j1)   dir = 1: operatorHp2 (push x2; bz = B3, bo = B1, j = j2) */
void orc2_hp12(const orc2_mesh *m, double *pa, int64_t n, int dir, double dt, double qm, double wscale, const double *bz,
               const double *bo, double *j)
{
    const int p = m->deg, d1 = p - 1, o = 1 - dir;
    const int nx = m->n[0];
    for (int64_t i = 0; i < n; ++i) {
        double *r = pa + 6 * i;
        const double x_old = r[dir], x_new = x_old + dt * r[2 + dir];
        long co, cn, ct;
        double to, tn, tt, bt0[4], bt1[4];
        locate(m, dir, x_old, &co, &to);
        locate(m, dir, x_new, &cn, &tn);
        locate(m, o, r[o], &ct, &tt);       /* transverse coordinate */
        basis(p, tt, bt0);
        basis(d1, tt, bt1);
        const double w = r[5] * wscale;
        double sum_z = 0.0, sum_o = 0.0;
        /* cells from co to cn in the direction of motion (pmc1d.jl:316-373) */
        const long step = cn >= co ? 1 : -1;
        for (long c = co;; c += step) {
            double lower, upper, sgn;
            if (co == cn) { lower = to; upper = tn; sgn = 1.0; }
            else if (step > 0) { lower = (c == co) ? to : 0.0; upper = (c == cn) ? tn : 1.0; sgn = 1.0; }
            else { lower = (c == cn) ? tn : 0.0; upper = (c == co) ? to : 1.0; sgn = -1.0; }
            double s[4];
            segment(d1, lower, upper, sgn * m->d[dir], s);
            for (int k = 0; k <= d1; ++k) {
                const long ia = imod(c - d1 + k, m->n[dir]);
                /* j: degree p transverse; bz = B3: degree p-1 transverse; bo: degree p transverse */
                for (int b = 0; b <= p; ++b) {
                    const long it = imod(ct - p + b, m->n[o]);
                    const long idx = dir == 0 ? ia + it * nx : it + ia * nx;
                    j[idx] += w * s[k] * bt0[b];
                    sum_o += bo[idx] * s[k] * bt0[b];
                }
                for (int b = 0; b <= d1; ++b) {
                    const long it = imod(ct - d1 + b, m->n[o]);
                    const long idx = dir == 0 ? ia + it * nx : it + ia * nx;
                    sum_z += bz[idx] * s[k] * bt1[b];
                }
            }
            if (c == cn) break;
        }
        if (dir == 0) {            /* v2 -= q/m int B3 dx1 ; v3 += q/m int B2 dx1 */
            r[3] -= qm * sum_z;
            r[4] += qm * sum_o;
        } else {                   /* v1 += q/m int B3 dx2 ; v3 -= q/m int B1 dx2 */
            r[2] += qm * sum_z;
            r[4] -= qm * sum_o;
        }
        /* periodic wrap into [xmin, xmin + L): whole periods are added / subtracted one at a time */
        double xw = x_new;
        const double lo = m->xmin[dir], hi = m->xmin[dir] + m->L[dir];
        while (xw < lo) xw += m->L[dir];
        while (xw >= hi) xw -= m->L[dir];
        r[dir] = xw;
    }
}

/* sum_p w v^2, sum_p w v_k (diagnostics.jl:197-211 extended to three velocity components) */
void orc2_moments(const double *pa, int64_t n, double *out4)
{
    out4[0] = out4[1] = out4[2] = out4[3] = 0.0;
    for (int64_t i = 0; i < n; ++i) {
        const double *r = pa + 6 * i;
        out4[0] += r[5] * (r[2] * r[2] + r[3] * r[3] + r[4] * r[4]);
        out4[1] += r[5] * r[2];
        out4[2] += r[5] * r[3];
        out4[3] += r[5] * r[4];
    }
}
