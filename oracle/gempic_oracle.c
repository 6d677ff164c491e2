/*
 * gempic_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C, fp64, serial restatement of the arithmetic of GEMPIC.jl's
 * particle-mesh hot path.  It exists only so that tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs can check and time the
 * reference algorithm.  Nothing under gempic.jl_b200/ may import, link or call it.
 *
 * The reference is 100% Julia and Julia is not installed in this image, so the
 * reference itself cannot be compiled into oracle/_ref/ ("unbuildable": needs the
 * Julia runtime).  Parity of this restatement is PINNED by the reference's own
 * golden vectors (tests/golden/*.json, transcribed from /root/reference/test):
 * see tests/test_oracle_golden.py.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference).  Summation order is the reference's serial particle order.
 * The particle array uses the reference's own layout: column-major
 * (D+V+W) x N, i.e. one contiguous record [x.., v.., w..] per particle
 * (src/particle_group.jl:15-46).
 *
 * Third-party arithmetic used by the reference and restated here:
 *   FFTW.jl "1"  (r2r R2HC/HC2R, src/maxwell_1d_fem.jl:101-102) -> plain O(n^2) DFT in
 *                the same half-complex layout; any correct DFT reproduces the goldens.
 *   FastGaussQuadrature "0.3-1" gausslegendre(n) (src/particle_mesh_coupling_1d.jl:72)
 *                -> Newton iteration on Legendre polynomials (orc_gausslegendre).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_MAXDEG 5
#define ORC_PI 3.14159265358979323846264338327950288

/* ------------------------------------------------------------------------- */
/* src/low_level_bsplines.jl:63-80  uniform_bsplines_eval_basis!              */
/* ------------------------------------------------------------------------- */
void orc_bsplines_eval_basis(int degree, double offset, double *bspl)
{
    bspl[0] = 1.0;
    for (int j = 1; j <= degree; ++j) {
        double xx = -offset;
        double j_real = (double)j;
        double inv_j = 1.0 / j_real;
        double saved = 0.0;
        for (int r = 0; r < j; ++r) {
            xx = xx + 1.0;
            double temp = bspl[r] * inv_j;
            bspl[r] = saved + xx * temp;
            saved = (j_real - xx) * temp;
        }
        bspl[j] = saved;
    }
}

/* gausslegendre(n): nodes ascending on [-1,1] and weights (FastGaussQuadrature). */
void orc_gausslegendre(int n, double *x, double *w)
{
    for (int i = 0; i < n; ++i) {
        long double z = cosl((long double)ORC_PI * (i + 0.75L) / (n + 0.5L));
        long double pp = 1.0L;
        for (int it = 0; it < 100; ++it) {
            long double p1 = 1.0L, p2 = 0.0L;
            for (int j = 1; j <= n; ++j) {
                long double p3 = p2;
                p2 = p1;
                p1 = ((2.0L * j - 1.0L) * z * p2 - (j - 1.0L) * p3) / j;
            }
            pp = n * (z * p1 - p2) / (z * z - 1.0L);
            long double z1 = z;
            z = z1 - p1 / pp;
            if (fabsl(z - z1) < 1e-19L) break;
        }
        /* descending roots -> ascending order */
        x[n - 1 - i] = (double)z;
        w[n - 1 - i] = (double)(2.0L / ((1.0L - z * z) * pp * pp));
    }
    if (n % 2 == 1) x[n / 2] = 0.0;
}

/* ------------------------------------------------------------------------- */
/* ParticleMeshCoupling1D  (src/particle_mesh_coupling_1d.jl:26-95)           */
/* ------------------------------------------------------------------------- */
typedef struct {
    double xmin, Lx, delta_x, scaling;
    int n_grid, degree, n_span, n_quad;
    double quad_x[4], quad_w[4];
} orc_pmc1d;

/* smoothing: 0 = :collocation, 1 = :galerkin ; returns -1 for anything else
 * (ArgumentError at src/particle_mesh_coupling_1d.jl:61) */
int orc_pmc1d_init(orc_pmc1d *p, double xmin, double xmax, int n_grid, int degree, int smoothing)
{
    if (degree < 0 || degree > ORC_MAXDEG) return -1;
    p->xmin = xmin;
    p->Lx = xmax - xmin;
    p->delta_x = (xmax - xmin) / n_grid;          /* :53 */
    p->n_grid = n_grid;
    p->degree = degree;
    p->n_span = degree + 1;
    if (smoothing == 0) p->scaling = 1.0 / p->delta_x;   /* :56-57 */
    else if (smoothing == 1) p->scaling = 1.0;           /* :58-59 */
    else return -1;
    p->n_quad = (degree + 2) / 2;                        /* :67 */
    orc_gausslegendre(p->n_quad, p->quad_x, p->quad_w);  /* :72 */
    return 0;
}

static inline long orc_mod(long a, long n) { long r = a % n; return r < 0 ? r + n : r; }

/* src/particle_mesh_coupling_1d.jl:261-280  add_charge! */
void orc_pmc1d_add_charge(const orc_pmc1d *p, double *rho, double position, double marker_charge)
{
    double bspl[ORC_MAXDEG + 1];
    double xi = (position - p->xmin) / p->delta_x;
    long index = (long)xi;                   /* trunc(Int, xi)  :268 */
    xi = xi - (double)index;
    index = index - p->degree;
    orc_bsplines_eval_basis(p->degree, xi, bspl);
    for (int i = 1; i <= p->n_span; ++i) {
        long idx = orc_mod(index + i - 1, p->n_grid);    /* mod1(index+i, nx) - 1 */
        rho[idx] += marker_charge * bspl[i - 1] * p->scaling;
    }
}

/* src/particle_mesh_coupling_1d.jl:438-453  evaluate */
double orc_pmc1d_evaluate(const orc_pmc1d *p, double position, const double *field)
{
    double bspl[ORC_MAXDEG + 1];
    double xi = (position - p->xmin) / p->delta_x;
    long index = (long)xi;
    double dxi = xi - (double)index;
    index = index - p->degree;
    orc_bsplines_eval_basis(p->degree, dxi, bspl);
    double value = 0.0;
    for (int i = 1; i <= p->n_span; ++i) {
        long idx = orc_mod(index + i - 1, p->n_grid);
        value += field[idx] * bspl[i - 1];
    }
    return value;
}

/* src/particle_mesh_coupling_1d.jl:385-425  update_jv!  (bfield == NULL gives the
 * 1d1v variant :538-580, which leaves vi untouched) */
static double orc_update_jv(const orc_pmc1d *p, double *j_dofs, double lower, double upper,
                            long index, double marker_charge, double qoverm, double sign,
                            double vi, const double *bfield)
{
    double val[ORC_MAXDEG + 1], more[ORC_MAXDEG + 1];
    int n_cells = p->n_grid;
    double c1 = 0.5 * (upper - lower);
    double c2 = 0.5 * (upper + lower);
    orc_bsplines_eval_basis(p->degree, c1 * p->quad_x[0] + c2, val);
    double f = p->quad_w[0] * c1;
    for (int k = 0; k < p->n_span; ++k) val[k] *= f;
    for (int q = 1; q < p->n_quad; ++q) {
        orc_bsplines_eval_basis(p->degree, c1 * p->quad_x[q] + c2, more);
        for (int k = 0; k < p->n_span; ++k) val[k] += more[k] * p->quad_w[q] * c1;
    }
    double s = sign * p->delta_x;
    for (int k = 0; k < p->n_span; ++k) val[k] *= s;
    int ind = 0;
    for (long g = index - p->degree; g <= index; ++g) {
        long i_mod = orc_mod(g, n_cells);
        j_dofs[i_mod] += marker_charge * val[ind] * p->scaling;
        if (bfield) vi = vi - qoverm * val[ind] * bfield[i_mod];
        ind++;
    }
    return vi;
}

/* src/particle_mesh_coupling_1d.jl:296-376  add_current_update_v! (with B).
 * new_floor != 0 selects the 1d1v variant (:471-529) which uses floor() for the new
 * index (:487) and no B (pass bfield = NULL). */
static double orc_add_current_impl(const orc_pmc1d *p, double *j_dofs, double position_old,
                                   double position_new, double marker_charge, double qoverm,
                                   const double *bfield, double vi, int new_floor)
{
    double xi = (position_old - p->xmin) / p->delta_x;
    long index_old = (long)xi;                            /* trunc :307 / :481 */
    double r_old = xi - (double)index_old;
    xi = (position_new - p->xmin) / p->delta_x;
    long index_new = new_floor ? (long)floor(xi) : (long)xi;   /* :313 / :487 */
    double r_new = xi - (double)index_new;

    if (index_old == index_new) {
        if (r_old < r_new)
            vi = orc_update_jv(p, j_dofs, r_old, r_new, index_old, marker_charge, qoverm, 1.0, vi, bfield);
        else
            vi = orc_update_jv(p, j_dofs, r_new, r_old, index_old, marker_charge, qoverm, -1.0, vi, bfield);
    } else if (index_old < index_new) {
        vi = orc_update_jv(p, j_dofs, r_old, 1.0, index_old, marker_charge, qoverm, 1.0, vi, bfield);
        vi = orc_update_jv(p, j_dofs, 0.0, r_new, index_new, marker_charge, qoverm, 1.0, vi, bfield);
        for (long ind = index_old + 1; ind <= index_new - 1; ++ind)
            vi = orc_update_jv(p, j_dofs, 0.0, 1.0, ind, marker_charge, qoverm, 1.0, vi, bfield);
    } else {
        vi = orc_update_jv(p, j_dofs, r_new, 1.0, index_new, marker_charge, qoverm, -1.0, vi, bfield);
        vi = orc_update_jv(p, j_dofs, 0.0, r_old, index_old, marker_charge, qoverm, -1.0, vi, bfield);
        for (long ind = index_new + 1; ind <= index_old - 1; ++ind)
            vi = orc_update_jv(p, j_dofs, 0.0, 1.0, ind, marker_charge, qoverm, -1.0, vi, bfield);
    }
    return vi;
}

double orc_pmc1d_add_current_update_v(const orc_pmc1d *p, double *j_dofs, double position_old,
                                      double position_new, double marker_charge, double qoverm,
                                      const double *bfield, double vi)
{
    return orc_add_current_impl(p, j_dofs, position_old, position_new, marker_charge, qoverm, bfield, vi, 0);
}

/* src/particle_mesh_coupling_1d.jl:471-529 (1d1v, no B) */
double orc_pmc1d_add_current_1d1v(const orc_pmc1d *p, double *j_dofs, double position_old,
                                  double position_new, double marker_charge, double qoverm, double vi)
{
    return orc_add_current_impl(p, j_dofs, position_old, position_new, marker_charge, qoverm, NULL, vi, 1);
}

/* Julia float mod(x, y), y > 0 (base/float.jl): rem, then shift into [0,y) */
static inline double orc_fmod_julia(double x, double y)
{
    double r = fmod(x, y);
    if (r == 0.0) return copysign(r, y);
    if ((r > 0.0) != (y > 0.0)) return r + y;
    return r;
}
double orc_mod_julia(double x, double y) { return orc_fmod_julia(x, y); }

/* ------------------------------------------------------------------------- */
/* Maxwell1DFEM  (src/maxwell_1d_fem.jl:29-177)                               */
/* ------------------------------------------------------------------------- */
typedef struct {
    double xmin, Lx, delta_x;
    int n_dofs, s_deg_0, s_deg_1;
    double mass_0[4], mass_1[4];
    double *eig_mass0, *eig_mass1, *eig_weak_ampere, *eig_weak_poisson;
    double *work, *wsave, *eigvals;
    double *cos_tab, *sin_tab; /* cos/sin(2 pi m / n), m = 0..n-1 */
} orc_maxwell1d;

void orc_maxwell1d_free(orc_maxwell1d *m)
{
    free(m->eig_mass0); free(m->eig_mass1); free(m->eig_weak_ampere); free(m->eig_weak_poisson);
    free(m->work); free(m->wsave); free(m->eigvals); free(m->cos_tab); free(m->sin_tab);
    memset(m, 0, sizeof(*m));
}

/* src/maxwell_1d_fem.jl:49-177 constructor */
int orc_maxwell1d_init(orc_maxwell1d *m, double xmin, double xmax, int n_dofs, int degree)
{
    memset(m, 0, sizeof(*m));
    if (degree < 1 || degree > 3) return -1;   /* ArgumentError :91 */
    if (n_dofs % 2 != 0 || n_dofs < 2) return -2;
    m->xmin = xmin;
    m->n_dofs = n_dofs;
    m->Lx = xmax - xmin;
    m->delta_x = m->Lx / n_dofs;
    m->s_deg_0 = degree;
    m->s_deg_1 = degree - 1;
    double *mass_0 = m->mass_0, *mass_1 = m->mass_1;
    if (degree == 1) {
        mass_0[0] = 4.0 / 6.0; mass_0[1] = 1.0 / 6.0;
        mass_1[0] = 1.0;
    } else if (degree == 2) {
        mass_0[0] = 66.0 / 120.0; mass_0[1] = 26.0 / 120.0; mass_0[2] = 1.0 / 120.0;
        mass_1[0] = 4.0 / 6.0; mass_1[1] = 1.0 / 6.0;
    } else {
        mass_0[0] = 2416.0 / 5040.0; mass_0[1] = 1191.0 / 5040.0;
        mass_0[2] = 120.0 / 5040.0;  mass_0[3] = 1.0 / 5040.0;
        mass_1[0] = 66.0 / 120.0; mass_1[1] = 26.0 / 120.0; mass_1[2] = 1.0 / 120.0;
    }
    size_t nb = sizeof(double) * (size_t)n_dofs;
    m->eig_mass0 = calloc(1, nb); m->eig_mass1 = calloc(1, nb);
    m->eig_weak_ampere = calloc(1, nb); m->eig_weak_poisson = calloc(1, nb);
    m->work = calloc(1, nb); m->wsave = calloc(1, nb); m->eigvals = calloc(1, nb);
    m->cos_tab = calloc(1, nb); m->sin_tab = calloc(1, nb);
    for (int k = 0; k < n_dofs; ++k) {
        m->cos_tab[k] = (double)cosl(2.0L * (long double)ORC_PI * k / n_dofs);
        m->sin_tab[k] = (double)sinl(2.0L * (long double)ORC_PI * k / n_dofs);
    }
    int n = n_dofs, s = degree;
    m->eig_weak_ampere[0] = 0.0;
    m->eig_weak_poisson[0] = 0.0;
    m->eig_mass0[0] = 1.0;
    m->eig_mass1[0] = 1.0;
    for (int k = 1; k <= n / 2 - 1; ++k) {               /* :111-134 */
        double coef0 = mass_0[0], coef1 = mass_1[0];
        for (int j = 1; j <= s - 1; ++j) {
            double cos_mode = cos(2 * ORC_PI * j * k / n);
            coef0 = coef0 + 2 * mass_0[j] * cos_mode;
            coef1 = coef1 + 2 * mass_1[j] * cos_mode;
        }
        int j = s;
        coef0 = coef0 + 2 * mass_0[j] * cos(2 * ORC_PI * j * k / n);
        m->eig_mass0[k] = coef0;     m->eig_mass0[n - k] = 0.0;
        m->eig_mass1[k] = coef1;     m->eig_mass1[n - k] = 0.0;
        double cos_mode = cos(2 * ORC_PI * k / n);
        double sin_mode = sin(2 * ORC_PI * k / n);
        m->eig_weak_ampere[k] = (coef1 / coef0) * (1 - cos_mode);
        m->eig_weak_ampere[n - k] = -(coef1 / coef0) * sin_mode;
        m->eig_weak_poisson[k] = 1.0 / (coef1 * ((1 - cos_mode) * (1 - cos_mode) + sin_mode * sin_mode));
        m->eig_weak_poisson[n - k] = 0.0;
    }
    /* N/2 mode :137-153 */
    double coef0 = mass_0[0], coef1 = mass_1[0];
    for (int j = 1; j <= s - 1; ++j) {
        coef0 = coef0 + 2 * mass_0[j] * cos(ORC_PI * j);
        coef1 = coef1 + 2 * mass_1[j] * cos(ORC_PI * j);
    }
    coef0 = coef0 + 2 * mass_0[s] * cos(ORC_PI * s);
    m->eig_mass0[n / 2] = coef0;
    m->eig_mass1[n / 2] = coef1;
    m->eig_weak_ampere[n / 2] = 2.0 * (coef1 / coef0);
    m->eig_weak_poisson[n / 2] = 1.0 / (coef1 * 4.0);
    return 0;
}

/* accessors for ctypes (struct layout stays private to C) */
double *orc_maxwell1d_table(orc_maxwell1d *m, int which)
{
    switch (which) {
    case 0: return m->eig_mass0;
    case 1: return m->eig_mass1;
    case 2: return m->eig_weak_ampere;
    case 3: return m->eig_weak_poisson;
    case 4: return m->work;
    default: return NULL;
    }
}
double orc_maxwell1d_delta_x(const orc_maxwell1d *m) { return m->delta_x; }

/* src/maxwell_1d_fem.jl:222-240  solve_circulant!  -- result in m->work.
 * FFTW R2HC: X_k = sum_j x_j exp(-2 pi i jk/n), real parts at [k], imaginary at [n-k];
 * HC2R is the unnormalised inverse. */
void orc_maxwell1d_solve_circulant(orc_maxwell1d *m, const double *eigvals, const double *rhs)
{
    /* The two DFTs accumulate in long double so that the O(n^2) sums are as close to the
     * exact transform as FFTW's O(n log n) butterflies are (both within ~1 ulp of exact);
     * the half-complex products in between are the reference's fp64 expressions. */
    int n = m->n_dofs;
    double *ws = m->wsave, *wk = m->work;
    for (int k = 0; k <= n / 2; ++k) {
        long double re = 0.0L, im = 0.0L;
        for (int j = 0; j < n; ++j) {
            int t = (int)(((long)j * k) % n);
            re += (long double)rhs[j] * cosl(2.0L * (long double)ORC_PI * t / n);
            im -= (long double)rhs[j] * sinl(2.0L * (long double)ORC_PI * t / n);
        }
        ws[k] = (double)re;
        if (k > 0 && k < n / 2) ws[n - k] = (double)im;
    }
    ws[0] = ws[0] * eigvals[0];
    for (int k = 1; k < n / 2; ++k) {          /* julia k = 2 : n/2 */
        double re_p = ws[k] * eigvals[k] - ws[n - k] * eigvals[n - k];
        double im_p = ws[k] * eigvals[n - k] + ws[n - k] * eigvals[k];
        ws[k] = re_p;
        ws[n - k] = im_p;
    }
    ws[n / 2] = ws[n / 2] * eigvals[n / 2];
    for (int j = 0; j < n; ++j) {
        long double acc = (long double)ws[0] + ((j & 1) ? -(long double)ws[n / 2] : (long double)ws[n / 2]);
        for (int k = 1; k < n / 2; ++k) {
            int t = (int)(((long)j * k) % n);
            acc += 2.0L * ((long double)ws[k] * cosl(2.0L * (long double)ORC_PI * t / n)
                           - (long double)ws[n - k] * sinl(2.0L * (long double)ORC_PI * t / n));
        }
        wk[j] = (double)acc / n;
    }
}

/* src/maxwell_1d_fem.jl:244-255 */
void orc_maxwell1d_compute_e_from_rho(orc_maxwell1d *m, double *e, const double *rho)
{
    int n = m->n_dofs;
    orc_maxwell1d_solve_circulant(m, m->eig_weak_poisson, rho);
    for (int i = 1; i < n; ++i) e[i] = m->work[i - 1] - m->work[i];
    e[0] = m->work[n - 1] - m->work[0];
}

/* src/maxwell_1d_fem.jl:263-289 ; returns -1 on bad component (ArgumentError :283) */
int orc_maxwell1d_compute_e_from_j(orc_maxwell1d *m, double *e, const double *current, int component)
{
    int n = m->n_dofs;
    memset(m->eigvals, 0, sizeof(double) * (size_t)n);
    if (component == 1) {
        for (int i = 0; i <= n / 2; ++i) m->eigvals[i] = 1.0 / m->eig_mass1[i];
    } else if (component == 2) {
        for (int i = 0; i <= n / 2; ++i) m->eigvals[i] = 1.0 / m->eig_mass0[i];
    } else return -1;
    orc_maxwell1d_solve_circulant(m, m->eigvals, current);
    for (int i = 0; i < n; ++i) m->work[i] /= m->delta_x;
    for (int i = 0; i < n; ++i) e[i] -= m->work[i];
    return 0;
}

/* src/maxwell_1d_fem.jl:384-396 */
void orc_maxwell1d_compute_e_from_b(orc_maxwell1d *m, double *field_out, double delta_t, const double *field_in)
{
    double coef = delta_t / m->delta_x;
    orc_maxwell1d_solve_circulant(m, m->eig_weak_ampere, field_in);
    for (int i = 0; i < m->n_dofs; ++i) field_out[i] += coef * m->work[i];
}

/* src/maxwell_1d_fem.jl:407-420 */
void orc_maxwell1d_compute_b_from_e(const orc_maxwell1d *m, double *field_out, double delta_t, const double *field_in)
{
    int n = m->n_dofs;
    double coef = delta_t / m->delta_x;
    for (int i = 1; i < n; ++i) field_out[i] = field_out[i] + coef * (field_in[i - 1] - field_in[i]);
    field_out[0] = field_out[0] + coef * (field_in[n - 1] - field_in[0]);
}

/* src/maxwell_1d_fem.jl:461-475 */
double orc_maxwell1d_inner_product(orc_maxwell1d *m, const double *c1, const double *c2, int degree)
{
    if (degree == m->s_deg_0) orc_maxwell1d_solve_circulant(m, m->eig_mass0, c2);
    else if (degree == m->s_deg_1) orc_maxwell1d_solve_circulant(m, m->eig_mass1, c2);
    double r = 0.0;
    for (int i = 0; i < m->n_dofs; ++i) r += c1[i] * m->work[i];
    return r * m->delta_x;
}

/* src/maxwell_1d_fem.jl:299-313 */
double orc_maxwell1d_l2norm_squared(orc_maxwell1d *m, const double *c, int degree)
{
    return orc_maxwell1d_inner_product(m, c, c, degree);
}

typedef double (*orc_func1d)(double x, void *ctx);

/* src/maxwell_1d_fem.jl:188-220 compute_rhs_from_function! */
void orc_maxwell1d_compute_rhs_from_function(orc_maxwell1d *m, double *coefs, orc_func1d f, void *ctx, int degree)
{
    double x[ORC_MAXDEG + 1], w[ORC_MAXDEG + 1], bspl[ORC_MAXDEG + 1][ORC_MAXDEG + 1];
    int np = degree + 1;
    orc_gausslegendre(np, x, w);
    for (int k = 0; k < np; ++k) { x[k] = 0.5 * (x[k] + 1.0); w[k] = 0.5 * w[k]; }
    for (int k = 0; k < np; ++k) orc_bsplines_eval_basis(degree, x[k], bspl[k]);
    for (int i = 1; i <= m->n_dofs; ++i) {
        double coef = 0.0;
        for (int j = 1; j <= degree + 1; ++j)
            for (int k = 1; k <= degree + 1; ++k)
                coef = coef + w[k - 1] * f(m->delta_x * (x[k - 1] + i + j - 2), ctx) * bspl[k - 1][degree + 1 - j];
        coefs[i - 1] = coef * m->delta_x;
    }
}

/* src/maxwell_1d_fem.jl:347-374 l2projection! ; -1 on bad degree (:366) */
int orc_maxwell1d_l2projection(orc_maxwell1d *m, double *coefs, orc_func1d f, void *ctx, int degree)
{
    int n = m->n_dofs;
    double *eig = calloc((size_t)n, sizeof(double));
    orc_maxwell1d_compute_rhs_from_function(m, coefs, f, ctx, degree);
    if (degree == m->s_deg_0) {
        for (int i = 0; i <= n / 2; ++i) eig[i] = 1.0 / m->eig_mass0[i];
    } else if (degree == m->s_deg_0 - 1) {
        for (int i = 0; i <= n / 2; ++i) eig[i] = 1.0 / m->eig_mass1[i];
    } else { free(eig); return -1; }
    /* NOTE reference quirk: solve_circulant! leaves its result in self.work and the
     * reference then rescales coefs_dofs (the *rhs*) by 1/dx (:369-373) -- i.e. the
     * mass-matrix inverse is NOT applied to the returned coefficients. Restated as is. */
    orc_maxwell1d_solve_circulant(m, eig, coefs);
    for (int i = 0; i < n; ++i) coefs[i] = coefs[i] / m->delta_x;
    free(eig);
    return 0;
}

static double orc_cos_mode(double x, void *ctx) { double *p = (double *)ctx; return p[0] * cos(p[1] * x); }
static double orc_sin_mode(double x, void *ctx) { double *p = (double *)ctx; return p[0] * sin(p[1] * x); }
/* l2projection! of amp*cos(k x) (kind 0) or amp*sin(k x) (kind 1): the Weibel B3(0)
 * of test/test_vm_1d2v.jl:33-34,80-84 */
int orc_maxwell1d_l2projection_mode(orc_maxwell1d *m, double *coefs, int kind, double amp, double k, int degree)
{
    double prm[2] = {amp, k};
    return orc_maxwell1d_l2projection(m, coefs, kind == 0 ? orc_cos_mode : orc_sin_mode, prm, degree);
}

/* ------------------------------------------------------------------------- */
/* ParticleGroup view (src/particle_group.jl:15-78)                           */
/* ------------------------------------------------------------------------- */
typedef struct {
    double *array;       /* (D+V+W) x N column-major */
    int64_t n_particles;
    int D, V, W;
    double charge, mass, common_weight, q_over_m;
} orc_pg;

void orc_pg_init(orc_pg *pg, double *array, int64_t n, int D, int V, int W, double charge, double mass, double common_weight)
{
    pg->array = array; pg->n_particles = n; pg->D = D; pg->V = V; pg->W = W;
    pg->charge = charge; pg->mass = mass;
    pg->common_weight = (common_weight == 0.0) ? 1.0 / (double)n : common_weight;   /* :30-32 */
    pg->q_over_m = charge / mass;
}
#define PGA(pg, row, i) ((pg)->array[(size_t)(i) * (size_t)((pg)->D + (pg)->V + (pg)->W) + (row)])

/* ------------------------------------------------------------------------- */
/* HamiltonianSplitting{1,2}  (src/hamiltonian_splitting.jl:20-108,           */
/*                             src/hamiltonian_splitting_1d2v.jl)             */
/* ------------------------------------------------------------------------- */
typedef struct {
    orc_maxwell1d *maxwell;
    orc_pmc1d *ks0, *ks1;    /* kernel_smoother_0 (deg p), kernel_smoother_1 (deg p-1) */
    orc_pg *pg;
    double *e1, *e2, *b;     /* aliased caller arrays (hamiltonian_splitting.jl:80-81) */
    double *j1, *j2;         /* owned j_dofs */
    double Lx;
    int n_chunks;            /* = nthreads() of the reference (:61-66); 1 = serial */
} orc_hs;

int orc_hs_init(orc_hs *h, orc_maxwell1d *m, orc_pmc1d *ks0, orc_pmc1d *ks1, orc_pg *pg,
                double *e1, double *e2, double *b, int n_chunks)
{
    if (ks0->n_grid != ks1->n_grid) return -1;         /* @assert :49 */
    if (n_chunks < 1 || pg->n_particles % n_chunks != 0) return -2;   /* @assert :64 */
    h->maxwell = m; h->ks0 = ks0; h->ks1 = ks1; h->pg = pg;
    h->e1 = e1; h->e2 = e2; h->b = b;
    h->j1 = calloc((size_t)ks0->n_grid, sizeof(double));
    h->j2 = calloc((size_t)ks0->n_grid, sizeof(double));
    h->Lx = m->Lx;
    h->n_chunks = n_chunks;
    return 0;
}
void orc_hs_free(orc_hs *h) { free(h->j1); free(h->j2); h->j1 = h->j2 = NULL; }
double *orc_hs_j(orc_hs *h, int which) { return which == 1 ? h->j1 : h->j2; }

/* chunked deposit driver: per-chunk private buffer, buffers summed in chunk order
 * (reduce(+, fetch.(tasks)), hamiltonian_splitting_1d2v.jl:88,108,169) */
typedef void (*orc_chunk_fn)(orc_hs *h, double dt, int64_t lo, int64_t hi, double *buffer);

static void orc_run_chunks(orc_hs *h, double dt, orc_chunk_fn fn, double *out)
{
    int n = h->ks0->n_grid;
    int nc = h->n_chunks;
    int64_t np = h->pg->n_particles, len = np / nc;
    double *bufs = calloc((size_t)nc * (size_t)n, sizeof(double));
#pragma omp parallel for schedule(static, 1) if (nc > 1)
    for (int c = 0; c < nc; ++c) fn(h, dt, c * len, (c + 1) * len, bufs + (size_t)c * n);
    for (int i = 0; i < n; ++i) {
        double acc = bufs[i];
        for (int c = 1; c < nc; ++c) acc = acc + bufs[(size_t)c * n + i];
        out[i] = acc;
    }
    free(bufs);
}

/* hamiltonian_splitting_1d2v.jl:52-82 */
static void orc_hp1_chunk_j(orc_hs *h, double dt, int64_t lo, int64_t hi, double *buffer)
{
    orc_pg *pg = h->pg;
    for (int64_t i = lo; i < hi; ++i) {
        double x_old = PGA(pg, 0, i), v1_old = PGA(pg, 1, i), v2_old = PGA(pg, 2, i);
        double x_new = x_old + dt * v1_old;
        double wi = PGA(pg, 3, i);
        wi = wi * pg->charge;
        wi = wi * pg->common_weight;
        double v2_new = orc_pmc1d_add_current_update_v(h->ks1, buffer, x_old, x_new, wi, pg->q_over_m, h->b, v2_old);
        x_new = orc_fmod_julia(x_new, h->Lx);
        PGA(pg, 0, i) = x_new;
        PGA(pg, 2, i) = v2_new;
    }
}
/* hamiltonian_splitting_1d2v.jl:93-104 */
static void orc_hp1_chunk_rho(orc_hs *h, double dt, int64_t lo, int64_t hi, double *buffer)
{
    (void)dt;
    orc_pg *pg = h->pg;
    for (int64_t i = lo; i < hi; ++i) {
        double x = PGA(pg, 0, i), w = PGA(pg, 3, i);
        w = w * pg->charge;
        w = w * pg->common_weight;
        orc_pmc1d_add_charge(h->ks0, buffer, x, w);
    }
}
/* hamiltonian_splitting_1d2v.jl:41-112 */
void orc_hs_operatorHp1(orc_hs *h, double dt)
{
    int n = h->ks0->n_grid;
    memset(h->j1, 0, sizeof(double) * (size_t)n);
    memset(h->j2, 0, sizeof(double) * (size_t)n);
    orc_run_chunks(h, dt, orc_hp1_chunk_j, h->j1);
    orc_run_chunks(h, dt, orc_hp1_chunk_rho, h->j2);
    orc_maxwell1d_compute_e_from_j(h->maxwell, h->e1, h->j1, 1);
}

/* hamiltonian_splitting_1d2v.jl:144-165 */
static void orc_hp2_chunk(orc_hs *h, double dt, int64_t lo, int64_t hi, double *buffer)
{
    orc_pg *pg = h->pg;
    double qm = pg->q_over_m;
    for (int64_t i = lo; i < hi; ++i) {
        double x1 = PGA(pg, 0, i), v1 = PGA(pg, 1, i), v2 = PGA(pg, 2, i);
        double b = orc_pmc1d_evaluate(h->ks1, x1, h->b);
        v1 = v1 + dt * qm * v2 * b;
        PGA(pg, 1, i) = v1;
        double w = PGA(pg, 3, i);
        w = w * pg->charge;
        w = w * pg->common_weight;
        w = w * v2;
        orc_pmc1d_add_charge(h->ks0, buffer, x1, w);
    }
}
/* hamiltonian_splitting_1d2v.jl:129-176 */
void orc_hs_operatorHp2(orc_hs *h, double dt)
{
    int n = h->ks0->n_grid;
    memset(h->j1, 0, sizeof(double) * (size_t)n);
    memset(h->j2, 0, sizeof(double) * (size_t)n);
    orc_run_chunks(h, dt, orc_hp2_chunk, h->j2);
    for (int i = 0; i < n; ++i) h->j2[i] = h->j2[i] * dt;
    orc_maxwell1d_compute_e_from_j(h->maxwell, h->e2, h->j2, 2);
}

/* hamiltonian_splitting_1d2v.jl:191-219 */
void orc_hs_operatorHE(orc_hs *h, double dt)
{
    orc_pg *pg = h->pg;
    double qm = pg->q_over_m;
    int64_t np = pg->n_particles;
#pragma omp parallel for schedule(static) if (h->n_chunks > 1)
    for (int64_t i = 0; i < np; ++i) {
        double v_old1 = PGA(pg, 1, i), v_old2 = PGA(pg, 2, i);
        double xi = PGA(pg, 0, i);
        double e1 = orc_pmc1d_evaluate(h->ks1, xi, h->e1);
        double e2 = orc_pmc1d_evaluate(h->ks0, xi, h->e2);
        PGA(pg, 1, i) = v_old1 + dt * qm * e1;
        PGA(pg, 2, i) = v_old2 + dt * qm * e2;
    }
    orc_maxwell1d_compute_b_from_e(h->maxwell, h->b, dt, h->e2);
}

/* hamiltonian_splitting_1d2v.jl:234-236 */
void orc_hs_operatorHB(orc_hs *h, double dt)
{
    orc_maxwell1d_compute_e_from_b(h->maxwell, h->e2, dt, h->b);
}

/* hamiltonian_splitting.jl:98-108 */
void orc_hs_strang_splitting(orc_hs *h, double dt, int number_steps)
{
    for (int s = 0; s < number_steps; ++s) {
        orc_hs_operatorHB(h, 0.5 * dt);
        orc_hs_operatorHE(h, 0.5 * dt);
        orc_hs_operatorHp2(h, 0.5 * dt);
        orc_hs_operatorHp1(h, 1.0 * dt);
        orc_hs_operatorHp2(h, 0.5 * dt);
        orc_hs_operatorHE(h, 0.5 * dt);
        orc_hs_operatorHB(h, 0.5 * dt);
    }
}

/* ------------------------------------------------------------------------- */
/* HamiltonianSplitting{1,1}  (src/hamiltonian_splitting_1d1v.jl)             */
/* particle record = [x, v, w]                                                */
/* ------------------------------------------------------------------------- */
/* hamiltonian_splitting_1d1v.jl:63-98 */
void orc_hs11_operatorHp1(orc_hs *h, double dt)
{
    orc_pg *pg = h->pg;
    int n = h->ks0->n_grid;
    memset(h->j1, 0, sizeof(double) * (size_t)n);
    memset(h->j2, 0, sizeof(double) * (size_t)n);
    for (int64_t i = 0; i < pg->n_particles; ++i) {
        double x_old = PGA(pg, 0, i), v_old = PGA(pg, 1, i);
        double x_new = x_old + dt * v_old;
        double wi = pg->charge * PGA(pg, 2, i) * pg->common_weight;    /* get_charge */
        orc_pmc1d_add_current_1d1v(h->ks1, h->j1, x_old, x_new, wi, pg->q_over_m, v_old);
        x_new = orc_fmod_julia(x_new, h->Lx);
        PGA(pg, 0, i) = x_new;
    }
    orc_maxwell1d_compute_e_from_j(h->maxwell, h->e1, h->j1, 1);
}
/* hamiltonian_splitting_1d1v.jl:113-125 (the E-kick; note: no q/m) */
void orc_hs11_operatorHB(orc_hs *h, double dt)
{
    orc_pg *pg = h->pg;
    for (int64_t i = 0; i < pg->n_particles; ++i) {
        double xi = PGA(pg, 0, i), vi = PGA(pg, 1, i);
        double e1 = orc_pmc1d_evaluate(h->ks1, xi, h->e1);
        vi = vi + dt * e1;
        PGA(pg, 1, i) = vi;
    }
}
/* hamiltonian_splitting_1d1v.jl:11-21 */
void orc_hs11_strang_splitting(orc_hs *h, double dt, int number_steps)
{
    for (int s = 0; s < number_steps; ++s) {
        orc_hs11_operatorHB(h, 0.5 * dt);
        orc_hs11_operatorHp1(h, dt);
        orc_hs11_operatorHB(h, 0.5 * dt);
    }
}

/* ------------------------------------------------------------------------- */
/* HamiltonianSplittingBoris  (src/hamiltonian_splitting_boris.jl)            */
/* ------------------------------------------------------------------------- */
typedef struct {
    orc_hs base;                 /* maxwell, ks0, ks1, pg, e1,e2,b (caller arrays), j1,j2 */
    double *e1_mid, *e2_mid, *b_mid;
} orc_boris;

int orc_boris_init(orc_boris *s, orc_maxwell1d *m, orc_pmc1d *ks0, orc_pmc1d *ks1, orc_pg *pg,
                   double *e1, double *e2, double *b)
{
    int rc = orc_hs_init(&s->base, m, ks0, ks1, pg, e1, e2, b, 1);
    if (rc) return rc;
    int n = ks0->n_grid;
    s->e1_mid = calloc((size_t)n, sizeof(double));
    s->e2_mid = calloc((size_t)n, sizeof(double));
    s->b_mid = calloc((size_t)n, sizeof(double));
    return 0;
}
void orc_boris_free(orc_boris *s)
{
    orc_hs_free(&s->base);
    free(s->e1_mid); free(s->e2_mid); free(s->b_mid);
}
double *orc_boris_field(orc_boris *s, int which)
{
    switch (which) {
    case 0: return s->e1_mid;
    case 1: return s->e2_mid;
    case 2: return s->b_mid;
    case 3: return s->base.j1;
    case 4: return s->base.j2;
    default: return NULL;
    }
}

/* hamiltonian_splitting_boris.jl:250-288 */
void orc_boris_push_x_accumulate_j(orc_boris *s, double dt)
{
    orc_hs *h = &s->base;
    orc_pg *pg = h->pg;
    int n = h->ks0->n_grid;
    memset(h->j1, 0, sizeof(double) * (size_t)n);
    memset(h->j2, 0, sizeof(double) * (size_t)n);
    for (int64_t i = 0; i < pg->n_particles; ++i) {
        double x_old = PGA(pg, 0, i), v1 = PGA(pg, 1, i), v2 = PGA(pg, 2, i);
        double x_new = x_old + dt * v1;
        double wi = pg->charge * PGA(pg, 3, i) * pg->common_weight;
        orc_pmc1d_add_charge(h->ks1, h->j1, (x_old + x_new) * 0.5, wi * v1);
        orc_pmc1d_add_charge(h->ks0, h->j2, (x_old + x_new) * 0.5, wi * v2);
        x_new = orc_fmod_julia(x_new, h->Lx);
        PGA(pg, 0, i) = x_new;
    }
}
/* hamiltonian_splitting_boris.jl:189-204 */
void orc_boris_push_v_epart(orc_boris *s, double dt)
{
    orc_hs *h = &s->base;
    orc_pg *pg = h->pg;
    double qm = pg->q_over_m;
    for (int64_t i = 0; i < pg->n_particles; ++i) {
        double xi = PGA(pg, 0, i);
        double ef1 = orc_pmc1d_evaluate(h->ks1, xi, s->e1_mid);
        double ef2 = orc_pmc1d_evaluate(h->ks0, xi, s->e2_mid);
        PGA(pg, 1, i) = PGA(pg, 1, i) + dt * qm * ef1;
        PGA(pg, 2, i) = PGA(pg, 2, i) + dt * qm * ef2;
    }
}
/* hamiltonian_splitting_boris.jl:211-233 */
void orc_boris_push_v_bpart(orc_boris *s, double dt)
{
    orc_hs *h = &s->base;
    orc_pg *pg = h->pg;
    double qmdt = pg->q_over_m * 0.5 * dt;
    for (int64_t i = 0; i < pg->n_particles; ++i) {
        double vi1 = PGA(pg, 1, i), vi2 = PGA(pg, 2, i), xi = PGA(pg, 0, i);
        double bfield = orc_pmc1d_evaluate(h->ks1, xi, s->b_mid);
        bfield = qmdt * bfield;
        double M11 = 1.0 / (1.0 + bfield * bfield);
        double M12 = M11 * bfield * 2.0;
        M11 = M11 * (1 - bfield * bfield);
        PGA(pg, 1, i) = M11 * vi1 + M12 * vi2;
        PGA(pg, 2, i) = -M12 * vi1 + M11 * vi2;
    }
}
/* hamiltonian_splitting_boris.jl:99-122 */
void orc_boris_staggering(orc_boris *s, double dt)
{
    orc_hs *h = &s->base;
    int n = h->ks0->n_grid;
    orc_boris_push_x_accumulate_j(s, dt * 0.5);
    for (int i = 0; i < n; ++i) { s->e1_mid[i] = h->e1[i]; s->e2_mid[i] = h->e2[i]; }
    for (int i = 0; i < n; ++i) { h->j1[i] = 0.5 * dt * h->j1[i]; h->j2[i] = 0.5 * dt * h->j2[i]; }
    orc_maxwell1d_compute_e_from_j(h->maxwell, s->e1_mid, h->j1, 1);
    orc_maxwell1d_compute_e_from_b(h->maxwell, s->e2_mid, 0.5 * dt, h->b);
    orc_maxwell1d_compute_e_from_j(h->maxwell, s->e2_mid, h->j2, 2);
}
/* hamiltonian_splitting_boris.jl:132-177 */
void orc_boris_strang_splitting(orc_boris *s, double dt, int number_steps)
{
    orc_hs *h = &s->base;
    int n = h->ks0->n_grid;
    for (int st = 0; st < number_steps; ++st) {
        for (int i = 0; i < n; ++i) s->b_mid[i] = h->b[i];
        orc_maxwell1d_compute_b_from_e(h->maxwell, h->b, dt, s->e2_mid);
        for (int i = 0; i < n; ++i) s->b_mid[i] = (s->b_mid[i] + h->b[i]) * 0.5;
        orc_boris_push_v_epart(s, 0.5 * dt);
        orc_boris_push_v_bpart(s, dt);
        orc_boris_push_v_epart(s, 0.5 * dt);
        orc_boris_push_x_accumulate_j(s, dt);
        for (int i = 0; i < n; ++i) { h->e1[i] = s->e1_mid[i]; h->e2[i] = s->e2_mid[i]; }
        for (int i = 0; i < n; ++i) { h->j1[i] = dt * h->j1[i]; h->j2[i] = dt * h->j2[i]; }
        orc_maxwell1d_compute_e_from_j(h->maxwell, s->e1_mid, h->j1, 1);
        orc_maxwell1d_compute_e_from_b(h->maxwell, s->e2_mid, dt, h->b);
        orc_maxwell1d_compute_e_from_j(h->maxwell, s->e2_mid, h->j2, 2);
    }
}

/* ------------------------------------------------------------------------- */
/* diagnostics (src/diagnostics.jl)                                           */
/* ------------------------------------------------------------------------- */
/* diagnostics.jl:15-31 solve_poisson! */
void orc_solve_poisson(double *efield, orc_pg *pg, const orc_pmc1d *ks0, orc_maxwell1d *m, double *rho)
{
    memset(rho, 0, sizeof(double) * (size_t)ks0->n_grid);
    for (int64_t i = 0; i < pg->n_particles; ++i) {
        double xi = PGA(pg, 0, i);
        double wi = pg->charge * PGA(pg, pg->D + pg->V, i) * pg->common_weight;
        orc_pmc1d_add_charge(ks0, rho, xi, wi);
    }
    orc_maxwell1d_compute_e_from_rho(m, efield, rho);
}

/* diagnostics.jl:186-250 write_step! for ParticleGroup{1,2}; out[11] =
 * Time, KineticEnergy, Momentum1, Momentum2, PotentialEnergyE1, E2, B3, Transfer, VVB,
 * Poynting, ErrorPoisson (:143-155) */
void orc_write_step(orc_pg *pg, orc_maxwell1d *m, const orc_pmc1d *ks0, const orc_pmc1d *ks1,
                    double time, int degree, const double *e1, const double *e2, const double *b,
                    const double *e1_n, const double *e2_n, const double *e_poisson, double *out)
{
    int n = m->n_dofs;
    double d0 = 0.0, d1 = 0.0, d2 = 0.0;
    for (int64_t i = 0; i < pg->n_particles; ++i) {           /* :197-211 */
        double v1 = PGA(pg, 1, i), v2 = PGA(pg, 2, i), wi = PGA(pg, 3, i);
        wi *= pg->mass;
        wi *= pg->common_weight;
        d0 += (v1 * v1 + v2 * v2) * wi;
        d1 += v1 * wi;
        d2 += v2 * wi;
    }
    double transfer = 0.0;                                     /* :45-63 */
    for (int64_t i = 0; i < pg->n_particles; ++i) {
        double xi = PGA(pg, 0, i);
        double wi = pg->charge * PGA(pg, 3, i) * pg->common_weight;
        double v1 = PGA(pg, 1, i), v2 = PGA(pg, 2, i);
        double ef1 = orc_pmc1d_evaluate(ks1, xi, e1);
        double ef2 = orc_pmc1d_evaluate(ks0, xi, e2);
        transfer += (v1 * ef1 + v2 * ef2) * wi;
    }
    double vvb = 0.0;                                          /* :75-92 */
    for (int64_t i = 0; i < pg->n_particles; ++i) {
        double xi = PGA(pg, 0, i), v1 = PGA(pg, 1, i), v2 = PGA(pg, 2, i), wi = PGA(pg, 3, i);
        wi *= pg->charge;
        wi *= pg->common_weight;
        double bf = orc_pmc1d_evaluate(ks1, xi, b);
        vvb += wi * v1 * v2 * bf;
    }
    double *scratch = calloc((size_t)n, sizeof(double));       /* :106-112 (similar() is
        uninitialised in Julia; compute_e_from_b! then does scratch .+= ..., so the reference
        value is formally undefined -- restated with a zeroed scratch) */
    orc_maxwell1d_compute_e_from_b(m, scratch, 1.0, b);
    double poynting = orc_maxwell1d_inner_product(m, e2, scratch, degree);
    free(scratch);
    double pe1 = orc_maxwell1d_inner_product(m, e1, e1_n, degree - 1);
    double pe2 = orc_maxwell1d_inner_product(m, e2, e2_n, degree);
    double pb3 = orc_maxwell1d_l2norm_squared(m, b, degree - 1);
    double err = 0.0;
    for (int i = 0; i < n; ++i) { double d = fabs(e1[i] - e_poisson[i]); if (d > err) err = d; }
    out[0] = time; out[1] = d0; out[2] = d1; out[3] = d2;
    out[4] = pe1; out[5] = pe2; out[6] = pb3;
    out[7] = transfer; out[8] = vvb; out[9] = poynting; out[10] = err;
}

/* ------------------------------------------------------------------------- */
/* ParticleMeshCoupling2D  (src/particle_mesh_coupling_2d.jl)                 */
/* ------------------------------------------------------------------------- */
typedef struct {
    double xmin, ymin, dx, dy, scaling;
    int nx, ny, degree;
} orc_pmc2d;

int orc_pmc2d_init(orc_pmc2d *p, double xmin, double xmax, int nx, double ymin, double ymax, int ny,
                   int degree, int smoothing)
{
    if (degree < 0 || degree > ORC_MAXDEG) return -1;
    p->xmin = xmin; p->ymin = ymin; p->nx = nx; p->ny = ny;
    p->dx = (xmax - xmin) / nx;                  /* mesh.jl:37-38 */
    p->dy = (ymax - ymin) / ny;
    p->degree = degree;
    if (smoothing == 0) p->scaling = 1.0 / (p->dx * p->dy);
    else if (smoothing == 1) p->scaling = 1.0;
    else return -1;
    return 0;
}

/* particle_mesh_coupling_2d.jl:54-65 compute_shape_factor (ceil convention) */
static void orc_pmc2d_shape(const orc_pmc2d *p, double xp, double yp, double *vx, double *vy, long *ix, long *iy)
{
    xp = (xp - p->xmin) / p->dx;
    yp = (yp - p->ymin) / p->dy;
    long ip = (long)ceil(xp), jp = (long)ceil(yp);
    double dxp = xp - (double)(ip - 1), dyp = yp - (double)(jp - 1);
    orc_bsplines_eval_basis(p->degree, dxp, vx);       /* low_level_bsplines.jl:82-110 */
    orc_bsplines_eval_basis(p->degree, dyp, vy);
    *ix = ip - p->degree;
    *iy = jp - p->degree;
}
void orc_pmc2d_shape_indices(const orc_pmc2d *p, double xp, double yp, long *ix, long *iy)
{
    double vx[ORC_MAXDEG + 1], vy[ORC_MAXDEG + 1];
    orc_pmc2d_shape(p, xp, yp, vx, vy, ix, iy);
}

/* particle_mesh_coupling_2d.jl:94-105 add_charge! */
void orc_pmc2d_add_charge(const orc_pmc2d *p, double *rho, double xp, double yp, double wp)
{
    double vx[ORC_MAXDEG + 1], vy[ORC_MAXDEG + 1];
    long ix, iy;
    orc_pmc2d_shape(p, xp, yp, vx, vy, &ix, &iy);
    for (int i1 = 1; i1 <= p->degree + 1; ++i1) {
        long a = orc_mod(ix + i1 - 2, p->nx);
        for (int i2 = 1; i2 <= p->degree + 1; ++i2) {
            long b = orc_mod(iy + i2 - 2, p->ny);
            rho[a + b * p->nx] += (wp * p->scaling * vx[i1 - 1] * vy[i2 - 1]);
        }
    }
}
/* particle_mesh_coupling_2d.jl:189-203 evaluate */
double orc_pmc2d_evaluate(const orc_pmc2d *p, double xp, double yp, const double *field)
{
    double vx[ORC_MAXDEG + 1], vy[ORC_MAXDEG + 1];
    long ix, iy;
    orc_pmc2d_shape(p, xp, yp, vx, vy, &ix, &iy);
    double value = 0.0;
    for (int i1 = 1; i1 <= p->degree + 1; ++i1) {
        long a = orc_mod(ix + i1 - 2, p->nx);
        for (int i2 = 1; i2 <= p->degree + 1; ++i2) {
            long b = orc_mod(iy + i2 - 2, p->ny);
            value += field[a + b * p->nx] * vx[i1 - 1] * vy[i2 - 1];
        }
    }
    return value;
}
/* particle_mesh_coupling_2d.jl:214-231 evaluate_multiple */
void orc_pmc2d_evaluate_multiple(const orc_pmc2d *p, double xp, double yp, const double *f1,
                                 const double *f2, double *out)
{
    double vx[ORC_MAXDEG + 1], vy[ORC_MAXDEG + 1];
    long ix, iy;
    orc_pmc2d_shape(p, xp, yp, vx, vy, &ix, &iy);
    double a1 = 0.0, a2 = 0.0;
    for (int i1 = 1; i1 <= p->degree + 1; ++i1) {
        long a = orc_mod(ix + i1 - 2, p->nx);
        for (int i2 = 1; i2 <= p->degree + 1; ++i2) {
            long b = orc_mod(iy + i2 - 2, p->ny);
            double c = vx[i1 - 1] * vy[i2 - 1];
            a1 += f1[a + b * p->nx] * c;
            a2 += f2[a + b * p->nx] * c;
        }
    }
    out[0] = a1; out[1] = a2;
}

/* ------------------------------------------------------------------------- */
/* batched helpers (host loops the reference callers write themselves)        */
/* ------------------------------------------------------------------------- */
void orc_pmc1d_add_charge_batch(const orc_pmc1d *p, double *rho, const double *x, const double *w, int64_t n)
{
    for (int64_t i = 0; i < n; ++i) orc_pmc1d_add_charge(p, rho, x[i], w[i]);
}
void orc_pmc1d_evaluate_batch(const orc_pmc1d *p, const double *x, const double *field, double *out, int64_t n)
{
    for (int64_t i = 0; i < n; ++i) out[i] = orc_pmc1d_evaluate(p, x[i], field);
}
void orc_pmc2d_add_charge_batch(const orc_pmc2d *p, double *rho, const double *x, const double *y,
                                const double *w, int64_t n)
{
    for (int64_t i = 0; i < n; ++i) orc_pmc2d_add_charge(p, rho, x[i], y[i], w[i]);
}
void orc_pmc2d_evaluate_batch(const orc_pmc2d *p, const double *x, const double *y, const double *field,
                              double *out, int64_t n)
{
    for (int64_t i = 0; i < n; ++i) out[i] = orc_pmc2d_evaluate(p, x[i], y[i], field);
}

/* sizes so that the ctypes side can allocate opaque storage */
int orc_sizeof(int what)
{
    switch (what) {
    case 0: return (int)sizeof(orc_pmc1d);
    case 1: return (int)sizeof(orc_maxwell1d);
    case 2: return (int)sizeof(orc_pg);
    case 3: return (int)sizeof(orc_hs);
    case 4: return (int)sizeof(orc_boris);
    case 5: return (int)sizeof(orc_pmc2d);
    default: return -1;
    }
}
/* pin the OpenMP team size explicitly (bench.py: an inherited OMP_NUM_THREADS=1 under torchrun must not shrink the
 * CPU arm silently) */
void orc_set_threads(int n)
{
#ifdef _OPENMP
    if (n >= 1) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
