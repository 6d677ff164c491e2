// fields2d.cu -- TwoDMaxwell / TwoDPoisson / TwoDLinearSolverSplineMass on the device
// (src/maxwell_2d_fem.jl, src/poisson_2d_fem.jl, src/linear_solver_spline_mass_2d.jl).
//
// Compiled with -fmad=false like fields1d.cu: the point-wise stencils (compute_b_from_e!, curl,
// div) reproduce the reference's fp64 expressions; the cost of these kernels is launch latency
// (nx*ny = 4096 dofs per component in BASELINE config 5).
//
// Mass solves.  The reference inverts the Kronecker mass matrices with a complex 2D FFT, a
// division by eig1[i]*eig2[j] and an inverse FFT (linear_solver_spline_mass_2d.jl:16-32).
// Because the symbol is a product, the inverse is the composition of two 1D circulant
// operators; their (real, symmetric) first columns irfft(1/eig) are computed once on the host
// in extended precision and applied on the device as periodic convolutions along x, then y.
// No FFT library, no CPU fallback.
//
// Poisson.  The symbol dtm1d_1[i]*m0_2[j] + m0_1[i]*dtm1d_2[j] (poisson_2d_fem.jl:245-250) is
// not a product, so compute_e_from_rho! runs a real 2D DFT as dense twiddle sums along x and
// y (k_dft_axis), scales the modes, and transforms -phi*d1 and -phi*d2 back.  It is used at
// set-up and in diagnostics only.
//
// Dof layout: flat nx*ny vector, x fastest (ind2d = (j-1)*nx + i, maxwell_2d_fem.jl:389-391).
#include <cmath>
#include <cstring>

#include "hostutil.hpp"
#include "objects.cuh"

namespace gempic {

constexpr int kT2 = 256;

struct AxisJob {
    const double *col;   // circulant first column (n_axis) or mass line (deg+1)
    const double *in;
    double *out;
    const double *base;  // for mode 1/2
    int mode;            // 0: out = r   1: out = base - r   2: out = base + scale*r
    int deg;             // banded kernels: number of off-diagonals
    double scale;
};
struct AxisJobs {
    AxisJob j[3];
};

// out[i,j] = sum_k col[(i-k) mod nx] in[k,j]   (AXIS 0)   /   sum_k col[(j-k) mod ny] in[i,k]   (AXIS 1)
template <int AXIS>
__global__ void k_circ_axis(const __grid_constant__ AxisJobs J, int nx, int ny)
{
    const AxisJob &q = J.j[blockIdx.y];
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nx * ny) return;
    const int j = idx / nx, i = idx - j * nx;
    const int n = AXIS == 0 ? nx : ny;
    const int me = AXIS == 0 ? i : j;
    const double *line = AXIS == 0 ? q.in + (size_t)j * nx : q.in + i;
    const int stride = AXIS == 0 ? 1 : nx;
    double acc = 0.0;
    int c = me;  // (me - k) mod n
    for (int k = 0; k < n; ++k) {
        acc += q.col[c] * line[(size_t)k * stride];
        c = c == 0 ? n - 1 : c - 1;
    }
    double r = acc;
    if (q.mode == 1) r = q.base[idx] - acc;
    else if (q.mode == 2) r = q.base[idx] + q.scale * acc;
    q.out[idx] = r;
}

// spline_fem_multiply_mass along one axis (maxwell_2d_fem.jl:300-338): periodic banded multiply,
// out[row] = mass[0] in[row] + sum_c mass[c] (in[row+c] + in[row-c])
template <int AXIS>
__global__ void k_band_axis(const __grid_constant__ AxisJobs J, int nx, int ny)
{
    const AxisJob &q = J.j[blockIdx.y];
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nx * ny) return;
    const int j = idx / nx, i = idx - j * nx;
    const int n = AXIS == 0 ? nx : ny;
    const int me = AXIS == 0 ? i : j;
    const double *line = AXIS == 0 ? q.in + (size_t)j * nx : q.in + i;
    const int stride = AXIS == 0 ? 1 : nx;
    double acc = q.col[0] * line[(size_t)me * stride];
    for (int c = 1; c <= q.deg; ++c) {
        int up = me + c, dn = me - c;
        up = up >= n ? up - n : up;
        dn = dn < 0 ? dn + n : dn;
        acc += q.col[c] * (line[(size_t)up * stride] + line[(size_t)dn * stride]);
    }
    q.out[idx] = acc;
}

// curl of the mass-weighted B (compute_e_from_b!, maxwell_2d_fem.jl:385-401)
__global__ void k_curl_b(const double *__restrict__ w1, const double *__restrict__ w2, const double *__restrict__ w3,
                         double *__restrict__ c1, double *__restrict__ c2, double *__restrict__ c3, int nx, int ny,
                         double dx1, double dx2)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nx * ny) return;
    const int j = idx / nx, i = idx - j * nx;
    const int i1 = idx + (i == nx - 1 ? 1 - nx : 1);
    const int i2 = idx + (j == ny - 1 ? nx * (1 - ny) : nx);
    c1[idx] = -(w3[idx] - w3[i2]) / dx2;
    c2[idx] = (w3[idx] - w3[i1]) / dx1;
    c3[idx] = (w1[idx] - w1[i2]) / dx2 - (w2[idx] - w2[i1]) / dx1;
}

// compute_b_from_e! (:423-444)
__global__ void k_b_from_e2d(double *__restrict__ b1, double *__restrict__ b2, double *__restrict__ b3,
                             const double *__restrict__ e1, const double *__restrict__ e2, const double *__restrict__ e3,
                             int nx, int ny, double dx1, double dx2, double dt)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nx * ny) return;
    const int j = idx / nx, i = idx - j * nx;
    const int i1 = idx + (i == 0 ? nx - 1 : -1);
    const int i2 = idx + (j == 0 ? nx * (ny - 1) : -nx);
    b1[idx] += -dt * (e3[idx] - e3[i2]) / dx2;
    b2[idx] += dt * (e3[idx] - e3[i1]) / dx1;
    b3[idx] += -dt * ((e2[idx] - e2[i1]) / dx1 - (e1[idx] - e1[i2]) / dx2);
}

// compute_rho_from_e! (:468-500) on the mass-weighted E
__global__ void k_div_e(const double *__restrict__ w1, const double *__restrict__ w2, double *__restrict__ rho, int nx,
                        int ny, double dx1, double dx2)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nx * ny) return;
    const int j = idx / nx, i = idx - j * nx;
    const int i1 = idx + (i == nx - 1 ? 1 - nx : 1);
    const int i2 = idx + (j == ny - 1 ? -nx * (ny - 1) : nx);
    const double r = (w1[idx] - w1[i1]) / dx1 + (w2[idx] - w2[i2]) / dx2;
    rho[idx] = r * -1.0;
}

// dense DFT along one axis: out[k] = sum_m in[m] exp(sign 2 pi i k m / n); `im` may be null
template <int AXIS>
__global__ void k_dft_axis(const double *__restrict__ cs, const double *__restrict__ sn, const double *__restrict__ re,
                           const double *__restrict__ im, double *__restrict__ ore, double *__restrict__ oim, int nx,
                           int ny, double sign, double scale)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nx * ny) return;
    const int j = idx / nx, i = idx - j * nx;
    const int n = AXIS == 0 ? nx : ny;
    const int k = AXIS == 0 ? i : j;
    const size_t base = AXIS == 0 ? (size_t)j * nx : (size_t)i;
    const int stride = AXIS == 0 ? 1 : nx;
    double ar = 0.0, ai = 0.0;
    int t = 0;  // (k*m) mod n
    for (int m = 0; m < n; ++m) {
        const double c = cs[t], s = sign * sn[t];
        const double xr = re[base + (size_t)m * stride];
        const double xi = im ? im[base + (size_t)m * stride] : 0.0;
        ar += xr * c - xi * s;
        ai += xr * s + xi * c;
        t += k;
        if (t >= n) t -= n;
    }
    ore[idx] = ar * scale;
    if (oim) oim[idx] = ai * scale;
}

// Poisson mode scaling (poisson_2d_fem.jl:242-256): s = rho_hat / eig, sx = -s d1[i], sy = -s d2[j]
__global__ void k_poisson_modes(const double *__restrict__ sre, const double *__restrict__ sim,
                                const double *__restrict__ dtm1, const double *__restrict__ m01,
                                const double *__restrict__ dtm2, const double *__restrict__ m02,
                                const double *__restrict__ d1re, const double *__restrict__ d1im,
                                const double *__restrict__ d2re, const double *__restrict__ d2im, double *__restrict__ xre,
                                double *__restrict__ xim, double *__restrict__ yre, double *__restrict__ yim, int nx, int ny)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nx * ny) return;
    const int j = idx / nx, i = idx - j * nx;
    double pr = 0.0, pi = 0.0;
    if (idx != 0) {
        const double eig = dtm1[i] * m02[j] + m01[i] * dtm2[j];
        pr = sre[idx] / eig;
        pi = sim[idx] / eig;
    }
    xre[idx] = -(pr * d1re[i] - pi * d1im[i]);
    xim[idx] = -(pr * d1im[i] + pi * d1re[i]);
    yre[idx] = -(pr * d2re[j] - pi * d2im[j]);
    yim[idx] = -(pr * d2im[j] + pi * d2re[j]);
}

// sum(c1 .* w) in index order per thread, fixed tree over threads (inner_product :574)
__global__ void k_dot(const double *__restrict__ a, const double *__restrict__ b, int n, double *__restrict__ out)
{
    __shared__ double red[kT2];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += kT2) s += a[i] * b[i];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int h = kT2 / 2; h > 0; h >>= 1) {
        if (threadIdx.x < h) red[threadIdx.x] += red[threadIdx.x + h];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = red[0];
}

// ---- object -----------------------------------------------------------------------------------
static std::vector<double> mass_line(int degree)   // spline_fem_mass_line, poisson_2d_fem.jl:87-106
{
    const int n = degree + 1;
    double x[8], w[8], val[8][8];
    legendre_nodes(n, x, w);
    for (int j = 0; j < n; ++j) {
        x[j] = (x[j] + 1.0) * 0.5;
        w[j] = w[j] * 0.5;
        host_bsplines(degree, x[j], val[j]);   // val[j][i] = spline i at node j
    }
    std::vector<double> line(degree + 1, 0.0);
    for (int j = 1; j <= degree + 1; ++j)
        for (int i = j; i <= degree + 1; ++i)
            for (int k = 0; k < n; ++k) line[j - 1] += val[k][i - 1] * val[k][i - j] * w[k];
    return line;
}

static std::vector<double> mass_eig(int n_cells, int degree, const std::vector<double> &line)   // :113-126
{
    std::vector<double> eig(n_cells);
    const double factor = 2.0 * 3.14159265358979323846 / n_cells;
    for (int k = 0; k < n_cells; ++k) {
        eig[k] = line[0];
        for (int j = 1; j <= degree; ++j) eig[k] += line[j] * 2 * std::cos(factor * k * j);
    }
    return eig;
}

// first column of the circulant with real symmetric symbol 1/eig: col[m] = (1/n) sum_k cos(2 pi k m / n) / eig[k]
static std::vector<double> inverse_column(const std::vector<double> &eig)
{
    const int n = (int)eig.size();
    const long double two_pi = 6.283185307179586476925286766559005768L;
    std::vector<double> col(n);
    for (int m = 0; m < n; ++m) {
        long double acc = 0.0L;
        for (int k = 0; k < n; ++k) acc += cosl(two_pi * (long double)(((long long)k * m) % n) / n) / (long double)eig[k];
        col[m] = (double)(acc / n);
    }
    return col;
}

std::unique_ptr<Maxwell2D> make_maxwell2d(double xmin, double xmax, int nx, double ymin, double ymax, int ny, int degree)
{
    GP_REQUIRE(degree >= 1 && degree <= 3, GEMPIC_EINVAL, "Wrong value of degree = %d  (1,2 or 3)", degree);
    GP_REQUIRE(nx >= 2 * degree + 1 && ny >= 2 * degree + 1, GEMPIC_EINVAL, "grid %d x %d too small for degree %d", nx, ny, degree);
    GP_REQUIRE(nx <= 4096 && ny <= 4096, GEMPIC_EINVAL, "grid %d x %d exceeds the supported 4096 per axis", nx, ny);
    auto m = std::make_unique<Maxwell2D>();
    m->nx = nx; m->ny = ny;
    m->xmin = xmin; m->ymin = ymin;
    m->Lx = xmax - xmin; m->Ly = ymax - ymin;
    m->dx = m->Lx / nx; m->dy = m->Ly / ny;
    m->s_deg_0 = degree; m->s_deg_1 = degree - 1;
    const int n_ax[2] = {nx, ny};
    const double d_ax[2] = {m->dx, m->dy};
    std::vector<double> l0 = mass_line(degree), l1 = mass_line(degree - 1);
    for (int a = 0; a < 2; ++a) {
        m->line[0][a] = l0;
        m->line[1][a] = l1;
        for (auto &v : m->line[0][a]) v *= d_ax[a];   // maxwell_2d_fem.jl:36-42
        for (auto &v : m->line[1][a]) v *= d_ax[a];
        m->eig[0][a] = mass_eig(n_ax[a], degree, m->line[0][a]);
        m->eig[1][a] = mass_eig(n_ax[a], degree - 1, m->line[1][a]);
    }
    // device tables
    //   inv columns [deg][axis] | mass lines [deg][axis] (4 each) | twiddles cos,sin per axis | poisson tables
    const size_t nmax = (size_t)std::max(nx, ny);
    std::vector<double> tab;
    auto push = [&](const std::vector<double> &v, size_t pad) {
        const size_t off = tab.size();
        tab.insert(tab.end(), v.begin(), v.end());
        tab.resize(off + pad, 0.0);
        return off;
    };
    for (int d = 0; d < 2; ++d)
        for (int a = 0; a < 2; ++a) m->off_inv[d][a] = push(inverse_column(m->eig[d][a]), nmax);
    for (int d = 0; d < 2; ++d)
        for (int a = 0; a < 2; ++a) m->off_line[d][a] = push(m->line[d][a], 4);
    const long double two_pi = 6.283185307179586476925286766559005768L;
    for (int a = 0; a < 2; ++a) {
        std::vector<double> cs(n_ax[a]), sn(n_ax[a]);
        for (int t = 0; t < n_ax[a]; ++t) {
            cs[t] = (double)cosl(two_pi * t / n_ax[a]);
            sn[t] = (double)sinl(two_pi * t / n_ax[a]);
        }
        m->off_cos[a] = push(cs, nmax);
        m->off_sin[a] = push(sn, nmax);
    }
    // TwoDPoisson tables (poisson_2d_fem.jl:49-81); its mass eigenvalues use mass_line * dx (same as above)
    for (int a = 0; a < 2; ++a) {
        const int n = n_ax[a];
        const double h = d_ax[a];
        std::vector<double> dre(n, 0.0), dim(n, 0.0), dtm(n, 0.0);
        for (int j = 1; j < n; ++j) {
            const double angle = 2 * 3.14159265358979323846 * j / n;
            dre[j] = (1 - std::cos(angle)) / h;
            dim[j] = std::sin(angle) / h;
            dtm[j] = 2 / (h * h) * (1 - std::cos(angle)) * m->eig[1][a][j];
        }
        m->off_dre[a] = push(dre, nmax);
        m->off_dim[a] = push(dim, nmax);
        m->off_dtm[a] = push(dtm, nmax);
        m->off_m0[a] = push(m->eig[0][a], nmax);
    }
    m->tab.alloc(tab.size());
    GP_CUDA(cudaMemcpyAsync(m->tab.p, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice, ctx().stream));
    GP_CUDA(cudaStreamSynchronize(ctx().stream));
    m->work.alloc((size_t)Maxwell2D::kWork * nx * ny + 8);
    return m;
}

static inline dim3 grid2(const Maxwell2D &m, int jobs) { return dim3((m.nx * m.ny + kT2 - 1) / kT2, jobs); }

// out_c = M^-1(degx_c, degy_c) in_c for up to 3 components, with the epilogue of `mode`
void m2d_solve_mass(const Maxwell2D &m, int njobs, const int (*deg)[2], const double *const *in, double *const *out,
                    const double *const *base, int mode, double scale)
{
    AxisJobs jx{}, jy{};
    for (int c = 0; c < njobs; ++c) {
        double *tmp = m.wk(Maxwell2D::kWork - 1 - c);
        jx.j[c] = AxisJob{m.inv_col(deg[c][0], 0), in[c], tmp, nullptr, 0, 0, 1.0};
        jy.j[c] = AxisJob{m.inv_col(deg[c][1], 1), tmp, out[c], base ? base[c] : nullptr, mode, 0, scale};
    }
    k_circ_axis<0><<<grid2(m, njobs), kT2, 0, ctx().stream>>>(jx, m.nx, m.ny);
    k_circ_axis<1><<<grid2(m, njobs), kT2, 0, ctx().stream>>>(jy, m.nx, m.ny);
    GP_CUDA(cudaGetLastError());
    count_launch(2);
}

// out_c = M(degx_c, degy_c) in_c  (multiply_mass_2dkron!, :343-363)
void m2d_multiply_mass(const Maxwell2D &m, int njobs, const int (*deg)[2], const double *const *in, double *const *out)
{
    AxisJobs jx{}, jy{};
    for (int c = 0; c < njobs; ++c) {
        double *tmp = m.wk(Maxwell2D::kWork - 1 - c);
        const int dx_ = deg[c][0] == 0 ? m.s_deg_0 : m.s_deg_1, dy_ = deg[c][1] == 0 ? m.s_deg_0 : m.s_deg_1;
        jx.j[c] = AxisJob{m.mass_line(deg[c][0], 0), in[c], tmp, nullptr, 0, dx_, 1.0};
        jy.j[c] = AxisJob{m.mass_line(deg[c][1], 1), tmp, out[c], nullptr, 0, dy_, 1.0};
    }
    k_band_axis<0><<<grid2(m, njobs), kT2, 0, ctx().stream>>>(jx, m.nx, m.ny);
    k_band_axis<1><<<grid2(m, njobs), kT2, 0, ctx().stream>>>(jy, m.nx, m.ny);
    GP_CUDA(cudaGetLastError());
    count_launch(2);
}

// which spline family (0: s_deg_0, 1: s_deg_1) a (component, form) uses along x and y
// (maxwell_2d_fem.jl:136-149)
void m2d_form_degrees(int component, int form, int out[2])
{
    GP_REQUIRE(form >= 0 && form <= 3, GEMPIC_EINVAL, " Wrong form ");
    GP_REQUIRE(component >= 1 && component <= 3, GEMPIC_EINVAL, "component %d not in 1:3", component);
    int d[2];
    if (form == 0) { d[0] = d[1] = 0; }
    else if (form == 1) { d[0] = d[1] = 0; if (component < 3) d[component - 1] = 1; }
    else if (form == 2) { d[0] = d[1] = 1; if (component < 3) d[component - 1] = 0; }
    else { d[0] = d[1] = 1; }
    out[0] = d[0];
    out[1] = d[1];
}

// compute_e_from_j! (:455-459): e -= inv_mass_1[component] j
void m2d_e_from_j(const Maxwell2D &m, double *e, const double *j, int component)
{
    int deg[1][2];
    m2d_form_degrees(component, 1, deg[0]);
    const double *in[1] = {j};
    double *out[1] = {e};
    const double *base[1] = {e};
    m2d_solve_mass(m, 1, deg, in, out, base, 1, 1.0);
}

// compute_e_from_b! (:370-412)
void m2d_e_from_b(const Maxwell2D &m, double *const e[3], double dt, const double *const b[3])
{
    int d2[3][2], d1[3][2];
    for (int c = 0; c < 3; ++c) {
        m2d_form_degrees(c + 1, 2, d2[c]);
        m2d_form_degrees(c + 1, 1, d1[c]);
    }
    double *w[3] = {m.wk(0), m.wk(1), m.wk(2)};
    double *cu[3] = {m.wk(3), m.wk(4), m.wk(5)};
    const double *bin[3] = {b[0], b[1], b[2]};
    m2d_multiply_mass(m, 3, d2, bin, w);
    k_curl_b<<<grid2(m, 1), kT2, 0, ctx().stream>>>(w[0], w[1], w[2], cu[0], cu[1], cu[2], m.nx, m.ny, m.dx, m.dy);
    GP_CUDA(cudaGetLastError());
    count_launch();
    const double *cin[3] = {cu[0], cu[1], cu[2]};
    double *eout[3] = {e[0], e[1], e[2]};
    const double *ebase[3] = {e[0], e[1], e[2]};
    m2d_solve_mass(m, 3, d1, cin, eout, ebase, 2, dt);
}

// compute_b_from_e! (:423-444)
void m2d_b_from_e(const Maxwell2D &m, double *const b[3], double dt, const double *const e[3])
{
    k_b_from_e2d<<<grid2(m, 1), kT2, 0, ctx().stream>>>(b[0], b[1], b[2], e[0], e[1], e[2], m.nx, m.ny, m.dx, m.dy, dt);
    GP_CUDA(cudaGetLastError());
    count_launch();
}

// compute_rho_from_e! (:468-500)
void m2d_rho_from_e(const Maxwell2D &m, double *rho, const double *const e[3])
{
    int d1[2][2];
    m2d_form_degrees(1, 1, d1[0]);
    m2d_form_degrees(2, 1, d1[1]);
    double *w[2] = {m.wk(0), m.wk(1)};
    const double *ein[2] = {e[0], e[1]};
    m2d_multiply_mass(m, 2, d1, ein, w);
    k_div_e<<<grid2(m, 1), kT2, 0, ctx().stream>>>(w[0], w[1], rho, m.nx, m.ny, m.dx, m.dy);
    GP_CUDA(cudaGetLastError());
    count_launch();
}

// compute_e_from_rho! (poisson_2d_fem.jl:237-262)
void m2d_e_from_rho(const Maxwell2D &m, double *e1, double *e2, const double *rho)
{
    cudaStream_t s = ctx().stream;
    const dim3 g = grid2(m, 1);
    const double *cx = m.tab.p + m.off_cos[0], *sx = m.tab.p + m.off_sin[0];
    const double *cy = m.tab.p + m.off_cos[1], *sy = m.tab.p + m.off_sin[1];
    double *ar = m.wk(0), *ai = m.wk(1), *br = m.wk(2), *bi = m.wk(3);
    double *xr = m.wk(4), *xi = m.wk(5), *yr = m.wk(6), *yi = m.wk(7);
    k_dft_axis<0><<<g, kT2, 0, s>>>(cx, sx, rho, nullptr, ar, ai, m.nx, m.ny, -1.0, 1.0);
    k_dft_axis<1><<<g, kT2, 0, s>>>(cy, sy, ar, ai, br, bi, m.nx, m.ny, -1.0, 1.0);
    k_poisson_modes<<<g, kT2, 0, s>>>(br, bi, m.tab.p + m.off_dtm[0], m.tab.p + m.off_m0[0], m.tab.p + m.off_dtm[1],
                                      m.tab.p + m.off_m0[1], m.tab.p + m.off_dre[0], m.tab.p + m.off_dim[0],
                                      m.tab.p + m.off_dre[1], m.tab.p + m.off_dim[1], xr, xi, yr, yi, m.nx, m.ny);
    // inverse transforms (ifft2d!, :184-206): along y, then x, real part kept
    k_dft_axis<1><<<g, kT2, 0, s>>>(cy, sy, xr, xi, ar, ai, m.nx, m.ny, 1.0, 1.0 / m.ny);
    k_dft_axis<0><<<g, kT2, 0, s>>>(cx, sx, ar, ai, e1, nullptr, m.nx, m.ny, 1.0, 1.0 / m.nx);
    k_dft_axis<1><<<g, kT2, 0, s>>>(cy, sy, yr, yi, ar, ai, m.nx, m.ny, 1.0, 1.0 / m.ny);
    k_dft_axis<0><<<g, kT2, 0, s>>>(cx, sx, ar, ai, e2, nullptr, m.nx, m.ny, 1.0, 1.0 / m.nx);
    GP_CUDA(cudaGetLastError());
    count_launch(7);
}

// inner_product (:512-575): sum(c1 .* (M c2)); result in out[0] (device)
void m2d_inner_product(const Maxwell2D &m, const double *c1, const double *c2, int component, int form, double *out)
{
    int deg[1][2];
    m2d_form_degrees(component, form, deg[0]);
    const double *in[1] = {c2};
    double *w[1] = {m.wk(0)};
    m2d_multiply_mass(m, 1, deg, in, w);
    k_dot<<<1, kT2, 0, ctx().stream>>>(c1, w[0], m.nx * m.ny, out);
    GP_CUDA(cudaGetLastError());
    count_launch();
}

}  // namespace gempic

// =============================== C ABI ==========================================================
using namespace gempic;

extern "C" {

int gempic_maxwell2d_create(double xmin, double xmax, int nx, double ymin, double ymax, int ny, int degree,
                            gempic_handle *out)
{
    GP_API_BEGIN
    require_init();
    GP_REQUIRE(out, GEMPIC_EINVAL, "null output handle");
    GP_REQUIRE(xmax > xmin && ymax > ymin, GEMPIC_EINVAL, "bad mesh");
    *out = register_object(make_maxwell2d(xmin, xmax, nx, ymin, ymax, ny, degree));
    GP_API_END
}

int gempic_maxwell2d_destroy(gempic_handle m)
{
    GP_API_BEGIN
    destroy(m, Kind::Maxwell2D, "TwoDMaxwell");
    GP_API_END
}

int gempic_maxwell2d_get_table(gempic_handle mh, int which, int axis, double *out, int *count)
{
    GP_API_BEGIN
    Maxwell2D *m = get<Maxwell2D>(mh, "TwoDMaxwell");
    GP_REQUIRE(out && count && (axis == 0 || axis == 1), GEMPIC_EINVAL, "bad argument");
    const std::vector<double> *t = nullptr;
    switch (which) {
    case 0: t = &m->line[0][axis]; break;   // mass_line_0
    case 1: t = &m->line[1][axis]; break;   // mass_line_1
    case 2: t = &m->eig[0][axis]; break;    // eig_values_mass_0
    case 3: t = &m->eig[1][axis]; break;    // eig_values_mass_1
    default: fail(GEMPIC_EINVAL, "unknown table %d", which);
    }
    std::memcpy(out, t->data(), sizeof(double) * t->size());
    *count = (int)t->size();
    GP_API_END
}

int gempic_maxwell2d_compute_e_from_rho(gempic_handle mh, double *e1, double *e2, const double *rho)
{
    GP_API_BEGIN
    require_init();
    Maxwell2D *m = get<Maxwell2D>(mh, "TwoDMaxwell");
    GP_REQUIRE(e1 && e2 && rho, GEMPIC_EINVAL, "null buffer");
    const size_t n = (size_t)m->nx * m->ny;
    Stage st(3 * n);
    double *dr = st.put(rho, n), *d1 = st.take(n), *d2 = st.take(n);
    m2d_e_from_rho(*m, d1, d2, dr);
    d2h(e1, d1, n);
    d2h(e2, d2, n);
    GP_API_END
}

int gempic_maxwell2d_compute_e_from_b(gempic_handle mh, double *e1, double *e2, double *e3, double dt, const double *b1,
                                      const double *b2, const double *b3)
{
    GP_API_BEGIN
    require_init();
    Maxwell2D *m = get<Maxwell2D>(mh, "TwoDMaxwell");
    GP_REQUIRE(e1 && e2 && e3 && b1 && b2 && b3, GEMPIC_EINVAL, "null buffer");
    const size_t n = (size_t)m->nx * m->ny;
    Stage st(6 * n);
    double *e[3] = {st.put(e1, n), st.put(e2, n), st.put(e3, n)};
    const double *b[3] = {st.put(b1, n), st.put(b2, n), st.put(b3, n)};
    m2d_e_from_b(*m, e, dt, b);
    d2h(e1, e[0], n);
    d2h(e2, e[1], n);
    d2h(e3, e[2], n);
    GP_API_END
}

int gempic_maxwell2d_compute_b_from_e(gempic_handle mh, double *b1, double *b2, double *b3, double dt, const double *e1,
                                      const double *e2, const double *e3)
{
    GP_API_BEGIN
    require_init();
    Maxwell2D *m = get<Maxwell2D>(mh, "TwoDMaxwell");
    GP_REQUIRE(e1 && e2 && e3 && b1 && b2 && b3, GEMPIC_EINVAL, "null buffer");
    const size_t n = (size_t)m->nx * m->ny;
    Stage st(6 * n);
    double *b[3] = {st.put(b1, n), st.put(b2, n), st.put(b3, n)};
    const double *e[3] = {st.put(e1, n), st.put(e2, n), st.put(e3, n)};
    m2d_b_from_e(*m, b, dt, e);
    d2h(b1, b[0], n);
    d2h(b2, b[1], n);
    d2h(b3, b[2], n);
    GP_API_END
}

int gempic_maxwell2d_compute_e_from_j(gempic_handle mh, double *e, const double *current, int component)
{
    GP_API_BEGIN
    require_init();
    Maxwell2D *m = get<Maxwell2D>(mh, "TwoDMaxwell");
    GP_REQUIRE(e && current, GEMPIC_EINVAL, "null buffer");
    GP_REQUIRE(component >= 1 && component <= 3, GEMPIC_EINVAL, "component %d not in 1:3", component);
    const size_t n = (size_t)m->nx * m->ny;
    Stage st(2 * n);
    double *de = st.put(e, n), *dj = st.put(current, n);
    m2d_e_from_j(*m, de, dj, component);
    d2h(e, de, n);
    GP_API_END
}

int gempic_maxwell2d_compute_rho_from_e(gempic_handle mh, double *rho, const double *e1, const double *e2, const double *e3)
{
    GP_API_BEGIN
    require_init();
    Maxwell2D *m = get<Maxwell2D>(mh, "TwoDMaxwell");
    GP_REQUIRE(rho && e1 && e2, GEMPIC_EINVAL, "null buffer");
    const size_t n = (size_t)m->nx * m->ny;
    Stage st(3 * n);
    const double *e[3] = {st.put(e1, n), st.put(e2, n), nullptr};
    double *dr = st.take(n);
    (void)e3;   // the weak Gauss law only involves components 1 and 2 (:479-497)
    m2d_rho_from_e(*m, dr, e);
    d2h(rho, dr, n);
    GP_API_END
}

int gempic_maxwell2d_inner_product(gempic_handle mh, const double *c1, const double *c2, int component, int form,
                                   double *out)
{
    GP_API_BEGIN
    require_init();
    Maxwell2D *m = get<Maxwell2D>(mh, "TwoDMaxwell");
    GP_REQUIRE(c1 && c2 && out, GEMPIC_EINVAL, "null buffer");
    const size_t n = (size_t)m->nx * m->ny;
    Stage st(2 * n + 8);
    double *d1 = st.put(c1, n), *d2 = st.put(c2, n), *dr = st.take(8);
    m2d_inner_product(*m, d1, d2, component, form, dr);
    d2h(out, dr, 1);
    GP_API_END
}

/* solve with inv_mass_1[component] (form 1) or inv_mass_2[component] (form 2): out = M^-1 rhs */
int gempic_maxwell2d_solve_mass(gempic_handle mh, double *out, const double *rhs, int component, int form)
{
    GP_API_BEGIN
    require_init();
    Maxwell2D *m = get<Maxwell2D>(mh, "TwoDMaxwell");
    GP_REQUIRE(out && rhs, GEMPIC_EINVAL, "null buffer");
    GP_REQUIRE(form == 1 || form == 2, GEMPIC_EINVAL, "form %d has no mass solver (1 or 2)", form);
    const size_t n = (size_t)m->nx * m->ny;
    Stage st(2 * n);
    int deg[1][2];
    m2d_form_degrees(component, form, deg[0]);
    const double *in[1] = {st.put(rhs, n)};
    double *o[1] = {st.take(n)};
    m2d_solve_mass(*m, 1, deg, in, o, nullptr, 0, 1.0);
    d2h(out, o[0], n);
    GP_API_END
}

int gempic_maxwell2d_multiply_mass(gempic_handle mh, double *out, const double *in_, int component, int form)
{
    GP_API_BEGIN
    require_init();
    Maxwell2D *m = get<Maxwell2D>(mh, "TwoDMaxwell");
    GP_REQUIRE(out && in_, GEMPIC_EINVAL, "null buffer");
    const size_t n = (size_t)m->nx * m->ny;
    Stage st(2 * n);
    int deg[1][2];
    m2d_form_degrees(component, form, deg[0]);
    const double *in[1] = {st.put(in_, n)};
    double *o[1] = {st.take(n)};
    m2d_multiply_mass(*m, 1, deg, in, o);
    d2h(out, o[0], n);
    GP_API_END
}

/* compute_rhs_from_function (:124-196) == compute_fem_rhs! (:211-277): host quadrature of a C callback (set-up path) */
int gempic_maxwell2d_compute_rhs_from_function(gempic_handle mh, double *coefs, gempic_func2d f, void *fctx, int component,
                                               int form)
{
    GP_API_BEGIN
    Maxwell2D *m = get<Maxwell2D>(mh, "TwoDMaxwell");
    GP_REQUIRE(coefs && f, GEMPIC_EINVAL, "null argument");
    int fam[2];
    m2d_form_degrees(component, form, fam);
    const int d1 = fam[0] == 0 ? m->s_deg_0 : m->s_deg_1, d2 = fam[1] == 0 ? m->s_deg_0 : m->s_deg_1;
    double x1[8], w1[8], b1[8][8], x2[8], w2[8], b2[8][8];
    legendre_nodes(d1 + 1, x1, w1);
    legendre_nodes(d2 + 1, x2, w2);
    for (int k = 0; k <= d1; ++k) {
        x1[k] = (x1[k] + 1.0) / 2.0;
        w1[k] = w1[k] / 2.0;
        host_bsplines(d1, x1[k], b1[k]);
    }
    for (int k = 0; k <= d2; ++k) {
        x2[k] = (x2[k] + 1.0) / 2.0;
        w2[k] = w2[k] / 2.0;
        host_bsplines(d2, x2[k], b2[k]);
    }
    size_t counter = 0;
    for (int i2 = 1; i2 <= m->ny; ++i2)
        for (int i1 = 1; i1 <= m->nx; ++i1) {
            double coef = 0.0;
            for (int j1 = 1; j1 <= d1 + 1; ++j1)
                for (int j2 = 1; j2 <= d2 + 1; ++j2)
                    for (int k1 = 0; k1 <= d1; ++k1)
                        for (int k2 = 0; k2 <= d2; ++k2) {
                            const double x = m->dx * (x1[k1] + i1 + j1 - 2);
                            const double y = m->dy * (x2[k2] + i2 + j2 - 2);
                            coef += w1[k1] * w2[k2] * f(x, y, fctx) * b1[k1][d1 + 1 - j1] * b2[k2][d2 + 1 - j2];
                        }
            coefs[counter++] = coef * m->dx * m->dy;
        }
    GP_API_END
}

/* l2projection (:283-293) */
int gempic_maxwell2d_l2projection(gempic_handle mh, double *coefs, gempic_func2d f, void *fctx, int component, int form)
{
    GP_API_BEGIN
    GP_REQUIRE(form == 1 || form == 2, GEMPIC_EINVAL, "l2projection: form %d has no mass solver (1 or 2)", form);
    int rc = gempic_maxwell2d_compute_rhs_from_function(mh, coefs, f, fctx, component, form);
    if (rc) return rc;
    rc = gempic_maxwell2d_solve_mass(mh, coefs, coefs, component, form);
    if (rc) return rc;
    GP_API_END
}

}  // extern "C"
