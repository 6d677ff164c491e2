// fields1d.cu -- Maxwell1DFEM on the device (src/maxwell_1d_fem.jl).
//
// Compiled with -fmad=false: these kernels touch n (= 32) doubles, so their cost is pure
// launch latency, and without contraction the point-wise updates (compute_b_from_e!,
// e .-= work, ...) reproduce the reference's fp64 expressions bit for bit.
//
// The reference applies each circulant operator as irfft(rfft(x) .* eig) with FFTW
// (solve_circulant!, :222-240).  Here the operator is applied as the equivalent periodic
// convolution y[i] = sum_j col[(i-j) mod n] x[j] with the first column `col` precomputed
// at construction from the same eigenvalue tables -- O(n^2) = 1024 FMAs for n = 32, one
// block, no FFT library, no CPU round trip.
#include <cmath>

#include "objects.cuh"

namespace gempic {

constexpr int kFieldThreads = 256;

// y[i] = sum_j col[(i-j) mod n] * x[j]  for i = tid, tid+blockDim, ...
__device__ __forceinline__ double circ_row(const double *__restrict__ col, const double *__restrict__ x, int i, int n)
{
    double acc = 0.0;
    int k = i;  // (i - j) mod n, starts at i for j = 0 and decreases
    for (int j = 0; j < n; ++j) {
        acc += col[k] * x[j];
        k = (k == 0) ? n - 1 : k - 1;
    }
    return acc;
}

// compute_e_from_rho! (:244-255): phi = circ(weak_poisson, rho); e[i] = phi[i-1] - phi[i]
__global__ void k_e_from_rho(double *__restrict__ e, const double *__restrict__ col, const double *__restrict__ rho, int n)
{
    extern __shared__ double phi[];
    for (int i = threadIdx.x; i < n; i += blockDim.x) phi[i] = circ_row(col, rho, i, n);
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) e[i] = phi[i == 0 ? n - 1 : i - 1] - phi[i];
}

// The three field updates of a splitting step as block-wide device functions (one block, `sh` = n doubles of shared
// scratch, callers synchronise before and after), shared by the stand-alone kernels and the fused per-step kernel below
// so that both perform the same fp64 operations in the same order.
//
// compute_e_from_j! (:263-289) with the optional `j .= j .* prescale` that precedes it in
// operatorHp2 (hamiltonian_splitting_1d2v.jl:173) and Boris (:160-161)
__device__ __forceinline__ void dev_e_from_j(double *e, const double *__restrict__ col, double *j, int n, double dx,
                                             double prescale, int do_prescale, double *sh)
{
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double v = j[i];
        if (do_prescale) {
            v = v * prescale;
            j[i] = v;
        }
        sh[i] = v;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double work = circ_row(col, sh, i, n);
        work = work / dx;
        e[i] = e[i] - work;
    }
    __syncthreads();
}
// compute_e_from_b! (:384-396)
__device__ __forceinline__ void dev_e_from_b(double *e, const double *__restrict__ col, const double *b, int n, double coef,
                                             double *sh)
{
    for (int i = threadIdx.x; i < n; i += blockDim.x) sh[i] = b[i];
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) e[i] = e[i] + coef * circ_row(col, sh, i, n);
    __syncthreads();
}
// compute_b_from_e! (:407-420)
__device__ __forceinline__ void dev_b_from_e(double *b, const double *e, int n, double coef)
{
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double prev = e[i == 0 ? n - 1 : i - 1];
        b[i] = b[i] + coef * (prev - e[i]);
    }
    __syncthreads();
}

__global__ void k_e_from_j(double *e, const double *__restrict__ col, double *j, int n, double dx, double prescale,
                           int do_prescale)
{
    extern __shared__ double sj[];
    dev_e_from_j(e, col, j, n, dx, prescale, do_prescale, sj);
}
__global__ void k_e_from_b(double *e, const double *__restrict__ col, const double *b, int n, double coef)
{
    extern __shared__ double sb[];
    dev_e_from_b(e, col, b, n, coef, sb);
}
__global__ void k_b_from_e(double *b, const double *e, int n, double coef) { dev_b_from_e(b, e, n, coef); }

// Every field-only update between two fused particle passes of strang_splitting! (hs1d.cu strang_fused) in ONE launch
// instead of nine 3 us ones:
//   solve   e2 -= M0^-1 (j2 * j2_scale) / dx ; e1 -= M1^-1 j1 / dx ; j_dofs[1] = 0        (Hp2, Hp1, Hp2 of the pass)
//   tail    eT = (e1, e2) ; b += dt_tail/dx D e2 ; e2 += dt_tail/dx A b                  (trailing HE field part, HB)
//   lead    e2 += dt_lead/dx A b ; b += dt_lead/dx D e2                                   (leading HB, HE field part)
__global__ void k_strang_fields(StrangFields F)
{
    // Everything lives in shared memory between the first load and the final store: the phases below would otherwise be
    // a chain of ~10 dependent global round trips.  Layout (n doubles each): scratch | e1 | e2 | b | acc (2n) | eT (2n) |
    // inv_mass0 | inv_mass1 | ampere
    extern __shared__ double sh[];
    const int n = F.n;
    double *e1 = sh + n, *e2 = e1 + n, *b = e2 + n, *acc = b + n, *eT = acc + 2 * n;
    double *c0 = eT + 2 * n, *c1 = c0 + n, *ca = c1 + n;
    if (F.n_partials >= 0) {
        // acc[g] = sum_b partials[b][g].  Thread t owns output t % n_acc and every (blockDim / n_acc)-th block row, so
        // all its loads are independent and in flight together (coalesced along g); the few row-group sums of an
        // output are then added in a fixed order.  (n_acc = 2n <= 256 threads; red = 4 * n_acc doubles after the tables)
        double *red = ca + n;
        const int groups = blockDim.x / F.n_acc;
        const int o = threadIdx.x % F.n_acc, grp = threadIdx.x / F.n_acc;
        if (grp < groups) {
            double sum = 0.0;
            for (int k = grp; k < F.n_partials; k += groups) sum += F.partials[(size_t)k * F.n_acc + o];
            red[grp * F.n_acc + o] = sum;
        }
        __syncthreads();
        for (int g = threadIdx.x; g < F.n_acc; g += blockDim.x) {
            double sum = 0.0;
            for (int q = 0; q < groups; ++q) sum += red[q * F.n_acc + g];
            acc[g] = sum;
        }
        if (F.x.n_ranks > 1) {   // sum over the ranks through peer memory (rank order: identical on every rank)
            __syncthreads();
            xchg_allreduce_block(F.x, acc, F.n_acc, 0, 0);
        }
    } else if (F.do_solve) {
        for (int i = threadIdx.x; i < 2 * n; i += blockDim.x) acc[i] = F.acc[i];
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        e1[i] = F.e1[i];
        e2[i] = F.e2[i];
        b[i] = F.b[i];
        c0[i] = F.inv_mass0[i];
        c1[i] = F.inv_mass1[i];
        ca[i] = F.ampere[i];
    }
    __syncthreads();
    if (F.do_solve) {
        dev_e_from_j(e2, c0, acc, n, F.dx, F.j2_scale, F.j2_scale != 1.0 ? 1 : 0, sh);
        dev_e_from_j(e1, c1, acc + n, n, F.dx, 1.0, 0, sh);
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            F.j1[i] = 0.0;
            F.acc[i] = acc[i];           // dt/2 (j2a + j2b), kept for inspection
            F.acc[n + i] = acc[n + i];
        }
    }
    if (F.do_tail) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            eT[i] = e1[i];
            eT[n + i] = e2[i];
        }
        __syncthreads();
        dev_b_from_e(b, e2, n, F.dt_tail / F.dx);
        dev_e_from_b(e2, ca, b, n, F.dt_tail / F.dx, sh);
        for (int i = threadIdx.x; i < 2 * n; i += blockDim.x) F.eT[i] = eT[i];
    }
    if (F.do_lead) {
        dev_e_from_b(e2, ca, b, n, F.dt_lead / F.dx, sh);
        dev_b_from_e(b, e2, n, F.dt_lead / F.dx);
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        F.e1[i] = e1[i];
        F.e2[i] = e2[i];
        F.b[i] = b[i];
    }
}

// inner_product (:461-475): sum(c1 .* circ(mass, c2)) * dx, summed in index order by one thread
__global__ void k_inner_product(const double *__restrict__ c1, const double *__restrict__ c2,
                                const double *__restrict__ col, int n, double dx, double *__restrict__ out)
{
    extern __shared__ double sm[];
    double *sx = sm, *sw = sm + n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) sx[i] = c2[i];
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) sw[i] = c1[i] * circ_row(col, sx, i, n);
    __syncthreads();
    if (threadIdx.x == 0) {
        double r = 0.0;
        for (int i = 0; i < n; ++i) r += sw[i];
        out[0] = r * dx;
    }
}

__global__ void k_axpby(double *__restrict__ y, double a, const double *__restrict__ x, double b, int n)
{
    for (int i = threadIdx.x; i < n; i += blockDim.x) y[i] = a * x[i] + b * y[i];
}
__global__ void k_copy(double *__restrict__ dst, const double *__restrict__ src, int n)
{
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
}
__global__ void k_max_abs_diff(const double *__restrict__ a, const double *__restrict__ b, int n, double *out)
{
    __shared__ double red[kFieldThreads];
    double m = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) m = fmax(m, fabs(a[i] - b[i]));
    red[threadIdx.x] = m;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + s]);
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = red[0];
}

// out[g] = sum_b partials[b][g]; one warp per dof, lanes stride over blocks, fixed shuffle tree
__global__ void k_reduce_partials(const double *__restrict__ partials, int n_blocks, int n_acc, double *__restrict__ out)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= n_acc) return;
    double s = 0.0;
    for (int b = lane; b < n_blocks; b += 32) s += partials[(size_t)b * n_acc + warp];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off);
    if (lane == 0) out[warp] = s;
}

static inline int field_threads(int n) { return n < kFieldThreads ? ((n + 31) / 32) * 32 : kFieldThreads; }

void field_e_from_rho(const Maxwell1D &m, double *e, const double *rho)
{
    k_e_from_rho<<<1, field_threads(m.n), m.n * sizeof(double), ctx().stream>>>(e, m.col(Maxwell1D::C_POISSON), rho, m.n);
    GP_CUDA(cudaGetLastError());
    count_launch();
}

void field_e_from_j(const Maxwell1D &m, double *e, double *j, int component, double prescale)
{
    GP_REQUIRE(component == 1 || component == 2, GEMPIC_EINVAL, "Component %d not implemented", component);
    const double *col = m.col(component == 1 ? Maxwell1D::C_INV_MASS1 : Maxwell1D::C_INV_MASS0);
    k_e_from_j<<<1, field_threads(m.n), m.n * sizeof(double), ctx().stream>>>(e, col, j, m.n, m.delta_x, prescale,
                                                                              prescale != 1.0 ? 1 : 0);
    GP_CUDA(cudaGetLastError());
    count_launch();
}

// The field-only part of HamiltonianSplittingBoris.strang_splitting! (hamiltonian_splitting_boris.jl:132-177) around the
// one-pass particle step, one launch:
//   post  (4) of the step just pushed: e = e_mid ; e1_mid -= M1^-1 (j1 dt)/dx ; e2_mid += dt/dx A b ; e2_mid -= M0^-1 (j2 dt)/dx
//   pre   (1) of the next step: b_mid = b ; b += dt/dx D e2_mid ; b_mid = (b_mid + b) * 0.5
__global__ void k_boris_fields(BorisFields F)
{
    extern __shared__ double sh[];
    const int n = F.n;
    if (F.n_partials >= 0) {
        // j1 | j2 (adjacent) = sum over the block rows of the pass, like k_strang_fields: thread t owns output t % 2n and
        // every 4th row, then the four row-group sums are added in a fixed order
        double *red = sh + n;
        const int n_acc = 2 * n, groups = blockDim.x / n_acc;
        const int o = threadIdx.x % n_acc, grp = threadIdx.x / n_acc;
        if (grp < groups) {
            double sum = 0.0;
            for (int k = grp; k < F.n_partials; k += groups) sum += F.partials[(size_t)k * n_acc + o];
            red[grp * n_acc + o] = sum;
        }
        __syncthreads();
        for (int g = threadIdx.x; g < n_acc; g += blockDim.x) {
            double sum = 0.0;
            for (int q = 0; q < groups; ++q) sum += red[q * n_acc + g];
            F.j1[g] = sum;   // j1 | j2 are adjacent
        }
        __syncthreads();
        if (F.x.n_ranks > 1) xchg_allreduce_block(F.x, F.j1, n_acc, 0, 0);
    }
    if (F.do_post) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            F.e1[i] = F.e1_mid[i];
            F.e2[i] = F.e2_mid[i];
        }
        __syncthreads();
        dev_e_from_j(F.e1_mid, F.inv_mass1, F.j1, n, F.dx, F.dt_post, F.dt_post != 1.0 ? 1 : 0, sh);
        dev_e_from_b(F.e2_mid, F.ampere, F.b, n, F.dt_post / F.dx, sh);
        dev_e_from_j(F.e2_mid, F.inv_mass0, F.j2, n, F.dx, F.dt_post, F.dt_post != 1.0 ? 1 : 0, sh);
    }
    if (F.do_pre) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) F.b_mid[i] = F.b[i];
        __syncthreads();
        dev_b_from_e(F.b, F.e2_mid, n, F.dt_pre / F.dx);
        for (int i = threadIdx.x; i < n; i += blockDim.x) F.b_mid[i] = (F.b_mid[i] + F.b[i]) * 0.5;
    }
}

void field_boris_fields(const Maxwell1D &m, BorisFields F)
{
    F.inv_mass0 = m.col(Maxwell1D::C_INV_MASS0);
    F.inv_mass1 = m.col(Maxwell1D::C_INV_MASS1);
    F.ampere = m.col(Maxwell1D::C_AMPERE);
    F.n = m.n;
    F.dx = m.delta_x;
    const bool red = F.n_partials >= 0;
    const int threads = red ? 8 * m.n : field_threads(m.n);
    GP_REQUIRE(threads <= 1024, GEMPIC_EINVAL, "fused field kernel: grid of %d dofs is too large", m.n);
    k_boris_fields<<<1, threads, (red ? 9 : 1) * (size_t)m.n * sizeof(double), ctx().stream>>>(F);
    GP_CUDA(cudaGetLastError());
    count_launch();
}

void field_strang_fields(const Maxwell1D &m, StrangFields F)
{
    F.inv_mass0 = m.col(Maxwell1D::C_INV_MASS0);
    F.inv_mass1 = m.col(Maxwell1D::C_INV_MASS1);
    F.ampere = m.col(Maxwell1D::C_AMPERE);
    F.n = m.n;
    F.dx = m.delta_x;
    // with partials to reduce: a multiple of n_acc = 2n threads (n <= 64 for the fused path), 4 row groups
    const int threads = F.n_partials >= 0 ? 4 * F.n_acc : field_threads(m.n);
    GP_REQUIRE(threads <= 1024, GEMPIC_EINVAL, "fused field kernel: grid of %d dofs is too large", m.n);
    k_strang_fields<<<1, threads, (12 * (size_t)m.n + 4 * (size_t)(F.n_partials >= 0 ? F.n_acc : 0)) * sizeof(double),
                      ctx().stream>>>(F);
    GP_CUDA(cudaGetLastError());
    count_launch();
}

void field_e_from_b(const Maxwell1D &m, double *e, double dt, const double *b)
{
    const double coef = dt / m.delta_x;
    k_e_from_b<<<1, field_threads(m.n), m.n * sizeof(double), ctx().stream>>>(e, m.col(Maxwell1D::C_AMPERE), b, m.n, coef);
    GP_CUDA(cudaGetLastError());
    count_launch();
}

void field_b_from_e(const Maxwell1D &m, double *b, double dt, const double *e)
{
    const double coef = dt / m.delta_x;
    k_b_from_e<<<1, field_threads(m.n), 0, ctx().stream>>>(b, e, m.n, coef);
    GP_CUDA(cudaGetLastError());
    count_launch();
}

void field_inner_product(const Maxwell1D &m, const double *c1, const double *c2, int degree, double *out)
{
    // reference: degree == s_deg_0 -> eig_mass0, degree == s_deg_1 -> eig_mass1, anything else
    // silently reuses a stale self.work (:463-468); the replacement rejects it instead.
    GP_REQUIRE(degree == m.s_deg_0 || degree == m.s_deg_1, GEMPIC_EINVAL, "degree %d not available", degree);
    const double *col = m.col(degree == m.s_deg_0 ? Maxwell1D::C_MASS0 : Maxwell1D::C_MASS1);
    k_inner_product<<<1, field_threads(m.n), 2 * m.n * sizeof(double), ctx().stream>>>(c1, c2, col, m.n, m.delta_x, out);
    GP_CUDA(cudaGetLastError());
    count_launch();
}

void field_axpby(double *y, double a, const double *x, double b, int n)
{
    k_axpby<<<1, field_threads(n), 0, ctx().stream>>>(y, a, x, b, n);
    GP_CUDA(cudaGetLastError());
    count_launch();
}
void field_copy(double *dst, const double *src, int n)
{
    k_copy<<<1, field_threads(n), 0, ctx().stream>>>(dst, src, n);
    GP_CUDA(cudaGetLastError());
    count_launch();
}
void field_max_abs_diff(const double *a, const double *b, int n, double *out)
{
    k_max_abs_diff<<<1, kFieldThreads, 0, ctx().stream>>>(a, b, n, out);
    GP_CUDA(cudaGetLastError());
    count_launch();
}

// ---- construction: eigenvalue tables (:49-177) and circulant first columns ---------------
static void first_column(const std::vector<double> &eig, int n, double *col)
{
    // y = irfft(rfft(x) .* lambda), lambda_k = eig[k] + i*eig[n-k] (0<k<n/2), lambda_0, lambda_{n/2} real
    // => col[m] = (1/n) [lambda_0 + (-1)^m lambda_{n/2} + 2 sum_k (Re cos(2 pi m k/n) - Im sin(2 pi m k/n))]
    const long double two_pi = 6.283185307179586476925286766559005768L;
    for (int mm = 0; mm < n; ++mm) {
        long double acc = (long double)eig[0] + ((mm & 1) ? -(long double)eig[n / 2] : (long double)eig[n / 2]);
        for (int k = 1; k < n / 2; ++k) {
            const int t = (int)(((long long)mm * k) % n);
            const long double ang = two_pi * t / n;
            acc += 2.0L * ((long double)eig[k] * cosl(ang) - (long double)eig[n - k] * sinl(ang));
        }
        col[mm] = (double)(acc / n);
    }
}

std::unique_ptr<Maxwell1D> make_maxwell1d(double xmin, double xmax, int n_dofs, int degree)
{
    GP_REQUIRE(degree >= 1 && degree <= 3, GEMPIC_EINVAL, "Wrong value of degree = %d  (1,2 or 3)", degree);
    GP_REQUIRE(n_dofs >= 2 && n_dofs % 2 == 0, GEMPIC_EINVAL, "n_dofs = %d must be even and >= 2", n_dofs);
    GP_REQUIRE(n_dofs <= 8192, GEMPIC_EINVAL, "n_dofs = %d exceeds the supported 8192", n_dofs);
    auto m = std::make_unique<Maxwell1D>();
    m->xmin = xmin;
    m->n = n_dofs;
    m->Lx = xmax - xmin;
    m->delta_x = m->Lx / n_dofs;
    m->s_deg_0 = degree;
    m->s_deg_1 = degree - 1;
    double mass_0[4] = {0, 0, 0, 0}, mass_1[4] = {0, 0, 0, 0};
    if (degree == 1) {
        mass_0[0] = 4.0 / 6.0; mass_0[1] = 1.0 / 6.0;
        mass_1[0] = 1.0;
    } else if (degree == 2) {
        mass_0[0] = 66.0 / 120.0; mass_0[1] = 26.0 / 120.0; mass_0[2] = 1.0 / 120.0;
        mass_1[0] = 4.0 / 6.0; mass_1[1] = 1.0 / 6.0;
    } else {
        mass_0[0] = 2416.0 / 5040.0; mass_0[1] = 1191.0 / 5040.0; mass_0[2] = 120.0 / 5040.0; mass_0[3] = 1.0 / 5040.0;
        mass_1[0] = 66.0 / 120.0; mass_1[1] = 26.0 / 120.0; mass_1[2] = 1.0 / 120.0;
    }
    const int n = n_dofs, s = degree;
    const double pi = 3.14159265358979323846;
    auto &e0 = m->eig_mass0, &e1 = m->eig_mass1, &ea = m->eig_weak_ampere, &ep = m->eig_weak_poisson;
    e0.assign(n, 0.0); e1.assign(n, 0.0); ea.assign(n, 0.0); ep.assign(n, 0.0);
    e0[0] = 1.0;
    e1[0] = 1.0;
    for (int k = 1; k <= n / 2 - 1; ++k) {
        double coef0 = mass_0[0], coef1 = mass_1[0];
        for (int j = 1; j <= s - 1; ++j) {
            const double cos_mode = std::cos(2 * pi * j * k / n);
            coef0 = coef0 + 2 * mass_0[j] * cos_mode;
            coef1 = coef1 + 2 * mass_1[j] * cos_mode;
        }
        coef0 = coef0 + 2 * mass_0[s] * std::cos(2 * pi * s * k / n);
        e0[k] = coef0;
        e1[k] = coef1;
        const double cos_mode = std::cos(2 * pi * k / n), sin_mode = std::sin(2 * pi * k / n);
        ea[k] = (coef1 / coef0) * (1 - cos_mode);
        ea[n - k] = -(coef1 / coef0) * sin_mode;
        ep[k] = 1.0 / (coef1 * ((1 - cos_mode) * (1 - cos_mode) + sin_mode * sin_mode));
    }
    double coef0 = mass_0[0], coef1 = mass_1[0];
    for (int j = 1; j <= s - 1; ++j) {
        coef0 = coef0 + 2 * mass_0[j] * std::cos(pi * j);
        coef1 = coef1 + 2 * mass_1[j] * std::cos(pi * j);
    }
    coef0 = coef0 + 2 * mass_0[s] * std::cos(pi * s);
    e0[n / 2] = coef0;
    e1[n / 2] = coef1;
    ea[n / 2] = 2.0 * (coef1 / coef0);
    ep[n / 2] = 1.0 / (coef1 * 4.0);

    // inverse mass tables exactly as compute_e_from_j! builds them (:272-280): 1/eig for k <= n/2
    std::vector<double> inv0(n, 0.0), inv1(n, 0.0);
    for (int i = 0; i <= n / 2; ++i) {
        inv0[i] = 1.0 / e0[i];
        inv1[i] = 1.0 / e1[i];
    }
    std::vector<double> cols((size_t)Maxwell1D::C_COUNT * n);
    first_column(e0, n, &cols[(size_t)Maxwell1D::C_MASS0 * n]);
    first_column(e1, n, &cols[(size_t)Maxwell1D::C_MASS1 * n]);
    first_column(inv0, n, &cols[(size_t)Maxwell1D::C_INV_MASS0 * n]);
    first_column(inv1, n, &cols[(size_t)Maxwell1D::C_INV_MASS1 * n]);
    first_column(ea, n, &cols[(size_t)Maxwell1D::C_AMPERE * n]);
    first_column(ep, n, &cols[(size_t)Maxwell1D::C_POISSON * n]);
    m->cols.alloc(cols.size());
    GP_CUDA(cudaMemcpyAsync(m->cols.p, cols.data(), cols.size() * sizeof(double), cudaMemcpyHostToDevice, ctx().stream));
    GP_CUDA(cudaStreamSynchronize(ctx().stream));
    m->tmp.alloc((size_t)4 * n + 8);
    return m;
}

}  // namespace gempic
