// sampling.cu -- the reference's particle samplers on the device:
//   sample!(pg::ParticleGroup{1,1} / {1,2}, alpha, k, sigma, mesh)   src/particle_sampling.jl:266-311
//   sample!(d::LandauDamping, pg)                                     src/landau_damping.jl:34-59
//       v_i = sigma sqrt(-2 log((i - 1/2)/N)), (r1, r2) = Sobol(2), theta = 2 pi r1, x = newton(r2): deterministic.
//   sample!(pg, ps::ParticleSampler, df::AbstractCosGaussian, mesh)   src/particle_sampling.jl:68-225
//       sample_all (:89-143) and the 8-fold antithetic sample_sym (:150-225), sampling_type :random / :sobol.
//
// Sobol.jl's SobolSeq(N) is the Gray-code (Antonov-Saleev) Sobol sequence with the Joe-Kuo direction numbers
// (new-joe-kuo-6.21201), whose first next!() returns point number 1 (the all-zero point 0 is skipped).  Point n is
//     x_n = XOR_{c : bit c of gray(n)} m_c 2^-(c+1),   gray(n) = n ^ (n >> 1),
// which every thread evaluates directly from the global particle index: the load is identical for any sharding of the
// index range, and bit-identical to Sobol.jl (the coordinates are dyadic rationals; no rounding is involved).
//
// Julia's MersenneTwister + ziggurat normals (rand!(rng, Normal(), v), :129,:198) cannot be reproduced outside Julia;
// the normal deviates and the :random uniforms come from a counter-based generator instead (splitmix64 of seed, index,
// stream; Box-Muller) -- statistical parity (test/test_sampling.jl:43-123), exact antithetic structure.
#include <cmath>

#include "objects.cuh"

namespace gempic {

constexpr int kSobolBits = 32;   // Sobol.jl keeps 32-bit direction numbers: indices below 2^32
constexpr int kSobolDims = 4;
constexpr int kMaxCos = 8, kMaxGauss = 8;

struct SobolTable {
    uint32_t v[kSobolDims][kSobolBits];   // direction numbers scaled to 32 bits: v_c = m_c << (31 - c)
};

// Joe-Kuo primitive polynomials (degree s, coefficient bits a) and initial m for dimensions 2..4; dimension 1 is the
// van der Corput sequence (m_c = 1).
static SobolTable make_sobol_table()
{
    SobolTable T{};
    const int s_[kSobolDims] = {0, 1, 2, 3};
    const uint32_t a_[kSobolDims] = {0, 0, 1, 1};
    const uint32_t m0[kSobolDims][3] = {{0, 0, 0}, {1, 0, 0}, {1, 3, 0}, {1, 3, 1}};
    for (int d = 0; d < kSobolDims; ++d) {
        uint32_t m[kSobolBits];
        if (d == 0) {
            for (int c = 0; c < kSobolBits; ++c) m[c] = 1;
        } else {
            const int s = s_[d];
            for (int c = 0; c < s; ++c) m[c] = m0[d][c];
            for (int c = s; c < kSobolBits; ++c) {
                uint32_t val = m[c - s] ^ (m[c - s] << s);
                for (int k = 1; k < s; ++k)
                    if ((a_[d] >> (s - 1 - k)) & 1u) val ^= m[c - k] << k;
                m[c] = val;
            }
        }
        for (int c = 0; c < kSobolBits; ++c) T.v[d][c] = m[c] << (31 - c);
    }
    return T;
}

__device__ __forceinline__ double sobol_coord(const SobolTable &T, int dim, uint64_t n)
{
    uint32_t g = (uint32_t)(n ^ (n >> 1)), x = 0;
    while (g) {
        const int c = __ffs(g) - 1;
        x ^= T.v[dim][c];
        g &= g - 1;
    }
    return (double)x * (1.0 / 4294967296.0);
}

__device__ __forceinline__ uint64_t mix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ double uniform01(uint64_t seed, uint64_t idx, uint64_t stream)
{
    const uint64_t h = mix64(mix64(seed ^ (stream * 0xD1B54A32D192ED03ull)) + idx);
    return ((double)(h >> 11) + 0.5) * (1.0 / 9007199254740992.0);  // (0,1)
}

// newton(r, alpha, k) of particle_sampling.jl:236-245 (and landau_damping.jl:37-46), iteration for iteration: two
// interleaved Newton sequences started at 0 and at 1, stopped when they agree to 1e-12.
__device__ __forceinline__ double newton_reference(double r, double alpha, double k)
{
    double x0 = 0.0, x1 = 1.0;
    r *= 6.283185307179586 / k;
    int guard = 0;
    while (fabs(x1 - x0) > 1e-12 && ++guard < 200) {
        const double p = x0 + alpha * sin(k * x0) / k;
        const double f = 1.0 + alpha * cos(k * x0);
        const double nx = x0 - (p - r) / f;
        x0 = x1;
        x1 = nx;
    }
    return x1;
}

struct LandauParams {
    double alpha, k, sigma, weight;
    int64_t first, n_global;
};

__global__ void k_sample_landau(double *__restrict__ data, size_t stride, int V, int64_t n, LandauParams s, SobolTable T)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t g = (uint64_t)(s.first + i) + 1;   // 1-based particle number = Sobol point number
        const double v = s.sigma * sqrt(-2.0 * log(((double)g - 0.5) / (double)s.n_global));
        const double r1 = sobol_coord(T, 0, g), r2 = sobol_coord(T, 1, g);
        const double theta = r1 * 6.283185307179586;
        data[i] = newton_reference(r2, s.alpha, s.k);
        data[stride + i] = v * cos(theta);
        if (V == 2) data[2 * stride + i] = v * sin(theta);
        data[(size_t)(1 + V) * stride + i] = s.weight;
    }
}

struct CosGaussParams {
    int sampling_type, symmetric, n_cos, n_gauss;
    double xmin, dimx;
    double k[kMaxCos], alpha[kMaxCos];
    double sigma[kMaxGauss][2], mu[kMaxGauss][2], delta_cum[kMaxGauss];
    uint64_t seed;
    int64_t first;
};

// one thread per DRAW: a draw is one particle (sample_all) or one group of 8 antithetic particles (sample_sym)
__global__ void k_sample_cos_gauss(double *__restrict__ data, size_t stride, int64_t n, CosGaussParams s, SobolTable T)
{
    const int per = s.symmetric ? 8 : 1;
    const int64_t n_draws = (n + per - 1) / per;
    const double two_pi = 6.283185307179586476925286766559;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n_draws; q += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t draw = (uint64_t)(s.first / per + q) + 1;   // 1-based draw number = Sobol point number
        double rx, rsel;
        if (s.sampling_type == 1) {
            rx = sobol_coord(T, 0, draw);                       // SobolSeq(ndx) (:107) / first coordinate of SobolSeq(4) (:167)
            rsel = s.symmetric ? sobol_coord(T, 3, draw) : 0.0;  // rdn[ndx+ndv+1] (:204); sample_all never fills rdn (:132)
        } else {
            rx = uniform01(s.seed, draw, 1);
            rsel = s.symmetric ? uniform01(s.seed, draw, 4) : 0.0;
        }
        double x = s.xmin + s.dimx * rx;
        double dens = 1.0;                                      // eval_x_density (distributions.jl:159-178)
        for (int j = 0; j < s.n_cos; ++j) dens += s.alpha[j] * cos(s.k[j] * x);
        const double w = dens * s.dimx;
        int ig = 0;
        while (ig < s.n_gauss - 1 && rsel > s.delta_cum[ig]) ++ig;
        const double u1 = uniform01(s.seed, draw, 20), u2 = uniform01(s.seed, draw, 21);
        const double rad = sqrt(-2.0 * log(u1));
        double v1 = rad * cos(two_pi * u2) * s.sigma[ig][0] + s.mu[ig][0];
        double v2 = rad * sin(two_pi * u2) * s.sigma[ig][1] + s.mu[ig][1];
        for (int m = 0; m < per; ++m) {
            const int64_t i = q * per + m;
            if (i >= n) break;
            // sample_sym :210-217: member 5 mirrors x, even members mirror v1, the other odd ones v2
            if (m == 4) x = s.dimx - x + 2.0 * s.xmin;
            else if (m > 0 && (m & 1)) v1 = -v1 + 2.0 * s.mu[ig][0];
            else if (m > 0) v2 = -v2 + 2.0 * s.mu[ig][1];
            data[i] = x;
            data[stride + i] = v1;
            data[2 * stride + i] = v2;
            data[3 * stride + i] = w;
        }
    }
}

}  // namespace gempic

using namespace gempic;

extern "C" {

int gempic_pg_sample_landau(gempic_handle h, double alpha, double k, double sigma, double weight, int64_t first_index,
                            int64_t n_global)
{
    GP_API_BEGIN
    require_init();
    ParticleGroup *pg = get_pg(h);
    GP_REQUIRE(pg->D == 1 && (pg->V == 1 || pg->V == 2), GEMPIC_EINVAL,
               "the Landau sampler is defined for ParticleGroup{1,1} and {1,2} (particle_sampling.jl:266-311)");
    GP_REQUIRE(k != 0.0, GEMPIC_EINVAL, "k must be non-zero");
    if (n_global <= 0) n_global = pg->n;
    GP_REQUIRE(first_index >= 0 && first_index + pg->n <= n_global, GEMPIC_EINVAL, "index range [%lld, %lld) outside [0, %lld)",
               (long long)first_index, (long long)(first_index + pg->n), (long long)n_global);
    GP_REQUIRE(n_global < ((int64_t)1 << kSobolBits), GEMPIC_EINVAL, "the Sobol sequence of Sobol.jl has 2^32 - 1 points");
    pg->sorted2d = false;
    particles_changed();
    if (pg->n == 0) return GEMPIC_OK;
    static const SobolTable T = make_sobol_table();
    LandauParams s{alpha, k, sigma, weight, first_index, n_global};
    k_sample_landau<<<ctx().sm_count * 8, 256, 0, ctx().stream>>>(pg->data.p, pg->stride, pg->V, pg->n, s, T);
    GP_CUDA(cudaGetLastError());
    count_launch();
    GP_API_END
}

int gempic_pg_sample_cos_gaussian(gempic_handle h, int sampling_type, int symmetric, uint64_t seed, double xmin, double dimx,
                                  int n_cos, const double *k, const double *alpha, int n_gaussians, const double *sigma,
                                  const double *mu, const double *delta, int64_t first_index)
{
    GP_API_BEGIN
    require_init();
    ParticleGroup *pg = get_pg(h);
    GP_REQUIRE(sampling_type == 0 || sampling_type == 1, GEMPIC_EINVAL, "Sampling type %d not implemented", sampling_type);   // :24-27
    GP_REQUIRE(pg->D == 1 && pg->V == 2, GEMPIC_EINVAL, "sample! is defined for ParticleGroup{1,2} (particle_sampling.jl:68-77)");
    GP_REQUIRE(n_cos >= 0 && n_cos <= kMaxCos && (n_cos == 0 || (k && alpha)), GEMPIC_EINVAL, "0..%d cosines", kMaxCos);
    GP_REQUIRE(n_gaussians >= 1 && n_gaussians <= kMaxGauss && sigma && mu, GEMPIC_EINVAL, "1..%d Gaussians", kMaxGauss);
    GP_REQUIRE(n_gaussians == 1 || delta, GEMPIC_EINVAL, "delta is required for more than one Gaussian");
    GP_REQUIRE(dimx > 0.0, GEMPIC_EINVAL, "domain length must be positive");
    const int per = symmetric ? 8 : 1;
    GP_REQUIRE(first_index >= 0 && first_index % per == 0, GEMPIC_EINVAL, "first_index must be a multiple of %d", per);
    GP_REQUIRE((first_index + pg->n) / per + 1 < ((int64_t)1 << kSobolBits), GEMPIC_EINVAL, "the Sobol sequence of Sobol.jl has 2^32 - 1 points");
    CosGaussParams s{};
    s.sampling_type = sampling_type; s.symmetric = symmetric ? 1 : 0; s.n_cos = n_cos; s.n_gauss = n_gaussians;
    s.xmin = xmin; s.dimx = dimx; s.seed = seed; s.first = first_index;
    for (int j = 0; j < n_cos; ++j) { s.k[j] = k[j]; s.alpha[j] = alpha[j]; }
    double cum = 0.0;
    for (int j = 0; j < n_gaussians; ++j) {
        GP_REQUIRE(sigma[2 * j] != 0.0 && sigma[2 * j + 1] != 0.0, GEMPIC_EASSERT, "all(sigma .!= 0.0) (distributions.jl:38)");
        s.sigma[j][0] = sigma[2 * j]; s.sigma[j][1] = sigma[2 * j + 1];
        s.mu[j][0] = mu[2 * j]; s.mu[j][1] = mu[2 * j + 1];
        cum += delta ? delta[j] : 1.0;
        s.delta_cum[j] = cum;                                 // :100-103, :160-163
    }
    GP_REQUIRE(cum == 1.0, GEMPIC_EASSERT, "sum(delta) == 1.0 (distributions.jl:44)");
    pg->sorted2d = false;
    particles_changed();
    if (pg->n == 0) return GEMPIC_OK;
    static const SobolTable T = make_sobol_table();
    k_sample_cos_gauss<<<ctx().sm_count * 8, 256, 0, ctx().stream>>>(pg->data.p, pg->stride, pg->n, s, T);
    GP_CUDA(cudaGetLastError());
    count_launch();
    GP_API_END
}

/* test hook: the first `n` points of SobolSeq(dims) as next!() returns them (row-major n x dims), dims <= 4 */
int gempic_sobol_points(int dims, int64_t first, int64_t n, double *out)
{
    GP_API_BEGIN
    GP_REQUIRE(dims >= 1 && dims <= kSobolDims && n >= 0 && first >= 0 && out, GEMPIC_EINVAL, "bad arguments");
    const SobolTable T = make_sobol_table();   // host evaluation of the same table the kernels use
    for (int64_t i = 0; i < n; ++i) {
        const uint64_t p = (uint64_t)(first + i) + 1;
        for (int d = 0; d < dims; ++d) {
            uint32_t g = (uint32_t)(p ^ (p >> 1)), x = 0;
            for (int c = 0; g; ++c, g >>= 1)
                if (g & 1u) x ^= T.v[d][c];
            out[i * dims + d] = (double)x * (1.0 / 4294967296.0);
        }
    }
    GP_API_END
}

}  // extern "C"
