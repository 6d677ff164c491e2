// hostutil.hpp -- host-side helpers shared by the extern "C" entry points: staging of small host
// vectors, Gauss-Legendre nodes and the de Boor recurrence for the set-up quadratures (cold path).
#pragma once
#include <cmath>

#include "common.cuh"

namespace gempic {

// host vectors of n doubles staged into one scratch allocation.  The scratch is the rank's grow-only buffer
// (Context::stage): a cudaMalloc + cudaFree per API call costs more than the small kernels it serves and synchronises
// the device (the diagnostics loop calls write_step! every step).  A second Stage alive at the same time gets its own.
struct Stage {
    DevBuf<double> own;
    double *base = nullptr;
    size_t used = 0;
    bool shared = false;
    explicit Stage(size_t total)
    {
        Context &c = ctx();
        if (!total) total = 1;
        if (!c.stage_busy && total <= ((size_t)1 << 22)) {   // up to 32 MB is kept; per-particle arrays come and go
            if (c.stage_n < total) {
                if (c.stage) cudaFree(c.stage);   // synchronises: nothing in flight still reads it
                c.stage = nullptr;
                c.stage_n = 0;
                const size_t n = total < 4096 ? 4096 : total;
                GP_CUDA(cudaMalloc(&c.stage, n * sizeof(double)));
                c.stage_n = n;
            }
            base = c.stage;
            shared = true;
            c.stage_busy = true;
        } else {
            own.alloc(total);
            base = own.p;
        }
    }
    Stage(const Stage &) = delete;
    Stage &operator=(const Stage &) = delete;
    ~Stage()
    {
        if (shared) ctx().stage_busy = false;
    }
    double *take(size_t n)
    {
        double *p = base + used;
        used += n;
        return p;
    }
    double *put(const double *host, size_t n)
    {
        double *p = take(n);
        if (n) h2d(p, host, n);
        return p;
    }
};

inline void legendre_nodes(int n, double *x, double *w)
{
    // Gauss-Legendre on [-1,1] (FastGaussQuadrature.gausslegendre), Newton on P_n
    const long double pi = 3.141592653589793238462643383279502884L;
    for (int i = 0; i < n; ++i) {
        long double z = cosl(pi * (i + 0.75L) / (n + 0.5L)), pp = 1.0L;
        for (int it = 0; it < 100; ++it) {
            long double p1 = 1.0L, p2 = 0.0L;
            for (int j = 1; j <= n; ++j) {
                const long double p3 = p2;
                p2 = p1;
                p1 = ((2.0L * j - 1.0L) * z * p2 - (j - 1.0L) * p3) / j;
            }
            pp = n * (z * p1 - p2) / (z * z - 1.0L);
            const long double z1 = z;
            z = z1 - p1 / pp;
            if (fabsl(z - z1) < 1e-19L) break;
        }
        x[n - 1 - i] = (double)z;
        w[n - 1 - i] = (double)(2.0L / ((1.0L - z * z) * pp * pp));
    }
    if (n % 2 == 1) x[n / 2] = 0.0;
}

inline void host_bsplines(int degree, double offset, double *b)
{
    b[0] = 1.0;
    for (int j = 1; j <= degree; ++j) {
        double xx = -offset, saved = 0.0;
        const double jr = (double)j, inv_j = 1.0 / jr;
        for (int r = 0; r < j; ++r) {
            xx = xx + 1.0;
            const double temp = b[r] * inv_j;
            b[r] = saved + xx * temp;
            saved = (jr - xx) * temp;
        }
        b[j] = saved;
    }
}


}  // namespace gempic
