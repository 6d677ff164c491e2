// pmc2d.cu -- ParticleMeshCoupling2D (src/particle_mesh_coupling_2d.jl): tensor-product
// B-spline deposit / gather on a periodic nx x ny grid (x fastest, :76-82).
//
// compute_shape_factor uses the CEIL convention (:54-65): ip = ceil(xp), offset = xp-(ip-1)
// in (0,1], first dof index ip - degree - 1 (0-based).  Reproduced exactly.
//
// Round-1 implementation: one thread per particle, fp64 RED.ADD to a global grid for the
// deposit (native on sm_100a), read-only cached loads for the gather. The tiled/sorted
// shared-memory version is the "next" row of SURVEY section 8.
#include "objects.cuh"

namespace gempic {

struct Mesh2D {
    double xmin, ymin, dx, dy;
    int nx, ny;
};

__device__ __forceinline__ int wrap2(int g, int n)
{
    g %= n;
    return g < 0 ? g + n : g;
}

template <int D>
__device__ __forceinline__ void shape_factor(double xp, double yp, const Mesh2D &m, double (&vx)[D + 1],
                                             double (&vy)[D + 1], int &ix, int &iy)
{
    xp = (xp - m.xmin) / m.dx;
    yp = (yp - m.ymin) / m.dy;
    const int ip = __double2int_ru(xp), jp = __double2int_ru(yp);
    bspline_basis<D>(xp - (double)(ip - 1), vx);
    bspline_basis<D>(yp - (double)(jp - 1), vy);
    ix = wrap2(ip - D - 1, m.nx);   // (ip - degree) + i1 - 2 with i1 = 1
    iy = wrap2(jp - D - 1, m.ny);
}

template <int D>
__global__ void k_pmc2d_add_charge(const double *__restrict__ x, const double *__restrict__ y,
                                   const double *__restrict__ w, int64_t n, double wscale, double scaling, Mesh2D m,
                                   double *__restrict__ rho)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double vx[D + 1], vy[D + 1];
        int ix, iy;
        shape_factor<D>(x[i], y[i], m, vx, vy, ix, iy);
        const double wp = w[i] * wscale;
        int a = ix;
#pragma unroll
        for (int i1 = 0; i1 <= D; ++i1) {
            int b = iy;
#pragma unroll
            for (int i2 = 0; i2 <= D; ++i2) {
                atomicAdd(&rho[a + b * m.nx], wp * scaling * vx[i1] * vy[i2]);   // :102
                b = b + 1 == m.ny ? 0 : b + 1;
            }
            a = a + 1 == m.nx ? 0 : a + 1;
        }
    }
}

template <int D, int NFIELD>
__global__ void k_pmc2d_evaluate(const double *__restrict__ x, const double *__restrict__ y, int64_t n, Mesh2D m,
                                 const double *__restrict__ f1, const double *__restrict__ f2, double *__restrict__ o1,
                                 double *__restrict__ o2)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double vx[D + 1], vy[D + 1];
        int ix, iy;
        shape_factor<D>(x[i], y[i], m, vx, vy, ix, iy);
        double v1 = 0.0, v2 = 0.0;
        int a = ix;
#pragma unroll
        for (int i1 = 0; i1 <= D; ++i1) {
            int b = iy;
#pragma unroll
            for (int i2 = 0; i2 <= D; ++i2) {
                if (NFIELD == 1) {
                    v1 += f1[a + b * m.nx] * vx[i1] * vy[i2];   // :198
                } else {
                    const double c = vx[i1] * vy[i2];           // :224-226
                    v1 += f1[a + b * m.nx] * c;
                    v2 += f2[a + b * m.nx] * c;
                }
                b = b + 1 == m.ny ? 0 : b + 1;
            }
            a = a + 1 == m.nx ? 0 : a + 1;
        }
        o1[i] = v1;
        if (NFIELD == 2) o2[i] = v2;
    }
}

struct Pmc2D : Object {
    static constexpr Kind kKind = Kind::Pmc2D;
    Mesh2D m;
    int degree;
    double scaling;
    Pmc2D() : Object(kKind) {}
};

static int grid_for(int64_t n)
{
    const int64_t need = (n + 255) / 256;
    const int64_t cap = (int64_t)ctx().sm_count * 8;
    return (int)std::max<int64_t>(1, std::min(need, cap));
}

static void add_charge_dev(Pmc2D &p, const double *x, const double *y, const double *w, int64_t n, double wscale,
                           double *rho_dev)
{
    if (n <= 0) return;
    GP_DISPATCH_DEGREE(p.degree, {
        k_pmc2d_add_charge<D><<<grid_for(n), 256, 0, ctx().stream>>>(x, y, w, n, wscale, p.scaling, p.m, rho_dev);
    });
    GP_CUDA(cudaGetLastError());
    count_launch();
}

static void evaluate_dev(Pmc2D &p, const double *x, const double *y, int64_t n, const double *f1, const double *f2,
                         double *o1, double *o2)
{
    if (n <= 0) return;
    GP_DISPATCH_DEGREE(p.degree, {
        if (f2) k_pmc2d_evaluate<D, 2><<<grid_for(n), 256, 0, ctx().stream>>>(x, y, n, p.m, f1, f2, o1, o2);
        else k_pmc2d_evaluate<D, 1><<<grid_for(n), 256, 0, ctx().stream>>>(x, y, n, p.m, f1, nullptr, o1, nullptr);
    });
    GP_CUDA(cudaGetLastError());
    count_launch();
}

}  // namespace gempic

using namespace gempic;

extern "C" {

int gempic_pmc2d_create(double xmin, double xmax, int nx, double ymin, double ymax, int ny, int spline_degree,
                        int smoothing_type, gempic_handle *out)
{
    GP_API_BEGIN
    require_init();
    GP_REQUIRE(out, GEMPIC_EINVAL, "null output handle");
    GP_REQUIRE(nx >= 1 && ny >= 1 && xmax > xmin && ymax > ymin, GEMPIC_EINVAL, "bad 2D mesh");
    GP_REQUIRE(spline_degree >= 0 && spline_degree <= kMaxDegree, GEMPIC_EINVAL, "unsupported spline degree %d", spline_degree);
    GP_REQUIRE(nx >= spline_degree && ny >= spline_degree, GEMPIC_EASSERT, "ncells >= degree (splinepp.jl:36)");
    auto p = std::make_unique<Pmc2D>();
    p->m = {xmin, ymin, (xmax - xmin) / nx, (ymax - ymin) / ny, nx, ny};
    p->degree = spline_degree;
    if (smoothing_type == GEMPIC_COLLOCATION) p->scaling = 1.0 / (p->m.dx * p->m.dy);
    else if (smoothing_type == GEMPIC_GALERKIN) p->scaling = 1.0;
    else fail(GEMPIC_EINVAL, "Smoothing Type %d not implemented for kernel_smoother_spline_2d", smoothing_type);
    *out = register_object(std::move(p));
    GP_API_END
}

int gempic_pmc2d_destroy(gempic_handle pmc)
{
    GP_API_BEGIN
    destroy(pmc, Kind::Pmc2D, "ParticleMeshCoupling2D");
    GP_API_END
}

int gempic_pmc2d_add_charge(gempic_handle pmc, const double *x, const double *y, const double *w, int64_t n, double *rho)
{
    GP_API_BEGIN
    require_init();
    Pmc2D *p = get<Pmc2D>(pmc, "ParticleMeshCoupling2D");
    GP_REQUIRE(rho && (n == 0 || (x && y && w)) && n >= 0, GEMPIC_EINVAL, "bad arguments");
    const size_t ng = (size_t)p->m.nx * p->m.ny;
    DevBuf<double> d((size_t)3 * std::max<int64_t>(n, 1) + ng);
    double *dx = d.p, *dy = d.p + n, *dw = d.p + 2 * n, *drho = d.p + 3 * std::max<int64_t>(n, 1);
    if (n) {
        h2d(dx, x, n);
        h2d(dy, y, n);
        h2d(dw, w, n);
    }
    h2d(drho, rho, ng);   // rho += ...
    add_charge_dev(*p, dx, dy, dw, n, 1.0, drho);
    d2h(rho, drho, ng);
    GP_API_END
}

int gempic_pmc2d_evaluate(gempic_handle pmc, const double *x, const double *y, int64_t n, const double *field, double *out)
{
    GP_API_BEGIN
    require_init();
    Pmc2D *p = get<Pmc2D>(pmc, "ParticleMeshCoupling2D");
    GP_REQUIRE(n >= 0 && field && (n == 0 || (x && y && out)), GEMPIC_EINVAL, "bad arguments");
    if (n == 0) return GEMPIC_OK;
    const size_t ng = (size_t)p->m.nx * p->m.ny;
    DevBuf<double> d((size_t)3 * n + ng);
    h2d(d.p, x, n);
    h2d(d.p + n, y, n);
    h2d(d.p + 3 * n, field, ng);
    evaluate_dev(*p, d.p, d.p + n, n, d.p + 3 * n, nullptr, d.p + 2 * n, nullptr);
    d2h(out, d.p + 2 * n, n);
    GP_API_END
}

int gempic_pmc2d_evaluate_multiple(gempic_handle pmc, const double *x, const double *y, int64_t n, const double *field1,
                                   const double *field2, double *out1, double *out2)
{
    GP_API_BEGIN
    require_init();
    Pmc2D *p = get<Pmc2D>(pmc, "ParticleMeshCoupling2D");
    GP_REQUIRE(n >= 0 && field1 && field2 && (n == 0 || (x && y && out1 && out2)), GEMPIC_EINVAL, "bad arguments");
    if (n == 0) return GEMPIC_OK;
    const size_t ng = (size_t)p->m.nx * p->m.ny;
    DevBuf<double> d((size_t)4 * n + 2 * ng);
    double *f1 = d.p + 4 * n, *f2 = f1 + ng;
    h2d(d.p, x, n);
    h2d(d.p + n, y, n);
    h2d(f1, field1, ng);
    h2d(f2, field2, ng);
    evaluate_dev(*p, d.p, d.p + n, n, f1, f2, d.p + 2 * n, d.p + 3 * n);
    d2h(out1, d.p + 2 * n, n);
    d2h(out2, d.p + 3 * n, n);
    GP_API_END
}

int gempic_pmc2d_add_charge_pg(gempic_handle pmc, gempic_handle pgh, double *rho)
{
    GP_API_BEGIN
    require_init();
    Pmc2D *p = get<Pmc2D>(pmc, "ParticleMeshCoupling2D");
    ParticleGroup *pg = get_pg(pgh);
    GP_REQUIRE(pg->D == 2, GEMPIC_EASSERT, "ParticleMeshCoupling2D needs a ParticleGroup{2,V}");
    GP_REQUIRE(rho, GEMPIC_EINVAL, "null rho");
    const size_t ng = (size_t)p->m.nx * p->m.ny;
    DevBuf<double> d(ng);
    h2d(d.p, rho, ng);
    add_charge_dev(*p, pg->row(0), pg->row(1), pg->row(pg->D + pg->V), pg->n, pg->charge * pg->common_weight, d.p);
    allreduce_sum(d.p, (int64_t)ng);
    d2h(rho, d.p, ng);
    GP_API_END
}

int gempic_pmc2d_evaluate_pg(gempic_handle pmc, gempic_handle pgh, const double *field, double *out)
{
    GP_API_BEGIN
    require_init();
    Pmc2D *p = get<Pmc2D>(pmc, "ParticleMeshCoupling2D");
    ParticleGroup *pg = get_pg(pgh);
    GP_REQUIRE(pg->D == 2, GEMPIC_EASSERT, "ParticleMeshCoupling2D needs a ParticleGroup{2,V}");
    GP_REQUIRE(field && out, GEMPIC_EINVAL, "null buffer");
    const size_t ng = (size_t)p->m.nx * p->m.ny;
    DevBuf<double> d(ng + (size_t)pg->n);
    h2d(d.p, field, ng);
    evaluate_dev(*p, pg->row(0), pg->row(1), pg->n, d.p, nullptr, d.p + ng, nullptr);
    d2h(out, d.p + ng, pg->n);
    GP_API_END
}

}  // extern "C"
