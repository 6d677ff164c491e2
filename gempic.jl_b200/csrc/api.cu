// api.cu -- extern "C" entry points of include/gempic_b200.h (1D path: ParticleGroup,
// ParticleMeshCoupling1D, Maxwell1DFEM, HamiltonianSplitting, HamiltonianSplittingBoris,
// diagnostics).  Argument checking mirrors the reference's ArgumentError / @assert sites.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "hostutil.hpp"
#include "objects.cuh"

using namespace gempic;

extern "C" {

// =============================== ParticleGroup ============================================
int gempic_pg_create(int D, int V, int n_weights, int64_t n_particles, double charge, double mass,
                     double common_weight, gempic_handle *out)
{
    GP_API_BEGIN
    require_init();
    GP_REQUIRE(out, GEMPIC_EINVAL, "null output handle");
    GP_REQUIRE(D >= 1 && D <= 3 && V >= 1 && V <= 3 && n_weights >= 1, GEMPIC_EINVAL, "bad dims (%d,%d,%d)", D, V, n_weights);
    GP_REQUIRE(n_particles >= 0, GEMPIC_EINVAL, "negative particle count");
    GP_REQUIRE(mass != 0.0, GEMPIC_EINVAL, "mass must be non-zero");
    auto pg = std::make_unique<ParticleGroup>();
    pg->D = D; pg->V = V; pg->W = n_weights; pg->n = n_particles;
    pg->charge = charge; pg->mass = mass;
    pg->common_weight = common_weight == 0.0 ? 1.0 / (double)n_particles : common_weight;   // particle_group.jl:30-32
    pg->q_over_m = charge / mass;
    pg->stride = ((size_t)std::max<int64_t>(n_particles, 1) + 31) / 32 * 32;
    pg->data.alloc(pg->stride * (size_t)pg->rows());
    pg->data.zero(ctx().stream);   // zeros(Float64, ...) :29
    *out = register_object(std::move(pg));
    GP_API_END
}

int gempic_pg_destroy(gempic_handle pg)
{
    GP_API_BEGIN
    destroy(pg, Kind::ParticleGroup, "ParticleGroup");
    GP_API_END
}

int gempic_pg_upload(gempic_handle h, const double *aos)
{
    GP_API_BEGIN
    require_init();
    ParticleGroup *pg = get_pg(h);
    GP_REQUIRE(aos || pg->n == 0, GEMPIC_EINVAL, "null particle array");
    if (pg->n) pg_upload(*pg, aos);
    GP_API_END
}

int gempic_pg_download(gempic_handle h, double *aos)
{
    GP_API_BEGIN
    require_init();
    ParticleGroup *pg = get_pg(h);
    GP_REQUIRE(aos || pg->n == 0, GEMPIC_EINVAL, "null particle array");
    if (pg->n) pg_download(*pg, aos);
    GP_API_END
}

int gempic_pg_set_row_device(gempic_handle h, int row, const double *dev_src)
{
    GP_API_BEGIN
    require_init();
    ParticleGroup *pg = get_pg(h);
    GP_REQUIRE(row >= 0 && row < pg->rows() && dev_src, GEMPIC_EINVAL, "bad row %d", row);
    pg->sorted2d = false;
    particles_changed();
    GP_CUDA(cudaMemcpyAsync(pg->row(row), dev_src, sizeof(double) * pg->n, cudaMemcpyDeviceToDevice, ctx().stream));
    GP_CUDA(cudaStreamSynchronize(ctx().stream));
    GP_API_END
}

int gempic_pg_get_row_device(gempic_handle h, int row, double *dev_dst)
{
    GP_API_BEGIN
    require_init();
    ParticleGroup *pg = get_pg(h);
    GP_REQUIRE(row >= 0 && row < pg->rows() && dev_dst, GEMPIC_EINVAL, "bad row %d", row);
    GP_CUDA(cudaMemcpyAsync(dev_dst, pg->row(row), sizeof(double) * pg->n, cudaMemcpyDeviceToDevice, ctx().stream));
    GP_CUDA(cudaStreamSynchronize(ctx().stream));
    GP_API_END
}

int gempic_pg_row_ptr(gempic_handle h, int row, double **dev_ptr)
{
    GP_API_BEGIN
    require_init();
    ParticleGroup *pg = get_pg(h);
    GP_REQUIRE(row >= 0 && row < pg->rows() && dev_ptr, GEMPIC_EINVAL, "bad row %d", row);
    pg->exposed = true;   // the caller may write through the pointer at any time: no cached particle sums from now on
    particles_changed();
    *dev_ptr = pg->row(row);
    GP_API_END
}

int gempic_pg_info(gempic_handle h, int *D, int *V, int *n_weights, int64_t *n_particles, double *charge, double *mass,
                   double *common_weight)
{
    GP_API_BEGIN
    ParticleGroup *pg = get_pg(h);
    if (D) *D = pg->D;
    if (V) *V = pg->V;
    if (n_weights) *n_weights = pg->W;
    if (n_particles) *n_particles = pg->n;
    if (charge) *charge = pg->charge;
    if (mass) *mass = pg->mass;
    if (common_weight) *common_weight = pg->common_weight;
    GP_API_END
}

int gempic_pg_sort(gempic_handle h, gempic_handle pmc)
{
    GP_API_BEGIN
    require_init();
    ParticleGroup *pg = get_pg(h);
    GP_REQUIRE(pg->D == 1, GEMPIC_EINVAL, "cell sort is implemented for D = 1 particle groups");
    Pmc1D *p = get<Pmc1D>(pmc, "ParticleMeshCoupling1D");
    pg_sort_1d(*pg, *p);
    GP_API_END
}

int gempic_pg_sample(gempic_handle h, int kind, double xmin, double L, double alpha, double k, const double *sigma,
                     uint64_t seed, int64_t first_index)
{
    GP_API_BEGIN
    require_init();
    ParticleGroup *pg = get_pg(h);
    GP_REQUIRE(L > 0.0, GEMPIC_EINVAL, "domain length must be positive");
    pg_sample(*pg, kind, xmin, L, alpha, k, sigma, seed, first_index);
    GP_API_END
}

// =============================== ParticleMeshCoupling1D ===================================
int gempic_pmc1d_create(double xmin, double xmax, int n_grid, int64_t no_particles, int spline_degree,
                        int smoothing_type, gempic_handle *out)
{
    GP_API_BEGIN
    require_init();
    (void)no_particles;
    GP_REQUIRE(out, GEMPIC_EINVAL, "null output handle");
    GP_REQUIRE(n_grid >= 1 && xmax > xmin, GEMPIC_EINVAL, "bad mesh [%g,%g] n=%d", xmin, xmax, n_grid);
    GP_REQUIRE(spline_degree >= 0 && spline_degree <= kMaxDegree, GEMPIC_EINVAL,
               "spline degree %d not supported (0..%d)", spline_degree, kMaxDegree);
    GP_REQUIRE(n_grid >= spline_degree, GEMPIC_EASSERT, "ncells >= degree (src/splinepp.jl:36)");
    auto p = std::make_unique<Pmc1D>();
    p->xmin = xmin; p->xmax = xmax; p->Lx = xmax - xmin;
    p->n_grid = n_grid; p->degree = spline_degree; p->smoothing = smoothing_type;
    p->delta_x = (xmax - xmin) / n_grid;   // :53
    if (smoothing_type == GEMPIC_COLLOCATION) p->scaling = 1.0 / p->delta_x;
    else if (smoothing_type == GEMPIC_GALERKIN) p->scaling = 1.0;
    else fail(GEMPIC_EINVAL, "Smoothing Type %d not implemented for kernel_smoother_spline_1d.", smoothing_type);   // :61
    p->grid_tmp.alloc((size_t)2 * n_grid);
    *out = register_object(std::move(p));
    GP_API_END
}

int gempic_pmc1d_destroy(gempic_handle pmc)
{
    GP_API_BEGIN
    destroy(pmc, Kind::Pmc1D, "ParticleMeshCoupling1D");
    GP_API_END
}

static void accumulate_host(double *host, const double *dev_new, int n, double *dev_old_tmp)
{
    // host[i] += dev_new[i], evaluated on the device in fp64 (x + y is exact IEEE either way)
    if (!host_out_enabled()) return;   // in-process multi-device run: rank 0 alone touches the caller's array
    h2d(dev_old_tmp, host, n);
    field_axpby(dev_old_tmp, 1.0, dev_new, 1.0, n);
    d2h(host, dev_old_tmp, n);
}

int gempic_pmc1d_add_charge(gempic_handle pmc, const double *x, const double *w, int64_t n, double *rho)
{
    GP_API_BEGIN
    require_init();
    Pmc1D *p = get<Pmc1D>(pmc, "ParticleMeshCoupling1D");
    GP_REQUIRE(n >= 0 && rho && (n == 0 || (x && w)), GEMPIC_EINVAL, "bad arguments");
    if (n == 0) return GEMPIC_OK;
    const size_t pad = ((size_t)n + 1) & ~(size_t)1;
    Stage st(2 * pad);
    double *dx = st.put(x, n);
    st.used = pad;
    double *dw = st.put(w, n);
    pmc1d_add_charge_dev(*p, dx, dw, n, 1.0, 1.0, p->grid_tmp.p);
    accumulate_host(rho, p->grid_tmp.p, p->n_grid, p->grid_tmp.p + p->n_grid);
    GP_API_END
}

int gempic_pmc1d_evaluate(gempic_handle pmc, const double *x, int64_t n, const double *field, double *out)
{
    GP_API_BEGIN
    require_init();
    Pmc1D *p = get<Pmc1D>(pmc, "ParticleMeshCoupling1D");
    GP_REQUIRE(n >= 0 && field && (n == 0 || (x && out)), GEMPIC_EINVAL, "bad arguments");
    if (n == 0) return GEMPIC_OK;
    const size_t pad = ((size_t)n + 1) & ~(size_t)1;
    Stage st(2 * pad + p->n_grid);
    double *dx = st.put(x, n);
    st.used = pad;
    double *dout = st.take(pad);
    double *df = st.put(field, p->n_grid);
    pmc1d_evaluate_dev(*p, dx, n, df, dout);
    d2h(out, dout, n);
    GP_API_END
}

static int add_current_common(gempic_handle pmc, const double *x_old, const double *x_new, const double *w, double qm,
                              const double *bfield, double *v, int64_t n, double *j, bool with_b)
{
    GP_API_BEGIN
    require_init();
    Pmc1D *p = get<Pmc1D>(pmc, "ParticleMeshCoupling1D");
    GP_REQUIRE(n >= 0 && j && (n == 0 || (x_old && x_new && w)), GEMPIC_EINVAL, "bad arguments");
    GP_REQUIRE(!with_b || (bfield && (v || n == 0)), GEMPIC_EINVAL, "bfield / v required");
    if (n == 0) return GEMPIC_OK;
    const size_t pad = ((size_t)n + 1) & ~(size_t)1;
    Stage st(4 * pad + p->n_grid);
    double *d0 = st.put(x_old, n); st.used = pad;
    double *d1 = st.put(x_new, n); st.used = 2 * pad;
    double *d2 = st.put(w, n); st.used = 3 * pad;
    double *dv = nullptr, *db = nullptr;
    if (with_b) {
        dv = st.put(v, n);
        st.used = 4 * pad;
        db = st.put(bfield, p->n_grid);
    }
    pmc1d_add_current_dev(*p, d0, d1, d2, qm, db, dv, n, p->grid_tmp.p);
    accumulate_host(j, p->grid_tmp.p, p->n_grid, p->grid_tmp.p + p->n_grid);
    if (with_b) d2h(v, dv, n);
    GP_API_END
}

int gempic_pmc1d_add_current_update_v(gempic_handle pmc, const double *x_old, const double *x_new, const double *w,
                                      double qoverm, const double *bfield, double *v, int64_t n, double *j)
{
    return add_current_common(pmc, x_old, x_new, w, qoverm, bfield, v, n, j, true);
}

int gempic_pmc1d_add_current(gempic_handle pmc, const double *x_old, const double *x_new, const double *w, int64_t n,
                             double *j)
{
    return add_current_common(pmc, x_old, x_new, w, 0.0, nullptr, nullptr, n, j, false);
}

int gempic_pmc1d_add_charge_pg(gempic_handle pmc, gempic_handle pgh, double *rho)
{
    GP_API_BEGIN
    require_init();
    Pmc1D *p = get<Pmc1D>(pmc, "ParticleMeshCoupling1D");
    ParticleGroup *pg = get_pg(pgh);
    GP_REQUIRE(pg->D == 1, GEMPIC_EASSERT, "ParticleMeshCoupling1D needs a ParticleGroup{1,V}");
    GP_REQUIRE(rho, GEMPIC_EINVAL, "null rho");
    pmc1d_add_charge_dev(*p, pg->row(0), pg->row(pg->D + pg->V), pg->n, pg->charge, pg->common_weight, p->grid_tmp.p);
    allreduce_sum(p->grid_tmp.p, p->n_grid);
    accumulate_host(rho, p->grid_tmp.p, p->n_grid, p->grid_tmp.p + p->n_grid);
    GP_API_END
}

int gempic_pmc1d_evaluate_pg(gempic_handle pmc, gempic_handle pgh, const double *field, double *out)
{
    GP_API_BEGIN
    require_init();
    Pmc1D *p = get<Pmc1D>(pmc, "ParticleMeshCoupling1D");
    ParticleGroup *pg = get_pg(pgh);
    GP_REQUIRE(pg->D == 1, GEMPIC_EASSERT, "ParticleMeshCoupling1D needs a ParticleGroup{1,V}");
    GP_REQUIRE(field && (out || pg->n == 0), GEMPIC_EINVAL, "null buffer");
    if (pg->n == 0) return GEMPIC_OK;
    const size_t pad = ((size_t)pg->n + 1) & ~(size_t)1;
    Stage st(pad + p->n_grid);
    double *dout = st.take(pad);
    double *df = st.put(field, p->n_grid);
    pmc1d_evaluate_dev(*p, pg->row(0), pg->n, df, dout);
    d2h(out, dout, pg->n);
    GP_API_END
}

// =============================== Maxwell1DFEM ==============================================
int gempic_maxwell1d_create(double xmin, double xmax, int n_dofs, int degree, gempic_handle *out)
{
    GP_API_BEGIN
    require_init();
    GP_REQUIRE(out, GEMPIC_EINVAL, "null output handle");
    GP_REQUIRE(xmax > xmin, GEMPIC_EINVAL, "bad mesh");
    *out = register_object(make_maxwell1d(xmin, xmax, n_dofs, degree));
    GP_API_END
}

int gempic_maxwell1d_destroy(gempic_handle m)
{
    GP_API_BEGIN
    destroy(m, Kind::Maxwell1D, "Maxwell1DFEM");
    GP_API_END
}

int gempic_maxwell1d_get_table(gempic_handle mh, int which, double *out)
{
    GP_API_BEGIN
    Maxwell1D *m = get<Maxwell1D>(mh, "Maxwell1DFEM");
    GP_REQUIRE(out, GEMPIC_EINVAL, "null output");
    const std::vector<double> *t = nullptr;
    switch (which) {
    case 0: t = &m->eig_mass0; break;
    case 1: t = &m->eig_mass1; break;
    case 2: t = &m->eig_weak_ampere; break;
    case 3: t = &m->eig_weak_poisson; break;
    default: fail(GEMPIC_EINVAL, "unknown table %d", which);
    }
    std::memcpy(out, t->data(), sizeof(double) * m->n);
    GP_API_END
}

int gempic_maxwell1d_compute_e_from_rho(gempic_handle mh, double *e, const double *rho)
{
    GP_API_BEGIN
    require_init();
    Maxwell1D *m = get<Maxwell1D>(mh, "Maxwell1DFEM");
    GP_REQUIRE(e && rho, GEMPIC_EINVAL, "null buffer");
    double *de = m->tmp.p, *dr = m->tmp.p + m->n;
    h2d(dr, rho, m->n);
    field_e_from_rho(*m, de, dr);
    d2h(e, de, m->n);
    GP_API_END
}

int gempic_maxwell1d_compute_e_from_j(gempic_handle mh, double *e, const double *j, int component)
{
    GP_API_BEGIN
    require_init();
    Maxwell1D *m = get<Maxwell1D>(mh, "Maxwell1DFEM");
    GP_REQUIRE(e && j, GEMPIC_EINVAL, "null buffer");
    GP_REQUIRE(component == 1 || component == 2, GEMPIC_EINVAL, "Component %d not implemented ", component);   // :283
    double *de = m->tmp.p, *dj = m->tmp.p + m->n;
    h2d(de, e, m->n);
    h2d(dj, j, m->n);
    field_e_from_j(*m, de, dj, component, 1.0);
    d2h(e, de, m->n);
    GP_API_END
}

int gempic_maxwell1d_compute_e_from_b(gempic_handle mh, double *e, double dt, const double *b)
{
    GP_API_BEGIN
    require_init();
    Maxwell1D *m = get<Maxwell1D>(mh, "Maxwell1DFEM");
    GP_REQUIRE(e && b, GEMPIC_EINVAL, "null buffer");
    double *de = m->tmp.p, *db = m->tmp.p + m->n;
    h2d(de, e, m->n);
    h2d(db, b, m->n);
    field_e_from_b(*m, de, dt, db);
    d2h(e, de, m->n);
    GP_API_END
}

int gempic_maxwell1d_compute_b_from_e(gempic_handle mh, double *b, double dt, const double *e)
{
    GP_API_BEGIN
    require_init();
    Maxwell1D *m = get<Maxwell1D>(mh, "Maxwell1DFEM");
    GP_REQUIRE(e && b, GEMPIC_EINVAL, "null buffer");
    double *db = m->tmp.p, *de = m->tmp.p + m->n;
    h2d(db, b, m->n);
    h2d(de, e, m->n);
    field_b_from_e(*m, db, dt, de);
    d2h(b, db, m->n);
    GP_API_END
}

int gempic_maxwell1d_inner_product(gempic_handle mh, const double *c1, const double *c2, int degree, double *out)
{
    GP_API_BEGIN
    require_init();
    Maxwell1D *m = get<Maxwell1D>(mh, "Maxwell1DFEM");
    GP_REQUIRE(c1 && c2 && out, GEMPIC_EINVAL, "null buffer");
    double *d1 = m->tmp.p, *d2 = m->tmp.p + m->n, *dr = m->tmp.p + 2 * m->n;
    h2d(d1, c1, m->n);
    h2d(d2, c2, m->n);
    field_inner_product(*m, d1, d2, degree, dr);
    d2h(out, dr, 1);
    GP_API_END
}

int gempic_maxwell1d_l2norm_squared(gempic_handle mh, const double *c, int degree, double *out)
{
    return gempic_maxwell1d_inner_product(mh, c, c, degree, out);
}

int gempic_maxwell1d_compute_rhs_from_function(gempic_handle mh, double *coefs, gempic_func1d f, void *fctx, int degree)
{
    GP_API_BEGIN
    Maxwell1D *m = get<Maxwell1D>(mh, "Maxwell1DFEM");
    GP_REQUIRE(coefs && f, GEMPIC_EINVAL, "null argument");
    GP_REQUIRE(degree >= 0 && degree <= 5, GEMPIC_EINVAL, "bad degree %d", degree);
    // set-up quadrature on the host (cold path; src/maxwell_1d_fem.jl:188-220)
    const int np = degree + 1;
    double x[8], w[8], bspl[8][8];
    legendre_nodes(np, x, w);
    for (int k = 0; k < np; ++k) {
        x[k] = 0.5 * (x[k] + 1.0);
        w[k] = 0.5 * w[k];
        host_bsplines(degree, x[k], bspl[k]);
    }
    for (int i = 1; i <= m->n; ++i) {
        double coef = 0.0;
        for (int j = 1; j <= np; ++j)
            for (int k = 1; k <= np; ++k)
                coef = coef + w[k - 1] * f(m->delta_x * (x[k - 1] + i + j - 2), fctx) * bspl[k - 1][degree + 1 - j];
        coefs[i - 1] = coef * m->delta_x;
    }
    GP_API_END
}

int gempic_maxwell1d_l2projection(gempic_handle mh, double *coefs, gempic_func1d f, void *fctx, int degree)
{
    GP_API_BEGIN
    Maxwell1D *m = get<Maxwell1D>(mh, "Maxwell1DFEM");
    GP_REQUIRE(degree == m->s_deg_0 || degree == m->s_deg_0 - 1, GEMPIC_EINVAL, "degree %d not available", degree);   // :366
    int rc = gempic_maxwell1d_compute_rhs_from_function(mh, coefs, f, fctx, degree);
    if (rc) return rc;
    // Reference behaviour (:369-373): solve_circulant! leaves M^{-1} rhs in self.work, but the
    // returned coefficients are the *rhs* scaled by 1/dx. Kept as is for drop-in parity.
    for (int i = 0; i < m->n; ++i) coefs[i] = coefs[i] / m->delta_x;
    GP_API_END
}

// =============================== HamiltonianSplitting =====================================
int gempic_hs_create(int D, int V, gempic_handle maxwell, gempic_handle pmc0, gempic_handle pmc1, gempic_handle pgh,
                     gempic_handle *out)
{
    GP_API_BEGIN
    require_init();
    GP_REQUIRE(out, GEMPIC_EINVAL, "null output handle");
    Maxwell1D *mx = get<Maxwell1D>(maxwell, "Maxwell1DFEM");
    Pmc1D *ks0 = get<Pmc1D>(pmc0, "ParticleMeshCoupling1D"), *ks1 = get<Pmc1D>(pmc1, "ParticleMeshCoupling1D");
    ParticleGroup *pg = get_pg(pgh);
    GP_REQUIRE(D == 1 && (V == 1 || V == 2), GEMPIC_EINVAL, "HamiltonianSplitting{%d,%d} is not defined by the reference", D, V);
    GP_REQUIRE(pg->D == D && pg->V == V, GEMPIC_EASSERT, "dims == particle_group.dims (hamiltonian_splitting.jl:47)");
    GP_REQUIRE(ks0->n_grid == ks1->n_grid, GEMPIC_EASSERT,
               "kernel_smoother_0.n_dofs == kernel_smoother_1.n_dofs (hamiltonian_splitting.jl:49)");
    GP_REQUIRE(ks0->n_grid == mx->n, GEMPIC_EASSERT, "kernel smoothers and Maxwell solver differ in n_dofs");
    GP_REQUIRE(ks0->xmin == ks1->xmin && ks0->xmax == ks1->xmax, GEMPIC_EINVAL,
               "both kernel smoothers must live on the same mesh");
    auto h = std::make_unique<Splitting>();
    h->maxwell = mx; h->ks0 = ks0; h->ks1 = ks1; h->pg = pg;
    retain(mx); retain(ks0); retain(ks1); retain(pg);
    h->D = D; h->V = V; h->n = ks0->n_grid;
    h->fuse = 1;   // the fused passes are the default in every front end (gempic_hs_set_fusion(h, 0): one pass per operator)
    h->fields.alloc((size_t)10 * h->n);
    h->fields.zero(ctx().stream);
    *out = register_object(std::move(h));
    GP_API_END
}

int gempic_hs_destroy(gempic_handle hs)
{
    GP_API_BEGIN
    Splitting *h = get<Splitting>(hs, "HamiltonianSplitting");
    if (h->pg->pending == h) pg_sync(*h->pg);   // a deferred HE kick must not be lost (rank-local pass)
    destroy(hs, Kind::Splitting, "HamiltonianSplitting");
    GP_API_END
}

int gempic_hs_set_fields(gempic_handle hs, const double *e1, const double *e2, const double *b)
{
    GP_API_BEGIN
    require_init();
    Splitting *h = get<Splitting>(hs, "HamiltonianSplitting");
    const double *src[3] = {e1, e2, b};
    h->fields_epoch++;
    h2d_vectors(h->fields.p, src, 3, h->n);   // e1 | e2 | b are adjacent
    GP_API_END
}

int gempic_hs_get_fields(gempic_handle hs, double *e1, double *e2, double *b, double *j1, double *j2)
{
    GP_API_BEGIN
    require_init();
    Splitting *h = get<Splitting>(hs, "HamiltonianSplitting");
    if (j2) hs_materialise_j2(*h);
    double *dst[5] = {e1, e2, b, j1, j2};
    d2h_vectors(dst, h->fields.p, 5, h->n);   // e1 | e2 | b | j1 | j2 are adjacent
    if (e1 && e2 && b) {   // remember what the caller now holds (loop_tail_pass, hs1d.cu)
        h->stash.resize((size_t)3 * h->n);
        if (host_out_enabled()) {
            std::memcpy(h->stash.data(), e1, sizeof(double) * h->n);
            std::memcpy(h->stash.data() + h->n, e2, sizeof(double) * h->n);
            std::memcpy(h->stash.data() + 2 * (size_t)h->n, b, sizeof(double) * h->n);
        } else {   // rank > 0 of an in-process multi-device run: rank 0 is writing the caller's arrays right now
            GP_CUDA(cudaMemcpyAsync(h->stash.data(), h->fields.p, sizeof(double) * 3 * h->n, cudaMemcpyDeviceToHost, ctx().stream));
            GP_CUDA(cudaStreamSynchronize(ctx().stream));
        }
        h->stash_epoch = h->fields_epoch;
    }
    GP_API_END
}

int gempic_hs_operator(gempic_handle hs, int op, double dt)
{
    GP_API_BEGIN
    require_init();
    Splitting *h = get<Splitting>(hs, "HamiltonianSplitting");
    hs_operator(*h, op, dt, false);
    GP_API_END
}

int gempic_hs_strang_splitting(gempic_handle hs, double dt, int64_t number_steps)
{
    GP_API_BEGIN
    require_init();
    Splitting *h = get<Splitting>(hs, "HamiltonianSplitting");
    GP_REQUIRE(number_steps >= 0, GEMPIC_EINVAL, "negative step count");
    hs_strang(*h, dt, number_steps);
    GP_API_END
}

int gempic_hs_operator_host(gempic_handle hs, int op, double dt, double *e1, double *e2, double *b, double *j1, double *j2)
{
    int rc = gempic_hs_set_fields(hs, e1, e2, b);
    if (rc) return rc;
    rc = gempic_hs_operator(hs, op, dt);
    if (rc) return rc;
    return gempic_hs_get_fields(hs, e1, e2, b, j1, j2);
}

int gempic_hs_strang_splitting_host(gempic_handle hs, double dt, int64_t number_steps, double *e1, double *e2, double *b,
                                    double *j1, double *j2)
{
    int rc = gempic_hs_set_fields(hs, e1, e2, b);
    if (rc) return rc;
    rc = gempic_hs_strang_splitting(hs, dt, number_steps);
    if (rc) return rc;
    return gempic_hs_get_fields(hs, e1, e2, b, j1, j2);
}

int gempic_hs_set_fusion(gempic_handle hs, int fuse)
{
    GP_API_BEGIN
    Splitting *h = get<Splitting>(hs, "HamiltonianSplitting");
    GP_REQUIRE(fuse == 0 || fuse == 1, GEMPIC_EINVAL, "fuse must be 0 or 1");
    h->fuse = fuse;
    GP_API_END
}

// =============================== HamiltonianSplittingBoris ================================
int gempic_boris_create(gempic_handle maxwell, gempic_handle pmc0, gempic_handle pmc1, gempic_handle pgh, gempic_handle *out)
{
    GP_API_BEGIN
    require_init();
    GP_REQUIRE(out, GEMPIC_EINVAL, "null output handle");
    Maxwell1D *mx = get<Maxwell1D>(maxwell, "Maxwell1DFEM");
    Pmc1D *ks0 = get<Pmc1D>(pmc0, "ParticleMeshCoupling1D"), *ks1 = get<Pmc1D>(pmc1, "ParticleMeshCoupling1D");
    ParticleGroup *pg = get_pg(pgh);
    GP_REQUIRE(pg->D == 1 && pg->V == 2, GEMPIC_EASSERT, "HamiltonianSplittingBoris needs a ParticleGroup{1,2}");
    GP_REQUIRE(ks0->n_grid == ks1->n_grid && ks0->n_grid == mx->n, GEMPIC_EASSERT, "n_dofs mismatch");
    GP_REQUIRE(ks0->xmin == ks1->xmin && ks0->xmax == ks1->xmax, GEMPIC_EINVAL,
               "both kernel smoothers must live on the same mesh");
    auto s = std::make_unique<Boris>();
    s->maxwell = mx; s->ks0 = ks0; s->ks1 = ks1; s->pg = pg;
    retain(mx); retain(ks0); retain(ks1); retain(pg);
    s->n = ks0->n_grid;
    s->fields.alloc((size_t)10 * s->n);
    s->fields.zero(ctx().stream);
    *out = register_object(std::move(s));
    GP_API_END
}

int gempic_boris_destroy(gempic_handle bs)
{
    GP_API_BEGIN
    destroy(bs, Kind::Boris, "HamiltonianSplittingBoris");
    GP_API_END
}

int gempic_boris_set_fields(gempic_handle bs, const double *e1, const double *e2, const double *b)
{
    GP_API_BEGIN
    require_init();
    Boris *s = get<Boris>(bs, "HamiltonianSplittingBoris");
    pg_sync(*s->pg);
    const double *src[3] = {e1, e2, b};
    h2d_vectors(s->f(GEMPIC_F_E1), src, 3, s->n);   // E1 | E2 | B are adjacent
    GP_API_END
}

int gempic_boris_get_field(gempic_handle bs, int which, double *out)
{
    GP_API_BEGIN
    require_init();
    Boris *s = get<Boris>(bs, "HamiltonianSplittingBoris");
    pg_sync(*s->pg);
    GP_REQUIRE(which >= 0 && which <= GEMPIC_F_B_MID && out, GEMPIC_EINVAL, "bad field selector %d", which);
    d2h(out, s->f(which), s->n);
    GP_API_END
}

#define GP_BORIS_CALL(NAME, FN)                                        \
    int NAME(gempic_handle bs, double dt)                              \
    {                                                                  \
        GP_API_BEGIN                                                   \
        require_init();                                                \
        Boris *s = get<Boris>(bs, "HamiltonianSplittingBoris");        \
        pg_sync(*s->pg);                                               \
        FN(*s, dt);                                                    \
        GP_API_END                                                     \
    }
GP_BORIS_CALL(gempic_boris_staggering, boris_staggering)
GP_BORIS_CALL(gempic_boris_push_v_epart, boris_push_v_epart)
GP_BORIS_CALL(gempic_boris_push_v_bpart, boris_push_v_bpart)
GP_BORIS_CALL(gempic_boris_push_x_accumulate_j, boris_push_x_accumulate_j)

int gempic_boris_strang_splitting(gempic_handle bs, double dt, int64_t number_steps)
{
    GP_API_BEGIN
    require_init();
    Boris *s = get<Boris>(bs, "HamiltonianSplittingBoris");
    pg_sync(*s->pg);
    GP_REQUIRE(number_steps >= 0, GEMPIC_EINVAL, "negative step count");
    boris_strang(*s, dt, number_steps);
    GP_API_END
}

static int boris_fields_out(gempic_handle bs, double *e1, double *e2, double *b)
{
    GP_API_BEGIN
    Boris *s = get<Boris>(bs, "HamiltonianSplittingBoris");
    double *dst[3] = {e1, e2, b};
    d2h_vectors(dst, s->f(GEMPIC_F_E1), 3, s->n);
    GP_API_END
}

int gempic_boris_staggering_host(gempic_handle bs, double dt, double *e1, double *e2, double *b)
{
    int rc = gempic_boris_set_fields(bs, e1, e2, b);
    if (rc) return rc;
    rc = gempic_boris_staggering(bs, dt);
    if (rc) return rc;
    return boris_fields_out(bs, e1, e2, b);
}

int gempic_boris_strang_splitting_host(gempic_handle bs, double dt, int64_t number_steps, double *e1, double *e2, double *b)
{
    int rc = gempic_boris_set_fields(bs, e1, e2, b);
    if (rc) return rc;
    rc = gempic_boris_strang_splitting(bs, dt, number_steps);
    if (rc) return rc;
    return boris_fields_out(bs, e1, e2, b);
}

// =============================== diagnostics ===============================================
int gempic_solve_poisson(gempic_handle pgh, gempic_handle pmc0, gempic_handle mh, double *efield, double *rho)
{
    GP_API_BEGIN
    require_init();
    ParticleGroup *pg = get_pg(pgh);
    Pmc1D *p = get<Pmc1D>(pmc0, "ParticleMeshCoupling1D");
    Maxwell1D *m = get<Maxwell1D>(mh, "Maxwell1DFEM");
    GP_REQUIRE(pg->D == 1, GEMPIC_EASSERT, "solve_poisson! is 1D");
    GP_REQUIRE(p->n_grid == m->n, GEMPIC_EASSERT, "n_dofs mismatch");
    GP_REQUIRE(efield, GEMPIC_EINVAL, "null efield");
    double *drho = p->grid_tmp.p, *de = m->tmp.p;
    {   // the loop-tail pass of a fused strang_splitting! (pg_sync above) has already deposited this rho
        auto &t = pg->tail;
        double key[12];
        tail_key(*p, *p, 0.0, key);
        bool hit = t.rho_valid && !pg->exposed && t.epoch == ctx().particle_epoch && t.buf.n >= (size_t)p->n_grid;
        for (int k = 0; k < 5 && hit; ++k) hit = key[k] == t.key[k];
        if (hit) GP_CUDA(cudaMemcpyAsync(drho, t.buf.p, sizeof(double) * p->n_grid, cudaMemcpyDeviceToDevice, ctx().stream));
        else pmc1d_add_charge_dev(*p, pg->row(0), pg->row(pg->D + pg->V), pg->n, pg->charge, pg->common_weight, drho);
    }
    allreduce_sum(drho, p->n_grid);
    field_e_from_rho(*m, de, drho);
    if (rho && host_out_enabled()) GP_CUDA(cudaMemcpyAsync(rho, drho, sizeof(double) * m->n, cudaMemcpyDeviceToHost, ctx().stream));
    d2h(efield, de, m->n);
    GP_API_END
}

int gempic_diag_write_step(gempic_handle pgh, gempic_handle mh, gempic_handle pmc0, gempic_handle pmc1, double time,
                           int degree, const double *e1, const double *e2, const double *b, const double *e1_n,
                           const double *e2_n, const double *e_poisson, double *out11)
{
    GP_API_BEGIN
    require_init();
    ParticleGroup *pg = get_pg(pgh);
    Maxwell1D *m = get<Maxwell1D>(mh, "Maxwell1DFEM");
    Pmc1D *ks0 = get<Pmc1D>(pmc0, "ParticleMeshCoupling1D");
    Pmc1D *ks1 = get<Pmc1D>(pmc1, "ParticleMeshCoupling1D");
    GP_REQUIRE(pg->D == 1 && pg->V == 2, GEMPIC_EASSERT, "write_step! needs a ParticleGroup{1,2}");
    GP_REQUIRE(e1 && e2 && b && e1_n && e2_n && e_poisson && out11, GEMPIC_EINVAL, "null buffer");
    GP_REQUIRE(degree == m->s_deg_0, GEMPIC_EINVAL, "degree %d must equal the Maxwell solver degree %d", degree, m->s_deg_0);
    const int n = m->n;
    Stage st((size_t)7 * n + 16);
    double *d_e1 = st.put(e1, n), *d_e2 = st.put(e2, n), *d_b = st.put(b, n);
    double *d_e1n = st.put(e1_n, n), *d_e2n = st.put(e2_n, n), *d_ep = st.put(e_poisson, n);
    double *d_scr = st.take(n), *d_out = st.take(16);
    GP_CUDA(cudaMemsetAsync(d_scr, 0, sizeof(double) * (n + 16), ctx().stream));
    // particle sums -> d_out[0..4] = KE, P1, P2, transfer, vvb; taken already by the loop-tail pass of a fused
    // strang_splitting! (pg_sync) when the caller passes the fields that call delivered
    {
        auto &t = pg->tail;
        double key[12];
        tail_key(*ks0, *ks1, m->Lx, key);
        bool hit = t.diag_valid && !pg->exposed && t.epoch == ctx().particle_epoch && t.fields.size() == (size_t)3 * n;
        for (int k = 0; k < 11 && hit; ++k) hit = key[k] == t.key[k];
        hit = hit && std::memcmp(e1, t.fields.data(), sizeof(double) * n) == 0 &&
              std::memcmp(e2, t.fields.data() + n, sizeof(double) * n) == 0 &&
              std::memcmp(b, t.fields.data() + 2 * (size_t)n, sizeof(double) * n) == 0;
        if (hit) {
            GP_CUDA(cudaMemcpyAsync(d_out, t.buf.p + n, sizeof(double) * 5, cudaMemcpyDeviceToDevice, ctx().stream));
            allreduce_sum(d_out, 5);
        } else {
            diag_particle_sums(*pg, *ks0, *ks1, *m, d_e1, d_e2, d_b, ks0->scratch, d_out);
        }
    }
    // poynting = inner_product(e2, M0^{-1} R^T b, degree)   (diagnostics.jl:106-112)
    field_e_from_b(*m, d_scr, 1.0, d_b);
    field_inner_product(*m, d_e2, d_scr, degree, d_out + 5);
    field_inner_product(*m, d_e1, d_e1n, degree - 1, d_out + 6);
    field_inner_product(*m, d_e2, d_e2n, degree, d_out + 7);
    field_inner_product(*m, d_b, d_b, degree - 1, d_out + 8);
    field_max_abs_diff(d_e1, d_ep, n, d_out + 9);
    double r[10];
    d2h(r, d_out, 10);
    if (!host_out_enabled()) return GEMPIC_OK;
    out11[0] = time;
    out11[1] = r[0]; out11[2] = r[1]; out11[3] = r[2];
    out11[4] = r[6]; out11[5] = r[7]; out11[6] = r[8];
    out11[7] = r[3]; out11[8] = r[4]; out11[9] = r[5]; out11[10] = r[9];
    GP_API_END
}

}  // extern "C"
