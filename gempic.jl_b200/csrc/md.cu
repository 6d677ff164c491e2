// md.cu -- the PUBLIC entry points of include/gempic_b200.h.
//
// The .cu files of this directory define the per-rank implementations `gempic_impl_X` (md_rename.inc renames their
// definitions).  With one rank per process (gempic_init) every public `gempic_X` simply forwards.  After
// gempic_init_devices(n, ids) one host process drives n devices: each public call is handed to the n worker threads
// (runtime.cu, namespace md), which run the same implementation on their own device -- the in-process form of the
// "one process per GPU" sharding of DESIGN.md section 5.  Most wrappers are generated (md_wrappers_gen.inc, classes
// ALL / RANK0 / CALLER, tools/gen_md_wrappers.py); the ones below need rank-specific arguments:
//   * a particle group is created with the GLOBAL particle count and sharded by index range (md::shard); upload,
//     download, the samplers and the per-particle evaluate address the global arrays;
//   * replicated host outputs are written by rank 0 only (host_out_enabled(), runtime.cu);
//   * raw device pointers make no sense across devices and are refused.
#define GEMPIC_NO_RENAME
#include <algorithm>
#include <mutex>
#include <unordered_map>

#include "common.cuh"
#include "md_impl_decls.inc"

namespace gempic {
namespace md {
void shutdown();

// global shape of every sharded particle group
struct PgShape {
    int64_t n;
    int rows;
};
static std::mutex g_mu;
static std::unordered_map<gempic_handle, PgShape> g_pg;
static PgShape pg_shape(gempic_handle h)
{
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_pg.find(h);
    if (it == g_pg.end()) return PgShape{-1, 0};
    return it->second;
}
static int refuse(const char *what)
{
    set_error("%s hands out / takes a device pointer of ONE device and is not available after gempic_init_devices", what);
    return GEMPIC_EINVAL;
}
// every rank writes its own part of a per-particle host array
struct HostOutScope {
    HostOutScope() { set_host_out(true); }
    ~HostOutScope() { set_host_out(ctx().rank == 0); }
};

}  // namespace md
}  // namespace gempic

using namespace gempic;

extern "C" {

#include "md_wrappers_gen.inc"

// ---- runtime -------------------------------------------------------------------------------------------------
const char *gempic_last_error(void) { return gempic_impl_last_error(); }
int gempic_version(void) { return gempic_impl_version(); }

int gempic_init(int device)
{
    if (md::dispatching()) {
        set_error("already initialised with gempic_init_devices");
        return GEMPIC_EINVAL;
    }
    return gempic_impl_init(device);
}

int gempic_finalize(void)
{
    if (!md::dispatching()) return gempic_impl_finalize();
    const int rc = md::run_all([&](int) { return gempic_impl_finalize(); });
    md::shutdown();
    std::lock_guard<std::mutex> lk(md::g_mu);
    md::g_pg.clear();
    return rc;
}

void *gempic_stream(void)
{
    if (!md::dispatching()) return gempic_impl_stream();
    void *s = nullptr;
    md::run_rank0([&](int) { s = gempic_impl_stream(); return 0; });
    return s;
}

int64_t gempic_launch_count(int reset)
{
    if (!md::dispatching()) return gempic_impl_launch_count(reset);
    std::vector<int64_t> v(md::n_ranks(), 0);
    md::run_all([&](int r) { v[r] = gempic_impl_launch_count(reset); return 0; });
    return v[0];   // every rank launches the same sequence
}

int gempic_comm_unique_id(void *id128)
{
    if (md::dispatching()) return md::run_rank0([&](int) { return gempic_impl_comm_unique_id(id128); });
    return gempic_impl_comm_unique_id(id128);
}
int gempic_comm_init(int n_ranks, int rank, const void *id128)
{
    if (md::dispatching()) {
        set_error("gempic_init_devices has already created the communicator over its devices");
        return GEMPIC_EINVAL;
    }
    return gempic_impl_comm_init(n_ranks, rank, id128);
}
int gempic_comm_finalize(void)
{
    if (md::dispatching()) return GEMPIC_OK;   // the communicator lives until gempic_finalize
    return gempic_impl_comm_finalize();
}
int gempic_comm_size(void)
{
    if (!md::dispatching()) return gempic_impl_comm_size();
    return md::n_ranks();
}
int gempic_sobol_points(int dims, int64_t first, int64_t n, double *out) { return gempic_impl_sobol_points(dims, first, n, out); }

// ---- ParticleGroup: sharded by index range ----------------------------------------------------------------------
int gempic_pg_create(int D, int V, int n_weights, int64_t n_particles, double charge, double mass, double common_weight,
                     gempic_handle *out)
{
    if (!md::dispatching()) return gempic_impl_pg_create(D, V, n_weights, n_particles, charge, mass, common_weight, out);
    if (!out || n_particles < 0) {
        set_error("bad arguments");
        return GEMPIC_EINVAL;
    }
    const double cw = common_weight == 0.0 ? 1.0 / (double)n_particles : common_weight;   // of the GLOBAL count (:30-32)
    std::vector<gempic_handle> hs(md::n_ranks(), 0);
    const int rc = md::run_all([&](int r) {
        int64_t first, count;
        md::shard(n_particles, r, first, count);
        return gempic_impl_pg_create(D, V, n_weights, count, charge, mass, cw, &hs[r]);
    });
    if (rc) return rc;
    for (gempic_handle h : hs)
        if (h != hs[0]) {
            set_error("internal error: the ranks disagree on a handle");
            return GEMPIC_ECUDA;
        }
    {
        std::lock_guard<std::mutex> lk(md::g_mu);
        md::g_pg[hs[0]] = md::PgShape{n_particles, D + V + n_weights};
    }
    *out = hs[0];
    return GEMPIC_OK;
}

int gempic_pg_destroy(gempic_handle pg)
{
    if (!md::dispatching()) return gempic_impl_pg_destroy(pg);
    {
        std::lock_guard<std::mutex> lk(md::g_mu);
        md::g_pg.erase(pg);
    }
    return md::run_all([&](int) { return gempic_impl_pg_destroy(pg); });
}

int gempic_pg_upload(gempic_handle pg, const double *aos)
{
    if (!md::dispatching()) return gempic_impl_pg_upload(pg, aos);
    const md::PgShape s = md::pg_shape(pg);
    return md::run_all([&](int r) {
        int64_t first = 0, count = 0;
        if (s.n >= 0) md::shard(s.n, r, first, count);
        return gempic_impl_pg_upload(pg, aos ? aos + (size_t)first * s.rows : nullptr);
    });
}

int gempic_pg_download(gempic_handle pg, double *aos)
{
    if (!md::dispatching()) return gempic_impl_pg_download(pg, aos);
    const md::PgShape s = md::pg_shape(pg);
    return md::run_all([&](int r) {
        md::HostOutScope scope;
        int64_t first = 0, count = 0;
        if (s.n >= 0) md::shard(s.n, r, first, count);
        return gempic_impl_pg_download(pg, aos ? aos + (size_t)first * s.rows : nullptr);
    });
}

int gempic_pg_info(gempic_handle pg, int *D, int *V, int *n_weights, int64_t *n_particles, double *charge, double *mass,
                   double *common_weight)
{
    if (!md::dispatching()) return gempic_impl_pg_info(pg, D, V, n_weights, n_particles, charge, mass, common_weight);
    const int rc = md::run_rank0([&](int) { return gempic_impl_pg_info(pg, D, V, n_weights, n_particles, charge, mass, common_weight); });
    if (rc == GEMPIC_OK && n_particles) *n_particles = md::pg_shape(pg).n;
    return rc;
}

int gempic_pg_sample(gempic_handle pg, int kind, double xmin, double L, double alpha, double k, const double *sigma,
                     uint64_t seed, int64_t first_index)
{
    if (!md::dispatching()) return gempic_impl_pg_sample(pg, kind, xmin, L, alpha, k, sigma, seed, first_index);
    const md::PgShape s = md::pg_shape(pg);
    return md::run_all([&](int r) {
        int64_t first = 0, count = 0;
        if (s.n >= 0) md::shard(s.n, r, first, count);
        return gempic_impl_pg_sample(pg, kind, xmin, L, alpha, k, sigma, seed, first_index + first);
    });
}

int gempic_pg_sample_landau(gempic_handle pg, double alpha, double k, double sigma, double weight, int64_t first_index,
                            int64_t n_global)
{
    if (!md::dispatching()) return gempic_impl_pg_sample_landau(pg, alpha, k, sigma, weight, first_index, n_global);
    const md::PgShape s = md::pg_shape(pg);
    return md::run_all([&](int r) {
        int64_t first = 0, count = 0;
        if (s.n >= 0) md::shard(s.n, r, first, count);
        return gempic_impl_pg_sample_landau(pg, alpha, k, sigma, weight, first_index + first, n_global > 0 ? n_global : s.n);
    });
}

int gempic_pg_sample_cos_gaussian(gempic_handle pg, int sampling_type, int symmetric, uint64_t seed, double xmin, double dimx,
                                  int n_cos, const double *k, const double *alpha, int n_gaussians, const double *sigma,
                                  const double *mu, const double *delta, int64_t first_index)
{
    if (!md::dispatching())
        return gempic_impl_pg_sample_cos_gaussian(pg, sampling_type, symmetric, seed, xmin, dimx, n_cos, k, alpha, n_gaussians, sigma,
                                                  mu, delta, first_index);
    const md::PgShape s = md::pg_shape(pg);
    return md::run_all([&](int r) {
        int64_t first = 0, count = 0;
        if (s.n >= 0) md::shard(s.n, r, first, count);
        return gempic_impl_pg_sample_cos_gaussian(pg, sampling_type, symmetric, seed, xmin, dimx, n_cos, k, alpha, n_gaussians, sigma,
                                                  mu, delta, first_index + first);
    });
}

int gempic_pg_row_ptr(gempic_handle pg, int row, double **dev_ptr)
{
    if (md::dispatching()) return md::refuse("gempic_pg_row_ptr");
    return gempic_impl_pg_row_ptr(pg, row, dev_ptr);
}
int gempic_pg_set_row_device(gempic_handle pg, int row, const double *dev_src)
{
    if (md::dispatching()) return md::refuse("gempic_pg_set_row_device");
    return gempic_impl_pg_set_row_device(pg, row, dev_src);
}
int gempic_pg_get_row_device(gempic_handle pg, int row, double *dev_dst)
{
    if (md::dispatching()) return md::refuse("gempic_pg_get_row_device");
    return gempic_impl_pg_get_row_device(pg, row, dev_dst);
}

// ---- per-particle host outputs --------------------------------------------------------------------------------
int gempic_pmc1d_evaluate_pg(gempic_handle pmc, gempic_handle pg, const double *field, double *out)
{
    if (!md::dispatching()) return gempic_impl_pmc1d_evaluate_pg(pmc, pg, field, out);
    const md::PgShape s = md::pg_shape(pg);
    return md::run_all([&](int r) {
        md::HostOutScope scope;
        int64_t first = 0, count = 0;
        if (s.n >= 0) md::shard(s.n, r, first, count);
        return gempic_impl_pmc1d_evaluate_pg(pmc, pg, field, out ? out + first : nullptr);
    });
}
int gempic_pmc2d_evaluate_pg(gempic_handle pmc, gempic_handle pg, const double *field, double *out)
{
    if (!md::dispatching()) return gempic_impl_pmc2d_evaluate_pg(pmc, pg, field, out);
    const md::PgShape s = md::pg_shape(pg);
    return md::run_all([&](int r) {
        md::HostOutScope scope;
        int64_t first = 0, count = 0;
        if (s.n >= 0) md::shard(s.n, r, first, count);
        return gempic_impl_pmc2d_evaluate_pg(pmc, pg, field, out ? out + first : nullptr);
    });
}

}  // extern "C"
