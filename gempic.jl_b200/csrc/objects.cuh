// objects.cuh -- the device-resident objects behind the C-ABI handles.
#pragma once
#include "common.cuh"
#include "launch.cuh"

namespace gempic {

// ParticleGroup{D,V} (src/particle_group.jl:15-46): device SoA, one fp64 row per
// coordinate, rows ordered x1..xD, v1..vV, w1..wW like the rows of the reference array.
struct Splitting;
struct Splitting2D;
struct ParticleGroup : Object {
    static constexpr Kind kKind = Kind::ParticleGroup;
    // A fused strang_splitting! call may leave its trailing operatorHE kick pending (hs1d.cu): the velocities lag
    // by that kick until the owner's next call folds it into its first pass, or pg_sync() applies it because
    // somebody else is about to look at the particles.
    Splitting *pending = nullptr;
    Splitting2D *pending2d = nullptr;   // the same for HamiltonianSplitting{2,3} (hs2d.cu)
    int D, V, W;
    int64_t n;
    double charge, mass, common_weight, q_over_m;
    DevBuf<double> data;  // (D+V+W) rows of `stride` doubles
    size_t stride = 0;    // row pitch in doubles (multiple of 32 -> 256 B aligned rows)
    DevBuf<double> sort_tmp;
    DevBuf<int> sort_keys;    // per-cell counters / cursors of the 2D sort
    // Results of the loop-tail pass (hs1d.cu loop_tail_pass, OpLoopTail): this rank's rho deposit and write_step! sums,
    // valid while no launch has written particle rows since (Context::particle_epoch) and the caller asks with the same
    // smoothers / fields.  gempic_solve_poisson and gempic_diag_write_step look here before running their own pass.
    struct TailCache {
        uint64_t epoch = 0;
        bool rho_valid = false, diag_valid = false;
        double key[12] = {0};          // degree, n, xmin, dx, scaling of both smoothers, Lx of the solver
        DevBuf<double> buf;            // [rho (n) | KE, P1, P2, transfer, vvb]
        std::vector<double> fields;    // host copy of e1 | e2 | b the sums were taken with
    } tail;
    bool exposed = false;     // a raw row pointer was handed out (gempic_pg_row_ptr): writes can no longer be tracked
    uint64_t generation = 0;  // bumped whenever the row pointers change (sort)
    bool sorted2d = false;    // rows are in 2D cell order (hs2d.cu keeps them so); cleared by whoever rewrites positions
    ParticleGroup() : Object(kKind) {}
    int rows() const { return D + V + W; }
    double *row(int r) { return data.p + (size_t)r * stride; }
    Rows rows1d()  // x, v1, (v2), w of a {1,1} or {1,2} group
    {
        Rows r;
        r.x = row(0);
        r.v1 = row(1);
        r.v2 = V >= 2 ? row(2) : nullptr;
        r.w = row(D + V);
        return r;
    }
};

// ParticleMeshCoupling1D (src/particle_mesh_coupling_1d.jl:26-95)
struct Pmc1D : Object {
    static constexpr Kind kKind = Kind::Pmc1D;
    double xmin, xmax, Lx, delta_x, scaling;
    int n_grid, degree, smoothing;
    PartialScratch scratch;
    DevBuf<double> grid_tmp;  // n_grid doubles
    Pmc1D() : Object(kKind) {}
    Mesh1D mesh(double Lmod) const
    {
        Mesh1D m;
        m.xmin = xmin;
        m.dx = delta_x;
        m.inv_dx = 1.0 / delta_x;
        m.Lx = Lmod;
        m.n = n_grid;
        m.pow2 = (n_grid & (n_grid - 1)) == 0;
        return m;
    }
};

// Maxwell1DFEM (src/maxwell_1d_fem.jl:29-177).  The reference diagonalises its circulant
// operators with FFTW; here each operator is kept as the first column of the circulant
// matrix (computed once on the host in extended precision from the same eigenvalue tables)
// and applied on the device as a periodic convolution: no FFT, no CPU fallback.
struct Maxwell1D : Object {
    static constexpr Kind kKind = Kind::Maxwell1D;
    double xmin, Lx, delta_x;
    int n, s_deg_0, s_deg_1;
    std::vector<double> eig_mass0, eig_mass1, eig_weak_ampere, eig_weak_poisson;  // half-complex layout
    // device first columns: [mass0, mass1, inv_mass0, inv_mass1, weak_ampere, weak_poisson]
    enum Col { C_MASS0 = 0, C_MASS1, C_INV_MASS0, C_INV_MASS1, C_AMPERE, C_POISSON, C_COUNT };
    DevBuf<double> cols;  // C_COUNT * n
    DevBuf<double> tmp;   // 4 * n scratch (host-buffer entry points)
    Maxwell1D() : Object(kKind) {}
    const double *col(int c) const { return cols.p + (size_t)c * n; }
};

std::unique_ptr<Maxwell1D> make_maxwell1d(double xmin, double xmax, int n_dofs, int degree);

// ---- field kernels (fields1d.cu, compiled without FMA contraction) ---------------------
void field_e_from_rho(const Maxwell1D &m, double *e, const double *rho);
// j *= prescale (when prescale != 1), then e -= circ(inv_mass, j) / dx   (component 1: mass1, 2: mass0)
void field_e_from_j(const Maxwell1D &m, double *e, double *j, int component, double prescale);
void field_e_from_b(const Maxwell1D &m, double *e, double dt, const double *b);
// all field-only updates between two fused particle passes of strang_splitting! in one launch (fields1d.cu)
struct StrangFields {
    double *e1, *e2, *b, *j1, *acc, *eT;   // acc = [j2 | j1] deposits of the fused pass; eT = [e1T | e2T]
    const double *inv_mass0, *inv_mass1, *ampere;
    int n;
    double dx;
    int do_solve, do_tail, do_lead;
    double j2_scale, dt_tail, dt_lead;     // dt_* = the (half) time steps of the HE/HB pair
    // single GPU: the per-block partial sums of the pass, reduced into acc by this kernel first (n_partials < 0: acc is
    // already reduced -- and all-reduced -- by the caller)
    const double *partials;
    int n_partials, n_acc;
    XchgDev x;   // several ranks, n_partials >= 0: the locally reduced acc is summed over the ranks inside the kernel (xchg.cuh)
};
void field_strang_fields(const Maxwell1D &m, StrangFields F);
// the same for HamiltonianSplittingBoris: step (4) of the step just pushed and step (1) of the next one
struct BorisFields {
    double *e1, *e2, *b, *j1, *j2, *e1_mid, *e2_mid, *b_mid;
    const double *inv_mass0, *inv_mass1, *ampere;
    int n;
    double dx;
    int do_post, do_pre;
    double dt_post, dt_pre;
    // single GPU: per-block partial sums [j1 | j2] of the pass, reduced into j1, j2 by this kernel (n_partials < 0: done)
    const double *partials;
    int n_partials;
    XchgDev x;   // like StrangFields::x
};
void field_boris_fields(const Maxwell1D &m, BorisFields F);
void field_b_from_e(const Maxwell1D &m, double *b, double dt, const double *e);
// out[0] = sum_i c1[i] * circ(mass_deg, c2)[i] * dx
void field_inner_product(const Maxwell1D &m, const double *c1, const double *c2, int degree, double *out);
void field_axpby(double *y, double a, const double *x, double b, int n);   // y = a*x + b*y
void field_copy(double *dst, const double *src, int n);
void field_max_abs_diff(const double *a, const double *b, int n, double *out);

// TwoDMaxwell (src/maxwell_2d_fem.jl:11-87) with its TwoDPoisson and TwoDLinearSolverSplineMass
// members.  Index [d][a]: spline family d (0: s_deg_0, 1: s_deg_1) along axis a (0: x, 1: y).
struct Maxwell2D : Object {
    static constexpr Kind kKind = Kind::Maxwell2D;
    static constexpr int kWork = 12;
    int nx, ny, s_deg_0, s_deg_1;
    double xmin, ymin, Lx, Ly, dx, dy;
    std::vector<double> line[2][2];   // mass_line_0 / mass_line_1 scaled by dx, dy (:36-42)
    std::vector<double> eig[2][2];    // spline_fem_compute_mass_eig of those lines (:51-54)
    DevBuf<double> tab;               // device tables, offsets below
    size_t off_inv[2][2], off_line[2][2], off_cos[2], off_sin[2], off_dre[2], off_dim[2], off_dtm[2], off_m0[2];
    DevBuf<double> work;              // kWork scratch vectors of nx*ny
    Maxwell2D() : Object(kKind) {}
    const double *inv_col(int d, int a) const { return tab.p + off_inv[d][a]; }
    const double *mass_line(int d, int a) const { return tab.p + off_line[d][a]; }
    double *wk(int i) const { return work.p + (size_t)i * nx * ny; }
};
std::unique_ptr<Maxwell2D> make_maxwell2d(double xmin, double xmax, int nx, double ymin, double ymax, int ny, int degree);
// device-pointer field operators (fields2d.cu); `deg[c]` = spline family along (x, y) of component c
void m2d_form_degrees(int component, int form, int out[2]);
void m2d_solve_mass(const Maxwell2D &m, int njobs, const int (*deg)[2], const double *const *in, double *const *out,
                    const double *const *base, int mode, double scale);
void m2d_multiply_mass(const Maxwell2D &m, int njobs, const int (*deg)[2], const double *const *in, double *const *out);
void m2d_e_from_j(const Maxwell2D &m, double *e, const double *j, int component);
void m2d_e_from_b(const Maxwell2D &m, double *const e[3], double dt, const double *const b[3]);
void m2d_b_from_e(const Maxwell2D &m, double *const b[3], double dt, const double *const e[3]);
void m2d_rho_from_e(const Maxwell2D &m, double *rho, const double *const e[3]);
void m2d_e_from_rho(const Maxwell2D &m, double *e1, double *e2, const double *rho);
void m2d_inner_product(const Maxwell2D &m, const double *c1, const double *c2, int component, int form, double *out);

// HamiltonianSplitting{1,2}/{1,1} (src/hamiltonian_splitting.jl:20-86)
struct Splitting : Object {
    static constexpr Kind kKind = Kind::Splitting;
    int D, V;
    Maxwell1D *maxwell = nullptr;   // retained (common.cuh)
    Pmc1D *ks0 = nullptr, *ks1 = nullptr;
    ParticleGroup *pg = nullptr;
    int n;
    DevBuf<double> fields;  // e1, e2, b, j1, j2, acc(3n), e1T, e2T   (10 * n)
    PartialScratch scratch;
    int fuse = 0;
    double pending_dt = 0.0;       // dt of the deferred trailing HE kick (ParticleGroup::pending == this)
    // after a fused pass j2() is not yet the reference's j_dofs[2]: hs_materialise_j2() / pg_sync() rebuild it from the
    // particles (hs1d.cu).  j2_stale implies ParticleGroup::pending == this.
    bool j2_stale = false;
    bool j2_unreduced = false;     // j2() holds this rank's share only; summed by hs_materialise_j2 (collective)
    // e1 | e2 | b as last delivered to the host (gempic_hs_get_fields) and the field epoch they belong to: lets the
    // loop-tail pass know that the caller's arrays are the fields it is about to take the write_step! sums with
    std::vector<double> stash;
    uint64_t stash_epoch = 0, fields_epoch = 1;
    double j2_scale = 0.0;
    Splitting() : Object(kKind) {}
    ~Splitting() override;   // drops a pending kick registration and releases maxwell, ks0, ks1, pg (hs1d.cu)
    double *e1() { return fields.p; }
    double *e2() { return fields.p + n; }
    double *b() { return fields.p + 2 * (size_t)n; }
    double *j1() { return fields.p + 3 * (size_t)n; }
    double *j2() { return fields.p + 4 * (size_t)n; }
    double *acc() { return fields.p + 5 * (size_t)n; }
    double *e1T() { return fields.p + 8 * (size_t)n; }
    double *e2T() { return fields.p + 9 * (size_t)n; }
    Mesh1D mesh() const { return ks0->mesh(maxwell->Lx); }
};

// HamiltonianSplittingBoris (src/hamiltonian_splitting_boris.jl:23-88)
struct Boris : Object {
    static constexpr Kind kKind = Kind::Boris;
    Maxwell1D *maxwell = nullptr;   // retained
    Pmc1D *ks0 = nullptr, *ks1 = nullptr;
    ParticleGroup *pg = nullptr;
    int n;
    DevBuf<double> fields;  // e1, e2, b, j1, j2, e1_mid, e2_mid, b_mid, acc(2n)  (10 * n)
    PartialScratch scratch;
    Boris() : Object(kKind) {}
    ~Boris() override;
    double *f(int which) { return fields.p + (size_t)which * n; }  // GEMPIC_F_* order
    double *acc() { return fields.p + 8 * (size_t)n; }
    Mesh1D mesh() const { return ks0->mesh(maxwell->Lx); }
};

// HamiltonianSplitting{2,3} on TwoDMaxwell (hs2d.cu)
struct Splitting2D : Object {
    static constexpr Kind kKind = Kind::Splitting2D;
    Maxwell2D *maxwell = nullptr;   // retained
    ParticleGroup *pg = nullptr;
    size_t nd = 0;             // nx * ny
    DevBuf<double> fields;     // e1 e2 e3 b1 b2 b3 j1 j2 j3 rho e1T e2T e3T (13 * nd) + 16 scalars
    int fuse = 2;              // 2: + sorted fast path (k2_sorted), 1: fused [HE,Hp3] pass and cross-step HE fold, 0: per operator
    DevBuf<double> celltab;    // per-cell field tables of the sorted fast path (hs2d.cu CellTab)
    PartialScratch scratch;
    int sort_interval = 1;     // cell-sort every k Strang steps (0: never)
    int64_t steps_done = 0;
    double pending_dt = 0.0;   // dt of the deferred trailing HE kick (ParticleGroup::pending2d == this; fields in eT)
    Splitting2D() : Object(kKind) {}
    ~Splitting2D() override;
    double *e(int c) { return fields.p + (size_t)c * nd; }
    double *b(int c) { return fields.p + (size_t)(3 + c) * nd; }
    double *j(int c) { return fields.p + (size_t)(6 + c) * nd; }
    double *eT(int c) { return fields.p + (size_t)(10 + c) * nd; }
};
void hs2d_operator(Splitting2D &h, int op, double dt);
void hs2d_apply_pending(ParticleGroup &pg);   // the deferred trailing HE kick of a fused 2d3v strang_splitting (hs2d.cu)
void hs2d_strang(Splitting2D &h, double dt, int64_t steps);
void pg_sort_2d(ParticleGroup &pg, const Maxwell2D &m);
void sort_scan(int *counts, int64_t total);   // exclusive prefix sum in place (particles.cu)

// applies a deferred trailing HE kick, if any (hs1d.cu); every entry point that reads or moves particles calls it
void pg_sync(ParticleGroup &pg);
inline ParticleGroup *get_pg(gempic_handle h)
{
    ParticleGroup *pg = get<ParticleGroup>(h, "ParticleGroup");
    pg_sync(*pg);
    return pg;
}

// operator implementations (hs1d.cu / boris.cu)
void hs_operator(Splitting &h, int op, double dt, bool inside_strang);
void hs_strang(Splitting &h, double dt, int64_t steps);
void hs_materialise_j2(Splitting &h);
void tail_key(const Pmc1D &ks0, const Pmc1D &ks1, double Lmod, double (&key)[12]);
void boris_push_v_epart(Boris &s, double dt);
void boris_push_v_bpart(Boris &s, double dt);
void boris_push_x_accumulate_j(Boris &s, double dt);
void boris_staggering(Boris &s, double dt);
void boris_strang(Boris &s, double dt, int64_t steps);

// batched coupling entry points on device arrays (pmc1d.cu)
void pmc1d_add_charge_dev(Pmc1D &p, const double *x, const double *w, int64_t n, double charge, double cw,
                          double *rho_out /* n_grid, overwritten */);
void pmc1d_evaluate_dev(Pmc1D &p, const double *x, int64_t n, const double *field, double *out);
void pmc1d_add_current_dev(Pmc1D &p, const double *x_old, const double *x_new, const double *w, double qm,
                           const double *bfield /* may be null: 1d1v variant */, double *v, int64_t n,
                           double *j_out /* n_grid, overwritten */);

// diagnostics (diag.cu): particle sums [KE, P1, P2, transfer, vvb] into out5 (device)
void diag_particle_sums(ParticleGroup &pg, Pmc1D &ks0, Pmc1D &ks1, const Maxwell1D &m, const double *e1,
                        const double *e2, const double *b, PartialScratch &scratch, double *out5);

// particle storage helpers (particles.cu)
void pg_upload(ParticleGroup &pg, const double *aos);
void pg_download(ParticleGroup &pg, double *aos);
void pg_sort_1d(ParticleGroup &pg, const Pmc1D &p);
void pg_sample(ParticleGroup &pg, int kind, double xmin, double L, double alpha, double k, const double *sigma,
               uint64_t seed, int64_t first_index);

}  // namespace gempic
