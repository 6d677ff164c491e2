/*
 * gempic_b200.h -- C ABI of libgempic_b200.so, the B200 (sm_100a) implementation of
 * GEMPIC.jl's particle-mesh hot path.
 *
 * GEMPIC.jl has no FFI layer of its own (it is pure Julia), so the boundary sits where a
 * Julia `ccall` shim replaces the reference's per-particle loops: one call per reference
 * *struct-level* operation.  Each entry point names the reference interface it replaces
 * (paths relative to the GEMPIC.jl tree).  INTEGRATION.md shows the Julia binding.
 *
 * Conventions
 *   - all floating point data is fp64; sizes are int64_t; handles are opaque uint64_t.
 *   - every function returns a gempic_status; on failure gempic_last_error() holds a
 *     thread-local message.  GEMPIC_EINVAL mirrors the reference's ArgumentError throws,
 *     GEMPIC_EASSERT its @assert failures.
 *   - host pointers are borrowed for the duration of the call only.
 *   - particle arrays cross the boundary in the reference layout: column-major
 *     (D+V+n_weights) x N  (src/particle_group.jl:29), i.e. one record per particle.
 *   - one device per process (gempic_init) or n devices driven by one process (gempic_init_devices); calls are
 *     synchronous and not re-entrant.
 */
#ifndef GEMPIC_B200_H
#define GEMPIC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef uint64_t gempic_handle;

typedef enum {
    GEMPIC_OK = 0,
    GEMPIC_EINVAL = 1,   /* ArgumentError in the reference */
    GEMPIC_EASSERT = 2,  /* @assert failure in the reference */
    GEMPIC_EHANDLE = 3,  /* unknown / wrong-type handle */
    GEMPIC_ECUDA = 4,    /* CUDA runtime error (message has the CUDA string) */
    GEMPIC_ENCCL = 5,    /* NCCL error */
    GEMPIC_ENOTINIT = 6  /* gempic_init not called / no CUDA device */
} gempic_status;

/* smoothing_type symbols of ParticleMeshCoupling1D/2D (src/particle_mesh_coupling_1d.jl:56-65) */
enum { GEMPIC_COLLOCATION = 0, GEMPIC_GALERKIN = 1 };
/* operators of src/hamiltonian_splitting_1d2v.jl / _1d1v.jl */
enum { GEMPIC_OP_HP1 = 1, GEMPIC_OP_HP2 = 2, GEMPIC_OP_HE = 3, GEMPIC_OP_HB = 4, GEMPIC_OP_HP3 = 5 /* 2d3v only */ };
/* field selectors for gempic_hs_get_field / gempic_boris_get_field */
enum {
    GEMPIC_F_E1 = 0, GEMPIC_F_E2 = 1, GEMPIC_F_B = 2, GEMPIC_F_J1 = 3, GEMPIC_F_J2 = 4,
    GEMPIC_F_E1_MID = 5, GEMPIC_F_E2_MID = 6, GEMPIC_F_B_MID = 7
};

/* ---- runtime ------------------------------------------------------------------------ */
const char *gempic_last_error(void);
int gempic_version(void);
/* Bind the process to CUDA device `device` and create the library stream. */
int gempic_init(int device);
/* ONE host process driving n devices (the reference chunks its particles over threads inside one process,
 * src/hamiltonian_splitting.jl:61-66): after this call every entry point below acts on all n devices -- objects are
 * replicated, particle groups are sharded by index range (gempic_pg_create takes the GLOBAL particle count; upload /
 * download / sample address the global array), deposits are all-reduced over NCCL (ncclCommInitAll), replicated host
 * outputs are written once.  Each device is driven by its own worker thread inside the library; calls stay synchronous
 * for the caller.  device_ids == NULL: devices 0 .. n-1.  Not combinable with gempic_init / gempic_comm_init; raw
 * device-pointer entry points (gempic_pg_row_ptr, gempic_pg_set/get_row_device) are refused in this mode. */
int gempic_init_devices(int n_devices, const int *device_ids);
/* number of ranks this process drives: 1 after gempic_init, n after gempic_init_devices(n, ..) */
int gempic_device_count(void);
int gempic_finalize(void);
int gempic_synchronize(void);
/* cudaStream_t the library launches on (for CUDA-event timing by the caller). */
void *gempic_stream(void);
int gempic_device_info(int *sm_count, int64_t *free_bytes, int64_t *total_bytes);
/* Number of kernels this library has launched since the last reset (for bench.py). */
int64_t gempic_launch_count(int reset);
/* Per-kernel device timing: CUDA events are recorded around every particle pass while enabled
 * (no host synchronisation). gempic_profile_read walks the slots (tag = reference operator
 * name, accumulated milliseconds, launches); it returns 100 past the last slot. */
int gempic_profile_enable(int on);
int gempic_profile_read(int slot, char *tag, int tag_len, double *ms, int64_t *launches);
/* Run-time options by name.  None is defined in this version (every name returns GEMPIC_EINVAL); the per-step
 * launch sequence is already 2-3 kernels, so there is nothing for a CUDA-graph switch to toggle. */
int gempic_set_option(const char *name, int64_t value);

/* ---- multi-GPU: particles sharded by rank, grid moments all-reduced ------------------
 * replaces reduce(+, fetch.(tasks)) (src/hamiltonian_splitting_1d2v.jl:88,108,169).
 * One process per GPU; rank 0 creates the id, the host side broadcasts it
 * (torch.distributed / MPI / any), every rank calls gempic_comm_init. */
int gempic_comm_unique_id(void *id128);
int gempic_comm_init(int n_ranks, int rank, const void *id128);
int gempic_comm_finalize(void);
int gempic_comm_size(void);
/* suspend = 1: keep the communicator but let this rank work on its own (all-reduces skipped) until suspend = 0 --
 * for an un-sharded check run next to a sharded one.  Every rank must resume before the next collective call. */
int gempic_comm_suspend(int suspend);
/* test hook: all-reduce (sum) a host vector through the same path the operators use */
int gempic_comm_allreduce(double *host_inout, int64_t n);

/* ---- ParticleGroup{D,V} (src/particle_group.jl:15-46) -------------------------------- */
/* common_weight == 0.0 selects the reference default 1/n_particles (:30-32).
 * n_particles is the number of particles held by THIS rank. For a sharded run pass the
 * global common weight explicitly (1/N_global). */
int gempic_pg_create(int D, int V, int n_weights, int64_t n_particles, double charge, double mass,
                     double common_weight, gempic_handle *out);
int gempic_pg_destroy(gempic_handle pg);
/* host AoS (D+V+W) x N  <->  device SoA */
int gempic_pg_upload(gempic_handle pg, const double *aos);
int gempic_pg_download(gempic_handle pg, double *aos);
/* one SoA row (0..D+V+W-1) from/to a DEVICE pointer (e.g. a torch CUDA tensor) */
int gempic_pg_set_row_device(gempic_handle pg, int row, const double *dev_src);
int gempic_pg_get_row_device(gempic_handle pg, int row, double *dev_dst);
int gempic_pg_row_ptr(gempic_handle pg, int row, double **dev_ptr);
int gempic_pg_info(gempic_handle pg, int *D, int *V, int *n_weights, int64_t *n_particles,
                   double *charge, double *mass, double *common_weight);
/* Periodic counting sort of the SoA rows by cell of x1 on the mesh of `pmc` (D = 1; gempic_pg_sort2d for D = 2).
 * Stable: particles of one cell keep their relative order. */
int gempic_pg_sort(gempic_handle pg, gempic_handle pmc);
/* Device-side synthetic loads: counter-based RNG, seed 1234 default.
 * kind 0: x ~ U[0,L), v ~ N(0, sigma_k); kind 1: Landau x by inverse CDF of 1+alpha*cos(kx).
 * weights = L (1D) ; first_index = global index of this rank's first particle. */
int gempic_pg_sample(gempic_handle pg, int kind, double xmin, double L, double alpha, double k,
                     const double *sigma, uint64_t seed, int64_t first_index);

/* The reference's samplers on the device (csrc/sampling.cu).  The Sobol coordinates are those of Sobol.jl's
 * SobolSeq (Gray-code order, Joe-Kuo direction numbers, point 0 skipped), evaluated from the GLOBAL particle index:
 * any sharding of the index range gives the same load.
 * sample!(pg::ParticleGroup{1,1} / {1,2}, alpha, k, sigma, mesh) (src/particle_sampling.jl:266-311): weight = mesh.dimx;
 * sample!(d::LandauDamping, pg) (src/landau_damping.jl:34-59): sigma = 1, weight = 2 pi / kx / nbpart.
 * Deterministic: v_i = sigma sqrt(-2 log((i - 1/2)/N)), (r1, r2) = Sobol(2), theta = 2 pi r1, x = newton(r2).
 * n_global <= 0: this group holds all particles. */
int gempic_pg_sample_landau(gempic_handle pg, double alpha, double k, double sigma, double weight, int64_t first_index,
                            int64_t n_global);
/* sample!(pg::ParticleGroup{1,2}, ps::ParticleSampler, df::AbstractCosGaussian, mesh) (src/particle_sampling.jl:68-225).
 * sampling_type 0 :random, 1 :sobol; symmetric 0 sample_all (:89-143), 1 sample_sym (:150-225, 8-fold antithetic).
 * df: n_cos wave numbers k / strengths alpha, n_gaussians x (sigma[2], mu[2]) row-major, portions delta (NULL for one
 * Gaussian) -- CosSumGaussian and SumCosGaussian share eval_x_density (src/distributions.jl:159-178).
 * Normal deviates (and the :random uniforms) come from a counter-based generator keyed by (seed, global index): Julia's
 * MersenneTwister stream is not reproducible outside Julia -- statistical parity, exact antithetic structure.
 * first_index must be a multiple of 8 for symmetric sampling. */
int gempic_pg_sample_cos_gaussian(gempic_handle pg, int sampling_type, int symmetric, uint64_t seed, double xmin,
                                  double dimx, int n_cos, const double *k, const double *alpha, int n_gaussians,
                                  const double *sigma, const double *mu, const double *delta, int64_t first_index);
/* test hook: points first+1 .. first+n of SobolSeq(dims) as Sobol.next! returns them (row-major n x dims, dims <= 4) */
int gempic_sobol_points(int dims, int64_t first, int64_t n, double *out);

/* ---- ParticleMeshCoupling1D (src/particle_mesh_coupling_1d.jl:26-95) ------------------ */
int gempic_pmc1d_create(double xmin, double xmax, int n_grid, int64_t no_particles, int spline_degree,
                        int smoothing_type, gempic_handle *out);
int gempic_pmc1d_destroy(gempic_handle pmc);
/* batched add_charge! (:261-280): rho[n_grid] += sum_i deposit(x_i, w_i); host arrays */
int gempic_pmc1d_add_charge(gempic_handle pmc, const double *x, const double *w, int64_t n, double *rho);
/* batched evaluate (:438-453): out[i] = field(x_i) */
int gempic_pmc1d_evaluate(gempic_handle pmc, const double *x, int64_t n, const double *field, double *out);
/* batched add_current_update_v! with B (:296-376): j += ..., v[i] updated in place */
int gempic_pmc1d_add_current_update_v(gempic_handle pmc, const double *x_old, const double *x_new,
                                      const double *w, double qoverm, const double *bfield, double *v,
                                      int64_t n, double *j);
/* 1d1v variant without B (:471-529) */
int gempic_pmc1d_add_current(gempic_handle pmc, const double *x_old, const double *x_new, const double *w,
                             int64_t n, double *j);
/* same on a device-resident ParticleGroup: marker charge = charge*w*common_weight (get_charge) */
int gempic_pmc1d_add_charge_pg(gempic_handle pmc, gempic_handle pg, double *rho);
int gempic_pmc1d_evaluate_pg(gempic_handle pmc, gempic_handle pg, const double *field, double *out);

/* ---- ParticleMeshCoupling2D (src/particle_mesh_coupling_2d.jl:12-231) ----------------- */
int gempic_pmc2d_create(double xmin, double xmax, int nx, double ymin, double ymax, int ny,
                        int spline_degree, int smoothing_type, gempic_handle *out);
int gempic_pmc2d_destroy(gempic_handle pmc);
int gempic_pmc2d_add_charge(gempic_handle pmc, const double *x, const double *y, const double *w, int64_t n,
                            double *rho);
int gempic_pmc2d_evaluate(gempic_handle pmc, const double *x, const double *y, int64_t n, const double *field,
                          double *out);
int gempic_pmc2d_evaluate_multiple(gempic_handle pmc, const double *x, const double *y, int64_t n,
                                   const double *field1, const double *field2, double *out1, double *out2);
int gempic_pmc2d_add_charge_pg(gempic_handle pmc, gempic_handle pg, double *rho);
int gempic_pmc2d_evaluate_pg(gempic_handle pmc, gempic_handle pg, const double *field, double *out);

/* ---- Maxwell1DFEM (src/maxwell_1d_fem.jl:29-177) -------------------------------------- */
int gempic_maxwell1d_create(double xmin, double xmax, int n_dofs, int degree, gempic_handle *out);
int gempic_maxwell1d_destroy(gempic_handle m);
/* which: 0 eig_mass0, 1 eig_mass1, 2 eig_weak_ampere, 3 eig_weak_poisson (FFTW half-complex layout) */
int gempic_maxwell1d_get_table(gempic_handle m, int which, double *out);
int gempic_maxwell1d_compute_e_from_rho(gempic_handle m, double *e, const double *rho);          /* :244-255 */
int gempic_maxwell1d_compute_e_from_j(gempic_handle m, double *e, const double *j, int component); /* :263-289 */
int gempic_maxwell1d_compute_e_from_b(gempic_handle m, double *e, double dt, const double *b);    /* :384-396 */
int gempic_maxwell1d_compute_b_from_e(gempic_handle m, double *b, double dt, const double *e);    /* :407-420 */
int gempic_maxwell1d_inner_product(gempic_handle m, const double *c1, const double *c2, int degree,
                                   double *out);                                                  /* :461-475 */
int gempic_maxwell1d_l2norm_squared(gempic_handle m, const double *c, int degree, double *out);   /* :299-313 */
/* set-up helpers (host quadrature of a C callback, then the device circulant solve) */
typedef double (*gempic_func1d)(double x, void *ctx);
int gempic_maxwell1d_compute_rhs_from_function(gempic_handle m, double *coefs, gempic_func1d f, void *ctx,
                                               int degree);                                       /* :188-220 */
int gempic_maxwell1d_l2projection(gempic_handle m, double *coefs, gempic_func1d f, void *ctx, int degree); /* :347-374 */

/* ---- TwoDMaxwell (src/maxwell_2d_fem.jl:11-87) with TwoDPoisson (src/poisson_2d_fem.jl) and
 *      TwoDLinearSolverSplineMass (src/linear_solver_spline_mass_2d.jl) ---------------------
 * dofs are flat nx*ny vectors, x fastest.  component is 1-based, form in 0..3 (:136-149). */
int gempic_maxwell2d_create(double xmin, double xmax, int nx, double ymin, double ymax, int ny, int degree,
                            gempic_handle *out);
int gempic_maxwell2d_destroy(gempic_handle m);
/* which: 0 mass_line_0, 1 mass_line_1, 2 eig_values_mass_0, 3 eig_values_mass_1; axis 0 (x) / 1 (y) */
int gempic_maxwell2d_get_table(gempic_handle m, int which, int axis, double *out, int *count);
int gempic_maxwell2d_compute_e_from_rho(gempic_handle m, double *e1, double *e2, const double *rho);   /* :199-201, poisson :237-262 */
int gempic_maxwell2d_compute_e_from_b(gempic_handle m, double *e1, double *e2, double *e3, double dt,
                                      const double *b1, const double *b2, const double *b3);         /* :370-412 */
int gempic_maxwell2d_compute_b_from_e(gempic_handle m, double *b1, double *b2, double *b3, double dt,
                                      const double *e1, const double *e2, const double *e3);         /* :423-444 */
int gempic_maxwell2d_compute_e_from_j(gempic_handle m, double *e, const double *current, int component); /* :455-459 */
int gempic_maxwell2d_compute_rho_from_e(gempic_handle m, double *rho, const double *e1, const double *e2,
                                        const double *e3);                                            /* :468-500 */
int gempic_maxwell2d_inner_product(gempic_handle m, const double *c1, const double *c2, int component, int form,
                                   double *out);                                                      /* :512-575 */
/* solve(inv_mass_1[component] | inv_mass_2[component], rhs)  (linear_solver_spline_mass_2d.jl:16-32) */
int gempic_maxwell2d_solve_mass(gempic_handle m, double *out, const double *rhs, int component, int form);
/* multiply_mass_2dkron! with the mass lines of (component, form)  (:343-363) */
int gempic_maxwell2d_multiply_mass(gempic_handle m, double *out, const double *in, int component, int form);
typedef double (*gempic_func2d)(double x, double y, void *ctx);
int gempic_maxwell2d_compute_rhs_from_function(gempic_handle m, double *coefs, gempic_func2d f, void *ctx,
                                               int component, int form);                              /* :124-196 */
int gempic_maxwell2d_l2projection(gempic_handle m, double *coefs, gempic_func2d f, void *ctx, int component,
                                  int form);                                                          /* :283-293 */

/* ---- HamiltonianSplitting{1,2} / {1,1} (src/hamiltonian_splitting.jl:20-108) ---------- */
int gempic_hs_create(int D, int V, gempic_handle maxwell, gempic_handle pmc0, gempic_handle pmc1,
                     gempic_handle pg, gempic_handle *out);
int gempic_hs_destroy(gempic_handle hs);
/* e_dofs/b_dofs are aliased caller arrays in the reference (:80-81): copy them in / out. */
int gempic_hs_set_fields(gempic_handle hs, const double *e1, const double *e2, const double *b);
/* NULL pointers are skipped.  j1/j2 = j_dofs (:51), scratch of the splitting object: after a fused strang_splitting
 * a non-NULL j2 costs one extra deposit pass that rebuilds the reference's j_dofs[2] from the particles. */
int gempic_hs_get_fields(gempic_handle hs, double *e1, double *e2, double *b, double *j1, double *j2);
/* device-resident calls: asynchronous on gempic_stream() */
int gempic_hs_operator(gempic_handle hs, int op, double dt);          /* operatorHp1/Hp2/HE/HB */
int gempic_hs_strang_splitting(gempic_handle hs, double dt, int64_t number_steps);   /* :98-108 */
/* drop-in calls with HOST field buffers: H2D(e1,e2,b) -> op -> D2H(e1,e2,b,j1,j2) -> sync.
 * NULL j1/j2 are skipped. */
int gempic_hs_operator_host(gempic_handle hs, int op, double dt, double *e1, double *e2, double *b,
                            double *j1, double *j2);
int gempic_hs_strang_splitting_host(gempic_handle hs, double dt, int64_t number_steps, double *e1, double *e2,
                                    double *b, double *j1, double *j2);
/* Same trajectory with the particle passes of a Strang step fused:
 * fuse = 1 (default, in every front end) fused [HE,Hp2,Hp1,Hp2] pass, 0 one kernel per reference operator. */
int gempic_hs_set_fusion(gempic_handle hs, int fuse);

/* ---- HamiltonianSplitting{2,3} on TwoDMaxwell (BASELINE config 5) ---------------------------
 * Fills the empty src/hamiltonian_splitting_2d3v.jl: the 2D extension of the operators of
 * src/hamiltonian_splitting_1d2v.jl:41-236 and of strang_splitting! (src/hamiltonian_splitting.jl:98-108)
 * on ParticleGroup{2,3} (rows x1,x2,v1,v2,v3,w); operator order HB HE Hp3 Hp2 Hp1 Hp2 Hp3 HE HB.
 * Field vectors are flat nx*ny dofs, x fastest: e = (E1,E2,E3) 1-form, b = (B1,B2,B3) 2-form. */
int gempic_hs2d_create(gempic_handle maxwell2d, gempic_handle pg, gempic_handle *out);
int gempic_hs2d_destroy(gempic_handle hs);
int gempic_hs2d_set_fields(gempic_handle hs, const double *e1, const double *e2, const double *e3, const double *b1,
                           const double *b2, const double *b3);
/* NULL pointers are skipped; j1,j2,j3 = currents of the last Hp1, Hp2, Hp3 */
int gempic_hs2d_get_fields(gempic_handle hs, double *e1, double *e2, double *e3, double *b1, double *b2, double *b3,
                           double *j1, double *j2, double *j3);
int gempic_hs2d_operator(gempic_handle hs, int op, double dt);                         /* GEMPIC_OP_* */
int gempic_hs2d_strang_splitting(gempic_handle hs, double dt, int64_t number_steps);
/* drop-in forms with HOST field buffers (the aliased e_dofs / b_dofs): H2D -> op -> D2H -> sync */
int gempic_hs2d_operator_host(gempic_handle hs, int op, double dt, double *e1, double *e2, double *e3, double *b1,
                              double *b2, double *b3);
int gempic_hs2d_strang_splitting_host(gempic_handle hs, double dt, int64_t number_steps, double *e1, double *e2,
                                      double *e3, double *b1, double *b2, double *b3);
/* cell-sort the particles every `interval` Strang steps (0: never; default 1) */
int gempic_hs2d_set_sort_interval(gempic_handle hs, int interval);
/* 2 (default): fused passes inside strang_splitting, the operators that start from the cell-sorted order on the
 * register-resident fast path; 1: fused [HE,Hp3] tile pass and cross-step HE fold only; 0: one pass per operator */
int gempic_hs2d_set_fusion(gempic_handle hs, int fuse);
/* add_charge! of all particles onto the degree p x p dofs (get_charge weights), summed over ranks */
int gempic_hs2d_charge_density(gempic_handle hs, double *rho);
/* out[4] = sum_p w |v|^2, sum_p w v1, sum_p w v2, sum_p w v3, summed over ranks (diagnostics.jl:197-211) */
int gempic_hs2d_moments(gempic_handle hs, double *out4);
/* periodic cell sort of a ParticleGroup{2,V} on the mesh of a TwoDMaxwell */
int gempic_pg_sort2d(gempic_handle pg, gempic_handle maxwell2d);

/* ---- HamiltonianSplittingBoris (src/hamiltonian_splitting_boris.jl:23-288) ------------ */
int gempic_boris_create(gempic_handle maxwell, gempic_handle pmc0, gempic_handle pmc1, gempic_handle pg,
                        gempic_handle *out);
int gempic_boris_destroy(gempic_handle bs);
int gempic_boris_set_fields(gempic_handle bs, const double *e1, const double *e2, const double *b);
/* which = GEMPIC_F_* (E1,E2,B,J1,J2,E1_MID,E2_MID,B_MID) */
int gempic_boris_get_field(gempic_handle bs, int which, double *out);
int gempic_boris_staggering(gempic_handle bs, double dt);                              /* :99-122 */
int gempic_boris_strang_splitting(gempic_handle bs, double dt, int64_t number_steps);  /* :132-177 */
int gempic_boris_push_v_epart(gempic_handle bs, double dt);                            /* :189-204 */
int gempic_boris_push_v_bpart(gempic_handle bs, double dt);                            /* :211-233 */
int gempic_boris_push_x_accumulate_j(gempic_handle bs, double dt);                     /* :250-288 */
int gempic_boris_staggering_host(gempic_handle bs, double dt, double *e1, double *e2, double *b);
int gempic_boris_strang_splitting_host(gempic_handle bs, double dt, int64_t number_steps, double *e1,
                                       double *e2, double *b);

/* ---- diagnostics (src/diagnostics.jl) -------------------------------------------------- */
/* solve_poisson! (:15-31): rho[n] and efield[n] are host outputs */
int gempic_solve_poisson(gempic_handle pg, gempic_handle pmc0, gempic_handle maxwell, double *efield,
                         double *rho);
/* write_step! (:186-250): out[11] = Time, KineticEnergy, Momentum1, Momentum2, PotentialEnergyE1,
 * PotentialEnergyE2, PotentialEnergyB3, Transfer, VVB, Poynting, ErrorPoisson (:143-155) */
int gempic_diag_write_step(gempic_handle pg, gempic_handle maxwell, gempic_handle pmc0, gempic_handle pmc1,
                           double time, int degree, const double *e1, const double *e2, const double *b,
                           const double *e1_n, const double *e2_n, const double *e_poisson, double *out11);

#ifdef __cplusplus
}
#endif
#endif /* GEMPIC_B200_H */
