// pmc1d.cu -- batched ParticleMeshCoupling1D entry points on plain device arrays
// (src/particle_mesh_coupling_1d.jl: add_charge! :261-280, evaluate :438-453,
// add_current_update_v! :296-376 and its 1d1v form :471-529).
#include "objects.cuh"

namespace gempic {

// out (v1 row) = field(x)
template <int D>
struct OpEval {
    static constexpr int READ = ROW_X, WRITE = ROW_V1;
    static constexpr int NF = 1, NG = 0, NS = 0;
    static constexpr bool DEPOSIT = false;
    struct Params { int unused; };
    template <bool LP>
    static __device__ __forceinline__ void apply(Particle &p, const PassParams<OpEval> &P, const double *sf, const Acc<LP> &)
    {
        const Pos ps = locate(p.x, P.m);
        double b[D + 1];
        basis_pp<D>(ps.t, b);
        p.v1 = gather_h<D>(sf, first_dof<D>(ps, P.m), b);
    }
};

// rows: x = x_old, v1 = x_new, v2 = v (updated), w = marker charge
template <int D, bool WITH_B>
struct OpCurrent {
    static constexpr int READ = ROW_X | ROW_V1 | ROW_W | (WITH_B ? ROW_V2 : 0), WRITE = WITH_B ? ROW_V2 : 0;
    static constexpr int NF = WITH_B ? 1 : 0, NG = 1, NS = 0;
    static constexpr bool DEPOSIT = true;
    struct Params { double qm_dx, scaling_dx; };
    template <bool LP>
    static __device__ __forceinline__ void apply(Particle &p, const PassParams<OpCurrent> &P, const double *sf, const Acc<LP> &acc)
    {
        const Pos po = locate(p.x, P.m);
        Pos pn;
        if (WITH_B) {
            pn = locate(p.v1, P.m);
            p.v2 = current_update_v<D, LP, true>(acc, 0, sf, po, pn, p.w * P.op.scaling_dx, P.op.qm_dx, p.v2, P.m);
        } else {
            cell_offset_floor(p.v1, P.m, pn.c, pn.t);   // 1d1v form: floor for the new index (:487)
            current_update_v<D, LP, false>(acc, 0, nullptr, po, pn, p.w * P.op.scaling_dx, 0.0, 0.0, P.m);
        }
    }
};

void pmc1d_add_charge_dev(Pmc1D &p, const double *x, const double *w, int64_t n, double charge, double cw, double *rho_out)
{
    GP_DISPATCH_DEGREE(p.degree, {
        using Op = OpCharge<D>;
        PassParams<Op> P{};
        P.r.x = const_cast<double *>(x);
        P.r.w = const_cast<double *>(w);
        P.n_particles = n;
        P.m = p.mesh(p.Lx);
        P.op = {charge * cw * p.scaling};
        launch_pass<Op>(P, &p.scratch, rho_out);
    });
}

void pmc1d_evaluate_dev(Pmc1D &p, const double *x, int64_t n, const double *field, double *out)
{
    GP_DISPATCH_DEGREE(p.degree, {
        using Op = OpEval<D>;
        PassParams<Op> P{};
        P.r.x = const_cast<double *>(x);
        P.r.v1 = out;
        P.n_particles = n;
        P.m = p.mesh(p.Lx);
        P.fields[0] = field;
        launch_pass<Op>(P, &p.scratch, nullptr);
    });
}

void pmc1d_add_current_dev(Pmc1D &p, const double *x_old, const double *x_new, const double *w, double qm,
                           const double *bfield, double *v, int64_t n, double *j_out)
{
    GP_DISPATCH_DEGREE(p.degree, {
        if (bfield) {
            using Op = OpCurrent<D, true>;
            PassParams<Op> P{};
            P.r.x = const_cast<double *>(x_old);
            P.r.v1 = const_cast<double *>(x_new);
            P.r.v2 = v;
            P.r.w = const_cast<double *>(w);
            P.n_particles = n;
            P.m = p.mesh(p.Lx);
            P.fields[0] = bfield;
            P.op = {qm * p.delta_x, p.scaling * p.delta_x};
            launch_pass<Op>(P, &p.scratch, j_out);
        } else {
            using Op = OpCurrent<D, false>;
            PassParams<Op> P{};
            P.r.x = const_cast<double *>(x_old);
            P.r.v1 = const_cast<double *>(x_new);
            P.r.w = const_cast<double *>(w);
            P.n_particles = n;
            P.m = p.mesh(p.Lx);
            P.op = {qm * p.delta_x, p.scaling * p.delta_x};
            launch_pass<Op>(P, &p.scratch, j_out);
        }
    });
}

// write_step! particle sums (src/diagnostics.jl:45-92,197-211)
void diag_particle_sums(ParticleGroup &pg, Pmc1D &ks0, Pmc1D &ks1, const Maxwell1D &m, const double *e1,
                        const double *e2, const double *b, PartialScratch &scratch, double *out5)
{
    GP_DISPATCH_DEGREES(ks0.degree, ks1.degree, {
        using Op = OpDiag<D0, D1>;
        PassParams<Op> P{};
        P.r = pg.rows1d();
        P.n_particles = pg.n;
        P.m = ks0.mesh(m.Lx);
        P.fields[0] = e1;
        P.fields[1] = e2;
        P.fields[2] = b;
        P.op = {pg.charge, pg.mass, pg.common_weight};
        launch_pass<Op>(P, &scratch, out5);
    });
    allreduce_sum(out5, 5);
}

}  // namespace gempic
