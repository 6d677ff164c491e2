// boris.cu -- HamiltonianSplittingBoris (src/hamiltonian_splitting_boris.jl).
//
// strang_splitting! runs push_v_epart!(dt/2), push_v_bpart!(dt), push_v_epart!(dt/2) and
// push_x_accumulate_j!(dt) back to back with no field update in between (:146-155), so one
// fused pass (OpBorisStep) performs the whole particle part of a step: 56 B/particle instead
// of 4 x 40.  The individual pushes remain available as separate entry points.
#include "objects.cuh"

namespace gempic {

template <class Op>
static PassParams<Op> base_params(Boris &s)
{
    PassParams<Op> P{};
    P.r = s.pg->rows1d();
    P.n_particles = s.pg->n;
    P.m = s.mesh();
    return P;
}

Boris::~Boris()
{
    if (maxwell) release(maxwell);
    if (ks0) release(ks0);
    if (ks1) release(ks1);
    if (pg) release(pg);
}

void boris_push_v_epart(Boris &s, double dt)   // :189-204
{
    const double dtqm = dt * s.pg->q_over_m;
    GP_DISPATCH_DEGREES(s.ks0->degree, s.ks1->degree, {
        using Op = OpHE<D0, D1>;
        auto P = base_params<Op>(s);
        P.fields[0] = s.f(GEMPIC_F_E1_MID);
        P.fields[1] = s.f(GEMPIC_F_E2_MID);
        P.op.dtqm = dtqm;
        launch_pass<Op>(P, &s.scratch, nullptr, "push_v_epart");
    });
}

void boris_push_v_bpart(Boris &s, double dt)   // :211-233
{
    const double qmdt = s.pg->q_over_m * 0.5 * dt;
    GP_DISPATCH_DEGREE(s.ks1->degree, {
        using Op = OpBorisB<D>;
        auto P = base_params<Op>(s);
        P.fields[0] = s.f(GEMPIC_F_B_MID);
        P.op.qmdt = qmdt;
        launch_pass<Op>(P, &s.scratch, nullptr, "push_v_bpart");
    });
}

void boris_push_x_accumulate_j(Boris &s, double dt)   // :250-288
{
    GP_DISPATCH_DEGREES(s.ks0->degree, s.ks1->degree, {
        using Op = OpBorisX<D0, D1>;
        auto P = base_params<Op>(s);
        const double cq = s.pg->charge * s.pg->common_weight;
        P.op = {dt, cq * s.ks0->scaling, cq * s.ks1->scaling};
        launch_pass<Op>(P, &s.scratch, s.f(GEMPIC_F_J1), "push_x_accumulate_j");   // j1 | j2 adjacent
    });
    allreduce_sum(s.f(GEMPIC_F_J1), 2 * s.n);
}

// (4) of strang_splitting! / staggering!: the field part after the particle push
static void boris_fields_after_push(Boris &s, double dt_scale, double dt_b)
{
    const Maxwell1D &m = *s.maxwell;
    field_e_from_j(m, s.f(GEMPIC_F_E1_MID), s.f(GEMPIC_F_J1), 1, dt_scale);   // j1 .*= dt ; e1_mid
    field_e_from_b(m, s.f(GEMPIC_F_E2_MID), dt_b, s.f(GEMPIC_F_B));
    field_e_from_j(m, s.f(GEMPIC_F_E2_MID), s.f(GEMPIC_F_J2), 2, dt_scale);
}

void boris_staggering(Boris &s, double dt)   // :99-122
{
    boris_push_x_accumulate_j(s, dt * 0.5);
    field_copy(s.f(GEMPIC_F_E1_MID), s.f(GEMPIC_F_E1), 2 * s.n);   // e_mid .= e (E1_MID|E2_MID and E1|E2 adjacent)
    boris_fields_after_push(s, 0.5 * dt, 0.5 * dt);
}

static void boris_fields(Boris &s, bool post, bool pre, double dt, const DeferredReduce *defer = nullptr)
{
    BorisFields F{};
    F.n_partials = defer ? defer->n_blocks : -1;
    F.partials = defer ? defer->partials : nullptr;
    F.x = defer ? xchg_next() : XchgDev{};
    F.e1 = s.f(GEMPIC_F_E1); F.e2 = s.f(GEMPIC_F_E2); F.b = s.f(GEMPIC_F_B);
    F.j1 = s.f(GEMPIC_F_J1); F.j2 = s.f(GEMPIC_F_J2);
    F.e1_mid = s.f(GEMPIC_F_E1_MID); F.e2_mid = s.f(GEMPIC_F_E2_MID); F.b_mid = s.f(GEMPIC_F_B_MID);
    F.do_post = post; F.do_pre = pre;
    F.dt_post = F.dt_pre = dt;
    field_boris_fields(*s.maxwell, F);
}

// (2)+(3) of a step: the fused particle pass while its lane-private grids and pp tables fit in shared memory
// (n <~ 40 cells at degree 3), else the four reference loops one by one
static void boris_particles(Boris &s, double dt, DeferredReduce *defer)
{
    bool fused = false;
    GP_DISPATCH_DEGREES(s.ks0->degree, s.ks1->degree, {
        using Op = OpBorisStep<D0, D1>;
        auto P = base_params<Op>(s);
        const size_t need = ((size_t)Op::NF * (P.m.n + Op::FIELD_HALO) * Op::FC + (size_t)acc_slots<Op>(P.m.n) * 32 * (Op::THREADS / 32)) *
                            sizeof(double);
        if (need <= kSmemMaxOptin) {
            fused = true;
            P.fields[0] = s.f(GEMPIC_F_E1_MID);
            P.fields[1] = s.f(GEMPIC_F_E2_MID);
            P.fields[2] = s.f(GEMPIC_F_B_MID);
            const double cq = s.pg->charge * s.pg->common_weight;
            P.op = {dt, (0.5 * dt) * s.pg->q_over_m, s.pg->q_over_m * 0.5 * dt, cq * s.ks0->scaling, cq * s.ks1->scaling};
            launch_pass<Op>(P, &s.scratch, s.f(GEMPIC_F_J1), "boris_step", defer);
        }
    });
    if (fused) {
        if (!defer) allreduce_sum(s.f(GEMPIC_F_J1), 2 * s.n);
    } else {
        if (defer) defer->n_blocks = -1;   // the separate passes reduce (and all-reduce) j1, j2 themselves
        boris_push_v_epart(s, 0.5 * dt);
        boris_push_v_bpart(s, dt);
        boris_push_v_epart(s, 0.5 * dt);
        boris_push_x_accumulate_j(s, dt);
    }
}

// strang_splitting! :132-177.  Per step: (1) b_mid = b ; b += dt/dx D e2_mid ; b_mid = (b_mid + b)/2, (2)+(3) the particle
// pass, (4) e = e_mid and the three field solves.  (4) of a step and (1) of the next are one launch (k_boris_fields).
void boris_strang(Boris &s, double dt, int64_t steps)
{
    if (steps <= 0) return;
    boris_fields(s, false, true, dt);
    // one GPU: the per-block partial sums of the pass are reduced by the field kernel itself (one launch less)
    const bool single = ctx().n_ranks == 1 || xchg_active();   // several GPUs: summed over the ranks in the same kernel (xchg.cuh)
    for (int64_t i = 0; i < steps; ++i) {
        DeferredReduce dr, *defer = single ? &dr : nullptr;
        boris_particles(s, dt, defer);
        boris_fields(s, true, i + 1 < steps, dt, defer && defer->n_blocks >= 0 ? defer : nullptr);
    }
}

}  // namespace gempic
