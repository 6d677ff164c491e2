// hs1d.cu -- HamiltonianSplitting{1,2} and {1,1}: the operators of
// src/hamiltonian_splitting_1d2v.jl / _1d1v.jl and strang_splitting!
// (src/hamiltonian_splitting.jl:98-108, src/hamiltonian_splitting_1d1v.jl:11-21).
//
// Each reference operator = one streaming particle pass (k_pass<Op>) + the tiny replicated
// field update.  Deposits go: block-private smem -> per-block partials -> fixed-order
// reduce -> (NCCL all-reduce over ranks) -> field solve, all on the library stream.
#include "objects.cuh"

namespace gempic {

template <class Op>
static PassParams<Op> base_params(Splitting &h)
{
    PassParams<Op> P{};
    P.r = h.pg->rows1d();
    P.n_particles = h.pg->n;
    P.m = h.mesh();
    P.copies = 0;
    P.partials = nullptr;
    return P;
}

// ---- {1,2} ----------------------------------------------------------------------------
static void op_HE_particles(Splitting &h, double dt, const double *e1, const double *e2)
{
    const double dtqm = dt * h.pg->q_over_m;
    GP_DISPATCH_DEGREES(h.ks0->degree, h.ks1->degree, {
        using Op = OpHE<D0, D1>;
        auto P = base_params<Op>(h);
        P.fields[0] = e1;
        P.fields[1] = e2;
        P.op.dtqm = dtqm;
        launch_pass<Op>(P, &h.scratch, nullptr, "operatorHE");
    });
}

static void op_HE(Splitting &h, double dt)
{
    op_HE_particles(h, dt, h.e1(), h.e2());
    field_b_from_e(*h.maxwell, h.b(), dt, h.e2());   // :218
}

static void op_HB(Splitting &h, double dt) { field_e_from_b(*h.maxwell, h.e2(), dt, h.b()); }   // :234-236

static void op_Hp2(Splitting &h, double dt)
{
    const double dtqm = dt * h.pg->q_over_m;
    GP_DISPATCH_DEGREES(h.ks0->degree, h.ks1->degree, {
        using Op = OpHp2<D0, D1>;
        auto P = base_params<Op>(h);
        P.fields[0] = h.b();
        P.op.dtqm = dtqm;
        P.op.wscale0 = h.pg->charge * h.pg->common_weight * h.ks0->scaling;
        launch_pass<Op>(P, &h.scratch, h.j2(), "operatorHp2");
    });
    h.j2_stale = h.j2_unreduced = false;   // j_dofs[2] is overwritten
    GP_CUDA(cudaMemsetAsync(h.j1(), 0, sizeof(double) * h.n, ctx().stream));   // fill!(j_dofs[1], 0) :132
    allreduce_sum(h.j2(), h.n);
    field_e_from_j(*h.maxwell, h.e2(), h.j2(), 2, dt);   // j2 .*= dt ; compute_e_from_j!(e2, j2, 2)  :173-175
}

template <class Op>
static void set_hp1_params(Splitting &h, double dt, typename Op::Params &op)
{
    const double cq = h.pg->charge * h.pg->common_weight;
    op.dt = dt;
    op.qm_dx = h.pg->q_over_m * h.ks1->delta_x;
    op.wscale0 = cq * h.ks0->scaling;
    op.wscale1_dx = cq * h.ks1->scaling * h.ks1->delta_x;
}

static void op_Hp1(Splitting &h, double dt, bool with_rho)
{
    GP_DISPATCH_DEGREES(h.ks0->degree, h.ks1->degree, {
        if (with_rho) {
            using Op = OpHp1<D0, D1, true>;
            auto P = base_params<Op>(h);
            P.fields[0] = h.b();
            set_hp1_params<Op>(h, dt, P.op);
            launch_pass<Op>(P, &h.scratch, h.j1(), "operatorHp1+rho");   // j1 | j2 are adjacent: out = [j1, rho]
            h.j2_stale = h.j2_unreduced = false;
        } else {
            using Op = OpHp1<D0, D1, false>;
            auto P = base_params<Op>(h);
            P.fields[0] = h.b();
            set_hp1_params<Op>(h, dt, P.op);
            launch_pass<Op>(P, &h.scratch, h.j1(), "operatorHp1");
        }
    });
    allreduce_sum(h.j1(), with_rho ? 2 * h.n : h.n);
    field_e_from_j(*h.maxwell, h.e1(), h.j1(), 1, 1.0);   // :111
}

// fused particle pass [HE x n_he, Hp2(dt/2), Hp1(dt), Hp2(dt/2)] + the three field solves.
// n_he = 1: fields e1,e2 = current.  n_he = 2: the first kick reads the snapshot (e1T, e2T)
// taken before the trailing HE's field update of the previous step.
static void fused_pass(Splitting &h, double dt, int n_he, double dt_T, DeferredReduce *defer)
{
    const double qm = h.pg->q_over_m;
    GP_DISPATCH_DEGREES(h.ks0->degree, h.ks1->degree, {
        if (n_he == 1) {
            using Op = OpStrangFused<D0, D1, 1>;
            auto P = base_params<Op>(h);
            P.fields[0] = h.e1();
            P.fields[1] = h.e2();
            P.fields[2] = h.b();
            P.op.dtqm_e[0] = P.op.dtqm_e[1] = 0.5 * dt * qm;
            P.op.dtqm_p2 = 0.5 * dt * qm;
            typename OpHp1<D0, D1, false>::Params hp;
            set_hp1_params<OpHp1<D0, D1, false>>(h, dt, hp);
            P.op.dt = hp.dt; P.op.qm_dx = hp.qm_dx; P.op.wscale0 = hp.wscale0; P.op.wscale1_dx = hp.wscale1_dx;
            launch_pass<Op>(P, &h.scratch, h.acc(), "fused[HE,Hp2,Hp1,Hp2]", defer);
        } else {
            using Op = OpStrangFused<D0, D1, 2>;
            auto P = base_params<Op>(h);
            P.fields[0] = h.e1T();
            P.fields[1] = h.e2T();
            P.fields[2] = h.e1();
            P.fields[3] = h.e2();
            P.fields[4] = h.b();
            P.op.dtqm_e[0] = 0.5 * dt_T * qm;   // trailing HE of the previous step (snapshot fields)
            P.op.dtqm_e[1] = 0.5 * dt * qm;
            P.op.dtqm_p2 = 0.5 * dt * qm;
            typename OpHp1<D0, D1, false>::Params hp;
            set_hp1_params<OpHp1<D0, D1, false>>(h, dt, hp);
            P.op.dt = hp.dt; P.op.qm_dx = hp.qm_dx; P.op.wscale0 = hp.wscale0; P.op.wscale1_dx = hp.wscale1_dx;
            launch_pass<Op>(P, &h.scratch, h.acc(), "fused[HE,HE,Hp2,Hp1,Hp2]", defer);
        }
    });
    if (!defer) allreduce_sum(h.acc(), 2 * h.n);
    // The field solves of the pass follow in strang_fields().  j_dofs as the reference leaves them: the last Hp2
    // zeroed j1 (:132) and holds dt/2 * j2b -- a deposit of the particle state this pass leaves behind, rebuilt by
    // materialise_j2() when somebody looks at it
    h.j2_stale = true;
    h.j2_unreduced = true;
    h.j2_scale = 0.5 * dt;
}

// Field-only updates around the fused pass, one launch (k_strang_fields, fields1d.cu):
//   solve   acc = [j2a + j2b | j1]: compute_e_from_j! is linear and nothing reads e2 between the two Hp2 half steps, so
//           their two solves (:173-175) collapse into one on the summed current; then the Hp1 solve (:111); j_dofs[1] = 0
//   tail    the trailing HE (field part, after a snapshot of e for its particle kick) and HB of a step
//   lead    the leading HB and HE (field part) of the next step
static void strang_fields(Splitting &h, bool solve, double dt, bool tail, double dt_tail, bool lead, double dt_lead,
                          const DeferredReduce *defer = nullptr)
{
    StrangFields F{};
    F.n_partials = defer ? defer->n_blocks : -1;
    F.partials = defer ? defer->partials : nullptr;
    F.x = defer ? xchg_next() : XchgDev{};
    F.n_acc = 2 * h.n;
    F.e1 = h.e1(); F.e2 = h.e2(); F.b = h.b(); F.j1 = h.j1(); F.acc = h.acc(); F.eT = h.e1T();   // e1T | e2T adjacent
    F.do_solve = solve; F.j2_scale = 0.5 * dt;
    F.do_tail = tail; F.dt_tail = 0.5 * dt_tail;
    F.do_lead = lead; F.dt_lead = 0.5 * dt_lead;
    field_strang_fields(*h.maxwell, F);
}

// j_dofs[2] after a fused pass = dt/2 * sum_p w v2 N(x) over the particles as the pass left them (the second Hp2 of
// hamiltonian_splitting_1d2v.jl:141-175 deposits after the push and does not change v2).  `kick_dt` != 0 applies the
// deferred trailing operatorHE kick in the same pass (after the deposit).
// RANK-LOCAL: this rank's share (already scaled by dt/2) lands in j2(); Splitting::j2_unreduced stays set until a
// collective entry point (hs_finish_j2) sums the shares.  pg_sync() -- reached from entry points that only one rank may
// call (download, save, a finalizer) -- therefore never issues a collective.
static void materialise_j2_pass(Splitting &h, double kick_dt)
{
    const double ws0 = h.pg->charge * h.pg->common_weight * h.ks0->scaling;
    GP_DISPATCH_DEGREES(h.ks0->degree, h.ks1->degree, {
        if (kick_dt != 0.0) {
            using Op = OpHEJ2<D0, D1>;
            auto P = base_params<Op>(h);
            P.fields[0] = h.e1T();
            P.fields[1] = h.e2T();
            P.op.dtqm = kick_dt * h.pg->q_over_m;
            P.op.wscale0 = ws0;
            launch_pass<Op>(P, &h.scratch, h.j2(), "operatorHE+j2");
        } else {
            using Op = OpJ2<D0>;
            auto P = base_params<Op>(h);
            P.op.wscale0 = ws0;
            launch_pass<Op>(P, &h.scratch, h.j2(), "j2 deposit");
        }
    });
    field_axpby(h.j2(), 0.0, h.j2(), h.j2_scale, h.n);
    h.j2_stale = false;
}

void tail_key(const Pmc1D &ks0, const Pmc1D &ks1, double Lmod, double (&key)[12])
{
    const Pmc1D *k[2] = {&ks0, &ks1};
    for (int i = 0; i < 2; ++i) {
        key[5 * i + 0] = k[i]->degree; key[5 * i + 1] = k[i]->n_grid; key[5 * i + 2] = k[i]->xmin;
        key[5 * i + 3] = k[i]->delta_x; key[5 * i + 4] = k[i]->scaling;
    }
    key[10] = Lmod;
    key[11] = 0.0;
}

// The deferred kick of a fused strang_splitting! applied together with everything a diagnostics loop body asks of the
// particles next (OpLoopTail): j_dofs[2] (rank-local share), the rho deposit of solve_poisson! and the write_step! sums.
// 48 B/particle instead of 48 (OpHEJ2) + 16 (OpCharge) + 32 (OpDiag).  RANK-LOCAL like materialise_j2_pass.
static void loop_tail_pass(Splitting &h, double kick_dt)
{
    ParticleGroup &pg = *h.pg;
    const int n = h.n;
    GP_DISPATCH_DEGREES(h.ks0->degree, h.ks1->degree, {
        using Op = OpLoopTail<D0, D1>;
        auto P = base_params<Op>(h);
        P.fields[0] = h.e1T();
        P.fields[1] = h.e2T();
        P.fields[2] = h.e1();
        P.fields[3] = h.e2();
        P.fields[4] = h.b();
        P.op.dtqm = kick_dt * pg.q_over_m;
        P.op.wscale0 = pg.charge * pg.common_weight * h.ks0->scaling;
        P.op.charge = pg.charge; P.op.mass = pg.mass; P.op.cw = pg.common_weight;
        launch_pass<Op>(P, &h.scratch, h.acc(), "loop tail [HE,rho,diag]");   // acc = [j2 | rho | 5 sums]  (2n + 5 <= 3n)
    });
    cudaStream_t st = ctx().stream;
    GP_CUDA(cudaMemcpyAsync(h.j2(), h.acc(), sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
    field_axpby(h.j2(), 0.0, h.j2(), h.j2_scale, n);
    h.j2_stale = false;
    auto &t = pg.tail;
    if (t.buf.n < (size_t)n + 8) t.buf.alloc((size_t)n + 8);
    GP_CUDA(cudaMemcpyAsync(t.buf.p, h.acc() + n, sizeof(double) * (n + 5), cudaMemcpyDeviceToDevice, st));
    tail_key(*h.ks0, *h.ks1, h.maxwell->Lx, t.key);
    t.epoch = ctx().particle_epoch;   // after the pass (which wrote the velocities)
    t.rho_valid = !pg.exposed;
    t.diag_valid = !pg.exposed && h.stash.size() == (size_t)3 * n && h.stash_epoch == h.fields_epoch;
    if (t.diag_valid) t.fields = h.stash;
}

// COLLECTIVE: j2() = the reference's j_dofs[2] on every rank.  Called by the entry points that hand j_dofs out
// (gempic_hs_get_fields with j2 != NULL); j2_unreduced only changes inside collective calls, so all ranks agree on it.
void hs_materialise_j2(Splitting &h)
{
    if (h.j2_stale) materialise_j2_pass(h, 0.0);
    if (h.j2_unreduced) {
        allreduce_sum(h.j2(), h.n);
        h.j2_unreduced = false;
    }
}

// The fused pass only pays while its three lane-private grids fit in shared memory (n <~ 35 cells at
// degree 3); larger grids run one pass per reference operator.
static bool fused_fits(Splitting &h)
{
    bool ok = false;
    GP_DISPATCH_DEGREES(h.ks0->degree, h.ks1->degree, {
        using Op = OpStrangFused<D0, D1, 2>;
        auto P = base_params<Op>(h);
        try {
            ok = plan_pass(P).lane_private;
        } catch (const Fail &) {   // the pp tables alone exceed shared memory: one pass per operator
            ok = false;
        }
    });
    return ok;
}

// strang_splitting! (hamiltonian_splitting.jl:98-108) with the particle passes fused.
// Per step the reference runs HB HE Hp2 Hp1 Hp2 HE HB (each dt/2 except Hp1).  Field-only work
// is hoisted around two kinds of particle pass:
//   first step      HB ; [b_from_e] ; fused{HE,Hp2,Hp1,Hp2}
//   between steps   snapshot eT=(e1,e2) ; b_from_e ; HB ; HB ; [b_from_e] ; fused{HE(eT),HE,Hp2,Hp1,Hp2}
//   last step       HE pass ; b_from_e ; HB
// compute_b_from_e! of the leading HE only reads e2 (which the kick also reads) and writes b
// (which the kick does not read), so b is simply advanced before the pass.
static void strang_fused(Splitting &h, double dt, int64_t steps)
{
    ParticleGroup &pg = *h.pg;
    if (pg.pending && pg.pending != &h) pg_sync(pg);
    // leading HB and HE (field part) of the first step; a pending trailing HE kick of the previous call (fields already
    // advanced, snapshot in e1T|e2T) rides in the first pass
    const bool pending = pg.pending != nullptr;
    pg.pending = nullptr;
    strang_fields(h, false, dt, false, dt, true, dt);
    // the per-block partial sums of the pass are reduced by the field kernel itself -- and, on several GPUs with mapped
    // exchange buffers, summed over the ranks in the same launch (xchg.cuh): pass + field kernel = 2 launches per step;
    // without the buffers: k_reduce_partials, ncclAllReduce, field kernel
    const bool single = ctx().n_ranks == 1 || xchg_active();
    for (int64_t s = 0; s < steps; ++s) {
        DeferredReduce dr, *defer = single ? &dr : nullptr;
        if (s == 0) fused_pass(h, dt, pending ? 2 : 1, pending ? h.pending_dt : dt, defer);
        else fused_pass(h, dt, 2, dt, defer);
        // solves of this pass, trailing HE (field part) + HB of this step, and -- between steps -- the leading HB + HE
        // of the next one.  The particle kick of the trailing HE only reads the snapshot: it is folded into the next
        // pass, or deferred to the next call of this splitting / applied by pg_sync() as soon as anybody else touches
        // the particles
        strang_fields(h, true, dt, true, dt, s + 1 < steps, dt, defer);
    }
    h.pending_dt = dt;
    pg.pending = &h;
}

Splitting::~Splitting()
{
    if (pg && pg->pending == this) pg->pending = nullptr;
    if (maxwell) release(maxwell);
    if (ks0) release(ks0);
    if (ks1) release(ks1);
    if (pg) release(pg);
}

void pg_sync(ParticleGroup &pg)
{
    if (pg.pending2d) hs2d_apply_pending(pg);
    Splitting *h = pg.pending;
    if (!h) return;
    pg.pending = nullptr;
    if (h->j2_stale && h->n >= 5) loop_tail_pass(*h, 0.5 * h->pending_dt);   // + what solve_poisson! / write_step! ask next
    else if (h->j2_stale) materialise_j2_pass(*h, 0.5 * h->pending_dt);   // the kick changes v2: rebuild j_dofs[2] first
    else op_HE_particles(*h, 0.5 * h->pending_dt, h->e1T(), h->e2T());
}

// ---- {1,1} ----------------------------------------------------------------------------
static void op_HB11(Splitting &h, double dt)
{
    GP_DISPATCH_DEGREE(h.ks1->degree, {
        using Op = OpHB11<D>;
        auto P = base_params<Op>(h);
        P.fields[0] = h.e1();
        P.op.dt = dt;
        launch_pass<Op>(P, &h.scratch, nullptr, "operatorHB{1,1}");
    });
}

static void op_Hp111(Splitting &h, double dt)
{
    GP_DISPATCH_DEGREE(h.ks1->degree, {
        using Op = OpHp111<D>;
        auto P = base_params<Op>(h);
        P.op = {dt, h.pg->charge * h.pg->common_weight * h.ks1->scaling * h.ks1->delta_x};
        launch_pass<Op>(P, &h.scratch, h.j1(), "operatorHp1{1,1}");
    });
    GP_CUDA(cudaMemsetAsync(h.j2(), 0, sizeof(double) * h.n, ctx().stream));   // fill!(j_dofs[2], 0) 1d1v.jl:68
    h.j2_stale = h.j2_unreduced = false;
    allreduce_sum(h.j1(), h.n);
    field_e_from_j(*h.maxwell, h.e1(), h.j1(), 1, 1.0);   // 1d1v.jl:96
}

void hs_operator(Splitting &h, int op, double dt, bool inside_strang)
{
    h.fields_epoch++;
    pg_sync(*h.pg);
    if (h.V == 2) {
        switch (op) {
        case GEMPIC_OP_HP1: op_Hp1(h, dt, !inside_strang); break;
        case GEMPIC_OP_HP2: op_Hp2(h, dt); break;
        case GEMPIC_OP_HE: op_HE(h, dt); break;
        case GEMPIC_OP_HB: op_HB(h, dt); break;
        default: fail(GEMPIC_EINVAL, "unknown operator %d", op);
        }
    } else {
        switch (op) {
        case GEMPIC_OP_HP1: op_Hp111(h, dt); break;
        case GEMPIC_OP_HB: op_HB11(h, dt); break;
        default: fail(GEMPIC_EINVAL, "operator %d is not defined for HamiltonianSplitting{1,1}", op);
        }
    }
}

static void strang_step(Splitting &h, double dt)
{
    if (h.V == 2) {   // hamiltonian_splitting.jl:98-108
        hs_operator(h, GEMPIC_OP_HB, 0.5 * dt, true);
        hs_operator(h, GEMPIC_OP_HE, 0.5 * dt, true);
        hs_operator(h, GEMPIC_OP_HP2, 0.5 * dt, true);
        hs_operator(h, GEMPIC_OP_HP1, 1.0 * dt, true);
        hs_operator(h, GEMPIC_OP_HP2, 0.5 * dt, true);
        hs_operator(h, GEMPIC_OP_HE, 0.5 * dt, true);
        hs_operator(h, GEMPIC_OP_HB, 0.5 * dt, true);
    } else {          // hamiltonian_splitting_1d1v.jl:11-21
        hs_operator(h, GEMPIC_OP_HB, 0.5 * dt, true);
        hs_operator(h, GEMPIC_OP_HP1, dt, true);
        hs_operator(h, GEMPIC_OP_HB, 0.5 * dt, true);
    }
}

void hs_strang(Splitting &h, double dt, int64_t steps)
{
    if (steps <= 0) return;
    h.fields_epoch++;
    if (h.fuse && h.V == 2 && fused_fits(h)) {
        strang_fused(h, dt, steps);
        return;
    }
    pg_sync(*h.pg);
    for (int64_t s = 0; s < steps; ++s) strang_step(h, dt);
}

}  // namespace gempic
