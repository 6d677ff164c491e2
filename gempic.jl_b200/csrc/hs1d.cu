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
    P.n_acc = 0;
    P.copies = 0;
    P.partials = nullptr;
    return P;
}

// ---- {1,2} ----------------------------------------------------------------------------
static void op_HE(Splitting &h, double dt)
{
    const double dtqm = dt * h.pg->q_over_m;
    GP_DISPATCH_DEGREES(h.ks0->degree, h.ks1->degree, {
        using Op = OpHE<D0, D1>;
        auto P = base_params<Op>(h);
        P.fields[0] = h.e1();
        P.fields[1] = h.e2();
        P.op.dtqm = dtqm;
        launch_pass<Op>(P, &h.scratch, nullptr, "operatorHE");
    });
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
        P.n_acc = h.n;
        P.op.dtqm = dtqm;
        P.op.charge = h.pg->charge;
        P.op.cw = h.pg->common_weight;
        P.op.scaling0 = h.ks0->scaling;
        launch_pass<Op>(P, &h.scratch, h.j2(), "operatorHp2");
    });
    GP_CUDA(cudaMemsetAsync(h.j1(), 0, sizeof(double) * h.n, ctx().stream));   // fill!(j_dofs[1], 0) :132
    allreduce_sum(h.j2(), h.n);
    field_e_from_j(*h.maxwell, h.e2(), h.j2(), 2, dt);   // j2 .*= dt ; compute_e_from_j!(e2, j2, 2)  :173-175
}

static void op_Hp1(Splitting &h, double dt, bool with_rho)
{
    GP_DISPATCH_DEGREES(h.ks0->degree, h.ks1->degree, {
        if (with_rho) {
            using Op = OpHp1<D0, D1, true>;
            auto P = base_params<Op>(h);
            P.fields[0] = h.b();
            P.n_acc = 2 * h.n;
            P.op = {dt, h.pg->q_over_m, h.pg->charge, h.pg->common_weight, h.ks0->scaling, h.ks1->scaling};
            launch_pass<Op>(P, &h.scratch, h.j1(), "operatorHp1+rho");   // j1 | j2 are adjacent: acc = [j1, rho]
        } else {
            using Op = OpHp1<D0, D1, false>;
            auto P = base_params<Op>(h);
            P.fields[0] = h.b();
            P.n_acc = h.n;
            P.op = {dt, h.pg->q_over_m, h.pg->charge, h.pg->common_weight, h.ks0->scaling, h.ks1->scaling};
            launch_pass<Op>(P, &h.scratch, h.j1(), "operatorHp1");
        }
    });
    allreduce_sum(h.j1(), with_rho ? 2 * h.n : h.n);
    field_e_from_j(*h.maxwell, h.e1(), h.j1(), 1, 1.0);   // :111
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
        P.n_acc = h.n;
        P.op = {dt, h.pg->charge, h.pg->common_weight, h.ks1->scaling};
        launch_pass<Op>(P, &h.scratch, h.j1(), "operatorHp1{1,1}");
    });
    GP_CUDA(cudaMemsetAsync(h.j2(), 0, sizeof(double) * h.n, ctx().stream));   // fill!(j_dofs[2], 0) 1d1v.jl:68
    allreduce_sum(h.j1(), h.n);
    field_e_from_j(*h.maxwell, h.e1(), h.j1(), 1, 1.0);   // 1d1v.jl:96
}

void hs_operator(Splitting &h, int op, double dt, bool inside_strang)
{
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
    for (int64_t s = 0; s < steps; ++s) strang_step(h, dt);
}

}  // namespace gempic
