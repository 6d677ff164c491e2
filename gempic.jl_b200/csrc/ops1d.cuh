// ops1d.cuh -- per-particle operators of the 1D (1d1v / 1d2v) paths, plugged into k_pass.
//
// Template parameters D0 / D1 are the spline degrees of kernel_smoother_0 / kernel_smoother_1
// (src/hamiltonian_splitting.jl:23-24): e2, j2 and rho live on degree D0, e1, b and j1 on
// degree D1 (src/hamiltonian_splitting_1d2v.jl:70,102,151,163,207-208).
#pragma once
#include "pass.cuh"

namespace gempic {

// rho[(cell-D+k) mod n] += (w * N_k(t)) * scaling        (add_charge!, pmc1d.jl:261-280)
template <int D, bool LP>
__device__ __forceinline__ void deposit(const Acc<LP> &acc, int goff, int cell, const double (&b)[D + 1], double w,
                                        double scaling, const Mesh1D &m)
{
    int g = wrap_index(cell - D, m);
#pragma unroll
    for (int k = 0; k <= D; ++k) {
        acc.add(goff + g, w * b[k] * scaling);
        g = wrap_next(g + 1, m.n);
    }
}

// one in-cell segment of the current line integral (update_jv!, pmc1d.jl:385-425 / :538-580)
template <int D, bool LP, bool WITH_B>
__device__ __forceinline__ double update_jv(const Acc<LP> &acc, const double *__restrict__ bfield, double lower,
                                            double upper, int cell, double w, double qm, double sign_dx,
                                            double scaling, double vi, const Mesh1D &m)
{
    double s[D + 1];
    segment_weights<D>(lower, upper, sign_dx, s);
    int g = wrap_index(cell - D, m);
#pragma unroll
    for (int k = 0; k <= D; ++k) {
        acc.add(g, w * s[k] * scaling);
        if (WITH_B) vi = vi - qm * s[k] * bfield[g];
        g = wrap_next(g + 1, m.n);
    }
    return vi;
}

// add_current_update_v! (pmc1d.jl:296-376); NEW_FLOOR selects the 1d1v variant (:471-529)
// whose new index uses floor (:487).  Segment order is the reference's: all lanes run
// segment A together, crossing lanes then run B, multi-cell crossers (rare) the loop.
template <int D, bool LP, bool WITH_B, bool NEW_FLOOR>
__device__ __forceinline__ double add_current_update_v(const Acc<LP> &acc, const double *__restrict__ bfield,
                                                       double x_old, double x_new, double w, double qm,
                                                       double scaling, double vi, const Mesh1D &m)
{
    int i_old, i_new;
    double r_old, r_new;
    cell_offset(x_old, m, i_old, r_old);
    if (NEW_FLOOR) {
        const double xi = (x_new - m.xmin) / m.dx;
        i_new = __double2int_rd(xi);
        r_new = xi - (double)i_new;
    } else {
        cell_offset(x_new, m, i_new, r_new);
    }
    double lo, up, sgn;
    int cell;
    if (i_old == i_new) {
        const bool fwd = r_old < r_new;
        lo = fwd ? r_old : r_new;
        up = fwd ? r_new : r_old;
        sgn = fwd ? m.dx : -m.dx;
        cell = i_old;
    } else if (i_old < i_new) {
        lo = r_old; up = 1.0; sgn = m.dx; cell = i_old;
    } else {
        lo = r_new; up = 1.0; sgn = -m.dx; cell = i_new;
    }
    vi = update_jv<D, LP, WITH_B>(acc, bfield, lo, up, cell, w, qm, sgn, scaling, vi, m);
    if (i_old != i_new) {
        if (i_old < i_new) {
            vi = update_jv<D, LP, WITH_B>(acc, bfield, 0.0, r_new, i_new, w, qm, m.dx, scaling, vi, m);
            for (int c = i_old + 1; c <= i_new - 1; ++c)
                vi = update_jv<D, LP, WITH_B>(acc, bfield, 0.0, 1.0, c, w, qm, m.dx, scaling, vi, m);
        } else {
            vi = update_jv<D, LP, WITH_B>(acc, bfield, 0.0, r_old, i_old, w, qm, -m.dx, scaling, vi, m);
            for (int c = i_new + 1; c <= i_old - 1; ++c)
                vi = update_jv<D, LP, WITH_B>(acc, bfield, 0.0, 1.0, c, w, qm, -m.dx, scaling, vi, m);
        }
    }
    return vi;
}

// ---- operatorHE{1,2}  (hamiltonian_splitting_1d2v.jl:198-215), also Boris push_v_epart!
//      (hamiltonian_splitting_boris.jl:189-204).  fields: [e1, e2]
template <int D0, int D1>
struct OpHE {
    static constexpr int READ = ROW_X | ROW_V1 | ROW_V2, WRITE = ROW_V1 | ROW_V2;
    static constexpr int NF = 2;
    static constexpr bool DEPOSIT = false;
    struct Params { double dtqm; };
    template <bool LP>
    static __device__ __forceinline__ void apply(Particle &p, const PassParams<OpHE> &P, const double *sf, const Acc<LP> &)
    {
        int c;
        double t;
        cell_offset(p.x, P.m, c, t);
        double b1[D1 + 1], b0[D0 + 1];
        bspline_basis<D1>(t, b1);
        bspline_basis<D0>(t, b0);
        const double e1 = gather<D1>(sf, c, b1, P.m);
        const double e2 = gather<D0>(sf + P.m.n, c, b0, P.m);
        p.v1 = p.v1 + P.op.dtqm * e1;
        p.v2 = p.v2 + P.op.dtqm * e2;
    }
};

// ---- operatorHp2{1,2}  (hamiltonian_splitting_1d2v.jl:141-167).  fields: [b]; deposit j2 (D0)
template <int D0, int D1>
struct OpHp2 {
    static constexpr int READ = ROW_X | ROW_V1 | ROW_V2 | ROW_W, WRITE = ROW_V1;
    static constexpr int NF = 1;
    static constexpr bool DEPOSIT = true;
    struct Params { double dtqm, charge, cw, scaling0; };
    template <bool LP>
    static __device__ __forceinline__ void apply(Particle &p, const PassParams<OpHp2> &P, const double *sf, const Acc<LP> &acc)
    {
        int c;
        double t;
        cell_offset(p.x, P.m, c, t);
        double b1[D1 + 1], b0[D0 + 1];
        bspline_basis<D1>(t, b1);
        bspline_basis<D0>(t, b0);
        const double b = gather<D1>(sf, c, b1, P.m);
        p.v1 = p.v1 + P.op.dtqm * p.v2 * b;
        double w = p.w * P.op.charge;
        w = w * P.op.cw;
        w = w * p.v2;
        deposit<D0, LP>(acc, 0, c, b0, w, P.op.scaling0, P.m);
    }
};

// ---- operatorHp1{1,2}  (hamiltonian_splitting_1d2v.jl:48-106).  fields: [b];
//      deposit j1 (D1) at acc[0..n) and, when RHO, rho (D0) at acc[n..2n).
//      RHO=false is used inside strang_splitting!, where that rho is dead data.
template <int D0, int D1, bool RHO>
struct OpHp1 {
    static constexpr int READ = ROW_X | ROW_V1 | ROW_V2 | ROW_W, WRITE = ROW_X | ROW_V2;
    static constexpr int NF = 1;
    static constexpr bool DEPOSIT = true;
    struct Params { double dt, qm, charge, cw, scaling0, scaling1; };
    template <bool LP>
    static __device__ __forceinline__ void apply(Particle &p, const PassParams<OpHp1> &P, const double *sf, const Acc<LP> &acc)
    {
        const double x_new = p.x + P.op.dt * p.v1;
        double wi = p.w * P.op.charge;
        wi = wi * P.op.cw;
        p.v2 = add_current_update_v<D1, LP, true, false>(acc, sf, p.x, x_new, wi, P.op.qm, P.op.scaling1, p.v2, P.m);
        p.x = mod_julia(x_new, P.m.Lx);
        if (RHO) {
            int c;
            double t;
            cell_offset(p.x, P.m, c, t);
            double b0[D0 + 1];
            bspline_basis<D0>(t, b0);
            deposit<D0, LP>(acc, P.m.n, c, b0, wi, P.op.scaling0, P.m);
        }
    }
};

// ---- 1d1v operatorHB = E-kick without q/m (hamiltonian_splitting_1d1v.jl:113-125). fields: [e1]
template <int D1>
struct OpHB11 {
    static constexpr int READ = ROW_X | ROW_V1, WRITE = ROW_V1;
    static constexpr int NF = 1;
    static constexpr bool DEPOSIT = false;
    struct Params { double dt; };
    template <bool LP>
    static __device__ __forceinline__ void apply(Particle &p, const PassParams<OpHB11> &P, const double *sf, const Acc<LP> &)
    {
        int c;
        double t;
        cell_offset(p.x, P.m, c, t);
        double b1[D1 + 1];
        bspline_basis<D1>(t, b1);
        p.v1 = p.v1 + P.op.dt * gather<D1>(sf, c, b1, P.m);
    }
};

// ---- 1d1v operatorHp1 (hamiltonian_splitting_1d1v.jl:70-94): j1 only, v unchanged
template <int D1>
struct OpHp111 {
    static constexpr int READ = ROW_X | ROW_V1 | ROW_W, WRITE = ROW_X;
    static constexpr int NF = 0;
    static constexpr bool DEPOSIT = true;
    struct Params { double dt, charge, cw, scaling1; };
    template <bool LP>
    static __device__ __forceinline__ void apply(Particle &p, const PassParams<OpHp111> &P, const double *, const Acc<LP> &acc)
    {
        const double x_new = p.x + P.op.dt * p.v1;
        const double wi = P.op.charge * p.w * P.op.cw;  // get_charge (particle_group.jl:67-69)
        add_current_update_v<D1, LP, false, true>(acc, nullptr, p.x, x_new, wi, 0.0, P.op.scaling1, p.v1, P.m);
        p.x = mod_julia(x_new, P.m.Lx);
    }
};

// ---- Boris push_v_bpart! (hamiltonian_splitting_boris.jl:211-233). fields: [b_mid]
template <int D1>
__device__ __forceinline__ void boris_rotate(Particle &p, double bfield, double qmdt)
{
    bfield = qmdt * bfield;
    double M11 = 1.0 / (1.0 + bfield * bfield);
    const double M12 = M11 * bfield * 2.0;
    M11 = M11 * (1 - bfield * bfield);
    const double v1 = M11 * p.v1 + M12 * p.v2;
    const double v2 = -M12 * p.v1 + M11 * p.v2;
    p.v1 = v1;
    p.v2 = v2;
}
template <int D1>
struct OpBorisB {
    static constexpr int READ = ROW_X | ROW_V1 | ROW_V2, WRITE = ROW_V1 | ROW_V2;
    static constexpr int NF = 1;
    static constexpr bool DEPOSIT = false;
    struct Params { double qmdt; };
    template <bool LP>
    static __device__ __forceinline__ void apply(Particle &p, const PassParams<OpBorisB> &P, const double *sf, const Acc<LP> &)
    {
        int c;
        double t;
        cell_offset(p.x, P.m, c, t);
        double b1[D1 + 1];
        bspline_basis<D1>(t, b1);
        boris_rotate<D1>(p, gather<D1>(sf, c, b1, P.m), P.op.qmdt);
    }
};

// ---- Boris push_x_accumulate_j! (hamiltonian_splitting_boris.jl:250-288):
//      deposits w*v1 (D1) -> acc[0..n) and w*v2 (D0) -> acc[n..2n) at the un-wrapped midpoint
template <int D0, int D1, bool LP>
__device__ __forceinline__ void boris_push_x(Particle &p, double dt, double charge, double cw, double scaling0,
                                             double scaling1, const Acc<LP> &acc, const Mesh1D &m)
{
    const double x_new = p.x + dt * p.v1;
    const double wi = charge * p.w * cw;
    int c;
    double t;
    cell_offset((p.x + x_new) * 0.5, m, c, t);
    double b1[D1 + 1], b0[D0 + 1];
    bspline_basis<D1>(t, b1);
    bspline_basis<D0>(t, b0);
    deposit<D1, LP>(acc, 0, c, b1, wi * p.v1, scaling1, m);
    deposit<D0, LP>(acc, m.n, c, b0, wi * p.v2, scaling0, m);
    p.x = mod_julia(x_new, m.Lx);
}
template <int D0, int D1>
struct OpBorisX {
    static constexpr int READ = ROW_X | ROW_V1 | ROW_V2 | ROW_W, WRITE = ROW_X;
    static constexpr int NF = 0;
    static constexpr bool DEPOSIT = true;
    struct Params { double dt, charge, cw, scaling0, scaling1; };
    template <bool LP>
    static __device__ __forceinline__ void apply(Particle &p, const PassParams<OpBorisX> &P, const double *, const Acc<LP> &acc)
    {
        boris_push_x<D0, D1, LP>(p, P.op.dt, P.op.charge, P.op.cw, P.op.scaling0, P.op.scaling1, acc, P.m);
    }
};

// ---- one whole Boris step per particle: epart(dt/2), bpart(dt), epart(dt/2), push_x(dt)
//      (hamiltonian_splitting_boris.jl:146-155; no field changes in between, so the four
//      reference loops collapse into one pass: 56 B/particle instead of 160).
//      fields: [e1_mid, e2_mid, b_mid]
template <int D0, int D1>
struct OpBorisStep {
    static constexpr int READ = ROW_X | ROW_V1 | ROW_V2 | ROW_W, WRITE = ROW_X | ROW_V1 | ROW_V2;
    static constexpr int NF = 3;
    static constexpr bool DEPOSIT = true;
    struct Params { double dt, half_dtqm, qmdt, charge, cw, scaling0, scaling1; };
    template <bool LP>
    static __device__ __forceinline__ void apply(Particle &p, const PassParams<OpBorisStep> &P, const double *sf, const Acc<LP> &acc)
    {
        int c;
        double t;
        cell_offset(p.x, P.m, c, t);
        double b1[D1 + 1], b0[D0 + 1];
        bspline_basis<D1>(t, b1);
        bspline_basis<D0>(t, b0);
        const double e1 = gather<D1>(sf, c, b1, P.m);
        const double e2 = gather<D0>(sf + P.m.n, c, b0, P.m);
        const double bf = gather<D1>(sf + 2 * P.m.n, c, b1, P.m);
        p.v1 = p.v1 + P.op.half_dtqm * e1;
        p.v2 = p.v2 + P.op.half_dtqm * e2;
        boris_rotate<D1>(p, bf, P.op.qmdt);
        p.v1 = p.v1 + P.op.half_dtqm * e1;
        p.v2 = p.v2 + P.op.half_dtqm * e2;
        boris_push_x<D0, D1, LP>(p, P.op.dt, P.op.charge, P.op.cw, P.op.scaling0, P.op.scaling1, acc, P.m);
    }
};

// ---- add_charge! over a ParticleGroup with marker charge get_charge (diagnostics.jl:24-28)
template <int D>
struct OpCharge {
    static constexpr int READ = ROW_X | ROW_W, WRITE = 0;
    static constexpr int NF = 0;
    static constexpr bool DEPOSIT = true;
    struct Params { double charge, cw, scaling; };
    template <bool LP>
    static __device__ __forceinline__ void apply(Particle &p, const PassParams<OpCharge> &P, const double *, const Acc<LP> &acc)
    {
        int c;
        double t;
        cell_offset(p.x, P.m, c, t);
        double b[D + 1];
        bspline_basis<D>(t, b);
        deposit<D, LP>(acc, 0, c, b, P.op.charge * p.w * P.op.cw, P.op.scaling, P.m);
    }
};

// ---- write_step! particle sums (diagnostics.jl:45-92,197-211): acc = [KE, P1, P2, transfer, vvb]
//      fields: [e1, e2, b]
template <int D0, int D1>
struct OpDiag {
    static constexpr int READ = ROW_X | ROW_V1 | ROW_V2 | ROW_W, WRITE = 0;
    static constexpr int NF = 3;
    static constexpr bool DEPOSIT = true;
    struct Params { double charge, mass, cw; };
    template <bool LP>
    static __device__ __forceinline__ void apply(Particle &p, const PassParams<OpDiag> &P, const double *sf, const Acc<LP> &acc)
    {
        double wm = p.w * P.op.mass;
        wm *= P.op.cw;
        acc.add(0, (p.v1 * p.v1 + p.v2 * p.v2) * wm);
        acc.add(1, p.v1 * wm);
        acc.add(2, p.v2 * wm);
        int c;
        double t;
        cell_offset(p.x, P.m, c, t);
        double b1[D1 + 1], b0[D0 + 1];
        bspline_basis<D1>(t, b1);
        bspline_basis<D0>(t, b0);
        const double e1 = gather<D1>(sf, c, b1, P.m);
        const double e2 = gather<D0>(sf + P.m.n, c, b0, P.m);
        const double bf = gather<D1>(sf + 2 * P.m.n, c, b1, P.m);
        const double wq = P.op.charge * p.w * P.op.cw;  // get_charge
        acc.add(3, (p.v1 * e1 + p.v2 * e2) * wq);
        double wv = p.w * P.op.charge;
        wv *= P.op.cw;
        acc.add(4, wv * p.v1 * p.v2 * bf);
    }
};

}  // namespace gempic
