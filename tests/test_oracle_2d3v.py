"""Pins the (unpinned-by-the-reference) 2d3v oracle, oracle/splitting2d3v.py:

  (a) operator by operator against the golden-pinned 1d2v oracle: with x2-independent field dofs the 2D
      operators HE, Hp1, Hp2 must reproduce the particle updates and the (x2-summed) currents of
      operatorHE/Hp1/Hp2 of src/hamiltonian_splitting_1d2v.jl;
  (b) Hp3 against a direct numpy evaluation of its point-wise formulas;
  (c) the invariant of the scheme: the discrete Gauss law residual G^T M1 e - rho(particles) is constant
      to round-off over full Strang steps with all six field components alive.
CPU only."""
import numpy as np

from oracle import maxwell2d as m2
from oracle import oracle as orc
from oracle import splitting2d3v as s2

from .helpers import landau_state

L1 = 4 * np.pi
NX, NY, DEG = 16, 12, 3


def make_2d(n, seed=7, ly=3.0, xmin=0.0, ymin=0.0):
    rng = np.random.default_rng(seed)
    st1 = landau_state(n, L1, seed=seed)
    pg = s2.ParticleGroup23(n)
    pg.array[0] = xmin + st1[0]
    pg.array[1] = ymin + rng.uniform(0, ly, n)
    pg.array[2], pg.array[3] = st1[1], st1[2]
    pg.array[4] = 0.0
    pg.array[5] = st1[3]
    mesh = orc.TwoDGrid(xmin, xmin + L1, NX, ymin, ymin + ly, NY)
    mx = m2.TwoDMaxwell(mesh, DEG)
    return st1, pg, mx


def make_1d(st1):
    mesh = orc.OneDGrid(0.0, L1, NX)
    n = st1.shape[1]
    pg = orc.ParticleGroup(1, 2, n)
    pg.array[:, :] = st1
    ks0 = orc.ParticleMeshCoupling1D(mesh, n, DEG, "galerkin")
    ks1 = orc.ParticleMeshCoupling1D(mesh, n, DEG - 1, "galerkin")
    mx = orc.Maxwell1DFEM(mesh, DEG)
    return pg, ks0, ks1, mx


def lift(v1d):
    """x2-independent 2D dofs from 1D dofs (x fastest)"""
    return np.tile(v1d, NY)


def test_operators_reduce_to_1d2v():
    n = 3000
    st1, pg2, mx2 = make_2d(n)
    pg1, ks0, ks1, mx1 = make_1d(st1)
    rng = np.random.default_rng(11)
    e1, e2, b = rng.normal(size=NX), rng.normal(size=NX), rng.normal(size=NX)
    h1 = orc.HamiltonianSplitting(1, 2, mx1, ks0, ks1, pg1, [e1.copy(), e2.copy()], b.copy())
    e2d = [lift(e1), lift(e2), np.zeros(NX * NY)]
    b2d = [np.zeros(NX * NY), np.zeros(NX * NY), lift(b)]
    h2 = s2.HamiltonianSplitting2D3V(mx2, pg2, e2d, b2d)
    tol = 2e-13

    def same_particles():
        a2, a1 = pg2.array, pg1.array
        dx = np.abs(a2[0] - a1[0])
        dx = np.minimum(dx, np.abs(dx - L1))
        assert np.max(dx) < tol * L1
        assert np.max(np.abs(a2[2] - a1[1])) < tol * max(1.0, np.max(np.abs(a1[1])))
        assert np.max(np.abs(a2[3] - a1[2])) < tol * max(1.0, np.max(np.abs(a1[2])))
        assert np.all(a2[4] == 0.0)   # v3 stays zero: B1 = B2 = E3 = 0

    h1.operatorHE(0.05)
    h2.operatorHE(0.05)
    same_particles()
    # keep the (by now x2-dependent) 2D fields in sync with the 1D ones: only the particle parts are compared
    def resync():
        h2.e_dofs[0][:], h2.e_dofs[1][:], h2.e_dofs[2][:] = lift(h1.e_dofs[0]), lift(h1.e_dofs[1]), 0.0
        h2.b_dofs[0][:], h2.b_dofs[1][:], h2.b_dofs[2][:] = 0.0, 0.0, lift(h1.b_dofs)

    resync()
    # The 1D reference locates x_new with trunc: a particle leaving through xmin is integrated
    # with the polynomial extension of cell 0.  The 2D operators use floor, so the comparison keeps such
    # particles inside the box.
    leaving = pg1.array[0] + 0.05 * pg1.array[1] < 0.0
    assert leaving.any()
    pg1.array[1, leaving] *= -1.0
    pg2.array[2, leaving] *= -1.0
    h1.operatorHp1(0.05)
    h2.operatorHp1(0.05)
    same_particles()
    j1_sum = h2.j_dofs[0].reshape(NY, NX).sum(axis=0)
    assert np.max(np.abs(j1_sum - h1.j_dofs[0])) < tol * np.max(np.abs(h1.j_dofs[0]))
    resync()
    x2_before = pg2.array[1].copy()
    h1.operatorHp2(0.05)
    h2.operatorHp2(0.05)
    same_particles()
    j2_sum = h2.j_dofs[1].reshape(NY, NX).sum(axis=0)
    assert np.max(np.abs(j2_sum - h1.j_dofs[1])) < tol * np.max(np.abs(h1.j_dofs[1]))   # 1D j2 already holds dt*j2
    moved = np.abs(pg2.array[1] - x2_before)
    assert np.max(np.minimum(moved, np.abs(moved - 3.0))) > 0.0


def test_hp3_pointwise():
    n = 500
    _, pg, mx = make_2d(n, seed=3)
    rng = np.random.default_rng(5)
    pg.array[4] = rng.normal(size=n)
    nd = NX * NY
    e = [np.zeros(nd) for _ in range(3)]
    b = [rng.normal(size=nd), rng.normal(size=nd), np.zeros(nd)]
    h = s2.HamiltonianSplitting2D3V(mx, pg, e, b)
    before = pg.array.copy()

    def eval_field(f, dgx, dgy, x, y):
        xi, yi = (x - 0.0) / mx.dx, (y - 0.0) / mx.dy
        cx, cy = int(np.floor(xi)), int(np.floor(yi))
        bx, by = orc.bsplines_eval_basis(dgx, xi - cx), orc.bsplines_eval_basis(dgy, yi - cy)
        F = f.reshape(NY, NX)
        return sum(F[(cy - dgy + q) % NY, (cx - dgx + a) % NX] * bx[a] * by[q] for a in range(dgx + 1) for q in range(dgy + 1))

    dt = 0.1
    h.operatorHp3(dt)
    for i in range(0, n, 37):
        x, y, v1, v2, v3 = before[0, i], before[1, i], before[2, i], before[3, i], before[4, i]
        B1 = eval_field(b[0], DEG, DEG - 1, x, y)
        B2 = eval_field(b[1], DEG - 1, DEG, x, y)
        assert abs(pg.array[2, i] - (v1 - dt * v3 * B2)) < 1e-13
        assert abs(pg.array[3, i] - (v2 + dt * v3 * B1)) < 1e-13
    # total j3 = dt * sum_p q w v3 (partition of unity)
    tot = dt * pg.charge * pg.common_weight * np.sum(before[5] * before[4])
    assert abs(h.j_dofs[2].sum() - tot) < 1e-12 * max(1.0, abs(tot))


def test_gauss_law_is_conserved_over_strang_steps():
    n = 4000
    _, pg, mx = make_2d(n, seed=21, xmin=-1.0, ymin=0.5)
    rng = np.random.default_rng(9)
    pg.array[4] = 0.3 * rng.normal(size=n)
    nd = NX * NY
    e = [0.1 * rng.normal(size=nd) for _ in range(3)]
    b = [0.1 * rng.normal(size=nd) for _ in range(3)]
    h = s2.HamiltonianSplitting2D3V(mx, pg, e, b)
    r0 = h.gauss_residual()
    k0 = sum(h.energies())
    h.strang_splitting(0.05, 4)
    r1 = h.gauss_residual()
    scale = np.max(np.abs(h.charge_density()))
    assert np.max(np.abs(r1 - r0)) < 1e-12 * max(scale, np.max(np.abs(r0)))
    # the splitting is symplectic, not energy conserving: the total energy only drifts at O(dt^2)
    k1 = sum(h.energies())
    assert abs(k1 - k0) < 2e-2 * abs(k0)
    # positions stay inside the periodic box
    assert np.all((pg.array[0] >= -1.0) & (pg.array[0] < -1.0 + L1))
    assert np.all((pg.array[1] >= 0.5) & (pg.array[1] < 3.5))


def _fresh(n, seed, amp=0.05):
    _, pg, mx = make_2d(n, seed=seed)
    rng = np.random.default_rng(seed + 100)
    pg.array[4] = 0.3 * rng.normal(size=n)
    nd = NX * NY
    # smooth fields (a few low modes) so that the O(dt^2) regime is reached at dt ~ 0.1
    ix, iy = np.arange(nd) % NX, np.arange(nd) // NX
    mode = lambda a, b, ph: np.cos(2 * np.pi * (a * ix / NX + b * iy / NY) + ph)
    e = [amp * mode(1, 0, 0.3), amp * mode(0, 1, 1.1), amp * mode(1, 1, 2.0)]
    b = [amp * mode(1, 1, 0.7), amp * mode(1, 0, 1.9), amp * mode(0, 1, 0.2)]
    return s2.HamiltonianSplitting2D3V(mx, pg, e, b)


def test_strang_step_is_time_reversible():
    """(d) every sub-flow is solved exactly and the Strang composition is symmetric, so S(-dt) S(dt) = id to round-off:
    particles, all six field components and the currents' effect come back"""
    h = _fresh(2000, seed=41)
    x0 = h.particle_group.array.copy()
    e0, b0 = [v.copy() for v in h.e_dofs], [v.copy() for v in h.b_dofs]
    h.strang_splitting(0.05, 3)
    assert np.max(np.abs(h.particle_group.array[2:5] - x0[2:5])) > 1e-4      # something happened
    h.strang_splitting(-0.05, 3)
    a = h.particle_group.array
    for d, L in ((0, L1), (1, 3.0)):
        dx = np.abs(a[d] - x0[d])
        assert np.max(np.minimum(dx, np.abs(dx - L))) < 1e-12
    assert np.max(np.abs(a[2:5] - x0[2:5])) < 1e-12
    for c in range(3):
        assert np.max(np.abs(h.e_dofs[c] - e0[c])) < 1e-12 and np.max(np.abs(h.b_dofs[c] - b0[c])) < 1e-12


def test_strang_splitting_is_second_order():
    """(e) global error against a dt/16 solution of the same semi-discrete system falls by ~4 when dt is halved"""
    T = 0.4

    def run(dt):
        h = _fresh(1500, seed=43)
        h.strang_splitting(dt, int(round(T / dt)))
        return np.concatenate([h.particle_group.array[2:5].ravel()] + [v for v in h.e_dofs] + [v for v in h.b_dofs])

    ref = run(T / 64)
    err = [np.max(np.abs(run(dt) - ref)) for dt in (T / 4, T / 8)]
    assert 3.0 < err[0] / err[1] < 5.5, err
