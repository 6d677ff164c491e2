"""The reference's own golden tests, driven through the C ABI on the GPU.

Each test mirrors one file of /root/reference/test and compares against the literals in
tests/golden/*.json with the reference's tolerance (written next to each assert)."""
import math

import numpy as np
import pytest

from .conftest import golden
from .test_oracle_golden import T1D, T2D, TBO, THS, _rho_ref_1d

pytestmark = pytest.mark.gpu


def _setup_1d(gp):
    n_cells, n_particles, degree = 10, 4, 3
    mesh = gp.OneDGrid(0.0, 2.0, n_cells)
    x_vec = golden(T1D, 11)
    v_vec = golden(T1D, 12).reshape(2, 4).T
    pg = gp.ParticleGroup(1, 2, n_particles)
    for i in range(n_particles):  # set_x!/set_v!/set_weights! (test :18-22)
        pg.set_x(i, x_vec[i])
        pg.set_v(i, v_vec[i])
        pg.set_weights(i, 1.0)
    kernel = gp.ParticleMeshCoupling1D(mesh, n_particles, degree, "collocation")
    return mesh, pg, kernel, n_cells, n_particles


def test_pmc1d_add_charge(gp):
    mesh, pg, kernel, n_cells, n_particles = _setup_1d(gp)
    rho = np.zeros(n_cells)
    for i in range(n_particles):  # per-particle calls, like the reference loop (:52-55)
        kernel.add_charge(rho, pg.get_x(i)[0], pg.get_charge(i))
    ref = _rho_ref_1d(n_cells, n_particles, mesh.xmax)
    assert np.max(np.abs(rho - ref)) < 1e-15  # test_particle_mesh_coupling_spline_1d.jl:64
    rho2 = np.zeros(n_cells)
    kernel.add_charge_pg(rho2, pg)  # batched, device-resident group
    assert np.max(np.abs(rho2 - ref)) < 1e-15


def test_pmc1d_add_current_update_v(gp):
    mesh, pg, kernel, n_cells, n_particles = _setup_1d(gp)
    j, b = np.zeros(n_cells), np.zeros(n_cells)
    for i in range(n_particles):
        xi = pg.get_x(i)[0]
        vi = pg.get_v(i)
        x_new = xi + vi[0] / 10.0
        vi[1] = kernel.add_current_update_v(j, xi, x_new, pg.get_charge(i), 1.0, b, vi[1])
    ref = golden(T1D, 82) + golden(T1D, 96) + golden(T1D, 110)
    assert np.max(np.abs(j - ref)) < 1e-15  # :123
    # batched call: all four particles in one launch
    jb = np.zeros(n_cells)
    arr = pg.to_host()
    kernel.add_current_update_v(jb, arr[0], arr[0] + arr[1] / 10.0, np.full(4, pg.get_charge(0)), 1.0, b, arr[2])
    assert np.max(np.abs(jb - ref)) < 1e-15


def test_pmc1d_evaluate(gp):
    mesh, pg, kernel, n_cells, n_particles = _setup_1d(gp)
    rho = np.zeros(n_cells)
    kernel.add_charge_pg(rho, pg)
    vals = np.array([kernel.evaluate(pg.get_x(i)[0], rho) for i in range(n_particles)])
    ref = golden(T1D, 137) / mesh.xmax
    assert np.max(np.abs(vals - ref)) < 1e-15  # :143
    assert np.max(np.abs(kernel.evaluate_pg(pg, rho) - ref)) < 1e-15
    # the `_pp` entry points (pmc1d.jl:106-122,242-250) are held to the same goldens by the reference (:57-65,137-144)
    pp = kernel.b_to_pp(rho)
    vals_pp = np.array([gp.evaluate_pp(kernel, pg.get_x(i)[0], pp) for i in range(n_particles)])
    assert np.max(np.abs(vals_pp - ref)) < 1e-15
    rho_pp = np.zeros(n_cells)
    for i in range(n_particles):
        gp.add_charge_pp(rho_pp, kernel, pg.get_x(i)[0], pg.get_charge(i))
    assert np.max(np.abs(rho_pp - rho)) < 1e-15


def test_pmc2d(gp):
    n_cells, n_particles, degree = 10, 4, 3
    grid = gp.TwoDGrid(0.0, 2.0, n_cells, 0.0, 1.0, n_cells)
    volume = 2.0
    x_vec = golden(T2D, 12).reshape(2, 4)
    v_vec = golden(T2D, 13).reshape(2, 4)
    pg = gp.ParticleGroup(2, 2, n_particles, charge=1.0, mass=1.0, n_weights=1)
    arr = np.zeros((5, n_particles))
    arr[0:2], arr[2:4], arr[4] = x_vec, v_vec, 1.0 / n_particles
    pg.upload(arr)
    kernel = gp.ParticleMeshCoupling2D(pg, grid, degree, "collocation")
    idx_ref = np.stack([golden(T2D, 30), golden(T2D, 31)]).astype(int)
    vg = np.zeros((4, 2, 4))
    vg[:, 0, 0] = golden(T2D, 34)
    vg[:, 0, 2] = vg[:, 0, 0]
    vg[:, 0, 3] = vg[:, 0, 0]
    vg[:, 0, 1] = golden(T2D, 42)
    vg[0, 1, :], vg[1, 1, :], vg[2, 1, :], vg[3, 1, :] = 0.0, 1 / 6, 2 / 3, 1 / 6
    rho = np.zeros(100)
    for i in range(n_particles):
        kernel.add_charge(rho, x_vec[0, i], x_vec[1, i], 1.0 / n_particles)
    ref = np.zeros(100)
    ref[7:10] = vg[0:3, 0, 0]
    ref[0] = vg[3, 0, 0]
    ref[0:4] += vg[:, 0, 1] + vg[:, 0, 2]
    ref[4:8] += vg[:, 0, 3]
    ref[70:80] = ref[0:10] / 6.0
    ref[80:90] = ref[0:10] * 2.0 / 3.0
    ref[90:100] = ref[0:10] / 6.0
    ref[0:10] = 0.0
    ref *= n_cells**2 / volume / n_particles
    np.testing.assert_allclose(rho, ref, rtol=1e-14, atol=1e-14)  # `≈` test_particle_mesh_coupling_spline_2d.jl:79
    vals_ref = np.zeros(4)
    for p in range(n_particles):
        for i in range(4):
            i1 = (idx_ref[0, p] + i - 1) % n_cells
            for jj in range(4):
                i2 = (idx_ref[1, p] + jj - 1) % n_cells
                vals_ref[p] += vg[i, 0, p] * vg[jj, 1, p] * ref[i1 + i2 * n_cells]
    vals = kernel.evaluate(x_vec[0], x_vec[1], rho)
    np.testing.assert_allclose(vals, vals_ref, rtol=1e-14)  # :117
    np.testing.assert_allclose(kernel.evaluate_pg(pg, rho), vals_ref, rtol=1e-14)
    v1, v2 = kernel.evaluate_multiple((x_vec[0], x_vec[1]), [rho, 2 * rho])
    np.testing.assert_allclose(v1, vals_ref, rtol=1e-14)
    np.testing.assert_allclose(v2, 2 * vals_ref, rtol=1e-14)


def _setup_hs(gp, num_cells, info):
    n_particles = info.shape[1]
    mesh = gp.OneDGrid(0.0, 4 * math.pi, num_cells)
    pg = gp.ParticleGroup(1, 2, n_particles, common_weight=1.0)
    pg.upload(info)
    ks1 = gp.ParticleMeshCoupling1D(mesh, n_particles, 2, "galerkin")
    ks0 = gp.ParticleMeshCoupling1D(mesh, n_particles, 3, "galerkin")
    maxwell = gp.Maxwell1DFEM(mesh, 3)
    e1, e2, b = np.ones(num_cells), np.ones(num_cells), np.ones(num_cells)
    rho = np.zeros(num_cells)
    ks0.add_charge_pg(rho, pg)
    maxwell.compute_e_from_rho(e1, rho)
    return mesh, pg, ks0, ks1, maxwell, e1, e2, b


@pytest.mark.parametrize("resident", [False, True])
def test_hamiltonian_splitting(gp, resident):
    info0 = golden(THS, 23).reshape(4, 2)
    mesh, pg, ks0, ks1, maxwell, e1, e2, b = _setup_hs(gp, 10, info0)
    h = gp.HamiltonianSplitting(1, 2, maxwell, ks0, ks1, pg, [e1, e2], b, resident=resident)
    dt = 0.1

    def check(line):
        ref = golden(THS, line).reshape(4, 2)
        for i in range(2):
            np.testing.assert_allclose(pg.get_x(i)[0], ref[0, i], rtol=1e-14)  # `≈` (rtol sqrt(eps)) in the reference
            np.testing.assert_allclose(pg.get_v(i), ref[1:3, i], rtol=1e-13)
            assert abs(pg.get_charge(i) - ref[3, i]) <= 1e-15

    gp.operatorHp1(h, dt)
    check(93)
    gp.operatorHp2(h, dt)
    check(128)
    gp.operatorHE(h, dt)
    check(163)
    gp.operatorHB(h, dt)
    check(198)
    if resident:
        h.sync_fields()
    b_ref = golden(THS, 228)
    e_ref = golden(THS, 242).reshape(2, 10)
    assert np.max(np.abs(b - b_ref)) < 1e-14  # test_hamiltonian_splitting.jl:268-270 (atol 1e-14)
    assert np.max(np.abs(e1 - e_ref[0])) < 1e-14
    assert np.max(np.abs(e2 - e_ref[1])) < 1e-14


def test_boris(gp):
    info0 = golden(TBO, 18).reshape(2, 4).T
    mesh, pg, ks0, ks1, maxwell, e1, e2, b = _setup_hs(gp, 16, info0)
    prop = gp.HamiltonianSplittingBoris(maxwell, ks0, ks1, pg, [e1, e2], b)
    dt = 0.1
    gp.staggering(prop, 0.5 * dt)
    gp.strang_splitting(prop, dt, 1)
    ref = golden(TBO, 84).reshape(2, 4).T
    for i in range(2):
        np.testing.assert_allclose(pg.get_x(i)[0], ref[0, i], rtol=1e-14)
        np.testing.assert_allclose(pg.get_v(i), ref[1:3, i], rtol=1e-13)
        assert abs(pg.get_charge(i) - ref[3, i]) <= 1e-15
    assert np.max(np.abs(b - golden(TBO, 113))) < 1e-15  # reference: `≈ 0.0` (:132)
    e_ref = golden(TBO, 135).reshape(2, 16)
    # :173-174 assert 1e-15 with FFTW; values are O(4.8) (1 ulp = 8.9e-16): allow 2 ulp for the
    # convolution form of the circulant solve
    assert np.max(np.abs(e1 - e_ref[0])) < 2e-15
    assert np.max(np.abs(e2 - e_ref[1])) < 1e-15


@pytest.mark.parametrize("n,deg,L", [(32, 3, 4 * math.pi), (256, 3, 2 * math.pi), (20, 2, 1.7), (8, 1, 3.0)])
def test_maxwell1d_tables_match_oracle(gp, orc, n, deg, L):
    """Maxwell1DFEM constructor tables (maxwell_1d_fem.jl:60-153: eig_mass0/1, eig_weak_ampere, eig_weak_poisson in the
    FFTW half-complex layout) read back through gempic_maxwell1d_get_table, against the CPU restatement"""
    mg = gp.Maxwell1DFEM(gp.OneDGrid(0.25, 0.25 + L, n), deg)
    mo = orc.Maxwell1DFEM(orc.OneDGrid(0.25, 0.25 + L, n), deg)
    for name in ("eig_mass0", "eig_mass1", "eig_weak_ampere", "eig_weak_poisson"):
        a, b = getattr(mg, name), getattr(mo, name)
        assert a.shape == b.shape == (n,)
        assert np.max(np.abs(a - b)) <= 4e-16 * max(1.0, np.max(np.abs(b))), name


def test_maxwell1d_analytic(gp):
    # test_maxwell_1d_fem.jl (Poisson, Ampere, 10 leap-frog steps) through the C ABI
    mode, n, deg = 2, 256, 3
    Lx = 2 * math.pi
    dx = Lx / n
    mx = gp.Maxwell1DFEM(gp.OneDGrid(0.0, Lx, n), deg)
    cos_k = lambda x: math.cos(mode * 2 * math.pi * x / Lx)
    xi = np.arange(n) * dx

    def bsp(degree):
        return {2: [0.5, 0.5, 0.0], 3: [1 / 6, 2 / 3, 1 / 6, 0.0]}[degree]

    def spline_curve(degree, coef):
        out = np.zeros(n)
        for j in range(1, degree + 1):
            out += bsp(degree)[j - 1] * np.roll(coef, j)
        return out

    rho, ex = np.zeros(n), np.zeros(n)
    gp.compute_rhs_from_function(rho, mx, cos_k, deg)
    gp.compute_e_from_rho(ex, mx, rho)
    assert np.max(np.abs(spline_curve(deg - 1, ex) - np.sin(mode * xi) / (2.0 * mode * math.pi / Lx))) < 1e-6
    dt = 0.5 * dx
    gp.compute_rhs_from_function(rho, mx, cos_k, deg - 1)
    ex[:] = 0.0
    gp.compute_e_from_j(ex, mx, dt * rho, 1)
    assert np.max(np.abs(spline_curve(deg - 1, ex) + np.cos(mode * xi) * dt)) < 1e-6
    assert abs(gp.l2norm_squared(mx, ex, deg - 1) - dt * dt * math.pi) < 1e-8
    ey, bz = np.zeros(n), np.zeros(n)
    gp.l2projection(bz, mx, cos_k, deg - 1)
    time = 0.0
    for _ in range(10):
        gp.compute_b_from_e(bz, mx, 0.5 * dt, ey)
        gp.compute_e_from_b(ey, mx, dt, bz)
        gp.compute_b_from_e(bz, mx, 0.5 * dt, ey)
        time += dt
        assert np.linalg.norm(spline_curve(deg, ey) - np.sin(mode * xi) * math.sin(mode * time)) < 1e-2
        assert np.linalg.norm(spline_curve(deg - 1, bz) - np.cos(mode * xi) * math.cos(mode * time)) < 1e-2


def test_error_behaviour(gp):
    mesh = gp.OneDGrid(0.0, 1.0, 8)
    with pytest.raises(gp.ArgumentError):  # pmc1d.jl:61
        gp.ParticleMeshCoupling1D(mesh, 4, 3, "nearest")
    with pytest.raises(gp.ArgumentError):  # maxwell_1d_fem.jl:91
        gp.Maxwell1DFEM(mesh, 4)
    mx = gp.Maxwell1DFEM(mesh, 3)
    with pytest.raises(gp.ArgumentError):  # maxwell_1d_fem.jl:283
        mx.compute_e_from_j(np.zeros(8), np.zeros(8), 3)
    pg = gp.ParticleGroup(1, 2, 4)
    ks0 = gp.ParticleMeshCoupling1D(mesh, 4, 3, "galerkin")
    ks1 = gp.ParticleMeshCoupling1D(gp.OneDGrid(0.0, 1.0, 16), 4, 2, "galerkin")
    with pytest.raises(gp.AssertionFailed):  # hamiltonian_splitting.jl:49
        gp.HamiltonianSplitting(1, 2, mx, ks0, ks1, pg, [np.zeros(8), np.zeros(8)], np.zeros(8))
    pg11 = gp.ParticleGroup(1, 1, 4)
    ks1b = gp.ParticleMeshCoupling1D(mesh, 4, 2, "galerkin")
    with pytest.raises(gp.AssertionFailed):  # hamiltonian_splitting.jl:47
        gp.HamiltonianSplitting(1, 2, mx, ks0, ks1b, pg11, [np.zeros(8), np.zeros(8)], np.zeros(8))
