"""Full runs and full sizes (BASELINE.json north_star: "the electric/magnetic field-energy and momentum diagnostics
track the reference over a full run"; parity at BASELINE's full sizes through size-independent properties).

  * config 1 (examples/strong_landau_damping_1d1v.jl: 1d1v, 32 cells, degree 3/2, 1e5 particles, dt 0.05): the
    PotentialEnergyE1 series of the CUDA path follows the oracle's over 300 steps;
  * 1d2v Weibel (test/test_vm_1d2v.jl set-up): all 11 write_step! columns follow the oracle's over 60 steps;
  * 1e8 particles on one GPU (config 2 size): charge and momentum bookkeeping, sortedness, fused == unfused."""
import math

import numpy as np
import pytest

from .helpers import Sim1D, landau_state, rel_err, weibel_state

pytestmark = pytest.mark.gpu

L_LANDAU = 4 * math.pi
L_WEIBEL = 2 * math.pi / 1.25


def test_config1_landau_1d1v_energy_series(orc, gp):
    n, steps, dt = 100_000, 300, 0.05
    state = landau_state(n, L_LANDAU, seed=1234, V=1)
    so, sg = Sim1D(orc, state, L_LANDAU, nx=32, V=1), Sim1D(gp, state, L_LANDAU, nx=32, V=1)
    so.init_fields(b_amp=0.0), sg.init_fields(b_amp=0.0)
    ho, hg = so.splitting(V=1), sg.splitting(V=1)
    pe_o, pe_g = [], []
    for _ in range(steps):
        ho.strang_splitting(dt, 1)
        hg.strang_splitting(dt, 1)
        pe_o.append(0.5 * so.mx.l2norm_squared(so.e1, 2))   # PotentialEnergyE1 (diagnostics.jl:226), degree - 1
        pe_g.append(0.5 * sg.mx.l2norm_squared(sg.e1, 2))
    pe_o, pe_g = np.array(pe_o), np.array(pe_g)
    # strong Landau damping: the field energy first decays by orders of magnitude, then grows again
    assert pe_o.min() < 0.05 * pe_o[0]
    assert np.max(np.abs(pe_g - pe_o) / pe_o) < 1e-8


def test_weibel_1d2v_diagnostics_series(orc, gp):
    n, steps, dt = 200_000, 60, 0.05
    state = weibel_state(n, L_WEIBEL, seed=3, sigma=(0.2, 0.005773502691896))
    so, sg = Sim1D(orc, state, L_WEIBEL, nx=32), Sim1D(gp, state, L_WEIBEL, nx=32)
    so.init_fields(b_amp=1e-3), sg.init_fields(b_amp=1e-3)
    ho, hg = so.splitting(), sg.splitting()
    th = gp.TimeHistoryDiagnostics(sg.pg, sg.mx, sg.ks0, sg.ks1)
    rows_o, rows_g = [], []
    epo, epg, rho_o, rho_g = np.zeros(32), np.zeros(32), np.zeros(32), np.zeros(32)
    for j in range(steps):
        e_n_o, e_n_g = [so.e1.copy(), so.e2.copy()], [sg.e1.copy(), sg.e2.copy()]
        ho.strang_splitting(dt, 1)
        hg.strang_splitting(dt, 1)
        orc.solve_poisson(epo, so.pg, so.ks0, so.mx, rho_o)
        gp.solve_poisson(epg, sg.pg, sg.ks0, sg.mx, rho_g)
        rows_o.append(orc.write_step(so.pg, so.mx, so.ks0, so.ks1, (j + 1) * dt, 3, [so.e1, so.e2], so.b, e_n_o, epo))
        rows_g.append(gp.write_step(th, (j + 1) * dt, 3, [sg.e1, sg.e2], sg.b, e_n_g, epg))
    ro, rg = np.array(rows_o), np.array(rows_g)
    ke = np.max(np.abs(ro[:, 1]))
    for k, name in enumerate(gp.DIAG_COLUMNS):
        if name == "Time":
            assert np.array_equal(ro[:, k], rg[:, k])
            continue
        # energies are compared relatively; momenta, transfer terms and the Poisson error (sums that cancel to
        # ~0) relative to the kinetic energy scale
        scale = np.maximum(np.abs(ro[:, k]), 1e-9 * ke)
        assert np.max(np.abs(rg[:, k] - ro[:, k]) / scale) < 1e-6, name
    # the run is not trivial: field energies move by more than rounding
    col = list(gp.DIAG_COLUMNS).index("PotentialEnergyB3")
    assert abs(ro[-1, col] - ro[0, col]) > 1e-3 * ro[0, col]


def test_full_size_1e8_properties(gp):
    """BASELINE config 2 size on one GPU: 1e8 particles, 32 cells"""
    n, nx, L = 100_000_000, 32, L_WEIBEL
    mesh = gp.OneDGrid(0.0, L, nx)
    ks0 = gp.ParticleMeshCoupling1D(mesh, n, 3, "galerkin")
    ks1 = gp.ParticleMeshCoupling1D(mesh, n, 2, "galerkin")
    mx = gp.Maxwell1DFEM(mesh, 3)

    def run(fuse):
        pg = gp.ParticleGroup(1, 2, n)
        pg.sample("uniform", 0.0, L, sigma=(0.2, 0.005773502691896), seed=1234)
        e1, e2, rho = np.zeros(nx), np.zeros(nx), np.zeros(nx)
        b = 1e-4 * np.cos(2 * math.pi * (np.arange(nx) + 0.5) / nx)
        gp.solve_poisson(e1, pg, ks0, mx, rho)
        assert abs(rho.sum() - L) < 1e-9 * L                      # total charge: sum_p q w / N = L
        h = gp.HamiltonianSplitting(1, 2, mx, ks0, ks1, pg, [e1, e2], b, resident=True)
        h.set_fusion(fuse)
        h.strang_splitting(0.05, 3)
        h.sync_fields()
        rho2, ep = np.zeros(nx), np.zeros(nx)
        gp.solve_poisson(ep, pg, ks0, mx, rho2)
        assert abs(rho2.sum() - L) < 1e-9 * L                     # ... and it is still there after 3 steps
        # Gauss law: E1 stays close to the Poisson field of rho.  Not to round-off: the reference locates x_new with
        # trunc (reproduced), so the ~0.1 % of particles leaving through xmin per step deposit a
        # slightly wrong current; on this noise-level field (uniform load) that is a few per cent.
        assert rel_err(e1, ep) < 0.2
        return pg, e1.copy(), e2.copy(), b.copy()

    pg_f, e1f, e2f, bf = run(True)
    pg_u, e1u, e2u, bu = run(False)
    assert rel_err(e1f, e1u) < 1e-10 and rel_err(bf, bu) < 1e-10
    assert np.max(np.abs(e2f - e2u)) < 1e-10 * max(np.max(np.abs(e2u)), np.max(np.abs(e1u)))
    del pg_u
    # periodic cell sort at full size: sorted, and a second sort changes nothing (idempotence)
    pg_f.sort(ks0)
    x = np.empty(n)
    row = gp.load()  # noqa: F841  (library handle kept alive)
    x[:] = pg_f.to_host()[0]
    cells = np.floor(x / (L / nx)).astype(np.int32)
    assert np.all(np.diff(cells) >= 0)
    chk = float(np.sum(x[::1000]))
    pg_f.sort(ks0)
    assert float(np.sum(pg_f.to_host()[0][::1000])) == chk
