"""Worker of tests/test_gpu_multi.py::test_one_process_drives_all_devices: runs the same script on
`n` devices of ONE process (gempic_init_devices) -- or on a single device (n = 1, gempic_init) -- and stores what a
caller can observe.  The test compares the two files."""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import __graft_entry__ as ge  # noqa: E402


def main(n_dev, out, force_md=False):
    gp = ge.load_package()
    if n_dev == 1 and not force_md:
        gp.init(0)
    else:
        gp.init_devices(n_dev)      # also with ONE device: every call then goes through the dispatcher and a worker thread
    assert gp.device_count() == n_dev
    res = {}
    nx, L = 32, 4 * math.pi
    n = 200_003
    rng = np.random.default_rng(5)
    # ---- device samplers address the global index range
    mesh = gp.OneDGrid(0.0, L, nx)
    pg = gp.ParticleGroup(1, 2, n)
    gp.sample(pg, 0.5, 0.5, 1.0, mesh)                      # Sobol + Newton Landau load
    res["landau_load"] = pg.to_host()
    ps = gp.ParticleSampler(1, 2, "sobol", True, 80_000)
    pgs = gp.ParticleGroup(1, 2, ps.n_particles)
    gp.sample(pgs, ps, gp.CosSumGaussian(1, 2, [[0.5]], [0.1], [[1.0, 0.5]], [[0.0, 0.0]]), mesh)
    res["sym_load"] = pgs.to_host()
    # ---- 1d2v: fused strang_splitting!, the diagnostics loop, j_dofs, evaluate
    ks0 = gp.ParticleMeshCoupling1D(mesh, n, 3, "galerkin")
    ks1 = gp.ParticleMeshCoupling1D(mesh, n, 2, "galerkin")
    mx = gp.Maxwell1DFEM(mesh, 3)
    e1, e2, rho, ep = np.zeros(nx), np.zeros(nx), np.zeros(nx), np.zeros(nx)
    b = 1e-2 * np.cos(2 * math.pi * (np.arange(nx) + 0.5) / nx)
    gp.solve_poisson(e1, pg, ks0, mx, rho)
    res["rho0"], res["e1_0"] = rho.copy(), e1.copy()
    h = gp.HamiltonianSplitting(1, 2, mx, ks0, ks1, pg, [e1, e2], b)
    th = gp.TimeHistoryDiagnostics(pg, mx, ks0, ks1)
    rows = []
    for j in range(3):
        e_n = [e1.copy(), e2.copy()]
        h.strang_splitting(0.05, 1)
        gp.solve_poisson(ep, pg, ks0, mx, rho)
        rows.append(gp.write_step(th, (j + 1) * 0.05, 3, [e1, e2], b, e_n, ep))
    h.strang_splitting(0.05, 2)
    res["diag"] = np.array(rows)
    res["hs_fields"] = np.stack([e1, e2, b] + [v.copy() for v in h.j_dofs])
    res["hs_particles"] = pg.to_host()
    res["evaluate_pg"] = ks0.evaluate_pg(pg, e2)
    rho_pg = np.zeros(nx)
    ks0.add_charge_pg(rho_pg, pg)
    res["add_charge_pg"] = rho_pg
    h.operatorHp1(0.03)                                      # a single operator (un-fused pass with the rho deposit)
    res["after_hp1"] = np.stack([e1, e2, b] + [v.copy() for v in h.j_dofs])
    # ---- Boris on an uploaded particle array
    st = np.stack([rng.uniform(0, L, n), rng.normal(size=n), rng.normal(size=n), np.full(n, L)])
    pgb = gp.ParticleGroup(1, 2, n)
    pgb.upload(st)
    f1, f2, fb = np.zeros(nx), 1e-3 * np.sin(2 * math.pi * np.arange(nx) / nx), b.copy()
    gp.solve_poisson(f1, pgb, ks0, mx, rho)
    hb = gp.HamiltonianSplittingBoris(mx, ks0, ks1, pgb, [f1, f2], fb)
    hb.staggering(0.05)
    hb.strang_splitting(0.05, 3)
    res["boris_fields"] = np.stack([f1, f2, fb, hb.e_dofs_mid[0], hb.e_dofs_mid[1], hb.b_dofs_mid] + hb.j_dofs)
    res["boris_particles"] = pgb.to_host()
    # ---- 2d3v with the riding sort and the sorted fast path
    n2, nx2 = 120_000, 16
    st2 = np.empty((6, n2))
    st2[0], st2[1] = rng.uniform(0, L, n2), rng.uniform(0, L, n2)
    for k in range(3):
        st2[2 + k] = rng.normal(size=n2)
    st2[5] = L * L * (1.0 + 0.1 * np.arange(n2) / n2)
    pg2 = gp.ParticleGroup(2, 3, n2, charge=-1.0)
    pg2.upload(st2)
    m2 = gp.TwoDMaxwell(gp.TwoDGrid(0.0, L, nx2, 0.0, L, nx2), 3)
    nd = nx2 * nx2
    ee = [0.05 * rng.normal(size=nd) for _ in range(3)]
    bb = [0.05 * rng.normal(size=nd) for _ in range(3)]
    h2 = gp.HamiltonianSplitting2D3V(m2, pg2, ee, bb)
    h2.strang_splitting(0.05, 2)
    h2.strang_splitting(0.05, 1)
    res["hs2d_fields"] = np.stack(ee + bb + h2.j_dofs + [h2.charge_density()])
    res["hs2d_moments"] = h2.moments()
    a = pg2.to_host()
    res["hs2d_particles"] = a[:, np.argsort(a[5])]
    # ---- error behaviour crosses the dispatcher
    try:
        gp.ParticleMeshCoupling1D(mesh, n, 7, "galerkin")
        res["error"] = np.array([0.0])
    except gp.ArgumentError:
        res["error"] = np.array([1.0])
    np.savez(out, **res)
    gp.finalize()


if __name__ == "__main__":
    main(int(sys.argv[1]), sys.argv[2], force_md="--md" in sys.argv[3:])
