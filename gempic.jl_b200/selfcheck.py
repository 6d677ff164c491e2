"""Self-check of the sharded (multi-GPU) path: sharded vs un-sharded runs of the three integrators on the same seeded
particle sets, product against product (no CPU checker involved).  Used by tests/dist_worker.py (NCCL parity test) and by
bench.py, which runs it before the timed region at N > 1 and reports the result in its JSON line
(`sharded_parity`), because the driver's test box has a single GPU."""
from __future__ import annotations

import ctypes as C
import math
import sys

import numpy as np

from ._lib import check


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def landau_state(n, L, seed=1234, alpha=0.5, k=0.5, sigma=(1.0, 1.0)):
    """seeded Landau-like load in the reference record layout (rows x, v1, v2, w)"""
    rng = np.random.default_rng(seed)
    u = rng.uniform(0, 1, n)
    x = u * L
    for _ in range(30):  # inverse CDF of 1 + alpha cos(kx)
        x -= (x + alpha / k * np.sin(k * x) - u * L) / (1 + alpha * np.cos(k * x))
    return np.stack([np.mod(x, L), sigma[0] * rng.normal(size=n), sigma[1] * rng.normal(size=n), np.full(n, L)])


def sharded_parity(dc, n=400_003, n2d=120_001, steps=3):
    """Sharded (this communicator) vs un-sharded (rank 0 alone, after the collective part) runs of the three
    integrators.  Returns, on rank 0, {case: max relative error}; every rank must call it.

    1d2v  HamiltonianSplitting{1,2}, fused and one pass per operator (hamiltonian_splitting.jl:98-108); after the fused
          run ONLY rank 0 downloads its particles first -- pg_sync must stay rank-local (no lone collective)
    boris HamiltonianSplittingBoris: staggering! + strang_splitting! (hamiltonian_splitting_boris.jl:99-177), the fused
          one-pass step and the separate pushes
    2d3v  HamiltonianSplitting{2,3} on TwoDMaxwell, fused and un-fused, riding sort on: fields, currents, charge,
          moments and rank 0's particles (identified by their distinct weights, since the cell sort permutes them)"""
    gp = sys.modules[__package__]
    nx, L = 32, 4 * math.pi
    state = landau_state(n, L, seed=7)
    first, count = dc.shard(n)
    rank0 = dc.rank == 0

    def build(cols, n_global):
        mesh = gp.OneDGrid(0.0, L, nx)
        pg = gp.ParticleGroup(1, 2, cols.shape[1], common_weight=1.0 / n_global)
        pg.upload(np.ascontiguousarray(cols))
        ks0 = gp.ParticleMeshCoupling1D(mesh, n_global, 3, "galerkin")
        ks1 = gp.ParticleMeshCoupling1D(mesh, n_global, 2, "galerkin")
        mx = gp.Maxwell1DFEM(mesh, 3)
        e1, e2, rho = np.zeros(nx), np.zeros(nx), np.zeros(nx)
        b = 1e-2 * np.cos(2 * math.pi * (np.arange(nx) + 0.5) / nx)
        return pg, ks0, ks1, mx, e1, e2, b, rho

    def run_hs(cols, fuse, lone_download):
        pg, ks0, ks1, mx, e1, e2, b, rho = build(cols, n)
        gp.solve_poisson(e1, pg, ks0, mx, rho)          # all-reduced rho -> replicated e1
        h = gp.HamiltonianSplitting(1, 2, mx, ks0, ks1, pg, [e1, e2], b, fuse=bool(fuse))
        h.strang_splitting(0.05, steps)
        parts = None
        if lone_download and rank0:
            parts = pg.to_host()       # applies the deferred kick on this rank only: must not hang
        j = [v.copy() for v in h.j_dofs]          # collective (all ranks): rebuilds / sums j_dofs[2]
        if parts is None:
            parts = pg.to_host()
        return [e1.copy(), e2.copy(), b.copy(), j[0], j[1]], parts

    def run_boris(cols, fused):
        pg, ks0, ks1, mx, e1, e2, b, rho = build(cols, n)
        gp.solve_poisson(e1, pg, ks0, mx, rho)
        e2[:] = 1e-3 * np.sin(2 * math.pi * np.arange(nx) / nx)
        h = gp.HamiltonianSplittingBoris(mx, ks0, ks1, pg, [e1, e2], b)
        h.staggering(0.05)
        if fused:
            h.strang_splitting(0.05, steps)
            out = [e1.copy(), e2.copy(), b.copy()]
        else:   # the separate pushes (src/hamiltonian_splitting_boris.jl:189-288), deposits all-reduced
            h.push_v_epart(0.025)
            h.push_v_bpart(0.05)
            h.push_v_epart(0.025)
            h.push_x_accumulate_j(0.05)
            out = [e1.copy(), e2.copy(), b.copy()]
        jj = h.j_dofs
        return out + [jj[0], jj[1], h.e_dofs_mid[0], h.e_dofs_mid[1], h.b_dofs_mid], pg.to_host()

    nx2, L2 = 16, 4 * math.pi
    nd = nx2 * nx2
    rng = np.random.default_rng(11)
    st2 = np.empty((6, n2d))
    st2[0], st2[1] = rng.uniform(0, L2, n2d), rng.uniform(0, L2, n2d)
    for k in range(3):
        st2[2 + k] = (1.0, 1.0, 0.5)[k] * rng.normal(size=n2d)
    st2[5] = L2 * L2 * (1.0 + 0.1 * np.arange(n2d) / n2d)   # distinct weights: the cell sort permutes, the weight identifies
    e0 = [0.05 * rng.normal(size=nd) for _ in range(3)]
    b0 = [0.05 * rng.normal(size=nd) for _ in range(3)]
    f2, c2 = dc.shard(n2d)

    def run_2d(cols, fuse):
        pg = gp.ParticleGroup(2, 3, cols.shape[1], charge=-1.0, common_weight=1.0 / n2d)
        pg.upload(np.ascontiguousarray(cols))
        mx = gp.TwoDMaxwell(gp.TwoDGrid(0.0, L2, nx2, 0.0, L2, nx2), 3)
        h = gp.HamiltonianSplitting2D3V(mx, pg, [v.copy() for v in e0], [v.copy() for v in b0])
        h.set_fusion(bool(fuse))
        h.strang_splitting(0.05, steps)
        out = [v.copy() for v in h.e_dofs + h.b_dofs] + h.j_dofs + [h.charge_density(), h.moments()]
        a = pg.to_host()
        return out, a[:, np.argsort(a[5])]

    cases = {"hs_fused": lambda c: run_hs(c, 1, True), "hs_unfused": lambda c: run_hs(c, 0, False),
             "boris_fused": lambda c: run_boris(c, True), "boris_pushes": lambda c: run_boris(c, False)}
    cases2 = {"2d3v_fused": lambda c: run_2d(c, 1), "2d3v_unfused": lambda c: run_2d(c, 0)}
    res = {}
    for name, fn in cases.items():
        res[name] = fn(state[:, first:first + count])
        for f in res[name][0]:   # replicas agree bitwise (the all-reduce result is identical on all ranks)
            ref = dc.broadcast_bytes(f.tobytes() if rank0 else None, f.nbytes, src=0)
            assert ref == f.tobytes(), f"{name}: replicated fields drifted between ranks"
    for name, fn in cases2.items():
        res[name] = fn(st2[:, f2:f2 + c2])
        for f in res[name][0]:
            ref = dc.broadcast_bytes(f.tobytes() if rank0 else None, f.nbytes, src=0)
            assert ref == f.tobytes(), f"{name}: replicated fields drifted between ranks"
    dc.barrier()
    # every rank suspends the library communicator; rank 0 then repeats the runs un-sharded on its own GPU
    check(gp.load().gempic_comm_suspend(C.c_int(1)))
    errs = {}
    if rank0:
        for name, fn in cases.items():
            fields, parts = fn(state)
            errs[name] = max([rel_err(a, b) if np.max(np.abs(b)) > 0 else float(np.max(np.abs(a)))
                              for a, b in zip(res[name][0], fields)] +
                             [rel_err(res[name][1][k], parts[k, first:first + count]) for k in range(0, 3)])
        for name, fn in cases2.items():
            fields, parts = fn(st2)
            e = [rel_err(a, b) for a, b in zip(res[name][0], fields)]
            # both runs cell-sort their particles; ordered by weight, rank 0's shard is the first c2 columns of either
            mine, sub = res[name][1], parts[:, :c2]
            assert np.array_equal(mine[5], sub[5])
            dx = np.abs(mine[:2] - sub[:2])
            e.append(float(np.max(np.minimum(dx, np.abs(dx - L2))) / L2))
            e.append(max(rel_err(mine[k], sub[k]) for k in range(2, 5)))
            errs[name] = max(e)
    dc.barrier()
    check(gp.load().gempic_comm_suspend(C.c_int(0)))
    return errs


