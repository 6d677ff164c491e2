"""Object lifetimes behind the C ABI: a garbage collector may finalise the host-side wrappers in any
order, and a script may simply end without gempic_finalize."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_destroy_in_any_order(gp):
    """the particle group, the smoothers and the solver are destroyed BEFORE the splitting that points to them: the
    splitting keeps working (the library retains what it points to) and goes last"""
    n, nx, L = 20_000, 32, 4 * np.pi
    rng = np.random.default_rng(3)
    st = np.stack([rng.uniform(0, L, n), rng.normal(size=n), rng.normal(size=n), np.full(n, L)])
    mesh = gp.OneDGrid(0.0, L, nx)
    pg = gp.ParticleGroup(1, 2, n)
    pg.upload(st)
    ks0, ks1 = gp.ParticleMeshCoupling1D(mesh, n, 3, "galerkin"), gp.ParticleMeshCoupling1D(mesh, n, 2, "galerkin")
    mx = gp.Maxwell1DFEM(mesh, 3)
    e1, e2, b, rho = np.zeros(nx), np.zeros(nx), 1e-2 * np.cos(np.arange(nx)), np.zeros(nx)
    gp.solve_poisson(e1, pg, ks0, mx, rho)
    h = gp.HamiltonianSplitting(1, 2, mx, ks0, ks1, pg, [e1, e2], b)
    hb = gp.HamiltonianSplittingBoris(mx, ks0, ks1, pg, [e1.copy(), e2.copy()], b.copy())
    h.strang_splitting(0.05, 2)                     # leaves a deferred kick pending on the group
    ref = e1.copy()
    for obj in (pg, ks0, ks1, mx):
        obj.close()                                 # handles gone, objects retained by the two splittings
    h.strang_splitting(0.05, 1)                     # still works on the retained objects
    assert np.all(np.isfinite(e1)) and not np.array_equal(e1, ref)
    j = h.j_dofs
    assert np.all(np.isfinite(j[1]))
    h.close()
    hb.close()                                      # the last user frees them
    # 2d3v the same way
    pg2 = gp.ParticleGroup(2, 3, 5000)
    pg2.sample("uniform", 0.0, L, sigma=(1.0, 1.0, 1.0))
    m2 = gp.TwoDMaxwell(gp.TwoDGrid(0.0, L, 8, 0.0, L, 8), 3)
    ee, bb = [np.zeros(64) for _ in range(3)], [np.zeros(64) for _ in range(3)]
    h2 = gp.HamiltonianSplitting2D3V(m2, pg2, ee, bb)
    h2.strang_splitting(0.05, 1)
    pg2.close()
    m2.close()
    h2.strang_splitting(0.05, 1)
    h2.close()


def test_exit_without_finalize_is_clean():
    """raw ctypes (no atexit hook of the Python mirror): objects alive, a kick pending, the process just ends"""
    code = textwrap.dedent("""
        import ctypes as C, sys
        L = C.CDLL(sys.argv[1])
        assert L.gempic_init(C.c_int(0)) == 0
        h = [C.c_uint64(0) for _ in range(5)]
        d = C.c_double
        assert L.gempic_pg_create(C.c_int(1), C.c_int(2), C.c_int(1), C.c_int64(10000), d(1.0), d(1.0), d(0.0), C.byref(h[0])) == 0
        sig = (d * 3)(1.0, 1.0, 1.0)
        assert L.gempic_pg_sample(h[0], C.c_int(0), d(0.0), d(12.0), d(0.0), d(1.0), sig, C.c_uint64(1), C.c_int64(0)) == 0
        assert L.gempic_pmc1d_create(d(0.0), d(12.0), C.c_int(32), C.c_int64(10000), C.c_int(3), C.c_int(1), C.byref(h[1])) == 0
        assert L.gempic_pmc1d_create(d(0.0), d(12.0), C.c_int(32), C.c_int64(10000), C.c_int(2), C.c_int(1), C.byref(h[2])) == 0
        assert L.gempic_maxwell1d_create(d(0.0), d(12.0), C.c_int(32), C.c_int(3), C.byref(h[3])) == 0
        assert L.gempic_hs_create(C.c_int(1), C.c_int(2), h[3], h[1], h[2], h[0], C.byref(h[4])) == 0
        assert L.gempic_hs_strang_splitting(h[4], d(0.05), C.c_int64(2)) == 0
        assert L.gempic_pg_destroy(h[0]) == 0        # still retained by the splitting
        assert L.gempic_synchronize() == 0
        print("ok")
    """)
    lib = os.path.join(ROOT, "gempic.jl_b200", "libgempic_b200.so")
    r = subprocess.run([sys.executable, "-c", code, lib], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, (r.returncode, r.stdout[-500:], r.stderr[-1500:])
