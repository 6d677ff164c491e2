"""Pins of the sampler restatement (oracle/sampling.py) -- CPU only.

  * Sobol.jl's sequence: bitwise equal to scipy's unscrambled Sobol points without their leading zero point (the same
    Gray-code / Joe-Kuo construction), from the incremental next!() form and from the direct
    (per-index) form the device kernels use;
  * the config-1/3 Landau load (particle_sampling.jl:266-311): known answers of SURVEY.md Appendix B;
  * sample_all / sample_sym: the assertions of test/test_sampling.jl:43-123."""
import math

import numpy as np
import pytest

from oracle import sampling as smp


def test_sobol_matches_scipy_bitwise():
    qmc = pytest.importorskip("scipy.stats.qmc")
    n = 4096
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = qmc.Sobol(4, scramble=False).random(n + 1)[1:]
    s = smp.SobolSeq(4)
    inc = np.array([s.next() for _ in range(n)])
    assert np.array_equal(inc, ref)
    assert np.array_equal(smp.sobol_points(4, n), ref)
    assert np.array_equal(smp.sobol_points(2, 1000, first=3000), ref[3000:4000, :2])
    assert smp.SobolSeq(2).next() == [0.5, 0.5]          # Sobol.jl skips the all-zero point


def test_library_sobol_table_matches(gp):
    """the direction numbers compiled into libgempic_b200 (host evaluation, no GPU needed)"""
    import ctypes as C

    out = np.zeros((5000, 4))
    rc = gp.load().gempic_sobol_points(C.c_int(4), C.c_int64(100), C.c_int64(5000), out.ctypes.data_as(C.POINTER(C.c_double)))
    assert rc == 0
    assert np.array_equal(out, smp.sobol_points(4, 5000, first=100))


def test_newton_inverts_the_cdf():
    for r in (0.0, 0.1, 0.37, 0.5, 0.93):
        x = smp.newton(r, 0.5, 0.5)
        assert abs(x + 0.5 * math.sin(0.5 * x) / 0.5 - r * 4 * math.pi) < 1e-11


def test_landau_load_known_answers(orc):
    """N = 1e5, alpha = 0.5, k = 0.5, sigma = 1, L = 4 pi, degree 3, 32 cells, galerkin"""
    n, L, nx = 100_000, 4 * math.pi, 32
    arr = np.zeros((4, n))
    smp.sample_landau(arr, 0.5, 0.5, 1.0, L)
    assert abs(arr[1].mean() - (-5.0e-5)) < 2e-6 and abs(arr[1].var() - 0.99990) < 2e-5
    mesh = orc.OneDGrid(0.0, L, nx)
    pg = orc.ParticleGroup(1, 2, n)
    pg.array[:, :] = arr
    ks0 = orc.ParticleMeshCoupling1D(mesh, n, 3, "galerkin")
    mx = orc.Maxwell1DFEM(mesh, 3)
    e1, rho = np.zeros(nx), np.zeros(nx)
    orc.solve_poisson(e1, pg, ks0, mx, rho)
    assert abs(rho.sum() - 12.56637061435917) < 1e-11
    assert abs(mx.inner_product(e1, e1, 2) - 6.28313) < 2e-5         # PotentialEnergyE1 (no 1/2)
    ke = float(np.sum((arr[1] ** 2 + arr[2] ** 2) * arr[3] / n))     # KineticEnergy (diagnostics.jl:206, no 1/2)
    assert abs(ke - 25.13265) < 2e-5
    # {1,1} form (:266-282) shares x and v1
    a11 = np.zeros((3, 1000))
    a12 = np.zeros((4, 1000))
    smp.sample_landau(a11, 0.5, 0.5, 1.0, L)
    smp.sample_landau(a12, 0.5, 0.5, 1.0, L)
    assert np.array_equal(a11[:2], a12[:2]) and np.all(a11[2] == L)


def moments(a):
    mean = a[:3].mean(axis=1)
    var = ((a[:3] - mean[:, None]) ** 2).sum(axis=1) / (a.shape[1] - 1)
    return mean, var


def test_reference_sampling_test():
    """test/test_sampling.jl:43-123"""
    n, xmin = 100_000, 1.0
    Lx = 4 * math.pi
    df1 = smp.CosGaussian([[0.5]], [0.01], [[0.1, 2.0]], [[0.0, 0.0]])
    mean_ref = np.array([Lx * 0.5 + xmin, 0.0, 0.0])
    for typ in ("sobol", "random"):
        for sym in (False, True):
            ps = smp.ParticleSampler(typ, sym, n)
            a = np.zeros((4, ps.n_particles))
            smp.sample(a, ps, df1, xmin, Lx)
            mean, var = moments(a)
            tol = 1e-12 if sym else 1e2 / math.sqrt(n)
            assert np.max(np.abs(mean - mean_ref)) < tol, (typ, sym)
            assert abs(var[0] - Lx ** 2 / 12) < 0.2 and abs(var[1] - 0.01) < 1e-3 and abs(var[2] - 4.0) < 0.1
    df2 = smp.CosGaussian([[0.5]], [0.01], [[0.1, 2.0], [2.0, 2.0]], [[0.0, 0.0], [1.0, 1.0]], [0.7, 0.3])
    for sym in (False, True):
        ps = smp.ParticleSampler("sobol", sym, n)
        a = np.zeros((4, ps.n_particles))
        smp.sample(a, ps, df2, xmin, Lx)
        mean, _ = moments(a)
        assert np.max(np.abs(mean - np.array([Lx * 0.5 + xmin, 0.3, 0.3]))) < 1e2 / math.sqrt(n)
    # ParticleSampler's particle-count rule (:33-38)
    assert smp.ParticleSampler("sobol", True, 100_003).n_particles == 100_006
    with pytest.raises(ValueError):
        smp.ParticleSampler("halton", False, 10)
