"""The reference's samplers on the device (csrc/sampling.cu) against the restatement (oracle/sampling.py) and against
the assertions of the reference's own test (test/test_sampling.jl:43-123).

Tolerances: Sobol coordinates, x of the symmetric / plain Sobol loads and the weights derived from them are exact up to
the last bits of cos (1e-15); the Sobol + Newton Landau load is deterministic on both sides and agrees to 1e-13 (device
sin/cos/log vs libm, amplified by the Newton iteration); normal deviates are statistical on both sides."""
import math

import numpy as np
import pytest

from oracle import sampling as smp

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("V", [1, 2])
def test_landau_sobol_newton_load_matches_restatement(gp, V):
    n, L = 20_000, 4 * math.pi
    ref = np.zeros((2 + V, n))
    smp.sample_landau(ref, 0.5, 0.5, 1.0, L)
    pg = gp.ParticleGroup(1, V, n)
    gp.sample(pg, 0.5, 0.5, 1.0, gp.OneDGrid(0.0, L, 32))
    a = pg.to_host()
    assert np.max(np.abs(a[0] - ref[0])) < 1e-13 * L
    assert np.max(np.abs(a[1:1 + V] - ref[1:1 + V])) < 1e-13 * np.max(np.abs(ref[1:1 + V]))
    assert np.array_equal(a[1 + V], ref[1 + V])
    # through ParticleSampler + CosSumGaussian for {1,1} (particle_sampling.jl:248-256)
    if V == 1:
        df = gp.CosSumGaussian(1, 1, [[0.5]], [0.5], [[1.0]], [[0.0]])
        pg2 = gp.ParticleGroup(1, 1, n)
        gp.sample(pg2, gp.ParticleSampler(1, 1, "sobol", False, n), df, gp.OneDGrid(0.0, L, 32))
        assert np.array_equal(pg2.to_host(), a)


def test_landau_load_is_sharding_independent_and_hits_the_known_answers(gp):
    """SURVEY.md Appendix B at N = 1e5; two shards give the same particles as one group"""
    n, L, nx = 100_000, 4 * math.pi, 32
    mesh = gp.OneDGrid(0.0, L, nx)
    pg = gp.ParticleGroup(1, 2, n)
    gp.sample(pg, 0.5, 0.5, 1.0, mesh)
    whole = pg.to_host()
    cut = 37_123
    pa, pb = gp.ParticleGroup(1, 2, cut, common_weight=1.0 / n), gp.ParticleGroup(1, 2, n - cut, common_weight=1.0 / n)
    gp.sample(pa, 0.5, 0.5, 1.0, mesh, first_index=0, n_global=n)
    gp.sample(pb, 0.5, 0.5, 1.0, mesh, first_index=cut, n_global=n)
    assert np.array_equal(np.concatenate([pa.to_host(), pb.to_host()], axis=1), whole)
    ks0 = gp.ParticleMeshCoupling1D(mesh, n, 3, "galerkin")
    mx = gp.Maxwell1DFEM(mesh, 3)
    e1, rho = np.zeros(nx), np.zeros(nx)
    gp.solve_poisson(e1, pg, ks0, mx, rho)
    assert abs(rho.sum() - 4 * math.pi) < 1e-11
    assert abs(mx.inner_product(e1, e1, 2) - 6.28313) < 2e-5
    assert abs(whole[1].mean() + 5.0e-5) < 2e-6 and abs(whole[1].var() - 0.99990) < 2e-5
    # sample!(d::LandauDamping, pg) (landau_damping.jl:34-59): same x, v; weights 2 pi / kx / N
    pg3 = gp.ParticleGroup(1, 2, n)
    gp.sample(gp.LandauDamping(0.5, 0.5), pg3)
    b = pg3.to_host()
    assert np.array_equal(b[:3], whole[:3]) and np.allclose(b[3], 2 * math.pi / 0.5 / n, rtol=1e-15)


def moments(a):
    mean = a[:3].mean(axis=1)
    var = ((a[:3] - mean[:, None]) ** 2).sum(axis=1) / (a.shape[1] - 1)
    return mean, var


@pytest.mark.parametrize("typ", ["sobol", "random"])
@pytest.mark.parametrize("sym", [False, True])
def test_reference_sampling_test_on_gpu(gp, typ, sym):
    """test/test_sampling.jl:43-123 with the device sampler; for :sobol the x coordinates and weights equal the
    restatement's value by value"""
    n, xmin = 100_000, 1.0
    Lx = 4 * math.pi
    mesh = gp.OneDGrid(xmin, xmin + Lx, 64)
    df1 = gp.CosSumGaussian(1, 2, [[0.5]], [0.01], [[0.1, 2.0]], [[0.0, 0.0]])
    ps = gp.ParticleSampler(1, 2, typ, sym, n)
    pg = gp.ParticleGroup(1, 2, ps.n_particles)
    gp.sample(pg, ps, df1, mesh)
    a = pg.to_host()
    mean, var = moments(a)
    tol = 1e-12 if sym else 1e2 / math.sqrt(n)
    assert np.max(np.abs(mean - np.array([Lx * 0.5 + xmin, 0.0, 0.0]))) < tol
    assert abs(var[0] - Lx ** 2 / 12) < 0.2 and abs(var[1] - 0.01) < 1e-3 and abs(var[2] - 4.0) < 0.1
    if typ == "sobol":
        ref = np.zeros((4, ps.n_particles))
        smp.sample(ref, smp.ParticleSampler(typ, sym, n), smp.CosGaussian([[0.5]], [0.01], [[0.1, 2.0]], [[0.0, 0.0]]), xmin, mesh.dimx)
        assert np.max(np.abs(a[0] - ref[0])) < 1e-14 * Lx
        assert np.max(np.abs(a[3] - ref[3])) < 1e-14 * Lx
    if sym:   # the 8-fold antithetic structure, exactly (particle_sampling.jl:210-217)
        g = a.reshape(4, -1, 8)
        assert np.array_equal(g[0, :, 4], mesh.dimx - g[0, :, 0] + 2.0 * xmin) and np.array_equal(g[0, :, 0], g[0, :, 3])
        assert np.array_equal(g[1, :, 1], -g[1, :, 0]) and np.array_equal(g[2, :, 2], -g[2, :, 1])
        assert np.array_equal(g[1, :, 7], g[1, :, 0]) and np.array_equal(g[2, :, 7], g[2, :, 0])
        assert np.all(g[3] == g[3, :, :1])
    # sharding independence
    cut = 40_000
    pa = gp.ParticleGroup(1, 2, cut, common_weight=1.0 / n)
    pb = gp.ParticleGroup(1, 2, ps.n_particles - cut, common_weight=1.0 / n)
    gp.sample(pa, ps, df1, mesh, first_index=0)
    gp.sample(pb, ps, df1, mesh, first_index=cut)
    assert np.array_equal(np.concatenate([pa.to_host(), pb.to_host()], axis=1), a)


def test_two_gaussians_and_errors(gp):
    n, xmin, Lx = 100_000, 1.0, 4 * math.pi
    mesh = gp.OneDGrid(xmin, xmin + Lx, 64)
    df2 = gp.CosSumGaussian(1, 2, [[0.5]], [0.01], [[0.1, 2.0], [2.0, 2.0]], [[0.0, 0.0], [1.0, 1.0]], [0.7, 0.3])
    for sym in (False, True):
        ps = gp.ParticleSampler(1, 2, "sobol", sym, n)
        pg = gp.ParticleGroup(1, 2, ps.n_particles)
        gp.sample(pg, ps, df2, mesh)
        mean, _ = moments(pg.to_host())
        assert np.max(np.abs(mean - np.array([Lx * 0.5 + xmin, 0.3, 0.3]))) < 1e2 / math.sqrt(n)   # test_sampling.jl:112-122
        if sym:   # the Gaussian is drawn from the 4th Sobol coordinate (:204-208): 30 % of the groups take the second one
            assert abs(mean[1] - 0.3) < 0.02 and abs(mean[2] - 0.3) < 0.02
    with pytest.raises(gp.ArgumentError):
        gp.ParticleSampler(1, 2, "halton", False, 10)
    with pytest.raises(gp.AssertionFailed):
        gp.CosSumGaussian(1, 2, [[0.5]], [0.01], [[0.1, 2.0]], [[0.0, 0.0]], [0.5])
    assert gp.ParticleSampler(1, 2, "sobol", True, 100_003).n_particles == 100_006
