"""Shared builders: the same set-up code drives the CPU oracle (`orc`) and the CUDA product
(`gp`), whose classes deliberately share the reference's names and signatures."""
import math

import numpy as np


def landau_state(n, L, seed=1234, alpha=0.5, k=0.5, sigma=(1.0, 1.0), V=2):
    """Seeded synthetic Landau-like load in the reference record layout (rows x, v.., w)."""
    rng = np.random.default_rng(seed)
    u = rng.uniform(0, 1, n)
    x = u * L
    for _ in range(30):  # inverse CDF of 1 + alpha cos(kx)
        x -= (x + alpha / k * np.sin(k * x) - u * L) / (1 + alpha * np.cos(k * x))
    x = np.mod(x, L)
    rows = [x] + [sigma[i] * rng.normal(size=n) for i in range(V)] + [np.full(n, L)]
    return np.stack(rows)


def weibel_state(n, L, seed=1234, sigma=(0.2, 0.005773502691896)):
    rng = np.random.default_rng(seed)
    return np.stack([rng.uniform(0, L, n), sigma[0] * rng.normal(size=n), sigma[1] * rng.normal(size=n), np.full(n, L)])


class Sim1D:
    """mesh + particle group + two smoothers + Maxwell solver + fields, for `mod` in (orc, gp)."""

    def __init__(self, mod, state, L, nx=32, deg0=3, deg1=None, smoothing="galerkin", xmin=0.0, V=2,
                 charge=1.0, mass=1.0, common_weight=0.0, maxwell_degree=None):
        deg1 = deg0 - 1 if deg1 is None else deg1
        n = state.shape[1]
        self.mod, self.n, self.nx, self.L = mod, n, nx, L
        self.mesh = mod.OneDGrid(xmin, xmin + L, nx)
        self.pg = mod.ParticleGroup(1, V, n, charge=charge, mass=mass, common_weight=common_weight)
        self.pg.array[:, :] = state
        self.ks0 = mod.ParticleMeshCoupling1D(self.mesh, n, deg0, smoothing)
        self.ks1 = mod.ParticleMeshCoupling1D(self.mesh, n, deg1, smoothing)
        self.mx = mod.Maxwell1DFEM(self.mesh, deg0 if maxwell_degree is None else maxwell_degree)
        self.e1, self.e2, self.b = np.zeros(nx), np.zeros(nx), np.zeros(nx)
        self.rho = np.zeros(nx)

    def init_fields(self, b_amp=1e-2, e2_amp=0.0):
        i = np.arange(self.nx)
        self.b[:] = b_amp * np.cos(2 * math.pi * (i + 0.5) / self.nx)
        self.e2[:] = e2_amp * np.sin(2 * math.pi * i / self.nx)
        self.mod.solve_poisson(self.e1, self.pg, self.ks0, self.mx, self.rho)
        return self

    def splitting(self, V=2, **kw):
        self.h = self.mod.HamiltonianSplitting(1, V, self.mx, self.ks0, self.ks1, self.pg, [self.e1, self.e2], self.b, **kw)
        return self.h

    def boris(self, **kw):
        self.h = self.mod.HamiltonianSplittingBoris(self.mx, self.ks0, self.ks1, self.pg, [self.e1, self.e2], self.b, **kw)
        return self.h

    def particles(self):
        pg = self.pg
        return pg.to_host() if hasattr(pg, "to_host") else pg.array.copy()


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = max(np.max(np.abs(b)), 1e-300)
    return float(np.max(np.abs(a - b)) / scale)


def particle_err(a, b, L):
    """max relative deviation of particle rows; positions compared modulo the period (a particle
    within 1e-16 of the boundary may wrap on one side only)."""
    a, b = np.asarray(a), np.asarray(b)
    dx = np.abs(a[0] - b[0])
    dx = np.minimum(dx, np.abs(dx - L))
    ex = np.max(dx) / L
    ev = max(rel_err(a[r], b[r]) for r in range(1, a.shape[0]))
    return max(ex, ev)
