"""CPU restatement of the reference's particle samplers -- TEST INFRASTRUCTURE ONLY (never imported by the product).

  SobolSeq / next            Sobol.jl (dependency of the reference, Project.toml; not vendored under /root/reference):
                             Gray-code Sobol sequence, Joe-Kuo direction numbers new-joe-kuo-6.21201, the all-zero
                             point skipped -- restated from the published algorithm (Joe & Kuo 2008; Antonov-Saleev)
                             and pinned against scipy.stats.qmc.Sobol(scramble=False) in tests/test_oracle_sampling.py
  newton                     src/particle_sampling.jl:236-245 (and landau_damping.jl:37-46), iteration for iteration
  sample_landau              src/particle_sampling.jl:266-282 ({1,1}) and :294-311 ({1,2})
  sample_landau_damping      src/landau_damping.jl:34-59
  sample_all / sample_sym    src/particle_sampling.jl:89-143 / :150-225; the normal deviates come from numpy's generator
                             (Julia's MersenneTwister + ziggurat stream is not reproducible outside Julia), so only the
                             Sobol coordinates, the weights and the antithetic structure are comparable value by value.
  ParticleSampler            src/particle_sampling.jl:14-59 (including its `n_particles += mod(n_particles, 8)` rule)

Pins: scipy's unscrambled Sobol points (bitwise); the known answers of SURVEY.md Appendix B for the config-1/3 load
(total charge 4 pi, PotentialEnergyE1 6.28313, KineticEnergy 25.13265, mean/var of v1); the moment tolerances of
test/test_sampling.jl:43-123."""
from __future__ import annotations

import math

import numpy as np

# Joe-Kuo: (s, a, m_init) for dimensions 2..4; dimension 1 is van der Corput
_JK = {2: (1, 0, [1]), 3: (2, 1, [1, 3]), 4: (3, 1, [1, 3, 1])}
_BITS = 32


def direction_numbers(dim: int):
    """m_c (c = 0..31) of dimension `dim` (1-based)"""
    if dim == 1:
        return [1] * _BITS
    s, a, m = _JK[dim]
    m = list(m)
    for c in range(s, _BITS):
        val = m[c - s] ^ (m[c - s] << s)
        for k in range(1, s):
            if (a >> (s - 1 - k)) & 1:
                val ^= m[c - k] << k
        m.append(val)
    return m


class SobolSeq:
    """SobolSeq(N) with next!() of Sobol.jl: state x_n = x_{n-1} XOR v_c, c = number of trailing zeros of n"""

    def __init__(self, ndims: int):
        assert 1 <= ndims <= 4
        self.ndims, self.n = ndims, 0
        self.v = [[m << (31 - c) for c, m in enumerate(direction_numbers(d + 1))] for d in range(ndims)]
        self.x = [0] * ndims

    def next(self):
        self.n += 1
        c = (self.n & -self.n).bit_length() - 1
        for d in range(self.ndims):
            self.x[d] ^= self.v[d][c]
        return [xi / 4294967296.0 for xi in self.x]


def sobol_points(ndims: int, n: int, first: int = 0) -> np.ndarray:
    """points first+1 .. first+n (vectorised: x_p = XOR over the set bits of gray(p))"""
    p = np.arange(first + 1, first + n + 1, dtype=np.uint64)
    g = p ^ (p >> np.uint64(1))
    out = np.zeros((n, ndims))
    for d in range(ndims):
        v = [m << (31 - c) for c, m in enumerate(direction_numbers(d + 1))]
        x = np.zeros(n, dtype=np.uint64)
        for c in range(_BITS):
            x ^= np.where((g >> np.uint64(c)) & np.uint64(1), np.uint64(v[c]), np.uint64(0))
        out[:, d] = x.astype(np.float64) / 4294967296.0
    return out


def newton(r: float, alpha: float, k: float) -> float:
    x0, x1 = 0.0, 1.0
    r *= 2 * math.pi / k
    while abs(x1 - x0) > 1e-12:
        p = x0 + alpha * math.sin(k * x0) / k
        f = 1 + alpha * math.cos(k * x0)
        x0, x1 = x1, x0 - (p - r) / f
    return x1


def sample_landau(array: np.ndarray, alpha: float, k: float, sigma: float, dimx: float):
    """array: (3, N) for {1,1} or (4, N) for {1,2}, filled in place"""
    V = array.shape[0] - 2
    nbpart = array.shape[1]
    s = SobolSeq(2)
    for i in range(1, nbpart + 1):
        v = sigma * math.sqrt(-2 * math.log((i - 0.5) / nbpart))
        r1, r2 = s.next()
        theta = r1 * 2 * math.pi
        array[0, i - 1] = newton(r2, alpha, k)
        array[1, i - 1] = v * math.cos(theta)
        if V == 2:
            array[2, i - 1] = v * math.sin(theta)
        array[1 + V, i - 1] = dimx


def sample_landau_damping(array: np.ndarray, alpha: float, kx: float):
    """sample!(d::LandauDamping, pg::ParticleGroup{1,2})"""
    nbpart = array.shape[1]
    sample_landau(array, alpha, kx, 1.0, 2 * math.pi / kx / nbpart)


class CosGaussian:
    """CosSumGaussian{1,2} / SumCosGaussian{1,2} parameters (src/distributions.jl:13-157); both share eval_x_density"""

    def __init__(self, k, alpha, sigma, mu, delta=(1.0,)):
        self.k = [float(np.atleast_1d(kk)[0]) for kk in k]
        self.alpha = [float(a) for a in alpha]
        self.sigma = [list(map(float, s)) for s in sigma]
        self.mu = [list(map(float, m)) for m in mu]
        self.delta = [float(d) for d in delta]
        assert len(self.k) == len(self.alpha) and len(self.sigma) == len(self.mu) == len(self.delta)
        assert sum(self.delta) == 1.0
        self.n_cos, self.n_gaussians = len(self.k), len(self.sigma)

    def eval_x_density(self, x):
        f = 1.0
        for j in range(self.n_cos):
            f += self.alpha[j] * math.cos(self.k[j] * x)
        return f


class ParticleSampler:
    def __init__(self, sampling_type: str, symmetric: bool, n_particles: int, seed: int = 1234):
        if sampling_type not in ("random", "sobol"):
            raise ValueError(f"Sampling type {sampling_type} not implemented")
        if symmetric:
            rem = n_particles % 8
            if rem != 0:
                n_particles += rem   # (sic) particle_sampling.jl:35-38
        self.sampling_type, self.symmetric, self.n_particles, self.seed = sampling_type, symmetric, n_particles, seed


def sample_all(ps: ParticleSampler, array: np.ndarray, df: CosGaussian, xmin: float, dimx: float):
    rng = np.random.default_rng(ps.seed)
    sob = SobolSeq(1) if ps.sampling_type == "sobol" else None
    for i in range(array.shape[1]):
        x = xmin + (sob.next()[0] if sob else rng.uniform()) * dimx
        w = df.eval_x_density(x) * dimx
        v = rng.normal(size=2)
        ig = 0   # rdn is never filled in the reference (:132): the first Gaussian is always taken
        array[0, i] = x
        array[1, i] = v[0] * df.sigma[ig][0] + df.mu[ig][0]
        array[2, i] = v[1] * df.sigma[ig][1] + df.mu[ig][1]
        array[3, i] = w


def sample_sym(ps: ParticleSampler, array: np.ndarray, df: CosGaussian, xmin: float, dimx: float):
    rng = np.random.default_rng(ps.seed)
    sob = SobolSeq(4) if ps.sampling_type == "sobol" else None
    dcum = np.cumsum(df.delta)
    x = v1 = v2 = wi = 0.0
    ig = 0
    for i_part in range(1, array.shape[1] + 1):
        ip = i_part % 8
        if ip == 1:
            rdn = sob.next() if sob else list(rng.uniform(size=4))
            x = xmin + dimx * rdn[0]
            wi = df.eval_x_density(x) * dimx
            v = rng.normal(size=2)
            ig = 0
            while ig < df.n_gaussians - 1 and rdn[3] > dcum[ig]:
                ig += 1
            v1 = v[0] * df.sigma[ig][0] + df.mu[ig][0]
            v2 = v[1] * df.sigma[ig][1] + df.mu[ig][1]
        elif ip == 5:
            x = dimx - x + 2.0 * xmin
        elif ip % 2 == 0:
            v1 = -v1 + 2.0 * df.mu[ig][0]
        else:
            v2 = -v2 + 2.0 * df.mu[ig][1]
        array[:, i_part - 1] = (x, v1, v2, wi)


def sample(array, ps: ParticleSampler, df: CosGaussian, xmin: float, dimx: float):
    (sample_sym if ps.symmetric else sample_all)(ps, array, df, xmin, dimx)
