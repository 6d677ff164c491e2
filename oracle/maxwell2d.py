"""numpy restatement of the reference's 2D field solvers (CPU ORACLE, part 2).

TEST INFRASTRUCTURE ONLY -- imported by tests/ (and bench.py's CPU legs), never by anything
under gempic.jl_b200/.

Follows, function by function (file:line into /root/reference):
  spline_fem_mass_line                src/poisson_2d_fem.jl:87-106
  spline_fem_compute_mass_eig         src/poisson_2d_fem.jl:113-126
  spline_fem_mixedmass_line           src/maxwell_2d_fem.jl:89-115
  TwoDLinearSolverSplineMass + solve  src/linear_solver_spline_mass_2d.jl:1-32
  TwoDPoisson + compute_e_from_rho!   src/poisson_2d_fem.jl:23-84,184-262
  TwoDMaxwell and its methods         src/maxwell_2d_fem.jl:11-87,124-575
  eval_uniform_periodic_spline_curve  src/low_level_bsplines.jl:120-137

Third-party arithmetic: FFTW.jl (`fft!`/`ifft!`, compat "1", no Manifest) is replaced by
numpy's pocketfft -- any correct DFT gives the same dofs to round-off; FastGaussQuadrature's
`gausslegendre` by the oracle's Newton iteration on Legendre polynomials (oracle.gausslegendre).

Parity status: the reference's own test of this path (test/test_maxwell_2d_fem.jl) is
*analytic* (rtol 1e-4 .. 1e-3 against plane-wave / standing-wave solutions, `rho ≈ rho_ref`
for the discrete Gauss law); tests/test_oracle_maxwell2d.py restates it line by line and the
oracle passes it.  There are no dof-level golden vectors for the 2D fields in the reference.

Dofs are flat vectors of nx*ny doubles, x index fastest (`ind2d = (j-1)*nx + i`,
src/maxwell_2d_fem.jl:389-391).
"""
from __future__ import annotations

import numpy as np

from . import oracle as _o


def _basis(deg: int, x: float) -> np.ndarray:
    return _o.bsplines_eval_basis(deg, x)


def _gauss01(n: int):
    """gausslegendre(n) rescaled to [0, 1] (maxwell_2d_fem.jl:95-98)"""
    x, w = _o.gausslegendre(n)
    return (x + 1.0) / 2.0, w / 2.0


def spline_fem_mass_line(degree: int) -> np.ndarray:
    """src/poisson_2d_fem.jl:87-106"""
    n = degree + 1
    x, w = _gauss01(n)
    val = np.zeros((degree + 1, n))
    for j in range(n):
        val[:, j] = _basis(degree, x[j])
    line = np.zeros(degree + 1)
    for j in range(1, degree + 2):
        for i in range(j, degree + 2):
            for k in range(n):
                line[j - 1] += val[i - 1, k] * val[i - j, k] * w[k]
    return line


def spline_fem_mixedmass_line(deg: int) -> np.ndarray:
    """src/maxwell_2d_fem.jl:89-115"""
    n = min(3 * deg + 1, 10)
    x, w = _gauss01(n)
    v0 = np.zeros((deg + 1, n))
    v1 = np.zeros((deg, n))
    for j in range(n):
        v0[:, j] = _basis(deg, x[j])
        v1[:, j] = _basis(deg - 1, x[j])
    line = np.zeros(2 * deg)
    for j in range(2, deg + 2):
        for i in range(j, deg + 2):
            for k in range(n):
                line[j + deg - 2] += v0[i - 1, k] * v1[i - j, k] * w[k]
    for j in range(-deg + 1, 1):
        for i in range(1, deg + j + 1):
            for k in range(n):
                line[j + deg - 1] += v0[i - 1, k] * v1[i - j - 1, k] * w[k]
    return line


def spline_fem_compute_mass_eig(n_cells: int, degree: int, mass_line: np.ndarray) -> np.ndarray:
    """src/poisson_2d_fem.jl:113-126"""
    eig = np.zeros(n_cells)
    factor = 2.0 * np.pi / n_cells
    for k in range(n_cells):
        eig[k] = mass_line[0]
        for j in range(1, degree + 1):
            eig[k] += mass_line[j] * 2.0 * np.cos(factor * k * j)
    return eig


def eval_uniform_periodic_spline_curve(degree: int, scoef: np.ndarray) -> np.ndarray:
    """src/low_level_bsplines.jl:120-137"""
    bspl = _basis(degree, 0.0)
    n = len(scoef)
    out = np.zeros(n)
    for i in range(n):
        val = 0.0
        for j in range(1, degree + 1):
            val += bspl[j - 1] * scoef[(i - j) % n]
        out[i] = val
    return out


def spline_fem_multiply_mass(n_cells: int, degree: int, mass: np.ndarray, invec: np.ndarray) -> np.ndarray:
    """src/maxwell_2d_fem.jl:300-338 (periodic banded multiply, same association order)"""
    out = np.zeros(n_cells)
    for row in range(n_cells):
        acc = mass[0] * invec[row]
        for col in range(1, degree + 1):
            acc += mass[col] * (invec[(row + col) % n_cells] + invec[(row - col) % n_cells])
        out[row] = acc
    return out


class TwoDLinearSolverSplineMass:
    """src/linear_solver_spline_mass_2d.jl:1-14"""

    def __init__(self, nx1, nx2, eigvals1, eigvals2):
        self.nx1, self.nx2 = nx1, nx2
        self.eigvals1, self.eigvals2 = np.array(eigvals1), np.array(eigvals2)

    def solve(self, rhs: np.ndarray) -> np.ndarray:
        """:16-32  fft2 -> divide by eig1[i]*eig2[j] -> ifft2 -> real"""
        wk = rhs.reshape(self.nx2, self.nx1).T.astype(complex)  # wk[i, j], column-major reshape
        wk = np.fft.fft2(wk)
        wk /= self.eigvals1[:, None] * self.eigvals2[None, :]
        wk = np.fft.ifft2(wk)
        return np.real(wk).T.reshape(-1).copy()


class TwoDPoisson:
    """src/poisson_2d_fem.jl:23-84"""

    def __init__(self, mesh, degree):
        nx, ny = mesh.nx, mesh.ny
        dx, dy = (mesh.xmax - mesh.xmin) / nx, (mesh.ymax - mesh.ymin) / ny
        self.nx, self.ny, self.degree, self.dx, self.dy = nx, ny, degree, dx, dy
        m0 = spline_fem_mass_line(degree)
        m1 = spline_fem_mass_line(degree - 1)
        self.eig_values_mass_0_1 = spline_fem_compute_mass_eig(nx, degree, m0 * dx)
        e11 = spline_fem_compute_mass_eig(nx, degree - 1, m1 * dx)
        self.eig_values_mass_0_2 = spline_fem_compute_mass_eig(ny, degree, m0 * dy)
        e12 = spline_fem_compute_mass_eig(ny, degree - 1, m1 * dy)
        self.eig_values_d1 = np.zeros(nx, complex)
        self.eig_values_dtm1d_1 = np.zeros(nx)
        for j in range(1, nx):
            a = 2.0 * np.pi * j / nx
            self.eig_values_d1[j] = (1 - np.cos(a)) / dx + 1j * np.sin(a) / dx
            self.eig_values_dtm1d_1[j] = 2.0 / dx**2 * (1 - np.cos(a)) * e11[j]
        self.eig_values_d2 = np.zeros(ny, complex)
        self.eig_values_dtm1d_2 = np.zeros(ny)
        for j in range(1, ny):
            a = 2.0 * np.pi * j / ny
            self.eig_values_d2[j] = (1 - np.cos(a)) / dy + 1j * np.sin(a) / dy
            self.eig_values_dtm1d_2[j] = 2.0 / dy**2 * (1 - np.cos(a)) * e12[j]

    def compute_e_from_rho(self, efield, rho):
        """:237-262"""
        nx, ny = self.nx, self.ny
        s = np.fft.fft2(rho.reshape(ny, nx).T.astype(complex))
        eig = (self.eig_values_dtm1d_1[:, None] * self.eig_values_mass_0_2[None, :]
               + self.eig_values_mass_0_1[:, None] * self.eig_values_dtm1d_2[None, :])
        eig[0, 0] = 1.0
        s = s / eig
        s[0, 0] = 0.0
        sx = -s * self.eig_values_d1[:, None]
        sy = -s * self.eig_values_d2[None, :]
        efield[0][:] = np.real(np.fft.ifft2(sx)).T.reshape(-1)
        efield[1][:] = np.real(np.fft.ifft2(sy)).T.reshape(-1)


def _form_degrees(deg0, deg1, component, form):
    """maxwell_2d_fem.jl:136-149 / 222-236 (component is 1-based)"""
    if form == 0:
        d = [deg0, deg0]
    elif form == 1:
        d = [deg0, deg0]
        if component < 3:
            d[component - 1] = deg1
    elif form == 2:
        d = [deg1, deg1]
        if component < 3:
            d[component - 1] = deg0
    elif form == 3:
        d = [deg1, deg1]
    else:
        raise ValueError("Wrong form")
    return d


class TwoDMaxwell:
    """src/maxwell_2d_fem.jl:11-87"""

    def __init__(self, mesh, degree):
        nx, ny = mesh.nx, mesh.ny
        dx, dy = (mesh.xmax - mesh.xmin) / nx, (mesh.ymax - mesh.ymin) / ny
        self.mesh, self.nx, self.ny, self.dx, self.dy = mesh, nx, ny, dx, dy
        self.s_deg_0, self.s_deg_1 = degree, degree - 1
        l0, l1 = spline_fem_mass_line(degree), spline_fem_mass_line(degree - 1)
        self.mass_line_0 = [l0 * dx, l0 * dy]
        self.mass_line_1 = [l1 * dx, l1 * dy]
        lm = spline_fem_mixedmass_line(degree)
        self.mass_line_mixed = [lm * dx, lm * dy]
        e01 = spline_fem_compute_mass_eig(nx, degree, self.mass_line_0[0])
        e02 = spline_fem_compute_mass_eig(ny, degree, self.mass_line_0[1])
        e11 = spline_fem_compute_mass_eig(nx, degree - 1, self.mass_line_1[0])
        e12 = spline_fem_compute_mass_eig(ny, degree - 1, self.mass_line_1[1])
        S = TwoDLinearSolverSplineMass
        self.inv_mass_1 = [S(nx, ny, e11, e02), S(nx, ny, e01, e12), S(nx, ny, e01, e02)]
        self.inv_mass_2 = [S(nx, ny, e01, e12), S(nx, ny, e11, e02), S(nx, ny, e11, e12)]
        self.poisson = TwoDPoisson(mesh, degree)

    # ---- right-hand sides and projections ---------------------------------------------------
    def compute_rhs_from_function(self, f, component, form):
        """:124-196 (== compute_fem_rhs! :211-277)"""
        nx, ny, dx, dy = self.nx, self.ny, self.dx, self.dy
        d1, d2 = _form_degrees(self.s_deg_0, self.s_deg_1, component, form)
        x1, w1 = _gauss01(d1 + 1)
        x2, w2 = _gauss01(d2 + 1)
        b1 = np.array([_basis(d1, x1[k]) for k in range(d1 + 1)])
        b2 = np.array([_basis(d2, x2[k]) for k in range(d2 + 1)])
        out = np.zeros(nx * ny)
        c = 0
        for i2 in range(1, ny + 1):
            for i1 in range(1, nx + 1):
                coef = 0.0
                for j1 in range(1, d1 + 2):
                    for j2 in range(1, d2 + 2):
                        for k1 in range(d1 + 1):
                            for k2 in range(d2 + 1):
                                x = dx * (x1[k1] + i1 + j1 - 2)
                                y = dy * (x2[k2] + i2 + j2 - 2)
                                coef += w1[k1] * w2[k2] * f(x, y) * b1[k1, d1 + 1 - j1] * b2[k2, d2 + 1 - j2]
                out[c] = coef * dx * dy
                c += 1
        return out

    def l2projection(self, f, component, form):
        """:283-293"""
        rhs = self.compute_rhs_from_function(f, component, form)
        if form == 1:
            return self.inv_mass_1[component - 1].solve(rhs)
        if form == 2:
            return self.inv_mass_2[component - 1].solve(rhs)
        raise ValueError("l2projection: form must be 1 or 2")

    # ---- mass multiply ----------------------------------------------------------------------
    def multiply_mass_2dkron(self, mass_line_1, mass_line_2, c_in):
        """:343-363"""
        nx, ny = self.nx, self.ny
        deg1, deg2 = len(mass_line_1) - 1, len(mass_line_2) - 1
        wk = c_in.reshape(ny, nx).T.copy()
        for j in range(ny):
            wk[:, j] = spline_fem_multiply_mass(nx, deg1, mass_line_1, wk[:, j].copy())
        for i in range(nx):
            wk[i, :] = spline_fem_multiply_mass(ny, deg2, mass_line_2, wk[i, :].copy())
        return wk.T.reshape(-1).copy()

    def _mass_lines(self, component, form):
        """the (mass_line_1, mass_line_2) pair used for a (component, form) (:506-570)"""
        d1, d2 = _form_degrees(0, 1, component, form)
        pick = lambda axis, which: (self.mass_line_0 if which == 0 else self.mass_line_1)[axis]
        return pick(0, d1), pick(1, d2)

    # ---- field updates ----------------------------------------------------------------------
    def compute_e_from_rho(self, efield, rho):
        """:199-201"""
        self.poisson.compute_e_from_rho(efield, rho)

    def compute_e_from_b(self, e, dt, b):
        """:370-412"""
        nx, ny, dx1, dx2 = self.nx, self.ny, self.dx, self.dy
        work = [self.multiply_mass_2dkron(self.mass_line_0[0], self.mass_line_1[1], b[0]),
                self.multiply_mass_2dkron(self.mass_line_1[0], self.mass_line_0[1], b[1]),
                self.multiply_mass_2dkron(self.mass_line_1[0], self.mass_line_1[1], b[2])]
        W = [w.reshape(ny, nx) for w in work]           # W[c][j, i]
        up1 = lambda a: np.roll(a, -1, axis=1)           # i+1 periodic
        up2 = lambda a: np.roll(a, -1, axis=0)           # j+1 periodic
        curl = [-(W[2] - up2(W[2])) / dx2,
                (W[2] - up1(W[2])) / dx1,
                (W[0] - up2(W[0])) / dx2 - (W[1] - up1(W[1])) / dx1]
        for c in range(3):
            e[c][:] = e[c] + dt * self.inv_mass_1[c].solve(curl[c].reshape(-1))

    def compute_b_from_e(self, b, dt, e):
        """:423-444"""
        nx, ny, dx1, dx2 = self.nx, self.ny, self.dx, self.dy
        E = [v.reshape(ny, nx) for v in e]
        dn1 = lambda a: np.roll(a, 1, axis=1)            # i-1 periodic
        dn2 = lambda a: np.roll(a, 1, axis=0)            # j-1 periodic
        b[0][:] += (-dt * (E[2] - dn2(E[2])) / dx2).reshape(-1)
        b[1][:] += (dt * (E[2] - dn1(E[2])) / dx1).reshape(-1)
        b[2][:] += (-dt * ((E[1] - dn1(E[1])) / dx1 - (E[0] - dn2(E[0])) / dx2)).reshape(-1)

    def compute_e_from_j(self, e, current, component):
        """:455-459"""
        e[:] = e - self.inv_mass_1[component - 1].solve(current)

    def compute_rho_from_e(self, rho, efield):
        """:468-500"""
        nx, ny, dx1, dx2 = self.nx, self.ny, self.dx, self.dy
        w1 = self.multiply_mass_2dkron(self.mass_line_1[0], self.mass_line_0[1], efield[0]).reshape(ny, nx)
        w2 = self.multiply_mass_2dkron(self.mass_line_0[0], self.mass_line_1[1], efield[1]).reshape(ny, nx)
        r = (w1 - np.roll(w1, -1, axis=1)) / dx1 + (w2 - np.roll(w2, -1, axis=0)) / dx2
        rho[:] = -r.reshape(-1)

    def inner_product(self, c1, c2, component, form):
        """:512-575"""
        l1, l2 = self._mass_lines(component, form)
        return float(np.sum(c1 * self.multiply_mass_2dkron(l1, l2, c2)))
