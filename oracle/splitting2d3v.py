"""CPU ORACLE (part 3): HamiltonianSplitting{2,3} = the C particle loops of gempic_oracle2d3v.c + the numpy
TwoDMaxwell of oracle/maxwell2d.py.

TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED at the integrator level: GEMPIC.jl has no 2d3v integrator
(src/hamiltonian_splitting_2d3v.jl is empty).  The operators are the 2D extension of
src/hamiltonian_splitting_1d2v.jl:41-236 and strang_splitting! of src/hamiltonian_splitting.jl:98-108;
tests/test_oracle_2d3v.py pins them (a) operator by operator to the golden-pinned 1d2v oracle on
x2-independent data and (b) to the invariants of the scheme (discrete Gauss law to round-off).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import maxwell2d as m2

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libgempic_oracle2d3v.so")
_dp = C.POINTER(C.c_double)


class _Mesh(C.Structure):
    _fields_ = [("xmin", C.c_double * 2), ("L", C.c_double * 2), ("d", C.c_double * 2), ("n", C.c_int * 2), ("deg", C.c_int)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "gempic_oracle2d3v.c")
        if not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
            subprocess.run(["make", "-C", _HERE, "libgempic_oracle2d3v.so", "-B"], check=True, capture_output=True)
        _lib = C.CDLL(_LIB)
    return _lib


def _p(a):
    assert a.dtype == np.float64
    return a.ctypes.data_as(_dp)


class ParticleGroup23:
    """ParticleGroup{2,3} (src/particle_group.jl:15-46): array is (6, N) column-major, rows x1,x2,v1,v2,v3,w."""

    def __init__(self, n_particles, charge=1.0, mass=1.0, common_weight=0.0):
        self.n_particles = int(n_particles)
        self._base = np.zeros((self.n_particles, 6))
        self.array = self._base.T
        self.charge, self.mass = float(charge), float(mass)
        self.common_weight = 1.0 / n_particles if common_weight == 0.0 else float(common_weight)
        self.q_over_m = self.charge / self.mass


class HamiltonianSplitting2D3V:
    """e_dofs, b_dofs: lists of three flat nx*ny arrays (aliased, like the 1D reference struct)."""

    def __init__(self, maxwell: m2.TwoDMaxwell, pg: ParticleGroup23, e_dofs, b_dofs):
        self.maxwell_solver, self.particle_group = maxwell, pg
        self.e_dofs, self.b_dofs = e_dofs, b_dofs
        mesh = maxwell.mesh
        self._m = _Mesh()
        self._m.xmin[:] = [mesh.xmin, mesh.ymin]
        self._m.L[:] = [mesh.xmax - mesh.xmin, mesh.ymax - mesh.ymin]
        self._m.d[:] = [maxwell.dx, maxwell.dy]
        self._m.n[:] = [maxwell.nx, maxwell.ny]
        self._m.deg = maxwell.s_deg_0
        n = maxwell.nx * maxwell.ny
        self.j_dofs = [np.zeros(n) for _ in range(3)]
        self._ws = pg.charge * pg.common_weight

    def _pa(self):
        return _p(self.particle_group._base)

    def operatorHE(self, dt):
        pg = self.particle_group
        lib().orc2_he(C.byref(self._m), self._pa(), C.c_int64(pg.n_particles), C.c_double(dt * pg.q_over_m),
                      _p(self.e_dofs[0]), _p(self.e_dofs[1]), _p(self.e_dofs[2]))
        self.maxwell_solver.compute_b_from_e(self.b_dofs, dt, self.e_dofs)

    def operatorHB(self, dt):
        self.maxwell_solver.compute_e_from_b(self.e_dofs, dt, self.b_dofs)

    def _hp12(self, d, dt):
        pg = self.particle_group
        j = self.j_dofs[d]
        j[:] = 0.0
        bo = self.b_dofs[1] if d == 0 else self.b_dofs[0]
        lib().orc2_hp12(C.byref(self._m), self._pa(), C.c_int64(pg.n_particles), C.c_int(d), C.c_double(dt),
                        C.c_double(pg.q_over_m), C.c_double(self._ws), _p(self.b_dofs[2]), _p(bo), _p(j))
        self.maxwell_solver.compute_e_from_j(self.e_dofs[d], j, d + 1)

    def operatorHp1(self, dt):
        self._hp12(0, dt)

    def operatorHp2(self, dt):
        self._hp12(1, dt)

    def operatorHp3(self, dt):
        pg = self.particle_group
        j = self.j_dofs[2]
        j[:] = 0.0
        lib().orc2_hp3(C.byref(self._m), self._pa(), C.c_int64(pg.n_particles), C.c_double(dt), C.c_double(pg.q_over_m),
                       C.c_double(self._ws), _p(self.b_dofs[0]), _p(self.b_dofs[1]), _p(j))
        j *= dt
        self.maxwell_solver.compute_e_from_j(self.e_dofs[2], j, 3)

    def strang_splitting(self, dt, number_steps):
        for _ in range(number_steps):
            self.operatorHB(0.5 * dt)
            self.operatorHE(0.5 * dt)
            self.operatorHp3(0.5 * dt)
            self.operatorHp2(0.5 * dt)
            self.operatorHp1(dt)
            self.operatorHp2(0.5 * dt)
            self.operatorHp3(0.5 * dt)
            self.operatorHE(0.5 * dt)
            self.operatorHB(0.5 * dt)

    # ---- diagnostics ------------------------------------------------------------------------------
    def charge_density(self):
        """rho dofs (degree p x p), add_charge! over all particles with get_charge weights"""
        pg = self.particle_group
        rho = np.zeros_like(self.j_dofs[0])
        lib().orc2_charge(C.byref(self._m), self._pa(), C.c_int64(pg.n_particles), C.c_double(self._ws), _p(rho))
        return rho

    def gauss_residual(self):
        """compute_rho_from_e!(E) - rho(particles): constant in time for the exact-line-integral scheme"""
        r = np.zeros_like(self.j_dofs[0])
        self.maxwell_solver.compute_rho_from_e(r, self.e_dofs)
        return r - self.charge_density()

    def energies(self):
        mx, pg = self.maxwell_solver, self.particle_group
        out = np.zeros(4)
        lib().orc2_moments(self._pa(), C.c_int64(pg.n_particles), _p(out))
        kin = 0.5 * pg.mass * pg.common_weight * out[0]
        ee = 0.5 * sum(mx.inner_product(self.e_dofs[c], self.e_dofs[c], c + 1, 1) for c in range(3))
        eb = 0.5 * sum(mx.inner_product(self.b_dofs[c], self.b_dofs[c], c + 1, 2) for c in range(3))
        return kin, ee, eb
