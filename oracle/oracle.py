"""ctypes front-end of the CPU ORACLE (oracle/gempic_oracle.c).

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module; nothing under
gempic.jl_b200/ does.  Class and function names mirror the reference
(GEMPIC.jl, /root/reference/src) so that the tests read like the reference's own.

Parity status: PINNED by the reference's golden vectors (tests/golden/*.json,
checked in tests/test_oracle_golden.py).  The reference itself (Julia) cannot run
in this image, so there is no oracle/_ref.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libgempic_oracle.so")

_dp = C.POINTER(C.c_double)
_vp = C.c_void_p


def build(force: bool = False) -> str:
    """Compile the C restatement with the committed Makefile."""
    src = os.path.join(_HERE, "gempic_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_maxwell1d_table.restype = _dp
        L.orc_maxwell1d_delta_x.restype = C.c_double
        L.orc_hs_j.restype = _dp
        L.orc_boris_field.restype = _dp
        L.orc_pmc1d_evaluate.restype = C.c_double
        L.orc_pmc1d_add_current_update_v.restype = C.c_double
        L.orc_pmc1d_add_current_1d1v.restype = C.c_double
        L.orc_maxwell1d_inner_product.restype = C.c_double
        L.orc_maxwell1d_l2norm_squared.restype = C.c_double
        L.orc_pmc2d_evaluate.restype = C.c_double
        L.orc_mod_julia.restype = C.c_double
        _lib = L
    return _lib


def _p(a: np.ndarray):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"] or a.flags["F_CONTIGUOUS"]
    return a.ctypes.data_as(_dp)


def _f(x):
    return C.c_double(float(x))


def _opaque(kind: int):
    n = lib().orc_sizeof(kind)
    return (C.c_char * n)()


def bsplines_eval_basis(degree: int, offset: float) -> np.ndarray:
    """src/low_level_bsplines.jl:63-80"""
    out = np.zeros(degree + 1)
    lib().orc_bsplines_eval_basis(C.c_int(degree), _f(offset), _p(out))
    return out


def gausslegendre(n: int):
    x = np.zeros(n)
    w = np.zeros(n)
    lib().orc_gausslegendre(C.c_int(n), _p(x), _p(w))
    return x, w


def mod_julia(x: float, y: float) -> float:
    return lib().orc_mod_julia(_f(x), _f(y))


SMOOTHING = {"collocation": 0, "galerkin": 1}


class OneDGrid:
    """src/mesh.jl:52-67 (only the scalar fields the hot path reads)"""

    def __init__(self, xmin, xmax, nx):
        self.xmin, self.xmax, self.nx = float(xmin), float(xmax), int(nx)


class TwoDGrid:
    """src/mesh.jl:17-45"""

    def __init__(self, xmin, xmax, nx, ymin, ymax, ny):
        self.xmin, self.xmax, self.nx = float(xmin), float(xmax), int(nx)
        self.ymin, self.ymax, self.ny = float(ymin), float(ymax), int(ny)


class ParticleGroup:
    """src/particle_group.jl:15-46.  `array` has the reference's shape (D+V+W, N) and
    the reference's memory layout (column-major)."""

    def __init__(self, D, V, n_particles, charge=1.0, mass=1.0, n_weights=1, common_weight=0.0):
        self.dims = (D, V)
        self.n_particles = int(n_particles)
        self.n_weights = n_weights
        self._base = np.zeros((self.n_particles, D + V + n_weights))
        self.array = self._base.T  # (D+V+W, N) view, column-major
        self.charge, self.mass = float(charge), float(mass)
        self.common_weight = 1.0 / n_particles if common_weight == 0.0 else float(common_weight)
        self.q_over_m = self.charge / self.mass
        self._c = _opaque(2)
        lib().orc_pg_init(self._c, _p(self._base), C.c_int64(self.n_particles), C.c_int(D), C.c_int(V),
                          C.c_int(n_weights), _f(charge), _f(mass), _f(self.common_weight))

    def get_charge(self, i):
        D, V = self.dims
        return self.charge * self.array[D + V, i] * self.common_weight


class ParticleMeshCoupling1D:
    """src/particle_mesh_coupling_1d.jl:26-95"""

    def __init__(self, mesh: OneDGrid, n_particles: int, degree: int, smoothing: str):
        if smoothing not in SMOOTHING:
            raise ValueError(f"Smoothing Type {smoothing} not implemented")  # ArgumentError :61
        self._c = _opaque(0)
        rc = lib().orc_pmc1d_init(self._c, _f(mesh.xmin), _f(mesh.xmax), C.c_int(mesh.nx), C.c_int(degree),
                                  C.c_int(SMOOTHING[smoothing]))
        if rc:
            raise ValueError("bad ParticleMeshCoupling1D arguments")
        self.n_dofs = self.n_grid = mesh.nx
        self.degree = degree
        self.smoothing = smoothing
        self.delta_x = (mesh.xmax - mesh.xmin) / mesh.nx
        self.xmin, self.Lx = mesh.xmin, mesh.xmax - mesh.xmin

    def add_charge(self, rho, x, w):
        lib().orc_pmc1d_add_charge(self._c, _p(rho), _f(x), _f(w))

    def add_charge_batch(self, rho, x, w):
        x = np.ascontiguousarray(x, dtype=np.float64)
        w = np.ascontiguousarray(w, dtype=np.float64)
        lib().orc_pmc1d_add_charge_batch(self._c, _p(rho), _p(x), _p(w), C.c_int64(x.size))

    def evaluate(self, x, field):
        return lib().orc_pmc1d_evaluate(self._c, _f(x), _p(field))

    def evaluate_batch(self, x, field):
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = np.zeros_like(x)
        lib().orc_pmc1d_evaluate_batch(self._c, _p(x), _p(field), _p(out), C.c_int64(x.size))
        return out

    def add_current_update_v(self, j, x_old, x_new, w, qm, b, v):
        return lib().orc_pmc1d_add_current_update_v(self._c, _p(j), _f(x_old), _f(x_new), _f(w), _f(qm), _p(b), _f(v))

    def add_current_update_v_1d1v(self, j, x_old, x_new, w, qm, v):
        return lib().orc_pmc1d_add_current_1d1v(self._c, _p(j), _f(x_old), _f(x_new), _f(w), _f(qm), _f(v))


class Maxwell1DFEM:
    """src/maxwell_1d_fem.jl:29-177"""

    def __init__(self, mesh: OneDGrid, degree: int):
        self._c = _opaque(1)
        rc = lib().orc_maxwell1d_init(self._c, _f(mesh.xmin), _f(mesh.xmax), C.c_int(mesh.nx), C.c_int(degree))
        if rc:
            raise ValueError(f"Wrong value of degree = {degree} (1,2 or 3) or odd n")
        self.n_dofs = mesh.nx
        self.s_deg_0, self.s_deg_1 = degree, degree - 1
        self.xmin, self.Lx = mesh.xmin, mesh.xmax - mesh.xmin
        self.delta_x = self.Lx / mesh.nx

    def __del__(self):
        try:
            lib().orc_maxwell1d_free(self._c)
        except Exception:
            pass

    def _table(self, which):
        ptr = lib().orc_maxwell1d_table(self._c, C.c_int(which))
        return np.ctypeslib.as_array(ptr, shape=(self.n_dofs,))

    eig_mass0 = property(lambda s: s._table(0))
    eig_mass1 = property(lambda s: s._table(1))
    eig_weak_ampere = property(lambda s: s._table(2))
    eig_weak_poisson = property(lambda s: s._table(3))
    work = property(lambda s: s._table(4))

    def solve_circulant(self, eigvals, rhs):
        lib().orc_maxwell1d_solve_circulant(self._c, _p(np.ascontiguousarray(eigvals)), _p(rhs))
        return self.work.copy()

    def compute_e_from_rho(self, e, rho):
        lib().orc_maxwell1d_compute_e_from_rho(self._c, _p(e), _p(rho))

    def compute_e_from_j(self, e, current, component):
        if lib().orc_maxwell1d_compute_e_from_j(self._c, _p(e), _p(current), C.c_int(component)):
            raise ValueError(f"Component {component} not implemented")

    def compute_e_from_b(self, field_out, dt, field_in):
        lib().orc_maxwell1d_compute_e_from_b(self._c, _p(field_out), _f(dt), _p(field_in))

    def compute_b_from_e(self, field_out, dt, field_in):
        lib().orc_maxwell1d_compute_b_from_e(self._c, _p(field_out), _f(dt), _p(field_in))

    def inner_product(self, c1, c2, degree):
        return lib().orc_maxwell1d_inner_product(self._c, _p(c1), _p(c2), C.c_int(degree))

    def l2norm_squared(self, c, degree):
        return lib().orc_maxwell1d_l2norm_squared(self._c, _p(c), C.c_int(degree))

    def compute_rhs_from_function(self, coefs, func, degree):
        cb = C.CFUNCTYPE(C.c_double, C.c_double, _vp)(lambda x, _ctx: float(func(x)))
        lib().orc_maxwell1d_compute_rhs_from_function(self._c, _p(coefs), cb, None, C.c_int(degree))

    def l2projection(self, coefs, func, degree):
        cb = C.CFUNCTYPE(C.c_double, C.c_double, _vp)(lambda x, _ctx: float(func(x)))
        if lib().orc_maxwell1d_l2projection(self._c, _p(coefs), cb, None, C.c_int(degree)):
            raise ValueError(f"degree {degree} not available")


class HamiltonianSplitting:
    """src/hamiltonian_splitting.jl:20-108 with the {1,2} operators of
    src/hamiltonian_splitting_1d2v.jl and the {1,1} operators of
    src/hamiltonian_splitting_1d1v.jl.  e_dofs/b_dofs are aliased, j_dofs owned."""

    def __init__(self, D, V, maxwell, ks0, ks1, pg, e_dofs, b_dofs, n_chunks=1):
        assert (D, V) == pg.dims
        self.dims = (D, V)
        self.maxwell_solver, self.kernel_smoother_0, self.kernel_smoother_1 = maxwell, ks0, ks1
        self.particle_group = pg
        self.e_dofs, self.b_dofs = e_dofs, b_dofs
        self._c = _opaque(3)
        rc = lib().orc_hs_init(self._c, maxwell._c, ks0._c, ks1._c, pg._c, _p(e_dofs[0]), _p(e_dofs[1]), _p(b_dofs),
                               C.c_int(n_chunks))
        if rc:
            raise AssertionError("HamiltonianSplitting: n_dofs mismatch or n_particles % n_chunks != 0")
        n = ks0.n_dofs
        self.j_dofs = [np.ctypeslib.as_array(lib().orc_hs_j(self._c, C.c_int(k)), shape=(n,)) for k in (1, 2)]

    def __del__(self):
        try:
            lib().orc_hs_free(self._c)
        except Exception:
            pass

    def operatorHp1(self, dt):
        (lib().orc_hs_operatorHp1 if self.dims == (1, 2) else lib().orc_hs11_operatorHp1)(self._c, _f(dt))

    def operatorHp2(self, dt):
        assert self.dims == (1, 2)
        lib().orc_hs_operatorHp2(self._c, _f(dt))

    def operatorHE(self, dt):
        assert self.dims == (1, 2)
        lib().orc_hs_operatorHE(self._c, _f(dt))

    def operatorHB(self, dt):
        (lib().orc_hs_operatorHB if self.dims == (1, 2) else lib().orc_hs11_operatorHB)(self._c, _f(dt))

    def strang_splitting(self, dt, number_steps):
        fn = lib().orc_hs_strang_splitting if self.dims == (1, 2) else lib().orc_hs11_strang_splitting
        fn(self._c, _f(dt), C.c_int(number_steps))


class HamiltonianSplittingBoris:
    """src/hamiltonian_splitting_boris.jl:23-288"""

    def __init__(self, maxwell, ks0, ks1, pg, e_dofs, b_dofs):
        self.maxwell_solver, self.kernel_smoother_0, self.kernel_smoother_1 = maxwell, ks0, ks1
        self.particle_group = pg
        self.e_dofs, self.b_dofs = e_dofs, b_dofs
        self._c = _opaque(4)
        if lib().orc_boris_init(self._c, maxwell._c, ks0._c, ks1._c, pg._c, _p(e_dofs[0]), _p(e_dofs[1]), _p(b_dofs)):
            raise AssertionError("HamiltonianSplittingBoris: bad arguments")
        n = ks0.n_dofs
        f = [np.ctypeslib.as_array(lib().orc_boris_field(self._c, C.c_int(k)), shape=(n,)) for k in range(5)]
        self.e_dofs_mid = [f[0], f[1]]
        self.b_dofs_mid = f[2]
        self.j_dofs = [f[3], f[4]]

    def __del__(self):
        try:
            lib().orc_boris_free(self._c)
        except Exception:
            pass

    def staggering(self, dt):
        lib().orc_boris_staggering(self._c, _f(dt))

    def strang_splitting(self, dt, number_steps):
        lib().orc_boris_strang_splitting(self._c, _f(dt), C.c_int(number_steps))

    def push_v_epart(self, dt):
        lib().orc_boris_push_v_epart(self._c, _f(dt))

    def push_v_bpart(self, dt):
        lib().orc_boris_push_v_bpart(self._c, _f(dt))

    def push_x_accumulate_j(self, dt):
        lib().orc_boris_push_x_accumulate_j(self._c, _f(dt))


def solve_poisson(efield, pg, ks0, maxwell, rho):
    """src/diagnostics.jl:15-31"""
    lib().orc_solve_poisson(_p(efield), pg._c, ks0._c, maxwell._c, _p(rho))


DIAG_COLUMNS = ("Time", "KineticEnergy", "Momentum1", "Momentum2", "PotentialEnergyE1", "PotentialEnergyE2",
                "PotentialEnergyB3", "Transfer", "VVB", "Poynting", "ErrorPoisson")


def write_step(pg, maxwell, ks0, ks1, time, degree, e_dofs, b_dofs, e_dofs_n, e_poisson):
    """src/diagnostics.jl:186-250 -> the 11 columns of :143-155"""
    out = np.zeros(11)
    lib().orc_write_step(pg._c, maxwell._c, ks0._c, ks1._c, _f(time), C.c_int(degree), _p(e_dofs[0]), _p(e_dofs[1]),
                         _p(b_dofs), _p(e_dofs_n[0]), _p(e_dofs_n[1]), _p(e_poisson), _p(out))
    return out


class ParticleMeshCoupling2D:
    """src/particle_mesh_coupling_2d.jl:12-45"""

    def __init__(self, grid: TwoDGrid, degree: int, smoothing: str):
        self._c = _opaque(5)
        rc = lib().orc_pmc2d_init(self._c, _f(grid.xmin), _f(grid.xmax), C.c_int(grid.nx), _f(grid.ymin),
                                  _f(grid.ymax), C.c_int(grid.ny), C.c_int(degree), C.c_int(SMOOTHING[smoothing]))
        if rc:
            raise ValueError("bad ParticleMeshCoupling2D arguments")
        self.grid, self.degree = grid, degree

    def shape_indices(self, xp, yp):
        ix, iy = C.c_long(), C.c_long()
        lib().orc_pmc2d_shape_indices(self._c, _f(xp), _f(yp), C.byref(ix), C.byref(iy))
        return ix.value, iy.value

    def add_charge(self, rho, xp, yp, wp):
        lib().orc_pmc2d_add_charge(self._c, _p(rho), _f(xp), _f(yp), _f(wp))

    def evaluate(self, xp, yp, field):
        return lib().orc_pmc2d_evaluate(self._c, _f(xp), _f(yp), _p(field))

    def evaluate_multiple(self, xp, yp, f1, f2):
        out = np.zeros(2)
        lib().orc_pmc2d_evaluate_multiple(self._c, _f(xp), _f(yp), _p(f1), _p(f2), _p(out))
        return out[0], out[1]

    def add_charge_batch(self, rho, x, y, w):
        x, y, w = (np.ascontiguousarray(a, dtype=np.float64) for a in (x, y, w))
        lib().orc_pmc2d_add_charge_batch(self._c, _p(rho), _p(x), _p(y), _p(w), C.c_int64(x.size))

    def evaluate_batch(self, x, y, field):
        x, y = (np.ascontiguousarray(a, dtype=np.float64) for a in (x, y))
        out = np.zeros_like(x)
        lib().orc_pmc2d_evaluate_batch(self._c, _p(x), _p(y), _p(field), _p(out), C.c_int64(x.size))
        return out


def max_threads() -> int:
    return lib().orc_max_threads()


def set_threads(n: int) -> int:
    """omp_set_num_threads(n); returns the team size now in force (1 without OpenMP)"""
    lib().orc_set_threads(C.c_int(int(n)))
    return lib().orc_max_threads()
