"""Host-side mirror of the reference's operator interface (GEMPIC.jl/src) for the hot path.

Every class keeps the reference's name and constructor arguments; every method is one call
into libgempic_b200.so.  Host numpy arrays are only ever *borrowed* for the duration of a
call; particle data lives on the GPU between calls.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import FUNC1D, FUNC2D, c_handle, check, dptr

__all__ = [
    "OneDGrid", "TwoDGrid", "ParticleGroup", "ParticleMeshCoupling1D", "ParticleMeshCoupling2D", "Maxwell1DFEM", "TwoDMaxwell", "HamiltonianSplitting2D3V",
    "HamiltonianSplitting", "HamiltonianSplittingBoris", "TimeHistoryDiagnostics", "strang_splitting",
    "staggering", "operatorHp1", "operatorHp2", "operatorHE", "operatorHB", "solve_poisson", "write_step",
    "add_charge", "evaluate", "add_current_update_v", "add_charge_pp", "evaluate_pp", "add_current_update_v_pp", "PPField", "compute_e_from_rho", "compute_e_from_j", "compute_e_from_b",
    "compute_b_from_e", "inner_product", "l2norm_squared", "l2projection", "compute_rhs_from_function",
    "ParticleSampler", "CosSumGaussian", "SumCosGaussian", "LandauDamping", "sample",
    "synchronize", "launch_count", "stream_ptr", "device_info", "set_option", "DIAG_COLUMNS", "save", "load_particles",
]

SMOOTHING = {"collocation": 0, "galerkin": 1}
OP_HP1, OP_HP2, OP_HE, OP_HB = 1, 2, 3, 4


def _L():
    _lib.init()
    return _lib.load()


def _f(x):
    return C.c_double(float(x))


def _vec(a, n, name):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if a.shape != (n,):
        raise ValueError(f"{name} must have shape ({n},), got {a.shape}")
    return a


def synchronize():
    check(_L().gempic_synchronize())


def launch_count(reset=False) -> int:
    return int(_L().gempic_launch_count(1 if reset else 0))


def stream_ptr() -> int:
    return int(_L().gempic_stream() or 0)


def device_info():
    sm, free, total = C.c_int(), C.c_int64(), C.c_int64()
    check(_L().gempic_device_info(C.byref(sm), C.byref(free), C.byref(total)))
    return {"sm_count": sm.value, "free_bytes": free.value, "total_bytes": total.value}


def set_option(name: str, value: int):
    check(_L().gempic_set_option(name.encode(), C.c_int64(value)))


class _Handle:
    _destroy = None

    def __init__(self):
        self._h = c_handle(0)

    @property
    def handle(self):
        return self._h

    def close(self):
        if self._h.value and self._destroy and _lib._initialised:
            getattr(_lib.load(), self._destroy)(self._h)
        self._h = c_handle(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class OneDGrid:
    """src/mesh.jl:52-67 -- only xmin, xmax, nx are read by the hot path."""

    def __init__(self, xmin, xmax, nx):
        self.xmin, self.xmax, self.nx = float(xmin), float(xmax), int(nx)
        self.dimx = self.xmax - self.xmin


class TwoDGrid:
    """src/mesh.jl:17-45"""

    def __init__(self, xmin, xmax, nx, ymin, ymax, ny):
        self.xmin, self.xmax, self.nx = float(xmin), float(xmax), int(nx)
        self.ymin, self.ymax, self.ny = float(ymin), float(ymax), int(ny)
        self.dx = (self.xmax - self.xmin) / self.nx
        self.dy = (self.ymax - self.ymin) / self.ny


class ParticleGroup(_Handle):
    """ParticleGroup{D,V}(n_particles; charge, mass, n_weights, common_weight)
    (src/particle_group.jl:15-46).  Device-resident fp64 SoA; `array` is a lazily
    synchronised host view with the reference's shape (D+V+W, N), column-major."""

    _destroy = "gempic_pg_destroy"

    def __init__(self, D, V, n_particles, charge=1.0, mass=1.0, n_weights=1, common_weight=0.0):
        super().__init__()
        self.dims = (int(D), int(V))
        self.n_particles = int(n_particles)
        self.n_weights = int(n_weights)
        self.charge, self.mass = float(charge), float(mass)
        self.common_weight = 1.0 / n_particles if common_weight == 0.0 else float(common_weight)
        self.q_over_m = self.charge / self.mass
        check(_L().gempic_pg_create(C.c_int(D), C.c_int(V), C.c_int(n_weights), C.c_int64(n_particles), _f(charge),
                                    _f(mass), _f(self.common_weight), C.byref(self._h)))
        self._host = None          # numpy (N, rows) C-order == (rows, N) column-major
        self._host_newer = False   # host copy may have been modified by the user
        self._dev_newer = False    # device copy has advanced since the last download

    # -- host mirror ------------------------------------------------------------------------
    @property
    def rows(self):
        return sum(self.dims) + self.n_weights

    @property
    def array(self):
        """Host view (rows, N). Reading it downloads if the device is ahead; because the caller may write into it, the
        next device operation uploads it again.  The view is only coherent until the next device operation: fetch
        `pg.array` again afterwards (a numpy view kept across a device call is neither refreshed nor uploaded; the Julia
        shim's MirrorArray tracks this, numpy cannot)."""
        a = self._mirror()
        self._host_newer = True
        return a

    def _mirror(self):
        """the host mirror for READING: synchronised with the device, not marked as written (get_x / get_v /
        get_charge / get_mass go through here, so a read-only diagnostic loop never re-uploads the particles)"""
        if self._host is None:
            self._host = np.zeros((self.n_particles, self.rows))
            self._dev_newer = True
        if self._dev_newer:
            check(_L().gempic_pg_download(self._h, dptr(self._host)))
            self._dev_newer = False
        return self._host.T

    def to_host(self):
        """Read-only snapshot (rows, N); does not mark the host copy as modified."""
        self._flush()
        out = np.zeros((self.n_particles, self.rows))
        check(_L().gempic_pg_download(self._h, dptr(out)))
        return out.T

    def upload(self, array):
        """array: (rows, N) like the reference's `pg.array`."""
        a = np.asarray(array, dtype=np.float64)
        if a.shape != (self.rows, self.n_particles):
            raise ValueError(f"expected shape {(self.rows, self.n_particles)}, got {a.shape}")
        base = np.ascontiguousarray(a.T)
        check(_L().gempic_pg_upload(self._h, dptr(base)))
        self._host, self._host_newer, self._dev_newer = None, False, False

    def _flush(self):
        if self._host_newer and self._host is not None:
            check(_L().gempic_pg_upload(self._h, dptr(self._host)))
        self._host_newer = False

    def _touched(self):
        self._dev_newer = True

    # -- reference accessors (particle_group.jl:53-150) --------------------------------------
    def get_x(self, i):
        return self._mirror()[0:self.dims[0], i].copy()

    def get_v(self, i):
        D, V = self.dims
        return self._mirror()[D:D + V, i].copy()

    def get_charge(self, i, i_wi=1):
        D, V = self.dims
        return self.charge * self._mirror()[D + V + i_wi - 1, i] * self.common_weight

    def get_mass(self, i, i_wi=1):
        D, V = self.dims
        return self.mass * self._mirror()[D + V + i_wi - 1, i] * self.common_weight

    def set_x(self, i, x):
        x = np.atleast_1d(np.asarray(x, dtype=np.float64))
        self.array[0:len(x), i] = x

    def set_v(self, i, v):
        v = np.atleast_1d(np.asarray(v, dtype=np.float64))
        D = self.dims[0]
        self.array[D:D + len(v), i] = v

    def set_weights(self, i, w):
        w = np.atleast_1d(np.asarray(w, dtype=np.float64))
        D, V = self.dims
        self.array[D + V:D + V + len(w), i] = w

    # -- device-side extras -------------------------------------------------------------------
    def sort(self, pmc):
        """cell sort on the mesh of a ParticleMeshCoupling1D (D = 1) or a TwoDMaxwell (D = 2)"""
        self._flush()
        if self.dims[0] == 2:
            check(_L().gempic_pg_sort2d(self._h, pmc.handle))
        else:
            check(_L().gempic_pg_sort(self._h, pmc.handle))
        self._touched()

    def sample(self, kind, xmin, L, alpha=0.0, k=1.0, sigma=(1.0, 1.0, 1.0), seed=1234, first_index=0):
        """kind: 'uniform' (x~U, v~N(0,sigma)) or 'landau' (x by inverse CDF of 1+alpha cos(kx))."""
        sig = (C.c_double * 3)(*(list(sigma) + [1.0, 1.0, 1.0])[:3])
        check(_L().gempic_pg_sample(self._h, C.c_int({"uniform": 0, "landau": 1}[kind]), _f(xmin), _f(L), _f(alpha),
                                    _f(k), sig, C.c_uint64(seed), C.c_int64(first_index)))
        self._host, self._host_newer, self._dev_newer = None, False, True

    def row_ptr(self, row) -> int:
        p = C.POINTER(C.c_double)()
        check(_L().gempic_pg_row_ptr(self._h, C.c_int(row), C.byref(p)))
        return C.cast(p, C.c_void_p).value

    def set_row_device(self, row, dev_ptr: int):
        check(_L().gempic_pg_set_row_device(self._h, C.c_int(row), C.cast(C.c_void_p(dev_ptr), _lib.c_dp)))
        self._host, self._host_newer, self._dev_newer = None, False, True


class PPField:
    """what b_to_pp (src/splinepp.jl:241-261, :392) returns here: the spline dofs behind an opaque name -- the device
    kernels build the per-cell polynomials themselves (cell_poly, csrc/ops1d.cuh)"""

    def __init__(self, dofs):
        self.dofs = dofs


class ParticleMeshCoupling1D(_Handle):
    """ParticleMeshCoupling1D(mesh, no_particles, spline_degree, smoothing_type)
    (src/particle_mesh_coupling_1d.jl:26-95)."""

    _destroy = "gempic_pmc1d_destroy"

    def __init__(self, mesh: OneDGrid, no_particles: int, spline_degree: int, smoothing_type: str):
        super().__init__()
        if smoothing_type not in SMOOTHING:
            raise _lib.ArgumentError(1, f"Smoothing Type {smoothing_type} not implemented for kernel_smoother_spline_1d.")
        check(_L().gempic_pmc1d_create(_f(mesh.xmin), _f(mesh.xmax), C.c_int(mesh.nx), C.c_int64(no_particles),
                                       C.c_int(spline_degree), C.c_int(SMOOTHING[smoothing_type]), C.byref(self._h)))
        self.dims = 1
        self.n_grid = self.n_dofs = mesh.nx
        self.no_particles = no_particles
        self.spline_degree = spline_degree
        self.n_span = spline_degree + 1
        self.xmin, self.Lx = mesh.xmin, mesh.xmax - mesh.xmin
        self.delta_x = self.Lx / mesh.nx
        self.scaling = 1.0 / self.delta_x if smoothing_type == "collocation" else 1.0

    def add_charge(self, rho_dofs, position, marker_charge):
        """add_charge!(rho_dofs, p, position, marker_charge) for scalars or equal-length vectors (:261-280)."""
        x = np.atleast_1d(np.asarray(position, dtype=np.float64)).copy()
        w = np.broadcast_to(np.asarray(marker_charge, dtype=np.float64), x.shape).copy()
        check(_L().gempic_pmc1d_add_charge(self._h, dptr(x), dptr(w), C.c_int64(x.size), dptr(rho_dofs)))

    def evaluate(self, position, field_dofs):
        """evaluate(p, position, field_dofs) (:438-453); vector form src/diagnostics.jl:259-268."""
        scalar = np.ndim(position) == 0
        x = np.atleast_1d(np.asarray(position, dtype=np.float64)).copy()
        out = np.zeros_like(x)
        check(_L().gempic_pmc1d_evaluate(self._h, dptr(x), C.c_int64(x.size), dptr(_vec(field_dofs, self.n_dofs, "field_dofs")),
                                         dptr(out)))
        return float(out[0]) if scalar else out

    def add_current_update_v(self, j_dofs, position_old, position_new, marker_charge, qoverm, bfield_dofs=None, vi=None):
        """add_current_update_v!(j_dofs, p, x_old, x_new, w, qoverm, bfield_dofs, vi) -> vi (:296-376);
        without bfield_dofs the 1d1v form (:471-529)."""
        scalar = np.ndim(position_old) == 0
        xo = np.atleast_1d(np.asarray(position_old, dtype=np.float64)).copy()
        xn = np.atleast_1d(np.asarray(position_new, dtype=np.float64)).copy()
        w = np.broadcast_to(np.asarray(marker_charge, dtype=np.float64), xo.shape).copy()
        if bfield_dofs is None:
            check(_L().gempic_pmc1d_add_current(self._h, dptr(xo), dptr(xn), dptr(w), C.c_int64(xo.size), dptr(j_dofs)))
            return vi
        v = np.broadcast_to(np.asarray(vi, dtype=np.float64), xo.shape).copy()
        check(_L().gempic_pmc1d_add_current_update_v(self._h, dptr(xo), dptr(xn), dptr(w), _f(qoverm),
                                                     dptr(_vec(bfield_dofs, self.n_dofs, "bfield_dofs")), dptr(v),
                                                     C.c_int64(xo.size), dptr(j_dofs)))
        return float(v[0]) if scalar else v

    # the `_pp` forms of the reference (:106-250) compute the same functions through the piecewise-polynomial tables; the
    # kernels here always evaluate in pp form, so they are aliases.  b_to_pp returns an opaque handle on the dofs.
    add_charge_pp = add_charge
    add_current_update_v_pp = add_current_update_v

    def b_to_pp(self, field_dofs):
        return PPField(_vec(field_dofs, self.n_dofs, "field_dofs").copy())

    def evaluate_pp(self, position, field_dofs_pp):
        return self.evaluate(position, field_dofs_pp.dofs)

    def add_charge_pg(self, rho_dofs, pg: ParticleGroup):
        """rho += sum_i add_charge!(rho, p, x_i, get_charge(pg, i)) on the device-resident group."""
        pg._flush()
        check(_L().gempic_pmc1d_add_charge_pg(self._h, pg.handle, dptr(rho_dofs)))

    def evaluate_pg(self, pg: ParticleGroup, field_dofs):
        pg._flush()
        out = np.zeros(pg.n_particles)
        check(_L().gempic_pmc1d_evaluate_pg(self._h, pg.handle, dptr(_vec(field_dofs, self.n_dofs, "field_dofs")), dptr(out)))
        return out


class ParticleMeshCoupling2D(_Handle):
    """ParticleMeshCoupling2D(pg, grid, degree, smoothing_type) (src/particle_mesh_coupling_2d.jl:12-45)."""

    _destroy = "gempic_pmc2d_destroy"

    def __init__(self, pg, grid: TwoDGrid, degree: int, smoothing_type: str):
        super().__init__()
        if smoothing_type not in SMOOTHING:
            raise _lib.ArgumentError(1, f"Smoothing Type {smoothing_type} not implemented for kernel_smoother_spline_2d.")
        if pg is not None and pg.dims[0] != 2:
            raise _lib.AssertionFailed(2, "ParticleMeshCoupling2D needs a ParticleGroup{2,V}")
        check(_L().gempic_pmc2d_create(_f(grid.xmin), _f(grid.xmax), C.c_int(grid.nx), _f(grid.ymin), _f(grid.ymax),
                                       C.c_int(grid.ny), C.c_int(degree), C.c_int(SMOOTHING[smoothing_type]),
                                       C.byref(self._h)))
        self.grid, self.degree, self.n_span = grid, degree, degree + 1
        self.npart = pg.n_particles if pg is not None else 0
        self.n_dofs = grid.nx * grid.ny

    def add_charge(self, rho_dofs, xp, yp, wp):
        x = np.atleast_1d(np.asarray(xp, dtype=np.float64)).copy()
        y = np.broadcast_to(np.asarray(yp, dtype=np.float64), x.shape).copy()
        w = np.broadcast_to(np.asarray(wp, dtype=np.float64), x.shape).copy()
        check(_L().gempic_pmc2d_add_charge(self._h, dptr(x), dptr(y), dptr(w), C.c_int64(x.size), dptr(rho_dofs)))

    def evaluate(self, xp, yp, field_dofs):
        scalar = np.ndim(xp) == 0
        x = np.atleast_1d(np.asarray(xp, dtype=np.float64)).copy()
        y = np.broadcast_to(np.asarray(yp, dtype=np.float64), x.shape).copy()
        out = np.zeros_like(x)
        check(_L().gempic_pmc2d_evaluate(self._h, dptr(x), dptr(y), C.c_int64(x.size), dptr(_vec(field_dofs, self.n_dofs, "field_dofs")),
                                         dptr(out)))
        return float(out[0]) if scalar else out

    # `_pp` forms (src/particle_mesh_coupling_2d.jl:107-188): same functions through the pp tables (floor instead of ceil
    # cell index, identical values by continuity of the splines)
    add_charge_pp = add_charge

    def b_to_pp(self, field_dofs):
        return PPField(_vec(field_dofs, self.n_dofs, "field_dofs").copy())

    def evaluate_pp(self, xp, yp, pp):
        return self.evaluate(xp, yp, pp.dofs)

    def evaluate_multiple(self, position, field_dofs):
        xp, yp = position
        scalar = np.ndim(xp) == 0
        x = np.atleast_1d(np.asarray(xp, dtype=np.float64)).copy()
        y = np.broadcast_to(np.asarray(yp, dtype=np.float64), x.shape).copy()
        o1, o2 = np.zeros_like(x), np.zeros_like(x)
        check(_L().gempic_pmc2d_evaluate_multiple(self._h, dptr(x), dptr(y), C.c_int64(x.size),
                                                  dptr(_vec(field_dofs[0], self.n_dofs, "field_dofs[1]")),
                                                  dptr(_vec(field_dofs[1], self.n_dofs, "field_dofs[2]")), dptr(o1), dptr(o2)))
        return (float(o1[0]), float(o2[0])) if scalar else (o1, o2)

    def add_charge_pg(self, rho_dofs, pg: ParticleGroup):
        pg._flush()
        check(_L().gempic_pmc2d_add_charge_pg(self._h, pg.handle, dptr(rho_dofs)))

    def evaluate_pg(self, pg: ParticleGroup, field_dofs):
        pg._flush()
        out = np.zeros(pg.n_particles)
        check(_L().gempic_pmc2d_evaluate_pg(self._h, pg.handle, dptr(_vec(field_dofs, self.n_dofs, "field_dofs")), dptr(out)))
        return out


class Maxwell1DFEM(_Handle):
    """Maxwell1DFEM(mesh, degree) (src/maxwell_1d_fem.jl:29-177)."""

    _destroy = "gempic_maxwell1d_destroy"

    def __init__(self, mesh: OneDGrid, degree: int):
        super().__init__()
        check(_L().gempic_maxwell1d_create(_f(mesh.xmin), _f(mesh.xmax), C.c_int(mesh.nx), C.c_int(degree), C.byref(self._h)))
        self.xmin, self.Lx = mesh.xmin, mesh.xmax - mesh.xmin
        self.n_dofs = mesh.nx
        self.delta_x = self.Lx / mesh.nx
        self.s_deg_0, self.s_deg_1 = degree, degree - 1

    def _table(self, which):
        out = np.zeros(self.n_dofs)
        check(_L().gempic_maxwell1d_get_table(self._h, C.c_int(which), dptr(out)))
        return out

    eig_mass0 = property(lambda s: s._table(0))
    eig_mass1 = property(lambda s: s._table(1))
    eig_weak_ampere = property(lambda s: s._table(2))
    eig_weak_poisson = property(lambda s: s._table(3))

    def compute_e_from_rho(self, e, rho):
        check(_L().gempic_maxwell1d_compute_e_from_rho(self._h, dptr(e), dptr(_vec(rho, self.n_dofs, "rho"))))

    def compute_e_from_j(self, e, current, component):
        check(_L().gempic_maxwell1d_compute_e_from_j(self._h, dptr(e), dptr(_vec(current, self.n_dofs, "current")), C.c_int(component)))

    def compute_e_from_b(self, field_out, delta_t, field_in):
        check(_L().gempic_maxwell1d_compute_e_from_b(self._h, dptr(field_out), _f(delta_t), dptr(_vec(field_in, self.n_dofs, "field_in"))))

    def compute_b_from_e(self, field_out, delta_t, field_in):
        check(_L().gempic_maxwell1d_compute_b_from_e(self._h, dptr(field_out), _f(delta_t), dptr(_vec(field_in, self.n_dofs, "field_in"))))

    def inner_product(self, coefs1_dofs, coefs2_dofs, degree):
        out = C.c_double()
        check(_L().gempic_maxwell1d_inner_product(self._h, dptr(_vec(coefs1_dofs, self.n_dofs, "coefs1")),
                                                  dptr(_vec(coefs2_dofs, self.n_dofs, "coefs2")), C.c_int(degree), C.byref(out)))
        return out.value

    def l2norm_squared(self, coefs_dofs, degree):
        return self.inner_product(coefs_dofs, coefs_dofs, degree)

    def compute_rhs_from_function(self, coefs_dofs, func, degree):
        cb = FUNC1D(lambda x, _ctx: float(func(x)))
        check(_L().gempic_maxwell1d_compute_rhs_from_function(self._h, dptr(coefs_dofs), cb, None, C.c_int(degree)))

    def l2projection(self, coefs_dofs, func, degree):
        cb = FUNC1D(lambda x, _ctx: float(func(x)))
        check(_L().gempic_maxwell1d_l2projection(self._h, dptr(coefs_dofs), cb, None, C.c_int(degree)))


class TwoDMaxwell(_Handle):
    """TwoDMaxwell(mesh, degree) (src/maxwell_2d_fem.jl:11-87); field vectors are lists of three flat
    nx*ny arrays, x fastest, like the reference's `efield`, `bfield`."""

    _destroy = "gempic_maxwell2d_destroy"

    def __init__(self, mesh: TwoDGrid, degree: int):
        super().__init__()
        check(_L().gempic_maxwell2d_create(_f(mesh.xmin), _f(mesh.xmax), C.c_int(mesh.nx), _f(mesh.ymin), _f(mesh.ymax),
                                           C.c_int(mesh.ny), C.c_int(degree), C.byref(self._h)))
        self.mesh = mesh
        self.nx, self.ny = mesh.nx, mesh.ny
        self.n_dofs = mesh.nx * mesh.ny
        self.s_deg_0, self.s_deg_1 = degree, degree - 1

    def _table(self, which, axis):
        out, cnt = np.zeros(max(self.nx, self.ny) + 8), C.c_int()
        check(_L().gempic_maxwell2d_get_table(self._h, C.c_int(which), C.c_int(axis), dptr(out), C.byref(cnt)))
        return out[:cnt.value].copy()

    mass_line_0 = property(lambda s: [s._table(0, 0), s._table(0, 1)])
    mass_line_1 = property(lambda s: [s._table(1, 0), s._table(1, 1)])

    def compute_e_from_rho(self, efield, rho):
        check(_L().gempic_maxwell2d_compute_e_from_rho(self._h, dptr(efield[0]), dptr(efield[1]), dptr(_vec(rho, self.n_dofs, "rho"))))

    def compute_e_from_b(self, e, dt, b):
        v = [_vec(x, self.n_dofs, "b") for x in b]
        check(_L().gempic_maxwell2d_compute_e_from_b(self._h, dptr(e[0]), dptr(e[1]), dptr(e[2]), _f(dt), dptr(v[0]), dptr(v[1]), dptr(v[2])))

    def compute_b_from_e(self, b, dt, e):
        v = [_vec(x, self.n_dofs, "e") for x in e]
        check(_L().gempic_maxwell2d_compute_b_from_e(self._h, dptr(b[0]), dptr(b[1]), dptr(b[2]), _f(dt), dptr(v[0]), dptr(v[1]), dptr(v[2])))

    def compute_e_from_j(self, e, current, component):
        check(_L().gempic_maxwell2d_compute_e_from_j(self._h, dptr(e), dptr(_vec(current, self.n_dofs, "current")), C.c_int(component)))

    def compute_rho_from_e(self, rho, efield):
        v = [_vec(x, self.n_dofs, "efield") for x in efield]
        check(_L().gempic_maxwell2d_compute_rho_from_e(self._h, dptr(rho), dptr(v[0]), dptr(v[1]), dptr(v[2])))

    def inner_product(self, coefs1_dofs, coefs2_dofs, component, form):
        out = C.c_double()
        check(_L().gempic_maxwell2d_inner_product(self._h, dptr(_vec(coefs1_dofs, self.n_dofs, "coefs1")),
                                                  dptr(_vec(coefs2_dofs, self.n_dofs, "coefs2")), C.c_int(component),
                                                  C.c_int(form), C.byref(out)))
        return out.value

    def solve_mass(self, rhs, component, form):
        """solve(solver.inv_mass_1[component] | inv_mass_2[component], rhs)"""
        out = np.zeros(self.n_dofs)
        check(_L().gempic_maxwell2d_solve_mass(self._h, dptr(out), dptr(_vec(rhs, self.n_dofs, "rhs")), C.c_int(component), C.c_int(form)))
        return out

    def multiply_mass(self, c_in, component, form):
        """multiply_mass_2dkron! with the mass lines of (component, form)"""
        out = np.zeros(self.n_dofs)
        check(_L().gempic_maxwell2d_multiply_mass(self._h, dptr(out), dptr(_vec(c_in, self.n_dofs, "c_in")), C.c_int(component), C.c_int(form)))
        return out

    def compute_rhs_from_function(self, func, component, form):
        out = np.zeros(self.n_dofs)
        cb = FUNC2D(lambda x, y, _ctx: float(func(x, y)))
        check(_L().gempic_maxwell2d_compute_rhs_from_function(self._h, dptr(out), cb, None, C.c_int(component), C.c_int(form)))
        return out

    def l2projection(self, func, component, form):
        out = np.zeros(self.n_dofs)
        cb = FUNC2D(lambda x, y, _ctx: float(func(x, y)))
        check(_L().gempic_maxwell2d_l2projection(self._h, dptr(out), cb, None, C.c_int(component), C.c_int(form)))
        return out


class HamiltonianSplitting(_Handle):
    """HamiltonianSplitting{D,V}(maxwell_solver, kernel_smoother_0, kernel_smoother_1, particle_group,
    e_dofs, b_dofs) (src/hamiltonian_splitting.jl:20-86).  `e_dofs` (list of two arrays) and `b_dofs`
    are the caller's arrays: like the reference, every operator reads them on entry and leaves the
    updated values in them (the drop-in, host-buffer path).  `resident=True` switches to the
    device-resident path: fields are uploaded once and only copied back by `sync_fields()`.
    `fuse=True` (the default of the library and of every front end) runs the particle passes of
    strang_splitting! fused (same trajectory); `fuse=False` runs one pass per operator."""

    _destroy = "gempic_hs_destroy"

    def __init__(self, D, V, maxwell_solver, kernel_smoother_0, kernel_smoother_1, particle_group, e_dofs, b_dofs,
                 resident=False, fuse=True):
        super().__init__()
        self.dims = (D, V)
        self.maxwell_solver = maxwell_solver
        self.kernel_smoother_0, self.kernel_smoother_1 = kernel_smoother_0, kernel_smoother_1
        self.particle_group = particle_group
        check(_L().gempic_hs_create(C.c_int(D), C.c_int(V), maxwell_solver.handle, kernel_smoother_0.handle,
                                    kernel_smoother_1.handle, particle_group.handle, C.byref(self._h)))
        n = kernel_smoother_0.n_dofs
        self.e_dofs = [_check_alias(e_dofs[0], n), _check_alias(e_dofs[1], n)]
        self.b_dofs = _check_alias(b_dofs, n)
        self._j = [np.zeros(n), np.zeros(n)]
        self.Lx, self.x_min = maxwell_solver.Lx, maxwell_solver.xmin
        self.delta_x = self.Lx / n
        self.resident = resident
        if not fuse:
            self.set_fusion(False)
        if resident:
            self.upload_fields()

    @property
    def j_dofs(self):
        """`j_dofs` of the reference struct (hamiltonian_splitting.jl:51): scratch owned by the splitting object, which
        no caller of the reference reads.  It lives on the device and is read back on access; after a fused
        strang_splitting! that access rebuilds j_dofs[2] from the particles (csrc/hs1d.cu, materialise_j2)."""
        check(_L().gempic_hs_get_fields(self._h, None, None, None, dptr(self._j[0]), dptr(self._j[1])))
        return self._j

    def upload_fields(self):
        check(_L().gempic_hs_set_fields(self._h, dptr(self.e_dofs[0]), dptr(self.e_dofs[1]), dptr(self.b_dofs)))

    def sync_fields(self):
        check(_L().gempic_hs_get_fields(self._h, dptr(self.e_dofs[0]), dptr(self.e_dofs[1]), dptr(self.b_dofs), None, None))

    def set_fusion(self, fuse: bool):
        check(_L().gempic_hs_set_fusion(self._h, C.c_int(1 if fuse else 0)))

    def _op(self, op, dt):
        pg = self.particle_group
        pg._flush()
        if self.resident:
            check(_L().gempic_hs_operator(self._h, C.c_int(op), _f(dt)))
        else:
            check(_L().gempic_hs_operator_host(self._h, C.c_int(op), _f(dt), dptr(self.e_dofs[0]), dptr(self.e_dofs[1]),
                                               dptr(self.b_dofs), None, None))
        pg._touched()

    def operatorHp1(self, dt):
        self._op(OP_HP1, dt)

    def operatorHp2(self, dt):
        self._op(OP_HP2, dt)

    def operatorHE(self, dt):
        self._op(OP_HE, dt)

    def operatorHB(self, dt):
        self._op(OP_HB, dt)

    def strang_splitting(self, dt, number_steps):
        pg = self.particle_group
        pg._flush()
        if self.resident:
            check(_L().gempic_hs_strang_splitting(self._h, _f(dt), C.c_int64(number_steps)))
        else:
            check(_L().gempic_hs_strang_splitting_host(self._h, _f(dt), C.c_int64(number_steps), dptr(self.e_dofs[0]),
                                                       dptr(self.e_dofs[1]), dptr(self.b_dofs), None, None))
        pg._touched()


class HamiltonianSplitting2D3V(_Handle):
    """HamiltonianSplitting{2,3}(maxwell_solver::TwoDMaxwell, particle_group::ParticleGroup{2,3}, e_dofs, b_dofs):
    the integrator the reference leaves empty (src/hamiltonian_splitting_2d3v.jl), with the field names and
    operator methods of src/hamiltonian_splitting.jl:20-108.  e_dofs / b_dofs are lists of three aliased
    flat nx*ny arrays."""

    _destroy = "gempic_hs2d_destroy"
    OP_HP3 = 5

    def __init__(self, maxwell_solver, particle_group, e_dofs, b_dofs, resident=False):
        super().__init__()
        self.maxwell_solver, self.particle_group = maxwell_solver, particle_group
        check(_L().gempic_hs2d_create(maxwell_solver.handle, particle_group.handle, C.byref(self._h)))
        n = maxwell_solver.n_dofs
        self.n = n
        self.e_dofs = [_check_alias(a, n) for a in e_dofs]
        self.b_dofs = [_check_alias(a, n) for a in b_dofs]
        assert len(self.e_dofs) == 3 and len(self.b_dofs) == 3
        self.resident = resident
        if resident:
            self.upload_fields()

    def _fp(self):
        return [dptr(a) for a in self.e_dofs + self.b_dofs]

    def upload_fields(self):
        check(_L().gempic_hs2d_set_fields(self._h, *self._fp()))

    def sync_fields(self):
        check(_L().gempic_hs2d_get_fields(self._h, *self._fp(), None, None, None))

    @property
    def j_dofs(self):
        out = [np.zeros(self.n) for _ in range(3)]
        check(_L().gempic_hs2d_get_fields(self._h, None, None, None, None, None, None, dptr(out[0]), dptr(out[1]), dptr(out[2])))
        return out

    def set_sort_interval(self, interval: int):
        check(_L().gempic_hs2d_set_sort_interval(self._h, C.c_int(interval)))

    def set_fusion(self, fuse):
        """True / 2: fused passes + sorted fast path (default); 1: fused [HE,Hp3] tile pass only; False / 0: per operator"""
        level = 2 if fuse is True else int(fuse)
        check(_L().gempic_hs2d_set_fusion(self._h, C.c_int(level)))

    def _op(self, op, dt):
        pg = self.particle_group
        pg._flush()
        if self.resident:
            check(_L().gempic_hs2d_operator(self._h, C.c_int(op), _f(dt)))
        else:
            check(_L().gempic_hs2d_operator_host(self._h, C.c_int(op), _f(dt), *self._fp()))
        pg._touched()

    def operatorHp1(self, dt):
        self._op(OP_HP1, dt)

    def operatorHp2(self, dt):
        self._op(OP_HP2, dt)

    def operatorHp3(self, dt):
        self._op(self.OP_HP3, dt)

    def operatorHE(self, dt):
        self._op(OP_HE, dt)

    def operatorHB(self, dt):
        self._op(OP_HB, dt)

    def strang_splitting(self, dt, number_steps):
        pg = self.particle_group
        pg._flush()
        if self.resident:
            check(_L().gempic_hs2d_strang_splitting(self._h, _f(dt), C.c_int64(number_steps)))
        else:
            check(_L().gempic_hs2d_strang_splitting_host(self._h, _f(dt), C.c_int64(number_steps), *self._fp()))
        pg._touched()

    # -- diagnostics ------------------------------------------------------------------------------
    def charge_density(self):
        self.particle_group._flush()
        rho = np.zeros(self.n)
        check(_L().gempic_hs2d_charge_density(self._h, dptr(rho)))
        return rho

    def gauss_residual(self):
        """compute_rho_from_e!(E) - rho(particles); constant in time (discrete Gauss law)"""
        if self.resident:
            self.sync_fields()
        r = np.zeros(self.n)
        self.maxwell_solver.compute_rho_from_e(r, self.e_dofs)
        return r - self.charge_density()

    def moments(self):
        self.particle_group._flush()
        out = np.zeros(4)
        check(_L().gempic_hs2d_moments(self._h, dptr(out)))
        return out

    def energies(self):
        """(kinetic, electric, magnetic) like TimeHistoryDiagnostics' KineticEnergy / PotentialEnergy columns"""
        if self.resident:
            self.sync_fields()
        mx, pg = self.maxwell_solver, self.particle_group
        kin = 0.5 * pg.mass * pg.common_weight * self.moments()[0]
        ee = 0.5 * sum(mx.inner_product(self.e_dofs[c], self.e_dofs[c], c + 1, 1) for c in range(3))
        eb = 0.5 * sum(mx.inner_product(self.b_dofs[c], self.b_dofs[c], c + 1, 2) for c in range(3))
        return kin, ee, eb


def _check_alias(a, n):
    if not (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.shape == (n,) and a.flags["C_CONTIGUOUS"]):
        raise ValueError(f"field dof vectors must be contiguous float64 arrays of shape ({n},) (they are aliased, "
                         "like in the reference)")
    return a


class HamiltonianSplittingBoris(_Handle):
    """HamiltonianSplittingBoris(maxwell_solver, kernel_smoother_0, kernel_smoother_1, particle_group,
    e_dofs, b_dofs) (src/hamiltonian_splitting_boris.jl:23-88)."""

    _destroy = "gempic_boris_destroy"
    F_E1, F_E2, F_B, F_J1, F_J2, F_E1_MID, F_E2_MID, F_B_MID = range(8)

    def __init__(self, maxwell_solver, kernel_smoother_0, kernel_smoother_1, particle_group, e_dofs, b_dofs, resident=False):
        super().__init__()
        self.maxwell_solver = maxwell_solver
        self.kernel_smoother_0, self.kernel_smoother_1 = kernel_smoother_0, kernel_smoother_1
        self.particle_group = particle_group
        check(_L().gempic_boris_create(maxwell_solver.handle, kernel_smoother_0.handle, kernel_smoother_1.handle,
                                       particle_group.handle, C.byref(self._h)))
        n = kernel_smoother_0.n_dofs
        self.n = n
        self.e_dofs = [_check_alias(e_dofs[0], n), _check_alias(e_dofs[1], n)]
        self.b_dofs = _check_alias(b_dofs, n)
        self.resident = resident
        if resident:
            check(_L().gempic_boris_set_fields(self._h, dptr(self.e_dofs[0]), dptr(self.e_dofs[1]), dptr(self.b_dofs)))

    def _field(self, which):
        out = np.zeros(self.n)
        check(_L().gempic_boris_get_field(self._h, C.c_int(which), dptr(out)))
        return out

    e_dofs_mid = property(lambda s: [s._field(s.F_E1_MID), s._field(s.F_E2_MID)])
    b_dofs_mid = property(lambda s: s._field(s.F_B_MID))
    j_dofs = property(lambda s: [s._field(s.F_J1), s._field(s.F_J2)])

    def sync_fields(self):
        self.e_dofs[0][:] = self._field(self.F_E1)
        self.e_dofs[1][:] = self._field(self.F_E2)
        self.b_dofs[:] = self._field(self.F_B)

    def staggering(self, dt):
        pg = self.particle_group
        pg._flush()
        if self.resident:
            check(_L().gempic_boris_staggering(self._h, _f(dt)))
        else:
            check(_L().gempic_boris_staggering_host(self._h, _f(dt), dptr(self.e_dofs[0]), dptr(self.e_dofs[1]), dptr(self.b_dofs)))
        pg._touched()

    def strang_splitting(self, dt, number_steps):
        pg = self.particle_group
        pg._flush()
        if self.resident:
            check(_L().gempic_boris_strang_splitting(self._h, _f(dt), C.c_int64(number_steps)))
        else:
            check(_L().gempic_boris_strang_splitting_host(self._h, _f(dt), C.c_int64(number_steps), dptr(self.e_dofs[0]),
                                                          dptr(self.e_dofs[1]), dptr(self.b_dofs)))
        pg._touched()

    def _push(self, name, dt):
        pg = self.particle_group
        pg._flush()
        check(getattr(_L(), name)(self._h, _f(dt)))
        pg._touched()

    def push_v_epart(self, dt):
        self._push("gempic_boris_push_v_epart", dt)

    def push_v_bpart(self, dt):
        self._push("gempic_boris_push_v_bpart", dt)

    def push_x_accumulate_j(self, dt):
        self._push("gempic_boris_push_x_accumulate_j", dt)


# ---- samplers (src/particle_sampling.jl, src/distributions.jl, src/landau_damping.jl) ---------------------------------
class _CosGaussian:
    """parameters of CosSumGaussian{D,V} / SumCosGaussian{D,V} (src/distributions.jl:13-157); k: one wave vector per
    cosine, alpha: strengths, sigma / mu: one velocity vector per Gaussian, delta: portions"""

    def __init__(self, D, V, k, alpha, sigma, mu, delta=(1.0,)):
        self.dims = (int(D), int(V))
        self.k = [np.atleast_1d(np.asarray(kk, dtype=np.float64)) for kk in k]
        self.alpha = np.asarray(alpha, dtype=np.float64)
        self.sigma = [np.atleast_1d(np.asarray(x, dtype=np.float64)) for x in sigma]
        self.mu = [np.atleast_1d(np.asarray(x, dtype=np.float64)) for x in mu]
        self.delta = np.asarray(delta, dtype=np.float64)
        self.n_cos, self.n_gaussians = len(self.k), len(self.sigma)
        if self.n_cos != len(self.alpha) or any(len(kk) != D for kk in self.k):
            raise _lib.AssertionFailed(2, "n_cos == length(alpha), length(k[i]) == dims[1] (distributions.jl:32-35)")
        if self.n_gaussians != len(self.mu) or any(len(x) != V for x in self.sigma) or any(np.any(x == 0.0) for x in self.sigma):
            raise _lib.AssertionFailed(2, "n_gaussians == length(mu), all(sigma .!= 0.0) (distributions.jl:36-41)")
        if float(np.sum(self.delta)) != 1.0 or len(self.delta) != self.n_gaussians:
            raise _lib.AssertionFailed(2, "sum(delta) == 1.0 (distributions.jl:44)")

    def eval_x_density(self, x):
        x = np.atleast_1d(np.asarray(x, dtype=np.float64))
        return 1.0 + sum(a * np.cos(np.sum(kk * x)) for a, kk in zip(self.alpha, self.k))


class CosSumGaussian(_CosGaussian):
    """CosSumGaussian{D,V}(k, alpha, sigma, mu, delta) (src/distributions.jl:88-107)"""


class SumCosGaussian(_CosGaussian):
    """SumCosGaussian{D,V}(k, alpha, sigma, mu, delta) (src/distributions.jl:142-157)"""


class LandauDamping:
    """LandauDamping(alpha, kx) (src/landau_damping.jl:10-13)"""

    def __init__(self, alpha, kx):
        self.alpha, self.kx = float(alpha), float(kx)


class ParticleSampler:
    """ParticleSampler{D,V}(sampling_type, symmetric, n_particles, seed = 1234) (src/particle_sampling.jl:14-59);
    sampling_type "random" or "sobol"."""

    def __init__(self, D, V, sampling_type, symmetric, n_particles, seed=1234):
        if sampling_type not in ("random", "sobol"):
            raise _lib.ArgumentError(1, f"Sampling type {sampling_type} not implemented")
        n_particles = int(n_particles)
        if symmetric:   # :33-38 (sic: the remainder is added)
            rem = n_particles % 2 ** (D + V)
            if rem != 0:
                n_particles += rem
        self.sampling_type, self.dims, self.n_particles = sampling_type, (int(D), int(V)), n_particles
        self.symmetric, self.seed = bool(symmetric), int(seed)


def sample(pg, *args, first_index=0, n_global=None):
    """sample!(pg, ps::ParticleSampler, df, mesh)            src/particle_sampling.jl:68-77 ({1,2}), :248-256 ({1,1})
    sample!(pg, alpha, k, sigma, mesh)                     :266-311  (Sobol(2) + Newton Landau load)
    sample!(d::LandauDamping, pg) -- called as sample(d, pg) -- src/landau_damping.jl:34-59
    on the device.  `first_index` / `n_global`: this group holds the particles first_index .. of a load of n_global
    (sharded runs; the load does not depend on the sharding)."""
    if isinstance(pg, LandauDamping):
        d, pg = pg, args[0]
        n_global = pg.n_particles if n_global is None else n_global
        _sample_landau(pg, d.alpha, d.kx, 1.0, 2 * np.pi / d.kx / n_global, first_index, n_global)
        return
    n_global = pg.n_particles if n_global is None else n_global
    if isinstance(args[0], ParticleSampler):
        ps, df, mesh = args
        if pg.dims == (1, 1):   # :248-256
            _sample_landau(pg, float(df.alpha[0]), float(df.k[0][0]), float(df.sigma[0][0]), mesh.dimx, first_index, n_global)
            return
        if pg.dims != (1, 2):
            raise _lib.ArgumentError(1, "sample! is defined for ParticleGroup{1,1} and {1,2}")
        kk = np.ascontiguousarray([x[0] for x in df.k], dtype=np.float64)
        al = np.ascontiguousarray(df.alpha, dtype=np.float64)
        sg = np.ascontiguousarray(np.stack(df.sigma), dtype=np.float64)
        mu = np.ascontiguousarray(np.stack(df.mu), dtype=np.float64)
        de = np.ascontiguousarray(df.delta, dtype=np.float64)
        check(_L().gempic_pg_sample_cos_gaussian(pg.handle, C.c_int(1 if ps.sampling_type == "sobol" else 0),
                                                 C.c_int(1 if ps.symmetric else 0), C.c_uint64(ps.seed), _f(mesh.xmin), _f(mesh.dimx),
                                                 C.c_int(df.n_cos), dptr(kk), dptr(al), C.c_int(df.n_gaussians), dptr(sg), dptr(mu),
                                                 dptr(de), C.c_int64(first_index)))
        pg._host, pg._host_newer, pg._dev_newer = None, False, True
        return
    alpha, k, sigma, mesh = args
    if pg.dims == (1, 2) and not np.isclose(mesh.dimx, 2 * np.pi / k):
        raise _lib.AssertionFailed(2, "mesh.dimx ≈ 2π / k (particle_sampling.jl:297)")
    _sample_landau(pg, alpha, k, sigma, mesh.dimx, first_index, n_global)


def _sample_landau(pg, alpha, k, sigma, weight, first_index, n_global):
    check(_L().gempic_pg_sample_landau(pg.handle, _f(alpha), _f(k), _f(sigma), _f(weight), C.c_int64(first_index), C.c_int64(n_global)))
    pg._host, pg._host_newer, pg._dev_newer = None, False, True


DIAG_COLUMNS = ("Time", "KineticEnergy", "Momentum1", "Momentum2", "PotentialEnergyE1", "PotentialEnergyE2",
                "PotentialEnergyB3", "Transfer", "VVB", "Poynting", "ErrorPoisson")


class TimeHistoryDiagnostics:
    """TimeHistoryDiagnostics(particle_group, maxwell_solver, kernel_smoother_0, kernel_smoother_1)
    (src/diagnostics.jl:127-170); `data` is a list of 11-column rows (:143-155)."""

    def __init__(self, particle_group, maxwell_solver, kernel_smoother_0, kernel_smoother_1):
        self.particle_group, self.maxwell_solver = particle_group, maxwell_solver
        self.kernel_smoother_0, self.kernel_smoother_1 = kernel_smoother_0, kernel_smoother_1
        self.data = []

    def write_step(self, time, degree, efield_dofs, bfield_dofs, efield_dofs_n, efield_poisson):
        pg = self.particle_group
        pg._flush()
        n = self.maxwell_solver.n_dofs
        out = np.zeros(11)
        check(_L().gempic_diag_write_step(pg.handle, self.maxwell_solver.handle, self.kernel_smoother_0.handle,
                                          self.kernel_smoother_1.handle, _f(time), C.c_int(degree),
                                          dptr(_vec(efield_dofs[0], n, "e1")), dptr(_vec(efield_dofs[1], n, "e2")),
                                          dptr(_vec(bfield_dofs, n, "b")), dptr(_vec(efield_dofs_n[0], n, "e1_n")),
                                          dptr(_vec(efield_dofs_n[1], n, "e2_n")), dptr(_vec(efield_poisson, n, "e_poisson")),
                                          dptr(out)))
        self.data.append(out)
        return out


# ---- free functions with the reference's names ------------------------------------------------
def strang_splitting(h, dt, number_steps):
    h.strang_splitting(dt, number_steps)


def staggering(h, dt):
    h.staggering(dt)


def operatorHp1(h, dt):
    h.operatorHp1(dt)


def operatorHp2(h, dt):
    h.operatorHp2(dt)


def operatorHE(h, dt):
    h.operatorHE(dt)


def operatorHB(h, dt):
    h.operatorHB(dt)


def save(file, step, p, **fields):
    """save(file, step, p::ParticleGroup) (src/particle_group.jl:152-165): particle dump "<file>-<step %06d>" with the
    reference's three entries "x" (D x N), "v" (V x N), "w" (W x N), read back from the device rows (a pending HE kick
    is applied first).  The container is .npz -- JLD2 needs Julia; the Julia shim writes the reference's .jld2 from the
    same download.  Extra keyword arrays (e1=..., b=...) are stored alongside, which makes the file a restart point."""
    D, V = p.dims
    a = p.to_host()
    datafile = "%s-%06d.npz" % (file, step)
    np.savez(datafile, x=a[:D], v=a[D:D + V], w=a[D + V:], **{k: np.asarray(v) for k, v in fields.items()})
    return datafile


def load_particles(file, step, p):
    """restart from a `save` dump: uploads x, v, w into the device rows of `p`, returns the extra field arrays"""
    D, V = p.dims
    with np.load("%s-%06d.npz" % (file, step)) as z:
        p.upload(np.concatenate([z["x"], z["v"], z["w"]], axis=0))
        return {k: z[k].copy() for k in z.files if k not in ("x", "v", "w")}


def solve_poisson(efield_dofs, particle_group, kernel_smoother_0, maxwell_solver, rho):
    """solve_poisson!(efield_dofs, particle_group, kernel_smoother_0, maxwell_solver, rho) (src/diagnostics.jl:15-31)"""
    particle_group._flush()
    check(_L().gempic_solve_poisson(particle_group.handle, kernel_smoother_0.handle, maxwell_solver.handle,
                                    dptr(efield_dofs), dptr(rho)))


def write_step(thdiag, time, degree, efield_dofs, bfield_dofs, efield_dofs_n, efield_poisson):
    return thdiag.write_step(time, degree, efield_dofs, bfield_dofs, efield_dofs_n, efield_poisson)


def add_charge(rho_dofs, p, *args):
    p.add_charge(rho_dofs, *args)


def evaluate(p, *args):
    return p.evaluate(*args)


def add_current_update_v(j_dofs, p, *args):
    return p.add_current_update_v(j_dofs, *args)


def add_charge_pp(rho_dofs, p, *args):
    p.add_charge_pp(rho_dofs, *args)


def evaluate_pp(p, *args):
    return p.evaluate_pp(*args)


def add_current_update_v_pp(j_dofs, p, *args):
    return p.add_current_update_v_pp(j_dofs, *args)


def compute_e_from_rho(e, m, rho):
    m.compute_e_from_rho(e, rho)


def compute_e_from_j(e, m, current, component):
    m.compute_e_from_j(e, current, component)


def compute_e_from_b(field_out, m, delta_t, field_in):
    m.compute_e_from_b(field_out, delta_t, field_in)


def compute_b_from_e(field_out, m, delta_t, field_in):
    m.compute_b_from_e(field_out, delta_t, field_in)


def inner_product(m, c1, c2, degree):
    return m.inner_product(c1, c2, degree)


def l2norm_squared(m, c, degree):
    return m.l2norm_squared(c, degree)


def l2projection(coefs, m, func, degree):
    m.l2projection(coefs, func, degree)


def compute_rhs_from_function(coefs, m, func, degree):
    m.compute_rhs_from_function(coefs, func, degree)
