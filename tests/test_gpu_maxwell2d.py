"""TwoDMaxwell on the device (gempic_maxwell2d_*, csrc/fields2d.cu) against
  (a) the reference's own test body (test/test_maxwell_2d_fem.jl, restated in test_oracle_maxwell2d.py), and
  (b) the numpy oracle dof by dof on random data.
Tolerance: 1e-12 relative to the largest dof of each vector (fp64; the device applies the mass inverse as
two 1D circulant convolutions instead of FFTs, which is the same operator up to rounding)."""
import numpy as np
import pytest

from oracle import maxwell2d as m2
from oracle import oracle as orc_mod

from .test_oracle_maxwell2d import run_reference_maxwell2d_test

pytestmark = pytest.mark.gpu

TOL = 1e-12


def rel(a, b):
    return np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(np.max(np.abs(b)), 1e-300)


def test_reference_maxwell2d_test_on_gpu(gp):
    """test_maxwell_2d_fem.jl:17-146 through the C ABI; the resulting dofs also match the oracle's."""
    _, e_g, b_g = run_reference_maxwell2d_test(lambda mesh, deg: gp.TwoDMaxwell(gp.TwoDGrid(mesh.xmin, mesh.xmax, mesh.nx, mesh.ymin, mesh.ymax, mesh.ny), deg), nsteps=300)
    _, e_o, b_o = run_reference_maxwell2d_test(lambda mesh, deg: m2.TwoDMaxwell(mesh, deg), nsteps=300)
    for c in range(3):
        assert rel(e_g[c], e_o[c]) < 1e-10   # 600 field updates
        assert rel(b_g[c], b_o[c]) < 1e-10


@pytest.mark.parametrize("nx,ny,deg", [(16, 32, 3), (64, 64, 3), (12, 10, 2), (8, 6, 1), (30, 18, 3)])
def test_field_operators_match_oracle(gp, nx, ny, deg):
    mesh_o = orc_mod.TwoDGrid(0.3, 0.3 + 4 * np.pi, nx, -1.0, 5.0, ny)
    mo = m2.TwoDMaxwell(mesh_o, deg)
    mg = gp.TwoDMaxwell(gp.TwoDGrid(mesh_o.xmin, mesh_o.xmax, nx, mesh_o.ymin, mesh_o.ymax, ny), deg)
    n = nx * ny
    rng = np.random.default_rng(nx * 100 + ny)
    # tables
    for a in range(2):
        assert np.allclose(mg.mass_line_0[a], mo.mass_line_0[a], rtol=1e-15, atol=0)
        assert np.allclose(mg.mass_line_1[a], mo.mass_line_1[a], rtol=1e-15, atol=0)
        # spline_fem_compute_mass_eig of those lines (maxwell_2d_fem.jl:51-54, poisson_2d_fem.jl:113-126)
        nn = (nx, ny)[a]
        assert np.allclose(mg._table(2, a), m2.spline_fem_compute_mass_eig(nn, deg, mo.mass_line_0[a]), rtol=1e-14, atol=0)
        assert np.allclose(mg._table(3, a), m2.spline_fem_compute_mass_eig(nn, deg - 1, mo.mass_line_1[a]), rtol=1e-14, atol=0)
    # mass multiply / solve for every (component, form)
    c = rng.normal(size=n)
    for form in (1, 2):
        for comp in (1, 2, 3):
            l1, l2 = mo._mass_lines(comp, form)
            ref = mo.multiply_mass_2dkron(l1, l2, c)
            assert rel(mg.multiply_mass(c, comp, form), ref) < TOL
            inv = (mo.inv_mass_1 if form == 1 else mo.inv_mass_2)[comp - 1]
            assert rel(mg.solve_mass(c, comp, form), inv.solve(c)) < TOL
    for form, comp in ((0, 1), (3, 1)):
        l1, l2 = mo._mass_lines(comp, form)
        assert rel(mg.multiply_mass(c, comp, form), mo.multiply_mass_2dkron(l1, l2, c)) < TOL
    # field updates
    e_o = [rng.normal(size=n) for _ in range(3)]
    b_o = [rng.normal(size=n) for _ in range(3)]
    e_g, b_g = [v.copy() for v in e_o], [v.copy() for v in b_o]
    mo.compute_b_from_e(b_o, 0.05, e_o)
    mg.compute_b_from_e(b_g, 0.05, e_g)
    for k in range(3):
        assert rel(b_g[k], b_o[k]) < TOL
    mo.compute_e_from_b(e_o, 0.05, b_o)
    mg.compute_e_from_b(e_g, 0.05, b_g)
    for k in range(3):
        assert rel(e_g[k], e_o[k]) < TOL
    for comp in (1, 2, 3):
        j = rng.normal(size=n)
        mo.compute_e_from_j(e_o[comp - 1], j, comp)
        mg.compute_e_from_j(e_g[comp - 1], j, comp)
        assert rel(e_g[comp - 1], e_o[comp - 1]) < TOL
    rho_o, rho_g = np.zeros(n), np.zeros(n)
    mo.compute_rho_from_e(rho_o, e_o)
    mg.compute_rho_from_e(rho_g, e_g)
    assert rel(rho_g, rho_o) < TOL
    # Poisson: zero-mean right-hand side
    rho = rng.normal(size=n)
    rho -= rho.mean()
    ep_o, ep_g = [np.zeros(n), np.zeros(n)], [np.zeros(n), np.zeros(n)]
    mo.compute_e_from_rho(ep_o, rho)
    mg.compute_e_from_rho(ep_g, rho)
    assert rel(ep_g[0], ep_o[0]) < 1e-11 and rel(ep_g[1], ep_o[1]) < 1e-11
    # inner products
    a, b = rng.normal(size=n), rng.normal(size=n)
    for form, comp in ((0, 1), (1, 1), (1, 2), (1, 3), (2, 1), (2, 2), (2, 3), (3, 1)):
        ref = mo.inner_product(a, b, comp, form)
        assert abs(mg.inner_product(a, b, comp, form) - ref) < 1e-12 * max(1.0, abs(ref), np.linalg.norm(a) * np.linalg.norm(b))
    # quadrature right-hand sides and projections
    f = lambda x, y: np.sin(0.5 * x) * np.cos(y) + 0.3
    for form, comp in ((0, 1), (1, 1), (2, 3)):
        assert rel(mg.compute_rhs_from_function(f, comp, form), mo.compute_rhs_from_function(f, comp, form)) < 1e-13
    assert rel(mg.l2projection(f, 2, 1), mo.l2projection(f, 2, 1)) < TOL


def test_discrete_gauss_law_is_preserved_by_the_field_updates(gp):
    """div-free property of the discrete curl: compute_e_from_b! does not change compute_rho_from_e!"""
    nx, ny = 32, 24
    mg = gp.TwoDMaxwell(gp.TwoDGrid(0.0, 2 * np.pi, nx, 0.0, 3.0, ny), 3)
    rng = np.random.default_rng(5)
    e = [rng.normal(size=nx * ny) for _ in range(3)]
    b = [rng.normal(size=nx * ny) for _ in range(3)]
    r0, r1 = np.zeros(nx * ny), np.zeros(nx * ny)
    mg.compute_rho_from_e(r0, e)
    mg.compute_e_from_b(e, 0.1, b)
    mg.compute_rho_from_e(r1, e)
    assert np.max(np.abs(r1 - r0)) < 1e-11 * np.max(np.abs(r0))


def test_argument_errors(gp):
    with pytest.raises(gp.ArgumentError):
        gp.TwoDMaxwell(gp.TwoDGrid(0.0, 1.0, 16, 0.0, 1.0, 16), 4)
    mg = gp.TwoDMaxwell(gp.TwoDGrid(0.0, 1.0, 16, 0.0, 1.0, 16), 3)
    with pytest.raises(gp.ArgumentError):
        mg.compute_e_from_j(np.zeros(256), np.zeros(256), 4)
    with pytest.raises(gp.ArgumentError):
        mg.l2projection(lambda x, y: 1.0, 1, 0)
