"""The reference's own TwoDMaxwell test (test/test_maxwell_2d_fem.jl), restated line by line against the
numpy oracle (oracle/maxwell2d.py).  CPU only.  The same body runs against the CUDA path in
tests/test_gpu_maxwell2d.py (shared through `run_reference_maxwell2d_test`)."""
import numpy as np

from oracle import maxwell2d as m2
from oracle import oracle as orc


def evaluate_spline_2d(nx1, nx2, degs, dofs):
    """test_maxwell_2d_fem.jl:5-15"""
    deg1, deg2 = degs
    vals = np.array(dofs).reshape(nx2, nx1).T.copy()
    for j in range(nx2):
        vals[:, j] = m2.eval_uniform_periodic_spline_curve(deg1, vals[:, j].copy())
    for i in range(nx1):
        vals[i, :] = m2.eval_uniform_periodic_spline_curve(deg2, vals[i, :].copy())
    return vals.T.reshape(-1)


def run_reference_maxwell2d_test(make_maxwell, nsteps=300):
    """test_maxwell_2d_fem.jl:17-146 with `make_maxwell(mesh, deg)` supplying the solver under test.
    The solver must offer the reference's method names."""
    x1min, x1max, nx1 = 0.0, 2 * np.pi, 16
    x2min, x2max, nx2 = 0.0, 2 * np.pi, 32
    mesh = orc.TwoDGrid(x1min, x1max, nx1, x2min, x2max, nx2)
    deg, delta_t = 3, 0.01
    maxwell = make_maxwell(mesh, deg)
    n = nx1 * nx2
    efield = [np.zeros(n) for _ in range(3)]
    bfield = [np.zeros(n) for _ in range(3)]
    xs = np.linspace(x1min, x1max, nx1 + 1)[:-1]
    ys = np.linspace(x2min, x2max, nx2 + 1)[:-1]
    X = np.tile(xs, nx2)            # flat, x fastest
    Y = np.repeat(ys, nx1)
    w1 = np.sqrt(3.0)
    t = {"time": 0.0}
    sin_k = lambda x, y: np.sin((x + y) - w1 * t["time"])
    cos_k = lambda x, y: np.cos((x + y) - w1 * t["time"])

    rho = maxwell.compute_rhs_from_function(cos_k, 1, 0)
    maxwell.compute_e_from_rho(efield, rho)
    v1 = evaluate_spline_2d(nx1, nx2, (deg - 1, deg), efield[0])
    v2 = evaluate_spline_2d(nx1, nx2, (deg, deg - 1), efield[1])
    v3 = evaluate_spline_2d(nx1, nx2, (deg, deg), efield[2])
    ref = sin_k(X, Y) / 2
    # Julia's `a ≈ b rtol=r` on vectors: norm(a-b) <= r*max(norm(a), norm(b))
    approx = lambda a, b, rtol: np.linalg.norm(a - b) <= rtol * max(np.linalg.norm(a), np.linalg.norm(b))
    assert approx(v1, ref, 1e-4)
    assert approx(v2, ref, 1e-4)
    assert np.linalg.norm(v3) == 0.0

    e1 = lambda x, y: np.cos(x) * np.sin(y) * np.sin(np.sqrt(2) * t["time"]) / np.sqrt(2)
    e2 = lambda x, y: -np.sin(x) * np.cos(y) * np.sin(np.sqrt(2) * t["time"]) / np.sqrt(2)
    b3 = lambda x, y: -np.cos(x) * np.cos(y) * np.cos(np.sqrt(2) * t["time"])

    t["time"] = -0.5 * delta_t
    bfield[0][:] = maxwell.l2projection(e1, 1, 2)
    bfield[1][:] = maxwell.l2projection(e2, 2, 2)
    bfield[2][:] = maxwell.l2projection(b3, 3, 2)
    t["time"] = 0.0
    efield[0][:] = maxwell.l2projection(e1, 1, 1)
    efield[1][:] = maxwell.l2projection(e2, 2, 1)
    efield[2][:] = maxwell.l2projection(b3, 3, 1)
    efield[2] *= -1

    for _ in range(nsteps):
        maxwell.compute_b_from_e(bfield, delta_t, efield)
        maxwell.compute_e_from_b(efield, delta_t, bfield)

    bv = [evaluate_spline_2d(nx1, nx2, (deg, deg - 1), bfield[0]),
          evaluate_spline_2d(nx1, nx2, (deg - 1, deg), bfield[1]),
          evaluate_spline_2d(nx1, nx2, (deg - 1, deg - 1), bfield[2])]
    t["time"] = (nsteps - 0.5) * delta_t
    assert approx(e1(X, Y), bv[0], 1e-4)
    assert approx(e2(X, Y), bv[1], 1e-4)
    assert approx(b3(X, Y), bv[2], 1e-3)
    ev = [evaluate_spline_2d(nx1, nx2, (deg - 1, deg), efield[0]),
          evaluate_spline_2d(nx1, nx2, (deg, deg - 1), efield[1]),
          evaluate_spline_2d(nx1, nx2, (deg, deg), efield[2])]
    t["time"] = nsteps * delta_t
    assert approx(e1(X, Y), ev[0], 1e-4)
    assert approx(e2(X, Y), ev[1], 1e-4)
    assert approx(-b3(X, Y), ev[2], 1e-3)

    efield[0][:] = maxwell.l2projection(cos_k, 1, 1)
    error2 = maxwell.inner_product(efield[0], efield[0], 1, 1) - 2 * np.pi**2
    assert abs(error2) < 1e-5

    rho = maxwell.compute_rhs_from_function(sin_k, 1, 1)
    maxwell.compute_e_from_j(efield[0], rho, 1)
    v1 = evaluate_spline_2d(nx1, nx2, (deg - 1, deg), efield[0])
    ref = cos_k(X, Y) - sin_k(X, Y)
    assert np.max(np.abs(v1 - ref)) < 1e-2

    t["time"] = 0.0
    rho_ref = 2.0 * maxwell.compute_rhs_from_function(cos_k, 1, 0)
    efield[0][:] = maxwell.l2projection(sin_k, 1, 1)
    efield[1][:] = maxwell.l2projection(sin_k, 2, 1)
    efield[2][:] = maxwell.l2projection(sin_k, 3, 1)
    rho = np.zeros(n)
    maxwell.compute_rho_from_e(rho, efield)
    assert approx(rho, rho_ref, np.sqrt(np.finfo(float).eps))
    return maxwell, efield, bfield


def test_reference_maxwell2d_test_against_oracle():
    run_reference_maxwell2d_test(lambda mesh, deg: m2.TwoDMaxwell(mesh, deg))


def test_mass_lines_closed_form():
    """spline_fem_mass_line reproduces the closed-form rows Maxwell1DFEM hard-codes
    (src/maxwell_1d_fem.jl:60-92): deg 3 -> (2416, 1191, 120, 1)/5040, deg 2 -> (66, 26, 1)/120"""
    assert np.allclose(m2.spline_fem_mass_line(3), np.array([2416, 1191, 120, 1]) / 5040, rtol=0, atol=1e-15)
    assert np.allclose(m2.spline_fem_mass_line(2), np.array([66, 26, 1]) / 120, rtol=0, atol=1e-15)
    assert np.allclose(m2.spline_fem_mass_line(1), np.array([4, 1]) / 6, rtol=0, atol=1e-15)


def test_mass_solve_inverts_mass_multiply():
    mesh = orc.TwoDGrid(0.0, 1.0, 8, 0.0, 2.0, 12)
    mx = m2.TwoDMaxwell(mesh, 3)
    rng = np.random.default_rng(3)
    c = rng.normal(size=8 * 12)
    for comp in (1, 2, 3):
        l1, l2 = mx._mass_lines(comp, 1)
        assert np.allclose(mx.inv_mass_1[comp - 1].solve(mx.multiply_mass_2dkron(l1, l2, c)), c, atol=1e-12)
        l1, l2 = mx._mass_lines(comp, 2)
        assert np.allclose(mx.inv_mass_2[comp - 1].solve(mx.multiply_mass_2dkron(l1, l2, c)), c, atol=1e-12)
