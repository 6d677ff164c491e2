"""HamiltonianSplitting{2,3} on the device (csrc/hs2d.cu) against the CPU oracle (oracle/splitting2d3v.py) on
identical seeded particle sets, plus the invariants of the scheme at a size the oracle cannot reach.

Tolerance: 1e-12 relative for particles and dofs after each operator (fp64; the device integrates the
splines with primitives, the oracle with the reference's Gauss-Legendre rule -- both exact for the
polynomial pieces; the RED / tree summation order is the only other difference)."""
import numpy as np
import pytest

from oracle import maxwell2d as m2
from oracle import oracle as orc_mod
from oracle import splitting2d3v as s2

pytestmark = pytest.mark.gpu

TOL = 1e-12


def rel(a, b):
    return np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(np.max(np.abs(b)), 1e-300)


def make_state(n, box, seed, vth=(1.0, 1.0, 0.5)):
    (xmin, lx), (ymin, ly) = box
    rng = np.random.default_rng(seed)
    st = np.empty((6, n))
    st[0] = xmin + rng.uniform(0, lx, n)
    st[1] = ymin + rng.uniform(0, ly, n)
    for k in range(3):
        st[2 + k] = vth[k] * rng.normal(size=n)
    st[5] = lx * ly * (1.0 + 0.1 * rng.uniform(size=n))
    return st


def build(gp, n, nx, ny, deg, seed, box=((0.0, 4 * np.pi), (0.0, 4 * np.pi)), resident=False, charge=-1.0, mass=1.0):
    (xmin, lx), (ymin, ly) = box
    st = make_state(n, box, seed)
    rng = np.random.default_rng(seed + 1)
    nd = nx * ny
    e0 = [0.1 * rng.normal(size=nd) for _ in range(3)]
    b0 = [0.1 * rng.normal(size=nd) for _ in range(3)]
    mesh_o = orc_mod.TwoDGrid(xmin, xmin + lx, nx, ymin, ymin + ly, ny)
    pg_o = s2.ParticleGroup23(n, charge=charge, mass=mass)
    pg_o.array[:, :] = st
    ho = s2.HamiltonianSplitting2D3V(m2.TwoDMaxwell(mesh_o, deg), pg_o, [v.copy() for v in e0], [v.copy() for v in b0])
    pg_g = gp.ParticleGroup(2, 3, n, charge=charge, mass=mass)
    pg_g.upload(st)
    mg = gp.TwoDMaxwell(gp.TwoDGrid(xmin, xmin + lx, nx, ymin, ymin + ly, ny), deg)
    hg = gp.HamiltonianSplitting2D3V(mg, pg_g, [v.copy() for v in e0], [v.copy() for v in b0], resident=resident)
    hg.set_sort_interval(0)   # per-particle comparison: keep the particle order
    return ho, hg


def check(ho, hg, box, tol=TOL, what=""):
    a, b = hg.particle_group.to_host(), ho.particle_group.array
    for d in range(2):
        L = box[d][1]
        dx = np.abs(a[d] - b[d])
        dx = np.minimum(dx, np.abs(dx - L))
        assert np.max(dx) / L < tol, f"{what}: x{d + 1}"
    for k in range(2, 5):
        assert rel(a[k], b[k]) < tol, f"{what}: v{k - 1}"
    for c in range(3):
        assert rel(hg.e_dofs[c], ho.e_dofs[c]) < tol, f"{what}: e{c + 1}"
        assert rel(hg.b_dofs[c], ho.b_dofs[c]) < tol, f"{what}: b{c + 1}"


BOX = ((0.0, 4 * np.pi), (0.0, 4 * np.pi))


@pytest.mark.parametrize("n,nx,ny,deg", [(20_000, 16, 16, 3), (9_999, 12, 20, 3), (5_000, 10, 8, 2), (4_000, 8, 8, 1), (1, 16, 16, 3)])
def test_operators_match_oracle(gp, n, nx, ny, deg):
    ho, hg = build(gp, n, nx, ny, deg, seed=n)
    for op, dt in (("operatorHB", 0.025), ("operatorHE", 0.025), ("operatorHp3", 0.025), ("operatorHp2", 0.025),
                   ("operatorHp1", 0.05), ("operatorHp2", 0.025), ("operatorHp3", 0.025), ("operatorHE", 0.025),
                   ("operatorHB", 0.025)):
        getattr(ho, op)(dt)
        getattr(hg, op)(dt)
        check(ho, hg, BOX, what=op)
    jo, jg = ho.j_dofs, hg.j_dofs
    for c in range(3):
        assert rel(jg[c], jo[c]) < TOL, f"j{c + 1}"
    assert rel(hg.charge_density(), ho.charge_density()) < TOL


def test_large_displacements_take_the_general_path(gp):
    """dt*v of several cells and an offset box: the slow path (global REDs, any number of crossings)"""
    box = ((-2.0, 3.0), (1.0, 2.0))
    ho, hg = build(gp, 3000, 12, 10, 3, seed=77, box=box)
    for op, dt in (("operatorHp1", 0.9), ("operatorHp2", 0.7), ("operatorHp1", -0.8), ("operatorHp2", -1.1)):
        getattr(ho, op)(dt)
        getattr(hg, op)(dt)
        check(ho, hg, box, tol=1e-11, what=op)


def test_strang_resident_sorted_matches_oracle(gp):
    """device-resident fields, cell sort every step: particles are permuted, so the sets are compared after
    ordering both by (x1, x2); fields must agree directly"""
    n = 30_000
    ho, hg = build(gp, n, 16, 16, 3, seed=5, resident=True)
    hg.set_sort_interval(1)
    ho.strang_splitting(0.05, 3)
    hg.strang_splitting(0.05, 3)
    hg.sync_fields()
    for c in range(3):
        assert rel(hg.e_dofs[c], ho.e_dofs[c]) < 1e-11
        assert rel(hg.b_dofs[c], ho.b_dofs[c]) < 1e-11
    a, b = hg.particle_group.to_host(), ho.particle_group.array
    ka, kb = np.lexsort((a[1], a[0])), np.lexsort((b[1], b[0]))
    assert np.max(np.abs(a[:, ka] - b[:, kb])) < 1e-9   # chaotic amplification over 3 steps; order must match
    # the sort rides in the last push of every step (sorting_hp2): the particles come back in cell order
    cell = np.floor(a[0] / (4 * np.pi / 16)).astype(int) + 16 * np.floor(a[1] / (4 * np.pi / 16)).astype(int)
    assert np.all(np.diff(cell) >= 0)
    # one step per call, with the particles replaced in between (stand-alone sort, then the riding one again)
    st = b.copy()
    hg.particle_group.upload(st)
    for _ in range(2):
        ho.strang_splitting(0.05, 1)
        hg.strang_splitting(0.05, 1)
    hg.sync_fields()
    for c in range(3):
        assert rel(hg.e_dofs[c], ho.e_dofs[c]) < 1e-10
    a, b = hg.particle_group.to_host(), ho.particle_group.array
    ka, kb = np.lexsort((a[1], a[0])), np.lexsort((b[1], b[0]))
    assert np.max(np.abs(a[:, ka] - b[:, kb])) < 1e-8
    cell = np.floor(a[0] / (4 * np.pi / 16)).astype(int) + 16 * np.floor(a[1] / (4 * np.pi / 16)).astype(int)
    assert np.all(np.diff(cell) >= 0)
    hg.particle_group.sort(hg.maxwell_solver)
    a = hg.particle_group.to_host()
    cell = np.floor(a[0] / (4 * np.pi / 16)).astype(int) + 16 * np.floor(a[1] / (4 * np.pi / 16)).astype(int)
    assert np.all(np.diff(cell) >= 0)


@pytest.mark.parametrize("deg,nx,ny", [(3, 8, 8), (2, 8, 6), (1, 6, 8), (3, 16, 16)])
def test_sorted_fast_path_matches_oracle(gp, deg, nx, ny):
    """fuse level 2 (default): [HE,(HE,)Hp3,Hp2] and the trailing Hp3 of a Strang step run from the cell-sorted order
    with register accumulators and warp-uniform field tables (k2_sorted).  Few cells and many particles per cell make
    long cell runs, so nearly every iteration takes the fast branch; cell-run boundaries take the general one.  The
    sort permutes the particles: they are matched through their distinct weights."""
    n = 160_000
    ho, hg = build(gp, n, nx, ny, deg, seed=deg * 100 + nx, resident=True)
    w = (4 * np.pi) ** 2 * (1.0 + np.random.default_rng(5).permutation(n) / n)
    ho.particle_group.array[5] = w
    st = ho.particle_group.array.copy()
    st[2:5] *= 1.5                                   # ~30 % of the particles change cell in a half step
    ho.particle_group.array[:, :] = st
    hg.particle_group.upload(st)
    hg.set_sort_interval(1)
    hg.set_fusion(2)
    ho.strang_splitting(0.05, 2)
    hg.strang_splitting(0.05, 2)                     # one call: the trailing HE rides in the second step's head pass
    ho.strang_splitting(0.04, 1)
    hg.strang_splitting(0.04, 1)                     # next call: the deferred kick rides in its head pass (other dt)
    hg.sync_fields()
    for c in range(3):
        assert rel(hg.e_dofs[c], ho.e_dofs[c]) < 1e-11, f"e{c + 1}"
        assert rel(hg.b_dofs[c], ho.b_dofs[c]) < 1e-11, f"b{c + 1}"
    jo, jg = ho.j_dofs, hg.j_dofs
    for c in range(3):
        assert rel(jg[c], jo[c]) < 1e-11, f"j{c + 1}"
    a, b = hg.particle_group.to_host(), ho.particle_group.array
    a, b = a[:, np.argsort(a[5])], b[:, np.argsort(b[5])]
    assert np.array_equal(a[5], b[5])
    for d in range(2):
        dx = np.abs(a[d] - b[d])
        assert np.max(np.minimum(dx, np.abs(dx - 4 * np.pi))) < 1e-10 * 4 * np.pi
    assert np.max(np.abs(a[2:5] - b[2:5])) < 1e-10 * np.max(np.abs(b[2:5]))
    # fuse levels agree with each other on the device as well
    hg1 = build(gp, n, nx, ny, deg, seed=deg * 100 + nx, resident=True)[1]
    hg1.particle_group.upload(st)
    hg1.set_sort_interval(1)
    hg1.set_fusion(1)
    hg1.strang_splitting(0.05, 2)
    hg1.strang_splitting(0.04, 1)
    hg1.sync_fields()
    for c in range(3):
        assert rel(hg1.e_dofs[c], hg.e_dofs[c]) < 1e-11


@pytest.mark.parametrize("fuse", [False, 1, True])
def test_strang_fused_and_unfused_match_oracle(gp, fuse):
    """strang_splitting!(h, dt, 4) in one call (the fused path folds the trailing HE into the next step's pass),
    particle order kept"""
    ho, hg = build(gp, 12_000, 16, 12, 3, seed=31, resident=True)
    hg.set_fusion(fuse)
    ho.strang_splitting(0.05, 4)
    hg.strang_splitting(0.05, 4)
    hg.sync_fields()
    check(ho, hg, BOX, tol=1e-10, what=f"strang fuse={fuse}")


def test_deferred_trailing_kick_is_invisible(gp):
    """the fused 2d3v strang_splitting! defers its trailing HE particle kick to its next call (hs2d.cu, pending2d): no
    sequence of calls may observe the lag -- varying dt, particle downloads, single operators and moments in between"""
    ho, hg = build(gp, 8_000, 16, 12, 3, seed=77, resident=False)
    for dt in (0.05, 0.03, 0.07):                      # back to back: the kick rides in the next call's first pass
        ho.strang_splitting(dt, 1), hg.strang_splitting(dt, 1)
        for c in range(3):
            assert rel(hg.e_dofs[c], ho.e_dofs[c]) < 1e-10 and rel(hg.b_dofs[c], ho.b_dofs[c]) < 1e-10
    check(ho, hg, BOX, tol=1e-10, what="download after three calls")      # the download applies the pending kick
    ho.strang_splitting(0.05, 2), hg.strang_splitting(0.05, 2)
    ho.operatorHp1(0.02), hg.operatorHp1(0.02)         # a single operator after a fused call
    check(ho, hg, BOX, tol=1e-10, what="operator after strang")
    ho.strang_splitting(0.04, 1), hg.strang_splitting(0.04, 1)
    a = ho.particle_group.array
    ke_ref = float(np.sum(a[5] * (a[2] ** 2 + a[3] ** 2 + a[4] ** 2)))
    m = hg.moments()                                   # sum_p w |v|^2 must see the kick
    assert abs(m[0] - ke_ref) < 1e-10 * ke_ref
    ho.strang_splitting(0.05, 1), hg.strang_splitting(0.05, 1)
    check(ho, hg, BOX, tol=1e-9, what="after moments")


def test_strang_round_trip_at_scale(gp):
    """size-independent property: every sub-flow is exact and the Strang composition symmetric, so three steps forward
    and three steps back (dt -> -dt) return 3e5 particles and all six fields to their start to round-off -- with the
    sort riding in the pushes (particles are matched through their distinct weights) and the deferred trailing kick"""
    n, nx = 300_000, 32
    box = ((0.0, 4 * np.pi), (0.0, 4 * np.pi))
    st = make_state(n, box, seed=91)
    st[5] = (4 * np.pi) ** 2 * (1.0 + np.random.default_rng(92).permutation(n) / n)   # distinct weights = particle ids
    nd = nx * nx
    ix, iy = np.arange(nd) % nx, np.arange(nd) // nx
    mode = lambda a, b, ph: 0.05 * np.cos(2 * np.pi * (a * ix + b * iy) / nx + ph)
    e0 = [mode(1, 0, 0.3), mode(0, 1, 1.1), mode(1, 1, 2.0)]
    b0 = [mode(1, 1, 0.7), mode(1, 0, 1.9), mode(0, 1, 0.2)]
    pg = gp.ParticleGroup(2, 3, n, charge=-1.0)
    pg.upload(st)
    mg = gp.TwoDMaxwell(gp.TwoDGrid(0.0, 4 * np.pi, nx, 0.0, 4 * np.pi, nx), 3)
    e, b = [v.copy() for v in e0], [v.copy() for v in b0]
    h = gp.HamiltonianSplitting2D3V(mg, pg, e, b, resident=True)
    h.strang_splitting(0.05, 2)
    h.strang_splitting(0.05, 1)
    mid = pg.to_host()
    assert np.max(np.abs(np.sort(mid[2]) - np.sort(st[2]))) > 1e-4            # something happened
    h.strang_splitting(-0.05, 1)
    h.strang_splitting(-0.05, 2)
    h.sync_fields()
    a = pg.to_host()
    ka, kb = np.argsort(a[5]), np.argsort(st[5])
    a, ref = a[:, ka], st[:, kb]
    assert np.array_equal(a[5], ref[5])
    for d in range(2):
        dx = np.abs(a[d] - ref[d])
        assert np.max(np.minimum(dx, np.abs(dx - 4 * np.pi))) < 1e-11
    assert np.max(np.abs(a[2:5] - ref[2:5])) < 1e-11
    for c in range(3):
        assert np.max(np.abs(e[c] - e0[c])) < 1e-11 and np.max(np.abs(b[c] - b0[c])) < 1e-11


def test_invariants_at_scale(gp):
    """2e6 particles on 64x64, degree 3 (the BASELINE config 5 grid): Gauss law conserved to round-off,
    total charge exact, energy drift O(dt^2)"""
    n, nx = 2_000_000, 64
    L = 4 * np.pi
    pg = gp.ParticleGroup(2, 3, n)
    pg.sample("landau", 0.0, L, alpha=0.5, k=0.5, sigma=(1.0, 1.0, 1.0), seed=1234)
    mg = gp.TwoDMaxwell(gp.TwoDGrid(0.0, L, nx, 0.0, L, nx), 3)
    nd = nx * nx
    e = [np.zeros(nd) for _ in range(3)]
    b = [np.zeros(nd) for _ in range(3)]
    h = gp.HamiltonianSplitting2D3V(mg, pg, e, b, resident=True)
    rho = h.charge_density()
    assert abs(rho.sum() - L * L) < 1e-9 * L * L            # sum_p q w / N = Lx Ly
    mg.compute_e_from_rho(e, rho - rho.mean())               # neutralising background
    b[2][:] = 1e-3 * np.cos(2 * np.pi * (np.arange(nd) % nx) / nx)
    h.upload_fields()
    r0 = h.gauss_residual()
    en0 = sum(h.energies())
    h.strang_splitting(0.05, 5)
    r1 = h.gauss_residual()
    assert np.max(np.abs(r1 - r0)) < 1e-11 * np.max(np.abs(rho))
    en1 = sum(h.energies())
    assert abs(en1 - en0) < 1e-3 * en0
    st = pg.to_host()
    assert np.all((st[0] >= 0) & (st[0] < L) & (st[1] >= 0) & (st[1] < L))
    assert np.all(np.isfinite(st))


def test_argument_errors(gp):
    mg = gp.TwoDMaxwell(gp.TwoDGrid(0.0, 1.0, 16, 0.0, 1.0, 16), 3)
    nd = 256
    f = lambda: [np.zeros(nd) for _ in range(3)]
    with pytest.raises(gp.AssertionFailed):
        gp.HamiltonianSplitting2D3V(mg, gp.ParticleGroup(1, 2, 8), f(), f())
    h = gp.HamiltonianSplitting2D3V(mg, gp.ParticleGroup(2, 3, 8), f(), f())
    with pytest.raises(gp.ArgumentError):
        h._op(9, 0.1)
