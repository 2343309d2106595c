"""CPU tests of the oracle itself.  The reference ships no golden vectors and cannot be built here (no DUNE), so the
oracle is PARITY UNPINNED against real PDELab; what pins it instead: patch tests, exact-derivative vs finite-difference
agreement, an independent 1-D boundary-value solve of the one-wall Poisson-Boltzmann problem (the analytic anchor of
/root/reference/test/one_wall_dh/one_wall.gp), h-convergence, and regression against the committed golden solutions."""
import os

import numpy as np
import pytest

import util
from oracle import binding as ora


def case(name, levels=0):
    m = ora.Mesh.from_arrays(**util.load_mesh_arrays(name)).refine(levels)
    return m, ora.Params.read(util.cfg_path(name))


def csr(rp, col, val):
    import scipy.sparse as sp
    return sp.csr_matrix((val, col, rp))


@pytest.mark.parametrize("name,nv,nT,nB", [("one_wall", 46, 64, 26), ("sphere", 213, 357, 67), ("cylinder", 311, 541, 79),
                                           ("pore_small", 320, 545, 93), ("pore", 3048, 5744, 350)])
def test_fixture_sizes_match_survey(name, nv, nT, nB):
    m, _ = case(name)
    assert (m.nv, m.nT, m.nB) == (nv, nT, nB)


def test_gmsh_reader_semantics(tmp_path):
    a = util.load_mesh_arrays("one_wall")
    path = str(tmp_path / "m.msh")
    util.write_gmsh(path, a, shuffle_nodes=True, extra_nodes=2)  # node order in the file and unused nodes are irrelevant
    m = ora.Mesh.read_gmsh(path)
    for k in a:
        assert np.array_equal(getattr(m, k), a[k]), k


@pytest.mark.parametrize("name", util.MESHES)
def test_gmsh_reader_on_the_reference_files(name):
    """tests/golden/msh/*.msh are the reference's own files (Gmsh 2.1 and 2.2 ASCII, copied verbatim by
    scripts/make_fixtures.py): reading them gives the committed fixture arrays."""
    m = ora.Mesh.read_gmsh(os.path.join(util.GOLDEN, "msh", name + ".msh"))
    a = util.load_mesh_arrays(name)
    for k in a:
        assert np.array_equal(getattr(m, k), a[k]), k


@pytest.mark.parametrize("name", util.MESHES)
def test_fast_pattern_equals_the_literal_restatement(name):
    """make_pattern() (vertex adjacency, O(nnz)) against the std::set-per-row restatement of PDELab's pattern rule."""
    m, p = case(name, 1)
    for F, comp0 in ((1, 0), (1, 1), (1, 2), (3, 0)):
        rp, col = ora.pattern(m, p, F, comp0)
        rp2, col2 = ora.pattern(m, p, F, comp0, literal=True)
        assert np.array_equal(rp, rp2) and np.array_equal(col, col2)


def test_pore_without_dna_mesh_is_reproducible_and_faithful_to_the_geo():
    """BASELINE config C4: the reference ships only pore_without_dna.geo; scripts/make_pore_without_dna_mesh.py
    triangulates it deterministically.  The committed fixture is what the script produces, the physical line tags are the
    .geo's (:69-74) and the domain is the .geo's polygon (box 100 x 55 minus the membrane with its two rounded corners)."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("mkmesh", os.path.join(root, "scripts", "make_pore_without_dna_mesh.py"))
    mk = importlib.util.module_from_spec(spec); spec.loader.exec_module(mk)
    g = mk.generate()
    a = util.load_mesh_arrays("pore_without_dna")
    for k in a:
        assert np.array_equal(g[k], a[k]), k
    x, y, tri = a["x"], a["y"], a["tri"]
    p0, p1, p2 = (np.stack([x[tri[:, i]], y[tri[:, i]]], axis=1) for i in range(3))
    det = (p1[:, 0] - p0[:, 0]) * (p2[:, 1] - p0[:, 1]) - (p2[:, 0] - p0[:, 0]) * (p1[:, 1] - p0[:, 1])
    assert np.all(det > 0)
    membrane = 20.0 * 45.0 - 2 * (1.0 - np.pi / 4)  # 20 x (55 - 10) block, two corners rounded with radius 1
    # each quarter circle is two chords: 4 circular segments of area (pi/4 - sin(pi/4)) / 2 are missing from the membrane
    assert abs(0.5 * det.sum() - (100.0 * 55.0 - membrane) - 2 * (np.pi / 4 - np.sin(np.pi / 4))) < 1e-9
    # tags: axis r = 0 -> 1, inflow z = -50 -> 2, outflow z = +50 -> 3, top left -> 4, top right -> 5, membrane -> 0
    mx, my = 0.5 * (x[a["ba"]] + x[a["bb"]]), 0.5 * (y[a["ba"]] + y[a["bb"]])
    ph = a["bphys"]
    assert np.all(ph[my == 0] == 1) and np.all(ph[mx == -50] == 2) and np.all(ph[mx == 50] == 3)
    assert np.all(ph[(my == 55) & (mx < 0)] == 4) and np.all(ph[(my == 55) & (mx > 0)] == 5)
    assert np.all((np.abs(mx[ph == 0]) <= 10 + 1e-9) & (my[ph == 0] >= 10 - 1e-9))
    assert sorted(set(ph.tolist())) == [0, 1, 2, 3, 4, 5]


@pytest.mark.parametrize("op", [ora.OP_PB, ora.OP_PNP])
def test_multicore_assembly_is_bit_identical_to_the_sequential_one(op):
    """The OpenMP variants of the CPU baseline (two phases, every dof sums its elements in ascending element order):
    same bits as residual() / jacobian() for 1, 3 and 8 threads."""
    m, p = case("pore", 1)
    u = np.random.RandomState(4).uniform(0.01, 1.0, ora.nfields(op) * m.nv)
    r0 = ora.residual(m, p, op, u)
    _, _, v0 = ora.jacobian(m, p, op, u, mode=0)
    for threads in (1, 3, 8):
        r, v = ora.assembly_par(m, p, op, u, threads, 0)
        assert np.array_equal(r, r0) and np.array_equal(v, v0)


def test_config_reader(tmp_path):
    p = ora.Params.read(util.cfg_path("pore"))
    assert p.sys[0] == 7 and p.sys[1] == 1 and p.sys[4] == 3.1415 and p.sys[7] == 1e-9 and p.sys[8] == 1e-8
    assert p.surf[0, 0] == 1 and p.surf[0, 1] == 1.1 and p.surf[4, 0] == 0 and p.surf[4, 2] == 24.1
    bad = tmp_path / "bad.cfg"
    bad.write_text("[system]\nn_surfaces=1\n[mesh]\nfilename=x.msh\n[surface_0]\ncoulombBtype=0\n")
    with pytest.raises(RuntimeError, match="coulombPotential"):  # only the value matching the type is required
        ora.Params.read(str(bad))
    with pytest.raises(RuntimeError):
        ora.Params.read(str(tmp_path / "missing.cfg"))


@pytest.mark.parametrize("name", ["one_wall", "pore_small"])
def test_refinement_counts_and_geometry(name):
    m, _ = case(name)
    f = m.refine(1)
    nE = m.nv + m.nT - 1  # Euler characteristic 1
    assert (f.nv, f.nT, f.nB) == (m.nv + nE, 4 * m.nT, 2 * m.nB)

    def area(mm):
        x, y, t = mm.x, mm.y, mm.tri
        return 0.5 * np.abs((x[t[:, 1]] - x[t[:, 0]]) * (y[t[:, 2]] - y[t[:, 0]]) - (x[t[:, 2]] - x[t[:, 0]]) * (y[t[:, 1]] - y[t[:, 0]])).sum()
    assert abs(area(f) - area(m)) <= 1e-12 * area(m)
    assert np.array_equal(f.bphys[0::2], m.bphys) and np.array_equal(f.bphys[1::2], m.bphys)
    assert np.array_equal(f.x[:m.nv], m.x)  # old vertices keep their index


def test_patch_test_poisson_laplace():
    """A linear field has zero Laplace residual at every vertex that is not on the boundary (planar case)."""
    m, p = case("sphere")
    u = 0.3 + 1.7 * m.x - 0.4 * m.y
    c = np.full(m.nv, 0.06)
    r, ab = ora.residual(m, p, ora.OP_POISSON, u, c, c, want_abs=True)
    onb = np.zeros(m.nv, bool); onb[m.ba] = True; onb[m.bb] = True
    assert np.max(np.abs(r[~onb]) / ab[~onb]) < 1e-13


def test_mass_operator_integrates_area():
    m, p = case("cylinder")
    # element contributions of M*1 are int phi_i > 0; their sum (taken before the Dirichlet rows are zeroed) is |Omega|
    _, contrib = ora.residual(m, p, ora.OP_MASS, np.ones(m.nv), want_abs=True)
    x, y, t = m.x, m.y, m.tri
    area = 0.5 * np.abs((x[t[:, 1]] - x[t[:, 0]]) * (y[t[:, 2]] - y[t[:, 0]]) - (x[t[:, 2]] - x[t[:, 0]]) * (y[t[:, 1]] - y[t[:, 0]])).sum()
    assert abs(contrib.sum() - area) <= 1e-12 * area


@pytest.mark.parametrize("op", [ora.OP_PB, ora.OP_POISSON, ora.OP_DIFFUSION, ora.OP_MASS, ora.OP_PNP])
@pytest.mark.parametrize("name", ["one_wall", "cylinder"])
def test_fd_jacobian_agrees_with_exact_derivative(name, op):
    m, p = case(name)
    rng = np.random.RandomState(0)
    F = ora.nfields(op)
    u = rng.uniform(-1, 1, F * m.nv); a0 = rng.uniform(0, 1, m.nv); a1 = rng.uniform(0, 1, m.nv)
    _, _, fd = ora.jacobian(m, p, op, u, a0, a1, valency=-1.0, mode=0, eps=1e-7)
    _, _, ex = ora.jacobian(m, p, op, u, a0, a1, valency=-1.0, mode=1)
    assert np.max(np.abs(fd - ex)) <= 2e-6 * np.max(np.abs(ex))
    # reference epsilon 1e-11 (PDELab <= 1.1): noise floor ~ eps_mach/eps = 1e-5
    _, _, fd11 = ora.jacobian(m, p, op, u, a0, a1, valency=-1.0, mode=0, eps=1e-11)
    assert np.max(np.abs(fd11 - ex)) <= 1e-3 * np.max(np.abs(ex))


def test_pnp_cross_coupling_blocks_are_exact_zeros_under_fd():
    """The (c+, c-) and (c-, c+) blocks stay exact zeros under the reference's finite differences -- the product stores 7 planes."""
    m, p = case("cylinder")
    u = np.random.RandomState(1).uniform(-1, 1, 3 * m.nv)
    rp, col, val = ora.jacobian(m, p, ora.OP_PNP, u, mode=0)
    A = csr(rp, col, val).toarray()
    nv = m.nv
    assert np.all(A[nv:2 * nv, 2 * nv:] == 0.0) and np.all(A[2 * nv:, nv:2 * nv] == 0.0)


def test_dirichlet_rows_and_columns():
    m, p = case("pore_small")
    u = np.random.RandomState(2).uniform(-1, 1, 3 * m.nv)
    rp, col, val = ora.jacobian(m, p, ora.OP_PNP, u)
    d = ora.dirichlet(m, p, 3)
    assert d.reshape(3, -1).sum(1).tolist() == [20, 20, 20]
    A = csr(rp, col, val).toarray()
    for i in np.where(d)[0]:
        row = A[i].copy(); row[i] -= 1.0
        assert not row.any() and not np.delete(A[:, i], i).any()
    r = ora.residual(m, p, ora.OP_PNP, u)
    assert not r[d].any()


def test_one_wall_pb_against_1d_boundary_value_problem():
    """PB between a charged wall (flux 0.1) and a grounded wall: the 2-D solve must reproduce the 1-D problem
    u'' = 8*PI*l_b*c0*sinh(u), u'(0) = +0.1 (weak form: du/dn = -j), u(5) = 0, to O(h^2)."""
    from scipy.integrate import solve_bvp
    k2 = 8 * 3.1415 * 1.0 * 0.06
    sol = solve_bvp(lambda x, y: np.vstack([y[1], k2 * np.sinh(y[0])]), lambda ya, yb: np.array([ya[1] - 0.1, yb[0]]),
                    np.linspace(0, 5, 200), np.zeros((2, 200)), tol=1e-10)
    errs = []
    for lev in (1, 2, 3):
        m, p = case("one_wall", lev)
        opts = ora.newton_opts(p, prec=ora.PREC_SSOR); opts[0], opts[2], opts[12] = 1e-10, 1e-8, 5000
        u, res = ora.newton(m, p, ora.OP_PB, np.zeros(m.nv), opts)
        assert res["converged"]
        errs.append(np.max(np.abs(u - sol.sol(m.x)[0])))
    assert errs[2] < 2e-4 * 0.0822 * 50  # |u|max = 0.082
    assert errs[0] / errs[1] > 2.5 and errs[1] / errs[2] > 2.5  # second-order convergence


def test_linear_solvers_against_dense_solve():
    m, p = case("sphere")
    c = np.full(m.nv, 0.06)
    rp, col, val = ora.jacobian(m, p, ora.OP_POISSON, np.zeros(m.nv), c, c, mode=1)
    A = csr(rp, col, val).toarray()
    b = np.random.RandomState(3).uniform(-1, 1, m.nv)
    x_ref = np.linalg.solve(A, b)
    for solver in (ora.SOLVER_BCGS, ora.SOLVER_CG):
        for prec in (ora.PREC_NONE, ora.PREC_JACOBI, ora.PREC_SSOR):
            x, info = ora.linsolve(rp, col, val, b, 1e-12, 2000, solver, prec)
            assert info["converged"] and np.linalg.norm(x - x_ref) <= 1e-8 * np.linalg.norm(x_ref)
    it_none = ora.linsolve(rp, col, val, b, 1e-8, 2000, ora.SOLVER_CG, ora.PREC_NONE)[1]["iterations"]
    it_ssor = ora.linsolve(rp, col, val, b, 1e-8, 2000, ora.SOLVER_CG, ora.PREC_SSOR)[1]["iterations"]
    assert it_ssor < it_none


def test_bcextension_interpolation_rules():
    m, p = case("pore_small")
    pb = np.random.RandomState(5).uniform(-1, 1, m.nv)
    d = ora.dirichlet(m, p, 1, 0)
    phi = ora.interpolate(m, p, 0, pb); cp = ora.interpolate(m, p, 1, pb); cm = ora.interpolate(m, p, 2, pb)
    # away from Dirichlet vertices: the PB-derived guess (dirichlet_bc.hh:99,107,115)
    assert np.array_equal(phi[~d], pb[~d])
    assert np.allclose(cp[~d], 0.06 * np.exp(-pb[~d]), rtol=1e-15) and np.allclose(cm[~d], 0.06 * np.exp(pb[~d]), rtol=1e-15)
    # on Dirichlet vertices: one of the configured boundary values
    assert set(np.unique(phi[d])) <= {0.0, 24.1} and set(np.unique(cp[d])) == {0.06}


@pytest.mark.parametrize("name", ["one_wall", "cylinder", "pore_small"])
def test_golden_solutions_regression(name):
    """The committed golden vectors (scripts/make_golden_solutions.py) are reproduced by the oracle."""
    g = np.load(os.path.join(util.GOLDEN, name + "_solution.npz"))
    m, p = case(name)
    p.sys[5] = 20000; p = ora.Params.from_flat(p.sys, p.surf)
    opts = ora.newton_opts(p, solver=ora.SOLVER_BCGS, prec=ora.PREC_SSOR); opts[0], opts[2] = 1e-11, 1e-9
    pb, r0 = ora.newton(m, p, ora.OP_PB, np.zeros(m.nv), opts)
    u0 = np.concatenate([ora.interpolate(m, p, k, pb) for k in range(3)])
    u, r = ora.newton(m, p, ora.OP_PNP, u0, opts)
    assert r0["iterations"] == int(g["pb_newton_iterations"]) and r["iterations"] == int(g["pnp_newton_iterations"])
    assert np.array_equal(pb, g["pb"]) and np.array_equal(u0, g["u0"]) and np.array_equal(u, g["u"])


@pytest.mark.parametrize("method,order", [(0, 2), (1, 1)])
def test_onestep_method_convergence_order_against_matrix_exponential(method, order):
    """Pins the restated OneStepMethod stage algebra: for the linear transport problem
    M x' = -K x the Alexander2 scheme converges with order 2 and implicit Euler with order 1 towards
    expm(-T M^-1 K) x0 (dense reference, cylinder.msh)."""
    import scipy.linalg
    import scipy.sparse as sp
    m = ora.Mesh.from_arrays(**util.load_mesh_arrays("cylinder"))
    p = ora.Params.read(util.cfg_path("cylinder"))
    nv = m.nv
    phi = 0.5 * np.sin(3 * m.x) * np.cos(2 * m.y)
    d = ora.dirichlet(m, p, 1, 1)
    rng = np.random.RandomState(5)
    x0 = rng.uniform(0.5, 1.5, nv); x0[d] = 0.0
    g = np.zeros(nv)
    rp, col, kv = ora.jacobian(m, p, ora.OP_DIFFUSION, x0, phi, valency=1.0, comp0=1, mode=1)
    _, _, mv = ora.jacobian(m, p, ora.OP_MASS, x0, comp0=1, mode=1)
    K = sp.csr_matrix((kv, col, rp), shape=(nv, nv)).toarray()
    M = sp.csr_matrix((mv, col, rp), shape=(nv, nv)).toarray()
    f = ~d
    T = 0.4
    exact = scipy.linalg.expm(-T * np.linalg.solve(M[np.ix_(f, f)], K[np.ix_(f, f)])) @ x0[f]
    errs = []
    for n in (4, 8, 16):
        x = x0.copy()
        for _ in range(n):
            x, res = ora.onestep(m, p, x, g, phi, 1.0, T / n, 1e-13, method, prec=ora.PREC_ILU0, jac_mode=1, comp0=1)
            assert all(r["converged"] for r in res)
        assert np.all(x[d] == 0.0)
        errs.append(np.linalg.norm(x[f] - exact))
    rates = [np.log2(errs[i] / errs[i + 1]) for i in range(2)]
    assert abs(rates[-1] - order) < 0.25, (errs, rates)


def test_ion_flux_pins():
    """calcIonFlux restatement (ionFlux.hh:8-96): (1) constant concentrations and a linear potential on a planar mesh: the
    surface currents sum to zero over the closed boundary (divergence theorem, exact for P1); (2) ip of c+ = exp(+phi)-like
    equilibrium vanishes to discretisation accuracy and im is insensitive to c+."""
    m = ora.Mesh.from_arrays(**util.load_mesh_arrays("cylinder"))
    p = ora.Params.read(util.cfg_path("cylinder"))
    assert p.sys[1] == 0  # planar
    phi = 0.3 * m.x - 0.2 * m.y
    c = np.full(m.nv, 0.06)
    ip, im = ora.ion_flux(m, p, phi, c, c)
    scale = 0.06 * np.hypot(0.3, 0.2) * (m.x.max() - m.x.min() + m.y.max() - m.y.min())
    assert abs(ip.sum()) <= 1e-12 * scale and abs(im.sum()) <= 1e-12 * scale
    assert np.allclose(ip, -im, rtol=0, atol=1e-13 * scale)  # ip = c grad(phi).n, im = -c grad(phi).n
    # linearity in the potential at fixed concentrations
    ip2, im2 = ora.ion_flux(m, p, 2 * phi, c, c)
    assert np.allclose(ip2, 2 * ip, rtol=1e-12, atol=1e-15) and np.allclose(im2, 2 * im, rtol=1e-12, atol=1e-15)
    # pure diffusion: phi = 0 -> ip = -grad(c+).n summed, independent of c-
    cp = 0.06 * (1 + 0.1 * m.x)
    ip3, _ = ora.ion_flux(m, p, 0 * phi, cp, c)
    ip4, _ = ora.ion_flux(m, p, 0 * phi, cp, 3 * c)
    assert np.array_equal(ip3, ip4) and abs(ip3.sum()) <= 1e-12 * scale  # grad(c+) constant: closed-surface sum vanishes


def test_write_cell_data_format(tmp_path):
    """DataWriter::writeData (datawriter.hh:45-94): one line per element, 'x y<TAB>value<TAB>gx gy', %.5e."""
    m = ora.Mesh.from_arrays(**util.load_mesh_arrays("one_wall"))
    u = 0.5 * m.x - 0.25 * m.y + 1.0
    fn = str(tmp_path / "phi.dat")
    ora.write_cell_data(m, u, fn)
    lines = open(fn).read().splitlines()
    assert len(lines) == m.nT
    import re
    num = r"-?\d\.\d{5}e[+-]\d{2}"
    pat = re.compile("^%s %s\t%s\t%s %s$" % (num, num, num, num, num))
    assert all(pat.match(l) for l in lines)
    first = [float(t) for t in lines[0].replace("\t", " ").split()]
    tri0 = m.tri[0]
    assert np.allclose(first[:2], [m.x[tri0].mean(), m.y[tri0].mean()], rtol=1e-5)
    assert np.allclose(first[2], u[tri0].mean(), rtol=1e-5) and np.allclose(first[3:], [0.5, -0.25], rtol=1e-5)


def test_one_wall_pb_against_the_references_gouy_chapman_curve():
    """The one analytic anchor the reference ships: test/one_wall_dh/one_wall.gp:4-12 plots the solution against the
    Gouy-Chapman profile g(x) = 2 ln((1 + tanh(phi0/4) e^{-kappa x}) / (1 - tanh(phi0/4) e^{-kappa x})) and its
    Debye-Hueckel limit F/kappa e^{-kappa x} (columns 1 and -3 of the DataWriter file).  The constants in the .gp file are
    stale; with the cfg's own l_b, c0, flux and the code's PI: kappa^2 = 8 PI l_b c0 and the Grahame relation
    phi0 = 2 asinh(j / (2 kappa)).  The domain is finite (phi = 0 at x = 5, kappa L = 6.1), which moves the curve by
    ~4 phi0 e^{-kappa L} = 0.2 % of phi0."""
    m, p = case("one_wall", 3)
    PI, l_b, c0 = p.sys[4], p.sys[2], p.sys[3]
    j = p.surf[0][1]                                  # coulombFlux of surface 0
    kappa = np.sqrt(8 * PI * l_b * c0)
    phi0 = 2 * np.arcsinh(j / (2 * kappa))
    g = lambda x: 2 * np.log((1 + np.tanh(0.25 * phi0) * np.exp(-kappa * x)) / (1 - np.tanh(0.25 * phi0) * np.exp(-kappa * x)))
    opts = ora.newton_opts(p, prec=ora.PREC_SSOR); opts[0], opts[2], opts[12] = 1e-10, 1e-8, 5000
    u, res = ora.newton(m, p, ora.OP_PB, np.zeros(m.nv), opts)
    assert res["converged"]
    assert abs(np.abs(u).max() - phi0) <= 5e-3 * phi0
    assert np.max(np.abs(np.abs(u) - g(m.x))) <= 5e-3 * phi0
    # Debye-Hueckel limit of the same file: j / kappa e^{-kappa x}, valid to O(phi0^2) relative
    assert np.max(np.abs(np.abs(u) - j / kappa * np.exp(-kappa * m.x))) <= (phi0 ** 2 / 8 + 5e-3) * phi0


def test_pb_solution_is_an_equilibrium_of_the_pnp_operator():
    """Cross-pin of two of the reference's operators (pb_operator.hh:116-118 against pnp_operator.hh:169-191): with
    c+- = c0 exp(+-u) the phi-row of PnpOperator is the PB residual (4 PI l_b (c+ - c-) = 8 PI l_b c0 sinh u, same
    boundary flux term) and the two Nernst-Planck rows vanish identically in the continuum, so the discrete PB solution
    makes the PNP residual O(h^2) small.  The sign is the PNP operator's own (phi -> -phi relative to BCExtension,
    which hands c+ = c0 exp(-pb) to the PNP Newton as the initial guess, dirichlet_bc.hh:106-107): with that sign the
    residual is four orders of magnitude larger and does not converge."""
    norms, wrong = [], []
    for lev in (1, 2, 3):
        m, p = case("one_wall", lev)
        opts = ora.newton_opts(p, prec=ora.PREC_SSOR); opts[0], opts[2], opts[12] = 1e-12, 1e-10, 20000
        u, res = ora.newton(m, p, ora.OP_PB, np.zeros(m.nv), opts)
        assert res["converged"]
        c0 = p.sys[3]
        r = ora.residual(m, p, ora.OP_PNP, np.concatenate([u, c0 * np.exp(u), c0 * np.exp(-u)])).reshape(3, -1)
        rw = ora.residual(m, p, ora.OP_PNP, np.concatenate([u, c0 * np.exp(-u), c0 * np.exp(u)])).reshape(3, -1)
        norms.append(np.linalg.norm(r, axis=1)); wrong.append(np.linalg.norm(rw, axis=1))
    norms, wrong = np.array(norms), np.array(wrong)
    # nodal residuals of an O(h^2)-consistent state shrink like h^3 (phi-row) resp. at least h^2 per refinement
    assert np.all(norms[:-1, 0] / norms[1:, 0] > 6.0) and np.all(norms[:-1, 1:] / norms[1:, 1:] > 3.0)
    assert np.all(wrong[-1] > 1e3 * norms[-1]) and np.all(wrong[:-1] / wrong[1:] < 2.5)


def test_pb_solution_is_a_steady_state_of_the_split_scheme():
    """Cross-pin of the operators of the reference's HEAD driver (instationary_pnp_md): with the concentrations that
    BCExtension derives from the PB field (c+ = c0 exp(-pb), c- = c0 exp(+pb), dirichlet_bc.hh:106-115) the Poisson
    residual (poisson_operator.hh:121-123) is the PB residual and both drift-diffusion residuals
    (diffusion_operator.hh:100,110, valency +-1) vanish in the continuum: discretely O(h^2)."""
    norms = []
    for lev in (1, 2, 3):
        m, p = case("one_wall", lev)
        opts = ora.newton_opts(p, prec=ora.PREC_SSOR); opts[0], opts[2], opts[12] = 1e-12, 1e-10, 20000
        pb, res = ora.newton(m, p, ora.OP_PB, np.zeros(m.nv), opts)
        cp, cm = ora.interpolate(m, p, 1, pb), ora.interpolate(m, p, 2, pb)
        assert np.allclose(cp, p.sys[3] * np.exp(-pb)) and np.allclose(cm, p.sys[3] * np.exp(pb))  # no Dirichlet c on one_wall
        rphi = ora.residual(m, p, ora.OP_POISSON, pb, cp, cm)
        rp = ora.residual(m, p, ora.OP_DIFFUSION, cp, pb, valency=1.0, comp0=1)
        rm = ora.residual(m, p, ora.OP_DIFFUSION, cm, pb, valency=-1.0, comp0=1)
        # scale: the same residuals with the opposite valency (a state that is NOT steady)
        sp = ora.residual(m, p, ora.OP_DIFFUSION, cp, pb, valency=-1.0, comp0=1)
        norms.append([np.linalg.norm(rphi), np.linalg.norm(rp), np.linalg.norm(rm), np.linalg.norm(sp)])
    norms = np.array(norms)
    assert np.all(norms[:-1, 0] / norms[1:, 0] > 6.0)          # Poisson row: h^3 in nodal terms
    assert np.all(norms[:-1, 1:3] / norms[1:, 1:3] > 3.0)      # transport rows: at least h^2
    assert np.all(norms[-1, 1:3] < 1e-3 * norms[-1, 3])
