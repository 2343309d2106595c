"""GPU parity tests of the quadratic- and cubic-element path (pnp_space_set_degree(ctx, 2 | 3): the reference's -DPDEGREE=2 / 3
programs, src/Makefile.am:54-110) against the oracle's Pk restatement (oracle/pnp_oracle_p2.hpp) on the same inputs, through the
C ABI.  Cubic elements run with the reference drivers' quadrature order (0: assembly parity only -- that rule makes the cubic
matrices indefinite, tests/test_oracle_p2.py) and with order 5 (solves).

Bars as for linear elements: dof numbering, constraints and pattern bit-exact; residual and Jacobian entries within 1e-12 of
the entry's scale (sum of |element contributions|); equal Newton iteration counts, converged fields within 1e-8 relative L2."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

import util
from oracle import binding as ora

pytestmark = pytest.mark.gpu

OPS = [ora.OP_PB, ora.OP_POISSON, ora.OP_DIFFUSION, ora.OP_MASS, ora.OP_PNP]
TOL = 1e-12
# exact-derivative mode (not in the reference): the PB entry is stiffness + kappa^2 cosh(u) * mass summed INSIDE one element, and
# the device's and glibc's cosh differ in the last bit; where the two parts cancel, the per-entry scale `ab` (sum over elements
# of |element contribution|) does not see the larger terms the rounding acted on.  Same bar as for linear elements.
TOL_EXACT = 1e-10


def _capi():
    from dune_pnp_b200 import capi
    return capi


def make_ctx(name, levels=0, degree=2):
    capi = _capi()
    a = util.load_mesh_arrays(name)
    c = capi.Context(0)
    c.mesh_set(**a)
    c.params_read(util.cfg_path(name))
    if levels:
        c.mesh_refine(levels)
    c.space_set_degree(degree)
    c.mesh_finalize(True)
    m = ora.Mesh.from_arrays(**a).refine(levels)
    p = ora.Params.read(util.cfg_path(name))
    return c, m, p, ora.P2(m, p, degree)


def rel_err(got, want, scale):
    e = np.abs(got - want)
    out = np.where(scale > 0, e / np.maximum(scale, 1e-300), e)
    return out.max() if out.size else 0.0


def _state(P, op, seed=0):
    rng = np.random.RandomState(seed)
    F = ora.nfields(op)
    return rng.uniform(-1, 1, F * P.nd), rng.uniform(0, 1, P.nd), rng.uniform(0, 1, P.nd)


def _operator(c, op, a0, a1, valency, comp0=0, intorder=0):
    capi = _capi()
    h = c.operator(op, comp0)
    if intorder:
        c.operator_set_intorder(h, intorder)
    if op == capi.OP_POISSON:
        c.operator_set_coefficient(h, 0, c.vec(1, a0)); c.operator_set_coefficient(h, 1, c.vec(1, a1))
    if op == capi.OP_DIFFUSION:
        c.operator_set_coefficient(h, 0, c.vec(1, a0)); c.operator_set_valency(h, valency)
    return h


@pytest.mark.parametrize("degree", [2, 3])
@pytest.mark.parametrize("name,levels", [(n, 0) for n in util.MESHES] + [("pore_small", 1)])
def test_space_pattern_and_constraints_bit_exact(name, levels, degree):
    capi = _capi()
    c, m, p, P = make_ctx(name, levels, degree)
    s = c.space_sizes()
    assert s == dict(degree=degree, n_edges=P.nE, ndof=P.nd) and c.space_offsets() == (P.eoff, P.voff)
    va, vb = c.space_edges()
    assert np.array_equal(va, P.eva) and np.array_equal(vb, P.evb)
    for op, F, comp0 in ((capi.OP_PB, 1, 0), (capi.OP_DIFFUSION, 1, 1), (capi.OP_MASS, 1, 2), (capi.OP_PNP, 3, 0)):
        h = c.operator(op, comp0)
        assert np.array_equal(c.constraints(h, F), P.dirichlet(F, comp0))
        rp, col = c.pattern(h, F)
        rp_o, col_o = P.pattern(F, comp0)
        assert np.array_equal(rp, rp_o) and np.array_equal(col, col_o)


DEG_ORDER = [(2, 0), (2, 5), (3, 0), (3, 5)]   # (degree, intorder): 0 = the reference drivers' order


@pytest.mark.parametrize("degree,intorder", DEG_ORDER)
@pytest.mark.parametrize("op", OPS)
@pytest.mark.parametrize("name,levels", [(n, 0) for n in util.MESHES] + [("pore_small", 1)])
def test_residual_parity(name, levels, op, degree, intorder):
    if (degree, intorder) != (2, 0) and name not in ("pore", "cylinder", "one_wall"):
        pytest.skip("the full mesh list runs for the reference configuration")
    c, m, p, P = make_ctx(name, levels, degree)
    F = ora.nfields(op)
    u, a0, a1 = _state(P, op)
    h = _operator(c, op, a0, a1, -1.0, intorder=intorder)
    vu, vr = c.vec(F, u), c.vec(F)
    c.residual(h, vu, vr)
    r = c.download(vr, F)
    r_o, ab = P.residual(op, u, a0, a1, valency=-1.0, want_abs=True, intorder=intorder or -1)
    assert rel_err(r, r_o, ab) <= TOL
    d = P.dirichlet(F, 0)
    assert not r[d].any() and not r_o[d].any()
    assert np.array_equal(c.download(vu, F), u)   # the upload / download pair is the identity in this numbering


@pytest.mark.parametrize("degree,intorder", DEG_ORDER)
@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("op", OPS)
@pytest.mark.parametrize("name,levels", [("one_wall", 0), ("cylinder", 0), ("pore", 0), ("pore_small", 1)])
def test_jacobian_parity(name, levels, op, mode, degree, intorder):
    if (degree, intorder) != (2, 0) and name not in ("one_wall", "pore"):
        pytest.skip("the full mesh list runs for the reference configuration")
    c, m, p, P = make_ctx(name, levels, degree)
    F = ora.nfields(op)
    u, a0, a1 = _state(P, op, seed=3)
    h = _operator(c, op, a0, a1, -1.0, intorder=intorder)
    vu, A = c.vec(F, u), c.matrix(h)
    c.jacobian(h, vu, A, mode, 1e-11)
    rp, col, val_o, ab = P.jacobian(op, u, a0, a1, valency=-1.0, mode=mode, eps=1e-11, want_abs=True, intorder=intorder or -1)
    val = c.matrix_values(h, A, len(col))
    assert rel_err(val, val_o, ab) <= (TOL if mode == 0 else TOL_EXACT)
    # same element order, same operation order, no FMA contraction: most entries reproduce the oracle bit for bit
    assert np.mean(val == val_o) > (0.9 if mode == 0 or op != ora.OP_PB else 0.0)
    # constrained rows are trivial
    d = P.dirichlet(F, 0)
    rows = np.repeat(np.arange(len(rp) - 1), np.diff(rp))
    assert np.all(val[d[rows]] == 1.0)


@pytest.mark.parametrize("op", [ora.OP_PB, ora.OP_PNP])
def test_spmv_and_dots(op):
    c, m, p, P = make_ctx("pore_small", 1)
    F = ora.nfields(op)
    u, a0, a1 = _state(P, op, seed=5)
    h = _operator(c, op, a0, a1, 1.0)
    vu, A = c.vec(F, u), c.matrix(h)
    c.jacobian(h, vu, A, 1, 1e-11)
    rp, col = c.pattern(h, F)
    J = sp.csr_matrix((c.matrix_values(h, A, len(col)), col, rp))
    x = np.random.RandomState(1).uniform(-1, 1, F * P.nd)
    vx, vy = c.vec(F, x), c.vec(F)
    c.spmv(A, vx, vy)
    y = c.download(vy, F)
    scale = abs(J) @ np.abs(x)
    assert rel_err(y, J @ x, scale) <= 1e-14
    assert abs(c.dot(vx, vy) - x @ y) <= 1e-12 * np.abs(x) @ np.abs(y)
    assert abs(c.norm(vy) - np.linalg.norm(y)) <= 1e-13 * np.linalg.norm(y)


@pytest.mark.parametrize("kind,prec", [(0, 0), (0, 1), (1, 1)])
def test_linear_solvers(kind, prec):
    """BiCGSTAB / CG with Richardson and Jacobi on the (symmetric positive definite) P2 Poisson matrix."""
    capi = _capi()
    c, m, p, P = make_ctx("cylinder")
    z = np.zeros(P.nd)
    h = _operator(c, capi.OP_POISSON, z, z, 1.0)
    vu, A = c.vec(1, z), c.matrix(h)
    c.jacobian(h, vu, A, 1, 1e-11)
    rp, col = c.pattern(h, 1)
    J = sp.csr_matrix((c.matrix_values(h, A, len(col)), col, rp))
    b = np.random.RandomState(2).uniform(-1, 1, P.nd)
    vz, vb = c.vec(1), c.vec(1, b)
    res = c.solve(c.solver(kind, prec, 20000), A, vz, vb, 1e-10)
    assert res.converged
    zsol = c.download(vz, 1)
    assert np.linalg.norm(J @ zsol - b) <= 2e-10 * np.linalg.norm(b)
    assert np.linalg.norm(zsol - spla.spsolve(J.tocsc(), b)) <= 1e-6 * np.linalg.norm(zsol)


@pytest.mark.parametrize("degree", [2, 3])
@pytest.mark.parametrize("prec,steps", [(2, 1), (2, 3), (3, 1)])
@pytest.mark.parametrize("name,levels,op", [("pore_small", 1, ora.OP_PB), ("pore", 0, ora.OP_PNP), ("cylinder", 0, ora.OP_DIFFUSION)])
def test_ssor_ilu0_application_equals_sequential_sweep(name, levels, op, prec, steps, degree):
    """SeqSSOR(n) / SeqILU0 in the row order of the P2 matrix: the level-scheduled device sweep does every row's operations in
    the sequential sweep's order (ascending columns, no FMA) -- bit-identical to the CPU sweep."""
    capi = _capi()
    c, m, p, P = make_ctx(name, levels, degree)
    F = ora.nfields(op)
    u, a0, a1 = _state(P, op, seed=9)
    if op == ora.OP_PNP:
        u[P.nd:] = 0.06 * (1 + 0.1 * u[P.nd:])
    h = _operator(c, op, a0, a1, 1.0, intorder=5 if degree == 3 else 0)
    vu, A = c.vec(F, 0.3 * u if op != ora.OP_PNP else u), c.matrix(h)
    c.jacobian(h, vu, A, 1, 1e-11)
    rp, col = c.pattern(h, F)
    val = c.matrix_values(h, A, len(col))
    d = np.random.RandomState(8).uniform(-1, 1, F * P.nd)
    s = c.solver(capi.SOLVER_BCGS, prec, 100, steps)
    vd, vv = c.vec(F, d), c.vec(F)
    c.precond_apply(s, A, vd, vv)
    v, v_o = c.download(vv, F), ora.prec_apply(rp, col, val, d, prec, steps)
    assert np.linalg.norm(v - v_o) <= 1e-13 * np.linalg.norm(v_o)
    assert np.mean(v == v_o) > 0.9
    nlev = c.solver_get(s, "ssor_levels" if prec == 2 else "ilu0_levels")
    assert 1 <= nlev <= 400


@pytest.mark.parametrize("kind,prec", [(0, 2), (1, 2), (0, 3), (1, 3)])
def test_krylov_with_ssor_ilu0_matches_oracle_iteration_counts(kind, prec):
    capi = _capi()
    c, m, p, P = make_ctx("sphere", 1)
    z = np.zeros(P.nd)
    h = _operator(c, capi.OP_POISSON, z, z, 1.0)
    vu, A = c.vec(1, z), c.matrix(h)
    c.jacobian(h, vu, A, 1, 1e-11)
    rp, col = c.pattern(h, 1)
    val = c.matrix_values(h, A, len(col))
    b = np.random.RandomState(2).uniform(-1, 1, P.nd)
    vz, vb = c.vec(1), c.vec(1, b)
    res = c.solve(c.solver(kind, prec, 2000), A, vz, vb, 1e-10)
    z_o, res_o = ora.linsolve(rp, col, val, b, 1e-10, 2000, kind, prec)
    assert res.converged and res_o["converged"]
    # same sweep bit for bit; the Krylov scalars feel the summation order of the device's dot products (tree) vs the CPU's (sequential)
    assert abs(res.iterations - res_o["iterations"]) <= max(3, res_o["iterations"] // 10)
    assert np.linalg.norm(c.download(vz, 1) - z_o) <= 1e-7 * np.linalg.norm(z_o)


@pytest.mark.parametrize("degree", [2, 3])
@pytest.mark.parametrize("name", ["one_wall", "cylinder", "pore_small"])
@pytest.mark.parametrize("mode", [0, 1])
def test_newton_pb_matches_oracle(name, mode, degree):
    capi = _capi()
    c, m, p, P = make_ctx(name, 0, degree)
    h = c.operator(capi.OP_PB, 0)
    intorder = 5 if degree == 3 else 0
    if intorder:
        c.operator_set_intorder(h, intorder)
    s = c.solver(capi.SOLVER_BCGS, capi.PREC_JACOBI, 20000)
    vu = c.vec(1)
    tight = mode == 1   # (FD noise moves iteration counts at tight tolerances: the tight comparison uses the exact derivative)
    kw = dict(reduction=1e-11, min_linear_reduction=1e-9) if tight else {}
    st, res = c.newton(h, vu, s, c.newton_opts(jac_mode=mode, **kw))
    opts = ora.newton_opts(p, solver=ora.SOLVER_BCGS, prec=ora.PREC_JACOBI, jac_mode=mode)
    opts[12] = 20000
    if tight:
        opts[0], opts[2] = 1e-11, 1e-9
    u_o, res_o = P.newton(ora.OP_PB, np.zeros(P.nd), opts, intorder=intorder or -1)
    assert res.converged and res_o["converged"]
    assert res.iterations == res_o["iterations"]
    u = c.download(vu, 1)
    tol = 1e-8 if tight else 10 * p.sys[7]
    assert np.linalg.norm(u - u_o) <= tol * np.linalg.norm(u_o) + 1e-14
    assert abs(res.first_defect - res_o["first_defect"]) <= 1e-12 * res_o["first_defect"]


@pytest.mark.parametrize("degree", [2, 3])
@pytest.mark.parametrize("comp", [0, 1, 2])
@pytest.mark.parametrize("name", ["pore", "sphere"])
def test_interpolate_bcext(name, comp, degree):
    c, m, p, P = make_ctx(name, 0, degree)
    pb = np.sin(0.1 * P.x) * np.cos(0.07 * P.y)
    vpb, vo = c.vec(1, pb), c.vec(1)
    c.interpolate_bcext(comp, vpb, vo)
    got, want = c.download(vo, 1), P.interpolate(comp, pb)
    assert np.allclose(got, want, rtol=1e-15, atol=0)
    d = P.dirichlet(1, comp)
    assert np.array_equal(got[d], want[d])


@pytest.mark.parametrize("degree,prec", [(2, 3), (2, 2), (3, 3)])
def test_newton_pnp_from_pb_matches_oracle(prec, degree):
    """The stationary program with quadratic elements (stationary_pnp_from_pb.hh:105-185, :344-360): PB Newton solve, the
    Boltzmann start from interpolate(BCExtension), then the coupled 3-field Newton solve -- exact derivative, the reference's
    default backend BiCGSTAB + SSOR(1) (~500 BiCGSTAB iterations per Newton step on this system) and BiCGSTAB + ILU0 (~50)."""
    capi = _capi()
    c, m, p, P = make_ctx("pore_small", 0, degree)
    io = 5 if degree == 3 else 0
    hpb = _operator(c, capi.OP_PB, None, None, 1.0, intorder=io)
    s = c.solver(capi.SOLVER_BCGS, prec, 50000, 1)
    vpb = c.vec(1)
    st, r0 = c.newton(hpb, vpb, s, c.newton_opts(jac_mode=1, reduction=1e-11, min_linear_reduction=1e-9))
    v = [c.vec(1) for _ in range(3)]
    for k in range(3):
        c.interpolate_bcext(k, vpb, v[k])
    vu = c.vec(3)
    c.pack3(vu, *v)
    hp = _operator(c, capi.OP_PNP, None, None, 1.0, intorder=io)
    st, res = c.newton(hp, vu, s, c.newton_opts(jac_mode=1, reduction=1e-10, min_linear_reduction=1e-9))
    # oracle
    opts = ora.newton_opts(p, solver=ora.SOLVER_BCGS, prec=prec, jac_mode=1); opts[12] = 50000
    opts[0], opts[2] = 1e-11, 1e-9
    pb_o, _ = P.newton(ora.OP_PB, np.zeros(P.nd), opts, intorder=io or -1)
    u0 = np.concatenate([P.interpolate(k, pb_o) for k in range(3)])
    assert np.linalg.norm(np.concatenate([c.download(x, 1) for x in v]) - u0) <= 1e-8 * np.linalg.norm(u0)
    opts[0] = 1e-10
    u_o, res_o = P.newton(ora.OP_PNP, u0, opts, intorder=io or -1)
    assert res.converged and res_o["converged"] and res.iterations == res_o["iterations"]
    u = c.download(vu, 3)
    assert np.linalg.norm(u - u_o) <= 1e-8 * np.linalg.norm(u_o)
    # pack / extract are the field blocks
    w = c.vec(1)
    for k in range(3):
        c.extract(vu, k, w)
        assert np.array_equal(c.download(w, 1), u[k * P.nd:(k + 1) * P.nd])


def test_onestep_implicit_euler_is_the_linear_algebra_it_claims():
    """One implicit-Euler step of the transport equation (OneStepGridOperator<DiffusionOperator, DiffusionTOperator>,
    instationary_pnp_from_pb_md.hh:368-391) with quadratic elements: both operators are linear, so the new state solves
    (M + dt A) x = M x_old on the free dofs with the boundary values on the constrained ones; checked with a direct solve
    on the exported matrices."""
    capi = _capi()
    c, m, p, P = make_ctx("pore_small")
    rng = np.random.RandomState(4)
    phi = 0.5 * np.sin(3 * P.x) * np.cos(2 * P.y)
    g = P.interpolate(1, phi)
    d = P.dirichlet(1, 1)
    x0 = g * (1 + 0.1 * rng.uniform(-1, 1, P.nd)); x0[d] = g[d]
    dt = 0.05
    h0 = _operator(c, capi.OP_DIFFUSION, phi, None, 1.0, comp0=1)
    h1 = c.operator(capi.OP_MASS, 1)
    s = c.solver(capi.SOLVER_BCGS, capi.PREC_JACOBI, 20000)
    vx0, vg, vx1 = c.vec(1, x0), c.vec(1, g), c.vec(1)
    res = c.onestep(h0, h1, s, dt, vx0, vg, vx1, 1e-12, capi.TIME_IMPLICIT_EULER, 1)
    assert len(res) == 1 and res[0].converged
    x1 = c.download(vx1, 1)
    assert np.array_equal(x1[d], g[d]) and np.array_equal(c.download(vx0, 1), x0)
    rp, col, a = P.jacobian(ora.OP_DIFFUSION, x0, phi, valency=1.0, mode=1, comp0=1)
    _, _, b = P.jacobian(ora.OP_MASS, x0, comp0=1, mode=1)
    A, M = sp.csr_matrix((a, col, rp)), sp.csr_matrix((b, col, rp))
    # free rows: (M + dt A)(x1) = M x0 where constrained columns (dropped from the pattern) carry g through the residual
    r1 = P.residual(ora.OP_MASS, x1, comp0=1) - P.residual(ora.OP_MASS, x0, comp0=1) + dt * P.residual(ora.OP_DIFFUSION, x1, phi, valency=1.0, comp0=1)
    scale = abs(M + dt * A) @ np.abs(x1)
    assert np.max(np.abs(r1[~d]) / scale[~d]) <= 1e-9


def test_solve_on_an_externally_assembled_matrix():
    capi = _capi()
    c, m, p, P = make_ctx("cylinder")
    u, a0, a1 = _state(P, ora.OP_PB, seed=7)
    rp, col, val = P.jacobian(ora.OP_PB, 0.1 * u, mode=1)
    h = c.operator(capi.OP_PB, 0)
    A = c.matrix(h)
    c.matrix_set_csr(h, A, rp, col, val)
    assert np.array_equal(c.matrix_values(h, A, len(col)), val)
    b = np.random.RandomState(3).uniform(-1, 1, P.nd)
    vz, vb = c.vec(1), c.vec(1, b)
    assert c.solve(c.solver(capi.SOLVER_BCGS, capi.PREC_JACOBI, 20000), A, vz, vb, 1e-10).converged
    J = sp.csr_matrix((val, col, rp))
    assert np.linalg.norm(J @ c.download(vz, 1) - b) <= 2e-10 * np.linalg.norm(b)
    bad = col.copy(); bad[1] += 1
    with pytest.raises(capi.PnpError):
        c.matrix_set_csr(h, A, rp, bad, val)


def test_errors_are_reported():
    capi = _capi()
    c, m, p, P = make_ctx("one_wall")
    with pytest.raises(capi.PnpError) as e:
        c.space_set_degree(1)                       # after finalize
    assert e.value.status == 8                      # PNP_E_ARG
    h = c.operator(capi.OP_PB, 0)
    vu = c.vec(1)
    # the p-multigrid re-discretises the last assembled Jacobian: an imported matrix is not one
    rp_, col_, val_ = P.jacobian(ora.OP_PB, np.zeros(P.nd), mode=1)
    Ai = c.matrix(h)
    c.matrix_set_csr(h, Ai, rp_, col_, val_)
    with pytest.raises(capi.PnpError) as e:
        c.solve(c.solver(capi.SOLVER_BCGS, capi.PREC_AMG, 100, 1), Ai, c.vec(1), c.vec(1, np.ones(P.nd)), 1e-8)
    assert e.value.status == 8 and "last assembled Jacobian" in str(e.value)
    with pytest.raises(capi.PnpError):
        c.carry_set([vu])                           # refinement carry-over is built for linear elements
    with pytest.raises(capi.PnpError):
        c.residual(h, c.vec(3), vu)                 # field count mismatch
    # a parameter change rebuilds the constraints and with them the patterns: matrices are detached, not left dangling
    A = c.matrix(h)
    c.jacobian(h, vu, A, 1, 1e-11)
    vx, vy = c.vec(1), c.vec(1)
    c.spmv(A, vx, vy)
    c.params_read(util.cfg_path("one_wall"))
    with pytest.raises(capi.PnpError):
        c.spmv(A, vx, vy)
    c.jacobian(h, vu, A, 1, 1e-11)
    c.spmv(A, vx, vy)
    c2 = capi.Context(0)
    with pytest.raises(capi.PnpError):
        c2.space_set_degree(4)
    with pytest.raises(capi.PnpError):
        c.operator_set_intorder(h, 4)               # tabulated: the reference drivers' order and 5
    c3, m3, p3 = __import__("test_gpu_parity").make_ctx("one_wall")
    with pytest.raises(capi.PnpError):
        c3.operator_set_intorder(c3.operator(capi.OP_PB, 0), 5)   # linear elements keep the drivers' order


# ---- the time loop's diagnostics with quadratic functions ----
@pytest.mark.parametrize("degree", [2, 3])
def test_outputs_match_oracle(tmp_path, degree):
    """calcIonFlux, DataWriter::writeData and the VTK vertex data for quadratic functions: fields and gradients are basis sums
    over the element's 6 dofs at the face / element centre; VTK vertex data are the vertex dofs."""
    c, m, p, P = make_ctx("pore", 0, degree)
    phi = np.cos(0.3 * P.x) * np.sin(0.2 * P.y); cp = 0.06 * np.exp(-phi); cm = 0.06 * np.exp(phi)
    vphi, vcp, vcm = c.vec(1, phi), c.vec(1, cp), c.vec(1, cm)
    ip, im = c.ion_flux(vphi, vcp, vcm)
    ip_o, im_o = P.ion_flux(phi, cp, cm)
    ip_a, im_a = P.ion_flux(np.abs(phi), np.abs(cp), np.abs(cm))
    scale = np.abs(ip_o).max() + np.abs(im_o).max() + np.abs(ip_a).max() + np.abs(im_a).max()
    assert np.all(np.abs(ip - ip_o) <= 1e-11 * scale) and np.all(np.abs(im - im_o) <= 1e-11 * scale) and np.any(ip_o != 0)
    # a quadratic field differs from its vertex interpolant at the face centres: the P1 formula would not pass
    ip1, _ = ora.ion_flux(m, p, phi[P.voff:], cp[P.voff:], cm[P.voff:])
    assert np.max(np.abs(ip1 - ip_o)) > 1e-6 * scale
    c.write_cell_data(vphi, str(tmp_path / "gpu.dat"))
    P.write_cell_data(phi, str(tmp_path / "ora.dat"))
    a = open(tmp_path / "gpu.dat").read().splitlines(); b = open(tmp_path / "ora.dat").read().splitlines()
    assert len(a) == len(b) == m.nT
    A = np.array([[float(t) for t in l.split()] for l in a]); B = np.array([[float(t) for t in l.split()] for l in b])
    assert np.allclose(A, B, rtol=2e-5, atol=1e-12)
    assert sum(la == lb for la, lb in zip(a, b)) >= 0.99 * len(a)
    for ascii_ in (True, False):
        g, o = str(tmp_path / ("gpu%d" % ascii_)), str(tmp_path / ("ora%d" % ascii_))
        c.write_vtk(g, [vphi, vcp], ["phi", "cp"], ascii=ascii_)
        ora.write_vtk(m, o, [phi[P.voff:], cp[P.voff:]], ["phi", "cp"], ascii=ascii_)
        assert open(g + ".vtu", "rb").read() == open(o + ".vtu", "rb").read()


@pytest.mark.parametrize("degree", [2, 3])
def test_driver_instationary_pnp_md_with_quadratic_elements(tmp_path, degree):
    """The reference binary built with -DPDEGREE=2 (src/Makefile.am:57-60: dune_pnp_BCGS_SSORk_2): PnpSolverMain::run on
    one_wall, three time steps with file output; the printed norms equal the same loop driven through the C ABI, the files
    are what the oracle's writers make of the fields."""
    import subprocess
    from test_gpu_parity import _build_example
    capi = _capi()
    a = util.load_mesh_arrays("one_wall")
    util.write_gmsh(str(tmp_path / "one_wall.msh"), a)
    cfg = open(util.cfg_path("one_wall")).read()
    if degree == 3:   # the cfg's 50 Krylov iterations are not enough for the cubic system on the refined mesh
        cfg = cfg.replace("linearSolverIterations=50", "linearSolverIterations=2000")
        assert "linearSolverIterations=2000" in cfg
    (tmp_path / "one_wall.cfg").write_text(cfg)
    io = 5 if degree == 3 else 0
    exe = _build_example("instationary_pnp_md", tmp_path, ("-DPDEGREE=%d" % degree,) + (("-DPNP_INTORDER=5",) if io else ()))
    out = subprocess.run([exe, "one_wall.cfg", "1", "3", "files"], cwd=str(tmp_path), capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = [l.split() for l in out.stdout.splitlines() if l.startswith("step")]
    assert len(lines) == 3
    c, m, p, P = make_ctx("one_wall", 1, degree)
    ls = c.solver(capi.SOLVER_BCGS, capi.PREC_SSOR, 2000 if degree == 3 else int(p.sys[5]), 1)
    vpb = c.vec(1)
    c.newton(_operator(c, capi.OP_PB, None, None, 1.0, intorder=io), vpb, ls, c.newton_opts(jac_mode=0))
    uphi, ucp, ucm, cpB, cmB, new = (c.vec(1) for _ in range(6))
    c.interpolate_bcext(0, vpb, uphi)
    c.interpolate_bcext(1, vpb, ucp); c.interpolate_bcext(1, vpb, cpB)
    c.interpolate_bcext(2, vpb, ucm); c.interpolate_bcext(2, vpb, cmB)
    hphi = c.operator(capi.OP_POISSON, 0)
    c.operator_set_coefficient(hphi, 0, ucp); c.operator_set_coefficient(hphi, 1, ucm)
    h0p, h0m = c.operator(capi.OP_DIFFUSION, 1), c.operator(capi.OP_DIFFUSION, 1)
    c.operator_set_coefficient(h0p, 0, uphi); c.operator_set_valency(h0p, 1.0)
    c.operator_set_coefficient(h0m, 0, uphi); c.operator_set_valency(h0m, -1.0)
    h1 = c.operator(capi.OP_MASS, 1)
    if io:
        for h in (hphi, h0p, h0m, h1):
            c.operator_set_intorder(h, io)
    for i in range(3):
        c.onestep(h0p, h1, ls, p.sys[11], ucp, cpB, new, 1e-5); c.vec_copy(ucp, new)
        c.onestep(h0m, h1, ls, p.sys[11], ucm, cmB, new, 1e-5); c.vec_copy(ucm, new)
        c.slp(hphi, uphi, ls, 1e-10)
        got = [float(lines[i][k]) for k in (6, 8, 10)]
        want = [c.norm(uphi), c.norm(ucp), c.norm(ucm)]
        assert np.allclose(got, want, rtol=1e-9), (i, got, want)
    assert c.space_sizes()["ndof"] == P.nd and len(c.download(uphi, 1)) == P.nd
    P.write_cell_data(c.download(ucp, 1), str(tmp_path / "ora_cp.dat"))
    a_, b_ = (tmp_path / "cp003.dat").read_text().splitlines(), (tmp_path / "ora_cp.dat").read_text().splitlines()
    assert len(a_) == len(b_) == m.nT and sum(x == y for x, y in zip(a_, b_)) >= 0.99 * len(a_)
    cur = (tmp_path / "current.dat").read_text().splitlines()
    ip_o, im_o = P.ion_flux(c.download(uphi, 1), c.download(ucp, 1), c.download(ucm, 1))
    last = [float(t) for t in cur[-1].split()]
    assert np.allclose(last[1::4], ip_o, rtol=1e-5, atol=1e-12) and np.allclose(last[3::4], im_o, rtol=1e-5, atol=1e-12)


@pytest.mark.parametrize("degree,mode", [(2, 0), (3, 1)])
def test_jacobian_scratch_chunks_give_the_same_bits(degree, mode):
    """The element matrices pass through a scratch block of bounded size (2 GB by default): chunks of elements in ascending order
    leave every entry's summation order unchanged."""
    capi = _capi()
    c, m, p, P = make_ctx("pore_small", 0, degree)
    u, a0, a1 = _state(P, ora.OP_PNP, seed=4)
    h = _operator(c, ora.OP_PNP, a0, a1, 1.0, intorder=5 if degree == 3 else 0)
    vu, A = c.vec(3, u), c.matrix(h)
    rp, col = c.pattern(h, 3)
    c.jacobian(h, vu, A, mode, 1e-11)
    ref = c.matrix_values(h, A, len(col))
    try:
        for chunk in (1, 37, m.nT - 1):
            capi.tune("p2_chunk", chunk)
            c.jacobian(h, vu, A, mode, 1e-11)
            assert np.array_equal(c.matrix_values(h, A, len(col)), ref)
    finally:
        capi.tune("p2_chunk", 0)


@pytest.mark.parametrize("degree,method,mode", [(2, 0, 0), (2, 1, 1), (3, 0, 1)])
def test_onestep_transport_matches_oracle(degree, method, mode):
    """One time step of the split scheme's ion transport (instationary_pnp_from_pb_md.hh:421-425) on the Pk spaces: SDIRK stages,
    each a StationaryLinearProblemSolver on a*M + b*dt*J0 with BiCGSTAB + ILU0, against the oracle's OneStepMethod."""
    capi = _capi()
    c, m, p, P = make_ctx("pore_small", 1, degree)
    io = 5 if degree == 3 else 0
    rng = np.random.RandomState(11)
    phi = 0.5 * np.sin(3 * P.x) * np.cos(2 * P.y)
    g = P.interpolate(1, phi)
    d = P.dirichlet(1, 1)
    x0 = g * (1 + 0.1 * rng.uniform(-1, 1, P.nd)); x0[d] = g[d]
    dt = p.sys[11]
    h0 = _operator(c, capi.OP_DIFFUSION, phi, None, 1.0, comp0=1, intorder=io)
    h1 = _operator(c, capi.OP_MASS, None, None, 1.0, comp0=1, intorder=io)
    s = c.solver(capi.SOLVER_BCGS, capi.PREC_ILU0, 5000, 1)
    vx0, vg, vx1 = c.vec(1, x0), c.vec(1, g), c.vec(1)
    res = c.onestep(h0, h1, s, dt, vx0, vg, vx1, 1e-10, method, mode)
    x1 = c.download(vx1, 1)
    x1_o, res_o = P.onestep(x0, g, phi, 1.0, dt, 1e-10, method, prec=ora.PREC_ILU0, maxit=5000, jac_mode=mode, comp0=1, intorder=io or -1)
    assert len(res) == len(res_o)
    for a, b in zip(res, res_o):
        assert a.converged and b["converged"] and abs(a.iterations - b["iterations"]) <= 1
    tol = 1e-6 if mode == 1 else 1e-3   # (FD Jacobian: 1e-5 relative noise in the matrix, amplified through the stage solves)
    assert np.linalg.norm(x1 - x1_o) <= tol * np.linalg.norm(x1_o - x0) + 1e-12 * np.linalg.norm(x1_o)
    assert np.array_equal(x1[d], g[d]) and np.array_equal(c.download(vx0, 1), x0)


@pytest.mark.parametrize("degree,op,kind", [(2, ora.OP_PB, 1), (3, ora.OP_PB, 1)])
def test_p_multigrid(degree, op, kind):
    """ISTLBackend_NOVLP_CG_AMG_SSOR with -DPDEGREE=2,3 (src/Makefile.am:106-110): SSOR smoothing on the Pk matrix, coarse correction
    in the P1 space of the same mesh through the star path's multigrid (scalar operators; the 3-field system answers PNP_E_ARG).
    Bar: the preconditioned Krylov method solves the assembled system in fewer iterations than SSOR alone, with mild growth."""
    capi = _capi()
    its = {}
    for levels in (1, 2):
        c, m, p, P = make_ctx("pore", levels, degree)
        F = ora.nfields(op)
        h = _operator(c, op, None, None, 1.0, intorder=5 if degree == 3 else 0)
        u0 = np.full(F * P.nd, 0.05)
        vu, A = c.vec(F, u0), c.matrix(h)
        c.jacobian(h, vu, A, 1, 1e-11)
        b = np.random.RandomState(0).uniform(-1, 1, F * P.nd)
        b[c.constraints(h, F)] = 0.0
        s = c.solver(kind, capi.PREC_AMG, 500, 1)
        z, r = c.vec(F), c.vec(F, b)
        res = c.solve(s, A, z, r, 1e-9)
        assert res.converged
        its[levels] = res.iterations
        y = c.vec(F)
        c.spmv(A, z, y)
        assert np.linalg.norm(c.download(y, F) - b) <= 2e-9 * np.linalg.norm(b)
        if levels == 1:
            z2, r2 = c.vec(F), c.vec(F, b)
            res2 = c.solve(c.solver(kind, capi.PREC_SSOR, 20000, 1), A, z2, r2, 1e-9)
            assert res2.converged and res.iterations < res2.iterations, (res.iterations, res2.iterations)
    # (the coarse solver is ONE cycle of the aggregation multigrid on the P1 matrix: mild growth with the mesh, as for linear elements)
    assert its[2] <= 2 * its[1] + 6, its


def test_newton_pb_with_p_multigrid():
    capi = _capi()
    c, m, p, P = make_ctx("pore", 1, 2)
    h = c.operator(capi.OP_PB, 0)
    vu = c.vec(1)
    st, res = c.newton(h, vu, c.solver(capi.SOLVER_CG, capi.PREC_AMG, 500, 1), c.newton_opts(jac_mode=1, reduction=1e-10, min_linear_reduction=1e-8))
    vw = c.vec(1)
    st2, res2 = c.newton(h, vw, c.solver(capi.SOLVER_BCGS, capi.PREC_ILU0, 5000, 1), c.newton_opts(jac_mode=1, reduction=1e-10, min_linear_reduction=1e-8))
    assert res.converged and res2.converged and res.iterations == res2.iterations
    assert res.linear_iterations <= 40 * res.iterations
    u, w = c.download(vu, 1), c.download(vw, 1)
    assert np.linalg.norm(u - w) <= 1e-7 * np.linalg.norm(w)
