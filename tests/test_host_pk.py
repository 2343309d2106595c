"""CPU tests of the product's quadratic / cubic element functions (dune_pnp_b200/csrc/pnp_elem_p2.cuh, compiled for the host by
tests/host_harness with -ffp-contract=off) against the oracle's Pk restatement on ONE triangle: element vector (alpha_volume +
alpha_boundary) and element matrix (NumericalJacobianVolume and the exact derivative), for every operator, both degrees, both
quadrature orders, both orientations and vertex orders that flip the edge-dof orientation.  One element, same operation
order, no FMA: the comparison is bit for bit.  Also pins the product's node table + numbering rule (pnp_p2.cu:p2_build)
against the oracle's Space2::sdof."""
import itertools

import numpy as np
import pytest

import harness
import util
from oracle import binding as ora

OPS = [ora.OP_PB, ora.OP_POISSON, ora.OP_DIFFUSION, ora.OP_MASS, ora.OP_PNP]
FACE_V = [(0, 1), (0, 2), (1, 2)]


def one_triangle(perm, seed):
    rng = np.random.RandomState(seed)
    xy = np.array([[0.1, 0.2], [1.3, 0.5], [0.4, 1.7]]) + 0.1 * rng.uniform(-1, 1, (3, 2))
    tri = np.array([list(perm)], dtype=np.int32)
    ba = np.array([tri[0][a] for a, b in FACE_V], dtype=np.int32)
    bb = np.array([tri[0][b] for a, b in FACE_V], dtype=np.int32)
    bphys = np.array([0, 1, 2], dtype=np.int32)        # three different surfaces of pore.cfg (Dirichlet / Neumann mixes)
    return ora.Mesh.from_arrays(xy[:, 0], xy[:, 1], tri, ba, bb, bphys), xy, tri[0]


def local_to_global(P, degree, tri):
    """The product's numbering rule: bubbles, then degree-1 dofs per edge counted from the smaller end vertex, then vertices."""
    out = []
    for n in range(P.NL):
        (kind, sub, idx), _ = harness.pk_node(degree, n)
        if kind == 2:
            out.append(0)
        elif kind == 0:
            out.append(P.voff + tri[sub])
        else:
            la, lb = FACE_V[sub]
            if degree == 3 and tri[la] > tri[lb]:
                idx = 1 - idx
            out.append(P.eoff + (degree - 1) * P.tedge[0][sub] + idx)
    return np.array(out)


@pytest.mark.parametrize("degree,intorder", [(2, 0), (2, 5), (3, 0), (3, 5)])
@pytest.mark.parametrize("op", OPS)
def test_element_vector_and_matrix_equal_the_oracle_bit_for_bit(op, degree, intorder):
    p = ora.Params.read(util.cfg_path("pore"))
    F = ora.nfields(op)
    for seed, perm in enumerate(itertools.permutations(range(3))):      # ccw and cw, every edge orientation
        m, xy, tri = one_triangle(perm, seed)
        P = ora.P2(m, p, degree)
        l2g = local_to_global(P, degree, tri)
        assert sorted(l2g) == list(range(P.nd))                          # the rule is a bijection onto the element's dofs
        rng = np.random.RandomState(10 + seed)
        u = rng.uniform(-1, 1, F * P.nd)
        if op == ora.OP_PNP:
            u[P.nd:] = 0.06 * (1 + 0.3 * u[P.nd:])
        a0, a1 = rng.uniform(0, 1, P.nd), rng.uniform(0, 1, P.nd)
        xl = np.concatenate([u[k * P.nd + l2g] for k in range(F)])
        caux = np.concatenate([a0[l2g] if op in (ora.OP_POISSON, ora.OP_DIFFUSION) else np.zeros(P.NL),
                               a1[l2g] if op == ora.OP_POISSON else np.zeros(P.NL)])
        phys = [p.sys[4], p.sys[2], p.sys[3], -1.0, p.sys[1]]
        fflux = np.array([[p.surf[f, 3 * c + 1] for c in range(3)] for f in range(3)]).ravel()
        fdir = [sum(int(p.surf[f, 3 * c] == 0) << c for c in range(3)) for f in range(3)]
        d = P.dirichlet(F, 0)
        glob = np.concatenate([k * P.nd + l2g for k in range(F)])
        r_o = P.residual(op, u, a0, a1, valency=-1.0, intorder=intorder or -1)
        for mode in (0, 1):
            r, A = harness.pk_element(degree, op, intorder, xy[tri].ravel(), phys, xl, caux, [1, 1, 1], fflux, fdir, 0, mode, 1e-11)
            free = ~d[glob]
            assert np.array_equal(r[free], r_o[glob][free]) and not r_o[glob][~free].any()
            rp, col, val = P.jacobian(op, u, a0, a1, valency=-1.0, mode=mode, eps=1e-11, intorder=intorder or -1)
            import scipy.sparse as sp
            J = sp.csr_matrix((val, col, rp), shape=(F * P.nd, F * P.nd)).toarray()
            Jl = J[np.ix_(glob, glob)]
            mask = np.outer(free, free)
            if mode == 0 or op != ora.OP_PB:
                assert np.array_equal(A[mask], Jl[mask])
            else:    # (the exact PB derivative calls cosh: the same libm here, but not a pinned operation sequence)
                assert np.allclose(A[mask], Jl[mask], rtol=1e-14, atol=0)
            assert np.array_equal(np.diag(Jl)[~free], np.ones((~free).sum()))


def test_node_tables():
    for degree, nl in ((2, 6), (3, 10)):
        kinds = [harness.pk_node(degree, n)[0][0] for n in range(nl)]
        assert kinds.count(0) == 3 and kinds.count(1) == 3 * (degree - 1) and kinds.count(2) == (1 if degree == 3 else 0)
        for n in range(nl):                                              # the node is where its basis function is 1
            (_, _, _), (x, y) = harness.pk_node(degree, n)
            phi, _ = ora.p2_basis(x, y, degree)
            assert np.allclose(phi, np.eye(nl)[n], atol=1e-14)
