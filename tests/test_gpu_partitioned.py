"""GPU tests of the partitioned (multi-GPU) code path on ONE device: the ranks are emulated as separate contexts
whose local meshes, ownership and halo plans come from dune_pnp_b200/partition.py; the halo exchange itself is done
by the test through host buffers (NCCL is exercised by `bench.py --gpus N`).  Owned rows of every emulated rank must
reproduce the oracle's rows of the global problem."""
import threading

import numpy as np
import pytest

import util
from oracle import binding as ora

pytestmark = pytest.mark.gpu


def _field(x, y, k=0):
    return np.sin(0.3 * x + k) + np.cos(0.17 * y - k) * 0.5 + 0.01 * x * y


def _plans(a, world, levels):
    from dune_pnp_b200 import partition
    slots = [None] * world
    barrier = threading.Barrier(world)
    lock = threading.Lock()
    store = {}

    def run(rank):
        calls = [0]

        def all_gather(obj):
            key = calls[0]; calls[0] += 1
            with lock:
                store.setdefault(key, [None] * world)[rank] = obj
            barrier.wait()
            out = list(store[key])
            barrier.wait()
            return out
        slots[rank] = partition.build_local(a, world, rank, levels, all_gather=all_gather)
    ts = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    [t.start() for t in ts]; [t.join() for t in ts]
    return slots


@pytest.mark.parametrize("world,levels", [(2, 1), (4, 2)])
def test_partitioned_rows_match_global_oracle(world, levels):
    from dune_pnp_b200 import capi, partition
    a = util.load_mesh_arrays("pore_small")
    plans = _plans(a, world, levels)
    gm = ora.Mesh.from_arrays(**a).refine(levels)
    p = ora.Params.read(util.cfg_path("pore_small"))
    gkeys = {k: i for i, k in enumerate(partition._coord_keys(gm.x, gm.y))}
    gu = np.concatenate([_field(gm.x, gm.y, k) for k in range(3)])
    gx = np.concatenate([_field(gm.x, gm.y, k + 5) for k in range(3)])
    r_glob, ab = ora.residual(gm, p, ora.OP_PNP, gu, want_abs=True)
    rp, col, val = ora.jacobian(gm, p, ora.OP_PNP, gu, mode=1)
    y_glob = ora.spmv(rp, col, val, gx)
    y_scale = ora.spmv(rp, col, np.abs(val), np.abs(gx))
    owned_total = 0
    for rank, plan in enumerate(plans):
        l2g = np.array([gkeys[k] for k in partition._coord_keys(plan.x, plan.y)])
        c = capi.Context(0)
        c.params_read(util.cfg_path("pore_small"))
        c.mesh_set_local(plan.n_own, plan.x, plan.y, plan.tri, plan.ba, plan.bb, plan.bphys)
        c.halo_set(plan.nbr, plan.send_ptr, plan.send_idx, plan.recv_ptr)
        c.mesh_finalize(True)
        assert c.mesh_owned() == plan.n_own
        own = l2g[:plan.n_own]
        h = c.operator(capi.OP_PNP, 0)
        # ghost values supplied by the test (what the halo exchange would deliver)
        u = c.vec(3, gu.reshape(3, -1)[:, l2g].reshape(-1)); r = c.vec(3)
        c.residual(h, u, r)
        r_loc = c.download(r, 3).reshape(3, -1)[:, :plan.n_own]
        assert np.all(np.abs(r_loc - r_glob.reshape(3, -1)[:, own]) <= 1e-12 * ab.reshape(3, -1)[:, own] + 1e-300)
        A = c.matrix(h)
        c.jacobian(h, u, A, capi.JAC_ANALYTIC, 0.0)
        x = c.vec(3, gx.reshape(3, -1)[:, l2g].reshape(-1)); y = c.vec(3)
        c.spmv(A, x, y)
        y_loc = c.download(y, 3).reshape(3, -1)[:, :plan.n_own]
        assert np.all(np.abs(y_loc - y_glob.reshape(3, -1)[:, own]) <= 1e-12 * y_scale.reshape(3, -1)[:, own] + 1e-300)
        # norms / dots run over owned dofs only
        assert abs(c.norm(x) - np.linalg.norm(gx.reshape(3, -1)[:, own])) <= 1e-12 * np.linalg.norm(gx)
        owned_total += plan.n_own
        c.close()
    assert owned_total == gm.nv


def test_partitioned_local_amg_solves_owned_block():
    """Block-Jacobi AMG: on one emulated rank, BiCGSTAB + AMG solves the owned-row system with ghost columns frozen."""
    from dune_pnp_b200 import capi
    a = util.load_mesh_arrays("pore_small")
    plan = _plans(a, 2, 2)[1]
    c = capi.Context(0)
    c.params_read(util.cfg_path("pore_small"))
    c.mesh_set_local(plan.n_own, plan.x, plan.y, plan.tri, plan.ba, plan.bb, plan.bphys)
    c.halo_set(plan.nbr, plan.send_ptr, plan.send_idx, plan.recv_ptr)
    c.mesh_finalize(True)
    h = c.operator(capi.OP_PB, 0)
    u = c.vec(1); A = c.matrix(h)
    c.jacobian(h, u, A, capi.JAC_ANALYTIC, 0.0)
    b = np.zeros(plan.nv); b[:plan.n_own] = np.random.RandomState(0).uniform(-1, 1, plan.n_own)
    s = c.solver(capi.SOLVER_BCGS, capi.PREC_AMG, 500, 2)
    z, rhs = c.vec(1), c.vec(1, b)
    res = c.solve(s, A, z, rhs, 1e-10)
    assert res.converged and res.iterations < 60
    assert np.all(c.download(z, 1)[plan.n_own:] == 0.0)  # ghost part untouched


@pytest.mark.parametrize("aggregated", [False, True])
@pytest.mark.parametrize("op_name,F", [("OP_PB", 1), ("OP_PNP", 3)])
def test_distributed_multigrid_on_one_rank(op_name, F, aggregated):
    """The distributed geometric multigrid (child contexts per level, re-discretised coarse operators, dense coarsest
    solve by global index) with a single subdomain: everything but NCCL runs."""
    from dune_pnp_b200 import capi, partition
    a = util.load_mesh_arrays("pore")
    plans = partition.build_hierarchy(a, 1, 0, 2)
    root = capi.Context(0)
    aggs = partition.aggregate_greedy(len(a["x"]), a["tri"]) if aggregated else None
    children = partition.setup_distributed(capi, root, plans, util.cfg_path("pore"), 0, 1, None, aggregates=aggs)
    assert len(children) == 2
    op = getattr(capi, op_name)
    h = root.operator(op, 0)
    nv = root.mesh_sizes()["nv"]
    u = root.vec(F); root.vec_set(u, 0.05)
    A = root.matrix(h)
    root.jacobian(h, u, A, capi.JAC_ANALYTIC, 0.0)
    b = np.random.RandomState(0).uniform(-1, 1, F * nv)
    b[root.constraints(h, F)] = 0.0
    s = root.solver(capi.SOLVER_BCGS, capi.PREC_AMG, 200, 2)
    z, r = root.vec(F), root.vec(F, b)
    res = root.solve(s, A, z, r, 1e-8)
    assert res.converged and res.iterations <= (16 if aggregated else 8)


@pytest.mark.parametrize("op_name,F", [("OP_PB", 1), ("OP_PNP", 3)])
def test_distributed_multigrid_replica_coarse_level(op_name, F):
    """Replica variant of the coarsest level: the distributed Gmsh level is gathered to a context holding the whole Gmsh
    mesh, below which the one-GPU hierarchy (aggregation levels + small dense LU) runs; "replica1": the hierarchy stops
    being distributed at refinement level 1 and the replica refines the Gmsh mesh itself.  One subdomain (no NCCL):
    converges like the exact dense coarse solve, and the solution solves the assembled system."""
    from dune_pnp_b200 import capi, partition
    a = util.load_mesh_arrays("pore")
    plans = partition.build_hierarchy(a, 1, 0, 2)
    its = {}
    for variant in ("dense", "replica", "replica1"):
        root = capi.Context(0)
        children = partition.setup_distributed(capi, root, plans, util.cfg_path("pore"), 0, 1, None,
                                               replica_mesh=None if variant == "dense" else a,
                                               replica_level=1 if variant == "replica1" else 0)
        assert len(children) == {"dense": 2, "replica": 3, "replica1": 2}[variant]
        op = getattr(capi, op_name)
        h = root.operator(op, 0)
        nv = root.mesh_sizes()["nv"]
        u = root.vec(F); root.vec_set(u, 0.05)
        A = root.matrix(h)
        root.jacobian(h, u, A, capi.JAC_ANALYTIC, 0.0)
        b = np.random.RandomState(0).uniform(-1, 1, F * nv)
        b[root.constraints(h, F)] = 0.0
        s = root.solver(capi.SOLVER_BCGS, capi.PREC_AMG, 200, 2)
        z, r, y = root.vec(F), root.vec(F, b), root.vec(F)
        res = root.solve(s, A, z, r, 1e-8)
        assert res.converged
        its[variant] = res.iterations
        root.spmv(A, z, y)
        assert np.linalg.norm(root.download(y, F) - b) <= 2e-8 * np.linalg.norm(b)
        del children
        root.close()
    assert its["replica"] <= its["dense"] + 4 and its["replica1"] <= its["dense"] + 4
