"""One rank of the multi-GPU parity test (launched by tests/test_gpu_nccl.py through torch.distributed.run, one process
per GPU, NCCL).  Every rank partitions the pore mesh, refines its part, sets up the library's distributed hierarchy and
checks -- against the oracle's GLOBAL objects -- the pieces that only exist with more than one rank:

  1. pnp_halo_exchange: ghost values arrive from their owners;
  2. assemble_residual / assemble_jacobian + SpMV on owned rows (the state's ghost part is poisoned first, so the values
     must come through NCCL), norms and dots summed over ranks;
  3. the Poisson operator's coefficient fields are refreshed at ghost vertices;
  4. BiCGSTAB + distributed multigrid on the PNP Jacobian, and the monolithic PNP Newton run from the oracle's
     interpolate(BCExtension) state: same Newton iteration count and fields as the global oracle run.

Prints "NCCL_WORKER_OK rank r" on success; any assertion kills the launcher with a non-zero exit code.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    import torch
    import torch.distributed as dist
    import util
    from dune_pnp_b200 import capi, partition
    from oracle import binding as ora
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    levels = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def all_gather(obj):
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    a = util.load_mesh_arrays("pore")
    cfg = util.cfg_path("pore")
    plans = partition.build_hierarchy(a, world, rank, levels, all_gather=all_gather)
    uid = [capi.Context.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    root = capi.Context(local)
    children = partition.setup_distributed(capi, root, plans, cfg, rank, world, uid[0], replica_mesh=a, replica_level=0)
    fine = plans[-1]
    n_own, nv = fine.n_own, fine.nv
    assert root.mesh_owned() == n_own and nv > n_own, "every rank must have ghosts"
    # global objects from the oracle (every rank computes them: the mesh is small)
    gm = ora.Mesh.from_arrays(**a).refine(levels)
    p = ora.Params.read(cfg)
    gkeys = {k: i for i, k in enumerate(partition._coord_keys(gm.x, gm.y))}
    l2g = np.array([gkeys[k] for k in partition._coord_keys(fine.x, fine.y)])
    own = l2g[:n_own]
    tot = torch.tensor([float(n_own)], device="cuda"); dist.all_reduce(tot)
    assert int(tot.item()) == gm.nv, "owned vertices must tile the global mesh"

    def loc(g, F):      # global [F][nv_g] -> local [F][nv]
        return g.reshape(F, -1)[:, l2g].reshape(-1)

    def poisoned(g, F):  # owned part right, ghost part garbage: only a halo exchange can fix it
        v = g.reshape(F, -1)[:, l2g].copy(); v[:, n_own:] = 1e30
        return v.reshape(-1)

    def owned(vec, F):
        return root.download(vec, F).reshape(F, -1)[:, :n_own]

    rng = np.random.RandomState(3)
    phi = 0.8 * np.sin(0.07 * gm.x) * np.cos(0.05 * gm.y)
    gu = np.concatenate([phi, 0.06 * np.exp(-phi), 0.06 * np.exp(phi)]) * (1 + 0.01 * rng.uniform(-1, 1, 3 * gm.nv))
    gx = rng.uniform(-1, 1, 3 * gm.nv)
    # 1. halo exchange
    v = root.vec(3, poisoned(gu, 3))
    root.halo_exchange(v)
    assert np.array_equal(root.download(v, 3), loc(gu, 3)), "halo exchange"
    # 2. residual, Jacobian, SpMV on owned rows; reductions over ranks
    h = root.operator(capi.OP_PNP, 0)
    vu, vr, A = root.vec(3, poisoned(gu, 3)), root.vec(3), root.matrix(h)
    root.residual(h, vu, vr)
    r_g, ab = ora.residual(gm, p, ora.OP_PNP, gu, want_abs=True)
    assert np.all(np.abs(owned(vr, 3) - r_g.reshape(3, -1)[:, own]) <= 1e-12 * ab.reshape(3, -1)[:, own] + 1e-300), "residual"
    assert abs(root.norm(vr) - np.linalg.norm(r_g)) <= 1e-12 * np.linalg.norm(r_g), "norm over ranks"
    for mode, tol in ((0, 1e-12), (1, 1e-10)):
        root.upload(vu, poisoned(gu, 3))
        root.jacobian(h, vu, A, mode, 1e-11)
        rp, col, val, jab = ora.jacobian(gm, p, ora.OP_PNP, gu, mode=mode, eps=1e-11, want_abs=True)
        vx, vy = root.vec(3, poisoned(gx, 3)), root.vec(3)
        root.spmv(A, vx, vy)
        y_g = ora.spmv(rp, col, val, gx)
        # entry errors weighted by |x| bound the row error: scale = (|A| + tol-free rounding) |x|
        scale = ora.spmv(rp, col, jab, np.abs(gx)) + ora.spmv(rp, col, np.abs(val), np.abs(gx))
        assert np.all(np.abs(owned(vy, 3) - y_g.reshape(3, -1)[:, own]) <= tol * scale.reshape(3, -1)[:, own] + 1e-300), "spmv %d" % mode
        assert abs(root.dot(vx, vy) - gx @ y_g) <= 1e-9 * (np.abs(gx) @ scale), "dot over ranks"
        root.vec_destroy(vx); root.vec_destroy(vy)
    # 3. Poisson with coefficient fields whose ghost part is stale
    cp, cm = 0.06 * np.exp(-phi), 0.06 * np.exp(phi)
    hp = root.operator(capi.OP_POISSON, 0)
    root.operator_set_coefficient(hp, 0, root.vec(1, poisoned(cp, 1))); root.operator_set_coefficient(hp, 1, root.vec(1, poisoned(cm, 1)))
    v1, r1 = root.vec(1, poisoned(phi, 1)), root.vec(1)
    root.residual(hp, v1, r1)
    rP, abP = ora.residual(gm, p, ora.OP_POISSON, phi, cp, cm, want_abs=True)
    assert np.all(np.abs(owned(r1, 1)[0] - rP[own]) <= 1e-12 * abP[own] + 1e-300), "Poisson residual with ghost coefficients"
    # 4a. linear solve with the distributed multigrid on the exact-derivative Jacobian
    root.upload(vu, loc(gu, 3))
    root.jacobian(h, vu, A, capi.JAC_ANALYTIC, 0.0)
    rp, col, val = ora.jacobian(gm, p, ora.OP_PNP, gu, mode=1)
    gb = rng.uniform(-1, 1, 3 * gm.nv); gb[ora.dirichlet(gm, p, 3, 0)] = 0.0
    s = root.solver(capi.SOLVER_BCGS, capi.PREC_AMG, 300, 2)
    vz, vb = root.vec(3), root.vec(3, loc(gb, 3))
    res = root.solve(s, A, vz, vb, 1e-10)
    assert res.converged and res.iterations <= 16, "distributed multigrid: %d iterations" % res.iterations
    z_g, ro = ora.linsolve(rp, col, val, gb, 1e-12, 20000, ora.SOLVER_BCGS, ora.PREC_ILU0)
    assert ro["converged"]
    zo = owned(vz, 3)
    err = torch.tensor([np.sum((zo - z_g.reshape(3, -1)[:, own]) ** 2)], device="cuda"); dist.all_reduce(err)
    assert np.sqrt(err.item()) <= 1e-7 * np.linalg.norm(z_g), "distributed solve"
    # 4b. monolithic PNP Newton from the oracle's interpolate(BCExtension) state, CONVERGED (reduction 1e-11, linear
    # reduction 1e-9: with pore.cfg's own 1e-9 / 1e-8 the last Newton step is decided by how far the two linear solvers
    # overshoot their tolerance -- the multigrid run stopped after 3 steps where the ILU0 run took 4)
    # exact-derivative Jacobian on both sides: the forward differences' noise (~1e-5 relative in the entries) decides
    # whether a tight run needs one Newton step more or less (measured here: 4 against 5)
    opts = ora.newton_opts(p, solver=ora.SOLVER_BCGS, prec=ora.PREC_ILU0, jac_mode=1); opts[12] = 20000
    opts[0], opts[2] = 1e-11, 1e-9
    pb_g, _ = ora.newton(gm, p, ora.OP_PB, np.zeros(gm.nv), opts)
    u0_g = np.concatenate([ora.interpolate(gm, p, k, pb_g) for k in range(3)])
    u_g, rn = ora.newton(gm, p, ora.OP_PNP, u0_g, opts)
    root.upload(vu, loc(u0_g, 3))
    st, rg = root.newton(h, vu, root.solver(capi.SOLVER_BCGS, capi.PREC_AMG, 20000, 2),
                         root.newton_opts(jac_mode=capi.JAC_ANALYTIC, reduction=1e-11, min_linear_reduction=1e-9))
    assert rg.converged and rn["converged"] and rg.iterations == rn["iterations"], "Newton counts %d vs %d" % (rg.iterations, rn["iterations"])
    assert abs(rg.first_defect - rn["first_defect"]) <= 1e-10 * rn["first_defect"], "first defect over ranks"
    uo = owned(vu, 3)
    for k in range(3):
        e = torch.tensor([np.sum((uo[k] - u_g.reshape(3, -1)[k, own]) ** 2)], device="cuda"); dist.all_reduce(e)
        assert np.sqrt(e.item()) <= 1e-8 * np.linalg.norm(u_g.reshape(3, -1)[k]), "Newton field %d" % k
    print("NCCL_WORKER_OK rank %d: %d owned + %d ghost vertices, Newton %d its, linear %s" % (
        rank, n_own, nv - n_own, rg.iterations, list(rg.linear_iterations_history[:rg.n_history])), flush=True)
    dist.barrier()
    sys.stdout.flush()
    # orderly tear-down: solvers (CUDA graphs with NCCL nodes) and contexts before the communicators
    for ch in children:
        ch.close()
    root.close()
    dist.destroy_process_group()
    os._exit(0)


if __name__ == "__main__":
    try:
        main()
    except BaseException:  # a failing rank must not leave its peers waiting in a collective: die at once, loudly
        import traceback
        traceback.print_exc()
        sys.stdout.flush(); sys.stderr.flush()
        os._exit(1)
