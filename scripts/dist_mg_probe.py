"""Development probe: distributed multigrid on ONE rank (no NCCL): iteration counts for coarse-level variants."""
import sys, os, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, util
from dune_pnp_b200 import capi, partition
levels = int(sys.argv[1]) if len(sys.argv) > 1 else 3
a = util.load_mesh_arrays("pore")
plans = partition.build_hierarchy(a, 1, 0, levels)
def aggregate_small(nv, tri, maxsize):
    agg, n = partition.aggregate_greedy(nv, tri)
    return agg, n
for variant in ("exact", "greedy", "greedy2"):
  for alpha in ((1.6,) if variant == "exact" else (1.0, 1.3, 1.6)):
    root = capi.Context(0)
    if variant == "exact": aggs = None
    elif variant == "greedy": aggs = partition.aggregate_greedy(len(a["x"]), a["tri"])
    else: aggs = partition.aggregate_greedy(len(a["x"]), a["tri"], leftovers_join=False)
    ch = partition.setup_distributed(capi, root, plans, util.cfg_path("pore"), 0, 1, None, aggregates=aggs)
    for op, F in ((capi.OP_PB, 1), (capi.OP_PNP, 3)):
        h = root.operator(op, 0); nv = root.mesh_sizes()["nv"]
        u = root.vec(F); root.vec_set(u, 0.05); A = root.matrix(h); root.jacobian(h, u, A, 1, 0.0)
        b = np.random.RandomState(0).uniform(-1, 1, F * nv); b[root.constraints(h, F)] = 0
        s = root.solver(capi.SOLVER_BCGS, capi.PREC_AMG, 200, 2); root.solver_set_option(s, "amg_alpha", alpha)
        z, r = root.vec(F), root.vec(F, b)
        res = root.solve(s, A, z, r, 1e-8)
        print(variant, "n_agg", None if aggs is None else aggs[1], "alpha", alpha, "op", op, "its", res.iterations, res.converged, flush=True)
    del ch; root.close()
