"""Development probe: one PNP Newton step from the nested-iteration state with different AMG settings."""
import os, sys, time, itertools
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from dune_pnp_b200 import capi
levels = int(sys.argv[1]) if len(sys.argv) > 1 else 5
a, cfg = bench.load_case()
c = capi.Context(0); c.mesh_set(**a); c.params_read(cfg)
h, s, us = bench.build_state(c, capi, levels, 3, capi.JAC_ANALYTIC, 2, False)
u = c.vec(3)
opts = c.newton_opts(jac_mode=capi.JAC_ANALYTIC, max_iterations=1)
configs = []
for geo, steps, gamma, wl, alpha in [(1, 2, 1, 99, 1.6), (1, 1, 1, 99, 1.6), (1, 3, 1, 99, 1.6), (0, 2, 2, 5, 1.1)]:
    configs.append(dict(amg_geometric=geo, steps=steps, amg_gamma=gamma, amg_wlevels=wl, amg_alpha=alpha))
for cf in configs:
    sv = c.solver(capi.SOLVER_BCGS, capi.PREC_AMG, 400, cf["steps"], 1)
    for k, v in cf.items():
        if k != "steps":
            c.solver_set_option(sv, k, v)
    c.vec_copy(u, us)
    st, r = c.newton(h, u, sv, opts, check=False)   # warm-up incl. symbolic setup
    c.vec_copy(u, us)
    t0 = time.perf_counter()
    st, r = c.newton(h, u, sv, opts, check=False)
    dt = time.perf_counter() - t0
    print("%-95s status %d its %4d  step %.3fs  solve %.3fs  defect %.2e" % (cf, st, r.linear_iterations, dt, r.seconds_solve, r.defect), flush=True)
