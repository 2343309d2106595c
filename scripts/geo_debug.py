import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, util
from dune_pnp_b200 import capi
for op, F in ((capi.OP_PB,1),(capi.OP_PNP,3)):
  for name, lev in (("pore_small",2),("pore",2)):
    for geo in (0,1):
        a = util.load_mesh_arrays(name)
        c = capi.Context(0); c.mesh_set(**a); c.params_read(util.cfg_path(name)); c.mesh_refine(lev); c.mesh_finalize(True)
        nv = c.mesh_sizes()["nv"]
        h = c.operator(op, 0); u = c.vec(F); c.vec_set(u, 0.05); A = c.matrix(h); c.jacobian(h, u, A, 1, 0.0)
        b = np.random.RandomState(0).uniform(-1,1,F*nv)
        d = c.constraints(h, F); b[d] = 0
        s = c.solver(capi.SOLVER_BCGS, capi.PREC_AMG, 300, 2, 1); c.solver_set_option(s, "amg_geometric", geo)
        z = c.vec(F); r = c.vec(F, b)
        try:
            res = c.solve(s, A, z, r, 1e-8)
            print(name, "op", op, "geo", geo, "its", res.iterations, "conv", res.converged, "red %.2e" % res.reduction, flush=True)
        except Exception as e:
            print(name, "op", op, "geo", geo, "ERR", e, flush=True)
