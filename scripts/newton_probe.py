"""Development probe: PB Newton -> interpolate(BCExtension) -> PNP Newton on the refined pore mesh with a chosen
preconditioner; prints iteration counts and timings."""
import os, sys, time
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import util
from dune_pnp_b200 import capi

levels = int(sys.argv[1]) if len(sys.argv) > 1 else 3
prec = int(sys.argv[2]) if len(sys.argv) > 2 else capi.PREC_AMG
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
mode = int(sys.argv[4]) if len(sys.argv) > 4 else 1
maxnewton = int(sys.argv[5]) if len(sys.argv) > 5 else 50
a = util.load_mesh_arrays("pore")
c = capi.Context(0)
c.mesh_set(**a); c.params_read(util.cfg_path("pore"))
c.mesh_refine(levels); c.mesh_finalize(True)
s = c.mesh_sizes()
print("levels", levels, s, "prec", prec, "steps", steps, "jac mode", mode, flush=True)
hpb = c.operator(capi.OP_PB, 0)
spb = c.solver(capi.SOLVER_BCGS, prec, 20000, steps, 1)
vpb = c.vec(1)
t0 = time.time()
st, r = c.newton(hpb, vpb, spb, c.newton_opts(jac_mode=mode), check=False)
print("PB: status %d its %d lin %s defects %s  %.3fs (asm %.3f solve %.3f)" % (st, r.iterations, list(r.linear_iterations_history[:r.n_history]),
      ["%.2e" % d for d in r.defect_history[:r.n_history]], time.time() - t0, r.seconds_assembly, r.seconds_solve), flush=True)
f = [c.vec(1) for _ in range(3)]
t0 = time.time()
for k in range(3):
    c.interpolate_bcext(k, vpb, f[k])
print("interpolate %.3fs" % (time.time() - t0))
vu = c.vec(3); c.pack3(vu, *f)
h = c.operator(capi.OP_PNP, 0)
sp = c.solver(capi.SOLVER_BCGS, prec, 20000, steps, 1)
t0 = time.time()
st, r = c.newton(h, vu, sp, c.newton_opts(jac_mode=mode, max_iterations=maxnewton), check=False)
print("PNP: status %d its %d lin %s defects %s ls %d  %.3fs (asm %.3f solve %.3f)" % (st, r.iterations, list(r.linear_iterations_history[:r.n_history]),
      ["%.2e" % d for d in r.defect_history[:r.n_history]], r.line_search_trials, time.time() - t0, r.seconds_assembly, r.seconds_solve), flush=True)
