"""Generates tests/golden/*_solution.npz with the CPU oracle: converged fields of the reference flow
(PB Newton -> interpolate(BCExtension) -> monolithic PNP Newton, BiCGSTAB + SSOR(1), FD Jacobian eps=1e-11,
Newton settings tightened to reduction 1e-11 / linear 1e-9 so the fields are converged to ~1e-10).
Used (a) as golden vectors for the GPU parity tests, (b) as the start state of bench.py's CPU arm."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
from oracle import binding as ora
import util
for name in ["one_wall", "cylinder", "pore_small", "pore"]:
    a = util.load_mesh_arrays(name); m = ora.Mesh.from_arrays(**a); p = ora.Params.read(util.cfg_path(name))
    p.sys[5] = 20000; p = ora.Params.from_flat(p.sys, p.surf)
    opts = ora.newton_opts(p, solver=ora.SOLVER_BCGS, prec=ora.PREC_SSOR); opts[0], opts[2] = 1e-11, 1e-9
    t = time.time()
    pb, r0 = ora.newton(m, p, ora.OP_PB, np.zeros(m.nv), opts)
    u0 = np.concatenate([ora.interpolate(m, p, k, pb) for k in range(3)])
    u, r = ora.newton(m, p, ora.OP_PNP, u0, opts)
    opts1 = opts.copy(); opts1[4] = 1
    u1, _ = ora.newton(m, p, ora.OP_PNP, u0, opts1)  # state after the first Newton iteration
    assert r0["converged"] and r["converged"], name
    np.savez_compressed(os.path.join(util.GOLDEN, name + "_solution.npz"), pb=pb, u0=u0, u1=u1, u=u,
                        pb_newton_iterations=r0["iterations"], pnp_newton_iterations=r["iterations"],
                        pnp_defects=r["defect_history"])
    print(name, r0["iterations"], r["iterations"], r["lin_iter_history"], "%.1fs" % (time.time() - t))
