"""Profiling target: the hot kernels once each on the refined pore mesh (default k = 7: 141 M dofs), for
`ncu --set full -k regex:'k_star_op|k_residual|k_jacobian'`: PNP residual, analytic Jacobian, FD-faithful Jacobian,
two fine-level SpMVs (k_star_op<7,0,0>) and, with a second argument "amg", one multigrid V(2,2) application whose first
k_star_op launches are the fine-level smoother step (k_star_op<7,2,0>) and residual (k_star_op<7,1,0>).
Prints CUDA-event times of the same launches when run without ncu."""
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import util  # noqa: E402
from dune_pnp_b200 import capi  # noqa: E402

levels = int(sys.argv[1]) if len(sys.argv) > 1 else 7
c = capi.Context(0)
c.mesh_set(**util.load_mesh_arrays("pore")); c.params_read(util.cfg_path("pore"))
c.mesh_refine(levels); c.mesh_finalize(True)
s = c.mesh_sizes()
nv, ns = s["nv"], s["nslots"]
h = c.operator(capi.OP_PNP, 0)
u, r, x, y, A = c.vec(3), c.vec(3), c.vec(3), c.vec(3), c.matrix(h)
c.vec_set(u, 0.05); c.vec_set(x, 1.0)


def timed(name, fn, nbytes):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("%-22s %8.3f ms  %7.1f GB/s on %.2f GB algorithmic" % (name, ms, nbytes / ms / 1e6, nbytes / 1e9), flush=True)


res_b = 4 * ns + (4 + 16 + 48) * nv
jac_b = res_b - 24 * nv + 56 * ns
spmv_b = 60 * ns + (4 + 48) * nv
print("levels %d nv %d nslots %d" % (levels, nv, ns))
timed("residual PNP", lambda: c.residual(h, u, r), res_b)
timed("jacobian analytic", lambda: c.jacobian(h, u, A, 1, 0.0), jac_b)
timed("jacobian fd-faithful", lambda: c.jacobian(h, u, A, 0, 1e-11), jac_b)
c.jacobian(h, u, A, 1, 0.0)
timed("spmv 3-field", lambda: c.spmv(A, x, y), spmv_b)
timed("spmv 3-field", lambda: c.spmv(A, x, y), spmv_b)
if len(sys.argv) > 2 and sys.argv[2] == "amg":
    s_amg = c.solver(capi.SOLVER_BCGS, capi.PREC_AMG, 100, 2)
    c.precond_apply(s_amg, A, x, y)   # setup (symbolic + numeric) and one cycle
    timed("multigrid setup + V(2,2)", lambda: c.precond_apply(s_amg, A, x, y), 4 * 27.9e9)
