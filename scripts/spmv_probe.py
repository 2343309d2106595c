"""Times the fine-level star SpMV kernels alone (CUDA events around every launch, pnp_profile_spmv) for each kernel
variant selected through pnp_tune: plain y = A x, residual b - A x, smoother step.  k = 7 by default (141 M dofs).
  python scripts/spmv_probe.py [levels] [variant ...]     variant = name=value[,name=value...]   e.g. tma=0 tma=1,tma_stages=2
"""
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import json  # noqa: E402

import util  # noqa: E402
from dune_pnp_b200 import capi  # noqa: E402

levels = int(sys.argv[1]) if len(sys.argv) > 1 else 7
variants = sys.argv[2:] or ["tma=0", "tma=1,tma_stages=3", "tma=1,tma_stages=2"]
c = capi.Context(0)
c.mesh_set(**util.load_mesh_arrays("pore")); c.params_read(util.cfg_path("pore"))
c.mesh_refine(levels); c.mesh_finalize(True)
sz = c.mesh_sizes()
nv, ns = sz["nv"], sz["nslots"]
h = c.operator(capi.OP_PNP, 0)
u, x, y, A = c.vec(3), c.vec(3), c.vec(3), c.matrix(h)
c.vec_set(u, 0.05); c.vec_set(x, 1.0)
c.jacobian(h, u, A, 1, 0.0)
base = 60 * ns + 52 * nv
nbytes = {"plain": base, "residual": base + 24 * nv, "smoother": base + 48 * nv}  # smoother: b and the rows' own x (D^-1 is computed in the kernel)
print("levels %d nv %d nslots %d" % (levels, nv, ns), flush=True)
for var in variants:
    for kv in var.split(","):
        k, v = kv.split("=")
        capi.tune(k, float(v))
    s = c.solver(capi.SOLVER_BCGS, capi.PREC_AMG, 100, 2)
    c.precond_apply(s, A, x, y)  # setup + warm-up
    for _ in range(3):
        c.spmv(A, x, y)
    c.profile_spmv(True)
    for _ in range(10):
        c.spmv(A, x, y)
    for _ in range(5):
        c.precond_apply(s, A, x, y)
    n, ms = c.profile_spmv_get()
    c.profile_spmv(False)
    out = {}
    for name, cnt, t in zip(("plain", "residual", "smoother"), n, ms):
        if cnt:
            out[name] = {"launches": cnt, "avg_ms": round(t / cnt, 4), "GB/s": round(nbytes[name] / (t / cnt) / 1e6, 1)}
    print(var, json.dumps(out), flush=True)
