"""Timings of the quadratic / cubic-element row on a B200: residual, both Jacobians, CSR SpMV and one
SSOR(1) application of the 3-field PNP operator on the pore mesh refined `--levels` times.  Device times (CUDA events through
pnp_timer_start/stop), medians of `--reps` runs after a warm-up; algorithmic bytes: CSR SpMV 12 B per non-zero + 20 B per row.
    python scripts/bench_pk.py --levels 4 > profiles/pk_bench_r02.json
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import util  # noqa: E402
from dune_pnp_b200 import capi  # noqa: E402


def timed(c, fn, reps):
    fn()
    ts = []
    for _ in range(reps):
        c.timer_start(); fn(); ts.append(c.timer_stop())
    return float(np.median(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--levels", type=int, default=4)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    out = {"mesh": "pore.msh refined %d times" % a.levels, "operator": "PnpOperator (3 fields)", "reps": a.reps, "degrees": {}}
    for degree in (2, 3):
        c = capi.Context(0)
        c.mesh_set(**util.load_mesh_arrays("pore"))
        c.params_read(util.cfg_path("pore"))
        c.mesh_refine(a.levels)
        c.space_set_degree(degree)
        import time
        t0 = time.time(); c.mesh_finalize(True); t_fin = time.time() - t0
        nd = c.ndof()
        h = c.operator(capi.OP_PNP, 0)
        if degree == 3:
            c.operator_set_intorder(h, 5)
        rng = np.random.RandomState(0)
        u0 = np.concatenate([0.1 * rng.uniform(-1, 1, nd), 0.06 * (1 + 0.1 * rng.uniform(-1, 1, 2 * nd))])
        u, r, A = c.vec(3, u0), c.vec(3), c.matrix(h)
        t0 = time.time(); rp, col = c.pattern(h, 3); t_pat = time.time() - t0
        nnz = len(col)
        res = {"dofs": 3 * nd, "nnz": nnz, "finalize_s_host": t_fin, "pattern_s_host": t_pat}
        res["residual_ms"] = timed(c, lambda: c.residual(h, u, r), a.reps)
        res["jacobian_exact_ms"] = timed(c, lambda: c.jacobian(h, u, A, 1, 1e-11), a.reps)
        res["jacobian_fd_ms"] = timed(c, lambda: c.jacobian(h, u, A, 0, 1e-11), a.reps)
        x, y = c.vec(3, rng.uniform(-1, 1, 3 * nd)), c.vec(3)
        ms = timed(c, lambda: c.spmv(A, x, y), a.reps)
        res["spmv_ms"] = ms
        res["spmv_gbs"] = (12.0 * nnz + 20.0 * 3 * nd) / ms / 1e6
        s = c.solver(capi.SOLVER_BCGS, capi.PREC_SSOR, 10, 1)
        c.precond_apply(s, A, x, y)
        res["ssor1_apply_ms"] = timed(c, lambda: c.precond_apply(s, A, x, y), a.reps)
        res["ssor_levels"] = c.solver_get(s, "ssor_levels")
        res["assembled_dofs_per_s_exact"] = 3 * nd / (res["jacobian_exact_ms"] + res["residual_ms"]) * 1e3
        out["degrees"][str(degree)] = res
        c.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
