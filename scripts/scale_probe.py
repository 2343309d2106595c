"""Development probe: builds the refined pore mesh on the device and times the individual hot kernels.
Not the bench (bench.py is); prints one line per kernel with achieved GB/s on the algorithmic bytes."""
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402  (CUDA events + sync only)

import util  # noqa: E402
from dune_pnp_b200 import capi  # noqa: E402


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    return float(np.median(ts))


def main():
    levels = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    renumber = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    a = util.load_mesh_arrays("pore")
    c = capi.Context(0)
    c.mesh_set(**a); c.params_read(util.cfg_path("pore"))
    t0 = time.perf_counter(); c.mesh_refine(levels); t1 = time.perf_counter(); c.mesh_finalize(bool(renumber)); t2 = time.perf_counter()
    s = c.mesh_sizes()
    nv, nT, ns = s["nv"], s["nT"], s["nslots"]
    print("levels %d renumber %d: nv %d nT %d nslots %d  refine %.2fs finalize %.2fs" % (levels, renumber, nv, nT, ns, t1 - t0, t2 - t1))
    free, total = torch.cuda.mem_get_info()
    print("mem used %.1f GB" % ((total - free) / 1e9))
    for op, F, NP in ((capi.OP_PB, 1, 1), (capi.OP_PNP, 3, 7)):
        h = c.operator(op, 0)
        u, r, x, y, A = c.vec(F), c.vec(F), c.vec(F), c.vec(F), c.matrix(h)
        c.vec_set(u, 0.05); c.vec_set(x, 1.0)
        t = timed(lambda: c.residual(h, u, r))
        b = 4 * ns + 4 * nv + 16 * nv + 16 * F * nv
        print("op %d residual      %8.3f ms  %7.1f GB/s (%.2f GB)" % (op, t * 1e3, b / t / 1e9, b / 1e9))
        for mode in (1, 0):
            t = timed(lambda: c.jacobian(h, u, A, mode, 1e-11), reps=3, warm=1)
            b = 4 * ns + 4 * nv + 16 * nv + 8 * F * nv + 8 * NP * ns
            print("op %d jacobian m%d   %8.3f ms  %7.1f GB/s (%.2f GB)" % (op, mode, t * 1e3, b / t / 1e9, b / 1e9))
        c.jacobian(h, u, A, 1, 0.0)
        t = timed(lambda: c.spmv(A, x, y), reps=10)
        b = (8 * NP + 4) * ns + 4 * nv + 16 * F * nv
        print("op %d spmv          %8.3f ms  %7.1f GB/s (%.2f GB)" % (op, t * 1e3, b / t / 1e9, b / 1e9))
        t = timed(lambda: c.dot(x, y), reps=10)
        print("op %d dot           %8.3f ms  %7.1f GB/s" % (op, t * 1e3, 16 * F * nv / t / 1e9))
        t = timed(lambda: c.axpy(y, 0.5, x), reps=10)
        print("op %d axpy          %8.3f ms  %7.1f GB/s" % (op, t * 1e3, 24 * F * nv / t / 1e9))
        for v in (u, r, x, y):
            c.vec_destroy(v)
        c.matrix_destroy(A)


if __name__ == "__main__":
    main()
