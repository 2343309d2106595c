"""Entry-wise parity figures for DESIGN.md section 6: besides the test tolerance (|gpu - oracle| <= 1e-12 * sum of the
|element contributions| of the entry), the plain relative error |gpu - oracle| / |oracle| of every residual and Jacobian
entry that is not a cancellation (|oracle| >= 1e-6 of its contribution sum), PNP and PB, pore mesh refined twice."""
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import util  # noqa: E402
from dune_pnp_b200 import capi  # noqa: E402
from oracle import binding as ora  # noqa: E402

a = util.load_mesh_arrays("pore")
c = capi.Context(0); c.mesh_set(**a); c.params_read(util.cfg_path("pore")); c.mesh_refine(2); c.mesh_finalize(True)
m = ora.Mesh.from_arrays(**a).refine(2); p = ora.Params.read(util.cfg_path("pore"))
rng = np.random.RandomState(21)
phi = 0.8 * np.sin(0.07 * m.x) * np.cos(0.05 * m.y) + 0.05 * rng.uniform(-1, 1, m.nv)
for op, F, u in ((capi.OP_PB, 1, phi), (capi.OP_PNP, 3, np.concatenate([phi, 0.06 * np.exp(-phi), 0.06 * np.exp(phi)]))):
    h = c.operator(op, 0)
    vu, vr, A = c.vec(F, u), c.vec(F), c.matrix(h)
    c.residual(h, vu, vr)
    r = c.download(vr, F)
    r_o, ab = ora.residual(m, p, op, u, want_abs=True)
    keep = np.abs(r_o) >= 1e-6 * ab
    print("op %d residual: max |d|/sum|contrib| %.2e   entry-wise max |d|/|oracle| %.2e over %d of %d entries" % (
        op, np.max(np.abs(r - r_o) / np.maximum(ab, 1e-300)), np.max(np.abs(r - r_o)[keep] / np.abs(r_o[keep])), keep.sum(), len(r)))
    for mode in (0, 1):
        c.jacobian(h, vu, A, mode, 1e-11)
        rp, col, val_o, jab = ora.jacobian(m, p, op, u, mode=mode, eps=1e-11, want_abs=True)
        val = c.matrix_values(h, A, len(col))
        keep = np.abs(val_o) >= 1e-6 * jab
        print("op %d jacobian mode %d: max |d|/sum|contrib| %.2e   entry-wise max |d|/|oracle| %.2e over %d of %d entries, bit-equal %.1f %%" % (
            op, mode, np.max(np.abs(val - val_o) / np.maximum(jab, 1e-300)), np.max(np.abs(val - val_o)[keep] / np.abs(val_o[keep])),
            keep.sum(), len(val), 100 * np.mean(val == val_o)))
