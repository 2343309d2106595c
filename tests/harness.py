"""ctypes wrapper of tests/host_harness/harness.cpp (TEST INFRASTRUCTURE).

Runs the product's __host__ __device__ per-item functions (element integrals, star walking, star
construction, refinement) on the CPU so that their logic is checked against the oracle without a
GPU.  The product library has no CPU path; this harness is never shipped or loaded by it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "host_harness", "harness.cpp")
_SO = os.path.join(_HERE, "host_harness", "libhostharness.so")
_LIB = None
VMASK = 0x07FFFFFF
PNP_PLANE = {(0, 0): 0, (0, 1): 1, (0, 2): 2, (1, 0): 3, (1, 1): 4, (2, 0): 5, (2, 2): 6}

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def lib():
    global _LIB
    if _LIB is None:
        csrc = os.path.join(_HERE, "..", "dune_pnp_b200", "csrc")
        deps = [_SRC] + [os.path.join(csrc, f) for f in ("pnp_elem.cuh", "pnp_star.cuh", "pnp_setup_algos.cuh", "pnp_sweep.cuh", "pnp_elem_p2.cuh")]
        if not os.path.exists(_SO) or any(os.path.getmtime(d) > os.path.getmtime(_SO) for d in deps):
            subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-x", "c++",
                                   _SRC, "-o", _SO])
        L = C.CDLL(_SO)
        L.hh_refine.restype = C.c_long
        L.hh_star_build.restype = C.c_void_p
        L.hh_star_nslots.restype = C.c_long
        _LIB = L
    return _LIB


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _i(a):
    return None if a is None else a.ctypes.data_as(_ip)


def refine(a):
    x = np.ascontiguousarray(a["x"], dtype=np.float64); y = np.ascontiguousarray(a["y"], dtype=np.float64)
    tri = np.ascontiguousarray(a["tri"], dtype=np.int32)
    ba = np.ascontiguousarray(a["ba"], dtype=np.int32); bb = np.ascontiguousarray(a["bb"], dtype=np.int32)
    ph = np.ascontiguousarray(a["bphys"], dtype=np.int32)
    nv, nT, nB = len(x), len(tri), len(ba)
    args = (C.c_long(nv), _d(x), _d(y), C.c_long(nT), _i(tri), C.c_long(nB), _i(ba), _i(bb), _i(ph))
    nE = lib().hh_refine(*args, None, None, None, None, None, None)
    ox = np.zeros(nv + nE); oy = np.zeros(nv + nE); ot = np.zeros((4 * nT, 3), dtype=np.int32)
    oa = np.zeros(2 * nB, dtype=np.int32); ob = np.zeros(2 * nB, dtype=np.int32); op = np.zeros(2 * nB, dtype=np.int32)
    lib().hh_refine(*args, _d(ox), _d(oy), _i(ot), _i(oa), _i(ob), _i(op))
    return dict(x=ox, y=oy, tri=ot, ba=oa, bb=ob, bphys=op)


class Star:
    def __init__(self, a, surf, renumber=True, n_own=None):
        """a: mesh arrays; surf: [ns][9] flat surface table (btype, flux, dirichlet) x 3 components."""
        self.a = {k: np.ascontiguousarray(v) for k, v in a.items()}
        x = self.a["x"].astype(np.float64); y = self.a["y"].astype(np.float64)
        tri = self.a["tri"].astype(np.int32); ba = self.a["ba"].astype(np.int32); bb = self.a["bb"].astype(np.int32)
        ph = self.a["bphys"].astype(np.int32)
        surf = np.asarray(surf, dtype=np.float64).reshape(-1, 9)
        bt = np.ascontiguousarray(surf[:, 0::3], dtype=np.int32)
        fl = np.ascontiguousarray(surf[:, 1::3], dtype=np.float64)
        err = C.c_int(0)
        self.nv, self.nT = len(x), len(tri)
        self.n_own = self.nv if n_own is None else int(n_own)
        h = lib().hh_star_build(C.c_long(self.nv), _d(x), _d(y), C.c_long(self.nT), _i(tri), C.c_long(len(ba)), _i(ba), _i(bb),
                                _i(ph), int(renumber), len(surf), _i(bt), _d(fl), C.c_long(self.n_own), C.byref(err))
        self.err = err.value
        if not h:
            raise RuntimeError("star build failed with mesh error %d" % err.value)
        self.h = C.c_void_p(h)
        self.nslots = lib().hh_star_nslots(self.h)
        self.rp = np.zeros(self.n_own + 1, dtype=np.int32); self.adj = np.zeros(self.nslots, dtype=np.uint32)
        self.int2ext = np.zeros(self.nv, dtype=np.int32); self.dmask = np.zeros(self.nv, dtype=np.uint8)
        lib().hh_star_get(self.h, _i(self.rp), self.adj.ctypes.data_as(C.POINTER(C.c_uint)), _i(self.int2ext),
                          self.dmask.ctypes.data_as(C.POINTER(C.c_ubyte)))
        self.ext2int = np.empty_like(self.int2ext); self.ext2int[self.int2ext] = np.arange(self.nv, dtype=np.int32)

    def __del__(self):
        try:
            lib().hh_star_free(self.h)
        except Exception:
            pass

    # ---- layout conversions (mirror vec_upload / vec_download / walk_pattern of pnp_setup.cu) ----
    def to_internal(self, lex, F):
        lex = np.asarray(lex, dtype=np.float64).reshape(F, self.nv)
        return np.ascontiguousarray(lex[:, self.int2ext].T).reshape(-1)

    def to_external(self, blk, F):
        out = np.zeros((F, self.nv))
        out[:, self.int2ext] = np.asarray(blk).reshape(self.nv, F).T
        return out.reshape(-1)

    def dirichlet(self, F, comp0=0):
        out = np.zeros((F, self.nv), dtype=bool)
        for k in range(F):
            out[k, self.int2ext] = (self.dmask >> (k if F == 3 else comp0)) & 1
        return out.reshape(-1)

    def export_csr(self, F, comp0=0, vals=None):
        """PDELab-1.1 pattern in reference numbering; vals = planes*nslots internal values (optional)."""
        nv = self.nv
        dirm = self.dirichlet(F, comp0).reshape(F, nv)
        rowptr = [0]; col = []; val = []
        cols_of = []
        for vi in range(nv):
            s0, s1 = self.rp[vi], self.rp[vi + 1]
            ext = self.int2ext[self.adj[s0:s1] & VMASK]
            order = np.argsort(ext, kind="stable")
            cols_of.append((ext[order], np.arange(s0, s1)[order]))
        for ki in range(F):
            for ve in range(nv):
                vi = self.ext2int[ve]
                if dirm[ki, ve]:
                    col.append(ki * nv + ve)
                    if vals is not None:
                        val.append(vals[(PNP_PLANE[(ki, ki)] if F == 3 else 0) * self.nslots + self.rp[vi]])
                else:
                    ext, slots = cols_of[vi]
                    for kj in range(F):
                        keep = ~dirm[kj, ext]
                        col.extend((kj * nv + ext[keep]).tolist())
                        if vals is not None:
                            pl = PNP_PLANE.get((ki, kj), -1) if F == 3 else 0
                            val.extend((vals[pl * self.nslots + slots[keep]] if pl >= 0 else np.zeros(keep.sum())).tolist())
                rowptr.append(len(col))
        out = (np.array(rowptr, dtype=np.int32), np.array(col, dtype=np.int32))
        return out + (np.array(val),) if vals is not None else out

    def _phys(self, params_sys, valency):
        return np.array([params_sys[4], params_sys[2], params_sys[3], valency, params_sys[1]], dtype=np.float64)

    def residual(self, op, params_sys, u_lex, aux0=None, aux1=None, valency=1.0, comp0=0):
        F = 3 if op == 4 else 1
        u = self.to_internal(u_lex, F)
        a0 = None if aux0 is None else self.to_internal(aux0, 1)
        a1 = None if aux1 is None else self.to_internal(aux1, 1)
        r = np.zeros_like(u)
        lib().hh_residual(self.h, op, _d(self._phys(params_sys, valency)), _d(u), _d(a0), _d(a1), comp0, _d(r))
        return self.to_external(r, F)

    def jacobian(self, op, params_sys, u_lex, aux0=None, aux1=None, valency=1.0, comp0=0, mode=0, eps=1e-11):
        F = 3 if op == 4 else 1
        u = self.to_internal(u_lex, F)
        a0 = None if aux0 is None else self.to_internal(aux0, 1)
        a1 = None if aux1 is None else self.to_internal(aux1, 1)
        vals = np.zeros((7 if F == 3 else 1) * self.nslots)
        lib().hh_jacobian(self.h, op, _d(self._phys(params_sys, valency)), _d(u), _d(a0), _d(a1), comp0, mode,
                          C.c_double(eps), _d(vals))
        self.last_vals = vals  # internal planes, for the sweep emulation
        return self.export_csr(F, comp0, vals)

    # ---- level-scheduled preconditioner sweeps (pnp_sweep.cuh) ----
    def sweep_levels(self, F, full):
        lev = np.zeros(F * self.n_own, dtype=np.int32)
        n = lib().hh_sweep_levels(self.h, F, int(full), _i(lev))
        return n, lev

    def ssor_apply(self, F, vals, steps, d_lex):
        """vals: NP planes in internal layout; d, result in reference numbering."""
        d = self.to_internal(d_lex, F); x = np.zeros_like(d)
        lib().hh_ssor_apply(self.h, F, _d(np.ascontiguousarray(vals, dtype=np.float64)), steps, _d(d), _d(x))
        return self.to_external(x, F)

    def ilu0_apply(self, F, vals, d_lex):
        vals = np.asarray(vals, dtype=np.float64).reshape(-1, self.nslots)
        if F == 3:  # 7 stored planes -> the 9 blocks of PDELab's pattern
            lu = np.zeros((9, self.nslots))
            for (f, g), pl in PNP_PLANE.items():
                lu[3 * f + g] = vals[pl]
        else:
            lu = vals.copy()
        lu = np.ascontiguousarray(lu)
        d = self.to_internal(d_lex, F); x = np.zeros_like(d)
        lib().hh_ilu0_apply(self.h, F, _d(lu), _d(d), _d(x))
        return self.to_external(x, F)


def pk_element(degree, op, intorder, xy, phys, xl, caux, fflag, fflux, fdir, comp0=0, mode=0, eps=1e-11):
    """The product's quadratic / cubic element functions on one triangle: (element vector, element matrix)."""
    nl = (degree + 1) * (degree + 2) // 2
    n = nl * (3 if op == 4 else 1)
    r = np.zeros(n); A = np.zeros(n * n)
    xy = np.ascontiguousarray(xy, dtype=np.float64); phys = np.ascontiguousarray(phys, dtype=np.float64)
    xl = np.ascontiguousarray(xl, dtype=np.float64); caux = np.ascontiguousarray(caux, dtype=np.float64)
    fflag = np.ascontiguousarray(fflag, dtype=np.int32); fdir = np.ascontiguousarray(fdir, dtype=np.int32)
    fflux = np.ascontiguousarray(fflux, dtype=np.float64)
    lib().hh_pk_element(degree, op, intorder, _d(xy), _d(phys), _d(xl), _d(caux), _i(fflag), _d(fflux), _i(fdir), comp0, mode,
                        C.c_double(eps), _d(r), _d(A))
    return r, A.reshape(n, n)


def pk_node(degree, n):
    key = np.zeros(3, dtype=np.int32); xy = np.zeros(2)
    lib().hh_pk_node(degree, n, _i(key), _d(xy))
    return tuple(int(k) for k in key), (xy[0], xy[1])
