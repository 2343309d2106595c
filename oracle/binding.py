"""ctypes binding of the CPU ORACLE (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product (dune_pnp_b200/) never imports this.
PARITY UNPINNED -- see the header of oracle/pnp_oracle.hpp.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

OP_PB, OP_POISSON, OP_DIFFUSION, OP_MASS, OP_PNP = range(5)
SOLVER_BCGS, SOLVER_CG = 0, 1
PREC_NONE, PREC_JACOBI, PREC_SSOR, PREC_ILU0 = 0, 1, 2, 3

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def build():
    subprocess.check_call(["make", "-C", _HERE, "-s"])


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        srcs = [os.path.join(_HERE, f) for f in ("oracle_capi.cpp", "pnp_oracle.hpp", "pnp_oracle_p2.hpp")]
        if not os.path.exists(so) or any(os.path.getmtime(f) > os.path.getmtime(so) + 1.0 for f in srcs if os.path.exists(f)):
            build()  # (a stale library would silently test yesterday's restatement)
        L = C.CDLL(so)
        L.ora_last_error.restype = C.c_char_p
        for f in ("ora_mesh_create", "ora_mesh_read_gmsh", "ora_mesh_refine", "ora_params_read", "ora_params_create"):
            getattr(L, f).restype = C.c_void_p
        L.ora_pattern.restype = C.c_long
        L.ora_pattern_sets.restype = C.c_long
        L.ora_bench_create.restype = C.c_void_p
        L.orak_pattern.restype = C.c_long
        L.ora_bench_nnz.restype = C.c_long
        L.ora_set_threads(1)
        _LIB = L
    return _LIB


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _i(a):
    return None if a is None else a.ctypes.data_as(_ip)


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _chk(rv):
    if rv is None or rv == -1:
        raise RuntimeError("oracle: " + lib().ora_last_error().decode())
    return rv


class Mesh:
    def __init__(self, handle):
        self.h = C.c_void_p(_chk(handle))
        nv, nT, nB = C.c_int(), C.c_int(), C.c_int()
        lib().ora_mesh_sizes(self.h, C.byref(nv), C.byref(nT), C.byref(nB))
        self.nv, self.nT, self.nB = nv.value, nT.value, nB.value
        self.x = np.empty(self.nv); self.y = np.empty(self.nv)
        self.tri = np.empty((self.nT, 3), dtype=np.int32)
        self.ba = np.empty(self.nB, dtype=np.int32); self.bb = np.empty(self.nB, dtype=np.int32)
        self.bphys = np.empty(self.nB, dtype=np.int32)
        lib().ora_mesh_get(self.h, _d(self.x), _d(self.y), _i(self.tri), _i(self.ba), _i(self.bb), _i(self.bphys))

    @staticmethod
    def from_arrays(x, y, tri, ba, bb, bphys):
        x = _f64(x); y = _f64(y)
        tri = np.ascontiguousarray(tri, dtype=np.int32); ba = np.ascontiguousarray(ba, dtype=np.int32)
        bb = np.ascontiguousarray(bb, dtype=np.int32); bphys = np.ascontiguousarray(bphys, dtype=np.int32)
        return Mesh(lib().ora_mesh_create(len(x), _d(x), _d(y), len(tri), _i(tri), len(ba), _i(ba), _i(bb), _i(bphys)))

    @staticmethod
    def read_gmsh(path):
        return Mesh(lib().ora_mesh_read_gmsh(path.encode()))

    def refine(self, k=1):
        m = self
        for _ in range(k):
            m = Mesh(lib().ora_mesh_refine(m.h))
        return m

    def __del__(self):
        try:
            lib().ora_mesh_free(self.h)
        except Exception:
            pass


class Params:
    """sys[16] / surf[ns][9] flat layout documented in oracle_capi.cpp."""

    def __init__(self, handle):
        self.h = C.c_void_p(_chk(handle))
        ns = lib().ora_params_nsurf(self.h)
        self.sys = np.zeros(16); self.surf = np.zeros((ns, 9))
        buf = C.create_string_buffer(512)
        lib().ora_params_get(self.h, _d(self.sys), _d(self.surf), buf, 512)
        self.meshfile = buf.value.decode()

    @staticmethod
    def read(path):
        return Params(lib().ora_params_read(path.encode()))

    @staticmethod
    def from_flat(sys, surf):
        sys = _f64(sys); surf = _f64(surf)
        return Params(lib().ora_params_create(_d(sys), _d(surf)))

    def __del__(self):
        try:
            lib().ora_params_free(self.h)
        except Exception:
            pass


def nfields(op):
    return 3 if op == OP_PNP else 1


def dirichlet(mesh, params, fields, comp0=0):
    out = np.zeros(fields * mesh.nv, dtype=np.int8)
    _chk(lib().ora_dirichlet(mesh.h, params.h, fields, comp0, out.ctypes.data_as(C.c_char_p)))
    return out.astype(bool)


def pattern(mesh, params, fields, comp0=0, literal=False):
    """literal=True: the std::set-per-row restatement of PDELab's pattern rule (slow; cross-check of the fast builder)."""
    fn = lib().ora_pattern_sets if literal else lib().ora_pattern
    N = fields * mesh.nv
    rowptr = np.zeros(N + 1, dtype=np.int32)
    nnz = _chk(fn(mesh.h, params.h, fields, comp0, _i(rowptr), None))
    col = np.zeros(nnz, dtype=np.int32)
    _chk(fn(mesh.h, params.h, fields, comp0, _i(rowptr), _i(col)))
    return rowptr, col


def residual(mesh, params, op, u, aux0=None, aux1=None, valency=1.0, intorder=-1, comp0=0, want_abs=False):
    u = _f64(u); aux0 = _f64(aux0); aux1 = _f64(aux1)
    r = np.zeros_like(u); ab = np.zeros_like(u) if want_abs else None
    _chk(lib().ora_residual(mesh.h, params.h, op, comp0, _d(u), _d(aux0), _d(aux1), C.c_double(valency), intorder,
                            _d(r), _d(ab)))
    return (r, ab) if want_abs else r


def jacobian(mesh, params, op, u, aux0=None, aux1=None, valency=1.0, intorder=-1, comp0=0, mode=0, eps=1e-11,
             want_abs=False):
    u = _f64(u); aux0 = _f64(aux0); aux1 = _f64(aux1)
    rowptr, col = pattern(mesh, params, nfields(op), comp0)
    val = np.zeros(len(col)); ab = np.zeros(len(col)) if want_abs else None
    _chk(lib().ora_jacobian(mesh.h, params.h, op, comp0, _d(u), _d(aux0), _d(aux1), C.c_double(valency), intorder,
                            mode, C.c_double(eps), _d(val), _d(ab)))
    return (rowptr, col, val, ab) if want_abs else (rowptr, col, val)


def interpolate(mesh, params, comp, pb=None):
    pb = _f64(pb)
    u = np.zeros(mesh.nv)
    _chk(lib().ora_interpolate(mesh.h, params.h, comp, _d(pb), _d(u)))
    return u


def linsolve(rowptr, col, val, b, reduction, maxit, solver=SOLVER_BCGS, prec=PREC_NONE, steps=1, x0=None):
    n = len(rowptr) - 1
    x = np.zeros(n) if x0 is None else _f64(x0).copy()
    bb = _f64(b).copy(); res = np.zeros(8)
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int32); col = np.ascontiguousarray(col, dtype=np.int32)
    val = _f64(val)
    _chk(lib().ora_linsolve(n, _i(rowptr), _i(col), _d(val), _d(x), _d(bb), C.c_double(reduction), maxit, solver,
                            prec, steps, _d(res)))
    return x, dict(converged=bool(res[0]), iterations=int(res[1]), reduction=res[2], conv_rate=res[3],
                   status=int(res[4]), seconds=res[5])


def prec_apply(rowptr, col, val, d, prec, steps=1):
    """v = M^-1 d, one application of an ISTL preconditioner (PREC_*)."""
    n = len(rowptr) - 1
    v = np.zeros(n); d = _f64(d); val = _f64(val)
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int32); col = np.ascontiguousarray(col, dtype=np.int32)
    _chk(lib().ora_prec_apply(n, _i(rowptr), _i(col), _d(val), prec, steps, _d(d), _d(v)))
    return v


def spmv(rowptr, col, val, x):
    n = len(rowptr) - 1
    y = np.zeros(n); x = _f64(x); val = _f64(val)
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int32); col = np.ascontiguousarray(col, dtype=np.int32)
    lib().ora_spmv(n, _i(rowptr), _i(col), _d(val), _d(x), _d(y))
    return y


def newton_opts(params, solver=SOLVER_BCGS, prec=PREC_NONE, prec_steps=1, jac_mode=0, fd_eps=1e-11, maxit=None,
                abs_limit=1e-12):
    """Newton set-up as the reference drivers do it (stationary_pnp_from_pb.hh:172-181, :344-360)."""
    s = params.sys
    return np.array([s[7], abs_limit, s[8], s[6], s[9] if maxit is None else maxit, s[10], 0.5, jac_mode, fd_eps,
                     solver, prec, prec_steps, s[5], s[15], 0, 0], dtype=np.float64)


def newton(mesh, params, op, u0, opts, aux0=None, aux1=None, valency=1.0, intorder=-1, comp0=0, cap=128):
    u = _f64(u0).copy(); aux0 = _f64(aux0); aux1 = _f64(aux1)
    res = np.zeros(16); hist = np.zeros(cap); lin = np.zeros(cap, dtype=np.int32)
    opts = _f64(opts)
    _chk(lib().ora_newton(mesh.h, params.h, op, comp0, _d(u), _d(aux0), _d(aux1), C.c_double(valency), intorder,
                          _d(opts), _d(res), _d(hist), _i(lin), cap))
    keys = ["status", "converged", "iterations", "first_defect", "defect", "reduction", "total_linear_iterations",
            "total_ls_trials", "jacobian_assemblies", "residual_assemblies", "seconds"]
    out = dict(zip(keys, res[:11]))
    for k in ("status", "iterations", "total_linear_iterations", "total_ls_trials", "jacobian_assemblies",
              "residual_assemblies"):
        out[k] = int(out[k])
    out["converged"] = bool(out["converged"])
    out["defect_history"] = hist[hist >= 0].copy()
    out["lin_iter_history"] = lin[lin >= 0].copy()
    return u, out


def slp(mesh, params, op, u0, reduction, solver=SOLVER_BCGS, prec=PREC_SSOR, steps=1, maxit=5000, aux0=None,
        aux1=None, valency=1.0, intorder=-1, comp0=0, jac_mode=0, eps=1e-11):
    u = _f64(u0).copy(); aux0 = _f64(aux0); aux1 = _f64(aux1); res = np.zeros(8)
    _chk(lib().ora_slp(mesh.h, params.h, op, comp0, _d(u), _d(aux0), _d(aux1), C.c_double(valency), intorder,
                       C.c_double(reduction), solver, prec, steps, maxit, jac_mode, C.c_double(eps), _d(res)))
    return u, dict(converged=bool(res[0]), iterations=int(res[1]), reduction=res[2])


def onestep(mesh, params, xold, g, phi, valency, dt, reduction=1e-5, method=0, solver=SOLVER_BCGS, prec=PREC_SSOR, steps=1,
            maxit=5000, jac_mode=0, eps=1e-11, comp0=1):
    """OneStepMethod<Alexander2>::apply (method 0; 1: implicit Euler) for the Nernst-Planck transport of one species."""
    xnew = np.zeros(mesh.nv); res = np.zeros(8)
    _chk(lib().ora_onestep(mesh.h, params.h, comp0, method, C.c_double(dt), _d(_f64(xold)), _d(_f64(g)), _d(_f64(phi)),
                           C.c_double(valency), _d(xnew), C.c_double(reduction), solver, prec, steps, maxit, jac_mode,
                           C.c_double(eps), _d(res)))
    nst = 1 if method == 1 else 2
    return xnew, [dict(converged=bool(res[2 * k]), iterations=int(res[2 * k + 1])) for k in range(nst)]


def ion_flux(mesh, params, phi, cp, cm):
    """calcIonFlux: (ip, im) per surface."""
    ns = int(params.sys[0])
    ip = np.zeros(ns); im = np.zeros(ns)
    _chk(lib().ora_ion_flux(mesh.h, params.h, _d(_f64(phi)), _d(_f64(cp)), _d(_f64(cm)), _d(ip), _d(im)))
    return ip, im


def write_cell_data(mesh, u, filename):
    _chk(lib().ora_write_cell_data(mesh.h, _d(_f64(u)), filename.encode()))


def write_vtk(mesh, name, fields, names, ascii=False):
    """Dune::VTKWriter vertex data: writes <name>.vtu."""
    arrs = [_f64(f) for f in fields]
    ptrs = (_dp * len(arrs))(*[_d(a) for a in arrs])
    nm = (C.c_char_p * len(names))(*[n.encode() for n in names])
    _chk(lib().ora_write_vtk(mesh.h, name.encode(), len(arrs), ptrs, nm, int(ascii)))


def max_threads():
    return int(lib().ora_max_threads())


def assembly_par(mesh, params, op, u, threads, mode=0):
    """Multi-core (OpenMP, two-phase) residual and Jacobian; must equal residual() / jacobian() bit for bit."""
    u = _f64(u)
    rowptr, col = pattern(mesh, params, nfields(op))
    r = np.zeros_like(u); val = np.zeros(len(col))
    _chk(lib().ora_assembly_par(mesh.h, params.h, op, _d(u), threads, mode, _d(r), _d(val)))
    return r, val


class Bench:
    """Phases of one PNP Newton step on `threads` host cores (CPU baseline of bench.py)."""

    def __init__(self, mesh, params, op=OP_PNP):
        self.mesh, self.params = mesh, params
        self.h = C.c_void_p(_chk(lib().ora_bench_create(mesh.h, params.h, op)))
        self.nnz = lib().ora_bench_nnz(self.h)

    def step(self, u, threads=1, jac_mode=0, krylov_iters=5, spmv_reps=3):
        out = np.zeros(8)
        _chk(lib().ora_bench_step(self.h, _d(_f64(u)), threads, jac_mode, krylov_iters, spmv_reps, _d(out)))
        return dict(jacobian_s=out[0], residual_s=out[1], krylov_s=out[2], spmv_s=out[3], total_s=out[4], defect=out[5],
                    reduction=out[6], krylov_iterations=int(out[7]))

    def __del__(self):
        try:
            lib().ora_bench_free(self.h)
        except Exception:
            pass


def sinh_shared(x):
    """The oracle's own sinh (pb_operator.hh:117 stand-in shared bit for bit with the device kernels)."""
    x = _f64(x); y = np.zeros_like(x)
    lib().ora_sinh_shared(len(x), _d(x), _d(y))
    return y


# ---------------------------------------------------------------------------------------------------------------
# quadratic elements (PDEGREE = 2, pnp_oracle_p2.hpp): dofs = edges [0, nE) then vertices [nE, nE + nv); fields lexicographic
# ---------------------------------------------------------------------------------------------------------------
class P2:
    """Pk space, k = degree (2 or 3).  Scalar dofs: [nT element bubbles (k = 3)] + (k-1) per edge from `eoff` + vertices from `voff`."""

    def __init__(self, mesh, params, degree=2):
        self.mesh, self.params, self.degree = mesh, params, int(degree)
        self.NL = (degree + 1) * (degree + 2) // 2
        sz = (C.c_long * 4)()
        _chk(lib().orak_sizes(self.degree, mesh.h, params.h, sz))
        self.nE, self.nd, self.voff, self.eoff = int(sz[0]), int(sz[1]), int(sz[2]), int(sz[3])
        self.eva = np.zeros(self.nE, dtype=np.int32); self.evb = np.zeros(self.nE, dtype=np.int32)
        self.tedge = np.zeros((mesh.nT, 3), dtype=np.int32)
        self.x = np.zeros(self.nd); self.y = np.zeros(self.nd)
        _chk(lib().orak_space(self.degree, mesh.h, params.h, _i(self.eva), _i(self.evb), _i(self.tedge), _d(self.x), _d(self.y)))

    def dirichlet(self, fields, comp0=0):
        out = np.zeros(fields * self.nd, dtype=np.int8)
        _chk(lib().orak_dirichlet(self.degree, self.mesh.h, self.params.h, fields, comp0, out.ctypes.data_as(C.c_char_p)))
        return out.astype(bool)

    def pattern(self, fields, comp0=0):
        rowptr = np.zeros(fields * self.nd + 1, dtype=np.int32)
        nnz = _chk(lib().orak_pattern(self.degree, self.mesh.h, self.params.h, fields, comp0, _i(rowptr), None))
        col = np.zeros(nnz, dtype=np.int32)
        _chk(lib().orak_pattern(self.degree, self.mesh.h, self.params.h, fields, comp0, _i(rowptr), _i(col)))
        return rowptr, col

    def residual(self, op, u, aux0=None, aux1=None, valency=1.0, intorder=-1, comp0=0, want_abs=False):
        u = _f64(u); aux0 = _f64(aux0); aux1 = _f64(aux1)
        r = np.zeros_like(u); ab = np.zeros_like(u) if want_abs else None
        _chk(lib().orak_residual(self.degree, self.mesh.h, self.params.h, op, comp0, _d(u), _d(aux0), _d(aux1), C.c_double(valency), intorder,
                                 _d(r), _d(ab)))
        return (r, ab) if want_abs else r

    def jacobian(self, op, u, aux0=None, aux1=None, valency=1.0, intorder=-1, comp0=0, mode=0, eps=1e-11, want_abs=False):
        u = _f64(u); aux0 = _f64(aux0); aux1 = _f64(aux1)
        rowptr, col = self.pattern(nfields(op), comp0)
        val = np.zeros(len(col)); ab = np.zeros(len(col)) if want_abs else None
        _chk(lib().orak_jacobian(self.degree, self.mesh.h, self.params.h, op, comp0, _d(u), _d(aux0), _d(aux1), C.c_double(valency), intorder,
                                 mode, C.c_double(eps), _d(val), _d(ab)))
        return (rowptr, col, val, ab) if want_abs else (rowptr, col, val)

    def interpolate(self, comp, pb=None):
        pb = _f64(pb)
        u = np.zeros(self.nd)
        _chk(lib().orak_interpolate(self.degree, self.mesh.h, self.params.h, comp, _d(pb), _d(u)))
        return u

    def onestep(self, xold, g, phi, valency, dt, reduction=1e-5, method=0, solver=SOLVER_BCGS, prec=PREC_SSOR, steps=1, maxit=5000,
                jac_mode=0, eps=1e-11, comp0=1, intorder=-1):
        """OneStepMethod::apply (method 0: Alexander2, 1: implicit Euler) for the transport of one species on this space."""
        xnew = np.zeros(self.nd); res = np.zeros(8)
        _chk(lib().orak_onestep(self.degree, self.mesh.h, self.params.h, comp0, method, C.c_double(dt), _d(_f64(xold)), _d(_f64(g)),
                                _d(_f64(phi)), C.c_double(valency), intorder, _d(xnew), C.c_double(reduction), solver, prec, steps,
                                maxit, jac_mode, C.c_double(eps), _d(res)))
        nst = 1 if method == 1 else 2
        return xnew, [dict(converged=bool(res[2 * k]), iterations=int(res[2 * k + 1])) for k in range(nst)]

    def ion_flux(self, phi, cp, cm):
        ns = int(self.params.sys[0])
        ip = np.zeros(ns); im = np.zeros(ns)
        _chk(lib().orak_ion_flux(self.degree, self.mesh.h, self.params.h, _d(_f64(phi)), _d(_f64(cp)), _d(_f64(cm)), _d(ip), _d(im)))
        return ip, im

    def write_cell_data(self, u, filename):
        _chk(lib().orak_write_cell_data(self.degree, self.mesh.h, self.params.h, _d(_f64(u)), filename.encode()))

    def newton(self, op, u0, opts, aux0=None, aux1=None, valency=1.0, intorder=-1, comp0=0, cap=128):
        u = _f64(u0).copy(); aux0 = _f64(aux0); aux1 = _f64(aux1)
        res = np.zeros(16); hist = np.zeros(cap); lin = np.zeros(cap, dtype=np.int32)
        _chk(lib().orak_newton(self.degree, self.mesh.h, self.params.h, op, comp0, _d(u), _d(aux0), _d(aux1), C.c_double(valency), intorder,
                               _d(_f64(opts)), _d(res), _d(hist), _i(lin), cap))
        keys = ["status", "converged", "iterations", "first_defect", "defect", "reduction", "total_linear_iterations",
                "total_ls_trials", "jacobian_assemblies", "residual_assemblies", "seconds"]
        out = dict(zip(keys, res[:11]))
        for k in ("status", "iterations", "total_linear_iterations", "total_ls_trials", "jacobian_assemblies", "residual_assemblies"):
            out[k] = int(out[k])
        out["converged"] = bool(out["converged"])
        out["defect_history"] = hist[hist >= 0].copy(); out["lin_iter_history"] = lin[lin >= 0].copy()
        return u, out


def p2_basis(x, y, degree=2):
    nl = (degree + 1) * (degree + 2) // 2
    phi = np.zeros(nl); g = np.zeros(2 * nl)
    lib().orak_basis(int(degree), C.c_double(x), C.c_double(y), _d(phi), _d(g))
    return phi, g.reshape(nl, 2)


def P3(mesh, params):
    return P2(mesh, params, 3)
