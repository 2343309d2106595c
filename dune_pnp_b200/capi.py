"""Thin ctypes binding of libpnp_b200.so (C ABI: include/pnp_b200.h).

This is harness plumbing for tests/ and bench.py; the product is the CUDA library.  There is no CPU
path: loading works anywhere (so symbols can be checked), creating a context needs a CUDA device.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpnp_b200.so")

OP_PB, OP_POISSON, OP_DIFFUSION, OP_MASS, OP_PNP = range(5)
JAC_FD_FAITHFUL, JAC_ANALYTIC = 0, 1
SOLVER_BCGS, SOLVER_CG = 0, 1
PREC_NONE, PREC_JACOBI, PREC_SSOR, PREC_ILU0, PREC_AMG = range(5)
TIME_ALEXANDER2, TIME_IMPLICIT_EULER = 0, 1
STATUS = {0: "PNP_OK", 1: "PNP_E_NOT_CONVERGED", 2: "PNP_E_LINEAR_SOLVER", 3: "PNP_E_LINE_SEARCH", 4: "PNP_E_NAN",
          5: "PNP_E_BREAKDOWN", 6: "PNP_E_CUDA", 7: "PNP_E_CONFIG", 8: "PNP_E_ARG", 9: "PNP_E_MESH"}

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


class LinResult(C.Structure):
    _fields_ = [("converged", C.c_int), ("iterations", C.c_int), ("reduction", C.c_double), ("conv_rate", C.c_double),
                ("seconds", C.c_double), ("status", C.c_int)]


class NewtonOpts(C.Structure):
    _fields_ = [("reduction", C.c_double), ("abs_limit", C.c_double), ("min_linear_reduction", C.c_double),
                ("reassemble_threshold", C.c_double), ("max_iterations", C.c_int),
                ("line_search_max_iterations", C.c_int), ("damping", C.c_double), ("jac_mode", C.c_int),
                ("fd_epsilon", C.c_double), ("verbosity", C.c_int), ("line_search_strategy", C.c_int)]


class NewtonResult(C.Structure):
    _fields_ = [("converged", C.c_int), ("iterations", C.c_int), ("first_defect", C.c_double), ("defect", C.c_double),
                ("reduction", C.c_double), ("linear_iterations", C.c_int), ("line_search_trials", C.c_int),
                ("jacobian_assemblies", C.c_int), ("residual_assemblies", C.c_int), ("seconds_assembly", C.c_double),
                ("seconds_solve", C.c_double), ("seconds_total", C.c_double), ("n_history", C.c_int),
                ("defect_history", C.c_double * 64), ("linear_iterations_history", C.c_int * 64)]


class PnpError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__("%s: %s" % (STATUS.get(status, status), msg))
        self.status = status


_LIB = None


def lib():
    """Loads the CUDA library; raises if it has not been built (there is no fallback)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libpnp_b200.so is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(make -C dune_pnp_b200/csrc)")
        # libpnp_b200.so needs libnccl.so.2 and libcusolver.so.11 (+ cublas/cusparse).  PyTorch bundles its own copies
        # under the same sonames; whichever copy is mapped first serves both, so map the bundled ones first (otherwise a
        # later `import torch` would find older system copies already mapped and fail to resolve its newer symbols).
        try:
            import importlib.util
            for pkg, names in (("nvidia.nccl", ["libnccl.so.2"]), ("nvidia.nvjitlink", ["libnvJitLink.so.12"]),
                               ("nvidia.cublas", ["libcublasLt.so.12", "libcublas.so.12"]),
                               ("nvidia.cusparse", ["libcusparse.so.12"]), ("nvidia.cusolver", ["libcusolver.so.11"])):
                spec = importlib.util.find_spec(pkg)
                if spec is None or not spec.submodule_search_locations:
                    continue
                for nm in names:
                    cand = os.path.join(list(spec.submodule_search_locations)[0], "lib", nm)
                    if os.path.exists(cand):
                        C.CDLL(cand, mode=C.RTLD_GLOBAL)
        except Exception:
            pass
        L = C.CDLL(LIB_PATH)
        L.pnp_last_error.restype = C.c_char_p
        L.pnp_launch_count.restype = C.c_long
        _LIB = L
    return _LIB


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _i(a):
    return None if a is None else a.ctypes.data_as(_ip)


def tune(name, value):
    """Process-wide kernel tuning knob (pnp_tune)."""
    st = lib().pnp_tune(name.encode(), C.c_double(value))
    if st != 0:
        raise PnpError(st, "unknown tuning knob " + name)


class Context:
    """One GPU, one mesh.  Mirrors the C ABI one to one; numpy arrays in the reference's numbering."""

    def __init__(self, device=0, parent=None):
        self._h = C.c_void_p()
        self._children = []
        if parent is not None:
            st = lib().pnp_ctx_create_child(parent._h, C.byref(self._h))
            parent._children.append(self)
        else:
            st = lib().pnp_ctx_create(device, C.byref(self._h))
        if st != 0:
            raise PnpError(st, "cannot create a context on CUDA device %d (no CPU fallback exists)" % device)

    def close(self):
        if self._h:
            lib().pnp_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, st):
        if st != 0:
            raise PnpError(st, lib().pnp_last_error(self._h).decode())

    def profile_spmv(self, enable):
        self._ck(lib().pnp_profile_spmv(self._h, int(enable)))

    def profile_spmv_get(self):
        n = (C.c_long * 3)(); ms = (C.c_double * 3)()
        self._ck(lib().pnp_profile_spmv_get(self._h, n, ms))
        return list(n), list(ms)

    def profile_bytes(self, reset=False):
        out = (C.c_double * 6)()
        self._ck(lib().pnp_profile_bytes(self._h, int(reset), out))
        return dict(zip(("spmv_fine", "spmv_coarse", "assembly", "blas1", "transfer", "dense"), list(out)))

    def profiler_range(self, start):
        self._ck(lib().pnp_profiler_range(self._h, int(start)))

    def timer_start(self):
        self._ck(lib().pnp_timer_start(self._h))

    def timer_stop(self):
        ms = C.c_double()
        self._ck(lib().pnp_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def launch_count(self):
        return lib().pnp_launch_count(self._h)

    # ---- mesh ----
    def mesh_set(self, x, y, tri, ba, bb, bphys):
        x = np.ascontiguousarray(x, dtype=np.float64); y = np.ascontiguousarray(y, dtype=np.float64)
        tri = np.ascontiguousarray(tri, dtype=np.int32); ba = np.ascontiguousarray(ba, dtype=np.int32)
        bb = np.ascontiguousarray(bb, dtype=np.int32); bphys = np.ascontiguousarray(bphys, dtype=np.int32)
        self._ck(lib().pnp_mesh_set(self._h, C.c_long(len(x)), _d(x), _d(y), C.c_long(len(tri)), _i(tri), C.c_long(len(ba)),
                                    _i(ba), _i(bb), _i(bphys)))

    def mesh_read_gmsh(self, path):
        self._ck(lib().pnp_mesh_read_gmsh(self._h, path.encode()))

    def mesh_refine(self, levels):
        self._ck(lib().pnp_mesh_refine(self._h, levels))

    def carry_set(self, vecs):
        arr = (C.c_int * len(vecs))(*vecs)
        self._ck(lib().pnp_carry_set(self._h, arr, len(vecs)))

    def carry_get(self, index, vec):
        self._ck(lib().pnp_carry_get(self._h, index, vec))

    def carry_set_host(self, field):
        field = np.ascontiguousarray(field, dtype=np.float64)
        self._ck(lib().pnp_carry_set_host(self._h, field.shape[0] if field.ndim == 2 else 1, _d(field)))

    def carry_get_host(self, index, fields):
        out = np.zeros((fields, self.mesh_sizes()["nv"]))
        self._ck(lib().pnp_carry_get_host(self._h, index, _d(out)))
        return out

    def mesh_set_local(self, n_own, x, y, tri, ba, bb, bphys):
        x = np.ascontiguousarray(x, dtype=np.float64); y = np.ascontiguousarray(y, dtype=np.float64)
        tri = np.ascontiguousarray(tri, dtype=np.int32); ba = np.ascontiguousarray(ba, dtype=np.int32)
        bb = np.ascontiguousarray(bb, dtype=np.int32); bphys = np.ascontiguousarray(bphys, dtype=np.int32)
        self._ck(lib().pnp_mesh_set_local(self._h, C.c_long(len(x)), C.c_long(n_own), _d(x), _d(y), C.c_long(len(tri)), _i(tri),
                                          C.c_long(len(ba)), _i(ba), _i(bb), _i(bphys)))

    def mesh_owned(self):
        n = C.c_long()
        self._ck(lib().pnp_mesh_owned(self._h, C.byref(n)))
        return n.value

    @staticmethod
    def comm_unique_id():
        buf = C.create_string_buffer(128)
        if lib().pnp_comm_unique_id(buf) != 0:
            raise PnpError(6, "ncclGetUniqueId failed")
        return buf.raw

    def comm_init(self, rank, world, unique_id):
        self._ck(lib().pnp_comm_init(self._h, rank, world, C.c_char_p(unique_id) if world > 1 else None))

    def halo_set(self, nbr, send_ptr, send_idx, recv_ptr):
        nbr = np.ascontiguousarray(nbr, dtype=np.int32); send_ptr = np.ascontiguousarray(send_ptr, dtype=np.int32)
        send_idx = np.ascontiguousarray(send_idx, dtype=np.int32); recv_ptr = np.ascontiguousarray(recv_ptr, dtype=np.int32)
        self._ck(lib().pnp_halo_set(self._h, len(nbr), _i(nbr), _i(send_ptr), _i(send_idx), _i(recv_ptr)))

    def mg_push_level(self, child, par0, par1):
        par0 = np.ascontiguousarray(par0, dtype=np.int32); par1 = np.ascontiguousarray(par1, dtype=np.int32)
        self._ck(lib().pnp_mg_push_level(self._h, child._h, _i(par0), _i(par1)))

    def mg_set_coarse_global(self, gid, n_global):
        gid = np.ascontiguousarray(gid, dtype=np.int32)
        self._ck(lib().pnp_mg_set_coarse_global(self._h, _i(gid), C.c_long(n_global)))

    def mg_set_coarse_replica(self, replica, gid, n_global):
        gid = np.ascontiguousarray(gid, dtype=np.int32)
        self._ck(lib().pnp_mg_set_coarse_replica(self._h, replica._h, _i(gid), C.c_long(n_global)))

    def mg_set_coarse_aggregates(self, agg, n_agg):
        agg = np.ascontiguousarray(agg, dtype=np.int32)
        self._ck(lib().pnp_mg_set_coarse_aggregates(self._h, _i(agg), C.c_long(n_agg)))

    def halo_exchange(self, vec):
        self._ck(lib().pnp_halo_exchange(self._h, vec))

    def mesh_finalize(self, renumber=True):
        self._ck(lib().pnp_mesh_finalize(self._h, int(renumber)))

    def mesh_sizes(self):
        v = [C.c_long() for _ in range(4)]
        self._ck(lib().pnp_mesh_sizes(self._h, *[C.byref(a) for a in v]))
        return dict(nv=v[0].value, nT=v[1].value, nB=v[2].value, nslots=v[3].value)

    def mesh_get(self):
        s = self.mesh_sizes()
        x = np.zeros(s["nv"]); y = np.zeros(s["nv"]); tri = np.zeros((s["nT"], 3), dtype=np.int32)
        ba = np.zeros(s["nB"], dtype=np.int32); bb = np.zeros(s["nB"], dtype=np.int32); ph = np.zeros(s["nB"], dtype=np.int32)
        self._ck(lib().pnp_mesh_get(self._h, _d(x), _d(y), _i(tri), _i(ba), _i(bb), _i(ph)))
        return dict(x=x, y=y, tri=tri, ba=ba, bb=bb, bphys=ph)

    # ---- parameters ----
    def params_set(self, sys, surf):
        sys = np.ascontiguousarray(sys, dtype=np.float64); surf = np.ascontiguousarray(surf, dtype=np.float64)
        self._ck(lib().pnp_params_set(self._h, _d(sys), _d(surf)))

    def params_read(self, path):
        self._ck(lib().pnp_params_read(self._h, path.encode()))

    def params_get(self):
        sys = np.zeros(16)
        self._ck(lib().pnp_params_get(self._h, _d(sys), None, None, 0))
        surf = np.zeros((int(sys[0]), 9)); buf = C.create_string_buffer(512)
        self._ck(lib().pnp_params_get(self._h, _d(sys), _d(surf), buf, 512))
        return sys, surf, buf.value.decode()

    # ---- operators / vectors / matrices ----
    def operator(self, op, comp0=0):
        h = C.c_int()
        self._ck(lib().pnp_operator_create(self._h, op, comp0, C.byref(h)))
        return h.value

    def operator_set_coefficient(self, op, which, vec):
        self._ck(lib().pnp_operator_set_coefficient(self._h, op, which, vec))

    def operator_set_valency(self, op, valency):
        self._ck(lib().pnp_operator_set_valency(self._h, op, C.c_double(valency)))

    def constraints(self, op, fields):
        out = np.zeros(fields * self.ndof(), dtype=np.int8)
        self._ck(lib().pnp_constraints_get(self._h, op, out.ctypes.data_as(C.c_char_p)))
        return out.astype(bool)

    def pattern(self, op, fields):
        nv = self.ndof()
        nnz = C.c_long(); rowptr = np.zeros(fields * nv + 1, dtype=np.int32)
        self._ck(lib().pnp_pattern_get(self._h, op, C.byref(nnz), _i(rowptr), None))
        col = np.zeros(nnz.value, dtype=np.int32)
        self._ck(lib().pnp_pattern_get(self._h, op, C.byref(nnz), _i(rowptr), _i(col)))
        return rowptr, col

    def vec(self, fields, host=None):
        h = C.c_int()
        self._ck(lib().pnp_vec_create(self._h, fields, C.byref(h)))
        if host is not None:
            self.upload(h.value, host)
        return h.value

    def vec_destroy(self, v):
        self._ck(lib().pnp_vec_destroy(self._h, v))

    def upload(self, v, host):
        host = np.ascontiguousarray(host, dtype=np.float64)
        self._ck(lib().pnp_vec_upload(self._h, v, _d(host)))

    def download(self, v, fields):
        out = np.zeros(fields * self.ndof())
        self._ck(lib().pnp_vec_download(self._h, v, _d(out)))
        return out

    def vec_set(self, v, value):
        self._ck(lib().pnp_vec_set(self._h, v, C.c_double(value)))

    def vec_copy(self, dst, src):
        self._ck(lib().pnp_vec_copy(self._h, dst, src))

    def axpy(self, y, a, x):
        self._ck(lib().pnp_vec_axpy(self._h, y, C.c_double(a), x))

    def norm(self, x):
        out = C.c_double()
        self._ck(lib().pnp_vec_norm(self._h, x, C.byref(out)))
        return out.value

    def dot(self, x, y):
        out = C.c_double()
        self._ck(lib().pnp_vec_dot(self._h, x, y, C.byref(out)))
        return out.value

    def pack3(self, dst3, phi, cp, cm):
        self._ck(lib().pnp_vec_pack3(self._h, dst3, phi, cp, cm))

    def extract(self, src3, field, dst1):
        self._ck(lib().pnp_vec_extract(self._h, src3, field, dst1))

    def matrix(self, op):
        h = C.c_int()
        self._ck(lib().pnp_matrix_create(self._h, op, C.byref(h)))
        return h.value

    def matrix_destroy(self, m):
        self._ck(lib().pnp_matrix_destroy(self._h, m))

    def residual(self, op, u, r):
        self._ck(lib().pnp_residual(self._h, op, u, r))

    def jacobian(self, op, u, A, mode=JAC_FD_FAITHFUL, eps=1e-11):
        self._ck(lib().pnp_jacobian(self._h, op, u, A, mode, C.c_double(eps)))

    def matrix_values(self, op, A, nnz):
        val = np.zeros(nnz)
        self._ck(lib().pnp_matrix_values_get(self._h, op, A, _d(val)))
        return val

    def spmv(self, A, x, y):
        self._ck(lib().pnp_spmv(self._h, A, x, y))

    # ---- solvers ----
    def solver(self, kind=SOLVER_BCGS, prec=PREC_NONE, maxit=5000, prec_steps=1, verbosity=0):
        h = C.c_int()
        self._ck(lib().pnp_solver_create(self._h, kind, prec, maxit, prec_steps, verbosity, C.byref(h)))
        return h.value

    def solver_set_option(self, solver, name, value):
        self._ck(lib().pnp_solver_set_option(self._h, solver, name.encode(), C.c_double(value)))

    def solver_get(self, solver, name):
        out = C.c_double(0)
        self._ck(lib().pnp_solver_get(self._h, solver, name.encode(), C.byref(out)))
        return out.value

    def precond_apply(self, solver, A, d, v):
        self._ck(lib().pnp_precond_apply(self._h, solver, A, d, v))

    def solve(self, solver, A, z, r, reduction):
        res = LinResult()
        self._ck(lib().pnp_solver_apply(self._h, solver, A, z, r, C.c_double(reduction), C.byref(res)))
        return res

    def newton_opts(self, **kw):
        o = NewtonOpts()
        self._ck(lib().pnp_newton_opts_from_params(self._h, C.byref(o)))
        for k, v in kw.items():
            setattr(o, k, v)
        return o

    def newton(self, op, u, solver, opts, check=True):
        res = NewtonResult()
        st = lib().pnp_newton_apply(self._h, op, u, solver, C.byref(opts), C.byref(res))
        if check:
            self._ck(st)
        return st, res

    def slp(self, op, u, solver, reduction, jac_mode=JAC_FD_FAITHFUL, eps=1e-11):
        res = LinResult()
        self._ck(lib().pnp_slp_apply(self._h, op, u, solver, C.c_double(reduction), jac_mode, C.c_double(eps), C.byref(res)))
        return res

    def onestep(self, op_space, op_time, solver, dt, x_old, dirichlet_values, x_new, reduction=1e-5, method=TIME_ALEXANDER2,
                jac_mode=JAC_FD_FAITHFUL, eps=1e-11, time=0.0):
        """OneStepMethod::apply; returns the stage solves' LinResults."""
        res = (LinResult * 2)()
        self._ck(lib().pnp_onestep_apply(self._h, method, op_space, op_time, solver, C.c_double(time), C.c_double(dt), x_old,
                                         dirichlet_values, x_new, C.c_double(reduction), jac_mode, C.c_double(eps), res))
        return list(res)[: (1 if method == TIME_IMPLICIT_EULER else 2)]

    def ion_flux(self, phi, cp, cm):
        ns = int(self.params_get()[0][0])
        ip = np.zeros(ns); im = np.zeros(ns)
        self._ck(lib().pnp_ion_flux(self._h, phi, cp, cm, _d(ip), _d(im)))
        return ip, im

    def write_cell_data(self, vec, filename):
        self._ck(lib().pnp_write_cell_data(self._h, vec, filename.encode()))

    def write_vtk(self, name, vecs, names, ascii=False):
        arr = (C.c_int * len(vecs))(*vecs)
        nm = (C.c_char_p * len(names))(*[n.encode() for n in names])
        self._ck(lib().pnp_write_vtk(self._h, name.encode(), len(vecs), arr, nm, int(ascii)))

    def matrix_set_csr(self, op, A, rowptr, col, val):
        rowptr = np.ascontiguousarray(rowptr, dtype=np.int32); col = np.ascontiguousarray(col, dtype=np.int32)
        val = np.ascontiguousarray(val, dtype=np.float64)
        self._ck(lib().pnp_matrix_set_csr(self._h, op, A, _i(rowptr), _i(col), _d(val)))

    def mesh_renumber(self, new_index):
        new_index = np.ascontiguousarray(new_index, dtype=np.int32)
        self._ck(lib().pnp_mesh_renumber(self._h, _i(new_index)))

    # ---- quadratic elements (-DPDEGREE=2) ----
    def space_set_degree(self, degree):
        self._ck(lib().pnp_space_set_degree(self._h, int(degree)))

    def space_sizes(self):
        deg = C.c_int(); nE = C.c_long(); nd = C.c_long()
        self._ck(lib().pnp_space_sizes(self._h, C.byref(deg), C.byref(nE), C.byref(nd)))
        return dict(degree=deg.value, n_edges=nE.value, ndof=nd.value)

    def space_offsets(self):
        eo = C.c_long(); vo = C.c_long()
        self._ck(lib().pnp_space_offsets(self._h, C.byref(eo), C.byref(vo)))
        return eo.value, vo.value

    def operator_set_intorder(self, op, intorder):
        self._ck(lib().pnp_operator_set_intorder(self._h, op, int(intorder)))

    def ndof(self):
        """Scalar dofs per field: vertices (degree 1), edges + vertices (2), elements + 2 per edge + vertices (3)."""
        return self.space_sizes()["ndof"]

    def space_edges(self):
        nE = self.space_sizes()["n_edges"]
        va = np.zeros(nE, dtype=np.int32); vb = np.zeros(nE, dtype=np.int32)
        self._ck(lib().pnp_space_edges(self._h, _i(va), _i(vb)))
        return va, vb

    def interpolate_bcext(self, component, pb_vec, out_vec):
        self._ck(lib().pnp_interpolate_bcext(self._h, component, -1 if pb_vec is None else pb_vec, out_vec))
