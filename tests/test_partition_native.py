"""The library's native partitioner (csrc/pnp_partition.cu, C ABI pnp_part_*) against the numpy statement of the same
decomposition (dune_pnp_b200/partition.py): every array of every level's plan is identical, for 1 to 5 ranks.  Host code
only -- no GPU involved; the ranks are threads exchanging through a barrier, as in tests/test_gpu_partitioned.py."""
import threading

import numpy as np
import pytest

import util


def _run_ranks(world, fn):
    out = [None] * world
    barrier = threading.Barrier(world)
    lock = threading.Lock()
    store = {}
    errs = []

    def run(rank):
        calls = [0]

        def all_gather(obj):
            key = calls[0]; calls[0] += 1
            with lock:
                store.setdefault(key, [None] * world)[rank] = obj
            barrier.wait()
            res = list(store[key])
            barrier.wait()
            return res
        try:
            out[rank] = fn(rank, all_gather)
        except Exception as e:  # pragma: no cover
            errs.append(e); barrier.abort()
    ts = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    [t.start() for t in ts]; [t.join() for t in ts]
    if errs:
        raise errs[0]
    return out


@pytest.mark.parametrize("name,world,levels", [("pore_small", 1, 2), ("pore_small", 2, 2), ("pore", 3, 1), ("pore", 4, 2),
                                               ("pore_without_dna", 5, 2), ("cylinder", 8, 1)])
def test_native_partitioner_equals_numpy_partitioner(name, world, levels):
    from dune_pnp_b200 import partition
    a = util.load_mesh_arrays(name)

    def fields(x, y):
        return {"u": np.stack([np.sin(0.3 * x) + y, np.cos(0.2 * y) * x, 0.01 * x * y])}
    ref = _run_ranks(world, lambda r, ag: partition.build_hierarchy(a, world, r, levels, all_gather=ag, fields_at=(0, fields)))
    nat = _run_ranks(world, lambda r, ag: partition.build_hierarchy_native(a, world, r, levels, all_gather=ag, fields_at=(0, fields)))
    total_owned = 0
    for r in range(world):
        assert len(ref[r]) == len(nat[r]) == levels + 1
        for l, (p, q) in enumerate(zip(ref[r], nat[r])):
            assert (p.nv, p.n_own, p.n_global) == (q.nv, q.n_own, q.n_global)
            for k in ("x", "y", "tri", "ba", "bb", "bphys", "nbr", "send_ptr", "send_idx", "recv_ptr"):
                assert np.array_equal(getattr(p, k), getattr(q, k)), (r, l, k)
            if l == 0:
                assert np.array_equal(p.gid, q.gid)
            else:
                assert np.array_equal(p.par, q.par), (r, l)
            assert np.array_equal(p.fields["u"], q.fields["u"]), (r, l)
        total_owned += nat[r][-1].n_own
    # owned vertices tile the global mesh
    from oracle import binding as ora
    assert total_owned == ora.Mesh.from_arrays(**a).refine(levels).nv


def test_native_partitioner_errors():
    import ctypes as C
    from dune_pnp_b200 import capi
    L = capi.lib()
    h = C.c_void_p()
    a = util.load_mesh_arrays("one_wall")
    x = np.ascontiguousarray(a["x"]); y = np.ascontiguousarray(a["y"]); tri = np.ascontiguousarray(a["tri"], dtype=np.int32)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    st = L.pnp_part_create(C.c_long(len(x)), x.ctypes.data_as(dp), y.ctypes.data_as(dp), C.c_long(len(tri)), tri.ctypes.data_as(ip),
                           C.c_long(0), None, None, None, 2, 5, 1, C.byref(h))   # rank outside the world
    assert st == 8
