"""N > 1 host logic on CPU: world_size-2/3 `gloo` runs of the partitioner and halo plan (dune_pnp_b200/partition.py),
checked against the GLOBAL oracle: every vertex owned exactly once, ghost values delivered by the plan equal the
owner's values, rows assembled on a rank's local mesh (product per-item logic via the host harness, owned rows only)
equal the oracle's rows of the global mesh, and nested-iteration fields are interpolated consistently."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import util


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _field(x, y, k=0):
    return np.sin(0.3 * x + k) + np.cos(0.17 * y - k) * 0.5 + 0.01 * x * y


def _worker(rank, world, port, levels, q):
    try:
        import sys
        here = os.path.dirname(os.path.abspath(__file__))
        for p_ in (os.path.dirname(here), here):
            if p_ not in sys.path:
                sys.path.insert(0, p_)
        import harness
        from dune_pnp_b200 import partition
        from oracle import binding as ora
        os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)

        def all_gather(obj):
            out = [None] * world
            dist.all_gather_object(out, obj)
            return out

        a = util.load_mesh_arrays("pore_small")
        params = ora.Params.read(util.cfg_path("pore_small"))
        u_coarse = np.stack([_field(a["x"], a["y"], k) for k in range(3)])
        plan = partition.build_local(a, world, rank, levels, fields={"u": u_coarse}, all_gather=all_gather)
        # global references
        gm = ora.Mesh.from_arrays(**a)
        gu = u_coarse.reshape(-1)
        import bench
        for _ in range(levels):
            gu = bench.carry_numpy(gm, gu, 3)
            gm = gm.refine(1)
        gkeys = {k: i for i, k in enumerate(partition._coord_keys(gm.x, gm.y))}
        loc2glob = np.array([gkeys[k] for k in partition._coord_keys(plan.x, plan.y)])
        # (a) every global vertex owned exactly once
        cnt = torch.zeros(gm.nv, dtype=torch.int64)
        cnt[torch.from_numpy(loc2glob[:plan.n_own])] += 1
        dist.all_reduce(cnt)
        assert bool((cnt == 1).all()), "ownership is not a partition of the vertices"
        # (b) carried field == global interpolation
        assert np.array_equal(plan.fields["u"], gu.reshape(3, -1)[:, loc2glob])
        # (c) the halo plan delivers the owners' values
        val = np.full((plan.nv, 3), np.nan)
        val[:plan.n_own] = np.stack([_field(plan.x[:plan.n_own], plan.y[:plan.n_own], k) for k in range(3)], axis=1)
        reqs, recv_bufs = [], []
        for i, r in enumerate(plan.nbr.tolist()):
            s0, s1, r0, r1 = plan.send_ptr[i], plan.send_ptr[i + 1], plan.recv_ptr[i], plan.recv_ptr[i + 1]
            if s1 > s0:
                reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(val[plan.send_idx[s0:s1]])), dst=r))
            if r1 > r0:
                buf = torch.empty((r1 - r0, 3), dtype=torch.float64); recv_bufs.append((r0, r1, buf))
                reqs.append(dist.irecv(buf, src=r))
        for rq in reqs:
            rq.wait()
        for r0, r1, buf in recv_bufs:
            val[plan.n_own + r0:plan.n_own + r1] = buf.numpy()
        want = np.stack([_field(plan.x, plan.y, k) for k in range(3)], axis=1)
        assert np.array_equal(val, want), "ghost values differ from the owners' values"
        # (d) owned rows assembled on the local mesh == rows of the global problem
        la = dict(x=plan.x, y=plan.y, tri=plan.tri, ba=plan.ba, bb=plan.bb, bphys=plan.bphys)
        S = harness.Star(la, params.surf, renumber=True, n_own=plan.n_own)
        gd = ora.dirichlet(gm, params, 3).reshape(3, -1)
        ld = S.dirichlet(3).reshape(3, -1)
        # Dirichlet flags must be right on owned vertices AND on ghost columns adjacent to owned rows
        adj_cols = np.unique(S.int2ext[S.adj & harness.VMASK])
        assert np.array_equal(ld[:, adj_cols], gd[:, loc2glob[adj_cols]])
        for op in (ora.OP_PB, ora.OP_PNP):
            F = 3 if op == ora.OP_PNP else 1
            gfield = np.concatenate([_field(gm.x, gm.y, k) for k in range(F)])
            r_glob, ab = ora.residual(gm, params, op, gfield, want_abs=True)
            lfield = np.concatenate([_field(plan.x, plan.y, k) for k in range(F)])
            r_loc = S.residual(op, params.sys, lfield).reshape(F, -1)[:, :plan.n_own]
            g = loc2glob[:plan.n_own]
            err = np.abs(r_loc - r_glob.reshape(F, -1)[:, g])
            assert np.all(err <= 1e-12 * ab.reshape(F, -1)[:, g] + 1e-300), "owned residual rows differ from the global ones"
        # (e) distributed multigrid hierarchy: parents of every local vertex exist on the next coarser local level
        plans = partition.build_hierarchy(a, world, rank, levels, all_gather=all_gather)
        assert plans[-1].n_own == plan.n_own and np.array_equal(plans[-1].x, plan.x)
        assert np.array_equal(np.asarray(a["x"])[plans[0].gid], plans[0].x)
        for l in range(1, len(plans)):
            f, cz = plans[l], plans[l - 1]
            p0, p1 = f.par[:, 0], f.par[:, 1]
            assert p0.min() >= 0 and p0.max() < cz.nv and p1.max() < cz.nv
            one = p1 < 0
            assert np.array_equal(f.x[one], cz.x[p0[one]]) and np.array_equal(f.y[one], cz.y[p0[one]])
            assert np.array_equal(f.x[~one], 0.5 * (cz.x[np.minimum(p0, p1)[~one]] + cz.x[np.maximum(p0, p1)[~one]]))
            # same owner on every level: an owned coarse vertex is the parent of its (owned) coincident fine vertex ...
            coinc = np.where(one)[0]
            assert np.array_equal(coinc < f.n_own, p0[coinc] < cz.n_own)
            # ... and all children of an owned coarse vertex are present: weights of P^T restricted to owned rows sum up
            w = np.zeros(cz.nv); np.add.at(w, p0, np.where(one, 1.0, 0.5)); np.add.at(w, p1[~one], 0.5)
            tw = torch.tensor([w[:cz.n_own].sum()], dtype=torch.float64); dist.all_reduce(tw)
            nf = torch.tensor([float(f.n_own)], dtype=torch.float64); dist.all_reduce(nf)
            assert abs(tw.item() - nf.item()) < 1e-9, "restriction weights do not add up to the number of fine vertices"
        q.put((rank, "ok", plan.n_own, plan.nv))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "fail: " + traceback.format_exc(), 0, 0))
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()


@pytest.mark.parametrize("world,levels", [(2, 1), (3, 2)])
def test_partition_and_halo_plan(world, levels):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, levels, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for r in res:
        assert r[1] == "ok", r[1]
    assert sum(r[2] for r in res) > 0
    # ghost overhead stays a boundary effect
    assert all(r[3] < 2.0 * r[2] for r in res)


def test_rcb_partition_is_balanced():
    from dune_pnp_b200 import partition
    a = util.load_mesh_arrays("pore")
    tri = a["tri"]
    part = partition.rcb_partition(a["x"][tri].mean(1), a["y"][tri].mean(1), 8)
    counts = np.bincount(part, minlength=8)
    assert counts.min() >= len(tri) // 8 - 1 and counts.max() <= len(tri) // 8 + 1
