"""Host-side domain decomposition for the multi-GPU path (one process per GPU).

Stands in for `grid->loadBalance()` (/root/reference/src/pnp_solver_main.cc:108) and PDELab's non-overlapping
ghost bookkeeping (`compute_ghosts`, stationary_pnp.hh:131).  Plumbing only -- numpy on the host, torch.distributed
object collectives for the one-off plan exchange; the data path (assembly, SpMV, Krylov) runs in the CUDA library.

Scheme (DESIGN.md "multi-GPU"):
  * every rank holds the same coarse mesh and computes the same partition of its TRIANGLES by recursive coordinate
    bisection of the centroids;
  * rank p keeps its triangles plus every triangle that touches a vertex of one of its triangles (one ghost layer),
    refines that local mesh uniformly (same midpoint rule as everywhere else, so shared vertices get bit-identical
    coordinates on all ranks) and trims the ghost layer back to one element after each level;
  * a vertex is owned by the lowest rank among the triangles around it; owners are resolved and the halo plan is
    built by matching the (bitwise) coordinates of ghost vertices against owned vertices of the other ranks.
"""
import numpy as np


def rcb_partition(cx, cy, nparts):
    """Recursive coordinate bisection of points (triangle centroids) into `nparts` equal-count parts."""
    part = np.zeros(len(cx), dtype=np.int32)

    def split(idx, lo, n):
        if n == 1:
            part[idx] = lo
            return
        nl = n // 2
        x, y = cx[idx], cy[idx]
        key = x if (x.max() - x.min()) >= (y.max() - y.min()) else y
        order = np.argsort(key, kind="stable")
        k = (len(idx) * nl) // n
        split(idx[order[:k]], lo, nl)
        split(idx[order[k:]], lo + nl, n - nl)

    split(np.arange(len(cx)), 0, nparts)
    return part


def _unique_inverse(keys):
    """Sorted unique values of an int64 array and, for every entry, its rank among them.  torch's multi-threaded sort does
    this several times faster than numpy on the 10^8 edge keys of the finest levels (exact integer work: same result)."""
    try:
        import os
        import torch
        if len(keys) > (1 << 20):
            nthreads = torch.get_num_threads()
            try:  # torchrun pins OMP_NUM_THREADS to 1; this one-off host step may use this rank's share of the cores
                torch.set_num_threads(max(1, (os.cpu_count() or 1) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))))
                u, inv = torch.unique(torch.from_numpy(np.ascontiguousarray(keys)), sorted=True, return_inverse=True)
            finally:
                torch.set_num_threads(nthreads)
            return u.numpy(), inv.numpy()
    except ImportError:
        pass
    return np.unique(keys, return_inverse=True)


def _edge_keys(tri):
    t = tri.astype(np.int64)
    e = np.concatenate([t[:, [0, 1]], t[:, [0, 2]], t[:, [1, 2]]])
    return (e.min(1) << 32) | e.max(1)


class LocalMesh:
    """A rank's piece of the mesh in its own vertex numbering: x, y, tri, tag (owner rank of each triangle), boundary
    segments ba/bb/bphys, carried nodal fields (dict name -> (F, nv) arrays)."""

    def __init__(self, x, y, tri, tag, ba, bb, bphys, fields=None, par=None, gid=None):
        self.x, self.y, self.tri, self.tag = x, y, tri.astype(np.int32), tag.astype(np.int32)
        self.ba, self.bb, self.bphys = ba.astype(np.int32), bb.astype(np.int32), bphys.astype(np.int32)
        self.fields = dict(fields or {})
        # par[v] = (p0, p1): parents of v in the numbering of the LocalMesh this one was refined from (p1 = -1 for a
        # vertex that coincides with its parent); gid[v] = vertex index in the global mesh (coarsest level only)
        self.par = par
        self.gid = gid

    @property
    def nv(self):
        return len(self.x)

    def _compact(self, keep_tri):
        """Keeps the given triangles, the vertices they use, and the boundary segments that are edges of kept triangles."""
        tri = self.tri[keep_tri]
        used = np.zeros(self.nv, dtype=bool)
        used[tri.ravel()] = True
        new_id = np.cumsum(used) - 1
        # a boundary segment stays if both end vertices stay: if one of them is owned, the segment's triangle touches an
        # owned vertex and was kept; segments between two far ghosts only contribute Dirichlet flags in the library
        keep_b = used[self.ba] & used[self.bb]
        fields = {k: v[:, used] for k, v in self.fields.items()}
        return LocalMesh(self.x[used], self.y[used], new_id[tri], self.tag[keep_tri], new_id[self.ba[keep_b]],
                         new_id[self.bb[keep_b]], self.bphys[keep_b], fields,
                         None if self.par is None else self.par[used], None if self.gid is None else self.gid[used])

    def trim(self, me):
        """One ghost layer: my triangles + every triangle touching a vertex of one of my triangles."""
        mine_v = np.zeros(self.nv, dtype=bool)
        mine_v[self.tri[self.tag == me].ravel()] = True
        keep = (self.tag == me) | mine_v[self.tri].any(axis=1)
        return self._compact(keep)

    def refine(self):
        """Uniform red refinement with the library's rule (new vertex = nv + rank of the edge key; children
        (a,ab,ac)(ab,b,bc)(ac,bc,c)(ab,bc,ac)); fields are P1-interpolated; children inherit tag / physical tag."""
        keys, rank = _unique_inverse(_edge_keys(self.tri))  # rank: edge index of (0,1), (0,2), (1,2) of every triangle
        a = (keys >> 32).astype(np.int64); b = (keys & 0xffffffff).astype(np.int64)
        nv = self.nv
        nT = len(self.tri)
        x = np.concatenate([self.x, 0.5 * (self.x[a] + self.x[b])]); y = np.concatenate([self.y, 0.5 * (self.y[a] + self.y[b])])
        fields = {k: np.concatenate([v, 0.5 * (v[:, a] + v[:, b])], axis=1) for k, v in self.fields.items()}

        def mid(p, q):
            p = p.astype(np.int64); q = q.astype(np.int64)
            return (nv + np.searchsorted(keys, (np.minimum(p, q) << 32) | np.maximum(p, q))).astype(np.int32)
        t = self.tri
        ab, ac, bc = ((nv + rank[k * nT:(k + 1) * nT]).astype(np.int32) for k in range(3))
        tri = np.stack([t[:, 0], ab, ac, ab, t[:, 1], bc, ac, bc, t[:, 2], ab, bc, ac], axis=1).reshape(-1, 3)
        m = mid(self.ba, self.bb)
        ba = np.stack([self.ba, m], axis=1).reshape(-1); bb = np.stack([m, self.bb], axis=1).reshape(-1)
        par = np.full((len(x), 2), -1, dtype=np.int64)
        par[:nv, 0] = np.arange(nv)
        par[nv:, 0] = a; par[nv:, 1] = b
        return LocalMesh(x, y, tri, np.repeat(self.tag, 4), ba, bb, np.repeat(self.bphys, 2), fields, par)


def extract_local(a, part, me, fields=None):
    """Rank `me`'s local mesh (with one ghost layer) from the global arrays `a` and the triangle partition `part`."""
    f = {k: np.asarray(v, dtype=np.float64).reshape(-1, len(a["x"])) for k, v in (fields or {}).items()}
    lm = LocalMesh(np.asarray(a["x"], float), np.asarray(a["y"], float), np.asarray(a["tri"]), part, np.asarray(a["ba"]),
                   np.asarray(a["bb"]), np.asarray(a["bphys"]), f, None, np.arange(len(a["x"]), dtype=np.int64))
    return lm.trim(me)


def _coord_keys(x, y):
    k = np.empty((len(x), 2), dtype=np.uint64)
    k[:, 0] = np.ascontiguousarray(x, dtype=np.float64).view(np.uint64)
    k[:, 1] = np.ascontiguousarray(y, dtype=np.float64).view(np.uint64)
    return [r.tobytes() for r in k]


class Plan:
    """Result of finalize(): the local mesh reordered (owned vertices first, ghosts grouped by owner rank) and the
    halo plan in the form pnp_halo_set() takes."""
    pass


def finalize(lm, me, world, all_gather=None):
    """Ownership + halo plan.  `all_gather(obj) -> list over ranks` (torch.distributed.all_gather_object wrapper);
    None for a single rank."""
    nv = lm.nv
    # owner of every vertex that touches one of my triangles: lowest tag around it (its star is complete locally)
    mine_v = np.zeros(nv, dtype=bool)
    mine_v[lm.tri[lm.tag == me].ravel()] = True
    owner = np.full(nv, np.iinfo(np.int32).max, dtype=np.int32)
    for r in np.unique(lm.tag)[::-1]:   # descending: the lowest tag is written last (ufunc.at is two orders slower)
        owner[lm.tri[lm.tag == r].ravel()] = r
    owned = mine_v & (owner == me)
    ghost_ids = np.where(~owned)[0]
    own_ids = np.where(owned)[0]
    p = Plan()
    if world == 1:
        assert len(ghost_ids) == 0
        order = own_ids
        nbr, send_ptr, send_idx, recv_ptr = [], [0], np.zeros(0, np.int32), [0]
    else:
        gk = _coord_keys(lm.x[ghost_ids], lm.y[ghost_ids])
        ghosts_all = all_gather(gk)
        # owned vertices another rank may need: those touching a triangle that is not mine
        foreign_v = np.zeros(nv, dtype=bool)
        foreign_v[lm.tri[lm.tag != me].ravel()] = True
        near_t = foreign_v[lm.tri].any(axis=1)   # triangles within one layer of another rank's triangles
        near_v = np.zeros(nv, dtype=bool); near_v[lm.tri[near_t].ravel()] = True
        cand = np.where(owned & near_v)[0]
        lookup = dict(zip(_coord_keys(lm.x[cand], lm.y[cand]), cand.tolist()))
        claims = []  # claims[q] = (positions in q's ghost list that I own, my local ids in that order)
        for q in range(world):
            if q == me:
                claims.append((np.zeros(0, np.int64), np.zeros(0, np.int64))); continue
            pos, ids = [], []
            for i, k in enumerate(ghosts_all[q]):
                v = lookup.get(k)
                if v is not None:
                    pos.append(i); ids.append(v)
            claims.append((np.array(pos, dtype=np.int64), np.array(ids, dtype=np.int64)))
        claimed_all = all_gather([c[0] for c in claims])  # claimed_all[r][q] = positions of q's ghosts owned by r
        # my ghosts, grouped by owner rank in the order the owner will send them
        ghost_order, recv_ptr, nbr_set = [], [0], []
        seen = np.zeros(len(ghost_ids), dtype=np.int32)
        for r in range(world):
            pos = claimed_all[r][me] if r != me else np.zeros(0, np.int64)
            seen[pos] += 1
            if len(pos) or len(claims[r][0]):
                nbr_set.append(r)
                ghost_order.append(ghost_ids[pos])
                recv_ptr.append(recv_ptr[-1] + len(pos))
        if not np.all(seen == 1):
            raise RuntimeError("halo plan: %d ghost vertices unclaimed, %d claimed twice" % ((seen == 0).sum(), (seen > 1).sum()))
        order = np.concatenate([own_ids] + ghost_order) if ghost_order else own_ids
        new_id = np.empty(nv, dtype=np.int64); new_id[order] = np.arange(nv)
        nbr = nbr_set
        send_ptr, send_idx = [0], []
        for r in nbr:
            send_idx.append(new_id[claims[r][1]])
            send_ptr.append(send_ptr[-1] + len(claims[r][1]))
        send_idx = np.concatenate(send_idx).astype(np.int32) if send_idx else np.zeros(0, np.int32)
    new_id = np.empty(nv, dtype=np.int64); new_id[order] = np.arange(nv)
    p.x, p.y = lm.x[order], lm.y[order]
    p.tri = new_id[lm.tri].astype(np.int32)
    p.ba, p.bb, p.bphys = new_id[lm.ba].astype(np.int32), new_id[lm.bb].astype(np.int32), lm.bphys
    p.tag = lm.tag
    p.fields = {k: v[:, order] for k, v in lm.fields.items()}
    p.n_own, p.nv = len(own_ids), nv
    p.old2new = new_id                                # LocalMesh numbering -> this plan's numbering
    p.par = None if lm.par is None else lm.par[order]  # parents, still in the coarser LocalMesh's numbering
    p.gid = None if lm.gid is None else lm.gid[order]
    p.nbr = np.array(nbr, dtype=np.int32); p.send_ptr = np.array(send_ptr, dtype=np.int32)
    p.send_idx = send_idx; p.recv_ptr = np.array(recv_ptr, dtype=np.int32)
    return p


def build_local(a, nparts, me, levels, fields=None, all_gather=None):
    """Partition the global mesh `a`, refine rank `me`'s part `levels` times, and build its halo plan."""
    tri = np.asarray(a["tri"])
    cx = np.asarray(a["x"])[tri].mean(1); cy = np.asarray(a["y"])[tri].mean(1)
    part = rcb_partition(cx, cy, nparts)
    lm = extract_local(a, part, me, fields)
    for _ in range(levels):
        lm = lm.refine().trim(me)
    return finalize(lm, me, nparts, all_gather)


def build_hierarchy(a, nparts, me, levels, all_gather=None, fields_at=None):
    """Distributed multigrid hierarchy: partitions the global mesh `a` (level 0), and returns the list of per-level plans
    [level 0 (coarsest), ..., level `levels` (finest)].  plan.par (levels >= 1) gives, for every local vertex, its parents
    in the numbering of the next coarser plan (-1: none).  plan.gid (level 0) are global vertex indices.
    fields_at = (level, lookup) injects nodal fields at that level: lookup(x, y) -> dict name -> (F, n) arrays; they
    are P1-interpolated to the finer levels."""
    tri = np.asarray(a["tri"])
    cx = np.asarray(a["x"])[tri].mean(1); cy = np.asarray(a["y"])[tri].mean(1)
    part = rcb_partition(cx, cy, nparts)
    lm = extract_local(a, part, me)
    plans = []
    prev = None
    n_global = len(a["x"])
    for l in range(levels + 1):
        if l > 0:
            lm = lm.refine().trim(me)
        if fields_at is not None and fields_at[0] == l:
            lm.fields = {k: np.asarray(v, dtype=np.float64) for k, v in fields_at[1](lm.x, lm.y).items()}
        p = finalize(lm, me, nparts, all_gather)
        if l > 0:
            par = p.par.copy()
            ok = par >= 0
            par[ok] = prev.old2new[par[ok]]
            p.par = par
        p.n_global = n_global
        plans.append(p)
        prev = p
    return plans


def setup_distributed(capi, root, plans, cfg_path, rank, world, unique_id, aggregates=None, replica_mesh=None,
                      replica_level=0):
    """Creates the library-side hierarchy from build_hierarchy() plans: `root` gets the finest level, one child context
    per coarser level; returns the list of child contexts (keep them alive as long as root).  aggregates = (agg, n_agg) from
    aggregate_greedy() on the global coarsest mesh: the coarsest level is then smoothed and its aggregate system is the
    replicated dense solve (cheap LU); None: the coarsest level itself is solved densely.  replica_mesh = arrays of the
    whole coarsest mesh: the coarsest level is gathered to a replica context that runs the one-GPU multigrid below it.
    replica_level = k > 0: the distributed hierarchy stops at level k; the replica refines the whole coarsest mesh k times
    on its device and continues below with its own geometric levels (no halo exchanges on the small levels)."""
    fine = plans[-1]
    root.params_read(cfg_path)
    root.mesh_set_local(fine.n_own, fine.x, fine.y, fine.tri, fine.ba, fine.bb, fine.bphys)
    root.comm_init(rank, world, unique_id)
    root.halo_set(fine.nbr, fine.send_ptr, fine.send_idx, fine.recv_ptr)
    root.mesh_finalize(True)
    children = []
    lowest = replica_level if replica_mesh is not None else 0
    assert 0 <= lowest <= len(plans) - 2
    for l in range(len(plans) - 2, lowest - 1, -1):
        p = plans[l]
        ch = capi.Context(parent=root)
        ch.mesh_set_local(p.n_own, p.x, p.y, p.tri, p.ba, p.bb, p.bphys)
        ch.halo_set(p.nbr, p.send_ptr, p.send_idx, p.recv_ptr)
        ch.mesh_finalize(True)
        root.mg_push_level(ch, plans[l + 1].par[:, 0], plans[l + 1].par[:, 1])
        children.append(ch)
    if replica_mesh is not None:
        # every rank also holds the WHOLE coarsest mesh: the one-GPU multigrid continues below it (redundantly)
        rep = capi.Context(parent=root)
        rep.mesh_set(**replica_mesh)
        if lowest == 0:
            gid, n_global = plans[0].gid, int(plans[0].n_global)
        else:  # vertices of the refined replica are matched by their (bitwise equal) coordinates
            rep.mesh_refine(lowest)
            g = rep.mesh_get()
            tab = {k: i for i, k in enumerate(_coord_keys(g["x"], g["y"]))}
            gid = np.array([tab[k] for k in _coord_keys(plans[lowest].x, plans[lowest].y)], dtype=np.int32)
            n_global = len(g["x"])
        rep.mesh_finalize(True)
        root.mg_set_coarse_replica(rep, gid, n_global)
        children.append(rep)
    elif aggregates is None:
        root.mg_set_coarse_global(plans[0].gid, int(plans[0].n_global))
    else:
        agg, n_agg = aggregates
        root.mg_set_coarse_aggregates(np.asarray(agg)[plans[0].gid], int(n_agg))
    return children


def aggregate_greedy(nv, tri, leftovers_join=True):
    """Deterministic greedy aggregation of the (global, coarsest) mesh graph: a vertex whose neighbourhood is still free
    becomes the root of an aggregate with all its neighbours; leftovers join an adjacent aggregate."""
    tri = np.asarray(tri, dtype=np.int64)
    e = np.concatenate([tri[:, [0, 1]], tri[:, [1, 2]], tri[:, [0, 2]]])
    e = np.unique(np.concatenate([e, e[:, ::-1]]), axis=0)
    ptr = np.searchsorted(e[:, 0], np.arange(nv + 1))
    nbr = e[:, 1]
    agg = np.full(nv, -1, dtype=np.int64)
    n_agg = 0
    for v in range(nv):
        nb = nbr[ptr[v]:ptr[v + 1]]
        if agg[v] < 0 and np.all(agg[nb] < 0):
            agg[v] = n_agg; agg[nb] = n_agg; n_agg += 1
    if not leftovers_join:  # leftovers first form aggregates among themselves (keeps the aggregates small)
        for v in range(nv):
            if agg[v] < 0:
                nb = nbr[ptr[v]:ptr[v + 1]]
                free = nb[agg[nb] < 0]
                if len(free) >= 2:
                    agg[v] = n_agg; agg[free] = n_agg; n_agg += 1
    size = np.bincount(agg[agg >= 0], minlength=n_agg)
    for v in range(nv):
        if agg[v] < 0:
            nb = nbr[ptr[v]:ptr[v + 1]]
            done = nb[agg[nb] >= 0]
            if len(done):
                best = done[np.argmin(size[agg[done]])]   # join the smallest adjacent aggregate
                agg[v] = agg[best]; size[agg[v]] += 1
            else:
                agg[v] = n_agg; n_agg += 1; size = np.append(size, 1)
    return agg.astype(np.int32), n_agg


# ---------------------------------------------------------------------------------------------------------------
# The same decomposition by the library's native partitioner (csrc/pnp_partition.cu, C ABI pnp_part_*): the plans it
# returns equal build_hierarchy()'s array by array (tests/test_partition_native.py); only the two small all-gathers stay here.
# ---------------------------------------------------------------------------------------------------------------
def build_hierarchy_native(a, nparts, me, levels, all_gather=None, fields_at=None):
    import ctypes as C
    from . import capi
    L = capi.lib()
    L.pnp_part_last_error.restype = C.c_char_p
    dp, ip, lp, up = C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_long), C.POINTER(C.c_ulonglong)

    def ck(st):
        if st != 0:
            raise capi.PnpError(st, L.pnp_part_last_error(h).decode())
    x = np.ascontiguousarray(a["x"], dtype=np.float64); y = np.ascontiguousarray(a["y"], dtype=np.float64)
    tri = np.ascontiguousarray(a["tri"], dtype=np.int32); ba = np.ascontiguousarray(a["ba"], dtype=np.int32)
    bb = np.ascontiguousarray(a["bb"], dtype=np.int32); ph = np.ascontiguousarray(a["bphys"], dtype=np.int32)
    h = C.c_void_p()
    st = L.pnp_part_create(C.c_long(len(x)), x.ctypes.data_as(dp), y.ctypes.data_as(dp), C.c_long(len(tri)), tri.ctypes.data_as(ip),
                           C.c_long(len(ba)), ba.ctypes.data_as(ip), bb.ctypes.data_as(ip), ph.ctypes.data_as(ip), nparts, me, levels,
                           C.byref(h))
    if st != 0:
        raise capi.PnpError(st, "pnp_part_create failed")
    plans = []
    try:
        for l in range(levels + 1):
            n = C.c_long()
            ck(L.pnp_part_ghost_keys(h, l, C.byref(n), None))
            keys = np.zeros(2 * n.value, dtype=np.uint64)
            ck(L.pnp_part_ghost_keys(h, l, C.byref(n), keys.ctypes.data_as(up)))
            if nparts > 1:
                allk = all_gather(keys)
                gptr = np.zeros(nparts + 1, dtype=np.int64); gptr[1:] = np.cumsum([len(k) // 2 for k in allk])
                cat = np.ascontiguousarray(np.concatenate(allk), dtype=np.uint64)
                cptr = np.zeros(nparts + 1, dtype=np.int64)
                ck(L.pnp_part_claim(h, l, gptr.ctypes.data_as(lp), cat.ctypes.data_as(up), cptr.ctypes.data_as(lp), None))
                cpos = np.zeros(max(1, cptr[-1]), dtype=np.int64)
                ck(L.pnp_part_claim(h, l, gptr.ctypes.data_as(lp), cat.ctypes.data_as(up), cptr.ctypes.data_as(lp), cpos.ctypes.data_as(lp)))
                claims_all = all_gather((cptr, cpos))            # every rank's claims about every rank's ghosts
                mine = [claims_all[r][1][claims_all[r][0][me]:claims_all[r][0][me + 1]] for r in range(nparts)]
                mptr = np.zeros(nparts + 1, dtype=np.int64); mptr[1:] = np.cumsum([len(m_) for m_ in mine])
                mpos = np.ascontiguousarray(np.concatenate(mine) if mptr[-1] else np.zeros(1), dtype=np.int64)
                ck(L.pnp_part_finalize(h, l, mptr.ctypes.data_as(lp), mpos.ctypes.data_as(lp)))
            else:
                ck(L.pnp_part_finalize(h, l, None, None))
            sz = (C.c_long * 8)()
            ck(L.pnp_part_sizes(h, l, sz))
            nv, n_own, nT, nB, n_nbr, n_send, n_global, has_gid = list(sz)
            p = Plan()
            p.x, p.y = np.zeros(nv), np.zeros(nv)
            p.tri = np.zeros((nT, 3), dtype=np.int32)
            p.ba, p.bb, p.bphys = (np.zeros(nB, dtype=np.int32) for _ in range(3))
            p.nbr = np.zeros(n_nbr, dtype=np.int32); p.send_ptr = np.zeros(n_nbr + 1, dtype=np.int32)
            p.send_idx = np.zeros(n_send, dtype=np.int32); p.recv_ptr = np.zeros(n_nbr + 1, dtype=np.int32)
            par0, par1 = np.zeros(nv, dtype=np.int32), np.zeros(nv, dtype=np.int32)
            gid = np.zeros(nv if has_gid else 1, dtype=np.int32)
            ck(L.pnp_part_get(h, l, p.x.ctypes.data_as(dp), p.y.ctypes.data_as(dp), p.tri.ctypes.data_as(ip), p.ba.ctypes.data_as(ip),
                              p.bb.ctypes.data_as(ip), p.bphys.ctypes.data_as(ip), p.nbr.ctypes.data_as(ip), p.send_ptr.ctypes.data_as(ip),
                              p.send_idx.ctypes.data_as(ip), p.recv_ptr.ctypes.data_as(ip), par0.ctypes.data_as(ip), par1.ctypes.data_as(ip),
                              gid.ctypes.data_as(ip) if has_gid else None))
            p.n_own, p.nv, p.n_global = n_own, nv, n_global
            p.par = None if l == 0 else np.stack([par0, par1], axis=1).astype(np.int64)
            p.gid = gid.astype(np.int64) if has_gid else None
            # nodal fields: injected at one level, P1-interpolated to the finer ones
            p.fields = {}
            if fields_at is not None:
                if fields_at[0] == l:
                    p.fields = {k: np.asarray(v, dtype=np.float64) for k, v in fields_at[1](p.x, p.y).items()}
                elif fields_at[0] < l:
                    two = par1 >= 0
                    for k, v in plans[-1].fields.items():
                        f = v[:, par0].copy()
                        f[:, two] = 0.5 * (v[:, par0[two]] + v[:, par1[two]])
                        p.fields[k] = f
            plans.append(p)
    finally:
        L.pnp_part_destroy(h)
    return plans
