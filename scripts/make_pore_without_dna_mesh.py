"""Mesher-free triangulation of /root/reference/test/pore_without_dna/pore_without_dna.geo (BASELINE config C4).

The reference ships only the .geo of this case (pore.cfg:21 names pore_without_dna.msh, which is not in the repository) and
gmsh is not available, so the mesh is generated here, deterministically:

  * geometry and characteristic lengths exactly as in the .geo (:1-16): a 100 x 55 half-plane box (axis r = 0), a
    membrane of length 20 with a pore of radius 10 whose two corners are rounded with radius 1; lengths 2 at the pore
    and axis points, 6 at the box corners, interpolated linearly along every curve (as gmsh does);
  * boundary curves are split into segments of the local length, in the order and orientation of the .geo's Line Loop
    (:66), physical line tags 0..5 as in the .geo (:69-74): 0 pore wall {111, 3, 2, 4, 1}, 1 axis {11, 7, 8}, 2 inflow
    {10}, 3 outflow {13}, 4 top left {9}, 5 top right {12};
  * interior vertices: dart throwing (numpy RandomState(20111), candidates on a jittered fine lattice in scan order) against
    the size field h(x) = harmonic-type interpolation of the boundary sizes (inverse-distance weights to the boundary
    vertices), minimum distance 0.8 h to accepted points and 0.7 h to the boundary;
  * scipy.spatial.Delaunay of all vertices, triangles whose centroid lies outside the polygon dropped, three sweeps of
    Laplacian smoothing of the interior vertices with re-triangulation; every boundary segment must be an edge of exactly
    one kept triangle (asserted); triangles are stored counter-clockwise.

Output: tests/golden/msh/pore_without_dna.msh (Gmsh 2.2 ASCII, the format GmshReader reads, pnp_solver_main.cc:82-91),
tests/golden/pore_without_dna.npz (the arrays the readers extract from it) and tests/golden/pore_without_dna.cfg
(the reference's pore.cfg of this case, keys normalised like the other fixtures).
"""
import os
import sys

import numpy as np
from scipy.spatial import Delaunay, cKDTree

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))

# ---- pore_without_dna.geo:1-16 ----
pore_ref, box_ref = 2.0, 6.0
pore_radius = pore_radius_left = 10.0
pore_length = dna_length = 20.0
box_z, box_r, rs = 100.0, 55.0, 1.0
P = {0: (-pore_length / 2, box_r, pore_ref), 1: (-pore_length / 2, pore_radius + rs, pore_ref),
     2: (-pore_length / 2 + rs, pore_radius, pore_ref), 3: (-pore_length / 2 + rs, pore_radius + rs, pore_ref),
     4: (pore_length / 2, box_r, pore_ref), 5: (pore_length / 2, pore_radius_left + rs, pore_ref),
     6: (pore_length / 2 - rs, pore_radius_left, pore_ref), 7: (pore_length / 2 - rs, pore_radius_left + rs, pore_ref),
     8: (-dna_length / 2, 0.0, pore_ref), 11: (dna_length / 2, 0.0, pore_ref),
     14: (-box_z / 2, 0.0, box_ref), 15: (-box_z / 2, box_r, box_ref), 16: (box_z / 2, 0.0, box_ref), 17: (box_z / 2, box_r, box_ref)}
LINES = {111: (0, 1), 1: (4, 5), 2: (2, 6), 7: (8, 11), 8: (8, 14), 9: (0, 15), 10: (14, 15), 11: (11, 16), 12: (4, 17), 13: (16, 17)}
CIRCLES = {3: (1, 3, 2), 4: (5, 7, 6)}   # start, centre, end
LOOP = [1, 4, -2, -3, -111, 9, -10, -8, 7, 11, 13, -12]
PHYS = {0: [111, 3, 2, 4, 1], 1: [11, 7, 8], 2: [10], 3: [13], 4: [9], 5: [12]}
TAG = {c: t for t, cs in PHYS.items() for c in cs}


def curve_points(cid):
    """Vertices along curve |cid| in loop direction, end point excluded; sizes interpolated linearly in arc length."""
    rev = cid < 0
    c = abs(cid)
    if c in LINES:
        a, b = LINES[c]
        if rev:
            a, b = b, a
        pa, pb = np.array(P[a][:2]), np.array(P[b][:2])
        ha, hb = P[a][2], P[b][2]
        length = np.linalg.norm(pb - pa)
        par = lambda s: pa + (pb - pa) * s  # noqa: E731
    else:
        a, m, b = CIRCLES[c]
        if rev:
            a, b = b, a
        ctr = np.array(P[m][:2])
        va, vb = np.array(P[a][:2]) - ctr, np.array(P[b][:2]) - ctr
        t0, t1 = np.arctan2(va[1], va[0]), np.arctan2(vb[1], vb[0])
        d = (t1 - t0 + np.pi) % (2 * np.pi) - np.pi   # the short way round (quarter circles)
        r = np.linalg.norm(va)
        ha, hb = P[a][2], P[b][2]
        length = abs(d) * r
        par = lambda s: ctr + r * np.array([np.cos(t0 + d * s), np.sin(t0 + d * s)])  # noqa: E731
    # number of segments: integral of ds / h(s) with linear h, at least 2 on the quarter circles
    n = max(int(round(length * (np.log(hb / ha) / (hb - ha) if hb != ha else 1.0 / ha))), 2 if c in CIRCLES else 1)
    # equidistribute: cumulative of 1/h
    s = np.linspace(0, 1, 2001)
    w = np.cumsum(np.concatenate([[0], 0.5 * (1 / (ha + (hb - ha) * s[1:]) + 1 / (ha + (hb - ha) * s[:-1])) * np.diff(s)]))
    tk = np.interp(np.linspace(0, w[-1], n + 1), w, s)
    pts = np.array([par(t) for t in tk[:-1]])
    hs = ha + (hb - ha) * tk[:-1]
    return pts, hs, TAG[c]


def inside(poly, q):
    """Even-odd rule for points q (n, 2) against the closed polygon poly (m, 2)."""
    x, y = q[:, 0], q[:, 1]
    inside_ = np.zeros(len(q), dtype=bool)
    for i in range(len(poly)):
        x0, y0 = poly[i]; x1, y1 = poly[(i + 1) % len(poly)]
        hit = ((y0 > y) != (y1 > y)) & (x < (x1 - x0) * (y - y0) / (y1 - y0 + 1e-300) + x0)
        inside_ ^= hit
    return inside_


def generate():
    """Returns the mesh arrays (x, y, tri, ba, bb, bphys) -- deterministic."""
    bpts, bh, btag = [], [], []
    for cid in LOOP:
        pts, hs, tag = curve_points(cid)
        bpts.append(pts); bh.append(hs); btag += [tag] * len(pts)
    bpts = np.concatenate(bpts); bh = np.concatenate(bh); btag = np.array(btag)
    nb = len(bpts)
    area2 = np.sum(bpts[:, 0] * np.roll(bpts[:, 1], -1) - np.roll(bpts[:, 0], -1) * bpts[:, 1])
    assert area2 != 0
    btree = cKDTree(bpts)

    def hfield(q):
        d, i = btree.query(q, k=8)
        w = 1.0 / (d + 1e-9) ** 2
        hb = (w * bh[i]).sum(1) / w.sum(1)
        # sizes grow away from the boundary at most like 1 + 0.25 * distance (gmsh-like gradation), capped by the box size
        return np.minimum(np.maximum(hb, 0.0), box_ref) * 1.0 + 0.0 * d[:, 0]

    rng = np.random.RandomState(20111)
    step = 0.5
    gx, gy = np.meshgrid(np.arange(-box_z / 2 + step / 2, box_z / 2, step), np.arange(step / 2, box_r, step))
    cand = np.stack([gx.ravel(), gy.ravel()], axis=1) + rng.uniform(-0.2, 0.2, (gx.size, 2))
    cand = cand[inside(bpts, cand)]
    hc = hfield(cand)
    db, _ = btree.query(cand)
    cand, hc = cand[db >= 0.7 * hc], hc[db >= 0.7 * hc]
    order = rng.permutation(len(cand))
    acc = []
    cell = 2.0
    grid = {}
    for i in order:
        p, h = cand[i], hc[i]
        kx, ky = int(np.floor(p[0] / cell)), int(np.floor(p[1] / cell))
        r = int(np.ceil(0.8 * h / cell))
        ok = True
        for ax in range(kx - r, kx + r + 1):
            for ay in range(ky - r, ky + r + 1):
                for j in grid.get((ax, ay), ()):
                    if (acc[j][0] - p[0]) ** 2 + (acc[j][1] - p[1]) ** 2 < (0.8 * min(h, acc[j][2])) ** 2:
                        ok = False; break
                if not ok:
                    break
            if not ok:
                break
        if ok:
            grid.setdefault((kx, ky), []).append(len(acc)); acc.append((p[0], p[1], h))
    ipts = np.array([(a[0], a[1]) for a in acc])

    def triangulate(ip):
        pts = np.concatenate([bpts, ip])
        tri = Delaunay(pts).simplices
        cen = pts[tri].mean(1)
        tri = tri[inside(bpts, cen)]
        a, b, c = pts[tri[:, 0]], pts[tri[:, 1]], pts[tri[:, 2]]
        det = (b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (c[:, 0] - a[:, 0]) * (b[:, 1] - a[:, 1])
        tri = tri[np.abs(det) > 1e-12]
        det = det[np.abs(det) > 1e-12]
        tri[det < 0] = tri[det < 0][:, [0, 2, 1]]
        return pts, tri

    for _ in range(3):  # Laplacian smoothing of interior vertices
        pts, tri = triangulate(ipts)
        nvt = len(pts)
        s = np.zeros((nvt, 2)); cnt = np.zeros(nvt)
        for i, j in ((0, 1), (1, 2), (2, 0)):
            np.add.at(s, tri[:, i], pts[tri[:, j]]); np.add.at(cnt, tri[:, i], 1)
            np.add.at(s, tri[:, j], pts[tri[:, i]]); np.add.at(cnt, tri[:, j], 1)
        new = s / np.maximum(cnt, 1)[:, None]
        ipts = 0.5 * ipts + 0.5 * new[nb:]
    pts, tri = triangulate(ipts)
    # every boundary segment is an edge of exactly one triangle, and no other edge is a boundary edge
    e = np.concatenate([tri[:, [0, 1]], tri[:, [1, 2]], tri[:, [2, 0]]])
    key = np.minimum(e[:, 0], e[:, 1]).astype(np.int64) * len(pts) + np.maximum(e[:, 0], e[:, 1])
    uk, cnt = np.unique(key, return_counts=True)
    bnd = set(uk[cnt == 1].tolist())
    ba = np.arange(nb); bb = (np.arange(nb) + 1) % nb
    want = set((np.minimum(ba, bb).astype(np.int64) * len(pts) + np.maximum(ba, bb)).tolist())
    assert bnd == want, "triangulation does not conform to the boundary: %d missing, %d extra" % (len(want - bnd), len(bnd - want))
    used = np.zeros(len(pts), dtype=bool); used[tri.ravel()] = True
    assert used.all()
    # minimum angle report
    a, b, c = pts[tri[:, 0]], pts[tri[:, 1]], pts[tri[:, 2]]

    def ang(u, v):
        return np.degrees(np.arccos(np.clip((u * v).sum(1) / np.linalg.norm(u, axis=1) / np.linalg.norm(v, axis=1), -1, 1)))
    amin = np.minimum(np.minimum(ang(b - a, c - a), ang(a - b, c - b)), ang(a - c, b - c))
    print("pore_without_dna: %d vertices, %d triangles, %d boundary segments, minimum angle %.1f deg" % (len(pts), len(tri), nb, amin.min()))
    assert amin.min() > 15.0
    return dict(x=pts[:, 0].copy(), y=pts[:, 1].copy(), tri=tri.astype(np.int32), ba=ba.astype(np.int32), bb=bb.astype(np.int32),
                bphys=btag.astype(np.int32))


def main():
    import util
    arr = generate()
    out = os.path.join(util.GOLDEN, "msh", "pore_without_dna.msh")
    util.write_gmsh(out, arr)
    # the fixture arrays are what the reader extracts from the FILE (coordinates round-trip through %.17g exactly)
    from oracle import binding as ora
    m = ora.Mesh.read_gmsh(out)
    np.savez_compressed(os.path.join(util.GOLDEN, "pore_without_dna.npz"), x=m.x, y=m.y, tri=m.tri, ba=m.ba, bb=m.bb, bphys=m.bphys)
    assert np.array_equal(m.x, arr["x"]) and np.array_equal(m.tri, arr["tri"]) and np.array_equal(m.bphys, arr["bphys"])
    # config: the reference's own pore.cfg of this case, keys normalised like the other fixtures
    ref = "/root/reference/test/pore_without_dna/pore.cfg"
    if os.path.exists(ref):
        keep = []
        for line in open(ref):
            t = line.strip()
            if not t or t.startswith("#"):
                continue
            keep.append(t)
        with open(os.path.join(util.GOLDEN, "pore_without_dna.cfg"), "w") as f:
            f.write("# test config for the pore_without_dna case (test/pore_without_dna/pore.cfg of the reference, comments dropped)\n")
            f.write("\n".join(keep) + "\n")


if __name__ == "__main__":
    main()
