"""Shared helpers for the tests: fixture loading and a Gmsh 2.2 ASCII writer."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MESHES = ["one_wall", "sphere", "cylinder", "pore_small", "pore", "pore_without_dna"]
# mesh fixture -> config that goes with it
CASES = {"one_wall": "one_wall", "sphere": "sphere", "cylinder": "cylinder", "pore_small": "pore", "pore": "pore",
         "pore_without_dna": "pore_without_dna"}  # (generated mesh: scripts/make_pore_without_dna_mesh.py)


def load_mesh_arrays(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in ("x", "y", "tri", "ba", "bb", "bphys")}


def cfg_path(name):
    return os.path.join(GOLDEN, CASES.get(name, name) + ".cfg")


def write_gmsh(path, a, node_id_offset=1, shuffle_nodes=False, extra_nodes=0):
    """Writes mesh arrays as Gmsh 2.2 ASCII: line elements first (file order = boundary segment order),
    then triangles.  Node ids are 1-based; `extra_nodes` appends unused nodes (the reader must skip them)."""
    nv = len(a["x"])
    ids = np.arange(nv) + node_id_offset
    with open(path, "w") as f:
        f.write("$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n%d\n" % (nv + extra_nodes))
        order = np.arange(nv)
        if shuffle_nodes:
            order = np.random.RandomState(0).permutation(nv)
        for i in order:
            f.write("%d %.17g %.17g 0\n" % (ids[i], a["x"][i], a["y"][i]))
        for k in range(extra_nodes):
            f.write("%d 99 99 0\n" % (nv + node_id_offset + k))
        f.write("$EndNodes\n$Elements\n%d\n" % (len(a["ba"]) + len(a["tri"])))
        eid = 1
        for s in range(len(a["ba"])):
            f.write("%d 1 2 %d %d %d %d\n" % (eid, a["bphys"][s], 100 + a["bphys"][s], ids[a["ba"][s]], ids[a["bb"][s]]))
            eid += 1
        for t in a["tri"]:
            f.write("%d 2 2 7 8 %d %d %d\n" % (eid, ids[t[0]], ids[t[1]], ids[t[2]]))
            eid += 1
        f.write("$EndElements\n")
