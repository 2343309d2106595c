"""Regenerates tests/golden/*.npz (mesh fixtures) from the reference's Gmsh files.

Runs only in the build container (needs /root/reference); the GPU box uses the committed .npz.
Each .npz holds exactly what GmshReader semantics extract from the .msh:
compressed vertex coordinates, triangles in file order, boundary segments in file order with their
physical tag.  The Gmsh reader itself is exercised in tests by writing these arrays back to a
temporary .msh (tests/util.py:write_gmsh) and re-reading it, and on the reference's own files: they are copied
verbatim (data fixtures, Gmsh 2.1 and 2.2 ASCII) to tests/golden/msh/ and the product reader must reproduce the .npz.
"""
import os, shutil, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from oracle import binding as ora

REF = "/root/reference/test"
OUT = os.path.join(os.path.dirname(__file__), "..", "tests", "golden")
MESHES = {
    "one_wall": "one_wall_dh/one_wall.msh",
    "sphere": "sphere_pb/sphere.msh",
    "cylinder": "cylinder.msh",
    "pore_small": "pore.msh",
    "pore": "pore_pnp/pore.msh",
}
for name, rel in MESHES.items():
    m = ora.Mesh.read_gmsh(os.path.join(REF, rel))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), x=m.x, y=m.y, tri=m.tri, ba=m.ba, bb=m.bb, bphys=m.bphys)
    os.makedirs(os.path.join(OUT, "msh"), exist_ok=True)
    shutil.copyfile(os.path.join(REF, rel), os.path.join(OUT, "msh", name + ".msh"))
    print(name, m.nv, m.nT, m.nB)
