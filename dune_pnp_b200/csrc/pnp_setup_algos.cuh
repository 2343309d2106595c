// pnp_setup_algos.cuh -- per-item logic of the mesh pipeline as __host__ __device__ functions:
// edge keys and red-refinement children, corner ("half-edge") records, fan ordering of a vertex
// star, boundary-face reconstruction.  pnp_setup.cu wraps them in kernels (cub sorts in between);
// tests/host_harness/ runs the very same functions in host loops (std::sort in between) so the
// logic can be checked against the CPU oracle on a machine without a GPU.
#pragma once
#include <stdint.h>

#include "pnp_elem.cuh"

namespace pnp {

// ring-slot encoding (see pnp_star.cuh)
constexpr unsigned STAR_VMASK = 0x07FFFFFFu;
constexpr unsigned STAR_HAS_TRI = 1u << 27;
constexpr int STAR_LI_SHIFT = 28;
constexpr unsigned STAR_CW = 1u << 30;
constexpr long STAR_MAX_VERTICES = 1l << 27;

PNP_HD uint64_t edge_key(int a, int b) {
  const unsigned lo = (unsigned)(a < b ? a : b), hi = (unsigned)(a < b ? b : a);
  return ((uint64_t)lo << 32) | (uint64_t)hi;
}
PNP_HD long lower_bound_u64(const uint64_t* a, long n, uint64_t key) {
  long lo = 0, hi = n;
  while (lo < hi) {
    const long mid = (lo + hi) >> 1;
    if (a[mid] < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}
// DUNE reference triangle: face f joins local vertices face_v(f,0) < face_v(f,1)
PNP_HD int face_v(int f, int l) { return l == 0 ? (f == 2 ? 1 : 0) : (f == 0 ? 1 : 2); }

// key of edge i (= 3*t + f) of the triangle list
PNP_HD uint64_t tri_edge_key(const int* tri, long i) {
  const long t = i / 3; const int f = (int)(i % 3);
  return edge_key(tri[3 * t + face_v(f, 0)], tri[3 * t + face_v(f, 1)]);
}
// Uniform red refinement (DESIGN.md "refinement rule"): the midpoint of the edge with rank k among the
// sorted unique edge keys becomes vertex nv + k; children of (a,b,c): (a,ab,ac)(ab,b,bc)(ac,bc,c)(ab,bc,ac)
PNP_HD void refine_children(const int* tri, long t, const uint64_t* ukeys, long nE, long nv, int* out12) {
  const int a = tri[3 * t], b = tri[3 * t + 1], c = tri[3 * t + 2];
  const int ab = (int)(nv + lower_bound_u64(ukeys, nE, edge_key(a, b)));
  const int ac = (int)(nv + lower_bound_u64(ukeys, nE, edge_key(a, c)));
  const int bc = (int)(nv + lower_bound_u64(ukeys, nE, edge_key(b, c)));
  out12[0] = a;  out12[1] = ab;  out12[2] = ac;
  out12[3] = ab; out12[4] = b;   out12[5] = bc;
  out12[6] = ac; out12[7] = bc;  out12[8] = c;
  out12[9] = ab; out12[10] = bc; out12[11] = ac;
}

// Corner records of triangle t: key = (v << 32 | from), payload = to | flags, where going counter-
// clockwise around v the triangle spans from neighbour `from` to neighbour `to`.
// Returns false for a degenerate triangle.
PNP_HD bool corner_records(const int* tri, long t, const int* ext2int, const double* x, const double* y, uint64_t* keys3,
                           unsigned* pay3) {
  const int e0 = tri[3 * t], e1 = tri[3 * t + 1], e2 = tri[3 * t + 2];
  const double det = (x[e1] - x[e0]) * (y[e2] - y[e0]) - (x[e2] - x[e0]) * (y[e1] - y[e0]);
  const bool cw = det < 0.0;
  const int v[3] = {ext2int[e0], ext2int[e1], ext2int[e2]};
  for (int li = 0; li < 3; li++) {
    const int n = v[(li + 1) % 3], p = v[(li + 2) % 3];
    const int from = cw ? p : n, to = cw ? n : p;
    keys3[li] = ((uint64_t)(unsigned)v[li] << 32) | (unsigned)from;
    pay3[li] = (unsigned)to | STAR_HAS_TRI | ((unsigned)li << STAR_LI_SHIFT) | (cw ? STAR_CW : 0u);
  }
  return !(det == 0.0 || e0 == e1 || e1 == e2 || e0 == e2);
}

// Records [b,e) (sorted by key) belong to one vertex.  Returns the index of the record that starts its
// fan: the one whose `from` is nobody's `to` (open fan), else the first record (closed fan).
// *open = 1 for an open fan; *bad = 1 if several fans meet at the vertex or it has no triangle.
PNP_HD int fan_start(const uint64_t* keys, const unsigned* pay, int b, int e, int* open, int* bad) {
  int starts = 0, first = b;
  for (int i = b; i < e; i++) {
    const unsigned from = (unsigned)(keys[i] & 0xffffffffu);
    bool has_pred = false;
    for (int j = b; j < e; j++) if ((pay[j] & STAR_VMASK) == from) { has_pred = true; break; }
    if (!has_pred) { if (starts == 0) first = i; starts++; }
  }
  *bad = (starts > 1 || e == b) ? 1 : 0;
  *open = starts > 0 ? 1 : 0;
  return first;
}
// Writes row v of the star (diagonal slot + ring in counter-clockwise order). Returns false if the
// records do not chain into a single fan.
PNP_HD bool ring_fill(const uint64_t* keys, const unsigned* pay, int b, int e, int v, unsigned* row) {
  int open, bad;
  int i = fan_start(keys, pay, b, e, &open, &bad);
  bool ok = !bad;
  int s = 0;
  row[s++] = (unsigned)v;
  unsigned to = 0;
  for (int k = 0; k < e - b; k++) {
    const unsigned from = (unsigned)(keys[i] & 0xffffffffu);
    row[s++] = from | (pay[i] & ~STAR_VMASK);
    to = pay[i] & STAR_VMASK;
    if (k + 1 < e - b) { // successor: the record whose `from` is this record's `to`
      int nx = -1;
      for (int j = b; j < e; j++) if ((unsigned)(keys[j] & 0xffffffffu) == to) { nx = j; break; }
      if (nx < 0) { ok = false; break; }
      i = nx;
    }
  }
  if (open) row[s++] = to;
  else if (e > b && to != (row[1] & STAR_VMASK)) ok = false;
  return ok;
}

// One boundary face in element terms (what alpha_boundary sees).
struct BFace {
  int v[3];   // internal vertex ids in ELEMENT-LOCAL order
  int f;      // DUNE face index 0:(0,1) 1:(0,2) 2:(1,2)
  int phys;   // physical tag = index of [surface_i]
  int seg;    // boundarySegmentIndex (file order of the line element)
  int a, b;   // the segment's end vertices (internal ids), also valid when no owned vertex touches the face
};
// Boundary segment (a,b) (internal ids): finds the one element that owns the edge from a's open fan.
// Returns false if (a,b) is not a boundary edge of the mesh.
PNP_HD bool boundary_face_of(const int* rp, const unsigned* adj, int a, int b, BFace* bf) {
  const int s0 = rp[a] + 1, s1 = rp[a + 1];
  bf->v[0] = bf->v[1] = bf->v[2] = -1; bf->f = 0; bf->a = a; bf->b = b;
  if (s1 - s0 < 2 || (adj[s1 - 1] & STAR_HAS_TRI)) return false; // a is not a boundary vertex
  int st;
  if ((int)(adj[s0] & STAR_VMASK) == b) st = s0;                // first edge of the open fan
  else if ((int)(adj[s1 - 1] & STAR_VMASK) == b) st = s1 - 2;   // last edge
  else return false;
  const unsigned fl = adj[st];
  const int cur = (int)(fl & STAR_VMASK), nxt = (int)(adj[st + 1] & STAR_VMASK);
  const int li = (fl >> STAR_LI_SHIFT) & 3;
  const bool cw = fl & STAR_CW;
  const int n = cw ? nxt : cur, p = cw ? cur : nxt;
  bf->v[li] = a; bf->v[(li + 1) % 3] = n; bf->v[(li + 2) % 3] = p;
  const int lb = bf->v[0] == b ? 0 : (bf->v[1] == b ? 1 : 2);
  bf->f = li + lb - 1;
  return true;
}

} // namespace pnp
