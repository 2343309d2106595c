// pnp_p2.cu -- quadratic and cubic elements (the reference's -DPDEGREE=2 / 3 builds, /root/reference/src/Makefile.am:54-110;
// instationary_pnp_from_pb_md.hh:26-28,125): the grid function space with edge (and element) dofs, its constraints and BCRS
// pattern, residual and Jacobian assembly of the five local operators with 6 / 10 local dofs per field, a general CSR SpMV.
//
// The P1 path lives on the vertex-star layout, which is both mesh and pattern for linear elements only; the P2 couplings
// (all 6 dofs of an element with each other) do not fit it, so this path uses the reference's own container layout on the
// device: dofs numbered codim by codim -- element bubbles (degree 3), then DEG-1 per edge (counted from the end vertex with
// the smaller index), then vertices -- fields lexicographic, matrices in
// scalar CSR with ascending columns (ISTLBCRSMatrixBackend<1,1>) -- what pnp_pattern_get returns IS the device layout.
// Assembly in two phases, deterministic and in the reference's summation order: (1) one thread per element computes the
// element vector / matrix (alpha_volume + alpha_boundary; NumericalJacobianVolume or the exact derivative) into a scratch
// block, (2) one thread per dof row sums its elements' contributions in ascending element order (the order of PDELab's
// element loop) into the residual entry / the row's CSR slots.  No atomics; constrained rows are trivial, constrained
// residual entries zero.  Compiled -fmad=false (Makefile): the FD Jacobian rounds like the CPU restatement.
// Set-up (edge numbering, incidence lists, patterns, Dirichlet flags) runs on the host from the canonical mesh arrays:
// quadratic elements are a functionality row here, the throughput path is the P1 star layout.
#include <algorithm>

#include "pnp_common.cuh"
#include "pnp_elem_p2.cuh"

namespace pnp {


struct P2Pattern {
  long nnz = 0;
  std::vector<int> h_rp, h_col;
  DBuf<int> rp, col;
  // level schedule of the row-order sweeps (SeqSSOR, SeqILU0), built on first use: rows grouped by level, diagonal positions
  int nlev = 0;
  std::vector<int> lev_ptr;
  DBuf<int> order, diag;
};

struct P2Space {
  int deg = 2, NL = 6;
  long nE = 0, nd = 0, nT = 0, nv = 0, eoff = 0, voff = 0; // first edge dof, first vertex dof
  std::vector<int> h_eva, h_evb, h_e2d, h_fphys, h_nbr; // edges; element -> NL scalar dofs; element face -> surface (-1 interior); face neighbours
  std::vector<int> h_edge_phys;                           // per edge: surface of a boundary edge, -1 inside
  std::vector<unsigned char> h_dir;                       // per scalar dof: bit c = Dirichlet for BC component c
  DBuf<int> e2d, fphys, inc_ptr, inc;                     // incidence: scalar dof -> (element * 16 + local node), elements ascending
  std::vector<int> h_bfaces; DBuf<int> bfaces;            // boundary faces, element * 4 + face, in element / intersection order
  DBuf<unsigned char> dir;
  std::map<int, std::unique_ptr<P2Pattern>> patterns;     // key = 4 * F + comp0
  DBuf<double> scratch;
};

namespace {

const int FACE_V2[3][2] = {{0, 1}, {0, 2}, {1, 2}};

template <int DEG, int OP>
__device__ __forceinline__ void p2_gather_element(int e, const int* tri, const double* cx, const double* cy, const int* e2d, long nd,
                                                  const double* u, const double* aux0, const double* aux1,
                                                  typename PkElem<DEG>::Geo2& G, double* xl, double (*caux)[PkElem<DEG>::NL]) {
  constexpr int F = OpTraits<OP>::F, NL = PkElem<DEG>::NL;
  const int a = tri[3 * e], b = tri[3 * e + 1], c = tri[3 * e + 2];
  G = PkElem<DEG>::make_geo2(cx[a], cy[a], cx[b], cy[b], cx[c], cy[c]);
  for (int i = 0; i < NL; i++) {
    const int d = e2d[NL * e + i];
    for (int k = 0; k < F; k++) xl[NL * k + i] = u[(long)k * nd + d];
    caux[0][i] = OpTraits<OP>::NAUX >= 1 ? aux0[d] : 0.0;
    caux[1][i] = OpTraits<OP>::NAUX >= 2 ? aux1[d] : 0.0;
  }
}

// phase 1, residual: element vector (alpha_volume + alpha_boundary in intersection order 0, 2, 1) -> scratch[e * n ..]
template <int DEG, int OP, int NQ>
__global__ void k_p2_elem_residual(int nT, const int* __restrict__ tri, const double* __restrict__ cx, const double* __restrict__ cy,
                                   const int* __restrict__ e2d, const int* __restrict__ fphys, const double* __restrict__ surf_flux,
                                   const unsigned char* __restrict__ surf_dir, PhysParams P, long nd, int comp0, int line_pts,
                                   const double* __restrict__ u, const double* __restrict__ aux0, const double* __restrict__ aux1,
                                   double* __restrict__ out) {
  using E = PkElem<DEG>;
  constexpr int F = OpTraits<OP>::F, NL = E::NL, n = NL * F;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nT; e += gridDim.x * blockDim.x) {
    typename E::Geo2 G; double xl[n], caux[2][NL], rl[n];
    p2_gather_element<DEG, OP>(e, tri, cx, cy, e2d, nd, u, aux0, aux1, G, xl, caux);
    for (int i = 0; i < n; i++) rl[i] = 0.0;
    E::template alpha_volume<OP, NQ>(G, P, xl, caux, rl);
    if (OP == OP_PB || OP == OP_POISSON || OP == OP_PNP) {
      const int order[3] = {0, 2, 1};
      for (int fi = 0; fi < 3; fi++) {
        const int f = order[fi], ph = fphys[3 * e + f];
        if (ph < 0) continue;
        const int va = tri[3 * e + (f == 2 ? 1 : 0)], vb = tri[3 * e + (f == 0 ? 1 : 2)];
        double j[3]; bool skip[3];
        for (int k = 0; k < F; k++) {
          const int comp = F == 3 ? k : comp0;
          j[k] = surf_flux[3 * ph + comp]; skip[k] = (surf_dir[ph] >> comp) & 1;
        }
        E::alpha_boundary(f, cx[va], cy[va], cx[vb], cy[vb], F, j, skip, P, line_pts, rl);
      }
    }
    for (int i = 0; i < n; i++) out[(long)e * n + i] = rl[i];
  }
}
// phase 2, residual: r[(k, d)] = sum over d's elements, ascending; constrained entries zero
__global__ void k_p2_gather_residual(long nd, int F, int NL, int comp0, const int* __restrict__ inc_ptr, const int* __restrict__ inc,
                                     const unsigned char* __restrict__ dir, const double* __restrict__ scratch, double* __restrict__ r) {
  const int n = NL * F;
  for (long g = blockIdx.x * (long)blockDim.x + threadIdx.x; g < (long)F * nd; g += (long)gridDim.x * blockDim.x) {
    const int k = (int)(g / nd); const long d = g - (long)k * nd;
    double sum = 0.0;
    for (int t = inc_ptr[d]; t < inc_ptr[d + 1]; t++) sum += scratch[(long)(inc[t] >> 4) * n + NL * k + (inc[t] & 15)];
    r[g] = ((dir[d] >> (F == 3 ? k : comp0)) & 1u) ? 0.0 : sum;
  }
}
// phase 1, Jacobian: element matrix -> scratch[e * n * n ..]
template <int DEG, int OP, int NQ, int MODE>
__global__ void k_p2_elem_jacobian(int e0, int ne, const int* __restrict__ tri, const double* __restrict__ cx, const double* __restrict__ cy,
                                   const int* __restrict__ e2d, PhysParams P, long nd, double eps, const double* __restrict__ u,
                                   const double* __restrict__ aux0, const double* __restrict__ aux1, double* __restrict__ out) {
  using E = PkElem<DEG>;
  constexpr int F = OpTraits<OP>::F, NL = E::NL, n = NL * F;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < ne; t += gridDim.x * blockDim.x) {
    const int e = e0 + t;
    typename E::Geo2 G; double xl[n], caux[2][NL];
    p2_gather_element<DEG, OP>(e, tri, cx, cy, e2d, nd, u, aux0, aux1, G, xl, caux);
    double* Ae = out + (long)t * n * n; // accumulated in place (one writer)
    for (int i = 0; i < n * n; i++) Ae[i] = 0.0;
    if (MODE == 0) E::template jacobian_fd<OP, NQ>(G, P, xl, caux, eps, Ae);
    else E::template jacobian_exact<OP, NQ>(G, P, xl, caux, Ae);
  }
}
// phase 2, Jacobian: row (ki, d) of the CSR matrix receives the contributions of the elements [e0, e1) (the scratch block
// holds that chunk); chunks come in ascending order, the first one (e0 == 0) initialises the row
__global__ void k_p2_gather_jacobian(int e0, int e1, long nd, int F, int NL, int comp0, const int* __restrict__ inc_ptr, const int* __restrict__ inc,
                                     const int* __restrict__ e2d, const unsigned char* __restrict__ dir, const int* __restrict__ rp,
                                     const int* __restrict__ col, const double* __restrict__ scratch, double* __restrict__ vals) {
  const int n = NL * F;
  for (long gi = blockIdx.x * (long)blockDim.x + threadIdx.x; gi < (long)F * nd; gi += (long)gridDim.x * blockDim.x) {
    const int ki = (int)(gi / nd); const long d = gi - (long)ki * nd;
    const int r0 = rp[gi], r1 = rp[gi + 1];
    if (e0 == 0) for (int s = r0; s < r1; s++) vals[s] = 0.0;
    if ((dir[d] >> (F == 3 ? ki : comp0)) & 1u) { vals[r0] = 1.0; continue; } // trivial row (its only entry is the diagonal)
    for (int t = inc_ptr[d]; t < inc_ptr[d + 1]; t++) {
      const int e = inc[t] >> 4, i = inc[t] & 15;
      if (e < e0 || e >= e1) continue;
      const double* Ae = scratch + (long)(e - e0) * n * n + (long)(NL * ki + i) * n;
      for (int kj = 0; kj < F; kj++)
        for (int j = 0; j < NL; j++) {
          const long dj = e2d[NL * e + j];
          if ((dir[dj] >> (F == 3 ? kj : comp0)) & 1u) continue;
          const int gj = (int)((long)kj * nd + dj);
          int lo = r0, hi = r1;
          while (lo < hi) { const int mid = (lo + hi) >> 1; if (col[mid] < gj) lo = mid + 1; else hi = mid; }
          vals[lo] += Ae[NL * kj + j];
        }
    }
  }
}
__global__ void k_csr_spmv(long N, const int* __restrict__ rp, const int* __restrict__ col, const double* __restrict__ vals,
                           const double* __restrict__ x, double* __restrict__ y) {
  for (long r = blockIdx.x * (long)blockDim.x + threadIdx.x; r < N; r += (long)gridDim.x * blockDim.x) {
    double sum = 0.0;
    for (int s = rp[r]; s < rp[r + 1]; s++) sum += vals[s] * x[col[s]];
    y[r] = sum;
  }
}
__global__ void k_csr_diag_inverse(long N, const int* __restrict__ rp, const int* __restrict__ col, const double* __restrict__ vals,
                                   double* __restrict__ dinv) {
  for (long r = blockIdx.x * (long)blockDim.x + threadIdx.x; r < N; r += (long)gridDim.x * blockDim.x) {
    double d = 0.0;
    for (int s = rp[r]; s < rp[r + 1]; s++) if (col[s] == r) d = vals[s];
    dinv[r] = d != 0.0 ? 1.0 / d : 0.0;
  }
}

// triangle rule of an operator: the reference drivers' order (OpTraits<OP>::NQ points) or, with intorder 5, the 7-point rule
template <int DEG, int OP, int NQ> void launch_elem_residual_q(Ctx& c, P2Space& S, const Operator& op, const PhysParams& P,
                                                               const double* u, const double* a0, const double* a1) {
  const int g = grid_for(S.nT, 128);
  k_p2_elem_residual<DEG, OP, NQ><<<g, 128, 0, c.stream>>>((int)S.nT, c.ctri.p, c.cx.p, c.cy.p, S.e2d.p, S.fphys.p, c.d_surf.p,
                                                           c.d_surf_dir.p, P, S.nd, op.comp0, op.intorder == 5 ? 3 : 2, u, a0, a1,
                                                           S.scratch.p);
  PNP_CHECK_LAUNCH(); c.launches++;
}
template <int DEG, int OP> void launch_elem_residual(Ctx& c, P2Space& S, const Operator& op, const PhysParams& P, const double* u,
                                                     const double* a0, const double* a1) {
  if (op.intorder == 5) launch_elem_residual_q<DEG, OP, 7>(c, S, op, P, u, a0, a1);
  else launch_elem_residual_q<DEG, OP, OpTraits<OP>::NQ>(c, S, op, P, u, a0, a1);
}
template <int DEG, int OP, int NQ> void launch_elem_jacobian_q(Ctx& c, P2Space& S, int e0, int ne, const PhysParams& P, int mode, double eps,
                                                               const double* u, const double* a0, const double* a1) {
  const int g = grid_for(ne, 64);
  if (mode == 0) k_p2_elem_jacobian<DEG, OP, NQ, 0><<<g, 64, 0, c.stream>>>(e0, ne, c.ctri.p, c.cx.p, c.cy.p, S.e2d.p, P, S.nd, eps, u, a0, a1, S.scratch.p);
  else k_p2_elem_jacobian<DEG, OP, NQ, 1><<<g, 64, 0, c.stream>>>(e0, ne, c.ctri.p, c.cx.p, c.cy.p, S.e2d.p, P, S.nd, eps, u, a0, a1, S.scratch.p);
  PNP_CHECK_LAUNCH(); c.launches++;
}
template <int DEG, int OP> void launch_elem_jacobian(Ctx& c, P2Space& S, int e0, int ne, const Operator& op, const PhysParams& P, int mode,
                                                     double eps, const double* u, const double* a0, const double* a1) {
  if (op.intorder == 5) launch_elem_jacobian_q<DEG, OP, 7>(c, S, e0, ne, P, mode, eps, u, a0, a1);
  else launch_elem_jacobian_q<DEG, OP, OpTraits<OP>::NQ>(c, S, e0, ne, P, mode, eps, u, a0, a1);
}
template <int DEG> void dispatch_residual(Ctx& c, P2Space& S, const Operator& op, const PhysParams& P, const double* u, const double* a0,
                                          const double* a1) {
  switch (op.op) {
    case OP_PB: launch_elem_residual<DEG, OP_PB>(c, S, op, P, u, a0, a1); break;
    case OP_POISSON: launch_elem_residual<DEG, OP_POISSON>(c, S, op, P, u, a0, a1); break;
    case OP_DIFFUSION: launch_elem_residual<DEG, OP_DIFFUSION>(c, S, op, P, u, a0, a1); break;
    case OP_MASS: launch_elem_residual<DEG, OP_MASS>(c, S, op, P, u, a0, a1); break;
    case OP_PNP: launch_elem_residual<DEG, OP_PNP>(c, S, op, P, u, a0, a1); break;
    default: PNP_REQUIRE(false, PNP_E_ARG, "unknown operator");
  }
}
template <int DEG> void dispatch_jacobian(Ctx& c, P2Space& S, int e0, int ne, const Operator& op, const PhysParams& P, int mode, double eps,
                                          const double* u, const double* a0, const double* a1) {
  switch (op.op) {
    case OP_PB: launch_elem_jacobian<DEG, OP_PB>(c, S, e0, ne, op, P, mode, eps, u, a0, a1); break;
    case OP_POISSON: launch_elem_jacobian<DEG, OP_POISSON>(c, S, e0, ne, op, P, mode, eps, u, a0, a1); break;
    case OP_DIFFUSION: launch_elem_jacobian<DEG, OP_DIFFUSION>(c, S, e0, ne, op, P, mode, eps, u, a0, a1); break;
    case OP_MASS: launch_elem_jacobian<DEG, OP_MASS>(c, S, e0, ne, op, P, mode, eps, u, a0, a1); break;
    case OP_PNP: launch_elem_jacobian<DEG, OP_PNP>(c, S, e0, ne, op, P, mode, eps, u, a0, a1); break;
    default: PNP_REQUIRE(false, PNP_E_ARG, "unknown operator");
  }
}

void coefficient_ptrs_p2(Ctx& c, const Operator& op, const double** a0, const double** a1) {
  *a0 = *a1 = nullptr;
  const int need = op.op == OP_POISSON ? 2 : (op.op == OP_DIFFUSION ? 1 : 0);
  if (need >= 1) { PNP_REQUIRE(op.aux0 >= 0 && c.vec(op.aux0).fields == 1, PNP_E_ARG, "operator coefficient 0 not set"); *a0 = c.vec(op.aux0).d.p; }
  if (need >= 2) { PNP_REQUIRE(op.aux1 >= 0 && c.vec(op.aux1).fields == 1, PNP_E_ARG, "operator coefficient 1 not set"); *a1 = c.vec(op.aux1).d.p; }
}

} // namespace

// Lagrange nodes of the space's degree (host side)
struct NodeInfo { int kind, sub, idx; double x, y; };
static NodeInfo node_info(int deg, int n) {
  NodeInfo t;
  if (deg == 2) { PkElem<2>::node_key(n, t.kind, t.sub, t.idx); t.x = PkElem<2>::node_x(n); t.y = PkElem<2>::node_y(n); }
  else { PkElem<3>::node_key(n, t.kind, t.sub, t.idx); t.x = PkElem<3>::node_x(n); t.y = PkElem<3>::node_y(n); }
  return t;
}

// ---- space: edges, element dof map, incidence lists, boundary faces, Dirichlet flags (host, from the canonical mesh) ----
void p2_build(Ctx& c) {
  PNP_REQUIRE(c.n_own == c.nv, PNP_E_ARG, "quadratic elements work on an unpartitioned mesh");
  auto S = std::make_shared<P2Space>();
  const long nv = c.nv, nT = c.nT, nB = c.nB;
  S->nT = nT; S->nv = nv; S->deg = c.degree; S->NL = (c.degree + 1) * (c.degree + 2) / 2;
  const int NL = S->NL, deg = S->deg;
  const std::vector<int> tri = c.ctri.to_host(c.stream), ba = c.cba.to_host(c.stream), bb = c.cbb.to_host(c.stream),
                         bph = c.cbphys.to_host(c.stream);
  auto key = [](int a, int b) { return ((uint64_t)std::min(a, b) << 32) | (uint64_t)std::max(a, b); };
  std::vector<uint64_t> keys(3 * (size_t)nT);
  for (long e = 0; e < nT; e++) for (int f = 0; f < 3; f++) keys[3 * e + f] = key(tri[3 * e + FACE_V2[f][0]], tri[3 * e + FACE_V2[f][1]]);
  std::vector<uint64_t> uk(keys);
  std::sort(uk.begin(), uk.end()); uk.erase(std::unique(uk.begin(), uk.end()), uk.end());
  S->nE = (long)uk.size();
  S->eoff = deg == 3 ? nT : 0; S->voff = S->eoff + (deg - 1) * S->nE; S->nd = S->voff + nv; // elements, edges, vertices
  PNP_REQUIRE(3 * S->nd < (1l << 31) && nT < (1l << 27), PNP_E_MESH, "too many dofs for 32-bit column indices");
  S->h_eva.resize(S->nE); S->h_evb.resize(S->nE);
  for (long k = 0; k < S->nE; k++) { S->h_eva[k] = (int)(uk[k] >> 32); S->h_evb[k] = (int)(uk[k] & 0xffffffffu); }
  std::vector<int> tedge(3 * (size_t)nT);
  for (size_t i = 0; i < keys.size(); i++) tedge[i] = (int)(std::lower_bound(uk.begin(), uk.end(), keys[i]) - uk.begin());
  S->h_e2d.resize(NL * (size_t)nT);
  for (long e = 0; e < nT; e++)
    for (int i = 0; i < NL; i++) {
      const NodeInfo t = node_info(deg, i);
      int d;
      if (t.kind == 2) d = (int)e;
      else if (t.kind == 0) d = (int)(S->voff + tri[3 * e + t.sub]);
      else { // edge dofs are counted from the end vertex with the smaller index (Pk2DLocalFiniteElementMap's variant choice)
        int idx = t.idx;
        if (deg == 3 && tri[3 * e + FACE_V2[t.sub][0]] > tri[3 * e + FACE_V2[t.sub][1]]) idx = 1 - idx;
        d = (int)(S->eoff + (deg - 1) * tedge[3 * e + t.sub] + idx);
      }
      S->h_e2d[NL * e + i] = d;
    }
  // boundary faces: the element face whose edge is a boundary segment carries that segment's surface; neighbours across faces
  std::vector<int> edge_phys(S->nE, -1), edge_elem0(S->nE, -1), edge_elem1(S->nE, -1);
  for (long s = 0; s < nB; s++) {
    auto it = std::lower_bound(uk.begin(), uk.end(), key(ba[s], bb[s]));
    PNP_REQUIRE(it != uk.end() && *it == key(ba[s], bb[s]), PNP_E_MESH, "a boundary segment is not an edge of the mesh");
    edge_phys[it - uk.begin()] = bph[s]; // (a segment listed twice: the later one wins)
  }
  for (long e = 0; e < nT; e++) for (int f = 0; f < 3; f++) { int& a = edge_elem0[tedge[3 * e + f]]; if (a < 0) a = (int)e; else edge_elem1[tedge[3 * e + f]] = (int)e; }
  S->h_fphys.assign(3 * (size_t)nT, -1); S->h_nbr.assign(3 * (size_t)nT, -1);
  for (long e = 0; e < nT; e++)
    for (int f = 0; f < 3; f++) {
      const int k = tedge[3 * e + f];
      if (edge_elem1[k] < 0) { PNP_REQUIRE(edge_phys[k] >= 0, PNP_E_MESH, "boundary face without boundary segment"); S->h_fphys[3 * e + f] = edge_phys[k]; }
      else S->h_nbr[3 * e + f] = edge_elem0[k] == e ? edge_elem1[k] : edge_elem0[k];
    }
  // incidence lists, elements ascending
  std::vector<int> ptr(S->nd + 1, 0), inc(NL * (size_t)nT);
  for (size_t i = 0; i < S->h_e2d.size(); i++) ptr[S->h_e2d[i] + 1]++;
  for (long d = 0; d < S->nd; d++) ptr[d + 1] += ptr[d];
  { std::vector<int> fill(ptr.begin(), ptr.end() - 1);
    for (long e = 0; e < nT; e++) for (int i = 0; i < NL; i++) inc[fill[S->h_e2d[NL * e + i]]++] = (int)(e * 16 + i); }
  { const int order[3] = {0, 2, 1};
    for (long e = 0; e < nT; e++) for (int fi = 0; fi < 3; fi++) if (S->h_fphys[3 * e + order[fi]] >= 0) S->h_bfaces.push_back((int)(e * 4 + order[fi])); }
  S->bfaces.alloc(S->h_bfaces.size()); S->bfaces.upload(S->h_bfaces.data(), S->h_bfaces.size(), c.stream);
  S->h_edge_phys.assign(S->nE, -1);
  for (long k = 0; k < S->nE; k++) if (edge_elem1[k] < 0) S->h_edge_phys[k] = edge_phys[k];
  S->e2d.alloc(S->h_e2d.size()); S->e2d.upload(S->h_e2d.data(), S->h_e2d.size(), c.stream);
  S->fphys.alloc(S->h_fphys.size()); S->fphys.upload(S->h_fphys.data(), S->h_fphys.size(), c.stream);
  S->inc_ptr.alloc(ptr.size()); S->inc_ptr.upload(ptr.data(), ptr.size(), c.stream);
  S->inc.alloc(inc.size()); S->inc.upload(inc.data(), inc.size(), c.stream);
  S->h_dir.assign(S->nd, 0);
  S->dir.alloc(S->nd); S->dir.zero(c.stream);
  PNP_CUDA(cudaStreamSynchronize(c.stream));
  c.p2 = S; c.p2_nd = S->nd;
}

// Dirichlet flags (ConformingDirichletConstraints): a Dirichlet face constrains its end vertices and its edge dofs
void p2_constraints(Ctx& c) {
  PNP_REQUIRE(c.p2 && c.params.set, PNP_E_ARG, "quadratic space / parameters not set");
  P2Space& S = *c.p2;
  S.h_dir.assign(S.nd, 0);
  for (long k = 0; k < S.nE; k++) {
    const int ph = S.h_edge_phys[k];
    if (ph < 0) continue;
    PNP_REQUIRE(ph < c.params.n_surfaces, PNP_E_CONFIG, "physical tag without [surface_i] section");
    unsigned char m = 0;
    for (int comp = 0; comp < 3; comp++) if (c.params.surfaces[ph].btype[comp] == 0) m |= (unsigned char)(1u << comp);
    for (int l = 0; l < S.deg - 1; l++) S.h_dir[S.eoff + (S.deg - 1) * k + l] |= m;
    S.h_dir[S.voff + S.h_eva[k]] |= m; S.h_dir[S.voff + S.h_evb[k]] |= m;
  }
  S.dir.upload(S.h_dir.data(), S.nd, c.stream);
  PNP_CUDA(cudaStreamSynchronize(c.stream));
  // the patterns depend on the constraints: drop them, and detach every matrix that points into one (it is re-initialised
  // by the next assembly / import; using it before that is an error, not a read of freed memory)
  S.patterns.clear();
  auto detach = [](Matrix& A) { A.csr_rp = nullptr; A.csr_col = nullptr; A.csr_pattern = nullptr; A.csr_nnz = 0; A.csr_n = 0; A.vals.release(); };
  for (auto& m : c.mats) if (m) detach(*m);
  detach(c.ws_A); detach(c.ws_B);
}

static P2Space& space(Ctx& c) {
  PNP_REQUIRE(c.degree >= 2 && c.p2, PNP_E_ARG, "no quadratic / cubic space (pnp_space_set_degree before pnp_mesh_finalize)");
  return *c.p2;
}

// FullVolumePattern + ISTLBCRSMatrixBackend<1,1>: rows and columns ascending, constrained links dropped
P2Pattern& p2_pattern(Ctx& c, int F, int comp0) {
  P2Space& S = space(c);
  auto& slot = S.patterns[4 * F + comp0];
  if (slot) return *slot;
  slot = std::make_unique<P2Pattern>();
  P2Pattern& Pn = *slot;
  const long nd = S.nd, nT = S.nT, N = F * nd;
  const int NL = S.NL;
  auto dirichlet = [&](int k, long d) { return (S.h_dir[d] >> (F == 3 ? k : comp0)) & 1; };
  PNP_REQUIRE((long)NL * NL * nT < (1l << 31), PNP_E_MESH, "too many element couplings for 32-bit pattern offsets");
  std::vector<int> cnt(nd + 1, 0);
  for (long e = 0; e < nT; e++) for (int i = 0; i < NL; i++) cnt[S.h_e2d[NL * e + i] + 1] += NL;
  for (long d = 0; d < nd; d++) cnt[d + 1] += cnt[d];
  std::vector<int> raw(cnt[nd]), fill(cnt.begin(), cnt.end() - 1), nptr(nd + 1, 0), nbr;
  for (long e = 0; e < nT; e++) for (int i = 0; i < NL; i++) for (int j = 0; j < NL; j++) raw[fill[S.h_e2d[NL * e + i]]++] = S.h_e2d[NL * e + j];
  for (long d = 0; d < nd; d++) {
    auto b = raw.begin() + cnt[d], e = raw.begin() + cnt[d + 1];
    std::sort(b, e); e = std::unique(b, e);
    nbr.insert(nbr.end(), b, e); nptr[d + 1] = (int)nbr.size();
  }
  Pn.h_rp.assign(N + 1, 0);
  for (int ki = 0; ki < F; ki++) for (long d = 0; d < nd; d++) {
    int len = 0;
    if (dirichlet(ki, d)) len = 1;
    else for (int kj = 0; kj < F; kj++) for (int t = nptr[d]; t < nptr[d + 1]; t++) len += !dirichlet(kj, nbr[t]);
    Pn.h_rp[ki * nd + d + 1] = len;
  }
  { long total = 0;
    for (long r = 0; r < N; r++) total += Pn.h_rp[r + 1];
    PNP_REQUIRE(total < (1l << 31), PNP_E_MESH, "too many matrix entries for 32-bit CSR offsets"); }
  for (long r = 0; r < N; r++) Pn.h_rp[r + 1] += Pn.h_rp[r];
  Pn.nnz = Pn.h_rp[N];
  Pn.h_col.resize(Pn.nnz);
  for (int ki = 0; ki < F; ki++) for (long d = 0; d < nd; d++) {
    int o = Pn.h_rp[ki * nd + d];
    if (dirichlet(ki, d)) { Pn.h_col[o] = (int)(ki * nd + d); continue; }
    for (int kj = 0; kj < F; kj++) for (int t = nptr[d]; t < nptr[d + 1]; t++) if (!dirichlet(kj, nbr[t])) Pn.h_col[o++] = (int)(kj * nd + nbr[t]);
  }
  Pn.rp.alloc(Pn.h_rp.size()); Pn.rp.upload(Pn.h_rp.data(), Pn.h_rp.size(), c.stream);
  Pn.col.alloc(Pn.h_col.size()); Pn.col.upload(Pn.h_col.data(), Pn.h_col.size(), c.stream);
  PNP_CUDA(cudaStreamSynchronize(c.stream));
  return Pn;
}

void p2_matrix_init(Ctx& c, Matrix& A, const Operator& op) {
  const int F = op_fields(op.op);
  P2Pattern& Pn = p2_pattern(c, F, F == 3 ? 0 : op.comp0);
  A.op = op.op; A.nplanes = op_planes(op.op); A.comp0 = op.comp0;
  A.csr_rp = Pn.rp.p; A.csr_col = Pn.col.p; A.csr_nnz = Pn.nnz; A.csr_n = F * space(c).nd; A.csr_pattern = &Pn;
  if (A.vals.n != (size_t)Pn.nnz) { A.vals.alloc(Pn.nnz); A.vals.zero(c.stream); }
}

void p2_assemble_residual(Ctx& c, const Operator& op, Vec& u, Vec& r) {
  P2Space& S = space(c);
  PNP_REQUIRE(c.params.set, PNP_E_ARG, "parameters not set");
  const int F = op_fields(op.op), n = S.NL * F;
  PNP_REQUIRE(u.fields == F && r.fields == F, PNP_E_ARG, "vector field count does not match the operator");
  const double *a0, *a1;
  coefficient_ptrs_p2(c, op, &a0, &a1);
  if (S.scratch.n < (size_t)n * S.nT) S.scratch.alloc((size_t)n * S.nT);
  const PhysParams P = c.phys(op.valency);
  if (S.deg == 2) dispatch_residual<2>(c, S, op, P, u.d.p, a0, a1);
  else dispatch_residual<3>(c, S, op, P, u.d.p, a0, a1);
  k_p2_gather_residual<<<grid_for(F * S.nd, 256), 256, 0, c.stream>>>(S.nd, F, S.NL, op.comp0, S.inc_ptr.p, S.inc.p, S.dir.p, S.scratch.p, r.d.p);
  PNP_CHECK_LAUNCH(); c.launches++;
  c.acct(Ctx::ACC_ASSEMBLY, (double)S.nT * (12 + 24 + 16.0 * n) + (double)F * S.nd * 16.0);
}

void p2_assemble_jacobian(Ctx& c, const Operator& op, Vec& u, Matrix& A, int mode, double eps) {
  P2Space& S = space(c);
  PNP_REQUIRE(c.params.set, PNP_E_ARG, "parameters not set");
  const int F = op_fields(op.op), n = S.NL * F;
  PNP_REQUIRE(u.fields == F && A.op == op.op, PNP_E_ARG, "vector / matrix do not match the operator");
  PNP_REQUIRE(mode == 0 || mode == 1, PNP_E_ARG, "unknown jacobian mode");
  p2_matrix_init(c, A, op);
  const double *a0, *a1;
  coefficient_ptrs_p2(c, op, &a0, &a1);
  // the element matrices (n*n doubles each: 2.6 KB for the quadratic, 7.2 KB for the cubic 3-field operator) pass through a
  // scratch block of bounded size: chunks of elements in ascending order, each gathered into the rows before the next
  long chunk = tune().p2_chunk > 0 ? tune().p2_chunk : (long)(2e9 / (8.0 * n * n));
  chunk = std::max(1l, std::min(chunk, S.nT));
  if (S.scratch.n < (size_t)n * n * chunk) S.scratch.alloc((size_t)n * n * chunk);
  const PhysParams P = c.phys(op.valency);
  for (long e0 = 0; e0 < S.nT; e0 += chunk) {
    const int ne = (int)std::min(chunk, S.nT - e0);
    if (S.deg == 2) dispatch_jacobian<2>(c, S, (int)e0, ne, op, P, mode, eps, u.d.p, a0, a1);
    else dispatch_jacobian<3>(c, S, (int)e0, ne, op, P, mode, eps, u.d.p, a0, a1);
    k_p2_gather_jacobian<<<grid_for(F * S.nd, 128), 128, 0, c.stream>>>((int)e0, (int)e0 + ne, S.nd, F, S.NL, op.comp0, S.inc_ptr.p, S.inc.p,
                                                                        S.e2d.p, S.dir.p, A.csr_rp, A.csr_col, S.scratch.p, A.vals.p);
    PNP_CHECK_LAUNCH(); c.launches++;
  }
  c.last_u = u.d.p; c.last_op = op; c.last_mode = mode; c.last_eps = eps; c.last_vals = A.vals.p; // (what the p-multigrid re-discretises)
  c.acct(Ctx::ACC_ASSEMBLY, (double)S.nT * 16.0 * n * n + 12.0 * (double)A.csr_nnz);
}

void csr_spmv(Ctx& c, const Matrix& A, const double* x, double* y) {
  k_csr_spmv<<<grid_for(A.csr_n, 256), 256, 0, c.stream>>>(A.csr_n, A.csr_rp, A.csr_col, A.vals.p, x, y);
  PNP_CHECK_LAUNCH(); c.launches++;
  c.acct(Ctx::ACC_SPMV_FINE, 12.0 * (double)A.csr_nnz + 20.0 * (double)A.csr_n);
}
void csr_diag_inverse(Ctx& c, const Matrix& A, double* dinv) {
  k_csr_diag_inverse<<<grid_for(A.csr_n, 256), 256, 0, c.stream>>>(A.csr_n, A.csr_rp, A.csr_col, A.vals.p, dinv);
  PNP_CHECK_LAUNCH(); c.launches++;
}

const unsigned char* p2_dirichlet_flags(Ctx& c) { return space(c).dir.p; }

// ---- SeqSSOR / SeqILU0 on the CSR matrix (the reference's default backend is BiCGSTAB + SSOR, instationary_pnp_from_pb_md.hh:
// 188-191).  A row-order sweep only depends on the relative order of COUPLED rows (pnp_sweep.cuh): level(i) = 1 + max level of
// the coupled rows before i; rows of one level are mutually uncoupled (the pattern is structurally symmetric) and run in
// parallel, levels in sequence -- ascending for the forward sweep, descending for the backward one.  One launch per level;
// every row does the oracle's operations in the oracle's order (ascending columns, no FMA), so the application is bit-identical
// to the sequential sweep.
namespace {
__global__ void k_csr_gs_level(const int* __restrict__ rows, int n, const int* __restrict__ rp, const int* __restrict__ col,
                               const int* __restrict__ diag, const double* __restrict__ vals, const double* __restrict__ d, double* x) {
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    const int i = rows[t];
    double sum = d[i];
    for (int k = rp[i]; k < rp[i + 1]; k++) sum -= vals[k] * x[col[k]];
    x[i] += sum / vals[diag[i]];
  }
}
// bilu0_decomposition of the rows of one level (dune-istl 2.2 ilu.hh; oracle ilu0_decompose)
__global__ void k_csr_ilu0_level(const int* __restrict__ rows, int n, const int* __restrict__ rp, const int* __restrict__ col,
                                 const int* __restrict__ diag, double* lu) {
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    const int i = rows[t], r1 = rp[i + 1];
    for (int kj = rp[i]; kj < r1 && col[kj] < i; kj++) {
      const int j = col[kj];
      const double l = lu[kj] * lu[diag[j]];
      lu[kj] = l;
      int ki = kj + 1;
      for (int kk = diag[j] + 1; kk < rp[j + 1]; kk++) {
        while (ki < r1 && col[ki] < col[kk]) ki++;
        if (ki < r1 && col[ki] == col[kk]) lu[ki] -= l * lu[kk];
      }
    }
    lu[diag[i]] = 1.0 / lu[diag[i]];
  }
}
__global__ void k_csr_ilu0_forward(const int* __restrict__ rows, int n, const int* __restrict__ rp, const int* __restrict__ col,
                                   const double* __restrict__ lu, const double* __restrict__ d, double* x) {
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    const int i = rows[t];
    double sum = d[i];
    for (int k = rp[i]; k < rp[i + 1] && col[k] < i; k++) sum -= lu[k] * x[col[k]];
    x[i] = sum;
  }
}
__global__ void k_csr_ilu0_backward(const int* __restrict__ rows, int n, const int* __restrict__ rp, const int* __restrict__ col,
                                    const int* __restrict__ diag, const double* __restrict__ lu, double* x) {
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    const int i = rows[t];
    double sum = x[i];
    for (int k = diag[i] + 1; k < rp[i + 1]; k++) sum -= lu[k] * x[col[k]];
    x[i] = sum * lu[diag[i]];
  }
}
P2Pattern& pattern_of(const Matrix& A) {
  PNP_REQUIRE(A.csr_pattern, PNP_E_ARG, "matrix has no CSR pattern");
  return *static_cast<P2Pattern*>(A.csr_pattern);
}
void build_levels(Ctx& c, P2Pattern& Pn) {
  if (Pn.nlev) return;
  const long N = (long)Pn.h_rp.size() - 1;
  std::vector<int> lev(N), dg(N, -1);
  int nlev = 0;
  for (long i = 0; i < N; i++) {
    int l = 0;
    for (int k = Pn.h_rp[i]; k < Pn.h_rp[i + 1]; k++) {
      const int j = Pn.h_col[k];
      if (j < i) l = std::max(l, lev[j] + 1);
      else if (j == i) dg[i] = k;
    }
    PNP_REQUIRE(dg[i] >= 0, PNP_E_ARG, "matrix row without diagonal entry");
    lev[i] = l; nlev = std::max(nlev, l + 1);
  }
  Pn.lev_ptr.assign(nlev + 1, 0);
  for (long i = 0; i < N; i++) Pn.lev_ptr[lev[i] + 1]++;
  for (int l = 0; l < nlev; l++) Pn.lev_ptr[l + 1] += Pn.lev_ptr[l];
  std::vector<int> order(N), fill(Pn.lev_ptr.begin(), Pn.lev_ptr.end() - 1);
  for (long i = 0; i < N; i++) order[fill[lev[i]]++] = (int)i;
  Pn.order.alloc(N); Pn.order.upload(order.data(), N, c.stream);
  Pn.diag.alloc(N); Pn.diag.upload(dg.data(), N, c.stream);
  PNP_CUDA(cudaStreamSynchronize(c.stream));
  Pn.nlev = nlev;
}
template <class Fn> void for_levels(Ctx& c, const P2Pattern& Pn, bool ascending, Fn fn) {
  for (int s = 0; s < Pn.nlev; s++) {
    const int l = ascending ? s : Pn.nlev - 1 - s, n = Pn.lev_ptr[l + 1] - Pn.lev_ptr[l];
    fn(Pn.order.p + Pn.lev_ptr[l], n, grid_for(n, 128));
    PNP_CHECK_LAUNCH(); c.launches++;
  }
}
} // namespace

void csr_sweep_setup(Ctx& c, Solver& S, const Matrix& A, bool ilu) {
  P2Pattern& Pn = pattern_of(A);
  build_levels(c, Pn);
  S.csr_levels = Pn.nlev;
  if (!ilu) return;
  if (S.csr_lu.n != (size_t)A.csr_nnz) S.csr_lu.alloc(A.csr_nnz);
  vec_copy(c, A.vals.p, S.csr_lu.p, A.csr_nnz);
  for_levels(c, Pn, true, [&](const int* rows, int n, int g) {
    k_csr_ilu0_level<<<g, 128, 0, c.stream>>>(rows, n, A.csr_rp, A.csr_col, Pn.diag.p, S.csr_lu.p);
  });
}
// v = 0; `prec_steps` x (forward sweep, backward sweep), relaxation factor 1
void csr_ssor_apply(Ctx& c, Solver& S, const Matrix& A, const double* d, double* y) {
  P2Pattern& Pn = pattern_of(A);
  vec_zero(c, y, A.csr_n);
  for (int s = 0; s < S.prec_steps; s++)
    for (int dir = 0; dir < 2; dir++)
      for_levels(c, Pn, dir == 0, [&](const int* rows, int n, int g) {
        k_csr_gs_level<<<g, 128, 0, c.stream>>>(rows, n, A.csr_rp, A.csr_col, Pn.diag.p, A.vals.p, d, y);
      });
  c.acct(Ctx::ACC_SPMV_FINE, 2.0 * S.prec_steps * (12.0 * (double)A.csr_nnz + 28.0 * (double)A.csr_n));
}
void csr_ilu0_apply(Ctx& c, Solver& S, const Matrix& A, const double* d, double* y) {
  P2Pattern& Pn = pattern_of(A);
  for_levels(c, Pn, true, [&](const int* rows, int n, int g) {
    k_csr_ilu0_forward<<<g, 128, 0, c.stream>>>(rows, n, A.csr_rp, A.csr_col, S.csr_lu.p, d, y);
  });
  for_levels(c, Pn, false, [&](const int* rows, int n, int g) {
    k_csr_ilu0_backward<<<g, 128, 0, c.stream>>>(rows, n, A.csr_rp, A.csr_col, Pn.diag.p, S.csr_lu.p, y);
  });
  c.acct(Ctx::ACC_SPMV_FINE, 12.0 * (double)A.csr_nnz + 32.0 * (double)A.csr_n);
}

// ---- the boundary's views of the space ----
void p2_sizes(Ctx& c, long* nE, long* nd) { P2Space& S = space(c); if (nE) *nE = S.nE; if (nd) *nd = S.nd; }
void p2_offsets(Ctx& c, long* eoff, long* voff) { P2Space& S = space(c); if (eoff) *eoff = S.eoff; if (voff) *voff = S.voff; }
void p2_edges(Ctx& c, int* eva, int* evb) {
  P2Space& S = space(c);
  if (eva) std::copy(S.h_eva.begin(), S.h_eva.end(), eva);
  if (evb) std::copy(S.h_evb.begin(), S.h_evb.end(), evb);
}
void p2_constraints_get(Ctx& c, const Operator& op, char* out) {
  P2Space& S = space(c);
  const int F = op_fields(op.op);
  for (int k = 0; k < F; k++) for (long d = 0; d < S.nd; d++) out[k * S.nd + d] = (char)((S.h_dir[d] >> (F == 3 ? k : op.comp0)) & 1);
}
long p2_pattern_export(Ctx& c, const Operator& op, int* rowptr, int* col) {
  const int F = op_fields(op.op);
  P2Pattern& Pn = p2_pattern(c, F, F == 3 ? 0 : op.comp0);
  if (rowptr) std::copy(Pn.h_rp.begin(), Pn.h_rp.end(), rowptr);
  if (col) std::copy(Pn.h_col.begin(), Pn.h_col.end(), col);
  return Pn.nnz;
}
void p2_matrix_import(Ctx& c, const Operator& op, Matrix& A, const int* rowptr, const int* col, const double* val) {
  p2_matrix_init(c, A, op);
  const int F = op_fields(op.op);
  P2Pattern& Pn = p2_pattern(c, F, F == 3 ? 0 : op.comp0);
  PNP_REQUIRE(std::equal(Pn.h_rp.begin(), Pn.h_rp.end(), rowptr) && std::equal(Pn.h_col.begin(), Pn.h_col.end(), col), PNP_E_ARG,
              "CSR pattern differs from the operator's pattern (pnp_pattern_get)");
  A.vals.upload(val, Pn.nnz, c.stream);
  PNP_CUDA(cudaStreamSynchronize(c.stream));
}

// ---- p-multigrid (ISTLBackend_NOVLP_CG_AMG_SSOR with -DPDEGREE=2,3: src/Makefile.am:106-110) ----
// V-cycle on two spaces: SSOR smoothing on the Pk matrix (the level-scheduled sweep above), coarse correction in the P1 space of
// the same mesh -- the P1 operator re-discretised at the vertex values of the state the Pk Jacobian was assembled at (the star
// path's own assembly on a child context), one cycle of the star path's multigrid as its solver.  Transfers: P = the P1
// interpolant at the Pk nodes (vertex dof: the vertex; edge dof: its two end vertices; bubble: the three vertices), stored per
// field as a CSR matrix with the rows of constrained Pk dofs empty, and its transpose with the rows of constrained vertices
// empty; both are applied by the CSR SpMV kernel.  Scalar operators only: with the 3-field PNP system the preconditioned BiCGSTAB
// broke down in the first measurements (not understood yet), so that combination answers PNP_E_ARG instead of failing late.
// Nothing here has a counterpart in ISTL's aggregation AMG beyond the role
// (a multigrid-preconditioned Krylov method for the higher-degree programs); the bar is convergence, as for the P1 multigrid.
void amg_setup(Ctx&, Solver&, const Matrix&);                        // pnp_amg.cu
void amg_apply(Ctx&, Solver&, const Matrix&, const double* d, double* y);
namespace {
struct PMg {
  Ctx* parent = nullptr; Ctx* child = nullptr; // the P1 context lives as long as this object (a solver of the parent)
  ~PMg() { if (parent && child) ctx_destroy_owned_child(*parent, child); }
  int F = 0, comp0 = -1;
  Operator op1; Vec u1, b1, e1; Matrix A1; Solver S1;
  DBuf<int> P_rp[3], P_col[3], PT_rp[3], PT_col[3];
  DBuf<double> P_val[3], PT_val[3];
  long P_nnz[3] = {0, 0, 0};
  DBuf<double> r, t, z, lex1;
};
void upload_csr(Ctx& c, const std::vector<int>& rp, const std::vector<int>& col, const std::vector<double>& val, DBuf<int>& drp,
                DBuf<int>& dcol, DBuf<double>& dval) {
  drp.alloc(rp.size()); drp.upload(rp.data(), rp.size(), c.stream);
  dcol.alloc(std::max<size_t>(1, col.size())); if (!col.empty()) dcol.upload(col.data(), col.size(), c.stream);
  dval.alloc(std::max<size_t>(1, val.size())); if (!val.empty()) dval.upload(val.data(), val.size(), c.stream);
}
// y[rows] = M x through the CSR SpMV kernel
void apply_csr(Ctx& c, long rows, const DBuf<int>& rp, const DBuf<int>& col, const DBuf<double>& val, long nnz, const double* x, double* y) {
  k_csr_spmv<<<grid_for(rows, 256), 256, 0, c.stream>>>(rows, rp.p, col.p, val.p, x, y);
  PNP_CHECK_LAUNCH(); c.launches++;
  c.acct(Ctx::ACC_TRANSFER, 12.0 * (double)nnz + 20.0 * (double)rows);
}
void build_transfers(Ctx& c, P2Space& S, PMg& M, int F, int comp0) {
  const long nd = S.nd, nv = S.nv, nT = S.nT;
  const std::vector<int> tri = c.ctri.to_host(c.stream);
  for (int k = 0; k < F; k++) {
    const int comp = F == 3 ? k : comp0;
    auto fixed = [&](long d) { return (S.h_dir[d] >> comp) & 1; };
    // rows of P: (column, weight) lists per Pk dof
    std::vector<int> rp(nd + 1, 0), col; std::vector<double> val;
    col.reserve(3 * (size_t)nd); val.reserve(3 * (size_t)nd);
    for (long d = 0; d < nd; d++) {
      if (!fixed(d)) {
        if (d >= S.voff) { col.push_back((int)(d - S.voff)); val.push_back(1.0); }
        else if (d >= S.eoff) {
          const long e = (d - S.eoff) / (S.deg - 1); const int idx = (int)((d - S.eoff) % (S.deg - 1));
          const double t = (idx + 1.0) / S.deg; // distance from the smaller end vertex, in edge lengths
          col.push_back(S.h_eva[e]); val.push_back(1.0 - t);
          col.push_back(S.h_evb[e]); val.push_back(t);
        } else {
          int v[3] = {tri[3 * d], tri[3 * d + 1], tri[3 * d + 2]};
          std::sort(v, v + 3);
          for (int i = 0; i < 3; i++) { col.push_back(v[i]); val.push_back(1.0 / 3.0); }
        }
      }
      rp[d + 1] = (int)col.size();
    }
    M.P_nnz[k] = (long)col.size();
    upload_csr(c, rp, col, val, M.P_rp[k], M.P_col[k], M.P_val[k]);
    // transpose, rows of constrained vertices empty
    std::vector<int> trp(nv + 1, 0), tcol(col.size()); std::vector<double> tval(col.size());
    for (size_t i = 0; i < col.size(); i++) if (!fixed(S.voff + col[i])) trp[col[i] + 1]++;
    for (long v = 0; v < nv; v++) trp[v + 1] += trp[v];
    std::vector<int> fill(trp.begin(), trp.end() - 1);
    for (long d = 0; d < nd; d++)
      for (int i = rp[d]; i < rp[d + 1]; i++) if (!fixed(S.voff + col[i])) { const int o = fill[col[i]]++; tcol[o] = (int)d; tval[o] = val[i]; }
    tcol.resize(trp[nv]); tval.resize(trp[nv]);
    upload_csr(c, trp, tcol, tval, M.PT_rp[k], M.PT_col[k], M.PT_val[k]);
  }
  (void)nT;
  PNP_CUDA(cudaStreamSynchronize(c.stream));
  M.F = F; M.comp0 = comp0;
}
PMg& pmg_of(Solver& S) { return *static_cast<PMg*>(S.pmg.get()); }
} // namespace

void pmg_setup(Ctx& c, Solver& S, const Matrix& A) {
  P2Space& Sp = space(c);
  PNP_REQUIRE(c.last_u && c.last_vals == A.vals.p, PNP_E_ARG,
              "quadratic / cubic elements: the multigrid re-discretises the last assembled Jacobian (assemble, then solve; combined "
              "or imported matrices take SSOR / ILU0)");
  PNP_REQUIRE(op_fields(c.last_op.op) == 1, PNP_E_ARG,
              "quadratic / cubic elements: the multigrid serves the scalar operators (PB, Poisson, transport, mass); the 3-field system "
              "takes SSOR / ILU0");
  if (!S.pmg) S.pmg = std::shared_ptr<void>(new PMg, [](void* p) { delete static_cast<PMg*>(p); });
  PMg& M = pmg_of(S);
  const Operator& op = c.last_op;
  const int F = op_fields(op.op), comp0 = F == 3 ? 0 : op.comp0;
  const long nd = Sp.nd, nv = Sp.nv;
  if (!M.child) { // the P1 space of the same mesh, on the star layout
    M.parent = &c; M.child = ctx_make_owned_child(c);
    const std::vector<double> x = c.cx.to_host(c.stream), y = c.cy.to_host(c.stream);
    const std::vector<int> tri = c.ctri.to_host(c.stream), ba = c.cba.to_host(c.stream), bb = c.cbb.to_host(c.stream),
                           bph = c.cbphys.to_host(c.stream);
    mesh_set(*M.child, nv, x.data(), y.data(), c.nT, tri.data(), c.nB, ba.data(), bb.data(), bph.data(), nv);
    mesh_finalize(*M.child, true);
    M.F = 0;
  }
  Ctx& c1 = *M.child;
  if (M.F != F || M.comp0 != comp0) {
    build_transfers(c, Sp, M, F, comp0);
    M.u1.fields = M.b1.fields = M.e1.fields = F;
    M.u1.d.alloc((size_t)F * nv); M.b1.d.alloc((size_t)F * nv); M.e1.d.alloc((size_t)F * nv);
    M.r.alloc((size_t)F * nd); M.t.alloc((size_t)F * nd); M.z.alloc((size_t)F * nd); M.lex1.alloc((size_t)F * nv);
    c1.vecs.clear();
    for (int a = 0; a < 2; a++) { auto v = std::make_unique<Vec>(); v->fields = 1; v->d.alloc(nv); v->d.zero(c.stream); c1.vecs.push_back(std::move(v)); }
  }
  // inject the state and the operator's coefficient fields: their vertex values
  auto inject = [&](const double* pk, int fields, Vec& dst) {
    for (int k = 0; k < fields; k++)
      PNP_CUDA(cudaMemcpyAsync(M.lex1.p + (long)k * nv, pk + (long)k * nd + Sp.voff, nv * sizeof(double), cudaMemcpyDeviceToDevice, c.stream));
    vec_from_lex_device(c1, dst, M.lex1.p);
  };
  inject(c.last_u, F, M.u1);
  M.op1 = op; M.op1.intorder = 0; M.op1.aux0 = M.op1.aux1 = -1;
  if (op.aux0 >= 0) { inject(c.vec(op.aux0).d.p, 1, *c1.vecs[0]); M.op1.aux0 = 0; }
  if (op.aux1 >= 0) { inject(c.vec(op.aux1).d.p, 1, *c1.vecs[1]); M.op1.aux1 = 1; }
  M.A1.op = op.op; M.A1.nplanes = op_planes(op.op);
  if (M.A1.vals.n != (size_t)M.A1.nplanes * c1.nslots) M.A1.vals.alloc((size_t)M.A1.nplanes * c1.nslots);
  assemble_jacobian(c1, M.op1, M.u1, M.A1, JAC_ANALYTIC, 1e-11);
  M.S1.prec = PNP_PREC_AMG; M.S1.prec_steps = (int)S.opt("pmg_coarse_steps", 2); M.S1.verbosity = S.verbosity;
  M.S1.opts = S.opts; // amg_* options reach the P1 multigrid
  amg_setup(c1, M.S1, M.A1);
  csr_sweep_setup(c, S, A, false);
  c.absorb(c1);
}

// y = M^-1 d: SSOR(prec_steps) from zero, coarse correction, SSOR(prec_steps) on the new defect
void pmg_apply(Ctx& c, Solver& S, const Matrix& A, const double* d, double* y) {
  P2Space& Sp = space(c);
  PMg& M = pmg_of(S);
  Ctx& c1 = *M.child;
  const long nd = Sp.nd, nv = Sp.nv, n = (long)M.F * nd;
  csr_ssor_apply(c, S, A, d, y);
  csr_spmv(c, A, y, M.t.p);
  vec_copy(c, d, M.r.p, n); vec_axpy(c, -1.0, M.t.p, M.r.p, n);
  for (int k = 0; k < M.F; k++) apply_csr(c, nv, M.PT_rp[k], M.PT_col[k], M.PT_val[k], M.P_nnz[k], M.r.p + (long)k * nd, M.lex1.p + (long)k * nv);
  vec_from_lex_device(c1, M.b1, M.lex1.p);
  amg_apply(c1, M.S1, M.A1, M.b1.d.p, M.e1.d.p);
  vec_to_lex_device(c1, M.e1, M.lex1.p);
  for (int k = 0; k < M.F; k++) apply_csr(c, nd, M.P_rp[k], M.P_col[k], M.P_val[k], M.P_nnz[k], M.lex1.p + (long)k * nv, M.t.p + (long)k * nd);
  vec_axpy(c, 1.0, M.t.p, y, n);
  csr_spmv(c, A, y, M.t.p);
  vec_copy(c, d, M.r.p, n); vec_axpy(c, -1.0, M.t.p, M.r.p, n);
  csr_ssor_apply(c, S, A, M.r.p, M.z.p);
  vec_axpy(c, 1.0, M.z.p, y, n);
  c.absorb(c1);
}

// ---- the time loop's diagnostics with quadratic functions (pnp_output.cu has the linear ones) ----
namespace {
// calcIonFlux (ionFlux.hh:51-83): one thread per boundary face; fields and gradients at the face centre
template <int DEG>
__global__ void k_p2_ion_flux(const int* __restrict__ faces, int nF, const int* __restrict__ tri, const double* __restrict__ cx,
                              const double* __restrict__ cy, const int* __restrict__ e2d, const double* __restrict__ phi,
                              const double* __restrict__ cp, const double* __restrict__ cm, int cylindrical, double PI,
                              double* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nF) return;
  const int e = faces[t] >> 2, f = faces[t] & 3;
  const int la = f == 2 ? 1 : 0, lb = f == 0 ? 1 : 2, lc = 3 - la - lb;
  const int v[3] = {tri[3 * e], tri[3 * e + 1], tri[3 * e + 2]};
  using E = PkElem<DEG>;
  constexpr int NL = E::NL;
  const typename E::Geo2 G = E::make_geo2(cx[v[0]], cy[v[0]], cx[v[1]], cy[v[1]], cx[v[2]], cy[v[2]]);
  const double VX[3] = {0.0, 1.0, 0.0}, VY[3] = {0.0, 0.0, 1.0};
  const double ax = cx[v[la]], ay = cy[v[la]], bx = cx[v[lb]], by = cy[v[lb]];
  const double ex = 0.5 * (ax + bx), ey = 0.5 * (ay + by);
  const typename E::BasisAt B = E::basis_at(G, 0.5 * (VX[la] + VX[lb]), 0.5 * (VY[la] + VY[lb]));
  double vcp = 0, vcm = 0, gphi[2] = {0, 0}, gcp[2] = {0, 0}, gcm[2] = {0, 0};
  for (int k = 0; k < NL; k++) {
    const int d = e2d[NL * e + k];
    vcp += cp[d] * B.phi[k]; vcm += cm[d] * B.phi[k];
    for (int r = 0; r < 2; r++) { gphi[r] += phi[d] * B.g[k][r]; gcp[r] += cp[d] * B.g[k][r]; gcm[r] += cm[d] * B.g[k][r]; }
  }
  const double len = sqrt((bx - ax) * (bx - ax) + (by - ay) * (by - ay));
  double factor = len;
  if (cylindrical) factor *= 2 * PI * ey;
  for (int r = 0; r < 2; r++) { gcp[r] *= -factor; gcm[r] *= -factor; gphi[r] *= factor; gphi[r] *= vcp; }
  double nx = (by - ay) / len, ny = -(bx - ax) / len;
  if (nx * (cx[v[lc]] - ex) + ny * (cy[v[lc]] - ey) > 0) { nx = -nx; ny = -ny; }
  out[2 * t] = (gcp[0] + gphi[0]) * nx + (gcp[1] + gphi[1]) * ny;
  const double ratio = vcm / vcp;
  for (int r = 0; r < 2; r++) gphi[r] *= ratio;
  out[2 * t + 1] = (gcm[0] - gphi[0]) * nx + (gcm[1] - gphi[1]) * ny;
}
// DataWriter::writeData: centre, value and gradient at the centre of every element
template <int DEG>
__global__ void k_p2_cell_data(int nT, const int* __restrict__ tri, const double* __restrict__ cx, const double* __restrict__ cy,
                               const int* __restrict__ e2d, const double* __restrict__ u, double* __restrict__ out) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nT; e += gridDim.x * blockDim.x) {
    const int a = tri[3 * e], b = tri[3 * e + 1], c = tri[3 * e + 2];
    using E = PkElem<DEG>;
    constexpr int NL = E::NL;
    const typename E::Geo2 G = E::make_geo2(cx[a], cy[a], cx[b], cy[b], cx[c], cy[c]);
    const typename E::BasisAt B = E::basis_at(G, 1.0 / 3.0, 1.0 / 3.0);
    double val = 0, gr[2] = {0, 0};
    for (int k = 0; k < NL; k++) {
      const double uk = u[e2d[NL * e + k]];
      val += uk * B.phi[k];
      for (int d = 0; d < 2; d++) gr[d] += uk * B.g[k][d];
    }
    out[5 * (long)e] = (cx[a] + cx[b] + cx[c]) / 3.0; out[5 * (long)e + 1] = (cy[a] + cy[b] + cy[c]) / 3.0;
    out[5 * (long)e + 2] = val; out[5 * (long)e + 3] = gr[0]; out[5 * (long)e + 4] = gr[1];
  }
}
} // namespace

void p2_ion_flux(Ctx& c, const Vec& phi, const Vec& cp, const Vec& cm, double* ip, double* im) {
  P2Space& S = space(c);
  PNP_REQUIRE(c.constraints_built, PNP_E_ARG, "parameters not set");
  PNP_REQUIRE(phi.fields == 1 && cp.fields == 1 && cm.fields == 1, PNP_E_ARG, "ion flux: three 1-field vectors expected");
  const int ns = c.params.n_surfaces, nF = (int)S.h_bfaces.size();
  for (int s = 0; s < ns; s++) ip[s] = im[s] = 0.0;
  if (!nF) return;
  DBuf<double> d(2 * (size_t)nF);
  if (S.deg == 2) k_p2_ion_flux<2><<<(nF + 127) / 128, 128, 0, c.stream>>>(S.bfaces.p, nF, c.ctri.p, c.cx.p, c.cy.p, S.e2d.p, phi.d.p, cp.d.p,
                                                                         cm.d.p, c.params.cylindrical, c.params.PI, d.p);
  else k_p2_ion_flux<3><<<(nF + 127) / 128, 128, 0, c.stream>>>(S.bfaces.p, nF, c.ctri.p, c.cx.p, c.cy.p, S.e2d.p, phi.d.p, cp.d.p, cm.d.p,
                                                              c.params.cylindrical, c.params.PI, d.p);
  PNP_CHECK_LAUNCH(); c.launches++;
  const std::vector<double> h = d.to_host(c.stream);
  for (int t = 0; t < nF; t++) { // per-surface sums in the element loop's order
    const int pg = S.h_fphys[3 * (S.h_bfaces[t] >> 2) + (S.h_bfaces[t] & 3)];
    if (pg >= ns) continue;
    ip[pg] += h[2 * t]; im[pg] += h[2 * t + 1];
  }
}
void p2_write_cell_data(Ctx& c, const Vec& u, const std::string& filename) {
  P2Space& S = space(c);
  PNP_REQUIRE(u.fields == 1, PNP_E_ARG, "writeData: a 1-field vector expected");
  DBuf<double> d(5 * (size_t)S.nT);
  if (S.deg == 2) k_p2_cell_data<2><<<grid_for(S.nT, 128), 128, 0, c.stream>>>((int)S.nT, c.ctri.p, c.cx.p, c.cy.p, S.e2d.p, u.d.p, d.p);
  else k_p2_cell_data<3><<<grid_for(S.nT, 128), 128, 0, c.stream>>>((int)S.nT, c.ctri.p, c.cx.p, c.cy.p, S.e2d.p, u.d.p, d.p);
  PNP_CHECK_LAUNCH(); c.launches++;
  const std::vector<double> h = d.to_host(c.stream);
  std::FILE* f = std::fopen(filename.c_str(), "w");
  PNP_REQUIRE(f, PNP_E_CONFIG, "cannot open " + filename);
  for (long e = 0; e < S.nT; e++) std::fprintf(f, "%.5e %.5e\t%.5e\t%.5e %.5e\n", h[5 * e], h[5 * e + 1], h[5 * e + 2], h[5 * e + 3], h[5 * e + 4]);
  std::fclose(f);
}
// values at the mesh vertices (what VTKWriter::addVertexData samples), reference numbering
void p2_vertex_values(Ctx& c, const Vec& u, double* out) {
  P2Space& S = space(c);
  PNP_REQUIRE(u.fields == 1, PNP_E_ARG, "a 1-field vector expected");
  PNP_CUDA(cudaMemcpyAsync(out, u.d.p + S.voff, S.nv * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
  PNP_CUDA(cudaStreamSynchronize(c.stream));
}

// interpolate(BCExtension) for quadratic elements (dirichlet_bc.hh:54-123; interpolate: element loop, node by node, later
// elements overwrite earlier ones); host code like the boundary band of the P1 version: an O(N) set-up step
void p2_interpolate_bcext(Ctx& c, int comp, const Vec* pb, Vec& out) {
  P2Space& S = space(c);
  PNP_REQUIRE(comp >= 0 && comp < 3 && out.fields == 1 && (!pb || pb->fields == 1), PNP_E_ARG, "interpolate works on 1-field vectors");
  const std::vector<double> x = c.cx.to_host(c.stream), y = c.cy.to_host(c.stream);
  const std::vector<int> tri = c.ctri.to_host(c.stream);
  std::vector<double> pbh;
  if (pb) pbh = pb->d.to_host(c.stream);
  std::vector<double> u(S.nd, 0.0);
  const HostParams& s = c.params;
  auto on_line = [&](double px, double py, int e, int f) { // globalOnIntersection, :21-33
    const int a = tri[3 * e + FACE_V2[f][0]], b = tri[3 * e + FACE_V2[f][1]];
    double vx = x[b] - x[a], vy = y[b] - y[a];
    const double nrm = std::sqrt(vx * vx + vy * vy);
    vx /= nrm; vy /= nrm;
    const double dx = px - x[a], dy = py - y[a];
    const double t = dx * vx + dy * vy;
    const double ex = vx * t - dx, ey = vy * t - dy;
    return std::sqrt(ex * ex + ey * ey) < 1e-9;
  };
  auto sticky = [&](int g) { return s.surfaces.at(g).btype[2] == 0; }; // bctype() falls through to minusDiffusionBtype (:40-51)
  const int order[3] = {0, 2, 1};
  for (long e = 0; e < S.nT; e++) {
    const int a = tri[3 * e], b = tri[3 * e + 1], cv = tri[3 * e + 2];
    for (int i = 0; i < S.NL; i++) {
      const NodeInfo nt = node_info(S.deg, i);
      const double px = x[a] + (x[b] - x[a]) * nt.x + (x[cv] - x[a]) * nt.y;
      const double py = y[a] + (y[b] - y[a]) * nt.x + (y[cv] - y[a]) * nt.y;
      int pg = -1;
      for (int fi = 0; fi < 3; fi++) {
        const int f = order[fi];
        if (S.h_fphys[3 * e + f] >= 0) {
          if (on_line(px, py, (int)e, f)) if (pg == -1 || !sticky(pg)) pg = S.h_fphys[3 * e + f];
        } else {
          const int o = S.h_nbr[3 * e + f];
          for (int gi = 0; gi < 3; gi++) {
            const int f2 = order[gi];
            if (S.h_fphys[3 * o + f2] >= 0 && on_line(px, py, o, f2)) if (pg == -1 || !sticky(pg)) pg = S.h_fphys[3 * o + f2];
          }
        }
      }
      const int d = S.h_e2d[S.NL * e + i];
      if (pg > -1 && s.surfaces.at(pg).btype[comp] == 0) { u[d] = s.surfaces[pg].dval[comp]; continue; }
      const double yv = pb ? pbh[d] : 0.0;
      u[d] = comp == 0 ? yv : (comp == 1 ? s.c0 * std::exp(-yv) : s.c0 * std::exp(+yv));
    }
  }
  out.d.upload(u.data(), S.nd, c.stream);
  PNP_CUDA(cudaStreamSynchronize(c.stream));
}

} // namespace pnp
