// harness.cpp -- TEST INFRASTRUCTURE: runs the product's __host__ __device__ per-item functions
// (dune_pnp_b200/csrc/pnp_elem.cuh, pnp_star.cuh, pnp_setup_algos.cuh) in plain host loops so their
// logic can be compared with the CPU oracle on a machine without a GPU.  It is compiled by
// tests/conftest.py with g++ -ffp-contract=off and is never linked into libpnp_b200.so; the
// product has no CPU path.
#include <algorithm>
#include <cstring>
#include <numeric>
#include <vector>

#include "../../dune_pnp_b200/csrc/pnp_star.cuh"
#include "../../dune_pnp_b200/csrc/pnp_sweep.cuh"
#include "../../dune_pnp_b200/csrc/pnp_elem_p2.cuh"

using namespace pnp;

namespace {
struct Star {
  long nv = 0, nT = 0, n_own = 0;
  std::vector<int> rp, int2ext, ext2int;
  std::vector<unsigned> adj;
  std::vector<XY> xy;
  std::vector<unsigned char> dmask;
  std::vector<BFace> bfaces;
  std::vector<int> bv, bv_ptr, bv_items;
  std::vector<double> surf_flux;
  std::vector<unsigned char> surf_dir;
  StarView view() const { return StarView{rp.data(), adj.data(), xy.data(), dmask.data(), (int)n_own}; }
};
} // namespace

// phys[5] = {PI, l_b, c0, valency, cylindrical}; vectors in INTERNAL blocked layout
static PhysParams mkphys(const double* p) { PhysParams P; P.PI = p[0]; P.l_b = p[1]; P.c0 = p[2]; P.valency = p[3]; P.cylindrical = (int)p[4]; return P; }

template <int OP> static void residual_t(Star* S, const PhysParams& P, const double* u, const double* a0, const double* a1,
                                         int comp0, double* r) {
  constexpr int F = OpTraits<OP>::F;
  const StarView M = S->view();
  for (int v = 0; v < M.nv; v++) {
    double out[F];
    residual_row<OP>(M, P, u, a0, a1, v, out);
    const unsigned db = dir_bits<OP>(M, v, comp0);
    for (int k = 0; k < F; k++) r[(long)F * v + k] = ((db >> k) & 1u) ? 0.0 : out[k];
  }
  if (OP == OP_PB || OP == OP_POISSON || OP == OP_PNP)
    for (size_t t = 0; t < S->bv.size(); t++) {
      double acc[3];
      boundary_vertex_sum(M, P, S->bfaces.data(), S->bv_items.data(), S->bv_ptr[t], S->bv_ptr[t + 1], S->surf_flux.data(),
                          S->surf_dir.data(), F, comp0, acc);
      const int v = S->bv[t];
      for (int k = 0; k < F; k++) {
        const int comp = F == 3 ? k : comp0;
        if (!((M.dmask[v] >> comp) & 1u)) r[(long)F * v + k] += acc[k];
      }
    }
}
template <int OP> static void jacobian_t(Star* S, const PhysParams& P, const double* u, const double* a0, const double* a1,
                                         int comp0, int mode, double eps, double* vals) {
  const StarView M = S->view();
  const long stride = (long)S->adj.size();
  for (int v = 0; v < M.nv; v++) {
    if (mode == JAC_FD_FAITHFUL) jacobian_row<OP, JAC_FD_FAITHFUL>(M, P, u, a0, a1, eps, comp0, v, vals, stride);
    else jacobian_row<OP, JAC_ANALYTIC>(M, P, u, a0, a1, eps, comp0, v, vals, stride);
  }
}
extern "C" {

// refinement with the product's per-item functions; outputs must be sized by the caller:
// x,y: nv+nE (call with out arrays NULL to get nE)
long hh_refine(long nv, const double* x, const double* y, long nT, const int* tri, long nB, const int* ba, const int* bb,
               const int* bphys, double* ox, double* oy, int* otri, int* oba, int* obb, int* obphys) {
  std::vector<uint64_t> keys(3 * nT);
  for (long i = 0; i < 3 * nT; i++) keys[i] = tri_edge_key(tri, i);
  std::sort(keys.begin(), keys.end());
  keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
  const long nE = (long)keys.size();
  if (!ox) return nE;
  std::copy(x, x + nv, ox); std::copy(y, y + nv, oy);
  for (long k = 0; k < nE; k++) {
    int a = (int)(keys[k] >> 32), b = (int)(keys[k] & 0xffffffffu);
    ox[nv + k] = 0.5 * (x[a] + x[b]); oy[nv + k] = 0.5 * (y[a] + y[b]);
  }
  for (long t = 0; t < nT; t++) refine_children(tri, t, keys.data(), nE, nv, otri + 12 * t);
  for (long s = 0; s < nB; s++) {
    int m = (int)(nv + lower_bound_u64(keys.data(), nE, edge_key(ba[s], bb[s])));
    oba[2 * s] = ba[s]; obb[2 * s] = m; obphys[2 * s] = bphys[s];
    oba[2 * s + 1] = m; obb[2 * s + 1] = bb[s]; obphys[2 * s + 1] = bphys[s];
  }
  return nE;
}

// builds the star exactly as mesh_finalize()/constraints_build() do (same per-item functions)
// surf_btype[ns][3], surf_flux[ns][3]; returns NULL on a mesh error (code in *err)
void* hh_star_build(long nv, const double* x, const double* y, long nT, const int* tri, long nB, const int* ba,
                    const int* bb, const int* bphys, int renumber, int ns, const int* surf_btype, const double* surf_flux,
                    long n_own, int* err) {
  Star* S = new Star; S->nv = nv; S->nT = nT; S->n_own = n_own; *err = 0;
  const long no = n_own;
  S->int2ext.resize(nv); S->ext2int.resize(nv);
  if (renumber) {
    std::vector<long> first(nv, 0x7fffffff);
    for (long i = 0; i < 3 * nT; i++) first[tri[i]] = std::min(first[tri[i]], i);
    std::iota(S->int2ext.begin(), S->int2ext.end(), 0);
    std::stable_sort(S->int2ext.begin(), S->int2ext.begin() + no, [&](int a, int b) { return first[a] < first[b]; });
  } else std::iota(S->int2ext.begin(), S->int2ext.end(), 0);
  for (long i = 0; i < nv; i++) S->ext2int[S->int2ext[i]] = (int)i;
  S->xy.resize(nv);
  for (long i = 0; i < nv; i++) S->xy[i] = XY{x[S->int2ext[i]], y[S->int2ext[i]]};
  const long nrec = 3 * nT;
  std::vector<uint64_t> keys(nrec); std::vector<unsigned> pay(nrec);
  for (long t = 0; t < nT; t++)
    if (!corner_records(tri, t, S->ext2int.data(), x, y, &keys[3 * t], &pay[3 * t])) *err = 1;
  std::vector<long> order(nrec);
  std::iota(order.begin(), order.end(), 0);
  std::sort(order.begin(), order.end(), [&](long a, long b) { return keys[a] < keys[b]; });
  std::vector<uint64_t> sk(nrec); std::vector<unsigned> sp(nrec);
  for (long i = 0; i < nrec; i++) { sk[i] = keys[order[i]]; sp[i] = pay[order[i]]; }
  std::vector<int> start(no + 1);
  for (long v = 0; v <= no; v++) start[v] = (int)lower_bound_u64(sk.data(), nrec, (uint64_t)v << 32);
  S->rp.assign(no + 1, 0);
  for (long v = 0; v < no; v++) {
    int open, bad;
    fan_start(sk.data(), sp.data(), start[v], start[v + 1], &open, &bad);
    if (bad) *err = 2;
    S->rp[v + 1] = S->rp[v] + 1 + (start[v + 1] - start[v]) + open;
  }
  S->adj.assign(S->rp[no], 0);
  for (long v = 0; v < no; v++)
    if (!ring_fill(sk.data(), sp.data(), start[v], start[v + 1], (int)v, &S->adj[S->rp[v]])) *err = 2;
  if (*err) { delete S; return nullptr; }
  long nopen = 0;
  for (long v = 0; v < no; v++) if (!(S->adj[S->rp[v + 1] - 1] & STAR_HAS_TRI)) nopen++;
  S->bfaces.resize(nB);
  for (long s = 0; s < nB; s++) { // mirrors k_bfaces
    BFace& bf = S->bfaces[s];
    const int a = S->ext2int[ba[s]], b = S->ext2int[bb[s]];
    bool ok;
    if (a < no) ok = boundary_face_of(S->rp.data(), S->adj.data(), a, b, &bf);
    else if (b < no) ok = boundary_face_of(S->rp.data(), S->adj.data(), b, a, &bf);
    else { bf.v[0] = bf.v[1] = bf.v[2] = -1; bf.f = 0; ok = no < nv; }
    if (!ok) *err = 3;
    bf.a = a; bf.b = b; bf.phys = bphys[s]; bf.seg = (int)s;
  }
  if (no == nv && nopen != nB && !*err) *err = 4;
  if (*err) { delete S; return nullptr; }
  // constraints + boundary incidence lists (mirrors constraints_build)
  S->dmask.assign(nv, 0);
  S->surf_flux.assign(surf_flux, surf_flux + 3 * ns);
  S->surf_dir.assign(ns, 0);
  for (int i = 0; i < ns; i++) for (int k = 0; k < 3; k++) if (surf_btype[3 * i + k] == 0) S->surf_dir[i] |= (1u << k);
  std::vector<std::vector<int>> items(nv);
  for (long s = 0; s < nB; s++) {
    const BFace& b = S->bfaces[s];
    S->dmask[b.a] |= S->surf_dir[b.phys]; S->dmask[b.b] |= S->surf_dir[b.phys];
    if (b.v[0] < 0) continue;
    for (int r = 0; r < 3; r++) if (b.v[r] < no) items[b.v[r]].push_back((int)s * 4 + r);
  }
  S->bv_ptr.push_back(0);
  for (long v = 0; v < nv; v++) if (!items[v].empty()) {
    S->bv.push_back((int)v);
    std::sort(items[v].begin(), items[v].end());
    S->bv_items.insert(S->bv_items.end(), items[v].begin(), items[v].end());
    S->bv_ptr.push_back((int)S->bv_items.size());
  }
  return S;
}
void hh_star_free(void* h) { delete (Star*)h; }
long hh_star_nslots(void* h) { return (long)((Star*)h)->adj.size(); }
void hh_star_get(void* h, int* rp, unsigned* adj, int* int2ext, unsigned char* dmask) {
  Star* S = (Star*)h;
  std::copy(S->rp.begin(), S->rp.end(), rp); std::copy(S->adj.begin(), S->adj.end(), adj);
  std::copy(S->int2ext.begin(), S->int2ext.end(), int2ext); std::copy(S->dmask.begin(), S->dmask.end(), dmask);
}

void hh_residual(void* h, int op, const double* phys, const double* u, const double* a0, const double* a1, int comp0,
                 double* r) {
  Star* S = (Star*)h; const PhysParams P = mkphys(phys);
  switch (op) {
    case OP_PB: residual_t<OP_PB>(S, P, u, a0, a1, comp0, r); break;
    case OP_POISSON: residual_t<OP_POISSON>(S, P, u, a0, a1, comp0, r); break;
    case OP_DIFFUSION: residual_t<OP_DIFFUSION>(S, P, u, a0, a1, comp0, r); break;
    case OP_MASS: residual_t<OP_MASS>(S, P, u, a0, a1, comp0, r); break;
    case OP_PNP: residual_t<OP_PNP>(S, P, u, a0, a1, comp0, r); break;
  }
}
void hh_jacobian(void* h, int op, const double* phys, const double* u, const double* a0, const double* a1, int comp0,
                 int mode, double eps, double* vals) {
  Star* S = (Star*)h; const PhysParams P = mkphys(phys);
  switch (op) {
    case OP_PB: jacobian_t<OP_PB>(S, P, u, a0, a1, comp0, mode, eps, vals); break;
    case OP_POISSON: jacobian_t<OP_POISSON>(S, P, u, a0, a1, comp0, mode, eps, vals); break;
    case OP_DIFFUSION: jacobian_t<OP_DIFFUSION>(S, P, u, a0, a1, comp0, mode, eps, vals); break;
    case OP_MASS: jacobian_t<OP_MASS>(S, P, u, a0, a1, comp0, mode, eps, vals); break;
    case OP_PNP: jacobian_t<OP_PNP>(S, P, u, a0, a1, comp0, mode, eps, vals); break;
  }
}

} // extern "C"

// ---- level-scheduled sweeps (pnp_sweep.cuh): the kernels of pnp_precond.cu emulated level by level ----
// Levels by fixpoint relaxation (as the device does); returns the number of levels.  lev: F*n_own ints.
template <int F> static int sweep_levels_t(Star* S, bool full, int* lev) {
  const SweepView V{S->rp.data(), S->adj.data(), S->int2ext.data(), (int)S->n_own};
  std::fill(lev, lev + F * S->n_own, 0);
  for (bool changed = true; changed;) {
    changed = false;
    for (int v = (int)S->n_own - 1; v >= 0; v--) // deliberately against the sweep order: worst case for the relaxation
      for (int f = 0; f < F; f++) {
        const int l = sweep_level_relax<F>(V, lev, v, f, full);
        if (l != lev[F * v + f]) { lev[F * v + f] = l; changed = true; }
      }
  }
  return 1 + *std::max_element(lev, lev + F * S->n_own);
}
// dofs grouped by level; inside a level in DESCENDING internal index (any order must give the same result)
static std::vector<std::vector<int>> level_sets(const int* lev, long n, int nlev) {
  std::vector<std::vector<int>> L(nlev);
  for (long i = n - 1; i >= 0; i--) L[lev[i]].push_back((int)i);
  return L;
}
extern "C" {
int hh_sweep_levels(void* h, int F, int full, int* lev) {
  return F == 1 ? sweep_levels_t<1>((Star*)h, full, lev) : sweep_levels_t<3>((Star*)h, full, lev);
}
// x = SSOR^steps applied to d (x starts at 0), vals: NP planes (internal layout)
void hh_ssor_apply(void* h, int F, const double* vals, int steps, const double* d, double* x) {
  Star* S = (Star*)h;
  const SweepView V{S->rp.data(), S->adj.data(), S->int2ext.data(), (int)S->n_own};
  const long n = (long)F * S->n_own, stride = (long)S->adj.size();
  std::vector<int> lev(n);
  const int nlev = hh_sweep_levels(h, F, 0, lev.data());
  const auto L = level_sets(lev.data(), n, nlev);
  std::fill(x, x + n, 0.0);
  for (int s = 0; s < steps; s++) {
    for (int l = 0; l < nlev; l++)
      for (int i : L[l]) { if (F == 1) gs_update<1>(V, vals, stride, d, x, i, 0); else gs_update<7>(V, vals, stride, d, x, i / 3, i % 3); }
    for (int l = nlev - 1; l >= 0; l--)
      for (int i : L[l]) { if (F == 1) gs_update<1>(V, vals, stride, d, x, i, 0); else gs_update<7>(V, vals, stride, d, x, i / 3, i % 3); }
  }
}
// x = (LU)^-1 d with the ILU(0) factor of the matrix given as F*F planes (overwritten by the factor)
void hh_ilu0_apply(void* h, int F, double* lu, const double* d, double* x) {
  Star* S = (Star*)h;
  const SweepView V{S->rp.data(), S->adj.data(), S->int2ext.data(), (int)S->n_own};
  const long n = (long)F * S->n_own, stride = (long)S->adj.size();
  std::vector<int> lev(n);
  const int nlev = hh_sweep_levels(h, F, 1, lev.data());
  const auto L = level_sets(lev.data(), n, nlev);
  for (int l = 0; l < nlev; l++)
    for (int i : L[l]) { if (F == 1) ilu0_row<1>(V, lu, stride, i, 0); else ilu0_row<3>(V, lu, stride, i / 3, i % 3); }
  for (int l = 0; l < nlev; l++)
    for (int i : L[l]) { if (F == 1) ilu0_forward<1>(V, lu, stride, d, x, i, 0); else ilu0_forward<3>(V, lu, stride, d, x, i / 3, i % 3); }
  for (int l = nlev - 1; l >= 0; l--)
    for (int i : L[l]) { if (F == 1) ilu0_backward<1>(V, lu, stride, x, i, 0); else ilu0_backward<3>(V, lu, stride, x, i / 3, i % 3); }
}
} // extern "C"

// the device's pnp_sinh compiled for the host (-ffp-contract=off): must equal the oracle's sinh_shared bit for bit
extern "C" void hh_sinh(int n, const double* x, double* y) { for (int i = 0; i < n; i++) y[i] = pnp::pnp_sinh(x[i]); }

// ---- quadratic / cubic elements: the product's element functions (pnp_elem_p2.cuh) on ONE triangle --------------------------
// xy[6] = vertex coordinates; xl = local coefficients (field-major, NL per field); caux = 2 x NL coefficient-field values;
// face f (DUNE order) is a boundary face when fflag[f] != 0 with fluxes fflux[3*f + k] and Dirichlet bits fdir[f];
// out_r (NL*F) = alpha_volume + alpha_boundary in intersection order 0, 2, 1; out_A (n*n) = the element matrix (mode 0: FD, 1: exact).
namespace {
template <int DEG, int OP, int NQ>
void pk_element_t(const double* xy, const PhysParams& P, const double* xl_in, const double* caux_in, const int* fflag, const double* fflux,
                  const int* fdir, int comp0, int line_pts, int mode, double eps, double* out_r, double* out_A) {
  using E = PkElem<DEG>;
  constexpr int F = OpTraits<OP>::F, NL = E::NL, n = NL * F;
  const typename E::Geo2 G = E::make_geo2(xy[0], xy[1], xy[2], xy[3], xy[4], xy[5]);
  double xl[n], caux[2][NL], rl[n];
  for (int i = 0; i < n; i++) { xl[i] = xl_in[i]; rl[i] = 0.0; }
  for (int a = 0; a < 2; a++) for (int i = 0; i < NL; i++) caux[a][i] = caux_in[a * NL + i];
  E::template alpha_volume<OP, NQ>(G, P, xl, caux, rl);
  if (OP == OP_PB || OP == OP_POISSON || OP == OP_PNP) {
    const int order[3] = {0, 2, 1};
    for (int fi = 0; fi < 3; fi++) {
      const int f = order[fi];
      if (!fflag[f]) continue;
      const int la = f == 2 ? 1 : 0, lb = f == 0 ? 1 : 2;
      double j[3]; bool skip[3];
      for (int k = 0; k < F; k++) { const int comp = F == 3 ? k : comp0; j[k] = fflux[3 * f + comp]; skip[k] = (fdir[f] >> comp) & 1; }
      E::alpha_boundary(f, xy[2 * la], xy[2 * la + 1], xy[2 * lb], xy[2 * lb + 1], F, j, skip, P, line_pts, rl);
    }
  }
  for (int i = 0; i < n; i++) out_r[i] = rl[i];
  for (int i = 0; i < n * n; i++) out_A[i] = 0.0;
  if (mode == 0) E::template jacobian_fd<OP, NQ>(G, P, xl, caux, eps, out_A);
  else E::template jacobian_exact<OP, NQ>(G, P, xl, caux, out_A);
}
template <int DEG, int OP, class... A> void pk_nq(int intorder, A... a) {
  if (intorder == 5) pk_element_t<DEG, OP, 7>(a...); else pk_element_t<DEG, OP, OpTraits<OP>::NQ>(a...);
}
template <int DEG, class... A> void pk_op(int op, int intorder, A... a) {
  switch (op) {
    case OP_PB: pk_nq<DEG, OP_PB>(intorder, a...); break;
    case OP_POISSON: pk_nq<DEG, OP_POISSON>(intorder, a...); break;
    case OP_DIFFUSION: pk_nq<DEG, OP_DIFFUSION>(intorder, a...); break;
    case OP_MASS: pk_nq<DEG, OP_MASS>(intorder, a...); break;
    default: pk_nq<DEG, OP_PNP>(intorder, a...); break;
  }
}
} // namespace
extern "C" void hh_pk_element(int degree, int op, int intorder, const double* xy, const double* phys, const double* xl, const double* caux,
                              const int* fflag, const double* fflux, const int* fdir, int comp0, int mode, double eps, double* out_r,
                              double* out_A) {
  const PhysParams P = mkphys(phys);
  const int line_pts = intorder == 5 ? 3 : 2;
  if (degree == 2) pk_op<2>(op, intorder, xy, P, xl, caux, fflag, fflux, fdir, comp0, line_pts, mode, eps, out_r, out_A);
  else pk_op<3>(op, intorder, xy, P, xl, caux, fflag, fflux, fdir, comp0, line_pts, mode, eps, out_r, out_A);
}
// local node -> (kind, sub, idx, x, y) of the product's node table
extern "C" void hh_pk_node(int degree, int n, int* key, double* xy) {
  if (degree == 2) { PkElem<2>::node_key(n, key[0], key[1], key[2]); xy[0] = PkElem<2>::node_x(n); xy[1] = PkElem<2>::node_y(n); }
  else { PkElem<3>::node_key(n, key[0], key[1], key[2]); xy[0] = PkElem<3>::node_x(n); xy[1] = PkElem<3>::node_y(n); }
}
