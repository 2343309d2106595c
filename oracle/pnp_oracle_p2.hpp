// pnp_oracle_p2.hpp -- TEST INFRASTRUCTURE: the reference path with quadratic elements (-DPDEGREE=2,
// /root/reference/src/Makefile.am:57-110; instationary_pnp_from_pb_md.hh:26-28,125: Pk2DLocalFiniteElementMap<GV,D,R,2>).
// Same local operators (the alpha_volume / alpha_boundary bodies of pnp_oracle.hpp, which are written for any
// lfsu.size()), 6 local dofs per field.  PARITY UNPINNED like the rest of the oracle; the upstream rules assumed here:
//   * Pk2DLocalBasis<D,R,2> [UPSTREAM, from memory, SURVEY A.6]: Lagrange nodes in lexicographic order (0,0),(1/2,0),(1,0),
//     (0,1/2),(1/2,1/2),(0,1) = vertex0, edge0, vertex1, edge1, edge2, vertex2; value of node (i,j):
//       prod_{a<i} (2x-a)/(i-a) * prod_{b<j} (2y-b)/(j-b) * prod_{g=i+j+1..2} (g-2x-2y)/(g-i-j)
//   * dof numbering: codim by codim, edges first [0,nE), then vertices [nE, nE+nv); the edge index is the rank
//     of (min vertex, max vertex) -- the refinement rule's edge order (the real UGGrid leaf edge index is not pinned, H2);
//     composite spaces lexicographic: g = field * (nE + nv) + scalar dof;
//   * constraints (A.5): a Dirichlet boundary face constrains its two end vertices AND its edge dof;
//   * interpolate (A.6): element loop, u[g(i)] = f(element, node_i), later elements overwrite earlier ones.
#pragma once
#include "pnp_oracle.hpp"

namespace pnpo {
template <int DEG> struct Pk {
  static_assert(DEG == 2 || DEG == 3, "Pk2DLocalFiniteElementMap degrees built here: 2, 3");

static constexpr int NL = (DEG + 1) * (DEG + 2) / 2;
// Lagrange nodes: lattice points (i, j)/DEG, i + j <= DEG, lexicographic (j outer, i inner).  What a node sits on: kind 0 =
// vertex (sub = local vertex), 1 = edge (sub = local edge, idx = position counted from the edge's FIRST local vertex),
// 2 = element interior (degree 3: the bubble at (1/3, 1/3)).
struct NodeTab { double x[NL], y[NL]; int kind[NL], sub[NL], idx[NL]; };
static const NodeTab& nodes() {
  static const NodeTab T = [] {
    NodeTab t; int n = 0;
    for (int j = 0; j <= DEG; j++) for (int i = 0; i <= DEG - j; i++, n++) {
      t.x[n] = (1.0 * i) / DEG; t.y[n] = (1.0 * j) / DEG; t.idx[n] = 0;
      if (i == 0 && j == 0) { t.kind[n] = 0; t.sub[n] = 0; }
      else if (i == DEG) { t.kind[n] = 0; t.sub[n] = 1; }
      else if (j == DEG) { t.kind[n] = 0; t.sub[n] = 2; }
      else if (j == 0) { t.kind[n] = 1; t.sub[n] = 0; t.idx[n] = i - 1; }      // edge0 = (v0, v1)
      else if (i == 0) { t.kind[n] = 1; t.sub[n] = 1; t.idx[n] = j - 1; }      // edge1 = (v0, v2)
      else if (i + j == DEG) { t.kind[n] = 1; t.sub[n] = 2; t.idx[n] = j - 1; } // edge2 = (v1, v2)
      else { t.kind[n] = 2; t.sub[n] = 0; }
    }
    return t;
  }();
  return T;
}

// Pk2DLocalBasis<D,R,k>::evaluateFunction / evaluateJacobian for any k [UPSTREAM dune-localfunctions 2.2 pk2dlocalbasis.hh, from
// memory]: pos[i] = i/k; node (i,j): prod_{a<i} (x-pos[a])/(pos[i]-pos[a]) * prod_{b<j} (y-pos[b])/(pos[j]-pos[b]) *
// prod_{g=i+j+1..k} (pos[g]-x-y)/(pos[g]-pos[i]-pos[j]); the derivatives by the product rule, factor by factor
static void basis_generic(double x, double y, double* phi) {
  double pos[DEG + 1];
  for (int i = 0; i <= DEG; i++) pos[i] = (1.0 * i) / DEG;
  int n = 0;
  for (int j = 0; j <= DEG; j++) for (int i = 0; i <= DEG - j; i++) {
    double out = 1.0;
    for (int a = 0; a < i; a++) out *= (x - pos[a]) / (pos[i] - pos[a]);
    for (int b = 0; b < j; b++) out *= (y - pos[b]) / (pos[j] - pos[b]);
    for (int g = i + j + 1; g <= DEG; g++) out *= (pos[g] - x - y) / (pos[g] - pos[i] - pos[j]);
    phi[n++] = out;
  }
}
static void basis_grad_generic(double x, double y, double (*gr)[2]) {
  double pos[DEG + 1];
  for (int i = 0; i <= DEG; i++) pos[i] = (1.0 * i) / DEG;
  int n = 0;
  for (int j = 0; j <= DEG; j++) for (int i = 0; i <= DEG - j; i++, n++) {
    for (int dir = 0; dir < 2; dir++) {
      // the direction's own factors are (dir == 0 ? the alpha product in x : the beta product in y); the other product is a factor
      const int own = dir == 0 ? i : j, oth = dir == 0 ? j : i;
      const double z = dir == 0 ? x : y, w = dir == 0 ? y : x;
      double factor = 1.0, sum = 0.0;
      for (int b = 0; b < oth; b++) factor *= (w - pos[b]) / (pos[oth] - pos[b]);
      for (int a = 0; a < own; a++) {
        double product = factor;
        for (int al = 0; al < own; al++)
          if (al == a) product *= 1.0 / (pos[own] - pos[al]);
          else product *= (z - pos[al]) / (pos[own] - pos[al]);
        for (int g = i + j + 1; g <= DEG; g++) product *= (pos[g] - x - y) / (pos[g] - pos[i] - pos[j]);
        sum += product;
      }
      for (int c = i + j + 1; c <= DEG; c++) {
        double product = factor;
        for (int al = 0; al < own; al++) product *= (z - pos[al]) / (pos[own] - pos[al]);
        for (int g = i + j + 1; g <= DEG; g++)
          if (g == c) product *= -1.0 / (pos[g] - pos[i] - pos[j]);
          else product *= (pos[g] - x - y) / (pos[g] - pos[i] - pos[j]);
        sum += product;
      }
      gr[n][dir] = sum;
    }
  }
}
static void basis(double x, double y, double* phi) {
  if (DEG != 2) { basis_generic(x, y, phi); return; }
  const double s = 2 * x + 2 * y;
  phi[0] = ((1 - s) / 1) * ((2 - s) / 2);
  phi[1] = (2 * x) * ((2 - s) / 1);
  phi[2] = (2 * x) * ((2 * x - 1) / 2);
  phi[3] = (2 * y) * ((2 - s) / 1);
  phi[4] = (2 * x) * (2 * y);
  phi[5] = (2 * y) * ((2 * y - 1) / 2);
}
// reference gradients (d/dx, d/dy) of the same products
static void basis_grad(double x, double y, double (*g)[2]) {
  if (DEG != 2) { basis_grad_generic(x, y, g); return; }
  const double s = 2 * x + 2 * y;
  // phi0 = (1-s)(2-s)/2 : d/ds = (2s-3)/2, ds/dx = ds/dy = 2
  g[0][0] = (2 * s - 3); g[0][1] = (2 * s - 3);
  // phi1 = 2x(2-s)
  g[1][0] = 2 * (2 - s) - 4 * x; g[1][1] = -4 * x;
  // phi2 = x(2x-1)*... = 2x(2x-1)/2 = x(2x-1)
  g[2][0] = 4 * x - 1; g[2][1] = 0.0;
  // phi3 = 2y(2-s)
  g[3][0] = -4 * y; g[3][1] = 2 * (2 - s) - 4 * y;
  // phi4 = 4xy
  g[4][0] = 4 * y; g[4][1] = 4 * x;
  // phi5 = y(2y-1)
  g[5][0] = 0.0; g[5][1] = 4 * y - 1;
}

struct Space2 {
  const Mesh* m = nullptr;
  const Sysparams* s = nullptr;
  int fields = 1, comp0 = 0;
  int nE = 0, nd = 0;          // edges; scalar dofs = [nT bubbles (degree 3)] + (DEG-1) per edge + nv
  int eoff = 0, voff = 0;      // first edge dof, first vertex dof (codim by codim -- elements, edges, vertices)
  std::vector<int> tedge;      // 3*nT: global edge of local edge f = (FACE_V[f][0], FACE_V[f][1])
  std::vector<int> eva, evb;   // edge -> end vertices (min, max)
  std::vector<char> dirichlet; // per dof, lexicographic [field][scalar dof]
  int N() const { return fields * nd; }
  int sdof(int e, int i) const { // scalar dof of local node i of element e
    const NodeTab& T = nodes();
    if (T.kind[i] == 2) return e;
    if (T.kind[i] == 0) return voff + m->tri[3 * e + T.sub[i]];
    // edge dofs are counted from the end vertex with the SMALLER global index (Pk2DLocalFiniteElementMap picks the local
    // coefficients variant by comparing the global vertex indices, so that neighbours agree on the order) [UPSTREAM]
    int idx = T.idx[i];
    const int f = T.sub[i];
    if (DEG == 3 && m->tri[3 * e + FACE_V[f][0]] > m->tri[3 * e + FACE_V[f][1]]) idx = 1 - idx;
    return eoff + (DEG - 1) * tedge[3 * e + f] + idx;
  }
  int gdof(int field, int sd) const { return field * nd + sd; }
};

static Space2 make_space2(const Mesh& m, const Sysparams& s, int fields, int comp0 = 0) {
  Space2 sp; sp.m = &m; sp.s = &s; sp.fields = fields; sp.comp0 = comp0;
  auto key = [](int a, int b) { return ((uint64_t)std::min(a, b) << 32) | (uint64_t)std::max(a, b); };
  std::vector<uint64_t> keys; keys.reserve(3 * (size_t)m.nT);
  for (int e = 0; e < m.nT; e++) for (int f = 0; f < 3; f++) keys.push_back(key(m.tri[3 * e + FACE_V[f][0]], m.tri[3 * e + FACE_V[f][1]]));
  std::vector<uint64_t> uk(keys);
  std::sort(uk.begin(), uk.end()); uk.erase(std::unique(uk.begin(), uk.end()), uk.end());
  sp.nE = (int)uk.size();
  sp.eoff = DEG == 3 ? m.nT : 0; sp.voff = sp.eoff + (DEG - 1) * sp.nE; sp.nd = sp.voff + m.nv;
  sp.tedge.resize(3 * (size_t)m.nT);
  for (size_t i = 0; i < keys.size(); i++) sp.tedge[i] = (int)(std::lower_bound(uk.begin(), uk.end(), keys[i]) - uk.begin());
  sp.eva.resize(sp.nE); sp.evb.resize(sp.nE);
  for (int k = 0; k < sp.nE; k++) { sp.eva[k] = (int)(uk[k] >> 32); sp.evb[k] = (int)(uk[k] & 0xffffffffu); }
  sp.dirichlet.assign(sp.N(), 0);
  for (int e = 0; e < m.nT; e++)
    for (int f = 0; f < 3; f++) {
      const int seg = m.fseg[3 * e + f];
      if (seg < 0) continue;
      const Surface& sf = s.surfaces.at(m.bphys[seg]);
      for (int k = 0; k < fields; k++) {
        const int comp = fields == 3 ? k : comp0;
        if (sf.btype(comp) != 0) continue;
        for (int l = 0; l < 2; l++) sp.dirichlet[sp.gdof(k, sp.voff + m.tri[3 * e + FACE_V[f][l]])] = 1;
        for (int l = 0; l < DEG - 1; l++) sp.dirichlet[sp.gdof(k, sp.eoff + (DEG - 1) * sp.tedge[3 * e + f] + l)] = 1;
      }
    }
  return sp;
}

// FullVolumePattern: all local pairs of every element; constrained links dropped, constrained rows keep the diagonal
static CSR make_pattern2(const Space2& sp) {
  const Mesh& m = *sp.m;
  const int nd = sp.nd, nf = sp.fields, N = sp.N();
  std::vector<int> cnt(nd + 1, 0);
  for (int e = 0; e < m.nT; e++) for (int i = 0; i < NL; i++) cnt[sp.sdof(e, i) + 1] += NL;
  for (int d = 0; d < nd; d++) cnt[d + 1] += cnt[d];
  std::vector<int> raw(cnt[nd]), fill(cnt.begin(), cnt.end() - 1);
  for (int e = 0; e < m.nT; e++) for (int i = 0; i < NL; i++) for (int j = 0; j < NL; j++) raw[fill[sp.sdof(e, i)]++] = sp.sdof(e, j);
  std::vector<int> nptr(nd + 1, 0), nbr;
  for (int d = 0; d < nd; d++) {
    auto b = raw.begin() + cnt[d], e = raw.begin() + cnt[d + 1];
    std::sort(b, e); e = std::unique(b, e);
    nbr.insert(nbr.end(), b, e); nptr[d + 1] = (int)nbr.size();
  }
  CSR A; A.n = N; A.rowptr.assign(N + 1, 0);
  for (int ki = 0; ki < nf; ki++) for (int d = 0; d < nd; d++) {
    const int gi = sp.gdof(ki, d);
    int len = 0;
    if (sp.dirichlet[gi]) len = 1;
    else for (int kj = 0; kj < nf; kj++) for (int t = nptr[d]; t < nptr[d + 1]; t++) len += !sp.dirichlet[sp.gdof(kj, nbr[t])];
    A.rowptr[gi + 1] = len;
  }
  for (int r = 0; r < N; r++) A.rowptr[r + 1] += A.rowptr[r];
  A.col.resize(A.rowptr[N]);
  for (int ki = 0; ki < nf; ki++) for (int d = 0; d < nd; d++) {
    const int gi = sp.gdof(ki, d);
    int o = A.rowptr[gi];
    if (sp.dirichlet[gi]) { A.col[o] = gi; continue; }
    for (int kj = 0; kj < nf; kj++) for (int t = nptr[d]; t < nptr[d + 1]; t++) {
      const int gj = sp.gdof(kj, nbr[t]);
      if (!sp.dirichlet[gj]) A.col[o++] = gj;
    }
  }
  A.val.assign(A.col.size(), 0.0);
  return A;
}

// basis values and transformed gradients at a reference point
struct BasisAt { double phi[NL], g[NL][2]; };
static BasisAt basis_at(const ElemGeo& G, double x, double y) {
  BasisAt B;
  basis(x, y, B.phi);
  double gh[NL][2];
  basis_grad(x, y, gh);
  for (int i = 0; i < NL; i++)
    for (int r = 0; r < 2; r++) { // FieldMatrix::mv
      double v = 0.0;
      v += G.jit[r][0] * gh[i][0];
      v += G.jit[r][1] * gh[i][1];
      B.g[i][r] = v;
    }
  return B;
}

// alpha_volume with 6 local dofs per field (the operator bodies of pnp_oracle.hpp, lfsu.size() = 6); the coefficient fields
// of the Poisson / diffusion operators are P2 functions too: caux[a][i] = their local coefficients
static void alpha_volume2(const OpCtx& c, int e, const double* xl, const double (*caux)[NL], double* rl) {
  const Mesh& m = *c.m; const Sysparams& s = *c.s;
  const ElemGeo G = elem_geo(m, e);
  const double PI = s.PI;
  auto dot = [](const double* a, const double* b) { double r = 0.0; r += a[0] * b[0]; r += a[1] * b[1]; return r; };
  for (const QP& q : tri_rule(c.order())) {
    const BasisAt B = basis_at(G, q.xi0, q.xi1);
    const double gy = G.y0 + (G.y1 - G.y0) * q.xi0 + (G.y2 - G.y0) * q.xi1;
    double factor = q.w * G.detabs;
    switch (c.op) {
      case OP_PNP: {
        if (s.cylindrical) factor *= gy * 2 * PI;
        double u[3], gu[3][2];
        for (int k = 0; k < 3; k++) {
          u[k] = 0.0; gu[k][0] = gu[k][1] = 0.0;
          for (int i = 0; i < NL; i++) u[k] += xl[NL * k + i] * B.phi[i];
          for (int i = 0; i < NL; i++) { gu[k][0] += xl[NL * k + i] * B.g[i][0]; gu[k][1] += xl[NL * k + i] * B.g[i][1]; }
        }
        for (int i = 0; i < NL; i++) rl[i] += (dot(gu[0], B.g[i]) + 4 * PI * s.l_b * (u[1] - u[2]) * B.phi[i]) * factor;
        for (int i = 0; i < NL; i++) rl[NL + i] += (dot(gu[1], B.g[i]) - u[1] * dot(gu[0], B.g[i])) * factor;
        for (int i = 0; i < NL; i++) rl[2 * NL + i] += (dot(gu[2], B.g[i]) + u[2] * dot(gu[0], B.g[i])) * factor;
        break;
      }
      case OP_PB: case OP_POISSON: {
        if (s.cylindrical) factor *= gy * 2 * PI;
        double u = 0.0, gu[2] = {0.0, 0.0};
        for (int i = 0; i < NL; i++) u += xl[i] * B.phi[i];
        for (int i = 0; i < NL; i++) { gu[0] += xl[i] * B.g[i][0]; gu[1] += xl[i] * B.g[i][1]; }
        double src;
        if (c.op == OP_PB) src = 8 * PI * s.l_b * s.c0 * sinh_shared(u);
        else {
          double cp = 0.0, cm = 0.0;
          for (int i = 0; i < NL; i++) cp += caux[0][i] * B.phi[i];
          for (int i = 0; i < NL; i++) cm += caux[1][i] * B.phi[i];
          src = 1 * s.l_b * 4 * PI * (cm - cp);
        }
        for (int i = 0; i < NL; i++) rl[i] += (dot(gu, B.g[i]) + src * B.phi[i]) * factor;
        break;
      }
      case OP_DIFFUSION: {
        double u = 0.0, gu[2] = {0.0, 0.0}, gP[2] = {0.0, 0.0};
        for (int i = 0; i < NL; i++) u += xl[i] * B.phi[i];
        for (int i = 0; i < NL; i++) { gu[0] += xl[i] * B.g[i][0]; gu[1] += xl[i] * B.g[i][1]; }
        for (int i = 0; i < NL; i++) { gP[0] += caux[0][i] * B.g[i][0]; gP[1] += caux[0][i] * B.g[i][1]; }
        const double a = 0;
        for (int i = 0; i < NL; i++) rl[i] += (dot(gu, B.g[i]) + u * c.valency * dot(gP, B.g[i]) + a * u * B.phi[i]) * factor;
        break;
      }
      case OP_MASS: {
        double u = 0.0;
        for (int i = 0; i < NL; i++) u += xl[i] * B.phi[i];
        for (int i = 0; i < NL; i++) rl[i] += u * B.phi[i] * factor;
        break;
      }
    }
  }
}

static void alpha_boundary2(const OpCtx& c, int e, int f, double* rl) {
  if (c.op == OP_DIFFUSION || c.op == OP_MASS) return;
  const Mesh& m = *c.m; const Sysparams& s = *c.s;
  const int* tv = &m.tri[3 * e];
  const int seg = m.fseg[3 * e + f];
  const Surface& sf = s.surfaces.at(m.bphys[seg]);
  const int va = tv[FACE_V[f][0]], vb = tv[FACE_V[f][1]];
  const double ax = m.x[va], ay = m.y[va], bx = m.x[vb], by = m.y[vb];
  const double len = std::sqrt((bx - ax) * (bx - ax) + (by - ay) * (by - ay));
  const int nf = op_fields(c.op);
  for (const QL& q : line_rule(c.order())) {
    double l0, l1;
    if (f == 0) { l0 = q.t; l1 = 0.0; } else if (f == 1) { l0 = 0.0; l1 = q.t; } else { l0 = 1.0 - q.t; l1 = q.t; }
    double phi[NL];
    basis(l0, l1, phi);
    const double gy = ay + q.t * (by - ay);
    double factor = q.w * len;
    if (s.cylindrical) factor *= gy * 2 * s.PI;
    for (int k = 0; k < nf; k++) {
      const int comp = (c.op == OP_PNP) ? k : 0;
      if (sf.btype(comp) == 0) continue;
      const double j = sf.flux(comp);
      for (int i = 0; i < NL; i++) rl[NL * k + i] += j * phi[i] * factor;
    }
  }
}

static void jacobian_volume_fd2(const OpCtx& c, int e, const double* xl, const double (*caux)[NL], double* Ae, double eps) {
  const int n = NL * op_fields(c.op);
  std::vector<double> u(xl, xl + n), down(n, 0.0), up(n);
  alpha_volume2(c, e, u.data(), caux, down.data());
  for (int j = 0; j < n; j++) {
    std::fill(up.begin(), up.end(), 0.0);
    const double delta = eps * (1.0 + std::fabs(u[j]));
    u[j] += delta;
    alpha_volume2(c, e, u.data(), caux, up.data());
    for (int i = 0; i < n; i++) Ae[i * n + j] += (up[i] - down[i]) / delta;
    u[j] = xl[j];
  }
}

// exact derivative of alpha_volume2 (not in the reference; validates the FD path and gives clean Newton comparisons)
static void jacobian_volume_exact2(const OpCtx& c, int e, const double* xl, const double (*caux)[NL], double* Ae) {
  const Mesh& m = *c.m; const Sysparams& s = *c.s;
  const ElemGeo G = elem_geo(m, e);
  const int n = NL * op_fields(c.op);
  for (const QP& q : tri_rule(c.order())) {
    const BasisAt B = basis_at(G, q.xi0, q.xi1);
    const double gy = G.y0 + (G.y1 - G.y0) * q.xi0 + (G.y2 - G.y0) * q.xi1;
    double factor = q.w * G.detabs;
    const bool cyl = s.cylindrical && c.op != OP_DIFFUSION && c.op != OP_MASS;
    if (cyl) factor *= gy * 2 * s.PI;
    auto K = [&](int i, int j) { return B.g[j][0] * B.g[i][0] + B.g[j][1] * B.g[i][1]; };
    if (c.op == OP_PNP) {
      double u[3] = {0, 0, 0}, gP[2] = {0, 0};
      for (int k = 0; k < 3; k++) for (int i = 0; i < NL; i++) u[k] += xl[NL * k + i] * B.phi[i];
      for (int i = 0; i < NL; i++) { gP[0] += xl[i] * B.g[i][0]; gP[1] += xl[i] * B.g[i][1]; }
      const double kap = 4 * s.PI * s.l_b;
      for (int i = 0; i < NL; i++) {
        const double dPi = gP[0] * B.g[i][0] + gP[1] * B.g[i][1];
        for (int j = 0; j < NL; j++) {
          Ae[i * n + j] += K(i, j) * factor;
          Ae[i * n + NL + j] += kap * B.phi[j] * B.phi[i] * factor;
          Ae[i * n + 2 * NL + j] -= kap * B.phi[j] * B.phi[i] * factor;
          Ae[(NL + i) * n + j] -= u[1] * K(i, j) * factor;
          Ae[(NL + i) * n + NL + j] += (K(i, j) - B.phi[j] * dPi) * factor;
          Ae[(2 * NL + i) * n + j] += u[2] * K(i, j) * factor;
          Ae[(2 * NL + i) * n + 2 * NL + j] += (K(i, j) + B.phi[j] * dPi) * factor;
        }
      }
    } else {
      double u = 0, gP[2] = {0, 0};
      for (int i = 0; i < NL; i++) u += xl[i] * B.phi[i];
      if (c.op == OP_DIFFUSION) for (int i = 0; i < NL; i++) { gP[0] += caux[0][i] * B.g[i][0]; gP[1] += caux[0][i] * B.g[i][1]; }
      for (int i = 0; i < NL; i++) for (int j = 0; j < NL; j++) {
        double v = 0;
        switch (c.op) {
          case OP_PB: v = K(i, j) + 8 * s.PI * s.l_b * s.c0 * std::cosh(u) * B.phi[j] * B.phi[i]; break;
          case OP_POISSON: v = K(i, j); break;
          case OP_DIFFUSION: v = K(i, j) + B.phi[j] * c.valency * (gP[0] * B.g[i][0] + gP[1] * B.g[i][1]); break;
          case OP_MASS: v = B.phi[j] * B.phi[i]; break;
        }
        Ae[i * n + j] += v * factor;
      }
    }
  }
}

// local coefficients of the operator's coefficient fields (c.cp / c.cm / c.uphi: P2 vectors of length nd)
static void gather_aux(const Space2& sp, const OpCtx& c, int e, double (*caux)[NL]) {
  for (int i = 0; i < NL; i++) {
    const int d = sp.sdof(e, i);
    caux[0][i] = c.op == OP_POISSON ? c.cp[d] : (c.op == OP_DIFFUSION ? c.uphi[d] : 0.0);
    caux[1][i] = c.op == OP_POISSON ? c.cm[d] : 0.0;
  }
}

static void residual2(const Space2& sp, const OpCtx& c, const double* u, double* r, double* absr = nullptr) {
  const Mesh& m = *sp.m;
  const int nf = sp.fields, n = NL * nf, N = sp.N();
  std::fill(r, r + N, 0.0);
  if (absr) std::fill(absr, absr + N, 0.0);
  std::vector<double> xl(n), rl(n);
  double caux[2][NL];
  for (int e = 0; e < m.nT; e++) {
    for (int k = 0; k < nf; k++) for (int i = 0; i < NL; i++) xl[NL * k + i] = u[sp.gdof(k, sp.sdof(e, i))];
    gather_aux(sp, c, e, caux);
    std::fill(rl.begin(), rl.end(), 0.0);
    alpha_volume2(c, e, xl.data(), caux, rl.data());
    for (int fi = 0; fi < 3; fi++) { const int f = FACE_ITER[fi]; if (m.fseg[3 * e + f] >= 0) alpha_boundary2(c, e, f, rl.data()); }
    for (int k = 0; k < nf; k++) for (int i = 0; i < NL; i++) {
      const int g = sp.gdof(k, sp.sdof(e, i));
      r[g] += rl[NL * k + i];
      if (absr) absr[g] += std::fabs(rl[NL * k + i]);
    }
  }
  for (int d = 0; d < N; d++) if (sp.dirichlet[d]) r[d] = 0.0;
}

static void jacobian2(const Space2& sp, const OpCtx& c, const double* u, CSR& A, int mode = 0, double eps = 1e-11,
                      std::vector<double>* absA = nullptr) {
  const Mesh& m = *sp.m;
  const int nf = sp.fields, n = NL * nf, N = sp.N();
  std::fill(A.val.begin(), A.val.end(), 0.0);
  if (absA) absA->assign(A.val.size(), 0.0);
  std::vector<double> xl(n), Ae((size_t)n * n);
  double caux[2][NL];
  for (int e = 0; e < m.nT; e++) {
    for (int k = 0; k < nf; k++) for (int i = 0; i < NL; i++) xl[NL * k + i] = u[sp.gdof(k, sp.sdof(e, i))];
    gather_aux(sp, c, e, caux);
    std::fill(Ae.begin(), Ae.end(), 0.0);
    if (mode == 0) jacobian_volume_fd2(c, e, xl.data(), caux, Ae.data(), eps);
    else jacobian_volume_exact2(c, e, xl.data(), caux, Ae.data());
    for (int ki = 0; ki < nf; ki++) for (int i = 0; i < NL; i++) {
      const int gi = sp.gdof(ki, sp.sdof(e, i));
      if (sp.dirichlet[gi]) continue;
      for (int kj = 0; kj < nf; kj++) for (int j = 0; j < NL; j++) {
        const int gj = sp.gdof(kj, sp.sdof(e, j));
        if (sp.dirichlet[gj]) continue;
        const int slot = A.find(gi, gj);
        A.val[slot] += Ae[(size_t)(NL * ki + i) * n + NL * kj + j];
        if (absA) (*absA)[slot] += std::fabs(Ae[(size_t)(NL * ki + i) * n + NL * kj + j]);
      }
    }
  }
  for (int d = 0; d < N; d++) if (sp.dirichlet[d]) A.val[A.find(d, d)] = 1.0;
}

// BCExtension<component>::evaluate at local node i of element e (dirichlet_bc.hh:54-123): the position is the node's;
// the PB field is a P2 function, whose value at a Lagrange node is its dof
static double bcext_eval2(const Space2& sp, int comp, const double* pb, int e, int i) {
  const Mesh& m = *sp.m; const Sysparams& s = *sp.s;
  const int a = m.tri[3 * e], b = m.tri[3 * e + 1], cv = m.tri[3 * e + 2];
  // geometry().global(local): v0 + J * local
  const double px = m.x[a] + (m.x[b] - m.x[a]) * nodes().x[i] + (m.x[cv] - m.x[a]) * nodes().y[i];
  const double py = m.y[a] + (m.y[b] - m.y[a]) * nodes().x[i] + (m.y[cv] - m.y[a]) * nodes().y[i];
  int pg = -1;
  auto sticky = [&](int g) { return s.surfaces.at(g).minusDiffusionBtype == 0; };
  for (int fi = 0; fi < 3; fi++) {
    const int f = FACE_ITER[fi];
    if (m.fseg[3 * e + f] >= 0) {
      if (global_on_intersection(m, px, py, e, f))
        if (pg == -1 || !sticky(pg)) pg = m.bphys[m.fseg[3 * e + f]];
    } else {
      const int o = m.nbr[3 * e + f];
      for (int gi = 0; gi < 3; gi++) {
        const int f2 = FACE_ITER[gi];
        if (m.fseg[3 * o + f2] >= 0 && global_on_intersection(m, px, py, o, f2))
          if (pg == -1 || !sticky(pg)) pg = m.bphys[m.fseg[3 * o + f2]];
      }
    }
  }
  if (pg > -1 && s.surfaces.at(pg).btype(comp) == 0) return s.surfaces[pg].dirichlet(comp);
  const double yv = pb ? pb[sp.sdof(e, i)] : 0.0;
  if (comp == 0) return yv;
  if (comp == 1) return s.c0 * std::exp(-yv);
  return s.c0 * std::exp(+yv);
}
static void interpolate_bcext2(const Space2& sp, int comp, const double* pb, double* u) {
  for (int e = 0; e < sp.m->nT; e++)
    for (int i = 0; i < NL; i++) u[sp.sdof(e, i)] = bcext_eval2(sp, comp, pb, e, i);
}

// OneStepMethod on OneStepGridOperator<DiffusionOperator, DiffusionTOperator> (instationary_pnp_from_pb_md.hh:368-391,421-425)
static OneStepResult onestep2(const Space2& sp, const OpCtx& c0, const OpCtx& c1, const TimeMethod& tm, double dt, const double* xold,
                              const double* g, double* xnew, double reduction, int solver, int prec, int steps, int maxit,
                              int jac_mode = 0, double eps = 1e-11) {
  return onestep_core(sp.N(), sp.dirichlet, make_pattern2(sp),
                      [&](const double* x, double* r) { residual2(sp, c0, x, r); }, [&](const double* x, double* r) { residual2(sp, c1, x, r); },
                      [&](const double* x, CSR& A) { jacobian2(sp, c0, x, A, jac_mode, eps); },
                      [&](const double* x, CSR& A) { jacobian2(sp, c1, x, A, jac_mode, eps); }, tm, dt, xold, g, xnew, reduction, solver,
                      prec, steps, maxit);
}

// calcIonFlux (ionFlux.hh:8-96) with quadratic functions: DiscreteGridFunction / DiscreteGridFunctionGradient evaluated at
// the local coordinates of the face centre -- the basis sum over the element's 6 dofs per field
static void ion_flux2(const Space2& sp, const double* phi, const double* cp, const double* cm, double* ip, double* im) {
  const Mesh& m = *sp.m; const Sysparams& s = *sp.s;
  static const double VX[3] = {0.0, 1.0, 0.0}, VY[3] = {0.0, 0.0, 1.0}; // reference vertices
  for (int i = 0; i < s.n_surfaces; i++) { ip[i] = 0; im[i] = 0; }
  for (int e = 0; e < m.nT; e++) {
    const ElemGeo g = elem_geo(m, e);
    const int* tv = &m.tri[3 * e];
    for (int gi = 0; gi < 3; gi++) {
      const int f = FACE_ITER[gi];
      const int seg = m.fseg[3 * e + f];
      if (seg < 0) continue;
      const int la = FACE_V[f][0], lb = FACE_V[f][1], lc = 3 - la - lb;
      const double ax = m.x[tv[la]], ay = m.y[tv[la]], bx = m.x[tv[lb]], by = m.y[tv[lb]];
      const double ex = 0.5 * (ax + bx), ey = 0.5 * (ay + by);
      const BasisAt B = basis_at(g, 0.5 * (VX[la] + VX[lb]), 0.5 * (VY[la] + VY[lb]));
      double vcp = 0, vcm = 0, gphi[2] = {0, 0}, gcp[2] = {0, 0}, gcm[2] = {0, 0};
      for (int k = 0; k < NL; k++) {
        const int d = sp.sdof(e, k);
        vcp += cp[d] * B.phi[k]; vcm += cm[d] * B.phi[k];
        for (int r = 0; r < 2; r++) { gphi[r] += phi[d] * B.g[k][r]; gcp[r] += cp[d] * B.g[k][r]; gcm[r] += cm[d] * B.g[k][r]; }
      }
      const double len = std::sqrt((bx - ax) * (bx - ax) + (by - ay) * (by - ay));
      double factor = len;
      if (s.cylindrical) factor *= 2 * s.PI * ey;
      for (int r = 0; r < 2; r++) { gcp[r] *= -factor; gcm[r] *= -factor; gphi[r] *= factor; gphi[r] *= vcp; }
      double nx = (by - ay) / len, ny = -(bx - ax) / len;
      if (nx * (m.x[tv[lc]] - ex) + ny * (m.y[tv[lc]] - ey) > 0) { nx = -nx; ny = -ny; }
      const int pg = m.bphys[seg];
      ip[pg] += (gcp[0] + gphi[0]) * nx + (gcp[1] + gphi[1]) * ny;
      const double ratio = vcm / vcp;
      for (int r = 0; r < 2; r++) gphi[r] *= ratio;
      im[pg] += (gcm[0] - gphi[0]) * nx + (gcm[1] - gphi[1]) * ny;
    }
  }
}

// DataWriter::writeData (datawriter.hh:45-94) with a quadratic function: value and gradient at the element centre
static void write_cell_data2(const Space2& sp, const double* u, const std::string& filename) {
  const Mesh& m = *sp.m;
  std::ofstream out(filename.c_str(), std::ios::out);
  out.precision(5);
  for (int e = 0; e < m.nT; e++) {
    const ElemGeo g = elem_geo(m, e);
    const double cx = (g.x0 + g.x1 + g.x2) / 3.0, cy = (g.y0 + g.y1 + g.y2) / 3.0;
    const BasisAt B = basis_at(g, 1.0 / 3.0, 1.0 / 3.0);
    double val = 0, gr[2] = {0, 0};
    for (int k = 0; k < NL; k++) {
      const double uk = u[sp.sdof(e, k)];
      val += uk * B.phi[k];
      for (int d = 0; d < 2; d++) gr[d] += uk * B.g[k][d];
    }
    out << std::left << std::scientific << cx << " " << cy << "\t";
    out << std::left << val << "\t";
    out << std::left << gr[0] << " " << gr[1] << std::endl;
  }
}

}; // struct Pk
using p2 = Pk<2>;
using p3 = Pk<3>;
} // namespace pnpo
