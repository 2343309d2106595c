// pnp_elem_p2.cuh -- element-level fp64 math of the five local operators with QUADRATIC and CUBIC elements (the reference's
// -DPDEGREE=2 / -DPDEGREE=3 builds: /root/reference/src/Makefile.am:54-110; Pk2DLocalFiniteElementMap<GV,D,R,PDEGREE>,
// instationary_pnp_from_pb_md.hh:26-28,125).  The operator bodies are the reference's alpha_volume / alpha_boundary
// (pnp_operator.hh:98-194, pb_operator.hh:74-120, poisson_operator.hh:74-126, diffusion_operator.hh:64-111,
// diffusion_toperator.hh:58-72) evaluated for lfsu.size() = 6 or 10; __host__ __device__ like pnp_elem.cuh, compiled without
// FMA contraction so that the FD Jacobian rounds like the CPU restatement.
//
// Local dof order (Pk2DLocalBasis<D,R,k>): Lagrange nodes (i, j)/k, i + j <= k, lexicographic with j outer.
// k = 2: (0,0),(1/2,0),(1,0),(0,1/2),(1/2,1/2),(0,1) = vertex0, edge0=(v0,v1), vertex1, edge1=(v0,v2), edge2=(v1,v2), vertex2.
// k = 3: v0, e0, e0, v1, e1, bubble, e2, e1, e2, v2.
// NQ is the triangle rule (3, 4 or 7 points: orders 2, 3, 5): the reference's constructors default to order 3 whatever PDEGREE is,
// which under-integrates the cubic stiffness and mass terms (and makes the matrices indefinite: the order-3 rule has a negative
// weight); pnp_operator_set_intorder(.., 5) selects the 7-point rule.
#pragma once
#include "pnp_elem.cuh"

namespace pnp {
template <int DEG> struct PkElem {
  static_assert(DEG == 2 || DEG == 3, "degrees 2 and 3");

static constexpr int NL = (DEG + 1) * (DEG + 2) / 2;

// node n <-> lattice point (i, j); what it sits on: kind 0 vertex (sub = local vertex), 1 edge (sub = local edge, idx = position
// counted from the edge's first local vertex), 2 element interior
PNP_HD static void node_ij(int n, int& i, int& j) {
  int c = 0;
  for (j = 0; j <= DEG; j++) for (i = 0; i <= DEG - j; i++) if (c++ == n) return;
}
PNP_HD static double node_x(int n) { int i, j; node_ij(n, i, j); return (1.0 * i) / DEG; }
PNP_HD static double node_y(int n) { int i, j; node_ij(n, i, j); return (1.0 * j) / DEG; }
PNP_HD static void node_key(int n, int& kind, int& sub, int& idx) {
  int i, j; node_ij(n, i, j);
  idx = 0;
  if (i == 0 && j == 0) { kind = 0; sub = 0; }
  else if (i == DEG) { kind = 0; sub = 1; }
  else if (j == DEG) { kind = 0; sub = 2; }
  else if (j == 0) { kind = 1; sub = 0; idx = i - 1; }
  else if (i == 0) { kind = 1; sub = 1; idx = j - 1; }
  else if (i + j == DEG) { kind = 1; sub = 2; idx = j - 1; }
  else { kind = 2; sub = 0; }
}

// Pk2DLocalBasis<D,R,k>::evaluateFunction / evaluateJacobian for any k: pos[i] = i/k; node (i,j):
// prod_{a<i} (x-pos[a])/(pos[i]-pos[a]) prod_{b<j} (y-pos[b])/(pos[j]-pos[b]) prod_{g=i+j+1..k} (pos[g]-x-y)/(pos[g]-pos[i]-pos[j]);
// derivatives by the product rule, factor by factor (the operation sequence of the CPU restatement)
PNP_HD static void basis_generic(double x, double y, double* phi) {
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
PNP_HD static void basis_grad_generic(double x, double y, double (*gr)[2]) {
  double pos[DEG + 1];
  for (int i = 0; i <= DEG; i++) pos[i] = (1.0 * i) / DEG;
  int n = 0;
  for (int j = 0; j <= DEG; j++) for (int i = 0; i <= DEG - j; i++, n++) {
    for (int dir = 0; dir < 2; dir++) {
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

// Pk2DLocalBasis<D,R,2>::evaluateFunction: node (i,j) -> prod_{a<i}(2x-a)/(i-a) prod_{b<j}(2y-b)/(j-b) prod_{g>i+j}(g-2x-2y)/(g-i-j)
PNP_HD static void basis(double x, double y, double* phi) {
  if (DEG != 2) { basis_generic(x, y, phi); return; }
  const double s = 2 * x + 2 * y;
  phi[0] = ((1 - s) / 1) * ((2 - s) / 2);
  phi[1] = (2 * x) * ((2 - s) / 1);
  phi[2] = (2 * x) * ((2 * x - 1) / 2);
  phi[3] = (2 * y) * ((2 - s) / 1);
  phi[4] = (2 * x) * (2 * y);
  phi[5] = (2 * y) * ((2 * y - 1) / 2);
}
PNP_HD static void basis_grad(double x, double y, double (*g)[2]) {
  if (DEG != 2) { basis_grad_generic(x, y, g); return; }
  const double s = 2 * x + 2 * y;
  g[0][0] = (2 * s - 3); g[0][1] = (2 * s - 3);
  g[1][0] = 2 * (2 - s) - 4 * x; g[1][1] = -4 * x;
  g[2][0] = 4 * x - 1; g[2][1] = 0.0;
  g[3][0] = -4 * y; g[3][1] = 2 * (2 - s) - 4 * y;
  g[4][0] = 4 * y; g[4][1] = 4 * x;
  g[5][0] = 0.0; g[5][1] = 4 * y - 1;
}

// affine geometry: J^{-T}, |det J|, the vertices' y for the cylindrical factor
struct Geo2 { double jit[2][2], detabs, y0, y1, y2; };
PNP_HD static Geo2 make_geo2(double x0, double y0, double x1, double y1, double x2, double y2) {
  Geo2 G;
  const double j00 = x1 - x0, j01 = x2 - x0, j10 = y1 - y0, j11 = y2 - y0;
  const double det = j00 * j11 - j01 * j10;
  const double di = 1.0 / det;
  G.jit[0][0] = j11 * di;  G.jit[0][1] = -j10 * di;
  G.jit[1][0] = -j01 * di; G.jit[1][1] = j00 * di;
  G.detabs = fabs(det);
  G.y0 = y0; G.y1 = y1; G.y2 = y2;
  return G;
}
struct BasisAt { double phi[NL], g[NL][2]; };
PNP_HD static BasisAt basis_at(const Geo2& G, double x, double y) {
  BasisAt B;
  basis(x, y, B.phi);
  double gh[NL][2];
  basis_grad(x, y, gh);
#pragma unroll
  for (int i = 0; i < NL; i++)
#pragma unroll
    for (int r = 0; r < 2; r++) { // FieldMatrix::mv
      double v = 0.0;
      v += G.jit[r][0] * gh[i][0];
      v += G.jit[r][1] * gh[i][1];
      B.g[i][r] = v;
    }
  return B;
}

// basis values / transformed gradients and integration factor (weight * |det J| [* 2 pi r]) at every point of the operator's rule
template <int NQ> struct QuadData { BasisAt B[NQ]; double factor[NQ]; };
template <int OP, int NQ>
PNP_HD static void quad_data(const Geo2& G, const PhysParams& P, QuadData<NQ>& Q) {
  for (int q = 0; q < NQ; q++) {
    double xi0, xi1, w;
    quad_point<NQ>(q, xi0, xi1, w);
    Q.B[q] = basis_at(G, xi0, xi1);
    const double gy = G.y0 + (G.y1 - G.y0) * xi0 + (G.y2 - G.y0) * xi1;
    double factor = w * G.detabs;
    if ((OP == OP_PNP || OP == OP_PB || OP == OP_POISSON) && P.cylindrical) factor *= gy * 2 * P.PI;
    Q.factor[q] = factor;
  }
}
// alpha_volume: xl[NL*k + i] = local coefficient of field k at node i; caux[a][i] = local coefficients of the operator's
// coefficient fields (Poisson: c+, c-; diffusion: Phi), Pk functions as well.  ACCUMULATES into rl[NL*k + i].
template <int OP, int NQ>
PNP_HD static void alpha_volume_q(const QuadData<NQ>& Q, const PhysParams& P, const double* xl, const double (*caux)[NL], double* rl) {
  const double PI = P.PI;
  for (int q = 0; q < NQ; q++) {
    const BasisAt& B = Q.B[q];
    const double factor = Q.factor[q];
    if (OP == OP_PNP) {
      double u[3], gu[3][2];
      for (int k = 0; k < 3; k++) {
        u[k] = 0.0; gu[k][0] = 0.0; gu[k][1] = 0.0;
        for (int i = 0; i < NL; i++) u[k] += xl[NL * k + i] * B.phi[i];
        for (int i = 0; i < NL; i++) { gu[k][0] += xl[NL * k + i] * B.g[i][0]; gu[k][1] += xl[NL * k + i] * B.g[i][1]; }
      }
      for (int i = 0; i < NL; i++) rl[i] += (dot2(gu[0], B.g[i]) + 4 * PI * P.l_b * (u[1] - u[2]) * B.phi[i]) * factor;
      for (int i = 0; i < NL; i++) rl[NL + i] += (dot2(gu[1], B.g[i]) - u[1] * dot2(gu[0], B.g[i])) * factor;
      for (int i = 0; i < NL; i++) rl[2 * NL + i] += (dot2(gu[2], B.g[i]) + u[2] * dot2(gu[0], B.g[i])) * factor;
    } else if (OP == OP_PB || OP == OP_POISSON) {
      double u = 0.0, gu[2] = {0.0, 0.0};
      for (int i = 0; i < NL; i++) u += xl[i] * B.phi[i];
      for (int i = 0; i < NL; i++) { gu[0] += xl[i] * B.g[i][0]; gu[1] += xl[i] * B.g[i][1]; }
      double src;
      if (OP == OP_PB) src = 8 * PI * P.l_b * P.c0 * pnp_sinh(u);
      else {
        double cp = 0.0, cm = 0.0;
        for (int i = 0; i < NL; i++) cp += caux[0][i] * B.phi[i];
        for (int i = 0; i < NL; i++) cm += caux[1][i] * B.phi[i];
        src = 1 * P.l_b * 4 * PI * (cm - cp);
      }
      for (int i = 0; i < NL; i++) rl[i] += (dot2(gu, B.g[i]) + src * B.phi[i]) * factor;
    } else if (OP == OP_DIFFUSION) {
      double u = 0.0, gu[2] = {0.0, 0.0}, gP[2] = {0.0, 0.0};
      for (int i = 0; i < NL; i++) u += xl[i] * B.phi[i];
      for (int i = 0; i < NL; i++) { gu[0] += xl[i] * B.g[i][0]; gu[1] += xl[i] * B.g[i][1]; }
      for (int i = 0; i < NL; i++) { gP[0] += caux[0][i] * B.g[i][0]; gP[1] += caux[0][i] * B.g[i][1]; }
      const double a = 0;
      for (int i = 0; i < NL; i++) rl[i] += (dot2(gu, B.g[i]) + u * P.valency * dot2(gP, B.g[i]) + a * u * B.phi[i]) * factor;
    } else { // OP_MASS
      double u = 0.0;
      for (int i = 0; i < NL; i++) u += xl[i] * B.phi[i];
      for (int i = 0; i < NL; i++) rl[i] += u * B.phi[i] * factor;
    }
  }
}
template <int OP, int NQ>
PNP_HD static void alpha_volume(const Geo2& G, const PhysParams& P, const double* xl, const double (*caux)[NL], double* rl) {
  QuadData<NQ> Q;
  quad_data<OP, NQ>(G, P, Q);
  alpha_volume_q<OP, NQ>(Q, P, xl, caux, rl);
}

// alpha_boundary of DUNE face f with end points (ax,ay)->(bx,by) in element order; j[k] = flux of field k, skip[k]: the face
// is Dirichlet for field k's component; npts = points of the Gauss-Legendre rule the operator's intorder asks for (2: order 3,
// 3: order 5).  ACCUMULATES into rl.
PNP_HD static void alpha_boundary(int f, double ax, double ay, double bx, double by, int nfields, const double* j, const bool* skip,
                                  const PhysParams& P, int npts, double* rl) {
  const double len = sqrt((bx - ax) * (bx - ax) + (by - ay) * (by - ay));
  const double t2[2] = {0.21132486540518711775, 0.78867513459481288225};
  const double t3[3] = {0.11270166537925831148, 0.5, 0.88729833462074168852};
  const double w3[3] = {5.0 / 18.0, 8.0 / 18.0, 5.0 / 18.0};
  for (int q = 0; q < npts; q++) {
    const double t = npts == 2 ? t2[q] : t3[q];
    double l0, l1;
    if (f == 0) { l0 = t; l1 = 0.0; } else if (f == 1) { l0 = 0.0; l1 = t; } else { l0 = 1.0 - t; l1 = t; }
    double phi[NL];
    basis(l0, l1, phi);
    const double gy = ay + t * (by - ay);
    double factor = (npts == 2 ? 0.5 : w3[q]) * len;
    if (P.cylindrical) factor *= gy * 2 * P.PI;
    for (int k = 0; k < nfields; k++) {
      if (skip[k]) continue;
      for (int i = 0; i < NL; i++) rl[NL * k + i] += j[k] * phi[i] * factor;
    }
  }
}

// NumericalJacobianVolume: Ae[i*n + j], n = NL * F, ACCUMULATED
template <int OP, int NQ>
PNP_HD static void jacobian_fd(const Geo2& G, const PhysParams& P, double* xl, const double (*caux)[NL], double eps, double* Ae) {
  constexpr int n = NL * OpTraits<OP>::F;
  double down[n], up[n];
  QuadData<NQ> Q; // the n + 1 evaluations share the basis tables (they do not depend on the state)
  quad_data<OP, NQ>(G, P, Q);
  for (int i = 0; i < n; i++) down[i] = 0.0;
  alpha_volume_q<OP, NQ>(Q, P, xl, caux, down);
  for (int j = 0; j < n; j++) {
    for (int i = 0; i < n; i++) up[i] = 0.0;
    const double keep = xl[j];
    const double delta = eps * (1.0 + fabs(keep));
    xl[j] = keep + delta;
    alpha_volume_q<OP, NQ>(Q, P, xl, caux, up);
    for (int i = 0; i < n; i++) Ae[i * n + j] += (up[i] - down[i]) / delta;
    xl[j] = keep;
  }
}

// exact derivative (not in the reference).  Every entry is the sum over the quadrature points, in point order, of the same terms
// the point-by-point accumulation adds -- summed in a register and added to Ae once.
template <int OP, int NQ>
PNP_HD static void jacobian_exact(const Geo2& G, const PhysParams& P, const double* xl, const double (*caux)[NL], double* Ae) {
  constexpr int n = NL * OpTraits<OP>::F;
  QuadData<NQ> Q;
  quad_data<OP, NQ>(G, P, Q);
  if (OP == OP_PNP) {
    double u1[NQ], u2[NQ], gP[NQ][2];
    for (int q = 0; q < NQ; q++) {
      const BasisAt& B = Q.B[q];
      double u[3] = {0, 0, 0};
      for (int k = 0; k < 3; k++) for (int i = 0; i < NL; i++) u[k] += xl[NL * k + i] * B.phi[i];
      u1[q] = u[1]; u2[q] = u[2];
      gP[q][0] = 0; gP[q][1] = 0;
      for (int i = 0; i < NL; i++) { gP[q][0] += xl[i] * B.g[i][0]; gP[q][1] += xl[i] * B.g[i][1]; }
    }
    const double kap = 4 * P.PI * P.l_b;
    for (int i = 0; i < NL; i++)
      for (int j = 0; j < NL; j++) {
        double a00 = 0, a01 = 0, a02 = 0, a10 = 0, a11 = 0, a20 = 0, a22 = 0;
        for (int q = 0; q < NQ; q++) {
          const BasisAt& B = Q.B[q];
          const double factor = Q.factor[q];
          const double dPi = gP[q][0] * B.g[i][0] + gP[q][1] * B.g[i][1];
          const double K = B.g[j][0] * B.g[i][0] + B.g[j][1] * B.g[i][1];
          a00 += K * factor;
          a01 += kap * B.phi[j] * B.phi[i] * factor;
          a02 -= kap * B.phi[j] * B.phi[i] * factor;
          a10 -= u1[q] * K * factor;
          a11 += (K - B.phi[j] * dPi) * factor;
          a20 += u2[q] * K * factor;
          a22 += (K + B.phi[j] * dPi) * factor;
        }
        Ae[i * n + j] += a00; Ae[i * n + NL + j] += a01; Ae[i * n + 2 * NL + j] += a02;
        Ae[(NL + i) * n + j] += a10; Ae[(NL + i) * n + NL + j] += a11;
        Ae[(2 * NL + i) * n + j] += a20; Ae[(2 * NL + i) * n + 2 * NL + j] += a22;
      }
  } else {
    double ch[NQ], gP[NQ][2];
    for (int q = 0; q < NQ; q++) {
      const BasisAt& B = Q.B[q];
      double u = 0;
      for (int i = 0; i < NL; i++) u += xl[i] * B.phi[i];
      gP[q][0] = 0; gP[q][1] = 0;
      if (OP == OP_DIFFUSION) for (int i = 0; i < NL; i++) { gP[q][0] += caux[0][i] * B.g[i][0]; gP[q][1] += caux[0][i] * B.g[i][1]; }
      ch[q] = OP == OP_PB ? 8 * P.PI * P.l_b * P.c0 * cosh(u) : 0.0;
    }
    for (int i = 0; i < NL; i++)
      for (int j = 0; j < NL; j++) {
        double acc = 0;
        for (int q = 0; q < NQ; q++) {
          const BasisAt& B = Q.B[q];
          const double K = B.g[j][0] * B.g[i][0] + B.g[j][1] * B.g[i][1];
          double v;
          if (OP == OP_PB) v = K + ch[q] * B.phi[j] * B.phi[i];
          else if (OP == OP_POISSON) v = K;
          else if (OP == OP_DIFFUSION) v = K + B.phi[j] * P.valency * (gP[q][0] * B.g[i][0] + gP[q][1] * B.g[i][1]);
          else v = B.phi[j] * B.phi[i];
          acc += v * Q.factor[q];
        }
        Ae[i * n + j] += acc;
      }
  }
}

}; // struct PkElem
} // namespace pnp
