// pnp_oracle.hpp -- CPU ORACLE (test infrastructure, NOT product code).
//
// A plain, sequential C++ restatement of the dune-pnp Newton-step hot path:
// the five local operators of the reference plus the DUNE/PDELab/ISTL semantics
// they rely on.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference leg may use anything in oracle/.
//
// PARITY UNPINNED: the reference ships no golden vectors and DUNE is not
// installable here, so every [UPSTREAM] rule below is a documented assumption
//  and the pins are analytic solutions, patch tests,
// FD-vs-analytic agreement and h-convergence (tests/test_oracle_*.py).
//
// Reference files followed (relative to /root/reference/src):
//   pnp_operator.hh:47-195,199-315    pb_operator.hh:46-120,126-192
//   poisson_operator.hh:46-127,131-199 diffusion_operator.hh:42-112
//   diffusion_toperator.hh:38-73      btype.hh:21-53   dirichlet_bc.hh:21-123
//   sysparams.cc:15-116               stationary_pnp_from_pb.hh:92-370
//   instationary_pnp_from_pb_md.hh:112-455
//
// Build: g++ -O2 -ffp-contract=off  (no FMA contraction: the GPU "faithful"
// path is compiled with -fmad=false so element results can be compared bitwise).
#pragma once
#include <algorithm>
#ifdef _OPENMP
#include <omp.h>
#endif
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace pnpo {

// ----------------------------------------------------------------------------
// Parameters (sysparams.hh:10-57, sysparams.cc:15-116)
// ----------------------------------------------------------------------------
struct Surface {
  // defaults: Surface::Surface() sysparams.cc:101-116 (all Neumann, zero flux)
  int coulombBtype = 1;
  double coulombFlux = 0, coulombPotential = 0;
  int plusDiffusionBtype = 1;
  double plusDiffusionFlux = 0, plusDiffusionConcentration = 0;
  int minusDiffusionBtype = 1;
  double minusDiffusionFlux = 0, minusDiffusionConcentration = 0;
  int btype(int comp) const {
    return comp == 0 ? coulombBtype : comp == 1 ? plusDiffusionBtype : minusDiffusionBtype;
  }
  double flux(int comp) const {
    return comp == 0 ? coulombFlux : comp == 1 ? plusDiffusionFlux : minusDiffusionFlux;
  }
  double dirichlet(int comp) const {
    return comp == 0 ? coulombPotential
                     : comp == 1 ? plusDiffusionConcentration : minusDiffusionConcentration;
  }
};

struct Sysparams {
  std::string meshfile;
  int n_surfaces = 0;
  int verbosity = 0;
  bool cylindrical = false;
  double l_b = 1.0;
  int linearSolverIterations = 50;
  double newtonReassembleThreshold = 0.0;
  double newtonReduction = 1e-5;
  double newtonMinLinearReduction = 1e-5;
  double newtonMaxIterations = 50;            // stored as double: sysparams.hh:28-29 (quirk B10)
  double newtonLineSearchMaxIteration = 500;
  double c0 = 0.06;
  double tau = 0.1;
  int outputFreq = 1, nSteps = 100, potentialUpdateFreq = 1, printStiffnessMatrix = 0;
  double PI = 3.1415;                         // #define PI 3.1415  pnp_operator.hh:20 (quirk B1)
  std::vector<Surface> surfaces;
};

// Minimal INI reader with the key set of sysparams.cc:31-94.  Keys missing from the
// stale sphere/cylinder cfgs get the one_wall.cfg values.
inline std::map<std::string, std::string> read_ini(const std::string& fn) {
  std::ifstream in(fn);
  if (!in) throw std::runtime_error("Could not read config file \"" + fn + "\"!");
  std::map<std::string, std::string> kv;
  std::string line, sec;
  auto trim = [](std::string s) {
    size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
    return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
  };
  while (std::getline(in, line)) {
    size_t h = line.find('#');
    if (h != std::string::npos) line = line.substr(0, h);
    line = trim(line);
    if (line.empty()) continue;
    if (line[0] == '[') { sec = trim(line.substr(1, line.find(']') - 1)); continue; }
    size_t e = line.find('=');
    if (e == std::string::npos) continue;
    kv[sec + "." + trim(line.substr(0, e))] = trim(line.substr(e + 1));
  }
  return kv;
}

inline Sysparams read_config(const std::string& fn) {
  auto kv = read_ini(fn);
  Sysparams s;
  auto has = [&](const std::string& k) { return kv.count(k) > 0; };
  auto req = [&](const std::string& k) -> std::string {
    if (!has(k)) throw std::runtime_error("config key missing: " + k);
    return kv[k];
  };
  auto getd = [&](const std::string& k, double dflt) { return has(k) ? std::stod(kv[k]) : dflt; };
  s.meshfile = req("mesh.filename");
  s.n_surfaces = std::stoi(req("system.n_surfaces"));
  s.verbosity = (int)getd("system.verbosity", 0);
  s.cylindrical = (int)getd("system.cylindrical", 0) != 0;
  s.l_b = getd("system.l_b", 1.0);
  s.linearSolverIterations = (int)getd("system.linearSolverIterations", 50);
  s.newtonReassembleThreshold = getd("system.newtonReassembleThreshold", 0.0);
  s.newtonReduction = getd("system.newtonReduction", 1e-5);
  s.newtonMinLinearReduction = getd("system.newtonMinLinearReduction", 1e-5);
  s.newtonMaxIterations = getd("system.newtonMaxIterations", 50);
  s.newtonLineSearchMaxIteration = getd("system.newtonLineSearchMaxIteration", 500);
  s.tau = getd("system.tau", 0.1);
  s.c0 = getd("system.c0", 0.06);
  s.nSteps = (int)getd("system.nSteps", 100);
  s.outputFreq = (int)getd("system.outputFreq", 1);
  s.potentialUpdateFreq = (int)getd("system.potentialUpdateFreq", 1);
  s.printStiffnessMatrix = (int)getd("system.printStiffnessMatrix", 0);
  s.surfaces.assign(s.n_surfaces, Surface());
  for (int i = 0; i < s.n_surfaces; i++) {
    std::string p = "surface_" + std::to_string(i);
    Surface& f = s.surfaces[i];
    // only the value matching the btype is read: sysparams.cc:70-93
    f.coulombBtype = std::stoi(req(p + ".coulombBtype"));
    if (f.coulombBtype == 0) f.coulombPotential = std::stod(req(p + ".coulombPotential"));
    if (f.coulombBtype == 1) f.coulombFlux = std::stod(req(p + ".coulombFlux"));
    f.plusDiffusionBtype = std::stoi(req(p + ".plusDiffusionBtype"));
    if (f.plusDiffusionBtype == 0)
      f.plusDiffusionConcentration = std::stod(req(p + ".plusDiffusionConcentration"));
    if (f.plusDiffusionBtype == 1) f.plusDiffusionFlux = std::stod(req(p + ".plusDiffusionFlux"));
    f.minusDiffusionBtype = std::stoi(req(p + ".minusDiffusionBtype"));
    if (f.minusDiffusionBtype == 0)
      f.minusDiffusionConcentration = std::stod(req(p + ".minusDiffusionConcentration"));
    if (f.minusDiffusionBtype == 1) f.minusDiffusionFlux = std::stod(req(p + ".minusDiffusionFlux"));
  }
  return s;
}

// ----------------------------------------------------------------------------
// Mesh: GmshReader semantics (pnp_solver_main.cc:82-91)
// ----------------------------------------------------------------------------
// DUNE reference triangle: face f -> local vertices {0,1},{0,2},{1,2} (App. A.6).
static const int FACE_V[3][2] = {{0, 1}, {0, 2}, {1, 2}};
// [UPSTREAM assumption] UGGrid iterates intersections in UG side order
// (v0v1),(v1v2),(v2v0) = DUNE faces 0,2,1.
static const int FACE_ITER[3] = {0, 2, 1};

struct Mesh {
  int nv = 0, nT = 0, nB = 0;
  std::vector<double> x, y;       // vertex coordinates (z dropped)
  std::vector<int> tri;           // 3*nT, file order, local order as in file
  std::vector<int> ba, bb, bphys; // boundary segments in file order: end vertices, physical tag (=b2e)
  // derived by finalize():
  std::vector<int> nbr;           // 3*nT: neighbour element across DUNE face f, -1 on boundary
  std::vector<int> fseg;          // 3*nT: boundarySegmentIndex of face f, -1 if interior
  void finalize() {
    nv = (int)x.size(); nT = (int)tri.size() / 3; nB = (int)ba.size();
    nbr.assign(3 * nT, -1); fseg.assign(3 * nT, -1);
    // (edge key, 3*e + f) sorted by key: the two faces of an interior edge are neighbours in the list
    auto key = [](int a, int b) { return ((uint64_t)std::min(a, b) << 32) | (uint64_t)std::max(a, b); };
    std::vector<std::pair<uint64_t, int>> faces; faces.reserve(3 * (size_t)nT);
    for (int e = 0; e < nT; e++)
      for (int f = 0; f < 3; f++) faces.push_back({key(tri[3 * e + FACE_V[f][0]], tri[3 * e + FACE_V[f][1]]), 3 * e + f});
    std::sort(faces.begin(), faces.end());
    std::vector<std::pair<uint64_t, int>> segs; segs.reserve(nB);
    for (int s = 0; s < nB; s++) segs.push_back({key(ba[s], bb[s]), s});
    std::sort(segs.begin(), segs.end()); // (a segment listed twice: the later one wins, as with the map this replaces)
    for (size_t i = 0; i < faces.size();) {
      size_t j = i + 1;
      while (j < faces.size() && faces[j].first == faces[i].first) j++;
      if (j - i > 2) throw std::runtime_error("edge shared by more than two elements");
      if (j - i == 2) {
        nbr[faces[i].second] = faces[i + 1].second / 3;
        nbr[faces[i + 1].second] = faces[i].second / 3;
      } else {
        auto it = std::upper_bound(segs.begin(), segs.end(), std::make_pair(faces[i].first, 0x7fffffff));
        if (it == segs.begin() || (it - 1)->first != faces[i].first) throw std::runtime_error("boundary face without boundary segment");
        fseg[faces[i].second] = (it - 1)->second;
      }
      i = j;
    }
  }
};

inline Mesh read_gmsh(const std::string& fn) {
  std::ifstream in(fn);
  if (!in) throw std::runtime_error("cannot open mesh " + fn);
  std::string tok;
  std::vector<long> nid; std::vector<double> nx, ny;
  struct El { int type, phys; std::vector<long> nodes; };
  std::vector<El> els;
  while (in >> tok) {
    if (tok == "$MeshFormat") { double v; int a, b; in >> v >> a >> b; }
    else if (tok == "$Nodes") {
      long n; in >> n; nid.resize(n); nx.resize(n); ny.resize(n);
      for (long i = 0; i < n; i++) { double z; in >> nid[i] >> nx[i] >> ny[i] >> z; }
    } else if (tok == "$Elements") {
      long n; in >> n; els.resize(n);
      static const int nn[16] = {0, 2, 3, 4, 4, 8, 6, 5, 3, 6, 9, 10, 27, 18, 14, 1};
      for (long i = 0; i < n; i++) {
        long id; int type, ntags; in >> id >> type >> ntags;
        std::vector<int> tags(ntags);
        for (int t = 0; t < ntags; t++) in >> tags[t];
        els[i].type = type; els[i].phys = ntags > 0 ? tags[0] : 0;
        if (type < 1 || type > 15) throw std::runtime_error("unsupported gmsh element type");
        els[i].nodes.resize(nn[type]);
        for (auto& v : els[i].nodes) in >> v;
      }
    }
  }
  // pass 1: nodes used by triangles; pass 2: compressed indices in node-id order
  std::map<long, int> used;
  for (auto& e : els) if (e.type == 2) for (long v : e.nodes) used[v] = -1;
  std::vector<std::pair<long, long>> order; // (node id, file position)
  for (size_t i = 0; i < nid.size(); i++) if (used.count(nid[i])) order.push_back({nid[i], (long)i});
  std::sort(order.begin(), order.end());
  Mesh m;
  for (size_t k = 0; k < order.size(); k++) {
    used[order[k].first] = (int)k;
    m.x.push_back(nx[order[k].second]); m.y.push_back(ny[order[k].second]);
  }
  for (auto& e : els) {
    if (e.type == 2) for (long v : e.nodes) m.tri.push_back(used[v]);
    if (e.type == 1) {
      m.ba.push_back(used.at(e.nodes[0])); m.bb.push_back(used.at(e.nodes[1]));
      m.bphys.push_back(e.phys);
    }
  }
  m.finalize();
  return m;
}

// Uniform red refinement (our synthetic-mesh rule, DESIGN.md "refinement"):
//  * edges numbered by ascending key (min vertex, max vertex); midpoint of edge k gets
//    vertex index nv_old + k; coordinates 0.5*(x_a + x_b);
//  * triangle t=(a,b,c) -> children 4t..4t+3: (a,mab,mac) (mab,b,mbc) (mac,mbc,c) (mab,mbc,mac);
//  * boundary segment s=(a,b) -> 2s:(a,mab) 2s+1:(mab,b), same physical tag.
inline Mesh refine(const Mesh& m) {
  std::vector<uint64_t> keys; keys.reserve(3 * (size_t)m.nT);
  auto key = [](int a, int b) { return ((uint64_t)std::min(a, b) << 32) | (uint64_t)std::max(a, b); };
  for (int e = 0; e < m.nT; e++)
    for (int f = 0; f < 3; f++) keys.push_back(key(m.tri[3 * e + FACE_V[f][0]], m.tri[3 * e + FACE_V[f][1]]));
  std::sort(keys.begin(), keys.end());
  keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
  auto mid = [&](int a, int b) {
    return m.nv + (int)(std::lower_bound(keys.begin(), keys.end(), key(a, b)) - keys.begin());
  };
  Mesh r;
  r.x = m.x; r.y = m.y;
  r.x.resize(m.nv + keys.size()); r.y.resize(m.nv + keys.size());
  for (size_t k = 0; k < keys.size(); k++) {
    int a = (int)(keys[k] >> 32), b = (int)(keys[k] & 0xffffffffu);
    r.x[m.nv + k] = 0.5 * (m.x[a] + m.x[b]);
    r.y[m.nv + k] = 0.5 * (m.y[a] + m.y[b]);
  }
  r.tri.reserve(12 * (size_t)m.nT);
  for (int e = 0; e < m.nT; e++) {
    int a = m.tri[3 * e], b = m.tri[3 * e + 1], c = m.tri[3 * e + 2];
    int ab = mid(a, b), ac = mid(a, c), bc = mid(b, c);
    int ch[12] = {a, ab, ac, ab, b, bc, ac, bc, c, ab, bc, ac};
    r.tri.insert(r.tri.end(), ch, ch + 12);
  }
  for (int s = 0; s < m.nB; s++) {
    int mm = mid(m.ba[s], m.bb[s]);
    r.ba.push_back(m.ba[s]); r.bb.push_back(mm); r.bphys.push_back(m.bphys[s]);
    r.ba.push_back(mm); r.bb.push_back(m.bb[s]); r.bphys.push_back(m.bphys[s]);
  }
  r.finalize();
  return r;
}

// ----------------------------------------------------------------------------
// Quadrature (dune-geometry QuadratureRules, from memory)
// ----------------------------------------------------------------------------
struct QP { double xi0, xi1, w; };
inline const std::vector<QP>& tri_rule(int order) {
  static const std::vector<QP> p2 = {{4.0 / 6.0, 1.0 / 6.0, 0.5 / 3.0},
                                     {1.0 / 6.0, 4.0 / 6.0, 0.5 / 3.0},
                                     {1.0 / 6.0, 1.0 / 6.0, 0.5 / 3.0}};
  static const std::vector<QP> p3 = {{10.0 / 30.0, 10.0 / 30.0, 0.5 * (-27.0 / 48.0)},
                                     {18.0 / 30.0, 6.0 / 30.0, 0.5 * (25.0 / 48.0)},
                                     {6.0 / 30.0, 18.0 / 30.0, 0.5 * (25.0 / 48.0)},
                                     {6.0 / 30.0, 6.0 / 30.0, 0.5 * (25.0 / 48.0)}};
  static const double a = 0.79742698535308732240, b = 0.10128650732345633880;
  static const double c = 0.05971587178976982045, d = 0.47014206410511508977;
  static const double wb = 0.5 * 0.12593918054482715260, wd = 0.5 * 0.13239415278850618074;
  static const std::vector<QP> p5 = {{1.0 / 3.0, 1.0 / 3.0, 0.5 * 0.225},
                                     {a, b, wb}, {b, a, wb}, {b, b, wb},
                                     {c, d, wd}, {d, c, wd}, {d, d, wd}};
  if (order <= 2) return p2;
  if (order == 3) return p3;
  if (order <= 5) return p5;
  throw std::runtime_error("triangle quadrature order not tabulated");
}
struct QL { double t, w; };
inline const std::vector<QL>& line_rule3() { // 2-pt Gauss-Legendre on [0,1] (order 3)
  static const std::vector<QL> g = {{0.21132486540518711775, 0.5}, {0.78867513459481288225, 0.5}};
  return g;
}
// QuadratureRules<DF,1>::rule(cube, order): Gauss-Legendre on [0,1], n points exact to degree 2n-1 (what the operators get for
// the faces with their `intorder`: 2 points for the drivers' default 3, 3 points for 5)
inline const std::vector<QL>& line_rule(int order) {
  static const std::vector<QL> g3 = {{0.11270166537925831148, 5.0 / 18.0}, {0.5, 8.0 / 18.0}, {0.88729833462074168852, 5.0 / 18.0}};
  if (order <= 3) return line_rule3();
  if (order <= 5) return g3;
  throw std::runtime_error("line quadrature order not tabulated");
}

// ----------------------------------------------------------------------------
// Local operators
// ----------------------------------------------------------------------------
enum Op { OP_PB = 0, OP_POISSON = 1, OP_DIFFUSION = 2, OP_MASS = 3, OP_PNP = 4 };
inline int op_fields(int op) { return op == OP_PNP ? 3 : 1; }

struct OpCtx {
  int op = OP_PB;
  const Sysparams* s = nullptr;
  const Mesh* m = nullptr;
  // coefficient fields (scalar P1 vectors of length nv)
  const double* cp = nullptr;   // poisson_operator.hh:97-100
  const double* cm = nullptr;
  const double* uphi = nullptr; // diffusion_operator.hh:103-105
  double valency = 1.0;
  int intorder = -1; // -1 -> the order the reference drivers end up with
  int order() const {
    if (intorder > 0) return intorder;
    switch (op) {
      case OP_DIFFUSION: return 2; // diffusion_operator.hh:36 default (quirk B8)
      case OP_MASS: return 5;      // DiffusionTOperator(5), instationary_pnp_from_pb_md.hh:362-366
      default: return 3;           // pnp/pb/poisson ctor default intorder_=3
    }
  }
};

struct ElemGeo {
  double x0, y0, x1, y1, x2, y2;
  double jit[2][2]; // J^{-T}
  double detabs;
  double gphi[3][2]; // transformed P1 gradients, gphi[i] = J^{-T} ghat_i
};
inline ElemGeo elem_geo(const Mesh& m, int e) {
  ElemGeo g;
  int a = m.tri[3 * e], b = m.tri[3 * e + 1], c = m.tri[3 * e + 2];
  g.x0 = m.x[a]; g.y0 = m.y[a]; g.x1 = m.x[b]; g.y1 = m.y[b]; g.x2 = m.x[c]; g.y2 = m.y[c];
  // J = [v1-v0 | v2-v0]  (columns);  J^{-T} via the 2x2 cofactor formula
  double j00 = g.x1 - g.x0, j01 = g.x2 - g.x0, j10 = g.y1 - g.y0, j11 = g.y2 - g.y0;
  double det = j00 * j11 - j01 * j10;
  double di = 1.0 / det;
  // J^{-1} = di*[[j11,-j01],[-j10,j00]] ; J^{-T} = transpose
  g.jit[0][0] = j11 * di;  g.jit[0][1] = -j10 * di;
  g.jit[1][0] = -j01 * di; g.jit[1][1] = j00 * di;
  g.detabs = std::fabs(det);
  static const double gh[3][2] = {{-1, -1}, {1, 0}, {0, 1}};
  for (int i = 0; i < 3; i++)
    for (int r = 0; r < 2; r++) { // FieldMatrix::mv : y[r] = 0; y[r] += A[r][c]*x[c]
      double y = 0.0;
      y += g.jit[r][0] * gh[i][0];
      y += g.jit[r][1] * gh[i][1];
      g.gphi[i][r] = y;
    }
  return g;
}

// alpha_volume for element e with local coefficients xl (child-major for PNP:
// [phi0 phi1 phi2 | cp0 cp1 cp2 | cm0 cm1 cm2]); ACCUMULATES into rl.
// sinh of pb_operator.hh:117.  The reference calls std::sinh; this restatement carries its own sinh made of +, -, *, /
// and floor only (|error| <= 4 ulp), because the device kernels must evaluate the SAME bits: NumericalJacobianVolume
// divides residual differences by delta ~ 1e-11, so a last-bit difference between two libm's would appear 1e4 times larger
// in the FD Jacobian and hide real discrepancies.  The CUDA side repeats this operation sequence
// (dune_pnp_b200/csrc/pnp_elem.cuh: pnp_sinh); tests/test_host_logic.py checks the two bit for bit and against libm.
inline double pow2_int(int k) { const uint64_t bits = (uint64_t)(k + 1023) << 52; double d; std::memcpy(&d, &bits, sizeof d); return d; }
inline double sinh_shared(double x) {
  const double a = x < 0.0 ? -x : x;
  double r;
  if (a < 0.35) {
    const double t = a * a;
    double p = 1.0 / 1307674368000.0;
    p = p * t + 1.0 / 6227020800.0;
    p = p * t + 1.0 / 39916800.0;
    p = p * t + 1.0 / 362880.0;
    p = p * t + 1.0 / 5040.0;
    p = p * t + 1.0 / 120.0;
    p = p * t + 1.0 / 6.0;
    r = a + a * (t * p);
  } else {
    const double kf = std::floor(a * 1.44269504088896338700e+00 + 0.5);
    const double s = (a - kf * 6.93147180369123816490e-01) - kf * 1.90821492927058770002e-10;
    double p = 1.0 / 87178291200.0;
    p = p * s + 1.0 / 6227020800.0;
    p = p * s + 1.0 / 479001600.0;
    p = p * s + 1.0 / 39916800.0;
    p = p * s + 1.0 / 3628800.0;
    p = p * s + 1.0 / 362880.0;
    p = p * s + 1.0 / 40320.0;
    p = p * s + 1.0 / 5040.0;
    p = p * s + 1.0 / 720.0;
    p = p * s + 1.0 / 120.0;
    p = p * s + 1.0 / 24.0;
    p = p * s + 1.0 / 6.0;
    p = p * s + 0.5;
    p = p * s + 1.0;
    p = p * s + 1.0;
    const int k = (int)kf;
    const double e = p * pow2_int(k / 2) * pow2_int(k - k / 2);
    r = 0.5 * e - 0.5 / e;
  }
  return x < 0.0 ? -r : r;
}

inline void alpha_volume(const OpCtx& c, int e, const double* xl, double* rl) {
  const Mesh& m = *c.m; const Sysparams& s = *c.s;
  const ElemGeo g = elem_geo(m, e);
  const double PI = s.PI;
  const int* tv = &m.tri[3 * e];
  for (const QP& q : tri_rule(c.order())) {
    double phi[3] = {1.0 - q.xi0 - q.xi1, q.xi0, q.xi1};
    double gy = g.y0 + (g.y1 - g.y0) * q.xi0 + (g.y2 - g.y0) * q.xi1; // geometry().global()[1]
    double factor = q.w * g.detabs;
    switch (c.op) {
      case OP_PNP: { // pnp_operator.hh:98-194
        if (s.cylindrical) factor *= gy * 2 * PI;
        double u[3], gu[3][2];
        for (int k = 0; k < 3; k++) {
          u[k] = 0.0;
          for (int i = 0; i < 3; i++) u[k] += xl[3 * k + i] * phi[i];
          gu[k][0] = gu[k][1] = 0.0;
          for (int i = 0; i < 3; i++) { // FieldVector::axpy
            gu[k][0] += xl[3 * k + i] * g.gphi[i][0];
            gu[k][1] += xl[3 * k + i] * g.gphi[i][1];
          }
        }
        auto dot = [](const double* a, const double* b) { double r = 0.0; r += a[0] * b[0]; r += a[1] * b[1]; return r; };
        for (int i = 0; i < 3; i++) // :167-173
          rl[i] += (dot(gu[0], g.gphi[i]) + 4 * PI * s.l_b * (u[1] - u[2]) * phi[i]) * factor;
        for (int i = 0; i < 3; i++) // :177-183
          rl[3 + i] += (dot(gu[1], g.gphi[i]) - u[1] * dot(gu[0], g.gphi[i])) * factor;
        for (int i = 0; i < 3; i++) // :187-193
          rl[6 + i] += (dot(gu[2], g.gphi[i]) + u[2] * dot(gu[0], g.gphi[i])) * factor;
        break;
      }
      case OP_PB: case OP_POISSON: { // pb_operator.hh:74-120, poisson_operator.hh:74-126
        if (s.cylindrical) factor *= gy * 2 * PI;
        double u = 0.0, gu[2] = {0.0, 0.0};
        for (int i = 0; i < 3; i++) u += xl[i] * phi[i];
        for (int i = 0; i < 3; i++) { gu[0] += xl[i] * g.gphi[i][0]; gu[1] += xl[i] * g.gphi[i][1]; }
        double src;
        if (c.op == OP_PB) src = 8 * PI * s.l_b * s.c0 * sinh_shared(u); // pb_operator.hh:117 (std::sinh there)
        else { // DiscreteGridFunction::evaluate = sum u_i phi_i  (poisson_operator.hh:97-100)
          double cp = 0.0, cm = 0.0;
          for (int i = 0; i < 3; i++) cp += c.cp[tv[i]] * phi[i];
          for (int i = 0; i < 3; i++) cm += c.cm[tv[i]] * phi[i];
          src = 1 * s.l_b * 4 * PI * (cm - cp); // poisson_operator.hh:122
        }
        for (int i = 0; i < 3; i++) {
          double d = 0.0; d += gu[0] * g.gphi[i][0]; d += gu[1] * g.gphi[i][1];
          rl[i] += (d + src * phi[i]) * factor;
        }
        break;
      }
      case OP_DIFFUSION: { // diffusion_operator.hh:64-111 (no cylindrical factor, quirk B6)
        double u = 0.0, gu[2] = {0.0, 0.0}, gP[2] = {0.0, 0.0};
        for (int i = 0; i < 3; i++) u += xl[i] * phi[i];
        for (int i = 0; i < 3; i++) { gu[0] += xl[i] * g.gphi[i][0]; gu[1] += xl[i] * g.gphi[i][1]; }
        // DiscreteGridFunctionGradient: sum_i uphi_i * J^{-T} ghat_i
        for (int i = 0; i < 3; i++) { gP[0] += c.uphi[tv[i]] * g.gphi[i][0]; gP[1] += c.uphi[tv[i]] * g.gphi[i][1]; }
        double a = 0;
        for (int i = 0; i < 3; i++) {
          double d = 0.0; d += gu[0] * g.gphi[i][0]; d += gu[1] * g.gphi[i][1];
          double dP = 0.0; dP += gP[0] * g.gphi[i][0]; dP += gP[1] * g.gphi[i][1];
          rl[i] += (d + u * c.valency * dP + a * u * phi[i]) * factor; // :110
        }
        break;
      }
      case OP_MASS: { // diffusion_toperator.hh:58-72
        double u = 0.0;
        for (int i = 0; i < 3; i++) u += xl[i] * phi[i];
        for (int i = 0; i < 3; i++) rl[i] += u * phi[i] * factor;
        break;
      }
    }
  }
}

// alpha_boundary for face f of element e, boundary segment seg (pnp_operator.hh:199-315,
// pb_operator.hh:126-192, poisson_operator.hh:131-199). diffusion/mass: doAlphaBoundary=false.
inline void alpha_boundary(const OpCtx& c, int e, int f, double* rl) {
  if (c.op == OP_DIFFUSION || c.op == OP_MASS) return;
  const Mesh& m = *c.m; const Sysparams& s = *c.s;
  const int* tv = &m.tri[3 * e];
  int seg = m.fseg[3 * e + f];
  const Surface& sf = s.surfaces.at(m.bphys[seg]);
  int va = tv[FACE_V[f][0]], vb = tv[FACE_V[f][1]];
  double ax = m.x[va], ay = m.y[va], bx = m.x[vb], by = m.y[vb];
  double len = std::sqrt((bx - ax) * (bx - ax) + (by - ay) * (by - ay));
  int nf = op_fields(c.op);
  for (const QL& q : line_rule3()) {
    double l0, l1; // geometryInInside().global(t)
    if (f == 0) { l0 = q.t; l1 = 0.0; } else if (f == 1) { l0 = 0.0; l1 = q.t; } else { l0 = 1.0 - q.t; l1 = q.t; }
    double phi[3] = {1.0 - l0 - l1, l0, l1};
    double gy = ay + q.t * (by - ay);
    double factor = q.w * len;
    if (s.cylindrical) factor *= gy * 2 * s.PI;
    for (int k = 0; k < nf; k++) {
      int comp = (c.op == OP_PNP) ? k : 0; // PB and Poisson are built with the component-0 BCType
      if (sf.btype(comp) == 0) continue;   // isDirichlet -> no flux term
      double j = sf.flux(comp);            // fluxContainer[seg][comp]
      for (int i = 0; i < 3; i++) rl[3 * k + i] += j * phi[i] * factor;
    }
  }
}

// NumericalJacobianVolume: Ae[i][j] (row-major n x n), n = 3*fields
inline void jacobian_volume_fd(const OpCtx& c, int e, const double* xl, double* Ae, double eps) {
  int n = 3 * op_fields(c.op);
  std::vector<double> u(xl, xl + n), down(n, 0.0), up(n);
  alpha_volume(c, e, u.data(), down.data());
  for (int j = 0; j < n; j++) {
    std::fill(up.begin(), up.end(), 0.0);
    double delta = eps * (1.0 + std::fabs(u[j]));
    u[j] += delta;
    alpha_volume(c, e, u.data(), up.data());
    for (int i = 0; i < n; i++) Ae[i * n + j] += (up[i] - down[i]) / delta;
    u[j] = xl[j];
  }
}

// Exact derivative of alpha_volume (NOT in the reference: used to validate the GPU
// analytic-Jacobian fast path and the FD noise floor).
inline void jacobian_volume_exact(const OpCtx& c, int e, const double* xl, double* Ae) {
  const Mesh& m = *c.m; const Sysparams& s = *c.s;
  const ElemGeo g = elem_geo(m, e);
  const int* tv = &m.tri[3 * e];
  int n = 3 * op_fields(c.op);
  double K[3][3];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++)
    K[i][j] = g.gphi[j][0] * g.gphi[i][0] + g.gphi[j][1] * g.gphi[i][1];
  for (const QP& q : tri_rule(c.order())) {
    double phi[3] = {1.0 - q.xi0 - q.xi1, q.xi0, q.xi1};
    double gy = g.y0 + (g.y1 - g.y0) * q.xi0 + (g.y2 - g.y0) * q.xi1;
    double factor = q.w * g.detabs;
    bool cyl = s.cylindrical && c.op != OP_DIFFUSION && c.op != OP_MASS;
    if (cyl) factor *= gy * 2 * s.PI;
    if (c.op == OP_PNP) {
      double u[3] = {0, 0, 0}, gphi_u[2] = {0, 0};
      for (int k = 0; k < 3; k++) for (int i = 0; i < 3; i++) u[k] += xl[3 * k + i] * phi[i];
      for (int i = 0; i < 3; i++) { gphi_u[0] += xl[i] * g.gphi[i][0]; gphi_u[1] += xl[i] * g.gphi[i][1]; }
      double kap = 4 * s.PI * s.l_b;
      for (int i = 0; i < 3; i++) {
        double dPi = gphi_u[0] * g.gphi[i][0] + gphi_u[1] * g.gphi[i][1];
        for (int j = 0; j < 3; j++) {
          Ae[(i)*n + j] += K[i][j] * factor;
          Ae[(i)*n + 3 + j] += kap * phi[j] * phi[i] * factor;
          Ae[(i)*n + 6 + j] -= kap * phi[j] * phi[i] * factor;
          Ae[(3 + i) * n + j] -= u[1] * K[i][j] * factor;
          Ae[(3 + i) * n + 3 + j] += (K[i][j] - phi[j] * dPi) * factor;
          Ae[(6 + i) * n + j] += u[2] * K[i][j] * factor;
          Ae[(6 + i) * n + 6 + j] += (K[i][j] + phi[j] * dPi) * factor;
        }
      }
    } else {
      double u = 0; for (int i = 0; i < 3; i++) u += xl[i] * phi[i];
      double gP[2] = {0, 0};
      if (c.op == OP_DIFFUSION)
        for (int i = 0; i < 3; i++) { gP[0] += c.uphi[tv[i]] * g.gphi[i][0]; gP[1] += c.uphi[tv[i]] * g.gphi[i][1]; }
      for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
        double v = 0;
        switch (c.op) {
          case OP_PB: v = K[i][j] + 8 * s.PI * s.l_b * s.c0 * std::cosh(u) * phi[j] * phi[i]; break;
          case OP_POISSON: v = K[i][j]; break;
          case OP_DIFFUSION: v = K[i][j] + phi[j] * c.valency * (gP[0] * g.gphi[i][0] + gP[1] * g.gphi[i][1]); break;
          case OP_MASS: v = phi[j] * phi[i]; break;
        }
        Ae[i * n + j] += v * factor;
      }
    }
  }
}

// ----------------------------------------------------------------------------
// Grid function space / constraints / pattern / grid operator (App. A.2, A.4, A.5)
// ----------------------------------------------------------------------------
struct CSR {
  int n = 0;
  std::vector<int> rowptr, col;
  std::vector<double> val;
  int find(int r, int c) const {
    auto b = col.begin() + rowptr[r], e = col.begin() + rowptr[r + 1];
    auto it = std::lower_bound(b, e, c);
    return (it != e && *it == c) ? (int)(it - col.begin()) : -1;
  }
  void mv(const double* xx, double* yy) const {
#pragma omp parallel for schedule(static)
    for (int r = 0; r < n; r++) {
      double sum = 0.0;
      for (int k = rowptr[r]; k < rowptr[r + 1]; k++) sum += val[k] * xx[col[k]];
      yy[r] = sum;
    }
  }
};

struct Space {
  const Mesh* m = nullptr;
  const Sysparams* s = nullptr;
  int fields = 1;
  int comp0 = 0;              // BCType component used by a scalar space
  std::vector<char> dirichlet; // per DOF, lexicographic [field][vertex]
  int N() const { return fields * m->nv; }
  int gdof(int field, int v) const { return field * m->nv + v; } // lexicographic mapper (A.4)
};

// constraints(param,gfs,cc,false): face-centre test, both end vertices (A.5; btype.hh:21-53)
inline Space make_space(const Mesh& m, const Sysparams& s, int fields, int comp0 = 0) {
  Space sp; sp.m = &m; sp.s = &s; sp.fields = fields; sp.comp0 = comp0;
  sp.dirichlet.assign(sp.N(), 0);
  for (int e = 0; e < m.nT; e++)
    for (int f = 0; f < 3; f++) {
      int seg = m.fseg[3 * e + f];
      if (seg < 0) continue;
      const Surface& sf = s.surfaces.at(m.bphys[seg]);
      for (int k = 0; k < fields; k++) {
        int comp = fields == 3 ? k : comp0;
        if (sf.btype(comp) == 0)
          for (int l = 0; l < 2; l++) sp.dirichlet[sp.gdof(k, m.tri[3 * e + FACE_V[f][l]])] = 1;
      }
    }
  return sp;
}

// FullVolumePattern + ISTLBCRSMatrixBackend<1,1>, PDELab-1.1 add_entry: links whose row or
// column is constrained are dropped, constrained rows keep their diagonal (A.4).
inline CSR make_pattern_sets(const Space& sp) { // literal restatement (std::set per row); make_pattern() is its fast equal
  const Mesh& m = *sp.m;
  int N = sp.N(), nf = sp.fields;
  std::vector<std::set<int>> rows(N);
  for (int e = 0; e < m.nT; e++)
    for (int ki = 0; ki < nf; ki++) for (int i = 0; i < 3; i++) {
      int gi = sp.gdof(ki, m.tri[3 * e + i]);
      if (sp.dirichlet[gi]) continue;
      for (int kj = 0; kj < nf; kj++) for (int j = 0; j < 3; j++) {
        int gj = sp.gdof(kj, m.tri[3 * e + j]);
        if (!sp.dirichlet[gj]) rows[gi].insert(gj);
      }
    }
  for (int d = 0; d < N; d++) if (sp.dirichlet[d]) rows[d].insert(d);
  CSR A; A.n = N; A.rowptr.assign(N + 1, 0);
  for (int r = 0; r < N; r++) A.rowptr[r + 1] = A.rowptr[r] + (int)rows[r].size();
  A.col.reserve(A.rowptr[N]);
  for (int r = 0; r < N; r++) A.col.insert(A.col.end(), rows[r].begin(), rows[r].end());
  A.val.assign(A.col.size(), 0.0);
  return A;
}

inline CSR make_pattern(const Space& sp) {
  const Mesh& m = *sp.m;
  const int N = sp.N(), nf = sp.fields, nv = m.nv;
  // vertex -> sorted vertices it shares an element with (itself included): the scalar P1 pattern.  The dof pattern
  // follows from it because all fields of a vertex couple to all fields of its neighbours (FullVolumePattern) and the
  // lexicographic mapper orders columns by (field, vertex).
  std::vector<int> cnt(nv + 1, 0);
  for (int e = 0; e < m.nT; e++) for (int i = 0; i < 3; i++) cnt[m.tri[3 * e + i] + 1] += 3;
  for (int v = 0; v < nv; v++) cnt[v + 1] += cnt[v];
  std::vector<int> raw(cnt[nv]), fill(cnt.begin(), cnt.end() - 1);
  for (int e = 0; e < m.nT; e++) for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) raw[fill[m.tri[3 * e + i]]++] = m.tri[3 * e + j];
  std::vector<int> nptr(nv + 1, 0), nbr; nbr.reserve(raw.size() / 2);
  for (int v = 0; v < nv; v++) {
    auto b = raw.begin() + cnt[v], e = raw.begin() + cnt[v + 1];
    std::sort(b, e);
    e = std::unique(b, e);
    nbr.insert(nbr.end(), b, e);
    nptr[v + 1] = (int)nbr.size();
  }
  CSR A; A.n = N; A.rowptr.assign(N + 1, 0);
  for (int ki = 0; ki < nf; ki++) for (int v = 0; v < nv; v++) {
    const int gi = sp.gdof(ki, v);
    int len = 0;
    if (sp.dirichlet[gi]) len = 1; // constrained rows keep their diagonal only
    else for (int kj = 0; kj < nf; kj++) for (int t = nptr[v]; t < nptr[v + 1]; t++) len += !sp.dirichlet[sp.gdof(kj, nbr[t])];
    A.rowptr[gi + 1] = len;
  }
  for (int r = 0; r < N; r++) A.rowptr[r + 1] += A.rowptr[r];
  A.col.resize(A.rowptr[N]);
  for (int ki = 0; ki < nf; ki++) for (int v = 0; v < nv; v++) {
    const int gi = sp.gdof(ki, v);
    int o = A.rowptr[gi];
    if (sp.dirichlet[gi]) { A.col[o] = gi; continue; }
    for (int kj = 0; kj < nf; kj++) for (int t = nptr[v]; t < nptr[v + 1]; t++) {
      const int gj = sp.gdof(kj, nbr[t]);
      if (!sp.dirichlet[gj]) A.col[o++] = gj;
    }
  }
  A.val.assign(A.col.size(), 0.0);
  return A;
}

// GridOperator::residual (A.2).  If absr != nullptr it receives sum |contribution| per DOF
// (the scale against which the 1e-12 parity tolerance is taken).
inline void residual(const Space& sp, const OpCtx& c, const double* u, double* r, double* absr = nullptr) {
  const Mesh& m = *sp.m;
  int nf = sp.fields, n = 3 * nf, N = sp.N();
  std::fill(r, r + N, 0.0);
  if (absr) std::fill(absr, absr + N, 0.0);
  std::vector<double> xl(n), rl(n);
  for (int e = 0; e < m.nT; e++) {
    for (int k = 0; k < nf; k++) for (int i = 0; i < 3; i++) xl[3 * k + i] = u[sp.gdof(k, m.tri[3 * e + i])];
    std::fill(rl.begin(), rl.end(), 0.0);
    alpha_volume(c, e, xl.data(), rl.data());
    for (int fi = 0; fi < 3; fi++) {
      int f = FACE_ITER[fi];
      if (m.fseg[3 * e + f] >= 0) alpha_boundary(c, e, f, rl.data());
    }
    for (int k = 0; k < nf; k++) for (int i = 0; i < 3; i++) {
      int g = sp.gdof(k, m.tri[3 * e + i]);
      r[g] += rl[3 * k + i];
      if (absr) absr[g] += std::fabs(rl[3 * k + i]);
    }
  }
  for (int d = 0; d < N; d++) if (sp.dirichlet[d]) r[d] = 0.0; // constrain_residual
}

// GridOperator::jacobian (A.2): mode 0 = NumericalJacobianVolume (reference), 1 = exact derivative.
inline void jacobian(const Space& sp, const OpCtx& c, const double* u, CSR& A, int mode = 0,
                     double eps = 1e-11, std::vector<double>* absA = nullptr) {
  const Mesh& m = *sp.m;
  int nf = sp.fields, n = 3 * nf, N = sp.N();
  std::fill(A.val.begin(), A.val.end(), 0.0);
  if (absA) absA->assign(A.val.size(), 0.0);
  std::vector<double> xl(n), Ae(n * n);
  for (int e = 0; e < m.nT; e++) {
    for (int k = 0; k < nf; k++) for (int i = 0; i < 3; i++) xl[3 * k + i] = u[sp.gdof(k, m.tri[3 * e + i])];
    std::fill(Ae.begin(), Ae.end(), 0.0);
    if (mode == 0) jacobian_volume_fd(c, e, xl.data(), Ae.data(), eps);
    else jacobian_volume_exact(c, e, xl.data(), Ae.data());
    // NumericalJacobianBoundary: boundary terms do not depend on x -> exact zeros (A.3)
    for (int ki = 0; ki < nf; ki++) for (int i = 0; i < 3; i++) {
      int gi = sp.gdof(ki, m.tri[3 * e + i]);
      if (sp.dirichlet[gi]) continue;
      for (int kj = 0; kj < nf; kj++) for (int j = 0; j < 3; j++) {
        int gj = sp.gdof(kj, m.tri[3 * e + j]);
        if (sp.dirichlet[gj]) continue;
        int slot = A.find(gi, gj);
        A.val[slot] += Ae[(3 * ki + i) * n + 3 * kj + j];
        if (absA) (*absA)[slot] += std::fabs(Ae[(3 * ki + i) * n + 3 * kj + j]);
      }
    }
  }
  for (int d = 0; d < N; d++) if (sp.dirichlet[d]) A.val[A.find(d, d)] = 1.0; // set_trivial_row
}

// ----------------------------------------------------------------------------
// Multi-core variants of the grid operator for the CPU baseline ("(ii) all host cores via
// OpenMP").  Two phases, no atomics: (1) element loop in parallel into per-element blocks, (2) row loop in parallel,
// every dof summing its elements' contributions in ascending element order -- the order of the sequential scatter, so
// the results equal residual() / jacobian() bit for bit whatever the thread count (tests/test_oracle.py).
// ----------------------------------------------------------------------------
struct Incidence { std::vector<int> ptr, elem; }; // vertex -> (element, local index) pairs, elements ascending: elem = 4*e + i
inline Incidence vertex_elements(const Mesh& m) {
  Incidence I; I.ptr.assign(m.nv + 1, 0);
  for (int e = 0; e < m.nT; e++) for (int i = 0; i < 3; i++) I.ptr[m.tri[3 * e + i] + 1]++;
  for (int v = 0; v < m.nv; v++) I.ptr[v + 1] += I.ptr[v];
  I.elem.resize(I.ptr[m.nv]);
  std::vector<int> fill(I.ptr.begin(), I.ptr.end() - 1);
  for (int e = 0; e < m.nT; e++) for (int i = 0; i < 3; i++) I.elem[fill[m.tri[3 * e + i]]++] = 4 * e + i;
  return I;
}
inline void residual_par(const Space& sp, const OpCtx& c, const Incidence& I, const double* u, double* r, std::vector<double>& scratch) {
  const Mesh& m = *sp.m;
  const int nf = sp.fields, n = 3 * nf;
  scratch.resize((size_t)n * m.nT);
#pragma omp parallel for schedule(static)
  for (int e = 0; e < m.nT; e++) {
    double xl[9], *rl = &scratch[(size_t)n * e];
    for (int k = 0; k < nf; k++) for (int i = 0; i < 3; i++) xl[3 * k + i] = u[sp.gdof(k, m.tri[3 * e + i])];
    for (int q = 0; q < n; q++) rl[q] = 0.0;
    alpha_volume(c, e, xl, rl);
    for (int fi = 0; fi < 3; fi++) { const int f = FACE_ITER[fi]; if (m.fseg[3 * e + f] >= 0) alpha_boundary(c, e, f, rl); }
  }
#pragma omp parallel for schedule(static)
  for (int v = 0; v < m.nv; v++)
    for (int k = 0; k < nf; k++) {
      const int g = sp.gdof(k, v);
      double sum = 0.0;
      for (int t = I.ptr[v]; t < I.ptr[v + 1]; t++) sum += scratch[(size_t)n * (I.elem[t] >> 2) + 3 * k + (I.elem[t] & 3)];
      r[g] = sp.dirichlet[g] ? 0.0 : sum;
    }
}
inline void jacobian_par(const Space& sp, const OpCtx& c, const Incidence& I, const double* u, CSR& A, int mode, double eps,
                         std::vector<double>& scratch) {
  const Mesh& m = *sp.m;
  const int nf = sp.fields, n = 3 * nf;
  scratch.resize((size_t)n * n * m.nT);
#pragma omp parallel for schedule(static)
  for (int e = 0; e < m.nT; e++) {
    double xl[9], *Ae = &scratch[(size_t)n * n * e];
    for (int k = 0; k < nf; k++) for (int i = 0; i < 3; i++) xl[3 * k + i] = u[sp.gdof(k, m.tri[3 * e + i])];
    for (int q = 0; q < n * n; q++) Ae[q] = 0.0;
    if (mode == 0) jacobian_volume_fd(c, e, xl, Ae, eps); else jacobian_volume_exact(c, e, xl, Ae);
  }
#pragma omp parallel for schedule(static)
  for (int v = 0; v < m.nv; v++)
    for (int ki = 0; ki < nf; ki++) {
      const int gi = sp.gdof(ki, v);
      for (int k = A.rowptr[gi]; k < A.rowptr[gi + 1]; k++) A.val[k] = 0.0;
      if (sp.dirichlet[gi]) { A.val[A.find(gi, gi)] = 1.0; continue; }
      for (int t = I.ptr[v]; t < I.ptr[v + 1]; t++) {
        const int e = I.elem[t] >> 2, i = I.elem[t] & 3;
        const double* Ae = &scratch[(size_t)n * n * e];
        for (int kj = 0; kj < nf; kj++) for (int j = 0; j < 3; j++) {
          const int gj = sp.gdof(kj, m.tri[3 * e + j]);
          if (!sp.dirichlet[gj]) A.val[A.find(gi, gj)] += Ae[(3 * ki + i) * n + 3 * kj + j];
        }
      }
    }
}

// ----------------------------------------------------------------------------
// BCExtension + interpolate (dirichlet_bc.hh:21-123; App. A.6 "interpolate")
// ----------------------------------------------------------------------------
inline bool global_on_intersection(const Mesh& m, double px, double py, int e, int f) { // :21-33
  int a = m.tri[3 * e + FACE_V[f][0]], b = m.tri[3 * e + FACE_V[f][1]];
  double c0x = m.x[a], c0y = m.y[a];
  double vx = m.x[b] - c0x, vy = m.y[b] - c0y;
  double nrm = std::sqrt(vx * vx + vy * vy);
  vx /= nrm; vy /= nrm;
  double dx = px - c0x, dy = py - c0y;
  double t = dx * vx + dy * vy;
  double ex = vx * t - dx, ey = vy * t - dy;
  return std::sqrt(ex * ex + ey * ey) < 1e-9;
}

// value of BCExtension<component>::evaluate at local vertex i of element e; pb = PB potential (nv)
inline double bcext_eval(const Mesh& m, const Sysparams& s, int comp, const double* pb, int e, int i) {
  int v = m.tri[3 * e + i];
  double px = m.x[v], py = m.y[v];
  int pg = -1;
  // bctype() always ends at minusDiffusionBtype (missing breaks, :40-51; quirk B2)
  auto sticky = [&](int g) { return s.surfaces.at(g).minusDiffusionBtype == 0; };
  for (int fi = 0; fi < 3; fi++) {
    int f = FACE_ITER[fi];
    if (m.fseg[3 * e + f] >= 0) {
      if (global_on_intersection(m, px, py, e, f))
        if (pg == -1 || !sticky(pg)) pg = m.bphys[m.fseg[3 * e + f]];
    } else {
      int o = m.nbr[3 * e + f];
      for (int gi = 0; gi < 3; gi++) {
        int f2 = FACE_ITER[gi];
        if (m.fseg[3 * o + f2] >= 0 && global_on_intersection(m, px, py, o, f2))
          if (pg == -1 || !sticky(pg)) pg = m.bphys[m.fseg[3 * o + f2]];
      }
    }
  }
  if (pg > -1 && s.surfaces.at(pg).btype(comp) == 0) return s.surfaces[pg].dirichlet(comp);
  double yv = pb ? pb[v] : 0.0; // phiDGF.evaluate at a vertex = nodal value
  if (comp == 0) return yv;
  if (comp == 1) return s.c0 * std::exp(-yv);
  return s.c0 * std::exp(+yv);
}

// interpolate(bce, gfs, u): element loop, later elements overwrite earlier ones
inline void interpolate_bcext(const Mesh& m, const Sysparams& s, int comp, const double* pb, double* u) {
  for (int e = 0; e < m.nT; e++)
    for (int i = 0; i < 3; i++) u[m.tri[3 * e + i]] = bcext_eval(m, s, comp, pb, e, i);
}

// ----------------------------------------------------------------------------
// ISTL solvers (App. A.7)
// ----------------------------------------------------------------------------
enum Prec { PREC_NONE = 0, PREC_JACOBI = 1, PREC_SSOR = 2, PREC_ILU0 = 3 };
struct LinResult { bool converged = false; int iterations = 0; double reduction = 1, conv_rate = 1; int status = 0; };

// [UPSTREAM dune-istl 2.2 ilu.hh] bilu0_decomposition: in-place ILU(0) on the pattern of A, rows ascending; for every
// lower entry (i,j), j ascending: a_ij *= a_jj^-1 (finished rows keep their diagonal INVERTED), then a_ik -= a_ij a_jk
// for the k > j present in both rows; finally a_ii is inverted.  (ILU0 is not one of the reference's five backends; the
// north star asks for it next to SSOR, so it is restated the way ISTL's SeqILU0 would run on the PDELab matrix.)
inline void ilu0_decompose(CSR& A) {
  for (int i = 0; i < A.n; i++) {
    for (int kj = A.rowptr[i]; kj < A.rowptr[i + 1] && A.col[kj] < i; kj++) {
      const int j = A.col[kj];
      A.val[kj] *= A.val[A.find(j, j)];
      int ki = kj + 1;
      for (int kk = A.find(j, j) + 1; kk < A.rowptr[j + 1]; kk++) {
        while (ki < A.rowptr[i + 1] && A.col[ki] < A.col[kk]) ki++;
        if (ki < A.rowptr[i + 1] && A.col[ki] == A.col[kk]) A.val[ki] -= A.val[kj] * A.val[kk];
      }
    }
    const int d = A.find(i, i);
    A.val[d] = 1.0 / A.val[d];
  }
}
// bilu_backsolve: L v = d (unit lower), then U v = v with the inverted diagonal
inline void ilu0_backsolve(const CSR& LU, double* v, const double* d) {
  for (int i = 0; i < LU.n; i++) {
    double sum = d[i];
    for (int k = LU.rowptr[i]; k < LU.rowptr[i + 1] && LU.col[k] < i; k++) sum -= LU.val[k] * v[LU.col[k]];
    v[i] = sum;
  }
  for (int i = LU.n - 1; i >= 0; i--) {
    double sum = v[i];
    const int dg = LU.find(i, i);
    for (int k = dg + 1; k < LU.rowptr[i + 1]; k++) sum -= LU.val[k] * v[LU.col[k]];
    v[i] = sum * LU.val[dg];
  }
}

struct Preconditioner {
  const CSR& A; int prec, steps; CSR LU;
  Preconditioner(const CSR& A_, int prec_, int steps_) : A(A_), prec(prec_), steps(steps_) {
    if (prec == PREC_ILU0) { LU = A; ilu0_decompose(LU); }
  }
  // v = M^-1 d  (ISTL: v = 0; prec.apply(v, d))
  void apply(double* v, const double* d) const {
    int N = A.n;
    std::fill(v, v + N, 0.0);
    if (prec == PREC_NONE) { std::copy(d, d + N, v); return; } // Richardson: v = d
    if (prec == PREC_JACOBI) { for (int i = 0; i < N; i++) v[i] = d[i] / A.val[A.find(i, i)]; return; }
    if (prec == PREC_ILU0) { ilu0_backsolve(LU, v, d); return; }
    for (int s = 0; s < steps; s++) { // SeqSSOR, w = 1
      for (int i = 0; i < N; i++) {
        double sum = d[i], dii = 0;
        for (int k = A.rowptr[i]; k < A.rowptr[i + 1]; k++) { if (A.col[k] == i) dii = A.val[k]; sum -= A.val[k] * v[A.col[k]]; }
        v[i] += sum / dii;
      }
      for (int i = N - 1; i >= 0; i--) {
        double sum = d[i], dii = 0;
        for (int k = A.rowptr[i]; k < A.rowptr[i + 1]; k++) { if (A.col[k] == i) dii = A.val[k]; sum -= A.val[k] * v[A.col[k]]; }
        v[i] += sum / dii;
      }
    }
  }
};
inline void prec_apply(const CSR& A, int prec, int steps, double* v, const double* d) { Preconditioner(A, prec, steps).apply(v, d); }
// (one thread -- the default, and what every parity test runs with -- gives the plain sequential sum)
inline double dot(int N, const double* a, const double* b) {
  double s = 0;
#pragma omp parallel for schedule(static) reduction(+ : s)
  for (int i = 0; i < N; i++) s += a[i] * b[i];
  return s;
}
inline double nrm2(int N, const double* a) { return std::sqrt(dot(N, a, a)); }

// BiCGSTABSolver::apply(x,b): b is overwritten with the residual
inline LinResult bicgstab(const CSR& A, double* x, double* b, double reduction, int maxit, int prec, int steps = 1) {
  int N = A.n; LinResult res; const Preconditioner P(A, prec, steps);
  std::vector<double> r(N), rt(N), p(N, 0.0), v(N, 0.0), t(N), y(N), tmp(N);
  A.mv(x, tmp.data());
#pragma omp parallel for schedule(static)
  for (int i = 0; i < N; i++) r[i] = b[i] - tmp[i];
  rt = r;
  double rho = 1, alpha = 1, omega = 1, rho_new, beta, h;
  double norm0 = nrm2(N, r.data()), norm = norm0;
  double it = 0;
  auto finish = [&](bool conv) {
    res.converged = conv; res.iterations = (int)std::ceil(it);
    res.reduction = norm0 > 0 ? norm / norm0 : 0; res.conv_rate = it > 0 ? std::pow(res.reduction, 1.0 / it) : 0;
    std::copy(r.begin(), r.end(), b); return res;
  };
  if (norm < reduction * norm0 || norm < 1e-30) return finish(true);
  for (it = 0.5; it < maxit; it += 0.5) {
    rho_new = dot(N, rt.data(), r.data());
    if (std::fabs(rho) <= 1e-80 || std::fabs(omega) <= 1e-80) { res.status = 2; return finish(false); }
    if (it < 1) p = r;
    else {
      beta = (rho_new / rho) * (alpha / omega);
#pragma omp parallel for schedule(static)
      for (int i = 0; i < N; i++) { p[i] += -omega * v[i]; p[i] *= beta; p[i] += r[i]; }
    }
    P.apply(y.data(), p.data());
    A.mv(y.data(), v.data());
    h = dot(N, rt.data(), v.data());
    if (std::fabs(h) < 1e-80) { res.status = 2; return finish(false); }
    alpha = rho_new / h;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < N; i++) { x[i] += alpha * y[i]; r[i] += -alpha * v[i]; }
    norm = nrm2(N, r.data());
    if (norm < reduction * norm0) return finish(true);
    it += 0.5;
    P.apply(y.data(), r.data());
    A.mv(y.data(), t.data());
    omega = dot(N, t.data(), r.data()) / dot(N, t.data(), t.data());
#pragma omp parallel for schedule(static)
    for (int i = 0; i < N; i++) { x[i] += omega * y[i]; r[i] += -omega * t[i]; }
    rho = rho_new;
    norm = nrm2(N, r.data());
    if (norm < reduction * norm0 || norm < 1e-30) return finish(true);
  }
  it = maxit;
  return finish(false);
}

// CGSolver::apply(x,b)
inline LinResult cg(const CSR& A, double* x, double* b, double reduction, int maxit, int prec, int steps = 1) {
  int N = A.n; LinResult res; const Preconditioner P(A, prec, steps);
  std::vector<double> p(N), q(N), tmp(N);
  A.mv(x, tmp.data());
  for (int i = 0; i < N; i++) b[i] -= tmp[i];
  double def0 = nrm2(N, b), def = def0;
  int i = 0;
  auto finish = [&](bool conv) {
    res.converged = conv; res.iterations = i; res.reduction = def0 > 0 ? def / def0 : 0;
    res.conv_rate = i > 0 ? std::pow(res.reduction, 1.0 / i) : 0; return res;
  };
  if (def0 < 1e-30) return finish(true);
  P.apply(p.data(), b);
  double rholast = dot(N, p.data(), b);
  for (i = 1; i <= maxit; i++) {
    A.mv(p.data(), q.data());
    double alpha = dot(N, p.data(), q.data());
    double lambda = rholast / alpha;
    for (int k = 0; k < N; k++) { x[k] += lambda * p[k]; b[k] -= lambda * q[k]; }
    def = nrm2(N, b);
    if (def < def0 * reduction || def < 1e-30) return finish(true);
    P.apply(q.data(), b);
    double rho = dot(N, q.data(), b);
    double beta = rho / rholast;
    for (int k = 0; k < N; k++) p[k] = beta * p[k] + q[k];
    rholast = rho;
  }
  i = maxit;
  return finish(false);
}

enum Solver { SOLVER_BCGS = 0, SOLVER_CG = 1 };
inline LinResult lin_solve(int solver, const CSR& A, double* x, double* b, double red, int maxit, int prec, int steps) {
  return solver == SOLVER_CG ? cg(A, x, b, red, maxit, prec, steps) : bicgstab(A, x, b, red, maxit, prec, steps);
}

// ----------------------------------------------------------------------------
// PDELab Newton (App. A.1)
// ----------------------------------------------------------------------------
struct NewtonOpts {
  double reduction = 1e-8, abs_limit = 1e-12, min_linear_reduction = 1e-3, reassemble_threshold = 0.0;
  int maxit = 40, ls_maxit = 10;
  double damping = 0.5;
  int jac_mode = 0; double fd_eps = 1e-11;
  int solver = SOLVER_BCGS, prec = PREC_NONE, prec_steps = 1, lin_maxit = 5000;
  int verbosity = 0;
  int line_search = 0; // 0 hackbuschReuskenAcceptBest (what the reference drivers select), 1 noLineSearch, 2 hackbuschReusken
};
struct NewtonResult {
  int status = 0; // 0 ok, 1 not converged, 2 linear solver, 3 line search, 4 nan
  bool converged = false; int iterations = 0;
  double first_defect = 0, defect = 0, reduction = 1;
  int total_linear_iterations = 0, total_ls_trials = 0, jacobian_assemblies = 0, residual_assemblies = 0;
  std::vector<double> defect_history; std::vector<int> lin_iter_history;
};

// the Newton loop itself, for any space: residual_fn(u, r), jacobian_fn(u, A) on the pattern A
template <class ResidualFn, class JacobianFn>
inline NewtonResult newton_core(int N, CSR A, ResidualFn residual_fn, JacobianFn jacobian_fn, double* u, const NewtonOpts& o) {
  NewtonResult R;
  std::vector<double> r(N), z(N), prev_u(N);
  auto defect = [&]() { residual_fn(u, r.data()); R.residual_assemblies++; return nrm2(N, r.data()); };
  R.defect = defect(); R.first_defect = R.defect; double prev_defect = R.defect;
  R.defect_history.push_back(R.defect);
  if (!std::isfinite(R.defect)) { R.status = 4; return R; }
  // newton.hh: a NewtonLineSearchError on a matrix that was not reassembled sets reassemble_threshold = 0 and retries
  double reassemble_threshold = o.reassemble_threshold; int ls_retries = 0;
  while (true) {
    R.converged = R.defect < o.abs_limit || R.defect < R.first_defect * o.reduction;
    if (R.converged) break;
    if (R.iterations >= o.maxit) { R.status = 1; break; }
    // prepare_step
    bool reassembled = false;
    if (R.defect / prev_defect > reassemble_threshold || R.jacobian_assemblies == 0) {
      jacobian_fn(u, A); R.jacobian_assemblies++; reassembled = true;
    }
    double stop_defect = std::max(R.first_defect * o.reduction, o.abs_limit);
    double linear_reduction;
    if (stop_defect / (10 * R.defect) > R.defect * R.defect / (prev_defect * prev_defect))
      linear_reduction = stop_defect / (10 * R.defect);
    else
      linear_reduction = std::min(o.min_linear_reduction, R.defect * R.defect / (prev_defect * prev_defect));
    prev_defect = R.defect;
    // linearSolve
    std::fill(z.begin(), z.end(), 0.0);
    LinResult lr = lin_solve(o.solver, A, z.data(), r.data(), linear_reduction, o.lin_maxit, o.prec, o.prec_steps);
    R.total_linear_iterations += lr.iterations; R.lin_iter_history.push_back(lr.iterations);
    if (!lr.converged) { R.status = 2; break; }
    // line_search: hackbuschReuskenAcceptBest
    double lambda = 1.0, best_lambda = 0.0, best_defect = R.defect;
    std::copy(u, u + N, prev_u.begin());
    int i = 0; bool ls_fail = false;
    while (true) {
      for (int k = 0; k < N; k++) u[k] += -lambda * z[k];
      R.defect = defect(); R.total_ls_trials++;
      if (o.line_search == 1) break; // noLineSearch: full step, whatever the defect does
      bool finite = std::isfinite(R.defect);
      if (finite && R.defect <= (1.0 - lambda / 4) * prev_defect) break;
      if (finite && R.defect < best_defect) { best_defect = R.defect; best_lambda = lambda; }
      if (++i >= o.ls_maxit) {
        if (best_lambda == 0.0 || o.line_search == 2) { std::copy(prev_u.begin(), prev_u.end(), u); R.defect = defect(); ls_fail = true; break; }
        if (best_lambda != lambda) {
          std::copy(prev_u.begin(), prev_u.end(), u);
          for (int k = 0; k < N; k++) u[k] += -best_lambda * z[k];
          R.defect = defect();
        }
        break;
      }
      lambda *= o.damping;
      std::copy(prev_u.begin(), prev_u.end(), u);
    }
    if (ls_fail) {
      if (reassembled || ++ls_retries > o.maxit) { R.status = 3; break; }
      reassemble_threshold = 0.0; continue;
    }
    reassemble_threshold = o.reassemble_threshold;
    R.reduction = R.defect / R.first_defect;
    R.iterations++;
    R.defect_history.push_back(R.defect);
    (void)reassembled;
  }
  return R;
}
inline NewtonResult newton(const Space& sp, const OpCtx& c, double* u, const NewtonOpts& o) {
  return newton_core(sp.N(), make_pattern(sp), [&](const double* uu, double* r) { residual(sp, c, uu, r); },
                     [&](const double* uu, CSR& A) { jacobian(sp, c, uu, A, o.jac_mode, o.fd_eps); }, u, o);
}

// StationaryLinearProblemSolver::apply (App. A.9): one Newton step without line search
inline LinResult slp_apply(const Space& sp, const OpCtx& c, double* u, double reduction, int solver, int prec,
                           int steps, int maxit, int jac_mode = 0, double eps = 1e-11) {
  int N = sp.N();
  CSR A = make_pattern(sp);
  jacobian(sp, c, u, A, jac_mode, eps);
  std::vector<double> r(N), z(N, 0.0);
  residual(sp, c, u, r.data());
  LinResult lr = lin_solve(solver, A, z.data(), r.data(), reduction, maxit, prec, steps);
  for (int i = 0; i < N; i++) u[i] -= z[i];
  return lr;
}

// ----------------------------------------------------------------------------
// OneStepMethod + OneStepGridOperator (App. A.9) [UPSTREAM dune-pdelab 1.1 instationary/onestep.hh,
// gridoperator/onestep.hh, timesteppingparameterinterface]; call sites
// instationary_pnp_from_pb_md.hh:368-391 (construction) and :421-425 (apply per time step).
// ----------------------------------------------------------------------------
struct TimeMethod { // TimeSteppingParameterInterface: s stages, d[0..s], a[r][0..s], b[r][0..s] (r = 0..s-1)
  int s = 1; std::vector<double> d; std::vector<std::vector<double>> a, b;
};
inline TimeMethod alexander2() { // Alexander2Parameter
  const double alpha = 1.0 - 0.5 * std::sqrt(2.0);
  TimeMethod m; m.s = 2; m.d = {0.0, alpha, 1.0};
  m.a = {{-1.0, 1.0, 0.0}, {-1.0, 0.0, 1.0}};
  m.b = {{0.0, alpha, 0.0}, {0.0, 1.0 - alpha, alpha}};
  return m;
}
inline TimeMethod implicit_euler() { // ImplicitEulerParameter
  TimeMethod m; m.s = 1; m.d = {0.0, 1.0}; m.a = {{-1.0, 1.0}}; m.b = {{0.0, 1.0}};
  return m;
}
struct OneStepResult { std::vector<LinResult> stage; };
// xold -> xnew over one step of size dt.  c0: spatial operator (DiffusionOperator), c1: temporal operator
// (DiffusionTOperator); g: Dirichlet values (interpolate(f, ...) at the constrained dofs; the reference's boundary
// function cpB is time independent); the stage problems are solved by StationaryLinearProblemSolver (one linear solve).
// (the method itself, for any discrete space: N dofs, their Dirichlet flags, the BCRS pattern, residual and Jacobian of the
// spatial (0) and the temporal (1) operator)
template <class Res0, class Res1, class Jac0, class Jac1>
inline OneStepResult onestep_core(int N, const std::vector<char>& dirichlet, const CSR& pattern, Res0 residual0, Res1 residual1,
                                  Jac0 jacobian0, Jac1 jacobian1, const TimeMethod& tm, double dt, const double* xold,
                                  const double* g, double* xnew, double reduction, int solver, int prec, int steps, int maxit) {
  OneStepResult out;
  std::vector<std::vector<double>> x(tm.s + 1, std::vector<double>(N));
  std::copy(xold, xold + N, x[0].begin());
  CSR A = pattern, B = pattern;
  std::vector<double> cst(N), r0(N), r1(N), res(N), z(N);
  for (int r = 1; r <= tm.s; r++) {
    // preStage: constant part of the residual from the earlier stages
    std::fill(cst.begin(), cst.end(), 0.0);
    for (int i = 0; i < r; i++) {
      const double ai = tm.a[r - 1][i], bi = tm.b[r - 1][i];
      if (std::fabs(ai) > 1e-6) { residual1(x[i].data(), r1.data()); for (int k = 0; k < N; k++) cst[k] += ai * r1[k]; }
      if (std::fabs(bi) > 1e-6) { residual0(x[i].data(), r0.data()); for (int k = 0; k < N; k++) cst[k] += bi * dt * r0[k]; }
    }
    // initial guess: previous stage; Dirichlet dofs from the boundary function, the rest copied (copy_nonconstrained_dofs)
    std::vector<double>& xn = x[r];
    xn = x[r - 1];
    for (int k = 0; k < N; k++) if (dirichlet[k]) xn[k] = g[k];
    // StationaryLinearProblemSolver on the one-step operator
    const double ar = tm.a[r - 1][r], br = tm.b[r - 1][r];
    jacobian1(xn.data(), A);
    jacobian0(xn.data(), B);
    for (size_t k = 0; k < A.val.size(); k++) A.val[k] = ar * A.val[k] + br * dt * B.val[k];
    for (int k = 0; k < N; k++) if (dirichlet[k]) A.val[A.find(k, k)] = 1.0; // constrained rows are trivial
    residual1(xn.data(), r1.data());
    residual0(xn.data(), r0.data());
    for (int k = 0; k < N; k++) res[k] = dirichlet[k] ? 0.0 : cst[k] + ar * r1[k] + br * dt * r0[k];
    std::fill(z.begin(), z.end(), 0.0);
    out.stage.push_back(lin_solve(solver, A, z.data(), res.data(), reduction, maxit, prec, steps));
    for (int k = 0; k < N; k++) xn[k] -= z[k];
  }
  std::copy(x[tm.s].begin(), x[tm.s].end(), xnew);
  return out;
}
inline OneStepResult onestep_apply(const Space& sp, const OpCtx& c0, const OpCtx& c1, const TimeMethod& tm, double dt,
                                   const double* xold, const double* g, double* xnew, double reduction, int solver,
                                   int prec, int steps, int maxit, int jac_mode = 0, double eps = 1e-11) {
  return onestep_core(sp.N(), sp.dirichlet, make_pattern(sp),
                      [&](const double* x, double* r) { residual(sp, c0, x, r); }, [&](const double* x, double* r) { residual(sp, c1, x, r); },
                      [&](const double* x, CSR& A) { jacobian(sp, c0, x, A, jac_mode, eps); },
                      [&](const double* x, CSR& A) { jacobian(sp, c1, x, A, jac_mode, eps); }, tm, dt, xold, g, xnew, reduction, solver,
                      prec, steps, maxit);
}

// ----------------------------------------------------------------------------
// Diagnostics of the time loop
// ----------------------------------------------------------------------------
// calcIonFlux (ionFlux.hh:8-96): element loop, intersections in iteration order, boundary intersections only (:74);
// everything is evaluated at the intersection centre mapped into the element (:51-53).  ip, im: n_surfaces entries
// (component 0 of the reference's FieldVectors).
inline void ion_flux(const Mesh& m, const Sysparams& s, const double* phi, const double* cp, const double* cm,
                     double* ip, double* im) {
  for (int i = 0; i < s.n_surfaces; i++) { ip[i] = 0; im[i] = 0; } // :40-43
  for (int e = 0; e < m.nT; e++) {
    const ElemGeo g = elem_geo(m, e);
    const int* tv = &m.tri[3 * e];
    for (int gi = 0; gi < 3; gi++) {
      const int f = FACE_ITER[gi];
      const int seg = m.fseg[3 * e + f];
      if (seg < 0) continue;
      const int la = FACE_V[f][0], lb = FACE_V[f][1], lc = 3 - la - lb;
      const double ax = m.x[tv[la]], ay = m.y[tv[la]], bx = m.x[tv[lb]], by = m.y[tv[lb]];
      const double ex = 0.5 * (ax + bx), ey = 0.5 * (ay + by);            // ii->geometry().center()
      // local = it->geometry().local(evalPos): barycentric weights (1/2, 1/2, 0) on the face's vertices
      double w[3] = {0, 0, 0}; w[la] = 0.5; w[lb] = 0.5;
      double vcp = 0, vcm = 0, gphi[2] = {0, 0}, gcp[2] = {0, 0}, gcm[2] = {0, 0};
      for (int k = 0; k < 3; k++) {
        vcp += cp[tv[k]] * w[k]; vcm += cm[tv[k]] * w[k];
        for (int d = 0; d < 2; d++) {
          gphi[d] += phi[tv[k]] * g.gphi[k][d]; gcp[d] += cp[tv[k]] * g.gphi[k][d]; gcm[d] += cm[tv[k]] * g.gphi[k][d];
        }
      }
      const double len = std::sqrt((bx - ax) * (bx - ax) + (by - ay) * (by - ay));
      double factor = len;                                                // ii->geometry().volume()
      if (s.cylindrical) factor *= 2 * s.PI * ey;                         // :63-64
      for (int d = 0; d < 2; d++) { gcp[d] *= -factor; gcm[d] *= -factor; gphi[d] *= factor; gphi[d] *= vcp; } // :65-69
      double nx = (by - ay) / len, ny = -(bx - ax) / len;                 // unitOuterNormal
      if (nx * (m.x[tv[lc]] - ex) + ny * (m.y[tv[lc]] - ey) > 0) { nx = -nx; ny = -ny; }
      const int pg = m.bphys[seg];                                        // :72
      ip[pg] += (gcp[0] + gphi[0]) * nx + (gcp[1] + gphi[1]) * ny;        // :73
      const double ratio = vcm / vcp;
      for (int d = 0; d < 2; d++) gphi[d] *= ratio;                       // :78
      im[pg] += (gcm[0] - gphi[0]) * nx + (gcm[1] - gphi[1]) * ny;        // :79
    }
  }
}

// DataWriter::writeData (datawriter.hh:45-94): element loop; centre, value at the centre, gradient; the stream keeps
// precision 5 / std::scientific; FieldVector prints its components separated by blanks; the groups are tab separated.
inline void write_cell_data(const Mesh& m, const double* u, const std::string& filename) {
  std::ofstream out(filename.c_str(), std::ios::out);
  out.precision(5);
  for (int e = 0; e < m.nT; e++) {
    const ElemGeo g = elem_geo(m, e);
    const int* tv = &m.tri[3 * e];
    const double cx = (g.x0 + g.x1 + g.x2) / 3.0, cy = (g.y0 + g.y1 + g.y2) / 3.0;
    double val = 0, gr[2] = {0, 0};
    for (int k = 0; k < 3; k++) {
      val += u[tv[k]] / 3.0;
      for (int d = 0; d < 2; d++) gr[d] += u[tv[k]] * g.gphi[k][d];
    }
    out << std::left << std::scientific << cx << " " << cy << "\t";
    out << std::left << val << "\t";
    out << std::left << gr[0] << " " << gr[1] << std::endl;
  }
}


// Dune::VTKWriter<GV>(gv, conforming).addVertexData(...).write(name, ascii | binaryappended)
// (instationary_pnp_from_pb_md.hh:337-340,440; stationary_pnp_from_pb.hh:190-192): one .vtu piece -- Float32 vertex data
// and points, Int32 connectivity / offsets, UInt8 cell types (5 = triangle); appended arrays are <uint32 bytes><raw>.
inline void write_vtk(const Mesh& m, const std::string& name, const std::vector<const double*>& fields,
                      const std::vector<std::string>& names, bool ascii) {
  std::ofstream f(name + ".vtu", std::ios::binary);
  if (!f) throw std::runtime_error("cannot open " + name + ".vtu");
  struct Blob { std::string bytes; };
  std::vector<Blob> blobs;
  unsigned long offset = 0;
  auto header = [&](const char* type, const std::string& nm, int ncomp) {
    f << "<DataArray type=\"" << type << "\" Name=\"" << nm << "\" NumberOfComponents=\"" << ncomp << "\" ";
  };
  auto put = [&](const char* type, const std::string& nm, int ncomp, const std::vector<double>& vals, int kind) {
    // kind 0: Float32, 1: Int32, 2: UInt8
    header(type, nm, ncomp);
    if (ascii) {
      f << "format=\"ascii\">\n";
      for (size_t i = 0; i < vals.size(); i++) {
        char buf[64];
        if (kind == 0) std::snprintf(buf, sizeof buf, "%g", (double)(float)vals[i]);
        else std::snprintf(buf, sizeof buf, "%d", (int)vals[i]);
        f << buf << (((i + 1) % 12 == 0 || i + 1 == vals.size()) ? '\n' : ' ');
      }
      f << "</DataArray>\n";
    } else {
      f << "format=\"appended\" offset=\"" << offset << "\" />\n";
      Blob b;
      for (double v : vals) {
        if (kind == 0) { float x = (float)v; b.bytes.append((const char*)&x, 4); }
        else if (kind == 1) { int x = (int)v; b.bytes.append((const char*)&x, 4); }
        else { unsigned char x = (unsigned char)v; b.bytes.append((const char*)&x, 1); }
      }
      offset += 4 + b.bytes.size();
      blobs.push_back(std::move(b));
    }
  };
  f << "<?xml version=\"1.0\"?>\n<VTKFile type=\"UnstructuredGrid\" version=\"0.1\" byte_order=\"LittleEndian\">\n";
  f << "<UnstructuredGrid>\n<Piece NumberOfCells=\"" << m.nT << "\" NumberOfPoints=\"" << m.nv << "\">\n";
  if (!fields.empty()) f << "<PointData Scalars=\"" << names[0] << "\">\n";
  for (size_t i = 0; i < fields.size(); i++) put("Float32", names[i], 1, std::vector<double>(fields[i], fields[i] + m.nv), 0);
  if (!fields.empty()) f << "</PointData>\n";
  std::vector<double> pts(3 * (size_t)m.nv), conn(3 * (size_t)m.nT), offs(m.nT), types(m.nT, 5.0);
  for (int v = 0; v < m.nv; v++) { pts[3 * v] = m.x[v]; pts[3 * v + 1] = m.y[v]; pts[3 * v + 2] = 0.0; }
  for (int e = 0; e < m.nT; e++) { for (int i = 0; i < 3; i++) conn[3 * e + i] = m.tri[3 * e + i]; offs[e] = 3 * (e + 1); }
  f << "<Points>\n"; put("Float32", "Coordinates", 3, pts, 0);
  f << "</Points>\n<Cells>\n";
  put("Int32", "connectivity", 1, conn, 1); put("Int32", "offsets", 1, offs, 1); put("UInt8", "types", 1, types, 2);
  f << "</Cells>\n</Piece>\n</UnstructuredGrid>\n";
  if (!ascii) {
    f << "<AppendedData encoding=\"raw\">\n_";
    for (auto& b : blobs) { unsigned n = (unsigned)b.bytes.size(); f.write((const char*)&n, 4); f.write(b.bytes.data(), b.bytes.size()); }
    f << "\n</AppendedData>\n";
  }
  f << "</VTKFile>\n";
}

} // namespace pnpo
