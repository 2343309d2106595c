// pnp_hostside.cu -- host-side pieces of the path that are O(boundary) or file parsing:
//   * Gmsh ASCII 2.x ingest with GmshReader<UGGrid<2>> semantics (pnp_solver_main.cc:82-91)
//   * the INI subset read by Sysparams::readConfigFile (sysparams.cc:15-98)
//   * interpolate(BCExtension) (dirichlet_bc.hh:54-123; stationary_pnp_from_pb.hh:235-270):
//     the bulk (PB-derived initial guess) is one streaming kernel; the boundary band, where the
//     reference's line-membership test and its missing-break precedence rule decide, is evaluated
//     on the host from a gathered O(sqrt N) subset and scattered back.
#include <algorithm>
#include <cmath>
#include <fstream>
#include <map>
#include <sstream>

#include "pnp_common.cuh"

namespace pnp {

void read_gmsh_file(const std::string& path, std::vector<double>& x, std::vector<double>& y, std::vector<int>& tri,
                    std::vector<int>& ba, std::vector<int>& bb, std::vector<int>& bphys) {
  std::ifstream in(path);
  PNP_REQUIRE((bool)in, PNP_E_CONFIG, "cannot open mesh file " + path);
  // nodes per Gmsh element type 1..15
  static const int nodes_of[16] = {0, 2, 3, 4, 4, 8, 6, 5, 3, 6, 9, 10, 27, 18, 14, 1};
  std::vector<long> node_id; std::vector<double> nx, ny;
  std::vector<long> tri_nodes, line_nodes; std::vector<int> line_phys;
  std::string tok;
  while (in >> tok) {
    if (tok == "$MeshFormat") {
      double ver; int ftype, dsize; in >> ver >> ftype >> dsize;
      PNP_REQUIRE(in && ver >= 2.0 && ver < 3.0 && ftype == 0, PNP_E_MESH, "only Gmsh ASCII format 2.x is supported");
    } else if (tok == "$Nodes") {
      long n; in >> n;
      node_id.resize(n); nx.resize(n); ny.resize(n);
      for (long i = 0; i < n; i++) { double z; in >> node_id[i] >> nx[i] >> ny[i] >> z; }
      PNP_REQUIRE((bool)in, PNP_E_MESH, "truncated $Nodes section");
    } else if (tok == "$Elements") {
      long n; in >> n;
      for (long i = 0; i < n; i++) {
        long id; int type, ntags; in >> id >> type >> ntags;
        PNP_REQUIRE(in && type >= 1 && type <= 15, PNP_E_MESH, "unsupported Gmsh element type");
        int phys = 0;
        for (int t = 0; t < ntags; t++) { int tag; in >> tag; if (t == 0) phys = tag; }
        long nd[27];
        for (int k = 0; k < nodes_of[type]; k++) in >> nd[k];
        if (type == 2) tri_nodes.insert(tri_nodes.end(), nd, nd + 3);      // insertElement, file order
        else if (type == 1) { line_nodes.insert(line_nodes.end(), nd, nd + 2); line_phys.push_back(phys); } // boundary segment
      }
      PNP_REQUIRE((bool)in, PNP_E_MESH, "truncated $Elements section");
    }
  }
  PNP_REQUIRE(!tri_nodes.empty(), PNP_E_MESH, "mesh file contains no triangles");
  // vertices = nodes referenced by triangles, compressed in ascending node-id order
  std::map<long, int> index;
  for (long v : tri_nodes) index[v] = 0;
  std::map<long, long> pos;
  for (size_t i = 0; i < node_id.size(); i++) pos[node_id[i]] = (long)i;
  int k = 0;
  x.clear(); y.clear();
  for (auto& kv : index) {
    auto it = pos.find(kv.first);
    PNP_REQUIRE(it != pos.end(), PNP_E_MESH, "triangle references an undefined node");
    kv.second = k++;
    x.push_back(nx[it->second]); y.push_back(ny[it->second]);
  }
  tri.resize(tri_nodes.size());
  for (size_t i = 0; i < tri_nodes.size(); i++) tri[i] = index[tri_nodes[i]];
  ba.clear(); bb.clear(); bphys = line_phys;
  for (size_t s = 0; s < line_phys.size(); s++) {
    auto a = index.find(line_nodes[2 * s]), b = index.find(line_nodes[2 * s + 1]);
    PNP_REQUIRE(a != index.end() && b != index.end(), PNP_E_MESH, "line element uses a node no triangle uses");
    ba.push_back(a->second); bb.push_back(b->second);
  }
}

void read_config_file(const std::string& path, HostParams& p) {
  std::ifstream in(path);
  PNP_REQUIRE((bool)in, PNP_E_CONFIG, "Could not read config file \"" + path + "\"!");
  std::map<std::string, std::string> kv;
  std::string line, section;
  auto strip = [](const std::string& s) {
    size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
    return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
  };
  while (std::getline(in, line)) {
    size_t hash = line.find('#');
    if (hash != std::string::npos) line.erase(hash);
    line = strip(line);
    if (line.empty()) continue;
    if (line.front() == '[') { section = strip(line.substr(1, line.find(']') - 1)); continue; }
    size_t eq = line.find('=');
    if (eq != std::string::npos) kv[section + "." + strip(line.substr(0, eq))] = strip(line.substr(eq + 1));
  }
  auto need = [&](const std::string& key) -> const std::string& {
    auto it = kv.find(key);
    PNP_REQUIRE(it != kv.end(), PNP_E_CONFIG, "config key missing: " + key);
    return it->second;
  };
  // keys the stale sphere/cylinder cfgs lack get the one_wall.cfg values
  auto opt = [&](const std::string& key, double dflt) {
    auto it = kv.find(key);
    return it == kv.end() ? dflt : std::stod(it->second);
  };
  HostParams q;
  q.meshfile = need("mesh.filename");
  q.n_surfaces = std::stoi(need("system.n_surfaces"));
  PNP_REQUIRE(q.n_surfaces >= 0, PNP_E_CONFIG, "system.n_surfaces must be non-negative");
  q.verbosity = (int)opt("system.verbosity", 0);
  q.cylindrical = (int)opt("system.cylindrical", 0) != 0;
  q.l_b = opt("system.l_b", 1.0);
  q.c0 = opt("system.c0", 0.06);
  q.linearSolverIterations = (int)opt("system.linearSolverIterations", 50);
  q.newtonReassembleThreshold = opt("system.newtonReassembleThreshold", 0.0);
  q.newtonReduction = opt("system.newtonReduction", 1e-5);
  q.newtonMinLinearReduction = opt("system.newtonMinLinearReduction", 1e-5);
  q.newtonMaxIterations = opt("system.newtonMaxIterations", 50);
  q.newtonLineSearchMaxIteration = opt("system.newtonLineSearchMaxIteration", 500);
  q.tau = opt("system.tau", 0.1);
  q.nSteps = (int)opt("system.nSteps", 100);
  q.outputFreq = (int)opt("system.outputFreq", 1);
  q.potentialUpdateFreq = (int)opt("system.potentialUpdateFreq", 1);
  q.surfaces.assign(q.n_surfaces, HostSurface());
  static const char* bt[3] = {"coulombBtype", "plusDiffusionBtype", "minusDiffusionBtype"};
  static const char* dv[3] = {"coulombPotential", "plusDiffusionConcentration", "minusDiffusionConcentration"};
  static const char* fl[3] = {"coulombFlux", "plusDiffusionFlux", "minusDiffusionFlux"};
  for (int i = 0; i < q.n_surfaces; i++) {
    const std::string sec = "surface_" + std::to_string(i) + ".";
    for (int k = 0; k < 3; k++) { // only the value that matches the type is read (sysparams.cc:70-93)
      q.surfaces[i].btype[k] = std::stoi(need(sec + bt[k]));
      if (q.surfaces[i].btype[k] == 0) q.surfaces[i].dval[k] = std::stod(need(sec + dv[k]));
      if (q.surfaces[i].btype[k] == 1) q.surfaces[i].flux[k] = std::stod(need(sec + fl[k]));
    }
  }
  q.set = true;
  p = q;
}

// ---------------------------------------------------------------------------------------------
namespace {

__global__ void k_pb_guess(const double* __restrict__ pb, long nv, int comp, double c0, double* __restrict__ out) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < nv; i += (long)gridDim.x * blockDim.x) {
    const double y = pb ? pb[i] : 0.0;
    out[i] = comp == 0 ? y : (comp == 1 ? c0 * exp(-y) : c0 * exp(+y)); // dirichlet_bc.hh:99,107,115
  }
}
__global__ void k_flag_vertices(const int* __restrict__ a, const int* __restrict__ b, long n, unsigned char* __restrict__ flag) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    flag[a[i]] = 1; flag[b[i]] = 1;
  }
}
__global__ void k_last_element(const int* __restrict__ tri, long nT, int* __restrict__ last) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < 3 * nT; i += (long)gridDim.x * blockDim.x)
    atomicMax(&last[tri[i]], (int)(i / 3));
}
__global__ void k_band_elements(const int* __restrict__ tri, long nT, const unsigned char* __restrict__ flag,
                                int* __restrict__ count, int* __restrict__ list, int cap) {
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < nT; t += (long)gridDim.x * blockDim.x)
    if (flag[tri[3 * t]] | flag[tri[3 * t + 1]] | flag[tri[3 * t + 2]]) {
      int k = atomicAdd(count, 1);
      if (k < cap) list[k] = (int)t;
    }
}
struct BandVertex { double x, y, pb; int last; int pad; };
__global__ void k_gather_band(const int* __restrict__ vext, int n, const double* __restrict__ x, const double* __restrict__ y,
                              const double* __restrict__ pb_int, const int* __restrict__ ext2int,
                              const int* __restrict__ last, BandVertex* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int v = vext[i];
  out[i] = BandVertex{x[v], y[v], pb_int ? pb_int[ext2int[v]] : 0.0, last[v], 0};
}
__global__ void k_gather_tri(const int* __restrict__ list, int n, const int* __restrict__ tri, int* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int k = 0; k < 3; k++) out[3 * i + k] = tri[3 * (long)list[i] + k];
}
__global__ void k_scatter_values(const int* __restrict__ vext, const double* __restrict__ val, int n,
                                 const int* __restrict__ ext2int, double* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[ext2int[vext[i]]] = val[i];
}

// global_on_intersection (dirichlet_bc.hh:21-33): distance to the infinite line through the face
bool on_line(double px, double py, double ax, double ay, double bx, double by) {
  double vx = bx - ax, vy = by - ay;
  const double nrm = std::sqrt(vx * vx + vy * vy);
  vx /= nrm; vy /= nrm;
  const double dx = px - ax, dy = py - ay;
  const double t = dx * vx + dy * vy;
  const double ex = vx * t - dx, ey = vy * t - dy;
  return std::sqrt(ex * ex + ey * ey) < 1e-9;
}

} // namespace

void interpolate_bcext(Ctx& c, int comp, const Vec* pb, Vec& out) {
  PNP_REQUIRE(c.constraints_built, PNP_E_ARG, "constraints not built");
  PNP_REQUIRE(comp >= 0 && comp < 3, PNP_E_ARG, "component must be 0, 1 or 2");
  PNP_REQUIRE(out.fields == 1 && (!pb || pb->fields == 1), PNP_E_ARG, "interpolate works on 1-field vectors");
  const long nv = c.nv, nT = c.nT, nB = c.nB;
  const double* pbp = pb ? pb->d.p : nullptr;
  k_pb_guess<<<grid_for(nv, 256), 256, 0, c.stream>>>(pbp, nv, comp, c.params.c0, out.d.p);
  PNP_CHECK_LAUNCH(); c.launches++;
  if (nB == 0) return;

  // --- boundary band: elements with at least one boundary vertex ---
  DBuf<unsigned char> flag(nv); flag.zero(c.stream);
  k_flag_vertices<<<grid_for(nB, 256), 256, 0, c.stream>>>(c.cba.p, c.cbb.p, nB, flag.p);
  DBuf<int> last(nv);
  PNP_CUDA(cudaMemsetAsync(last.p, 0xff, nv * sizeof(int), c.stream)); // -1
  k_last_element<<<grid_for(3 * nT, 256), 256, 0, c.stream>>>(c.ctri.p, nT, last.p);
  int cap = (int)std::min<long>(nT, 64 * nB + 1024);
  DBuf<int> count(1), list(cap);
  int nband = 0;
  for (int attempt = 0; attempt < 2; attempt++) {
    count.zero(c.stream);
    k_band_elements<<<grid_for(nT, 256), 256, 0, c.stream>>>(c.ctri.p, nT, flag.p, count.p, list.p, cap);
    PNP_CHECK_LAUNCH();
    count.download(&nband, 1, c.stream);
    if (nband <= cap) break;
    cap = nband; list.alloc(cap);
  }
  c.launches += 4;
  std::vector<int> band(nband);
  list.download(band.data(), nband, c.stream);
  std::sort(band.begin(), band.end());
  list.upload(band.data(), nband, c.stream);
  DBuf<int> dtri(3 * (size_t)nband);
  k_gather_tri<<<(nband + 255) / 256, 256, 0, c.stream>>>(list.p, nband, c.ctri.p, dtri.p);
  std::vector<int> btri = dtri.to_host(c.stream);
  std::vector<int> bvert(btri);
  std::sort(bvert.begin(), bvert.end());
  bvert.erase(std::unique(bvert.begin(), bvert.end()), bvert.end());
  const int nbv = (int)bvert.size();
  DBuf<int> dv(nbv); dv.upload(bvert.data(), nbv, c.stream);
  DBuf<BandVertex> dbv(nbv);
  k_gather_band<<<(nbv + 255) / 256, 256, 0, c.stream>>>(dv.p, nbv, c.cx.p, c.cy.p, pbp, c.ext2int.p, last.p, dbv.p);
  PNP_CHECK_LAUNCH(); c.launches += 2;
  std::vector<BandVertex> hv = dbv.to_host(c.stream);
  std::vector<int> hba = c.cba.to_host(c.stream), hbb = c.cbb.to_host(c.stream), hph = c.cbphys.to_host(c.stream);
  auto vidx = [&](int vext) { return (int)(std::lower_bound(bvert.begin(), bvert.end(), vext) - bvert.begin()); };

  // edge -> boundary segment ; edge -> band elements sharing it
  static const int FV[3][2] = {{0, 1}, {0, 2}, {1, 2}};
  static const int FACE_ITER[3] = {0, 2, 1}; // UG side order (v0v1),(v1v2),(v2v0) in DUNE face numbers
  auto ekey = [](int a, int b) { return std::make_pair(std::min(a, b), std::max(a, b)); };
  std::map<std::pair<int, int>, int> seg_of;
  for (long s = 0; s < nB; s++) seg_of[ekey(hba[s], hbb[s])] = (int)s;
  std::map<std::pair<int, int>, std::pair<int, int>> elems_of; // edge -> (band idx, band idx)
  for (int e = 0; e < nband; e++)
    for (int f = 0; f < 3; f++) {
      auto key = ekey(btri[3 * e + FV[f][0]], btri[3 * e + FV[f][1]]);
      auto it = elems_of.find(key);
      if (it == elems_of.end()) elems_of[key] = {e, -1}; else it->second.second = e;
    }
  auto sticky = [&](int pg) { return c.params.surfaces[pg].btype[2] == 0; }; // bctype() falls through to minusDiffusion (:40-51)
  auto line_test = [&](double px, double py, int e, int f) {
    const BandVertex& A = hv[vidx(btri[3 * e + FV[f][0]])];
    const BandVertex& B = hv[vidx(btri[3 * e + FV[f][1]])];
    return on_line(px, py, A.x, A.y, B.x, B.y);
  };
  std::vector<int> fix_v; std::vector<double> fix_val;
  for (int e = 0; e < nband; e++)
    for (int i = 0; i < 3; i++) {
      const int v = btri[3 * e + i];
      const BandVertex& V = hv[vidx(v)];
      if (V.last != band[e]) continue; // a later element overwrites this one (interpolate order)
      int pg = -1;
      for (int fi = 0; fi < 3; fi++) {
        const int f = FACE_ITER[fi];
        auto key = ekey(btri[3 * e + FV[f][0]], btri[3 * e + FV[f][1]]);
        auto sg = seg_of.find(key);
        if (sg != seg_of.end()) { // ii->boundary()
          if (line_test(V.x, V.y, e, f) && (pg == -1 || !sticky(pg))) pg = hph[sg->second];
        } else {                  // ii->neighbor(): scan the neighbour's boundary intersections
          auto el = elems_of[key];
          const int o = el.first == e ? el.second : el.first;
          if (o < 0) continue;    // neighbour has no boundary vertex, hence no boundary face
          for (int gi = 0; gi < 3; gi++) {
            const int f2 = FACE_ITER[gi];
            auto s2 = seg_of.find(ekey(btri[3 * o + FV[f2][0]], btri[3 * o + FV[f2][1]]));
            if (s2 != seg_of.end() && line_test(V.x, V.y, o, f2) && (pg == -1 || !sticky(pg))) pg = hph[s2->second];
          }
        }
      }
      if (pg > -1 && c.params.surfaces[pg].btype[comp] == 0) {
        fix_v.push_back(v); fix_val.push_back(c.params.surfaces[pg].dval[comp]);
      }
    }
  if (!fix_v.empty()) {
    const int n = (int)fix_v.size();
    DBuf<int> fv(n); DBuf<double> fx(n);
    fv.upload(fix_v.data(), n, c.stream); fx.upload(fix_val.data(), n, c.stream);
    k_scatter_values<<<(n + 255) / 256, 256, 0, c.stream>>>(fv.p, fx.p, n, c.ext2int.p, out.d.p);
    PNP_CHECK_LAUNCH(); c.launches++;
  }
  PNP_CUDA(cudaStreamSynchronize(c.stream));
}

} // namespace pnp
