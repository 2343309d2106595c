// pnp_output.cu -- the two diagnostics the reference's time loop produces from the fields:
//   calcIonFlux          /root/reference/src/ionFlux.hh:8-96        per-surface ion currents -> current.dat
//   DataWriter::writeData /root/reference/src/datawriter.hh:45-94   cell-centre "x y  value  gradx grady" text files
//   Dune::VTKWriter<GV>(gv, conforming) + addVertexData + write(name, binaryappended)
//                        /root/reference/src/instationary_pnp_from_pb_md.hh:337-340,440; stationary_pnp_from_pb.hh:190-192
// The ion current is a reduction over the O(sqrt N) boundary faces: one thread per face on the device, the per-surface
// sums on the host in face order (deterministic).  writeData is host I/O by nature: the field is downloaded once and the
// element loop runs on the host.
#include <cmath>
#include <cstdio>
#include <cstring>

#include "pnp_common.cuh"

namespace pnp {

namespace {

// contribution of one boundary face to (ip, im) of its surface; evaluation at the face centre (ionFlux.hh:51-83)
__global__ void k_ion_flux(const BFace* __restrict__ faces, int nB, const XY* __restrict__ xy, const double* __restrict__ phi,
                           const double* __restrict__ cp, const double* __restrict__ cm, int cylindrical, double PI, int n_own,
                           double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nB) return;
  const BFace b = faces[i];
  double ip = 0.0, im = 0.0;
  if (b.a < n_own) { // every face is counted by the rank that owns its first end vertex
    const XY P0 = xy[b.v[0]], P1 = xy[b.v[1]], P2 = xy[b.v[2]];
    const Geo G = make_geo(P0.x, P0.y, P1.x, P1.y, P2.x, P2.y);
    const int ia = face_v(b.f, 0), ib = face_v(b.f, 1), ic = 3 - ia - ib;
    const XY A = xy[b.v[ia]], B = xy[b.v[ib]], C = xy[b.v[ic]];
    const double cx = 0.5 * (A.x + B.x), cy = 0.5 * (A.y + B.y); // ii->geometry().center()
    // P1 values at the face centre (the local coordinates of the centre are (1/2, 1/2, 0) on the face's vertices)
    const double vcp = 0.5 * (cp[b.v[ia]] + cp[b.v[ib]]), vcm = 0.5 * (cm[b.v[ia]] + cm[b.v[ib]]);
    double gphi[2] = {0, 0}, gcp[2] = {0, 0}, gcm[2] = {0, 0};
    for (int k = 0; k < 3; k++)
      for (int d = 0; d < 2; d++) {
        gphi[d] += phi[b.v[k]] * G.g[k][d]; gcp[d] += cp[b.v[k]] * G.g[k][d]; gcm[d] += cm[b.v[k]] * G.g[k][d];
      }
    const double ex = B.x - A.x, ey = B.y - A.y, len = sqrt(ex * ex + ey * ey);
    double nx = ey / len, ny = -ex / len; // unit normal; outer = pointing away from the third vertex
    if (nx * (C.x - cx) + ny * (C.y - cy) > 0.0) { nx = -nx; ny = -ny; }
    double factor = len;                   // ii->geometry().volume()
    if (cylindrical) factor *= 2 * PI * cy;
    for (int d = 0; d < 2; d++) { gcp[d] *= -factor; gcm[d] *= -factor; gphi[d] *= factor; gphi[d] *= vcp; }
    ip = (gcp[0] + gphi[0]) * nx + (gcp[1] + gphi[1]) * ny;
    const double ratio = vcm / vcp;        // gradphi *= cm/cp (:78)
    for (int d = 0; d < 2; d++) gphi[d] *= ratio;
    im = (gcm[0] - gphi[0]) * nx + (gcm[1] - gphi[1]) * ny;
  }
  out[2 * i] = ip; out[2 * i + 1] = im;
}

} // namespace

// ip[s], im[s] for every surface s (ionFlux.hh accumulates into component 0 of ip[pg], im[pg])
void ion_flux(Ctx& c, const Vec& phi, const Vec& cp, const Vec& cm, double* ip, double* im) {
  PNP_REQUIRE(c.constraints_built, PNP_E_ARG, "mesh not finalized / parameters not set");
  if (c.degree >= 2) { p2_ion_flux(c, phi, cp, cm, ip, im); return; }
  PNP_REQUIRE(phi.fields == 1 && cp.fields == 1 && cm.fields == 1, PNP_E_ARG, "ion flux: three 1-field vectors expected");
  const int ns = c.params.n_surfaces, nB = (int)c.bfaces.size();
  for (int s = 0; s < ns; s++) ip[s] = im[s] = 0.0;
  std::vector<double> h(2 * (size_t)nB);
  if (nB) {
    DBuf<double> d(2 * (size_t)nB);
    k_ion_flux<<<(nB + 127) / 128, 128, 0, c.stream>>>(c.d_bfaces.p, nB, c.xy.p, phi.d.p, cp.d.p, cm.d.p, c.params.cylindrical,
                                                     c.params.PI, (int)c.n_own, d.p);
    PNP_CHECK_LAUNCH(); c.launches++;
    d.download(h.data(), h.size(), c.stream);
  }
  std::vector<double> sums(2 * (size_t)ns, 0.0);
  for (int i = 0; i < nB; i++) {
    const int pg = c.bfaces[i].phys;
    if (pg < 0 || pg >= ns) continue;
    sums[2 * pg] += h[2 * i]; sums[2 * pg + 1] += h[2 * i + 1];
  }
  if (c.world > 1) { // (the reference prints rank 0's partial sums only; here the surface currents are summed over the ranks)
    DBuf<double> d(sums.size());
    d.upload(sums.data(), sums.size(), c.stream);
    allreduce_sum(c, d.p, sums.size());
    d.download(sums.data(), sums.size(), c.stream);
  }
  for (int s = 0; s < ns; s++) { ip[s] = sums[2 * s]; im[s] = sums[2 * s + 1]; }
}

// DataWriter::writeData: one line per element in grid order: centre, value at the centre, gradient; std::scientific with
// precision 5, FieldVector components separated by blanks, the three groups by tabs.  (The "This is intro" line the
// reference writes first is lost when it reopens the file with std::ios::out, datawriter.hh:52-60; same here.)
void write_cell_data(Ctx& c, const Vec& u, const std::string& filename) {
  PNP_REQUIRE(c.finalized && u.fields == 1, PNP_E_ARG, "writeData: finalized mesh and a 1-field vector expected");
  PNP_REQUIRE(c.n_own == c.nv, PNP_E_ARG, "writeData walks the global element order: one subdomain only");
  if (c.degree >= 2) { p2_write_cell_data(c, u, filename); return; }
  std::vector<double> lex((size_t)c.nv);
  vec_download(c, u, lex.data());
  const std::vector<double> x = c.cx.to_host(c.stream), y = c.cy.to_host(c.stream);
  const std::vector<int> tri = c.ctri.to_host(c.stream);
  std::FILE* f = std::fopen(filename.c_str(), "w");
  PNP_REQUIRE(f, PNP_E_CONFIG, "cannot open " + filename);
  for (long e = 0; e < c.nT; e++) {
    const int a = tri[3 * e], b = tri[3 * e + 1], d = tri[3 * e + 2];
    const Geo G = make_geo(x[a], y[a], x[b], y[b], x[d], y[d]);
    const double cx = (x[a] + x[b] + x[d]) / 3.0, cy = (y[a] + y[b] + y[d]) / 3.0;
    const double val = (lex[a] + lex[b] + lex[d]) / 3.0;
    const double gx = lex[a] * G.g[0][0] + lex[b] * G.g[1][0] + lex[d] * G.g[2][0];
    const double gy = lex[a] * G.g[0][1] + lex[b] * G.g[1][1] + lex[d] * G.g[2][1];
    std::fprintf(f, "%.5e %.5e\t%.5e\t%.5e %.5e\n", cx, cy, val, gx, gy);
  }
  std::fclose(f);
}

// Dune::VTKWriter (dune-grid 2.2, conforming): one .vtu piece with vertex data (Float32, as DUNE writes it), points,
// cells (triangles: connectivity, offsets, types = 5).  binaryappended: every array is "<uint32 byte count><raw bytes>"
// behind the '_' of <AppendedData encoding="raw">, the DataArray offsets count from there; ascii: the same arrays in
// place.  Several ranks: rank r writes the piece "s<world>:p<rank>:<name>.vtu" (its owned + ghost vertices and all its
// elements) and rank 0 the "s<world>:<name>.pvtu" index, as DUNE's parallel writer names them.
void write_vtk(Ctx& c, const std::string& name, int nfields, const Vec* const* fields, const char* const* names, int ascii) {
  PNP_REQUIRE(c.finalized && nfields >= 0, PNP_E_ARG, "VTK writer: finalized mesh expected");
  for (int i = 0; i < nfields; i++)
    PNP_REQUIRE(fields[i] && fields[i]->fields == 1 && names[i], PNP_E_ARG, "VTK writer: 1-field vectors with names expected");
  const long nv = c.nv, nT = c.nT;
  std::vector<std::vector<float>> data(nfields, std::vector<float>((size_t)nv));
  std::vector<double> lex((size_t)nv);
  for (int i = 0; i < nfields; i++) {
    if (c.degree >= 2) p2_vertex_values(c, *fields[i], lex.data()); // vertex data of a quadratic function = its vertex dofs
    else vec_download(c, *fields[i], lex.data());
    for (long v = 0; v < nv; v++) data[i][v] = (float)lex[v];
  }
  const std::vector<double> x = c.cx.to_host(c.stream), y = c.cy.to_host(c.stream);
  const std::vector<int> tri = c.ctri.to_host(c.stream);
  std::vector<float> pts(3 * (size_t)nv);
  for (long v = 0; v < nv; v++) { pts[3 * v] = (float)x[v]; pts[3 * v + 1] = (float)y[v]; pts[3 * v + 2] = 0.0f; }
  std::vector<int> offs((size_t)nT);
  for (long e = 0; e < nT; e++) offs[e] = (int)(3 * (e + 1));
  std::vector<unsigned char> types((size_t)nT, 5); // VTK_TRIANGLE
  char piece[512];
  if (c.world > 1) std::snprintf(piece, sizeof piece, "s%04d:p%04d:%s.vtu", c.world, c.rank, name.c_str());
  else std::snprintf(piece, sizeof piece, "%s.vtu", name.c_str());
  // a directory part of `name` stays in front of the DUNE prefix
  std::string dir, base = name;
  const size_t slash = name.find_last_of('/');
  if (slash != std::string::npos) { dir = name.substr(0, slash + 1); base = name.substr(slash + 1); }
  if (c.world > 1) std::snprintf(piece, sizeof piece, "%ss%04d:p%04d:%s.vtu", dir.c_str(), c.world, c.rank, base.c_str());
  std::FILE* f = std::fopen(piece, "wb");
  PNP_REQUIRE(f, PNP_E_CONFIG, std::string("cannot open ") + piece);
  unsigned long offset = 0;
  std::vector<std::pair<const void*, unsigned>> blobs; // appended arrays in file order
  auto array = [&](const char* type, const char* nm, int ncomp, const void* ptr, size_t count, size_t elsize, auto print) {
    std::fprintf(f, "<DataArray type=\"%s\" Name=\"%s\" NumberOfComponents=\"%d\" ", type, nm, ncomp);
    if (ascii) {
      std::fprintf(f, "format=\"ascii\">\n");
      for (size_t i = 0; i < count; i++) { print(i); std::fputc((i + 1) % 12 == 0 || i + 1 == count ? '\n' : ' ', f); }
      std::fprintf(f, "</DataArray>\n");
    } else {
      std::fprintf(f, "format=\"appended\" offset=\"%lu\" />\n", offset);
      blobs.push_back({ptr, (unsigned)(count * elsize)});
      offset += 4 + count * elsize;
    }
  };
  std::fprintf(f, "<?xml version=\"1.0\"?>\n<VTKFile type=\"UnstructuredGrid\" version=\"0.1\" byte_order=\"LittleEndian\">\n");
  std::fprintf(f, "<UnstructuredGrid>\n<Piece NumberOfCells=\"%ld\" NumberOfPoints=\"%ld\">\n", nT, nv);
  if (nfields > 0) std::fprintf(f, "<PointData Scalars=\"%s\">\n", names[0]);
  for (int i = 0; i < nfields; i++)
    array("Float32", names[i], 1, data[i].data(), (size_t)nv, 4, [&](size_t k) { std::fprintf(f, "%g", (double)data[i][k]); });
  if (nfields > 0) std::fprintf(f, "</PointData>\n");
  std::fprintf(f, "<Points>\n");
  array("Float32", "Coordinates", 3, pts.data(), 3 * (size_t)nv, 4, [&](size_t k) { std::fprintf(f, "%g", (double)pts[k]); });
  std::fprintf(f, "</Points>\n<Cells>\n");
  array("Int32", "connectivity", 1, tri.data(), 3 * (size_t)nT, 4, [&](size_t k) { std::fprintf(f, "%d", tri[k]); });
  array("Int32", "offsets", 1, offs.data(), (size_t)nT, 4, [&](size_t k) { std::fprintf(f, "%d", offs[k]); });
  array("UInt8", "types", 1, types.data(), (size_t)nT, 1, [&](size_t k) { std::fprintf(f, "%d", (int)types[k]); });
  std::fprintf(f, "</Cells>\n</Piece>\n</UnstructuredGrid>\n");
  if (!ascii) {
    std::fprintf(f, "<AppendedData encoding=\"raw\">\n_");
    for (auto& b : blobs) { std::fwrite(&b.second, 4, 1, f); std::fwrite(b.first, 1, b.second, f); }
    std::fprintf(f, "\n</AppendedData>\n");
  }
  std::fprintf(f, "</VTKFile>\n");
  std::fclose(f);
  if (c.world > 1 && c.rank == 0) {
    char idx[512];
    std::snprintf(idx, sizeof idx, "%ss%04d:%s.pvtu", dir.c_str(), c.world, base.c_str());
    std::FILE* g = std::fopen(idx, "w");
    PNP_REQUIRE(g, PNP_E_CONFIG, std::string("cannot open ") + idx);
    std::fprintf(g, "<?xml version=\"1.0\"?>\n<VTKFile type=\"PUnstructuredGrid\" version=\"0.1\" byte_order=\"LittleEndian\">\n<PUnstructuredGrid GhostLevel=\"0\">\n");
    if (nfields > 0) std::fprintf(g, "<PPointData Scalars=\"%s\">\n", names[0]);
    for (int i = 0; i < nfields; i++) std::fprintf(g, "<PDataArray type=\"Float32\" Name=\"%s\" NumberOfComponents=\"1\"/>\n", names[i]);
    if (nfields > 0) std::fprintf(g, "</PPointData>\n");
    std::fprintf(g, "<PPoints>\n<PDataArray type=\"Float32\" Name=\"Coordinates\" NumberOfComponents=\"3\"/>\n</PPoints>\n");
    for (int r = 0; r < c.world; r++) std::fprintf(g, "<Piece Source=\"s%04d:p%04d:%s.vtu\"/>\n", c.world, r, base.c_str());
    std::fprintf(g, "</PUnstructuredGrid>\n</VTKFile>\n");
    std::fclose(g);
  }
}

} // namespace pnp
