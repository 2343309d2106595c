// pnp_output.cu -- the two diagnostics the reference's time loop produces from the fields:
//   calcIonFlux          /root/reference/src/ionFlux.hh:8-96        per-surface ion currents -> current.dat
//   DataWriter::writeData /root/reference/src/datawriter.hh:45-94   cell-centre "x y  value  gradx grady" text files
// The ion current is a reduction over the O(sqrt N) boundary faces: one thread per face on the device, the per-surface
// sums on the host in face order (deterministic).  writeData is host I/O by nature: the field is downloaded once and the
// element loop runs on the host.
#include <cmath>
#include <cstdio>

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

} // namespace pnp
