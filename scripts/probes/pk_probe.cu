// probe: PkElem<3>::alpha_boundary / basis on the device against the same functions on the host
#include <cstdio>
#include "../../dune_pnp_b200/csrc/pnp_elem_p2.cuh"
using namespace pnp;
__global__ void k(int npts, int f, double* out, double* ph) {
  double rl[10] = {};
  double j[3] = {0.37, 0, 0}; bool skip[3] = {false, false, false};
  PhysParams P; P.PI = 3.1415; P.cylindrical = 0;
  PkElem<3>::alpha_boundary(f, 0.1, 0.2, 1.3, 0.5, 1, j, skip, P, npts, rl);
  for (int i = 0; i < 10; i++) out[i] = rl[i];
  PkElem<3>::basis(0.11270166537925831148, 0.0, ph);
}
int main() {
  double *d, *p; cudaMalloc(&d, 80); cudaMalloc(&p, 80);
  for (int npts = 2; npts <= 3; npts++) for (int f = 0; f < 3; f++) {
    k<<<1, 1>>>(npts, f, d, p);
    double h[10], hp[10], r[10] = {}, rp[10];
    cudaMemcpy(h, d, 80, cudaMemcpyDeviceToHost); cudaMemcpy(hp, p, 80, cudaMemcpyDeviceToHost);
    double j[3] = {0.37, 0, 0}; bool skip[3] = {false, false, false};
    PhysParams P; P.PI = 3.1415; P.cylindrical = 0;
    PkElem<3>::alpha_boundary(f, 0.1, 0.2, 1.3, 0.5, 1, j, skip, P, npts, r);
    PkElem<3>::basis(0.11270166537925831148, 0.0, rp);
    for (int i = 0; i < 10; i++) std::printf("npts %d f %d i %d dev % .17g host % .17g diff %g | phi dev % .17g host % .17g\n", npts, f, i, h[i], r[i], h[i] - r[i], hp[i], rp[i]);
  }
  return 0;
}
