// stationary_pnp -- /root/reference/src/stationary_pnp.hh:91-365 on the B200 backend: interpolate(BCExtension) without a PB
// stage, then the monolithic PNP Newton with ISTLBackend_NOVLP_BCGS_NOPREC as the reference selects (:254-256).
#include <cstdio>
#include <string>

#include "pnp_b200/drivers.hh"

using namespace Dune::PNPB200;

int main(int argc, char** argv) {
  if (argc < 3) { std::printf("usage: %s <config.cfg> <mesh.msh> [refinements]\n", argv[0]); return 1; }
  try {
    Grid grid(0);
    grid.readConfigFile(argv[1]);
    grid.readGmsh(argv[2]);
    if (argc > 3) grid.globalRefine(std::stoi(argv[3]));
    grid.finalize();
    ISTLBackend_NOVLP_BCGS_NOPREC ls(grid, (unsigned)sysparam(grid, 5), (int)sysparam(grid, 15));
    Vector u(grid, 3);
    const pnp_newton_result r = stationary_pnp(grid, u, ls);
    std::printf("PNP Newton: converged %d, %d iterations, defect %.3e -> %.3e, %d linear iterations\n", r.converged, r.iterations,
                r.first_defect, r.defect, r.linear_iterations);
  } catch (const Exception& e) {
    std::printf("Dune reported error: %s\n", e.what());
    return 1;
  }
  return 0;
}
