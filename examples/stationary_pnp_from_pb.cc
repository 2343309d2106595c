// stationary_pnp_from_pb -- the reference driver of the same name (/root/reference/src/stationary_pnp_from_pb.hh:92-370) on the
// B200 backend: PB Newton solve, BCExtension interpolation, monolithic PNP Newton solve (drivers.hh).  The backend is picked
// by typedef as in the reference; the multigrid-preconditioned BiCGSTAB is the one that scales.  Build: see INTEGRATION.md.
// -DPDEGREE=2 builds the quadratic-element program like the reference's Makefile.am (:57-110) does.
#include <cstdio>
#include <string>

#include "pnp_b200/drivers.hh"

using namespace Dune::PNPB200;

int main(int argc, char** argv) {
  if (argc < 3) { std::printf("usage: %s <config.cfg> <mesh.msh> [refinements]\n", argv[0]); return 1; }
  try {
    Grid grid(0);
    grid.readConfigFile(argv[1]);                       // Sysparams::readConfigFile
    grid.readGmsh(argv[2]);                             // GmshReader<UGGrid<2>>::read + createGrid
    if (argc > 3) grid.globalRefine(std::stoi(argv[3]));
    grid.setDegree(PDEGREE);                            // Pk2DLocalFiniteElementMap<GV, D, R, PDEGREE>
    grid.finalize();
#if PDEGREE == 1
    typedef ISTLBackend_NOVLP_BCGS_AMG LS;              // reference: ISTLBackend_NOVLP_BCGS_SSORk<GO> (LINEARSOLVER == 1)
    LS pbls(grid, 2, 5000, 0), ls(grid, 2, 20000, 0);
#else
    typedef ISTLBackend_NOVLP_BCGS_ILU0 LS;             // (the multigrid is built for linear elements)
    LS pbls(grid, 5000, 0), ls(grid, 50000, 0);
#endif
    Vector u(grid, 3);
    const pnp_newton_result r = stationary_pnp_from_pb(grid, u, pbls, ls, PNP_JAC_ANALYTIC);
    std::printf("PNP Newton: %d iterations, defect %.3e -> %.3e, %d linear iterations, %.3f s\n", r.iterations,
                r.first_defect, r.defect, r.linear_iterations, r.seconds_total);
  } catch (const Exception& e) {
    std::printf("Dune reported error: %s\n", e.what());
    return 1;
  }
  return 0;
}
