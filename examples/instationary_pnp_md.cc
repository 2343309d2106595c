// instationary_pnp_md -- what the reference binary does at HEAD (dune_pnp.cc:18-38 -> PnpSolverMain::run,
// /root/reference/src/pnp_solver_main.cc:70-116 -> instationary_pnp_md, instationary_pnp_from_pb_md.hh:112-455) on the B200
// backend: config -> Gmsh mesh -> PB Newton -> operator-split time loop.  Build: see INTEGRATION.md.
//   usage: instationary_pnp_md <config.cfg> [refinements] [nSteps]     (the mesh file is named by the config)
#include <cstdio>
#include <string>

#include "pnp_b200/drivers.hh"

using namespace Dune::PNPB200;

int main(int argc, char** argv) {
  if (argc < 2) { std::printf("usage: %s <config.cfg> [refinements] [nSteps]\n", argv[0]); return 1; }
  try {
    PnpSolverMain solver(0);
    solver.run(argv[1], argc > 2 ? std::stoi(argv[2]) : 0, argc > 3 ? std::stoi(argv[3]) : -1,
               [](int step, double time, Vector& uphi, Vector& ucp, Vector& ucm) {
                 // here the reference writes phiNNN.dat / cpNNN.dat / cmNNN.dat, dataNNN.vtu and current.dat (:430-452)
                 std::printf("step %4d  t = %-8g  |phi| %.10e  |c+| %.10e  |c-| %.10e\n", step, time, uphi.two_norm(),
                             ucp.two_norm(), ucm.two_norm());
               });
  } catch (const Exception& e) {  // dune_pnp.cc:33-38
    std::printf("Dune reported error: %s\n", e.what());
    return 1;
  }
  return 0;
}
