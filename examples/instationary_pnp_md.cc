// instationary_pnp_md -- what the reference binary does at HEAD (dune_pnp.cc:18-38 -> PnpSolverMain::run,
// /root/reference/src/pnp_solver_main.cc:70-116 -> instationary_pnp_md, instationary_pnp_from_pb_md.hh:112-455) on the B200
// backend: config -> Gmsh mesh -> PB Newton -> operator-split time loop.  Build: see INTEGRATION.md.
//   usage: instationary_pnp_md <config.cfg> [refinements] [nSteps] [files]     (the mesh file is named by the config;
//   "files": also write phiNNN.dat / cpNNN.dat / cmNNN.dat and current.dat into the working directory)
#include <cstdio>
#include <string>
#include <vector>

#include "pnp_b200/drivers.hh"

using namespace Dune::PNPB200;

int main(int argc, char** argv) {
  if (argc < 2) { std::printf("usage: %s <config.cfg> [refinements] [nSteps] [files]\n", argv[0]); return 1; }
  try {
    PnpSolverMain solver(0);
    const bool write_files = argc > 4 && std::string(argv[4]) == "files";
    int output_counter = 0;
    std::FILE* current = write_files ? std::fopen("current.dat", "w") : nullptr;
    solver.run(argv[1], argc > 2 ? std::stoi(argv[2]) : 0, argc > 3 ? std::stoi(argv[3]) : -1,
               [&](int step, double time, Vector& uphi, Vector& ucp, Vector& ucm) {
                 std::printf("step %4d  t = %-8g  |phi| %.10e  |c+| %.10e  |c-| %.10e\n", step, time, uphi.two_norm(),
                             ucp.two_norm(), ucm.two_norm());
                 if (!write_files) return;
                 // what the reference writes here (:430-452): phiNNN.dat / cpNNN.dat / cmNNN.dat and a line of current.dat
                 // (the VTK file is left to the host-side tooling)
                 char name[32];
                 output_counter++;
                 std::snprintf(name, sizeof name, "phi%03d.dat", output_counter); writeData(uphi, name);
                 std::snprintf(name, sizeof name, "cp%03d.dat", output_counter); writeData(ucp, name);
                 std::snprintf(name, sizeof name, "cm%03d.dat", output_counter); writeData(ucm, name);
                 std::vector<double> ip, im;
                 calcIonFlux(uphi, ucp, ucm, ip, im);
                 std::fprintf(current, "%g", time);
                 for (size_t s = 0; s < ip.size(); s++) std::fprintf(current, " %g 0 %g 0", ip[s], im[s]); // FieldVector<2>: second entry stays 0
                 std::fprintf(current, "\n"); std::fflush(current);
               });
    if (current) std::fclose(current);
  } catch (const Exception& e) {  // dune_pnp.cc:33-38
    std::printf("Dune reported error: %s\n", e.what());
    return 1;
  }
  return 0;
}
