// stationary_pnp_from_pb -- the reference driver of the same name (/root/reference/src/stationary_pnp_from_pb.hh:92-370,
// launched from PnpSolverMain::run, pnp_solver_main.cc:70-116) written against the facade: PB Newton solve, BCExtension
// interpolation, monolithic PNP Newton solve.  Build: see INTEGRATION.md.
#include <cstdio>
#include <string>

#include "pnp_b200/pdelab_facade.hh"

using namespace Dune::PNPB200;

int main(int argc, char** argv) {
  if (argc < 3) { std::printf("usage: %s <config.cfg> <mesh.msh> [refinements]\n", argv[0]); return 1; }
  try {
    Grid grid(0);
    grid.readConfigFile(argv[1]);                       // Sysparams::readConfigFile
    grid.readGmsh(argv[2]);                             // GmshReader<UGGrid<2>>::read + createGrid
    if (argc > 3) grid.globalRefine(std::stoi(argv[3]));
    grid.finalize();
    pnp_newton_opts fromcfg; pnp_newton_opts_from_params(grid.ctx(), &fromcfg);
    auto configure = [&](auto& newton) {                // stationary_pnp_from_pb.hh:172-181
      newton.setLineSearchStrategy(newton.hackbuschReuskenAcceptBest);
      newton.setReassembleThreshold(fromcfg.reassemble_threshold);
      newton.setVerbosityLevel(fromcfg.verbosity);
      newton.setReduction(fromcfg.reduction);
      newton.setMinLinearReduction(fromcfg.min_linear_reduction);
      newton.setMaxIterations(fromcfg.max_iterations);
      newton.setLineSearchMaxIterations(fromcfg.line_search_max_iterations);
    };
    // --- PB stage (:105-185)
    GridOperator<PBOperator> pbgo(grid, 0);
    Vector pbu(grid, 1, 0.0);
    ISTLBackend_NOVLP_BCGS_AMG pbls(grid, 2, 5000, 0);
    Newton<GridOperator<PBOperator>, ISTLBackend_NOVLP_BCGS_AMG> pbnewton(pbgo, pbu, pbls);
    configure(pbnewton);
    pbnewton.apply();
    // --- initial guess + Dirichlet values (:235-270)
    Vector phi(grid, 1), cp(grid, 1), cm(grid, 1), u(grid, 3);
    interpolate_bcext(grid, 0, &pbu, phi); interpolate_bcext(grid, 1, &pbu, cp); interpolate_bcext(grid, 2, &pbu, cm);
    check(grid.ctx(), pnp_vec_pack3(grid.ctx(), u.handle(), phi.handle(), cp.handle(), cm.handle()));
    // --- monolithic PNP Newton (:310-360)
    GridOperator<PnpOperator> go(grid);
    ISTLBackend_NOVLP_BCGS_AMG ls(grid, 2, 20000, 0);
    Newton<GridOperator<PnpOperator>, ISTLBackend_NOVLP_BCGS_AMG> newton(go, u, ls);
    configure(newton);
    newton.apply();
    const pnp_newton_result& r = newton.result();
    std::printf("PNP Newton: %d iterations, defect %.3e -> %.3e, %d linear iterations, %.3f s\n", r.iterations,
                r.first_defect, r.defect, r.linear_iterations, r.seconds_total);
  } catch (const Exception& e) {
    std::printf("Something has happened: %s\n", e.what());  // the reference swallows Newton errors the same way
    return 2;
  }
  return 0;
}
