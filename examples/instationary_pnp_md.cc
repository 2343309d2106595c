// instationary_pnp_md -- the driver the reference binary runs at HEAD (/root/reference/src/instationary_pnp_from_pb_md.hh:112-455,
// selected in pnp_solver_main.cc:116) written against the facade: PB Newton solve, BCExtension interpolation, then the
// operator-split time loop: two Alexander2 transport steps (c+, c-) and a linear Poisson update per time step.
// File output (DataWriter, VTK, current.dat) stays on the host side of the reference and is left out.
// Build: see INTEGRATION.md.
#include <cstdio>
#include <string>

#include "pnp_b200/pdelab_facade.hh"

using namespace Dune::PNPB200;

int main(int argc, char** argv) {
  if (argc < 3) { std::printf("usage: %s <config.cfg> <mesh.msh> [refinements] [nSteps]\n", argv[0]); return 1; }
  try {
    Grid grid(0);
    grid.readConfigFile(argv[1]);
    grid.readGmsh(argv[2]);
    if (argc > 3) grid.globalRefine(std::stoi(argv[3]));
    grid.finalize();
    double sys[16];
    check(grid.ctx(), pnp_params_get(grid.ctx(), sys, nullptr, nullptr, 0));
    const double tau = sys[11];
    const int nSteps = argc > 4 ? std::stoi(argv[4]) : (int)sys[12];
    const int potentialUpdateFreq = (int)sys[14];
    const unsigned maxit = (unsigned)sys[5];
    typedef ISTLBackend_NOVLP_BCGS_SSORk PbLS;           // LINEARSOLVER == 1 (:188-191)
    PbLS pbls(grid, maxit, 1, 0);
    // --- PB Newton (:213-228)
    GridOperator<PBOperator> pbgo(grid, 0);
    Vector pbu(grid, 1, 0.0);
    Newton<GridOperator<PBOperator>, PbLS> pbnewton(pbgo, pbu, pbls);
    pnp_newton_opts o; pnp_newton_opts_from_params(grid.ctx(), &o);
    pbnewton.setLineSearchStrategy(pbnewton.hackbuschReuskenAcceptBest);
    pbnewton.setReassembleThreshold(o.reassemble_threshold);
    pbnewton.setReduction(o.reduction);
    pbnewton.setMinLinearReduction(o.min_linear_reduction);
    pbnewton.setMaxIterations(o.max_iterations);
    pbnewton.setLineSearchMaxIterations(o.line_search_max_iterations);
    pbnewton.apply();
    // --- initial state and Dirichlet values: interpolate(phiB / cpB / cmB) (:329-331)
    Vector uphi(grid, 1), ucp(grid, 1), ucm(grid, 1), cpB(grid, 1), cmB(grid, 1), ucpNew(grid, 1), ucmNew(grid, 1);
    interpolate_bcext(grid, 0, &pbu, uphi);
    interpolate_bcext(grid, 1, &pbu, ucp); interpolate_bcext(grid, 1, &pbu, cpB);
    interpolate_bcext(grid, 2, &pbu, ucm); interpolate_bcext(grid, 2, &pbu, cmB);
    // --- Poisson problem for the potential (:343-350)
    GridOperator<PoissonOperator> phigo(grid, 0);
    phigo.setCoefficient(0, ucp); phigo.setCoefficient(1, ucm);
    StationaryLinearProblemSolver<GridOperator<PoissonOperator>, PbLS> slp(phigo, uphi, pbls, 1e-10);
    // --- transport problems (:357-391); both species use the c+ constraints (cpB_t), quirk kept
    GridOperator<DiffusionOperator> cpgo0(grid, 1), cmgo0(grid, 1);
    cpgo0.setCoefficient(0, uphi); cpgo0.setValency(1.0);
    cmgo0.setCoefficient(0, uphi); cmgo0.setValency(-1.0);
    GridOperator<DiffusionTOperator> cpgo1(grid, 1), cmgo1(grid, 1);
    typedef OneStepGridOperator<GridOperator<DiffusionOperator>, GridOperator<DiffusionTOperator>> IGO;
    IGO cpigo(cpgo0, cpgo1), cmigo(cmgo0, cmgo1);
    Alexander2Parameter method;
    OneStepMethod<Alexander2Parameter, IGO, PbLS> cposm(method, cpigo, pbls, 1e-5), cmosm(method, cmigo, pbls, 1e-5);
    // --- time loop (:421-454)
    double time = 0.0;
    for (int i = 0; i < nSteps; i++) {
      cposm.apply(time, tau, ucp, cpB, ucpNew);
      check(grid.ctx(), pnp_vec_copy(grid.ctx(), ucp.handle(), ucpNew.handle()));
      cmosm.apply(time, tau, ucm, cmB, ucmNew);
      check(grid.ctx(), pnp_vec_copy(grid.ctx(), ucm.handle(), ucmNew.handle()));
      time += tau;
      if (i % potentialUpdateFreq == 0) slp.apply();
      std::printf("step %4d  t = %-8g  |phi| %.10e  |c+| %.10e  |c-| %.10e  (stage its %d %d / %d %d)\n", i, time, uphi.two_norm(),
                  ucp.two_norm(), ucm.two_norm(), cposm.stageResults()[0].iterations, cposm.stageResults()[1].iterations,
                  cmosm.stageResults()[0].iterations, cmosm.stageResults()[1].iterations);
    }
    slp.apply();
  } catch (const Exception& e) {
    std::printf("Something has happened: %s\n", e.what());
    return 2;
  }
  return 0;
}
