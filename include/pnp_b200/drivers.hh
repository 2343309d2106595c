// drivers.hh -- the reference's driver functions on the facade (pdelab_facade.hh), same names and flow:
//   stationary_pnp          /root/reference/src/stationary_pnp.hh:91-365        interpolate(BCExtension) -> PNP Newton
//   stationary_pnp_from_pb  stationary_pnp_from_pb.hh:92-440 (named stationary_pnp there)   PB Newton -> interpolate -> PNP Newton
//   instationary_pnp_md     instationary_pnp_from_pb_md.hh:112-455   PB Newton -> interpolate -> operator-split time loop
//   PnpSolverMain::run      pnp_solver_main.cc:70-116                config -> Gmsh mesh -> instationary_pnp_md
// What the reference does around these calls on the host (DataWriter / VTK files, calcIonFlux -> current.dat) is reached
// through the `output` callback of instationary_pnp_md; the solver path itself never leaves the device.
#pragma once
#include <functional>
#include <string>

#include "pdelab_facade.hh"

// polynomial degree of the Pk2DLocalFiniteElementMap, a compile-time switch as in the reference
// (instationary_pnp_from_pb_md.hh:26-28; src/Makefile.am:54-110 builds every program with -DPDEGREE=1,2,3): 1 or 2 here
#ifndef PDEGREE
#define PDEGREE 1
#endif

namespace Dune {
namespace PNPB200 {

// Newton configured from Sysparams exactly as the drivers do (stationary_pnp.hh:280-288)
template <class NEWTON> void configure_newton(Grid& grid, NEWTON& newton) {
  pnp_newton_opts o; check(grid.ctx(), pnp_newton_opts_from_params(grid.ctx(), &o));
  newton.setLineSearchStrategy(newton.hackbuschReuskenAcceptBest);
  newton.setReassembleThreshold(o.reassemble_threshold);
  newton.setVerbosityLevel(o.verbosity);
  newton.setReduction(o.reduction);
  newton.setMinLinearReduction(o.min_linear_reduction);
  newton.setMaxIterations(o.max_iterations);
  newton.setLineSearchMaxIterations(o.line_search_max_iterations);
}
inline double sysparam(Grid& grid, int index) { // pnp_params_get layout (pnp_b200.h)
  double sys[16]; check(grid.ctx(), pnp_params_get(grid.ctx(), sys, nullptr, nullptr, 0)); return sys[index];
}

// u (3 fields) = interpolate(BCExtension) of the PB field `pb` (nullptr: zero PB field)
inline void interpolate_initial(Grid& grid, const Vector* pb, Vector& u) {
  Vector phi(grid, 1), cp(grid, 1), cm(grid, 1);
  interpolate_bcext(grid, 0, pb, phi); interpolate_bcext(grid, 1, pb, cp); interpolate_bcext(grid, 2, pb, cm);
  check(grid.ctx(), pnp_vec_pack3(grid.ctx(), u.handle(), phi.handle(), cp.handle(), cm.handle()));
}

// Monolithic PNP Newton from the interpolated boundary data.  ls: the reference builds ISTLBackend_NOVLP_BCGS_NOPREC
// (gfs, s.linearSolverIterations, s.verbosity) (stationary_pnp.hh:254-256); any backend of the facade works.  Newton
// errors are swallowed as the reference does ("Something has happened", :290-294) unless rethrow is set.
template <class LS>
pnp_newton_result stationary_pnp(Grid& grid, Vector& u, LS& ls, int jac_mode = PNP_JAC_FD_FAITHFUL, bool rethrow = false) {
  interpolate_initial(grid, nullptr, u);
  GridOperator<PnpOperator> go(grid);
  Newton<GridOperator<PnpOperator>, LS> newton(go, u, ls);
  configure_newton(grid, newton);
  newton.setJacobianMode(jac_mode);
  try { newton.apply(); } catch (const Exception&) { if (rethrow) throw; }
  return newton.result();
}

// PB Newton solve into pbu (stationary_pnp_from_pb.hh:105-185)
template <class LS> pnp_newton_result pb_newton(Grid& grid, Vector& pbu, LS& ls, int jac_mode = PNP_JAC_FD_FAITHFUL) {
  GridOperator<PBOperator> pbgo(grid, 0);
  Newton<GridOperator<PBOperator>, LS> newton(pbgo, pbu, ls);
  configure_newton(grid, newton);
  newton.setJacobianMode(jac_mode);
  newton.apply();
  return newton.result();
}

template <class LS>
pnp_newton_result stationary_pnp_from_pb(Grid& grid, Vector& u, LS& pbls, LS& ls, int jac_mode = PNP_JAC_FD_FAITHFUL,
                                         bool rethrow = false) {
  Vector pbu(grid, 1, 0.0);
  pb_newton(grid, pbu, pbls, jac_mode);
  interpolate_initial(grid, &pbu, u);
  GridOperator<PnpOperator> go(grid);
  Newton<GridOperator<PnpOperator>, LS> newton(go, u, ls);
  configure_newton(grid, newton);
  newton.setJacobianMode(jac_mode);
  try { newton.apply(); } catch (const Exception&) { if (rethrow) throw; }
  return newton.result();
}

// The driver of the reference binary: PB Newton, interpolation, then per time step two Alexander2 transport steps and
// (every potentialUpdateFreq steps) a linear Poisson solve; `output(step, time, uphi, ucp, ucm)` is called where the
// reference writes its files (every outputFreq steps, :430-452).  nSteps < 0: Sysparams::nSteps.
template <class LS>
void instationary_pnp_md(Grid& grid, LS& pbls, Vector& uphi, Vector& ucp, Vector& ucm, int nSteps = -1,
                         const std::function<void(int, double, Vector&, Vector&, Vector&)>& output = nullptr,
                         int jac_mode = PNP_JAC_FD_FAITHFUL) {
  const double tau = sysparam(grid, 11);
  if (nSteps < 0) nSteps = (int)sysparam(grid, 12);
  const int outputFreq = (int)sysparam(grid, 13) > 0 ? (int)sysparam(grid, 13) : 1;
  const int potentialUpdateFreq = (int)sysparam(grid, 14) > 0 ? (int)sysparam(grid, 14) : 1;
  Vector pbu(grid, 1, 0.0);
  pb_newton(grid, pbu, pbls, jac_mode);                                       // :213-228
  Vector cpB(grid, 1), cmB(grid, 1), ucpNew(grid, 1), ucmNew(grid, 1);
  interpolate_bcext(grid, 0, &pbu, uphi);                                     // :329-331
  interpolate_bcext(grid, 1, &pbu, ucp); interpolate_bcext(grid, 1, &pbu, cpB);
  interpolate_bcext(grid, 2, &pbu, ucm); interpolate_bcext(grid, 2, &pbu, cmB);
  GridOperator<PoissonOperator> phigo(grid, 0);                               // :343-350
  phigo.setCoefficient(0, ucp); phigo.setCoefficient(1, ucm);
  StationaryLinearProblemSolver<GridOperator<PoissonOperator>, LS> slp(phigo, uphi, pbls, 1e-10);
  GridOperator<DiffusionOperator> cpgo0(grid, 1), cmgo0(grid, 1);             // :357-391; both species on the c+ constraints
  cpgo0.setCoefficient(0, uphi); cpgo0.setValency(1.0);
  cmgo0.setCoefficient(0, uphi); cmgo0.setValency(-1.0);
  GridOperator<DiffusionTOperator> cpgo1(grid, 1), cmgo1(grid, 1);
  typedef OneStepGridOperator<GridOperator<DiffusionOperator>, GridOperator<DiffusionTOperator>> IGO;
  IGO cpigo(cpgo0, cpgo1), cmigo(cmgo0, cmgo1);
  Alexander2Parameter method;
  OneStepMethod<Alexander2Parameter, IGO, LS> cposm(method, cpigo, pbls, 1e-5), cmosm(method, cmigo, pbls, 1e-5);
  cposm.setJacobianMode(jac_mode); cmosm.setJacobianMode(jac_mode);
  double time = 0.0;
  for (int i = 0; i < nSteps; i++) {                                          // :421-453
    cposm.apply(time, tau, ucp, cpB, ucpNew);
    check(grid.ctx(), pnp_vec_copy(grid.ctx(), ucp.handle(), ucpNew.handle()));
    cmosm.apply(time, tau, ucm, cmB, ucmNew);
    check(grid.ctx(), pnp_vec_copy(grid.ctx(), ucm.handle(), ucmNew.handle()));
    time += tau;
    if (i % potentialUpdateFreq == 0) slp.apply(jac_mode);
    if (output && i % outputFreq == 0) output(i, time, uphi, ucp, ucm);
  }
  slp.apply(jac_mode);                                                        // :454
}

// PnpSolverMain::run(configfile): Sysparams -> GmshReader -> instationary_pnp_md (pnp_solver_main.cc:70-116)
class PnpSolverMain {
 public:
  explicit PnpSolverMain(int device = 0) : device_(device) {}
  void run(const std::string& configfile, int refinements = 0, int nSteps = -1,
           const std::function<void(int, double, Vector&, Vector&, Vector&)>& output = nullptr) {
    Grid grid(device_);
    grid.readConfigFile(configfile);
    char meshfile[1024] = {0};
    check(grid.ctx(), pnp_params_get(grid.ctx(), nullptr, nullptr, meshfile, (int)sizeof meshfile));
    grid.readGmsh(meshfile);  // path as written in the config, relative to the working directory (sysparams.cc:44)
    if (refinements > 0) grid.globalRefine(refinements);
    grid.setDegree(PDEGREE);
    grid.finalize();
    Vector uphi(grid, 1), ucp(grid, 1), ucm(grid, 1);
    ISTLBackend_NOVLP_BCGS_SSORk pbls(grid, (unsigned)sysparam(grid, 5), 1, (int)sysparam(grid, 15)); // LINEARSOLVER == 1 (:188-191)
    instationary_pnp_md(grid, pbls, uphi, ucp, ucm, nSteps, output);
  }
 private:
  int device_;
};

}  // namespace PNPB200
}  // namespace Dune
