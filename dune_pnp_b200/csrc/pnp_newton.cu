// pnp_newton.cu -- Dune::PDELab::Newton and StationaryLinearProblemSolver on the device.
//
// Flow restated from PDELab 1.1 newton.hh / linearproblem.hh as the
// reference drives them: /root/reference/src/stationary_pnp_from_pb.hh:172-185 (PB) and :344-360
// (PNP); instationary_pnp_from_pb_md.hh:214-228, :349-350.  u, z, r, the trial iterates of the
// Hackbusch-Reusken line search and the matrix stay in HBM; only the scalar defect norms and the
// Krylov scalars reach the host.
#include <chrono>
#include <cmath>

#include "pnp_common.cuh"

namespace pnp {

namespace {
double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
} // namespace

int newton_apply(Ctx& c, const Operator& op, Vec& u, Solver& S, const pnp_newton_opts& o, pnp_newton_result& R) {
  const int F = op_fields(op.op);
  PNP_REQUIRE(u.fields == F, PNP_E_ARG, "vector field count does not match the operator");
  const long n = c.n_own * F, nall = c.nv * F;
  R = pnp_newton_result();
  Vec &r = c.ws_r, &z = c.ws_z, &prev_u = c.ws_prev;
  Matrix& A = c.ws_A;
  r.fields = z.fields = prev_u.fields = F;
  if (r.d.n != (size_t)nall) { r.d.alloc(nall); z.d.alloc(nall); prev_u.d.alloc(nall); }
  r.d.zero(c.stream); z.d.zero(c.stream);
  A.op = op.op; A.nplanes = op_planes(op.op);
  if (A.vals.n != (size_t)A.nplanes * c.nslots) A.vals.alloc((size_t)A.nplanes * c.nslots);
  const double t_start = now();
  auto sync = [&] { PNP_CUDA(cudaStreamSynchronize(c.stream)); };
  auto defect = [&]() {
    const double t0 = now();
    assemble_residual(c, op, u, r);
    const double d = vec_norm(c, r.d.p, n);
    R.seconds_assembly += now() - t0;
    R.residual_assemblies++;
    return d;
  };
  auto record = [&](double d, int lin_its) {
    if (R.n_history < 64) { R.defect_history[R.n_history] = d; R.linear_iterations_history[R.n_history] = lin_its; R.n_history++; }
  };
  int status = PNP_OK;
  R.defect = defect();
  R.first_defect = R.defect;
  double prev_defect = R.defect;
  record(R.defect, 0);
  if (!std::isfinite(R.defect)) { R.seconds_total = now() - t_start; return PNP_E_NAN; }
  if (o.verbosity >= 2) std::printf("  Initial defect: %12.4e\n", R.defect);
  while (true) {
    R.converged = R.defect < o.abs_limit || R.defect < R.first_defect * o.reduction;
    if (R.converged) break;
    if (R.iterations >= o.max_iterations) { status = PNP_E_NOT_CONVERGED; break; }
    // prepare_step
    bool reassembled = false;
    if (R.defect / prev_defect > o.reassemble_threshold || R.jacobian_assemblies == 0) {
      const double t0 = now();
      assemble_jacobian(c, op, u, A, o.jac_mode, o.fd_epsilon);
      sync();
      R.seconds_assembly += now() - t0;
      R.jacobian_assemblies++;
      reassembled = true;
    }
    const double stop_defect = std::max(R.first_defect * o.reduction, o.abs_limit);
    const double ratio2 = R.defect * R.defect / (prev_defect * prev_defect);
    const double linear_reduction =
        stop_defect / (10 * R.defect) > ratio2 ? stop_defect / (10 * R.defect) : std::min(o.min_linear_reduction, ratio2);
    prev_defect = R.defect;
    // linearSolve
    vec_zero(c, z.d.p, n);
    const double t1 = now();
    const LinResult lr = solver_apply(c, S, A, z, r, linear_reduction);
    R.seconds_solve += now() - t1;
    R.linear_iterations += lr.iterations;
    if (o.verbosity >= 3)
      std::printf("  linear solve: %d its, reduction %.3e (asked %.3e)\n", lr.iterations, lr.reduction, linear_reduction);
    if (!lr.converged) {
      record(R.defect, lr.iterations);
      status = lr.status == PNP_E_BREAKDOWN ? PNP_E_BREAKDOWN : PNP_E_LINEAR_SOLVER;
      break;
    }
    // line_search, strategy hackbuschReuskenAcceptBest
    double lambda = 1.0, best_lambda = 0.0, best_defect = R.defect;
    vec_copy(c, u.d.p, prev_u.d.p, n);
    int i = 0;
    bool ls_failed = false;
    while (true) {
      vec_axpy(c, -lambda, z.d.p, u.d.p, n);
      R.defect = defect();
      R.line_search_trials++;
      const bool finite = std::isfinite(R.defect);
      if (finite && R.defect <= (1.0 - lambda / 4) * prev_defect) break;
      if (finite && R.defect < best_defect) { best_defect = R.defect; best_lambda = lambda; }
      if (++i >= o.line_search_max_iterations) {
        if (best_lambda == 0.0) {
          vec_copy(c, prev_u.d.p, u.d.p, n);
          R.defect = defect();
          ls_failed = true;
          break;
        }
        if (best_lambda != lambda) {
          vec_copy(c, prev_u.d.p, u.d.p, n);
          vec_axpy(c, -best_lambda, z.d.p, u.d.p, n);
          R.defect = defect();
        }
        break;
      }
      lambda *= o.damping;
      vec_copy(c, prev_u.d.p, u.d.p, n);
    }
    if (ls_failed) {
      if (reassembled) { status = PNP_E_LINE_SEARCH; break; }
      continue; // retry with a freshly assembled matrix
    }
    R.reduction = R.defect / R.first_defect;
    R.iterations++;
    record(R.defect, lr.iterations);
    if (o.verbosity >= 2)
      std::printf("  Newton iteration %2d.  New defect: %12.4e.  Reduction (total): %12.4e  lambda %g  lin its %d\n",
                  R.iterations, R.defect, R.reduction, lambda, lr.iterations);
  }
  sync();
  R.seconds_total = now() - t_start;
  return status;
}

// StationaryLinearProblemSolver::apply: A = J(u); r = R(u); solve A z = r to `reduction`; u -= z
LinResult slp_apply(Ctx& c, const Operator& op, Vec& u, Solver& S, double reduction, int jac_mode, double eps) {
  const int F = op_fields(op.op);
  PNP_REQUIRE(u.fields == F, PNP_E_ARG, "vector field count does not match the operator");
  const long n = c.n_own * F, nall = c.nv * F;
  Vec &r = c.ws_r, &z = c.ws_z;
  Matrix& A = c.ws_A;
  r.fields = z.fields = F;
  if (r.d.n != (size_t)nall) { r.d.alloc(nall); z.d.alloc(nall); c.ws_prev.d.alloc(nall); }
  r.d.zero(c.stream); z.d.zero(c.stream);
  A.op = op.op; A.nplanes = op_planes(op.op);
  if (A.vals.n != (size_t)A.nplanes * c.nslots) A.vals.alloc((size_t)A.nplanes * c.nslots);
  assemble_jacobian(c, op, u, A, jac_mode, eps);
  assemble_residual(c, op, u, r);
  vec_zero(c, z.d.p, n);
  LinResult lr = solver_apply(c, S, A, z, r, reduction);
  vec_axpy(c, -1.0, z.d.p, u.d.p, n);
  PNP_CUDA(cudaStreamSynchronize(c.stream));
  return lr;
}

} // namespace pnp
