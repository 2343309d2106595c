// pdelab_facade.hh -- header-only C++ facade over the C ABI (pnp_b200.h) that keeps the PDELab / ISTL names the
// dune-pnp drivers use, so that stationary_pnp / stationary_pnp_from_pb / instationary_pnp_md read the same after the
// backend switch.  Reference call sites (relative to /root/reference/src):
//   GridOperator<...>(gfsu,cu,gfsv,cv,lop), go.residual(u,r), go.jacobian(u,A)      stationary_pnp.hh:240-246
//   ISTLBackend_NOVLP_BCGS_SSORk / _BCGS_NOPREC / _CG_NOPREC / _CG_Jacobi / _CG_AMG_SSOR   instationary_pnp_from_pb_md.hh:188-211
//   Newton<GO,LS,U>(go,u,ls) + setters + apply()                                    stationary_pnp.hh:280-294
//   StationaryLinearProblemSolver<GO,LS,U>(go,u,ls,red).apply()                      instationary_pnp_from_pb_md.hh:349-350
//   interpolate(BCExtension, gfs, u)                                                 stationary_pnp_from_pb.hh:270
//   OneStepGridOperator<GO0,GO1>, OneStepMethod<...>(Alexander2Parameter, igo, pdesolver).apply(t,dt,x,f,xnew)   instationary_pnp_from_pb_md.hh:368-391,421-425
// Errors are rethrown as the same-named exception types PDELab throws.
#pragma once
#include <stdexcept>
#include <string>
#include <vector>

#include "../pnp_b200.h"

namespace Dune {
namespace PNPB200 {

struct Exception : std::runtime_error { using std::runtime_error::runtime_error; };
struct NewtonNotConverged : Exception { using Exception::Exception; };
struct NewtonLinearSolverError : Exception { using Exception::Exception; };
struct NewtonLineSearchError : Exception { using Exception::Exception; };
struct NewtonDefectError : Exception { using Exception::Exception; };
struct ISTLError : Exception { using Exception::Exception; };

inline void check(pnp_ctx* c, pnp_status st) {
  if (st == PNP_OK) return;
  const std::string m = pnp_last_error(c);
  switch (st) {
    case PNP_E_NOT_CONVERGED: throw NewtonNotConverged(m);
    case PNP_E_LINEAR_SOLVER: throw NewtonLinearSolverError(m);
    case PNP_E_LINE_SEARCH: throw NewtonLineSearchError(m);
    case PNP_E_NAN: throw NewtonDefectError(m);
    case PNP_E_BREAKDOWN: throw ISTLError(m);
    default: throw Exception(m);
  }
}

// Grid + Sysparams: GmshReader/GridFactory (pnp_solver_main.cc:82-114) and Sysparams::readConfigFile
class Grid {
 public:
  explicit Grid(int device = 0) { if (pnp_ctx_create(device, &c_) != PNP_OK) throw Exception("no CUDA device"); }
  ~Grid() { pnp_ctx_destroy(c_); }
  Grid(const Grid&) = delete;
  Grid& operator=(const Grid&) = delete;
  void readConfigFile(const std::string& cfg) { check(c_, pnp_params_read(c_, cfg.c_str())); }
  void readGmsh(const std::string& msh) { check(c_, pnp_mesh_read_gmsh(c_, msh.c_str())); }
  void globalRefine(int levels) { check(c_, pnp_mesh_refine(c_, levels)); }
  // Pk2DLocalFiniteElementMap<GV, D, R, PDEGREE> (instationary_pnp_from_pb_md.hh:26-28,125): 1 (default), 2 or 3, before finalize()
  void setDegree(int pdegree) { check(c_, pnp_space_set_degree(c_, pdegree)); }
  void finalize(bool renumber = true) { check(c_, pnp_mesh_finalize(c_, renumber)); }
  // gfs.size() per field: vertices, or edges + vertices for quadratic elements
  long dofs() const { long nd = 0; check(c_, pnp_space_sizes(c_, nullptr, nullptr, &nd)); return nd; }
  long size() const { long nv = 0; pnp_mesh_sizes(c_, &nv, nullptr, nullptr, nullptr); return nv; }
  long ownedSize() const { long n = 0; pnp_mesh_owned(c_, &n); return n; }
  // MPIHelper + grid->loadBalance() (pnp_solver_main.cc:93-108): one process per GPU.  The NCCL id travels through a file
  // all ranks see (no MPI needed); loadBalance partitions the Gmsh mesh this grid holds, refines this rank's part `levels`
  // times and registers the coarser levels as multigrid levels (replicated below `replicaLevel`).
  void initCommunication(int rank, int world, const std::string& rendezvousFile) {
    check(c_, pnp_comm_init_file(c_, rank, world, rendezvousFile.c_str()));
  }
  // returns the handle of a start vector when a field was given (its values on the unpartitioned level `fieldLevel` mesh)
  int loadBalance(int levels, int replicaLevel, int fieldLevel = 0, int nfields = 0, const std::vector<double>* fx = nullptr,
                  const std::vector<double>* fy = nullptr, const std::vector<double>* fields = nullptr) {
    int start = -1;
    check(c_, pnp_partition_build(c_, levels, replicaLevel, fieldLevel, nfields, fx ? (long)fx->size() : 0, fx ? fx->data() : nullptr,
                                  fy ? fy->data() : nullptr, fields ? fields->data() : nullptr, &start));
    return start;
  }
  std::vector<double> coordinates(int which) const { // 0: x, 1: y (reference numbering)
    std::vector<double> v(size());
    check(c_, which == 0 ? pnp_mesh_get(c_, v.data(), nullptr, nullptr, nullptr, nullptr, nullptr)
                         : pnp_mesh_get(c_, nullptr, v.data(), nullptr, nullptr, nullptr, nullptr));
    return v;
  }
  pnp_ctx* ctx() const { return c_; }
 private:
  pnp_ctx* c_ = nullptr;
};

// BackendVectorSelector<GFS,double>::Type
class Vector {
 public:
  Vector(Grid& g, int fields, double value = 0.0) : g_(g), fields_(fields) {
    check(g.ctx(), pnp_vec_create(g.ctx(), fields, &h_));
    if (value != 0.0) check(g.ctx(), pnp_vec_set(g.ctx(), h_, value));
  }
  Vector(Grid& g, int fields, int existingHandle, bool) : g_(g), fields_(fields), h_(existingHandle) {} // adopts a library-made vector
  ~Vector() { pnp_vec_destroy(g_.ctx(), h_); }
  Vector(const Vector&) = delete;
  void set(const std::vector<double>& host) { check(g_.ctx(), pnp_vec_upload(g_.ctx(), h_, host.data())); }
  std::vector<double> get() const { std::vector<double> v(fields_ * g_.dofs()); check(g_.ctx(), pnp_vec_download(g_.ctx(), h_, v.data())); return v; }
  void axpy(double a, const Vector& x) { check(g_.ctx(), pnp_vec_axpy(g_.ctx(), h_, a, x.h_)); }
  double two_norm() const { double n; check(g_.ctx(), pnp_vec_norm(g_.ctx(), h_, &n)); return n; }
  int handle() const { return h_; }
  int fields() const { return fields_; }
  Grid& grid() const { return g_; }
 private:
  Grid& g_; int fields_; int h_ = -1;
};

// local operator tags (files of the same name in the reference)
struct PBOperator { static constexpr int op = PNP_OP_PB, fields = 1; };
struct PoissonOperator { static constexpr int op = PNP_OP_POISSON, fields = 1; };
struct DiffusionOperator { static constexpr int op = PNP_OP_DIFFUSION, fields = 1; };
struct DiffusionTOperator { static constexpr int op = PNP_OP_MASS, fields = 1; };
struct PnpOperator { static constexpr int op = PNP_OP_PNP, fields = 3; };

template <class LOP> class GridOperator;

// GridOperator::Traits::Jacobian / MatrixContainer
template <class LOP> class Matrix {
 public:
  explicit Matrix(GridOperator<LOP>& go);
  ~Matrix() { pnp_matrix_destroy(g_.ctx(), h_); }
  int handle() const { return h_; }
 private:
  Grid& g_; int h_ = -1;
};

template <class LOP> class GridOperator {
 public:
  // bcComponent: the BCType<...,component> a scalar operator is built with (btype.hh:5)
  GridOperator(Grid& g, int bcComponent = 0) : g_(g) {
    check(g.ctx(), pnp_operator_create(g.ctx(), LOP::op, bcComponent, &h_));
#ifdef PNP_INTORDER  // build-wide `intorder` of the local operators (their last constructor argument), e.g. -DPDEGREE=3 -DPNP_INTORDER=5
    setIntegrationOrder(PNP_INTORDER);
#endif
  }
  // the local operator's `intorder` constructor argument (pb_operator.hh:39): 0 = the reference drivers' default, 5 for degree >= 2
  void setIntegrationOrder(int intorder) { check(g_.ctx(), pnp_operator_set_intorder(g_.ctx(), h_, intorder)); }
  void setCoefficient(int which, const Vector& v) { check(g_.ctx(), pnp_operator_set_coefficient(g_.ctx(), h_, which, v.handle())); }
  void setValency(double z) { check(g_.ctx(), pnp_operator_set_valency(g_.ctx(), h_, z)); }
  void residual(const Vector& u, Vector& r) const { check(g_.ctx(), pnp_residual(g_.ctx(), h_, u.handle(), r.handle())); }
  void jacobian(const Vector& u, Matrix<LOP>& A, int mode = PNP_JAC_FD_FAITHFUL, double eps = 1e-11) const {
    check(g_.ctx(), pnp_jacobian(g_.ctx(), h_, u.handle(), A.handle(), mode, eps));
  }
  Grid& grid() const { return g_; }
  int handle() const { return h_; }
 private:
  Grid& g_; int h_ = -1;
};
template <class LOP> Matrix<LOP>::Matrix(GridOperator<LOP>& go) : g_(go.grid()) { check(g_.ctx(), pnp_matrix_create(g_.ctx(), go.handle(), &h_)); }

// ISTLBackend_NOVLP_*: apply(A,z,r,reduction), result()
struct LinearSolverResult { bool converged = false; int iterations = 0; double reduction = 1, conv_rate = 1, elapsed = 0; };
class LinearSolverBackend {
 public:
  LinearSolverBackend(Grid& g, int kind, int prec, unsigned maxiter, int steps, int verbose) : g_(g) {
    check(g.ctx(), pnp_solver_create(g.ctx(), kind, prec, (int)maxiter, steps, verbose, &h_));
  }
  template <class M> void apply(M& A, Vector& z, Vector& r, double reduction) {
    pnp_lin_result lr{};
    check(g_.ctx(), pnp_solver_apply(g_.ctx(), h_, A.handle(), z.handle(), r.handle(), reduction, &lr));
    res_ = {lr.converged != 0, lr.iterations, lr.reduction, lr.conv_rate, lr.seconds};
  }
  double norm(const Vector& v) const { return v.two_norm(); }
  const LinearSolverResult& result() const { return res_; }
  int handle() const { return h_; }
 private:
  Grid& g_; int h_ = -1; LinearSolverResult res_;
};
struct ISTLBackend_NOVLP_BCGS_NOPREC : LinearSolverBackend {
  ISTLBackend_NOVLP_BCGS_NOPREC(Grid& g, unsigned maxiter = 5000, int verbose = 1) : LinearSolverBackend(g, PNP_SOLVER_BCGS, PNP_PREC_NONE, maxiter, 1, verbose) {}
};
struct ISTLBackend_NOVLP_CG_NOPREC : LinearSolverBackend {
  ISTLBackend_NOVLP_CG_NOPREC(Grid& g, unsigned maxiter = 5000, int verbose = 1) : LinearSolverBackend(g, PNP_SOLVER_CG, PNP_PREC_NONE, maxiter, 1, verbose) {}
};
struct ISTLBackend_NOVLP_CG_Jacobi : LinearSolverBackend {
  ISTLBackend_NOVLP_CG_Jacobi(Grid& g, unsigned maxiter = 5000, int verbose = 1) : LinearSolverBackend(g, PNP_SOLVER_CG, PNP_PREC_JACOBI, maxiter, 1, verbose) {}
};
struct ISTLBackend_NOVLP_BCGS_SSORk : LinearSolverBackend {
  ISTLBackend_NOVLP_BCGS_SSORk(Grid& g, unsigned maxiter = 5000, int steps = 5, int verbose = 1) : LinearSolverBackend(g, PNP_SOLVER_BCGS, PNP_PREC_SSOR, maxiter, steps, verbose) {}
};
struct ISTLBackend_NOVLP_CG_AMG_SSOR : LinearSolverBackend {
  // the backend's smoother is SeqSSOR: on one GPU the multigrid smooths with it (amg_smoother = 2); a partitioned grid keeps the
  // damped point-block Jacobi smoother (a Gauss-Seidel sweep is sequential over one GPU's rows).  setSmoother(0) for the fast one.
  ISTLBackend_NOVLP_CG_AMG_SSOR(Grid& g, int smoothsteps = 2, unsigned maxiter = 5000, int verbose = 1) : LinearSolverBackend(g, PNP_SOLVER_CG, PNP_PREC_AMG, maxiter, smoothsteps, verbose), grid_(g) {
    if (g.ownedSize() == g.size()) setSmoother(2);
  }
  void setSmoother(int kind) { check(grid_.ctx(), pnp_solver_set_option(grid_.ctx(), handle(), "amg_smoother", (double)kind)); }
 private:
  Grid& grid_;
};
struct ISTLBackend_NOVLP_BCGS_AMG : LinearSolverBackend {  // new: what bench.py runs for the non-symmetric PNP system
  ISTLBackend_NOVLP_BCGS_AMG(Grid& g, int smoothsteps = 2, unsigned maxiter = 5000, int verbose = 1) : LinearSolverBackend(g, PNP_SOLVER_BCGS, PNP_PREC_AMG, maxiter, smoothsteps, verbose) {}
};

struct ISTLBackend_NOVLP_BCGS_ILU0 : LinearSolverBackend {  // new (north star): SeqILU0 in the reference's row order
  ISTLBackend_NOVLP_BCGS_ILU0(Grid& g, unsigned maxiter = 5000, int verbose = 1) : LinearSolverBackend(g, PNP_SOLVER_BCGS, PNP_PREC_ILU0, maxiter, 1, verbose) {}
};

// Dune::PDELab::Newton
template <class GO, class LS> class Newton {
 public:
  enum Strategy { noLineSearch, hackbuschReusken, hackbuschReuskenAcceptBest };
  Newton(GO& go, Vector& u, LS& ls) : go_(go), u_(u), ls_(ls) {
    pnp_newton_opts_default(&o_);
    o_.line_search_strategy = PNP_LS_HACKBUSCH_REUSKEN; // PDELab's default until setLineSearchStrategy is called
  }
  void setLineSearchStrategy(Strategy s) {
    o_.line_search_strategy = s == noLineSearch ? PNP_LS_NONE : (s == hackbuschReusken ? PNP_LS_HACKBUSCH_REUSKEN : PNP_LS_HACKBUSCH_REUSKEN_ACCEPT_BEST);
  }
  void setReassembleThreshold(double v) { o_.reassemble_threshold = v; }
  void setVerbosityLevel(int v) { o_.verbosity = v; }
  void setReduction(double v) { o_.reduction = v; }
  void setMinLinearReduction(double v) { o_.min_linear_reduction = v; }
  void setMaxIterations(unsigned v) { o_.max_iterations = (int)v; }
  void setLineSearchMaxIterations(unsigned v) { o_.line_search_max_iterations = (int)v; }
  void setJacobianMode(int mode, double eps = 1e-11) { o_.jac_mode = mode; o_.fd_epsilon = eps; }
  void apply() { check(go_.grid().ctx(), pnp_newton_apply(go_.grid().ctx(), go_.handle(), u_.handle(), ls_.handle(), &o_, &r_)); }
  const pnp_newton_result& result() const { return r_; }
 private:
  GO& go_; Vector& u_; LS& ls_; pnp_newton_opts o_; pnp_newton_result r_{};
};

template <class GO, class LS> class StationaryLinearProblemSolver {
 public:
  StationaryLinearProblemSolver(GO& go, Vector& u, LS& ls, double reduction) : go_(go), u_(u), ls_(ls), red_(reduction) {}
  void apply(int jac_mode = PNP_JAC_FD_FAITHFUL, double eps = 1e-11) {
    pnp_lin_result lr{};
    check(go_.grid().ctx(), pnp_slp_apply(go_.grid().ctx(), go_.handle(), u_.handle(), ls_.handle(), red_, jac_mode, eps, &lr));
  }
 private:
  GO& go_; Vector& u_; LS& ls_; double red_;
};

// Alexander2Parameter / ImplicitEulerParameter + OneStepGridOperator<GO0,GO1> + OneStepMethod<Real,IGO,PDESOLVER,U,U>
// (instationary_pnp_from_pb_md.hh:368-391).  The stage solver is the StationaryLinearProblemSolver the reference builds
// with (igo, ls, 1e-5); it is folded into the method here: OneStepMethod(method, igo, ls, reduction).
struct Alexander2Parameter { static constexpr int method = PNP_TIME_ALEXANDER2; };
struct ImplicitEulerParameter { static constexpr int method = PNP_TIME_IMPLICIT_EULER; };
template <class GO0, class GO1> struct OneStepGridOperator {
  OneStepGridOperator(GO0& go0_, GO1& go1_) : go0(go0_), go1(go1_) {}
  GO0& go0; GO1& go1;
};
template <class Method, class IGO, class LS> class OneStepMethod {
 public:
  OneStepMethod(const Method&, IGO& igo, LS& ls, double reduction) : igo_(igo), ls_(ls), red_(reduction) {}
  void setJacobianMode(int mode, double eps = 1e-11) { mode_ = mode; eps_ = eps; }
  // apply(time, dt, xold, f, xnew): f = boundary function interpolated to the dofs (Dirichlet values)
  void apply(double time, double dt, const Vector& xold, const Vector& f, Vector& xnew) {
    pnp_ctx* c = igo_.go0.grid().ctx();
    check(c, pnp_onestep_apply(c, Method::method, igo_.go0.handle(), igo_.go1.handle(), ls_.handle(), time, dt, xold.handle(),
                               f.handle(), xnew.handle(), red_, mode_, eps_, stage_));
  }
  const pnp_lin_result* stageResults() const { return stage_; }
 private:
  IGO& igo_; LS& ls_; double red_; int mode_ = PNP_JAC_FD_FAITHFUL; double eps_ = 1e-11; pnp_lin_result stage_[2] = {};
};

// DataWriter<GV>::writeData(gfs, u, filename) (datawriter.hh:45-94) and calcIonFlux(...) (ionFlux.hh:8-96)
inline void writeData(const Vector& u, const std::string& filename) {
  pnp_ctx* c = u.grid().ctx();
  check(c, pnp_write_cell_data(c, u.handle(), filename.c_str()));
}
inline void calcIonFlux(const Vector& uphi, const Vector& ucp, const Vector& ucm, std::vector<double>& ip, std::vector<double>& im) {
  pnp_ctx* c = uphi.grid().ctx();
  double sys[16]; check(c, pnp_params_get(c, sys, nullptr, nullptr, 0));
  ip.assign((size_t)sys[0], 0.0); im.assign((size_t)sys[0], 0.0);
  check(c, pnp_ion_flux(c, uphi.handle(), ucp.handle(), ucm.handle(), ip.data(), im.data()));
}

// interpolate(BCExtension<...,component,PbDGF>, gfs, u)
inline void interpolate_bcext(Grid& g, int component, const Vector* pb, Vector& out) {
  check(g.ctx(), pnp_interpolate_bcext(g.ctx(), component, pb ? pb->handle() : -1, out.handle()));
}

}  // namespace PNPB200
}  // namespace Dune
