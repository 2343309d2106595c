// oracle_capi.cpp -- C entry points of the CPU ORACLE for ctypes (test infrastructure only;
// see the header of pnp_oracle.hpp).  Nothing under dune_pnp_b200/ may link or load this.
#include "pnp_oracle.hpp"
#include "pnp_oracle_p2.hpp"
#include <chrono>
#include <cstring>

using namespace pnpo;

namespace {
thread_local std::string g_err;
struct Params { Sysparams s; };
OpCtx make_ctx(const Mesh* m, const Sysparams* s, int op, const double* aux0, const double* aux1,
               double valency, int intorder) {
  OpCtx c; c.op = op; c.s = s; c.m = m; c.valency = valency; c.intorder = intorder;
  if (op == OP_POISSON) { c.cp = aux0; c.cm = aux1; }
  if (op == OP_DIFFUSION) { c.uphi = aux0; }
  return c;
}
int comp0_of(int op) { (void)op; return 0; }
} // namespace

#define ORA_TRY try {
#define ORA_CATCH(rv) } catch (const std::exception& ex) { g_err = ex.what(); return rv; }

extern "C" {

const char* ora_last_error() { return g_err.c_str(); }

void* ora_mesh_create(int nv, const double* x, const double* y, int nT, const int* tri, int nB,
                      const int* ba, const int* bb, const int* bphys) {
  ORA_TRY
  Mesh* m = new Mesh;
  m->x.assign(x, x + nv); m->y.assign(y, y + nv); m->tri.assign(tri, tri + 3 * nT);
  m->ba.assign(ba, ba + nB); m->bb.assign(bb, bb + nB); m->bphys.assign(bphys, bphys + nB);
  m->finalize();
  return m;
  ORA_CATCH(nullptr)
}
void* ora_mesh_read_gmsh(const char* path) {
  ORA_TRY return new Mesh(read_gmsh(path)); ORA_CATCH(nullptr)
}
void* ora_mesh_refine(void* h) {
  ORA_TRY return new Mesh(refine(*(Mesh*)h)); ORA_CATCH(nullptr)
}
void ora_mesh_sizes(void* h, int* nv, int* nT, int* nB) {
  Mesh* m = (Mesh*)h; *nv = m->nv; *nT = m->nT; *nB = m->nB;
}
void ora_mesh_get(void* h, double* x, double* y, int* tri, int* ba, int* bb, int* bphys) {
  Mesh* m = (Mesh*)h;
  std::copy(m->x.begin(), m->x.end(), x); std::copy(m->y.begin(), m->y.end(), y);
  std::copy(m->tri.begin(), m->tri.end(), tri);
  std::copy(m->ba.begin(), m->ba.end(), ba); std::copy(m->bb.begin(), m->bb.end(), bb);
  std::copy(m->bphys.begin(), m->bphys.end(), bphys);
}
void ora_mesh_free(void* h) { delete (Mesh*)h; }

// flat parameter layout shared with the tests:
//  sys[16]  = {n_surfaces, cylindrical, l_b, c0, PI, linearSolverIterations, newtonReassembleThreshold,
//              newtonReduction, newtonMinLinearReduction, newtonMaxIterations,
//              newtonLineSearchMaxIteration, tau, nSteps, outputFreq, potentialUpdateFreq, verbosity}
//  surf[ns][9] = per component {btype, flux, dirichlet value}
void* ora_params_read(const char* path) {
  ORA_TRY Params* p = new Params; p->s = read_config(path); return p; ORA_CATCH(nullptr)
}
void* ora_params_create(const double* sys, const double* surf) {
  ORA_TRY
  Params* p = new Params; Sysparams& s = p->s;
  s.n_surfaces = (int)sys[0]; s.cylindrical = sys[1] != 0; s.l_b = sys[2]; s.c0 = sys[3]; s.PI = sys[4];
  s.linearSolverIterations = (int)sys[5]; s.newtonReassembleThreshold = sys[6]; s.newtonReduction = sys[7];
  s.newtonMinLinearReduction = sys[8]; s.newtonMaxIterations = sys[9]; s.newtonLineSearchMaxIteration = sys[10];
  s.tau = sys[11]; s.nSteps = (int)sys[12]; s.outputFreq = (int)sys[13]; s.potentialUpdateFreq = (int)sys[14];
  s.verbosity = (int)sys[15];
  s.surfaces.assign(s.n_surfaces, Surface());
  for (int i = 0; i < s.n_surfaces; i++) {
    const double* f = surf + 9 * i; Surface& q = s.surfaces[i];
    q.coulombBtype = (int)f[0]; q.coulombFlux = f[1]; q.coulombPotential = f[2];
    q.plusDiffusionBtype = (int)f[3]; q.plusDiffusionFlux = f[4]; q.plusDiffusionConcentration = f[5];
    q.minusDiffusionBtype = (int)f[6]; q.minusDiffusionFlux = f[7]; q.minusDiffusionConcentration = f[8];
  }
  return p;
  ORA_CATCH(nullptr)
}
int ora_params_nsurf(void* h) { return ((Params*)h)->s.n_surfaces; }
void ora_params_get(void* h, double* sys, double* surf, char* meshfile, int meshfile_cap) {
  const Sysparams& s = ((Params*)h)->s;
  double v[16] = {(double)s.n_surfaces, (double)s.cylindrical, s.l_b, s.c0, s.PI, (double)s.linearSolverIterations,
                  s.newtonReassembleThreshold, s.newtonReduction, s.newtonMinLinearReduction, s.newtonMaxIterations,
                  s.newtonLineSearchMaxIteration, s.tau, (double)s.nSteps, (double)s.outputFreq,
                  (double)s.potentialUpdateFreq, (double)s.verbosity};
  std::copy(v, v + 16, sys);
  for (int i = 0; i < s.n_surfaces; i++) {
    const Surface& q = s.surfaces[i]; double* f = surf + 9 * i;
    f[0] = q.coulombBtype; f[1] = q.coulombFlux; f[2] = q.coulombPotential;
    f[3] = q.plusDiffusionBtype; f[4] = q.plusDiffusionFlux; f[5] = q.plusDiffusionConcentration;
    f[6] = q.minusDiffusionBtype; f[7] = q.minusDiffusionFlux; f[8] = q.minusDiffusionConcentration;
  }
  if (meshfile && meshfile_cap > 0) { std::strncpy(meshfile, s.meshfile.c_str(), meshfile_cap - 1); meshfile[meshfile_cap - 1] = 0; }
}
void ora_params_free(void* h) { delete (Params*)h; }

int ora_dirichlet(void* mh, void* ph, int fields, int comp0, char* out) {
  ORA_TRY
  Space sp = make_space(*(Mesh*)mh, ((Params*)ph)->s, fields, comp0);
  std::copy(sp.dirichlet.begin(), sp.dirichlet.end(), out);
  return 0;
  ORA_CATCH(-1)
}
// pattern: call with col == nullptr to get nnz
long ora_pattern(void* mh, void* ph, int fields, int comp0, int* rowptr, int* col) {
  ORA_TRY
  Space sp = make_space(*(Mesh*)mh, ((Params*)ph)->s, fields, comp0);
  CSR A = make_pattern(sp);
  if (rowptr) std::copy(A.rowptr.begin(), A.rowptr.end(), rowptr);
  if (col) std::copy(A.col.begin(), A.col.end(), col);
  return (long)A.col.size();
  ORA_CATCH(-1)
}
// the literal (std::set per row) restatement of the pattern rule: cross-check of make_pattern() in the tests
long ora_pattern_sets(void* mh, void* ph, int fields, int comp0, int* rowptr, int* col) {
  ORA_TRY
  Space sp = make_space(*(Mesh*)mh, ((Params*)ph)->s, fields, comp0);
  CSR A = make_pattern_sets(sp);
  if (rowptr) std::copy(A.rowptr.begin(), A.rowptr.end(), rowptr);
  if (col) std::copy(A.col.begin(), A.col.end(), col);
  return (long)A.col.size();
  ORA_CATCH(-1)
}
int ora_residual(void* mh, void* ph, int op, int comp0, const double* u, const double* aux0, const double* aux1,
                 double valency, int intorder, double* r, double* absr) {
  ORA_TRY
  const Mesh* m = (Mesh*)mh; const Sysparams* s = &((Params*)ph)->s;
  Space sp = make_space(*m, *s, op_fields(op), comp0);
  OpCtx c = make_ctx(m, s, op, aux0, aux1, valency, intorder);
  residual(sp, c, u, r, absr);
  return 0;
  ORA_CATCH(-1)
}
// values are written in the pattern order of ora_pattern()
int ora_jacobian(void* mh, void* ph, int op, int comp0, const double* u, const double* aux0, const double* aux1,
                 double valency, int intorder, int mode, double eps, double* val, double* absval) {
  ORA_TRY
  const Mesh* m = (Mesh*)mh; const Sysparams* s = &((Params*)ph)->s;
  Space sp = make_space(*m, *s, op_fields(op), comp0);
  OpCtx c = make_ctx(m, s, op, aux0, aux1, valency, intorder);
  CSR A = make_pattern(sp);
  std::vector<double> ab;
  jacobian(sp, c, u, A, mode, eps, absval ? &ab : nullptr);
  std::copy(A.val.begin(), A.val.end(), val);
  if (absval) std::copy(ab.begin(), ab.end(), absval);
  return 0;
  ORA_CATCH(-1)
}
int ora_interpolate(void* mh, void* ph, int comp, const double* pb, double* u) {
  ORA_TRY interpolate_bcext(*(Mesh*)mh, ((Params*)ph)->s, comp, pb, u); return 0; ORA_CATCH(-1)
}
// result[8] = {converged, iterations, reduction, conv_rate, status, seconds, 0, 0}
int ora_linsolve(int n, const int* rowptr, const int* col, const double* val, double* x, double* b,
                 double reduction, int maxit, int solver, int prec, int steps, double* result) {
  ORA_TRY
  CSR A; A.n = n; A.rowptr.assign(rowptr, rowptr + n + 1); A.col.assign(col, col + rowptr[n]);
  A.val.assign(val, val + rowptr[n]);
  auto t0 = std::chrono::steady_clock::now();
  LinResult lr = lin_solve(solver, A, x, b, reduction, maxit, prec, steps);
  double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  result[0] = lr.converged; result[1] = lr.iterations; result[2] = lr.reduction; result[3] = lr.conv_rate;
  result[4] = lr.status; result[5] = sec;
  return 0;
  ORA_CATCH(-1)
}
// v = M^-1 d for one of the ISTL preconditioners (PREC_*), applied once
int ora_prec_apply(int n, const int* rowptr, const int* col, const double* val, int prec, int steps, const double* d,
                   double* v) {
  ORA_TRY
  CSR A; A.n = n; A.rowptr.assign(rowptr, rowptr + n + 1); A.col.assign(col, col + rowptr[n]);
  A.val.assign(val, val + rowptr[n]);
  prec_apply(A, prec, steps, v, d);
  return 0;
  ORA_CATCH(-1)
}
void ora_spmv(int n, const int* rowptr, const int* col, const double* val, const double* x, double* y) {
  for (int r = 0; r < n; r++) {
    double s = 0.0;
    for (int k = rowptr[r]; k < rowptr[r + 1]; k++) s += val[k] * x[col[k]];
    y[r] = s;
  }
}
// opts[16] = {reduction, abs_limit, min_linear_reduction, reassemble_threshold, maxit, ls_maxit, damping,
//             jac_mode, fd_eps, solver, prec, prec_steps, lin_maxit, verbosity, line_search_strategy}
// result[16] = {status, converged, iterations, first_defect, defect, reduction, total_linear_iterations,
//               total_ls_trials, jacobian_assemblies, residual_assemblies, seconds}
// hist (optional, cap entries): defect history; lin_hist: linear iterations per step
int ora_newton(void* mh, void* ph, int op, int comp0, double* u, const double* aux0, const double* aux1,
               double valency, int intorder, const double* opts, double* result, double* hist, int* lin_hist, int cap) {
  ORA_TRY
  const Mesh* m = (Mesh*)mh; const Sysparams* s = &((Params*)ph)->s;
  Space sp = make_space(*m, *s, op_fields(op), comp0);
  OpCtx c = make_ctx(m, s, op, aux0, aux1, valency, intorder);
  NewtonOpts o;
  o.reduction = opts[0]; o.abs_limit = opts[1]; o.min_linear_reduction = opts[2]; o.reassemble_threshold = opts[3];
  o.maxit = (int)opts[4]; o.ls_maxit = (int)opts[5]; o.damping = opts[6]; o.jac_mode = (int)opts[7]; o.fd_eps = opts[8];
  o.solver = (int)opts[9]; o.prec = (int)opts[10]; o.prec_steps = (int)opts[11]; o.lin_maxit = (int)opts[12];
  o.verbosity = (int)opts[13]; o.line_search = (int)opts[14];
  auto t0 = std::chrono::steady_clock::now();
  NewtonResult R = newton(sp, c, u, o);
  double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  double v[11] = {(double)R.status, (double)R.converged, (double)R.iterations, R.first_defect, R.defect, R.reduction,
                  (double)R.total_linear_iterations, (double)R.total_ls_trials, (double)R.jacobian_assemblies,
                  (double)R.residual_assemblies, sec};
  std::copy(v, v + 11, result);
  if (hist) for (int i = 0; i < cap; i++) hist[i] = i < (int)R.defect_history.size() ? R.defect_history[i] : -1.0;
  if (lin_hist) for (int i = 0; i < cap; i++) lin_hist[i] = i < (int)R.lin_iter_history.size() ? R.lin_iter_history[i] : -1;
  return 0;
  ORA_CATCH(-1)
}
int ora_slp(void* mh, void* ph, int op, int comp0, double* u, const double* aux0, const double* aux1, double valency,
            int intorder, double reduction, int solver, int prec, int steps, int maxit, int jac_mode, double eps,
            double* result) {
  ORA_TRY
  const Mesh* m = (Mesh*)mh; const Sysparams* s = &((Params*)ph)->s;
  Space sp = make_space(*m, *s, op_fields(op), comp0);
  OpCtx c = make_ctx(m, s, op, aux0, aux1, valency, intorder);
  LinResult lr = slp_apply(sp, c, u, reduction, solver, prec, steps, maxit, jac_mode, eps);
  result[0] = lr.converged; result[1] = lr.iterations; result[2] = lr.reduction; result[3] = lr.conv_rate; result[4] = lr.status;
  return 0;
  ORA_CATCH(-1)
}

// OneStepMethod<Alexander2 (method 0) | ImplicitEuler (1)>::apply for the scalar transport problem:
// spatial operator DiffusionOperator(phi, valency), temporal operator DiffusionTOperator; result[2*stage + {0,1}] =
// {converged, iterations} of the stage solves
int ora_onestep(void* mh, void* ph, int comp0, int method, double dt, const double* xold, const double* g,
                const double* phi, double valency, double* xnew, double reduction, int solver, int prec, int steps,
                int maxit, int jac_mode, double eps, double* result) {
  ORA_TRY
  const Mesh* m = (Mesh*)mh; const Sysparams* s = &((Params*)ph)->s;
  Space sp = make_space(*m, *s, 1, comp0);
  OpCtx c0 = make_ctx(m, s, OP_DIFFUSION, phi, nullptr, valency, -1);
  OpCtx c1 = make_ctx(m, s, OP_MASS, nullptr, nullptr, 1.0, -1);
  TimeMethod tm = method == 1 ? implicit_euler() : alexander2();
  OneStepResult R = onestep_apply(sp, c0, c1, tm, dt, xold, g, xnew, reduction, solver, prec, steps, maxit, jac_mode, eps);
  for (size_t k = 0; k < R.stage.size(); k++) { result[2 * k] = R.stage[k].converged; result[2 * k + 1] = R.stage[k].iterations; }
  return 0;
  ORA_CATCH(-1)
}

int ora_ion_flux(void* mh, void* ph, const double* phi, const double* cp, const double* cm, double* ip, double* im) {
  ORA_TRY ion_flux(*(Mesh*)mh, ((Params*)ph)->s, phi, cp, cm, ip, im); return 0; ORA_CATCH(-1)
}
int ora_write_cell_data(void* mh, const double* u, const char* filename) {
  ORA_TRY write_cell_data(*(Mesh*)mh, u, filename); return 0; ORA_CATCH(-1)
}

int ora_write_vtk(void* mh, const char* name, int n, const double* const* fields, const char* const* names, int ascii) {
  ORA_TRY
  std::vector<const double*> f(fields, fields + n);
  std::vector<std::string> nm;
  for (int i = 0; i < n; i++) nm.push_back(names[i]);
  write_vtk(*(Mesh*)mh, name, f, nm, ascii != 0);
  return 0;
  ORA_CATCH(-1)
}

// ---- CPU baseline of bench.py: the phases of one PNP Newton step, timed on `threads` cores ----
struct BenchState {
  const Mesh* m; const Sysparams* s; Space sp; OpCtx c; CSR A; Incidence I; std::vector<double> scratch, r, z, y;
};
void ora_set_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n < 1 ? 1 : n);
#else
  (void)n;
#endif
}
int ora_max_threads() {
#ifdef _OPENMP
  return omp_get_num_procs();
#else
  return 1;
#endif
}
void* ora_bench_create(void* mh, void* ph, int op) {
  ORA_TRY
  auto* b = new BenchState;
  b->m = (Mesh*)mh; b->s = &((Params*)ph)->s;
  b->sp = make_space(*b->m, *b->s, op_fields(op), 0);
  b->c = make_ctx(b->m, b->s, op, nullptr, nullptr, 1.0, -1);
  b->A = make_pattern(b->sp);
  b->I = vertex_elements(*b->m);
  b->r.assign(b->sp.N(), 0.0); b->z.assign(b->sp.N(), 0.0); b->y.assign(b->sp.N(), 0.0);
  return b;
  ORA_CATCH(nullptr)
}
void ora_bench_free(void* h) { delete (BenchState*)h; }
long ora_bench_nnz(void* h) { return (long)((BenchState*)h)->A.col.size(); }
// One bounded sample of a Newton step at state u: 1 Jacobian assembly (jac_mode 0: the reference's FD jacobian_volume),
// 2 residual assemblies, `krylov_iters` iterations of BiCGSTAB + SSOR(1) (the reference's default backend) on A z = r,
// and `spmv_reps` extra SpMVs timed on their own.  out[8] = {t_jacobian, t_residual (one), t_krylov (all iterations),
// t_spmv (one, median), t_total (jacobian + 2 residuals + krylov), defect, reduction reached by the bounded solve, 0}
int ora_bench_step(void* h, const double* u, int threads, int jac_mode, int krylov_iters, int spmv_reps, double* out) {
  ORA_TRY
  BenchState& b = *(BenchState*)h;
  ora_set_threads(threads);
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto sec = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point e) { return std::chrono::duration<double>(e - a).count(); };
  const int N = b.sp.N();
  auto t0 = now();
  jacobian_par(b.sp, b.c, b.I, u, b.A, jac_mode, 1e-11, b.scratch);
  auto t1 = now();
  residual_par(b.sp, b.c, b.I, u, b.r.data(), b.scratch);
  auto t2 = now();
  residual_par(b.sp, b.c, b.I, u, b.y.data(), b.scratch); // the line search's trial residual
  auto t3 = now();
  const double defect = nrm2(N, b.r.data());
  std::fill(b.z.begin(), b.z.end(), 0.0);
  std::vector<double> rhs = b.r;
  auto t4 = now();
  LinResult lr = bicgstab(b.A, b.z.data(), rhs.data(), 1e-30, krylov_iters, PREC_SSOR, 1);
  auto t5 = now();
  std::vector<double> ts;
  for (int k = 0; k < spmv_reps; k++) { auto a0 = now(); b.A.mv(b.z.data(), b.y.data()); ts.push_back(sec(a0, now())); }
  std::sort(ts.begin(), ts.end());
  out[0] = sec(t0, t1); out[1] = 0.5 * sec(t1, t3); out[2] = sec(t4, t5); out[3] = ts.empty() ? 0.0 : ts[ts.size() / 2];
  out[4] = sec(t0, t3) + sec(t4, t5); out[5] = defect; out[6] = lr.reduction; out[7] = lr.iterations;
  ora_set_threads(1);
  return 0;
  ORA_CATCH(-1)
}
// bit-identity check of the multi-core assembly against the sequential one (tests)
int ora_assembly_par(void* mh, void* ph, int op, const double* u, int threads, int mode, double* r, double* val) {
  ORA_TRY
  const Mesh* m = (Mesh*)mh; const Sysparams* s = &((Params*)ph)->s;
  Space sp = make_space(*m, *s, op_fields(op), 0);
  OpCtx c = make_ctx(m, s, op, nullptr, nullptr, 1.0, -1);
  CSR A = make_pattern(sp);
  Incidence I = vertex_elements(*m);
  std::vector<double> scratch;
  ora_set_threads(threads);
  residual_par(sp, c, I, u, r, scratch);
  jacobian_par(sp, c, I, u, A, mode, 1e-11, scratch);
  ora_set_threads(1);
  std::copy(A.val.begin(), A.val.end(), val);
  return 0;
  ORA_CATCH(-1)
}

void ora_sinh_shared(int n, const double* x, double* y) { for (int i = 0; i < n; i++) y[i] = sinh_shared(x[i]); }

// ---- quadratic and cubic elements (PDEGREE = 2, 3): pnp_oracle_p2.hpp; `degree` picks Pk<2> or Pk<3> ----
static OpCtx make_ctx2(const Mesh* m, const Sysparams* s, int op, const double* a0, const double* a1, double valency, int intorder) {
  return make_ctx(m, s, op, a0, a1, valency, intorder); // (coefficient vectors are P2 vectors of length nE + nv here)
}
// sizes[4] = {nE, scalar dofs, first vertex dof, first edge dof}
extern "C++" template <class P> int pk_sizes(void* mh, void* ph, long* sizes) {
  ORA_TRY
  typename P::Space2 sp = P::make_space2(*(Mesh*)mh, ((Params*)ph)->s, 1, 0);
  sizes[0] = sp.nE; sizes[1] = sp.nd; sizes[2] = sp.voff; sizes[3] = sp.eoff;
  return 0;
  ORA_CATCH(-1)
}
int orak_sizes(int degree, void* mh, void* ph, long* sizes) {
  return degree == 3 ? pk_sizes<p3>(mh, ph, sizes) : pk_sizes<p2>(mh, ph, sizes);
}
// edges (end vertices), element -> local edge index, coordinates of the scalar dofs
extern "C++" template <class P> int pk_space(void* mh, void* ph, int* eva, int* evb, int* tedge, double* dx, double* dy) {
  ORA_TRY
  const Mesh& m = *(Mesh*)mh;
  typename P::Space2 sp = P::make_space2(m, ((Params*)ph)->s, 1, 0);
  if (eva) std::copy(sp.eva.begin(), sp.eva.end(), eva);
  if (evb) std::copy(sp.evb.begin(), sp.evb.end(), evb);
  if (tedge) std::copy(sp.tedge.begin(), sp.tedge.end(), tedge);
  if (dx && dy) // position of every scalar dof = its Lagrange node: geometry().global(node) from any element that holds it
    for (int e = 0; e < m.nT; e++) {
      const int a = m.tri[3 * e], b = m.tri[3 * e + 1], c = m.tri[3 * e + 2];
      for (int i = 0; i < P::NL; i++) {
        const double lx = P::nodes().x[i], ly = P::nodes().y[i];
        const int d = sp.sdof(e, i);
        if (P::nodes().kind[i] == 0) { dx[d] = m.x[m.tri[3 * e + P::nodes().sub[i]]]; dy[d] = m.y[m.tri[3 * e + P::nodes().sub[i]]]; continue; }
        dx[d] = m.x[a] + (m.x[b] - m.x[a]) * lx + (m.x[c] - m.x[a]) * ly;
        dy[d] = m.y[a] + (m.y[b] - m.y[a]) * lx + (m.y[c] - m.y[a]) * ly;
      }
    }
  return 0;
  ORA_CATCH(-1)
}
int orak_space(int degree, void* mh, void* ph, int* eva, int* evb, int* tedge, double* dx, double* dy) {
  return degree == 3 ? pk_space<p3>(mh, ph, eva, evb, tedge, dx, dy) : pk_space<p2>(mh, ph, eva, evb, tedge, dx, dy);
}
extern "C++" template <class P> int pk_dirichlet(void* mh, void* ph, int fields, int comp0, char* out) {
  ORA_TRY
  typename P::Space2 sp = P::make_space2(*(Mesh*)mh, ((Params*)ph)->s, fields, comp0);
  std::copy(sp.dirichlet.begin(), sp.dirichlet.end(), out);
  return 0;
  ORA_CATCH(-1)
}
int orak_dirichlet(int degree, void* mh, void* ph, int fields, int comp0, char* out) {
  return degree == 3 ? pk_dirichlet<p3>(mh, ph, fields, comp0, out) : pk_dirichlet<p2>(mh, ph, fields, comp0, out);
}
extern "C++" template <class P> long pk_pattern(void* mh, void* ph, int fields, int comp0, int* rowptr, int* col) {
  ORA_TRY
  typename P::Space2 sp = P::make_space2(*(Mesh*)mh, ((Params*)ph)->s, fields, comp0);
  CSR A = P::make_pattern2(sp);
  if (rowptr) std::copy(A.rowptr.begin(), A.rowptr.end(), rowptr);
  if (col) std::copy(A.col.begin(), A.col.end(), col);
  return (long)A.col.size();
  ORA_CATCH(-1)
}
long orak_pattern(int degree, void* mh, void* ph, int fields, int comp0, int* rowptr, int* col) {
  return degree == 3 ? pk_pattern<p3>(mh, ph, fields, comp0, rowptr, col) : pk_pattern<p2>(mh, ph, fields, comp0, rowptr, col);
}
extern "C++" template <class P> int pk_residual(void* mh, void* ph, int op, int comp0, const double* u, const double* aux0, const double* aux1, double valency,
                  int intorder, double* r, double* absr) {
  ORA_TRY
  const Mesh* m = (Mesh*)mh; const Sysparams* s = &((Params*)ph)->s;
  typename P::Space2 sp = P::make_space2(*m, *s, op_fields(op), comp0);
  OpCtx c = make_ctx2(m, s, op, aux0, aux1, valency, intorder);
  P::residual2(sp, c, u, r, absr);
  return 0;
  ORA_CATCH(-1)
}
int orak_residual(int degree, void* mh, void* ph, int op, int comp0, const double* u, const double* aux0, const double* aux1, double valency,
                  int intorder, double* r, double* absr) {
  return degree == 3 ? pk_residual<p3>(mh, ph, op, comp0, u, aux0, aux1, valency, intorder, r, absr) : pk_residual<p2>(mh, ph, op, comp0, u, aux0, aux1, valency, intorder, r, absr);
}
extern "C++" template <class P> int pk_jacobian(void* mh, void* ph, int op, int comp0, const double* u, const double* aux0, const double* aux1, double valency,
                  int intorder, int mode, double eps, double* val, double* absval) {
  ORA_TRY
  const Mesh* m = (Mesh*)mh; const Sysparams* s = &((Params*)ph)->s;
  typename P::Space2 sp = P::make_space2(*m, *s, op_fields(op), comp0);
  OpCtx c = make_ctx2(m, s, op, aux0, aux1, valency, intorder);
  CSR A = P::make_pattern2(sp);
  std::vector<double> ab;
  P::jacobian2(sp, c, u, A, mode, eps, absval ? &ab : nullptr);
  std::copy(A.val.begin(), A.val.end(), val);
  if (absval) std::copy(ab.begin(), ab.end(), absval);
  return 0;
  ORA_CATCH(-1)
}
int orak_jacobian(int degree, void* mh, void* ph, int op, int comp0, const double* u, const double* aux0, const double* aux1, double valency,
                  int intorder, int mode, double eps, double* val, double* absval) {
  return degree == 3 ? pk_jacobian<p3>(mh, ph, op, comp0, u, aux0, aux1, valency, intorder, mode, eps, val, absval) : pk_jacobian<p2>(mh, ph, op, comp0, u, aux0, aux1, valency, intorder, mode, eps, val, absval);
}
extern "C++" template <class P> int pk_interpolate(void* mh, void* ph, int comp, const double* pb, double* u) {
  ORA_TRY
  typename P::Space2 sp = P::make_space2(*(Mesh*)mh, ((Params*)ph)->s, 1, comp);
  P::interpolate_bcext2(sp, comp, pb, u);
  return 0;
  ORA_CATCH(-1)
}
int orak_interpolate(int degree, void* mh, void* ph, int comp, const double* pb, double* u) {
  return degree == 3 ? pk_interpolate<p3>(mh, ph, comp, pb, u) : pk_interpolate<p2>(mh, ph, comp, pb, u);
}
extern "C++" template <class P> int pk_newton(void* mh, void* ph, int op, int comp0, double* u, const double* aux0, const double* aux1, double valency,
                int intorder, const double* opts, double* result, double* hist, int* lin_hist, int cap) {
  ORA_TRY
  const Mesh* m = (Mesh*)mh; const Sysparams* s = &((Params*)ph)->s;
  typename P::Space2 sp = P::make_space2(*m, *s, op_fields(op), comp0);
  OpCtx c = make_ctx2(m, s, op, aux0, aux1, valency, intorder);
  NewtonOpts o;
  o.reduction = opts[0]; o.abs_limit = opts[1]; o.min_linear_reduction = opts[2]; o.reassemble_threshold = opts[3];
  o.maxit = (int)opts[4]; o.ls_maxit = (int)opts[5]; o.damping = opts[6]; o.jac_mode = (int)opts[7]; o.fd_eps = opts[8];
  o.solver = (int)opts[9]; o.prec = (int)opts[10]; o.prec_steps = (int)opts[11]; o.lin_maxit = (int)opts[12];
  o.verbosity = (int)opts[13]; o.line_search = (int)opts[14];
  auto t0 = std::chrono::steady_clock::now();
  NewtonResult R = newton_core(sp.N(), P::make_pattern2(sp), [&](const double* uu, double* r) { P::residual2(sp, c, uu, r); },
                               [&](const double* uu, CSR& A) { P::jacobian2(sp, c, uu, A, o.jac_mode, o.fd_eps); }, u, o);
  double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  double v[11] = {(double)R.status, (double)R.converged, (double)R.iterations, R.first_defect, R.defect, R.reduction,
                  (double)R.total_linear_iterations, (double)R.total_ls_trials, (double)R.jacobian_assemblies,
                  (double)R.residual_assemblies, sec};
  std::copy(v, v + 11, result);
  if (hist) for (int i = 0; i < cap; i++) hist[i] = i < (int)R.defect_history.size() ? R.defect_history[i] : -1.0;
  if (lin_hist) for (int i = 0; i < cap; i++) lin_hist[i] = i < (int)R.lin_iter_history.size() ? R.lin_iter_history[i] : -1;
  return 0;
  ORA_CATCH(-1)
}
int orak_newton(int degree, void* mh, void* ph, int op, int comp0, double* u, const double* aux0, const double* aux1, double valency,
                int intorder, const double* opts, double* result, double* hist, int* lin_hist, int cap) {
  return degree == 3 ? pk_newton<p3>(mh, ph, op, comp0, u, aux0, aux1, valency, intorder, opts, result, hist, lin_hist, cap) : pk_newton<p2>(mh, ph, op, comp0, u, aux0, aux1, valency, intorder, opts, result, hist, lin_hist, cap);
}
extern "C++" template <class P> int pk_onestep(void* mh, void* ph, int comp0, int method, double dt, const double* xold, const double* g,
                                               const double* phi, double valency, int intorder, double* xnew, double reduction, int solver,
                                               int prec, int steps, int maxit, int jac_mode, double eps, double* result) {
  ORA_TRY
  const Mesh* m = (Mesh*)mh; const Sysparams* s = &((Params*)ph)->s;
  typename P::Space2 sp = P::make_space2(*m, *s, 1, comp0);
  OpCtx c0 = make_ctx2(m, s, OP_DIFFUSION, phi, nullptr, valency, intorder);
  OpCtx c1 = make_ctx2(m, s, OP_MASS, nullptr, nullptr, 1.0, intorder);
  TimeMethod tm = method == 1 ? implicit_euler() : alexander2();
  OneStepResult R = P::onestep2(sp, c0, c1, tm, dt, xold, g, xnew, reduction, solver, prec, steps, maxit, jac_mode, eps);
  for (size_t k = 0; k < R.stage.size(); k++) { result[2 * k] = R.stage[k].converged; result[2 * k + 1] = R.stage[k].iterations; }
  return 0;
  ORA_CATCH(-1)
}
int orak_onestep(int degree, void* mh, void* ph, int comp0, int method, double dt, const double* xold, const double* g, const double* phi,
                 double valency, int intorder, double* xnew, double reduction, int solver, int prec, int steps, int maxit, int jac_mode,
                 double eps, double* result) {
  return degree == 3 ? pk_onestep<p3>(mh, ph, comp0, method, dt, xold, g, phi, valency, intorder, xnew, reduction, solver, prec, steps, maxit, jac_mode, eps, result)
                     : pk_onestep<p2>(mh, ph, comp0, method, dt, xold, g, phi, valency, intorder, xnew, reduction, solver, prec, steps, maxit, jac_mode, eps, result);
}
extern "C++" template <class P> int pk_ion_flux(void* mh, void* ph, const double* phi, const double* cp, const double* cm, double* ip, double* im) {
  ORA_TRY
  typename P::Space2 sp = P::make_space2(*(Mesh*)mh, ((Params*)ph)->s, 1, 0);
  P::ion_flux2(sp, phi, cp, cm, ip, im);
  return 0;
  ORA_CATCH(-1)
}
int orak_ion_flux(int degree, void* mh, void* ph, const double* phi, const double* cp, const double* cm, double* ip, double* im) {
  return degree == 3 ? pk_ion_flux<p3>(mh, ph, phi, cp, cm, ip, im) : pk_ion_flux<p2>(mh, ph, phi, cp, cm, ip, im);
}
extern "C++" template <class P> int pk_write_cell_data(void* mh, void* ph, const double* u, const char* filename) {
  ORA_TRY
  typename P::Space2 sp = P::make_space2(*(Mesh*)mh, ((Params*)ph)->s, 1, 0);
  P::write_cell_data2(sp, u, filename);
  return 0;
  ORA_CATCH(-1)
}
int orak_write_cell_data(int degree, void* mh, void* ph, const double* u, const char* filename) {
  return degree == 3 ? pk_write_cell_data<p3>(mh, ph, u, filename) : pk_write_cell_data<p2>(mh, ph, u, filename);
}
extern "C++" template <class P> void pk_basis(double x, double y, double* phi, double* grad) {
  P::basis(x, y, phi);
  double g[P::NL][2]; P::basis_grad(x, y, g);
  for (int i = 0; i < P::NL; i++) { grad[2 * i] = g[i][0]; grad[2 * i + 1] = g[i][1]; }
}
void orak_basis(int degree, double x, double y, double* phi, double* grad) {
  if (degree == 3) pk_basis<p3>(x, y, phi, grad); else pk_basis<p2>(x, y, phi, grad);
}

} // extern "C"
