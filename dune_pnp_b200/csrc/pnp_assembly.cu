// pnp_assembly.cu -- residual and Jacobian assembly kernels (fp64, vertex-parallel gather).
//
// Replaces GridOperator::residual / GridOperator::jacobian driving the reference's local operators
// (/root/reference/src/stationary_pnp.hh:240-246; instationary_pnp_from_pb_md.hh:183-186,343-366):
// one thread owns one vertex row, walks the vertex star, recomputes the incident element integrals
// for its own test function and writes each result exactly once.  HBM-bound: per vertex it streams
// the ring (4 B/slot), gathers coordinates and coefficients (L1/L2 hits after the locality
// renumbering) and writes F residual entries or NPLANES*rowlen matrix values.
//
// This source is compiled TWICE (Makefile): with -DPNP_ASM_FAITHFUL -fmad=false it provides the FD-faithful Jacobian
// kernels, which must round like the CPU restatement (no FMA contraction, the reference's operation order); without
// the macro (FMA allowed) it provides the residual and exact-derivative Jacobian kernels plus the host-side dispatch.
#include "pnp_common.cuh"

namespace pnp {

// implemented in the other compilation of this file
void launch_jacobian_fd(Ctx& c, const StarView& M, const Operator& op, const double* u, const double* a0, const double* a1,
                        double* vals, long stride, double eps);

namespace {

// One block assembles the rows of 128 consecutive vertices per pass.  Their slots form ONE contiguous range of every
// value plane, so the rows are first written into a shared-memory tile (a thread's row is 7 doubles per plane: stored
// straight to global memory every store instruction of a warp touches 32 different sectors, and L2 evicts them half
// written -- the first version of this kernel read 15.7 GB and wrote 27.2 GB for an 18.4 GB matrix) and then copied out
// with fully coalesced stores.  Chunks with more than JAC_CAP slots (average valence > 8) are written directly.
constexpr int JAC_CHUNK = 128;   // = block size
constexpr int JAC_CAP = 1152;    // staged slots per plane

template <int OP, int MODE>
__global__ void __launch_bounds__(JAC_CHUNK)
k_jacobian(StarView M, PhysParams P, const double* __restrict__ u, const double* __restrict__ aux0,
           const double* __restrict__ aux1, double eps, int comp0, double* __restrict__ vals, long stride) {
  constexpr int NP = OpTraits<OP>::NPLANES;
  extern __shared__ double stage[]; // NP * JAC_CAP
  const int nchunks = (M.nv + JAC_CHUNK - 1) / JAC_CHUNK;
  for (int ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
    const int v0 = ch * JAC_CHUNK, v1 = min(v0 + JAC_CHUNK, M.nv);
    const int sb = M.rp[v0], n = M.rp[v1] - sb;
    const int v = v0 + threadIdx.x;
    if (n <= JAC_CAP) { // block-uniform
      if (v < v1) jacobian_row<OP, MODE>(M, P, u, aux0, aux1, eps, comp0, v, stage, JAC_CAP, sb);
      __syncthreads();
#pragma unroll
      for (int p = 0; p < NP; p++)
        for (int i = threadIdx.x; i < n; i += JAC_CHUNK) vals[p * stride + sb + i] = stage[p * JAC_CAP + i];
      __syncthreads();
    } else if (v < v1) {
      jacobian_row<OP, MODE>(M, P, u, aux0, aux1, eps, comp0, v, vals, stride);
    }
  }
}

// M: the star to assemble on -- the context's own (finest) one, or a coarser refinement level's (multigrid)
template <int OP, int MODE>
void launch_jac(Ctx& c, const StarView& M, const Operator& op, const double* u, const double* a0, const double* a1,
                double* vals, long stride, double eps) {
  const PhysParams P = c.phys(op.valency);
  const int smem = OpTraits<OP>::NPLANES * JAC_CAP * (int)sizeof(double);
  // (set on every launch: the attribute is per device and a process may drive several contexts)
  PNP_CUDA(cudaFuncSetAttribute(k_jacobian<OP, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int grid = grid_for(M.nv, JAC_CHUNK, c.sm_count * 16);
  k_jacobian<OP, MODE><<<grid, JAC_CHUNK, smem, c.stream>>>(M, P, u, a0, a1, eps, op.comp0, vals, stride);
  PNP_CHECK_LAUNCH(); c.launches++;
  // ring (4 B/slot), row pointer + coordinates + Dirichlet mask + state (+ coefficient fields) per vertex, NP planes written
  c.acct(Ctx::ACC_ASSEMBLY, (4.0 + 8.0 * OpTraits<OP>::NPLANES) * (double)stride + (21.0 + 8.0 * OpTraits<OP>::F + 8.0 * OpTraits<OP>::NAUX) * (double)M.nv);
}
template <int MODE>
void dispatch_jac(Ctx& c, const StarView& M, const Operator& op, const double* u, const double* a0, const double* a1,
                  double* vals, long stride, double eps) {
  switch (op.op) {
    case OP_PB: launch_jac<OP_PB, MODE>(c, M, op, u, a0, a1, vals, stride, eps); break;
    case OP_POISSON: launch_jac<OP_POISSON, MODE>(c, M, op, u, a0, a1, vals, stride, eps); break;
    case OP_DIFFUSION: launch_jac<OP_DIFFUSION, MODE>(c, M, op, u, a0, a1, vals, stride, eps); break;
    case OP_MASS: launch_jac<OP_MASS, MODE>(c, M, op, u, a0, a1, vals, stride, eps); break;
    case OP_PNP: launch_jac<OP_PNP, MODE>(c, M, op, u, a0, a1, vals, stride, eps); break;
    default: PNP_REQUIRE(false, PNP_E_ARG, "unknown operator");
  }
}

} // namespace

#ifdef PNP_ASM_FAITHFUL

void launch_jacobian_fd(Ctx& c, const StarView& M, const Operator& op, const double* u, const double* a0, const double* a1,
                        double* vals, long stride, double eps) {
  dispatch_jac<JAC_FD_FAITHFUL>(c, M, op, u, a0, a1, vals, stride, eps);
}

#else

namespace {

template <int OP>
__global__ void __launch_bounds__(128)
k_residual(StarView M, PhysParams P, const double* __restrict__ u, const double* __restrict__ aux0,
           const double* __restrict__ aux1, int comp0, double* __restrict__ r) {
  constexpr int F = OpTraits<OP>::F;
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < M.nv; v += gridDim.x * blockDim.x) {
    double out[F];
    residual_row<OP, false>(M, P, u, aux0, aux1, v, out);
    const unsigned db = dir_bits<OP>(M, v, comp0);
#pragma unroll
    for (int k = 0; k < F; k++) r[(long)F * v + k] = ((db >> k) & 1u) ? 0.0 : out[k]; // constrain_residual
  }
}

// alpha_boundary: one thread per vertex that occurs in a boundary face (boundary_vertex_sum, pnp_star.cuh)
__global__ void k_boundary(StarView M, PhysParams P, const BFace* __restrict__ faces, const int* __restrict__ bv,
                           const int* __restrict__ ptr, const int* __restrict__ items, int n_bv,
                           const double* __restrict__ surf_flux, const unsigned char* __restrict__ surf_dir, int F,
                           int comp0, double* __restrict__ r) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_bv) return;
  double acc[3];
  boundary_vertex_sum(M, P, faces, items, ptr[t], ptr[t + 1], surf_flux, surf_dir, F, comp0, acc);
  const int v = bv[t];
  const unsigned m = M.dmask[v];
  for (int k = 0; k < F; k++) {
    const int comp = F == 3 ? k : comp0;
    if (!((m >> comp) & 1u)) r[(long)F * v + k] += acc[k];
  }
}

} // namespace
static void coefficient_ptrs(Ctx& c, const Operator& op, const double** a0, const double** a1) {
  *a0 = *a1 = nullptr;
  const int need = op.op == OP_POISSON ? 2 : (op.op == OP_DIFFUSION ? 1 : 0);
  if (need >= 1) {
    PNP_REQUIRE(op.aux0 >= 0, PNP_E_ARG, "operator coefficient 0 not set");
    PNP_REQUIRE(c.vec(op.aux0).fields == 1, PNP_E_ARG, "coefficient must be a 1-field vector");
    *a0 = c.vec(op.aux0).d.p;
  }
  if (need >= 2) {
    PNP_REQUIRE(op.aux1 >= 0, PNP_E_ARG, "operator coefficient 1 not set");
    PNP_REQUIRE(c.vec(op.aux1).fields == 1, PNP_E_ARG, "coefficient must be a 1-field vector");
    *a1 = c.vec(op.aux1).d.p;
  }
  // partitioned mesh: the coefficient fields are read at ghost vertices too, and solver results only cover owned dofs
  if (c.n_own < c.nv) {
    if (*a0) halo_exchange(c, const_cast<double*>(*a0), 1);
    if (*a1) halo_exchange(c, const_cast<double*>(*a1), 1);
  }
}

void assemble_residual(Ctx& c, const Operator& op, Vec& u, Vec& r) {
  PNP_REQUIRE(c.constraints_built, PNP_E_ARG, "constraints not built (mesh finalized + parameters set?)");
  const int F = op_fields(op.op);
  PNP_REQUIRE(u.fields == F && r.fields == F, PNP_E_ARG, "vector field count does not match the operator");
  const double *a0, *a1;
  coefficient_ptrs(c, op, &a0, &a1);
  const StarView M = c.star();
  const PhysParams P = c.phys(op.valency);
  const int block = 128, grid = grid_for(c.n_own, block, c.sm_count * 16);
  halo_exchange(c, u.d.p, F); // ghost values of the state (no-op on one GPU)
  switch (op.op) {
    case OP_PB: k_residual<OP_PB><<<grid, block, 0, c.stream>>>(M, P, u.d.p, a0, a1, op.comp0, r.d.p); break;
    case OP_POISSON: k_residual<OP_POISSON><<<grid, block, 0, c.stream>>>(M, P, u.d.p, a0, a1, op.comp0, r.d.p); break;
    case OP_DIFFUSION: k_residual<OP_DIFFUSION><<<grid, block, 0, c.stream>>>(M, P, u.d.p, a0, a1, op.comp0, r.d.p); break;
    case OP_MASS: k_residual<OP_MASS><<<grid, block, 0, c.stream>>>(M, P, u.d.p, a0, a1, op.comp0, r.d.p); break;
    case OP_PNP: k_residual<OP_PNP><<<grid, block, 0, c.stream>>>(M, P, u.d.p, a0, a1, op.comp0, r.d.p); break;
    default: PNP_REQUIRE(false, PNP_E_ARG, "unknown operator");
  }
  PNP_CHECK_LAUNCH(); c.launches++;
  c.acct(Ctx::ACC_ASSEMBLY, 4.0 * (double)c.nslots + (21.0 + 16.0 * F + 8.0 * (op.op == OP_POISSON ? 2 : (op.op == OP_DIFFUSION ? 1 : 0))) * (double)c.n_own);
  // doAlphaBoundary is false for the diffusion and mass operators (diffusion_operator.hh:34)
  if ((op.op == OP_PB || op.op == OP_POISSON || op.op == OP_PNP) && c.n_bv > 0) {
    // Poisson is constructed with the PB BCType (component 0): instationary_pnp_from_pb_md.hh:343-344
    k_boundary<<<(c.n_bv + 127) / 128, 128, 0, c.stream>>>(M, P, c.d_bfaces.p, c.d_bv.p, c.d_bv_ptr.p, c.d_bv_items.p,
                                                          c.n_bv, c.d_surf.p, c.d_surf_dir.p, F, op.comp0, r.d.p);
    PNP_CHECK_LAUNCH(); c.launches++;
  }
}

void assemble_jacobian(Ctx& c, const Operator& op, Vec& u, Matrix& A, int mode, double eps) {
  PNP_REQUIRE(c.constraints_built, PNP_E_ARG, "constraints not built (mesh finalized + parameters set?)");
  PNP_REQUIRE(u.fields == op_fields(op.op), PNP_E_ARG, "vector field count does not match the operator");
  PNP_REQUIRE(A.op == op.op, PNP_E_ARG, "matrix belongs to another operator type");
  PNP_REQUIRE(mode == JAC_FD_FAITHFUL || mode == JAC_ANALYTIC, PNP_E_ARG, "unknown jacobian mode");
  const double *a0, *a1;
  coefficient_ptrs(c, op, &a0, &a1);
  A.comp0 = op.comp0;
  c.last_u = u.d.p; c.last_op = op; c.last_mode = mode; c.last_eps = eps; c.last_vals = A.vals.p;
  halo_exchange(c, u.d.p, u.fields);
  if (mode == JAC_FD_FAITHFUL) launch_jacobian_fd(c, c.star(), op, u.d.p, a0, a1, A.vals.p, c.nslots, eps);
  else dispatch_jac<JAC_ANALYTIC>(c, c.star(), op, u.d.p, a0, a1, A.vals.p, c.nslots, eps);
}

// The same operator on another star of the same mesh hierarchy (a coarser refinement level held by the multigrid):
// u in that level's numbering, vals = NPLANES planes of `stride` slots.  Operators with coefficient fields are not
// supported here (their coefficients live on the finest level only).
void assemble_jacobian_on(Ctx& c, const StarView& M, long stride, const Operator& op, const double* u, double* vals, int mode,
                          double eps) {
  PNP_REQUIRE(op.op == OP_PB || op.op == OP_PNP || op.op == OP_MASS, PNP_E_ARG, "no coefficient fields on coarse levels");
  if (mode == JAC_FD_FAITHFUL) launch_jacobian_fd(c, M, op, u, nullptr, nullptr, vals, stride, eps);
  else dispatch_jac<JAC_ANALYTIC>(c, M, op, u, nullptr, nullptr, vals, stride, eps);
}

#endif // PNP_ASM_FAITHFUL

} // namespace pnp
