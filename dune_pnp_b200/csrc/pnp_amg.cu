// pnp_amg.cu -- aggregation AMG preconditioner on the device (PNP_PREC_AMG).
//
// Stands in for Dune::Amg::AMG as selected by ISTLBackend_NOVLP_CG_AMG_SSOR
// (/root/reference/src/instationary_pnp_from_pb_md.hh:208-211): unsmoothed
// aggregation, Galerkin coarse operators, V-cycle, `prec_steps` smoothing sweeps before and after the
// coarse correction, coarse correction scaled by dune-istl's default prolongation damping 1.6.
// What differs, because the sequential pieces of ISTL have no parallel equivalent: aggregates come
// from a parallel maximal independent set (root + its strongest-coupled neighbours) instead of
// ISTL's sequential front-growing, and the smoother is damped Jacobi instead of SSOR.  Linear
// iteration counts therefore differ from ISTL's; Newton counts do not.
//
// Point-block hierarchy: all fields of a vertex are aggregated together, so every level keeps the
// 1-plane (scalar) or 7-plane (PNP) layout of the fine matrix and runs the same SpMV.
// Symbolic phase (aggregates, coarse patterns, gather lists) once per mesh/operator type;
// numeric phase (Galerkin sums, inverse diagonals) per Jacobian -- deterministic gathers, no atomics.
#include <cub/cub.cuh>

#include "pnp_common.cuh"

namespace pnp {

namespace {

constexpr int BLK = 256;

struct Level {
  int nv = 0; long nslots = 0;
  // matrix (level 0 borrows the context's star arrays and the caller's values)
  const int* rp = nullptr; const unsigned* col = nullptr; const double* vals = nullptr;
  DBuf<int> rp_own; DBuf<unsigned> col_own; DBuf<double> vals_own;
  DBuf<double> dinv, x, x2, b, r;
  // transfer to the next coarser level
  DBuf<int> agg, agg_ptr, agg_mem;   // vertex -> aggregate; members of each aggregate
  DBuf<int> seg_ptr, seg_items;      // coarse slot -> fine slots summed into it
};

} // namespace

struct Amg {
  int NP = 1, F = 1;
  bool symbolic = false;
  long nv0 = -1, nslots0 = -1;
  double omega = 0.7, alpha = 1.6;
  int coarse_sweeps = 40;
  int comp0 = 0;
  std::vector<std::unique_ptr<Level>> L;
  DBuf<unsigned char> tmp;
  void* temp(size_t bytes) { if (bytes > tmp.n) tmp.alloc(bytes); return tmp.p; }
};

namespace {

__device__ __forceinline__ unsigned hash32(unsigned x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
// Luby rounds. state: 0 undecided, 1 root, 2 covered
__global__ void k_mis_select(const int* __restrict__ rp, const unsigned* __restrict__ col, int nv,
                             const unsigned char* __restrict__ st, unsigned char* __restrict__ st_out) {
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < nv; v += gridDim.x * blockDim.x) {
    unsigned char s = st[v];
    if (s == 0) {
      const unsigned pv = hash32((unsigned)v);
      bool win = true;
      for (int k = rp[v] + 1; k < rp[v + 1]; k++) {
        const int w = (int)(col[k] & STAR_VMASK);
        if (w >= nv || w == v || st[w] != 0) continue; // (ghost columns take no part in the local hierarchy)
        const unsigned pw = hash32((unsigned)w);
        if (pw > pv || (pw == pv && w > v)) { win = false; break; }
      }
      if (win) s = 1;
    }
    st_out[v] = s;
  }
}
__global__ void k_mis_cover(const int* __restrict__ rp, const unsigned* __restrict__ col, int nv,
                            unsigned char* __restrict__ st, int* __restrict__ undecided) {
  int local = 0;
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < nv; v += gridDim.x * blockDim.x) {
    if (st[v] != 0) continue;
    bool covered = false;
    for (int k = rp[v] + 1; k < rp[v + 1]; k++) {
      const int w = (int)(col[k] & STAR_VMASK);
      if (w < nv && st[w] == 1) { covered = true; break; }
    }
    if (covered) st[v] = 2; else local++;
  }
  if (local) atomicAdd(undecided, local);
}
__global__ void k_root_flag(const unsigned char* __restrict__ st, int nv, int* __restrict__ flag) {
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < nv; v += gridDim.x * blockDim.x) flag[v] = st[v] == 1;
}
// every covered vertex joins the adjacent root it is most strongly coupled to (|plane 0|)
__global__ void k_attach(const int* __restrict__ rp, const unsigned* __restrict__ col, const double* __restrict__ vals,
                         int nv, const unsigned char* __restrict__ st, const int* __restrict__ root_id,
                         int* __restrict__ agg) {
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < nv; v += gridDim.x * blockDim.x) {
    if (st[v] == 1) { agg[v] = root_id[v]; continue; }
    double best = -1.0; int bw = -1;
    for (int k = rp[v] + 1; k < rp[v + 1]; k++) {
      const int w = (int)(col[k] & STAR_VMASK);
      if (w >= nv || st[w] != 1) continue;
      const double s = fabs(vals[k]);
      if (s > best || (s == best && w < bw)) { best = s; bw = w; }
    }
    agg[v] = root_id[bw];
  }
}
// coarse-slot key of every fine slot: (I << 32) | (J == I ? 0 : J + 1)  -> the diagonal sorts first
__global__ void k_coarse_keys(const int* __restrict__ rp, const unsigned* __restrict__ col, int nv,
                              const int* __restrict__ agg, uint64_t* __restrict__ keys, int* __restrict__ slot_id) {
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < nv; v += gridDim.x * blockDim.x) {
    const unsigned I = (unsigned)agg[v];
    for (int k = rp[v]; k < rp[v + 1]; k++) {
      const unsigned w = col[k] & STAR_VMASK;
      if (w >= (unsigned)nv) { keys[k] = ~0ull; slot_id[k] = k; continue; } // ghost column: dropped (block-Jacobi over ranks)
      const unsigned J = (unsigned)agg[w];
      keys[k] = ((uint64_t)I << 32) | (J == I ? 0u : J + 1u);
      slot_id[k] = k;
    }
  }
}
__global__ void k_coarse_cols(const uint64_t* __restrict__ ukeys, long n, unsigned* __restrict__ col) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const unsigned lo = (unsigned)(ukeys[i] & 0xffffffffu);
    col[i] = lo == 0 ? (unsigned)(ukeys[i] >> 32) : lo - 1u;
  }
}
__global__ void k_row_starts(const uint64_t* __restrict__ ukeys, long n, int nc, int* __restrict__ rp) {
  for (int I = blockIdx.x * blockDim.x + threadIdx.x; I <= nc; I += gridDim.x * blockDim.x)
    rp[I] = (int)lower_bound_u64(ukeys, n, (uint64_t)(unsigned)I << 32);
}
__global__ void k_lower_bounds_i32(const int* __restrict__ sorted, int n, int m, int* __restrict__ out) {
  for (int I = blockIdx.x * blockDim.x + threadIdx.x; I <= m; I += gridDim.x * blockDim.x) {
    int lo = 0, hi = n;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (sorted[mid] < I) lo = mid + 1; else hi = mid; }
    out[I] = lo;
  }
}
__global__ void k_iota(int* a, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) a[i] = i;
}

// Galerkin sums A_c = P^T A P for piecewise-constant P: coarse slot cs = sum of its fine slots.
// At level 0 Dirichlet dofs are left out of the hierarchy (their rows/columns are already decoupled;
// only the unit diagonal has to be skipped): dmask/rp0 non-null, F = 3 uses bits 0..2, F = 1 uses bit comp0.
template <int NP>
__global__ void k_galerkin(const int* __restrict__ seg_ptr, const int* __restrict__ seg_items, long nsc,
                           const double* __restrict__ vf, long nsf, double* __restrict__ vc,
                           const unsigned char* __restrict__ dmask, const unsigned* __restrict__ colf, const int* __restrict__ rpf,
                           int comp0) {
  for (long cs = blockIdx.x * (long)blockDim.x + threadIdx.x; cs < nsc; cs += (long)gridDim.x * blockDim.x) {
    double acc[NP];
#pragma unroll
    for (int p = 0; p < NP; p++) acc[p] = 0.0;
    for (int t = seg_ptr[cs]; t < seg_ptr[cs + 1]; t++) {
      const int s = seg_items[t];
      double v[NP];
#pragma unroll
      for (int p = 0; p < NP; p++) v[p] = vf[p * nsf + s];
      if (dmask) { // level 0: skip the unit diagonal of constrained dofs (slot s is a diagonal slot iff s == rp[col[s]])
        const unsigned w = colf[s] & STAR_VMASK;
        const unsigned m = dmask[w];
        if (m && rpf[w] == s) {
          if (NP == 1) { if ((m >> comp0) & 1u) v[0] = 0.0; }
          else {
            if (m & 1u) v[0] = 0.0;
            if (m & 2u) v[NP == 7 ? 4 : 0] = 0.0;
            if (m & 4u) v[NP == 7 ? 6 : 0] = 0.0;
          }
        }
      }
#pragma unroll
      for (int p = 0; p < NP; p++) acc[p] += v[p];
    }
#pragma unroll
    for (int p = 0; p < NP; p++) vc[p * nsc + cs] = acc[p];
  }
}
template <int NP>
__global__ void k_dinv(const int* __restrict__ rp, const double* __restrict__ vals, long stride, int nv,
                       double* __restrict__ dinv) {
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < nv; v += gridDim.x * blockDim.x) {
    const int s = rp[v];
    if (NP == 1) { const double d = vals[s]; dinv[v] = d != 0.0 ? 1.0 / d : 0.0; }
    else {
      const double d0 = vals[s], d1 = vals[4 * stride + s], d2 = vals[6 * stride + s];
      dinv[3l * v] = d0 != 0.0 ? 1.0 / d0 : 0.0;
      dinv[3l * v + 1] = d1 != 0.0 ? 1.0 / d1 : 0.0;
      dinv[3l * v + 2] = d2 != 0.0 ? 1.0 / d2 : 0.0;
    }
  }
}

// y = A x with an epilogue: EPI 0: y = A x; 1: y = b - A x; 2: y = x + omega*dinv*(b - A x)
template <int NP, int EPI, int LANES>
__global__ void __launch_bounds__(BLK)
k_level_op(const int* __restrict__ rp, const unsigned* __restrict__ col, const double* __restrict__ vals, long stride,
           const double* __restrict__ x, const double* __restrict__ b, const double* __restrict__ dinv, double omega,
           double* __restrict__ y, int nv) {
  constexpr int F = NP == 1 ? 1 : 3;
  constexpr int RPW = 32 / LANES;
  const int lane = threadIdx.x & (LANES - 1);
  const int grp = (blockIdx.x * blockDim.x + threadIdx.x) / LANES;
  const int ngrp = gridDim.x * blockDim.x / LANES;
  for (int base = grp - (grp % RPW); base < nv; base += ngrp) {
    const int row = base + (grp % RPW);
    double acc[F];
#pragma unroll
    for (int k = 0; k < F; k++) acc[k] = 0.0;
    if (row < nv) {
      for (int s = rp[row] + lane; s < rp[row + 1]; s += LANES) {
        const long c = col[s] & STAR_VMASK;
        if (NP == 1) acc[0] += vals[s] * x[c];
        else {
          const double x0 = x[3 * c], x1 = x[3 * c + 1], x2 = x[3 * c + 2];
          acc[0] += vals[s] * x0 + vals[stride + s] * x1 + vals[2 * stride + s] * x2;
          acc[1] += vals[3 * stride + s] * x0 + vals[4 * stride + s] * x1;
          acc[2] += vals[5 * stride + s] * x0 + vals[6 * stride + s] * x2;
        }
      }
    }
#pragma unroll
    for (int k = 0; k < F; k++)
#pragma unroll
      for (int o = LANES / 2; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
    if (lane == 0 && row < nv) {
#pragma unroll
      for (int k = 0; k < F; k++) {
        const long i = (long)F * row + k;
        if (EPI == 0) y[i] = acc[k];
        else if (EPI == 1) y[i] = b[i] - acc[k];
        else y[i] = x[i] + omega * dinv[i] * (b[i] - acc[k]);
      }
    }
  }
}
// first sweep from a zero initial guess: x = omega * dinv * b
__global__ void k_jacobi0(const double* __restrict__ dinv, const double* __restrict__ b, double omega,
                          double* __restrict__ x, long n) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    x[i] = omega * dinv[i] * b[i];
}
template <int F>
__global__ void k_restrict(const int* __restrict__ agg_ptr, const int* __restrict__ agg_mem, int nc,
                           const double* __restrict__ r, double* __restrict__ bc) {
  for (int I = blockIdx.x * blockDim.x + threadIdx.x; I < nc; I += gridDim.x * blockDim.x) {
    double acc[F];
#pragma unroll
    for (int k = 0; k < F; k++) acc[k] = 0.0;
    for (int t = agg_ptr[I]; t < agg_ptr[I + 1]; t++) {
      const long v = agg_mem[t];
#pragma unroll
      for (int k = 0; k < F; k++) acc[k] += r[F * v + k];
    }
#pragma unroll
    for (int k = 0; k < F; k++) bc[(long)F * I + k] = acc[k];
  }
}
// x += alpha * P x_c ; constrained dofs (level 0 only) receive no correction
template <int F>
__global__ void k_prolong(const int* __restrict__ agg, int nv, const double* __restrict__ xc, double alpha,
                          double* __restrict__ x, const unsigned char* __restrict__ dmask, int comp0) {
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < nv; v += gridDim.x * blockDim.x) {
    const long I = agg[v];
    const unsigned m = dmask ? (F == 3 ? dmask[v] : (dmask[v] >> comp0) & 1u) : 0u;
#pragma unroll
    for (int k = 0; k < F; k++)
      if (!((m >> k) & 1u)) x[(long)F * v + k] += alpha * xc[F * I + k];
  }
}

#define KL(c, kern, n, ...)                                                                              \
  do { kern<<<grid_for((n), BLK), BLK, 0, (c).stream>>>(__VA_ARGS__); PNP_CHECK_LAUNCH(); (c).launches++; } while (0)

template <int EPI>
void level_op(Ctx& c, const Amg& A, const Level& l, const double* x, const double* b, double* y) {
  constexpr int LANES = 8;
  const int grid = grid_for((long)l.nv * LANES, BLK, c.sm_count * 8);
  const bool fine = &l == A.L[0].get();
  if (fine) c.prof_mark();
  if (A.NP == 1)
    k_level_op<1, EPI, LANES><<<grid, BLK, 0, c.stream>>>(l.rp, l.col, l.vals, l.nslots, x, b, l.dinv.p, A.omega, y, l.nv);
  else
    k_level_op<7, EPI, LANES><<<grid, BLK, 0, c.stream>>>(l.rp, l.col, l.vals, l.nslots, x, b, l.dinv.p, A.omega, y, l.nv);
  PNP_CHECK_LAUNCH(); c.launches++;
  if (fine) c.prof_mark();
}

// builds level li+1 from level li (symbolic); returns false if coarsening stalled
bool coarsen(Ctx& c, Amg& A, int li) {
  Level& f = *A.L[li];
  const int nv = f.nv;
  DBuf<unsigned char> st(nv), st2(nv);
  st.zero(c.stream);
  DBuf<int> d_und(1);
  for (int round = 0; round < 200; round++) {
    KL(c, k_mis_select, nv, f.rp, f.col, nv, st.p, st2.p);
    std::swap(st.p, st2.p);
    d_und.zero(c.stream);
    KL(c, k_mis_cover, nv, f.rp, f.col, nv, st.p, d_und.p);
    int und = 0;
    d_und.download(&und, 1, c.stream);
    if (und == 0) break;
    PNP_REQUIRE(round < 199, PNP_E_ARG, "AMG: independent-set selection did not terminate");
  }
  DBuf<int> flag(nv), root_id(nv);
  KL(c, k_root_flag, nv, st.p, nv, flag.p);
  size_t bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, bytes, flag.p, root_id.p, nv, c.stream);
  PNP_CUDA(cub::DeviceScan::ExclusiveSum(A.temp(bytes), bytes, flag.p, root_id.p, nv, c.stream));
  int last_id = 0, last_flag = 0;
  PNP_CUDA(cudaMemcpyAsync(&last_id, root_id.p + nv - 1, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
  PNP_CUDA(cudaMemcpyAsync(&last_flag, flag.p + nv - 1, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
  PNP_CUDA(cudaStreamSynchronize(c.stream));
  const int nc = last_id + last_flag;
  if (nc < 1 || nc > 0.8 * nv) return false;
  f.agg.alloc(nv);
  KL(c, k_attach, nv, f.rp, f.col, f.vals, nv, st.p, root_id.p, f.agg.p);
  // members of each aggregate
  {
    DBuf<int> ids(nv), agg_sorted(nv);
    f.agg_mem.alloc(nv); f.agg_ptr.alloc(nc + 1);
    KL(c, k_iota, nv, ids.p, nv);
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, f.agg.p, agg_sorted.p, ids.p, f.agg_mem.p, nv, 0, 32, c.stream);
    PNP_CUDA(cub::DeviceRadixSort::SortPairs(A.temp(bytes), bytes, f.agg.p, agg_sorted.p, ids.p, f.agg_mem.p, nv, 0, 32,
                                            c.stream));
    KL(c, k_lower_bounds_i32, nc + 1, agg_sorted.p, nv, nc, f.agg_ptr.p);
  }
  // coarse pattern + gather lists
  auto nl = std::make_unique<Level>();
  {
    const long ns = f.nslots;
    PNP_REQUIRE(ns < (1l << 31), PNP_E_MESH, "AMG: too many matrix slots on one level");
    DBuf<uint64_t> keys(ns), skeys(ns), ukeys(ns);
    DBuf<int> slot(ns), counts(ns + 1), d_nu(1);
    f.seg_items.alloc(ns);
    KL(c, k_coarse_keys, nv, f.rp, f.col, nv, f.agg.p, keys.p, slot.p);
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, keys.p, skeys.p, slot.p, f.seg_items.p, (int)ns, 0, 64, c.stream);
    PNP_CUDA(cub::DeviceRadixSort::SortPairs(A.temp(bytes), bytes, keys.p, skeys.p, slot.p, f.seg_items.p, (int)ns, 0, 64,
                                            c.stream));
    cub::DeviceRunLengthEncode::Encode(nullptr, bytes, skeys.p, ukeys.p, counts.p, d_nu.p, (int)ns, c.stream);
    PNP_CUDA(cub::DeviceRunLengthEncode::Encode(A.temp(bytes), bytes, skeys.p, ukeys.p, counts.p, d_nu.p, (int)ns, c.stream));
    int nu = 0;
    d_nu.download(&nu, 1, c.stream);
    { // the run of dropped ghost-column slots (key ~0) sorts last: leave it out
      uint64_t lastkey = 0;
      PNP_CUDA(cudaMemcpyAsync(&lastkey, ukeys.p + (nu - 1), sizeof(uint64_t), cudaMemcpyDeviceToHost, c.stream));
      PNP_CUDA(cudaStreamSynchronize(c.stream));
      if (lastkey == ~0ull) nu--;
    }
    nl->nv = nc; nl->nslots = nu;
    PNP_CUDA(cudaMemsetAsync(counts.p + nu, 0, sizeof(int), c.stream));
    f.seg_ptr.alloc((size_t)nu + 1);
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, counts.p, f.seg_ptr.p, nu + 1, c.stream);
    PNP_CUDA(cub::DeviceScan::ExclusiveSum(A.temp(bytes), bytes, counts.p, f.seg_ptr.p, nu + 1, c.stream));
    nl->rp_own.alloc((size_t)nc + 1); nl->col_own.alloc(nu);
    KL(c, k_row_starts, nc + 1, ukeys.p, (long)nu, nc, nl->rp_own.p);
    KL(c, k_coarse_cols, nu, ukeys.p, (long)nu, nl->col_own.p);
    c.launches += 8;
    PNP_CUDA(cudaStreamSynchronize(c.stream));
  }
  nl->vals_own.alloc((size_t)A.NP * nl->nslots);
  nl->rp = nl->rp_own.p; nl->col = nl->col_own.p; nl->vals = nl->vals_own.p;
  A.L.push_back(std::move(nl));
  return true;
}

void alloc_work(Amg& A, Level& l) {
  const size_t n = (size_t)A.F * l.nv;
  l.dinv.alloc(n); l.x.alloc(n); l.x2.alloc(n); l.b.alloc(n); l.r.alloc(n);
}

void numeric(Ctx& c, Amg& A, int comp0) {
  for (size_t li = 0; li < A.L.size(); li++) {
    Level& l = *A.L[li];
    if (li + 1 < A.L.size()) {
      Level& n = *A.L[li + 1];
      const unsigned char* dm = li == 0 ? c.dmask.p : nullptr;
      if (A.NP == 1)
        KL(c, k_galerkin<1>, n.nslots, l.seg_ptr.p, l.seg_items.p, n.nslots, l.vals, l.nslots, n.vals_own.p, dm, l.col, l.rp, comp0);
      else
        KL(c, k_galerkin<7>, n.nslots, l.seg_ptr.p, l.seg_items.p, n.nslots, l.vals, l.nslots, n.vals_own.p, dm, l.col, l.rp, comp0);
    }
    if (A.NP == 1) KL(c, k_dinv<1>, l.nv, l.rp, l.vals, l.nslots, l.nv, l.dinv.p);
    else KL(c, k_dinv<7>, l.nv, l.rp, l.vals, l.nslots, l.nv, l.dinv.p);
  }
}

// one V-cycle on level li: solves A x = b approximately, x from a zero initial guess; result in l.x
void vcycle(Ctx& c, Amg& A, int li, int nu, int comp0) {
  Level& l = *A.L[li];
  const long n = (long)A.F * l.nv;
  const bool coarsest = li + 1 == (int)A.L.size();
  const int pre = coarsest ? A.coarse_sweeps : nu;
  KL(c, k_jacobi0, n, l.dinv.p, l.b.p, A.omega, l.x.p, n);
  for (int s = 1; s < pre; s++) { level_op<2>(c, A, l, l.x.p, l.b.p, l.x2.p); std::swap(l.x.p, l.x2.p); }
  if (coarsest) return;
  Level& nx = *A.L[li + 1];
  level_op<1>(c, A, l, l.x.p, l.b.p, l.r.p);
  if (A.F == 1) KL(c, k_restrict<1>, nx.nv, l.agg_ptr.p, l.agg_mem.p, nx.nv, l.r.p, nx.b.p);
  else KL(c, k_restrict<3>, nx.nv, l.agg_ptr.p, l.agg_mem.p, nx.nv, l.r.p, nx.b.p);
  vcycle(c, A, li + 1, nu, comp0);
  const unsigned char* dm = li == 0 ? c.dmask.p : nullptr;
  if (A.F == 1) KL(c, k_prolong<1>, l.nv, l.agg.p, l.nv, nx.x.p, A.alpha, l.x.p, dm, comp0);
  else KL(c, k_prolong<3>, l.nv, l.agg.p, l.nv, nx.x.p, A.alpha, l.x.p, dm, comp0);
  for (int s = 0; s < nu; s++) { level_op<2>(c, A, l, l.x.p, l.b.p, l.x2.p); std::swap(l.x.p, l.x2.p); }
}

} // namespace

void amg_setup(Ctx& c, Solver& S, const Matrix& M) {
  if (!S.amg) S.amg = std::make_shared<Amg>();
  Amg& A = *S.amg;
  const int comp0 = M.comp0;
  A.comp0 = comp0;
  if (!A.symbolic || A.NP != M.nplanes || A.nv0 != c.n_own || A.nslots0 != c.nslots) {
    A.L.clear();
    A.NP = M.nplanes; A.F = M.nplanes == 1 ? 1 : 3;
    A.nv0 = c.n_own; A.nslots0 = c.nslots;
    auto l0 = std::make_unique<Level>();
    l0->nv = (int)c.n_own; l0->nslots = c.nslots; l0->rp = c.rp.p; l0->col = c.adj.p; l0->vals = M.vals.p;
    A.L.push_back(std::move(l0));
    // the hierarchy's strength of connection is read from the current matrix values, level by level
    for (int li = 0; li < 24; li++) {
      if (A.L[li]->nv <= 64) break;
      if (!coarsen(c, A, li)) break;
      // numeric values of the new level are needed before it can be coarsened further
      Level& f = *A.L[li]; Level& n = *A.L[li + 1];
      const unsigned char* dm = li == 0 ? c.dmask.p : nullptr;
      if (A.NP == 1)
        KL(c, k_galerkin<1>, n.nslots, f.seg_ptr.p, f.seg_items.p, n.nslots, f.vals, f.nslots, n.vals_own.p, dm, f.col, f.rp, comp0);
      else
        KL(c, k_galerkin<7>, n.nslots, f.seg_ptr.p, f.seg_items.p, n.nslots, f.vals, f.nslots, n.vals_own.p, dm, f.col, f.rp, comp0);
    }
    for (auto& l : A.L) alloc_work(A, *l);
    { // level 0 iterates are SpMV inputs whose ghost columns must read as zero
      Level& l0r = *A.L[0];
      const size_t nall = (size_t)A.F * c.nv;
      l0r.x.alloc(nall); l0r.x2.alloc(nall);
      l0r.x.zero(c.stream); l0r.x2.zero(c.stream);
    }
    A.symbolic = true;
    if (S.verbosity > 0) {
      std::printf("AMG hierarchy:");
      for (auto& l : A.L) std::printf(" %d", l->nv);
      std::printf("\n");
    }
  }
  A.L[0]->vals = M.vals.p;
  numeric(c, A, comp0);
}

// y = M^{-1} d : `prec_steps` pre- and post-smoothing sweeps per level
void amg_apply(Ctx& c, Solver& S, const Matrix&, const double* d, double* y) {
  Amg& A = *S.amg;
  Level& l0 = *A.L[0];
  const long n = (long)A.F * l0.nv;
  PNP_CUDA(cudaMemcpyAsync(l0.b.p, d, n * sizeof(double), cudaMemcpyDeviceToDevice, c.stream));
  vcycle(c, A, 0, S.prec_steps > 0 ? S.prec_steps : 1, A.comp0);
  PNP_CUDA(cudaMemcpyAsync(y, l0.x.p, n * sizeof(double), cudaMemcpyDeviceToDevice, c.stream));
}

} // namespace pnp
