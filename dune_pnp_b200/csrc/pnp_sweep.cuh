// pnp_sweep.cuh -- per-dof logic of the sequential-order preconditioners (SSOR, ILU0) as __host__ __device__ functions.
//
// ISTL's SeqSSOR / SeqILU0 (wrapped by ISTLBackend_NOVLP_BCGS_SSORk, /root/reference/src/instationary_pnp_from_pb_md.hh:188-191;
// SURVEY App. A.7-A.8) sweep the dofs in the order of the matrix rows.  The result of such a sweep only depends on the
// relative order of COUPLED dofs, so it is reproduced exactly by "level scheduling": level(i) = 1 + max level of the coupled
// dofs that precede i in the reference's order; dofs of one level are mutually uncoupled and are updated in parallel, levels
// run in sequence (forward: ascending, backward: descending).  The order is the reference's, not the internal one:
// dof (field f, vertex v) sits at f*nv + ext(v) (GridFunctionSpaceLexicographicMapper, stationary_pnp.hh:126-129).
//
// pnp_precond.cu wraps these functions in kernels; tests/host_harness runs them in host loops against the CPU oracle.
#pragma once
#include "pnp_setup_algos.cuh"

namespace pnp {

struct SweepView {
  const int* rp;        // n_own + 1
  const unsigned* adj;  // nslots (column vertex in the low 27 bits)
  const int* int2ext;   // nv: reference index of an internal vertex (null: the row order itself is the sweep order)
  int n_own;            // rows; columns >= n_own are ghosts and are skipped (the preconditioner acts on the local block)
};

// plane of the 7-plane PNP matrix that holds block (f, g); -1: the block is an exact zero and is not stored
PNP_HD int pnp_plane7(int f, int g) {
  return f == 0 ? g : (g == 0 ? (f == 1 ? 3 : 5) : (f == g ? (f == 1 ? 4 : 6) : -1));
}

// One relaxation of the level recurrence for dof (f, v).  full = all F x F field blocks couple (PDELab's pattern, what
// ILU0 fills); otherwise the (c+, c-) / (c-, c+) blocks are left out (they are exact zeros in the assembled matrix).
template <int F>
PNP_HD int sweep_level_relax(const SweepView& S, const int* lev, int v, int f, bool full) {
  int best = -1;
  const int ev = S.int2ext ? S.int2ext[v] : v;
  for (int s = S.rp[v]; s < S.rp[v + 1]; s++) {
    const int w = (int)(S.adj[s] & STAR_VMASK);
    if (w >= S.n_own) continue;
    const bool before = (S.int2ext ? S.int2ext[w] : w) < ev;
#pragma unroll
    for (int g = 0; g < F; g++) {
      if (g > f || (g == f && !before)) continue;
      if (g < f && g != 0 && !full) continue;
      const int l = lev[F * w + g];
      best = l > best ? l : best;
    }
  }
  return best + 1;
}

// SeqSSOR / bsorf-bsorb item (w = 1):  x_i += (d_i - sum_k A_ik x_k) / A_ii   over the stored planes (NP = 1 or 7)
template <int NP>
PNP_HD void gs_update(const SweepView& S, const double* vals, long stride, const double* d, double* x, int v, int f) {
  constexpr int F = NP == 1 ? 1 : 3;
  const int s0 = S.rp[v], s1 = S.rp[v + 1];
  double sum = d[(long)F * v + f];
  for (int s = s0; s < s1; s++) {
    const long w = (long)(S.adj[s] & STAR_VMASK);
    if (w >= S.n_own) continue;
    if (NP == 1) sum -= vals[s] * x[w];
    else if (f == 0) sum -= vals[s] * x[3 * w] + vals[stride + s] * x[3 * w + 1] + vals[2 * stride + s] * x[3 * w + 2];
    else sum -= vals[(2 * f + 1) * stride + s] * x[3 * w] + vals[(2 * f + 2) * stride + s] * x[3 * w + f];
  }
  const double diag = vals[(NP == 1 ? 0 : (f == 0 ? 0 : 2 * f + 2)) * stride + s0];
  if (diag != 0.0) x[(long)F * v + f] += sum / diag; // (a multigrid level may hold an empty row: an all-constrained aggregate)
}

// slot of column vertex x in row v, or -1
PNP_HD int sweep_row_find(const SweepView& S, int v, int x) {
  for (int s = S.rp[v]; s < S.rp[v + 1]; s++)
    if ((int)(S.adj[s] & STAR_VMASK) == x) return s;
  return -1;
}

// ILU(0) of row i = (f, v), dune-istl ilu.hh bilu0_decomposition: for every lower entry (i, j) in ascending order
// a_ij *= a_jj^-1 (the diagonal of a finished row is stored inverted), then a_ik -= a_ij a_jk for every k > j present in
// both rows; finally the diagonal of row i is inverted.  lu: F*F planes (block (f, g) in plane F*f + g).
template <int F>
PNP_HD void ilu0_row(const SweepView& S, double* lu, long stride, int v, int f) {
  const int s0 = S.rp[v], s1 = S.rp[v + 1];
  const int ev = S.int2ext[v];
  for (int g = 0; g <= f; g++) {
    int last = -1;
    for (;;) {
      int sb = -1, eb = 0x7fffffff;
      for (int s = s0; s < s1; s++) {
        const int w = (int)(S.adj[s] & STAR_VMASK);
        if (w >= S.n_own) continue;
        const int e = S.int2ext[w];
        if (e > last && e < eb && (g < f || e < ev)) { sb = s; eb = e; }
      }
      if (sb < 0) break;
      last = eb;
      const int w = (int)(S.adj[sb] & STAR_VMASK);
      const double l = lu[(long)(F * f + g) * stride + sb] * lu[(long)(F * g + g) * stride + S.rp[w]];
      lu[(long)(F * f + g) * stride + sb] = l;
      for (int t = S.rp[w]; t < S.rp[w + 1]; t++) {
        const int x = (int)(S.adj[t] & STAR_VMASK);
        if (x >= S.n_own) continue;
        const int sx = x == v ? s0 : sweep_row_find(S, v, x);
        if (sx < 0) continue;
        const bool after = S.int2ext[x] > eb;
        for (int h = g; h < F; h++) {
          if (h == g && !after) continue;
          lu[(long)(F * f + h) * stride + sx] -= l * lu[(long)(F * g + h) * stride + t];
        }
      }
    }
  }
  lu[(long)(F * f + f) * stride + s0] = 1.0 / lu[(long)(F * f + f) * stride + s0];
}

// bilu_backsolve, lower part:  x_i = d_i - sum_{j<i} L_ij x_j
template <int F>
PNP_HD void ilu0_forward(const SweepView& S, const double* lu, long stride, const double* d, double* x, int v, int f) {
  const int ev = S.int2ext[v];
  double sum = d[(long)F * v + f];
  for (int s = S.rp[v]; s < S.rp[v + 1]; s++) {
    const long w = (long)(S.adj[s] & STAR_VMASK);
    if (w >= S.n_own) continue;
    const bool before = S.int2ext[w] < ev;
    for (int g = 0; g <= f; g++) {
      if (g == f && !before) continue;
      sum -= lu[(long)(F * f + g) * stride + s] * x[F * w + g];
    }
  }
  x[(long)F * v + f] = sum;
}
// upper part:  x_i = a_ii^-1 (x_i - sum_{j>i} U_ij x_j)
template <int F>
PNP_HD void ilu0_backward(const SweepView& S, const double* lu, long stride, double* x, int v, int f) {
  const int ev = S.int2ext[v], s0 = S.rp[v];
  double sum = x[(long)F * v + f];
  for (int s = s0; s < S.rp[v + 1]; s++) {
    const long w = (long)(S.adj[s] & STAR_VMASK);
    if (w >= S.n_own) continue;
    const bool after = S.int2ext[w] > ev;
    for (int g = f; g < F; g++) {
      if (g == f && !after) continue;
      sum -= lu[(long)(F * f + g) * stride + s] * x[F * w + g];
    }
  }
  x[(long)F * v + f] = sum * lu[(long)(F * f + f) * stride + s0];
}

} // namespace pnp
