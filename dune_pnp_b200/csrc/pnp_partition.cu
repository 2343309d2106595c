// pnp_partition.cu -- native domain decomposition for the multi-GPU path (host code; no kernel in this file).
//
// Stands in for `grid->loadBalance()` (/root/reference/src/pnp_solver_main.cc:93-108) and PDELab's non-overlapping ghost
// bookkeeping (stationary_pnp.hh:131): one process per GPU, every rank runs the same deterministic steps on the same
// coarse (Gmsh) mesh and keeps only its own part:
//   1. recursive coordinate bisection of the triangle centroids into `world` equal-count parts;
//   2. rank p keeps its triangles plus every triangle touching one of their vertices (one ghost layer);
//   3. `levels` uniform red refinements of that local mesh with the library's rule (new vertex = nv + rank of the edge key;
//      children (a,ab,ac)(ab,b,bc)(ac,bc,c)(ab,bc,ac), boundary segments split in order), trimmed back to one ghost layer
//      after each -- no mesh larger than a rank's share is ever built, and shared vertices get bit-identical coordinates
//      on all ranks;
//   4. a vertex belongs to the lowest rank among the triangles around it; the halo plan comes from matching the bitwise
//      coordinates of each rank's ghost vertices against the interface vertices the other ranks own.  Step 4 needs two
//      small all-gathers (ghost keys, claims); the transport is the caller's (pnp_part_* hands the buffers over):
//      torch.distributed in the Python launchers, pnp_comm_allgatherv (NCCL) in the C++ drivers.
// dune_pnp_b200/partition.py is the numpy statement of the same steps; tests/test_partition_native.py holds the two against
// each other array by array.
#include <algorithm>
#include <cstring>
#include <numeric>
#include <unordered_map>

#include "pnp_common.cuh"

namespace pnp {
namespace {

struct Key2 { uint64_t a, b; bool operator==(const Key2& o) const { return a == o.a && b == o.b; } };
struct Key2Hash { size_t operator()(const Key2& k) const { return (size_t)(k.a * 0x9E3779B97F4A7C15ull ^ (k.b + 0x7F4A7C15ull + (k.a << 6) + (k.a >> 2))); } };
inline uint64_t dbits(double d) { uint64_t u; std::memcpy(&u, &d, sizeof u); return u; }

// a rank's piece of one mesh level in its own vertex numbering
struct LocalMesh {
  std::vector<double> x, y;
  std::vector<int> tri, tag;        // 3 per triangle; owner rank of each triangle
  std::vector<int> ba, bb, bphys;
  std::vector<long> par0, par1;     // parents in the numbering of the LocalMesh this one was refined from (-1: none)
  std::vector<long> gid;            // global vertex index (coarsest level only)
  long nv() const { return (long)x.size(); }
  long nT() const { return (long)tag.size(); }
};

// keeps the given triangles, the vertices they use (ascending) and the boundary segments whose two ends stay
LocalMesh compact(const LocalMesh& m, const std::vector<char>& keep) {
  const long nv = m.nv(), nT = m.nT();
  std::vector<char> used(nv, 0);
  for (long t = 0; t < nT; t++) if (keep[t]) for (int i = 0; i < 3; i++) used[m.tri[3 * t + i]] = 1;
  std::vector<int> nid(nv, -1);
  long n = 0;
  for (long v = 0; v < nv; v++) if (used[v]) nid[v] = (int)n++;
  LocalMesh r;
  r.x.reserve(n); r.y.reserve(n);
  for (long v = 0; v < nv; v++) if (used[v]) { r.x.push_back(m.x[v]); r.y.push_back(m.y[v]); }
  if (!m.par0.empty()) for (long v = 0; v < nv; v++) if (used[v]) { r.par0.push_back(m.par0[v]); r.par1.push_back(m.par1[v]); }
  if (!m.gid.empty()) for (long v = 0; v < nv; v++) if (used[v]) r.gid.push_back(m.gid[v]);
  long nk = 0;
  for (long t = 0; t < nT; t++) nk += keep[t] != 0;
  r.tri.reserve(3 * (size_t)nk); r.tag.reserve(nk);
  for (long t = 0; t < nT; t++) if (keep[t]) {
    for (int i = 0; i < 3; i++) r.tri.push_back(nid[m.tri[3 * t + i]]);
    r.tag.push_back(m.tag[t]);
  }
  for (size_t s = 0; s < m.ba.size(); s++)
    if (used[m.ba[s]] && used[m.bb[s]]) { r.ba.push_back(nid[m.ba[s]]); r.bb.push_back(nid[m.bb[s]]); r.bphys.push_back(m.bphys[s]); }
  return r;
}

// one ghost layer: my triangles + every triangle touching a vertex of one of my triangles
LocalMesh trim(const LocalMesh& m, int me) {
  const long nT = m.nT();
  std::vector<char> mine_v(m.nv(), 0), keep(nT, 0);
  for (long t = 0; t < nT; t++) if (m.tag[t] == me) for (int i = 0; i < 3; i++) mine_v[m.tri[3 * t + i]] = 1;
  for (long t = 0; t < nT; t++)
    keep[t] = m.tag[t] == me || mine_v[m.tri[3 * t]] || mine_v[m.tri[3 * t + 1]] || mine_v[m.tri[3 * t + 2]];
  return compact(m, keep);
}

// uniform red refinement with the library's rule.  Edges are numbered in the order of their keys (min vertex << 32 | max
// vertex); instead of sorting all 3 nT keys, the larger end points are bucketed by the smaller one (a counting sort over the
// vertices) and each short bucket is sorted and made unique: the same edge order in linear time with vertex-local accesses.
LocalMesh refine(const LocalMesh& m) {
  const long nv = m.nv(), nT = m.nT();
  static const int EV[3][2] = {{0, 1}, {0, 2}, {1, 2}};
  std::vector<long> ptr(nv + 1, 0);
  for (long t = 0; t < nT; t++)
    for (int e = 0; e < 3; e++) ptr[std::min(m.tri[3 * t + EV[e][0]], m.tri[3 * t + EV[e][1]]) + 1]++;
  for (long v = 0; v < nv; v++) ptr[v + 1] += ptr[v];
  std::vector<int> hi(ptr[nv]);
  {
    std::vector<long> fill(ptr.begin(), ptr.end() - 1);
    for (long t = 0; t < nT; t++)
      for (int e = 0; e < 3; e++) {
        const int a = m.tri[3 * t + EV[e][0]], b = m.tri[3 * t + EV[e][1]];
        hi[fill[std::min(a, b)]++] = std::max(a, b);
      }
  }
  std::vector<long> eptr(nv + 1, 0); // edges whose smaller end is v: [eptr[v], eptr[v+1]) of ehi, ascending larger end
  std::vector<int> ehi; ehi.reserve(hi.size() / 2 + 16);
  for (long v = 0; v < nv; v++) {
    std::sort(hi.begin() + ptr[v], hi.begin() + ptr[v + 1]);
    auto last = std::unique(hi.begin() + ptr[v], hi.begin() + ptr[v + 1]);
    ehi.insert(ehi.end(), hi.begin() + ptr[v], last);
    eptr[v + 1] = (long)ehi.size();
  }
  const long nE = (long)ehi.size();
  auto mid = [&](int a, int b) {
    const int lo = std::min(a, b), h = std::max(a, b);
    long k = eptr[lo];
    while (ehi[k] != h) k++; // (the edge exists; buckets hold ~3 entries)
    return (int)(nv + k);
  };
  LocalMesh r;
  r.x = m.x; r.y = m.y; r.x.resize(nv + nE); r.y.resize(nv + nE);
  r.par0.assign(nv + nE, -1); r.par1.assign(nv + nE, -1);
  for (long v = 0; v < nv; v++) {
    r.par0[v] = v;
    for (long k = eptr[v]; k < eptr[v + 1]; k++) {
      const long b = ehi[k];
      r.x[nv + k] = 0.5 * (m.x[v] + m.x[b]); r.y[nv + k] = 0.5 * (m.y[v] + m.y[b]);
      r.par0[nv + k] = v; r.par1[nv + k] = b;
    }
  }
  r.tri.resize(12 * (size_t)nT); r.tag.resize(4 * (size_t)nT);
  for (long t = 0; t < nT; t++) {
    const int a = m.tri[3 * t], b = m.tri[3 * t + 1], c = m.tri[3 * t + 2];
    const int ab = mid(a, b), ac = mid(a, c), bc = mid(b, c);
    const int ch[12] = {a, ab, ac, ab, b, bc, ac, bc, c, ab, bc, ac};
    std::copy(ch, ch + 12, r.tri.begin() + 12 * t);
    for (int q = 0; q < 4; q++) r.tag[4 * t + q] = m.tag[t];
  }
  r.ba.reserve(2 * m.ba.size()); r.bb.reserve(2 * m.ba.size()); r.bphys.reserve(2 * m.ba.size());
  for (size_t s = 0; s < m.ba.size(); s++) {
    const int mm = mid(m.ba[s], m.bb[s]);
    r.ba.push_back(m.ba[s]); r.bb.push_back(mm); r.bphys.push_back(m.bphys[s]);
    r.ba.push_back(mm); r.bb.push_back(m.bb[s]); r.bphys.push_back(m.bphys[s]);
  }
  return r;
}

// recursive coordinate bisection of the triangle centroids into equal-count parts (stable sorts: ties keep their order)
void rcb(const std::vector<double>& cx, const std::vector<double>& cy, std::vector<long>& idx, long lo_i, long hi_i, int lo, int n,
         std::vector<int>& part) {
  if (n == 1) { for (long i = lo_i; i < hi_i; i++) part[idx[i]] = lo; return; }
  const int nl = n / 2;
  double xmin = 1e300, xmax = -1e300, ymin = 1e300, ymax = -1e300;
  for (long i = lo_i; i < hi_i; i++) {
    xmin = std::min(xmin, cx[idx[i]]); xmax = std::max(xmax, cx[idx[i]]);
    ymin = std::min(ymin, cy[idx[i]]); ymax = std::max(ymax, cy[idx[i]]);
  }
  const std::vector<double>& k = (xmax - xmin) >= (ymax - ymin) ? cx : cy;
  std::stable_sort(idx.begin() + lo_i, idx.begin() + hi_i, [&](long a, long b) { return k[a] < k[b]; });
  const long cut = lo_i + ((hi_i - lo_i) * nl) / n;
  rcb(cx, cy, idx, lo_i, cut, lo, nl, part);
  rcb(cx, cy, idx, cut, hi_i, lo + nl, n - nl, part);
}

} // namespace
} // namespace pnp

using namespace pnp;

// ownership, ordering and halo plan of one level
struct PartLevel {
  LocalMesh lm;
  // finalize state
  std::vector<long> own_ids, ghost_ids;              // LocalMesh numbering, ascending
  std::vector<std::vector<long>> claim_pos, claim_id; // per rank q: positions in q's ghost list owned here, my ids in that order
  // the plan (this rank's numbering: owned first, ghosts grouped by owner rank)
  std::vector<long> order, old2new;
  long n_own = 0;
  std::vector<int> nbr, send_ptr, send_idx, recv_ptr;
  std::vector<int> tri, ba, bb;
  std::vector<double> x, y;
  std::vector<int> par0, par1;                        // parents in the next coarser PLAN's numbering (-1: none)
  std::vector<int> gid;
  bool finalized = false;
};
struct pnp_part {
  int world = 1, rank = 0, levels = 0;
  long n_global = 0;
  std::vector<PartLevel> L; // coarsest first
  std::string err;
};

#define PART_TRY try {
#define PART_CATCH(p)                                                                  \
    return PNP_OK;                                                                     \
  } catch (const pnp::Error& e) { if (p) (p)->err = e.what(); return e.code; }          \
  catch (const std::exception& e) { if (p) (p)->err = e.what(); return PNP_E_ARG; }

extern "C" {

pnp_status pnp_part_create(long nv, const double* x, const double* y, long nT, const int* tri, long nB, const int* ba, const int* bb,
                           const int* bphys, int world, int rank, int levels, pnp_part** out) {
  if (!out) return PNP_E_ARG;
  *out = nullptr;
  pnp_part* P = new pnp_part;
  PART_TRY
  PNP_REQUIRE(nv > 0 && nT > 0 && world >= 1 && rank >= 0 && rank < world && levels >= 0, PNP_E_ARG, "bad partition arguments");
  P->world = world; P->rank = rank; P->levels = levels; P->n_global = nv;
  std::vector<double> cx(nT), cy(nT);
  for (long t = 0; t < nT; t++) { // numpy's mean(1): ((a + b) + c) / 3
    cx[t] = ((x[tri[3 * t]] + x[tri[3 * t + 1]]) + x[tri[3 * t + 2]]) / 3.0;
    cy[t] = ((y[tri[3 * t]] + y[tri[3 * t + 1]]) + y[tri[3 * t + 2]]) / 3.0;
  }
  std::vector<long> idx(nT);
  std::iota(idx.begin(), idx.end(), 0l);
  std::vector<int> part(nT, 0);
  rcb(cx, cy, idx, 0, nT, 0, world, part);
  LocalMesh g;
  g.x.assign(x, x + nv); g.y.assign(y, y + nv); g.tri.assign(tri, tri + 3 * nT); g.tag = part;
  g.ba.assign(ba, ba + nB); g.bb.assign(bb, bb + nB); g.bphys.assign(bphys, bphys + nB);
  g.gid.resize(nv); std::iota(g.gid.begin(), g.gid.end(), 0l);
  LocalMesh lm = trim(g, rank);
  P->L.resize(levels + 1);
  for (int l = 0; l <= levels; l++) {
    if (l > 0) lm = trim(refine(lm), rank);
    P->L[l].lm = lm; // (the next level is refined from this LocalMesh: its par arrays index THIS numbering)
  }
  *out = P;
    return PNP_OK;
  } catch (const pnp::Error& e) { const int code = e.code; delete P; return code; }
  catch (const std::exception&) { delete P; return PNP_E_ARG; }
}

void pnp_part_destroy(pnp_part* P) { delete P; }
const char* pnp_part_last_error(pnp_part* P) { return P ? P->err.c_str() : "null partition"; }

// ownership of level l; hands out the coordinate keys (x bits, y bits) of this rank's ghost vertices
pnp_status pnp_part_ghost_keys(pnp_part* P, int level, long* n, unsigned long long* keys) {
  if (!P) return PNP_E_ARG;
  PART_TRY
  PNP_REQUIRE(level >= 0 && level <= P->levels && n, PNP_E_ARG, "bad level");
  PartLevel& L = P->L[level];
  const LocalMesh& m = L.lm;
  const long nv = m.nv(), nT = m.nT();
  if (L.own_ids.empty() && L.ghost_ids.empty()) {
    std::vector<char> mine_v(nv, 0);
    std::vector<int> owner(nv, 0x7fffffff);
    for (long t = 0; t < nT; t++)
      for (int i = 0; i < 3; i++) {
        const int v = m.tri[3 * t + i];
        owner[v] = std::min(owner[v], m.tag[t]);
        if (m.tag[t] == P->rank) mine_v[v] = 1;
      }
    for (long v = 0; v < nv; v++) (mine_v[v] && owner[v] == P->rank ? L.own_ids : L.ghost_ids).push_back(v);
  }
  *n = (long)L.ghost_ids.size();
  if (keys)
    for (size_t i = 0; i < L.ghost_ids.size(); i++) { keys[2 * i] = dbits(m.x[L.ghost_ids[i]]); keys[2 * i + 1] = dbits(m.y[L.ghost_ids[i]]); }
  PART_CATCH(P)
}

// all ranks' ghost keys in (ghost_ptr[q] .. ghost_ptr[q+1]) pairs -> for every rank q the positions of q's ghosts this rank
// owns: claim_ptr[world + 1], claim_pos[claim_ptr[world]] (call with claim_pos == NULL for the sizes)
pnp_status pnp_part_claim(pnp_part* P, int level, const long* ghost_ptr, const unsigned long long* ghost_keys_all, long* claim_ptr,
                          long* claim_pos) {
  if (!P) return PNP_E_ARG;
  PART_TRY
  PNP_REQUIRE(level >= 0 && level <= P->levels && ghost_ptr && claim_ptr, PNP_E_ARG, "bad arguments");
  PartLevel& L = P->L[level];
  const LocalMesh& m = L.lm;
  const long nv = m.nv(), nT = m.nT();
  const int me = P->rank, world = P->world;
  if (L.claim_pos.empty()) {
    PNP_REQUIRE(world == 1 || ghost_keys_all, PNP_E_ARG, "ghost keys missing");
    // owned vertices another rank may need: those within one layer of a triangle that is not mine
    std::vector<char> foreign_v(nv, 0), near_v(nv, 0), owned(nv, 0);
    for (long v : L.own_ids) owned[v] = 1;
    for (long t = 0; t < nT; t++) if (m.tag[t] != me) for (int i = 0; i < 3; i++) foreign_v[m.tri[3 * t + i]] = 1;
    for (long t = 0; t < nT; t++)
      if (foreign_v[m.tri[3 * t]] || foreign_v[m.tri[3 * t + 1]] || foreign_v[m.tri[3 * t + 2]])
        for (int i = 0; i < 3; i++) near_v[m.tri[3 * t + i]] = 1;
    std::unordered_map<Key2, long, Key2Hash> lookup;
    for (long v = 0; v < nv; v++) if (owned[v] && near_v[v]) lookup[Key2{dbits(m.x[v]), dbits(m.y[v])}] = v;
    L.claim_pos.assign(world, {}); L.claim_id.assign(world, {});
    for (int q = 0; q < world; q++) {
      if (q == me) continue;
      for (long i = ghost_ptr[q]; i < ghost_ptr[q + 1]; i++) {
        auto it = lookup.find(Key2{ghost_keys_all[2 * i], ghost_keys_all[2 * i + 1]});
        if (it != lookup.end()) { L.claim_pos[q].push_back(i - ghost_ptr[q]); L.claim_id[q].push_back(it->second); }
      }
    }
  }
  claim_ptr[0] = 0;
  for (int q = 0; q < world; q++) claim_ptr[q + 1] = claim_ptr[q] + (long)L.claim_pos[q].size();
  if (claim_pos)
    for (int q = 0; q < world; q++) std::copy(L.claim_pos[q].begin(), L.claim_pos[q].end(), claim_pos + claim_ptr[q]);
  PART_CATCH(P)
}

// claims about MY ghosts: mine_ptr[r] .. mine_ptr[r+1] = positions (in my ghost list) of the ghosts rank r owns.  Builds the
// level's plan; levels must be finalized coarsest first (the parents are translated into the coarser plan's numbering).
pnp_status pnp_part_finalize(pnp_part* P, int level, const long* mine_ptr, const long* mine_pos) {
  if (!P) return PNP_E_ARG;
  PART_TRY
  PNP_REQUIRE(level >= 0 && level <= P->levels, PNP_E_ARG, "bad level");
  PNP_REQUIRE(level == 0 || P->L[level - 1].finalized, PNP_E_ARG, "finalize the levels coarsest first");
  PartLevel& L = P->L[level];
  const LocalMesh& m = L.lm;
  const long nv = m.nv();
  const int me = P->rank, world = P->world;
  PNP_REQUIRE(world == 1 || (mine_ptr && !L.claim_pos.empty()), PNP_E_ARG, "claims missing (pnp_part_claim first)");
  L.order = L.own_ids;
  L.nbr.clear(); L.recv_ptr.assign(1, 0); L.send_ptr.assign(1, 0); L.send_idx.clear();
  std::vector<int> seen(L.ghost_ids.size(), 0);
  std::vector<int> nbrs;
  if (world > 1) {
    for (int r = 0; r < world; r++) {
      if (r == me) continue;
      const long n = mine_ptr[r + 1] - mine_ptr[r];
      if (n == 0 && L.claim_pos[r].empty()) continue;
      nbrs.push_back(r);
      for (long i = mine_ptr[r]; i < mine_ptr[r + 1]; i++) {
        PNP_REQUIRE(mine_pos[i] >= 0 && mine_pos[i] < (long)L.ghost_ids.size(), PNP_E_ARG, "claim position out of range");
        seen[mine_pos[i]]++;
        L.order.push_back(L.ghost_ids[mine_pos[i]]);
      }
      L.recv_ptr.push_back(L.recv_ptr.back() + (int)n);
    }
    for (int s : seen) PNP_REQUIRE(s == 1, PNP_E_MESH, "halo plan: a ghost vertex is unclaimed or claimed twice");
  } else PNP_REQUIRE(L.ghost_ids.empty(), PNP_E_MESH, "one rank cannot have ghosts");
  L.old2new.assign(nv, -1);
  for (size_t i = 0; i < L.order.size(); i++) L.old2new[L.order[i]] = (long)i;
  L.nbr = nbrs;
  for (int r : nbrs) {
    for (long id : L.claim_id[r]) L.send_idx.push_back((int)L.old2new[id]);
    L.send_ptr.push_back((int)L.send_idx.size());
  }
  L.n_own = (long)L.own_ids.size();
  L.x.resize(nv); L.y.resize(nv);
  for (long i = 0; i < nv; i++) { L.x[i] = m.x[L.order[i]]; L.y[i] = m.y[L.order[i]]; }
  L.tri.resize(m.tri.size());
  for (size_t i = 0; i < m.tri.size(); i++) L.tri[i] = (int)L.old2new[m.tri[i]];
  L.ba.resize(m.ba.size()); L.bb.resize(m.bb.size());
  for (size_t i = 0; i < m.ba.size(); i++) { L.ba[i] = (int)L.old2new[m.ba[i]]; L.bb[i] = (int)L.old2new[m.bb[i]]; }
  L.par0.assign(nv, -1); L.par1.assign(nv, -1);
  if (level > 0) {
    const PartLevel& C = P->L[level - 1];
    for (long i = 0; i < nv; i++) {
      const long p0 = m.par0[L.order[i]], p1 = m.par1[L.order[i]];
      L.par0[i] = p0 >= 0 ? (int)C.old2new[p0] : -1;
      L.par1[i] = p1 >= 0 ? (int)C.old2new[p1] : -1;
    }
  }
  L.gid.clear();
  if (!m.gid.empty()) { L.gid.resize(nv); for (long i = 0; i < nv; i++) L.gid[i] = (int)m.gid[L.order[i]]; }
  L.finalized = true;
  PART_CATCH(P)
}

// sizes[8] = {nv, n_own, nT, nB, n_nbr, n_send, n_global, has_gid}
pnp_status pnp_part_sizes(pnp_part* P, int level, long* sizes) {
  if (!P) return PNP_E_ARG;
  PART_TRY
  PNP_REQUIRE(level >= 0 && level <= P->levels && sizes && P->L[level].finalized, PNP_E_ARG, "level not finalized");
  const PartLevel& L = P->L[level];
  sizes[0] = (long)L.x.size(); sizes[1] = L.n_own; sizes[2] = (long)L.tri.size() / 3; sizes[3] = (long)L.ba.size();
  sizes[4] = (long)L.nbr.size(); sizes[5] = (long)L.send_idx.size(); sizes[6] = P->n_global; sizes[7] = L.gid.empty() ? 0 : 1;
  PART_CATCH(P)
}
// any output may be NULL
pnp_status pnp_part_get(pnp_part* P, int level, double* x, double* y, int* tri, int* ba, int* bb, int* bphys, int* nbr, int* send_ptr,
                        int* send_idx, int* recv_ptr, int* par0, int* par1, int* gid) {
  if (!P) return PNP_E_ARG;
  PART_TRY
  PNP_REQUIRE(level >= 0 && level <= P->levels && P->L[level].finalized, PNP_E_ARG, "level not finalized");
  const PartLevel& L = P->L[level];
  auto put = [](auto* dst, const auto& v) { if (dst) std::copy(v.begin(), v.end(), dst); };
  put(x, L.x); put(y, L.y); put(tri, L.tri); put(ba, L.ba); put(bb, L.bb); put(bphys, L.lm.bphys);
  put(nbr, L.nbr); put(send_ptr, L.send_ptr); put(send_idx, L.send_idx); put(recv_ptr, L.recv_ptr);
  put(par0, L.par0); put(par1, L.par1); put(gid, L.gid);
  PART_CATCH(P)
}

// The whole decomposition in one call, for drivers without a Python launcher (PnpSolverMain::run with N ranks,
// pnp_solver_main.cc:93-108).  `root` holds the GLOBAL coarse mesh (pnp_mesh_set / pnp_mesh_read_gmsh), the parameters and
// an initialised communicator (pnp_comm_init / pnp_comm_init_file) on every rank.  Afterwards root holds this rank's part
// of the mesh refined `levels` times (finalized, halo plan set), the coarser levels are registered as multigrid levels down
// to `replica_level`, below which a replica of the whole mesh continues on every rank (pnp_mg_set_coarse_replica).  The
// level contexts belong to root and go with it.
// Nodal fields (nfields > 0) given on the unpartitioned mesh of refinement level `field_level` -- values [nfields][nf] at
// the vertices (fx, fy)[nf], e.g. a coarse solution for nested iteration -- are injected there by bitwise coordinate match
// and P1-interpolated to the finest level; *start_vec receives a new vector of root holding them.
pnp_status pnp_partition_build(pnp_ctx* root, int levels, int replica_level, int field_level, int nfields, long nf,
                               const double* fx, const double* fy, const double* fields, int* start_vec) {
  if (!root) return PNP_E_ARG;
  Ctx& c = root->c;
  pnp_part* P = nullptr;
  try {
    PNP_CUDA(cudaSetDevice(c.device));
    PNP_REQUIRE(c.nv > 0 && c.n_own == c.nv && c.hier.empty(), PNP_E_ARG, "root must hold the unrefined global mesh");
    PNP_REQUIRE(c.params.set, PNP_E_ARG, "parameters not set");
    PNP_REQUIRE(levels >= 1 && replica_level >= 0 && replica_level <= levels - 1, PNP_E_ARG, "bad level arguments");
    PNP_REQUIRE(nfields == 0 || (nfields == 1 || nfields == 3), PNP_E_ARG, "1 or 3 fields");
    PNP_REQUIRE(nfields == 0 || (field_level >= 0 && field_level <= levels && fx && fy && fields && start_vec), PNP_E_ARG, "bad field arguments");
    const int world = c.world, me = c.rank;
    const long gnv = c.nv, gnT = c.nT, gnB = c.nB;
    std::vector<double> gx = c.cx.to_host(c.stream), gy = c.cy.to_host(c.stream);
    std::vector<int> gtri = c.ctri.to_host(c.stream), gba = c.cba.to_host(c.stream), gbb = c.cbb.to_host(c.stream), gph = c.cbphys.to_host(c.stream);
    pnp_status st = pnp_part_create(gnv, gx.data(), gy.data(), gnT, gtri.data(), gnB, gba.data(), gbb.data(), gph.data(), world, me, levels, &P);
    PNP_REQUIRE(st == PNP_OK, st, "partitioning failed");
    auto pck = [&](pnp_status s_) { if (s_ != PNP_OK) throw Error(s_, P->err); };
    for (int l = 0; l <= levels; l++) {
      long ng = 0;
      pck(pnp_part_ghost_keys(P, l, &ng, nullptr));
      std::vector<unsigned long long> keys(2 * (size_t)ng + 1);
      pck(pnp_part_ghost_keys(P, l, &ng, keys.data()));
      if (world > 1) {
        std::vector<long> cnt;
        std::vector<unsigned char> allk = comm_allgatherv(c, keys.data(), 16 * ng, cnt);
        std::vector<long> gptr(world + 1, 0);
        for (int r = 0; r < world; r++) gptr[r + 1] = gptr[r] + cnt[r] / 16;
        std::vector<long> cptr(world + 1, 0);
        pck(pnp_part_claim(P, l, gptr.data(), (const unsigned long long*)allk.data(), cptr.data(), nullptr));
        std::vector<long> blob(world + 1 + cptr[world]); // claim_ptr followed by claim_pos
        pck(pnp_part_claim(P, l, gptr.data(), (const unsigned long long*)allk.data(), blob.data(), blob.data() + world + 1));
        std::vector<unsigned char> allc = comm_allgatherv(c, blob.data(), (long)(blob.size() * sizeof(long)), cnt);
        std::vector<long> mptr(world + 1, 0), mpos;
        size_t off = 0;
        for (int r = 0; r < world; r++) {
          const long* b = (const long*)(allc.data() + off);
          mpos.insert(mpos.end(), b + world + 1 + b[me], b + world + 1 + b[me + 1]);
          mptr[r + 1] = (long)mpos.size();
          off += cnt[r];
        }
        mpos.push_back(0);
        pck(pnp_part_finalize(P, l, mptr.data(), mpos.data()));
      } else pck(pnp_part_finalize(P, l, nullptr, nullptr));
    }
    // nodal fields: injection at field_level, P1 interpolation upwards
    std::vector<double> f_fine;
    if (nfields > 0) {
      std::unordered_map<Key2, long, Key2Hash> at;
      at.reserve((size_t)nf * 2);
      for (long i = 0; i < nf; i++) at[Key2{dbits(fx[i]), dbits(fy[i])}] = i;
      const PartLevel& L0 = P->L[field_level];
      long nvl = (long)L0.x.size();
      std::vector<double> f((size_t)nfields * nvl);
      for (long v = 0; v < nvl; v++) {
        auto it = at.find(Key2{dbits(L0.x[v]), dbits(L0.y[v])});
        PNP_REQUIRE(it != at.end(), PNP_E_ARG, "a vertex of the field level has no counterpart in the given field");
        for (int k = 0; k < nfields; k++) f[(size_t)k * nvl + v] = fields[(size_t)k * nf + it->second];
      }
      for (int l = field_level + 1; l <= levels; l++) {
        const PartLevel& L = P->L[l];
        const long nvf = (long)L.x.size();
        std::vector<double> g((size_t)nfields * nvf);
        for (long v = 0; v < nvf; v++)
          for (int k = 0; k < nfields; k++) {
            const double a0 = f[(size_t)k * nvl + L.par0[v]];
            g[(size_t)k * nvf + v] = L.par1[v] < 0 ? a0 : 0.5 * (a0 + f[(size_t)k * nvl + L.par1[v]]);
          }
        f.swap(g); nvl = nvf;
      }
      f_fine.swap(f);
    }
    // library-side hierarchy: root <- finest level; one child context per coarser distributed level; the replica
    auto set_level = [&](pnp_ctx* h, const PartLevel& L) {
      Ctx& k = h->c;
      mesh_set(k, (long)L.x.size(), L.x.data(), L.y.data(), (long)L.tri.size() / 3, L.tri.data(), (long)L.ba.size(), L.ba.data(),
               L.bb.data(), L.lm.bphys.data(), L.n_own);
      halo_set(k, (int)L.nbr.size(), L.nbr.data(), L.send_ptr.data(), L.send_idx.data(), L.recv_ptr.data());
      mesh_finalize(k, true); // (builds the constraints as well: the parameters are set)
    };
    set_level(root, P->L[levels]);
    auto cck = [&](pnp_status s_, pnp_ctx* h) { if (s_ != PNP_OK) throw Error(s_, h->c.err); };
    for (int l = levels - 1; l >= replica_level; l--) {
      pnp_ctx* ch = nullptr;
      cck(pnp_ctx_create_child(root, &ch), root);
      c.owned_children.push_back(ch);
      set_level(ch, P->L[l]);
      cck(pnp_mg_push_level(root, ch, P->L[l + 1].par0.data(), P->L[l + 1].par1.data()), root);
    }
    {
      pnp_ctx* rep = nullptr;
      cck(pnp_ctx_create_child(root, &rep), root);
      c.owned_children.push_back(rep);
      Ctx& rc = rep->c;
      mesh_set(rc, gnv, gx.data(), gy.data(), gnT, gtri.data(), gnB, gba.data(), gbb.data(), gph.data(), gnv);
      std::vector<int> gid;
      long n_global = gnv;
      const PartLevel& LR = P->L[replica_level];
      if (replica_level == 0) gid = LR.gid;
      else { // vertices of the refined replica are matched by their (bitwise equal) coordinates
        mesh_refine(rc, replica_level);
        std::vector<double> rx = rc.cx.to_host(c.stream), ry = rc.cy.to_host(c.stream);
        std::unordered_map<Key2, long, Key2Hash> at;
        at.reserve(rx.size() * 2);
        for (size_t i = 0; i < rx.size(); i++) at[Key2{dbits(rx[i]), dbits(ry[i])}] = (long)i;
        gid.resize(LR.x.size());
        for (size_t v = 0; v < LR.x.size(); v++) {
          auto it = at.find(Key2{dbits(LR.x[v]), dbits(LR.y[v])});
          PNP_REQUIRE(it != at.end(), PNP_E_MESH, "replica level does not contain a vertex of the distributed level");
          gid[v] = (int)it->second;
        }
        n_global = (long)rx.size();
      }
      mesh_finalize(rc, true);
      cck(pnp_mg_set_coarse_replica(root, rep, gid.data(), n_global), root);
    }
    if (nfields > 0) {
      cck(pnp_vec_create(root, nfields, start_vec), root);
      vec_upload(c, c.vec(*start_vec), f_fine.data());
    }
    pnp_part_destroy(P);
    return PNP_OK;
  } catch (const pnp::Error& e) { c.err = e.what(); if (P) pnp_part_destroy(P); return e.code; }
  catch (const std::exception& e) { c.err = e.what(); if (P) pnp_part_destroy(P); return PNP_E_ARG; }
}

} // extern "C"
