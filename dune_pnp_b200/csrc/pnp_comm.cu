// pnp_comm.cu -- the only inter-GPU traffic of the path: halo values of a vector before an SpMV or an
// assembly, and the scalar sums of dot products / norms.  One process per GPU; NCCL over NVLink/NVSwitch.
//
// Replaces PDELab's non-overlapping parallel layer that the reference selects with GridOperator<...,true>
// and the ISTLBackend_NOVLP_* backends (/root/reference/src/stationary_pnp.hh:240, :254-256;
// instationary_pnp_from_pb_md.hh:185-211): there every rank assembles its interior
// elements and border dofs are made consistent by an ADD-exchange after each operator application; here
// every rank owns whole rows (it holds all elements around its owned vertices), so one COPY-exchange of
// the SpMV input replaces the add-exchange of the output, and assembly needs no communication beyond the
// ghost values of the state.
#include <nccl.h>

#include <cstring>
#include <ctime>

#include "pnp_common.cuh"

namespace pnp {

#define PNP_NCCL(call)                                                                                 \
  do {                                                                                                 \
    ncclResult_t r_ = (call);                                                                          \
    if (r_ != ncclSuccess)                                                                             \
      throw ::pnp::Error(PNP_E_CUDA, std::string(__FILE__) + ":" + std::to_string(__LINE__) + " " +    \
                                         #call + " -> " + ncclGetErrorString(r_));                     \
  } while (0)

namespace {
__global__ void k_pack(const double* __restrict__ x, const int* __restrict__ idx, long n, int F, double* __restrict__ buf) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n * F; i += (long)gridDim.x * blockDim.x) {
    const long v = idx[i / F]; const int k = (int)(i % F);
    buf[i] = x[F * v + k];
  }
}
} // namespace

void comm_unique_id(char* out128) {
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  ncclUniqueId id;
  PNP_NCCL(ncclGetUniqueId(&id));
  std::memcpy(out128, &id, 128);
}

void comm_init(Ctx& c, int rank, int world, const char* unique_id128) {
  PNP_REQUIRE(world >= 1 && rank >= 0 && rank < world, PNP_E_ARG, "bad rank / world size");
  c.rank = rank; c.world = world;
  if (world == 1) return;
  ncclUniqueId id;
  std::memcpy(&id, unique_id128, 128);
  ncclComm_t comm;
  PNP_NCCL(ncclCommInitRank(&comm, world, id, rank));
  c.nccl = comm; c.owns_comm = true;
}
// Tear-down of the communicator a context created.  ncclCommDestroy is collective in effect (it waits for the peers'
// side of the communicator), and contexts are destroyed whenever their owner lets go of them -- a Python finaliser at
// interpreter exit, in any order across ranks -- where that wait never ends (measured: both ranks of a finished 2-GPU
// run sat in it until the launcher was killed).  The stream has been synchronised, nothing is in flight:
// ncclCommAbort frees the local resources without waiting for anyone.
void comm_destroy(Ctx& c) {
  if (c.nccl && c.owns_comm) ncclCommAbort((ncclComm_t)c.nccl);
  c.nccl = nullptr; c.owns_comm = false;
}

// NCCL needs its unique id on every rank before the communicator exists.  A DUNE build would broadcast it over MPI; the
// Python launchers use torch.distributed; a plain C++ driver started by any process launcher uses a file all ranks see:
// rank 0 writes the id (write + rename: never seen half written), the others wait for it.
void comm_bootstrap_file(Ctx& c, int rank, int world, const std::string& path) {
  if (world == 1) { comm_init(c, 0, 1, nullptr); return; }
  char id[128];
  if (rank == 0) {
    comm_unique_id(id);
    const std::string tmp = path + ".tmp";
    std::FILE* f = std::fopen(tmp.c_str(), "wb");
    PNP_REQUIRE(f, PNP_E_CONFIG, "cannot write " + tmp);
    std::fwrite(id, 1, 128, f); std::fclose(f);
    PNP_REQUIRE(std::rename(tmp.c_str(), path.c_str()) == 0, PNP_E_CONFIG, "cannot rename " + tmp);
  } else {
    std::FILE* f = nullptr;
    for (int tries = 0; tries < 6000 && !f; tries++) { // up to 10 minutes
      f = std::fopen(path.c_str(), "rb");
      if (!f) { struct timespec ts = {0, 100000000}; nanosleep(&ts, nullptr); }
    }
    PNP_REQUIRE(f, PNP_E_CONFIG, "rendezvous file " + path + " never appeared");
    const size_t n = std::fread(id, 1, 128, f); std::fclose(f);
    PNP_REQUIRE(n == 128, PNP_E_CONFIG, "rendezvous file " + path + " is truncated");
  }
  comm_init(c, rank, world, id);
}

// all-gather of byte blocks of different lengths (set-up plumbing: partition plans): counts[r] = length of rank r's block,
// returned: the blocks in rank order
std::vector<unsigned char> comm_allgatherv(Ctx& c, const void* send, long nbytes, std::vector<long>& counts) {
  counts.assign(c.world, 0);
  if (c.world == 1) { counts[0] = nbytes; return std::vector<unsigned char>((const unsigned char*)send, (const unsigned char*)send + nbytes); }
  PNP_REQUIRE(c.nccl, PNP_E_ARG, "no communicator (pnp_comm_init)");
  ncclComm_t comm = (ncclComm_t)c.nccl;
  DBuf<long> d_cnt(c.world), d_mine(1);
  d_mine.upload(&nbytes, 1, c.stream);
  PNP_NCCL(ncclAllGather(d_mine.p, d_cnt.p, 1, ncclInt64, comm, c.stream));
  d_cnt.download(counts.data(), c.world, c.stream);
  long mx = 0;
  for (long v : counts) mx = std::max(mx, v);
  mx = (mx + 15) & ~15l;
  if (mx == 0) return {};
  DBuf<unsigned char> d_send(mx), d_recv((size_t)mx * c.world);
  if (nbytes) d_send.upload((const unsigned char*)send, nbytes, c.stream);
  PNP_NCCL(ncclAllGather(d_send.p, d_recv.p, mx, ncclChar, comm, c.stream));
  std::vector<unsigned char> padded = d_recv.to_host(c.stream), out;
  for (int r = 0; r < c.world; r++) out.insert(out.end(), padded.begin() + (size_t)r * mx, padded.begin() + (size_t)r * mx + counts[r]);
  return out;
}

void halo_set(Ctx& c, int n_nbr, const int* nbr, const int* send_ptr, const int* send_idx, const int* recv_ptr) {
  PNP_REQUIRE(n_nbr >= 0, PNP_E_ARG, "negative neighbour count");
  c.halo_nbr.assign(nbr, nbr + n_nbr);
  c.halo_send_ptr.assign(send_ptr, send_ptr + n_nbr + 1);
  c.halo_recv_ptr.assign(recv_ptr, recv_ptr + n_nbr + 1);
  c.halo_send_ext.assign(send_idx, send_idx + send_ptr[n_nbr]);
  PNP_REQUIRE(recv_ptr[n_nbr] == c.nv - c.n_own, PNP_E_ARG, "halo plan does not cover the ghost vertices");
  for (int v : c.halo_send_ext) PNP_REQUIRE(v >= 0 && v < c.n_own, PNP_E_ARG, "send list names a vertex that is not owned");
  if (c.finalized) halo_finalize(c);
}

// translates the send list to internal numbering (owned vertices are renumbered by mesh_finalize)
void halo_finalize(Ctx& c) {
  if (c.halo_send_ext.empty()) return;
  std::vector<int> e2i = c.ext2int.to_host(c.stream), idx(c.halo_send_ext.size());
  for (size_t i = 0; i < idx.size(); i++) idx[i] = e2i[c.halo_send_ext[i]];
  c.halo_send_idx.alloc(idx.size());
  c.halo_send_idx.upload(idx.data(), idx.size(), c.stream);
  c.halo_send_buf.alloc(3 * idx.size());
  PNP_CUDA(cudaStreamSynchronize(c.stream));
}

void halo_exchange(Ctx& c, double* x, int F) {
  if (c.world == 1 || c.halo_nbr.empty()) return;
  PNP_REQUIRE(c.nccl, PNP_E_ARG, "partitioned mesh without a communicator (pnp_comm_init)");
  ncclComm_t comm = (ncclComm_t)c.nccl;
  const long ns = (long)c.halo_send_ext.size();
  if (ns) {
    k_pack<<<grid_for(ns * F, 256), 256, 0, c.stream>>>(x, c.halo_send_idx.p, ns, F, c.halo_send_buf.p);
    PNP_CHECK_LAUNCH(); c.launches++;
  }
  PNP_NCCL(ncclGroupStart());
  for (size_t i = 0; i < c.halo_nbr.size(); i++) {
    const long s0 = c.halo_send_ptr[i], s1 = c.halo_send_ptr[i + 1], r0 = c.halo_recv_ptr[i], r1 = c.halo_recv_ptr[i + 1];
    if (s1 > s0) PNP_NCCL(ncclSend(c.halo_send_buf.p + F * s0, (size_t)F * (s1 - s0), ncclDouble, c.halo_nbr[i], comm, c.stream));
    if (r1 > r0) PNP_NCCL(ncclRecv(x + F * (c.n_own + r0), (size_t)F * (r1 - r0), ncclDouble, c.halo_nbr[i], comm, c.stream));
  }
  PNP_NCCL(ncclGroupEnd());
}

void allreduce_sum(Ctx& c, double* dev, size_t n) {
  if (c.world == 1) return;
  PNP_REQUIRE(c.nccl, PNP_E_ARG, "no communicator (pnp_comm_init)");
  PNP_NCCL(ncclAllReduce(dev, dev, n, ncclDouble, ncclSum, (ncclComm_t)c.nccl, c.stream));
}

void allreduce_max_u64(Ctx& c, unsigned long long* dev, size_t n) {
  if (c.world == 1) return;
  PNP_REQUIRE(c.nccl, PNP_E_ARG, "no communicator (pnp_comm_init)");
  PNP_NCCL(ncclAllReduce(dev, dev, n, ncclUint64, ncclMax, (ncclComm_t)c.nccl, c.stream));
}

} // namespace pnp
