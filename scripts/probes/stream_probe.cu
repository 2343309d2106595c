// stream_probe.cu -- how fast can ONE persistent CTA per SM pull a contiguous stream from HBM into shared memory?
// Variants: (a) cp.async.bulk (TMA) issued by one thread, S stages of B bytes, K copies per stage;
//           (b) cp.async 16 B (LDGSTS) issued by P producer warps; (c) plain LDG.128 + sum (registers).
// Consumers only wait for a stage and release it (no compute): this is the feed rate the SpMV design can count on.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ unsigned s32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, unsigned c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, unsigned n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, unsigned par) {
  unsigned done;
  do { asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(s32(b)), "r"(par) : "memory"); } while (!done);
}
__device__ __forceinline__ void bulk(void* d, const void* s, unsigned n, uint64_t* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(d)), "l"(s), "r"(n), "r"(s32(b)) : "memory");
}

// (a) TMA: thread 0 of warp `nw` produces, nw consumer warps release
__global__ void k_tma(const char* src, size_t total, int stage_bytes, int nstages, int ncopies, int nw, unsigned long long* sink, size_t stream_stride = 0, int src_off = 0, int shrink = 0) {
  extern __shared__ __align__(128) unsigned char sm[];
  uint64_t* full = (uint64_t*)sm; uint64_t* empty = full + 8;
  unsigned char* buf = sm + 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { for (int s = 0; s < nstages; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], nw); } asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  const long ntiles = (stream_stride ? stream_stride * ncopies : total) / stage_bytes;
  if (warp == nw) {
    if (lane == 0) {
      int it = 0;
      for (long t = blockIdx.x; t < ntiles; t += gridDim.x, it++) {
        const int st = it % nstages; const unsigned par = (it / nstages) & 1;
        mbar_wait(&empty[st], par ^ 1u);
        const int cb = stage_bytes / ncopies;
        mbar_expect(&full[st], (cb - shrink) * ncopies);
        for (int k = 0; k < ncopies; k++) {
          const char* sk = stream_stride ? src + (size_t)k * stream_stride + (size_t)t * cb : src + (size_t)t * stage_bytes + (size_t)k * cb;
          bulk(buf + (size_t)st * stage_bytes + (size_t)k * cb, sk + src_off, cb - shrink, &full[st]);
        }
      }
    }
  } else {
    int it = 0; unsigned long long acc = 0;
    for (long t = blockIdx.x; t < ntiles; t += gridDim.x, it++) {
      const int st = it % nstages; const unsigned par = (it / nstages) & 1;
      mbar_wait(&full[st], par);
      acc += buf[(size_t)st * stage_bytes + threadIdx.x];
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[st]);
    }
    if (acc == 0x123456789ull) *sink = acc;
  }
}

// (b) LDGSTS: np producer warps issue 16-byte cp.async; completion through cp.async.mbarrier.arrive.noinc
__global__ void k_ldgsts(const char* src, size_t total, int stage_bytes, int nstages, int np, int nw, unsigned long long* sink) {
  extern __shared__ __align__(128) unsigned char sm[];
  uint64_t* full = (uint64_t*)sm; uint64_t* empty = full + 8;
  unsigned char* buf = sm + 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { for (int s = 0; s < nstages; s++) { mbar_init(&full[s], np * 32); mbar_init(&empty[s], nw); } asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  const long ntiles = total / stage_bytes;
  if (warp >= nw) {
    const int pt = (warp - nw) * 32 + lane, npt = np * 32;
    int it = 0;
    for (long t = blockIdx.x; t < ntiles; t += gridDim.x, it++) {
      const int st = it % nstages; const unsigned par = (it / nstages) & 1;
      mbar_wait(&empty[st], par ^ 1u);
      const char* s0 = src + (size_t)t * stage_bytes; unsigned char* d0 = buf + (size_t)st * stage_bytes;
      for (int o = pt * 16; o < stage_bytes; o += npt * 16)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s32(d0 + o)), "l"(s0 + o) : "memory");
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(s32(&full[st])) : "memory");
    }
  } else {
    int it = 0; unsigned long long acc = 0;
    for (long t = blockIdx.x; t < ntiles; t += gridDim.x, it++) {
      const int st = it % nstages; const unsigned par = (it / nstages) & 1;
      mbar_wait(&full[st], par);
      acc += buf[(size_t)st * stage_bytes + threadIdx.x];
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[st]);
    }
    if (acc == 0x123456789ull) *sink = acc;
  }
}

// (c) plain 16-byte loads into registers
__global__ void k_ldg(const uint4* src, size_t n16, unsigned long long* sink) {
  unsigned acc = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) { uint4 v = src[i]; acc += v.x ^ v.y ^ v.z ^ v.w; }
  if (acc == 0x12345678u) *sink = acc;
}

int main() {
  const size_t total = 16ull << 30; // 16 GiB stream
  char* src; CK(cudaMalloc(&src, total)); CK(cudaMemset(src, 1, total));
  unsigned long long* sink; CK(cudaMalloc(&sink, 8));
  int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto run = [&](const char* name, auto launch) {
    launch(); CK(cudaDeviceSynchronize());
    cudaEventRecord(e0); launch(); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%-60s %8.3f ms %8.1f GB/s\n", name, ms, total / ms / 1e6); fflush(stdout);
  };
  CK(cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  CK(cudaFuncSetAttribute(k_ldgsts, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  char nm[200];
  run("ldg.128, 148*8 blocks x 256", [&] { k_ldg<<<sms * 8, 256>>>((const uint4*)src, total / 16, sink); });
  for (int ctas = 1; ctas <= 1; ctas *= 2)
    for (int sb : {65536})
      for (int ns : {2, 3})
        for (int nc : {1, 8}) {
          const int smem = 128 + sb * ns;
          if (smem * ctas > 220 * 1024) continue;
          snprintf(nm, sizeof nm, "tma  ctas/SM %d stage %3d KB x %d stages, %d copies/stage", ctas, sb / 1024, ns, nc);
          run(nm, [&] { k_tma<<<sms * ctas, 5 * 32, smem>>>(src, total, sb, ns, nc, 4, sink); });
        }
  // the SpMV's pattern: 8 streams far apart, misaligned starts, odd sizes
  for (int ns : {2, 3})
    for (int mode = 0; mode < 4; mode++) {
      const int sb = 65536, nc = 8;
      const size_t stride = (mode & 1) ? (total / 8) : 0;
      const int off = (mode & 2) ? 16 : 0, shr = (mode & 2) ? 48 : 0;
      snprintf(nm, sizeof nm, "tma 64 KB x %d stages, 8 copies: %s, %s", ns, stride ? "8 streams 2 GiB apart" : "contiguous", off ? "misaligned +16 B, size -48 B" : "aligned");
      run(nm, [&] { k_tma<<<sms, 5 * 32, 128 + sb * ns>>>(src, total, sb, ns, nc, 4, sink, stride, off, shr); });
    }
  for (int ctas = 1; ctas <= 1; ctas *= 2)
    for (int sb : {65536})
      for (int ns : {2})
        for (int np : {2}) {
          const int smem = 128 + sb * ns;
          if (smem * ctas > 220 * 1024) continue;
          snprintf(nm, sizeof nm, "ldgsts ctas/SM %d stage %3d KB x %d stages, %d producer warps", ctas, sb / 1024, ns, np);
          run(nm, [&] { k_ldgsts<<<sms * ctas, (4 + np) * 32, smem>>>(src, total, sb, ns, np, 4, sink); });
        }
  return 0;
}
