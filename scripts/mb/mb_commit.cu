// Microbenchmark: latency of tcgen05.commit -> mbarrier phase flip (no instructions pending, and behind N MMAs),
// and of a plain mbarrier arrive -> try_wait wake-up in another warp.  Development aid.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool testw(uint32_t bar, uint32_t par) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(par) : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool tryw(uint32_t bar, uint32_t par) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(par), "r"(100000u) : "memory");
  return ok != 0;
}
__device__ __forceinline__ uint64_t mkdesc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__global__ void __launch_bounds__(128, 1) k(int nmma, int iters, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ unsigned long long bar, bar2, bar3;
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += 128) ((uint32_t*)smem)[i] = 0x3C003C00u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar2)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar3)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tslot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = tslot;
  const uint32_t idesc = (1u << 4) | ((uint32_t)(240 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  long long tot_commit = 0, tot_pp = 0;
  if (warp == 0 && lane == 0) {
    const uint32_t sb = s32(smem);
    uint32_t par = 0;
    for (int it = 0; it < iters; ++it) {
      for (int p = 0; p < nmma; ++p)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tbase), "l"(mkdesc(sb, 2048u, 128u)), "l"(mkdesc(sb + 8192u + p * 48u, 1920u, 128u)), "r"(idesc), "r"(1u) : "memory");
      const long long t0 = clock64();
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
      while (!testw(s32(&bar), par)) {}
      tot_commit += clock64() - t0;
      par ^= 1u;
    }
    out[0] = tot_commit / iters;
  }
  // ping-pong between warp 1 (lane 0) and warp 2 (lane 0) through two mbarriers: round trip / 2 = arrive -> wake
  if (warp == 1 && lane == 0) {
    uint32_t par = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&bar2)) : "memory");
      while (!tryw(s32(&bar3), par)) {}
      par ^= 1u;
    }
    tot_pp = clock64() - t0;
    out[1] = tot_pp / iters;
  }
  if (warp == 2 && lane == 0) {
    uint32_t par = 0;
    for (int it = 0; it < iters; ++it) {
      while (!tryw(s32(&bar2), par)) {}
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&bar3)) : "memory");
      par ^= 1u;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512) : "memory");
}
int main() {
  long long* d; CK(cudaMalloc(&d, 64));
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  for (int n : {0, 1, 2, 4, 7, 14, 25}) {
    k<<<1, 128, 64 * 1024>>>(n, 200, d);
    CK(cudaDeviceSynchronize());
    long long h[2]; CK(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
    printf("MMAs before commit %2d: commit -> flip observed %6lld clk (per MMA %.1f) | arrive/try_wait ping-pong round trip %lld clk\n", n, h[0], n ? (double)h[0] / n : 0.0, h[1]);
  }
  return 0;
}
