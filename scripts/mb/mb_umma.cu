// Microbenchmark: tcgen05.mma throughput for 8-bit operands (kind::i8 / kind::f8f6f4), M=128, K=32 per
// instruction, as a function of N and of the shared-memory layout (no swizzle vs 64B / 128B swizzle).
// Timing only: operand contents are arbitrary.  Development aid, not part of the library.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool tryw(uint32_t bar, uint32_t par) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(par) : "memory");
  return ok != 0;
}
// layout: 0 none, 2 = 128B swizzle, 4 = 64B, 6 = 32B (sm_100 descriptor layout_type, bits 61-63)
__device__ __forceinline__ uint64_t mkdesc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) |
         (1ull << 46) | ((uint64_t)layout << 61);
}
template <int KIND>   // 0: i8, 1: f8f6f4 (e4m3 x e4m3 -> f32)
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (KIND == 0)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
template <int KIND>
__global__ void __launch_bounds__(128, 1) k(int N, int layout, int iters, int kchunks, long long* cycles, int a_shift, int M, int commit_each, int b_shift) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ unsigned long long bar;
  __shared__ unsigned long long bar2[8];
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += 128) ((uint32_t*)smem)[i] = 0x01010101u;
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar2[i]))); asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tslot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = tslot;
  // idesc: c fmt (bits 4-5): 2 = s32 (i8) / 1 = f32; a fmt (7-9), b fmt (10-12): i8: 1 = s8; f8f6f4: 0 = e4m3
  uint32_t idesc;
  if (KIND == 0) idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  else idesc = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  long long t0 = 0, t1 = 0;
  if (warp == 0) {
    const uint32_t sb = s32(smem);
    // A region: 4 stages x 32 KB at 0; B region: 4 stages x 16 KB at 128 KB
    t0 = clock64();
    if (lane == 0) {
      for (int it = 0; it < iters; ++it) {
        const uint32_t a0 = sb + (uint32_t)(it & 1) * 32768u + (uint32_t)a_shift * ((it >> 1) & 7), b0 = sb + 65536u + (uint32_t)(it & 3) * 32768u + (uint32_t)b_shift * ((it >> 1) & 3);
        for (int kc = 0; kc < kchunks; ++kc) {
          uint64_t ad, bd;
          if (layout == 0) {   // K-major, core matrices 8 x 16 B contiguous, LBO between K-adjacent (2 KB / N*16), SBO 128 B
            ad = mkdesc(a0 + kc * 2 * 2048u, 2048u, 128u, 0);
            bd = mkdesc(b0 + kc * 2 * (uint32_t)N * 16u, (uint32_t)N * 16u, 128u, 0);
          } else if (layout == 2) {   // 128-byte rows, 8-row atoms of 1024 B, K slice = 32 B inside the row
            ad = mkdesc(a0 + kc * 32u, 16u, 1024u, 2);
            bd = mkdesc(b0 + kc * 32u, 16u, 1024u, 2);
          } else {                    // 64-byte rows, 8-row atoms of 512 B
            ad = mkdesc(a0 + (kc & 1) * 32u, 16u, 512u, 4);
            bd = mkdesc(b0 + (kc & 1) * 32u, 16u, 512u, 4);
          }
          umma<KIND>(tbase, ad, bd, idesc, (it | kc) != 0);
        }
        if (commit_each) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar2[it & 7])) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
    }
    __syncwarp();
    while (!tryw(s32(&bar), 0)) {}
    t1 = clock64();
    if (lane == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512) : "memory");
}
template <int KIND>
void run(const char* name, int N, int layout, long long* dcyc, int a_shift = 0, int M = 128, int commit_each = 0, int b_shift = 0) {
  const int iters = 2000, kch = 4;
  CK(cudaFuncSetAttribute(k<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  k<KIND><<<148, 128, 200 * 1024>>>(N, layout, 200, kch, dcyc, a_shift, M, commit_each, b_shift);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<KIND><<<148, 128, 200 * 1024>>>(N, layout, iters, kch, dcyc, a_shift, M, commit_each, b_shift);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long cyc; CK(cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost));
  const double ops = 2.0 * M * N * 32 * (double)iters * kch;
  printf("%-40s N=%3d  %8.1f clk/MMA  %7.0f ops/clk/SM  %6.2f POP/s (148 SMs, event time %.3f ms)\n", name, N, (double)cyc / (iters * kch),
         ops / cyc, ops * 148 / (ms * 1e-3) / 1e15, ms);
}
int main() {
  long long* dcyc; CK(cudaMalloc(&dcyc, 8));
  run<0>("i8 M128 N256 aligned", 256, 0, dcyc);
  run<0>("i8 M128 N256 B start + k*32 B", 256, 0, dcyc, 0, 128, 0, 32);
  run<0>("i8 M128 N256 B start + k*16 B", 256, 0, dcyc, 0, 128, 0, 16);
  run<0>("i8 M128 N256 A and B shifted", 256, 0, dcyc, 48, 128, 0, 32);
  run<0>("i8 M128 N256 64B swizzle aligned", 256, 4, dcyc);
  run<0>("i8 M128 N256 64B swizzle B + k*64 B", 256, 4, dcyc, 0, 128, 0, 64);
  return 0;
}
