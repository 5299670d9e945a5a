// Microbenchmark: tcgen05.mma kind::f16 (K = 16 per instruction = 32 operand bytes per row), no-swizzle K-major
// operands as the stem kernel lays them out: time per instruction as a function of M (128 / 64), N, and of where A
// comes from (shared memory descriptor or tensor memory).  Timing only.  Development aid, not part of the library.
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
__device__ __forceinline__ uint64_t mkdesc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// mode 0: A from shared memory; 1: A from tensor memory (columns 256..); 2: A from smem, both K chunks the same B (LBO = 0)
__global__ void __launch_bounds__(128, 1) k(int M, int N, int mode, int iters, int per_tile, long long* cycles, int bstep = 48, int lbo = 1920) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ unsigned long long bar;
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += 128) ((uint32_t*)smem)[i] = 0x3C003C00u;
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tslot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = tslot;
  const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);   // f16 x f16 -> f32
  long long t0 = 0, t1 = 0;
  if (warp == 0) {
    const uint32_t sb = s32(smem);
    t0 = clock64();
    if (lane == 0) {
      for (int it = 0; it < iters; ++it) {
        for (int p = 0; p < per_tile; ++p) {
          // A: 25 pair images of 4 KB at 0; B: a 16-byte-per-position patch at 112 KB, taps = shifts
          const uint32_t a0 = sb + (uint32_t)p * 4096u, b0 = sb + 114688u + (uint32_t)(it & 3) * 16384u + (uint32_t)p * (uint32_t)bstep;
          const uint64_t ad = mkdesc(a0, 2048u, 128u);
          const uint64_t bd = mkdesc(b0, mode == 2 ? 0u : (uint32_t)lbo, 128u);
          const uint32_t acc = (it | p) != 0;
          if (mode == 1) {
            const uint32_t at = tbase + 256u + (uint32_t)p * 8u;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tbase), "r"(at), "l"(bd), "r"(idesc), "r"(acc) : "memory");
          } else {
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tbase), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
          }
        }
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
void run(const char* name, int M, int N, int mode, long long* dcyc, int bstep = 48, int lbo = 1920) {
  const int iters = 400, per = 25;
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  k<<<148, 128, 200 * 1024>>>(M, N, mode, 40, per, dcyc, bstep, lbo);
  CK(cudaDeviceSynchronize());
  k<<<148, 128, 200 * 1024>>>(M, N, mode, iters, per, dcyc, bstep, lbo);
  CK(cudaDeviceSynchronize());
  long long cyc; CK(cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost));
  printf("%-44s M=%3d N=%3d  %7.1f clk/MMA  %6.0f MAC/clk/SM\n", name, M, N, (double)cyc / (iters * per), (double)M * N * 16 * iters * per / cyc);
}
int main() {
  long long* dcyc; CK(cudaMalloc(&dcyc, 8));
  run("f16 A smem, B 128-B aligned, LBO 1920", 128, 240, 0, dcyc, 128, 1920);
  run("f16 A smem, B +16 B steps, LBO 1920", 128, 240, 0, dcyc, 16, 1920);
  run("f16 A smem, B 128-B aligned, LBO 16", 128, 240, 0, dcyc, 128, 16);
  run("f16 A smem, B +16 B steps, LBO 16", 128, 240, 0, dcyc, 16, 16);
  run("f16 A smem, B +48 B steps, LBO 32", 128, 240, 0, dcyc, 48, 32);
  run("f16 A smem, B aligned, LBO 128", 128, 240, 0, dcyc, 128, 128);
  run("f16 A tmem, B aligned, LBO 1920", 128, 240, 1, dcyc, 128, 1920);
  for (int N : {240}) {
    run("f16 A smem", 128, N, 0, dcyc);
    run("f16 A tmem", 128, N, 1, dcyc);
    run("f16 A smem M64", 64, N, 0, dcyc);
    run("f16 A smem M64, both K chunks one B", 64, N, 2, dcyc);
    run("f16 A smem M128, both K chunks one B", 128, N, 2, dcyc);
  }
  return 0;
}
