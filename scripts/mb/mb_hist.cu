// Microbenchmark: what bounds the solver's histogram pass?  (development aid, not part of the library)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
constexpr int T = 512, NB = 8192, LB = 8, SF = 3072;

__global__ void gen(float* x, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint32_t h = (uint32_t)i * 2654435761u; h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
    uint32_t g = h * 1664525u + 1013904223u; g ^= g >> 15; g *= 0x2c1b3c6du; g ^= g >> 12;
    float u1 = (h >> 8) * (1.0f / 16777216.0f) + 1e-7f, u2 = (g >> 8) * (1.0f / 16777216.0f);
    float v = sqrtf(-2.0f * logf(u1)) * cosf(6.2831853f * u2);
    x[i] = fminf(fmaxf(v, -3.0f), 3.0f);
  }
}
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool tryw(uint32_t bar, uint32_t par) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(par) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void issue(const float* xr, long long len, uint32_t c, float* st, unsigned long long* bar) {
  long long base = (long long)c * SF; uint32_t bytes = (uint32_t)(min((long long)SF, len - base) * 4); uint32_t b = s32(bar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(st)), "l"(xr + base), "r"(bytes), "r"(b) : "memory");
}
// MODE bits: 1 fp64 S,Q  2 hist count atomic  4 sum atomic  8 match_any aggregation  16 64 bins/octave remap (fewer conflicts)
template <int MODE>
__device__ __forceinline__ void body(float v, uint32_t* hist, uint32_t* bsum, double& ls, double& lq, float& fs, uint32_t& kmn, uint32_t& kmx) {
  float a = fabsf(v); uint32_t k = __float_as_uint(a);
  fs += a;
  if (MODE & 1) { ls += (double)a; lq += (double)a * (double)a; kmn = min(kmn, k); kmx = max(kmx, k); }
  uint32_t b = k >> 18;
  if (MODE & 8) {
    uint32_t peers = __match_any_sync(__activemask(), b);
    int leader = __ffs(peers) - 1, lane = threadIdx.x & 31;
    uint32_t m9 = (k & 0x7FFFFFu) >> 9;
    // sum over peers: do a simple loop over set bits (leader only adds)
    uint32_t tot = 0, cnt = __popc(peers);
    for (uint32_t p = peers; p; p &= p - 1) tot += __shfl_sync(peers, m9, __ffs(p) - 1);
    if (lane == leader) { if (MODE & 2) atomicAdd(&hist[b], cnt); if (MODE & 4) atomicAdd(&bsum[b], tot); }
  } else {
    if (MODE & 2) atomicAdd(&hist[b], 1u);
    if (MODE & 4) atomicAdd(&bsum[b], (k & 0x7FFFFFu) >> 9);
  }
}
template <int MODE, bool TMA, int NST>
__global__ void __launch_bounds__(T, 2) pass1(const float* __restrict__ x, long long len, int skip, float* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint32_t* hist = (uint32_t*)smem; uint32_t* bsum = hist + NB;
  float* ring = (float*)(smem + NB * 8);
  __shared__ unsigned long long full[8]; __shared__ uint32_t done[8];
  const int tid = threadIdx.x, lane = tid & 31;
  const float* xr = x + (long long)blockIdx.x * len;
  const uint32_t n = (uint32_t)((len + skip - 1) / skip);
  for (int b = tid; b < NB; b += T) { hist[b] = 0; bsum[b] = 0; }
  double ls = 0, lq = 0; float fs = 0; uint32_t kmn = ~0u, kmx = 0;
  if (!TMA) {
    __syncthreads();
    for (uint32_t eb = 0; eb < n; eb += T * LB) {
      float v[LB];
#pragma unroll
      for (int u = 0; u < LB; ++u) { uint32_t e = eb + tid + u * T; v[u] = e < n ? __ldg(xr + (long long)e * skip) : 0.f; }
#pragma unroll
      for (int u = 0; u < LB; ++u) { if (eb + tid + u * T >= n) break; body<MODE>(v[u], hist, bsum, ls, lq, fs, kmn, kmx); }
    }
  } else {
    const uint32_t nch = (uint32_t)((len + SF - 1) / SF);
    if (tid == 0) {
      for (int s = 0; s < NST; ++s) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&full[s]))); done[s] = 0; }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      for (uint32_t c = 0; c < NST && c < nch; ++c) issue(xr, len, c, ring + c * SF, &full[c]);
    }
    __syncthreads();
    uint32_t s = 0, par = 0;
    for (uint32_t c = 0; c < nch; ++c) {
      while (!tryw(s32(&full[s]), par)) {}
      const float* st = ring + s * SF;
      long long base = (long long)c * SF, endf = min(len, base + SF);
      uint32_t e_lo = (uint32_t)((base + skip - 1) / skip), e_hi = (uint32_t)((endf + skip - 1) / skip);
      float v[LB];
#pragma unroll
      for (int u = 0; u < LB; ++u) { uint32_t e = e_lo + tid + u * T; v[u] = e < e_hi ? st[(long long)e * skip - base] : 0.f; }
#pragma unroll
      for (int u = 0; u < LB; ++u) { if (e_lo + tid + u * T >= e_hi) break; body<MODE>(v[u], hist, bsum, ls, lq, fs, kmn, kmx); }
      __syncwarp();
      if (lane == 0) { uint32_t old = atomicAdd(&done[s], 1u); if ((old + 1) % (T / 32) == 0 && c + NST < nch) issue(xr, len, c + NST, ring + s * SF, &full[s]); }
      if (++s == NST) { s = 0; par ^= 1; }
    }
  }
  __syncthreads();
  uint32_t hs = 0; for (int b = tid; b < NB; b += T) hs += hist[b] + bsum[b];
  if (fs + (float)ls + (float)lq + hs + kmn + kmx == 12345.678f) out[blockIdx.x] = 1.f;
}
template <int MODE, bool TMA, int NST>
void run(const char* name, const float* x, int rows, long long len, float* out) {
  size_t smem = NB * 8 + (TMA ? NST * SF * 4 : 0);
  if (smem < 100 * 1024) smem = 100 * 1024;    // two CTAs per SM like the solver
  CK(cudaFuncSetAttribute(pass1<MODE, TMA, NST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 2; ++i) pass1<MODE, TMA, NST><<<rows, T, smem>>>(x, len, 3, out);
  CK(cudaEventRecord(e0));
  for (int i = 0; i < 3; ++i) pass1<MODE, TMA, NST><<<rows, T, smem>>>(x, len, 3, out);
  CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 3;
  printf("%-34s %8.1f us  %6.0f GB/s (row bytes)\n", name, ms * 1e3, rows * len * 4.0 / ms / 1e6);
}
int main() {
  const int rows = 592; const long long len = 200704;
  float *x, *out; CK(cudaMalloc(&x, rows * len * 4)); CK(cudaMalloc(&out, rows * 4));
  gen<<<2048, 256>>>(x, (size_t)rows * len); CK(cudaDeviceSynchronize());
  run<0, false, 1>("direct loads only", x, rows, len, out);
  run<1, false, 1>("direct + fp64 S,Q", x, rows, len, out);
  run<3, false, 1>("direct + fp64 + hist", x, rows, len, out);
  run<7, false, 1>("direct + fp64 + hist + sum", x, rows, len, out);
  run<6, false, 1>("direct + hist + sum (no fp64)", x, rows, len, out);
  run<15, false, 1>("direct + all, match_any aggregated", x, rows, len, out);
  run<0, true, 3>("tma x3 loads only", x, rows, len, out);
  run<0, true, 6>("tma x6 loads only (1 CTA/SM?)", x, rows, len, out);
  run<7, true, 3>("tma x3 + all", x, rows, len, out);
  run<15, true, 3>("tma x3 + all, match_any", x, rows, len, out);
  return 0;
}
