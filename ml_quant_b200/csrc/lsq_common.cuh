// Shared helpers for the lsq_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include "../../include/lsq_b200.h"

namespace lsq {

void set_error(const char* fmt, ...);

#define LSQ_CHECK_ARG(cond, ...)                 \
  do {                                           \
    if (!(cond)) {                               \
      lsq::set_error(__VA_ARGS__);               \
      return LSQ_ERR_ARG;                        \
    }                                            \
  } while (0)

#define LSQ_CUDA_LAUNCH_CHECK(what)                                                   \
  do {                                                                                \
    cudaError_t e__ = cudaGetLastError();                                             \
    if (e__ != cudaSuccess) {                                                         \
      lsq::set_error("%s: %s", what, cudaGetErrorString(e__));                        \
      return LSQ_ERR_CUDA;                                                            \
    }                                                                                 \
  } while (0)

// Per-device facts that never change, looked up once (immutable after initialisation: the library still
// holds no mutable state that calls could race on).
inline int device_index() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev;
}
inline int device_sms() {
  static std::atomic<int> cache[64];
  const int dev = device_index();
  if (dev < 0 || dev >= 64) return 148;
  int v = cache[dev].load(std::memory_order_relaxed);
  if (v == 0) {
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    cache[dev].store(v, std::memory_order_relaxed);
  }
  return v;
}
// Opt a kernel into the full 227 KB of dynamic shared memory, once per device (`done` = one bit per device).
template <class K>
inline cudaError_t ensure_max_smem(K kernel, std::atomic<unsigned long long>& done) {
  const unsigned long long bit = 1ull << (device_index() & 63);
  if (done.load(std::memory_order_acquire) & bit) return cudaSuccess;
  const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e == cudaSuccess) done.fetch_or(bit, std::memory_order_release);
  return e;
}

// sign with sign(0) = +1, as a float (+-1).  quant/binary/ste.py:16-18
__device__ __forceinline__ float sign_pm1(float x) { return x >= 0.0f ? 1.0f : -1.0f; }

// clamp(x, -alpha, alpha) when alpha > 0.  quant/binary/quantization.py:22-24
__device__ __forceinline__ float clamp_sym(float x, float alpha) {
  return alpha > 0.0f ? fminf(fmaxf(x, -alpha), alpha) : x;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ uint32_t warp_min_u32(uint32_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ uint32_t warp_max_u32(uint32_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Deterministic block-wide sum (fixed tree).  `red` holds >= 32 doubles.  Result valid in all threads.
__device__ __forceinline__ double block_sum(double v, double* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  double t = (lane < nw) ? red[lane] : 0.0;
  t = warp_sum(t);
  return t;
}

// Fused per-channel affine prologue (eval BatchNorm) and convolution epilogue, device-side copies.
struct Prologue {
  const float* a;
  const float* b;
  int channels;
  long long inner;
  unsigned long long magic;   // ceil(2^40 / inner): index / inner == (index * magic) >> 40 while index * inner < 2^40;
                              // 0 when the row is too long for that: the kernels then divide exactly
};
// channel of element `index_in_row` (row = channels * inner elements)
__device__ __forceinline__ unsigned prologue_channel(const Prologue& p, long long index_in_row) {
  const unsigned c = p.magic ? (unsigned)(((unsigned long long)index_in_row * p.magic) >> 40)
                             : (unsigned)((unsigned long long)index_in_row / (unsigned long long)p.inner);
  return min(c, (unsigned)p.channels - 1u);
}
__host__ inline Prologue to_dev(const lsq_prologue* p) {
  Prologue d{nullptr, nullptr, 1, 1, 1ull << 40};
  if (p && p->d_ch_scale && p->d_ch_shift && p->inner > 0) {
    d.a = p->d_ch_scale; d.b = p->d_ch_shift; d.channels = p->channels; d.inner = p->inner;
    d.magic = ((1ull << 40) + (unsigned long long)p->inner - 1ull) / (unsigned long long)p->inner;
    // (index * magic) >> 40 is exact while index * (magic * inner - 2^40) < 2^40, i.e. surely while
    // row_length * inner <= 2^40; longer rows (feature maps beyond ~1000 x 1000) take the exact division
    if ((double)p->channels * (double)p->inner * (double)p->inner > 1099511627776.0) d.magic = 0ull;
  }
  return d;
}
// the caller guarantees row length == channels * inner (one row = one sample), so index / inner < channels
__device__ __forceinline__ float apply_prologue(const Prologue& p, float x, long long index_in_row) {
  if (p.a == nullptr) return x;
  const unsigned c = prologue_channel(p, index_in_row);
  return fmaf(x, __ldg(p.a + c), __ldg(p.b + c));
}
struct Epilogue {
  const float* residual;
  const float* prelu;
  int n_prelu, act, residual_after_act;
};
__host__ inline Epilogue to_dev(const lsq_epilogue* e) {
  Epilogue d{nullptr, nullptr, 0, 0, 1};
  if (e) { d.residual = e->d_residual; d.prelu = e->d_prelu; d.n_prelu = e->n_prelu; d.act = e->act; d.residual_after_act = e->residual_after_act; }
  return d;
}
// r = vw * sum(s_j I_j) + bias already formed; apply activation / residual in the requested order
__device__ __forceinline__ float apply_epilogue_res(const Epilogue& e, float r, int channel, float res) {
  if (!e.residual_after_act) r += res;
  if (e.act == 1) r = fmaxf(r, 0.0f);
  else if (e.act == 2) r = r >= 0.0f ? r : r * __ldg(e.prelu + (e.n_prelu > 1 ? channel : 0));
  if (e.residual_after_act) r += res;
  return r;
}
__device__ __forceinline__ float apply_epilogue(const Epilogue& e, float r, int channel, long long out_index) {
  const float res = e.residual ? __ldg(e.residual + out_index) : 0.0f;
  if (!e.residual_after_act) r += res;
  if (e.act == 1) r = fmaxf(r, 0.0f);
  else if (e.act == 2) r = r >= 0.0f ? r : r * __ldg(e.prelu + (e.n_prelu > 1 ? channel : 0));
  if (e.residual_after_act) r += res;
  return r;
}

// Geometry helpers shared by the encoder and the convolution kernels (see lsq_b200.h).
struct ActGeom {
  int n, c, h, w, kh, kw, stride, pad, ho, wo, cw, nphase, hv, wv, ph, pitch, rps, lead;
  long long vtot;
};
__host__ __device__ inline ActGeom to_dev(const lsq_act_geom& g) {
  ActGeom d;
  d.n = g.n; d.c = g.c; d.h = g.h; d.w = g.w; d.kh = g.kh; d.kw = g.kw; d.stride = g.stride;
  d.pad = g.pad; d.ho = g.ho; d.wo = g.wo; d.cw = g.cw; d.nphase = g.nphase; d.hv = g.hv;
  d.wv = g.wv; d.ph = g.ph; d.pitch = g.pitch; d.rps = g.rows_per_sample; d.lead = g.lead;
  d.vtot = g.vtot;
  return d;
}
// virtual position of phase coordinate (a, b) of sample s; a in [-ph, hv), b in [-ph, wv+ph)
__host__ __device__ inline long long vpos(const ActGeom& g, int s, int a, int b) {
  return (long long)g.lead + ((long long)s * g.rps + g.ph + a) * g.pitch + b;
}

// Streaming loads of the quantizer sweeps: read-only path with a 256-byte L2 prefetch hint (the sweeps are bound by the
// number of outstanding requests per SM, not by bytes: fetching the neighbouring 128 bytes into L2 with the same request
// measured -2 % on the 64 x 56 x 56 rows and -5 % on the 512 x 7 x 7 rows; adding L1::no_allocate cost 14 %, the
// sampled sweep re-uses the lines it fetched)
__device__ __forceinline__ float4 ldg_stream4(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L2::256B.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float ldg_stream(const float* p) {
  float v;
  asm volatile("ld.global.nc.L2::256B.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}

}  // namespace lsq
