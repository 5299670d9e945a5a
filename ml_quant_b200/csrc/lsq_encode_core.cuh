// The activation encoder's work item, shared by lsq_quant.cu (stand-alone encoder) and lsq_qact.cu (fused
// solve + encode): one thread turns 32 channels x VEC pixels of an NCHW fp32 tensor into bit-plane words.
#pragma once
#include "lsq_common.cuh"

namespace lsq {

// Activation encoder.  One work item = VEC consecutive pixels x one 32-channel group: the thread loads the
// 32 channels (VEC = 4: one 16-byte load per channel, coalesced along W of the NCHW input, 8 loads in
// flight), folds the per-channel affine prologue and the clamp, and shifts one sign bit per element and plane
// into a register word with a funnel shift (channels are walked from 31 down to 0, so channel c lands on bit
// c).  The sign bit IS the reference's sign(x) = [x >= 0] because no value tested here can be -0.0: the
// prologue is always applied as fma(x, a, b) with b = -0.0 replaced by +0.0 (identity: a = 1, b = +0.0, which
// maps -0.0 to +0.0 and every other float to itself), and u - v is never -0.0 for u != -0.0.
// s * sign(d) is formed exactly by xor-ing d's sign bit into s.
template <int NPL, int VEC, bool FULL>
__device__ __forceinline__ void encode_group(const float* __restrict__ xp, long long cstride, int cn,
                                             const float2* __restrict__ ab, float alpha, const float (&sc)[NPL], int ns,
                                             uint32_t (&word)[VEC][NPL], float (&gsum)[VEC]) {
#pragma unroll
  for (int p = 0; p < VEC; ++p) {
    gsum[p] = 0.0f;
#pragma unroll
    for (int j = 0; j < NPL; ++j) word[p][j] = 0u;
  }
  constexpr int kBatch = 8;
#pragma unroll 1
  for (int c0 = 32 - kBatch; c0 >= 0; c0 -= kBatch) {
    float raw[kBatch][VEC];
#pragma unroll
    for (int u = kBatch - 1; u >= 0; --u) {
      const int cc = c0 + u;
      if (FULL || cc < cn) {
        if (VEC == 4) {
          const float4 t = ldg_stream4(reinterpret_cast<const float4*>(xp + (long long)cc * cstride));
          raw[u][0] = t.x; raw[u][1 % VEC] = t.y; raw[u][2 % VEC] = t.z; raw[u][3 % VEC] = t.w;
        } else {
          raw[u][0] = __ldg(xp + (long long)cc * cstride);
        }
      }
    }
#pragma unroll
    for (int u = kBatch - 1; u >= 0; --u) {
      const int cc = c0 + u;
      const bool have = FULL || cc < cn;
      const float2 k = have ? ab[cc] : make_float2(0.0f, 0.0f);
#pragma unroll
      for (int p = 0; p < VEC; ++p) {
        if (!have) {   // channel beyond C: bit 0 in every plane, nothing added to the residual sum
#pragma unroll
          for (int j = 0; j < NPL; ++j) word[p][j] = __funnelshift_l(0x80000000u, word[p][j], 1);
          continue;
        }
        const float val = clamp_sym(fmaf(raw[u][p], k.x, k.y), alpha);
        float acc = 0.0f, res = val;
#pragma unroll
        for (int j = 0; j < NPL; ++j) {
          const float d = (j == 0) ? val : __fsub_rn(val, acc);
          word[p][j] = __funnelshift_l(__float_as_uint(d), word[p][j], 1);
          if (j < ns) {
            const float t = __uint_as_float(__float_as_uint(sc[j]) ^ (__float_as_uint(d) & 0x80000000u));
            acc = (j == 0) ? t : __fadd_rn(acc, t);
            res = __fsub_rn(res, __uint_as_float(__float_as_uint(sc[j]) ^ (__float_as_uint(res) & 0x80000000u)));
          }
        }
        gsum[p] += fabsf(res);
      }
    }
  }
#pragma unroll
  for (int p = 0; p < VEC; ++p)
#pragma unroll
    for (int j = 0; j < NPL; ++j) word[p][j] = ~word[p][j];
}

// Both sign planes of the items [it_lo, it_hi) of sample `s` (item = 32-channel group x VEC-pixel group of the row xr)
// for the 2-bit / ternary code with first scale v1, written in the convolution's raster; returns this thread's share of
// sum |x - v1 sign(x)| (the v2 numerator).  Work is strided over the CTA's threads.  Shared by the fused quantizer
// (lsq_qact.cu) and the generic fallback for the rows it marks (lsq_solve.cu).
template <int VEC>
__device__ __forceinline__ double encode2_row_part(const float* __restrict__ xr, const ActGeom& g, int s, uint32_t hw,
                                                   uint32_t nq, uint32_t it_lo, uint32_t it_hi,
                                                   const float2* __restrict__ ab, float alpha, float v1,
                                                   uint32_t* __restrict__ planes) {
  const float sc[2] = {v1, 0.0f};
  double acc_sum = 0.0;
  for (uint32_t item = it_lo + threadIdx.x; item < it_hi; item += blockDim.x) {
    const int cgi = (int)(item / nq), q = (int)(item - (uint32_t)cgi * nq);
    const int p0 = q * VEC;
    const int cbase = cgi * 32;
    const int cn = min(32, g.c - cbase);
    const float* xp = xr + (long long)cbase * hw + p0;
    uint32_t word[VEC][2];
    float gsum[VEC];
    if (cn == 32) encode_group<2, VEC, true>(xp, (long long)hw, cn, ab + cbase, alpha, sc, 1, word, gsum);
    else encode_group<2, VEC, false>(xp, (long long)hw, cn, ab + cbase, alpha, sc, 1, word, gsum);
    int yi = p0 / g.w, xi = p0 - yi * g.w;
#pragma unroll
    for (int p = 0; p < VEC; ++p) {
      if (p0 + p < (int)hw) {
        int phase = 0, a = yi, b = xi;
        if (g.nphase == 4) { phase = ((yi & 1) << 1) | (xi & 1); a = yi >> 1; b = xi >> 1; }
        const long long v = vpos(g, s, a, b);
#pragma unroll
        for (int j = 0; j < 2; ++j)
          planes[(((long long)j * g.nphase + phase) * g.vtot + v) * g.cw + cgi] = word[p][j];
        acc_sum += (double)gsum[p];
      }
      if (++xi == g.w) { xi = 0; ++yi; }
    }
  }
  return acc_sum;
}

}  // namespace lsq
