// Binary convolution, CUDA-core variant, and the weight packer shared with the tensor-core variant.
//
// QuantConv2d.forward (quant/binary/binary_conv.py:161-173) with ls-1 weights and a k-plane
// activation code is   y = vw[c] * sum_j s_j[n] * I_j + bias[c],   I_j = conv(plane_j, sign(W)),
// an exact integer.  Zero padding contributes 0, so a tap outside the
// image is skipped rather than counted as -1.
#include "lsq_common.cuh"

namespace lsq {

int bconv2d_tc_launch(const uint32_t* d_planes, const lsq_act_geom* g, int nplanes, const float* d_act_scales,
                      const void* d_wpack, const float* d_w_scale, const float* d_bias, int cout, float* d_y,
                      const Epilogue& epi, cudaStream_t stream);
bool bconv2d_tc_supported(const lsq_act_geom* g, int nplanes, int cout);

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// --- weight image layout ---------------------------------------------------------------------
// bits : uint32[cout][kh*kw][cw]
// i8   : int8 [ceil(cout/128)][cin/64][kh*kw][4][128][16]   (only when cout % 64 == 0 and cin % 64 == 0)
//        one (channel tile, 64-channel block, tap) slab is the K-major, non-swizzled tcgen05 shared-memory
//        operand (8-row x 16-byte core matrices, LBO = 128*16, SBO = 128) with M = 128 rows, so the
//        tensor-core kernel fetches it with a single bulk copy.  A 64-channel layer stores channel r in rows
//        r and r + 64: both halves of the accumulator then hold the same channels and all eight epilogue
//        warps share the work.
__host__ __device__ inline int wpack_rows(int cout) { return cout < 128 ? 128 : cout; }
__host__ __device__ inline bool wpack_has_i8(int cout, int cin) { return (cout % 64 == 0) && (cin % 64 == 0) && (cout <= 128 || cout % 128 == 0); }
static inline size_t wpack_bits_bytes(int cout, int cin, int kh, int kw) {
  return align_up((size_t)cout * kh * kw * ((cin + 31) / 32) * 4, 1024);
}

__global__ void pack_weights_kernel(const float* __restrict__ w, int cout, int cin, int taps, int cw,
                                    uint32_t* __restrict__ bits, int8_t* __restrict__ i8) {
  // one thread per (co, tap, word)
  const long long total = (long long)cout * taps * cw;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int wd = (int)(idx % cw);
    const int tap = (int)((idx / cw) % taps);
    const int co = (int)(idx / ((long long)cw * taps));
    uint32_t word = 0u;
    for (int cc = 0; cc < 32; ++cc) {
      const int c = wd * 32 + cc;
      if (c >= cin) break;
      const float v = w[((long long)co * cin + c) * taps + tap];
      const bool pos = v >= 0.0f;
      word |= (pos ? 1u : 0u) << cc;
      if (i8) {
        const int ctile = co >> 7, r = co & 127, cb = c >> 6, j = (c & 63) >> 4, byte = c & 15;
        const long long off = ((((long long)ctile * (cin >> 6) + cb) * taps + tap) * 4 + j) * (128ll * 16) + (long long)r * 16 + byte;
        i8[off] = pos ? (int8_t)1 : (int8_t)-1;
        if (cout < 128) i8[off + (long long)cout * 16] = pos ? (int8_t)1 : (int8_t)-1;
      }
    }
    bits[idx] = word;
  }
}

// Inverse of the `bits` image: w[co][c][tap] = bit ? +scale[co] : -scale[co]  (scale == NULL: +-1).
__global__ void unpack_weights_kernel(const uint32_t* __restrict__ bits, const float* __restrict__ scale, int cout,
                                      int cin, int taps, int cw, float* __restrict__ w) {
  const long long total = (long long)cout * cin * taps;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int tap = (int)(idx % taps);
    const int c = (int)((idx / taps) % cin);
    const int co = (int)(idx / ((long long)taps * cin));
    const uint32_t word = __ldg(bits + ((long long)co * taps + tap) * cw + (c >> 5));
    float v = scale ? __ldg(scale + co) : 1.0f;
    if (!(v > 0.0f)) v = 1.0f;              // an unset scale must not erase the sign (-0.0 >= 0)
    w[idx] = ((word >> (c & 31)) & 1u) ? v : -v;
  }
}

// One thread per output element; all planes at once.  Used for shapes outside the tensor-core
// kernel (few channels, 5x5 LeNet layer) and as the on-device cross-check of that kernel.
template <int NPL>
__global__ void __launch_bounds__(256)
bconv_simple_kernel(const uint32_t* __restrict__ planes, ActGeom g, const float* __restrict__ act_scales,
                    const uint32_t* __restrict__ wbits, const float* __restrict__ w_scale,
                    const float* __restrict__ bias, int cout, float* __restrict__ y, Epilogue epi) {
  const long long total = (long long)g.n * cout * g.ho * g.wo;
  const int taps = g.kh * g.kw;
  const uint32_t tail_mask = (g.c & 31) ? ((1u << (g.c & 31)) - 1u) : 0xFFFFFFFFu;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int xo = (int)(idx % g.wo);
    const int yo = (int)((idx / g.wo) % g.ho);
    const int co = (int)((idx / ((long long)g.wo * g.ho)) % cout);
    const int s = (int)(idx / ((long long)g.wo * g.ho * cout));
    int acc[NPL];
#pragma unroll
    for (int j = 0; j < NPL; ++j) acc[j] = 0;
    for (int dy = 0; dy < g.kh; ++dy) {
      const int yi = yo * g.stride + dy - g.pad;
      if (yi < 0 || yi >= g.h) continue;
      for (int dx = 0; dx < g.kw; ++dx) {
        const int xi = xo * g.stride + dx - g.pad;
        if (xi < 0 || xi >= g.w) continue;
        int phase = 0, a = yi, b = xi;
        if (g.nphase == 4) { phase = ((yi & 1) << 1) | (xi & 1); a = yi >> 1; b = xi >> 1; }
        const long long v = vpos(g, s, a, b);
        const uint32_t* wp = wbits + ((long long)co * taps + dy * g.kw + dx) * g.cw;
        for (int wd = 0; wd < g.cw; ++wd) {
          const uint32_t m = (wd == g.cw - 1) ? tail_mask : 0xFFFFFFFFu;
          const uint32_t ww = __ldg(wp + wd);
          const int nvalid = __popc(m);
#pragma unroll
          for (int j = 0; j < NPL; ++j) {
            const uint32_t av = __ldg(planes + (((long long)j * g.nphase + phase) * g.vtot + v) * g.cw + wd);
            acc[j] += nvalid - 2 * __popc((av ^ ww) & m);
          }
        }
      }
    }
    float t = 0.0f;
#pragma unroll
    for (int j = 0; j < NPL; ++j) t = fmaf(__ldg(act_scales + (long long)j * g.n + s), (float)acc[j], t);
    float r = __ldg(w_scale + co) * t;
    if (bias) r += __ldg(bias + co);
    y[idx] = apply_epilogue(epi, r, co, idx);
  }
}

}  // namespace lsq

using namespace lsq;

extern "C" {

size_t lsq_wpack_bytes(int cout, int cin, int kh, int kw) {
  if (cout <= 0 || cin <= 0 || kh <= 0 || kw <= 0) return 0;
  size_t b = wpack_bits_bytes(cout, cin, kh, kw);
  if (wpack_has_i8(cout, cin)) b += align_up((size_t)wpack_rows(cout) * cin * kh * kw, 1024);
  return b;
}

int lsq_pack_weights(const float* d_w, int cout, int cin, int kh, int kw, void* d_wpack, void* stream) {
  LSQ_CHECK_ARG(d_w && d_wpack, "lsq_pack_weights: null pointer");
  LSQ_CHECK_ARG(cout > 0 && cin > 0 && kh > 0 && kw > 0, "lsq_pack_weights: bad shape");
  LSQ_CHECK_ARG(((uintptr_t)d_wpack & 1023) == 0, "lsq_pack_weights: d_wpack must be 1024-byte aligned");
  const int cw = (cin + 31) / 32, taps = kh * kw;
  uint32_t* bits = (uint32_t*)d_wpack;
  int8_t* i8 = wpack_has_i8(cout, cin) ? (int8_t*)((char*)d_wpack + wpack_bits_bytes(cout, cin, kh, kw)) : nullptr;
  const long long total = (long long)cout * taps * cw;
  unsigned grid = (unsigned)((total + 255) / 256);
  if (grid > 148 * 8) grid = 148 * 8;
  pack_weights_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_w, cout, cin, taps, cw, bits, i8);
  LSQ_CUDA_LAUNCH_CHECK("pack_weights_kernel");
  return LSQ_OK;
}

size_t lsq_wbits_bytes(int cout, int cin, int kh, int kw) {
  if (cout <= 0 || cin <= 0 || kh <= 0 || kw <= 0) return 0;
  return (size_t)cout * kh * kw * ((cin + 31) / 32) * 4;
}

int lsq_unpack_weights(const uint32_t* d_bits, const float* d_scale, int cout, int cin, int kh, int kw, float* d_w,
                       void* stream) {
  LSQ_CHECK_ARG(d_bits && d_w, "lsq_unpack_weights: null pointer");
  LSQ_CHECK_ARG(cout > 0 && cin > 0 && kh > 0 && kw > 0, "lsq_unpack_weights: bad shape");
  const long long total = (long long)cout * cin * kh * kw;
  unsigned grid = (unsigned)((total + 255) / 256);
  if (grid > 148 * 16) grid = 148 * 16;
  unpack_weights_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_bits, d_scale, cout, cin, kh * kw, (cin + 31) / 32, d_w);
  LSQ_CUDA_LAUNCH_CHECK("unpack_weights_kernel");
  return LSQ_OK;
}

int lsq_bconv2d_tc_supported(const lsq_act_geom* g, int nplanes, int cout) {
  return (g && bconv2d_tc_supported(g, nplanes, cout)) ? 1 : 0;
}

int lsq_bconv2d_fwd(const uint32_t* d_planes, const lsq_act_geom* g, int nplanes, const float* d_act_scales,
                    const void* d_wpack, const float* d_w_scale, const float* d_bias, int cout, float* d_y,
                    int impl, void* stream) {
  return lsq_bconv2d_fwd_ex(d_planes, g, nplanes, d_act_scales, d_wpack, d_w_scale, d_bias, cout, d_y, impl, nullptr, stream);
}

int lsq_bconv2d_fwd_ex(const uint32_t* d_planes, const lsq_act_geom* g, int nplanes, const float* d_act_scales,
                       const void* d_wpack, const float* d_w_scale, const float* d_bias, int cout, float* d_y,
                       int impl, const lsq_epilogue* epi, void* stream) {
  LSQ_CHECK_ARG(d_planes && g && d_act_scales && d_wpack && d_w_scale && d_y, "lsq_bconv2d_fwd: null pointer");
  LSQ_CHECK_ARG(nplanes >= 1 && nplanes <= 4 && cout > 0, "lsq_bconv2d_fwd: bad nplanes/cout");
  LSQ_CHECK_ARG(impl >= 0 && impl <= 2, "lsq_bconv2d_fwd: bad impl %d", impl);
  cudaStream_t st = (cudaStream_t)stream;
  const Epilogue de = to_dev(epi);
  LSQ_CHECK_ARG(de.act >= 0 && de.act <= 2 && (de.act != 2 || (de.prelu && (de.n_prelu == 1 || de.n_prelu == cout))), "lsq_bconv2d_fwd: bad epilogue");
  const bool tc_ok = bconv2d_tc_supported(g, nplanes, cout);
  if (impl == 2 && !tc_ok) {
    set_error("lsq_bconv2d_fwd: shape not supported by the tensor-core kernel (cin=%d cout=%d k=%dx%d)", g->c, cout, g->kh, g->kw);
    return LSQ_ERR_UNSUPPORTED;
  }
  if (impl == 2 || (impl == 0 && tc_ok))
    return bconv2d_tc_launch(d_planes, g, nplanes, d_act_scales, d_wpack, d_w_scale, d_bias, cout, d_y, de, st);
  ActGeom dg = to_dev(*g);
  const long long total = (long long)g->n * cout * g->ho * g->wo;
  unsigned grid = (unsigned)((total + 255) / 256);
  if (grid > 148u * 32u) grid = 148u * 32u;
  const uint32_t* wbits = (const uint32_t*)d_wpack;
  switch (nplanes) {
    case 1: bconv_simple_kernel<1><<<grid, 256, 0, st>>>(d_planes, dg, d_act_scales, wbits, d_w_scale, d_bias, cout, d_y, de); break;
    case 2: bconv_simple_kernel<2><<<grid, 256, 0, st>>>(d_planes, dg, d_act_scales, wbits, d_w_scale, d_bias, cout, d_y, de); break;
    case 3: bconv_simple_kernel<3><<<grid, 256, 0, st>>>(d_planes, dg, d_act_scales, wbits, d_w_scale, d_bias, cout, d_y, de); break;
    default: bconv_simple_kernel<4><<<grid, 256, 0, st>>>(d_planes, dg, d_act_scales, wbits, d_w_scale, d_bias, cout, d_y, de); break;
  }
  LSQ_CUDA_LAUNCH_CHECK("bconv_simple_kernel");
  return LSQ_OK;
}

}  // extern "C"
