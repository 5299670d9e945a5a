// Fused ImageNet stem of QResNet: conv 7x7 / stride 2 / pad 3 (3 -> 64 channels, eval BatchNorm folded
// into weights and bias) -> max-pool 3x3 / stride 2 / pad 1 -> ReLU, fp32 in, fp32 out.
// (quant/models/resnet.py:283-308: blocks[0] = Sequential(conv1, bn1, ReLU, maxpool); ReLU and max-pool
//  commute.)  The 1.6 GB conv output of a 512-image batch never reaches HBM: a CTA computes the conv
// outputs feeding a 2 x 14 tile of pooled pixels in shared memory and writes only the pooled tile.
//
// Arithmetic: implicit GEMM (M = conv pixels, N = 64, K = 147 padded to 152) on the tensor cores with
// mma.sync m16n8k8 TF32 and the 3xTF32 split (x = hi + lo, x*w ~ hi*hi + hi*lo + lo*hi), which restores
// fp32-level accuracy (~1e-6 relative) while the reference's own GPU path (cuDNN, TF32 allowed by
// default) is at ~1e-3.  This layer is outside the quantized path proper; it is fused
// because after the quantized layers were fused it was the largest item of the forward step.
#include "lsq_common.cuh"

namespace lsq {

constexpr int kStemThreads = 160;   // 5 warps x 2 m16 tiles = 160 conv-pixel rows (145 used)
constexpr int kPH = 2, kPW = 14;    // pooled tile
constexpr int kCR = 2 * kPH + 1;    // conv rows per tile (5)
constexpr int kCC = 2 * kPW + 1;    // conv cols per tile (29)
constexpr int kPR = 4 * kPH + 7;    // input rows (15)
constexpr int kPC = 4 * kPW + 7;    // input cols (63)
constexpr int kPPitch = 64;         // patch row pitch (floats)
constexpr int kK = 152;             // 3*7*7 = 147 padded to a multiple of 8
constexpr int kWPitch = 156;        // weight row pitch (floats): conflict-free B fragments
constexpr int kCPitch = 66;         // conv staging pitch (floats per pixel)

constexpr int kPatchElems = 3 * kPR * kPPitch;
struct StemSmem {                       // 102 KB: two CTAs per SM, their load / GEMM / pool phases interleave
  float w[64 * kWPitch];                // folded weights, resident: 39.9 KB
  float patch[2][kPatchElems];          // double-buffered input patch: 2 x 11.5 KB
  float conv[kCR * kCC * kCPitch];      // 38.3 KB
  int koff[kK];
  float bias[64];
};

// 4-byte async copy global -> shared, zero-filled when !valid (cp.async with src-size 0)
__device__ __forceinline__ void cp_async_f32(float* dst, const float* src, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  const int sz = valid ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}

__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// x = hi + lo with hi the TF32 truncation of x (the tensor core ignores the low 13 mantissa bits)
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x) & 0xFFFFE000u;
  lo = __float_as_uint(__fsub_rn(x, __uint_as_float(hi)));
}

// Persistent: two CTAs per SM keep the weights in shared memory and walk over the tiles; the input patch
// of the next tile is fetched with cp.async while the tensor cores work on the current one.
__global__ void __launch_bounds__(kStemThreads, 2)
stem_kernel(const float* __restrict__ x, const float* __restrict__ wg, const float* __restrict__ bias,
            float* __restrict__ out, int n, int h, int w, int hc, int wc, int hp, int wp, int tiles_x, int tiles_y,
            int n_tiles) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  StemSmem& sm = *reinterpret_cast<StemSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  for (int i = tid; i < 64 * kK; i += kStemThreads) sm.w[(i / kK) * kWPitch + (i % kK)] = __ldg(wg + i);
  for (int i = tid; i < 64; i += kStemThreads) sm.bias[i] = __ldg(bias + i);
  for (int k = tid; k < kK; k += kStemThreads) {
    const int c = k / 49, r = k - c * 49, ky = r / 7, kx = r - ky * 7;
    sm.koff[k] = (k < 147) ? (c * kPR * kPPitch + ky * kPPitch + kx) : 0;
  }

  const int gid = lane >> 2, tig = lane & 3;
  int base[2][2];
#pragma unroll
  for (int t = 0; t < 2; ++t)
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      int m = (warp * 2 + t) * 16 + gid + hh * 8;
      if (m >= kCR * kCC) m = 0;                      // padding rows compute garbage that is never stored
      const int r = m / kCC, cx = m - r * kCC;
      base[t][hh] = 2 * r * kPPitch + 2 * cx;
    }

  auto prefetch_patch = [&](int tile, int buf) {
    const int tx = tile % tiles_x;
    const int ty = (tile / tiles_x) % tiles_y;
    const int s = tile / (tiles_x * tiles_y);
    const int iy0 = 4 * (ty * kPH) - 5, ix0 = 4 * (tx * kPW) - 5;    // first input row / col of the patch
    const float* xs = x + (long long)s * 3 * h * w;
    for (int i = tid; i < kPatchElems; i += kStemThreads) {
      const int col = i % kPPitch, row = (i / kPPitch) % kPR, c = i / (kPPitch * kPR);
      const int iy = iy0 + row, ix = ix0 + col;
      const bool ok = col < kPC && iy >= 0 && iy < h && ix >= 0 && ix < w;
      cp_async_f32(&sm.patch[buf][i], ok ? xs + ((long long)c * h + iy) * w + ix : xs, ok);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  int buf = 0;
  if ((int)blockIdx.x < n_tiles) prefetch_patch(blockIdx.x, 0);
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, buf ^= 1) {
    const int tx = tile % tiles_x;
    const int ty = (tile / tiles_x) % tiles_y;
    const int s = tile / (tiles_x * tiles_y);
    const int py0 = ty * kPH, px0 = tx * kPW;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();          // patch[buf] visible; the previous tile's pooling is done with sm.conv
    if (tile + (int)gridDim.x < n_tiles) prefetch_patch(tile + gridDim.x, buf ^ 1);
    const float* patch = sm.patch[buf];

    // ---- implicit GEMM: each warp owns two m16 tiles (32 conv pixels), all 64 output channels --------
    float acc[2][8][4];
#pragma unroll
    for (int t = 0; t < 2; ++t)
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[t][nt][e] = 0.0f;

#pragma unroll 1
    for (int ks = 0; ks < kK / 8; ++ks) {
      const int k0 = ks * 8;
      const int o0 = sm.koff[k0 + tig], o1 = sm.koff[k0 + tig + 4];
      uint32_t ahi[2][4], alo[2][4];
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        split_tf32(patch[base[t][0] + o0], ahi[t][0], alo[t][0]);
        split_tf32(patch[base[t][1] + o0], ahi[t][1], alo[t][1]);
        split_tf32(patch[base[t][0] + o1], ahi[t][2], alo[t][2]);
        split_tf32(patch[base[t][1] + o1], ahi[t][3], alo[t][3]);
      }
      uint2 b0[8], b1[8];   // (hi, lo)
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        split_tf32(sm.w[(nt * 8 + gid) * kWPitch + k0 + tig], b0[nt].x, b0[nt].y);
        split_tf32(sm.w[(nt * 8 + gid) * kWPitch + k0 + tig + 4], b1[nt].x, b1[nt].y);
      }
      // three passes, small terms first; consecutive MMAs touch different accumulators
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int t = 0; t < 2; ++t) mma_tf32(acc[t][nt], alo[t], b0[nt].x, b1[nt].x);
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int t = 0; t < 2; ++t) mma_tf32(acc[t][nt], ahi[t], b0[nt].y, b1[nt].y);
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int t = 0; t < 2; ++t) mma_tf32(acc[t][nt], ahi[t], b0[nt].x, b1[nt].x);
    }

    // ---- conv + bias -> shared staging -------------------------------------------------------------
#pragma unroll
    for (int t = 0; t < 2; ++t)
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int m = (warp * 2 + t) * 16 + gid + hh * 8;
        if (m < kCR * kCC) {
#pragma unroll
          for (int nt = 0; nt < 8; ++nt) {
            const int ch = nt * 8 + 2 * tig;
            *reinterpret_cast<float2*>(&sm.conv[m * kCPitch + ch]) =
                make_float2(acc[t][nt][hh * 2 + 0] + sm.bias[ch], acc[t][nt][hh * 2 + 1] + sm.bias[ch + 1]);
          }
        }
      }
    __syncthreads();

    // ---- 3x3 / stride 2 / pad 1 max-pool + ReLU, pooled tile -> NCHW ----------------------------------
    float* os = out + (long long)s * 64 * hp * wp;
    for (int i = tid; i < 64 * kPH * kPW; i += kStemThreads) {
      const int pxl = i % kPW, pyl = (i / kPW) % kPH, ch = i / (kPW * kPH);
      const int py = py0 + pyl, px = px0 + pxl;
      if (py >= hp || px >= wp) continue;
      float m = -INFINITY;
#pragma unroll
      for (int dy = 0; dy < 3; ++dy) {
        const int cy = 2 * py - 1 + dy;
        if (cy < 0 || cy >= hc) continue;
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const int cxg = 2 * px - 1 + dx;
          if (cxg < 0 || cxg >= wc) continue;
          m = fmaxf(m, sm.conv[((2 * pyl + dy) * kCC + (2 * pxl + dx)) * kCPitch + ch]);
        }
      }
      os[((long long)ch * hp + py) * wp + px] = fmaxf(m, 0.0f);
    }
  }
}

}  // namespace lsq

using namespace lsq;

extern "C" int lsq_stem_fwd(const float* d_x, int n, int h, int w, const float* d_w, const float* d_bias,
                            float* d_out, void* stream) {
  LSQ_CHECK_ARG(d_x && d_w && d_bias && d_out, "lsq_stem_fwd: null pointer");
  LSQ_CHECK_ARG(n > 0 && h >= 7 && w >= 7, "lsq_stem_fwd: bad shape");
  const int hc = (h + 6 - 7) / 2 + 1, wc = (w + 6 - 7) / 2 + 1;
  const int hp = (hc + 2 - 3) / 2 + 1, wp = (wc + 2 - 3) / 2 + 1;
  const int tiles_x = (wp + kPW - 1) / kPW, tiles_y = (hp + kPH - 1) / kPH;
  const long long tiles = (long long)n * tiles_x * tiles_y;
  LSQ_CHECK_ARG(tiles < (1ll << 31), "lsq_stem_fwd: too many tiles");
  const size_t smem = sizeof(StemSmem);
  cudaError_t e = cudaFuncSetAttribute(stem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("lsq_stem_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return LSQ_ERR_CUDA; }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const unsigned grid = (unsigned)(tiles < 2 * sms ? tiles : 2 * sms);
  stem_kernel<<<grid, kStemThreads, smem, (cudaStream_t)stream>>>(d_x, d_w, d_bias, d_out, n, h, w, hc, wc, hp, wp,
                                                                  tiles_x, tiles_y, (int)tiles);
  LSQ_CUDA_LAUNCH_CHECK("stem_kernel");
  return LSQ_OK;
}
