// Full-precision pointwise (1x1) strided convolution on the 5th-generation tensor cores: the downsampling
// shortcut of the residual blocks (quant/models/resnet.py:24-39: Conv2d(kernel_size=1, stride=2) + BatchNorm2d,
// BatchNorm folded into weights / bias by the caller).  Outside the quantized path proper; it is
// here because cuDNN / cuBLAS ran it at 1.0-1.2 ms of a 10 ms forward step (layout transposes or SIMT fp32 GEMMs).
//
//   y[n, co, oy, ox] = sum_ci w[co, ci] * x[n, ci, stride*oy, stride*ox] + b[co]
// as a GEMM  D[128 out-channels, 256 output positions] += W[128, K] * X[K, 256]  with tcgen05.mma kind::tf32 and
// the 3xTF32 split (x = hi + lo: lo*hi + hi*lo + hi*hi, fp32-level accuracy).  K runs over the input channels in
// stages of 16: the producers gather the strided pixels of 16 channels for 256 positions into K-major chunks
// ([4 channels = 16 bytes] per position, hi and lo images), the weights of a stage arrive as one bulk copy of a
// pre-packed slab.  Same warp roles and epilogue as lsq_bconv_tc.cu.
#include "lsq_common.cuh"
#include "lsq_tc.cuh"

namespace lsq {

constexpr int kPwThreads = 576;        // warps 0-3, 10-13 epilogue; 4 MMA; 5 weights; 6-9, 14-17 producers
constexpr int kPwTile = 256;          // output positions per tile = N
constexpr int kPwStages = 3;          // patch / weight ring depth (per 16 input channels)
constexpr int kPwOutPitch = 36;
constexpr uint32_t kPwPatchBytes = 2u * 4u * kPwTile * 16u;   // [hi, lo][chunk 4][256 positions][16 B] = 32 KB
constexpr uint32_t kPwSlabBytes = 2u * 4u * 128u * 16u;       // [hi, lo][chunk 4][128 rows][16 B]     = 16 KB

struct PwParams {
  int n, cin, h, w, stride, cout, ho, wo;
  int p_tiles, n_ctiles, kstages;
  long long positions;                 // n * ho * wo
  unsigned long long pos_magic;        // ceil(2^40 / (ho*wo))
  unsigned long long wo_magic;         // ceil(2^40 / wo)
  uint32_t smem_p, smem_w, smem_bar, smem_out;
};

// image[ctile][kstage][hl][chunk][row][k]: w[ctile*128 + row][kstage*16 + chunk*4 + k] as TF32 hi / fp32 lo
__global__ void pw_pack_kernel(const float* __restrict__ w, int cout, int cin, float* __restrict__ image) {
  const long long total = (long long)cout * cin * 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i & 3), row = (int)((i >> 2) & 127), chunk = (int)((i >> 9) & 3), hl = (int)((i >> 11) & 1);
    const long long st = i >> 12;                       // ctile * kstages + kstage
    const int kstages = cin >> 4;
    const int ctile = (int)(st / kstages), ks = (int)(st - (long long)ctile * kstages);
    const float v = w[(long long)(ctile * 128 + row) * cin + ks * 16 + chunk * 4 + k];
    const float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    image[i] = hl == 0 ? hi : __fsub_rn(v, hi);
  }
}

__global__ void __launch_bounds__(kPwThreads, 1)
pwconv_kernel(const float* __restrict__ x, PwParams P, const float* __restrict__ wimage, const float* __restrict__ bias,
              float* __restrict__ y) {
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // barriers: p_full[3] p_empty[3] w_full[3] w_empty[3] acc_full[2] acc_empty[2] | tmem base
  const uint32_t bar0 = sbase + P.smem_bar;
  auto p_full = [&](int s) { return bar0 + 8u * s; };
  auto p_empty = [&](int s) { return bar0 + 8u * (kPwStages + s); };
  auto w_full = [&](int s) { return bar0 + 8u * (2 * kPwStages + s); };
  auto w_empty = [&](int s) { return bar0 + 8u * (3 * kPwStages + s); };
  auto acc_full = [&](int s) { return bar0 + 8u * (4 * kPwStages + s); };
  auto acc_empty = [&](int s) { return bar0 + 8u * (4 * kPwStages + 2 + s); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + P.smem_bar + 8u * (4 * kPwStages + 4));
  if (threadIdx.x == 0) {
    for (int s = 0; s < kPwStages; ++s) {
      mbar_init(p_full(s), 8); mbar_init(p_empty(s), 1); mbar_init(w_full(s), 1); mbar_init(w_empty(s), 1);
    }
    for (int s = 0; s < 2; ++s) { mbar_init(acc_full(s), 1); mbar_init(acc_empty(s), 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) tmem_alloc(smem_u32((const void*)tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  int* const err = nullptr;
  const int n_items = P.p_tiles * P.n_ctiles;
  const int hwo = P.ho * P.wo;

  if (warp < 4 || (warp >= 10 && warp < 14)) {
    // ===================== epilogue (8 warps): + bias, transposed, one 128-byte row segment per store ========
    const int quarter = warp & 3, half = warp < 4 ? 0 : 1;
    const int ewarp = quarter + 4 * half;
    float* const outt = reinterpret_cast<float*>(smem + P.smem_out) + (size_t)ewarp * 32 * kPwOutPitch;
    Ring acc(2);
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int ptile = item / P.n_ctiles, ctile = item - ptile * P.n_ctiles;
      const float bs = __ldg(bias + ctile * 128 + 32 * quarter + lane);
      mbar_wait(acc_full(acc.stage), acc.phase, err, 1);
      tc_fence_after();
      for (int st = 0; st < 4; ++st) {
        const int p0 = half * 128 + 32 * st;
        float4* orow = reinterpret_cast<float4*>(outt + lane * kPwOutPitch);
#pragma unroll
        for (int sub = 0; sub < 2; ++sub) {
          uint32_t rr[16];
          tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc.stage * kPwTile + p0 + 16 * sub), rr);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 4; ++j)
            orow[4 * sub + j] = make_float4(__uint_as_float(rr[4 * j]) + bs, __uint_as_float(rr[4 * j + 1]) + bs,
                                            __uint_as_float(rr[4 * j + 2]) + bs, __uint_as_float(rr[4 * j + 3]) + bs);
        }
        __syncwarp();
        const long long p = (long long)ptile * kPwTile + p0 + lane;
        if (p < P.positions) {
          const unsigned s = (unsigned)(((unsigned long long)p * P.pos_magic) >> 40);
          const int rem = (int)(p - (long long)s * hwo);
          float* yp = y + ((long long)s * P.cout + ctile * 128 + 32 * quarter) * hwo + rem;
          const float* ot = outt + lane;
#pragma unroll
          for (int c0 = 0; c0 < 32; c0 += 16) {
            float v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = ot[(c0 + i) * kPwOutPitch];
#pragma unroll
            for (int i = 0; i < 16; ++i) { *yp = v[i]; yp += hwo; }
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty(acc.stage));
      acc.advance();
    }
  } else if (warp == 4) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      Ring acc(2), rp(kPwStages), rw(kPwStages);
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kPwTile >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        mbar_wait(acc_empty(acc.stage), acc.phase ^ 1u, err, 2);
        tc_fence_after();
        const uint32_t d0 = tmem_base + (uint32_t)(acc.stage * kPwTile);
        for (int ks = 0; ks < P.kstages; ++ks) {
          mbar_wait(p_full(rp.stage), rp.phase, err, 3);
          mbar_wait(w_full(rw.stage), rw.phase, err, 4);
          tc_fence_after();
          const uint32_t pst = sbase + P.smem_p + (uint32_t)rp.stage * kPwPatchBytes;
          const uint32_t wst = sbase + P.smem_w + (uint32_t)rw.stage * kPwSlabBytes;
#pragma unroll
          for (int term = 0; term < 3; ++term) {          // small terms first: W_lo X_hi, W_hi X_lo, W_hi X_hi
            const uint32_t wa = wst + (term == 0 ? kPwSlabBytes / 2 : 0u);
            const uint32_t pb = pst + (term == 1 ? kPwPatchBytes / 2 : 0u);
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
              const uint64_t ad = make_desc(wa + (uint32_t)kk * 2u * 2048u, 2048u, 128u);
              const uint64_t bd = make_desc(pb + (uint32_t)kk * 2u * 4096u, 4096u, 128u);
              umma_tf32(d0, ad, bd, idesc, (ks | term | kk) != 0 ? 1u : 0u);
            }
          }
          umma_commit(p_empty(rp.stage));
          umma_commit(w_empty(rw.stage));
          rp.advance(); rw.advance();
        }
        umma_commit(acc_full(acc.stage));
        acc.advance();
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    // ===================== weight loader =====================
    if (lane == 0) {
      Ring rw(kPwStages);
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int ctile = item % P.n_ctiles;
        const float* wsrc = wimage + (size_t)ctile * P.kstages * (kPwSlabBytes / 4);
        for (int ks = 0; ks < P.kstages; ++ks) {
          mbar_wait(w_empty(rw.stage), rw.phase ^ 1u, err, 5);
          mbar_expect_tx(w_full(rw.stage), kPwSlabBytes);
          bulk_g2s(sbase + P.smem_w + (uint32_t)rw.stage * kPwSlabBytes, wsrc + (size_t)ks * (kPwSlabBytes / 4), kPwSlabBytes, w_full(rw.stage));
          rw.advance();
        }
      }
    }
  } else {
    // ===================== producers (256 threads): strided gather of 16 channels x 256 positions ==========
    // thread = position.  The 16 loads of the NEXT stage are issued before the current one is converted and stored
    // (two register buffers): the gather's latency, not its volume, is what a stage waits for.
    Ring rp(kPwStages);
    const int pt = warp < 10 ? threadIdx.x - 6 * 32 : threadIdx.x - 14 * 32 + 128;       // warps 6-9 and 14-17
    const long long plane = (long long)P.h * P.w;
    int item = blockIdx.x, ks = 0;
    const float* src = x;
    bool ok = false;
    auto locate = [&]() {                               // this thread's pixel of channel 0 for `item`
      const long long p = (long long)(item / P.n_ctiles) * kPwTile + pt;
      ok = item < n_items && p < P.positions;
      const long long pc = ok ? p : 0;
      const unsigned sm = (unsigned)(((unsigned long long)pc * P.pos_magic) >> 40);
      const unsigned rem = (unsigned)(pc - (long long)sm * hwo);
      const unsigned oy = (unsigned)(((unsigned long long)rem * P.wo_magic) >> 40);
      const unsigned ox = rem - oy * (unsigned)P.wo;
      src = x + ((long long)sm * P.cin * P.h + (long long)P.stride * oy) * P.w + (long long)P.stride * ox;
    };
    auto issue = [&](float (&v)[16]) {
#pragma unroll
      for (int c = 0; c < 16; ++c) v[c] = ok ? __ldg(src + (long long)(ks * 16 + c) * plane) : 0.0f;
    };
    auto next = [&]() {
      if (++ks == P.kstages) { ks = 0; item += gridDim.x; locate(); }
    };
    auto emit = [&](const float (&v)[16]) {
      mbar_wait(p_empty(rp.stage), rp.phase ^ 1u, err, 6);
      float4* hi = reinterpret_cast<float4*>(smem + P.smem_p + (size_t)rp.stage * kPwPatchBytes);
      float4* lo = hi + kPwPatchBytes / 32;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float h4[4], l4[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          h4[k] = __uint_as_float(__float_as_uint(v[4 * j + k]) & 0xFFFFE000u);
          l4[k] = __fsub_rn(v[4 * j + k], h4[k]);
        }
        hi[j * kPwTile + pt] = make_float4(h4[0], h4[1], h4[2], h4[3]);
        lo[j * kPwTile + pt] = make_float4(l4[0], l4[1], l4[2], l4[3]);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full(rp.stage));
      rp.advance();
    };
    float va[16], vb[16];
    locate();
    if (item < n_items) issue(va);
    while (item < n_items) {
      next();
      if (item < n_items) issue(vb);
      emit(va);
      if (item >= n_items) break;
      next();
      if (item < n_items) issue(va);
      emit(vb);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, 512);
}

}  // namespace lsq

using namespace lsq;

extern "C" int lsq_pwconv_supported(int cin, int cout) { return (cin > 0 && cout > 0 && cin % 16 == 0 && cout % 128 == 0) ? 1 : 0; }

extern "C" size_t lsq_pwconv_image_bytes(int cout, int cin) {
  return lsq_pwconv_supported(cin, cout) ? (size_t)cout * cin * 2 * sizeof(float) : 0;
}

extern "C" int lsq_pwconv_pack_weights(const float* d_w, int cout, int cin, float* d_image, void* stream) {
  LSQ_CHECK_ARG(d_w && d_image, "lsq_pwconv_pack_weights: null pointer");
  LSQ_CHECK_ARG(lsq_pwconv_supported(cin, cout), "lsq_pwconv_pack_weights: needs cin %% 16 == 0 and cout %% 128 == 0 (got %d, %d)", cin, cout);
  const long long total = (long long)cout * cin * 2;
  unsigned grid = (unsigned)((total + 255) / 256);
  if (grid > 148u * 8u) grid = 148u * 8u;
  pw_pack_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_w, cout, cin, d_image);
  LSQ_CUDA_LAUNCH_CHECK("pw_pack_kernel");
  return LSQ_OK;
}

extern "C" int lsq_pwconv_fwd(const float* d_x, int n, int cin, int h, int w, int stride, const float* d_image,
                              const float* d_bias, int cout, float* d_y, void* stream) {
  LSQ_CHECK_ARG(d_x && d_image && d_bias && d_y, "lsq_pwconv_fwd: null pointer");
  LSQ_CHECK_ARG(n > 0 && h > 0 && w > 0 && stride >= 1, "lsq_pwconv_fwd: bad shape");
  LSQ_CHECK_ARG(lsq_pwconv_supported(cin, cout), "lsq_pwconv_fwd: needs cin %% 16 == 0 and cout %% 128 == 0 (got %d, %d)", cin, cout);
  LSQ_CHECK_ARG(((uintptr_t)d_image & 15) == 0, "lsq_pwconv_fwd: weight image must be 16-byte aligned");
  PwParams P;
  P.n = n; P.cin = cin; P.h = h; P.w = w; P.stride = stride; P.cout = cout;
  P.ho = (h - 1) / stride + 1; P.wo = (w - 1) / stride + 1;
  P.positions = (long long)n * P.ho * P.wo;
  LSQ_CHECK_ARG(P.positions < (1ll << 31), "lsq_pwconv_fwd: too many positions");
  P.p_tiles = (int)((P.positions + kPwTile - 1) / kPwTile);
  P.n_ctiles = cout / 128; P.kstages = cin / 16;
  const unsigned long long hwo = (unsigned long long)P.ho * P.wo;
  P.pos_magic = ((1ull << 40) + hwo - 1ull) / hwo;
  P.wo_magic = ((1ull << 40) + (unsigned long long)P.wo - 1ull) / (unsigned long long)P.wo;
  uint32_t o = 0;
  P.smem_p = o; o += kPwStages * kPwPatchBytes;
  P.smem_w = o; o += kPwStages * kPwSlabBytes;
  P.smem_bar = o; o += 256;
  P.smem_out = o; o += 8 * 32 * kPwOutPitch * 4;
  const size_t smem = o;
  static std::atomic<unsigned long long> smem_set{0ull};
  const cudaError_t e = ensure_max_smem(pwconv_kernel, smem_set);
  if (e != cudaSuccess) { set_error("lsq_pwconv_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return LSQ_ERR_CUDA; }
  const int sms = device_sms();
  const int n_items = P.p_tiles * P.n_ctiles;
  const int grid = n_items < sms ? n_items : sms;
  pwconv_kernel<<<grid, kPwThreads, smem, (cudaStream_t)stream>>>(d_x, P, d_image, d_bias, d_y);
  LSQ_CUDA_LAUNCH_CHECK("pwconv_kernel");
  return LSQ_OK;
}
