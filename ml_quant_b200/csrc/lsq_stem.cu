// ImageNet stem of QResNet on the 5th-generation tensor cores: conv 7x7 / stride 2 / pad 3 (3 -> 64 channels,
// eval BatchNorm folded into weights and bias) + ReLU as an implicit GEMM with tcgen05.mma kind::tf32 and the
// 3xTF32 split (x = hi + lo, x*w ~ lo*hi + hi*lo + hi*hi: fp32-level accuracy, ~1e-6 relative), followed by a
// small max-pool kernel.  (quant/models/resnet.py:283-308: blocks[0] = Sequential(conv1, bn1, ReLU, maxpool).)
// This layer is outside the quantized path proper; it is here because after the quantized
// layers were fused it was the largest item of the forward step (legacy mma.sync TF32 runs at CUDA-core rate
// on sm_100a: the previous mma.sync kernel took 4.2 ms of a 13 ms step).
//
// Same construction as the binary convolution (lsq_bconv_tc.cu): the input is viewed as 4 stride phases in the
// "virtual raster" of lsq_act_geometry, in which a tap is a uniform shift of the position index, so a tile of
// 256 consecutive output positions needs ONE patch per phase in shared memory: [position][c0 c1 c2 0] fp32
// (16 bytes = one K chunk of 4), once as TF32 "hi" and once as the fp32 remainder "lo".
//     D[64 channels (M = 128, upper half don't-care), N = 256 positions] += W[., K = 8] * P[K = 8, N]
// One MMA contracts TWO taps: its two K chunks are the same patch at two shifts -- the descriptor's start
// address selects the first tap, its leading-dimension byte offset (LBO) the distance to the second.
// The 64 channels use only half of the M = 128 rows, so the other half carries the second weight term:
// rows 0..63 = W_hi, rows 64..127 = W_lo; with B = P_hi and then B = P_lo the accumulator holds
// W_hi (P_hi + P_lo) in lanes 0..63 and W_lo (P_hi + P_lo) in lanes 64..127 -- all four product terms in TWO MMAs
// per tap pair instead of three.  49 taps -> 25 pairs x 2 = 50 MMAs per tile; the epilogue adds the two lane
// halves (warp pairs exchange through shared memory).  All weights (100 KB) stay in shared memory.
#include <stdlib.h>
#include <cuda_fp16.h>
#include "lsq_common.cuh"
#include "lsq_tc.cuh"

namespace lsq {

constexpr int kStThreads = 800;      // 8 + 8 epilogue warps (lane halves), MMA warp 16, producer warps 17-24
constexpr int kStTile = 256;        // output positions per tile = N
constexpr int kStPairs = 25;
constexpr int kStPStages = 4;       // ring of per-phase patches
constexpr int kStOutPitch = 20;
constexpr int kStProducerWarps = 8;
constexpr uint32_t kStWeightBytes = kStPairs * 4096;       // [pair][chunk 2][128 rows: 64 hi, 64 lo][4 floats]

struct StemTaps {                   // static description of the 7x7 / stride 2 / pad 3 taps
  int tap[kStPairs][2];             // ky*7+kx of the two chunks of a pair, -1 = none (zero weights)
  int phase[kStPairs];
  int qy[kStPairs][2], qx[kStPairs][2];
  int first[4], count[4];           // pairs of a phase
};

static StemTaps stem_taps() {
  StemTaps T;
  auto fdiv = [](int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); };
  int np = 0;
  for (int phase = 0; phase < 4; ++phase) {
    T.first[phase] = np;
    int list[16][3], nl = 0;        // (tap, qy, qx) in ascending (qy, qx) = ascending offset
    for (int ky = 0; ky < 7; ++ky)
      for (int kx = 0; kx < 7; ++kx) {
        const int ey = ky - 3, ex = kx - 3;
        const int qy = fdiv(ey, 2), qx = fdiv(ex, 2);
        if ((ey - 2 * qy) * 2 + (ex - 2 * qx) != phase) continue;
        list[nl][0] = ky * 7 + kx; list[nl][1] = qy; list[nl][2] = qx; ++nl;
      }
    for (int i = 0; i < nl; i += 2, ++np) {
      T.phase[np] = phase;
      T.tap[np][0] = list[i][0]; T.qy[np][0] = list[i][1]; T.qx[np][0] = list[i][2];
      if (i + 1 < nl) { T.tap[np][1] = list[i + 1][0]; T.qy[np][1] = list[i + 1][1]; T.qx[np][1] = list[i + 1][2]; }
      else { T.tap[np][1] = -1; T.qy[np][1] = list[i][1]; T.qx[np][1] = list[i][2] + 1; }   // zero weights x the next position
    }
    T.count[phase] = np - T.first[phase];
  }
  return T;   // np == 25
}

struct StemParams {
  ActGeom g;
  int hc, wc, p_tiles, pp;
  long long q_begin;
  int dmin[4], first[4], count[4];
  int pair_off[kStPairs], pair_lbo[kStPairs];   // first-tap offset (relative to dmin of the phase) and distance to the second, in positions
  uint32_t phase_bytes, stage_bytes, smem_w, smem_p, smem_bar, smem_out;
  unsigned long long pitch_magic, rps_magic;
};

struct StPos { int s, a, col; bool in_range; };
__device__ __forceinline__ StPos st_decode(const StemParams& P, long long q) {
  StPos r;
  const long long rel = q - P.g.lead;
  r.in_range = rel >= 0;
  const unsigned long long urel = r.in_range ? (unsigned long long)rel : 0ull;
  const unsigned R = (unsigned)((urel * P.pitch_magic) >> 40);
  r.col = (int)(urel - (unsigned long long)R * (unsigned)P.g.pitch);
  r.s = (int)(((unsigned long long)R * P.rps_magic) >> 40);
  r.a = (int)R - r.s * P.g.rps - P.g.ph;
  return r;
}

// image[pair][chunk][row][k]: weights of tap(pair, chunk), channel row & 63, input channel k (k = 3: zero);
// rows 0..63 TF32 hi, rows 64..127 fp32 remainder lo
__global__ void stem_pack_kernel(const float* __restrict__ w, float* __restrict__ image, StemTaps T) {
  const int total = kStPairs * 2 * 2 * 64 * 4;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int k = i & 3, row = (i >> 2) & 63, hl = (i >> 8) & 1, chunk = (i >> 9) & 1, pair = i >> 10;
    const int tap = T.tap[pair][chunk];
    const float v = (k < 3 && tap >= 0) ? w[row * 147 + k * 49 + tap] : 0.0f;
    const float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    image[i] = hl == 0 ? hi : __fsub_rn(v, hi);
  }
}

__global__ void __launch_bounds__(kStThreads, 1)
stem_conv_kernel(const float* __restrict__ x, StemParams P, const float* __restrict__ wimage, const float* __restrict__ bias,
                 float* __restrict__ conv, long long* __restrict__ diag) {
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const ActGeom& g = P.g;
  long long w0 = 0, w1 = 0;
  const long long t_start = LSQ_TC_CLOCK();
  // barriers: p_full[4] p_empty[4] acc_full[2] acc_empty[2] | tmem base
  const uint32_t bar0 = sbase + P.smem_bar;
  auto p_full = [&](int s) { return bar0 + 8u * s; };
  auto p_empty = [&](int s) { return bar0 + 8u * (kStPStages + s); };
  auto acc_full = [&](int s) { return bar0 + 8u * (2 * kStPStages + s); };
  auto acc_empty = [&](int s) { return bar0 + 8u * (2 * kStPStages + 2 + s); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + P.smem_bar + 8u * (2 * kStPStages + 4));
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStPStages; ++s) { mbar_init(p_full(s), kStProducerWarps); mbar_init(p_empty(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(acc_full(s), 1); mbar_init(acc_empty(s), 16); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // resident weights
  {
    const float4* src = reinterpret_cast<const float4*>(wimage);
    float4* dst = reinterpret_cast<float4*>(smem + P.smem_w);
    for (int i = threadIdx.x; i < (int)(kStWeightBytes / 16); i += kStThreads) dst[i] = __ldg(src + i);
  }
  fence_proxy_async();
  if (warp == 16) tmem_alloc(smem_u32((const void*)tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  int* const err = nullptr;

  if (warp < 16) {
    // ===================== epilogue (16 warps) =====================
    // Warps with (warp & 3) < 2 read accumulator lanes 0..63 (the W_hi half), their partners warp + 2 lanes
    // 64..127 (the W_lo half) of the same 32 channels x 64 positions; the partner parks its half in the pair's
    // transposition tile, the main warp adds it, applies bias + ReLU, transposes and stores along positions.
    const bool upper = (warp & 3) >= 2;
    const int chg = warp & 1, part = warp >> 2;
    const int ewarp = chg + 2 * part;                  // pair index 0..7
    const int bar_id = 1 + ewarp;                      // named barrier of the pair (64 threads)
    float* const outt = reinterpret_cast<float*>(smem + P.smem_out) + (size_t)ewarp * 32 * kStOutPitch;
    const float bs = __ldg(bias + 32 * chg + lane);
    const int ch_sub = lane >> 4, pl16 = lane & 15;
    const long long cstride = (long long)P.hc * P.wc;
    Ring acc(2);
    for (int tile = blockIdx.x; tile < P.p_tiles; tile += gridDim.x) {
      mbar_wait_t(acc_full(acc.stage), acc.phase, err, 1, w0);
      tc_fence_after();
      for (int st = 0; st < 4; ++st) {
        const int p0 = part * 64 + 16 * st;
        uint32_t rr[16];
        tmem_ld16(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(acc.stage * kStTile + p0), rr);
        tmem_ld_wait();
        float4* orow = reinterpret_cast<float4*>(outt + lane * kStOutPitch);
        if (upper) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            orow[j] = make_float4(__uint_as_float(rr[4 * j]), __uint_as_float(rr[4 * j + 1]), __uint_as_float(rr[4 * j + 2]), __uint_as_float(rr[4 * j + 3]));
        }
        asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");      // half parked (both warps of the pair: one instruction)
        if (!upper) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 u = orow[j];
            orow[j] = make_float4(fmaxf(__uint_as_float(rr[4 * j]) + u.x + bs, 0.0f), fmaxf(__uint_as_float(rr[4 * j + 1]) + u.y + bs, 0.0f),
                                  fmaxf(__uint_as_float(rr[4 * j + 2]) + u.z + bs, 0.0f), fmaxf(__uint_as_float(rr[4 * j + 3]) + u.w + bs, 0.0f));
          }
          __syncwarp();
          const StPos pi = st_decode(P, P.q_begin + (long long)tile * kStTile + p0 + pl16);
          if (pi.in_range && pi.s < g.n && pi.a >= 0 && pi.a < P.hc && pi.col < P.wc) {
            const float* ot = outt + (4 * ch_sub) * kStOutPitch + pl16;
            float v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = ot[(8 * (i >> 2) + (i & 3)) * kStOutPitch];
            float* yp = conv + (((long long)pi.s * 64 + 32 * chg + 4 * ch_sub) * P.hc + pi.a) * P.wc + pi.col;
            const long long s5 = 5 * cstride;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              *yp = v[i];
              yp += ((i & 3) == 3) ? s5 : cstride;
            }
          }
        }
        __syncwarp();                                  // reconverge before the aligned barrier
        asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");      // tile consumed: may be overwritten
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty(acc.stage));
      acc.advance();
    }
  } else if (warp == 16) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      Ring acc(2), rp(kStPStages);
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kStTile >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t wbase = sbase + P.smem_w;
      for (int tile = blockIdx.x; tile < P.p_tiles; tile += gridDim.x) {
        mbar_wait_t(acc_empty(acc.stage), acc.phase ^ 1u, err, 2, w0);
        tc_fence_after();
        const uint32_t d0 = tmem_base + (uint32_t)(acc.stage * kStTile);
        for (int phase = 0; phase < 4; ++phase) {
          mbar_wait_t(p_full(rp.stage), rp.phase, err, 3, w1);
          tc_fence_after();
          const uint32_t p_hi = sbase + P.smem_p + (uint32_t)rp.stage * P.stage_bytes, p_lo = p_hi + P.phase_bytes;
#pragma unroll 1
          for (int term = 0; term < 2; ++term) {        // small operand first: [W_hi; W_lo] P_lo, then [W_hi; W_lo] P_hi
            const uint32_t pb = term == 0 ? p_lo : p_hi;
            for (int k = 0; k < P.count[phase]; ++k) {
              const int pair = P.first[phase] + k;
              const uint64_t ad = make_desc(wbase + (uint32_t)pair * 4096u, 2048u, 128u);
              const uint64_t bd = make_desc(pb + (uint32_t)P.pair_off[pair] * 16u, (uint32_t)P.pair_lbo[pair] * 16u, 128u);
              umma_tf32(d0, ad, bd, idesc, (phase | term | k) != 0 ? 1u : 0u);
            }
          }
          umma_commit(p_empty(rp.stage));
          rp.advance();
        }
        umma_commit(acc_full(acc.stage));
        acc.advance();
      }
    }
    __syncwarp();
  } else {
    // ===================== patch producers (warps 17-24) =====================
    const int pw = warp - 17;
    const int pt = pw * 32 + lane;
    const long long plane = (long long)g.h * g.w;
    Ring rp(kStPStages);
    for (int tile = blockIdx.x; tile < P.p_tiles; tile += gridDim.x) {
      const long long q0 = P.q_begin + (long long)tile * kStTile;
      for (int phase = 0; phase < 4; ++phase) {
        mbar_wait_t(p_empty(rp.stage), rp.phase ^ 1u, err, 6, w0);
        float4* hi = reinterpret_cast<float4*>(smem + P.smem_p + (size_t)rp.stage * P.stage_bytes);
        float4* lo = reinterpret_cast<float4*>(smem + P.smem_p + (size_t)rp.stage * P.stage_bytes + P.phase_bytes);
        const int py = phase >> 1, px = phase & 1;
        constexpr int kB = 4;
        for (int pos0 = pt; pos0 < P.pp; pos0 += kB * kStProducerWarps * 32) {
          float v[kB][3];
#pragma unroll
          for (int u = 0; u < kB; ++u) {
            const int pos = pos0 + u * kStProducerWarps * 32;
            v[u][0] = v[u][1] = v[u][2] = 0.0f;
            if (pos < P.pp) {
              const StPos pi = st_decode(P, q0 + P.dmin[phase] + pos);
              const int iy = 2 * pi.a + py, ix = 2 * pi.col + px;
              if (pi.in_range && pi.s < g.n && pi.a >= 0 && iy < g.h && ix < g.w) {
                const float* xp = x + ((long long)pi.s * 3 * g.h + iy) * g.w + ix;
                v[u][0] = __ldg(xp); v[u][1] = __ldg(xp + plane); v[u][2] = __ldg(xp + 2 * plane);
              }
            }
          }
#pragma unroll
          for (int u = 0; u < kB; ++u) {
            const int pos = pos0 + u * kStProducerWarps * 32;
            if (pos < P.pp) {
              float h3[3], l3[3];
#pragma unroll
              for (int c = 0; c < 3; ++c) {
                h3[c] = __uint_as_float(__float_as_uint(v[u][c]) & 0xFFFFE000u);
                l3[c] = __fsub_rn(v[u][c], h3[c]);
              }
              hi[pos] = make_float4(h3[0], h3[1], h3[2], 0.0f);
              lo[pos] = make_float4(l3[0], l3[1], l3[2], 0.0f);
            }
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full(rp.stage));
        rp.advance();
      }
    }
  }
#ifdef LSQ_TC_DIAG
  if (diag && lane == 0 && blockIdx.x == 0) { diag[warp * 3] = clock64() - t_start; diag[warp * 3 + 1] = w0; diag[warp * 3 + 2] = w1; }
#else
  (void)t_start; (void)w0; (void)w1; (void)diag;
#endif
  tc_fence_before();
  __syncthreads();
  if (warp == 16) tmem_dealloc(tmem_base, 512);
}

// max-pool 3x3 / stride 2 / pad 1 of the (already rectified, >= 0) convolution output
__global__ void __launch_bounds__(256)
stem_pool_kernel(const float* __restrict__ conv, float* __restrict__ out, int planes, int hc, int wc, int hp, int wp) {
  const long long total = (long long)planes * hp * wp;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int px = (int)(i % wp), py = (int)((i / wp) % hp);
    const long long pl = i / ((long long)wp * hp);
    const float* c = conv + pl * hc * wc;
    float m = 0.0f;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int cy = 2 * py - 1 + dy;
      if (cy < 0 || cy >= hc) continue;
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int cx = 2 * px - 1 + dx;
        if (cx < 0 || cx >= wc) continue;
        m = fmaxf(m, __ldg(c + (long long)cy * wc + cx));
      }
    }
    out[i] = m;
  }
}

// same, one CTA per (sample, channel) plane staged in shared memory: one coalesced read of the plane, one
// coalesced write of the pooled plane (HBM bound)
__global__ void __launch_bounds__(256)
stem_pool_plane_kernel(const float* __restrict__ conv, float* __restrict__ out, int hc, int wc, int hp, int wp) {
  extern __shared__ __align__(16) float plane[];
  const int np = hc * wc;
  const float* c = conv + (long long)blockIdx.x * np;
  if ((np & 3) == 0) {
    const float4* c4 = reinterpret_cast<const float4*>(c);
    float4* p4 = reinterpret_cast<float4*>(plane);
    for (int i = threadIdx.x; i < (np >> 2); i += 256) p4[i] = __ldg(c4 + i);
  } else {
    for (int i = threadIdx.x; i < np; i += 256) plane[i] = __ldg(c + i);
  }
  __syncthreads();
  float* o = out + (long long)blockIdx.x * hp * wp;
  for (int i = threadIdx.x; i < hp * wp; i += 256) {
    const int py = i / wp, px = i - py * wp;
    float m = 0.0f;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int cy = 2 * py - 1 + dy;
      if (cy < 0 || cy >= hc) continue;
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int cx = 2 * px - 1 + dx;
        if (cx < 0 || cx >= wc) continue;
        m = fmaxf(m, plane[cy * wc + cx]);
      }
    }
    o[i] = m;
  }
}

static bool stem_plan(int n, int h, int w, StemParams& P, size_t& smem_bytes) {
  lsq_act_geom g;
  if (lsq_act_geometry(n, 3, h, w, 7, 7, 2, 3, &g) != LSQ_OK) return false;
  if ((long long)g.n * g.rows_per_sample * g.pitch > (1ll << 31) - 4096) return false;
  P.g = to_dev(g);
  P.hc = g.ho; P.wc = g.wo;
  const StemTaps T = stem_taps();
  int dmax[4];
  for (int ph = 0; ph < 4; ++ph) { P.dmin[ph] = 1 << 30; dmax[ph] = -(1 << 30); P.first[ph] = T.first[ph]; P.count[ph] = T.count[ph]; }
  int off[kStPairs][2];
  for (int p = 0; p < kStPairs; ++p)
    for (int c = 0; c < 2; ++c) {
      off[p][c] = T.qy[p][c] * g.pitch + T.qx[p][c];
      const int ph = T.phase[p];
      if (off[p][c] < P.dmin[ph]) P.dmin[ph] = off[p][c];
      if (off[p][c] > dmax[ph]) dmax[ph] = off[p][c];
    }
  int span = 0;
  for (int ph = 0; ph < 4; ++ph) if (dmax[ph] - P.dmin[ph] > span) span = dmax[ph] - P.dmin[ph];
  for (int p = 0; p < kStPairs; ++p) {
    P.pair_off[p] = off[p][0] - P.dmin[T.phase[p]];
    P.pair_lbo[p] = off[p][1] - off[p][0];
    if (P.pair_lbo[p] <= 0 || P.pair_lbo[p] >= 0x3FFF) return false;
  }
  P.pp = (kStTile + span + 7) / 8 * 8;
  P.phase_bytes = (uint32_t)P.pp * 16u;
  P.stage_bytes = 2u * P.phase_bytes;
  uint32_t o = 0;
  P.smem_w = o; o += kStWeightBytes;
  P.smem_p = o; o += kStPStages * P.stage_bytes;
  P.smem_bar = o; o += 256;
  P.smem_out = o; o += 8 * 32 * kStOutPitch * 4;
  smem_bytes = o;
  if (smem_bytes > 227 * 1024) return false;
  P.pitch_magic = ((1ull << 40) + (unsigned long long)g.pitch - 1ull) / (unsigned long long)g.pitch;
  P.rps_magic = ((1ull << 40) + (unsigned long long)g.rows_per_sample - 1ull) / (unsigned long long)g.rows_per_sample;
  P.q_begin = (long long)g.lead + (long long)g.ph * g.pitch;
  const long long qspan = (long long)g.n * g.rows_per_sample * g.pitch;
  P.p_tiles = (int)((qspan + kStTile - 1) / kStTile);
  return true;
}

// ================================================================================================================
// Fused stem for images up to 250 pixels wide: conv 7x7 / 2 + bias + ReLU + max-pool 3x3 / 2 in ONE kernel.
// The convolution output (1.6 GB at batch 512) never reaches HBM.
//  * operands are fp16 pairs: x = hi + lo and w = hi + lo (11 + 11 significant bits; every weight row is first
//    scaled by a power of two so that its largest entry sits in [2^13, 2^14) and its lo parts stay normal; the
//    epilogue undoes the scale exactly).  A patch position holds 16 bytes [xh0 xh1 xh2 0 xl0 xl1 xl2 0], a weight
//    row [w0 w1 w2 0 w0 w1 w2 0], so one K chunk of kind::f16 contracts w * (x_hi + x_lo) for one tap; the two K
//    chunks of an instruction are two taps (two shifts of the same patch, as above); rows 0..63 carry W_hi and
//    rows 64..127 W_lo as above.  All four product terms in 25 instructions per tile (the TF32 kernel: 50).
//  * the kernel has its own raster: pitch pw = wc + 3 rounded up to 8, a tile = 2 * pw positions = the two
//    convolution rows (2k, 2k+1) of one sample = exactly what pooled row k needs besides row 2k-1.  A CTA walks
//    a contiguous range of tiles; the horizontal maxima of row 2k-1 stay in the epilogue threads' registers from
//    the previous tile (a range that starts inside a sample recomputes one tile to get them).
//  * epilogue: thread = channel (the accumulator's own layout), so both pooling directions are register maxima;
//    the W_lo half of the accumulator is added through a shared-memory exchange between warp pairs first (the sum
//    must precede the ReLU / max).  The pooled row (64 x wp) is transposed through shared memory and stored
//    along px.
constexpr int kSfXPitch = 20;                  // exchange tile: [32 channels][8 columns x 2 rows + 2 left neighbours], float4 conflict-free
constexpr int kSfPoolPitch = 68;               // pooled tile: [64 channels][<= 64 px], rows 16-byte aligned for the float4 store
constexpr int kSfProducerWarps = 6;
constexpr int kSfThreads = (17 + kSfProducerWarps) * 32;     // 16 epilogue warps, MMA warp 16, producer warps 17-22
constexpr uint32_t kSfImageOffset = kStWeightBytes;          // fp16 image behind the TF32 image
constexpr uint32_t kSfScaleOffset = 2 * kStWeightBytes;      // 64 x 2^-e (fp32)

struct SfParams {
  int n, h, w, hc, wc, hp, wp;
  int pw, ncols, pp, nch, tiles;
  uint32_t pw_magic;
  int first[4], count[4];
  int pair_off[kStPairs], pair_lbo[kStPairs];
  uint32_t stage_bytes, smem_w, smem_p, smem_bar, smem_x, smem_pool, smem_lut;
};

__device__ __forceinline__ int stem_scale_exp(const float* __restrict__ w, int ch) {
  float mx = 0.0f;
  for (int j = 0; j < 147; ++j) mx = fmaxf(mx, fabsf(__ldg(w + ch * 147 + j)));
  if (!(mx > 0.0f) || !(mx < 3.0e38f)) return 0;
  const int e = 13 - ilogbf(mx);
  return e < -100 ? -100 : (e > 100 ? 100 : e);
}

// fp16 image[pair][chunk][row][slot]: slots 0..2 and 4..6 = the three input channels of tap(pair, chunk) of output
// channel row & 63, scaled by 2^e(channel); rows 0..63 the fp16 value, rows 64..127 the fp16 remainder
__global__ void stem_pack_f16_kernel(const float* __restrict__ w, __half* __restrict__ image, float* __restrict__ inv_scale,
                                     StemTaps T) {
  const int total = kStPairs * 2 * 128 * 8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int slot = i & 7, row = (i >> 3) & 127, chunk = (i >> 10) & 1, pair = i >> 11;
    const int ch = row & 63, c = slot & 3;
    const int tap = T.tap[pair][chunk];
    const int e = stem_scale_exp(w, ch);
    const float v = (c < 3 && tap >= 0) ? __ldg(w + ch * 147 + c * 49 + tap) * ldexpf(1.0f, e) : 0.0f;
    const __half hi = __float2half_rn(v);
    image[i] = row < 64 ? hi : __float2half_rn(__fsub_rn(v, __half2float(hi)));
    if (i < 64) inv_scale[i] = ldexpf(1.0f, -stem_scale_exp(w, i));
  }
}

// U8: the input is uint8 pixels [n,3,h,w] and `lut` float[3][256] gives the fp32 value of every pixel level of every channel
// (the caller's ToTensor + Normalize, evaluated once per level): the producers look the fp16 (hi, lo) pair of a pixel up in
// a shared-memory table built from it, so the result equals the fp32 route on lut[c][pixel] bit for bit.
template <bool U8>
__global__ void __launch_bounds__(kSfThreads, 1)
stem_fused_kernel(const float* __restrict__ x, const unsigned char* __restrict__ xu8, const float* __restrict__ lut,
                  SfParams P, const unsigned char* __restrict__ wimage,
                  const float* __restrict__ inv_scale, const float* __restrict__ bias, float* __restrict__ out,
                  long long* __restrict__ diag) {
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long w0 = 0, w1 = 0;
  const long long t_start = LSQ_TC_CLOCK();
#ifdef LSQ_TC_DIAG
  long long sec[6] = {0, 0, 0, 0, 0, 0}, tq = 0;
#define SF_T0() tq = clock64()
#define SF_T(i) do { const long long tn = clock64(); sec[i] += tn - tq; tq = tn; } while (0)
#else
#define SF_T0()
#define SF_T(i)
#endif
  // Barriers.  The patches have one buffer per stride phase; the phases of one input row parity (a "group": phases
  // 2g, 2g+1) are produced and released together.  g_full[2]: patches of group g written; g_empty: the instructions
  // of group 0 have read them; group 1 is released by acc_full (its instructions are the last of a tile) -- two
  // tcgen05.commit per tile, each costs the issuing thread several hundred cycles.  acc_full[2] acc_empty[2] | tmem base
  const uint32_t bar0 = sbase + P.smem_bar;
  auto g_full = [&](int g) { return bar0 + 8u * g; };
  const uint32_t g_empty = bar0 + 8u * 2;
  auto acc_full = [&](int s) { return bar0 + 8u * (8 + s); };
  auto acc_empty = [&](int s) { return bar0 + 8u * (10 + s); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + P.smem_bar + 8u * 12);
  if (threadIdx.x == 0) {
    for (int g = 0; g < 2; ++g) mbar_init(g_full(g), kSfProducerWarps / 2);
    mbar_init(g_empty, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(acc_full(s), 1); mbar_init(acc_empty(s), 16); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  {
    const float4* src = reinterpret_cast<const float4*>(wimage);
    float4* dst = reinterpret_cast<float4*>(smem + P.smem_w);
    for (int i = threadIdx.x; i < (int)(kStWeightBytes / 16); i += kSfThreads) dst[i] = __ldg(src + i);
  }
  // operand descriptors of the 25 instructions of a tile: constant for the whole kernel (one ring stage per phase)
  unsigned long long* const desc_tab = reinterpret_cast<unsigned long long*>(smem + P.smem_bar + 128);
  if (threadIdx.x < kStPairs) {
    const int pair = threadIdx.x;
    int phase = 0;
    while (phase < 3 && pair >= P.first[phase + 1]) ++phase;
    const uint32_t pb = sbase + P.smem_p + (uint32_t)phase * P.stage_bytes;
    desc_tab[2 * pair] = make_desc(sbase + P.smem_w + (uint32_t)pair * 4096u, 2048u, 128u);
    desc_tab[2 * pair + 1] = make_desc(pb + (uint32_t)P.pair_off[pair] * 16u, (uint32_t)P.pair_lbo[pair] * 16u, 128u);
  }
  uint32_t* const lutp = reinterpret_cast<uint32_t*>(smem + P.smem_lut);      // U8: [3][256] fp16 hi | lo << 16
  if (U8) {
    for (int i = threadIdx.x; i < 768; i += kSfThreads) {
      const float f = fminf(fmaxf(__ldg(lut + i), -65504.0f), 65504.0f);
      const __half hi = __float2half_rn(f);
      const __half lo = __float2half_rn(__fsub_rn(f, __half2float(hi)));
      lutp[i] = (uint32_t)__half_as_ushort(hi) | ((uint32_t)__half_as_ushort(lo) << 16);
    }
  }
  fence_proxy_async();
  if (warp == 16) tmem_alloc(smem_u32((const void*)tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  int* const err = nullptr;
  // this CTA's tiles: [t_own, t_end) are stored; a range that starts inside a sample first redoes the tile before it
  const int t_own = (int)((long long)P.tiles * blockIdx.x / gridDim.x);
  const int t_end = (int)((long long)P.tiles * (blockIdx.x + 1) / gridDim.x);
  const int t_begin = (t_own < t_end && (t_own % P.hp) != 0) ? t_own - 1 : t_own;

  if (warp < 16) {
    // ===================== epilogue (16 warps) =====================
    // Warp pair (qd, qd + 2) of a part owns 32 channels x up to 32 convolution columns (two 16-column chunks) of
    // both rows: the W_hi warp finishes chunk 0, the W_lo warp chunk 1, each after receiving the other's raw half
    // of its chunk through the pair's exchange tile.  Pooling runs on the raw sums (the map acc -> relu(acc * 2^-e +
    // bias) is monotone, so it commutes with max exactly); positions outside the map count as -inf.
    const int qd = warp & 3, part = warp >> 2;
    const int mine = qd >> 1;                          // chunk this warp finishes (0: W_hi warp, 1: W_lo warp)
    const int chg = qd & 1;
    const int pairid = chg + 2 * part;
    const int bar_id = 1 + pairid;                     // named barrier of the warp pair (64 threads)
    float* const xpair = reinterpret_cast<float*>(smem + P.smem_x) + (size_t)pairid * 2 * 32 * kSfXPitch + lane * kSfXPitch;
    float4* const xsend = reinterpret_cast<float4*>(xpair + mine * 32 * kSfXPitch);          // my half of the partner's chunk
    const float4* const xrecv = reinterpret_cast<const float4*>(xpair + (mine ^ 1) * 32 * kSfXPitch);
    float* const pool = reinterpret_cast<float*>(smem + P.smem_pool);
    const float bs = __ldg(bias + 32 * chg + lane), inv = __ldg(inv_scale + 32 * chg + lane);
    const uint32_t trow = tmem_base + ((uint32_t)(qd * 32) << 16);
    const int pw = P.pw, wc = P.wc, wp = P.wp;
    const int cb_m = 16 * (part * P.nch + mine), cb_o = 16 * (part * P.nch + (mine ^ 1));   // first columns of the chunks
    const bool have_m = mine < P.nch && cb_m < pw, have_o = (mine ^ 1) < P.nch && cb_o < pw;
    const float ninf = __int_as_float(0xff800000);
    float carry[8];                                    // raw horizontal maxima of convolution row 2k-1, px = cb_m / 2 + i
    // pooled row -> global: float4 units when rows are 16-byte multiples, else a scalar walk
    // pooled row -> global: thread t stores channel t >> 3, 16-byte quads (t & 7) and (t & 7) + 8 of the row (rows are
    // 16-byte multiples and at most 16 quads wide), else a scalar walk
    const bool vec_store = (wp & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    const int wq = wp >> 2;
    const int st_q = (int)threadIdx.x & 7, st_ch = (int)threadIdx.x >> 3;
    const int so0 = st_ch * kSfPoolPitch + 4 * st_q;
    const long long go0 = (long long)st_ch * P.hp * wp + 4 * st_q;
    Ring acc(2);
    uint32_t emitted = 0;
    for (int tile = t_begin; tile < t_end; ++tile) {
      const int s = tile / P.hp, k = tile - s * P.hp;
      const bool emit = tile >= t_own;
      if (k == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) carry[i] = ninf;   // no row -1
      }
      float* const ptile = pool + (size_t)(emitted & 1u) * 64 * kSfPoolPitch;
      mbar_wait_t(acc_full(acc.stage), acc.phase, err, 1, w0);
      tc_fence_after();
      // two 8-column halves of the chunk, both convolution rows at once (keeps the live registers few)
      float left0 = ninf, left1 = ninf;                // column before the current half: row 2k, row 2k+1
      const uint32_t tcol = trow + (uint32_t)(acc.stage * 256);
      float* const prow = ptile + (32 * chg + lane) * kSfPoolPitch + (cb_m >> 1);
      const bool row1 = 2 * k + 1 < P.hc;
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        SF_T0();
        const int cm = cb_m + 8 * hf, co = cb_o + 8 * hf;
        const bool do_m = have_m && cm < pw, do_o = have_o && co < pw;
        uint32_t a0[8], a1[8], b0[8], b1[8], al0 = 0u, al1 = 0u, bl0 = 0u, bl1 = 0u;
        if (do_o) {
          tmem_ld8(tcol + co, b0);
          tmem_ld8(tcol + pw + co, b1);
          if (hf == 0 && co > 0) { tmem_ld1(tcol + co - 1, bl0); tmem_ld1(tcol + pw + co - 1, bl1); }
        }
        if (do_m) {
          tmem_ld8(tcol + cm, a0);
          tmem_ld8(tcol + pw + cm, a1);
          if (hf == 0 && cm > 0) { tmem_ld1(tcol + cm - 1, al0); tmem_ld1(tcol + pw + cm - 1, al1); }
        }
        tmem_ld_wait();
        SF_T(0);
        if (do_o) {
          xsend[0] = make_float4(__uint_as_float(b0[0]), __uint_as_float(b0[1]), __uint_as_float(b0[2]), __uint_as_float(b0[3]));
          xsend[1] = make_float4(__uint_as_float(b0[4]), __uint_as_float(b0[5]), __uint_as_float(b0[6]), __uint_as_float(b0[7]));
          xsend[2] = make_float4(__uint_as_float(b1[0]), __uint_as_float(b1[1]), __uint_as_float(b1[2]), __uint_as_float(b1[3]));
          xsend[3] = make_float4(__uint_as_float(b1[4]), __uint_as_float(b1[5]), __uint_as_float(b1[6]), __uint_as_float(b1[7]));
          if (hf == 0) reinterpret_cast<float2*>(xsend)[8] = make_float2(__uint_as_float(bl0), __uint_as_float(bl1));
        }
        asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");      // halves parked
        SF_T(1);
        if (do_m) {
          const float4 u0 = xrecv[0], u1 = xrecv[1], u2 = xrecv[2], u3 = xrecv[3];
          float v0[8], v1[8];
          v0[0] = __uint_as_float(a0[0]) + u0.x; v0[1] = __uint_as_float(a0[1]) + u0.y;
          v0[2] = __uint_as_float(a0[2]) + u0.z; v0[3] = __uint_as_float(a0[3]) + u0.w;
          v0[4] = __uint_as_float(a0[4]) + u1.x; v0[5] = __uint_as_float(a0[5]) + u1.y;
          v0[6] = __uint_as_float(a0[6]) + u1.z; v0[7] = __uint_as_float(a0[7]) + u1.w;
          v1[0] = __uint_as_float(a1[0]) + u2.x; v1[1] = __uint_as_float(a1[1]) + u2.y;
          v1[2] = __uint_as_float(a1[2]) + u2.z; v1[3] = __uint_as_float(a1[3]) + u2.w;
          v1[4] = __uint_as_float(a1[4]) + u3.x; v1[5] = __uint_as_float(a1[5]) + u3.y;
          v1[6] = __uint_as_float(a1[6]) + u3.z; v1[7] = __uint_as_float(a1[7]) + u3.w;
          if (hf == 0 && cm > 0) {
            const float2 ul = reinterpret_cast<const float2*>(xrecv)[8];
            left0 = __uint_as_float(al0) + ul.x;
            left1 = __uint_as_float(al1) + ul.y;
          }
          if (cm + 8 > wc) {                                             // right edge of the map (uniform)
#pragma unroll
            for (int i = 0; i < 8; ++i)
              if (cm + i >= wc) { v0[i] = ninf; v1[i] = ninf; }
          }
          if (!row1) {                                                   // odd map height: no row 2k+1
            left1 = ninf;
#pragma unroll
            for (int i = 0; i < 8; ++i) v1[i] = ninf;
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float p0 = i == 0 ? left0 : v0[2 * i - 1], p1 = i == 0 ? left1 : v1[2 * i - 1];
            const float h1 = fmaxf(fmaxf(p1, v1[2 * i]), v1[2 * i + 1]);
            const float h0 = fmaxf(fmaxf(p0, v0[2 * i]), v0[2 * i + 1]);
            const float hm = fmaxf(fmaxf(h0, h1), carry[4 * hf + i]);
            carry[4 * hf + i] = h1;
            if (emit) prow[4 * hf + i] = fmaxf(fmaf(hm, inv, bs), 0.0f);
          }
          left0 = v0[7]; left1 = v1[7];
        }
        SF_T(2);
        asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");      // halves consumed
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty(acc.stage));
      acc.advance();
      SF_T(3);
      if (emit) {
#ifdef LSQ_TC_DIAG
        const long long tb = clock64();
#endif
        asm volatile("bar.sync 15, 512;" ::: "memory");                  // pooled row complete
#ifdef LSQ_TC_DIAG
        w1 += clock64() - tb;
#endif
        SF_T0();
        float* const ob = out + ((long long)s * 64 * P.hp + k) * wp;
        if (vec_store) {
          if (st_q < wq) *reinterpret_cast<float4*>(ob + go0) = *reinterpret_cast<const float4*>(ptile + so0);
          if (st_q + 8 < wq) *reinterpret_cast<float4*>(ob + go0 + 32) = *reinterpret_cast<const float4*>(ptile + so0 + 32);
        } else {
          const long long chs = (long long)P.hp * wp;
          const int dch = 512 / wp, dpx = 512 - dch * wp;
          int ch = (int)threadIdx.x / wp, px = (int)threadIdx.x - ch * wp;
          while (ch < 64) {
            ob[ch * chs + px] = ptile[ch * kSfPoolPitch + px];
            px += dpx; ch += dch;
            if (px >= wp) { px -= wp; ++ch; }
          }
        }
        SF_T(4);
        ++emitted;
      }
    }
  } else if (warp == 16) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      Ring acc(2);
      uint32_t par = 0;
      const uint32_t idesc = (1u << 4) | ((uint32_t)(P.ncols >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // f16 x f16 -> f32
      const ulonglong2* const dt = reinterpret_cast<const ulonglong2*>(desc_tab);
      const int pfirst[5] = {P.first[0], P.first[1], P.first[2], P.first[3], kStPairs};
      for (int tile = t_begin; tile < t_end; ++tile) {
        mbar_wait_t(acc_empty(acc.stage), acc.phase ^ 1u, err, 2, w0);
        tc_fence_after();
        const uint32_t d0 = tmem_base + (uint32_t)(acc.stage * 256);
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          mbar_wait_t(g_full(g), par, err, 3, w1);
          tc_fence_after();
#pragma unroll 4
          for (int pair = pfirst[2 * g]; pair < pfirst[2 * g + 2]; ++pair) {
            const ulonglong2 d = dt[pair];
            umma_f16(d0, d.x, d.y, idesc, pair != 0 ? 1u : 0u);
          }
          if (g == 0) umma_commit(g_empty);
        }
        umma_commit(acc_full(acc.stage));
        acc.advance();
        par ^= 1u;
      }
    }
    __syncwarp();
  } else {
    // ===================== patch producers (warps 17-22) =====================
    // Three warps per input row parity py.  A lane loads pixel pairs (ix, ix + 1) = the px = 0 and px = 1 phase
    // entries of one position (8 contiguous bytes when the row pitch allows), so every fetched sector is used once;
    // the loads of the next tile are issued before waiting for the ring slots (their latency is what a tile waits for).
    const int pwarp = warp - 17;
    const int py = pwarp & 1, third = pwarp >> 1;
    const long long plane = (long long)P.h * P.w;
    uint4* const patch0 = reinterpret_cast<uint4*>(smem + P.smem_p + (size_t)(2 * py) * P.stage_bytes);
    uint4* const patch1 = reinterpret_cast<uint4*>(smem + P.smem_p + (size_t)(2 * py + 1) * P.stage_bytes);
    const bool pair_loads = (P.w & 1) == 0 && (U8 ? (reinterpret_cast<uintptr_t>(xu8) & 1) == 0 : (reinterpret_cast<uintptr_t>(x) & 7) == 0);
    uint32_t par = 0;
    Ring prev(2);                                        // accumulator stage of the previous tile (group 1's release)
    constexpr int kB = 7;                                // 96 * 7 >= the largest patch (648 positions)
    for (int tile = t_begin; tile < t_end; ++tile) {
      const int s = tile / P.hp, k = tile - s * P.hp;
      const float* const xs = x + (U8 ? 0ll : (long long)s * 3 * plane);
      const unsigned char* const us = xu8 + (U8 ? (long long)s * 3 * plane : 0ll);
      float2 v[kB][3];                                   // fp32 route: pixel pairs of the three channels
      uint32_t ra[kB][3];                                // uint8 route: the raw pixel pair of each channel (x | y << 8), bit 31: outside,
                                                         // bit 30: second pixel outside -- untouched until the slot wait below, so that
                                                         // all loads of a tile are in flight together
#pragma unroll
      for (int u = 0; u < kB; ++u) {
        const int pos = third * 32 + lane + 96 * u;
        const int prow = (int)(((uint32_t)pos * P.pw_magic) >> 20);
        const int bcol = pos - prow * P.pw - 2;
        const int a = 2 * k - 2 + prow;
        const int iy = 2 * a + py, ix = 2 * bcol;
        const bool inside = pos < P.pp && bcol >= 0 && ix < P.w && a >= 0 && iy < P.h;
        if (U8) {
          ra[u][0] = ra[u][1] = ra[u][2] = 0x80000000u;
          if (inside) {
            const unsigned char* up = us + (long long)iy * P.w + ix;
            if (pair_loads) {                            // w even: ix + 1 < w
#pragma unroll
              for (int c = 0; c < 3; ++c) ra[u][c] = (uint32_t)__ldg(reinterpret_cast<const unsigned short*>(up + c * plane));
            } else {
              const bool second = ix + 1 < P.w;
#pragma unroll
              for (int c = 0; c < 3; ++c) {
                ra[u][c] = (uint32_t)__ldg(up + c * plane);
                if (second) ra[u][c] |= (uint32_t)__ldg(up + c * plane + 1) << 8; else ra[u][c] |= 0x40000000u;
              }
            }
          }
        } else {
          v[u][0] = v[u][1] = v[u][2] = make_float2(0.0f, 0.0f);
          if (inside) {
            const float* xp = xs + (long long)iy * P.w + ix;
            if (pair_loads) {                            // w even: ix + 1 < w
#pragma unroll
              for (int c = 0; c < 3; ++c) v[u][c] = __ldg(reinterpret_cast<const float2*>(xp + c * plane));
            } else {
              const bool second = ix + 1 < P.w;
#pragma unroll
              for (int c = 0; c < 3; ++c) {
                v[u][c].x = __ldg(xp + c * plane);
                if (second) v[u][c].y = __ldg(xp + c * plane + 1);
              }
            }
          }
        }
      }
      // the previous tile's instructions of this group have read the patches
      if (py == 0) mbar_wait_t(g_empty, par ^ 1u, err, 6, w0);
      else if (tile > t_begin) { mbar_wait_t(acc_full(prev.stage), prev.phase, err, 7, w0); prev.advance(); }
#pragma unroll
      for (int u = 0; u < kB; ++u) {
        const int pos = third * 32 + lane + 96 * u;
        if (pos < P.pp) {
          uint4 q[2];
          if (U8) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              q[e] = make_uint4(0u, 0u, 0u, 0u);         // outside the image: zero padding of the NORMALISED tensor
              const bool out = (ra[u][0] & 0x80000000u) != 0u || (e == 1 && (ra[u][0] & 0x40000000u) != 0u);
              if (!out) {
                const uint32_t e0 = lutp[(ra[u][0] >> (8 * e)) & 0xFFu], e1 = lutp[256 + ((ra[u][1] >> (8 * e)) & 0xFFu)],
                               e2 = lutp[512 + ((ra[u][2] >> (8 * e)) & 0xFFu)];
                q[e] = make_uint4((e0 & 0xFFFFu) | (e1 << 16), e2 & 0xFFFFu, (e0 >> 16) | (e1 & 0xFFFF0000u), e2 >> 16);
              }
            }
          } else {
            // hi = fp16(x), lo = fp16(x - hi), two values per conversion instruction; |x| clamped to the fp16 range
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const float f0 = fminf(fmaxf(e == 0 ? v[u][0].x : v[u][0].y, -65504.0f), 65504.0f);
              const float f1 = fminf(fmaxf(e == 0 ? v[u][1].x : v[u][1].y, -65504.0f), 65504.0f);
              const float f2 = fminf(fmaxf(e == 0 ? v[u][2].x : v[u][2].y, -65504.0f), 65504.0f);
              const __half2 h01 = __floats2half2_rn(f0, f1), h2z = __floats2half2_rn(f2, 0.0f);
              const float2 b01 = __half22float2(h01), b2z = __half22float2(h2z);
              const __half2 l01 = __floats2half2_rn(__fsub_rn(f0, b01.x), __fsub_rn(f1, b01.y));
              const __half2 l2z = __floats2half2_rn(__fsub_rn(f2, b2z.x), 0.0f);
              q[e] = make_uint4(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h2z),
                                *reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l2z));
            }
          }
          patch0[pos] = q[0];
          patch1[pos] = q[1];
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(g_full(py));
      par ^= 1u;
    }
  }
#ifdef LSQ_TC_DIAG
  if (diag && lane == 0 && blockIdx.x == 0) { diag[warp * 3] = clock64() - t_start; diag[warp * 3 + 1] = w0; diag[warp * 3 + 2] = w1; }
  if (diag && lane == 0 && blockIdx.x == 0 && warp == 0) for (int i = 0; i < 6; ++i) diag[75 + i] = sec[i];
#else
  (void)t_start; (void)w0; (void)w1; (void)diag;
#endif
#undef SF_T0
#undef SF_T
  tc_fence_before();
  __syncthreads();
  if (warp == 16) tmem_dealloc(tmem_base, 512);
}

// out[p][i] = lut[(p % c) * 256 + x[p][i]]: uint8 planes -> fp32 through a per-channel table of the 256 pixel levels
// (ToTensor + Normalize of the caller, evaluated once per level).  16 pixels per thread per step.
__global__ void __launch_bounds__(256)
u8_expand_kernel(const unsigned char* __restrict__ x, long long planes, int c, long long inner, const float* __restrict__ lut,
                 float* __restrict__ out) {
  extern __shared__ float slut[];
  for (int i = threadIdx.x; i < c * 256; i += blockDim.x) slut[i] = __ldg(lut + i);
  __syncthreads();
  const long long total = planes * inner;
  const bool vec = (inner & 15) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  if (vec) {
    const long long n16 = total >> 4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (long long)gridDim.x * blockDim.x) {
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(x) + i);
      const float* const t = slut + (int)(((i << 4) / inner) % c) * 256;
      const uint32_t w[4] = {q.x, q.y, q.z, q.w};
      float4* o = reinterpret_cast<float4*>(out) + (i << 2);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        o[j] = make_float4(t[w[j] & 0xFFu], t[(w[j] >> 8) & 0xFFu], t[(w[j] >> 16) & 0xFFu], t[w[j] >> 24]);
    }
  } else {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
      out[i] = slut[(int)((i / inner) % c) * 256 + x[i]];
  }
}

static bool stem_fused_plan(int n, int h, int w, SfParams& P, size_t& smem_bytes) {
  if (n <= 0 || h < 7 || w < 7) return false;
  P.n = n; P.h = h; P.w = w;
  P.hc = (h - 1) / 2 + 1; P.wc = (w - 1) / 2 + 1;
  P.hp = (P.hc - 1) / 2 + 1; P.wp = (P.wc - 1) / 2 + 1;
  P.pw = (P.wc + 3 + 7) / 8 * 8;
  if (P.pw > 128) return false;                              // two convolution rows must fit one N = 256 tile
  P.ncols = 2 * P.pw;
  P.nch = (P.wp + 31) / 32;                                  // 4 epilogue parts x nch chunks x 8 px >= wp
  if ((long long)n * P.hp > (1ll << 30)) return false;
  P.tiles = n * P.hp;
  P.pw_magic = ((1u << 20) + (uint32_t)P.pw - 1u) / (uint32_t)P.pw;
  const StemTaps T = stem_taps();
  int max_off = 0;
  for (int ph = 0; ph < 4; ++ph) { P.first[ph] = T.first[ph]; P.count[ph] = T.count[ph]; }
  for (int p = 0; p < kStPairs; ++p) {
    const int o0 = (T.qy[p][0] + 2) * P.pw + T.qx[p][0] + 2, o1 = (T.qy[p][1] + 2) * P.pw + T.qx[p][1] + 2;
    P.pair_off[p] = o0;
    P.pair_lbo[p] = o1 - o0;
    if (o0 < 0 || P.pair_lbo[p] <= 0 || P.pair_lbo[p] >= 0x3FFF) return false;
    if (o1 > max_off) max_off = o1;
  }
  P.pp = (P.ncols + max_off + 7) / 8 * 8;
  if ((long long)P.pp * P.pw_magic >= (1ll << 32)) return false;
  P.stage_bytes = (uint32_t)P.pp * 16u;
  uint32_t o = 0;
  P.smem_w = o; o += kStWeightBytes;
  P.smem_p = o; o += kStPStages * P.stage_bytes;
  P.smem_bar = o; o += 128 + kStPairs * 16 + 48;      // barriers + tmem slot, then the descriptor table
  P.smem_x = o; o += 8 * 2 * 32 * kSfXPitch * 4;
  o = (o + 15u) & ~15u;
  P.smem_pool = o; o += 2 * 64 * kSfPoolPitch * 4;
  P.smem_lut = o; o += 3 * 256 * 4;
  smem_bytes = o;
  return smem_bytes <= 227 * 1024;
}

}  // namespace lsq

using namespace lsq;

extern "C" size_t lsq_stem_image_bytes(void) { return 2 * (size_t)kStWeightBytes + 64 * sizeof(float); }

extern "C" int lsq_stem_is_fused(int n, int h, int w) {
  SfParams F;
  size_t smem = 0;
  return stem_fused_plan(n, h, w, F, smem) ? 1 : 0;
}

extern "C" size_t lsq_stem_workspace_bytes(int n, int h, int w) {
  if (n <= 0 || h < 7 || w < 7) return 0;
  if (lsq_stem_is_fused(n, h, w)) return 256;      // the fused kernel keeps the convolution output on chip
  const size_t hc = (size_t)(h - 1) / 2 + 1, wc = (size_t)(w - 1) / 2 + 1;
  return (size_t)n * 64 * hc * wc * sizeof(float);
}

extern "C" int lsq_stem_supported(int n, int h, int w) {
  StemParams P;
  size_t smem = 0;
  return (n > 0 && h >= 7 && w >= 7 && (lsq_stem_is_fused(n, h, w) || stem_plan(n, h, w, P, smem))) ? 1 : 0;
}

extern "C" int lsq_stem_pack_weights(const float* d_w, float* d_image, void* stream) {
  LSQ_CHECK_ARG(d_w && d_image, "lsq_stem_pack_weights: null pointer");
  LSQ_CHECK_ARG(((uintptr_t)d_image & 15) == 0, "lsq_stem_pack_weights: image must be 16-byte aligned");
  stem_pack_kernel<<<50, 256, 0, (cudaStream_t)stream>>>(d_w, d_image, stem_taps());
  LSQ_CUDA_LAUNCH_CHECK("stem_pack_kernel");
  unsigned char* const base = reinterpret_cast<unsigned char*>(d_image);
  stem_pack_f16_kernel<<<50, 256, 0, (cudaStream_t)stream>>>(d_w, reinterpret_cast<__half*>(base + kSfImageOffset),
                                                           reinterpret_cast<float*>(base + kSfScaleOffset), stem_taps());
  LSQ_CUDA_LAUNCH_CHECK("stem_pack_f16_kernel");
  return LSQ_OK;
}

static int stem_fused_launch(const float* d_x, const unsigned char* d_xu8, const float* d_lut, int n, int h, int w,
                             const float* d_image, const float* d_bias, float* d_out, void* stream) {
  SfParams F;
  size_t smem = 0;
  if (!stem_fused_plan(n, h, w, F, smem)) { set_error("lsq_stem_fwd: image %dx%d does not fit the one-kernel route", h, w); return LSQ_ERR_UNSUPPORTED; }
  static std::atomic<unsigned long long> fused_set{0ull}, fused_set_u8{0ull};
  const cudaError_t fe = d_xu8 ? ensure_max_smem(stem_fused_kernel<true>, fused_set_u8) : ensure_max_smem(stem_fused_kernel<false>, fused_set);
  if (fe != cudaSuccess) { set_error("lsq_stem_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(fe)); return LSQ_ERR_CUDA; }
  const int sms_f = device_sms();
  const unsigned char* const base = reinterpret_cast<const unsigned char*>(d_image);
  const int grid = F.tiles < sms_f ? F.tiles : sms_f;
  long long* d_fdiag = nullptr;
#ifdef LSQ_TC_DIAG
  static const bool want_fdiag = getenv("LSQ_TC_DIAG") != nullptr;   // development build: wait cycles per warp of CTA 0
  if (want_fdiag) { cudaMalloc(&d_fdiag, 32 * 3 * sizeof(long long)); cudaMemsetAsync(d_fdiag, 0, 32 * 3 * sizeof(long long), (cudaStream_t)stream); }
#endif
  if (d_xu8)
    stem_fused_kernel<true><<<grid, kSfThreads, smem, (cudaStream_t)stream>>>(
        nullptr, d_xu8, d_lut, F, base + kSfImageOffset, reinterpret_cast<const float*>(base + kSfScaleOffset), d_bias, d_out, d_fdiag);
  else
    stem_fused_kernel<false><<<grid, kSfThreads, smem, (cudaStream_t)stream>>>(
        d_x, nullptr, nullptr, F, base + kSfImageOffset, reinterpret_cast<const float*>(base + kSfScaleOffset), d_bias, d_out, d_fdiag);
  LSQ_CUDA_LAUNCH_CHECK("stem_fused_kernel");
#ifdef LSQ_TC_DIAG
  if (want_fdiag) {
    long long hd[32 * 3];
    cudaStreamSynchronize((cudaStream_t)stream);
    cudaMemcpy(hd, d_fdiag, sizeof(hd), cudaMemcpyDeviceToHost);
    cudaFree(d_fdiag);
    fprintf(stderr, "[fused stem diag] tiles %d pw %d pp %d | warp 0 sections: tmem %lld pairbar %lld math %lld pool %lld store %lld\n", F.tiles, F.pw, F.pp, hd[75], hd[76], hd[77], hd[78], hd[79]);
    for (int wi = 0; wi < kSfThreads / 32; ++wi)
      fprintf(stderr, "   warp %2d (%s) total %9lld  wait0 %9lld  wait1 %9lld\n", wi,
              wi < 16 ? ((wi & 3) < 2 ? "epi-hi" : "epi-lo") : (wi == 16 ? "mma" : "producer"), hd[wi * 3], hd[wi * 3 + 1], hd[wi * 3 + 2]);
  }
#endif
  return LSQ_OK;
}

extern "C" int lsq_u8_expand(const unsigned char* d_x, int64_t planes, int c, int64_t inner, const float* d_lut, float* d_out,
                             void* stream) {
  LSQ_CHECK_ARG(d_x && d_lut && d_out, "lsq_u8_expand: null pointer");
  LSQ_CHECK_ARG(planes > 0 && inner > 0 && c > 0 && c <= 32 && planes % c == 0, "lsq_u8_expand: bad shape planes=%lld c=%d inner=%lld",
                (long long)planes, c, (long long)inner);
  const long long work = (planes * inner + 15) / 16;
  long long grid = (work + 255) / 256;
  const long long cap = (long long)device_sms() * 8;
  if (grid > cap) grid = cap;
  u8_expand_kernel<<<(unsigned)grid, 256, (size_t)c * 256 * sizeof(float), (cudaStream_t)stream>>>(d_x, planes, c, inner, d_lut, d_out);
  LSQ_CUDA_LAUNCH_CHECK("u8_expand_kernel");
  return LSQ_OK;
}

extern "C" int lsq_stem_fwd_u8(const unsigned char* d_x, int n, int h, int w, const float* d_lut, const float* d_image,
                               const float* d_bias, float* d_out, void* stream) {
  LSQ_CHECK_ARG(d_x && d_lut && d_image && d_bias && d_out, "lsq_stem_fwd_u8: null pointer");
  LSQ_CHECK_ARG(n > 0 && h >= 7 && w >= 7, "lsq_stem_fwd_u8: bad shape");
  LSQ_CHECK_ARG(((uintptr_t)d_image & 15) == 0, "lsq_stem_fwd_u8: weight image must be 16-byte aligned");
  return stem_fused_launch(nullptr, d_x, d_lut, n, h, w, d_image, d_bias, d_out, stream);
}

extern "C" int lsq_stem_fwd(const float* d_x, int n, int h, int w, const float* d_image, const float* d_bias,
                            float* d_conv_ws, float* d_out, void* stream) {
  LSQ_CHECK_ARG(d_x && d_image && d_bias && d_conv_ws && d_out, "lsq_stem_fwd: null pointer");
  LSQ_CHECK_ARG(n > 0 && h >= 7 && w >= 7, "lsq_stem_fwd: bad shape");
  LSQ_CHECK_ARG(((uintptr_t)d_image & 15) == 0, "lsq_stem_fwd: weight image must be 16-byte aligned");
  size_t smem = 0;
  {
    static const bool no_fuse = getenv("LSQ_STEM_UNFUSED") != nullptr;     // development: force the two-kernel route
    if (!no_fuse && lsq_stem_is_fused(n, h, w)) return stem_fused_launch(d_x, nullptr, nullptr, n, h, w, d_image, d_bias, d_out, stream);
  }
  StemParams P;
  if (!stem_plan(n, h, w, P, smem)) {
    set_error("lsq_stem_fwd: image %dx%d not supported (patch does not fit shared memory)", h, w);
    return LSQ_ERR_UNSUPPORTED;
  }
  static std::atomic<unsigned long long> smem_set{0ull}, pool_set{0ull};
  cudaError_t e = ensure_max_smem(stem_conv_kernel, smem_set);
  if (e != cudaSuccess) { set_error("lsq_stem_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return LSQ_ERR_CUDA; }
  const int sms = device_sms();
  const int grid = P.p_tiles < sms ? P.p_tiles : sms;
  long long* d_diag = nullptr;
#ifdef LSQ_TC_DIAG
  static const bool want_diag = getenv("LSQ_TC_DIAG") != nullptr;     // development build: wait cycles per warp of CTA 0
  if (want_diag) { cudaMalloc(&d_diag, 32 * 3 * sizeof(long long)); cudaMemsetAsync(d_diag, 0, 32 * 3 * sizeof(long long), (cudaStream_t)stream); }
#endif
  stem_conv_kernel<<<grid, kStThreads, smem, (cudaStream_t)stream>>>(d_x, P, d_image, d_bias, d_conv_ws, d_diag);
  LSQ_CUDA_LAUNCH_CHECK("stem_conv_kernel");
#ifdef LSQ_TC_DIAG
  if (want_diag) {
    long long h[32 * 3];
    cudaStreamSynchronize((cudaStream_t)stream);
    cudaMemcpy(h, d_diag, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(d_diag);
    fprintf(stderr, "[stem diag] tiles %d grid %d pp %d\n", P.p_tiles, grid, P.pp);
    for (int w = 0; w < kStThreads / 32; ++w)
      fprintf(stderr, "   warp %2d (%s) total %9lld  wait0 %9lld  wait1(p_full) %9lld\n", w,
              w < 16 ? ((w & 3) < 2 ? "epi-main" : "epi-upper") : (w == 16 ? "mma" : "producer"), h[w * 3], h[w * 3 + 1], h[w * 3 + 2]);
  }
#endif
  const int hp = (P.hc - 1) / 2 + 1, wp = (P.wc - 1) / 2 + 1;
  const size_t plane_bytes = (size_t)P.hc * P.wc * sizeof(float);
  if (plane_bytes <= 56 * 1024 && ((uintptr_t)d_conv_ws & 15) == 0) {
    e = ensure_max_smem(stem_pool_plane_kernel, pool_set);
    if (e != cudaSuccess) { set_error("lsq_stem_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return LSQ_ERR_CUDA; }
    stem_pool_plane_kernel<<<(unsigned)(n * 64), 256, plane_bytes, (cudaStream_t)stream>>>(d_conv_ws, d_out, P.hc, P.wc, hp, wp);
  } else {
    const long long total = (long long)n * 64 * hp * wp;
    unsigned pgrid = (unsigned)((total + 255) / 256);
    if (pgrid > 148u * 32u) pgrid = 148u * 32u;
    stem_pool_kernel<<<pgrid, 256, 0, (cudaStream_t)stream>>>(d_conv_ws, d_out, n * 64, P.hc, P.wc, hp, wp);
  }
  LSQ_CUDA_LAUNCH_CHECK("stem_pool_kernel");
  return LSQ_OK;
}
