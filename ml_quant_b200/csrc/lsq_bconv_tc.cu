// Binary convolution on the 5th-generation tensor cores (tcgen05.mma kind::i8, sm_100a).
//
// sm_100a has no 1-bit MMA; the +-1 planes are expanded to int8 on
// the fly and contracted with int8 +-1 weights, int32 accumulators in TMEM -- exact integers.
//
// Operand roles.  One tcgen05.mma with 8-bit operands (K = 32) takes ~135 clocks whatever M <= 128 and
// N <= 256 are (measured, scripts/mb/mb_umma.cu: 4.3 POP/s at M = 128, N = 256, half of that at N = 128), so
// the instruction must be as large as it can be for EVERY layer, also those with 64 or 128 output channels:
//     D[128 out-channels, N = positions x planes] += W_tap[128, 64] * P_tap[64, N]
//   A = weights  : M = 128 rows (a 64-channel layer stores its rows twice, lsq_bconv.cu: lanes 64..127 of the
//                  accumulator repeat lanes 0..63 and the epilogue warps of those lanes take other positions),
//   B = patch    : N = 256 columns = 128 consecutive positions x 2 planes interleaved (1 plane: 256 positions),
// summed over 64-channel blocks and taps.  The activation planes live in HBM as bits in a "virtual raster"
// (lsq_b200.h) in which a tap is a uniform shift of the position index, so ONE expanded int8 patch per channel
// block in shared memory (tile + halo, zero at padding positions) serves all kh*kw taps: the B operand of tap t
// is the same patch read through a shared-memory matrix descriptor whose start address is shifted by d(t)
// positions (K-major, no swizzle: a row is 16 bytes per 16-channel chunk, 8-row core matrices are contiguous,
// so any row shift is a 16-byte multiple).
//
// Warp roles (448 threads, one CTA per SM, persistent over (position tile, channel tile) items):
//   warps 0-3, 10-13  epilogue : thread = output channel (TMEM lane).  tcgen05.ld 16 accumulator columns ->
//                     vw[c] * (s1[n] I1 + s2[n] I2) + bias[c] -> transposed through shared memory so that
//                     global accesses run along positions (NCHW rows); activation and residual (requested at the
//                     start of the step, held in registers) are applied on the way out
//   warp  4           MMA      : one thread issues tcgen05.mma / tcgen05.commit; owns the TMEM allocation
//   warp  5           weights  : cp.async.bulk (TMA engine, 1-D) of pre-packed operand slabs, mbarrier tx
//   warps 6-9         patches  : plane bits (L2) -> int8 patch in shared memory, fence.proxy.async
// Pipelines: patch ring (per channel block), weight ring (per channel block x tap group), 2 accumulator
// stages in TMEM so the epilogue of item i overlaps the MMAs of item i+1.
#include <stdlib.h>
#include "lsq_common.cuh"
#include "lsq_tc.cuh"

namespace lsq {

constexpr int kTcThreads = 576;   // 4 epilogue + MMA + loader + 4 producer + 4 more epilogue + 4 more producer warps
constexpr int kProdThreads = 256; // (issue slots are less than half used: more warps hide the producers' L2 latency)
constexpr int kMaxTaps = 9;
constexpr int kMaxWStages = 6;
constexpr int kAccStages = 2;
constexpr int kMaxPStages = 3;
constexpr int kResStages = 4;     // residual staging: 32-position steps in flight per epilogue warp (4 KB each)
constexpr int kOutPitch = 36;     // transposition tile: 32 channels x 32 positions; 36 words: 16-byte rows, conflict-free both ways

struct TcParams {
  ActGeom g;
  int npl, cout, creal, n_ctiles, ncb, taps, tps;   // creal: distinct channels per 128-row tile (64 or 128)
  int tp, p_tiles;               // positions per tile (N = tp * npl), number of position tiles
  long long q_begin;             // first output position
  int pp;                        // patch positions per phase
  int p_stages, w_stages, r_stages;
  int w_resident;                // all weight slabs of the layer stay in shared memory (one channel tile, they fit): loaded once per CTA
  int dmin[4];                   // per phase: smallest tap offset (positions)
  int tap_phase[kMaxTaps];
  int tap_off[kMaxTaps];         // tap offset relative to dmin of its phase (>= 0)
  uint32_t lbo_p, phase_bytes, p_stage_bytes, w_slab_bytes, w_stage_bytes;
  uint32_t smem_p, smem_w, smem_bar, smem_tab, smem_scl, smem_out, smem_res, smem_as;  // offsets in dynamic smem
  int as_n;                      // samples whose activation scales are cached in shared memory (0: read from global)
  unsigned long long pitch_magic, rps_magic;   // ceil(2^40 / d): x / d == (x * magic) >> 40 for x * d < 2^40
};

// 4 bits -> 4 bytes: bit 1 -> 0x01 (+1), bit 0 -> 0xFF (-1)
__device__ __forceinline__ uint32_t expand4(uint32_t nib) {
  const uint32_t y = (nib * 0x00204081u) & 0x01010101u;
  return 0xFFFFFFFFu - y * 0xFEu;
}

// (n, row, col) of a virtual position, by multiplication (positions < 2^31, divisors <= 2^8..2^15)
struct PosInfo { int s, a, col; bool in_range; };
__device__ __forceinline__ PosInfo decode_pos(const TcParams& P, long long q) {
  PosInfo r;
  const long long rel = q - P.g.lead;
  r.in_range = rel >= 0;
  const unsigned long long urel = r.in_range ? (unsigned long long)rel : 0ull;
  const unsigned R = (unsigned)((urel * P.pitch_magic) >> 40);
  r.col = (int)(urel - (unsigned long long)R * (unsigned)P.g.pitch);
  r.s = (int)(((unsigned long long)R * P.rps_magic) >> 40);
  r.a = (int)R - r.s * P.g.rps - P.g.ph;
  return r;
}


__global__ void __launch_bounds__(kTcThreads, 1)
bconv_tc_kernel(const uint32_t* __restrict__ planes, TcParams P, const float* __restrict__ act_scales,
                const int8_t* __restrict__ wi8, const float* __restrict__ w_scale, const float* __restrict__ bias,
                float* __restrict__ y, int* __restrict__ err, Epilogue epi, long long* __restrict__ diag) {
  long long w0 = 0, w1 = 0, w2 = 0;
  const long long t_start = LSQ_TC_CLOCK();
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const ActGeom& g = P.g;

  // barrier block: p_full[kMaxPStages] p_empty[kMaxPStages] w_full[kMaxWStages] w_empty[kMaxWStages]
  //                acc_full[2] acc_empty[2] | tmem base
  const uint32_t bar0 = sbase + P.smem_bar;
  auto p_full = [&](int s) { return bar0 + 8u * s; };
  auto p_empty = [&](int s) { return bar0 + 8u * (kMaxPStages + s); };
  auto w_full = [&](int s) { return bar0 + 8u * (2 * kMaxPStages + s); };
  auto w_empty = [&](int s) { return bar0 + 8u * (2 * kMaxPStages + kMaxWStages + s); };
  auto acc_full = [&](int s) { return bar0 + 8u * (2 * kMaxPStages + 2 * kMaxWStages + s); };
  auto acc_empty = [&](int s) { return bar0 + 8u * (2 * kMaxPStages + 2 * kMaxWStages + kAccStages + s); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + P.smem_bar + 8u * (2 * kMaxPStages + 2 * kMaxWStages + 2 * kAccStages));

  if (threadIdx.x == 0) {
    for (int s = 0; s < kMaxPStages; ++s) { mbar_init(p_full(s), kProdThreads / 32); mbar_init(p_empty(s), 1); }
    for (int s = 0; s < kMaxWStages; ++s) { mbar_init(w_full(s), 1); mbar_init(w_empty(s), 1); }
    for (int s = 0; s < kAccStages; ++s) { mbar_init(acc_full(s), 1); mbar_init(acc_empty(s), 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // per-output-channel epilogue constants (weight scale, bias, PReLU slope)
  float4* const ctab = reinterpret_cast<float4*>(smem + P.smem_tab);
  for (int c = threadIdx.x; c < P.cout; c += kTcThreads)
    ctab[c] = make_float4(__ldg(w_scale + c), bias ? __ldg(bias + c) : 0.0f,
                          epi.act == 2 ? __ldg(epi.prelu + (epi.n_prelu > 1 ? c : 0)) : 0.0f, 0.0f);
  // Operand descriptors of every (weight stage, tap-in-stage, K half) and (patch stage, tap, K half): built once, so the
  // issuing thread's loop is two shared-memory loads and one tcgen05.mma per instruction (indexed parameter loads and
  // descriptor arithmetic in that single thread cost more than the instruction takes to execute).
  unsigned long long* const atab = reinterpret_cast<unsigned long long*>(smem + P.smem_bar + 512);
  unsigned long long* const btab = atab + kMaxWStages * 3 * 2;
  for (int i = threadIdx.x; i < kMaxWStages * 3 * 2 + kMaxPStages * kMaxTaps * 2; i += kTcThreads) {
    if (i < kMaxWStages * 3 * 2) {
      const int h = i & 1, tt = (i >> 1) % 3, ws = (i >> 1) / 3;
      const uint32_t a_addr = sbase + P.smem_w + (uint32_t)ws * P.w_stage_bytes + (uint32_t)tt * P.w_slab_bytes;
      atab[i] = make_desc(a_addr + (uint32_t)(2 * h) * (128u * 16u), 128u * 16u, 128u);
    } else {
      const int j = i - kMaxWStages * 3 * 2;
      const int h = j & 1, tap = (j >> 1) % kMaxTaps, ps = (j >> 1) / kMaxTaps;
      const uint32_t b_addr = sbase + P.smem_p + (uint32_t)ps * P.p_stage_bytes + (uint32_t)P.tap_phase[tap] * P.phase_bytes +
                              (uint32_t)(P.tap_off[tap] * P.npl) * 16u;
      btab[j] = make_desc(b_addr + (uint32_t)(2 * h) * P.lbo_p, P.lbo_p, 128u);
    }
  }
  // per-sample activation scales [planes][n]: a copy in shared memory (the epilogue looks two of them up per position
  // and item; from global memory that dependent L2 round trip sat in front of every item's accumulator wait)
  float* const as_tab = reinterpret_cast<float*>(smem + P.smem_as);
  for (int i = threadIdx.x; i < P.as_n * P.npl; i += kTcThreads) {
    const int pl = i / P.as_n, smp = i - pl * P.as_n;
    as_tab[i] = __ldg(act_scales + (long long)pl * g.n + smp);
  }
  if (warp == 4) tmem_alloc(smem_u32((const void*)tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int n_items = P.p_tiles * P.n_ctiles;
  const uint32_t acc_cols = (uint32_t)(P.tp * P.npl);

  if (warp < 4 || (warp >= 10 && warp < 14)) {
    // ===================== epilogue (8 warps) =====================
    // quarter = TMEM lane quarter; 128-channel tiles: quarter = channel group, the two warps of a quarter split
    // the positions in 2; 64-channel tiles: lanes 64..127 repeat lanes 0..63, positions are split in 4.
    // A warp walks its positions in steps of 32: thread = channel converts and scales two 16-position halves into
    // a 32 x 32 transposition tile, then lane = position applies activation / residual and stores one 128-byte
    // row segment per channel.
    const int quarter = warp & 3;
    const int half = warp < 4 ? 0 : 1;
    const int ewarp = quarter + 4 * half;
    const int cq = P.creal >> 5;                        // quarters holding distinct channels (2 or 4)
    const int chb = 32 * (quarter % cq);                // first channel (within the tile) of this warp
    const int nparts = 8 / cq;                          // position parts per tile (2 or 4)
    const int part = half * (nparts >> 1) + quarter / cq;
    const int tph = P.tp / nparts;                      // positions of this warp per item (multiple of 32)
    const int pbase = part * tph;
    const int spi = tph >> 5;                           // steps per item
    const bool has_res = epi.residual != nullptr;
    const long long cstride = (long long)g.ho * g.wo;
    float* const outt = reinterpret_cast<float*>(smem + P.smem_out) + (size_t)ewarp * 32 * kOutPitch;
    float2* const scl = reinterpret_cast<float2*>(smem + P.smem_scl) + (size_t)ewarp * tph;     // tph entries per warp

    // where does position `p` of tile `ptile` land in the output?  (offset of channel 0 of the channel tile, or -1)
    auto out_offset = [&](int ptile, int ctile, int p, int& sample) -> long long {
      const PosInfo pi = decode_pos(P, P.q_begin + (long long)ptile * P.tp + p);
      sample = pi.s;
      const bool valid = pi.in_range && (pi.s < g.n) && (pi.a >= 0) && (pi.a < g.ho) && (pi.col < g.wo);
      if (!valid) return -1;
      return (((long long)pi.s * P.cout + (long long)ctile * 128) * g.ho + pi.a) * g.wo + pi.col;
    };
    // exact int32 -> float for |i| < 2^22 with full-rate instructions (I2F issues at a quarter of the rate)
    auto i2f = [](uint32_t i) -> float { return __uint_as_float(0x4B400000u + i) - 12582912.0f; };
    Ring acc(kAccStages);
    float ws = 0.0f, bs = 0.0f;
    const int ppsub = 16 / P.npl;                       // positions per 16-column TMEM load
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      int ptile, ctile;
      if (P.n_ctiles == 1) { ptile = item; ctile = 0; }  // layers of up to 128 channels: no division
      else { ptile = item / P.n_ctiles; ctile = item - ptile * P.n_ctiles; }
      for (int step = 0; step < spi; ++step) {
        int sample;
        const long long off = out_offset(ptile, ctile, pbase + 32 * step + lane, sample);
        // the residual values of this lane's position (32 channels) are requested first and stay in registers: their
        // latency hides behind the accumulator conversion below (round 1 staged them through shared memory with
        // cp.async: one issue + one shared-memory read per element and a cursor to keep several steps in flight)
        float rs[32];
        if (has_res) {
          const float* rp = epi.residual + (off >= 0 ? off + (long long)chb * cstride : 0);
          const long long st = off >= 0 ? cstride : 0;
#pragma unroll
          for (int i = 0; i < 32; ++i) rs[i] = ldg_stream(rp + i * st);
        }
        if (step == 0) {
          // per-position activation scales of this warp's positions, per-thread channel constants
          for (int p = lane; p < tph; p += 32) {
            int smp = sample;
            const long long o2 = (p == lane) ? off : out_offset(ptile, ctile, pbase + p, smp);
            float2 s2 = make_float2(0.0f, 0.0f);
            if (o2 >= 0) {
              if (P.as_n) {
                s2.x = as_tab[smp];
                if (P.npl > 1) s2.y = as_tab[P.as_n + smp];
              } else {
                s2.x = __ldg(act_scales + smp);
                if (P.npl > 1) s2.y = __ldg(act_scales + g.n + smp);
              }
            }
            scl[p] = s2;
          }
          const float4 k = ctab[ctile * 128 + chb + lane];
          ws = k.x; bs = k.y;
          __syncwarp();
          mbar_wait_t(acc_full(acc.stage), acc.phase, err, 1, w0);
          tc_fence_after();
        }
        // ---- thread = channel: 16 accumulator columns at a time from TMEM, scale, bias -> transposition tile ----
        float* orow = outt + lane * kOutPitch;
        for (int sub = 0; sub < 2 * P.npl; ++sub) {
          const int pl0 = 32 * step + ppsub * sub;          // first position (within this warp's part) of the load
          const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)acc.stage * acc_cols +
                                 (uint32_t)((pbase + pl0) * P.npl);
          uint32_t rr[16];
          tmem_ld16(taddr, rr);
          const float4* sp = reinterpret_cast<const float4*>(scl + pl0);     // (s1, s2) of two positions per float4
          if (P.npl > 1) {
            float4 s4[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) s4[j] = sp[j];
            tmem_ld_wait();
            float o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float sx = (j & 1) ? s4[j >> 1].z : s4[j >> 1].x, sy = (j & 1) ? s4[j >> 1].w : s4[j >> 1].y;
              o[j] = fmaf(ws, fmaf(sy, i2f(rr[2 * j + 1]), sx * i2f(rr[2 * j])), bs);
            }
            reinterpret_cast<float4*>(orow + 8 * sub)[0] = make_float4(o[0], o[1], o[2], o[3]);
            reinterpret_cast<float4*>(orow + 8 * sub)[1] = make_float4(o[4], o[5], o[6], o[7]);
          } else {
            float4 s4[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) s4[j] = sp[j];
            tmem_ld_wait();
            float o[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float sx = (j & 1) ? s4[j >> 1].z : s4[j >> 1].x;
              o[j] = fmaf(ws, sx * i2f(rr[j]), bs);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
              reinterpret_cast<float4*>(orow + 16 * sub)[j] = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
          }
        }
        __syncwarp();
        // ---- lane = position: activation, residual, one 128-byte row segment per channel ----
        if (off >= 0) {
          const float* ot = outt + lane;
          const float4* tab = ctab + ctile * 128 + chb;
          float* yp = y + off + (long long)chb * cstride;
#pragma unroll
          for (int c0 = 0; c0 < 32; c0 += 16) {
            float v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = ot[(c0 + i) * kOutPitch];
            if (has_res && !epi.residual_after_act) {
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] += rs[c0 + i];
            }
            if (epi.act == 1) {
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.0f);
            } else if (epi.act == 2) {
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = v[i] >= 0.0f ? v[i] : v[i] * tab[c0 + i].z;
            }
            if (has_res && epi.residual_after_act) {
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] += rs[c0 + i];
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              *yp = v[i];
              yp += cstride;
            }
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
      Ring acc(kAccStages), rp(P.p_stages), rw(P.w_stages);
      const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((P.tp * P.npl) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const int ngroups = P.taps / P.tps;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        mbar_wait_t(acc_empty(acc.stage), acc.phase ^ 1u, err, 2, w0);
        tc_fence_after();
        const uint32_t d0 = tmem_base + (uint32_t)acc.stage * acc_cols;
        for (int cb = 0; cb < P.ncb; ++cb) {
          mbar_wait_t(p_full(rp.stage), rp.phase, err, 3, w1);
          tc_fence_after();
          const ulonglong2* const bt = reinterpret_cast<const ulonglong2*>(btab + rp.stage * (kMaxTaps * 2));
          for (int tg = 0; tg < ngroups; ++tg) {
            if (!P.w_resident || item == (int)blockIdx.x) {          // resident slabs: landed during the first item
              mbar_wait_t(w_full(rw.stage), rw.phase, err, 4, w2);
              tc_fence_after();
            }
            const ulonglong2* const at = reinterpret_cast<const ulonglong2*>(atab + rw.stage * 6);
            if (P.tps == 3) {
#pragma unroll
              for (int tt = 0; tt < 3; ++tt) {
                const ulonglong2 ad = at[tt], bd = bt[tg * 3 + tt];
                umma_i8(d0, ad.x, bd.x, idesc, (cb | tg | tt) != 0 ? 1u : 0u);
                umma_i8(d0, ad.y, bd.y, idesc, 1u);
              }
            } else {
              const ulonglong2 ad = at[0], bd = bt[tg];
              umma_i8(d0, ad.x, bd.x, idesc, (cb | tg) != 0 ? 1u : 0u);
              umma_i8(d0, ad.y, bd.y, idesc, 1u);
            }
            if (!P.w_resident) umma_commit(w_empty(rw.stage));
            rw.advance();
          }
          umma_commit(p_empty(rp.stage));
          rp.advance();
        }
        umma_commit(acc_full(acc.stage));
        acc.advance();
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    // ===================== weight loader =====================
    Ring rw(P.w_stages);
    if (lane == 0) {
      const int ngroups = P.taps / P.tps;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        if (P.w_resident && item != (int)blockIdx.x) break;         // loaded once
        const int ctile = item % P.n_ctiles;
        const int8_t* wsrc = wi8 + (long long)ctile * P.ncb * P.taps * P.w_slab_bytes;
        for (int cb = 0; cb < P.ncb; ++cb)
          for (int tg = 0; tg < ngroups; ++tg) {
            mbar_wait_t(w_empty(rw.stage), rw.phase ^ 1u, err, 5, w0);
            mbar_expect_tx(w_full(rw.stage), P.w_stage_bytes);
            bulk_g2s(sbase + P.smem_w + (uint32_t)rw.stage * P.w_stage_bytes,
                     wsrc + ((long long)cb * P.taps + tg * P.tps) * P.w_slab_bytes, P.w_stage_bytes, w_full(rw.stage));
            rw.advance();
          }
      }
    }
  } else {
    // ===================== patch producers (128 threads) =====================
    Ring rp(P.p_stages);
    const int pt = (warp < 10 ? threadIdx.x - 6 * 32 : threadIdx.x - 14 * 32 + 128);     // warps 6-9 and 14-17
    const int ntask = P.npl * g.nphase * P.pp;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int ptile = item / P.n_ctiles;
      const long long q0 = P.q_begin + (long long)ptile * P.tp;
      for (int cb = 0; cb < P.ncb; ++cb) {
        mbar_wait_t(p_empty(rp.stage), rp.phase ^ 1u, err, 6, w0);
        unsigned char* p_stage = smem + P.smem_p + (size_t)rp.stage * P.p_stage_bytes;
        constexpr int kPB = 4;   // tasks whose plane-bit loads are in flight together
        for (int task0 = pt; task0 < ntask; task0 += kProdThreads * kPB) {
          uint2 bits[kPB];
          uint32_t vm[kPB];
          unsigned char* dst[kPB];
#pragma unroll
          for (int u = 0; u < kPB; ++u) {
            const int task = task0 + u * kProdThreads;
            bits[u] = make_uint2(0u, 0u);
            vm[u] = 0u;
            dst[u] = nullptr;
            if (task < ntask) {
              // task -> (phase, position, plane) with the plane fastest: consecutive lanes write consecutive
              // 16-byte rows of the patch (position-fastest lanes stored at a 32-byte stride: 2-way bank conflicts)
              const int phase = task / (P.pp * P.npl);
              const int rem = task - phase * (P.pp * P.npl);
              const int pos = (P.npl == 2) ? (rem >> 1) : rem, pl = (P.npl == 2) ? (rem & 1) : 0;
              const int pp_idx = pl * g.nphase + phase;
              const long long q = q0 + P.dmin[phase] + pos;
              const PosInfo pi = decode_pos(P, q);
              int hv_p = g.h, wv_p = g.w;
              if (g.nphase == 4) { hv_p = (g.h - (phase >> 1) + 1) >> 1; wv_p = (g.w - (phase & 1) + 1) >> 1; }
              const bool valid = pi.in_range && (pi.s < g.n) && (pi.a >= 0) && (pi.a < hv_p) && (pi.col < wv_p);
              if (valid) {
                bits[u] = __ldg(reinterpret_cast<const uint2*>(planes + ((long long)pp_idx * g.vtot + q) * g.cw + cb * 2));
                vm[u] = 0xFFFFFFFFu;
              }
              // row = position * planes + plane inside the phase's [16-channel chunk][row][16 B] block
              dst[u] = p_stage + (size_t)phase * P.phase_bytes + ((size_t)pos * P.npl + pl) * 16;
            }
          }
#pragma unroll
          for (int u = 0; u < kPB; ++u) {
            if (dst[u] == nullptr) continue;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint32_t w = (j < 2) ? bits[u].x : bits[u].y;
              const uint32_t h16 = (w >> ((j & 1) * 16)) & 0xFFFFu;
              uint4 v;
              v.x = expand4(h16 & 0xFu) & vm[u];
              v.y = expand4((h16 >> 4) & 0xFu) & vm[u];
              v.z = expand4((h16 >> 8) & 0xFu) & vm[u];
              v.w = expand4((h16 >> 12) & 0xFu) & vm[u];
              *reinterpret_cast<uint4*>(dst[u] + (size_t)j * P.lbo_p) = v;
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
  if (diag && lane == 0 && blockIdx.x == 0) {
    long long* d = diag + warp * 4;
    d[0] = clock64() - t_start; d[1] = w0; d[2] = w1; d[3] = w2;
  }
#else
  (void)t_start; (void)w0; (void)w1; (void)w2; (void)diag;
#endif
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, 512);
}

// Shared-memory plan for a tile of `tp` positions; returns false when it does not fit.
static bool tc_plan(const lsq_act_geom* g, int nplanes, int cout, bool has_res, int tp, TcParams& P) {
  P.g = to_dev(*g);
  P.npl = nplanes; P.cout = cout; P.creal = cout < 128 ? cout : 128; P.n_ctiles = (cout + 127) / 128;
  P.ncb = g->c / 64; P.taps = g->kh * g->kw;
  P.tp = tp;
  if (tp / (8 / (P.creal >> 5)) < 32) return false;       // every epilogue warp takes whole 32-position steps
  // per-tap phase and position offset: input coordinate = stride*out + d - pad
  auto fdiv = [](int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); };
  for (int ph = 0; ph < 4; ++ph) P.dmin[ph] = 0;
  int off[kMaxTaps];
  bool seen[4] = {false, false, false, false};
  int dmax[4] = {0, 0, 0, 0};
  for (int dy = 0; dy < g->kh; ++dy)
    for (int dx = 0; dx < g->kw; ++dx) {
      const int t = dy * g->kw + dx;
      const int ey = dy - g->pad, ex = dx - g->pad;
      const int qy = fdiv(ey, g->stride), qx = fdiv(ex, g->stride);
      const int py = ey - qy * g->stride, px = ex - qx * g->stride;
      const int phase = g->nphase == 4 ? (py * 2 + px) : 0;
      P.tap_phase[t] = phase;
      off[t] = qy * g->pitch + qx;
      if (!seen[phase]) { seen[phase] = true; P.dmin[phase] = off[t]; dmax[phase] = off[t]; }
      if (off[t] < P.dmin[phase]) P.dmin[phase] = off[t];
      if (off[t] > dmax[phase]) dmax[phase] = off[t];
    }
  int span = 0;
  for (int ph = 0; ph < g->nphase; ++ph)
    if (seen[ph] && dmax[ph] - P.dmin[ph] > span) span = dmax[ph] - P.dmin[ph];
  for (int t = 0; t < P.taps; ++t) P.tap_off[t] = off[t] - P.dmin[P.tap_phase[t]];
  for (int t = P.taps; t < kMaxTaps; ++t) { P.tap_phase[t] = 0; P.tap_off[t] = 0; }
  P.pp = (tp + span + 7) / 8 * 8;
  P.lbo_p = (uint32_t)P.pp * nplanes * 16u;
  if (P.lbo_p > 0x3FFFu * 16u) return false;
  P.phase_bytes = 4u * P.lbo_p;
  P.p_stage_bytes = (uint32_t)g->nphase * P.phase_bytes;
  P.w_slab_bytes = 128u * 64u;
  const size_t total = 225 * 1024;
  const size_t bar_bytes = 512 + (kMaxWStages * 3 * 2 + kMaxPStages * kMaxTaps * 2) * 8, tab_bytes = (size_t)cout * 16;
  const size_t scl_bytes = (size_t)8 * (tp / (8 / (P.creal >> 5))) * sizeof(float2), out_bytes = 8 * 32 * kOutPitch * sizeof(float);
  const size_t slack = 0;
  P.as_n = (size_t)g->n * nplanes * sizeof(float) <= 8192 ? g->n : 0;       // up to 1024 samples x 2 planes
  const size_t as_bytes = ((size_t)P.as_n * nplanes * sizeof(float) + 15) & ~(size_t)15;
  const size_t fixed = bar_bytes + tab_bytes + scl_bytes + out_bytes + as_bytes + slack;
  // minimum: 2 patch stages (1 when there is a single channel block and nothing to overlap with is no option:
  // the next item's patch is built while this one is multiplied), 2 weight stages of one tap, 2 residual steps
  const size_t res_min = has_res ? (size_t)8 * 2 * 4096 : 0;
  if (fixed + 2 * (size_t)P.p_stage_bytes + 2 * (size_t)P.w_slab_bytes + res_min > total) return false;
  size_t left = total - fixed - res_min - 2 * (size_t)P.p_stage_bytes;
  // weights: the whole layer when it fits (one channel tile: every item uses the same slabs -- streaming them again for
  // every position tile costs ~1 GB of L2 -> shared-memory traffic per layer); else a ring of whole tap rows per stage when
  // they fit twice (fewer commits), else of single taps
  const size_t all_w = (size_t)P.ncb * P.taps * P.w_slab_bytes;
  static const bool no_resident = getenv("LSQ_BCONV_STREAM_W") != nullptr;       // development: force the ring
  P.w_resident = (!no_resident && P.n_ctiles == 1 && P.taps % g->kw == 0 && P.ncb * (P.taps / g->kw) <= kMaxWStages && all_w <= left) ? 1 : 0;
  if (P.w_resident) {
    P.tps = g->kw;
    P.w_stage_bytes = (uint32_t)P.tps * P.w_slab_bytes;
    P.w_stages = P.ncb * (P.taps / g->kw);
    left -= all_w;
  } else {
    P.tps = (P.taps % g->kw == 0 && 2 * (size_t)g->kw * P.w_slab_bytes <= left) ? g->kw : 1;
    P.w_stage_bytes = (uint32_t)P.tps * P.w_slab_bytes;
    P.w_stages = 2;
    left -= 2 * (size_t)P.w_stage_bytes;
  }
  // then: deeper residual staging, a third patch stage, more weight stages
  P.r_stages = has_res ? 2 : 1;
  if (has_res) {
    const int want[2] = {kResStages, 3};
    for (int k = 0; k < 2; ++k) {
      const size_t extra = (size_t)8 * (want[k] - 2) * 4096;
      if (extra <= left) { P.r_stages = want[k]; left -= extra; break; }
    }
  }
  P.p_stages = 2;
  if (P.ncb >= 3 && P.p_stage_bytes <= left) { P.p_stages = 3; left -= P.p_stage_bytes; }
  while (!P.w_resident && P.w_stages < kMaxWStages && P.w_stages * P.tps < 2 * P.taps && P.w_stage_bytes <= left) { ++P.w_stages; left -= P.w_stage_bytes; }
  uint32_t o = 0;
  P.smem_p = o; o += (uint32_t)P.p_stages * P.p_stage_bytes; o = (o + 127u) / 128u * 128u;
  P.smem_w = o; o += (uint32_t)P.w_stages * P.w_stage_bytes + (uint32_t)slack;
  P.smem_bar = o; o += (uint32_t)bar_bytes;
  P.smem_tab = o; o += (uint32_t)tab_bytes;
  P.smem_scl = o; o += (uint32_t)scl_bytes;
  P.smem_out = o; o += (uint32_t)out_bytes;
  P.smem_res = o; o += has_res ? (uint32_t)(8 * P.r_stages * 4096) : 0u;
  P.smem_as = o; o += (uint32_t)as_bytes;
  return o <= 227 * 1024;
}

static int tc_pick_tile(const lsq_act_geom* g, int nplanes, int cout, bool has_res, TcParams& P) {
  for (int tp = 256 / nplanes; tp >= 64; tp >>= 1)
    if (tc_plan(g, nplanes, cout, has_res, tp, P)) return tp;
  return 0;
}

bool bconv2d_tc_supported(const lsq_act_geom* g, int nplanes, int cout) {
  if (nplanes < 1 || nplanes > 2) return false;
  if (g->c % 64 != 0 || cout % 64 != 0) return false;
  if (cout > 128 && cout % 128 != 0) return false;
  if (g->kh * g->kw > kMaxTaps) return false;
  if (g->stride != 1 && g->stride != 2) return false;
  if ((long long)g->n * g->rows_per_sample * g->pitch > (1ll << 31) - 4096) return false;
  TcParams P;
  return tc_pick_tile(g, nplanes, cout, false, P) != 0;
}

int bconv2d_tc_launch(const uint32_t* d_planes, const lsq_act_geom* g, int nplanes, const float* d_act_scales,
                      const void* d_wpack, const float* d_w_scale, const float* d_bias, int cout, float* d_y,
                      const Epilogue& epi, cudaStream_t stream) {
  TcParams P;
  if (tc_pick_tile(g, nplanes, cout, false, P) == 0) {      // residual values travel through registers: no staging area
    set_error("bconv2d_tc: patch does not fit shared memory");
    return LSQ_ERR_UNSUPPORTED;
  }
  P.pitch_magic = ((1ull << 40) + (unsigned long long)g->pitch - 1ull) / (unsigned long long)g->pitch;
  P.rps_magic = ((1ull << 40) + (unsigned long long)g->rows_per_sample - 1ull) / (unsigned long long)g->rows_per_sample;
  P.q_begin = (long long)g->lead + (long long)g->ph * g->pitch;
  const long long qspan = (long long)g->n * g->rows_per_sample * g->pitch;
  P.p_tiles = (int)((qspan + P.tp - 1) / P.tp);
  const int n_items = P.p_tiles * P.n_ctiles;
  const size_t smem_bytes = (size_t)P.smem_as + (((size_t)P.as_n * P.npl * sizeof(float) + 15) & ~(size_t)15);

  // the weight image follows the bit image inside d_wpack (lsq_bconv.cu)
  const size_t bits_bytes = ((size_t)cout * g->kh * g->kw * g->cw * 4 + 1023) / 1024 * 1024;
  const int8_t* wi8 = (const int8_t*)d_wpack + bits_bytes;

  static std::atomic<unsigned long long> smem_set{0ull};
  const cudaError_t e = ensure_max_smem(bconv_tc_kernel, smem_set);
  if (e != cudaSuccess) { set_error("bconv2d_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return LSQ_ERR_CUDA; }
  const int sms = device_sms();
  const int grid = n_items < sms ? n_items : sms;
  long long* d_diag = nullptr;
#ifdef LSQ_TC_DIAG
  static const bool want_diag = getenv("LSQ_TC_DIAG") != nullptr;     // development build: per-role wait cycles of CTA 0
  if (want_diag) { cudaMalloc(&d_diag, 22 * 4 * sizeof(long long)); cudaMemsetAsync(d_diag, 0, 22 * 4 * sizeof(long long), stream); }
#endif
  bconv_tc_kernel<<<grid, kTcThreads, smem_bytes, stream>>>(d_planes, P, d_act_scales, wi8, d_w_scale, d_bias, d_y, nullptr, epi, d_diag);
  LSQ_CUDA_LAUNCH_CHECK("bconv_tc_kernel");
#ifdef LSQ_TC_DIAG
  if (want_diag) {
    long long h[22 * 4];
    cudaStreamSynchronize(stream);
    cudaMemcpy(h, d_diag, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(d_diag);
    const char* role[22] = {"epi0", "epi1", "epi2", "epi3", "mma", "wload", "prod0", "prod1", "prod2", "prod3", "epi4", "epi5", "epi6", "epi7",
                            "prod4", "prod5", "prod6", "prod7", "prod8", "prod9", "prod10", "prod11"};
    fprintf(stderr, "[bconv_tc diag] cin %d cout %d %dx%d stride %d items %d grid %d tp %d pp %d p_stages %d w_stages %d x %d taps r_stages %d smem %zu\n",
            g->c, cout, g->h, g->w, g->stride, n_items, grid, P.tp, P.pp, P.p_stages, P.w_stages, P.tps, P.r_stages, smem_bytes);
    for (int w = 0; w < kTcThreads / 32; ++w)
      fprintf(stderr, "   %-6s total %9lld  wait0 %9lld  wait1(p_full) %9lld  wait2(w_full) %9lld\n", role[w], h[w * 4], h[w * 4 + 1], h[w * 4 + 2], h[w * 4 + 3]);
  }
#endif
  return LSQ_OK;
}

}  // namespace lsq
