// Binary convolution on the 5th-generation tensor cores (tcgen05.mma kind::i8, sm_100a).
//
// sm_100a has no 1-bit MMA; the +-1 planes are expanded to int8 on
// the fly and contracted with int8 +-1 weights, int32 accumulators in TMEM -- exact integers.
//
// Implicit GEMM without im2col:  D[128 positions, NT out-channels] += A_tap[128, 64] * B_tap[64, NT]
// summed over 64-channel blocks and taps.  The activation planes live in HBM as bits in a "virtual
// raster" (lsq_b200.h) in which a tap is a uniform shift of the position index, so ONE expanded
// int8 patch per channel block in shared memory (tile + halo, zero at padding positions) serves all
// kh*kw taps: the A operand of tap t is the same patch read through a shared-memory matrix
// descriptor whose start address is shifted by d(t) rows (K-major, no swizzle: a row is 16 bytes per
// 16-channel chunk, 8-row core matrices are contiguous, so any row shift is a 16-byte offset).
//
// Warp roles (320 threads, one CTA per SM, persistent over (m-tile, n-tile) items):
//   warps 0-3  epilogue : tcgen05.ld accumulators -> vw[c]*(s1[n]*I1 + s2[n]*I2) + bias[c] -> NCHW fp32
//   warp  4    MMA      : one elected thread issues tcgen05.mma / tcgen05.commit; owns TMEM alloc
//   warp  5    weights  : cp.async.bulk (TMA engine, 1-D) of pre-packed operand slabs, mbarrier tx
//   warps 6-9  patches  : plane bits (L2) -> int8 patch in shared memory, fence.proxy.async
// Pipelines: patch ring (per channel block), weight ring (per channel block x tap), 2 accumulator
// stages in TMEM so the epilogue of item i overlaps the MMAs of item i+1.
#include "lsq_common.cuh"

namespace lsq {

constexpr int kTcThreads = 448;   // 4 epilogue + MMA + loader + 4 producer + 4 more epilogue warps
constexpr int kTileM = 128;
constexpr int kMaxTaps = 9;
constexpr int kBStages = 4;
constexpr int kAccStages = 2;
constexpr int kMaxAStages = 3;
constexpr unsigned long long kWatchdogCycles = 4000000000ull;  // ~2 s: trap instead of hanging

struct TcParams {
  ActGeom g;
  int npl, cout, nt, n_ntiles, ncb, taps;
  int m_tiles;
  long long q_begin;             // first output position
  int pp;                        // patch positions per (plane, phase)
  int a_stages;
  int dmin[4];                   // per phase: smallest tap offset (positions)
  int tap_phase[kMaxTaps];
  int tap_off[kMaxTaps];         // tap offset relative to dmin of its phase (>= 0)
  uint32_t a_stage_bytes, b_stage_bytes;
  uint32_t smem_a, smem_b, smem_bar;  // offsets in dynamic smem
  unsigned long long pitch_magic, rps_magic;   // ceil(2^40 / d): x / d == (x * magic) >> 40 for x * d < 2^40
};

// ---- PTX wrappers -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* err, int code) {
  if (mbar_try(bar, parity)) return;
  const unsigned long long t0 = clock64();
  while (!mbar_try(bar, parity)) {
    if (clock64() - t0 > kWatchdogCycles) {
      if (err) atomicExch(err, code);
      __threadfence_system();
      asm volatile("trap;");
    }
  }
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, SWIZZLE_NONE shared-memory matrix descriptor (sm_100 format: version 1 at bit 46)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}

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

struct Ring {
  int stage, n;
  uint32_t phase;
  __device__ Ring(int n_) : stage(0), n(n_), phase(0) {}
  __device__ void advance() {
    if (++stage == n) { stage = 0; phase ^= 1u; }
  }
};

__global__ void __launch_bounds__(kTcThreads, 1)
bconv_tc_kernel(const uint32_t* __restrict__ planes, TcParams P, const float* __restrict__ act_scales,
                const int8_t* __restrict__ wi8, const float* __restrict__ w_scale, const float* __restrict__ bias,
                float* __restrict__ y, int* __restrict__ err, Epilogue epi) {
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const ActGeom& g = P.g;

  // barrier block: a_full[kMaxAStages] a_empty[kMaxAStages] b_full[kBStages] b_empty[kBStages]
  //                acc_full[2] acc_empty[2] | tmem base
  const uint32_t bar0 = sbase + P.smem_bar;
  auto a_full = [&](int s) { return bar0 + 8u * s; };
  auto a_empty = [&](int s) { return bar0 + 8u * (kMaxAStages + s); };
  auto b_full = [&](int s) { return bar0 + 8u * (2 * kMaxAStages + s); };
  auto b_empty = [&](int s) { return bar0 + 8u * (2 * kMaxAStages + kBStages + s); };
  auto acc_full = [&](int s) { return bar0 + 8u * (2 * kMaxAStages + 2 * kBStages + s); };
  auto acc_empty = [&](int s) { return bar0 + 8u * (2 * kMaxAStages + 2 * kBStages + kAccStages + s); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + P.smem_bar + 8u * (2 * kMaxAStages + 2 * kBStages + 2 * kAccStages));

  if (threadIdx.x == 0) {
    for (int s = 0; s < kMaxAStages; ++s) { mbar_init(a_full(s), 4); mbar_init(a_empty(s), 1); }
    for (int s = 0; s < kBStages; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1); }
    for (int s = 0; s < kAccStages; ++s) { mbar_init(acc_full(s), 1); mbar_init(acc_empty(s), 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) tmem_alloc(smem_u32((const void*)tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int n_items = P.m_tiles * P.n_ntiles;
  const uint32_t acc_cols = (uint32_t)(P.npl * P.nt);

  if (warp < 4 || warp >= 10) {
    // ===================== epilogue (8 warps: two per TMEM lane quarter) =====================
    // Warp w reads TMEM lanes 32*(w%4)..+31 (= tile rows); the two warps of a quarter split the channels
    // in interleaved groups of 32.  Residual values are prefetched one group ahead, across tiles, so the
    // loads of the next group are in flight while this one is scaled and stored.
    const int quarter = warp & 3;
    const int egrp = warp < 4 ? 0 : 1;
    const int ngroups = P.nt / 64;                     // groups of 32 channels handled by this warp per item
    const bool has_res = epi.residual != nullptr;
    const long long cstride = (long long)g.ho * g.wo;
    struct Row { bool valid; long long yoff; float sc0, sc1; int ntile; };
    auto decode_row = [&](int item) {
      Row r;
      const int mt = item / P.n_ntiles;
      r.ntile = item - mt * P.n_ntiles;
      const PosInfo pi = decode_pos(P, P.q_begin + (long long)mt * kTileM + quarter * 32 + lane);
      r.valid = pi.in_range && (pi.s < g.n) && (pi.a >= 0) && (pi.a < g.ho) && (pi.col < g.wo);
      r.sc0 = 0.0f; r.sc1 = 0.0f;
      if (r.valid) {
        r.sc0 = __ldg(act_scales + pi.s);
        if (P.npl > 1) r.sc1 = __ldg(act_scales + g.n + pi.s);
      }
      r.yoff = (((long long)pi.s * P.cout + (long long)r.ntile * P.nt) * g.ho + pi.a) * g.wo + pi.col;
      return r;
    };
    auto load_res = [&](float (&dst)[32], const Row& r, int c0) {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        dst[j] = (has_res && r.valid) ? __ldg(epi.residual + r.yoff + (long long)(c0 + j) * cstride) : 0.0f;
    };
    Ring acc(kAccStages);
    int item = blockIdx.x;
    if (item < n_items) {
      Row cur = decode_row(item);
      float res[32];
      load_res(res, cur, egrp * 32);
      for (; item < n_items; item += gridDim.x) {
        Row nxt = cur;
        const uint32_t tbase = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)acc.stage * acc_cols;
        for (int gi = 0; gi < ngroups; ++gi) {
          const int c0 = egrp * 32 + gi * 64;
          float nres[32];
          if (gi + 1 < ngroups) {
            load_res(nres, cur, c0 + 64);
          } else if (item + (int)gridDim.x < n_items) {
            nxt = decode_row(item + (int)gridDim.x);
            load_res(nres, nxt, egrp * 32);
          }
          if (gi == 0) {
            mbar_wait(acc_full(acc.stage), acc.phase, err, 1);
            tc_fence_after();
          }
          float* yrow = y + cur.yoff;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t r0[16], r1[16];
            tmem_ld16(tbase + (uint32_t)(c0 + 16 * h), r0);
            if (P.npl > 1) tmem_ld16(tbase + (uint32_t)(P.nt + c0 + 16 * h), r1);
            tmem_ld_wait();
            if (cur.valid) {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const int c = cur.ntile * P.nt + c0 + 16 * h + j;
                float t = cur.sc0 * (float)(int)r0[j];
                if (P.npl > 1) t = fmaf(cur.sc1, (float)(int)r1[j], t);
                float o = __ldg(w_scale + c) * t;
                if (bias) o += __ldg(bias + c);
                yrow[(long long)(c0 + 16 * h + j) * cstride] = apply_epilogue_res(epi, o, c, res[16 * h + j]);
              }
            }
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) res[j] = nres[j];
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty(acc.stage));
        acc.advance();
        cur = nxt;
      }
    }
  } else if (warp == 4) {
    // ===================== MMA issuer =====================
    Ring acc(kAccStages), ra(P.a_stages), rb(kBStages);
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(P.nt >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
    const uint32_t lbo_a = (uint32_t)P.pp * 16u, lbo_b = (uint32_t)P.nt * 16u;
    const uint32_t plane_phase_bytes = (uint32_t)P.pp * 64u;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      mbar_wait(acc_empty(acc.stage), acc.phase ^ 1u, err, 2);
      tc_fence_after();
      const uint32_t d0 = tmem_base + (uint32_t)acc.stage * acc_cols;
      for (int cb = 0; cb < P.ncb; ++cb) {
        mbar_wait(a_full(ra.stage), ra.phase, err, 3);
        tc_fence_after();
        const uint32_t a_stage = sbase + P.smem_a + (uint32_t)ra.stage * P.a_stage_bytes;
        for (int tap = 0; tap < P.taps; ++tap) {
          mbar_wait(b_full(rb.stage), rb.phase, err, 4);
          tc_fence_after();
          if (lane == 0) {
            const uint32_t b_stage = sbase + P.smem_b + (uint32_t)rb.stage * P.b_stage_bytes;
            for (int pl = 0; pl < P.npl; ++pl) {
              const uint32_t a_addr = a_stage + (uint32_t)(pl * g.nphase + P.tap_phase[tap]) * plane_phase_bytes +
                                      (uint32_t)P.tap_off[tap] * 16u;
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const uint64_t ad = make_desc(a_addr + (uint32_t)(2 * h) * lbo_a, lbo_a, 128u);
                const uint64_t bd = make_desc(b_stage + (uint32_t)(2 * h) * lbo_b, lbo_b, 128u);
                umma_i8(d0 + (uint32_t)(pl * P.nt), ad, bd, idesc, (cb | tap | h) != 0 ? 1u : 0u);
              }
            }
            umma_commit(b_empty(rb.stage));
            if (tap == P.taps - 1) umma_commit(a_empty(ra.stage));
            if (tap == P.taps - 1 && cb == P.ncb - 1) umma_commit(acc_full(acc.stage));
          }
          __syncwarp();
          rb.advance();
        }
        ra.advance();
      }
      acc.advance();
    }
  } else if (warp == 5) {
    // ===================== weight loader =====================
    Ring rb(kBStages);
    if (lane == 0) {
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int ntile = item % P.n_ntiles;
        const int8_t* wsrc = wi8 + (long long)ntile * P.ncb * P.taps * P.b_stage_bytes;
        for (int cb = 0; cb < P.ncb; ++cb)
          for (int tap = 0; tap < P.taps; ++tap) {
            mbar_wait(b_empty(rb.stage), rb.phase ^ 1u, err, 5);
            mbar_expect_tx(b_full(rb.stage), P.b_stage_bytes);
            bulk_g2s(sbase + P.smem_b + (uint32_t)rb.stage * P.b_stage_bytes,
                     wsrc + ((long long)cb * P.taps + tap) * P.b_stage_bytes, P.b_stage_bytes, b_full(rb.stage));
            rb.advance();
          }
      }
    }
  } else {
    // ===================== patch producers (128 threads) =====================
    Ring ra(P.a_stages);
    const int pt = threadIdx.x - 6 * 32;
    const int ntask = P.npl * g.nphase * P.pp;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int mt = item / P.n_ntiles;
      const long long q0 = P.q_begin + (long long)mt * kTileM;
      for (int cb = 0; cb < P.ncb; ++cb) {
        mbar_wait(a_empty(ra.stage), ra.phase ^ 1u, err, 6);
        unsigned char* a_stage = smem + P.smem_a + (size_t)ra.stage * P.a_stage_bytes;
        constexpr int kPB = 4;   // tasks whose plane-bit loads are in flight together
        for (int task0 = pt; task0 < ntask; task0 += 128 * kPB) {
          uint2 bits[kPB];
          uint32_t vm[kPB];
          unsigned char* dst[kPB];
#pragma unroll
          for (int u = 0; u < kPB; ++u) {
            const int task = task0 + u * 128;
            bits[u] = make_uint2(0u, 0u);
            vm[u] = 0u;
            dst[u] = nullptr;
            if (task < ntask) {
              const int pp_idx = task / P.pp;          // pl * nphase + phase
              const int pos = task - pp_idx * P.pp;
              const int phase = pp_idx % g.nphase;
              const long long q = q0 + P.dmin[phase] + pos;
              const PosInfo pi = decode_pos(P, q);
              int hv_p = g.h, wv_p = g.w;
              if (g.nphase == 4) { hv_p = (g.h - (phase >> 1) + 1) >> 1; wv_p = (g.w - (phase & 1) + 1) >> 1; }
              const bool valid = pi.in_range && (pi.s < g.n) && (pi.a >= 0) && (pi.a < hv_p) && (pi.col < wv_p);
              if (valid) {
                bits[u] = __ldg(reinterpret_cast<const uint2*>(planes + ((long long)pp_idx * g.vtot + q) * g.cw + cb * 2));
                vm[u] = 0xFFFFFFFFu;
              }
              dst[u] = a_stage + (size_t)pp_idx * ((size_t)P.pp * 64) + (size_t)pos * 16;
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
              *reinterpret_cast<uint4*>(dst[u] + (size_t)j * ((size_t)P.pp * 16)) = v;
            }
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(a_full(ra.stage));
        ra.advance();
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, 512);
}

bool bconv2d_tc_supported(const lsq_act_geom* g, int nplanes, int cout) {
  if (nplanes < 1 || nplanes > 2) return false;
  if (g->c % 64 != 0 || cout % 64 != 0) return false;
  if (cout > 128 && cout % 128 != 0) return false;
  if (g->kh * g->kw > kMaxTaps) return false;
  if (g->stride != 1 && g->stride != 2) return false;
  // patch + weight ring must fit shared memory
  const int span = g->ph * g->pitch + g->ph;  // |tap offset| bound in one phase
  const int pp = kTileM + 2 * span + 2;
  const size_t a_stage = (size_t)nplanes * g->nphase * pp * 64;
  const int nt = cout < 128 ? cout : 128;
  const size_t b = (size_t)kBStages * nt * 64;
  if (a_stage + b + 1024 > 220 * 1024) return false;
  if ((size_t)pp * 16 > 0x3FFF * 16) return false;
  if ((long long)g->n * g->rows_per_sample * g->pitch > (1ll << 31) - 4096) return false;
  return true;
}

int bconv2d_tc_launch(const uint32_t* d_planes, const lsq_act_geom* g, int nplanes, const float* d_act_scales,
                      const void* d_wpack, const float* d_w_scale, const float* d_bias, int cout, float* d_y,
                      const Epilogue& epi, cudaStream_t stream) {
  TcParams P;
  P.g = to_dev(*g);
  P.npl = nplanes; P.cout = cout; P.nt = cout < 128 ? cout : 128; P.n_ntiles = cout / P.nt;
  P.ncb = g->c / 64; P.taps = g->kh * g->kw;
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
  P.pp = (kTileM + span + 7) / 8 * 8;
  P.a_stage_bytes = (uint32_t)((size_t)nplanes * g->nphase * P.pp * 64);
  P.b_stage_bytes = (uint32_t)(P.nt * 64);
  const size_t budget = 220 * 1024;
  const size_t fixed = (size_t)kBStages * P.b_stage_bytes + 1024;
  int a_st = (int)((budget - fixed) / P.a_stage_bytes);
  if (a_st < 1) { set_error("bconv2d_tc: patch does not fit shared memory"); return LSQ_ERR_UNSUPPORTED; }
  if (a_st > kMaxAStages) a_st = kMaxAStages;
  if (a_st > P.ncb && P.ncb >= 1) a_st = P.ncb < 2 ? 2 : P.ncb;  // more stages than blocks buys nothing
  if (a_st > kMaxAStages) a_st = kMaxAStages;
  if ((size_t)a_st * P.a_stage_bytes + fixed > budget) a_st = (int)((budget - fixed) / P.a_stage_bytes);
  P.a_stages = a_st;
  P.smem_a = 0;
  P.smem_b = (uint32_t)((size_t)a_st * P.a_stage_bytes);
  P.smem_b = (P.smem_b + 127u) / 128u * 128u;
  P.smem_bar = P.smem_b + (uint32_t)kBStages * P.b_stage_bytes;
  const size_t smem_bytes = (size_t)P.smem_bar + 256;
  P.pitch_magic = ((1ull << 40) + (unsigned long long)g->pitch - 1ull) / (unsigned long long)g->pitch;
  P.rps_magic = ((1ull << 40) + (unsigned long long)g->rows_per_sample - 1ull) / (unsigned long long)g->rows_per_sample;
  P.q_begin = (long long)g->lead + (long long)g->ph * g->pitch;
  const long long qspan = (long long)g->n * g->rows_per_sample * g->pitch;
  P.m_tiles = (int)((qspan + kTileM - 1) / kTileM);
  const int n_items = P.m_tiles * P.n_ntiles;

  // the weight image follows the bit image inside d_wpack (lsq_bconv.cu)
  const size_t bits_bytes = ((size_t)cout * g->kh * g->kw * g->cw * 4 + 1023) / 1024 * 1024;
  const int8_t* wi8 = (const int8_t*)d_wpack + bits_bytes;

  cudaError_t e = cudaFuncSetAttribute(bconv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
  if (e != cudaSuccess) { set_error("bconv2d_tc: cudaFuncSetAttribute(%zu): %s", smem_bytes, cudaGetErrorString(e)); return LSQ_ERR_CUDA; }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = n_items < sms ? n_items : sms;
  bconv_tc_kernel<<<grid, kTcThreads, smem_bytes, stream>>>(d_planes, P, d_act_scales, wi8, d_w_scale, d_bias, d_y, nullptr, epi);
  LSQ_CUDA_LAUNCH_CHECK("bconv_tc_kernel");
  return LSQ_OK;
}

}  // namespace lsq
