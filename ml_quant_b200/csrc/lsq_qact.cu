// Fused activation quantizer for the 2-bit (ls-2) and ternary (ls-T) schemes: one kernel per QuantConv2d input.
//
// Replaces, on the QuantConv2d input (quant/binary/binary_conv.py:163 -> activation_quantization.py:99-100 ->
// quantization.py:59-115 -> optimal.py:121-155), the chain  clamp -> opt_v1 (sort, cumsum, masks, cost tensor,
// .tolist() sync) -> v2 = mean|x - v1 sign(x)| -> binarize / residual / binarize  and, fused in front of it, the
// eval-mode BatchNorm of the caller (quant/models/resnet.py:180-190).  Round 1 ran this as three kernels that each
// streamed the tensor from HBM (histogram pass, collection pass, encoder).  Here a row (= one sample) is handled by one
// CTA (default) or by one thread-block cluster (LSQ_QACT_MODE=cluster, see qact_shape below):
//
//   sweep 1 (HBM)  every 3rd element (optimal.py:134) of |clamp(bn(x))|: float bit patterns are monotone keys;
//                  bins of 2^14 keys anchored at the clamp bound (512 bins per octave, 8 octaves), count + exact
//                  integer sum of the low 14 key bits per bin (native shared-memory atomics, order independent);
//                  the bin index of every sampled element is kept (2 bytes per sample; shared memory or, for rows
//                  too long for that, a global scratch row that is read back L2-hot);
//   merge          (cluster only) the CTAs add up their histograms through distributed shared memory;
//   flag           prefix counts / EXACT prefix sums at the bin edges bound both threshold functions of
//                  optimal.py:63-80 over every bin (may_hold, conservative): a handful of bins can hold a candidate;
//   collect        walk the bin indices, queue the hits, re-read the few hundred flagged elements in one go and file
//                  their exact keys into one list segment per flagged bin;
//   solve          one warp per segment sorts it, gives every element the reference's own fp32 candidate test on
//                  (float) prefix sums and the closed-form cost in fp64; first minimum wins (torch.argmin) -> v1;
//   sweep 2        the encoder's work item (lsq_encode_core.cuh): both sign planes in the convolution's raster and
//                  sum |x - v1 sign(x)| -> v2 (fp64, fixed summation tree, deterministic and batch invariant).
//
// Rows the fast path cannot decide (a candidate may sit below the bin window, too many flagged elements, ...) are
// marked in d_row_status and redone by the generic kernels of lsq_solve.cu / lsq_quant.cu, launched behind this one on
// the marked rows only.
#include <cooperative_groups.h>
#include <stdlib.h>
#include "lsq_common.cuh"
#include "lsq_encode_core.cuh"
#include "lsq_solve_core.cuh"

namespace cg = cooperative_groups;

namespace lsq {

constexpr int kQThreads = 512;
constexpr int kQBins = 4096, kQShift = 14;
constexpr int kQIdsCap = 17408;                       // sampled elements of a row handled by one CTA
constexpr int kQListCap = 2048;                       // collected elements (sorted and tested one by one)
constexpr int kQMaxFlag = 64;
constexpr int kQMaxGroups = 128;
constexpr int kQMaxCluster = 8;
constexpr int kQMaxChannels = 512;
constexpr uint32_t kIdBelow = 0xFFFFu, kIdNone = 0xFFFEu;

struct QParams {
  long long len;          // elements per row (= c * h * w)
  uint32_t n_s;           // sampled elements per row: ceil(len / 3)
  uint32_t groups;        // 12-element groups per row: ceil(len / 12)
  uint32_t klo;           // first key of the bin window
  uint32_t hw, nq, nitems;
  unsigned long long hw_magic;   // ceil(2^40 / hw), or 0: divide exactly
};

constexpr int kQSegMax = 256;                         // elements of one flagged bin a warp sorts
constexpr int kQQueueTotal = kQBins;                  // hits a CTA may queue in the collection walk (split over its warps;
                                                      // the queue reuses the bin counters, dead by then)
template <int T, int IDS>
struct QSmemT {
  uint32_t hcnt[kQBins];
  union {
    uint32_t hrem[kQBins];         // sweep 1 .. flagging: sum of the low key bits per bin
    uint8_t fmap[kQBins];          // afterwards: bin -> 1 + flagged index, or 65 + index of the flagged bin it follows
  };
  uint16_t ids[IDS];
  uint32_t list[(T * 4 > kQListCap) ? T * 4 : kQListCap];   // collected keys, one power-of-two padded segment per flagged bin (layout shared by
                                   // all CTAs of the cluster); during the scan: per-thread group prefixes
  float2 ab[kQMaxChannels];
  double red[32];
  double wsum[32];
  uint32_t wcnt[32];
  uint32_t wfirst[32];
  // per-rank partial results, written into rank 0's copy through distributed shared memory
  double part_sb[kQMaxCluster];
  double part_v2[kQMaxCluster];
  uint32_t part_cb[kQMaxCluster], part_kmin[kQMaxCluster], part_kmax[kQMaxCluster];
  uint32_t kmin, kmax, cb;
  int nflag, ngroup;
  uint16_t glist[kQMaxGroups];
  // flagged bins as found (unordered) ...
  uint16_t fbin[kQMaxFlag];
  uint32_t fexcl[kQMaxFlag], fnb[kQMaxFlag];
  double fsumb[kQMaxFlag];
  int ford[kQMaxFlag];
  // ... and in ascending order, with their list segments
  uint16_t sbin[kQMaxFlag], snb[kQMaxFlag];       // bin, next non-empty bin (0xFFFF: none)
  uint32_t sexcl[kQMaxFlag], seg_start[kQMaxFlag], seg_cnt[kQMaxFlag];
  double spref[kQMaxFlag];
  uint32_t lfill[kQMaxFlag], lsmin[kQMaxFlag];    // this CTA: elements collected per segment, smallest key of the follower bin
  uint32_t nlist;
  int status;                      // 0 = solved here, != 0: reason the row goes to the generic kernels
  float v1;
  double s_tot;
  double best_cost[32];
  uint32_t best_pos[32], best_key[32], ncand;
};
static_assert(sizeof(QSmemT<kQThreads, kQIdsCap>) <= 113 * 1024, "two CTAs per SM");
static_assert(sizeof(QSmemT<256, 64>) <= 56 * 1024, "four CTAs per SM");

__device__ __forceinline__ uint32_t q_channel(const QParams& qp, unsigned long long idx) {
  return qp.hw_magic ? (uint32_t)((idx * qp.hw_magic) >> 40) : (uint32_t)(idx / qp.hw);
}

// exact sum of the `cnt` keys of bin b as an integer multiple of the bin's ulp 2^(e-150): all keys of a bin share the
// exponent e and the upper mantissa bits of the bin's first key; `rem` is the sum of their low 14 bits
__device__ __forceinline__ unsigned long long q_bin_isum(uint32_t klo, uint32_t b, uint32_t cnt, uint32_t rem) {
  const uint32_t kb = klo + (b << kQShift);
  return (unsigned long long)cnt * (unsigned long long)(0x800000u | (kb & 0x7FFFFFu)) + rem;
}
__device__ __forceinline__ double q_bin_scale(uint32_t klo, uint32_t b) {       // 2^(e-150) of bin b (e >= 1)
  return pow2d((int)((klo + (b << kQShift)) >> 23) - 150);
}
__device__ __forceinline__ double q_bin_sum(uint32_t klo, uint32_t b, uint32_t cnt, uint32_t rem) {
  return (double)q_bin_isum(klo, b, cnt, rem) * q_bin_scale(klo, b);
}

template <bool TERN, int VEC, int T, int IDS>
__global__ void __launch_bounds__(T, 1024 / T)
quant_act_kernel(const float* __restrict__ x, ActGeom g, float alpha, Prologue pro, int cs, QParams qp,
                 uint32_t* __restrict__ planes, float* __restrict__ scales, int* __restrict__ row_status,
                 int* __restrict__ diag, uint16_t* __restrict__ gids) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using QSmem = QSmemT<T, IDS>;
  constexpr int kBinsPerThread = kQBins / T;          // consecutive bins per thread in the scan (8 or 16)
  constexpr int kQWarpQueue = kQQueueTotal / (T / 32);
  QSmem& sm = *reinterpret_cast<QSmem*>(smem_raw);
  cg::cluster_group cluster = cg::this_cluster();
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int rank = (cs > 1) ? (int)cluster.block_rank() : 0;
  const long long row = (long long)blockIdx.x / cs;
  QSmem* const r0 = (cs > 1) ? cluster.map_shared_rank(&sm, 0) : &sm;
  const float* xr = x + row * qp.len;
  const uint32_t klo = qp.klo, n = qp.n_s;
  // bin index of every sampled element: in shared memory when the CTA's share of the row fits, else (one CTA per
  // long row) in a global scratch row that is written in sweep 1 and read back, L2 hot, by the collection walk
  uint16_t* const ids = gids ? gids + row * ((4ll * qp.groups + 7ll) & ~7ll) : sm.ids;      // 16-byte aligned rows
  auto csync = [&]() { if (cs > 1) cluster.sync(); else __syncthreads(); };
  // phase cycle counters of rank 0 (diagnostics only: one clock read per phase by one thread)
  long long tprev = 0, tph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const bool timing = diag != nullptr && tid == 0 && rank == 0;
  if (timing) tprev = clock64();
#define LSQ_QTICK(i) do { if (timing) { const long long tn = clock64(); tph[i] += tn - tprev; tprev = tn; } } while (0)

  // this CTA's share of the row, and the loads of its first trip of sweep 1: requested before the setup below so that their
  // latency passes under it (every later trip requests the next one's loads before it processes its own)
  const uint32_t g_lo = (uint32_t)(((unsigned long long)qp.groups * (unsigned)rank) / (unsigned)cs);
  const uint32_t g_hi = (uint32_t)(((unsigned long long)qp.groups * (unsigned)(rank + 1)) / (unsigned)cs);
#ifndef LSQ_QACT_KU
#define LSQ_QACT_KU 2
#endif
  constexpr int kU = LSQ_QACT_KU;   // groups per thread and trip: 4 independent loads each in flight
  auto issue_trip = [&](uint32_t gb, float (&raw)[kU][4]) {
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const uint32_t gi = gb + u * T;
      const long long i0 = (long long)gi * 12;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const long long idx = i0 + 3 * j;
        raw[u][j] = (gi < g_hi && idx < qp.len) ? ldg_stream(xr + idx) : 0.0f;
      }
    }
  };
  float nxt[kU][4];
  issue_trip(g_lo + tid, nxt);

  // ---- setup --------------------------------------------------------------------------------------------------
  for (int b = tid; b < kQBins; b += T) { sm.hcnt[b] = 0u; sm.hrem[b] = 0u; }
  for (int c = tid; c < g.cw * 32; c += T) {
    float2 k = make_float2(1.0f, 0.0f);
    if (pro.a && c < g.c) k = make_float2(__ldg(pro.a + c), __ldg(pro.b + c) + 0.0f);
    sm.ab[c] = k;
  }
  if (tid == 0) {
    sm.kmin = kNoKey; sm.kmax = 0u; sm.cb = 0u; sm.nflag = 0; sm.ngroup = 0; sm.nlist = 0u;
    sm.status = 0; sm.v1 = 0.0f; sm.ncand = 0u;
  }
  __syncthreads();

  // ---- sweep 1: histogram of the sampled keys of this CTA's share of the row -------------------------------------
  {
    uint32_t kmn = kNoKey, kmx = 0u, cb = 0u;
    double lb = 0.0;
    auto sample = [&](float raw, uint32_t c) -> uint32_t {
      const float2 k = sm.ab[c];
      const float a = fabsf(clamp_sym(fmaf(raw, k.x, k.y), alpha));
      const uint32_t key = __float_as_uint(a);
      kmn = min(kmn, key); kmx = max(kmx, key);
      if (key >= klo) {
        const uint32_t b = min((key - klo) >> kQShift, (uint32_t)(kQBins - 1));
        atomicAdd(&sm.hcnt[b], 1u);
        atomicAdd(&sm.hrem[b], key & ((1u << kQShift) - 1u));
        return b;
      }
      ++cb; lb += (double)a;
      return kIdBelow;
    };
    for (uint32_t gb = g_lo + tid; gb < g_hi; gb += kU * T) {
      float raw[kU][4];
#pragma unroll
      for (int u = 0; u < kU; ++u)
#pragma unroll
        for (int j = 0; j < 4; ++j) raw[u][j] = nxt[u][j];
      if (gb + kU * T < g_hi) issue_trip(gb + kU * T, nxt);
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const uint32_t gi = gb + u * T;
        if (gi >= g_hi) break;
        const unsigned long long i0 = (unsigned long long)gi * 12ull;
        uint32_t c = q_channel(qp, i0);
        uint32_t r = (uint32_t)(i0 - (unsigned long long)c * qp.hw);
        uint32_t id[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if ((long long)(i0 + 3 * j) < qp.len) {
            while (r >= qp.hw) { r -= qp.hw; ++c; }
            id[j] = sample(raw[u][j], c);
          } else {
            id[j] = kIdNone;
          }
          r += 3;
        }
        *reinterpret_cast<uint2*>(&ids[4 * (gi - g_lo)]) = make_uint2(id[0] | (id[1] << 16), id[2] | (id[3] << 16));
      }
    }
    const double sb = block_sum(lb, sm.red);
    cb = (uint32_t)__reduce_add_sync(0xffffffffu, cb);
    kmn = warp_min_u32(kmn); kmx = warp_max_u32(kmx);
    if (lane == 0) { atomicAdd(&sm.cb, cb); atomicMin(&sm.kmin, kmn); atomicMax(&sm.kmax, kmx); }
    __syncthreads();
    if (tid == 0) {
      r0->part_sb[rank] = sb; r0->part_cb[rank] = sm.cb; r0->part_kmin[rank] = sm.kmin; r0->part_kmax[rank] = sm.kmax;
    }
  }
  LSQ_QTICK(0);   // setup + sweep 1
  csync();
  LSQ_QTICK(1);   // first cluster barrier (waits for the slowest CTA's sweep)

  // ---- merge: CTA r sums slice r of all histograms into rank 0's ---------------------------------------------------
  if (cs > 1) {
    const int slice = kQBins / cs;
    for (int b = rank * slice + tid; b < (rank + 1) * slice; b += T) {
      uint32_t cq[kQMaxCluster], rq[kQMaxCluster];
#pragma unroll
      for (int q = 0; q < kQMaxCluster; ++q) {       // all remote loads in flight together (one round trip)
        cq[q] = 0u; rq[q] = 0u;
        if (q < cs) {
          const QSmem* p = cluster.map_shared_rank(&sm, q);
          cq[q] = p->hcnt[b]; rq[q] = p->hrem[b];
        }
      }
      uint32_t c = 0u, r = 0u;
#pragma unroll
      for (int q = 0; q < kQMaxCluster; ++q) { c += cq[q]; r += rq[q]; }
      r0->hcnt[b] = c; r0->hrem[b] = r;
    }
    cluster.sync();
  }
  LSQ_QTICK(2);   // merge

  // ---- rank 0: scan the bins, flag those that can hold a candidate, lay out one list segment per flagged bin -------
  if (rank == 0) {
    double sum_below = 0.0;
    uint32_t cnt_below = 0u, kmin = kNoKey, kmax = 0u;
    for (int q = 0; q < cs; ++q) {
      sum_below += sm.part_sb[q]; cnt_below += sm.part_cb[q];
      kmin = min(kmin, sm.part_kmin[q]); kmax = max(kmax, sm.part_kmax[q]);
    }
    // a thread's 8 consecutive bins lie in one octave (512 bins, aligned): their sums add up as integers
    uint32_t ct = 0u, fn = kNoKey;
    unsigned long long it = 0ull;
#pragma unroll
    for (int j = 0; j < kBinsPerThread; ++j) {
      const uint32_t b = tid * kBinsPerThread + j, cnt = sm.hcnt[b];
      if (cnt != 0u) {
        ct += cnt; it += q_bin_isum(klo, b, cnt, sm.hrem[b]);
        if (fn == kNoKey) fn = b;
      }
    }
    const double stt = (double)it * q_bin_scale(klo, tid * kBinsPerThread);
    uint32_t ci = ct;
    double si = stt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t tc = __shfl_up_sync(0xffffffffu, ci, o);
      const double ts = __shfl_up_sync(0xffffffffu, si, o);
      if (lane >= o) { ci += tc; si += ts; }
    }
    uint32_t sfx = fn;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_down_sync(0xffffffffu, sfx, o);
      if (lane + o < 32) sfx = min(sfx, t);
    }
    uint32_t nxt_in_warp = __shfl_down_sync(0xffffffffu, sfx, 1);
    if (lane == 31) nxt_in_warp = kNoKey;
    if (lane == 31) { sm.wcnt[wid] = ci; sm.wsum[wid] = si; }
    if (lane == 0) sm.wfirst[wid] = sfx;
    __syncthreads();
    uint32_t coff = 0u;
    double soff = 0.0, s_bins = 0.0;
    for (int w = 0; w < T / 32; ++w) {
      if (w < wid) { coff += sm.wcnt[w]; soff += sm.wsum[w]; }
      s_bins += sm.wsum[w];
    }
    uint32_t nxt_after = nxt_in_warp;
    for (int w = wid + 1; w < T / 32 && nxt_after == kNoKey; ++w) nxt_after = sm.wfirst[w];
    const double s_tot = sum_below + s_bins;
    const uint32_t excl0 = cnt_below + coff + ci - ct;
    const double pref0 = sum_below + soff + si - stt;

    auto bin_upper = [&](uint32_t b) -> float {       // largest value a key of bin b can have
      const uint32_t nh = min(klo + ((b + 1u) << kQShift) - 1u, kmax);
      return fmaxf(key_val(nh), key_val(klo + (b << kQShift)));
    };
    auto test_bin = [&](uint32_t b, uint32_t cnt, double s, uint32_t excl, double pref, uint32_t nb) {
      const uint32_t elo_k = klo + (b << kQShift);
      const uint32_t ehi_k = min(klo + ((b + 1u) << kQShift) - 1u, kmax);
      const float nxt_hi = (nb != kNoKey) ? bin_upper(nb) : key_val(kmax);
      if (!may_hold<TERN>(elo_k, ehi_k, cnt, s, excl, pref, nxt_hi, n, s_tot, kmax, 0.0f)) return;
      const int slot = atomicAdd(&sm.nflag, 1);
      if (slot < kQMaxFlag) { sm.fbin[slot] = (uint16_t)b; sm.fexcl[slot] = excl; sm.fsumb[slot] = pref; sm.fnb[slot] = nb; }
    };
    // two levels: one coarse test per thread on the union of its bins, the fine test only for flagged groups
    uint32_t* const gexcl = sm.list;
    uint32_t* const gnext = sm.list + T;
    double* const gpref = reinterpret_cast<double*>(sm.list + 2 * T);
    gexcl[tid] = excl0; gnext[tid] = nxt_after; gpref[tid] = pref0;
    if (ct != 0u) {
      const uint32_t b0 = tid * kBinsPerThread;
      const uint32_t elo_k = klo + (b0 << kQShift);
      const uint32_t ehi_k = min(klo + ((b0 + kBinsPerThread) << kQShift) - 1u, kmax);
      const float nxt_hi = (nxt_after != kNoKey) ? bin_upper(nxt_after) : key_val(kmax);
      if (may_hold<TERN>(elo_k, ehi_k, ct, stt, excl0, pref0, nxt_hi, n, s_tot, kmax, 0.0f)) {
        const int slot = atomicAdd(&sm.ngroup, 1);
        if (slot < kQMaxGroups) sm.glist[slot] = (uint16_t)tid;
      }
    }
    __syncthreads();
    const int ngroup = sm.ngroup;
    const uint32_t nfine = (uint32_t)min(ngroup, kQMaxGroups) * (uint32_t)kBinsPerThread;
    for (uint32_t idx = tid; idx < nfine; idx += T) {
      const uint32_t gid = sm.glist[idx / kBinsPerThread], j = idx % kBinsPerThread;
      const uint32_t b = gid * kBinsPerThread + j, cnt = sm.hcnt[b];
      if (cnt == 0u) continue;
      uint32_t excl = gexcl[gid];
      unsigned long long ib = 0ull;
      for (uint32_t j1 = 0; j1 < j; ++j1) {
        const uint32_t b1 = gid * kBinsPerThread + j1, c1 = sm.hcnt[b1];
        if (c1 != 0u) { excl += c1; ib += q_bin_isum(klo, b1, c1, sm.hrem[b1]); }
      }
      const double pref = gpref[gid] + (double)ib * q_bin_scale(klo, b);
      uint32_t nb = kNoKey;
      for (uint32_t j2 = kBinsPerThread - 1; j2 > j; --j2)
        if (sm.hcnt[gid * kBinsPerThread + j2] != 0u) nb = gid * kBinsPerThread + j2;
      if (nb == kNoKey) nb = gnext[gid];
      test_bin(b, cnt, q_bin_sum(klo, b, cnt, sm.hrem[b]), excl, pref, nb);
    }
    __syncthreads();
    const int nflag = sm.nflag;
    // warps 0-1: ascending order of the flagged bins and one power-of-two padded list segment for each (two lanes
    // steps of a 64-entry scan); warp 2: the pseudo bin below the window; in parallel
    if (tid == 0) { sm.s_tot = s_tot; sm.kmin = kmin; sm.kmax = kmax; sm.nlist = 0u; }
    int bad = 0;
    if (n < 3u) bad = 1;
    if (ngroup > kQMaxGroups || nflag > kQMaxFlag) bad = 2;
    if (bad == 0 && wid == 0) {
      uint32_t run = 0u;
      for (int base = 0; base < nflag; base += 32) {
        const int t = base + lane;
        uint32_t lp = 0u, cnt = 0u;
        int rk = 0;
        if (t < nflag) {
          const uint32_t mine = sm.fbin[t];
          for (int i = 0; i < nflag; ++i) rk += (sm.fbin[i] < mine) ? 1 : 0;   // bins are distinct
          cnt = sm.hcnt[mine];
          lp = 2u;
          while (lp < cnt) lp <<= 1;
          if (cnt > (uint32_t)kQSegMax) bad = 4;
          sm.sbin[rk] = (uint16_t)mine;
          sm.snb[rk] = (sm.fnb[t] != kNoKey) ? (uint16_t)sm.fnb[t] : (uint16_t)0xFFFFu;
          sm.sexcl[rk] = sm.fexcl[t]; sm.spref[rk] = sm.fsumb[t];
          sm.seg_cnt[rk] = cnt;
          sm.ford[rk] = (int)lp;         // padded size by rank, turned into offsets below
        }
      }
      __syncwarp();
      for (int base = 0; base < nflag; base += 32) {
        const int t = base + lane;
        const uint32_t lp = (t < nflag) ? (uint32_t)sm.ford[t] : 0u;
        uint32_t inc = lp;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
          if (lane >= o) inc += v;
        }
        if (t < nflag) sm.seg_start[t] = run + inc - lp;
        run += __shfl_sync(0xffffffffu, inc, 31);
      }
      bad = __reduce_max_sync(0xffffffffu, bad);
      if (bad == 0 && run > (uint32_t)kQListCap) bad = 4;
      if (lane == 0) { sm.nlist = run; if (bad) atomicMax(&sm.status, bad); }
    } else if (bad == 0 && wid == 2 && lane == 0 && cnt_below != 0u) {
      // everything below the bin window is one pseudo bin: if it could hold a candidate the row needs finer
      // treatment there (never seen on clamped BatchNorm outputs)
      uint32_t fb = kNoKey;
      for (int w = 0; w < T / 32 && fb == kNoKey; ++w) fb = sm.wfirst[w];
      const float nxt_hi = (fb != kNoKey) ? bin_upper(fb) : key_val(kmax);
      if (may_hold<TERN>(0u, klo - 1u, cnt_below, sum_below, 0u, 0.0, nxt_hi, n, s_tot, kmax, 0.0f)) atomicMax(&sm.status, 3);
    }
    if (bad != 0 && tid == 0) atomicMax(&sm.status, bad);
    __syncthreads();
  }
  LSQ_QTICK(3);   // scan + flag + segments
  csync();

  // ---- collect: exact keys of the sampled elements of the flagged bins -> this CTA's copy of the segments ------------
  int status = r0->status;
  const int nflag = (status == 0) ? r0->nflag : 0;
  if (nflag > 0) {
    // every CTA: bin -> segment map (the histogram's remainder sums are dead by now)
    for (int b = tid; b < kQBins / 4; b += T) reinterpret_cast<uint32_t*>(sm.fmap)[b] = 0u;
    if (rank != 0 && tid < nflag) { sm.sbin[tid] = r0->sbin[tid]; sm.snb[tid] = r0->snb[tid]; sm.seg_start[tid] = r0->seg_start[tid]; }
    if (tid < kQMaxFlag) { sm.lfill[tid] = 0u; sm.lsmin[tid] = kNoKey; }
    __syncthreads();
    if (tid < nflag && sm.snb[tid] != 0xFFFFu) sm.fmap[sm.snb[tid]] = (uint8_t)(65 + tid);
    __syncthreads();
    if (tid < nflag) sm.fmap[sm.sbin[tid]] = (uint8_t)(1 + tid);     // a flagged bin that follows another one is a segment
    __syncthreads();
    // phase A: walk the bin indices; a warp queues its hits (sample, segment) without touching global memory ...
    const uint32_t ns_cta = 4u * (g_hi - g_lo);
    uint32_t* const wq = sm.hcnt + wid * kQWarpQueue;     // (sample index << 8 | segment code) of this warp's hits
    uint32_t nq = 0u;                                  // warp uniform
    // a lane takes 8 consecutive samples per 16-byte load (global scratch rows: L2 latency), two loads in flight
    constexpr int kW = 2;
    for (uint32_t base0 = 256u * wid; base0 < ns_cta; base0 += kW * 8u * T) {
      uint4 wv[kW];
#pragma unroll
      for (int u = 0; u < kW; ++u) {
        const uint32_t s8 = base0 + u * 8u * T + 8u * lane;
        wv[u] = make_uint4(0xFFFEFFFEu, 0xFFFEFFFEu, 0xFFFEFFFEu, 0xFFFEFFFEu);
        if (s8 + 8u <= ns_cta) wv[u] = *reinterpret_cast<const uint4*>(&ids[s8]);
        else if (s8 < ns_cta) { const uint2 h = *reinterpret_cast<const uint2*>(&ids[s8]); wv[u].x = h.x; wv[u].y = h.y; }
      }
#pragma unroll
      for (int u = 0; u < kW; ++u) {
        const uint32_t s8 = base0 + u * 8u * T + 8u * lane;
        if (base0 + u * 8u * T >= ns_cta) break;       // warp uniform
        const uint32_t w4[4] = {wv[u].x, wv[u].y, wv[u].z, wv[u].w};
        uint32_t m8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {                  // eight independent lookups, then the ballots
          const uint32_t id = (j & 1) ? (w4[j >> 1] >> 16) : (w4[j >> 1] & 0xFFFFu);
          m8[j] = (id < (uint32_t)kQBins) ? sm.fmap[id] : 0u;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t hit = __ballot_sync(0xffffffffu, m8[j] != 0u);
          if (hit != 0u) {
            const uint32_t pos = nq + __popc(hit & ((1u << lane) - 1u));
            if (m8[j] != 0u && pos < (uint32_t)kQWarpQueue) wq[pos] = ((s8 + j) << 8) | m8[j];
            nq += __popc(hit);
          }
        }
      }
    }
    if (nq > (uint32_t)kQWarpQueue) { if (lane == 0) atomicMax(&r0->status, 6); nq = kQWarpQueue; }
    __syncwarp();
    if (timing && cs == 1) { const long long tn = clock64(); tph[7] = tn - tprev; }      // map build + index walk
    // ... phase B: the exact keys of all queued hits, their loads in flight together (one trip to L2)
    for (uint32_t e = lane; e < nq; e += 32u) {
      const uint32_t ent = wq[e], m = ent & 0xFFu;
      const unsigned long long idx = 3ull * (4ull * g_lo + (ent >> 8));
      const uint32_t c = q_channel(qp, idx);
      const float2 k = sm.ab[c];
      const uint32_t key = __float_as_uint(fabsf(clamp_sym(fmaf(__ldg(xr + idx), k.x, k.y), alpha)));
      if (m <= 64u) {
        const uint32_t pos = atomicAdd(&sm.lfill[m - 1u], 1u);
        if (pos < (uint32_t)kQSegMax) sm.list[sm.seg_start[m - 1u] + pos] = key;
      } else {
        atomicMin(&sm.lsmin[m - 65u], key);
      }
    }
  }
  csync();
  LSQ_QTICK(4);   // collect (incl. the barriers around it)

  // ---- rank 0: gather the segments, sort each in one warp, the reference's candidate test, closed-form cost -> v1 ---
  if (rank == 0) {
    Best best{1e300, 0xFFFFFFFFu, 0u};
    uint32_t ncand = 0u;
    const double s_tot = sm.s_tot;
    if (status == 0) {
      for (int i = wid; i < nflag; i += T / 32) {
        const uint32_t base = sm.seg_start[i], cnt = sm.seg_cnt[i];
        uint32_t fill = sm.lfill[i], smin = sm.lsmin[i];
        if (cs > 1) {
          // lane q (1 <= q < cs) fetches CTA q's count and follower minimum, then the warp copies all remote
          // elements with independent loads: two round trips through distributed shared memory per segment
          uint32_t cq = 0u, sq = kNoKey;
          if (lane >= 1 && lane < cs) {
            const QSmem* p = cluster.map_shared_rank(&sm, lane);
            cq = p->lfill[i]; sq = p->lsmin[i];
          }
          uint32_t inc = cq;
#pragma unroll
          for (int o = 1; o < kQMaxCluster; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += v;
          }
          const uint32_t total_remote = __shfl_sync(0xffffffffu, inc, kQMaxCluster - 1);
          smin = min(smin, warp_min_u32(sq));
          uint32_t endt[kQMaxCluster];                 // inclusive prefix of CTA t, known to every lane
#pragma unroll
          for (int t = 0; t < kQMaxCluster; ++t) endt[t] = __shfl_sync(0xffffffffu, inc, t);
          for (uint32_t e = lane; e < total_remote; e += 32u) {
            int q = 1;
            uint32_t off = 0u;
#pragma unroll
            for (int t = 1; t < kQMaxCluster; ++t)
              if (e >= endt[t]) { q = t + 1; off = endt[t]; }
            if (q < cs && fill + e < (uint32_t)kQSegMax) {
              const QSmem* p = cluster.map_shared_rank(&sm, q);
              sm.list[base + fill + e] = p->list[base + (e - off)];
            }
          }
          fill += total_remote;
        }
        uint32_t lp = 2;
        while (lp < cnt) lp <<= 1;
        for (uint32_t e = cnt + lane; e < lp; e += 32) sm.list[base + e] = kNoKey;
        __syncwarp();                              // every lane has read lfill / lsmin of the segment (racecheck: WAR)
        if (lane == 0) {
          sm.lsmin[i] = smin;
          if (fill != cnt) sm.status = 5;        // cannot happen: the histogram and the bin indices disagree
        }
        __syncwarp();
        uint32_t* keys = sm.list + base;
        for (uint32_t k = 2; k <= lp; k <<= 1)
          for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t t = lane; t < (lp >> 1); t += 32) {
              const uint32_t a0 = ((t & ~(j - 1)) << 1) | (t & (j - 1));
              const uint32_t p1 = a0 | j;
              const uint32_t ka = keys[a0], kb = keys[p1];
              const bool up = ((a0 & k) == 0);
              if ((ka > kb) == up) { keys[a0] = kb; keys[p1] = ka; }
            }
            __syncwarp();
          }
      }
    }
    __syncthreads();
    status = sm.status;
    if (status == 0) {
      for (int i = wid; i < nflag; i += T / 32) {
        const uint32_t base = sm.seg_start[i], cnt = sm.seg_cnt[i];
        const uint32_t* keys = sm.list + base;
        // successor of the segment's last element: the smallest key of the next non-empty bin
        uint32_t succ = sm.kmax;
        const uint32_t nb = sm.snb[i];
        if (nb != 0xFFFFu) {
          const uint32_t m = sm.fmap[nb];
          succ = (m >= 1u && m <= 64u) ? sm.list[sm.seg_start[m - 1u]] : sm.lsmin[i];
        }
        const uint32_t per = (cnt + 31u) / 32u;
        const uint32_t j0 = min((uint32_t)lane * per, cnt), j1 = min(j0 + per, cnt);
        double loc = 0.0;
        for (uint32_t j = j0; j < j1; ++j) loc += (double)key_val(keys[j]);
        double inc = loc;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const double t = __shfl_up_sync(0xffffffffu, inc, o);
          if (lane >= o) inc += t;
        }
        double run = sm.spref[i] + inc - loc;
        const uint32_t excl = sm.sexcl[i];
        for (uint32_t j = j0; j < j1; ++j) {
          const uint32_t kj = keys[j];
          run += (double)key_val(kj);
          try_position<TERN>(best, ncand, kj, (j + 1u < cnt) ? keys[j + 1u] : succ, excl + j, run, n, s_tot, 0.0);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double oc = __shfl_xor_sync(0xffffffffu, best.cost, o);
        const uint32_t op = __shfl_xor_sync(0xffffffffu, best.pos, o);
        const uint32_t ok = __shfl_xor_sync(0xffffffffu, best.key, o);
        best.offer(oc, op, ok);
      }
      ncand = (uint32_t)__reduce_add_sync(0xffffffffu, ncand);
      if (lane == 0) {
        sm.best_cost[wid] = best.cost; sm.best_pos[wid] = best.pos; sm.best_key[wid] = best.key;
        atomicAdd(&sm.ncand, ncand);
      }
      __syncthreads();
      if (tid == 0) {
        Best b{1e300, 0xFFFFFFFFu, 0u};
        for (int w = 0; w < T / 32; ++w) b.offer(sm.best_cost[w], sm.best_pos[w], sm.best_key[w]);
        uint32_t nc_tot = sm.ncand;
        if (TERN) {
          // optimal.py:86-118: when min > mean/2 the value mean/2 (not a data element) is appended last
          const float mean = (float)(s_tot / (double)n);
          const float half_mean = __fmul_rn(0.5f, mean);
          if (key_val(sm.kmin) > half_mean) {
            ++nc_tot;
            b.offer(closed_cost2<true>((double)half_mean, 0.0, 0.0, (double)n, s_tot, 0.0), n, __float_as_uint(half_mean));
          }
        }
        const float v1 = (nc_tot > 0u) ? key_val(b.key) : 0.0f;
        sm.v1 = v1;
        scales[row] = v1;
        if (TERN) scales[(long long)g.n + row] = v1;
        sm.ncand = nc_tot;
      }
    }
    if (tid == 0) {
      row_status[row] = sm.status;
      if (diag) {
        int* d = diag + row * 16;
        d[0] = sm.status; d[1] = sm.nflag; d[2] = (int)sm.nlist; d[3] = (int)sm.ncand; d[4] = 0; d[5] = cs;
        d[6] = (int)sm.part_cb[0]; d[7] = sm.ngroup;
      }
    }
    __syncthreads();
  }
  LSQ_QTICK(5);   // sort + evaluate
  csync();
  status = r0->status;
  const float v1 = r0->v1;
  csync();                          // rank 0 must not leave (and release its shared memory) before everyone has read
  if (status != 0) return;          // the generic kernels redo this row (uniform over the cluster)

  // ---- sweep 2 (L2): both sign planes and sum |x - v1 sign(x)| ----------------------------------------------------
  {
    const uint32_t it_lo = (uint32_t)(((unsigned long long)qp.nitems * (unsigned)rank) / (unsigned)cs);
    const uint32_t it_hi = (uint32_t)(((unsigned long long)qp.nitems * (unsigned)(rank + 1)) / (unsigned)cs);
    const double acc_sum = encode2_row_part<VEC>(xr, g, (int)row, qp.hw, qp.nq, it_lo, it_hi, sm.ab, alpha, v1, planes);
    if (!TERN) {
      const double tot = block_sum(acc_sum, sm.red);
      if (tid == 0) r0->part_v2[rank] = tot;
      csync();
      if (rank == 0 && tid == 0) {
        double t = 0.0;
        for (int q = 0; q < cs; ++q) t += sm.part_v2[q];
        scales[(long long)g.n + row] = (float)(t / (double)qp.len);
      }
    }
  }
  LSQ_QTICK(6);   // sweep 2 (+ barriers)
  if (timing) {
    int* d = diag + row * 16;
    for (int i = 0; i < 7; ++i) d[8 + i] = (int)tph[i];
    if (cs == 1) d[10] = (int)tph[7];       // single-CTA rows have no merge: the slot reports the index walk instead
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    d[15] = (int)smid;
  }
#undef LSQ_QTICK
}

}  // namespace lsq

using namespace lsq;

// implemented in lsq_solve.cu / lsq_quant.cu: the generic kernels restricted to the rows marked in d_row_status
namespace lsq {
int solve_v1_marked_rows(const float* d_x, int64_t rows, int64_t len, int skip, int ternary, float alpha, float* d_v1,
                         float* d_v1_dup, const lsq_prologue* pro, const int* d_row_status, cudaStream_t stream);
int encode_act_marked_rows(const float* d_x, const lsq_act_geom* g, float alpha, const float* d_scales, int nscales,
                           int nplanes, uint32_t* d_planes, float* d_last_scale, void* d_ws, size_t ws_bytes,
                           const lsq_prologue* pro, const int* d_row_status, cudaStream_t stream);
// one small launch: generic solve + encode of the rows marked in d_row_status (lsq_solve.cu)
int qact_fallback_launch(const float* d_x, const lsq_act_geom* g, float alpha, int ternary, int skip, uint32_t* d_planes,
                         float* d_scales, const lsq_prologue* pro, const int* d_row_status, bool vec4, cudaStream_t stream);
}

// Launch shape for rows of `len` elements.  Rows whose sampled third fits the bin-index buffer of a CTA (17 408
// samples: every layer input of the CIFAR network, stages 3-4 of the ImageNet one) run one CTA per row with the indices
// in shared memory.  Longer rows (64 x 56 x 56 and 128 x 28 x 28):
//   default   still one CTA per row, two per SM, the bin indices in a global scratch row (2 bytes per sample, written
//             in sweep 1, read back L2-hot by the collection walk): 296 rows in flight, every SM streams all the time,
//             no cross-CTA traffic -- but the rows in flight (238 / 119 MB) exceed L2, so the second sweep comes from
//             HBM again: 2 x the input in DRAM traffic;
//   cluster   LSQ_QACT_MODE=cluster: a cluster of 2-8 CTAs per row, sized so that the rows in flight stay L2 resident:
//             DRAM traffic 1.07 x the input, but three of four CTAs idle while rank 0 solves and seven cluster
//             barriers per row (measured 1.3 x slower than the default: DESIGN.md section 5).
struct QactShape { int cs; bool global_ids; bool t256; };
static QactShape qact_shape(int64_t len, int sms) {
  static const int forced = getenv("LSQ_QACT_CS") ? atoi(getenv("LSQ_QACT_CS")) : 0;   // development override
  static const bool want_cluster = getenv("LSQ_QACT_MODE") && getenv("LSQ_QACT_MODE")[0] == 'c';
  static const int tmode = getenv("LSQ_QACT_T") ? atoi(getenv("LSQ_QACT_T")) : 0;      // development: 256-thread CTAs, 4 per SM
  const int64_t groups = (len + 11) / 12;
  if (!want_cluster && forced == 0 && tmode == 256) return {1, true, true};
  if (4 * (groups + 1) <= (int64_t)kQIdsCap && forced <= 1) return {1, false, false};
  if (!want_cluster && forced == 0) return {1, true, false};
  int cs = 1;
  while (cs < kQMaxCluster && 4 * ((groups + cs - 1) / cs + 1) > (int64_t)kQIdsCap) cs <<= 1;
  const double l2_budget = 64.0 * 1024 * 1024;
  while (cs < kQMaxCluster && (double)len * 4.0 * (2.0 * sms / cs) > l2_budget) cs <<= 1;
  if (forced == 1 || forced == 2 || forced == 4 || forced == 8) {
    if (4 * ((groups + forced - 1) / forced + 1) <= (int64_t)kQIdsCap) cs = forced;
  }
  return {cs, false, false};
}

static bool qact_fast_path(const lsq_act_geom* g, float alpha, int skip) {
  static const bool off = getenv("LSQ_QACT_OFF") != nullptr;      // development: always take the generic kernels
  if (off) return false;
  const int64_t len = (int64_t)g->c * g->h * g->w;
  if (skip != 3 || !(alpha >= 1e-30f) || !(alpha < 3e38f)) return false;
  if (len < 2048 || (len + 2) / 3 > 131071) return false;       // short rows: the sorting kernel; 32-bit bin sums
  if (g->cw * 32 > kQMaxChannels) return false;
  return true;
}

extern "C" size_t lsq_quantize_act_workspace_bytes(const lsq_act_geom* g) {
  if (!g) return 0;
  const int64_t len = (int64_t)g->c * g->h * g->w;
  const size_t ids = (size_t)g->n * (size_t)((((len + 11) / 12) * 4 + 7) & ~7ll) * sizeof(uint16_t);   // bin indices of long rows
  return (size_t)g->n * sizeof(int) + 512 + lsq_reduce_workspace_bytes(g->n, len) + ids;
}

extern "C" int lsq_quantize_act(const float* d_x, const lsq_act_geom* g, float alpha, int ternary, int skip,
                                uint32_t* d_planes, float* d_scales, void* d_ws, size_t ws_bytes,
                                const lsq_prologue* pro, int32_t* d_diag, void* stream) {
  LSQ_CHECK_ARG(d_x && g && d_planes && d_scales && d_ws, "lsq_quantize_act: null pointer");
  LSQ_CHECK_ARG(skip >= 1, "lsq_quantize_act: bad skip %d", skip);
  if (ws_bytes < lsq_quantize_act_workspace_bytes(g)) {
    set_error("lsq_quantize_act: workspace %zu < %zu", ws_bytes, lsq_quantize_act_workspace_bytes(g));
    return LSQ_ERR_WORKSPACE;
  }
  const int64_t len = (int64_t)g->c * g->h * g->w;
  if (pro && pro->d_ch_scale && (pro->channels != g->c || pro->inner != (int64_t)g->h * g->w)) {
    set_error("lsq_quantize_act: prologue must have %d channels of %d elements", g->c, g->h * g->w);
    return LSQ_ERR_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  // workspace: [reduce workspace of the encoder (arrival counters first, zero at rest)][row status, one int per row]
  const size_t red_bytes = (lsq_reduce_workspace_bytes(g->n, len) + 255) / 256 * 256;
  int* d_status = (int*)((char*)d_ws + red_bytes);
  uint16_t* d_ids = (uint16_t*)((char*)d_ws + red_bytes + ((size_t)g->n * sizeof(int) + 255) / 256 * 256);
  float* d_v2 = ternary ? nullptr : d_scales + g->n;
  if (!qact_fast_path(g, alpha, skip)) {
    // not the shape the fused kernel is built for: the generic kernels on every row (d_row_status = NULL)
    int rc = solve_v1_marked_rows(d_x, g->n, len, skip, ternary, alpha, d_scales, ternary ? d_scales + g->n : nullptr,
                                  pro, nullptr, st);
    if (rc != LSQ_OK) return rc;
    return encode_act_marked_rows(d_x, g, alpha, d_scales, 1, 2, d_planes, d_v2, d_ws, red_bytes, pro, nullptr, st);
  }
  const int sms = device_sms();
  const QactShape shape = qact_shape(len, sms);
  const int cs = shape.cs;
  QParams qp;
  qp.len = len; qp.n_s = (uint32_t)((len + 2) / 3); qp.groups = (uint32_t)((len + 11) / 12);
  const uint32_t ka = __builtin_bit_cast(uint32_t, alpha) >> kQShift;
  // window of kQBins bins whose last bin holds key(alpha); its first bin is rounded up to a multiple of 16 absolute
  // bins so that the 8 or 16 consecutive bins a thread scans never straddle an octave (exponent) boundary
  qp.klo = (((ka + 1u - (uint32_t)kQBins) + 15u) & ~15u) << kQShift;
  qp.hw = (uint32_t)(g->h * g->w);
  const bool vec4 = (qp.hw % 4 == 0) && (reinterpret_cast<uintptr_t>(d_x) % 16 == 0);
  qp.nq = vec4 ? qp.hw / 4 : qp.hw;
  qp.nitems = qp.nq * (uint32_t)g->cw;
  qp.hw_magic = ((double)len * (double)qp.hw <= 1099511627776.0) ? ((1ull << 40) + qp.hw - 1ull) / qp.hw : 0ull;
  const ActGeom dg = to_dev(*g);
  const Prologue dp = to_dev(pro);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)((int64_t)g->n * cs));
  cfg.blockDim = dim3(shape.t256 ? 256 : kQThreads);
  cfg.dynamicSmemBytes = shape.t256 ? sizeof(QSmemT<256, 64>) : sizeof(QSmemT<kQThreads, kQIdsCap>);
  uint16_t* gids = shape.global_ids ? d_ids : nullptr;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = cs > 1 ? 1 : 0;
  cudaError_t e = cudaSuccess;
#define LSQ_QACT(TERNARY, V, THR, IDS)                                                                             \
  do {                                                                                                             \
    static std::atomic<unsigned long long> smem_set{0ull};                                                         \
    e = ensure_max_smem(quant_act_kernel<TERNARY, V, THR, IDS>, smem_set);                                         \
    if (e == cudaSuccess)                                                                                          \
      e = cudaLaunchKernelEx(&cfg, quant_act_kernel<TERNARY, V, THR, IDS>, d_x, dg, alpha, dp, cs, qp, d_planes,   \
                             d_scales, d_status, (int*)d_diag, gids);                                              \
  } while (0)
#define LSQ_QACT_V(TERNARY, V)                                                                                     \
  do {                                                                                                             \
    if (shape.t256) LSQ_QACT(TERNARY, V, 256, 64); else LSQ_QACT(TERNARY, V, kQThreads, kQIdsCap);                 \
  } while (0)
  if (ternary) { if (vec4) LSQ_QACT_V(true, 4); else LSQ_QACT_V(true, 1); }
  else { if (vec4) LSQ_QACT_V(false, 4); else LSQ_QACT_V(false, 1); }
#undef LSQ_QACT_V
#undef LSQ_QACT
  if (e != cudaSuccess) {
    set_error("lsq_quantize_act: launch (cluster %d): %s", cs, cudaGetErrorString(e));
    return LSQ_ERR_CUDA;
  }
  LSQ_CUDA_LAUNCH_CHECK("quant_act_kernel");
  // rows the fused kernel could not decide (normally none): one small launch that looks at the marks and redoes the
  // marked rows with the generic solver + encoder
  return qact_fallback_launch(d_x, g, alpha, ternary, skip, d_planes, d_scales, pro, d_status, vec4, st);
}
