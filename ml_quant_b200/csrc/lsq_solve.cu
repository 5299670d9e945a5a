// Least-squares optimal v1 for the 2-bit / ternary quantizer: one CTA per row, no global sort.
//
// Replaces opt_v1 -> compute_mask -> cost_function (quant/binary/optimal.py:16-155), which sorts
// every row, forms fp32 prefix means, picks the positions i with  a_i <= thr(i) <= a_{i+1}  for
// thr = hi_i/2 (and (lo_i+hi_i)/2 for 2 bits) and evaluates an fp32 cost on a [rows, cand, n] tensor.
//
// Here (tests/solver_model.py is the executable specification of this file):
//   1. one pass histograms the float bit patterns of a = |clamp(x)|[::skip] (monotone keys) into
//      8192 bins (count + integer sum of the top 14 mantissa bits: native shared-memory atomics, order
//      independent) and accumulates S = sum a, Q = sum a^2 in fp64;
//   2. prefix counts / sums at bin edges bound both threshold functions over each bin, which flags
//      the few bins that can hold a candidate (a conservative fp32 test, 1e-4 margin on the sums);
//   3. a second pass collects only the flagged elements (<= 8192, list A) into shared memory together
//      with the exact fp64 sum of everything below each flagged range; each range is then refined once
//      more IN SHARED MEMORY with 8192 finer bins whose sums are exact integers (mantissa sums per
//      exponent), which leaves a few dozen elements (list B); those are sorted (bitonic) and every one
//      gets the reference's own fp32 candidate test on (float) prefix sums, so the candidate set is the
//      reference's;
//   4. the cost of a candidate is the closed form of optimal.py:31-38 in fp64; the
//      first minimum in ascending order wins, as torch.argmin does.
//   Rows with n <= 8192 skip 1-2 (the whole row is list A).  Ranges too large to collect are
//   refined in child windows with a finer bin shift; at shift 0 a bin is a run of equal values and
//   is evaluated directly.  No host synchronisation (the reference does .tolist(), optimal.py:147).
#include "lsq_common.cuh"
#include "lsq_solve_core.cuh"
#include "lsq_encode_core.cuh"

namespace lsq {

constexpr int kSolveThreads = 512;    // two CTAs per SM: one row's serial phases overlap the other's passes
constexpr int kListBins = 1024;   // bins of a shared-memory refinement window
constexpr int kListBinsLog2 = 10;
constexpr int kFineCap = 2048;  // list B: elements that are sorted and evaluated one by one
constexpr int kMaxFlag = 64;
constexpr int kStack = 24;
constexpr int kLoadBatch = 8;
constexpr int kMaxGroups = 128;        // flagged 16-bin groups whose bins get the fine test in parallel
constexpr int kMaxProChannels = 512;   // prologue tables up to this many channels are staged in shared memory
constexpr int kFineShift = 14, kFineListShift = 16;   // bin width of a top window anchored at the clamp bound
constexpr uint32_t kMaxExactN = 262144;  // sum(m >> 9) of a bin fits 32 bits up to this many elements
struct Window {       // a Span to histogram with bins of 2^shift keys starting at klo
  Span base;          // elements are taken from base (global row: the whole key space)
  uint32_t klo;
  int shift;
  int from_list;      // 0: elements come from the global row, 1: from list A restricted to base
};

// Shared memory.  Two layouts (template parameter of the kernel):
//   L = SolveLayout<13, 8192, 0, 512 threads, 2 CTAs/SM>  115 KB: rows of up to 21504 sampled elements stay entirely in
//                              shared memory (one pass over HBM), 8192 bins, list A of 8192 elements;
//   L = SolveLayout<12, 2048, 0, 256 threads, 4 CTAs/SM>  55 KB: long rows (two streaming passes).  512 rows over
//                              592 slots run as ONE wave (296 slots of 512 threads needed two, the second 73 %
//                              full), and four rows per SM overlap their serial phases; 4096 bins still cover 16
//                              octaves below the clamp bound at 256 bins per octave.  (A cp.async staging ring in
//                              the spare memory of a 2-CTA variant brought 3 %: the passes are instruction bound.)
// Lifetimes overlap as little as possible so regions are reused:
//   list_a       : per-thread bin-group prefixes while a global window is scanned
//   bins.l.ext   : tail of list A when the whole sampled row lives in shared memory
template <int LOG2_BINS, int CAP, int NSTAGE, int THREADS, int MINBLOCKS, int PROCH>
struct SolveLayout {
  static constexpr int kStages = NSTAGE;
  static constexpr int kThreads = THREADS, kMinBlocks = MINBLOCKS, kProChannels = PROCH;
  static constexpr int kBinsLog2 = LOG2_BINS;
  static constexpr int kBins = 1 << LOG2_BINS;
  static constexpr int kBinsPerThread = kBins / THREADS;
  static constexpr int kCap = CAP;                               // list A: elements of the flagged ranges of a global window
  static constexpr int kSmallCap = CAP + 2 * kBins - 3 * kListBins;   // list A when it holds a whole sampled row
  static constexpr int kTopShift = 31 - LOG2_BINS;
  union Bins {
    struct { uint32_t hist[kBins]; uint32_t bsum[kBins]; } g;     // global windows: counts, sum(m >> 9)
    struct {
      uint32_t ext[2 * kBins - 3 * kListBins];
      uint32_t hist[kListBins], lo[kListBins], hi[kListBins];     // list windows: counts, sum(m & 0xFFF), sum(m >> 12)
    } l;
  };
  struct Smem {
    uint32_t list_b[kFineCap];
    uint32_t list_a[kCap];
    Bins bins;
    float2 ab[PROCH];                      // per-channel (scale, shift) of the fused prologue
    double red[32];
    double wsum[32];
    uint32_t wcnt[32];
    uint32_t wfirst[32];
    uint32_t wnz[32];
    int nflag, ngroup;
    uint16_t glist[kMaxGroups];
    uint32_t fmin, fmax;
    uint16_t fbin[kMaxFlag];
    uint32_t fexcl[kMaxFlag];
    uint32_t fnext[kMaxFlag];
    double fsumb[kMaxFlag];
    int ford[kMaxFlag];
    int nrange;
    Range rng[kMaxRanges];
    double seg_base[kMaxRanges];
    int nstack;
    Window stack[kStack];
    Window cur;
    uint32_t cnt_below, min_above, win_min, kmin, kmax, nlist_a, nlist_b;
    int direct_eval, flags, action;   // action: 0 none, 1 collect ranges into a list
    double best_cost[32];
    uint32_t best_pos[32], best_key[32], ncand;
    float stage[NSTAGE > 0 ? NSTAGE * THREADS * kLoadBatch : 1];   // cp.async staging ring of the row passes
  };
  static_assert(sizeof(Bins) == 2 * kBins * 4, "Bins views must have the same size");
  static_assert(offsetof(Smem, bins) == (kFineCap + kCap) * 4, "list_a must run into bins.l.ext");
  static_assert(THREADS * 16 <= kCap * 4, "group prefix records must fit list A");
};
using LayoutBig = SolveLayout<13, 8192, 0, 512, 2, kMaxProChannels>;
using LayoutSmall = SolveLayout<12, 2048, 0, 256, 4, 256>;
static_assert(sizeof(LayoutBig::Smem) <= 113 * 1024, "two CTAs per SM need <= 113 KB each (228 KB - 2 x 1 KB reserved)");
static_assert(sizeof(LayoutSmall::Smem) <= 55 * 1024 + 768, "four CTAs per SM need <= 55.75 KB each");

// Calls body(v, e0, e_end) on batches of sampled elements: this thread's elements are e = e0 + u * blockDim.x
// (u < kLoadBatch, valid while e < e_end), v[u] = x[e * skip]; kLoadBatch independent loads are in flight per
// thread.  Trip counts are warp uniform.  (A cp.async.bulk ring feeding the same loop was measured slower:
// scripts/mb/mb_hist.cu, 3.4 vs 4.7 TB/s.)
template <int NSTAGE, int THREADS, class Body>
__device__ __forceinline__ void sweep_row(const float* __restrict__ xr, int skip, uint32_t n, float* stage, Body&& body) {
  const int tid = threadIdx.x;
  constexpr uint32_t kBatch = THREADS * kLoadBatch;
  if (NSTAGE == 0) {
    for (uint32_t eb = 0; eb < n; eb += kBatch) {
      float v[kLoadBatch];
      const uint32_t e0 = eb + tid;
#pragma unroll
      for (int u = 0; u < kLoadBatch; ++u) {
        const uint32_t e = e0 + u * THREADS;
        v[u] = (e < n) ? __ldg(xr + (long long)e * skip) : 0.0f;
      }
      body(v, e0, min(n, eb + kBatch));
    }
    return;
  }
  // staged: every thread copies its own elements of the next NSTAGE-1 batches into shared memory with cp.async
  // and reads back only what it copied itself, so cp.async.wait_group is the only synchronisation needed
  auto issue = [&](uint32_t eb, int slot) {
    if (eb < n) {
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(stage + slot * kBatch + tid);
#pragma unroll
      for (int u = 0; u < kLoadBatch; ++u) {
        const uint32_t e = eb + tid + u * THREADS;
        const bool ok = e < n;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst + (uint32_t)u * THREADS * 4u),
                     "l"(xr + (ok ? (long long)e * skip : 0ll)), "r"(ok ? 4 : 0) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  int islot = 0, cslot = 0;
  uint32_t ieb = 0;
  for (int k = 0; k + 1 < NSTAGE; ++k) { issue(ieb, islot); ieb += kBatch; if (++islot == NSTAGE) islot = 0; }
  for (uint32_t eb = 0; eb < n; eb += kBatch) {
    issue(ieb, islot); ieb += kBatch; if (++islot == NSTAGE) islot = 0;
    if (NSTAGE == 3) asm volatile("cp.async.wait_group 2;" ::: "memory");
    else if (NSTAGE == 2) asm volatile("cp.async.wait_group 1;" ::: "memory");
    else asm volatile("cp.async.wait_group 0;" ::: "memory");
    float v[kLoadBatch];
    const float* sp = stage + cslot * kBatch + tid;
#pragma unroll
    for (int u = 0; u < kLoadBatch; ++u) v[u] = sp[u * THREADS];
    body(v, eb + tid, min(n, eb + kBatch));
    if (++cslot == NSTAGE) cslot = 0;
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}


// One row solved by one CTA of LAY::kThreads threads (smem_raw: LAY::Smem).  v1_dup != NULL receives a second copy of
// v1 (ternary scale table).  Every thread of the CTA must call it (block barriers inside).
template <bool TERN, class LAY>
__device__ void solve_v1_row(unsigned char* smem_raw, const float* __restrict__ x, long long len, int skip, float alpha,
                             float* __restrict__ v1_out, int* __restrict__ diag, const Prologue& pro, const long long row,
                             float* __restrict__ v1_dup) {
  using SolveSmem = typename LAY::Smem;
  constexpr int kT = LAY::kThreads;
  constexpr int kBins = LAY::kBins, kBinsPerThread = LAY::kBinsPerThread, kCap = LAY::kCap, kSmallCap = LAY::kSmallCap, kTopShift = LAY::kTopShift;
  SolveSmem& sm = *reinterpret_cast<SolveSmem*>(smem_raw);
  const float* xr = x + row * len;
  const uint32_t n = (uint32_t)((len + skip - 1) / skip);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  // fused per-channel affine prologue: the table is staged in shared memory (global loads of it miss L1
  // behind the streamed row and stalled the passes)
  const bool pro_on = pro.a != nullptr, pro_smem = pro_on && pro.channels <= LAY::kProChannels;
  if (pro_smem)
    for (int c = tid; c < pro.channels; c += blockDim.x) sm.ab[c] = make_float2(__ldg(pro.a + c), __ldg(pro.b + c));
  auto prologue = [&](float v, long long index_in_row) -> float {
    if (!pro_on) return v;
    const unsigned c = prologue_channel(pro, index_in_row);
    if (pro_smem) { const float2 k = sm.ab[c]; return fmaf(v, k.x, k.y); }
    return fmaf(v, __ldg(pro.a + c), __ldg(pro.b + c));
  };
  uint32_t* const list_a = sm.list_a;      // runs on into sm.bins.l.ext for rows held entirely in shared memory
  Best best{1e300, 0xFFFFFFFFu, 0u};
  uint32_t ncand = 0;
  int passes = 0;
  uint32_t collected = 0;
  long long tacc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  long long tprev = clock64();
#define LSQ_TICK(i) do { if (diag && tid == 0) { const long long tn = clock64(); tacc[i] += tn - tprev; tprev = tn; } } while (0)

  if (tid == 0) {
    sm.kmin = kNoKey; sm.kmax = 0u; sm.nstack = 0; sm.flags = 0; sm.ncand = 0u; sm.nlist_a = 0u; sm.nlist_b = 0u;
  }
  __syncthreads();
  if (n < 3) {
    if (tid == 0) {
      v1_out[row] = 0.0f;
      if (v1_dup) v1_dup[row] = 0.0f;
      if (diag) for (int i = 0; i < 16; ++i) diag[row * 16 + i] = 0;
    }
    return;
  }

  double s_tot = 0.0, q_tot = 0.0;
  bool first = true;          // row totals not yet known
  if (n <= (uint32_t)kSmallCap) {
    // ---- small row: list A is the whole row --------------------------------------------------
    double ls = 0.0, lq = 0.0;
    uint32_t kmn = kNoKey, kmx = 0u;
    sweep_row<0, kT>(xr, skip, n, nullptr, [&](const float (&v)[kLoadBatch], uint32_t e0, uint32_t e_end) {
#pragma unroll
      for (int u = 0; u < kLoadBatch; ++u) {       // kLoadBatch loads in flight (one at a time left this phase latency bound)
        const uint32_t e = e0 + u * kT;
        if (e >= e_end) break;
        const float a = fabsf(clamp_sym(prologue(v[u], (long long)e * skip), alpha));
        const uint32_t k = __float_as_uint(a);
        ls += (double)a; lq += (double)a * (double)a;
        kmn = min(kmn, k); kmx = max(kmx, k);
        list_a[e] = k;
      }
    });
    s_tot = block_sum(ls, sm.red);
    q_tot = block_sum(lq, sm.red);
    kmn = warp_min_u32(kmn); kmx = warp_max_u32(kmx);
    if (lane == 0) { atomicMin(&sm.kmin, kmn); atomicMax(&sm.kmax, kmx); }
    __syncthreads();
    passes = 1;
    first = false;
    if (tid == 0) {
      sm.nlist_a = n;
      Window w;
      w.base.klo = 0ull; w.base.khi = 1ull << 32; w.base.sum_below = 0.0; w.base.cnt_below = 0u; w.base.next_key = sm.kmax;
      w.klo = 0u; w.shift = 31 - kListBinsLog2; w.from_list = 1;
      const uint32_t ka = __float_as_uint(alpha) >> kFineListShift;     // clamped rows: all keys <= key(alpha)
      if (alpha > 0.0f && ka + 1u >= (uint32_t)kListBins) { w.klo = (ka + 1u - kListBins) << kFineListShift; w.shift = kFineListShift; }
      sm.stack[0] = w; sm.nstack = 1;
    }
  } else if (tid == 0) {
    Window w;
    w.base.klo = 0ull; w.base.khi = 1ull << 32; w.base.sum_below = 0.0; w.base.cnt_below = 0u; w.base.next_key = 0u;
    w.klo = 0u; w.shift = kTopShift; w.from_list = 0;
    const uint32_t ka = __float_as_uint(alpha) >> kFineShift;
    if (alpha > 0.0f && ka + 1u >= (uint32_t)kBins) { w.klo = (ka + 1u - kBins) << kFineShift; w.shift = kFineShift; }
    sm.stack[0] = w; sm.nstack = 1;
  }
  __syncthreads();

  while (true) {
    __syncthreads();
    const int pending = sm.nstack;
    __syncthreads();              // every thread has read the depth before thread 0 pops
    if (pending == 0) break;
    if (tid == 0) {
      sm.cur = sm.stack[--sm.nstack];
      sm.cnt_below = 0u; sm.min_above = kNoKey; sm.win_min = kNoKey; sm.nflag = 0; sm.fmin = kNoKey; sm.fmax = 0u;
      sm.nrange = 0; sm.direct_eval = 0; sm.action = 0;
    }
    __syncthreads();              // sm.cur is published
    const Window W = sm.cur;
    const uint32_t klo = W.klo;
    const int shift = W.shift;
    const bool from_list = W.from_list != 0;
    // bin arrays of this window: counts, low sums, (list windows) high sums
    uint32_t* const hist = from_list ? sm.bins.l.hist : sm.bins.g.hist;
    uint32_t* const blo_sum = from_list ? sm.bins.l.lo : sm.bins.g.bsum;
    uint32_t* const bhi_sum = sm.bins.l.hi;
    if (from_list) {
      for (int b = tid; b < kListBins; b += blockDim.x) { hist[b] = 0u; blo_sum[b] = 0u; bhi_sum[b] = 0u; }
    } else {
      for (int b = tid; b < kBins; b += blockDim.x) { hist[b] = 0u; blo_sum[b] = 0u; }
    }
    __syncthreads();
    LSQ_TICK(0);   // pop + zero
    const int nbins = from_list ? kListBins : kBins;
    const unsigned long long khi = min((unsigned long long)klo + ((unsigned long long)nbins << shift), W.base.khi);
    // how bin sums are known: 2 = exact integers (list windows), 1 = mantissa sums truncated to 14 bits
    // (global windows, relative error < 3.1e-5), 0 = counts only (rows too long for 32-bit sums)
    const int sum_mode = from_list ? 2 : (n <= kMaxExactN ? 1 : 0);
    const uint32_t khi_incl = (uint32_t)min(khi - 1ull, 0xFFFFFFFFull);
    const uint32_t base_lo = (uint32_t)W.base.klo, base_hi_incl = (uint32_t)min(W.base.khi - 1ull, 0xFFFFFFFFull);

    // ---- histogram pass over the window's source ---------------------------------------------
    double ls = 0.0, lq = 0.0, lb = 0.0;
    uint32_t kmn = kNoKey, kmx = 0u, cb = 0u, mab = kNoKey, kwin = kNoKey;
    if (!from_list) {
      sweep_row<LAY::kStages, kT>(xr, skip, n, sm.stage,
                [&](const float (&v)[kLoadBatch], uint32_t e0, uint32_t e_end) {
#pragma unroll
        for (int u = 0; u < kLoadBatch; ++u) {
          const uint32_t e = e0 + u * kT;
          if (e >= e_end) break;
          const float a = fabsf(clamp_sym(prologue(v[u], (long long)e * skip), alpha));
          const uint32_t k = __float_as_uint(a);
          if (first) { ls += (double)a; lq += (double)a * (double)a; kmn = min(kmn, k); kmx = max(kmx, k); }
          if (k < klo) { ++cb; lb += (double)a; }
          else if (k > khi_incl) { mab = min(mab, k); }
          else {
            const uint32_t b = (k - klo) >> shift;
            kwin = min(kwin, k);
            atomicAdd(&hist[b], 1u);
            if (sum_mode == 1) atomicAdd(&blo_sum[b], (k & 0x7FFFFFu) >> 9);
          }
        }
      });
      ++passes;
    } else {
      const uint32_t la = sm.nlist_a;
      for (uint32_t e = tid; e < la; e += blockDim.x) {
        const uint32_t k = list_a[e];
        if (k < base_lo || k > base_hi_incl) continue;
        if (k < klo) { ++cb; lb += (double)key_val(k); }
        else if (k > khi_incl) { mab = min(mab, k); }
        else {
          const uint32_t b = (k - klo) >> shift;
          const uint32_t m = k & 0x7FFFFFu;
          kwin = min(kwin, k);
          atomicAdd(&hist[b], 1u);
          atomicAdd(&blo_sum[b], m & 0xFFFu);
          atomicAdd(&bhi_sum[b], m >> 12);
        }
      }
    }
    if (from_list) LSQ_TICK(2); else LSQ_TICK(1);   // histogram pass (global / list)
    if (first) {
      s_tot = block_sum(ls, sm.red);
      q_tot = block_sum(lq, sm.red);
      kmn = warp_min_u32(kmn); kmx = warp_max_u32(kmx);
      if (lane == 0) { atomicMin(&sm.kmin, kmn); atomicMax(&sm.kmax, kmx); }
      first = false;
    }
    // sum of everything below the window.  Global rows: per-thread order is fixed -> deterministic.
    // List windows normally start at their base span (nothing below, lb = 0); only the rare child of a
    // list window sums list elements here, in list order.
    const double sum_below_w = W.base.sum_below + block_sum(lb, sm.red);
    cb = (uint32_t)__reduce_add_sync(0xffffffffu, cb);
    mab = warp_min_u32(mab);
    kwin = warp_min_u32(kwin);
    if (lane == 0) { atomicAdd(&sm.cnt_below, cb); atomicMin(&sm.min_above, mab); atomicMin(&sm.win_min, kwin); }
    __syncthreads();
    const uint32_t kmax = sm.kmax;
    const uint32_t base_next = (W.base.next_key != 0u) ? W.base.next_key : kmax;   // top window: maximum not known at push time
    const uint32_t min_above = (sm.min_above != kNoKey) ? sm.min_above : base_next;
    const uint32_t cnt_below_w = W.base.cnt_below + sm.cnt_below;

    LSQ_TICK(3);   // reductions
    // ---- S1: prefix counts / sums over the bins (thread owns bpt consecutive bins) ------------------
    const int bpt = from_list ? (kListBins / kT) : kBinsPerThread;
    auto bin_sum = [&](uint32_t b, uint32_t cnt) -> double {
      if (shift == 0) return (double)cnt * (double)key_val(klo + b);
      if (sum_mode == 2) return exact_bin_sum(klo + (b << shift), cnt, blo_sum[b], bhi_sum[b]);
      if (sum_mode == 1) return approx_bin_sum(klo + (b << shift), cnt, blo_sum[b]);
      const unsigned long long e1 = min((unsigned long long)klo + ((unsigned long long)(b + 1) << shift) - 1ull, 0x7F800000ull);
      return 0.5 * (double)cnt * ((double)key_val(klo + (b << shift)) + (double)key_val((uint32_t)e1));
    };
    uint32_t ct = 0, nz = 0, fn = kNoKey;
    double stt = 0.0;
    for (int j = 0; j < bpt; ++j) {
      const uint32_t b = tid * bpt + j, cnt = hist[b];
      if (cnt != 0u) {
        ct += cnt; stt += bin_sum(b, cnt); ++nz;
        if (fn == kNoKey) fn = b;
      }
    }
    uint32_t ci = ct, zi = nz;
    double si = stt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t tc = __shfl_up_sync(0xffffffffu, ci, o);
      const uint32_t tz = __shfl_up_sync(0xffffffffu, zi, o);
      const double ts = __shfl_up_sync(0xffffffffu, si, o);
      if (lane >= o) { ci += tc; zi += tz; si += ts; }
    }
    uint32_t sfx = fn;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_down_sync(0xffffffffu, sfx, o);
      if (lane + o < 32) sfx = min(sfx, t);
    }
    uint32_t nxt_in_warp = __shfl_down_sync(0xffffffffu, sfx, 1);
    if (lane == 31) nxt_in_warp = kNoKey;
    __syncthreads();
    if (lane == 31) { sm.wcnt[wid] = ci; sm.wsum[wid] = si; sm.wnz[wid] = zi; }
    if (lane == 0) sm.wfirst[wid] = sfx;
    __syncthreads();
    uint32_t coff = 0;
    double soff = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
      if (w < wid) { coff += sm.wcnt[w]; soff += sm.wsum[w]; }
    }
    uint32_t nxt_after = nxt_in_warp;
    for (int w = wid + 1; w < (int)(blockDim.x >> 5) && nxt_after == kNoKey; ++w) nxt_after = sm.wfirst[w];
    const uint32_t excl0 = cnt_below_w + coff + ci - ct;
    const double pref0 = sum_below_w + soff + si - stt;

    // ---- S2: flag the bins that can hold a candidate ------------------------------------------------
    // relative uncertainty of the prefix sums (see sum_mode); the fp32 threshold arithmetic adds eps (may_hold)
    const float marg = (shift == 0 || sum_mode == 2) ? 0.0f : (sum_mode == 1 ? 1e-4f : 0.02f);
    auto bin_upper = [&](uint32_t b) -> float {      // largest value a key of bin b can have
      const unsigned long long nh = min((unsigned long long)klo + ((unsigned long long)(b + 1) << shift) - 1ull, (unsigned long long)kmax);
      const unsigned long long nl = min((unsigned long long)klo + ((unsigned long long)b << shift), 0x7F800000ull);
      return fmaxf(key_val((uint32_t)nh), key_val((uint32_t)nl));
    };
    auto test_bin = [&](uint32_t b, uint32_t cnt, double s, uint32_t excl, double pref, uint32_t nb) {
      const unsigned long long elo_k = min((unsigned long long)klo + ((unsigned long long)b << shift), 0x7F800000ull);
      const unsigned long long ehi_k = min((unsigned long long)klo + ((unsigned long long)(b + 1) << shift) - 1ull, (unsigned long long)kmax);
      const float nxt_hi = (nb != kNoKey) ? bin_upper(nb) : key_val(min_above);
      if (!may_hold<TERN>((uint32_t)elo_k, (uint32_t)ehi_k, cnt, s, excl, pref, nxt_hi, n, s_tot, kmax, marg)) return;
      const int slot = atomicAdd(&sm.nflag, 1);
      atomicMin(&sm.fmin, b); atomicMax(&sm.fmax, b);
      if (slot < kMaxFlag) {
        sm.fbin[slot] = (uint16_t)b; sm.fexcl[slot] = excl; sm.fsumb[slot] = pref;
        sm.fnext[slot] = (nb != kNoKey) ? (uint32_t)min((unsigned long long)klo + ((unsigned long long)nb << shift), 0xFFFFFFFEull) : min_above;
      } else if (shift == 0) {
        const uint32_t kv = klo + b;
        const uint32_t knx = (nb != kNoKey) ? klo + nb : min_above;
        for (uint32_t jj = 0; jj < cnt; ++jj)
          try_position<TERN>(best, ncand, kv, (jj + 1 < cnt) ? kv : knx, excl + jj,
                             pref + (double)(jj + 1) * (double)key_val(kv), n, s_tot, q_tot);
      }
    };
    if (from_list) {
      // two bins per thread: tested in place (list A is the source of this window and stays intact)
      uint32_t excl = excl0;
      double pref = pref0;
      for (int j = 0; j < bpt; ++j) {
        const uint32_t b = tid * bpt + j, cnt = hist[b];
        if (cnt == 0u) continue;
        uint32_t nb = kNoKey;
        for (int j2 = bpt - 1; j2 > j; --j2)
          if (hist[tid * bpt + j2] != 0u) nb = tid * bpt + j2;
        if (nb == kNoKey) nb = nxt_after;
        const double s = bin_sum(b, cnt);
        test_bin(b, cnt, s, excl, pref, nb);
        excl += cnt; pref += s;
      }
    } else {
      // Two levels: every thread tests the union of its own bins as one coarse bin (one balanced test per
      // thread; may_hold is conservative for any bin width); only the bins of the few flagged groups get the
      // fine test, spread over the block.  Group prefixes are parked in the idle list A.
      uint32_t* const gexcl = list_a;
      uint32_t* const gnext = list_a + kT;
      double* const gpref = reinterpret_cast<double*>(list_a + 2 * kT);
      if (tid == 0) sm.ngroup = 0;
      __syncthreads();
      gexcl[tid] = excl0; gnext[tid] = nxt_after; gpref[tid] = pref0;
      if (nz != 0u) {
        const unsigned long long elo_k = min((unsigned long long)klo + ((unsigned long long)(tid * bpt) << shift), 0x7F800000ull);
        const unsigned long long ehi_k = min((unsigned long long)klo + ((unsigned long long)(tid * bpt + bpt) << shift) - 1ull, (unsigned long long)kmax);
        const float nxt_hi = (nxt_after != kNoKey) ? bin_upper(nxt_after) : key_val(min_above);
        if (may_hold<TERN>((uint32_t)elo_k, (uint32_t)ehi_k, ct, stt, excl0, pref0, nxt_hi, n, s_tot, kmax, marg)) {
          const int slot = atomicAdd(&sm.ngroup, 1);
          if (slot < kMaxGroups) sm.glist[slot] = (uint16_t)tid;
          else {
            // more flagged groups than the list holds (not seen in practice): the owner tests its bins itself
            uint32_t excl = excl0;
            double pref = pref0;
            for (int j = 0; j < bpt; ++j) {
              const uint32_t b = tid * bpt + j, cnt = hist[b];
              if (cnt == 0u) continue;
              uint32_t nb = kNoKey;
              for (int j2 = bpt - 1; j2 > j; --j2)
                if (hist[tid * bpt + j2] != 0u) nb = tid * bpt + j2;
              if (nb == kNoKey) nb = nxt_after;
              const double s = bin_sum(b, cnt);
              test_bin(b, cnt, s, excl, pref, nb);
              excl += cnt; pref += s;
            }
          }
        }
      }
      __syncthreads();
      const uint32_t nfine = (uint32_t)min(sm.ngroup, kMaxGroups) * (uint32_t)bpt;
      for (uint32_t idx = tid; idx < nfine; idx += blockDim.x) {
        const uint32_t gid = sm.glist[idx / bpt], j = idx % bpt;
        const uint32_t b = gid * bpt + j, cnt = hist[b];
        if (cnt == 0u) continue;
        uint32_t excl = gexcl[gid];
        double pref = gpref[gid];
        for (uint32_t j1 = 0; j1 < j; ++j1) {
          const uint32_t c1 = hist[gid * bpt + j1];
          if (c1 != 0u) { excl += c1; pref += bin_sum(gid * bpt + j1, c1); }
        }
        uint32_t nb = kNoKey;
        for (uint32_t j2 = bpt - 1; j2 > j; --j2)
          if (hist[gid * bpt + j2] != 0u) nb = gid * bpt + j2;
        if (nb == kNoKey) nb = gnext[gid];
        test_bin(b, cnt, bin_sum(b, cnt), excl, pref, nb);
      }
      __syncthreads();
    }
    // the part of the base span below the window (anchored top windows) is one pseudo bin: if it could hold
    // a candidate it gets its own window (never seen on clamped activations; kept for exactness)
    if (tid == 0 && sm.cnt_below != 0u && (unsigned long long)klo > W.base.klo && shift != 0) {
      uint32_t fb = kNoKey;
      for (int w = 0; w < (int)(blockDim.x >> 5) && fb == kNoKey; ++w) fb = sm.wfirst[w];
      const float nxt_hi = (fb != kNoKey) ? bin_upper(fb) : key_val(min_above);
      if (may_hold<TERN>((uint32_t)W.base.klo, klo - 1u, sm.cnt_below, sum_below_w - W.base.sum_below, W.base.cnt_below,
                         W.base.sum_below, nxt_hi, n, s_tot, kmax, 0.0f)) {
        if (sm.nstack < kStack) {
          Window ch;
          ch.base = W.base; ch.base.khi = klo; ch.base.next_key = (sm.win_min != kNoKey) ? sm.win_min : min_above;
          ch.klo = (uint32_t)W.base.klo; ch.shift = from_list ? 31 - kListBinsLog2 : kTopShift; ch.from_list = W.from_list;
          sm.stack[sm.nstack++] = ch;
        } else {
          sm.flags |= 1;
        }
      }
    }
    __syncthreads();

    if (sm.nflag > 1 && sm.nflag <= kMaxFlag && tid < sm.nflag) {
      const uint32_t mine = sm.fbin[tid];
      int rank = 0;
      for (int i = 0; i < sm.nflag; ++i) rank += (sm.fbin[i] < mine) ? 1 : 0;   // bins are distinct
      sm.ford[rank] = tid;
    } else if (sm.nflag == 1 && tid == 0) {
      sm.ford[0] = 0;
    }
    __syncthreads();
    LSQ_TICK(4);   // scan + flag
    // ---- thread 0: flagged bins -> ranges -> collect / refine ---------------------------------------
    if (tid == 0 && sm.nflag > 0) {
      const int nf = min(sm.nflag, kMaxFlag);
      if (shift == 0) {
        sm.direct_eval = nf;
      } else {
        int nr = 0;
        Range* R = sm.rng;
        if (sm.nflag > kMaxFlag) {
          uint32_t cbw = cnt_below_w, cnt = 0;
          for (uint32_t b = 0; b < sm.fmin; ++b) cbw += hist[b];
          for (uint32_t b = sm.fmin; b <= sm.fmax; ++b) cnt += hist[b];
          R[0].blo = sm.fmin; R[0].bhi = sm.fmax; R[0].cnt_below = cbw; R[0].count = cnt;
          nr = 1;
        } else {
          const int* ord = sm.ford;
          // consecutive non-empty flagged bins form one range (kept in place in fexcl/fnext as scratch)
          uint32_t r_end = 0;
          for (int i = 0; i < nf; ++i) {
            const int sl = ord[i];
            const uint32_t b = sm.fbin[sl], ex = sm.fexcl[sl], en = ex + hist[b];
            if (nr > 0 && r_end == ex && nr <= kMaxRanges) { R[nr - 1].bhi = b; R[nr - 1].count = en - R[nr - 1].cnt_below; }
            else if (nr < kMaxRanges) { R[nr].blo = b; R[nr].bhi = b; R[nr].cnt_below = ex; R[nr].count = en - ex; ++nr; }
            else {
              // more than kMaxRanges runs: extend the last range over the gap (it then also holds unflagged bins)
              R[nr - 1].bhi = b; R[nr - 1].count = en - R[nr - 1].cnt_below;
            }
            r_end = en;
          }
        }
        // collect what fits into the destination list, refine the rest in child windows
        const uint32_t cap = from_list ? (uint32_t)kFineCap : (uint32_t)kCap;
        uint32_t budget = cap, lstart = 0;
        int nc = 0;
        for (int g = 0; g < nr; ++g) {
          Range r = R[g];
          r.span.klo = (unsigned long long)klo + ((unsigned long long)r.blo << shift);
          r.span.khi = min((unsigned long long)klo + ((unsigned long long)(r.bhi + 1) << shift), W.base.khi);
          r.span.cnt_below = r.cnt_below; r.span.sum_below = 0.0; r.span.next_key = kNoKey;
          if (r.count <= budget) {
            budget -= r.count;
            r.list_start = lstart; lstart += r.count;
            R[nc++] = r;
          } else {
            const unsigned long long span = r.span.khi - r.span.klo;
            int lg = 0;
            while ((1ull << lg) < span) ++lg;
            int nshift = lg - (from_list ? kListBinsLog2 : LAY::kBinsLog2);
            if (nshift < 0) nshift = 0;
            if (nshift >= shift) nshift = shift - 1;
            const unsigned long long wspan = (unsigned long long)nbins << nshift;
            for (unsigned long long o = 0; o < span; o += wspan) {
              if (sm.nstack < kStack) {
                Window ch;
                ch.base = W.base; ch.klo = (uint32_t)(r.span.klo + o); ch.shift = nshift; ch.from_list = W.from_list;
                sm.stack[sm.nstack++] = ch;
              } else {
                sm.flags |= 1;
              }
            }
          }
        }
        sm.nrange = nc;
        sm.action = nc > 0 ? 1 : 0;
      }
    }
    __syncthreads();

    LSQ_TICK(5);   // thread-0 bookkeeping
    // ---- shift 0: runs of equal values straight from the histogram --------------------------------
    if (sm.direct_eval > 0) {
      const int nf = sm.direct_eval;
      for (int sl = 0; sl < nf; ++sl) {
        const uint32_t b = sm.fbin[sl], cnt = hist[b], kv = klo + b;
        const double base = sm.fsumb[sl], v = (double)key_val(kv);
        for (uint32_t jj = tid; jj < cnt; jj += blockDim.x)
          try_position<TERN>(best, ncand, kv, (jj + 1 < cnt) ? kv : sm.fnext[sl], sm.fexcl[sl] + jj,
                             base + (double)(jj + 1) * v, n, s_tot, q_tot);
      }
    }

    // ---- collection ----------------------------------------------------------------------------------
    if (sm.action == 1) {
      const int nc = sm.nrange;
      double sb[kMaxRanges];
      uint32_t ma[kMaxRanges];
#pragma unroll
      for (int g = 0; g < kMaxRanges; ++g) { sb[g] = 0.0; ma[g] = kNoKey; }
      if (tid == 0) { if (from_list) sm.nlist_b = 0u; else sm.nlist_a = 0u; }
      uint32_t rlo[kMaxRanges], rhi[kMaxRanges];             // [rlo, rhi] inclusive, 32-bit keys
#pragma unroll
      for (int g = 0; g < kMaxRanges; ++g) {
        rlo[g] = (g < nc) ? (uint32_t)sm.rng[g].span.klo : kNoKey;
        rhi[g] = (g < nc) ? (uint32_t)(sm.rng[g].span.khi - 1ull) : kNoKey;
      }
      __syncthreads();
      if (!from_list) {
        sweep_row<LAY::kStages, kT>(xr, skip, n, sm.stage,
                  [&](const float (&v)[kLoadBatch], uint32_t e0, uint32_t e_end) {
          uint32_t matched = 0u;
          uint32_t keys[kLoadBatch];
#pragma unroll
          for (int u = 0; u < kLoadBatch; ++u) {
            if (e0 + u * kT >= e_end) break;
            const float a = fabsf(clamp_sym(prologue(v[u], (long long)(e0 + u * kT) * skip), alpha));
            const uint32_t k = __float_as_uint(a);
            keys[u] = k;
#pragma unroll
            for (int g = 0; g < kMaxRanges; ++g) {
              if (g >= nc) continue;
              if (k < rlo[g]) sb[g] += (double)a;
              else if (k <= rhi[g]) matched |= 1u << u;
              else ma[g] = min(ma[g], k);
            }
          }
          // one shared-memory atomic per warp and batch: slots from a warp prefix sum of the match counts
          const uint32_t cnt = __popc(matched);
          uint32_t inc = cnt;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
          }
          const uint32_t total = __shfl_sync(0xffffffffu, inc, 31);
          if (total != 0u) {
            uint32_t base = 0u;
            if (lane == 31) base = atomicAdd(&sm.nlist_a, total);
            base = __shfl_sync(0xffffffffu, base, 31);
            uint32_t pos = base + inc - cnt;
#pragma unroll
            for (int u = 0; u < kLoadBatch; ++u)
              if ((matched >> u) & 1u) {
                if (pos < (uint32_t)kCap) list_a[pos] = keys[u];
                ++pos;
              }
          }
        });
        ++passes;
      } else {
        const uint32_t la = sm.nlist_a;
        for (uint32_t e = tid; e < la; e += blockDim.x) {
          const uint32_t k = list_a[e];
          if (k < base_lo || k > base_hi_incl) continue;
#pragma unroll
          for (int g = 0; g < kMaxRanges; ++g) {
            if (g >= nc) continue;
            if (k >= rlo[g] && k <= rhi[g]) {
              const uint32_t slot = atomicAdd(&sm.nlist_b, 1u);
              if (slot < (uint32_t)kFineCap) sm.list_b[slot] = k;
            } else if (k > rhi[g]) ma[g] = min(ma[g], k);
          }
        }
      }
      if (from_list) LSQ_TICK(7); else LSQ_TICK(6);   // collection pass (global / list)
#pragma unroll
      for (int g = 0; g < kMaxRanges; ++g) {
        if (g < nc) {
          const double t = from_list ? 0.0 : block_sum(sb[g], sm.red);
          uint32_t m = warp_min_u32(ma[g]);
          __syncthreads();
          if (tid == 0) { sm.rng[g].span.sum_below = W.base.sum_below + t; sm.rng[g].span.next_key = kNoKey; }
          __syncthreads();
          if (lane == 0) atomicMin(&sm.rng[g].span.next_key, m);
        }
      }
      __syncthreads();
      if (tid == 0) {
        for (int g = 0; g < nc; ++g)
          if (sm.rng[g].span.next_key == kNoKey) sm.rng[g].span.next_key = base_next;
      }
      __syncthreads();
      LSQ_TICK(8);   // collection reductions
      if (!from_list) {
        // ranges of the global row are now in list A
        const uint32_t L = min(sm.nlist_a, (uint32_t)kCap);   // shadows the layout parameter in this scope
        if (sm.nlist_a <= (uint32_t)kFineCap) {
          // few elements (the usual case with a fine top window): sort them -- every range becomes one
          // ascending segment -- and give each element the reference's own candidate test
          uint32_t lp = 2;
          while (lp < L) lp <<= 1;
          for (uint32_t e = L + tid; e < lp; e += blockDim.x) list_a[e] = kNoKey;
          __syncthreads();
          bitonic_sort(list_a, lp);
          evaluate_list<TERN>(sm, list_a, L, nc, n, s_tot, q_tot, best, ncand);
          LSQ_TICK(9);   // sort + evaluate
        } else if (tid == 0) {
          // too many: refine each range in shared memory with exact integer bin sums
          if (sm.nlist_a > (uint32_t)kCap) { sm.flags |= 2; sm.nlist_a = kCap; }
          for (int g = nc - 1; g >= 0; --g) {
            const Range& r = sm.rng[g];
            const unsigned long long span = r.span.khi - r.span.klo;
            int lg = 0;
            while ((1ull << lg) < span) ++lg;
            int nshift = lg - kListBinsLog2;
            if (nshift < 0) nshift = 0;
            if (sm.nstack < kStack) {
              Window ch;
              ch.base = r.span; ch.klo = (uint32_t)r.span.klo; ch.shift = nshift; ch.from_list = 1;
              sm.stack[sm.nstack++] = ch;
            } else {
              sm.flags |= 1;
            }
          }
        }
        collected += L;
      } else {
        // list B holds the few elements that can be candidates: exact sums below each range come from
        // the bins of this window (integers), then sort and test every element
        const uint32_t L = min(sm.nlist_b, (uint32_t)kFineCap);
        if (tid == 0 && sm.nlist_b > (uint32_t)kFineCap) sm.flags |= 2;
        // sum below range g = window prefix at bin blo: recompute with one warp-free pass per range
        for (int g = 0; g < nc; ++g) {
          double part = 0.0;
          const uint32_t blo = sm.rng[g].blo;
          for (uint32_t b = tid; b < blo; b += blockDim.x) {
            const uint32_t cc = hist[b];
            if (cc) part += exact_bin_sum(klo + (b << shift), cc, blo_sum[b], bhi_sum[b]);
          }
          const double t = block_sum(part, sm.red);
          __syncthreads();
          if (tid == 0) sm.rng[g].span.sum_below = sum_below_w + t;
        }
        uint32_t lp = 2;
        while (lp < L) lp <<= 1;
        for (uint32_t e = L + tid; e < lp; e += blockDim.x) sm.list_b[e] = kNoKey;
        __syncthreads();
        bitonic_sort(sm.list_b, lp);
        evaluate_list<TERN>(sm, sm.list_b, L, nc, n, s_tot, q_tot, best, ncand);
        LSQ_TICK(9);   // exact range sums + sort + evaluate
      }
    }
  }

  // ---- reduce the best candidate over the block ------------------------------------------------
  __syncthreads();
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
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) b.offer(sm.best_cost[w], sm.best_pos[w], sm.best_key[w]);
    uint32_t nc_tot = sm.ncand;
    if (TERN) {
      // optimal.py:86-118: when min > mean/2 the value mean/2 (not a data element) is appended last
      const float mean = (float)(s_tot / (double)n);
      const float half_mean = __fmul_rn(0.5f, mean);
      if (key_val(sm.kmin) > half_mean) {
        ++nc_tot;
        b.offer(closed_cost2<true>((double)half_mean, 0.0, 0.0, (double)n, s_tot, q_tot), n, __float_as_uint(half_mean));
      }
    }
    v1_out[row] = (nc_tot > 0) ? key_val(b.key) : 0.0f;
    if (v1_dup) v1_dup[row] = (nc_tot > 0) ? key_val(b.key) : 0.0f;
    if (diag) {
      diag[row * 16 + 0] = passes; diag[row * 16 + 1] = (int)collected; diag[row * 16 + 2] = (int)nc_tot;
      diag[row * 16 + 3] = sm.flags;
      for (int i = 0; i < 10; ++i) diag[row * 16 + 4 + i] = (int)tacc[i];
      diag[row * 16 + 14] = (int)(clock64() - tprev); diag[row * 16 + 15] = 0;
    }
  }
}

template <bool TERN, class LAY>
__global__ void __launch_bounds__(LAY::kThreads, LAY::kMinBlocks)
solve_v1_kernel(const float* __restrict__ x, long long len, int skip, float alpha, float* __restrict__ v1_out,
                int* __restrict__ diag, Prologue pro, const int* __restrict__ row_status, float* __restrict__ v1_dup) {
  // row_status != NULL: only the rows marked non-zero are solved
  if (row_status && row_status[blockIdx.x] == 0) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  solve_v1_row<TERN, LAY>(smem_raw, x, len, skip, alpha, v1_out, diag, pro, (long long)blockIdx.x, v1_dup);
}

// Fallback of the fused activation quantizer (lsq_qact.cu): ONE small launch whose CTAs walk the samples, skip the
// unmarked ones (normally all) and redo a marked sample with the generic solver above followed by the encoder's work
// items -- the same arithmetic as lsq_solve_v1_ex + lsq_encode_act_ex (nscales = 1, nplanes = 2).
template <bool TERN, int VEC>
__global__ void __launch_bounds__(LayoutBig::kThreads, 1)
qact_fallback_kernel(const float* __restrict__ x, ActGeom g, long long len, int skip, float alpha, Prologue pro,
                     uint32_t* __restrict__ planes, float* __restrict__ scales, const int* __restrict__ row_status) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* const ab = reinterpret_cast<float2*>(smem_raw + ((sizeof(LayoutBig::Smem) + 15) & ~(size_t)15));      // encoder's (scale, shift) table
  double* const red = reinterpret_cast<double*>(ab + g.cw * 32);
  const uint32_t hw = (uint32_t)(g.h * g.w), nq = VEC == 4 ? hw / 4 : hw, nitems = nq * (uint32_t)g.cw;
  bool table_ready = false;
  for (long long row = blockIdx.x; row < g.n; row += gridDim.x) {
    if (row_status[row] == 0) continue;                 // uniform over the CTA
    if (!table_ready) {
      for (int c = threadIdx.x; c < g.cw * 32; c += blockDim.x) {
        float2 k = make_float2(1.0f, 0.0f);
        if (pro.a && c < g.c) k = make_float2(__ldg(pro.a + c), __ldg(pro.b + c) + 0.0f);
        ab[c] = k;
      }
      table_ready = true;
    }
    __syncthreads();
    solve_v1_row<TERN, LayoutBig>(smem_raw, x, len, skip, alpha, scales, nullptr, pro, row, TERN ? scales + g.n : nullptr);
    __syncthreads();
    const float v1 = *reinterpret_cast<volatile float*>(scales + row);       // written by thread 0 before the barrier
    const double part = encode2_row_part<VEC>(x + row * len, g, (int)row, hw, nq, 0u, nitems, ab, alpha, v1, planes);
    if (!TERN) {
      const double tot = block_sum(part, red);
      if (threadIdx.x == 0) scales[(long long)g.n + row] = (float)(tot / (double)len);
    }
    __syncthreads();
  }
}

// ---- short rows (<= 2048 sampled elements: every conv / linear weight row with skip 3) ---------------------
// One WARP per row (round 2; round 1 used a 128-thread CTA per row whose ~60 block barriers per sort dominated):
// the sampled row is sorted in the warp's slice of shared memory with __syncwarp-only bitonic steps and every
// position gets the reference's own candidate test -- the reference algorithm itself (optimal.py:41-83), with no
// histogram machinery and no block-wide synchronisation; 4 rows per CTA, up to 28 rows per SM in flight.
constexpr int kSmallThreads = 128;
constexpr int kSmallWarps = kSmallThreads / 32;
constexpr int kSmallRow = 2048;

// Sort keys[0, lp) (lp a power of two) with one warp.  Element e belongs to lane e & 31 ("column" ownership): a bitonic
// compare-exchange at distance >= 32 pairs two elements of the SAME lane (plain loads / stores of the lane's own column, no
// synchronisation), the five shorter distances of a stage are shuffles on a register.  No lane ever touches another lane's
// column, so the whole sort needs no barrier; ~3 instructions per element and step (the network with one __syncwarp per
// step and pair addressing through shared memory spent ~10).
__device__ __forceinline__ void warp_sort_columns(uint32_t* keys, uint32_t lp, int lane) {
  const uint32_t rows = (lp + 31u) >> 5;
  for (uint32_t i = 0; i < rows; ++i) {                // stages k = 2 .. 32 inside every row of 32
    const uint32_t e = i * 32u + (uint32_t)lane;
    uint32_t r = e < lp ? keys[e] : kNoKey;
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
      for (int j = k >> 1; j > 0; j >>= 1) {
        const uint32_t o = __shfl_xor_sync(0xffffffffu, r, j);
        const bool lower = (lane & j) == 0;
        const bool up = k < 32 ? ((lane & k) == 0) : ((i & 1u) == 0u);
        r = (lower == up) ? min(r, o) : max(r, o);
      }
    }
    if (e < lp) keys[e] = r;
  }
  for (uint32_t k = 64; k <= lp; k <<= 1) {
    for (uint32_t j = k >> 1; j >= 32u; j >>= 1) {     // distances of whole rows: lane-private
      const uint32_t jj = j >> 5;
#pragma unroll 4
      for (uint32_t t = 0; t < (rows >> 1); ++t) {
        const uint32_t i = ((t & ~(jj - 1u)) << 1) | (t & (jj - 1u));
        uint32_t* const pa = keys + i * 32u + lane;
        uint32_t* const pb = pa + jj * 32u;
        const uint32_t a = *pa, b2 = *pb;
        const bool up = ((i << 5) & k) == 0u;
        const uint32_t lo = min(a, b2), hi = max(a, b2);
        *pa = up ? lo : hi;
        *pb = up ? hi : lo;
      }
    }
#pragma unroll 2
    for (uint32_t i = 0; i < rows; ++i) {              // distances 16 .. 1 inside every row
      uint32_t r = keys[i * 32u + lane];
      const bool up = ((i << 5) & k) == 0u;
#pragma unroll
      for (int j = 16; j > 0; j >>= 1) {
        const uint32_t o = __shfl_xor_sync(0xffffffffu, r, j);
        const bool lower = (lane & j) == 0;
        r = (lower == up) ? min(r, o) : max(r, o);
      }
      keys[i * 32u + lane] = r;
    }
  }
}

// One row solved by one warp.  keys = the warp's shared-memory slice (>= next power of two of the sampled length).
template <bool TERN>
__device__ __forceinline__ void solve_row_warp(uint32_t* keys, const float* __restrict__ xr, long long len, int skip,
                                               float alpha, float* __restrict__ v1_dst, float* __restrict__ v1_dup,
                                               int* __restrict__ diag_row, const Prologue& pro) {
  const uint32_t n = (uint32_t)((len + skip - 1) / skip);
  const int lane = threadIdx.x & 31;
  if (n < 3) {
    if (lane == 0) {
      *v1_dst = 0.0f;
      if (v1_dup) *v1_dup = 0.0f;
      if (diag_row) for (int i = 0; i < 16; ++i) diag_row[i] = 0;
    }
    return;
  }
  double ls = 0.0, lq = 0.0;
  uint32_t kmn = kNoKey, kmx = 0u;
  for (uint32_t e = lane; e < n; e += 32) {
    const long long idx = (long long)e * skip;
    const float a = fabsf(clamp_sym(apply_prologue(pro, __ldg(xr + idx), idx), alpha));
    const uint32_t k = __float_as_uint(a);
    ls += (double)a; lq += (double)a * (double)a;
    kmn = min(kmn, k); kmx = max(kmx, k);
    keys[e] = k;
  }
  uint32_t lp = 2;
  while (lp < n) lp <<= 1;
  for (uint32_t e = n + lane; e < lp; e += 32) keys[e] = kNoKey;
  const double s_tot = warp_sum(ls), q_tot = warp_sum(lq);
  kmn = warp_min_u32(kmn); kmx = warp_max_u32(kmx);
  __syncwarp();
  warp_sort_columns(keys, lp, lane);
  __syncwarp();
  // every position of the sorted row: contiguous chunk per lane, fp64 prefix sums by shuffle
  const uint32_t per = (n + 31u) / 32u;
  const uint32_t j0 = min((uint32_t)lane * per, n), j1 = min(j0 + per, n);
  double loc = 0.0;
  for (uint32_t j = j0; j < j1; ++j) loc += (double)key_val(keys[j]);
  double inc = loc;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  double run = inc - loc;
  Best best{1e300, 0xFFFFFFFFu, 0u};
  uint32_t ncand = 0;
  for (uint32_t j = j0; j < j1; ++j) {
    const uint32_t kj = keys[j];
    run += (double)key_val(kj);
    try_position<TERN>(best, ncand, kj, (j + 1u < n) ? keys[j + 1u] : kmx, j, run, n, s_tot, q_tot);
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
    uint32_t nc_tot = ncand;
    if (TERN) {
      // optimal.py:86-118: when min > mean/2 the value mean/2 (not a data element) is appended last
      const float mean = (float)(s_tot / (double)n);
      const float half_mean = __fmul_rn(0.5f, mean);
      if (key_val(kmn) > half_mean) {
        ++nc_tot;
        best.offer(closed_cost2<true>((double)half_mean, 0.0, 0.0, (double)n, s_tot, q_tot), n, __float_as_uint(half_mean));
      }
    }
    const float v1 = (nc_tot > 0) ? key_val(best.key) : 0.0f;
    *v1_dst = v1;
    if (v1_dup) *v1_dup = v1;
    if (diag_row) {
      for (int i = 0; i < 16; ++i) diag_row[i] = 0;
      diag_row[0] = 1; diag_row[2] = (int)nc_tot;
    }
  }
}

template <bool TERN>
__global__ void __launch_bounds__(kSmallThreads)
solve_v1_small_kernel(const float* __restrict__ x, long long rows, long long len, int skip, float alpha,
                      float* __restrict__ v1_out, int* __restrict__ diag, Prologue pro, const int* __restrict__ row_status,
                      float* __restrict__ v1_dup, int lp_cap) {
  extern __shared__ __align__(16) uint32_t small_keys[];
  const long long row = (long long)blockIdx.x * kSmallWarps + (threadIdx.x >> 5);
  if (row >= rows) return;
  if (row_status && row_status[row] == 0) return;
  solve_row_warp<TERN>(small_keys + (size_t)(threadIdx.x >> 5) * lp_cap, x + row * len, len, skip, alpha, v1_out + row,
                       v1_dup ? v1_dup + row : nullptr, diag ? diag + row * 16 : nullptr, pro);
}

// Multi-tensor launch (weight tensors of a whole network in ONE grid): the tensor table travels by value in the
// kernel parameters, every warp finds its tensor by bisection over the cumulative row counts.  Same per-row code as
// the single-tensor kernel, so the results are bit-identical.
constexpr int kMultiMax = 112;
struct MultiTab {
  const float* x[kMultiMax];
  float* out[kMultiMax];
  int len[kMultiMax];
  int first_row[kMultiMax + 1];
  int n;
};
__device__ __forceinline__ int multi_find(const MultiTab& tab, int row) {
  int lo = 0, hi = tab.n;                 // first_row[lo] <= row < first_row[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (tab.first_row[mid] <= row) lo = mid; else hi = mid;
  }
  return lo;
}
template <bool TERN>
__global__ void __launch_bounds__(kSmallThreads)
solve_v1_multi_kernel(const __grid_constant__ MultiTab tab, int skip, float alpha, int lp_cap) {
  extern __shared__ __align__(16) uint32_t small_keys[];
  const int grow = (int)blockIdx.x * kSmallWarps + (int)(threadIdx.x >> 5);
  if (grow >= tab.first_row[tab.n]) return;
  const int t = multi_find(tab, grow);
  const int r = grow - tab.first_row[t];
  const long long len = tab.len[t];
  const Prologue none{nullptr, nullptr, 1, 1, 1ull << 40};
  solve_row_warp<TERN>(small_keys + (size_t)(threadIdx.x >> 5) * lp_cap, tab.x[t] + (long long)r * len, len, skip, alpha,
                       tab.out[t] + r, nullptr, nullptr, none);
}

}  // namespace lsq

using namespace lsq;

extern "C" int lsq_solve_v1(const float* d_x, int64_t rows, int64_t len, int skip, int ternary, float alpha,
                            float* d_v1, int32_t* d_diag, void* stream) {
  return lsq_solve_v1_ex(d_x, rows, len, skip, ternary, alpha, d_v1, d_diag, nullptr, stream);
}

static int solve_v1_launch(const float* d_x, int64_t rows, int64_t len, int skip, int ternary, float alpha,
                           float* d_v1, int32_t* d_diag, const lsq_prologue* pro, const int* d_row_status,
                           float* d_v1_dup, void* stream);

extern "C" int lsq_solve_v1_ex(const float* d_x, int64_t rows, int64_t len, int skip, int ternary, float alpha,
                               float* d_v1, int32_t* d_diag, const lsq_prologue* pro, void* stream) {
  return solve_v1_launch(d_x, rows, len, skip, ternary, alpha, d_v1, d_diag, pro, nullptr, nullptr, stream);
}

namespace lsq {
int solve_v1_marked_rows(const float* d_x, int64_t rows, int64_t len, int skip, int ternary, float alpha, float* d_v1,
                         float* d_v1_dup, const lsq_prologue* pro, const int* d_row_status, cudaStream_t stream) {
  return solve_v1_launch(d_x, rows, len, skip, ternary, alpha, d_v1, nullptr, pro, d_row_status, d_v1_dup, (void*)stream);
}
}  // namespace lsq

namespace lsq {
int qact_fallback_launch(const float* d_x, const lsq_act_geom* g, float alpha, int ternary, int skip, uint32_t* d_planes,
                         float* d_scales, const lsq_prologue* pro, const int* d_row_status, bool vec4, cudaStream_t stream) {
  const int64_t len = (int64_t)g->c * g->h * g->w;
  const Prologue dp = to_dev(pro);
  const ActGeom dg = to_dev(*g);
  const size_t smem = ((sizeof(LayoutBig::Smem) + 15) & ~(size_t)15) + (size_t)g->cw * 32 * sizeof(float2) + 32 * sizeof(double);
  int grid = device_sms();
  if (grid > g->n) grid = g->n;
  cudaError_t e = cudaSuccess;
#define LSQ_FB(T, V)                                                                                               \
  do {                                                                                                             \
    static std::atomic<unsigned long long> smem_set{0ull};                                                         \
    e = ensure_max_smem(qact_fallback_kernel<T, V>, smem_set);                                                     \
    if (e == cudaSuccess)                                                                                          \
      qact_fallback_kernel<T, V><<<grid, LayoutBig::kThreads, smem, stream>>>(d_x, dg, len, skip, alpha, dp, d_planes, \
                                                                              d_scales, d_row_status);             \
  } while (0)
  if (ternary) { if (vec4) LSQ_FB(true, 4); else LSQ_FB(true, 1); }
  else { if (vec4) LSQ_FB(false, 4); else LSQ_FB(false, 1); }
#undef LSQ_FB
  if (e != cudaSuccess) { set_error("lsq_quantize_act: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return LSQ_ERR_CUDA; }
  LSQ_CUDA_LAUNCH_CHECK("qact_fallback_kernel");
  return LSQ_OK;
}
}  // namespace lsq

static int solve_v1_launch(const float* d_x, int64_t rows, int64_t len, int skip, int ternary, float alpha,
                           float* d_v1, int32_t* d_diag, const lsq_prologue* pro, const int* d_row_status,
                           float* d_v1_dup, void* stream) {
  LSQ_CHECK_ARG(d_x && d_v1, "lsq_solve_v1: null pointer");
  LSQ_CHECK_ARG(rows > 0 && len > 0 && skip >= 1, "lsq_solve_v1: bad shape rows=%lld len=%lld skip=%d", (long long)rows, (long long)len, skip);
  LSQ_CHECK_ARG((len + skip - 1) / skip < (1ll << 31), "lsq_solve_v1: row too long");
  if (pro && pro->d_ch_scale && ((int64_t)pro->channels * pro->inner != len || len >= (1ll << 31))) {
    set_error("lsq_solve_v1: prologue needs len == channels * inner (< 2^31)");
    return LSQ_ERR_ARG;
  }
  const Prologue dp = to_dev(pro);
  dim3 grid((unsigned)rows);
  if ((len + skip - 1) / skip <= (int64_t)kSmallRow) {
    int lp_cap = 2;
    while (lp_cap < (int)((len + skip - 1) / skip)) lp_cap <<= 1;
    const size_t sm_bytes = (size_t)kSmallWarps * lp_cap * sizeof(uint32_t);
    const dim3 wgrid((unsigned)((rows + kSmallWarps - 1) / kSmallWarps));
    if (ternary) solve_v1_small_kernel<true><<<wgrid, kSmallThreads, sm_bytes, (cudaStream_t)stream>>>(d_x, rows, len, skip, alpha, d_v1, d_diag, dp, d_row_status, d_v1_dup, lp_cap);
    else solve_v1_small_kernel<false><<<wgrid, kSmallThreads, sm_bytes, (cudaStream_t)stream>>>(d_x, rows, len, skip, alpha, d_v1, d_diag, dp, d_row_status, d_v1_dup, lp_cap);
    LSQ_CUDA_LAUNCH_CHECK("solve_v1_small_kernel");
    return LSQ_OK;
  }
  // rows whose sampled elements fit shared memory take the layout that keeps them there (one pass over HBM)
  // (unclamped long rows keep the 8192-bin layout: their top window cannot be anchored and stays coarse)
  const bool big = (len + skip - 1) / skip <= (int64_t)LayoutBig::kSmallCap || !(alpha > 0.0f);
  const size_t smem = big ? sizeof(LayoutBig::Smem) : sizeof(LayoutSmall::Smem);
  cudaError_t e;
#define LSQ_SOLVE(T, LAY)                                                                                          \
  do {                                                                                                             \
    static std::atomic<unsigned long long> smem_set{0ull};                                                         \
    e = ensure_max_smem(solve_v1_kernel<T, LAY>, smem_set);                                                        \
    if (e == cudaSuccess)                                                                                          \
      solve_v1_kernel<T, LAY><<<grid, LAY::kThreads, smem, (cudaStream_t)stream>>>(d_x, len, skip, alpha, d_v1, d_diag, dp, d_row_status, d_v1_dup); \
  } while (0)
  if (ternary) { if (big) LSQ_SOLVE(true, LayoutBig); else LSQ_SOLVE(true, LayoutSmall); }
  else { if (big) LSQ_SOLVE(false, LayoutBig); else LSQ_SOLVE(false, LayoutSmall); }
#undef LSQ_SOLVE
  if (e != cudaSuccess) {
    set_error("lsq_solve_v1: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    return LSQ_ERR_CUDA;
  }
  LSQ_CUDA_LAUNCH_CHECK("solve_v1_kernel");
  return LSQ_OK;
}

// Several tensors, one launch per `kMultiMax` small-row tensors (host table, passed by value; nothing is read from
// the table after the call returns).  Tensors whose sampled rows exceed the small-row kernel get their own launch.
extern "C" int lsq_solve_v1_multi(const lsq_row_tensor* tensors, int ntensors, int skip, int ternary, float alpha,
                                  void* stream) {
  LSQ_CHECK_ARG(tensors && ntensors > 0 && skip >= 1, "lsq_solve_v1_multi: bad arguments");
  for (int i = 0; i < ntensors; ++i)
    LSQ_CHECK_ARG(tensors[i].d_x && tensors[i].d_out && tensors[i].rows > 0 && tensors[i].len > 0,
                  "lsq_solve_v1_multi: tensor %d: null pointer or empty shape", i);
  // longest rows first: the grid drains evenly
  int order[kMultiMax];
  int done = 0;
  while (done < ntensors) {
    MultiTab tab;
    tab.n = 0;
    tab.first_row[0] = 0;
    for (; done < ntensors && tab.n < kMultiMax; ++done) {
      const lsq_row_tensor& T = tensors[done];
      if (((int64_t)T.len + skip - 1) / skip > (int64_t)kSmallRow) {
        const int st = lsq_solve_v1_ex(T.d_x, T.rows, T.len, skip, ternary, alpha, T.d_out, nullptr, nullptr, stream);
        if (st != LSQ_OK) return st;
        continue;
      }
      if ((int64_t)tab.first_row[tab.n] + T.rows > (int64_t)0x7fffffff) break;
      order[tab.n] = done;
      ++tab.n;
      tab.first_row[tab.n] = tab.first_row[tab.n - 1] + T.rows;
    }
    if (tab.n == 0) continue;
    // insertion sort of the batch by row length, descending (stable)
    for (int i = 1; i < tab.n; ++i) {
      const int o = order[i];
      int j = i - 1;
      for (; j >= 0 && tensors[order[j]].len < tensors[o].len; --j) order[j + 1] = order[j];
      order[j + 1] = o;
    }
    for (int i = 0; i < tab.n; ++i) {
      const lsq_row_tensor& T = tensors[order[i]];
      tab.x[i] = T.d_x; tab.out[i] = T.d_out; tab.len[i] = T.len;
      tab.first_row[i + 1] = tab.first_row[i] + T.rows;
    }
    int lp_cap = 2;                          // the longest sampled row of this batch (tab.len is sorted descending)
    while (lp_cap < (tab.len[0] + skip - 1) / skip) lp_cap <<= 1;
    const size_t sm_bytes = (size_t)kSmallWarps * lp_cap * sizeof(uint32_t);
    dim3 grid((unsigned)((tab.first_row[tab.n] + kSmallWarps - 1) / kSmallWarps));
    if (ternary) solve_v1_multi_kernel<true><<<grid, kSmallThreads, sm_bytes, (cudaStream_t)stream>>>(tab, skip, alpha, lp_cap);
    else solve_v1_multi_kernel<false><<<grid, kSmallThreads, sm_bytes, (cudaStream_t)stream>>>(tab, skip, alpha, lp_cap);
    LSQ_CUDA_LAUNCH_CHECK("solve_v1_multi_kernel");
  }
  return LSQ_OK;
}
