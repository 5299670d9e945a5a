// Device-side building blocks of the v1 solve shared by lsq_solve.cu (one CTA per row, generic) and lsq_qact.cu
// (fused activation quantizer): the reference's own candidate test in fp32, the closed-form cost, the conservative
// bin test, the in-shared-memory sort and the evaluation of a sorted list.  See lsq_solve.cu for the algorithm.
#pragma once
#include "lsq_common.cuh"

namespace lsq {

constexpr int kMaxRanges = 4;
constexpr uint32_t kNoKey = 0xFFFFFFFFu;

// A contiguous run of the sorted row: keys in [klo, khi), preceded by cnt_below elements whose
// exact sum is sum_below, followed by next_key (smallest key >= khi, or the row maximum).
struct Span {
  unsigned long long klo, khi;
  double sum_below;
  uint32_t cnt_below, next_key;
};
struct Range {        // flagged bins [blo, bhi] of the current window
  uint32_t blo, bhi, cnt_below, count;
  Span span;
  uint32_t list_start;
};

__device__ __forceinline__ float key_val(uint32_t k) { return __uint_as_float(k); }

// 2^e as a double (-1022 <= e <= 1023): scaling by it is exact, like ldexp
__device__ __forceinline__ double pow2d(int e) { return __hiloint2double((e + 1023) << 20, 0); }

// exact value sum of `cnt` keys sharing one exponent: lo = sum(m & 0xFFF), hi = sum(m >> 12)
__device__ __forceinline__ double exact_bin_sum(uint32_t any_key, uint32_t cnt, uint32_t lo, uint32_t hi) {
  const int e = (int)(any_key >> 23);
  const double msum = (double)hi * 4096.0 + (double)lo;
  if (e == 0) return msum * pow2d(-149);
  return ((double)cnt * 8388608.0 + msum) * pow2d(e - 150);
}

// value sum of `cnt` keys sharing one exponent from s9 = sum(m >> 9): midpoint of the possible range,
// relative error < 2^8 / 2^23 = 3.1e-5
__device__ __forceinline__ double approx_bin_sum(uint32_t any_key, uint32_t cnt, uint32_t s9) {
  const int e = (int)(any_key >> 23);
  const double msum = (double)s9 * 512.0 + 256.0 * (double)cnt;
  if (e == 0) return msum * pow2d(-149);
  return ((double)cnt * 8388608.0 + msum) * pow2d(e - 150);
}

struct Best {
  double cost;
  uint32_t pos, key;
  __device__ void offer(double c, uint32_t p, uint32_t k) {
    if (c < cost || (c == cost && p < pos)) { cost = c; pos = p; key = k; }
  }
};

// fp32 emulation of optimal.py:56-80 at sorted position i = k-1 (1 <= i <= n-2)
template <bool TERN>
__device__ __forceinline__ bool is_candidate(float a_i, float a_next, uint32_t k, double s_i, uint32_t n, float tot) {
  const float cum = (float)s_i;
  const float m2 = __fdiv_rn(__fsub_rn(tot, cum), (float)(n - k));
  const float half = __fmul_rn(0.5f, m2);
  bool ok = (a_i <= half) && (half <= a_next);
  if (!TERN) {
    const float m1 = __fdiv_rn(cum, (float)k);
    const float mid = __fmul_rn(0.5f, __fadd_rn(m1, m2));
    ok = ok || ((a_i <= mid) && (mid <= a_next));
  }
  return ok;
}

// closed form of cost^2 (optimal.py:31-38) for candidate value c with k elements <= c
template <bool TERN>
__device__ __forceinline__ double closed_cost2(double c, double k, double s_i, double n, double s_tot, double q_tot) {
  const double sabs = (s_tot - s_i - (n - k) * c) + (k * c - s_i);
  const double sq = q_tot - 2.0 * c * s_tot + n * c * c;
  if (TERN) return sq - 2.0 * c * sabs + n * c * c;
  return sq - sabs * sabs / n;
}

// Can the bin of keys [elo_k, ehi_k] -- cnt elements summing to s, preceded in sorted order by excl elements
// summing to pref (both known to a relative `marg`) and followed by a key of value <= nxt_hi -- contain a
// position the reference accepts as a candidate (optimal.py:73-80)?  Conservative: intervals of the two
// threshold functions over the bin are compared with the bin's value range.
template <bool TERN>
__device__ __forceinline__ bool may_hold(uint32_t elo_k, uint32_t ehi_k, uint32_t cnt, double s, uint32_t excl, double pref,
                                         float nxt_hi, uint32_t n, double s_tot, uint32_t kmax, float marg) {
  const uint32_t k0 = max(excl, 1u), k1 = min(excl + cnt, n - 1);
  if (k0 > k1) return false;
  const float eps = 4e-6f;
  const float edge_lo = key_val(elo_k);
  const float edge_hi = fmaxf(key_val(ehi_k), edge_lo);
  // prefix sums at the first (k0) and last (k1) split of the bin, as [lo, hi] intervals
  const float p0 = (float)pref, p1 = (float)(pref + s), rest0 = (float)(s_tot - pref), rest1 = (float)(s_tot - pref - s);
  float lo0, hi0, r0lo, r0hi, lo1, hi1, r1lo, r1hi;
  const float dm0 = marg * p0, dm1 = marg * p1;
  if (excl >= 1u) { lo0 = p0 - dm0; hi0 = p0 + dm0; r0lo = rest0 - dm0; r0hi = rest0 + dm0; }
  else { lo0 = p0 - dm0 + edge_lo; hi0 = p0 + dm0 + edge_hi; r0lo = rest0 - dm0 - edge_hi; r0hi = rest0 + dm0 - edge_lo; }
  if (excl + cnt <= n - 1) { lo1 = p1 - dm1; hi1 = p1 + dm1; r1lo = rest1 - dm1; r1hi = rest1 + dm1; }
  else { r1lo = r1hi = key_val(kmax); lo1 = hi1 = (float)(s_tot - (double)key_val(kmax)); }
  // fast division: its 2 ulp are far inside eps
  const float ih0 = __fdividef(0.5f, (float)(n - k0)), ih1 = __fdividef(0.5f, (float)(n - k1));
  const float half_min = r0lo * ih0, half_max = r1hi * ih1;      // hi/2 is monotone in the split
  // some threshold must reach the bin from above, and either pass below a later element of the bin or -- for
  // the last element, whose threshold is at least r1lo * ih1 -- below the next key after the bin
  bool hit = (half_max * (1.0f + eps) >= edge_lo) &&
             (half_min * (1.0f - eps) <= edge_hi || r1lo * ih1 * (1.0f - eps) <= nxt_hi);
  if (!TERN) {
    const float ik0 = __fdividef(0.5f, (float)k0), ik1 = __fdividef(0.5f, (float)k1);
    const float mid_min = fminf(r0lo * ih0 + lo0 * ik0, r0hi * ih0 + hi0 * ik0);
    const float mid_max = fmaxf(r1lo * ih1 + lo1 * ik1, r1hi * ih1 + hi1 * ik1);
    // (lo, rest) move in opposite directions: the extremes over the interval are at its ends,
    // paired as (lo0, r0hi) / (hi0, r0lo); take the enclosing values to stay conservative
    const float mid_min2 = fminf(mid_min, fminf(r0hi * ih0 + lo0 * ik0, r0lo * ih0 + hi0 * ik0));
    const float mid_max2 = fmaxf(mid_max, fmaxf(r1hi * ih1 + lo1 * ik1, r1lo * ih1 + hi1 * ik1));
    hit = hit || ((mid_max2 * (1.0f + eps) >= edge_lo) &&
                  (mid_min2 * (1.0f - eps) <= edge_hi || (r1lo * ih1 + lo1 * ik1) * (1.0f - eps) <= nxt_hi));
  }
  return hit;
}

template <bool TERN>
__device__ __forceinline__ void try_position(Best& best, uint32_t& ncand, uint32_t key_i, uint32_t key_next,
                                             uint32_t i, double s_i, uint32_t n, double s_tot, double q_tot) {
  if (i < 1 || i + 2 > n) return;
  const float a_i = key_val(key_i);
  if (is_candidate<TERN>(a_i, key_val(key_next), i + 1, s_i, n, (float)s_tot)) {
    ++ncand;
    best.offer(closed_cost2<TERN>((double)a_i, (double)(i + 1), s_i, (double)n, s_tot, q_tot), i, key_i);
  }
}

__device__ __forceinline__ void bitonic_sort(uint32_t* keys, uint32_t lp) {
  for (uint32_t k = 2; k <= lp; k <<= 1) {
    for (uint32_t j = k >> 1; j > 0; j >>= 1) {
      for (uint32_t t = threadIdx.x; t < (lp >> 1); t += blockDim.x) {
        const uint32_t i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const uint32_t p = i | j;
        const uint32_t a = keys[i], b = keys[p];
        const bool up = ((i & k) == 0);
        if ((a > b) == up) { keys[i] = b; keys[p] = a; }
      }
      __syncthreads();
    }
  }
}

// Evaluate the sorted list keys[0..L) made of the `nseg` ranges sm.rng[] (ascending key order).
template <bool TERN, class SM>
__device__ void evaluate_list(SM& sm, const uint32_t* keys, uint32_t L, int nseg, uint32_t n, double s_tot,
                              double q_tot, Best& best, uint32_t& ncand) {
  const uint32_t per = (L + blockDim.x - 1) / blockDim.x;
  const uint32_t j0 = min(threadIdx.x * per, L), j1 = min(j0 + per, L);
  double loc = 0.0;
  for (uint32_t j = j0; j < j1; ++j) loc += (double)key_val(keys[j]);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  double inc = loc;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    double t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();
  if (lane == 31) sm.wsum[wid] = inc;
  __syncthreads();
  double off = 0.0;
  for (int w = 0; w < wid; ++w) off += sm.wsum[w];
  double run = off + inc - loc;  // exclusive prefix at j0
  for (int g = 0; g < nseg; ++g) {
    const uint32_t st = sm.rng[g].list_start;
    if (st >= j0 && st < j1) {
      double r = run;
      for (uint32_t j = j0; j < st; ++j) r += (double)key_val(keys[j]);
      sm.seg_base[g] = r;
    }
  }
  __syncthreads();
  for (uint32_t j = j0; j < j1; ++j) {
    const uint32_t kj = keys[j];
    run += (double)key_val(kj);
    int g = 0;
    while (g + 1 < nseg && j >= sm.rng[g + 1].list_start) ++g;
    const Range& R = sm.rng[g];
    const uint32_t seg_end = R.list_start + R.count;
    const uint32_t i = R.span.cnt_below + (j - R.list_start);
    const double s_i = R.span.sum_below + (run - sm.seg_base[g]);
    const uint32_t knext = (j + 1 < seg_end) ? keys[j + 1] : R.span.next_key;
    try_position<TERN>(best, ncand, kj, knext, i, s_i, n, s_tot, q_tot);
  }
  __syncthreads();
}

}  // namespace lsq
