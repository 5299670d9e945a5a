"""Serial numpy model of the CUDA least-squares solver (ml_quant_b200/csrc/lsq_solve.cu).

Test infrastructure: it documents, step for step, the sort-free algorithm the kernel runs per row and
lets the algorithm be validated against the oracle on the CPU (tests/test_solver_model.py) before any
GPU time is spent.  The kernel must make the same decisions (same bins, same flags, same fp32-emulated
candidate tests, same fp64 closed-form costs); only the summation order of the fp64 sums differs.

Algorithm (per row; a = |clamp(x)|[::skip], n elements, key = float bits of a, monotone in a):
  window(klo, shift): histogram the keys inside [klo, klo + NBINS<<shift) into NBINS bins,
      exact fp64 sum / count of everything below the window, min key above it;
      bound the prefix sums at bin edges (exact when shift == 0) -> bound the two threshold
      functions  half(i) = hi_i/2  and  mid(i) = (lo_i+hi_i)/2  over each bin -> flag the bins that
      can contain a reference candidate  a_i <= thr(i) <= a_{i+1}  (quant/binary/optimal.py:73-80);
      flagged runs of bins are either collected (few elements: sort + evaluate every element),
      evaluated directly (shift == 0: a bin is a run of equal values) or refined in a child window.
  evaluate(i): the reference's fp32 arithmetic on (float)prefix sums (optimal.py:56-80), so the
      candidate set equals the reference's; the cost of a candidate is the closed form of
      optimal.py:31-38 in fp64, first minimum in ascending order wins.

The kernel adds performance-only refinements on top of this logic (none changes which positions are evaluated
exactly): for clamped rows the top window is anchored at the clamp bound with 256 bins per octave and everything
below it is one pseudo bin; the bins of a window are first tested in groups (a conservative coarse test), only
flagged groups get the per-bin test; few collected elements are sorted directly instead of being refined again;
rows of up to 21.5 K elements are held in shared memory, rows of up to 2 K are simply sorted.
"""
import numpy as np

NBINS = 8192
CAP = 8192          # collected-list capacity (elements)
MAXR = 4            # flagged ranges handled by one collection pass
TOP_SHIFT = 18      # 31 key bits - 13 bin bits
F32 = np.float32


def keys_of(a):
    return a.astype(F32).view(np.uint32).astype(np.int64)


def val_of(k):
    return np.asarray(k, dtype=np.uint32).view(F32).astype(np.float64)


class Best:
    def __init__(self):
        self.cost, self.pos, self.val = np.inf, np.iinfo(np.int64).max, F32(0)
        self.ncand = 0
        self.cands = []

    def offer(self, cost, pos, val):
        self.ncand += 1
        self.cands.append(F32(val))
        if cost < self.cost or (cost == self.cost and pos < self.pos):
            self.cost, self.pos, self.val = cost, pos, F32(val)


def closed_cost2(c, k, s_i, n, s_tot, q_tot, ternary):
    """cost^2 of candidate value c sitting at sorted position k-1 (k elements <= c)."""
    c = float(c)
    sabs = (s_tot - s_i - (n - k) * c) + (k * c - s_i)
    sq = q_tot - 2.0 * c * s_tot + n * c * c
    if ternary:
        return sq - 2.0 * c * sabs + n * c * c
    return sq - sabs * sabs / n


def is_candidate(a_i, a_next, k, s_i, n, s_tot, ternary):
    """fp32 emulation of optimal.py:56-80 at sorted position i = k-1 (interior only)."""
    cum = F32(s_i)
    tot = F32(s_tot)
    m2 = F32(F32(tot - cum) / F32(n - k))
    half = F32(F32(0.5) * m2)
    ok = (F32(a_i) <= half) and (half <= F32(a_next))
    if not ternary:
        m1 = F32(cum / F32(k))
        mid = F32(F32(0.5) * F32(m1 + m2))
        ok = ok or ((F32(a_i) <= mid) and (mid <= F32(a_next)))
    return ok


def thr_bounds(k, s_lo, s_hi, n, s_tot, ternary):
    """Per threshold function, [min, max] over the admissible prefix sum at split size k (1 <= k <= n-1).
    Returns [(min_half, max_half)] or [(min_half, max_half), (min_mid, max_mid)]."""
    halves, mids = [], []
    for s in (s_lo, s_hi):
        hi = (s_tot - s) / (n - k)
        halves.append(0.5 * hi)
        mids.append(0.5 * (s / k + hi))
    out = [(min(halves), max(halves))]
    if not ternary:
        out.append((min(mids), max(mids)))
    return out


def solve_row(a, ternary, stats=None):
    a = np.asarray(a, dtype=F32)
    n = a.size
    best = Best()
    if n < 3:
        return F32(0), best
    key = keys_of(a)
    vals = a.astype(np.float64)
    s_tot = float(vals.sum())
    q_tot = float((vals * vals).sum())
    kmax = int(key.max())
    kmin = int(key.min())

    def evaluate_sorted(sk, cnt_below, sum_below, next_key):
        """sk: sorted keys of one contiguous run of the sorted row starting at position cnt_below."""
        v = val_of(sk)
        cs = sum_below + np.cumsum(v)
        for j in range(sk.size):
            i = cnt_below + j
            if i < 1 or i > n - 2:
                continue
            nk = sk[j + 1] if j + 1 < sk.size else next_key
            a_i, a_n = F32(val_of(sk[j])), F32(val_of(nk))
            if is_candidate(a_i, a_n, i + 1, cs[j], n, s_tot, ternary):
                best.offer(closed_cost2(a_i, i + 1, cs[j], n, s_tot, q_tot, ternary), i, a_i)

    if n <= CAP:
        evaluate_sorted(np.sort(key), 0, 0.0, kmax)
    else:
        stack = [(0, TOP_SHIFT)]
        while stack:
            klo, shift = stack.pop()
            khi = klo + (NBINS << shift)
            inwin = (key >= klo) & (key < khi)
            cnt_b0 = int((key < klo).sum())
            sum_b0 = float(vals[key < klo].sum())
            above = key[key >= khi]
            min_above = int(above.min()) if above.size else kmax
            hist = np.bincount(((key[inwin] - klo) >> shift), minlength=NBINS)
            if stats is not None:
                stats['passes'] = stats.get('passes', 0) + 1
            excl = np.concatenate([[0], np.cumsum(hist)[:-1]]) + cnt_b0
            bidx = np.arange(NBINS)
            edge_lo = val_of(np.minimum(klo + (bidx << shift), 0x7F800000))
            edge_hi = val_of(np.minimum(klo + ((bidx + 1) << shift) - 1, max(kmax, 0)))
            edge_hi = np.maximum(edge_hi, edge_lo)
            # per-bin sums: the kernel accumulates them with fp32 shared-memory atomics (order dependent,
            # ~1e-6 relative), so they only steer the flagging, with a 1e-4 safety margin; at shift 0 a bin
            # holds one value and count*value is exact.
            bsum = np.bincount(((key[inwin] - klo) >> shift), weights=vals[inwin], minlength=NBINS)
            if shift == 0:
                bsum = hist * np.where(hist > 0, edge_lo, 0.0)
            pref = sum_b0 + np.concatenate([[0.0], np.cumsum(bsum)])
            marg = 0.0 if shift == 0 else 1e-4
            slo, shi = pref * (1 - marg), pref * (1 + marg)           # bounds on the prefix sum below bin b
            nz = np.nonzero(hist)[0]
            flagged = []
            for t, b in enumerate(nz):
                k0 = max(int(excl[b]), 1)                    # split sizes possible inside the bin
                k1 = min(int(excl[b] + hist[b]), n - 1)
                if k0 > k1:
                    continue
                # prefix sum when k0 / k1 elements are below the split
                if excl[b] >= 1:
                    lo0, hi0 = slo[b], shi[b]
                else:                                         # k0 = 1 takes the first element of the bin
                    lo0, hi0 = slo[b] + edge_lo[b], shi[b] + edge_hi[b]
                if excl[b] + hist[b] <= n - 1:
                    lo1, hi1 = slo[b + 1], shi[b + 1]
                else:                                         # k1 = n-1 leaves only the maximum above
                    lo1 = hi1 = s_tot - float(val_of(kmax))
                b_lo = thr_bounds(k0, lo0, hi0, n, s_tot, ternary)
                b_hi = thr_bounds(k1, lo1, hi1, n, s_tot, ternary)
                nxt_hi = edge_hi[nz[t + 1]] if t + 1 < nz.size else float(val_of(min_above))
                eps = 1e-6
                hit = False
                for (tmin, _), (_, tmax) in zip(b_lo, b_hi):
                    # some a_i <= thr(i): thr(k1) >= L.  thr(i) <= a_{i+1}: inside the bin a_{i+1} <= U,
                    # for the last element of the bin a_{i+1} <= U_next and thr = thr(k1).
                    hit = hit or (tmax * (1 + eps) >= edge_lo[b] and
                                  (tmin * (1 - eps) <= edge_hi[b] or tmax * (1 - eps) <= nxt_hi))
                if hit:
                    flagged.append(t)
            if stats is not None:
                stats['flagged_bins'] = stats.get('flagged_bins', 0) + len(flagged)
            # group flagged entries of the compacted list into ranges of consecutive nonempty bins
            ranges = []
            for t in flagged:
                if ranges and ranges[-1][1] == t - 1:
                    ranges[-1][1] = t
                else:
                    ranges.append([t, t])
            while len(ranges) > MAXR:                          # merge the two closest ranges
                gaps = [ranges[g + 1][0] - ranges[g][1] for g in range(len(ranges) - 1)]
                g = int(np.argmin(gaps))
                ranges[g][1] = ranges[g + 1][1]
                del ranges[g + 1]
            collect, budget = [], CAP
            for t0, t1 in ranges:
                b0, b1 = int(nz[t0]), int(nz[t1])
                cnt = int(hist[b0:b1 + 1].sum())
                if shift == 0:
                    # every bin is a run of one value: evaluate the runs directly
                    for b in nz[t0:t1 + 1]:
                        nk = int(klo + nz[np.searchsorted(nz, b) + 1]) if b != nz[-1] else min_above
                        sk = np.full(int(hist[b]), klo + int(b), dtype=np.int64)
                        evaluate_sorted(sk, int(excl[b]), float(pref[b]), nk)
                elif cnt <= budget:
                    budget -= cnt
                    collect.append((b0, b1))
                else:
                    span = (b1 - b0 + 1) << shift
                    nshift = max(0, int(np.ceil(np.log2(span))) - 13)
                    stack.append((klo + (b0 << shift), nshift))
            if collect:
                if stats is not None:
                    stats['passes'] = stats.get('passes', 0) + 1
                for b0, b1 in collect:
                    lo_k, hi_k = klo + (b0 << shift), klo + ((b1 + 1) << shift)
                    sel = (key >= lo_k) & (key < hi_k)
                    sk = np.sort(key[sel])
                    if stats is not None:
                        stats['collected'] = stats.get('collected', 0) + sk.size
                    ab = key[key >= hi_k]
                    nk = int(ab.min()) if ab.size else kmax
                    evaluate_sorted(sk, int((key < lo_k).sum()), float(vals[key < lo_k].sum()), nk)
    if ternary:
        mean = s_tot / n
        if float(val_of(kmin)) > 0.5 * mean:                  # optimal.py:95-116
            c = F32(0.5 * mean)
            # all elements are above c: k = 0
            best.offer(closed_cost2(c, 0, 0.0, n, s_tot, q_tot, True), n, c)
    if best.ncand == 0:
        return F32(0), best
    return best.val, best


def solve(rows, ternary, skip=1, stats=None):
    rows = np.asarray(rows, dtype=F32)
    out = np.zeros(rows.shape[0], dtype=F32)
    infos = []
    for r in range(rows.shape[0]):
        out[r], b = solve_row(np.abs(rows[r, ::skip]), ternary, stats)
        infos.append(b)
    return out, infos
