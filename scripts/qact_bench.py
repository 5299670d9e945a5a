"""Time the fused activation quantizer (lsq_quantize_act) against the round-1 kernel pair (lsq_solve_v1_ex +
lsq_encode_act_ex) on the QuantConv2d input shapes of the benchmark networks.  CUDA events, inputs larger than L2.

    python scripts/qact_bench.py [--batch 512] [--only N]      (LSQ_QACT_CS=1|2|4|8 forces a cluster size)
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ml_quant_b200 import ops  # noqa: E402

SHAPES = [('imagenet s1 64x56x56', 64, 56, 56, 1, 3.0), ('imagenet s1->s2 64x56x56 /2', 64, 56, 56, 2, 3.0),
          ('imagenet s2 128x28x28', 128, 28, 28, 1, 3.0), ('imagenet s3 256x14x14', 256, 14, 14, 1, 3.0),
          ('imagenet s4 512x7x7', 512, 7, 7, 1, 3.0), ('cifar s1 64x32x32 (b256)', 64, 32, 32, 1, 2.0),
          ('cifar s2 128x16x16 (b256)', 128, 16, 16, 1, 2.0), ('cifar s3 256x8x8 (b256)', 256, 8, 8, 1, 2.0),
          ('cifar s4 512x4x4 (b256)', 512, 4, 4, 1, 2.0)]


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return sorted(ts)[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=512)
    ap.add_argument('--only', type=int, default=-1)
    args = ap.parse_args()
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
    for i, (name, c, h, w, st, alpha) in enumerate(SHAPES):
        if args.only >= 0 and i != args.only:
            continue
        n = 256 if 'cifar' in name else args.batch
        x = torch.randn(n, c, h, w, device=dev)
        a = torch.rand(c, device=dev) + 0.5
        b = torch.randn(c, device=dev) * 0.2
        pro = (a, b, h * w)
        g = ops.act_geometry(n, c, h, w, 3, 3, st, 1)
        rows = x.reshape(n, -1)
        tab = torch.empty(2, n, device=dev)

        def old():
            flush.zero_() if x.numel() * 4 < 200e6 else None
            ops.solve_v1(rows, False, 3, alpha, prologue=pro, out=tab[0])
            ops.encode_act(x, g, tab[:1], 2, alpha, True, None, pro, next_scale_out=tab[1])

        def new():
            flush.zero_() if x.numel() * 4 < 200e6 else None
            ops.quantize_act(x, g, False, alpha, 3, None, pro)

        def fl():
            flush.zero_() if x.numel() * 4 < 200e6 else None
        t_fl = timed(fl)
        t_old, t_new = timed(old) - t_fl, timed(new) - t_fl
        _, t2, dg = ops.quantize_act(x, g, False, alpha, 3, None, pro, diag=True)
        torch.cuda.synchronize()
        gb = x.numel() * 4 / 1e9
        print(f'{name:32s} n={n:4d}  old {t_old:8.1f} us ({gb / t_old * 1e6:6.0f} GB/s)   fused {t_new:8.1f} us '
              f'({gb / t_new * 1e6:6.0f} GB/s)  cs={int(dg[0, 5])} odd_rows={int((dg[:, 0] != 0).sum())} '
              f'collected~{float(dg[:, 2].float().mean()):.0f} flagged~{float(dg[:, 1].float().mean()):.1f}', flush=True)
        ph = dg[:, 8:15].float().mean(0).tolist()
        print('      mean cycles/row (rank 0): sweep1 %.0f  wait1 %.0f  merge %.0f  scan %.0f  collect %.0f  sort+eval %.0f  '
              'sweep2 %.0f  | total %.0f, SMs used %d' % (*ph, sum(ph), len(set(dg[:, 15].tolist()))), flush=True)


if __name__ == '__main__':
    main()
