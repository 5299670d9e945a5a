"""CPU estimate for DESIGN.md section 10 item 3: accuracy of a stem convolution whose operands are split into fp16
hi/lo pairs (tcgen05 kind::f16, fp32 accumulation; three product terms hi*hi + hi*lo + lo*hi) against the 3xTF32
split the stem kernel uses now and against plain fp32, all measured against an fp64 convolution.
Runs on the CPU (torch), no GPU needed:  python scripts/dev/stem_f16_split_error.py"""
import torch
import torch.nn.functional as F

torch.manual_seed(0)


def tf32(t):          # round-to-nearest-even to 10 mantissa bits
    b = t.float().view(torch.int32)
    b = (b + 0x0FFF + ((b >> 13) & 1)) & ~0x1FFF
    return b.view(torch.float32)


def conv(x, w):       # products exact in fp64, the sum rounded once: an upper bound for fp32 accumulation quality
    return F.conv2d(x.double(), w.double(), stride=2, padding=3)


for name, xs in (('N(0,1) pixels', 1.0), ('pixels in [0, 255]', None)):
    x = torch.randn(2, 3, 96, 96) if xs else torch.rand(2, 3, 96, 96) * 255.0
    w = torch.randn(64, 3, 7, 7) * (2.0 / 147) ** 0.5 * (torch.rand(64, 1, 1, 1) + 0.5)      # BatchNorm-folded magnitudes
    ref = conv(x, w)
    scale = ref.abs().max()
    # fp16 hi/lo
    xh = x.half(); xl = (x - xh.float()).half()
    wh = w.half(); wl = (w - wh.float()).half()
    y16 = conv(xh, wh) + conv(xh, wl) + conv(xl, wh)
    y16_4 = y16 + conv(xl, wl)
    # 3xTF32
    xt = tf32(x); xtl = tf32(x - xt)
    wt = tf32(w); wtl = tf32(w - wt)
    yt = conv(xt, wt) + conv(xt, wtl) + conv(xtl, wt)
    y32 = F.conv2d(x, w, stride=2, padding=3).double()
    y16_1 = conv(xh, wh)
    print(f'{name}: max|err| / max|y|   fp32 conv {float((y32 - ref).abs().max() / scale):.2e}   3xTF32 {float((yt - ref).abs().max() / scale):.2e}   '
          f'fp16 hi/lo, 3 terms {float((y16 - ref).abs().max() / scale):.2e}   4 terms {float((y16_4 - ref).abs().max() / scale):.2e}   '
          f'single fp16 {float((y16_1 - ref).abs().max() / scale):.2e}   (fp16 lo parts below the normal range: x {float((xl.float().abs() < 6.1e-5).float().mean()):.2f}, w {float((wl.float().abs() < 6.1e-5).float().mean()):.2f})')
