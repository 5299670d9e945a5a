"""Every kernel of the library once at small sizes, for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool memcheck python scripts/sanitize_small.py
"""
import os
import sys

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ml_quant_b200 import ops, runtime  # noqa: E402
from ml_quant_b200.binary.binary_conv import QuantConv2d, QuantLinear  # noqa: E402

DEV = torch.device('cuda:0')
torch.manual_seed(0)
with torch.no_grad():
    # stem: one-kernel route (fp32 and uint8 pixels), two-kernel route
    wt = torch.randn(64, 3, 7, 7, device=DEV) * 0.1
    b = torch.randn(64, device=DEV)
    img = ops.stem_pack(wt)
    for n, h, w in [(2, 64, 64), (1, 33, 75), (1, 40, 300)]:
        ops.stem_fwd(torch.randn(n, 3, h, w, device=DEV), img, b)
    lut = runtime.pixel_lut((0.485, 0.456, 0.406), (0.229, 0.224, 0.225))
    u8 = torch.randint(0, 256, (2, 3, 61, 75), dtype=torch.uint8, device=DEV)
    ops.stem_fwd_u8(u8, lut, img, b)
    ops.u8_expand(u8, lut)
    # shortcut convolution
    x = torch.randn(2, 64, 28, 28, device=DEV)
    ops.pwconv_fwd(x, ops.pwconv_pack(torch.randn(128, 64, device=DEV)), torch.randn(128, device=DEV), 128, 2)
    # fused blocks: ls-2 / ls-T / ls-1 activations, stride 1 and 2, 64 / 128 / 256 channels, residual + ReLU
    for a_q, cin, cout, hw, st in [('ls-2', 64, 64, 28, 1), ('ls-T', 64, 128, 28, 2), ('ls-1', 128, 128, 14, 1),
                                   ('ls-2', 128, 256, 14, 2), ('ls-2', 256, 256, 8, 1)]:
        conv = QuantConv2d(a_q, 'ls-1', cin, cout, 3, {'kind': 'symmetric', 'alpha': 2.0}, stride=st, padding=1).to(DEV).eval()
        conv.w_approximate.v1.copy_(conv.weight.abs().mean(dim=(1, 2, 3)))
        bn = nn.BatchNorm2d(cin).to(DEV).eval()
        x = torch.randn(3, cin, hw, hw, device=DEV)
        x[1] = 0.5                                    # a row the fused quantizer hands to the fallback kernel
        ho = (hw - 1) // st + 1
        conv.forward_fused(x, bn, nn.ReLU(), torch.randn(3, cout, ho, ho, device=DEV), True)
        conv(x)
    # multi-plane weights, QuantLinear, generic solver on long and short rows, multi-tensor solves
    conv = QuantConv2d('ls-2', 'ls-2', 64, 64, 3, {'kind': 'symmetric', 'alpha': 2.0}, padding=1).to(DEV)
    conv.train()
    conv(torch.randn(2, 64, 16, 16, device=DEV))
    conv.eval()
    conv(torch.randn(2, 64, 16, 16, device=DEV))
    lin = QuantLinear('ls-2', 'ls-1', 256, 128, {'kind': 'symmetric', 'alpha': 2.0}).to(DEV).eval()
    lin.w_approximate.v1.copy_(lin.weight.abs().mean(dim=1))
    lin(torch.randn(70, 256, device=DEV))
    for rows, ln in [(5, 70001), (9, 30000), (40, 4608), (33, 576), (7, 64), (3, 2)]:
        xr = torch.randn(rows, ln, device=DEV)
        for tern in (False, True):
            ops.solve_v1(xr, tern, 3, 2.0)
            ops.solve_v1(xr, tern, 1, None)
        ops.row_absmean(xr)
    ws = [torch.randn(r, ln, device=DEV) for r, ln in [(64, 576), (128, 1152), (16, 27), (256, 2304)]]
    ops.solve_v1_multi(ws, False, 3)
    ops.solve_v1_multi(ws, True, 3)
    ops.row_absmean_multi(ws)
    ops.plane_mean(torch.randn(3, 70, 7, 7, device=DEV))
torch.cuda.synchronize()
print('sanitize_small: all kernels ran')
