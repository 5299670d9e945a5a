"""Time one fused QuantConv2d layer's binary convolution (development): LSQ_C, LSQ_HW, LSQ_COUT, LSQ_STRIDE."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn
from ml_quant_b200 import ops
from ml_quant_b200.binary.binary_conv import QuantConv2d
c, hw, n = int(os.environ.get('LSQ_C', '64')), int(os.environ.get('LSQ_HW', '56')), int(os.environ.get('LSQ_N', '512'))
co, st = int(os.environ.get('LSQ_COUT', str(c))), int(os.environ.get('LSQ_STRIDE', '1'))
dev = torch.device('cuda:0')
torch.manual_seed(0)
conv = QuantConv2d('ls-2', 'ls-1', c, co, 3, {'kind': 'symmetric', 'alpha': 3.0}, stride=st, padding=1).to(dev).eval()
conv.w_approximate.v1.copy_(conv.weight.detach().abs().mean(dim=(1, 2, 3)))
bn = nn.BatchNorm2d(c).to(dev).eval()
x = torch.randn(n, c, hw, hw, device=dev)
res = torch.randn(n, co, (hw - 1) // st + 1, (hw - 1) // st + 1, device=dev)
with torch.no_grad():
    for _ in range(3):
        y = conv.forward_fused(x, bn, nn.ReLU(), res, True)
    torch.cuda.synchronize()
    ops.PROFILE = []
    for _ in range(10):
        y = conv.forward_fused(x, bn, nn.ReLU(), res, True)
    torch.cuda.synchronize()
agg = {}
for name, e0, e1, b, o in ops.PROFILE:
    agg[name] = agg.get(name, 0.0) + e0.elapsed_time(e1) / 10
print('dbg', os.environ.get('LSQ_BCONV_DBG'), 'c', c, 'hw', hw, 'cout', co, 'stride', st, {k: round(v * 1000, 1) for k, v in agg.items()}, 'us')
