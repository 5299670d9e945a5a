"""One fused QuantConv2d layer (ls-1 weights / ls-2 activations, BN prologue, ReLU + residual epilogue) of the
ImageNet ResNet-18 at batch 512, for Nsight Compute:  LSQ_C=128 LSQ_HW=28 python profiles/profile_bconv.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn
from ml_quant_b200.binary.binary_conv import QuantConv2d
c, hw, n = int(os.environ.get('LSQ_C', '128')), int(os.environ.get('LSQ_HW', '28')), int(os.environ.get('LSQ_N', '512'))
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
print('ok', tuple(y.shape))
