"""Fused stem (tcgen05 conv + pool) on a 512 x 3 x 224 x 224 batch: timing and a check against fp32 ATen."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from ml_quant_b200 import ops, runtime
runtime.strict_fp32()
torch.manual_seed(0)
n = int(os.environ.get('LSQ_N', '512'))
x = torch.randn(n, 3, 224, 224, device='cuda:0')
w = torch.randn(64, 3, 7, 7, device='cuda:0') * 0.1
b = torch.randn(64, device='cuda:0')
img = ops.stem_pack(w)
for _ in range(3):
    y = ops.stem_fwd(x, img, b)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    y = ops.stem_fwd(x, img, b)
e1.record(); torch.cuda.synchronize()
print('stem ms', e0.elapsed_time(e1) / 5)
want = F.relu(F.max_pool2d(F.conv2d(x[:8], w, b, 2, 3), 3, 2, 1))
print('rel err', float((y[:8] - want).abs().max() / want.abs().max()))
