"""One launch of the fused activation quantizer on a benchmark layer shape (for ncu captures)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ml_quant_b200 import ops  # noqa: E402

c, h, w = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (256, 14, 14)))
n = int(sys.argv[4]) if len(sys.argv) > 4 else 512
dev = torch.device('cuda:0')
torch.manual_seed(0)
x = torch.randn(n, c, h, w, device=dev)
pro = (torch.rand(c, device=dev) + 0.5, torch.randn(c, device=dev) * 0.2, h * w)
g = ops.act_geometry(n, c, h, w, 3, 3, 1, 1)
for _ in range(3):
    planes, tab, dg = ops.quantize_act(x, g, False, 3.0, 3, None, pro, diag=True)
torch.cuda.synchronize()
print('groups flagged per row (mean/max):', float(dg[:, 7].float().mean()), int(dg[:, 7].max()),
      ' flagged bins:', float(dg[:, 1].float().mean()), ' status != 0:', int((dg[:, 0] != 0).sum()))
