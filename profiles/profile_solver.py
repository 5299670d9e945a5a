"""solve_v1 on one ImageNet layer-1 sized activation tensor (512 rows x 200704), for Nsight Compute."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ml_quant_b200 import ops
rows, length = int(os.environ.get('LSQ_ROWS', '512')), int(os.environ.get('LSQ_LEN', '200704'))
torch.manual_seed(0)
x = torch.randn(rows, length, device='cuda:0')
a = torch.rand(64, device='cuda:0') + 0.5
b = torch.randn(64, device='cuda:0') * 0.1
for _ in range(3):
    ops.solve_v1(x, False, 3, 3.0, prologue=(a, b, length // 64))
torch.cuda.synchronize()
