import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from ml_quant_b200 import ops
x = torch.randn(512, 3, 224, 224, device='cuda:0')
w = F.pad((torch.randn(64, 3, 7, 7, device='cuda:0') * 0.1).reshape(64, 147), (0, 5)).contiguous()
b = torch.randn(64, device='cuda:0')
for _ in range(3):
    ops.stem_fwd(x, w, b)
torch.cuda.synchronize()
