"""Role wait-cycle diagnostics of the stem kernel (needs a library built with LSQ_NVCC_EXTRA=-DLSQ_TC_DIAG)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ml_quant_b200 import ops
DEV = torch.device('cuda:0')
torch.manual_seed(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
x = torch.randn(n, 3, 224, 224, device=DEV)
wt = torch.randn(64, 3, 7, 7, device=DEV) * 0.1
b = torch.randn(64, device=DEV)
img = ops.stem_pack(wt)
ops.stem_fwd(x, img, b)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
os.environ['QUIET'] = '1'
