"""Time lsq_stem_fwd at batch 512 (development)."""
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
for _ in range(3): ops.stem_fwd(x, img, b)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): ops.stem_fwd(x, img, b)
e1.record(); torch.cuda.synchronize()
print('stem b%d dbg=%s: %.3f ms' % (n, os.environ.get('LSQ_STEM_DBG'), e0.elapsed_time(e1) / 10))
