import sys, torch, time
sys.path.insert(0, '/root/repo')
import torch.nn.functional as F
from ml_quant_b200 import ops
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.enabled = False
DEV = torch.device('cuda:0')
torch.manual_seed(9)
for n, h, w in [(3, 224, 224), (2, 64, 64), (2, 61, 75), (1, 32, 40), (2, 250, 250), (1, 260, 300), (5, 7, 7), (300, 96, 96)]:
    x = torch.randn(n, 3, h, w, device=DEV)
    wt = torch.randn(64, 3, 7, 7, device=DEV) * 0.1
    wt[5] *= 1e-3; wt[6] *= 300.0
    b = torch.randn(64, device=DEV)
    want = F.relu(F.max_pool2d(F.conv2d(x, wt, b, 2, 3), 3, 2, 1))
    got = ops.stem_fwd(x, ops.stem_pack(wt), b)
    torch.cuda.synchronize()
    err = float((got - want).abs().max() / want.abs().max())
    errc = float(((got - want).abs().amax(dim=(0, 2, 3)) / want.abs().amax(dim=(0, 2, 3)).clamp_min(1e-30)).max())
    print(n, h, w, 'fused' if ops._C.lib().lsq_stem_is_fused(n, h, w) else 'two-kernel', 'err', err, 'per-channel err', errc, flush=True)
x = torch.randn(512, 3, 224, 224, device=DEV)
img = ops.stem_pack(wt)
for _ in range(3): ops.stem_fwd(x, img, b)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): ops.stem_fwd(x, img, b)
e1.record(); torch.cuda.synchronize()
print('stem b512: %.3f ms' % (e0.elapsed_time(e1) / 10))
