import sys, torch
sys.path.insert(0, '/root/repo')
from ml_quant_b200 import ops
x = torch.randn(512, 512, 7, 7, device='cuda')
for f, name in ((lambda: ops.plane_mean(x), 'plane_mean'), (lambda: x.mean((2, 3)), 'aten mean')):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): f()
    e1.record(); torch.cuda.synchronize()
    print(name, '%.1f us' % (e0.elapsed_time(e1) * 50))
