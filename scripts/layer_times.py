"""Per-launch CUDA-event times of our kernels over one eager forward of the benchmark workload."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ml_quant_b200 import ops, runtime
cfg = os.environ.get('LSQ_CFG', 'imagenet_resnet18_ls1w_ls2a')
B = int(os.environ.get('LSQ_BATCH', '512'))
HW = int(os.environ.get('LSQ_HW', '224'))
dev = torch.device('cuda:0')
torch.backends.cudnn.benchmark = True
model = runtime.build_model(cfg, dev)
runtime.calibrate(model, (3, HW, HW))
runtime.optimize_for_inference(model)
x = torch.randn(B, 3, HW, HW, device=dev)
with torch.no_grad():
    for _ in range(3):
        model(x)
    torch.cuda.synchronize()
    ops.PROFILE = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); model(x); e1.record()
    torch.cuda.synchronize()
prof, ops.PROFILE = ops.PROFILE, None
tot = {}
for name, a, b, nbytes, nops in prof:
    ms = a.elapsed_time(b)
    tot[name] = tot.get(name, 0.0) + ms
    extra = f'{nops / ms / 1e9:8.0f} TOP/s' if nops else ''
    print(f'{name:14s} {ms * 1e3:8.1f} us  {nbytes / 1e6:8.1f} MB  {nbytes / ms / 1e6:7.0f} GB/s {extra}')
fwd = runtime.GraphedForward(model, x)
for _ in range(3):
    fwd()
g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
g0.record()
for _ in range(10):
    fwd()
g1.record(); torch.cuda.synchronize()
print(cfg, 'batch', B, 'graph step ms', g0.elapsed_time(g1) / 10, 'images/s', B * 10 / g0.elapsed_time(g1) * 1e3)
print('eager step ms', e0.elapsed_time(e1), {k: round(v, 3) for k, v in tot.items()}, 'sum', round(sum(tot.values()), 3))
