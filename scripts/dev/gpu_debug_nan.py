"""Debug helper: find the first QuantConv2d producing NaN in the plain (unfused) CIFAR network and dump why."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from ml_quant_b200 import runtime, ops
from ml_quant_b200.binary.binary_conv import QuantConv2d
dev = torch.device('cuda:0')
runtime.strict_fp32()
cfg, shape = 'cifar100_resnet18_ls1w_ls2a', (3, 32, 32)
torch.manual_seed(7)
model = runtime.build_model(cfg, dev)
runtime.calibrate(model, shape, batches=2, batch=32)
x = torch.randn(8, *shape, device=dev)
bad = []
def hook(name):
    def f(m, inp, out):
        xi = inp[0]
        ni, no = bool(torch.isnan(xi).any()), bool(torch.isnan(out).any())
        if no and not ni and not bad:
            bad.append((name, m, xi.clone()))
        print(name, 'in nan', ni, 'out nan', no, 'w.v1 nan', bool(torch.isnan(m.w_approximate.v1).any()), flush=True)
    return f
for n, m in model.named_modules():
    if isinstance(m, QuantConv2d):
        m.register_forward_hook(hook(n))
with torch.no_grad():
    model(x)
    if bad:
        name, m, xi = bad[0]
        print('first bad', name, tuple(xi.shape), m.x_quant, m.clamp_alpha, m.stride, m.padding)
        g = m._packed_geometry(xi)
        rows = xi.reshape(xi.shape[0], -1)
        v1, dg = ops.solve_v1(rows, False, 3, m.clamp_alpha, diag=True)
        print('v1', v1.tolist()); print('diag', dg[:, :4].tolist())
        planes, v2 = ops.encode_act(xi, g, [v1], 2, m.clamp_alpha, True)
        print('v2', v2.tolist())
        tab = torch.stack([v1, v2])
        for impl in (1, 2):
            y = ops.bconv2d(planes, g, 2, tab, m.packed_weights(), m.w_approximate.v1, m.bias, m.out_channels, impl)
            nn_ = torch.isnan(y)
            print('impl', impl, 'nan count', int(nn_.sum()), 'per sample', nn_.flatten(1).sum(1).tolist())
        for i in range(5):
            y = m(xi)
            print('repeat', i, int(torch.isnan(y).sum()))
