import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from ml_quant_b200 import runtime
dev = torch.device('cuda:0')
runtime.strict_fp32()
for cfg, shape in [('cifar100_resnet18_ls1w_ls2a', (3, 32, 32)), ('imagenet_resnet18_ls1w_ls1a', (3, 64, 64))]:
    model = runtime.build_model(cfg, dev)
    runtime.calibrate(model, shape, batches=1, batch=16)
    print(cfg, 'bn stats finite', all(bool(torch.isfinite(b).all()) for b in model.buffers()))
    x = torch.randn(8, *shape, device=dev)
    with torch.no_grad():
        h = x
        for i, blk in enumerate(model.blocks):
            h = blk(h)
            print('  plain block', i, 'nan', bool(torch.isnan(h).any()), 'inf', bool(torch.isinf(h).any()), float(h.abs().max()), flush=True)
        plain = model(x)
        fused = runtime.optimize_for_inference(model)(x)
    print(' plain nan', bool(torch.isnan(plain).any()), 'fused nan', bool(torch.isnan(fused).any()), float((plain - fused).abs().max()), float(plain.abs().max()))
