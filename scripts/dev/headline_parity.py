"""Development aid: full-size headline network vs the oracle (plain and fused), error and top-1 agreement."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from oracle import lsq_oracle as O
from ml_quant_b200 import configs, runtime
runtime.strict_fp32()
cfg = 'imagenet_resnet18_ls1w_ls2a'
dev = torch.device('cuda:0')
for seed in (0, 1):
    model = runtime.build_model(cfg, dev, seed=seed)
    runtime.calibrate(model, (3, 224, 224), batches=1, batch=8)
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    g = torch.Generator().manual_seed(1234 + seed)
    x = torch.randn(8, 3, 224, 224, generator=g)
    y_ref = O.resnet_forward(sd, configs.arch(cfg), x)
    with torch.no_grad():
        y = model(x.to(dev)).cpu()
        yf = runtime.optimize_for_inference(model)(x.to(dev)).cpu()
    for name, out in (('plain', y), ('fused', yf)):
        e = (out - y_ref).abs().flatten(1).max(1).values / y_ref.abs().max()
        print(seed, name, 'per-sample err', [round(float(v), 4) for v in e], 'top1 equal', (out.argmax(1) == y_ref.argmax(1)).tolist(),
              'logit range', float(y_ref.min()), float(y_ref.max()))
