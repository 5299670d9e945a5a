import sys; import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from oracle import lsq_oracle as O
from ml_quant_b200 import configs, runtime
cfg='imagenet_resnet18_ls1w_ls2a'
model = runtime.build_model(cfg, None, seed=0)
# calibrate on CPU impossible (no CPU path) -> emulate: weight scales from weights, BN stats default
sd = {k: v.clone() for k, v in model.state_dict().items()}
for k in list(sd):
    if k.endswith('w_approximate.v1'):
        w = sd[k[:-len('w_approximate.v1')] + 'weight']; sd[k] = w.abs().mean(dim=(1,2,3))
g = torch.Generator().manual_seed(1234)
x = torch.randn(4,3,224,224, generator=g)
y32 = O.resnet_forward(sd, configs.arch(cfg), x)
# perturb the input by one ulp-scale relative noise: how far do the logits move?
x2 = x * (1 + 1e-7 * torch.randn_like(x))
y32b = O.resnet_forward(sd, configs.arch(cfg), x2)
e = (y32b - y32).abs().flatten(1).max(1).values / y32.abs().max()
print('logit range', float(y32.min()), float(y32.max()))
print('rel change of logits for a 1e-7 relative input perturbation (oracle vs oracle):', [round(float(v),4) for v in e], 'top1 equal', (y32.argmax(1)==y32b.argmax(1)).tolist())
