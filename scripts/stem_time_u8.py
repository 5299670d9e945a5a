"""Time lsq_stem_fwd against lsq_stem_fwd_u8 at batch 512 (development)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ml_quant_b200 import ops, runtime
DEV = torch.device('cuda:0')
torch.manual_seed(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
u8 = torch.randint(0, 256, (n, 3, 224, 224), dtype=torch.uint8, device=DEV)
lut = runtime.pixel_lut((0.485, 0.456, 0.406), (0.229, 0.224, 0.225)).to(DEV)
x = ops.u8_expand(u8, lut)
wt = torch.randn(64, 3, 7, 7, device=DEV) * 0.1
b = torch.randn(64, device=DEV)
img = ops.stem_pack(wt)
for name, fn in [('fp32', lambda: ops.stem_fwd(x, img, b)), ('uint8', lambda: ops.stem_fwd_u8(u8, lut, img, b)),
                 ('expand', lambda: ops.u8_expand(u8, lut))]:
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    print('stem b%d %s: %.3f ms' % (n, name, e0.elapsed_time(e1) / 10))
