"""Time lsq_pwconv_fwd on the three downsampling shortcuts of ResNet-18 at batch 512 (development)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ml_quant_b200 import ops
DEV = torch.device('cuda:0')
torch.manual_seed(0)
for cin, cout, hw in [(64, 128, 56), (128, 256, 28), (256, 512, 14)]:
    x = torch.randn(512, cin, hw, hw, device=DEV)
    wt = torch.randn(cout, cin, device=DEV) * (1.0 / cin) ** 0.5
    b = torch.randn(cout, device=DEV)
    img = ops.pwconv_pack(wt)
    for _ in range(3): y = ops.pwconv_fwd(x, img, b, cout, 2)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): y = ops.pwconv_fwd(x, img, b, cout, 2)
    e1.record(); torch.cuda.synchronize()
    want = torch.nn.functional.conv2d(x[:8], wt.view(cout, cin, 1, 1), b, 2)
    err = float((y[:8] - want).abs().max() / want.abs().max())
    print('pwconv %d->%d @%d: %.1f us  err %.2e' % (cin, cout, hw, e0.elapsed_time(e1) * 100, err))
