import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
dev = 'cuda:0'
torch.backends.cudnn.benchmark = True
x = torch.randn(512, 3, 224, 224, device=dev)
w = torch.randn(64, 3, 7, 7, device=dev) * 0.05
b = torch.randn(64, device=dev)
def timeit(name, fn, n=5):
    for _ in range(3): y = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): y = fn()
    e1.record(); torch.cuda.synchronize()
    print(f'{name}: {e0.elapsed_time(e1)/n:.3f} ms', tuple(y.shape), y.is_contiguous(), flush=True)
    return y
with torch.no_grad():
    y0 = timeit('conv nchw fp32(tf32 default)', lambda: F.conv2d(x, w, b, 2, 3))
    timeit('maxpool nchw', lambda: F.max_pool2d(y0, 3, 2, 1))
    p0 = F.max_pool2d(y0, 3, 2, 1)
    timeit('relu_ pooled', lambda: F.relu_(p0))
    xcl = x.contiguous(memory_format=torch.channels_last); wcl = w.contiguous(memory_format=torch.channels_last)
    timeit('to channels_last (input)', lambda: x.contiguous(memory_format=torch.channels_last))
    y1 = timeit('conv channels_last', lambda: F.conv2d(xcl, wcl, b, 2, 3))
    timeit('maxpool channels_last', lambda: F.max_pool2d(y1, 3, 2, 1))
    p1 = F.max_pool2d(y1, 3, 2, 1)
    timeit('pooled cl -> nchw contiguous', lambda: p1.contiguous())
    xb = xcl.bfloat16(); wb = wcl.bfloat16()
    y2 = timeit('conv channels_last bf16', lambda: F.conv2d(xb, wb, b.bfloat16(), 2, 3))
    timeit('whole stem nchw', lambda: F.relu_(F.max_pool2d(F.conv2d(x, w, b, 2, 3), 3, 2, 1)))
    timeit('whole stem cl', lambda: F.relu_(F.max_pool2d(F.conv2d(x.contiguous(memory_format=torch.channels_last), wcl, b, 2, 3), 3, 2, 1)).contiguous())
    print('cl vs nchw max diff', float((y1 - y0).abs().max()), float(y0.abs().max()))
