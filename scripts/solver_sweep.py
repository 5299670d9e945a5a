"""BASELINE.json configs[4]: least-squares solver throughput over the 53 conv weights of a ResNet-50
(23.45 M elements, 26 560 rows, row length 64..4608), k in {ls-1, ls-2, ls-T}, skip 3 and 1.
Prints Melem/s and GB/s on 4 B per element (one algorithmic read), CUDA events, L2 flushed between runs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ml_quant_b200 import ops
SHAPES = {(64, 3, 7, 7): 1, (64, 64, 1, 1): 1, (64, 64, 3, 3): 3, (64, 256, 1, 1): 2, (128, 128, 3, 3): 4, (128, 256, 1, 1): 1,
          (128, 512, 1, 1): 3, (256, 64, 1, 1): 4, (256, 256, 3, 3): 6, (256, 512, 1, 1): 1, (256, 1024, 1, 1): 5,
          (512, 128, 1, 1): 4, (512, 256, 1, 1): 1, (512, 512, 3, 3): 3, (512, 1024, 1, 1): 1, (512, 2048, 1, 1): 2,
          (1024, 256, 1, 1): 6, (1024, 512, 1, 1): 1, (2048, 512, 1, 1): 3, (2048, 1024, 1, 1): 1}
dev = torch.device('cuda:0')
torch.manual_seed(0)
ws = [torch.randn(s[0], s[1] * s[2] * s[3], device=dev) * 0.05 for s, c in SHAPES.items() for _ in range(c)]
elems = sum(w.numel() for w in ws)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timed(fn, reps=5):
    best = 1e9
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for w in ws:
            fn(w)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
print(f'{len(ws)} tensors, {sum(w.shape[0] for w in ws)} rows, {elems / 1e6:.2f} M elements')
for name, fn in [('ls-1 (row_absmean)', lambda w: ops.row_absmean(w)),
                 ('ls-2 skip 3', lambda w: ops.solve_v1(w, False, 3)), ('ls-2 skip 1', lambda w: ops.solve_v1(w, False, 1)),
                 ('ls-T skip 3', lambda w: ops.solve_v1(w, True, 3)), ('ls-T skip 1', lambda w: ops.solve_v1(w, True, 1))]:
    ms = timed(fn)
    # the same 53 launches replayed as one CUDA graph (no host time between them)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for w in ws:
            fn(w)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        keep = [fn(w) for w in ws]
    best = 1e9
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); graph.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f'{name:20s} eager {ms:7.3f} ms {elems / ms / 1e3:8.0f} Melem/s | graph {best:7.3f} ms {elems / best / 1e3:8.0f} Melem/s '
          f'{4.0 * elems / best / 1e6:7.0f} GB/s (53 launches)')
# all 53 tensors in ONE launch (lsq_solve_v1_multi / lsq_row_absmean_multi; skip 1 keeps own launches for the
# tensors whose rows exceed the small-row kernel)
for name, fn in [('ls-1 multi', lambda: ops.row_absmean_multi(ws)),
                 ('ls-2 skip 3 multi', lambda: ops.solve_v1_multi(ws, False, 3)), ('ls-2 skip 1 multi', lambda: ops.solve_v1_multi(ws, False, 1)),
                 ('ls-T skip 3 multi', lambda: ops.solve_v1_multi(ws, True, 3)), ('ls-T skip 1 multi', lambda: ops.solve_v1_multi(ws, True, 1))]:
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); keep = fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        keep = fn()
    gbest = 1e9
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); graph.replay(); e1.record(); torch.cuda.synchronize()
        gbest = min(gbest, e0.elapsed_time(e1))
    print(f'{name:20s} eager {best:7.3f} ms {elems / best / 1e3:8.0f} Melem/s | graph {gbest:7.3f} ms {elems / gbest / 1e3:8.0f} Melem/s '
          f'{4.0 * elems / gbest / 1e6:7.0f} GB/s')
if '--cpu' in sys.argv:
    import time
    from oracle import lsq_oracle as O          # checker used as the CPU baseline of this sweep (test infrastructure)
    torch.set_num_threads(os.cpu_count() or 1)
    cw = [w.cpu() for w in ws]
    for name, tern in (('ls-2 skip 3', False), ('ls-T skip 3', True)):
        t = time.time()
        for w in cw:
            O.solve_v1(w, tern, 3, chunk=64)
        dt = time.time() - t
        print(f'CPU oracle {name:12s} {dt * 1e3:8.1f} ms  {elems / dt / 1e6:9.1f} Melem/s  ({os.cpu_count()} threads)')
