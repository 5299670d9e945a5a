"""Phase breakdown of solve_v1 (cycle counters of thread 0, lsq_solve.cu LSQ_TICK) on ImageNet ResNet-18 row shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ml_quant_b200 import ops
dev = torch.device('cuda:0')
names = ['pop+zero', 'hist(glob)', 'hist(list)', 'reduce', 'scan+flag', 'thread0', 'collect(glob)', 'collect(list)', 'coll.reduce', 'sort+eval', 'tail']
for rows, length in [(512, 200704), (512, 100352), (512, 50176), (512, 25088)]:
    torch.manual_seed(0)
    x = torch.randn(rows, length, device=dev)
    for _ in range(2):
        v1, dg = ops.solve_v1(x, False, 3, 3.0, diag=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.solve_v1(x, False, 3, 3.0); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    d = dg.float().mean(0).tolist()
    tot = sum(d[4:15])
    print(f'rows {rows} len {length}: {ms*1e3:.0f} us  ({4.0*rows*length/ms/1e6:.0f} GB/s alg)  passes {d[0]:.2f} collected {d[1]:.0f} ncand {d[2]:.1f} flags {d[3]:.2f}  total cycles/row {tot:.0f}')
    print('   ' + '  '.join(f'{n} {c:.0f}' for n, c in zip(names, d[4:15])))
