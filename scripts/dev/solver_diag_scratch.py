import os, sys
sys.path.insert(0, '/root/repo')
exec(open('/root/repo/scripts/solver_diag.py').read().split("for rows, length in")[0])
for scratch in (False, True):
    ops.SOLVE_SCRATCH = scratch
    print('scratch', scratch)
    for rows, length in [(512, 200704), (512, 100352)]:
        torch.manual_seed(0)
        x = torch.randn(rows, length, device=dev)
        for _ in range(2):
            v1, dg = ops.solve_v1(x, False, 3, 3.0, diag=True)
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ops.solve_v1(x, False, 3, 3.0); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        d = dg.float().mean(0).tolist()
        tot = sum(d[4:15])
        print(f'rows {rows} len {length}: {best*1e3:.0f} us passes {d[0]:.2f} collected {d[1]:.0f} total cycles/row {tot:.0f}')
        print('   ' + '  '.join(f'{n} {c:.0f}' for n, c in zip(names, d[4:15])))
