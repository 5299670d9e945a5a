"""One markdown table row per profiled launch of one or more .ncu-rep files (key raw metrics):
    python scripts/ncu_table.py out.md rep1.ncu-rep [rep2 ...]
Run on the GPU box right after the capture so that only the small table has to travel back."""
import csv
import os
import subprocess
import sys

COLS = [('gpu__time_duration.sum', 'us'), ('dram__bytes_read.sum', 'rd MB'), ('dram__bytes_write.sum', 'wr MB'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram %'),
        ('sm__ops_path_tensor_op_utcimma_src_int8_sparsity_off.avg.pct_of_peak_sustained_elapsed', 'utcimma %'),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'tensor %'),
        ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue %'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps %'),
        ('lts__t_sector_hit_rate.pct', 'L2 hit %'), ('launch__grid_size', 'grid'), ('launch__block_size', 'block'),
        ('launch__cluster_size', 'cluster'), ('launch__registers_per_thread', 'regs'),
        ('launch__shared_mem_per_block_dynamic', 'smem')]
SCALE = {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3, 'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'usecond': 1.0, 'msecond': 1e3, 'nsecond': 1e-3}
out, reps = sys.argv[1], sys.argv[2:]
lines = ['# ncu --set full, one row per profiled launch (`--clock-control none`; cold-cache, serialised launches)\n']
for rep in reps:
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    if len(rows) < 3:
        lines.append(f'\n## {os.path.basename(rep)}: no launches\n')
        continue
    H, U = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(H)}
    lines.append(f'\n## {os.path.basename(rep)}\n')
    lines.append('| # | kernel | ' + ' | '.join(c for _, c in COLS) + ' |')
    lines.append('|---|---|' + '---|' * len(COLS))
    for n, r in enumerate(rows[2:]):
        name = r[idx['Kernel Name']] if 'Kernel Name' in idx else '?'
        name = name.split('(')[0].replace('void ', '').replace('lsq::', '')[:48]
        cells = []
        for key, _ in COLS:
            if key not in idx:
                cells.append('')
                continue
            v, u = r[idx[key]].replace(',', ''), U[idx[key]]
            try:
                f = float(v) * SCALE.get(u, 1.0)
                cells.append(f'{f:.1f}' if abs(f) < 1e6 and f != int(f) else f'{int(f)}')
            except ValueError:
                cells.append(v)
        lines.append(f'| {n} | `{name}` | ' + ' | '.join(cells) + ' |')
open(out, 'w').write('\n'.join(lines) + '\n')
print(open(out).read()[:3000])
