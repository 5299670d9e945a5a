"""Summarise .ncu-rep files: python scripts/ncu_summary.py out.md rep1 rep2 ...  (key raw metrics per kernel)."""
import csv, subprocess, sys, json, os
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__ops_path_tensor_op_utcimma_src_int8_sparsity_off.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'lts__t_sector_hit_rate.pct', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.avg',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed.avg.per_cycle_elapsed']
out, reps = sys.argv[1], sys.argv[2:]
lines, traffic = ['# ncu --set full summaries (round 1, final kernels)\n'], {}
for rep in reps:
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    H, U, V = rows[0], rows[1], rows[-1]
    d = {k: (v, u) for k, u, v in zip(H, U, V)}
    name = d.get('Kernel Name', ('?', ''))[0]
    lines.append(f'\n## {os.path.basename(rep)} — `{name[:90]}`\n\n| metric | value | unit |\n|---|---|---|')
    for k in KEYS:
        if k in d:
            lines.append(f'| {k} | {d[k][0]} | {d[k][1]} |')
    try:
        def tobytes(k):
            v, u = d[k]
            return float(v.replace(',', '')) * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]
        traffic[os.path.basename(rep)] = tobytes('dram__bytes_read.sum') + tobytes('dram__bytes_write.sum')
    except Exception as e:  # noqa: BLE001
        print('traffic', rep, e)
open(out, 'w').write('\n'.join(lines) + '\n')
print(json.dumps(traffic, indent=1))
