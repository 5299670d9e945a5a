"""Summarise an ncu launch list (gpu__time_duration per launch) by kernel: python scripts/launch_summary.py csv [out.md]"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
h = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
H = rows[h]; ki, vi, ui = H.index('Kernel Name'), H.index('Metric Value'), H.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows[h + 1:]:
    if len(r) <= vi:
        continue
    try:
        v = float(r[vi].replace(',', ''))
    except ValueError:
        continue
    v *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 'usecond': 1e-3, 'nsecond': 1e-6, 'msecond': 1.0}.get(r[ui], 1e-6)
    a = agg.setdefault(r[ki][:100], [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
lines = [f'total {tot:.3f} ms over {sum(a[0] for a in agg.values())} launches', '', '| share | ms | launches | kernel |', '|---|---|---|---|']
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append(f'| {100 * a[1] / tot:.2f}% | {a[1]:.3f} | {a[0]} | `{k}` |')
text = '\n'.join(lines)
print(text)
if len(sys.argv) > 2:
    open(sys.argv[2], 'w').write('# One forward step, ImageNet ResNet-18 ls1w/ls2a, batch 512, fused blocks, eager (no graph)\n\n'
                                 '`ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none python profiles/profile_step.py` '
                                 '(cold-cache, serialised: compare shares)\n\n' + text + '\n')
