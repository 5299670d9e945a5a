"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export by CUDA source line:
    ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass > src.csv ; python scripts/ncu_lines.py src.csv [top]
Prints the lines with the most warp-stall samples (all samples), their share, executed instructions and the dominant
stall reasons -- what a kernel's time is spent on, line by line."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
fname, hdr = None, None
agg = collections.OrderedDict()
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        fname = r[1].split('/')[-1]
        continue
    if r[0] == 'Line No':
        hdr = r
        continue
    if hdr is None or r[0] == 'Function Name' or r[0] == '':
        continue
    d = {k: ('0' if v in ('-', '') else v) for k, v in zip(hdr, r)}
    if not r[0].isdigit():
        continue
    key = (fname, int(r[0]))
    try:
        samples = int(d.get('Warp Stall Sampling (All Samples)', '0') or 0)
        inst = int(d.get('Instructions Executed', '0') or 0)
        stalls = {k[6:]: int(v or 0) for k, v in d.items() if k.startswith('stall_') and 'Not Issued' not in k}
    except ValueError:      # a source line with unescaped quotes (inline asm) shifted the columns: skip it
        continue
    e = agg.setdefault(key, {'src': r[1].strip(), 'samples': 0, 'inst': 0, 'stalls': collections.Counter()})
    e['samples'] += samples
    e['inst'] += inst
    e['stalls'].update(stalls)
total = sum(e['samples'] for e in agg.values()) or 1
tinst = sum(e['inst'] for e in agg.values()) or 1
print(f'total samples {total}, warp instructions {tinst}')
for (f, ln), e in sorted(agg.items(), key=lambda kv: -kv[1]['samples'])[:top]:
    st = ', '.join(f'{k} {v * 100 // max(e["samples"], 1)}%' for k, v in e['stalls'].most_common(3) if v)
    print(f'{e["samples"] * 100.0 / total:5.1f}%  inst {e["inst"] * 100.0 / tinst:5.1f}%  {f}:{ln:<4d} {e["src"][:90]:90s} | {st}')
