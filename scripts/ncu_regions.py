"""Warp-stall samples of bconv_tc by kernel region (roles / epilogue phases): python scripts/ncu_regions.py rep
Regions are line ranges of ml_quant_b200/csrc/lsq_bconv_tc.cu found from its section comments."""
import csv, os, re, subprocess, sys
rep = sys.argv[1]
src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'ml_quant_b200', 'csrc', 'lsq_bconv_tc.cu')).read().split('\n')
def line_of(pat, start=0):
    for i in range(start, len(src)):
        if re.search(pat, src[i]):
            return i + 1
    raise SystemExit(f'pattern not found: {pat}')
epi = line_of(r'=+ epilogue \(8 warps\)')
loop = line_of(r'float ws = 0.0f, bs = 0.0f;')
scl = line_of(r'per-position activation scales of this warp', loop)
tm = line_of(r'thread = channel: two halves', loop)
st = line_of(r'lane = position: activation, residual', loop)
tail = line_of(r'co.step == spi - 1', loop)
mma = line_of(r'=+ MMA issuer')
wl = line_of(r'=+ weight loader')
pr = line_of(r'=+ patch producers')
end = line_of(r'^static|^bool|tc_plan', pr)
regions = [('epilogue: cursor setup, residual prefetch (cp.async), scale requests', epi, scl - 2), ('epilogue: scale table, accumulator wait', scl - 1, tm - 1),
           ('epilogue: TMEM -> scaled transposition tile', tm, st - 1), ('epilogue: activation, residual, global stores', st, tail - 1),
           ('epilogue: accumulator release, advance', tail, mma - 1), ('MMA issuer', mma, wl - 1), ('weight loader', wl, pr - 1), ('patch producers', pr, end)]
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'], capture_output=True, text=True).stdout
hdr, fname, data = None, '', []
for r in csv.reader(out.splitlines()):
    if len(r) == 2 and r[0] == 'File Path':
        fname = r[1].split('/')[-1]
    elif r and r[0] == 'Line No':
        hdr = r
    elif hdr and len(r) == len(hdr) and r[0].strip().isdigit():
        data.append((fname, int(r[0]), r))
idx = {n: i for i, n in enumerate(hdr)}
cols = [c for c in hdr if c.startswith('stall_') and 'Not Issued' not in c]
def num(v):
    try:
        return int(v)
    except ValueError:
        return 0
total = sum(num(r[4]) for _, _, r in data)
print(f'| region (lsq_bconv_tc.cu lines) | samples | share | top stall reasons |\n|---|---|---|---|')
rows = []
for name, lo, hi in regions:
    sel = [r for f, ln, r in data if f == 'lsq_bconv_tc.cu' and lo <= ln <= hi]
    rows.append((f'{name} ({lo}-{hi})', sel))
rows.append(('inlined helpers: decode_pos / out_offset arithmetic (lsq_bconv_tc.cu < %d)' % epi, [r for f, ln, r in data if f == 'lsq_bconv_tc.cu' and ln < epi]))
rows.append(('lsq_tc.cuh (mbarrier waits, tcgen05 / TMA wrappers)', [r for f, ln, r in data if f == 'lsq_tc.cuh']))
for name, sel in rows:
    n = sum(num(r[4]) for r in sel)
    by = sorted(((sum(num(r[idx[c]]) for r in sel), c.replace('stall_', '')) for c in cols), reverse=True)[:4]
    print(f'| {name} | {n} | {100.0 * n / max(total, 1):.1f} % | ' + ', '.join(f'{c} {v}' for v, c in by) + ' |')
print(f'\ntotal samples {total}')
