"""Rank source lines of an .ncu-rep by warp-stall samples: python scripts/ncu_lines.py rep [top]"""
import csv, subprocess, sys
rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, fname, data = None, '', []
for r in rows:
    if len(r) == 2 and r[0] == 'File Path':
        fname = r[1].split('/')[-1]
    elif r and r[0] == 'Line No':
        hdr = r
    elif hdr and len(r) == len(hdr) and r[0].strip().isdigit():
        try:
            data.append((int(r[4]), fname, r))
        except ValueError:
            pass
H = hdr
idx = {n: i for i, n in enumerate(H)}
tot = sum(d[0] for d in data)
print('total samples', tot)
cols = ['stall_long_sb', 'stall_short_sb', 'stall_barrier', 'stall_wait', 'stall_mio', 'stall_lg', 'stall_math', 'stall_branch_resolving', 'stall_no_inst', 'stall_not_selected', 'stall_selected']
print('   n     %   file:line  ' + ' '.join(c.replace('stall_', '')[:7] for c in cols))
for n, f, r in sorted(data, key=lambda t: -t[0])[:top]:
    st = ' '.join(f'{r[idx[c]]:>7s}' for c in cols if c in idx)
    print(f'{n:6d} {100 * n / tot:5.1f}% {f}:{r[0]:>4s} {st} | {r[1].strip()[:90]}')
