"""Per-source-line instruction counts / stall samples from `ncu --page source --csv --print-source cuda,sass`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
div = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hdr = None
lines = []
for r in rows:
    if r and r[0] == 'Line No':
        hdr = r; continue
    if hdr is None or len(r) < len(hdr): continue
    if r[0] != '':  # a CUDA line summary row
        lines.append(r)
ie = hdr.index('Instructions Executed'); ss = hdr.index('# Samples')
tot = sum(int(r[ie]) for r in lines if r[ie].isdigit()); tots = sum(int(r[ss]) for r in lines if r[ss].isdigit())
print('total inst', tot, 'per-unit', tot / div, 'samples', tots)
stall_cols = [i for i, c in enumerate(hdr) if c.startswith('stall_') and 'Not Issued' not in c]
for r in lines:
    if not r[ie].isdigit(): continue
    if int(r[ie]) > tot * 0.003 or int(r[ss]) > tots * 0.005:
        st = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stall_cols), reverse=True)[:3]
        sts = ' '.join(f"{n}:{v}" for v, n in st if v)
        print(f"{r[0]:>5} {int(r[ie]) / div:8.1f} {100 * int(r[ss]) / tots:5.1f}%  {r[1][:90]:90s} {sts}")
