"""Top stall-sampled SASS instructions from `ncu --page source --csv` output."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
h = rows[hdr]
si, ss = h.index('Source'), h.index('Warp Stall Sampling (All Samples)')
data = [(int(r[ss] or 0), i, r[si].strip()) for i, r in enumerate(rows[hdr + 1:]) if len(r) > ss]
total = sum(d[0] for d in data)
print('total samples', total)
for s, i, src in sorted(data, reverse=True)[:top]:
    print(f"{s:7d} {100.0 * s / max(total, 1):5.1f}%  #{i:5d}  {src[:110]}")
