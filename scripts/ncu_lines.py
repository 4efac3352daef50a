"""Per-CUDA-source-line sample totals from `ncu --page source --csv --print-source cuda,sass`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if '# Samples' in r][0]
hdr = rows[hi]
si = hdr.index('# Samples'); ii = hdr.index('Instructions Executed')
lines = []
for r in rows[hi + 1:]:
    if len(r) > si and r[0] not in ('', 'Line No'):
        try:
            lines.append((int(r[0]), r[1], float(r[si]), float(r[ii] or 0)))
        except ValueError:
            pass
tot = sum(l[2] for l in lines)
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.8
print("total samples", tot)
for ln, src, s, ex in lines:
    if s / tot * 100 >= thr:
        print(f"{s / tot * 100:5.1f}%  inst {ex / 1e6:8.1f}M  L{ln:<4d} {src.strip()[:110]}")
