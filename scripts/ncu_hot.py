"""Summarise an `ncu --page source --csv` export: stall-reason totals and the hottest SASS lines."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[2:]:
    try:
        data.append((float(r[ci["# Samples"]]), r))
    except Exception:
        pass
tot = sum(d[0] for d in data)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: 0.0 for s in stalls}
for s_, r in data:
    for s in stalls:
        try: agg[s] += float(r[ci[s]])
        except Exception: pass
print("total samples", tot)
for s, v in sorted(agg.items(), key=lambda t: -t[1])[:8]:
    print(f"  {s:28s} {v / max(tot,1) * 100:5.1f}%")
data.sort(key=lambda t: -t[0])
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
for s_, r in data[:n]:
    top = sorted(((float(r[ci[s]] or 0), s) for s in stalls), reverse=True)[:2]
    print(f"{s_ / tot * 100:5.1f}% {r[ci['Source']][:90]:90s} {top[0][1]}:{top[0][0]:.0f} {top[1][1]}:{top[1][0]:.0f}")
