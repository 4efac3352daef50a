"""Write a text summary of an .ncu-rep (raw page key metrics + stall mix + hottest source lines)."""
import csv, io, subprocess, sys, json
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]
d = {}
lines = [f"# ncu summary of {rep} (ncu --set full --clock-control none; one launch, cold-cache, serialised)"]
for h, u, v in zip(hdr, units, vals):
    if h in want:
        d[h] = v
        lines.append(f"{h:95s} {v} {u}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(srows) if "# Samples" in r][0]
sh = srows[hi]; ci = {h: i for i, h in enumerate(sh)}
stalls = [h for h in sh if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: 0.0 for s in stalls}; tot = 0.0; ops = {}
for r in srows[hi + 1:]:
    try: s_ = float(r[ci["# Samples"]])
    except Exception: continue
    tot += s_
    t = r[ci["Source"]].split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    ops[op] = ops.get(op, 0.0) + s_
    for s in stalls:
        try: agg[s] += float(r[ci[s]])
        except Exception: pass
lines.append("\n# warp-stall sampling (share of all samples)")
for s, v in sorted(agg.items(), key=lambda t: -t[1])[:7]:
    lines.append(f"{s:30s} {v / tot * 100:5.1f}%")
lines.append("\n# samples by SASS opcode")
for op, v in sorted(ops.items(), key=lambda t: -t[1])[:10]:
    lines.append(f"{op:12s} {v / tot * 100:5.1f}%")
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
