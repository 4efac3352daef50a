#!/usr/bin/env python
"""Copy the summaries of one scripts/collect_profiles.sh run from gpurun_out/ into profiles/ (tracked) and rebuild
profiles/traffic_<tag>.json (DRAM bytes per launch of every captured kernel next to its algorithmic bytes).
usage: python scripts/summarise_profiles.py [tag]"""
import json
import os
import re
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
src, dst = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
copies = {f"{tag}_bench.json": f"bench_{tag}.json", f"{tag}_bench_ref.json": f"bench_ref_{tag}.json", f"{tag}_launches.csv": f"launches_{tag}.csv",
          f"{tag}_lims_loose.json": f"lims_loose_{tag}.json", f"{tag}_lims_tight.json": f"lims_tight_{tag}.json", f"{tag}_ltv.log": f"ltv_{tag}.log",
          f"{tag}_memcheck.log": f"sanitizer_memcheck_{tag}.log", f"{tag}_racecheck.log": f"sanitizer_racecheck_{tag}.log",
          f"{tag}_smoke.log": f"smoke_{tag}.log", f"{tag}_solve_c3.json": f"solve_c3_{tag}.json",
          f"{tag}_synccheck.log": f"sanitizer_synccheck_{tag}.log", f"{tag}_c1.json": f"c1_{tag}.json", f"{tag}_single.json": f"single_trajectory_{tag}.json"}
for k in ("bp_tile", "fwd_lin", "bp_small", "fwd_pend", "kl_tile", "kl_cached", "bp_tile_lims", "bp_tile_gps"):
    copies[f"{tag}_{k}.txt"] = f"ncu_{k}_{tag}.txt"
for a, b in copies.items():
    if os.path.exists(os.path.join(src, a)):
        shutil.copyfile(os.path.join(src, a), os.path.join(dst, b))
    else:
        print("missing", a)

n, m, T = 32, 8, 256
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
ALG = {  # algorithmic bytes per trajectory and the batch of the capture
    "bp_tile": ("bp_tile32x8_kernel<LTI> at B = 65536", 8.0 * (n * n + n * m + (T - 1) * (n + m) + n + (T - 1) * (n * m + m + n) + n + 2 + n * n) + 4, 65536),
    "fwd_lin": ("fwd_lin32x8_kernel at B = 65536", None, 65536),
    "bp_small": ("bp_small_kernel<4,1> at B = 262144", None, 262144),
    "fwd_pend": ("fwd_pend_staged_kernel at B = 262144", None, 262144),
    "kl_tile": ("kl_tile32x8_kernel<0> at B = 16384", None, 16384),
    "kl_cached": ("kl_tile32x8_kernel<2> on the 50912 cached trajectories of the C4 batch", 8.0 * (T * (2 * 256 + 2 * 32 + 8 + 3 * 64 + 528) + 1), 50912),
}
old = {}
for t in (tag, "r02"):
    pth = os.path.join(dst, f"traffic_{t}.json")
    if os.path.exists(pth):
        old = json.load(open(pth))
        break
out = {}
for k, (name, alg, B) in ALG.items():
    f = os.path.join(dst, f"ncu_{k}_{tag}.txt")
    if not os.path.exists(f):
        continue
    txt = open(f).read()
    g = lambda key: (lambda mm: float(mm.group(1)) * UNIT[mm.group(2)] if mm else None)(re.search(key + r"\s+([0-9.]+) (\w+)", txt))
    rd, wr = g("dram__bytes_read.sum"), g("dram__bytes_write.sum")
    ms = float(re.search(r"gpu__time_duration.sum\s+([0-9.]+) ms", txt).group(1))
    algb = old.get(k, {}).get("algorithmic_bytes_per_launch") or (alg * B if alg else None)      # (bench.py's BYTES_* constants x batch)
    out[k] = dict(kernel=name, dram_bytes_read=rd, dram_bytes_write=wr, dram_bytes_per_launch=rd + wr, algorithmic_bytes_per_launch=algb,
                  ratio=(rd + wr) / algb if algb else None, ncu_ms=ms, dram_gbs_under_ncu=(rd + wr) / ms * 1e-6)
for k in ("bp_tile32x8_kernel_dram_bytes_per_launch", "how"):
    if k in old:
        out[k] = old[k]
if "bp_tile" in out:
    out["bp_tile32x8_kernel_dram_bytes_per_launch"] = out["bp_tile"]["dram_bytes_per_launch"]
out["how"] = f"scripts/summarise_profiles.py {tag}: dram__bytes_read.sum + dram__bytes_write.sum of the ncu --set full captures of scripts/collect_profiles.sh {tag}"
json.dump(out, open(os.path.join(dst, f"traffic_{tag}.json"), "w"), indent=1)
for k, v in out.items():
    if isinstance(v, dict):
        print(k, f"{v['ncu_ms']:.2f} ms", f"ratio {v['ratio']:.3f}" if v["ratio"] else "")
