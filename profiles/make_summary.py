#!/usr/bin/env python
"""Assembles profiles/r01_<tag>_summary.md from one gpurun call's outputs:
    gpurun_out/bench_r01_<tag>.log            (bench.py without profiler)
    gpurun_out/launches_r01_<tag>.csv         (ncu --metrics gpu__time_duration.sum launch list)
    gpurun_out/prof_r01_<tag>_cycle.ncu-rep   (ncu --set full, the 16 kernels of one cycle)
and refreshes profiles/traffic.json.   python profiles/make_summary.py v5 "<what changed>"
"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
changed = sys.argv[2] if len(sys.argv) > 2 else ""
go = os.path.join(ROOT, "gpurun_out")
bench = None
for line in open(os.path.join(go, f"bench_r01_{tag}.log")):
    if line.startswith('{"metric"'):
        bench = json.loads(line)
assert bench, "no bench line"
shutil.copy(os.path.join(go, f"launches_r01_{tag}.csv"), os.path.join(ROOT, "profiles", f"r01_{tag}_ncu_launches.csv"))
table = subprocess.run([sys.executable, os.path.join(ROOT, "profiles", "summarise_ncu.py"), os.path.join(go, f"prof_r01_{tag}_cycle.ncu-rep"),
                        "--traffic", os.path.join(ROOT, "profiles", "traffic.json")], capture_output=True, text=True).stdout
rows = [r for r in csv.reader(open(os.path.join(ROOT, "profiles", f"r01_{tag}_ncu_launches.csv"))) if len(r) > 8]
h = rows[0]
iK, iV, iG = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size")
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for r in rows[1:]:
    k = (r[iK][:60], r[iG])
    v = float(r[iV].replace(",", "")) / 1e3
    agg[k][0] += 1
    agg[k][1] += v
    tot += v
shares = ["| share | launches | avg us | kernel |", "|---|---|---|---|"]
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:18]:
    shares.append(f"| {100 * t / tot:.1f} % | {n} | {t / n:.1f} | {k[0]} grid {k[1]} |")
rf, e2e, cb = bench["roofline"], bench["e2e"], bench.get("cpu_baseline", {})
bc = bench.get("batched_chains", {})
doc = f"""# Round 1, {tag} kernels -- 4096^2, m = 32, 8 levels, V(1,1) SSOR (config C3)

{changed}

Commands (each after the same command line exited 0 without ncu, same gpurun call):
`python bench.py --steps {bench['steps']} --warmup {bench['warmup']}` -> below
`ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-batched` -> `r01_{tag}_ncu_launches.csv`
`ncu --set full --clock-control none --import-source on -k regex:"fused_smooth_kernel|trimv_kernel" -s 48 -c 16 ...` (the 16 kernels of one MGMC cycle) -> table below, `traffic.json` (`profiles/summarise_ncu.py`)

bench.py on the same build without profiler (B200, {bench['clocks']['sm_mhz']:.0f} MHz, throttle reasons {bench['clocks']['reasons']}): **{bench['value']:.1f} samples/s**
device-resident ({bench['ms_per_step']:.3f} ms / cycle, {bench['site_updates_per_sec']:.3g} site-updates/s, **{100 * bench['cycle_roofline_frac']:.1f} % of the measured
{rf['peak']:.0f} GB/s** against the algorithmic 134 B/site model), dominant kernel {rf['kernel']} at {rf['achieved']:.0f} GB/s algorithmic =
**{100 * rf['frac']:.1f} % of measured peak** ({rf['avg_launch_ms'] * 1e3:.0f} us per launch, real DRAM traffic {rf['traffic'] / 1e6:.0f} MB against {rf['algorithmic_bytes_per_launch'] / 1e6:.0f} MB algorithmic);
e2e (literal apply(f, x) with {e2e['h2d_bytes_per_step'] / 1e6:.0f} MB host x each way, PCIe-bound) {e2e['value']:.1f} samples/s; device-resident loop with host QoI series
{bench['e2e_resident']['value']:.0f} samples/s; 4 chains per launch {bc.get('value', float('nan')):.0f} chain-samples/s; CPU oracle {cb.get('value', float('nan')):.3f} samples/s-equivalent ({cb.get('cores', 1)} core).

## One cycle (ncu --set full; per-launch times are cold-cache and serialised)

{table}
## Launch list shares (first 400 launches incl. set-up)

""" + "\n".join(shares) + "\n"
open(os.path.join(ROOT, "profiles", f"r01_{tag}_summary.md"), "w").write(doc)
print(doc[:1500])
