#!/usr/bin/env python
"""Turns an `ncu --set full` capture of the kernels of ONE unit of the device-resident sampling loop (merged level-0
launch, 2 x 6 fused launches of levels 1-6, the coarse phase; any cyclic order) into the per-launch table of
profiles/r02_summary.md and into profiles/traffic.json (DRAM bytes per launch, read by bench.py for roofline.traffic).

    python profiles/summarise_ncu_r02.py gpurun_out/r02_prof.ncu-rep [--traffic profiles/traffic.json]
"""
import csv
import io
import json
import re
import subprocess
import sys


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, units = rows[0], rows[1]
    recs = [dict(zip(h, r)) for r in rows[2:]]

    def grid(r):
        return int(r["Grid Size"].strip("()").split(",")[0])

    def targs(r):
        m = re.search(r"<([^>]*)>", r["Kernel Name"])
        return [int(a) for a in m.group(1).split(",")] if m else []

    # levels of the 4-colour launches: grids in descending order within the pre- / post-smoothing launches
    pre = sorted({grid(r) for r in recs if "fused_smooth" in r["Kernel Name"] and targs(r)[0] == 4 and targs(r)[3] == 1}, reverse=True)
    post = sorted({grid(r) for r in recs if "fused_smooth" in r["Kernel Name"] and targs(r)[0] == 4 and targs(r)[2] == 1}, reverse=True)
    seen, table = set(), []
    for r in recs:
        kn = r["Kernel Name"]
        if "fused_smooth" in kn:
            nc, gib, pr, rs = targs(r)[:4]
            if nc == 2:
                name = "gibbs_rb8+prolong+restrict/L0" if (pr and rs) else ("gibbs_rb4+restrict/L0" if rs else "gibbs_rb4+prolong/L0")
            else:
                name = f"gibbs_4c8+restrict/L{1 + pre.index(grid(r))}" if rs else f"gibbs_4c8+prolong/L{1 + post.index(grid(r))}"
        elif "tail_kernel" in kn:
            name = "coarse sample (tail_kernel, 1 phase)"
        else:
            name = kn.split("(")[0].replace("void ", "")[:28]
        if name in seen:
            continue  # (the capture window may wrap around the unit)
        seen.add(name)
        table.append((name, r))

    def f(r, key, scale=1.0):
        try:
            return float(r[key].replace(",", "")) * scale
        except Exception:
            return float("nan")

    def unit_scale(key, want):  # ncu prints bytes in the unit of the column header row
        u = units[h.index(key)].split("/")[0]
        return {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u] / want

    print("| slot | us | DRAM read MB | DRAM write MB | DRAM GB/s | issue active % | fp64 pipe % | LSU wavefronts % | warps active % | inst (M warp) | grid | regs | smem KB |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
    traffic = {}
    tot = 0.0
    for name, r in table:
        rd = f(r, "dram__bytes_read.sum", unit_scale("dram__bytes_read.sum", 1e6))
        wr = f(r, "dram__bytes_write.sum", unit_scale("dram__bytes_write.sum", 1e6))
        us = f(r, "gpu__time_duration.sum") * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(units[h.index("gpu__time_duration.sum")], 1.0)
        tot += us
        traffic[name] = (rd + wr) * 1e6
        print(f"| {name} | {us:.1f} | {rd:.1f} | {wr:.1f} | {(rd + wr) / us * 1e3:.0f} | {f(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f} | "
              f"{f(r, 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active'):.1f} | "
              f"{f(r, 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed'):.1f} | "
              f"{f(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} | {f(r, 'smsp__inst_executed.sum') / 1e6:.1f} | "
              f"{r.get('Grid Size', '')} | {r.get('launch__registers_per_thread', '')} | "
              f"{f(r, 'launch__shared_mem_per_block_dynamic', unit_scale('launch__shared_mem_per_block_dynamic', 1e3)):.1f} |")
    print(f"\nsum of the serialised, cold-cache launch times of one unit: {tot:.1f} us")
    if "--traffic" in sys.argv:
        path = sys.argv[sys.argv.index("--traffic") + 1]
        try:
            old = json.load(open(path))
        except Exception:
            old = {}
        old.update(traffic)
        json.dump(old, open(path, "w"), indent=1, sort_keys=True)

    # stall reasons of the dominant launch
    top = max(table, key=lambda t: f(t[1], "gpu__time_duration.sum"))
    print(f"\nstall reasons of {top[0]} (warps per issue-active cycle):")
    st = sorted(((k, f(top[1], k)) for k in h if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and "not_issued" not in k),
                key=lambda kv: -kv[1] if kv[1] == kv[1] else 0)
    for k, v in st[:12]:
        print(f"  {k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):24s} {v:.2f}")


if __name__ == "__main__":
    main()
