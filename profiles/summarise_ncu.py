#!/usr/bin/env python
"""Turns an `ncu --set full` capture of the kernels of ONE MGMC cycle (14 fused launches + the two coarse
triangular mat-vecs, in launch order) into the per-launch table of profiles/*_summary.md and into
profiles/traffic.json (DRAM bytes per launch, read by bench.py for roofline.traffic).

    python profiles/summarise_ncu.py gpurun_out/prof_cycle.ncu-rep [--traffic profiles/traffic.json]
"""
import csv
import io
import json
import subprocess
import sys


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, units = rows[0], rows[1]
    recs = [dict(zip(h, r)) for r in rows[2:]]
    fused = [r for r in recs if "fused_smooth_kernel" in r["Kernel Name"]]
    nlev = len(fused) // 2
    names = {}
    k = 0
    for r in recs:
        if "fused_smooth_kernel" in r["Kernel Name"]:
            nc = "rb4" if "<2," in r["Kernel Name"] else "4c8"
            names[r["ID"]] = f"gibbs_{nc}+restrict/L{k}" if k < nlev else f"gibbs_{nc}+prolong/L{2 * nlev - 1 - k}"
            k += 1
        else:
            names[r["ID"]] = r["Kernel Name"].split("(")[0].replace("void ", "")[:28]

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
    for r in recs:
        rd = f(r, "dram__bytes_read.sum", unit_scale("dram__bytes_read.sum", 1e6))
        wr = f(r, "dram__bytes_write.sum", unit_scale("dram__bytes_write.sum", 1e6))
        us = f(r, "gpu__time_duration.sum") * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(units[h.index("gpu__time_duration.sum")], 1.0)
        tot += us
        traffic[names[r["ID"]]] = (rd + wr) * 1e6
        print(f"| {names[r['ID']]} | {us:.1f} | {rd:.1f} | {wr:.1f} | {(rd + wr) / us * 1e3:.0f} | {f(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f} | "
              f"{f(r, 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active'):.1f} | "
              f"{f(r, 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed'):.1f} | "
              f"{f(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} | {f(r, 'smsp__inst_executed.sum') / 1e6:.1f} | "
              f"{r.get('Grid Size', '')} | {r.get('launch__registers_per_thread', '')} | "
              f"{f(r, 'launch__shared_mem_per_block_dynamic', unit_scale('launch__shared_mem_per_block_dynamic', 1e3)):.1f} |")
    print(f"\nsum of the serialised, cold-cache launch times: {tot:.1f} us")
    if "--traffic" in sys.argv:
        json.dump(traffic, open(sys.argv[sys.argv.index("--traffic") + 1], "w"), indent=1)


if __name__ == "__main__":
    main()
