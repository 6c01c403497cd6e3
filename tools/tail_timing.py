"""Development aid: summarise the per-CTA phase stamps of the persistent tail kernel written by a
-DMGMC_TILE_TIMING build (tools/build_timing.sh, MGMC_TIMING_FILE=<prefix> -> <prefix>.tail.txt)."""
import sys
import numpy as np

for path in sys.argv[1:]:
    a = np.loadtxt(path)
    nph = int(a[:, 0].max()) + 1
    t00 = a[a[:, 0] == 0][:, 4].min()
    print(path)
    tot = 0.0
    for p in range(nph):
        r = a[a[:, 0] == p]
        kind, nt = int(r[0, 1]), int(r[0, 2])
        start, bbar, abar = r[:, 4], r[:, 5], r[:, 6]
        ps = start.min()
        line = f"ph{p} kind{kind} nt{nt:4d} at {(ps - t00) / 1e3:7.2f}us work(max) {(bbar.max() - ps) / 1e3:6.2f}"
        if kind == 0:
            busy = r[:nt]
            st = busy[:, 8:20]
            st = st[st[:, 0] > 0]
            rel = (st - st[:, [0]]) / 1e3
            rel[rel < 0] = np.nan
            med, mx = np.nanmedian(rel, axis=0), np.nanmax(rel, axis=0)
            names = ["", "load", "seg0", "fix0", "seg1", "fix1", "seg2", "pre-st", "store", "resid", "lr-res", "end"]
            line += " med: " + " ".join(f"{n}={v:5.2f}" for n, v in zip(names[1:], med[1:])) + " | max end " + f"{mx[11]:5.2f}"
        if p < nph - 1:
            line += f" | barrier {(abar - bbar).min() / 1e3:5.2f}"
        print(line)
    print("total", (a[:, 5].max() - t00) / 1e3, "us")
