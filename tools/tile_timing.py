"""Development aid: summarise the per-CTA phase stamps of a fused level launch written by a -DMGMC_TILE_TIMING build
(tools/build_timing.sh, MGMC_TIMING_FILE=<prefix> -> <prefix>.<kernel name>.txt)."""
import sys
import numpy as np

names = ["start", "load", "seg0", "fix0", "seg1", "fix1", "seg2", "pre-st", "store", "resid", "lr-res", "end"]
for f in sys.argv[1:]:
    a = np.loadtxt(f)
    a = a[a[:, 0] > 0]
    st = a[:, 0:12].copy()
    st[st <= 0] = np.nan
    rel = (st - st[:, [0]]) / 1e3
    med = np.nanmedian(rel, axis=0)
    print(f, len(a), "tiles; kernel span us", (np.nanmax(st[:, 11]) - np.nanmin(st[:, 0])) / 1e3)
    print("  median stamps us:", " ".join(f"{n}={v:.2f}" for n, v in zip(names, med)))
    print("  warp 0 cycles in pass set-up / rows / barrier wait (median):", np.median(a[:, 12]), np.median(a[:, 13]), np.median(a[:, 14]))
