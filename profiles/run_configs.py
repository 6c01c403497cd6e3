#!/usr/bin/env python
"""Times the BASELINE configurations other than the bench line (C1, C2, C4, C5) on one
GPU through the C ABI and prints one JSON line per configuration (device-resident, CUDA events)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import multigridmc_b200 as m  # noqa: E402

PEAK = 6542.1
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def mgmc(name, n, nlevel, nchains=1, pde="shiftedlaplace_fd", steps=200, **kw):
    ctx = m.Context(n, n, nlevel, Lambda=0.2, pde=pde, nchains=nchains, **kw)
    nd = ctx.ndof()
    xs = np.arange(1, n) / n
    u = np.outer(np.sin(np.pi * xs), np.sin(np.pi * xs)).ravel()
    f = ctx.op_apply(0, np.tile(u, nchains))
    ctx.set_rhs(f)
    ctx.set_state(np.zeros(nd * nchains))
    ctx.set_qoi([nd // 2], [1.0])
    ctx.sample(20, series=False)
    ms, _ = ctx.sample_timed(steps)
    byts, upd = ctx.cycle_model()
    t = ms / steps * 1e-3
    print(json.dumps({"config": name, "ms_per_cycle": ms / steps, "chain_samples_per_s": nchains / t, "site_updates_per_s": nchains * upd / t,
                      "algorithmic_gbs": nchains * byts / t / 1e9, "frac_of_measured_hbm": nchains * byts / t / 1e9 / PEAK}), flush=True)
    ctx.close()


def mg(name, n, nlevel):
    ctx = m.Context(n, n, nlevel, Lambda=0.2, npresmooth=2, npostsmooth=2)
    rng = np.random.default_rng(0)
    b = rng.standard_normal(ctx.ndof())
    ctx.loop_solve(b, rtol=1e-12, atol=1e-15, maxiter=3)
    t0 = time.perf_counter()
    x, hist, it, cv = ctx.loop_solve(b, rtol=1e-12, atol=1e-15, maxiter=100)
    dt = time.perf_counter() - t0
    print(json.dumps({"config": name, "iterations": len(hist), "ms_per_iteration_incl_host_norm_readback": 1e3 * dt / len(hist),
                      "residual_reduction": hist[-1] / hist[0], "rate_first_10": (hist[10] / hist[0]) ** 0.1}), flush=True)
    ctx.close()


if __name__ == "__main__":
    mgmc("C1 driver_mgmc 64x64 L3 prior V(1,1) SSOR", 64, 3, steps=2000)
    mg("C2 driver_mg 1024x1024 L6 V(2,2) SSOR", 1024, 6)
    mgmc("C4 (single GPU) squared_shiftedlaplace_fd 2048x2048 L7 V(1,1) SSOR", 2048, 7, pde="squared_shiftedlaplace_fd", steps=20)
    mgmc("C5 256 chains x 512x512 L5 per GPU", 512, 5, nchains=256, steps=20)
    mgmc("C5' 32 chains x 512x512 L5", 512, 5, nchains=32, steps=50)
