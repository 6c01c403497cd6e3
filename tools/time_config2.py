import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import multigridmc_b200 as m
ctx = m.Context(1024, 1024, 6, Lambda=0.2, npresmooth=2, npostsmooth=2)
rng = np.random.default_rng(0)
b = rng.standard_normal(ctx.ndof())
ctx.loop_solve(b, rtol=1e-12, atol=1e-15, maxiter=3)
t0 = time.perf_counter()
x, hist, it, cv = ctx.loop_solve(b, rtol=1e-12, atol=1e-15, maxiter=100)
dt = time.perf_counter() - t0
print("C2", os.environ.get("MGMC_TAIL"), "iterations", len(hist), "ms/iter", 1e3 * dt / len(hist), "reduction", hist[-1] / hist[0])
prof = ctx.profile_cycle(3) if False else None
