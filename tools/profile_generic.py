"""Per-launch CUDA-event profile of one MGMC cycle for the operator families that run one launch per colour
(radius-2 / per-vertex / 3d): python tools/profile_generic.py {c4|3d|3dfem|periodic}"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import multigridmc_b200 as m

which = sys.argv[1] if len(sys.argv) > 1 else "3d"
if which == "c4":
    ctx = m.Context(2048, 2048, 7, Lambda=0.2, pde="squared_shiftedlaplace_fd")
elif which == "3d":
    ctx = m.Context(128, 128, 5, nz=128, Lambda=0.2)
elif which == "3dfem":
    ctx = m.Context(128, 128, 5, nz=128, Lambda=0.2, pde="shiftedlaplace_fem")
else:
    ctx = m.Context(2048, 2048, 7, Lambda=0.2, kappa_sq=m.periodic_kappa_sq(2048, 2048, 0.1, 0.4))
nd = ctx.ndof()
rng = np.random.default_rng(0)
ctx.set_rhs(rng.standard_normal(nd))
ctx.set_state(np.zeros(nd))
ctx.set_qoi([nd // 2], [1.0])
ctx.sample(5, series=False)
ms, _ = ctx.sample_timed(20)
print(which, "ms/cycle (graph)", ms / 20)
prof = ctx.profile_cycle(6)
tot = sum(p[1] for p in prof)
print("sum of the eager per-launch times: %.1f us/cycle, %d launches/cycle" % (1e3 * tot / 6, sum(p[2] for p in prof) / 6))
for n, t, l, b in sorted(prof, key=lambda p: -p[1])[:40]:
    print(f"  {n:36s} {1e3 * t / 6:8.1f} us/cycle  {l / 6:6.2f} launches/cycle  {1e3 * t / max(l, 1):7.2f} us/launch  {100 * t / tot:5.1f} %")
