import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import multigridmc_b200 as m
nch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ctx = m.Context(512, 512, 5, Lambda=0.2, nchains=nch)
nd = ctx.ndof()
rng = np.random.default_rng(0)
ctx.set_rhs(np.tile(rng.standard_normal(nd), nch))
ctx.set_state(np.zeros(nd * nch))
ctx.set_qoi([nd // 2], [1.0])
ctx.sample(5, series=False)
ms, _ = ctx.sample_timed(20)
print("C5", nch, "chains: ms/cycle", ms / 20, "chain-samples/s", nch * 20 / ms * 1e3)
prof = ctx.profile_cycle(6)
tot = sum(p[1] for p in prof)
for n, t, l, b in sorted(prof, key=lambda p: -p[1]):
    print(f"  {n:36s} {1e3 * t / 6:8.1f} us/cycle  {l / 6:.2f} launches/cycle  {100 * t / tot:.1f} %")
