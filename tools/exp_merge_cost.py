"""Development aid: what would one level-0 launch per cycle cost?  Times the level-0 launches of V(1,1) and V(2,2) cycles
without measurements on 4096^2 (a V(2,2) launch is 8 colour passes, 5 of them live: the pass count of the merged
post-smoothing + pre-smoothing launch)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import multigridmc_b200 as m

for pre, post in [(1, 1), (2, 2)]:
    ctx = m.Context(4096, 4096, 8, seed=1, npresmooth=pre, npostsmooth=post)
    nd = ctx.ndof()
    rng = np.random.default_rng(0)
    ctx.set_rhs(rng.standard_normal(nd))
    ctx.set_state(np.zeros(nd))
    ctx.sample(3, series=False)
    prof = ctx.profile_cycle(10)
    print(f"V({pre},{post}) m=0:", [(n, round(1e3 * ms / 10, 1)) for n, ms, l, b in prof if n.endswith("/L0") or n.endswith("/L1")])
    ctx.close()
