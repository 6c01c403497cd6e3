"""Development aid: where does a chain with merged level-0 launches differ from the unmerged one?"""
import os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SNIP = r'''
import sys, numpy as np
sys.path.insert(0, %(root)r)
import multigridmc_b200 as m
from multigridmc_b200 import workloads as w
n, nlevel, nmeas, K = %(n)d, %(nlevel)d, %(nmeas)d, %(K)d
loc, _, _, var = w.measurement_set(nmeas)
B = w.point_measurement_matrix(n, n, loc, var, 1e-3)
ctx = m.Context(n, n, nlevel, B=B, seed=4711)
rng = np.random.default_rng(3)
nd = ctx.ndof()
ctx.set_rhs(rng.standard_normal(nd))
ctx.set_state(rng.standard_normal(nd))
ctx.set_qoi([nd // 2], [1.0])
ctx.set_philox_position(0)
z = ctx.sample(K)
np.save(%(out)r, np.concatenate([ctx.get_state(), np.asarray(z).ravel()]))
'''
n, nlevel, nmeas = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
sys.path.insert(0, ROOT)
from multigridmc_b200 import workloads as w
loc = np.asarray(w.measurement_set(nmeas)[0])
print("measurement sites (i, j):", [(round(float(x) * n, 2), round(float(y) * n, 2)) for x, y in loc])
for K in [int(k) for k in sys.argv[4].split(",")]:
    res = {}
    for tag, env in {"merged": {}, "unmerged": {"MGMC_NO_MERGE": "1"}, "merged_all": {"MGMC_NO_DEAD_PASS": "1"}, "merged_nograph": {"MGMC_NO_GRAPH": "1"}, "merged_again": {}}.items():
        out = f"/tmp/dbg_{tag}.npy"
        e = dict(os.environ); e.update(env)
        subprocess.check_call([sys.executable, "-c", SNIP % dict(root=ROOT, n=n, nlevel=nlevel, nmeas=nmeas, K=K, out=out)], env=e)
        res[tag] = np.load(out)
    w_ = n - 1
    for tag in ("merged", "merged_all", "merged_nograph", "merged_again"):
        d = np.abs(res[tag][:w_ * w_] - res["unmerged"][:w_ * w_]).reshape(w_, w_)
        bad = np.argwhere(d > 0)
        print(f"K={K} {tag}: max diff {d.max():.3e}, {len(bad)} sites differ; series diff {np.abs(res[tag][w_ * w_:] - res['unmerged'][w_ * w_:])}")
        if len(bad):
            jj, ii = bad[:, 0] + 1, bad[:, 1] + 1
            print("   rows", jj.min(), "..", jj.max(), " cols", ii.min(), "..", ii.max())
            big = np.argwhere(d > 0.5 * d.max())
            print("   largest at (i, j):", [(int(b[1]) + 1, int(b[0]) + 1) for b in big[:10]])
