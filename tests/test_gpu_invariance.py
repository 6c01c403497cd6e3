"""The chain must not depend on HOW the tile kernel covers the lattice: tile heights, skipping of dead colour passes
(plan_stages), residual folding.  The switches are read once per process, so every variant runs in its own process;
the states after a few cycles must agree BIT FOR BIT (every site's update is a pure function of its inputs and of the
Philox counters).  Size-independent property test of the fused kernel."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SNIPPET = r'''
import sys, numpy as np
sys.path.insert(0, %(root)r)
import multigridmc_b200 as m
from multigridmc_b200 import workloads as w
n, nlevel, nmeas, omega, out = %(n)d, %(nlevel)d, %(nmeas)d, %(omega)r, %(out)r
B = None
if nmeas:
    loc, _, _, var = w.measurement_set(nmeas)
    B = w.point_measurement_matrix(n, n, loc, var, 1e-3)
ctx = m.Context(n, n, nlevel, B=B, seed=4711, omega=omega)
rng = np.random.default_rng(3)
nd = ctx.ndof()
ctx.set_rhs(rng.standard_normal(nd))
ctx.set_state(rng.standard_normal(nd))
ctx.set_qoi([nd // 2], [1.0])
ctx.set_philox_position(0)
z = ctx.sample(4)
np.save(out, np.concatenate([ctx.get_state(), np.asarray(z).ravel()]))
'''


def _parallel(fn, variants, workers=4):
    """{tag: result}: every variant in its own process (the switches are read once per process), up to `workers` of them at a
    time on the same GPU -- the results are pure functions of the inputs, so concurrency must not change a bit."""
    from concurrent.futures import ThreadPoolExecutor

    with ThreadPoolExecutor(max_workers=workers) as ex:
        futs = {tag: ex.submit(fn, tag, env) for tag, env in variants.items()}
        return {tag: f.result() for tag, f in futs.items()}


def _run(tmp_path, tag, env, **kw):
    out = str(tmp_path / f"{tag}.npy")
    e = dict(os.environ)
    e.update(env)
    subprocess.check_call([sys.executable, "-c", SNIPPET % dict(root=ROOT, out=out, **kw)], env=e)
    return np.load(out)


@pytest.mark.gpu
@pytest.mark.parametrize("n,nlevel,nmeas,omega", [(512, 5, 0, 1.0), (512, 5, 8, 1.0), (1024, 6, 8, 1.0), (512, 5, 8, 1.3)])
def test_chain_independent_of_tiling_and_dead_pass_skipping(tmp_path, n, nlevel, nmeas, omega):
    kw = dict(n=n, nlevel=nlevel, nmeas=nmeas, omega=omega)
    variants = {
        "default": {},
        "all_passes": {"MGMC_NO_DEAD_PASS": "1"},
        "low_tiles": {"MGMC_TILE_ROWS": "32,32,16,8,32,32,16"},
        "tall_tiles": {"MGMC_TILE_ROWS": "24,40,32,16,36,24,24"},
        "no_fold": {"MGMC_NO_RES_FOLD": "1", "MGMC_TILE_ROWS": "16,16,8,8,16,16,8"},
        # the small levels and the coarse solve as phases of ONE persistent cooperative kernel (tail.cuh) instead of a
        # launch per level visit
        "tail": {"MGMC_TAIL": "1"},
        "short_tail": {"MGMC_TAIL": "1", "MGMC_TAIL_MAX_SITES": "20000"},
        # post-smoothing of cycle k and pre-smoothing of cycle k + 1 as separate level-0 launches (the default merges
        # them into one launch per cycle, the observed sites recorded inside it)
        "no_merge": {"MGMC_NO_MERGE": "1"},
        "no_merge_all_passes": {"MGMC_NO_MERGE": "1", "MGMC_NO_DEAD_PASS": "1"},
    }
    res = _parallel(lambda tag, env: _run(tmp_path, tag, env, **kw), variants)
    ref = res.pop("default")
    assert np.all(np.isfinite(ref))
    for tag, x in res.items():
        if tag == "no_fold" and nmeas == 0 and omega == 1.0:
            # the folded residual (-noise) and the stencil residual differ in the last bits: same chain to rounding
            assert np.max(np.abs(x - ref)) <= 1e-9 * np.max(np.abs(ref)), tag
        else:
            assert np.array_equal(x, ref), f"{tag}: max abs diff {np.max(np.abs(x - ref)):.3e}"


SNIPPET_AHEAD = r'''
import sys, numpy as np
sys.path.insert(0, %(root)r)
import multigridmc_b200 as m
from multigridmc_b200 import workloads as w
n, nlevel, nmeas, omega, out = %(n)d, %(nlevel)d, %(nmeas)d, %(omega)r, %(out)r
B = None
if nmeas:
    loc, _, _, var = w.measurement_set(nmeas)
    B = w.point_measurement_matrix(n, n, loc, var, 1e-3)
ctx = m.Context(n, n, nlevel, B=B, seed=99, omega=omega)
rng = np.random.default_rng(5)
nd = ctx.ndof()
f = rng.standard_normal(nd)
ctx.set_rhs(f)
ctx.set_state(rng.standard_normal(nd))
ctx.set_qoi([nd // 3], [1.0])
ctx.set_philox_position(0)
z1 = ctx.sample(3)                      # graph replays
x = ctx.mgmc_apply(f, ctx.get_state())  # one cycle outside the graph, host vectors
ctx.set_state(x)
ctx.set_philox_position(100)            # the sample index jumps
z2 = ctx.sample(2)
np.save(out, np.concatenate([ctx.get_state(), x, np.asarray(z1).ravel(), np.asarray(z2).ravel()]))
'''


@pytest.mark.gpu
@pytest.mark.parametrize("n,nlevel,nmeas,omega", [(2048, 6, 8, 1.0), (2048, 6, 8, 1.3)])  # (40 s per case on a B200)
def test_chain_independent_of_noise_generated_ahead(tmp_path, n, nlevel, nmeas, omega):
    """Normals of the small levels generated ahead of their launches by a second branch of the cycle graph
    (noise_ahead.cuh) vs generated in registers by the launches themselves: the same chain bit for bit -- through graph
    replays, a cycle outside the graph and a jump of the sample index."""
    kw = dict(n=n, nlevel=nlevel, nmeas=nmeas, omega=omega)

    def run(tag, env):
        out = str(tmp_path / f"{tag}.npy")
        e = dict(os.environ)
        e.update(env)
        subprocess.check_call([sys.executable, "-c", SNIPPET_AHEAD % dict(root=ROOT, out=out, **kw)], env=e)
        return np.load(out)

    on = {"MGMC_NOISE_AHEAD": "1"}  # (opt-in: measured, no gain -- profiles/r02_noise_ahead.md)
    res = _parallel(run, {"in_register": {}, "ahead": on, "ahead_no_merge": dict(on, MGMC_NO_MERGE="1"), "ahead_no_graph": dict(on, MGMC_NO_GRAPH="1")})
    ref = res.pop("in_register")
    assert np.all(np.isfinite(ref))
    for tag, x in res.items():
        assert np.array_equal(x, ref), f"{tag}: max abs diff {np.max(np.abs(x - ref)):.3e}"


@pytest.mark.gpu
@pytest.mark.parametrize("n,nlevel,nmeas,nchains", [(1024, 6, 8, 1), (512, 5, 8, 3)])
def test_repeated_runs_are_bit_identical(n, nlevel, nmeas, nchains):
    """Race detector of our own (compute-sanitizer is closed on the GPU pool): the chain is a pure function of its inputs,
    so two contexts fed the same inputs must agree bit for bit -- run after run, with the scheduling noise of 60-1000
    CTAs per launch in between.  A data race between CTAs (e.g. one tile zeroing what another still reads) or between
    the warps of a CTA shows up as a difference here within a few repetitions."""
    import multigridmc_b200 as m
    from multigridmc_b200 import workloads as w

    loc, _, _, var = w.measurement_set(nmeas)
    B = w.point_measurement_matrix(n, n, loc, var, 1e-3)
    rng = np.random.default_rng(11)
    ref = None
    for rep in range(4):
        ctx = m.Context(n, n, nlevel, B=B, seed=99, nchains=nchains)
        nd = ctx.ndof()
        if rep == 0:
            f, x0 = rng.standard_normal(nd * nchains), rng.standard_normal(nd * nchains)
        ctx.set_rhs(f)
        ctx.set_state(x0)
        ctx.set_qoi([nd // 2, nd // 3], [1.0, -0.5])
        ctx.set_philox_position(0)
        z = ctx.sample(2 + rep % 2 * 3)[:2]  # (2 or 5 cycles: one or several merged level-0 launches; compare the first two samples)
        out = np.concatenate([z.ravel()] + ([ctx.get_state()] if rep % 2 == 0 else []))
        ctx.close()
        if ref is None:
            ref, ref_z = out, z.copy()
        elif rep % 2 == 0:
            assert np.array_equal(out, ref), f"repetition {rep}: max abs diff {np.max(np.abs(out - ref)):.3e}"
        else:
            assert np.array_equal(z, ref_z), f"repetition {rep}: series differs"


SNIPPET_ROWS = r'''
import sys, numpy as np
sys.path.insert(0, %(root)r)
import multigridmc_b200 as m
family, omega, nmeas, out = %(family)r, %(omega)r, %(nmeas)d, %(out)r
rng = np.random.default_rng(3)
B = None
if family == "radius2":
    n = (192, 160, None)
    kw = dict(pde="squared_shiftedlaplace_fd")
elif family == "3d":
    n = (32, 24, 16)
    kw = {}
else:
    n = (160, 96, None)
    kw = dict(kappa_sq=m.periodic_kappa_sq(160, 96, 0.1, 0.4))
nd = (n[0] - 1) * (n[1] - 1) * ((n[2] - 1) if n[2] else 1)
if nmeas:
    rows = rng.choice(nd, size=nmeas, replace=False)
    B = (rows, np.arange(nmeas), np.ones(nmeas), 1e-3 * (1.0 + rng.random(nmeas)))
ctx = m.Context(n[0], n[1], 3, nz=n[2], B=B, seed=4711, omega=omega, smoother=%(smoother)r, npresmooth=2, **kw)
ctx.set_rhs(rng.standard_normal(nd))
ctx.set_state(rng.standard_normal(nd))
ctx.set_qoi([nd // 2], [1.0])
ctx.set_philox_position(0)
z = ctx.sample(3)
xs = ctx.smoother_apply(0, "SSOR", rng.standard_normal(nd), rng.standard_normal(nd), omega=omega, nsmooth=2)
np.save(out, np.concatenate([ctx.get_state(), np.asarray(z).ravel(), xs]))
'''


@pytest.mark.gpu
@pytest.mark.parametrize("family,omega,nmeas,smoother", [("radius2", 1.0, 0, "SSOR"), ("radius2", 1.2, 3, "SOR"), ("3d", 1.0, 0, "SSOR"), ("3d", 0.9, 3, "SSOR"),
                                                         ("periodic", 1.0, 0, "SSOR"), ("periodic", 1.1, 3, "SOR")])
def test_chain_independent_of_row_class_launches(tmp_path, family, omega, nmeas, smoother):
    """The colour passes of one row class in one launch (rowfuse.cuh: 9-colour radius-2, 8-colour 3d, 4-colour per-vertex levels)
    against one launch per colour: the same chain bit for bit, with and without dead-pass skipping."""

    def run(tag, env):
        out = str(tmp_path / f"{tag}.npy")
        e = dict(os.environ)
        e.update(env)
        subprocess.check_call([sys.executable, "-c", SNIPPET_ROWS % dict(root=ROOT, out=out, family=family, omega=omega, nmeas=nmeas, smoother=smoother)], env=e)
        return np.load(out)

    res = _parallel(run, {"rows": {}, "per_colour": {"MGMC_NO_ROWFUSE": "1"}, "rows_all_passes": {"MGMC_NO_DEAD_PASS": "1"},
                          "per_colour_all_passes": {"MGMC_NO_ROWFUSE": "1", "MGMC_NO_DEAD_PASS": "1"}})
    ref, x = res["rows"], res["per_colour"]
    assert np.all(np.isfinite(ref))
    assert np.array_equal(x, ref), f"per_colour: max abs diff {np.max(np.abs(x - ref)):.3e}"
    # Running the dead pass as well: these kernels update x_i += omega (b - (A x)_i) / a_ii with the own value inside (A x)_i, so for
    # omega = 1 the skipped pass changes the next update of the site in the last bits only (the tile kernel's omega = 1 passes do
    # not read the own value at all) -- the same chain to rounding; the two launch schemes stay bit-identical to each other.
    a, b = res["rows_all_passes"], res["per_colour_all_passes"]
    assert np.array_equal(a, b), f"all passes: max abs diff {np.max(np.abs(a - b)):.3e}"
    assert np.max(np.abs(a - ref)) <= 1e-9 * np.max(np.abs(ref))
