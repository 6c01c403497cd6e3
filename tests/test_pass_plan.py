"""Pass planning of the fused tile kernel (plan_stages, mgmc_b200.cu) checked on the CPU: a numpy model
of ONE tile that loads the planned input halo, runs every pass only inside its planned rectangle and skips the passes
the plan declares dead must reproduce the global multicolour sweep on the tile (+ the halo a fused residual needs),
bit for bit -- for red-black / 4-colour orderings, omega = 1 and omega != 1, with and without restriction."""
import numpy as np
import pytest

FULL, SKIP, SPARSE = 0, 1, 2


def colour_of(nc, i, j):
    return (i + j) & 1 if nc == 2 else ((j & 1) << 1) | (i & 1)


def sweep_pass(x, f, nc, colour, omega, w, rect, rng_noise):
    """one colour pass on the rectangle rect = (i0, i1, j0, j1) (inclusive) of the padded arrays (in place)"""
    i0, i1, j0, j1 = rect
    jj, ii = np.mgrid[j0:j1 + 1, i0:i1 + 1]
    mask = colour_of(nc, ii, jj) == colour
    xs = x[j0 - 1:j1 + 2, i0 - 1:i1 + 2]
    s = np.zeros_like(xs[1:-1, 1:-1])
    for dj in (-1, 0, 1):
        for di in (-1, 0, 1):
            if (dj or di) and (nc == 4 or dj == 0 or di == 0):
                s = s + w[dj + 1, di + 1] * xs[1 + dj:xs.shape[0] - 1 + dj, 1 + di:xs.shape[1] - 1 + di]
    b = f[j0:j1 + 1, i0:i1 + 1] + rng_noise[j0:j1 + 1, i0:i1 + 1]
    old = x[j0:j1 + 1, i0:i1 + 1]
    new = old + omega * (b - s - w[1, 1] * old) / w[1, 1] if omega != 1.0 else (b - s) / w[1, 1]
    x[j0:j1 + 1, i0:i1 + 1] = np.where(mask, new, old)


@pytest.mark.parametrize("nc,seq", [(2, [0, 1, 1, 0]), (2, [0, 1]), (2, [0, 1, 1, 0, 0, 1, 1, 0]), (4, [0, 1, 2, 3, 3, 2, 1, 0]),
                                    (4, [3, 2, 1, 0]), (4, [0, 1, 2, 3, 0, 1, 2, 3]),
                                    (4, [0, 1, 2, 3, 3, 2, 1, 0, 0, 1, 2, 3, 3, 2, 1, 0])])  # (16 passes: a V(2,2) smoothing step of a small level in one launch)
@pytest.mark.parametrize("omega", [1.0, 1.3])
@pytest.mark.parametrize("restrict_behind", [False, True])
def test_planned_tile_reproduces_global_sweep(nc, seq, omega, restrict_behind):
    from multigridmc_b200 import capi

    modes, margins, halo = capi.plan_passes(nc, seq, omega_is_one=(omega == 1.0), restrict_behind=restrict_behind)
    assert all(m in (FULL, SKIP) for m in modes)
    if omega == 1.0 and len(seq) >= 2 and any(a == b for a, b in zip(seq, seq[1:])):
        assert SKIP in modes  # the first of two consecutive passes of one colour is dead
    if omega != 1.0:
        assert SKIP not in modes
    rng = np.random.default_rng(len(seq) * 7 + nc)
    n = 64
    w = rng.uniform(-1.0, -0.2, (3, 3))
    w[1, 1] = 9.0
    x0, f = rng.standard_normal((n, n)), rng.standard_normal((n, n))
    noise = [rng.standard_normal((n, n)) for _ in seq]
    # global sweep on the interior [1, n - 2]^2 (the outer ring plays the role of fixed boundary values)
    xg = x0.copy()
    for s, c in enumerate(seq):
        sweep_pass(xg, f, nc, c, omega, w, (1, n - 2, 1, n - 2), noise[s])
    # one tile in the middle of the lattice
    ti0, ti1, tj0, tj1 = 24, 39, 25, 40
    xt = np.full((n, n), np.nan)
    hxl, hxh, hyl, hyh = halo
    xt[tj0 - hyl:tj1 + hyh + 1, ti0 - hxl:ti1 + hxh + 1] = x0[tj0 - hyl:tj1 + hyh + 1, ti0 - hxl:ti1 + hxh + 1]
    for s, c in enumerate(seq):
        if modes[s] == SKIP:
            continue
        xl, xh, yl, yh = margins[s]
        sweep_pass(xt, f, nc, c, omega, w, (ti0 - xl, ti1 + xh, tj0 - yl, tj1 + yh), noise[s])
    ex = (2, 1, 1, 2) if restrict_behind else (0, 0, 0, 0)  # what the fused residual / restriction reads beyond the tile
    sl = (slice(tj0 - ex[2], tj1 + ex[3] + 1), slice(ti0 - ex[0], ti1 + ex[1] + 1))
    assert not np.isnan(xt[sl]).any()
    assert np.array_equal(xt[sl], xg[sl])


def test_plan_with_fixups_keeps_measurement_sites_alive():
    from multigridmc_b200 import capi

    modes, margins, halo = capi.plan_passes(2, [0, 1, 1, 0], fix_after=[1, 3], lr_mx=2, lr_my=1, restrict_behind=True)
    assert modes == [FULL, SPARSE, FULL, FULL]          # the dead pass is still run on supp(B_k) of the owned measurements
    assert halo[1] >= 2 + 1 and halo[3] >= 1 + 1        # supp(B_k) (+ its neighbours) lies inside the loaded region
    modes4, _, halo4 = capi.plan_passes(4, [0, 1, 2, 3, 3, 2, 1, 0])
    assert modes4.count(SKIP) == 1 and halo4 == [7, 7, 3, 3]  # directional margins: y-halo 3 rows for 8 passes


def test_plan_of_the_merged_level0_launch():
    """Post-smoothing of cycle k + pre-smoothing of cycle k + 1 in one launch (emit_merged_level0, mgmc_b200.cu): red-black
    SSOR V(1,1) is R B | B R || R B | B R with a low-rank fix-up behind every sweep; with omega = 1 three passes are dead
    (run on the measured / observed sites only), and the 5 live passes + the fused residual need 13 halo rows."""
    from multigridmc_b200 import capi

    modes, margins, halo = capi.plan_passes(2, [0, 1, 1, 0, 0, 1, 1, 0], fix_after=[1, 3, 5, 7], restrict_behind=True)
    assert modes == [FULL, SPARSE, FULL, SPARSE, FULL, SPARSE, FULL, FULL]
    assert halo == [7, 6, 6, 7]
    live = [mg for md, mg in zip(modes, margins) if md == FULL]
    assert all(a[k] >= b[k] for a, b in zip(live, live[1:]) for k in range(4))  # the rectangles shrink towards the tile
    # without measurements the dead passes are skipped altogether
    assert capi.plan_passes(2, [0, 1, 1, 0, 0, 1, 1, 0], restrict_behind=True)[0] == [FULL, SKIP, FULL, SKIP, FULL, SKIP, FULL, FULL]
