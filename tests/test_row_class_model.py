"""CPU property test of the claim behind csrc/rowfuse.cuh: for the colourings the device uses -- 9 colours
(i % 3) + 3 (j % 3) with a radius-2 stencil, 8 colours (i & 1) + 2 (j & 1) + 4 (k & 1) with a 3d radius-1 stencil, 4 colours
(i & 1) + 2 (j & 1) with a 2d radius-1 stencil -- the colour passes of one ROW CLASS may run row by row, each row doing all its
passes back to back before any other row starts (the most adversarial schedule of independent CTAs), and the result is bit for
bit the global sweep "one colour over the whole lattice at a time".  Numpy model, no GPU, no library call."""
import itertools

import numpy as np
import pytest


def _colour(family, idx):
    if family == "radius2":
        return (idx[0] % 3) + 3 * (idx[1] % 3)
    if family == "3d":
        return (idx[0] & 1) + 2 * (idx[1] & 1) + 4 * (idx[2] & 1)
    return (idx[0] & 1) + 2 * (idx[1] & 1)


def _setup(family, rng):
    if family == "radius2":
        shape, rad, ncol, per_row = (14, 11), 2, 9, 3
    elif family == "3d":
        shape, rad, ncol, per_row = (9, 7, 6), 1, 8, 2
    else:
        shape, rad, ncol, per_row = (13, 10), 1, 4, 2
    dim = len(shape)
    offs = [o for o in itertools.product(range(-rad, rad + 1), repeat=dim)]
    # per-site coefficients (position classes / per-vertex operators are special cases), diagonally dominant
    coef = {o: rng.standard_normal(shape) for o in offs}
    coef[(0,) * dim] = 40.0 + rng.random(shape)
    return shape, offs, coef, ncol, per_row


def _update(x, f, noise, coef, offs, shape, idx, omega):
    s = 0.0
    for o in offs:
        nb = tuple(a + b for a, b in zip(idx, o))
        if all(0 <= v < n for v, n in zip(nb, shape)):
            s += coef[o][idx] * x[nb]
    d = coef[(0,) * len(shape)][idx]
    x[idx] = x[idx] + omega * (f[idx] + noise[idx] - s) / d


def _passes(ncol, sweeps, omega):
    """(colour, sweep index) of the live passes: as emit_smoothing_r2 -- the last colour of a sweep is dead when omega = 1 and the
    next sweep runs the other way (here the model keeps the own-value term, so dead passes are simply kept: same pass list for both
    schedules)."""
    out = []
    for si, fwd in enumerate(sweeps):
        for cc in range(ncol):
            out.append((cc if fwd else ncol - 1 - cc, si))
    return out


@pytest.mark.parametrize("family", ["radius2", "3d", "4colour"])
@pytest.mark.parametrize("sweeps,omega", [((True, False), 1.0), ((True, False, True, False), 1.3), ((True, True), 0.8), ((False,), 1.0)])
def test_row_class_schedule_equals_global_colour_sweep(family, sweeps, omega):
    rng = np.random.default_rng(11)
    shape, offs, coef, ncol, per_row = _setup(family, rng)
    f = rng.standard_normal(shape)
    x0 = rng.standard_normal(shape)
    noise = [rng.standard_normal(shape) for _ in sweeps]  # a pure function of (site, sweep), like the Philox stream
    sites = list(itertools.product(*[range(n) for n in shape]))
    passes = _passes(ncol, sweeps, omega)
    # (a) global: one colour over the whole lattice at a time
    xa = x0.copy()
    for colour, si in passes:
        for idx in sites:
            if _colour(family, idx) == colour:
                _update(xa, f, noise[si], coef, offs, shape, idx, omega)
    # (b) row classes: consecutive passes of one class form a launch; inside a launch every row runs ALL its passes before the next row
    xb = x0.copy()
    groups, cur = [], []
    for colour, si in passes:
        if cur and colour // per_row != cur[0][0] // per_row:
            groups.append(cur)
            cur = []
        cur.append((colour, si))
    groups.append(cur)
    assert max(len(g) for g in groups) <= 6  # kRowPassMax
    rows = sorted({idx[1:] for idx in sites})
    for g in groups:
        for row in reversed(rows):  # (any row order)
            for colour, si in g:
                for i in range(shape[0]):
                    idx = (i,) + row
                    if _colour(family, idx) == colour:
                        _update(xb, f, noise[si], coef, offs, shape, idx, omega)
    assert np.array_equal(xa, xb)
    # the launch counts DESIGN.md 4.9 quotes for an SSOR step (without the dead pass: one pass less, same launches)
    if sweeps == (True, False):
        assert len(groups) == {"radius2": 5, "3d": 7, "4colour": 3}[family]
