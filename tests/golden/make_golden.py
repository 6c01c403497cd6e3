#!/usr/bin/env python
"""Generates tests/golden/*.npz from the CPU oracle (oracle/, the restatement of the reference).

The reference repository holds no golden vectors of its own and cannot be built in
this image (Eigen / libconfig++ absent), so these fixtures pin the ORACLE: lexicographic reference ordering +
std::mt19937_64 chain (what the reference computes) and the colour-ordered / Philox twin the GPU path is
compared with.  Regenerate with `python tests/golden/make_golden.py`; tests/test_golden.py checks both the
oracle (CPU) and the CUDA path (GPU) against the committed files.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import oracle as orc  # noqa: E402


def case(n, nlevel, pde, n_meas, seed):
    rng = np.random.default_rng(seed)
    op = orc.Operator.prior(n, pde, Lambda=0.2)
    locs = var = None
    if n_meas:
        locs = 0.1 + 0.8 * rng.random((n_meas, len(n)))
        var = 1.0 + rng.random(n_meas)
        op = op.measured(locs, var, variance_scaling=1e-3)
    nd = op.ndof
    x, f = rng.standard_normal(nd), rng.standard_normal(nd)
    out = dict(n=np.array(n), nlevel=nlevel, pde=pde, n_meas=n_meas, x=x, f=f)
    if n_meas:
        out.update(locs=locs, var=var)
    out["apply"] = op.apply(x)
    for name, order in (("lex", orc.LEX), ("col", orc.COLOUR)):
        H = orc.Hierarchy(op, nlevel, order)
        if name == "lex":
            out["restrict"] = H.restrict(0, x)
            xc = rng.standard_normal(H.level_op(1).ndof)
            out["xc"] = xc
            out["prolongate_add"] = H.prolongate_add(0, 0.7, xc, x)
            out["residual_restrict"] = H.restrict(0, f - op.apply(x))
        out[f"ssor_{name}"] = H.smoother(0, "SSOR", 0.9, 1, 1).apply(f, x)
        prec = H.preconditioner(npresmooth=1, npostsmooth=1)
        out[f"mgprec_{name}"] = prec.apply(f, np.zeros(nd))
        b = orc.StdRng(1482817).normal(nd)
        _, hist, _, _ = orc.loop_solve(op, prec, b, rtol=1e-10, atol=1e300, maxiter=25)
        out[f"history_{name}"] = hist
    # the reference chain: lexicographic sweeps, one std::mt19937_64(5418513) shared by all samplers
    H = orc.Hierarchy(op, nlevel, orc.LEX)
    s = H.mgmc(rng=orc.StdRng(5418513))
    xs = x
    for _ in range(2):
        xs = s.apply(f, xs)
    out["mgmc_lex_mt19937_2samples"] = xs
    # its colour-ordered Philox twin (what the GPU path computes)
    H = orc.Hierarchy(op, nlevel, orc.COLOUR)
    s = H.mgmc(rng=None, philox_seed=5418513)
    xs = x
    for _ in range(2):
        xs = s.apply(f, xs)
    out["mgmc_col_philox_2samples"] = xs
    return out


if __name__ == "__main__":
    np.savez_compressed(os.path.join(HERE, "laplace_32x32_l3_m0.npz"), **case((32, 32), 3, "shiftedlaplace_fd", 0, 1))
    np.savez_compressed(os.path.join(HERE, "laplace_64x32_l3_m3.npz"), **case((64, 32), 3, "shiftedlaplace_fd", 3, 2))
    np.savez_compressed(os.path.join(HERE, "squared_32x32_l2_m0.npz"), **case((32, 32), 2, "squared_shiftedlaplace_fd", 0, 3))
    # round 2: FEM operator and 3d lattices (the CUDA path of these families is compared with the live oracle in
    # tests/test_fem_operator.py / tests/test_lattice3d.py; these files pin the oracle's restatement of them)
    np.savez_compressed(os.path.join(HERE, "oracle_only_fem_32x16_l2_m2.npz"), **case((32, 16), 2, "shiftedlaplace_fem", 2, 4))
    np.savez_compressed(os.path.join(HERE, "oracle_only_laplace_16x16x16_l3_m2.npz"), **case((16, 16, 16), 3, "shiftedlaplace_fd", 2, 5))
    np.savez_compressed(os.path.join(HERE, "oracle_only_fem_8x16x8_l2_m0.npz"), **case((8, 16, 8), 2, "shiftedlaplace_fem", 0, 6))
    print("written")
