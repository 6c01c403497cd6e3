"""Committed golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py from the CPU oracle):
CPU: the oracle still reproduces them (regression pin of the restatement, incl. the reference's
lexicographic std::mt19937_64 chain); GPU: the CUDA path reproduces the deterministic kernels and the
colour-ordered Philox chain stored in them -- without rebuilding the expected values on the GPU box."""
import glob
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ALL = sorted(f for f in glob.glob(os.path.join(HERE, "golden", "*.npz")) if not os.path.basename(f).startswith("c5_"))
# oracle_only_*: FEM operator / 3d lattices (round 2) -- regression pins of the oracle; the CUDA path of these families is
# compared with the live oracle in tests/test_fem_operator.py and tests/test_lattice3d.py
FILES = [f for f in ALL if not os.path.basename(f).startswith("oracle_only_")]


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def _operator(orc, g):
    op = orc.Operator.prior(tuple(int(v) for v in g["n"]), str(g["pde"]), Lambda=0.2)
    if int(g["n_meas"]):
        op = op.measured(g["locs"], g["var"], variance_scaling=1e-3)
    return op


def test_fixtures_exist():
    assert len(FILES) >= 3 and len(ALL) >= 6


@pytest.mark.parametrize("path", ALL, ids=[os.path.basename(p) for p in ALL])
def test_oracle_reproduces_golden(oracle, path):
    g = np.load(path)
    op = _operator(oracle, g)
    nlevel, x, f = int(g["nlevel"]), g["x"], g["f"]
    assert rel(op.apply(x), g["apply"]) < 1e-14
    H = oracle.Hierarchy(op, nlevel, oracle.LEX)
    assert rel(H.restrict(0, x), g["restrict"]) < 1e-14
    assert rel(H.prolongate_add(0, 0.7, g["xc"], x), g["prolongate_add"]) < 1e-14
    assert rel(H.smoother(0, "SSOR", 0.9, 1, 1).apply(f, x), g["ssor_lex"]) < 1e-13
    s = H.mgmc(rng=oracle.StdRng(5418513))
    xs = x
    for _ in range(2):
        xs = s.apply(f, xs)
    assert rel(xs, g["mgmc_lex_mt19937_2samples"]) < 1e-12  # the reference's own chain, draw for draw


@pytest.mark.gpu
@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(p) for p in FILES])
def test_gpu_reproduces_golden(oracle, path):
    import multigridmc_b200 as m

    g = np.load(path)
    n, nlevel, x, f = tuple(int(v) for v in g["n"]), int(g["nlevel"]), g["x"], g["f"]
    B = _operator(oracle, g).B() if int(g["n_meas"]) else None  # host-side assembly of B only
    ctx = m.Context(n[0], n[1], nlevel, Lambda=0.2, B=B, pde=str(g["pde"]), seed=5418513)
    assert rel(ctx.op_apply(0, x), g["apply"]) < 1e-12
    assert rel(ctx.restrict(0, x), g["restrict"]) < 1e-12
    assert rel(ctx.prolongate_add(0, 0.7, g["xc"], x), g["prolongate_add"]) < 1e-12
    assert rel(ctx.residual_restrict(0, f, x), g["residual_restrict"]) < 1e-11
    assert rel(ctx.smoother_apply(0, "SSOR", f, x, omega=0.9), g["ssor_col"]) < 1e-12
    assert rel(ctx.mgprec_apply(f), g["mgprec_col"]) < 1e-10
    b = oracle.StdRng(1482817).normal(len(x))
    _, hist, _, _ = ctx.loop_solve(b, rtol=1e-10, atol=1e300, maxiter=25)
    ref = g["history_col"]
    assert len(hist) == len(ref) and np.abs(hist - ref).max() < 1e-11 * ref[0]
    # convergence RATE agrees with the lexicographic reference ordering (the iterates cannot)
    lex = g["history_lex"]
    rate = lambda h: (h[min(len(h), 10) - 1] / h[0]) ** (1.0 / (min(len(h), 10) - 1))
    assert abs(rate(hist) - rate(lex)) < 0.15
    ctx.set_philox_position(0)
    xs = x
    for _ in range(2):
        xs = ctx.mgmc_apply(f, xs)
    assert rel(xs, g["mgmc_col_philox_2samples"]) < 1e-9


def test_c5_reference_chain_fixture(oracle):
    """tests/golden/c5_reference_chains.npz (reference chains of BASELINE config 5, made by make_c5_chains.py): the
    oracle still produces the first cycles of chain 0, and the recorded chains have the exact mean within their own
    error bars (f = A u with u = 1 at the observed vertex)."""
    from multigridmc_b200 import workloads as w

    g = np.load(os.path.join(HERE, "golden", "c5_reference_chains.npz"))
    n, nlevel = int(g["n"]), int(g["nlevel"])
    op = oracle.Operator.prior((n, n), "shiftedlaplace_fd", Lambda=0.2)
    H = oracle.Hierarchy(op, nlevel, oracle.LEX)
    s = H.mgmc(rng=oracle.StdRng(int(g["seeds"][0])))
    xs = np.arange(1, n) / n
    u = np.outer(np.sin(np.pi * xs), np.sin(np.pi * xs)).ravel()
    q = w.nearest_vertex(n, n, [0.5, 0.5])
    assert q == int(g["qoi_index"])
    b_obs = np.zeros(op.ndof)
    b_obs[q] = 1.0
    _, z = s.run(op.apply(u), np.zeros(op.ndof), b_obs, 3)
    assert np.abs(np.asarray(z) - g["first_cycles"]).max() < 1e-12
    series = g["series"]
    means = series.mean(axis=1)
    assert abs(means.mean() - u[q]) < 4 * means.std(ddof=1) / np.sqrt(len(means))
    assert all(oracle.tau_int(series[c], 20) < 2.0 for c in range(series.shape[0]))
