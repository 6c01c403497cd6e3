"""ShiftedLaplaceFEMOperator (linear_operator/shiftedlaplace_fem_operator.cc:9-150; the operator of the reference's own MGMC sampler
test, test_sampler.hh:260-323) with a constant correlation length, 2d and 3d: the fine level already carries a uniform 9-point /
27-point stencil.  CPU: the stencil algebra against the oracle's assembled matrix and its R A R^T.  GPU: every C-ABI entry point
against the oracle on the same inputs (same orderings, same Philox stream), as tests/test_gpu_parity.py does for the FD operator.
"""
import numpy as np
import pytest

import multigridmc_b200 as m
from multigridmc_b200 import capi

from .test_host_setup import _stencil_matrix
from .test_lattice3d import _stencil_matrix3, rel

TOL = 1e-12
PDE = "shiftedlaplace_fem"


@pytest.mark.parametrize("n,nlevel", [((16, 16), 3), ((32, 16), 2), ((8, 8, 8), 2), ((16, 8, 12), 2)])
def test_fem_stencils_match_oracle(oracle, n, nlevel):
    op = oracle.Operator.prior(n, PDE, Lambda=0.2)
    H = oracle.Hierarchy(op, nlevel, oracle.COLOUR)
    d3 = len(n) == 3
    desc = capi.make_desc(n[0], n[1], nlevel, pde=PDE, Lambda=0.2, nz=n[2] if d3 else None)
    shape = list(n)
    for level in range(nlevel):
        st, nc = m.host_stencil3(desc, level) if d3 else m.host_stencil(desc, level)
        A = _stencil_matrix3(st, *shape) if d3 else _stencil_matrix(st, *shape)
        A_ref = H.level_op(level).csr().toarray()
        assert np.abs(A - A_ref).max() <= 1e-12 * np.abs(A_ref).max(), f"level {level}"
        assert nc == H.ncolours(level) == (8 if d3 else 4)
        shape = [v // 2 for v in shape]


def test_fem_with_variable_correlation_length_is_reported_unsupported():
    desc = capi.make_desc(16, 16, 2, pde=PDE, kappa_sq=np.full(15 * 15, 25.0))
    with pytest.raises(m.MgmcError) as e:
        m.host_stencil(desc, 0)
    assert e.value.code == -2


def _setup(oracle, n, nlevel, n_meas=0, **kw):
    op = oracle.Operator.prior(n, PDE, Lambda=0.2)
    if n_meas:
        rng = np.random.default_rng(7)
        locs = 0.15 + 0.7 * rng.random((n_meas, len(n)))
        op = op.measured(locs, 1.0 + rng.random(n_meas), variance_scaling=1e-4, radius=0.0)
    H = oracle.Hierarchy(op, nlevel, oracle.COLOUR)
    ctx = m.Context(n[0], n[1], nlevel, nz=n[2] if len(n) == 3 else None, pde=PDE, Lambda=0.2, B=op.B() if n_meas else None, **kw)
    return op, H, ctx


@pytest.mark.gpu
@pytest.mark.parametrize("n,nlevel,n_meas", [((64, 64), 3, 0), ((96, 32), 3, 0), ((128, 128), 4, 5), ((16, 16, 16), 3, 0), ((32, 16, 16), 2, 3)])
def test_single_level_operations_fem(oracle, n, nlevel, n_meas):
    seed = 4711
    op, H, ctx = _setup(oracle, n, nlevel, n_meas, seed=seed)
    rng = np.random.default_rng(1)
    for level in range(nlevel):
        lop = H.level_op(level)
        nd = lop.ndof
        assert ctx.ndof(level) == nd
        assert ctx.level_info(level)[3] == H.ncolours(level)
        x, b = rng.standard_normal(nd), rng.standard_normal(nd)
        assert rel(ctx.op_apply(level, x), lop.apply(x)) < TOL
        for kind, direction, nsmooth, omega in (("SOR", 1, 1, 1.0), ("SOR", 2, 1, 0.8), ("SSOR", 1, 2, 1.0), ("SSOR", 1, 1, 0.9)):
            ref = H.smoother(level, kind, omega, nsmooth, direction).apply(b, x)
            assert rel(ctx.smoother_apply(level, kind, b, x, omega=omega, nsmooth=nsmooth, direction=direction), ref) < TOL, (level, kind, omega)
            s = H.sampler(level, kind, omega=omega, nsmooth=nsmooth, direction=direction, rng=None, philox_seed=seed)
            s.set_philox_position(3, 0, 1)
            ctx.set_philox_position(3, 1)
            got = ctx.sampler_apply(level, kind, b, x, omega=omega, nsmooth=nsmooth, direction=direction)
            assert rel(got, s.apply(b, x)) < 1e-11, (level, kind, omega)
        if level < nlevel - 1:
            xc = rng.standard_normal(H.level_op(level + 1).ndof)
            assert rel(ctx.residual_restrict(level, b, x), H.restrict(level, b - lop.apply(x))) < TOL
            assert rel(ctx.prolongate_add(level, 0.7, xc, x), H.prolongate_add(level, 0.7, xc, x)) < TOL
    x_exact = rng.standard_normal(op.ndof)
    assert rel(ctx.smoother_apply(0, "SSOR", op.apply(x_exact), x_exact, omega=0.8), x_exact) < (1e-9 if n_meas else TOL)


@pytest.mark.gpu
@pytest.mark.parametrize("n,nlevel,n_meas,kw", [
    ((64, 64), 3, 0, {}),
    ((16, 8), 2, 4, {}),  # (the lattice of the reference's MGMC sampler test, test_sampler.hh:260-323: 16 x 8 FEM + 4 measurements)
    ((128, 128), 4, 5, dict(npresmooth=2, npostsmooth=2)),
    ((64, 128), 4, 0, dict(smoother="SOR", cycle=2, npresmooth=2, omega=0.9)),
    ((32, 32, 32), 3, 4, {}),
])
def test_multigrid_solver_and_mgmc_chain_fem(oracle, n, nlevel, n_meas, kw):
    seed = 5418513
    op, H, ctx = _setup(oracle, n, nlevel, n_meas, seed=seed, **kw)
    b = oracle.StdRng(1482817).normal(op.ndof)
    prec = H.preconditioner(**kw)
    assert rel(ctx.mgprec_apply(b), prec.apply(b, np.zeros_like(b))) < 1e-11
    x_ref, h_ref, it_ref, cv_ref = oracle.loop_solve(op, prec, b, rtol=1e-12, atol=1e-15, maxiter=20)
    x, h, it, cv = ctx.loop_solve(b, rtol=1e-12, atol=1e-15, maxiter=20)
    assert len(h) == len(h_ref) and it == it_ref and cv == cv_ref
    assert np.abs(h - h_ref).max() < 1e-12 * np.linalg.norm(b)
    assert rel(x, x_ref) < 1e-11
    rng = np.random.default_rng(8)
    f, x0 = rng.standard_normal(op.ndof), rng.standard_normal(op.ndof)
    sampler = H.mgmc(rng=None, philox_seed=seed, **kw)
    ctx.set_philox_position(0)
    xr, xg = x0, x0
    for k in range(3):
        xr = sampler.apply(f, xr)
        xg = ctx.mgmc_apply(f, xg)
        assert rel(xg, xr) < 1e-10, k
    idx = np.array([op.ndof // 2 + 3, 5])
    val = np.array([1.0, -0.5])
    ctx.set_qoi(idx, val)
    ctx.set_rhs(f)
    ctx.set_state(xg)
    series = ctx.sample(4)[:, 0]
    b_obs = np.zeros(op.ndof)
    b_obs[idx] = val
    xr2, series_ref = sampler.run(f, xr, b_obs, 4)
    assert rel(ctx.get_state(), xr2) < 1e-10
    assert np.abs(series - series_ref).max() < 1e-10 * np.abs(series_ref).max()


@pytest.mark.gpu
def test_fem_mgmc_sampler_covariance_like_the_reference_test(oracle):
    """test_sampler.hh:260-323 (MultigridMC on the 16 x 8 FEM lattice with 4 measurements: sample mean and covariance against the exact
    posterior, 2e-3 at 5e5 samples in the reference): here 64 chains x 4000 samples of the QoI set {x_k} at 6 vertices -- mean against
    A^-1 f and covariance against A^-1 (dense inverse in the oracle) within 5 standard errors."""
    n, nlevel, nchains, nsamples = (16, 8), 2, 64, 4000
    op, H, ctx0 = _setup(oracle, n, nlevel, 4)
    ctx0.close()
    cov = op.covariance()  # (A_0 + B Sigma^-1 B^T)^-1, dense (linear_operator.hh:153-174)
    rng = np.random.default_rng(3)
    f = rng.standard_normal(op.ndof)
    mean = cov @ f
    sites = [3, 17, 40, 52, 77, 100]
    ctx = m.Context(n[0], n[1], nlevel, pde=PDE, Lambda=0.2, B=op.B(), nchains=nchains, seed=99)
    ctx.set_rhs(np.tile(f, nchains))
    ctx.set_state(np.zeros(op.ndof * nchains))
    ctx.sample(50, series=False)
    acc1, acc2 = np.zeros(len(sites)), np.zeros((len(sites), len(sites)))
    for _ in range(nsamples // 40):
        ctx.sample(40, series=False)  # thinned: one state per 40 cycles and chain
        X = ctx.get_state().reshape(nchains, op.ndof)[:, sites]
        acc1 += X.sum(axis=0)
        acc2 += X.T @ X
    N = nchains * (nsamples // 40)
    mu = acc1 / N
    C = acc2 / N - np.outer(mu, mu)
    Cx = cov[np.ix_(sites, sites)]
    se_mu = np.sqrt(np.diag(Cx) / N)
    assert np.all(np.abs(mu - mean[sites]) < 5 * se_mu)
    se_C = np.sqrt((Cx ** 2 + np.outer(np.diag(Cx), np.diag(Cx))) / N)
    assert np.all(np.abs(C - Cx) < 5 * se_C)


@pytest.mark.parametrize("n", [(8, 8), (16, 12), (8, 8, 8), (8, 12, 16)])
def test_coarsened_fem_operator_is_the_rediscretised_one(n):
    """test_intergrid.hh:172-207 (TestCoarsenOperator2d / 3d, Lambda = 1): coarsening the FEM operator with constant coefficients
    gives the FEM operator of the next-coarser lattice, 1e-12 -- here for the product's stencil algebra (setup.hh), no oracle."""
    d3 = len(n) == 3
    fine = capi.make_desc(n[0], n[1], 2, pde=PDE, Lambda=1.0, nz=n[2] if d3 else None)
    coarse = capi.make_desc(n[0] // 2, n[1] // 2, 1, pde=PDE, Lambda=1.0, nz=n[2] // 2 if d3 else None)
    get = m.host_stencil3 if d3 else m.host_stencil
    a, _ = get(fine, 1)
    b, _ = get(coarse, 0)
    interior = a if d3 else a[4]
    ref = b if d3 else b[4]
    assert np.abs(interior - ref).max() < 1e-12 * np.abs(ref).max()
