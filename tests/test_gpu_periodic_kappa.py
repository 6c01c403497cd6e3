"""GPU parity for a correlation length that varies in space: PeriodicCorrelationLengthModel
(linear_operator/correlationlength_model.hh:83-113), the model every sampler / solver / smoother test of the reference
uses (test_sampler.hh:286, test_solver.hh:41).  The operators of all levels carry per-vertex coefficients
(csrc/varcoef.cuh); every C-ABI entry point is compared with the CPU oracle on the same inputs, exactly as
tests/test_gpu_parity.py does for the constant model: 1e-12 relative for the deterministic kernels, the oracle fed by
the same Philox stream for the samplers.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = 1e-12
LMIN, LMAX = 0.1, 0.4


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.fixture(scope="module")
def m():
    import multigridmc_b200 as mod

    mod.lib()
    return mod


def _setup(oracle, m, n, nlevel, n_meas=0, radius=0.0, **kw):
    op = oracle.Operator.prior(n, "shiftedlaplace_fd", Lambda_min=LMIN, Lambda_max=LMAX)
    if n_meas:
        rng = np.random.default_rng(7)
        locs = 0.1 + 0.8 * rng.random((n_meas, 2))
        op = op.measured(locs, 1.0 + rng.random(n_meas), variance_scaling=1e-6, radius=radius)
    H = oracle.Hierarchy(op, nlevel, oracle.COLOUR)
    ctx = m.Context(n[0], n[1], nlevel, B=op.B() if n_meas else None, kappa_sq=m.periodic_kappa_sq(n[0], n[1], LMIN, LMAX), **kw)
    return op, H, ctx


# (last case: the coarse lattice has a single interior column -- a 9-plane operator without corner couplings, swept red-black)
CASES = [((64, 64), 3, 0), ((96, 32), 3, 0), ((128, 128), 4, 5), ((48, 80), 2, 3), ((4, 64), 2, 0)]


@pytest.mark.parametrize("n,nlevel,n_meas", CASES)
def test_single_level_operations(oracle, m, n, nlevel, n_meas):
    seed = 4711
    op, H, ctx = _setup(oracle, m, n, nlevel, n_meas, seed=seed)
    rng = np.random.default_rng(1)
    for level in range(nlevel):
        lop = H.level_op(level)
        nd = lop.ndof
        assert ctx.ndof(level) == nd
        assert ctx.level_info(level)[3] == H.ncolours(level) == (2 if level == 0 or n[0] == 4 else 4)
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
            ncoarse = H.level_op(level + 1).ndof
            xc = rng.standard_normal(ncoarse)
            assert rel(ctx.residual_restrict(level, b, x), H.restrict(level, b - lop.apply(x))) < TOL
            assert rel(ctx.prolongate_add(level, 0.7, xc, x), H.prolongate_add(level, 0.7, xc, x)) < TOL
    # SSOR leaves the exact solution invariant (test_smoother.hh:90-114, which runs on this correlation length model)
    x_exact = rng.standard_normal(op.ndof)
    assert rel(ctx.smoother_apply(0, "SSOR", op.apply(x_exact), x_exact, omega=0.8), x_exact) < (1e-10 if n_meas else TOL)
    # coarsest level: dense factor of the per-vertex operator
    lc = nlevel - 1
    nd = H.level_op(lc).ndof
    b = rng.standard_normal(nd)
    assert rel(ctx.coarse_solve(b), H.cholesky_solver(lc).apply(b, np.zeros(nd))) < 1e-10
    s = H.sampler(lc, "Cholesky", rng=None, philox_seed=seed)
    s.set_philox_position(5, 0, 2)
    ctx.set_philox_position(5, 2)
    assert rel(ctx.coarse_sample(b), s.apply(b, np.zeros(nd))) < 1e-10


@pytest.mark.parametrize("n,nlevel,n_meas,kw", [
    ((64, 64), 3, 0, {}),
    ((64, 64), 4, 0, dict(smoother="SOR", cycle=2, npresmooth=2, omega=0.9)),
    ((128, 128), 4, 5, dict(npresmooth=2, npostsmooth=2)),
    ((128, 64), 3, 4, dict(coarse_solver="SSOR", ncoarsesmooth=2)),
])
def test_multigrid_solver_and_mgmc_chain(oracle, m, n, nlevel, n_meas, kw):
    """MultigridPreconditioner + LoopSolver (test_solver.hh:41-90 runs them on the periodic model) and three MGMC samples
    + the graph-replayed device loop against the oracle chain on the same Philox stream."""
    seed = 5418513
    op, H, ctx = _setup(oracle, m, n, nlevel, n_meas, seed=seed, **kw)
    b = oracle.StdRng(1482817).normal(op.ndof)
    prec = H.preconditioner(**kw)
    assert rel(ctx.mgprec_apply(b), prec.apply(b, np.zeros_like(b))) < 1e-11
    x_ref, h_ref, it_ref, cv_ref = oracle.loop_solve(op, prec, b, rtol=1e-12, atol=1e-15, maxiter=30)
    x, h, it, cv = ctx.loop_solve(b, rtol=1e-12, atol=1e-15, maxiter=30)
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


def test_constant_kappa_field_reproduces_the_stencil_path(m):
    """kappa^2 = 1 / Lambda^2 at every vertex: the per-vertex kernels and the tile kernels run the same orderings and the
    same noise streams, so the chains agree to rounding (different summation order only)."""
    n, nlevel, seed = 128, 4, 99
    rng = np.random.default_rng(3)
    f, x0 = rng.standard_normal((n - 1) ** 2), rng.standard_normal((n - 1) ** 2)
    out = []
    for kw in (dict(Lambda=0.2), dict(kappa_sq=np.full((n - 1) ** 2, 1.0 / 0.2 ** 2))):
        ctx = m.Context(n, n, nlevel, seed=seed, **kw)
        ctx.set_philox_position(0)
        x = x0
        for _ in range(3):
            x = ctx.mgmc_apply(f, x)
        out.append(x)
        ctx.close()
    assert rel(out[1], out[0]) < 1e-11


def test_many_chains_and_large_lattice_properties(m):
    """1024 x 1024, 6 levels (size-independent identities, no oracle): the operator is symmetric, SSOR keeps the exact
    solution, the multigrid-preconditioned Richardson iteration converges to it; batched chains equal single chains."""
    n, nlevel = 1024, 6
    ks = m.periodic_kappa_sq(n, n, LMIN, LMAX)
    ctx = m.Context(n, n, nlevel, kappa_sq=ks, npresmooth=2, npostsmooth=2)
    rng = np.random.default_rng(11)
    nd = ctx.ndof(0)
    x, y = rng.standard_normal(nd), rng.standard_normal(nd)
    Ax, Ay = ctx.op_apply(0, x), ctx.op_apply(0, y)
    assert abs(y.dot(Ax) - x.dot(Ay)) < 1e-11 * abs(y.dot(Ax))
    assert rel(ctx.smoother_apply(0, "SSOR", Ax, x, omega=0.8), x) < 1e-11
    xs, h, it, cv = ctx.loop_solve(Ax, rtol=1e-11, atol=1e300, maxiter=40)
    assert cv and it <= 20
    assert rel(xs, x) < 1e-8
    ctx.close()
    n, nlevel, seed = 128, 4, 5
    ks = m.periodic_kappa_sq(n, n, LMIN, LMAX)
    nd = (n - 1) ** 2
    f = rng.standard_normal(nd)
    batch = m.Context(n, n, nlevel, kappa_sq=ks, seed=seed, nchains=3)
    batch.set_rhs(np.tile(f, 3))
    batch.set_state(np.zeros(3 * nd))
    batch.sample(3, series=False)
    xb = batch.get_state().reshape(3, nd)
    for ch in range(3):
        single = m.Context(n, n, nlevel, kappa_sq=ks, seed=seed, first_chain=ch)
        single.set_rhs(f)
        single.set_state(np.zeros(nd))
        single.sample(3, series=False)
        assert np.array_equal(single.get_state(), xb[ch]), ch
        single.close()
    assert rel(xb[0], xb[1]) > 1e-3


def test_statistics_agree_with_reference_chain_periodic_and_global(oracle, m):
    """North-star check (SURVEY.md section 8c procedure) on the operator family of the reference's own sampler tests
    (test_sampler.hh:286: periodic correlation length) with point measurements AND the global measurement: the coloured /
    Philox GPU chains and the reference's lexicographic / std::mt19937_64 chain (oracle in reference ordering and RNG call
    order) sample the same law -- QoI mean, variance, integrated autocorrelation time within Monte-Carlo error bars -- and both
    agree with the exact posterior mean / variance of the observation (linear_operator.hh:153-174, dense solve in the oracle)."""
    n, nlevel, nchains, nsamples, nref = (32, 32), 3, 16, 3000, 12000
    prior = oracle.Operator.prior(n, "shiftedlaplace_fd", Lambda_min=LMIN, Lambda_max=LMAX)
    rng = np.random.default_rng(12)
    locs, var, y = 0.15 + 0.7 * rng.random((5, 2)), 1.0 + rng.random(5), 1.0 + 3.0 * rng.random(6)
    op = prior.measured(locs, var, variance_scaling=1e-2, measure_global=True, variance_global=1e-2)
    assert op.m_lowrank == 6
    xbar = op.mean(np.zeros(op.ndof), y)
    f = op.apply(xbar)
    b_obs = op.measurement_vector([0.4, 0.6], 0.0)
    # reference chain on the CPU
    H = oracle.Hierarchy(op, nlevel, oracle.LEX)
    s = H.mgmc(rng=oracle.StdRng(5418513))
    x, _ = s.run(f, np.zeros(op.ndof), b_obs, 200)
    _, z_ref = s.run(f, x, b_obs, nref)
    tau_ref = oracle.tau_int(z_ref, 20)
    se_mean_ref = np.sqrt(z_ref.var(ddof=1) * max(tau_ref, 1.0) / nref)
    se_var_ref = z_ref.var(ddof=1) * np.sqrt(2.0 * max(tau_ref, 1.0) / nref)
    # GPU chains
    ctx = m.Context(n[0], n[1], nlevel, B=op.B(), kappa_sq=m.periodic_kappa_sq(n[0], n[1], LMIN, LMAX), nchains=nchains, seed=777)
    idx = np.nonzero(b_obs)[0]
    ctx.set_qoi(idx, b_obs[idx])
    ctx.set_rhs(np.tile(f, nchains))
    ctx.set_state(np.zeros(op.ndof * nchains))
    ctx.sample(200, series=False)
    z = ctx.sample(nsamples)
    means, variances = z.mean(axis=0), z.var(axis=0, ddof=1)
    se_mean = means.std(ddof=1) / np.sqrt(nchains)
    se_var = variances.std(ddof=1) / np.sqrt(nchains)
    tau = np.array([oracle.tau_int(z[:, c], 20) for c in range(nchains)])
    assert abs(means.mean() - z_ref.mean()) < 4 * np.hypot(se_mean, se_mean_ref)
    assert abs(variances.mean() - z_ref.var(ddof=1)) < 4 * np.hypot(se_var, se_var_ref)
    assert tau.mean() <= tau_ref + 3 * (tau.std(ddof=1) / np.sqrt(nchains) + 0.15)
    assert tau.mean() < 2.0 and tau_ref < 2.0
    # exact posterior mean of the observation: b^T xbar
    assert abs(means.mean() - b_obs.dot(xbar)) < 4 * se_mean
