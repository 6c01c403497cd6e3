"""GPU parity tests: every C-ABI entry point against the CPU oracle on the same seeded inputs.

Tolerance: 1e-12 relative in fp64 (BASELINE.json north_star) for the deterministic kernels.  Sweeps are
compared with the oracle run in the SAME multicolour ordering (a coloured sweep is
a different splitting than the reference's lexicographic one, so only rates -- not iterates -- can
match the lexicographic chain); Gibbs sweeps are compared with the oracle fed by the same Philox
stream, which turns the sampler into a deterministic map that must agree to rounding.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = 1e-12


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def _measurements(oracle, prior, n_meas, seed=7, radius=0.0):
    rng = np.random.default_rng(seed)
    locs = 0.1 + 0.8 * rng.random((n_meas, 2))
    var = 1.0 + rng.random(n_meas)
    return prior.measured(locs, var, variance_scaling=1e-6, radius=radius)


def _setup(oracle, m, n, nlevel, n_meas=0, radius=0.0, pde="shiftedlaplace_fd", **kw):
    """oracle hierarchy (colour ordering) + GPU context for the same problem"""
    op = oracle.Operator.prior(n, pde, Lambda=0.2)
    if n_meas:
        op = _measurements(oracle, op, n_meas, radius=radius)
    H = oracle.Hierarchy(op, nlevel, oracle.COLOUR)
    B = op.B() if n_meas else None
    ctx = m.Context(n[0], n[1], nlevel, Lambda=0.2, B=B, pde=pde, **kw)
    return op, H, ctx


@pytest.fixture(scope="module")
def m():
    import multigridmc_b200 as mod

    mod.lib()
    return mod


CASES = [((64, 64), 3, 0), ((64, 32), 3, 0), ((128, 128), 4, 5), ((48, 80), 2, 3)]


@pytest.mark.parametrize("n,nlevel,n_meas", CASES)
def test_operator_apply_all_levels(oracle, m, n, nlevel, n_meas):
    op, H, ctx = _setup(oracle, m, n, nlevel, n_meas)
    rng = np.random.default_rng(1)
    for level in range(nlevel):
        lop = H.level_op(level)
        assert ctx.ndof(level) == lop.ndof
        assert ctx.level_info(level)[3] == H.ncolours(level)
        x = rng.standard_normal(lop.ndof)
        assert rel(ctx.op_apply(level, x), lop.apply(x)) < TOL


@pytest.mark.parametrize("n,nlevel,n_meas", CASES)
def test_intergrid(oracle, m, n, nlevel, n_meas):
    op, H, ctx = _setup(oracle, m, n, nlevel, n_meas)
    rng = np.random.default_rng(2)
    for level in range(nlevel - 1):
        nf, ncoarse = H.level_op(level).ndof, H.level_op(level + 1).ndof
        r, xc, x, f = rng.standard_normal(nf), rng.standard_normal(ncoarse), rng.standard_normal(nf), rng.standard_normal(nf)
        assert rel(ctx.restrict(level, r), H.restrict(level, r)) < TOL
        assert rel(ctx.prolongate_add(level, 0.7, xc, x), H.prolongate_add(level, 0.7, xc, x)) < TOL
        # fused residual + restrict (multigridmc_sampler.cc:118-120)
        ref = H.restrict(level, f - H.level_op(level).apply(x))
        assert rel(ctx.residual_restrict(level, f, x), ref) < TOL
        # adjointness <x_c, R r> = <R^T x_c, r>  (test_intergrid.hh:153-171)
        lhs = xc.dot(ctx.restrict(level, r))
        rhs = ctx.prolongate_add(level, 1.0, xc, np.zeros(nf)).dot(r)
        assert abs(lhs - rhs) < 1e-10 * abs(lhs)


@pytest.mark.parametrize("n,nlevel,n_meas", CASES)
@pytest.mark.parametrize("omega", [1.0, 0.8])
def test_smoothers(oracle, m, n, nlevel, n_meas, omega):
    op, H, ctx = _setup(oracle, m, n, nlevel, n_meas)
    rng = np.random.default_rng(3)
    for level in range(nlevel):
        nd = H.level_op(level).ndof
        b, x = rng.standard_normal(nd), rng.standard_normal(nd)
        for kind, direction, nsmooth in (("SOR", 1, 1), ("SOR", 2, 1), ("SOR", 1, 2), ("SSOR", 1, 1), ("SSOR", 1, 2)):
            ref = H.smoother(level, kind, omega, nsmooth, direction).apply(b, x)
            got = ctx.smoother_apply(level, kind, b, x, omega=omega, nsmooth=nsmooth, direction=direction)
            assert rel(got, ref) < TOL, (level, kind, direction, nsmooth)


@pytest.mark.parametrize("n,nlevel,n_meas", [((64, 64), 3, 0), ((128, 128), 4, 5)])
def test_ssor_fixed_point(oracle, m, n, nlevel, n_meas):
    """test_smoother.hh:90-114: SSOR (omega = 0.8) leaves the exact solution invariant."""
    op, H, ctx = _setup(oracle, m, n, nlevel, n_meas)
    x_exact = np.random.default_rng(4).standard_normal(op.ndof)
    b = op.apply(x_exact)
    x = ctx.smoother_apply(0, "SSOR", b, x_exact, omega=0.8)
    assert rel(x, x_exact) < (1e-10 if n_meas else TOL)


@pytest.mark.parametrize("n,nlevel,n_meas", CASES)
@pytest.mark.parametrize("omega", [1.0, 0.8])
def test_gibbs_sweeps_same_philox_stream(oracle, m, n, nlevel, n_meas, omega):
    seed = 1234567
    op, H, ctx = _setup(oracle, m, n, nlevel, n_meas, seed=seed)
    rng = np.random.default_rng(5)
    for level in range(nlevel):
        nd = H.level_op(level).ndof
        f, x = rng.standard_normal(nd), rng.standard_normal(nd)
        for kind, direction, nsmooth in (("SOR", 1, 1), ("SOR", 2, 2), ("SSOR", 1, 1)):
            s = H.sampler(level, kind, omega=omega, nsmooth=nsmooth, direction=direction, rng=None, philox_seed=seed)
            s.set_philox_position(11, 0, 3)
            ref = s.apply(f, x)
            ctx.set_philox_position(11, 3)
            got = ctx.sampler_apply(level, kind, f, x, omega=omega, nsmooth=nsmooth, direction=direction)
            assert rel(got, ref) < 1e-11, (level, kind, direction, nsmooth)
            # and the noise is really there
            assert rel(got, ctx.smoother_apply(level, kind, f, x, omega=omega, nsmooth=1, direction=direction)) > 1e-3


@pytest.mark.parametrize("n,nlevel,n_meas", CASES)
def test_coarse_solve_and_sample(oracle, m, n, nlevel, n_meas):
    seed = 99
    op, H, ctx = _setup(oracle, m, n, nlevel, n_meas, seed=seed)
    lc = nlevel - 1
    nd = H.level_op(lc).ndof
    b = np.random.default_rng(6).standard_normal(nd)
    assert rel(ctx.coarse_solve(b), H.cholesky_solver(lc).apply(b, np.zeros(nd))) < 1e-10
    s = H.sampler(lc, "Cholesky", rng=None, philox_seed=seed)
    s.set_philox_position(5, 0, 2)
    ctx.set_philox_position(5, 2)
    assert rel(ctx.coarse_sample(b), s.apply(b, np.zeros(nd))) < 1e-10


@pytest.mark.parametrize("n,nlevel,n_meas,kw", [
    ((64, 64), 3, 0, {}),
    ((64, 64), 4, 0, dict(smoother="SOR", cycle=2, npresmooth=2)),
    ((128, 128), 4, 5, {}),
    ((128, 64), 3, 4, dict(omega=0.9, npostsmooth=2, coarse_scaling=0.9)),
])
def test_multigrid_preconditioner_and_loop_solver(oracle, m, n, nlevel, n_meas, kw):
    op, H, ctx = _setup(oracle, m, n, nlevel, n_meas, **kw)
    b = oracle.StdRng(1482817).normal(op.ndof)  # driver_mg.cc:165-172
    prec = H.preconditioner(**kw)
    assert rel(ctx.mgprec_apply(b), prec.apply(b, np.zeros_like(b))) < 1e-11
    x_ref, h_ref, it_ref, cv_ref = oracle.loop_solve(op, prec, b, rtol=1e-12, atol=1e-15, maxiter=30)
    x, h, it, cv = ctx.loop_solve(b, rtol=1e-12, atol=1e-15, maxiter=30)
    assert len(h) == len(h_ref) and it == it_ref and cv == cv_ref
    r0 = np.linalg.norm(b)
    # residual history: identical to the oracle (same ordering) to 1e-12 of the initial residual ...
    assert np.abs(h - h_ref).max() < 1e-12 * r0
    # ... and to 1e-9 relative for every entry that is still well above the rounding floor
    big = h_ref > 1e-6 * r0
    assert np.abs(h[big] / h_ref[big] - 1).max() < 1e-9
    assert rel(x, x_ref) < 1e-11


@pytest.mark.parametrize("n,nlevel,n_meas,kw", [
    ((64, 64), 3, 0, {}),
    ((64, 64), 4, 0, dict(smoother="SOR", cycle=2, npresmooth=2, omega=0.9)),
    ((128, 128), 4, 5, {}),
    ((64, 64), 3, 4, dict(coarse_solver="SSOR", ncoarsesmooth=2)),
])
def test_mgmc_cycle_same_philox_stream(oracle, m, n, nlevel, n_meas, kw):
    """MultigridMCSampler::apply: three consecutive samples must follow the oracle's chain when both
    use the colour ordering and the same Philox stream."""
    seed = 5418513
    op, H, ctx = _setup(oracle, m, n, nlevel, n_meas, seed=seed, **kw)
    rng = np.random.default_rng(8)
    f, x0 = rng.standard_normal(op.ndof), rng.standard_normal(op.ndof)
    sampler = H.mgmc(rng=None, philox_seed=seed, **kw)
    ctx.set_philox_position(0)
    xr, xg = x0, x0
    for k in range(3):
        xr = sampler.apply(f, xr)
        xg = ctx.mgmc_apply(f, xg)
        assert rel(xg, xr) < 1e-10, k
    # device-resident loop (CUDA graph replay + device QoI) continues the same chain
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


# ---- squared shifted Laplacian (biharmonic-type prior, BASELINE config 4): 13 / 21-point stencils with
#      boundary-ring classes, 9-colour sweeps; never exercised inside multigrid by the reference's own tests ----
SQ = "squared_shiftedlaplace_fd"
SQ_CASES = [((64, 64), 3, 0), ((64, 32), 2, 0), ((128, 128), 4, 4)]


@pytest.mark.parametrize("n,nlevel,n_meas", SQ_CASES)
def test_squared_operator_single_level_ops(oracle, m, n, nlevel, n_meas):
    seed = 77
    op, H, ctx = _setup(oracle, m, n, nlevel, n_meas, pde=SQ, seed=seed)
    rng = np.random.default_rng(21)
    for level in range(nlevel):
        lop = H.level_op(level)
        nd = lop.ndof
        assert ctx.ndof(level) == nd and ctx.level_info(level)[3] == H.ncolours(level) == 9
        x, b = rng.standard_normal(nd), rng.standard_normal(nd)
        assert rel(ctx.op_apply(level, x), lop.apply(x)) < TOL
        for kind, direction, nsmooth, omega in (("SOR", 1, 1, 1.0), ("SOR", 2, 1, 0.8), ("SSOR", 1, 2, 0.9)):
            ref = H.smoother(level, kind, omega, nsmooth, direction).apply(b, x)
            assert rel(ctx.smoother_apply(level, kind, b, x, omega=omega, nsmooth=nsmooth, direction=direction), ref) < TOL, (level, kind)
            s = H.sampler(level, kind, omega=omega, nsmooth=nsmooth, direction=direction, rng=None, philox_seed=seed)
            s.set_philox_position(3, 0, 1)
            ctx.set_philox_position(3, 1)
            assert rel(ctx.sampler_apply(level, kind, b, x, omega=omega, nsmooth=nsmooth, direction=direction), s.apply(b, x)) < 1e-11, (level, kind)
        if level < nlevel - 1:
            ref = H.restrict(level, b - lop.apply(x))
            assert rel(ctx.residual_restrict(level, b, x), ref) < 1e-11
    # SSOR fixed point (test_smoother.hh:90-114) on the squared operator
    x_exact = rng.standard_normal(op.ndof)
    assert rel(ctx.smoother_apply(0, "SSOR", op.apply(x_exact), x_exact, omega=0.8), x_exact) < 1e-9


@pytest.mark.parametrize("n,nlevel,n_meas,kw", [
    ((64, 64), 3, 0, {}),
    ((128, 128), 4, 4, dict(npresmooth=2, npostsmooth=2)),
])
def test_squared_operator_multigrid_and_mgmc(oracle, m, n, nlevel, n_meas, kw):
    seed = 5418513
    op, H, ctx = _setup(oracle, m, n, nlevel, n_meas, pde=SQ, seed=seed, **kw)
    b = oracle.StdRng(1482817).normal(op.ndof)
    prec = H.preconditioner(**kw)
    assert rel(ctx.mgprec_apply(b), prec.apply(b, np.zeros_like(b))) < 1e-10
    # the V-cycle of the squared operator contracts slowly (factor ~0.7 with coarse_scaling = 1, in the
    # lexicographic reference ordering as well): same history, same iteration count as the oracle
    x_ref, h_ref, it_ref, cv_ref = oracle.loop_solve(op, prec, b, rtol=1e-6, atol=1e300, maxiter=80)
    x, h, it, cv = ctx.loop_solve(b, rtol=1e-6, atol=1e300, maxiter=80)
    assert (it, cv) == (it_ref, cv_ref) and cv
    assert np.abs(h - h_ref).max() < 1e-10 * np.linalg.norm(b)
    assert rel(x, x_ref) < 1e-9
    rng = np.random.default_rng(22)
    f, x0 = rng.standard_normal(op.ndof), rng.standard_normal(op.ndof)
    sampler = H.mgmc(rng=None, philox_seed=seed, **kw)
    ctx.set_philox_position(0)
    xr, xg = x0, x0
    for k in range(2):
        xr = sampler.apply(f, xr)
        xg = ctx.mgmc_apply(f, xg)
        assert rel(xg, xr) < 1e-9, k
    ctx.set_rhs(f)
    ctx.set_state(xg)
    ctx.sample(2, series=False)  # graph replay continues the chain
    xr = sampler.apply(f, sampler.apply(f, xr))
    assert rel(ctx.get_state(), xr) < 1e-9


def test_chains_are_independent_and_reproducible(m):
    """nchains > 1: chain c of a batched context equals a single-chain context with first_chain = c."""
    n, nlevel = 64, 3
    rng = np.random.default_rng(9)
    batched = m.Context(n, n, nlevel, nchains=3, seed=42)
    nd = batched.ndof()
    f, x0 = rng.standard_normal(nd), rng.standard_normal(nd)
    batched.set_rhs(np.tile(f, 3))
    batched.set_state(np.tile(x0, 3))
    batched.set_philox_position(0)
    batched.sample(3, series=False)
    xs = batched.get_state().reshape(3, nd)
    assert rel(xs[0], xs[1]) > 1e-3
    for c in range(3):
        single = m.Context(n, n, nlevel, nchains=1, first_chain=c, seed=42)
        single.set_rhs(f)
        single.set_state(x0)
        single.set_philox_position(0)
        single.sample(3, series=False)
        assert rel(xs[c], single.get_state()) < 1e-13


def test_batched_chains_with_measurements_equal_single_chains(m):
    """Multi-tile lattice with measurements, several chains per launch: every chain is BIT-identical to the same chain
    run on its own, and the run is reproducible (regression: a low-rank tile flag shared by the chains could change
    under a CTA that was still reading it)."""
    from multigridmc_b200 import workloads as w

    n, nlevel, nb = 1024, 6, 4
    loc, _, _, var = w.measurement_set(8)
    B = w.point_measurement_matrix(n, n, loc, var, 1e-3)
    rng = np.random.default_rng(21)
    nd = (n - 1) ** 2
    f, x0 = rng.standard_normal(nd), rng.standard_normal(nd)

    def run(nchains, first):
        ctx = m.Context(n, n, nlevel, B=B, nchains=nchains, first_chain=first, seed=7)
        ctx.set_rhs(np.tile(f, nchains))
        ctx.set_state(np.tile(x0, nchains))
        ctx.set_philox_position(0)
        ctx.sample(3, series=False)
        x = ctx.get_state().reshape(nchains, nd)
        ctx.close()
        return x

    xa, xb = run(nb, 0), run(nb, 0)
    assert np.array_equal(xa, xb)
    for c in (0, nb - 1):
        assert np.array_equal(xa[c], run(1, c)[0])


def test_statistics_small_lattice(oracle, m):
    """Sampled QoI mean / variance against the exact posterior values (linear_operator.hh:153-174)
    within Monte-Carlo error bars: 16 independent chains x 4000 samples on a 32x32 posterior."""
    n, nlevel, nchains, nsamples = (32, 32), 3, 16, 4000
    prior = oracle.Operator.prior(n, "shiftedlaplace_fd", Lambda=0.2)
    rng = np.random.default_rng(10)
    locs = 0.15 + 0.7 * rng.random((6, 2))
    var = 1.0 + rng.random(6)
    op = prior.measured(locs, var, variance_scaling=1e-3)
    y = 1.0 + 3.0 * rng.random(6)
    xbar = np.zeros(op.ndof)
    mean_exact = op.mean(xbar, y)
    f = op.apply(mean_exact)
    b_obs = op.measurement_vector([0.5, 0.5], 0.0)
    z_mean, z_var = op.observed_mean_and_variance(xbar, y, b_obs)
    ctx = m.Context(n[0], n[1], nlevel, Lambda=0.2, B=op.B(), nchains=nchains, seed=2024)
    idx = np.nonzero(b_obs)[0]
    ctx.set_qoi(idx, b_obs[idx])
    ctx.set_rhs(np.tile(f, nchains))
    ctx.set_state(np.zeros(op.ndof * nchains))
    ctx.sample(200, series=False)  # warm-up
    z = ctx.sample(nsamples)       # (nsamples, nchains)
    chain_means = z.mean(axis=0)
    se_mean = chain_means.std(ddof=1) / np.sqrt(nchains)
    assert abs(chain_means.mean() - z_mean) < 4 * se_mean + 1e-12
    chain_vars = z.var(axis=0, ddof=1)
    se_var = chain_vars.std(ddof=1) / np.sqrt(nchains)
    assert abs(chain_vars.mean() - z_var) < 4 * se_var
    # integrated autocorrelation time (statistics.cc:65-79) must be close to 1 for MGMC
    tau = np.mean([oracle.tau_int(z[:, c], 20) for c in range(nchains)])
    assert tau < 2.0


def test_statistics_agree_with_reference_chain(oracle, m):
    """BASELINE north star: the coloured / Philox GPU chain and the reference's lexicographic / std::mt19937_64
    chain (oracle, reference ordering and RNG call order) sample the same law: QoI mean, variance and
    integrated autocorrelation time agree within Monte-Carlo error bars (SURVEY.md section 8c procedure)."""
    n, nlevel, nchains, nsamples, nref = (32, 32), 3, 16, 3000, 12000
    prior = oracle.Operator.prior(n, "shiftedlaplace_fd", Lambda=0.2)
    rng = np.random.default_rng(12)
    locs, var, y = 0.15 + 0.7 * rng.random((5, 2)), 1.0 + rng.random(5), 1.0 + 3.0 * rng.random(5)
    op = prior.measured(locs, var, variance_scaling=1e-2)
    f = op.apply(op.mean(np.zeros(op.ndof), y))
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
    ctx = m.Context(n[0], n[1], nlevel, Lambda=0.2, B=op.B(), nchains=nchains, seed=777)
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
    # tau_int of a coloured sweep need not equal the lexicographic one: it must not be worse beyond its error
    assert tau.mean() <= tau_ref + 3 * (tau.std(ddof=1) / np.sqrt(nchains) + 0.15)
    assert tau.mean() < 2.0 and tau_ref < 2.0


@pytest.mark.parametrize("n,nlevel", [((1024, 1024), 6), ((4096, 4096), 8)])
def test_full_size_properties(m, n, nlevel):
    """BASELINE sizes, checked through size-independent identities (no oracle at this size):
    adjointness of restrict / prolongate, SSOR fixed point, and MG convergence to a known solution."""
    ctx = m.Context(n[0], n[1], nlevel, Lambda=0.2, npresmooth=2, npostsmooth=2)
    rng = np.random.default_rng(11)
    nd, ndc = ctx.ndof(0), ctx.ndof(1)
    r, xc = rng.standard_normal(nd), rng.standard_normal(ndc)
    lhs = xc.dot(ctx.restrict(0, r))
    rhs = ctx.prolongate_add(0, 1.0, xc, np.zeros(nd)).dot(r)
    assert abs(lhs - rhs) < 1e-10 * abs(lhs)
    x_exact = rng.standard_normal(nd)
    b = ctx.op_apply(0, x_exact)
    assert rel(ctx.smoother_apply(0, "SSOR", b, x_exact, omega=0.8), x_exact) < 1e-12
    x, hist, niter, conv = ctx.loop_solve(b, rtol=1e-11, atol=1e300, maxiter=40)
    assert conv and niter < 25
    assert rel(x, x_exact) < 1e-8


def test_series_shorter_than_a_later_run_without_series(m):
    """Round-1 advisor finding: the QoI series buffer is sized by the call that asks for it; a later call WITHOUT series
    that runs more cycles must not write past it (several chains per launch).  The chain must simply continue: the state
    after series(4) + no-series(12) + series(2) equals the state after one run of 18 cycles, and the recorded values are
    the first 4 and the last 2 of that run's series."""
    from multigridmc_b200 import workloads as w

    n, nlevel, nchains = 256, 4, 4
    loc, _, _, var = w.measurement_set(8)
    B = w.point_measurement_matrix(n, n, loc, var, 1e-3)
    nd = (n - 1) ** 2
    rng = np.random.default_rng(0)
    f = np.tile(rng.standard_normal(nd), nchains)

    def fresh():
        ctx = m.Context(n, n, nlevel, B=B, nchains=nchains, seed=3)
        ctx.set_rhs(f)
        ctx.set_state(np.zeros(nd * nchains))
        ctx.set_qoi([nd // 2, 7], [1.0, 0.5])
        ctx.set_philox_position(0)
        return ctx

    a = fresh()
    z1 = a.sample(4, series=True)
    a.sample(12, series=False)
    z2 = a.sample(2, series=True)
    xa = a.get_state()
    b = fresh()
    zb = b.sample(18, series=True)
    assert np.array_equal(xa, b.get_state())
    assert np.array_equal(z1, zb[:4]) and np.array_equal(z2, zb[16:])
    # setting the same functional again is a no-op (no re-capture, no new buffers); a different one takes effect
    b.set_qoi([nd // 2, 7], [1.0, 0.5])
    z3 = b.sample(1, series=True)
    b.set_qoi([7], [2.0])
    z4 = b.sample(1, series=True)
    assert np.all(np.isfinite(z3)) and np.allclose(z4[0], 2.0 * b.get_state().reshape(nchains, nd)[:, 7])
