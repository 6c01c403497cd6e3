"""GPU parity with a GLOBAL measurement (MeasurementParameters::measure_global, measured_operator.cc:31-46): one column of
B is dense (the average of the field over the domain), so W = M_0^{-1} B has a dense column as well and the measurements
interact on every level.  The in-kernel owner / consumer scheme of the tile kernel does not apply (supports wider than a
tile): the sweeps run one launch each, followed by the stand-alone fix-up kernels.  Compared with the oracle on the same
inputs as tests/test_gpu_parity.py does for point measurements."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.fixture(scope="module")
def m():
    import multigridmc_b200 as mod

    mod.lib()
    return mod


def _setup(oracle, m, n, nlevel, n_meas, radius=0.0, **kw):
    op = oracle.Operator.prior(n, "shiftedlaplace_fd", Lambda=0.2)
    rng = np.random.default_rng(7)
    locs = 0.1 + 0.8 * rng.random((n_meas, 2))
    op = op.measured(locs, 1.0 + rng.random(n_meas), variance_scaling=1e-3, radius=radius, measure_global=True, variance_global=2e-3)
    assert op.m_lowrank == n_meas + 1
    H = oracle.Hierarchy(op, nlevel, oracle.COLOUR)
    ctx = m.Context(n[0], n[1], nlevel, Lambda=0.2, B=op.B(), **kw)
    return op, H, ctx


@pytest.mark.parametrize("n,nlevel,n_meas,kw", [
    ((64, 64), 3, 3, {}),
    ((128, 64), 3, 0, dict(npresmooth=2, omega=0.9)),
    ((256, 256), 5, 4, dict(radius=0.03)),
])
def test_global_measurement_against_oracle(oracle, m, n, nlevel, n_meas, kw):
    seed = 2024
    radius = kw.pop("radius", 0.0)
    op, H, ctx = _setup(oracle, m, n, nlevel, n_meas, radius=radius, seed=seed, **kw)
    rng = np.random.default_rng(1)
    for level in range(nlevel):
        lop = H.level_op(level)
        nd = lop.ndof
        x, b = rng.standard_normal(nd), rng.standard_normal(nd)
        assert rel(ctx.op_apply(level, x), lop.apply(x)) < 1e-12
        for kind, direction, omega in (("SOR", 1, 1.0), ("SOR", 2, 0.8), ("SSOR", 1, 1.0)):
            ref = H.smoother(level, kind, omega, 1, direction).apply(b, x)
            assert rel(ctx.smoother_apply(level, kind, b, x, omega=omega, direction=direction), ref) < 1e-11, (level, kind)
            s = H.sampler(level, kind, omega=omega, nsmooth=1, direction=direction, rng=None, philox_seed=seed)
            s.set_philox_position(3, 0, 1)
            ctx.set_philox_position(3, 1)
            assert rel(ctx.sampler_apply(level, kind, b, x, omega=omega, direction=direction), s.apply(b, x)) < 1e-10, (level, kind)
        if level < nlevel - 1:
            assert rel(ctx.residual_restrict(level, b, x), H.restrict(level, b - lop.apply(x))) < 1e-11
    b = oracle.StdRng(1482817).normal(op.ndof)
    prec = H.preconditioner(**kw)
    assert rel(ctx.mgprec_apply(b), prec.apply(b, np.zeros_like(b))) < 1e-10
    x_ref, h_ref, it_ref, cv_ref = oracle.loop_solve(op, prec, b, rtol=1e-11, atol=1e300, maxiter=40)
    x, h, it, cv = ctx.loop_solve(b, rtol=1e-11, atol=1e300, maxiter=40)
    assert (it, cv) == (it_ref, cv_ref)
    assert np.abs(h - h_ref).max() < 1e-11 * np.linalg.norm(b)
    f, x0 = rng.standard_normal(op.ndof), rng.standard_normal(op.ndof)
    sampler = H.mgmc(rng=None, philox_seed=seed, **kw)
    ctx.set_philox_position(0)
    xr, xg = x0, x0
    for k in range(3):
        xr = sampler.apply(f, xr)
        xg = ctx.mgmc_apply(f, xg)
        assert rel(xg, xr) < 1e-9, k
    ctx.set_rhs(f)
    ctx.set_state(xg)
    ctx.sample(2, series=False)
    xr = sampler.apply(f, sampler.apply(f, xr))
    assert rel(ctx.get_state(), xr) < 1e-9
