"""3d lattices (Lattice3d, lattice/lattice3d.hh; ShiftedLaplaceFDOperator with dim = 3, shiftedlaplace_fd_operator.cc:33-56;
trilinear IntergridOperatorLinear, intergrid_operator_linear.cc:8-30): the host set-up algebra against the oracle's sparse
triple product R A R^T (CPU), and every C-ABI entry point against the CPU oracle on the same inputs (GPU): 1e-12 relative for
the deterministic kernels, the oracle fed by the same Philox stream for the samplers -- as tests/test_gpu_parity.py does in 2d.
"""
import numpy as np
import pytest

import multigridmc_b200 as m
from multigridmc_b200 import capi

TOL = 1e-12


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def _stencil_matrix3(st, nx, ny, nz):
    """Dense matrix of a uniform radius-1 stencil [dk + 1, dj + 1, di + 1] on the interior vertices (lattice3d.hh:122-135)."""
    w, h, d = nx - 1, ny - 1, nz - 1
    A = np.zeros((w * h * d, w * h * d))
    for k in range(1, nz):
        for j in range(1, ny):
            for i in range(1, nx):
                row = ((k - 1) * h + j - 1) * w + i - 1
                for dk in (-1, 0, 1):
                    for dj in (-1, 0, 1):
                        for di in (-1, 0, 1):
                            ii, jj, kk = i + di, j + dj, k + dk
                            if 1 <= ii < nx and 1 <= jj < ny and 1 <= kk < nz:
                                A[row, ((kk - 1) * h + jj - 1) * w + ii - 1] = st[dk + 1, dj + 1, di + 1]
    return A


@pytest.mark.parametrize("n,nlevel", [((8, 8, 8), 2), ((16, 8, 12), 2), ((16, 16, 16), 3)])
def test_galerkin_stencils_3d_match_oracle_triple_product(oracle, n, nlevel):
    """LinearOperator::coarsen (linear_operator.cc:10-23) in 3d: the 27-point stencil algebra reproduces R A R^T of the oracle
    on every level, entry by entry; colours as the oracle's scan of the matrix rows (2 on the fine level, 8 below)."""
    op = oracle.Operator.prior(n, "shiftedlaplace_fd", Lambda=0.2)
    H = oracle.Hierarchy(op, nlevel, oracle.COLOUR)
    desc = capi.make_desc(n[0], n[1], nlevel, Lambda=0.2, nz=n[2])
    nx, ny, nz = n
    for level in range(nlevel):
        st, nc = m.host_stencil3(desc, level)
        A_ref = H.level_op(level).csr().toarray()
        A = _stencil_matrix3(st, nx, ny, nz)
        assert np.abs(A - A_ref).max() <= 1e-12 * np.abs(A_ref).max(), f"level {level}"
        assert nc == H.ncolours(level) == (2 if level == 0 else 8)
        nx, ny, nz = nx // 2, ny // 2, nz // 2


def test_create_3d_rejects_what_the_reference_rejects():
    # lattice3d.hh:242-257: odd extent / no interior vertex -> exit(-1) in the reference
    for n, nlevel in (((8, 6, 8), 3), ((4, 8, 8), 3)):
        desc = capi.make_desc(n[0], n[1], nlevel, nz=n[2])
        with pytest.raises(m.MgmcError) as e:
            m.host_stencil3(desc, 0)
        assert e.value.code == -1
    # not on the device path: the squared operator in 3d (2d only in the reference as well)
    desc = capi.make_desc(8, 8, 2, nz=8, pde="squared_shiftedlaplace_fd")
    with pytest.raises(m.MgmcError) as e:
        m.host_stencil3(desc, 0)
    assert e.value.code == -2


# ---------------------------------------------------------------------------------------------------------------------
# GPU parity
# ---------------------------------------------------------------------------------------------------------------------
def _setup(oracle, n, nlevel, n_meas=0, radius=0.0, measure_global=False, **kw):
    op = oracle.Operator.prior(n, "shiftedlaplace_fd", Lambda=0.2)
    if n_meas:
        # MeasuredOperator on a 3d lattice (measured_operator.cc:9-49, :69-170 with dim = 3): point / ball measurements,
        # optionally the global average (a dense column of B)
        rng = np.random.default_rng(7)
        locs = 0.15 + 0.7 * rng.random((n_meas, 3))
        op = op.measured(locs, 1.0 + rng.random(n_meas), variance_scaling=1e-4, radius=radius, measure_global=measure_global, variance_global=1e-3)
    H = oracle.Hierarchy(op, nlevel, oracle.COLOUR)
    ctx = m.Context(n[0], n[1], nlevel, nz=n[2], Lambda=0.2, B=op.B() if n_meas else None, **kw)
    return op, H, ctx


# (n, nlevel, measurements, radius, global measurement)
CASES = [((16, 16, 16), 3, 0, 0.0, False), ((24, 8, 16), 2, 0, 0.0, False), ((32, 16, 8), 2, 0, 0.0, False), ((12, 20, 28), 2, 0, 0.0, False),
         ((16, 16, 16), 3, 4, 0.0, False), ((32, 16, 16), 2, 3, 0.12, False), ((16, 16, 16), 2, 2, 0.0, True)]


@pytest.mark.gpu
@pytest.mark.parametrize("n,nlevel,n_meas,radius,glob", CASES)
def test_single_level_operations_3d(oracle, n, nlevel, n_meas, radius, glob):
    seed = 4711
    op, H, ctx = _setup(oracle, n, nlevel, n_meas, radius, glob, seed=seed)
    rng = np.random.default_rng(1)
    for level in range(nlevel):
        lop = H.level_op(level)
        nd = lop.ndof
        assert ctx.ndof(level) == nd
        assert ctx.level_info(level)[3] == H.ncolours(level) == (2 if level == 0 else 8)
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
            assert rel(ctx.restrict(level, x), H.restrict(level, x)) < TOL
            assert rel(ctx.residual_restrict(level, b, x), H.restrict(level, b - lop.apply(x))) < TOL
            assert rel(ctx.prolongate_add(level, 0.7, xc, x), H.prolongate_add(level, 0.7, xc, x)) < TOL
    # SSOR leaves the exact solution invariant (test_smoother.hh:90-114)
    x_exact = rng.standard_normal(op.ndof)
    assert rel(ctx.smoother_apply(0, "SSOR", op.apply(x_exact), x_exact, omega=0.8), x_exact) < (1e-9 if n_meas else TOL)
    # intergrid adjointness <R x, y> = <x, R^T y> (test_intergrid.hh:87-120) on the device transfers
    if nlevel > 1:
        xf, yc = rng.standard_normal(op.ndof), rng.standard_normal(H.level_op(1).ndof)
        lhs = ctx.restrict(0, xf).dot(yc)
        rhs = xf.dot(ctx.prolongate_add(0, 1.0, yc, np.zeros(op.ndof)))
        assert abs(lhs - rhs) < 1e-12 * abs(lhs)
    # coarsest level: dense factor
    lc = nlevel - 1
    nd = H.level_op(lc).ndof
    b = rng.standard_normal(nd)
    assert rel(ctx.coarse_solve(b), H.cholesky_solver(lc).apply(b, np.zeros(nd))) < 1e-10
    s = H.sampler(lc, "Cholesky", rng=None, philox_seed=seed)
    s.set_philox_position(5, 0, 2)
    ctx.set_philox_position(5, 2)
    assert rel(ctx.coarse_sample(b), s.apply(b, np.zeros(nd))) < 1e-10


@pytest.mark.gpu
@pytest.mark.parametrize("n,nlevel,n_meas,kw", [
    ((16, 16, 16), 3, 0, {}),
    ((32, 32, 32), 3, 0, dict(npresmooth=2, npostsmooth=2)),
    ((16, 32, 16), 3, 0, dict(smoother="SOR", cycle=2, npresmooth=2, omega=0.9)),
    ((32, 16, 16), 2, 0, dict(coarse_solver="SSOR", ncoarsesmooth=2)),
    ((32, 32, 32), 3, 5, {}),
    ((16, 16, 32), 3, 3, dict(npresmooth=2, omega=0.9, measure_global=True)),
])
def test_multigrid_solver_and_mgmc_chain_3d(oracle, n, nlevel, n_meas, kw):
    """MultigridPreconditioner + LoopSolver (multigrid_preconditioner.cc:74-101, loop_solver.cc:9-53) and three MGMC samples
    (multigridmc_sampler.cc:103-138) + the graph-replayed device loop against the oracle chain on the same Philox stream."""
    seed = 5418513
    glob = kw.pop("measure_global", False)
    op, H, ctx = _setup(oracle, n, nlevel, n_meas, 0.0, glob, seed=seed, **kw)
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
    idx = np.array([op.ndof // 2 + 3, 5, op.ndof - 1])
    val = np.array([1.0, -0.5, 0.25])
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
def test_large_lattice_properties_and_chains_3d():
    """128^3, 5 levels (size-independent identities, no oracle): the operator is symmetric, SSOR keeps the exact solution, the
    multigrid-preconditioned Richardson iteration converges to it; batched chains equal single chains bit for bit."""
    n, nlevel = 128, 5
    ctx = m.Context(n, n, nlevel, nz=n, npresmooth=2, npostsmooth=2)
    rng = np.random.default_rng(11)
    nd = ctx.ndof(0)
    assert nd == (n - 1) ** 3
    x, y = rng.standard_normal(nd), rng.standard_normal(nd)
    Ax, Ay = ctx.op_apply(0, x), ctx.op_apply(0, y)
    assert abs(y.dot(Ax) - x.dot(Ay)) < 1e-11 * abs(y.dot(Ax))
    assert rel(ctx.smoother_apply(0, "SSOR", Ax, x, omega=0.8), x) < 1e-11
    xs, h, it, cv = ctx.loop_solve(Ax, rtol=1e-11, atol=1e300, maxiter=40)
    assert cv and it <= 20
    assert rel(xs, x) < 1e-8
    ctx.close()
    n, nlevel, seed = 32, 3, 5
    nd = (n - 1) ** 3
    f = rng.standard_normal(nd)
    batch = m.Context(n, n, nlevel, nz=n, seed=seed, nchains=3)
    batch.set_rhs(np.tile(f, 3))
    batch.set_state(np.zeros(3 * nd))
    batch.sample(3, series=False)
    xb = batch.get_state().reshape(3, nd)
    for ch in range(3):
        single = m.Context(n, n, nlevel, nz=n, seed=seed, first_chain=ch)
        single.set_rhs(f)
        single.set_state(np.zeros(nd))
        single.sample(3, series=False)
        assert np.array_equal(single.get_state(), xb[ch]), ch
        single.close()
    assert rel(xb[0], xb[1]) > 1e-3


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(64, 64, None), (16, 16, 16)])
def test_moment_fields_equal_manual_accumulation(shape):
    """mgmc_sample_moments (driver_mgmc.cc:146-151: running mean / second moment of the field over the samples) against the same
    chain advanced sample by sample with the fields accumulated on the host -- 2d and 3d layouts."""
    nx, ny, nz = shape
    rng = np.random.default_rng(5)
    out = []
    for mode in range(2):
        ctx = m.Context(nx, ny, 3, nz=nz, seed=77)
        nd = ctx.ndof()
        if mode == 0:
            f, x0 = rng.standard_normal(nd), rng.standard_normal(nd)
        ctx.set_rhs(f)
        ctx.set_state(x0)
        ctx.set_philox_position(0)
        if mode == 0:
            out.append(ctx.sample_moments(6))
        else:
            mean, second = np.zeros(nd), np.zeros(nd)
            for k in range(6):
                ctx.sample(1, series=False)
                v = ctx.get_state()
                mean += (v - mean) / (k + 1.0)
                second += (v * v - second) / (k + 1.0)
            out.append((mean, second))
        ctx.close()
    assert rel(out[0][0], out[1][0]) < 1e-13 and rel(out[0][1], out[1][1]) < 1e-13
    assert np.all(out[0][1] - out[0][0] ** 2 > -1e-12)
