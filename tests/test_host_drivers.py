"""Host drop-in layer (host/): the reference's class names, drivers and .cfg formats over the C ABI.

CPU part: the drivers build, the libconfig-subset reader parses the reference's file format, lattice
known answers (test_lattice.hh:166).  GPU part: driver_mg reproduces the oracle's residual history
(colour ordering), driver_mgmc reproduces the exact posterior mean / variance of the observation within
Monte-Carlo error bars and writes the reference's output files."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "host")
CONFIGS = os.path.join(HOST, "configs")


@pytest.fixture(scope="module")
def built():
    import multigridmc_b200 as m

    m.build()
    subprocess.check_call(["make", "-C", HOST], stdout=subprocess.DEVNULL)
    return HOST


def test_drivers_build(built):
    for exe in ("driver_mg", "driver_mgmc", "test_config"):
        assert os.access(os.path.join(built, exe), os.X_OK)


def test_config_reader_reads_reference_format(built):
    from multigridmc_b200 import workloads as w

    out = subprocess.check_output([os.path.join(built, "test_config"), "c3_mgmc_4096.cfg"], cwd=CONFIGS, text=True)
    assert "dim=2 operator=posterior do_ssor=0 do_multigridmc=1" in out
    assert "lattice=4096,4096,4096" in out
    assert "multigrid nlevel=8 smoother=SSOR coarse=Cholesky pre=1 post=1 ncoarse=1 cycle=1 scaling=1 omega=1" in out
    assert "solver rtol=9.9999999999999998e-13 atol=1.0000000000000001e-15 maxiter=100 verbose=2" in out
    assert "sampling nsamples=1000 nwarmup=100 nsteps=8 nconv=20" in out
    loc, sample, mean, var = w.measurement_set(32)
    rows = [l.split() for l in out.splitlines() if l.startswith("meas ")]
    assert len(rows) == 32
    got = np.array([[float(v) for v in r[2:]] for r in rows])
    assert np.array_equal(got[:, :2], loc) and np.array_equal(got[:, 2], mean) and np.array_equal(got[:, 3], var)
    m = re.search(r"sample=([0-9.e+-]+),([0-9.e+-]+)", out)
    assert (float(m.group(1)), float(m.group(2))) == (float(sample[0]), float(sample[1]))
    # test_lattice.hh:166 (2d lattice 4 x 5): fine_vertex_idx(7) == 38
    assert "lattice2d Nvertex=12 Ncell=20 fine_vertex_idx(7)=38" in out
    # test_lattice.hh:30-101 (1d lattice, n = 6)
    assert ("lattice1d Nvertex=5 Ncell=6 cell5=5 cell(3)=3 cellshifts=4,2,5,3 vertex4=5 vertex(3)=2 shifts=4,2,5,3 fine=7,1,5 "
            "info='1d lattice,    6 points,    5 unknowns'") in out
    assert "lattice3d cellshifts(59)=63,55,60,58,79,39" in out  # test_lattice.hh:189-204
    # test_lattice.hh:171-242 (3d lattice 4 x 5 x 6)
    assert "lattice3d Nvertex=60 Ncell=120 cell53=1,3,2 cell(1,3,2)=53 vertex23=3,4,2 vertex(3,4,2)=23" in out
    assert "lattice3d shifts(23)=26,20,24,22,35,11 fine_vertex_idx(23)=243" in out
    assert "lattice3d coords(23)=0.7500,0.8000,0.3333 coarse(8,4,6)=4,2,3 info='3d lattice,    4 x    5 x    6 points,   60 unknowns'" in out


def test_config_reader_errors_like_the_reference(built, tmp_path):
    """parameters.cc:25-47: message + exit(-1) on a missing file or a missing setting."""
    r = subprocess.run([os.path.join(built, "test_config"), "does_not_exist.cfg"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 255 and "cannot open configuration file" in r.stderr
    (tmp_path / "broken.cfg").write_text("general = { dim = 2; }\n")
    r = subprocess.run([os.path.join(built, "test_config"), "broken.cfg"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 255 and "cannot read configuration" in r.stderr


def _write_cfg(path, text_from, **repl):
    s = open(os.path.join(CONFIGS, text_from)).read()
    for k, v in repl.items():
        s, n = re.subn(rf"(\b{k}\s*=\s*)[^;]+;", rf"\g<1>{v};", s, count=1)
        assert n == 1, k
    open(path, "w").write(s)


@pytest.mark.gpu
def test_driver_mg_residual_history_matches_oracle(built, oracle, tmp_path):
    """driver_mg (config C2 at 256^2, 5 levels, V(2,2) SSOR): printed ||r_k|| == oracle LoopSolver in the same ordering."""
    n, nlevel = 256, 5
    _write_cfg(tmp_path / "mg.cfg", "c2_mg_1024.cfg", nx=n, ny=n, nlevel=nlevel, maxiter=12, filename=f'"{CONFIGS}/measurements_8.cfg"')
    out = subprocess.check_output([os.path.join(built, "driver_mg"), "mg.cfg"], cwd=tmp_path, text=True)
    hist = np.array([float(l.split()[1]) for l in out.splitlines() if re.match(r"^\s*\d+\s+\d\.\d+e[+-]\d+\s", l)])
    op = oracle.Operator.prior((n, n), "shiftedlaplace_fd", Lambda=0.2)
    H = oracle.Hierarchy(op, nlevel, oracle.COLOUR)
    b = oracle.StdRng(1482817).normal(op.ndof)
    prec = H.preconditioner(npresmooth=2, npostsmooth=2)
    _, h_ref, _, _ = oracle.loop_solve(op, prec, b, rtol=1e-12, atol=1e-15, maxiter=12)
    assert len(hist) == len(h_ref) == 12
    big = h_ref > 1e-11 * h_ref[0]  # entries above the rounding floor; printed with 4 significant digits
    assert big.sum() >= 8 and np.abs(hist[big] / h_ref[big] - 1).max() < 2e-3
    assert os.path.exists(tmp_path / "solution.vtk")
    assert hist[-1] / hist[0] < 1e-8


@pytest.mark.gpu
def test_driver_mg_periodic_correlation_length(built, oracle, tmp_path):
    """driver_mg with `correlationlengthmodel = "periodic"` (parameters.cc:225-243, correlationlength_model.hh:83-113; the
    model of the reference's own solver tests, test_solver.hh:41): the host classes evaluate kappa^2 at the vertices, the
    device runs per-vertex coefficients; printed ||r_k|| == oracle LoopSolver on the same operator."""
    n, nlevel = 128, 4
    _write_cfg(tmp_path / "mg.cfg", "c2_mg_1024.cfg", nx=n, ny=n, nlevel=nlevel, maxiter=12, filename=f'"{CONFIGS}/measurements_8.cfg"',
               correlationlengthmodel='"periodic"', Lambda_min=0.15, Lambda_max=0.45)
    out = subprocess.check_output([os.path.join(built, "driver_mg"), "mg.cfg"], cwd=tmp_path, text=True)
    assert "correlation length model = periodic" in out
    hist = np.array([float(l.split()[1]) for l in out.splitlines() if re.match(r"^\s*\d+\s+\d\.\d+e[+-]\d+\s", l)])
    op = oracle.Operator.prior((n, n), "shiftedlaplace_fd", Lambda_min=0.15, Lambda_max=0.45)
    H = oracle.Hierarchy(op, nlevel, oracle.COLOUR)
    b = oracle.StdRng(1482817).normal(op.ndof)
    prec = H.preconditioner(npresmooth=2, npostsmooth=2)
    _, h_ref, _, _ = oracle.loop_solve(op, prec, b, rtol=1e-12, atol=1e-15, maxiter=12)
    assert len(hist) == len(h_ref) == 12
    big = h_ref > 1e-11 * h_ref[0]
    assert big.sum() >= 8 and np.abs(hist[big] / h_ref[big] - 1).max() < 2e-3


@pytest.mark.gpu
def test_driver_mgmc_statistics_and_files(built, tmp_path):
    """driver_mgmc on a 128^2 posterior: sampled mean / variance of the observation agree with the exact
    values the driver prints (computed by device MG solves, linear_operator.hh:153-174) within error bars."""
    _write_cfg(tmp_path / "mgmc.cfg", "small_posterior_128.cfg", filename=f'"{CONFIGS}/measurements_8.cfg"')
    out = subprocess.check_output([os.path.join(built, "driver_mgmc"), "mgmc.cfg"], cwd=tmp_path, text=True)
    for label, fname in (("MultigridMC", "timeseries_multigridmc.txt"), ("SSOR", "timeseries_ssor.txt")):
        blk = out[out.index("**** Multigrid MC ****" if label == "MultigridMC" else "**** SSOR ****"):]
        mean, err = map(float, re.search(rf"{label} mean\s+=\s+(\S+) \+/-\s+(\S+)", blk).groups())
        mean_exact = float(re.search(r"exact mean\s+=\s+(\S+)", blk).group(1))
        var = float(re.search(rf"{label} variance =\s+(\S+)", blk).group(1))
        var_exact = float(re.search(r"exact variance =\s+(\S+)", blk).group(1))
        tau = float(re.search(rf"{label} tau_int\s+=\s+(\S+)", blk).group(1))
        series = np.loadtxt(tmp_path / fname)
        nsamp = len(series)
        assert nsamp == 4000
        assert abs(series.mean() - mean) < 2e-4 * max(abs(mean), 1.0) and abs(series.var(ddof=1) / var - 1) < 2e-3  # printed (5 digits) == file
        if label == "MultigridMC":
            # independent-looking samples (tau_int ~ 1): standard errors from the printed error bar / the Gaussian
            # fourth moment, inflated by tau_int (statistics.cc:65-79)
            assert tau < 2.0
            assert abs(mean - mean_exact) < 4.5 * err * np.sqrt(max(tau, 1.0))
            assert abs(var / var_exact - 1) < 4.5 * np.sqrt(2.0 * max(tau, 1.0) / nsamp)  # SE ~ 2.2 % -> 10 %
        else:
            # the plain SSOR (Gibbs) sampler mixes slowly: its windowed tau_int estimate saturates, so the error bars
            # come from batch means (20 batches of 200 consecutive samples) -- same law, wide bars
            nb = 20
            bm = series.reshape(nb, -1).mean(axis=1)
            se_mean = bm.std(ddof=1) / np.sqrt(nb)
            assert abs(series.mean() - mean_exact) < 5 * se_mean
            bv = ((series.reshape(nb, -1) - mean_exact) ** 2).mean(axis=1)
            se_var = bv.std(ddof=1) / np.sqrt(nb)
            assert abs(bv.mean() - var_exact) < 5 * se_var
            assert tau > 1.5  # (and that is why the driver compares it with MGMC)
    conv = open(tmp_path / "convergence_multigridmc.txt").read()
    assert "q_k = |E[z^k] - E[z]|" in conv and "q_k = |Var[z^k] - Var[z]|" in conv


@pytest.mark.gpu
def test_driver_mgmc_global_measurement(built, tmp_path):
    """driver_mgmc with `measure_global = true` (measured_operator.cc:31-46: the average of the field over the domain is
    observed as well -- a dense column of B): the MultigridMC mean / variance of the observation agree with the exact
    values within error bars, as for point measurements."""
    _write_cfg(tmp_path / "mgmc.cfg", "small_posterior_128.cfg", filename=f'"{CONFIGS}/measurements_8.cfg"', measure_global="true", do_ssor="false")
    out = subprocess.check_output([os.path.join(built, "driver_mgmc"), "mgmc.cfg"], cwd=tmp_path, text=True)
    blk = out[out.index("**** Multigrid MC ****"):]
    mean, err = map(float, re.search(r"MultigridMC mean\s+=\s+(\S+) \+/-\s+(\S+)", blk).groups())
    mean_exact = float(re.search(r"exact mean\s+=\s+(\S+)", blk).group(1))
    var = float(re.search(r"MultigridMC variance =\s+(\S+)", blk).group(1))
    var_exact = float(re.search(r"exact variance =\s+(\S+)", blk).group(1))
    tau = float(re.search(r"MultigridMC tau_int\s+=\s+(\S+)", blk).group(1))
    nsamp = len(np.loadtxt(tmp_path / "timeseries_multigridmc.txt"))
    assert nsamp == 4000 and tau < 2.0
    assert abs(mean - mean_exact) < 4.5 * err * np.sqrt(max(tau, 1.0))
    assert abs(var / var_exact - 1) < 4.5 * np.sqrt(2.0 * max(tau, 1.0) / nsamp)


@pytest.mark.gpu
def test_cholesky_solver_and_sampler_classes(built):
    """Host classes CholeskySolver / Dense-, SparseCholeskySampler + factories (solver/cholesky_solver.hh:21-68,
    sampler/cholesky_sampler.hh:27-196) over mgmc_coarse_solve / mgmc_coarse_sample: exact solve with and without the
    measurement term (test_solver.hh:93-113: 1e-12 there, the device factor gives 1e-10), sample mean and variance
    (test_sampler.hh:215-283)."""
    out = subprocess.check_output([os.path.join(built, "test_cholesky")], text=True)
    sol = re.findall(r"solver (\d): \|x - x_exact\| / \|x_exact\| = ([0-9.e+-]+)\s+\|A x - b\| / \|b\| = ([0-9.e+-]+)", out)
    smp = re.findall(r"sampler (\d): \|mean - A\^-1 f\| / \|A\^-1 f\| = ([0-9.e+-]+)\s+var / exact = ([0-9.]+)", out)
    assert len(sol) == 2 and len(smp) == 2, out
    for _, ex, ax in sol:
        assert float(ex) < 1e-9 and float(ax) < 1e-10
    for _, em, vr in smp:
        # 20000 samples: relative error of the mean field ~ sqrt(tr A^-1 / N) / |A^-1 f|, variance of one entry +- 1 % (1 sigma)
        assert float(em) < 0.05 and abs(float(vr) - 1.0) < 0.04


MEAS_3D = """// four point measurements in the unit cube (format of measurements_template.cfg with dim = 3)
dim = 3;
n = 4;
measurement_locations = [0.25, 0.25, 0.25, 0.75, 0.25, 0.5, 0.25, 0.75, 0.75, 0.6, 0.6, 0.4];
mean = [1.0, 2.0, 3.0, 4.0];
variance = [1.0, 1.5, 1.2, 1.8];
"""


@pytest.mark.gpu
def test_driver_mg_3d_residual_history_matches_oracle(built, oracle, tmp_path):
    """driver_mg with `dim = 3` (Lattice3d, driver_mg.cc:383-392): 32^3, 3 levels, V(2,2) SSOR; printed ||r_k|| == oracle
    LoopSolver in the same (red-black / 8-colour) ordering; solution.vtk in the 3d writer's format (vtk_writer3d.cc:8-58)."""
    n, nlevel = 32, 3
    (tmp_path / "meas3d.cfg").write_text(MEAS_3D)
    _write_cfg(tmp_path / "mg.cfg", "c2_mg_1024.cfg", dim=3, nx=n, ny=n, nz=n, nlevel=nlevel, maxiter=12, filename='"meas3d.cfg"',
               sample_location="[0.5, 0.5, 0.5]")
    out = subprocess.check_output([os.path.join(built, "driver_mg"), "mg.cfg"], cwd=tmp_path, text=True)
    hist = np.array([float(l.split()[1]) for l in out.splitlines() if re.match(r"^\s*\d+\s+\d\.\d+e[+-]\d+\s", l)])
    op = oracle.Operator.prior((n, n, n), "shiftedlaplace_fd", Lambda=0.2)
    H = oracle.Hierarchy(op, nlevel, oracle.COLOUR)
    b = oracle.StdRng(1482817).normal(op.ndof)
    prec = H.preconditioner(npresmooth=2, npostsmooth=2)
    _, h_ref, _, _ = oracle.loop_solve(op, prec, b, rtol=1e-12, atol=1e-15, maxiter=12)
    assert len(hist) == len(h_ref) == 12
    big = h_ref > 1e-11 * h_ref[0]
    assert big.sum() >= 8 and np.abs(hist[big] / h_ref[big] - 1).max() < 2e-3
    vtk = open(tmp_path / "solution.vtk").read().splitlines()
    assert vtk[4] == f"DIMENSIONS {n + 1} {n + 1} {n + 1}" and vtk[5] == "ORIGIN -0.5 -0.5 -5.0"
    assert f"POINT_DATA {(n + 1) ** 3}" in vtk[:10]


@pytest.mark.gpu
def test_driver_mgmc_3d_prior_statistics(built, tmp_path):
    """driver_mgmc with `dim = 3`, operator = "prior" (32^3, 3 levels): the sampled variance of the observation at the vertex
    nearest (0.5, 0.5, 0.5) agrees with the exact b^T A^-1 b the driver prints (device MG solves), the mean with 0."""
    # (the driver's "exact" values are those of the MEASURED operator whatever `operator` says, driver_mgmc.cc:56-58,90-94:
    #  measurement variances of 1e12 make it the prior to 12 digits)
    (tmp_path / "meas3d.cfg").write_text(MEAS_3D.replace("variance = [1.0, 1.5, 1.2, 1.8]", "variance = [1.0e+12, 1.5e+12, 1.2e+12, 1.8e+12]"))
    _write_cfg(tmp_path / "mgmc.cfg", "c1_mgmc_64.cfg", dim=3, nx=32, ny=32, nz=32, nlevel=3, do_ssor="false", measure_convergence="false",
               nsamples=4000, nwarmup=100, filename='"meas3d.cfg"', sample_location="[0.5, 0.5, 0.5]")
    out = subprocess.check_output([os.path.join(built, "driver_mgmc"), "mgmc.cfg"], cwd=tmp_path, text=True)
    assert "3d lattice,   32 x   32 x   32 points" in out
    blk = out[out.index("**** Multigrid MC ****"):]
    mean, err = map(float, re.search(r"MultigridMC mean\s+=\s+(\S+) \+/-\s+(\S+)", blk).groups())
    mean_exact = float(re.search(r"exact mean\s+=\s+(\S+)", blk).group(1))
    var = float(re.search(r"MultigridMC variance =\s+(\S+)", blk).group(1))
    var_exact = float(re.search(r"exact variance =\s+(\S+)", blk).group(1))
    tau = float(re.search(r"MultigridMC tau_int\s+=\s+(\S+)", blk).group(1))
    series = np.loadtxt(tmp_path / "timeseries_multigridmc.txt")
    assert len(series) == 4000
    assert tau < 2.0 and abs(mean_exact) < 1e-8
    assert abs(mean - mean_exact) < 4.5 * err * np.sqrt(max(tau, 1.0))
    assert abs(var / var_exact - 1) < 4.5 * np.sqrt(2.0 * max(tau, 1.0) / len(series))


@pytest.mark.gpu
def test_driver_mgmc_3d_posterior_statistics(built, tmp_path):
    """driver_mgmc with `dim = 3`, operator = "posterior" (32^3, 3 levels, four point measurements; MeasuredOperator on a Lattice3d,
    measured_operator.cc:9-49): sampled mean / variance of the observation at (0.45, 0.55, 0.5) agree with the exact posterior values
    the driver prints (Woodbury with device MG solves, linear_operator.hh:153-174) within Monte-Carlo error bars."""
    (tmp_path / "meas3d.cfg").write_text(MEAS_3D)
    _write_cfg(tmp_path / "mgmc.cfg", "small_posterior_128.cfg", dim=3, nx=32, ny=32, nz=32, nlevel=3, do_ssor="false", filename='"meas3d.cfg"',
               sample_location="[0.45, 0.55, 0.5]", variance_scaling="1.E-3")
    out = subprocess.check_output([os.path.join(built, "driver_mgmc"), "mgmc.cfg"], cwd=tmp_path, text=True)
    blk = out[out.index("**** Multigrid MC ****"):]
    mean, err = map(float, re.search(r"MultigridMC mean\s+=\s+(\S+) \+/-\s+(\S+)", blk).groups())
    mean_exact = float(re.search(r"exact mean\s+=\s+(\S+)", blk).group(1))
    var = float(re.search(r"MultigridMC variance =\s+(\S+)", blk).group(1))
    var_exact = float(re.search(r"exact variance =\s+(\S+)", blk).group(1))
    tau = float(re.search(r"MultigridMC tau_int\s+=\s+(\S+)", blk).group(1))
    series = np.loadtxt(tmp_path / "timeseries_multigridmc.txt")
    assert len(series) == 4000 and tau < 2.0
    assert abs(mean_exact) > 1e-3  # (the measurements pull the mean away from zero)
    assert abs(mean - mean_exact) < 4.5 * err * np.sqrt(max(tau, 1.0))
    assert abs(var / var_exact - 1) < 4.5 * np.sqrt(2.0 * max(tau, 1.0) / len(series))


@pytest.mark.gpu
def test_driver_mg_fem_operator(built, oracle, tmp_path):
    """driver_mg with `pdemodel = "shiftedlaplace_fem"` (driver_mg.cc:129-139; ShiftedLaplaceFEMOperator, constant correlation length):
    9-point stencils on every level, 4-colour sweeps; printed ||r_k|| == oracle LoopSolver on the same operator."""
    n, nlevel = 128, 4
    _write_cfg(tmp_path / "mg.cfg", "c2_mg_1024.cfg", nx=n, ny=n, nlevel=nlevel, maxiter=12, filename=f'"{CONFIGS}/measurements_8.cfg"',
               pdemodel='"shiftedlaplace_fem"')
    out = subprocess.check_output([os.path.join(built, "driver_mg"), "mg.cfg"], cwd=tmp_path, text=True)
    hist = np.array([float(l.split()[1]) for l in out.splitlines() if re.match(r"^\s*\d+\s+\d\.\d+e[+-]\d+\s", l)])
    op = oracle.Operator.prior((n, n), "shiftedlaplace_fem", Lambda=0.2)
    H = oracle.Hierarchy(op, nlevel, oracle.COLOUR)
    b = oracle.StdRng(1482817).normal(op.ndof)
    prec = H.preconditioner(npresmooth=2, npostsmooth=2)
    _, h_ref, _, _ = oracle.loop_solve(op, prec, b, rtol=1e-12, atol=1e-15, maxiter=12)
    assert len(hist) == len(h_ref) == 12
    big = h_ref > 1e-11 * h_ref[0]
    assert big.sum() >= 8 and np.abs(hist[big] / h_ref[big] - 1).max() < 2e-3


def test_host_measurement_functional_matches_oracle(built, oracle):
    """MeasuredOperator::measurement_vector of the host layer (measured_operator.cc:69-170: closest vertex for radius ~ 0, ball average
    against the multilinear hat functions otherwise) on 2d and 3d lattices against the oracle's restatement, entry by entry (no GPU)."""
    out = subprocess.check_output([os.path.join(built, "test_measure")], text=True)
    got = {}
    for line in out.splitlines():
        _, tag, ell, val = line.split()
        got.setdefault(tag, {})[int(ell)] = float(val)
    cases = {
        "2d_point": ((16, 12), [0.37, 0.62], 0.0), "2d_ball": ((16, 12), [0.37, 0.62], 0.15), "2d_edge": ((16, 12), [0.97, 0.02], 0.1),
        "3d_point": ((8, 12, 10), [0.37, 0.62, 0.48], 0.0), "3d_corner_point": ((8, 12, 10), [0.999, 0.001, 1.0], 0.0),
        "3d_ball": ((8, 12, 10), [0.37, 0.62, 0.48], 0.2), "3d_edge": ((8, 12, 10), [0.95, 0.05, 0.5], 0.15),
    }
    assert set(got) == set(cases)
    for tag, (n, x0, radius) in cases.items():
        op = oracle.Operator.prior(n, "shiftedlaplace_fd", Lambda=0.2)
        ref = op.measurement_vector(x0, radius)
        vec = np.zeros(op.ndof)
        for ell, v in got[tag].items():
            vec[ell] = v
        assert np.abs(vec - ref).max() <= 1e-14 * max(np.abs(ref).max(), 1.0), tag
        assert np.count_nonzero(ref) > 0


REFERENCE_SRC = "/root/reference/src"


@pytest.mark.skipif(not os.path.isdir(REFERENCE_SRC), reason="the reference checkout only exists in the build container")
def test_reference_lattice_tests_run_against_the_host_layer(built, tmp_path):
    """Drop-in check at the source level: the reference's OWN test file src/lattice/test_lattice.hh (21 googletest cases: index maps,
    shifts, fine_vertex_idx on 1d / 2d / 3d lattices) compiles UNMODIFIED against the host layer -- `lattice*.hh`, `<Eigen/Dense>` and
    `<gtest/gtest.h>` are served by tests/ref_compat/ -- and every case passes.  The file is compiled from a temporary copy (so that its
    quoted includes resolve to the drop-in headers, not to the reference's own); nothing of it is stored in this repository."""
    import shutil

    shutil.copy(os.path.join(REFERENCE_SRC, "lattice", "test_lattice.hh"), tmp_path / "test_lattice.hh")
    compat = os.path.join(ROOT, "tests", "ref_compat")
    libdir = os.path.join(ROOT, "multigridmc_b200", "csrc")
    exe = str(tmp_path / "run_lattice_tests")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", compat, "-I", HOST, "-I", str(tmp_path), os.path.join(compat, "main_lattice.cc"), "-o", exe,
                           "-L" + libdir, "-lmgmc_b200", "-Wl,-rpath," + libdir])
    out = subprocess.check_output([exe], text=True)
    assert "[  PASSED  ] 21 tests, 0 failed" in out, out
    assert out.count("[       OK ] LatticeTest.") == 21


@pytest.mark.skipif(not os.path.isdir(REFERENCE_SRC), reason="the reference checkout only exists in the build container")
def test_reference_quadrature_tests_run_against_the_host_layer(built, tmp_path):
    """The reference's own src/auxilliary/test_quadrature.hh (Gauss-Legendre rules of order 0 / 1 / 2 integrate monomials up to degree
    1 / 3 / 5 on the unit cube exactly, 1e-12) compiled unmodified against the host layer's GaussLegendreQuadrature."""
    import shutil

    shutil.copy(os.path.join(REFERENCE_SRC, "auxilliary", "test_quadrature.hh"), tmp_path / "test_quadrature.hh")
    compat = os.path.join(ROOT, "tests", "ref_compat")
    libdir = os.path.join(ROOT, "multigridmc_b200", "csrc")
    exe = str(tmp_path / "run_quadrature_tests")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", compat, "-I", HOST, "-I", str(tmp_path), os.path.join(compat, "main_quadrature.cc"), "-o", exe,
                           "-L" + libdir, "-lmgmc_b200", "-Wl,-rpath," + libdir])
    out = subprocess.check_output([exe], text=True)
    assert "[  PASSED  ] 3 tests, 0 failed" in out, out


def _ref_compat_build(tmp_path, rel_test, main_cc, exe_name, syntax_only=False):
    import shutil

    shutil.copy(os.path.join(REFERENCE_SRC, rel_test), tmp_path / os.path.basename(rel_test))
    compat = os.path.join(ROOT, "tests", "ref_compat")
    libdir = os.path.join(ROOT, "multigridmc_b200", "csrc")
    if syntax_only:
        src = tmp_path / f"syntax_{exe_name}.cc"
        src.write_text(f'#include "{os.path.basename(rel_test)}"\nint main() {{ return 0; }}\n')
        subprocess.check_call(["g++", "-std=c++17", "-fsyntax-only", "-I", compat, "-I", HOST, "-I", str(tmp_path), str(src)])
        return None
    exe = str(tmp_path / exe_name)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", compat, "-I", HOST, "-I", str(tmp_path), os.path.join(compat, main_cc), "-o", exe,
                           "-L" + libdir, "-lmgmc_b200", "-Wl,-rpath," + libdir])
    return exe


@pytest.mark.skipif(not os.path.isdir(REFERENCE_SRC), reason="the reference checkout only exists in the build container")
def test_reference_coarsen_operator_tests_run_against_the_library(built, tmp_path):
    """The reference's own TestCoarsenOperator2d / TestCoarsenOperator3d (src/intergrid/test_intergrid.hh:172-207: the Galerkin-coarsened
    FEM operator equals the rediscretised FEM operator of the coarse lattice, 1e-12) compiled unmodified against the host layer and run
    on the CPU: LinearOperator::coarsen / get_sparse go through the library's host-side stencil algebra (mgmc_host_stencil / _stencil3),
    i.e. the reference's test pins the product's own R A R^T.  (The other cases of the file need a device: the file compiles, they are
    filtered out here.)"""
    exe = _ref_compat_build(tmp_path, "intergrid/test_intergrid.hh", "main_intergrid.cc", "run_intergrid_tests")
    out = subprocess.check_output([exe, "TestCoarsenOperator"], text=True)
    assert "[  PASSED  ] 2 tests, 0 failed" in out, out
    assert "IntergridTest.TestCoarsenOperator2d" in out and "IntergridTest.TestCoarsenOperator3d" in out


@pytest.mark.skipif(not os.path.isdir(REFERENCE_SRC), reason="the reference checkout only exists in the build container")
@pytest.mark.parametrize("rel_test", ["smoother/test_smoother.hh", "solver/test_solver.hh", "linear_operator/test_linear_operator.hh", "intergrid/test_intergrid.hh"])
def test_reference_test_files_compile_against_the_host_layer(built, tmp_path, rel_test):
    """Source-level drop-in check of the interfaces that need a device at run time: the reference's own smoother / solver / linear-operator /
    intergrid test files compile UNMODIFIED against host/mgmc_host.hh (class names, constructor signatures, apply / coarsen / restrict /
    prolongate_add / measurement_vector ..., parameter structs) -- compile only; their GPU twins are tests/test_gpu_*.py.  (test_sampler.hh
    and test_cholesky_wrapper.hh define matrix-based operators of their own on Eigen sparse matrices and stay with the oracle.)"""
    _ref_compat_build(tmp_path, rel_test, None, os.path.basename(rel_test).replace(".hh", ""), syntax_only=True)
