#!/usr/bin/env python
"""Writes the parameter / measurement files of the BASELINE configurations in the
reference's libconfig format (parameters_template.cfg, measurements_template.cfg).

  python host/configs/make_configs.py        # regenerates host/configs/*.cfg
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from multigridmc_b200 import workloads as w  # noqa: E402

TEMPLATE = """// {title}
general = {{
    dim = 2;
    do_cholesky = false;
    do_ssor = {do_ssor};
    do_multigridmc = true;
    save_posterior_statistics = false;
    measure_convergence = true;
    operator = "{operator}";
}}
lattice = {{
    nx = {n};
    ny = {n};
    nz = {n};
}}
cholesky = {{
    factorisation = "dense";
}}
smoother = {{
    nsmooth = 1;
    omega = 1.0;
}}
iterative_solver = {{
    rtol = 1.E-12;
    atol = 1.E-15;
    maxiter = {maxiter};
    verbose = 2;
}}
multigrid = {{
    smoother = "SSOR";
    coarse_solver = "Cholesky";
    npresmooth = {nsmooth};
    npostsmooth = {nsmooth};
    ncoarsesmooth = 1;
    omega = 1.0;
    nlevel = {nlevel};
    cycle = 1;
    coarse_scaling = 1.0;
    verbose = 1;
}}
sampling = {{
    timeseries = {{
        nsamples = {nsamples};
        nwarmup = {nwarmup};
    }}
    convergence = {{
        nsteps = 8;
        nsamples = {nconv};
    }}
}}
prior = {{
    pdemodel = "shiftedlaplace_fd";
    correlationlengthmodel = "constant";
}}
constantcorrelationlengthmodel = {{
    Lambda = 0.2;
}}
periodiccorrelationlengthmodel = {{
    Lambda_min = 0.2;
    Lambda_max = 0.4;
}}
measurements = {{
    radius = 0.0;
    sample_location = [{sx}, {sy}];
    variance_scaling = {vscale};
    measure_global = false;
    mean_global = 1.0;
    variance_global = 0.01;
    filename = "{mfile}";
}}
"""


def write_measurements(path, nmeas):
    loc, sample, mean, var = w.measurement_set(nmeas)
    with open(path, "w") as f:
        f.write("// measurement set of python/generate_measurements.py (seeds 2154157 / 2513267 / 2511541)\n")
        f.write("dim = 2;\nn = %d;\n" % nmeas)
        f.write("measurement_locations = [%s];\n" % ", ".join(repr(float(v)) for v in loc.ravel()))
        f.write("mean = [%s];\n" % ", ".join(repr(float(v)) for v in mean))
        f.write("variance = [%s];\n" % ", ".join(repr(float(v)) for v in var))
    return sample


def main():
    s8 = write_measurements(os.path.join(HERE, "measurements_8.cfg"), 8)
    s32 = write_measurements(os.path.join(HERE, "measurements_32.cfg"), 32)
    cases = [
        ("c1_mgmc_64.cfg", dict(title="C1: driver_mgmc, 64x64, 3 levels, prior, SSOR V(1,1)", n=64, nlevel=3, operator="prior", do_ssor="true",
                                nsmooth=1, maxiter=100, nsamples=10000, nwarmup=1000, nconv=200, sx=0.5, sy=0.5, vscale=1.0, mfile="measurements_8.cfg")),
        ("c2_mg_1024.cfg", dict(title="C2: driver_mg, 1024x1024, 6 levels, V(2,2) SSOR", n=1024, nlevel=6, operator="prior", do_ssor="false",
                                nsmooth=2, maxiter=100, nsamples=100, nwarmup=10, nconv=10, sx=0.5, sy=0.5, vscale=1.0, mfile="measurements_8.cfg")),
        ("c3_mgmc_4096.cfg", dict(title="C3: driver_mgmc, 4096x4096, 8 levels, posterior with 32 point measurements", n=4096, nlevel=8,
                                  operator="posterior", do_ssor="false", nsmooth=1, maxiter=100, nsamples=1000, nwarmup=100, nconv=20,
                                  sx=repr(float(s32[0])), sy=repr(float(s32[1])), vscale="1.E-6", mfile="measurements_32.cfg")),
        ("small_posterior_128.cfg", dict(title="test case: 128x128 posterior with 8 measurements", n=128, nlevel=4, operator="posterior", do_ssor="true",
                                         nsmooth=1, maxiter=50, nsamples=4000, nwarmup=200, nconv=100, sx=repr(float(s8[0])), sy=repr(float(s8[1])),
                                         vscale="1.E-3", mfile="measurements_8.cfg")),
    ]
    for name, kw in cases:
        with open(os.path.join(HERE, name), "w") as f:
            f.write(TEMPLATE.format(**kw))


if __name__ == "__main__":
    main()
