#!/usr/bin/env python
"""bench.py -- MGMC samples/s on the BASELINE workload:
2d shifted-Laplace 4096 x 4096, MeasuredOperator with 32 point measurements, 8 levels,
SSOR V(1,1) Gibbs smoothing, dense Cholesky sampler on the coarsest level, fp64.

A "step" is one MGMC sample (one V-cycle over the whole hierarchy).  Contract: see the task
statement; one JSON line on stdout from rank 0.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--n 4096] [--nlevel 8]

N > 1 (launched under torchrun, one rank per GPU): ONE chain of the benchmark lattice on row strips over all ranks
("large lattices are domain-decomposed" of the north star; strong scaling: `value` = samples/s of that chain, halo
rows exchanged inside the tile kernel over NVLink peer memory).  The other mode of the north star -- independent
chains per GPU, no data-path collective, weak scaling -- is timed in the same run and reported under
`independent_chains`; `--decomp chains` makes it the headline instead.

`--impl reference`: the CPU oracle (restatement of the reference; the reference itself needs Eigen 3.4 + libconfig++,
absent here) on the SAME workload -- 4096 x 4096, 8 levels, 32 measurements -- for --steps samples after --warmup,
single-threaded like the reference (one sequential RNG stream, lexicographic Gauss-Seidel); "time per sample" as
driver_mgmc.cc:72-80 defines it.  Needs ~30 GB of host memory (dense n x m low-rank matrices, sor_smoother.cc:17-38);
on a smaller host the largest lattice that fits is run and named in config.workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

NPROF = 11  # cycles of the per-kernel profile: a run of K cycles is K - 1 merged level-0 launches + one split pair

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=200)
    p.add_argument("--warmup", type=int, default=5)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--n", type=int, default=4096)
    p.add_argument("--nlevel", type=int, default=8)
    p.add_argument("--nmeas", type=int, default=32)
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-batched", action="store_true", help="skip the 4-chains-per-launch figure")
    p.add_argument("--decomp", default="strips", choices=["chains", "strips"],
                   help="N > 1: ONE chain on row strips of the lattice (strong scaling, default) or independent chains per GPU (weak scaling)")
    p.add_argument("--cpu-n", type=int, default=0, help="lattice of the CPU arm (0: the benchmark lattice if the host memory allows)")
    p.add_argument("--cpu-samples", type=int, default=2, help="samples of the in-line cpu_baseline (bounded sample of the workload)")
    p.add_argument("--no-configs", action="store_true", help="skip the sub-records of the other BASELINE configurations")
    return p.parse_args()


def workload_config(a):
    return {
        "workload": f"driver_mgmc: 2D shifted Laplace {a.n}x{a.n} with MeasuredOperator ({a.nmeas} point measurements), "
                    f"{a.nlevel} levels, SSOR V(1,1), coarse Cholesky sampler",
        "lattice": [a.n, a.n], "nlevel": a.nlevel, "m_lowrank": a.nmeas, "smoother": "SSOR", "npresmooth": 1, "npostsmooth": 1,
        "cycle": 1, "omega": 1.0, "Lambda": 0.2, "variance_scaling": 1e-6,
        "l2_policy": "inputs larger than L2 (x and f of the fine level are 2 x 134 MB > 126 MB L2); no flush needed",
    }


def cycle_model(n, nlevel):
    """algorithmic bytes / site updates per V(1,1)-SSOR cycle"""
    sites = [(n // 2 ** l - 1) ** 2 for l in range(nlevel)]
    s = sum(sites[:-1])
    return 134.0 * s + 8.0 * sites[-1] ** 2, 4.0 * s + sites[-1]


# ------------------------------------------------------------------------------------ CPU arm
def host_info():
    info = {"cores_available": os.cpu_count()}
    try:
        import cpuinfo

        info["cpu"] = cpuinfo.get_cpu_info().get("brand_raw")
    except Exception:
        pass
    try:
        import psutil

        info["mem_available_gb"] = round(psutil.virtual_memory().available / 1e9, 1)
    except Exception:
        pass
    return info


def cpu_lattice(a):
    """The lattice the CPU arm runs: the benchmark lattice itself unless the host cannot hold the reference's data
    structures (4 dense n x m matrices per level: ~1.8 KB per fine-level unknown at m = 32 -> 30 GB at 4096^2)."""
    n, nlevel = a.n, a.nlevel
    if a.cpu_n:
        while n > a.cpu_n and nlevel > 2:
            n, nlevel = n // 2, nlevel - 1
        return n, nlevel
    try:
        import psutil

        avail = psutil.virtual_memory().available
    except Exception:
        avail = 64e9
    need = lambda n_: 1.25 * (30e9 * (n_ / 4096.0) ** 2 * max(a.nmeas, 4) / 32.0)
    while need(n) > avail and nlevel > 2:
        n, nlevel = n // 2, nlevel - 1
    return n, nlevel


def cpu_reference_run(a, nsamples, nwarm):
    """Times the CPU oracle (faithful restatement of the reference: lexicographic sweeps, std::mt19937_64, CSR Galerkin
    hierarchy, dense n x m low-rank correction; one thread) on the benchmark workload: MultigridMCSampler::apply in the
    loop of measure_sampling_time (driver_mgmc.cc:66-80), "time per sample" = elapsed / nsamples."""
    from multigridmc_b200 import workloads as w
    from oracle import oracle as orc

    n, nlevel = cpu_lattice(a)
    loc, sample_loc, mean, var = w.measurement_set(a.nmeas) if a.nmeas else (None, np.array([0.5, 0.5]), None, None)
    t0 = time.time()
    prior = orc.Operator.prior((n, n), "shiftedlaplace_fd", Lambda=0.2)
    op = prior.measured(loc, var, variance_scaling=1e-6) if a.nmeas else prior
    H = orc.Hierarchy(op, nlevel, orc.LEX)
    rng = orc.StdRng(5418513)
    sampler = H.mgmc(rng=rng, smoother="SSOR", coarse_solver="Cholesky", npresmooth=1, npostsmooth=1, cycle=1, omega=1.0)
    t_setup = time.time() - t0
    xs = np.arange(1, n) / n
    u = np.outer(np.sin(np.pi * xs), np.sin(np.pi * xs)).ravel()
    f = op.apply(u)
    b_obs = op.measurement_vector(sample_loc, 0.0)
    x = np.zeros(op.ndof)
    if nwarm > 0:
        x, _ = sampler.run(f, x, b_obs, nwarm)
    t0 = time.time()
    x, series = sampler.run(f, x, b_obs, nsamples)
    dt = time.time() - t0
    _, upd_run = cycle_model(n, nlevel)
    _, upd_full = cycle_model(a.n, a.nlevel)
    same = (n, nlevel) == (a.n, a.nlevel)
    updates_per_s = upd_run * nsamples / dt
    value = nsamples / dt  # samples/s on the lattice that was run
    value_full = value if same else updates_per_s / upd_full
    what = (f"{nsamples} MGMC samples (after {nwarm} warm-up) of the workload itself: {n}x{n}, {nlevel} levels, m={a.nmeas}, "
            f"time per sample {1e3 * dt / nsamples:.0f} ms (driver_mgmc.cc:72-80)" if same else
            f"{nsamples} MGMC samples (after {nwarm} warm-up) on {n}x{n}, {nlevel} levels, m={a.nmeas} -- the largest lattice of the "
            f"family this host's memory holds; site-updates/s converted to samples/s of {a.n}x{a.n}")
    return {
        "value": value, "value_full_lattice": value_full, "same_config": same, "lattice": [n, n], "nlevel": nlevel,
        "site_updates_per_s": updates_per_s, "ms_per_sample": 1e3 * dt / nsamples,
        "sample": what + f"; set-up {t_setup:.0f} s untimed (threaded over the measurements; sampling on 1 thread)",
        "seconds": dt, "setup_seconds": t_setup, "host": host_info(),
    }


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference_run(a, max(a.steps, 1), max(a.warmup, 0))
    cfg = workload_config(a)
    if not r["same_config"]:
        n, nl = r["lattice"][0], r["nlevel"]
        cfg["workload"] = cfg["workload"].replace(f"{a.n}x{a.n}", f"{n}x{n}").replace(f"{a.nlevel} levels", f"{nl} levels")
        cfg["lattice"], cfg["nlevel"] = r["lattice"], nl
    line = {
        "impl": "reference", "metric": "mgmc_samples_per_sec", "value": r["value"], "unit": "samples/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": r["ms_per_sample"],
        # (same key as the B200 arm prints for this N: one chain of one lattice -- strong -- when the lattice is decomposed)
        "higher_is_better": True, "scaling": "strong" if (a.gpus > 1 and a.decomp == "strips") else "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": cfg,
        "site_updates_per_sec": r["site_updates_per_s"],
        "cpu_baseline": {"value": r["value"], "unit": "samples/s", "cores": 1, "kind": "port", "sample": r["sample"], "host": r["host"]},
        "e2e": {"value": r["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference is single-threaded by construction (one sequential RNG stream, lexicographic Gauss-Seidel); "
                "the reference itself needs Eigen 3.4 + libconfig++ (absent): this is the oracle port, timed per sample like driver_mgmc.cc:72-80",
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------ GPU arm
class ClockSampler:
    def __init__(self, device):
        self.device = device
        self.rows = []
        self.proc = None

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.device}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                smax = float(r[1])
                for k, nm in enumerate(names):
                    if r[3 + k].lower().startswith("active"):
                        reasons.add(nm)
            except (ValueError, IndexError):
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


def other_configs(m, peak):
    """Driver-run numbers for the other BASELINE configurations (C1, C2, C4, C5) on one GPU, through
    the C ABI, device-timed -- sub-records of the bench line, not bench lines of their own."""
    out = []

    def mgmc(name, n, nlevel, nchains=1, pde="shiftedlaplace_fd", steps=100, target=None, kappa_sq=None, nz=None):
        try:
            ctx = m.Context(n, n, nlevel, Lambda=0.2, pde=pde, nchains=nchains, kappa_sq=kappa_sq, nz=nz)
            nd = ctx.ndof()
            xs = np.arange(1, n) / n
            u = np.outer(np.sin(np.pi * xs), np.sin(np.pi * xs)).ravel()
            if nz is not None:
                u = np.outer(np.sin(np.pi * np.arange(1, nz) / nz), u).ravel()
            ctx.set_rhs(ctx.op_apply(0, np.tile(u, nchains)))
            ctx.set_state(np.zeros(nd * nchains))
            ctx.set_qoi([nd // 2], [1.0])
            ctx.sample(10, series=False)
            ms, _ = ctx.sample_timed(steps)
            byts, upd = ctx.cycle_model()
            t = ms / steps * 1e-3
            rec = {"config": name, "ms_per_cycle": ms / steps, "chain_samples_per_s": nchains / t, "site_updates_per_s": nchains * upd / t,
                   "algorithmic_gbs": nchains * byts / t / 1e9, "frac_of_peak": nchains * byts / t / 1e9 / peak}
            if target:
                rec["target"] = target
            ctx.close()
        except m.MgmcError as e:
            rec = {"config": name, "unavailable": str(e)}
        out.append(rec)

    mgmc("C1 driver_mgmc 64x64, 3 levels, prior, V(1,1) SSOR (latency bound: us per cycle, not a roofline fraction)", 64, 3, steps=1000)
    try:
        n, nlevel = 1024, 6
        ctx = m.Context(n, n, nlevel, Lambda=0.2, npresmooth=2, npostsmooth=2)
        b = np.random.default_rng(1482817).standard_normal(ctx.ndof())  # (driver_mg.cc:165-172 draws it from std::mt19937_64(1482817))
        ctx.loop_solve(b, rtol=1e-12, atol=1e-15, maxiter=3)
        t0 = time.perf_counter()
        x, hist, it, cv = ctx.loop_solve(b, rtol=1e-12, atol=1e-15, maxiter=100)
        dt = time.perf_counter() - t0
        out.append({"config": "C2 driver_mg 1024x1024, 6 levels, V(2,2) SSOR, rtol 1e-12 / atol 1e-15 / maxiter 100", "iterations": len(hist),
                    "ms_per_iteration": 1e3 * dt / len(hist), "residual_reduction": float(hist[-1] / hist[0]),
                    "rate_first_10": float((hist[10] / hist[0]) ** 0.1), "timed": "wall clock of the whole LoopSolver call incl. H2D of b and D2H of x",
                    "target": "<= 0.09 ms / iteration at 60 % of the model"})
        ctx.close()
    except Exception as e:  # noqa: BLE001
        out.append({"config": "C2", "unavailable": str(e)})
    mgmc("C4 squared_shiftedlaplace_fd 2048x2048, 7 levels, prior, V(1,1) SSOR, one GPU", 2048, 7, pde="squared_shiftedlaplace_fd", steps=20,
         target="<= 0.193 ms / cycle at 60 % of the model")
    mgmc("C5 256 chains x 512x512, 5 levels, prior, one GPU", 512, 5, nchains=256, steps=20, target=">= 8.4e4 chain-samples/s per GPU at 60 % of the model")
    # not a BASELINE configuration: the correlation length model of the reference's own sampler / solver tests
    # (PeriodicCorrelationLengthModel), per-vertex coefficients on every level, one launch per colour (csrc/varcoef.cuh)
    mgmc("periodic correlation length (Lambda 0.1 .. 0.4), 2048x2048, 7 levels, prior, V(1,1) SSOR, one GPU (per-vertex coefficients: first correct path)",
         2048, 7, steps=20, kappa_sq=m.periodic_kappa_sq(2048, 2048, 0.1, 0.4))
    # not a BASELINE configuration either: a 3d lattice (Lattice3d; 7-point fine / 27-point Galerkin operators, red-black / 8-colour
    # sweeps, one launch per colour, csrc/lattice3d.cuh)
    mgmc("3d lattice 128x128x128, 5 levels, prior, V(1,1) SSOR, one GPU (first correct path)", 128, 5, steps=20, nz=128)
    return out


def run_b200(a):
    import torch

    import multigridmc_b200 as m
    from multigridmc_b200 import workloads as w

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the MGMC path has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    n, nlevel = a.n, a.nlevel
    loc, sample_loc, mean, var = w.measurement_set(a.nmeas) if a.nmeas else (None, np.array([0.5, 0.5]), None, None)
    B = w.point_measurement_matrix(n, n, loc, var, 1e-6) if a.nmeas else None
    strips_on = world > 1 and a.decomp == "strips"
    if strips_on:
        from multigridmc_b200 import strips

        ctx = m.Context(n, n, nlevel, Lambda=0.2, B=B, smoother="SSOR", coarse_solver="Cholesky", npresmooth=1, npostsmooth=1,
                        cycle=1, omega=1.0, seed=5418513, device=local, nchains=1, first_chain=0, strip_rank=rank, strip_nranks=world)
        strips.connect(ctx, dist, torch.device("cuda", local))
    else:
        ctx = m.Context(n, n, nlevel, Lambda=0.2, B=B, smoother="SSOR", coarse_solver="Cholesky", npresmooth=1, npostsmooth=1,
                        cycle=1, omega=1.0, seed=5418513, device=local, nchains=1, first_chain=rank)
    nchains_total = 1 if strips_on else world
    strip_err = 0
    nd = ctx.ndof()
    # synthetic right-hand side f = A u, u = sin(pi x) sin(pi y); pinned host buffers for the e2e path
    xs = np.arange(1, n) / n
    u = np.outer(np.sin(np.pi * xs), np.sin(np.pi * xs)).ravel()
    f_pin = torch.empty(nd, dtype=torch.float64).pin_memory()
    x_pin = torch.zeros(nd, dtype=torch.float64).pin_memory()
    f_np, x_np = f_pin.numpy(), x_pin.numpy()
    f_np[:] = ctx.op_apply(0, u)
    ctx.set_rhs(f_np)
    ctx.set_state(x_np)
    ctx.set_qoi([w.nearest_vertex(n, n, sample_loc)], [1.0])
    ctx.set_philox_position(0)

    # ---- device-resident throughput ("value"): inputs already in HBM, CUDA-graph replay ----
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()  # nvidia-smi needs a few 100 ms to come up: start it before the warm-up
    ctx.sample(a.warmup, series=False)
    launches0 = ctx.launch_count()
    barrier()
    ms, series = ctx.sample_timed(a.steps, series=True)
    barrier()
    clk = clocks.stop() if rank == 0 else None
    launches = ctx.launch_count() - launches0
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = nchains_total * a.steps / (ms_max * 1e-3)
    if strips_on:
        series = strips.reduce_series(series, dist, torch.device("cuda", local))
        strip_err = ctx.strip_error()
        if strip_err:
            raise SystemExit("bench.py: a device-side wait for a neighbour rank timed out")

    # ---- end to end through the reference-facing call with HOST buffers ----
    # MultigridMCSampler::apply(f, x) after fix_rhs(f) (driver_mgmc.cc:65,75): per step H2D of the chain
    # state x from pinned memory, one cycle, D2H of x; QoI evaluated on the host like the driver does.
    import ctypes as C
    L = m.lib()
    c_dp = C.POINTER(C.c_double)
    xp = x_np.ctypes.data_as(c_dp)
    e2e_steps = max(3, min(a.steps, 20))
    qidx = w.nearest_vertex(n, n, sample_loc)
    e2e_value = None
    if not strips_on:  # (row strips: the chain state is distributed; only the device-resident loop below applies)
        for _ in range(2):
            assert L.mgmc_sampler_mgmc_apply(ctx.h, None, xp) == 0
        barrier()
        t0 = time.perf_counter()
        acc = 0.0
        for _ in range(e2e_steps):
            assert L.mgmc_sampler_mgmc_apply(ctx.h, None, xp) == 0
            acc += x_np[qidx]
        barrier()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_value = nchains_total * e2e_steps / float(t.item())

    # ---- the same metric through the device-resident form of the driver's hot loop (Sampler::sample_series,
    #      host/mgmc_host.hh): chain state stays in HBM, only the QoI series crosses PCIe ----
    ctx.set_state(x_np)
    barrier()
    t0 = time.perf_counter()
    series_res = ctx.sample(a.steps, series=True)
    barrier()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_resident = nchains_total * a.steps / float(t.item())

    # ---- N > 1: the other multi-GPU mode of the north star next to the headline one.  Headline = row strips (default):
    #      additionally every rank advances an independent chain (Philox chain id = rank; no data-path collective, weak
    #      scaling).  Headline = chains (--decomp chains): additionally ONE chain on row strips. ----
    strips_extra = None
    chains_extra = None
    if world > 1 and strips_on:
        try:
            cctx = m.Context(n, n, nlevel, Lambda=0.2, B=B, smoother="SSOR", coarse_solver="Cholesky", npresmooth=1, npostsmooth=1,
                             cycle=1, omega=1.0, seed=5418513, device=local, nchains=1, first_chain=rank)
            cctx.set_rhs(f_np)
            cctx.set_state(np.zeros(nd))
            cctx.set_philox_position(0)
            cctx.sample(a.warmup, series=False)
            barrier()
            cms, _ = cctx.sample_timed(a.steps, series=False)
            barrier()
            t = torch.tensor([cms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            chains_extra = {"value": world * a.steps / (float(t.item()) * 1e-3), "unit": "samples/s", "ms_per_step": float(t.item()) / a.steps,
                            "scaling": "weak", "what": "every rank advances an independent chain of the same lattice (Philox chain id = rank), no data-path collective"}
            cctx.close()
        except m.MgmcError as e:
            chains_extra = {"unavailable": str(e)}
    if world > 1 and not strips_on:
        from multigridmc_b200 import strips as _strips

        try:
            sctx = m.Context(n, n, nlevel, Lambda=0.2, B=B, smoother="SSOR", coarse_solver="Cholesky", npresmooth=1, npostsmooth=1,
                             cycle=1, omega=1.0, seed=5418513, device=local, nchains=1, first_chain=0, strip_rank=rank, strip_nranks=world)
            _strips.connect(sctx, dist, torch.device("cuda", local))
            sctx.set_rhs(f_np)
            sctx.set_state(np.zeros(nd))
            sctx.set_philox_position(0)
            barrier()
            sctx.sample(a.warmup, series=False)
            barrier()
            sms, _ = sctx.sample_timed(a.steps, series=False)
            barrier()
            t = torch.tensor([sms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            strips_extra = {"value": a.steps / (float(t.item()) * 1e-3), "unit": "samples/s", "ms_per_step": float(t.item()) / a.steps, "scaling": "strong",
                            "error_flag": int(sctx.strip_error()),
                            "what": "ONE chain of the same lattice on row strips over all ranks (bit-identical to the single-GPU chain)"}
            sctx.close()
        except m.MgmcError as e:
            strips_extra = {"unavailable": str(e)}

    # ---- N = 1: the same workload with several chains per launch (blockIdx.z).  Levels 3-7 are latency bound, so
    #      further chains ride along at almost no cost there: throughput of the sampler when the user wants more than
    #      one chain (independent chains per GPU of the north star) -- reported next to `value`, not instead of it ----
    batched = None
    if world == 1 and not a.no_batched:
        try:
            nb = 4
            bctx = m.Context(n, n, nlevel, Lambda=0.2, B=B, smoother="SSOR", coarse_solver="Cholesky", npresmooth=1, npostsmooth=1,
                             cycle=1, omega=1.0, seed=5418513, device=local, nchains=nb, first_chain=0)
            bctx.set_rhs(np.tile(f_np, nb))
            bctx.set_state(np.zeros(nd * nb))
            bctx.set_qoi([qidx], [1.0])
            bctx.set_philox_position(0)
            bctx.sample(max(2, a.warmup // 2), series=False)
            bsteps = max(10, a.steps // 4)
            bms, _ = bctx.sample_timed(bsteps, series=False)
            batched = {"chains": nb, "value": nb * bsteps / (bms * 1e-3), "unit": "chain-samples/s", "ms_per_step": bms / bsteps,
                       "what": "same lattice and measurements, 4 independent chains advanced by every launch"}
            bctx.close()
        except m.MgmcError as e:
            batched = {"unavailable": str(e)}

    prof = None
    if strips_on:  # cooperative: every rank has to run the profiled cycles
        barrier()
        prof = ctx.profile_cycle(nsamples=NPROF)
        barrier()
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- per-kernel CUDA-event timing of the cycle (rank 0): roofline of the dominant kernel ----
    if prof is None:
        prof = ctx.profile_cycle(nsamples=NPROF)
    total_ms = sum(p[1] for p in prof)
    top = max(prof, key=lambda p: p[1])
    byts, upd = ctx.cycle_model()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (of fallback)"
    # algorithmic bytes per launch of the dominant kernel: counted by the library itself in the model of
    # SURVEY.md section 8(d) (24 B per site and sweep, 18 B prolongate_add, 18 + 2 B residual + restrict),
    # see DESIGN.md "Kernels"; the duration is the CUDA-event average over the launches of that slot
    per_launch = top[3] / top[2] if top[3] > 0 else None
    avg_ms = top[1] / top[2]
    achieved = (per_launch / (avg_ms * 1e-3) / 1e9) if per_launch else None
    traffic = None
    try:  # DRAM bytes per launch of the same kernel from the committed ncu --set full capture
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(top[0])
    except Exception:
        pass
    cycle_gbs = byts / (ms_max / a.steps * 1e-3) / 1e9

    line = {
        "metric": "mgmc_samples_per_sec", "value": value, "unit": "samples/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms_max / a.steps, "higher_is_better": True, "scaling": "strong" if strips_on else "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": dict(workload_config(a), multi_gpu=("row strips of one lattice, halo rows stored into the neighbours' memory over NVLink (CUDA IPC), "
                                                                           "device-side flags, coarse levels replicated" if strips_on else
                                                                           ("independent chains per GPU, no data-path collective" if world > 1 else "single GPU"))),
        "site_updates_per_sec": value * upd,
        "site_updates_note": "reference-equivalent site updates (4 colour-pass equivalents per SSOR step as the reference sweeps; the kernels skip the dead pass between "
                             "a forward and a backward sweep, exact for omega = 1)",
        "cycle_algorithmic_gbs": cycle_gbs, "cycle_roofline_frac": cycle_gbs / peak / (world if strips_on else 1),
        "e2e": ({"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": 8 * nd, "d2h_bytes_per_step": 8 * nd,
                 "call": "mgmc_sampler_mgmc_apply(ctx, NULL /*rhs fixed by fix_rhs*/, x) with pinned host x, QoI read on the host",
                 "steps": e2e_steps} if e2e_value is not None else
                {"value": e2e_resident, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 8,
                 "call": "mgmc_sample(ctx, K, qoi_host) on every rank + sum of the partial QoI series (state distributed over the ranks)"}),
        "e2e_resident": {"value": e2e_resident, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 8,
                         "call": "mgmc_sample(ctx, K, qoi_host): K cycles + device QoI, one D2H of the K-entry series (wall clock, host buffers)"},
        "gpu_launches": int(launches),
        "clocks": clk,
        "roofline": {"bound": "hbm", "kernel": top[0], "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": (achieved / peak) if achieved else None, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": per_launch, "avg_launch_ms": avg_ms,
                     "kernel_share_of_cycle": top[1] / total_ms if total_ms else None},
        "kernels": [{"name": p[0], "ms_per_cycle": p[1] / NPROF, "launches_per_cycle": p[2] / NPROF,
                     "algorithmic_gbs": (p[3] / (p[1] * 1e-3) / 1e9) if p[1] > 0 else None} for p in sorted(prof, key=lambda p: -p[1])[:24]],
        "qoi_mean": float(np.mean(series)),
    }
    if strips_extra is not None:
        line["strips"] = strips_extra
    if chains_extra is not None:
        line["independent_chains"] = chains_extra
    if strips_on:
        line["strips_error_flag"] = int(strip_err)
    if batched is not None:
        line["batched_chains"] = batched
    if world == 1 and not a.no_configs:
        line["other_configs"] = other_configs(m, peak)
    if not a.no_cpu_baseline and world == 1:
        r = cpu_reference_run(a, max(a.cpu_samples, 1), 0)
        line["cpu_baseline"] = {"value": r["value_full_lattice"], "unit": "samples/s", "cores": 1, "kind": "port", "sample": r["sample"],
                                "same_config": r["same_config"], "site_updates_per_sec": r["site_updates_per_s"], "host": r["host"]}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
