#!/usr/bin/env python
"""bench.py -- MGMC samples/s on the BASELINE workload:
2d shifted-Laplace 4096 x 4096, MeasuredOperator with 32 point measurements, 8 levels,
SSOR V(1,1) Gibbs smoothing, dense Cholesky sampler on the coarsest level, fp64.

A "step" is one MGMC sample (one V-cycle over the whole hierarchy).  Contract: see the task
statement; one JSON line on stdout from rank 0.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--n 4096] [--nlevel 8]

N > 1 (launched under torchrun, one rank per GPU): every rank advances an independent Markov chain
on the same lattice (Philox chain id = rank) -- "small lattices run independent chains per GPU" of
the north star applied to the benchmark lattice; no data-path collective, weak scaling.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

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
    p.add_argument("--decomp", default="chains", choices=["chains", "strips"],
                   help="N > 1: independent chains per GPU (weak scaling, default) or ONE chain on row strips of the lattice (strong scaling)")
    p.add_argument("--cpu-n", type=int, default=1024, help="lattice of the bounded CPU sample")
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
def cpu_reference_run(a, nsamples, nwarm):
    """Times the CPU oracle (faithful restatement of the reference: lexicographic sweeps,
    std::mt19937_64, CSR Galerkin hierarchy, dense n x m low-rank correction) on a bounded sample:
    the same operator family on a cpu_n x cpu_n lattice with nlevel chosen so that the coarsest level
    is the same 31 x 31; converted to samples/s of the benchmark lattice through site-updates/s."""
    from multigridmc_b200 import workloads as w
    from oracle import oracle as orc

    n = a.cpu_n
    nlevel = a.nlevel - int(round(np.log2(a.n / n)))
    loc, sample_loc, mean, var = w.measurement_set(a.nmeas)
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
    x, _ = sampler.run(f, x, b_obs, nwarm)
    t0 = time.time()
    x, series = sampler.run(f, x, b_obs, nsamples)
    dt = time.time() - t0
    _, upd_small = cycle_model(n, nlevel)
    _, upd_full = cycle_model(a.n, a.nlevel)
    updates_per_s = upd_small * nsamples / dt
    return {
        "samples_per_s_equiv": updates_per_s / upd_full,
        "site_updates_per_s": updates_per_s,
        "ms_per_sample_on_sample_lattice": 1e3 * dt / nsamples,
        "sample": f"{nsamples} MGMC samples (after {nwarm} warm-up) of the same operator on {n}x{n}, {nlevel} levels, m={a.nmeas}; "
                  f"site-updates/s converted to samples/s of the {a.n}x{a.n} workload; setup {t_setup:.1f} s untimed",
        "seconds": dt,
    }


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference_run(a, max(a.steps, 1), max(a.warmup, 0))
    cfg = workload_config(a)
    line = {
        "impl": "reference", "metric": "mgmc_samples_per_sec", "value": r["samples_per_s_equiv"], "unit": "samples/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 / r["samples_per_s_equiv"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
        "site_updates_per_sec": r["site_updates_per_s"],
        "cpu_baseline": {"value": r["samples_per_s_equiv"], "unit": "samples/s", "cores": 1, "kind": "port", "sample": r["sample"]},
        "e2e": {"value": r["samples_per_s_equiv"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference is single-threaded by construction (one sequential RNG stream, lexicographic Gauss-Seidel); "
                "the reference itself needs Eigen 3.4 + libconfig++ (absent): this is the oracle port",
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
        if ctx.strip_error():
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

    # ---- N > 1, default mode (independent chains): additionally time ONE chain on row strips of the same lattice
    #      (strong scaling; halo rows exchanged inside the tile kernel over NVLink peer memory) ----
    strips_extra = None
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
        prof = ctx.profile_cycle(nsamples=3)
        barrier()
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- per-kernel CUDA-event timing of the cycle (rank 0): roofline of the dominant kernel ----
    if prof is None:
        prof = ctx.profile_cycle(nsamples=3)
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
        "kernels": [{"name": p[0], "ms_per_cycle": p[1] / 3.0, "launches_per_cycle": p[2] / 3.0,
                     "algorithmic_gbs": (p[3] / (p[1] * 1e-3) / 1e9) if p[1] > 0 else None} for p in sorted(prof, key=lambda p: -p[1])[:12]],
        "qoi_mean": float(np.mean(series)),
    }
    if strips_extra is not None:
        line["strips"] = strips_extra
    if batched is not None:
        line["batched_chains"] = batched
    if not a.no_cpu_baseline and world == 1:
        r = cpu_reference_run(a, 10, 1)
        line["cpu_baseline"] = {"value": r["samples_per_s_equiv"], "unit": "samples/s", "cores": 1, "kind": "port", "sample": r["sample"],
                                "site_updates_per_sec": r["site_updates_per_s"]}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
