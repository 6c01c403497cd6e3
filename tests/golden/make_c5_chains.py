#!/usr/bin/env python
"""Generates tests/golden/c5_reference_chains.npz: the REFERENCE chain of BASELINE config 5 -- MultigridMCSampler with
lexicographic SSOR sweeps and one std::mt19937_64 stream per chain (the oracle's restatement of
multigridmc_sampler.cc:103-138 / sor_sampler.cc:37-58) on the 2d shifted Laplacian, 512 x 512, 5 levels, prior --
8 independent chains x 2000 recorded QoI values (x at the vertex nearest (0.5, 0.5)) after 200 warm-up cycles.
tests/test_gpu_benchmark_scale.py compares the statistics of the GPU chains with these.  ~2 min per chain on one core;
the chains run in parallel processes:  python tests/golden/make_c5_chains.py
"""
import multiprocessing as mp
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

N, NLEVEL, NWARM, NSAMPLES, NCHAINS = 512, 5, 200, 2000, 8
SEEDS = [5418513 + 1000 * k for k in range(NCHAINS)]  # chain 0 = the driver's seed (driver_mgmc.cc:448)


def chain(seed):
    from multigridmc_b200 import workloads as w
    from oracle import oracle as orc

    op = orc.Operator.prior((N, N), "shiftedlaplace_fd", Lambda=0.2)
    H = orc.Hierarchy(op, NLEVEL, orc.LEX)
    s = H.mgmc(rng=orc.StdRng(seed), smoother="SSOR", coarse_solver="Cholesky", npresmooth=1, npostsmooth=1, cycle=1, omega=1.0)
    xs = np.arange(1, N) / N
    u = np.outer(np.sin(np.pi * xs), np.sin(np.pi * xs)).ravel()
    f = op.apply(u)
    q = w.nearest_vertex(N, N, [0.5, 0.5])
    b_obs = np.zeros(op.ndof)
    b_obs[q] = 1.0
    x, zw = s.run(f, np.zeros(op.ndof), b_obs, NWARM)
    _, z = s.run(f, x, b_obs, NSAMPLES)
    return q, np.asarray(z), np.asarray(zw)[:3]


if __name__ == "__main__":
    with mp.Pool(min(NCHAINS, os.cpu_count() or 1)) as pool:
        out = pool.map(chain, SEEDS)
    series = np.stack([o[1] for o in out])
    # first_cycles: the QoI of the first three cycles of chain 0 from x = 0 (cheap regression pin, tests/test_golden.py)
    np.savez_compressed(os.path.join(HERE, "c5_reference_chains.npz"), n=N, nlevel=NLEVEL, nwarm=NWARM, seeds=np.array(SEEDS),
                        qoi_index=out[0][0], series=series, first_cycles=out[0][2])
    print("written", series.shape, "mean", series.mean(), "var", series.var(axis=1, ddof=1).mean())
