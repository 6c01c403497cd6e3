// Gibbs noise of the small levels generated AHEAD of the launches that consume it.
//
// The noise of a site is a pure function of (seed, chain, sample, level, sweep, site) (philox.cuh) -- it does not depend
// on the iterate.  The launches of the small levels (everything below ~512 x 512) are latency bound: a warp has one row
// per colour pass, and a pass is the in-order latency of ONE Philox + Box-Muller + update chain (~1400 cycles, 80 % of it
// the normals) times 7 live passes.  So while the big levels of the cycle run, a second branch of the cycle graph
// generates the normals of all launches of the small levels of this cycle (a few hundred thousand pairs, a few
// microseconds of the chip, parked in L2: 16 bytes per pair of same-colour sites of an aligned group of 4 columns, one
// plane per live colour pass), and those launches (fused_smooth_kernel<..., NZG = true>) read their normals -- one
// pass ahead, into registers -- instead of generating them: same counters, same function (normal_pair), hence the same
// chain bit for bit (reference: the noise term of SORSampler::apply, sor_sampler.cc:42-46).
//
// Tried first and dropped (profiles/r02_noise_ahead.md): the same for the BIG levels, generated in the idle issue slots
// of the small ones.  The generator runs at ~140 pairs of normals per ns whatever the occupancy (fp64 pipe), level 0
// needs ~90 us of the whole chip per launch, the co-resident small-level launches slow down 2 x, and the level-0 tiles
// do not get faster: without the generator their colour passes are bound by shared-memory bandwidth instead.
#pragma once
#include "fused.cuh"

namespace mgmc {

struct NzJob {
  double2 *buf;         // plane of one colour pass: [row index][gp] pairs, group p at p + kNzgPad
  int colour;
  uint32_t c1;          // (level << 24) | sweep counter of the sweep the pass belongs to
};

constexpr int kMaxNzJobs = 16;
constexpr int kNzGenThreads = 512;

struct NzGenP {
  NoiseP nz;
  int nx, ny, gp;
  int njobs;
  NzJob job[kMaxNzJobs];
};

// NI = independent Philox / Box-Muller chains per warp in flight; THREADS / MINB: block size and blocks per SM the
// register budget is derived from (tools/micro/noise_gen_bench.cu measures the variants)
template <int NC, int NI = 2, int THREADS = kNzGenThreads, int MINB = 2>
__global__ void __launch_bounds__(THREADS, MINB) noise_gen_kernel(const __grid_constant__ NzGenP P) {
  __shared__ __align__(16) double ntab[128];
  if (threadIdx.x < 128) ntab[threadIdx.x] = kNormalTabDev[threadIdx.x];
  __syncthreads();
  const uint32_t sample0 = *P.nz.sample;
  const int lane = threadIdx.x & 31;
  const int G = (int)P.nz.G;
  const int CH = (G + 31) / 32;                              // chunks of 32 groups per row
  const int RJ = (NC == 2) ? (P.ny - 1) : (P.ny / 2);        // rows of a colour pass (4 colours: every other row)
  // (32-bit work-item arithmetic: lattices are limited to ~46000^2 by the site counter, 16 jobs x rows x chunks < 2^31)
  const int per_job = RJ * CH;
  const int total = per_job * P.njobs;
  const int nwarp = (int)gridDim.x * (THREADS / 32);
  const int w0 = (int)blockIdx.x * (THREADS / 32) + (int)(threadIdx.x >> 5);
  for (int u0 = w0; u0 < total; u0 += NI * nwarp) {
    // (both chains are evaluated unconditionally -- surplus work items recompute a valid one -- so that the compiler
    //  interleaves them; only the stores are predicated)
    double z0[NI], z1[NI];
    double2 *dst[NI];
    bool ok[NI];
#pragma unroll
    for (int k = 0; k < NI; ++k) {
      int u = u0 + k * nwarp;
      ok[k] = u < total;
      if (!ok[k]) u = u0;
      const int jb = u / per_job;
      const int rem = u - jb * per_job;
      const int r = rem / CH;
      int pg = (rem - r * CH) * 32 + lane;
      const int colour = P.job[jb].colour;
      int j = (NC == 2) ? (1 + r) : (2 * r + ((colour >> 1) ? 1 : 2));
      ok[k] = ok[k] && (j <= P.ny - 1) && (pg < G);
      j = min(j, P.ny - 1);
      pg = min(pg, G - 1);
      const uint32_t q = (NC == 2) ? (uint32_t)((colour ^ j) & 1) : (uint32_t)(colour & 1);
      normal_pair(P.nz.keys, (((uint32_t)j * (uint32_t)G + (uint32_t)pg) << 1) | q, P.job[jb].c1, sample0, P.nz.chain0, P.nz.mc, ntab,
                  z0[k], z1[k]);
      dst[k] = P.job[jb].buf + ((long long)((NC == 4) ? (j >> 1) : j) * P.gp + (pg + kNzgPad));
    }
#pragma unroll
    for (int k = 0; k < NI; ++k)
      if (ok[k]) *dst[k] = make_double2(z0[k], z1[k]);  // (stays in L2 for the launches that read it)
  }
}

}  // namespace mgmc
