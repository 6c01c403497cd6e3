// Gibbs noise of the big levels generated AHEAD of the launches that consume it.
//
// The noise of a site is a pure function of (seed, chain, sample, level, sweep, site) (philox.cuh) -- it does not depend
// on the iterate.  The colour passes of the big levels are bound by the fp64 pipe, and 3/4 of their instructions are
// Philox + Box-Muller; the small levels of the cycle (everything below ~512 x 512 and the coarse solve, ~25 % of the
// cycle time) are latency bound and leave > 80 % of the chip idle.  So while the small levels run, a second branch of
// the cycle graph fills those idle issue slots with the normals of the NEXT fine-level launches -- the post-smoothing
// of this cycle and the pre-smoothing of the next one (sample index + 1) -- and parks them in HBM (16 bytes per pair
// of same-colour sites of an aligned group of 4 columns, one plane per live colour pass).  The fine-level launches
// (fused_smooth_kernel<..., NZG = true>) then read their normals instead of generating them: same counters, same
// function (normal_pair), hence the same chain bit for bit (reference: the noise term of SORSampler::apply,
// sor_sampler.cc:42-46).
//
// The kernel is persistent with one CTA per SM (the dynamic shared memory it claims keeps a second one off the SM) at
// 64 registers per thread, so that the one CTA per SM of a small-level launch still finds its registers and shared
// memory.  A plane is valid for one sample index: the TAG of a job list records it, and the same kernel launched with
// only_if_stale at the head of the cycle regenerates a plane that is not the one the cycle is about to read (first
// cycle of a context, sample index moved by the API).
#pragma once
#include "fused.cuh"

namespace mgmc {

struct NzJob {
  double2 *buf;         // plane of one colour pass: [row index][gp] pairs, group p at p + kNzgPad
  int colour;
  uint32_t c1;          // (level << 24) | sweep counter of the sweep the pass belongs to
  uint32_t sample_off;  // the plane is for sample index *sample + sample_off
};

constexpr int kMaxNzJobs = 16;
constexpr int kNzGenThreads = 512;
constexpr int kNzGenSmem = 116 * 1024;  // claimed, not used: one CTA per SM

struct NzGenP {
  NoiseP nz;
  int nc, nx, ny, gp;
  int njobs;
  NzJob job[kMaxNzJobs];
  uint32_t *tag;        // sample index the tagged planes hold (nullptr: none of the jobs is tagged)
  uint32_t tag_off;     // the tagged planes are generated for *sample + tag_off
  int only_if_stale;    // return at once if *tag == *sample + tag_off
  unsigned int *ticket; // arrival counter of the CTAs: the last one writes the tag (nullptr: this launch does not write it)
};

// NI = independent Philox / Box-Muller chains per warp in flight; THREADS / MINB: block size and blocks per SM the
// register budget is derived from (tools/micro/noise_gen_bench.cu measures the variants)
template <int NC, int NI = 2, int THREADS = kNzGenThreads, int MINB = 2>
__global__ void __launch_bounds__(THREADS, MINB) noise_gen_kernel(const __grid_constant__ NzGenP P) {
  __shared__ __align__(16) double ntab[128];
  __shared__ uint32_t s_sample;
  if (threadIdx.x == 0) s_sample = *P.nz.sample;
  if (threadIdx.x < 128) ntab[threadIdx.x] = kNormalTabDev[threadIdx.x];
  __syncthreads();
  const uint32_t sample0 = s_sample;
  // (the tag is only written after every CTA has arrived at the end of the kernel, i.e. after every CTA has read it)
  if (P.only_if_stale && P.tag && *reinterpret_cast<volatile uint32_t *>(P.tag) == sample0 + P.tag_off) return;
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
      normal_pair(P.nz.keys, (((uint32_t)j * (uint32_t)G + (uint32_t)pg) << 1) | q, P.job[jb].c1, sample0 + P.job[jb].sample_off, P.nz.chain0, P.nz.mc, ntab,
                  z0[k], z1[k]);
      dst[k] = P.job[jb].buf + ((long long)((NC == 4) ? (j >> 1) : j) * P.gp + (pg + kNzgPad));
    }
#pragma unroll
    for (int k = 0; k < NI; ++k)
      if (ok[k]) __stcs(dst[k], make_double2(z0[k], z1[k]));
  }
  if (P.tag && P.ticket) {
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(P.ticket, 1u) == gridDim.x - 1) {
      *P.ticket = 0u;
      *P.tag = sample0 + P.tag_off;
    }
  }
}

}  // namespace mgmc
