// Persistent kernel for the latency-bound tail of a multigrid cycle: every level below ~512 x 512 and the
// coarsest-level solve run as PHASES of ONE cooperative launch, separated by grid-wide barriers instead of
// kernel boundaries (multigridmc_sampler.cc:103-130 / multigrid_preconditioner.cc:74-101 below the big levels).
//
// Why: on these levels a launch moves < 2 MB and is pure latency -- kernel start, cold instruction cache, the first
// touch of every table -- 20-27 us per launch and 13 launches per cycle (profiles/r01_v5_summary.md).  Here the CTAs
// stay resident, the code stays in the instruction cache, and a phase boundary costs one L2 round trip.
//
// A phase is one of
//   TAIL_FUSED   the tile jobs of one fused smoothing launch (fused_tile, fused.cuh): [prolongation] colour passes
//                [residual + restriction], low-rank fix-ups by the owner / consumer protocol between the tiles
//   TAIL_COARSE  x = A^{-1} f + L^{-T} xi on the coarsest level in ONE pass over the chip: the two dependent
//                triangular products of CholeskySampler::apply (cholesky_sampler.hh:50-66), x = L^{-T}(xi + L^{-1} f),
//                are merged through A^{-1} = L^{-T} L^{-1} (formed on the host at set-up); CholeskySolver::apply
//                (cholesky_solver.cc:30-41) is the same without the noise term
//   TAIL_COPY / TAIL_ZERO   x_primary = x / x = 0 (bookkeeping between the sweeps of the recursion)
// The phase table travels in the kernel parameters (__grid_constant__, up to 32 KB since CUDA 12.1): nothing to
// upload, and a CUDA graph of the cycle carries it inside the kernel node.
#pragma once
#include "fused.cuh"

namespace mgmc {

enum { TAIL_FUSED = 0, TAIL_COARSE = 1, TAIL_COPY = 2, TAIL_ZERO = 3 };

struct CoarseP {
  const double *Ainv;  // A^{-1}, Np x Np row-major (symmetric)
  const double *TT;    // L^{-T}, Np x Np row-major (upper triangular)
  int N, Np, w;        // unknowns, padded row length, interior vertices per lattice row
  int pitch;
  int h, prow;         // 3d lattices (planes stacked in the row direction): interior rows per plane, rows between planes; 2d: INT_MAX, 0
  long long stride;
  const double *f;  // padded lattice layout
  double *x;
  int stage;        // ensembles: the matrix rows of a CTA are staged in shared memory once and reused for every chain
  const double *xi_pre;  // ensembles: the normals of all chains, generated once by coarse_xi_kernel ([chain][Np]); nullptr:
                         // every CTA generates the normals of the chains of a batch itself (one chain: cheaper than a launch)
};

// padded-layout offset of unknown e of the coarsest level (lexicographic: lattice2d.hh:96-103, lattice3d.hh:122-135)
__device__ __forceinline__ long long coarse_site(const CoarseP &C, int e) {
  const int r = e / C.w;
  return (long long)((r / C.h) * C.prow + (r % C.h) + 1 + C.prow) * C.pitch + (e % C.w + 1);
}

struct TailPhase {
  int kind;
  int nc;      // TAIL_FUSED: colours of the level (2 / 4)
  int ntiles;  // TAIL_FUSED: tile jobs per chain
  int c1;      // TAIL_COARSE: Philox word (level << 24 | sweep counter) of the coarse sampler; nc = 1: CholeskySampler
               // (adds L^{-T} xi), nc = 0: CholeskySolver
  FusedP P;    // TAIL_FUSED: the launch; TAIL_COPY / TAIL_ZERO: P.g, P.x_in (source), P.x_out (destination)
};

constexpr int kMaxTailPhases = 18;

struct TailP {
  int nphase, nchains;
  unsigned long long *bar;  // grid barrier: monotonic arrival counter (a multiple of gridDim.x between launches)
  long long *stamps;        // optional: globaltimer at the start and after every phase (CTA 0), for the per-phase profile
  int prefetch_coarse;      // a coarse phase follows later in this launch
  long long *cta_stamps;    // MGMC_TILE_TIMING builds: [phase][CTA][4] globaltimer at phase start / before / after the barrier
  CoarseP coarse;
  NoiseP nz;  // keys / constants of the coarse sampler
  TailPhase ph[kMaxTailPhases];
};
static_assert(sizeof(TailP) <= 32000, "phase table must fit into the kernel parameter space");

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// Grid-wide barrier of a cooperative launch (all CTAs resident).  Writes of every thread of the CTA before the
// barrier are visible to every thread of every CTA after it: bar.sync orders them before thread 0's gpu-scope fence
// and arrival; the acquire load that sees the last arrival orders thread 0 -- and through the second bar.sync the
// whole CTA -- after them.
__device__ __forceinline__ void grid_barrier(unsigned long long *bar, unsigned long long &target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += gridDim.x;
    __threadfence();
    atomicAdd(bar, 1ull);
    while (ld_acquire_u64(bar) < target) {
    }
  }
  __syncthreads();
}

// x = A^{-1} f (+ L^{-T} xi): rows dealt out to the CTAs, one warp per (row, matrix) pair, vectors in shared memory.
// Chains are processed in batches of kCoarseBatch: a matrix element is loaded once and used for every chain of the batch.
constexpr int kCoarseBatch = 4;
constexpr int kCoarseBatchEnsemble = 8;  // chains per matrix pass when there are at least that many (64 accumulator registers)
inline int coarse_batch(int nchains) { return nchains >= kCoarseBatchEnsemble ? kCoarseBatchEnsemble : (nchains < kCoarseBatch ? nchains : kCoarseBatch); }
inline size_t coarse_phase_smem(int Np, int N, int ncta, int nchains, bool stage = false) {
  const int nb = coarse_batch(nchains);
  const size_t rows = (size_t)((N + ncta - 1) / ncta);
  return ((size_t)2 * nb * Np + (size_t)2 * nb * rows + 2 + (stage ? 2 * rows * Np : 0)) * sizeof(double);
}
// Many chains per launch: streaming the two dense matrices from L2 once per batch of chains is what the phase costs
// (256 chains: 64 x 15 MB).  The 7 rows per matrix a CTA works on are 108 KB: staged in shared memory once, every chain
// of the ensemble reuses them -- same products, same summation order, bit-identical chains.
inline bool coarse_phase_stage(int Np, int N, int ncta, int nchains, size_t smem_max) {
  return nchains > kCoarseBatch && coarse_phase_smem(Np, N, ncta, nchains, true) <= smem_max;
}

// Threads of a tail CTA.  The coarse phase (and the copy / zero phases) work with whatever the launch gives them; launches
// that run tile phases (fused_tile) use kFusedThreads, the block size the tile code is written for.
constexpr int kTailThreads = (kFusedThreads > 512) ? kFusedThreads : 512;

template <bool GIBBS, int NB>
__device__ __forceinline__ void coarse_phase(const CoarseP &C, const NoiseP &nz, bool with_noise, uint32_t c1, int nchains, double *sm, const double *ntab,
                                             long long *dbg = nullptr) {
  const int G = gridDim.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rpc = (C.N + G - 1) / G;
  const int r0 = blockIdx.x * rpc, r1 = min(r0 + rpc, C.N);
  if (r0 >= r1) return;  // (uniform over the CTA; the grid barrier is outside)
  const bool sample = GIBBS && with_noise;
  const int Np = C.Np, nbmax = min(nchains, NB);
  double *fv = sm, *xi = sm + (size_t)nbmax * Np, *part = xi + (size_t)nbmax * Np;  // [nb][Np], [nb][Np], [tasks][nb]
  const int nseg = sample ? 2 : 1;
  const int ntask = (r1 - r0) * nseg;
  double *Mst = part + (size_t)2 * nbmax * rpc + 2;  // [row - r0][seg][Np] (C.stage)
  if (C.stage) {
    for (int task = warp; task < ntask; task += (int)(blockDim.x >> 5)) {
      const int row = r0 + task / nseg, seg = task % nseg;
      const double *__restrict__ M = (seg ? C.TT : C.Ainv) + (long long)row * Np;
      double *dst = Mst + (size_t)task * Np;
      for (int c = lane; c < Np; c += 32) dst[c] = M[c];
    }
    // (published by the barrier behind the vector loads of the first batch)
  }
  for (int ch0 = 0; ch0 < nchains; ch0 += NB) {
    const int nb = min(NB, nchains - ch0);
    for (int idx = threadIdx.x; idx < nb * Np; idx += (int)blockDim.x) {
      const int b = idx / Np, e = idx - b * Np;
      fv[idx] = (e < C.N) ? C.f[(long long)(ch0 + b) * C.stride + coarse_site(C, e)] : 0.0;
    }
    if (sample && C.xi_pre) {
      for (int idx = threadIdx.x; idx < nb * Np; idx += (int)blockDim.x) xi[idx] = C.xi_pre[(size_t)ch0 * Np + idx];
    } else if (sample) {
      // xi_row: Philox counter 0x40000000 | row / 2 (philox.cuh), normal = (row & 1) ? z1 : z0
      const int hp = Np / 2;
      for (int idx = threadIdx.x; idx < nb * hp; idx += (int)blockDim.x) {
        const int b = idx / hp, p = idx - b * hp;
        double z0 = 0.0, z1 = 0.0;
        if (2 * p < C.N) normal_pair(nz.keys, 0x40000000u | (uint32_t)p, c1, *nz.sample, nz.chain0 + ch0 + b, nz.mc, ntab, z0, z1);
        xi[b * Np + 2 * p] = z0;
        xi[b * Np + 2 * p + 1] = (2 * p + 1 < C.N) ? z1 : 0.0;
      }
    }
    __syncthreads();
#ifdef MGMC_TILE_TIMING
    if (dbg && threadIdx.x == 0) dbg[0] = gtimer();
#endif
    for (int task = warp; task < ntask; task += (int)(blockDim.x >> 5)) {
      const int row = r0 + task / nseg, seg = task % nseg;
      const double *__restrict__ M = C.stage ? (Mst + (size_t)task * Np) : ((seg ? C.TT : C.Ainv) + (long long)row * Np);
      const double *v = seg ? xi : fv;
      int c = (seg ? (row & ~31) : 0) + lane;  // L^{-T} is upper triangular: row `row` starts at column `row`
      // (the summation order of a chain must not depend on the batch: chains are compared bit for bit with the same
      //  chain run on its own -- four partial sums per chain, element c + 32 k into partial sum k mod 4)
      double acc[NB][4];
#pragma unroll
      for (int b = 0; b < NB; ++b) acc[b][0] = acc[b][1] = acc[b][2] = acc[b][3] = 0.0;
      for (; c + 224 < Np; c += 256) {  // 8 independent loads in flight per lane
        double t[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) t[k] = M[c + 32 * k];
#pragma unroll
        for (int b = 0; b < NB; ++b)
          if (b < nb) {
            const double *vb = v + b * Np + c;
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[b][k & 3] = fma(t[k], vb[32 * k], acc[b][k & 3]);
          }
      }
      for (; c < Np; c += 32) {
        const double t0 = M[c];
#pragma unroll
        for (int b = 0; b < NB; ++b)
          if (b < nb) acc[b][0] = fma(t0, v[b * Np + c], acc[b][0]);
      }
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        double a = (acc[b][0] + acc[b][1]) + (acc[b][2] + acc[b][3]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0 && b < nb) part[task * nb + b] = a;
      }
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < (r1 - r0) * nb; idx += (int)blockDim.x) {
      const int r = idx / nb, b = idx - r * nb, row = r0 + r;
      const double v = part[(r * nseg) * nb + b] + (sample ? part[(r * nseg + 1) * nb + b] : 0.0);
      C.x[(long long)(ch0 + b) * C.stride + coarse_site(C, row)] = v;
    }
    __syncthreads();
  }
}

// Ensembles: the normals of the coarse sampler for all chains, once (inside the coarse phase every CTA would generate
// the normals of all chains again: 148-fold).  Same counters and pairing as in coarse_phase.
__global__ void __launch_bounds__(256) coarse_xi_kernel(NoiseP nz, uint32_t c1, int N, int Np, int nchains, double *__restrict__ xi) {
  __shared__ __align__(16) double ntab[128];
  if (threadIdx.x < 128) ntab[threadIdx.x] = kNormalTabDev[threadIdx.x];
  __syncthreads();
  const int hp = Np / 2;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < nchains * hp; idx += gridDim.x * blockDim.x) {
    const int b = idx / hp, p = idx - b * hp;
    double z0 = 0.0, z1 = 0.0;
    if (2 * p < N) normal_pair(nz.keys, 0x40000000u | (uint32_t)p, c1, *nz.sample, nz.chain0 + b, nz.mc, ntab, z0, z1);
    xi[(size_t)b * Np + 2 * p] = z0;
    xi[(size_t)b * Np + 2 * p + 1] = (2 * p + 1 < N) ? z1 : 0.0;
  }
}

template <bool GIBBS, bool LOWRANK>
__global__ void __launch_bounds__(kTailThreads, 1) tail_kernel(const __grid_constant__ TailP T) {
  extern __shared__ double sm[];
  __shared__ int lr_cnt[4];
  __shared__ __align__(16) double ntab[128];
  // the description of the current phase, copied from the parameter space once per phase: the tile code reads it all the
  // time, and a constant-bank access with a run-time phase index costs an indexed LDC (several of them in a dependent
  // chain per colour pass) where shared memory costs an LDS
  __shared__ __align__(16) unsigned long long Pbuf[(sizeof(FusedP) + 7) / 8];
  const FusedP &Ps = *reinterpret_cast<const FusedP *>(Pbuf);
  if (GIBBS && threadIdx.x < 128) ntab[threadIdx.x] = kNormalTabDev[threadIdx.x];
  unsigned long long target = 0;
  if (threadIdx.x == 0) target = (ld_acquire_u64(T.bar) / gridDim.x) * gridDim.x;
  if (T.stamps && blockIdx.x == 0 && threadIdx.x == 0) T.stamps[0] = gtimer();
  __syncthreads();
  const int G = gridDim.x;
  if (T.prefetch_coarse) {
    // The coarse matrices were last read a whole cycle (several hundred MB of traffic) ago: pull the rows of this CTA
    // back into L2 now, while the first phases run
    const CoarseP &C = T.coarse;
    const int rpc = (C.N + G - 1) / G;
    const int r0 = blockIdx.x * rpc, r1 = min(r0 + rpc, C.N);
    const int lines = (C.Np * 8 + 127) / 128;
    for (int k = threadIdx.x; k < (r1 - r0) * lines * 2; k += (int)blockDim.x) {
      const int mat = k / ((r1 - r0) * lines), rem = k - mat * (r1 - r0) * lines;
      const int row = r0 + rem / lines, ln = rem % lines;
      if (mat == 1 && (ln + 1) * 16 <= (row & ~31)) continue;  // (zero part of the upper triangle)
      const char *ptr = reinterpret_cast<const char *>((mat ? C.TT : C.Ainv) + (long long)row * C.Np) + (long long)ln * 128;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
    }
  }
  for (int p = 0; p < T.nphase; ++p) {
    const TailPhase &ph = T.ph[p];
#ifdef MGMC_TILE_TIMING
    if (T.cta_stamps && threadIdx.x == 0) T.cta_stamps[((long long)p * G + blockIdx.x) * 4 + 0] = gtimer();
#endif
    if (ph.kind == TAIL_FUSED) {
      {
        const unsigned long long *src = reinterpret_cast<const unsigned long long *>(&ph.P);
        for (int k = threadIdx.x; k < (int)(sizeof(FusedP) / 8); k += (int)blockDim.x) Pbuf[k] = src[k];
        __syncthreads();
      }
      // whole chains per wave: the owner / consumer exchange of the low-rank fix-ups needs every tile of a chain resident
      const int nt = ph.ntiles;
      const int cpw = max(1, G / nt);
      const int slot = blockIdx.x;
      for (int c0 = 0; c0 < T.nchains; c0 += cpw) {
        const int chain = c0 + slot / nt, tile = slot % nt;
        if (slot < cpw * nt && chain < T.nchains) {
          if (ph.nc == 2) fused_tile<2, GIBBS, 2, 2, LOWRANK>(Ps, tile, chain, nt, sm, lr_cnt, ntab);
          else fused_tile<4, GIBBS, 2, 2, LOWRANK>(Ps, tile, chain, nt, sm, lr_cnt, ntab);
        }
        __syncthreads();
      }
    } else if (ph.kind == TAIL_COARSE) {
#ifdef MGMC_TILE_TIMING
      coarse_phase<GIBBS, kCoarseBatch>(T.coarse, T.nz, ph.nc != 0, (uint32_t)ph.c1, T.nchains, sm, ntab, T.cta_stamps ? T.cta_stamps + ((long long)p * G + blockIdx.x) * 4 + 3 : nullptr);
#else
      if (T.nchains >= kCoarseBatchEnsemble) coarse_phase<GIBBS, kCoarseBatchEnsemble>(T.coarse, T.nz, ph.nc != 0, (uint32_t)ph.c1, T.nchains, sm, ntab);
      else coarse_phase<GIBBS, kCoarseBatch>(T.coarse, T.nz, ph.nc != 0, (uint32_t)ph.c1, T.nchains, sm, ntab);
#endif
    } else {
      // x_out = x_in (TAIL_COPY) or x_out = 0 (TAIL_ZERO) on the interior
      const GridP &g = ph.P.g;
      const long long nrow = (long long)(g.ny - 1) * T.nchains;
      for (long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < nrow; r += (long long)G * (blockDim.x >> 5)) {
        const long long o = (r / (g.ny - 1)) * g.stride + (r % (g.ny - 1) + 1) * g.pitch;
        for (int i = 1 + (threadIdx.x & 31); i < g.nx; i += 32) ph.P.x_out[o + i] = (ph.kind == TAIL_COPY) ? ph.P.x_in[o + i] : 0.0;
      }
    }
#ifdef MGMC_TILE_TIMING
    __syncthreads();
    if (T.cta_stamps && threadIdx.x == 0) T.cta_stamps[((long long)p * G + blockIdx.x) * 4 + 1] = gtimer();
#endif
    if (p + 1 < T.nphase) grid_barrier(T.bar, target);
#ifdef MGMC_TILE_TIMING
    if (T.cta_stamps && threadIdx.x == 0) T.cta_stamps[((long long)p * G + blockIdx.x) * 4 + 2] = gtimer();
#endif
    if (T.stamps && blockIdx.x == 0 && threadIdx.x == 0) T.stamps[p + 1] = gtimer();
  }
}

}  // namespace mgmc
