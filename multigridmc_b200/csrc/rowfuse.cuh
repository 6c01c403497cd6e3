// Colour passes of one ROW CLASS in one launch (sm_100a, fp64) for the operator families that sweep colour by colour
// (radius-2 operators: 9 colours (i % 3) + 3 (j % 3); 3d 27-point operators: 8 colours (i & 1) + 2 (j & 1) + 4 (k & 1);
// 2d per-vertex 9-point operators: 4 colours (i & 1) + 2 (j & 1)).
//
// Sites of the colours that share the row part of the colour index -- (., j % 3), (., j & 1, k & 1), (., j & 1) -- lie in the
// same lattice rows, and a stencil couples two such sites only if they lie in the SAME row (a different row within the
// stencil radius has a different row class).  Consecutive colour passes of one row class therefore only depend on each
// other through the row they sit in: one CTA per row runs them back to back with a CTA barrier in between, in place, and
// the rows of the class proceed independently.  A 9-colour SSOR step (17 live passes for omega = 1) becomes 5 launches --
// (0 1 2) (3 4 5) (6 7 | 8 7 6) (5 4 3) (2 1 0): the passes of the last row class of the forward sweep and of the first row class of
// the backward sweep share a launch -- an 8-colour one 7 instead of 15, a 4-colour one 3 instead of 7.  Same ordering, same
// Philox counters, same arithmetic per site as the one-launch-per-colour kernels (kernels.cuh, lattice3d.cuh,
// varcoef.cuh): the chain does not change (tests/test_gpu_invariance.py, MGMC_NO_ROWFUSE=1).
#pragma once
#include "kernels.cuh"
#include "lattice3d.cuh"
#include "varcoef.cuh"

namespace mgmc {

constexpr int kRowPassMax = 6;  // (6 7 8 | 8 7 6) is the longest run of one row class
struct RowPasses {
  int n;
  int ci[kRowPassMax];       // column part of the colour of pass p
  uint32_t c1[kRowPassMax];  // Philox counter word 1 of the sweep pass p belongs to ((level << 24) | sweep counter)
};

// radius-2 operators (position classes, kernels.cuh stencil25): row class cj = j % 3, rows j = jfirst, jfirst + 3, ...
template <bool GIBBS, bool PRE>
__global__ void __launch_bounds__(256) sweep_rows25_kernel(GridP g, const double *__restrict__ st, double *x, const double *__restrict__ f, int jfirst, double omega,
                                                           NoiseP nz, RowPasses P) {
  const int j = jfirst + 3 * (int)blockIdx.x;
  if (j >= g.ny) return;
  const long long orow = (long long)blockIdx.z * g.stride + (long long)j * g.pitch;
  const int cy = 3 * pos_class_dev(j, g.ny);
  if (PRE) {
    // short rows (one site per thread and pass): right-hand sides + noise of ALL passes first -- they do not depend on x, so the
    // generator's latency is paid once, with the passes' chains in flight together, instead of once per pass between the barriers
    double bb[kRowPassMax];
#pragma unroll
    for (int p = 0; p < kRowPassMax; ++p) {
      bb[p] = 0.0;
      const int ci = P.ci[p < P.n ? p : 0];
      const int i = ((ci == 0) ? 3 : ci) + 3 * (int)threadIdx.x;
      if (p < P.n && i < g.nx) {
        const long long o = orow + i;
        double b = f[o];
        if (GIBBS) {
          const double diag = st[25 * (pos_class_dev(i, g.nx) + cy) + 12];
          double z0, z1;
          normal_pair(nz.keys, (((uint32_t)j * nz.G + (uint32_t)(i >> 2)) << 1) | (uint32_t)(i & 1), P.c1[p], *nz.sample, nz.chain0 + blockIdx.z, nz.mc, kNormalTabDev, z0, z1);
          b = fma(sqrt(diag * (2. - omega) / omega), (i & 2) ? z1 : z0, b);
        }
        bb[p] = b;
      }
    }
#pragma unroll
    for (int p = 0; p < kRowPassMax; ++p) {
      if (p < P.n) {
        const int ci = P.ci[p];
        const int i = ((ci == 0) ? 3 : ci) + 3 * (int)threadIdx.x;
        if (i < g.nx) {
          const long long o = orow + i;
          const double *a = st + 25 * (pos_class_dev(i, g.nx) + cy);
          x[o] = x[o] + omega * (bb[p] - stencil25(a, x + o, g.pitch)) / a[12];
        }
        __syncthreads();
      }
    }
    return;
  }
  for (int p = 0; p < P.n; ++p) {
    const int ci = P.ci[p];
    for (int i = ((ci == 0) ? 3 : ci) + 3 * (int)threadIdx.x; i < g.nx; i += 3 * (int)blockDim.x) {
      const long long o = orow + i;
      const double *a = st + 25 * (pos_class_dev(i, g.nx) + cy);
      const double diag = a[12];
      double b = f[o];
      if (GIBBS) {
        double z0, z1;
        normal_pair(nz.keys, (((uint32_t)j * nz.G + (uint32_t)(i >> 2)) << 1) | (uint32_t)(i & 1), P.c1[p], *nz.sample, nz.chain0 + blockIdx.z, nz.mc, kNormalTabDev, z0, z1);
        b = fma(sqrt(diag * (2. - omega) / omega), (i & 2) ? z1 : z0, b);  // sor_sampler.cc:24-27
      }
      x[o] = x[o] + omega * (b - stencil25(a, x + o, g.pitch)) / diag;
    }
    __syncthreads();
  }
}

// 3d 27-point operators (lattice3d.cuh): row class (cj, ck), rows (j, k) = (j0 + 2 blockIdx.x, k0 + 2 blockIdx.y)
template <bool GIBBS, bool PRE>
__global__ void __launch_bounds__(256) sweep_rows27_kernel(GridP g, Grid3 q, Coef27 c, double *x, const double *__restrict__ f, int j0, int k0, double omega, NoiseP nz,
                                                           RowPasses P) {
  const int j = j0 + 2 * (int)blockIdx.x, k = k0 + 2 * (int)blockIdx.y;
  if (j >= q.ny || k >= q.nz) return;
  const int J = k * (q.ny + 1) + j;
  const long long orow = (long long)blockIdx.z * g.stride + (long long)J * g.pitch;
  const long long plane = (long long)(q.ny + 1) * g.pitch;
  const double diag = c.a[13];
  if (PRE) {  // (see sweep_rows25_kernel)
    double bb[kRowPassMax];
#pragma unroll
    for (int p = 0; p < kRowPassMax; ++p) {
      bb[p] = 0.0;
      const int i = (P.ci[p < P.n ? p : 0] ? 1 : 2) + 2 * (int)threadIdx.x;
      if (p < P.n && i < g.nx) {
        double b = f[orow + i];
        if (GIBBS) {
          double z0, z1;
          normal_pair(nz.keys, (((uint32_t)J * nz.G + (uint32_t)(i >> 2)) << 1) | (uint32_t)(i & 1), P.c1[p], *nz.sample, nz.chain0 + blockIdx.z, nz.mc, kNormalTabDev, z0, z1);
          b = fma(sqrt(diag * (2. - omega) / omega), (i & 2) ? z1 : z0, b);
        }
        bb[p] = b;
      }
    }
#pragma unroll
    for (int p = 0; p < kRowPassMax; ++p) {
      if (p < P.n) {
        const int i = (P.ci[p] ? 1 : 2) + 2 * (int)threadIdx.x;
        if (i < g.nx) {
          const long long o = orow + i;
          x[o] = x[o] + omega * (bb[p] - stencil27<true>(c, x + o, g.pitch, plane)) / diag;
        }
        __syncthreads();
      }
    }
    return;
  }
  for (int p = 0; p < P.n; ++p) {
    for (int i = (P.ci[p] ? 1 : 2) + 2 * (int)threadIdx.x; i < g.nx; i += 2 * (int)blockDim.x) {
      const long long o = orow + i;
      double b = f[o];
      if (GIBBS) {
        double z0, z1;
        normal_pair(nz.keys, (((uint32_t)J * nz.G + (uint32_t)(i >> 2)) << 1) | (uint32_t)(i & 1), P.c1[p], *nz.sample, nz.chain0 + blockIdx.z, nz.mc, kNormalTabDev, z0, z1);
        b = fma(sqrt(diag * (2. - omega) / omega), (i & 2) ? z1 : z0, b);
      }
      x[o] = x[o] + omega * (b - stencil27<true>(c, x + o, g.pitch, plane)) / diag;
    }
    __syncthreads();
  }
}

// 2d per-vertex 9-point operators (varcoef.cuh), 4 colours: row class cj = j & 1, rows j = j0, j0 + 2, ...
template <bool NINE, bool GIBBS, bool PRE>
__global__ void __launch_bounds__(256) sweep_rows9v_kernel(GridP g, VarCoef vc, double *x, const double *__restrict__ f, int j0, double omega, NoiseP nz, RowPasses P) {
  const int j = j0 + 2 * (int)blockIdx.x;
  if (j >= g.ny) return;
  const long long orow0 = (long long)j * g.pitch;
  const long long orow = (long long)blockIdx.z * g.stride + orow0;
  if (PRE) {  // (see sweep_rows25_kernel)
    double bb[kRowPassMax];
#pragma unroll
    for (int p = 0; p < kRowPassMax; ++p) {
      bb[p] = 0.0;
      const int i = (P.ci[p < P.n ? p : 0] ? 1 : 2) + 2 * (int)threadIdx.x;
      if (p < P.n && i < g.nx) {
        double b = f[orow + i];
        if (GIBBS) {
          const double diag = diag9v<NINE>(vc, orow0 + i);
          double z0, z1;
          normal_pair(nz.keys, (((uint32_t)j * nz.G + (uint32_t)(i >> 2)) << 1) | (uint32_t)(i & 1), P.c1[p], *nz.sample, nz.chain0 + blockIdx.z, nz.mc, kNormalTabDev, z0, z1);
          b = fma(sqrt(diag * (2. - omega) / omega), (i & 2) ? z1 : z0, b);
        }
        bb[p] = b;
      }
    }
#pragma unroll
    for (int p = 0; p < kRowPassMax; ++p) {
      if (p < P.n) {
        const int i = (P.ci[p] ? 1 : 2) + 2 * (int)threadIdx.x;
        if (i < g.nx) {
          const long long o = orow0 + i, oc = orow + i;
          x[oc] = x[oc] + omega * (bb[p] - stencil9v<NINE>(vc, o, x + oc, g.pitch)) / diag9v<NINE>(vc, o);
        }
        __syncthreads();
      }
    }
    return;
  }
  for (int p = 0; p < P.n; ++p) {
    for (int i = (P.ci[p] ? 1 : 2) + 2 * (int)threadIdx.x; i < g.nx; i += 2 * (int)blockDim.x) {
      const long long o = orow0 + i, oc = orow + i;
      const double diag = diag9v<NINE>(vc, o);
      double b = f[oc];
      if (GIBBS) {
        double z0, z1;
        normal_pair(nz.keys, (((uint32_t)j * nz.G + (uint32_t)(i >> 2)) << 1) | (uint32_t)(i & 1), P.c1[p], *nz.sample, nz.chain0 + blockIdx.z, nz.mc, kNormalTabDev, z0, z1);
        b = fma(sqrt(diag * (2. - omega) / omega), (i & 2) ? z1 : z0, b);
      }
      x[oc] = x[oc] + omega * (b - stencil9v<NINE>(vc, o, x + oc, g.pitch)) / diag;
    }
    __syncthreads();
  }
}

}  // namespace mgmc
