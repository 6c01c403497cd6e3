// 3d lattices (sm_100a, fp64): ShiftedLaplaceFDOperator on a Lattice3d (lattice/lattice3d.hh:43-270,
// linear_operator/shiftedlaplace_fd_operator.cc:33-56 with dim = 3) and its Galerkin coarsenings R A R^T with the
// trilinear full weighting (intergrid/intergrid_operator_linear.cc:8-30, linear_operator.cc:12-15).
//
// Layout: the planes k = 0 .. nz of a level are stacked in the row direction of the padded 2d layout of the level
// vectors: vertex (i, j, k) lives in row J = k (ny + 1) + j, column i.  Boundary vertices (i, j or k on the boundary)
// hold zeros and are never written, so a uniform stencil serves every interior vertex (entries that point to the
// boundary multiply zeros) and the lattice-wide helpers of the 2d path (zero, copy, axpy, norms, moments, the observed
// sites, the dense coarse solve) work on these vectors unchanged.  The operator is a uniform radius-1 stencil: 7 points
// on the fine level, 27 on the coarse levels (setup.hh fine_stencil3 / coarsen_stencil3).
//
// First correct path for this lattice family (like the radius-2 and per-vertex kernels): one launch per colour, in
// place -- red-black ((i + j + k) & 1) for the 7-point operator, 8 colours ((i & 1) + 2 (j & 1) + 4 (k & 1)) for the
// 27-point operators -- and separate transfer kernels.  The noise of a site is the same pure function of (row J, column i)
// as in every other sweep kernel (philox.cuh), so the chain does not depend on the launch geometry.
#pragma once
#include "kernels.cuh"

namespace mgmc {

struct Grid3 {
  int ny, nz;  // cells in y and z; rows per plane = ny + 1
};
struct Coef27 {
  double a[27];  // (di, dj, dk) at [(dk + 1) * 9 + (dj + 1) * 3 + (di + 1)]
};

// row J -> (j, k); true for rows that hold interior vertices
__device__ __forceinline__ bool rows3(const Grid3 &q, int J, int &j, int &k) {
  const int pr = q.ny + 1;
  k = J / pr;
  j = J - k * pr;
  return j >= 1 && j < q.ny && k >= 1 && k < q.nz;
}

template <bool FULL>
__device__ __forceinline__ double stencil27(const Coef27 &c, const double *__restrict__ p, int pitch, long long plane) {
  double s = c.a[13] * p[0];
  s = fma(c.a[12], p[-1], s);
  s = fma(c.a[14], p[1], s);
  s = fma(c.a[10], p[-pitch], s);
  s = fma(c.a[16], p[pitch], s);
  s = fma(c.a[4], p[-plane], s);
  s = fma(c.a[22], p[plane], s);
  if (FULL) {
#pragma unroll
    for (int dk = -1; dk <= 1; ++dk)
#pragma unroll
      for (int dj = -1; dj <= 1; ++dj)
#pragma unroll
        for (int di = -1; di <= 1; ++di) {
          if ((di != 0) + (dj != 0) + (dk != 0) < 2) continue;
          s = fma(c.a[(dk + 1) * 9 + (dj + 1) * 3 + (di + 1)], p[dk * plane + dj * pitch + di], s);
        }
  }
  return s;
}

// y = A x (LinearOperator::apply sparse part, linear_operator.hh:69) or r = f - A x
template <bool FULL, bool RESIDUAL>
__global__ void __launch_bounds__(256) apply27_kernel(GridP g, Grid3 q, Coef27 c, const double *__restrict__ x, const double *__restrict__ f, double *__restrict__ y) {
  const int i = 1 + blockIdx.x * 64 + threadIdx.x;
  const int J = 1 + blockIdx.y * 4 + threadIdx.y;
  int j, k;
  if (i >= g.nx || J >= g.ny || !rows3(q, J, j, k)) return;
  const long long o = (long long)blockIdx.z * g.stride + (long long)J * g.pitch + i;
  const double s = stencil27<FULL>(c, x + o, g.pitch, (long long)(q.ny + 1) * g.pitch);
  y[o] = RESIDUAL ? (f[o] - s) : s;
}

// r = A x - b with per-block partial sums of r^2 (LoopSolver, loop_solver.cc:26-28)
template <bool FULL>
__global__ void __launch_bounds__(256) residual_norm27_kernel(GridP g, Grid3 q, Coef27 c, const double *__restrict__ x, const double *__restrict__ b, double *__restrict__ r,
                                                              double *__restrict__ partial) {
  const int i = 1 + blockIdx.x * 64 + threadIdx.x;
  const int J = 1 + blockIdx.y * 4 + threadIdx.y;
  double v = 0.0;
  int j, k;
  if (i < g.nx && J < g.ny && rows3(q, J, j, k)) {
    const long long o = (long long)J * g.pitch + i;
    v = stencil27<FULL>(c, x + o, g.pitch, (long long)(q.ny + 1) * g.pitch) - b[o];
    r[o] = v;
  }
  v = v * v;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  __shared__ double ws[8];
  const int t = threadIdx.y * 64 + threadIdx.x;
  if ((t & 31) == 0) ws[t >> 5] = v;
  __syncthreads();
  if (t == 0) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += ws[w];
    partial[blockIdx.y * gridDim.x + blockIdx.x] = s;
  }
}

// One colour of a SOR / Gibbs sweep (sor_smoother.cc:41-78, sor_sampler.cc:37-58 in the multicolour ordering).
// FULL = false: red-black, colour = (i + j + k) & 1; FULL = true: 8 colours, colour = (i & 1) + 2 (j & 1) + 4 (k & 1).
template <bool FULL, bool GIBBS>
__global__ void __launch_bounds__(256) sweep_colour27_kernel(GridP g, Grid3 q, Coef27 c, double *__restrict__ x, const double *__restrict__ f, int colour, double omega,
                                                             NoiseP nz) {
  const int J = 1 + blockIdx.y * 4 + threadIdx.y;
  int j, k;
  if (J >= g.ny || !rows3(q, J, j, k)) return;
  if (FULL && (((j & 1) != ((colour >> 1) & 1)) || ((k & 1) != ((colour >> 2) & 1)))) return;
  const int ipar = FULL ? (colour & 1) : ((colour ^ j ^ k) & 1);  // parity of the columns of this colour in row J
  const int i = (ipar ? 1 : 2) + 2 * (blockIdx.x * 64 + threadIdx.x);
  if (i >= g.nx) return;
  const long long o = (long long)blockIdx.z * g.stride + (long long)J * g.pitch + i;
  const double diag = c.a[13];
  double b = f[o];
  if (GIBBS) {
    double z0, z1;
    normal_pair(nz.keys, (((uint32_t)J * nz.G + (uint32_t)(i >> 2)) << 1) | (uint32_t)(i & 1), nz.c1, *nz.sample, nz.chain0 + blockIdx.z, nz.mc, kNormalTabDev, z0, z1);
    b = fma(sqrt(diag * (2. - omega) / omega), (i & 2) ? z1 : z0, b);  // sor_sampler.cc:24-27
  }
  x[o] = x[o] + omega * (b - stencil27<FULL>(c, x + o, g.pitch, (long long)(q.ny + 1) * g.pitch)) / diag;
}

// f_c = R r (IntergridOperator::restrict, intergrid_operator.hh:74-88): weights {1/2, 1, 1/2}^(x)3 around the fine vertex 2 I
__global__ void __launch_bounds__(256) restrict27_kernel(GridP g, Grid3 q, GridP gc, Grid3 qc, const double *__restrict__ r, double *__restrict__ fc) {
  const int I = 1 + blockIdx.x * 64 + threadIdx.x;
  const int JC = 1 + blockIdx.y * 4 + threadIdx.y;
  int Jj, K;
  if (I >= gc.nx || JC >= gc.ny || !rows3(qc, JC, Jj, K)) return;
  const long long plane = (long long)(q.ny + 1) * g.pitch;
  const double *p = r + (long long)blockIdx.z * g.stride + (long long)(2 * K) * plane + (long long)(2 * Jj) * g.pitch + 2 * I;
  double s = 0.0;
#pragma unroll
  for (int dk = -1; dk <= 1; ++dk)
#pragma unroll
    for (int dj = -1; dj <= 1; ++dj) {
      const double *pr = p + dk * plane + dj * g.pitch;
      const double w = (dk ? 0.5 : 1.0) * (dj ? 0.5 : 1.0);
      s += w * (0.5 * pr[-1] + pr[0] + 0.5 * pr[1]);
    }
  fc[(long long)blockIdx.z * gc.stride + (long long)JC * gc.pitch + I] = s;
}

// x += alpha R^T x_c in gather form (IntergridOperator::prolongate_add, intergrid_operator.hh:106-120): every fine vertex
// reads its (up to) 8 coarse parents; boundary parents are the zero boundary planes / lines
__global__ void __launch_bounds__(256) prolongate_add27_kernel(GridP g, Grid3 q, GridP gc, Grid3 qc, double alpha, const double *__restrict__ xc, double *__restrict__ x) {
  const int i = 1 + blockIdx.x * 64 + threadIdx.x;
  const int J = 1 + blockIdx.y * 4 + threadIdx.y;
  int j, k;
  if (i >= g.nx || J >= g.ny || !rows3(q, J, j, k)) return;
  const double *c = xc + (long long)blockIdx.z * gc.stride;
  const long long planec = (long long)(qc.ny + 1) * gc.pitch;
  const int I0 = i >> 1, I1 = (i + 1) >> 1, J0 = j >> 1, J1 = (j + 1) >> 1, K0 = k >> 1, K1 = (k + 1) >> 1;
  // even index: the two parents coincide -> weight 1/2 + 1/2 = 1; odd index: the two neighbours with weight 1/2
  double v = 0.0;
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const double *row = c + (long long)(a ? K1 : K0) * planec + (long long)(b ? J1 : J0) * gc.pitch;
      v += row[I0] + row[I1];
    }
  x[(long long)blockIdx.z * g.stride + (long long)J * g.pitch + i] += alpha * (0.125 * v);
}

}  // namespace mgmc
