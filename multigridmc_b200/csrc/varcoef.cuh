// Radius-1 operators with PER-VERTEX coefficients (sm_100a, fp64): ShiftedLaplaceFDOperator with a correlation length
// that varies in space (PeriodicCorrelationLengthModel, linear_operator/correlationlength_model.hh:83-113: kappa^2(x) on
// the diagonal of the fine matrix, shiftedlaplace_fd_operator.cc:33-56) and its Galerkin coarsenings R A R^T
// (linear_operator.cc:12-15), whose nine coefficients all vary from vertex to vertex.
//
// Layout: nine coefficient planes in the padded layout of the level vectors, a[k * plane + j * pitch + i] with
// k = (dj + 1) * 3 + (di + 1); boundary / ghost vertices hold zeros, and so do the entries that point to a boundary
// vertex.  The planes are shared by all chains.  On the fine level only the diagonal varies (kappa^2(x)): the 5-point
// operator (NINE = false) reads ONE plane, the diagonal, and takes the four neighbour coefficients from the kernel
// parameters -- towards the boundary they multiply the zero ghost lines of x, as in the constant-coefficient kernels.
//
// First correct path for this operator family (as the radius-2 kernels in kernels.cuh): one launch per colour, in
// place -- sites of one colour do not couple (red-black for the 5-point fine operator, 4 colours for the 9-point
// Galerkin operators: the orderings of the constant-coefficient tile kernel, so the chain does not depend on which
// path runs) -- and separate transfer kernels.  HBM-bound: a colour pass moves 8 * (1 or 9 coefficients) + 24 bytes per site.
#pragma once
#include "kernels.cuh"

namespace mgmc {

struct VarCoef {
  const double *a;     // origin (vertex i = 0, j = 0) of plane 0 (NINE) / of the diagonal plane (5-point fine operator)
  long long plane;     // doubles between planes
  double w, e, s, n;   // 5-point fine operator: the constant neighbour coefficients
};

template <bool NINE>
__device__ __forceinline__ double diag9v(const VarCoef &vc, long long o) {
  return NINE ? vc.a[4 * vc.plane + o] : vc.a[o];
}

template <bool NINE>
__device__ __forceinline__ double stencil9v(const VarCoef &vc, long long o, const double *__restrict__ p, int pitch) {
  const double *a = vc.a + o;
  if (!NINE) {
    double s = a[0] * p[0];
    s = fma(vc.w, p[-1], s);
    s = fma(vc.e, p[1], s);
    s = fma(vc.s, p[-pitch], s);
    s = fma(vc.n, p[pitch], s);
    return s;
  }
  double s = a[4 * vc.plane] * p[0];
  s = fma(a[3 * vc.plane], p[-1], s);
  s = fma(a[5 * vc.plane], p[1], s);
  s = fma(a[1 * vc.plane], p[-pitch], s);
  s = fma(a[7 * vc.plane], p[pitch], s);
  {
    s = fma(a[0 * vc.plane], p[-pitch - 1], s);
    s = fma(a[2 * vc.plane], p[-pitch + 1], s);
    s = fma(a[6 * vc.plane], p[pitch - 1], s);
    s = fma(a[8 * vc.plane], p[pitch + 1], s);
  }
  return s;
}

// y = A_0 x (LinearOperator::apply sparse part, linear_operator.hh:69) or r = f - A_0 x
template <bool NINE, bool RESIDUAL>
__global__ void __launch_bounds__(256) apply9v_kernel(GridP g, VarCoef vc, const double *__restrict__ x, const double *__restrict__ f, double *__restrict__ y) {
  const int i = 1 + blockIdx.x * 64 + threadIdx.x;
  const int j = 1 + blockIdx.y * 4 + threadIdx.y;
  if (i >= g.nx || j >= g.ny) return;
  const long long o = (long long)j * g.pitch + i;
  const long long oc = (long long)blockIdx.z * g.stride + o;
  const double s = stencil9v<NINE>(vc, o, x + oc, g.pitch);
  y[oc] = RESIDUAL ? (f[oc] - s) : s;
}

// r = A_0 x - b with per-block partial sums of r^2 (LoopSolver, loop_solver.cc:26-28)
template <bool NINE>
__global__ void __launch_bounds__(256) residual_norm9v_kernel(GridP g, VarCoef vc, const double *__restrict__ x, const double *__restrict__ b, double *__restrict__ r,
                                                              double *__restrict__ partial) {
  const int i = 1 + blockIdx.x * 64 + threadIdx.x;
  const int j = 1 + blockIdx.y * 4 + threadIdx.y;
  double v = 0.0;
  if (i < g.nx && j < g.ny) {
    const long long o = (long long)j * g.pitch + i;
    v = stencil9v<NINE>(vc, o, x + o, g.pitch) - b[o];
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
    for (int k = 0; k < 8; ++k) s += ws[k];
    partial[blockIdx.y * gridDim.x + blockIdx.x] = s;
  }
}

// One colour of a SOR / Gibbs sweep (sor_smoother.cc:41-78, sor_sampler.cc:37-58 in the multicolour ordering).
// nc = 2: red-black, colour = (i + j) & 1, every row holds sites of the colour (jstep = 1);
// nc = 4: 4 colours, colour = (i & 1) + 2 (j & 1), the rows j = j0, j0 + 2, ... (jstep = 2).
// (NINE only selects how the coefficients are stored: a degenerate coarse lattice with a single interior line has a
//  9-plane operator without corner couplings, which sweeps red-black.)
// The noise of a site is the same pure function of the site as in every other sweep kernel (philox.cuh).
template <bool NINE, bool GIBBS>
__global__ void __launch_bounds__(256) sweep_colour9v_kernel(GridP g, VarCoef vc, double *__restrict__ x, const double *__restrict__ f, int colour, int nc, double omega,
                                                             NoiseP nz, int j0, int jstep) {
  const int j = j0 + jstep * (blockIdx.y * 4 + threadIdx.y);
  if (j >= g.ny) return;
  const int ipar = (nc == 4) ? (colour & 1) : ((colour ^ j) & 1);  // parity of the columns of this colour in row j
  const int i = (ipar ? 1 : 2) + 2 * (blockIdx.x * 64 + threadIdx.x);
  if (i >= g.nx) return;
  const long long o = (long long)j * g.pitch + i;
  const long long oc = (long long)blockIdx.z * g.stride + o;
  const double diag = diag9v<NINE>(vc, o);
  double b = f[oc];
  if (GIBBS) {
    double z0, z1;
    normal_pair(nz.keys, (((uint32_t)j * nz.G + (uint32_t)(i >> 2)) << 1) | (uint32_t)(i & 1), nz.c1, *nz.sample, nz.chain0 + blockIdx.z, nz.mc, kNormalTabDev, z0, z1);
    b = fma(sqrt(diag * (2. - omega) / omega), (i & 2) ? z1 : z0, b);  // sor_sampler.cc:24-27
  }
  x[oc] = x[oc] + omega * (b - stencil9v<NINE>(vc, o, x + oc, g.pitch)) / diag;
}

}  // namespace mgmc
