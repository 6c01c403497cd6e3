// Device kernels of the MGMC hot path (sm_100a, fp64, bandwidth-bound stencil work: no tensor cores).
//
// Data layout: every level vector is a padded row-major array addressed as v[j * pitch + i] with
// (i, j) the Euclidean vertex index of the reference's Lattice2d (lattice/lattice2d.hh:96-103),
// i in [0, nx], j in [0, ny]; the Dirichlet boundary lines i = 0, nx and j = 0, ny and two further
// ghost lines on every side hold zeros that are never written, so stencils need no boundary
// branches.  Rows start 128-byte aligned.  Chains are stacked with a fixed stride.
#pragma once
#include <cstdint>

#include "philox.cuh"

namespace mgmc {

// radius-1 stencil with constant coefficients (5-point: corners are zero)
struct Coef9 {
  double c, w, e, s, n, sw, se, nw, ne;
};

struct NoiseP {
  PhiloxKeys keys;        // round keys of the 64-bit seed
  NormalConsts mc;        // full-width constants of the normal generator (philox.cuh)
  uint32_t c1;            // (level << 24) | sweep counter
  const uint32_t *sample; // device-resident sample index (so that CUDA graphs can be replayed)
  uint32_t chain0;        // global id of chain 0
  uint32_t G;             // groups per row = nx / 4 + 1
};

// device-side waits (row-strip flags, low-rank packets) give up after this many SM clocks and raise the context's error word
// instead of hanging the GPU (~3 s at 1.9 GHz)
constexpr long long kWaitTimeoutClocks = 6000000000ll;

struct GridP {
  int nx, ny, pitch;
  long long stride;  // doubles between chains
};

__device__ __forceinline__ int ld_acquire_sys(const int *p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// ------------------------------------------------------------------------------------------------
// y = A_0 x  (LinearOperator::apply sparse part, linear_operator.hh:69) or r = f - A_0 x
// ------------------------------------------------------------------------------------------------
template <bool NINE, bool RESIDUAL>
__global__ void __launch_bounds__(256) apply_kernel(GridP g, Coef9 a, const double *__restrict__ x, const double *__restrict__ f,
                                                   double *__restrict__ y) {
  const int i = 1 + blockIdx.x * 64 + threadIdx.x;
  const int j = 1 + blockIdx.y * 4 + threadIdx.y;
  if (i >= g.nx || j >= g.ny) return;
  const long long o = (long long)blockIdx.z * g.stride + (long long)j * g.pitch + i;
  const double *p = x + o;
  double s = a.c * p[0] + a.w * p[-1] + a.e * p[1] + a.s * p[-g.pitch] + a.n * p[g.pitch];
  if (NINE) s += a.sw * p[-g.pitch - 1] + a.se * p[-g.pitch + 1] + a.nw * p[g.pitch - 1] + a.ne * p[g.pitch + 1];
  y[o] = RESIDUAL ? (f[o] - s) : s;
}

// r = A_0 x - b together with per-block partial sums of r^2 (LoopSolver, loop_solver.cc:26-28)
template <bool NINE>
__global__ void __launch_bounds__(256) residual_norm_kernel(GridP g, Coef9 a, const double *__restrict__ x, const double *__restrict__ b,
                                                           double *__restrict__ r, double *__restrict__ partial) {
  const int i = 1 + blockIdx.x * 64 + threadIdx.x;
  const int j = 1 + blockIdx.y * 4 + threadIdx.y;
  double v = 0.0;
  if (i < g.nx && j < g.ny) {
    const long long o = (long long)j * g.pitch + i;
    const double *p = x + o;
    double s = a.c * p[0] + a.w * p[-1] + a.e * p[1] + a.s * p[-g.pitch] + a.n * p[g.pitch];
    if (NINE) s += a.sw * p[-g.pitch - 1] + a.se * p[-g.pitch + 1] + a.nw * p[g.pitch - 1] + a.ne * p[g.pitch + 1];
    v = s - b[o];
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

// Device-resident control of the Richardson loop (LoopSolver::apply, loop_solver.cc:21-41): the residual norm of every
// evaluation goes into a device-side history and the convergence test "rel < rtol AND abs < atol" (loop_solver.cc:33) is
// made on the device; once it holds the update x -= P r is suppressed, so the host may launch iterations in batches and
// look at the outcome afterwards -- x stays the iterate at which the reference would have stopped.
struct SolverCtl {
  double r0, rtol, atol;
  int maxiter;  // residual evaluations recorded at most
  int iter;     // residual evaluations so far
  int conv;     // converged
  int it_conv;  // iteration index at which the test held
};

// deterministic final reduction of the per-block partial sums of r^2, then history + convergence test
__global__ void __launch_bounds__(1024) reduce_check_kernel(const double *__restrict__ partial, int n, SolverCtl *ctl, double *__restrict__ hist) {
  __shared__ double ws[32];
  double v = 0.0;
  for (int k = threadIdx.x; k < n; k += 1024) v += partial[k];
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    v = ws[threadIdx.x];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    if (threadIdx.x == 0 && !ctl->conv && ctl->iter < ctl->maxiter) {
      const double r = sqrt(v);
      const int k = ctl->iter;
      hist[k] = r;
      ctl->iter = k + 1;
      if ((r / ctl->r0 < ctl->rtol) && (r < ctl->atol)) {
        ctl->conv = 1;
        ctl->it_conv = k;
      }
    }
  }
}

// x -= y unless the loop has converged (LoopSolver update x -= Pr, loop_solver.cc:41)
__global__ void __launch_bounds__(256) axpy_guard_kernel(GridP g, double *__restrict__ x, const double *__restrict__ y, const SolverCtl *ctl) {
  if (ctl->conv) return;
  const int i = 1 + blockIdx.x * 64 + threadIdx.x;
  const int j = 1 + blockIdx.y * 4 + threadIdx.y;
  if (i >= g.nx || j >= g.ny) return;
  const long long o = (long long)j * g.pitch + i;
  x[o] -= y[o];
}

// x -= y on the interior; or x = 0
template <int MODE>  // 0: x -= y, 1: x = 0, 2: x = y
__global__ void __launch_bounds__(256) axpy_kernel(GridP g, double *__restrict__ x, const double *__restrict__ y) {
  const int i = 1 + blockIdx.x * 64 + threadIdx.x;
  const int j = 1 + blockIdx.y * 4 + threadIdx.y;
  if (i >= g.nx || j >= g.ny) return;
  const long long o = (long long)blockIdx.z * g.stride + (long long)j * g.pitch + i;
  if (MODE == 0) x[o] -= y[o];
  else if (MODE == 1) x[o] = 0.0;
  else x[o] = y[o];
}

// ------------------------------------------------------------------------------------------------
// Radius-2 operators (SquaredShiftedLaplaceFDOperator, squared_shiftedlaplace_fd_operator.cc:9-96, and
// its Galerkin coarsenings): 13-point stencil on the finest level, 21-point below, constant in the
// interior but with different coefficients on the first / last interior line of either direction
// (SURVEY.md section 7.3 H1).  `st` = 9 position classes x 25 coefficients, class = cx + 3 cy,
// coefficient (di, dj) at [(dj + 2) * 5 + (di + 2)].  Sweeps use 9 colours, colour = i % 3 + 3 (j % 3).
// First correct path: one launch per colour, in place (same-colour sites are >= 3 apart).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int pos_class_dev(int i, int n) { return i == 1 ? 0 : (i == n - 1 ? 2 : 1); }

__device__ __forceinline__ double stencil25(const double *__restrict__ a, const double *p, int pitch) {
  double s = 0.0;
#pragma unroll
  for (int dj = -2; dj <= 2; ++dj)
#pragma unroll
    for (int di = -2; di <= 2; ++di) {
      if ((di == -2 || di == 2) && (dj == -2 || dj == 2)) continue;  // the corners of the 5 x 5 box are never populated
      s = fma(a[(dj + 2) * 5 + (di + 2)], p[dj * pitch + di], s);
    }
  return s;
}

// y = A_0 x, or r = f - A_0 x
// (row strips: only the rows [rows.j0, rows.j1]; the grid is sized for that range)
struct RowRange {
  int j0, j1;
};
template <bool RESIDUAL>
__global__ void __launch_bounds__(256) apply25_kernel(GridP g, const double *__restrict__ st, const double *__restrict__ x, const double *__restrict__ f,
                                                     double *__restrict__ y, RowRange rows) {
  const int i = 1 + blockIdx.x * 64 + threadIdx.x;
  const int j = rows.j0 + blockIdx.y * 4 + threadIdx.y;
  if (i >= g.nx || j > rows.j1) return;
  const long long o = (long long)blockIdx.z * g.stride + (long long)j * g.pitch + i;
  const double s = stencil25(st + 25 * (pos_class_dev(i, g.nx) + 3 * pos_class_dev(j, g.ny)), x + o, g.pitch);
  y[o] = RESIDUAL ? (f[o] - s) : s;
}

// r = A_0 x - b with per-block partial sums of r^2 (LoopSolver)
__global__ void __launch_bounds__(256) residual_norm25_kernel(GridP g, const double *__restrict__ st, const double *__restrict__ x, const double *__restrict__ b,
                                                             double *__restrict__ r, double *__restrict__ partial) {
  const int i = 1 + blockIdx.x * 64 + threadIdx.x;
  const int j = 1 + blockIdx.y * 4 + threadIdx.y;
  double v = 0.0;
  if (i < g.nx && j < g.ny) {
    const long long o = (long long)j * g.pitch + i;
    v = stencil25(st + 25 * (pos_class_dev(i, g.nx) + 3 * pos_class_dev(j, g.ny)), x + o, g.pitch) - b[o];
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

// Row strips (one process per GPU): the colour launch itself exchanges the halo.  A rank sweeps its own rows
// [lo, hi]; the sites it updates in the `halo` rows next to a neighbour are stored a second time straight into the
// neighbour's copy of x (CUDA IPC mapping, NVLink) -- the per-colour halo exchange.  CTAs that touch those rows first
// wait (device side) until the neighbour has finished the previous launch -- its rows have arrived AND it no longer
// reads the rows about to be overwritten -- and the last of them raises the neighbour's flag.  Same counting protocol
// as the tile kernel (fused.cuh StripK): the flag value launch k of a cycle must see is cycle * per_cycle + k.
struct StripR2 {
  int on, lo, hi, halo;
  double *peer_dn, *peer_up;                 // the neighbours' x (same layout); nullptr at the ends of the lattice
  int *peer_flag_dn, *peer_flag_up;
  const int *flag_from_dn, *flag_from_up;
  const int *cycle_no;
  int per_cycle, index;
  unsigned int *ticket_dn, *ticket_up;
  unsigned int n_edge_dn, n_edge_up;         // CTAs of this launch that touch the rows next to the neighbour
  int *err;
};
__device__ __forceinline__ void r2_spin(const int *flag, int target, int *err) {
  const long long t0 = clock64();
  while (ld_acquire_sys(flag) < target) {
    if (clock64() - t0 > kWaitTimeoutClocks) {  // a peer died: raise the error word instead of hanging the GPU
      *err = 1;
      break;
    }
  }
}
__device__ __forceinline__ void r2_arrive(unsigned int *ticket, unsigned int n_expected, int *peer_flag) {
  if (atomicAdd(ticket, 1u) == n_expected - 1) {
    *ticket = 0u;
    __threadfence_system();
    atomicAdd_system(peer_flag, 1);
  }
}

// Transfers of a distributed radius-2 level run without exchange on own + mirrored rows, but they must not start before
// the neighbours have finished the launch that wrote those rows (sync), and -- where they modify mirrored rows
// (prolongation) -- the neighbours' next colour launch must not push into them before they are done (raise).
__global__ void strip_sync_kernel(const int *flag_dn, const int *flag_up, const int *cycle_no, int per_cycle, int index, int *err) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const int target = *cycle_no * per_cycle + index;
  if (flag_dn) r2_spin(flag_dn, target, err);
  if (flag_up) r2_spin(flag_up, target, err);
}
// one CTA: copy `n` doubles per chain into a neighbour's array (the first own row of the restricted right-hand side:
// the neighbour below forms its residual one row beyond its strip), then raise both neighbours' flags
__global__ void __launch_bounds__(256) strip_row_push_kernel(const double *__restrict__ src, double *__restrict__ dst, long long n, long long stride, int nchains,
                                                             int *peer_flag_dn, int *peer_flag_up) {
  if (dst)
    for (int ch = 0; ch < nchains; ++ch)
      for (long long k = threadIdx.x; k < n; k += blockDim.x) dst[ch * stride + k] = src[ch * stride + k];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    if (peer_flag_dn) atomicAdd_system(peer_flag_dn, 1);
    if (peer_flag_up) atomicAdd_system(peer_flag_up, 1);
  }
}
__global__ void strip_raise_kernel(int *peer_flag_dn, int *peer_flag_up) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  __threadfence_system();
  if (peer_flag_dn) atomicAdd_system(peer_flag_dn, 1);
  if (peer_flag_up) atomicAdd_system(peer_flag_up, 1);
}

// one colour (0..8) of a 9-colour SOR / Gibbs sweep; the noise of a site is the same pure function of the
// site as in the radius-1 kernels (philox.cuh).  jfirst = first row of this colour in the launch.
template <bool GIBBS>
__global__ void __launch_bounds__(256) sweep_colour25_kernel(GridP g, const double *__restrict__ st, double *__restrict__ x, const double *__restrict__ f, int colour,
                                                            double omega, NoiseP nz, int jfirst, StripR2 sk) {
  const int ci = colour % 3;
  const int i = ((ci == 0) ? 3 : ci) + 3 * (blockIdx.x * 64 + threadIdx.x);
  // row strips: the CTA rows at both ends of the strip run first (they feed the neighbours, whose next launch waits)
  const int by = sk.on ? ((blockIdx.y & 1) ? (int)(gridDim.y - 1 - (blockIdx.y >> 1)) : (int)(blockIdx.y >> 1)) : (int)blockIdx.y;
  const int j = jfirst + 3 * (by * 4 + threadIdx.y);
  const int jmax = sk.on ? min(sk.hi, g.ny - 1) : g.ny - 1;
  bool edge_dn = false, edge_up = false;
  if (sk.on) {
    const int jc0 = jfirst + 12 * by, jc1 = jc0 + 9;
    edge_dn = sk.peer_dn && jc0 < sk.lo + sk.halo;
    edge_up = sk.peer_up && jc1 > sk.hi - sk.halo && jc0 <= sk.hi;
    if (edge_dn || edge_up) {
      if (threadIdx.x == 0 && threadIdx.y == 0) {
        const int target = *sk.cycle_no * sk.per_cycle + sk.index;
        if (edge_dn) r2_spin(sk.flag_from_dn, target, sk.err);
        if (edge_up) r2_spin(sk.flag_from_up, target, sk.err);
      }
      __syncthreads();
    }
  }
  if (i < g.nx && j <= jmax) {
    const long long o = (long long)blockIdx.z * g.stride + (long long)j * g.pitch + i;
    const double *a = st + 25 * (pos_class_dev(i, g.nx) + 3 * pos_class_dev(j, g.ny));
    const double diag = a[12];
    double b = f[o];
    if (GIBBS) {
      double z0, z1;
      normal_pair(nz.keys, (((uint32_t)j * nz.G + (uint32_t)(i >> 2)) << 1) | (uint32_t)(i & 1), nz.c1, *nz.sample, nz.chain0 + blockIdx.z, nz.mc, kNormalTabDev, z0, z1);
      b = fma(sqrt(diag * (2. - omega) / omega), (i & 2) ? z1 : z0, b);  // sor_sampler.cc:24-27
    }
    const double v = x[o] + omega * (b - stencil25(a, x + o, g.pitch)) / diag;
    x[o] = v;
    if (sk.on) {
      if (sk.peer_dn && j < sk.lo + sk.halo) sk.peer_dn[o] = v;
      if (sk.peer_up && j > sk.hi - sk.halo) sk.peer_up[o] = v;
    }
  }
  if (edge_dn || edge_up) {
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0) {
      if (edge_dn) r2_arrive(sk.ticket_dn, sk.n_edge_dn, sk.peer_flag_dn);
      if (edge_up) r2_arrive(sk.ticket_up, sk.n_edge_up, sk.peer_flag_up);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Restriction f_c = R r with R = {1/2,1,1/2}^(x)2 un-normalised (IntergridOperator::restrict,
// intergrid_operator.hh:74-88, intergrid_operator_linear.cc:13): one thread per coarse vertex
// (I, J) <-> fine (2I, 2J).  (Inside a cycle the residual is never stored: the tile kernel restricts it
// from shared memory, fused.cuh.)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) restrict_kernel(GridP g, GridP gc, const double *__restrict__ r, double *__restrict__ fc, RowRange crows) {
  const int I = 1 + blockIdx.x * 64 + threadIdx.x;
  const int J = crows.j0 + blockIdx.y * 4 + threadIdx.y;
  if (I >= gc.nx || J > crows.j1) return;
  const long long of = (long long)blockIdx.z * g.stride + (long long)(2 * J) * g.pitch + 2 * I;
  double acc = 0.0;
#pragma unroll
  for (int dj = -1; dj <= 1; ++dj)
#pragma unroll
    for (int di = -1; di <= 1; ++di) acc += ((di == 0) ? 1.0 : 0.5) * ((dj == 0) ? 1.0 : 0.5) * r[of + (long long)dj * g.pitch + di];
  fc[(long long)blockIdx.z * gc.stride + (long long)J * gc.pitch + I] = acc;
}

// ------------------------------------------------------------------------------------------------
// x += alpha R^T x_c in gather form (IntergridOperator::prolongate_add, intergrid_operator.hh:106-120):
// every fine vertex reads its (up to) 4 coarse parents; boundary parents are the zero ghost lines.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) prolongate_add_kernel(GridP g, GridP gc, double alpha, const double *__restrict__ xc, double *__restrict__ x,
                                                            RowRange rows) {
  const int i = 1 + blockIdx.x * 64 + threadIdx.x;
  const int j = rows.j0 + blockIdx.y * 4 + threadIdx.y;
  if (i >= g.nx || j > rows.j1) return;
  const double *c = xc + (long long)blockIdx.z * gc.stride;
  const int I0 = i >> 1, J0 = j >> 1, I1 = (i + 1) >> 1, J1 = (j + 1) >> 1;
  // even index: I0 == I1 -> weight 1/2 + 1/2 = 1; odd index: the two neighbours with weight 1/2
  const double v = 0.25 * (c[(long long)J0 * gc.pitch + I0] + c[(long long)J0 * gc.pitch + I1] + c[(long long)J1 * gc.pitch + I0] +
                           c[(long long)J1 * gc.pitch + I1]);
  x[(long long)blockIdx.z * g.stride + (long long)j * g.pitch + i] += alpha * v;
}

// ------------------------------------------------------------------------------------------------
// Low-rank term.  Sparse column storage of a n x m matrix on the device: entries of column k are
// (site[e], val[e]) for e in [colptr[k], colptr[k+1]); "site" is the offset j * pitch + i.
// Row-grouped storage ("usites") for deterministic scatter: unique site u has entries
// (ucol[e], uval[e]) for e in [uptr[u], uptr[u+1]).
// ------------------------------------------------------------------------------------------------
struct SparseCols {
  int m;
  const int *colptr;
  const long long *site;
  const double *val;
};
struct SparseRows {
  int nu;
  const long long *usite;
  const int *uptr;
  const int *ucol;
  const double *uval;
};

// y += B Sigma^{-1} B^T x  (linear_operator.hh:71-75).  One block per chain; m <= 1024.
__global__ void __launch_bounds__(256) lowrank_apply_kernel(SparseCols B, SparseRows Br, const double *__restrict__ sigma_inv, long long stride,
                                                           const double *__restrict__ x, double *__restrict__ y) {
  extern __shared__ double sh[];
  double *t = sh;
  const double *xc = x + (long long)blockIdx.x * stride;
  double *yc = y + (long long)blockIdx.x * stride;
  for (int k = threadIdx.x; k < B.m; k += blockDim.x) {
    double s = 0.0;
    for (int e = B.colptr[k]; e < B.colptr[k + 1]; ++e) s += B.val[e] * xc[B.site[e]];
    t[k] = s * sigma_inv[k];
  }
  __syncthreads();
  for (int u = threadIdx.x; u < Br.nu; u += blockDim.x) {
    double s = 0.0;
    for (int e = Br.uptr[u]; e < Br.uptr[u + 1]; ++e) s += Br.uval[e] * t[Br.ucol[e]];
    yc[Br.usite[u]] += s;
  }
}

// low-rank part of the restricted residual: f_c -= R B Sigma^{-1} B^T x = B_c (Sigma^{-1} B^T x)
// (R B is exactly the coarse-level B, linear_operator.cc:19), completing multigridmc_sampler.cc:118-120
__global__ void __launch_bounds__(256) lowrank_restrict_kernel(SparseCols B, SparseRows Bc, const double *__restrict__ sigma_inv, long long stride,
                                                              long long stride_c, const double *__restrict__ x, double *__restrict__ fc) {
  extern __shared__ double sh[];
  double *t = sh;
  const double *xc = x + (long long)blockIdx.x * stride;
  double *fcc = fc + (long long)blockIdx.x * stride_c;
  for (int k = threadIdx.x; k < B.m; k += blockDim.x) {
    double s = 0.0;
    for (int e = B.colptr[k]; e < B.colptr[k + 1]; ++e) s += B.val[e] * xc[B.site[e]];
    t[k] = s * sigma_inv[k];
  }
  __syncthreads();
  for (int u = threadIdx.x; u < Bc.nu; u += blockDim.x) {
    double s = 0.0;
    for (int e = Bc.uptr[u]; e < Bc.uptr[u + 1]; ++e) s += Bc.uval[e] * t[Bc.ucol[e]];
    fcc[Bc.usite[u]] -= s;
  }
}

// Woodbury fix-up after a sweep on A_0 (SORSmoother::apply, sor_smoother.cc:47-51) merged with the
// low-rank part of the Gibbs noise (sor_sampler.cc:48-56):
//   y = x + W s,  s = Sigma^{-1/2} xi           (sweep is linear in its rhs: M_0^{-1} B s = W s)
//   x_new = y - W K B^T y = x + W d,   d = (I - K G) s - K B^T x,   W = M_0^{-1} B, G = B^T W,
//   K = (Sigma + G)^{-1}
// where M_0 = D/omega + L in the colour ordering, so W is sparse.
// Mneg = -K and Ms = I - K G are formed on the host.  One CTA per chain.  The kernel is pure latency
// (a handful of dependent loads), so B and W are stored padded with a fixed number of entries per
// column / per touched site: every thread issues all its index loads at once, the two m x m
// matrices are prefetched into shared memory meanwhile, and only x[site] depends on them.
struct LowRankFix {
  int m, EB, nu, EW;
  const long long *bsite;  // [m * EB]   sites of column k (padded with a repeated site, value 0)
  const double *bval;      // [m * EB]
  const long long *usite;  // [nu]       unique sites touched by W
  const int *wcol;         // [nu * EW]  (padded with column 0, value 0)
  const double *wval;      // [nu * EW]
  const double *Mneg, *Ms; // [m * m] row-major
  const double *sigma_inv_sqrt;
  int mats_in_smem;
};

template <bool GIBBS>
__global__ void __launch_bounds__(256) lowrank_fix_kernel(LowRankFix F, long long stride, double *__restrict__ x, NoiseP nz) {
  extern __shared__ double sh[];
  const int m = F.m;
  double *prod = sh;               // m * EB
  double *t = prod + m * F.EB;     // m
  double *s = t + m;               // m
  double *d = s + m;               // m
  double *Mn = d + m;              // m * m (if mats_in_smem)
  double *Mss = Mn + m * m;        // m * m
  double *xc = x + (long long)blockIdx.x * stride;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // independent loads first
  for (int e = threadIdx.x; e < m * F.EB; e += 256) prod[e] = F.bval[e] * xc[F.bsite[e]];
  if (F.mats_in_smem) {
    for (int e = threadIdx.x; e < m * m; e += 256) {
      Mn[e] = F.Mneg[e];
      if (GIBBS) Mss[e] = F.Ms[e];
    }
  }
  if (GIBBS) {
    const uint32_t sample = *nz.sample;
    for (int k = threadIdx.x; k < m; k += 256) {
      double z0, z1;
      normal_pair(nz.keys, 0x80000000u | ((uint32_t)k >> 1), nz.c1, sample, nz.chain0 + blockIdx.x, nz.mc, kNormalTabDev, z0, z1);
      s[k] = F.sigma_inv_sqrt[k] * ((k & 1) ? z1 : z0);
    }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < m; k += 256) {
    double acc = 0.0;
    for (int e = 0; e < F.EB; ++e) acc += prod[k * F.EB + e];
    t[k] = acc;
  }
  __syncthreads();
  const double *An = F.mats_in_smem ? Mn : F.Mneg, *As = F.mats_in_smem ? Mss : F.Ms;
  for (int k = warp; k < m; k += 8) {
    double acc = 0.0;
    for (int c = lane; c < m; c += 32) {
      acc = fma(An[k * m + c], t[c], acc);
      if (GIBBS) acc = fma(As[k * m + c], s[c], acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) d[k] = acc;
  }
  __syncthreads();
  for (int u = threadIdx.x; u < F.nu; u += 256) {
    double acc = 0.0;
    for (int e = 0; e < F.EW; ++e) acc += F.wval[u * F.EW + e] * d[F.wcol[u * F.EW + e]];
    xc[F.usite[u]] += acc;
  }
}

// ------------------------------------------------------------------------------------------------
// Wide supports: a GLOBAL measurement (MeasurementParameters::measure_global, measured_operator.cc:31-46) is a dense
// column of B, and then W = M_0^{-1} B has a dense column as well.  The padded one-CTA kernels above do not scale to
// that (they stage m x max-column-length products in shared memory); these three do the same algebra over the whole
// chip, deterministically (fixed reduction trees, no atomics):
//   1. lowrank_bt_partial_kernel   partial[chain][k][b] = sum of B_ek x_e over chunk b of column k
//   2. lowrank_d_kernel            t = sum_b partial;  MODE 0: d = scale * t  (apply / residual / restriction)
//                                  MODE 1 / 2: d = -K t (+ (I - K G) Sigma^{-1/2} xi)  (Woodbury fix-up, as lowrank_fix_kernel)
//   3. lowrank_scatter_kernel      y_u += sign * sum_k R_uk d_k  over the unique sites of R (B, the coarse B, or W)
// ------------------------------------------------------------------------------------------------
constexpr int kLrWideBlocks = 64;

__global__ void __launch_bounds__(256) lowrank_bt_partial_kernel(SparseCols B, long long stride, const double *__restrict__ x, double *__restrict__ partial) {
  const int k = blockIdx.y;
  const double *xc = x + (long long)blockIdx.z * stride;
  const int e0 = B.colptr[k], e1 = B.colptr[k + 1];
  const int chunk = (e1 - e0 + (int)gridDim.x - 1) / (int)gridDim.x;
  const int lo = e0 + (int)blockIdx.x * chunk, hi = min(lo + chunk, e1);
  double v = 0.0;
  for (int e = lo + (int)threadIdx.x; e < hi; e += 256) v = fma(B.val[e], xc[B.site[e]], v);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  __shared__ double ws[8];
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += ws[w];
    partial[((long long)blockIdx.z * B.m + k) * gridDim.x + blockIdx.x] = s;
  }
}

template <int MODE>
__global__ void __launch_bounds__(256) lowrank_d_kernel(int m, int nblk, const double *__restrict__ partial, const double *__restrict__ scale,
                                                       const double *__restrict__ Mneg, const double *__restrict__ Ms, const double *__restrict__ sigma_inv_sqrt, NoiseP nz,
                                                       double *__restrict__ d_out) {
  extern __shared__ double sh[];
  double *t = sh, *s = sh + m;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int k = threadIdx.x; k < m; k += 256) {
    double acc = 0.0;
    for (int b = 0; b < nblk; ++b) acc += partial[((long long)blockIdx.x * m + k) * nblk + b];
    t[k] = acc;
    if (MODE == 0) d_out[(long long)blockIdx.x * m + k] = scale[k] * acc;
    if (MODE == 2) {
      double z0, z1;
      normal_pair(nz.keys, 0x80000000u | ((uint32_t)k >> 1), nz.c1, *nz.sample, nz.chain0 + blockIdx.x, nz.mc, kNormalTabDev, z0, z1);
      s[k] = sigma_inv_sqrt[k] * ((k & 1) ? z1 : z0);
    }
  }
  if (MODE == 0) return;
  __syncthreads();
  for (int k = warp; k < m; k += 8) {
    double acc = 0.0;
    for (int c = lane; c < m; c += 32) {
      acc = fma(Mneg[k * m + c], t[c], acc);
      if (MODE == 2) acc = fma(Ms[k * m + c], s[c], acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) d_out[(long long)blockIdx.x * m + k] = acc;
  }
}

__global__ void __launch_bounds__(256) lowrank_scatter_kernel(SparseRows R, int m, const double *__restrict__ d, long long stride, double *__restrict__ y, double sign) {
  const int u = blockIdx.x * 256 + threadIdx.x;
  if (u >= R.nu) return;
  const double *dc = d + (long long)blockIdx.y * m;
  double acc = 0.0;
  for (int e = R.uptr[u]; e < R.uptr[u + 1]; ++e) acc = fma(R.uval[e], dc[R.ucol[e]], acc);
  y[(long long)blockIdx.y * stride + R.usite[u]] += sign * acc;
}

// ------------------------------------------------------------------------------------------------
// Row-strip domain decomposition over several GPUs (one process per GPU, peer memory mapped with CUDA
// IPC).  After a fused launch a rank stores the boundary rows of its output straight into the
// neighbours' arrays over NVLink and raises their flag; before a fused launch it waits (on the device,
// no host round trip, CUDA-graph capturable) until both neighbours have delivered the rows of the
// previous launch.  All ranks run the same launch sequence, so counting launches is the whole protocol.
// ------------------------------------------------------------------------------------------------
struct StripSeg {
  const double *src;
  double *dst;     // peer memory
  long long n;     // doubles, multiple of 2 (rows are 128-byte aligned and a multiple of 16 doubles long)
};
struct StripPush {
  StripSeg seg[8];
  int nseg;
  int *flag[8];    // peer flags to increment (system scope) once every segment has landed
  int nflag;
  unsigned int *ticket;  // local arrival counter of the CTAs of this launch
};

__global__ void __launch_bounds__(256) strip_push_kernel(const __grid_constant__ StripPush P) {
  for (int s = 0; s < P.nseg; ++s) {
    const double2 *src = reinterpret_cast<const double2 *>(P.seg[s].src);
    double2 *dst = reinterpret_cast<double2 *>(P.seg[s].dst);
    const long long n2 = P.seg[s].n >> 1;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n2; k += (long long)gridDim.x * blockDim.x) dst[k] = src[k];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(P.ticket, 1u);
    if (t == gridDim.x - 1) {  // last CTA: every other CTA has fenced its stores before taking its ticket
      *P.ticket = 0u;
      __threadfence_system();
      for (int f = 0; f < P.nflag; ++f) atomicAdd_system(P.flag[f], 1);
    }
  }
}


// wait until flag[f] >= per_wait[f] * (number of this wait, counted in *waitno) - lag[f]; a time-out
// (a peer died) raises *err instead of hanging the GPU
__global__ void strip_wait_kernel(const int *flag0, const int *flag1, int per_wait0, int per_wait1, int lag0, int lag1, int *waitno, int *err) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const int k = *waitno + 1;
  *waitno = k;
  const long long t0 = clock64();
  const int *flags[2] = {flag0, flag1};
  const int target[2] = {per_wait0 * k - lag0, per_wait1 * k - lag1};
  for (int f = 0; f < 2; ++f) {
    if (!flags[f]) continue;
    while (ld_acquire_sys(flags[f]) < target[f]) {
      if (clock64() - t0 > kWaitTimeoutClocks) {
        *err = 1;
        return;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// End of an MGMC cycle: z = sample_vector . x for every chain (driver_mgmc.cc:76), stored at
// series[pos * nchains + chain]; then advance the device-resident sample index and series position.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) end_of_cycle_kernel(int nnz, const long long *__restrict__ qsite, const double *__restrict__ qval,
                                                          const double *__restrict__ x, long long stride, int nchains, double *__restrict__ series,
                                                          long long series_cap, uint32_t *sample, unsigned long long *pos, int *cycle_no,
                                                          const double *__restrict__ qpart, int *lr_epoch) {
  // series_cap = capacity of `series` in SAMPLES (rows of nchains values): a call that does not ask for the series may
  // run more cycles than the buffer of an earlier call holds -- they are not recorded
  // qpart != nullptr (after a merged level-0 launch, fused.cuh): the sample only existed inside that launch, which
  // recorded the observed sites: x_site(e) of chain c at qpart[c * nnz + e]; same summation order as below
  const unsigned long long p = *pos;
  if (series != nullptr && nnz > 0 && p < (unsigned long long)series_cap) {
    for (int c = threadIdx.x >> 5; c < nchains; c += 8) {
      double v = 0.0;
      for (int e = threadIdx.x & 31; e < nnz; e += 32) v += qval[e] * (qpart ? qpart[(long long)c * nnz + e] : x[(long long)c * stride + qsite[e]]);
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
      if ((threadIdx.x & 31) == 0) series[p * nchains + c] = v;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    *sample = *sample + 1u;
    *pos = p + 1ull;
    if (cycle_no) *cycle_no = *cycle_no + 1;  // row strips: flag values are counted in cycles
    if (lr_epoch) *lr_epoch = *lr_epoch + 1;  // merged launches: the next unit starts a new epoch of the low-rank packets
  }
}

__global__ void bump_kernel(int *counter) { *counter = *counter + 1; }

// running mean / second moment fields (posterior_statistics, driver_mgmc.cc:146-151)
__global__ void __launch_bounds__(256) moments_kernel(GridP g, const double *__restrict__ x, double *__restrict__ mean, double *__restrict__ second,
                                                     double inv_count) {
  const int i = 1 + blockIdx.x * 64 + threadIdx.x;
  const int j = 1 + blockIdx.y * 4 + threadIdx.y;
  if (i >= g.nx || j >= g.ny) return;
  const long long o = (long long)j * g.pitch + i;
  const double v = x[o];
  mean[o] += (v - mean[o]) * inv_count;
  second[o] += (v * v - second[o]) * inv_count;
}

}  // namespace mgmc
