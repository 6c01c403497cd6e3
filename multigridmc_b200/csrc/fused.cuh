// Fused smoothing kernel: one CTA owns a tile of a level, stages the tile plus a halo of x and f in
// shared memory once, runs a whole sequence of colour passes (any mix of forward / backward sweeps)
// on it, and writes the tile back.  Optionally the prolongation of the coarse correction is fused in
// front (x += alpha R^T x_c while loading) and the residual + restriction behind
// (f_c = R (f - A_0 x), x_c = 0), so that a V(1,1) level visit of the prior operator is two kernels
// that move x, f once each way instead of 2 x 4 colour passes + 3 transfer kernels:
//   reference: SSORSampler::apply + LinearOperator::apply + IntergridOperator::restrict /
//   prolongate_add (ssor_sampler.cc:9-16, multigridmc_sampler.cc:116-127).
//
// Correctness of the overlapped tiles: stage k of S may only update sites whose neighbours held
// correct values after stage k-1, so the updated region shrinks by the stencil radius (1) per stage
// from tile +- (S - 1 [+ extra for the residual]) down to the tile; halo sites are recomputed by the
// neighbouring tiles with IDENTICAL results because the Gibbs noise is a pure function of
// (seed, chain, sample, level, sweep, site) -- see philox.cuh.  Because tiles overlap the sweep is out
// of place (x_in -> x_out, ping-pong).
//
// Geometry (all index arithmetic is compile-time): a region is always 128 columns = 32 aligned groups
// of 4 columns wide (one warp lane per group -- the unit that shares a Philox call) and RY rows tall.
// Shared memory holds row r as 4 planes of 32 doubles, plane k = columns 4p + k: for every neighbour
// access consecutive lanes read consecutive doubles of one plane (bank-conflict free), and every
// global access is a 128-bit load / store of an aligned column group.
// Tiles are group-aligned in x (i_t0 = TX * bx, TX = 128 - HXL - HXR) and start on odd rows in y
// (j_t0 = 1 + TY * by), which fixes the extra halo the fused residual needs: 2 columns left / 1 right,
// 1 row below / 2 above.
#pragma once
#include "kernels.cuh"

namespace mgmc {

constexpr int kGX = 16;  // allocated doubles left of i = 0
constexpr int kGY = 2;   // allocated rows below j = 0 / above j = ny
constexpr int kFusedThreads = 512;
constexpr int kFusedWarps = kFusedThreads / 32;

struct Stage {
  int colour;
  uint32_t c1;  // (level << 24) | sweep counter of the sweep this colour pass belongs to
};

struct FusedP {
  GridP g, gc;  // this level, next coarser level
  Coef9 a;
  const double *x_in;
  double *x_out;
  const double *f;
  const double *xc_in;  // PROLONG: coarse correction
  double alpha;
  double *fc_out;       // RESTRICT: coarse right-hand side
  double *xc_zero;      // RESTRICT: coarse iterate, set to zero (multigridmc_sampler.cc:122)
  int nstages;
  Stage st[8];
  double winv, noise_scale;  // omega / a_ii, sqrt(a_ii (2 - omega) / omega)
  NoiseP nz;
  int HXL, TX, TY, RY, hl;  // region geometry (host-computed, identical for all tiles)
  int omega_is_one;
  long long *timing;  // MGMC_TILE_TIMING builds only: 8 clock64 stamps + smid per CTA
};

// element at column offset K (-1..4) of group p in a shared-memory row
template <int K>
__device__ __forceinline__ double &sat(double *row, int p) {
  constexpr int plane = (K + 4) & 3;
  constexpr int dp = (K < 0) ? -1 : ((K >= 4) ? 1 : 0);
  return row[plane * 32 + p + dp];
}

// stencil sum  sum_j a_ij x_j  for the site at column offset K of group p
template <bool NINE, int K>
__device__ __forceinline__ double stencil_at(const Coef9 &a, double *row, int p) {
  double s = a.c * sat<K>(row, p) + a.w * sat<K - 1>(row, p) + a.e * sat<K + 1>(row, p) + a.s * sat<K>(row - 128, p) + a.n * sat<K>(row + 128, p);
  if (NINE)
    s += a.sw * sat<K - 1>(row - 128, p) + a.se * sat<K + 1>(row - 128, p) + a.nw * sat<K - 1>(row + 128, p) + a.ne * sat<K + 1>(row + 128, p);
  return s;
}

// off-diagonal part  sum_{j != i} a_ij x_j
template <bool NINE, int K>
__device__ __forceinline__ double offdiag_at(const Coef9 &a, double *row, int p) {
  double s = a.w * sat<K - 1>(row, p) + a.e * sat<K + 1>(row, p) + a.s * sat<K>(row - 128, p) + a.n * sat<K>(row + 128, p);
  if (NINE)
    s += a.sw * sat<K - 1>(row - 128, p) + a.se * sat<K + 1>(row - 128, p) + a.nw * sat<K - 1>(row + 128, p) + a.ne * sat<K + 1>(row + 128, p);
  return s;
}

// update the two sites (offset Q and Q + 2) of one colour in group p of one row:
//   x_i += omega (b_i - sum_j a_ij x_j) / a_ii  (sor_smoother.cc:75); for omega = 1 the old value drops
//   out, x_i = (b_i - sum_{j != i} a_ij x_j) / a_ii, which saves reading it from shared memory
template <bool NINE, bool GIBBS, bool W1, int Q>
__device__ __forceinline__ void update_pair(const Coef9 &a, double *xrow, const double *frow, int p, bool v0, bool v1, double winv, double nscale, double z0,
                                            double z1) {
  if (v0) {
    double b = frow[Q * 32 + p];
    if (GIBBS) b = fma(nscale, z0, b);
    if (W1) sat<Q>(xrow, p) = winv * (b - offdiag_at<NINE, Q>(a, xrow, p));
    else sat<Q>(xrow, p) += winv * (b - stencil_at<NINE, Q>(a, xrow, p));
  }
  if (v1) {
    double b = frow[(Q + 2) * 32 + p];
    if (GIBBS) b = fma(nscale, z1, b);
    if (W1) sat<Q + 2>(xrow, p) = winv * (b - offdiag_at<NINE, Q + 2>(a, xrow, p));
    else sat<Q + 2>(xrow, p) += winv * (b - stencil_at<NINE, Q + 2>(a, xrow, p));
  }
}

template <int NC, bool GIBBS, bool PROLONG, bool RESTRICT>
__global__ void __launch_bounds__(kFusedThreads, 2) fused_smooth_kernel(const __grid_constant__ FusedP P) {
  extern __shared__ double sm[];
  constexpr bool NINE = (NC == 4);
#ifdef MGMC_TILE_TIMING
  const int cta_id = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
#define TSTAMP(k) if (threadIdx.x == 0 && P.timing) P.timing[(long long)cta_id * 10 + (k)] = clock64();
  if (threadIdx.x == 0 && P.timing) { unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid)); P.timing[(long long)cta_id * 10 + 9] = smid; }
#else
#define TSTAMP(k)
#endif
  TSTAMP(0)
  const int RY = P.RY, TX = P.TX, TY = P.TY;
  double *xs = sm;
  double *fs = sm + RY * 128;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nx = P.g.nx, ny = P.g.ny, pitch = P.g.pitch;
  const int i_t0 = TX * blockIdx.x, j_t0 = 1 + TY * blockIdx.y;
  const int i_r0 = i_t0 - P.HXL, j_r0 = j_t0 - P.hl;
  const long long cbase = (long long)blockIdx.z * P.g.stride;
  const double *xg = P.x_in + cbase;
  double *xo = P.x_out + cbase;
  const double *fg = P.f + cbase;
  const int gi0 = i_r0 + 4 * lane;  // first global column of this lane's group
  const bool cols_alloc = (gi0 >= -kGX) && (gi0 + 3 < pitch - kGX);

  // ---- stage the region: one warp per row.  Lane l loads the column pairs (2l, 2l+1) and
  //      (64+2l, 64+2l+1): each 128-bit load instruction covers 512 contiguous bytes (fully coalesced);
  //      the pair is then scattered into planes 2(l&1), 2(l&1)+1 at index l/2 (+16), conflict free.
  //      Two rows per iteration keep 8 independent 128-bit loads per lane in flight. ----
  const int pa = 2 * (lane & 1), ia = lane >> 1;          // plane / index of column 2l; column 64+2l is at index ia + 16
  const int gia = i_r0 + 2 * lane, gib = gia + 64;        // global columns of the two pairs
  const bool oka = (gia >= -kGX) && (gia + 1 < pitch - kGX), okb = (gib >= -kGX) && (gib + 1 < pitch - kGX);
  for (int r = warp; r < RY; r += 2 * kFusedWarps) {
    double2 xa[2], xb[2], fa[2], fb[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int gj = j_r0 + r + u * kFusedWarps;
      xa[u] = xb[u] = fa[u] = fb[u] = make_double2(0.0, 0.0);
      if (gj >= -kGY && gj <= ny + kGY && r + u * kFusedWarps < RY) {
        const long long o = (long long)gj * pitch;
        if (oka) {
          xa[u] = *reinterpret_cast<const double2 *>(xg + o + gia);
          fa[u] = *reinterpret_cast<const double2 *>(fg + o + gia);
        }
        if (okb) {
          xb[u] = *reinterpret_cast<const double2 *>(xg + o + gib);
          fb[u] = *reinterpret_cast<const double2 *>(fg + o + gib);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int rr = r + u * kFusedWarps;
      if (rr >= RY) break;
      const int gj = j_r0 + rr;
      if (PROLONG) {
        if (gj >= 1 && gj < ny) {
          // x += alpha R^T x_c in gather form: every fine vertex reads its (up to) 4 coarse parents
          const double *xc = P.xc_in + (long long)blockIdx.z * P.gc.stride;
          const double *r0 = xc + (long long)(gj >> 1) * P.gc.pitch, *r1 = xc + (long long)((gj + 1) >> 1) * P.gc.pitch;
          const double al = P.alpha;
          if (oka && gia >= -1 && gia <= nx) {  // gia is even: coarse columns I, I + 1
            const int I = gia >> 1;
            const double q0 = 0.5 * (r0[I] + r1[I]), q1 = 0.5 * (r0[I + 1] + r1[I + 1]);
            if (gia >= 1 && gia < nx) xa[u].x += al * q0;
            if (gia + 1 >= 1 && gia + 1 < nx) xa[u].y += al * (0.5 * (q0 + q1));
          }
          if (okb && gib >= -1 && gib <= nx) {
            const int I = gib >> 1;
            const double q0 = 0.5 * (r0[I] + r1[I]), q1 = 0.5 * (r0[I + 1] + r1[I + 1]);
            if (gib >= 1 && gib < nx) xb[u].x += al * q0;
            if (gib + 1 >= 1 && gib + 1 < nx) xb[u].y += al * (0.5 * (q0 + q1));
          }
        }
      }
      double *xr = xs + rr * 128 + pa * 32 + ia, *fr = fs + rr * 128 + pa * 32 + ia;
      xr[0] = xa[u].x;
      xr[32] = xa[u].y;
      xr[16] = xb[u].x;
      xr[48] = xb[u].y;
      fr[0] = fa[u].x;
      fr[32] = fa[u].y;
      fr[16] = fb[u].x;
      fr[48] = fb[u].y;
    }
  }
  __syncthreads();
  TSTAMP(1)

  // ---- colour passes: one warp per row, lane = group ----
  const double winv = P.winv;
  const double nscale = P.noise_scale;
  const int S = P.nstages;
  constexpr int EXLX = RESTRICT ? 2 : 0, EXHX = RESTRICT ? 1 : 0, EXLY = RESTRICT ? 1 : 0, EXHY = RESTRICT ? 2 : 0;
  const uint32_t pg = (uint32_t)((i_r0 >> 2) + lane);
  const uint32_t sample = GIBBS ? *P.nz.sample : 0u;
  const uint32_t chain = P.nz.chain0 + blockIdx.z;
  for (int s = 0; s < S; ++s) {
    const int colour = P.st[s].colour;
    const uint32_t c1 = P.st[s].c1;
    const int m = S - 1 - s;
    const int ilo = max(1, i_t0 - m - EXLX), ihi = min(nx - 1, i_t0 + TX - 1 + m + EXHX);
    int jlo = max(1, j_t0 - m - EXLY);
    const int jhi = min(ny - 1, j_t0 + TY - 1 + m + EXHY);
    int step = 1;
    if (NC == 4) {
      step = 2;
      if ((jlo & 1) != (colour >> 1)) ++jlo;
    }
    for (int j = jlo + warp * step; j <= jhi; j += kFusedWarps * step) {
      const int q = (NC == 2) ? ((colour ^ j) & 1) : (colour & 1);
      const int i0 = gi0 + q;
      const bool v0 = (i0 >= ilo) && (i0 <= ihi), v1 = (i0 + 2 >= ilo) && (i0 + 2 <= ihi);
      if (!(v0 || v1)) continue;
      double z0 = 0.0, z1 = 0.0;
      if (GIBBS) normal_pair(P.nz.keys, (((uint32_t)j * P.nz.G + pg) << 1) | (uint32_t)q, c1, sample, chain, z0, z1);
      double *xr = xs + (j - j_r0) * 128;
      const double *fr = fs + (j - j_r0) * 128;
      if (P.omega_is_one) {
        if (q == 0) update_pair<NINE, GIBBS, true, 0>(P.a, xr, fr, lane, v0, v1, winv, nscale, z0, z1);
        else update_pair<NINE, GIBBS, true, 1>(P.a, xr, fr, lane, v0, v1, winv, nscale, z0, z1);
      } else {
        if (q == 0) update_pair<NINE, GIBBS, false, 0>(P.a, xr, fr, lane, v0, v1, winv, nscale, z0, z1);
        else update_pair<NINE, GIBBS, false, 1>(P.a, xr, fr, lane, v0, v1, winv, nscale, z0, z1);
      }
    }
    __syncthreads();
    TSTAMP(2 + (s < 4 ? s : 3))
  }

  // ---- write the tile to the output buffer (skipped by a pure residual + restrict launch):
  //      same coalesced column-pair mapping as the load ----
  if (PROLONG || S > 0) {
    const bool minea = (gia >= i_t0) && (gia < i_t0 + TX) && (gia <= nx) && oka;
    const bool mineb = (gib >= i_t0) && (gib < i_t0 + TX) && (gib <= nx) && okb;
    for (int rr = warp; rr < TY; rr += kFusedWarps) {
      const int gj = j_t0 + rr;
      if (gj >= ny) break;
      const double *xr = xs + (gj - j_r0) * 128 + pa * 32 + ia;
      const long long o = (long long)gj * pitch;
      // boundary / pad columns inside a pair hold the zeros they were loaded with
      if (minea) *reinterpret_cast<double2 *>(xo + o + gia) = make_double2(xr[0], xr[32]);
      if (mineb) *reinterpret_cast<double2 *>(xo + o + gib) = make_double2(xr[16], xr[48]);
    }
  }

  TSTAMP(6)
  // ---- residual on [i_t0 - 1, i_t0 + TX - 1] x [j_t0, j_t0 + TY], then full-weighting restriction ----
  if (RESTRICT) {
    for (int rr = warp; rr <= TY; rr += kFusedWarps) {
      const int gj = j_t0 + rr;
      double *__restrict__ xr = xs + (gj - j_r0) * 128;
      double *__restrict__ fr = fs + (gj - j_r0) * 128;
      double res[4] = {0.0, 0.0, 0.0, 0.0};
      if (gj < ny) {  // (lanes 0 / 31 read in-bounds garbage for columns that are masked out below)
        res[0] = fr[lane] - stencil_at<NINE, 0>(P.a, xr, lane);
        res[1] = fr[32 + lane] - stencil_at<NINE, 1>(P.a, xr, lane);
        res[2] = fr[64 + lane] - stencil_at<NINE, 2>(P.a, xr, lane);
        res[3] = fr[96 + lane] - stencil_at<NINE, 3>(P.a, xr, lane);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int gi = gi0 + k;
        const bool ok = (gi >= max(1, i_t0 - 1)) && (gi <= min(nx - 1, i_t0 + TX - 1)) && (gj < ny);
        fr[k * 32 + lane] = ok ? res[k] : 0.0;
      }
    }
    __syncthreads();
    const long long ccb = (long long)blockIdx.z * P.gc.stride;
    const int I = gi0 >> 1;  // coarse columns I (fine 4p) and I + 1 (fine 4p + 2) of this lane
    const bool mine = (gi0 >= i_t0) && (gi0 < i_t0 + TX);
    for (int rr = warp; rr < TY / 2; rr += kFusedWarps) {
      const int J = (j_t0 + 1) / 2 + rr;
      if (!mine || J >= P.gc.ny) continue;
      double *fr = fs + (2 * J - j_r0) * 128;
      const double a0 = sat<0>(fr, lane) + 0.5 * (sat<-1>(fr, lane) + sat<1>(fr, lane) + sat<0>(fr - 128, lane) + sat<0>(fr + 128, lane)) +
                        0.25 * (sat<-1>(fr - 128, lane) + sat<1>(fr - 128, lane) + sat<-1>(fr + 128, lane) + sat<1>(fr + 128, lane));
      const double a1 = sat<2>(fr, lane) + 0.5 * (sat<1>(fr, lane) + sat<3>(fr, lane) + sat<2>(fr - 128, lane) + sat<2>(fr + 128, lane)) +
                        0.25 * (sat<1>(fr - 128, lane) + sat<3>(fr - 128, lane) + sat<1>(fr + 128, lane) + sat<3>(fr + 128, lane));
      const long long o = ccb + (long long)J * P.gc.pitch + I;
      if (I >= 1 && I < P.gc.nx) {
        P.fc_out[o] = a0;
        if (P.xc_zero) P.xc_zero[o] = 0.0;
      }
      if (I + 1 < P.gc.nx) {
        P.fc_out[o + 1] = a1;
        if (P.xc_zero) P.xc_zero[o + 1] = 0.0;
      }
    }
  }
  TSTAMP(7)
}

}  // namespace mgmc
