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

// Device-resident data of the low-rank (measurement) term of one level for the in-kernel Woodbury
// fix-up (see "patch CTAs" below).  Sparse matrices are stored with explicit (i, j) coordinates.
struct LowRankTile {
  int m, EB;
  const int *b_i, *b_j;      // [m * EB]  entries of column k of B, padded by repeating an entry with value 0
  const double *b_val;       // [m * EB]
  const int *bbox;           // [m * 4]   i0, i1, j0, j1 of supp(B_k)
  int nbu;                   // B grouped by unique site (race-free scatter of B u into the residual)
  const int *bu_i, *bu_j, *bu_ptr, *bu_col;
  const double *bu_val;
  int nu[2], EW[2];          // W = M_0^{-1} B per sweep direction, grouped by unique site, EW padded entries each
  const int *w_i[2], *w_j[2], *w_col[2];
  const double *w_val[2];
  const double *Mneg[2], *Ms[2];  // m x m row-major: d = Ms s + Mneg (B^T x)
  int diag[2];               // Mneg, Ms are diagonal (B^T W is: the measurements do not interact on this level)
  const int *wbox[2];        // [m * 4]   i0, i1, j0, j1 of supp(W_k)
  // per measurement k: the W sites within 8 sites of supp(B_k), flattened (site coordinates, EW padded entries each)
  const int *wl_ptr[2], *wl_i[2], *wl_j[2], *wl_col[2];
  const double *wl_val[2];
  const double *sigma_inv, *sigma_inv_sqrt;
  double *dbuf;              // [nslots][nchains][m]      fix-up coefficients d (and u = Sigma^{-1} B^T x for the residual)
  double *tbuf;              // [nslots][nchains][2 m]    t = B^T x and s exchanged between patch CTAs
  int *flags;                // [nslots][nchains]         1 once dbuf[slot] is complete
  int *counters;             // [nslots][nchains]         arrival counter of the patch CTAs
};

// Row-strip decomposition (one process per GPU): the tile kernel itself exchanges the halo rows.  Tiles
// whose region reaches beyond the rank's own rows first wait (device-side, system-scope acquire) until
// the neighbour has delivered the rows of the previous launch; tiles that own rows of the neighbours'
// halo store them a second time, straight into the neighbour's array over NVLink (CUDA IPC mapping),
// and the last of those CTAs raises the neighbour's flag.  Every rank runs the same launch sequence, so
// the flag value a launch has to see is (cycle number) * (launches per cycle) + (its index in the cycle).
struct StripK {
  int on;
  int own_lo, own_hi, tiles_y;     // own rows of this level, own tile rows
  int halo;                        // rows of x mirrored into a neighbour
  int clo, chi, chalo;             // RESTRICT: own coarse rows, rows of f_c mirrored into a neighbour
  double *peer_x_dn, *peer_x_up;   // the neighbours' x_out (same layout); nullptr at the ends of the lattice
  double *peer_fc_dn, *peer_fc_up; // the neighbours' f of the coarser level (RESTRICT, coarser level distributed)
  int *peer_flag_dn, *peer_flag_up;          // flags to raise in the neighbours' memory
  const int *flag_from_dn, *flag_from_up;    // own flags, raised by the neighbours
  const int *cycle_no;
  int per_cycle, index;            // flag value to wait for = *cycle_no * per_cycle + index
  int edge_rows;                   // tile rows at either end of the strip that mirror rows
  unsigned int *ticket_dn, *ticket_up;
  int *err;
};

__device__ __forceinline__ int ld_acquire_sys_i(const int *p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// spin until *flag >= target; a time-out (a peer died) raises *err instead of hanging the GPU
__device__ __forceinline__ void strip_spin(const int *flag, int target, int *err) {
  const long long t0 = clock64();
  while (ld_acquire_sys_i(flag) < target) {
    if (clock64() - t0 > 6000000000ll) {
      *err = 1;
      break;
    }
  }
}

// arrival of one CTA that read halo rows and / or mirrored rows into a neighbour; the last one raises the flag
__device__ __forceinline__ void strip_arrive(unsigned int *ticket, unsigned int n_expected, int *peer_flag) {
  if (atomicAdd(ticket, 1u) == n_expected - 1) {
    *ticket = 0u;
    __threadfence_system();
    atomicAdd_system(peer_flag, 1);
  }
}

__device__ __forceinline__ long long gtimer() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

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
  int tiles_x;        // the grid is 1-d: npatch patch CTAs followed by tiles_x * tiles_y tile CTAs
  int by0;            // first tile row of this launch (row-strip decomposition: the rank's own tile rows)
  StripK sk;
  // low-rank term (LOWRANK kernels): fix-up q follows stage fix_stage[q]
  const LowRankTile *lr;
  int npatch, wpw, wcap;  // patch CTAs, windows per warp, doubles per window array
  int nfix;
  int fix_stage[2], fix_dir[2];
  uint32_t fix_c1[2];
  int lr_slot, nchains;   // first dbuf / flag slot of this launch (nfix fix-ups, then u)
};

__device__ __forceinline__ int ld_acquire(const int *p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int *p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

// ------------------------------------------------------------------------------------------------
// Patch CTAs: the Woodbury fix-up after a sweep (sor_smoother.cc:47-51, sor_sampler.cc:48-56),
//   x += W d,  d = (I - K G) Sigma^{-1/2} xi - K B^T x,
// needs B^T x of the freshly swept x: a grid-wide dependency in the middle of a fused launch.  But
// B^T x only involves the few sites of supp(B_k), and their values after `S` colour stages depend on
// the input x within distance S of them.  The first `npatch` CTAs of the grid therefore re-run the
// launch on small windows (supp(B_k) dilated by S sites) around every measurement -- same input, same
// Philox noise, same fix-ups -- publish d for every fix-up (and u = Sigma^{-1} B^T x of the final
// state for the low-rank part of the residual) in global memory and raise a flag; tile CTAs whose
// region contains a site of W (or B) wait for the flag and apply x += W d between two stages.
// Patch CTAs have the lowest block indices of the grid, so they are resident before any tile CTA can wait.
// ------------------------------------------------------------------------------------------------
template <int NC, bool GIBBS, bool PROLONG, bool RESTRICT>
__device__ __noinline__ void patch_cta(const FusedP &P, double *sm, int patch_id, int chainz) {
  const LowRankTile &R = *P.lr;
  const int m = R.m, S = P.nstages, wcap = P.wcap;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nx = P.g.nx, ny = P.g.ny, pitch = P.g.pitch;
  // shared memory: t, s, d (m each) | per window: x, f, noise of sweep 0, noise of sweep 1 (wcap each) | meta (wcap ints) | geometry
  double *tsm = sm, *ssm = sm + m, *dsm = sm + 2 * m;
  double *wins = sm + 3 * m;
  const int nwin = P.wpw;  // windows of this patch CTA
  int *meta = reinterpret_cast<int *>(wins + (size_t)nwin * 4 * wcap);
  int *geo = meta + (size_t)nwin * wcap;  // per window: wi0, wj0, wx, wy
  double *lrn = reinterpret_cast<double *>((reinterpret_cast<uintptr_t>(geo + 4 * nwin) + 7) & ~uintptr_t(7));  // [2][nwin] low-rank noise of the fix-ups
  const int k0 = patch_id * nwin;
  const int nloc = min(nwin, m - k0);
  const long long cbase = (long long)chainz * P.g.stride;
  const double *xg = P.x_in + cbase, *fg = P.f + cbase;
  const uint32_t sample = GIBBS ? *P.nz.sample : 0u;
  const uint32_t chain = P.nz.chain0 + chainz;
  const Coef9 &a = P.a;
  const double winv = P.winv, nscale = P.noise_scale;
#ifdef MGMC_TILE_TIMING
#define PSTAMP(k) if (threadIdx.x == 0 && P.timing) P.timing[(long long)(chainz * gridDim.x + patch_id) * 10 + (k)] = gtimer();
#else
#define PSTAMP(k)
#endif
  PSTAMP(0)

  for (int k = threadIdx.x; k < 3 * m; k += kFusedThreads) sm[k] = 0.0;  // padded W entries read d[0] with weight 0
  for (int w = threadIdx.x; w < nloc; w += kFusedThreads) {
    const int k = k0 + w;
    const int wi0 = max(0, R.bbox[4 * k] - S), wi1 = min(nx, R.bbox[4 * k + 1] + S);
    const int wj0 = max(0, R.bbox[4 * k + 2] - S), wj1 = min(ny, R.bbox[4 * k + 3] + S);
    geo[4 * w] = wi0;
    geo[4 * w + 1] = wj0;
    geo[4 * w + 2] = wi1 - wi0 + 1;
    geo[4 * w + 3] = wj1 - wj0 + 1;
  }
  // Everything a fix-up needs besides x is fetched now, while the windows load: the entries of B_k (one per lane),
  // the head of the W list of the window, the diagonal of the capacitance matrices (warp w <-> window w)
  const int kw = k0 + warp;               // measurement of this warp (windows per CTA <= warps per CTA)
  const bool has_win = warp < nloc;
  int pb_i = 0, pb_j = 0;
  double pb_val = 0.0;
  if (has_win && lane < R.EB) {
    pb_i = R.b_i[kw * R.EB + lane];
    pb_j = R.b_j[kw * R.EB + lane];
    pb_val = R.b_val[kw * R.EB + lane];
  }
  int wl0[2] = {0, 0}, wl1[2] = {0, 0};
  double dg_ms[2] = {0.0, 0.0}, dg_mn[2] = {0.0, 0.0}, sg_is = 0.0, sg_i = 0.0;
  if (has_win) {
#pragma unroll
    for (int dir = 0; dir < 2; ++dir) {
      wl0[dir] = R.wl_ptr[dir][kw];
      wl1[dir] = R.wl_ptr[dir][kw + 1];
      dg_ms[dir] = R.Ms[dir][(size_t)kw * m + kw];
      dg_mn[dir] = R.Mneg[dir][(size_t)kw * m + kw];
    }
    sg_is = R.sigma_inv_sqrt[kw];
    sg_i = R.sigma_inv[kw];
  }
  __syncthreads();
  // ---- load the windows; meta = (stage of the site's colour within a sweep) | (distance to supp(B_k)) << 8,
  //      -1 for sites that are never updated ----
  for (int t = threadIdx.x; t < nloc * wcap; t += kFusedThreads) {
    const int w = t / wcap, idx = t - w * wcap;
    const int wi0 = geo[4 * w], wj0 = geo[4 * w + 1], wx = geo[4 * w + 2], wy = geo[4 * w + 3];
    int mt = -1;
    if (idx < wx * wy) {
      const int k = k0 + w;
      const int i = wi0 + idx % wx, j = wj0 + idx / wx;
      double xv = xg[(long long)j * pitch + i];
      const double fv = fg[(long long)j * pitch + i];
      if (PROLONG && i >= 1 && i < nx && j >= 1 && j < ny) {
        const double *xc = P.xc_in + (long long)chainz * P.gc.stride;
        const double *r0 = xc + (long long)(j >> 1) * P.gc.pitch, *r1 = xc + (long long)((j + 1) >> 1) * P.gc.pitch;
        const int I0 = i >> 1, I1 = (i + 1) >> 1;
        xv += P.alpha * (0.25 * ((r0[I0] + r1[I0]) + (r0[I1] + r1[I1])));
      }
      double *xw = wins + (size_t)w * 4 * wcap;
      xw[idx] = xv;
      xw[wcap + idx] = fv;
      // window edges that are not the Dirichlet boundary are never updated
      const bool inner = (i > wi0) && (i < wi0 + wx - 1) && (j > wj0) && (j < wj0 + wy - 1);
      if (inner) {
        const int dist = max(max(R.bbox[4 * k] - i, i - R.bbox[4 * k + 1]), max(max(R.bbox[4 * k + 2] - j, j - R.bbox[4 * k + 3]), 0));
        const int colour = (NC == 2) ? ((i + j) & 1) : ((i & 1) + 2 * (j & 1));
        mt = colour | (dist << 8);
      }
    }
    meta[t] = mt;
  }
  __syncthreads();
  PSTAMP(1)
  if (P.sk.on && threadIdx.x == 0) {
    // row strips: this CTA has read everything it needs from the halo rows -- it counts as an edge CTA, so
    // that a neighbour cannot overwrite those rows (next launch) before the windows are loaded
    const unsigned int n_edge = (unsigned int)(P.sk.edge_rows * P.tiles_x + P.npatch);
    if (P.sk.peer_flag_dn) strip_arrive(P.sk.ticket_dn, n_edge, P.sk.peer_flag_dn);
    if (P.sk.peer_flag_up) strip_arrive(P.sk.ticket_up, n_edge, P.sk.peer_flag_up);
  }
  // ---- noise of every (sweep, site) that lies in the dependence cone of supp(B_k): the value of a B site
  //      after stage S - 1 depends on stage s only within distance S - 1 - s ----
  if (GIBBS) {
    const int nsweeps = S / NC;
    for (int t = threadIdx.x; t < nsweeps * nloc * wcap; t += kFusedThreads) {
      const int sw = t / (nloc * wcap), r = t - sw * (nloc * wcap);
      const int mt = meta[r];
      if (mt < 0) continue;
      const int colour = mt & 255, dist = mt >> 8;
      int s = sw * NC;
      while (P.st[s].colour != colour) ++s;
      if (dist > S - 1 - s) continue;
      const int w = r / wcap, idx = r - w * wcap;
      const int i = geo[4 * w] + idx % geo[4 * w + 2], j = geo[4 * w + 1] + idx / geo[4 * w + 2];
      double z0, z1;
      normal_pair(P.nz.keys, (((uint32_t)j * P.nz.G + (uint32_t)(i >> 2)) << 1) | (uint32_t)(i & 1), P.st[s].c1, sample, chain, z0, z1);
      wins[(size_t)w * 4 * wcap + (2 + sw) * wcap + idx] = (i & 2) ? z1 : z0;
    }
    // low-rank noise of every (fix-up, window): Sigma^{-1/2} xi (sor_sampler.cc:48-56), kept in ssm until the fix-up
    for (int t = threadIdx.x; t < P.nfix * nloc; t += kFusedThreads) {
      const int q = t / nloc, k = k0 + t % nloc;
      double z0, z1;
      normal_pair(P.nz.keys, 0x80000000u | ((uint32_t)k >> 1), P.fix_c1[q], sample, chain, z0, z1);
      lrn[q * nwin + t % nloc] = R.sigma_inv_sqrt[k] * ((k & 1) ? z1 : z0);
    }
    __syncthreads();
  }
  PSTAMP(2)

  int fixq = 0;
  for (int s = 0; s <= S; ++s) {
    if (s < S) {
      const int colour = P.st[s].colour;
      const int key_lo = colour, key_hi = colour | ((S - 1 - s) << 8);  // colour matches and dist <= S - 1 - s
      for (int t = threadIdx.x; t < nloc * wcap; t += kFusedThreads) {
        const int mt = meta[t];
        if (mt < 0 || (mt & 255) != key_lo || mt > key_hi) continue;
        const int w = t / wcap, idx = t - w * wcap;
        const int wx = geo[4 * w + 2];
        double *q = wins + (size_t)w * 4 * wcap + idx;
        double b = q[wcap];
        if (GIBBS) b = fma(nscale, q[(2 + s / NC) * wcap], b);
        double off = a.w * q[-1] + a.e * q[1] + a.s * q[-wx] + a.n * q[wx];
        if (NC == 4) off += a.sw * q[-wx - 1] + a.se * q[-wx + 1] + a.nw * q[wx - 1] + a.ne * q[wx + 1];
        if (P.omega_is_one) q[0] = winv * (b - off);
        else q[0] += winv * (b - (a.c * q[0] + off));
      }
      __syncthreads();
    }
    const bool fix_here = (s < S) && (fixq < P.nfix) && (P.fix_stage[fixq] == s);
    const bool u_here = RESTRICT && (s == S);
    if (!(fix_here || u_here)) continue;
    if (fixq == 0) { PSTAMP(5) }
    // ---- t = B^T x on every window (and the low-rank noise s): one warp per window ----
    const int slot = P.lr_slot + (fix_here ? fixq : P.nfix);
    double *tb = R.tbuf + ((size_t)slot * P.nchains + chainz) * 2 * m;
    double *db = R.dbuf + ((size_t)slot * P.nchains + chainz) * m;
    const int dir = fix_here ? P.fix_dir[fixq] : 0;
    const bool diag = fix_here && R.diag[dir];
    for (int w = warp; w < nloc; w += kFusedWarps) {
      const int k = k0 + w;
      const double *xw = wins + (size_t)w * 4 * wcap;
      const int wi0 = geo[4 * w], wj0 = geo[4 * w + 1], wx = geo[4 * w + 2];
      double acc = 0.0;
      if (lane < R.EB) acc = pb_val * xw[(pb_j - wj0) * wx + (pb_i - wi0)];
      for (int e = lane + 32; e < R.EB; e += 32) acc += R.b_val[k * R.EB + e] * xw[(R.b_j[k * R.EB + e] - wj0) * wx + (R.b_i[k * R.EB + e] - wi0)];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) {
        if (u_here) {
          db[k] = acc * sg_i;  // u_k: low-rank part of the residual, r -= B u
        } else {
          const double sv = GIBBS ? lrn[fixq * nwin + w] : 0.0;
          if (diag) {
            // capacitance matrix Sigma + B^T W is diagonal (measurements do not interact on this level)
            const double dk = fma(dir ? dg_ms[1] : dg_ms[0], sv, (dir ? dg_mn[1] : dg_mn[0]) * acc);
            dsm[k] = dk;
            db[k] = dk;
          } else {
            tsm[k] = acc;
            ssm[k] = sv;
            if (P.npatch > 1) {
              tb[k] = acc;
              tb[m + k] = sv;
            }
          }
        }
      }
    }
    int *counter = R.counters + (size_t)slot * P.nchains + chainz;
    int *flag = R.flags + (size_t)slot * P.nchains + chainz;
    if (fixq == 0) { PSTAMP(7) }
    if (u_here || diag) {
      // the last patch CTA to arrive publishes u / d
      __syncthreads();
      if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(counter, 1) == P.npatch - 1) st_release(flag, 1);
      }
      if (u_here) continue;
    } else {
      if (P.npatch > 1) {
        // all measurements are coupled through K: exchange t and s between the patch CTAs
        __syncthreads();
        if (threadIdx.x == 0) {
          __threadfence();
          atomicAdd(counter, 1);
          while (ld_acquire(counter) < P.npatch) {
          }
        }
        __syncthreads();
        for (int k = threadIdx.x; k < 2 * m; k += kFusedThreads) tsm[k] = __ldcg(tb + k);  // tsm and ssm are contiguous
      }
      __syncthreads();
      if (fixq == 0) { PSTAMP(8) }
      // ---- d = Ms s + Mneg t ----
      for (int k = warp; k < m; k += kFusedWarps) {
        double acc = 0.0;
        for (int c = lane; c < m; c += 32) {
          acc = fma(R.Mneg[dir][(size_t)k * m + c], tsm[c], acc);
          if (GIBBS) acc = fma(R.Ms[dir][(size_t)k * m + c], ssm[c], acc);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) {
          dsm[k] = acc;
          if (patch_id == 0) db[k] = acc;
        }
      }
      __syncthreads();
      if (patch_id == 0 && threadIdx.x == 0) {
        __threadfence();
        st_release(flag, 1);
      }
      if (fixq == 0) { PSTAMP(9) }
    }
    // ---- x += W d on every window: the W sites near window k are listed per window (with a diagonal
    //      capacitance matrix these belong to column k alone and only d_k is needed) ----
    for (int w = warp; w < nloc; w += kFusedWarps) {
      const int k = k0 + w;
      double *xw = wins + (size_t)w * 4 * wcap;
      const int wi0 = geo[4 * w], wj0 = geo[4 * w + 1], wx = geo[4 * w + 2], wy = geo[4 * w + 3];
      const int EW = R.EW[dir];
      for (int q = (dir ? wl0[1] : wl0[0]) + lane; q < (dir ? wl1[1] : wl1[0]); q += 32) {
        const int i = R.wl_i[dir][q] - wi0, j = R.wl_j[dir][q] - wj0;
        double acc = 0.0;
        for (int e = 0; e < EW; ++e) acc += R.wl_val[dir][(size_t)q * EW + e] * dsm[R.wl_col[dir][(size_t)q * EW + e]];
        if (i < 0 || i >= wx || j < 0 || j >= wy) continue;
        xw[j * wx + i] += acc;
      }
    }
    __syncthreads();  // tsm / ssm / dsm are reused by the next fix-up
    PSTAMP(3 + fixq)
    ++fixq;
  }
  PSTAMP(6)
}

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

template <int NC, bool GIBBS, bool PROLONG, bool RESTRICT, bool LOWRANK>
__global__ void __launch_bounds__(kFusedThreads, 2) fused_smooth_kernel(const __grid_constant__ FusedP P) {
  extern __shared__ double sm[];
  __shared__ int lr_hits;  // LOWRANK: bit 0 / 1: the region meets supp(W) of the forward / backward sweep, bit 2: the tile meets supp(B)
  constexpr bool NINE = (NC == 4);
  if (LOWRANK && (int)blockIdx.x < P.npatch) {
    if (P.sk.on) {  // the windows near the strip boundaries read halo rows
      if (threadIdx.x == 0) {
        const int target = *P.sk.cycle_no * P.sk.per_cycle + P.sk.index;
        if (P.sk.flag_from_dn) strip_spin(P.sk.flag_from_dn, target, P.sk.err);
        if (P.sk.flag_from_up) strip_spin(P.sk.flag_from_up, target, P.sk.err);
      }
      __syncthreads();
    }
    patch_cta<NC, GIBBS, PROLONG, RESTRICT>(P, sm, blockIdx.x, blockIdx.z);
    return;
  }
  const int tile_id = (int)blockIdx.x - (LOWRANK ? P.npatch : 0);
  int tile_row = tile_id / P.tiles_x;
  // row strips: the tile rows at both ends of the strip run first (they feed the neighbours)
  if (P.sk.on) tile_row = (tile_row & 1) ? (P.sk.tiles_y - 1 - (tile_row >> 1)) : (tile_row >> 1);
  const int tile_bx = tile_id % P.tiles_x, tile_by = tile_row + P.by0;
#ifdef MGMC_TILE_TIMING
  const int cta_id = blockIdx.z * gridDim.x + blockIdx.x;
#define TSTAMP(k) if (threadIdx.x == 0 && P.timing) P.timing[(long long)cta_id * 10 + (k)] = gtimer();
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
  const int i_t0 = TX * tile_bx, j_t0 = 1 + TY * tile_by;
  const int i_r0 = i_t0 - P.HXL, j_r0 = j_t0 - P.hl;
  const long long cbase = (long long)blockIdx.z * P.g.stride;
  const double *xg = P.x_in + cbase;
  double *xo = P.x_out + cbase;
  const double *fg = P.f + cbase;
  const int gi0 = i_r0 + 4 * lane;  // first global column of this lane's group
  const bool cols_alloc = (gi0 >= -kGX) && (gi0 + 3 < pitch - kGX);
  if (LOWRANK) {
    // does this tile meet a site touched by the low-rank fix-ups?  Tested up front (its loads overlap the tile
    // load, the barrier that ends the load publishes the result)
    if (threadIdx.x == 0) lr_hits = 0;
    __syncthreads();
    const LowRankTile &R = *P.lr;
    int bits = 0;
    for (int k = threadIdx.x; k < R.m; k += kFusedThreads) {
      const int4 b0 = reinterpret_cast<const int4 *>(R.wbox[0])[k], b1 = reinterpret_cast<const int4 *>(R.wbox[1])[k];
      bits |= (b0.y >= i_r0 && b0.x < i_r0 + 128 && b0.w >= j_r0 && b0.z < j_r0 + RY) ? 1 : 0;
      bits |= (b1.y >= i_r0 && b1.x < i_r0 + 128 && b1.w >= j_r0 && b1.z < j_r0 + RY) ? 2 : 0;
      if (RESTRICT) {
        const int4 bb = reinterpret_cast<const int4 *>(R.bbox)[k];  // i0, i1, j0, j1 of supp(B_k)
        bits |= (bb.y >= max(1, i_t0 - 1) && bb.x <= min(nx - 1, i_t0 + TX - 1) && bb.w >= j_t0 && bb.z <= min(j_t0 + TY, ny - 1)) ? 4 : 0;
      }
    }
    if (bits) atomicOr(&lr_hits, bits);
  }
  if (P.sk.on) {
    const bool wdn = P.sk.flag_from_dn && (j_r0 < P.sk.own_lo), wup = P.sk.flag_from_up && (j_r0 + RY - 1 > P.sk.own_hi);
    if (wdn || wup) {
      if (threadIdx.x == 0) {
        const int target = *P.sk.cycle_no * P.sk.per_cycle + P.sk.index;
        if (wdn) strip_spin(P.sk.flag_from_dn, target, P.sk.err);
        if (wup) strip_spin(P.sk.flag_from_up, target, P.sk.err);
      }
      __syncthreads();
    }
  }

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
  // does the region of this tile contain a site touched by the low-rank fix-ups?
  bool lr_need[2] = {false, false};
  int fixq = 0;
  if (LOWRANK) {
    lr_need[0] = (lr_hits & 1) != 0;
    lr_need[1] = (lr_hits & 2) != 0;
  }
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
    if (LOWRANK && fixq < P.nfix && P.fix_stage[fixq] == s) {
      // x += W d with d published by the patch CTAs
      const int dir = P.fix_dir[fixq];
      if (lr_need[dir]) {
        const LowRankTile &R = *P.lr;
        const size_t slot = (size_t)(P.lr_slot + fixq) * P.nchains + blockIdx.z;
        const double *db = R.dbuf + slot * R.m;
        const int EW = R.EW[dir], nu = R.nu[dir];
        // the W entries of this thread are fetched before the flag is polled: afterwards only d is missing
        int pidx[4], pcol[4];
        double pval[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int u = threadIdx.x + q * kFusedThreads;
          pidx[q] = -1;
          pcol[q] = 0;
          pval[q] = 0.0;
          if (u < nu) {
            const int i = R.w_i[dir][u], j = R.w_j[dir][u];
            pcol[q] = R.w_col[dir][(size_t)u * EW];
            pval[q] = R.w_val[dir][(size_t)u * EW];
            const int di = i - i_r0;
            if (i >= i_r0 && i < i_r0 + 128 && j >= j_r0 && j < j_r0 + RY) pidx[q] = (j - j_r0) * 128 + (di & 3) * 32 + (di >> 2);
          }
        }
        if (threadIdx.x == 0) {
          while (ld_acquire(R.flags + slot) == 0) {
          }
        }
        __syncthreads();
        TSTAMP(8)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (pidx[q] < 0) continue;
          const int u = threadIdx.x + q * kFusedThreads;
          double acc = pval[q] * __ldcg(db + pcol[q]);
          for (int e = 1; e < EW; ++e) acc += R.w_val[dir][(size_t)u * EW + e] * __ldcg(db + R.w_col[dir][(size_t)u * EW + e]);
          xs[pidx[q]] += acc;
        }
        for (int u = threadIdx.x + 4 * kFusedThreads; u < nu; u += kFusedThreads) {
          const int i = R.w_i[dir][u], j = R.w_j[dir][u];
          if (!(i >= i_r0 && i < i_r0 + 128 && j >= j_r0 && j < j_r0 + RY)) continue;
          double acc = 0.0;
          for (int e = 0; e < EW; ++e) acc += R.w_val[dir][(size_t)u * EW + e] * __ldcg(db + R.w_col[dir][(size_t)u * EW + e]);
          const int di = i - i_r0;
          xs[(j - j_r0) * 128 + (di & 3) * 32 + (di >> 2)] += acc;
        }
        __syncthreads();
      }
      ++fixq;
    }
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
      if (P.sk.on) {  // rows of the neighbours' halo: second store into their memory (pad columns of the pair included)
        double *pd = nullptr;
        if (P.sk.peer_x_dn && gj < P.sk.own_lo + P.sk.halo) pd = P.sk.peer_x_dn;
        if (pd) {
          if (minea) *reinterpret_cast<double2 *>(pd + cbase + o + gia) = make_double2(xr[0], xr[32]);
          if (mineb) *reinterpret_cast<double2 *>(pd + cbase + o + gib) = make_double2(xr[16], xr[48]);
        }
        if (P.sk.peer_x_up && gj > P.sk.own_hi - P.sk.halo) {
          pd = P.sk.peer_x_up;
          if (minea) *reinterpret_cast<double2 *>(pd + cbase + o + gia) = make_double2(xr[0], xr[32]);
          if (mineb) *reinterpret_cast<double2 *>(pd + cbase + o + gib) = make_double2(xr[16], xr[48]);
        }
      }
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
    if (LOWRANK) {
      // low-rank part of the residual: r -= B u, u = Sigma^{-1} B^T x of the final state (linear_operator.hh:71-75)
      const LowRankTile &R = *P.lr;
      if (lr_hits & 4) {  // (bounding boxes: a superset of the sites tested below)
        const size_t slot = (size_t)(P.lr_slot + P.nfix) * P.nchains + blockIdx.z;
        if (threadIdx.x == 0) {
          while (ld_acquire(R.flags + slot) == 0) {
          }
        }
        __syncthreads();
        const double *ub = R.dbuf + slot * R.m;
        for (int u = threadIdx.x; u < R.nbu; u += kFusedThreads) {
          const int i = R.bu_i[u], j = R.bu_j[u];
          if (!(i >= max(1, i_t0 - 1) && i <= min(nx - 1, i_t0 + TX - 1) && j >= j_t0 && j <= j_t0 + TY && j < ny)) continue;
          double acc = 0.0;
          for (int e = R.bu_ptr[u]; e < R.bu_ptr[u + 1]; ++e) acc += R.bu_val[e] * __ldcg(ub + R.bu_col[e]);
          const int di = i - i_r0;
          fs[(j - j_r0) * 128 + (di & 3) * 32 + (di >> 2)] -= acc;
        }
        __syncthreads();
      }
    }
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
      double *pf = nullptr;
      if (P.sk.on) {
        if (P.sk.peer_fc_dn && J < P.sk.clo + P.sk.chalo) pf = P.sk.peer_fc_dn;
        if (P.sk.peer_fc_up && J > P.sk.chi - P.sk.chalo) pf = P.sk.peer_fc_up;  // (strips are taller than two halos)
      }
      if (I >= 1 && I < P.gc.nx) {
        P.fc_out[o] = a0;
        if (P.xc_zero) P.xc_zero[o] = 0.0;
        if (pf) pf[o] = a0;
      }
      if (I + 1 < P.gc.nx) {
        P.fc_out[o + 1] = a1;
        if (P.xc_zero) P.xc_zero[o + 1] = 0.0;
        if (pf) pf[o + 1] = a1;
      }
    }
  }
  if (P.sk.on) {
    // the last CTA of the tile rows that mirror into a neighbour raises that neighbour's flag
    const int trow = tile_by - P.by0;
    const bool edge_dn = P.sk.peer_flag_dn && (trow < P.sk.edge_rows), edge_up = P.sk.peer_flag_up && (trow >= P.sk.tiles_y - P.sk.edge_rows);
    if (edge_dn || edge_up) {
      __threadfence_system();
      __syncthreads();
      if (threadIdx.x == 0) {
        const unsigned int n_edge = (unsigned int)(P.sk.edge_rows * P.tiles_x + (LOWRANK ? P.npatch : 0));
        if (edge_dn) strip_arrive(P.sk.ticket_dn, n_edge, P.sk.peer_flag_dn);
        if (edge_up) strip_arrive(P.sk.ticket_up, n_edge, P.sk.peer_flag_up);
      }
    }
  }
  TSTAMP(7)
}

}  // namespace mgmc
