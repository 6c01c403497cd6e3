// Fused smoothing kernel: one CTA owns a tile of a level, stages the tile plus a halo of x and f in
// shared memory once, runs a whole sequence of colour passes (any mix of forward / backward sweeps)
// on it, and writes the tile back.  Optionally the prolongation of the coarse correction is fused in
// front (x += alpha R^T x_c while loading) and the residual + restriction behind
// (f_c = R (f - A_0 x), x_c = 0), so that a V(1,1) level visit of the prior operator is two kernels
// that move x, f once each way instead of 2 x 4 colour passes + 3 transfer kernels:
//   reference: SSORSampler::apply + LinearOperator::apply + IntergridOperator::restrict /
//   prolongate_add (ssor_sampler.cc:9-16, multigridmc_sampler.cc:116-127).
//
// Correctness of the overlapped tiles: a colour pass may only update sites whose neighbours hold correct values, so
// the updated rectangle shrinks from pass to pass down to the tile (+ the halo of the fused residual); the host
// plans the rectangles backwards from what the launch must deliver (plan_stages in mgmc_b200.cu, see "Stage" below:
// directional margins, dead passes).  Halo sites are recomputed by the neighbouring tiles with IDENTICAL results
// because the Gibbs noise is a pure function of (seed, chain, sample, level, sweep, site) -- see philox.cuh.
// Because tiles overlap the sweep is out of place (x_in -> x_out, ping-pong).
//
// Geometry: a region is always 128 columns = 32 aligned groups
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
// 256 threads (8 warps) per CTA, 2 CTAs per SM, up to 128 registers per thread, two rows of a warp per pass iteration
// (MGMC_PASS_ILP 2): measured 1307 samples/s on C3 against 1267 with 512 threads / 64 registers / one row per iteration --
// a warp runs twice as many rows per colour pass, so the per-pass set-up (rectangle geometry, generator constants) is
// amortised over twice the work, and the two independent Philox / Box-Muller chains hide the fp64 latency that the
// second half of the warps used to hide (profiles/r02_summary.md)
#ifndef MGMC_FUSED_THREADS
#define MGMC_FUSED_THREADS 256
#endif
#ifndef MGMC_LOAD_ROWS
#define MGMC_LOAD_ROWS 2
#endif
#ifndef MGMC_PASS_PIPELINE
#define MGMC_PASS_PIPELINE 0
#endif
#ifndef MGMC_PASS_ILP
#define MGMC_PASS_ILP 2
#endif
#ifndef MGMC_FUSED_MINBLOCKS
#define MGMC_FUSED_MINBLOCKS 2  // CTAs per SM the tile kernel is compiled for (register cap); 1 with -DMGMC_FUSED_THREADS=1024
#endif
constexpr int kFusedThreads = MGMC_FUSED_THREADS;
constexpr int kFusedWarps = kFusedThreads / 32;

struct LrPkt;

// One colour pass of a fused launch.  The host plans the passes backwards from what the launch must deliver
// (plan_stages, mgmc_b200.cu): a pass only updates the rectangle tile + (xl, xh, yl, yh) whose values a later
// pass, the residual or the output can still see.  With omega = 1 a site update does not read the site's own
// value, so a pass whose colour is updated again before any other colour moves (the last colour of a forward
// sweep followed by a backward sweep) is dead: it is skipped, or -- if a low-rank fix-up reads x in between --
// restricted to supp(B_k) of the measurements the tile owns.  The results are bit-identical to running it.
enum { STAGE_FULL = 0, STAGE_SKIP = 1, STAGE_SPARSE = 2 };
struct Stage {
  int colour;
  uint32_t c1;  // (level << 24) | sweep counter of the sweep this colour pass belongs to
  short xl, xh, yl, yh;
  int mode;
  uint32_t soff;  // added to the sample index (1: the pass belongs to the pre-smoothing of the NEXT cycle, merged level-0 launch)
};
constexpr int kMaxStages = 16;  // colour passes per launch (8 on the big levels; the latency-bound small 4-colour levels take the 16
                                // passes of a V(2,2) smoothing step in one launch instead of two)
constexpr int kMaxFix = 4;   // low-rank fix-ups per launch (merged level-0 launch: 2 sweeps of cycle k + 2 of cycle k + 1)
constexpr int kMaxQoi = 8;   // observed sites a merged level-0 launch can record (more: the launches are not merged)

// Device-resident data of the low-rank (measurement) term of one level for the in-kernel Woodbury
// fix-up (see "Low-rank term inside the launch" below).  Sparse matrices carry explicit (i, j) coordinates.
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
  const int *wbox[2];        // [m * 4]   i0, i1, j0, j1 of supp(W_k)
  const double *Mneg[2], *Ms[2];  // m x m row-major: d = Ms s + Mneg (B^T x)
  int diag[2];               // Mneg, Ms are diagonal (B^T W is: the measurements do not interact on this level)
  const double *sigma_inv, *sigma_inv_sqrt;
  struct LrPkt *vbuf;        // [nslots][nchains][2 m]  published per fix-up: d_k (diagonal case) or t_k = (B^T x)_k and
                             //                         the noise s_k; for the residual slot: u_k = (Sigma^{-1} B^T x)_k
  const int *epoch;          // advanced once per cycle / API call: a packet is valid iff it carries the current epoch
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
  long long lr_peer_dn, lr_peer_up;  // byte offsets from this rank's vbuf / flags to the neighbours' copies (0: none)
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
    if (clock64() - t0 > kWaitTimeoutClocks) {
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
  Coef9 aw;   // a * winv (omega = 1 updates)
  double wn;  // noise_scale * winv
  const double *x_in;
  double *x_out;
  const double *f;
  const double *xc_in;  // PROLONG: coarse correction
  double alpha;
  double *fc_out;       // RESTRICT: coarse right-hand side
  double *xc_zero;      // RESTRICT: coarse iterate, set to zero (multigridmc_sampler.cc:122)
  int nstages;
  int res_stage;  // RESTRICT, omega = 1: index of the last pass (its sites get their residual for free), else -1
  Stage st[kMaxStages];
  double winv, noise_scale;  // omega / a_ii, sqrt(a_ii (2 - omega) / omega)
  NoiseP nz;
  int HXL, TX, TY, RY, hl;  // region geometry (host-computed, identical for all tiles)
  int omega_is_one;
  long long *timing;  // MGMC_TILE_TIMING builds only: 8 clock64 stamps + smid per CTA
  int tiles_x;        // the grid is 1-d: tiles_x * tiles_y tile CTAs
  int by0;            // first tile row of this launch (row-strip decomposition: the rank's own tile rows)
  StripK sk;
  // low-rank term (LOWRANK kernels): fix-up q follows stage fix_stage[q]
  LowRankTile lr;         // by value: its pointers sit in the constant bank, no dependent load to reach the tables
  int lr_mx, lr_my;       // extent of supp(B_k) beyond its lower left corner: extra halo on the high sides
  int nfix;
  int fix_stage[kMaxFix], fix_dir[kMaxFix];
  uint32_t fix_c1[kMaxFix], fix_soff[kMaxFix];
  // Merged level-0 launch (post-smoothing of cycle k + pre-smoothing of cycle k + 1): the sample x^(k) only exists
  // inside the launch, after stage qoi_stage (and its fix-up).  The tile that holds an observed site records its value
  // there: qoi_out[chain * nqoi + e] (end_of_cycle_kernel forms sample_vector . x from them, driver_mgmc.cc:76).
  int nqoi, qoi_stage;
  int qoi_i[kMaxQoi], qoi_j[kMaxQoi];
  double *qoi_out;
  // The iterate this launch starts from is zero by construction (multigridmc_sampler.cc:122, x_{l+1} = 0 before the
  // recursion) and is not read.  Set for the first launch of level 1 behind a merged level-0 launch, which cannot zero
  // the coarse iterate itself: its tiles read it (prolongation) while others would already be zeroing it.
  int x_in_zero;
  unsigned char *lr_flags; // [chains][tiles of the launch] see "Per-tile flag" in the kernel
  int lr_u_from_fix;      // RESTRICT: the launch ends with a fix-up, u = s - d (see the fix-up block)
  int lr_slot, nchains;   // first vbuf / flag slot of this launch (nfix fix-ups, then u)
  int rt_prolong, rt_restrict;  // persistent kernel of the small levels (tail.cuh): the flavour of the phase
  int chain_off;                // first chain of this launch (chains launched in groups)
  int *err;                     // error word of the context: a device-side wait that timed out sets it
  int nz_off, nz_cap;           // persistent kernel: noise generated ahead of the passes -- buffer offset in doubles behind
                                // the tile (0: off) and its capacity in (row, pass) items
  // NZG kernels (small levels): the normals of the full colour passes were generated ahead of the launch by
  // noise_gen_kernel (noise_ahead.cuh) while the big levels ran.  nzg[s] = plane of pass s, one pair per
  // (row, aligned group of 4 columns): [row index][nzg_gp] double2, group p at index p + kNzgPad
  const double2 *nzg[8];
  int nzg_gp;
};
constexpr int kNzgPad = 4;   // groups left of column 0 in a noise plane (= kGX / 4)
constexpr int kNzgRows = 3;  // rows of a warp per colour pass a NZG kernel can hold in registers (host: RY <= 16 * kNzgRows ...)

__device__ __forceinline__ int ld_acquire(const int *p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int *p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

// What an owner tile publishes: value and epoch travel in ONE aligned 16-byte store, so no fence / release is
// needed between "value written" and "flag raised" (a naturally aligned 128-bit access is a single transaction,
// also across NVLink); consumers poll the packet with volatile 128-bit loads (served by L2) until the epoch matches.
struct __align__(16) LrPkt {
  double v;
  int epoch, pad;
};
__device__ __forceinline__ void pkt_store(LrPkt *p, double v, int epoch) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"((unsigned)b), "r"((unsigned)(b >> 32)), "r"(epoch), "r"(0) : "memory");
}
// bounded wait: a bug (or a dead peer) must not hang the GPU -- a time-out raises the context's error word (the API
// call that finds it set fails with MGMC_ERR_CUDA instead of returning a chain built on a stale value).  The epoch
// comparison is wrap-safe.
__device__ __forceinline__ double pkt_wait(const LrPkt *p, int epoch, int *err) {
  unsigned lo, hi;
  int e, pad;
  const long long t0 = clock64();
  while (true) {
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(lo), "=r"(hi), "=r"(e), "=r"(pad) : "l"(p) : "memory");
    if ((int)(e - epoch) >= 0) break;
    if (clock64() - t0 > kWaitTimeoutClocks) {
      if (err) *err = 2;
      break;
    }
  }
  return __longlong_as_double((long long)(((unsigned long long)hi << 32) | lo));
}

// ------------------------------------------------------------------------------------------------
// Low-rank term inside the launch.  The Woodbury fix-up after a sweep (sor_smoother.cc:47-51,
// sor_sampler.cc:48-56),   x += W d,  d = (I - K G) Sigma^{-1/2} xi - K B^T x,
// needs t = B^T x of the freshly swept x: a grid-wide dependency between two colour passes of a fused
// launch.  But t_k only involves the few sites of supp(B_k).  Every measurement has an OWNER tile (the one
// that contains the lower left corner of supp(B_k); its halo is widened so that all of supp(B_k) is exact
// in its shared memory at every fix-up).  After the sweep the owner computes t_k from shared memory and
// publishes it -- d_k itself when the capacitance matrix is diagonal, i.e. the measurements do not
// interact on this level, else t_k and the noise s_k -- with a release store of the current epoch into
// flag k; every tile whose region meets supp(W) acquires the flags of the measurements it needs
// (all of them in the coupled case, where it forms the needed rows of d = Ms s + Mneg t itself) and
// applies W d in shared memory.  Tiles elsewhere never wait.  u = Sigma^{-1} B^T x of the final state
// (low-rank part of the residual) is published the same way.  Owners publish before they wait for
// anything, and only the few tiles next to a measurement wait, so the scheme cannot deadlock as long as
// waiting tiles do not fill the chip (host: coupled levels must fit on the chip at once).
// ------------------------------------------------------------------------------------------------
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
// RES (omega = 1, last pass before a fused residual): the residual of a site that has just been updated is known
// without a stencil evaluation, f_i - sum_j a_ij x_j = f_i - b_i = -(noise), so the pass leaves it in the slot of f_i
template <bool NINE, bool GIBBS, bool W1, int Q, bool RES>
__device__ __forceinline__ void update_pair(const Coef9 &a, const Coef9 &aw, double *xrow, double *frow, int p, bool v0, bool v1, double winv, double nscale,
                                            double wn, double z0, double z1) {
  // both sites are evaluated before either is stored (same-colour sites are never neighbours): the neighbour the two
  // stencils share is loaded once, the two dependent chains overlap, and the stores are the only predicated part
  // (masked lanes read in-bounds shared memory and drop the result)
  double b0 = frow[Q * 32 + p], b1 = frow[(Q + 2) * 32 + p];
  double r0, r1;
  if (W1) {
    // omega = 1: x_i = (f_i + n z_i - sum_{j != i} a_ij x_j) / a_ii with the division folded into the coefficients
    // (aw = a / a_ii, wn = n / a_ii: one instruction less per site)
    r0 = fma(winv, b0, -offdiag_at<NINE, Q>(aw, xrow, p));
    r1 = fma(winv, b1, -offdiag_at<NINE, Q + 2>(aw, xrow, p));
    if (GIBBS) {
      r0 = fma(wn, z0, r0);
      r1 = fma(wn, z1, r1);
    }
  } else {
    if (GIBBS) {
      b0 = fma(nscale, z0, b0);
      b1 = fma(nscale, z1, b1);
    }
    r0 = sat<Q>(xrow, p) + winv * (b0 - stencil_at<NINE, Q>(a, xrow, p));
    r1 = sat<Q + 2>(xrow, p) + winv * (b1 - stencil_at<NINE, Q + 2>(a, xrow, p));
  }
  if (v0) {
    sat<Q>(xrow, p) = r0;
    if (RES) frow[Q * 32 + p] = GIBBS ? -(nscale * z0) : 0.0;
  }
  if (v1) {
    sat<Q + 2>(xrow, p) = r1;
    if (RES) frow[(Q + 2) * 32 + p] = GIBBS ? -(nscale * z1) : 0.0;
  }
}

// one colour pass over the rows of a warp: xl / fl point at the lane's group (plane 0) in the first row
// PRE: the normals were generated ahead of the passes (fused_tile "noise ahead of the passes"); zp points at the pair of
// this lane for the first row, consecutive rows of the warp are dz pairs apart
template <bool NINE, bool GIBBS, bool W1, int Q, bool RES = false, bool PRE = false>
__device__ __forceinline__ void pass_rows(const FusedP &P, double *xl, double *fl, int nrows, int dl, uint32_t c0, uint32_t dc0, uint32_t c1,
                                          uint32_t sample, uint32_t chain, const double *ntab, bool v0, bool v1, const double2 *zp = nullptr, int dz = 0) {
  const double winv = P.winv, nscale = P.noise_scale, wn = P.wn;
  int n = 0;
  if (PRE) {
    for (; n < nrows; ++n) {
      const double2 z = *zp;
      update_pair<NINE, GIBBS, W1, Q, RES>(P.a, P.aw, xl, fl, 0, v0, v1, winv, nscale, wn, z.x, z.y);
      xl += dl;
      fl += dl;
      zp += dz;
    }
    return;
  }
#if MGMC_PASS_ILP == 2
  // two rows of the warp per iteration: two independent Philox / Box-Muller / update chains in flight
  for (; n + 1 < nrows; n += 2) {
    double z0 = 0.0, z1 = 0.0, y0 = 0.0, y1 = 0.0;
    if (GIBBS) {
      normal_pair(P.nz.keys, c0, c1, sample, chain, P.nz.mc, ntab, z0, z1);
      normal_pair(P.nz.keys, c0 + dc0, c1, sample, chain, P.nz.mc, ntab, y0, y1);
    }
    update_pair<NINE, GIBBS, W1, Q, RES>(P.a, P.aw, xl, fl, 0, v0, v1, winv, nscale, wn, z0, z1);
    update_pair<NINE, GIBBS, W1, Q, RES>(P.a, P.aw, xl + dl, fl + dl, 0, v0, v1, winv, nscale, wn, y0, y1);
    xl += 2 * dl;
    fl += 2 * dl;
    c0 += 2 * dc0;
  }
#endif
#if MGMC_PASS_PIPELINE
  // software pipeline: the normals of the next row are generated before the current row is updated, so that the
  // shared-memory latency and the dependent chain of the update overlap the (independent) Philox rounds
  double z0 = 0.0, z1 = 0.0;
  if (GIBBS && n < nrows) normal_pair(P.nz.keys, c0, c1, sample, chain, P.nz.mc, ntab, z0, z1);
  for (; n < nrows; ++n) {
    double y0 = 0.0, y1 = 0.0;
    c0 += dc0;
    if (GIBBS && n + 1 < nrows) normal_pair(P.nz.keys, c0, c1, sample, chain, P.nz.mc, ntab, y0, y1);
    update_pair<NINE, GIBBS, W1, Q, RES>(P.a, P.aw, xl, fl, 0, v0, v1, winv, nscale, wn, z0, z1);
    xl += dl;
    fl += dl;
    z0 = y0;
    z1 = y1;
  }
#else
  for (; n < nrows; ++n) {
    double z0 = 0.0, z1 = 0.0;
    if (GIBBS) normal_pair(P.nz.keys, c0, c1, sample, chain, P.nz.mc, ntab, z0, z1);
    update_pair<NINE, GIBBS, W1, Q, RES>(P.a, P.aw, xl, fl, 0, v0, v1, winv, nscale, wn, z0, z1);
    xl += dl;
    fl += dl;
    c0 += dc0;
  }
#endif
}

// NZG kernels: one colour pass over the (at most kNzgRows) rows of a warp with the normals already in registers
template <bool NINE, bool W1, int Q, bool RES>
__device__ __forceinline__ void pass_rows_z(const FusedP &P, double *xl, double *fl, int nrows, int dl, bool v0, bool v1, const double2 (&z)[kNzgRows]) {
  const double winv = P.winv, nscale = P.noise_scale, wn = P.wn;
#pragma unroll
  for (int n = 0; n < kNzgRows; ++n) {
    if (n < nrows) update_pair<NINE, true, W1, Q, RES>(P.a, P.aw, xl + n * dl, fl + n * dl, 0, v0, v1, winv, nscale, wn, z[n].x, z[n].y);
  }
}

// One tile job: tile `tile_id` of chain `chain` of the launch / phase described by P.  `njob_x` = tiles per chain
// (index of the per-tile flags).  sm = dynamic shared memory of the CTA; lr_cnt[4], ntab[128] = static shared memory
// (LOWRANK: measurements owned by this tile, needed for the forward / backward fix-up, for the residual; GIBBS: tables
// of the normal generator (philox.cuh), published by the barrier of the tile load).
// PR / RS = 0, 1: prolongation in front / residual + restriction behind compiled out / in; 2: decided at run time by
// P.rt_prolong / P.rt_restrict (the persistent kernel of the small levels runs both flavours, tail.cuh).
template <int NC, bool GIBBS, int PR, int RS, bool LOWRANK, bool NZG = false>
__device__ __forceinline__ void fused_tile(const FusedP &P, const int tile_id, const int chain_idx, const int njob_x, double *sm, int *lr_cnt, double *ntab) {
  const bool PROLONG = (PR == 2) ? (P.rt_prolong != 0) : (PR == 1);
  const bool RESTRICT = (RS == 2) ? (P.rt_restrict != 0) : (RS == 1);
  constexpr bool NINE = (NC == 4);
  // Persistent kernel of the small levels: a warp has at most a row or two per colour pass there, and a pass is the
  // in-order latency of ONE Philox + Box-Muller + update chain (~1400 cycles, 80 % of it the normals) times the number
  // of passes.  The normals do not depend on x: all of them -- every (row, pass) item of the launch -- are generated
  // ahead of the passes by all warps, two independent chains per warp in flight, and parked in shared memory.
  constexpr bool TAILMODE = (PR == 2);
  int tile_row = tile_id / P.tiles_x;
  // row strips: the tile rows at both ends of the strip run first (they feed the neighbours)
  if (P.sk.on) tile_row = (tile_row & 1) ? (P.sk.tiles_y - 1 - (tile_row >> 1)) : (tile_row >> 1);
  const int tile_bx = tile_id % P.tiles_x, tile_by = tile_row + P.by0;
#ifdef MGMC_TILE_TIMING
  const int cta_id = chain_idx * njob_x + tile_id;
#define TSTAMP(k) if (threadIdx.x == 0 && P.timing) P.timing[(long long)cta_id * 16 + (k)] = gtimer();
  if (threadIdx.x == 0 && P.timing) { unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid)); P.timing[(long long)cta_id * 16 + 15] = smid; }
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
  const long long cbase = (long long)chain_idx * P.g.stride;
  const double *xg = P.x_in + cbase;
  double *xo = P.x_out + cbase;
  const double *fg = P.f + cbase;
  const int gi0 = i_r0 + 4 * lane;  // first global column of this lane's group
  const bool cols_alloc = (gi0 >= -kGX) && (gi0 + 3 < pitch - kGX);
  // shared memory behind the tile: d (or u) of the current fix-up, t and s (coupled case), the low-rank noise and the
  // diagonal capacitance entries of the owned measurements, then the index lists
  // (the pointers are re-derived inside every low-rank block instead of being kept live across the colour passes:
  //  the passes run at the 64-register limit)
#define MGMC_LR_PTRS                                                                                                  \
  const int lrm = P.lr.m;                                                                                             \
  double *darr = sm + 2 * P.RY * 128, *tarr = darr + lrm, *sarr = tarr + lrm;                                         \
  double *spre = sarr + lrm;    /* [nfix][m] noise of fix-up q for owned k (sized by the fix-ups of THIS launch: */   \
  double *cms = spre + P.nfix * lrm; /* [nfix][m] Ms_kk per fix-up (owned k)     1.5 KB more per CTA push the two-sweep */ \
  double *cmn = cms + P.nfix * lrm;  /* [nfix][m] Mneg_kk                        launches of level 0 over a carve-out step) */ \
  double *uarr = cmn + P.nfix * lrm; /* u_k = s_k - d_k of the last fix-up (low-rank part of the residual) */          \
  int *own_list = reinterpret_cast<int *>(uarr + lrm), *need_list = own_list + lrm; /* need_list: [3][m] */           \
  int *is_own = need_list + 3 * lrm;                                                                                  \
  const int lr_epoch = *P.lr.epoch;                                                                                   \
  const bool lr_to_dn = P.sk.on && P.sk.lr_peer_dn != 0 && (tile_by - P.by0 < P.sk.edge_rows);                        \
  const bool lr_to_up = P.sk.on && P.sk.lr_peer_up != 0 && (tile_by - P.by0 >= P.sk.tiles_y - P.sk.edge_rows);        \
  (void)darr; (void)tarr; (void)sarr; (void)spre; (void)cms; (void)cmn; (void)uarr; (void)own_list; (void)need_list; (void)is_own; \
  (void)lr_epoch; (void)lr_to_dn; (void)lr_to_up;
  // Per-tile flag of this launch geometry (self-initialising: 0xFF = not known yet): 0 = the tile is nowhere near a
  // measurement and skips every low-rank block -- set-up, tests and their loads included -- on one uniform branch
  int lr_flag = 0;
  // (one flag per tile AND chain: a flag shared by the chains could change under the threads of a CTA that is still
  //  reading it -- the CTA of another chain writes it -- and split the CTA at the barriers below)
  const size_t lr_flag_idx = (size_t)chain_idx * njob_x + tile_id;
  if (LOWRANK) lr_flag = (P.lr.m > 0) ? (P.lr_flags ? (int)P.lr_flags[lr_flag_idx] : 0xFF) : 0;  // (m = 0: a phase without low-rank work inside a LOWRANK kernel)
  if (LOWRANK && lr_flag) {
    MGMC_LR_PTRS
    if (threadIdx.x < 4) lr_cnt[threadIdx.x] = 0;
    for (int k = threadIdx.x; k < lrm; k += kFusedThreads) {
      darr[k] = 0.0;
      is_own[k] = 0;
    }
    __syncthreads();
  }
  // ---- persistent kernel of the small levels: pass descriptors and noise ahead of the passes ----
  // A warp has a row or two per colour pass there, and a pass costs the in-order latency of its set-up (~60 dependent
  // integer instructions) plus one Philox + Box-Muller + update chain (~1400 cycles, 80 % of it the normals).  Both are
  // taken off the pass loop: one thread per (pass, warp) works out what the warp does in the pass (descriptor), and
  // all normals of the launch -- one item per (row, live pass) -- are generated by all warps, three independent
  // chains per warp in flight, and parked in shared memory.
  double2 *nzb = nullptr;
  int *aux = nullptr;  // [0..7] first item of a pass, [8..15] first row, [16..23] rows, [32 + 4 (16 s + w)] descriptor of
                       // (pass s, warp w): row offset, rows | q << 16, ilo, ihi; [32 + 512 + 2 t] (row, pass) of item t
  bool pre_on = false;
  if (TAILMODE) {
    nzb = reinterpret_cast<double2 *>(sm + P.nz_off);
    aux = reinterpret_cast<int *>(nzb + (size_t)P.nz_cap * 32);
    const int S0 = P.nstages;
    if ((int)threadIdx.x < S0 * kFusedWarps) {
      const int st = threadIdx.x / kFusedWarps, w = threadIdx.x % kFusedWarps;
      int n = 0, jlo = 0, rowoff = 0, nrows = 0, q = 0, ilo = 0, ihi = -1;
      if (P.st[st].mode == STAGE_FULL) {
        const int colour = P.st[st].colour;
        ilo = max(1, i_t0 - P.st[st].xl);
        ihi = min(nx - 1, i_t0 + TX - 1 + P.st[st].xh);
        jlo = max(1, j_t0 - P.st[st].yl);
        const int jhi = min(ny - 1, j_t0 + TY - 1 + P.st[st].yh);
        const int step = (NC == 4) ? 2 : 1;
        if (NC == 4 && (jlo & 1) != (colour >> 1)) ++jlo;
        n = (jhi >= jlo) ? (jhi - jlo) / step + 1 : 0;
        const int jw = jlo + w * step;
        if (jw <= jhi) {
          nrows = (jhi - jw) / (kFusedWarps * step) + 1;
          rowoff = (jw - j_r0) * 128;
          q = (NC == 2) ? ((colour ^ jw) & 1) : (colour & 1);
        }
      }
      reinterpret_cast<int4 *>(aux + 32)[threadIdx.x] = make_int4(rowoff, nrows | (q << 16), ilo, ihi);
      if (w == 0) {
        aux[8 + st] = jlo;
        aux[16 + st] = n;
      }
    }
    __syncthreads();
    if (GIBBS && P.nz_cap > 0) {
      int tot = 0;
      {
        // item t = (pass, k-th row of the pass): every thread walks the (at most 8) passes
        const int t = threadIdx.x;
        for (int st = 0; st < S0; ++st) {
          const int n = aux[16 + st];
          if (t == 0) aux[st] = tot;
          if (t >= tot && t < tot + n) {
            aux[32 + 512 + 2 * t] = aux[8 + st] + (t - tot) * ((NC == 4) ? 2 : 1);
            aux[33 + 512 + 2 * t] = st;
          }
          tot += n;
        }
      }
      pre_on = tot <= P.nz_cap;  // (uniform)
      __syncthreads();
      if (pre_on) {
        const uint32_t pg0 = (uint32_t)((i_r0 >> 2) + lane);
        const uint32_t sample0 = *P.nz.sample, chain0 = P.nz.chain0 + chain_idx;
        constexpr int NI = 3;
        for (int it0 = warp; it0 < tot; it0 += NI * kFusedWarps) {
          double z0[NI], z1[NI];
#pragma unroll
          for (int u = 0; u < NI; ++u) {
            const int it = min(it0 + u * kFusedWarps, tot - 1);  // (a surplus chain recomputes the last item)
            const int jr = aux[32 + 512 + 2 * it], sr = aux[33 + 512 + 2 * it];
            const int cr = P.st[sr].colour;
            const uint32_t qr = (NC == 2) ? ((cr ^ jr) & 1) : (cr & 1);
            normal_pair(P.nz.keys, (((uint32_t)jr * P.nz.G + pg0) << 1) | qr, P.st[sr].c1, sample0 + P.st[sr].soff, chain0, P.nz.mc, ntab, z0[u], z1[u]);
          }
#pragma unroll
          for (int u = 0; u < NI; ++u) {
            const int it = it0 + u * kFusedWarps;
            if (it < tot) nzb[it * 32 + lane] = make_double2(z0[u], z1[u]);
          }
        }
      }
    }
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

  // ---- NZG kernels: what this warp does in colour pass s, and the normals of that pass (generated ahead of the launch,
  //      noise_ahead.cuh) fetched into registers one pass ahead -- the first pass before the tile load ----
  auto pass_geom = [&](int s_, int &jw_, int &nrows_, int &q_, bool &v0_, bool &v1_) {
    const int colour_ = P.st[s_].colour;
    const int ilo_ = max(1, i_t0 - P.st[s_].xl), ihi_ = min(nx - 1, i_t0 + TX - 1 + P.st[s_].xh);
    int jlo_ = max(1, j_t0 - P.st[s_].yl);
    const int jhi_ = min(ny - 1, j_t0 + TY - 1 + P.st[s_].yh);
    const int step_ = (NC == 4) ? 2 : 1;
    if (NC == 4 && (jlo_ & 1) != (colour_ >> 1)) ++jlo_;
    jw_ = jlo_ + warp * step_;
    nrows_ = (jw_ <= jhi_) ? (jhi_ - jw_) / (kFusedWarps * step_) + 1 : 0;
    q_ = (NC == 2) ? ((colour_ ^ jw_) & 1) : (colour_ & 1);
    const int i0_ = gi0 + q_;
    v0_ = (i0_ >= ilo_) && (i0_ <= ihi_);
    v1_ = (i0_ + 2 >= ilo_) && (i0_ + 2 <= ihi_);
  };
  // (two register sets used alternately: a copy "current = next" would wait for the loads in flight)
  double2 za[kNzgRows], zb[kNzgRows];
  bool z_flip = false;  // false: the current pass reads za and loads the next one into zb
  auto nzg_load = [&](int s_, double2 (&zd)[kNzgRows]) {
    int jw_, nrows_, q_;
    bool v0_, v1_;
    pass_geom(s_, jw_, nrows_, q_, v0_, v1_);
    const double2 *zp = P.nzg[s_] + ((long long)((NC == 4) ? (jw_ >> 1) : jw_) * P.nzg_gp + ((i_r0 >> 2) + lane + kNzgPad));
#pragma unroll
    for (int n = 0; n < kNzgRows; ++n)
      if (n < nrows_ && (v0_ || v1_)) zd[n] = __ldg(zp + (long long)n * kFusedWarps * P.nzg_gp);
  };
  auto next_full = [&](int s_) {
    while (s_ < P.nstages && P.st[s_].mode != STAGE_FULL) ++s_;
    return s_;
  };
  if (NZG) {
#pragma unroll
    for (int n = 0; n < kNzgRows; ++n) za[n] = zb[n] = make_double2(0.0, 0.0);
    const int s0_ = next_full(0);
    if (s0_ < P.nstages) nzg_load(s0_, za);
  }

  // ---- stage the region: one warp per row.  Lane l loads the column pairs (2l, 2l+1) and
  //      (64+2l, 64+2l+1): each 128-bit load instruction covers 512 contiguous bytes (fully coalesced);
  //      the pair is then scattered into planes 2(l&1), 2(l&1)+1 at index l/2 (+16), conflict free.
  //      Two rows per iteration keep 8 independent 128-bit loads per lane in flight. ----
  const int pa = 2 * (lane & 1), ia = lane >> 1;          // plane / index of column 2l; column 64+2l is at index ia + 16
  const int gia = i_r0 + 2 * lane, gib = gia + 64;        // global columns of the two pairs
  const bool oka = (gia >= -kGX) && (gia + 1 < pitch - kGX), okb = (gib >= -kGX) && (gib + 1 < pitch - kGX);
  constexpr int LU = MGMC_LOAD_ROWS;  // rows of a warp whose loads are in flight together (4 x 128-bit loads per lane and row)
  for (int r = warp; r < RY; r += LU * kFusedWarps) {
    double2 xa[LU], xb[LU], fa[LU], fb[LU];
#pragma unroll
    for (int u = 0; u < LU; ++u) {
      const int gj = j_r0 + r + u * kFusedWarps;
      xa[u] = xb[u] = fa[u] = fb[u] = make_double2(0.0, 0.0);
      if (gj >= -kGY && gj <= ny + kGY && r + u * kFusedWarps < RY) {
        const long long o = (long long)gj * pitch;
        if (oka) {
          if (!P.x_in_zero) xa[u] = *reinterpret_cast<const double2 *>(xg + o + gia);
          fa[u] = *reinterpret_cast<const double2 *>(fg + o + gia);
        }
        if (okb) {
          if (!P.x_in_zero) xb[u] = *reinterpret_cast<const double2 *>(xg + o + gib);
          fb[u] = *reinterpret_cast<const double2 *>(fg + o + gib);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < LU; ++u) {
      const int rr = r + u * kFusedWarps;
      if (rr >= RY) break;
      const int gj = j_r0 + rr;
      if (PROLONG) {
        if (gj >= 1 && gj < ny) {
          // x += alpha R^T x_c in gather form: every fine vertex reads its (up to) 4 coarse parents
          const double *xc = P.xc_in + (long long)chain_idx * P.gc.stride;
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
  if (LOWRANK && lr_flag && (int)threadIdx.x < P.lr.m) {
    // Which measurements does this tile own / need?  (after the tile load has been issued; published by its barrier)
    MGMC_LR_PTRS
    const LowRankTile &R = P.lr;
    const int k = threadIdx.x;
    const int4 bb = reinterpret_cast<const int4 *>(R.bbox)[k];  // i0, i1, j0, j1 of supp(B_k)
    const int4 b0 = reinterpret_cast<const int4 *>(R.wbox[0])[k], b1 = reinterpret_cast<const int4 *>(R.wbox[1])[k];
    const bool hit_res = RESTRICT && bb.y >= max(1, i_t0 - 1) && bb.x <= min(nx - 1, i_t0 + TX - 1) && bb.w >= j_t0 && bb.z <= min(j_t0 + TY, ny - 1);
    // (a measurement that enters the residual of this tile is also needed at the fix-ups: u_k = s_k - d_k below)
    if (hit_res || (b0.y >= i_r0 && b0.x < i_r0 + 128 && b0.w >= j_r0 && b0.z < j_r0 + RY)) need_list[atomicAdd(&lr_cnt[1], 1)] = k;
    if (hit_res || (b1.y >= i_r0 && b1.x < i_r0 + 128 && b1.w >= j_r0 && b1.z < j_r0 + RY)) need_list[lrm + atomicAdd(&lr_cnt[2], 1)] = k;
    if (hit_res) need_list[2 * lrm + atomicAdd(&lr_cnt[3], 1)] = k;
    const bool owner = bb.x >= i_t0 && bb.x < i_t0 + TX && bb.z >= j_t0 && bb.z < j_t0 + TY;  // the tile that holds the lower left corner
    if (owner) {
      own_list[atomicAdd(&lr_cnt[0], 1)] = k;
      is_own[k] = 1;
    }
#pragma unroll 1  // (not unrolled: every copy carries a normal_pair, and the code size of the kernel costs instruction-cache misses in the pass loops)
    for (int q = 0; q < P.nfix; ++q) {
      const int dir = P.fix_dir[q];
      if (owner) {
        cms[q * lrm + k] = R.Ms[dir][(size_t)k * lrm + k];
        cmn[q * lrm + k] = R.Mneg[dir][(size_t)k * lrm + k];
      }
      // the low-rank noise s_k = Sigma_k^{-1/2} xi_k is a pure function of the counters: every tile forms it itself
      double sv = 0.0;
      if (GIBBS) {
        double z0, z1;
        normal_pair(P.nz.keys, 0x80000000u | ((uint32_t)k >> 1), P.fix_c1[q], *P.nz.sample + P.fix_soff[q], P.nz.chain0 + chain_idx, P.nz.mc, ntab, z0, z1);
        sv = R.sigma_inv_sqrt[k] * ((k & 1) ? z1 : z0);
      }
      spre[q * lrm + k] = sv;
    }
  }
  __syncthreads();
  TSTAMP(1)
  // most tiles are nowhere near a measurement: they skip every low-rank block below (and its loads) on one uniform test
  const bool lr_own = LOWRANK && lr_flag && lr_cnt[0] > 0;
  const bool lr_tile = LOWRANK && lr_flag && (lr_cnt[0] | lr_cnt[1] | lr_cnt[2]) != 0;
  const bool lr_res = LOWRANK && lr_flag && lr_cnt[3] > 0;
  if (LOWRANK && lr_flag == 0xFF && P.lr_flags && threadIdx.x == 0) P.lr_flags[lr_flag_idx] = (lr_cnt[0] | lr_cnt[1] | lr_cnt[2] | lr_cnt[3]) ? 1 : 0;

  // ---- colour passes: one warp per row, lane = group ----
  const int S = P.nstages;
  const uint32_t pg = (uint32_t)((i_r0 >> 2) + lane);
  const uint32_t sample = GIBBS ? *P.nz.sample : 0u;
  const uint32_t chain = P.nz.chain0 + chain_idx;
  // the colour passes run in segments that end at a low-rank fix-up (or at the last pass): the fix-up code stays out
  // of the body of the pass loop
#ifdef MGMC_TILE_TIMING
  long long tacc_setup = 0, tacc_pass = 0, tacc_bar = 0, tq0 = 0, tq1 = 0, tq2 = 0;
#define TCLK(v) v = clock64();
#else
#define TCLK(v)
#endif
  // observed sites of a merged level-0 launch that this tile holds (uniform over the CTA)
  bool tile_qoi = false;
  for (int e = 0; e < P.nqoi; ++e)
    tile_qoi = tile_qoi || (P.qoi_i[e] >= i_t0 && P.qoi_i[e] < i_t0 + TX && P.qoi_j[e] >= j_t0 && P.qoi_j[e] < j_t0 + TY);
  // segments of passes: each ends at a low-rank fix-up; without fix-ups, at the stage after which the observed sites
  // are recorded
  const int nsegfix = LOWRANK ? P.nfix : 0;
  const int nbreak = nsegfix > 0 ? nsegfix : (P.nqoi > 0 ? 1 : 0);
  int s = 0;
  for (int seg = 0; seg <= nbreak; ++seg) {
   const int s_end = (seg < nbreak) ? ((nsegfix > 0) ? P.fix_stage[seg] : P.qoi_stage) + 1 : S;
   for (; s < s_end; ++s) {
    TCLK(tq0)
    const int colour = P.st[s].colour;
    const uint32_t c1 = P.st[s].c1;
    const uint32_t smp = sample + P.st[s].soff;
    const int mode = P.st[s].mode;
    if (mode == STAGE_SKIP) continue;  // (uniform over the launch)
    if (mode == STAGE_SPARSE) {
      // dead pass that a fix-up (or the recording of the observed sites) still looks at: only supp(B_k) of the owned
      // measurements and the observed sites of this tile (omega = 1 here)
      if (!(lr_own || tile_qoi)) continue;
      auto sparse_box = [&](int bi0, int bi1, int bj0, int bj1) {
        for (int j = bj0; j <= bj1; ++j) {
          if (NC == 4 && (j & 1) != (colour >> 1)) continue;
          const int q = (NC == 2) ? ((colour ^ j) & 1) : (colour & 1);
          const int i0 = gi0 + q;
          const bool v0 = (i0 >= bi0) && (i0 <= bi1), v1 = (i0 + 2 >= bi0) && (i0 + 2 <= bi1);
          if (!(v0 || v1)) continue;
          double *xl = xs + (j - j_r0) * 128 + lane;
          double *fl = fs + (j - j_r0) * 128 + lane;
          const uint32_t c0 = (((uint32_t)j * P.nz.G + pg) << 1) | (uint32_t)q;
          if (q == 0) pass_rows<NINE, GIBBS, true, 0>(P, xl, fl, 1, 0, c0, 0u, c1, smp, chain, ntab, v0, v1);
          else pass_rows<NINE, GIBBS, true, 1>(P, xl, fl, 1, 0, c0, 0u, c1, smp, chain, ntab, v0, v1);
        }
      };
      // (an observed site inside supp(B_k) of an owned measurement is updated by that box: the boxes of one tile must
      //  not update a site twice in a pass -- a second update would draw the same normal and give the same value, but
      //  two warps would write it concurrently; same value, benign, and excluded here for the observed sites)
      // (one loop and one call site for both kinds of boxes: every copy of the pass code carries a normal_pair)
      int n_own = 0;
      const int *own = nullptr;
      if (LOWRANK && lr_own) {
        MGMC_LR_PTRS
        n_own = lr_cnt[0];
        own = own_list;
      }
      for (int o = warp; o < n_own + (tile_qoi ? P.nqoi : 0); o += kFusedWarps) {
        int bi0, bi1, bj0, bj1;
        if (o < n_own) {
          const int4 bb = reinterpret_cast<const int4 *>(P.lr.bbox)[own[o]];  // i0, i1, j0, j1
          bi0 = bb.x, bi1 = bb.y, bj0 = bb.z, bj1 = bb.w;
        } else {
          bi0 = bi1 = P.qoi_i[o - n_own];
          bj0 = bj1 = P.qoi_j[o - n_own];
          if (!(bi0 >= i_t0 && bi0 < i_t0 + TX && bj0 >= j_t0 && bj0 < j_t0 + TY)) continue;
        }
        sparse_box(bi0, bi1, bj0, bj1);
      }
      __syncthreads();
      continue;
    }
    if (TAILMODE) {
      // persistent kernel of the small levels: what this warp does in the pass was worked out ahead (descriptor)
      const int4 d = reinterpret_cast<const int4 *>(aux + 32)[s * kFusedWarps + warp];
      const int nrows = d.y & 0xffff, q = d.y >> 16;
      const int i0 = gi0 + q;
      const bool v0 = (i0 >= d.z) && (i0 <= d.w), v1 = (i0 + 2 >= d.z) && (i0 + 2 <= d.w);
      TCLK(tq1)
      if (nrows > 0 && (v0 || v1)) {
        double *xl = xs + d.x + lane, *fl = fs + d.x + lane;
        const int dl = kFusedWarps * ((NC == 4) ? 2 : 1) * 128;
        const bool res = RESTRICT && s == P.res_stage && !lr_tile;
        if (GIBBS && pre_on) {
          // normals generated ahead of the passes: item of this warp's first row, 16 items per round of rows
          const double2 *zp = nzb + (size_t)(aux[s] + warp) * 32 + lane;
          const int dz = kFusedWarps * 32;
#define MGMC_TAIL_PASS(W1_, RES_)                                                                                           \
  {                                                                                                                         \
    if (q == 0) pass_rows<NINE, GIBBS, W1_, 0, RES_, true>(P, xl, fl, nrows, dl, 0u, 0u, c1, smp, chain, ntab, v0, v1, zp, dz); \
    else pass_rows<NINE, GIBBS, W1_, 1, RES_, true>(P, xl, fl, nrows, dl, 0u, 0u, c1, smp, chain, ntab, v0, v1, zp, dz);       \
  }
          if (res) MGMC_TAIL_PASS(true, true)
          else if (P.omega_is_one) MGMC_TAIL_PASS(true, false)
          else MGMC_TAIL_PASS(false, false)
#undef MGMC_TAIL_PASS
        } else {
          const int jw = d.x / 128 + j_r0;
          const uint32_t c0 = (((uint32_t)jw * P.nz.G + pg) << 1) | (uint32_t)q, dc0 = ((uint32_t)(kFusedWarps * ((NC == 4) ? 2 : 1)) * P.nz.G) << 1;
#define MGMC_TAIL_PASS(W1_, RES_)                                                                                     \
  {                                                                                                                   \
    if (q == 0) pass_rows<NINE, GIBBS, W1_, 0, RES_>(P, xl, fl, nrows, dl, c0, dc0, c1, smp, chain, ntab, v0, v1); \
    else pass_rows<NINE, GIBBS, W1_, 1, RES_>(P, xl, fl, nrows, dl, c0, dc0, c1, smp, chain, ntab, v0, v1);       \
  }
          if (res) MGMC_TAIL_PASS(true, true)
          else if (P.omega_is_one) MGMC_TAIL_PASS(true, false)
          else MGMC_TAIL_PASS(false, false)
#undef MGMC_TAIL_PASS
        }
      }
      TCLK(tq2)
      __syncthreads();
#ifdef MGMC_TILE_TIMING
      tacc_setup += tq1 - tq0;
      tacc_pass += tq2 - tq1;
      tacc_bar += clock64() - tq2;
#endif
      continue;
    }
    if (NZG) {
      int jw, nrows, q;
      bool v0, v1;
      pass_geom(s, jw, nrows, q, v0, v1);
      const int s2 = next_full(s + 1);
      auto run_pass = [&](const double2 (&zc)[kNzgRows], double2 (&zd)[kNzgRows]) {
        if (s2 < S) nzg_load(s2, zd);  // in flight while this pass runs
        TCLK(tq1)
        if (nrows > 0 && (v0 || v1)) {
          double *xl = xs + (jw - j_r0) * 128 + lane;
          double *fl = fs + (jw - j_r0) * 128 + lane;
          const int dl = kFusedWarps * ((NC == 4) ? 2 : 1) * 128;
          if (RESTRICT && s == P.res_stage && !lr_tile) {
            if (q == 0) pass_rows_z<NINE, true, 0, true>(P, xl, fl, nrows, dl, v0, v1, zc);
            else pass_rows_z<NINE, true, 1, true>(P, xl, fl, nrows, dl, v0, v1, zc);
          } else if (P.omega_is_one) {
            if (q == 0) pass_rows_z<NINE, true, 0, false>(P, xl, fl, nrows, dl, v0, v1, zc);
            else pass_rows_z<NINE, true, 1, false>(P, xl, fl, nrows, dl, v0, v1, zc);
          } else {
            if (q == 0) pass_rows_z<NINE, false, 0, false>(P, xl, fl, nrows, dl, v0, v1, zc);
            else pass_rows_z<NINE, false, 1, false>(P, xl, fl, nrows, dl, v0, v1, zc);
          }
        }
      };
      if (z_flip) run_pass(zb, za);
      else run_pass(za, zb);
      z_flip = !z_flip;
      TCLK(tq2)
      __syncthreads();
#ifdef MGMC_TILE_TIMING
      tacc_setup += tq1 - tq0;
      tacc_pass += tq2 - tq1;
      tacc_bar += clock64() - tq2;
#endif
      continue;
    }
    const int ilo = max(1, i_t0 - P.st[s].xl), ihi = min(nx - 1, i_t0 + TX - 1 + P.st[s].xh);
    int jlo = max(1, j_t0 - P.st[s].yl);
    const int jhi = min(ny - 1, j_t0 + TY - 1 + P.st[s].yh);
    int step = 1;
    if (NC == 4) {
      step = 2;
      if ((jlo & 1) != (colour >> 1)) ++jlo;
    }
    // the rows of a warp advance by an even number: the column parity q of the colour, the validity of the two sites
    // of the lane and the stride of every pointer are invariants of the pass
    const int jw = jlo + warp * step;
    TCLK(tq1)
    if (jw <= jhi) {
      const int q = (NC == 2) ? ((colour ^ jw) & 1) : (colour & 1);
      const int i0 = gi0 + q;
      const bool v0 = (i0 >= ilo) && (i0 <= ihi), v1 = (i0 + 2 >= ilo) && (i0 + 2 <= ihi);
      if (v0 || v1) {
        const int nrows = (jhi - jw) / (kFusedWarps * step) + 1;
        double *xl = xs + (jw - j_r0) * 128 + lane;
        double *fl = fs + (jw - j_r0) * 128 + lane;
        const uint32_t c0 = (((uint32_t)jw * P.nz.G + pg) << 1) | (uint32_t)q, dc0 = ((uint32_t)(kFusedWarps * step) * P.nz.G) << 1;
        const int dl = kFusedWarps * step * 128;
        if (RESTRICT && s == P.res_stage && !lr_tile) {  // last pass before the residual (omega = 1): leaves the residual of its sites in fs
          if (q == 0) pass_rows<NINE, GIBBS, true, 0, true>(P, xl, fl, nrows, dl, c0, dc0, c1, smp, chain, ntab, v0, v1);
          else pass_rows<NINE, GIBBS, true, 1, true>(P, xl, fl, nrows, dl, c0, dc0, c1, smp, chain, ntab, v0, v1);
        } else if (P.omega_is_one) {
          if (q == 0) pass_rows<NINE, GIBBS, true, 0>(P, xl, fl, nrows, dl, c0, dc0, c1, smp, chain, ntab, v0, v1);
          else pass_rows<NINE, GIBBS, true, 1>(P, xl, fl, nrows, dl, c0, dc0, c1, smp, chain, ntab, v0, v1);
        } else {
          if (q == 0) pass_rows<NINE, GIBBS, false, 0>(P, xl, fl, nrows, dl, c0, dc0, c1, smp, chain, ntab, v0, v1);
          else pass_rows<NINE, GIBBS, false, 1>(P, xl, fl, nrows, dl, c0, dc0, c1, smp, chain, ntab, v0, v1);
        }
      }
    }
    TCLK(tq2)
    __syncthreads();
#ifdef MGMC_TILE_TIMING
    tacc_setup += tq1 - tq0;
    tacc_pass += tq2 - tq1;
    tacc_bar += clock64() - tq2;
#endif
   }
    if (seg < 2) { TSTAMP(2 + 2 * seg) }
    if (LOWRANK && seg < nsegfix && lr_tile) {
      MGMC_LR_PTRS
      const int fixq = seg;
      const LowRankTile &R = P.lr;
      const int dir = P.fix_dir[fixq];
      const bool diag = R.diag[dir] != 0;
      const size_t slot = (size_t)(P.lr_slot + fixq) * P.nchains + chain_idx;
      LrPkt *vb = R.vbuf + slot * 2 * lrm;
      // (1) owners: t_k = (B^T x)_k from shared memory, published as d_k (diagonal capacitance matrix) or as (t_k, s_k);
      //     the tile keeps its own values in shared memory
      for (int o = warp; o < lr_cnt[0]; o += kFusedWarps) {
        const int k = own_list[o];
        double acc = 0.0;
        for (int e = lane; e < R.EB; e += 32) {
          const int bi = R.b_i[k * R.EB + e] - i_r0, bj = R.b_j[k * R.EB + e] - j_r0;
          acc += R.b_val[k * R.EB + e] * xs[bj * 128 + (bi & 3) * 32 + (bi >> 2)];
        }
#pragma unroll
        for (int o2 = 16; o2 > 0; o2 >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o2);
        if (lane == 0) {
          const double sv = spre[fixq * lrm + k];
          const double v0 = diag ? fma(cms[fixq * lrm + k], sv, cmn[fixq * lrm + k] * acc) : acc;
          if (diag) {
            darr[k] = v0;
          } else {
            tarr[k] = v0;
            sarr[k] = sv;
          }
          pkt_store(vb + k, v0, lr_epoch);
          if (!diag) pkt_store(vb + lrm + k, sv, lr_epoch);
#pragma unroll
          for (int side = 0; side < 2; ++side) {
            if (!(side ? lr_to_up : lr_to_dn)) continue;
            LrPkt *pvb = reinterpret_cast<LrPkt *>(reinterpret_cast<char *>(vb) + (side ? P.sk.lr_peer_up : P.sk.lr_peer_dn));
            pkt_store(pvb + k, v0, lr_epoch);
            if (!diag) pkt_store(pvb + lrm + k, sv, lr_epoch);
          }
        }
      }
      const int n_need = lr_cnt[1 + dir];
      const int *nlist = need_list + dir * lrm;
      if (n_need > 0) {
        const int EW = R.EW[dir], nu = R.nu[dir];
        // the W entries of this thread are fetched before any flag is polled: afterwards only d is missing
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
        // (2) consumers: acquire what the owners published
        if (diag) {
          for (int n = threadIdx.x; n < n_need; n += kFusedThreads) {
            const int k = nlist[n];
            if (!is_own[k]) darr[k] = pkt_wait(vb + k, lr_epoch, P.err);
          }
        } else {
          // the measurements interact: every needed d_k is a full row of Ms s + Mneg t
          for (int k = threadIdx.x; k < lrm; k += kFusedThreads) {
            if (is_own[k]) continue;
            tarr[k] = pkt_wait(vb + k, lr_epoch, P.err);
            sarr[k] = pkt_wait(vb + lrm + k, lr_epoch, P.err);
          }
          __syncthreads();
          for (int n = warp; n < n_need; n += kFusedWarps) {
            const int k = nlist[n];
            double acc = 0.0;
            for (int c = lane; c < lrm; c += 32) {
              acc = fma(R.Mneg[dir][(size_t)k * lrm + c], tarr[c], acc);
              if (GIBBS) acc = fma(R.Ms[dir][(size_t)k * lrm + c], sarr[c], acc);
            }
#pragma unroll
            for (int o2 = 16; o2 > 0; o2 >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o2);
            if (lane == 0) darr[k] = acc;
          }
        }
        __syncthreads();
        // (3) x += W d  (columns this tile does not need keep d = 0: their W entries lie outside the region or are padding)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (pidx[q] < 0) continue;
          const int u = threadIdx.x + q * kFusedThreads;
          double acc = pval[q] * darr[pcol[q]];
          for (int e = 1; e < EW; ++e) acc += R.w_val[dir][(size_t)u * EW + e] * darr[R.w_col[dir][(size_t)u * EW + e]];
          xs[pidx[q]] += acc;
        }
        for (int u = threadIdx.x + 4 * kFusedThreads; u < nu; u += kFusedThreads) {
          const int i = R.w_i[dir][u], j = R.w_j[dir][u];
          if (!(i >= i_r0 && i < i_r0 + 128 && j >= j_r0 && j < j_r0 + RY)) continue;
          double acc = 0.0;
          for (int e = 0; e < EW; ++e) acc += R.w_val[dir][(size_t)u * EW + e] * darr[R.w_col[dir][(size_t)u * EW + e]];
          const int di = i - i_r0;
          xs[(j - j_r0) * 128 + (di & 3) * 32 + (di >> 2)] += acc;
        }
        __syncthreads();
        // Low-rank part of the residual for free: after the last fix-up x = x' + W d, hence Sigma^{-1} B^T x =
        // K (t + G s) = s - d  (d = (I - K G) s - K t, K = (Sigma + G)^{-1}): no second exchange for u.
        if (RESTRICT && P.lr_u_from_fix && seg == P.nfix - 1)
          for (int n = threadIdx.x; n < n_need; n += kFusedThreads) uarr[nlist[n]] = spre[fixq * lrm + nlist[n]] - darr[nlist[n]];
        for (int n = threadIdx.x; n < n_need; n += kFusedThreads) darr[nlist[n]] = 0.0;  // (ordered before the next use by the pass barriers)
      }
    }
    if (P.nqoi > 0 && seg < nbreak && s_end - 1 == P.qoi_stage && tile_qoi) {
      // the sample x^(k) of a merged launch: record the observed sites this tile holds (tile_qoi is uniform over the CTA)
      if ((int)threadIdx.x < P.nqoi) {
        const int qi = P.qoi_i[threadIdx.x], qj = P.qoi_j[threadIdx.x];
        if (qi >= i_t0 && qi < i_t0 + TX && qj >= j_t0 && qj < j_t0 + TY) {
          const int di = qi - i_r0;
          P.qoi_out[(size_t)chain_idx * P.nqoi + threadIdx.x] = xs[(qj - j_r0) * 128 + (di & 3) * 32 + (di >> 2)];
        }
      }
      __syncthreads();  // (the next pass may overwrite the site)
    }
    if (seg < 2) { TSTAMP(3 + 2 * seg) }
  }

  if (LOWRANK && RESTRICT && lr_own && !P.lr_u_from_fix) {
    // (launch without a closing fix-up) owners: u_k = (Sigma^{-1} B^T x)_k of the final iterate, for the low-rank part of the residual
    MGMC_LR_PTRS
    const LowRankTile &R = P.lr;
    const size_t slot = (size_t)(P.lr_slot + P.nfix) * P.nchains + chain_idx;
    LrPkt *vb = R.vbuf + slot * 2 * lrm;
    for (int o = warp; o < lr_cnt[0]; o += kFusedWarps) {
      const int k = own_list[o];
      double acc = 0.0;
      for (int e = lane; e < R.EB; e += 32) {
        const int bi = R.b_i[k * R.EB + e] - i_r0, bj = R.b_j[k * R.EB + e] - j_r0;
        acc += R.b_val[k * R.EB + e] * xs[bj * 128 + (bi & 3) * 32 + (bi >> 2)];
      }
#pragma unroll
      for (int o2 = 16; o2 > 0; o2 >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o2);
      if (lane == 0) {
        const double v0 = acc * R.sigma_inv[k];
        tarr[k] = v0;  // (kept for this tile's own residual)
        pkt_store(vb + k, v0, lr_epoch);
        if (lr_to_dn) pkt_store(reinterpret_cast<LrPkt *>(reinterpret_cast<char *>(vb + k) + P.sk.lr_peer_dn), v0, lr_epoch);
        if (lr_to_up) pkt_store(reinterpret_cast<LrPkt *>(reinterpret_cast<char *>(vb + k) + P.sk.lr_peer_up), v0, lr_epoch);
      }
    }
  }

  TSTAMP(7 + 0)
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

  TSTAMP(8)
  // ---- residual on [i_t0 - 1, i_t0 + TX - 1] x [j_t0, j_t0 + TY], then full-weighting restriction ----
  if (RESTRICT) {
    for (int rr = warp; rr <= TY; rr += kFusedWarps) {
      const int gj = j_t0 + rr;
      double *__restrict__ xr = xs + (gj - j_r0) * 128;
      double *__restrict__ fr = fs + (gj - j_r0) * 128;
      // sites of the last pass already hold their residual (unless this tile takes part in a low-rank fix-up)
      const int lastc = (P.res_stage >= 0 && !lr_tile) ? P.st[P.res_stage].colour : -1;
      int hq = -1;  // column parity of the sites of this row that hold their residual already
      if (lastc >= 0) {
        if (NC == 2) hq = (lastc ^ gj) & 1;
        else if ((gj & 1) == (lastc >> 1)) hq = lastc & 1;
      }
      const bool have0 = (hq == 0), have1 = (hq == 1);
      double res[4] = {0.0, 0.0, 0.0, 0.0};
      if (gj < ny) {  // (lanes 0 / 31 read in-bounds garbage for columns that are masked out below)
        res[0] = fr[lane];
        res[1] = fr[32 + lane];
        res[2] = fr[64 + lane];
        res[3] = fr[96 + lane];
        if (!have0) {  // (warp-uniform branches: the stencils are really skipped)
          res[0] -= stencil_at<NINE, 0>(P.a, xr, lane);
          res[2] -= stencil_at<NINE, 2>(P.a, xr, lane);
        }
        if (!have1) {
          res[1] -= stencil_at<NINE, 1>(P.a, xr, lane);
          res[3] -= stencil_at<NINE, 3>(P.a, xr, lane);
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int gi = gi0 + k;
        const bool ok = (gi >= max(1, i_t0 - 1)) && (gi <= min(nx - 1, i_t0 + TX - 1)) && (gj < ny);
        fr[k * 32 + lane] = ok ? res[k] : 0.0;
      }
    }
    __syncthreads();
    TSTAMP(9)
    if (lr_res) {
      // low-rank part of the residual: r -= B u, u = Sigma^{-1} B^T x of the final state (linear_operator.hh:71-75)
      MGMC_LR_PTRS
      const LowRankTile &R = P.lr;
      const size_t slot = (size_t)(P.lr_slot + P.nfix) * P.nchains + chain_idx;
      const LrPkt *vb = R.vbuf + slot * 2 * lrm;
      const int *nlist = need_list + 2 * lrm;
      for (int n = threadIdx.x; n < lr_cnt[3]; n += kFusedThreads) {
        const int k = nlist[n];
        darr[k] = P.lr_u_from_fix ? uarr[k] : (is_own[k] ? tarr[k] : pkt_wait(vb + k, lr_epoch, P.err));
      }
      __syncthreads();
      for (int u = threadIdx.x; u < R.nbu; u += kFusedThreads) {
        const int i = R.bu_i[u], j = R.bu_j[u];
        if (!(i >= max(1, i_t0 - 1) && i <= min(nx - 1, i_t0 + TX - 1) && j >= j_t0 && j <= j_t0 + TY && j < ny)) continue;
        double acc = 0.0;
        for (int e = R.bu_ptr[u]; e < R.bu_ptr[u + 1]; ++e) acc += R.bu_val[e] * darr[R.bu_col[e]];
        const int di = i - i_r0;
        fs[(j - j_r0) * 128 + (di & 3) * 32 + (di >> 2)] -= acc;
      }
      __syncthreads();
    }
    TSTAMP(10)
    const long long ccb = (long long)chain_idx * P.gc.stride;
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
        const unsigned int n_edge = (unsigned int)(P.sk.edge_rows * P.tiles_x);
        if (edge_dn) strip_arrive(P.sk.ticket_dn, n_edge, P.sk.peer_flag_dn);
        if (edge_up) strip_arrive(P.sk.ticket_up, n_edge, P.sk.peer_flag_up);
      }
    }
  }
  TSTAMP(11)
#ifdef MGMC_TILE_TIMING
  if (threadIdx.x == 0 && P.timing) {  // warp 0: cycles spent in pass set-up / rows / waiting at the pass barriers
    P.timing[(long long)cta_id * 16 + 12] = tacc_setup;
    P.timing[(long long)cta_id * 16 + 13] = tacc_pass;
    P.timing[(long long)cta_id * 16 + 14] = tacc_bar;
  }
#endif
}

template <int NC, bool GIBBS, bool PROLONG, bool RESTRICT, bool LOWRANK, bool NZG = false>
__global__ void __launch_bounds__(kFusedThreads, NZG ? 1 : MGMC_FUSED_MINBLOCKS) fused_smooth_kernel(const __grid_constant__ FusedP P) {  // (NZG: small levels, at most one CTA per SM -- no register cap)
  extern __shared__ double sm[];
  __shared__ int lr_cnt[4];
  __shared__ __align__(16) double ntab[128];
  if (GIBBS && threadIdx.x < 128) ntab[threadIdx.x] = kNormalTabDev[threadIdx.x];
  fused_tile<NC, GIBBS, PROLONG ? 1 : 0, RESTRICT ? 1 : 0, LOWRANK, NZG>(P, (int)blockIdx.x, (int)blockIdx.z + P.chain_off, (int)gridDim.x, sm, lr_cnt, ntab);
}

}  // namespace mgmc
