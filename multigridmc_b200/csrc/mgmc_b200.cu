// C-ABI implementation of the B200 MGMC hot path (include/mgmc_b200.h).
// Context = operator hierarchy resident in HBM + one CUDA stream; the multigrid recursion of the
// reference (multigridmc_sampler.cc:103-130, multigrid_preconditioner.cc:74-101) is unrolled on the
// host into a fixed launch sequence that is captured once into a CUDA graph and replayed per sample.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/mgmc_b200.h"
#include "fused.cuh"
#include "kernels.cuh"
#include "setup.hh"
#include "tail.cuh"
#include "noise_ahead.cuh"
#include "varcoef.cuh"
#include "lattice3d.cuh"
#include "rowfuse.cuh"
#include <climits>

#include <set>

using namespace mgmc;

static thread_local std::string g_last_error;
extern "C" const char *mgmc_last_error(void) { return g_last_error.c_str(); }

namespace {

struct MgmcError {
  int code;
  std::string msg;
};
[[noreturn]] void fail(int code, const std::string &msg) { throw MgmcError{code, msg}; }

#define CUDA_CHECK(expr)                                                                                   \
  do {                                                                                                     \
    cudaError_t err__ = (expr);                                                                            \
    if (err__ != cudaSuccess) fail(MGMC_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(err__)); \
  } while (0)

constexpr int GX = 16;  // doubles left of i = 0 (keeps i = 0 128-byte aligned)
constexpr int GY = 2;   // ghost rows below j = 0 and above j = ny

struct DevSparse {  // device copies of a sparse n x m matrix in both groupings
  SparseCols cols{0, nullptr, nullptr, nullptr};
  SparseRows rows{0, nullptr, nullptr, nullptr, nullptr};
};

struct LowRankDev {
  LowRankFix fix[2];  // [0] forward, [1] backward
  size_t smem = 0;
  // in-kernel fix-up (owner / consumer tiles of the fused kernel)
  LowRankTile tile;               // descriptor of the in-kernel fix-up (passed to the kernel by value)
  int bw = 0, bh = 0;             // largest extent of supp(B_k)
  int wreach = 0;                 // largest distance of a site of supp(W_k) from the bounding box of supp(B_k)
  bool diag[2] = {false, false};  // capacitance matrix diagonal (measurements do not interact on this level)
  // wide supports (a global measurement = a dense column of B): chip-wide kernels instead of the padded one-CTA ones
  bool wide = false;
  DevSparse W[2];                 // W = M_0^{-1} B per sweep direction (row grouping used)
};

struct DevLevel {
  HostLevel h;
  GridP g;
  double *x = nullptr, *f = nullptr, *r = nullptr;  // origin pointers (element i = 0, j = 0 of chain 0)
  double *x_alt = nullptr, *x_primary = nullptr;     // ping-pong partner of x / the buffer x must be in between calls
  Coef9 coef;
  bool nine = false;
  bool r2 = false;             // radius-2 stencil with position classes: generic kernels, 9 colours
  double *d_st = nullptr;      // [9][25] stencil classes (radius-2 levels)
  bool vc = false;             // per-vertex coefficients (variable kappa): generic kernels of varcoef.cuh, 2 / 4 colours
  VarCoef dvc{nullptr, 0, 0, 0, 0, 0};  // coefficient planes in the layout of the level vectors
  bool vc_full = false;        // ... all nine planes (the Galerkin operators); false: the 5-point fine operator, diagonal plane only
  bool d3 = false;             // 3d lattice: planes stacked in the row direction (lattice3d.cuh), generic kernels, 2 / 8 colours
  Grid3 q{0, 0};
  Coef27 c27;
  bool full27 = false;         // 27-point (Galerkin) operator; false: the 7-point fine operator
  bool generic() const { return r2 || vc || d3; }  // colour-by-colour launches instead of the fused tile kernel
  DevSparse B;
  bool lr_wide = false;        // some column of B has more entries than the padded low-rank kernels stage (kernels.cuh "Wide supports")
  std::map<double, LowRankDev> lowrank;  // keyed by omega
};

// Row-strip decomposition plan (same on every rank): rank r owns the rows (J0_r / 2^l, J1_r / 2^l] of the
// distributed levels l < ndist, J0_r = r * ny_0 / nranks; strips start on tile boundaries so that a
// rank simply runs its own tile rows.  Levels >= ndist are replicated.
struct StripPlan {
  int nranks = 1, rank = 0, ndist = 0;
  std::vector<int> lo, hi, halo;  // per distributed level: own rows [lo, hi], rows exchanged with a neighbour
  bool on() const { return nranks > 1; }
};

struct ProfSlot {
  std::string name;
  double ms = 0.0;
  int64_t launches = 0;
  double bytes = 0.0;  // algorithmic bytes summed over the launches
};

struct ProfEvent {
  std::string name;
  cudaEvent_t e0, e1;
  double bytes;
};

}  // namespace

struct mgmc_ctx {
  mgmc_desc d;
  std::vector<double> Sigma;
  int device = 0;
  int num_sms = 148;
  cudaStream_t stream = nullptr;
  std::vector<DevLevel> lv;
  std::vector<void *> allocs;
  // coarse factor
  int Nc = 0, Ncp = 0;
  double *dTT = nullptr;  // L^{-T} of the coarsest level (dense, row-major upper triangular)
  double *d_sigma_inv = nullptr, *d_sigma_inv_sqrt = nullptr, *d_sigma_inv_neg = nullptr;
  // in-kernel low-rank fix-up: slots of one cycle / API call (d vectors, exchange buffers, flags)
  static constexpr int kLrSlots = 1024;
  LrPkt *d_lr_vbuf = nullptr;   // [kLrSlots][nchains][2 m] (value, epoch) packets published by the owner tiles
  int *d_lr_epoch = nullptr;    // advanced at the start of every cycle / API call
  int lr_slot_next = 0;
  bool lr_fuse = true;        // MGMC_NO_LR_FUSE=1: separate fix-up launches (fallback path, perf experiments)
  // row-strip decomposition: arena shared with the neighbours through CUDA IPC
  StripPlan strip;
  char *arena = nullptr;
  size_t arena_bytes = 0;
  std::vector<char *> peer_arena;  // per rank (nullptr for self / unconnected)
  bool strip_connected = false;
  int *d_strip_ctl = nullptr;      // in the arena: [0] flag from below, [1] flag from above, [2] all-gather count,
                                   // [3] error, [5] waitno (all-gather), [6] all-gather push ticket, [7] / [8] tickets of
                                   // the edge CTAs (below / above), [9] cycle number
  int strip_index = 0;             // distributed fused launches emitted so far in this cycle
  int strip_per_cycle = 0;         // ... per cycle (counted by a dry run before the first capture)
  int *d_err = nullptr;  // error word: set by a device-side wait (low-rank packets, strip flags) that timed out
  // noise position
  uint32_t *d_sample = nullptr;
  uint32_t h_sample = 0;
  std::vector<uint32_t> sweep_counter;
  PhiloxKeys keys;
  // QoI / series
  std::vector<long long> h_qsite;  // what d_qsite / d_qval hold (mgmc_set_qoi returns early when nothing changes)
  std::vector<double> h_qval;
  int qoi_nnz = 0;
  long long *d_qsite = nullptr;
  double *d_qval = nullptr;
  double *d_series = nullptr;
  long long series_cap = 0;
  unsigned long long *d_pos = nullptr;
  // loop solver scratch
  double *sol_x = nullptr, *sol_b = nullptr, *d_partial = nullptr, *d_sol_hist = nullptr;
  SolverCtl *d_solver = nullptr;
  int npartial = 0, sol_hist_cap = 0;
  // moments
  double *d_mean = nullptr, *d_second = nullptr;
  // graph of one MGMC cycle (+ end-of-cycle kernel)
  cudaGraphExec_t graph = nullptr;
  // Merged level-0 launch: the post-smoothing of cycle k and the pre-smoothing of cycle k + 1 are adjacent launches on
  // level 0 -- one launch does both (prolongation, the sweeps of both, residual + restriction): one load / store of the
  // fine level and, with omega = 1, one colour pass less per cycle.  A run of K cycles is
  //   pre-smoothing, levels >= 1 | (K - 1) x { merged launch, end of cycle, levels >= 1 } | post-smoothing, end of cycle
  // and the unit in braces is the CUDA graph (one per parity of the ping-pong buffers of level 0).
  bool merge_on = false;
  int merge_qoi_stage = -1;          // >= 0 while the merged launch is being emitted: last stage of cycle k
  bool next_x_zero = false;          // the next fused launch starts from a zero iterate that nobody has zeroed (FusedP::x_in_zero)
  cudaGraphExec_t graph_unit[2] = {nullptr, nullptr};
  int64_t unit_launches = 0;
  double *d_qpart = nullptr;         // [nchains][kMaxQoi] observed sites recorded by the merged launch
  // per-tile low-rank flags of the fused launches (fused.cuh): a pool allocated with the context, handed out per launch geometry
  unsigned char *d_lr_flag_pool = nullptr;
  size_t lr_flag_pool_size = 0, lr_flag_pool_used = 0;
  std::map<std::string, size_t> lr_flag_slots;
  cudaGraphExec_t mg_graph = nullptr;  // one LoopSolver iteration: V-cycle, x -= Pr, next residual and its norm
  int64_t mg_graph_launches = 0;
  bool use_graph = true;
  // persistent kernel of the small levels (tail.cuh): levels >= tail_level and the coarse solve are phases of one launch
  int tail_level = -1;             // -1: not planned yet; nlevel: off
  bool tail_rec = false;           // the recursion below tail_level is being recorded instead of launched
  std::vector<TailPhase> tail_ph;
  size_t tail_smem = 0;
  double tail_bytes = 0.0;
  bool tail_lowrank = false;
  unsigned long long *d_tail_bar = nullptr;
  long long *d_tail_stamps = nullptr;  // MGMC_TAIL_STAMPS=1: per-phase time stamps of the last tail launch
  std::vector<int> tail_stamp_kinds;
  double *dAinv = nullptr;         // A^{-1} of the coarsest level (one-pass coarse phase)
  double *d_lr_partial = nullptr, *d_lr_d = nullptr;  // wide low-rank supports: partial sums of B^T x, coefficients d
  double *d_coarse_xi = nullptr;   // ensembles: normals of the coarse sampler for all chains (coarse_xi_kernel)
  std::set<const void *> func_attr_done;  // kernels whose dynamic shared memory limit has been raised on this device
  // noise of the small levels generated ahead of their launches (noise_ahead.cuh): a second branch of the cycle graph
  // generates the normals of the latency-bound levels while the big levels run
  struct NzaSlot {
    int level = 0;
    std::vector<NzJob> jobs;         // one per full colour pass of the launch
    std::vector<int> stage;          // stage index of jobs[k] in the launch
  };
  int nza_planned = 0;               // 0: not yet, 1: planned (nza_on says whether it is in use)
  bool nza_on = false;
  int nza_fork_level = -1;           // the branch forks before the first launch of this level
  std::vector<int> nza_levels;
  std::vector<NzaSlot> nza_slots;    // in launch order per level
  std::vector<int> nza_cursor;       // per level: launches of the level emitted so far in this cycle
  bool nza_dry = false;              // planning run: dev_fused records the launch instead of emitting it
  bool nza_forked = false, nza_joined = false, nza_in_cycle = false;
  cudaStream_t stream2 = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_gen = nullptr;
  bool perf_no_noise = false;  // MGMC_PERF_NO_NOISE=1: run the sampling cycle with the deterministic kernels (perf experiments only)
  // instrumentation
  int64_t launch_count = 0;
  int64_t launches_per_cycle = 0;
  bool prof_on = false;
  std::vector<ProfEvent> prof_events;

  template <class T>
  T *dalloc(size_t n, bool zero = true) {
    void *p = nullptr;
    CUDA_CHECK(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
    allocs.push_back(p);
    if (zero) CUDA_CHECK(cudaMemsetAsync(p, 0, std::max<size_t>(n, 1) * sizeof(T), stream));
    return (T *)p;
  }
  template <class T>
  T *dupload(const std::vector<T> &v) {
    T *p = dalloc<T>(v.size(), false);
    if (!v.empty()) CUDA_CHECK(cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, stream));
    CUDA_CHECK(cudaStreamSynchronize(stream));  // v may be a temporary
    return p;
  }
  void sync() { CUDA_CHECK(cudaStreamSynchronize(stream)); }
  void dfree(void *p) {
    if (!p) return;
    for (size_t k = 0; k < allocs.size(); ++k)
      if (allocs[k] == p) {
        allocs.erase(allocs.begin() + k);
        break;
      }
    cudaFree(p);
  }

  // alg_bytes: algorithmic bytes of this launch in the model of SURVEY.md section 8(d)
  template <class F>
  void launch(const char *name, int level, F &&fn, double alg_bytes = 0.0) {
    ++launch_count;
    if (prof_on) {
      cudaEvent_t e0, e1;
      CUDA_CHECK(cudaEventCreate(&e0));
      CUDA_CHECK(cudaEventCreate(&e1));
      CUDA_CHECK(cudaEventRecord(e0, stream));
      fn();
      CUDA_CHECK(cudaEventRecord(e1, stream));
      prof_events.push_back({std::string(name) + "/L" + std::to_string(level), e0, e1, alg_bytes});
    } else {
      fn();
    }
    CUDA_CHECK(cudaGetLastError());
  }
};

namespace {

dim3 grid_sites(const GridP &g, int nch) { return dim3((g.nx - 1 + 63) / 64, (g.ny - 1 + 3) / 4, nch); }
const dim3 kBlockSites(64, 4, 1);

DevSparse upload_sparse(mgmc_ctx *c, const std::vector<SEntry> &E, int m, int pitch) {
  DevSparse out;
  // column grouping
  std::vector<int> colptr(m + 1, 0);
  for (const SEntry &e : E) colptr[e.col + 1]++;
  for (int k = 0; k < m; ++k) colptr[k + 1] += colptr[k];
  std::vector<long long> site(E.size());
  std::vector<double> val(E.size());
  std::vector<int> fill(colptr.begin(), colptr.end() - 1);
  for (const SEntry &e : E) {
    const int p = fill[e.col]++;
    site[p] = (long long)e.j * pitch + e.i;
    val[p] = e.val;
  }
  out.cols.m = m;
  out.cols.colptr = c->dupload(colptr);
  out.cols.site = c->dupload(site);
  out.cols.val = c->dupload(val);
  // row grouping (unique sites, deterministic scatter)
  std::map<long long, std::vector<std::pair<int, double>>> bysite;
  for (const SEntry &e : E) bysite[(long long)e.j * pitch + e.i].push_back({e.col, e.val});
  std::vector<long long> usite;
  std::vector<int> uptr(1, 0), ucol;
  std::vector<double> uval;
  for (auto &kv : bysite) {
    usite.push_back(kv.first);
    for (auto &cv : kv.second) {
      ucol.push_back(cv.first);
      uval.push_back(cv.second);
    }
    uptr.push_back((int)ucol.size());
  }
  out.rows.nu = (int)usite.size();
  out.rows.usite = c->dupload(usite);
  out.rows.uptr = c->dupload(uptr);
  out.rows.ucol = c->dupload(ucol);
  out.rows.uval = c->dupload(uval);
  return out;
}

Coef9 to_coef9(const StencilSet &s) {
  Coef9 a;
  a.c = s.at(4, 0, 0);
  a.w = s.at(4, -1, 0);
  a.e = s.at(4, 1, 0);
  a.s = s.at(4, 0, -1);
  a.n = s.at(4, 0, 1);
  a.sw = s.at(4, -1, -1);
  a.se = s.at(4, 1, -1);
  a.nw = s.at(4, -1, 1);
  a.ne = s.at(4, 1, 1);
  return a;
}

inline int fused_tile_rows(int ny, int nc, bool strips, bool restrict_, bool merged = false, int nchains = 1);

StripPlan make_strip_plan(const mgmc_desc &d, const std::vector<HostLevel> &H) {
  StripPlan p;
  p.nranks = std::max(d.strip_nranks, 1);
  p.rank = d.strip_rank;
  if (p.nranks == 1) return p;
  if (p.rank < 0 || p.rank >= p.nranks) fail(MGMC_ERR_INVALID, "strip_rank out of range");
  if (d.nchains != 1) fail(MGMC_ERR_UNSUPPORTED, "row strips advance one chain (nchains = 1)");
  static const char *mr = std::getenv("MGMC_STRIP_MIN_ROWS");
  const int min_rows = mr ? std::atoi(mr) : 64;
  // levels below ~1M sites are latency bound on one GPU already: splitting them only adds NVLink round trips
  static const char *ms = std::getenv("MGMC_STRIP_MIN_SITES");
  const long long min_sites = ms ? std::atoll(ms) : (1ll << 20);
  for (int l = 0; l + 1 < (int)H.size(); ++l) {  // the coarsest level is always replicated
    const HostLevel &h = H[l];
    const bool r2 = h.st.radius > 1;  // radius-2 operators: per-colour exchange inside the colour launches (kernels.cuh StripR2)
    if (r2 && d.m_lowrank > 0) break;
    const int nc = h.st.ncolours;
    const int ty = r2 ? 2 : fused_tile_rows(h.ny, nc, true, true);
    const int rows = h.ny / p.nranks;
    // rows exchanged with a neighbour: what a launch of 2 sweeps (+ residual) reads beyond the own rows; with a
    // low-rank term additionally the windows of the measurements near the strip boundary.  Radius 2: the stencil
    // reach (2) + the residual row beyond the strip (1) + prolongation of the mirrored rows.
    const int halo = r2 ? 4 : ((d.m_lowrank > 0) ? ((nc == 2) ? 12 : 20) : ((nc == 2) ? 8 : 12));
    if (h.ny % p.nranks || rows % ty || rows < std::max(min_rows, 2 * halo)) break;
    if (l > 0 && (long long)h.nx * h.ny < min_sites) break;
    // every rank runs the measurement windows near its strip on its own: only while the measurements do not interact
    if (d.m_lowrank > 0 && !lowrank_is_diagonal(h, std::vector<double>(d.Sigma, d.Sigma + d.m_lowrank), d.omega)) break;
    p.lo.push_back(p.rank * rows + 1);
    p.hi.push_back(std::min((p.rank + 1) * rows, h.ny - 1));
    p.halo.push_back(halo);
    p.ndist = l + 1;
  }
  if (p.ndist == 0) fail(MGMC_ERR_UNSUPPORTED, "lattice cannot be split into row strips for this number of ranks (rows per rank must be a multiple of the tile height and >= MGMC_STRIP_MIN_ROWS)");
  return p;
}

std::vector<HostLevel> build_host_levels(const mgmc_desc &d) {
  if (d.dim != 2 && d.dim != 3) fail(MGMC_ERR_UNSUPPORTED, "only dim = 2 and dim = 3 lattices are implemented on the device path");
  if (d.nx < 2 || d.ny < 2 || (d.dim == 3 && d.nz < 2)) fail(MGMC_ERR_INVALID, "invalid lattice size");
  if (d.nlevel < 1) fail(MGMC_ERR_INVALID, "nlevel must be >= 1");
  if (d.pde_model != MGMC_PDE_SHIFTEDLAPLACE_FD && d.pde_model != MGMC_PDE_SQUARED_SHIFTEDLAPLACE_FD && d.pde_model != MGMC_PDE_SHIFTEDLAPLACE_FEM)
    fail(MGMC_ERR_INVALID, "invalid pde_model");
  if (!(d.Lambda > 0.0)) fail(MGMC_ERR_INVALID, "Lambda must be positive");
  if (!(d.omega > 0.0 && d.omega < 2.0)) fail(MGMC_ERR_INVALID, "omega must be in (0,2)");
  if (d.nx >= (1 << 20) || d.ny >= (1 << 20)) fail(MGMC_ERR_INVALID, "lattice too large");
  // Philox counter word 0 of a site is ((j G + i / 4) << 1) | (i & 1), G = nx / 4 + 1; 0x40000000 / 0x80000000 tag the
  // coarse-sampler / low-rank streams (philox.cuh): the site counters must stay below 2^30
  if (2ll * ((long long)d.ny + 1) * ((long long)d.nx / 4 + 1) >= (1ll << 30)) fail(MGMC_ERR_UNSUPPORTED, "lattice too large for the 32-bit site counter of the noise streams (about 46000 x 46000)");
  std::vector<HostLevel> L(d.nlevel);
  if (d.dim == 3) {
    // Lattice3d (lattice/lattice3d.hh): shiftedlaplace_fd with a constant correlation length, no measurements yet
    if (d.pde_model != MGMC_PDE_SHIFTEDLAPLACE_FD && d.pde_model != MGMC_PDE_SHIFTEDLAPLACE_FEM)
      fail(MGMC_ERR_UNSUPPORTED, "3d lattices are implemented for shiftedlaplace_fd and shiftedlaplace_fem only");
    if (d.kappa_sq) fail(MGMC_ERR_UNSUPPORTED, "3d lattices: a variable correlation length is not implemented");
    if (std::max(d.strip_nranks, 1) > 1) fail(MGMC_ERR_UNSUPPORTED, "3d lattices: row strips are not implemented");
    if (d.nz >= (1 << 20)) fail(MGMC_ERR_INVALID, "lattice too large");
    // (the stacked planes are the y dimension of the launch grids: 4 rows per CTA, at most 65535 CTAs)
    if (((long long)d.nz + 1) * ((long long)d.ny + 1) > 4ll * 65535) fail(MGMC_ERR_UNSUPPORTED, "3d lattice too large for the launch geometry of the first 3d path ((ny + 1)(nz + 1) <= 262140)");
    if (2ll * ((long long)d.nz + 1) * ((long long)d.ny + 1) * ((long long)d.nx / 4 + 1) >= (1ll << 30)) fail(MGMC_ERR_UNSUPPORTED, "lattice too large for the 32-bit site counter of the noise streams");
    for (int l = 0; l < d.nlevel; ++l) {
      HostLevel &h = L[l];
      std::memset(h.st.a, 0, sizeof(h.st.a));
      if (l == 0) {
        h.nx = d.nx;
        h.ny = d.ny;
        h.nz = d.nz;
        if (d.pde_model == MGMC_PDE_SHIFTEDLAPLACE_FEM) {
          const int n3[3] = {d.nx, d.ny, d.nz};
          fem_stencil(3, n3, d.Lambda, h.st3);
        } else {
          fine_stencil3(d.nx, d.ny, d.nz, d.Lambda, h.st3);
        }
        const long long w3 = d.nx - 1, h3 = d.ny - 1;
        for (int64_t e = 0; e < d.B_nnz; ++e) {
          const int64_t row = d.B_rows[e];
          if (row < 0 || row >= w3 * h3 * (d.nz - 1) || d.B_cols[e] < 0 || d.B_cols[e] >= d.m_lowrank) fail(MGMC_ERR_INVALID, "B entry out of range");
          // (entries of B / W on a 3d lattice: (i, row of the stacked planes), setup.hh coarsen_B3)
          h.B.push_back({int(row % w3) + 1, int(row / (w3 * h3) + 1) * (d.ny + 1) + int((row / w3) % h3) + 1, d.B_cols[e], d.B_vals[e]});
        }
      } else {
        const HostLevel &f = L[l - 1];
        // Lattice3d::get_coarse_lattice (lattice3d.hh:242-257)
        if ((f.nx % 2) || (f.ny % 2) || (f.nz % 2)) fail(MGMC_ERR_INVALID, "cannot coarsen lattice of size " + std::to_string(f.nx) + " x " + std::to_string(f.ny) + " x " + std::to_string(f.nz) + " [one of the extents is odd]");
        if (!(f.nx / 2 > 1 && f.ny / 2 > 1 && f.nz / 2 > 1)) fail(MGMC_ERR_INVALID, "cannot coarsen lattice [resulting lattice would have no interior vertices]");
        h.nx = f.nx / 2;
        h.ny = f.ny / 2;
        h.nz = f.nz / 2;
        coarsen_stencil3(f.st3, h.st3);
        h.B = coarsen_B3(f.B, f.nx, f.ny, f.nz);
      }
      h.st.radius = 1;
      h.st.uniform = true;
      h.st.ncolours = colours_stencil3(h.st3);
    }
    return L;
  }
  L[0].nx = d.nx;
  L[0].ny = d.ny;
  L[0].st = fine_stencil(d.pde_model, d.nx, d.ny, d.Lambda);
  if (d.kappa_sq) {
    // correlation length that varies in space: kappa^2 per interior vertex instead of 1 / Lambda^2
    if (d.pde_model != MGMC_PDE_SHIFTEDLAPLACE_FD)
      fail(MGMC_ERR_UNSUPPORTED, "a variable correlation length is implemented for shiftedlaplace_fd only (the FEM operator evaluates kappa^2 at the quadrature points)");
    if (std::max(d.strip_nranks, 1) > 1) fail(MGMC_ERR_UNSUPPORTED, "row strips of an operator with per-vertex coefficients are not implemented");
    for (long long k = 0; k < (long long)(d.nx - 1) * (d.ny - 1); ++k)
      if (!(d.kappa_sq[k] >= 0.0)) fail(MGMC_ERR_INVALID, "kappa_sq entries must be non-negative");
    fine_varcoef(L[0], d.kappa_sq);
  }
  const int w = d.nx - 1;
  for (int64_t e = 0; e < d.B_nnz; ++e) {
    const int64_t row = d.B_rows[e];
    if (row < 0 || row >= (int64_t)(d.nx - 1) * (d.ny - 1) || d.B_cols[e] < 0 || d.B_cols[e] >= d.m_lowrank) fail(MGMC_ERR_INVALID, "B entry out of range");
    L[0].B.push_back({int(row % w) + 1, int(row / w) + 1, d.B_cols[e], d.B_vals[e]});
  }
  for (int l = 1; l < d.nlevel; ++l) {
    const HostLevel &f = L[l - 1];
    // Lattice2d::get_coarse_lattice (lattice2d.hh:198-213)
    if ((f.nx % 2) || (f.ny % 2)) fail(MGMC_ERR_INVALID, "cannot coarsen lattice of size " + std::to_string(f.nx) + " x " + std::to_string(f.ny) + " [one of the extents is odd]");
    if (!(f.nx / 2 > 1 && f.ny / 2 > 1)) fail(MGMC_ERR_INVALID, "cannot coarsen lattice [resulting lattice would have no interior vertices]");
    L[l].nx = f.nx / 2;
    L[l].ny = f.ny / 2;
    if (f.varcoef()) coarsen_varcoef(f, L[l]);
    else L[l].st = coarsen_stencil(f.st, f.nx, f.ny);
    L[l].B = coarsen_B(f.B, f.nx, f.ny);
  }
  return L;
}

const LowRankDev &get_lowrank(mgmc_ctx *c, int level, double omega) {
  DevLevel &L = c->lv[level];
  auto it = L.lowrank.find(omega);
  if (it != L.lowrank.end()) return it->second;
  LowRankDev dev;
  const int m = c->d.m_lowrank;
  if (L.lr_wide) {
    // a dense column of B (global measurement): W and the m x m matrices only, for the chip-wide kernels
    dev.wide = true;
    dev.bw = dev.bh = 1 << 20;  // (never fusable into the tile kernel)
    for (int dir = 0; dir < 2; ++dir) {
      LowRankDir h = lowrank_setup(L.h, c->Sigma, omega, dir == 0);
      dev.W[dir] = upload_sparse(c, h.W, m, L.g.pitch);
      std::memset(&dev.fix[dir], 0, sizeof(dev.fix[dir]));
      dev.fix[dir].m = m;
      dev.fix[dir].Mneg = c->dupload(h.Mneg);
      dev.fix[dir].Ms = c->dupload(h.Ms);
      dev.fix[dir].sigma_inv_sqrt = c->d_sigma_inv_sqrt;
    }
    return L.lowrank.emplace(omega, dev).first->second;
  }
  // B padded to EB entries per column
  std::vector<std::vector<SEntry>> cols(m);
  for (const SEntry &e : L.h.B) cols[e.col].push_back(e);
  int EB = 1;
  for (auto &v : cols) EB = std::max<int>(EB, (int)v.size());
  std::vector<long long> bsite((size_t)m * EB, 0);
  std::vector<double> bval((size_t)m * EB, 0.0);
  for (int k = 0; k < m; ++k)
    for (int e = 0; e < EB; ++e) {
      const bool has = e < (int)cols[k].size();
      const SEntry &src = has ? cols[k][e] : (cols[k].empty() ? SEntry{1, 1, k, 0.0} : cols[k][0]);
      bsite[(size_t)k * EB + e] = (long long)src.j * L.g.pitch + src.i;
      bval[(size_t)k * EB + e] = has ? src.val : 0.0;
    }
  const long long *d_bsite = c->dupload(bsite);
  const double *d_bval = c->dupload(bval);
  for (int dir = 0; dir < 2; ++dir) {
    LowRankDir h = lowrank_setup(L.h, c->Sigma, omega, dir == 0);
    std::map<long long, std::vector<std::pair<int, double>>> bysite;
    for (const SEntry &e : h.W) bysite[(long long)e.j * L.g.pitch + e.i].push_back({e.col, e.val});
    int EW = 1;
    for (auto &kv : bysite) EW = std::max<int>(EW, (int)kv.second.size());
    std::vector<long long> usite;
    std::vector<int> wcol;
    std::vector<double> wval;
    for (auto &kv : bysite) {
      usite.push_back(kv.first);
      for (int e = 0; e < EW; ++e) {
        const bool has = e < (int)kv.second.size();
        wcol.push_back(has ? kv.second[e].first : 0);
        wval.push_back(has ? kv.second[e].second : 0.0);
      }
    }
    LowRankFix &F = dev.fix[dir];
    F.m = m;
    F.EB = EB;
    F.nu = (int)usite.size();
    F.EW = EW;
    F.bsite = d_bsite;
    F.bval = d_bval;
    F.usite = c->dupload(usite);
    F.wcol = c->dupload(wcol);
    F.wval = c->dupload(wval);
    F.Mneg = c->dupload(h.Mneg);
    F.Ms = c->dupload(h.Ms);
    F.sigma_inv_sqrt = c->d_sigma_inv_sqrt;
    F.mats_in_smem = (m <= 48) ? 1 : 0;
  }
  dev.smem = ((size_t)m * EB + 3 * (size_t)m + (m <= 48 ? 2 * (size_t)m * m : 0)) * sizeof(double);
  {
    // descriptor of the in-kernel fix-up: same matrices with explicit (i, j) coordinates
    LowRankTile T;
    std::memset(&T, 0, sizeof(T));
    T.m = m;
    T.EB = EB;
    std::vector<int> bi((size_t)m * EB), bj((size_t)m * EB), bbox((size_t)m * 4);
    for (int k = 0; k < m; ++k) {
      int i0 = 1 << 30, i1 = -1, j0 = 1 << 30, j1 = -1;
      for (int e = 0; e < EB; ++e) {
        const bool has = e < (int)cols[k].size();
        const SEntry &src = has ? cols[k][e] : (cols[k].empty() ? SEntry{1, 1, k, 0.0} : cols[k][0]);
        bi[(size_t)k * EB + e] = src.i;
        bj[(size_t)k * EB + e] = src.j;
        i0 = std::min(i0, src.i);
        i1 = std::max(i1, src.i);
        j0 = std::min(j0, src.j);
        j1 = std::max(j1, src.j);
      }
      bbox[4 * k] = i0;
      bbox[4 * k + 1] = i1;
      bbox[4 * k + 2] = j0;
      bbox[4 * k + 3] = j1;
      dev.bw = std::max(dev.bw, i1 - i0 + 1);
      dev.bh = std::max(dev.bh, j1 - j0 + 1);
    }
    T.b_i = c->dupload(bi);
    T.b_j = c->dupload(bj);
    T.b_val = d_bval;
    T.bbox = c->dupload(bbox);
    {
      std::map<std::pair<int, int>, std::vector<std::pair<int, double>>> bysite;  // (j, i) -> (col, val)
      for (const SEntry &e : L.h.B) bysite[{e.j, e.i}].push_back({e.col, e.val});
      std::vector<int> ui, uj, uptr(1, 0), ucol;
      std::vector<double> uval;
      for (auto &kv : bysite) {
        ui.push_back(kv.first.second);
        uj.push_back(kv.first.first);
        for (auto &cv : kv.second) {
          ucol.push_back(cv.first);
          uval.push_back(cv.second);
        }
        uptr.push_back((int)ucol.size());
      }
      T.nbu = (int)ui.size();
      T.bu_i = c->dupload(ui);
      T.bu_j = c->dupload(uj);
      T.bu_ptr = c->dupload(uptr);
      T.bu_col = c->dupload(ucol);
      T.bu_val = c->dupload(uval);
    }
    for (int dir = 0; dir < 2; ++dir) {
      const LowRankFix &F = dev.fix[dir];
      std::vector<long long> usite(F.nu);
      CUDA_CHECK(cudaMemcpy(usite.data(), F.usite, sizeof(long long) * F.nu, cudaMemcpyDeviceToHost));
      std::vector<int> wi(F.nu), wj(F.nu);
      for (int u = 0; u < F.nu; ++u) {
        wi[u] = (int)(usite[u] % L.g.pitch);
        wj[u] = (int)(usite[u] / L.g.pitch);
      }
      T.nu[dir] = F.nu;
      T.EW[dir] = F.EW;
      T.w_i[dir] = c->dupload(wi);
      T.w_j[dir] = c->dupload(wj);
      T.w_col[dir] = F.wcol;
      T.w_val[dir] = F.wval;
      T.Mneg[dir] = F.Mneg;
      T.Ms[dir] = F.Ms;
      // bounding boxes of supp(W_k), W sites near every measurement, diagonality of the capacitance matrix
      std::vector<int> wcolh((size_t)F.nu * F.EW);
      std::vector<double> wvalh((size_t)F.nu * F.EW), Mn((size_t)m * m), Msh((size_t)m * m);
      CUDA_CHECK(cudaMemcpy(wcolh.data(), F.wcol, sizeof(int) * wcolh.size(), cudaMemcpyDeviceToHost));
      CUDA_CHECK(cudaMemcpy(wvalh.data(), F.wval, sizeof(double) * wvalh.size(), cudaMemcpyDeviceToHost));
      CUDA_CHECK(cudaMemcpy(Mn.data(), F.Mneg, sizeof(double) * Mn.size(), cudaMemcpyDeviceToHost));
      CUDA_CHECK(cudaMemcpy(Msh.data(), F.Ms, sizeof(double) * Msh.size(), cudaMemcpyDeviceToHost));
      std::vector<int> wbox((size_t)m * 4);
      for (int k = 0; k < m; ++k) {
        wbox[4 * k] = wbox[4 * k + 2] = 1 << 30;
        wbox[4 * k + 1] = wbox[4 * k + 3] = -(1 << 30);
      }
      for (int u = 0; u < F.nu; ++u)
        for (int e = 0; e < F.EW; ++e) {
          if (wvalh[(size_t)u * F.EW + e] == 0.0) continue;
          const int k = wcolh[(size_t)u * F.EW + e];
          wbox[4 * k] = std::min(wbox[4 * k], wi[u]);
          wbox[4 * k + 1] = std::max(wbox[4 * k + 1], wi[u]);
          wbox[4 * k + 2] = std::min(wbox[4 * k + 2], wj[u]);
          wbox[4 * k + 3] = std::max(wbox[4 * k + 3], wj[u]);
        }
      T.wbox[dir] = c->dupload(wbox);
      for (int k = 0; k < m; ++k)
        if (wbox[4 * k + 1] >= wbox[4 * k])
          dev.wreach = std::max(dev.wreach, std::max(std::max(bbox[4 * k] - wbox[4 * k], wbox[4 * k + 1] - bbox[4 * k + 1]),
                                                     std::max(bbox[4 * k + 2] - wbox[4 * k + 2], wbox[4 * k + 3] - bbox[4 * k + 3])));
      bool diag = true;
      for (int r = 0; r < m && diag; ++r)
        for (int q = 0; q < m; ++q)
          if (r != q && (Mn[(size_t)r * m + q] != 0.0 || Msh[(size_t)r * m + q] != 0.0)) {
            diag = false;
            break;
          }
      T.diag[dir] = diag ? 1 : 0;
      dev.diag[dir] = diag;
    }
    T.sigma_inv = c->d_sigma_inv;
    T.sigma_inv_sqrt = c->d_sigma_inv_sqrt;
    if (!c->d_lr_vbuf) fail(MGMC_ERR_INVALID, "internal: low-rank buffers missing");
    T.vbuf = c->d_lr_vbuf;
    T.epoch = c->d_lr_epoch;
    dev.tile = T;
  }
  return L.lowrank.emplace(omega, dev).first->second;
}

// ---------------------------------------------------------------------------------------------
// host <-> device transfers: lexicographic interior vector <-> padded device layout
// ---------------------------------------------------------------------------------------------
void upload_vec(mgmc_ctx *c, int level, double *dev, const double *host) {
  const DevLevel &L = c->lv[level];
  if (L.d3) {  // plane by plane (lexicographic: lattice3d.hh:122-135)
    const size_t w = L.g.nx - 1, h = L.q.ny - 1, d = L.q.nz - 1;
    for (int ch = 0; ch < c->d.nchains; ++ch)
      for (size_t k = 1; k <= d; ++k)
        CUDA_CHECK(cudaMemcpy2DAsync(dev + (size_t)ch * L.g.stride + (k * (L.q.ny + 1) + 1) * L.g.pitch + 1, L.g.pitch * sizeof(double),
                                     host + ((size_t)ch * d + (k - 1)) * w * h, w * sizeof(double), w * sizeof(double), h, cudaMemcpyHostToDevice, c->stream));
    return;
  }
  const size_t w = L.g.nx - 1, h = L.g.ny - 1;
  for (int ch = 0; ch < c->d.nchains; ++ch)
    CUDA_CHECK(cudaMemcpy2DAsync(dev + (size_t)ch * L.g.stride + L.g.pitch + 1, L.g.pitch * sizeof(double), host + (size_t)ch * w * h, w * sizeof(double),
                                 w * sizeof(double), h, cudaMemcpyHostToDevice, c->stream));
}
void download_vec(mgmc_ctx *c, int level, const double *dev, double *host) {
  const DevLevel &L = c->lv[level];
  if (L.d3) {
    const size_t w = L.g.nx - 1, h = L.q.ny - 1, d = L.q.nz - 1;
    for (int ch = 0; ch < c->d.nchains; ++ch)
      for (size_t k = 1; k <= d; ++k)
        CUDA_CHECK(cudaMemcpy2DAsync(host + ((size_t)ch * d + (k - 1)) * w * h, w * sizeof(double),
                                     dev + (size_t)ch * L.g.stride + (k * (L.q.ny + 1) + 1) * L.g.pitch + 1, L.g.pitch * sizeof(double), w * sizeof(double), h,
                                     cudaMemcpyDeviceToHost, c->stream));
    return;
  }
  const size_t w = L.g.nx - 1, h = L.g.ny - 1;
  for (int ch = 0; ch < c->d.nchains; ++ch)
    CUDA_CHECK(cudaMemcpy2DAsync(host + (size_t)ch * w * h, w * sizeof(double), dev + (size_t)ch * L.g.stride + L.g.pitch + 1, L.g.pitch * sizeof(double),
                                 w * sizeof(double), h, cudaMemcpyDeviceToHost, c->stream));
}

NoiseP noise_params(mgmc_ctx *c, int level, uint32_t c1) {
  NoiseP nz;
  nz.keys = c->keys;
  nz.mc = kNormalConstsHost;
  nz.c1 = c1;
  nz.sample = c->d_sample;
  nz.chain0 = (uint32_t)c->d.first_chain;
  nz.G = (uint32_t)(c->lv[level].g.nx / 4 + 1);
  return nz;
}
uint32_t next_c1(mgmc_ctx *c, int level, bool advance) {
  const uint32_t c1 = ((uint32_t)level << 24) | (c->sweep_counter[level] & 0xFFFFFFu);
  if (advance) c->sweep_counter[level]++;
  return c1;
}

// ---------------------------------------------------------------------------------------------
// single-level building blocks (device vectors)
// ---------------------------------------------------------------------------------------------
void tail_push_simple(mgmc_ctx *c, int kind, const GridP &g, const double *src, double *dst) {
  TailPhase ph;
  std::memset(&ph, 0, sizeof(ph));
  ph.kind = kind;
  ph.P.g = g;
  ph.P.x_in = src;
  ph.P.x_out = dst;
  c->tail_ph.push_back(ph);
}

void dev_zero(mgmc_ctx *c, int level, double *x) {
  const DevLevel &L = c->lv[level];
  if (c->tail_rec) return tail_push_simple(c, TAIL_ZERO, L.g, nullptr, x);
  c->launch("zero", level, [&] { axpy_kernel<1><<<grid_sites(L.g, c->d.nchains), kBlockSites, 0, c->stream>>>(L.g, x, nullptr); });
}

// keep the invariant "between API calls / cycles the iterate of a level lives in its primary buffer"
// (the fused sweeps are out of place and ping-pong between x and x_alt)
void normalize_x(mgmc_ctx *c, int level) {
  DevLevel &L = c->lv[level];
  if (L.x == L.x_primary) return;
  if (c->tail_rec) {
    tail_push_simple(c, TAIL_COPY, L.g, L.x, L.x_primary);
    std::swap(L.x, L.x_alt);
    return;
  }
  c->launch("copy_back", level, [&] { axpy_kernel<2><<<grid_sites(L.g, c->d.nchains), kBlockSites, 0, c->stream>>>(L.g, L.x_primary, L.x); });
  std::swap(L.x, L.x_alt);
}

// ---- wide low-rank supports (kernels.cuh "Wide supports") ----
// partial sums of B^T x (B of `level`) for every chain of x
void dev_lowrank_wide_partial(mgmc_ctx *c, int level, const double *x, int nch) {
  const DevLevel &L = c->lv[level];
  const int m = c->d.m_lowrank;
  if (!c->d_lr_partial) fail(MGMC_ERR_INVALID, "internal: buffers of the wide low-rank kernels missing");
  c->launch("lowrank_bt", level, [&] { lowrank_bt_partial_kernel<<<dim3(kLrWideBlocks, m, nch), 256, 0, c->stream>>>(L.B.cols, L.g.stride, x, c->d_lr_partial); });
}
// y += sign * R diag(scale) B^T x with R = the row grouping of B on this level or of the coarse B (restriction)
void dev_lowrank_wide_apply(mgmc_ctx *c, int level, const double *x, const SparseRows &R, long long stride_y, double *y, const double *scale, double sign, int nch) {
  const int m = c->d.m_lowrank;
  dev_lowrank_wide_partial(c, level, x, nch);
  c->launch("lowrank_d", level, [&] {
    lowrank_d_kernel<0><<<nch, 256, 2 * m * sizeof(double), c->stream>>>(m, kLrWideBlocks, c->d_lr_partial, scale, nullptr, nullptr, nullptr, NoiseP{}, c->d_lr_d);
  });
  c->launch("lowrank_scatter", level, [&] { lowrank_scatter_kernel<<<dim3((R.nu + 255) / 256, nch), 256, 0, c->stream>>>(R, m, c->d_lr_d, stride_y, y, sign); });
}

// y = A x (incl. low-rank term)
void dev_apply(mgmc_ctx *c, int level, const double *x, double *y) {
  const DevLevel &L = c->lv[level];
  c->launch("apply", level, [&] {
    if (L.d3 && L.full27) apply27_kernel<true, false><<<grid_sites(L.g, c->d.nchains), kBlockSites, 0, c->stream>>>(L.g, L.q, L.c27, x, nullptr, y);
    else if (L.d3) apply27_kernel<false, false><<<grid_sites(L.g, c->d.nchains), kBlockSites, 0, c->stream>>>(L.g, L.q, L.c27, x, nullptr, y);
    else if (L.r2) apply25_kernel<false><<<grid_sites(L.g, c->d.nchains), kBlockSites, 0, c->stream>>>(L.g, L.d_st, x, nullptr, y, RowRange{1, L.g.ny - 1});
    else if (L.vc && L.vc_full) apply9v_kernel<true, false><<<grid_sites(L.g, c->d.nchains), kBlockSites, 0, c->stream>>>(L.g, L.dvc, x, nullptr, y);
    else if (L.vc) apply9v_kernel<false, false><<<grid_sites(L.g, c->d.nchains), kBlockSites, 0, c->stream>>>(L.g, L.dvc, x, nullptr, y);
    else if (L.nine) apply_kernel<true, false><<<grid_sites(L.g, c->d.nchains), kBlockSites, 0, c->stream>>>(L.g, L.coef, x, nullptr, y);
    else apply_kernel<false, false><<<grid_sites(L.g, c->d.nchains), kBlockSites, 0, c->stream>>>(L.g, L.coef, x, nullptr, y);
  });
  if (c->d.m_lowrank > 0 && L.lr_wide) {
    dev_lowrank_wide_apply(c, level, x, L.B.rows, L.g.stride, y, c->d_sigma_inv, 1.0, c->d.nchains);
  } else if (c->d.m_lowrank > 0)
    c->launch("lowrank_apply", level, [&] {
      lowrank_apply_kernel<<<c->d.nchains, 256, c->d.m_lowrank * sizeof(double), c->stream>>>(L.B.cols, L.B.rows, c->d_sigma_inv, L.g.stride, x, y);
    });
}

// ---- fused tile kernel dispatch ----
// tile height: 32 rows on the big (bandwidth / issue bound) levels; the small levels are latency bound,
// there short tiles give every warp at most one row per colour pass and spread over more SMs
inline int fused_tile_rows(int ny, int nc, bool strips, bool restrict_, bool merged, int nchains) {
  if (strips) return ny > 1024 ? 32 : (ny > 256 ? 16 : 8);  // strips must start on tile boundaries: powers of two only
  static const char *ov = std::getenv("MGMC_TILE_ROWS");    // perf experiments: "rb_big,4c_big,4c_mid,small,rb_big_prolong,rb_big_merged"
  static int t[8] = {36, 46, 24, 8, 40, 40, 36, 38};  // (... ,rb_big_merged,4c_1024,4c_big_restrict)
  static bool parsed = false;
  if (!parsed) {
    parsed = true;
    if (ov) std::sscanf(ov, "%d,%d,%d,%d,%d,%d,%d,%d", &t[0], &t[1], &t[2], &t[3], &t[4], &t[5], &t[6], &t[7]);
  }
  // (merged level-0 launch, restrict_ and prolongation: 40 rows + the 13 halo rows of its 5 live passes + residual is
  //  the tallest tile that leaves two CTAs per SM; also on mid-size lattices -- the 16-row tiles of their two-sweep
  //  launches would nearly double the rows a merged launch stages)
  if (ny > 256 && nc == 2 && merged) return t[5];
  // Many chains per launch (ensembles: BASELINE config 5): every level has thousands of tiles, nothing is latency bound --
  // tall tiles everywhere (the 8 / 16 / 24-row tiles of a single chain's small levels stage twice the rows they own),
  // the rows of the level dealt out evenly
  static const bool tall_off = std::getenv("MGMC_NO_TALL_ENSEMBLE_TILES") != nullptr;
  if (nchains >= 8 && ny >= 32 && !tall_off) {
    const int cap = (nc == 2) ? (restrict_ ? t[0] : t[4]) : t[1];
    const int parts = (ny - 1 + cap - 1) / cap;
    return std::min(cap, ((ny - 1 + parts - 1) / parts + 1) / 2 * 2);
  }
  // The rows of a colour pass are dealt out to 16 warps, so the tile heights are chosen to make the passes come out
  // at whole rounds (plan_stages: a red-black launch with restriction updates TY + 7 / 5 / 3 rows, without TY + 4 / 2 / 0;
  // a 4-colour launch every other row of TY + 7 ... TY + 3).
  if (ny > 2048 && nc == 2) return restrict_ ? t[0] : t[4];
  // (a 2048 x 2048 4-colour level with restriction behind stages TY + 15 rows: 38 is the tallest tile that fits)
  if (ny > 1024) return nc == 2 ? 32 : (restrict_ ? t[7] : t[1]);
  // (a 1024 x 1024 4-colour level in 36-row tiles is 290 tiles: one wave of 2 CTAs per SM instead of 1.45 with 24 rows)
  if (ny > 512) return nc == 2 ? 16 : t[6];
  if (ny > 256) return nc == 2 ? 16 : t[2];
  return t[3];
}
#ifndef MGMC_FUSED_SMEM_KB
#define MGMC_FUSED_SMEM_KB 112  // 2 CTAs per SM: 2 x (112 + 1 KB reserved) <= 227 KB  (experiment: 224 with one 1024-thread CTA per SM)
#endif
constexpr int kFusedSmemMax = MGMC_FUSED_SMEM_KB * 1024;
constexpr int kTailTileRowsMax = 40;
constexpr int kTailSmemMax = 200 * 1024;  // the persistent kernel runs one CTA per SM

// Persistent kernel of the small levels (tail.cuh): every tile job of a chain must be resident at once (the low-rank
// exchange between the tiles may wait for any of them), so the tiles are made as tall as that needs (dev_fused).
// Conservative tile count of one chain for the planning: tiles at least 96 columns wide, as tall as allowed.
inline int tail_tiles_bound(int nx, int ny, int nc) {
  const int ty = std::max(std::min(fused_tile_rows(ny, nc, false, true), fused_tile_rows(ny, nc, false, false)), 8);
  int best = ((nx + 95) / 96) * ((ny - 1 + ty - 1) / ty);
  for (int t = ty; t <= kTailTileRowsMax; t += 2) best = std::min(best, ((nx + 95) / 96) * ((ny - 1 + t - 1) / t));
  return best;
}

// coop: tiles of this launch wait for packets of ANY other tile of their chain (interacting measurements): launched
// cooperatively, so that the runtime guarantees (and checks) that the whole grid is resident at once
template <int NC, bool G, bool PR, bool RS, bool LR, bool NZ = false>
void launch_fused_t(mgmc_ctx *c, const FusedP &P, dim3 grid, size_t smem, bool coop = false) {
  // (the attribute is per device and function: tracked per context, not per process)
  const void *fn = (const void *)fused_smooth_kernel<NC, G, PR, RS, LR, NZ>;
  if (c->func_attr_done.insert(fn).second) {
    CUDA_CHECK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, kFusedSmemMax));
    // (with the noise branch in use, always the largest shared-memory carve-out: an SM that has to change its L1 / shared
    //  split for a new CTA first drains -- a small-level launch would wait for a co-resident noise_gen_kernel CTA to finish.
    //  Not otherwise: the prolongation gathers reuse the coarse iterate through L1, and the smaller L1 costs the launches
    //  that do not fill the shared memory anyway 4-12 %)
    static const bool nza_env = std::getenv("MGMC_NOISE_AHEAD") != nullptr;
    if (nza_env) CUDA_CHECK(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  }
  if (coop) {
    cudaLaunchConfig_t cfg;
    std::memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = dim3(kFusedThreads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = c->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative;
    at[0].val.cooperative = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    CUDA_CHECK(cudaLaunchKernelEx(&cfg, fused_smooth_kernel<NC, G, PR, RS, LR, NZ>, P));
    return;
  }
  fused_smooth_kernel<NC, G, PR, RS, LR, NZ><<<grid, kFusedThreads, smem, c->stream>>>(P);
}

struct FixSpec {
  int stage;    // the fix-up follows this stage of the launch
  int dir;      // 0: forward sweep, 1: backward sweep
  uint32_t c1;  // Philox word of the sweep (low-rank noise)
  uint32_t soff = 0;  // added to the sample index (merged level-0 launch: sweeps of the next cycle)
};

// start of every cycle and API call: a new epoch invalidates everything the owner tiles published before
void lr_begin_epoch(mgmc_ctx *c) {
  c->lr_slot_next = 0;
  if (c->d_lr_epoch) c->launch("lr_epoch", 0, [&] { bump_kernel<<<1, 1, 0, c->stream>>>(c->d_lr_epoch); });
}

// shared memory the low-rank bookkeeping of a tile needs behind the tile itself
inline size_t lr_tile_smem(int m, int nfix) { return (size_t)(4 + 3 * nfix) * m * sizeof(double) + (size_t)6 * m * sizeof(int); }

// Can the fix-ups of this level run inside the fused launch?  Always when the measurements do not interact;
// otherwise every tile that needs a fix-up waits for ALL owner tiles of its chain, so all tiles of a chain must be
// resident at once: the tiles of such a level are made tall enough for one CTA per SM (dev_fused), and chains are
// launched in groups that fit.  The answer must not depend on nchains or on whether the level runs as a phase of the
// persistent kernel: the in-kernel and the separate fix-up differ in rounding, and chains are compared bit for bit.
bool lr_fusable(mgmc_ctx *c, const LowRankDev &lr, int level) {
  const int m = c->d.m_lowrank;
  if (!c->lr_fuse || m > 256 || lr.bw > 16 || lr.bh > 16) return false;
  if (lr.diag[0] && lr.diag[1]) return true;
  const DevLevel &L = c->lv[level];
  return tail_tiles_bound(L.g.nx, L.g.ny, L.h.st.ncolours) <= c->num_sms;
}

const LowRankDev &get_lowrank(mgmc_ctx *c, int level, double omega);

// ---- row strips: device-side wait for the neighbours' rows, push of the own boundary rows ----
template <class T>
T *peer_ptr(mgmc_ctx *c, int rank, T *mine) {
  return (T *)(c->peer_arena[rank] + ((char *)mine - c->arena));
}

// rows [j0, j1] of a level array as one contiguous segment (full padded rows)
StripSeg row_segment(mgmc_ctx *c, const DevLevel &L, double *arr, int j0, int j1, int dst_rank) {
  StripSeg sg;
  double *p = arr + (long long)j0 * L.g.pitch - GX;
  sg.src = p;
  sg.dst = peer_ptr(c, dst_rank, p);
  sg.n = (long long)(j1 - j0 + 1) * L.g.pitch;
  return sg;
}

// restriction into the first replicated level: every rank sends its rows of f to every other rank,
// zeroes the whole coarse iterate and waits for the rows of the others
void strip_allgather_rhs(mgmc_ctx *c, int level /* fine, distributed */) {
  const StripPlan &sp = c->strip;
  DevLevel &L = c->lv[level];
  DevLevel &C = c->lv[level + 1];
  const int clo = (sp.lo[level] - 1) / 2 + 1, chi = std::min(sp.hi[level] / 2, C.g.ny - 1);  // coarse rows my tiles produce
  int *ctl = c->d_strip_ctl;
  for (int base = 0; base < sp.nranks; base += 8) {
    StripPush P;
    std::memset(&P, 0, sizeof(P));
    for (int r = base; r < std::min(base + 8, sp.nranks); ++r) {
      if (r == sp.rank) continue;
      P.seg[P.nseg++] = row_segment(c, C, C.f, clo, chi, r);
      P.flag[P.nflag++] = peer_ptr(c, r, ctl + 2);
    }
    P.ticket = (unsigned int *)(ctl + 6);
    if (P.nseg) c->launch("strip_allgather", level, [&] { strip_push_kernel<<<16, 256, 0, c->stream>>>(P); });
  }
  (void)L;
  dev_zero(c, level + 1, C.x);
  c->launch("strip_wait_all", level, [&] { strip_wait_kernel<<<1, 32, 0, c->stream>>>(ctl + 2, nullptr, sp.nranks - 1, 0, 0, 0, ctl + 5, ctl + 3); });
}

// Plans the colour passes of one fused launch backwards from what the launch must deliver (fused.cuh "Stage").
// A rectangle is the margin (xl, xh, yl, yh) around the tile inside which the sites of a colour must be exact.
//   end:      every colour on the tile (+ the halo of the fused residual / restriction, + supp(B_k) of the owned
//             measurements for the low-rank part of the residual)
//   pass s:   updates its colour where it is needed; for that the neighbouring colours must be exact one site
//             further out -- only in the directions in which they are neighbours (4-colour ordering: the colour
//             that differs in the column parity is a neighbour in x only, ...), which keeps the y-halo of an
//             8-pass launch at 3 rows instead of 7.  With omega = 1 the colour's own old values are not read:
//             nothing is required of it before the pass, and a pass nobody needs is dead (skipped).
//   fix-up:   x += W d is pointwise (requirements pass through), but d needs x on supp(B_k) of the owned
//             measurements: every colour becomes "sparsely" needed there (inside tile + (0, lr_mx, 0, lr_my)).
// Returns the halo of the input the launch must load.
struct Margin {
  bool on = false;
  int v[4] = {0, 0, 0, 0};
  void join(const Margin &o) {
    if (!o.on) return;
    if (!on) { *this = o; return; }
    for (int k = 0; k < 4; ++k) v[k] = std::max(v[k], o.v[k]);
  }
  Margin grown(int dx, int dy) const {
    Margin m = *this;
    m.v[0] += dx; m.v[1] += dx; m.v[2] += dy; m.v[3] += dy;
    return m;
  }
};
Margin plan_stages(int nc, std::vector<Stage> &st, const std::vector<FixSpec> &fixes, bool use_lr, bool w1, bool restrict_, int lr_mx, int lr_my, int qoi_stage = -1) {
  static const bool noskip = std::getenv("MGMC_NO_DEAD_PASS") != nullptr;  // (experiments: run every pass in full)
  Margin end;
  end.on = true;
  if (restrict_) { end.v[0] = 2; end.v[1] = 1; end.v[2] = 1; end.v[3] = 2; }
  Margin elr;  // where the owned measurements live
  elr.on = true;
  elr.v[1] = lr_mx;
  elr.v[3] = lr_my;
  if (use_lr && restrict_) end.join(elr);
  std::vector<Margin> req(nc, end);
  std::vector<char> sparse(nc, 0);
  for (int s = (int)st.size() - 1; s >= 0; --s) {
    if (use_lr)
      for (const FixSpec &fx : fixes)
        if (fx.stage == s) std::fill(sparse.begin(), sparse.end(), 1);
    // (merged level-0 launch: the observed sites are recorded after this stage -- every colour is needed there)
    if (s == qoi_stage) std::fill(sparse.begin(), sparse.end(), 1);
    const int cs = st[s].colour;
    Margin r = req[cs];
    const bool sp = sparse[cs] != 0;
    if (noskip && !r.on) r = end;
    st[s].mode = r.on ? STAGE_FULL : (sp ? STAGE_SPARSE : STAGE_SKIP);
    if (r.on && sp) r.join(elr);
    Margin src = r;
    if (!r.on && sp) src = elr;
    st[s].xl = (short)r.v[0]; st[s].xh = (short)r.v[1]; st[s].yl = (short)r.v[2]; st[s].yh = (short)r.v[3];
    if (src.on) {
      for (int c2 = 0; c2 < nc; ++c2) {
        if (c2 == cs) continue;
        const int dx = (nc == 2) ? 1 : ((c2 ^ cs) & 1), dy = (nc == 2) ? 1 : (((c2 ^ cs) >> 1) & 1);
        req[c2].join(src.grown(dx, dy));
      }
      if (w1) {
        req[cs] = Margin();
        sparse[cs] = 0;
      } else {
        req[cs].join(src);
      }
    }
  }
  Margin in;
  for (int c2 = 0; c2 < nc; ++c2) {
    in.join(req[c2]);
    if (sparse[c2]) in.join(elr);
  }
  return in;
}

void dev_fused(mgmc_ctx *c, int level, const std::vector<Stage> &stages, const std::vector<FixSpec> &fixes, bool use_lr, bool gibbs, double omega, bool prolong,
               double alpha, bool restrict_) {
  DevLevel &L = c->lv[level];
  const int nc = L.h.st.ncolours;
  const int S = (int)stages.size();
  if (S > kMaxStages || (S > 8 && c->tail_rec)) fail(MGMC_ERR_INVALID, "internal: too many stages in one fused launch");
  if (S == 0 && !prolong && !restrict_) return;
  FusedP P;
  std::memset(&P, 0, sizeof(P));
  P.g = L.g;
  P.a = L.coef;
  P.x_in = L.x;
  P.x_out = L.x_alt;
  P.f = L.f;
  if (prolong || restrict_) {
    DevLevel &C = c->lv[level + 1];
    P.gc = C.g;
    P.xc_in = C.x;
    P.alpha = alpha;
    P.fc_out = C.f;
    P.xc_zero = C.x;
  }
  if (prolong && restrict_) P.xc_zero = nullptr;  // (merged level-0 launch: see FusedP::x_in_zero)
  if (c->next_x_zero) {
    if (prolong || level == 0) fail(MGMC_ERR_INVALID, "internal: zero-iterate launch out of place");
    P.x_in_zero = 1;
    c->next_x_zero = false;
  }
  P.nstages = S;
  P.winv = omega / L.coef.c;
  P.omega_is_one = (omega == 1.0) ? 1 : 0;
  P.noise_scale = std::sqrt(L.coef.c * (2. - omega) / omega);  // sor_sampler.cc:24-27
  P.wn = P.noise_scale * P.winv;
  {
    const double *src = &L.coef.c;
    double *dst = &P.aw.c;
    for (int k = 0; k < 9; ++k) dst[k] = src[k] * P.winv;
  }
  P.nz = noise_params(c, level, 0);
  auto up4 = [](int v) { return (v + 3) / 4 * 4; };
  if (use_lr) {
    // the owner tile of a measurement must hold all of supp(B_k) exactly at every fix-up: it reaches this far
    // beyond its lower left corner
    const LowRankDev &lr = get_lowrank(c, level, omega);
    P.lr_mx = lr.bw - 1;
    P.lr_my = lr.bh - 1;
  }
  std::vector<Stage> plan = stages;
  const int nqoi = (c->merge_qoi_stage >= 0) ? c->qoi_nnz : 0;
  const Margin halo = plan_stages(nc, plan, fixes, use_lr, omega == 1.0, restrict_, P.lr_mx, P.lr_my, nqoi > 0 ? c->merge_qoi_stage : -1);
  if (nqoi > 0) {
    if (nqoi > kMaxQoi) fail(MGMC_ERR_INVALID, "internal: too many observed sites for a merged launch");
    P.nqoi = nqoi;
    P.qoi_stage = c->merge_qoi_stage;
    for (int e = 0; e < nqoi; ++e) {
      P.qoi_i[e] = (int)(c->h_qsite[e] % L.g.pitch);
      P.qoi_j[e] = (int)(c->h_qsite[e] / L.g.pitch);
    }
    P.qoi_out = c->d_qpart;
  }
  for (int k = 0; k < S; ++k) P.st[k] = plan[k];
  static const bool nofold = std::getenv("MGMC_NO_RES_FOLD") != nullptr;
  // (not with a low-rank term: tiles next to a measurement could not fold, and which tiles those are depends on the
  //  tiling -- the folded residual differs from the stencil one in the last bit, and the chain must not depend on the
  //  tiling / the strip decomposition)
  P.res_stage = (restrict_ && omega == 1.0 && S > 0 && plan[S - 1].mode == STAGE_FULL && !use_lr && !nofold) ? S - 1 : -1;
  P.HXL = up4(halo.v[0]);
  const int HXR = up4(halo.v[1]);
  P.TX = 128 - P.HXL - HXR;
  P.TY = fused_tile_rows(L.g.ny, nc, c->strip.on() && !c->tail_rec, restrict_, prolong && restrict_, c->d.nchains);
  // persistent kernel of the small levels / interacting measurements: all tiles of a chain must be resident at once
  // (one CTA per SM)
  bool lr_coupled = false;
  if (use_lr) {
    const LowRankDev &lr = get_lowrank(c, level, omega);
    lr_coupled = !(lr.diag[0] && lr.diag[1]);
  }
  if (c->tail_rec) {
    // persistent kernel: the phases are latency bound -- as many (low) tiles as there are SMs for the chains of a wave
    const int target = std::max(1, c->num_sms / c->d.nchains);
    P.TY = 2;
    while (((L.g.nx + P.TX - 1) / P.TX) * ((L.g.ny - 1 + P.TY - 1) / P.TY) > target && P.TY < kTailTileRowsMax) P.TY += 2;
  }
  if (c->tail_rec || lr_coupled)
    while (((L.g.nx + P.TX - 1) / P.TX) * ((L.g.ny - 1 + P.TY - 1) / P.TY) > c->num_sms && P.TY < kTailTileRowsMax) P.TY += 2;
  P.hl = halo.v[2];
  const int hh = halo.v[3];
  P.RY = P.TY + P.hl + hh;
  size_t smem = (size_t)2 * P.RY * 128 * sizeof(double) + (use_lr ? lr_tile_smem(c->d.m_lowrank, (int)fixes.size()) : 0);
  while (smem > (size_t)kFusedSmemMax && P.TY > 8) {  // (many measurements: a lower tile makes room for their bookkeeping)
    P.TY -= 8;
    P.RY = P.TY + P.hl + hh;
    smem = (size_t)2 * P.RY * 128 * sizeof(double) + (use_lr ? lr_tile_smem(c->d.m_lowrank, (int)fixes.size()) : 0);
  }
  if (c->strip.on() && !c->tail_rec && P.TY != fused_tile_rows(L.g.ny, nc, true, restrict_)) fail(MGMC_ERR_UNSUPPORTED, "row strips: too many measurements for the tile geometry");
  if (smem > (size_t)kFusedSmemMax) fail(MGMC_ERR_INVALID, "internal: fused tile does not fit in shared memory");
  // ---- normals generated ahead of the launch (noise_ahead.cuh) ----
  bool nzg = false;
  if (gibbs && nc == 4 && S <= 8 && !c->tail_rec && (c->nza_dry || (c->nza_on && c->nza_in_cycle)) && std::find(c->nza_levels.begin(), c->nza_levels.end(), level) != c->nza_levels.end()) {
    const bool fits = ((P.RY + 1) / 2 + kFusedWarps - 1) / kFusedWarps <= kNzgRows;  // rows of a warp per pass held in registers
    if (c->nza_dry) {
      // planning run: record the launch (one plane of normals per full colour pass) instead of emitting it
      mgmc_ctx::NzaSlot sl;
      sl.level = level;
      if (fits) {
        const size_t rows = (size_t)L.g.ny / 2 + 1;
        for (int k = 0; k < S; ++k) {
          if (plan[k].mode != STAGE_FULL) continue;
          NzJob jb;
          jb.buf = c->dalloc<double2>(rows * (size_t)(L.g.pitch / 4));
          jb.colour = plan[k].colour;
          jb.c1 = plan[k].c1;
          sl.jobs.push_back(jb);
          sl.stage.push_back(k);
        }
      }
      c->nza_slots.push_back(sl);
      return;
    }
    int idx = c->nza_cursor[level]++, at = -1;
    for (size_t q = 0; q < c->nza_slots.size(); ++q)
      if (c->nza_slots[q].level == level && idx-- == 0) at = (int)q;
    if (at < 0) fail(MGMC_ERR_INVALID, "internal: launch without a planned noise slot");
    const mgmc_ctx::NzaSlot &sl = c->nza_slots[at];
    if (!sl.jobs.empty() && c->nza_forked) {
      size_t nfull = 0;
      for (int k = 0; k < S; ++k) nfull += plan[k].mode == STAGE_FULL;
      if (nfull != sl.jobs.size()) fail(MGMC_ERR_INVALID, "internal: noise slot does not match the launch");
      for (size_t q = 0; q < sl.jobs.size(); ++q) {
        const Stage &st = plan[sl.stage[q]];
        if (st.mode != STAGE_FULL || st.colour != sl.jobs[q].colour || st.c1 != sl.jobs[q].c1 || st.soff != 0u) fail(MGMC_ERR_INVALID, "internal: noise slot does not match the launch");
        P.nzg[sl.stage[q]] = sl.jobs[q].buf;
      }
      P.nzg_gp = L.g.pitch / 4;
      nzg = true;
      if (!c->nza_joined) {
        CUDA_CHECK(cudaStreamWaitEvent(c->stream, c->ev_gen, 0));  // the planes of this cycle are complete
        c->nza_joined = true;
      }
    }
  }
  P.tiles_x = (L.g.nx + P.TX - 1) / P.TX;
  int tiles_y = (L.g.ny - 1 + P.TY - 1) / P.TY;
  const bool strip_level = c->strip.on() && c->strip_connected && level < c->strip.ndist;
  if (strip_level) {
    // only this rank's tile rows; the halo rows come from the neighbours
    P.by0 = (c->strip.lo[level] - 1) / P.TY;
    tiles_y = (c->strip.hi[level] - c->strip.lo[level] + 1 + P.TY - 1) / P.TY;
    const int need = std::max(P.hl, hh) + 1;
    int lr_reach = 0;
    if (use_lr) {
      // owner tiles near the strip boundary publish into the neighbours' buffers as well: every measurement whose
      // fix-up reaches a neighbour's tile regions must be owned by one of the edge tile rows
      const LowRankDev &lr = get_lowrank(c, level, omega);
      if (!(lr.diag[0] && lr.diag[1])) fail(MGMC_ERR_UNSUPPORTED, "row strips: measurements interact on a distributed level (raise MGMC_STRIP_MIN_ROWS to replicate it)");
      lr_reach = need + lr.wreach + lr.bh + 1;
    } else if (c->d.m_lowrank > 0 && (S > 0 || restrict_)) {
      fail(MGMC_ERR_UNSUPPORTED, "row strips need the in-kernel low-rank fix-up");
    }
    if (need > c->strip.halo[level]) fail(MGMC_ERR_INVALID, "internal: strip halo too small");
    const StripPlan &sp = c->strip;
    const bool has_dn = sp.rank > 0, has_up = sp.rank + 1 < sp.nranks;
    StripK &K = P.sk;
    K.on = 1;
    K.own_lo = sp.lo[level];
    K.own_hi = sp.hi[level];
    K.tiles_y = tiles_y;
    K.halo = sp.halo[level];
    int E = std::max(K.halo, lr_reach);
    if (use_lr) {
      if (has_dn) K.lr_peer_dn = (long long)(c->peer_arena[sp.rank - 1] - c->arena);
      if (has_up) K.lr_peer_up = (long long)(c->peer_arena[sp.rank + 1] - c->arena);
    }
    // the output buffer of this launch is L.x_alt if the launch writes x (it is swapped in below)
    double *xout = (S > 0 || prolong) ? L.x_alt : L.x;
    if (has_dn) K.peer_x_dn = peer_ptr(c, sp.rank - 1, xout);
    if (has_up) K.peer_x_up = peer_ptr(c, sp.rank + 1, xout);
    if (restrict_ && level + 1 < sp.ndist) {
      DevLevel &C = c->lv[level + 1];
      K.clo = sp.lo[level + 1];
      K.chi = sp.hi[level + 1];
      K.chalo = sp.halo[level + 1];
      E = std::max(E, 2 * K.chalo + 2);
      if (has_dn) K.peer_fc_dn = peer_ptr(c, sp.rank - 1, C.f);
      if (has_up) K.peer_fc_up = peer_ptr(c, sp.rank + 1, C.f);
      // the neighbours zero their rows of x_{l+1} themselves: zero my copies of them.  Safe before this launch
      // (nobody reads x_{l+1} any more) and ordered before my flags go up, i.e. before a neighbour can mirror
      // rows of x_{l+1} again.
      const size_t rowb = (size_t)C.g.pitch * sizeof(double);
      if (has_dn) CUDA_CHECK(cudaMemsetAsync(C.x + (long long)(K.clo - K.chalo) * C.g.pitch - GX, 0, rowb * K.chalo, c->stream));
      if (has_up) CUDA_CHECK(cudaMemsetAsync(C.x + (long long)(K.chi + 1) * C.g.pitch - GX, 0, rowb * K.chalo, c->stream));
    }
    int *ctl = c->d_strip_ctl;
    if (has_dn) {
      K.peer_flag_dn = peer_ptr(c, sp.rank - 1, ctl + 1);  // I am the neighbour above rank - 1
      K.flag_from_dn = ctl + 0;
    }
    if (has_up) {
      K.peer_flag_up = peer_ptr(c, sp.rank + 1, ctl + 0);
      K.flag_from_up = ctl + 1;
    }
    K.cycle_no = ctl + 9;
    K.per_cycle = c->strip_per_cycle;
    K.index = c->strip_index++;
    K.edge_rows = std::min((E + P.TY - 1) / P.TY, tiles_y);
    K.ticket_dn = (unsigned int *)(ctl + 7);
    K.ticket_up = (unsigned int *)(ctl + 8);
    K.err = ctl + 3;
  }
  P.nchains = c->d.nchains;
  P.err = c->d_err;
  if (use_lr) {
    const LowRankDev &lr = get_lowrank(c, level, omega);
    P.lr = lr.tile;
    P.nfix = (int)fixes.size();
    if (P.nfix > kMaxFix) fail(MGMC_ERR_INVALID, "internal: too many fix-ups in one fused launch");
    for (int q = 0; q < P.nfix; ++q) {
      P.fix_stage[q] = fixes[q].stage;
      P.fix_dir[q] = fixes[q].dir;
      P.fix_c1[q] = fixes[q].c1;
      P.fix_soff[q] = fixes[q].soff;
    }
    P.lr_u_from_fix = (restrict_ && P.nfix > 0 && fixes.back().stage == S - 1) ? 1 : 0;
    P.lr_slot = c->lr_slot_next;
    c->lr_slot_next += P.nfix + 1;
    if (c->d_lr_flag_pool) {
      // flags are a function of the launch geometry and of what the launch looks at (fix-up directions, residual)
      char key[160];
      std::snprintf(key, sizeof(key), "%d:%.17g:%d,%d,%d,%d,%d,%d,%d,%d:%d,%d,%d:%d,%d,%d", level, omega, P.HXL, P.TX, P.TY, P.hl, P.RY, P.tiles_x, tiles_y, P.by0, P.nfix,
                    P.nfix > 0 ? P.fix_dir[0] : -1, P.nfix > 1 ? P.fix_dir[1] : -1, (int)restrict_, P.sk.on, (int)prolong);
      const size_t ntile = (size_t)P.tiles_x * tiles_y * c->d.nchains;  // one flag per tile and chain
      auto it = c->lr_flag_slots.find(key);
      if (it == c->lr_flag_slots.end() && c->lr_flag_pool_used + ntile <= c->lr_flag_pool_size) {
        it = c->lr_flag_slots.emplace(key, c->lr_flag_pool_used).first;
        c->lr_flag_pool_used += (ntile + 15) / 16 * 16;
      }
      if (it != c->lr_flag_slots.end()) P.lr_flags = c->d_lr_flag_pool + it->second;  // (pool exhausted: every tile runs the tests)
    }
    if (c->lr_slot_next > mgmc_ctx::kLrSlots) fail(MGMC_ERR_UNSUPPORTED, "too many low-rank fix-ups in one cycle (W-cycle too deep)");
  }
  dim3 grid(P.tiles_x * tiles_y, 1, c->d.nchains);
  // algorithmic bytes of this launch: 24 B per site and sweep, 18 B prolongate_add,
  // 18 + 2 B residual + restrict + coarse zeroing -- fixed by the model, not by what the kernel moves
  const bool strip_own = c->strip.on() && c->strip_connected && level < c->strip.ndist;
  const double nsites = strip_own ? (double)(c->strip.hi[level] - c->strip.lo[level] + 1) * (L.g.nx - 1) : (double)L.h.ndof();
  const double alg_bytes = nsites * c->d.nchains * (24.0 * S / nc + (prolong ? 18.0 : 0.0) + (restrict_ ? 20.0 : 0.0));
  std::string name = std::string(gibbs ? "gibbs" : "sor") + (nc == 2 ? "_rb" : "_4c") + std::to_string(S) + (prolong ? "+prolong" : "") + (restrict_ ? "+restrict" : "");
#ifdef MGMC_TILE_TIMING
  // debug build: dump per-CTA phase time stamps of the first level-0 launch of every kernel flavour
  static std::map<std::string, int> dumped;
  const char *tfile = std::getenv("MGMC_TIMING_FILE");
  long long *d_timing = nullptr;
  const size_t ncta = (size_t)grid.x * grid.z;
  static const char *tlev = std::getenv("MGMC_TIMING_LEVEL");
  cudaStreamCaptureStatus cap_status = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(c->stream, &cap_status);  // (the dump allocates and synchronises: eager launches only)
  if (cap_status == cudaStreamCaptureStatusNone && tfile && level == (tlev ? std::atoi(tlev) : 0) && dumped[name]++ == 3) {
    CUDA_CHECK(cudaMalloc(&d_timing, ncta * 16 * sizeof(long long)));
    CUDA_CHECK(cudaMemset(d_timing, 0, ncta * 16 * sizeof(long long)));
    P.timing = d_timing;
  }
#endif
  if (c->tail_rec) {
    if (strip_level) fail(MGMC_ERR_INVALID, "internal: distributed level inside the persistent tail kernel");
    {
      // behind the tile: noise generated ahead of the passes (fused.cuh; one item per (row, live pass), 32 pairs of
      // normals each) and the pass descriptors
      int items = 0;
      const int step = (nc == 4) ? 2 : 1;
      if (gibbs)
        for (int k = 0; k < S; ++k)
          if (plan[k].mode == STAGE_FULL) items += (P.TY - 1 + plan[k].yl + plan[k].yh) / step + 1;
      if (items > kFusedThreads) items = 0;  // (cannot happen with the tile heights of the persistent kernel)
      const size_t off = (smem + 15) / 16 * 16;
      P.nz_off = (int)(off / sizeof(double));
      P.nz_cap = items;
      smem = off + (size_t)items * 32 * sizeof(double2) + (32 + 512 + 2 * kFusedThreads) * sizeof(int);
    }
    TailPhase ph;
    std::memset(&ph, 0, sizeof(ph));
    ph.kind = TAIL_FUSED;
    ph.nc = nc;
    ph.ntiles = P.tiles_x * tiles_y;
    if (ph.ntiles > c->num_sms) fail(MGMC_ERR_INVALID, "internal: level too large for the persistent tail kernel");
    ph.P = P;
    ph.P.rt_prolong = prolong ? 1 : 0;
    ph.P.rt_restrict = restrict_ ? 1 : 0;
    c->tail_ph.push_back(ph);
    c->tail_smem = std::max(c->tail_smem, smem);
    c->tail_bytes += alg_bytes;
    c->tail_lowrank = c->tail_lowrank || use_lr;
    if (S > 0 || prolong) std::swap(L.x, L.x_alt);
    return;
  }
  // interacting measurements: the chains are launched in groups whose tiles are all resident at once
  int chain_group = c->d.nchains;
  if (lr_coupled) {
    if ((int)grid.x > c->num_sms) fail(MGMC_ERR_INVALID, "internal: level with interacting measurements does not fit on the chip");
    chain_group = std::max(1, c->num_sms / (int)grid.x);
  }
  const bool coop = lr_coupled;
  c->launch(name.c_str(), level, [&] {
   for (int c0 = 0; c0 < c->d.nchains; c0 += chain_group) {
    P.chain_off = c0;
    grid.z = std::min(chain_group, c->d.nchains - c0);
#define FUSED_CASE(NC_, G_, PR_, RS_)                                             \
  if (nc == NC_ && gibbs == G_ && prolong == PR_ && restrict_ == RS_) {            \
    if (G_ && nzg && NC_ == 4) { /* (the small levels are 4-colour levels: Galerkin 9-point operators) */ \
      if (use_lr) launch_fused_t<NC_, G_, PR_, RS_, true, G_ && NC_ == 4>(c, P, grid, smem, coop); \
      else launch_fused_t<NC_, G_, PR_, RS_, false, G_ && NC_ == 4>(c, P, grid, smem);        \
    } else if (use_lr) launch_fused_t<NC_, G_, PR_, RS_, true>(c, P, grid, smem, coop); \
    else launch_fused_t<NC_, G_, PR_, RS_, false>(c, P, grid, smem);              \
  }
    FUSED_CASE(2, false, false, false) FUSED_CASE(2, false, false, true) FUSED_CASE(2, false, true, false) FUSED_CASE(2, false, true, true)
    FUSED_CASE(2, true, false, false) FUSED_CASE(2, true, false, true) FUSED_CASE(2, true, true, false) FUSED_CASE(2, true, true, true)
    FUSED_CASE(4, false, false, false) FUSED_CASE(4, false, false, true) FUSED_CASE(4, false, true, false) FUSED_CASE(4, false, true, true)
    FUSED_CASE(4, true, false, false) FUSED_CASE(4, true, false, true) FUSED_CASE(4, true, true, false) FUSED_CASE(4, true, true, true)
#undef FUSED_CASE
   }
  }, alg_bytes);
#ifdef MGMC_TILE_TIMING
  if (d_timing) {
    c->sync();
    std::vector<long long> h(ncta * 16);
    CUDA_CHECK(cudaMemcpy(h.data(), d_timing, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    cudaFree(d_timing);
    FILE *fp = std::fopen((std::string(tfile) + "." + name + ".txt").c_str(), "w");
    for (size_t k = 0; k < ncta; ++k) {
      for (int q = 0; q < 16; ++q) std::fprintf(fp, "%lld ", h[k * 16 + q]);
      std::fprintf(fp, "\n");
    }
    std::fclose(fp);
  }
#endif
  if (S > 0 || prolong) std::swap(L.x, L.x_alt);
  if (strip_level && restrict_ && level + 1 == c->strip.ndist) strip_allgather_rhs(c, level);
}

void dev_lowrank_fix(mgmc_ctx *c, int level, bool fwd, bool gibbs, double omega, uint32_t c1) {
  DevLevel &L = c->lv[level];
  const LowRankDev &lr = get_lowrank(c, level, omega);
  const LowRankFix &F = lr.fix[fwd ? 0 : 1];
  NoiseP nz = noise_params(c, level, c1);
  if (lr.wide) {
    const int m = c->d.m_lowrank, nch = c->d.nchains;
    const SparseRows &R = lr.W[fwd ? 0 : 1].rows;
    dev_lowrank_wide_partial(c, level, L.x, nch);
    c->launch("lowrank_d", level, [&] {
      if (gibbs) lowrank_d_kernel<2><<<nch, 256, 2 * m * sizeof(double), c->stream>>>(m, kLrWideBlocks, c->d_lr_partial, nullptr, F.Mneg, F.Ms, F.sigma_inv_sqrt, nz, c->d_lr_d);
      else lowrank_d_kernel<1><<<nch, 256, 2 * m * sizeof(double), c->stream>>>(m, kLrWideBlocks, c->d_lr_partial, nullptr, F.Mneg, F.Ms, F.sigma_inv_sqrt, nz, c->d_lr_d);
    });
    c->launch("lowrank_scatter", level, [&] { lowrank_scatter_kernel<<<dim3((R.nu + 255) / 256, nch), 256, 0, c->stream>>>(R, m, c->d_lr_d, L.g.stride, L.x, 1.0); });
    return;
  }
  c->launch("lowrank_fix", level, [&] {
    if (gibbs) lowrank_fix_kernel<true><<<c->d.nchains, 256, lr.smem, c->stream>>>(F, L.g.stride, L.x, nz);
    else lowrank_fix_kernel<false><<<c->d.nchains, 256, lr.smem, c->stream>>>(F, L.g.stride, L.x, nz);
  });
}

struct SweepSpec {
  bool fwd;
  bool fix_after;
};

// Sweep lists with the reference's semantics, including the nsmooth^2 quirk of the deterministic
// SORSmoother (sor_smoother.cc:43,64) which the samplers do not have (sor_sampler.cc:28,39)
std::vector<SweepSpec> sweep_list(int kind, int direction, int nsmooth, bool gibbs) {
  std::vector<SweepSpec> out;
  if (kind == MGMC_SMOOTHER_SOR) {
    const bool fwd = (direction == MGMC_FORWARD);
    for (int k = 0; k < nsmooth; ++k) {
      if (gibbs) out.push_back({fwd, true});
      else
        for (int q = 0; q < nsmooth; ++q) out.push_back({fwd, q == nsmooth - 1});
    }
  } else {
    for (int k = 0; k < nsmooth; ++k) {
      out.push_back({true, true});
      out.push_back({false, true});
    }
  }
  return out;
}

// A smoothing step of a level: [prolongate_add] sweeps... [residual + restrict].  Consecutive sweeps are
// fused into launches of up to 2 sweeps (4 / 8 colour passes).  The Woodbury fix-up of the low-rank
// term after a sweep is a grid-wide dependency: normally it is resolved inside the launch by the owner
// tiles of the measurements (fused.cuh 4.2); where that scheme does not apply, every sweep becomes its own launch followed
// by a fix-up kernel, and the low-rank part of the residual is a separate kernel.
void emit_smoothing_r2(mgmc_ctx *c, int level, const std::vector<SweepSpec> &sweeps, bool gibbs, double omega, bool prolong, double alpha, bool restrict_);

void emit_smoothing(mgmc_ctx *c, int level, const std::vector<SweepSpec> &sweeps, bool gibbs, double omega, bool prolong, double alpha, bool restrict_) {
  DevLevel &L = c->lv[level];
  if (L.generic()) return emit_smoothing_r2(c, level, sweeps, gibbs, omega, prolong, alpha, restrict_);
  const int nc = L.h.st.ncolours;
  const bool lowrank = c->d.m_lowrank > 0;
  // colour passes per launch: 2 sweeps (tile + halo of x and f stay below ~100 KB, 2 CTAs / SM, and a launch carries at
  // most two low-rank fix-ups); red-black levels without a low-rank term take 4 sweeps -- V(2,2): with omega = 1 only
  // 5 of the 8 passes are live (plan_stages), one launch instead of two per smoothing step
  // Small 4-colour levels (latency bound: a launch costs its in-order latency, not its passes) take up to 16 passes --
  // the 4 sweeps of a V(2,2) smoothing step, at most 4 fix-ups -- in one launch
  static const bool no16 = std::getenv("MGMC_NO_LONG_LAUNCHES") != nullptr;
  const bool small_level = nc == 4 && !c->tail_rec && !no16 && c->d.nchains < 8 && (long long)L.g.nx * L.g.ny <= 512ll * 512ll && !(c->strip.on() && level < c->strip.ndist);
  const int max_stages = (nc == 2 && (lowrank || omega != 1.0)) ? 4 : (small_level ? kMaxStages : 8);
  bool fusedlr = false;
  if (lowrank) {
    fusedlr = lr_fusable(c, get_lowrank(c, level, omega), level);
  }
  std::vector<Stage> cur;
  std::vector<FixSpec> fixes;
  bool pending_prolong = prolong;
  auto flush = [&](bool with_restrict) {
    const bool use_lr = fusedlr && (!fixes.empty() || with_restrict);
    dev_fused(c, level, cur, fixes, use_lr, gibbs, omega, pending_prolong, alpha, with_restrict);
    pending_prolong = false;
    cur.clear();
    fixes.clear();
  };
  for (const SweepSpec &sw : sweeps) {
    const uint32_t c1 = next_c1(c, level, gibbs);
    if (!cur.empty() && ((int)cur.size() + nc > max_stages || (lowrank && !fusedlr))) flush(false);
    for (int cc = 0; cc < nc; ++cc) cur.push_back(Stage{sw.fwd ? cc : nc - 1 - cc, c1});
    if (lowrank) {
      if (fusedlr) {
        if (sw.fix_after) fixes.push_back(FixSpec{(int)cur.size() - 1, sw.fwd ? 0 : 1, c1});
      } else {
        flush(false);
        if (sw.fix_after) dev_lowrank_fix(c, level, sw.fwd, gibbs, omega, c1);
      }
    }
  }
  if (!lowrank || fusedlr) {
    flush(restrict_);
  } else {
    if (pending_prolong) flush(false);
    if (restrict_) {
      flush(true);
      DevLevel &C = c->lv[level + 1];
      if (L.lr_wide) dev_lowrank_wide_apply(c, level, L.x, C.B.rows, C.g.stride, C.f, c->d_sigma_inv, -1.0, c->d.nchains);
      else
        c->launch("lowrank_restrict", level, [&] {
          lowrank_restrict_kernel<<<c->d.nchains, 256, c->d.m_lowrank * sizeof(double), c->stream>>>(L.B.cols, C.B.rows, c->d_sigma_inv, L.g.stride, C.g.stride, L.x, C.f);
        });
    }
  }
}

void dev_restrict_plain(mgmc_ctx *c, int level, const double *r, double *fc);

// Radius-2 levels (squared shifted Laplacian): colour-by-colour launches of the generic kernels, separate
// transfer kernels -- the first correct path for this operator family, not yet tiled / fused.
void emit_smoothing_r2(mgmc_ctx *c, int level, const std::vector<SweepSpec> &sweeps, bool gibbs, double omega, bool prolong, double alpha, bool restrict_) {
  DevLevel &L = c->lv[level];
  const int nch = c->d.nchains;
  const bool lowrank = c->d.m_lowrank > 0;
  // Row strips: this rank sweeps its own rows; the colour launches exchange the `halo` rows next to a neighbour
  // themselves (kernels.cuh StripR2).  Transfers need no exchange: the prolongation is also applied to the mirrored
  // rows (same arithmetic as on the owner), the residual is formed one row beyond the own rows for the restriction.
  const bool strip_level = c->strip.on() && c->strip_connected && level < c->strip.ndist;
  const StripPlan &sp = c->strip;
  const int lo = strip_level ? sp.lo[level] : 1, hi = strip_level ? sp.hi[level] : L.g.ny - 1;
  const int halo = strip_level ? sp.halo[level] : 0;
  const bool has_dn = strip_level && sp.rank > 0, has_up = strip_level && sp.rank + 1 < sp.nranks;
  if (strip_level && lowrank) fail(MGMC_ERR_UNSUPPORTED, "row strips of a radius-2 operator with a low-rank term are not implemented");
  if (strip_level && L.vc) fail(MGMC_ERR_UNSUPPORTED, "row strips of an operator with per-vertex coefficients are not implemented");
  const int ncol = L.h.st.ncolours;  // 9: radius 2; 2 / 4: per-vertex coefficients (varcoef.cuh)
  const double n = L.d3 ? (double)L.h.ndof() * nch : (double)(hi - lo + 1) * (L.g.nx - 1) * nch;
  auto rows_grid = [&](int j0, int j1) { return dim3((L.g.nx - 1 + 63) / 64, (std::max(j1 - j0 + 1, 1) + 3) / 4, nch); };
  int *ctl = c->d_strip_ctl;
  auto strip_sync = [&] {  // wait until the neighbours have finished every distributed launch emitted so far
    if (!strip_level) return;
    const int idx = c->strip_index, per = c->strip_per_cycle;
    c->launch("strip_sync", level, [&] { strip_sync_kernel<<<1, 32, 0, c->stream>>>(has_dn ? ctl + 0 : nullptr, has_up ? ctl + 1 : nullptr, ctl + 9, per, idx, ctl + 3); });
  };
  if (prolong) {
    DevLevel &C = c->lv[level + 1];
    const RowRange rr{std::max(1, lo - (has_dn ? halo : 0)), std::min(L.g.ny - 1, hi + (has_up ? halo : 0))};
    strip_sync();
    c->launch("prolongate_add", level, [&] {
      if (L.d3) prolongate_add27_kernel<<<grid_sites(L.g, nch), kBlockSites, 0, c->stream>>>(L.g, L.q, C.g, C.q, alpha, C.x, L.x);
      else prolongate_add_kernel<<<rows_grid(rr.j0, rr.j1), kBlockSites, 0, c->stream>>>(L.g, C.g, alpha, C.x, L.x, rr);
    }, (L.d3 ? 17.0 : 18.0) * n);
    if (strip_level) {  // counts as a distributed launch: the neighbours' next colour launch waits for it
      c->strip_index++;
      c->launch("strip_raise", level, [&] {
        strip_raise_kernel<<<1, 32, 0, c->stream>>>(has_dn ? peer_ptr(c, sp.rank - 1, ctl + 1) : nullptr, has_up ? peer_ptr(c, sp.rank + 1, ctl + 0) : nullptr);
      });
    }
  }
  // The coarse iterate of a distributed next level is zeroed BEFORE the sweeps: the neighbours' first coarse colour
  // launch pushes into my mirrored rows of it as soon as my last colour launch here has raised their flag.
  const bool zero_first = strip_level && restrict_ && level + 1 < sp.ndist;
  if (zero_first) dev_zero(c, level + 1, c->lv[level + 1].x);
  static const bool noskip = std::getenv("MGMC_NO_DEAD_PASS") != nullptr;
  // Colour passes of one row class in one launch (rowfuse.cuh): 9-colour radius-2 levels, 8-colour 3d levels, 4-colour
  // per-vertex levels on one GPU.  (Row strips keep one launch per colour: the halo exchange is per colour.)
  static const bool no_rowfuse = std::getenv("MGMC_NO_ROWFUSE") != nullptr;
  const bool rowfuse = !no_rowfuse && !strip_level && ((L.r2 && ncol == 9) || (L.d3 && L.full27) || (L.vc && ncol == 4));
  const int per_row = L.r2 ? 3 : 2;  // colours per row class
  RowPasses grp;
  grp.n = 0;
  int grp_class = -1;
  auto flush_rows = [&] {
    if (grp.n == 0) return;
    const RowPasses P = grp;
    const int cls = grp_class;
    const NoiseP nz0 = noise_params(c, level, 0);
    const int per_colour = (L.g.nx - 1 + per_row - 1) / per_row;  // sites of a colour in a row (upper bound)
    const int threads = std::min(256, std::max(32, (per_colour + 31) / 32 * 32));  // (up to 512 threads per row measured: slower)
    const bool pre = per_colour <= threads;  // one site per thread and pass: right-hand sides + noise of all passes up front
    if (L.r2) {
      const int jfirst = (cls == 0) ? 3 : cls;
      const int nrows = (L.g.ny - 1 >= jfirst) ? (L.g.ny - 1 - jfirst) / 3 + 1 : 0;
      const dim3 gridr(std::max(nrows, 1), 1, nch);
      c->launch(gibbs ? "gibbs_9c_rows" : "sor_9c_rows", level, [&] {
        if (gibbs && pre) sweep_rows25_kernel<true, true><<<gridr, threads, 0, c->stream>>>(L.g, L.d_st, L.x, L.f, jfirst, omega, nz0, P);
        else if (gibbs) sweep_rows25_kernel<true, false><<<gridr, threads, 0, c->stream>>>(L.g, L.d_st, L.x, L.f, jfirst, omega, nz0, P);
        else if (pre) sweep_rows25_kernel<false, true><<<gridr, threads, 0, c->stream>>>(L.g, L.d_st, L.x, L.f, jfirst, omega, nz0, P);
        else sweep_rows25_kernel<false, false><<<gridr, threads, 0, c->stream>>>(L.g, L.d_st, L.x, L.f, jfirst, omega, nz0, P);
      }, 24.0 * n / ncol * P.n);
    } else if (L.d3) {
      const int j0 = (cls & 1) ? 1 : 2, k0 = (cls & 2) ? 1 : 2;
      const int nj = (L.q.ny - 1 >= j0) ? (L.q.ny - 1 - j0) / 2 + 1 : 0, nk = (L.q.nz - 1 >= k0) ? (L.q.nz - 1 - k0) / 2 + 1 : 0;
      const dim3 gridr(std::max(nj, 1), std::max(nk, 1), nch);
      c->launch(gibbs ? "gibbs_8c_rows" : "sor_8c_rows", level, [&] {
        if (gibbs && pre) sweep_rows27_kernel<true, true><<<gridr, threads, 0, c->stream>>>(L.g, L.q, L.c27, L.x, L.f, j0, k0, omega, nz0, P);
        else if (gibbs) sweep_rows27_kernel<true, false><<<gridr, threads, 0, c->stream>>>(L.g, L.q, L.c27, L.x, L.f, j0, k0, omega, nz0, P);
        else if (pre) sweep_rows27_kernel<false, true><<<gridr, threads, 0, c->stream>>>(L.g, L.q, L.c27, L.x, L.f, j0, k0, omega, nz0, P);
        else sweep_rows27_kernel<false, false><<<gridr, threads, 0, c->stream>>>(L.g, L.q, L.c27, L.x, L.f, j0, k0, omega, nz0, P);
      }, 24.0 * n / ncol * P.n);
    } else {
      const int j0 = (cls & 1) ? 1 : 2;
      const int nrows = (L.g.ny - 1 >= j0) ? (L.g.ny - 1 - j0) / 2 + 1 : 0;
      const dim3 gridr(std::max(nrows, 1), 1, nch);
      c->launch(gibbs ? "gibbs_4cv_rows" : "sor_4cv_rows", level, [&] {
#define MGMC_ROWS9V(NINE_, GIBBS_)                                                                                                                       \
  do {                                                                                                                                                   \
    if (pre) sweep_rows9v_kernel<NINE_, GIBBS_, true><<<gridr, threads, 0, c->stream>>>(L.g, L.dvc, L.x, L.f, j0, omega, nz0, P);                          \
    else sweep_rows9v_kernel<NINE_, GIBBS_, false><<<gridr, threads, 0, c->stream>>>(L.g, L.dvc, L.x, L.f, j0, omega, nz0, P);                             \
  } while (0)
        if (L.vc_full) {
          if (gibbs) MGMC_ROWS9V(true, true);
          else MGMC_ROWS9V(true, false);
        } else {
          if (gibbs) MGMC_ROWS9V(false, true);
          else MGMC_ROWS9V(false, false);
        }
#undef MGMC_ROWS9V
      }, 24.0 * n / ncol * P.n);
    }
    grp.n = 0;
    grp_class = -1;
  };
  for (size_t si = 0; si < sweeps.size(); ++si) {
    const SweepSpec &sw = sweeps[si];
    const uint32_t c1 = next_c1(c, level, gibbs);
    NoiseP nz = noise_params(c, level, c1);
    for (int cc = 0; cc < ncol; ++cc) {
      const int colour = sw.fwd ? cc : ncol - 1 - cc;
      // omega = 1: an update does not read the site's own value, so the last colour of this sweep is dead if the next
      // sweep starts with the same colour and nothing reads x in between (see plan_stages for the tile kernel)
      if (cc == ncol - 1 && omega == 1.0 && !noskip && si + 1 < sweeps.size() && sweeps[si + 1].fwd != sw.fwd && !(lowrank && sw.fix_after)) continue;
      if (rowfuse) {
        const int cls = colour / per_row;
        if (cls != grp_class || grp.n == kRowPassMax) flush_rows();
        grp_class = cls;
        grp.ci[grp.n] = colour % per_row;
        grp.c1[grp.n] = c1;
        ++grp.n;
        continue;
      }
      if (L.d3) {
        // 3d lattice: red-black (7-point) or 8 colours (27-point), whole lattice, rows = stacked planes (lattice3d.cuh)
        dim3 grid3((L.g.nx / 2 + 1 + 63) / 64, (L.g.ny - 1 + 3) / 4, nch);
        c->launch(gibbs ? (L.full27 ? "gibbs_8c1" : "gibbs_rb1/3d") : (L.full27 ? "sor_8c1" : "sor_rb1/3d"), level, [&] {
          if (L.full27) {
            if (gibbs) sweep_colour27_kernel<true, true><<<grid3, kBlockSites, 0, c->stream>>>(L.g, L.q, L.c27, L.x, L.f, colour, omega, nz);
            else sweep_colour27_kernel<true, false><<<grid3, kBlockSites, 0, c->stream>>>(L.g, L.q, L.c27, L.x, L.f, colour, omega, nz);
          } else {
            if (gibbs) sweep_colour27_kernel<false, true><<<grid3, kBlockSites, 0, c->stream>>>(L.g, L.q, L.c27, L.x, L.f, colour, omega, nz);
            else sweep_colour27_kernel<false, false><<<grid3, kBlockSites, 0, c->stream>>>(L.g, L.q, L.c27, L.x, L.f, colour, omega, nz);
          }
        }, 24.0 * n / ncol);
        continue;
      }
      if (L.vc) {
        // per-vertex coefficients: red-black (every row) or 4 colours (every other row), whole lattice
        const bool four = (ncol == 4);
        const int j0 = four ? ((colour >> 1) ? 1 : 2) : 1, jstep = four ? 2 : 1;
        const int nrows = (L.g.ny - 1 - j0) / jstep + 1;
        dim3 gridv((L.g.nx / 2 + 1 + 63) / 64, std::max((nrows + 3) / 4, 1), nch);
        c->launch(gibbs ? (four ? "gibbs_4c1v" : "gibbs_rb1v") : (four ? "sor_4c1v" : "sor_rb1v"), level, [&] {
          if (L.vc_full) {
            if (gibbs) sweep_colour9v_kernel<true, true><<<gridv, kBlockSites, 0, c->stream>>>(L.g, L.dvc, L.x, L.f, colour, ncol, omega, nz, j0, jstep);
            else sweep_colour9v_kernel<true, false><<<gridv, kBlockSites, 0, c->stream>>>(L.g, L.dvc, L.x, L.f, colour, ncol, omega, nz, j0, jstep);
          } else {
            if (gibbs) sweep_colour9v_kernel<false, true><<<gridv, kBlockSites, 0, c->stream>>>(L.g, L.dvc, L.x, L.f, colour, ncol, omega, nz, j0, jstep);
            else sweep_colour9v_kernel<false, false><<<gridv, kBlockSites, 0, c->stream>>>(L.g, L.dvc, L.x, L.f, colour, ncol, omega, nz, j0, jstep);
          }
        }, 24.0 * n / ncol);
        continue;
      }
      const int cj = colour / 3;
      // first row of this colour at or above lo: rows j = cj (mod 3), j >= 1 (row 0 is the boundary)
      int jfirst = (cj == 0) ? 3 : cj;
      if (jfirst < lo) jfirst += (lo - jfirst + 2) / 3 * 3;
      const int nrows3 = (hi >= jfirst) ? (hi - jfirst) / 3 + 1 : 0;
      dim3 grid((L.g.nx / 3 + 1 + 63) / 64, std::max((nrows3 + 3) / 4, 1), nch);
      StripR2 K;
      std::memset(&K, 0, sizeof(K));
      if (strip_level) {
        K.on = 1;
        K.lo = lo;
        K.hi = hi;
        K.halo = halo;
        if (has_dn) {
          K.peer_dn = peer_ptr(c, sp.rank - 1, L.x);
          K.peer_flag_dn = peer_ptr(c, sp.rank - 1, ctl + 1);
          K.flag_from_dn = ctl + 0;
        }
        if (has_up) {
          K.peer_up = peer_ptr(c, sp.rank + 1, L.x);
          K.peer_flag_up = peer_ptr(c, sp.rank + 1, ctl + 0);
          K.flag_from_up = ctl + 1;
        }
        K.cycle_no = ctl + 9;
        K.per_cycle = c->strip_per_cycle;
        K.index = c->strip_index++;
        K.ticket_dn = (unsigned int *)(ctl + 7);
        K.ticket_up = (unsigned int *)(ctl + 8);
        K.err = ctl + 3;
        unsigned int ndn = 0, nup = 0;
        for (unsigned int by = 0; by < grid.y; ++by) {  // the kernel's own tests
          const int jc0 = jfirst + 12 * (int)by, jc1 = jc0 + 9;
          if (has_dn && jc0 < lo + halo) ++ndn;
          if (has_up && jc1 > hi - halo && jc0 <= hi) ++nup;
        }
        K.n_edge_dn = ndn * grid.x * grid.z;
        K.n_edge_up = nup * grid.x * grid.z;
        if ((has_dn && ndn == 0) || (has_up && nup == 0)) fail(MGMC_ERR_INVALID, "internal: strip too short for the radius-2 halo exchange");
      }
      c->launch(gibbs ? "gibbs_9c1" : "sor_9c1", level, [&] {
        if (gibbs) sweep_colour25_kernel<true><<<grid, kBlockSites, 0, c->stream>>>(L.g, L.d_st, L.x, L.f, colour, omega, nz, jfirst, K);
        else sweep_colour25_kernel<false><<<grid, kBlockSites, 0, c->stream>>>(L.g, L.d_st, L.x, L.f, colour, omega, nz, jfirst, K);
      }, 24.0 * n / 9.0);
    }
    if (lowrank && sw.fix_after) {
      flush_rows();
      dev_lowrank_fix(c, level, sw.fwd, gibbs, omega, c1);
    }
  }
  flush_rows();
  if (restrict_) {
    DevLevel &C = c->lv[level + 1];
    // the restriction to the own coarse rows reads the residual one fine row beyond the own rows
    const RowRange rr{lo, std::min(L.g.ny - 1, hi + (has_up ? 1 : 0))};
    strip_sync();
    c->launch("residual", level, [&] {
      if (L.d3 && L.full27) apply27_kernel<true, true><<<grid_sites(L.g, nch), kBlockSites, 0, c->stream>>>(L.g, L.q, L.c27, L.x, L.f, L.r);
      else if (L.d3) apply27_kernel<false, true><<<grid_sites(L.g, nch), kBlockSites, 0, c->stream>>>(L.g, L.q, L.c27, L.x, L.f, L.r);
      else if (L.vc && L.vc_full) apply9v_kernel<true, true><<<grid_sites(L.g, nch), kBlockSites, 0, c->stream>>>(L.g, L.dvc, L.x, L.f, L.r);
      else if (L.vc) apply9v_kernel<false, true><<<grid_sites(L.g, nch), kBlockSites, 0, c->stream>>>(L.g, L.dvc, L.x, L.f, L.r);
      else apply25_kernel<true><<<rows_grid(rr.j0, rr.j1), kBlockSites, 0, c->stream>>>(L.g, L.d_st, L.x, L.f, L.r, rr);
    }, 16.0 * n);
    if (lowrank && L.lr_wide) dev_lowrank_wide_apply(c, level, L.x, L.B.rows, L.g.stride, L.r, c->d_sigma_inv, -1.0, nch);
    else if (lowrank)
      c->launch("lowrank_residual", level, [&] {
        lowrank_apply_kernel<<<nch, 256, c->d.m_lowrank * sizeof(double), c->stream>>>(L.B.cols, L.B.rows, c->d_sigma_inv_neg, L.g.stride, L.x, L.r);
      });
    if (strip_level) {
      const RowRange cr{(lo - 1) / 2 + 1, std::min(hi / 2, C.g.ny - 1)};
      c->launch("restrict", level, [&] {
        restrict_kernel<<<dim3((C.g.nx - 1 + 63) / 64, (cr.j1 - cr.j0 + 1 + 3) / 4, nch), kBlockSites, 0, c->stream>>>(L.g, C.g, L.r, C.f, cr);
      });
      if (level + 1 == sp.ndist) {
        strip_allgather_rhs(c, level);  // (zeroes the coarse iterate as well)
      } else {
        // the neighbour below forms its coarse residual one row beyond its strip: it needs my first row of f_c
        // (counts as a distributed launch: both flags go up)
        c->strip_index++;
        const long long off = (long long)cr.j0 * C.g.pitch - GX;
        c->launch("strip_row_push", level, [&] {
          strip_row_push_kernel<<<1, 256, 0, c->stream>>>(C.f + off, has_dn ? peer_ptr(c, sp.rank - 1, C.f + off) : nullptr, C.g.pitch, C.g.stride, nch,
                                                        has_dn ? peer_ptr(c, sp.rank - 1, ctl + 1) : nullptr, has_up ? peer_ptr(c, sp.rank + 1, ctl + 0) : nullptr);
        });
      }
    } else {
      dev_restrict_plain(c, level, L.r, C.f);
      dev_zero(c, level + 1, C.x);
    }
  }
}

void dev_restrict_plain(mgmc_ctx *c, int level, const double *r, double *fc) {
  const DevLevel &L = c->lv[level], &C = c->lv[level + 1];
  c->launch("restrict", level, [&] {
    if (L.d3) restrict27_kernel<<<grid_sites(C.g, c->d.nchains), kBlockSites, 0, c->stream>>>(L.g, L.q, C.g, C.q, r, fc);
    else restrict_kernel<<<grid_sites(C.g, c->d.nchains), kBlockSites, 0, c->stream>>>(L.g, C.g, r, fc, RowRange{1, C.g.ny - 1});
  });
}

// dense factor of the coarsest level (cholesky_sampler.cc:25-38), built at first use: contexts that only
// serve single-level operations (apply, smoothers, transfers) never need it
void ensure_coarse(mgmc_ctx *c) {
  if (c->dTT) return;
  const DevLevel &LC = c->lv[c->d.nlevel - 1];
  if (LC.h.ndof() > 4096) fail(MGMC_ERR_UNSUPPORTED, "coarsest level has more than 4096 unknowns: increase nlevel (dense coarse factor)");
  CoarseFactor cf;
  try {
    cf = coarse_factor(LC.h, c->Sigma);
  } catch (const std::exception &e) {
    fail(MGMC_ERR_INVALID, e.what());
  }
  c->Nc = cf.N;
  c->Ncp = cf.Np;
  c->dTT = c->dupload(cf.TT);
  {
    // A^{-1} = L^{-T} L^{-1} for the one-pass coarse phase (tail.cuh): row i of TT times T
    const int Np = cf.Np;
    std::vector<double> Ainv((size_t)Np * Np, 0.0);
    for (int i = 0; i < cf.N; ++i) {
      double *ai = &Ainv[(size_t)i * Np];
      for (int k = i; k < cf.N; ++k) {
        const double a = cf.TT[(size_t)i * Np + k];
        const double *tk = &cf.T[(size_t)k * Np];
        for (int j = 0; j <= k; ++j) ai[j] += a * tk[j];
      }
    }
    c->dAinv = c->dupload(Ainv);
  }
  if (!c->d_tail_bar) c->d_tail_bar = c->dalloc<unsigned long long>(1);
  if (c->d.nchains > kCoarseBatch) c->d_coarse_xi = c->dalloc<double>((size_t)c->d.nchains * cf.Np);
  c->sync();
}

void tail_flush(mgmc_ctx *c, bool gibbs, int level);

// Coarsest level: x = A^{-1} f (+ L^{-T} xi) in one pass over the chip (tail.cuh coarse_phase) -- a phase of the
// persistent kernel of the small levels, or a launch of its own when called on its own
void dev_coarse(mgmc_ctx *c, bool sample, const double *f, double *x) {
  ensure_coarse(c);
  const int lc = c->d.nlevel - 1;
  const DevLevel &L = c->lv[lc];
  if (f != L.f || x != L.x) fail(MGMC_ERR_INVALID, "internal: coarse phase on foreign vectors");
  c->next_x_zero = false;  // (the coarse solve overwrites x: nothing to zero)
  const uint32_t c1 = next_c1(c, lc, sample);
  const bool standalone = !c->tail_rec;
  if (sample && !c->perf_no_noise && c->d.nchains > kCoarseBatch) {
    // ensembles: the normals of all chains once, ahead of the phase (a launch of its own also in front of the
    // persistent kernel: they only depend on the counters)
    if (!c->d_coarse_xi) fail(MGMC_ERR_INVALID, "internal: coarse noise buffer missing");
    const int pairs = c->d.nchains * (c->Ncp / 2);
    c->launch("coarse_xi", lc, [&] {
      coarse_xi_kernel<<<std::min((pairs + 255) / 256, 4 * c->num_sms), 256, 0, c->stream>>>(noise_params(c, lc, 0), c1, c->Nc, c->Ncp, c->d.nchains, c->d_coarse_xi);
    });
  }
  TailPhase ph;
  std::memset(&ph, 0, sizeof(ph));
  ph.kind = TAIL_COARSE;
  ph.nc = sample ? 1 : 0;
  ph.c1 = (int)c1;
  c->tail_ph.push_back(ph);
  c->tail_smem = std::max(c->tail_smem, coarse_phase_smem(c->Ncp, c->Nc, c->num_sms, c->d.nchains, coarse_phase_stage(c->Ncp, c->Nc, c->num_sms, c->d.nchains, kTailSmemMax)));
  c->tail_bytes += 8.0 * c->Nc * c->Nc * c->d.nchains;
  if (standalone) tail_flush(c, sample, lc);
}

// ---------------------------------------------------------------------------------------------
// persistent kernel of the small levels (tail.cuh)
// ---------------------------------------------------------------------------------------------
// First level of the tail: every level from there down must be a radius-1 level that is not distributed over row
// strips, small enough that all tile jobs of a chain are resident at once (<= one CTA per SM) and that the chains
// take at most two waves, and -- with a low-rank term -- able to run its fix-ups inside the launch.
void plan_tail(mgmc_ctx *c) {
  if (c->tail_level >= 0) return;
  const int nl = c->d.nlevel;
  c->tail_level = nl;
  // Measured (profiles/r02_tail.md): a phase of the persistent kernel costs what a launch inside the CUDA graph costs
  // (5-7 us: the in-order latency of the tile code, not the launch), so the kernel is at parity with the launch
  // sequence -- it is opt-in (MGMC_TAIL=1) until the per-phase latency is below that of a launch; the coarsest-level
  // solve always runs as its one-pass phase.
  static const bool on = std::getenv("MGMC_TAIL") != nullptr && std::getenv("MGMC_NO_TAIL") == nullptr;
  if (!on) return;
  static const char *ms = std::getenv("MGMC_TAIL_MAX_SITES");  // (perf experiments)
  const long long max_sites = ms ? std::atoll(ms) : 512ll * 512ll;
  for (int l = nl - 1; l >= 0; --l) {
    const DevLevel &L = c->lv[l];
    if (L.generic() || (c->strip.on() && l < c->strip.ndist)) break;
    const bool smoothed = (l < nl - 1) || c->d.coarse_solver != MGMC_COARSE_CHOLESKY;
    if (l == nl - 1 && !smoothed && L.h.ndof() > 4096) break;
    if (smoothed) {
      if ((long long)L.g.nx * L.g.ny > max_sites) break;
      const int nt = tail_tiles_bound(L.g.nx, L.g.ny, L.h.st.ncolours);
      if (nt > c->num_sms || (long long)nt * c->d.nchains > 2ll * c->num_sms) break;
      if (c->d.m_lowrank > 0) {
        const LowRankDev &lr = get_lowrank(c, l, c->d.omega);
        if (!lr_fusable(c, lr, l)) break;
      }
    }
    c->tail_level = l;
  }
  if (c->tail_level < nl) {
    if (!c->d_tail_bar) c->d_tail_bar = c->dalloc<unsigned long long>(1);
    static const bool stamps = std::getenv("MGMC_TAIL_STAMPS") != nullptr;
    if (stamps && !c->d_tail_stamps) c->d_tail_stamps = c->dalloc<long long>(256);
    if (c->d.coarse_solver == MGMC_COARSE_CHOLESKY) ensure_coarse(c);
    c->sync();
  }
}

template <bool G_, bool LR_>
void launch_tail_t(mgmc_ctx *c, const TailP &T, size_t smem) {
  const void *fn = (const void *)tail_kernel<G_, LR_>;
  if (c->func_attr_done.insert(fn).second) CUDA_CHECK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, kTailSmemMax));
  cudaLaunchConfig_t cfg;
  std::memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(c->num_sms, 1, 1);  // one CTA per SM: the cooperative launch guarantees that all of them are resident
  // (tile phases need the block size of the tile code; the coarse solve on its own takes the full kTailThreads)
  bool tiles = false;
  for (int k = 0; k < T.nphase; ++k) tiles = tiles || (T.ph[k].kind == TAIL_FUSED);
  cfg.blockDim = dim3(tiles ? kFusedThreads : kTailThreads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = c->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative;
  at[0].val.cooperative = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  CUDA_CHECK(cudaLaunchKernelEx(&cfg, tail_kernel<G_, LR_>, T));
}

// launch what the recursion below tail_level has recorded
void tail_flush(mgmc_ctx *c, bool gibbs, int level) {
  const std::vector<TailPhase> ph = c->tail_ph;
  const size_t smem = std::max(c->tail_smem, (size_t)1024);
  const double bytes = c->tail_bytes;
  const bool lr = c->tail_lowrank;
  c->tail_ph.clear();
  c->tail_smem = 0;
  c->tail_bytes = 0.0;
  c->tail_lowrank = false;
  if (smem > (size_t)kTailSmemMax) fail(MGMC_ERR_INVALID, "internal: tail phase does not fit in shared memory");
  const int lc = c->d.nlevel - 1;
  for (size_t p0 = 0; p0 < ph.size(); p0 += kMaxTailPhases) {
    // (more phases than the parameter space holds -- deep W-cycles: several launches, the kernel boundary is the barrier)
    TailP T;
    std::memset(&T, 0, sizeof(T));
    T.nphase = (int)std::min<size_t>(kMaxTailPhases, ph.size() - p0);
    T.nchains = c->d.nchains;
    T.bar = c->d_tail_bar;
    T.stamps = (p0 == 0) ? c->d_tail_stamps : nullptr;
    T.nz = noise_params(c, lc, 0);
    if (c->dAinv) {
      const DevLevel &LC = c->lv[lc];
      T.coarse.Ainv = c->dAinv;
      T.coarse.TT = c->dTT;
      T.coarse.N = c->Nc;
      T.coarse.Np = c->Ncp;
      T.coarse.w = LC.g.nx - 1;
      T.coarse.h = LC.h.d3() ? LC.h.ny - 1 : INT_MAX;
      T.coarse.prow = LC.h.d3() ? LC.h.ny + 1 : 0;
      T.coarse.pitch = LC.g.pitch;
      T.coarse.stride = LC.g.stride;
      T.coarse.f = LC.f;
      T.coarse.x = LC.x;
      T.coarse.xi_pre = c->d_coarse_xi;  // (nullptr for up to kCoarseBatch chains)
      T.coarse.stage = coarse_phase_stage(c->Ncp, c->Nc, c->num_sms, c->d.nchains, kTailSmemMax) ? 1 : 0;
    }
    if (p0 == 0) c->tail_stamp_kinds.clear();
    for (int k = 0; k < T.nphase; ++k) {
      T.ph[k] = ph[p0 + k];
      if (k > 0 && T.ph[k].kind == TAIL_COARSE) T.prefetch_coarse = 1;
      if (p0 == 0) c->tail_stamp_kinds.push_back(T.ph[k].kind == TAIL_FUSED ? 100 * (T.ph[k].P.rt_restrict ? 1 : 2) + T.ph[k].ntiles : -T.ph[k].kind);
    }
    const std::string name = std::string(gibbs ? "tail_gibbs" : "tail_sor") + std::to_string(T.nphase);
#ifdef MGMC_TILE_TIMING
    // debug build: per-CTA phase stamps of the 6th multi-phase tail launch (tile stamps of fused.cuh + barrier stamps)
    static int tail_dumped = 0;
    const char *tfile = std::getenv("MGMC_TIMING_FILE");
    long long *d_tt = nullptr, *d_cs = nullptr;
    const size_t G_ = (size_t)c->num_sms;
    if (tfile && T.nphase > 1 && tail_dumped++ == 5) {
      CUDA_CHECK(cudaMalloc(&d_tt, T.nphase * G_ * 16 * sizeof(long long)));
      CUDA_CHECK(cudaMemset(d_tt, 0, T.nphase * G_ * 16 * sizeof(long long)));
      CUDA_CHECK(cudaMalloc(&d_cs, T.nphase * G_ * 4 * sizeof(long long)));
      CUDA_CHECK(cudaMemset(d_cs, 0, T.nphase * G_ * 4 * sizeof(long long)));
      for (int k = 0; k < T.nphase; ++k) T.ph[k].P.timing = d_tt + (size_t)k * G_ * 16;
      T.cta_stamps = d_cs;
    }
#endif
    c->launch(name.c_str(), level, [&] {
      if (gibbs) {
        if (lr) launch_tail_t<true, true>(c, T, smem);
        else launch_tail_t<true, false>(c, T, smem);
      } else {
        if (lr) launch_tail_t<false, true>(c, T, smem);
        else launch_tail_t<false, false>(c, T, smem);
      }
    }, p0 == 0 ? bytes : 0.0);
#ifdef MGMC_TILE_TIMING
    if (d_tt) {
      c->sync();
      std::vector<long long> h(T.nphase * G_ * 16), hc(T.nphase * G_ * 4);
      CUDA_CHECK(cudaMemcpy(h.data(), d_tt, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
      CUDA_CHECK(cudaMemcpy(hc.data(), d_cs, hc.size() * sizeof(long long), cudaMemcpyDeviceToHost));
      cudaFree(d_tt);
      cudaFree(d_cs);
      FILE *fp = std::fopen((std::string(tfile) + ".tail.txt").c_str(), "w");
      for (int k = 0; k < T.nphase; ++k)
        for (size_t b = 0; b < G_; ++b) {
          std::fprintf(fp, "%d %d %d %zu ", k, T.ph[k].kind, T.ph[k].ntiles, b);
          for (int q = 0; q < 4; ++q) std::fprintf(fp, "%lld ", hc[(k * G_ + b) * 4 + q]);
          for (int q = 0; q < 16; ++q) std::fprintf(fp, "%lld ", h[(k * G_ + b) * 16 + q]);
          std::fprintf(fp, "\n");
        }
      std::fclose(fp);
    }
#endif
  }
}

// ---------------------------------------------------------------------------------------------
// noise of the small levels ahead of their launches (noise_ahead.cuh)
// ---------------------------------------------------------------------------------------------
// In use for one chain per context when the cycle has big levels to hide the generation behind: level 0 has >= 1 M
// sites and the first small level (<= 512 x 512 sites) is level 2 or deeper.  V-cycles only (a W-cycle visits the
// small levels several times per cycle).  The branch forks where the recursion enters level 1.
void plan_nza(mgmc_ctx *c) {
  if (c->nza_planned) return;
  c->nza_planned = 1;
  // Measured (profiles/r02_noise_ahead.md): no gain -- a small-level launch is a chain of many latencies (tile load,
  // pass set-up, fix-up exchanges) of which the normals are one; reading them costs an L2 round trip per pass instead.
  // Opt-in (MGMC_NOISE_AHEAD=1) for experiments; tests/test_gpu_invariance.py keeps it bit-identical.
  static const bool on = std::getenv("MGMC_NOISE_AHEAD") != nullptr;
  const mgmc_desc &d = c->d;
  if (!on || c->perf_no_noise || d.nchains != 1 || d.cycle != 1 || d.nlevel < 4) return;
  if ((long long)c->lv[0].g.nx * c->lv[0].g.ny < (1ll << 20)) return;
  const int last = (d.coarse_solver == MGMC_COARSE_CHOLESKY) ? d.nlevel - 2 : d.nlevel - 1;  // last smoothed level
  for (int l = 2; l <= last; ++l) {
    const DevLevel &L = c->lv[l];
    if ((long long)L.g.nx * L.g.ny > 512ll * 512ll) continue;
    if (L.generic() || L.h.st.ncolours != 4 || (c->strip.on() && l < c->strip.ndist) || (c->tail_level >= 0 && l >= c->tail_level)) continue;
    c->nza_levels.push_back(l);
  }
  if (c->nza_levels.empty()) return;
  if (d.m_lowrank > 0)
    for (int l : c->nza_levels) get_lowrank(c, l, d.omega);
  // planning run: the launches of these levels in one cycle, in order
  const std::vector<uint32_t> sweeps0 = c->sweep_counter;
  const int64_t count0 = c->launch_count;
  const int lr_slot0 = c->lr_slot_next;
  c->nza_dry = true;
  try {
    for (int l : c->nza_levels) {
      c->sweep_counter[l] = 0u;
      if (l == d.nlevel - 1) {
        emit_smoothing(c, l, sweep_list(MGMC_SMOOTHER_SSOR, MGMC_FORWARD, d.ncoarsesmooth, true), true, d.omega, false, 0.0, false);
      } else {
        emit_smoothing(c, l, sweep_list(d.smoother, MGMC_FORWARD, d.npresmooth, true), true, d.omega, false, 0.0, true);
        emit_smoothing(c, l, sweep_list(d.smoother, MGMC_BACKWARD, d.npostsmooth, true), true, d.omega, true, d.coarse_scaling, false);
      }
    }
  } catch (...) {
    c->nza_dry = false;
    c->sweep_counter = sweeps0;
    throw;
  }
  c->nza_dry = false;
  c->sweep_counter = sweeps0;
  c->launch_count = count0;
  c->lr_slot_next = lr_slot0;
  bool any = false;
  for (const auto &sl : c->nza_slots) any = any || !sl.jobs.empty();
  if (!any) return;
  int lo = 0, hi = 0;
  CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  CUDA_CHECK(cudaStreamCreateWithPriority(&c->stream2, cudaStreamNonBlocking, lo));  // lowest priority: fills idle SMs
  CUDA_CHECK(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
  CUDA_CHECK(cudaEventCreateWithFlags(&c->ev_gen, cudaEventDisableTiming));
  c->nza_cursor.assign(d.nlevel, 0);
  c->nza_fork_level = 1;
  c->nza_on = true;
  c->sync();
}

// generate the planes of all launches of a level in this cycle
void nza_launch_gen(mgmc_ctx *c, int level, cudaStream_t stream) {
  std::vector<NzJob> jobs;
  for (const auto &sl : c->nza_slots)
    if (sl.level == level) jobs.insert(jobs.end(), sl.jobs.begin(), sl.jobs.end());
  if (jobs.empty()) return;
  const DevLevel &L = c->lv[level];
  for (size_t j0 = 0; j0 < jobs.size(); j0 += kMaxNzJobs) {
    NzGenP G;
    std::memset(&G, 0, sizeof(G));
    G.nz = noise_params(c, level, 0);
    G.nx = L.g.nx;
    G.ny = L.g.ny;
    G.gp = L.g.pitch / 4;
    G.njobs = (int)std::min<size_t>(kMaxNzJobs, jobs.size() - j0);
    for (int k = 0; k < G.njobs; ++k) G.job[k] = jobs[j0 + k];
    const int units = G.njobs * (L.g.ny / 2) * ((int)(G.nz.G + 31) / 32);  // (job, row, chunk of 32 groups) items, one per warp
    const int grid = std::max(1, std::min(2 * c->num_sms, (units + 2 * (kNzGenThreads / 32) - 1) / (2 * (kNzGenThreads / 32))));
    ++c->launch_count;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (c->prof_on) {  // (timed on the stream it runs on)
      CUDA_CHECK(cudaEventCreate(&e0));
      CUDA_CHECK(cudaEventCreate(&e1));
      CUDA_CHECK(cudaEventRecord(e0, stream));
    }
    noise_gen_kernel<4><<<grid, kNzGenThreads, 0, stream>>>(G);
    CUDA_CHECK(cudaGetLastError());
    if (c->prof_on) {
      CUDA_CHECK(cudaEventRecord(e1, stream));
      c->prof_events.push_back({"noise_ahead/L" + std::to_string(level), e0, e1, 0.0});
    }
  }
}

void nza_begin(mgmc_ctx *c) {
  if (!c->nza_on) return;
  std::fill(c->nza_cursor.begin(), c->nza_cursor.end(), 0);
  c->nza_forked = c->nza_joined = false;
  c->nza_in_cycle = true;
}

// the branch: the normals of this cycle's launches of the small levels, the levels visited first first
void nza_fork(mgmc_ctx *c) {
  c->nza_forked = true;
  CUDA_CHECK(cudaEventRecord(c->ev_fork, c->stream));
  CUDA_CHECK(cudaStreamWaitEvent(c->stream2, c->ev_fork, 0));
  for (int l : c->nza_levels) nza_launch_gen(c, l, c->stream2);
  CUDA_CHECK(cudaEventRecord(c->ev_gen, c->stream2));
}

void nza_end(mgmc_ctx *c) {
  if (!c->nza_on) return;
  c->nza_in_cycle = false;
  if (c->nza_forked && !c->nza_joined) CUDA_CHECK(cudaStreamWaitEvent(c->stream, c->ev_gen, 0));  // (a captured branch must be joined)
}

// ---------------------------------------------------------------------------------------------
// multilevel recursions
// ---------------------------------------------------------------------------------------------
void mgmc_sample_level_body(mgmc_ctx *c, int level);
void mgmc_sample_level(mgmc_ctx *c, int level) {  // multigridmc_sampler.cc:103-130
  if (c->nza_on && c->nza_in_cycle && level == c->nza_fork_level && !c->nza_forked) nza_fork(c);
  if (level == c->tail_level && !c->tail_rec) {
    // this level and everything below it: phases of one persistent launch
    c->tail_rec = true;
    try {
      mgmc_sample_level_body(c, level);
    } catch (...) {
      c->tail_rec = false;
      c->tail_ph.clear();
      throw;
    }
    c->tail_rec = false;
    tail_flush(c, !c->perf_no_noise, level);
    return;
  }
  mgmc_sample_level_body(c, level);
}
void mgmc_sample_level_body(mgmc_ctx *c, int level) {
  const mgmc_desc &d = c->d;
  if (level == d.nlevel - 1) {
    DevLevel &L = c->lv[level];
    if (d.coarse_solver == MGMC_COARSE_CHOLESKY) {
      dev_coarse(c, true, L.f, L.x);
    } else {
      emit_smoothing(c, level, sweep_list(MGMC_SMOOTHER_SSOR, MGMC_FORWARD, d.ncoarsesmooth, true), true, d.omega, false, 0.0, false);
      normalize_x(c, level);
    }
    return;
  }
  const int cycle_ = (level > 0) ? d.cycle : 1;
  for (int j = 0; j < cycle_; ++j) {
    // presampler + residual + restrict (+ x_{l+1} = 0)
    const bool gibbs = !c->perf_no_noise;
    emit_smoothing(c, level, sweep_list(d.smoother, MGMC_FORWARD, d.npresmooth, true), gibbs, d.omega, false, 0.0, true);
    mgmc_sample_level(c, level + 1);
    // prolongate_add + postsampler
    emit_smoothing(c, level, sweep_list(d.smoother, MGMC_BACKWARD, d.npostsmooth, true), gibbs, d.omega, true, d.coarse_scaling, false);
    normalize_x(c, level);
  }
}

void mg_solve_level_body(mgmc_ctx *c, int level);
void mg_solve_level(mgmc_ctx *c, int level) {  // multigrid_preconditioner.cc:74-101
  if (level == c->tail_level && !c->tail_rec) {
    c->tail_rec = true;
    try {
      mg_solve_level_body(c, level);
    } catch (...) {
      c->tail_rec = false;
      c->tail_ph.clear();
      throw;
    }
    c->tail_rec = false;
    tail_flush(c, false, level);
    return;
  }
  mg_solve_level_body(c, level);
}
void mg_solve_level_body(mgmc_ctx *c, int level) {
  const mgmc_desc &d = c->d;
  DevLevel &L = c->lv[level];
  if (level == d.nlevel - 1) {
    dev_coarse(c, false, L.f, L.x);  // writes every interior entry; ghost lines stay zero
    return;
  }
  if (level == 0) {
    lr_begin_epoch(c);
    dev_zero(c, level, L.x);  // deeper levels are zeroed by the restriction that feeds them
  }
  const int cycle_ = (level > 0) ? d.cycle : 1;
  for (int j = 0; j < cycle_; ++j) {
    emit_smoothing(c, level, sweep_list(d.smoother, MGMC_FORWARD, d.npresmooth, false), false, d.omega, false, 0.0, true);
    mg_solve_level(c, level + 1);
    emit_smoothing(c, level, sweep_list(d.smoother, MGMC_BACKWARD, d.npostsmooth, false), false, d.omega, true, d.coarse_scaling, false);
    normalize_x(c, level);
  }
}

void emit_mgmc_cycle(mgmc_ctx *c) {
  lr_begin_epoch(c);
  c->strip_index = 0;
  std::fill(c->sweep_counter.begin(), c->sweep_counter.end(), 0u);
  nza_begin(c);
  try {
    mgmc_sample_level(c, 0);
  } catch (...) {
    c->nza_in_cycle = false;
    throw;
  }
  nza_end(c);
}

void emit_end_of_cycle(mgmc_ctx *c, bool merged = false) {
  const DevLevel &L = c->lv[0];
  c->launch("end_of_cycle", 0, [&] {
    end_of_cycle_kernel<<<1, 256, 0, c->stream>>>(c->qoi_nnz, c->d_qsite, c->d_qval, L.x, L.g.stride, c->d.nchains, c->d_series, c->series_cap / c->d.nchains, c->d_sample,
                                                 c->d_pos, (c->strip.on() && c->strip_connected) ? c->d_strip_ctl + 9 : nullptr, merged ? c->d_qpart : nullptr,
                                                 merged ? c->d_lr_epoch : nullptr);
  });
  c->h_sample++;
}

// ---- merged level-0 launches (mgmc_ctx::merge_on) ----
// Level 0 is visited once per cycle (also in a W-cycle): post-smoothing of cycle k, then pre-smoothing of cycle k + 1.
// Merged they are one launch of at most 8 colour passes -- red-black levels, SOR / SSOR with one sweep pair each --
// with up to 4 low-rank fix-ups; with omega = 1 the passes that are recomputed before anybody reads them are dead
// (plan_stages): red-black SSOR V(1,1) runs 5 full passes of 8 (+ the observed / measured sites of a sixth).
void plan_merge(mgmc_ctx *c) {
  static const bool off = std::getenv("MGMC_NO_MERGE") != nullptr;
  const mgmc_desc &d = c->d;
  c->merge_on = false;
  if (off || c->strip.on() || d.nlevel < 2 || c->tail_level == 0 || c->perf_no_noise) return;
  const DevLevel &L = c->lv[0];
  if (L.generic() || L.h.st.ncolours != 2 || d.omega != 1.0) return;
  const size_t npre = sweep_list(d.smoother, MGMC_FORWARD, d.npresmooth, true).size(), npost = sweep_list(d.smoother, MGMC_BACKWARD, d.npostsmooth, true).size();
  if (npre == 0 || npost == 0 || 2 * (npre + npost) > 8) return;
  if (d.m_lowrank > 0 && ((int)(npre + npost) > kMaxFix || !lr_fusable(c, get_lowrank(c, 0, d.omega), 0))) return;
  if (c->qoi_nnz > kMaxQoi) return;
  if (!c->d_qpart) c->d_qpart = c->dalloc<double>((size_t)d.nchains * kMaxQoi);
  c->merge_on = true;
}

void emit_merged_level0(mgmc_ctx *c) {
  const mgmc_desc &d = c->d;
  const bool lowrank = d.m_lowrank > 0;
  const std::vector<SweepSpec> pre = sweep_list(d.smoother, MGMC_FORWARD, d.npresmooth, true), post = sweep_list(d.smoother, MGMC_BACKWARD, d.npostsmooth, true);
  std::vector<Stage> st;
  std::vector<FixSpec> fixes;
  auto add = [&](const std::vector<SweepSpec> &sweeps, uint32_t first_sweep, uint32_t soff) {
    uint32_t k = first_sweep;
    for (const SweepSpec &sw : sweeps) {
      const uint32_t c1 = k++ & 0xFFFFFFu;  // level 0: (0 << 24) | sweep counter (next_c1)
      for (int cc = 0; cc < 2; ++cc) {
        Stage sg;
        std::memset(&sg, 0, sizeof(sg));
        sg.colour = sw.fwd ? cc : 1 - cc;
        sg.c1 = c1;
        sg.soff = soff;
        st.push_back(sg);
      }
      if (lowrank && sw.fix_after) {
        FixSpec fx{(int)st.size() - 1, sw.fwd ? 0 : 1, c1};
        fx.soff = soff;
        fixes.push_back(fx);
      }
    }
  };
  add(post, (uint32_t)pre.size(), 0u);  // cycle k: the sweep counters of level 0 continue behind the pre-smoothing sweeps
  c->merge_qoi_stage = (int)st.size() - 1;
  add(pre, 0u, 1u);                     // cycle k + 1
  try {
    dev_fused(c, 0, st, fixes, lowrank, true, d.omega, true, d.coarse_scaling, true);
  } catch (...) {
    c->merge_qoi_stage = -1;
    throw;
  }
  c->merge_qoi_stage = -1;
  c->sweep_counter[0] = (uint32_t)pre.size();
}

// pre-smoothing of the first cycle + the levels below
void emit_merge_prologue(mgmc_ctx *c) {
  const mgmc_desc &d = c->d;
  lr_begin_epoch(c);
  std::fill(c->sweep_counter.begin(), c->sweep_counter.end(), 0u);
  emit_smoothing(c, 0, sweep_list(d.smoother, MGMC_FORWARD, d.npresmooth, true), true, d.omega, false, 0.0, true);
  nza_begin(c);
  try {
    mgmc_sample_level(c, 1);
  } catch (...) {
    c->nza_in_cycle = false;
    throw;
  }
  nza_end(c);
  // (the merged launch that follows reuses the packet slots of the pre-smoothing: new epoch)
  if (c->d_lr_epoch) c->launch("lr_epoch", 0, [&] { bump_kernel<<<1, 1, 0, c->stream>>>(c->d_lr_epoch); });
}

// the unit of a run: merged launch (end of cycle k, start of cycle k + 1), end of cycle k, levels >= 1 of cycle k + 1
void emit_merge_unit(mgmc_ctx *c) {
  c->lr_slot_next = 0;
  std::fill(c->sweep_counter.begin(), c->sweep_counter.end(), 0u);
  emit_merged_level0(c);
  emit_end_of_cycle(c, true);
  c->next_x_zero = true;  // the merged launch does not zero the iterate of level 1
  nza_begin(c);
  try {
    mgmc_sample_level(c, 1);
  } catch (...) {
    c->next_x_zero = false;
    c->nza_in_cycle = false;
    throw;
  }
  nza_end(c);
  if (c->next_x_zero) {
    c->next_x_zero = false;
    fail(MGMC_ERR_INVALID, "internal: level 1 did not start with a fused launch");
  }
}

// post-smoothing of the last cycle
void emit_merge_epilogue(mgmc_ctx *c) {
  const mgmc_desc &d = c->d;
  c->lr_slot_next = 0;
  c->sweep_counter[0] = (uint32_t)sweep_list(d.smoother, MGMC_FORWARD, d.npresmooth, true).size();
  emit_smoothing(c, 0, sweep_list(d.smoother, MGMC_BACKWARD, d.npostsmooth, true), true, d.omega, true, d.coarse_scaling, false);
  normalize_x(c, 0);
  emit_end_of_cycle(c);
}

// after the stream has been synchronised: did a device-side wait time out?  (then the state is invalid)
void check_device_error(mgmc_ctx *c) {
  if (!c->d_err || c->d.m_lowrank == 0) return;
  int e = 0;
  CUDA_CHECK(cudaMemcpy(&e, c->d_err, sizeof(int), cudaMemcpyDeviceToHost));
  if (e) {
    CUDA_CHECK(cudaMemset(c->d_err, 0, sizeof(int)));
    fail(MGMC_ERR_CUDA, "a device-side wait for a low-rank packet of another tile timed out (co-residency lost or a peer died); the chain state is invalid");
  }
}

void set_sample_index(mgmc_ctx *c, uint32_t s) {
  c->h_sample = s;
  CUDA_CHECK(cudaMemcpyAsync(c->d_sample, &c->h_sample, sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
  c->sync();
}

void drop_graph(mgmc_ctx *c) {
  if (c->graph) {
    cudaGraphExecDestroy(c->graph);
    c->graph = nullptr;
  }
  for (cudaGraphExec_t &g : c->graph_unit)
    if (g) {
      cudaGraphExecDestroy(g);
      g = nullptr;
    }
}

void ensure_series(mgmc_ctx *c, long long n) {
  if (n > c->series_cap) {
    c->sync();
    drop_graph(c);
    c->dfree(c->d_series);
    c->d_series = c->dalloc<double>((size_t)n);
    c->series_cap = n;
  }
}

// row strips: the number of distributed fused launches per cycle enters the flag protocol; count it by
// capturing (and discarding) one cycle
void strip_count_launches(mgmc_ctx *c) {
  if (!(c->strip.on() && c->strip_connected) || c->strip_per_cycle > 0) return;
  if (c->d.m_lowrank > 0)
    for (int l = 0; l < c->d.nlevel; ++l) get_lowrank(c, l, c->d.omega);
  if (c->d.coarse_solver == MGMC_COARSE_CHOLESKY) ensure_coarse(c);
  c->sync();
  const int64_t count0 = c->launch_count;
  const std::vector<uint32_t> sweeps0 = c->sweep_counter;
  const bool prof0 = c->prof_on;
  c->prof_on = false;
  cudaGraph_t g = nullptr;
  CUDA_CHECK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
  try {
    emit_mgmc_cycle(c);
  } catch (...) {
    cudaStreamEndCapture(c->stream, &g);
    if (g) cudaGraphDestroy(g);
    c->prof_on = prof0;
    throw;
  }
  CUDA_CHECK(cudaStreamEndCapture(c->stream, &g));
  CUDA_CHECK(cudaGraphDestroy(g));
  c->prof_on = prof0;
  c->strip_per_cycle = c->strip_index;
  c->launch_count = count0;
  c->sweep_counter = sweeps0;
}

// K >= 2 cycles with merged level-0 launches (mgmc_ctx::merge_on); nsamples = 0: only instantiate the graphs
void run_cycles_merged(mgmc_ctx *c, int64_t nsamples) {
  DevLevel &L0 = c->lv[0];
  if (L0.x != L0.x_primary) fail(MGMC_ERR_INVALID, "internal: level-0 iterate not in its primary buffer at the start of a run");
  const bool use_g = c->use_graph && !c->prof_on;
  if (use_g && !c->graph_unit[0]) {
    // (lazily built low-rank data must exist before capture: uploads are not capturable)
    if (c->d.m_lowrank > 0)
      for (int l = 0; l < c->d.nlevel; ++l) get_lowrank(c, l, c->d.omega);
    if (c->d.coarse_solver == MGMC_COARSE_CHOLESKY) ensure_coarse(c);
    c->sync();
    const int64_t count0 = c->launch_count;
    const uint32_t sample0 = c->h_sample;
    const std::vector<uint32_t> sweeps0 = c->sweep_counter;
    const int slot0 = c->lr_slot_next;
    // one graph per parity of the ping-pong buffers of level 0 (the merged launch is out of place: x -> x_alt); the
    // capture swaps the host-side pointers like a launch does, so the second capture is the other parity
    for (int parity = 0; parity < 2; ++parity) {
      const int64_t before = c->launch_count;
      cudaGraph_t g = nullptr;
      CUDA_CHECK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
      try {
        emit_merge_unit(c);
      } catch (...) {
        cudaStreamEndCapture(c->stream, &g);
        if (g) cudaGraphDestroy(g);
        if (L0.x != L0.x_primary) std::swap(L0.x, L0.x_alt);
        throw;
      }
      CUDA_CHECK(cudaStreamEndCapture(c->stream, &g));
      CUDA_CHECK(cudaGraphInstantiate(&c->graph_unit[parity], g, 0));
      CUDA_CHECK(cudaGraphDestroy(g));
      c->unit_launches = c->launch_count - before;
    }
    c->launch_count = count0;  // capture itself launched nothing
    c->h_sample = sample0;
    c->sweep_counter = sweeps0;
    c->lr_slot_next = slot0;
    if (L0.x != L0.x_primary) fail(MGMC_ERR_INVALID, "internal: parity of the level-0 buffers after capture");
  }
  if (nsamples == 0) return;
  emit_merge_prologue(c);
  for (int64_t k = 1; k < nsamples; ++k) {
    if (use_g) {
      CUDA_CHECK(cudaGraphLaunch(c->graph_unit[L0.x == L0.x_primary ? 0 : 1], c->stream));
      c->launch_count += c->unit_launches;
      c->h_sample++;
      std::swap(L0.x, L0.x_alt);
    } else {
      const int64_t before = c->launch_count;
      emit_merge_unit(c);
      c->unit_launches = c->launch_count - before;
    }
  }
  emit_merge_epilogue(c);
  c->launches_per_cycle = c->unit_launches;
}

void run_cycles(mgmc_ctx *c, int64_t nsamples) {
  plan_tail(c);
  plan_nza(c);
  strip_count_launches(c);
  plan_merge(c);
  const unsigned long long zero = 0ull;
  CUDA_CHECK(cudaMemcpyAsync(c->d_pos, &zero, sizeof(zero), cudaMemcpyHostToDevice, c->stream));
  if (c->merge_on && nsamples != 1) {
    run_cycles_merged(c, nsamples);
    if (nsamples > 0) return;
  }
  if (c->use_graph && !c->prof_on) {
    if (!c->graph) {
      // make sure lazily built low-rank data exists before capture (uploads are not capturable)
      if (c->d.m_lowrank > 0)
        for (int l = 0; l < c->d.nlevel; ++l) get_lowrank(c, l, c->d.omega);
      if (c->d.coarse_solver == MGMC_COARSE_CHOLESKY) ensure_coarse(c);
      c->sync();
      const int64_t count0 = c->launch_count;
      const uint32_t sample0 = c->h_sample;
      cudaGraph_t g = nullptr;
      CUDA_CHECK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
      try {
        emit_mgmc_cycle(c);
        emit_end_of_cycle(c);
      } catch (...) {
        cudaStreamEndCapture(c->stream, &g);
        if (g) cudaGraphDestroy(g);
        throw;
      }
      CUDA_CHECK(cudaStreamEndCapture(c->stream, &g));
      CUDA_CHECK(cudaGraphInstantiate(&c->graph, g, 0));
      CUDA_CHECK(cudaGraphDestroy(g));
      c->launches_per_cycle = c->launch_count - count0;
      c->launch_count = count0;  // capture itself launched nothing
      c->h_sample = sample0;
    }
  }
  for (int64_t k = 0; k < nsamples; ++k) {
    if (c->graph && !c->prof_on) {
      CUDA_CHECK(cudaGraphLaunch(c->graph, c->stream));
      c->launch_count += c->launches_per_cycle;
      c->h_sample++;
    } else {
      const int64_t before = c->launch_count;
      emit_mgmc_cycle(c);
      emit_end_of_cycle(c);
      c->launches_per_cycle = c->launch_count - before;
    }
  }
}

}  // namespace

// =================================================================================================
// C ABI
// =================================================================================================
#define API_BEGIN try {
#define API_END                                   \
  }                                               \
  catch (const MgmcError &e) {                    \
    g_last_error = e.msg;                         \
    return e.code;                                \
  }                                               \
  catch (const std::exception &e) {               \
    g_last_error = e.what();                      \
    return MGMC_ERR_INVALID;                      \
  }                                               \
  return MGMC_OK;

// every entry point that takes a context passes through here: argument check + the context's device becomes current
// (a host thread may hold contexts on several devices)
static void check_level(const mgmc_ctx *c, int level, bool need_coarser = false) {
  if (!c) fail(MGMC_ERR_INVALID, "null context");
  CUDA_CHECK(cudaSetDevice(c->device));
  if (level < 0 || level >= c->d.nlevel - (need_coarser ? 1 : 0)) fail(MGMC_ERR_INVALID, "level out of range");
}

extern "C" {

int mgmc_create(const mgmc_desc *desc, mgmc_ctx **out) {
  mgmc_ctx *c = nullptr;
  try {
    if (!desc || !out) fail(MGMC_ERR_INVALID, "null argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) fail(MGMC_ERR_CUDA, "no CUDA device available: the MGMC path has no CPU fallback");
    if (desc->device < 0 || desc->device >= ndev) fail(MGMC_ERR_INVALID, "invalid device ordinal");
    if (desc->nchains < 1) fail(MGMC_ERR_INVALID, "nchains must be >= 1");
    if (desc->m_lowrank < 0 || desc->m_lowrank > 1024) fail(MGMC_ERR_UNSUPPORTED, "m_lowrank must be in [0, 1024]");
    if (desc->B_nnz < 0 || (desc->m_lowrank > 0 && !desc->Sigma) || (desc->B_nnz > 0 && (!desc->B_rows || !desc->B_cols || !desc->B_vals)))
      fail(MGMC_ERR_INVALID, "low-rank term: Sigma / B_rows / B_cols / B_vals must not be null when m_lowrank / B_nnz > 0");
    if (desc->m_lowrank == 0 && desc->B_nnz > 0) fail(MGMC_ERR_INVALID, "B entries given but m_lowrank = 0");
    if (desc->smoother != MGMC_SMOOTHER_SOR && desc->smoother != MGMC_SMOOTHER_SSOR) fail(MGMC_ERR_INVALID, "invalid smoother");
    if (desc->coarse_solver != MGMC_COARSE_SSOR && desc->coarse_solver != MGMC_COARSE_CHOLESKY) fail(MGMC_ERR_INVALID, "invalid coarse solver");
    std::vector<HostLevel> H = build_host_levels(*desc);
    for (const HostLevel &h : H)
      if (h.st.radius <= 1 && !h.st.uniform) fail(MGMC_ERR_UNSUPPORTED, "internal: radius-1 stencil with position classes");
    c = new mgmc_ctx();
    c->d = *desc;
    c->d.B_rows = nullptr;
    c->d.B_cols = nullptr;
    c->d.B_vals = nullptr;
    c->d.Sigma = nullptr;
    c->d.kappa_sq = nullptr;
    c->Sigma.assign(desc->Sigma, desc->Sigma + desc->m_lowrank);
    for (double s : c->Sigma)
      if (!(s > 0.0)) fail(MGMC_ERR_INVALID, "Sigma entries must be positive");
    c->device = desc->device;
    CUDA_CHECK(cudaSetDevice(c->device));
    CUDA_CHECK(cudaDeviceGetAttribute(&c->num_sms, cudaDevAttrMultiProcessorCount, c->device));
    {
      // (highest priority: the branch that generates noise ahead of the launches runs at the lowest, noise_ahead.cuh)
      int plo = 0, phi = 0;
      CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&plo, &phi));
      CUDA_CHECK(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, phi));
    }
    c->use_graph = (std::getenv("MGMC_NO_GRAPH") == nullptr);
    c->perf_no_noise = (std::getenv("MGMC_PERF_NO_NOISE") != nullptr);
    c->lr_fuse = (std::getenv("MGMC_NO_LR_FUSE") == nullptr);
    c->sweep_counter.assign(desc->nlevel, 0u);
    c->keys = philox_round_keys(desc->seed);
    c->lv.resize(desc->nlevel);
    c->strip = make_strip_plan(*desc, H);
    size_t arena_off = 4096;  // control words first
    if (c->strip.on()) {
      // x, x_alt and f of every level live in ONE allocation with rank-independent offsets: a neighbour's
      // array is its arena base (one CUDA IPC handle per rank) plus the same offset
      size_t bytes = arena_off;
      for (int l = 0; l < desc->nlevel; ++l) {
        const size_t pitch = ((GX + H[l].nx + 1 + 2 + 15) / 16) * 16;
        bytes += 3 * (((size_t)H[l].ny + 1 + 2 * GY) * pitch * sizeof(double) + 256);
      }
      // (+ what the owner tiles of the low-rank fix-ups publish: edge tiles write into the neighbours' copies)
      bytes += (size_t)mgmc_ctx::kLrSlots * desc->nchains * desc->m_lowrank * 2 * sizeof(LrPkt) + 512;
      c->arena_bytes = bytes;
      c->arena = (char *)c->dalloc<char>(bytes);
      c->d_strip_ctl = (int *)c->arena;
      c->peer_arena.assign(c->strip.nranks, nullptr);
    }
    auto carve = [&](size_t total) {
      double *p = (double *)(c->arena + arena_off);
      arena_off += (total * sizeof(double) + 255) / 256 * 256;
      return p;
    };
    if (desc->m_lowrank > 0) {
      const size_t n = (size_t)mgmc_ctx::kLrSlots * desc->nchains * desc->m_lowrank;
      c->d_lr_vbuf = c->strip.on() ? (LrPkt *)carve(4 * n) : c->dalloc<LrPkt>(2 * n);
      c->d_lr_epoch = c->dalloc<int>(1);
      c->lr_flag_pool_size = (size_t)1 << 20;
      c->d_lr_flag_pool = c->dalloc<unsigned char>(c->lr_flag_pool_size, false);
      CUDA_CHECK(cudaMemsetAsync(c->d_lr_flag_pool, 0xFF, c->lr_flag_pool_size, c->stream));  // 0xFF: not known yet
    }
    for (int l = 0; l < desc->nlevel; ++l) {
      DevLevel &L = c->lv[l];
      L.h = H[l];
      L.g.nx = L.h.nx;
      L.g.ny = L.h.ny;
      if (L.h.d3()) {
        // planes k = 0 .. nz stacked in the row direction: the last row (k = nz, j = ny) is a boundary row, like row 0
        L.d3 = true;
        L.q = Grid3{L.h.ny, L.h.nz};
        L.g.ny = (L.h.nz + 1) * (L.h.ny + 1) - 1;
        std::memcpy(L.c27.a, L.h.st3, sizeof(L.c27.a));
        L.full27 = (L.h.st.ncolours == 8);
      }
      L.g.pitch = ((GX + L.h.nx + 1 + 2 + 15) / 16) * 16;
      const size_t rows = (size_t)L.g.ny + 1 + 2 * GY;
      L.g.stride = (long long)rows * L.g.pitch;
      const size_t total = (size_t)L.g.stride * desc->nchains;
      const size_t origin = (size_t)GY * L.g.pitch + GX;
      L.x = (c->strip.on() ? carve(total) : c->dalloc<double>(total)) + origin;
      L.x_primary = L.x;
      L.x_alt = (c->strip.on() ? carve(total) : c->dalloc<double>(total)) + origin;
      L.f = (c->strip.on() ? carve(total) : c->dalloc<double>(total)) + origin;
      L.r = c->dalloc<double>(total) + origin;
      L.coef = to_coef9(L.h.st);
      L.nine = (L.h.st.ncolours == 4);
      L.r2 = (L.h.st.radius > 1);
      if (L.r2) {
        std::vector<double> st(&L.h.st.a[0][0], &L.h.st.a[0][0] + 225);
        L.d_st = c->dupload(st);
      }
      L.vc = L.h.varcoef();
      if (L.vc) {
        // coefficient planes in the padded layout of the vectors (zeros on the boundary / ghost lines): all nine for the
        // 9-point Galerkin operators, the diagonal alone for the 5-point fine operator (varcoef.cuh)
        const size_t np = (size_t)(L.h.nx + 1) * (L.h.ny + 1);
        L.vc_full = (l > 0);
        const int k0 = L.vc_full ? 0 : 4, k1 = L.vc_full ? 9 : 5;
        std::vector<double> planes((size_t)(k1 - k0) * L.g.stride, 0.0);
        for (int k = k0; k < k1; ++k)
          for (int j = 0; j <= L.h.ny; ++j)
            std::memcpy(&planes[(size_t)(k - k0) * L.g.stride + origin + (size_t)j * L.g.pitch], &L.h.vc[(size_t)k * np + (size_t)j * (L.h.nx + 1)], sizeof(double) * (L.h.nx + 1));
        L.dvc.a = c->dupload(planes) + origin;
        L.dvc.plane = L.g.stride;
        // (5-point: the neighbour coefficients of a vertex in the interior -- constant, shiftedlaplace_fd_operator.cc:46)
        const int im = std::min(2, L.h.nx - 1), jm = std::min(2, L.h.ny - 1);
        L.dvc.w = L.h.coef(im, jm, -1, 0);
        L.dvc.e = L.h.coef(std::max(L.h.nx - 2, 1), jm, +1, 0);
        L.dvc.s = L.h.coef(im, jm, 0, -1);
        L.dvc.n = L.h.coef(im, std::max(L.h.ny - 2, 1), 0, +1);
      }
      if (desc->m_lowrank > 0) {
        L.B = upload_sparse(c, L.h.B, desc->m_lowrank, L.g.pitch);
        // the padded one-CTA kernels stage m x (longest column) products in shared memory (lowrank_fix_kernel)
        std::vector<int> cnt(desc->m_lowrank, 0);
        int longest = 0;
        for (const SEntry &e : L.h.B) longest = std::max(longest, ++cnt[e.col]);
        const size_t mm = (size_t)desc->m_lowrank;
        L.lr_wide = (mm * longest + 3 * mm + (mm <= 48 ? 2 * mm * mm : 0)) * sizeof(double) > 40 * 1024;
        if (L.lr_wide && !c->d_lr_partial) {  // (allocated here: the first use may be inside a stream capture)
          c->d_lr_partial = c->dalloc<double>((size_t)desc->nchains * mm * kLrWideBlocks);
          c->d_lr_d = c->dalloc<double>((size_t)desc->nchains * mm);
        }
      }
    }
    if (desc->m_lowrank > 0) {
      std::vector<double> si(desc->m_lowrank), sis(desc->m_lowrank);
      for (int k = 0; k < desc->m_lowrank; ++k) {
        si[k] = 1.0 / c->Sigma[k];
        sis[k] = std::sqrt(1.0 / c->Sigma[k]);
      }
      c->d_sigma_inv = c->dupload(si);
      std::vector<double> sin(si);
      for (double &v : sin) v = -v;
      c->d_sigma_inv_neg = c->dupload(sin);
      c->d_sigma_inv_sqrt = c->dupload(sis);
    }
    c->d_sample = c->dalloc<uint32_t>(1);
    c->d_pos = c->dalloc<unsigned long long>(1);
    c->d_err = c->dalloc<int>(1);
    c->sync();
    *out = c;
    return MGMC_OK;
  } catch (const MgmcError &e) {
    g_last_error = e.msg;
    if (c) mgmc_destroy(c);
    return e.code;
  } catch (const std::exception &e) {
    g_last_error = e.what();
    if (c) mgmc_destroy(c);
    return MGMC_ERR_INVALID;
  }
}

void mgmc_destroy(mgmc_ctx *c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->graph) cudaGraphExecDestroy(c->graph);
  for (cudaGraphExec_t g : c->graph_unit)
    if (g) cudaGraphExecDestroy(g);
  if (c->mg_graph) cudaGraphExecDestroy(c->mg_graph);
  for (char *p : c->peer_arena)
    if (p) cudaIpcCloseMemHandle(p);
  for (void *p : c->allocs) cudaFree(p);
  if (c->stream2) {
    cudaStreamSynchronize(c->stream2);
    cudaStreamDestroy(c->stream2);
  }
  for (cudaEvent_t e : {c->ev_fork, c->ev_gen})
    if (e) cudaEventDestroy(e);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

int mgmc_level_info(const mgmc_ctx *c, int level, int *nx, int *ny, int64_t *ndof, int *ncolours) {
  API_BEGIN
  check_level(c, level);
  const HostLevel &h = c->lv[level].h;
  if (nx) *nx = h.nx;
  if (ny) *ny = h.ny;
  if (ndof) *ndof = h.ndof();
  if (ncolours) *ncolours = h.st.ncolours;
  API_END
}

int mgmc_level_nz(const mgmc_ctx *c, int level, int *nz) {
  API_BEGIN
  check_level(c, level);
  if (nz) *nz = c->lv[level].h.nz;
  API_END
}

int mgmc_host_stencil3(const mgmc_desc *desc, int level, double *out27, int *ncolours) {
  API_BEGIN
  if (!desc || !out27) fail(MGMC_ERR_INVALID, "null argument");
  if (desc->dim != 3) fail(MGMC_ERR_INVALID, "mgmc_host_stencil3 needs a 3d lattice (2d: mgmc_host_stencil)");
  std::vector<HostLevel> H = build_host_levels(*desc);
  if (level < 0 || level >= (int)H.size()) fail(MGMC_ERR_INVALID, "level out of range");
  std::memcpy(out27, H[level].st3, sizeof(double) * 27);
  if (ncolours) *ncolours = H[level].st.ncolours;
  API_END
}

int mgmc_get_stencil(const mgmc_ctx *c, int level, double *out225) {
  API_BEGIN
  check_level(c, level);
  if (c->lv[level].d3) fail(MGMC_ERR_INVALID, "mgmc_get_stencil: 2d lattices only (3d: mgmc_host_stencil3)");
  std::memcpy(out225, c->lv[level].h.st.a, sizeof(double) * 225);
  API_END
}

/* host-only twin of mgmc_get_stencil: runs the setup algebra without a CUDA device (used by the
 * CPU test-suite to check the Galerkin stencils against the oracle's sparse triple product) */
int mgmc_host_stencil(const mgmc_desc *desc, int level, double *out225, int *ncolours) {
  API_BEGIN
  if (desc && desc->dim == 3) fail(MGMC_ERR_INVALID, "mgmc_host_stencil: 2d lattices only (3d: mgmc_host_stencil3)");
  std::vector<HostLevel> H = build_host_levels(*desc);
  if (level < 0 || level >= (int)H.size()) fail(MGMC_ERR_INVALID, "level out of range");
  std::memcpy(out225, H[level].st.a, sizeof(double) * 225);
  if (ncolours) *ncolours = H[level].st.ncolours;
  API_END
}

int mgmc_host_coefficients(const mgmc_desc *desc, int level, double *out, int *ncolours) {
  API_BEGIN
  if (!desc || !out) fail(MGMC_ERR_INVALID, "null argument");
  if (!desc->kappa_sq) fail(MGMC_ERR_INVALID, "mgmc_host_coefficients needs desc->kappa_sq (constant coefficients: mgmc_host_stencil)");
  std::vector<HostLevel> H = build_host_levels(*desc);
  if (level < 0 || level >= (int)H.size()) fail(MGMC_ERR_INVALID, "level out of range");
  std::memcpy(out, H[level].vc.data(), sizeof(double) * H[level].vc.size());
  if (ncolours) *ncolours = H[level].st.ncolours;
  API_END
}

int mgmc_op_apply(mgmc_ctx *c, int level, const double *x, double *y) {
  API_BEGIN
  check_level(c, level);
  DevLevel &L = c->lv[level];
  upload_vec(c, level, L.x, x);
  dev_apply(c, level, L.x, L.r);
  download_vec(c, level, L.r, y);
  c->sync();
  API_END
}

int mgmc_restrict(mgmc_ctx *c, int level, const double *x_fine, double *x_coarse) {
  API_BEGIN
  check_level(c, level, true);
  upload_vec(c, level, c->lv[level].r, x_fine);
  dev_restrict_plain(c, level, c->lv[level].r, c->lv[level + 1].f);
  download_vec(c, level + 1, c->lv[level + 1].f, x_coarse);
  c->sync();
  API_END
}

int mgmc_prolongate_add(mgmc_ctx *c, int level, double alpha, const double *x_coarse, double *x_fine) {
  API_BEGIN
  check_level(c, level, true);
  upload_vec(c, level, c->lv[level].x, x_fine);
  upload_vec(c, level + 1, c->lv[level + 1].x, x_coarse);
  lr_begin_epoch(c);
  emit_smoothing(c, level, {}, false, c->d.omega, true, alpha, false);
  normalize_x(c, level);
  download_vec(c, level, c->lv[level].x, x_fine);
  c->sync();
  check_device_error(c);
  API_END
}

int mgmc_residual_restrict(mgmc_ctx *c, int level, const double *f, const double *x, double *f_coarse) {
  API_BEGIN
  check_level(c, level, true);
  upload_vec(c, level, c->lv[level].f, f);
  upload_vec(c, level, c->lv[level].x, x);
  lr_begin_epoch(c);
  emit_smoothing(c, level, {}, false, c->d.omega, false, 0.0, true);
  download_vec(c, level + 1, c->lv[level + 1].f, f_coarse);
  c->sync();
  check_device_error(c);
  API_END
}

static void check_smoother_args(int kind, int direction, double omega, int nsmooth) {
  if (kind != MGMC_SMOOTHER_SOR && kind != MGMC_SMOOTHER_SSOR) fail(MGMC_ERR_INVALID, "invalid smoother kind");
  if (direction != MGMC_FORWARD && direction != MGMC_BACKWARD) fail(MGMC_ERR_INVALID, "invalid direction");
  if (!(omega > 0.0 && omega < 2.0)) fail(MGMC_ERR_INVALID, "omega must be in (0,2)");
  if (nsmooth < 0) fail(MGMC_ERR_INVALID, "nsmooth must be >= 0");
}

int mgmc_smoother_apply(mgmc_ctx *c, int level, int kind, int direction, double omega, int nsmooth, const double *b, double *x) {
  API_BEGIN
  check_level(c, level);
  check_smoother_args(kind, direction, omega, nsmooth);
  DevLevel &L = c->lv[level];
  upload_vec(c, level, L.f, b);
  upload_vec(c, level, L.x, x);
  lr_begin_epoch(c);
  emit_smoothing(c, level, sweep_list(kind, direction, nsmooth, false), false, omega, false, 0.0, false);
  normalize_x(c, level);
  download_vec(c, level, L.x, x);
  c->sync();
  check_device_error(c);
  API_END
}

int mgmc_sampler_apply(mgmc_ctx *c, int level, int kind, int direction, double omega, int nsmooth, const double *f, double *x) {
  API_BEGIN
  check_level(c, level);
  check_smoother_args(kind, direction, omega, nsmooth);
  DevLevel &L = c->lv[level];
  upload_vec(c, level, L.f, f);
  upload_vec(c, level, L.x, x);
  lr_begin_epoch(c);
  emit_smoothing(c, level, sweep_list(kind, direction, nsmooth, true), true, omega, false, 0.0, false);
  normalize_x(c, level);
  download_vec(c, level, L.x, x);
  c->sync();
  check_device_error(c);
  API_END
}

int mgmc_coarse_solve(mgmc_ctx *c, const double *b, double *x) {
  API_BEGIN
  check_level(c, 0);
  const int lc = c->d.nlevel - 1;
  upload_vec(c, lc, c->lv[lc].f, b);
  dev_coarse(c, false, c->lv[lc].f, c->lv[lc].x);
  download_vec(c, lc, c->lv[lc].x, x);
  c->sync();
  API_END
}

int mgmc_coarse_sample(mgmc_ctx *c, const double *f, double *x) {
  API_BEGIN
  check_level(c, 0);
  const int lc = c->d.nlevel - 1;
  upload_vec(c, lc, c->lv[lc].f, f);
  dev_coarse(c, true, c->lv[lc].f, c->lv[lc].x);
  download_vec(c, lc, c->lv[lc].x, x);
  c->sync();
  API_END
}

int mgmc_sampler_mgmc_apply(mgmc_ctx *c, const double *f, double *x) {
  API_BEGIN
  check_level(c, 0);
  if (c->strip.on() && c->strip_connected) fail(MGMC_ERR_UNSUPPORTED, "row strips: the chain state is distributed; use mgmc_set_state / mgmc_sample / mgmc_get_state");
  plan_tail(c);
  plan_nza(c);
  if (f) upload_vec(c, 0, c->lv[0].f, f);  // f == NULL: right-hand side fixed earlier (Sampler::fix_rhs, sampler.hh:56)
  upload_vec(c, 0, c->lv[0].x, x);
  emit_mgmc_cycle(c);
  download_vec(c, 0, c->lv[0].x, x);
  set_sample_index(c, c->h_sample + 1);
  check_device_error(c);
  API_END
}

int mgmc_mgprec_apply(mgmc_ctx *c, const double *b, double *x) {
  API_BEGIN
  check_level(c, 0);
  plan_tail(c);
  upload_vec(c, 0, c->lv[0].f, b);
  mg_solve_level(c, 0);
  download_vec(c, 0, c->lv[0].x, x);
  c->sync();
  check_device_error(c);
  API_END
}

int mgmc_loop_solve(mgmc_ctx *c, const double *b, double *x, double rtol, double atol, int maxiter, double *history, int *nhist, int *niter,
                    int *converged) {
  API_BEGIN
  check_level(c, 0);
  if (c->d.nchains != 1) fail(MGMC_ERR_UNSUPPORTED, "mgmc_loop_solve works on a single right-hand side (nchains = 1)");
  if (maxiter < 0) fail(MGMC_ERR_INVALID, "maxiter must be >= 0");
  plan_tail(c);
  DevLevel &L = c->lv[0];
  const size_t total = (size_t)L.g.stride;
  const size_t origin = (size_t)GY * L.g.pitch + GX;
  if (!c->sol_x) {
    c->sol_x = c->dalloc<double>(total) + origin;
    c->sol_b = c->dalloc<double>(total) + origin;
    dim3 g = grid_sites(L.g, 1);
    c->npartial = (int)(g.x * g.y);
    c->d_partial = c->dalloc<double>(c->npartial);
    c->d_solver = c->dalloc<SolverCtl>(1);
  }
  if (maxiter + 1 > c->sol_hist_cap) {
    // (the captured iteration holds the pointer: a larger history needs a new graph)
    if (c->mg_graph) {
      cudaGraphExecDestroy(c->mg_graph);
      c->mg_graph = nullptr;
    }
    c->sol_hist_cap = std::max(maxiter + 1, 256);
    c->d_sol_hist = c->dalloc<double>(c->sol_hist_cap);
  }
  upload_vec(c, 0, c->sol_b, b);
  c->launch("zero", 0, [&] { axpy_kernel<1><<<grid_sites(L.g, 1), kBlockSites, 0, c->stream>>>(L.g, c->sol_x, nullptr); });
  double r0 = 0.0;
  const size_t n = (size_t)L.h.ndof();
  for (size_t k = 0; k < n; ++k) r0 += b[k] * b[k];
  r0 = std::sqrt(r0);
  SolverCtl ctl;
  std::memset(&ctl, 0, sizeof(ctl));
  ctl.r0 = r0;
  ctl.rtol = rtol;
  ctl.atol = atol;
  ctl.maxiter = maxiter;
  CUDA_CHECK(cudaMemcpyAsync(c->d_solver, &ctl, sizeof(ctl), cudaMemcpyHostToDevice, c->stream));
  c->sync();  // (ctl is a stack variable)
  // r = A x - b (into f_ell[0], the preconditioner's input), ||r|| into the device-side history + convergence test
  auto emit_residual = [&] {
    c->launch("residual_norm", 0, [&] {
      if (L.d3 && L.full27) residual_norm27_kernel<true><<<grid_sites(L.g, 1), kBlockSites, 0, c->stream>>>(L.g, L.q, L.c27, c->sol_x, c->sol_b, L.f, c->d_partial);
      else if (L.d3) residual_norm27_kernel<false><<<grid_sites(L.g, 1), kBlockSites, 0, c->stream>>>(L.g, L.q, L.c27, c->sol_x, c->sol_b, L.f, c->d_partial);
      else if (L.r2) residual_norm25_kernel<<<grid_sites(L.g, 1), kBlockSites, 0, c->stream>>>(L.g, L.d_st, c->sol_x, c->sol_b, L.f, c->d_partial);
      else if (L.vc && L.vc_full) residual_norm9v_kernel<true><<<grid_sites(L.g, 1), kBlockSites, 0, c->stream>>>(L.g, L.dvc, c->sol_x, c->sol_b, L.f, c->d_partial);
      else if (L.vc) residual_norm9v_kernel<false><<<grid_sites(L.g, 1), kBlockSites, 0, c->stream>>>(L.g, L.dvc, c->sol_x, c->sol_b, L.f, c->d_partial);
      else if (L.nine) residual_norm_kernel<true><<<grid_sites(L.g, 1), kBlockSites, 0, c->stream>>>(L.g, L.coef, c->sol_x, c->sol_b, L.f, c->d_partial);
      else residual_norm_kernel<false><<<grid_sites(L.g, 1), kBlockSites, 0, c->stream>>>(L.g, L.coef, c->sol_x, c->sol_b, L.f, c->d_partial);
    }, 24.0 * (double)L.h.ndof());
    if (c->d.m_lowrank > 0) {
      // low-rank part of A x is added to r, then the norm is recomputed from r
      if (L.lr_wide) dev_lowrank_wide_apply(c, 0, c->sol_x, L.B.rows, L.g.stride, L.f, c->d_sigma_inv, 1.0, 1);
      else
        c->launch("lowrank_apply", 0, [&] {
          lowrank_apply_kernel<<<1, 256, c->d.m_lowrank * sizeof(double), c->stream>>>(L.B.cols, L.B.rows, c->d_sigma_inv, L.g.stride, c->sol_x, L.f);
        });
      c->launch("norm", 0, [&] {
        residual_norm_kernel<false><<<grid_sites(L.g, 1), kBlockSites, 0, c->stream>>>(L.g, Coef9{0, 0, 0, 0, 0, 0, 0, 0, 0}, c->sol_x, L.f, L.r, c->d_partial);
      });
    }
    c->launch("reduce_check", 0, [&] { reduce_check_kernel<<<1, 1024, 0, c->stream>>>(c->d_partial, c->npartial, c->d_solver, c->d_sol_hist); });
  };
  // one iteration = V-cycle on r, x -= Pr (loop_solver.cc:40-41; suppressed once converged), residual of the new iterate:
  // captured once, replayed
  auto emit_iteration = [&] {
    mg_solve_level(c, 0);  // Pr = x_ell[0]
    c->launch("axpy", 0, [&] { axpy_guard_kernel<<<grid_sites(L.g, 1), kBlockSites, 0, c->stream>>>(L.g, c->sol_x, L.x, c->d_solver); });
    emit_residual();
  };
  if (c->use_graph && !c->mg_graph && !c->prof_on) {
    if (c->d.m_lowrank > 0)
      for (int l = 0; l < c->d.nlevel; ++l) get_lowrank(c, l, c->d.omega);
    ensure_coarse(c);
    c->sync();
    const int64_t count0 = c->launch_count;
    cudaGraph_t g = nullptr;
    CUDA_CHECK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    try {
      emit_iteration();
    } catch (...) {
      cudaStreamEndCapture(c->stream, &g);
      if (g) cudaGraphDestroy(g);
      throw;
    }
    CUDA_CHECK(cudaStreamEndCapture(c->stream, &g));
    CUDA_CHECK(cudaGraphInstantiate(&c->mg_graph, g, 0));
    CUDA_CHECK(cudaGraphDestroy(g));
    c->mg_graph_launches = c->launch_count - count0;
    c->launch_count = count0;
  }
  emit_residual();
  // The reference evaluates the residual, tests, updates -- at most maxiter times.  Here the iterations are launched in
  // batches without a host round trip in between; the batch size follows the observed convergence rate, so that only a
  // few iterations run past the one at which the test holds (they leave x alone).
  std::vector<double> hist(std::max(maxiter, 1), 0.0);
  int launched = 0, batch = std::min(4, maxiter);
  while (true) {
    for (int k = 0; k < batch; ++k) {
      if (c->mg_graph && !c->prof_on) {
        CUDA_CHECK(cudaGraphLaunch(c->mg_graph, c->stream));
        c->launch_count += c->mg_graph_launches;
      } else {
        emit_iteration();
      }
    }
    launched += batch;
    CUDA_CHECK(cudaMemcpyAsync(&ctl, c->d_solver, sizeof(ctl), cudaMemcpyDeviceToHost, c->stream));
    c->sync();
    if (ctl.conv || launched >= maxiter) break;
    CUDA_CHECK(cudaMemcpy(hist.data(), c->d_sol_hist, sizeof(double) * ctl.iter, cudaMemcpyDeviceToHost));
    int pred = 8;
    const int q = std::min(ctl.iter - 1, 4);
    if (q >= 1 && hist[ctl.iter - 1] > 0.0 && hist[ctl.iter - 1 - q] > hist[ctl.iter - 1]) {
      const double rho = std::pow(hist[ctl.iter - 1] / hist[ctl.iter - 1 - q], 1.0 / q);
      const double target = std::min(rtol * r0, atol);
      if (target > 0.0 && rho < 1.0) pred = (int)std::ceil(std::log(target / hist[ctl.iter - 1]) / std::log(rho));
    }
    batch = std::min(std::max(pred, 1), std::min(32, maxiter - launched));
  }
  const int nh = ctl.iter;
  if (history && nh > 0) CUDA_CHECK(cudaMemcpy(history, c->d_sol_hist, sizeof(double) * nh, cudaMemcpyDeviceToHost));
  download_vec(c, 0, c->sol_x, x);
  c->sync();
  check_device_error(c);
  if (nhist) *nhist = nh;
  if (niter) *niter = ctl.conv ? ctl.it_conv : maxiter;
  if (converged) *converged = ctl.conv ? 1 : 0;
  API_END
}

int mgmc_set_philox_position(mgmc_ctx *c, uint32_t sample, uint32_t sweep_counter) {
  API_BEGIN
  check_level(c, 0);
  std::fill(c->sweep_counter.begin(), c->sweep_counter.end(), sweep_counter);
  set_sample_index(c, sample);
  API_END
}

int mgmc_set_rhs(mgmc_ctx *c, const double *f) {
  API_BEGIN
  check_level(c, 0);
  upload_vec(c, 0, c->lv[0].f, f);
  c->sync();
  API_END
}
int mgmc_set_state(mgmc_ctx *c, const double *x) {
  API_BEGIN
  check_level(c, 0);
  upload_vec(c, 0, c->lv[0].x, x);
  c->sync();
  API_END
}
int mgmc_get_state(mgmc_ctx *c, double *x) {
  API_BEGIN
  check_level(c, 0);
  download_vec(c, 0, c->lv[0].x, x);
  c->sync();
  API_END
}

int mgmc_set_qoi(mgmc_ctx *c, int64_t nnz, const int64_t *idx, const double *val) {
  API_BEGIN
  check_level(c, 0);
  const DevLevel &L = c->lv[0];
  const int w = L.g.nx - 1;
  std::vector<long long> site(nnz);
  std::vector<double> v(val, val + nnz);
  for (int64_t e = 0; e < nnz; ++e) {
    if (idx[e] < 0 || idx[e] >= L.h.ndof()) fail(MGMC_ERR_INVALID, "QoI index out of range");
    int j = (int)(idx[e] / w) + 1;
    if (L.d3) {  // lexicographic 3d index (lattice3d.hh:122-135) -> row of the stacked planes
      const int r = (int)(idx[e] / w), h = L.q.ny - 1;
      j = (r / h + 1) * (L.q.ny + 1) + (r % h) + 1;
    }
    site[e] = (long long)j * L.g.pitch + (idx[e] % w + 1);
    // row strips: every rank sums the entries it owns; the caller adds the partial series of all ranks
    if (c->strip.on() && (j < c->strip.lo[0] || j > c->strip.hi[0])) v[e] = 0.0;
  }
  // (the drivers set the same functional before every series: nothing to do then -- no new buffers, no re-capture)
  if (c->d_qsite && site == c->h_qsite && v == c->h_qval) return MGMC_OK;
  c->sync();
  drop_graph(c);  // the captured cycle holds the old pointers
  c->dfree(c->d_qsite);
  c->dfree(c->d_qval);
  c->d_qsite = c->dupload(site);
  c->d_qval = c->dupload(v);
  c->h_qsite = site;
  c->h_qval = v;
  c->qoi_nnz = (int)nnz;
  API_END
}

int mgmc_sample(mgmc_ctx *c, int64_t nsamples, double *qoi_series) {
  API_BEGIN
  check_level(c, 0);
  if (nsamples < 0) fail(MGMC_ERR_INVALID, "nsamples must be >= 0");
  if (qoi_series) ensure_series(c, nsamples * c->d.nchains);
  run_cycles(c, nsamples);
  if (qoi_series && c->qoi_nnz > 0)
    CUDA_CHECK(cudaMemcpyAsync(qoi_series, c->d_series, sizeof(double) * nsamples * c->d.nchains, cudaMemcpyDeviceToHost, c->stream));
  c->sync();
  check_device_error(c);
  API_END
}

int mgmc_sample_timed(mgmc_ctx *c, int64_t nsamples, double *qoi_series, double *elapsed_ms) {
  API_BEGIN
  check_level(c, 0);
  if (nsamples < 0) fail(MGMC_ERR_INVALID, "nsamples must be >= 0");
  if (qoi_series) ensure_series(c, nsamples * c->d.nchains);
  if (c->use_graph && !c->graph) run_cycles(c, 0);  // instantiate the graph outside the timed region
  cudaEvent_t e0, e1;
  CUDA_CHECK(cudaEventCreate(&e0));
  CUDA_CHECK(cudaEventCreate(&e1));
  c->sync();
  CUDA_CHECK(cudaEventRecord(e0, c->stream));
  run_cycles(c, nsamples);
  CUDA_CHECK(cudaEventRecord(e1, c->stream));
  CUDA_CHECK(cudaEventSynchronize(e1));
  float ms = 0.f;
  CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (elapsed_ms) *elapsed_ms = ms;
  if (qoi_series && c->qoi_nnz > 0)
    CUDA_CHECK(cudaMemcpyAsync(qoi_series, c->d_series, sizeof(double) * nsamples * c->d.nchains, cudaMemcpyDeviceToHost, c->stream));
  c->sync();
  check_device_error(c);
  API_END
}

int mgmc_sample_moments(mgmc_ctx *c, int64_t nsamples, double *mean_field, double *second_moment_field) {
  API_BEGIN
  check_level(c, 0);
  DevLevel &L = c->lv[0];
  const size_t origin = (size_t)GY * L.g.pitch + GX;
  if (!c->d_mean) {
    c->d_mean = c->dalloc<double>((size_t)L.g.stride) + origin;
    c->d_second = c->dalloc<double>((size_t)L.g.stride) + origin;
  }
  CUDA_CHECK(cudaMemsetAsync(c->d_mean - origin, 0, sizeof(double) * L.g.stride, c->stream));
  CUDA_CHECK(cudaMemsetAsync(c->d_second - origin, 0, sizeof(double) * L.g.stride, c->stream));
  for (int64_t k = 0; k < nsamples; ++k) {
    run_cycles(c, 1);
    c->launch("moments", 0, [&] { moments_kernel<<<grid_sites(L.g, 1), kBlockSites, 0, c->stream>>>(L.g, L.x, c->d_mean, c->d_second, 1.0 / (k + 1.0)); });
  }
  if (L.d3) {
    if (c->d.nchains != 1) fail(MGMC_ERR_UNSUPPORTED, "3d lattices: moment fields need nchains = 1");
    download_vec(c, 0, c->d_mean, mean_field);
    download_vec(c, 0, c->d_second, second_moment_field);
  } else {
    const size_t w = L.g.nx - 1, h = L.g.ny - 1;
    CUDA_CHECK(cudaMemcpy2DAsync(mean_field, w * sizeof(double), c->d_mean + L.g.pitch + 1, L.g.pitch * sizeof(double), w * sizeof(double), h,
                                 cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaMemcpy2DAsync(second_moment_field, w * sizeof(double), c->d_second + L.g.pitch + 1, L.g.pitch * sizeof(double), w * sizeof(double), h,
                                 cudaMemcpyDeviceToHost, c->stream));
  }
  c->sync();
  check_device_error(c);
  API_END
}

int mgmc_strip_partition(const mgmc_desc *desc, int level, int rank, int *row_lo, int *row_hi, int *distributed) {
  API_BEGIN
  if (!desc) fail(MGMC_ERR_INVALID, "null argument");
  std::vector<HostLevel> H = build_host_levels(*desc);
  if (level < 0 || level >= (int)H.size()) fail(MGMC_ERR_INVALID, "level out of range");
  mgmc_desc d = *desc;
  d.strip_rank = rank;
  StripPlan p = make_strip_plan(d, H);
  const bool dist = p.on() && level < p.ndist;
  if (row_lo) *row_lo = dist ? p.lo[level] : 1;
  if (row_hi) *row_hi = dist ? p.hi[level] : H[level].ny - 1;
  if (distributed) *distributed = dist ? 1 : 0;
  API_END
}

int mgmc_plan_passes(int ncolours, int npass, const int *colours, int nfix, const int *fix_after, int omega_is_one, int restrict_behind, int lr_mx,
                     int lr_my, int *mode, int *margins, int *halo) {
  API_BEGIN
  if ((ncolours != 2 && ncolours != 4) || npass < 0 || npass > kMaxStages || !colours || !mode || !margins || !halo || nfix < 0 || (nfix > 0 && !fix_after))
    fail(MGMC_ERR_INVALID, "invalid argument");
  std::vector<Stage> st(npass);
  for (int s = 0; s < npass; ++s) {
    if (colours[s] < 0 || colours[s] >= ncolours) fail(MGMC_ERR_INVALID, "colour out of range");
    st[s] = Stage{colours[s], 0u, 0, 0, 0, 0, STAGE_FULL};
  }
  std::vector<FixSpec> fixes;
  for (int q = 0; q < nfix; ++q) fixes.push_back(FixSpec{fix_after[q], 0, 0u});
  const Margin in = plan_stages(ncolours, st, fixes, nfix > 0, omega_is_one != 0, restrict_behind != 0, lr_mx, lr_my);
  for (int s = 0; s < npass; ++s) {
    mode[s] = st[s].mode;
    margins[4 * s] = st[s].xl;
    margins[4 * s + 1] = st[s].xh;
    margins[4 * s + 2] = st[s].yl;
    margins[4 * s + 3] = st[s].yh;
  }
  for (int k = 0; k < 4; ++k) halo[k] = in.v[k];
  API_END
}

int mgmc_strip_handle_bytes(void) { return (int)sizeof(cudaIpcMemHandle_t); }

int mgmc_strip_export(mgmc_ctx *c, void *handle_out) {
  API_BEGIN
  check_level(c, 0);
  if (!c->strip.on()) fail(MGMC_ERR_INVALID, "context was not created with strip_nranks > 1");
  cudaIpcMemHandle_t h;
  CUDA_CHECK(cudaIpcGetMemHandle(&h, c->arena));
  std::memcpy(handle_out, &h, sizeof(h));
  API_END
}

int mgmc_strip_connect(mgmc_ctx *c, const void *all_handles) {
  API_BEGIN
  check_level(c, 0);
  if (!c->strip.on()) fail(MGMC_ERR_INVALID, "context was not created with strip_nranks > 1");
  CUDA_CHECK(cudaSetDevice(c->device));
  for (int r = 0; r < c->strip.nranks; ++r) {
    if (r == c->strip.rank || c->peer_arena[r]) continue;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, (const char *)all_handles + (size_t)r * sizeof(h), sizeof(h));
    void *p = nullptr;
    CUDA_CHECK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    c->peer_arena[r] = (char *)p;
  }
  c->strip_connected = true;
  drop_graph(c);
  API_END
}

int mgmc_strip_error(mgmc_ctx *c) {
  if (!c || !c->strip.on()) return 0;
  int e = 0;
  if (cudaMemcpy(&e, c->d_strip_ctl + 3, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return 1;
  if (e) cudaMemset(c->d_strip_ctl + 3, 0, sizeof(int));
  return e;
}

int64_t mgmc_launch_count(const mgmc_ctx *c) { return c ? c->launch_count : 0; }

int mgmc_profile_cycle(mgmc_ctx *c, int nsamples, int nslots_max, char *names, double *ms_total, int64_t *launches, double *alg_bytes, int *nslots) {
  API_BEGIN
  check_level(c, 0);
  c->sync();
  c->prof_on = true;
  c->prof_events.clear();
  try {
    run_cycles(c, nsamples);
    c->sync();
  } catch (...) {
    c->prof_on = false;
    throw;
  }
  c->prof_on = false;
  std::vector<ProfSlot> slots;
  std::map<std::string, int> index;
  for (auto &ev : c->prof_events) {
    float ms = 0.f;
    CUDA_CHECK(cudaEventElapsedTime(&ms, ev.e0, ev.e1));
    cudaEventDestroy(ev.e0);
    cudaEventDestroy(ev.e1);
    auto it = index.find(ev.name);
    if (it == index.end()) {
      it = index.emplace(ev.name, (int)slots.size()).first;
      slots.push_back(ProfSlot{ev.name, 0.0, 0, 0.0});
    }
    slots[it->second].ms += ms;
    slots[it->second].launches++;
    slots[it->second].bytes += ev.bytes;
  }
  c->prof_events.clear();
  const int n = std::min<int>((int)slots.size(), nslots_max);
  for (int k = 0; k < n; ++k) {
    std::snprintf(names + (size_t)k * 64, 64, "%s", slots[k].name.c_str());
    ms_total[k] = slots[k].ms;
    launches[k] = slots[k].launches;
    if (alg_bytes) alg_bytes[k] = slots[k].bytes;
  }
  if (nslots) *nslots = n;
  API_END
}

int mgmc_tail_stamps(mgmc_ctx *c, int nmax, int *kinds, double *us, int *nphases) {
  API_BEGIN
  check_level(c, 0);
  c->sync();
  const int n = (int)c->tail_stamp_kinds.size();
  if (!c->d_tail_stamps || n == 0) {
    if (nphases) *nphases = 0;
    return MGMC_OK;
  }
  std::vector<long long> st(n + 1);
  CUDA_CHECK(cudaMemcpy(st.data(), c->d_tail_stamps, sizeof(long long) * (n + 1), cudaMemcpyDeviceToHost));
  for (int k = 0; k < n && k < nmax; ++k) {
    kinds[k] = c->tail_stamp_kinds[k];
    us[k] = 1e-3 * (double)(st[k + 1] - st[k]);
  }
  if (nphases) *nphases = std::min(n, nmax);
  API_END
}

int mgmc_cycle_model(const mgmc_ctx *c, double *bytes, double *site_updates) {
  API_BEGIN
  check_level(c, 0);
  // 24 B/site per sweep, 18 B residual+restrict, 18 B prolongate_add, 2 B coarse
  // zeroing; per level visit.  Level l > 0 is visited cycle^l times.
  const mgmc_desc &d = c->d;
  const int per_smooth = (d.smoother == MGMC_SMOOTHER_SSOR) ? 2 : 1;
  double B = 0.0, U = 0.0, visits = 1.0;  // visits = number of passes through the loop body of a level
  for (int l = 0; l < d.nlevel - 1; ++l) {
    if (l > 0) visits *= d.cycle;  // multigridmc_sampler.cc:112
    const double n = (double)c->lv[l].h.ndof();
    const double sweeps = per_smooth * (d.npresmooth + d.npostsmooth);
    B += visits * n * (24.0 * sweeps + (c->lv[l].d3 ? 35.0 : 38.0));  // (3d: the coarse lattice has 1/8 of the sites)
    U += visits * n * sweeps;
  }
  const double nc = (double)c->lv[d.nlevel - 1].h.ndof();
  if (d.coarse_solver == MGMC_COARSE_CHOLESKY) {
    B += visits * 8.0 * nc * nc;
    U += visits * nc;
  } else {
    B += visits * nc * 24.0 * 2 * d.ncoarsesmooth;
    U += visits * nc * 2 * d.ncoarsesmooth;
  }
  if (bytes) *bytes = B;
  if (site_updates) *site_updates = U;
  API_END
}

}  // extern "C"
