/* mgmc_b200.h -- C ABI of the B200-native MultigridMC hot path.
 *
 * Drop-in boundary for the sampling / solve path of nilsfriess/MultigridMC.
 * Plain pointers and sizes only; every host vector is in the reference's LEXICOGRAPHIC
 * interior-vertex order (lattice/lattice2d.hh:96-103); the padded device layout and the multicolour
 * sweep ordering are internal.  All entry points return 0 on success and a negative MGMC_ERR_* code
 * otherwise (mgmc_last_error() gives the message); nothing here ever calls exit().
 *
 * Each entry point names the reference interface it replaces (paths relative to
 * /root/reference/src).  The C++ classes in host/ wrap these calls under the reference's own class
 * names; INTEGRATION.md shows the binding a reference maintainer would add.
 */
#ifndef MGMC_B200_H
#define MGMC_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct mgmc_ctx mgmc_ctx;

enum {
  MGMC_OK = 0,
  MGMC_ERR_INVALID = -1,     /* invalid argument / configuration (reference: message + exit(-1)) */
  MGMC_ERR_UNSUPPORTED = -2, /* valid reference configuration that the device path does not cover yet */
  MGMC_ERR_CUDA = -3,        /* CUDA runtime error */
  MGMC_ERR_NOTCONVERGED = -4
};

/* shiftedlaplace_fd_operator.cc, squared_shiftedlaplace_fd_operator.cc, shiftedlaplace_fem_operator.cc (the FEM operator with a constant
 * correlation length: a uniform 9-point (2d) / 27-point (3d) stencil on the fine level as well) */
enum { MGMC_PDE_SHIFTEDLAPLACE_FD = 0, MGMC_PDE_SQUARED_SHIFTEDLAPLACE_FD = 1, MGMC_PDE_SHIFTEDLAPLACE_FEM = 2 };
enum { MGMC_SMOOTHER_SOR = 0, MGMC_SMOOTHER_SSOR = 1 };
enum { MGMC_COARSE_SSOR = 0, MGMC_COARSE_CHOLESKY = 1 };
enum { MGMC_FORWARD = 1, MGMC_BACKWARD = 2 }; /* smoother/sor_smoother.hh:50-54 */
enum { MGMC_VEC_X = 0, MGMC_VEC_F = 1, MGMC_VEC_R = 2 };

/* Problem + hierarchy description.  Mirrors LatticeParameters, PriorParameters,
 * ConstantCorrelationLengthModelParameters, MultigridParameters (auxilliary/parameters.hh:94-277)
 * and the host-assembled low-rank part of MeasuredOperator (linear_operator/measured_operator.cc:9-49). */
typedef struct {
  int dim;                /* 2 (Lattice2d, lattice/lattice2d.hh) or 3 (Lattice3d, lattice/lattice3d.hh: shiftedlaplace_fd with a
                           * constant correlation length and no measurements; host vectors in the reference's order
                           * ell = (k-1)(nx-1)(ny-1) + (j-1)(nx-1) + (i-1), lattice3d.hh:122-135); 1 is MGMC_ERR_UNSUPPORTED */
  int nx, ny, nz;         /* cells per direction (nz: dim = 3 only) */
  int pde_model;          /* MGMC_PDE_* */
  double Lambda;          /* constant correlation length: kappa^2 = 1/Lambda^2 */
  /* low-rank term B Sigma^{-1} B^T: B as COO triplets (row = lexicographic vertex index) */
  int m_lowrank;
  int64_t B_nnz;
  const int64_t *B_rows;
  const int32_t *B_cols;
  const double *B_vals;
  const double *Sigma;    /* m_lowrank diagonal entries (already multiplied by variance_scaling) */
  /* MultigridParameters (auxilliary/parameters.hh:145-174) */
  int nlevel;
  int smoother;           /* MGMC_SMOOTHER_* */
  int coarse_solver;      /* MGMC_COARSE_* (the preconditioner always uses Cholesky, multigrid_preconditioner.cc:41-45) */
  int npresmooth, npostsmooth, ncoarsesmooth;
  int cycle;              /* 1 = V, 2 = W (below level 0 only, multigridmc_sampler.cc:112) */
  double coarse_scaling;
  double omega;
  /* noise + placement */
  uint64_t seed;          /* Philox key (the reference's std::mt19937_64 seed plays this role) */
  int device;             /* CUDA device ordinal */
  int nchains;            /* independent Markov chains advanced together (>= 1) */
  int first_chain;        /* global id of chain 0 of this context (Philox counter word 3) */
  /* row-strip domain decomposition of ONE lattice over the GPUs of a node (one process per GPU):
   * rank strip_rank of strip_nranks owns a contiguous block of rows of every distributed level; the
   * coarse levels are replicated.  0 / 1 ranks = off.  See mgmc_strip_* below. */
  int strip_rank, strip_nranks;
  /* Correlation length that varies in space (CorrelationLengthModel::kappa_sq, linear_operator/
   * correlationlength_model.hh:45-113; PeriodicCorrelationLengthModel :83-113): kappa^2 at every interior vertex of the
   * finest lattice, lexicographic, (nx-1)*(ny-1) entries, as ShiftedLaplaceFDOperator evaluates it
   * (shiftedlaplace_fd_operator.cc:35-36).  NULL: constant, 1 / Lambda^2.  shiftedlaplace_fd only; the operators of all
   * levels then carry per-vertex coefficients (Galerkin products R A R^T on the host, csrc/varcoef.cuh on the device). */
  const double *kappa_sq;
} mgmc_desc;

const char *mgmc_last_error(void);

/* MultigridMCSampler::MultigridMCSampler / MultigridPreconditioner ctor: builds lattice hierarchy,
 * Galerkin stencils (LinearOperator::coarsen, linear_operator.cc:10-23), low-rank smoother data
 * (SORSmoother ctor, sor_smoother.cc:9-39) and the coarse dense factor (cholesky_sampler.cc:25-38). */
int mgmc_create(const mgmc_desc *desc, mgmc_ctx **out);
void mgmc_destroy(mgmc_ctx *);

/* level geometry: cells, unknowns, colours used by the sweeps (2: red-black, 4, 9) */
int mgmc_level_info(const mgmc_ctx *, int level, int *nx, int *ny, int64_t *ndof, int *ncolours);
/* cells in z of a level of a 3d hierarchy (Lattice3d::get_coarse_lattice, lattice3d.hh:242-257); 0 on 2d lattices.
 * 3d levels sweep in 2 colours ((i + j + k) & 1, 7-point fine operator) or 8 ((i & 1) + 2 (j & 1) + 4 (k & 1), 27-point). */
int mgmc_level_nz(const mgmc_ctx *, int level, int *nz);
/* Galerkin stencil of a level: 9 position classes x 25 coefficients, class = cx + 3*cy with
 * cx,cy in {0: first interior line, 1: interior, 2: last interior line}; coefficient (di,dj) at
 * [(dj+2)*5 + (di+2)].  Lets tests compare against the oracle's R A R^T (linear_operator.cc:12-15). */
int mgmc_get_stencil(const mgmc_ctx *, int level, double *out225);
/* same algebra run on the host only (needs no CUDA device; desc->B_* may be empty) */
int mgmc_host_stencil(const mgmc_desc *desc, int level, double *out225, int *ncolours);
/* 3d twin: the uniform radius-1 stencil of `level` (7-point fine operator, shiftedlaplace_fd_operator.cc:33-56; 27-point
 * Galerkin products R A R^T with the trilinear full weighting, linear_operator.cc:12-15), coefficient (di, dj, dk) at
 * [(dk+1)*9 + (dj+1)*3 + (di+1)].  Host only. */
int mgmc_host_stencil3(const mgmc_desc *desc, int level, double *out27, int *ncolours);
/* Operators with per-vertex coefficients (desc->kappa_sq != NULL): the matrix of `level` as nine planes, entry
 * A[(i, j), (i + di, j + dj)] at out[((dj + 1) * 3 + (di + 1)) * (nx_l + 1) * (ny_l + 1) + j * (nx_l + 1) + i] with (i, j)
 * the Euclidean vertex index of that level (zero on the boundary).  Host only; lets the CPU tests compare with the
 * oracle's R A R^T (linear_operator.cc:12-15). */
int mgmc_host_coefficients(const mgmc_desc *desc, int level, double *out, int *ncolours);

/* ---- single-level operations on HOST vectors (lexicographic, length ndof(level) * nchains) ---- */
/* LinearOperator::apply (linear_operator.hh:66-76): y = A_0 x + B Sigma^{-1} B^T x */
int mgmc_op_apply(mgmc_ctx *, int level, const double *x, double *y);
/* IntergridOperator::restrict (intergrid_operator.hh:74-88) */
int mgmc_restrict(mgmc_ctx *, int level, const double *x_fine, double *x_coarse);
/* IntergridOperator::prolongate_add (intergrid_operator.hh:106-120): x += alpha R^T x_coarse */
int mgmc_prolongate_add(mgmc_ctx *, int level, double alpha, const double *x_coarse, double *x_fine);
/* fused r = f - A x, f_coarse = R r (multigridmc_sampler.cc:118-120) */
int mgmc_residual_restrict(mgmc_ctx *, int level, const double *f, const double *x, double *f_coarse);
/* SORSmoother::apply / SSORSmoother::apply (sor_smoother.cc:41-53, ssor_smoother.cc:9-16) in the
 * multicolour ordering; x is in/out */
int mgmc_smoother_apply(mgmc_ctx *, int level, int kind, int direction, double omega, int nsmooth, const double *b, double *x);
/* SORSampler::apply / SSORSampler::apply (sor_sampler.cc:37-58, ssor_sampler.cc:9-16), Philox noise;
 * x is in/out.  Uses and advances the context's Philox position (sample index, sweep counters). */
int mgmc_sampler_apply(mgmc_ctx *, int level, int kind, int direction, double omega, int nsmooth, const double *f, double *x);
/* CholeskySolver::apply (cholesky_solver.cc:30-41) and CholeskySampler::apply
 * (cholesky_sampler.hh:50-66) on the coarsest level */
int mgmc_coarse_solve(mgmc_ctx *, const double *b, double *x);
int mgmc_coarse_sample(mgmc_ctx *, const double *f, double *x);

/* ---- multilevel operations on HOST vectors ---- */
/* MultigridMCSampler::apply(f, x) (multigridmc_sampler.cc:133-138), x in/out = chain state.
 * f == NULL keeps the right-hand side set by mgmc_set_rhs (Sampler::fix_rhs, sampler.hh:56). */
int mgmc_sampler_mgmc_apply(mgmc_ctx *, const double *f, double *x);
/* MultigridPreconditioner::apply(b, x) (multigrid_preconditioner.cc:104-108) */
int mgmc_mgprec_apply(mgmc_ctx *, const double *b, double *x);
/* LoopSolver::apply(b, x) (loop_solver.cc:9-53).  history receives ||r_k||, one entry per evaluated
 * residual (at most maxiter); stop iff rel < rtol AND abs < atol, as the reference. */
int mgmc_loop_solve(mgmc_ctx *, const double *b, double *x, double rtol, double atol, int maxiter, double *history, int *nhist,
                    int *niter, int *converged);

/* ---- Philox position (sample index / per-level sweep counters); see DESIGN.md "Noise" ---- */
int mgmc_set_philox_position(mgmc_ctx *, uint32_t sample, uint32_t sweep_counter);

/* ---- device-resident sampling: the hot loop of measure_sampling_time (driver_mgmc.cc:66-77) ---- */
int mgmc_set_rhs(mgmc_ctx *, const double *f);   /* f_ell[0] <- f      (H2D) */
int mgmc_set_state(mgmc_ctx *, const double *x); /* x_ell[0] <- x      (H2D) */
int mgmc_get_state(mgmc_ctx *, double *x);       /* x <- x_ell[0]      (D2H) */
/* observation functional z = sample_vector . x (driver_mgmc.cc:58,76), sparse, lexicographic indices */
int mgmc_set_qoi(mgmc_ctx *, int64_t nnz, const int64_t *idx, const double *val);
/* advance every chain by nsamples MGMC cycles; qoi_series (nullable) receives nsamples * nchains
 * values, sample-major.  Synchronous at return. */
int mgmc_sample(mgmc_ctx *, int64_t nsamples, double *qoi_series);
/* same, additionally accumulating running mean / second moment fields of chain 0
 * (posterior_statistics, driver_mgmc.cc:146-151); fields are lexicographic, length ndof(0) */
int mgmc_sample_moments(mgmc_ctx *, int64_t nsamples, double *mean_field, double *second_moment_field);
/* like mgmc_sample but timed on the device with CUDA events on the launching stream */
int mgmc_sample_timed(mgmc_ctx *, int64_t nsamples, double *qoi_series, double *elapsed_ms);

/* ---- row-strip decomposition over several GPUs (no reference counterpart: the
 *      reference is single-threaded) ----
 * Every rank creates its context with the same desc except device / strip_rank.  The ranks then exchange
 * the opaque handle blobs (e.g. torch.distributed.all_gather) and connect; afterwards mgmc_sample /
 * mgmc_sample_timed advance ONE chain cooperatively: per fused launch the boundary rows of the iterate are
 * stored straight into the neighbours' memory over NVLink (CUDA IPC mappings) and a flag is raised; the
 * neighbour's next launch waits on the flag on the device.  The chain is bit-identical to the single-GPU
 * chain (the noise of a site does not depend on the decomposition). */
/* rows [row_lo, row_hi] of `level` owned by `rank` (host-only, needs no device); *distributed = 0 for
 * the replicated coarse levels (every rank owns all rows) */
int mgmc_strip_partition(const mgmc_desc *desc, int level, int rank, int *row_lo, int *row_hi, int *distributed);
int mgmc_strip_handle_bytes(void);
int mgmc_strip_export(mgmc_ctx *, void *handle_out);
/* all_handles: strip_nranks blobs of mgmc_strip_handle_bytes() bytes, ordered by rank */
int mgmc_strip_connect(mgmc_ctx *, const void *all_handles);
/* nonzero if a device-side wait for a neighbour timed out since the last call (then the state is invalid) */
int mgmc_strip_error(mgmc_ctx *);

/* ---- introspection (host-only, needs no device; used by the CPU tests) ---- */
/* Plan of the colour passes of one fused launch (DESIGN.md 4.1 "Pass planning"): for the colour sequence
 * colours[0..npass) of a level with `ncolours` colours (2: red-black, 4: 4-colour ordering), fix-ups after the passes
 * fix_after[0..nfix), omega == 1 or not, a fused residual + restriction behind or not, and the extent (lr_mx, lr_my) of
 * supp(B_k) beyond its lower left corner (0, 0 without a low-rank term), returns per pass mode[s] (0 = full rectangle,
 * 1 = skipped, 2 = only supp(B_k) of the owned measurements) and the margins margins[4 s .. 4 s + 3] = (xl, xh, yl, yh)
 * of the updated rectangle around the tile, and in halo[0..3] the margins of the input the launch loads. */
int mgmc_plan_passes(int ncolours, int npass, const int *colours, int nfix, const int *fix_after, int omega_is_one, int restrict_behind, int lr_mx,
                     int lr_my, int *mode, int *margins, int *halo);

/* ---- instrumentation ---- */
/* number of kernels launched by this context so far */
int64_t mgmc_launch_count(const mgmc_ctx *);
/* per-kernel CUDA-event timing of nsamples MGMC cycles: for each distinct (kernel, level) slot of
 * one cycle returns total ms, launch count and (nullable) the algorithmic bytes of those launches
 * in the model of mgmc_cycle_model.  names: caller buffer of nslots_max * 64 chars. */
int mgmc_profile_cycle(mgmc_ctx *, int nsamples, int nslots_max, char *names, double *ms_total, int64_t *launches, double *alg_bytes,
                       int *nslots);
/* Levels below ~512 x 512 and the coarsest-level solve run as phases of ONE persistent cooperative launch
 * (csrc/tail.cuh).  With MGMC_TAIL_STAMPS=1 in the environment the kernel records a time stamp per phase; this
 * returns the duration (us) of every phase of the most recent such launch and its kind: 100 + tiles = smoothing +
 * residual + restriction, 200 + tiles = prolongation + smoothing, -1 = coarse solve, -2 / -3 = copy / zero. */
int mgmc_tail_stamps(mgmc_ctx *, int nmax, int *kinds, double *us, int *nphases);
/* algorithmic bytes and site updates of one MGMC cycle */
int mgmc_cycle_model(const mgmc_ctx *, double *bytes, double *site_updates);

#ifdef __cplusplus
}
#endif
#endif
