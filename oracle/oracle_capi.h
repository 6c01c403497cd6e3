/* ORACLE -- TEST INFRASTRUCTURE ONLY.  C ABI of the CPU restatement (oracle_core.hh) so that
 * tests/ and bench.py's cpu_baseline leg can drive it through ctypes.  Nothing in the product
 * path (multigridmc_b200/, include/, host/) may include or link this. */
#ifndef MGMC_ORACLE_CAPI_H
#define MGMC_ORACLE_CAPI_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_op orc_op;         /* a LinearOperator (any level) */
typedef struct orc_hier orc_hier;     /* operator hierarchy + intergrid operators + orderings */
typedef struct orc_rng orc_rng;       /* std::mt19937_64 / std::mt19937 + persistent distributions */
typedef struct orc_obj orc_obj;       /* smoother / sampler / preconditioner object */

const char *orc_last_error(void);

/* ---- std RNG streams used by the reference's drivers and test fixtures ---- */
orc_rng *orc_rng_create(int bits /*32|64*/, uint64_t seed);
void orc_rng_destroy(orc_rng *);
int orc_rng_normal(orc_rng *, long n, double *out);  /* std::normal_distribution<double>(0,1), one persistent object */
int orc_rng_uniform(orc_rng *, long n, double *out); /* std::uniform_real_distribution<double>(0,1) */

/* ---- lattice index algebra (lattice/lattice{1d,2d,3d}.hh) ---- */
long orc_lattice_nvertex(int dim, const int *n);
long orc_lattice_ncell(int dim, const int *n);
long orc_lattice_vertex_e2l(int dim, const int *n, const int *idx);
int orc_lattice_vertex_l2e(int dim, const int *n, long ell, int *idx);
long orc_lattice_cell_e2l(int dim, const int *n, const int *idx);
int orc_lattice_cell_l2e(int dim, const int *n, long ell, int *idx);
long orc_lattice_shift_vertexidx(int dim, const int *n, long ell, const int *shift); /* index arithmetic, unchecked */
long orc_lattice_shifted_vertex_internal(int dim, const int *n, long ell, const int *shift); /* -1 if not interior */
long orc_lattice_shift_cellidx(int dim, const int *n, long ell, const int *shift);
long orc_lattice_corner_vertex(int dim, const int *n, long cell, const int *corner); /* -1 if not interior */
long orc_lattice_fine_vertex_idx(int dim, const int *n, long ell);
int orc_lattice_vertex_coordinates(int dim, const int *n, long ell, double *x);
int orc_lattice_coarsen(int dim, const int *n, int *n_coarse); /* nonzero status on the reference's exit(-1) paths */

/* ---- operators ---- */
/* pde: 0 shiftedlaplace_fd, 1 squared_shiftedlaplace_fd, 2 shiftedlaplace_fem, 3 TestOperator1d (prior part)
 * kappa_model: 0 constant (lam0 = Lambda), 1 periodic (lam0 = Lambda_min, lam1 = Lambda_max) */
orc_op *orc_op_create_prior(int dim, const int *n, int pde, int kappa_model, double lam0, double lam1);
/* MeasuredOperator(base, params): variance is the UN-scaled vector, scaled inside by variance_scaling */
orc_op *orc_op_create_measured(const orc_op *base, int n_meas, const double *locations, const double *variance, double variance_scaling,
                               double radius, int measure_global, double variance_global);
orc_op *orc_op_create_test1d(int lowrank);
void orc_op_destroy(orc_op *);
long orc_op_ndof(const orc_op *);
int orc_op_m_lowrank(const orc_op *);
int orc_op_lattice(const orc_op *, int *dim, int *n);
int orc_op_apply(const orc_op *, const double *x, double *y);
long orc_op_nnz(const orc_op *);
int orc_op_get_csr(const orc_op *, long *rowptr, int *col, double *val);
long orc_op_B_nnz(const orc_op *);
int orc_op_get_B(const orc_op *, long *rows, int *cols, double *vals, double *sigma); /* COO triplets of B + diag(Sigma) */
int orc_op_precision(const orc_op *, double *dense_rowmajor);
int orc_op_covariance(const orc_op *, double *dense_rowmajor);
int orc_op_mean(const orc_op *, const double *xbar, const double *y, double *out);
int orc_op_observed_mean_and_variance(const orc_op *, const double *xbar, const double *y, const double *b_obs, double *mean, double *variance);
int orc_measurement_vector(const orc_op *, const double *x0, double radius, double *dense_out);

/* ---- hierarchy ---- */
orc_hier *orc_hier_create(const orc_op *fine, int nlevel, int ordering /*0 lexicographic (reference), 1 multicolour*/);
void orc_hier_destroy(orc_hier *);
const orc_op *orc_hier_op(const orc_hier *, int level); /* borrowed */
int orc_hier_ncolours(const orc_hier *, int level);
int orc_hier_order(const orc_hier *, int level, long *order);
int orc_hier_restrict(const orc_hier *, int level, const double *x_fine, double *x_coarse);
int orc_hier_prolongate_add(const orc_hier *, int level, double alpha, const double *x_coarse, double *x_fine);

/* ---- parameters (auxilliary/parameters.hh:145-174) ---- */
typedef struct {
  int nlevel;
  int smoother;      /* 0 "SOR", 1 "SSOR" */
  int coarse_solver; /* 0 "SSOR", 1 "Cholesky" */
  int npresmooth, npostsmooth, ncoarsesmooth;
  int cycle;
  double coarse_scaling;
  double omega;
} orc_mg_params;

/* ---- smoothers / samplers / solvers; level refers to the hierarchy ---- */
/* kind: 0 SOR, 1 SSOR;  direction: 1 forward, 2 backward (ignored for SSOR) */
orc_obj *orc_smoother_create(const orc_hier *, int level, int kind, double omega, int nsmooth, int direction);
/* noise: rng != NULL -> reference mode (shared engine); rng == NULL -> philox(seed, chain) */
orc_obj *orc_sampler_create(const orc_hier *, int level, int kind /*0 SOR,1 SSOR,2 Cholesky*/, double omega, int nsmooth, int direction,
                            orc_rng *rng, uint64_t philox_seed);
orc_obj *orc_mgmc_create(const orc_hier *, const orc_mg_params *, orc_rng *rng, uint64_t philox_seed);
orc_obj *orc_mgprec_create(const orc_hier *, const orc_mg_params *);
orc_obj *orc_cholesky_solver_create(const orc_hier *, int level);
void orc_obj_destroy(orc_obj *);
/* apply(b_or_f, x): x is in/out for smoothers and samplers, out for preconditioner / Cholesky solver */
int orc_obj_apply(orc_obj *, const double *b, double *x);
/* philox-mode samplers only: set (sample index, chain id) and reset / read the per-level sweep counters */
int orc_obj_set_philox_position(orc_obj *, uint32_t sample, uint32_t chain, uint32_t sweep_counter);
/* run nsamples of sampler->apply(f,x) recording z_k = dot(b_obs, x) (driver_mgmc.cc:73-77) */
int orc_sampler_run(orc_obj *, const double *f, double *x, const double *b_obs, long nsamples, double *series);
/* mean_covariance_error harness (sampler/test_sampler.hh:113-153): accumulates E[x], E[x x^T] */
int orc_sampler_moments(orc_obj *, const double *f, double *x, long nwarmup, long nsamples, double *Ex, double *Exx);

/* LoopSolver (solver/loop_solver.cc:9-53): history must hold maxiter doubles */
int orc_loop_solve(const orc_op *, orc_obj *prec, double rtol, double atol, int maxiter, int verbose, const double *b, double *x,
                   double *history, int *nhist, int *niter, int *converged);

/* ---- statistics (auxilliary/statistics.cc:65-79) ---- */
double orc_tau_int(const double *series, long n, int window);

/* ---- philox noise (counter layout in oracle_core.hh) ---- */
void orc_philox_normal_pair(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, double *z0, double *z1);
void orc_philox_raw(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t *out4);

#ifdef __cplusplus
}
#endif
#endif
