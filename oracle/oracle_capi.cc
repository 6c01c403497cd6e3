// ORACLE -- TEST INFRASTRUCTURE ONLY.  C ABI wrappers around oracle_core.hh (see oracle_capi.h).
#include "oracle_capi.h"

#include <cstring>

#include "oracle_core.hh"

using namespace orc;

static thread_local std::string g_err;
const char *orc_last_error(void) { return g_err.c_str(); }

#define ORC_TRY try {
#define ORC_CATCH(ret)                  \
  }                                     \
  catch (const std::exception &e) {     \
    g_err = e.what();                   \
    return ret;                         \
  }

struct orc_op {
  std::shared_ptr<LinearOperator> p;
};
struct orc_hier {
  std::shared_ptr<Hierarchy> H;
  std::vector<orc_op> views;
};
struct orc_rng {
  int bits;
  std::mt19937_64 e64;
  std::mt19937 e32;
  std::normal_distribution<double> normal{0.0, 1.0};
  std::uniform_real_distribution<double> uniform{0.0, 1.0};
};
struct orc_obj {
  std::shared_ptr<Hierarchy> H;
  std::shared_ptr<PhiloxCtx> px;
  std::shared_ptr<SORSmoother> sor_smoother;
  std::shared_ptr<SSORSmoother> ssor_smoother;
  std::shared_ptr<SORSampler> sor_sampler;
  std::shared_ptr<SSORSampler> ssor_sampler;
  std::shared_ptr<CholeskySampler> chol_sampler;
  std::shared_ptr<CholeskySolver> chol_solver;
  std::shared_ptr<MultigridMCSampler> mgmc;
  std::shared_ptr<MultigridPreconditioner> mgprec;
  bool is_sampler() const { return sor_sampler || ssor_sampler || chol_sampler || mgmc; }
  void sample(const double *f, double *x) {
    if (sor_sampler) sor_sampler->apply(f, x);
    else if (ssor_sampler) ssor_sampler->apply(f, x);
    else if (chol_sampler) chol_sampler->apply(f, x);
    else if (mgmc) mgmc->apply(f, x);
    else throw std::runtime_error("object is not a sampler");
    if (px && !mgmc) {  // standalone philox samplers: one "sample" per apply
      px->sample++;
      std::fill(px->sweep_counter.begin(), px->sweep_counter.end(), 0u);
    }
  }
};

extern "C" {

orc_rng *orc_rng_create(int bits, uint64_t seed) {
  orc_rng *r = new orc_rng();
  r->bits = bits;
  r->e64.seed(seed);
  r->e32.seed((uint32_t)seed);
  return r;
}
void orc_rng_destroy(orc_rng *r) { delete r; }
int orc_rng_normal(orc_rng *r, long n, double *out) {
  for (long k = 0; k < n; ++k) out[k] = (r->bits == 64) ? r->normal(r->e64) : r->normal(r->e32);
  return 0;
}
int orc_rng_uniform(orc_rng *r, long n, double *out) {
  for (long k = 0; k < n; ++k) out[k] = (r->bits == 64) ? r->uniform(r->e64) : r->uniform(r->e32);
  return 0;
}

long orc_lattice_nvertex(int dim, const int *n) { return Lattice(dim, n).Nvertex(); }
long orc_lattice_ncell(int dim, const int *n) { return Lattice(dim, n).Ncell(); }
long orc_lattice_vertex_e2l(int dim, const int *n, const int *idx) { return Lattice(dim, n).vertex_e2l(idx); }
int orc_lattice_vertex_l2e(int dim, const int *n, long ell, int *idx) {
  Lattice(dim, n).vertex_l2e(ell, idx);
  return 0;
}
long orc_lattice_cell_e2l(int dim, const int *n, const int *idx) { return Lattice(dim, n).cell_e2l(idx); }
int orc_lattice_cell_l2e(int dim, const int *n, long ell, int *idx) {
  Lattice(dim, n).cell_l2e(ell, idx);
  return 0;
}
long orc_lattice_shift_vertexidx(int dim, const int *n, long ell, const int *shift) { return Lattice(dim, n).shift_vertexidx(ell, shift); }
long orc_lattice_shifted_vertex_internal(int dim, const int *n, long ell, const int *shift) {
  long out = -1;
  return Lattice(dim, n).shifted_vertex_is_internal(ell, shift, out) ? out : -1;
}
long orc_lattice_shift_cellidx(int dim, const int *n, long ell, const int *shift) { return Lattice(dim, n).shift_cellidx(ell, shift); }
long orc_lattice_corner_vertex(int dim, const int *n, long cell, const int *corner) {
  long out = -1;
  return Lattice(dim, n).corner_is_internal_vertex(cell, corner, out) ? out : -1;
}
long orc_lattice_fine_vertex_idx(int dim, const int *n, long ell) { return Lattice(dim, n).fine_vertex_idx(ell); }
int orc_lattice_vertex_coordinates(int dim, const int *n, long ell, double *x) {
  Lattice(dim, n).vertex_coordinates(ell, x);
  return 0;
}
int orc_lattice_coarsen(int dim, const int *n, int *nc) {
  ORC_TRY
  Lattice c = Lattice(dim, n).coarse();
  for (int d = 0; d < dim; ++d) nc[d] = c.n[d];
  return 0;
  ORC_CATCH(1)
}

orc_op *orc_op_create_prior(int dim, const int *n, int pde, int kappa_model, double lam0, double lam1) {
  ORC_TRY
  Lattice lat(dim, n);
  KappaModel km;
  km.kind = kappa_model;
  km.Lambda = lam0;
  km.Lambda_min = lam0;
  km.Lambda_max = lam1;
  orc_op *o = new orc_op();
  if (pde == 0) o->p = std::make_shared<LinearOperator>(make_shiftedlaplace_fd(lat, km));
  else if (pde == 1) o->p = std::make_shared<LinearOperator>(make_squared_shiftedlaplace_fd(lat, km));
  else if (pde == 2) o->p = std::make_shared<LinearOperator>(make_shiftedlaplace_fem(lat, km));
  else {
    delete o;
    throw std::runtime_error("invalid pde model");
  }
  return o;
  ORC_CATCH(nullptr)
}
orc_op *orc_op_create_measured(const orc_op *base, int n_meas, const double *locations, const double *variance, double variance_scaling,
                               double radius, int measure_global, double variance_global) {
  ORC_TRY
  std::vector<double> v(n_meas);
  for (int k = 0; k < n_meas; ++k) v[k] = variance_scaling * variance[k];  // measured_operator.cc:19
  orc_op *o = new orc_op();
  o->p = std::make_shared<LinearOperator>(make_measured_operator(*base->p, n_meas, locations, v.data(), radius, measure_global != 0, variance_global));
  return o;
  ORC_CATCH(nullptr)
}
orc_op *orc_op_create_test1d(int lowrank) {
  orc_op *o = new orc_op();
  o->p = std::make_shared<LinearOperator>(make_test_operator_1d(lowrank != 0));
  return o;
}
void orc_op_destroy(orc_op *o) { delete o; }
long orc_op_ndof(const orc_op *o) { return o->p->ndof(); }
int orc_op_m_lowrank(const orc_op *o) { return o->p->m_lowrank; }
int orc_op_lattice(const orc_op *o, int *dim, int *n) {
  *dim = o->p->lattice.dim;
  for (int d = 0; d < 3; ++d) n[d] = o->p->lattice.n[d];
  return 0;
}
int orc_op_apply(const orc_op *o, const double *x, double *y) {
  o->p->apply(x, y);
  return 0;
}
long orc_op_nnz(const orc_op *o) { return o->p->A.nnz(); }
int orc_op_get_csr(const orc_op *o, long *rowptr, int *col, double *val) {
  const CSR &A = o->p->A;
  std::memcpy(rowptr, A.rowptr.data(), sizeof(long) * (A.rows + 1));
  std::memcpy(col, A.col.data(), sizeof(int) * A.nnz());
  std::memcpy(val, A.val.data(), sizeof(double) * A.nnz());
  return 0;
}
long orc_op_B_nnz(const orc_op *o) { return o->p->m_lowrank > 0 ? o->p->B.nnz() : 0; }
int orc_op_get_B(const orc_op *o, long *rows, int *cols, double *vals, double *sigma) {
  const LinearOperator &op = *o->p;
  if (op.m_lowrank == 0) return 0;
  long q = 0;
  for (long r = 0; r < op.B.rows; ++r)
    for (long k = op.B.rowptr[r]; k < op.B.rowptr[r + 1]; ++k, ++q) {
      rows[q] = r;
      cols[q] = op.B.col[k];
      vals[q] = op.B.val[k];
    }
  for (int k = 0; k < op.m_lowrank; ++k) sigma[k] = op.Sigma[k];
  return 0;
}
int orc_op_precision(const orc_op *o, double *D) {
  ORC_TRY
  std::vector<double> Q = o->p->precision();
  std::memcpy(D, Q.data(), sizeof(double) * Q.size());
  return 0;
  ORC_CATCH(1)
}
int orc_op_covariance(const orc_op *o, double *D) {
  ORC_TRY
  std::vector<double> Q = o->p->covariance();
  std::memcpy(D, Q.data(), sizeof(double) * Q.size());
  return 0;
  ORC_CATCH(1)
}
int orc_op_mean(const orc_op *o, const double *xbar, const double *y, double *out) {
  ORC_TRY
  Vec m = o->p->mean(xbar, y);
  std::memcpy(out, m.data(), sizeof(double) * m.size());
  return 0;
  ORC_CATCH(1)
}
int orc_op_observed_mean_and_variance(const orc_op *o, const double *xbar, const double *y, const double *b_obs, double *mean, double *variance) {
  ORC_TRY
  o->p->observed_mean_and_variance(xbar, y, b_obs, *mean, *variance);
  return 0;
  ORC_CATCH(1)
}
int orc_measurement_vector(const orc_op *o, const double *x0, double radius, double *dense_out) {
  ORC_TRY
  std::vector<long> idx;
  std::vector<double> val;
  measurement_vector(o->p->lattice, x0, radius, idx, val);
  std::fill(dense_out, dense_out + o->p->ndof(), 0.0);
  for (size_t q = 0; q < idx.size(); ++q) dense_out[idx[q]] = val[q];
  return 0;
  ORC_CATCH(1)
}

orc_hier *orc_hier_create(const orc_op *fine, int nlevel, int ordering) {
  ORC_TRY
  orc_hier *h = new orc_hier();
  try {
    h->H = std::make_shared<Hierarchy>(fine->p, nlevel, ordering);
  } catch (...) {
    delete h;
    throw;
  }
  for (auto &op : h->H->ops) h->views.push_back(orc_op{op});
  return h;
  ORC_CATCH(nullptr)
}
void orc_hier_destroy(orc_hier *h) { delete h; }
const orc_op *orc_hier_op(const orc_hier *h, int level) { return &h->views[level]; }
int orc_hier_ncolours(const orc_hier *h, int level) {
  ORC_TRY
  return colour_count_2d(h->H->ops[level]->lattice, h->H->ops[level]->A);
  ORC_CATCH(-1)
}
int orc_hier_order(const orc_hier *h, int level, long *order) {
  std::memcpy(order, h->H->orders[level].data(), sizeof(long) * h->H->orders[level].size());
  return 0;
}
int orc_hier_restrict(const orc_hier *h, int level, const double *xf, double *xc) {
  h->H->intergrids[level].restrict(xf, xc);
  return 0;
}
int orc_hier_prolongate_add(const orc_hier *h, int level, double alpha, const double *xc, double *xf) {
  h->H->intergrids[level].prolongate_add(alpha, xc, xf);
  return 0;
}

static MultigridParameters to_params(const orc_hier *h, const orc_mg_params *p) {
  MultigridParameters q;
  q.nlevel = p->nlevel;
  q.smoother = p->smoother;
  q.coarse_solver = p->coarse_solver;
  q.npresmooth = p->npresmooth;
  q.npostsmooth = p->npostsmooth;
  q.ncoarsesmooth = p->ncoarsesmooth;
  q.cycle = p->cycle;
  q.coarse_scaling = p->coarse_scaling;
  q.omega = p->omega;
  (void)h;
  return q;
}

orc_obj *orc_smoother_create(const orc_hier *h, int level, int kind, double omega, int nsmooth, int direction) {
  ORC_TRY
  orc_obj *o = new orc_obj();
  o->H = h->H;
  const LinearOperator *op = h->H->ops[level].get();
  if (kind == 0) o->sor_smoother = std::make_shared<SORSmoother>(op, omega, nsmooth, (Direction)direction, h->H->orders[level]);
  else o->ssor_smoother = std::make_shared<SSORSmoother>(op, omega, nsmooth, h->H->orders[level]);
  return o;
  ORC_CATCH(nullptr)
}

orc_obj *orc_sampler_create(const orc_hier *h, int level, int kind, double omega, int nsmooth, int direction, orc_rng *rng, uint64_t philox_seed) {
  ORC_TRY
  orc_obj *o = new orc_obj();
  o->H = h->H;
  NoiseSource ns;
  if (rng) {
    ns.engine = &rng->e64;
  } else {
    o->px = std::make_shared<PhiloxCtx>();
    o->px->seed = philox_seed;
    o->px->sweep_counter.assign(h->H->ops.size(), 0u);
    ns.px = o->px.get();
  }
  const LinearOperator *op = h->H->ops[level].get();
  if (kind == 0) o->sor_sampler = std::make_shared<SORSampler>(op, ns, omega, nsmooth, (Direction)direction, h->H->orders[level], level);
  else if (kind == 1) o->ssor_sampler = std::make_shared<SSORSampler>(op, ns, omega, nsmooth, h->H->orders[level], level);
  else o->chol_sampler = std::make_shared<CholeskySampler>(op, ns, level);
  return o;
  ORC_CATCH(nullptr)
}

orc_obj *orc_mgmc_create(const orc_hier *h, const orc_mg_params *p, orc_rng *rng, uint64_t philox_seed) {
  ORC_TRY
  orc_obj *o = new orc_obj();
  o->H = h->H;
  MultigridParameters q = to_params(h, p);
  o->mgmc = std::make_shared<MultigridMCSampler>(h->H, rng ? &rng->e64 : nullptr, q, rng == nullptr, philox_seed);
  o->px = o->mgmc->px;
  return o;
  ORC_CATCH(nullptr)
}

orc_obj *orc_mgprec_create(const orc_hier *h, const orc_mg_params *p) {
  ORC_TRY
  orc_obj *o = new orc_obj();
  o->H = h->H;
  o->mgprec = std::make_shared<MultigridPreconditioner>(h->H, to_params(h, p));
  return o;
  ORC_CATCH(nullptr)
}

orc_obj *orc_cholesky_solver_create(const orc_hier *h, int level) {
  ORC_TRY
  orc_obj *o = new orc_obj();
  o->H = h->H;
  o->chol_solver = std::make_shared<CholeskySolver>(h->H->ops[level].get());
  return o;
  ORC_CATCH(nullptr)
}

void orc_obj_destroy(orc_obj *o) { delete o; }

int orc_obj_apply(orc_obj *o, const double *b, double *x) {
  ORC_TRY
  if (o->sor_smoother) o->sor_smoother->apply(b, x);
  else if (o->ssor_smoother) o->ssor_smoother->apply(b, x);
  else if (o->chol_solver) o->chol_solver->apply(b, x);
  else if (o->mgprec) o->mgprec->apply(b, x);
  else o->sample(b, x);
  return 0;
  ORC_CATCH(1)
}

int orc_obj_set_philox_position(orc_obj *o, uint32_t sample, uint32_t chain, uint32_t sweep_counter) {
  if (!o->px) {
    g_err = "object has no philox stream";
    return 1;
  }
  o->px->sample = sample;
  o->px->chain = chain;
  std::fill(o->px->sweep_counter.begin(), o->px->sweep_counter.end(), sweep_counter);
  return 0;
}

int orc_sampler_run(orc_obj *o, const double *f, double *x, const double *b_obs, long nsamples, double *series) {
  ORC_TRY
  const long n = o->H->ops[0]->ndof();
  for (long k = 0; k < nsamples; ++k) {
    o->sample(f, x);
    if (series) {
      double z = 0.0;
      for (long i = 0; i < n; ++i) z += b_obs[i] * x[i];
      series[k] = z;
    }
  }
  return 0;
  ORC_CATCH(1)
}

int orc_sampler_moments(orc_obj *o, const double *f, double *x, long nwarmup, long nsamples, double *Ex, double *Exx) {
  ORC_TRY
  long n = 0;
  if (o->mgmc) n = o->H->ops[0]->ndof();
  else if (o->sor_sampler) n = o->sor_sampler->op->ndof();
  else if (o->ssor_sampler) n = o->ssor_sampler->fwd.op->ndof();
  else if (o->chol_sampler) n = o->chol_sampler->n;
  else throw std::runtime_error("object is not a sampler");
  std::fill(Ex, Ex + n, 0.0);
  std::fill(Exx, Exx + n * n, 0.0);
  for (long k = 0; k < nwarmup; ++k) o->sample(f, x);
  for (long k = 0; k < nsamples; ++k) {
    o->sample(f, x);
    const double w = 1. / (k + 1);
    for (long i = 0; i < n; ++i) Ex[i] += w * (x[i] - Ex[i]);
    for (long i = 0; i < n; ++i)
      for (long j = 0; j < n; ++j) Exx[i * n + j] += w * (x[i] * x[j] - Exx[i * n + j]);
  }
  return 0;
  ORC_CATCH(1)
}

int orc_loop_solve(const orc_op *op, orc_obj *prec, double rtol, double atol, int maxiter, int verbose, const double *b, double *x,
                   double *history, int *nhist, int *niter, int *converged) {
  ORC_TRY
  if (!prec->mgprec) throw std::runtime_error("orc_loop_solve: preconditioner object required");
  LoopSolverResult r = loop_solve(*op->p, *prec->mgprec, rtol, atol, maxiter, verbose, b, x);
  if (history) std::memcpy(history, r.history.data(), sizeof(double) * r.history.size());
  if (nhist) *nhist = (int)r.history.size();
  if (niter) *niter = r.niter;
  if (converged) *converged = r.converged ? 1 : 0;
  return 0;
  ORC_CATCH(1)
}

double orc_tau_int(const double *series, long n, int window) { return tau_int_scalar(series, n, window); }

void orc_philox_normal_pair(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, double *z0, double *z1) {
  Philox::normal_pair(seed, c0, c1, c2, c3, *z0, *z1);
}
void orc_philox_raw(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t *out4) {
  uint32_t c[4] = {c0, c1, c2, c3};
  Philox::philox4x32(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  for (int k = 0; k < 4; ++k) out4[k] = c[k];
}

}  // extern "C"
