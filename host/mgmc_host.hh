// Host-side drop-in layer: the reference's C++ class names, constructor signatures and apply()
// semantics implemented as thin handles over the C ABI of include/mgmc_b200.h.
// All numerical work happens on the B200 behind mgmc_*; host vectors are lexicographic
// interior-vertex vectors exactly as in the reference.  A nonzero ABI status is converted to the
// reference's error convention: message on stdout + exit(-1).
//
// Deviations from the reference headers (documented):
//  * LinearOperator has no Eigen::SparseMatrix inside; get_sparse() does not exist.  coarsen() returns
//    the next level of the same device hierarchy (matrix-free Galerkin stencil, linear_operator.cc:10-23).
//  * Samplers take the reference's std::mt19937_64& for signature compatibility and use ONE draw from
//    it as the Philox key; the chain is not the reference's mt19937 chain (multicolour + Philox).
//  * Supported on the device path: dim = 2, shiftedlaplace_fd and squared_shiftedlaplace_fd with constant
//    correlation length, sparse measurement matrices.  Everything else exits with the ABI's "unsupported" message.
#ifndef MGMC_HOST_HH
#define MGMC_HOST_HH
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <random>
#include <string>
#include <vector>

#include "../include/mgmc_b200.h"
#include "config.hh"
#include "vector.hh"

namespace mgmc_host {
inline void check(int status, const char *what) {
  if (status != MGMC_OK) {
    std::cout << "ERROR: " << what << ": " << mgmc_last_error() << std::endl;
    exit(-1);
  }
}
}  // namespace mgmc_host

// ------------------------------------------------------------------------------------------------
// Lattice (lattice/lattice.hh:18-129, lattice/lattice2d.hh:43-235)
// ------------------------------------------------------------------------------------------------
class Lattice {
 public:
  Lattice(const unsigned int Ncell_, const unsigned int Nvertex_) : Ncell(Ncell_), Nvertex(Nvertex_) {}
  virtual ~Lattice() = default;
  virtual Eigen::VectorXi shape() const = 0;
  virtual int dim() const { return (int)shape().size(); }
  double cell_volume() const {
    Eigen::VectorXi s = shape();
    double v = 1.0;
    for (int d = 0; d < dim(); ++d) v /= s[d];
    return v;
  }
  virtual Eigen::VectorXi cellidx_linear2euclidean(const unsigned int ell) const = 0;
  virtual unsigned int cellidx_euclidean2linear(const Eigen::VectorXi idx) const = 0;
  virtual Eigen::VectorXi vertexidx_linear2euclidean(const unsigned int ell) const = 0;
  virtual unsigned int vertexidx_euclidean2linear(const Eigen::VectorXi idx) const = 0;
  virtual unsigned int shift_vertexidx(const unsigned int ell, const Eigen::VectorXi shift) const = 0;
  virtual bool shifted_vertex_is_internal_vertex(const unsigned int ell, const Eigen::VectorXi shift, unsigned int &idx_vertex) const = 0;
  virtual unsigned int fine_vertex_idx(const unsigned int ell) const = 0;
  virtual Eigen::VectorXd vertex_coordinates(const unsigned int ell) const = 0;
  virtual std::shared_ptr<Lattice> get_coarse_lattice() const = 0;
  virtual std::string get_info() const = 0;
  const unsigned int Ncell;
  const unsigned int Nvertex;
};

class Lattice2d : public Lattice {
 public:
  Lattice2d(const unsigned int nx_, const unsigned int ny_) : Lattice(nx_ * ny_, (nx_ - 1) * (ny_ - 1)), nx(nx_), ny(ny_), hx(1. / double(nx_)), hy(1. / double(ny_)) {}
  Eigen::VectorXi shape() const override { return Eigen::VectorXi({(int)nx, (int)ny}); }
  Eigen::VectorXi cellidx_linear2euclidean(const unsigned int ell) const override { return Eigen::VectorXi({(int)(ell % nx), (int)(ell / nx)}); }
  unsigned int cellidx_euclidean2linear(const Eigen::VectorXi idx) const override { return idx[1] * nx + idx[0]; }
  Eigen::VectorXi vertexidx_linear2euclidean(const unsigned int ell) const override {
    return Eigen::VectorXi({(int)(ell % (nx - 1)) + 1, (int)(ell / (nx - 1)) + 1});
  }
  unsigned int vertexidx_euclidean2linear(const Eigen::VectorXi idx) const override { return (idx[1] - 1) * (nx - 1) + (idx[0] - 1); }
  unsigned int shift_cellidx(const unsigned int ell, const Eigen::VectorXi shift) const {  // lattice2d.hh:106-118
    return ((int)(ell / nx) + shift[1]) * nx + ((int)(ell % nx) + shift[0]);
  }
  unsigned int shift_vertexidx(const unsigned int ell, const Eigen::VectorXi shift) const override {
    const int i = (int)(ell % (nx - 1)) + shift[0] + 1, j = (int)(ell / (nx - 1)) + shift[1] + 1;
    return (j - 1) * (nx - 1) + (i - 1);
  }
  bool shifted_vertex_is_internal_vertex(const unsigned int ell, const Eigen::VectorXi shift, unsigned int &idx_vertex) const override {
    const int i = (int)(ell % (nx - 1)) + shift[0] + 1, j = (int)(ell / (nx - 1)) + shift[1] + 1;
    idx_vertex = (j - 1) * (nx - 1) + (i - 1);
    return (i > 0) && (i < (int)nx) && (j > 0) && (j < (int)ny);
  }
  unsigned int fine_vertex_idx(const unsigned int ell) const override {
    const int i = (int)(ell % (nx - 1)) + 1, j = (int)(ell / (nx - 1)) + 1;
    return (2 * j - 1) * (2 * nx - 1) + (2 * i - 1);
  }
  Eigen::VectorXd vertex_coordinates(const unsigned int ell) const override {
    return Eigen::VectorXd({(ell % (nx - 1) + 1.0) * hx, (ell / (nx - 1) + 1.0) * hy});
  }
  std::shared_ptr<Lattice> get_coarse_lattice() const override {  // lattice2d.hh:198-213
    if (!((nx % 2 == 0) && (ny % 2 == 0))) {
      std::cout << "ERROR: cannot coarsen lattice of size " << nx << " x " << ny << " [one of the extents is odd]" << std::endl;
      exit(-1);
    }
    if (!((nx / 2 > 1) && (ny / 2 > 1))) {
      std::cout << "ERROR: cannot coarsen lattice of size " << nx << " x " << ny << " [resulting lattice would have no interior vertices]" << std::endl;
      exit(-1);
    }
    return std::make_shared<Lattice2d>(nx / 2, ny / 2);
  }
  std::string get_info() const override {
    char b[128];
    std::snprintf(b, 128, "2d lattice, %4d x %4d points, %8d cells, %8d vertices", nx, ny, Ncell, Nvertex);
    return std::string(b);
  }
  const unsigned int nx, ny;
  const double hx, hy;
};

/** Lattice1d (lattice/lattice1d.hh:27-190): index arithmetic only -- operators on 1d lattices are not on the device path
 *  (mgmc_create reports dim = 1 as MGMC_ERR_UNSUPPORTED) */
class Lattice1d : public Lattice {
 public:
  explicit Lattice1d(const unsigned int n_) : Lattice(n_, n_ - 1), n(n_), h(1. / double(n_)) {}
  Eigen::VectorXi shape() const override { return Eigen::VectorXi({(int)n}); }
  Eigen::VectorXi cellidx_linear2euclidean(const unsigned int ell) const override { return Eigen::VectorXi({(int)ell}); }
  unsigned int cellidx_euclidean2linear(const Eigen::VectorXi idx) const override { return idx[0]; }
  Eigen::VectorXi vertexidx_linear2euclidean(const unsigned int ell) const override { return Eigen::VectorXi({(int)ell + 1}); }
  unsigned int vertexidx_euclidean2linear(const Eigen::VectorXi idx) const override { return idx[0] - 1; }
  unsigned int shift_cellidx(const unsigned int ell, const Eigen::VectorXi shift) const { return ell + shift[0]; }
  unsigned int shift_vertexidx(const unsigned int ell, const Eigen::VectorXi shift) const override { return ell + shift[0]; }
  bool shifted_vertex_is_internal_vertex(const unsigned int ell, const Eigen::VectorXi shift, unsigned int &idx_vertex) const override {
    const int i = (int)ell + shift[0] + 1;
    idx_vertex = i - 1;
    return (i > 0) && (i < (int)n);
  }
  unsigned int fine_vertex_idx(const unsigned int ell) const override { return 2 * ell + 1; }
  Eigen::VectorXd vertex_coordinates(const unsigned int ell) const override { return Eigen::VectorXd({(ell + 1.) * h}); }
  std::shared_ptr<Lattice> get_coarse_lattice() const override {  // lattice1d.hh:155-170
    if (!(n % 2 == 0)) {
      std::cout << "ERROR: cannot coarsen lattice of size " << n << " [extent is odd]" << std::endl;
      exit(-1);
    }
    if (!(n / 2 > 1)) {
      std::cout << "ERROR: cannot coarsen lattice of size " << n << " [resulting lattice would have no interior vertices]" << std::endl;
      exit(-1);
    }
    return std::make_shared<Lattice1d>(n / 2);
  }
  std::string get_info() const override {  // lattice1d.cc:8-14
    char b[128];
    std::snprintf(b, 128, "1d lattice, %4d points, %4d unknowns", n, Nvertex);
    return std::string(b);
  }
  const unsigned int n;
  const double h;
};

/** Lattice3d (lattice/lattice3d.hh:43-270): interior vertices in the order ell = (k-1)(nx-1)(ny-1) + (j-1)(nx-1) + (i-1) */
class Lattice3d : public Lattice {
 public:
  Lattice3d(const unsigned int nx_, const unsigned int ny_, const unsigned int nz_)
      : Lattice(nx_ * ny_ * nz_, (nx_ - 1) * (ny_ - 1) * (nz_ - 1)), nx(nx_), ny(ny_), nz(nz_), hx(1. / double(nx_)), hy(1. / double(ny_)), hz(1. / double(nz_)) {}
  Eigen::VectorXi shape() const override { return Eigen::VectorXi({(int)nx, (int)ny, (int)nz}); }
  Eigen::VectorXi cellidx_linear2euclidean(const unsigned int ell) const override {
    return Eigen::VectorXi({(int)((ell % (nx * ny)) % nx), (int)((ell % (nx * ny)) / nx), (int)(ell / (nx * ny))});
  }
  unsigned int cellidx_euclidean2linear(const Eigen::VectorXi idx) const override { return idx[2] * nx * ny + idx[1] * nx + idx[0]; }
  Eigen::VectorXi vertexidx_linear2euclidean(const unsigned int ell) const override {
    const unsigned int w = nx - 1, h = ny - 1;
    return Eigen::VectorXi({(int)((ell % (w * h)) % w) + 1, (int)((ell % (w * h)) / w) + 1, (int)(ell / (w * h)) + 1});
  }
  unsigned int vertexidx_euclidean2linear(const Eigen::VectorXi idx) const override {
    return (idx[2] - 1) * (nx - 1) * (ny - 1) + (idx[1] - 1) * (nx - 1) + (idx[0] - 1);
  }
  unsigned int shift_cellidx(const unsigned int ell, const Eigen::VectorXi shift) const {  // lattice3d.hh:138-156
    const Eigen::VectorXi c = cellidx_linear2euclidean(ell);
    return (c[2] + shift[2]) * nx * ny + (c[1] + shift[1]) * nx + (c[0] + shift[0]);
  }
  unsigned int shift_vertexidx(const unsigned int ell, const Eigen::VectorXi shift) const override {
    const Eigen::VectorXi p = vertexidx_linear2euclidean(ell);
    return (p[2] + shift[2] - 1) * (nx - 1) * (ny - 1) + (p[1] + shift[1] - 1) * (nx - 1) + (p[0] + shift[0] - 1);
  }
  bool shifted_vertex_is_internal_vertex(const unsigned int ell, const Eigen::VectorXi shift, unsigned int &idx_vertex) const override {
    const Eigen::VectorXi p = vertexidx_linear2euclidean(ell);
    const int i = p[0] + shift[0], j = p[1] + shift[1], k = p[2] + shift[2];
    idx_vertex = (k - 1) * (nx - 1) * (ny - 1) + (j - 1) * (nx - 1) + (i - 1);
    return (i > 0) && (i < (int)nx) && (j > 0) && (j < (int)ny) && (k > 0) && (k < (int)nz);
  }
  unsigned int fine_vertex_idx(const unsigned int ell) const override {  // lattice3d.hh:217-225
    const Eigen::VectorXi p = vertexidx_linear2euclidean(ell);
    return (2 * p[2] - 1) * (2 * nx - 1) * (2 * ny - 1) + (2 * p[1] - 1) * (2 * nx - 1) + 2 * p[0] - 1;
  }
  Eigen::VectorXd vertex_coordinates(const unsigned int ell) const override {
    const Eigen::VectorXi p = vertexidx_linear2euclidean(ell);
    return Eigen::VectorXd({p[0] * hx, p[1] * hy, p[2] * hz});
  }
  std::shared_ptr<Lattice> get_coarse_lattice() const override {  // lattice3d.hh:242-257
    if (!((nx % 2 == 0) && (ny % 2 == 0) && (nz % 2 == 0))) {
      std::cout << "ERROR: cannot coarsen lattice of size " << nx << " x " << ny << " x " << nz << " [one of the extents is odd]" << std::endl;
      exit(-1);
    }
    if (!((nx / 2 > 1) && (ny / 2 > 1) && (nz / 2 > 1))) {
      std::cout << "ERROR: cannot coarsen lattice of size " << nx << " x " << ny << " x " << nz << " [resulting lattice would have no interior vertices]" << std::endl;
      exit(-1);
    }
    return std::make_shared<Lattice3d>(nx / 2, ny / 2, nz / 2);
  }
  std::string get_info() const override {  // lattice3d.cc:8-14
    char b[128];
    std::snprintf(b, 128, "3d lattice, %4d x %4d x %4d points, %4d unknowns", nx, ny, nz, Nvertex);
    return std::string(b);
  }
  const unsigned int nx, ny, nz;
  const double hx, hy, hz;
};

// ------------------------------------------------------------------------------------------------
// GaussLegendreQuadrature (auxilliary/quadrature.hh:22-47, quadrature.cc:11-59): tensor-product Gauss-Legendre rule with
// order + 1 points per direction on the unit cube [0,1]^dim (order 0, 1, 2) -- the rule of the FEM assembly and of the
// measurement functional (the host layer's own assembly loops use the order-1 rule inline)
// ------------------------------------------------------------------------------------------------
class GaussLegendreQuadrature {
 public:
  GaussLegendreQuadrature(const int dim_, const int order_) : dim(dim_), order(order_) {
    static const double node[3][3] = {{0.0, 0.0, 0.0}, {-0.57735026918962576451, +0.57735026918962576451, 0.0}, {-0.77459666924148337704, 0.0, +0.77459666924148337704}};
    static const double weight[3][3] = {{2.0, 0.0, 0.0}, {1.0, 1.0, 0.0}, {5.0 / 9.0, 8.0 / 9.0, 5.0 / 9.0}};
    if (dim < 1 || order < 0 || order > 2) {
      std::cout << "ERROR: GaussLegendreQuadrature needs dim > 0 and 0 <= order < 3" << std::endl;
      exit(-1);
    }
    const int n1 = order + 1;
    long total = 1;
    for (int d = 0; d < dim; ++d) total *= n1;
    for (long q = 0; q < total; ++q) {  // first direction slowest (the order of a cartesian product)
      Eigen::VectorXd pt(dim);
      double w = 1.0;
      long rem = q;
      for (int d = dim - 1; d >= 0; --d) {
        const int k = (int)(rem % n1);
        rem /= n1;
        pt[d] = 0.5 * (node[order][k] + 1.0);  // [-1, +1] -> [0, 1]
        w *= 0.5 * weight[order][k];
      }
      weights.push_back(w);
      points.push_back(pt);
    }
  }
  std::vector<double> get_weights() const { return weights; }
  std::vector<Eigen::VectorXd> get_points() const { return points; }

 protected:
  const int dim, order;
  std::vector<double> weights;
  std::vector<Eigen::VectorXd> points;
};

// ------------------------------------------------------------------------------------------------
// Correlation length models (linear_operator/correlationlength_model.hh:45-113)
// ------------------------------------------------------------------------------------------------
class CorrelationLengthModel {
 public:
  virtual ~CorrelationLengthModel() = default;
  virtual double kappa_sq(const Eigen::VectorXd x) const = 0;
  virtual bool is_constant(double &Lambda) const = 0;
};
class ConstantCorrelationLengthModel : public CorrelationLengthModel {
 public:
  explicit ConstantCorrelationLengthModel(const ConstantCorrelationLengthModelParameters p) : Lambda_(p.Lambda), kappa_sq_(1. / std::pow(p.Lambda, 2)) {}
  double kappa_sq(const Eigen::VectorXd) const override { return kappa_sq_; }
  bool is_constant(double &Lambda) const override {
    Lambda = Lambda_;
    return true;
  }

 protected:
  const double Lambda_, kappa_sq_;
};
class PeriodicCorrelationLengthModel : public CorrelationLengthModel {
 public:
  explicit PeriodicCorrelationLengthModel(const PeriodicCorrelationLengthModelParameters p)
      : Lambda_1(0.5 * (p.Lambda_max + p.Lambda_min)), Lambda_2(0.5 * (p.Lambda_max - p.Lambda_min)) {}
  double kappa_sq(const Eigen::VectorXd x) const override {
    double L = Lambda_2;
    for (int d = 0; d < (int)x.size(); ++d) L *= std::cos(M_PI * x[d]);
    L += Lambda_1;
    return 1. / (L * L);
  }
  bool is_constant(double &) const override { return false; }

 protected:
  const double Lambda_1, Lambda_2;
};

// ------------------------------------------------------------------------------------------------
// Device hierarchy shared by the operator / smoother / sampler / solver handles
// ------------------------------------------------------------------------------------------------
struct OperatorData {
  unsigned int nx = 0, ny = 0;
  unsigned int nz = 0;  // > 0: Lattice3d
  int pde_model = MGMC_PDE_SHIFTEDLAPLACE_FD;
  double Lambda = 1.0;
  bool constant_kappa = true;
  std::vector<double> kappa_sq;  // variable correlation length: kappa^2 at every interior vertex (lexicographic)
  std::vector<int64_t> B_rows;
  std::vector<int32_t> B_cols;
  std::vector<double> B_vals, Sigma;
};

class DeviceHierarchy {
 public:
  DeviceHierarchy(const OperatorData &d, const MultigridParameters &p, uint64_t seed) : data(d), params(p) {
    mgmc_desc desc;
    std::memset(&desc, 0, sizeof(desc));
    desc.kappa_sq = d.constant_kappa ? nullptr : d.kappa_sq.data();
    desc.dim = d.nz ? 3 : 2;
    desc.nx = (int)d.nx;
    desc.ny = (int)d.ny;
    desc.nz = d.nz ? (int)d.nz : 1;
    desc.pde_model = d.pde_model;
    desc.Lambda = d.Lambda;
    desc.m_lowrank = (int)d.Sigma.size();
    desc.B_nnz = (int64_t)d.B_vals.size();
    desc.B_rows = d.B_rows.data();
    desc.B_cols = d.B_cols.data();
    desc.B_vals = d.B_vals.data();
    desc.Sigma = d.Sigma.data();
    desc.nlevel = (int)p.nlevel;
    desc.smoother = (p.smoother == "SOR") ? MGMC_SMOOTHER_SOR : MGMC_SMOOTHER_SSOR;
    desc.coarse_solver = (p.coarse_solver == "SSOR") ? MGMC_COARSE_SSOR : MGMC_COARSE_CHOLESKY;
    desc.npresmooth = (int)p.npresmooth;
    desc.npostsmooth = (int)p.npostsmooth;
    desc.ncoarsesmooth = (int)p.ncoarsesmooth;
    desc.cycle = (int)p.cycle;
    desc.coarse_scaling = p.coarse_scaling;
    desc.omega = p.omega;
    desc.seed = seed;
    desc.device = 0;
    desc.nchains = 1;
    desc.first_chain = 0;
    mgmc_host::check(mgmc_create(&desc, &ctx), "mgmc_create");
  }
  ~DeviceHierarchy() { mgmc_destroy(ctx); }
  DeviceHierarchy(const DeviceHierarchy &) = delete;
  mgmc_ctx *ctx = nullptr;
  OperatorData data;
  MultigridParameters params;
};

class IntergridOperator;

// ------------------------------------------------------------------------------------------------
// LinearOperator A = A_0 + B Sigma^{-1} B^T (linear_operator/linear_operator.hh:28-198)
// ------------------------------------------------------------------------------------------------
class LinearOperator {
 public:
  LinearOperator(const std::shared_ptr<Lattice> lattice_, const unsigned int m_lowrank_ = 0)
      : lattice(lattice_), m_lowrank(m_lowrank_), data(std::make_shared<OperatorData>()), level(0) {}
  virtual ~LinearOperator() = default;
  std::shared_ptr<Lattice> get_lattice() const { return lattice; }
  unsigned int get_ndof() const { return lattice->Nvertex; }
  unsigned int get_m_lowrank() const { return m_lowrank; }
  const OperatorData &get_data() const { return *data; }
  int get_level() const { return level; }

  /** y = A x (linear_operator.hh:66-76) */
  void apply(const Eigen::VectorXd &x, Eigen::VectorXd &y) {
    mgmc_host::check(mgmc_op_apply(hierarchy(level + 1)->ctx, level, x.data(), y.data()), "LinearOperator::apply");
  }
  /** Galerkin coarsening (linear_operator.cc:10-23): next level of the same device hierarchy */
  LinearOperator coarsen(const std::shared_ptr<IntergridOperator>) const {
    LinearOperator c(lattice->get_coarse_lattice(), m_lowrank);
    c.data = data;
    c.level = level + 1;
    c.dev = dev;
    return c;
  }
  /** the sparse part A_0 of this level as an explicit matrix (linear_operator.hh:79 get_sparse).  The device path is matrix-free: the
   *  matrix is assembled on demand from the host-only stencil algebra of the library (mgmc_host_stencil / mgmc_host_stencil3: the
   *  fine stencil and its Galerkin products R A R^T) -- no device needed; constant correlation length only */
  typedef Eigen::SparseMatrix<double> SparseMatrixType;
  const SparseMatrixType &get_sparse() const {
    if (sparse_cache) return *sparse_cache;
    if (!data->constant_kappa) {
      std::cout << "ERROR: get_sparse() of the host layer needs a constant correlation length (per-vertex operators: mgmc_host_coefficients)" << std::endl;
      exit(-1);
    }
    mgmc_desc desc;
    std::memset(&desc, 0, sizeof(desc));
    desc.dim = data->nz ? 3 : 2;
    desc.nx = (int)data->nx;
    desc.ny = (int)data->ny;
    desc.nz = data->nz ? (int)data->nz : 1;
    desc.pde_model = data->pde_model;
    desc.Lambda = data->Lambda;
    desc.nlevel = level + 1;
    desc.smoother = MGMC_SMOOTHER_SSOR;
    desc.coarse_solver = MGMC_COARSE_CHOLESKY;
    desc.npresmooth = desc.npostsmooth = desc.ncoarsesmooth = 1;
    desc.cycle = 1;
    desc.coarse_scaling = 1.0;
    desc.omega = 1.0;
    desc.nchains = 1;
    const Eigen::VectorXi shp = lattice->shape();
    const long n = lattice->Nvertex;
    auto M = std::make_shared<SparseMatrixType>(n, n);
    int nc = 0;
    if (data->nz) {
      double st[27];
      mgmc_host::check(mgmc_host_stencil3(&desc, level, st, &nc), "LinearOperator::get_sparse");
      const int nx = shp[0], ny = shp[1], nz = shp[2];
      for (int k = 1; k < nz; ++k)
        for (int j = 1; j < ny; ++j)
          for (int i = 1; i < nx; ++i)
            for (int dk = -1; dk <= 1; ++dk)
              for (int dj = -1; dj <= 1; ++dj)
                for (int di = -1; di <= 1; ++di) {
                  const int ii = i + di, jj = j + dj, kk = k + dk;
                  const double v = st[(dk + 1) * 9 + (dj + 1) * 3 + (di + 1)];
                  if (v == 0.0 || ii < 1 || ii >= nx || jj < 1 || jj >= ny || kk < 1 || kk >= nz) continue;
                  M->coeffRef(((long)(k - 1) * (ny - 1) + (j - 1)) * (nx - 1) + (i - 1), ((long)(kk - 1) * (ny - 1) + (jj - 1)) * (nx - 1) + (ii - 1)) = v;
                }
    } else {
      double st[225];
      mgmc_host::check(mgmc_host_stencil(&desc, level, st, &nc), "LinearOperator::get_sparse");
      const int nx = shp[0], ny = shp[1];
      auto cls = [](int i, int n_) { return i == 1 ? 0 : (i == n_ - 1 ? 2 : 1); };
      for (int j = 1; j < ny; ++j)
        for (int i = 1; i < nx; ++i) {
          const double *a = st + 25 * (cls(i, nx) + 3 * cls(j, ny));
          for (int dj = -2; dj <= 2; ++dj)
            for (int di = -2; di <= 2; ++di) {
              const int ii = i + di, jj = j + dj;
              const double v = a[(dj + 2) * 5 + (di + 2)];
              if (v == 0.0 || ii < 1 || ii >= nx || jj < 1 || jj >= ny) continue;
              M->coeffRef((long)(j - 1) * (nx - 1) + (i - 1), (long)(jj - 1) * (nx - 1) + (ii - 1)) = v;
            }
        }
    }
    sparse_cache = M;
    return *sparse_cache;
  }
  /** device hierarchy with at least nlevel levels (created lazily, shared between the handles) */
  std::shared_ptr<DeviceHierarchy> hierarchy(int nlevel) const {
    if (!dev || (int)dev->params.nlevel < nlevel) {
      MultigridParameters p;
      p.nlevel = (unsigned int)nlevel;
      dev = std::make_shared<DeviceHierarchy>(*data, p, 0);
    }
    return dev;
  }

  /** posterior mean (linear_operator.hh:119-139); the reference factorises the fine matrix, here the
   *  m + 1 prior solves run as multigrid-preconditioned Richardson iterations on the device */
  Eigen::VectorXd mean(const Eigen::VectorXd &xbar, const Eigen::VectorXd &y) const;
  /** mean and variance of z = b^T x (linear_operator.hh:153-174) */
  void observed_mean_and_variance(const Eigen::VectorXd &xbar, const Eigen::VectorXd &y, const Eigen::SparseVector<double> &b_obs, double &mean_,
                                  double &variance) const;

 protected:
  std::vector<Eigen::VectorXd> prior_solve_columns(const std::vector<Eigen::VectorXd> &rhs) const;
  const std::shared_ptr<Lattice> lattice;
  const unsigned int m_lowrank;
  std::shared_ptr<OperatorData> data;
  int level;
  mutable std::shared_ptr<DeviceHierarchy> dev;
  mutable std::shared_ptr<SparseMatrixType> sparse_cache;
};

/** ShiftedLaplaceFDOperator (linear_operator/shiftedlaplace_fd_operator.hh:28-42) */
class ShiftedLaplaceFDOperator : public LinearOperator {
 public:
  ShiftedLaplaceFDOperator(const std::shared_ptr<Lattice> lattice_, const std::shared_ptr<CorrelationLengthModel> clm, const int verbose = 0) : LinearOperator(lattice_) {
    (void)verbose;
    Eigen::VectorXi s = lattice->shape();
    if (lattice->dim() != 2 && lattice->dim() != 3) {
      std::cout << "ERROR: the device path supports dim = 2 and dim = 3 only" << std::endl;
      exit(-1);
    }
    data->nx = s[0];
    data->ny = s[1];
    data->nz = (lattice->dim() == 3) ? s[2] : 0;
    data->pde_model = MGMC_PDE_SHIFTEDLAPLACE_FD;
    data->constant_kappa = clm->is_constant(data->Lambda);
    if (!data->constant_kappa) {
      // kappa^2 at every interior vertex, evaluated as the reference's assembly loop does (shiftedlaplace_fd_operator.cc:33-36)
      data->kappa_sq.resize(lattice->Nvertex);
      for (unsigned int ell = 0; ell < lattice->Nvertex; ++ell) data->kappa_sq[ell] = clm->kappa_sq(lattice->vertex_coordinates(ell));
    }
  }
};

/** ShiftedLaplaceFEMOperator (linear_operator/shiftedlaplace_fem_operator.hh:31-75): multilinear elements; on the device with a constant
 *  correlation length (a uniform 9-point / 27-point stencil), dim = 2 and 3 */
class ShiftedLaplaceFEMOperator : public LinearOperator {
 public:
  ShiftedLaplaceFEMOperator(const std::shared_ptr<Lattice> lattice_, const std::shared_ptr<CorrelationLengthModel> clm, const int verbose = 0) : LinearOperator(lattice_) {
    (void)verbose;
    Eigen::VectorXi s = lattice->shape();
    if (lattice->dim() != 2 && lattice->dim() != 3) {
      std::cout << "ERROR: the device path supports dim = 2 and dim = 3 only" << std::endl;
      exit(-1);
    }
    data->nx = s[0];
    data->ny = s[1];
    data->nz = (lattice->dim() == 3) ? s[2] : 0;
    data->pde_model = MGMC_PDE_SHIFTEDLAPLACE_FEM;
    data->constant_kappa = clm->is_constant(data->Lambda);
    if (!data->constant_kappa) {
      std::cout << "ERROR: pdemodel 'shiftedlaplace_fem' is on the device path with a constant correlation length only" << std::endl;
      exit(-1);
    }
  }
};

/** SquaredShiftedLaplaceFDOperator (squared_shiftedlaplace_fd_operator.hh): 13 / 21-point stencils with
 *  boundary-ring classes, 9-colour sweeps on the device */
class SquaredShiftedLaplaceFDOperator : public LinearOperator {
 public:
  SquaredShiftedLaplaceFDOperator(const std::shared_ptr<Lattice> lattice_, const std::shared_ptr<CorrelationLengthModel> clm, const int verbose = 0) : LinearOperator(lattice_) {
    (void)verbose;
    Eigen::VectorXi s = lattice->shape();
    if (lattice->dim() != 2) {  // (the reference's assembly is 2d only as well: squared_shiftedlaplace_fd_operator.cc:16-20)
      std::cout << "ERROR: the squared shifted Laplace operator is implemented for dim = 2 only" << std::endl;
      exit(-1);
    }
    data->nx = s[0];
    data->ny = s[1];
    data->pde_model = MGMC_PDE_SQUARED_SHIFTEDLAPLACE_FD;
    data->constant_kappa = clm->is_constant(data->Lambda);
  }
};

/** MeasuredOperator (linear_operator/measured_operator.hh:26-80, measured_operator.cc:9-170): B is
 *  assembled on the host (setup-time, tiny) and handed to the device as COO triplets */
class MeasuredOperator : public LinearOperator {
 public:
  MeasuredOperator(const std::shared_ptr<LinearOperator> base_operator_, const MeasurementParameters params_)
      : LinearOperator(base_operator_->get_lattice(), (unsigned int)(params_.measurement_locations.size() + params_.measure_global)), params(params_) {
    *data = base_operator_->get_data();
    data->B_rows.clear();
    data->B_cols.clear();
    data->B_vals.clear();
    const unsigned int n_meas = (unsigned int)params.measurement_locations.size();
    data->Sigma.assign(n_meas + params.measure_global, 0.0);
    for (unsigned int k = 0; k < n_meas; ++k) {
      data->Sigma[k] = params.variance_scaling * params.variance[k];
      Eigen::SparseVector<double> r = measurement_vector(params.measurement_locations[k], params.radius);
      for (auto &e : r.entries()) {
        data->B_rows.push_back((int64_t)e.first);
        data->B_cols.push_back((int32_t)k);
        data->B_vals.push_back(e.second);
      }
    }
    if (params.measure_global) {  // measured_operator.cc:31-46 (a dense column of B: the device runs its chip-wide low-rank kernels)
      const double cv = lattice->cell_volume();
      for (unsigned int ell = 0; ell < lattice->Nvertex; ++ell) {
        data->B_rows.push_back(ell);
        data->B_cols.push_back((int32_t)n_meas);
        data->B_vals.push_back(cv);
      }
      data->Sigma[n_meas] = params.variance_global;
    }
  }

  /** measurement functional in dual space (measured_operator.cc:69-170) */
  Eigen::SparseVector<double> measurement_vector(const Eigen::VectorXd x0, const double radius) const {
    Eigen::SparseVector<double> r(lattice->Nvertex);
    const Eigen::VectorXi shape = lattice->shape();
    if (lattice->dim() == 3) return measurement_vector3(x0, radius);
    const int nx = shape[0], ny = shape[1];
    const double hx = 1. / nx, hy = 1. / ny;
    if (radius < 1.E-12) {
      // closest vertex, first minimum in lexicographic order (only the 2 x 2 candidates can win)
      double d_min = 2.0;
      unsigned int ell_min = 0;
      const int ci = (int)std::floor(x0[0] * nx), cj = (int)std::floor(x0[1] * ny);
      // (candidates clamped to the interior: a point on or beyond the upper boundary picks the last interior line, as
      //  the reference's scan over all vertices does)
      const int j0 = std::min(std::max(cj, 1), ny - 1), j1 = std::min(std::max(cj + 1, 1), ny - 1);
      const int i0 = std::min(std::max(ci, 1), nx - 1), i1 = std::min(std::max(ci + 1, 1), nx - 1);
      for (int j = j0; j <= j1; ++j)
        for (int i = i0; i <= i1; ++i) {
          const double dist = std::sqrt((i * hx - x0[0]) * (i * hx - x0[0]) + (j * hy - x0[1]) * (j * hy - x0[1]));
          if (dist < d_min) {
            d_min = dist;
            ell_min = (j - 1) * (nx - 1) + (i - 1);
          }
        }
      r.coeffRef(ell_min) = 1.0;
      return r;
    }
    // average over a ball of the given radius against the bilinear hat functions, 2-point Gauss rule
    const double cell_volume = lattice->cell_volume();
    const double normalisation = 1. / (M_PI * radius * radius);
    const double gp[2] = {0.5 * (1.0 - 1.0 / std::sqrt(3.0)), 0.5 * (1.0 + 1.0 / std::sqrt(3.0))};
    const int i_lo = std::max(0, (int)std::floor((x0[0] - radius) * nx) - 1), i_hi = std::min(nx - 1, (int)std::floor((x0[0] + radius) * nx) + 1);
    const int j_lo = std::max(0, (int)std::floor((x0[1] - radius) * ny) - 1), j_hi = std::min(ny - 1, (int)std::floor((x0[1] + radius) * ny) + 1);
    for (int cj = j_lo; cj <= j_hi; ++cj)
      for (int ci = i_lo; ci <= i_hi; ++ci) {
        bool overlap = false;
        for (int oy = 0; oy < 2; ++oy)
          for (int ox = 0; ox < 2; ++ox) {
            const double dx = hx * (ci + ox) - x0[0], dy = hy * (cj + oy) - x0[1];
            overlap = overlap || (std::sqrt(dx * dx + dy * dy) < radius);
          }
        const bool centre_in_cell = (hx * ci <= x0[0]) && (x0[0] <= hx * (ci + 1)) && (hy * cj <= x0[1]) && (x0[1] <= hy * (cj + 1));
        if (!(overlap || centre_in_cell)) continue;
        for (int ax = 0; ax < 2; ++ax)
          for (int ay = 0; ay < 2; ++ay) {
            const int i = ci + ax, j = cj + ay;
            if (!(i > 0 && i < nx && j > 0 && j < ny)) continue;
            double local = 0.0;
            for (int qx = 0; qx < 2; ++qx)
              for (int qy = 0; qy < 2; ++qy) {
                const double dx = hx * (gp[qx] + ci) - x0[0], dy = hy * (gp[qy] + cj) - x0[1];
                if (std::sqrt(dx * dx + dy * dy) / radius < 1.0) {
                  const double phi = (ax == 0 ? 1.0 - gp[qx] : gp[qx]) * (ay == 0 ? 1.0 - gp[qy] : gp[qy]);
                  local += phi * 0.25 * cell_volume * normalisation;
                }
              }
            r.coeffRef((j - 1) * (nx - 1) + (i - 1)) += local;
          }
      }
    return r;
  }

  /** measurement functional on a Lattice3d (measured_operator.cc:69-170 with dim = 3): closest vertex for radius ~ 0 (first
   *  minimum in lexicographic order; only the 2 x 2 x 2 candidates around x0 can win), else the average over a ball of the
   *  given radius against the trilinear hat functions, 2-point Gauss rule per direction, normalised by 4/3 pi r^3 */
  Eigen::SparseVector<double> measurement_vector3(const Eigen::VectorXd x0, const double radius) const {
    Eigen::SparseVector<double> r(lattice->Nvertex);
    const Eigen::VectorXi shape = lattice->shape();
    const int n[3] = {shape[0], shape[1], shape[2]};
    const double h[3] = {1. / n[0], 1. / n[1], 1. / n[2]};
    auto lin = [&](int i, int j, int k) { return (unsigned int)(((k - 1) * (n[1] - 1) + (j - 1)) * (n[0] - 1) + (i - 1)); };
    if (radius < 1.E-12) {
      double d_min = 3.0;
      unsigned int ell_min = 0;
      int lo[3], hi[3];
      for (int d = 0; d < 3; ++d) {
        const int c = (int)std::floor(x0[d] * n[d]);
        lo[d] = std::min(std::max(c, 1), n[d] - 1);
        hi[d] = std::min(std::max(c + 1, 1), n[d] - 1);
      }
      for (int k = lo[2]; k <= hi[2]; ++k)
        for (int j = lo[1]; j <= hi[1]; ++j)
          for (int i = lo[0]; i <= hi[0]; ++i) {
            const double dx = i * h[0] - x0[0], dy = j * h[1] - x0[1], dz = k * h[2] - x0[2];
            const double dist = std::sqrt(dx * dx + dy * dy + dz * dz);
            if (dist < d_min) {
              d_min = dist;
              ell_min = lin(i, j, k);
            }
          }
      r.coeffRef(ell_min) = 1.0;
      return r;
    }
    const double cell_volume = lattice->cell_volume();
    const double normalisation = 1. / (4. / 3. * M_PI * radius * radius * radius);
    const double gp[2] = {0.5 * (1.0 - 1.0 / std::sqrt(3.0)), 0.5 * (1.0 + 1.0 / std::sqrt(3.0))};
    int c_lo[3], c_hi[3];
    for (int d = 0; d < 3; ++d) {
      c_lo[d] = std::max(0, (int)std::floor((x0[d] - radius) * n[d]) - 1);
      c_hi[d] = std::min(n[d] - 1, (int)std::floor((x0[d] + radius) * n[d]) + 1);
    }
    for (int ck = c_lo[2]; ck <= c_hi[2]; ++ck)
      for (int cj = c_lo[1]; cj <= c_hi[1]; ++cj)
        for (int ci = c_lo[0]; ci <= c_hi[0]; ++ci) {
          const int cc[3] = {ci, cj, ck};
          bool overlap = false, centre_in_cell = true;
          for (int o = 0; o < 8; ++o) {
            double d2 = 0.0;
            for (int d = 0; d < 3; ++d) {
              const double dx = h[d] * (cc[d] + ((o >> d) & 1)) - x0[d];
              d2 += dx * dx;
            }
            overlap = overlap || (std::sqrt(d2) < radius);
          }
          for (int d = 0; d < 3; ++d) centre_in_cell = centre_in_cell && (h[d] * cc[d] <= x0[d]) && (x0[d] <= h[d] * (cc[d] + 1));
          if (!(overlap || centre_in_cell)) continue;
          for (int a = 0; a < 8; ++a) {
            const int v[3] = {ci + (a & 1), cj + ((a >> 1) & 1), ck + ((a >> 2) & 1)};
            if (!(v[0] > 0 && v[0] < n[0] && v[1] > 0 && v[1] < n[1] && v[2] > 0 && v[2] < n[2])) continue;
            double local = 0.0;
            for (int q = 0; q < 8; ++q) {
              double d2 = 0.0, phi = 1.0;
              for (int d = 0; d < 3; ++d) {
                const double xh = gp[(q >> d) & 1];
                const double dx = h[d] * (xh + cc[d]) - x0[d];
                d2 += dx * dx;
                phi *= ((a >> d) & 1) ? xh : (1.0 - xh);
              }
              if (std::sqrt(d2) / radius < 1.0) local += phi * 0.125 * cell_volume * normalisation;
            }
            r.coeffRef(lin(v[0], v[1], v[2])) += local;
          }
        }
    return r;
  }

 protected:
  MeasurementParameters params;
};

// ------------------------------------------------------------------------------------------------
// IntergridOperator (intergrid/intergrid_operator.hh:40-144, intergrid_operator_linear.hh)
// ------------------------------------------------------------------------------------------------
class IntergridOperator {
 public:
  explicit IntergridOperator(const std::shared_ptr<Lattice> lattice_) : lattice(lattice_) {}
  virtual ~IntergridOperator() = default;
  /** attach the device hierarchy of the operator whose levels this transfer connects */
  void bind(const LinearOperator &fine_op) {
    dev = fine_op.hierarchy(fine_op.get_level() + 2);
    level = fine_op.get_level();
  }
  virtual void restrict(const Eigen::VectorXd &x, Eigen::VectorXd &x_coarse) {
    need();
    mgmc_host::check(mgmc_restrict(dev->ctx, level, x.data(), x_coarse.data()), "IntergridOperator::restrict");
  }
  virtual void prolongate_add(const double alpha, const Eigen::VectorXd &x_coarse, Eigen::VectorXd &x) {
    need();
    mgmc_host::check(mgmc_prolongate_add(dev->ctx, level, alpha, x_coarse.data(), x.data()), "IntergridOperator::prolongate_add");
  }

 protected:
  void need() {
    if (!dev) {  // standalone use (test_intergrid.hh): a prior hierarchy on this lattice carries the transfer
      OperatorData d;
      Eigen::VectorXi s = lattice->shape();
      if (lattice->dim() != 2 && lattice->dim() != 3) {
        std::cout << "ERROR: the device path supports dim = 2 and dim = 3 only" << std::endl;
        exit(-1);
      }
      d.nx = s[0];
      d.ny = s[1];
      d.nz = (lattice->dim() == 3) ? s[2] : 0;
      MultigridParameters p;
      p.nlevel = 2;
      dev = std::make_shared<DeviceHierarchy>(d, p, 0);
      level = 0;
    }
  }
  const std::shared_ptr<Lattice> lattice;
  std::shared_ptr<DeviceHierarchy> dev;
  int level = 0;
};
class IntergridOperatorLinear : public IntergridOperator {
 public:
  explicit IntergridOperatorLinear(const std::shared_ptr<Lattice> lattice_) : IntergridOperator(lattice_) {}
};
class IntergridOperatorFactory {
 public:
  virtual ~IntergridOperatorFactory() = default;
  virtual std::shared_ptr<IntergridOperator> get(std::shared_ptr<Lattice> lattice) = 0;
};
class IntergridOperatorLinearFactory : public IntergridOperatorFactory {
 public:
  std::shared_ptr<IntergridOperator> get(std::shared_ptr<Lattice> lattice) override { return std::make_shared<IntergridOperatorLinear>(lattice); }
};

// ------------------------------------------------------------------------------------------------
// Smoothers (smoother/smoother.hh, sor_smoother.hh, ssor_smoother.hh)
// ------------------------------------------------------------------------------------------------
enum Direction { forward = 1, backward = 2 };

class Smoother {
 public:
  explicit Smoother(const std::shared_ptr<LinearOperator> linear_operator_) : linear_operator(linear_operator_) {}
  virtual ~Smoother() = default;
  virtual void apply(const Eigen::VectorXd &b, Eigen::VectorXd &x) const = 0;

 protected:
  const std::shared_ptr<LinearOperator> linear_operator;
};
class SORSmoother : public Smoother {
 public:
  SORSmoother(const std::shared_ptr<LinearOperator> op, const double omega_, const int nsmooth_, const Direction direction_)
      : Smoother(op), omega(omega_), nsmooth(nsmooth_), direction(direction_) {}
  void apply(const Eigen::VectorXd &b, Eigen::VectorXd &x) const override {
    const int l = linear_operator->get_level();
    mgmc_host::check(mgmc_smoother_apply(linear_operator->hierarchy(l + 1)->ctx, l, MGMC_SMOOTHER_SOR, direction, omega, nsmooth, b.data(), x.data()), "SORSmoother::apply");
  }

 protected:
  const double omega;
  const int nsmooth;
  const Direction direction;
};
class SSORSmoother : public Smoother {
 public:
  SSORSmoother(const std::shared_ptr<LinearOperator> op, const double omega_, const int nsmooth_) : Smoother(op), omega(omega_), nsmooth(nsmooth_) {}
  void apply(const Eigen::VectorXd &b, Eigen::VectorXd &x) const override {
    const int l = linear_operator->get_level();
    mgmc_host::check(mgmc_smoother_apply(linear_operator->hierarchy(l + 1)->ctx, l, MGMC_SMOOTHER_SSOR, MGMC_FORWARD, omega, nsmooth, b.data(), x.data()), "SSORSmoother::apply");
  }

 protected:
  const double omega;
  const int nsmooth;
};
class SmootherFactory {
 public:
  virtual ~SmootherFactory() = default;
  virtual std::shared_ptr<Smoother> get(std::shared_ptr<LinearOperator> linear_operator) = 0;
};
class SORSmootherFactory : public SmootherFactory {
 public:
  SORSmootherFactory(const double omega_, const int nsmooth_, const Direction direction_) : omega(omega_), nsmooth(nsmooth_), direction(direction_) {}
  std::shared_ptr<Smoother> get(std::shared_ptr<LinearOperator> op) override { return std::make_shared<SORSmoother>(op, omega, nsmooth, direction); }

 private:
  const double omega;
  const int nsmooth;
  const Direction direction;
};
class SSORSmootherFactory : public SmootherFactory {
 public:
  SSORSmootherFactory(const double omega_, const int nsmooth_) : omega(omega_), nsmooth(nsmooth_) {}
  std::shared_ptr<Smoother> get(std::shared_ptr<LinearOperator> op) override { return std::make_shared<SSORSmoother>(op, omega, nsmooth); }

 private:
  const double omega;
  const int nsmooth;
};

// ------------------------------------------------------------------------------------------------
// Samplers (sampler/sampler.hh, sor_sampler.hh, ssor_sampler.hh, multigridmc_sampler.hh)
// ------------------------------------------------------------------------------------------------
class Sampler {
 public:
  Sampler(const std::shared_ptr<LinearOperator> linear_operator_, std::mt19937_64 &rng_) : linear_operator(linear_operator_), rng(rng_) {}
  virtual ~Sampler() = default;
  virtual void apply(const Eigen::VectorXd &f, Eigen::VectorXd &x) const = 0;
  std::shared_ptr<LinearOperator> get_linear_operator() const { return linear_operator; }
  virtual void fix_rhs(const Eigen::VectorXd &) {}
  virtual void unfix_rhs() {}
  /** device-resident hot loop of measure_sampling_time (driver_mgmc.cc:73-77): nsamples x apply with
   *  z_k = sample_vector . x evaluated on the device; x is the chain state in / out */
  virtual void sample_series(const Eigen::VectorXd &f, Eigen::VectorXd &x, const Eigen::SparseVector<double> &sample_vector, std::vector<double> &data) const {
    for (size_t k = 0; k < data.size(); ++k) {
      apply(f, x);
      data[k] = sample_vector.dot(x);
    }
  }

 protected:
  const std::shared_ptr<LinearOperator> linear_operator;
  std::mt19937_64 &rng;
};

/** SOR / SSOR Gibbs samplers on one level (sor_sampler.cc:37-58, ssor_sampler.cc:9-16) */
class SORSampler : public Sampler {
 public:
  SORSampler(const std::shared_ptr<LinearOperator> op, std::mt19937_64 &rng_, const double omega_, const unsigned int nsmooth_, const Direction direction_)
      : Sampler(op, rng_), omega(omega_), nsmooth(nsmooth_), direction(direction_), kind(MGMC_SMOOTHER_SOR) {
    init();
  }
  void apply(const Eigen::VectorXd &f, Eigen::VectorXd &x) const override {
    mgmc_host::check(mgmc_set_philox_position(dev->ctx, sample_index++, 0), "set_philox_position");
    mgmc_host::check(mgmc_sampler_apply(dev->ctx, linear_operator->get_level(), kind, direction, omega, (int)nsmooth, f.data(), x.data()), "Sampler::apply");
  }

 protected:
  SORSampler(const std::shared_ptr<LinearOperator> op, std::mt19937_64 &rng_, const double omega_, const unsigned int nsmooth_, int kind_)
      : Sampler(op, rng_), omega(omega_), nsmooth(nsmooth_), direction(forward), kind(kind_) {
    init();
  }
  void init() {
    MultigridParameters p;
    p.nlevel = (unsigned int)linear_operator->get_level() + 1;
    dev = std::make_shared<DeviceHierarchy>(linear_operator->get_data(), p, rng());  // one draw = Philox key
  }
  const double omega;
  const unsigned int nsmooth;
  const Direction direction;
  const int kind;
  std::shared_ptr<DeviceHierarchy> dev;
  mutable uint32_t sample_index = 0;
};
class SSORSampler : public SORSampler {
 public:
  SSORSampler(const std::shared_ptr<LinearOperator> op, std::mt19937_64 &rng_, const double omega_, const unsigned int nsmooth_)
      : SORSampler(op, rng_, omega_, nsmooth_, MGMC_SMOOTHER_SSOR) {}
};

/** MultigridMCSampler (sampler/multigridmc_sampler.hh:34-72, multigridmc_sampler.cc:8-138) */
class MultigridMCSampler : public Sampler {
 public:
  MultigridMCSampler(std::shared_ptr<LinearOperator> op, std::mt19937_64 &rng_, const MultigridParameters params_, const CholeskyParameters cholesky_params_)
      : Sampler(op, rng_), params(params_), cholesky_params(cholesky_params_) {
    if (params.verbose > 0) std::cout << "Setting up Multilevel MC sampler " << std::endl;
    dev = std::make_shared<DeviceHierarchy>(op->get_data(), params, rng());  // one draw = Philox key
    if (params.verbose > 0) {
      std::shared_ptr<Lattice> lattice = op->get_lattice();
      for (unsigned int level = 0; level < params.nlevel; ++level) {
        std::cout << "  level " << level << " lattice : " << lattice->get_info() << std::endl;
        if (level + 1 < params.nlevel) lattice = lattice->get_coarse_lattice();
      }
    }
  }
  void apply(const Eigen::VectorXd &f, Eigen::VectorXd &x) const override {
    mgmc_host::check(mgmc_sampler_mgmc_apply(dev->ctx, rhs_fixed ? nullptr : f.data(), x.data()), "MultigridMCSampler::apply");
  }
  void fix_rhs(const Eigen::VectorXd &f) override {
    mgmc_host::check(mgmc_set_rhs(dev->ctx, f.data()), "MultigridMCSampler::fix_rhs");
    rhs_fixed = true;
  }
  void unfix_rhs() override { rhs_fixed = false; }
  void sample_series(const Eigen::VectorXd &f, Eigen::VectorXd &x, const Eigen::SparseVector<double> &sample_vector, std::vector<double> &data) const override {
    std::vector<int64_t> idx;
    std::vector<double> val;
    for (auto &e : sample_vector.entries()) {
      idx.push_back((int64_t)e.first);
      val.push_back(e.second);
    }
    mgmc_host::check(mgmc_set_qoi(dev->ctx, (int64_t)idx.size(), idx.data(), val.data()), "mgmc_set_qoi");
    if (!rhs_fixed) mgmc_host::check(mgmc_set_rhs(dev->ctx, f.data()), "mgmc_set_rhs");
    mgmc_host::check(mgmc_set_state(dev->ctx, x.data()), "mgmc_set_state");
    mgmc_host::check(mgmc_sample(dev->ctx, (int64_t)data.size(), data.data()), "mgmc_sample");
    mgmc_host::check(mgmc_get_state(dev->ctx, x.data()), "mgmc_get_state");
  }
  /** running mean / second moment fields over nsamples (posterior_statistics, driver_mgmc.cc:146-151) */
  void sample_moments(const Eigen::VectorXd &f, Eigen::VectorXd &x, unsigned int nsamples, Eigen::VectorXd &mean, Eigen::VectorXd &second) const {
    mgmc_host::check(mgmc_set_rhs(dev->ctx, f.data()), "mgmc_set_rhs");
    mgmc_host::check(mgmc_set_state(dev->ctx, x.data()), "mgmc_set_state");
    mgmc_host::check(mgmc_sample_moments(dev->ctx, nsamples, mean.data(), second.data()), "mgmc_sample_moments");
    mgmc_host::check(mgmc_get_state(dev->ctx, x.data()), "mgmc_get_state");
  }
  mgmc_ctx *context() const { return dev->ctx; }

 protected:
  const MultigridParameters params;
  const CholeskyParameters cholesky_params;
  std::shared_ptr<DeviceHierarchy> dev;
  bool rhs_fixed = false;
};

class SamplerFactory {
 public:
  virtual ~SamplerFactory() = default;
  virtual std::shared_ptr<Sampler> get(std::shared_ptr<LinearOperator> linear_operator) = 0;
};
class SSORSamplerFactory : public SamplerFactory {
 public:
  SSORSamplerFactory(std::mt19937_64 &rng_, const double omega_, const int nsmooth_) : rng(rng_), omega(omega_), nsmooth(nsmooth_) {}
  std::shared_ptr<Sampler> get(std::shared_ptr<LinearOperator> op) override { return std::make_shared<SSORSampler>(op, rng, omega, nsmooth); }

 protected:
  std::mt19937_64 &rng;
  const double omega;
  const int nsmooth;
};
class SORSamplerFactory : public SamplerFactory {
 public:
  SORSamplerFactory(std::mt19937_64 &rng_, const double omega_, const int nsmooth_, const Direction direction_) : rng(rng_), omega(omega_), nsmooth(nsmooth_), direction(direction_) {}
  std::shared_ptr<Sampler> get(std::shared_ptr<LinearOperator> op) override { return std::make_shared<SORSampler>(op, rng, omega, nsmooth, direction); }

 protected:
  std::mt19937_64 &rng;
  const double omega;
  const int nsmooth;
  const Direction direction;
};

/** CholeskySampler (sampler/cholesky_sampler.hh:27-146, cholesky_sampler.cc:9-38): exact sampler x = A^{-1} f + L^{-T} xi for a
 *  small operator (<= 4096 unknowns: the dense factor, low-rank term folded in, is resident on the device -- the coarse
 *  phase of the V-cycle, csrc/tail.cuh).  Sparse / Dense name the reference's two factorisation back ends; here both are
 *  the same device factor.  xi is drawn from the Philox stream keyed by one draw of rng (as for the other samplers). */
class CholeskySampler : public Sampler {
 public:
  CholeskySampler(const std::shared_ptr<LinearOperator> op, std::mt19937_64 &rng_) : Sampler(op, rng_) {
    MultigridParameters p;
    p.nlevel = (unsigned int)op->get_level() + 1;  // the operator is the coarsest level of this hierarchy
    p.coarse_solver = "Cholesky";
    dev = std::make_shared<DeviceHierarchy>(op->get_data(), p, rng());
  }
  void apply(const Eigen::VectorXd &f, Eigen::VectorXd &x) const override {
    mgmc_host::check(mgmc_set_philox_position(dev->ctx, sample_index++, 0), "set_philox_position");
    mgmc_host::check(mgmc_coarse_sample(dev->ctx, f.data(), x.data()), "CholeskySampler::apply");
  }

 protected:
  std::shared_ptr<DeviceHierarchy> dev;
  mutable uint32_t sample_index = 0;
};
class SparseCholeskySampler : public CholeskySampler {
 public:
  using CholeskySampler::CholeskySampler;
};
class DenseCholeskySampler : public CholeskySampler {
 public:
  using CholeskySampler::CholeskySampler;
};
/** factories (sampler/cholesky_sampler.hh:151-196) */
class SparseCholeskySamplerFactory : public SamplerFactory {
 public:
  explicit SparseCholeskySamplerFactory(std::mt19937_64 &rng_) : rng(rng_) {}
  std::shared_ptr<Sampler> get(std::shared_ptr<LinearOperator> op) override { return std::make_shared<SparseCholeskySampler>(op, rng); }

 protected:
  std::mt19937_64 &rng;
};
class DenseCholeskySamplerFactory : public SamplerFactory {
 public:
  explicit DenseCholeskySamplerFactory(std::mt19937_64 &rng_) : rng(rng_) {}
  std::shared_ptr<Sampler> get(std::shared_ptr<LinearOperator> op) override { return std::make_shared<DenseCholeskySampler>(op, rng); }

 protected:
  std::mt19937_64 &rng;
};

// ------------------------------------------------------------------------------------------------
// Preconditioner / solvers (preconditioner/multigrid_preconditioner.hh, solver/loop_solver.hh)
// ------------------------------------------------------------------------------------------------
class Preconditioner {
 public:
  explicit Preconditioner(std::shared_ptr<LinearOperator> linear_operator_) : linear_operator(linear_operator_) {}
  virtual ~Preconditioner() = default;
  virtual void apply(const Eigen::VectorXd &b, Eigen::VectorXd &x) = 0;

 protected:
  std::shared_ptr<LinearOperator> linear_operator;
};
class MultigridPreconditioner : public Preconditioner {
 public:
  MultigridPreconditioner(std::shared_ptr<LinearOperator> op, const MultigridParameters params_) : Preconditioner(op), params(params_) {
    if (params.coarse_solver != "Cholesky")  // multigrid_preconditioner.cc:41-45
      std::cout << "WARNING: ignoring coarse solver setting '" << params.coarse_solver << "', using Choleksy." << std::endl;
    dev = std::make_shared<DeviceHierarchy>(op->get_data(), params, 0);
    if (params.verbose > 0) {
      std::shared_ptr<Lattice> lattice = op->get_lattice();
      for (unsigned int level = 0; level < params.nlevel; ++level) {
        std::cout << "  level " << level << " lattice : " << lattice->get_info() << std::endl;
        if (level + 1 < params.nlevel) lattice = lattice->get_coarse_lattice();
      }
    }
  }
  void apply(const Eigen::VectorXd &b, Eigen::VectorXd &x) override { mgmc_host::check(mgmc_mgprec_apply(dev->ctx, b.data(), x.data()), "MultigridPreconditioner::apply"); }
  mgmc_ctx *context() const { return dev->ctx; }

 protected:
  const MultigridParameters params;
  std::shared_ptr<DeviceHierarchy> dev;
};

class LinearSolver {
 public:
  explicit LinearSolver(std::shared_ptr<LinearOperator> linear_operator_) : linear_operator(linear_operator_) {}
  virtual ~LinearSolver() = default;
  virtual void apply(const Eigen::VectorXd &b, Eigen::VectorXd &x) = 0;

 protected:
  std::shared_ptr<LinearOperator> linear_operator;
};

/** LoopSolver (solver/loop_solver.cc:9-53): preconditioned Richardson iteration, same printed history */
class LoopSolver : public LinearSolver {
 public:
  LoopSolver(std::shared_ptr<LinearOperator> op, std::shared_ptr<Preconditioner> preconditioner_, const IterativeSolverParameters params_)
      : LinearSolver(op), preconditioner(preconditioner_), params(params_) {}
  void apply(const Eigen::VectorXd &b, Eigen::VectorXd &x) override {
    auto mg = std::dynamic_pointer_cast<MultigridPreconditioner>(preconditioner);
    if (!mg) {
      std::cout << "ERROR: LoopSolver on the device path needs a MultigridPreconditioner" << std::endl;
      exit(-1);
    }
    const double r0_nrm = b.norm();
    if (params.verbose >= 2) printf("Initial residual ||r_0|| =  %12.4f\n", r0_nrm);
    std::vector<double> history(params.maxiter + 1);
    int nhist = 0, niter = 0, converged = 0;
    mgmc_host::check(mgmc_loop_solve(mg->context(), b.data(), x.data(), params.rtol, params.atol, (int)params.maxiter, history.data(), &nhist, &niter, &converged),
                     "LoopSolver::apply");
    if (params.verbose >= 2) {
      printf("%5s   %8s   %12s   %6s\n", "iter", "||r||", "||r||/||r_0||", "rho");
      double rold = r0_nrm;
      for (int k = 0; k < nhist; ++k) {
        printf("%5d   %8.3e   %12.3e   %6.3f\n", k, history[k], history[k] / r0_nrm, history[k] / rold);
        rold = history[k];
      }
    }
    residual_history.assign(history.begin(), history.begin() + nhist);
    if (params.verbose >= 1) {
      if (converged) printf("Solver converged after %5d iterations\n||r|| = %8.3e, ||r||/||r_0|| = %8.3e\n", niter, history[nhist - 1], history[nhist - 1] / r0_nrm);
      else printf("Solver failed to converge after %5d iterations\n", params.maxiter);
    }
  }
  std::vector<double> residual_history;

 protected:
  std::shared_ptr<Preconditioner> preconditioner;
  const IterativeSolverParameters params;
};

/** linear solver factory base class (solver/linear_solver.hh:43-48) */
class LinearSolverFactory {
 public:
  virtual ~LinearSolverFactory() = default;
  virtual std::shared_ptr<LinearSolver> get(std::shared_ptr<LinearOperator> linear_operator) = 0;
};

/** CholeskySolver (solver/cholesky_solver.hh:21-52, cholesky_solver.cc:8-41): direct solve A x = b for a small operator
 *  (<= 4096 unknowns).  The reference factorises A_0 and treats the low-rank term by the Woodbury identity; the device
 *  factor has the term folded in -- the same x to rounding. */
class CholeskySolver : public LinearSolver {
 public:
  explicit CholeskySolver(std::shared_ptr<LinearOperator> op) : LinearSolver(op) {
    MultigridParameters p;
    p.nlevel = (unsigned int)op->get_level() + 1;
    p.coarse_solver = "Cholesky";
    dev = std::make_shared<DeviceHierarchy>(op->get_data(), p, 0);
  }
  void apply(const Eigen::VectorXd &b, Eigen::VectorXd &x) override {
    mgmc_host::check(mgmc_coarse_solve(dev->ctx, b.data(), x.data()), "CholeskySolver::apply");
  }

 protected:
  std::shared_ptr<DeviceHierarchy> dev;
};
/** CholeskySolverFactory (solver/cholesky_solver.hh:57-68) */
class CholeskySolverFactory : public LinearSolverFactory {
 public:
  std::shared_ptr<LinearSolver> get(std::shared_ptr<LinearOperator> op) override { return std::make_shared<CholeskySolver>(op); }
};

// ---- LinearOperator members that need the solver classes ----
inline std::vector<Eigen::VectorXd> LinearOperator::prior_solve_columns(const std::vector<Eigen::VectorXd> &rhs) const {
  // A_0^{-1} rhs_k by multigrid-preconditioned Richardson on the prior operator
  OperatorData prior = *data;
  prior.B_rows.clear();
  prior.B_cols.clear();
  prior.B_vals.clear();
  prior.Sigma.clear();
  MultigridParameters p;
  unsigned int n = std::min(data->nx, data->ny), nl = 1;
  unsigned int nx = data->nx, ny = data->ny, nz = data->nz;
  if (nz) {
    // Lattice3d: coarsen until the dense coarse factor fits (at most 4096 unknowns: 16^3 cells)
    while (nx % 2 == 0 && ny % 2 == 0 && nz % 2 == 0 && std::min(std::min(nx, ny), nz) / 2 >= 2 && (unsigned long long)(nx - 1) * (ny - 1) * (nz - 1) > 4096ull) {
      nx /= 2;
      ny /= 2;
      nz /= 2;
      ++nl;
    }
  }
  while (!nz && nx % 2 == 0 && ny % 2 == 0 && std::min(nx, ny) / 2 >= 16 && n > 32) {
    nx /= 2;
    ny /= 2;
    n /= 2;
    ++nl;
  }
  p.nlevel = nl;
  p.npresmooth = p.npostsmooth = 2;
  DeviceHierarchy H(prior, p, 0);
  std::vector<Eigen::VectorXd> out;
  for (const Eigen::VectorXd &b : rhs) {
    Eigen::VectorXd x(b.size());
    int nh = 0, it = 0, conv = 0;
    mgmc_host::check(mgmc_loop_solve(H.ctx, b.data(), x.data(), 1e-13, 1e300, 200, nullptr, &nh, &it, &conv), "prior solve");
    out.push_back(x);
  }
  return out;
}

inline Eigen::VectorXd LinearOperator::mean(const Eigen::VectorXd &xbar, const Eigen::VectorXd &y) const {
  if (m_lowrank == 0) return xbar;
  const unsigned int n = get_ndof(), m = m_lowrank;
  std::vector<Eigen::VectorXd> Bcols(m, Eigen::VectorXd(n));
  for (auto &c : Bcols) c.setZero();
  for (size_t e = 0; e < data->B_vals.size(); ++e) Bcols[data->B_cols[e]][data->B_rows[e]] = data->B_vals[e];
  std::vector<Eigen::VectorXd> Bbar = prior_solve_columns(Bcols);
  // S = Sigma + B^T Bbar, solve S z = y - B^T xbar (dense m x m, Gaussian elimination with pivoting)
  std::vector<double> S(m * m), rhs(m);
  for (unsigned int a = 0; a < m; ++a) {
    for (unsigned int b = 0; b < m; ++b) S[a * m + b] = Bcols[a].dot(Bbar[b]) + (a == b ? data->Sigma[a] : 0.0);
    rhs[a] = y[a] - Bcols[a].dot(xbar);
  }
  for (unsigned int c = 0; c < m; ++c) {
    unsigned int piv = c;
    for (unsigned int r = c + 1; r < m; ++r)
      if (std::fabs(S[r * m + c]) > std::fabs(S[piv * m + c])) piv = r;
    for (unsigned int k = 0; k < m; ++k) std::swap(S[piv * m + k], S[c * m + k]);
    std::swap(rhs[piv], rhs[c]);
    for (unsigned int r = c + 1; r < m; ++r) {
      const double f = S[r * m + c] / S[c * m + c];
      for (unsigned int k = c; k < m; ++k) S[r * m + k] -= f * S[c * m + k];
      rhs[r] -= f * rhs[c];
    }
  }
  std::vector<double> z(m);
  for (int r = (int)m - 1; r >= 0; --r) {
    double s = rhs[r];
    for (unsigned int k = r + 1; k < m; ++k) s -= S[r * m + k] * z[k];
    z[r] = s / S[r * m + r];
  }
  Eigen::VectorXd out = xbar;
  for (unsigned int k = 0; k < m; ++k) out += z[k] * Bbar[k];
  return out;
}

inline void LinearOperator::observed_mean_and_variance(const Eigen::VectorXd &xbar, const Eigen::VectorXd &y, const Eigen::SparseVector<double> &b_obs, double &mean_,
                                                       double &variance) const {
  const unsigned int n = get_ndof(), m = m_lowrank;
  std::vector<Eigen::VectorXd> cols(m + 1, Eigen::VectorXd(n));
  for (auto &c : cols) c.setZero();
  for (size_t e = 0; e < data->B_vals.size(); ++e) cols[data->B_cols[e]][data->B_rows[e]] = data->B_vals[e];
  for (auto &e : b_obs.entries()) cols[m][e.first] = e.second;
  std::vector<Eigen::VectorXd> sol = prior_solve_columns(cols);
  const Eigen::VectorXd &bbar = sol[m];
  mean_ = b_obs.dot(xbar);
  variance = b_obs.dot(bbar);
  if (m == 0) return;
  std::vector<double> S(m * m), Sinv(m * m, 0.0);
  for (unsigned int a = 0; a < m; ++a)
    for (unsigned int b = 0; b < m; ++b) S[a * m + b] = cols[a].dot(sol[b]) + (a == b ? data->Sigma[a] : 0.0);
  // invert S (Gauss-Jordan)
  for (unsigned int i = 0; i < m; ++i) Sinv[i * m + i] = 1.0;
  for (unsigned int c = 0; c < m; ++c) {
    unsigned int piv = c;
    for (unsigned int r = c + 1; r < m; ++r)
      if (std::fabs(S[r * m + c]) > std::fabs(S[piv * m + c])) piv = r;
    for (unsigned int k = 0; k < m; ++k) {
      std::swap(S[piv * m + k], S[c * m + k]);
      std::swap(Sinv[piv * m + k], Sinv[c * m + k]);
    }
    const double inv = 1.0 / S[c * m + c];
    for (unsigned int k = 0; k < m; ++k) {
      S[c * m + k] *= inv;
      Sinv[c * m + k] *= inv;
    }
    for (unsigned int r = 0; r < m; ++r) {
      if (r == c) continue;
      const double f = S[r * m + c];
      for (unsigned int k = 0; k < m; ++k) {
        S[r * m + k] -= f * S[c * m + k];
        Sinv[r * m + k] -= f * Sinv[c * m + k];
      }
    }
  }
  std::vector<double> BTbbar(m), resid(m);
  for (unsigned int a = 0; a < m; ++a) {
    BTbbar[a] = cols[a].dot(bbar);
    resid[a] = y[a] - cols[a].dot(xbar);
  }
  for (unsigned int a = 0; a < m; ++a)
    for (unsigned int b = 0; b < m; ++b) {
      mean_ += BTbbar[a] * Sinv[a * m + b] * resid[b];
      variance -= BTbbar[a] * Sinv[a * m + b] * BTbbar[b];
    }
}

// ------------------------------------------------------------------------------------------------
// Statistics for a scalar time series (auxilliary/statistics.cc:4-79): running averages of the lagged
// products with window k_max, auto-covariance C(k) and integrated autocorrelation time
// ------------------------------------------------------------------------------------------------
class Statistics {
 public:
  Statistics(const std::string label_, const unsigned int k_max_) : label(label_), k_max(k_max_), n_samples(0), avg(0.0), avg2(0.0) {}
  void record_sample(const double Q) {
    n_samples++;
    avg += (Q - avg) / (1.0 * n_samples);
    avg2 += (Q * Q - avg2) / (1.0 * n_samples);
    Q_k.push_front(Q);
    if (Q_k.size() > k_max) Q_k.pop_back();
    for (unsigned int k = 0; k < Q_k.size(); ++k) {
      const unsigned int N_k = n_samples - k;
      if (N_k == 1) S_k.push_back(Q_k[0] * Q_k[k]);
      else S_k[k] += (Q_k[0] * Q_k[k] - S_k[k]) / (1.0 * N_k);
    }
  }
  double variance() const { return 1.0 * n_samples / (n_samples - 1.0) * (avg2 - avg * avg); }
  double average() const { return avg; }
  std::vector<double> auto_covariance() const {
    std::vector<double> c;
    for (double s : S_k) c.push_back(s - avg * avg);
    return c;
  }
  double tau_int() const {
    const std::vector<double> C = auto_covariance();
    double tau = 1.0;
    const unsigned int kmax = (unsigned int)C.size();
    for (unsigned int k = 1; k < kmax; ++k) tau += 2 * (1. - k / (1.0 * kmax)) * C[k] / C[0];
    return tau;
  }
  unsigned int samples() const { return n_samples; }

 private:
  const std::string label;
  const unsigned int k_max;
  unsigned int n_samples;
  double avg, avg2;
  std::deque<double> Q_k;
  std::vector<double> S_k;
};

/** legacy-VTK STRUCTURED_POINTS writer for vertex fields (auxilliary/vtk_writer2d.cc) */
class VTKWriter2d {
 public:
  VTKWriter2d(const std::string filename_, const std::shared_ptr<Lattice> lattice_, const int verbose = 0) : filename(filename_), lattice(lattice_) { (void)verbose; }
  void add_state(const Eigen::VectorXd &phi, const std::string label) { states.push_back({label, phi}); }
  void write() const {
    const Eigen::VectorXi s = lattice->shape();
    if (lattice->dim() == 3) return write3d();
    const int nx = s[0], ny = s[1];
    std::ofstream out(filename);
    out << "# vtk DataFile Version 2.0\nSample state\nASCII\nDATASET STRUCTURED_POINTS\n";
    out << "DIMENSIONS " << nx + 1 << " " << ny + 1 << " 1\nORIGIN 0.0 0.0 0.0\nSPACING " << 1. / nx << " " << 1. / ny << " 0\n\nPOINT_DATA " << (nx + 1) * (ny + 1) << "\n";
    for (auto &st : states) {
      out << "SCALARS " << st.first << " double 1\nLOOKUP_TABLE default\n";
      for (int j = 0; j <= ny; ++j)
        for (int i = 0; i <= nx; ++i) out << ((i > 0 && i < nx && j > 0 && j < ny) ? st.second[(j - 1) * (nx - 1) + (i - 1)] : 0.0) << "\n";
    }
  }

 protected:
  /** VTKWriter3d::write (auxilliary/vtk_writer3d.cc:8-58): same header, origin and |data| < 1e-20 -> 0 clamp */
  void write3d() const {
    const Eigen::VectorXi s = lattice->shape();
    const int nx = s[0], ny = s[1], nz = s[2];
    std::ofstream out(filename);
    out << "# vtk DataFile Version 2.0\nSample state\nASCII\nDATASET STRUCTURED_POINTS\n";
    out << "DIMENSIONS " << nx + 1 << " " << ny + 1 << " " << nz + 1 << "\nORIGIN -0.5 -0.5 -5.0\nSPACING " << 1. / nx << " " << 1. / ny << " " << 1. / nz << "\n\nPOINT_DATA "
        << (nx + 1) * (ny + 1) * (nz + 1) << "\n";
    for (auto &st : states) {
      out << "SCALARS " << st.first << " double 1\nLOOKUP_TABLE default\n";
      for (int k = 0; k <= nz; ++k)
        for (int j = 0; j <= ny; ++j)
          for (int i = 0; i <= nx; ++i) {
            double data = 0.0;
            if (i > 0 && i < nx && j > 0 && j < ny && k > 0 && k < nz) data = st.second[((k - 1) * (ny - 1) + (j - 1)) * (nx - 1) + (i - 1)];
            if (std::fabs(data) < 1.0E-20) data = 0.0;
            out << data << "\n";
          }
    }
  }
  const std::string filename;
  const std::shared_ptr<Lattice> lattice;
  std::vector<std::pair<std::string, Eigen::VectorXd>> states;
};
/** VTKWriter3d (auxilliary/vtk_writer3d.hh): the writer above picks the format from the lattice */
class VTKWriter3d : public VTKWriter2d {
 public:
  using VTKWriter2d::VTKWriter2d;
};
#endif
