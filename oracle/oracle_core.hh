// ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// CPU restatement (plain C++17, no Eigen) of the MultigridMC sampling hot path of
// nilsfriess/MultigridMC.  Every function cites the reference file:line it follows
// (paths relative to /root/reference/src).  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may use this code.
//
// Parity status: the reference holds NO golden vectors; the oracle is
// pinned against every known-answer / identity test the reference's own test-suite holds for
// this path (tests/test_oracle_*.py restate them one by one).  Bit-level chain reproduction
// against a compiled reference is "parity unpinned": the reference needs Eigen 3.4 +
// libconfig++, neither of which exists in this image (DESIGN.md section "Oracle").
#ifndef MGMC_ORACLE_CORE_HH
#define MGMC_ORACLE_CORE_HH

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>

namespace orc {

typedef std::vector<double> Vec;

// ---------------------------------------------------------------------------------------------
// Lattice (lattice/lattice.hh:18-129, lattice1d.hh, lattice2d.hh:43-235, lattice3d.hh:60-286)
// Interior vertices, lexicographic: ell = (k-1)(nx-1)(ny-1) + (j-1)(nx-1) + (i-1)
// ---------------------------------------------------------------------------------------------
struct Lattice {
  int dim;
  int n[3];  // cells per dimension (unused dims = 1... but stored as 2 so that n-1 = 1)
  Lattice() : dim(0) { n[0] = n[1] = n[2] = 2; }
  Lattice(int dim_, const int *n_) : dim(dim_) {
    n[0] = n[1] = n[2] = 2;
    for (int d = 0; d < dim; ++d) n[d] = n_[d];
  }
  long Ncell() const {
    long c = 1;
    for (int d = 0; d < dim; ++d) c *= n[d];
    return c;
  }
  long Nvertex() const {
    long c = 1;
    for (int d = 0; d < dim; ++d) c *= (n[d] - 1);
    return c;
  }
  // lattice.hh:31-40
  double cell_volume() const {
    double v = 1.0;
    for (int d = 0; d < dim; ++d) v /= n[d];
    return v;
  }
  // lattice2d.hh:86-93, lattice3d.hh:113-119 (Euclidean index starts at 1)
  void vertex_l2e(long ell, int *idx) const {
    for (int d = 0; d < dim; ++d) {
      idx[d] = int(ell % (n[d] - 1)) + 1;
      ell /= (n[d] - 1);
    }
  }
  // lattice2d.hh:99-106
  long vertex_e2l(const int *idx) const {
    long ell = 0, stride = 1;
    for (int d = 0; d < dim; ++d) {
      assert(idx[d] > 0 && idx[d] < n[d]);
      ell += stride * (idx[d] - 1);
      stride *= (n[d] - 1);
    }
    return ell;
  }
  void cell_l2e(long ell, int *idx) const {  // lattice2d.hh:62-69
    for (int d = 0; d < dim; ++d) {
      idx[d] = int(ell % n[d]);
      ell /= n[d];
    }
  }
  long cell_e2l(const int *idx) const {  // lattice2d.hh:75-78
    long ell = 0, stride = 1;
    for (int d = 0; d < dim; ++d) {
      ell += stride * idx[d];
      stride *= n[d];
    }
    return ell;
  }
  long shift_cellidx(long ell, const int *shift) const {  // lattice2d.hh:115-121
    int idx[3];
    cell_l2e(ell, idx);
    for (int d = 0; d < dim; ++d) idx[d] += shift[d];
    return cell_e2l(idx);
  }
  // lattice2d.hh:146-154: true iff shifted vertex is interior; writes its index
  bool shifted_vertex_is_internal(long ell, const int *shift, long &out) const {
    int idx[3];
    vertex_l2e(ell, idx);
    bool ok = true;
    for (int d = 0; d < dim; ++d) {
      idx[d] += shift[d];
      ok = ok && (idx[d] > 0) && (idx[d] < n[d]);
    }
    if (ok) out = vertex_e2l(idx);
    return ok;
  }
  // lattice2d.hh:128-138: pure index arithmetic (the reference only assert()s the bounds, and its
  // own test test_lattice.hh:222-229 shifts onto the boundary in a release build)
  long shift_vertexidx(long ell, const int *shift) const {
    int idx[3];
    vertex_l2e(ell, idx);
    long out = 0, stride = 1;
    for (int d = 0; d < dim; ++d) {
      out += stride * (idx[d] + shift[d] - 1);
      stride *= (n[d] - 1);
    }
    return out;
  }
  // lattice2d.hh:166-175.  NOTE: Lattice1d::corner_is_internal_vertex (lattice1d.hh:132-139)
  // has an off-by-one (treats vertex index 0 as non-internal); we use the 2d/3d-consistent rule
  // in every dimension (affects only 1d FEM / radius>0, which no driver or test uses).
  bool corner_is_internal_vertex(long cell, const int *corner, long &out) const {
    int idx[3];
    cell_l2e(cell, idx);
    bool ok = true;
    for (int d = 0; d < dim; ++d) {
      idx[d] += corner[d];
      ok = ok && (idx[d] > 0) && (idx[d] < n[d]);
    }
    if (ok) out = vertex_e2l(idx);
    return ok;
  }
  // lattice2d.hh:178-187: index of coarse vertex ell on the next-FINER lattice
  long fine_vertex_idx(long ell) const {
    int idx[3];
    vertex_l2e(ell, idx);
    long out = 0, stride = 1;
    for (int d = 0; d < dim; ++d) {
      out += stride * (2 * idx[d] - 1);
      stride *= (2 * n[d] - 1);
    }
    return out;
  }
  void vertex_coordinates(long ell, double *x) const {  // lattice2d.hh:188-195
    int idx[3];
    vertex_l2e(ell, idx);
    for (int d = 0; d < dim; ++d) x[d] = idx[d] * (1.0 / double(n[d]));
  }
  // lattice2d.hh:198-213 (errors are exceptions here; the reference prints + exit(-1))
  Lattice coarse() const {
    int c[3] = {2, 2, 2};
    for (int d = 0; d < dim; ++d) {
      if (n[d] % 2 != 0) throw std::runtime_error("cannot coarsen lattice [one of the extents is odd]");
      if (!(n[d] / 2 > 1)) throw std::runtime_error("cannot coarsen lattice [resulting lattice would have no interior vertices]");
      c[d] = n[d] / 2;
    }
    return Lattice(dim, c);
  }
};

// ---------------------------------------------------------------------------------------------
// CSR matrix (stands in for Eigen::SparseMatrix<double>; all operator matrices are symmetric so
// Eigen's column-major storage read row-wise in sor_smoother.cc:58-61 is the same thing)
// ---------------------------------------------------------------------------------------------
struct Triplet {
  long r, c;
  double v;
};
struct CSR {
  long rows = 0, cols = 0;
  std::vector<long> rowptr;
  std::vector<int> col;
  std::vector<double> val;
  long nnz() const { return (long)val.size(); }
  // Eigen setFromTriplets semantics: duplicates are summed, inner indices sorted
  static CSR from_triplets(long rows, long cols, std::vector<Triplet> &t) {
    CSR A;
    A.rows = rows;
    A.cols = cols;
    std::sort(t.begin(), t.end(), [](const Triplet &a, const Triplet &b) { return a.r != b.r ? a.r < b.r : a.c < b.c; });
    A.rowptr.assign(rows + 1, 0);
    for (size_t k = 0; k < t.size(); ++k) {
      if (k > 0 && t[k].r == t[k - 1].r && t[k].c == t[k - 1].c) {
        A.val.back() += t[k].v;
      } else {
        A.col.push_back((int)t[k].c);
        A.val.push_back(t[k].v);
        A.rowptr[t[k].r + 1]++;
      }
    }
    for (long r = 0; r < rows; ++r) A.rowptr[r + 1] += A.rowptr[r];
    return A;
  }
  void matvec(const double *x, double *y) const {
    for (long r = 0; r < rows; ++r) {
      double s = 0.0;
      for (long k = rowptr[r]; k < rowptr[r + 1]; ++k) s += val[k] * x[col[k]];
      y[r] = s;
    }
  }
  double coeff(long r, long c) const {
    for (long k = rowptr[r]; k < rowptr[r + 1]; ++k)
      if (col[k] == c) return val[k];
    return 0.0;
  }
  double &coeff_ref(long r, long c) {
    for (long k = rowptr[r]; k < rowptr[r + 1]; ++k)
      if (col[k] == c) return val[k];
    throw std::runtime_error("coeff_ref: entry not in sparsity pattern");
  }
  Vec diagonal() const {
    Vec d(rows, 0.0);
    for (long r = 0; r < rows; ++r) d[r] = coeff(r, r);
    return d;
  }
  CSR transpose() const {
    CSR T;
    T.rows = cols;
    T.cols = rows;
    T.rowptr.assign(cols + 1, 0);
    for (long k = 0; k < nnz(); ++k) T.rowptr[col[k] + 1]++;
    for (long r = 0; r < cols; ++r) T.rowptr[r + 1] += T.rowptr[r];
    T.col.resize(nnz());
    T.val.resize(nnz());
    std::vector<long> fill(T.rowptr.begin(), T.rowptr.end() - 1);
    for (long r = 0; r < rows; ++r)
      for (long k = rowptr[r]; k < rowptr[r + 1]; ++k) {
        long p = fill[col[k]]++;
        T.col[p] = (int)r;
        T.val[p] = val[k];
      }
    return T;
  }
  // sparse * sparse (Gustavson), result rows sorted by column (as Eigen's product + pruning-free)
  CSR multiply(const CSR &B) const {
    assert(cols == B.rows);
    CSR C;
    C.rows = rows;
    C.cols = B.cols;
    C.rowptr.assign(rows + 1, 0);
    std::vector<double> acc(B.cols, 0.0);
    std::vector<long> mark(B.cols, -1);
    std::vector<int> touched;
    for (long r = 0; r < rows; ++r) {
      touched.clear();
      for (long k = rowptr[r]; k < rowptr[r + 1]; ++k) {
        const double a = val[k];
        const long j = col[k];
        for (long q = B.rowptr[j]; q < B.rowptr[j + 1]; ++q) {
          const int c = B.col[q];
          if (mark[c] != r) {
            mark[c] = r;
            acc[c] = 0.0;
            touched.push_back(c);
          }
          acc[c] += a * B.val[q];
        }
      }
      std::sort(touched.begin(), touched.end());
      for (int c : touched) {
        C.col.push_back(c);
        C.val.push_back(acc[c]);
      }
      C.rowptr[r + 1] = (long)C.col.size();
    }
    return C;
  }
  std::vector<double> to_dense() const {  // row-major
    std::vector<double> D(rows * cols, 0.0);
    for (long r = 0; r < rows; ++r)
      for (long k = rowptr[r]; k < rowptr[r + 1]; ++k) D[r * cols + col[k]] += val[k];
    return D;
  }
};

// ---------------------------------------------------------------------------------------------
// small dense helpers (stand in for Eigen LLT / .inverse())
// ---------------------------------------------------------------------------------------------
// in-place lower Cholesky of row-major n x n SPD matrix (Eigen::LLT<.,Lower>, cholesky_wrapper.cc:131-135)
inline void dense_cholesky(std::vector<double> &A, long n) {
  for (long j = 0; j < n; ++j) {
    double d = A[j * n + j];
    for (long k = 0; k < j; ++k) d -= A[j * n + k] * A[j * n + k];
    if (!(d > 0.0)) throw std::runtime_error("dense_cholesky: matrix not positive definite");
    d = std::sqrt(d);
    A[j * n + j] = d;
    for (long i = j + 1; i < n; ++i) {
      double s = A[i * n + j];
      const double *ai = &A[i * n], *aj = &A[j * n];
      for (long k = 0; k < j; ++k) s -= ai[k] * aj[k];
      A[i * n + j] = s / d;
    }
    for (long k = j + 1; k < n; ++k) A[j * n + k] = 0.0;
  }
}
inline void dense_solveL(const std::vector<double> &L, long n, const double *b, double *x) {  // L x = b
  for (long i = 0; i < n; ++i) {
    double s = b[i];
    const double *li = &L[i * n];
    for (long k = 0; k < i; ++k) s -= li[k] * x[k];
    x[i] = s / li[i];
  }
}
inline void dense_solveLT(const std::vector<double> &L, long n, const double *b, double *x) {  // L^T x = b
  for (long i = n - 1; i >= 0; --i) {
    double s = b[i];
    for (long k = i + 1; k < n; ++k) s -= L[k * n + i] * x[k];
    x[i] = s / L[i * n + i];
  }
}
// general inverse by Gauss-Jordan with partial pivoting (row-major, n small)
inline std::vector<double> dense_inverse(std::vector<double> A, long n) {
  std::vector<double> I(n * n, 0.0);
  for (long i = 0; i < n; ++i) I[i * n + i] = 1.0;
  for (long c = 0; c < n; ++c) {
    long p = c;
    for (long r = c + 1; r < n; ++r)
      if (std::fabs(A[r * n + c]) > std::fabs(A[p * n + c])) p = r;
    if (A[p * n + c] == 0.0) throw std::runtime_error("dense_inverse: singular matrix");
    if (p != c)
      for (long k = 0; k < n; ++k) {
        std::swap(A[p * n + k], A[c * n + k]);
        std::swap(I[p * n + k], I[c * n + k]);
      }
    const double inv = 1.0 / A[c * n + c];
    for (long k = 0; k < n; ++k) {
      A[c * n + k] *= inv;
      I[c * n + k] *= inv;
    }
    for (long r = 0; r < n; ++r) {
      if (r == c) continue;
      const double f = A[r * n + c];
      if (f == 0.0) continue;
      for (long k = 0; k < n; ++k) {
        A[r * n + k] -= f * A[c * n + k];
        I[r * n + k] -= f * I[c * n + k];
      }
    }
  }
  return I;
}

// ---------------------------------------------------------------------------------------------
// cartesian_product (auxilliary/common.hh:27-52): last index fastest... NO: the recursion appends
// the new factor LAST and iterates it fastest, i.e. the FIRST component is the slowest.
// ---------------------------------------------------------------------------------------------
template <class T>
std::vector<std::vector<T>> cartesian_product(const std::vector<T> &v, int n) {
  std::vector<std::vector<T>> prod;
  if (n == 1) {
    for (const T &x : v) prod.push_back(std::vector<T>{x});
  } else {
    std::vector<std::vector<T>> prev = cartesian_product(v, n - 1);
    for (auto &s : prev)
      for (const T &x : v) {
        std::vector<T> sj = s;
        sj.push_back(x);
        prod.push_back(sj);
      }
  }
  return prod;
}

// Gauss-Legendre quadrature on [0,1]^d (auxilliary/quadrature.cc:13-56)
struct GaussLegendreQuadrature {
  std::vector<double> weights;
  std::vector<std::vector<double>> points;
  GaussLegendreQuadrature(int dim, int order) {
    std::vector<double> w1, p1;
    switch (order) {
      case 0: w1 = {2.0}; p1 = {0.0}; break;
      case 1: w1 = {1.0, 1.0}; p1 = {-1.0 / std::sqrt(3.0), +1.0 / std::sqrt(3.0)}; break;
      case 2: w1 = {5.0 / 9.0, 8.0 / 9.0, 5.0 / 9.0}; p1 = {-std::sqrt(3.0 / 5.0), 0.0, +std::sqrt(3.0 / 5.0)}; break;
      default: throw std::runtime_error("quadrature order must be 0,1,2");
    }
    for (auto &w : cartesian_product(w1, dim)) {
      double x = 1.0;
      for (double s : w) x *= 0.5 * s;
      weights.push_back(x);
    }
    for (auto &p : cartesian_product(p1, dim)) {
      std::vector<double> q(dim);
      for (int j = 0; j < dim; ++j) q[j] = 0.5 * (p[j] + 1.0);
      points.push_back(q);
    }
  }
};

// ---------------------------------------------------------------------------------------------
// Correlation length model (linear_operator/correlationlength_model.hh:45-113)
// ---------------------------------------------------------------------------------------------
struct KappaModel {
  int kind = 0;  // 0 constant, 1 periodic
  double Lambda = 1.0, Lambda_min = 0.0, Lambda_max = 0.0;
  double kappa_sq(const double *x, int dim) const {
    if (kind == 0) return 1. / std::pow(Lambda, 2);
    const double L1 = 0.5 * (Lambda_max + Lambda_min), L2 = 0.5 * (Lambda_max - Lambda_min);
    double L = L2;
    for (int d = 0; d < dim; ++d) L *= std::cos(M_PI * x[d]);
    L += L1;
    return 1. / (L * L);
  }
};

// ---------------------------------------------------------------------------------------------
// Intergrid operator (intergrid/intergrid_operator.hh:40-144, intergrid_operator.cc:8-19,
// intergrid_operator_linear.cc:8-30): un-normalised full weighting {1/2,1,1/2}^(x)d
// ---------------------------------------------------------------------------------------------
struct Intergrid {
  Lattice fine, coarse;
  int stencil_size;
  std::vector<double> matrix;
  std::vector<long> colidx;
  Intergrid() : stencil_size(0) {}
  explicit Intergrid(const Lattice &lat) : fine(lat), coarse(lat.coarse()) {
    const int dim = lat.dim;
    stencil_size = 1;
    for (int d = 0; d < dim; ++d) stencil_size *= 3;
    const double stencil1d[3] = {0.5, 1.0, 0.5};
    const int shift1d[3] = {-1, 0, +1};
    matrix.resize(stencil_size);
    std::vector<std::vector<int>> shift;
    for (int j = 0; j < stencil_size; ++j) {  // intergrid_operator_linear.cc:16-29
      matrix[j] = 1.0;
      std::vector<int> s(3, 0);
      int mu = j;
      for (int d = 0; d < dim; ++d) {
        matrix[j] *= stencil1d[mu % 3];
        s[d] = shift1d[mu % 3];
        mu /= 3;
      }
      shift.push_back(s);
    }
    const long Nc = coarse.Nvertex();
    colidx.resize(Nc * stencil_size);
    for (long ec = 0; ec < Nc; ++ec) {  // intergrid_operator.cc:8-19
      int idx[3];
      coarse.vertex_l2e(ec, idx);
      for (int d = 0; d < dim; ++d) idx[d] *= 2;
      const long ell = fine.vertex_e2l(idx);
      for (int j = 0; j < stencil_size; ++j) colidx[ec * stencil_size + j] = fine.shift_vertexidx(ell, shift[j].data());
    }
  }
  void restrict(const double *x, double *xc) const {  // intergrid_operator.hh:74-88
    const long Nc = coarse.Nvertex();
    for (long ec = 0; ec < Nc; ++ec) {
      double result = 0;
      for (int k = 0; k < stencil_size; ++k) result += matrix[k] * x[colidx[ec * stencil_size + k]];
      xc[ec] = result;
    }
  }
  void prolongate_add(double alpha, const double *xc, double *x) const {  // intergrid_operator.hh:106-120
    const long Nc = coarse.Nvertex();
    for (long ec = 0; ec < Nc; ++ec) {
      const double v = xc[ec];
      for (int k = 0; k < stencil_size; ++k) x[colidx[ec * stencil_size + k]] += alpha * matrix[k] * v;
    }
  }
  CSR to_sparse() const {  // intergrid_operator.hh:123-144
    std::vector<Triplet> t;
    const long Nc = coarse.Nvertex();
    t.reserve(Nc * stencil_size);
    for (long ec = 0; ec < Nc; ++ec)
      for (int k = 0; k < stencil_size; ++k) t.push_back({ec, colidx[ec * stencil_size + k], matrix[k]});
    return CSR::from_triplets(Nc, fine.Nvertex(), t);
  }
};

// ---------------------------------------------------------------------------------------------
// LinearOperator A = A_0 + B Sigma^{-1} B^T   (linear_operator/linear_operator.hh:28-198)
// ---------------------------------------------------------------------------------------------
struct LinearOperator {
  Lattice lattice;
  int m_lowrank = 0;
  CSR A;              // A_sparse
  CSR B;              // n x m
  CSR BT;             // m x n (B^T; rows = measurement columns)
  Vec Sigma;          // diagonal of Sigma
  long ndof() const { return A.rows; }
  void set_B(const CSR &B_) {
    B = B_;
    BT = B.transpose();
  }
  // linear_operator.hh:66-76
  void apply(const double *x, double *y) const {
    A.matvec(x, y);
    if (m_lowrank > 0) {
      Vec t(m_lowrank);
      BT.matvec(x, t.data());
      for (int k = 0; k < m_lowrank; ++k) t[k] /= Sigma[k];
      for (long r = 0; r < B.rows; ++r)
        for (long q = B.rowptr[r]; q < B.rowptr[r + 1]; ++q) y[r] += B.val[q] * t[B.col[q]];
    }
  }
  // linear_operator.cc:10-23
  LinearOperator coarsen(const Intergrid &ig) const {
    const CSR R = ig.to_sparse();
    const CSR P = R.transpose();
    LinearOperator c;
    c.lattice = lattice.coarse();
    c.m_lowrank = m_lowrank;
    c.A = R.multiply(A).multiply(P);
    if (m_lowrank > 0) c.set_B(R.multiply(B));
    c.Sigma = Sigma;
    return c;
  }
  // linear_operator.cc:26-33 (row-major dense)
  std::vector<double> precision() const {
    const long n = ndof();
    std::vector<double> Q = A.to_dense();
    for (int k = 0; k < m_lowrank; ++k)
      for (long p = BT.rowptr[k]; p < BT.rowptr[k + 1]; ++p)
        for (long q = BT.rowptr[k]; q < BT.rowptr[k + 1]; ++q) Q[(long)BT.col[p] * n + BT.col[q]] += BT.val[p] * BT.val[q] / Sigma[k];
    return Q;
  }
  std::vector<double> covariance() const {  // linear_operator.hh:180-183
    const long n = ndof();
    std::vector<double> L = precision();
    dense_cholesky(L, n);
    std::vector<double> C(n * n);
    Vec e(n), y(n), x(n);
    for (long j = 0; j < n; ++j) {
      std::fill(e.begin(), e.end(), 0.0);
      e[j] = 1.0;
      dense_solveL(L, n, e.data(), y.data());
      dense_solveLT(L, n, y.data(), x.data());
      for (long i = 0; i < n; ++i) C[i * n + j] = x[i];
    }
    return C;
  }
  // A_0^{-1} applied to the columns of B and to extra vectors, via dense Cholesky of A_0 (the
  // reference uses Eigen::SimplicialLLT, linear_operator.hh:122-127; same mathematics)
  struct PriorSolve {
    std::vector<double> L;
    long n;
    void solve(const double *b, double *x) const {
      Vec y(n);
      dense_solveL(L, n, b, y.data());
      dense_solveLT(L, n, y.data(), x);
    }
  };
  PriorSolve prior_factor() const {
    if (ndof() > 20000) throw std::runtime_error("oracle: dense prior factorisation limited to ndof <= 20000");
    PriorSolve s;
    s.n = ndof();
    s.L = A.to_dense();
    dense_cholesky(s.L, s.n);
    return s;
  }
  // linear_operator.hh:119-139:  x|y = xbar + A0^{-1} B (Sigma + B^T A0^{-1} B)^{-1} (y - B^T xbar)
  Vec mean(const double *xbar, const double *y) const {
    const long n = ndof();
    Vec out(xbar, xbar + n);
    if (m_lowrank == 0) return out;
    const int m = m_lowrank;
    PriorSolve ps = prior_factor();
    std::vector<Vec> Bbar(m, Vec(n));
    Vec col(n);
    for (int k = 0; k < m; ++k) {
      std::fill(col.begin(), col.end(), 0.0);
      for (long p = BT.rowptr[k]; p < BT.rowptr[k + 1]; ++p) col[BT.col[p]] = BT.val[p];
      ps.solve(col.data(), Bbar[k].data());
    }
    std::vector<double> S(m * m, 0.0);
    for (int a = 0; a < m; ++a)
      for (int b = 0; b < m; ++b) {
        double s = (a == b) ? Sigma[a] : 0.0;
        for (long p = BT.rowptr[a]; p < BT.rowptr[a + 1]; ++p) s += BT.val[p] * Bbar[b][BT.col[p]];
        S[a * m + b] = s;
      }
    std::vector<double> Sinv = dense_inverse(S, m);
    Vec rhs(m), z(m, 0.0);
    BT.matvec(xbar, rhs.data());
    for (int k = 0; k < m; ++k) rhs[k] = y[k] - rhs[k];
    for (int a = 0; a < m; ++a)
      for (int b = 0; b < m; ++b) z[a] += Sinv[a * m + b] * rhs[b];
    for (int k = 0; k < m; ++k)
      for (long i = 0; i < n; ++i) out[i] += Bbar[k][i] * z[k];
    return out;
  }
  // linear_operator.hh:153-174
  void observed_mean_and_variance(const double *xbar, const double *y, const double *b_obs, double &mean_, double &variance) const {
    const long n = ndof();
    const int m = m_lowrank;
    PriorSolve ps = prior_factor();
    Vec bbar(n);
    ps.solve(b_obs, bbar.data());
    mean_ = 0.0;
    variance = 0.0;
    for (long i = 0; i < n; ++i) {
      mean_ += b_obs[i] * xbar[i];
      variance += b_obs[i] * bbar[i];
    }
    if (m > 0) {
      std::vector<Vec> Bbar(m, Vec(n));
      Vec col(n);
      for (int k = 0; k < m; ++k) {
        std::fill(col.begin(), col.end(), 0.0);
        for (long p = BT.rowptr[k]; p < BT.rowptr[k + 1]; ++p) col[BT.col[p]] = BT.val[p];
        ps.solve(col.data(), Bbar[k].data());
      }
      std::vector<double> S(m * m, 0.0);
      for (int a = 0; a < m; ++a)
        for (int b = 0; b < m; ++b) {
          double s = (a == b) ? Sigma[a] : 0.0;
          for (long p = BT.rowptr[a]; p < BT.rowptr[a + 1]; ++p) s += BT.val[p] * Bbar[b][BT.col[p]];
          S[a * m + b] = s;
        }
      std::vector<double> Sinv = dense_inverse(S, m);
      Vec rhs(m), BTbbar(m);
      BT.matvec(xbar, rhs.data());
      for (int k = 0; k < m; ++k) rhs[k] = y[k] - rhs[k];
      BT.matvec(bbar.data(), BTbbar.data());
      for (int a = 0; a < m; ++a)
        for (int b = 0; b < m; ++b) {
          mean_ += BTbbar[a] * Sinv[a * m + b] * rhs[b];
          variance -= BTbbar[a] * Sinv[a * m + b] * BTbbar[b];
        }
    }
  }
};

// shiftedlaplace_fd_operator.cc:9-56
LinearOperator make_shiftedlaplace_fd(const Lattice &lattice, const KappaModel &km);
// squared_shiftedlaplace_fd_operator.cc:9-96
LinearOperator make_squared_shiftedlaplace_fd(const Lattice &lattice, const KappaModel &km);
// shiftedlaplace_fem_operator.cc:9-187
LinearOperator make_shiftedlaplace_fem(const Lattice &lattice, const KappaModel &km);
// sampler/test_sampler.hh:30-67 (TestOperator1d)
LinearOperator make_test_operator_1d(bool lowrank);

// measured_operator.cc:69-170: returns sparse vector as (index,value) lists
void measurement_vector(const Lattice &lattice, const double *x0, double radius, std::vector<long> &idx, std::vector<double> &val);
// measured_operator.cc:9-49
LinearOperator make_measured_operator(const LinearOperator &base, int n_meas, const double *locations, const double *variance_scaled,
                                      double radius, bool measure_global, double variance_global);

// ---------------------------------------------------------------------------------------------
// Site orderings for the sweeps.  ORDER_LEX is the reference (sor_smoother.cc:66-69).
// ORDER_COLOUR is the multicolour ordering of the B200 path: the reference sweep run on
// P A P^T.  `order[k]` = k-th row visited by a forward sweep.
// ---------------------------------------------------------------------------------------------
enum { ORDER_LEX = 0, ORDER_COLOUR = 1 };
// number of colours the B200 path uses for a 2d stencil matrix: 2 (5-pt), 4 (9-pt), 9 (radius 2)
int colour_count_2d(const Lattice &lat, const CSR &A);
std::vector<long> make_order(const Lattice &lat, const CSR &A, int ordering);

// ---------------------------------------------------------------------------------------------
// Noise sources.  NormalSource mirrors Sampler's (rng&, per-object normal_distribution)
// (sampler/sampler.hh:31-34,69-71).  PhiloxNoise is the counter-based stream of the B200 path.
// ---------------------------------------------------------------------------------------------
struct Philox {
  // Philox4x32-R (Salmon et al. 2011, Random123), key = 64-bit seed, counter = 4 x 32 bit
  static inline void round_(uint32_t &c0, uint32_t &c1, uint32_t &c2, uint32_t &c3, uint32_t k0, uint32_t k1) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
  }
  // rounds: the CUDA path runs Philox4x32-7 (multigridmc_b200/csrc/philox.cuh kPhiloxRounds; same switch)
#ifndef MGMC_PHILOX_ROUNDS
#define MGMC_PHILOX_ROUNDS 7
#endif
  static inline void philox4x32(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int r = 0; r < MGMC_PHILOX_ROUNDS; ++r) {
      round_(c[0], c[1], c[2], c[3], k0, k1);
      k0 += 0x9E3779B9u;
      k1 += 0xBB67AE85u;
    }
  }
  // two N(0,1) from one counter: Box-Muller on two 52-bit uniforms (k + 1/2) 2^-52 in (0,1)
  static inline void normal_pair(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, double &z0, double &z1) {
    uint32_t c[4] = {c0, c1, c2, c3};
    philox4x32(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    const uint64_t a = (uint64_t)c[0] | ((uint64_t)c[1] << 32), b = (uint64_t)c[2] | ((uint64_t)c[3] << 32);
    const double u1 = ((double)(a >> 12) + 0.5) * (1.0 / 4503599627370496.0);
    const double u2 = ((double)(b >> 12) + 0.5) * (1.0 / 4503599627370496.0);
    const double r = std::sqrt(-2.0 * std::log(u1));
    const double t = 2.0 * M_PI * u2;
    z0 = r * std::cos(t);
    z1 = r * std::sin(t);
  }
};

// Counter layout shared with the CUDA path (multigridmc_b200/csrc/philox.cuh):
//   c0 = stream-local index (see below), c1 = (level << 24) | sweep_counter, c2 = sample index,
//   c3 = chain id;  key = seed.
//   site noise (2d):   c0 = ((j * G + (i >> 2)) << 1) | (i & 1),  G = nx/4 + 1, normal = (i & 2) ? z1 : z0
//                      (i,j = Euclidean vertex index, 1..n-1): the pair (i, i+2) inside an aligned
//                      group of 4 columns shares one Philox call.
//   low-rank noise:    c0 = 0x80000000 | (k >> 1), normal = (k & 1) ? z1 : z0        (k-th measurement)
//   coarse Cholesky:   c0 = 0x40000000 | (ell >> 1), normal = (ell & 1) ? z1 : z0
struct PhiloxCtx {
  uint64_t seed = 0;
  uint32_t sample = 0, chain = 0;
  std::vector<uint32_t> sweep_counter;  // per level, reset at the start of every sample
};

struct NoiseSource {
  // reference mode
  std::mt19937_64 *engine = nullptr;
  std::normal_distribution<double> dist{0.0, 1.0};
  // philox mode
  PhiloxCtx *px = nullptr;
  bool philox() const { return px != nullptr; }
};

// ---------------------------------------------------------------------------------------------
// Smoothers (smoother/sor_smoother.{hh,cc}, smoother/ssor_smoother.{hh,cc})
// ---------------------------------------------------------------------------------------------
enum Direction { forward = 1, backward = 2 };

struct SORSmoother {
  const LinearOperator *op;
  double omega;
  int nsmooth;
  Direction direction;
  std::vector<long> order;  // forward visiting order (identity = reference)
  Vec diag;
  std::vector<Vec> B_bar;   // m columns of length n (dense, as sor_smoother.cc:28-35)
  SORSmoother(const LinearOperator *op_, double omega_, int nsmooth_, Direction dir_, const std::vector<long> &order_);
  void apply(const double *b, double *x) const;         // sor_smoother.cc:41-53
  void apply_sparse(const double *b, double *x) const;  // sor_smoother.cc:56-78
  void sweep_once(const double *b, double *x) const;
};

struct SSORSmoother {  // ssor_smoother.hh:44-49, ssor_smoother.cc:9-16
  int nsmooth;
  SORSmoother fwd, bwd;
  SSORSmoother(const LinearOperator *op, double omega, int nsmooth_, const std::vector<long> &order)
      : nsmooth(nsmooth_), fwd(op, omega, 1, forward, order), bwd(op, omega, 1, backward, order) {}
  void apply(const double *b, double *x) const {
    for (int k = 0; k < nsmooth; ++k) {
      fwd.apply(b, x);
      bwd.apply(b, x);
    }
  }
};

// ---------------------------------------------------------------------------------------------
// Samplers (sampler/sor_sampler.{hh,cc}, ssor_sampler.{hh,cc}, cholesky_sampler.{hh,cc})
// ---------------------------------------------------------------------------------------------
struct SORSampler {
  const LinearOperator *op;
  mutable NoiseSource noise;
  double omega;
  Direction direction;
  int nsmooth;
  int level;  // only used for the philox counter
  Vec sqrt_precision_diag;
  Vec Sigma_inv_sqrt;
  SORSmoother smoother;
  mutable Vec c_rhs, xi;
  SORSampler(const LinearOperator *op_, NoiseSource noise_, double omega_, int nsmooth_, Direction dir_, const std::vector<long> &order, int level_);
  void apply(const double *f, double *x) const;  // sor_sampler.cc:37-58
};

struct SSORSampler {  // ssor_sampler.hh:30-36, ssor_sampler.cc:9-16
  int nsmooth;
  SORSampler fwd, bwd;
  SSORSampler(const LinearOperator *op, NoiseSource noise, double omega, int nsmooth_, const std::vector<long> &order, int level)
      : nsmooth(nsmooth_), fwd(op, noise, omega, 1, forward, order, level), bwd(op, noise, omega, 1, backward, order, level) {}
  void apply(const double *f, double *x) const {
    for (int k = 0; k < nsmooth; ++k) {
      fwd.apply(f, x);
      bwd.apply(f, x);
    }
  }
};

// Dense Cholesky sampler (cholesky_sampler.hh:30-101, cholesky_sampler.cc:25-38).  The reference's
// "sparse" variants (CHOLMOD / SimplicialLLT with AMD permutation) draw from the same
// distribution with a permuted factor; the oracle always uses the un-permuted dense factor.
struct CholeskySampler {
  const LinearOperator *op;
  mutable NoiseSource noise;
  int level;
  long n;
  std::vector<double> L;
  mutable Vec xi, g;
  bool rhs_fixed = false;
  Vec g_rhs;
  CholeskySampler(const LinearOperator *op_, NoiseSource noise_, int level_);
  void apply(const double *f, double *x) const;  // cholesky_sampler.hh:50-66
  void fix_rhs(const double *f);                 // cholesky_sampler.hh:75-80
  void unfix_rhs() { rhs_fixed = false; }
};

// cholesky_solver.cc:8-41
struct CholeskySolver {
  const LinearOperator *op;
  long n;
  std::vector<double> L;
  std::vector<Vec> B_bar;
  explicit CholeskySolver(const LinearOperator *op_);
  void apply(const double *b, double *x) const;
};

// ---------------------------------------------------------------------------------------------
// Multigrid parameters (auxilliary/parameters.hh:145-174)
// ---------------------------------------------------------------------------------------------
struct MultigridParameters {
  int nlevel = 2;
  int smoother = 1;       // 0 "SOR", 1 "SSOR"
  int coarse_solver = 1;  // 0 "SSOR", 1 "Cholesky"
  int npresmooth = 1, npostsmooth = 1, ncoarsesmooth = 1;
  int cycle = 1;
  double coarse_scaling = 1.0;
  double omega = 1.0;
  int ordering = ORDER_LEX;
};

struct Hierarchy {
  std::vector<std::shared_ptr<LinearOperator>> ops;
  std::vector<Intergrid> intergrids;
  std::vector<std::vector<long>> orders;
  Hierarchy(const std::shared_ptr<LinearOperator> &fine, int nlevel, int ordering);
};

// sampler/multigridmc_sampler.{hh,cc}
struct MultigridMCSampler {
  std::shared_ptr<Hierarchy> H;
  MultigridParameters params;
  std::mt19937_64 *engine;
  std::shared_ptr<PhiloxCtx> px;  // non-null => philox noise
  struct LevelSamplers {
    std::shared_ptr<SORSampler> sor_pre, sor_post;
    std::shared_ptr<SSORSampler> ssor_pre, ssor_post;
  };
  std::vector<LevelSamplers> samplers;
  std::shared_ptr<CholeskySampler> coarse_cholesky;
  std::shared_ptr<SSORSampler> coarse_ssor;
  mutable std::vector<Vec> x_ell, f_ell, r_ell;
  MultigridMCSampler(const std::shared_ptr<Hierarchy> &H_, std::mt19937_64 *engine_, const MultigridParameters &p, bool use_philox, uint64_t philox_seed);
  void apply(const double *f, double *x) const;  // multigridmc_sampler.cc:133-138
  void sample(int level) const;                  // multigridmc_sampler.cc:103-130
};

// preconditioner/multigrid_preconditioner.{hh,cc}
struct MultigridPreconditioner {
  std::shared_ptr<Hierarchy> H;
  MultigridParameters params;
  struct LevelSmoothers {
    std::shared_ptr<SORSmoother> sor_pre, sor_post;
    std::shared_ptr<SSORSmoother> ssor_pre, ssor_post;
  };
  std::vector<LevelSmoothers> smoothers;
  std::shared_ptr<CholeskySolver> coarse_solver;
  std::vector<Vec> x_ell, b_ell, r_ell;
  MultigridPreconditioner(const std::shared_ptr<Hierarchy> &H_, const MultigridParameters &p);
  void apply(const double *b, double *x);  // multigrid_preconditioner.cc:104-108
  void solve(int level);                   // multigrid_preconditioner.cc:74-101
};

// solver/loop_solver.cc:9-53.  Returns number of iterations (k at convergence, or maxiter) and
// fills history with ||r_k|| for k = 0.. (one entry per evaluated residual).
struct LoopSolverResult {
  bool converged = false;
  int niter = 0;
  double r0_nrm = 0.0;
  std::vector<double> history;
};
LoopSolverResult loop_solve(const LinearOperator &op, MultigridPreconditioner &prec, double rtol, double atol, int maxiter, int verbose,
                            const double *b, double *x);

// auxilliary/statistics.cc:65-79 for a scalar time series (window = number of lags k_max)
double tau_int_scalar(const double *series, long n, int k_max);

}  // namespace orc
#endif
