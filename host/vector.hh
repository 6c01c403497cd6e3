// Minimal stand-ins for the Eigen types that appear in the reference's public interfaces
// (Eigen::VectorXd / VectorXi / SparseVector<double>).  Eigen is not installed in the build image; when
// it is, define MGMC_USE_EIGEN and the real headers are used instead -- every host class below only
// relies on size(), data(), operator[] / (), setZero(), norm(), dot() and element-wise arithmetic.
#ifndef MGMC_HOST_VECTOR_HH
#define MGMC_HOST_VECTOR_HH
#if defined(MGMC_USE_EIGEN)
#include <Eigen/Dense>
#include <Eigen/Sparse>
#else
#include <cmath>
#include <cstddef>
#include <initializer_list>
#include <map>
#include <utility>
#include <vector>

namespace Eigen {

template <typename T>
class VectorX {
 public:
  VectorX() {}
  explicit VectorX(std::ptrdiff_t n) : v_(n) {}
  VectorX(std::ptrdiff_t rows, std::ptrdiff_t /*cols: Eigen's (rows, cols) form, used by test_solver.hh:45*/) : v_(rows) {}
  VectorX(std::initializer_list<T> l) : v_(l) {}
  std::ptrdiff_t size() const { return (std::ptrdiff_t)v_.size(); }
  void resize(std::ptrdiff_t n) { v_.resize(n); }
  T *data() { return v_.data(); }
  const T *data() const { return v_.data(); }
  T &operator[](std::ptrdiff_t i) { return v_[i]; }
  const T &operator[](std::ptrdiff_t i) const { return v_[i]; }
  T &operator()(std::ptrdiff_t i) { return v_[i]; }
  const T &operator()(std::ptrdiff_t i) const { return v_[i]; }
  void setZero() {
    for (auto &e : v_) e = T(0);
  }
  void setConstant(T c) {
    for (auto &e : v_) e = c;
  }
  double dot(const VectorX &o) const {
    double s = 0;
    for (size_t i = 0; i < v_.size(); ++i) s += double(v_[i]) * double(o.v_[i]);
    return s;
  }
  double norm() const { return std::sqrt(dot(*this)); }
  VectorX &operator+=(const VectorX &o) {
    for (size_t i = 0; i < v_.size(); ++i) v_[i] += o.v_[i];
    return *this;
  }
  VectorX &operator-=(const VectorX &o) {
    for (size_t i = 0; i < v_.size(); ++i) v_[i] -= o.v_[i];
    return *this;
  }
  VectorX &operator*=(T a) {
    for (auto &e : v_) e *= a;
    return *this;
  }
  friend VectorX operator+(VectorX a, const VectorX &b) { return a += b; }
  friend VectorX operator-(VectorX a, const VectorX &b) { return a -= b; }
  friend VectorX operator*(T a, VectorX b) { return b *= a; }
  friend VectorX operator*(VectorX b, T a) { return b *= a; }
  friend VectorX operator/(VectorX b, T a) { return b *= (T(1) / a); }
  friend bool operator==(const VectorX &a, const VectorX &b) { return a.v_ == b.v_; }
  friend bool operator!=(const VectorX &a, const VectorX &b) { return !(a == b); }
  VectorX cwiseProduct(const VectorX &o) const {
    VectorX r(*this);
    for (size_t i = 0; i < v_.size(); ++i) r.v_[i] *= o.v_[i];
    return r;
  }

 private:
  std::vector<T> v_;
};
typedef VectorX<double> VectorXd;
typedef VectorX<int> VectorXi;
// (fixed-size names used by the reference's tests: same dynamic type here)
typedef VectorX<int> Vector2i;
typedef VectorX<int> Vector3i;
typedef VectorX<double> Vector2d;
typedef VectorX<double> Vector3d;

// sparse matrix as a (row, col) -> value map: what LinearOperator::get_sparse() returns in the host layer (difference and
// Frobenius norm are all the reference's tests ask of it, test_intergrid.hh:188,206)
template <typename T>
class SparseMatrix {
 public:
  SparseMatrix(std::ptrdiff_t rows = 0, std::ptrdiff_t cols = 0) : rows_(rows), cols_(cols) {}
  std::ptrdiff_t rows() const { return rows_; }
  std::ptrdiff_t cols() const { return cols_; }
  T &coeffRef(std::ptrdiff_t i, std::ptrdiff_t j) { return e_[{i, j}]; }
  T coeff(std::ptrdiff_t i, std::ptrdiff_t j) const {
    auto it = e_.find({i, j});
    return it == e_.end() ? T(0) : it->second;
  }
  std::ptrdiff_t nonZeros() const { return (std::ptrdiff_t)e_.size(); }
  double norm() const {
    double s = 0;
    for (auto &kv : e_) s += double(kv.second) * double(kv.second);
    return std::sqrt(s);
  }
  friend SparseMatrix operator-(SparseMatrix a, const SparseMatrix &b) {
    for (auto &kv : b.e_) a.e_[kv.first] -= kv.second;
    return a;
  }
  const std::map<std::pair<std::ptrdiff_t, std::ptrdiff_t>, T> &entries() const { return e_; }

 private:
  std::ptrdiff_t rows_, cols_;
  std::map<std::pair<std::ptrdiff_t, std::ptrdiff_t>, T> e_;
};

// sparse vector as (index, value) pairs -- what MeasuredOperator::measurement_vector returns
template <typename T>
class SparseVector {
 public:
  explicit SparseVector(std::ptrdiff_t n = 0) : n_(n) {}
  std::ptrdiff_t size() const { return n_; }
  T &coeffRef(std::ptrdiff_t i) {
    for (auto &e : e_)
      if (e.first == i) return e.second;
    e_.push_back({i, T(0)});
    return e_.back().second;
  }
  double dot(const VectorXd &x) const {
    double s = 0;
    for (auto &e : e_) s += e.second * x[e.first];
    return s;
  }
  const std::vector<std::pair<std::ptrdiff_t, T>> &entries() const { return e_; }

 private:
  std::ptrdiff_t n_;
  std::vector<std::pair<std::ptrdiff_t, T>> e_;
};

}  // namespace Eigen
#endif
#endif
