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
#include <utility>
#include <vector>

namespace Eigen {

template <typename T>
class VectorX {
 public:
  VectorX() {}
  explicit VectorX(std::ptrdiff_t n) : v_(n) {}
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
