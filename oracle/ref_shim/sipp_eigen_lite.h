// Minimal stand-in for the parts of Eigen 3 that the reference's hot-path translation units use (TEST INFRASTRUCTURE: it exists
// only so that oracle/_ref can compile the reference's own evaluator sources unmodified; Eigen itself is not installed in this
// image). Fixed-size vectors with Eigen's operation order: squaredNorm() follows Eigen's vectorised redux on x86-64
// ((x*x + y*y) + z*z for a 3-vector), norm() = sqrt of it, quaternion * vector uses Eigen's
// v + w * 2(q x v) + q x 2(q x v) form (exact for the identity rotation the shipped cfgs use).
#ifndef SIPP_EIGEN_LITE_H_
#define SIPP_EIGEN_LITE_H_
#include <math.h>

#include <cmath>
#include <cstddef>
#include <iostream>
#include <vector>

namespace Eigen {

template <class T, int N>
struct Vec;

template <class T, int N, int K>
struct HeadRef {   // v.head(K) as an lvalue / rvalue
  T* p;
  HeadRef& operator=(const Vec<T, K>& o) {
    for (int i = 0; i < K; ++i) p[i] = o.v[i];
    return *this;
  }
  operator Vec<T, K>() const {
    Vec<T, K> r;
    for (int i = 0; i < K; ++i) r.v[i] = p[i];
    return r;
  }
};

template <class T, int N>
struct Vec {
  T v[N];
  Vec() {
    for (int i = 0; i < N; ++i) v[i] = T();
  }
  Vec(T a, T b) {
    static_assert(N == 2, "two coefficients");
    v[0] = a; v[1] = b;
  }
  Vec(T a, T b, T c) {
    static_assert(N == 3, "three coefficients");
    v[0] = a; v[1] = b; v[2] = c;
  }
  static Vec Zero() { return Vec(); }
  T& operator[](int i) { return v[i]; }
  const T& operator[](int i) const { return v[i]; }
  T& operator()(int i) { return v[i]; }
  const T& operator()(int i) const { return v[i]; }
  T& x() { return v[0]; }
  T& y() { return v[1]; }
  T& z() { static_assert(N >= 3, "z"); return v[2]; }
  const T& x() const { return v[0]; }
  const T& y() const { return v[1]; }
  const T& z() const { static_assert(N >= 3, "z"); return v[2]; }
  HeadRef<T, N, 2> head(int k) { (void)k; return HeadRef<T, N, 2>{v}; }
  Vec<T, 2> head(int k) const {
    (void)k;
    Vec<T, 2> r;
    r.v[0] = v[0]; r.v[1] = v[1];
    return r;
  }
  template <int K>
  HeadRef<T, N, K> head() { return HeadRef<T, N, K>{v}; }
  template <class U>
  Vec<U, N> cast() const {
    Vec<U, N> r;
    for (int i = 0; i < N; ++i) r.v[i] = static_cast<U>(v[i]);
    return r;
  }
  // Eigen 3.3 (the reference's Docker image: Ubuntu 18.04, x86-64, SSE2 => EIGEN_VECTORIZE, EIGEN_UNALIGNED_VECTORIZE) reduces a
  // fixed-size double vector with LinearVectorizedTraversal: Packet2d partial sums, predux, then the leftover scalar — for a
  // 3-vector (x*x + y*y) + z*z, for a 2-vector x*x + y*y (Redux.h, redux_impl<..., LinearVectorizedTraversal, CompleteUnrolling>)
  T squaredNorm() const {
    static_assert(N <= 3, "only the 2- and 3-vectors of the reference are covered");
    T s = v[0] * v[0];
    for (int i = 1; i < N; ++i) s = s + v[i] * v[i];
    return s;
  }
  T norm() const { return std::sqrt(squaredNorm()); }
  Vec normalized() const {
    const T n = norm();
    return n > T(0) ? (*this) / n : *this;
  }
  Vec cross(const Vec& o) const {
    static_assert(N == 3, "cross");
    return Vec(v[1] * o.v[2] - v[2] * o.v[1], v[2] * o.v[0] - v[0] * o.v[2], v[0] * o.v[1] - v[1] * o.v[0]);
  }
  Vec operator+(const Vec& o) const { Vec r; for (int i = 0; i < N; ++i) r.v[i] = v[i] + o.v[i]; return r; }
  Vec operator-(const Vec& o) const { Vec r; for (int i = 0; i < N; ++i) r.v[i] = v[i] - o.v[i]; return r; }
  Vec operator-() const { Vec r; for (int i = 0; i < N; ++i) r.v[i] = -v[i]; return r; }
  Vec operator*(T s) const { Vec r; for (int i = 0; i < N; ++i) r.v[i] = v[i] * s; return r; }
  Vec operator/(T s) const { Vec r; for (int i = 0; i < N; ++i) r.v[i] = v[i] / s; return r; }
  Vec& operator+=(const Vec& o) { for (int i = 0; i < N; ++i) v[i] = v[i] + o.v[i]; return *this; }
  Vec& operator-=(const Vec& o) { for (int i = 0; i < N; ++i) v[i] = v[i] - o.v[i]; return *this; }
  Vec& operator*=(T s) { for (int i = 0; i < N; ++i) v[i] = v[i] * s; return *this; }
  bool operator==(const Vec& o) const { for (int i = 0; i < N; ++i) if (!(v[i] == o.v[i])) return false; return true; }
  bool operator!=(const Vec& o) const { return !(*this == o); }
};
template <class T, int N>
Vec<T, N> operator*(T s, const Vec<T, N>& a) { return a * s; }
template <class T, int N>
std::ostream& operator<<(std::ostream& os, const Vec<T, N>& a) {
  for (int i = 0; i < N; ++i) os << (i ? " " : "") << a.v[i];
  return os;
}

typedef Vec<double, 2> Vector2d;
typedef Vec<double, 3> Vector3d;
typedef Vec<int, 2> Vector2i;
typedef Vec<int, 3> Vector3i;

struct Quaterniond {
  double w_, x_, y_, z_;
  Quaterniond() : w_(1.0), x_(0.0), y_(0.0), z_(0.0) {}
  Quaterniond(double w, double x, double y, double z) : w_(w), x_(x), y_(y), z_(z) {}
  static Quaterniond Identity() { return Quaterniond(); }
  double w() const { return w_; } double x() const { return x_; } double y() const { return y_; } double z() const { return z_; }
  double& w() { return w_; } double& x() { return x_; } double& y() { return y_; } double& z() { return z_; }
  Vector3d vec() const { return Vector3d(x_, y_, z_); }
  Vector3d operator*(const Vector3d& v) const {     // QuaternionBase::_transformVector
    Vector3d uv = vec().cross(v);
    uv += uv;
    return v + uv * w_ + vec().cross(uv);
  }
  Quaterniond operator*(const Quaterniond& b) const {
    return Quaterniond(w_ * b.w_ - x_ * b.x_ - y_ * b.y_ - z_ * b.z_, w_ * b.x_ + x_ * b.w_ + y_ * b.z_ - z_ * b.y_,
                       w_ * b.y_ + y_ * b.w_ + z_ * b.x_ - x_ * b.z_, w_ * b.z_ + z_ * b.w_ + x_ * b.y_ - y_ * b.x_);
  }
  void normalize() {
    const double n = std::sqrt(w_ * w_ + x_ * x_ + y_ * y_ + z_ * z_);
    if (n > 0) { w_ /= n; x_ /= n; y_ /= n; z_ /= n; }
  }
};

// column-major dynamic 2-D array, (row, col) indexing like Eigen::ArrayXX*
template <class T>
struct ArrayXX {
  std::vector<T> a;
  std::ptrdiff_t rows_ = 0, cols_ = 0;
  ArrayXX() {}
  ArrayXX(std::ptrdiff_t r, std::ptrdiff_t c) : a((size_t)(r * c)), rows_(r), cols_(c) {}
  static ArrayXX Zero(std::ptrdiff_t r, std::ptrdiff_t c) { return ArrayXX(r, c); }
  std::ptrdiff_t rows() const { return rows_; }
  std::ptrdiff_t cols() const { return cols_; }
  // An out-of-range index is undefined behaviour in Eigen (the reference's unchecked path-map lookup, map_planexp.cpp:226-229, can
  // produce one for a voxel outside the map server's bounds); here it reads a zero and is counted instead of touching foreign memory.
  T& operator()(std::ptrdiff_t r, std::ptrdiff_t c) {
    if (r < 0 || r >= rows_ || c < 0 || c >= cols_) { ++oob_reads(); scratch_ = T(); return scratch_; }
    return a[(size_t)(c * rows_ + r)];
  }
  const T& operator()(std::ptrdiff_t r, std::ptrdiff_t c) const { return const_cast<ArrayXX*>(this)->operator()(r, c); }
  static long& oob_reads() { static long n = 0; return n; }
  T scratch_ = T();
};
typedef ArrayXX<int> ArrayXXi;
typedef ArrayXX<double> ArrayXXd;

}  // namespace Eigen
#endif  // SIPP_EIGEN_LITE_H_
