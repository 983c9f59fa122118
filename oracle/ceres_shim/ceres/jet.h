// TEST INFRASTRUCTURE — part of the CPU oracle, never linked into the product library.
//
// Minimal forward-mode dual number ("Jet") used by the oracle's autodiff cost
// functions.  Restates the semantics of upstream ceres-solver's ceres::Jet<T,N>
// [Ceres-upstream, version unpinned: CMakeLists.txt:7 of the reference only says
// find_package(Ceres REQUIRED)]: a scalar part `a` and an N-vector infinitesimal
// part `v`; every arithmetic operation propagates first derivatives exactly.
// Only the operations reached from the reference's functors are provided
// (snavely_reprojection_error.hh:38-118, hemisphere_radius.hh:18-28, and the
// rotation helpers in rotation.h).
#ifndef ORACLE_CERES_SHIM_JET_H_
#define ORACLE_CERES_SHIM_JET_H_

#include <cmath>

namespace ceres {

template <typename T, int N>
struct Jet {
  T a;
  T v[N];

  Jet() : a() {
    for (int i = 0; i < N; ++i) v[i] = T();
  }
  // NOLINTNEXTLINE: implicit by design, mirrors ceres::Jet.
  Jet(const T& value) : a(value) {
    for (int i = 0; i < N; ++i) v[i] = T();
  }
  Jet(const T& value, int k) : a(value) {
    for (int i = 0; i < N; ++i) v[i] = T();
    v[k] = T(1.0);
  }

  Jet& operator+=(const Jet& y) {
    a += y.a;
    for (int i = 0; i < N; ++i) v[i] += y.v[i];
    return *this;
  }
  Jet& operator-=(const Jet& y) {
    a -= y.a;
    for (int i = 0; i < N; ++i) v[i] -= y.v[i];
    return *this;
  }
  Jet& operator*=(const Jet& y) {
    *this = *this * y;
    return *this;
  }
  Jet& operator/=(const Jet& y) {
    *this = *this / y;
    return *this;
  }
};

// unary
template <typename T, int N>
inline Jet<T, N> operator+(const Jet<T, N>& f) {
  return f;
}
template <typename T, int N>
inline Jet<T, N> operator-(const Jet<T, N>& f) {
  Jet<T, N> h;
  h.a = -f.a;
  for (int i = 0; i < N; ++i) h.v[i] = -f.v[i];
  return h;
}

// Jet (+,-,*,/) Jet
template <typename T, int N>
inline Jet<T, N> operator+(const Jet<T, N>& f, const Jet<T, N>& g) {
  Jet<T, N> h;
  h.a = f.a + g.a;
  for (int i = 0; i < N; ++i) h.v[i] = f.v[i] + g.v[i];
  return h;
}
template <typename T, int N>
inline Jet<T, N> operator-(const Jet<T, N>& f, const Jet<T, N>& g) {
  Jet<T, N> h;
  h.a = f.a - g.a;
  for (int i = 0; i < N; ++i) h.v[i] = f.v[i] - g.v[i];
  return h;
}
template <typename T, int N>
inline Jet<T, N> operator*(const Jet<T, N>& f, const Jet<T, N>& g) {
  Jet<T, N> h;
  h.a = f.a * g.a;
  for (int i = 0; i < N; ++i) h.v[i] = f.a * g.v[i] + f.v[i] * g.a;
  return h;
}
template <typename T, int N>
inline Jet<T, N> operator/(const Jet<T, N>& f, const Jet<T, N>& g) {
  // (f/g)' = (f' - (f/g) g') / g ; upstream multiplies by 1/g.
  Jet<T, N> h;
  const T g_a_inverse = T(1.0) / g.a;
  const T f_a_by_g_a = f.a * g_a_inverse;
  h.a = f_a_by_g_a;
  for (int i = 0; i < N; ++i) h.v[i] = (f.v[i] - f_a_by_g_a * g.v[i]) * g_a_inverse;
  return h;
}

// Jet op scalar, scalar op Jet
template <typename T, int N>
inline Jet<T, N> operator+(const Jet<T, N>& f, T s) {
  Jet<T, N> h = f;
  h.a += s;
  return h;
}
template <typename T, int N>
inline Jet<T, N> operator+(T s, const Jet<T, N>& f) {
  Jet<T, N> h = f;
  h.a += s;
  return h;
}
template <typename T, int N>
inline Jet<T, N> operator-(const Jet<T, N>& f, T s) {
  Jet<T, N> h = f;
  h.a -= s;
  return h;
}
template <typename T, int N>
inline Jet<T, N> operator-(T s, const Jet<T, N>& f) {
  Jet<T, N> h;
  h.a = s - f.a;
  for (int i = 0; i < N; ++i) h.v[i] = -f.v[i];
  return h;
}
template <typename T, int N>
inline Jet<T, N> operator*(const Jet<T, N>& f, T s) {
  Jet<T, N> h;
  h.a = f.a * s;
  for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * s;
  return h;
}
template <typename T, int N>
inline Jet<T, N> operator*(T s, const Jet<T, N>& f) {
  Jet<T, N> h;
  h.a = f.a * s;
  for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * s;
  return h;
}
template <typename T, int N>
inline Jet<T, N> operator/(const Jet<T, N>& f, T s) {
  const T s_inverse = T(1.0) / s;
  Jet<T, N> h;
  h.a = f.a * s_inverse;
  for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * s_inverse;
  return h;
}
template <typename T, int N>
inline Jet<T, N> operator/(T s, const Jet<T, N>& g) {
  Jet<T, N> h;
  const T minus_s_g_a_inverse2 = -s / (g.a * g.a);
  h.a = s / g.a;
  for (int i = 0; i < N; ++i) h.v[i] = g.v[i] * minus_s_g_a_inverse2;
  return h;
}

// comparisons act on the scalar part only
#define ORACLE_JET_COMPARISON(op)                                         \
  template <typename T, int N>                                            \
  inline bool operator op(const Jet<T, N>& f, const Jet<T, N>& g) {       \
    return f.a op g.a;                                                    \
  }                                                                       \
  template <typename T, int N>                                            \
  inline bool operator op(const T& s, const Jet<T, N>& g) {               \
    return s op g.a;                                                      \
  }                                                                       \
  template <typename T, int N>                                            \
  inline bool operator op(const Jet<T, N>& f, const T& s) {               \
    return f.a op s;                                                      \
  }
ORACLE_JET_COMPARISON(<)
ORACLE_JET_COMPARISON(<=)
ORACLE_JET_COMPARISON(>)
ORACLE_JET_COMPARISON(>=)
ORACLE_JET_COMPARISON(==)
ORACLE_JET_COMPARISON(!=)
#undef ORACLE_JET_COMPARISON

// elementary functions
template <typename T, int N>
inline Jet<T, N> sqrt(const Jet<T, N>& f) {
  Jet<T, N> h;
  h.a = std::sqrt(f.a);
  const T two_a_inverse = T(1.0) / (T(2.0) * h.a);
  for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * two_a_inverse;
  return h;
}
template <typename T, int N>
inline Jet<T, N> cos(const Jet<T, N>& f) {
  Jet<T, N> h;
  h.a = std::cos(f.a);
  const T minus_sin = -std::sin(f.a);
  for (int i = 0; i < N; ++i) h.v[i] = minus_sin * f.v[i];
  return h;
}
template <typename T, int N>
inline Jet<T, N> sin(const Jet<T, N>& f) {
  Jet<T, N> h;
  h.a = std::sin(f.a);
  const T cos_a = std::cos(f.a);
  for (int i = 0; i < N; ++i) h.v[i] = cos_a * f.v[i];
  return h;
}
template <typename T, int N>
inline Jet<T, N> atan2(const Jet<T, N>& g, const Jet<T, N>& f) {
  // d atan2(g, f) = (f dg - g df) / (f^2 + g^2)
  Jet<T, N> h;
  const T tmp = T(1.0) / (f.a * f.a + g.a * g.a);
  h.a = std::atan2(g.a, f.a);
  for (int i = 0; i < N; ++i) h.v[i] = tmp * (-g.a * f.v[i] + f.a * g.v[i]);
  return h;
}
template <typename T, int N>
inline Jet<T, N> abs(const Jet<T, N>& f) {
  return f.a < T(0.0) ? -f : f;
}

// So that templated code can call ceres::sqrt etc. on plain doubles too.
using std::atan2;
using std::cos;
using std::sin;
using std::sqrt;

}  // namespace ceres

#endif  // ORACLE_CERES_SHIM_JET_H_
