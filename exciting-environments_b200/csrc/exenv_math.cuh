// exenv_math.cuh -- scalar building blocks shared by all kernels (device) and by the test-only host harness.
// Nothing here touches memory; every function is a pure scalar transform so that the same source is
// compiled by nvcc for sm_100a (product) and by g++ for tests/hostcheck (device-math host harness).
#pragma once
#include <type_traits>

#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>

#if defined(__CUDACC__)
#define EX_HD __host__ __device__ __forceinline__
#else
#define EX_HD inline
#endif

namespace exenv {

// ---- dtype-dispatched libm -----------------------------------------------------------------------------------
template <typename T>
struct Mth;

template <>
struct Mth<float> {
  static EX_HD float sin(float x) { return ::sinf(x); }
  static EX_HD float cos(float x) { return ::cosf(x); }
  static EX_HD void sincos(float x, float *s, float *c) {
#if defined(__CUDA_ARCH__)
    ::sincosf(x, s, c);
#else
    *s = ::sinf(x);
    *c = ::cosf(x);
#endif
  }
  static EX_HD float sqrt(float x) { return ::sqrtf(x); }
  static EX_HD float fmod(float x, float y) { return ::fmodf(x, y); }
  static EX_HD float atan2(float y, float x) { return ::atan2f(y, x); }
  static EX_HD float abs(float x) { return ::fabsf(x); }
};

template <>
struct Mth<double> {
  static EX_HD double sin(double x) { return ::sin(x); }
  static EX_HD double cos(double x) { return ::cos(x); }
  static EX_HD void sincos(double x, double *s, double *c) {
#if defined(__CUDA_ARCH__)
    ::sincos(x, s, c);
#else
    *s = ::sin(x);
    *c = ::cos(x);
#endif
  }
  static EX_HD double sqrt(double x) { return ::sqrt(x); }
  static EX_HD double fmod(double x, double y) { return ::fmod(x, y); }
  static EX_HD double atan2(double y, double x) { return ::atan2(y, x); }
  static EX_HD double abs(double x) { return ::fabs(x); }
};

template <typename T>
EX_HD T nan_of() {
  return (T)NAN;
}

// MinMaxNormalization (reference utils.py:13-17).  True divisions are kept: norm(denorm(z)) == z exactly at the
// default resets only holds with IEEE division.
template <typename T>
EX_HD T norm_mm(T x, T mn, T mx) {
  return (T)2 * (x - mn) / (mx - mn) - (T)1;
}
template <typename T>
EX_HD T denorm_mm(T z, T mn, T mx) {
  return (z + (T)1) / (T)2 * (mx - mn) + mn;
}

// Division by a loop-invariant divisor.  {d, r = RN(1/d)} is prepared once; q = RN(x*r), rem = x - d*q (exact, one
// FMA), q' = RN(q + rem*r) is the correctly rounded x/d (Markstein) outside of the denormal / overflow corner cases
// the full IEEE sequence guards with its slow path.  Used by the compute-bound sim_ahead kernel, where the plain
// `/` costs ~13 instructions per use; the bandwidth-bound step kernels keep the plain operator.
template <typename T>
struct InvDiv {
  T d, r;
};
template <typename T>
EX_HD InvDiv<T> make_invdiv(T d) {
  InvDiv<T> v;
  v.d = d;
  v.r = (T)1 / d;
  return v;
}
EX_HD float fma_(float a, float b, float c) { return ::fmaf(a, b, c); }
EX_HD double fma_(double a, double b, double c) { return ::fma(a, b, c); }
template <bool FAST, typename T>
EX_HD T div_inv(T x, const InvDiv<T> &c) {
  if constexpr (FAST) {
    const T q = x * c.r;
    const T rem = fma_(-c.d, q, x);
    return fma_(rem, c.r, q);
  } else {
    return x / c.d;
  }
}
// x / d as x * RN(1 / d): one rounding more than div_inv (<= 1.5 ulp from the quotient).  The FAST right-hand sides of
// the instruction-bound environments use it where the quotient only feeds a derivative.
template <typename T>
EX_HD T div_mul(T x, const InvDiv<T> &c) {
  return x * c.r;
}
// 1 / d for a state-dependent denominator without the IEEE division's range checks and slow path: the hardware
// approximation refined by Newton steps (float: one, about 1 ulp; double: two from the 2^-23 seed).  The caller
// guarantees a normal, non-zero d (physical inertia terms); inf / nan propagate.
EX_HD float rcp_nr(float d) {
#if defined(__CUDA_ARCH__)
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
  return ::fmaf(r, ::fmaf(-d, r, 1.0f), r);
#else
  return 1.0f / d;
#endif
}
EX_HD double rcp_nr(double d) {
#if defined(__CUDA_ARCH__)
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  r = ::fma(r, ::fma(-d, r, 1.0), r);
  return ::fma(r, ::fma(-d, r, 1.0), r);
#else
  return 1.0 / d;
#endif
}
template <bool FAST, typename T>
EX_HD T norm_inv(T x, T mn, const InvDiv<T> &span) {
  return div_inv<FAST, T>((T)2 * (x - mn), span) - (T)1;
}

// The same normalisation as ONE fused multiply-add, 2 (x - mn) / (mx - mn) - 1 = c1 x + c0 with c1 = 2 / (mx - mn) and
// c0 = -2 mn / (mx - mn) - 1 prepared by true divisions (c0 is exactly 0 for the symmetric ranges, exactly -1 for
// ranges starting at 0).  One rounding of the exact c1 x + c0 instead of three roundings: within 3 ulp of 1 (3.6e-7
// in fp32) of the reference's value (observations live in [-1, 1]; the north star's tolerance is 1e-6).  Used by the time-major
// trajectory kernels, where the correctly rounded form costs six instructions per observation entry; the truncation
// flags those kernels emit are computed from the very observation values they return.
template <typename T>
struct NormFma {
  T c1, c0;
};
template <typename T>
EX_HD NormFma<T> make_normfma(T mn, T mx) {
  NormFma<T> n;
  n.c1 = (T)2 / (mx - mn);
  n.c0 = -(((T)2 * mn) / (mx - mn)) - (T)1;
  return n;
}
template <typename T>
EX_HD T norm_fma(T x, const NormFma<T> &n) {
  return fma_(x, n.c1, n.c0);
}

// Single IEEE operations the compiler must not contract into a fused multiply-add, whatever -fmad says: the truncation
// flags are bit-exact statements about un-fused expressions (the host harness is compiled with -ffp-contract=off).
EX_HD float mul_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fmul_rn(a, b);
#else
  return a * b;
#endif
}
EX_HD double mul_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dmul_rn(a, b);
#else
  return a * b;
#endif
}
EX_HD float add_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fadd_rn(a, b);
#else
  return a + b;
#endif
}
EX_HD double add_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dadd_rn(a, b);
#else
  return a + b;
#endif
}

// truncated = abs(2 (x - min) / (max - min) - 1) > 1 as two comparisons.  Every operation of the reference expression
// is a monotone (non-decreasing) function of x under round-to-nearest, so the flag is "x >= hi_first or x <= lo_last"
// for two thresholds found ONCE on the host by bisection over the floating-point numbers with the reference expression
// itself (make_trunc_thresholds): bit-exact by construction, two compares instead of a correctly rounded division per
// observation entry.  NaN compares false on both sides, like abs(NaN) > 1.
template <typename T>
struct TruncThr {
  T lo_last, hi_first;
};
template <typename T>
EX_HD bool trunc_by_threshold(T x, const TruncThr<T> &t) {
  return (x >= t.hi_first) || (x <= t.lo_last);
}
namespace detail {
inline uint32_t ord_of(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
inline float of_ord(uint32_t o, float) {
  const uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
  float f;
  memcpy(&f, &u, 4);
  return f;
}
inline uint64_t ord_of(double f) {
  uint64_t u;
  memcpy(&u, &f, 8);
  return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}
inline double of_ord(uint64_t o, double) {
  const uint64_t u = (o & 0x8000000000000000ull) ? (o & 0x7fffffffffffffffull) : ~o;
  double f;
  memcpy(&f, &u, 8);
  return f;
}
}  // namespace detail
// Host only (compiled with -ffp-contract=off).  Returns false when the range is degenerate (span not a positive finite
// number): the caller then takes the path that evaluates the expression itself.
template <typename T>
inline bool make_trunc_thresholds(T mn, T mx, TruncThr<T> &out) {
  const T span = mx - mn;
  if (!(span > (T)0) || !(span < std::numeric_limits<T>::infinity()) || mn != mn) return false;
  auto g = [&](T x) -> T {
    volatile T d = x - mn;
    volatile T n = (T)2 * d;
    volatile T q = n / span;
    volatile T r = q - (T)1;
    return r;
  };
  const T inf = std::numeric_limits<T>::infinity();
  auto lo_o = detail::ord_of(-inf), hi_o = detail::ord_of(inf);
  using U = decltype(lo_o);
  {  // smallest x with g(x) > 1 (exists: g(+inf) = +inf; g(-inf) = -inf)
    U a = lo_o, b = hi_o;  // g(a) > 1 false, g(b) > 1 true
    while (b - a > 1) {
      const U m = a + (b - a) / 2;
      if (g(detail::of_ord(m, (T)0)) > (T)1) b = m;
      else a = m;
    }
    out.hi_first = detail::of_ord(b, (T)0);
  }
  {  // largest x with g(x) < -1
    U a = lo_o, b = hi_o;  // g(a) < -1 true, g(b) < -1 false
    while (b - a > 1) {
      const U m = a + (b - a) / 2;
      if (g(detail::of_ord(m, (T)0)) < (T)-1) a = m;
      else b = m;
    }
    out.lo_last = detail::of_ord(a, (T)0);
  }
  return true;
}

// C fmod(x, y) for y > 0, exactly, without libm's iterative reduction when the quotient is moderate (angles: a few
// turns at most).  q' = trunc(RN(|x| * RN(1/y)) * (1 + 2^-21)) is the true integer quotient q or q + 1 (the upward bias
// beats the two roundings of the product; it needs |x| / y < 2^20); |x| - q' * y is then a multiple of ulp(y) of
// magnitude below y, so the FMA delivers it exactly, and so does the single correction r + y.  The sign of x is
// restored last (fmod(-2 pi, 2 pi) = -0.0).  Larger quotients take libm.
template <typename T>
struct FmodK;
template <>
struct FmodK<float> {
  static constexpr float BIAS = 1.000000476837158203125f;  // 1 + 2^-21
  static constexpr float QMAX = 1048576.0f;                  // 2^20
};
template <>
struct FmodK<double> {
  static constexpr double BIAS = 1.0 + 8.8817841970012523e-16;  // 1 + 2^-50
  static constexpr double QMAX = 562949953421312.0;              // 2^49
};
EX_HD float trunc_(float x) { return ::truncf(x); }
EX_HD double trunc_(double x) { return ::trunc(x); }
EX_HD float copysign_(float x, float s) { return ::copysignf(x, s); }
EX_HD double copysign_(double x, double s) { return ::copysign(x, s); }
template <typename T>
EX_HD T fmod_pos(T x, T y, T inv_y) {
  const T ax = Mth<T>::abs(x);
  if (!(ax < FmodK<T>::QMAX * y)) return Mth<T>::fmod(x, y);  // huge, inf or nan
  // no shortcut for |x| < y: q is 0 or 1 there and the same two exact operations return |x| (branch-free)
  const T q = trunc_(ax * inv_y * FmodK<T>::BIAS);
  T r = fma_(-q, y, ax);
  if (r < (T)0) r += y;
  return copysign_(r, x);
}

// ---- two environments per value: sm_100a packed fp32 arithmetic ------------------------------------------------
// FFMA2 / FMUL2 / FADD2 perform two IEEE fp32 operations per instruction (same results as the scalar instructions, at
// the same arithmetic rate, for half the issue slots); a scalar operand is broadcast (`R.F32`, immediates and uniform
// registers included), so shared parameters and solver coefficients stay scalars.  The environment math is written
// against a scalar type T; instantiating it with f32x2 steps two environments at once.
struct f32x2 {
  float x, y;
  f32x2() = default;
  EX_HD constexpr f32x2(float s) : x(s), y(s) {}
  EX_HD constexpr f32x2(double s) : x((float)s), y((float)s) {}
  EX_HD constexpr f32x2(int s) : x((float)s), y((float)s) {}
  EX_HD constexpr f32x2(float a, float b) : x(a), y(b) {}
};
#if defined(__CUDACC__)
EX_HD f32x2 operator+(f32x2 a, f32x2 b) {
#if defined(__CUDA_ARCH__)
  const float2 r = __fadd2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y));
  return f32x2(r.x, r.y);
#else
  return f32x2(a.x + b.x, a.y + b.y);
#endif
}
EX_HD f32x2 operator-(f32x2 a) { return f32x2(-a.x, -a.y); }
EX_HD f32x2 operator-(f32x2 a, f32x2 b) { return a + (-b); }
EX_HD f32x2 operator*(f32x2 a, f32x2 b) {
#if defined(__CUDA_ARCH__)
  const float2 r = __fmul2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y));
  return f32x2(r.x, r.y);
#else
  return f32x2(a.x * b.x, a.y * b.y);
#endif
}
EX_HD f32x2 operator/(f32x2 a, f32x2 b) { return f32x2(a.x / b.x, a.y / b.y); }  // set-up code only
EX_HD f32x2 &operator+=(f32x2 &a, f32x2 b) { return a = a + b; }
EX_HD f32x2 fma_(f32x2 a, f32x2 b, f32x2 c) {
#if defined(__CUDA_ARCH__)
  const float2 r = __ffma2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y), make_float2(c.x, c.y));
  return f32x2(r.x, r.y);
#else
  return f32x2(::fmaf(a.x, b.x, c.x), ::fmaf(a.y, b.y, c.y));
#endif
}
// (z + 1) / 2 * (max - min) + min: the halving as a multiplication (the same value; a packed division would not be
// strength-reduced by the compiler)
template <>
EX_HD f32x2 denorm_mm<f32x2>(f32x2 z, f32x2 mn, f32x2 mx) {
  return (z + f32x2(1.0f)) * f32x2(0.5f) * (mx - mn) + mn;
}
#endif

#if defined(__CUDACC__)
// ---- packed transcendental helpers: what the right-hand sides of pendulum / fluid tank / cart-pole / acrobot need ----
EX_HD int f32_bits(float x) {
#if defined(__CUDA_ARCH__)
  return __float_as_int(x);
#else
  int i;
  memcpy(&i, &x, 4);
  return i;
#endif
}
EX_HD float f32_from_bits(int i) {
#if defined(__CUDA_ARCH__)
  return __int_as_float(i);
#else
  float x;
  memcpy(&x, &i, 4);
  return x;
#endif
}
// sin and cos of two angles at once.  Cody-Waite reduction by pi/2 with three FMAs on the packed pipe (k = rint(2x/pi)
// by the magic-number addition; pi/2 = HI + MID + LO to ~2^-76, the products exact inside the FMAs), both minimax
// polynomials of the Cephes single-precision sinf / cosf on |r| <= pi/4 as FFMA2 chains, and only the quadrant fix-up
// (two selects and two sign flips) per lane: ~16 packed + ~12 scalar instructions for two angles, against ~25 scalar
// instructions per angle for CUDA's sincosf.  Maximum error 1.6 ulp for |x| < 1e5 (tests/test_hostcheck.py; CUDA's
// sinf / cosf: 1.5 ulp); larger arguments, inf and nan take libm per lane.
// The polynomial path alone (every argument; garbage beyond the reduction's range).
EX_HD void sincos2_poly(f32x2 x, f32x2 *sn, f32x2 *cs) {
  const f32x2 kf = fma_(x, f32x2(0.63661977236758134f), f32x2(12582912.0f));
  const f32x2 k = kf + f32x2(-12582912.0f);
  f32x2 r = fma_(k, f32x2(-1.5707963705062866f), x);
  r = fma_(k, f32x2(4.371138828673793e-08f), r);
  r = fma_(k, f32x2(1.7151245100058819e-15f), r);
  const f32x2 z = r * r;
  f32x2 ps = fma_(z, f32x2(-1.9515295891e-4f), f32x2(8.3321608736e-3f));
  ps = fma_(ps, z, f32x2(-1.6666654611e-1f));
  const f32x2 s = fma_(r * z, ps, r);
  f32x2 pc = fma_(z, f32x2(2.443315711809948e-5f), f32x2(-1.388731625493765e-3f));
  pc = fma_(pc, z, f32x2(4.166664568298827e-2f));
  const f32x2 c = fma_(z * z, pc, fma_(z, f32x2(-0.5f), f32x2(1.0f)));
  auto lane = [](float kfl, float sl, float cl, float *so, float *co) {
    const int q = f32_bits(kfl);  // the low mantissa bits of x * 2/pi + 1.5 * 2^23 hold rint(x * 2/pi)
    const float a = (q & 1) ? cl : sl, b = (q & 1) ? sl : cl;
    *so = f32_from_bits(f32_bits(a) ^ ((q & 2) << 30));
    *co = f32_from_bits(f32_bits(b) ^ (((q + 1) & 2) << 30));
  };
  lane(kf.x, s.x, c.x, &sn->x, &cs->x);
  lane(kf.y, s.y, c.y, &sn->y, &cs->y);
}
#if defined(__CUDACC__)
// Large arguments (|x| >= 1e5), inf, nan: libm per lane.  Out of line: the Payne-Hanek reduction of sincosf is a few
// hundred instructions, and sincos2 is inlined 7-14 times per solver step -- inline, these never-executed blocks made
// the step bodies of the pendulum / cart-pole / acrobot kernels 70-140 KB, several times the instruction cache.  Takes
// the two angles only (two argument registers to set up at every call site) and redoes the polynomial path for a lane
// that is in range: the same operations, the same value.
static __device__ __noinline__ float4 sincos2_slow(float x, float y) {
  f32x2 s, c;
  sincos2_poly(f32x2(x, y), &s, &c);
  if (!(::fabsf(x) < 100000.0f)) ::sincosf(x, &s.x, &c.x);
  if (!(::fabsf(y) < 100000.0f)) ::sincosf(y, &s.y, &c.y);
  return make_float4(s.x, s.y, c.x, c.y);
}
#endif
EX_HD void sincos2(f32x2 x, f32x2 *sn, f32x2 *cs) {
  sincos2_poly(x, sn, cs);
  // beyond the reduction's range (or inf / nan): libm, per lane.  One warp-wide vote decides whether anybody has to go
  // there, so that the common case costs a uniform branch instead of four divergent ones.
  const bool big = !(::fabsf(x.x) < 100000.0f) || !(::fabsf(x.y) < 100000.0f);
#if defined(__CUDA_ARCH__)
  if (__any_sync(__activemask(), big)) {
    const float4 r = sincos2_slow(x.x, x.y);
    sn->x = r.x;
    sn->y = r.y;
    cs->x = r.z;
    cs->y = r.w;
  }
#else
  if (big) {
    if (!(::fabsf(x.x) < 100000.0f)) {
      sn->x = ::sinf(x.x);
      cs->x = ::cosf(x.x);
    }
    if (!(::fabsf(x.y) < 100000.0f)) {
      sn->y = ::sinf(x.y);
      cs->y = ::cosf(x.y);
    }
  }
#endif
}
template <>
struct Mth<f32x2> {
  static EX_HD f32x2 sin(f32x2 x) {
    f32x2 s, c;
    sincos2(x, &s, &c);
    return s;
  }
  static EX_HD f32x2 cos(f32x2 x) {
    f32x2 s, c;
    sincos2(x, &s, &c);
    return c;
  }
  static EX_HD void sincos(f32x2 x, f32x2 *s, f32x2 *c) { sincos2(x, s, c); }
  static EX_HD f32x2 sqrt(f32x2 x) { return f32x2(::sqrtf(x.x), ::sqrtf(x.y)); }
  static EX_HD f32x2 abs(f32x2 x) { return f32x2(::fabsf(x.x), ::fabsf(x.y)); }
};
// Newton-refined reciprocal: the two hardware approximations per lane, the refinement packed
EX_HD f32x2 rcp_nr(f32x2 d) {
#if defined(__CUDA_ARCH__)
  f32x2 r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.x) : "f"(d.x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.y) : "f"(d.y));
  return fma_(r, fma_(-d, r, f32x2(1.0f)), r);
#else
  return f32x2(1.0f / d.x, 1.0f / d.y);
#endif
}
#endif  // __CUDACC__

EX_HD float max_(float a, float b) { return ::fmaxf(a, b); }
EX_HD double max_(double a, double b) { return ::fmax(a, b); }
EX_HD float min_(float a, float b) { return ::fminf(a, b); }
EX_HD double min_(double a, double b) { return ::fmin(a, b); }

// max(x, 0) as the reference writes it for the tank height (jnp.clip(h, 0): NaN stays NaN)
EX_HD float relu_(float x) { return x < 0.0f ? 0.0f : x; }
EX_HD double relu_(double x) { return x < 0.0 ? 0.0 : x; }
#if defined(__CUDACC__)
EX_HD f32x2 relu_(f32x2 x) { return f32x2(relu_(x.x), relu_(x.y)); }
#endif

// jnp `%`: C fmod followed by the sign correction (jax.numpy.remainder); SURVEY Appendix A.
template <typename T>
EX_HD T pymod(T x, T y) {
  T r = Mth<T>::fmod(x, y);
  if (r != (T)0 && ((r < (T)0) != (y < (T)0))) r += y;
  return r;
}
// pymod(x, 2 pi) through fmod_pos (identical values)
template <typename T>
EX_HD T pymod_two_pi(T x) {
  const T y = (T)(2 * 3.141592653589793);
  T r = fmod_pos<T>(x, y, (T)1 / y);
  if (r < (T)0) r += y;
  return r;
}

template <typename T>
struct Cst {
  static constexpr double PI_D = 3.141592653589793;
  static EX_HD T pi() { return (T)PI_D; }
  static EX_HD T two_pi() { return (T)(2 * PI_D); }
  static EX_HD T half_pi() { return (T)(PI_D / 2); }
};

// theta <- ((theta + pi) % (2 pi)) - pi   (pendulum_env.py:188, cart_pole_env.py:229, acrobot_env.py:247-248,
// pmsm_env.py:571)
template <typename T>
EX_HD T wrap_pi(T th) {
  return pymod_two_pi<T>(th + Cst<T>::pi()) - Cst<T>::pi();
}

template <typename T>
EX_HD T sign_of(T v) {
  return (T)((v > (T)0) - (v < (T)0));
}
#if defined(__CUDACC__)
template <>
EX_HD f32x2 wrap_pi<f32x2>(f32x2 th) {
  return f32x2(wrap_pi<float>(th.x), wrap_pi<float>(th.y));
}
template <>
EX_HD f32x2 sign_of<f32x2>(f32x2 v) {
  return f32x2(sign_of<float>(v.x), sign_of<float>(v.y));
}
#endif

// Tableau coefficients a_ij (row-major below the diagonal; row 7 = the solution weights b_j).  fp32 / packed fp32: literals
// (32-bit immediates of FFMA / FFMA2).  fp64 on the device: a __constant__ table -- DFMA takes a constant-bank operand but
// no 64-bit immediate, and as literals the 21 coefficients cost two moves each per solver step (88 of the 267 warp
// instructions per env-step of the fp64 mass-spring-damper trajectory kernel, profiles/README.md).
#define EXENV_TSIT5_A                                                                                                 \
  {0.161, -0.008480655492356989, 0.335480655492357, 2.8971530571054935, -6.359448489975075, 4.3622954328695815,       \
   5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525, 5.86145544294642,                \
   -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383, 0.09646076681806523, 0.01,       \
   0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774}
#if defined(__CUDACC__)
static __constant__ double k_tsit5_a_f64[21] = EXENV_TSIT5_A;
#endif
template <typename T>
EX_HD T tsit5_a(int i) {
#if defined(__CUDA_ARCH__)
  if constexpr (std::is_same<T, double>::value) return k_tsit5_a_f64[i];
#endif
  constexpr double v[21] = EXENV_TSIT5_A;
  return (T)v[i];
}

// ---- Tsit5 (diffrax.Tsit5; Tsitouras 2011).  SURVEY Appendix B.2 --------------------------------------------
// Stage increments are k_i = f_i * dt (diffrax "vf_prod"); stage states y0 + sum_j a_ij k_j; the solution is the
// stage-7 state (FSAL/SSAL tableau).  `rhs(stage, y, f)` evaluates the vector field for stage index 1..6
// (0-based; stage 0's derivative f0 is supplied by the caller: FSAL carry or a fresh evaluation).
// k[0..6][N] are returned for dense output; f_last = f(t0 + dt, y1).
template <typename T, int N, class F>
EX_HD void tsit5_step(const T *y0, const T *f0, T dt, F &&rhs, T *y1, T *f_last, T (*k)[N]) {
  const T a21 = tsit5_a<T>(0);
  const T a31 = tsit5_a<T>(1), a32 = tsit5_a<T>(2);
  const T a41 = tsit5_a<T>(3), a42 = tsit5_a<T>(4), a43 = tsit5_a<T>(5);
  const T a51 = tsit5_a<T>(6), a52 = tsit5_a<T>(7), a53 = tsit5_a<T>(8), a54 = tsit5_a<T>(9);
  const T a61 = tsit5_a<T>(10), a62 = tsit5_a<T>(11), a63 = tsit5_a<T>(12), a64 = tsit5_a<T>(13), a65 = tsit5_a<T>(14);
  const T a71 = tsit5_a<T>(15), a72 = tsit5_a<T>(16), a73 = tsit5_a<T>(17), a74 = tsit5_a<T>(18), a75 = tsit5_a<T>(19),
          a76 = tsit5_a<T>(20);
  T ys[N], f[N];
#pragma unroll
  for (int i = 0; i < N; ++i) k[0][i] = f0[i] * dt;
#pragma unroll
  for (int i = 0; i < N; ++i) ys[i] = y0[i] + (a21 * k[0][i]);
  rhs(1, ys, f);
#pragma unroll
  for (int i = 0; i < N; ++i) k[1][i] = f[i] * dt;
#pragma unroll
  for (int i = 0; i < N; ++i) ys[i] = y0[i] + (a31 * k[0][i] + a32 * k[1][i]);
  rhs(2, ys, f);
#pragma unroll
  for (int i = 0; i < N; ++i) k[2][i] = f[i] * dt;
#pragma unroll
  for (int i = 0; i < N; ++i) ys[i] = y0[i] + (a41 * k[0][i] + a42 * k[1][i] + a43 * k[2][i]);
  rhs(3, ys, f);
#pragma unroll
  for (int i = 0; i < N; ++i) k[3][i] = f[i] * dt;
#pragma unroll
  for (int i = 0; i < N; ++i) ys[i] = y0[i] + (a51 * k[0][i] + a52 * k[1][i] + a53 * k[2][i] + a54 * k[3][i]);
  rhs(4, ys, f);
#pragma unroll
  for (int i = 0; i < N; ++i) k[4][i] = f[i] * dt;
#pragma unroll
  for (int i = 0; i < N; ++i)
    ys[i] = y0[i] + (a61 * k[0][i] + a62 * k[1][i] + a63 * k[2][i] + a64 * k[3][i] + a65 * k[4][i]);
  rhs(5, ys, f);
#pragma unroll
  for (int i = 0; i < N; ++i) k[5][i] = f[i] * dt;
#pragma unroll
  for (int i = 0; i < N; ++i)
    ys[i] = y0[i] + (a71 * k[0][i] + a72 * k[1][i] + a73 * k[2][i] + a74 * k[3][i] + a75 * k[4][i] + a76 * k[5][i]);
  rhs(6, ys, f);
#pragma unroll
  for (int i = 0; i < N; ++i) {
    k[6][i] = f[i] * dt;
    y1[i] = ys[i];
    f_last[i] = f[i];
  }
}

// The same step in "derivative form", used by the trajectory kernels: the stage derivatives f_j are kept instead of the
// increments k_j = f_j dt, and a stage state is y0 + dt (sum_j a_ij f_j) -- one multiplication by dt per stage and
// component instead of one per derivative (7 fewer operations per component and step).  Same tableau, same stage
// order; the value differs from diffrax's y0 + sum_j a_ij (f_j dt) by the rounding of the increment (a few ulp of an
// increment that is itself ~dt |f| of the state: far inside the 1e-6 per-step tolerance; tests/test_hostcheck.py and
// tests/test_sim_gpu.py hold it against the oracle, which keeps diffrax's form).
//   f[0][.] = FSAL derivative at (t0, y0) on entry; f[1..6][.] are filled; f[6] = f(t0 + dt, y1).
template <typename T, int N, class F>
EX_HD void tsit5_step_f(const T *y0, T dt, F &&rhs, T *y1, T (*f)[N]) {
  const T a21 = tsit5_a<T>(0);
  const T a31 = tsit5_a<T>(1), a32 = tsit5_a<T>(2);
  const T a41 = tsit5_a<T>(3), a42 = tsit5_a<T>(4), a43 = tsit5_a<T>(5);
  const T a51 = tsit5_a<T>(6), a52 = tsit5_a<T>(7), a53 = tsit5_a<T>(8), a54 = tsit5_a<T>(9);
  const T a61 = tsit5_a<T>(10), a62 = tsit5_a<T>(11), a63 = tsit5_a<T>(12), a64 = tsit5_a<T>(13), a65 = tsit5_a<T>(14);
  const T a71 = tsit5_a<T>(15), a72 = tsit5_a<T>(16), a73 = tsit5_a<T>(17), a74 = tsit5_a<T>(18), a75 = tsit5_a<T>(19),
          a76 = tsit5_a<T>(20);
  T ys[N];
#pragma unroll
  for (int i = 0; i < N; ++i) ys[i] = fma_(dt, a21 * f[0][i], y0[i]);
  rhs(1, ys, f[1]);
#pragma unroll
  for (int i = 0; i < N; ++i) ys[i] = fma_(dt, fma_(a32, f[1][i], a31 * f[0][i]), y0[i]);
  rhs(2, ys, f[2]);
#pragma unroll
  for (int i = 0; i < N; ++i) ys[i] = fma_(dt, fma_(a43, f[2][i], fma_(a42, f[1][i], a41 * f[0][i])), y0[i]);
  rhs(3, ys, f[3]);
#pragma unroll
  for (int i = 0; i < N; ++i)
    ys[i] = fma_(dt, fma_(a54, f[3][i], fma_(a53, f[2][i], fma_(a52, f[1][i], a51 * f[0][i]))), y0[i]);
  rhs(4, ys, f[4]);
#pragma unroll
  for (int i = 0; i < N; ++i)
    ys[i] = fma_(dt, fma_(a65, f[4][i], fma_(a64, f[3][i], fma_(a63, f[2][i], fma_(a62, f[1][i], a61 * f[0][i])))),
                 y0[i]);
  rhs(5, ys, f[5]);
#pragma unroll
  for (int i = 0; i < N; ++i)
    y1[i] = fma_(dt,
                 fma_(a76, f[5][i],
                      fma_(a75, f[4][i], fma_(a74, f[3][i], fma_(a73, f[2][i], fma_(a72, f[1][i], a71 * f[0][i]))))),
                 y0[i]);
  rhs(6, y1, f[6]);
}

// Dense-output weights b_i(theta) of diffrax's _Tsit5Interpolation.
template <typename T>
EX_HD void tsit5_dense_weights(T t, T *b) {
  const T t2 = t * t;
  b[0] = (T)-1.0530884977290216 * t * (t - (T)1.3299890189751412) * (t2 - (T)1.4364028541716351 * t + (T)0.7139816917074209);
  b[1] = (T)0.1017 * t2 * (t2 - (T)2.1966568338249754 * t + (T)1.2949852507374631);
  b[2] = (T)2.490627285651252793 * t2 * (t2 - (T)2.38535645472061657 * t + (T)1.57803468208092486);
  b[3] = (T)-16.54810288924490272 * (t - (T)1.21712927295533244) * (t - (T)0.61620406037800089) * t2;
  b[4] = (T)47.37952196281928122 * (t - (T)1.203071208372362603) * (t - (T)0.658047292653547382) * t2;
  b[5] = (T)-34.87065786149660974 * (t - (T)1.2) * (t - (T)0.666666666666666667) * t2;
  b[6] = (T)2.5 * (t - (T)1.0) * (t - (T)0.6) * t2;
}

}  // namespace exenv
