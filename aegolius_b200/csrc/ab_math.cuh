// ab_math.cuh — the arithmetic layer of the SDF interpreter.
//
//   Pack<T,W>  : W grid points per thread held in registers (W consecutive points along the fastest grid axis, so one
//                thread owns one 128-bit store). For T=float the add/mul/fma lanes are issued as packed f32x2
//                instructions (FADD2/FMUL2/FFMA2 on sm_100a): one issue slot does two points. Measured on B200
//                (profiles/r01_ubench_pipes.txt): FFMA2 issues at the same rate as FFMA, i.e. 2x the FP32 FMA work per
//                issue slot, which is the binding resource for deep trees.
//   Dual<P,K>  : forward-mode dual number over a Pack: value + K tangents. K=3 seeds d/dx,d/dy,d/dz (analytic SDF
//                gradient, replaces np.gradient-based from_sdf, vector_functions.py:130-139); K=1 seeds one geometry
//                parameter (replaces jacfwd(geometry, argnums=k), Code/examples/autodiff/gradient_map_3D.py:84).
//
// Every interpreter op is written once, generically, over a "scalar" S that is either a Pack or a Dual<Pack>.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#define AB_DEV __device__ __forceinline__

#ifndef AB_USE_F32X2
#define AB_USE_F32X2 1
#endif
#ifndef AB_FAST_SQRT
#define AB_FAST_SQRT 1 /* sqrt.approx.ftz.f32 (1 MUFU, <=1 ulp) instead of the IEEE sequence (~11 issue slots) */
#endif
#ifndef AB_FAST_ATAN2
#define AB_FAST_ATAN2 1 /* packed polynomial atan2 for fp32 (see atan2_ below) instead of atan2f */
#endif
#ifndef AB_FAST_SINCOS
#define AB_FAST_SINCOS 1 /* packed Cody-Waite + polynomial sincos for fp32 (see sincos_ below) instead of sincosf per lane */
#endif
#ifndef AB_FAST_DIV
#define AB_FAST_DIV 1 /* rcp.approx-based division (<=2 ulp) instead of the IEEE sequence (~14 issue slots) */
#endif

namespace ab {

// ------------------------------------------------------------------------------------------------------------------
// scalar helpers (T = float | double)

AB_DEV float s_sqrt(float a) {
#if AB_FAST_SQRT
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
  return r;
#else
  return sqrtf(a);
#endif
}
AB_DEV double s_sqrt(double a) { return sqrt(a); }
AB_DEV float s_div(float a, float b) {
#if AB_FAST_DIV
  return __fdividef(a, b);
#else
  return a / b;
#endif
}
AB_DEV double s_div(double a, double b) { return a / b; }
AB_DEV float s_rcp(float a) {
#if AB_FAST_DIV
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
  return r;
#else
  return 1.0f / a;
#endif
}
AB_DEV double s_rcp(double a) { return 1.0 / a; }
AB_DEV float s_fma(float a, float b, float c) { return fmaf(a, b, c); }
AB_DEV double s_fma(double a, double b, double c) { return fma(a, b, c); }
AB_DEV float s_min(float a, float b) { return fminf(a, b); }
AB_DEV double s_min(double a, double b) { return fmin(a, b); }
AB_DEV float s_max(float a, float b) { return fmaxf(a, b); }
AB_DEV double s_max(double a, double b) { return fmax(a, b); }
AB_DEV float s_abs(float a) { return fabsf(a); }
AB_DEV double s_abs(double a) { return fabs(a); }
AB_DEV float s_floor(float a) { return floorf(a); }
AB_DEV double s_floor(double a) { return floor(a); }
// The big libm bodies (sincos with its Payne-Hanek slow path, atan2, exp, pow) are deliberately NOT inlined: inlined
// W times at every call site they made the executed path of a deep tree larger than the instruction cache
// ("no_instructions" was the top stall of the C5 kernel, profiles/r01_c5_dual_kernel_ncu_summary.json). One shared copy
// per kernel, results returned in registers.
#ifndef AB_OUTLINE_LIBM
#define AB_OUTLINE_LIBM 1
#endif
#if AB_OUTLINE_LIBM
#define AB_LIBM static __device__ __noinline__
#else
#define AB_LIBM static __device__ __forceinline__
#endif
AB_LIBM float2 ab_sincosf(float a) {
  float2 r;
  sincosf(a, &r.x, &r.y);
  return r;
}
AB_LIBM double2 ab_sincos(double a) {
  double2 r;
  sincos(a, &r.x, &r.y);
  return r;
}
AB_LIBM float ab_atan2f(float y, float x) { return atan2f(y, x); }
AB_LIBM double ab_atan2(double y, double x) { return atan2(y, x); }
AB_LIBM float ab_expf(float a) { return expf(a); }
AB_LIBM double ab_exp(double a) { return exp(a); }
AB_LIBM float ab_powf(float a, float b) { return powf(a, b); }
AB_LIBM double ab_pow(double a, double b) { return pow(a, b); }
AB_DEV void s_sincos(float a, float& s, float& c) {
  const float2 r = ab_sincosf(a);
  s = r.x;
  c = r.y;
}
AB_DEV void s_sincos(double a, double& s, double& c) {
  const double2 r = ab_sincos(a);
  s = r.x;
  c = r.y;
}
AB_DEV float s_atan2(float y, float x) { return ab_atan2f(y, x); }
AB_DEV double s_atan2(double y, double x) { return ab_atan2(y, x); }
AB_DEV float s_exp(float a) { return ab_expf(a); }
AB_DEV double s_exp(double a) { return ab_exp(a); }
AB_DEV float s_pow(float a, float b) { return ab_powf(a, b); }
AB_DEV double s_pow(double a, double b) { return ab_pow(a, b); }
AB_DEV float s_log(float a) { return logf(a); }
AB_DEV double s_log(double a) { return log(a); }

// np.mod semantics (floor-mod: result in [0, b) for b > 0, in (b, 0] for b < 0 — the sign of the divisor):
// r = a - floor(a/b)*b evaluated with one FMA (exact when the quotient is right) and repaired when the rounded quotient
// is off by one.
template <typename T>
AB_DEV T s_mod(T a, T b) {
  T k = s_floor(s_div(a, b));
  T r = s_fma(-k, b, a);
  if (b > T(0)) {
    if (r < T(0)) r += b;
    if (r >= b) r -= b;
  } else {
    if (r > T(0)) r += b;
    if (r <= b) r -= b;
  }
  return r;
}

// ------------------------------------------------------------------------------------------------------------------
// Pack

template <int W>
struct Mask {
  bool m[W];
};

template <typename T, int W>
struct alignas(sizeof(T) * W >= 16 ? 16 : sizeof(T) * W) Pack {
  T v[W];
  typedef T scalar;
  static constexpr int width = W;
  AB_DEV Pack() {}
  AB_DEV Pack(T s) {
#pragma unroll
    for (int i = 0; i < W; i++) v[i] = s;
  }
};

#define AB_PACK_LOOP for (int i = 0; i < W; i++)

// --- add / sub / mul / fma: packed f32x2 for float, scalar otherwise. Always the explicitly rounded (_rn) forms: the
// compiler then never contracts a multiply of one op with an add of the next into an FMA, so the interpreter and the
// program-compiled straight-line kernels (codegen.py), which inline the same op bodies back to back, return identical
// bits. Every FMA the ops want is written as fma_().
template <int W>
AB_DEV Pack<float, W> p_add(const Pack<float, W>& a, const Pack<float, W>& b) {
  Pack<float, W> r;
#if AB_USE_F32X2
  if constexpr (W % 2 == 0) {
#pragma unroll
    for (int i = 0; i < W; i += 2) {
      float2 t = __fadd2_rn(make_float2(a.v[i], a.v[i + 1]), make_float2(b.v[i], b.v[i + 1]));
      r.v[i] = t.x;
      r.v[i + 1] = t.y;
    }
    return r;
  } else
#endif
  {
#pragma unroll
    AB_PACK_LOOP r.v[i] = __fadd_rn(a.v[i], b.v[i]);
    return r;
  }
}
template <int W>
AB_DEV Pack<float, W> p_mul(const Pack<float, W>& a, const Pack<float, W>& b) {
  Pack<float, W> r;
#if AB_USE_F32X2
  if constexpr (W % 2 == 0) {
#pragma unroll
    for (int i = 0; i < W; i += 2) {
      float2 t = __fmul2_rn(make_float2(a.v[i], a.v[i + 1]), make_float2(b.v[i], b.v[i + 1]));
      r.v[i] = t.x;
      r.v[i + 1] = t.y;
    }
    return r;
  } else
#endif
  {
#pragma unroll
    AB_PACK_LOOP r.v[i] = __fmul_rn(a.v[i], b.v[i]);
    return r;
  }
}
template <int W>
AB_DEV Pack<float, W> p_fma(const Pack<float, W>& a, const Pack<float, W>& b, const Pack<float, W>& c) {
  Pack<float, W> r;
#if AB_USE_F32X2
  if constexpr (W % 2 == 0) {
#pragma unroll
    for (int i = 0; i < W; i += 2) {
      float2 t = __ffma2_rn(make_float2(a.v[i], a.v[i + 1]), make_float2(b.v[i], b.v[i + 1]),
                            make_float2(c.v[i], c.v[i + 1]));
      r.v[i] = t.x;
      r.v[i + 1] = t.y;
    }
    return r;
  } else
#endif
  {
#pragma unroll
    AB_PACK_LOOP r.v[i] = fmaf(a.v[i], b.v[i], c.v[i]);
    return r;
  }
}
template <int W>
AB_DEV Pack<double, W> p_add(const Pack<double, W>& a, const Pack<double, W>& b) {
  Pack<double, W> r;
#pragma unroll
  AB_PACK_LOOP r.v[i] = __dadd_rn(a.v[i], b.v[i]);
  return r;
}
template <int W>
AB_DEV Pack<double, W> p_mul(const Pack<double, W>& a, const Pack<double, W>& b) {
  Pack<double, W> r;
#pragma unroll
  AB_PACK_LOOP r.v[i] = __dmul_rn(a.v[i], b.v[i]);
  return r;
}
template <int W>
AB_DEV Pack<double, W> p_fma(const Pack<double, W>& a, const Pack<double, W>& b, const Pack<double, W>& c) {
  Pack<double, W> r;
#pragma unroll
  AB_PACK_LOOP r.v[i] = fma(a.v[i], b.v[i], c.v[i]);
  return r;
}

template <typename T, int W>
AB_DEV Pack<T, W> operator+(const Pack<T, W>& a, const Pack<T, W>& b) { return p_add(a, b); }
template <typename T, int W>
AB_DEV Pack<T, W> operator*(const Pack<T, W>& a, const Pack<T, W>& b) { return p_mul(a, b); }
template <typename T, int W>
AB_DEV Pack<T, W> operator-(const Pack<T, W>& a) {
  Pack<T, W> r;
#pragma unroll
  AB_PACK_LOOP r.v[i] = -a.v[i];
  return r;
}
template <typename T, int W>
AB_DEV Pack<T, W> operator-(const Pack<T, W>& a, const Pack<T, W>& b) { return p_add(a, -b); }
template <typename T, int W>
AB_DEV Pack<T, W> operator+(const Pack<T, W>& a, T b) { return p_add(a, Pack<T, W>(b)); }
template <typename T, int W>
AB_DEV Pack<T, W> operator-(const Pack<T, W>& a, T b) { return p_add(a, Pack<T, W>(-b)); }
template <typename T, int W>
AB_DEV Pack<T, W> operator-(T a, const Pack<T, W>& b) { return p_add(Pack<T, W>(a), -b); }
template <typename T, int W>
AB_DEV Pack<T, W> operator*(const Pack<T, W>& a, T b) { return p_mul(a, Pack<T, W>(b)); }
template <typename T, int W>
AB_DEV Pack<T, W> operator*(T a, const Pack<T, W>& b) { return p_mul(Pack<T, W>(a), b); }
template <typename T, int W>
AB_DEV Pack<T, W> fma_(const Pack<T, W>& a, const Pack<T, W>& b, const Pack<T, W>& c) { return p_fma(a, b, c); }
template <typename T, int W>
AB_DEV Pack<T, W> fma_(const Pack<T, W>& a, T b, const Pack<T, W>& c) { return p_fma(a, Pack<T, W>(b), c); }
template <typename T, int W>
AB_DEV Pack<T, W> fma_(const Pack<T, W>& a, T b, T c) { return p_fma(a, Pack<T, W>(b), Pack<T, W>(c)); }

#define AB_PACK_UNARY(name, expr)                         \
  template <typename T, int W>                            \
  AB_DEV Pack<T, W> name(const Pack<T, W>& a) {           \
    Pack<T, W> r;                                         \
    _Pragma("unroll") AB_PACK_LOOP r.v[i] = expr(a.v[i]); \
    return r;                                             \
  }
AB_PACK_UNARY(abs_, s_abs)
AB_PACK_UNARY(sqrt_, s_sqrt)
AB_PACK_UNARY(floor_, s_floor)
AB_PACK_UNARY(exp_, s_exp)
AB_PACK_UNARY(rcp_, s_rcp)

#define AB_PACK_BINARY(name, expr)                                \
  template <typename T, int W>                                    \
  AB_DEV Pack<T, W> name(const Pack<T, W>& a, const Pack<T, W>& b) { \
    Pack<T, W> r;                                                 \
    _Pragma("unroll") AB_PACK_LOOP r.v[i] = expr(a.v[i], b.v[i]); \
    return r;                                                     \
  }                                                               \
  template <typename T, int W>                                    \
  AB_DEV Pack<T, W> name(const Pack<T, W>& a, T b) {              \
    Pack<T, W> r;                                                 \
    _Pragma("unroll") AB_PACK_LOOP r.v[i] = expr(a.v[i], b);      \
    return r;                                                     \
  }
AB_PACK_BINARY(min_, s_min)
AB_PACK_BINARY(max_, s_max)
AB_PACK_BINARY(div_, s_div)
AB_PACK_BINARY(mod_, s_mod)
AB_PACK_BINARY(pow_, s_pow)

// 1 / n for a Euclidean norm n >= 0 that multiplies quantities vanishing with n (components of the vector whose norm n is):
// the reciprocal of max(n, tiny) instead of a compare and a select around it (tiny = the square root of the smallest
// normal number: below it the squares that make up n underflow). At n = 0 the products are exactly 0, as before.
template <typename T, int W>
AB_DEV Pack<T, W> rcp_norm_(const Pack<T, W>& n) { return rcp_(max_(n, T(sizeof(T) == 4 ? 1e-18 : 1e-150))); }

// atan2 — fp64: libm per lane. fp32: atan(t) = t + t^3 Q(t^2) on t = min/max in [0,1] (Q of degree 8 fitted to 0.9 ulp,
// tools/fit_atan.py), quadrant fixed up afterwards; the Horner chain runs as packed FFMA2 (two points per issue slot).
// ~23 issue slots per pair of points instead of ~47 per point for atan2f; total error <= ~2 ulp like atan2f.
template <int W>
AB_DEV Pack<double, W> atan2_(const Pack<double, W>& y, const Pack<double, W>& x) {
  Pack<double, W> r;
#pragma unroll
  AB_PACK_LOOP r.v[i] = s_atan2(y.v[i], x.v[i]);
  return r;
}
template <int W>
AB_DEV Pack<float, W> atan2_(const Pack<float, W>& y, const Pack<float, W>& x) {
#if AB_FAST_ATAN2
  typedef Pack<float, W> P;
  const P ax = abs_(x), ay = abs_(y);
  const P mx = max_(ax, ay), mn = min_(ax, ay);
  P t;
#pragma unroll
  AB_PACK_LOOP t.v[i] = mx.v[i] > 0.0f ? mn.v[i] * s_rcp(mx.v[i]) : 0.0f;
  const P u = t * t;
  P q = fma_(u, -1.6024069807e-03f, 1.0025515875e-02f);
  q = fma_(q, u, P(-2.9446289233e-02f));
  q = fma_(q, u, P(5.6127113276e-02f));
  q = fma_(q, u, P(-8.2897455093e-02f));
  q = fma_(q, u, P(1.0910273375e-01f));
  q = fma_(q, u, P(-1.4255461991e-01f));
  q = fma_(q, u, P(1.9997619923e-01f));
  q = fma_(q, u, P(-3.3333262863e-01f));
  P r = fma_(t * u, q, t);
#pragma unroll
  AB_PACK_LOOP {
    float a = r.v[i];
    a = ay.v[i] > ax.v[i] ? 1.57079632679489661923f - a : a;
    a = x.v[i] < 0.0f ? 3.14159265358979323846f - a : a;
    r.v[i] = copysignf(a, y.v[i]);
  }
  return r;
#else
  Pack<float, W> r;
#pragma unroll
  AB_PACK_LOOP r.v[i] = s_atan2(y.v[i], x.v[i]);
  return r;
#endif
}

template <typename T, int W>
AB_DEV Pack<T, W> operator/(const Pack<T, W>& a, const Pack<T, W>& b) { return div_(a, b); }
template <typename T, int W>
AB_DEV Pack<T, W> operator/(const Pack<T, W>& a, T b) { return a * s_rcp(b); }

template <int W>
AB_DEV void sincos_(const Pack<double, W>& a, Pack<double, W>& s, Pack<double, W>& c) {
#pragma unroll
  AB_PACK_LOOP s_sincos(a.v[i], s.v[i], c.v[i]);
}
// sincos — fp32: q = rint(a 2/pi) by the 1.5 * 2^23 trick (its low mantissa bits are the quadrant), r = a - q pi/2 in three
// FMA steps (pi/2 split into three floats), sin r = r + r^3 S(r^2), cos r = 1 - r^2/2 + r^4 C(r^2) (degree-2 S and C, the
// Cephes single-precision sets), quadrant swap / signs per lane. All the arithmetic runs packed (two points per issue
// slot): ~17 packed + 9 per-lane instructions per pair instead of two ~35-instruction sincosf calls. Error <= 1.6 ulp for
// |a| <= 2000 (checked against fp64 on 8 M samples); larger arguments take sincosf.
template <int W>
AB_DEV void sincos_(const Pack<float, W>& a, Pack<float, W>& s, Pack<float, W>& c) {
#if AB_FAST_SINCOS
  typedef Pack<float, W> P;
  bool small = true;
#pragma unroll
  AB_PACK_LOOP small = small && (fabsf(a.v[i]) <= 2000.0f);
  if (small) {
    const P t = fma_(a, 0.636619772367581343f, P(12582912.0f));
    const P q = t - 12582912.0f;
    P r = fma_(q, -1.57079637050628662109375f, a);
    r = fma_(q, 4.37113900018624283e-8f, r);
    r = fma_(q, 1.71512449173013125e-15f, r);
    const P r2 = r * r;
    P sp = fma_(r2, -1.9515295891e-4f, P(8.3321608736e-3f));
    sp = fma_(sp, r2, P(-1.6666654611e-1f));
    sp = fma_(r * r2, sp, r);
    P cp = fma_(r2, 2.443315711809948e-5f, P(-1.388731625493765e-3f));
    cp = fma_(cp, r2, P(4.166664568298827e-2f));
    cp = fma_(r2 * r2, cp, fma_(r2, -0.5f, P(1.0f)));
#pragma unroll
    AB_PACK_LOOP {
      const unsigned n = __float_as_uint(t.v[i]);
      const bool swap = n & 1u;
      const float ss = swap ? cp.v[i] : sp.v[i], cc = swap ? sp.v[i] : cp.v[i];
      s.v[i] = __uint_as_float(__float_as_uint(ss) ^ ((n & 2u) << 30));
      c.v[i] = __uint_as_float(__float_as_uint(cc) ^ (((n + 1u) & 2u) << 30));
    }
    return;
  }
#endif
#pragma unroll
  AB_PACK_LOOP s_sincos(a.v[i], s.v[i], c.v[i]);
}

// comparisons -> Mask
#define AB_PACK_CMP(name, op)                                         \
  template <typename T, int W>                                        \
  AB_DEV Mask<W> name(const Pack<T, W>& a, const Pack<T, W>& b) {     \
    Mask<W> r;                                                        \
    _Pragma("unroll") AB_PACK_LOOP r.m[i] = a.v[i] op b.v[i];         \
    return r;                                                         \
  }                                                                   \
  template <typename T, int W>                                        \
  AB_DEV Mask<W> name(const Pack<T, W>& a, T b) {                     \
    Mask<W> r;                                                        \
    _Pragma("unroll") AB_PACK_LOOP r.m[i] = a.v[i] op b;              \
    return r;                                                         \
  }
AB_PACK_CMP(lt_, <)
AB_PACK_CMP(le_, <=)
AB_PACK_CMP(gt_, >)
AB_PACK_CMP(ge_, >=)
AB_PACK_CMP(eq_, ==)

template <int W>
AB_DEV Mask<W> operator&(const Mask<W>& a, const Mask<W>& b) {
  Mask<W> r;
#pragma unroll
  AB_PACK_LOOP r.m[i] = a.m[i] && b.m[i];
  return r;
}
template <int W>
AB_DEV Mask<W> operator|(const Mask<W>& a, const Mask<W>& b) {
  Mask<W> r;
#pragma unroll
  AB_PACK_LOOP r.m[i] = a.m[i] || b.m[i];
  return r;
}
template <int W>
AB_DEV Mask<W> operator!(const Mask<W>& a) {
  Mask<W> r;
#pragma unroll
  AB_PACK_LOOP r.m[i] = !a.m[i];
  return r;
}
template <int W>
AB_DEV bool any_(const Mask<W>& a) {
  bool r = false;
#pragma unroll
  AB_PACK_LOOP r = r || a.m[i];
  return r;
}
template <typename T, int W>
AB_DEV Pack<T, W> select_(const Mask<W>& m, const Pack<T, W>& a, const Pack<T, W>& b) {
  Pack<T, W> r;
#pragma unroll
  AB_PACK_LOOP r.v[i] = m.m[i] ? a.v[i] : b.v[i];
  return r;
}
// np.sign: -1, 0, +1
template <typename T, int W>
AB_DEV Pack<T, W> sign_(const Pack<T, W>& a) {
  Pack<T, W> r;
#pragma unroll
  AB_PACK_LOOP r.v[i] = (a.v[i] > T(0)) ? T(1) : ((a.v[i] < T(0)) ? T(-1) : T(0));
  return r;
}
template <typename T, int W>
AB_DEV Pack<T, W> value_of(const Pack<T, W>& a) { return a; }
template <typename T, int W>
AB_DEV Pack<T, W> constant_like(const Pack<T, W>&, T c) { return Pack<T, W>(c); }

// ------------------------------------------------------------------------------------------------------------------
// Dual

template <typename P, int K>
struct Dual {
  P v;
  P d[K];
  typedef typename P::scalar scalar;
  static constexpr int width = P::width;
  AB_DEV Dual() {}
  AB_DEV Dual(scalar s) : v(s) {
#pragma unroll
    for (int k = 0; k < K; k++) d[k] = P(scalar(0));
  }
  AB_DEV explicit Dual(const P& p) : v(p) {
#pragma unroll
    for (int k = 0; k < K; k++) d[k] = P(scalar(0));
  }
};

#define AB_DK for (int k = 0; k < K; k++)
#define AB_DUAL_T typename P::scalar

template <typename P, int K>
AB_DEV Dual<P, K> operator+(const Dual<P, K>& a, const Dual<P, K>& b) {
  Dual<P, K> r;
  r.v = a.v + b.v;
#pragma unroll
  AB_DK r.d[k] = a.d[k] + b.d[k];
  return r;
}
template <typename P, int K>
AB_DEV Dual<P, K> operator-(const Dual<P, K>& a, const Dual<P, K>& b) {
  Dual<P, K> r;
  r.v = a.v - b.v;
#pragma unroll
  AB_DK r.d[k] = a.d[k] - b.d[k];
  return r;
}
template <typename P, int K>
AB_DEV Dual<P, K> operator-(const Dual<P, K>& a) {
  Dual<P, K> r;
  r.v = -a.v;
#pragma unroll
  AB_DK r.d[k] = -a.d[k];
  return r;
}
template <typename P, int K>
AB_DEV Dual<P, K> operator*(const Dual<P, K>& a, const Dual<P, K>& b) {
  Dual<P, K> r;
  r.v = a.v * b.v;
#pragma unroll
  AB_DK r.d[k] = fma_(a.v, b.d[k], a.d[k] * b.v);
  return r;
}
template <typename P, int K>
AB_DEV Dual<P, K> operator+(const Dual<P, K>& a, AB_DUAL_T b) {
  Dual<P, K> r = a;
  r.v = a.v + b;
  return r;
}
template <typename P, int K>
AB_DEV Dual<P, K> operator-(const Dual<P, K>& a, AB_DUAL_T b) {
  Dual<P, K> r = a;
  r.v = a.v - b;
  return r;
}
template <typename P, int K>
AB_DEV Dual<P, K> operator-(AB_DUAL_T a, const Dual<P, K>& b) {
  Dual<P, K> r = -b;
  r.v = r.v + a;
  return r;
}
template <typename P, int K>
AB_DEV Dual<P, K> operator*(const Dual<P, K>& a, AB_DUAL_T b) {
  Dual<P, K> r;
  r.v = a.v * b;
#pragma unroll
  AB_DK r.d[k] = a.d[k] * b;
  return r;
}
template <typename P, int K>
AB_DEV Dual<P, K> operator*(AB_DUAL_T a, const Dual<P, K>& b) { return b * a; }
template <typename P, int K>
AB_DEV Dual<P, K> operator/(const Dual<P, K>& a, AB_DUAL_T b) { return a * s_rcp(b); }
template <typename P, int K>
AB_DEV Dual<P, K> fma_(const Dual<P, K>& a, const Dual<P, K>& b, const Dual<P, K>& c) { return a * b + c; }
template <typename P, int K>
AB_DEV Dual<P, K> fma_(const Dual<P, K>& a, AB_DUAL_T b, const Dual<P, K>& c) {
  Dual<P, K> r;
  r.v = fma_(a.v, b, c.v);
#pragma unroll
  AB_DK r.d[k] = fma_(a.d[k], b, c.d[k]);
  return r;
}
template <typename P, int K>
AB_DEV Dual<P, K> fma_(const Dual<P, K>& a, AB_DUAL_T b, AB_DUAL_T c) {
  Dual<P, K> r;
  r.v = fma_(a.v, b, c);
#pragma unroll
  AB_DK r.d[k] = a.d[k] * b;
  return r;
}
template <typename P, int K>
AB_DEV Dual<P, K> div_(const Dual<P, K>& a, const Dual<P, K>& b) {
  Dual<P, K> r;
  P ib = rcp_(b.v);
  r.v = a.v * ib;
#pragma unroll
  AB_DK r.d[k] = (a.d[k] - r.v * b.d[k]) * ib;
  return r;
}
template <typename P, int K>
AB_DEV Dual<P, K> div_(const Dual<P, K>& a, AB_DUAL_T b) { return a * s_rcp(b); }
template <typename P, int K>
AB_DEV Dual<P, K> operator/(const Dual<P, K>& a, const Dual<P, K>& b) { return div_(a, b); }

template <typename P, int K>
AB_DEV Dual<P, K> select_(const Mask<P::width>& m, const Dual<P, K>& a, const Dual<P, K>& b) {
  Dual<P, K> r;
  r.v = select_(m, a.v, b.v);
#pragma unroll
  AB_DK r.d[k] = select_(m, a.d[k], b.d[k]);
  return r;
}
template <typename P, int K>
AB_DEV Dual<P, K> min_(const Dual<P, K>& a, const Dual<P, K>& b) { return select_(le_(a.v, b.v), a, b); }
template <typename P, int K>
AB_DEV Dual<P, K> max_(const Dual<P, K>& a, const Dual<P, K>& b) { return select_(ge_(a.v, b.v), a, b); }
template <typename P, int K>
AB_DEV Dual<P, K> min_(const Dual<P, K>& a, AB_DUAL_T b) { return select_(le_(a.v, b), a, Dual<P, K>(b)); }
template <typename P, int K>
AB_DEV Dual<P, K> max_(const Dual<P, K>& a, AB_DUAL_T b) { return select_(ge_(a.v, b), a, Dual<P, K>(b)); }
template <typename P, int K>
AB_DEV Dual<P, K> abs_(const Dual<P, K>& a) { return select_(ge_(a.v, AB_DUAL_T(0)), a, -a); }
template <typename P, int K>
AB_DEV Dual<P, K> sqrt_(const Dual<P, K>& a) {
  Dual<P, K> r;
  r.v = sqrt_(a.v);
  // d sqrt = da / (2 sqrt); at the kink (sqrt == 0) the tangent is set to 0 instead of inf/NaN
  P h = select_(gt_(r.v, AB_DUAL_T(0)), rcp_(r.v + r.v), P(AB_DUAL_T(0)));
#pragma unroll
  AB_DK r.d[k] = a.d[k] * h;
  return r;
}
template <typename P, int K>
AB_DEV Dual<P, K> floor_(const Dual<P, K>& a) { return Dual<P, K>(floor_(a.v)); }
template <typename P, int K>
AB_DEV Dual<P, K> exp_(const Dual<P, K>& a) {
  Dual<P, K> r;
  r.v = exp_(a.v);
#pragma unroll
  AB_DK r.d[k] = a.d[k] * r.v;
  return r;
}
template <typename P, int K>
AB_DEV void sincos_(const Dual<P, K>& a, Dual<P, K>& s, Dual<P, K>& c) {
  sincos_(a.v, s.v, c.v);
#pragma unroll
  AB_DK {
    s.d[k] = a.d[k] * c.v;
    c.d[k] = -(a.d[k] * s.v);
  }
}
template <typename P, int K>
AB_DEV Dual<P, K> atan2_(const Dual<P, K>& y, const Dual<P, K>& x) {
  Dual<P, K> r;
  r.v = atan2_(y.v, x.v);
  P n = fma_(x.v, x.v, y.v * y.v);
  P in = select_(gt_(n, AB_DUAL_T(0)), rcp_(n), P(AB_DUAL_T(0)));
#pragma unroll
  AB_DK r.d[k] = (x.v * y.d[k] - y.v * x.d[k]) * in;
  return r;
}
template <typename P, int K>
AB_DEV Dual<P, K> mod_(const Dual<P, K>& a, AB_DUAL_T b) {
  Dual<P, K> r = a;
  r.v = mod_(a.v, b);
  return r;
}
// floor-mod by a dual divisor: a - b floor(a/b); the quotient is piecewise constant
template <typename P, int K>
AB_DEV Dual<P, K> mod_(const Dual<P, K>& a, const Dual<P, K>& b) {
  Dual<P, K> r;
  r.v = mod_(a.v, b.v);
  const P q = floor_(div_(a.v, b.v));
#pragma unroll
  AB_DK r.d[k] = a.d[k] - q * b.d[k];
  return r;
}
template <typename P, int K>
AB_DEV Dual<P, K> constant_like(const Dual<P, K>&, const Dual<P, K>& c) { return c; }
template <typename P, int K>
AB_DEV Dual<P, K> pow_(const Dual<P, K>& a, AB_DUAL_T b) {  // a >= 0
  Dual<P, K> r;
  r.v = pow_(a.v, b);
  P g = select_(gt_(a.v, AB_DUAL_T(0)), r.v * rcp_(a.v) * b, P(AB_DUAL_T(0)));
#pragma unroll
  AB_DK r.d[k] = a.d[k] * g;
  return r;
}
template <typename P, int K>
AB_DEV Dual<P, K> sign_(const Dual<P, K>& a) { return Dual<P, K>(sign_(a.v)); }

#define AB_DUAL_CMP(name)                                                                        \
  template <typename P, int K>                                                                   \
  AB_DEV Mask<P::width> name(const Dual<P, K>& a, const Dual<P, K>& b) { return name(a.v, b.v); } \
  template <typename P, int K>                                                                   \
  AB_DEV Mask<P::width> name(const Dual<P, K>& a, AB_DUAL_T b) { return name(a.v, b); }
AB_DUAL_CMP(lt_)
AB_DUAL_CMP(le_)
AB_DUAL_CMP(gt_)
AB_DUAL_CMP(ge_)

template <typename P, int K>
AB_DEV P value_of(const Dual<P, K>& a) { return a.v; }
template <typename P, int K>
AB_DEV Dual<P, K> constant_like(const Dual<P, K>&, AB_DUAL_T c) { return Dual<P, K>(c); }

// per-lane constants (a Pack without tangents) combined with S
template <typename T, int W>
AB_DEV Pack<T, W> mul_lane(const Pack<T, W>& a, const Pack<T, W>& c) { return a * c; }
template <typename T, int W>
AB_DEV Pack<T, W> add_lane(const Pack<T, W>& a, const Pack<T, W>& c) { return a + c; }
template <typename T, int W>
AB_DEV Pack<T, W> value_sign(const Pack<T, W>& a) { return sign_(a); }
template <typename P, int K>
AB_DEV Dual<P, K> mul_lane(const Dual<P, K>& a, const P& c) {
  Dual<P, K> r;
  r.v = a.v * c;
#pragma unroll
  AB_DK r.d[k] = a.d[k] * c;
  return r;
}
template <typename P, int K>
AB_DEV Dual<P, K> add_lane(const Dual<P, K>& a, const P& c) {
  Dual<P, K> r = a;
  r.v = a.v + c;
  return r;
}
template <typename P, int K>
AB_DEV P value_sign(const Dual<P, K>& a) { return sign_(a.v); }
// a * c + b with a per-lane constant c, as explicit FMAs (written as mul_lane + operator+ the adds stay separate
// instructions whenever the product has a second use)
template <typename T, int W>
AB_DEV Pack<T, W> fma_lane(const Pack<T, W>& a, const Pack<T, W>& c, const Pack<T, W>& b) { return fma_(a, c, b); }
template <typename P, int K>
AB_DEV Dual<P, K> fma_lane(const Dual<P, K>& a, const P& c, const Dual<P, K>& b) {
  Dual<P, K> r;
  r.v = fma_(a.v, c, b.v);
#pragma unroll
  AB_DK r.d[k] = fma_(a.d[k], c, b.d[k]);
  return r;
}

// ------------------------------------------------------------------------------------------------------------------
// shared helpers written once for both kinds of S

template <typename S, typename L, typename H>
AB_DEV S clamp_(const S& a, const L& lo, const H& hi) { return min_(max_(a, lo), hi); }
template <typename S>
AB_DEV S norm2_(const S& a, const S& b) { return sqrt_(fma_(a, a, b * b)); }
template <typename S>
AB_DEV S norm3_(const S& a, const S& b, const S& c) { return sqrt_(fma_(c, c, fma_(b, b, a * a))); }  // (z last, see op_affine)
// Euclidean norms of dual numbers in closed form: d|v| = (a da + b db [+ c dc]) / |v| (0 at the kink |v| = 0). Same value
// expression as the generic form; the tangents take 3 instructions per component instead of squaring duals and scaling
// by 1 / (2 sqrt) afterwards (about half the instructions of a sphere / torus / box tangent).
template <typename P, int K>
AB_DEV Dual<P, K> norm2_(const Dual<P, K>& a, const Dual<P, K>& b) {
  Dual<P, K> r;
  r.v = sqrt_(fma_(a.v, a.v, b.v * b.v));
  const P inv = rcp_norm_(r.v);
#pragma unroll
  AB_DK r.d[k] = fma_(a.v, a.d[k], b.v * b.d[k]) * inv;
  return r;
}
template <typename P, int K>
AB_DEV Dual<P, K> norm3_(const Dual<P, K>& a, const Dual<P, K>& b, const Dual<P, K>& c) {
  Dual<P, K> r;
  r.v = sqrt_(fma_(c.v, c.v, fma_(b.v, b.v, a.v * a.v)));
  const P inv = rcp_norm_(r.v);
#pragma unroll
  AB_DK r.d[k] = fma_(c.v, c.d[k], fma_(b.v, b.d[k], a.v * a.d[k])) * inv;
  return r;
}

}  // namespace ab
