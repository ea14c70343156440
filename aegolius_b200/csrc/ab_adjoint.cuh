// ab_adjoint.cuh — gradient pull-backs for the program-compiled field + gradient kernels (codegen.py, "adjoint" builds).
//
// The interpreter (and the first compiled kernels) carry three spatial tangents through EVERY op: a coordinate transform
// with Jacobian J costs J * (3 tangent vectors) = up to 27 multiply-adds per point on top of its value. A generated
// kernel is straight-line code, so the chain rule can be applied the other way round without any run-time tape:
//
//   * coordinate ops run on PLAIN points (exactly the value-only arithmetic) and leave what their Jacobian needs in
//     named registers (fwd_* below: a sine / cosine pair, a sector's rotation, an instance index, a few masks);
//   * a primitive is evaluated on a dual point seeded with the identity, i.e. it returns its value and its gradient in
//     its OWN local coordinates; value ops keep working on that dual number;
//   * when two values meet in a combine, and at the end of the program, the gradient 3-vector is pulled back through
//     the coordinate ops that separate its frame from the common one: g_in = J^T g_out (pb_* below), 9 multiply-adds
//     for a general affine map, 4 for a plane rotation, a select for mirrors / elongations, nothing for translations
//     and repetitions.
//
// The FIELD VALUE is computed by the same expressions as everywhere else (bit-identical to the interpreter under the same
// rules as the value kernels); the gradient is the same derivative evaluated in a different association order, so it
// differs from the interpreter's forward-mode gradient by rounding only (tests/test_gpu_jit.py states the bound).
// Reference semantics of every op: ab_ops.cuh (paths into Code/spomso/spomso/cores/ are cited there).
#pragma once
#include "ab_ops.cuh"

namespace ab {

// identity-seeded dual copy of a plain point: tangents = d / d(local x, y, z)
template <typename T, int W>
AB_DEV void seed_local(Pt<Dual<Pack<T, W>, 3>>& q, const Pt<Pack<T, W>>& p) { seed(q, p.x, p.y, p.z); }

// ---- maps whose Jacobian is constant --------------------------------------------------------------------------------
// op_affine: out_i = sum_j a[3 i + j] p_j + a[9 + i]  ->  g_j = sum_i a[3 i + j] g_i
template <typename P, typename T>
AB_DEV void pb_affine(Dual<P, 3>& v, const T* a) {
  const P gx = v.d[0], gy = v.d[1], gz = v.d[2];
  v.d[0] = fma_(gz, a[6], fma_(gy, a[3], gx * a[0]));
  v.d[1] = fma_(gz, a[7], fma_(gy, a[4], gx * a[1]));
  v.d[2] = fma_(gz, a[8], fma_(gy, a[5], gx * a[2]));
}
template <typename P, typename T>
AB_DEV void pb_scale(Dual<P, 3>& v, T s) {
#pragma unroll
  for (int k = 0; k < 3; k++) v.d[k] = v.d[k] * s;
}
template <typename P>
AB_DEV void pb_zero_z(Dual<P, 3>& v) { v.d[2] = P(typename P::scalar(0)); }

// ---- mirrors: |x| (ABSX_SUB, SYMMETRY). Dual abs_ flips the tangent where the value is negative.
template <int AXIS, typename P>
AB_DEV void pb_flip(Dual<P, 3>& v, const Mask<P::width>& neg) { v.d[AXIS] = select_(neg, -v.d[AXIS], v.d[AXIS]); }

// ---- elongation: the tangent vanishes inside [lo, hi] (elongate_axis)
template <typename P>
struct TapeElongate {
  Mask<P::width> in[3];
};
template <typename P, typename T>
AB_DEV void fwd_elongate(Pt<P>& p, const T* a, TapeElongate<P>& t) {
  op_elongate(p, a);
  // q - clamp(q, lo, hi) is exactly 0 inside [lo, hi] (the clamp returns q itself there): one compare per coordinate
  t.in[0] = eq_(p.x, T(0));
  t.in[1] = eq_(p.y, T(0));
  t.in[2] = eq_(p.z, T(0));
}
template <typename P>
AB_DEV void pb_elongate(Dual<P, 3>& v, const TapeElongate<P>& t) {
  typedef typename P::scalar T;
#pragma unroll
  for (int k = 0; k < 3; k++) v.d[k] = select_(t.in[k], P(T(0)), v.d[k]);
}

// ---- twist (op_twist): theta = k z, (x, y) -> (c x - s y, s x + c y)
//   d out / d (x, y) = [[c, -s], [s, c]],  d out / d z = k (-out_y, out_x)
template <typename P>
struct TapeTwist {
  P s, c, ox, oy;
};
template <typename P, typename T>
AB_DEV void fwd_twist(Pt<P>& p, const T* a, TapeTwist<P>& t) {
  sincos_(p.z * a[0], t.s, t.c);
  const P nx = t.c * p.x - t.s * p.y;
  const P ny = t.s * p.x + t.c * p.y;
  p.x = t.ox = nx;
  p.y = t.oy = ny;
}
template <typename P, typename T>
AB_DEV void pb_twist(Dual<P, 3>& v, const T* a, const TapeTwist<P>& t) {
  const P gx = v.d[0], gy = v.d[1];
  v.d[0] = fma_(t.s, gy, t.c * gx);
  v.d[1] = fma_(t.c, gy, -(t.s * gx));
  v.d[2] = fma_(fma_(t.ox, gy, -(t.oy * gx)), a[0], v.d[2]);
}

// ---- bend (op_bend; args r, angle/2, cos, sin, r*angle/2, r*sin, r*(1-cos), r*(angle/2))
//   curved part: out = (r atan2(x, -qy), |(x, qy)| - r), qy = y - r
//   straight part (|out_x| >= r angle/2): a rotation by -+ angle/2 about the end of the arc
template <typename P>
struct TapeBend {
  P x, qy, n;
  Mask<P::width> straight, pos;
  bool any_straight;
};
template <typename P, typename T>
AB_DEV void fwd_bend(Pt<P>& p, const T* a, TapeBend<P>& t) {
  const T r = a[0], c = a[2], s = a[3], thr = a[4], rs = a[5], r1c = a[6], rha = a[7];
  t.x = p.x;
  t.qy = p.y - r;
  const P phi = atan2_(p.x, -t.qy);
  t.n = norm2_(p.x, t.qy);
  P ny = t.n - r;
  P nx = phi * r;
  t.straight = ge_(abs_(nx), thr);
  t.pos = ge_(p.x, T(0));
  t.any_straight = any_(t.straight);
  if (t.any_straight) {
    const P sg = sign_(p.x);
    const P w0 = p.x - sg * rs;
    const P w1 = p.y - r1c;
    const P sw1 = w1 * s, sw0 = w0 * s;
    P wr0 = select_(t.pos, fma_(w0, c, sw1), fma_(w0, c, -sw1));
    const P wr1 = select_(t.pos, fma_(w1, c, -sw0), fma_(w1, c, sw0));
    wr0 = wr0 + sg * rha;
    nx = select_(t.straight, wr0, nx);
    ny = select_(t.straight, wr1, ny);
  }
  p.x = nx;
  p.y = ny;
}
template <typename P, typename T>
AB_DEV void pb_bend(Dual<P, 3>& v, const T* a, const TapeBend<P>& t) {
  const T r = a[0], c = a[2], s = a[3];
  const P gx = v.d[0], gy = v.d[1];
  const P inv = rcp_norm_(t.n);
  const P ux = t.x * inv, uy = t.qy * inv;  // unit vector from the bend centre: d out_y / d (x, y)
  const P k = (gx * r) * inv;               // d out_x / d (x, y) = r (-uy, ux) / n
  P ix = fma_(gy, ux, -(k * uy));
  P iy = fma_(gy, uy, k * ux);
  if (t.any_straight) {
    const P ss = select_(t.pos, P(s), P(-s));
    const P sx = fma_(gx, c, -(ss * gy));
    const P sy = fma_(gy, c, ss * gx);
    ix = select_(t.straight, sx, ix);
    iy = select_(t.straight, sy, iy);
  }
  v.d[0] = ix;
  v.d[1] = iy;
}

// ---- rotational symmetry (op_rotsym): out = (c x + s y - rad, -s x + c y) with the sector's (c, s)
template <typename P>
struct TapeRot {
  P c, s;
};
template <typename P, typename T>
AB_DEV void fwd_rotsym(Pt<P>& p, const T* a, TapeRot<P>& t) { op_rotsym(p, a, t.c, t.s); }
template <typename P>
AB_DEV void pb_rot(Dual<P, 3>& v, const TapeRot<P>& t) {
  const P gx = v.d[0], gy = v.d[1];
  v.d[0] = fma_(t.c, gx, -(t.s * gy));
  v.d[1] = fma_(t.s, gx, t.c * gy);
}

// ---- revolution (op_revolve): out = (|(x, z)| - R, y, 0)
template <typename P>
struct TapeRevolve {
  P ux, uz;
};
template <typename P, typename T>
AB_DEV void fwd_revolve(Pt<P>& p, const T* a, TapeRevolve<P>& t) {
  const P n = norm2_(p.x, p.z);
  const P inv = rcp_norm_(n);
  t.ux = p.x * inv;
  t.uz = p.z * inv;
  p.x = n - a[0];
  p.z = P(T(0));
}
template <typename P>
AB_DEV void pb_revolve(Dual<P, 3>& v, const TapeRevolve<P>& t) {
  const P gx = v.d[0];
  v.d[0] = t.ux * gx;
  v.d[2] = t.uz * gx;
}

// ---- instancing along a curve (op_curve_inst): out = R (p - o) for the nearest instance; the index is the tape, the
// frame rows are read again from the table in shared memory (three 128-bit loads per distinct index)
template <typename T, int W>
AB_DEV void fwd_curve_inst(Pt<Pack<T, W>>& p, const T* a, int mode, int (&idx)[W]) {
  curve_search(p, a, mode, idx);
  curve_frames(p, a + 4, idx, mode);
}
template <typename T, int W>
AB_DEV void pb_curve_inst(Dual<Pack<T, W>, 3>& v, const T* a, const int (&idx)[W]) {
  typedef Pack<T, W> P;
  if (same_index<W>(idx)) {  // one instance for the thread's points: broadcast record (curve_frames)
    P r[12];
    broadcast_record(a + 4, idx[0], 1, r);
    const P gx = v.d[0], gy = v.d[1], gz = v.d[2];
    v.d[0] = fma_(gz, r[9], fma_(gy, r[6], gx * r[3]));
    v.d[1] = fma_(gz, r[10], fma_(gy, r[7], gx * r[4]));
    v.d[2] = fma_(gz, r[11], fma_(gy, r[8], gx * r[5]));
    return;
  }
  constexpr int WC = (W % 2 == 0) ? 2 : 1;  // two points at a time, like curve_frames
#pragma unroll
  for (int c0 = 0; c0 < W; c0 += WC) {
    Pack<T, WC> r[12];
    gather_records(a + 4, idx + c0, 1, r);
    Pack<T, WC> g[3];
#pragma unroll
    for (int i = 0; i < WC; i++) {
      g[0].v[i] = v.d[0].v[c0 + i];
      g[1].v[i] = v.d[1].v[c0 + i];
      g[2].v[i] = v.d[2].v[c0 + i];
    }
    const Pack<T, WC> ox = fma_(g[2], r[9], fma_(g[1], r[6], g[0] * r[3]));
    const Pack<T, WC> oy = fma_(g[2], r[10], fma_(g[1], r[7], g[0] * r[4]));
    const Pack<T, WC> oz = fma_(g[2], r[11], fma_(g[1], r[8], g[0] * r[5]));
#pragma unroll
    for (int i = 0; i < WC; i++) {
      v.d[0].v[c0 + i] = ox.v[i];
      v.d[1].v[c0 + i] = oy.v[i];
      v.d[2].v[c0 + i] = oz.v[i];
    }
  }
}

// ---- axis revolution (op_axis_revolve; args radius, cos, sin): (xr, yr) = R(-a) (x, y), m = |(xr, z)|,
// out = (c m - s yr - rad, s m + c yr, 0)
template <typename P, typename T>
AB_DEV void fwd_axis_revolve(Pt<P>& p, const T* a, TapeRevolve<P>& t) {
  const T c = a[1], s = a[2];
  const P xr = fma_(p.x, c, p.y * s);
  const P yr = fma_(p.y, c, -(p.x * s));
  const P m = norm2_(xr, p.z);
  const P inv = rcp_norm_(m);
  t.ux = xr * inv;
  t.uz = p.z * inv;
  p.x = fma_(m, c, -(yr * s)) - a[0];
  p.y = fma_(m, s, yr * c);
  p.z = P(T(0));
}
template <typename P, typename T>
AB_DEV void pb_axis_revolve(Dual<P, 3>& v, const T* a, const TapeRevolve<P>& t) {
  const T c = a[1], s = a[2];
  const P gx = v.d[0], gy = v.d[1];
  const P gm = fma_(gy, s, gx * c);      // d / d m
  const P gyr = fma_(gy, c, -(gx * s));  // d / d yr
  const P gxr = t.ux * gm;
  v.d[0] = fma_(gxr, c, -(gyr * s));
  v.d[1] = fma_(gxr, s, gyr * c);
  v.d[2] = t.uz * gm;
}

// ---- primitives with their local gradient written out (the identity-seeded dual evaluation multiplies by the zeros and
// ones of the seed: a sphere's tangents cost 12 instructions that way, 3 here). Same value expressions as ab_ops.cuh; the
// selects follow the dual min_ / max_ / abs_ (ties go the same way).
template <typename P, typename T>
AB_DEV Dual<P, 3> grad_sphere(const Pt<P>& p, const T* a) {
  Dual<P, 3> r;
  const P n = norm3_(p.x, p.y, p.z);
  const P inv = rcp_norm_(n);
  r.v = n - a[0];
  r.d[0] = p.x * inv;
  r.d[1] = p.y * inv;
  r.d[2] = p.z * inv;
  return r;
}
template <typename P, typename T>
AB_DEV Dual<P, 3> grad_torus(const Pt<P>& p, const T* a) {
  Dual<P, 3> r;
  const P m = norm2_(p.x, p.y);
  const P q = m - a[0];
  const P n = norm2_(q, p.z);
  const P im = rcp_norm_(m);
  const P in = rcp_norm_(n);
  const P gq = (q * in) * im;
  r.v = n - a[1];
  r.d[0] = gq * p.x;
  r.d[1] = gq * p.y;
  r.d[2] = p.z * in;
  return r;
}
template <typename P, typename T>
AB_DEV Dual<P, 3> grad_box(const Pt<P>& p, const T* a) {
  constexpr int W = P::width;
  Dual<P, 3> r;
  const P q0 = abs_(p.x) - a[0], q1 = abs_(p.y) - a[1], q2 = abs_(p.z) - a[2];
  const P o0 = max_(q0, T(0)), o1 = max_(q1, T(0)), o2 = max_(q2, T(0));
  const P n = norm3_(o0, o1, o2);
  const P m12 = max_(q1, q2);
  const P mx = max_(q0, m12);
  r.v = n + min_(mx, T(0));
  const P inv = rcp_norm_(n);
  const Mask<W> inner = le_(mx, T(0));        // min_(mx, 0) passes mx's tangent
  const Mask<W> first = ge_(q0, m12);         // max_(q0, m12) picks q0
  const Mask<W> second = ge_(q1, q2);         // max_(q1, q2) picks q1
  const P one(T(1)), zero(T(0));
  P g0 = fma_(o0, inv, select_(inner & first, one, zero));
  P g1 = fma_(o1, inv, select_(inner & !first & second, one, zero));
  P g2 = fma_(o2, inv, select_(inner & !first & !second, one, zero));
  r.d[0] = select_(ge_(p.x, T(0)), g0, -g0);
  r.d[1] = select_(ge_(p.y, T(0)), g1, -g1);
  r.d[2] = select_(ge_(p.z, T(0)), g2, -g2);
  return r;
}
template <typename P, typename T>
AB_DEV Dual<P, 3> grad_cylinder(const Pt<P>& p, const T* a) {
  constexpr int W = P::width;
  Dual<P, 3> r;
  const P m = norm2_(p.x, p.y);
  const P d0 = m - a[0];
  const P d1 = abs_(p.z) - a[1];
  const P o0 = max_(d0, T(0)), o1 = max_(d1, T(0));
  const P n = norm2_(o0, o1);
  const P mx = max_(d0, d1);
  r.v = min_(mx, T(0)) + n;
  const P inv = rcp_norm_(n);
  const P im = rcp_norm_(m);
  const Mask<W> inner = le_(mx, T(0));
  const Mask<W> first = ge_(d0, d1);
  const P one(T(1)), zero(T(0));
  const P g0 = fma_(o0, inv, select_(inner & first, one, zero)) * im;
  const P g1 = fma_(o1, inv, select_(inner & !first, one, zero));
  r.d[0] = g0 * p.x;
  r.d[1] = g0 * p.y;
  r.d[2] = select_(ge_(p.z, T(0)), g1, -g1);
  return r;
}

}  // namespace ab
