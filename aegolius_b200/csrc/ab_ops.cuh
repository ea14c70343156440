// ab_ops.cuh — the interpreter's operations, written once over S (Pack or Dual<Pack>) and the scalar type T of the
// argument pool. Each function restates the reference semantics cited next to it (paths relative to
// Code/spomso/spomso/cores/); host-side constant folding is described in aegolius_b200/program.py.
#pragma once
#include "ab_math.cuh"

namespace ab {

template <typename S>
struct Pt {
  S x, y, z;
};

// ---- argument access ------------------------------------------------------------------------------------------------
// Ops read their arguments through `A a`: either a plain `const T*` (value / spatial-gradient kernels: a[i] is a scalar)
// or ArgD<S> (parameter-tangent kernel, AB_GRAD_PARAM: a[i] is a dual number whose tangent is d arg_i / d theta, so the
// op bodies below propagate d/d theta through every arithmetic use of a parameter without being rewritten).
template <typename S>
struct ArgD {
  typedef typename S::scalar T;
  const T* v;
  const T* d;
  AB_DEV S operator[](int i) const {
    S r(v[i]);
    r.d[0] = decltype(r.v)(d[i]);
    return r;
  }
  AB_DEV ArgD operator+(int k) const { return ArgD{v + k, d + k}; }
};
template <typename T>
AB_DEV const T* raw_args(const T* a) { return a; }  // tables (instances, polylines, vertices) carry no tangents
template <typename S>
AB_DEV const typename S::scalar* raw_args(const ArgD<S>& a) { return a.v; }
// uniform value of an argument (for counts, flags and branch selection)
AB_DEV float aval(float x) { return x; }
AB_DEV double aval(double x) { return x; }
template <typename P, int K>
AB_DEV typename P::scalar aval(const Dual<P, K>& x) { return x.v.v[0]; }
// reciprocal of an argument
AB_DEV float rcp_arg(float x) { return s_rcp(x); }
AB_DEV double rcp_arg(double x) { return s_rcp(x); }
template <typename P, int K>
AB_DEV Dual<P, K> rcp_arg(const Dual<P, K>& x) { return div_(Dual<P, K>(typename P::scalar(1)), x); }

// ---- coordinate ops -------------------------------------------------------------------------------------------------

// do the W points of a thread share one table index (instance, sector)? Then the entry is read once and enters the packed
// arithmetic as a broadcast operand.
template <int W>
AB_DEV bool same_index(const int* idx) {
  bool same = true;
#pragma unroll
  for (int i = 1; i < W; i++) same = same && (idx[i] == idx[0]);
  return same;
}

// p = M p + b : folded apply_ec_transforms (transformations.py:232-242), shears, frames
template <typename S, typename A>
AB_DEV void op_affine(Pt<S>& p, A a) {
  typedef typename S::scalar T;
  // z (the fastest grid axis, the one that varies along a thread's run of points) enters last: on a grid the x / y part of
  // the first transform of a tree is then the same for all W points of a thread and is evaluated once (codegen.py, rowsplit)
  S nx = fma_(p.z, a[2], fma_(p.y, a[1], fma_(p.x, a[0], a[9])));
  S ny = fma_(p.z, a[5], fma_(p.y, a[4], fma_(p.x, a[3], a[10])));
  S nz = fma_(p.z, a[8], fma_(p.y, a[7], fma_(p.x, a[6], a[11])));
  p.x = nx;
  p.y = ny;
  p.z = nz;
}
template <typename S, typename A>
AB_DEV void op_translate(Pt<S>& p, A a) {
  typedef typename S::scalar T;
  p.x = p.x + a[0];
  p.y = p.y + a[1];
  p.z = p.z + a[2];
}
template <typename S, typename A>
AB_DEV void op_scale_p(Pt<S>& p, A a) {
  typedef typename S::scalar T;
  p.x = p.x * a[0];
  p.y = p.y * a[0];
  p.z = p.z * a[0];
}
// modifications.py:91-93  q = p - clip(p, -e/2, e/2)
template <typename S, typename U>
AB_DEV S elongate_axis(const S& q, const U& lo, const U& hi) { return q - clamp_(q, lo, hi); }
// dual coordinate, plain bounds: the tangent passes through outside [lo, hi] and vanishes inside (one select per
// component instead of two dual min / max and a dual subtraction)
template <typename P, int K>
AB_DEV Dual<P, K> elongate_axis(const Dual<P, K>& q, const typename P::scalar& lo, const typename P::scalar& hi) {
  typedef typename P::scalar T;
  Dual<P, K> r;
  r.v = q.v - clamp_(q.v, lo, hi);
  const Mask<P::width> inside = ge_(q.v, lo) & le_(q.v, hi);
#pragma unroll
  for (int k = 0; k < K; k++) r.d[k] = select_(inside, P(T(0)), q.d[k]);
  return r;
}
template <typename S, typename A>
AB_DEV void op_elongate(Pt<S>& p, A a) {
  typedef typename S::scalar T;
  p.x = elongate_axis(p.x, a[0], a[3]);
  p.y = elongate_axis(p.y, a[1], a[4]);
  p.z = elongate_axis(p.z, a[2], a[5]);
}
// modifications.py:516-522
template <typename S, typename A>
AB_DEV void op_twist(Pt<S>& p, A a) {
  typedef typename S::scalar T;
  S s, c;
  sincos_(p.z * a[0], s, c);
  S nx = c * p.x - s * p.y;
  S ny = s * p.x + c * p.y;
  p.x = nx;
  p.y = ny;
}
// modifications.py:545-573; args r, angle/2, cos, sin, r*angle/2, r*sin, r*(1-cos), r*(angle/2)
template <typename S, typename A>
AB_DEV void op_bend(Pt<S>& p, A a) {
  typedef typename S::scalar T;
  const auto r = a[0], c = a[2], s = a[3], thr = a[4], rs = a[5], r1c = a[6], rha = a[7];
  S qy = p.y - r;
  S phi = atan2_(p.x, -qy);
  S ny = norm2_(p.x, qy) - r;
  S nx = phi * r;
  Mask<S::width> straight = ge_(abs_(nx), thr);
  if (any_(straight)) {
    S sg = sign_(p.x);
    S w0 = p.x - sg * rs;
    S w1 = p.y - r1c;
    Mask<S::width> pos = ge_(p.x, T(0));
    S sw1 = w1 * s, sw0 = w0 * s;
    S wr0 = select_(pos, fma_(w0, c, sw1), fma_(w0, c, -sw1));
    S wr1 = select_(pos, fma_(w1, c, -sw0), fma_(w1, c, sw0));
    wr0 = wr0 + sg * rha;
    nx = select_(straight, wr0, nx);
    ny = select_(straight, wr1, ny);
  }
  p.x = nx;
  p.y = ny;
}
// modifications.py:991-993
template <typename S, typename A>
AB_DEV void op_absx_sub(Pt<S>& p, A a) { typedef typename S::scalar T; p.x = abs_(p.x) - a[0]; }
// modifications.py:1023-1029 (pre-rotation emitted as AFFINE); args angle, radius, n_sectors, pad, (cos, sin)[n_sectors].
// The reference maps phi -> mod(phi, angle) - angle/2 and rebuilds (r cos, r sin); that is a rotation by
// -theta_k, theta_k = k*angle + angle/2, k = floor(phi/angle): the sector comes from atan2, the rotation from the
// host-computed table (fewer roundings than the polar round trip, no sqrt / sincos / mod).
template <typename S, typename A>
AB_DEV void op_rotsym(Pt<S>& p, A a, Pack<typename S::scalar, S::width>& c, Pack<typename S::scalar, S::width>& s) {
  typedef typename S::scalar T;
  constexpr int W = S::width;
  const T* tab = raw_args(a);  // angle and sector table are structural: no parameter tangent flows through them
  const T ang = tab[0];
  const auto rad = a[1];
  const int nsec = (int)tab[2];
  const T inv = s_rcp(ang);
  auto vx = value_of(p.x), vy = value_of(p.y);
  const Pack<T, W> ph = atan2_(vy, vx);
  int ks[W];
#pragma unroll
  for (int i = 0; i < W; i++) {
    T phi = ph.v[i];
    int k;
    if constexpr (sizeof(T) == 4) {
      // fp32: the angle itself carries ~1e-7 of rounding, so a sector decision that close to a boundary is arbitrary either
      // way (the parity tests mask it through the oracle's margin): no exact re-check of the quotient
      T t = phi * inv;
      if (phi < T(0)) t += (T)nsec;
      k = (int)t;
      k = k < 0 ? 0 : (k >= nsec ? nsec - 1 : k);
    } else {
      if (phi < T(0)) phi += T(6.283185307179586476925286766559);
      k = (int)(phi * inv);
      k = k < 0 ? 0 : (k >= nsec ? nsec - 1 : k);
      // one more exact step: phi*inv may round across an integer
      if (phi < (T)k * ang && k > 0) k--;
      else if (phi >= (T)(k + 1) * ang && k + 1 < nsec) k++;
    }
    ks[i] = k;
  }
  if (same_index<W>(ks)) {  // the thread's points share the sector: one table read, broadcast operands
    c = Pack<T, W>(tab[4 + 2 * ks[0]]);
    s = Pack<T, W>(tab[5 + 2 * ks[0]]);
  } else {
#pragma unroll
    for (int i = 0; i < W; i++) {
      c.v[i] = tab[4 + 2 * ks[i]];
      s.v[i] = tab[5 + 2 * ks[i]];
    }
  }
  S nx = fma_lane(p.y, s, mul_lane(p.x, c)) - rad;
  S ny = fma_lane(p.x, -s, mul_lane(p.y, c));
  p.x = nx;
  p.y = ny;
}
template <typename S, typename A>
AB_DEV void op_rotsym(Pt<S>& p, A a) {
  Pack<typename S::scalar, S::width> c, s;
  op_rotsym(p, a, c, s);
}
// modifications.py:427-431
template <typename S, typename A>
AB_DEV void op_revolve(Pt<S>& p, A a) {
  typedef typename S::scalar T;
  p.x = norm2_(p.x, p.z) - a[0];
  p.z = constant_like(p.z, T(0));
}
// modifications.py:455-467; args radius, cos, sin
template <typename S, typename A>
AB_DEV void op_axis_revolve(Pt<S>& p, A a) {
  typedef typename S::scalar T;
  const auto rad = a[0], c = a[1], s = a[2];
  S xr = fma_(p.x, c, p.y * s);
  S yr = fma_(p.y, c, -(p.x * s));
  S m = norm2_(xr, p.z);
  p.x = fma_(m, c, -(yr * s)) - rad;
  p.y = fma_(m, s, yr * c);
  p.z = constant_like(p.z, T(0));
}
// modifications.py:819-820; args d(3), d/2(3)
template <typename S, typename A>
AB_DEV void op_rep_inf(Pt<S>& p, A a) {
  typedef typename S::scalar T;
  p.x = mod_(p.x + a[3], a[0]) - a[3];
  p.y = mod_(p.y + a[4], a[1]) - a[4];
  p.z = mod_(p.z + a[5], a[2]) - a[5];
}
// modifications.py:847-868 per axis; args c(3), d(3), s(3), s/2(3)
template <typename S, typename U>
AB_DEV S rep_fin_axis(const S& q, const U& c, const U& d, const U& s, const U& sh) {
  typedef typename S::scalar T;
  Mask<S::width> inner = ge_(q, -d) & le_(q, d);
  S v = abs_(q) - c;
  v = select_(lt_(q, T(0)), -v, v);
  S u = mod_(q - d, s) - sh;
  return select_(inner, u, v);
}
template <typename S, typename A>
AB_DEV void op_rep_fin(Pt<S>& p, A a) {
  typedef typename S::scalar T;
  p.x = rep_fin_axis(p.x, a[0], a[3], a[6], a[9]);
  p.y = rep_fin_axis(p.y, a[1], a[4], a[7], a[10]);
  p.z = rep_fin_axis(p.z, a[2], a[5], a[8], a[11]);
}
// modifications.py:1069-1083 (frame emitted as AFFINE); args l/2, s, d, lo, hi, off ; inner = (n > 2)
template <typename S, typename A>
AB_DEV void op_lin_inst(Pt<S>& p, A a, int inner_on) {
  typedef typename S::scalar T;
  S v = abs_(p.x) - a[0];
  v = select_(lt_(p.x, T(0)), -v, v);
  if (inner_on) {
    S u = mod_(p.x - a[5], a[1]) - a[2];
    v = select_(ge_(p.x, a[3]) & le_(p.x, a[4]), u, v);
  }
  p.x = v;
}
// min / max over the warp of a NON-NEGATIVE value: the bit patterns of non-negative floats order like unsigned integers, so
// fp32 takes one REDUX instruction instead of a five-round shuffle butterfly
AB_DEV float warp_min_nonneg(float v) { return __uint_as_float(__reduce_min_sync(0xffffffffu, __float_as_uint(v))); }
AB_DEV float warp_max_nonneg(float v) { return __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(v))); }
AB_DEV double warp_min_nonneg(double v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, off));
  return v;
}
AB_DEV double warp_max_nonneg(double v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, off));
  return v;
}
// the first four scalars of a 16-byte aligned record with one 128-bit load
AB_DEV void load4(const float* q, float& a, float& b, float& c, float& d) {
  const float4 v = *reinterpret_cast<const float4*>(q);
  a = v.x; b = v.y; c = v.z; d = v.w;
}
AB_DEV void load4(const double* q, double& a, double& b, double& c, double& d) {
  const double2 u = *reinterpret_cast<const double2*>(q), v = *reinterpret_cast<const double2*>(q + 2);
  a = u.x; b = u.y; c = v.x; d = v.y;
}

// second half of curve instancing: subtract the chosen instance's position and, for the aligned variants (mode 1), rotate
// into its frame (rows dx, dy, dz). The 12-scalar records come in as three 128-bit loads; consecutive points mostly share
// one record. Plain packs are processed two points at a time so that only 24 record registers are live (8 points per
// thread would otherwise hold 96), duals (at most 2 points per thread) in one go.
template <typename T, int WC>
AB_DEV void gather_records(const T* rec, const int* idx, int mode, Pack<T, WC> (&r)[12]) {
  if (mode) {
#pragma unroll
    for (int i = 0; i < WC; i++) {
      if (i > 0 && idx[i] == idx[i - 1]) {
#pragma unroll
        for (int k = 0; k < 12; k++) r[k].v[i] = r[k].v[i - 1];
      } else {
        const T* q = rec + idx[i] * 12;
        load4(q, r[0].v[i], r[1].v[i], r[2].v[i], r[3].v[i]);
        load4(q + 4, r[4].v[i], r[5].v[i], r[6].v[i], r[7].v[i]);
        load4(q + 8, r[8].v[i], r[9].v[i], r[10].v[i], r[11].v[i]);
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < WC; i++) {
      const T* q = rec + idx[i] * 3;
      r[0].v[i] = q[0];
      r[1].v[i] = q[1];
      r[2].v[i] = q[2];
    }
  }
}
template <typename S, typename PK>
AB_DEV void frame_change(Pt<S>& p, const PK (&r)[12], int mode) {
  S dx = add_lane(p.x, -r[0]), dy = add_lane(p.y, -r[1]), dz = add_lane(p.z, -r[2]);
  if (mode) {
    p.x = fma_lane(dz, r[5], fma_lane(dy, r[4], mul_lane(dx, r[3])));
    p.y = fma_lane(dz, r[8], fma_lane(dy, r[7], mul_lane(dx, r[6])));
    p.z = fma_lane(dz, r[11], fma_lane(dy, r[10], mul_lane(dx, r[9])));
  } else {
    p.x = dx;
    p.y = dy;
    p.z = dz;
  }
}
// all W points of a thread in one instance (the rule: a thread's points are neighbours): the record is loaded once as
// scalars and enters the packed arithmetic as a broadcast operand (no per-lane register shuffling)
template <typename T, int W>
AB_DEV void broadcast_record(const T* rec, int j, int mode, Pack<T, W> (&r)[12]) {
  T q[12];
  if (mode) {
    load4(rec + j * 12, q[0], q[1], q[2], q[3]);
    load4(rec + j * 12 + 4, q[4], q[5], q[6], q[7]);
    load4(rec + j * 12 + 8, q[8], q[9], q[10], q[11]);
  } else {
#pragma unroll
    for (int k = 0; k < 3; k++) q[k] = rec[j * 3 + k];
#pragma unroll
    for (int k = 3; k < 12; k++) q[k] = T(0);
  }
#pragma unroll
  for (int k = 0; k < 12; k++) r[k] = Pack<T, W>(q[k]);
}
template <typename P, int K>
AB_DEV void curve_frames(Pt<Dual<P, K>>& p, const typename P::scalar* rec, const int* idx, int mode) {
  P r[12];
  if (same_index<P::width>(idx)) broadcast_record(rec, idx[0], mode, r);
  else gather_records(rec, idx, mode, r);
  frame_change(p, r, mode);
}
template <typename T, int W>
AB_DEV void curve_frames(Pt<Pack<T, W>>& p, const T* rec, const int* idx, int mode) {
  if (same_index<W>(idx)) {
    Pack<T, W> r[12];
    broadcast_record(rec, idx[0], mode, r);
    frame_change(p, r, mode);
    return;
  }
  constexpr int WC = (W % 2 == 0) ? 2 : 1;
#pragma unroll
  for (int c = 0; c < W; c += WC) {
    Pack<T, WC> r[12];
    gather_records(rec, idx + c, mode, r);
    Pt<Pack<T, WC>> q;
#pragma unroll
    for (int i = 0; i < WC; i++) {
      q.x.v[i] = p.x.v[c + i];
      q.y.v[i] = p.y.v[c + i];
      q.z.v[i] = p.z.v[c + i];
    }
    frame_change(q, r, mode);
#pragma unroll
    for (int i = 0; i < WC; i++) {
      p.x.v[c + i] = q.x.v[i];
      p.y.v[c + i] = q.y.v[i];
      p.z.v[c + i] = q.z.v[i];
    }
  }
}

// modifications.py:1120-1127 / 1183-1191 / 1253-1261: nearest instance by position (the reference asks a KD-tree), then
// subtract its position and, for the aligned variants, rotate into its frame. Record = pos(3) [+ rows dx,dy,dz (9)].
//
// Warp-cooperative exact search. A warp's 32*W points are spatially close (a tile of grid points), so the few instances
// that can be nearest for ANY of them are found once per warp: lane l measures instance l (32 per round) from a warp
// reference point C; i* = the instance nearest to C, R = max distance of the warp's points from C. Instance j can only
// win at a point of the ball B(C, R) if the bisector plane of (i*, j) cuts the ball:
//     (d(C, j)^2 - d(C, i*)^2) / (2 |q_j - q_i*|) <= R,   tested squared (no square root),
// which is exact and much tighter than the triangle inequality d(C, j) <= d(C, i*) + 2R far from the curve, where many
// instances are almost equidistant (that test kept 2-3 candidates for most warps of the headline grid). Every thread then
// scans just the candidate list for its W points. Cost ~ n/32 + (#candidates) per point instead of n; scattered point
// sets degrade gracefully to the full scan.
template <typename S, typename A>
AB_DEV void curve_search(const Pt<S>& p, A a, int mode, int (&idx)[S::width]) {
  typedef typename S::scalar T;
  const T* tab = raw_args(a);  // instance table: structural (no parameter tangents)
  const int n = (int)tab[0];
  const int stride = mode ? 12 : 3;
  const T* rec = tab + 4;  // records start on a 16-byte boundary (and stay on one for mode 1: 12 scalars each)
  constexpr int W = S::width;
  constexpr unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  auto vx = value_of(p.x), vy = value_of(p.y), vz = value_of(p.z);
  // reference point: the first point of the middle lane (halves the warp radius against taking lane 0's)
  const T Cx = __shfl_sync(FULL, vx.v[0], 16), Cy = __shfl_sync(FULL, vy.v[0], 16), Cz = __shfl_sync(FULL, vz.v[0], 16);
  T r2 = T(0);
#pragma unroll
  for (int i = 0; i < W; i++) {
    const T dx = vx.v[i] - Cx, dy = vy.v[i] - Cy, dz = vz.v[i] - Cz;
    r2 = s_max(r2, s_fma(dx, dx, s_fma(dy, dy, dz * dz)));
  }
  // lane l measures instance base + l from the reference point; the first round's distance is kept for the second pass
  auto lane_d2 = [&](int j) {
    const T dx = Cx - rec[j * stride], dy = Cy - rec[j * stride + 1], dz = Cz - rec[j * stride + 2];
    return s_fma(dx, dx, s_fma(dy, dy, dz * dz));
  };
  // lane l measures instances l, l + 32, ... from the reference point and keeps its nearest
  const T d_first = lane < n ? lane_d2(lane) : T(3.0e38);
  T dl = d_first;
  int jl = lane;
  for (int base = 32; base < n; base += 32)
    if (base + lane < n) {
      const T d = lane_d2(base + lane);
      if (d < dl) {
        dl = d;
        jl = base + lane;
      }
    }
  r2 = warp_max_nonneg(r2);
  const T dmin = warp_min_nonneg(dl);
  const unsigned who = __ballot_sync(FULL, dl == dmin);
  const int istar = __shfl_sync(FULL, jl, who ? __ffs(who) - 1 : 0);  // (who == 0 only for NaN coordinates)
  const T sx = rec[istar * stride], sy = rec[istar * stride + 1], sz = rec[istar * stride + 2];
  // widened a little against rounding (of the squared distances and of R): it only admits extra candidates
  const T lo = dmin * T(1.000002), r2w = T(4.002) * r2;
  auto candidate = [&](int j, T dj) {
    const T ex = rec[j * stride] - sx, ey = rec[j * stride + 1] - sy, ez = rec[j * stride + 2] - sz;
    const T e2 = s_fma(ex, ex, s_fma(ey, ey, ez * ez));
    const T diff = s_fma(dj, T(0.999998), -lo);
    return diff <= T(0) || diff * diff <= s_fma(r2w, e2, T(1e-30));
  };
  unsigned m = __ballot_sync(FULL, lane < n && candidate(lane, d_first));
  if (n <= 32 && (m & (m - 1)) == 0) {
    // one candidate for the whole warp (the common case away from the cell boundaries of the instances): it is the
    // nearest instance of every point, no distance needs to be evaluated
    const int j = m ? __ffs(m) - 1 : 0;  // (m == 0 only for NaN coordinates)
#pragma unroll
    for (int i = 0; i < W; i++) idx[i] = j;
  } else {
    T bst[W];
#pragma unroll
    for (int i = 0; i < W; i++) {
      bst[i] = T(3.0e38);
      idx[i] = 0;
    }
    for (int base = 0; base < n; base += 32) {
      if (base) m = __ballot_sync(FULL, base + lane < n && candidate(base + lane, lane_d2(base + lane)));
      while (m) {  // warp-uniform loop over the candidates, in index order (ties resolve to the lowest index)
        const int j = base + __ffs(m) - 1;
        m &= m - 1;
        const T qx = rec[j * stride], qy = rec[j * stride + 1], qz = rec[j * stride + 2];
        const Pack<T, W> dx = vx - qx, dy = vy - qy, dz = vz - qz;  // packed f32x2 lanes
        const Pack<T, W> d2 = fma_(dx, dx, fma_(dy, dy, dz * dz));
#pragma unroll
        for (int i = 0; i < W; i++) {
          if (d2.v[i] < bst[i]) {
            bst[i] = d2.v[i];
            idx[i] = j;
          }
        }
      }
    }
  }
}
template <typename S, typename A>
AB_DEV void op_curve_inst(Pt<S>& p, A a, int mode) {
  int idx[S::width];
  curve_search(p, a, mode, idx);
  curve_frames(p, raw_args(a) + 4, idx, mode);
}

// ---- value ops --------------------------------------------------------------------------------------------------------

// modifications.py:489-497
template <typename S, typename T>
AB_DEV S op_extrude_end(const S& d, const S& w1) {
  S o0 = max_(d, T(0)), o1 = max_(w1, T(0));
  return min_(max_(d, w1), T(0)) + norm2_(o0, o1);
}

// ---- post-processing value maps (post_processing.py:380-560; wrappers modifications.py:1361-1587) ---------------------------
// args as laid out by program.py::mod_pre: (amplitude, width), (threshold), (width), (b, width), (width, b/width, ground)
template <typename S, typename A>
AB_DEV S pp_sigmoid(const S& v, A a) { typedef typename S::scalar T; return div_(constant_like(v, a[0]), exp_(v * (T(4) * rcp_arg(a[1]))) + T(1)); }
template <typename S, typename A>
AB_DEV S pp_pos_sigmoid(const S& v, A a) { typedef typename S::scalar T; return div_(constant_like(v, a[0]), exp_((v - a[1]) * (T(4) * rcp_arg(a[1]))) + T(1)); }
template <typename S, typename A>
AB_DEV S pp_capped_exp(const S& v, A a) { typedef typename S::scalar T; return min_(exp_(v * (T(-4) * rcp_arg(a[1]))), T(1)) * a[0]; }
template <typename S, typename A>
AB_DEV S pp_hard_bin(const S& v, A a) { typedef typename S::scalar T; return select_(le_(v, a[0]), constant_like(v, T(1)), constant_like(v, T(0))); }
template <typename S, typename A>
AB_DEV S pp_linear(const S& v, A a) { typedef typename S::scalar T; return clamp_(T(1) - v * rcp_arg(a[1]), T(0), T(1)) * a[0]; }
template <typename S, typename A>
AB_DEV S pp_relu(const S& v, A a) { typedef typename S::scalar T; return max_(v * rcp_arg(a[0]), T(0)); }
template <typename S, typename A>
AB_DEV S pp_smooth_relu(const S& v, A a) { typedef typename S::scalar T;
  S u = v * rcp_arg(a[1]);
  return (u + sqrt_(fma_(u, u, constant_like(u, a[0])))) * T(0.5);
}
template <typename S, typename A>
AB_DEV S pp_slowstart(const S& v, A a) { typedef typename S::scalar T;
  S u = max_(v * rcp_arg(a[0]), T(0));
  return sqrt_(fma_(u, u, constant_like(u, a[1]))) - a[2];
}
template <typename S, typename A>
AB_DEV S pp_gauss_boundary(const S& v, A a) { typedef typename S::scalar T;
  S u = v * rcp_arg(a[1]);
  return exp_(u * u * T(-4)) * a[0];
}
template <typename S, typename A>
AB_DEV S pp_gauss_falloff(const S& v, A a) { typedef typename S::scalar T;
  S u = max_(v, T(0)) * rcp_arg(a[1]);
  return exp_(u * u * T(-4)) * a[0];
}

// ---- combine ops (combine.py:12-78) --------------------------------------------------------------------------------------
template <typename S, typename U>
AB_DEV S smin_poly2(const S& x, const S& y, const U& w) {  // combine.py:12-18
  typedef typename S::scalar T;
  S h = max_(w - abs_(x - y), T(0)) * rcp_arg(w);
  return min_(x, y) - h * h * (w * T(0.25));
}
template <typename S, typename U>
AB_DEV S smin_poly3(const S& x, const S& y, const U& w) {  // combine.py:20-26
  typedef typename S::scalar T;
  S h = max_(w - abs_(x - y), T(0)) * rcp_arg(w);
  return min_(x, y) - h * h * h * (w * T(1.0 / 6.0));
}
// polynomial smooth minima of dual numbers with a plain width, in closed form: with u = x - y, h = max(w - |u|, 0) / w,
//   d smin_n = d min(x, y) + (n h^(n-1) / (2 n)) sign(u) (dx - dy)        (n = 2: h / 2 ... n = 3: h^2 / 2)
// (same value expression as the generic form; 3 instructions per tangent component instead of a chain of dual products)
template <typename P, int K>
AB_DEV Dual<P, K> smin_poly_dual(const Dual<P, K>& x, const Dual<P, K>& y, const typename P::scalar& w, int order) {
  typedef typename P::scalar T;
  const P u = x.v - y.v;
  const P h = max_(w - abs_(u), T(0)) * s_rcp(w);
  Dual<P, K> r;
  const Mask<P::width> xle = le_(x.v, y.v);
  r.v = order == 2 ? min_(x.v, y.v) - h * h * (w * T(0.25)) : min_(x.v, y.v) - h * h * h * (w * T(1.0 / 6.0));
  // sign(u) from the comparison that picks the minimum (at u == 0 this gives the mean of the two tangents, which is the
  // derivative of the smooth minimum there); 0 where h == 0: plain min
  const P gm = (order == 2 ? h : h * h) * T(0.5);
  const P g = select_(xle, -gm, gm);
#pragma unroll
  for (int k = 0; k < K; k++) r.d[k] = fma_(g, x.d[k] - y.d[k], select_(xle, x.d[k], y.d[k]));
  return r;
}
template <typename P, int K>
AB_DEV Dual<P, K> smin_poly2(const Dual<P, K>& x, const Dual<P, K>& y, const typename P::scalar& w) { return smin_poly_dual(x, y, w, 2); }
template <typename P, int K>
AB_DEV Dual<P, K> smin_poly3(const Dual<P, K>& x, const Dual<P, K>& y, const typename P::scalar& w) { return smin_poly_dual(x, y, w, 3); }
// combine.py:29-34, evaluated in the shifted form (x e^{(x-m)/a} + y e^{(y-m)/a}) / (e^{(x-m)/a} + e^{(y-m)/a}),
// m = max(x,y): algebraically identical, but does not overflow in fp32 where exp(x/a) would for x/a > 88.
template <typename S, typename U>
AB_DEV S smax_boltz(const S& x, const S& y, const U& w) {
  typedef typename S::scalar T;
  const U iw = rcp_arg(w);
  S m = (aval(w) > T(0)) ? max_(x, y) : min_(x, y);
  S e1 = exp_((x - m) * iw);
  S e2 = exp_((y - m) * iw);
  return div_(x * e1 + y * e2, e1 + e2);
}

// ---- 3D primitives (sdf_3D.py) ----------------------------------------------------------------------------------------------
template <typename S, typename A>
AB_DEV S prim_sphere(const Pt<S>& p, A a) { typedef typename S::scalar T; return norm3_(p.x, p.y, p.z) - a[0]; }  // :25-27
template <typename S, typename A>
AB_DEV S prim_cylinder(const Pt<S>& p, A a) { typedef typename S::scalar T;  // :30-37 ; args radius, height/2
  S d0 = norm2_(p.x, p.y) - a[0];
  S d1 = abs_(p.z) - a[1];
  return min_(max_(d0, d1), T(0)) + norm2_(max_(d0, T(0)), max_(d1, T(0)));
}
template <typename S, typename A>
AB_DEV S prim_box(const Pt<S>& p, A a) { typedef typename S::scalar T;  // :40-47 ; args half sizes
  S q0 = abs_(p.x) - a[0], q1 = abs_(p.y) - a[1], q2 = abs_(p.z) - a[2];
  return norm3_(max_(q0, T(0)), max_(q1, T(0)), max_(q2, T(0))) + min_(max_(q0, max_(q1, q2)), T(0));
}
template <typename S, typename A>
AB_DEV S prim_torus(const Pt<S>& p, A a) { typedef typename S::scalar T;  // :50-53
  return norm2_(norm2_(p.x, p.y) - a[0], p.z) - a[1];
}
template <typename S, typename A>
AB_DEV S prim_chainlink(const Pt<S>& p, A a) { typedef typename S::scalar T;  // :56-61 ; args R, r, length/2
  S xx = p.x - clamp_(p.x, -a[2], a[2]);
  return norm2_(norm2_(xx, p.y) - a[0], p.z) - a[1];
}
template <typename S, typename A>
AB_DEV S prim_braid(const Pt<S>& p, A a) { typedef typename S::scalar T;  // :64-75 ; args length/2, R, r, pitch
  S s, c;
  sincos_(p.z * a[3], s, c);
  S xr = c * p.x - s * p.y;
  S yr = s * p.x + c * p.y;
  S zz = p.z - clamp_(p.z, -a[0], a[0]);
  return norm2_(norm2_(xr, zz) - a[1], yr) - a[2];
}
// shared by sdf_arc (sdf_2D.py:85-102) and sdf_arc_3d (sdf_3D.py:78-96): rotate by the centre angle, fold, subtract
// the closest point on the arc. args C, S, R, ea
template <typename S, typename U>
AB_DEV void arc_core(const S& x, const S& y, const U& C, const U& Sn, const U& R, const U& ea, S& dx, S& dy) {
  typedef typename S::scalar T;
  S xr = fma_(x, C, y * Sn);
  S yr = abs_(fma_(y, C, -(x * Sn)));
  S psi = clamp_(atan2_(yr, xr), T(0), ea);
  S s, c;
  sincos_(psi, s, c);
  dx = xr - c * R;
  dy = yr - s * R;
}
template <typename S, typename A>
AB_DEV S prim_arc3d(const Pt<S>& p, A a) { typedef typename S::scalar T;  // args R, r, C, S, ea
  S dx, dy;
  arc_core(p.x, p.y, a[2], a[3], a[0], a[4], dx, dy);
  return norm3_(dx, dy, p.z) - a[1];
}
template <typename S, typename A>
AB_DEV S prim_plane(const Pt<S>& p, A a) { typedef typename S::scalar T;  // :99-102
  return fma_(p.x, a[0], fma_(p.y, a[1], p.z * a[2])) - a[3];
}
template <typename S, typename A>
AB_DEV S prim_uplane(const Pt<S>& p, A a) { typedef typename S::scalar T;  // :105-108
  return abs_(fma_(p.x, a[0], fma_(p.y, a[1], p.z * a[2]))) - a[3];
}
template <typename S, typename A>
AB_DEV S prim_segment(const Pt<S>& p, A a) { typedef typename S::scalar T;  // :111-118 ; args a(3), ba(3), dot(ba,ba)
  S px = p.x - a[0], py = p.y - a[1], pz = p.z - a[2];
  S h = clamp_(fma_(px, a[3], fma_(py, a[4], pz * a[5])) * rcp_arg(a[6]), T(0), T(1));
  return norm3_(px - h * a[3], py - h * a[4], pz - h * a[5]);
}
template <typename S, typename A>
AB_DEV S prim_cone(const Pt<S>& p, A a) { typedef typename S::scalar T;  // :121-136 ; args q0, q1, zoff, dot(q,q)
  const auto q0 = a[0], q1 = a[1];
  S w0 = norm2_(p.x, p.y), w1 = p.z - a[2];
  S t = clamp_(fma_(w0, q0, w1 * q1) * rcp_arg(a[3]), T(0), T(1));
  S a0 = w0 - t * q0, a1 = w1 - t * q1;
  S b0 = w0 - clamp_(w0 * rcp_arg(q0), T(0), T(1)) * q0, b1 = w1 - q1;
  S d = min_(fma_(a0, a0, a1 * a1), fma_(b0, b0, b1 * b1));
  S s = max_(-(w0 * q1 - w1 * q0), -(w1 - q1));
  return mul_lane(sqrt_(d), value_sign(s));
}
template <typename S, typename A>
AB_DEV S prim_inf_cone(const Pt<S>& p, A a, bool oriented) { typedef typename S::scalar T;  // :139-157 ; args sin, cos
  const auto v0 = a[0], v1 = a[1];
  S q0 = norm2_(p.x, p.y), q1 = -p.z;
  S t = max_(fma_(q0, v0, q1 * v1), T(0));
  S d = norm2_(q0 - t * v0, q1 - t * v1);
  if (oriented) d = select_(lt_(q0 * v1 - q1 * v0, T(0)), -d, d);
  return d;
}
// shared by sdf_sector (sdf_2D.py:105-129) and sdf_solid_angle (sdf_3D.py:160-183) after rotation + fold
template <typename S, typename U>
AB_DEV S sector_core(const S& x, const S& y, const U& radius, const U& ad, const U& cad, const U& sad) {
  typedef typename S::scalar T;
  S phi = atan2_(y, x);
  S psi = clamp_(phi, T(0), ad);
  S s, c;
  sincos_(psi, s, c);
  S length = norm2_(x - c * radius, y - s * radius);
  S t = clamp_(fma_(x, cad, y * sad), T(0), radius);
  S m = norm2_(x - t * cad, y - t * sad);
  Mask<S::width> msk = le_(norm2_(x, y), radius) & le_(phi, ad);
  S out = min_(m, length);
  return select_(msk, -out, out);
}
template <typename S, typename A>
AB_DEV S prim_solid_angle(const Pt<S>& p, A a) { typedef typename S::scalar T;  // args radius, C, S, ad, cos ad, sin ad
  S xr = fma_(p.x, a[1], p.y * a[2]);
  S yr = norm2_(fma_(p.y, a[1], -(p.x * a[2])), p.z);
  return sector_core(xr, yr, a[0], a[3], a[4], a[5]);
}
template <typename S, typename T>
AB_DEV S dot3(const T* u, const S& x, const S& y, const S& z) { return fma_(x, u[0], fma_(y, u[1], z * u[2])); }
template <typename S, typename T>
AB_DEV S edge_sq(const T* s, T ss, const S& x, const S& y, const S& z) {  // one term of sdf_3D.py:203-206
  S h = clamp_(dot3(s, x, y, z) * s_rcp(ss), T(0), T(1));
  S t0 = h * s[0] - x, t1 = h * s[1] - y, t2 = h * s[2] - z;
  return fma_(t0, t0, fma_(t1, t1, t2 * t2));
}
template <typename S, typename A>
AB_DEV S prim_triangle3d(const Pt<S>& p, A g_) { typedef typename S::scalar T;  // :186-214 ; layout in program.py
  const T* g = raw_args(g_);  // vertex-derived table: structural
  S ax = p.x - g[0], ay = p.y - g[1], az = p.z - g[2];
  S bx = p.x - g[3], by = p.y - g[4], bz = p.z - g[5];
  S cx = p.x - g[6], cy = p.y - g[7], cz = p.z - g[8];
  auto sg = value_of(sign_(dot3(g + 21, ax, ay, az))) + value_of(sign_(dot3(g + 24, bx, by, bz))) +
            value_of(sign_(dot3(g + 27, cx, cy, cz)));
  S ex1 = min_(min_(edge_sq(g + 9, g[30], ax, ay, az), edge_sq(g + 12, g[31], bx, by, bz)),
               edge_sq(g + 15, g[32], cx, cy, cz));
  S dn = dot3(g + 18, ax, ay, az);
  S ex2 = dn * dn * rcp_arg(g[33]);
  return sqrt_(select_(lt_(sg, T(2)), ex1, ex2));
}
template <typename S, typename A>
AB_DEV S prim_quad3d(const Pt<S>& p, A g_) { typedef typename S::scalar T;  // :217-250
  const T* g = raw_args(g_);
  S ax = p.x - g[0], ay = p.y - g[1], az = p.z - g[2];
  S bx = p.x - g[3], by = p.y - g[4], bz = p.z - g[5];
  S cx = p.x - g[6], cy = p.y - g[7], cz = p.z - g[8];
  S dx = p.x - g[9], dy = p.y - g[10], dz = p.z - g[11];
  auto sg = value_of(sign_(dot3(g + 27, ax, ay, az))) + value_of(sign_(dot3(g + 30, bx, by, bz))) +
            value_of(sign_(dot3(g + 33, cx, cy, cz))) + value_of(sign_(dot3(g + 36, dx, dy, dz)));
  S ex1 = min_(min_(edge_sq(g + 21, g[42], dx, dy, dz), edge_sq(g + 18, g[41], cx, cy, cz)),
               min_(edge_sq(g + 12, g[39], ax, ay, az), edge_sq(g + 15, g[40], bx, by, bz)));
  S dn = dot3(g + 24, ax, ay, az);
  S ex2 = dn * dn * rcp_arg(g[43]);
  return sqrt_(select_(lt_(sg, T(3)), ex1, ex2));
}
template <typename S, typename A>
AB_DEV S prim_segline(const Pt<S>& p, A a_, int dim) { typedef typename S::scalar T;  // sdf_3D.py:264-271 / sdf_2D.py:191-198
  const T* a = raw_args(a_);  // vertex table: structural
  const int n = (int)a[0];
  const T* pts = a + 1;
  S best = constant_like(p.x, T(1e32));  // squared: the reference starts from 1e16 on the distance
  for (int i = 0; i + 1 < n; i++) {
    const T* u = pts + i * dim;
    const T* v = u + dim;
    T bx = v[0] - u[0], by = v[1] - u[1], bz = dim == 3 ? v[2] - u[2] : T(0);
    T bb = s_fma(bx, bx, s_fma(by, by, bz * bz));
    S px = p.x - u[0], py = p.y - u[1];
    S d2;
    if (dim == 3) {
      S pz = p.z - u[2];
      S h = clamp_(fma_(px, bx, fma_(py, by, pz * bz)) * s_rcp(bb), T(0), T(1));
      S t0 = px - h * bx, t1 = py - h * by, t2 = pz - h * bz;
      d2 = fma_(t0, t0, fma_(t1, t1, t2 * t2));
    } else {
      S h = clamp_(fma_(px, bx, py * by) * s_rcp(bb), T(0), T(1));
      S t0 = px - h * bx, t1 = py - h * by;
      d2 = fma_(t0, t0, t1 * t1);
    }
    best = min_(best, d2);
  }
  return sqrt_(best);
}

// ---- 2D primitives (sdf_2D.py) ----------------------------------------------------------------------------------------------
template <typename S, typename A>
AB_DEV S prim_circle(const Pt<S>& p, A a) { typedef typename S::scalar T; return norm2_(p.x, p.y) - a[0]; }  // :12-14
template <typename S, typename A>
AB_DEV S prim_neu_circle(const Pt<S>& p, A a) { typedef typename S::scalar T;  // :17-19 ; args radius, order
  const T ord = aval(a[1]);
  S ax = abs_(p.x), ay = abs_(p.y);
  if (isinf(ord)) return max_(ax, ay) - a[0];
  if (ord == T(1)) return ax + ay - a[0];
  if (ord == T(2)) return norm2_(ax, ay) - a[0];
  return pow_(pow_(ax, ord) + pow_(ay, ord), s_rcp(ord)) - a[0];
}
template <typename S, typename A>
AB_DEV S prim_box2d(const Pt<S>& p, A a) { typedef typename S::scalar T;  // :22-28
  S d0 = abs_(p.x) - a[0], d1 = abs_(p.y) - a[1];
  return norm2_(max_(d0, T(0)), max_(d1, T(0))) + min_(max_(d0, d1), T(0));
}
template <typename S, typename A>
AB_DEV S prim_segment2d(const Pt<S>& p, A a) { typedef typename S::scalar T;  // :31-38 ; args a(2), ba(2), dot
  S px = p.x - a[0], py = p.y - a[1];
  S h = clamp_(fma_(px, a[2], py * a[3]) * rcp_arg(a[4]), T(0), T(1));
  return norm2_(px - h * a[2], py - h * a[3]);
}
template <typename S, typename A>
AB_DEV S prim_rbox2d(const Pt<S>& p, A a_) { typedef typename S::scalar T;  // :41-57 ; args hx, hy, r0..r3
  const T* a = raw_args(a_);  // per-quadrant radius selection: structural
  constexpr int W = S::width;
  auto vx = value_of(p.x), vy = value_of(p.y);
  Pack<T, W> r;
#pragma unroll
  for (int i = 0; i < W; i++)
    r.v[i] = (vy.v[i] > T(0)) ? ((vx.v[i] < T(0)) ? a[5] : a[4]) : ((vx.v[i] > T(0)) ? a[3] : a[2]);
  S d0 = add_lane(abs_(p.x) - a[0], r), d1 = add_lane(abs_(p.y) - a[1], r);
  return norm2_(max_(d0, T(0)), max_(d1, T(0))) + add_lane(min_(max_(d0, d1), T(0)), -r);
}
template <typename S, typename A>
AB_DEV S prim_triangle2d(const Pt<S>& p, A g_) { typedef typename S::scalar T;  // :60-82 ; args p0,p1,p2,e0,e1,e2,ee(3),s
  const T* g = raw_args(g_);  // vertex-derived table: structural
  S d0, d1;
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const T ex = g[6 + 2 * i], ey = g[7 + 2 * i];
    S v0 = p.x - g[2 * i], v1 = p.y - g[2 * i + 1];
    S h = clamp_(fma_(v0, ex, v1 * ey) * rcp_arg(g[12 + i]), T(0), T(1));
    S q0 = v0 - h * ex, q1 = v1 - h * ey;
    S dd = fma_(q0, q0, q1 * q1);
    S cr = (v0 * ey - v1 * ex) * g[15];
    if (i == 0) {
      d0 = dd;
      d1 = cr;
    } else {
      d0 = min_(d0, dd);
      d1 = min_(d1, cr);
    }
  }
  return -mul_lane(sqrt_(d0), value_sign(d1));
}
template <typename S, typename A>
AB_DEV S prim_arc(const Pt<S>& p, A a) { typedef typename S::scalar T;  // :85-102 ; args R, C, S, ea
  S dx, dy;
  arc_core(p.x, p.y, a[1], a[2], a[0], a[3], dx, dy);
  return norm2_(dx, dy);
}
template <typename S, typename A>
AB_DEV S prim_sector(const Pt<S>& p, A a) { typedef typename S::scalar T;  // :105-129
  S xr = fma_(p.x, a[1], p.y * a[2]);
  S yr = abs_(fma_(p.y, a[1], -(p.x * a[2])));
  return sector_core(xr, yr, a[0], a[3], a[4], a[5]);
}
template <typename S, typename A>
AB_DEV S prim_inf_sector(const Pt<S>& p, A a) { typedef typename S::scalar T;  // :132-150 ; args C, S, ad, cos ad, sin ad
  S xr = fma_(p.x, a[0], p.y * a[1]);
  S yr = abs_(fma_(p.y, a[0], -(p.x * a[1])));
  S phi = atan2_(yr, xr);
  S t = max_(fma_(xr, a[3], yr * a[4]), T(0));
  S m = norm2_(xr - t * a[3], yr - t * a[4]);
  return mul_lane(m, value_sign(phi - a[2]));
}
template <typename S, typename A>
AB_DEV S prim_ngon(const Pt<S>& p, A a) { typedef typename S::scalar T;  // :153-177 ; args radius, alpha, tx, ty, nox, noy, l
  S phi = atan2_(p.y, p.x);
  phi = select_(lt_(phi, T(0)), phi + T(6.283185307179586476925286766559), phi);
  phi = mod_(phi, a[1]);
  S rr = norm2_(p.x, p.y);
  S s, c;
  sincos_(phi, s, c);
  S q0 = c * rr - a[0], q1 = s * rr;
  S h = clamp_(fma_(q0, a[2], q1 * a[3]), T(0), a[6]);
  S len = norm2_(q0 - h * a[2], q1 - h * a[3]);
  return mul_lane(len, value_sign(fma_(q0, a[4], q1 * a[5])));
}

// interior of a simple closed polygon by the crossing-number rule (horizontal ray towards +x, half-open in y): equals the
// union of the reference's ear-clipping triangles (triangulation_functions.py:390-430) off the boundary
template <typename T, int W>
AB_DEV void polygon_inside(const Pack<T, W>& vx, const Pack<T, W>& vy, const T* pts, int n, bool (&inside)[W]) {
#pragma unroll
  for (int j = 0; j < W; j++) inside[j] = false;
  for (int i = 0; i < n; i++) {
    const int k = (i + 1 == n) ? 0 : i + 1;
    const T ax = pts[2 * i], ay = pts[2 * i + 1], bx = pts[2 * k] - ax, by = pts[2 * k + 1] - ay;
#pragma unroll
    for (int j = 0; j < W; j++) {
      const bool ca = ay > vy.v[j], cb = (ay + by) > vy.v[j];
      if (ca != cb) {
        const T xi = bx * (vy.v[j] - ay) / by + ax;
        inside[j] = inside[j] != (vx.v[j] < xi);
      }
    }
  }
}

// sdf_polygon_2d (sdf_2D.py:201-218) for simple polygons: unsigned distance = min over the closed edge loop, interior
// by the crossing-number rule. args: n, then n (x, y) vertices.
template <typename S, typename A>
AB_DEV S prim_polygon2d(const Pt<S>& p, A a_) {
  typedef typename S::scalar T;
  constexpr int W = S::width;
  const T* a = raw_args(a_);  // vertex table: structural
  const int n = (int)a[0];
  const T* pts = a + 1;
  auto vx = value_of(p.x), vy = value_of(p.y);
  S best = constant_like(p.x, T(1e32));
  for (int i = 0; i < n; i++) {
    const int k = (i + 1 == n) ? 0 : i + 1;
    const T ax = pts[2 * i], ay = pts[2 * i + 1], bx = pts[2 * k] - ax, by = pts[2 * k + 1] - ay;
    const T bb = s_fma(bx, bx, by * by);
    S px = p.x - ax, py = p.y - ay;
    S h = clamp_(fma_(px, bx, py * by) * s_rcp(bb), T(0), T(1));
    S t0 = px - h * bx, t1 = py - h * by;
    best = min_(best, fma_(t0, t0, t1 * t1));
  }
  bool inside[W];
  polygon_inside(vx, vy, pts, n, inside);
  Pack<T, W> sg;
#pragma unroll
  for (int j = 0; j < W; j++) sg.v[j] = inside[j] ? T(-1) : T(1);
  return mul_lane(sqrt_(best), sg);
}

// POLY_SIGN: acc * interior sign evaluated at saved coordinates (x, y).
//   rule 0: SegmentedLine.polygon() / SegmentedParametricCurve.polygon() (geom_2d.py:530-555, 601-626): interior_polygon of
//           the control points, -1 inside / +1 outside; table: n, then n (x, y) vertices;
//   rule 1: ParametricCurve.shape() (geom_2d.py:440-452), term by term: for every segment i of the sampled curve whose
//           half-open x interval [lx, ux) holds the point, multiply by sign(dot(p - P_i, n_i)), n_i = (-t_y, |t_x|) (the
//           reference takes the absolute value of the normal's y component only); table: n, then n records
//           (P_ix, P_iy, lx, ux, n_ix, n_iy).
template <typename S, typename A>
AB_DEV S op_poly_sign(const S& acc, const S& sx, const S& sy, A a_, int rule) {
  typedef typename S::scalar T;
  constexpr int W = S::width;
  const T* a = raw_args(a_);
  const int n = (int)a[0];
  auto vx = value_of(sx), vy = value_of(sy);
  Pack<T, W> sg;
  if (rule == 0) {
    bool inside[W];
    polygon_inside(vx, vy, a + 1, n, inside);
#pragma unroll
    for (int j = 0; j < W; j++) sg.v[j] = inside[j] ? T(-1) : T(1);
  } else {
#pragma unroll
    for (int j = 0; j < W; j++) sg.v[j] = T(1);
    for (int i = 0; i < n; i++) {
      const T* r = a + 1 + 6 * i;
#pragma unroll
      for (int j = 0; j < W; j++) {
        if (vx.v[j] >= r[2] && vx.v[j] < r[3]) {
          const T d = s_fma(vx.v[j] - r[0], r[4], (vy.v[j] - r[1]) * r[5]);
          sg.v[j] = sg.v[j] * (d > T(0) ? T(1) : (d < T(0) ? T(-1) : T(0)));
        }
      }
    }
  }
  return mul_lane(acc, sg);
}

}  // namespace ab
