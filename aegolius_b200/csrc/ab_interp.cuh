// ab_interp.cuh — the SDF interpreter kernel (sm_100a).
//
// One launch evaluates one flattened geometry tree on a slab of grid points (or on an explicit point list). The whole
// program (op list + argument pool) is a __grid_constant__ kernel parameter: it lives in constant bank 0, costs no
// separate H2D copy and cannot race between concurrent launches. Control flow is identical for every point, so opcode
// dispatch is warp-uniform; per thread W consecutive points along the fastest grid axis live in registers (Pack<T,W>),
// the saved-coordinate / saved-value stacks live in shared memory as conflict-free 16-byte columns, and the field is
// written with one 128-bit streaming store per thread (coalesced: a warp writes 512 contiguous bytes).
//
// Grid layout (generate_grid, helper_functions.py:23-93): flat k = (ix*ny + iy)*nz + iz, z fastest; the slab is a range
// of ix planes, i.e. one contiguous range of k. Coordinates are regenerated from (ix,iy,iz): 0 bytes read per point.
#pragma once
#include "../../include/aegolius_b200.h"
#include "ab_ops.cuh"

namespace ab {

struct GridK {
  // index axes (ix, iy, iz), iz fastest. 2D grids are (nx, ny, 1) with size_z = 0, so z == 0.
  uint32_t n1, n2;          // ny, nz
  uint32_t plane;           // ny * nz
  uint32_t i0_begin;        // first ix plane of the slab
  uint32_t last[3];         // res-1 per axis (for the exact linspace end point)
  double start[3], step[3], stop[3];  // linspace parameters per axis (fp64 path: bit-identical to numpy)
  float hi[3], lo[3], centre[3];      // fp32 path: x = (i - centre) * (hi + lo), one rounding
  double inv_plane, inv_n2;           // reciprocals for the index split
};

template <typename T>
struct Vec4;
template <>
struct Vec4<float> {
  typedef float4 type;
};
template <>
struct Vec4<double> {
  typedef double4 type;
};

template <typename T>
struct KParams {
  uint64_t n;        // points in this launch
  T* out;            // (n,)
  T* grad;           // (K, grad_stride) or nullptr
  uint64_t grad_stride;
  const void* co;    // points mode: (3, co_stride) of float/double
  uint64_t co_stride;
  int32_t co_is_f64;
  int32_t grid_mode;
  GridK g;
  uint32_t n_ops, n_pslots, n_vslots;
  const void* blob[AB_MAX_BLOBS];  // (x, y, z, 0) records of T
  uint32_t blob_count[AB_MAX_BLOBS];
  ab_op ops[AB_MAX_OPS];
  T args[AB_MAX_ARGS];
};

// ---- coordinate generation ------------------------------------------------------------------------------------------
AB_DEV float grid_coord(const GridK& g, int c, uint32_t i, float) {
  float k = (float)i - g.centre[c];
  return fmaf(k, g.hi[c], k * g.lo[c]);
}
AB_DEV double grid_coord(const GridK& g, int c, uint32_t i, double) {
  // np.linspace: arange(n) * step + start, last sample forced to stop (no FMA contraction: bit-identical)
  return (i == g.last[c]) ? g.stop[c] : __dadd_rn(__dmul_rn((double)i, g.step[c]), g.start[c]);
}

// ---- stack in shared memory: element [slot][tid] is one 16-byte Pack column --------------------------------------------
template <typename P>
AB_DEV void st_pack(P* base, int slot, int nt, const P& v) { base[slot * nt + threadIdx.x] = v; }
template <typename P>
AB_DEV P ld_pack(const P* base, int slot, int nt) { return base[slot * nt + threadIdx.x]; }

template <typename T, int W>
struct StackIO {
  typedef Pack<T, W> P;
  static constexpr int cols = 1;
  static AB_DEV void st(P* base, int slot, int nt, const P& v) { st_pack(base, slot, nt, v); }
  static AB_DEV P ld(const P* base, int slot, int nt) { return ld_pack(base, slot, nt); }
};
template <typename S>
struct StackOf;
template <typename T, int W>
struct StackOf<Pack<T, W>> {
  typedef Pack<T, W> P;
  static constexpr int cols = 1;
  static AB_DEV void st(P* base, int slot, int nt, const P& v) { st_pack(base, slot, nt, v); }
  static AB_DEV P ld(const P* base, int slot, int nt) { return ld_pack(base, slot, nt); }
};
template <typename Pk, int K>
struct StackOf<Dual<Pk, K>> {
  typedef Pk P;
  static constexpr int cols = 1 + K;
  static AB_DEV void st(P* base, int slot, int nt, const Dual<Pk, K>& v) {
    st_pack(base, slot * cols, nt, v.v);
#pragma unroll
    for (int k = 0; k < K; k++) st_pack(base, slot * cols + 1 + k, nt, v.d[k]);
  }
  static AB_DEV Dual<Pk, K> ld(const P* base, int slot, int nt) {
    Dual<Pk, K> r;
    r.v = ld_pack(base, slot * cols, nt);
#pragma unroll
    for (int k = 0; k < K; k++) r.d[k] = ld_pack(base, slot * cols + 1 + k, nt);
    return r;
  }
};

// ---- 128-bit streaming stores ---------------------------------------------------------------------------------------------
AB_DEV void store_pack(float* dst, const Pack<float, 4>& v, uint64_t idx, uint64_t n, bool aligned) {
  if (aligned && idx + 4 <= n) {
    __stcs(reinterpret_cast<float4*>(dst + idx), make_float4(v.v[0], v.v[1], v.v[2], v.v[3]));
  } else {
#pragma unroll
    for (int i = 0; i < 4; i++)
      if (idx + i < n) __stcs(dst + idx + i, v.v[i]);
  }
}
AB_DEV void store_pack(float* dst, const Pack<float, 2>& v, uint64_t idx, uint64_t n, bool aligned) {
  if (aligned && idx + 2 <= n) {
    __stcs(reinterpret_cast<float2*>(dst + idx), make_float2(v.v[0], v.v[1]));
  } else {
#pragma unroll
    for (int i = 0; i < 2; i++)
      if (idx + i < n) __stcs(dst + idx + i, v.v[i]);
  }
}
AB_DEV void store_pack(double* dst, const Pack<double, 2>& v, uint64_t idx, uint64_t n, bool aligned) {
  if (aligned && idx + 2 <= n) {
    __stcs(reinterpret_cast<double2*>(dst + idx), make_double2(v.v[0], v.v[1]));
  } else {
#pragma unroll
    for (int i = 0; i < 2; i++)
      if (idx + i < n) __stcs(dst + idx + i, v.v[i]);
  }
}
template <typename T>
AB_DEV void store_pack(T* dst, const Pack<T, 1>& v, uint64_t idx, uint64_t n, bool) {
  if (idx < n) __stcs(dst + idx, v.v[0]);
}

// ---- seeding ------------------------------------------------------------------------------------------------------------
template <typename T, int W>
AB_DEV void seed(Pt<Pack<T, W>>& p, const Pack<T, W>& x, const Pack<T, W>& y, const Pack<T, W>& z) {
  p.x = x;
  p.y = y;
  p.z = z;
}
template <typename T, int W, int K>
AB_DEV void seed(Pt<Dual<Pack<T, W>, K>>& p, const Pack<T, W>& x, const Pack<T, W>& y, const Pack<T, W>& z) {
  typedef Dual<Pack<T, W>, K> D;
  p.x = D(x);
  p.y = D(y);
  p.z = D(z);
  if (K == 3) {  // spatial tangents: d/dx, d/dy, d/dz
    p.x.d[0] = Pack<T, W>(T(1));
    p.y.d[1 % K] = Pack<T, W>(T(1));
    p.z.d[2 % K] = Pack<T, W>(T(1));
  }
}
template <typename T, int W>
AB_DEV void emit(const KParams<T>& kp, const Pack<T, W>& acc, uint64_t idx, bool aligned) {
  store_pack(kp.out, acc, idx, kp.n, aligned);
}
template <typename T, int W, int K>
AB_DEV void emit(const KParams<T>& kp, const Dual<Pack<T, W>, K>& acc, uint64_t idx, bool aligned) {
  store_pack(kp.out, acc.v, idx, kp.n, aligned);
  const bool ga = aligned && (kp.grad_stride % W == 0);
#pragma unroll
  for (int k = 0; k < K; k++) store_pack(kp.grad + (uint64_t)k * kp.grad_stride, acc.d[k], idx, kp.n, ga);
}

// brute-force nearest cloud point inside the interpreter (sdf_3D.py:283-286): uniform (broadcast) loads of float4
// records, exact (q-p)^2 form. The dedicated kernel in ab_nn.cu is the fast path for a bare cloud; this one lets a
// cloud sit anywhere in a tree.
template <typename S, typename T>
AB_DEV S prim_point_cloud(const Pt<S>& p, const void* __restrict__ cloud_v, uint32_t m, int dim) {
  typedef typename Vec4<T>::type V4;
  const V4* __restrict__ cloud = reinterpret_cast<const V4*>(cloud_v);
  constexpr int W = S::width;
  auto vx = value_of(p.x), vy = value_of(p.y), vz = value_of(p.z);
  T best[W];
  uint32_t bi[W];
#pragma unroll
  for (int i = 0; i < W; i++) {
    best[i] = T(3.0e38);
    bi[i] = 0;
  }
  for (uint32_t j = 0; j < m; j++) {
    const V4 c = cloud[j];
#pragma unroll
    for (int i = 0; i < W; i++) {
      T dx = vx.v[i] - (T)c.x, dy = vy.v[i] - (T)c.y, dz = (dim == 3) ? vz.v[i] - (T)c.z : T(0);
      T d2 = s_fma(dx, dx, s_fma(dy, dy, dz * dz));
      if (d2 < best[i]) {
        best[i] = d2;
        bi[i] = j;
      }
    }
  }
  // rebuild the distance to the winner with tangents (d|q-c|/dq = (q-c)/|q-c|)
  Pack<T, W> cx, cy, cz;
#pragma unroll
  for (int i = 0; i < W; i++) {
    const V4 c = cloud[bi[i]];
    cx.v[i] = (T)c.x;
    cy.v[i] = (T)c.y;
    cz.v[i] = (dim == 3) ? (T)c.z : T(0);
  }
  S dx = add_lane(p.x, -cx), dy = add_lane(p.y, -cy);
  if (dim == 3) return norm3_(dx, dy, add_lane(p.z, -cz));
  return norm2_(dx, dy);
}

// ---- the interpreter -------------------------------------------------------------------------------------------------------
template <typename S, typename T>
__global__ void __launch_bounds__(128) ab_interp_kernel(const __grid_constant__ KParams<T> kp) {
  constexpr int W = S::width;
  const int NT = blockDim.x;
  typedef Pack<T, W> P;
  typedef StackOf<S> SK;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  P* pstack = reinterpret_cast<P*>(smem_raw);                         // [n_pslots*3*cols][NT]
  P* vstack = pstack + (size_t)kp.n_pslots * 3 * SK::cols * NT;       // [n_vslots*cols][NT]

  const uint64_t tile_pts = (uint64_t)NT * W;
  const uint64_t n_tiles = (kp.n + tile_pts - 1) / tile_pts;

  for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const uint64_t idx = tile * tile_pts + (uint64_t)threadIdx.x * W;  // first local point of this thread
    P cx, cy, cz;
    if (kp.grid_mode) {
      // split the flat index of the first point, then walk W points with carries
      uint64_t k = idx < kp.n ? idx : (kp.n - 1);
      uint32_t i0 = (uint32_t)((double)k * kp.g.inv_plane);
      uint64_t rem64 = k - (uint64_t)i0 * kp.g.plane;
      if ((int64_t)rem64 < 0) { i0--; rem64 += kp.g.plane; }
      if (rem64 >= kp.g.plane) { i0++; rem64 -= kp.g.plane; }
      uint32_t rem = (uint32_t)rem64;
      uint32_t i1 = (uint32_t)((double)rem * kp.g.inv_n2);
      int32_t r2 = (int32_t)(rem - i1 * kp.g.n2);
      if (r2 < 0) { i1--; r2 += kp.g.n2; }
      if ((uint32_t)r2 >= kp.g.n2) { i1++; r2 -= kp.g.n2; }
      uint32_t i2 = (uint32_t)r2;
      i0 += kp.g.i0_begin;
      T c0 = T(0), c1 = T(0);
      bool fresh = true;
#pragma unroll
      for (int j = 0; j < W; j++) {
        if (fresh) {
          c0 = grid_coord(kp.g, 0, i0, T());
          c1 = grid_coord(kp.g, 1, i1, T());
          fresh = false;
        }
        cx.v[j] = c0;
        cy.v[j] = c1;
        cz.v[j] = grid_coord(kp.g, 2, i2, T());
        if (++i2 == kp.g.n2) {
          i2 = 0;
          fresh = true;
          if (++i1 == kp.g.n1) {
            i1 = 0;
            ++i0;
          }
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < W; j++) {
        uint64_t k = idx + j < kp.n ? idx + j : kp.n - 1;
        if (kp.co_is_f64) {
          const double* c = (const double*)kp.co;
          cx.v[j] = (T)__ldcs(c + k);
          cy.v[j] = (T)__ldcs(c + kp.co_stride + k);
          cz.v[j] = (T)__ldcs(c + 2 * kp.co_stride + k);
        } else {
          const float* c = (const float*)kp.co;
          cx.v[j] = (T)__ldcs(c + k);
          cy.v[j] = (T)__ldcs(c + kp.co_stride + k);
          cz.v[j] = (T)__ldcs(c + 2 * kp.co_stride + k);
        }
      }
    }

    Pt<S> p;
    seed(p, cx, cy, cz);
    S acc = constant_like(p.x, T(0));

    for (uint32_t pc = 0; pc < kp.n_ops; pc++) {
      const ab_op op = kp.ops[pc];
      const T* a = kp.args + op.arg;
      const int sa = op.a;
      switch (op.opcode) {
        case AB_OP_END: pc = kp.n_ops; break;
        case AB_OP_SAVE_P:
          SK::st(pstack, sa * 3 + 0, NT, p.x);
          SK::st(pstack, sa * 3 + 1, NT, p.y);
          SK::st(pstack, sa * 3 + 2, NT, p.z);
          break;
        case AB_OP_LOAD_P:
          p.x = SK::ld(pstack, sa * 3 + 0, NT);
          p.y = SK::ld(pstack, sa * 3 + 1, NT);
          p.z = SK::ld(pstack, sa * 3 + 2, NT);
          break;
        case AB_OP_PUSH_V: SK::st(vstack, sa, NT, acc); break;
        // coordinate ops
        case AB_OP_AFFINE: op_affine(p, a); break;
        case AB_OP_TRANSLATE: op_translate(p, a); break;
        case AB_OP_SCALE_P: op_scale_p(p, a); break;
        case AB_OP_ELONGATE: op_elongate(p, a); break;
        case AB_OP_TWIST: op_twist(p, a); break;
        case AB_OP_BEND: op_bend(p, a); break;
        case AB_OP_ABSX_SUB: op_absx_sub(p, a); break;
        case AB_OP_SYMMETRY:
          if (sa == 0) p.x = abs_(p.x);
          else if (sa == 1) p.y = abs_(p.y);
          else p.z = abs_(p.z);
          break;
        case AB_OP_ROTSYM: op_rotsym(p, a); break;
        case AB_OP_REVOLVE: op_revolve(p, a); break;
        case AB_OP_AXIS_REVOLVE: op_axis_revolve(p, a); break;
        case AB_OP_REP_INF: op_rep_inf(p, a); break;
        case AB_OP_REP_FIN: op_rep_fin(p, a); break;
        case AB_OP_LIN_INST: op_lin_inst(p, a, sa); break;
        case AB_OP_CURVE_INST: op_curve_inst(p, a, sa); break;
        case AB_OP_ZERO_Z: p.z = constant_like(p.z, T(0)); break;
        // value ops
        case AB_OP_ROUND: acc = acc - a[0]; break;
        case AB_OP_ABS: acc = abs_(acc); break;
        case AB_OP_NEG: acc = -acc; break;
        case AB_OP_SIGN: acc = sign_(acc); break;
        case AB_OP_ONION: acc = abs_(acc) - a[0]; break;
        case AB_OP_CONCENTRIC: acc = abs_(acc - a[0]); break;
        case AB_OP_SCALE_V: acc = acc * a[0]; break;
        case AB_OP_EXTRUDE_BEGIN:
          SK::st(vstack, sa, NT, abs_(p.z) - a[0]);
          p.z = constant_like(p.z, T(0));
          break;
        case AB_OP_EXTRUDE_END: acc = op_extrude_end<S, T>(acc, SK::ld(vstack, sa, NT)); break;
        // post-processing (post_processing.py:380-560)
        case AB_OP_PP_SIGMOID: acc = div_(constant_like(acc, a[0]), exp_(acc * (T(4) * s_rcp(a[1]))) + T(1)); break;
        case AB_OP_PP_POS_SIGMOID:
          acc = div_(constant_like(acc, a[0]), exp_((acc - a[1]) * (T(4) * s_rcp(a[1]))) + T(1));
          break;
        case AB_OP_PP_CAPPED_EXP: acc = min_(exp_(acc * (T(-4) * s_rcp(a[1]))), T(1)) * a[0]; break;
        case AB_OP_PP_HARD_BIN:
          acc = select_(le_(acc, a[0]), constant_like(acc, T(1)), constant_like(acc, T(0)));
          break;
        case AB_OP_PP_LINEAR: acc = clamp_(T(1) - acc * s_rcp(a[1]), T(0), T(1)) * a[0]; break;
        case AB_OP_PP_RELU: acc = max_(acc * s_rcp(a[0]), T(0)); break;
        case AB_OP_PP_SMOOTH_RELU: {
          S v = acc * s_rcp(a[1]);
          acc = (v + sqrt_(fma_(v, v, constant_like(v, a[0])))) * T(0.5);
        } break;
        case AB_OP_PP_SLOWSTART: {
          S v = max_(acc * s_rcp(a[0]), T(0));
          acc = sqrt_(fma_(v, v, constant_like(v, a[1]))) - a[2];
        } break;
        case AB_OP_PP_GAUSS_BOUNDARY: {
          S v = acc * s_rcp(a[1]);
          acc = exp_(v * v * T(-4)) * a[0];
        } break;
        case AB_OP_PP_GAUSS_FALLOFF: {
          S v = max_(acc, T(0)) * s_rcp(a[1]);
          acc = exp_(v * v * T(-4)) * a[0];
        } break;
        // combine: acc = f(V[a], acc)
        case AB_OP_C_UNION: acc = min_(SK::ld(vstack, sa, NT), acc); break;
        case AB_OP_C_INTERSECT: acc = max_(SK::ld(vstack, sa, NT), acc); break;
        case AB_OP_C_SUBTRACT: acc = max_(SK::ld(vstack, sa, NT), -acc); break;
        case AB_OP_C_SUM: acc = SK::ld(vstack, sa, NT) + acc; break;
        case AB_OP_C_DIFF: acc = SK::ld(vstack, sa, NT) - acc; break;
        case AB_OP_C_SMIN2: acc = smin_poly2(SK::ld(vstack, sa, NT), acc, a[0]); break;
        case AB_OP_C_SMIN3: acc = smin_poly3(SK::ld(vstack, sa, NT), acc, a[0]); break;
        case AB_OP_C_SMAX3: acc = -smin_poly3(-SK::ld(vstack, sa, NT), -acc, a[0]); break;
        case AB_OP_C_SSUB3: acc = -smin_poly3(-SK::ld(vstack, sa, NT), acc, a[0]); break;
        case AB_OP_C_BOLTZ_INT: acc = smax_boltz(SK::ld(vstack, sa, NT), acc, a[0]); break;
        case AB_OP_C_BOLTZ_SUB: acc = smax_boltz(SK::ld(vstack, sa, NT), -acc, a[0]); break;
        // 3D primitives
        case AB_OP_P_SPHERE: acc = prim_sphere(p, a); break;
        case AB_OP_P_CYLINDER: acc = prim_cylinder(p, a); break;
        case AB_OP_P_BOX: acc = prim_box(p, a); break;
        case AB_OP_P_TORUS: acc = prim_torus(p, a); break;
        case AB_OP_P_CHAINLINK: acc = prim_chainlink(p, a); break;
        case AB_OP_P_BRAID: acc = prim_braid(p, a); break;
        case AB_OP_P_ARC3D: acc = prim_arc3d(p, a); break;
        case AB_OP_P_PLANE: acc = prim_plane(p, a); break;
        case AB_OP_P_UPLANE: acc = prim_uplane(p, a); break;
        case AB_OP_P_SEGMENT: acc = prim_segment(p, a); break;
        case AB_OP_P_CONE: acc = prim_cone(p, a); break;
        case AB_OP_P_OINF_CONE: acc = prim_inf_cone(p, a, true); break;
        case AB_OP_P_INF_CONE: acc = prim_inf_cone(p, a, false); break;
        case AB_OP_P_SOLID_ANGLE: acc = prim_solid_angle(p, a); break;
        case AB_OP_P_TRIANGLE3D: acc = prim_triangle3d(p, a); break;
        case AB_OP_P_QUAD3D: acc = prim_quad3d(p, a); break;
        case AB_OP_P_SEGLINE: acc = prim_segline(p, a, 3); break;
        case AB_OP_P_AXIS: acc = (sa == 0 ? p.x : (sa == 1 ? p.y : p.z)) - a[0]; break;
        case AB_OP_P_POINT_CLOUD: acc = prim_point_cloud<S, T>(p, kp.blob[op.b], kp.blob_count[op.b], sa); break;
        // 2D primitives
        case AB_OP_P_CIRCLE: acc = prim_circle(p, a); break;
        case AB_OP_P_NEU_CIRCLE: acc = prim_neu_circle(p, a); break;
        case AB_OP_P_BOX2D: acc = prim_box2d(p, a); break;
        case AB_OP_P_SEGMENT2D: acc = prim_segment2d(p, a); break;
        case AB_OP_P_RBOX2D: acc = prim_rbox2d(p, a); break;
        case AB_OP_P_TRIANGLE2D: acc = prim_triangle2d(p, a); break;
        case AB_OP_P_ARC: acc = prim_arc(p, a); break;
        case AB_OP_P_SECTOR: acc = prim_sector(p, a); break;
        case AB_OP_P_INF_SECTOR: acc = prim_inf_sector(p, a); break;
        case AB_OP_P_NGON: acc = prim_ngon(p, a); break;
        case AB_OP_P_SEGLINE2D: acc = prim_segline(p, a, 2); break;
        default: break;  // unknown opcodes are rejected on the host (AB_EUNSUPPORTED_OP)
      }
    }
    const bool aligned = ((reinterpret_cast<uintptr_t>(kp.out) & 15) == 0);
    emit(kp, acc, idx, aligned);
  }
}

// ---- launcher (one explicit instantiation per translation unit: ab_interp_*.cu) ---------------------------------------------
struct LaunchCfg {
  int sms;
  size_t smem_optin;
};
// returns cudaSuccess or the CUDA error; *status is AB_OK / AB_ETOOLARGE
template <typename S, typename T>
cudaError_t launch_interp(const KParams<T>& kp, const LaunchCfg& cfg, cudaStream_t st, int* status);

#ifdef AB_INTERP_INSTANTIATE
template <typename S, typename T>
cudaError_t launch_interp(const KParams<T>& kp, const LaunchCfg& cfg, cudaStream_t st, int* status) {
  typedef StackOf<S> SK;
  *status = AB_OK;
  const size_t per_thread = (size_t)sizeof(typename SK::P) * SK::cols * ((size_t)kp.n_pslots * 3 + kp.n_vslots);
  // prefer 128 threads; shrink the CTA when the stacks would not leave room for >= 2 CTAs per SM
  int nt = 128;
  if (per_thread * 128 > 96 * 1024) nt = 64;
  if (per_thread * nt > cfg.smem_optin) nt = 32;
  const size_t smem = per_thread * nt;
  if (smem > cfg.smem_optin) {
    *status = AB_ETOOLARGE;
    return cudaSuccess;
  }
  auto kern = ab_interp_kernel<S, T>;
  cudaError_t e;
  if (smem > 48 * 1024) {
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  int occ = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, nt, smem);
  if (e != cudaSuccess) return e;
  if (occ < 1) {
    *status = AB_ETOOLARGE;
    return cudaSuccess;
  }
  const uint64_t tile_pts = (uint64_t)nt * S::width;
  const uint64_t n_tiles = (kp.n + tile_pts - 1) / tile_pts;
  const uint64_t resident = (uint64_t)cfg.sms * occ;  // persistent CTAs: a whole number of resident waves
  const unsigned grid = (unsigned)(n_tiles < resident ? n_tiles : resident);
  kern<<<grid, nt, smem, st>>>(kp);
  return cudaGetLastError();
}
#endif

}  // namespace ab
