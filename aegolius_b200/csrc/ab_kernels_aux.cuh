// ab_kernels_aux.cuh — the two non-interpreter kernels of the path:
//   * ab_nn_kernel : point cloud -> unsigned distance field (sdf_point_cloud_3d/2d, sdf_3D.py:283-286,
//                    sdf_2D.py:221-224) by tiled brute force. Cloud tiles are staged in shared memory; every thread
//                    keeps Q query points in registers and streams the tile through broadcast LDS. fp32 lanes are
//                    issued as packed f32x2 (two queries per FADD2/FMUL2/FFMA2); the exact (q-p)^2 form is kept
//                    (the expanded |q|^2-2qp+|p|^2 form loses ~1e-7*|p|^2 on d^2).
//   * ab_fd_kernel : from_sdf (vector_functions.py:130-139): np.gradient with unit spacing (2nd-order central inside,
//                    1st-order one-sided on the faces) + batch_normalize (vector_modification_functions.py:14-20).
#pragma once
#include "ab_interp.cuh"

namespace ab {

template <typename T>
struct NNParams {
  uint64_t n;
  T* out;
  const typename Vec4<T>::type* cloud;  // records (x, y, z, 0); dim 2 clouds have z = 0 and queries use z = 0
  uint32_t m;
  int32_t dim;
  int32_t grid_mode;
  GridK g;
  const void* co;
  uint64_t co_stride;
  int32_t co_is_f64;
};

// NN_TILE cloud points per shared-memory tile, double buffered (2 x 16 KB for both fp32/1024 and fp64/512)
template <typename T, int Q, int NT, int NN_TILE>
__global__ void __launch_bounds__(NT) ab_nn_kernel(const __grid_constant__ NNParams<T> kp) {
  typedef typename Vec4<T>::type V4;
  __shared__ V4 tile[2][NN_TILE];
  const uint64_t tile_pts = (uint64_t)NT * Q;
  const uint64_t n_tiles = (kp.n + tile_pts - 1) / tile_pts;
  for (uint64_t qt = blockIdx.x; qt < n_tiles; qt += gridDim.x) {
    const uint64_t idx = qt * tile_pts + (uint64_t)threadIdx.x * Q;
    T qx[Q], qy[Q], qz[Q], best[Q];
    if (kp.grid_mode) {
      uint64_t k = idx < kp.n ? idx : (kp.n - 1);
      uint32_t i0 = (uint32_t)(k / kp.g.plane);
      uint32_t rem = (uint32_t)(k - (uint64_t)i0 * kp.g.plane);
      uint32_t i1 = rem / kp.g.n2, i2 = rem - i1 * kp.g.n2;
      i0 += kp.g.i0_begin;
      i1 += kp.g.i1_begin;
#pragma unroll
      for (int j = 0; j < Q; j++) {
        if (kp.g.is2d) {  // index axes (1, nx, ny): parameter sets 1 and 2 are x and y
          qx[j] = grid_coord(kp.g, 1, i1, T());
          qy[j] = grid_coord(kp.g, 2, i2, T());
          qz[j] = T(0);
        } else {
          qx[j] = grid_coord(kp.g, 0, i0, T());
          qy[j] = grid_coord(kp.g, 1, i1, T());
          qz[j] = kp.dim == 3 ? grid_coord(kp.g, 2, i2, T()) : T(0);
        }
        if (++i2 == kp.g.n2) {
          i2 = 0;
          if (++i1 == kp.g.n1 + kp.g.i1_begin) {
            i1 = kp.g.i1_begin;
            ++i0;
          }
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < Q; j++) {
        uint64_t k = idx + j < kp.n ? idx + j : kp.n - 1;
        if (kp.co_is_f64) {
          const double* c = (const double*)kp.co;
          qx[j] = (T)c[k];
          qy[j] = (T)c[kp.co_stride + k];
          qz[j] = kp.dim == 3 ? (T)c[2 * kp.co_stride + k] : T(0);
        } else {
          const float* c = (const float*)kp.co;
          qx[j] = (T)c[k];
          qy[j] = (T)c[kp.co_stride + k];
          qz[j] = kp.dim == 3 ? (T)c[2 * kp.co_stride + k] : T(0);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < Q; j++) best[j] = T(3.0e38);

    const uint32_t n_ct = (kp.m + NN_TILE - 1) / NN_TILE;
    // prologue: stage tile 0
    for (int i = threadIdx.x; i < NN_TILE; i += NT) {
      uint32_t s = i < kp.m ? i : kp.m - 1;  // pad with a repeated real point: harmless for a min
      tile[0][i] = kp.cloud[s];
    }
    __syncthreads();
    for (uint32_t ct = 0; ct < n_ct; ct++) {
      const int cur = ct & 1;
      if (ct + 1 < n_ct) {  // stage the next tile while this one is consumed
        const uint32_t base = (ct + 1) * NN_TILE;
        for (int i = threadIdx.x; i < NN_TILE; i += NT) {
          uint32_t s = base + i < kp.m ? base + i : kp.m - 1;
          tile[cur ^ 1][i] = kp.cloud[s];
        }
      }
#pragma unroll 4
      for (int i = 0; i < NN_TILE; i++) {
        const V4 c = tile[cur][i];
#pragma unroll
        for (int j = 0; j < Q; j++) {
          T dx = qx[j] - c.x, dy = qy[j] - c.y, dz = qz[j] - c.z;
          best[j] = s_min(best[j], s_fma(dx, dx, s_fma(dy, dy, dz * dz)));
        }
      }
      __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < Q; j++)
      if (idx + j < kp.n) __stcs(kp.out + idx + j, s_sqrt(best[j]));
  }
}

// fp32 fast path. Two queries share every FADD2/FMUL2/FFMA2 (f32x2) and two cloud points share every FMNMX3, so one
// (query, point) pair costs 3.5 issue slots instead of 7: (3 FADD2 + FMUL2 + 2 FFMA2) per point per query pair, plus one
// 3-input min per query per two points. Tiles are staged already negated and duplicated, (-x,-x,-y,-y) + (-z,-z), so the
// packed adds take them straight from a broadcast LDS.128 + LDS.64 with no register shuffling.
AB_DEV float fmin3(float a, float b, float c) {
  float r;
  asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

template <int Q, int NT, int NN_TILE>
__global__ void __launch_bounds__(NT) ab_nn_kernel_f32x2(const __grid_constant__ NNParams<float> kp) {
  static_assert(Q % 2 == 0 && NN_TILE % 2 == 0, "pairs");
  __shared__ float4 sxy[2][NN_TILE];
  __shared__ float2 sz[2][NN_TILE];
  const uint64_t tile_pts = (uint64_t)NT * Q;
  const uint64_t n_tiles = (kp.n + tile_pts - 1) / tile_pts;
  for (uint64_t qt = blockIdx.x; qt < n_tiles; qt += gridDim.x) {
    const uint64_t idx = qt * tile_pts + (uint64_t)threadIdx.x * Q;
    float2 qx[Q / 2], qy[Q / 2], qz[Q / 2];
    float best[Q];
    {
      float x[Q], y[Q], z[Q];
      if (kp.grid_mode) {
        uint64_t k = idx < kp.n ? idx : (kp.n - 1);
        uint32_t i0 = (uint32_t)(k / kp.g.plane);
        uint32_t rem = (uint32_t)(k - (uint64_t)i0 * kp.g.plane);
        uint32_t i1 = rem / kp.g.n2, i2 = rem - i1 * kp.g.n2;
        i0 += kp.g.i0_begin;
        i1 += kp.g.i1_begin;
#pragma unroll
        for (int j = 0; j < Q; j++) {
          if (kp.g.is2d) {
            x[j] = grid_coord(kp.g, 1, i1, 0.0f);
            y[j] = grid_coord(kp.g, 2, i2, 0.0f);
            z[j] = 0.0f;
          } else {
            x[j] = grid_coord(kp.g, 0, i0, 0.0f);
            y[j] = grid_coord(kp.g, 1, i1, 0.0f);
            z[j] = kp.dim == 3 ? grid_coord(kp.g, 2, i2, 0.0f) : 0.0f;
          }
          if (++i2 == kp.g.n2) {
            i2 = 0;
            if (++i1 == kp.g.n1 + kp.g.i1_begin) {
              i1 = kp.g.i1_begin;
              ++i0;
            }
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < Q; j++) {
          const uint64_t k = idx + j < kp.n ? idx + j : kp.n - 1;
          if (kp.co_is_f64) {
            const double* c = (const double*)kp.co;
            x[j] = (float)c[k];
            y[j] = (float)c[kp.co_stride + k];
            z[j] = kp.dim == 3 ? (float)c[2 * kp.co_stride + k] : 0.0f;
          } else {
            const float* c = (const float*)kp.co;
            x[j] = c[k];
            y[j] = c[kp.co_stride + k];
            z[j] = kp.dim == 3 ? c[2 * kp.co_stride + k] : 0.0f;
          }
        }
      }
#pragma unroll
      for (int j = 0; j < Q / 2; j++) {
        qx[j] = make_float2(x[2 * j], x[2 * j + 1]);
        qy[j] = make_float2(y[2 * j], y[2 * j + 1]);
        qz[j] = make_float2(z[2 * j], z[2 * j + 1]);
      }
    }
#pragma unroll
    for (int j = 0; j < Q; j++) best[j] = 3.0e38f;

    const uint32_t n_ct = (kp.m + NN_TILE - 1) / NN_TILE;
    auto stage = [&](int buf, uint32_t base) {
      for (int i = threadIdx.x; i < NN_TILE; i += NT) {
        float4 c = make_float4(1.0e18f, 1.0e18f, 1.0e18f, 0.0f);  // padding: farther than any real point
        if (base + i < kp.m) c = kp.cloud[base + i];
        sxy[buf][i] = make_float4(-c.x, -c.x, -c.y, -c.y);
        sz[buf][i] = make_float2(-c.z, -c.z);
      }
    };
    __syncthreads();  // the previous query tile may still be reading the buffers
    stage(0, 0);
    __syncthreads();
    for (uint32_t ct = 0; ct < n_ct; ct++) {
      const int cur = ct & 1;
      if (ct + 1 < n_ct) stage(cur ^ 1, (ct + 1) * NN_TILE);  // overlaps with the scan of the current tile
#pragma unroll 2
      for (int i = 0; i < NN_TILE; i += 2) {
        const float4 a = sxy[cur][i], b = sxy[cur][i + 1];
        const float2 az = sz[cur][i], bz = sz[cur][i + 1];
#pragma unroll
        for (int j = 0; j < Q / 2; j++) {
          float2 dxa = __fadd2_rn(qx[j], make_float2(a.x, a.y)), dya = __fadd2_rn(qy[j], make_float2(a.z, a.w));
          float2 dza = __fadd2_rn(qz[j], az);
          float2 da = __ffma2_rn(dza, dza, __ffma2_rn(dya, dya, __fmul2_rn(dxa, dxa)));
          float2 dxb = __fadd2_rn(qx[j], make_float2(b.x, b.y)), dyb = __fadd2_rn(qy[j], make_float2(b.z, b.w));
          float2 dzb = __fadd2_rn(qz[j], bz);
          float2 db = __ffma2_rn(dzb, dzb, __ffma2_rn(dyb, dyb, __fmul2_rn(dxb, dxb)));
          best[2 * j] = fmin3(best[2 * j], da.x, db.x);
          best[2 * j + 1] = fmin3(best[2 * j + 1], da.y, db.y);
        }
      }
      __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < Q; j++)
      if (idx + j < kp.n) __stcs(kp.out + idx + j, s_sqrt(best[j]));
  }
}

// Few queries x many points (points mode with a short list, coarse grids): parallelism has to come from the cloud. Every
// warp owns QW queries (held by all lanes) and one slice of the cloud; lane l scans points l, l+32, ... of the slice, the
// 32 partial minima are combined with warp-shuffle min-reductions and lane 0 folds the result into the output with an
// atomic min on the float bit pattern (monotonic for non-negative values). Output must be pre-filled with 0x7f7f7f7f
// and is finalised (sqrt) by ab_nn_finalize_kernel.
template <int QW, int NT>
__global__ void __launch_bounds__(NT) ab_nn_split_kernel_f32(const __grid_constant__ NNParams<float> kp, uint32_t slice) {
  static_assert(QW % 2 == 0, "pairs");
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint64_t q0 = (uint64_t)blockIdx.x * QW;
  const uint32_t begin = (blockIdx.y * (NT / 32) + warp) * slice;
  if (begin >= kp.m) return;
  const uint32_t end = begin + slice < kp.m ? begin + slice : kp.m;
  float2 qx[QW / 2], qy[QW / 2], qz[QW / 2];
  float best[QW];
#pragma unroll
  for (int j = 0; j < QW; j++) {
    const uint64_t k = q0 + j < kp.n ? q0 + j : kp.n - 1;
    float x, y, z;
    if (kp.grid_mode) {
      uint32_t i0 = (uint32_t)(k / kp.g.plane);
      uint32_t rem = (uint32_t)(k - (uint64_t)i0 * kp.g.plane);
      uint32_t i1 = rem / kp.g.n2, i2 = rem - i1 * kp.g.n2;
      i0 += kp.g.i0_begin;
      i1 += kp.g.i1_begin;
      if (kp.g.is2d) {
        x = grid_coord(kp.g, 1, i1, 0.0f);
        y = grid_coord(kp.g, 2, i2, 0.0f);
        z = 0.0f;
      } else {
        x = grid_coord(kp.g, 0, i0, 0.0f);
        y = grid_coord(kp.g, 1, i1, 0.0f);
        z = kp.dim == 3 ? grid_coord(kp.g, 2, i2, 0.0f) : 0.0f;
      }
    } else if (kp.co_is_f64) {
      const double* c = (const double*)kp.co;
      x = (float)c[k];
      y = (float)c[kp.co_stride + k];
      z = kp.dim == 3 ? (float)c[2 * kp.co_stride + k] : 0.0f;
    } else {
      const float* c = (const float*)kp.co;
      x = c[k];
      y = c[kp.co_stride + k];
      z = kp.dim == 3 ? c[2 * kp.co_stride + k] : 0.0f;
    }
    if (j & 1) {
      qx[j / 2].y = x;
      qy[j / 2].y = y;
      qz[j / 2].y = z;
    } else {
      qx[j / 2].x = x;
      qy[j / 2].x = y;
      qz[j / 2].x = z;
    }
    best[j] = 3.0e38f;
  }
  for (uint32_t i = begin + lane; i < end; i += 32) {
    const float4 c = kp.cloud[i];  // coalesced: a warp reads 512 contiguous bytes
    const float2 nx = make_float2(-c.x, -c.x), ny = make_float2(-c.y, -c.y), nz = make_float2(-c.z, -c.z);
#pragma unroll
    for (int j = 0; j < QW / 2; j++) {
      const float2 dx = __fadd2_rn(qx[j], nx), dy = __fadd2_rn(qy[j], ny), dz = __fadd2_rn(qz[j], nz);
      const float2 d = __ffma2_rn(dz, dz, __ffma2_rn(dy, dy, __fmul2_rn(dx, dx)));
      best[2 * j] = fminf(best[2 * j], d.x);
      best[2 * j + 1] = fminf(best[2 * j + 1], d.y);
    }
  }
#pragma unroll
  for (int j = 0; j < QW; j++) {
    float b = best[j];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) b = fminf(b, __shfl_xor_sync(0xffffffffu, b, off));
    if (lane == 0 && q0 + j < kp.n) atomicMin(reinterpret_cast<unsigned int*>(kp.out) + q0 + j, __float_as_uint(b));
  }
}

__global__ void ab_nn_finalize_kernel(float* out, uint64_t n) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    out[i] = s_sqrt(out[i]);
}

// ---- from_sdf -----------------------------------------------------------------------------------------------------------
// The field is seen as a C-ordered (n0, n1, n2) array; a 2D (nx, ny) field arrives as (1, nx, ny) so that the fastest axis
// is always the long one. Thread = one (i1, i2) column position (i2 from blockIdx.x/threadIdx.x: coalesced; i1 =
// blockIdx.y), marching over a chunk of i0 planes (blockIdx.z) with the three i0-neighbours kept in registers, so every
// field sample is fetched from DRAM once and there is no integer division anywhere.
template <typename T>
struct FDParams {
  const T* field;        // samples [o0.., o1.., 0..) of the whole grid: buffer origin (o0, o1)
  uint32_t n0, n1, n2;   // whole-grid extents of the view
  uint32_t o0, o1;       // first i0 / i1 held by `field`
  uint32_t b0, e0, b1, e1;  // output range [b0, e0) x [b1, e1) x [0, n2)
  int32_t has0;          // 1: 3D (three components, row 0 = d/di0); 0: 2D view (rows = d/di1, d/di2)
  int32_t normalize;
  uint32_t chunk;        // i0 planes per CTA
  T* out;                // (dims, out_stride), indexed from the first output sample
  uint64_t out_stride;
};

// np.gradient, unit spacing, edge_order=1: interior (f[i+1] - f[i-1]) / 2, faces f[1] - f[0] and f[n-1] - f[n-2]. Written as
// (p - m) * s with the missing neighbour replaced by the centre sample and s = 1 on a face, 1/2 inside (bit-identical to
// the three-way form, no branches): the face tests depend on the thread's (i1, i2) only and leave the plane loop.
// 1 / |g| from |g|^2: fp32 one MUFU.RSQ (<= 2 ulp; the tolerance is 2e-6), fp64 IEEE sqrt + division (results match
// np.gradient / norm to 1e-13)
AB_DEV float fd_inv_norm(float m2) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(m2));
  return r;
}
AB_DEV double fd_inv_norm(double m2) { return 1.0 / sqrt(m2); }

// HAS0 / NORM: kp.has0 / kp.normalize as compile-time switches (the launcher picks the instantiation)
template <typename T, bool HAS0, bool NORM, int U = 4>
__global__ void __launch_bounds__(256) ab_fd_kernel(const __grid_constant__ FDParams<T> kp) {
  const uint32_t i2 = blockIdx.x * blockDim.x + threadIdx.x;
  if (i2 >= kp.n2) return;
  const uint32_t i1 = kp.b1 + blockIdx.y;
  const uint32_t x_begin = kp.b0 + blockIdx.z * kp.chunk;
  const uint32_t x_end = x_begin + kp.chunk < kp.e0 ? x_begin + kp.chunk : kp.e0;
  const uint32_t w1 = kp.e1 - kp.b1;                  // output rows per plane
  const int64_t fplane = (int64_t)(kp.n1 - kp.o1) * kp.n2;  // the buffer holds rows [o1, n1) of every plane
  // thread-constant face handling for the two in-plane axes: offsets of the neighbours (0 = use the centre) and scales
  const int64_t d1m = i1 > 0 ? -(int64_t)kp.n2 : 0, d1p = i1 + 1 < kp.n1 ? (int64_t)kp.n2 : 0;
  const int64_t d2m = i2 > 0 ? -1 : 0, d2p = i2 + 1 < kp.n2 ? 1 : 0;
  const T s1 = (d1m != 0 && d1p != 0) ? T(0.5) : T(1), s2 = (d2m != 0 && d2p != 0) ? T(0.5) : T(1);
  const T* __restrict__ pc = kp.field + (int64_t)(i1 - kp.o1) * kp.n2 + i2 + (int64_t)(x_begin - kp.o0) * fplane;  // plane x_begin
  const uint64_t oplane = (uint64_t)w1 * kp.n2;
  T* __restrict__ po = kp.out + ((uint64_t)(x_begin - kp.b0) * w1 + (i1 - kp.b1)) * kp.n2 + i2;
  // four planes per step with all their loads issued up front: the march is latency-bound otherwise (one DRAM load in
  // flight per thread); per plane the centre of the next plane + 4 in-plane neighbours = 20 independent loads per thread
  T fc = pc[0];
  T fm = (HAS0 && x_begin > 0) ? pc[-fplane] : fc;  // below the grid: the centre stands in (scale 1 there)
  for (uint32_t i0 = x_begin; i0 < x_end; i0 += U) {
    T c[U + 2];  // planes i0-1 .. i0+U
    c[0] = fm;
    c[1] = fc;
    T a1m[U], a1p[U], a2m[U], a2p[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const uint32_t x = i0 + u;
      const bool live = x < x_end;
      const T* p = live ? pc + (int64_t)u * fplane : pc;  // (dead planes of the last step re-read a valid address)
      c[u + 2] = (live && x + 1 < kp.n0) ? p[fplane] : T(0);
      a1m[u] = p[d1m];
      a1p[u] = p[d1p];
      a2m[u] = p[d2m];
      a2p[u] = p[d2p];
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      const uint32_t x = i0 + u;
      if (x >= x_end) break;
      const T cc = c[u + 1];
      T g0 = T(0);
      if (HAS0) {
        const bool lo = x == 0, hi = x + 1 >= kp.n0;
        g0 = ((hi ? cc : c[u + 2]) - (lo ? cc : c[u])) * ((lo || hi) ? T(1) : T(0.5));
      }
      T g1 = (a1p[u] - a1m[u]) * s1;
      T g2 = (a2p[u] - a2m[u]) * s2;
      if (NORM) {
        const T m2 = HAS0 ? s_fma(g0, g0, s_fma(g1, g1, g2 * g2)) : s_fma(g1, g1, g2 * g2);
        if (m2 > T(0)) {  // zero vectors stay zero (batch_normalize)
          const T im = fd_inv_norm(m2);
          g0 *= im;
          g1 *= im;
          g2 *= im;
        }
      }
      T* o = po + (uint64_t)u * oplane;
      if (HAS0) {
        __stcs(o, g0);
        __stcs(o + kp.out_stride, g1);
        __stcs(o + 2 * kp.out_stride, g2);
      } else {
        __stcs(o, g1);
        __stcs(o + kp.out_stride, g2);
      }
    }
    fm = c[U];
    fc = c[U + 1];
    pc += (int64_t)U * fplane;
    po += (uint64_t)U * oplane;
  }
}

}  // namespace ab
