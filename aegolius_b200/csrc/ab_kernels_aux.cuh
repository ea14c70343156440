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
template <typename T>
struct FDParams {
  const T* field;        // planes [plane0, ...) of the whole grid
  uint32_t plane0;
  uint32_t res[3];
  uint32_t slab_begin, slab_end;
  int32_t dims;          // 2 or 3
  int32_t normalize;
  T* out;                // (dims, out_stride), indexed from the slab's first point
  uint64_t out_stride;
};

template <typename T>
AB_DEV T fd_axis(const T* f, uint64_t k, uint64_t stride, uint32_t i, uint32_t n) {
  // np.gradient, unit spacing, edge_order=1
  if (n < 2) return T(0);
  if (i == 0) return f[k + stride] - f[k];
  if (i == n - 1) return f[k] - f[k - stride];
  return (f[k + stride] - f[k - stride]) * T(0.5);
}

template <typename T, int NT>
__global__ void __launch_bounds__(NT) ab_fd_kernel(const __grid_constant__ FDParams<T> kp) {
  const uint32_t n0 = kp.res[0], n1 = kp.res[1], n2 = kp.dims == 3 ? kp.res[2] : 1;
  const uint64_t plane = (uint64_t)n1 * n2;
  const uint64_t n = (uint64_t)(kp.slab_end - kp.slab_begin) * plane;
  for (uint64_t l = (uint64_t)blockIdx.x * NT + threadIdx.x; l < n; l += (uint64_t)gridDim.x * NT) {
    uint32_t i0 = (uint32_t)(l / plane);
    uint32_t rem = (uint32_t)(l - (uint64_t)i0 * plane);
    uint32_t i1 = rem / n2, i2 = rem - i1 * n2;
    i0 += kp.slab_begin;
    const uint64_t k = (uint64_t)(i0 - kp.plane0) * plane + rem;  // index into the field buffer
    T g0 = fd_axis(kp.field, k, plane, i0, n0);
    T g1 = fd_axis(kp.field, k, (uint64_t)n2, i1, n1);
    T g2 = kp.dims == 3 ? fd_axis(kp.field, k, (uint64_t)1, i2, n2) : T(0);
    if (kp.normalize) {
      T m = s_sqrt(s_fma(g0, g0, s_fma(g1, g1, g2 * g2)));
      if (m != T(0)) {
        T im = T(1) / m;
        g0 *= im;
        g1 *= im;
        g2 *= im;
      }
    }
    __stcs(kp.out + l, g0);
    __stcs(kp.out + kp.out_stride + l, g1);
    if (kp.dims == 3) __stcs(kp.out + 2 * kp.out_stride + l, g2);
  }
}

}  // namespace ab
