// ab_nn_tree.cuh — exact nearest neighbour through an implicit octree (quadtree for 2D clouds).
//
// sdf_point_cloud_3d/2d (sdf_3D.py:283-286, sdf_2D.py:221-224) ask a k-d tree for the nearest cloud point of every query.
// The brute-force kernels of ab_kernels_aux.cuh run at the FMA-pipe limit but do n*m pair tests; this path does the
// reference's O(n log m) instead and returns THE SAME bits: the result is min over the cloud of one fixed d^2 expression
// (order independent), and a subtree is skipped only when the same expression evaluated on its box is already >= the
// running minimum (every operation in it is monotone in |dx|,|dy|,|dz|, and the box is widened by a slack that covers the
// rounding of the binning), so no point that could lower the minimum is ever skipped.
//
// Structure (rebuilt per call on the caller's stream, a few kernel launches over the cloud):
//   * root cube = bounding cube of the cloud; L levels; finest cells indexed by Morton code (x bit lowest);
//   * the cloud is counting-sorted by finest cell into `pts`, `start[code]` = first point of the cell (cells+1 entries);
//   * node (level l, code c) owns pts[start[c << B(L-l)] .. start[(c+1) << B(L-l)]) with B = bits per level (3 or 2), so
//     the one table serves every level and there are no node records at all.
//   * occ[] holds one byte per inner node: which of its children are non-empty.
// Query: one thread per query, depth-first, children visited nearest-octant first (child = k XOR preferred); the box bound
// of a child is assembled from six per-node half distances; the children still to visit at every level live in one
// 64-bit register used as a stack; ranges of <= leaf points are scanned directly.
#pragma once
#include "ab_kernels_aux.cuh"

namespace ab {

template <typename T>
struct TreeParams {
  NNParams<T> q;  // queries + output (q.cloud = the caller's unsorted cloud)
  const TreeRef<T>* tree;
};

// ---- order-preserving encoding of doubles for atomicMin/Max ------------------------------------------------------------
AB_DEV unsigned long long ord_encode(double d) {
  long long b = __double_as_longlong(d);
  return b < 0 ? ~(unsigned long long)b : ((unsigned long long)b | 0x8000000000000000ull);
}
AB_DEV double ord_decode(unsigned long long u) {
  return __longlong_as_double((u & 0x8000000000000000ull) ? (long long)(u & 0x7fffffffffffffffull) : (long long)~u);
}

// bbox[0..2] = min, bbox[3..5] = max (encoded); initialise to ~0 / 0. One atomic per axis and CTA.
template <typename T>
__global__ void __launch_bounds__(256) ab_tree_bbox_kernel(const typename Vec4<T>::type* __restrict__ cloud, uint32_t m,
                                                           unsigned long long* bbox) {
  T lo[3] = {T(3.0e38), T(3.0e38), T(3.0e38)}, hi[3] = {T(-3.0e38), T(-3.0e38), T(-3.0e38)};
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
    const auto c = cloud[i];
    lo[0] = s_min(lo[0], c.x); hi[0] = s_max(hi[0], c.x);
    lo[1] = s_min(lo[1], c.y); hi[1] = s_max(hi[1], c.y);
    lo[2] = s_min(lo[2], c.z); hi[2] = s_max(hi[2], c.z);
  }
  __shared__ T slo[8][3], shi[8][3];
#pragma unroll
  for (int a = 0; a < 3; a++) {
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      lo[a] = s_min(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
      hi[a] = s_max(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
    }
    if ((threadIdx.x & 31) == 0) {
      slo[threadIdx.x >> 5][a] = lo[a];
      shi[threadIdx.x >> 5][a] = hi[a];
    }
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    const int a = threadIdx.x;
    T l = slo[0][a], h = shi[0][a];
    for (int w = 1; w < 8; w++) {
      l = s_min(l, slo[w][a]);
      h = s_max(h, shi[w][a]);
    }
    atomicMin(&bbox[a], ord_encode((double)l));
    atomicMax(&bbox[3 + a], ord_encode((double)h));
  }
}

template <typename T>
__global__ void ab_tree_geom_kernel(const unsigned long long* bbox, int levels, uint32_t leaf, const typename Vec4<T>::type* pts,
                                    const uint32_t* start, const uint8_t* occ, TreeRef<T>* ref) {
  double lo[3], hi[3], ext = 0.0, mag = 0.0;
  for (int a = 0; a < 3; a++) {
    lo[a] = ord_decode(bbox[a]);
    hi[a] = ord_decode(bbox[3 + a]);
    ext = fmax(ext, hi[a] - lo[a]);
    mag = fmax(mag, fmax(fabs(lo[a]), fabs(hi[a])));
  }
  const double eps = sizeof(T) == 4 ? 1.1920928955078125e-7 : 2.220446049250313e-16;
  ext = fmax(ext * (1.0 + 1.0e-6), fmax(mag * 64.0 * eps, 1.0e-30));  // never zero (single point / coincident points)
  const double cells = (double)(1u << levels);
  TreeGeom<T> g;
  for (int a = 0; a < 3; a++) g.org[a] = (T)lo[a];
  g.cell = (T)(ext / cells);
  g.inv_cell = (T)(cells / ext);
  g.slack = (T)(ext * 9.5367431640625e-7 + 8.0 * eps * (mag + ext));
  ref->geom = g;
  ref->pts = pts;
  ref->start = start;
  ref->occ = occ;
  ref->levels = levels;
  ref->leaf = leaf;
}

// bit spreading for Morton codes: x bit i -> bit B*i
AB_DEV uint32_t spread3(uint32_t v) {  // 8 bits -> 24
  v = (v | (v << 16)) & 0x030000ffu;
  v = (v | (v << 8)) & 0x0300f00fu;
  v = (v | (v << 4)) & 0x030c30c3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}
AB_DEV uint32_t spread2(uint32_t v) {  // 12 bits -> 24
  v = (v | (v << 8)) & 0x00ff00ffu;
  v = (v | (v << 4)) & 0x0f0f0f0fu;
  v = (v | (v << 2)) & 0x33333333u;
  v = (v | (v << 1)) & 0x55555555u;
  return v;
}
template <typename T>
AB_DEV uint32_t cell_index(T x, T org, T inv_cell, uint32_t last) {
  T f = (x - org) * inv_cell;
  int i = (int)f;  // truncation; negatives clamp to 0 below
  i = i < 0 ? 0 : i;
  return (uint32_t)i > last ? last : (uint32_t)i;
}
template <typename T, int DIM>
AB_DEV uint32_t morton_of(T x, T y, T z, const TreeGeom<T>& g, int levels) {
  const uint32_t last = (1u << levels) - 1u;
  const uint32_t ix = cell_index(x, g.org[0], g.inv_cell, last), iy = cell_index(y, g.org[1], g.inv_cell, last);
  if constexpr (DIM == 3) {
    const uint32_t iz = cell_index(z, g.org[2], g.inv_cell, last);
    return spread3(ix) | (spread3(iy) << 1) | (spread3(iz) << 2);
  } else {
    return spread2(ix) | (spread2(iy) << 1);
  }
}

// counts[code]++ and remember the arrival rank inside the cell (it becomes the offset of the scatter)
template <typename T, int DIM>
__global__ void ab_tree_count_kernel(const typename Vec4<T>::type* __restrict__ cloud, uint32_t m, const TreeRef<T>* ref,
                                     int levels, uint32_t* counts, uint32_t* __restrict__ key, uint32_t* __restrict__ rank) {
  const TreeGeom<T> g = ref->geom;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
    const auto c = cloud[i];
    const uint32_t k = morton_of<T, DIM>(c.x, c.y, c.z, g, levels);
    key[i] = k;
    rank[i] = atomicAdd(&counts[k], 1u);
  }
}

template <typename T>
__global__ void ab_tree_scatter_kernel(const typename Vec4<T>::type* __restrict__ cloud, uint32_t m,
                                       const uint32_t* __restrict__ start, const uint32_t* __restrict__ key,
                                       const uint32_t* __restrict__ rank, typename Vec4<T>::type* __restrict__ pts) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x)
    pts[start[key[i]] + rank[i]] = cloud[i];
}

// ---- exclusive prefix sum over uint32 (three launches: chunk totals, scan of the totals, apply) ---------------------------
constexpr int kScanNT = 1024, kScanIPT = 8, kScanChunk = kScanNT * kScanIPT;

AB_DEV uint32_t block_exclusive_scan(uint32_t v, uint32_t* total) {  // kScanNT threads
  __shared__ uint32_t warp_tot[32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_tot[w] = inc;
  __syncthreads();
  if (w == 0) {
    uint32_t t = warp_tot[lane], ti = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t u = __shfl_up_sync(0xffffffffu, ti, o);
      if (lane >= o) ti += u;
    }
    warp_tot[lane] = ti - t;  // exclusive
    if (lane == 31 && total) *total = ti;
  }
  __syncthreads();
  const uint32_t r = warp_tot[w] + inc - v;
  __syncthreads();
  return r;
}

// data is padded to a multiple of kScanChunk
__global__ void __launch_bounds__(kScanNT) ab_scan_totals_kernel(const uint32_t* __restrict__ data, uint32_t* __restrict__ totals) {
  const uint4* p = reinterpret_cast<const uint4*>(data + (size_t)blockIdx.x * kScanChunk + threadIdx.x * kScanIPT);
  const uint4 a = p[0], b = p[1];
  uint32_t s = a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w;
  __shared__ uint32_t tot;
  block_exclusive_scan(s, &tot);
  if (threadIdx.x == 0) totals[blockIdx.x] = tot;
}
__global__ void __launch_bounds__(kScanNT) ab_scan_offsets_kernel(uint32_t* totals, uint32_t nb) {
  const uint32_t per = (nb + kScanNT - 1) / kScanNT;
  const uint32_t b = threadIdx.x * per, e = b + per < nb ? b + per : nb;
  uint32_t s = 0;
  for (uint32_t i = b; i < e; i++) s += totals[i];
  uint32_t run = block_exclusive_scan(s, nullptr);
  for (uint32_t i = b; i < e; i++) {
    const uint32_t t = totals[i];
    totals[i] = run;
    run += t;
  }
}
__global__ void __launch_bounds__(kScanNT) ab_scan_apply_kernel(uint32_t* data, const uint32_t* __restrict__ offsets) {
  uint4* p = reinterpret_cast<uint4*>(data + (size_t)blockIdx.x * kScanChunk + threadIdx.x * kScanIPT);
  uint4 a = p[0], b = p[1];
  const uint32_t s = a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w;
  uint32_t run = offsets[blockIdx.x] + block_exclusive_scan(s, nullptr);
  uint32_t t;
  t = a.x; a.x = run; run += t;
  t = a.y; a.y = run; run += t;
  t = a.z; a.z = run; run += t;
  t = a.w; a.w = run; run += t;
  t = b.x; b.x = run; run += t;
  t = b.y; b.y = run; run += t;
  t = b.z; b.z = run; run += t;
  t = b.w; b.w = run;
  p[0] = a;
  p[1] = b;
}

// ---- the query kernel ---------------------------------------------------------------------------------------------------------
template <typename T>
AB_DEV void nn_query_point(const NNParams<T>& kp, uint64_t k, T& x, T& y, T& z) {
  if (kp.grid_mode) {
    uint32_t i0 = (uint32_t)(k / kp.g.plane);
    uint32_t rem = (uint32_t)(k - (uint64_t)i0 * kp.g.plane);
    uint32_t i1 = rem / kp.g.n2, i2 = rem - i1 * kp.g.n2;
    i0 += kp.g.i0_begin;
    i1 += kp.g.i1_begin;
    if (kp.g.is2d) {
      x = grid_coord(kp.g, 1, i1, T());
      y = grid_coord(kp.g, 2, i2, T());
      z = T(0);
    } else {
      x = grid_coord(kp.g, 0, i0, T());
      y = grid_coord(kp.g, 1, i1, T());
      z = kp.dim == 3 ? grid_coord(kp.g, 2, i2, T()) : T(0);
    }
  } else if (kp.co_is_f64) {
    const double* c = (const double*)kp.co;
    x = (T)c[k];
    y = (T)c[kp.co_stride + k];
    z = kp.dim == 3 ? (T)c[2 * kp.co_stride + k] : T(0);
  } else {
    const float* c = (const float*)kp.co;
    x = (T)c[k];
    y = (T)c[kp.co_stride + k];
    z = kp.dim == 3 ? (T)c[2 * kp.co_stride + k] : T(0);
  }
}

// occ[level_offset(l) + code] = which children of node (l, code) hold points (bit c = child c), for l < L
template <int DIM>
__global__ void ab_tree_occupancy_kernel(const uint32_t* __restrict__ start, int levels, uint8_t* __restrict__ occ) {
  constexpr uint32_t NC = 1u << DIM;
  const uint32_t total = level_offset<DIM>(levels);
  for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    int l = 0;
    while (level_offset<DIM>(l + 1) <= t) l++;
    const uint32_t code = t - level_offset<DIM>(l);
    const int sh = DIM * (levels - l - 1);
    uint32_t mask = 0;
    uint32_t prev = start[(size_t)(code * NC) << sh];
#pragma unroll
    for (uint32_t c = 0; c < NC; c++) {
      const uint32_t next = start[(size_t)(code * NC + c + 1) << sh];
      mask |= (next != prev ? 1u : 0u) << c;
      prev = next;
    }
    occ[t] = (uint8_t)mask;
  }
}

// Point-list queries (no spatial order to share a walk): one thread per query, ab_tree.cuh's walk
template <typename T, int DIM, int NT>
__global__ void __launch_bounds__(NT) ab_nn_tree_kernel(const __grid_constant__ TreeParams<T> tp) {
  const uint64_t k = (uint64_t)blockIdx.x * NT + threadIdx.x;
  if (k >= tp.q.n) return;
  const TreeRef<T> t = *tp.tree;
  T qx, qy, qz, best;
  uint32_t bi;
  nn_query_point(tp.q, k, qx, qy, qz);
  tree_nearest<T, DIM, 1, false>(t, qx, qy, qz, best, bi);
  __stcs(tp.q.out + k, s_sqrt(best));
}

// Grid queries: packet traversal. A warp owns a compact 2x4x4 (1x4x8 on 2D grids) block of samples and walks the tree
// ONCE for all of them: node state, child order (taken from lane 0's sample) and the stack are warp-uniform, every lane
// tests its own sample against the child's box and the child is entered when any lane still needs it (one vote). Leaf
// points are fetched once per warp (same address in every lane) and tested by all 32 samples. The lanes of such a block
// need almost the same nodes, so the votes waste little and nothing diverges.
template <typename T, int DIM, int NT>
__global__ void __launch_bounds__(NT) ab_nn_tree_packet_kernel(const __grid_constant__ TreeParams<T> tp) {
  const TreeRef<T> tr = *tp.tree;

  const GridK& gk = tp.q.g;
  const uint32_t n0 = (uint32_t)(tp.q.n / gk.plane), n1 = gk.n1, n2 = gk.n2;
  const uint32_t lane = threadIdx.x & 31u;
  uint32_t b0, b1, b2, l0, l1, l2;
  if (n0 == 1) { b0 = 1; b1 = 4; b2 = 8; l0 = 0; l1 = lane >> 3; l2 = lane & 7u; }
  else { b0 = 2; b1 = 4; b2 = 4; l0 = lane >> 4; l1 = (lane >> 2) & 3u; l2 = lane & 3u; }
  const uint32_t t1 = (n1 + b1 - 1) / b1, t2 = (n2 + b2 - 1) / b2;
  const uint64_t w = ((uint64_t)blockIdx.x * NT + threadIdx.x) >> 5;
  const uint32_t w2 = (uint32_t)(w % t2), w1 = (uint32_t)((w / t2) % t1);
  const uint64_t w0 = w / ((uint64_t)t1 * t2);
  if (w0 * b0 >= n0) return;  // whole warp out of range (uniform)
  uint64_t i0 = w0 * b0 + l0;
  uint32_t i1 = w1 * b1 + l1, i2 = w2 * b2 + l2;
  const bool valid = i0 < n0 && i1 < n1 && i2 < n2;
  // lanes hanging over the edge of the grid repeat an in-range sample: they must stay in the votes
  i0 = i0 < n0 ? i0 : n0 - 1;
  i1 = i1 < n1 ? i1 : n1 - 1;
  i2 = i2 < n2 ? i2 : n2 - 1;
  const uint64_t k = (i0 * n1 + i1) * n2 + i2;
  T qx, qy, qz, best;
  uint32_t bi;
  nn_query_point(tp.q, k, qx, qy, qz);
  tree_nearest_packet<T, DIM, 1, false>(tr, qx, qy, qz, best, bi);
  if (valid) __stcs(tp.q.out + k, s_sqrt(best));
}

}  // namespace ab
