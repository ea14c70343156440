// ab_tree.cuh — the implicit octree / quadtree over a point cloud, device side: the structure and the per-thread exact
// nearest-neighbour walk. Shared by the dedicated kernels of ab_nn_tree.cuh and by the interpreter's point-cloud leaf
// (a cloud may sit anywhere inside a tree, where queries arrive already warped by the ops above it).
// See ab_nn_tree.cuh for the construction and for why pruning never changes the result.
#pragma once

namespace ab {

template <typename T>
struct TreeGeom {
  T org[3];
  T cell;      // finest cell edge
  T inv_cell;
  T slack;
};

// device-resident descriptor (filled in by ab_tree_geom_kernel); start == nullptr: no tree
template <typename T>
struct TreeRef {
  const typename Vec4<T>::type* pts;  // the cloud in cell order
  const uint32_t* start;              // first point of every finest cell (cells + 1 entries)
  const uint8_t* occ;                 // child occupancy of every inner node, level by level
  TreeGeom<T> geom;
  int32_t levels;
  uint32_t leaf;  // ranges of at most this many points are scanned instead of subdivided
};

AB_DEV float max3_(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
AB_DEV double max3_(double a, double b, double c) { return fmax(fmax(a, b), c); }

// first node of level l in the occupancy table: (NC^l - 1) / (NC - 1)
template <int DIM>
AB_DEV constexpr uint32_t level_offset(int l) {
  return ((1u << (DIM * l)) - 1u) / ((1u << DIM) - 1u);
}

// bit k of the result = bit (k XOR pref) of m: the children in the order they are visited
template <int DIM>
AB_DEV uint32_t xor_permute(uint32_t m, uint32_t pref) {
  if (pref & 1u) m = ((m & 0x55u) << 1) | ((m & 0xaau) >> 1);
  if (pref & 2u) m = ((m & 0x33u) << 2) | ((m & 0xccu) >> 2);
  if constexpr (DIM == 3)
    if (pref & 4u) m = ((m & 0x0fu) << 4) | ((m & 0xf0u) >> 4);
  return m;
}

// per-axis distances from r to the low and the high half of the node [mid - cs, mid + cs), mid = (2i+1)*cs, both widened
// by the slack (css = cs + slack): with t = r - mid, low half: max(t, -t - cs), high half: max(-t, t - cs)
template <typename T>
AB_DEV void half_distances(T r, uint32_t i, T cs, T css, T slack, T& d0, T& d1, uint32_t& high) {
  const T t = s_fma(-(T)(2 * i + 1), cs, r);
  d0 = max3_(t - slack, -t - css, T(0));
  d1 = max3_(-t - slack, t - css, T(0));
  high = t >= T(0) ? 1u : 0u;
}

// The squared-distance expression. FORM 1 = the fp32 brute-force kernel's nesting (ab_nn_kernel_f32x2), FORM 0 = the
// generic kernel's and the interpreter's. The same expression bounds the boxes, so pruning is exact for either.
template <int FORM>
AB_DEV float tree_d2(float dx, float dy, float dz) {
  if constexpr (FORM == 1) return __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
  else return __fmaf_rn(dx, dx, __fmaf_rn(dy, dy, __fmul_rn(dz, dz)));
}
template <int FORM>
AB_DEV double tree_d2(double dx, double dy, double dz) {
  return __fma_rn(dx, dx, __fma_rn(dy, dy, __dmul_rn(dz, dz)));
}

// One query, depth first, children nearest-octant first (child = k XOR preferred); the children still to visit at every
// level live in one 64-bit register used as a stack. TRACK: also return the position of the winner in t.pts.
template <typename T, int DIM, int FORM, bool TRACK>
AB_DEV void tree_nearest(const TreeRef<T>& t, T qx, T qy, T qz, T& best, uint32_t& best_i) {
  typedef typename Vec4<T>::type V4;
  constexpr int B = DIM;            // Morton bits per level
  constexpr uint32_t NC = 1u << B;  // children per node = bits of one stack entry
  const int L = t.levels;
  const uint32_t* __restrict__ start = t.start;
  const uint8_t* __restrict__ occ = t.occ;
  const V4* __restrict__ pts = t.pts;
  // work relative to the cube's corner
  const T rx = qx - t.geom.org[0], ry = qy - t.geom.org[1], rz = DIM == 3 ? qz - t.geom.org[2] : T(0);
  // rounding of rx/ry/rz grows with the query's distance from the corner: widen the boxes accordingly
  const T slack = t.geom.slack + T(sizeof(T) == 4 ? 4.8e-7 : 8.9e-16) * (s_abs(rx) + s_abs(ry) + s_abs(rz));
  best = T(3.0e38);
  best_i = 0;
  uint32_t ix = 0, iy = 0, iz = 0, code = 0;
  uint64_t stack = 0;
  int l = 0;             // level of the current node; its children live on level l+1
  uint32_t todo = 0;     // children still to visit, in visiting order (bit k = child k XOR pref)
  bool fresh = true;     // just descended: fetch the occupancy
  while (true) {
    // per-node values (recomputed after coming back up: cheaper than keeping seven registers per level)
    const T cs = t.geom.cell * (T)(1u << (L - l - 1));  // child edge
    const T css = cs + slack;
    T ax0, ax1, ay0, ay1, az0 = T(0), az1 = T(0);
    uint32_t hx, hy, hz = 0;
    half_distances(rx, ix, cs, css, slack, ax0, ax1, hx);
    half_distances(ry, iy, cs, css, slack, ay0, ay1, hy);
    if constexpr (DIM == 3) half_distances(rz, iz, cs, css, slack, az0, az1, hz);
    const uint32_t pref = hx | (hy << 1) | (hz << 2);
    if (fresh) todo = xor_permute<DIM>(occ[level_offset<DIM>(l) + code], pref);
    bool descended = false;
    while (todo) {
      const uint32_t c = (uint32_t)(__ffs((int)todo) - 1) ^ pref;
      todo &= todo - 1u;
      const T bx = (c & 1u) ? ax1 : ax0, by = (c & 2u) ? ay1 : ay0, bz = (c & 4u) ? az1 : az0;
      if (tree_d2<FORM>(bx, by, bz) >= best) continue;
      const uint32_t ccode = (code << B) | c;
      const int shift = B * (L - l - 1);
      const uint32_t s = start[(size_t)ccode << shift], e = start[(size_t)(ccode + 1) << shift];
      if (shift == 0 || e - s <= t.leaf) {
        for (uint32_t i = s; i < e; i++) {
          const V4 p = pts[i];
          const T dx = qx - p.x, dy = qy - p.y, dz = DIM == 3 ? qz - p.z : T(0);
          const T d2 = tree_d2<FORM>(dx, dy, dz);
          if constexpr (TRACK) {
            if (d2 < best) {
              best = d2;
              best_i = i;
            }
          } else {
            best = s_min(best, d2);
          }
        }
        continue;
      }
      stack = (stack << NC) | todo;
      l++;
      ix = 2 * ix + (c & 1u);
      iy = 2 * iy + ((c >> 1) & 1u);
      if constexpr (DIM == 3) iz = 2 * iz + (c >> 2);
      code = ccode;
      descended = true;
      break;
    }
    if (descended) {
      fresh = true;
      continue;
    }
    if (l == 0) break;
    l--;
    ix >>= 1;
    iy >>= 1;
    iz >>= 1;
    code >>= B;
    todo = (uint32_t)(stack & (uint64_t)((1u << NC) - 1u));
    stack >>= NC;
    fresh = false;
  }
}

// The same walk for a whole warp at once (all 32 lanes must call it together): node state, child order (from lane 0's
// query) and the stack are warp-uniform, every lane tests its own query against the child's box and a child is entered when
// any lane still needs it; leaf points are fetched once per warp. For queries that are close together (grid samples) the
// lanes need nearly the same nodes, so little is wasted and nothing diverges.
template <typename T, int DIM, int FORM, bool TRACK>
AB_DEV void tree_nearest_packet(const TreeRef<T>& t, T qx, T qy, T qz, T& best, uint32_t& best_i) {
  typedef typename Vec4<T>::type V4;
  constexpr int B = DIM;
  constexpr uint32_t NC = 1u << B;
  constexpr uint32_t kFull = 0xffffffffu;
  const int L = t.levels;
  const uint32_t* __restrict__ start = t.start;
  const uint8_t* __restrict__ occ = t.occ;
  const V4* __restrict__ pts = t.pts;
  const T rx = qx - t.geom.org[0], ry = qy - t.geom.org[1], rz = DIM == 3 ? qz - t.geom.org[2] : T(0);
  const T slack = t.geom.slack + T(sizeof(T) == 4 ? 4.8e-7 : 8.9e-16) * (s_abs(rx) + s_abs(ry) + s_abs(rz));
  const T ux = __shfl_sync(kFull, rx, 0), uy = __shfl_sync(kFull, ry, 0), uz = __shfl_sync(kFull, rz, 0);
  best = T(3.0e38);
  best_i = 0;
  uint32_t ix = 0, iy = 0, iz = 0, code = 0;
  uint64_t stack = 0;
  int l = 0;
  uint32_t todo = 0;
  bool fresh = true;
  while (true) {
    const T cs = t.geom.cell * (T)(1u << (L - l - 1));
    const T css = cs + slack;
    T ax0, ax1, ay0, ay1, az0 = T(0), az1 = T(0);
    uint32_t hx, hy, hz = 0;
    half_distances(rx, ix, cs, css, slack, ax0, ax1, hx);
    half_distances(ry, iy, cs, css, slack, ay0, ay1, hy);
    if constexpr (DIM == 3) half_distances(rz, iz, cs, css, slack, az0, az1, hz);
    (void)hx, (void)hy, (void)hz;
    uint32_t pref = (ux >= (T)(2 * ix + 1) * cs ? 1u : 0u) | (uy >= (T)(2 * iy + 1) * cs ? 2u : 0u);
    if constexpr (DIM == 3) pref |= uz >= (T)(2 * iz + 1) * cs ? 4u : 0u;
    if (fresh) todo = xor_permute<DIM>(occ[level_offset<DIM>(l) + code], pref);
    bool descended = false;
    while (todo) {
      const uint32_t c = (uint32_t)(__ffs((int)todo) - 1) ^ pref;
      todo &= todo - 1u;
      const T bx = (c & 1u) ? ax1 : ax0, by = (c & 2u) ? ay1 : ay0, bz = (c & 4u) ? az1 : az0;
      if (!__any_sync(kFull, tree_d2<FORM>(bx, by, bz) < best)) continue;
      const uint32_t ccode = (code << B) | c;
      const int shift = B * (L - l - 1);
      const uint32_t s = start[(size_t)ccode << shift], e = start[(size_t)(ccode + 1) << shift];
      if (shift == 0 || e - s <= t.leaf) {
        for (uint32_t i = s; i < e; i++) {
          const V4 p = pts[i];
          const T dx = qx - p.x, dy = qy - p.y, dz = DIM == 3 ? qz - p.z : T(0);
          const T d2 = tree_d2<FORM>(dx, dy, dz);
          if constexpr (TRACK) {
            if (d2 < best) {
              best = d2;
              best_i = i;
            }
          } else {
            best = s_min(best, d2);
          }
        }
        continue;
      }
      stack = (stack << NC) | todo;
      l++;
      ix = 2 * ix + (c & 1u);
      iy = 2 * iy + ((c >> 1) & 1u);
      if constexpr (DIM == 3) iz = 2 * iz + (c >> 2);
      code = ccode;
      descended = true;
      break;
    }
    if (descended) {
      fresh = true;
      continue;
    }
    if (l == 0) break;
    l--;
    ix >>= 1;
    iy >>= 1;
    iz >>= 1;
    code >>= B;
    todo = (uint32_t)(stack & (uint64_t)((1u << NC) - 1u));
    stack >>= NC;
    fresh = false;
  }
}

}  // namespace ab
