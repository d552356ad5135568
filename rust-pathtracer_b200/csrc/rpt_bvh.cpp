// rpt_bvh.cpp — see rpt_bvh.h. Host only; runs once per scene at rpt_scene_create.
#include "rpt_bvh.h"

#include <algorithm>
#include <cmath>
#include <limits>

namespace rpt {
namespace {

constexpr float kInf = std::numeric_limits<float>::infinity();
constexpr int kBuckets = 6;  // bvh.rs:381

inline Box empty_box() { return Box{{kInf, kInf, kInf}, {-kInf, -kInf, -kInf}}; }
inline void merge(Box &a, const Box &b) {
  for (int k = 0; k < 3; ++k) {
    a.mn[k] = std::fmin(a.mn[k], b.mn[k]);
    a.mx[k] = std::fmax(a.mx[k], b.mx[k]);
  }
}
inline void centroid(const Box &b, float c[3]) {  // AABB::center = min + size / 2 (aabb.rs:95-97)
  for (int k = 0; k < 3; ++k) c[k] = b.mn[k] + (b.mx[k] - b.mn[k]) / 2.0f;
}
inline float area(const Box &b) {  // aabb.rs:99-102; an empty box yields +inf (and 0 * inf = NaN below)
  float sx = b.mx[0] - b.mn[0], sy = b.mx[1] - b.mn[1], sz = b.mx[2] - b.mn[2];
  return 2.0f * (sx * sy + sx * sz + sy * sz);
}

struct Builder {
  const std::vector<Box> &shapes;
  BuiltBvh out;
  uint32_t next_order = 0;

  explicit Builder(const std::vector<Box> &s) : shapes(s) { out.order.assign(s.size(), 0); }

  Box bounds_of(const uint32_t *idx, size_t n) const {
    Box b = empty_box();
    for (size_t i = 0; i < n; ++i) merge(b, shapes[idx[i]]);
    return b;
  }

  // Returns the child-ref of the subtree over idx[0..n). Leaves are numbered in DFS (left first)
  // order, which is exactly the order FlatBVH::traverse reports candidates in.
  int32_t build(std::vector<uint32_t> idx, uint32_t depth) {
    out.max_depth = std::max(out.max_depth, depth);
    if (idx.size() == 1) {
      out.order[idx[0]] = next_order++;
      return ~(int32_t)idx[0];
    }
    Box cb = empty_box();
    for (uint32_t i : idx) {
      float c[3];
      centroid(shapes[i], c);
      for (int k = 0; k < 3; ++k) {
        cb.mn[k] = std::fmin(cb.mn[k], c[k]);
        cb.mx[k] = std::fmax(cb.mx[k], c[k]);
      }
    }
    float ext[3] = {cb.mx[0] - cb.mn[0], cb.mx[1] - cb.mn[1], cb.mx[2] - cb.mn[2]};
    float widest = std::fmax(std::fmax(ext[0], ext[1]), ext[2]);
    int axis = -1;  // highest-index axis whose extent equals the maximum (bvh.rs:346-351)
    if (widest > 0.0f)
      for (int k = 0; k < 3; ++k)
        if (ext[k] >= widest) axis = k;

    std::vector<uint32_t> left, right;
    if (axis < 0 || ext[axis] < 0.00001f) {  // bvh.rs:359-377
      size_t half = idx.size() / 2;
      left.assign(idx.begin(), idx.begin() + half);
      right.assign(idx.begin() + half, idx.end());
    } else {
      Box whole = bounds_of(idx.data(), idx.size());
      std::vector<uint32_t> bins[kBuckets];
      Box bin_box[kBuckets];
      size_t bin_n[kBuckets] = {0, 0, 0, 0, 0, 0};
      for (auto &b : bin_box) b = empty_box();
      for (uint32_t i : idx) {
        float c[3];
        centroid(shapes[i], c);
        float rel = (c[axis] - cb.mn[axis]) / ext[axis];
        size_t b = (size_t)(rel * ((float)kBuckets - 0.01f));  // bvh.rs:398
        if (b >= (size_t)kBuckets) b = kBuckets - 1;
        bins[b].push_back(i);
        merge(bin_box[b], shapes[i]);
        bin_n[b] += 1;
      }
      int best = 0;
      float best_cost = kInf;
      for (int s = 0; s < kBuckets - 1; ++s) {  // bvh.rs:410-423
        Box lb = empty_box(), rb = empty_box();
        size_t ln = 0, rn = 0;
        for (int k = 0; k <= s; ++k) {
          merge(lb, bin_box[k]);
          ln += bin_n[k];
        }
        for (int k = s + 1; k < kBuckets; ++k) {
          merge(rb, bin_box[k]);
          rn += bin_n[k];
        }
        float cost = ((float)ln * area(lb) + (float)rn * area(rb)) / area(whole);
        if (cost < best_cost) {  // NaN (an empty side) never wins, as in the reference
          best = s;
          best_cost = cost;
        }
      }
      for (int k = 0; k <= best; ++k) left.insert(left.end(), bins[k].begin(), bins[k].end());
      for (int k = best + 1; k < kBuckets; ++k) right.insert(right.end(), bins[k].begin(), bins[k].end());
    }

    int32_t me = (int32_t)out.nodes.size();
    out.nodes.emplace_back();
    Box lb = bounds_of(left.data(), left.size()), rb = bounds_of(right.data(), right.size());
    int32_t l = build(std::move(left), depth + 1);
    int32_t r = build(std::move(right), depth + 1);
    HostNode &n = out.nodes[me];
    for (int k = 0; k < 3; ++k) {
      n.lmin[k] = lb.mn[k];
      n.lmax[k] = lb.mx[k];
      n.rmin[k] = rb.mn[k];
      n.rmax[k] = rb.mx[k];
    }
    n.left = l;
    n.right = r;
    return me;
  }
};

}  // namespace

BuiltBvh build_bvh(const std::vector<Box> &shapes) {
  Builder b(shapes);
  if (shapes.empty()) return std::move(b.out);
  std::vector<uint32_t> idx(shapes.size());
  for (size_t i = 0; i < idx.size(); ++i) idx[i] = (uint32_t)i;
  b.out.root = b.build(std::move(idx), 0);
  return std::move(b.out);
}

}  // namespace rpt
