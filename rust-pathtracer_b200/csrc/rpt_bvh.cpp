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

namespace {

struct SahBuilder {
  const std::vector<Box> &shapes;
  std::vector<float> cx, cy, cz;  // centroids
  BuiltBvh out;
  static constexpr int kBins = 16;

  explicit SahBuilder(const std::vector<Box> &s) : shapes(s), cx(s.size()), cy(s.size()), cz(s.size()) {
    for (size_t i = 0; i < s.size(); ++i) {
      cx[i] = 0.5f * (s[i].mn[0] + s[i].mx[0]);
      cy[i] = 0.5f * (s[i].mn[1] + s[i].mx[1]);
      cz[i] = 0.5f * (s[i].mn[2] + s[i].mx[2]);
    }
  }
  float cen(uint32_t i, int axis) const { return axis == 0 ? cx[i] : (axis == 1 ? cy[i] : cz[i]); }
  static float half_area(const Box &b) {
    float sx = b.mx[0] - b.mn[0], sy = b.mx[1] - b.mn[1], sz = b.mx[2] - b.mn[2];
    return sx * sy + sx * sz + sy * sz;
  }

  int32_t build(uint32_t *idx, size_t n, uint32_t depth) {
    out.max_depth = std::max(out.max_depth, depth);
    if (n == 1) return ~(int32_t)idx[0];
    Box cb = empty_box();
    for (size_t i = 0; i < n; ++i)
      for (int k = 0; k < 3; ++k) {
        float c = cen(idx[i], k);
        cb.mn[k] = std::fmin(cb.mn[k], c);
        cb.mx[k] = std::fmax(cb.mx[k], c);
      }
    int best_axis = -1, best_split = 0;
    float best_cost = kInf;
    for (int axis = 0; axis < 3; ++axis) {
      float ext = cb.mx[axis] - cb.mn[axis];
      if (!(ext > 0.0f)) continue;
      Box bb[kBins];
      size_t bn[kBins] = {0};
      for (auto &b : bb) b = empty_box();
      float scale = (float)kBins / ext;
      for (size_t i = 0; i < n; ++i) {
        int b = std::min(kBins - 1, (int)((cen(idx[i], axis) - cb.mn[axis]) * scale));
        merge(bb[b], shapes[idx[i]]);
        bn[b]++;
      }
      float right_area[kBins];
      size_t right_n[kBins];
      Box acc = empty_box();
      size_t cnt = 0;
      for (int b = kBins - 1; b > 0; --b) {
        merge(acc, bb[b]);
        cnt += bn[b];
        right_area[b] = half_area(acc);
        right_n[b] = cnt;
      }
      acc = empty_box();
      cnt = 0;
      for (int b = 0; b < kBins - 1; ++b) {
        merge(acc, bb[b]);
        cnt += bn[b];
        if (cnt == 0 || right_n[b + 1] == 0) continue;
        float cost = (float)cnt * half_area(acc) + (float)right_n[b + 1] * right_area[b + 1];
        if (cost < best_cost) {
          best_cost = cost;
          best_axis = axis;
          best_split = b;
        }
      }
    }
    size_t mid;
    if (best_axis < 0) {
      mid = n / 2;  // all centroids coincide
    } else {
      float ext = cb.mx[best_axis] - cb.mn[best_axis], scale = (float)kBins / ext, lo = cb.mn[best_axis];
      int axis = best_axis, split = best_split;
      uint32_t *m = std::partition(idx, idx + n, [&](uint32_t i) { return std::min(kBins - 1, (int)((cen(i, axis) - lo) * scale)) <= split; });
      mid = (size_t)(m - idx);
      if (mid == 0 || mid == n) mid = n / 2;
    }
    int32_t me = (int32_t)out.nodes.size();
    out.nodes.emplace_back();
    Box lb = empty_box(), rb = empty_box();
    for (size_t i = 0; i < mid; ++i) merge(lb, shapes[idx[i]]);
    for (size_t i = mid; i < n; ++i) merge(rb, shapes[idx[i]]);
    int32_t l = build(idx, mid, depth + 1);
    int32_t r = build(idx + mid, n - mid, depth + 1);
    HostNode &nd = out.nodes[me];
    for (int k = 0; k < 3; ++k) {
      nd.lmin[k] = lb.mn[k];
      nd.lmax[k] = lb.mx[k];
      nd.rmin[k] = rb.mn[k];
      nd.rmax[k] = rb.mx[k];
    }
    nd.left = l;
    nd.right = r;
    return me;
  }
};

}  // namespace

BuiltBvh build_bvh_sah(const std::vector<Box> &shapes) {
  SahBuilder b(shapes);
  if (shapes.empty()) return std::move(b.out);
  std::vector<uint32_t> idx(shapes.size());
  for (size_t i = 0; i < idx.size(); ++i) idx[i] = (uint32_t)i;
  b.out.root = b.build(idx.data(), idx.size(), 0);
  return std::move(b.out);
}

BuiltBvh build_bvh(const std::vector<Box> &shapes) {
  Builder b(shapes);
  if (shapes.empty()) return std::move(b.out);
  std::vector<uint32_t> idx(shapes.size());
  for (size_t i = 0; i < idx.size(); ++i) idx[i] = (uint32_t)i;
  b.out.root = b.build(std::move(idx), 0);
  return std::move(b.out);
}

namespace {

struct Collapser {
  const BuiltBvh &in;
  WideBvh out;
  struct Slot {
    Box box;
    int32_t ref;
  };
  static float half_area(const Box &b) {
    float sx = b.mx[0] - b.mn[0], sy = b.mx[1] - b.mn[1], sz = b.mx[2] - b.mn[2];
    return sx * sy + sx * sz + sy * sz;
  }
  static void children_of(const HostNode &n, Slot &l, Slot &r) {
    for (int k = 0; k < 3; ++k) {
      l.box.mn[k] = n.lmin[k];
      l.box.mx[k] = n.lmax[k];
      r.box.mn[k] = n.rmin[k];
      r.box.mx[k] = n.rmax[k];
    }
    l.ref = n.left;
    r.ref = n.right;
  }
  // returns the wide child-ref of two-wide ref `ref`; *need = stack entries the walk below it can hold
  int32_t convert(int32_t ref, uint32_t *need) {
    if (ref < 0) {
      *need = 0;
      return ref;
    }
    Slot s[4];
    int n = 2;
    children_of(in.nodes[ref], s[0], s[1]);
    while (n < 4) {
      int pick = -1;
      float best = -1.0f;
      for (int i = 0; i < n; ++i) {
        if (s[i].ref < 0) continue;
        float a = half_area(s[i].box);
        if (!(a <= best)) {  // NaN / inf areas (degenerate boxes) still get picked
          best = a;
          pick = i;
        }
      }
      if (pick < 0) break;
      Slot l, r;
      children_of(in.nodes[s[pick].ref], l, r);
      s[pick] = l;
      s[n++] = r;
    }
    int32_t me = (int32_t)out.nodes.size();
    out.nodes.emplace_back();
    uint32_t deepest = 0;
    int32_t refs[4];
    for (int i = 0; i < n; ++i) {
      uint32_t sub = 0;
      refs[i] = convert(s[i].ref, &sub);
      deepest = std::max(deepest, sub);
    }
    WideNode &w = out.nodes[me];
    for (int i = 0; i < 4; ++i) {
      for (int k = 0; k < 3; ++k) {
        w.plane[k][i] = i < n ? s[i].box.mn[k] : 0.0f;
        w.plane[3 + k][i] = i < n ? s[i].box.mx[k] : 0.0f;
      }
      w.child[i] = i < n ? refs[i] : kEmptyChild;
      w.pad[i] = 0;
    }
    *need = (uint32_t)(n - 1) + deepest;  // the siblings wait on the stack while the first child is walked
    return me;
  }
};

}  // namespace

WideBvh collapse_bvh4(const BuiltBvh &bvh) {
  Collapser c{bvh, {}};
  c.out.root = c.convert(bvh.root, &c.out.stack_need);
  return std::move(c.out);
}

}  // namespace rpt
