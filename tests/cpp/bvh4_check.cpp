// CPU check of rpt::collapse_bvh4 (rust-pathtracer_b200/csrc/rpt_bvh.cpp), built and run by tests/test_host_logic.py.
// For random shape sets: every leaf of the two-wide tree appears exactly once in the four-wide tree, each child box is the
// two-wide tree's box of that subtree, and for random rays the un-pruned walks of both trees reach the same leaves while the
// nearest-first walk of the wide tree never holds more refs than WideBvh::stack_need.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <set>
#include <vector>

#include "../../rust-pathtracer_b200/csrc/rpt_bvh.h"

using namespace rpt;

static bool slab(const float mn[3], const float mx[3], const float o[3], const float inv[3], float &tn) {
  float t0 = 0.0f, t1 = INFINITY;
  for (int k = 0; k < 3; ++k) {
    float a = (mn[k] - o[k]) * inv[k], b = (mx[k] - o[k]) * inv[k];
    if (a > b) std::swap(a, b);
    if (a == a) t0 = std::max(t0, a);
    if (b == b) t1 = std::min(t1, b);
  }
  tn = t0;
  return t0 <= t1 * 1.000001f;
}

static void walk2(const BuiltBvh &b, int32_t ref, const float o[3], const float inv[3], std::set<int32_t> &leaves) {
  if (ref < 0) {
    leaves.insert(~ref);
    return;
  }
  const HostNode &n = b.nodes[ref];
  float t;
  if (slab(n.lmin, n.lmax, o, inv, t)) walk2(b, n.left, o, inv, leaves);
  if (slab(n.rmin, n.rmax, o, inv, t)) walk2(b, n.right, o, inv, leaves);
}

static uint32_t walk4(const WideBvh &w, const float o[3], const float inv[3], std::set<int32_t> &leaves) {
  std::vector<int32_t> stack;
  uint32_t high = 0;
  int32_t cur = w.root;
  while (true) {
    if (cur < 0) {
      leaves.insert(~cur);
      if (stack.empty()) break;
      cur = stack.back();
      stack.pop_back();
      continue;
    }
    const WideNode &n = w.nodes[cur];
    std::vector<std::pair<float, int32_t>> hits;
    for (int c = 0; c < 4; ++c) {
      if (n.child[c] == kEmptyChild) continue;
      float mn[3] = {n.plane[0][c], n.plane[1][c], n.plane[2][c]}, mx[3] = {n.plane[3][c], n.plane[4][c], n.plane[5][c]};
      float t;
      if (slab(mn, mx, o, inv, t)) hits.push_back({t, n.child[c]});
    }
    std::sort(hits.begin(), hits.end());
    if (hits.empty()) {
      if (stack.empty()) break;
      cur = stack.back();
      stack.pop_back();
      continue;
    }
    for (size_t i = hits.size(); i-- > 1;) stack.push_back(hits[i].second);
    high = std::max<uint32_t>(high, (uint32_t)stack.size());
    cur = hits[0].second;
  }
  return high;
}

static bool same_box(const float a[3], const float b[3]) { return a[0] == b[0] && a[1] == b[1] && a[2] == b[2]; }

// the two-wide tree's box of subtree `ref` as stored in its parent
struct ParentBox {
  float mn[3], mx[3];
};

int main() {
  std::mt19937 rng(7);
  std::uniform_real_distribution<float> U(-1.0f, 1.0f);
  int fails = 0;
  for (size_t n : {size_t(1), size_t(2), size_t(3), size_t(4), size_t(5), size_t(9), size_t(33), size_t(1000), size_t(20000)}) {
    std::vector<Box> shapes(n);
    for (auto &s : shapes) {
      float c[3] = {U(rng), U(rng), U(rng)}, e = 0.02f + 0.05f * std::fabs(U(rng));
      for (int k = 0; k < 3; ++k) {
        s.mn[k] = c[k] - e * std::fabs(U(rng));
        s.mx[k] = c[k] + e * std::fabs(U(rng));
      }
    }
    for (int which = 0; which < 2; ++which) {
      BuiltBvh b = which ? build_bvh(shapes) : build_bvh_sah(shapes);
      WideBvh w = collapse_bvh4(b);
      // (1) leaves exactly once, (2) child boxes are the two-wide boxes of those refs
      std::vector<ParentBox> box_of_leaf(n), box_of_inner(b.nodes.size());
      for (auto &nd : b.nodes) {
        for (int side = 0; side < 2; ++side) {
          int32_t r = side ? nd.right : nd.left;
          ParentBox pb;
          for (int k = 0; k < 3; ++k) {
            pb.mn[k] = side ? nd.rmin[k] : nd.lmin[k];
            pb.mx[k] = side ? nd.rmax[k] : nd.lmax[k];
          }
          if (r < 0) box_of_leaf[~r] = pb; else box_of_inner[r] = pb;
        }
      }
      std::vector<int> seen(n, 0);
      size_t inner_children = 0;
      if (w.root < 0) seen[~w.root]++;
      for (auto &wn : w.nodes) {
        int kids = 0;
        for (int c = 0; c < 4; ++c) {
          if (wn.child[c] == kEmptyChild) continue;
          ++kids;
          float mn[3] = {wn.plane[0][c], wn.plane[1][c], wn.plane[2][c]}, mx[3] = {wn.plane[3][c], wn.plane[4][c], wn.plane[5][c]};
          if (wn.child[c] < 0) {
            int32_t leaf = ~wn.child[c];
            seen[leaf]++;
            if (!same_box(mn, box_of_leaf[leaf].mn) || !same_box(mx, box_of_leaf[leaf].mx)) ++fails;
          } else {
            ++inner_children;
          }
        }
        if (kids < 2) ++fails;
      }
      for (size_t i = 0; i < n; ++i)
        if (seen[i] != 1) ++fails;
      if (n > 1 && inner_children + 1 != w.nodes.size()) ++fails;  // a tree: every node but the root has one parent
      if (n >= 4 && w.nodes.size() * 2 > b.nodes.size() + 2) ++fails;  // it did collapse
      // (3) random rays, some axis-aligned
      uint32_t high = 0;
      for (int r = 0; r < 2000; ++r) {
        float o[3] = {1.5f * U(rng), 1.5f * U(rng), 1.5f * U(rng)}, d[3] = {U(rng), U(rng), U(rng)};
        if (r % 7 == 0) d[r % 3] = 0.0f;
        if (r % 11 == 0) d[(r + 1) % 3] = 0.0f;
        if (d[0] == 0.0f && d[1] == 0.0f && d[2] == 0.0f) d[0] = 1.0f;
        float inv[3] = {1.0f / d[0], 1.0f / d[1], 1.0f / d[2]};
        std::set<int32_t> l2, l4;
        walk2(b, b.root, o, inv, l2);
        high = std::max(high, walk4(w, o, inv, l4));
        if (l2 != l4) ++fails;
      }
      if (high > w.stack_need) ++fails;
      std::printf("n=%zu builder=%d nodes2=%zu nodes4=%zu stack_need=%u observed=%u depth2=%u fails=%d\n", n, which, b.nodes.size(), w.nodes.size(),
                  w.stack_need, high, b.max_depth, fails);
    }
  }
  return fails ? 1 : 0;
}
