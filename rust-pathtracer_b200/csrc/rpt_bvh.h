// rpt_bvh.h — host-side BVH construction for the device scene.
//
// Row a7 of SURVEY.md §8: the binary tree is built with the reference's own algorithm
// (src/accelerator/bvh.rs:299-457: SAH over 6 buckets on the longest centroid axis, split-in-half
// fallback when the centroid extent is < 1e-5, one shape per leaf) so that
//  (1) the candidate order of the reference's flattened tree (src/accelerator/lbvh.rs:87-134) is
//      known — it only matters for exact-t tie-breaking (rpt_device.cuh tie_key), and
//  (2) the device tree has the same SAH quality the reference traverses.
// The tree is then re-encoded into 64-byte two-child nodes (DevNode) that hold both children's boxes.
#pragma once

#include <cstdint>
#include <vector>

namespace rpt {

struct Box {
  float mn[3], mx[3];
};

struct HostNode {  // mirrors DevNode (rpt_device.cuh)
  float lmin[3], lmax[3];
  float rmin[3], rmax[3];
  int32_t left, right;  // >= 0 inner node index (relative to this tree), < 0 leaf: ~shape
};

struct BuiltBvh {
  std::vector<HostNode> nodes;
  int32_t root = -1;            // child-ref of the root
  std::vector<uint32_t> order;  // order[shape] = position in the reference's flat candidate order
  uint32_t max_depth = 0;
};

BuiltBvh build_bvh(const std::vector<Box> &shapes);

// Traversal-quality builder for the DEVICE tree: binned SAH over all three axes (16 bins), one shape per leaf.
// Which tree the device walks does not change any result (closest hit + tie keys are tree-independent); the
// reference-order tree above is still built, but only to number the shapes for tie-breaking.
BuiltBvh build_bvh_sah(const std::vector<Box> &shapes);

}  // namespace rpt
