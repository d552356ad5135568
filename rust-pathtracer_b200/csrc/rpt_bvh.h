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

// Four-wide form of a built tree for the device's wide traversal (rpt_device.cuh TravT<true>): 128 bytes per node, laid out
// as the kernel reads it — rows [min x, min y, min z, max x, max y, max z][child], then the four child refs (same convention as
// HostNode; kEmptyChild marks an unused slot), then padding to one cache line.
struct WideNode {
  float plane[6][4];
  int32_t child[4];
  int32_t pad[4];
};
static_assert(sizeof(WideNode) == 128, "WideNode is one 128-byte line");
constexpr int32_t kEmptyChild = INT32_MIN;  // == RPT_DONE on the device

struct WideBvh {
  std::vector<WideNode> nodes;
  int32_t root = -1;        // child-ref of the root (a leaf ref when the tree has one shape)
  uint32_t stack_need = 0;  // most refs the nearest-first walk can hold on its stack at once
};

// Collapses the two-wide tree: every inner node adopts its grandchildren, largest surface area first, until it has four
// children or only leaves are left. Boxes and leaf refs are copied bit for bit, so the wide tree encloses exactly what the
// two-wide one does.
WideBvh collapse_bvh4(const BuiltBvh &bvh);

}  // namespace rpt
