// kdtree_host.h -- host-side construction of the model kd-tree in the reference's exact shape.
//
// The nearest-neighbour INDEX the reference returns (incl. which of several equidistant model
// points wins) is a function of the tree layout produced by its old nanoflann fork
// (nanoflann_goicp.hpp:927-1111: divideTree / middleSplit_ / planeSplit, leaf size 10,
// jly_icp3d.hpp:151).  This builder reproduces that layout -- including the split-axis choice
// that results from middleSplit_ measuring the spread along the *current* cut axis instead of
// the loop axis (:1049) -- and flattens it for the device traversal in icp_kernels.cu.
#pragma once
#include <vector>
#include <cstdint>
#include "goicp_kernels.h"

namespace goicp {

struct HostKdTree {
    std::vector<KdNode> nodes;
    std::vector<int32_t> vind;          // permutation of model indices; leaves own [left,right)
    std::vector<float> boxes;           // 8 floats per node: tight bounding box of the subtree's points {lo xyz, m, hi xyz, 0}; interior nodes: m = int bits of the first leaf-order position of child2
    float bb_lo[3], bb_hi[3];           // root bounding box after the build tightened it
    void build(const float* xyz, int n, int leaf_max = 10);
private:
    const float* pts_ = nullptr;
};

} // namespace goicp
