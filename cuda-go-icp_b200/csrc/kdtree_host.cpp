// kdtree_host.cpp -- see kdtree_host.h.  Host logic only; all searches run on the GPU.
#include "kdtree_host.h"
#include <algorithm>
#include <cstring>

namespace goicp {

namespace {

struct Span { float lo, hi; };

inline Span extent(const float* pts, const int32_t* ind, int count, int axis)      // computeMinMax (:977-986)
{
    Span s{pts[3 * ind[0] + axis], pts[3 * ind[0] + axis]};
    for (int i = 1; i < count; ++i) {
        const float v = pts[3 * ind[i] + axis];
        if (v < s.lo) s.lo = v;
        if (v > s.hi) s.hi = v;
    }
    return s;
}

// Three-way partition around `cut` on `axis` with the reference's two Hoare-style sweeps
// (planeSplit, :1084-1111): afterwards [0,lim1) < cut, [lim1,lim2) == cut, [lim2,count) > cut.
void plane_split(const float* pts, int32_t* ind, int count, int axis, float cut, int& lim1, int& lim2)
{
    auto key = [&](long i) { return pts[3 * ind[i] + axis]; };
    long lo = 0, hi = count - 1;
    while (true) {
        while (lo <= hi && key(lo) < cut) ++lo;
        while (hi != 0 && lo <= hi && key(hi) >= cut) --hi;
        if (lo > hi || hi == 0) break;
        std::swap(ind[lo], ind[hi]);
        ++lo; --hi;
    }
    lim1 = (int)lo;
    hi = count - 1;
    while (true) {
        while (lo <= hi && key(lo) <= cut) ++lo;
        while (hi != 0 && lo <= hi && key(hi) > cut) --hi;
        if (lo > hi || hi == 0) break;
        std::swap(ind[lo], ind[hi]);
        ++lo; --hi;
    }
    lim2 = (int)lo;
}

} // namespace

int HostKdTree::divide(int left, int right, float* lo, float* hi, int leaf_max)
{
    const int id = (int)nodes.size();
    nodes.push_back(KdNode{});
    boxes.resize(boxes.size() + 8, 0.0f);
    // on return lo/hi hold the tight box of the points below this node (the reference tightens them the same
    // way, :939-946,:965-968); the device search uses it to skip subtrees that cannot hold a closer point
    auto keep_box = [&]() { for (int a = 0; a < 3; ++a) { boxes[8 * (size_t)id + a] = lo[a]; boxes[8 * (size_t)id + 4 + a] = hi[a]; } };
    if (right - left <= leaf_max) {                                  // leaf (:932-947)
        KdNode& n = nodes[id];
        n.child1 = n.child2 = -1; n.left = left; n.right = right; n.divfeat = 0; n.divlow = n.divhigh = 0.0f; n.pad = 0;
        for (int a = 0; a < 3; ++a) lo[a] = hi[a] = pts_[3 * vind[left] + a];
        for (int k = left + 1; k < right; ++k)
            for (int a = 0; a < 3; ++a) {
                const float v = pts_[3 * vind[k] + a];
                if (lo[a] > v) lo[a] = v;
                if (hi[a] < v) hi[a] = v;
            }
        keep_box();
        return id;
    }
    // ---- split selection (middleSplit_, :1033-1072) ----
    int32_t* ind = vind.data() + left;
    const int count = right - left;
    const float eps = (float)0.00001;
    float widest = hi[0] - lo[0];
    for (int a = 1; a < 3; ++a) widest = std::max(widest, hi[a] - lo[a]);
    int axis = 0;
    float best_spread = -1.0f;
    for (int a = 0; a < 3; ++a) {
        if (hi[a] - lo[a] > (1 - eps) * widest) {
            const Span s = extent(pts_, ind, count, axis);           // sic: measured along the running `axis`
            const float spread = s.hi - s.lo;
            if (spread > best_spread) { axis = a; best_spread = spread; }
        }
    }
    const float mid = (lo[axis] + hi[axis]) / 2;
    const Span s = extent(pts_, ind, count, axis);
    const float cut = mid < s.lo ? s.lo : (mid > s.hi ? s.hi : mid);
    int lim1, lim2;
    plane_split(pts_, ind, count, axis, cut, lim1, lim2);
    const int half = count / 2;
    const int split = lim1 > half ? lim1 : (lim2 < half ? lim2 : half);

    float llo[3], lhi[3], rlo[3], rhi[3];
    std::memcpy(llo, lo, sizeof llo); std::memcpy(lhi, hi, sizeof lhi); lhi[axis] = cut;
    const int c1 = divide(left, left + split, llo, lhi, leaf_max);
    std::memcpy(rlo, lo, sizeof rlo); std::memcpy(rhi, hi, sizeof rhi); rlo[axis] = cut;
    const int c2 = divide(left + split, right, rlo, rhi, leaf_max);
    KdNode& n = nodes[id];
    n.child1 = c1; n.child2 = c2; n.left = n.right = 0; n.divfeat = axis; n.pad = 0;
    n.divlow = lhi[axis]; n.divhigh = rlo[axis];                     // tightened child boxes (:962-963)
    for (int a = 0; a < 3; ++a) { lo[a] = std::min(llo[a], rlo[a]); hi[a] = std::max(lhi[a], rhi[a]); }
    keep_box();
    const int32_t mid_pos = left + split;                            // first leaf-order position of child2's points
    std::memcpy(&boxes[8 * (size_t)id + 3], &mid_pos, sizeof mid_pos);
    return id;
}

void HostKdTree::build(const float* xyz, int n, int leaf_max)
{
    pts_ = xyz;
    nodes.clear();
    boxes.clear();
    vind.resize(n);
    for (int i = 0; i < n; ++i) vind[i] = i;
    for (int a = 0; a < 3; ++a) bb_lo[a] = bb_hi[a] = xyz[a];        // computeBoundingBox (:895-917)
    for (int k = 1; k < n; ++k)
        for (int a = 0; a < 3; ++a) {
            const float v = xyz[3 * k + a];
            if (v < bb_lo[a]) bb_lo[a] = v;
            if (v > bb_hi[a]) bb_hi[a] = v;
        }
    divide(0, n, bb_lo, bb_hi, leaf_max);                             // root box is tightened in place (:761)
}

} // namespace goicp
