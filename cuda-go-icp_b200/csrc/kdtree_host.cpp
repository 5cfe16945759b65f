// kdtree_host.cpp -- see kdtree_host.h.  Host logic only; all searches run on the GPU.
#include "kdtree_host.h"
#include <algorithm>
#include <cstring>
#include <thread>

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

// One builder = one contiguous block of nodes in pre-order (a node, its whole left subtree, its whole right subtree).
// Large right subtrees are built by another builder on its own thread and spliced in behind the left one: same node
// numbering as a sequential build (child1 == parent + 1), a 1 M-point model in 0.58 s -> a fraction of that on the
// host cores a registration otherwise leaves idle.
struct Builder {
    const float* pts; int32_t* vind;                 // vind: the shared permutation; builders own disjoint ranges of it
    std::vector<KdNode> nodes; std::vector<float> boxes;
    int divide(int left, int right, float* lo, float* hi, int leaf_max, int depth);
};

int Builder::divide(int left, int right, float* lo, float* hi, int leaf_max, int depth)
{
    const float* pts_ = pts;
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
    int32_t* ind = vind + left;
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
    std::memcpy(rlo, lo, sizeof rlo); std::memcpy(rhi, hi, sizeof rhi); rlo[axis] = cut;
    int c1, c2;
    std::thread worker;
    bool threaded = depth < 3 && count - split >= 16384;
    Builder right_b{pts, vind, {}, {}};
    if (threaded) {
        try { worker = std::thread([&]() { right_b.divide(left + split, right, rlo, rhi, leaf_max, depth + 1); }); }
        catch (...) { threaded = false; }                            // no thread to be had: build it here
    }
    if (threaded) {
        c1 = divide(left, left + split, llo, lhi, leaf_max, depth + 1);
        worker.join();
        c2 = (int)nodes.size();                                      // splice: the right block follows the left one
        for (KdNode k : right_b.nodes) {
            if (k.child1 >= 0) k.child1 += c2;
            if (k.child2 >= 0) k.child2 += c2;
            nodes.push_back(k);
        }
        boxes.insert(boxes.end(), right_b.boxes.begin(), right_b.boxes.end());
    } else {
        c1 = divide(left, left + split, llo, lhi, leaf_max, depth + 1);
        c2 = divide(left + split, right, rlo, rhi, leaf_max, depth + 1);
    }
    KdNode& n = nodes[id];
    n.child1 = c1; n.child2 = c2; n.left = n.right = 0; n.divfeat = axis; n.pad = 0;
    n.divlow = lhi[axis]; n.divhigh = rlo[axis];                     // tightened child boxes (:962-963)
    for (int a = 0; a < 3; ++a) { lo[a] = std::min(llo[a], rlo[a]); hi[a] = std::max(lhi[a], rhi[a]); }
    keep_box();
    const int32_t mid_pos = left + split;                            // first leaf-order position of child2's points
    std::memcpy(&boxes[8 * (size_t)id + 3], &mid_pos, sizeof mid_pos);
    return id;
}

} // namespace

void HostKdTree::build(const float* xyz, int n, int leaf_max)
{
    pts_ = xyz;
    vind.resize(n);
    for (int i = 0; i < n; ++i) vind[i] = i;
    for (int a = 0; a < 3; ++a) bb_lo[a] = bb_hi[a] = xyz[a];        // computeBoundingBox (:895-917)
    for (int k = 1; k < n; ++k)
        for (int a = 0; a < 3; ++a) {
            const float v = xyz[3 * k + a];
            if (v < bb_lo[a]) bb_lo[a] = v;
            if (v > bb_hi[a]) bb_hi[a] = v;
        }
    Builder b{xyz, vind.data(), {}, {}};
    b.divide(0, n, bb_lo, bb_hi, leaf_max, 0);                        // root box is tightened in place (:761)
    nodes.swap(b.nodes);
    boxes.swap(b.boxes);
}

} // namespace goicp
