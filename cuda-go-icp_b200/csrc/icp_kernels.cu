// icp_kernels.cu -- ICP refinement on sm_100a: reference-ordered kd-tree nearest neighbour,
// fused transform + NN + moment reduction, device-side 3x3 SVD / Procrustes and SE(3) compose,
// all iterations inside ONE cooperative kernel (no host round trip per iteration).
//
// Replaces ICP3D<float>::Run (jly_icp3d.hpp:180-295), KDTreeSingleIndexAdaptor::searchLevel
// (nanoflann_goicp.hpp:1136-1184) and Matrix::svd (matrix.cpp:602-830) for the GPU.
#include <cooperative_groups.h>
#include <algorithm>
#include <cstdlib>
#include <cstddef>
#include "goicp_kernels.h"

namespace cg = cooperative_groups;

namespace goicp {

constexpr int kIcpThreads = 512;
constexpr int kKdStack = 64;

// ------------------------------------------------------------------------------------------
// Exact 1-NN with the reference's traversal order.  The reference recursion
//     search(best child); if (mindistsq + cut - dists[idx] <= worst) search(other child)
// is unrolled onto an explicit stack; leaves are scanned in ascending vind order and a point is
// accepted only if strictly closer than the incumbent, so the first-visited of several
// exactly-equidistant model points wins, as in the reference (nanoflann_goicp.hpp:95-116,1143-1150).
//
// Subtree skipping.  The reference prunes with slab distances only (divlow/divhigh along the split
// axes of the path), so a query at distance r from the surface walks every cell under the ball's
// shadow -- O(r^2 * density) leaves, none of which can hold a closer point.  Each node also carries
// the tight box of its points (kdtree_host.cpp); a subtree whose box is farther than the incumbent
// (with a 1e-5 relative margin, two orders above the float rounding of either side) cannot change
// (worst, best): every `dist < worst` test below it fails and the reference restores its per-axis
// offsets on the way out.  Skipping it leaves the traversal's state -- hence the returned index,
// ties included -- exactly the reference's.
//
// `cap` (optional, FLT_MAX = none) is any known upper bound of the nearest distance -- the distance to
// some model point, e.g. the previous iteration's correspondent.  Subtrees whose box is farther than
// min(worst, cap) are skipped as well.  That no longer mirrors the reference's intermediate `worst`
// (ours stays >= its value at every point of the common depth-first order, so we prune a subset of
// what it prunes by the slab test), but the RESULT is the same: both traversals end at the smallest
// float distance d*, no skipped subtree holds a point at d* (all of its points are > cap >= d*), the
// extra subtrees we enter hold only points the reference had already beaten, and a point replaces the
// incumbent only if strictly closer -- so both return the first point at d* in that common order.
// (Only a query whose runner-up is within float rounding, 6e-7 relative, of d* AND whose offset to
// the winner is parallel to the cut axes could make the reference's rounded slab bound skip the
// winner itself; the caller's margin for treating a result as a tie is above that.)
// ------------------------------------------------------------------------------------------
// One pending far-side branch of the walk, 16 bytes so that a push or a pop is ONE local-memory access: x = far child |
// split axis << 28 | visited << 30, y/z/w = float bits of its bound, of the axis offset inside it and of the offset to
// restore afterwards.  (As six scalars a frame cost six scattered 4-byte accesses per lane; with 512 walks per SM the
// stacks do not fit L1 and that traffic, not the tree, was what the walk waited for.)
typedef uint4 KdFrame;

// Visit the points [left,right) of a leaf in order.  The loads of five points are issued together: a leaf (<= 10
// points, jly_icp3d.hpp:151) costs two round trips to L2 instead of one per point -- the compiler does not batch
// the loads of a loop whose trip count it does not know.
template <typename F>
__device__ __forceinline__ void for_leaf_points(const float4* __restrict__ leaf, int left, int right, F f)
{
    for (int base = left; base < right; base += 5) {
        float4 p[5];
#pragma unroll
        for (int k = 0; k < 5; k++) if (base + k < right) p[k] = leaf[base + k];
#pragma unroll
        for (int k = 0; k < 5; k++) if (base + k < right) f(base + k, p[k]);
    }
}

__device__ __forceinline__ float sel3(float a, float b, float c, int i) { return i == 0 ? a : (i == 1 ? b : c); }

// `budget` bounds the node visits (<= 0: unlimited): a walk that exceeds it returns -1 and the caller hands the
// query to the warp-cooperative search below.  pos_out = leaf-order position of the returned point.
__device__ __forceinline__ int kd_nearest(const KdView& kd, const KdNode* __restrict__ nodes, const float4* __restrict__ leaf, float qx, float qy, float qz, float cap,
                          float& d2_out, int budget = 0, int* pos_out = nullptr)
{
    int best_pos = 0;
    if (budget <= 0) budget = 0x7fffffff;
    float worst = 3.402823466e+38f;     // KNNResultSet::init (nanoflann_goicp.hpp:79)
    int best = 0;
    float ds0 = 0.0f, ds1 = 0.0f, ds2 = 0.0f;
    float distsq = 0.0f;                // computeInitialDistances (:1113-1130)
    {
        const float q[3] = {qx, qy, qz};
        float ds[3] = {0.0f, 0.0f, 0.0f};
#pragma unroll
        for (int i = 0; i < 3; i++) {
            if (q[i] < kd.bb_lo[i]) { float d = q[i] - kd.bb_lo[i]; ds[i] = d * d; distsq += ds[i]; }
            if (q[i] > kd.bb_hi[i]) { float d = q[i] - kd.bb_hi[i]; ds[i] = d * d; distsq += ds[i]; }
        }
        ds0 = ds[0]; ds1 = ds[1]; ds2 = ds[2];
    }
    KdFrame stack[kKdStack];
    int sp = 0;
    int cur = 0;
    float cur_min = distsq;
    for (;;) {
        // descend along the preferred children
        for (;;) {
            if (--budget < 0) return -1;
            KdNode nd;
            {
                const uint4 a = reinterpret_cast<const uint4*>(nodes + cur)[0], b = reinterpret_cast<const uint4*>(nodes + cur)[1];
                nd.child1 = (int)a.x; nd.child2 = (int)a.y; nd.left = (int)a.z; nd.right = (int)a.w;
                nd.divfeat = (int)b.x; nd.divlow = __uint_as_float(b.y); nd.divhigh = __uint_as_float(b.z); nd.pad = 0;
            }
            if (kd.boxes != nullptr && (worst < 3.0e+38f || cap < 3.0e+38f)) {
                const float4 lo = __ldg(kd.boxes + 2 * (size_t)cur), hi = __ldg(kd.boxes + 2 * (size_t)cur + 1);
                const float bx = fmaxf(fmaxf(lo.x - qx, qx - hi.x), 0.0f);
                const float by = fmaxf(fmaxf(lo.y - qy, qy - hi.y), 0.0f);
                const float bz = fmaxf(fmaxf(lo.z - qz, qz - hi.z), 0.0f);
                if (bx * bx + by * by + bz * bz > fminf(worst, cap) * 1.00001f) break;      // nothing below can be the answer
            }
            if (nd.child1 < 0 && nd.child2 < 0) {
                const float worst_at_entry = worst;          // cached once per leaf (:1143)
                for (int i = nd.left; i < nd.right; i++) {
                    const float4 p = leaf[i];
                    const float d0 = qx - p.x, d1 = qy - p.y, d2 = qz - p.z;
                    const float dist = d0 * d0 + d1 * d1 + d2 * d2;     // kdtree_distance (jly_icp3d.hpp:48-54)
                    if (dist < worst_at_entry && worst > dist) { worst = dist; best = __float_as_int(p.w); best_pos = i; }
                }
                break;
            }
            const int idx = nd.divfeat;
            const float val = sel3(qx, qy, qz, idx);
            const float diff1 = val - nd.divlow, diff2 = val - nd.divhigh;
            int bestc, otherc; float cut;
            if ((diff1 + diff2) < 0) { bestc = nd.child1; otherc = nd.child2; cut = (val - nd.divhigh) * (val - nd.divhigh); }
            else                     { bestc = nd.child2; otherc = nd.child1; cut = (val - nd.divlow) * (val - nd.divlow); }
            const float dst = sel3(ds0, ds1, ds2, idx);
            if (sp < kKdStack)
                stack[sp++] = make_uint4((unsigned)otherc | ((unsigned)idx << 28), __float_as_uint(cur_min + cut - dst), __float_as_uint(cut), __float_as_uint(dst));
            cur = bestc;
        }
        // unwind
        bool descend = false;
        while (sp > 0) {
            const KdFrame f = stack[sp - 1];
            const int fidx = (int)((f.x >> 28) & 3u);
            if ((f.x >> 30) == 0u) {
                const float m2 = __uint_as_float(f.y);
                if (m2 * 1.0f <= worst) {         // epsError = 1 (:809, :1178)
                    stack[sp - 1].x = f.x | (1u << 30);
                    const float fcut = __uint_as_float(f.z);
                    if (fidx == 0) ds0 = fcut; else if (fidx == 1) ds1 = fcut; else ds2 = fcut;
                    cur = (int)(f.x & 0x0fffffffu); cur_min = m2; descend = true;
                    break;
                }
                sp--;
            } else {
                const float fdst = __uint_as_float(f.w);
                if (fidx == 0) ds0 = fdst; else if (fidx == 1) ds1 = fdst; else ds2 = fdst;
                sp--;
            }
        }
        if (!descend) break;
    }
    d2_out = worst;
    if (pos_out) *pos_out = best_pos;
    return best;
}

// ------------------------------------------------------------------------------------------
// Warp-cooperative search for the queries whose walk is long (far from the surface: hundreds of
// dependent node visits).  Same answer as kd_nearest, found differently:
//  1. the minimum distance D1, its point and the runner-up distance D2 by a breadth-first frontier
//     expansion -- 32 nodes per round trip to L2 instead of one; a node is dropped when its tight box
//     is farther than the best distance so far (1e-5 relative margin), so every point within that
//     margin of D1 is scanned and D2 is exact whenever it matters;
//  2. D2 > D1 (1 + 1e-5): the point at D1 is what the reference returns -- its walk ends at the smallest
//     float distance, and its slab bound, a lower bound of every distance below a node up to ~2e-6 of
//     rounding, cannot have excluded that point against an incumbent >= D2;
//  3. D1 < D2 <= D1 (1 + 1e-5): replay the reference's slab bound along the winner's root path (the
//     bound depends on the path only): if every far-side step has bound <= D2 <= the incumbent at that
//     moment, the reference does reach the winner's leaf, and the strictly smallest distance wins;
//  4. two points at exactly D1 (0.1 % of far queries on a 1 M-point model): the same replay, following both root paths
//     until they part -- the point under the preferred child (same leaf: the lower position) is met first and keeps
//     the title -- with the incumbent >= D1 until then;
//  5. three or more exact ties, a failed replay or a frontier that outgrows its queue: returns -1 and the best bound
//     found in d2_out; the caller runs the reference walk itself (kd_nearest) capped by it.
// q: this warp's queue in shared memory (kCoopQ ints).  seed_pos: leaf-order position of any model
// point (an upper bound of D1), or -1.  All lanes return the same values.
// ------------------------------------------------------------------------------------------
constexpr int kCoopQ = 1024;        // ring entries per warp; frontiers of far queries on a 1 M-point model reach ~430
__device__ __forceinline__ int kd_coop_nearest(const KdView& kd, int* q, float qx, float qy, float qz, int seed_pos, int lane, float& d2_out, int& pos_out)
{
    const unsigned full = 0xffffffffu;
    const float kInfF = 3.402823466e+38f;
    const KdNode* __restrict__ nodes = kd.nodes; const float4* __restrict__ leaf = kd.pts_leaf;
    float bd1 = kInfF, bd2 = kInfF, lim = kInfF; int bpos = -1, bpos2 = -1, ntie = 0;     // ntie: points at exactly bd1 seen by this lane
    if (seed_pos >= 0) {
        const float4 p = __ldg(leaf + seed_pos);
        const float e0 = qx - p.x, e1 = qy - p.y, e2 = qz - p.z;
        lim = e0 * e0 + e1 * e1 + e2 * e2;
        if (lane == 0) { bd1 = lim; bpos = seed_pos; ntie = 1; }
    }
    int head = 0, tail = 1, live = 1;                  // ring positions in [0, kCoopQ); live = entries queued
    if (kd.n_top == 32) { q[lane] = kd.top[lane]; tail = live = 32; }     // start five levels down: a full round at once
    else if (lane == 0) q[0] = 0;
    __syncwarp();
    bool overflow = false;
    while (live > 0) {
        const int n = min(32, live);
        int node = -1;
        if (lane < n) { int at = head + lane; if (at >= kCoopQ) at -= kCoopQ; node = q[at]; }
        head += n; if (head >= kCoopQ) head -= kCoopQ;
        live -= n;
        int c1 = -1, c2 = -1;
        if (node >= 0) {
            const uint4 a = __ldg(reinterpret_cast<const uint4*>(nodes + node));
            const float4 lo = __ldg(kd.boxes + 2 * (size_t)node), hi = __ldg(kd.boxes + 2 * (size_t)node + 1);
            const float bx = fmaxf(fmaxf(lo.x - qx, qx - hi.x), 0.0f);
            const float by = fmaxf(fmaxf(lo.y - qy, qy - hi.y), 0.0f);
            const float bz = fmaxf(fmaxf(lo.z - qz, qz - hi.z), 0.0f);
            if (bx * bx + by * by + bz * bz <= lim * 1.00001f) {
                if ((int)a.x < 0 && (int)a.y < 0) {
                    for_leaf_points(leaf, (int)a.z, (int)a.w, [&](int i, const float4& p) {
                        if (i == seed_pos) return;                     // already lane 0's candidate
                        const float e0 = qx - p.x, e1 = qy - p.y, e2 = qz - p.z;
                        const float dist = e0 * e0 + e1 * e1 + e2 * e2; // kdtree_distance (jly_icp3d.hpp:48-54)
                        if (dist < bd1) { bd2 = bd1; bd1 = dist; bpos = i; ntie = 1; }
                        else if (dist == bd1) { bd2 = dist; bpos2 = i; ntie++; }
                        else if (dist < bd2) bd2 = dist;
                    });
                } else { c1 = (int)a.x; c2 = (int)a.y; }
            }
        }
        // distances are non-negative floats: their bit patterns order like the values
        lim = fminf(lim, __uint_as_float(__reduce_min_sync(full, __float_as_uint(bd1))));
        const unsigned pm = __ballot_sync(full, c1 >= 0);
        const int total = 2 * __popc(pm);
        if (live + total > kCoopQ) { overflow = true; break; }
        if (c1 >= 0) {
            int at = tail + 2 * __popc(pm & ((1u << lane) - 1u));
            if (at >= kCoopQ) at -= kCoopQ;
            q[at] = c1;
            at++; if (at >= kCoopQ) at -= kCoopQ;
            q[at] = c2;
        }
        tail += total; if (tail >= kCoopQ) tail -= kCoopQ;
        live += total;
        __syncwarp();
    }
    float D1 = kInfF, D2 = kInfF; int p1 = 0; bool settled = false;
    if (!overflow) {
        D1 = __uint_as_float(__reduce_min_sync(full, __float_as_uint(bd1)));
        const unsigned at_min = __ballot_sync(full, bd1 == D1);
        const int win = __ffs(at_min) - 1;
        D2 = __uint_as_float(__reduce_min_sync(full, __float_as_uint(lane == win ? bd2 : bd1)));
        p1 = __shfl_sync(full, bpos, win);
        const int ties = __reduce_add_sync(full, bd1 == D1 ? ntie : 0);            // points at exactly D1
        int p2 = -1;                                                                  // the other one of a two-way exact tie
        if (ties == 2) {
            const int win_n = __shfl_sync(full, ntie, win);
            const int other = win_n == 2 ? win : __ffs(at_min & (at_min - 1)) - 1;
            p2 = __shfl_sync(full, win_n == 2 ? bpos2 : bpos, other);
        }
        if (D2 > D1 * 1.00001f) settled = true;
        else if (D2 > D1 || ties == 2) {
            // Replay of the reference's bound along the root path (computeInitialDistances + searchLevel,
            // nanoflann_goicp.hpp:1113-1184); every lane walks the same path.  With a unique minimum the incumbent the
            // reference holds when it tests a far-side step is >= D2; with two points at exactly D1 the walk follows both
            // until their paths part -- the one under the preferred child is visited first and, strict '<', keeps the
            // title (same leaf: the lower position) -- and the incumbent is >= D1 until that point is reached.
            const float bound_ok = ties == 2 ? D1 : D2;
            float ds0 = 0.0f, ds1 = 0.0f, ds2 = 0.0f, cur_min = 0.0f;
            {
                const float qq[3] = {qx, qy, qz};
                float ds[3] = {0.0f, 0.0f, 0.0f};
#pragma unroll
                for (int i = 0; i < 3; i++) {
                    if (qq[i] < kd.bb_lo[i]) { float d = qq[i] - kd.bb_lo[i]; ds[i] = d * d; cur_min += ds[i]; }
                    if (qq[i] > kd.bb_hi[i]) { float d = qq[i] - kd.bb_hi[i]; ds[i] = d * d; cur_min += ds[i]; }
                }
                ds0 = ds[0]; ds1 = ds[1]; ds2 = ds[2];
            }
            int cur = 0; settled = true;
            for (;;) {
                const uint4 a = __ldg(reinterpret_cast<const uint4*>(nodes + cur)), b = __ldg(reinterpret_cast<const uint4*>(nodes + cur) + 1);
                if ((int)a.x < 0 && (int)a.y < 0) { if (p2 >= 0) p1 = min(p1, p2); break; }     // leaves are scanned in ascending position
                const int mid = __float_as_int(__ldg(kd.boxes + 2 * (size_t)cur).w);
                const int idx = (int)b.x; const float divlow = __uint_as_float(b.y), divhigh = __uint_as_float(b.z);
                const float val = sel3(qx, qy, qz, idx);
                const float diff1 = val - divlow, diff2 = val - divhigh;
                int bestc, otherc; float cut;
                if ((diff1 + diff2) < 0) { bestc = (int)a.x; otherc = (int)a.y; cut = (val - divhigh) * (val - divhigh); }
                else                     { bestc = (int)a.y; otherc = (int)a.x; cut = (val - divlow) * (val - divlow); }
                int target = p1 < mid ? (int)a.x : (int)a.y;
                if (p2 >= 0) {
                    const int target2 = p2 < mid ? (int)a.x : (int)a.y;
                    if (target2 != target) {                             // the paths part here: the preferred child's point comes first
                        if (target2 == bestc) { p1 = p2; target = target2; }
                        p2 = -1;
                    }
                }
                if (target == bestc) { cur = bestc; continue; }       // the preferred child is always searched
                const float dst = sel3(ds0, ds1, ds2, idx);
                const float m2 = cur_min + cut - dst;
                if (!(m2 <= bound_ok)) { settled = false; break; }     // cannot prove the reference gets here
                if (idx == 0) ds0 = cut; else if (idx == 1) ds1 = cut; else ds2 = cut;
                cur_min = m2; cur = otherc;
            }
        }
    }
    if (settled) { d2_out = D1; pos_out = p1; return __float_as_int(__ldg(leaf + p1).w); }
    d2_out = overflow ? lim : D1;          // not settled here: the caller runs the reference walk, capped by this bound
    pos_out = 0;
    return -1;
}

__global__ void nn_kernel(KdView kd, const float* __restrict__ q, int n, int32_t* __restrict__ idx, float* __restrict__ d2)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float d;
    idx[i] = kd_nearest(kd, kd.nodes, kd.pts_leaf, q[3 * i], q[3 * i + 1], q[3 * i + 2], 3.402823466e+38f, d);
    d2[i] = d;
}
// the same answers through the warp-cooperative search, one query per warp at a time (parity tests: GOICP_NN_COOP=1)
__global__ void __launch_bounds__(128) nn_coop_kernel(KdView kd, const float* __restrict__ q, int n, int32_t* __restrict__ idx, float* __restrict__ d2)
{
    __shared__ int queue[4][kCoopQ];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = blockIdx.x * 4 + warp; i < n; i += gridDim.x * 4) {
        float d; int pos;
        int id = kd_coop_nearest(kd, queue[warp], q[3 * i], q[3 * i + 1], q[3 * i + 2], -1, lane, d, pos);
        if (id < 0) id = kd_nearest(kd, kd.nodes, kd.pts_leaf, q[3 * i], q[3 * i + 1], q[3 * i + 2], d, d);      // every lane the same walk
        if (lane == 0) { idx[i] = id; d2[i] = d; }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
// 3x3 SVD: Numerical-Recipes svdcmp in float with the reference's double promotions, followed by
// the descending sort and the sign normalisation (matrix.cpp:602-830).  One thread.
// ------------------------------------------------------------------------------------------
__device__ float svd_pythag(float a, float b)
{
    float absa = fabsf(a), absb = fabsf(b);
    if (absa > absb) { float r = absb / absa; return (float)((double)absa * sqrt(1.0 + (r == 0.0f ? 0.0 : (double)(r * r)))); }
    if (absb == 0.0f) return 0.0f;
    float r = absa / absb;
    return (float)((double)absb * sqrt(1.0 + (r == 0.0f ? 0.0 : (double)(r * r))));
}
__device__ __forceinline__ float svd_sign(float a, float b) { return b >= 0.0f ? fabsf(a) : -fabsf(a); }

// Register-resident form: every index below is a compile-time constant once the loops are unrolled -- the three passes over
// k are unrolled and the data-dependent lower end l of a QR sweep is dispatched to a <K, L> instantiation -- so U, V, w and
// rv1 live in registers instead of local memory (the one-thread solve sat on the critical path of every ICP iteration:
// 14.7 k cycles with the arrays in local memory).  The floating-point operations and their order are the reference's.
template <int K, int L>
__device__ __forceinline__ bool svd_qr_sweep(float (&U)[3][3], float (&V)[3][3], float (&w)[3], float (&rv1)[3], const float anorm, const int flag)
{
    constexpr int m = 3, n = 3;
    float c, f, g, h, s, x, y, z;
    if (flag) {
        constexpr int nm = L - 1;                        // flag is only set with L >= 1 (rv1[0] == 0 always ends the search at l = 0 with flag = 0)
        c = 0.0f; s = 1.0f;
#pragma unroll
        for (int i = L; i <= K; i++) {
            f = s * rv1[i];
            rv1[i] = c * rv1[i];
            if ((float)(fabsf(f) + anorm) == anorm) break;
            g = w[i];
            h = svd_pythag(f, g);
            w[i] = h;
            h = (float)(1.0 / (double)h);
            c = g * h;
            s = -f * h;
#pragma unroll
            for (int j = 0; j < m; j++) {
                y = U[j][nm < 0 ? 0 : nm]; z = U[j][i];
                U[j][nm < 0 ? 0 : nm] = y * c + z * s;
                U[j][i] = z * c - y * s;
            }
        }
    }
    z = w[K];
    if (L == K) {
        if (z < 0.0f) {
            w[K] = -z;
#pragma unroll
            for (int j = 0; j < n; j++) V[j][K] = -V[j][K];
        }
        return true;                                     // converged for this k
    }
    constexpr int nm = K - 1 < 0 ? 0 : K - 1;            // (L < K implies K >= 1)
    x = w[L]; y = w[nm]; g = rv1[nm]; h = rv1[K];
    f = (float)((double)((y - z) * (y + z) + (g - h) * (g + h)) / (2.0 * (double)h * (double)y));
    g = svd_pythag(f, 1.0f);
    f = ((x - z) * (x + z) + h * ((y / (f + svd_sign(g, f))) - h)) / x;
    c = s = 1.0f;
#pragma unroll
    for (int j = L; j <= nm; j++) {
        const int i = j + 1 > 2 ? 2 : j + 1;
        g = rv1[i]; y = w[i];
        h = s * g; g = c * g;
        z = svd_pythag(f, h);
        rv1[j] = z;
        c = f / z; s = h / z;
        f = x * c + g * s;
        g = g * c - x * s;
        h = y * s;
        y *= c;
#pragma unroll
        for (int jj = 0; jj < n; jj++) {
            x = V[jj][j]; z = V[jj][i];
            V[jj][j] = x * c + z * s;
            V[jj][i] = z * c - x * s;
        }
        z = svd_pythag(f, h);
        w[j] = z;
        if (z != 0.0f) { z = (float)(1.0 / (double)z); c = f * z; s = h * z; }
        f = c * g + s * y;
        x = c * y - s * g;
#pragma unroll
        for (int jj = 0; jj < m; jj++) {
            y = U[jj][j]; z = U[jj][i];
            U[jj][j] = y * c + z * s;
            U[jj][i] = z * c - y * s;
        }
    }
    rv1[L] = 0.0f; rv1[K] = f; w[K] = x;
    return false;
}

template <int K>
__device__ __forceinline__ void svd_qr_k(float (&U)[3][3], float (&V)[3][3], float (&w)[3], float (&rv1)[3], const float anorm)
{
    for (int its = 0; its < 30; its++) {
        // the reference's search for l (:741-749), over constant indices
        int flag = 1, l = -1;
#pragma unroll
        for (int ll = K; ll >= 0; ll--) {
            if (l < 0) {
                if ((float)(fabsf(rv1[ll]) + anorm) == anorm) { flag = 0; l = ll; }
                else if (ll >= 1 && (float)(fabsf(w[ll >= 1 ? ll - 1 : 0]) + anorm) == anorm) l = ll;
            }
        }
        if (l < 0) l = 0;
        bool done;
        if (l == 0) done = svd_qr_sweep<K, 0>(U, V, w, rv1, anorm, flag);
        else if (l == 1) done = svd_qr_sweep<K, (K >= 1 ? 1 : 0)>(U, V, w, rv1, anorm, flag);
        else done = svd_qr_sweep<K, (K >= 2 ? 2 : 0)>(U, V, w, rv1, anorm, flag);
        if (done) break;
    }
}

__device__ void svd3_ref(const float* A9, float* U9, float* W3, float* V9)
{
    constexpr int m = 3, n = 3;
    float U[3][3], V[3][3], w[3], rv1[3];
    float anorm, f, g, h, s, scale;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) { U[i][j] = A9[3 * i + j]; V[i][j] = 0.0f; }
    g = scale = anorm = 0.0f;
#pragma unroll
    for (int i = 0; i < n; i++) {
        const int l = i + 1;
        rv1[i] = scale * g;
        g = s = scale = 0.0f;
        {
#pragma unroll
            for (int k = i; k < m; k++) scale += fabsf(U[k][i]);
            if (scale != 0.0f) {
#pragma unroll
                for (int k = i; k < m; k++) { U[k][i] /= scale; s += U[k][i] * U[k][i]; }
                f = U[i][i];
                g = -svd_sign(sqrtf(s), f);
                h = f * g - s;
                U[i][i] = f - g;
#pragma unroll
                for (int j = l; j < n; j++) {
                    s = 0.0f;
#pragma unroll
                    for (int k = i; k < m; k++) s += U[k][i] * U[k][j];
                    f = s / h;
#pragma unroll
                    for (int k = i; k < m; k++) U[k][j] += f * U[k][i];
                }
#pragma unroll
                for (int k = i; k < m; k++) U[k][i] *= scale;
            }
        }
        w[i] = scale * g;
        g = s = scale = 0.0f;
        if (i != n - 1) {
#pragma unroll
            for (int k = l; k < n; k++) scale += fabsf(U[i][k]);
            if (scale != 0.0f) {
#pragma unroll
                for (int k = l; k < n; k++) { U[i][k] /= scale; s += U[i][k] * U[i][k]; }
                f = U[i][l < 3 ? l : 2];
                g = -svd_sign(sqrtf(s), f);
                h = f * g - s;
                U[i][l < 3 ? l : 2] = f - g;
#pragma unroll
                for (int k = l; k < n; k++) rv1[k] = U[i][k] / h;
#pragma unroll
                for (int j = l; j < m; j++) {
                    s = 0.0f;
#pragma unroll
                    for (int k = l; k < n; k++) s += U[j][k] * U[i][k];
#pragma unroll
                    for (int k = l; k < n; k++) U[j][k] += s * rv1[k];
                }
#pragma unroll
                for (int k = l; k < n; k++) U[i][k] *= scale;
            }
        }
        { float t2 = fabsf(w[i]) + fabsf(rv1[i]); anorm = anorm > t2 ? anorm : t2; }
    }
    // accumulation of the right-hand transformations (:689-706); entering with g = the last value set above, l = n
#pragma unroll
    for (int i = n - 1; i >= 0; i--) {
        const int l = i + 1;
        if (i < n - 1) {
            if (g != 0.0f) {
#pragma unroll
                for (int j = l; j < n; j++) V[j][i] = (U[i][j] / U[i][l < 3 ? l : 2]) / g;
#pragma unroll
                for (int j = l; j < n; j++) {
                    s = 0.0f;
#pragma unroll
                    for (int k = l; k < n; k++) s += U[i][k] * V[k][j];
#pragma unroll
                    for (int k = l; k < n; k++) V[k][j] += s * V[k][i];
                }
            }
#pragma unroll
            for (int j = l; j < n; j++) V[i][j] = V[j][i] = 0.0f;
        }
        V[i][i] = 1.0f;
        g = rv1[i];
    }
    // accumulation of the left-hand transformations (:707-726)
#pragma unroll
    for (int i = 2; i >= 0; i--) {
        const int l = i + 1;
        g = w[i];
#pragma unroll
        for (int j = l; j < n; j++) U[i][j] = 0.0f;
        if (g != 0.0f) {
            g = (float)(1.0 / (double)g);
#pragma unroll
            for (int j = l; j < n; j++) {
                s = 0.0f;
#pragma unroll
                for (int k = l; k < m; k++) s += U[k][i] * U[k][j];
                f = (s / U[i][i]) * g;
#pragma unroll
                for (int k = i; k < m; k++) U[k][j] += f * U[k][i];
            }
#pragma unroll
            for (int j = i; j < m; j++) U[j][i] *= g;
        } else {
#pragma unroll
            for (int j = i; j < m; j++) U[j][i] = 0.0f;
        }
        U[i][i] = U[i][i] + 1.0f;
    }
    // diagonalisation of the bidiagonal form (:727-781)
    svd_qr_k<2>(U, V, w, rv1, anorm);
    svd_qr_k<1>(U, V, w, rv1, anorm);
    svd_qr_k<0>(U, V, w, rv1, anorm);
    // descending sort (insertion with gap 1 for n = 3) + sign normalisation (:783-818)
    {   // i = 1
        const float sw = w[1], su0 = U[0][1], su1 = U[1][1], su2 = U[2][1], sv0 = V[0][1], sv1 = V[1][1], sv2 = V[2][1];
        if (w[0] < sw) {
            w[1] = w[0]; U[0][1] = U[0][0]; U[1][1] = U[1][0]; U[2][1] = U[2][0]; V[0][1] = V[0][0]; V[1][1] = V[1][0]; V[2][1] = V[2][0];
            w[0] = sw; U[0][0] = su0; U[1][0] = su1; U[2][0] = su2; V[0][0] = sv0; V[1][0] = sv1; V[2][0] = sv2;
        }
    }
    {   // i = 2
        const float sw = w[2], su0 = U[0][2], su1 = U[1][2], su2 = U[2][2], sv0 = V[0][2], sv1 = V[1][2], sv2 = V[2][2];
        if (w[1] < sw) {
            w[2] = w[1]; U[0][2] = U[0][1]; U[1][2] = U[1][1]; U[2][2] = U[2][1]; V[0][2] = V[0][1]; V[1][2] = V[1][1]; V[2][2] = V[2][1];
            if (w[0] < sw) {
                w[1] = w[0]; U[0][1] = U[0][0]; U[1][1] = U[1][0]; U[2][1] = U[2][0]; V[0][1] = V[0][0]; V[1][1] = V[1][0]; V[2][1] = V[2][0];
                w[0] = sw; U[0][0] = su0; U[1][0] = su1; U[2][0] = su2; V[0][0] = sv0; V[1][0] = sv1; V[2][0] = sv2;
            } else {
                w[1] = sw; U[0][1] = su0; U[1][1] = su1; U[2][1] = su2; V[0][1] = sv0; V[1][1] = sv1; V[2][1] = sv2;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < n; k++) {
        int s2 = 0;
#pragma unroll
        for (int i = 0; i < 3; i++) if (U[i][k] < 0.0f) s2++;
#pragma unroll
        for (int j = 0; j < 3; j++) if (V[j][k] < 0.0f) s2++;
        if (s2 > 3) {
#pragma unroll
            for (int i = 0; i < 3; i++) U[i][k] = -U[i][k];
#pragma unroll
            for (int j = 0; j < 3; j++) V[j][k] = -V[j][k];
        }
    }
#pragma unroll
    for (int i = 0; i < 3; i++) {
        W3[i] = w[i];
#pragma unroll
        for (int j = 0; j < 3; j++) { U9[3 * i + j] = U[i][j]; V9[3 * i + j] = V[i][j]; }
    }
}

__device__ void mat3_mul(const float* A, const float* B, float* C)   // accumulate from 0 in k order (matrix.cpp:287-301)
{
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
        float acc = 0.0f;
        for (int k = 0; k < 3; k++) acc += A[3 * i + k] * B[3 * k + j];
        C[3 * i + j] = acc;
    }
}

__device__ void procrustes_jacobi(const double* H, float* Rn);

// Procrustes step + SE(3) composition (jly_icp3d.hpp:268-291) from H = q_d^T q_m and the means.
// jacobi: GOICP_NUM_JACOBI_SVD -- the rotation from the engine's own solver instead of the reference's svdcmp (experiments:
// with the reference-order sums kept, does the trajectory survive a solver that is merely accurate?  scripts/parity_modes.py)
__device__ void icp_update(IcpState* st, const float* H, bool jacobi, const float* Rc /* current R */, const float* tc /* current t */, const float* mu /* mu_m, mu_d */)
{
    float U[9], W[3], V[9], Ut[9], Rn[9], VT[9], tmp[9];
    if (jacobi) {
        double Hd[9];
        for (int i = 0; i < 9; i++) Hd[i] = (double)H[i];
        procrustes_jacobi(Hd, Rn);
    } else {
    svd3_ref(H, U, W, V);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Ut[3 * j + i] = U[3 * i + j];
    mat3_mul(V, Ut, Rn);                                                     // R_ = V * ~U
    float da = Rn[0] * (Rn[4] * Rn[8] - Rn[5] * Rn[7]);
    float db = -Rn[1] * (Rn[3] * Rn[8] - Rn[5] * Rn[6]);
    float dc = Rn[2] * (Rn[3] * Rn[7] - Rn[4] * Rn[6]);
    float det = da + db + dc;
    float D[9] = {1, 0, 0, 0, 1, 0, 0, 0, det};
    mat3_mul(V, D, VT);
    mat3_mul(VT, Ut, Rn);                                                    // R_ = V * diag(1,1,det) * ~U
    }
    float tn[3], tt[3];
    for (int a = 0; a < 3; a++) {                                            // t_ = ~mu_m - R_ * ~mu_d
        float acc = 0.0f;
        for (int k = 0; k < 3; k++) acc += Rn[3 * a + k] * mu[3 + k];
        tn[a] = mu[a] - acc;
    }
    mat3_mul(Rn, Rc, tmp);                                                   // R = R_ * R
    for (int a = 0; a < 3; a++) {                                            // t = R_ * t + t_
        float acc = 0.0f;
        for (int k = 0; k < 3; k++) acc += Rn[3 * a + k] * tc[k];
        tt[a] = acc + tn[a];
    }
    for (int i = 0; i < 9; i++) st->R[i] = tmp[i];
    for (int i = 0; i < 3; i++) st->t[i] = tt[i];
}

// ------------------------------------------------------------------------------------------
// ICP3D::Run as ONE cooperative kernel.  Per iteration:
//  (all CTAs)  transform every data point in float (reference order), exact kd-tree NN, store
//              (query, index, d^2) and a 64-bit sort key (d^2 bits, point index);
//  grid.sync
//  (CTA 0)     bitonic sort of the keys == the reference's qsort by distance (stable for ties,
//              like glibc's merge sort with the reference's never-zero comparator, :156-160,:238);
//              then the reference's SEQUENTIAL float accumulations in that order -- means on top of
//              the previous means (never reset, :205-206,:244-252), error (:254), H (:268) -- one
//              lane per accumulator; convergence test (:257); SVD, det fix, compose (:268-291);
//  grid.sync   publishes the new pose.
// The whole refinement therefore follows the reference's trajectory bit for bit, which is what
// makes the final pose agree to 1e-4: ICP stops on a loose relative criterion far from its fixed
// point, so its end point depends on the exact path.
// ------------------------------------------------------------------------------------------
constexpr int kIcpChunk = 384;
constexpr int kColPitch = kIcpChunk + 4;       // column pitch of a transposed chunk: 7 columns fit the 8 * kIcpChunk floats, lanes read distinct banks

// ------------------------------------------------------------------------------------------
// Nearest neighbour on small models (the linear-scan range), two stages, both exact:
//  1. grid_nn -- eight lanes per query look at the (2r+1)^3 grid cells around the query, r chosen from the query's nearest
//     distance in the previous iteration (it moves little between iterations; any value is only a hint).  The result is
//     final when the nearest point found is closer than the edge of that block of cells: every point outside the block is
//     at least `margin` away, so none can beat it, nor come within the 1e-5 near-tie band that sends a query to the
//     reference-ordered kd-tree walk.  A few dozen points are looked at instead of the whole model.
//  2. the rest (a wrong hint, or a query so far from the surface that its block would be most of the grid): a linear scan of
//     the whole model by all the warps of the CTA together, one query after the other, merged through shared memory
//     (block_nn_scan); the results are parked and emitted in parallel afterwards.
// Same distances (same float expression), same near-tie rule, hence the same indices as the plain scan.
// ------------------------------------------------------------------------------------------
// what a CTA of the ICP kernel has in (or reaches through) shared memory, see icp_carve
struct IcpSmem { const KdNode* nodes; const float4* leaf; const unsigned short* gstart; const float4* gpts; float* sstage; unsigned* sradix; int* squeue; float4* qpts; float* qhint; };
constexpr int kGridMaxR = 5;             // beyond 11^3 cells the block-wide scan is cheaper
__device__ __forceinline__ void nn_merge(float& d1, int& i1, float& d2, int o)
{
    const unsigned full = 0xffffffffu;
    const float od1 = __shfl_xor_sync(full, d1, o), od2 = __shfl_xor_sync(full, d2, o);
    const int oi1 = __shfl_xor_sync(full, i1, o);
    const float nd2 = fminf(fmaxf(d1, od1), fminf(d2, od2));         // second smallest of the union (an exact tie of the two minima counts)
    if (od1 < d1) { d1 = od1; i1 = oi1; }
    d2 = nd2;
}
// One query per warp, one x-run of cells per lane (two or more beyond r = 2); the points of a run are fetched four at a
// time so that their loads and distance arithmetic overlap -- a one-by-one loop is a chain of ~130 dependent cycles per
// point.  r < 0: skip (the caller already knows the block would be too large); all 32 lanes must call.
// gl = lanes per query: 32 (the whole warp) or 16 (two queries side by side, lanes 0-15 and 16-31 -- a block of 3^3 cells has nine x-runs,
// so half a warp is enough for the usual radius).
__device__ __forceinline__ bool grid_nn(const KdView& kd, const unsigned short* gstart, const float4* gpts, float qx, float qy, float qz, int r, int lane, int gl,
                                        float& D1, int& I1, float& D2)
{
    const float kInfF = 3.402823466e+38f;
    const float q[3] = {qx, qy, qz};
    int c[3]; float margin = kInfF;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const float u = fminf(fmaxf((q[a] - kd.glo[a]) * kd.ginv_h, -1.0f), (float)kd.gdim[a]);
        const int ca = min(max((int)floorf(u), 0), kd.gdim[a] - 1);
        c[a] = ca;
        // a side of the block is closed only if a cell lies beyond it: no model point is below cell 0 or above the last cell
        if (ca - r - 1 >= 0) margin = fminf(margin, q[a] - (kd.glo[a] + (float)(ca - r) * kd.gh));
        if (ca + r + 1 < kd.gdim[a]) margin = fminf(margin, (kd.glo[a] + (float)(ca + r + 1) * kd.gh) - q[a]);
    }
    float d1 = kInfF, d2 = kInfF; int i1 = 0;
    if (r >= 0) {
        const int x0 = max(c[0] - r, 0), x1 = min(c[0] + r, kd.gdim[0] - 1);
        const int side = 2 * r + 1;
        for (int run = lane & (gl - 1); run < side * side; run += gl) {  // x-runs: the cells (x0..x1, cy, cz) are contiguous
            const int cy = c[1] + (run % side) - r, cz = c[2] + (run / side) - r;
            if (cy < 0 || cy >= kd.gdim[1] || cz < 0 || cz >= kd.gdim[2]) continue;
            const int base = (cz * kd.gdim[1] + cy) * kd.gdim[0];
            const int e = gstart[base + x1 + 1];
            for (int m = gstart[base + x0]; m < e; m += 4) {
                float4 pm[4]; float dist[4];
#pragma unroll
                for (int u = 0; u < 4; u++) pm[u] = gpts[min(m + u, e - 1)];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const float e0 = qx - pm[u].x, e1 = qy - pm[u].y, e2 = qz - pm[u].z;
                    dist[u] = m + u < e ? e0 * e0 + e1 * e1 + e2 * e2 : kInfF;    // kdtree_distance (jly_icp3d.hpp:48-54)
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    if (dist[u] < d1) { d2 = d1; d1 = dist[u]; i1 = m + u; }
                    else if (dist[u] < d2) d2 = dist[u];
                }
            }
        }
    }
    if (gl == 32) nn_merge(d1, i1, d2, 16);
    nn_merge(d1, i1, d2, 8); nn_merge(d1, i1, d2, 4); nn_merge(d1, i1, d2, 2); nn_merge(d1, i1, d2, 1);
    D1 = d1; D2 = d2; I1 = __shfl_sync(0xffffffffu, i1, lane & ~(gl - 1));
    const float me = margin - 1e-4f * kd.gh;                        // binning and edge arithmetic round at ~1e-6 of a cell
    return r >= 0 && me > 0.0f && d1 * 1.00001f < me * me;
}
// one query, all the warps of the CTA: every thread scans its stride of the model, warps merge by shuffles, the warps'
// results go through merge[3][16] in shared memory; returns (in warp 0, all lanes) the two smallest distances and lane
// 0's index.  One block barrier per call; the caller alternates between two merge buffers.
__device__ __forceinline__ void block_nn_scan(const float4* __restrict__ pts, int nm, float qx, float qy, float qz, float* merge, int lane, int warp,
                                              float& D1, int& I1, float& D2)
{
    const float kInfF = 3.402823466e+38f;
    float d1 = kInfF, d2 = kInfF; int i1 = 0;
    for (int m = threadIdx.x; m < nm; m += kIcpThreads) {
        const float4 pm = pts[m];
        const float e0 = qx - pm.x, e1 = qy - pm.y, e2 = qz - pm.z;
        const float dist = e0 * e0 + e1 * e1 + e2 * e2;
        if (dist < d1) { d2 = d1; d1 = dist; i1 = m; }
        else if (dist < d2) d2 = dist;
    }
    nn_merge(d1, i1, d2, 16); nn_merge(d1, i1, d2, 8); nn_merge(d1, i1, d2, 4); nn_merge(d1, i1, d2, 2); nn_merge(d1, i1, d2, 1);
    if (lane == 0) { merge[warp] = d1; merge[16 + warp] = d2; merge[32 + warp] = __int_as_float(i1); }
    __syncthreads();
    if (warp == 0) {
        d1 = lane < kIcpThreads / 32 ? merge[lane] : kInfF; d2 = lane < kIcpThreads / 32 ? merge[16 + lane] : kInfF;
        i1 = lane < kIcpThreads / 32 ? __float_as_int(merge[32 + lane]) : 0;
        nn_merge(d1, i1, d2, 8); nn_merge(d1, i1, d2, 4); nn_merge(d1, i1, d2, 2); nn_merge(d1, i1, d2, 1);
        I1 = __shfl_sync(0xffffffffu, i1, 0);
    }
    D1 = d1; D2 = d2;
}
// The NN phase of one ICP iteration for the queries of this CTA (i = q_begin + k * gridDim.x + blockIdx.x < q_end).
// emit(i, model index, model xyz, d^2, query xyz) is called by exactly one thread per query: lane 0 of warp k mod 16 when the
// grid settles it, else thread k mod 512 -- a fixed assignment, so whatever emit accumulates is reproducible from run to run.
// d2_hint: the queries' squared nearest distances of the previous iteration (any content is safe; sm.qhint caches the first
// ones on chip, as sm.qpts caches their data points).  park: >= ceil(nd / gridDim.x) 64-bit words of scratch owned by this
// CTA: 0 = settled by the grid, else bit 63 | d^2 bits << 32 | near-tie << 31 | position of the winner in the scanned array.
template <typename Emit>
__device__ __forceinline__ void icp_nn_small_model(const KdView& kd, const IcpSmem& sm, int qcache_n, const float4* __restrict__ data, int q_begin, int q_end, const float (&R)[9], const float (&t)[3],
                                                   float* d2_hint, unsigned long long* park, float* merge /* [2][3][16] */, Emit emit)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nq = q_end - q_begin > (int)blockIdx.x ? (q_end - q_begin - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const KdNode* nodes = sm.nodes; const float4* leaf = sm.leaf; const unsigned short* gstart = sm.gstart; const float4* gpts = sm.gpts;
    const float4* scan = gpts ? gpts : leaf;                 // either way: x, y, z, original index
    auto query = [&](int k, int i, float& qx, float& qy, float& qz) {
        const float4 p = k < qcache_n ? sm.qpts[k] : __ldg(data + i);
        // query = R p + t, (((r0*x + r1*y) + r2*z) + t) in float (jly_icp3d.hpp:219-221)
        qx = __fadd_rn(dot3_ref(R[0], R[1], R[2], p.x, p.y, p.z), t[0]);
        qy = __fadd_rn(dot3_ref(R[3], R[4], R[5], p.x, p.y, p.z), t[1]);
        qz = __fadd_rn(dot3_ref(R[6], R[7], R[8], p.x, p.y, p.z), t[2]);
    };
    // winner at position `pos` of the scanned array, or -- near tie -- whatever the reference-ordered kd-tree walk returns
    auto finish = [&](int k, int i, int pos, bool tie, float D1, float qx, float qy, float qz) {
        if (k < qcache_n) sm.qhint[k] = D1; else d2_hint[i] = D1;
        const float4 w = scan[pos];
        int I1 = __float_as_int(w.w); float mx = w.x, my = w.y, mz = w.z;
        if (tie) {
            I1 = kd_nearest(kd, nodes, leaf, qx, qy, qz, D1, D1);
            mx = __ldg(kd.model + 3 * I1); my = __ldg(kd.model + 3 * I1 + 1); mz = __ldg(kd.model + 3 * I1 + 2);
        }
        emit(i, I1, mx, my, mz, D1, qx, qy, qz);
    };
    __shared__ int n_park_sm;
    if (threadIdx.x == 0) n_park_sm = 0;
    __syncthreads();
    if (gstart) {
        // one query per warp while there are warps enough, else two per warp (half a warp each)
        const int gl = nq > kIcpThreads / 32 ? 16 : 32, per_warp = 32 / gl, sub = lane / gl;
        for (int kb = warp * per_warp; kb < nq; kb += per_warp * (kIcpThreads / 32)) {
            const int k = kb + sub;
            const bool valid = k < nq;
            const int i = q_begin + k * gridDim.x + blockIdx.x;
            float qx = 0.0f, qy = 0.0f, qz = 0.0f;
            // radius of the block of cells from last iteration's distance: the block's edge is >= r cells away
            int r = -1;
            if (valid) {
                query(k, i, qx, qy, qz);
                r = 1;
                const float hint = k < qcache_n ? sm.qhint[k] : __ldcg(d2_hint + i);
                if (hint >= 0.0f && hint < 1.0e30f) r = (int)(sqrtf(hint) * 1.02f * kd.ginv_h) + 1;
                if (r > kGridMaxR) r = -1;
            }
            float D1, D2; int P1;
            const bool ok = grid_nn(kd, gstart, gpts, qx, qy, qz, r, lane, gl, D1, P1, D2);
            if ((lane & (gl - 1)) == 0 && valid) {
                park[k] = ok ? 0ull : 1ull;
                if (ok) finish(k, i, P1, D2 <= D1 * 1.00001f, D1, qx, qy, qz);
                else atomicAdd(&n_park_sm, 1);
            }
            __syncwarp();
        }
        __syncthreads();
        if (n_park_sm == 0) return;                           // the usual case after the first iteration
    }
    int par = 0;
    for (int k = 0; k < nq; k++) {
        if (gstart && park[k] == 0ull) continue;              // block-uniform
        const int i = q_begin + k * gridDim.x + blockIdx.x;
        float qx, qy, qz;
        query(k, i, qx, qy, qz);
        float D1, D2; int P1;
        block_nn_scan(scan, kd.nm, qx, qy, qz, merge + par * 48, lane, warp, D1, P1, D2);
        par ^= 1;
        if (threadIdx.x == 0)
            park[k] = (1ull << 63) | ((unsigned long long)__float_as_uint(D1) << 32) | (D2 <= D1 * 1.00001f ? 0x80000000ull : 0ull) | (unsigned)P1;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < nq; k += kIcpThreads) {
        const unsigned long long w = park[k];
        if (!(w >> 63)) continue;
        const int i = q_begin + k * gridDim.x + blockIdx.x;
        float qx, qy, qz;
        query(k, i, qx, qy, qz);
        finish(k, i, (int)(w & 0x7fffffffu), (w & 0x80000000ull) != 0, __uint_as_float((unsigned)(w >> 32) & 0x7fffffffu), qx, qy, qz);
    }
}

// dynamic shared memory plan of the ICP kernel (decided on the host)
struct IcpSmemPlan { int tree_nodes; int tree_bytes; int stage_bytes /* keys of the counting sort */; int brute_force; int radix_bytes; int nn_budget; int queue_bytes;
                     int grid_bytes /* NN grid: cell table + cell-sorted points */; int qcache_n /* this CTA's first data points + distance hints kept on chip */; };
constexpr int kQCacheEntry = (int)sizeof(float4) + (int)sizeof(float);
__host__ __device__ inline size_t icp_plan_bytes(const IcpSmemPlan& p)
{
    return (size_t)p.grid_bytes + p.tree_bytes + p.stage_bytes + p.radix_bytes + p.queue_bytes + (((size_t)p.qcache_n * kQCacheEntry + 15) & ~(size_t)15);
}
// carve the dynamic shared memory [NN grid][kd-tree][sort keys][radix][search rings][query cache] and fill the read-only parts
__device__ __forceinline__ IcpSmem icp_carve(const KdView& kd, const IcpSmemPlan& plan, unsigned char* sp, const float4* __restrict__ data, int q_begin, int q_end)
{
    IcpSmem m;
    m.nodes = kd.nodes; m.leaf = kd.pts_leaf; m.gstart = kd.grid_start; m.gpts = kd.grid_pts;
    if (plan.grid_bytes) {
        unsigned short* sg = reinterpret_cast<unsigned short*>(sp);
        float4* sl = reinterpret_cast<float4*>(sp + (((size_t)(kd.gcells + 1) * sizeof(unsigned short) + 15) & ~(size_t)15));
        for (int i = threadIdx.x; i < kd.gcells + 1; i += blockDim.x) sg[i] = kd.grid_start[i];
        for (int i = threadIdx.x; i < kd.nm; i += blockDim.x) sl[i] = kd.grid_pts[i];
        m.gstart = sg; m.gpts = sl; sp += plan.grid_bytes;
    }
    if (plan.tree_bytes) {
        KdNode* sn = reinterpret_cast<KdNode*>(sp);
        float4* sl = reinterpret_cast<float4*>(sp + (size_t)plan.tree_nodes * sizeof(KdNode));
        for (int i = threadIdx.x; i < plan.tree_nodes * 2; i += blockDim.x) reinterpret_cast<uint4*>(sn)[i] = reinterpret_cast<const uint4*>(kd.nodes)[i];
        for (int i = threadIdx.x; i < kd.nm; i += blockDim.x) sl[i] = kd.pts_leaf[i];
        m.nodes = sn; m.leaf = sl; sp += plan.tree_bytes;
    }
    m.sstage = plan.stage_bytes ? reinterpret_cast<float*>(sp) : nullptr; sp += plan.stage_bytes;
    m.sradix = plan.radix_bytes ? reinterpret_cast<unsigned*>(sp) : nullptr; sp += plan.radix_bytes;
    m.squeue = plan.queue_bytes ? reinterpret_cast<int*>(sp) : nullptr; sp += plan.queue_bytes;
    m.qpts = nullptr; m.qhint = nullptr;
    if (plan.qcache_n) {
        // the queries of this CTA never change (i = q_begin + k * gridDim.x + blockIdx.x): keep the first qcache_n data points on chip,
        // next to their nearest distance of the previous iteration (the grid search's radius hint; -1 = none yet)
        m.qpts = reinterpret_cast<float4*>(sp); m.qhint = reinterpret_cast<float*>(sp + (size_t)plan.qcache_n * sizeof(float4));
        for (int k = threadIdx.x; k < plan.qcache_n; k += blockDim.x) {
            const long long i = (long long)q_begin + (long long)k * gridDim.x + blockIdx.x;
            m.qpts[k] = i < q_end ? __ldg(data + i) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            m.qhint[k] = -1.0f;
        }
    }
    __syncthreads();
    return m;
}

// ------------------------------------------------------------------------------------------
// Phase B for clouds too large to rank by counting (the keys no longer fit in shared memory and
// the O(nd^2) count would stream them through L2 nd times): a grid-wide, STABLE least-significant
// -digit radix sort of the 64-bit keys over the 32 distance bits, 8 bits per pass.  The keys enter
// in point order, so stability reproduces the (d^2, index) order the counting version produces.
// Every warp of the grid owns one contiguous slice of the array:
//   1. per-warp digit histogram (shared-memory atomics)           -> hist[digit][warp]   grid.sync
//   2. exclusive scan of that digit-major table, one chunk per CTA -> + chunk totals      grid.sync
//   3. each warp re-walks its slice in order, 32 keys at a time: match_any groups equal digits,
//      rank within the group = earlier lanes, destination = scanned base + running count  grid.sync
// ------------------------------------------------------------------------------------------
constexpr int kRadixBlockOffs = 1024;          // max CTAs of the cooperative grid the scan supports
constexpr int kRadixSmemBytes = (kIcpThreads / 32 * 256 + 96 + kRadixBlockOffs) * 4;

__device__ void icp_radix_sort(cg::grid_group& grid, const IcpWork& wk, int nd, unsigned* sm)
{
    constexpr int kWarps = kIcpThreads / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int NW = (int)gridDim.x * kWarps, gw = (int)blockIdx.x * kWarps + warp;
    int per = (nd + NW - 1) / NW; per = (per + 31) & ~31;
    const int r0 = (int)min((long long)gw * per, (long long)nd), r1 = min(r0 + per, nd);
    unsigned* wh = sm + warp * 256;                     // this warp's histogram, then its running bases
    unsigned* sscan = sm + kWarps * 256;                // [0..31] warp totals, [32..63] exclusive, [64] block total
    unsigned* boff = sscan + 96;                        // exclusive scan of the chunk totals
    const int M = 256 * NW;
    const int chunk = (M + (int)gridDim.x - 1) / (int)gridDim.x;
    unsigned long long* src = wk.keys; unsigned long long* dst = wk.keys2;
    for (int pass = 0; pass < 4; pass++) {
        const int shift = 32 + 8 * pass;
        // ---- 1. histogram of this warp's slice
        for (int d = lane; d < 256; d += 32) wh[d] = 0u;
        __syncwarp();
        for (int i = r0 + lane; i < r1; i += 32) atomicAdd(&wh[(unsigned)(__ldcg(src + i) >> shift) & 255u], 1u);
        __syncwarp();
        for (int d = lane; d < 256; d += 32) wk.hist[(size_t)d * NW + gw] = wh[d];
        grid.sync();
        // ---- 2. chunk-local exclusive scan (in place) + chunk total
        {
            const int e0 = min((int)blockIdx.x * chunk, M), e1 = min(e0 + chunk, M);
            const int ept = (chunk + kIcpThreads - 1) / kIcpThreads;
            const int t0 = min(e0 + (int)threadIdx.x * ept, e1), t1 = min(t0 + ept, e1);
            unsigned s = 0;
            for (int e = t0; e < t1; e++) s += __ldcg(wk.hist + e);
            unsigned incl = s;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const unsigned v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
            if (lane == 31) sscan[warp] = incl;
            __syncthreads();
            if (warp == 0) {
                const unsigned v = lane < kWarps ? sscan[lane] : 0u;
                unsigned iv = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const unsigned u = __shfl_up_sync(0xffffffffu, iv, o); if (lane >= o) iv += u; }
                sscan[32 + lane] = iv - v;
                if (lane == 31) sscan[64] = iv;
            }
            __syncthreads();
            unsigned run = sscan[32 + warp] + incl - s;
            for (int e = t0; e < t1; e++) { const unsigned v = __ldcg(wk.hist + e); wk.hist[e] = run; run += v; }
            if (threadIdx.x == 0) wk.blocksum[blockIdx.x] = sscan[64];
        }
        grid.sync();
        // ---- 3. stable scatter
        if (warp == 0) {
            unsigned carry = 0;
            for (int b0 = 0; b0 < (int)gridDim.x; b0 += 32) {
                const unsigned v = b0 + lane < (int)gridDim.x ? __ldcg(wk.blocksum + b0 + lane) : 0u;
                unsigned iv = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const unsigned u = __shfl_up_sync(0xffffffffu, iv, o); if (lane >= o) iv += u; }
                if (b0 + lane < (int)gridDim.x) boff[b0 + lane] = carry + iv - v;
                carry += __shfl_sync(0xffffffffu, iv, 31);
            }
        }
        __syncthreads();
        for (int d = lane; d < 256; d += 32) { const int e = d * NW + gw; wh[d] = __ldcg(wk.hist + e) + boff[e / chunk]; }
        __syncwarp();
        for (int i0 = r0; i0 < r1; i0 += 32) {
            const int i = i0 + lane;
            const bool valid = i < r1;
            const unsigned long long k = valid ? __ldcg(src + i) : 0ull;
            const unsigned dg = valid ? ((unsigned)(k >> shift) & 255u) : (256u + (unsigned)lane);
            const unsigned m = __match_any_sync(0xffffffffu, dg);
            const unsigned rank = __popc(m & ((1u << lane) - 1u));
            const unsigned pos = valid ? wh[dg] + rank : 0u;
            __syncwarp();
            if (valid) {
                dst[pos] = k;
                if (pass == 3) {                              // final position: publish the index and the correspondence row
                    const unsigned i = (unsigned)k;
                    wk.order[pos] = (int)i;
                    const float4* row = reinterpret_cast<const float4*>(wk.q) + 2 * (size_t)i;
                    const float4 lo = __ldcg(row), hi = __ldcg(row + 1);
                    reinterpret_cast<float4*>(wk.stage)[2 * (size_t)pos] = lo; reinterpret_cast<float4*>(wk.stage)[2 * (size_t)pos + 1] = hi;
                }
                if (rank == 0) wh[dg] += __popc(m);
            }
            __syncwarp();
        }
        __syncthreads();                                 // boff / sscan are rewritten by the next pass
        grid.sync();
        unsigned long long* tmp = src; src = dst; dst = tmp;
    }
}

// acc += x[0], x[1], ... (cnt terms, x 16-byte aligned) strictly in order.  The accumulation buffers are laid out one
// COLUMN per adding lane, so a lane fetches four operands per shared-memory instruction; 16 operands are in flight
// ahead of the dependent chain of adds, which is then all that is left: ~4 cycles per term.
__device__ __forceinline__ float seq_add_contig(float acc, const float* x, int cnt)
{
    const float4* x4 = reinterpret_cast<const float4*>(x);
    int rr = 0;
    if (cnt >= 16) {
        float4 a0 = x4[0], a1 = x4[1], a2 = x4[2], a3 = x4[3];
        for (rr = 16; rr + 16 <= cnt; rr += 16) {
            const float4 b0 = x4[rr / 4], b1 = x4[rr / 4 + 1], b2 = x4[rr / 4 + 2], b3 = x4[rr / 4 + 3];
            acc = __fadd_rn(acc, a0.x); acc = __fadd_rn(acc, a0.y); acc = __fadd_rn(acc, a0.z); acc = __fadd_rn(acc, a0.w);
            acc = __fadd_rn(acc, a1.x); acc = __fadd_rn(acc, a1.y); acc = __fadd_rn(acc, a1.z); acc = __fadd_rn(acc, a1.w);
            acc = __fadd_rn(acc, a2.x); acc = __fadd_rn(acc, a2.y); acc = __fadd_rn(acc, a2.z); acc = __fadd_rn(acc, a2.w);
            acc = __fadd_rn(acc, a3.x); acc = __fadd_rn(acc, a3.y); acc = __fadd_rn(acc, a3.z); acc = __fadd_rn(acc, a3.w);
            a0 = b0; a1 = b1; a2 = b2; a3 = b3;
        }
        acc = __fadd_rn(acc, a0.x); acc = __fadd_rn(acc, a0.y); acc = __fadd_rn(acc, a0.z); acc = __fadd_rn(acc, a0.w);
        acc = __fadd_rn(acc, a1.x); acc = __fadd_rn(acc, a1.y); acc = __fadd_rn(acc, a1.z); acc = __fadd_rn(acc, a1.w);
        acc = __fadd_rn(acc, a2.x); acc = __fadd_rn(acc, a2.y); acc = __fadd_rn(acc, a2.z); acc = __fadd_rn(acc, a2.w);
        acc = __fadd_rn(acc, a3.x); acc = __fadd_rn(acc, a3.y); acc = __fadd_rn(acc, a3.z); acc = __fadd_rn(acc, a3.w);
    }
    for (; rr < cnt; rr++) acc = __fadd_rn(acc, x[rr]);
    return acc;
}

__global__ void __launch_bounds__(kIcpThreads)
icp_kernel(KdView kd, const float4* __restrict__ data, int nd, IcpState* st, IcpWork wk,
           int max_iter, float err_diff, int num, int flags /* bit 0: sort (the reference's do_trim), bit 1: Jacobi solver */, IcpSmemPlan plan)
{
    const int do_sort = flags & 1;
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) unsigned char icp_smem[];
    __shared__ float sh_H[9];
    __shared__ float sh_acc[8];
    __shared__ float sh_mu[6];                                   // block 0: mu_m, mu_d of this iteration
    __shared__ int sh_conv;                                      // block 0: converged in this iteration
    __shared__ float sh_err;                                     // block 0: err of the previous iteration
    __shared__ __align__(16) float chunk[2][kIcpChunk * 8];      // double buffer of the streamed (large-cloud) accumulation
    __shared__ int n_deferred, n_unsettled;
    __shared__ float nn_merge_sm[2 * 48];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int kWarps = kIcpThreads / 32;
    if (threadIdx.x == 0) { sh_conv = 0; sh_err = st->err; }
    if (threadIdx.x < 6) sh_mu[threadIdx.x] = threadIdx.x < 3 ? st->mu_m[threadIdx.x] : st->mu_d[threadIdx.x - 3];

    // ---- carve dynamic shared memory: [kd-tree nodes | leaf points] [staged rows]
    const IcpSmem sm = icp_carve(kd, plan, icp_smem, data, 0, nd);
    const KdNode* nodes = sm.nodes; const float4* leaf = sm.leaf;
    float* sstage = sm.sstage; unsigned* sradix = sm.sradix; int* squeue = sm.squeue;

    long long c_nn = 0, c_wait = 0, c_sort = 0, c_p1 = 0, c_p2 = 0, c_acc1 = 0, c_svd = 0; const long long c_begin = clock64();
    for (int iter = 0; iter < max_iter; iter++) {
        long long c0 = clock64();
        float R[9], t[3];
        {   // the pose (and, from the second iteration on, the convergence flag) block 0 wrote before the grid barrier: four
            // 16-byte L2 loads in flight together (volatile scalar loads are issued one after the other -- thirteen L2 round
            // trips, ~10 k cycles per iteration)
            const float4* sp4 = reinterpret_cast<const float4*>(st);
            const float4 s0 = __ldcg(sp4), s1 = __ldcg(sp4 + 1), s2 = __ldcg(sp4 + 2);
            static_assert(offsetof(IcpState, converged) == 84 && offsetof(IcpState, R) == 0 && offsetof(IcpState, t) == 36, "IcpState layout");
            if (iter > 0 && __float_as_int(__ldcg(sp4 + 5).y)) break;
            R[0] = s0.x; R[1] = s0.y; R[2] = s0.z; R[3] = s0.w; R[4] = s1.x; R[5] = s1.y; R[6] = s1.z; R[7] = s1.w; R[8] = s2.x;
            t[0] = s2.y; t[1] = s2.z; t[2] = s2.w;
        }
        // ---- phase A: transform + nearest neighbour; queries interleaved over the CTAs ----------
        if (plan.brute_force) {
            unsigned long long* park = wk.keys2 + (size_t)blockIdx.x * ((nd + gridDim.x - 1) / gridDim.x);      // wk.keys2 is idle until the sort
            icp_nn_small_model(kd, sm, plan.qcache_n, data, 0, nd, R, t, wk.d2, park, nn_merge_sm, [&](int i, int I1, float mx, float my, float mz, float D1, float qx, float qy, float qz) {
                wk.nn[i] = I1;
                wk.keys[i] = ((unsigned long long)__float_as_uint(D1) << 32) | (unsigned)i;
                // correspondence row (model point, query, d^2) in query order
                float4* row = reinterpret_cast<float4*>(wk.q) + 2 * (size_t)i;
                row[0] = make_float4(mx, my, mz, qx);
                row[1] = make_float4(qy, qz, D1, 0.0f);
            });
        } else {
            // Tree search.  With a visit budget > 1 every thread first walks the reference's traversal for its own queries
            // (subtree skipping, capped by the distance to the point's last correspondent; any model point bounds the
            // nearest distance from above, a stale or never-written position is just a loose cap) and a walk that exceeds
            // plan.nn_budget node visits is put on the CTA's list; the listed queries (by default: all of them) are
            // answered by the warp-cooperative search, one query per warp at a time.
            int* deferred = wk.order + (size_t)blockIdx.x * ((nd + gridDim.x - 1) / gridDim.x);   // wk.order is idle until the sort
            if (threadIdx.x == 0) { n_deferred = 0; n_unsettled = 0; }
            __syncthreads();
            auto emit = [&](int i, int id, int pos, float d2, float qx, float qy, float qz) {
                wk.nn[i] = id; wk.pos[i] = pos; wk.d2[i] = d2;
                wk.keys[i] = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned)i;
                float4* row = reinterpret_cast<float4*>(wk.q) + 2 * (size_t)i;
                row[0] = make_float4(__ldg(kd.model + 3 * id), __ldg(kd.model + 3 * id + 1), __ldg(kd.model + 3 * id + 2), qx);
                row[1] = make_float4(qy, qz, d2, 0.0f);
            };
            // pass 0: every thread its own queries, within the visit budget; then the cooperative search of the deferred
            // ones; pass 1: the few it could not settle (exact ties, ...) walk the reference traversal to the end.
            unsigned long long* unsettled = wk.keys2 + (size_t)blockIdx.x * ((nd + gridDim.x - 1) / gridDim.x);   // idle until the sort
            for (int pass = 0; pass < 2; pass++) {
                const int count = pass == 0 ? nd : n_unsettled;
                for (int g = pass == 0 ? threadIdx.x * gridDim.x + blockIdx.x : threadIdx.x; g < count; g += (pass == 0 ? gridDim.x : 1) * blockDim.x) {
                    const unsigned long long rec = pass == 0 ? 0ull : unsettled[g];
                    const int i = pass == 0 ? g : (int)(unsigned)rec;
                    if (pass == 0 && plan.nn_budget == 1) { deferred[atomicAdd(&n_deferred, 1)] = i; continue; }   // everything goes to the cooperative search
                    const float4 p = __ldg(data + i);
                    const float qx = __fadd_rn(dot3_ref(R[0], R[1], R[2], p.x, p.y, p.z), t[0]);
                    const float qy = __fadd_rn(dot3_ref(R[3], R[4], R[5], p.x, p.y, p.z), t[1]);
                    const float qz = __fadd_rn(dot3_ref(R[6], R[7], R[8], p.x, p.y, p.z), t[2]);
                    float d2, cap; int pos;
                    if (pass == 0) {
                        const unsigned sp = (unsigned)__ldcg(wk.pos + i) < (unsigned)kd.nm ? (unsigned)__ldcg(wk.pos + i) : 0u;
                        const float4 m = __ldg(leaf + sp);
                        const float e0 = qx - m.x, e1 = qy - m.y, e2 = qz - m.z;
                        cap = e0 * e0 + e1 * e1 + e2 * e2;
                    } else cap = __uint_as_float((unsigned)(rec >> 32));
                    const int id = kd_nearest(kd, nodes, leaf, qx, qy, qz, cap, d2, pass == 0 ? plan.nn_budget : 0, &pos);
                    if (id < 0) { deferred[atomicAdd(&n_deferred, 1)] = i; continue; }
                    emit(i, id, pos, d2, qx, qy, qz);
                }
                if (pass == 1) break;
                __syncthreads();
                const int ndef = n_deferred;
                int* queue = squeue + warp * kCoopQ;
                for (int k = warp; k < ndef; k += kWarps) {
                    const int i = deferred[k];
                    const float4 p = __ldg(data + i);
                    const float qx = __fadd_rn(dot3_ref(R[0], R[1], R[2], p.x, p.y, p.z), t[0]);
                    const float qy = __fadd_rn(dot3_ref(R[3], R[4], R[5], p.x, p.y, p.z), t[1]);
                    const float qz = __fadd_rn(dot3_ref(R[6], R[7], R[8], p.x, p.y, p.z), t[2]);
                    const int sp = (unsigned)__ldcg(wk.pos + i) < (unsigned)kd.nm ? __ldcg(wk.pos + i) : 0;
                    float d2; int pos;
                    const int id = kd_coop_nearest(kd, queue, qx, qy, qz, sp, lane, d2, pos);
                    if (lane == 0) {
                        if (id >= 0) emit(i, id, pos, d2, qx, qy, qz);
                        else unsettled[atomicAdd(&n_unsettled, 1)] = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned)i;
                    }
                    __syncwarp();
                }
                __syncthreads();
            }
            __syncthreads();                                                         // the lists live in wk.order / wk.keys2
        }
        long long c1 = clock64(); c_nn += c1 - c0;
        grid.sync();
        c0 = clock64(); c_wait += c0 - c1;
        // ---- phase B: the reference's qsort by distance (stable for ties) as a distributed rank
        // count: rank(i) = #{j : key_j < key_i}, keys = (d^2 bits, index) are unique.  Every CTA ranks
        // its own queries against all keys; order[rank] = i.
        // (the reference sorts only when do_trim is set, jly_icp3d.hpp:236-239: without it the accumulations below run in
        // data order, i.e. straight over the rows phase A wrote)
        if (!do_sort) {
        } else if (sradix) {
            icp_radix_sort(grid, wk, nd, sradix);        // large clouds: stable LSD radix sort (ends with a grid.sync)
        } else {
            const unsigned long long* keys = wk.keys;
            const bool keys_in_smem = sstage != nullptr && (size_t)plan.stage_bytes >= (size_t)nd * sizeof(unsigned long long);
            if (keys_in_smem) {                      // the staging area is idle until phase C
                unsigned long long* sk = reinterpret_cast<unsigned long long*>(sstage);
                for (int j = threadIdx.x; j < nd; j += blockDim.x) sk[j] = __ldcg(wk.keys + j);
                __syncthreads();
                keys = sk;
            }
            for (int qq = warp; qq * (int)gridDim.x + (int)blockIdx.x < nd; qq += kWarps) {
                const int i = qq * gridDim.x + blockIdx.x;
                const unsigned long long ki = keys[i];
                int cnt = 0;
                if (keys_in_smem) { for (int j = lane; j < nd; j += 32) cnt += keys[j] < ki; }
                else              { for (int j = lane; j < nd; j += 32) cnt += __ldcg(keys + j) < ki; }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
                if (lane == 0) wk.order[cnt] = i;
                if (lane < 2) reinterpret_cast<float4*>(wk.stage)[2 * (size_t)cnt + lane] = __ldcg(reinterpret_cast<const float4*>(wk.q) + 2 * (size_t)i + lane);
            }
            if (keys_in_smem) __syncthreads();       // block 0 overwrites the area in phase C
        }
        c1 = clock64(); c_sort += c1 - c0;
        if (do_sort && !sradix) grid.sync();
        c0 = clock64(); c_wait += c0 - c1;
        if (blockIdx.x == 0) {
            // ---- phase C: reference-order accumulations.  The sort left the correspondences (model
            // point, query, d^2) in wk.stage in sorted order; they are read into shared memory -- all at
            // once when they fit, else streamed in double-buffered chunks of kIcpChunk rows that warps
            // 1.. fetch while warp 0 adds the previous chunk -- and one lane per accumulator adds them
            // up strictly sequentially.
            // The adding lanes (warp 0) run a chain of dependent FADDs, 4 cycles apart when nothing else wants their issue
            // port.  Warps 4, 8 and 12 share that port (warp id mod 4 = SM sub-partition), so they stay out of the staging work:
            // the 12 warps of the other three sub-partitions do it.
            constexpr int kHelpers = (kWarps - 4) * 32;
            const int helper = (warp & 3) != 0 ? (warp - 1 - (warp >> 2)) * 32 + lane : -1;
            float acc = 0.0f;
            if (warp == 0 && lane < 7) acc = lane < 6 ? sh_mu[lane] : 0.0f;                // never reset between iterations (:205-206)
            const float4* srows = reinterpret_cast<const float4*>(do_sort ? wk.stage : wk.q);
            // err_new += dis is float += double in the reference (:254); the double sum of two floats is
            // exact (or differs from either by < 2^-29), so rounding it to float equals the float sum --
            // one add per element, like the other accumulators; operands are fetched 16 at a time ahead
            // of the dependent chain of adds
            // Rows are handed to the adding lanes in chunks of kIcpChunk, TRANSPOSED: column c of the chunk (model x,y,z, query
            // x,y,z, d^2) is contiguous at chunk[buf] + c * kColPitch, so lane c reads four terms per instruction.  Warps 1..
            // fill chunk k+1 (from the staged rows in shared memory, or from global memory for large clouds) while warp 0
            // adds chunk k.
            auto fill = [&](int k, int first, int nthr) {
                const int base = k * kIcpChunk, cnt = min(kIcpChunk, num - base);
                float* cb = chunk[k & 1];
                for (int r = first; r < cnt; r += nthr) {
                    const float4 lo = __ldcg(srows + 2 * (size_t)(base + r)), hi = __ldcg(srows + 2 * (size_t)(base + r) + 1);
                    cb[r] = lo.x; cb[kColPitch + r] = lo.y; cb[2 * kColPitch + r] = lo.z; cb[3 * kColPitch + r] = lo.w;
                    cb[4 * kColPitch + r] = hi.x; cb[5 * kColPitch + r] = hi.y; cb[6 * kColPitch + r] = hi.z;
                }
            };
            {
                const int nchunks = (num + kIcpChunk - 1) / kIcpChunk;
                fill(0, threadIdx.x, blockDim.x);
                __syncthreads();
                const long long ca0 = clock64();
                for (int k = 0; k < nchunks; k++) {
                    if (warp == 0) { if (lane < 7) acc = seq_add_contig(acc, chunk[k & 1] + lane * kColPitch, min(kIcpChunk, num - k * kIcpChunk)); }
                    else if (k + 1 < nchunks && helper >= 0) fill(k + 1, helper, kHelpers);
                    __syncthreads();
                }
                c_acc1 += clock64() - ca0;
            }
            if (warp == 0) {
                if (lane < 7) sh_acc[lane] = acc;
                __syncwarp();
                if (lane == 0) {
                    const float err_new = sh_acc[6];
                    st->err_new = err_new;
                    st->iter = iter;
                    if (sh_err > 0.0f && sh_err - err_new < err_diff * (float)num) { st->converged = 1; sh_conv = 1; }          // :257
                    else {
                        st->err = sh_err = err_new;
                        for (int c = 0; c < 3; c++) { sh_mu[c] = st->mu_m[c] = sh_acc[c] / (float)nd; sh_mu[3 + c] = st->mu_d[c] = sh_acc[3 + c] / (float)nd; }   // :262-263
                    }
                    __threadfence_block();
                }
            }
            __syncthreads();
            c1 = clock64(); c_p1 += c1 - c0; c0 = c1;
            if (!sh_conv) {
                // H = sum over the sorted rows of (q - mu_d)(m - mu_m)^T, each of its 9 entries a strictly
                // sequential float sum (:268).  Warps 1.. form the products of block k+1 into a
                // shared-memory ring while lanes 0..8 of warp 0 add block k: the dependent chain of adds
                // is all that is left on the critical path.
                acc = 0.0f;
                constexpr int kProd = 336, kProdPitch = kProd + 4;   // rows per ring half; 2 * 9 columns of 340 floats fit `chunk`; the pitch keeps the lanes' float4 reads on distinct banks
                static_assert(2 * 9 * kProdPitch <= 2 * kIcpChunk * 8, "product ring exceeds `chunk`");
                float* ring = &chunk[0][0];
                const int nblk = (num + kProd - 1) / kProd;
                const float md0 = sh_mu[3], md1 = sh_mu[4], md2 = sh_mu[5], mm0 = sh_mu[0], mm1 = sh_mu[1], mm2 = sh_mu[2];   // the means just formed (:262-263)
                auto produce = [&](int k) {
                    const int base = k * kProd, cnt = min(kProd, num - base);
                    float* dst = ring + (k & 1) * (kProdPitch * 9);
                    for (int r = helper; r < cnt; r += kHelpers) {
                        const float4 lo = __ldcg(srows + 2 * (size_t)(base + r)), hi = __ldcg(srows + 2 * (size_t)(base + r) + 1);
                        const float q0 = __fsub_rn(lo.w, md0), q1 = __fsub_rn(hi.x, md1), q2 = __fsub_rn(hi.y, md2);
                        const float m0 = __fsub_rn(lo.x, mm0), m1 = __fsub_rn(lo.y, mm1), m2 = __fsub_rn(lo.z, mm2);
                        float* d = dst + r;                                 // entry e of H: column e of the ring half
                        d[0] = __fmul_rn(q0, m0); d[kProdPitch] = __fmul_rn(q0, m1); d[2 * kProdPitch] = __fmul_rn(q0, m2);
                        d[3 * kProdPitch] = __fmul_rn(q1, m0); d[4 * kProdPitch] = __fmul_rn(q1, m1); d[5 * kProdPitch] = __fmul_rn(q1, m2);
                        d[6 * kProdPitch] = __fmul_rn(q2, m0); d[7 * kProdPitch] = __fmul_rn(q2, m1); d[8 * kProdPitch] = __fmul_rn(q2, m2);
                    }
                };
                if (helper >= 0) produce(0);
                __syncthreads();
                for (int k = 0; k < nblk; k++) {
                    if (warp == 0) {
                        if (lane < 9) {
                            const int cnt = min(kProd, num - k * kProd);
                            acc = seq_add_contig(acc, ring + (k & 1) * (kProdPitch * 9) + lane * kProdPitch, cnt);
                        }
                    } else if (k + 1 < nblk && helper >= 0) produce(k + 1);
                    __syncthreads();
                }
                if (warp == 0) {
                    if (lane < 9) sh_H[lane] = acc;
                    __syncwarp();
                    if (lane == 0) { const long long cu0 = clock64(); icp_update(st, sh_H, (flags & 2) != 0, R, t, sh_mu); c_wait += 0; c_svd += clock64() - cu0; }
                }
            }
            c1 = clock64(); c_p2 += c1 - c0;
            if (threadIdx.x == 0) __threadfence();
        }
        grid.sync();
        if (iter == max_iter - 1) { if (blockIdx.x == 0 && threadIdx.x == 0 && !sh_conv) st->iter = max_iter; }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { st->dbg[0] = c_nn; st->dbg[1] = c_wait; st->dbg[2] = c_sort; st->dbg[3] = c_p1; st->dbg[4] = c_p2; st->dbg[5] = c_acc1; st->dbg[0] = c_nn + (c_svd << 32); }
}

// ==========================================================================================
// GOICP_NUM_FAST_ICP: the same iteration -- same queries, same exact nearest neighbours, same convergence test, same
// stale-mean quirk -- with the sums formed in parallel and an original closed 3x3 solver.
//
//  * Moments.  Every CTA adds up, for the queries it answered, 16 numbers: sum m, sum q, sum d^2 and the 9 products
//    sum (q - c_d)(m - c_m)^T about the previous iteration's means c (so nothing cancels when the clouds sit away from
//    the origin); float per thread, shuffle tree per warp, then ONE grid barrier, after which every CTA adds the
//    per-CTA partials of all CTAs in the same fixed order in double and solves for the new pose by itself.  All CTAs
//    execute the same arithmetic on the same numbers, so they agree bit for bit without a second barrier and without a
//    broadcast: one grid.sync per iteration instead of three, nothing sequential but a 16 x gridDim reduction.
//  * Rotation.  R_ = V diag(1,1,det) U^T of H = U W V^T (jly_icp3d.hpp:268-285) is the proper rotation that maximises
//    tr(R_ H^T)...  computed here as: cyclic Jacobi eigen-decomposition of H^T H in double (V, det +1, eigenvalues
//    sorted), u1 = H v1 / |H v1|, u2 = H v2 made orthogonal to u1, u3 = u1 x u2 -- which IS the reference's
//    reflection fix (its third column flips sign exactly when det(V U^T) < 0).
// With trimming (num < nd) the correspondences are sorted by the existing phase B and the moments run over the first
// `num` sorted rows; the sort costs its grid barriers, the rest is the same.
// ==========================================================================================
__device__ void procrustes_jacobi(const double* H /* 3x3 row-major: rows = data, cols = model */, float* Rn)
{
    // A = H^T H (symmetric)
    double A[3][3], V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) A[i][j] = H[0 + i] * H[0 + j] + H[3 + i] * H[3 + j] + H[6 + i] * H[6 + j];
    for (int sweep = 0; sweep < 12; sweep++) {
        const double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
        if (off <= 1e-30 * (fabs(A[0][0]) + fabs(A[1][1]) + fabs(A[2][2])) || off == 0.0) break;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const int p = k == 2 ? 1 : 0, q = k == 0 ? 1 : 2;            // (0,1), (0,2), (1,2)
            if (A[p][q] == 0.0) continue;
            const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
            const double tt = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
            const double c = 1.0 / sqrt(tt * tt + 1.0), sn = tt * c;
            const int r = 3 - p - q;
            const double app = A[p][p], aqq = A[q][q], apq = A[p][q], arp = A[r][p], arq = A[r][q];
            A[p][p] = app - tt * apq; A[q][q] = aqq + tt * apq; A[p][q] = A[q][p] = 0.0;
            A[r][p] = A[p][r] = c * arp - sn * arq; A[r][q] = A[q][r] = sn * arp + c * arq;
#pragma unroll
            for (int i = 0; i < 3; i++) { const double vp = V[i][p], vq = V[i][q]; V[i][p] = c * vp - sn * vq; V[i][q] = sn * vp + c * vq; }
        }
    }
    // eigenvalues descending; an odd permutation is undone by negating a column (keeps det V = +1)
    int o0 = 0, o1 = 1, o2 = 2; bool odd = false;
    if (A[o0][o0] < A[o1][o1]) { const int t = o0; o0 = o1; o1 = t; odd = !odd; }
    if (A[o1][o1] < A[o2][o2]) { const int t = o1; o1 = o2; o2 = t; odd = !odd; }
    if (A[o0][o0] < A[o1][o1]) { const int t = o0; o0 = o1; o1 = t; odd = !odd; }
    double v1[3], v2[3], v3[3];
    for (int i = 0; i < 3; i++) { v1[i] = V[i][o0]; v2[i] = V[i][o1]; v3[i] = odd ? -V[i][o2] : V[i][o2]; }
    // U: images of v1, v2 under H, orthonormalised; u3 completes a right-handed frame
    double u1[3], u2[3], u3[3];
    for (int i = 0; i < 3; i++) { u1[i] = H[3 * i] * v1[0] + H[3 * i + 1] * v1[1] + H[3 * i + 2] * v1[2]; u2[i] = H[3 * i] * v2[0] + H[3 * i + 1] * v2[1] + H[3 * i + 2] * v2[2]; }
    double n1 = sqrt(u1[0] * u1[0] + u1[1] * u1[1] + u1[2] * u1[2]);
    if (n1 > 0.0) { u1[0] /= n1; u1[1] /= n1; u1[2] /= n1; } else { u1[0] = 1.0; u1[1] = u1[2] = 0.0; }
    const double d12 = u1[0] * u2[0] + u1[1] * u2[1] + u1[2] * u2[2];
    for (int i = 0; i < 3; i++) u2[i] -= d12 * u1[i];
    double n2 = sqrt(u2[0] * u2[0] + u2[1] * u2[1] + u2[2] * u2[2]);
    if (n2 > 1e-150) { u2[0] /= n2; u2[1] /= n2; u2[2] /= n2; }
    else {      // rank one: any unit vector orthogonal to u1
        const int a = fabs(u1[0]) <= fabs(u1[1]) && fabs(u1[0]) <= fabs(u1[2]) ? 0 : (fabs(u1[1]) <= fabs(u1[2]) ? 1 : 2);
        double e[3] = {0, 0, 0}; e[a] = 1.0;
        const double d = u1[a];
        for (int i = 0; i < 3; i++) u2[i] = e[i] - d * u1[i];
        n2 = sqrt(u2[0] * u2[0] + u2[1] * u2[1] + u2[2] * u2[2]);
        u2[0] /= n2; u2[1] /= n2; u2[2] /= n2;
    }
    u3[0] = u1[1] * u2[2] - u1[2] * u2[1]; u3[1] = u1[2] * u2[0] - u1[0] * u2[2]; u3[2] = u1[0] * u2[1] - u1[1] * u2[0];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Rn[3 * i + j] = (float)(v1[i] * u1[j] + v2[i] * u2[j] + v3[i] * u3[j]);   // V U^T
}

struct IcpFastShared { float R[9], t[3], mu_m[3], mu_d[3]; float err, err_new; int iter, converged; };

// New pose from the 16 moments of one iteration (see icp_fast_kernel); thread-serial, identical on every CTA / rank.
__device__ void icp_fast_solve(IcpFastShared& cur, const double* tot, int nd, int num, float err_diff, int iter)
{
    const float err_new = (float)tot[6];
    cur.err_new = err_new; cur.iter = iter;
    // jly_icp3d.hpp:257; err_diff < 0 selects fgoicp's relative rule instead (icp3d.cu:96: stop when the improvement is
    // at most |err_diff| of the previous error)
    if (cur.err > 0.0f && (err_diff >= 0.0f ? cur.err - err_new < err_diff * (float)num : cur.err - err_new <= -err_diff * cur.err)) { cur.converged = 1; return; }
    cur.err = err_new;
    // means on top of the previous means, divided by n (the reference never resets them, :205-206, :244-263)
    double mm[3], md[3];
    for (int c = 0; c < 3; c++) { mm[c] = ((double)cur.mu_m[c] + tot[c]) / (double)nd; md[c] = ((double)cur.mu_d[c] + tot[3 + c]) / (double)nd; }
    // H = sum (q - md)(m - mm)^T from the moments about (c_d, c_m) = the previous means
    double H[9];
    for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) {
        const double dd = md[a] - (double)cur.mu_d[a], dm = mm[b] - (double)cur.mu_m[b];
        const double Sq = tot[3 + a] - (double)num * (double)cur.mu_d[a], Sm = tot[b] - (double)num * (double)cur.mu_m[b];
        H[3 * a + b] = tot[7 + 3 * a + b] - dd * Sm - Sq * dm + (double)num * dd * dm;
    }
    float Rn[9];
    procrustes_jacobi(H, Rn);
    float tn[3], tt[3], tmp[9];
    for (int a = 0; a < 3; a++) tn[a] = (float)(mm[a] - ((double)Rn[3 * a] * md[0] + (double)Rn[3 * a + 1] * md[1] + (double)Rn[3 * a + 2] * md[2]));   // t_ = mu_m - R_ mu_d
    mat3_mul(Rn, cur.R, tmp);                                                                   // R = R_ R
    for (int a = 0; a < 3; a++) tt[a] = Rn[3 * a] * cur.t[0] + Rn[3 * a + 1] * cur.t[1] + Rn[3 * a + 2] * cur.t[2] + tn[a];   // t = R_ t + t_
    for (int i = 0; i < 9; i++) cur.R[i] = tmp[i];
    for (int i = 0; i < 3; i++) { cur.t[i] = tt[i]; cur.mu_m[i] = (float)mm[i]; cur.mu_d[i] = (float)md[i]; }
}
// Sharded ICP (the queries of an iteration dealt over W GPUs): every rank runs this on the all-gathered W x 16 moments, added
// in rank order -- the same arithmetic on the same numbers everywhere, so the ranks' poses stay bit-identical.
__global__ void icp_fast_solve_kernel(IcpState* st, const double* __restrict__ xch_all, int W, int nd, int num, float err_diff, int iter, int max_iter)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    IcpFastShared cur;
    for (int i = 0; i < 9; i++) cur.R[i] = st->R[i];
    for (int i = 0; i < 3; i++) { cur.t[i] = st->t[i]; cur.mu_m[i] = st->mu_m[i]; cur.mu_d[i] = st->mu_d[i]; }
    cur.err = st->err; cur.err_new = st->err_new; cur.iter = st->iter; cur.converged = 0;
    double tot[16];
    for (int v = 0; v < 16; v++) { double a = 0.0; for (int r = 0; r < W; r++) a += xch_all[(size_t)r * 16 + v]; tot[v] = a; }
    icp_fast_solve(cur, tot, nd, num, err_diff, iter);
    if (!cur.converged && iter == max_iter - 1) cur.iter = max_iter;
    for (int i = 0; i < 9; i++) st->R[i] = cur.R[i];
    for (int i = 0; i < 3; i++) { st->t[i] = cur.t[i]; st->mu_m[i] = cur.mu_m[i]; st->mu_d[i] = cur.mu_d[i]; }
    st->err = cur.err; st->err_new = cur.err_new; st->iter = cur.iter; st->converged = cur.converged;
}

__global__ void __launch_bounds__(kIcpThreads)
icp_fast_kernel(KdView kd, const float4* __restrict__ data, int nd, IcpState* st, IcpWork wk,
                int max_iter, float err_diff, int num, int do_sort, IcpSmemPlan plan, float* __restrict__ partials /* 2 x gridDim.x x 16 */,
                int q_begin, int q_end, double* __restrict__ xch_out /* shard mode: one NN pass over [q_begin, q_end), its 16 moments go here */)
{
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) unsigned char icp_smem[];
    __shared__ int n_deferred, n_unsettled;
    __shared__ float nn_merge_sm[2 * 48];
    __shared__ IcpFastShared cur;
    __shared__ float red[kIcpThreads / 32][16];
    __shared__ double tot[16];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int kWarps = kIcpThreads / 32;
    const bool sorting = do_sort && num < nd;          // an untrimmed sum does not depend on the order

    const IcpSmem sm = icp_carve(kd, plan, icp_smem, data, q_begin, q_end);
    const KdNode* nodes = sm.nodes; const float4* leaf = sm.leaf;
    float* sstage = sm.sstage; unsigned* sradix = sm.sradix; int* squeue = sm.squeue;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 9; i++) cur.R[i] = st->R[i];
        for (int i = 0; i < 3; i++) { cur.t[i] = st->t[i]; cur.mu_m[i] = st->mu_m[i]; cur.mu_d[i] = st->mu_d[i]; }
        cur.err = st->err; cur.err_new = st->err_new; cur.iter = 0; cur.converged = 0;
    }
    __syncthreads();

    for (int iter = 0; iter < max_iter; iter++) {
        float R[9], t[3];
#pragma unroll
        for (int i = 0; i < 9; i++) R[i] = cur.R[i];
#pragma unroll
        for (int i = 0; i < 3; i++) t[i] = cur.t[i];
        const float cm0 = cur.mu_m[0], cm1 = cur.mu_m[1], cm2 = cur.mu_m[2], cd0 = cur.mu_d[0], cd1 = cur.mu_d[1], cd2 = cur.mu_d[2];
        float acc[16];
#pragma unroll
        for (int k = 0; k < 16; k++) acc[k] = 0.0f;
        auto add_row = [&](float mx, float my, float mz, float qx, float qy, float qz, float d2) {
            acc[0] += mx; acc[1] += my; acc[2] += mz; acc[3] += qx; acc[4] += qy; acc[5] += qz; acc[6] += d2;
            const float a0 = qx - cd0, a1 = qy - cd1, a2 = qz - cd2, b0 = mx - cm0, b1 = my - cm1, b2 = mz - cm2;
            acc[7] += a0 * b0; acc[8] += a0 * b1; acc[9] += a0 * b2;
            acc[10] += a1 * b0; acc[11] += a1 * b1; acc[12] += a1 * b2;
            acc[13] += a2 * b0; acc[14] += a2 * b1; acc[15] += a2 * b2;
        };
        // ---- phase A: transform + exact nearest neighbour (same searches as icp_kernel) ------------------------
        if (plan.brute_force) {
            unsigned long long* park = wk.keys2 + (size_t)blockIdx.x * ((nd + gridDim.x - 1) / gridDim.x);
            icp_nn_small_model(kd, sm, plan.qcache_n, data, q_begin, q_end, R, t, wk.d2, park, nn_merge_sm, [&](int i, int I1, float mx, float my, float mz, float D1, float qx, float qy, float qz) {
                wk.nn[i] = I1;
                if (sorting) {
                    wk.keys[i] = ((unsigned long long)__float_as_uint(D1) << 32) | (unsigned)i;
                    float4* row = reinterpret_cast<float4*>(wk.q) + 2 * (size_t)i;
                    row[0] = make_float4(mx, my, mz, qx);
                    row[1] = make_float4(qy, qz, D1, 0.0f);
                } else add_row(mx, my, mz, qx, qy, qz, D1);
            });
        } else {
            int* deferred = wk.order + (size_t)blockIdx.x * ((nd + gridDim.x - 1) / gridDim.x);
            if (threadIdx.x == 0) { n_deferred = 0; n_unsettled = 0; }
            __syncthreads();
            auto emit = [&](int i, int id, int pos, float d2, float qx, float qy, float qz) {
                wk.nn[i] = id; wk.pos[i] = pos; wk.d2[i] = d2;
                const float mx = __ldg(kd.model + 3 * id), my = __ldg(kd.model + 3 * id + 1), mz = __ldg(kd.model + 3 * id + 2);
                if (sorting) {
                    wk.keys[i] = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned)i;
                    float4* row = reinterpret_cast<float4*>(wk.q) + 2 * (size_t)i;
                    row[0] = make_float4(mx, my, mz, qx);
                    row[1] = make_float4(qy, qz, d2, 0.0f);
                } else add_row(mx, my, mz, qx, qy, qz, d2);
            };
            unsigned long long* unsettled = wk.keys2 + (size_t)blockIdx.x * ((nd + gridDim.x - 1) / gridDim.x);
            for (int pass = 0; pass < 2; pass++) {
                const int count = pass == 0 ? q_end - q_begin : n_unsettled;
                for (int g = pass == 0 ? threadIdx.x * gridDim.x + blockIdx.x : threadIdx.x; g < count; g += (pass == 0 ? gridDim.x : 1) * blockDim.x) {
                    const unsigned long long rec = pass == 0 ? 0ull : unsettled[g];
                    const int i = pass == 0 ? q_begin + g : (int)(unsigned)rec;
                    if (pass == 0 && plan.nn_budget == 1) { deferred[atomicAdd(&n_deferred, 1)] = i; continue; }
                    const float4 p = __ldg(data + i);
                    const float qx = __fadd_rn(dot3_ref(R[0], R[1], R[2], p.x, p.y, p.z), t[0]);
                    const float qy = __fadd_rn(dot3_ref(R[3], R[4], R[5], p.x, p.y, p.z), t[1]);
                    const float qz = __fadd_rn(dot3_ref(R[6], R[7], R[8], p.x, p.y, p.z), t[2]);
                    float d2, cap; int pos;
                    if (pass == 0) {
                        const unsigned sp2 = (unsigned)__ldcg(wk.pos + i) < (unsigned)kd.nm ? (unsigned)__ldcg(wk.pos + i) : 0u;
                        const float4 m = __ldg(leaf + sp2);
                        const float e0 = qx - m.x, e1 = qy - m.y, e2 = qz - m.z;
                        cap = e0 * e0 + e1 * e1 + e2 * e2;
                    } else cap = __uint_as_float((unsigned)(rec >> 32));
                    const int id = kd_nearest(kd, nodes, leaf, qx, qy, qz, cap, d2, pass == 0 ? plan.nn_budget : 0, &pos);
                    if (id < 0) { deferred[atomicAdd(&n_deferred, 1)] = i; continue; }
                    emit(i, id, pos, d2, qx, qy, qz);
                }
                if (pass == 1) break;
                __syncthreads();
                const int ndef = n_deferred;
                int* queue = squeue + warp * kCoopQ;
                for (int k = warp; k < ndef; k += kWarps) {
                    const int i = deferred[k];
                    const float4 p = __ldg(data + i);
                    const float qx = __fadd_rn(dot3_ref(R[0], R[1], R[2], p.x, p.y, p.z), t[0]);
                    const float qy = __fadd_rn(dot3_ref(R[3], R[4], R[5], p.x, p.y, p.z), t[1]);
                    const float qz = __fadd_rn(dot3_ref(R[6], R[7], R[8], p.x, p.y, p.z), t[2]);
                    const int sp2 = (unsigned)__ldcg(wk.pos + i) < (unsigned)kd.nm ? __ldcg(wk.pos + i) : 0;
                    float d2; int pos;
                    const int id = kd_coop_nearest(kd, queue, qx, qy, qz, sp2, lane, d2, pos);
                    if (lane == 0) {
                        if (id >= 0) emit(i, id, pos, d2, qx, qy, qz);
                        else unsettled[atomicAdd(&n_unsettled, 1)] = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned)i;
                    }
                    __syncwarp();
                }
                __syncthreads();
            }
            __syncthreads();
        }
        if (sorting) {
            // ---- phase B (trimming only): the existing sorts leave the correspondences in wk.stage by (d^2, index)
            grid.sync();
            if (sradix) icp_radix_sort(grid, wk, nd, sradix);
            else {
                const unsigned long long* keys = wk.keys;
                const bool keys_in_smem = sstage != nullptr && (size_t)plan.stage_bytes >= (size_t)nd * sizeof(unsigned long long);
                if (keys_in_smem) {
                    unsigned long long* sk = reinterpret_cast<unsigned long long*>(sstage);
                    for (int j = threadIdx.x; j < nd; j += blockDim.x) sk[j] = __ldcg(wk.keys + j);
                    __syncthreads();
                    keys = sk;
                }
                for (int qq = warp; qq * (int)gridDim.x + (int)blockIdx.x < nd; qq += kWarps) {
                    const int i = qq * gridDim.x + blockIdx.x;
                    const unsigned long long ki = keys[i];
                    int cnt = 0;
                    if (keys_in_smem) { for (int j = lane; j < nd; j += 32) cnt += keys[j] < ki; }
                    else              { for (int j = lane; j < nd; j += 32) cnt += __ldcg(keys + j) < ki; }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
                    if (lane < 2) reinterpret_cast<float4*>(wk.stage)[2 * (size_t)cnt + lane] = __ldcg(reinterpret_cast<const float4*>(wk.q) + 2 * (size_t)i + lane);
                }
                __syncthreads();
                grid.sync();
            }
            const float4* srows = reinterpret_cast<const float4*>(wk.stage);
            for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < num; r += gridDim.x * blockDim.x) {
                const float4 lo = __ldcg(srows + 2 * (size_t)r), hi = __ldcg(srows + 2 * (size_t)r + 1);
                add_row(lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z);
            }
        }
        // ---- per-CTA partial moments -> global, one grid barrier, then every CTA reduces all of them by itself ---
        warp_reduce16(acc, lane);
        if ((lane & 1) == 0) red[warp][(lane >> 1) & 15] = acc[0];
        __syncthreads();
        float* mine = partials + ((size_t)(iter & 1) * gridDim.x + blockIdx.x) * 16;
        if (threadIdx.x < 16) {
            float sacc = 0.0f;
#pragma unroll
            for (int w = 0; w < kWarps; w++) sacc += red[w][threadIdx.x];
            __stcg(mine + threadIdx.x, sacc);
        }
        grid.sync();
        if (xch_out && blockIdx.x != 0) return;                 // shard mode: CTA 0 adds this GPU's partials up and hands them to the exchange
        {
            // warp w sums value w over the CTAs: lane-strided doubles, then a fixed shuffle tree
            const float* all = partials + (size_t)(iter & 1) * gridDim.x * 16;
            double sacc = 0.0;
            for (int b = lane; b < (int)gridDim.x; b += 32) sacc += (double)__ldcg(all + (size_t)b * 16 + warp);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, o);
            if (lane == 0) tot[warp] = sacc;
            if (xch_out) { if (lane == 0) xch_out[warp] = sacc; return; }
        }
        __syncthreads();
        if (threadIdx.x == 0) icp_fast_solve(cur, tot, nd, num, err_diff, iter);
        __syncthreads();
        if (cur.converged) break;
        if (iter == max_iter - 1 && threadIdx.x == 0) cur.iter = max_iter;
    }
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (int i = 0; i < 9; i++) st->R[i] = cur.R[i];
        for (int i = 0; i < 3; i++) { st->t[i] = cur.t[i]; st->mu_m[i] = cur.mu_m[i]; st->mu_d[i] = cur.mu_d[i]; }
        st->err = cur.err; st->err_new = cur.err_new; st->iter = cur.iter; st->converged = cur.converged;
    }
}

__global__ void svd3_kernel(const float* __restrict__ H9, int n, float* __restrict__ U9, float* __restrict__ W3, float* __restrict__ V9)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float H[9], U[9], W[3], V[9];
    for (int k = 0; k < 9; k++) H[k] = H9[9 * i + k];
    svd3_ref(H, U, W, V);
    for (int k = 0; k < 9; k++) { U9[9 * i + k] = U[k]; V9[9 * i + k] = V[k]; }
    for (int k = 0; k < 3; k++) W3[3 * i + k] = W[k];
}
cudaError_t launch_svd3(const float* d_H9, int n, float* d_U9, float* d_W3, float* d_V9, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    svd3_kernel<<<(n + 63) / 64, 64, 0, s>>>(d_H9, n, d_U9, d_W3, d_V9);
    return cudaGetLastError();
}

cudaError_t launch_nn(const KdView& kd, const float* d_q, int n, int32_t* d_idx, float* d_d2, bool cooperative, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    if (cooperative) nn_coop_kernel<<<min((n + 3) / 4, 148 * 16), 128, 0, s>>>(kd, d_q, n, d_idx, d_d2);
    else nn_kernel<<<(n + 127) / 128, 128, 0, s>>>(kd, d_q, n, d_idx, d_d2);
    return cudaGetLastError();
}
static IcpSmemPlan icp_plan(const KdView& kd, int n_nodes, int nd, int num, int smem_limit)
{
    // node visits a thread spends on one query before the warp-cooperative search takes it over; 0 = walk to the end, 1 = do
    // not try at all.  Measured (ICP seconds, budgets 0 / 96 / 1): 10 k x 100 k 0.36 / 0.24 / 0.066, 100 k x 1 M 3.1 / 1.8 / 1.2,
    // 1 M x 1 M - / 26.3 / 16.2, bunny 40 k x 40 k 0.034 / 0.031 / 0.026: with 32 divergent walks per warp there is nearly
    // always a lane in a leaf scan or an unwind, so the warp pays the longest branch at every step; the cooperative search
    // keeps the lanes in lock step on one query.  GOICP_NN_BUDGET overrides (the parity tests force each path).
    IcpSmemPlan p = {0, 0, 0, 0, 0, 1, 0, 0, 0};
    (void)num;
    if (const char* f = getenv("GOICP_NN_BUDGET")) p.nn_budget = atoi(f);
    int left = smem_limit;
    // models beyond the linear-scan range are searched through the tree (it cannot fit in shared memory at that size):
    // the cooperative search's per-warp rings come first
    if (kd.nm > 16384) { p.queue_bytes = (kIcpThreads / 32) * kCoopQ * (int)sizeof(int); left -= p.queue_bytes; }
    p.brute_force = kd.nm <= 16384 ? 1 : 0;
    // sort: rank by counting while the cloud is small and its keys can sit in shared memory, radix sort beyond
    // (GOICP_ICP_RADIX=1/0 forces the choice -- used by the parity tests to cover both on the same input)
    const size_t keys = ((size_t)nd * sizeof(unsigned long long) + 15) & ~(size_t)15;
    bool radix = nd > 6144 || keys + 8192 > (size_t)left;
    if (const char* f = getenv("GOICP_ICP_RADIX")) radix = f[0] == '1' || keys > (size_t)left;
    if (radix) { p.radix_bytes = kRadixSmemBytes; left -= kRadixSmemBytes; }
    else { p.stage_bytes = (int)keys; left -= (int)keys; }
    if (p.brute_force) {
        // this CTA's data points + distance hints (at most ceil(nd / 128) + 1 queries per CTA matter, 512 kept)
        p.qcache_n = std::min(512, nd / 128 + 2);
        left -= (int)(((size_t)p.qcache_n * kQCacheEntry + 15) & ~(size_t)15);
        if (kd.grid_start) {
            const size_t grid = (((size_t)(kd.gcells + 1) * sizeof(unsigned short) + 15) & ~(size_t)15) + (size_t)kd.nm * sizeof(float4);
            if (grid <= (size_t)left) { p.grid_bytes = (int)grid; left -= (int)grid; }
        }
        // the kd-tree (walked only by near-tied queries) on chip as well when there is room
        const size_t tree = (size_t)n_nodes * sizeof(KdNode) + (size_t)kd.nm * sizeof(float4);
        if (tree <= (size_t)left) { p.tree_nodes = n_nodes; p.tree_bytes = (int)((tree + 15) / 16 * 16); left -= p.tree_bytes; }
    }
    return p;
}
int icp_threads() { return kIcpThreads; }
int icp_max_blocks_supported() { return kRadixBlockOffs; }
int icp_max_grid_blocks(int device, const KdView& kd, int n_nodes, int nd, int num, int smem_optin, bool fast)
{
    const void* kern = fast ? (const void*)icp_fast_kernel : (const void*)icp_kernel;
    cudaFuncAttributes a;
    if (cudaFuncGetAttributes(&a, kern) != cudaSuccess) return 0;
    const int limit = smem_optin - (int)a.sharedSizeBytes - 1024;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, limit);
    const IcpSmemPlan p = icp_plan(kd, n_nodes, nd, num, limit);
    int per_sm = 0, sms = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kIcpThreads, icp_plan_bytes(p));
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    return per_sm * sms;
}
// one iteration's NN pass of the fast ICP over the queries [q_begin, q_end) (this rank's share); the 16 moments land in d_xch_out
cudaError_t launch_icp_fast_shard(const KdView& kd, int n_nodes, const float4* d_data, int nd, IcpState* d_state, const IcpWork& work,
                                  int num_inliers, int grid_blocks, int smem_optin, float* d_partials, int q_begin, int q_end, double* d_xch_out, cudaStream_t s)
{
    const void* kern = (const void*)icp_fast_kernel;
    cudaFuncAttributes a;
    cudaError_t e = cudaFuncGetAttributes(&a, kern);
    if (e != cudaSuccess) return e;
    const int limit = smem_optin - (int)a.sharedSizeBytes - 1024;
    IcpSmemPlan plan = icp_plan(kd, n_nodes, nd, num_inliers, limit);
    KdView kdv = kd; IcpWork wk = work;
    int max_iter = 1, do_sort = 0; float err_diff = 0.0f;
    void* args[] = {(void*)&kdv, (void*)&d_data, (void*)&nd, (void*)&d_state, (void*)&wk, (void*)&max_iter, (void*)&err_diff, (void*)&num_inliers, (void*)&do_sort,
                    (void*)&plan, (void*)&d_partials, (void*)&q_begin, (void*)&q_end, (void*)&d_xch_out};
    return cudaLaunchCooperativeKernel(kern, dim3(grid_blocks), dim3(kIcpThreads), args, icp_plan_bytes(plan), s);
}
cudaError_t launch_icp_fast_solve(IcpState* d_state, const double* d_xch_all, int W, int nd, int num, float err_diff, int iter, int max_iter, cudaStream_t s)
{
    icp_fast_solve_kernel<<<1, 32, 0, s>>>(d_state, d_xch_all, W, nd, num, err_diff, iter, max_iter);
    return cudaGetLastError();
}
cudaError_t launch_icp(const KdView& kd, int n_nodes, const float4* d_data, int nd, IcpState* d_state, const IcpWork& work,
                       int max_iter, float err_diff, int num_inliers, int do_sort, int grid_blocks, int smem_optin, bool fast, float* d_partials, cudaStream_t s)
{
    const void* kern = fast ? (const void*)icp_fast_kernel : (const void*)icp_kernel;
    cudaFuncAttributes a;
    cudaError_t e = cudaFuncGetAttributes(&a, kern);
    if (e != cudaSuccess) return e;
    const int limit = smem_optin - (int)a.sharedSizeBytes - 1024;
    IcpSmemPlan plan = icp_plan(kd, n_nodes, nd, num_inliers, limit);
    KdView kdv = kd; IcpWork wk = work;
    int q_begin = 0, q_end = nd; double* d_xch = nullptr;               // (the last three only exist in icp_fast_kernel's signature)
    void* args[] = {(void*)&kdv, (void*)&d_data, (void*)&nd, (void*)&d_state, (void*)&wk,
                    (void*)&max_iter, (void*)&err_diff, (void*)&num_inliers, (void*)&do_sort, (void*)&plan, (void*)&d_partials, (void*)&q_begin, (void*)&q_end, (void*)&d_xch};
    return cudaLaunchCooperativeKernel(kern, dim3(grid_blocks), dim3(kIcpThreads), args,
                                       icp_plan_bytes(plan), s);
}

} // namespace goicp
