// bnb_kernels.cu -- the Go-ICP hot path on sm_100a: distance-transform gathers, bound
// evaluation and the persistent translation branch-and-bound.
//
// Replaces, for the GPU: DT3D::Distance (jly_3ddt.cpp:981-1026), the bound-evaluation loop
// body (jly_goicp.cpp:265-336) and GoICP::InnerBnB (jly_goicp.cpp:227-340).  Roofline: random
// 4-byte gathers into the S^3 float grid (L2-resident at S=300, HBM at S=512); no dense
// contraction anywhere, so no tensor cores (DESIGN.md "Kernels").
#include <cstdio>
#include "goicp_types.h"
#include "goicp_kernels.h"
#include "strict_sum.cuh"
#include <cooperative_groups.h>
#include <cstdlib>

namespace goicp {

// ------------------------------------------------------------------------------------------
// DT3D::Distance for a batch of points (+ raw voxel indices for the bit-exactness tests)
// ------------------------------------------------------------------------------------------
__global__ void dt_lookup_kernel(DtView dt, const float* __restrict__ q, int n, float* __restrict__ out, int32_t* __restrict__ idx)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float x = q[3 * i], y = q[3 * i + 1], z = q[3 * i + 2];
    out[i] = dt_distance(dt, x, y, z);
    if (idx) {
        idx[3 * i] = dt_axis_raw(x, dt.xmin, dt.scale);
        idx[3 * i + 1] = dt_axis_raw(y, dt.ymin, dt.scale);
        idx[3 * i + 2] = dt_axis_raw(z, dt.zmin, dt.scale);
    }
}

// ------------------------------------------------------------------------------------------
// block-wide reduction of 16 floats per thread (fixed order => bitwise reproducible)
// red: [kBnbWarps][16] shared; result of value k lands in tot[k] (shared), valid after the
// trailing __syncthreads() for warp 0 only unless sync_all.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void block_reduce16(float (&acc)[16], float (*red)[16], float* tot, int warp, int lane)
{
    warp_reduce16(acc, lane);
    if ((lane & 1) == 0) red[warp][(lane >> 1) & 15] = acc[0];
    __syncthreads();
    if (warp == 0) {
        if (lane < 16) {
            float s = 0.0f;
#pragma unroll
            for (int w = 0; w < kBnbWarps; w++) s += red[w][lane];
            tot[lane] = s;
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
// Generic (rotation, translation cube) pairs: one CTA per pair.  goicp_eval_bounds().
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBnbThreads)
pair_bounds_kernel(BnbConst c, const PairTask* __restrict__ tasks, float2* __restrict__ out)
{
    __shared__ float red[kBnbWarps][16];
    __shared__ float tot[16];
    const PairTask& t = tasks[blockIdx.x];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float w = t.tc[3];
    const float half = w / 2;
    const float tx = __fadd_rn(t.tc[0], half), ty = __fadd_rn(t.tc[1], half), tz = __fadd_rn(t.tc[2], half);
    const float gt = max_trans_dis(w);
    const float cg = t.level >= 0 ? c.cgamma[t.level] : 0.0f;
    float acc[16];
#pragma unroll
    for (int k = 0; k < 16; k++) acc[k] = 0.0f;
    for (int i = threadIdx.x; i < c.nd; i += kBnbThreads) {
        float4 p = __ldg(c.data + i);
        float rx = dot3_ref(t.R[0], t.R[1], t.R[2], p.x, p.y, p.z);
        float ry = dot3_ref(t.R[3], t.R[4], t.R[5], p.x, p.y, p.z);
        float rz = dot3_ref(t.R[6], t.R[7], t.R[8], p.x, p.y, p.z);
        accumulate_point1(c.dt, rx, ry, rz, __fmul_rn(cg, p.w), tx, ty, tz, gt, acc[0], acc[1]);
    }
    block_reduce16(acc, red, tot, warp, lane);
    if (threadIdx.x == 0) out[blockIdx.x] = make_float2(tot[0], tot[1]);
}

// ------------------------------------------------------------------------------------------
// Expansion of translation cubes: one CTA per (rotation, PARENT translation cube); evaluates
// the 8 children exactly like one iteration of InnerBnB's while-loop.  Used by the bench to
// measure the DT-gather roofline in isolation and by the tests; out16 = ub[0..7], lb[0..7].
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBnbThreads)
expand_bounds_kernel(BnbConst c, const PairTask* __restrict__ tasks, float* __restrict__ out16)
{
    __shared__ float red[kBnbWarps][16];
    __shared__ float tot[16];
    const PairTask& t = tasks[blockIdx.x];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float w = t.tc[3] / 2;                     // child width (jly_goicp.cpp:262)
    const float half = w / 2;
    __shared__ float tr[8];
    if (threadIdx.x < 6) {                           // nodeTrans.x = parent.x + bit*w ; transX = nodeTrans.x + w/2
        const int b = threadIdx.x & 1;
        tr[threadIdx.x] = __fadd_rn(__fadd_rn(t.tc[threadIdx.x >> 1], b ? w : 0.0f), half);
    }
    if (threadIdx.x == 6) tr[6] = max_trans_dis(w);
    __syncthreads();
    const float cg = t.level >= 0 ? c.cgamma[t.level] : 0.0f;
    float acc[16];
#pragma unroll
    for (int k = 0; k < 16; k++) acc[k] = 0.0f;
    for (int i = threadIdx.x; i < c.nd; i += kBnbThreads) {
        float4 p = __ldg(c.data + i);
        float rx = dot3_ref(t.R[0], t.R[1], t.R[2], p.x, p.y, p.z);
        float ry = dot3_ref(t.R[3], t.R[4], t.R[5], p.x, p.y, p.z);
        float rz = dot3_ref(t.R[6], t.R[7], t.R[8], p.x, p.y, p.z);
        accumulate_point8(c.dt, rx, ry, rz, __fmul_rn(cg, p.w), tr, acc);
    }
    block_reduce16(acc, red, tot, warp, lane);
    if (threadIdx.x < 16) out16[blockIdx.x * 16 + threadIdx.x] = tot[threadIdx.x];
}

// ------------------------------------------------------------------------------------------
// Translation-BnB heap: libstdc++'s binary-heap algorithms (std::push_heap / std::pop_heap as
// used by std::priority_queue), restated so that nodes with equal (lb, w) are visited in the
// same order as the reference.  First `cap_sm` entries live in shared memory, the rest spill
// to a per-CTA region of global memory.
// ------------------------------------------------------------------------------------------
struct Heap {
    HeapEntry* sm; HeapEntry* gl; int cap_sm; int cap_total; int n;
    __device__ __forceinline__ HeapEntry get(int i) const { return i < cap_sm ? sm[i] : gl[i - cap_sm]; }
    __device__ __forceinline__ void set(int i, const HeapEntry& e) { if (i < cap_sm) sm[i] = e; else gl[i - cap_sm] = e; }
};
// operator< of TRANSNODE (jly_goicp.h:63-71): "a has lower priority than b"
__device__ __forceinline__ bool node_less(const HeapEntry& a, const HeapEntry& b)
{
    return a.lb != b.lb ? a.lb > b.lb : a.level > b.level;
}
__device__ __forceinline__ void heap_sift_up(Heap& h, int hole, const HeapEntry& v)
{
    while (hole > 0) {
        int parent = (hole - 1) / 2;
        HeapEntry p = h.get(parent);
        if (!node_less(p, v)) break;
        h.set(hole, p);
        hole = parent;
    }
    h.set(hole, v);
}
__device__ __forceinline__ bool heap_push(Heap& h, const HeapEntry& v)
{
    if (h.n >= h.cap_total) return false;
    heap_sift_up(h, h.n++, v);
    return true;
}
__device__ __forceinline__ HeapEntry heap_pop(Heap& h)
{
    HeapEntry top = h.get(0);
    int len = --h.n;
    if (len == 0) return top;
    HeapEntry v = h.get(len);
    int hole = 0, child = 0;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        HeapEntry a = h.get(child), b = h.get(child - 1);
        if (node_less(a, b)) { child--; a = b; }
        h.set(hole, a);
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        h.set(hole, h.get(child - 1));
        hole = child - 1;
    }
    heap_sift_up(h, hole, v);
    return top;
}

// Warp-cooperative versions (all 32 lanes call with identical arguments; same resulting array as
// the sequential routines above).  push: every ancestor of the new slot is loaded by its own lane,
// a ballot finds how far the new entry rises, the displaced ancestors move down in one step.
// pop: lane 0 walks the preferred-child path to a leaf (keys only), then the entries on the path
// and the re-inserted last element are placed in parallel.
// Entries move as 16-byte vectors; comparisons need only the first 8 bytes (lb, level).  While the
// queue fits the shared-memory part (the common case) no address-space select is involved.
__device__ __forceinline__ uint4 he_pack(const HeapEntry& e) { return make_uint4(__float_as_uint(e.lb), e.level, e.path_lo, e.path_hi); }
__device__ __forceinline__ HeapEntry he_unpack(const uint4& u) { HeapEntry e; e.lb = __uint_as_float(u.x); e.level = u.y; e.path_lo = u.z; e.path_hi = u.w; return e; }
__device__ __forceinline__ bool key_less(float lb_a, uint32_t lv_a, float lb_b, uint32_t lv_b) { return lb_a != lb_b ? lb_a > lb_b : lv_a > lv_b; }

__device__ __forceinline__ bool wheap_push(Heap& h, const HeapEntry& v, int lane)
{
    if (h.n >= h.cap_total) return false;
    const int n = h.n;
    const int depth = 31 - __clz(n + 1);                 // number of ancestors of slot n
    uint4* A = reinterpret_cast<uint4*>(h.sm);
    const bool in_sm = n < h.cap_sm;
    uint4 e = he_pack(v); bool pred = false;
    if (lane >= 1 && lane <= depth) {
        const int a = ((n + 1) >> lane) - 1;
        e = in_sm ? A[a] : he_pack(h.get(a));
        pred = key_less(__uint_as_float(e.x), e.y, v.lb, v.level);
    }
    const unsigned b = __ballot_sync(0xffffffffu, pred) >> 1;
    const int m = __ffs(~b) - 1;                         // ancestors 1..m are displaced
    if (lane >= 1 && lane <= m) { const int d = ((n + 1) >> (lane - 1)) - 1; if (in_sm) A[d] = e; else h.set(d, he_unpack(e)); }
    if (lane == 0) { const int d = ((n + 1) >> m) - 1; if (in_sm) A[d] = e; else h.set(d, v); }
    h.n = n + 1;
    __syncwarp();
    return true;
}
// Up to NB pushes in one shared-memory round trip (entries e[0..k-1] pushed in that order; same
// resulting array as k wheap_push calls).  Lane l holds, for every new slot, the slot's ancestor at
// tree level l -- all loaded up front.  The pushes are then resolved one after the other in
// registers: a ballot finds how far the new entry rises, every displaced ancestor moves one level
// down its path (one shuffle), and a lane that rewrites an index patches its copies of that index
// held for the later pushes (an index lives at one level, hence in one lane).  Requires the heap to
// stay in shared memory and n >= 7 (then no new slot is an ancestor of another new slot).
template <int NB>
__device__ __forceinline__ void wheap_push_batch(Heap& h, const HeapEntry* e, int k, int lane)
{
    const unsigned full = 0xffffffffu;
    uint4* A = reinterpret_cast<uint4*>(h.sm);
    const int n0 = h.n;
    uint4 anc[NB]; int aidx[NB];
#pragma unroll
    for (int p = 0; p < NB; p++) {
        aidx[p] = -1; anc[p] = make_uint4(0u, 0u, 0u, 0u);
        if (p < k) {
            const int s = n0 + p, D = 31 - __clz(s + 1);
            if (lane < D) { aidx[p] = ((s + 1) >> (D - lane)) - 1; anc[p] = A[aidx[p]]; }
            else if (lane == D) aidx[p] = s;
        }
    }
#pragma unroll
    for (int p = 0; p < NB; p++) {
        if (p < k) {
            const uint4 vv = he_pack(e[p]);
            const int s = n0 + p, D = 31 - __clz(s + 1);
            const bool pred = lane < D && key_less(__uint_as_float(anc[p].x), anc[p].y, e[p].lb, e[p].level);
            const unsigned bits = __ballot_sync(full, pred);
            const int rise = __clz(~(bits << (32 - D)));          // displaced ancestors, counted from the parent upwards
            const int top = D - rise;                              // level the new entry lands on
            uint4 up;
            up.x = __shfl_up_sync(full, anc[p].x, 1); up.y = __shfl_up_sync(full, anc[p].y, 1);
            up.z = __shfl_up_sync(full, anc[p].z, 1); up.w = __shfl_up_sync(full, anc[p].w, 1);
            if (lane >= top && lane <= D) {
                const uint4 w = lane == top ? vv : up;
                A[aidx[p]] = w;
#pragma unroll
                for (int q = p + 1; q < NB; q++) if (aidx[q] == aidx[p]) anc[q] = w;
            }
        }
    }
    h.n = n0 + k;
    __syncwarp();
}
__device__ __forceinline__ HeapEntry wheap_pop(Heap& h, int lane, int* path /* shared, >= 34 ints */)
{
    uint4* A = reinterpret_cast<uint4*>(h.sm);
    const bool in_sm = h.n <= h.cap_sm;
    const HeapEntry top = in_sm ? he_unpack(A[0]) : h.get(0);
    const int len = h.n - 1;
    h.n = len;
    if (len == 0) return top;
    const uint4 v = in_sm ? A[len] : he_pack(h.get(len));
    if (lane == 0) {
        int t = 0, child = 0;
        path[0] = 0;
        const int last_parent = (len - 1) / 2;
        if (in_sm) {
            const uint2* K = reinterpret_cast<const uint2*>(h.sm);          // first 8 bytes of entry i at K[2*i]
            while (child < last_parent) {
                child = 2 * (child + 1);
                const uint2 a = K[2 * child], b = K[2 * (child - 1)];
                if (key_less(__uint_as_float(a.x), a.y, __uint_as_float(b.x), b.y)) child--;
                path[++t] = child;
            }
        } else {
            while (child < last_parent) {
                child = 2 * (child + 1);
                const HeapEntry a = h.get(child), b = h.get(child - 1);
                if (node_less(a, b)) child--;
                path[++t] = child;
            }
        }
        if ((len & 1) == 0 && child == (len - 2) / 2) { child = 2 * (child + 1); path[++t] = child - 1; }
        path[33] = t;
    }
    __syncwarp();
    const int t = path[33];
    uint4 e = v; bool pred = false; int pos = 0;
    if (lane >= 1 && lane <= t) {
        pos = path[lane];
        e = in_sm ? A[pos] : he_pack(h.get(pos));
        pred = key_less(__uint_as_float(e.x), e.y, __uint_as_float(v.x), v.y);
    }
    // the last element climbs from the leaf while the entry above it (path entry i, now one level up) is lower
    const unsigned b = __ballot_sync(0xffffffffu, pred);  // bit i = entry i is lower than v
    int j = t;
    while (j > 0 && ((b >> j) & 1u)) j--;
    if (lane >= 1 && lane <= j) { const int d = path[lane - 1]; if (in_sm) A[d] = e; else h.set(d, he_unpack(e)); }
    if (lane == 0) { const int d = path[j]; if (in_sm) A[d] = v; else h.set(d, he_unpack(v)); }
    __syncwarp();
    return top;
}

// ------------------------------------------------------------------------------------------
// Persistent translation branch-and-bound: ONE THREAD-BLOCK CLUSTER runs one whole
// GoICP::InnerBnB call (jly_goicp.cpp:227-340) -- pop, expand into 8 children, 8*Nd DT gathers,
// reduce, prune, push -- without leaving the chip.
//
// Why a cluster: a scattered 4-byte gather costs one L1TEX sector lookup, and an SM retires about
// one such lookup per clock (measured: the pure gather kernel runs at 0.98 lookups/clk/SM).  One
// expansion step is 8*Nd lookups, so a single CTA needs >= 8*Nd clocks per step and the longest
// inner BnB of a round (hundreds of steps) bounds the whole round.  A cluster of C CTAs on C SMs
// splits the data points C ways: every CTA keeps its slice of the rotated points in its own shared
// memory, gathers for it, and sends 16 partial sums to the leader CTA through distributed shared
// memory; the leader owns the priority queue (libstdc++ heap order) and broadcasts the next cube.
// Two cluster barriers per step.
// ------------------------------------------------------------------------------------------
struct InnerCtrl {
    float tr[8];         // child translations per axis bit (x0,x1,y0,y1,z0,z1) and maxTransDis
    int done;
};
// State only the leader's thread 0 touches; kept in shared memory so that it costs the gathering
// threads no registers.
struct OwnerState {
    float opt_t; float best[4];
    uint32_t pops, evals, max_heap; int status; int heap_n;
    float px, py, pz, cw; uint32_t plevel, ppath_lo, ppath_hi;
    int n_cand; uint32_t flags;
};
constexpr int kMaxCluster = 16;

template <bool PTS_SMEM, bool TRIM>
__global__ void __launch_bounds__(kBnbThreads, 2)
inner_bnb_kernel(BnbConst c, const InnerTask* __restrict__ tasks, InnerResult* __restrict__ results,
                 int heap_cap_sm, HeapEntry* __restrict__ spill, int spill_cap, CandList* __restrict__ cands, unsigned* __restrict__ gkeys)
{
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int C = (int)cluster.num_blocks();
    const int rank = (int)cluster.block_rank();
    const int task_id = blockIdx.x / C;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ float red[kBnbWarps][16];
    __shared__ float tot[16];
    __shared__ InnerCtrl ctrl;                       // written by the leader into every CTA of the cluster
    __shared__ float partials[kMaxCluster][16];      // leader: one row per CTA
    __shared__ OwnerState own;
    __shared__ float4 cand_node[kMaxCand];
    __shared__ float cand_ub[kMaxCand];
    // trimming (inlier_num < nd): cluster-wide radix select of the inlier_num smallest residuals per child
    __shared__ int hist[8][256];
    __shared__ unsigned sel_prefix[8];
    __shared__ int sel_k[8];

    HeapEntry* hsm = reinterpret_cast<HeapEntry*>(smem_raw);
    float4* pts = reinterpret_cast<float4*>(smem_raw + (size_t)heap_cap_sm * sizeof(HeapEntry));

    const long long t_begin = clock64();
    const InnerTask& task = tasks[task_id];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool leader = rank == 0;
    const bool ub_pass = task.level < 0;
    // |sequential float sum - tree float sum| <= ~Nd*2^-24 relative; contenders within twice that
    const float cand_eps = fminf(1e-2f, 2.0f * (float)c.nd * 5.9604645e-8f) + 1e-6f;
    // this CTA's slice of the data points
    const int per = (c.nd + C - 1) / C;
    const int p_begin = min(rank * per, c.nd), p_end = min(p_begin + per, c.nd);
    const int np = p_end - p_begin;
    constexpr bool trim = TRIM;
    // residual keys of this CTA's points for the 8 children (trimming only): in shared memory after the points (if staged),
    // or -- clouds whose slice of keys does not fit next to the queue -- in this CTA's slab of global memory (L2-resident
    // between the passes of the select)
    unsigned* mk = gkeys ? gkeys + (size_t)blockIdx.x * per * 8
                         : reinterpret_cast<unsigned*>(smem_raw + (size_t)heap_cap_sm * sizeof(HeapEntry) + (PTS_SMEM ? (size_t)per * sizeof(float4) : 0));

    if (PTS_SMEM) {
        // rotate once (jly_goicp.cpp:470-476) and keep (R p, gamma) on chip for the whole search
        const float cg_ = task.level >= 0 ? c.cgamma[task.level] : 0.0f;
        for (int i = tid; i < np; i += kBnbThreads) {
            float4 p = __ldg(c.data + p_begin + i);
            pts[i] = make_float4(dot3_ref(task.R[0], task.R[1], task.R[2], p.x, p.y, p.z), dot3_ref(task.R[3], task.R[4], task.R[5], p.x, p.y, p.z),
                                 dot3_ref(task.R[6], task.R[7], task.R[8], p.x, p.y, p.z), __fmul_rn(cg_, p.w));
        }
    }
    HeapEntry* const spill_mine = spill + (size_t)task_id * spill_cap;
    if (leader && tid == 0) {
        own.opt_t = task.opt_error;
        own.best[0] = own.best[1] = own.best[2] = own.best[3] = 0.0f;
        own.pops = own.evals = own.max_heap = 0; own.status = 0; own.n_cand = 0; own.flags = 0;
        Heap heap; heap.sm = hsm; heap.gl = spill_mine; heap.cap_sm = heap_cap_sm; heap.cap_total = heap_cap_sm + spill_cap; heap.n = 0;
        HeapEntry root; root.lb = 0.0f; root.level = 0; root.path_lo = 0; root.path_hi = 0;   // initNodeTrans.lb = 0 (jly_goicp.cpp:63)
        heap_push(heap, root);
        own.heap_n = heap.n;
    }

    for (;;) {
        if (leader && tid == 0) {
            Heap heap; heap.sm = hsm; heap.gl = spill_mine; heap.cap_sm = heap_cap_sm; heap.cap_total = heap_cap_sm + spill_cap; heap.n = own.heap_n;
            InnerCtrl next; next.done = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) next.tr[k] = 0.0f;
            if (own.status) next.done = 1;                                        // heap capacity exceeded: give up loudly
            else if (heap.n == 0) next.done = 1;                                  // :243-244
            else {
                HeapEntry e = heap_pop(heap);
                own.pops++;                                                       // tNodeCount++ (:248)
                if (__fsub_rn(own.opt_t, e.lb) < c.sse_thresh) next.done = 1;     // :257
                else if (e.level >= (uint32_t)kMaxTransLevel) { next.done = 1; own.status = 4; }
                else {
                    // rebuild the cube corner by replaying the reference's float additions
                    // parent.x + (j&1)*w down the octant path (:267-269)
                    float x = c.tx, y = c.ty, z = c.tz, w = c.tw;
                    unsigned long long path = ((unsigned long long)e.path_hi << 32) | e.path_lo;
                    for (uint32_t l = 0; l < e.level; l++) {
                        w = w / 2;
                        unsigned b = (unsigned)(path >> (3 * l)) & 7u;
                        x = __fadd_rn(x, (b & 1) ? w : 0.0f);
                        y = __fadd_rn(y, (b & 2) ? w : 0.0f);
                        z = __fadd_rn(z, (b & 4) ? w : 0.0f);
                    }
                    const float cw = w / 2;                                       // nodeTrans.w = parent.w/2 (:262)
                    own.px = x; own.py = y; own.pz = z; own.cw = cw;
                    own.plevel = e.level; own.ppath_lo = e.path_lo; own.ppath_hi = e.path_hi;
                    const float half = cw / 2;
#pragma unroll
                    for (int b = 0; b < 2; b++) {
                        next.tr[b] = __fadd_rn(__fadd_rn(x, b ? cw : 0.0f), half);
                        next.tr[2 + b] = __fadd_rn(__fadd_rn(y, b ? cw : 0.0f), half);
                        next.tr[4 + b] = __fadd_rn(__fadd_rn(z, b ? cw : 0.0f), half);
                    }
                    next.tr[6] = max_trans_dis(cw);
                }
            }
            own.heap_n = heap.n;
            for (int r = 0; r < C; r++) *cluster.map_shared_rank(&ctrl, r) = next;   // broadcast through DSMEM
        }
        cluster.sync();
        if (ctrl.done) break;

        // ---- 8 x (Nd / C) distance-transform gathers ------------------------------------
        float acc[16];
#pragma unroll
        for (int k = 0; k < 16; k++) acc[k] = 0.0f;
        if (!trim) {
            if (PTS_SMEM) {
                for (int i = tid; i < np; i += kBnbThreads) {
                    const float4 p = pts[i];
                    accumulate_point8(c.dt, p.x, p.y, p.z, p.w, ctrl.tr, acc);
                }
            } else {
                const float cg_ = task.level >= 0 ? c.cgamma[task.level] : 0.0f;
                for (int i = p_begin + tid; i < p_end; i += kBnbThreads) {
                    const float4 p = __ldg(c.data + i);
                    accumulate_point8(c.dt, dot3_ref(task.R[0], task.R[1], task.R[2], p.x, p.y, p.z), dot3_ref(task.R[3], task.R[4], task.R[5], p.x, p.y, p.z),
                                      dot3_ref(task.R[6], task.R[7], task.R[8], p.x, p.y, p.z), __fmul_rn(cg_, p.w), ctrl.tr, acc);
                }
            }
        } else {
            // Trimmed bounds (jly_goicp.cpp:293-315): only the inlier_num smallest residuals of each
            // child count.  Residuals go to shared memory as order-preserving keys (non-negative
            // floats), the k-th smallest is found by a 4 x 8-bit radix select whose histograms are
            // merged in the leader through DSMEM atomics, then everything below the threshold is
            // summed and the leader adds the ties it still needs.
            const float cg_ = task.level >= 0 ? c.cgamma[task.level] : 0.0f;
            for (int i = tid; i < np; i += kBnbThreads) {
                float px, py, pz, gm;
                if (PTS_SMEM) { const float4 p = pts[i]; px = p.x; py = p.y; pz = p.z; gm = p.w; }
                else {
                    const float4 p = __ldg(c.data + p_begin + i);
                    px = dot3_ref(task.R[0], task.R[1], task.R[2], p.x, p.y, p.z); py = dot3_ref(task.R[3], task.R[4], task.R[5], p.x, p.y, p.z);
                    pz = dot3_ref(task.R[6], task.R[7], task.R[8], p.x, p.y, p.z); gm = __fmul_rn(cg_, p.w);
                }
                float m[8];
                point_residuals8(c.dt, px, py, pz, gm, ctrl.tr, m);
#pragma unroll
                for (int j = 0; j < 8; j++) mk[j * per + i] = __float_as_uint(m[j]);
            }
            if (tid < 8) { sel_prefix[tid] = 0u; sel_k[tid] = c.inlier_num; }
            for (int pass = 0; pass < 4; pass++) {
                const int shift = 24 - 8 * pass;
                __syncthreads();                                  // the leader's scan of the previous pass is done with hist
                for (int b = tid; b < 8 * 256; b += kBnbThreads) (&hist[0][0])[b] = 0;
                cluster.sync();                                   // histograms zeroed everywhere, keys / prefixes visible
                const unsigned himask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
                for (int i = tid; i < np; i += kBnbThreads)
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const unsigned key = mk[j * per + i];
                        if ((key & himask) == sel_prefix[j]) atomicAdd(&hist[j][(key >> shift) & 255u], 1);
                    }
                __syncthreads();
                if (!leader) {                                    // merge into the leader's histogram
                    int* lh = cluster.map_shared_rank(&hist[0][0], 0);
                    for (int b = tid; b < 8 * 256; b += kBnbThreads) { const int v = (&hist[0][0])[b]; if (v) atomicAdd(lh + b, v); }
                }
                cluster.sync();
                if (leader && warp < 8) {                         // warp j: find the digit holding the k-th smallest of child j
                    const int j = warp;
                    int carry = 0, found = -1, kk = sel_k[j];
                    for (int base = 0; base < 256 && found < 0; base += 32) {
                        const int v = hist[j][base + lane];
                        int incl = v;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
                        const unsigned hit = __ballot_sync(0xffffffffu, carry + incl >= kk);
                        if (hit) {
                            const int l0 = __ffs(hit) - 1;
                            const int before = carry + __shfl_sync(0xffffffffu, incl - v, l0);
                            found = base + l0;
                            if (lane == 0) {
                                const unsigned np_ = sel_prefix[j] | ((unsigned)found << shift);
                                const int nk = kk - before;
                                for (int r = 0; r < C; r++) { *cluster.map_shared_rank(&sel_prefix[j], r) = np_; *cluster.map_shared_rank(&sel_k[j], r) = nk; }
                            }
                        } else carry += __shfl_sync(0xffffffffu, incl, 31);
                    }
                }
                // the next pass's first cluster.sync() publishes the new prefixes
            }
            cluster.sync();
            // sel_prefix[j] is now the threshold key T_j, sel_k[j] the number of values == T_j to include
            for (int i = tid; i < np; i += kBnbThreads)
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const unsigned key = mk[j * per + i];
                    if (key < sel_prefix[j]) {
                        const float m = __uint_as_float(key);
                        acc[j] = __fadd_rn(acc[j], __fmul_rn(m, m));
                        const float e = __fsub_rn(m, ctrl.tr[6]);
                        if (e > 0.0f) acc[8 + j] = __fadd_rn(acc[8 + j], __fmul_rn(e, e));
                    }
                }
        }
        block_reduce16(acc, red, tot, warp, lane);      // contains a __syncthreads(); tot valid in warp 0
        if (warp == 0 && lane < 16) cluster.map_shared_rank(&partials[0][0], 0)[rank * 16 + lane] = tot[lane];
        cluster.sync();

        // ---- sequential bookkeeping of the 8 children (jly_goicp.cpp:317-336) -------------
        if (leader && tid == 0) {
            Heap heap; heap.sm = hsm; heap.gl = spill_mine; heap.cap_sm = heap_cap_sm; heap.cap_total = heap_cap_sm + spill_cap; heap.n = own.heap_n;
            float opt_t = own.opt_t;
            const float px = own.px, py = own.py, pz = own.pz, cw = own.cw;
            own.evals += 8;
            for (int j = 0; j < 8; j++) {
                float ub = 0.0f, lb = 0.0f;                 // fixed-order sum over the cluster's CTAs
                for (int r = 0; r < C; r++) { ub += partials[r][j]; lb += partials[r][8 + j]; }
                if (trim) {                                 // the ties at the trimming threshold that still count
                    const float tv = __uint_as_float(sel_prefix[j]), cnt = (float)sel_k[j];
                    ub += cnt * (tv * tv);
                    const float e = __fsub_rn(tv, ctrl.tr[6]);
                    if (e > 0.0f) lb += cnt * (e * e);
                }
                if (ub < opt_t) {
                    opt_t = ub;
                    own.best[0] = __fadd_rn(px, (j & 1) ? cw : 0.0f);
                    own.best[1] = __fadd_rn(py, (j & 2) ? cw : 0.0f);
                    own.best[2] = __fadd_rn(pz, (j & 4) ? cw : 0.0f);
                    own.best[3] = cw;
                }
                if (ub_pass && ub <= opt_t * (1.0f + cand_eps)) {
                    int n_cand = own.n_cand;
                    if (n_cand == kMaxCand) {           // drop contenders the running minimum has left behind
                        int k = 0;
                        for (int q = 0; q < n_cand; q++)
                            if (cand_ub[q] <= opt_t * (1.0f + cand_eps)) { cand_node[k] = cand_node[q]; cand_ub[k] = cand_ub[q]; k++; }
                        n_cand = k;
                    }
                    if (n_cand < kMaxCand) {
                        cand_node[n_cand] = make_float4(__fadd_rn(px, (j & 1) ? cw : 0.0f), __fadd_rn(py, (j & 2) ? cw : 0.0f),
                                                        __fadd_rn(pz, (j & 4) ? cw : 0.0f), cw);
                        cand_ub[n_cand] = ub; n_cand++;
                    } else own.flags |= 1u;             // could not keep every contender
                    own.n_cand = n_cand;
                }
                if (lb >= opt_t) continue;
                if ((int)own.plevel + 1 > c.trans_cutoff_level) continue;          // span cut-off (fgoicp-style search): evaluated, not subdivided
                HeapEntry e; e.lb = lb; e.level = own.plevel + 1;
                unsigned long long path = (((unsigned long long)own.ppath_hi << 32) | own.ppath_lo) | ((unsigned long long)j << (3 * own.plevel));
                e.path_lo = (uint32_t)path; e.path_hi = (uint32_t)(path >> 32);
                if (!heap_push(heap, e)) { own.status = 3; break; }
            }
            own.opt_t = opt_t;
            own.heap_n = heap.n;
            if ((uint32_t)heap.n > own.max_heap) own.max_heap = heap.n;
        }
        // a failed push (status != 0) ends the task at the top of the next iteration
    }
    if (!leader) return;

    // ---- results.  The search used fixed-order tree sums.  For an upper-bound pass the cubes
    // whose sum is within rounding distance of the minimum are handed to the host: if this call
    // turns out to improve the global optimum, strict_eval_kernel re-evaluates exactly those in
    // the reference's summation order to settle the arg-min (strict_sum.cuh).
    if (ub_pass) {
        if (tid == 0) {
            int k = 0;
            for (int q = 0; q < own.n_cand; q++)
                if (cand_ub[q] <= own.opt_t * (1.0f + cand_eps)) { cand_node[k] = cand_node[q]; cand_ub[k] = cand_ub[q]; k++; }
            own.n_cand = k;
        }
        __syncthreads();
        CandList& cl = cands[task_id];
        for (int q = tid; q < own.n_cand; q += kBnbThreads) { cl.node[q] = cand_node[q]; cl.ub[q] = cand_ub[q]; }
        if (tid == 0) { cl.n = own.n_cand; cl.flags = own.flags; cl.final_fast = own.opt_t; cl.eps = cand_eps; }
    }
    if (tid == 0) {
        InnerResult r;
        r.value = own.opt_t; r.node[0] = own.best[0]; r.node[1] = own.best[1]; r.node[2] = own.best[2]; r.node[3] = own.best[3];
        r.pops = own.pops; r.evals = own.evals; r.status = own.status; r.max_heap = own.max_heap;
        r.pad[0] = own.flags; r.pad[1] = ub_pass ? (uint32_t)own.n_cand : 0u;
        r.kcycles = (uint32_t)((clock64() - t_begin) >> 10);
        r.reuse_gt = 3.402823466e+38f; r.reuse_poplb = 0.0f;       // this kernel does not track it: never reusable
        results[task_id] = r;
    }
}

// ------------------------------------------------------------------------------------------
// Pipelined variant (no trimming): warp 0 of every CTA is the OWNER warp and does not gather; in
// the leader CTA it runs the priority queue.  The cluster barrier is split into arrive / wait so
// that the leader's queue maintenance (8 pushes + 1 pop in libstdc++ order) happens in the SHADOW
// of the cluster's gathers for the next cube:
//     owner:    decide next cube -> write ctrl to all CTAs -> arrive(A) ; queue maintenance ; wait(A) ;
//               arrive(B) ; wait(B) -> partial sums are in -> decide ...
//     gatherers:            arrive(A) ; wait(A) ; gather + reduce + send partials ; arrive(B) ; wait(B)
// "Decide" picks the node the queue WILL pop next -- the better of the current top and the children
// about to be pushed -- which is unambiguous unless two candidates tie on (lb, level); on a tie the
// maintenance runs first (the tie is then resolved by the heap exactly like in the reference).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
// arrival that publishes nothing (no fence needed on the arriving side)
__device__ __forceinline__ void cluster_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }

// LOWLAT: for rounds bound by their longest inner BnB (not by throughput) the kernel trades the second resident CTA
// for 128 registers, so that the batched queue maintenance of the owner warp stays out of local memory (local traffic
// would queue behind the gathers in L1TEX).  Measured on the bunny config: the longest task of a round gets 1.4x
// faster, the sum of all task cycles 1.5x smaller, with half the resident clusters; the engine picks per round.
template <bool PTS_SMEM, bool LOWLAT, int THREADS = kBnbThreads, int MINB = (LOWLAT ? 1 : 2)>
__global__ void __launch_bounds__(THREADS, MINB)
inner_bnb_pipelined_kernel(BnbConst c, const InnerTask* __restrict__ tasks, InnerResult* __restrict__ results,
                           int heap_cap_sm, HeapEntry* __restrict__ spill, int spill_cap, CandList* __restrict__ cands)
{
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int C = (int)cluster.num_blocks();
    const int rank = (int)cluster.block_rank();
    const int task_id = blockIdx.x / C;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int kVirtWarps = kBnbWarps - 1, kGW = THREADS / 32 - 1;     // virtual / real gather warps
    static_assert(kVirtWarps % kGW == 0, "the real gather warps must share the 15 virtual ones evenly");
    __shared__ float red[kVirtWarps][16];
    __shared__ InnerCtrl ctrl;                       // written by the leader into every CTA of the cluster
    __shared__ float partials[kMaxCluster][16];      // leader: one row per CTA
    __shared__ float4 cand_node[kMaxCand];
    __shared__ float cand_ub[kMaxCand];

    HeapEntry* hsm = reinterpret_cast<HeapEntry*>(smem_raw);
    float4* pts = reinterpret_cast<float4*>(smem_raw + (size_t)heap_cap_sm * sizeof(HeapEntry));

    const long long t_begin = clock64();
    const InnerTask& task = tasks[task_id];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool leader = rank == 0;
    const bool ub_pass = task.level < 0;
    const float cand_eps = fminf(1e-2f, 2.0f * (float)c.nd * 5.9604645e-8f) + 1e-6f;
    const int per = (c.nd + C - 1) / C;
    const int p_begin = min(rank * per, c.nd), p_end = min(p_begin + per, c.nd);
    const int np = p_end - p_begin;

    if (PTS_SMEM) {
        const float cg_ = task.level >= 0 ? c.cgamma[task.level] : 0.0f;
        for (int i = tid; i < np; i += THREADS) {
            float4 p = __ldg(c.data + p_begin + i);
            pts[i] = make_float4(dot3_ref(task.R[0], task.R[1], task.R[2], p.x, p.y, p.z), dot3_ref(task.R[3], task.R[4], task.R[5], p.x, p.y, p.z),
                                 dot3_ref(task.R[6], task.R[7], task.R[8], p.x, p.y, p.z), __fmul_rn(cg_, p.w));
        }
    }
    __syncthreads();

    if (warp != 0) {
        // ================================ gather warps ======================================
        long long g_waitA = 0, g_gather = 0, g_reduce = 0, g_waitB = 0;      // cycle breakdown (reported when c.dbg is set)
        for (;;) {
            long long g0 = clock64();
            cluster_arrive_relaxed(); cluster_wait();               // (A) next cube published
            if (ctrl.done) break;
            long long g1 = clock64(); g_waitA += g1 - g0;
            // The sums are formed by kVirtWarps = 15 VIRTUAL gather warps whatever the CTA size: virtual thread g = 32 * vw + lane
            // adds up the points g, g + 480, ... and its warp reduces by shuffles; a CTA with fewer real gather warps (5 or 3,
            // the dense variants) lets each of them play 3 or 5 virtual warps in turn.  Same numbers in the same order in every
            // variant, so the engine's per-round choice of variant (a function of measured cycles) never changes a result.
#pragma unroll 1
            for (int v = 0; v < kVirtWarps / kGW; v++) {
                const int vw = warp - 1 + v * kGW;
                const int vgt = vw * 32 + lane;
                float acc[16];
#pragma unroll
                for (int k = 0; k < 16; k++) acc[k] = 0.0f;
                if (PTS_SMEM) {
                    for (int i = vgt; i < np; i += kVirtWarps * 32) {
                        const float4 p = pts[i];
                        accumulate_point8(c.dt, p.x, p.y, p.z, p.w, ctrl.tr, acc);
                    }
                } else {
                    const float cg_ = task.level >= 0 ? c.cgamma[task.level] : 0.0f;
                    for (int i = p_begin + vgt; i < p_end; i += kVirtWarps * 32) {
                        const float4 p = __ldg(c.data + i);
                        accumulate_point8(c.dt, dot3_ref(task.R[0], task.R[1], task.R[2], p.x, p.y, p.z), dot3_ref(task.R[3], task.R[4], task.R[5], p.x, p.y, p.z),
                                          dot3_ref(task.R[6], task.R[7], task.R[8], p.x, p.y, p.z), __fmul_rn(cg_, p.w), ctrl.tr, acc);
                    }
                }
                if (kGW == kVirtWarps) { g0 = clock64(); g_gather += g0 - g1; }
                warp_reduce16(acc, lane);
                if ((lane & 1) == 0) red[vw][(lane >> 1) & 15] = acc[0];
            }
            if (kGW != kVirtWarps) { g0 = clock64(); g_gather += g0 - g1; }
            asm volatile("bar.sync 1, %0;" :: "r"(THREADS - 32) : "memory");
            if (warp == 1 && lane < 16) {
                float s = 0.0f;
#pragma unroll
                for (int w = 0; w < kVirtWarps; w++) s += red[w][lane];
                cluster.map_shared_rank(&partials[0][0], 0)[rank * 16 + lane] = s;
            }
            g1 = clock64(); g_reduce += g1 - g0;
            cluster_arrive(); cluster_wait();                       // (B) partial sums delivered
            g_waitB += clock64() - g1;
        }
        if (c.dbg && leader && warp == 1 && lane == 0) {
            unsigned long long* d = c.dbg + (size_t)task_id * 12;
            d[4] = g_waitA; d[5] = g_gather; d[6] = g_reduce; d[7] = g_waitB;
        }
        return;
    }
    if (!leader) {
        // ============================ idle owner warp of a helper CTA ========================
        for (;;) {
            cluster_arrive_relaxed(); cluster_wait();
            if (ctrl.done) break;
            cluster_arrive_relaxed(); cluster_wait();
        }
        return;
    }

    // ===================================== leader's owner warp =============================
    // All 32 lanes run the same scalar bookkeeping redundantly (no divergence); the lanes share the
    // work where there is any: summing the CTAs' partials, heap operations, the ctrl broadcast.
    __shared__ float tot16[16];
    __shared__ int pop_path[34];
    Heap heap; heap.sm = hsm; heap.gl = spill + (size_t)task_id * spill_cap; heap.cap_sm = heap_cap_sm; heap.cap_total = heap_cap_sm + spill_cap; heap.n = 0;
    float opt_t = task.opt_error;
    float best[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    uint32_t pops = 0, evals = 0, max_heap = 0, flags = 0; int status = 0, n_cand = 0;
    float px = c.tx, py = c.ty, pz = c.tz, cw = c.tw / 2;              // cube being expanded
    uint32_t plevel = 0; unsigned long long ppath = 0ull;
    // pushes deferred into the shadow of the gathers: children pmask of one parent; their lower bounds stay in
    // tot16[8..15] (not rewritten before the flush), level and path are rebuilt from the parent's -- no local array
    uint32_t pmask = 0, pend_level = 0; unsigned long long pend_base = 0ull; int pend_shift = 0; bool need_pop = false;
    HeapEntry expect; expect.lb = 0.0f; expect.level = 0; expect.path_lo = expect.path_hi = 0;
    bool done = false;
    // validity range of the result under a smaller initial optErrorT (InnerResult::reuse_*): tracked while opt_t is still the initial value
    bool first_phase = true; float reuse_gt = 0.0f, reuse_poplb = 0.0f;

    auto publish = [&](bool fin) {                                       // lane r writes the control block of CTA r
        if (lane < C) {
            InnerCtrl next; next.done = fin ? 1 : 0;
            const float half = cw / 2;
#pragma unroll
            for (int b = 0; b < 2; b++) {
                next.tr[b] = __fadd_rn(__fadd_rn(px, b ? cw : 0.0f), half);
                next.tr[2 + b] = __fadd_rn(__fadd_rn(py, b ? cw : 0.0f), half);
                next.tr[4 + b] = __fadd_rn(__fadd_rn(pz, b ? cw : 0.0f), half);
            }
            next.tr[6] = max_trans_dis(cw); next.tr[7] = 0.0f;
            *cluster.map_shared_rank(&ctrl, lane) = next;
        }
    };
    // adopt `e` as the cube to expand next (the reference's pop, :246-262), or finish
    auto adopt = [&](const HeapEntry& e) {
        pops++;                                                               // tNodeCount++ (:248)
        if (__fsub_rn(opt_t, e.lb) < c.sse_thresh) { done = true; return; }  // :257
        if (first_phase) reuse_poplb = fmaxf(reuse_poplb, e.lb);
        if (e.level >= (uint32_t)kMaxTransLevel) { done = true; status = 4; return; }
        // corner: replay the reference's float additions parent.x + (j&1)*w down the octant path (:267-269)
        float x = c.tx, y = c.ty, z = c.tz, w = c.tw;
        const unsigned long long path = ((unsigned long long)e.path_hi << 32) | e.path_lo;
        for (uint32_t l = 0; l < e.level; l++) {
            w = w / 2;
            const unsigned b = (unsigned)(path >> (3 * l)) & 7u;
            x = __fadd_rn(x, (b & 1) ? w : 0.0f); y = __fadd_rn(y, (b & 2) ? w : 0.0f); z = __fadd_rn(z, (b & 4) ? w : 0.0f);
        }
        px = x; py = y; pz = z; cw = w / 2; plevel = e.level; ppath = path;
    };
    auto pend_entry = [&](int j) {
        HeapEntry e; e.lb = tot16[8 + j]; e.level = pend_level;
        const unsigned long long path = pend_base | ((unsigned long long)j << pend_shift);
        e.path_lo = (uint32_t)path; e.path_hi = (uint32_t)(path >> 32);
        return e;
    };
    auto flush_pending = [&]() {
        uint32_t m = pmask;
        // common case: the queue stays in shared memory -> the pushes go four at a time, one shared-memory round trip each
        while (LOWLAT && m && !status && heap.n >= 7 && heap.n + 4 <= heap.cap_sm) {
            HeapEntry e4[4]; int k = 0;
#pragma unroll
            for (int u = 0; u < 4; u++) if (m) { e4[u] = pend_entry(__ffs(m) - 1); m &= m - 1; k = u + 1; }
            wheap_push_batch<4>(heap, e4, k, lane);
        }
        for (; m && !status; m &= m - 1) if (!wheap_push(heap, pend_entry(__ffs(m) - 1), lane)) status = 3;
        pmask = 0;
        if ((uint32_t)heap.n > max_heap) max_heap = heap.n;
    };

    {   // push + pop of initNodeTrans (:241-247)
        HeapEntry root; root.lb = 0.0f; root.level = 0; root.path_lo = root.path_hi = 0;
        adopt(root);
        publish(done);
    }
    long long o_maint = 0, o_waitA = 0, o_waitB = 0, o_book = 0, o_arrive = 0, o_push = 0, o_npush = 0, o_pop = 0;
    for (;;) {
        __syncwarp();
        long long o0 = clock64();
        cluster_arrive();                                                     // (A) cube published
        const long long oa = clock64(); o_arrive += oa - o0;
        if (!done) {
            // ---- queue maintenance in the shadow of the gathers --------------------------
            o_npush += __popc(pmask);
            flush_pending();
            const long long ob = clock64();
            o_push += ob - oa;
            if (need_pop && !status) {
                const HeapEntry e = wheap_pop(heap, lane, pop_path);
                o_pop += clock64() - ob;
                if (e.lb != expect.lb || e.level != expect.level || e.path_lo != expect.path_lo || e.path_hi != expect.path_hi) status = 5;   // cannot happen: the prediction was unambiguous
                need_pop = false;
            }
        }
        long long o1 = clock64(); o_maint += o1 - o0;
        cluster_wait();
        if (done) break;
        o0 = clock64(); o_waitA += o0 - o1;
        cluster_arrive_relaxed(); cluster_wait();                             // (B) partial sums are in
        o1 = clock64(); o_waitB += o1 - o0;
        // ---- fixed-order sum over the cluster's CTAs, one lane per value ------------------------
        float sacc = 0.0f;
        if (lane < 16) { for (int r = 0; r < C; r++) sacc += partials[r][lane]; tot16[lane] = sacc; }   // tot16[8..15]: lbs of the deferred pushes
        // ---- bookkeeping of the 8 children (jly_goicp.cpp:317-336), one lane per child -----------
        // The reference walks j = 0..7 with a running optErrorT; lane j needs that running value right after child j's
        // own update = min(optErrorT, ub_0..ub_j): an 8-wide prefix minimum.  Everything else is a ballot.
        constexpr unsigned kFull = 0xffffffffu;
        const int cj = lane & 7;
        const float ub = __shfl_sync(kFull, sacc, cj), lb = __shfl_sync(kFull, sacc, 8 + cj);
        float run = ub;
#pragma unroll
        for (int d = 1; d < 8; d <<= 1) { const float o = __shfl_up_sync(kFull, run, d, 8); if (cj >= d && o < run) run = o; }
        if (opt_t < run) run = opt_t;                                                // optErrorT after child cj was looked at
        evals += 8;
        if (first_phase) {
            // children the reference looks at while optErrorT is still the caller's: all 8, or those up to the first one whose ub
            // improves it (that one's ub and the queued lbs before it bound the smaller incumbents this call is also valid for)
            const unsigned imp = __ballot_sync(kFull, lane < 8 && ub < opt_t);
            const int jfirst = imp ? __ffs(imp) - 1 : 8;
            unsigned key = lane < jfirst && lane < 8 && lb < opt_t ? __float_as_uint(lb) : 0u;       // lbs are >= +0: bit patterns order like values
            if (imp && lane == jfirst) key = __float_as_uint(ub);
            reuse_gt = fmaxf(reuse_gt, __uint_as_float(__reduce_max_sync(kFull, key)));
            if (imp) first_phase = false;
        }
        {
            const float fin = __shfl_sync(kFull, run, 7);
            if (fin < opt_t) {                                                       // taken by the first child that reaches the final value
                const int jb = __ffs(__ballot_sync(kFull, lane < 8 && ub == fin)) - 1;
                opt_t = fin;
                best[0] = __fadd_rn(px, (jb & 1) ? cw : 0.0f); best[1] = __fadd_rn(py, (jb & 2) ? cw : 0.0f);
                best[2] = __fadd_rn(pz, (jb & 4) ? cw : 0.0f); best[3] = cw;
            }
        }
        if (ub_pass) {
            const float lim = run * (1.0f + cand_eps);
            const unsigned cm = __ballot_sync(kFull, lane < 8 && ub <= lim);
            if (n_cand + __popc(cm) <= kMaxCand) {
                if ((cm >> lane) & 1u) {
                    const int q = n_cand + __popc(cm & ((1u << lane) - 1u));
                    cand_node[q] = make_float4(__fadd_rn(px, (lane & 1) ? cw : 0.0f), __fadd_rn(py, (lane & 2) ? cw : 0.0f), __fadd_rn(pz, (lane & 4) ? cw : 0.0f), cw);
                    cand_ub[q] = ub;
                }
                n_cand += __popc(cm);
                __syncwarp();
            } else {
                // the list is (nearly) full: the sequential form, which purges entries the running optimum has left behind
                for (int j = 0; j < 8; j++) {
                    const float uj = __shfl_sync(kFull, ub, j), oj = __shfl_sync(kFull, run, j);
                    if (!(uj <= oj * (1.0f + cand_eps))) continue;
                    if (n_cand == kMaxCand) {
                        int k = 0;
                        for (int q = 0; q < n_cand; q++)
                            if (cand_ub[q] <= oj * (1.0f + cand_eps)) { if (lane == 0) { cand_node[k] = cand_node[q]; cand_ub[k] = cand_ub[q]; } k++; }
                        n_cand = k;
                        __syncwarp();
                    }
                    if (n_cand < kMaxCand) {
                        if (lane == 0) {
                            cand_node[n_cand] = make_float4(__fadd_rn(px, (j & 1) ? cw : 0.0f), __fadd_rn(py, (j & 2) ? cw : 0.0f), __fadd_rn(pz, (j & 4) ? cw : 0.0f), cw);
                            cand_ub[n_cand] = uj;
                        }
                        n_cand++;
                        __syncwarp();
                    } else flags |= 1u;
                }
            }
        }
        // children that go into the queue (`if (lb >= optErrorT) continue;`, and the span cut-off of the fgoicp-style search)
        pmask = __ballot_sync(kFull, lane < 8 && !(lb >= run) && (int)plevel + 1 <= c.trans_cutoff_level);
        pend_level = plevel + 1; pend_base = ppath; pend_shift = 3 * (int)plevel;
        __syncwarp();                                                                // tot16 is read by the deferred pushes
        // ---- which node will the queue pop next? ------------------------------------------------
        if (status) { done = true; publish(true); }
        else if (heap.n == 0 && pmask == 0) { done = true; publish(true); }          // queue empty (:243-244)
        else {
            // best candidate among the current top (lane 8) and the pending children (lanes 0..7): smallest lb, then smallest
            // level; two candidates with the same (lb, level) are left to the heap's own arrangement
            HeapEntry top; top.lb = 0.0f; top.level = 0; top.path_lo = top.path_hi = 0;
            if (heap.n > 0) top = heap.get(0);
            const bool valid = lane < 8 ? ((pmask >> lane) & 1u) != 0 : (lane == 8 && heap.n > 0);
            const float klb = lane == 8 ? top.lb : lb;
            const uint32_t klv = lane == 8 ? top.level : pend_level;
            bool in = valid;
            {   // lower bounds are sums of squares (>= +0): their bit patterns order like the values
                const unsigned kb = in ? __float_as_uint(klb) : 0xffffffffu;
                const unsigned kb_min = __reduce_min_sync(kFull, kb);              // (not inside an `&&`: every lane takes part)
                in = in && kb == kb_min;
                const unsigned kl = in ? klv : 0xffffffffu;
                const unsigned kl_min = __reduce_min_sync(kFull, kl);
                in = in && kl == kl_min;
            }
            const unsigned win = __ballot_sync(kFull, in);
            if (__popc(win) > 1) {
                flush_pending();                                               // the heap's own arrangement decides
                if (status) { done = true; publish(true); }
                else { const HeapEntry e = wheap_pop(heap, lane, pop_path); adopt(e); publish(done); }
            } else {
                const int wl = __ffs(win) - 1;
                if (wl == 8) { expect = top; need_pop = true; adopt(top); }
                else {
                    // a child of the cube just expanded: its corner is one more step of the reference's additions (:267-269)
                    expect = pend_entry(wl); need_pop = true;
                    pops++;                                                               // tNodeCount++ (:248)
                    if (__fsub_rn(opt_t, expect.lb) < c.sse_thresh) done = true;         // :257
                    else if (expect.level >= (uint32_t)kMaxTransLevel) { done = true; status = 4; }
                    else {
                        if (first_phase) reuse_poplb = fmaxf(reuse_poplb, expect.lb);
                        px = __fadd_rn(px, (wl & 1) ? cw : 0.0f); py = __fadd_rn(py, (wl & 2) ? cw : 0.0f); pz = __fadd_rn(pz, (wl & 4) ? cw : 0.0f);
                        cw = cw / 2; plevel = expect.level; ppath = ((unsigned long long)expect.path_hi << 32) | expect.path_lo;
                    }
                }
                publish(done);
            }
        }
        o_book += clock64() - o1;
    }
    if (c.dbg && lane == 0) {
        unsigned long long* d = c.dbg + (size_t)task_id * 12;
        d[0] = o_maint; d[1] = o_waitA; d[2] = o_waitB; d[3] = o_book; d[8] = o_arrive; d[9] = o_push; d[10] = o_npush; d[11] = o_pop;
    }

    // ---- results (see inner_bnb_kernel) ---------------------------------------------------------
    if (ub_pass) {
        __syncwarp();
        {
            int k = 0;
            for (int q = 0; q < n_cand; q++)
                if (cand_ub[q] <= opt_t * (1.0f + cand_eps)) { if (lane == 0) { cand_node[k] = cand_node[q]; cand_ub[k] = cand_ub[q]; } k++; }
            n_cand = k;
            __syncwarp();
        }
        CandList& cl = cands[task_id];
        for (int q = lane; q < n_cand; q += 32) { cl.node[q] = cand_node[q]; cl.ub[q] = cand_ub[q]; }
        if (lane == 0) { cl.n = n_cand; cl.flags = flags; cl.final_fast = opt_t; cl.eps = cand_eps; }
    }
    if (lane == 0) {
        InnerResult r;
        r.value = opt_t; r.node[0] = best[0]; r.node[1] = best[1]; r.node[2] = best[2]; r.node[3] = best[3];
        r.pops = pops; r.evals = evals; r.status = status == 5 ? 3 : status; r.max_heap = max_heap;
        r.pad[0] = flags | (status == 5 ? 0x100u : 0u); r.pad[1] = ub_pass ? (uint32_t)n_cand : 0u;
        r.kcycles = (uint32_t)((clock64() - t_begin) >> 10);
        r.reuse_gt = (status || (flags & 1u)) ? 3.402823466e+38f : reuse_gt; r.reuse_poplb = reuse_poplb;
        results[task_id] = r;
    }
}

// ------------------------------------------------------------------------------------------
// Strict resolution of one upper-bound pass (on demand, when the pass may improve the optimum).
// One CTA per contender: gathers in parallel into shared memory, then one thread applies the
// reference's intro_select + sequential float sum.  strict_pick_kernel replays the reference's
// `if (ub < optErrorT)` over the contenders in evaluation order.
// ------------------------------------------------------------------------------------------
constexpr int kStrictStaticSmem = 2 * kSsWin * (int)sizeof(float) + 1024;       // windows of the global-memory select + slack
__global__ void __launch_bounds__(256)
strict_eval_kernel(BnbConst c, const InnerTask* __restrict__ task_p, const CandList* __restrict__ cl, float* __restrict__ strict_ub,
                   float* __restrict__ gscratch, int use_smem /* 2: residuals + the select's position lists in shared memory, 1: residuals only, 0: neither */)
{
    extern __shared__ float m_sm[];
    __shared__ unsigned long long sel_sh[40];
    const int q = blockIdx.x;
    if (q >= cl->n) return;
    if (!(cl->ub[q] <= cl->final_fast * (1.0f + cl->eps))) { if (threadIdx.x == 0) strict_ub[q] = 3.402823466e+38f; return; }
    const InnerTask& task = *task_p;
    float* m = use_smem ? m_sm : gscratch + (size_t)q * 3 * c.nd;
    int* sel_idx = use_smem == 2 ? reinterpret_cast<int*>(m_sm + c.nd) : reinterpret_cast<int*>(gscratch + (size_t)q * 3 * c.nd + c.nd);
    const float4 nd4 = cl->node[q];
    const float half = nd4.w / 2;
    const float tx = __fadd_rn(nd4.x, half), ty = __fadd_rn(nd4.y, half), tz = __fadd_rn(nd4.z, half);   // :270-272
    for (int i = threadIdx.x; i < c.nd; i += blockDim.x) {
        const float4 p = __ldg(c.data + i);
        const float rx = dot3_ref(task.R[0], task.R[1], task.R[2], p.x, p.y, p.z);
        const float ry = dot3_ref(task.R[3], task.R[4], task.R[5], p.x, p.y, p.z);
        const float rz = dot3_ref(task.R[6], task.R[7], task.R[8], p.x, p.y, p.z);
        const float d = dt_distance(c.dt, __fadd_rn(rx, tx), __fadd_rn(ry, ty), __fadd_rn(rz, tz));
        m[i] = d < 0.0f ? 0.0f : d;
    }
    __syncthreads();
    __shared__ float win[2 * kSsWin];                     // staging of the sequential sum when the residuals live in global memory
    if (c.do_trim) ss_intro_select_cta(m, 0, c.nd - 1, c.inlier_num - 1, sel_idx, sel_sh);
    if (threadIdx.x == 0) {
        float ub, lb;
        ss_sum_selected(m, c.do_trim ? c.inlier_num : c.nd, 0.0f, false, ub, lb, use_smem ? nullptr : win);
        strict_ub[q] = ub;
    }
}
__global__ void strict_pick_kernel(const InnerTask* __restrict__ task_p, const CandList* __restrict__ cl, const float* __restrict__ strict_ub, float* __restrict__ out5)
{
    float so = task_p->opt_error;                    // optErrorT starts at optError (:238)
    float best[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    for (int q = 0; q < cl->n; q++)
        if (strict_ub[q] < so) { so = strict_ub[q]; best[0] = cl->node[q].x; best[1] = cl->node[q].y; best[2] = cl->node[q].z; best[3] = cl->node[q].w; }   // :319-324
    out5[0] = so; out5[1] = best[0]; out5[2] = best[1]; out5[3] = best[2]; out5[4] = best[3];
}

// ------------------------------------------------------------------------------------------
// (Trimmed) sum of squared DT distances of the data under a pose: the initial error
// (jly_goicp.cpp:357-371) and the DT re-scoring of GoICP::ICP (:100-131).  These few values
// become optError itself, so they are formed in the reference's order: gathers in parallel, then
// intro_select + sequential float sum by one thread (strict_sum.cuh).  One CTA per pose;
// use_pose[k]==0 scores the raw data.  scratch: nposes * nd floats.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBnbThreads)
dt_score_kernel(BnbConst c, const float* __restrict__ Rt12, const int* __restrict__ use_pose, float* __restrict__ scratch, float* __restrict__ out, int use_smem)
{
    extern __shared__ float m_sm[];
    __shared__ unsigned long long sel_sh[40];
    const float* Rt = Rt12 + 12 * blockIdx.x;
    const bool pose = use_pose[blockIdx.x] != 0;
    float* m = use_smem ? m_sm : scratch + (size_t)blockIdx.x * 3 * c.nd;
    int* sel_idx = use_smem == 2 ? reinterpret_cast<int*>(m_sm + c.nd) : reinterpret_cast<int*>(scratch + (size_t)blockIdx.x * 3 * c.nd + c.nd);
    for (int i = threadIdx.x; i < c.nd; i += kBnbThreads) {
        float4 p = __ldg(c.data + i);
        float x = p.x, y = p.y, z = p.z;
        if (pose) {
            x = __fadd_rn(dot3_ref(Rt[0], Rt[1], Rt[2], p.x, p.y, p.z), Rt[9]);
            y = __fadd_rn(dot3_ref(Rt[3], Rt[4], Rt[5], p.x, p.y, p.z), Rt[10]);
            z = __fadd_rn(dot3_ref(Rt[6], Rt[7], Rt[8], p.x, p.y, p.z), Rt[11]);
        }
        m[i] = dt_distance(c.dt, x, y, z);
    }
    __syncthreads();
    __shared__ float win[2 * kSsWin];                     // staging of the sequential sum when the residuals live in global memory
    const long long t0 = clock64();
    if (c.do_trim) ss_intro_select_cta(m, 0, c.nd - 1, c.inlier_num - 1, sel_idx, sel_sh);
    if (threadIdx.x == 0) {
        float ub, lb;
        const long long t1 = clock64();
        ss_sum_selected(m, c.do_trim ? c.inlier_num : c.nd, 0.0f, false, ub, lb, use_smem ? nullptr : win);
#ifdef GOICP_SCORE_TRACE
        printf("[dt_score] nd %d: select %lld cycles, sum %lld cycles\n", c.nd, t1 - t0, clock64() - t1);
#else
        (void)t0; (void)t1;
#endif
        out[blockIdx.x] = ub;
    }
}

// ------------------------------------------------------------------------------------------
// The same score with the sums in a fixed parallel order (GOICP_NUM_FAST_SUMS): every thread adds its strided share of
// the squared residuals, then a warp-shuffle tree and a fixed-order sum over the warps.  With trimming the inlier_num
// smallest residuals are found by a 4 x 8-bit radix select over their bit patterns (non-negative floats order like
// their bits); everything strictly below the threshold is summed and the ties at the threshold that still count are
// added as a product.  One CTA per pose; scratch (nposes * nd floats) is only touched when trimming.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBnbThreads)
dt_score_fast_kernel(BnbConst c, const float* __restrict__ Rt12, const int* __restrict__ use_pose, float* __restrict__ scratch, float* __restrict__ out)
{
    __shared__ float red[kBnbWarps];
    __shared__ int hist[256];
    __shared__ unsigned s_prefix; __shared__ int s_k;
    const float* Rt = Rt12 + 12 * blockIdx.x;
    const bool pose = use_pose[blockIdx.x] != 0;
    const bool trim = c.inlier_num < c.nd;
    float* m = scratch + (size_t)blockIdx.x * c.nd;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float acc = 0.0f;
    for (int i = threadIdx.x; i < c.nd; i += kBnbThreads) {
        float4 p = __ldg(c.data + i);
        float x = p.x, y = p.y, z = p.z;
        if (pose) {
            x = __fadd_rn(dot3_ref(Rt[0], Rt[1], Rt[2], p.x, p.y, p.z), Rt[9]);
            y = __fadd_rn(dot3_ref(Rt[3], Rt[4], Rt[5], p.x, p.y, p.z), Rt[10]);
            z = __fadd_rn(dot3_ref(Rt[6], Rt[7], Rt[8], p.x, p.y, p.z), Rt[11]);
        }
        const float d = dt_distance(c.dt, x, y, z);
        if (trim) m[i] = d; else acc = __fadd_rn(acc, __fmul_rn(d, d));
    }
    float tie_part = 0.0f;
    if (trim) {
        if (threadIdx.x == 0) { s_prefix = 0u; s_k = c.inlier_num; }
        for (int pass = 0; pass < 4; pass++) {
            const int shift = 24 - 8 * pass;
            for (int b = threadIdx.x; b < 256; b += kBnbThreads) hist[b] = 0;
            __syncthreads();
            const unsigned himask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
            const unsigned prefix = s_prefix;
            for (int i = threadIdx.x; i < c.nd; i += kBnbThreads) {
                const unsigned key = __float_as_uint(m[i]);
                if ((key & himask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1);
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                int k = s_k, b = 0, before = 0;
                for (; b < 256; b++) { if (before + hist[b] >= k) break; before += hist[b]; }
                s_prefix = prefix | ((unsigned)b << shift); s_k = k - before;
            }
            __syncthreads();
        }
        const unsigned T = s_prefix;
        for (int i = threadIdx.x; i < c.nd; i += kBnbThreads) {
            const float d = m[i];
            if (__float_as_uint(d) < T) acc = __fadd_rn(acc, __fmul_rn(d, d));
        }
        const float tv = __uint_as_float(T);
        tie_part = (float)s_k * (tv * tv);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, o));
    if (lane == 0) red[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.0f;
        for (int w = 0; w < kBnbWarps; w++) s = __fadd_rn(s, red[w]);
        out[blockIdx.x] = __fadd_rn(s, tie_part);
    }
}

// ------------------------------------------------------------------------------------------
// Roofline denominator of the DT gathers, measured rather than assumed: uniformly random 4-byte loads over a buffer of a
// given size and NOTHING else -- no voxel-index arithmetic, no reduction.  Eight independent addresses per step from one
// 32-bit state (multiplicative hashes), so a thread always has eight loads in flight.  goicp_measure_gather().
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512)
gather_peak_kernel(const float* __restrict__ buf, unsigned n, int iters, float* __restrict__ sink)
{
    unsigned s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 0x9e3779b9u;
    float acc = 0.0f;
    for (int it = 0; it < iters; it++) {
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const unsigned r = (s ^ (0x85ebca6bu * (unsigned)(k + 1))) * (0xc2b2ae35u + 2u * (unsigned)k);
            v[k] = __ldg(buf + (unsigned)(((unsigned long long)(r ^ (r >> 15)) * n) >> 32));
        }
#pragma unroll
        for (int k = 0; k < 8; k++) acc += v[k];
        s = s * 1664525u + 1013904223u;
    }
    if (acc == 1.2345678e33f) *sink = acc;               // never true: keeps the loads alive
}
cudaError_t launch_gather_peak(const float* d_buf, unsigned n, int iters, int blocks, float* d_sink, cudaStream_t s)
{
    gather_peak_kernel<<<blocks, 512, 0, s>>>(d_buf, n, iters, d_sink);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// host-callable launchers
// ------------------------------------------------------------------------------------------
cudaError_t launch_dt_lookup(const DtView& dt, const float* d_q, int n, float* d_out, int32_t* d_idx, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    dt_lookup_kernel<<<(n + 255) / 256, 256, 0, s>>>(dt, d_q, n, d_out, d_idx);
    return cudaGetLastError();
}
cudaError_t launch_pair_bounds(const BnbConst& c, const PairTask* d_tasks, int n, float2* d_out, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    pair_bounds_kernel<<<n, kBnbThreads, 0, s>>>(c, d_tasks, d_out);
    return cudaGetLastError();
}
cudaError_t launch_expand_bounds(const BnbConst& c, const PairTask* d_tasks, int n, float* d_out16, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    expand_bounds_kernel<<<n, kBnbThreads, 0, s>>>(c, d_tasks, d_out16);
    return cudaGetLastError();
}
// Opts the kernels into the full shared-memory carve-out; returns the dynamic bytes one CTA may use.
template <bool P, bool T>
static cudaError_t configure_one(int smem_optin, int* stat_out)
{
    cudaFuncAttributes a;
    cudaError_t e = cudaFuncGetAttributes(&a, inner_bnb_kernel<P, T>);
    if (e != cudaSuccess) return e;
    *stat_out = (int)a.sharedSizeBytes;
    e = cudaFuncSetAttribute(inner_bnb_kernel<P, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin - (int)a.sharedSizeBytes);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(inner_bnb_kernel<P, T>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
}
// Opts the kernels into the full shared-memory carve-out; returns the dynamic bytes one CTA may use.
cudaError_t inner_bnb_configure(int smem_optin, int* max_dyn_out)
{
    int st[4] = {0, 0, 0, 0};
    cudaError_t e;
    if ((e = configure_one<true, false>(smem_optin, &st[0])) != cudaSuccess) return e;
    if ((e = configure_one<false, false>(smem_optin, &st[1])) != cudaSuccess) return e;
    if ((e = configure_one<true, true>(smem_optin, &st[2])) != cudaSuccess) return e;
    if ((e = configure_one<false, true>(smem_optin, &st[3])) != cudaSuccess) return e;
    int stat = 0;
    for (int i = 0; i < 4; i++) stat = st[i] > stat ? st[i] : stat;
    {
        cudaFuncAttributes a;
        auto conf = [&](auto kern) -> cudaError_t {
            cudaError_t e2 = cudaFuncGetAttributes(&a, kern);
            if (e2 != cudaSuccess) return e2;
            stat = (int)a.sharedSizeBytes > stat ? (int)a.sharedSizeBytes : stat;
            if ((e2 = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin - (int)a.sharedSizeBytes)) != cudaSuccess) return e2;
            return cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        };
        if ((e = conf(inner_bnb_pipelined_kernel<true, false>)) != cudaSuccess) return e;
        if ((e = conf(inner_bnb_pipelined_kernel<true, true>)) != cudaSuccess) return e;
        if ((e = conf(inner_bnb_pipelined_kernel<false, false>)) != cudaSuccess) return e;
        if ((e = conf(inner_bnb_pipelined_kernel<false, true>)) != cudaSuccess) return e;
        if ((e = conf(inner_bnb_pipelined_kernel<true, false, 192, 5>)) != cudaSuccess) return e;
        if ((e = conf(inner_bnb_pipelined_kernel<false, false, 192, 5>)) != cudaSuccess) return e;
        if ((e = conf(inner_bnb_pipelined_kernel<true, false, 128, 8>)) != cudaSuccess) return e;
        if ((e = conf(inner_bnb_pipelined_kernel<false, false, 128, 8>)) != cudaSuccess) return e;
    }
    // both carry the select's two global-memory windows as static shared memory (strict_sum.cuh)
    e = cudaFuncSetAttribute(strict_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin - 1024 - kStrictStaticSmem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(dt_score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin - 1024 - kStrictStaticSmem);
    *max_dyn_out = smem_optin - stat;
    return e;
}
cudaError_t launch_inner_bnb(const BnbConst& c, const InnerTask* d_tasks, InnerResult* d_results, int n, int cluster,
                             bool pts_in_smem, int heap_cap_sm, HeapEntry* d_spill, int spill_cap, CandList* d_cands, int variant /* 0: 512 threads x 2 CTAs per SM, 1: low latency (512 x 1, 128 registers), 2: 192 x 5, 3: 128 x 8 */, unsigned* d_trim_keys, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    const bool low_latency = variant == 1;
    const bool trim = c.inlier_num < c.nd;
    const int per = (c.nd + cluster - 1) / cluster;
    size_t smem = (size_t)heap_cap_sm * sizeof(HeapEntry) + (pts_in_smem ? (size_t)per * sizeof(float4) : 0)
                + (trim && !d_trim_keys ? (size_t)per * 8 * sizeof(unsigned) : 0);
    // GOICP_BNB_MIN_SMEM_KB (experiment): pad the request so that fewer CTAs share an SM (and its L1TEX pipe)
    static const size_t min_smem = getenv("GOICP_BNB_MIN_SMEM_KB") ? (size_t)atoi(getenv("GOICP_BNB_MIN_SMEM_KB")) << 10 : 0;
    if (smem < min_smem) smem = min_smem;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)n * cluster); cfg.blockDim = dim3(kBnbThreads); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    BnbConst cc = c;
    static const bool legacy = getenv("GOICP_NO_PIPELINE") != nullptr;     // A/B switch for profiling
    if (!trim && !legacy) {
        if (variant >= 2) {
            cfg.blockDim = dim3(variant == 2 ? 192 : 128);
            if (variant == 2) return pts_in_smem ? cudaLaunchKernelEx(&cfg, inner_bnb_pipelined_kernel<true, false, 192, 5>, cc, d_tasks, d_results, heap_cap_sm, d_spill, spill_cap, d_cands)
                                                 : cudaLaunchKernelEx(&cfg, inner_bnb_pipelined_kernel<false, false, 192, 5>, cc, d_tasks, d_results, heap_cap_sm, d_spill, spill_cap, d_cands);
            return pts_in_smem ? cudaLaunchKernelEx(&cfg, inner_bnb_pipelined_kernel<true, false, 128, 8>, cc, d_tasks, d_results, heap_cap_sm, d_spill, spill_cap, d_cands)
                               : cudaLaunchKernelEx(&cfg, inner_bnb_pipelined_kernel<false, false, 128, 8>, cc, d_tasks, d_results, heap_cap_sm, d_spill, spill_cap, d_cands);
        }
        if (pts_in_smem) return low_latency ? cudaLaunchKernelEx(&cfg, inner_bnb_pipelined_kernel<true, true>, cc, d_tasks, d_results, heap_cap_sm, d_spill, spill_cap, d_cands)
                                            : cudaLaunchKernelEx(&cfg, inner_bnb_pipelined_kernel<true, false>, cc, d_tasks, d_results, heap_cap_sm, d_spill, spill_cap, d_cands);
        return low_latency ? cudaLaunchKernelEx(&cfg, inner_bnb_pipelined_kernel<false, true>, cc, d_tasks, d_results, heap_cap_sm, d_spill, spill_cap, d_cands)
                           : cudaLaunchKernelEx(&cfg, inner_bnb_pipelined_kernel<false, false>, cc, d_tasks, d_results, heap_cap_sm, d_spill, spill_cap, d_cands);
    }
    if (pts_in_smem && !trim) return cudaLaunchKernelEx(&cfg, inner_bnb_kernel<true, false>, cc, d_tasks, d_results, heap_cap_sm, d_spill, spill_cap, d_cands, d_trim_keys);
    if (!pts_in_smem && !trim) return cudaLaunchKernelEx(&cfg, inner_bnb_kernel<false, false>, cc, d_tasks, d_results, heap_cap_sm, d_spill, spill_cap, d_cands, d_trim_keys);
    if (pts_in_smem) return cudaLaunchKernelEx(&cfg, inner_bnb_kernel<true, true>, cc, d_tasks, d_results, heap_cap_sm, d_spill, spill_cap, d_cands, d_trim_keys);
    return cudaLaunchKernelEx(&cfg, inner_bnb_kernel<false, true>, cc, d_tasks, d_results, heap_cap_sm, d_spill, spill_cap, d_cands, d_trim_keys);
}
// intro_select on a caller's array (parity hook, goicp_intro_select): the CTA-wide select of the strict kernels on its own.
// in_smem: the array and the position lists are staged in dynamic shared memory (3 * n floats), else they stay in global memory.
__global__ void __launch_bounds__(1024)
select_test_kernel(float* __restrict__ a, int n, int k, int* __restrict__ idx, int in_smem)
{
    extern __shared__ float m_sm[];
    __shared__ unsigned long long sel_sh[40];
    float* m = in_smem ? m_sm : a;
    if (in_smem) { for (int i = threadIdx.x; i < n; i += blockDim.x) m[i] = a[i]; __syncthreads(); }
    ss_intro_select_cta(m, 0, n - 1, k, in_smem ? reinterpret_cast<int*>(m_sm + n) : idx, sel_sh);
    if (in_smem) for (int i = threadIdx.x; i < n; i += blockDim.x) a[i] = m[i];
}
cudaError_t launch_select_test(float* d_a, int n, int k, int* d_idx, int threads, int smem_limit, bool allow_smem, cudaStream_t s)
{
    const size_t need = 3 * (size_t)n * sizeof(float);
    const int in_smem = allow_smem && need + 1024 <= (size_t)smem_limit ? 1 : 0;
    if (in_smem) { cudaError_t e = cudaFuncSetAttribute(select_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need); if (e != cudaSuccess) return e; }
    select_test_kernel<<<1, threads, in_smem ? need : 0, s>>>(d_a, n, k, d_idx, in_smem);
    return cudaGetLastError();
}
// where the residuals (nd floats) and the position lists of the CTA-wide select (2 * nd ints) of one strict evaluation live:
// 2 = all in shared memory, 1 = residuals in shared memory, 0 = all in the caller's scratch (3 * nd floats per evaluation)
int strict_smem_mode(int nd, int smem_limit)
{
    const size_t need = (size_t)nd * sizeof(float);
    return 3 * need + kStrictStaticSmem <= (size_t)smem_limit ? 2 : (need + kStrictStaticSmem <= (size_t)smem_limit ? 1 : 0);
}
// out5 = {strict optErrorT, node x, y, z, w}; d_strict: kMaxCand floats; d_scratch: kMaxCand*nd floats (only if nd does not fit in smem)
cudaError_t launch_strict_resolve(const BnbConst& c, const InnerTask* d_task, const CandList* d_list, float* d_strict, float* d_scratch,
                                  float* d_out5, int smem_limit, cudaStream_t s)
{
    const size_t need = (size_t)c.nd * sizeof(float);
    const int use_smem = strict_smem_mode(c.nd, smem_limit);
    strict_eval_kernel<<<kMaxCand, 256, use_smem == 2 ? 3 * need : (use_smem == 1 ? need : 0), s>>>(c, d_task, d_list, d_strict, d_scratch, use_smem);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    strict_pick_kernel<<<1, 1, 0, s>>>(d_task, d_list, d_strict, d_out5);
    return cudaGetLastError();
}
cudaError_t launch_dt_score(const BnbConst& c, const float* d_Rt12, const int* d_use_pose, int nposes, float* d_scratch, float* d_out, int smem_limit, bool fast_sums, cudaStream_t s)
{
    if (nposes <= 0) return cudaSuccess;
    if (fast_sums) { dt_score_fast_kernel<<<nposes, kBnbThreads, 0, s>>>(c, d_Rt12, d_use_pose, d_scratch, d_out); return cudaGetLastError(); }
    const size_t need = (size_t)c.nd * sizeof(float);
    const int use_smem = strict_smem_mode(c.nd, smem_limit);
    dt_score_kernel<<<nposes, kBnbThreads, use_smem == 2 ? 3 * need : (use_smem == 1 ? need : 0), s>>>(c, d_Rt12, d_use_pose, d_scratch, d_out, use_smem);
    return cudaGetLastError();
}

} // namespace goicp
