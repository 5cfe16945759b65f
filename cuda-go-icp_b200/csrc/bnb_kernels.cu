// bnb_kernels.cu -- the Go-ICP hot path on sm_100a: distance-transform gathers, bound
// evaluation and the persistent translation branch-and-bound.
//
// Replaces, for the GPU: DT3D::Distance (jly_3ddt.cpp:981-1026), the bound-evaluation loop
// body (jly_goicp.cpp:265-336) and GoICP::InnerBnB (jly_goicp.cpp:227-340).  Roofline: random
// 4-byte gathers into the S^3 float grid (L2-resident at S=300, HBM at S=512); no dense
// contraction anywhere, so no tensor cores (DESIGN.md "Kernels").
#include "goicp_types.h"
#include "goicp_kernels.h"
#include "strict_sum.cuh"

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
    float tx[2], ty[2], tz[2];
#pragma unroll
    for (int b = 0; b < 2; b++) {                    // nodeTrans.x = parent.x + bit*w ; transX = nodeTrans.x + w/2
        tx[b] = __fadd_rn(__fadd_rn(t.tc[0], b ? w : 0.0f), half);
        ty[b] = __fadd_rn(__fadd_rn(t.tc[1], b ? w : 0.0f), half);
        tz[b] = __fadd_rn(__fadd_rn(t.tc[2], b ? w : 0.0f), half);
    }
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
        accumulate_point8(c.dt, rx, ry, rz, __fmul_rn(cg, p.w), tx, ty, tz, gt, acc);
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

// ------------------------------------------------------------------------------------------
// Persistent translation branch-and-bound: ONE CTA runs one whole GoICP::InnerBnB call
// (jly_goicp.cpp:227-340) -- pop, expand into 8 children, 8*Nd DT gathers, reduce, prune,
// push -- without leaving the SM.  Thread 0 owns the priority queue; all threads gather.
// ------------------------------------------------------------------------------------------
struct InnerCtrl {
    float tx[2], ty[2], tz[2];
    float gt;
    int done;
    int n_cand;          // upper-bound pass: contenders for the arg-min (see strict_sum.cuh)
    float final_fast;
};
constexpr int kMaxCand = 128;

template <bool PTS_SMEM>
__global__ void __launch_bounds__(kBnbThreads, 2)
inner_bnb_kernel(BnbConst c, const InnerTask* __restrict__ tasks, InnerResult* __restrict__ results,
                 int heap_cap_sm, HeapEntry* __restrict__ spill, int spill_cap)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ float red[kBnbWarps][16];
    __shared__ float tot[16];
    __shared__ InnerCtrl ctrl;
    __shared__ float4 cand_node[kMaxCand];
    __shared__ float cand_ub[kMaxCand];

    HeapEntry* hsm = reinterpret_cast<HeapEntry*>(smem_raw);
    float4* pts = reinterpret_cast<float4*>(smem_raw + (size_t)heap_cap_sm * sizeof(HeapEntry));

    const InnerTask& task = tasks[blockIdx.x];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float cg = task.level >= 0 ? c.cgamma[task.level] : 0.0f;
    const float R0 = task.R[0], R1 = task.R[1], R2 = task.R[2], R3 = task.R[3], R4 = task.R[4], R5 = task.R[5],
                R6 = task.R[6], R7 = task.R[7], R8 = task.R[8];

    if (PTS_SMEM) {
        // rotate once (jly_goicp.cpp:470-476) and keep (R p, gamma) on chip for the whole search
        for (int i = tid; i < c.nd; i += kBnbThreads) {
            float4 p = __ldg(c.data + i);
            pts[i] = make_float4(dot3_ref(R0, R1, R2, p.x, p.y, p.z), dot3_ref(R3, R4, R5, p.x, p.y, p.z),
                                 dot3_ref(R6, R7, R8, p.x, p.y, p.z), __fmul_rn(cg, p.w));
        }
    }

    // ---- thread-0 state -----------------------------------------------------------------
    Heap heap;
    heap.sm = hsm; heap.gl = spill + (size_t)blockIdx.x * spill_cap; heap.cap_sm = heap_cap_sm;
    heap.cap_total = heap_cap_sm + spill_cap; heap.n = 0;
    float opt_t = task.opt_error;
    float best[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    uint32_t pops = 0, evals = 0, max_heap = 0;
    int status = 0;
    const bool ub_pass = task.level < 0;
    // |sequential float sum - tree float sum| <= ~Nd*2^-24 relative; contenders within twice that
    const float cand_eps = fminf(1e-2f, 2.0f * (float)c.nd * 5.9604645e-8f) + 1e-6f;
    int n_cand = 0; uint32_t flags = 0;
    // parent being expanded
    float px = 0, py = 0, pz = 0, cw = 0;
    uint32_t plevel = 0, ppath_lo = 0, ppath_hi = 0;

    if (tid == 0) {
        HeapEntry root; root.lb = 0.0f; root.level = 0; root.path_lo = 0; root.path_hi = 0;   // initNodeTrans.lb = 0 (jly_goicp.cpp:63)
        heap_push(heap, root);
    }

    for (;;) {
        if (tid == 0) {
            int done = 0;
            if (status) done = 1;                                                 // heap capacity exceeded: give up loudly
            else if (heap.n == 0) done = 1;                                      // :243-244
            else {
                HeapEntry e = heap_pop(heap);
                pops++;                                                           // tNodeCount++ (:248)
                if (__fsub_rn(opt_t, e.lb) < c.sse_thresh) done = 1;              // :257
                else if (e.level >= (uint32_t)kMaxTransLevel) { done = 1; status = 4; }
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
                    px = x; py = y; pz = z; cw = w / 2;                           // nodeTrans.w = parent.w/2 (:262)
                    plevel = e.level; ppath_lo = e.path_lo; ppath_hi = e.path_hi;
                    const float half = cw / 2;
#pragma unroll
                    for (int b = 0; b < 2; b++) {
                        ctrl.tx[b] = __fadd_rn(__fadd_rn(px, b ? cw : 0.0f), half);
                        ctrl.ty[b] = __fadd_rn(__fadd_rn(py, b ? cw : 0.0f), half);
                        ctrl.tz[b] = __fadd_rn(__fadd_rn(pz, b ? cw : 0.0f), half);
                    }
                    ctrl.gt = max_trans_dis(cw);
                }
            }
            ctrl.done = done;
        }
        __syncthreads();
        if (ctrl.done) break;

        // ---- 8 x Nd distance-transform gathers -----------------------------------------
        float acc[16];
#pragma unroll
        for (int k = 0; k < 16; k++) acc[k] = 0.0f;
        {
            const float tx[2] = {ctrl.tx[0], ctrl.tx[1]}, ty[2] = {ctrl.ty[0], ctrl.ty[1]}, tz[2] = {ctrl.tz[0], ctrl.tz[1]};
            const float gt = ctrl.gt;
            if (PTS_SMEM) {
                for (int i = tid; i < c.nd; i += kBnbThreads) {
                    float4 p = pts[i];
                    accumulate_point8(c.dt, p.x, p.y, p.z, p.w, tx, ty, tz, gt, acc);
                }
            } else {
                for (int i = tid; i < c.nd; i += kBnbThreads) {
                    float4 p = __ldg(c.data + i);
                    accumulate_point8(c.dt, dot3_ref(R0, R1, R2, p.x, p.y, p.z), dot3_ref(R3, R4, R5, p.x, p.y, p.z),
                                      dot3_ref(R6, R7, R8, p.x, p.y, p.z), __fmul_rn(cg, p.w), tx, ty, tz, gt, acc);
                }
            }
        }
        block_reduce16(acc, red, tot, warp, lane);      // contains a __syncthreads()

        // ---- sequential bookkeeping of the 8 children (jly_goicp.cpp:317-336) -------------
        if (tid == 0) {
            evals += 8;
            for (int j = 0; j < 8; j++) {
                const float ub = tot[j], lb = tot[8 + j];
                if (ub < opt_t) {
                    opt_t = ub;
                    best[0] = __fadd_rn(px, (j & 1) ? cw : 0.0f);
                    best[1] = __fadd_rn(py, (j & 2) ? cw : 0.0f);
                    best[2] = __fadd_rn(pz, (j & 4) ? cw : 0.0f);
                    best[3] = cw;
                }
                if (ub_pass && ub <= opt_t * (1.0f + cand_eps)) {
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
                    } else flags |= 1u;                 // could not keep every contender: result is the tree-sum one
                }
                if (lb >= opt_t) continue;
                HeapEntry e; e.lb = lb; e.level = plevel + 1;
                unsigned long long path = (((unsigned long long)ppath_hi << 32) | ppath_lo) | ((unsigned long long)j << (3 * plevel));
                e.path_lo = (uint32_t)path; e.path_hi = (uint32_t)(path >> 32);
                if (!heap_push(heap, e)) { status = 3; break; }
            }
            if ((uint32_t)heap.n > max_heap) max_heap = heap.n;
        }
        // a failed push (status != 0) ends the task at the top of the next iteration
    }

    // ---- strict resolution of the arg-min (upper-bound pass only) ---------------------------
    // The search above used fixed-order tree sums.  The reference's optErrorT is the first strict
    // minimum of its own sequential sums over the same evaluated cubes; re-evaluate the contenders
    // in reference order (strict_sum.cuh) and replay that rule.
    float strict_value = opt_t;
    if (ub_pass) {
        if (tid == 0) {
            int k = 0;
            for (int q = 0; q < n_cand; q++)
                if (cand_ub[q] <= opt_t * (1.0f + cand_eps)) { cand_node[k] = cand_node[q]; cand_ub[k] = cand_ub[q]; k++; }
            ctrl.n_cand = (flags & 1u) ? 0 : k;
        }
        __syncthreads();
        const int nc = ctrl.n_cand;
        float* scratch = reinterpret_cast<float*>(spill + (size_t)blockIdx.x * spill_cap);
        const size_t scratch_floats = (size_t)spill_cap * (sizeof(HeapEntry) / sizeof(float));
        int per_chunk = (int)min((size_t)kBnbWarps, scratch_floats / (size_t)max(c.nd, 1));
        if (nc > 0 && per_chunk == 0) { if (tid == 0) flags |= 2u; }
        else if (nc > 0) {
            float so = task.opt_error;                       // thread 0: the reference's running optErrorT
            bool have = false;
            for (int base = 0; base < nc; base += per_chunk) {
                const int cnt = min(per_chunk, nc - base);
                for (int q = 0; q < cnt; q++) {
                    const float4 nd4 = cand_node[base + q];
                    const float half = nd4.w / 2;
                    const float tx = __fadd_rn(nd4.x, half), ty = __fadd_rn(nd4.y, half), tz = __fadd_rn(nd4.z, half);
                    float* m = scratch + (size_t)q * c.nd;
                    for (int i = tid; i < c.nd; i += kBnbThreads) {
                        float rx, ry, rz;
                        if (PTS_SMEM) { const float4 p = pts[i]; rx = p.x; ry = p.y; rz = p.z; }
                        else {
                            const float4 p = __ldg(c.data + i);
                            rx = dot3_ref(R0, R1, R2, p.x, p.y, p.z); ry = dot3_ref(R3, R4, R5, p.x, p.y, p.z); rz = dot3_ref(R6, R7, R8, p.x, p.y, p.z);
                        }
                        float d = dt_distance(c.dt, __fadd_rn(rx, tx), __fadd_rn(ry, ty), __fadd_rn(rz, tz));
                        m[i] = d < 0.0f ? 0.0f : d;
                    }
                }
                __syncthreads();
                if (lane == 0 && warp < cnt) {
                    float ub, lb;
                    ss_select_and_sum(scratch + (size_t)warp * c.nd, c.nd, c.inlier_num, c.do_trim != 0, 0.0f, false, ub, lb);
                    cand_ub[base + warp] = ub;                   // now the reference-order sum
                }
                __syncthreads();
                if (tid == 0) {
                    for (int q = 0; q < cnt; q++)
                        if (cand_ub[base + q] < so) {            // `if(ub < optErrorT)` in evaluation order (:319-324)
                            so = cand_ub[base + q]; have = true;
                            best[0] = cand_node[base + q].x; best[1] = cand_node[base + q].y; best[2] = cand_node[base + q].z; best[3] = cand_node[base + q].w;
                        }
                }
            }
            if (tid == 0) { strict_value = so; if (!have) { best[0] = best[1] = best[2] = best[3] = 0.0f; } }
        }
    }

    if (tid == 0) {
        InnerResult r;
        r.value = strict_value; r.node[0] = best[0]; r.node[1] = best[1]; r.node[2] = best[2]; r.node[3] = best[3];
        r.pops = pops; r.evals = evals; r.status = status; r.max_heap = max_heap; r.pad[0] = flags; r.pad[1] = __float_as_uint(opt_t);
        results[blockIdx.x] = r;
    }
}

// ------------------------------------------------------------------------------------------
// (Trimmed) sum of squared DT distances of the data under a pose: the initial error
// (jly_goicp.cpp:357-371) and the DT re-scoring of GoICP::ICP (:100-131).  These few values
// become optError itself, so they are formed in the reference's order: gathers in parallel, then
// intro_select + sequential float sum by one thread (strict_sum.cuh).  One CTA per pose;
// use_pose[k]==0 scores the raw data.  scratch: nposes * nd floats.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBnbThreads)
dt_score_kernel(BnbConst c, const float* __restrict__ Rt12, const int* __restrict__ use_pose, float* __restrict__ scratch, float* __restrict__ out)
{
    const float* Rt = Rt12 + 12 * blockIdx.x;
    const bool pose = use_pose[blockIdx.x] != 0;
    float* m = scratch + (size_t)blockIdx.x * c.nd;
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
    if (threadIdx.x == 0) {
        float ub, lb;
        ss_select_and_sum(m, c.nd, c.do_trim ? c.inlier_num : c.nd, c.do_trim != 0, 0.0f, false, ub, lb);
        out[blockIdx.x] = ub;
    }
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
cudaError_t inner_bnb_configure(int smem_optin, int* max_dyn_out)
{
    cudaFuncAttributes a0, a1;
    cudaError_t e = cudaFuncGetAttributes(&a0, inner_bnb_kernel<true>);
    if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(&a1, inner_bnb_kernel<false>);
    if (e != cudaSuccess) return e;
    const int stat = (int)(a0.sharedSizeBytes > a1.sharedSizeBytes ? a0.sharedSizeBytes : a1.sharedSizeBytes);
    const int dyn = smem_optin - stat;
    e = cudaFuncSetAttribute(inner_bnb_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(inner_bnb_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
    *max_dyn_out = dyn;
    return e;
}
cudaError_t launch_inner_bnb(const BnbConst& c, const InnerTask* d_tasks, InnerResult* d_results, int n,
                             bool pts_in_smem, int heap_cap_sm, HeapEntry* d_spill, int spill_cap, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    size_t smem = (size_t)heap_cap_sm * sizeof(HeapEntry) + (pts_in_smem ? (size_t)c.nd * sizeof(float4) : 0);
    if (pts_in_smem) inner_bnb_kernel<true><<<n, kBnbThreads, smem, s>>>(c, d_tasks, d_results, heap_cap_sm, d_spill, spill_cap);
    else             inner_bnb_kernel<false><<<n, kBnbThreads, smem, s>>>(c, d_tasks, d_results, heap_cap_sm, d_spill, spill_cap);
    return cudaGetLastError();
}
cudaError_t launch_dt_score(const BnbConst& c, const float* d_Rt12, const int* d_use_pose, int nposes, float* d_scratch, float* d_out, cudaStream_t s)
{
    if (nposes <= 0) return cudaSuccess;
    dt_score_kernel<<<nposes, kBnbThreads, 0, s>>>(c, d_Rt12, d_use_pose, d_scratch, d_out);
    return cudaGetLastError();
}

} // namespace goicp
