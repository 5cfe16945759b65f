// goicp_kernels.h -- host-callable launchers of the CUDA kernels (implemented in *.cu).
#pragma once
#include "goicp_types.h"

namespace goicp {

// ---- bnb_kernels.cu ---------------------------------------------------------------------
cudaError_t launch_dt_lookup(const DtView& dt, const float* d_q, int n, float* d_out, int32_t* d_idx, cudaStream_t s);
cudaError_t launch_pair_bounds(const BnbConst& c, const PairTask* d_tasks, int n, float2* d_out, cudaStream_t s);
cudaError_t launch_expand_bounds(const BnbConst& c, const PairTask* d_tasks, int n, float* d_out16, cudaStream_t s);
cudaError_t inner_bnb_configure(int smem_optin, int* max_dyn_out);
cudaError_t launch_inner_bnb(const BnbConst& c, const InnerTask* d_tasks, InnerResult* d_results, int n, int cluster,
                             bool pts_in_smem, int heap_cap_sm, HeapEntry* d_spill, int spill_cap, CandList* d_cands, int variant,
                             unsigned* d_trim_keys /* trimming: per-CTA slabs of 8 * ceil(nd/cluster) keys in global memory, or null = shared memory */, cudaStream_t s);
cudaError_t launch_select_test(float* d_a, int n, int k, int* d_idx, int threads, int smem_limit, bool allow_smem, cudaStream_t s);
int strict_smem_mode(int nd, int smem_limit);        // 2 / 1 / 0: see bnb_kernels.cu
cudaError_t launch_strict_resolve(const BnbConst& c, const InnerTask* d_task, const CandList* d_list, float* d_strict, float* d_scratch,
                                  float* d_out5, int smem_limit, cudaStream_t s);
cudaError_t launch_dt_score(const BnbConst& c, const float* d_Rt12, const int* d_use_pose, int nposes, float* d_scratch, float* d_out, int smem_limit, bool fast_sums, cudaStream_t s);

cudaError_t launch_gather_peak(const float* d_buf, unsigned n, int iters, int blocks, float* d_sink, cudaStream_t s);

// ---- icp_kernels.cu ---------------------------------------------------------------------
// Flattened copy of the reference-ordered kd-tree (built on the host by kdtree_host.cpp).
struct KdNode {             // 32 B
    int32_t child1, child2; // -1/-1 = leaf
    int32_t left, right;    // leaf: [left,right) into vind
    int32_t divfeat;
    float divlow, divhigh;
    int32_t pad;
};
struct KdView {
    const KdNode* __restrict__ nodes;
    const int32_t* __restrict__ vind;
    const float4* __restrict__ pts_leaf;   // model points permuted into leaf order: x,y,z,(original index as int bits)
    const float* __restrict__ model;       // original xyz triples
    const float4* __restrict__ boxes;      // 2 per node: tight box of the subtree's points {lo xyz, m}, {hi xyz, 0}; interior nodes: m = int bits of
                                           // the first leaf-order position of child2's points (kdtree_host.cpp); may be null
    int nm;
    float bb_lo[3], bb_hi[3];
    int n_top; int top[32];                // the nodes five levels below the root (n_top = 32), or n_top = 0: where the cooperative search's frontier starts
    // Uniform grid over the model's bounding box (models of the linear-scan range only, else grid_start = null): cell c holds
    // grid_pts[grid_start[c] .. grid_start[c+1]), cells ordered x fastest.  A query first looks at the 3x3x3 cells around its
    // own; if the nearest point found there is closer than the edge of that block, no point outside can beat or tie it.
    const unsigned short* __restrict__ grid_start;
    const float4* __restrict__ grid_pts;   // model points sorted by cell: x, y, z, (original index as int bits)
    int gdim[3]; int gcells;
    float glo[3]; float gh, ginv_h;
};
struct IcpState {           // lives in device memory; written by block 0 of the ICP kernel
    float R[9], t[3];
    float mu_m[3], mu_d[3];
    float err, err_new;
    int iter, converged;
    long long dbg[6];       // cycles spent by CTA 0 in: NN, wait, sort, pass 1, pass 2 + SVD, total
};
struct IcpWork {            // per-iteration device scratch of the ICP kernel
    float* q;               // 8*nd: correspondence rows in query order (model xyz, query xyz, d^2, pad)
    int32_t* nn;            // nd nearest model indices
    int32_t* pos;           // nd leaf-order positions of those points (seed of the next iteration's search)
    float* d2;              // nd squared distances
    unsigned long long* keys;   // nd sort keys (d2 bits << 32 | point index)
    int32_t* order;         // nd: order[rank] = point index, ranks by key
    float* stage;           // 8*nd: correspondences in sorted order
    unsigned long long* keys2;  // nd: second key buffer of the radix sort (large clouds)
    unsigned* hist;         // 256 * (warps of the grid): digit-major histogram / scanned bases of the radix sort
    unsigned* blocksum;     // 1024: per-CTA chunk totals of that scan
};
cudaError_t launch_svd3(const float* d_H9, int n, float* d_U9, float* d_W3, float* d_V9, cudaStream_t s);
cudaError_t launch_nn(const KdView& kd, const float* d_q, int n, int32_t* d_idx, float* d_d2, bool cooperative, cudaStream_t s);
// Whole ICP3D::Run as ONE cooperative kernel (grid-synchronous iterations, no host round trips).
// do_sort = the reference's do_trim (it sorts the correspondences only then, jly_icp3d.hpp:236-239); fast = GOICP_NUM_FAST_ICP
// (parallel moments + Jacobi Procrustes, d_partials: 2 * grid_blocks * 16 floats).
cudaError_t launch_icp(const KdView& kd, int n_nodes, const float4* d_data, int nd, IcpState* d_state, const IcpWork& work,
                       int max_iter, float err_diff, int num_inliers, int do_sort, int grid_blocks, int smem_optin, bool fast, float* d_partials, cudaStream_t s);
// Sharded fast ICP (one iteration per call): NN + moments of the queries [q_begin, q_end) -> 16 doubles in d_xch_out; after the
// ranks' blocks have been all-gathered, icp_fast_solve forms the new pose from the W x 16 moments (same result on every rank).
cudaError_t launch_icp_fast_shard(const KdView& kd, int n_nodes, const float4* d_data, int nd, IcpState* d_state, const IcpWork& work,
                                  int num_inliers, int grid_blocks, int smem_optin, float* d_partials, int q_begin, int q_end, double* d_xch_out, cudaStream_t s);
cudaError_t launch_icp_fast_solve(IcpState* d_state, const double* d_xch_all, int W, int nd, int num, float err_diff, int iter, int max_iter, cudaStream_t s);
int icp_threads();
int icp_max_blocks_supported();   // CTAs the in-kernel radix sort's scan supports
int icp_max_grid_blocks(int device, const KdView& kd, int n_nodes, int nd, int num, int smem_optin, bool fast);

} // namespace goicp
