// goicp_types.h -- plain structs shared between the host engine and the kernels.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>
#include "goicp_device.cuh"

namespace goicp {

constexpr int kMaxRotLevel = 20;        // MAXROTLEVEL, jly_goicp.h:78
constexpr int kMaxTransLevel = 21;      // 3 path bits per level in a 64-bit word
constexpr int kBnbThreads = 512;        // CTA size of the translation-BnB / bound kernels
constexpr int kBnbWarps = kBnbThreads / 32;

// Everything a bound evaluation needs that is constant over a Register() call.
struct BnbConst {
    DtView dt;
    const float4* __restrict__ data;    // data cloud: x, y, z, ||p|| (normData, jly_goicp.cpp:143-147)
    int nd;
    int inlier_num;                     // == nd unless trimming
    int do_trim;                        // GoICP::doTrim: the reference runs intro_select iff set (even when nothing is trimmed)
    float sse_thresh;                   // SSEThresh (jly_goicp.cpp:208)
    float tx, ty, tz, tw;               // initNodeTrans (jly_goicp.cpp:50-53)
    float cgamma[kMaxRotLevel];         // 2*sinf(maxAngle_l/2): maxRotDis[l][i] = cgamma[l]*||p_i|| (jly_goicp.cpp:150-160)
    int trans_cutoff_level;             // translation cubes deeper than this are evaluated but not queued (kMaxTransLevel = no cut-off; the
                                        // fgoicp-style search stops subdividing below a span, fgoicp.cpp:156-157)
    unsigned long long* dbg;            // optional (GOICP_ROUND_STATS): 8 cycle counters per task of the pipelined inner BnB, else nullptr
};

// One inner (translation) BnB to run: a rotation-cube child and a pass.
struct InnerTask {
    float R[9];          // rotation of the cube centre, row-major (computed on the host with glibc
                         // sincosf exactly like jly_goicp.cpp:449-467)
    int32_t level;       // rotation level for the lower-bound pass, -1 for the upper-bound pass
    float opt_error;     // optError the sequential reference would hold when making this call
    int32_t pad;
};

struct InnerResult {
    float value;         // optErrorT returned by InnerBnB
    float node[4];       // best translation cube (x,y,z,w) of the ub pass
    uint32_t pops;       // tNodeCount increments of this call
    uint32_t evals;      // bound evaluations of this call
    int32_t status;      // goicp_status
    uint32_t max_heap;
    uint32_t pad[2];     // [0] flags, [1] number of arg-min contenders (ub pass)
    uint32_t kcycles;    // SM cycles / 1024 this task took
    // For which smaller incumbents E' < opt_error the call would have run EXACTLY the same way (same pops, same pushes, same
    // arg-min; `value` the same unless it is opt_error itself, in which case it becomes E'): every decision that involves the
    // initial optErrorT directly is taken before the call's first own improvement, and it comes out the same iff
    //   E' > reuse_gt       (no queued child's lb lies in [E', opt_error); the first improving ub is below E' too), and
    //   E' - reuse_poplb >= SSEThresh in float (no popped node of that phase would have ended the call, jly_goicp.cpp:257).
    // The engine keeps speculative results across an improvement of the incumbent when both hold (engine.cu: revalidate).
    float reuse_gt, reuse_poplb;
};

// Translation-BnB heap entry (16 B): lb, level, and the octant path from the root cube.
// Priority order of std::priority_queue<TRANSNODE> (jly_goicp.h:60-72): lower lb first,
// equal lb -> larger w first == smaller level first.
struct __align__(16) HeapEntry {
    float lb;
    uint32_t level;
    uint32_t path_lo, path_hi;
};

// Contenders for the arg-min of one upper-bound pass (see strict_sum.cuh), in evaluation order.
constexpr int kMaxCand = 128;
struct CandList {
    int32_t n; uint32_t flags; float final_fast; float eps;
    float4 node[kMaxCand];   // translation cube x,y,z,w
    float ub[kMaxCand];      // its tree-sum upper bound
};

// A generic (rotation, translation cube) pair for goicp_eval_bounds.
struct PairTask {
    float R[9];
    int32_t level;
    float tc[4];
    int32_t pad[2];
};

} // namespace goicp
