/*
 * goicp_b200.h -- C ABI of the B200-native Go-ICP registration engine.
 *
 * The reference (zjsun1017/CUDA-Go-ICP) has no FFI: its boundary is the C++ object protocol
 * that src/main.cpp and src/goicp_kernel.cu use on `class GoICP` (src/goicp/jly_goicp.h:82-141).
 * Every entry point below names the piece of that protocol it replaces.  Plain pointers and
 * sizes only; all buffers are HOST memory owned by the caller unless stated otherwise; the
 * library copies what it needs to the device.  Every call returns a goicp_status; nothing
 * exits the process or throws across the boundary (the reference exit()s on CUDA errors,
 * src/kernel.cu:29-38).  Calls on one handle must be serialised by the caller, except
 * goicp_poll() and goicp_cancel(), which may be called from another thread while
 * goicp_register() runs (replaces the unlocked polling of optR/optT/optError from the GL
 * thread, src/goicp_kernel.cu:82-149, and the global `goicp_finished`, jly_goicp.cpp:36,400).
 *
 * There is NO CPU fallback: every numeric entry point runs hand-written sm_100a CUDA and
 * returns GOICP_ERR_CUDA if no usable device is present.
 */
#ifndef GOICP_B200_H
#define GOICP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct goicp_handle goicp_handle;

typedef enum {
    GOICP_OK = 0,
    GOICP_ERR_INVALID = 1,      /* bad argument / call order (e.g. register before build_dt) */
    GOICP_ERR_CUDA = 2,         /* CUDA runtime failure or no device; see goicp_last_error() */
    GOICP_ERR_CAPACITY = 3,     /* a device queue/heap overflowed its configured capacity */
    GOICP_ERR_DEPTH = 4,        /* search needed rotation level >= 20 (reference reads maxRotDis[20] out of bounds, jly_goicp.cpp:551) or translation level > 21 */
    GOICP_ERR_IO = 5,           /* file / TOML problems (std::runtime_error in the reference, src/common.cpp:25-226) */
    GOICP_ERR_CANCELLED = 6
} goicp_status;

/* Which exit the search took (jly_goicp.cpp:402-406, :416-420, :527-530, :400). */
typedef enum {
    GOICP_EXIT_NONE = 0,
    GOICP_EXIT_CERTIFIED = 1,         /* optError - lowest lb <= SSEThresh : global-optimality certificate */
    GOICP_EXIT_EARLY_SSE = 2,         /* optError < SSEThresh right after an ICP refinement */
    GOICP_EXIT_QUEUE_EMPTY = 3,
    GOICP_EXIT_CANCELLED = 4
} goicp_exit_path;

typedef enum {
    GOICP_DT_REFERENCE = 0,   /* bit-exact with DT3D::Build's sequential vector propagation (jly_3ddt.cpp:710-742) */
    GOICP_DT_EXACT_EDT = 1,   /* exact Euclidean DT (fast, fully parallel); differs from the reference on a few voxels */
    GOICP_DT_EXACT_EDT_REFSEED = 2   /* exact EDT of the seeds the reference binary really has: the model voxels plus voxel (0,0,0), which
                                        its mask function seeds by returning an uninitialised struct (jly_3ddt.cpp:469-470).
                                        Differs from GOICP_DT_REFERENCE only where the reference's propagation is not exact (~1e-5 of the voxels) */
} goicp_dt_mode;

/* Where float sums are formed in the reference's own order (bit flags; 0 = everywhere a decision hinges on them). */
typedef enum {
    GOICP_NUM_STRICT = 0,
    GOICP_NUM_FAST_SUMS = 1,  /* near-tied upper bounds of an improving pass and the DT scores (optError itself) keep their fixed-order
                                 tree sums: no single-thread emulation of intro_select + sequential adds (jly_goicp.cpp:293-315) */
    GOICP_NUM_FAST_ICP = 2,   /* ICP3D::Run: means / covariance by parallel reduction and a Jacobi 3x3 SVD instead of the reference's
                                 sorted sequential accumulations and Matrix::svd (jly_icp3d.hpp:238-291); NN indices stay bit-exact */
    GOICP_NUM_JACOBI_SVD = 4  /* reference-order ICP sums, but the Procrustes rotation from the Jacobi solver (an experiment switch:
                                 shows that Matrix::svd's own rounding is part of the trajectory, profiles/r2_parity_modes.md) */
} goicp_numerics;

/* Which search goicp_register runs. */
typedef enum {
    GOICP_SEARCH_GOICP = 0,   /* GoICP::Register of src/goicp (jly_goicp.cpp:342-585): angle-axis rotation cubes, nested translation BnB,
                                 results identical to the reference's -- the default, and the only mode with an executable oracle */
    GOICP_SEARCH_FGOICP = 1   /* the strategy of the reference's own GPU path, icp::FastGoICP (src/fgoicp/fgoicp.cpp:9-181): unit-quaternion
                                 cube [-1,1]^3 with its SO(3) overlap tests, rotation span cut-off 0.1 and translation span cut-off 0.12,
                                 translation domain [-1,1]^3, the relaxed trigger "ub < 2 * best => ICP" (500 iterations, relative
                                 stop 1e-3), error = sum of squared NEAREST-NEIGHBOUR distances.  Bounds come from this library's distance
                                 transform instead of fgoicp's 300^3 trilinear table over [0,1]^3 (registration.hpp:67, "TODO: needs fix"),
                                 so it is a compatible strategy, not a bit-compatible restatement */
} goicp_search_mode;

/* Replaces the public tunables of class GoICP (jly_goicp.h:85-120) and its constructor
 * defaults (jly_goicp.cpp:40-72). */
typedef struct {
    float  mse_threshold;    /* GoICP(float mse_threshold); SSEThresh = mse*inlierNum (jly_goicp.cpp:208) */
    float  trim_fraction;    /* GoICP::trimFraction (default 0) */
    int    do_trim;          /* GoICP::doTrim (default 1) */
    int    dt_size;          /* GoICP::dt.SIZE (default 300) */
    double dt_expand;        /* GoICP::dt.expandFactor (default 2.0) */
    float  rot_cube[4];      /* initNodeRot a,b,c,w  (default -pi,-pi,-pi,2pi) */
    float  trans_cube[4];    /* initNodeTrans x,y,z,w (default -0.5,-0.5,-0.5,1) */
    int    icp_max_iter;     /* ICP3D::max_iter_def (10000, jly_icp3d.hpp:113) */
    int    device;           /* CUDA device ordinal */
    int    spec_cubes;       /* rotation cubes expanded speculatively per round (0 = auto) */
    int    cluster_size;     /* CTAs (SMs) cooperating on one translation BnB via a thread-block cluster (0 = auto, max 16) */
    int    dt_mode;          /* goicp_dt_mode used by goicp_build_dt (default GOICP_DT_EXACT_EDT_REFSEED: on all of the reference's golden
                                runs it yields the results of GOICP_DT_REFERENCE bit for bit, 200x sooner -- profiles/r2_parity_modes.md) */
    /* multi-GPU sharding of the rotation frontier (all ranks hold identical inputs) */
    int    rank, world_size;
    int    numerics;         /* goicp_numerics bit flags (default 0: strict) */
    int    search_mode;      /* goicp_search_mode (default GOICP_SEARCH_GOICP) */
} goicp_params;

/* Replaces the public results of class GoICP (optR/optT/optError/optNodeRot/optNodeTrans,
 * jly_goicp.h:107-120) plus the counters the reference prints (tNodeCount/rNodeCount,
 * jly_goicp.cpp:34-35,579-580). */
typedef struct {
    float  R[9];             /* row-major, x' = R x + t maps data -> model (jly_goicp.cpp:105-107) */
    float  t[3];
    float  sse;              /* optError */
    float  sse_thresh;       /* SSEThresh */
    float  best_lb;          /* lower bound of the cube popped at the certificate (0 otherwise) */
    int    exit_path;        /* goicp_exit_path */
    int64_t rot_pops;        /* rNodeCount */
    int64_t trans_pops;      /* tNodeCount (committed = what the sequential reference would count) */
    int64_t bound_evals;     /* committed bound evaluations (body of jly_goicp.cpp:265-336) */
    int64_t bound_evals_executed;   /* incl. speculative work that was later discarded */
    int64_t icp_calls;
    int64_t rounds;          /* device rounds (one batch of rotation cubes each) */
    int64_t kernel_launches; /* CUDA kernels this call launched */
    double seconds_total, seconds_bnb_kernels, seconds_icp;
    /* accounting (not in the reference) */
    int64_t bound_evals_executed_local;  /* the part of bound_evals_executed this rank's GPU ran (== bound_evals_executed on one GPU) */
    int64_t strict_resolves;             /* upper-bound passes whose near-tied contenders were re-evaluated in the reference's order */
    int64_t contender_overflows;         /* ... of which the kernel could not keep every contender (more than 128 within rounding of the
                                            minimum): the arg-min was then chosen among the 128 kept -- reported, never silent */
    double seconds_dt_score, seconds_strict, seconds_setup;   /* DT scoring of poses / strict re-evaluations / upload + kd-tree */
    double seconds_host;                 /* the rest of seconds_total: the host side of the rotation frontier (queue, prune, commit in the
                                            reference's order, task lists) -- the part a device-side frontier would absorb */
    int64_t bnb_kernel_variants;         /* which translation-BnB kernels ran: bit 0 inner_bnb_pipelined_kernel<1,1> (points in shared memory,
                                            low-latency), bit 1 <1,0>, bit 2 <0,1>, bit 3 <0,0>, bit 4 inner_bnb_kernel (trimming / GOICP_NO_PIPELINE),
                                            bit 5 / 6 the dense shape <1,0,192,5> / <0,0,192,5> (192 threads, five CTAs per SM) */
} goicp_result;

typedef struct {
    float R[9], t[3];
    float err;               /* return value of ICP3D::Run: sum of squared NN distances of the last iteration */
    int   iterations;
} goicp_icp_result;

typedef struct {
    float R[9], t[3];        /* best so far (optR/optT) */
    float sse;               /* optError */
    int64_t rot_pops, trans_pops, bound_evals;
    int   finished;          /* GoICP::finished */
} goicp_snapshot;

/* One inner (translation) branch-and-bound: replaces GoICP::InnerBnB (jly_goicp.cpp:227-340). */
typedef struct {
    float value;             /* returned optErrorT */
    float node[4];           /* nodeTransOut x,y,z,w (upper-bound pass only) */
    uint32_t pops, evals;
    int32_t status;
    /* For which SMALLER starting optErrors E' the call would have run exactly the same way (same pops, evaluations, arg-min;
       `value` unchanged unless it is the starting optError itself, which becomes E'): E' > reuse_gt and, in float,
       E' - reuse_poplb >= SSEThresh.  This is what lets goicp_register keep speculative results across an improvement of
       the incumbent; FLT_MAX when the call cannot vouch for itself (trimming kernel, overflow). */
    float reuse_gt, reuse_poplb;
} goicp_inner_result;

void goicp_default_params(goicp_params* p);

/* GoICP::GoICP + ~GoICP */
int goicp_create(const goicp_params* p, goicp_handle** out);
int goicp_destroy(goicp_handle* h);
const char* goicp_last_error(const goicp_handle* h);

/* pModel/Nm, pData/Nd (jly_goicp.h:85-86; borrowed glm::vec3* in the reference, main.cpp:50-53):
 * packed xyz float triples, copied. */
int goicp_set_model(goicp_handle* h, const float* xyz, int n);
int goicp_set_data(goicp_handle* h, const float* xyz, int n);

/* GoICP::BuildDT (jly_goicp.cpp:75-90) on the GPU.  In GOICP_DT_REFERENCE mode, when the data cloud is already set (the
 * order of main.cpp:47-57), the ICP from the identity pose that GoICP::Register begins with (jly_goicp.cpp:378-391; it
 * reads no DT) runs on the SMs the single-CTA DT propagation leaves idle; the first goicp_register after this call picks
 * its result up.  Setting either cloud again discards it.  GOICP_NO_PREFETCH=1 in the environment turns the overlap off. */
int goicp_build_dt(goicp_handle* h);
/* Install / read back a distance grid ([z][y][x] floats + {xMin,yMin,zMin,scale}); lets a caller
 * cache the model-only precompute across runs (the reference cannot). */
int goicp_set_dt(goicp_handle* h, const float* grid, int size, const double meta4[4]);
int goicp_get_dt(goicp_handle* h, float* grid_out, double meta4_out[4]);
int goicp_dt_size(const goicp_handle* h);

/* DT3D::Distance (jly_3ddt.cpp:981-1026) for n query points; optionally also the raw voxel
 * indices ROUND((q-min)*scale) (3 ints per query, un-clamped). */
int goicp_dt_distance(goicp_handle* h, const float* q_xyz, int n, float* dist_out, int32_t* ixyz_out);

/* Batched bound evaluation over (rotation cube x translation cube) pairs -- the unit of work
 * of the metric (loop body jly_goicp.cpp:265-336).  Pair k: rotation R9[9k..] (row-major, as
 * produced at jly_goicp.cpp:449-467), rotation level level[k] (-1: no rotation uncertainty, i.e.
 * the upper-bound pass), translation cube tcube[4k..] = x,y,z,w.  Outputs ub[k], lb[k]. */
int goicp_eval_bounds(goicp_handle* h, int npairs, const float* R9, const int32_t* level,
                      const float* tcube, float* ub_out, float* lb_out);

/* One expansion step of InnerBnB for n (rotation, PARENT translation cube) pairs: the 8 octant
 * children of tcube[4k..] are evaluated together (6 voxel-index computations serve 8 gathers per
 * point).  out16[16k..] = ub[0..7], lb[0..7] in the child order j of jly_goicp.cpp:265-269.
 * If device_ms != NULL the kernel is launched `repeats` times on resident inputs and the mean
 * CUDA-event time of one launch is returned (used by bench.py for the DT-gather roofline). */
int goicp_expand_bounds(goicp_handle* h, int n, const float* R9, const int32_t* level, const float* tcube,
                        float* out16, int repeats, float* device_ms);

/* n independent inner BnBs (GoICP::InnerBnB): rotation R9[9k..], level[k] (-1 = ub pass),
 * starting optError opt_error[k]. */
int goicp_inner_bnb(goicp_handle* h, int n, const float* R9, const int32_t* level,
                    const float* opt_error, goicp_inner_result* out);

/* Nearest model point of each query in the reference's kd-tree visiting order (ties included):
 * KDTreeSingleIndexAdaptor::knnSearch(k=1) (nanoflann_goicp.hpp:821-826). */
int goicp_nn(goicp_handle* h, const float* q_xyz, int n, int32_t* idx_out, float* d2_out);

/* The model kd-tree exactly as ICP3D::Build lays it out (jly_icp3d.hpp:129-154; divideTree / middleSplit_ / planeSplit,
 * nanoflann_goicp.hpp:927-1111, leaf size 10): built on the host, no device needed -- the layout decides which of several
 * equidistant points goicp_nn returns, so it is exposed for parity checks.  nodes7_out: 7 int32 per node {child1, child2,
 * left, right, divfeat, float bits of divlow, float bits of divhigh} (children -1/-1 and [left,right) into vind for a leaf),
 * node 0 = root, capacity_nodes entries available; vind_out: n indices; bbox6_out: root box lo xyz, hi xyz.
 * Returns the number of nodes (the tree is not written if that exceeds capacity_nodes), or a negative goicp status. */
int goicp_kdtree_host(const float* model_xyz, int n, int32_t* nodes7_out, int capacity_nodes, int32_t* vind_out, float* bbox6_out);

/* Matrix::svd (matrix.cpp:602-830) of n 3x3 float matrices (row-major, 9 floats each) exactly as ICP3D::Run uses it: U, singular
 * values (descending) and V with the reference's float roundings, sort and sign normalisation.  Exposed for parity tests. */
int goicp_svd3(goicp_handle* h, const float* H9, int n, float* U9_out, float* W3_out, float* V9_out);

/* intro_select (jly_sorting.hpp:228-313) of a[0..n) for position k, in place: the permutation the reference leaves behind decides the
 * order of its sequential residual sums (jly_goicp.cpp:293-315), so the strict kernels reproduce it -- with the partition sweep done by a
 * whole thread block (csrc/strict_sum.cuh).  threads: block size (multiple of 32, <= 1024); in_global != 0 keeps the array in device
 * global memory (the path of clouds beyond shared memory).  Exposed for parity tests. */
int goicp_intro_select(goicp_handle* h, float* a, int n, int k, int threads, int in_global);
/* Diagnostic (host only, no device needed): the engine's per-round choice of inner-BnB kernel shape -- 1 = low latency (512
   threads, one CTA per SM), 0 = 512 x 2, 2 = dense (192 x 5) -- from the forecast pops of the round's longest task and of
   all of this rank's tasks, the previous round's longest task, and the number of clusters a GPU holds in the low-latency
   shape (SMs / cluster size).  No reference counterpart: the reference runs its InnerBnB calls one after the other
   (jly_goicp.cpp:471-551); every shape returns the same bits, so this only decides speed. */
int goicp_bnb_shape_rule(double forecast_longest_pops, double forecast_sum_pops_this_rank, double previous_longest_pops, int cluster_slots);

/* ICP3D<float>::Run (jly_icp3d.hpp:180-295) from (R0,t0). err_diff<0 -> mse_threshold/10000
 * (jly_goicp.cpp:186); max_iter<=0 -> params.icp_max_iter. */
int goicp_icp(goicp_handle* h, const float R0[9], const float t0[3], int max_iter, float err_diff,
              goicp_icp_result* out);
/* GoICP::ICP (jly_goicp.cpp:93-132): ICP then (trimmed) DT re-scoring; returns the DT error. */
int goicp_icp_dt(goicp_handle* h, float R[9], float t[3], float* dt_error_out);
/* (trimmed) sum of squared DT distances under (R,t); R==NULL scores the raw data (jly_goicp.cpp:357-371). */
int goicp_dt_score(goicp_handle* h, const float* R, const float* t, float* sse_out);

/* GoICP::Register (jly_goicp.cpp:569-585): Initialize + OuterBnB; blocking. */
int goicp_register(goicp_handle* h, goicp_result* out);
int goicp_poll(goicp_handle* h, goicp_snapshot* out);
int goicp_cancel(goicp_handle* h);

/* Device and pinned buffers of destroyed handles are cached process-wide and reused by the next
 * handle (the reference new/deletes its 324 MB grid per GoICP object, jly_3ddt.cpp:932-935; here a
 * cudaMalloc costs a driver lock).  This returns the cached blocks to the driver. */
int goicp_trim_memory(void);

/* Measurement helpers (bench.py).  goicp_transfer_bytes: host->device / device->host bytes this handle has copied so far
 * (counted at every copy the library issues).  goicp_measure_gather: rate of uniformly random 4-byte loads over a device
 * buffer of `bytes` bytes with no other work -- the measured ceiling of the distance-transform gathers (L2-resident for a
 * 300^3 grid, HBM-bound for 512^3); returns look-ups per second. */
int goicp_transfer_bytes(const goicp_handle* h, int64_t* h2d_out, int64_t* d2h_out);
int goicp_measure_gather(goicp_handle* h, size_t bytes, int repeats, double* lookups_per_s_out);

/* Multi-GPU: every rank evaluates its slice of each round's cubes; `exchange` must all-gather
 * `bytes_per_rank` bytes from every rank into recv (rank-major).  The Python driver implements it
 * with torch.distributed (NCCL on device buffers is_device=1, gloo on host buffers). */
typedef int (*goicp_allgather_fn)(void* user, const void* send, void* recv, size_t bytes_per_rank, int is_device);
int goicp_set_exchange(goicp_handle* h, goicp_allgather_fn fn, void* user, int use_device_buffers);

/* Native exchange over NCCL (NVLink / NVSwitch): rank 0 obtains a 128-byte ncclUniqueId, hands it to the
 * other ranks by any means (torch.distributed broadcast, MPI, a file), then every rank calls goicp_nccl_init
 * (collective).  From then on goicp_register shards each round's inner BnBs over the ranks and all-gathers
 * the 48-byte result records with ncclAllGather on the engine's own stream, right behind the kernel; every
 * rank commits the same results in the same order, so the best upper bound (the prune threshold), the
 * rotation queue and the certificate stay replicated -- this is the per-round best-bound exchange of
 * SURVEY.md section 8e.  libnccl.so.2 is loaded on demand; single-GPU use never touches it. */
int goicp_nccl_unique_id(void* id128_out);
int goicp_nccl_init(goicp_handle* h, const void* id128, int rank, int world_size);

/* Host-only self test of the sharding + exchange plumbing used by multi-GPU rounds (no GPU needed): the round-robin
 * deal and all-gather of n result records, and the hand-round of a contender list from each rank in turn. */
int goicp_selftest_shard(int rank, int world, int n, goicp_allgather_fn fn, void* user, int* mismatches);

/* Convenience driver over the reference's TOML keys (src/common.cpp:39-74) and cloud formats
 * (src/common.cpp:79-228): loads [io].target/source (.txt / .ply), applies subsample (seeded)
 * and resize, builds the DT, registers, writes [io].output if non-empty. */
int goicp_run_toml(const char* toml_path, unsigned seed_model, unsigned seed_data, goicp_result* out);
/* load_cloud (src/common.cpp:205-228) with a seeded subsample; *xyz_out is malloc'ed, release with goicp_free_cloud. */
int goicp_load_cloud(const char* path, float subsample, float resize, unsigned seed, float** xyz_out, int* n_out);
void goicp_free_cloud(float* xyz);
const char* goicp_io_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* GOICP_B200_H */
