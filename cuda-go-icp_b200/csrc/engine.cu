// engine.cu -- host side of the B200 Go-ICP engine and its C ABI (include/goicp_b200.h).
//
// Mirrors the object protocol of the reference's `class GoICP` (jly_goicp.h:82-141):
// set clouds -> BuildDT -> Register (Initialize + OuterBnB) -> optR/optT/optError.
//
// Search strategy.  The reference's outer loop is a strictly sequential best-first search whose
// result (which cube triggers which ICP refinement) depends on the visiting order, including
// the order std::priority_queue gives to nodes with equal (lb, w).  Between two improvements of
// the best error E, however, every inner BnB is a pure function of (rotation cube, pass, E).  The
// engine therefore keeps the reference's queue semantics *exactly* (same libstdc++ heap
// algorithms, same commit order) but evaluates cubes ahead of time: each device round expands
// the cube at the top of the queue plus the next-best `spec_cubes-1` queue entries, all
// (child, pass) inner BnBs of a round running concurrently, one CTA each.  Results are cached
// per cube and committed strictly in the reference's order; an improvement of E invalidates
// the cache (epoch bump).  The committed trajectory, counters and certificate are those of the
// sequential algorithm; speculative work only shows up in bound_evals_executed.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/goicp_b200.h"
#include "goicp_kernels.h"
#include "kdtree_host.h"
#include "dt_kernels.h"
#include "mem_pool.h"

using namespace goicp;

namespace {

constexpr double kPi = 3.1415926536;       // PI,    jly_goicp.h:35
constexpr double kSqrt3 = 1.732050808;     // SQRT3, jly_goicp.h:36

double now_s()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct RotNode {                                                       // ROTNODE, jly_goicp.h:44-58
    float a, b, c, w, ub, lb; int l;
    uint32_t pops_ub = 0, pops_lb = 0;   // not in the reference: translation pops of this cube's own two passes, a cost hint for its children's
};
struct RotLower {                                                      // its operator<
    bool operator()(const RotNode& n1, const RotNode& n2) const { return n1.lb != n2.lb ? n1.lb > n2.lb : n1.w < n2.w; }
};

struct CubeKey {
    uint32_t a, b, c; int l;
    bool operator==(const CubeKey& o) const { return a == o.a && b == o.b && c == o.c && l == o.l; }
};
struct CubeKeyHash {
    size_t operator()(const CubeKey& k) const
    {
        uint64_t h = 1469598103934665603ULL;
        for (uint32_t v : {k.a, k.b, k.c, (uint32_t)k.l}) { h ^= v; h *= 1099511628211ULL; }
        return (size_t)h;
    }
};
inline uint32_t fbits(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline CubeKey key_of(const RotNode& n) { return CubeKey{fbits(n.a), fbits(n.b), fbits(n.c), n.l}; }

struct ChildEval {
    bool skipped = true;            // outside the pi-ball (jly_goicp.cpp:443)
    float R[9];
    InnerResult ub, lb;
    bool have_ub = false, have_lb = false;   // result present and valid under the current incumbent
    float ub_opt_error = 0, lb_opt_error = 0;   // optError the pass was run with (or has been re-validated for)
    std::shared_ptr<CandList> cands;    // arg-min contenders of the ub pass (when it has any)
};
struct CubeEval {
    long epoch = -1; bool committed = false; ChildEval ch[8];
    bool complete() const { for (const ChildEval& c : ch) if (!c.skipped && !(c.have_ub && c.have_lb)) return false; return true; }
};
// Would this inner BnB, run with initial optErrorT = e_old, have gone exactly the same way with e_new < e_old?  (InnerResult::reuse_*)
inline bool reusable_under(const InnerResult& r, float e_new, float sse_thresh)
{
    return r.status == 0 && e_new > r.reuse_gt && !((float)(e_new - r.reuse_poplb) < sse_thresh);
}

// Angle-axis vector of the cube centre -> rotation matrix, in float with glibc sinf/cosf exactly
// like jly_goicp.cpp:437-467.  false = cube wholly outside the pi-ball.
bool cube_rotation(float a, float b, float c, float w, float* R)
{
    float v1 = a + w / 2, v2 = b + w / 2, v3 = c + w / 2;
    if ((double)std::sqrt(v1 * v1 + v2 * v2 + v3 * v3) - kSqrt3 * (double)w / 2 > kPi) return false;
    float t = std::sqrt(v1 * v1 + v2 * v2 + v3 * v3);
    if (t > 0) {
        v1 /= t; v2 /= t; v3 /= t;
        const float ct = std::cos(t), ct2 = 1 - ct, st = std::sin(t);
        const float tmp121 = v1 * v2 * ct2, tmp122 = v3 * st, tmp131 = v1 * v3 * ct2, tmp132 = v2 * st,
                    tmp231 = v2 * v3 * ct2, tmp232 = v1 * st;
        R[0] = ct + v1 * v1 * ct2; R[1] = tmp121 - tmp122;    R[2] = tmp131 + tmp132;
        R[3] = tmp121 + tmp122;    R[4] = ct + v2 * v2 * ct2; R[5] = tmp231 - tmp232;
        R[6] = tmp131 - tmp132;    R[7] = tmp231 + tmp232;    R[8] = ct + v3 * v3 * ct2;
    } else {
        for (int i = 0; i < 9; i++) R[i] = (i % 4 == 0) ? 1.0f : 0.0f;
    }
    return true;
}

template <typename T>
struct DevBuf {
    T* p = nullptr; size_t n = 0;
    cudaError_t reserve(size_t count)
    {
        if (count <= n) return cudaSuccess;
        if (p) pool_free(p);
        p = nullptr; n = 0;
        cudaError_t e = pool_alloc((void**)&p, count * sizeof(T));
        if (e == cudaSuccess) n = count;
        return e;
    }
    void release() { if (p) pool_free(p); p = nullptr; n = 0; }
};

} // namespace

// ---------------------------------------------------------------------------------------------
// NCCL, bound at run time (dlopen): a single-GPU user needs no NCCL at all, and inside a
// torch.distributed process the already-loaded libnccl.so.2 is picked up by soname.  Only the
// stable C ABI is used: ncclGetUniqueId / ncclCommInitRank / ncclAllGather / ncclCommDestroy.
// ---------------------------------------------------------------------------------------------
namespace {
struct NcclUniqueId { char internal[128]; };             // ncclUniqueId (NCCL_UNIQUE_ID_BYTES = 128)
typedef struct ncclComm* NcclComm;
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    std::string err;
};
NcclApi* nccl_api()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) { api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (api.lib) break; }
        if (!api.lib) { api.err = std::string("cannot load libnccl.so.2: ") + dlerror(); return; }
        api.GetUniqueId = (int (*)(NcclUniqueId*))dlsym(api.lib, "ncclGetUniqueId");
        api.CommInitRank = (int (*)(NcclComm*, int, NcclUniqueId, int))dlsym(api.lib, "ncclCommInitRank");
        api.AllGather = (int (*)(const void*, void*, size_t, int, NcclComm, cudaStream_t))dlsym(api.lib, "ncclAllGather");
        api.CommDestroy = (int (*)(NcclComm))dlsym(api.lib, "ncclCommDestroy");
        api.GetErrorString = (const char* (*)(int))dlsym(api.lib, "ncclGetErrorString");
        if (!api.GetUniqueId || !api.CommInitRank || !api.AllGather || !api.CommDestroy) api.err = "libnccl is missing a required symbol";
    });
    return &api;
}
constexpr int kNcclUint8 = 1;                            // ncclUint8
} // namespace

struct goicp_handle {
    goicp_params p;
    std::string err;
    bool cuda_ready = false;
    int sm_count = 0, max_smem_optin = 0, inner_dyn_smem = 0;
    cudaStream_t stream = nullptr;

    std::vector<float> model, data;     // xyz triples
    int nm = 0, nd = 0;

    // distance transform
    DevBuf<float> d_dt; int dt_size = 0; double dt_meta[4] = {0, 0, 0, 0}; bool have_dt = false;
    cudaTextureObject_t dt_tex = 0; const float* dt_tex_ptr = nullptr; int dt_tex_size = 0;      // the grid as a linear texture (dt_texture())
    // data cloud on device (x,y,z,norm)
    DevBuf<float4> d_data; bool data_uploaded = false;
    // kd-tree
    std::vector<unsigned short> grid_start; int gdim[3] = {0, 0, 0}; float gh = 0; DevBuf<unsigned short> d_grid_start; DevBuf<float4> d_grid_pts;     // NN grid (small models)
    HostKdTree kd_host; DevBuf<KdNode> d_kd_nodes; DevBuf<float4> d_kd_boxes; DevBuf<int32_t> d_kd_vind; DevBuf<float4> d_kd_leaf; DevBuf<float> d_model;
    bool kd_ready = false;
    // scratch
    DevBuf<InnerTask> d_tasks; DevBuf<InnerResult> d_results; DevBuf<HeapEntry> d_spill; int spill_cap = 0; int spill_slots = 0;
    DevBuf<CandList> d_cands; DevBuf<unsigned> d_trim_keys; DevBuf<unsigned long long> d_dbg; DevBuf<float> d_strict;   // d_strict: [0..127] strict sums, [128..135] pick result, then optional scratch
    int64_t strict_resolves = 0, cand_overflows = 0, bnb_variants = 0;
    DevBuf<PairTask> d_pairs; DevBuf<float> d_f32a, d_f32b, d_score_scratch; DevBuf<int32_t> d_i32; DevBuf<float> d_q;
    DevBuf<IcpState> d_icp_state; DevBuf<float> d_icp_q, d_icp_d2, d_icp_stage, d_icp_partials; DevBuf<double> d_icp_xch; DevBuf<int32_t> d_icp_nn, d_icp_pos, d_icp_order; DevBuf<unsigned long long> d_icp_keys, d_icp_keys2; DevBuf<unsigned> d_icp_hist; int icp_blocks = 0;
    InnerResult* h_results = nullptr; size_t h_results_n = 0;       // pinned
    InnerTask* h_tasks = nullptr; size_t h_tasks_n = 0;             // pinned

    // Initialize() products
    float sse_thresh = 0; int inlier_num = 0; float cgamma[kMaxRotLevel]; bool initialized = false;

    // live state for poll/cancel
    std::mutex snap_mtx; goicp_snapshot snap; std::atomic<int> cancel_flag{0};

    // multi-GPU exchange
    goicp_allgather_fn xchg = nullptr; void* xchg_user = nullptr; int xchg_device = 0;
    bool low_latency = true;             // inner-BnB kernel variant of the next round (see run_inner_batch)
    int dense = 0;                       // when not low_latency: 0 = 512 threads x 2 CTAs per SM, 2 = 192 threads x 5, 3 = 128 x 8
    double prev_max_pops = 0;            // longest inner BnB of the previous batch (pops)
    double pred_max = 0, pred_sum = 0;   // expected cost of the next batch's longest task / of all its tasks (parent pops; 0 = unknown)
    NcclComm nccl = nullptr; DevBuf<InnerResult> d_gather; DevBuf<unsigned char> d_share;      // native exchange: all-gather of the round's result records on the stream

    // timing
    double t_kernels = 0, t_icp = 0, t_score = 0, t_strict = 0, t_setup = 0;     // t_score.. only for GOICP_ROUND_STATS
    // ICP from the identity pose, run by goicp_build_dt next to the distance-transform build (it needs no DT) and
    // consumed by the first goicp_register after it
    cudaStream_t stream_dt = nullptr; bool icp0_valid = false; goicp_icp_result icp0;
    int64_t launches = 0;
    int64_t bytes_h2d = 0, bytes_d2h = 0;      // cumulative over the handle's life (xfer)
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};

namespace {

int fail(goicp_handle* h, int code, const std::string& msg) { if (h) h->err = msg; return code; }
#define CUDA_TRY(h, expr)                                                                         \
    do { cudaError_t _e = (expr);                                                                 \
         if (_e != cudaSuccess) return fail(h, GOICP_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); } while (0)

// every host<->device copy of a handle goes through here, so that the bytes a call really moved can be reported
// (goicp_transfer_bytes; bench.py's e2e.h2d_bytes_per_step / d2h_bytes_per_step are read from it, not estimated)
inline cudaError_t xfer(goicp_handle* h, void* dst, const void* src, size_t bytes, cudaMemcpyKind kind, cudaStream_t s)
{
    if (kind == cudaMemcpyHostToDevice) h->bytes_h2d += (int64_t)bytes; else if (kind == cudaMemcpyDeviceToHost) h->bytes_d2h += (int64_t)bytes;
    return cudaMemcpyAsync(dst, src, bytes, kind, s);
}

// Per-device facts and launch configuration are looked up once per process, and the streams / events of destroyed
// handles are kept for the next one: cudaGetDeviceProperties, cudaFuncSetAttribute and stream creation take the
// driver's global lock, and a registration through a fresh handle spent 3-110 ms here (measured; whenever a
// monitoring process happened to hold that lock) before any work started.
struct DeviceFacts { bool ready = false; int sm_count = 0, max_smem_optin = 0, inner_dyn_smem = 0; };
struct StreamSet { cudaStream_t stream = nullptr, stream_dt = nullptr; cudaEvent_t ev0 = nullptr, ev1 = nullptr; };
std::mutex g_dev_mtx;
std::unordered_map<int, DeviceFacts> g_dev_facts;
std::unordered_map<int, std::vector<StreamSet>> g_stream_sets;

int ensure_cuda(goicp_handle* h)
{
    if (h->cuda_ready) return GOICP_OK;
    std::lock_guard<std::mutex> lk(g_dev_mtx);
    DeviceFacts& f = g_dev_facts[h->p.device];
    if (!f.ready) {
        int count = 0;
        cudaError_t e = cudaGetDeviceCount(&count);
        if (e != cudaSuccess || count == 0)
            return fail(h, GOICP_ERR_CUDA, std::string("no CUDA device available (this engine has no CPU fallback): ") + cudaGetErrorString(e));
        CUDA_TRY(h, cudaSetDevice(h->p.device));
        int major = 0;
        CUDA_TRY(h, cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, h->p.device));
        if (major < 10) return fail(h, GOICP_ERR_CUDA, "device is not sm_100-class; this library ships sm_100a code only");
        CUDA_TRY(h, cudaDeviceGetAttribute(&f.sm_count, cudaDevAttrMultiProcessorCount, h->p.device));
        CUDA_TRY(h, cudaDeviceGetAttribute(&f.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->p.device));
        CUDA_TRY(h, inner_bnb_configure(f.max_smem_optin, &f.inner_dyn_smem));
        f.ready = true;
    } else {
        CUDA_TRY(h, cudaSetDevice(h->p.device));
    }
    h->sm_count = f.sm_count; h->max_smem_optin = f.max_smem_optin; h->inner_dyn_smem = f.inner_dyn_smem;
    std::vector<StreamSet>& spare = g_stream_sets[h->p.device];
    if (!spare.empty()) {
        const StreamSet ss = spare.back(); spare.pop_back();
        h->stream = ss.stream; h->stream_dt = ss.stream_dt; h->ev0 = ss.ev0; h->ev1 = ss.ev1;
    } else {
        CUDA_TRY(h, cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        CUDA_TRY(h, cudaStreamCreateWithFlags(&h->stream_dt, cudaStreamNonBlocking));
        CUDA_TRY(h, cudaEventCreate(&h->ev0));
        CUDA_TRY(h, cudaEventCreate(&h->ev1));
    }
    h->cuda_ready = true;
    return GOICP_OK;
}

int upload_data(goicp_handle* h)
{
    if (h->data_uploaded) return GOICP_OK;
    if (h->nd <= 0) return fail(h, GOICP_ERR_INVALID, "no data cloud set");
    std::vector<float4> tmp(h->nd);
    for (int i = 0; i < h->nd; i++) {
        const float x = h->data[3 * i], y = h->data[3 * i + 1], z = h->data[3 * i + 2];
        tmp[i] = make_float4(x, y, z, std::sqrt(x * x + y * y + z * z));     // normData (jly_goicp.cpp:143-147)
    }
    CUDA_TRY(h, h->d_data.reserve(h->nd));
    CUDA_TRY(h, xfer(h, h->d_data.p, tmp.data(), sizeof(float4) * h->nd, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->data_uploaded = true;
    return GOICP_OK;
}

int ensure_kdtree(goicp_handle* h)
{
    if (h->kd_ready) return GOICP_OK;
    if (h->nm <= 0) return fail(h, GOICP_ERR_INVALID, "no model cloud set");
    h->kd_host.build(h->model.data(), h->nm, 10);                              // ICP3D::Build (jly_icp3d.hpp:129-154)
    std::vector<float4> leaf(h->nm);
    for (int i = 0; i < h->nm; i++) {
        const int id = h->kd_host.vind[i];
        float w; std::memcpy(&w, &id, 4);
        leaf[i] = make_float4(h->model[3 * id], h->model[3 * id + 1], h->model[3 * id + 2], w);
    }
    CUDA_TRY(h, h->d_kd_nodes.reserve(h->kd_host.nodes.size()));
    CUDA_TRY(h, h->d_kd_boxes.reserve(2 * h->kd_host.nodes.size()));
    CUDA_TRY(h, h->d_kd_vind.reserve(h->nm));
    CUDA_TRY(h, h->d_kd_leaf.reserve(h->nm));
    CUDA_TRY(h, h->d_model.reserve((size_t)3 * h->nm));
    CUDA_TRY(h, xfer(h, h->d_kd_nodes.p, h->kd_host.nodes.data(), sizeof(KdNode) * h->kd_host.nodes.size(), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, xfer(h, h->d_kd_boxes.p, h->kd_host.boxes.data(), sizeof(float) * h->kd_host.boxes.size(), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, xfer(h, h->d_kd_vind.p, h->kd_host.vind.data(), sizeof(int32_t) * h->nm, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, xfer(h, h->d_kd_leaf.p, leaf.data(), sizeof(float4) * h->nm, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, xfer(h, h->d_model.p, h->model.data(), sizeof(float) * 3 * h->nm, cudaMemcpyHostToDevice, h->stream));
    // NN grid of the linear-scan range (icp_kernels.cu: grid_nn): cells of edge h = longest bounding-box side / G,
    // G ~ sqrt(nm / 8) (a surface-like cloud then has a few points per occupied cell); points binned with the same float
    // expression the kernel applies to a query.
    h->grid_start.clear(); h->gdim[0] = h->gdim[1] = h->gdim[2] = 0;
    if (h->nm <= 16384 && h->nm >= 64) {
        int G = (int)std::lround(std::sqrt((double)h->nm / 8.0));
        if (const char* e = getenv("GOICP_NN_GRID")) G = atoi(e);
        if (G >= 2) {
            G = std::min(G, 40);
            const float* lo = h->kd_host.bb_lo; const float* hi = h->kd_host.bb_hi;
            float ext = std::max(std::max(hi[0] - lo[0], hi[1] - lo[1]), hi[2] - lo[2]);
            if (ext > 0) {
                const float gh = ext / (float)G, inv = 1.0f / gh;
                for (int a = 0; a < 3; a++) h->gdim[a] = std::max(1, std::min(G, (int)std::floor((hi[a] - lo[a]) * inv) + 1));
                const int nc = h->gdim[0] * h->gdim[1] * h->gdim[2];
                std::vector<int> cell(h->nm); std::vector<int> count(nc + 1, 0);
                for (int i = 0; i < h->nm; i++) {
                    int c[3];
                    for (int a = 0; a < 3; a++) c[a] = std::min(std::max((int)std::floor((h->model[3 * i + a] - lo[a]) * inv), 0), h->gdim[a] - 1);
                    cell[i] = (c[2] * h->gdim[1] + c[1]) * h->gdim[0] + c[0];
                    count[cell[i] + 1]++;
                }
                for (int c = 0; c < nc; c++) count[c + 1] += count[c];
                h->grid_start.assign(count.begin(), count.end());
                std::vector<float4> gp(h->nm); std::vector<int> fill(count.begin(), count.end() - 1);
                for (int i = 0; i < h->nm; i++) {
                    float w; std::memcpy(&w, &i, 4);
                    gp[fill[cell[i]]++] = make_float4(h->model[3 * i], h->model[3 * i + 1], h->model[3 * i + 2], w);
                }
                h->gh = gh;
                CUDA_TRY(h, h->d_grid_start.reserve(h->grid_start.size() + 8));
                CUDA_TRY(h, h->d_grid_pts.reserve(h->nm));
                CUDA_TRY(h, xfer(h, h->d_grid_start.p, h->grid_start.data(), sizeof(unsigned short) * h->grid_start.size(), cudaMemcpyHostToDevice, h->stream));
                CUDA_TRY(h, xfer(h, h->d_grid_pts.p, gp.data(), sizeof(float4) * h->nm, cudaMemcpyHostToDevice, h->stream));
                CUDA_TRY(h, cudaStreamSynchronize(h->stream));
            }
        }
    }
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    CUDA_TRY(h, h->d_icp_state.reserve(1));
    h->kd_ready = true;
    return GOICP_OK;
}

KdView kd_view(const goicp_handle* h)
{
    KdView v;
    v.nodes = h->d_kd_nodes.p; v.boxes = h->d_kd_boxes.p; v.vind = h->d_kd_vind.p; v.pts_leaf = h->d_kd_leaf.p; v.model = h->d_model.p; v.nm = h->nm;
    for (int i = 0; i < 3; i++) { v.bb_lo[i] = h->kd_host.bb_lo[i]; v.bb_hi[i] = h->kd_host.bb_hi[i]; }
    v.grid_start = nullptr; v.grid_pts = nullptr; v.gcells = 0; v.gh = 0; v.ginv_h = 0;
    for (int i = 0; i < 3; i++) { v.gdim[i] = 0; v.glo[i] = h->kd_host.bb_lo[i]; }
    if (!h->grid_start.empty()) {
        v.grid_start = h->d_grid_start.p; v.grid_pts = h->d_grid_pts.p;
        for (int i = 0; i < 3; i++) v.gdim[i] = h->gdim[i];
        v.gcells = h->gdim[0] * h->gdim[1] * h->gdim[2]; v.gh = h->gh; v.ginv_h = 1.0f / h->gh;
    }
    // the 32 nodes five levels down, when every node above them is interior: the first five rounds of a breadth-first
    // search from the root would keep 1, 2, 4, 8, 16 lanes busy
    v.n_top = 0;
    {
        std::vector<int> cur{0};
        bool ok = !h->kd_host.nodes.empty();
        for (int d = 0; d < 5 && ok; d++) {
            std::vector<int> next;
            for (int n : cur) {
                const KdNode& k = h->kd_host.nodes[n];
                if (k.child1 < 0 || k.child2 < 0) { ok = false; break; }
                next.push_back(k.child1); next.push_back(k.child2);
            }
            if (ok) cur.swap(next);
        }
        if (ok && cur.size() == 32) { v.n_top = 32; for (int i = 0; i < 32; i++) v.top[i] = cur[i]; }
        else for (int i = 0; i < 32; i++) v.top[i] = 0;
    }
    return v;
}

// Initialize() (jly_goicp.cpp:134-209): thresholds and the per-level rotation-uncertainty scale.
int initialize(goicp_handle* h)
{
    int rc = ensure_cuda(h); if (rc) return rc;
    rc = upload_data(h); if (rc) return rc;
    for (int i = 0; i < kMaxRotLevel; i++) {
        float sigma = (float)((double)h->p.rot_cube[3] / std::pow(2.0, i) / 2.0);
        float max_angle = (float)(kSqrt3 * (double)sigma);
        if ((double)max_angle > kPi) max_angle = (float)kPi;
        h->cgamma[i] = 2 * std::sin(max_angle / 2);              // maxRotDis[i][j] = cgamma[i]*normData[j]
    }
    h->inlier_num = h->p.do_trim ? (int)((float)h->nd * (1 - h->p.trim_fraction)) : h->nd;
    h->sse_thresh = h->p.mse_threshold * (float)h->inlier_num;
    h->initialized = true;
    return GOICP_OK;
}

// The DT grid as a 1-D linear texture: the bound kernels fetch their look-ups through the TEX pipe (same values, same L1TEX
// sector rate -- measured 296 vs 293 G look-ups/s for the plain gather kernel), which leaves the LSU pipe to the shared-memory
// traffic and shuffles of the inner BnB's owner warp: its queue maintenance no longer queues behind the CTA's own gathers
// (7.2 k -> 5.7 k cycles per expansion).  GOICP_DT_TEX=0 keeps the plain loads (A/B, and the fallback
// for grids beyond the device's linear-texture width).  Returns 0 when no texture is in use.
cudaTextureObject_t dt_texture(goicp_handle* h)
{
    static const bool want = !(getenv("GOICP_DT_TEX") && atoi(getenv("GOICP_DT_TEX")) == 0);
    if (!want || !h->have_dt) return 0;
    if (h->dt_tex && (h->dt_tex_ptr != h->d_dt.p || h->dt_tex_size != h->dt_size)) { cudaDestroyTextureObject(h->dt_tex); h->dt_tex = 0; h->dt_tex_ptr = nullptr; }
    if (!h->dt_tex && (h->dt_tex_ptr != h->d_dt.p || h->dt_tex_size != h->dt_size)) {  // (one attempt per grid)
        h->dt_tex_ptr = h->d_dt.p; h->dt_tex_size = h->dt_size;
        const size_t n3 = (size_t)h->dt_size * h->dt_size * h->dt_size;
        cudaChannelFormatDesc fd = cudaCreateChannelDesc<float>();
        size_t max_width = 0;
        if (cudaDeviceGetTexture1DLinearMaxWidth(&max_width, &fd, h->p.device) != cudaSuccess || n3 > max_width) { cudaGetLastError(); return 0; }
        cudaResourceDesc rd; std::memset(&rd, 0, sizeof rd);
        rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = h->d_dt.p; rd.res.linear.desc = fd; rd.res.linear.sizeInBytes = n3 * sizeof(float);
        cudaTextureDesc td; std::memset(&td, 0, sizeof td);
        td.readMode = cudaReadModeElementType; td.filterMode = cudaFilterModePoint; td.addressMode[0] = cudaAddressModeClamp;
        if (cudaCreateTextureObject(&h->dt_tex, &rd, &td, nullptr) != cudaSuccess) { h->dt_tex = 0; cudaGetLastError(); }
    }
    return h->dt_tex;
}

int make_const(goicp_handle* h, BnbConst& c)
{
    if (!h->have_dt) return fail(h, GOICP_ERR_INVALID, "distance transform not built (call goicp_build_dt or goicp_set_dt first)");
    if (!h->initialized) { int rc = initialize(h); if (rc) return rc; }
    c.dt.grid = h->d_dt.p; c.dt.S = h->dt_size; c.dt.S2 = h->dt_size * h->dt_size;
    c.dt.tex = dt_texture(h);
    c.dt.xmin = h->dt_meta[0]; c.dt.ymin = h->dt_meta[1]; c.dt.zmin = h->dt_meta[2]; c.dt.scale = h->dt_meta[3]; c.dt.inv_scale = 1.0 / h->dt_meta[3];
    c.data = h->d_data.p; c.nd = h->nd; c.inlier_num = h->inlier_num; c.do_trim = h->p.do_trim; c.sse_thresh = h->sse_thresh;
    c.tx = h->p.trans_cube[0]; c.ty = h->p.trans_cube[1]; c.tz = h->p.trans_cube[2]; c.tw = h->p.trans_cube[3];
    for (int i = 0; i < kMaxRotLevel; i++) c.cgamma[i] = h->cgamma[i];
    c.trans_cutoff_level = kMaxTransLevel;
    c.dbg = nullptr;
    return GOICP_OK;
}

// Shared-memory plan of the persistent inner-BnB kernel: rotated points on chip when two CTAs
// per SM still fit, the rest of the per-CTA budget goes to the priority queue.
struct InnerPlan { bool pts_smem; bool keys_smem; int heap_cap_sm; int cluster; };
InnerPlan plan_inner(const goicp_handle* h, int ctas_per_sm = 2)
{
    InnerPlan p;
    // Cluster size: one expansion step costs 8*Nd scattered lookups and an SM retires ~1 per clock,
    // so split each inner BnB over as many SMs as keeps a slice worth a CTA (>= ~256 points).
    int cl = h->p.cluster_size;
    // (measured on the bunny config, Nd = 3019: clusters of 4 beat 8 and 16 -- beyond ~750 points per CTA the
    // extra barrier/DSMEM traffic costs more than the shorter gather saves)
    if (cl <= 0) { cl = 1; while (cl < 16 && h->nd / (cl * 2) >= 512) cl *= 2; }
    if (cl > 16) cl = 16;
    p.cluster = cl;
    // two CTAs per SM: each may take half of the SM's shared memory minus its static part and the 1 KB the driver reserves per CTA.
    // Signed arithmetic: what does not fit next to a 16 KB queue stays in global memory (the trimming keys first claim the
    // space -- the radix select reads them five times per expansion --, then the rotated points).
    const long stat = (long)h->max_smem_optin - h->inner_dyn_smem;
    const long per_cta = (long)h->max_smem_optin / ctas_per_sm - stat - 2048;
    const long per = (h->nd + cl - 1) / cl;
    const long pts = per * (long)sizeof(float4);
    const bool trim = h->initialized && h->inlier_num < h->nd;
    const long keys = trim ? per * 8 * (long)sizeof(unsigned) : 0;          // residual keys of the radix select
    const long min_heap = 16 * 1024;
    p.keys_smem = trim && keys + min_heap <= per_cta;
    long left = per_cta - (p.keys_smem ? keys : 0);
    p.pts_smem = pts + min_heap <= left;
    if (p.pts_smem) left -= pts;
    p.heap_cap_sm = (int)std::max<long>(64, std::min<long>(left / (long)sizeof(HeapEntry), 4096));
    return p;
}

int ensure_task_buffers(goicp_handle* h, size_t n)
{
    CUDA_TRY(h, h->d_tasks.reserve(n + 1));
    CUDA_TRY(h, h->d_results.reserve(n));
    CUDA_TRY(h, h->d_cands.reserve(n + 1));
    if (h->h_results_n < n) {
        if (h->h_results) pool_free_host(h->h_results);
        if (h->h_tasks) pool_free_host(h->h_tasks);
        h->h_results = nullptr; h->h_tasks = nullptr;
        size_t cap = std::max<size_t>(n, 1024);
        CUDA_TRY(h, pool_alloc_host((void**)&h->h_results, cap * sizeof(InnerResult)));
        CUDA_TRY(h, pool_alloc_host((void**)&h->h_tasks, cap * sizeof(InnerTask)));
        h->h_results_n = cap; h->h_tasks_n = cap;
    }
    // spill region: one slab per concurrently launched CTA (grid == n)
    const int spill_cap = 1 << 16;        // 65536 entries (1 MiB) per task beyond the ~4000 kept in shared memory; overflow is reported, not hidden
    if ((size_t)h->spill_slots < n || h->spill_cap != spill_cap) {
        size_t slots = std::max<size_t>(n, 640);
        CUDA_TRY(h, h->d_spill.reserve(slots * (size_t)spill_cap));
        h->spill_slots = (int)slots; h->spill_cap = spill_cap;
    }
    return GOICP_OK;
}

// Round-robin sharding of a round's tasks: rank r owns tasks r, r+W, r+2W, ...  Every rank sends
// its results (padded to ceil(n/W) records); after the all-gather task t is found at
// recv[(t % W) * per_rank + t / W].
int shard_exchange(goicp_allgather_fn fn, void* user, int W, int n, const InnerResult* mine, int n_mine, InnerResult* all)
{
    const int per_rank = (n + W - 1) / W;
    std::vector<InnerResult> send(per_rank), recv((size_t)per_rank * W);
    std::memset(send.data(), 0, sizeof(InnerResult) * per_rank);
    if (n_mine > 0) std::memcpy(send.data(), mine, sizeof(InnerResult) * n_mine);
    if (fn(user, send.data(), recv.data(), sizeof(InnerResult) * per_rank, 0) != 0) return 1;
    for (int t = 0; t < n; t++) all[t] = recv[(size_t)(t % W) * per_rank + t / W];
    return 0;
}

// Shape of the inner-BnB kernel for a batch (1 = low latency, 0 = 512 threads x 2 CTAs per SM, 2 = dense 192 x 5) from the
// forecast cost of its longest task and of all this rank's tasks (pops) and the previous batch's longest task; `slots` =
// clusters one GPU holds in the low-latency shape.  A pure function: tests/test_host_io.py scores it on the committed
// per-round measurements (profiles/r2x_rounds/) through goicp_bnb_shape_rule.
int bnb_shape_rule(double pred_max, double pred_sum_rank, double prev_max_pops, double slots)
{
    const double pm = std::max(pred_max, 0.5 * prev_max_pops);
    const double t_lat = std::max(pm, pred_sum_rank / slots), t_thr = std::max(1.4 * pm, 1.45 * pred_sum_rank / (2 * slots)),
                 t_q = std::max(2.0 * pm, 2.1 * pred_sum_rank / (5 * slots));
    if (t_lat <= t_thr && t_lat <= t_q) return 1;
    return t_q < t_thr ? 2 : 0;
}

// Runs `n` inner BnBs (tasks in h->h_tasks) and leaves the results in h->h_results.
// With an exchange hook installed, rank r runs tasks r, r+W, r+2W, ... and the per-rank result
// blocks are all-gathered, so every rank ends up with all n results.
int run_inner_batch(goicp_handle* h, const BnbConst& c, int n, int64_t* executed_evals, std::vector<std::shared_ptr<CandList>>* lists = nullptr, int64_t* executed_local = nullptr)
{
    if (n <= 0) return GOICP_OK;
    int rc = ensure_task_buffers(h, (size_t)n + 64); if (rc) return rc;
    const bool trim_run = c.inlier_num < c.nd;
    {
        // Kernel variant of this batch.  Three shapes of the same kernel (same sums in the same order, so the choice never
        // changes a result): low latency (512 threads, 1 CTA per SM, 128 registers), 512 x 2 and the dense one, 192 x 5.  Per
        // task the denser shapes are slower (measured on the bunny and spanner configs: 1.4x / 2.0x) and their summed task
        // cycles larger (1.45x / 2.1x) but 2 / 5 clusters share an SM: a batch costs about max(longest task, sum / resident
        // clusters) in each.  The expected cost of a task is what the same pass cost on the parent cube (pops; the caller
        // leaves it in pred_max / pred_sum); without a forecast the previous batch's measured cycles decide (below).
        static const char* force = getenv("GOICP_BNB_VARIANT");       // "lat" / "thr" / "q5" / "q3": pin the variant (experiments)
        if (force) { h->low_latency = force[0] == 'l'; h->dense = force[0] == 'q' ? (force[1] == '3' ? 3 : 2) : 0; }
        else if (h->pred_sum > 0 && !trim_run && plan_inner(h, 5).pts_smem && !(getenv("GOICP_BNB_FORECAST") && atoi(getenv("GOICP_BNB_FORECAST")) == 0)) {
            // (clouds whose rotated points would not fit a fifth of an SM's shared memory keep the cycles-only rule and the two 512-thread shapes: measured on
            // the 1e5-point sweep case, 512^3 grid in HBM, the dense shape loses -- BnB kernels 0.62 -> 0.77 s)
            const int Wq = ((h->xchg || h->nccl) && h->p.world_size > 1) ? h->p.world_size : 1;
            const int clq = plan_inner(h).cluster;
            const double slots = std::max(1, h->sm_count / clq), sum = h->pred_sum / Wq;
            // the longest task is the hard one to foresee (a child's search can be ten times its parent's): the forecast is
            // floored by half the previous batch's longest (fitted on per-round times of six golden runs in all three shapes,
            // profiles/r2x_shape_rule_fit.txt: 3 % above the per-round oracle on average, the cycles-only rule 10 %)
            const int shape = bnb_shape_rule(h->pred_max, sum, h->prev_max_pops, slots);
            h->low_latency = shape == 1;
            h->dense = shape == 2 ? 2 : 0;
        }
    }
    const int variant = h->low_latency || trim_run ? (h->low_latency ? 1 : 0) : h->dense;
    const double pred_max_used = h->pred_max, pred_sum_used = h->pred_sum;
    const InnerPlan plan = plan_inner(h, variant == 2 ? 5 : (variant == 3 ? 8 : 2));
    const int W = ((h->xchg || h->nccl) && h->p.world_size > 1) ? h->p.world_size : 1;
    const int r = W > 1 ? h->p.rank : 0;
    const int per_rank = (n + W - 1) / W;
    int mine = 0;
    std::vector<InnerTask> local;
    const InnerTask* src = h->h_tasks;
    if (W > 1) {
        local.reserve(per_rank);
        for (int t = r; t < n; t += W) local.push_back(h->h_tasks[t]);
        mine = (int)local.size();
        src = local.data();
    } else mine = n;
    const bool stats = getenv("GOICP_ROUND_STATS") != nullptr;
    BnbConst cdbg = c;
    if (stats) { CUDA_TRY(h, h->d_dbg.reserve((size_t)12 * (n + 64))); CUDA_TRY(h, cudaMemsetAsync(h->d_dbg.p, 0, sizeof(unsigned long long) * 12 * n, h->stream)); cdbg.dbg = h->d_dbg.p; }
    CUDA_TRY(h, cudaEventRecord(h->ev0, h->stream));
    if (mine > 0) {
        CUDA_TRY(h, xfer(h, h->d_tasks.p, src, sizeof(InnerTask) * mine, cudaMemcpyHostToDevice, h->stream));
        const bool trim = c.inlier_num < c.nd;
        {
            static const bool legacy = getenv("GOICP_NO_PIPELINE") != nullptr;
            h->bnb_variants |= (trim || legacy) ? 16 : (variant >= 2 ? (plan.pts_smem ? 32 : 64) : (plan.pts_smem ? (variant == 1 ? 1 : 2) : (variant == 1 ? 4 : 8)));
        }
        if (trim && !plan.keys_smem) {
            // trimming a cloud whose keys do not fit in shared memory: per-CTA key slabs in global memory, the round launched
            // in chunks so that the slabs stay below 1 GiB
            const size_t per_task = (size_t)plan.cluster * ((h->nd + plan.cluster - 1) / plan.cluster) * 8;
            const int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)mine, ((size_t)1 << 28) / per_task));
            CUDA_TRY(h, h->d_trim_keys.reserve(per_task * chunk));
            for (int off = 0; off < mine; off += chunk) {
                CUDA_TRY(h, launch_inner_bnb(cdbg, h->d_tasks.p + off, h->d_results.p + off, std::min(chunk, mine - off), plan.cluster, plan.pts_smem, plan.heap_cap_sm,
                                             h->d_spill.p, h->spill_cap, h->d_cands.p + off, variant, h->d_trim_keys.p, h->stream));
                h->launches++;
            }
        } else {
            CUDA_TRY(h, launch_inner_bnb(cdbg, h->d_tasks.p, h->d_results.p, mine, plan.cluster, plan.pts_smem, plan.heap_cap_sm, h->d_spill.p, h->spill_cap, h->d_cands.p, variant, nullptr, h->stream));
            h->launches++;
        }
    }
    if (W == 1) {
        CUDA_TRY(h, xfer(h, h->h_results, h->d_results.p, sizeof(InnerResult) * n, cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(h, cudaEventRecord(h->ev1, h->stream));
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    } else if (h->nccl) {
        // native exchange: the per-rank result blocks (padded to per_rank records) are all-gathered by NCCL on the
        // engine's stream right behind the kernel; one D2H copy brings every rank's results back.  This all-gather of
        // 48-byte records is the round's best-bound exchange: every rank commits the same results in the same order.
        CUDA_TRY(h, h->d_gather.reserve((size_t)per_rank * W));
        CUDA_TRY(h, h->d_results.reserve((size_t)per_rank));
        if (mine < per_rank) CUDA_TRY(h, cudaMemsetAsync(h->d_results.p + mine, 0, sizeof(InnerResult) * (per_rank - mine), h->stream));
        const int nrc = nccl_api()->AllGather(h->d_results.p, h->d_gather.p, sizeof(InnerResult) * per_rank, kNcclUint8, h->nccl, h->stream);
        if (nrc != 0) return fail(h, GOICP_ERR_CUDA, std::string("ncclAllGather: ") + (nccl_api()->GetErrorString ? nccl_api()->GetErrorString(nrc) : "error"));
        std::vector<InnerResult> recv((size_t)per_rank * W);
        CUDA_TRY(h, xfer(h, recv.data(), h->d_gather.p, sizeof(InnerResult) * per_rank * W, cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(h, cudaEventRecord(h->ev1, h->stream));
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
        for (int t = 0; t < n; t++) h->h_results[t] = recv[(size_t)(t % W) * per_rank + t / W];
    } else {
        std::vector<InnerResult> mine_res(std::max(mine, 1));
        if (mine > 0) CUDA_TRY(h, xfer(h, mine_res.data(), h->d_results.p, sizeof(InnerResult) * mine, cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(h, cudaEventRecord(h->ev1, h->stream));
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
        if (shard_exchange(h->xchg, h->xchg_user, W, n, mine_res.data(), mine, h->h_results) != 0)
            return fail(h, GOICP_ERR_INVALID, "exchange callback failed");
    }
    float ms = 0; cudaEventElapsedTime(&ms, h->ev0, h->ev1); h->t_kernels += ms * 1e-3;
    if (getenv("GOICP_ROUND_STATS")) {
        uint32_t maxc = 0, maxp = 0, sump = 0, maxh = 0; uint64_t sumc = 0; int arg = 0, flagged = 0;
        for (int t = 0; t < n; t++) { const InnerResult& q = h->h_results[t]; sumc += q.kcycles; sump += q.pops; if (q.kcycles > maxc) { maxc = q.kcycles; arg = t; } if (q.pops > maxp) maxp = q.pops; if (q.max_heap > maxh) maxh = q.max_heap; if (q.pad[1]) flagged++; }
        fprintf(stderr, "[round] shape %d forecast max %.0f sum %.0f; ", variant, pred_max_used, pred_sum_used);
        fprintf(stderr, "[round] tasks %d kernel %.3f ms; slowest task %.3f Mcyc (pops %u, level %d); max pops %u; total pops %u; sum task cycles %.1f Mcyc; max heap %u; tasks with contenders %d\n",
                n, ms, maxc * 1024e-6, h->h_results[arg].pops, h->h_tasks[arg].level, maxp, sump, sumc * 1024e-6, maxh, flagged);
        if (W == 1) {
            unsigned long long d[12];
            cudaMemcpy(d, h->d_dbg.p + (size_t)12 * arg, sizeof d, cudaMemcpyDeviceToHost);
            const double steps = std::max(1u, h->h_results[arg].evals / 8);
            fprintf(stderr, "        slowest task, cycles per expansion (cluster %d): owner maint %.0f (arrive %.0f, pushes %.0f = %.1f entries, pop %.0f) waitA %.0f waitB %.0f bookkeeping %.0f | gather warp waitA %.0f gather %.0f reduce %.0f waitB %.0f\n",
                    plan.cluster, d[0] / steps, d[8] / steps, d[9] / steps, d[10] / steps, d[11] / steps, d[1] / steps, d[2] / steps, d[3] / steps, d[4] / steps, d[5] / steps, d[6] / steps, d[7] / steps);
        }
    }
    {
        // Variant for the next round, from this round's task cycle counts (all ranks' tasks).  In low-latency units
        // (measured ratios on the bunny config: a task is 1.43x slower, the sum of task cycles 1.54x larger in the
        // two-CTAs-per-SM variant) a round costs about max(longest task, sum / resident clusters) in either variant.
        double maxc = 0, sumc = 0;
        for (int t = 0; t < n; t++) { const double k = h->h_results[t].kcycles; sumc += k; if (k > maxc) maxc = k; }
        if (variant == 0) { maxc /= 1.43; sumc /= 1.54; }
        else if (variant >= 2) { maxc /= 2.0; sumc /= 2.1; }
        const double per_rank_sum = sumc / W;
        const double res_lat = std::max(1, h->sm_count / plan.cluster), res_thr = std::max(1, 2 * h->sm_count / plan.cluster);
        const double t_lat = std::max(maxc, per_rank_sum / res_lat), t_thr = std::max(1.43 * maxc, 1.54 * per_rank_sum / res_thr);
        h->low_latency = t_lat <= t_thr; h->dense = 0;
    }
    h->pred_max = h->pred_sum = 0;                                   // a forecast is for one batch
    { double mp = 0; for (int t = 0; t < n; t++) mp = std::max(mp, (double)h->h_results[t].pops); h->prev_max_pops = mp; }
    if (lists) {
        // contender lists of the upper-bound passes that have any (local tasks only; a rank that
        // needs a list it does not hold re-runs that one task, see commit)
        lists->assign(n, nullptr);
        for (int k = 0; k < mine; k++) {
            const int t = W > 1 ? r + k * W : k;
            if (h->h_tasks[t].level < 0 && h->h_results[t].pad[1] > 0) {
                auto cl = std::make_shared<CandList>();
                CUDA_TRY(h, xfer(h, cl.get(), h->d_cands.p + k, sizeof(CandList), cudaMemcpyDeviceToHost, h->stream));
                (*lists)[t] = cl;
            }
        }
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    }
    for (int t = 0; t < n; t++) {
        if (h->h_results[t].status == 3 && (h->h_results[t].pad[0] & 0x100u)) return fail(h, GOICP_ERR_CUDA, "internal: the pipelined translation search mispredicted a queue pop (please report)");
        if (h->h_results[t].status == 3) return fail(h, GOICP_ERR_CAPACITY, "translation priority queue overflowed its device capacity");
        if (h->h_results[t].status == 4) return fail(h, GOICP_ERR_DEPTH, "translation search exceeded 21 levels");
        if (executed_evals) *executed_evals += h->h_results[t].evals;
        if (executed_local && t % W == r) *executed_local += h->h_results[t].evals;
    }
    return GOICP_OK;
}

// All ranks hold the same result records but only the rank that ran an upper-bound pass holds its contender list.
// When the commit needs that list (same moment on every rank: the decision depends on replicated data only), the
// holder's copy is handed round with one small all-gather over the handle's exchange -- instead of every other rank
// re-running the whole inner BnB just to re-derive it (as long as the round's longest task at Nd = 1e5).
int share_cand_list(goicp_handle* h, std::shared_ptr<CandList>& cl)
{
    const int W = h->p.world_size;
    struct Block { int32_t have; int32_t pad; CandList list; };
    std::vector<Block> recv((size_t)W);
    Block mine; std::memset(&mine, 0, sizeof mine);
    if (cl) { mine.have = 1; mine.list = *cl; }
    if (h->nccl) {
        CUDA_TRY(h, h->d_share.reserve((size_t)(W + 1) * sizeof(Block)));
        CUDA_TRY(h, xfer(h, h->d_share.p, &mine, sizeof mine, cudaMemcpyHostToDevice, h->stream));
        const int nrc = nccl_api()->AllGather(h->d_share.p, h->d_share.p + sizeof(Block), sizeof(Block), kNcclUint8, h->nccl, h->stream);
        if (nrc != 0) return fail(h, GOICP_ERR_CUDA, "ncclAllGather (contender list) failed");
        CUDA_TRY(h, xfer(h, recv.data(), h->d_share.p + sizeof(Block), sizeof(Block) * W, cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    } else if (h->xchg) {
        if (h->xchg(h->xchg_user, &mine, recv.data(), sizeof(Block), 0) != 0) return fail(h, GOICP_ERR_INVALID, "exchange callback failed");
    } else return GOICP_OK;
    if (!cl)
        for (int r = 0; r < W; r++)
            if (recv[r].have) { cl = std::make_shared<CandList>(recv[r].list); break; }
    return GOICP_OK;
}

// Settles value and arg-min cube of one upper-bound pass in the reference's summation order
// (strict_sum.cuh).  `list` are the contenders the search kernel reported for that task.
int resolve_strict(goicp_handle* h, const BnbConst& c, const InnerTask& task, const CandList& list, float* value, float* node4)
{
    struct Acc { double& a; double t0; ~Acc() { a += now_s() - t0; } } acc{h->t_strict, now_s()};
    const size_t scratch = strict_smem_mode(h->nd, h->max_smem_optin - 2048) == 2 ? 0 : (size_t)kMaxCand * 3 * h->nd;      // residuals + the select's position lists
    CUDA_TRY(h, h->d_strict.reserve(256 + scratch));
    CUDA_TRY(h, h->d_tasks.reserve(1)); CUDA_TRY(h, h->d_cands.reserve(1));
    InnerTask* d_task = h->d_tasks.p + (h->d_tasks.n - 1);          // last slots are reserved for this
    CandList* d_list = h->d_cands.p + (h->d_cands.n - 1);
    CUDA_TRY(h, xfer(h, d_task, &task, sizeof task, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, xfer(h, d_list, &list, sizeof list, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, launch_strict_resolve(c, d_task, d_list, h->d_strict.p, h->d_strict.p + 256, h->d_strict.p + 128, h->max_smem_optin - 2048, h->stream));
    h->launches += 2; h->strict_resolves++;
    float out5[5];
    CUDA_TRY(h, xfer(h, out5, h->d_strict.p + 128, sizeof out5, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    *value = out5[0]; node4[0] = out5[1]; node4[1] = out5[2]; node4[2] = out5[3]; node4[3] = out5[4];
    return GOICP_OK;
}

int score_pose(goicp_handle* h, const BnbConst& c, const float* R, const float* t, float* out)
{
    const double t_begin = now_s();
    struct Acc { double& a; double t0; ~Acc() { a += now_s() - t0; } } acc{h->t_score, t_begin};
    float Rt[12]; int use = R ? 1 : 0;
    for (int i = 0; i < 9; i++) Rt[i] = R ? R[i] : 0.0f;
    for (int i = 0; i < 3; i++) Rt[9 + i] = R ? t[i] : 0.0f;
    CUDA_TRY(h, h->d_f32a.reserve(64));
    CUDA_TRY(h, h->d_i32.reserve(16));
    CUDA_TRY(h, xfer(h, h->d_f32a.p, Rt, sizeof Rt, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, xfer(h, h->d_i32.p, &use, sizeof use, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, h->d_score_scratch.reserve((size_t)3 * h->nd));                     // residuals + the select's position lists (strict_smem_mode < 2)
    CUDA_TRY(h, launch_dt_score(c, h->d_f32a.p, h->d_i32.p, 1, h->d_score_scratch.p, h->d_f32a.p + 16, h->max_smem_optin - 2048, (h->p.numerics & GOICP_NUM_FAST_SUMS) != 0, h->stream));
    h->launches++;
    CUDA_TRY(h, xfer(h, out, h->d_f32a.p + 16, sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return GOICP_OK;
}

int run_icp(goicp_handle* h, const float* R0, const float* t0, int max_iter, float err_diff, goicp_icp_result* out, int blocks_cap = 1 << 30, bool force_fast = false)
{
    int rc = ensure_cuda(h); if (rc) return rc;
    rc = upload_data(h); if (rc) return rc;
    rc = ensure_kdtree(h); if (rc) return rc;
    const int num = h->p.do_trim ? (int)((float)h->nd * (1 - h->p.trim_fraction)) : h->nd;     // jly_icp3d.hpp:189-196
    size_t npad = 1; while (npad < (size_t)h->nd) npad <<= 1;
    CUDA_TRY(h, h->d_icp_q.reserve((size_t)8 * h->nd)); CUDA_TRY(h, h->d_icp_d2.reserve(h->nd)); CUDA_TRY(h, h->d_icp_nn.reserve(h->nd)); CUDA_TRY(h, h->d_icp_pos.reserve(h->nd));
    CUDA_TRY(h, h->d_icp_keys.reserve(npad)); CUDA_TRY(h, h->d_icp_stage.reserve((size_t)8 * h->nd)); CUDA_TRY(h, h->d_icp_order.reserve((size_t)h->nd + 2048));     // + slack: doubles as the per-CTA lists of deferred queries
    // radix-sort scratch of the large-cloud path: second key buffer, digit-major histogram (256 x warps of the grid) + 1024 chunk totals
    const size_t hist_n = (size_t)256 * icp_max_blocks_supported() * (icp_threads() / 32);
    CUDA_TRY(h, h->d_icp_keys2.reserve((size_t)h->nd + 2048));     // + slack: doubles as the per-CTA lists of unsettled queries
    CUDA_TRY(h, h->d_icp_hist.reserve(hist_n + 1024));
    IcpWork wk; wk.keys2 = h->d_icp_keys2.p; wk.hist = h->d_icp_hist.p; wk.blocksum = h->d_icp_hist.p + hist_n; wk.q = h->d_icp_q.p; wk.nn = h->d_icp_nn.p; wk.pos = h->d_icp_pos.p; wk.d2 = h->d_icp_d2.p; wk.keys = h->d_icp_keys.p; wk.stage = h->d_icp_stage.p; wk.order = h->d_icp_order.p;
    IcpState st; std::memset(&st, 0, sizeof st);
    for (int i = 0; i < 9; i++) st.R[i] = R0[i];
    for (int i = 0; i < 3; i++) st.t[i] = t0[i];
    st.err = -1.0f;
    const double t_begin = now_s();
    CUDA_TRY(h, xfer(h, h->d_icp_state.p, &st, sizeof st, cudaMemcpyHostToDevice, h->stream));
    const int n_nodes = (int)h->kd_host.nodes.size();
    const bool fast = force_fast || (h->p.numerics & GOICP_NUM_FAST_ICP) != 0;
    const int max_blocks = icp_max_grid_blocks(h->p.device, kd_view(h), n_nodes, h->nd, num, h->max_smem_optin, fast);
    if (max_blocks <= 0) return fail(h, GOICP_ERR_CUDA, "cooperative ICP kernel cannot be resident");
    // queries are interleaved over the CTAs, one per warp and pass (16 per CTA): use every SM the cooperative launch allows
    int blocks = std::max(1, std::min(std::min(std::min(max_blocks, icp_max_blocks_supported()), (h->nd + 15) / 16), blocks_cap));
    if (const char* e = getenv("GOICP_ICP_BLOCKS")) blocks = std::max(1, std::min(std::min(max_blocks, icp_max_blocks_supported()), std::min(atoi(e), blocks_cap)));   // experiments
    if (fast) CUDA_TRY(h, h->d_icp_partials.reserve((size_t)2 * 16 * blocks));
    // Multi-GPU, tolerance numerics, large cloud: the nearest-neighbour search -- all that is left of an iteration once the
    // sums are parallel -- is dealt over the ranks by query range.  Per iteration every rank reduces its queries to 16
    // moments, the W x 128 bytes are all-gathered (NCCL on the engine stream, or the host hook) and every rank solves for
    // the same new pose; the host loop reads the convergence flag back each iteration (~50 us, against milliseconds of
    // search per iteration at these sizes).  Strict ICP is not sharded: its sorted sequential sums need every row on one GPU.
    const int W = ((h->xchg || h->nccl) && h->p.world_size > 1) ? h->p.world_size : 1;
    static const int shard_min = getenv("GOICP_ICP_SHARD_MIN") ? atoi(getenv("GOICP_ICP_SHARD_MIN")) : 20000;
    if (fast && W > 1 && h->nd >= shard_min && !(h->p.do_trim && num < h->nd)) {
        const int per = (h->nd + W - 1) / W, q0 = std::min(h->p.rank * per, h->nd), q1 = std::min(q0 + per, h->nd);
        CUDA_TRY(h, h->d_icp_xch.reserve((size_t)16 * W));
        std::vector<double> hx((size_t)16 * W);
        int it = 0;
        for (; it < max_iter; it++) {
            CUDA_TRY(h, launch_icp_fast_shard(kd_view(h), n_nodes, h->d_data.p, h->nd, h->d_icp_state.p, wk, num, blocks, h->max_smem_optin, h->d_icp_partials.p, q0, q1,
                                              h->d_icp_xch.p + (size_t)16 * h->p.rank, h->stream));
            if (h->nccl) {
                const int nrc = nccl_api()->AllGather(h->d_icp_xch.p + (size_t)16 * h->p.rank, h->d_icp_xch.p, 16 * sizeof(double), kNcclUint8, h->nccl, h->stream);
                if (nrc != 0) return fail(h, GOICP_ERR_CUDA, "ncclAllGather (ICP moments) failed");
            } else {
                CUDA_TRY(h, xfer(h, hx.data() + (size_t)16 * h->p.rank, h->d_icp_xch.p + (size_t)16 * h->p.rank, 16 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
                CUDA_TRY(h, cudaStreamSynchronize(h->stream));
                std::vector<double> mine(hx.begin() + 16 * h->p.rank, hx.begin() + 16 * (h->p.rank + 1));
                if (h->xchg(h->xchg_user, mine.data(), hx.data(), 16 * sizeof(double), 0) != 0) return fail(h, GOICP_ERR_INVALID, "exchange callback failed");
                CUDA_TRY(h, xfer(h, h->d_icp_xch.p, hx.data(), sizeof(double) * 16 * W, cudaMemcpyHostToDevice, h->stream));
            }
            CUDA_TRY(h, launch_icp_fast_solve(h->d_icp_state.p, h->d_icp_xch.p, W, h->nd, num, err_diff, it, max_iter, h->stream));
            h->launches += 2;
            CUDA_TRY(h, xfer(h, &st, h->d_icp_state.p, sizeof st, cudaMemcpyDeviceToHost, h->stream));
            CUDA_TRY(h, cudaStreamSynchronize(h->stream));
            if (st.converged) break;
        }
        h->t_icp += now_s() - t_begin;
        for (int i = 0; i < 9; i++) out->R[i] = st.R[i];
        for (int i = 0; i < 3; i++) out->t[i] = st.t[i];
        out->err = st.err_new; out->iterations = st.iter;
        return GOICP_OK;
    }
    CUDA_TRY(h, launch_icp(kd_view(h), n_nodes, h->d_data.p, h->nd, h->d_icp_state.p, wk, max_iter, err_diff, num, (h->p.do_trim ? 1 : 0) | ((h->p.numerics & GOICP_NUM_JACOBI_SVD) ? 2 : 0), blocks, h->max_smem_optin, fast, h->d_icp_partials.p, h->stream));
    h->launches++;
    CUDA_TRY(h, xfer(h, &st, h->d_icp_state.p, sizeof st, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->t_icp += now_s() - t_begin;
    for (int i = 0; i < 9; i++) out->R[i] = st.R[i];
    for (int i = 0; i < 3; i++) out->t[i] = st.t[i];
    out->err = st.err_new; out->iterations = st.iter;
    if (getenv("GOICP_ICP_STATS"))
        fprintf(stderr, "[icp] iters %d blocks %d cycles: nn %lld wait %lld sort %lld pass1 %lld pass2 %lld acc1 %lld svd+compose %lld\n", st.iter, blocks,
                st.dbg[0] & 0xffffffffll, st.dbg[1], st.dbg[2], st.dbg[3], st.dbg[4], st.dbg[5], st.dbg[0] >> 32);
    return GOICP_OK;
}

// GoICP::ICP (jly_goicp.cpp:93-132)
int icp_then_dt(goicp_handle* h, const BnbConst& c, float* R, float* t, float* err, const goicp_icp_result* done = nullptr)
{
    goicp_icp_result r;
    if (done) r = *done;
    else {
        int rc = run_icp(h, R, t, h->p.icp_max_iter, h->p.mse_threshold / 10000, &r);  // err_diff_def = MSEThresh/10000 (jly_goicp.cpp:186)
        if (rc) return rc;
    }
    std::memcpy(R, r.R, sizeof r.R); std::memcpy(t, r.t, sizeof r.t);
    return score_pose(h, c, R, t, err);
}

void publish(goicp_handle* h, const goicp_result& res, int finished)
{
    std::lock_guard<std::mutex> lk(h->snap_mtx);
    std::memcpy(h->snap.R, res.R, sizeof res.R); std::memcpy(h->snap.t, res.t, sizeof res.t);
    h->snap.sse = res.sse; h->snap.rot_pops = res.rot_pops; h->snap.trans_pops = res.trans_pops;
    h->snap.bound_evals = res.bound_evals; h->snap.finished = finished;
}


// ---------------------------------------------------------------------------------------------
// GOICP_SEARCH_FGOICP: the search strategy of the reference's GPU path, icp::FastGoICP::run (src/fgoicp/fgoicp.cpp:9-181),
// on this engine's kernels.  Rotation nodes are cubes (centre q, half edge `span`) of the vector part of a unit quaternion
// in [-1,1]^3 (fgoicp_common.hpp:31-105); the translation domain is [-1,1]^3; a popped rotation node spawns 8 children
// unless their span falls below 0.1 (:50); a child outside the unit ball is dropped, one whose centre is outside it is
// queued unevaluated with its parent's bounds (:58-63); otherwise an upper-bound translation BnB (rotation fixed), the
// relaxed refinement trigger `ub < 2 * best => ICP` (:75-91, 500 iterations, relative stop 1e-3), then a lower-bound
// translation BnB with the rotation-uncertainty radius of the cube; children with lb >= best are dropped (:97).  The search
// ends on `best - lb <= sse_threshold` (:44) or when the queue runs dry.  Nodes order by lb, ties by larger span (:88-95).
// The translation BnB is this library's persistent inner-BnB kernel with fgoicp's span cut-off (translation cubes of half
// edge < 0.12 are evaluated but not subdivided, :156-157) -- up to 8 queue nodes' children (128 inner BnBs) per device round,
// committed in queue order.  Deviations, all deliberate: distances come from the distance transform (fgoicp's table only
// covers [0,1]^3); the rotation-uncertainty radius is 2 |p| sin(min(sqrt3 pi span / 2, pi / 2)) (registration.cu:41-43 uses
// |p|^2 with a "needs examination" note); the error of a pose is ICP's sum of squared nearest-neighbour distances, as there.
// ---------------------------------------------------------------------------------------------
struct QNode { float x, y, z, span, lb, ub; };
struct QLower { bool operator()(const QNode& a, const QNode& b) const { return a.lb == b.lb ? a.span < b.span : a.lb > b.lb; } };

void quat_rotation(float x, float y, float z, float* R)
{
    const float r = x * x + y * y + z * z, ww = 1.0f - r, w = std::sqrt(std::max(ww, 0.0f));
    const float wx = w * x, xx = x * x, wy = w * y, xy = x * y, yy = y * y, wz = w * z, xz = x * z, yz = y * z, zz = z * z;
    R[0] = ww + xx - yy - zz; R[1] = 2 * (xy - wz);      R[2] = 2 * (xz + wy);
    R[3] = 2 * (xy + wz);     R[4] = ww - xx + yy - zz;  R[5] = 2 * (yz - wx);
    R[6] = 2 * (xz - wy);     R[7] = 2 * (yz + wx);      R[8] = ww - xx - yy + zz;
}

int register_fgoicp(goicp_handle* h, goicp_result* out)
{
    const double t_begin = now_s();
    h->initialized = false;
    int rc = initialize(h); if (rc) return rc;
    BnbConst c; rc = make_const(h, c); if (rc) return rc;
    rc = ensure_kdtree(h); if (rc) return rc;
    h->t_setup = now_s() - t_begin;
    // fgoicp's constants
    const float kRotSpanMin = 0.1f, kIcpRel = 1e-3f;
    const int kIcpIter0 = 1000, kIcpIter = 500, kBatch = 8;
    c.tx = c.ty = c.tz = -1.0f; c.tw = 2.0f;                                   // TransNode(0,0,0, span 1)
    c.trans_cutoff_level = 3;                                                  // half edges 1, .5, .25, .125 are subdivided; .0625 < 0.12 is not
    c.inlier_num = h->nd; c.do_trim = 0;
    const float sse_threshold = (float)h->nd * h->p.mse_threshold;             // fgoicp.hpp:24
    c.sse_thresh = sse_threshold;
    for (int d = 0; d < kMaxRotLevel; d++) {                                   // rotation-uncertainty scale of a cube at depth d (half edge 2^-d)
        const double half_angle = std::min(kSqrt3 * kPi / 2.0 * std::pow(0.5, d), kPi / 2.0);
        c.cgamma[d] = (float)(2.0 * std::sin(half_angle));
    }
    goicp_result res; std::memset(&res, 0, sizeof res);
    res.sse_thresh = sse_threshold;
    float best = 1e+10f, optR[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, optT[3] = {0, 0, 0};
    auto sync_res = [&]() { std::memcpy(res.R, optR, sizeof optR); std::memcpy(res.t, optT, sizeof optT); res.sse = best; };
    {
        goicp_icp_result r;
        rc = run_icp(h, optR, optT, kIcpIter0, -kIcpRel, &r, 1 << 30, true); if (rc) return rc;
        res.icp_calls++;
        best = r.err; std::memcpy(optR, r.R, sizeof optR); std::memcpy(optT, r.t, sizeof optT);
    }
    sync_res(); publish(h, res, 0);
    int exit_path = GOICP_EXIT_NONE; float exit_lb = 0;
    if (!(best > sse_threshold)) exit_path = GOICP_EXIT_EARLY_SSE;             // fgoicp.cpp:21: the BnB only runs when the first ICP is not good enough
    std::vector<QNode> heap; QLower lower;
    heap.push_back(QNode{0, 0, 0, 1.0f, 0.0f, best}); std::push_heap(heap.begin(), heap.end(), lower);
    struct Child { int parent; QNode n; bool evaluate; float R[9]; int depth; int task_ub, task_lb; };
    while (exit_path == GOICP_EXIT_NONE) {
        if (h->cancel_flag.load()) { exit_path = GOICP_EXIT_CANCELLED; break; }
        if (heap.empty()) { exit_path = GOICP_EXIT_QUEUE_EMPTY; break; }
        // pop up to kBatch nodes (the first decides the certificate, :44)
        std::vector<QNode> popped;
        while (!heap.empty() && (int)popped.size() < kBatch) {
            const QNode top = heap.front();
            if (best - top.lb <= sse_threshold) { if (popped.empty()) { exit_path = GOICP_EXIT_CERTIFIED; exit_lb = top.lb; res.rot_pops++; } break; }
            std::pop_heap(heap.begin(), heap.end(), lower); heap.pop_back();
            popped.push_back(top);
        }
        if (exit_path != GOICP_EXIT_NONE) break;
        std::vector<Child> kids;
        int n_tasks = 0;
        rc = ensure_task_buffers(h, (size_t)kBatch * 16 + 64); if (rc) return rc;
        for (size_t pi = 0; pi < popped.size(); pi++) {
            const QNode& P = popped[pi];
            const float span = P.span / 2.0f;
            if (span < kRotSpanMin) continue;
            int depth = 0; { float sp = 1.0f; while (sp > span * 1.5f && depth < kMaxRotLevel - 1) { sp *= 0.5f; depth++; } }
            for (int j = 0; j < 8; j++) {
                Child k; k.parent = (int)pi; k.depth = depth; k.task_ub = k.task_lb = -1;
                k.n = QNode{P.x - span + (j & 1) * P.span, P.y - span + (j >> 1 & 1) * P.span, P.z - span + (j >> 2 & 1) * P.span, span, P.lb, P.ub};
                // nearest point of the cube to the origin inside the unit ball?  (overlaps_SO3, fgoicp_common.hpp:103-107, stated geometrically)
                const float nx = std::max(std::fabs(k.n.x) - span, 0.0f), ny = std::max(std::fabs(k.n.y) - span, 0.0f), nz = std::max(std::fabs(k.n.z) - span, 0.0f);
                if (nx * nx + ny * ny + nz * nz > 1.0f) continue;
                k.evaluate = k.n.x * k.n.x + k.n.y * k.n.y + k.n.z * k.n.z <= 1.0f;      // in_SO3
                if (k.evaluate) {
                    quat_rotation(k.n.x, k.n.y, k.n.z, k.R);
                    for (int pass = 0; pass < 2; pass++) {
                        InnerTask& t = h->h_tasks[n_tasks];
                        std::memcpy(t.R, k.R, sizeof k.R);
                        t.level = pass == 0 ? -1 : depth;
                        t.opt_error = pass == 0 ? 2.0f * best : best;           // the ub pass must report values up to 2 * best (:75)
                        t.pad = 0;
                        (pass == 0 ? k.task_ub : k.task_lb) = n_tasks++;
                    }
                }
                kids.push_back(k);
            }
        }
        res.rot_pops += (int64_t)popped.size();
        if (n_tasks > 0) { rc = run_inner_batch(h, c, n_tasks, &res.bound_evals_executed, nullptr, &res.bound_evals_executed_local); if (rc) return rc; res.rounds++; }
        const std::vector<InnerResult> results(h->h_results, h->h_results + n_tasks);
        for (const Child& k : kids) {
            QNode nd = k.n;
            if (!k.evaluate) { heap.push_back(nd); std::push_heap(heap.begin(), heap.end(), lower); continue; }
            const InnerResult& ru = results[k.task_ub]; const InnerResult& rl = results[k.task_lb];
            res.trans_pops += ru.pops + rl.pops; res.bound_evals += ru.evals + rl.evals;
            const float ub = ru.value;
            if (ub < 2.0f * best && ru.node[3] > 0.0f) {
                float t0[3] = {ru.node[0] + ru.node[3] / 2, ru.node[1] + ru.node[3] / 2, ru.node[2] + ru.node[3] / 2};
                goicp_icp_result r;
                rc = run_icp(h, k.R, t0, kIcpIter, -kIcpRel, &r, 1 << 30, true); if (rc) return rc;
                res.icp_calls++;
                if (r.err < best) { best = r.err; std::memcpy(optR, r.R, sizeof optR); std::memcpy(optT, r.t, sizeof optT); sync_res(); publish(h, res, 0); }
            }
            const float lb = rl.value;
            if (lb >= best) continue;
            nd.lb = lb; nd.ub = ub;
            heap.push_back(nd); std::push_heap(heap.begin(), heap.end(), lower);
        }
    }
    sync_res();
    res.exit_path = exit_path; res.best_lb = exit_lb; res.kernel_launches = h->launches;
    res.seconds_total = now_s() - t_begin; res.seconds_bnb_kernels = h->t_kernels; res.seconds_icp = h->t_icp; res.seconds_setup = h->t_setup;
    res.bnb_kernel_variants = h->bnb_variants;
    publish(h, res, 1);
    *out = res;
    return exit_path == GOICP_EXIT_CANCELLED ? GOICP_ERR_CANCELLED : GOICP_OK;
}

} // namespace

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

void goicp_default_params(goicp_params* p)
{
    std::memset(p, 0, sizeof *p);
    p->mse_threshold = 1e-3f;
    p->trim_fraction = 0.0f; p->do_trim = 1;
    p->dt_size = 300; p->dt_expand = 2.0;
    p->rot_cube[0] = p->rot_cube[1] = p->rot_cube[2] = (float)-kPi; p->rot_cube[3] = (float)(2 * kPi);
    p->trans_cube[0] = p->trans_cube[1] = p->trans_cube[2] = -0.5f; p->trans_cube[3] = 1.0f;
    p->icp_max_iter = 10000;
    p->device = 0; p->spec_cubes = 0; p->cluster_size = 0; p->dt_mode = GOICP_DT_EXACT_EDT_REFSEED;
    p->rank = 0; p->world_size = 1;
    p->numerics = GOICP_NUM_STRICT;
    p->search_mode = GOICP_SEARCH_GOICP;
}

int goicp_create(const goicp_params* p, goicp_handle** out)
{
    if (!p || !out) return GOICP_ERR_INVALID;
    goicp_handle* h = new goicp_handle();
    h->p = *p;
    std::memset(&h->snap, 0, sizeof h->snap);
    *out = h;
    return GOICP_OK;
}

int goicp_destroy(goicp_handle* h)
{
    if (!h) return GOICP_OK;
    if (h->cuda_ready) {
        cudaSetDevice(h->p.device);
        // nothing queued on this handle's streams may still touch a buffer when it goes back to the process-wide pool
        if (h->stream) cudaStreamSynchronize(h->stream);
        if (h->stream_dt) cudaStreamSynchronize(h->stream_dt);
        h->nccl = nullptr;                        // communicators are shared process-wide (goicp_nccl_init)
        h->d_gather.release(); h->d_share.release();
        if (h->dt_tex) { cudaDestroyTextureObject(h->dt_tex); h->dt_tex = 0; } h->dt_tex_ptr = nullptr;
        h->d_dt.release(); h->d_data.release(); h->d_kd_nodes.release(); h->d_kd_boxes.release(); h->d_kd_vind.release(); h->d_kd_leaf.release(); h->d_model.release(); h->d_grid_start.release(); h->d_grid_pts.release();
        h->d_tasks.release(); h->d_results.release(); h->d_spill.release(); h->d_cands.release(); h->d_trim_keys.release(); h->d_dbg.release(); h->d_strict.release(); h->d_pairs.release(); h->d_f32a.release(); h->d_f32b.release();
        h->d_i32.release(); h->d_q.release(); h->d_score_scratch.release(); h->d_icp_state.release(); h->d_icp_q.release(); h->d_icp_d2.release(); h->d_icp_stage.release(); h->d_icp_nn.release(); h->d_icp_pos.release(); h->d_icp_keys.release(); h->d_icp_keys2.release(); h->d_icp_hist.release(); h->d_icp_order.release(); h->d_icp_partials.release(); h->d_icp_xch.release();
        if (h->h_results) pool_free_host(h->h_results);
        if (h->h_tasks) pool_free_host(h->h_tasks);
        if (h->stream && h->stream_dt && h->ev0 && h->ev1) {       // kept for the next handle on this device (ensure_cuda)
            std::lock_guard<std::mutex> lk(g_dev_mtx);
            StreamSet ss; ss.stream = h->stream; ss.stream_dt = h->stream_dt; ss.ev0 = h->ev0; ss.ev1 = h->ev1;
            g_stream_sets[h->p.device].push_back(ss);
        } else {
            if (h->ev0) cudaEventDestroy(h->ev0);
            if (h->ev1) cudaEventDestroy(h->ev1);
            if (h->stream) cudaStreamDestroy(h->stream);
            if (h->stream_dt) cudaStreamDestroy(h->stream_dt);
        }
    }
    delete h;
    return GOICP_OK;
}

const char* goicp_last_error(const goicp_handle* h) { return h ? h->err.c_str() : "null handle"; }

int goicp_set_model(goicp_handle* h, const float* xyz, int n)
{
    if (!h || !xyz || n <= 0) return fail(h, GOICP_ERR_INVALID, "set_model: bad arguments");
    h->model.assign(xyz, xyz + 3 * (size_t)n); h->nm = n; h->kd_ready = false; h->icp0_valid = false;
    return GOICP_OK;
}
int goicp_set_data(goicp_handle* h, const float* xyz, int n)
{
    if (!h || !xyz || n <= 0) return fail(h, GOICP_ERR_INVALID, "set_data: bad arguments");
    h->data.assign(xyz, xyz + 3 * (size_t)n); h->nd = n; h->data_uploaded = false; h->initialized = false; h->icp0_valid = false;
    return GOICP_OK;
}

int goicp_set_dt(goicp_handle* h, const float* grid, int size, const double meta4[4])
{
    if (!h || !grid || size <= 1 || !meta4) return fail(h, GOICP_ERR_INVALID, "set_dt: bad arguments");
    int rc = ensure_cuda(h); if (rc) return rc;
    const size_t n3 = (size_t)size * size * size;
    CUDA_TRY(h, h->d_dt.reserve(n3));
    CUDA_TRY(h, xfer(h, h->d_dt.p, grid, n3 * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->dt_size = size; std::memcpy(h->dt_meta, meta4, sizeof h->dt_meta); h->have_dt = true;
    return GOICP_OK;
}
int goicp_get_dt(goicp_handle* h, float* grid_out, double meta4_out[4])
{
    if (!h || !h->have_dt) return fail(h, GOICP_ERR_INVALID, "get_dt: no distance transform");
    if (meta4_out) std::memcpy(meta4_out, h->dt_meta, sizeof h->dt_meta);
    if (grid_out) {
        const size_t n3 = (size_t)h->dt_size * h->dt_size * h->dt_size;
        CUDA_TRY(h, cudaMemcpy(grid_out, h->d_dt.p, n3 * sizeof(float), cudaMemcpyDeviceToHost));
        h->bytes_d2h += (int64_t)(n3 * sizeof(float));
    }
    return GOICP_OK;
}
int goicp_dt_size(const goicp_handle* h) { return h && h->have_dt ? h->dt_size : 0; }

int goicp_build_dt(goicp_handle* h)
{
    if (!h || h->nm <= 0) return fail(h, GOICP_ERR_INVALID, "build_dt: no model cloud set");
    const bool trace = getenv("GOICP_TRACE_DT") != nullptr;
    const double tr0 = now_s();
    int rc = ensure_cuda(h); if (rc) return rc;
    if (trace) fprintf(stderr, "[dt trace] ensure_cuda %.3f ms\n", 1e3 * (now_s() - tr0));
    const int S = h->p.dt_size;
    if (S < 2 || S > 1024) return fail(h, GOICP_ERR_INVALID, "build_dt: dt_size out of range [2,1024]");
    const size_t n3 = (size_t)S * S * S;
    h->have_dt = false;                              // a failed rebuild must not leave a half-written grid in use
    CUDA_TRY(h, h->d_dt.reserve(n3));
    std::string msg;
    cudaError_t e = cudaSuccess;
    h->icp0_valid = false;
    const bool no_overlap = getenv("GOICP_NO_PREFETCH") != nullptr;
    if (h->nd > 0 && !no_overlap && h->p.dt_mode == GOICP_DT_REFERENCE) {
        // GoICP::Register starts with an ICP from the identity pose (jly_goicp.cpp:378-391) that reads the clouds and
        // the kd-tree but not the DT.  The reference-order DT propagation is one CTA on one SM for most of the build
        // (DESIGN.md section 5), so that refinement runs here, on the other SMs, while the DT is being built: the DT
        // build goes to its own stream on a helper thread, this thread builds the kd-tree and drives the ICP kernel
        // (capped to the SMs the DT leaves free, so that the cooperative launch can be resident next to it).
        std::thread worker;
        bool threaded = true;
        try {
            worker = std::thread([&]() {
                cudaSetDevice(h->p.device);
                e = dt_build_device(h->model.data(), h->nm, S, h->p.dt_expand, h->p.dt_mode, h->d_dt.p, h->dt_meta, h->stream_dt, msg);
            });
        } catch (...) { threaded = false; }                 // no thread to be had: build the DT here, no head start
        if (threaded) {
            const float R0[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, t0[3] = {0, 0, 0};
            const double tr1 = now_s();
            const int rc_icp = run_icp(h, R0, t0, h->p.icp_max_iter, h->p.mse_threshold / 10000, &h->icp0, std::max(1, h->sm_count - 1));
            if (trace) fprintf(stderr, "[dt trace] overlapped ICP (upload + kd-tree + kernel) %.3f ms\n", 1e3 * (now_s() - tr1));
            worker.join();
            h->icp0_valid = rc_icp == GOICP_OK;           // a failed head start is not a failed DT build: Register runs that ICP itself
            if (rc_icp) h->err.clear();
        } else {
            e = dt_build_device(h->model.data(), h->nm, S, h->p.dt_expand, h->p.dt_mode, h->d_dt.p, h->dt_meta, h->stream, msg);
        }
    } else {
        e = dt_build_device(h->model.data(), h->nm, S, h->p.dt_expand, h->p.dt_mode, h->d_dt.p, h->dt_meta, h->stream, msg);
    }
    if (e != cudaSuccess) { h->icp0_valid = false; return fail(h, GOICP_ERR_CUDA, "dt_build_device: " + msg + ": " + cudaGetErrorString(e)); }
    h->bytes_h2d += (int64_t)sizeof(float) * 3 * h->nm;          // dt_build_device uploads the model cloud itself
    h->dt_size = S; h->have_dt = true;
    if (trace) fprintf(stderr, "[dt trace] goicp_build_dt total %.3f ms\n", 1e3 * (now_s() - tr0));
    return GOICP_OK;
}

int goicp_dt_distance(goicp_handle* h, const float* q_xyz, int n, float* dist_out, int32_t* ixyz_out)
{
    if (!h || !q_xyz || !dist_out || n < 0) return fail(h, GOICP_ERR_INVALID, "dt_distance: bad arguments");
    if (!h->have_dt) return fail(h, GOICP_ERR_INVALID, "dt_distance: no distance transform");
    if (n == 0) return GOICP_OK;
    DtView dt; dt.tex = 0; dt.grid = h->d_dt.p; dt.S = h->dt_size; dt.S2 = h->dt_size * h->dt_size;
    dt.xmin = h->dt_meta[0]; dt.ymin = h->dt_meta[1]; dt.zmin = h->dt_meta[2]; dt.scale = h->dt_meta[3]; dt.inv_scale = 1.0 / h->dt_meta[3];
    CUDA_TRY(h, h->d_q.reserve((size_t)3 * n));
    CUDA_TRY(h, h->d_f32b.reserve(n));
    CUDA_TRY(h, h->d_i32.reserve((size_t)3 * n + 16));
    CUDA_TRY(h, xfer(h, h->d_q.p, q_xyz, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, launch_dt_lookup(dt, h->d_q.p, n, h->d_f32b.p, ixyz_out ? h->d_i32.p : nullptr, h->stream));
    CUDA_TRY(h, xfer(h, dist_out, h->d_f32b.p, sizeof(float) * n, cudaMemcpyDeviceToHost, h->stream));
    if (ixyz_out) CUDA_TRY(h, xfer(h, ixyz_out, h->d_i32.p, sizeof(int32_t) * 3 * n, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return GOICP_OK;
}

int goicp_eval_bounds(goicp_handle* h, int npairs, const float* R9, const int32_t* level, const float* tcube, float* ub_out, float* lb_out)
{
    if (!h || npairs < 0 || !R9 || !level || !tcube || !ub_out || !lb_out) return fail(h, GOICP_ERR_INVALID, "eval_bounds: bad arguments");
    if (npairs == 0) return GOICP_OK;
    BnbConst c; int rc = make_const(h, c); if (rc) return rc;
    if (c.inlier_num < c.nd) return fail(h, GOICP_ERR_INVALID, "eval_bounds: trimmed bounds are evaluated by goicp_inner_bnb / goicp_register");
    std::vector<PairTask> tasks(npairs);
    for (int k = 0; k < npairs; k++) {
        if (level[k] >= kMaxRotLevel) return fail(h, GOICP_ERR_DEPTH, "eval_bounds: rotation level >= 20");
        std::memcpy(tasks[k].R, R9 + 9 * (size_t)k, 9 * sizeof(float));
        tasks[k].level = level[k];
        std::memcpy(tasks[k].tc, tcube + 4 * (size_t)k, 4 * sizeof(float));
        tasks[k].pad[0] = tasks[k].pad[1] = 0;
    }
    CUDA_TRY(h, h->d_pairs.reserve(npairs));
    CUDA_TRY(h, h->d_f32b.reserve((size_t)2 * npairs));
    CUDA_TRY(h, xfer(h, h->d_pairs.p, tasks.data(), sizeof(PairTask) * npairs, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, launch_pair_bounds(c, h->d_pairs.p, npairs, reinterpret_cast<float2*>(h->d_f32b.p), h->stream));
    std::vector<float> out((size_t)2 * npairs);
    CUDA_TRY(h, xfer(h, out.data(), h->d_f32b.p, sizeof(float) * 2 * npairs, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    for (int k = 0; k < npairs; k++) { ub_out[k] = out[2 * k]; lb_out[k] = out[2 * k + 1]; }
    return GOICP_OK;
}

int goicp_expand_bounds(goicp_handle* h, int n, const float* R9, const int32_t* level, const float* tcube, float* out16, int repeats, float* device_ms)
{
    if (!h || n < 0 || !R9 || !level || !tcube || !out16) return fail(h, GOICP_ERR_INVALID, "expand_bounds: bad arguments");
    if (n == 0) return GOICP_OK;
    BnbConst c; int rc = make_const(h, c); if (rc) return rc;
    if (c.inlier_num < c.nd) return fail(h, GOICP_ERR_INVALID, "expand_bounds: trimmed bounds are evaluated by goicp_inner_bnb / goicp_register");
    std::vector<PairTask> tasks(n);
    for (int k = 0; k < n; k++) {
        if (level[k] >= kMaxRotLevel) return fail(h, GOICP_ERR_DEPTH, "expand_bounds: rotation level >= 20");
        std::memcpy(tasks[k].R, R9 + 9 * (size_t)k, 9 * sizeof(float));
        tasks[k].level = level[k];
        std::memcpy(tasks[k].tc, tcube + 4 * (size_t)k, 4 * sizeof(float));
        tasks[k].pad[0] = tasks[k].pad[1] = 0;
    }
    CUDA_TRY(h, h->d_pairs.reserve(n));
    CUDA_TRY(h, h->d_f32b.reserve((size_t)16 * n));
    CUDA_TRY(h, xfer(h, h->d_pairs.p, tasks.data(), sizeof(PairTask) * n, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, launch_expand_bounds(c, h->d_pairs.p, n, h->d_f32b.p, h->stream));
    if (device_ms) {
        if (repeats < 1) repeats = 1;
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
        CUDA_TRY(h, cudaEventRecord(h->ev0, h->stream));
        for (int r = 0; r < repeats; r++) CUDA_TRY(h, launch_expand_bounds(c, h->d_pairs.p, n, h->d_f32b.p, h->stream));
        CUDA_TRY(h, cudaEventRecord(h->ev1, h->stream));
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
        float ms = 0; CUDA_TRY(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
        *device_ms = ms / repeats;
    }
    CUDA_TRY(h, xfer(h, out16, h->d_f32b.p, sizeof(float) * 16 * n, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return GOICP_OK;
}

int goicp_inner_bnb(goicp_handle* h, int n, const float* R9, const int32_t* level, const float* opt_error, goicp_inner_result* out)
{
    if (!h || n < 0 || !R9 || !level || !opt_error || !out) return fail(h, GOICP_ERR_INVALID, "inner_bnb: bad arguments");
    if (n == 0) return GOICP_OK;
    BnbConst c; int rc = make_const(h, c); if (rc) return rc;
    rc = ensure_task_buffers(h, (size_t)n + 64); if (rc) return rc;
    for (int k = 0; k < n; k++) {
        if (level[k] >= kMaxRotLevel) return fail(h, GOICP_ERR_DEPTH, "inner_bnb: rotation level >= 20");
        std::memcpy(h->h_tasks[k].R, R9 + 9 * (size_t)k, 9 * sizeof(float));
        h->h_tasks[k].level = level[k]; h->h_tasks[k].opt_error = opt_error[k]; h->h_tasks[k].pad = 0;
    }
    goicp_allgather_fn saved = h->xchg; h->xchg = nullptr;       // this entry point is single-GPU
    NcclComm saved_nccl = h->nccl; h->nccl = nullptr;
    std::vector<std::shared_ptr<CandList>> lists;
    rc = run_inner_batch(h, c, n, nullptr, &lists);
    h->xchg = saved; h->nccl = saved_nccl;
    if (rc) return rc;
    std::vector<InnerTask> tasks(h->h_tasks, h->h_tasks + n);
    std::vector<InnerResult> results(h->h_results, h->h_results + n);
    for (int k = 0; k < n; k++)
        if (lists[k] && !(h->p.numerics & GOICP_NUM_FAST_SUMS)) { rc = resolve_strict(h, c, tasks[k], *lists[k], &results[k].value, results[k].node); if (rc) return rc; }
    std::memcpy(h->h_results, results.data(), sizeof(InnerResult) * n);
    for (int k = 0; k < n; k++) {
        out[k].value = h->h_results[k].value; std::memcpy(out[k].node, h->h_results[k].node, sizeof out[k].node);
        out[k].pops = h->h_results[k].pops; out[k].evals = h->h_results[k].evals; out[k].status = h->h_results[k].status;
        out[k].reuse_gt = (h->h_results[k].pad[0] & 1u) ? 3.402823466e+38f : h->h_results[k].reuse_gt; out[k].reuse_poplb = h->h_results[k].reuse_poplb;
    }
    return GOICP_OK;
}

int goicp_nn(goicp_handle* h, const float* q_xyz, int n, int32_t* idx_out, float* d2_out)
{
    if (!h || !q_xyz || n < 0 || !idx_out || !d2_out) return fail(h, GOICP_ERR_INVALID, "nn: bad arguments");
    if (n == 0) return GOICP_OK;
    int rc = ensure_cuda(h); if (rc) return rc;
    rc = ensure_kdtree(h); if (rc) return rc;
    CUDA_TRY(h, h->d_q.reserve((size_t)3 * n));
    CUDA_TRY(h, h->d_f32b.reserve(n));
    CUDA_TRY(h, h->d_i32.reserve((size_t)n + 16));
    CUDA_TRY(h, xfer(h, h->d_q.p, q_xyz, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, launch_nn(kd_view(h), h->d_q.p, n, h->d_i32.p, h->d_f32b.p, getenv("GOICP_NN_COOP") != nullptr, h->stream));
    CUDA_TRY(h, xfer(h, idx_out, h->d_i32.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, xfer(h, d2_out, h->d_f32b.p, sizeof(float) * n, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return GOICP_OK;
}

int goicp_svd3(goicp_handle* h, const float* H9, int n, float* U9_out, float* W3_out, float* V9_out)
{
    if (!h || !H9 || n < 0 || !U9_out || !W3_out || !V9_out) return fail(h, GOICP_ERR_INVALID, "svd3: bad arguments");
    if (n == 0) return GOICP_OK;
    int rc = ensure_cuda(h); if (rc) return rc;
    CUDA_TRY(h, h->d_q.reserve((size_t)30 * n));
    float* dH = h->d_q.p; float* dU = dH + (size_t)9 * n; float* dW = dU + (size_t)9 * n; float* dV = dW + (size_t)3 * n;
    CUDA_TRY(h, xfer(h, dH, H9, sizeof(float) * 9 * n, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, launch_svd3(dH, n, dU, dW, dV, h->stream));
    CUDA_TRY(h, xfer(h, U9_out, dU, sizeof(float) * 9 * n, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, xfer(h, W3_out, dW, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, xfer(h, V9_out, dV, sizeof(float) * 9 * n, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return GOICP_OK;
}

int goicp_bnb_shape_rule(double forecast_longest_pops, double forecast_sum_pops_this_rank, double previous_longest_pops, int cluster_slots)
{
    return bnb_shape_rule(forecast_longest_pops, forecast_sum_pops_this_rank, previous_longest_pops, (double)std::max(1, cluster_slots));
}
int goicp_intro_select(goicp_handle* h, float* a, int n, int k, int threads, int in_global)
{
    if (!h || !a || n <= 0 || k < 0 || k >= n || threads < 32 || threads > 1024 || (threads & 31)) return fail(h, GOICP_ERR_INVALID, "intro_select: bad arguments");
    int rc = ensure_cuda(h); if (rc) return rc;
    CUDA_TRY(h, h->d_q.reserve((size_t)n));
    CUDA_TRY(h, h->d_i32.reserve((size_t)2 * n));
    CUDA_TRY(h, xfer(h, h->d_q.p, a, sizeof(float) * n, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, launch_select_test(h->d_q.p, n, k, h->d_i32.p, threads, h->max_smem_optin - 2048, !in_global, h->stream));
    CUDA_TRY(h, xfer(h, a, h->d_q.p, sizeof(float) * n, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return GOICP_OK;
}

int goicp_kdtree_host(const float* model_xyz, int n, int32_t* nodes7_out, int capacity_nodes, int32_t* vind_out, float* bbox6_out)
{
    if (!model_xyz || n <= 0) return -GOICP_ERR_INVALID;
    HostKdTree t;
    t.build(model_xyz, n, 10);
    const int nn = (int)t.nodes.size();
    if (nn > capacity_nodes || !nodes7_out || !vind_out || !bbox6_out) return nn;
    for (int i = 0; i < nn; i++) {
        const KdNode& k = t.nodes[i];
        int32_t* o = nodes7_out + 7 * (size_t)i;
        o[0] = k.child1; o[1] = k.child2; o[2] = k.left; o[3] = k.right; o[4] = k.divfeat;
        std::memcpy(&o[5], &k.divlow, 4); std::memcpy(&o[6], &k.divhigh, 4);
    }
    std::memcpy(vind_out, t.vind.data(), sizeof(int32_t) * (size_t)n);
    for (int a = 0; a < 3; a++) { bbox6_out[a] = t.bb_lo[a]; bbox6_out[3 + a] = t.bb_hi[a]; }
    return nn;
}

int goicp_icp(goicp_handle* h, const float R0[9], const float t0[3], int max_iter, float err_diff, goicp_icp_result* out)
{
    if (!h || !R0 || !t0 || !out) return fail(h, GOICP_ERR_INVALID, "icp: bad arguments");
    if (max_iter <= 0) max_iter = h->p.icp_max_iter;
    if (err_diff < 0) err_diff = h->p.mse_threshold / 10000;
    return run_icp(h, R0, t0, max_iter, err_diff, out);
}

int goicp_dt_score(goicp_handle* h, const float* R, const float* t, float* sse_out)
{
    if (!h || !sse_out || (R && !t)) return fail(h, GOICP_ERR_INVALID, "dt_score: bad arguments");
    BnbConst c; int rc = make_const(h, c); if (rc) return rc;
    return score_pose(h, c, R, t, sse_out);
}

int goicp_icp_dt(goicp_handle* h, float R[9], float t[3], float* dt_error_out)
{
    if (!h || !R || !t || !dt_error_out) return fail(h, GOICP_ERR_INVALID, "icp_dt: bad arguments");
    BnbConst c; int rc = make_const(h, c); if (rc) return rc;
    return icp_then_dt(h, c, R, t, dt_error_out);
}

int goicp_nccl_unique_id(void* id128_out)
{
    if (!id128_out) return GOICP_ERR_INVALID;
    NcclApi* a = nccl_api();
    if (!a->err.empty()) return GOICP_ERR_CUDA;
    NcclUniqueId id;
    if (a->GetUniqueId(&id) != 0) return GOICP_ERR_CUDA;
    std::memcpy(id128_out, &id, sizeof id);
    return GOICP_OK;
}

int goicp_nccl_init(goicp_handle* h, const void* id128, int rank, int world_size)
{
    if (!h || !id128 || world_size < 1 || rank < 0 || rank >= world_size) return fail(h, GOICP_ERR_INVALID, "nccl_init: bad arguments");
    NcclApi* a = nccl_api();
    if (!a->err.empty()) return fail(h, GOICP_ERR_CUDA, a->err);
    int rc = ensure_cuda(h); if (rc) return rc;
    // communicators are process-wide, keyed by the unique id: creating one is a collective costing ~0.1 s, and the
    // handles of one process (one GPU) share it -- calls on them are serialised by the caller anyway
    static std::mutex comm_mtx;
    static std::unordered_map<std::string, NcclComm> comms;
    NcclUniqueId id; std::memcpy(&id, id128, sizeof id);
    const std::string key(id.internal, sizeof id.internal);
    std::lock_guard<std::mutex> lk(comm_mtx);
    auto it = comms.find(key);
    if (it == comms.end()) {
        NcclComm comm = nullptr;
        const int nrc = a->CommInitRank(&comm, world_size, id, rank);
        if (nrc != 0) return fail(h, GOICP_ERR_CUDA, std::string("ncclCommInitRank: ") + (a->GetErrorString ? a->GetErrorString(nrc) : "error"));
        it = comms.emplace(key, comm).first;
    }
    h->nccl = it->second;
    h->p.rank = rank; h->p.world_size = world_size;
    return GOICP_OK;
}

int goicp_set_exchange(goicp_handle* h, goicp_allgather_fn fn, void* user, int use_device_buffers)
{
    if (!h) return GOICP_ERR_INVALID;
    h->xchg = fn; h->xchg_user = user; h->xchg_device = use_device_buffers;
    return GOICP_OK;
}

// Host-only self test of the multi-rank result exchange (no GPU): fabricates the results of `n`
// tasks, runs the same shard_exchange() as a real round and counts wrong records.
int goicp_selftest_shard(int rank, int world, int n, goicp_allgather_fn fn, void* user, int* mismatches)
{
    if (!fn || !mismatches || world < 1 || rank < 0 || rank >= world) return GOICP_ERR_INVALID;
    auto fake = [](int t) { InnerResult r; std::memset(&r, 0, sizeof r); r.value = (float)t * 0.5f; r.pops = 7u * t + 1; r.evals = 8u * t; r.node[3] = (float)t; return r; };
    std::vector<InnerResult> mine, all(std::max(n, 1));
    for (int t = rank; t < n; t += world) mine.push_back(fake(t));
    if (shard_exchange(fn, user, world, n, mine.data(), (int)mine.size(), all.data()) != 0) return GOICP_ERR_INVALID;
    int bad = 0;
    for (int t = 0; t < n; t++) { InnerResult e = fake(t); if (std::memcmp(&e, &all[t], sizeof e) != 0) bad++; }
    // the hand-round of an upper-bound pass's contender list (share_cand_list): whoever holds it, every rank ends with it
    {
        goicp_handle* h = new goicp_handle();
        h->p.world_size = world; h->p.rank = rank; h->xchg = fn; h->xchg_user = user;
        auto pattern = [](int owner) { auto cl = std::make_shared<CandList>(); std::memset(cl.get(), 0, sizeof(CandList));
                                       cl->n = 3 + owner; cl->flags = 0x55u + owner; cl->final_fast = 1.5f * owner; cl->eps = 1e-6f;
                                       for (int q = 0; q < cl->n; q++) { cl->node[q] = make_float4(0.25f * q, -0.5f * owner, 0.125f, 1.0f / (q + 1)); cl->ub[q] = 10.0f + q + owner; }
                                       return cl; };
        for (int owner = 0; owner < world; owner++) {
            std::shared_ptr<CandList> cl = rank == owner ? pattern(owner) : nullptr;
            if (share_cand_list(h, cl) != GOICP_OK || !cl || std::memcmp(cl.get(), pattern(owner).get(), sizeof(CandList)) != 0) bad++;
        }
        delete h;
    }
    *mismatches = bad;
    return GOICP_OK;
}

int goicp_trim_memory(void) { pool_trim(); return GOICP_OK; }

// Random 4-byte gather rate over a `bytes`-sized device buffer (the DT gathers' roofline denominator; bnb_kernels.cu)
int goicp_measure_gather(goicp_handle* h, size_t bytes, int repeats, double* lookups_per_s_out)
{
    if (!h || bytes < 4096 || !lookups_per_s_out) return fail(h, GOICP_ERR_INVALID, "measure_gather: bad arguments");
    int rc = ensure_cuda(h); if (rc) return rc;
    if (repeats < 1) repeats = 1;
    float* buf = nullptr;
    CUDA_TRY(h, pool_alloc((void**)&buf, bytes + 256));
    const unsigned n = (unsigned)std::min<size_t>(bytes / 4, 0xffffffffu);
    cudaError_t e = cudaMemsetAsync(buf, 0, bytes, h->stream);
    const int blocks = h->sm_count * 4, iters = 256;
    float* sink = buf + bytes / 4;
    if (e == cudaSuccess) e = launch_gather_peak(buf, n, iters, blocks, sink, h->stream);          // warm-up (fills L2 when the buffer fits)
    if (e == cudaSuccess) e = launch_gather_peak(buf, n, iters, blocks, sink, h->stream);
    if (e == cudaSuccess) e = cudaEventRecord(h->ev0, h->stream);
    for (int r = 0; r < repeats && e == cudaSuccess; r++) e = launch_gather_peak(buf, n, iters, blocks, sink, h->stream);
    if (e == cudaSuccess) e = cudaEventRecord(h->ev1, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    float ms = 0;
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, h->ev0, h->ev1);
    pool_free(buf);
    if (e != cudaSuccess) return fail(h, GOICP_ERR_CUDA, std::string("measure_gather: ") + cudaGetErrorString(e));
    *lookups_per_s_out = (double)blocks * 512.0 * 8.0 * iters * repeats / (ms * 1e-3);
    return GOICP_OK;
}

int goicp_transfer_bytes(const goicp_handle* h, int64_t* h2d_out, int64_t* d2h_out)
{
    if (!h) return GOICP_ERR_INVALID;
    if (h2d_out) *h2d_out = h->bytes_h2d;
    if (d2h_out) *d2h_out = h->bytes_d2h;
    return GOICP_OK;
}

int goicp_cancel(goicp_handle* h) { if (!h) return GOICP_ERR_INVALID; h->cancel_flag.store(1); return GOICP_OK; }
int goicp_poll(goicp_handle* h, goicp_snapshot* out)
{
    if (!h || !out) return GOICP_ERR_INVALID;
    std::lock_guard<std::mutex> lk(h->snap_mtx);
    *out = h->snap;
    return GOICP_OK;
}

// GoICP::Register = Initialize + OuterBnB (jly_goicp.cpp:569-585, 342-567)
int goicp_register(goicp_handle* h, goicp_result* out)
{
    if (!h || !out) return fail(h, GOICP_ERR_INVALID, "register: bad arguments");
    std::memset(out, 0, sizeof *out);
    h->cancel_flag.store(0);
    h->t_kernels = 0; h->t_icp = 0; h->launches = 0; h->low_latency = true; h->dense = 0; h->prev_max_pops = 0; h->pred_max = h->pred_sum = 0; h->t_score = h->t_strict = 0; h->strict_resolves = 0; h->cand_overflows = 0; h->bnb_variants = 0;
    if (h->p.search_mode == GOICP_SEARCH_FGOICP) return register_fgoicp(h, out);
    const double t_begin = now_s();
    h->initialized = false;
    int rc = initialize(h); if (rc) return rc;
    BnbConst c; rc = make_const(h, c); if (rc) return rc;
    rc = ensure_kdtree(h); if (rc) return rc;
    h->t_setup = now_s() - t_begin;

    goicp_result res; std::memset(&res, 0, sizeof res);
    res.sse_thresh = h->sse_thresh;
    float E = 1e+10f;                                    // optError
    float optR[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, optT[3] = {0, 0, 0};
    auto sync_res = [&]() { std::memcpy(res.R, optR, sizeof optR); std::memcpy(res.t, optT, sizeof optT); res.sse = E; };

    // initial error (jly_goicp.cpp:357-372) and ICP from the identity (:378-391)
    rc = score_pose(h, c, nullptr, nullptr, &E); if (rc) return rc;
    {
        float R_icp[9], t_icp[3], error;
        std::memcpy(R_icp, optR, sizeof optR); std::memcpy(t_icp, optT, sizeof optT);
        // goicp_build_dt may already have run exactly this refinement next to the DT build (one use only)
        const bool pre = h->icp0_valid; h->icp0_valid = false;
        rc = icp_then_dt(h, c, R_icp, t_icp, &error, pre ? &h->icp0 : nullptr); if (rc) return rc;
        res.icp_calls++;
        if (error < E) { E = error; std::memcpy(optR, R_icp, sizeof optR); std::memcpy(optT, t_icp, sizeof optT); }
    }
    sync_res(); publish(h, res, 0);

    std::vector<RotNode> heap;                            // std::priority_queue<ROTNODE> == vector + push_heap/pop_heap
    RotLower lower;
    RotNode init; init.a = h->p.rot_cube[0]; init.b = h->p.rot_cube[1]; init.c = h->p.rot_cube[2]; init.w = h->p.rot_cube[3];
    init.ub = 0; init.lb = 0; init.l = 0;
    heap.push_back(init); std::push_heap(heap.begin(), heap.end(), lower);

    std::unordered_map<CubeKey, CubeEval, CubeKeyHash> cache;
    long epoch = 0;
    const bool reuse_ok = !(getenv("GOICP_NO_REUSE") && atoi(getenv("GOICP_NO_REUSE")) != 0);      // A/B switch (tests): 1 = drop everything on an improvement
    long st_cubes_evaluated = 0, st_cubes_dropped_by_improvement = 0, st_tasks = 0, st_tasks_reused = 0, st_tasks_kept = 0, st_tasks_invalid = 0;       // GOICP_ROUND_STATS
    // default speculation width: ~4 SMs' worth of clusters per cube (36 cubes = 576 inner BnBs per round on 148 SMs)
    int spec = h->p.spec_cubes > 0 ? h->p.spec_cubes : std::max(1, h->sm_count / 4);
    if ((h->xchg || h->nccl) && h->p.world_size > 1) spec *= h->p.world_size;

    // Evaluate `first` (must be evaluated) plus the best not-yet-evaluated queue entries.
    auto evaluate_round = [&](const RotNode& first) -> int {
        std::vector<RotNode> cubes;
        cubes.push_back(first);
        if (spec > 1 && !heap.empty()) {
            std::vector<const RotNode*> cand;
            cand.reserve(heap.size());
            const CubeKey fk = key_of(first);
            for (const RotNode& n : heap) {
                const CubeKey k = key_of(n);
                if (k == fk) continue;
                auto it = cache.find(k);
                if (it != cache.end() && it->second.epoch == epoch && it->second.complete()) continue;
                if ((E - n.lb) <= h->sse_thresh) continue;         // would certify, never expanded
                cand.push_back(&n);
            }
            const size_t want = std::min<size_t>(cand.size(), (size_t)spec - 1);
            std::partial_sort(cand.begin(), cand.begin() + want, cand.end(),
                              [&](const RotNode* x, const RotNode* y) { return lower(*y, *x); });
            for (size_t i = 0; i < want; i++) cubes.push_back(*cand[i]);
        }
        // task list: for every in-ball child an upper-bound pass and a lower-bound pass -- those the cache does not already
        // hold (a cube evaluated before the last improvement keeps the passes that are provably unchanged, see revalidate)
        struct Slot { size_t cube; int child; int pass; };
        std::vector<Slot> slots;
        std::vector<CubeEval> evs(cubes.size());
        int rcl = ensure_task_buffers(h, cubes.size() * 16 + 64); if (rcl) return rcl;
        int n = 0;
        for (size_t ci = 0; ci < cubes.size(); ci++) {
            const RotNode& P = cubes[ci];
            if (P.l + 1 >= kMaxRotLevel) return fail(h, GOICP_ERR_DEPTH, "rotation search reached level 20 (the reference indexes maxRotDis out of bounds there)");
            {
                auto it = cache.find(key_of(P));
                if (it != cache.end() && it->second.epoch == epoch) evs[ci] = it->second;
            }
            const float w = P.w / 2;
            for (int j = 0; j < 8; j++) {
                const float a = P.a + (j & 1) * w, b = P.b + (j >> 1 & 1) * w, cc = P.c + (j >> 2 & 1) * w;
                ChildEval& ce = evs[ci].ch[j];
                ce.skipped = !cube_rotation(a, b, cc, w, ce.R);
                if (ce.skipped) continue;
                for (int pass = 0; pass < 2; pass++) {
                    if (pass == 0 ? ce.have_ub : ce.have_lb) { st_tasks_reused++; continue; }
                    InnerTask& t = h->h_tasks[n++];
                    std::memcpy(t.R, ce.R, sizeof ce.R);
                    t.level = pass == 0 ? -1 : P.l + 1;              // maxRotDis[nodeRot.l] (jly_goicp.cpp:551)
                    t.opt_error = E; t.pad = 0;
                    slots.push_back(Slot{ci, j, pass});
                }
            }
        }
        // Launch order = longest expected first.  Clusters are dispatched in task order and a round lasts until its
        // last task retires; an inner BnB of a child costs about what the same pass cost on its parent (same
        // neighbourhood of rotation space, half the uncertainty radius), so sorting by the parent's pop counts keeps
        // the long searches out of the tail.  Results are per task, so the order changes no value; every rank
        // derives the same permutation from replicated data.
        {
            std::vector<int> ord(n);
            for (int t = 0; t < n; t++) ord[t] = t;
            auto cost = [&](int t) { const RotNode& P = cubes[slots[t].cube]; return slots[t].pass == 0 ? P.pops_ub : P.pops_lb; };
            std::stable_sort(ord.begin(), ord.end(), [&](int x, int y) { return cost(x) > cost(y); });
            std::vector<InnerTask> tk(n); std::vector<Slot> sl(n);
            for (int t = 0; t < n; t++) { tk[t] = h->h_tasks[ord[t]]; sl[t] = slots[ord[t]]; }
            std::memcpy(h->h_tasks, tk.data(), sizeof(InnerTask) * n); slots.swap(sl);
            h->pred_max = n > 0 ? (double)cost(0) : 0; h->pred_sum = 0;
            for (int t = 0; t < n; t++) h->pred_sum += (double)cost(t);
        }
        std::vector<std::shared_ptr<CandList>> lists;
        rcl = run_inner_batch(h, c, n, &res.bound_evals_executed, &lists, &res.bound_evals_executed_local); if (rcl) return rcl;
        res.rounds++; st_cubes_evaluated += (long)cubes.size(); st_tasks += n;
        for (int t = 0; t < n; t++) {
            ChildEval& ce = evs[slots[t].cube].ch[slots[t].child];
            (slots[t].pass == 0 ? ce.ub : ce.lb) = h->h_results[t];
            if (slots[t].pass == 0) { ce.cands = lists[t]; ce.ub_opt_error = E; ce.have_ub = true; }
            else { ce.lb_opt_error = E; ce.have_lb = true; }
        }
        for (size_t ci = 0; ci < cubes.size(); ci++) { evs[ci].epoch = epoch; cache[key_of(cubes[ci])] = evs[ci]; }
        return GOICP_OK;
    };

    int exit_path = GOICP_EXIT_NONE;
    float exit_lb = 0;
    while (exit_path == GOICP_EXIT_NONE) {
        if (h->cancel_flag.load()) { exit_path = GOICP_EXIT_CANCELLED; break; }           // goicp_finished (jly_goicp.cpp:400)
        if (heap.empty()) { exit_path = GOICP_EXIT_QUEUE_EMPTY; break; }                  // :402-407
        RotNode P = heap.front();
        if ((E - P.lb) <= h->sse_thresh) {                                                // :416-420
            std::pop_heap(heap.begin(), heap.end(), lower); heap.pop_back(); res.rot_pops++;
            exit_path = GOICP_EXIT_CERTIFIED; exit_lb = P.lb; break;
        }
        {
            auto it = cache.find(key_of(P));
            if (it == cache.end() || it->second.epoch != epoch || !it->second.complete()) { rc = evaluate_round(P); if (rc) return rc; }
        }
        std::pop_heap(heap.begin(), heap.end(), lower); heap.pop_back(); res.rot_pops++;
        const CubeKey pk = key_of(P);
        cache[pk].committed = true;
        int j0 = 0; bool skip_ub = false;
        bool restart = true;
        while (restart && exit_path == GOICP_EXIT_NONE) {
            restart = false;
            const CubeEval ev = cache[pk];
            const float w = P.w / 2;
            for (int j = j0; j < 8; j++) {
                const ChildEval& ce = ev.ch[j];
                if (ce.skipped) continue;
                RotNode nd; nd.w = w; nd.l = P.l + 1;
                nd.a = P.a + (j & 1) * w; nd.b = P.b + (j >> 1 & 1) * w; nd.c = P.c + (j >> 2 & 1) * w;
                float ub = ce.ub.value;
                float ub_node[4] = {ce.ub.node[0], ce.ub.node[1], ce.ub.node[2], ce.ub.node[3]};
                if (!(j == j0 && skip_ub)) {
                    res.trans_pops += ce.ub.pops; res.bound_evals += ce.ub.evals;
                    if (ce.ub.pad[1] > 0 && !(h->p.numerics & GOICP_NUM_FAST_SUMS)) {
                        // this pass may improve the optimum: settle value and arg-min cube in the
                        // reference's summation order before deciding (strict_sum.cuh)
                        InnerTask task; std::memcpy(task.R, ce.R, sizeof ce.R); task.level = -1; task.opt_error = ce.ub_opt_error; task.pad = 0;
                        std::shared_ptr<CandList> cl = ce.cands;
                        if ((h->xchg || h->nccl) && h->p.world_size > 1) { rc = share_cand_list(h, cl); if (rc) return rc; }   // collective: every rank is here
                        if (cl) {
                            if (cl->flags & 1u) h->cand_overflows++;          // the list is the first 128 contenders only (goicp_result.contender_overflows)
                            rc = resolve_strict(h, c, task, *cl, &ub, ub_node); if (rc) return rc;
                        }
                    }
                    if (ub < E) {                                                          // :495-544
                        E = ub;
                        std::memcpy(optR, ce.R, sizeof optR);
                        optT[0] = ub_node[0] + ub_node[3] / 2; optT[1] = ub_node[1] + ub_node[3] / 2; optT[2] = ub_node[2] + ub_node[3] / 2;
                        float R_icp[9], t_icp[3], error;
                        std::memcpy(R_icp, optR, sizeof optR); std::memcpy(t_icp, optT, sizeof optT);
                        rc = icp_then_dt(h, c, R_icp, t_icp, &error); if (rc) return rc;
                        res.icp_calls++;
                        if (error < E) { E = error; std::memcpy(optR, R_icp, sizeof optR); std::memcpy(optT, t_icp, sizeof optT); }
                        sync_res(); publish(h, res, 0);
                        if (E < h->sse_thresh) { exit_path = GOICP_EXIT_EARLY_SSE; break; }       // :527-530
                        {   // discard queue nodes with lb >= E, rebuilding in pop order (:533-543)
                            std::vector<RotNode> fresh;
                            while (!heap.empty()) {
                                std::pop_heap(heap.begin(), heap.end(), lower);
                                RotNode n2 = heap.back(); heap.pop_back();
                                if (n2.lb < E) { fresh.push_back(n2); std::push_heap(fresh.begin(), fresh.end(), lower); }
                                else break;
                            }
                            heap.swap(fresh);
                        }
                        {
                            // Speculative results were computed under the old incumbent.  An inner BnB depends on its initial
                            // optErrorT only through decisions taken before its own first improvement; the kernel reports the
                            // range of smaller values for which every one of them comes out the same (InnerResult::reuse_*).
                            // Passes inside that range are kept -- same pops, same arg-min, and a value that was the initial
                            // optErrorT itself becomes the new one --, the others are run again when their cube is needed.
                            // (Measured on test/spanner_goicp.toml: one improvement of 8e-4 relative after 1 921 rotation pops
                            // found 2 490 evaluated cubes waiting in the cache -- a third of the run's work.)
                            long kept = 0, invalid = 0, dropped = 0;
                            for (auto it = cache.begin(); it != cache.end();) {
                                CubeEval& cev = it->second;
                                if (cev.committed && !(it->first == pk)) { it = cache.erase(it); continue; }
                                bool any = false;
                                for (int q = 0; q < 8; q++) {
                                    ChildEval& cq = cev.ch[q];
                                    if (cq.skipped) continue;
                                    if (it->first == pk && (q < j || (q == j))) {       // P itself: children before j are done, child j's ub pass too
                                        if (q < j) { cq.have_ub = cq.have_lb = true; continue; }
                                        cq.have_ub = true;
                                    } else if (cq.have_ub) {
                                        if (reuse_ok && reusable_under(cq.ub, E, h->sse_thresh) && !(cq.ub.pad[0] & 1u)) {
                                            if (cq.ub.value == cq.ub_opt_error) cq.ub.value = E;
                                            cq.ub_opt_error = E; kept++; any = true;
                                        } else { cq.have_ub = false; cq.cands.reset(); invalid++; }
                                    }
                                    if (cq.have_lb) {
                                        if (reuse_ok && reusable_under(cq.lb, E, h->sse_thresh)) {
                                            if (cq.lb.value == cq.lb_opt_error) cq.lb.value = E;
                                            cq.lb_opt_error = E; kept++; any = true;
                                        } else { cq.have_lb = false; invalid++; }
                                    }
                                }
                                if (!any && !(it->first == pk)) { dropped++; it = cache.erase(it); continue; }
                                cev.epoch = epoch + 1;
                                ++it;
                            }
                            st_cubes_dropped_by_improvement += dropped; st_tasks_kept += kept; st_tasks_invalid += invalid;
                            if (getenv("GOICP_ROUND_STATS")) fprintf(stderr, "[improvement] after %lld rotation pops: E -> %.9g; cached passes kept %ld, to be re-run %ld, cubes dropped %ld\n", (long long)res.rot_pops, (double)E, kept, invalid, dropped);
                        }
                        epoch++;
                        if (!cache[pk].complete()) { rc = evaluate_round(P); if (rc) return rc; }                        // rest of P under the new E
                        j0 = j; skip_ub = true; restart = true;
                        break;
                    }
                }
                const float lb = ce.lb.value;                                              // :551
                res.trans_pops += ce.lb.pops; res.bound_evals += ce.lb.evals;
                if (lb >= E) continue;                                                    // :554
                nd.ub = ub; nd.lb = lb; nd.pops_ub = ce.ub.pops; nd.pops_lb = ce.lb.pops;
                heap.push_back(nd); std::push_heap(heap.begin(), heap.end(), lower);      // :560-562
            }
            if (!restart) break;
        }
        if ((res.rot_pops & 15) == 0) { sync_res(); publish(h, res, 0); }
    }

    sync_res();
    res.exit_path = exit_path; res.best_lb = exit_lb;
    res.kernel_launches = h->launches;
    res.seconds_total = now_s() - t_begin; res.seconds_bnb_kernels = h->t_kernels; res.seconds_icp = h->t_icp;
    res.strict_resolves = h->strict_resolves; res.contender_overflows = h->cand_overflows; res.bnb_kernel_variants = h->bnb_variants;
    res.seconds_dt_score = h->t_score; res.seconds_strict = h->t_strict; res.seconds_setup = h->t_setup;
    res.seconds_host = res.seconds_total - h->t_setup - h->t_kernels - h->t_icp - h->t_score - h->t_strict;
    if (getenv("GOICP_ROUND_STATS")) {
        long left = 0;
        for (const auto& kv : cache) left += !kv.second.committed;
        fprintf(stderr, "[register] cubes evaluated %ld (tasks run %ld, taken from the cache across an improvement %ld of %ld kept; %ld invalidated) in %d rounds, committed %lld, cubes dropped by an improvement %ld, evaluated but never popped %ld\n",
                st_cubes_evaluated, st_tasks, st_tasks_reused, st_tasks_kept, st_tasks_invalid, (int)res.rounds, (long long)res.rot_pops, st_cubes_dropped_by_improvement, left);
    }
    if (getenv("GOICP_ROUND_STATS"))
        fprintf(stderr, "[register] total %.3f s: setup (upload, gamma table, kd-tree) %.3f, BnB kernels + exchange %.3f, ICP %.3f, DT scoring %.3f, strict resolves %.3f (%lld), rest (host commit, copies) %.3f\n",
                res.seconds_total, h->t_setup, h->t_kernels, h->t_icp, h->t_score, h->t_strict, (long long)h->strict_resolves,
                res.seconds_total - h->t_setup - h->t_kernels - h->t_icp - h->t_score - h->t_strict);
    publish(h, res, 1);
    *out = res;
    return exit_path == GOICP_EXIT_CANCELLED ? GOICP_ERR_CANCELLED : GOICP_OK;
}

} // extern "C"
