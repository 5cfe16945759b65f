// dt_kernels.cu -- distance-transform builders on sm_100a (GoICP::BuildDT, jly_goicp.cpp:75-90).
//
// GOICP_DT_REFERENCE reproduces DT3D::Build (jly_3ddt.cpp:889-979) bit for bit.  The reference
// transform is a *sequential* signed-less vector propagation ("DEuclidean", :710-742): every
// voxel update reads voxels updated earlier in the same raster scan, each x-slice is swept by
// rows forward then backward, each row along z forward then backward, and the whole volume in
// +x then -x.  Its output depends on that order (it is not the exact EDT), so the order is
// kept: one persistent CTA walks slices and rows in the reference order; within a row all the
// candidates that do not depend on the running scan are evaluated in parallel and only the true
// recurrence -- "previous voxel of this scan + (0,0,1)" -- is resolved, as a fixed point.
// Comparisons use the integer squared norm: the reference compares float(sqrt(v^2+h^2+d^2)), a
// strictly increasing function of that integer for grids up to 1024^3, with strict '<' in mask
// order (first minimum wins).
//
// GOICP_DT_EXACT_EDT is the separable exact squared-Euclidean transform (three 1-D lower-envelope
// passes over integer squared distances), fully parallel.
#include <ctime>
#include "dt_kernels.h"
#include "mem_pool.h"
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

namespace goicp {

namespace {

constexpr int kInf = 0x3fffffff;       // "no candidate" squared norm
constexpr int kMaxS = 1024;

// Working voxel, in HBM and on chip alike, 8 bytes: the vector to the nearest seed found so far
// (component magnitudes, 10 bits each: w = v | h<<10 | d<<20) and its squared norm n.  "unset"
// (no seed seen yet) = all-ones components with n = kInf.  Shared-memory bandwidth and warp
// shuffles are what the propagation kernel runs out of first, so the record is as small as the
// arithmetic allows; a candidate still costs a handful of integer instructions:
//   |(v+a, h+b, d+c)|^2 = n + 2(a v + b h + c d) + (a+b+c)   for a,b,c in {0,1},
// and adding (a,b,c) to the vector is one add on w.
struct __align__(8) V2 { unsigned w; int n; };
constexpr unsigned kUnsetW = 0x3fffffffu;

__device__ __forceinline__ V2 v2_unset() { V2 u; u.w = kUnsetW; u.n = kInf; return u; }
__device__ __forceinline__ int pitch_of(int S) { return (S + 1) & ~1; }       // rows of G are 16-byte aligned for the bulk copies

__global__ void dt_init_kernel(V2* G, size_t n, int corner_seed)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    V2 u = v2_unset();
    // As compiled (g++ 13.3 -O2) the reference's mask function returns (0,0,0) for the very first
    // voxel it visits, where no candidate qualifies and its result struct is uninitialised
    // (jly_3ddt.cpp:469-470): voxel (0,0,0) acts as one extra seed.  Pinned against oracle/_ref.
    if (i == 0 && corner_seed) { u.w = 0; u.n = 0; }
    G[i] = u;
}

// seeds: ROUND((p - min)*scale) in double, points outside the grid skipped (jly_3ddt.cpp:952-966)
__global__ void dt_seed_kernel(V2* G, int S, int P, const float* __restrict__ model, int nm, double xmin, double ymin, double zmin, double scale)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nm) return;
    int x = __double2int_rz(__dadd_rn(__dmul_rn(__dsub_rn((double)model[3 * i], xmin), scale), 0.5));
    int y = __double2int_rz(__dadd_rn(__dmul_rn(__dsub_rn((double)model[3 * i + 1], ymin), scale), 0.5));
    int z = __double2int_rz(__dadd_rn(__dmul_rn(__dsub_rn((double)model[3 * i + 2], zmin), scale), 0.5));
    if (x < 0 || x >= S || y < 0 || y >= S || z < 0 || z >= S) return;
    V2 s; s.w = 0; s.n = 0;
    G[((size_t)x * S + y) * P + z] = s;
}

// ---- the sequential propagation ------------------------------------------------------------
// A row scan of the reference is the recurrence
//   state_k = (state_{k-1} + (0,0,1) beats T_k) ? state_{k-1} + (0,0,1) : own_k
// where own_k is the first minimum (mask order) of all candidates that do not depend on the
// running scan and T_k the squared norm the running-scan candidate must beat at k (ties by mask
// order).  own/T of a whole row are computed in parallel; the recurrence has a unique solution,
// so any fixed point of the parallel update "every voxel re-evaluates the rule against its
// predecessor's current state" IS the sequential result, and starting from state = own that
// iteration converges in (longest run + 1) steps -- runs are short (1.3 voxels on average on the
// bunny), most rows have none.
// Mapping: each warp owns 30 consecutive voxels of the row; lanes 0 and 31 shadow the boundary
// voxels of the neighbouring warps, so a warp resolves both scans of a row with shuffles only,
// speculating that no run crosses into it.  The owner of a boundary voxel knows whether that
// held (its state differs from its own value); the flags are gathered by the single barrier of
// the row, which also publishes the row to the neighbours.  Rows where the speculation failed
// (3 %) exchange the warps' boundary states through shared memory and re-resolve.
constexpr int kOwned = 30;             // voxels owned per warp
constexpr int kMaxRefS = 32 * kOwned;  // 1024 threads; also keeps every component below the 10-bit unset marker

// IV*v + IH*h + ID*d out of the packed vector
template <int IV, int IH, int ID>
__device__ __forceinline__ int vec_dot(unsigned w)
{
    if (IV && IH && ID) return (int)((w & 1023u) + ((w >> 10) & 1023u) + (w >> 20));
    if (IV && IH) { const unsigned y = w & 0xfffffu; return (int)((y & 1023u) + (y >> 10)); }
    if (IV && ID) return (int)((w & 1023u) + (w >> 20));
    if (IH && ID) { const unsigned x = w >> 10; return (int)((x & 1023u) + (x >> 10)); }
    if (IV) return (int)(w & 1023u);
    if (IH) return (int)((w >> 10) & 1023u);
    if (ID) return (int)(w >> 20);
    return 0;
}
// candidate = source + (IV,IH,ID); strict '<' keeps the first minimum in mask order.  An unset
// source (n = kInf) scores above kInf and never wins (the reference scores it ~56755 > 32767);
// its components overflow their fields harmlessly.
template <int IV, int IH, int ID>
__device__ __forceinline__ void consider(V2& best, const V2 s)
{
    const int n = s.n + 2 * vec_dot<IV, IH, ID>(s.w) + (IV + IH + ID);
    const bool take = n < best.n;
    best.n = take ? n : best.n; best.w = take ? s.w + (unsigned)(IV | (IH << 10) | (ID << 20)) : best.w;
}
// fold an already-incremented candidate (first minimum in order)
__device__ __forceinline__ void fold(V2& best, const V2 c)
{
    const bool take = c.n < best.n;
    best.n = take ? c.n : best.n; best.w = take ? c.w : best.w;
}

// Row-scan kinds (mask functions of jly_3ddt.cpp and where the running-scan entry sits in the
// tie order): F1 = MINforwardDE1 (:502-706), F3 = MINforwardDE3 (:51-131), B1 = MINbackwardDE1
// (:295-500), B3 = MINbackwardDE3 (:171-252), C_UP = MINforwardDE4 (:133-169, scan z ascending),
// C_DN = MINforwardDE2 (:254-293, scan z descending).  The two "quirk" entries of DE3/backwardDE1
// (read (z+1,y) but add (0,1,1)) can never win: the entry just before them offers the same source
// voxel with a strictly smaller increment; they are omitted.
enum ScanKind { F1 = 0, F3 = 1, B1 = 2, B3 = 3, C_UP = 4, C_DN = 5 };

extern __shared__ __align__(16) unsigned char dt_smem[];

#ifdef GOICP_DT_INSTRUMENT
__device__ unsigned long long g_dt_stats_buf[16];
#define DT_STAT(i, v) atomicAdd(&g_dt_stats_buf[i], (unsigned long long)(v))
#else
#define DT_STAT(i, v) ((void)0)
#endif

// One scan resolved inside a warp (DIR = +1: the predecessor is lane-1), state carried through two
// shuffles.  The lane at the upstream end has no predecessor in the warp and keeps its own value
// -- the speculation.
template <int DIR>
__device__ __forceinline__ V2 warp_resolve(const V2 own, const int T, const int lane)
{
    const unsigned full = 0xffffffffu;
    const bool has_pred = DIR > 0 ? lane > 0 : lane < 31;
    V2 st = own;
    while (true) {
        const unsigned ow = DIR > 0 ? __shfl_up_sync(full, st.w, 1) : __shfl_down_sync(full, st.w, 1);
        const int on = DIR > 0 ? __shfl_up_sync(full, st.n, 1) : __shfl_down_sync(full, st.n, 1);
        const int nc = on + 2 * (int)(ow >> 20) + 1;         // an unset predecessor gives nc > kInf >= T
        const bool take = has_pred & (nc < T);
        const unsigned nw = take ? ow + (1u << 20) : own.w;
        const bool changed = nw != st.w;
        st.w = nw; st.n = take ? nc : own.n;
        if (!__any_sync(full, changed)) break;
    }
    return st;
}

// lane / voxel bookkeeping of a row-recurrence thread
struct Lane {
    int lane, z, zc;
    bool valid, owned, check_up, check_dn;
    __device__ __forceinline__ Lane(int warp, int lane_, int S) : lane(lane_)
    {
        z = warp * kOwned + lane - 1;
        valid = z >= 0 && z < S;                              // a voxel of the row (owned or shadow)
        owned = valid && lane >= 1 && lane <= kOwned;
        zc = min(max(z, -1), S) + 2;                          // padded index; lanes outside the row read the pads
        // the owner of a warp's downstream boundary voxel validates the neighbour warp's speculation
        check_up = owned && lane == kOwned && z + 1 < S;      // scans with DIR = +1
        check_dn = owned && lane == 1 && z > 0;               // scans with DIR = -1
    }
};

// Both scans of one row for this lane's voxel, warp-local.  pL/pC/pR = final previous row of the
// pass at z-1, z, z+1; p9 = folded candidates of the adjacent slice; self = the voxel's value so far.
// bad: bit 0 = the speculation of scan K1 failed at this (boundary) voxel, bit 1 = that of scan K2.
template <int K1, int K2>
__device__ __forceinline__ void row_recurrence(const Lane& ln, const V2 pL, const V2 pC, const V2 pR, const V2 p9, const V2 self,
                                               V2& own1, int& T1, V2& st1, V2& fin, int& bad)
{
    constexpr bool HAS_XS = (K1 == F1 || K1 == B1);
    constexpr int DIR1 = (K1 == F1 || K1 == B3) ? +1 : -1;                       // +1: recurrence reads z-1
    constexpr int DIR2 = (K2 == C_UP) ? +1 : -1;
    static_assert(DIR1 == -DIR2, "the two scans of a row run in opposite directions");
    V2 P = v2_unset(), Q = v2_unset();
    if (HAS_XS) P = p9;
    if (K1 == F1 || K1 == B3) {                  // previous row, then self; recurrence comes last
        V2 P2 = v2_unset();
        consider<0, 1, 1>(P, pL); consider<0, 1, 0>(P, pC);
        consider<0, 1, 1>(P2, pR); consider<0, 0, 0>(P2, self);
        fold(P, P2);
    } else {                                     // recurrence first, then (z,y+-1), self, (z-1,y+-1)
        consider<0, 1, 0>(Q, pC); consider<0, 0, 0>(Q, self); consider<0, 1, 1>(Q, pL);
    }
    own1 = P.n <= Q.n ? P : Q;
    T1 = ln.valid ? min(P.n, Q.n + 1) : 0;       // recurrence wins iff nc < nP and nc <= nQ; nothing passes a lane outside the row
    // chain scan K2: its input is the state after K1, its only other candidate the voxel itself.
    // First step of both scans at once, K2 under the assumption that K1 changes nothing: if no voxel
    // of the warp takes its predecessor's candidate in either scan, the row keeps its own values.
    const unsigned full = 0xffffffffu;
    const unsigned up_w = __shfl_up_sync(full, own1.w, 1), dn_w = __shfl_down_sync(full, own1.w, 1);
    const int up_n = __shfl_up_sync(full, own1.n, 1), dn_n = __shfl_down_sync(full, own1.n, 1);
    const int T2 = ln.valid ? min(kInf, own1.n + 1) : 0;
    const int nc_up = up_n + 2 * (int)(up_w >> 20) + 1, nc_dn = dn_n + 2 * (int)(dn_w >> 20) + 1;
    const bool ch_up = (ln.lane > 0) & (nc_up < (DIR1 > 0 ? T1 : T2)) & (up_w + (1u << 20) != own1.w);
    const bool ch_dn = (ln.lane < 31) & (nc_dn < (DIR1 > 0 ? T2 : T1)) & (dn_w + (1u << 20) != own1.w);
    const bool any1 = __any_sync(full, DIR1 > 0 ? ch_up : ch_dn);
    const bool any2 = __any_sync(full, DIR1 > 0 ? ch_dn : ch_up);
    st1 = own1; fin = own1; bad = 0;
    if (any1 | any2) {
        if (any1) st1 = warp_resolve<DIR1>(own1, T1, ln.lane);
        fin = warp_resolve<DIR2>(st1, ln.valid ? min(kInf, st1.n + 1) : 0, ln.lane);
        bad = (((DIR1 > 0 ? ln.check_up : ln.check_dn) && st1.w != own1.w) ? 1 : 0) | (((DIR2 > 0 ? ln.check_up : ln.check_dn) && fin.w != st1.w) ? 2 : 0);
    }
}

__device__ __forceinline__ void bar_all() { asm volatile("bar.sync 0;" ::: "memory"); }
__device__ __forceinline__ void bar_group(int n) { asm volatile("bar.sync 1, %0;" :: "r"(n) : "memory"); }
__device__ __forceinline__ bool bar_group_or(int n, bool pred)
{
    unsigned r;
    asm volatile("{ .reg .pred p, q; setp.ne.u32 q, %1, 0; bar.red.or.pred p, 1, %2, q; selp.u32 %0, 1, 0, p; }"
                 : "=r"(r) : "r"((unsigned)pred), "r"(n) : "memory");
    return r != 0;
}

// Redo one scan of a row whose speculation failed: exchange the states of the warps' downstream
// boundary voxels and re-resolve inside the warps with the upstream shadow lane pinned to its
// voxel's true state, until no boundary state moves (one round per warp boundary a run crosses).
// Collective over `nthreads` threads on named barrier 1.
template <int DIR>
__device__ __noinline__ V2 boundary_resolve(V2* xchg, const Lane& ln, const V2 own, const int T, V2 st, const int nthreads)
{
    const bool is_out = ln.owned && (DIR > 0 ? ln.lane == kOwned : ln.lane == 1);
    const bool is_in = ln.valid && (DIR > 0 ? ln.lane == 0 : ln.lane == 31);
    V2 in_state = own;                                       // what the warp assumed about its upstream shadow
    while (true) {
        if (is_out) xchg[ln.z + 2] = st;
        bar_group(nthreads);
        bool changed = false;
        if (is_in) { const V2 t = xchg[ln.z + 2]; changed = t.w != in_state.w; in_state = t; }
        if (!bar_group_or(nthreads, changed)) break;
        st = warp_resolve<DIR>(is_in ? in_state : own, T, ln.lane);
    }
    return st;
}
// both scans redone; `which`: bit 0 = scan K1 failed somewhere, bit 1 = scan K2 failed somewhere
template <int DIR1>
__device__ __forceinline__ V2 settle_row(V2* xchg, const Lane& ln, const int which, const V2 own1, const int T1, const V2 st1w, const V2 fin_w, const int nthreads)
{
    V2 st1 = st1w, fin = fin_w;
    if (which & 1) {
        st1 = boundary_resolve<DIR1>(xchg, ln, own1, T1, st1w, nthreads);
        // the shadow lanes' copies of the neighbours' scan-K1 states feed scan K2: refresh them
        if (ln.owned) xchg[ln.z + 2] = st1;
        bar_group(nthreads);
        if (ln.valid) st1 = xchg[ln.z + 2];
        bar_group(nthreads);
        fin = warp_resolve<-DIR1>(st1, ln.valid ? min(kInf, st1.n + 1) : 0, ln.lane);
    }
    return boundary_resolve<-DIR1>(xchg, ln, st1, ln.valid ? min(kInf, st1.n + 1) : 0, fin, nthreads);
}

// ---- variant 1: every warp does everything (any S <= 960) -------------------------------------
// State the row recurrence needs lives in registers: the final previous row at z-1, z, z+1 and
// the folded candidates of the adjacent slice (three rows x three z-neighbours, folded per row
// into A (row offset +-1) and B (same row) when a row enters the window, one row ahead of its
// use).  Every global access is a thread's own column, prefetched one row ahead.
// Row buffers: voxel z sits at index z+2, two "unset" pads on each side (a shadow lane reads the
// neighbour of a voxel one outside the row).  `fin` and `xrow` alternate by row parity: a row is
// written before the row's barrier and read after it, the next row writes the other copy.
struct RowSmem {
    int stride;          // voxels of one padded row
    __device__ __forceinline__ V2* base() const { return reinterpret_cast<V2*>(dt_smem); }
    __device__ __forceinline__ V2* fin(int par) const { return base() + par * stride; }
    __device__ __forceinline__ V2* xrow(int par) const { return base() + (2 + par) * stride; }
    __device__ __forceinline__ V2* xchg() const { return base() + 4 * stride; }
    __device__ __forceinline__ volatile int* flag() const { return reinterpret_cast<volatile int*>(base() + 5 * stride); }
};
size_t dt_propagate_smem(int S) { return 5 * (size_t)(S + 6) * sizeof(V2) + 16; }

// one pass over the rows of slice x: scan K1 then chain scan K2 on every row, rows in direction YDIR
template <int K1, int K2, int YDIR>
__device__ void slice_pass(V2* G, int S, int x, int xs, const RowSmem& sh, const int serial)
{
    constexpr bool HAS_XS = (K1 == F1 || K1 == B1);
    constexpr int DIR1 = (K1 == F1 || K1 == B3) ? +1 : -1;
    const Lane ln(threadIdx.x >> 5, threadIdx.x & 31, S);
    const int z = ln.z, P = pitch_of(S);
    const bool use_xs = HAS_XS && xs >= 0 && xs < S;
    const int y0 = YDIR > 0 ? 0 : S - 1;
    auto gload = [&](int xx, int yy) -> V2 { return (ln.valid && yy >= 0 && yy < S) ? G[((size_t)xx * S + yy) * P + z] : v2_unset(); };
    // folded candidates of one adjacent-slice row for column z: A = row offset +-1, B = same row
    auto fold_xrow = [&](const V2* xr, V2& A, V2& B) {
        A = v2_unset(); B = v2_unset();
        const V2 l = xr[ln.zc - 1], c = xr[ln.zc], r = xr[ln.zc + 1];
        consider<1, 1, 1>(A, l); consider<1, 1, 0>(A, c); consider<1, 1, 1>(A, r);
        consider<1, 0, 1>(B, l); consider<1, 0, 0>(B, c); consider<1, 0, 1>(B, r);
    };
    // adjacent-slice window: rows y-YDIR ("behind", outside the grid at the first row), y, y+YDIR
    V2 A_behind = v2_unset(), A_cur = v2_unset(), B_cur = v2_unset(), A_ahead = v2_unset(), B_ahead = v2_unset();
    if (use_xs) {
        if (ln.owned) { sh.xrow(0)[z + 2] = gload(xs, y0); sh.xrow(1)[z + 2] = gload(xs, y0 + YDIR); }
        __syncthreads();
        fold_xrow(sh.xrow(0), A_cur, B_cur);
        fold_xrow(sh.xrow(1), A_ahead, B_ahead);
        __syncthreads();
    }
    V2 self_next = gload(x, y0);
    V2 xs_next = (use_xs && ln.owned) ? gload(xs, y0 + 2 * YDIR) : v2_unset();
    V2 pL = v2_unset(), pC = v2_unset(), pR = v2_unset();    // final previous row of this pass at z-1, z, z+1
    for (int i = 0, y = y0; i < S; i++, y += YDIR) {
        const int par = i & 1;
        const V2 self = self_next;
        const V2 xs_new = xs_next;                          // adjacent-slice row y + 2*YDIR: enters the window at the next row
        self_next = gload(x, y + YDIR);                     // own column only
        xs_next = (use_xs && ln.owned) ? gload(xs, y + 3 * YDIR) : v2_unset();
        {   // the volume exceeds L2 at S=300: pull the rows needed some iterations from now into L2
            const int yf = y + 16 * YDIR;
            if (ln.owned && (z & 15) == 0 && yf >= 0 && yf < S) {
                asm volatile("prefetch.global.L2 [%0];" :: "l"(G + ((size_t)x * S + yf) * P + z));
                if (use_xs) asm volatile("prefetch.global.L2 [%0];" :: "l"(G + ((size_t)xs * S + yf) * P + z));
            }
        }
        V2 p9 = v2_unset();
        if (HAS_XS) {                                       // mask order: row y-1, row y, row y+1 of the adjacent slice
            if (YDIR > 0) { fold(p9, A_behind); fold(p9, B_cur); fold(p9, A_ahead); }
            else          { fold(p9, A_ahead); fold(p9, B_cur); fold(p9, A_behind); }
        }
        V2 own1, st1w, fin; int T1, bad;
        row_recurrence<K1, K2>(ln, pL, pC, pR, p9, self, own1, T1, st1w, fin, bad);
        V2* fin_row = sh.fin(par);
        if (ln.owned) { fin_row[z + 2] = fin; if (use_xs) sh.xrow(par)[z + 2] = xs_new; }
        if (bad) { if (bad & 1) sh.flag()[par] = serial + i; else sh.flag()[2 + par] = serial + i; }
        __syncthreads();
        const int which = (sh.flag()[par] == serial + i ? 1 : 0) | (sh.flag()[2 + par] == serial + i ? 2 : 0);
        DT_STAT(4, threadIdx.x == 0); DT_STAT(5, threadIdx.x == 0 && which);
        if (which) {                                 // a run crossed a warp boundary: settle the boundaries across the CTA
            fin = settle_row<DIR1>(sh.xchg(), ln, which, own1, T1, st1w, fin, blockDim.x);
            if (ln.owned) fin_row[z + 2] = fin;
            __syncthreads();
        }
        if (ln.owned) G[((size_t)x * S + y) * P + z] = fin;
        pL = fin_row[ln.zc - 1]; pC = fin_row[ln.zc]; pR = fin_row[ln.zc + 1];
        if (HAS_XS) {                                // slide the adjacent-slice window
            A_behind = A_cur; A_cur = A_ahead; B_cur = B_ahead;
            if (use_xs) fold_xrow(sh.xrow(par), A_ahead, B_ahead); else { A_ahead = v2_unset(); B_ahead = v2_unset(); }
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(1024)
dt_propagate_kernel(V2* G, int S)
{
    RowSmem sh;
    sh.stride = S + 6;
    // pads of the row buffers stay "unset" for the whole kernel
    for (int i = threadIdx.x; i < 5 * (S + 6); i += blockDim.x) sh.base()[i] = v2_unset();
    if (threadIdx.x < 4) sh.flag()[threadIdx.x] = -1;
    __syncthreads();
    int serial = 0;                                                 // rows get unique ids across passes
    for (int x = 0; x < S; x++) {                                   // jly_3ddt.cpp:719-728
        slice_pass<F1, C_DN, +1>(G, S, x, x - 1, sh, serial); serial += 1024;
        slice_pass<F3, C_UP, -1>(G, S, x, -1, sh, serial); serial += 1024;
    }
    for (int x = S - 1; x >= 0; x--) {                              // :729-739
        slice_pass<B1, C_UP, -1>(G, S, x, x + 1, sh, serial); serial += 1024;
        slice_pass<B3, C_DN, +1>(G, S, x, -1, sh, serial); serial += 1024;
    }
}

// ---- variant 2: warp-specialised, rows moved by the TMA (S <= 640) ----------------------------
// Everything that does not depend on the row recurrence is taken off its critical path.  The CTA
// is split into
//   CONSUMER warps   the recurrence proper: own/T from the previous row, the two scans, the
//                    speculation check;
//   PRODUCER warps   one row ahead: fold the 3x3 candidates of the adjacent slice into one
//                    candidate per voxel (p9), and write final rows back to HBM two rows behind;
//   one COPY warp    a single lane streams rows HBM -> shared memory with 1-D bulk copies
//                    (cp.async.bulk + mbarrier complete_tx) into two rings, R rows deep: the rows
//                    of the slice being updated (read by the consumers as "self") and the rows of
//                    the adjacent slice (read by the producers).  Nobody computes a global address
//                    or converts a format per voxel; the copy lane absorbs all HBM/L2 latency.
// One CTA-wide barrier per row step is the only synchronisation (the copy lane waits for the bulk
// copies a step before their rows are used, the barrier publishes them); settling a row whose
// speculation failed uses a named barrier among the consumer warps only.
constexpr int kSplitMaxS = 640;
constexpr int kRing = 8;               // rows of lookahead of the bulk copies

struct SplitSmem {
    int stride;          // voxels of one padded row (pitch + 4), voxel z at index z+2; rows are 16-byte aligned
    __device__ __forceinline__ V2* base() const { return reinterpret_cast<V2*>(dt_smem); }
    __device__ __forceinline__ V2* fin(int k) const { return base() + k * stride; }                // final rows, ring of 4
    __device__ __forceinline__ V2* p9(int k) const { return base() + (4 + k) * stride; }           // folded adjacent-slice candidate, by row parity
    __device__ __forceinline__ V2* xchg() const { return base() + 6 * stride; }                    // boundary exchange (consumers only)
    __device__ __forceinline__ V2* selfv(int k) const { return base() + (7 + k) * stride; }        // ring: rows of the slice before this pass
    __device__ __forceinline__ V2* xrow(int k) const { return base() + (7 + kRing + k) * stride; } // ring: rows of the adjacent slice
    __device__ __forceinline__ unsigned char* tail() const { return reinterpret_cast<unsigned char*>(base() + (7 + 2 * kRing) * stride); }
    __device__ __forceinline__ volatile int* flag() const { return reinterpret_cast<volatile int*>(tail()); }
    __device__ __forceinline__ unsigned full_s(int k) const { return (unsigned)__cvta_generic_to_shared(tail() + 16 + 8 * k); }
    __device__ __forceinline__ unsigned full_x(int k) const { return (unsigned)__cvta_generic_to_shared(tail() + 16 + 8 * (kRing + k)); }
};
size_t dt_split_smem(int S) { return (size_t)(7 + 2 * kRing) * (((S + 1) & ~1) + 4) * sizeof(V2) + 16 + 16 * (size_t)kRing; }

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
    asm volatile("{\n.reg .pred P1;\nDT_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DT_DONE;\nbra DT_WAIT;\nDT_DONE:\n}"
                 :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(unsigned dst, const void* src, unsigned bytes, unsigned bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_async_proxy() { asm volatile("fence.proxy.async;" ::: "memory"); }

// Row steps s = -2 .. S+1 of one pass (rows indexed along the pass: y = y0 + i*YDIR; rows outside
// [0,S) are "unset" / skipped).  At step s
//   consumers  resolve row s                      (reads self row s, p9 of row s, final row s-1)
//   producers  fold adjacent row s+2 -> p9 of row s+1; store final row s-2
//   copy lane  refills the ring slots freed at step s-1, then waits for self row s+1 and adjacent
//              row s+3, which are first read at step s+1.
template <int K1, int K2, int YDIR>
__device__ __forceinline__ void consumer_pass(int S, const SplitSmem& sh, const int ncons_threads, const int serial)
{
    constexpr int DIR1 = (K1 == F1 || K1 == B3) ? +1 : -1;
    const Lane ln(threadIdx.x >> 5, threadIdx.x & 31, S);
    const int z = ln.z;
    if (ln.owned) sh.fin(3)[z + 2] = v2_unset();             // "row -1" of this pass
    bar_all();                                               // prologue
    bar_all(); bar_all();                                    // steps -2, -1: the producers and the copy lane fill the pipeline
    // inputs of the coming step, loaded right after the barrier that publishes them
    V2 pL = v2_unset(), pC = pL, pR = pL;
    V2 self = sh.selfv(0)[ln.zc], p9 = sh.p9(0)[ln.zc];
#ifdef GOICP_DT_INSTRUMENT
    long long a_work = 0, a_bar = 0, a_post = 0, a_rows = 0, a_settled = 0; long long c_prev = clock64();
#endif
    for (int s = 0; s < S; s++) {
#ifdef GOICP_DT_INSTRUMENT
        const long long c_w0 = clock64(); a_post += c_w0 - c_prev;
#endif
        V2 own1, st1w, fin_w; int T1, bad;
        row_recurrence<K1, K2>(ln, pL, pC, pR, p9, self, own1, T1, st1w, fin_w, bad);
        V2* cur = sh.fin(s & 3);
        if (ln.owned) cur[z + 2] = fin_w;
        if (bad) { if (bad & 1) sh.flag()[s & 1] = serial + s; else sh.flag()[2 + (s & 1)] = serial + s; }
#ifdef GOICP_DT_INSTRUMENT
        const long long c_w1 = clock64();
        bar_all();
        c_prev = clock64(); a_work += c_w1 - c_w0; a_bar += c_prev - c_w1;
#else
        bar_all();
#endif
        const int flag1 = sh.flag()[s & 1], flag2 = sh.flag()[2 + (s & 1)];
        pL = cur[ln.zc - 1]; pC = cur[ln.zc]; pR = cur[ln.zc + 1];
        self = sh.selfv((s + 1) & (kRing - 1))[ln.zc];
        p9 = sh.p9((s + 1) & 1)[ln.zc];
        const int which = (flag1 == serial + s ? 1 : 0) | (flag2 == serial + s ? 2 : 0);
#ifdef GOICP_DT_INSTRUMENT
        a_rows++; a_settled += which ? 1 : 0;
#endif
        if (which) {                                 // a run crossed a warp boundary: settle the boundaries among the consumers
            const V2 fin = settle_row<DIR1>(sh.xchg(), ln, which, own1, T1, st1w, fin_w, ncons_threads);
            if (ln.owned) cur[z + 2] = fin;
            bar_group(ncons_threads);
            pL = cur[ln.zc - 1]; pC = cur[ln.zc]; pR = cur[ln.zc + 1];
        }
    }
#ifdef GOICP_DT_INSTRUMENT
    if (threadIdx.x == 0) { DT_STAT(4, a_rows); DT_STAT(5, a_settled); DT_STAT(6, a_work); DT_STAT(7, a_bar); DT_STAT(8, a_post); }
#endif
    bar_all(); bar_all();                                    // steps S, S+1: the producers store the last rows
}

template <int K1, int YDIR, int VPT>
__device__ __forceinline__ void producer_pass(V2* G, int S, int x, int xs, const SplitSmem& sh, const int ptid, const int nprod_threads)
{
    constexpr bool HAS_XS = (K1 == F1 || K1 == B1);
    const bool use_xs = HAS_XS && xs >= 0 && xs < S;
    const int y0 = YDIR > 0 ? 0 : S - 1, P = pitch_of(S);
    V2 A_prev[VPT], X_prev[VPT];
#pragma unroll
    for (int j = 0; j < VPT; j++) {
        A_prev[j] = v2_unset(); X_prev[j] = v2_unset();
        const int z = ptid + j * nprod_threads;
        if (HAS_XS && !use_xs && z < S) { sh.p9(0)[z + 2] = v2_unset(); sh.p9(1)[z + 2] = v2_unset(); }   // no adjacent slice: nothing to fold
    }
    bar_all();
#ifdef GOICP_DT_INSTRUMENT
    long long b_work = 0, b_bar = 0;
#endif
    for (int s = -2; s < S + 2; s++) {
#ifdef GOICP_DT_INSTRUMENT
        const long long c_p0 = clock64();
#endif
#pragma unroll
        for (int j = 0; j < VPT; j++) {
            const int z = ptid + j * nprod_threads;
            if (z < S) {
                if (use_xs && s + 1 < S) {
                    V2 A = v2_unset(), B = v2_unset();
                    if (s + 2 < S) {                                              // row S is outside the grid
                        const V2* xr = sh.xrow((s + 2) & (kRing - 1));
                        const V2 l = xr[z + 1], c = xr[z + 2], r = xr[z + 3];
                        consider<1, 1, 1>(A, l); consider<1, 1, 0>(A, c); consider<1, 1, 1>(A, r);
                        consider<1, 0, 1>(B, l); consider<1, 0, 0>(B, c); consider<1, 0, 1>(B, r);
                    }
                    // mask order of row s+1: adjacent rows y-1, y, y+1 = (s, s+1, s+2) for YDIR>0, reversed otherwise
                    V2 p9, X;
                    if (YDIR > 0) { p9 = X_prev[j]; fold(p9, A); X = A_prev[j]; fold(X, B); }
                    else          { p9 = A; fold(p9, X_prev[j]); X = B; fold(X, A_prev[j]); }
                    if (s + 1 >= 0) sh.p9((s + 1) & 1)[z + 2] = p9;
                    X_prev[j] = X; A_prev[j] = A;
                }
                if (s - 2 >= 0 && s - 2 < S) G[((size_t)x * S + (y0 + (s - 2) * YDIR)) * P + z] = sh.fin((s - 2) & 3)[z + 2];
            }
        }
        if (s == S + 1) fence_async_proxy();         // the rows just written are bulk-read by the next pass
#ifdef GOICP_DT_INSTRUMENT
        const long long c_p1 = clock64();
        bar_all();
        if (s >= 0 && s < S) { b_work += c_p1 - c_p0; b_bar += clock64() - c_p1; }
#else
        bar_all();
#endif
    }
#ifdef GOICP_DT_INSTRUMENT
    if (ptid == 0) { DT_STAT(2, b_work); DT_STAT(3, b_bar); }
#endif
}

template <int K1, int YDIR>
__device__ __forceinline__ void copy_pass(const V2* G, int S, int x, int xs, const SplitSmem& sh, unsigned& ph_s, unsigned& ph_x)
{
    constexpr bool HAS_XS = (K1 == F1 || K1 == B1);
    constexpr int R = kRing, Rm = kRing - 1;
    const bool use_xs = HAS_XS && xs >= 0 && xs < S;
    const int y0 = YDIR > 0 ? 0 : S - 1, P = pitch_of(S);
    const unsigned row_bytes = (unsigned)P * (unsigned)sizeof(V2);           // the pad voxel of an odd row is "unset" and lands on a pad
    const bool lead = (threadIdx.x & 31) == 0;
    auto issue_self = [&](int r) {
        if (r < 0 || r >= S) return;
        const unsigned bar = sh.full_s(r & Rm);
        mbar_expect_tx(bar, row_bytes);
        bulk_load((unsigned)__cvta_generic_to_shared(sh.selfv(r & Rm) + 2), G + ((size_t)x * S + (y0 + r * YDIR)) * P, row_bytes, bar);
    };
    auto issue_x = [&](int r) {
        if (!use_xs || r < 0 || r >= S) return;
        const unsigned bar = sh.full_x(r & Rm);
        mbar_expect_tx(bar, row_bytes);
        bulk_load((unsigned)__cvta_generic_to_shared(sh.xrow(r & Rm) + 2), G + ((size_t)xs * S + (y0 + r * YDIR)) * P, row_bytes, bar);
    };
    auto wait_self = [&](int r) {
        if (r < 0 || r >= S) return;
        mbar_wait(sh.full_s(r & Rm), (ph_s >> (r & Rm)) & 1u); ph_s ^= 1u << (r & Rm);
    };
    auto wait_x = [&](int r) {
        if (!use_xs || r < 0 || r >= S) return;
        mbar_wait(sh.full_x(r & Rm), (ph_x >> (r & Rm)) & 1u); ph_x ^= 1u << (r & Rm);
    };
    if (lead) {
        fence_async_proxy();
        for (int r = 0; r <= R - 4; r++) issue_self(r);      // step s issues self row s-1+R and adjacent row s+1+R
        for (int r = 0; r <= R - 2; r++) issue_x(r);
        wait_x(0);                                           // folded at step -2
    }
    __syncwarp();
    bar_all();
#ifdef GOICP_DT_INSTRUMENT
    long long k_work = 0, k_bar = 0;
#endif
    for (int s = -2; s < S + 2; s++) {
#ifdef GOICP_DT_INSTRUMENT
        const long long c_k0 = clock64();
#endif
        if (lead) {
            issue_self(s - 1 + R); issue_x(s + 1 + R);
            wait_self(s + 1); wait_x(s + 3);
        }
        __syncwarp();
#ifdef GOICP_DT_INSTRUMENT
        const long long c_k1 = clock64();
        bar_all();
        if (s >= 0 && s < S) { k_work += c_k1 - c_k0; k_bar += clock64() - c_k1; }
#else
        bar_all();
#endif
    }
#ifdef GOICP_DT_INSTRUMENT
    if (lead) { DT_STAT(9, k_work); DT_STAT(10, k_bar); }
#endif
}

template <int VPT, int MAXT>
__global__ void __launch_bounds__(MAXT)
dt_propagate_split_kernel(V2* G, int S, int ncons_warps, int nprod_warps)
{
    SplitSmem sh;
    sh.stride = pitch_of(S) + 4;
    // pads of the row buffers stay "unset" for the whole kernel
    for (int i = threadIdx.x; i < (7 + 2 * kRing) * sh.stride; i += blockDim.x) sh.base()[i] = v2_unset();
    if (threadIdx.x < 4) sh.flag()[threadIdx.x] = -1;
    if (threadIdx.x == 0) {
        for (int k = 0; k < kRing; k++) { mbar_init(sh.full_s(k), 1); mbar_init(sh.full_x(k), 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_async_proxy();
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    const int role = warp < ncons_warps ? 0 : (warp < ncons_warps + nprod_warps ? 1 : 2);
    const int nct = ncons_warps * 32;
    const int ptid = (int)threadIdx.x - nct, npt = nprod_warps * 32;
    int serial = 0;                                                // row steps get unique ids across passes
    unsigned ph_s = 0, ph_x = 0;                                   // mbarrier phase bits per ring slot (copy lane)
#define DT_PASS(K1, K2, YD, XX, XS) do { \
        if (role == 0) consumer_pass<K1, K2, YD>(S, sh, nct, serial); \
        else if (role == 1) producer_pass<K1, YD, VPT>(G, S, XX, XS, sh, ptid, npt); \
        else copy_pass<K1, YD>(G, S, XX, XS, sh, ph_s, ph_x); \
        serial += 1024; } while (0)
    for (int x = 0; x < S; x++) {                                   // jly_3ddt.cpp:719-728
        DT_PASS(F1, C_DN, +1, x, x - 1);
        DT_PASS(F3, C_UP, -1, x, -1);
    }
    for (int x = S - 1; x >= 0; x--) {                              // :729-739
        DT_PASS(B1, C_UP, -1, x, x + 1);
        DT_PASS(B3, C_DN, +1, x, -1);
    }
#undef DT_PASS
}

// distance = float( double(float(sqrt(double(n2)))) / scale ), clamped at 0 (jly_3ddt.cpp:970-978);
// also transposes the working [x][y][z] layout into the reference's [z][y][x].
__global__ void dt_finalize_kernel(const V2* __restrict__ G, int S, int P, double scale, float* __restrict__ out)
{
    __shared__ float tile[32][33];
    const int y = blockIdx.z;
    const int x0 = blockIdx.x * 32, z0 = blockIdx.y * 32;
    {
        const int x = x0 + threadIdx.y, z = z0 + threadIdx.x;
        if (x < S && z < S) {
            const V2 a = G[((size_t)x * S + y) * P + z];
            const double n2 = (double)a.n;                   // = v^2 + h^2 + d^2 exactly, as the reference's double sum
            float dv = __double2float_rn(sqrt(n2));
            float r = __double2float_rn(__ddiv_rn((double)dv, scale));
            tile[threadIdx.y][threadIdx.x] = r < 0.0f ? 0.0f : r;
        }
    }
    __syncthreads();
    {
        const int x = x0 + threadIdx.x, z = z0 + threadIdx.y;
        if (x < S && z < S) out[((size_t)z * S + y) * S + x] = tile[threadIdx.x][threadIdx.y];
    }
}

// ---- exact EDT (separable, integer squared distances; Meijster et al.'s two-scan lower envelope) -------------
// Works IN PLACE on the output grid, in the output's own [z][y][x] layout (x fastest), as int32 squared distances
// until the last pass converts to the reference's float metric distance -- no working volume, no transposition:
//   pass x   one warp per (z,y) line: nearest seed to the left / right by ballots                 (contiguous)
//   pass y   one thread per (z,x) line, threads of a warp = neighbouring x                          (coalesced)
//   pass z   one thread per (y,x) line, same, and writes float(double(float(sqrt(n2))) / scale)     (coalesced)
// A line pass is D'(u) = min_i D(i) + (u-i)^2, the lower envelope of S parabolas: the forward scan keeps the
// parabolas that own a stretch of the line on a stack (s = apex, t = first owned position, gs = D at the apex;
// integer Sep() = floor of the crossing point, exact), the backward scan evaluates.  O(S) per line, S^2 lines.
constexpr int kEdtInf = 0x3f3f3f3f;    // "no seed on this line yet"; byte-uniform so that cudaMemset can write it, and
                                       // kEdtInf + 1023^2 stays far below 2^31

__global__ void edt_seed_kernel(int* __restrict__ D, int S, const float* __restrict__ model, int nm, double xmin, double ymin, double zmin, double scale, int corner_seed)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0 && corner_seed) D[0] = 0;                       // see dt_init_kernel
    if (i >= nm) return;
    int x = __double2int_rz(__dadd_rn(__dmul_rn(__dsub_rn((double)model[3 * i], xmin), scale), 0.5));
    int y = __double2int_rz(__dadd_rn(__dmul_rn(__dsub_rn((double)model[3 * i + 1], ymin), scale), 0.5));
    int z = __double2int_rz(__dadd_rn(__dmul_rn(__dsub_rn((double)model[3 * i + 2], zmin), scale), 0.5));
    if (x < 0 || x >= S || y < 0 || y >= S || z < 0 || z >= S) return;
    D[((size_t)z * S + y) * S + x] = 0;
}

__global__ void __launch_bounds__(256) edt_pass_x(int* __restrict__ D, int S)
{
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int nw = gridDim.x * (blockDim.x >> 5);
    for (int line = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); line < S * S; line += nw) {
        int* d = D + (size_t)line * S;
        const int far = 1 << 20;
        int last = -far;                                        // nearest seed at or left of x, chunk by chunk
        for (int x0 = 0; x0 < S; x0 += 32) {
            const int x = x0 + lane;
            const bool seed = x < S && d[x] == 0;
            const unsigned m = __ballot_sync(full, seed);
            const unsigned upto = m & (0xffffffffu >> (31 - lane));
            const int left = upto ? x0 + 31 - __clz(upto) : last;
            if (x < S) d[x] = x - left;                         // >= far when there is none
            if (m) last = x0 + 31 - __clz(m);
        }
        int next = far + far;                                   // nearest seed at or right of x
        for (int x0 = ((S - 1) >> 5) << 5; x0 >= 0; x0 -= 32) {
            const int x = x0 + lane;
            const int dl = x < S ? d[x] : far;
            const unsigned m = __ballot_sync(full, dl == 0);
            const unsigned from = m & (0xffffffffu << lane);
            const int right = from ? x0 + __ffs(from) - 1 : next;
            if (x < S) { const int k = min(dl, right - x); d[x] = k >= far ? kEdtInf : k * k; }
            if (m) next = x0 + __ffs(m) - 1;
        }
    }
}

template <int MAXS, bool FINAL>
__global__ void __launch_bounds__(128) edt_pass_line(int* __restrict__ D, int S, size_t stride_elem, size_t stride_outer, double scale)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= S * S) return;
    const size_t base = (size_t)(tid / S) * stride_outer + (size_t)(tid % S);
    short s[MAXS], t[MAXS]; int gs[MAXS];
    int q = 0;
    s[0] = 0; t[0] = 0; gs[0] = D[base];
    int g_next = S > 1 ? D[base + stride_elem] : 0;
    for (int u = 1; u < S; u++) {
        const int gu = g_next;
        if (u + 1 < S) g_next = D[base + (size_t)(u + 1) * stride_elem];          // one element ahead of the dependent chain
        while (q >= 0) {
            const int a = t[q] - s[q], b = t[q] - u;
            if (a * a + gs[q] > b * b + gu) q--; else break;
        }
        if (q < 0) { q = 0; s[0] = (short)u; gs[0] = gu; }
        else {
            const int i = s[q];
            const int w = 1 + (u * u - i * i + gu - gs[q]) / (2 * (u - i));      // numerator >= 0 here: the parabola at i owns t[q] >= 0
            if (w < S) { q++; s[q] = (short)u; t[q] = (short)w; gs[q] = gu; }
        }
    }
    for (int u = S - 1; u >= 0; u--) {
        const int e = u - s[q];
        int v = e * e + gs[q];
        if (v > kEdtInf) v = kEdtInf;
        if (FINAL) {
            // distance = float( double(float(sqrt(double(n2)))) / scale ) (jly_3ddt.cpp:970-978)
            const float dv = __double2float_rn(sqrt((double)v));
            reinterpret_cast<float*>(D)[base + (size_t)u * stride_elem] = __double2float_rn(__ddiv_rn((double)dv, scale));
        } else D[base + (size_t)u * stride_elem] = v;
        if (u == t[q]) q--;
    }
}

} // namespace

void dt_frame_host(const float* m, int nm, int S, double expand, double* meta4)
{
    double xMin = m[0], xMax = m[0], yMin = m[1], yMax = m[1], zMin = m[2], zMax = m[2];
    for (int i = 1; i < nm; i++) {
        const double x = m[3 * i], y = m[3 * i + 1], z = m[3 * i + 2];
        if (xMin > x) xMin = x;
        if (xMax < x) xMax = x;
        if (yMin > y) yMin = y;
        if (yMax < y) yMax = y;
        if (zMin > z) zMin = z;
        if (zMax < z) zMax = z;
    }
    const double xc = (xMin + xMax) / 2, yc = (yMin + yMax) / 2, zc = (zMin + zMax) / 2;
    xMin = xc - expand * (xMax - xc); xMax = xc + expand * (xMax - xc);
    yMin = yc - expand * (yMax - yc); yMax = yc + expand * (yMax - yc);
    zMin = zc - expand * (zMax - zc); zMax = zc + expand * (zMax - zc);
    double side = xMax - xMin > yMax - yMin ? xMax - xMin : yMax - yMin;
    side = side > zMax - zMin ? side : zMax - zMin;
    meta4[0] = xc - side / 2; meta4[1] = yc - side / 2; meta4[2] = zc - side / 2;
    meta4[3] = S / side;
}

cudaError_t dt_build_device(const float* model, int nm, int S, double expand, int mode,
                            float* d_out, double* meta, cudaStream_t stream, std::string& msg)
{
    cudaError_t e;
    const bool trace = getenv("GOICP_TRACE_DT") != nullptr;
    auto tnow = []() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; };
    double tprev = tnow();
    auto mark = [&](const char* what) { if (trace) { const double t = tnow(); fprintf(stderr, "[dt trace] %-28s %.3f ms\n", what, 1e3 * (t - tprev)); tprev = t; } };
    if (S > kMaxS) { msg = "dt_size > 1024"; return cudaErrorInvalidValue; }
    if (mode == 0 && S > kMaxRefS) { msg = "dt_size > 960 in reference-order mode (use the exact EDT mode)"; return cudaErrorInvalidValue; }
    dt_frame_host(model, nm, S, expand, meta);
    const size_t n3 = (size_t)S * S * S;
    const int P = (S + 1) & ~1;
    const size_t ng = (size_t)S * S * P;
    V2* G = nullptr; float* d_model = nullptr;
    auto cleanup = [&]() { cudaStreamSynchronize(stream);      /* blocks go back to a shared pool: nothing may still use them */ pool_free(G); pool_free(d_model); };
#define DT_TRY(expr) do { e = (expr); if (e != cudaSuccess) { msg = #expr; cleanup(); return e; } } while (0)
    if (mode != 0) {
        // exact EDT, in place on the output grid (see edt_pass_line)
        DT_TRY(pool_alloc((void**)&d_model, (size_t)3 * nm * sizeof(float)));
        mark("frame + pool_alloc");
        DT_TRY(cudaMemcpyAsync(d_model, model, (size_t)3 * nm * sizeof(float), cudaMemcpyHostToDevice, stream));
        int* D = reinterpret_cast<int*>(d_out);
        DT_TRY(cudaMemsetAsync(D, 0x3f, n3 * sizeof(int), stream));
        edt_seed_kernel<<<(nm + 255) / 256, 256, 0, stream>>>(D, S, d_model, nm, meta[0], meta[1], meta[2], meta[3], mode == 2 ? 1 : 0);
        DT_TRY(cudaGetLastError());
        edt_pass_x<<<std::min((S * S + 7) / 8, 148 * 8), 256, 0, stream>>>(D, S);
        DT_TRY(cudaGetLastError());
        const unsigned nb = (unsigned)((S * S + 127) / 128);
#define DT_EDT_PASSES(MAXS) do { \
            edt_pass_line<MAXS, false><<<nb, 128, 0, stream>>>(D, S, (size_t)S, (size_t)S * S, meta[3]);      /* along y: lines (z, x) */ \
            DT_TRY(cudaGetLastError()); \
            edt_pass_line<MAXS, true><<<nb, 128, 0, stream>>>(D, S, (size_t)S * S, (size_t)S, meta[3]);       /* along z: lines (y, x) */ \
            DT_TRY(cudaGetLastError()); } while (0)
        if (S <= 320) DT_EDT_PASSES(320); else if (S <= 512) DT_EDT_PASSES(512); else DT_EDT_PASSES(1024);
#undef DT_EDT_PASSES
    } else {
    DT_TRY(pool_alloc((void**)&G, ng * sizeof(V2)));
    DT_TRY(pool_alloc((void**)&d_model, (size_t)3 * nm * sizeof(float)));
    mark("frame + pool_alloc");
    DT_TRY(cudaMemcpyAsync(d_model, model, (size_t)3 * nm * sizeof(float), cudaMemcpyHostToDevice, stream));
    dt_init_kernel<<<(unsigned)((ng + 255) / 256), 256, 0, stream>>>(G, ng, 1);
    DT_TRY(cudaGetLastError());
    dt_seed_kernel<<<(nm + 255) / 256, 256, 0, stream>>>(G, S, P, d_model, nm, meta[0], meta[1], meta[2], meta[3]);
    DT_TRY(cudaGetLastError());
    mark("memcpy + init + seed launch");
    {
        const bool timing = getenv("GOICP_DT_TIMING") != nullptr;
        cudaEvent_t ev0 = nullptr, ev1 = nullptr;
        if (timing) { cudaEventCreate(&ev0); cudaEventCreate(&ev1); cudaEventRecord(ev0, stream); }
        const int ncons = (S + kOwned - 1) / kOwned;
        if (S <= kSplitMaxS && getenv("GOICP_DT_UNSPLIT") == nullptr) {
            // producer warps: one voxel per thread while 32 warps suffice, else up to three; one copy warp
            int nprod = (S + 31) / 32;
            if (ncons + nprod + 1 > 32) nprod = 32 - 1 - ncons;
            const int vpt = (S + nprod * 32 - 1) / (nprod * 32);                // <= 3 for S <= kSplitMaxS
            const size_t smem = dt_split_smem(S);
            const int threads = (ncons + nprod + 1) * 32;
#define DT_LAUNCH_SPLIT(VPT, MAXT) do { \
                DT_TRY(cudaFuncSetAttribute(dt_propagate_split_kernel<VPT, MAXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
                dt_propagate_split_kernel<VPT, MAXT><<<1, threads, smem, stream>>>(G, S, ncons, nprod); } while (0)
            if (vpt == 1 && threads <= 704) DT_LAUNCH_SPLIT(1, 704);            // S <= 320: more registers per thread
            else if (vpt == 1) DT_LAUNCH_SPLIT(1, 1024);
            else if (vpt == 2) DT_LAUNCH_SPLIT(2, 1024);
            else DT_LAUNCH_SPLIT(3, 1024);
#undef DT_LAUNCH_SPLIT
        } else {
            const size_t smem = dt_propagate_smem(S);
            DT_TRY(cudaFuncSetAttribute(dt_propagate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            dt_propagate_kernel<<<1, ncons * 32, smem, stream>>>(G, S);
        }
        DT_TRY(cudaGetLastError());
        mark("propagate launch");
        if (timing) {
            cudaEventRecord(ev1, stream); cudaEventSynchronize(ev1);
            float ms = 0; cudaEventElapsedTime(&ms, ev0, ev1);
            fprintf(stderr, "[dt timing] S=%d propagate kernel %.3f ms\n", S, ms);
            cudaEventDestroy(ev0); cudaEventDestroy(ev1);
        }
#ifdef GOICP_DT_INSTRUMENT
        {
            unsigned long long hs[16];
            DT_TRY(cudaStreamSynchronize(stream));
            DT_TRY(cudaMemcpyFromSymbol(hs, g_dt_stats_buf, sizeof hs));
            fprintf(stderr, "[dt stats] rows %llu, settled across warps %llu; per row: consumer work %.0f barrier %.0f post-barrier %.0f | producer work %.0f barrier %.0f | copy lane issue+wait %.0f barrier %.0f cycles\n", hs[4], hs[5], (double)hs[6] / hs[4], (double)hs[7] / hs[4], (double)hs[8] / hs[4], (double)hs[2] / hs[4], (double)hs[3] / hs[4], (double)hs[9] / hs[4], (double)hs[10] / hs[4]);
        }
#endif
        dim3 grid((S + 31) / 32, (S + 31) / 32, S), block(32, 32);
        dt_finalize_kernel<<<grid, block, 0, stream>>>(G, S, P, meta[3], d_out);
        DT_TRY(cudaGetLastError());
    }
    }
    mark("finalize launch");
    DT_TRY(cudaStreamSynchronize(stream));
    mark("stream sync");
#undef DT_TRY
    cleanup();
    mark("cleanup");
    return cudaSuccess;
}

} // namespace goicp
