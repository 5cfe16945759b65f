// dt_kernels.cu -- distance-transform builders on sm_100a (GoICP::BuildDT, jly_goicp.cpp:75-90).
//
// GOICP_DT_REFERENCE reproduces DT3D::Build (jly_3ddt.cpp:889-979) bit for bit.  The reference
// transform is a *sequential* signed-less vector propagation ("DEuclidean", :710-742): every
// voxel update reads voxels updated earlier in the same raster scan, each x-slice is swept by
// rows forward then backward, each row along z forward then backward, and the whole volume in
// +x then -x.  Its output depends on that order (it is not the exact EDT), so the order is
// kept: one persistent CTA walks slices and rows in the reference order; within a row all the
// candidates that do not depend on the running scan are evaluated in parallel (one thread per
// voxel) and only the true recurrence -- "previous voxel of this scan + (0,0,1)" -- is resolved
// serially, and only for rows where it can matter.  Comparisons use the integer squared norm:
// the reference compares float(sqrt(v^2+h^2+d^2)), a strictly increasing function of that
// integer for grids up to 1024^3, with strict '<' in mask order (first minimum wins).
//
// GOICP_DT_EXACT_EDT is the separable exact squared-Euclidean transform (three 1-D lower-envelope
// passes over integer squared distances), fully parallel.
#include "dt_kernels.h"
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

namespace goicp {

namespace {

constexpr int kInf = 0x3fffffff;       // "no candidate" squared norm
constexpr int kMaxS = 1024;

struct __align__(8) Vox { short v, h, d, pad; };

__device__ __forceinline__ int vox_norm2(const Vox& a) { return (int)a.v * a.v + (int)a.h * a.h + (int)a.d * a.d; }
__device__ __forceinline__ bool vox_unset(const Vox& a) { return a.v == 32767; }

__global__ void dt_init_kernel(Vox* G, size_t n3)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n3) return;
    Vox u; u.v = u.h = u.d = 32767; u.pad = 0;
    // As compiled (g++ 13.3 -O2) the reference's mask function returns (0,0,0) for the very first
    // voxel it visits, where no candidate qualifies and its result struct is uninitialised
    // (jly_3ddt.cpp:469-470): voxel (0,0,0) acts as one extra seed.  Pinned against oracle/_ref.
    if (i == 0) u.v = u.h = u.d = 0;
    G[i] = u;
}

// seeds: ROUND((p - min)*scale) in double, points outside the grid skipped (jly_3ddt.cpp:952-966)
__global__ void dt_seed_kernel(Vox* G, int S, const float* __restrict__ model, int nm, double xmin, double ymin, double zmin, double scale)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nm) return;
    int x = __double2int_rz(__dadd_rn(__dmul_rn(__dsub_rn((double)model[3 * i], xmin), scale), 0.5));
    int y = __double2int_rz(__dadd_rn(__dmul_rn(__dsub_rn((double)model[3 * i + 1], ymin), scale), 0.5));
    int z = __double2int_rz(__dadd_rn(__dmul_rn(__dsub_rn((double)model[3 * i + 2], zmin), scale), 0.5));
    if (x < 0 || x >= S || y < 0 || y >= S || z < 0 || z >= S) return;
    Vox s; s.v = s.h = s.d = 0; s.pad = 0;
    G[((size_t)x * S + y) * S + z] = s;
}

// ---- the sequential propagation ------------------------------------------------------------
// One persistent CTA, thread z owns column z of every row it visits.  Shared memory holds the
// three relevant rows of the adjacent slice and the previous row of the current slice (padded
// with "unset" voxels so that the reference's bounds tests need no branches); every global
// access is a thread's own column, prefetched one row ahead, so HBM/L2 latency stays off the
// critical path.  A row scan is resolved as
//   own_z   = first minimum (mask order) of all candidates that do not depend on the scan,
//   T_z     = the squared norm the running-scan candidate must beat at z (ties by mask order),
//   chain   = runs: a run starts at s when (own_s + (0,0,1)) beats T_{s+1} and lives while
//             own_s + (0,0,k) beats T_{s+k}; where it dies the voxel falls back to own.
// Runs are found in parallel (one thread per possible run head, first wins are exactly the
// "static" wins against the neighbour's own value), then the few heads are filtered in scan
// order so that a head inside an earlier run is ignored -- the sequential semantics exactly.
// Working representation on chip: 32-bit components plus the squared norm, 16 bytes per voxel
// (one LDS.128), so that a candidate costs a handful of integer instructions:
//   |(v+a, h+b, d+c)|^2 = n + 2(a v + b h + c d) + (a+b+c)   for a,b,c in {0,1}.
// Global memory keeps the compact 8-byte (short) form.
struct __align__(16) V4 { int v, h, d, n; };

__device__ __forceinline__ V4 v4_unset() { V4 u; u.v = u.h = u.d = 32767; u.n = kInf; return u; }
__device__ __forceinline__ V4 v4_from(const Vox& a) { V4 r; r.v = a.v; r.h = a.h; r.d = a.d; r.n = vox_unset(a) ? kInf : vox_norm2(a); return r; }
__device__ __forceinline__ Vox v4_to(const V4& a) { Vox r; r.v = (short)a.v; r.h = (short)a.h; r.d = (short)a.d; r.pad = 0; return r; }

// candidate = source + (IV,IH,ID); strict '<' keeps the first minimum in mask order
template <int IV, int IH, int ID>
__device__ __forceinline__ void consider(V4& best, const V4 s)
{
    // unset source (n == kInf): scores ~56755 > 32767 in the reference and is never chosen; kInf + increments stays > any real norm
    const int n = s.n + 2 * (IV * s.v + IH * s.h + ID * s.d) + (IV + IH + ID);
    const bool take = (s.n < kInf) & (n < best.n);
    best.n = take ? n : best.n; best.v = take ? s.v + IV : best.v; best.h = take ? s.h + IH : best.h; best.d = take ? s.d + ID : best.d;
}

// Row-scan kinds (mask functions of jly_3ddt.cpp and where the running-scan entry sits in the
// tie order): F1 = MINforwardDE1 (:502-706), F3 = MINforwardDE3 (:51-131), B1 = MINbackwardDE1
// (:295-500), B3 = MINbackwardDE3 (:171-252), C_UP = MINforwardDE4 (:133-169, scan z ascending),
// C_DN = MINforwardDE2 (:254-293, scan z descending).  The two "quirk" entries of DE3/backwardDE1
// (read (z+1,y) but add (0,1,1)) can never win: the entry just before them offers the same source
// voxel with a strictly smaller increment; they are omitted.
enum ScanKind { F1 = 0, F3 = 1, B1 = 2, B3 = 3, C_UP = 4, C_DN = 5 };

// All row buffers live in one dynamic shared array; the layout is a set of byte offsets so that
// every access is derived from the __shared__ symbol (LDS/STS, not generic loads).
extern __shared__ __align__(16) unsigned char dt_smem[];
struct RowSmem {
    int o_xs[3];         // rows (y-1, y, y+1) % 3 of the adjacent slice, padded: index z+1
    int o_prev;          // final previous row of the current slice in this pass, padded
    int o_own[2];        // per scan position k: the voxel's own (recurrence-free) result; double-buffered across the two scans of a row
    int o_T[2];          // the squared norm the recurrence must beat at k
    int o_run_start, o_run_end, o_accept, o_wmask;
    __device__ __forceinline__ V4* xs(int r) const { return reinterpret_cast<V4*>(dt_smem + o_xs[r]); }
    __device__ __forceinline__ V4* prev() const { return reinterpret_cast<V4*>(dt_smem + o_prev); }
    __device__ __forceinline__ V4* own(int b) const { return reinterpret_cast<V4*>(dt_smem + o_own[b]); }
    __device__ __forceinline__ int* T(int b) const { return reinterpret_cast<int*>(dt_smem + o_T[b]); }
    __device__ __forceinline__ short* run_start() const { return reinterpret_cast<short*>(dt_smem + o_run_start); }
    __device__ __forceinline__ short* run_end() const { return reinterpret_cast<short*>(dt_smem + o_run_end); }
    __device__ __forceinline__ unsigned char* accept() const { return dt_smem + o_accept; }
    __device__ __forceinline__ unsigned* wmask() const { return reinterpret_cast<unsigned*>(dt_smem + o_wmask); }
};

#ifdef GOICP_DT_INSTRUMENT
__device__ unsigned long long g_dt_stats_buf[8];
#define DT_STAT(i, v) atomicAdd(&g_dt_stats_buf[i], (unsigned long long)(v))
#else
#define DT_STAT(i, v) ((void)0)
#endif

// One row scan.  `buf` selects the record buffer (0 for the first scan of a row, 1 for the second)
// so that the second scan may start writing its records while slow threads still read the first's.
// If last_of_row the result is also stored as the row's entry of sh.prev -- optimistically BEFORE
// the barrier that decides whether any run exists, so the common run-free row needs no extra
// barrier; rows with runs rewrite it afterwards.
template <int KIND>
__device__ __forceinline__ V4 row_scan(const RowSmem& sh, int S, int y, const V4 self, int buf, bool last_of_row, bool& had_runs)
{
    const int z = threadIdx.x;
    const bool active = z < S;
    constexpr int dir = (KIND == F1 || KIND == B3 || KIND == C_UP) ? +1 : -1;     // +1: recurrence reads z-1
    const int k = dir > 0 ? z : S - 1 - z;                                      // position in scan order
    V4* own = sh.own(buf); int* T = sh.T(buf);
    V4 P = v4_unset(), Q = v4_unset();
    V4 out = P; int Tk = kInf;
    if (active) {
        const int zp = z + 1;                                                    // padded index
        if (KIND == F1 || KIND == B1) {
            const V4* r0 = sh.xs((y - 1 + 3) % 3); const V4* r1 = sh.xs((y + 3) % 3); const V4* r2 = sh.xs((y + 1 + 3) % 3);
            consider<1, 1, 1>(P, r0[zp - 1]); consider<1, 1, 0>(P, r0[zp]); consider<1, 1, 1>(P, r0[zp + 1]);
            consider<1, 0, 1>(P, r1[zp - 1]); consider<1, 0, 0>(P, r1[zp]); consider<1, 0, 1>(P, r1[zp + 1]);
            consider<1, 1, 1>(P, r2[zp - 1]); consider<1, 1, 0>(P, r2[zp]); consider<1, 1, 1>(P, r2[zp + 1]);
        }
        if (KIND == F1 || KIND == B3) {              // previous row y-1, then self; recurrence comes last
            consider<0, 1, 1>(P, sh.prev()[zp - 1]);
            consider<0, 1, 0>(P, sh.prev()[zp]);
            consider<0, 1, 1>(P, sh.prev()[zp + 1]);
            consider<0, 0, 0>(P, self);
        } else if (KIND == F3 || KIND == B1) {       // recurrence first, then (z,y+1), self, (z-1,y+1)
            consider<0, 1, 0>(Q, sh.prev()[zp]);
            consider<0, 0, 0>(Q, self);
            consider<0, 1, 1>(Q, sh.prev()[zp - 1]);
        } else {                                     // pure chains: recurrence first, then self
            consider<0, 0, 0>(Q, self);
        }
        out = P.n <= Q.n ? P : Q;
        Tk = min(P.n, Q.n + 1);                      // recurrence wins iff nc < nP and nc <= nQ
        own[k] = out; T[k] = Tk;
    }
    __syncthreads();
    // A run can only begin with a win against the neighbour's OWN value.  A "win" whose vector is
    // identical to the voxel's own result changes nothing (state == own either way) and is not a
    // head; likewise a run that arrives at a voxel carrying exactly the vector the voxel would
    // hold anyway simply ends there.  Most recurrence wins are such ties (the chain entry precedes
    // `self` in four of the six masks), so real heads are few and runs short.
    int win = 0, nc = 0;
    V4 o = out;
    if (active && k >= 1) {
        o = own[k - 1];
        if (o.n < kInf) {
            nc = o.n + 2 * o.d + 1;
            win = (nc < Tk) && !(o.v == out.v && o.h == out.h && o.d + 1 == out.d);
        }
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, win);
    if ((threadIdx.x & 31) == 0) sh.wmask()[threadIdx.x >> 5] = ballot;
    if (last_of_row && active) sh.prev()[z + 1] = out;       // optimistic: valid unless a run covers z
    const int any = __syncthreads_or(win);
    DT_STAT(0, threadIdx.x == 0); DT_STAT(1, threadIdx.x == 0 && any);
    had_runs = any != 0;
    if (any) {
        if (active) sh.run_start()[k] = -1;
        if (win) {                                   // extent of the run that starts at k-1
            int n = nc, d = o.d + 1, pos = k + 1;
            while (pos < S) {
                n += 2 * d + 1; d++;
                if (!(n < T[pos])) break;
                const V4 w = own[pos];
                if (n == w.n && w.v == o.v && w.h == o.h && w.d == d) break;
                pos++;
            }
            sh.run_end()[k] = (short)pos;
            DT_STAT(2, 1); DT_STAT(3, pos - k);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            // heads in scan order; a head whose start lies inside an accepted run is void
            int last_end = -1;
            const int nw = (S + 31) / 32;
            if (dir > 0) {
                for (int w = 0; w < nw; w++) {
                    unsigned m = sh.wmask()[w];
                    while (m) { const int b = __ffs(m) - 1; m &= m - 1; const int kk = w * 32 + b;      // kk == z
                        const bool ok = kk - 1 >= last_end; sh.accept()[kk] = ok; if (ok) last_end = sh.run_end()[kk]; }
                }
            } else {
                for (int w = nw - 1; w >= 0; w--) {
                    unsigned m = sh.wmask()[w];
                    while (m) { const int b = 31 - __clz(m); m &= ~(1u << b); const int kk = S - 1 - (w * 32 + b);   // thread z -> k
                        const bool ok = kk - 1 >= last_end; sh.accept()[kk] = ok; if (ok) last_end = sh.run_end()[kk]; }
                }
            }
        }
        __syncthreads();
        if (win && sh.accept()[k]) { const int e = sh.run_end()[k]; for (int pos = k; pos < e; pos++) sh.run_start()[pos] = (short)(k - 1); }
        __syncthreads();
        if (active) {
            const int rs = sh.run_start()[k];
            if (rs >= 0) {
                const V4 s0 = own[rs];
                const int m = k - rs;
                out.v = s0.v; out.h = s0.h; out.d = s0.d + m; out.n = s0.n + 2 * s0.d * m + m * m;
                if (last_of_row) sh.prev()[z + 1] = out;
            }
        }
    }
    return out;
}

// one pass over the rows of slice x: K1 then K2 on every row, rows in direction ydir
template <int K1, int K2>
__device__ void slice_pass(Vox* G, int S, int x, int xs, int ydir, const RowSmem& sh)
{
    const int z = threadIdx.x;
    const bool active = z < S;
    Vox unset_g; unset_g.v = unset_g.h = unset_g.d = 32767; unset_g.pad = 0;
    const bool use_xs = (K1 == F1 || K1 == B1) && xs >= 0 && xs < S;
    const int y0 = ydir > 0 ? 0 : S - 1;
    auto gload = [&](int xx, int yy) -> Vox { return (active && yy >= 0 && yy < S) ? G[((size_t)xx * S + yy) * S + z] : unset_g; };
    // prime the shared rows: adjacent-slice rows y0-1, y0, y0+1 ; previous row of this slice = outside
    if (active) {
        for (int r = -1; r <= 1; r++) sh.xs((y0 + r + 3) % 3)[z + 1] = use_xs ? v4_from(gload(xs, y0 + r)) : v4_unset();
        sh.prev()[z + 1] = v4_unset();
    }
    Vox self_next = gload(x, y0);
    Vox xs_next = use_xs ? gload(xs, y0 + 2 * ydir) : unset_g;
    __syncthreads();
    for (int i = 0, y = y0; i < S; i++, y += ydir) {
        const V4 self = v4_from(self_next);
        const V4 xs_row = v4_from(xs_next);                 // adjacent-slice row y + 2*ydir: needed from the next row on
        self_next = gload(x, y + ydir);                     // prefetch: own column only
        xs_next = use_xs ? gload(xs, y + 3 * ydir) : unset_g;
        {   // the volume (8 B/voxel) exceeds L2 at S=300: pull the rows needed a few iterations from now into L2
            const int yf = y + 8 * ydir;
            if (active && (z & 15) == 0 && yf >= 0 && yf < S) {
                asm volatile("prefetch.global.L2 [%0];" :: "l"(G + ((size_t)x * S + yf) * S + z));
                if (use_xs) asm volatile("prefetch.global.L2 [%0];" :: "l"(G + ((size_t)xs * S + yf) * S + z));
            }
        }
        bool runs1, runs2;
        V4 v = row_scan<K1>(sh, S, y, self, 0, false, runs1);
        // the adjacent-slice row that leaves the 3-row window is replaced by the incoming one; nobody
        // reads that slot again in this row, and the next row's reads come after >= 2 barriers
        if ((K1 == F1 || K1 == B1) && active) sh.xs((y + 2 * ydir + 3 + 3) % 3)[z + 1] = xs_row;
        v = row_scan<K2>(sh, S, y, v, 1, true, runs2);
        if (active) G[((size_t)x * S + y) * S + z] = v4_to(v);
        if (runs2) __syncthreads();                         // prev[] was patched after the deciding barrier
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kMaxS)
dt_propagate_kernel(Vox* G, int S)
{
    RowSmem sh;
    {
        int p = 0;
        const int row = (S + 2) * (int)sizeof(V4);
        for (int r = 0; r < 3; r++) { sh.o_xs[r] = p; p += row; }
        sh.o_prev = p; p += row;
        for (int b = 0; b < 2; b++) { sh.o_own[b] = p; p += S * (int)sizeof(V4); }
        for (int b = 0; b < 2; b++) { sh.o_T[b] = p; p += S * (int)sizeof(int); }
        sh.o_wmask = p; p += 32 * (int)sizeof(unsigned);
        sh.o_run_start = p; p += (S + 2) / 2 * 2 * (int)sizeof(short);
        sh.o_run_end = p; p += (S + 2) / 2 * 2 * (int)sizeof(short);
        sh.o_accept = p;
    }
    // pads of the shared rows stay "unset" for the whole kernel
    if (threadIdx.x == 0) {
        const V4 u = v4_unset();
        for (int r = 0; r < 3; r++) { sh.xs(r)[0] = u; sh.xs(r)[S + 1] = u; }
        sh.prev()[0] = u; sh.prev()[S + 1] = u;
    }
    __syncthreads();
    for (int x = 0; x < S; x++) {                                   // jly_3ddt.cpp:719-728
        slice_pass<F1, C_DN>(G, S, x, x - 1, +1, sh);
        slice_pass<F3, C_UP>(G, S, x, -1, -1, sh);
    }
    for (int x = S - 1; x >= 0; x--) {                              // :729-739
        slice_pass<B1, C_UP>(G, S, x, x + 1, -1, sh);
        slice_pass<B3, C_DN>(G, S, x, -1, +1, sh);
    }
}

static size_t dt_propagate_smem(int S)
{
    return 4 * (size_t)(S + 2) * 16 + 2 * (size_t)S * 16 + 2 * (size_t)S * sizeof(int) + 32 * sizeof(unsigned)
         + 2 * (size_t)(S + 2) * sizeof(short) + (size_t)S + 64;
}

// distance = float( double(float(sqrt(double(n2)))) / scale ), clamped at 0 (jly_3ddt.cpp:970-978);
// also transposes the working [x][y][z] layout into the reference's [z][y][x].
__global__ void dt_finalize_kernel(const Vox* __restrict__ G, int S, double scale, float* __restrict__ out)
{
    __shared__ float tile[32][33];
    const int y = blockIdx.z;
    const int x0 = blockIdx.x * 32, z0 = blockIdx.y * 32;
    {
        const int x = x0 + threadIdx.y, z = z0 + threadIdx.x;
        if (x < S && z < S) {
            const Vox a = G[((size_t)x * S + y) * S + z];
            const double n2 = (double)a.v * (double)a.v + (double)((int)a.h * a.h) + (double)((int)a.d * a.d);
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

// ---- exact EDT (separable, integer squared distances) ----------------------------------------
// pass along z per (x,y) column: squared distance to the nearest seed in the column
__global__ void edt_pass_z(const Vox* __restrict__ G, int S, int* __restrict__ D)
{
    const int col = blockIdx.x * blockDim.x + threadIdx.x;      // x*S + y
    if (col >= S * S) return;
    const Vox* g = G + (size_t)col * S;
    int* d = D + (size_t)col * S;
    int last = -kMaxS * 4;
    for (int z = 0; z < S; z++) { if (g[z].v == 0 && g[z].h == 0 && g[z].d == 0) last = z; int k = z - last; d[z] = k > 4 * kMaxS - 1 ? kInf : k * k; }
    last = kMaxS * 8;
    for (int z = S - 1; z >= 0; z--) { if (g[z].v == 0 && g[z].h == 0 && g[z].d == 0) last = z; int k = last - z; int v = k > 4 * kMaxS - 1 ? kInf : k * k; if (v < d[z]) d[z] = v; }
}
// generic 1-D min-plus pass along a strided line: out[i] = min_j in[j] + (i-j)^2
__global__ void edt_pass_line(const int* __restrict__ in, int* __restrict__ out, int S, size_t line_stride_a, size_t line_stride_b, size_t elem_stride)
{
    extern __shared__ int line[];
    const size_t base = (size_t)blockIdx.x * line_stride_a + (size_t)blockIdx.y * line_stride_b;
    for (int i = threadIdx.x; i < S; i += blockDim.x) line[i] = in[base + (size_t)i * elem_stride];
    __syncthreads();
    for (int i = threadIdx.x; i < S; i += blockDim.x) {
        int best = kInf;
        for (int j = 0; j < S; j++) { const int v = line[j]; if (v < kInf) { const int c = v + (i - j) * (i - j); if (c < best) best = c; } }
        out[base + (size_t)i * elem_stride] = best;
    }
}
__global__ void edt_finalize_kernel(const int* __restrict__ D, int S, double scale, float* __restrict__ out)
{
    // D is [x][y][z]; out is [z][y][x]
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t n3 = (size_t)S * S * S;
    if (i >= n3) return;
    const int x = (int)(i % S), y = (int)((i / S) % S), z = (int)(i / ((size_t)S * S));
    const int n2 = D[((size_t)x * S + y) * S + z];
    float dv = __double2float_rn(sqrt((double)n2));
    out[i] = __double2float_rn(__ddiv_rn((double)dv, scale));
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
    if (S > kMaxS) { msg = "dt_size > 1024"; return cudaErrorInvalidValue; }
    dt_frame_host(model, nm, S, expand, meta);
    const size_t n3 = (size_t)S * S * S;
    Vox* G = nullptr; float* d_model = nullptr; int* D0 = nullptr; int* D1 = nullptr;
    auto cleanup = [&]() { if (G) cudaFree(G); if (d_model) cudaFree(d_model); if (D0) cudaFree(D0); if (D1) cudaFree(D1); };
#define DT_TRY(expr) do { e = (expr); if (e != cudaSuccess) { msg = #expr; cleanup(); return e; } } while (0)
    DT_TRY(cudaMalloc((void**)&G, n3 * sizeof(Vox)));
    DT_TRY(cudaMalloc((void**)&d_model, (size_t)3 * nm * sizeof(float)));
    DT_TRY(cudaMemcpyAsync(d_model, model, (size_t)3 * nm * sizeof(float), cudaMemcpyHostToDevice, stream));
    dt_init_kernel<<<(unsigned)((n3 + 255) / 256), 256, 0, stream>>>(G, n3);
    DT_TRY(cudaGetLastError());
    dt_seed_kernel<<<(nm + 255) / 256, 256, 0, stream>>>(G, S, d_model, nm, meta[0], meta[1], meta[2], meta[3]);
    DT_TRY(cudaGetLastError());
    if (mode == 0) {
        const int threads = ((S + 31) / 32) * 32;
        const size_t smem = dt_propagate_smem(S);
        DT_TRY(cudaFuncSetAttribute(dt_propagate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dt_propagate_kernel<<<1, threads, smem, stream>>>(G, S);
        DT_TRY(cudaGetLastError());
#ifdef GOICP_DT_INSTRUMENT
        {
            unsigned long long hs[8];
            DT_TRY(cudaStreamSynchronize(stream));
            DT_TRY(cudaMemcpyFromSymbol(hs, g_dt_stats_buf, sizeof hs));
            fprintf(stderr, "[dt stats] row scans %llu, with runs %llu, run heads %llu, total run length %llu\n", hs[0], hs[1], hs[2], hs[3]);
        }
#endif
        dim3 grid((S + 31) / 32, (S + 31) / 32, S), block(32, 32);
        dt_finalize_kernel<<<grid, block, 0, stream>>>(G, S, meta[3], d_out);
        DT_TRY(cudaGetLastError());
    } else {
        // the extra corner seed is an artefact of the reference binary, not part of an exact EDT
        Vox unset; unset.v = unset.h = unset.d = 32767; unset.pad = 0;
        DT_TRY(cudaMemcpyAsync(G, &unset, sizeof(Vox), cudaMemcpyHostToDevice, stream));
        dt_seed_kernel<<<(nm + 255) / 256, 256, 0, stream>>>(G, S, d_model, nm, meta[0], meta[1], meta[2], meta[3]);
        DT_TRY(cudaMalloc((void**)&D0, n3 * sizeof(int)));
        DT_TRY(cudaMalloc((void**)&D1, n3 * sizeof(int)));
        edt_pass_z<<<(S * S + 127) / 128, 128, 0, stream>>>(G, S, D0);
        DT_TRY(cudaGetLastError());
        // along y: lines indexed by (x, z): base = x*S*S + z, element stride S
        edt_pass_line<<<dim3(S, S), 128, S * sizeof(int), stream>>>(D0, D1, S, (size_t)S * S, 1, (size_t)S);
        DT_TRY(cudaGetLastError());
        // along x: lines indexed by (y, z): base = y*S + z, element stride S*S
        edt_pass_line<<<dim3(S, S), 128, S * sizeof(int), stream>>>(D1, D0, S, (size_t)S, 1, (size_t)S * S);
        DT_TRY(cudaGetLastError());
        edt_finalize_kernel<<<(unsigned)((n3 + 255) / 256), 256, 0, stream>>>(D0, S, meta[3], d_out);
        DT_TRY(cudaGetLastError());
    }
    DT_TRY(cudaStreamSynchronize(stream));
#undef DT_TRY
    cleanup();
    return cudaSuccess;
}

} // namespace goicp
