// strict_sum.cuh -- reference-order ("strict") evaluation of a bound: the reference permutes the
// per-point residuals with its intro_select (jly_sorting.hpp:228-313) and then adds them up
// SEQUENTIALLY in float (jly_goicp.cpp:293-315).  The float result therefore depends on that
// permutation.  Wherever a decision of the search hinges on which of two nearly equal sums is
// smaller (the arg-min translation cube of an improving upper-bound pass, the error after an ICP
// refinement), the engine re-evaluates the few contenders with this single-thread emulation so
// the decision -- and hence the refinement the search performs next -- is the reference's.
// One thread per array; arrays of Nd floats in shared or global memory.
#pragma once
#include <cuda_runtime.h>

namespace goicp {

__device__ inline int ss_med_ends(const float* a, int st, int en)            // median_of_st_mid_en (:27-52)
{
    const int mid = (st + en) / 2;
    if (a[st] < a[en]) { if (a[mid] < a[st]) return st; return a[mid] < a[en] ? mid : en; }
    if (a[mid] < a[en]) return en;
    return a[mid] < a[st] ? mid : st;
}
// median_of_3 (:55-80); the reference's else-branch indexes the array BASE (data[1], data[2], data[0])
__device__ inline int ss_med3(const float* a, int st)
{
    const float* b = a + st;
    if (b[0] < b[2]) { if (b[1] < b[0]) return st; return b[1] < b[2] ? st + 1 : st + 2; }
    if (a[1] < a[2]) return st + 2;
    return a[1] < a[0] ? st + 1 : st;
}
__device__ inline int ss_med5(float* a, int st)                                // median_of_5 (:83-146)
{
    float* b = a + st; float t;
    if (b[0] > b[1]) { t = b[0]; b[0] = b[1]; b[1] = t; }
    if (b[2] > b[3]) { t = b[2]; b[2] = b[3]; b[3] = t; }
    if (b[0] < b[2]) { t = b[4]; b[4] = b[0]; if (t < b[1]) b[0] = t; else { b[0] = b[1]; b[1] = t; } }
    else             { t = b[4]; b[4] = b[2]; if (t < b[3]) b[2] = t; else { b[2] = b[3]; b[3] = t; } }
    if (b[0] < b[2]) return b[1] < b[2] ? st + 1 : st + 2;
    return b[0] < b[3] ? st : st + 3;
}
__device__ inline int ss_mom(float* a, int st, int en)                         // median_of_medians (:148-210), tail call unrolled
{
    for (;;) {
        const int l = en - st + 1;
        const int groups = l / 5 + (l % 5 != 0);
        int sub = st, i, med; float t;
        for (i = 0; i + 1 < groups; i++, sub += 5) {
            med = ss_med5(a, sub);
            t = a[st + i]; a[st + i] = a[med]; a[med] = t;
        }
        const int rest = en - sub + 1;
        if (rest == 3 || rest == 4) med = ss_med3(a, sub);
        else if (rest == 5) med = ss_med5(a, sub);
        else med = sub;
        t = a[st + i]; a[st + i] = a[med]; a[med] = t;
        if (groups > 5) { en = st + groups - 1; continue; }
        if (groups == 3 || groups == 4) return ss_med3(a, st);
        if (groups == 5) return ss_med5(a, st);
        return st;
    }
}
// intro_select (:228-313)
__device__ inline void ss_intro_select(float* a, int st, int en, int k)
{
    int l_pre = en - st + 1, tries = 0;
    bool quick = true;
    for (;;) {
        if (st >= en) break;
        if (en - st <= 5) {                                                     // insertion_sort (:212-224)
            for (int i = st + 1; i <= en; i++)
                for (int j = i; j > st && a[j - 1] > a[j]; j--) { float t = a[j - 1]; a[j - 1] = a[j]; a[j] = t; }
            return;
        }
        if (quick && tries++ == 5) {
            const int l = en - st + 1;
            if (l * 2 > l_pre) quick = false;
            l_pre = l; tries = 0;
        }
        const int med = quick ? ss_med_ends(a, st, en) : ss_mom(a, st, en);
        float t;
        if (med != st) { t = a[st]; a[st] = a[med]; a[med] = t; }
        int left = st + 1, right = en;
        const float pivot = a[st];
        for (;;) {
            while (left < right && pivot >= a[left]) ++left;
            while (left < right && pivot <= a[right]) --right;
            if (left >= right) break;
            t = a[left]; a[left] = a[right]; a[right] = t;
        }
        int s = left - 1;
        if (a[left] < pivot) s = left;
        a[st] = a[s]; a[s] = pivot;
        if (s < k) st = s + 1; else if (s > k) en = s - 1; else break;
    }
}

// ---- the same for an array in GLOBAL memory (clouds beyond ~55 k points) ------------------------
// One thread walking a global array pays an L2 round trip per element (the accesses depend on the
// comparisons), 75x a CPU core.  The Hoare sweep only ever touches two fronts, so the thread stages
// them through two shared-memory windows with plain copy loops (independent loads: they pipeline)
// and runs the reference's loop on the windows; a range of <= 2 windows is staged whole.  Same
// comparisons, same swaps, same order -- only where the operands sit differs.
constexpr int kSsWin = 2048;                 // floats per window; callers provide 2 * kSsWin floats of shared memory

// One thread moving a window: what it needs is loads in flight, so the bulk goes as 16 independent 16-byte accesses
// per step on whichever side is in global memory (g = 16-byte aligned part of that side); the other side is shared
// memory, accessed as scalars because the two alignments differ.
__device__ inline void ss_copy(float* dst, const float* src, int n, bool src_is_global = true)
{
    const float* g = src_is_global ? src : dst;
    int head = (int)((16 - ((size_t)g & 15)) & 15) / 4;            // floats until the global side is 16-byte aligned
    if (head > n) head = n;
    int i = 0;
    for (; i < head; i++) dst[i] = src[i];
    if (src_is_global) {
        for (; i + 64 <= n; i += 64) {
            float4 v[16];
#pragma unroll
            for (int u = 0; u < 16; u++) v[u] = *reinterpret_cast<const float4*>(src + i + 4 * u);
#pragma unroll
            for (int u = 0; u < 16; u++) { dst[i + 4 * u] = v[u].x; dst[i + 4 * u + 1] = v[u].y; dst[i + 4 * u + 2] = v[u].z; dst[i + 4 * u + 3] = v[u].w; }
        }
    } else {
        for (; i + 4 <= n; i += 4)
            *reinterpret_cast<float4*>(dst + i) = make_float4(src[i], src[i + 1], src[i + 2], src[i + 3]);
    }
    for (; i < n; i++) dst[i] = src[i];
}
// the partition sweep of intro_select (:283-296) on a[left..right]; returns with left/right where the reference's
// `for (;;)` leaves them
__device__ inline void ss_partition_global(float* a, int& left, int& right, const float pivot, float* buf)
{
    for (;;) {
        const int n = right - left + 1;
        if (n <= 2 * kSsWin) {                                   // the rest fits: the reference's loop verbatim, on the staged copy
            if (n <= 0) return;
            ss_copy(buf, a + left, n);
            int l = 0, r = n - 1;
            for (;;) {
                while (l < r && pivot >= buf[l]) ++l;
                while (l < r && pivot <= buf[r]) --r;
                if (l >= r) break;
                const float t = buf[l]; buf[l] = buf[r]; buf[r] = t;
            }
            ss_copy(a + left, buf, n, false);
            right = left + r; left += l;
            return;
        }
        // two disjoint fronts (left < right holds throughout): L = a[left, left+W), R = a(right-W, right]
        float* L = buf; float* R = buf + kSsWin;
        ss_copy(L, a + left, kSsWin);
        ss_copy(R, a + right - kSsWin + 1, kSsWin);
        int l = 0, r = kSsWin - 1;
        for (;;) {
            while (l < kSsWin && pivot >= L[l]) ++l;
            if (l == kSsWin) break;                               // left front leaves its window in the middle of its scan
            while (r >= 0 && pivot <= R[r]) --r;
            if (r < 0) break;                                     // right front leaves its window; the left one rests on a[left] > pivot
            const float t = L[l]; L[l] = R[r]; R[r] = t;
        }
        ss_copy(a + left, L, kSsWin, false);
        ss_copy(a + right - kSsWin + 1, R, kSsWin, false);
        left += l; right -= kSsWin - 1 - r;                       // resuming with the left scan is what the reference does in both cases
    }
}
__device__ inline void ss_intro_select_global(float* a, int st, int en, int k, float* buf)
{
    int l_pre = en - st + 1, tries = 0;
    bool quick = true;
    for (;;) {
        if (st >= en) break;
        if (en - st <= 5) {                                                     // insertion_sort (:212-224)
            for (int i = st + 1; i <= en; i++)
                for (int j = i; j > st && a[j - 1] > a[j]; j--) { float t = a[j - 1]; a[j - 1] = a[j]; a[j] = t; }
            return;
        }
        if (quick && tries++ == 5) {
            const int l = en - st + 1;
            if (l * 2 > l_pre) quick = false;
            l_pre = l; tries = 0;
        }
        const int med = quick ? ss_med_ends(a, st, en) : ss_mom(a, st, en);
        float t;
        if (med != st) { t = a[st]; a[st] = a[med]; a[med] = t; }
        int left = st + 1, right = en;
        const float pivot = a[st];
        ss_partition_global(a, left, right, pivot, buf);
        int s = left - 1;
        if (a[left] < pivot) s = left;
        a[st] = a[s]; a[s] = pivot;
        if (s < k) st = s + 1; else if (s > k) en = s - 1; else break;
    }
}

// The reference's (trimmed) sums over the selected residuals: ub (jly_goicp.cpp:302-306) and,
// when want_lb, lb with the translation radius gt (:308-315).
// win: 2 * kSsWin floats of shared memory when m is in global memory, nullptr when m itself is in shared memory.
__device__ inline void ss_select_and_sum(float* m, int nd, int inlier_num, bool do_select, float gt, bool want_lb,
                                         float& ub_out, float& lb_out, float* win = nullptr)
{
    if (do_select) { if (win) ss_intro_select_global(m, 0, nd - 1, inlier_num - 1, win); else ss_intro_select(m, 0, nd - 1, inlier_num - 1); }
    float ub = 0.0f, lb = 0.0f;
    if (win) {
        for (int base = 0; base < inlier_num; base += 2 * kSsWin) {
            const int n = inlier_num - base < 2 * kSsWin ? inlier_num - base : 2 * kSsWin;
            ss_copy(win, m + base, n);
            for (int i = 0; i < n; i++) ub = __fadd_rn(ub, __fmul_rn(win[i], win[i]));
        }
    } else
    for (int i = 0; i < inlier_num; i++) ub = __fadd_rn(ub, __fmul_rn(m[i], m[i]));
    if (want_lb)
        for (int i = 0; i < inlier_num; i++) { const float e = __fsub_rn(m[i], gt); if (e > 0.0f) lb = __fadd_rn(lb, __fmul_rn(e, e)); }
    ub_out = ub; lb_out = lb;
}

} // namespace goicp
