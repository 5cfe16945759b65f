// strict_sum.cuh -- reference-order ("strict") evaluation of a bound: the reference permutes the
// per-point residuals with its intro_select (jly_sorting.hpp:228-313) and then adds them up
// SEQUENTIALLY in float (jly_goicp.cpp:293-315).  The float result therefore depends on that
// permutation.  Wherever a decision of the search hinges on which of two nearly equal sums is
// smaller (the arg-min translation cube of an improving upper-bound pass, the error after an ICP
// refinement), the engine re-evaluates the few contenders in that order so the decision -- and
// hence the refinement the search performs next -- is the reference's.
// The selection runs block-wide (ss_intro_select_cta: each partition sweep as two ranked position
// lists and K independent exchanges); what is inherently serial -- pivot rule, ranges of a few
// dozen elements, the final sum -- stays with one thread.  Arrays of Nd floats in shared or global memory.
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
// intro_select (:228-313); l_pre / tries / quick: the state of its pivot-rule switch (the CTA-wide form below hands the tail of a
// selection to one thread in mid-flight)
__device__ inline void ss_intro_select_from(float* a, int st, int en, int k, int l_pre, int tries, bool quick)
{
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
__device__ inline void ss_intro_select(float* a, int st, int en, int k) { ss_intro_select_from(a, st, en, k, en - st + 1, 0, true); }

// ---- the same selection by a whole CTA ---------------------------------------------------------------------------
// What the reference's partition sweep (:283-296) does to the array, said without its two walking indices: the sweep stops
// its left index on elements > pivot and its right index on elements < pivot (elements equal to the pivot stop neither), and
// swaps the pair it stopped on.  A swapped-in element lies behind the index that passes it, so the stops are those of the
// ORIGINAL array: the sweep exchanges the k-th element > pivot counted from the left with the k-th element < pivot counted
// from the right, for k = 1, 2, ... as long as the former lies left of the latter (K exchanges), and its indices meet on the
// (K+1)-th element > pivot if that lies left of the K-th partner (of the range's end for K = 0), else on that partner.
// Two ranked position lists (warp ballots + one scan over the warps' counts), K independent exchanges, one thread for the
// meeting point: same array, same pivot position, for any input including ties and all-equal ranges
// (tests/test_gpu_gaps.py::test_cta_select_matches_reference_permutation).  Pivot choice, the placing of the pivot, ranges
// below kSsSeqBelow elements and the median-of-medians fallback stay with thread 0 (the routines above).
// Every thread of the CTA calls with the same arguments.  a: shared or global memory.  idx: 2 * (en - st + 1) ints, shared or
// global.  sh: 40 64-bit words of shared memory.
constexpr int kSsSeqBelow = 64;
__device__ inline int ss_partition_cta(float* a, int lo, int hi, float pivot, int* idx, unsigned long long* sh)
{
    const unsigned full = 0xffffffffu;
    const int n = hi - lo + 1, T = (int)blockDim.x, t = (int)threadIdx.x, lane = t & 31, warp = t >> 5, nw = T >> 5;
    const int cw = (((n + nw - 1) / nw) + 31) & ~31;                      // elements per warp, whole 32-element rows
    const int wb = min(lo + warp * cw, hi + 1), we = min(wb + cw, hi + 1);
    const unsigned lt = (1u << lane) - 1u;
    int nl = 0, nr = 0;                                                      // stops in this warp's stretch (warp-uniform)
    for (int p0 = wb; p0 < we; p0 += 128) {
        float v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) { const int p = p0 + 32 * u + lane; v[u] = p < we ? a[p] : pivot; }
#pragma unroll
        for (int u = 0; u < 4; u++) { nl += __popc(__ballot_sync(full, v[u] > pivot)); nr += __popc(__ballot_sync(full, v[u] < pivot)); }
    }
    if (lane == 0) sh[warp] = (unsigned long long)(unsigned)nl | ((unsigned long long)(unsigned)nr << 32);
    __syncthreads();
    if (warp == 0) {
        const unsigned long long w = lane < nw ? sh[lane] : 0ull; unsigned long long inc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned long long y = __shfl_up_sync(full, inc, o); if (lane >= o) inc += y; }
        sh[lane] = inc - w;                                                  // stops in the warps to the left
        if (lane == 31) sh[32] = inc;
    }
    __syncthreads();
    const unsigned long long before = sh[warp], tot = sh[32];
    const int NL = (int)(unsigned)tot, NR = (int)(tot >> 32);
    int* lpos = idx; int* rpos = idx + n;                                    // k-th stop from the left / from the right
    int il = (int)(unsigned)before, ir = (int)(before >> 32);
    for (int p0 = wb; p0 < we; p0 += 128) {
        float v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) { const int p = p0 + 32 * u + lane; v[u] = p < we ? a[p] : pivot; }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const unsigned bl = __ballot_sync(full, v[u] > pivot), br = __ballot_sync(full, v[u] < pivot);
            const int p = p0 + 32 * u + lane;
            if (v[u] > pivot) lpos[il + __popc(bl & lt)] = p;
            else if (v[u] < pivot) rpos[NR - 1 - (ir + __popc(br & lt))] = p;
            il += __popc(bl); ir += __popc(br);
        }
    }
    __syncthreads();
    const int m = min(NL, NR);
    int fail = m;                                                            // first k whose pair has crossed (the lists are monotone)
    for (int k = t; k < m; k += T) {
        const int p = lpos[k], q = rpos[k];
        if (p < q) { const float vp = a[p]; a[p] = a[q]; a[q] = vp; } else { fail = k; break; }
    }
    fail = __reduce_min_sync(full, (unsigned)fail);
    if (lane == 0) sh[warp] = (unsigned long long)(unsigned)fail;
    __syncthreads();
    if (t == 0) {
        int K = m;
        for (int w = 0; w < nw; w++) K = min(K, (int)(unsigned)sh[w]);
        const int rK = K > 0 ? rpos[K - 1] : hi;
        sh[33] = (unsigned long long)(unsigned)((K < NL && lpos[K] < rK) ? lpos[K] : rK);
    }
    __syncthreads();
    return (int)(unsigned)sh[33];
}
// median_of_medians (:148-210) by the whole CTA.  One level of the reference's loop takes the median of every group of five
// (ss_med5, which also reorders its group) and swaps it to the front, a[st + i] <-> a[med_i], group after group.  Group i lies
// wholly behind a[st + i], so every group is still untouched when its turn comes: the medians are independent.  The swaps
// chain only through the front: the value swap i finds at a[st + i] is the one swap floor(i / 5) left there if that swap's
// median position was a[st + i] (and so on down that chain, <= log5 n steps), else the group-reordered original.  So: all
// groups in parallel, then every i resolves (median value, displaced value) by walking its chain, then two conflict-free
// write phases (displaced values first, the front -- which later swaps own -- last).  The ragged last group and ranges of a
// few hundred elements stay with thread 0, in the reference's order.  scratch: 3 * groups 32-bit words.
constexpr int kSsMomSeqBelow = 320;
__device__ inline int ss_mom_cta(float* a, int st, int en, int* scratch, unsigned long long* sh)
{
    const int T = (int)blockDim.x, t = (int)threadIdx.x;
    for (;;) {
        const int l = en - st + 1;
        if (l < kSsMomSeqBelow) {
            if (t == 0) sh[36] = (unsigned long long)(unsigned)ss_mom(a, st, en);
            __syncthreads();
            const int med = (int)(unsigned)sh[36];
            __syncthreads();
            return med;
        }
        const int groups = l / 5 + (l % 5 != 0), full = groups - 1;        // groups 0 .. full-1 are whole
        int* medpos = scratch; float* Mv = reinterpret_cast<float*>(scratch + groups); float* Vv = reinterpret_cast<float*>(scratch + 2 * groups);
        for (int i = t; i < full; i += T) {
            float* g = a + st + 5 * i;
            float b[5] = {g[0], g[1], g[2], g[3], g[4]};
            float x; int m;
            if (b[0] > b[1]) { x = b[0]; b[0] = b[1]; b[1] = x; }
            if (b[2] > b[3]) { x = b[2]; b[2] = b[3]; b[3] = x; }
            if (b[0] < b[2]) { x = b[4]; b[4] = b[0]; if (x < b[1]) b[0] = x; else { b[0] = b[1]; b[1] = x; } }
            else             { x = b[4]; b[4] = b[2]; if (x < b[3]) b[2] = x; else { b[2] = b[3]; b[3] = x; } }
            if (b[0] < b[2]) m = b[1] < b[2] ? 1 : 2; else m = b[0] < b[3] ? 0 : 3;
            g[0] = b[0]; g[1] = b[1]; g[2] = b[2]; g[3] = b[3]; g[4] = b[4];
            medpos[i] = st + 5 * i + m;
        }
        __syncthreads();
        for (int i = t; i < full; i += T) {
            Mv[i] = a[medpos[i]];
            int ii = i;
            while (ii > 0 && medpos[ii / 5] == st + ii) ii /= 5;           // swap ii/5 left its displaced value on a[st + ii]
            Vv[i] = a[st + ii];
        }
        __syncthreads();
        for (int i = t; i < full; i += T) a[medpos[i]] = Vv[i];
        __syncthreads();
        for (int i = t; i < full; i += T) a[st + i] = Mv[i];
        __syncthreads();
        if (t == 0) {                                                       // the last group, as the reference does it
            const int sub = st + 5 * full, rest = en - sub + 1;
            int med;
            if (rest == 3 || rest == 4) med = ss_med3(a, sub);
            else if (rest == 5) med = ss_med5(a, sub);
            else med = sub;
            const float x = a[st + full]; a[st + full] = a[med]; a[med] = x;
        }
        __syncthreads();
        en = st + groups - 1;                                               // groups > 5 here: next level over the medians
    }
}
__device__ inline void ss_intro_select_cta(float* a, int st, int en, int k, int* idx, unsigned long long* sh)
{
    int l_pre = en - st + 1, tries = 0;
    bool quick = true;
    for (;;) {
        if (st >= en) break;
        if (en - st < kSsSeqBelow) { if (threadIdx.x == 0) ss_intro_select_from(a, st, en, k, l_pre, tries, quick); break; }
        if (quick && tries++ == 5) {
            const int l = en - st + 1;
            if (l * 2 > l_pre) quick = false;
            l_pre = l; tries = 0;
        }
        const int mom = quick ? -1 : ss_mom_cta(a, st, en, idx, sh);
        if (threadIdx.x == 0) {
            const int med = quick ? ss_med_ends(a, st, en) : mom;
            if (med != st) { const float t = a[st]; a[st] = a[med]; a[med] = t; }
            sh[34] = (unsigned long long)__float_as_uint(a[st]);
        }
        __syncthreads();
        const float pivot = __uint_as_float((unsigned)sh[34]);
        const int left = ss_partition_cta(a, st + 1, en, pivot, idx, sh);
        if (threadIdx.x == 0) {
            int s = left - 1;
            if (a[left] < pivot) s = left;
            a[st] = a[s]; a[s] = pivot;
            sh[35] = (unsigned long long)(unsigned)s;
        }
        __syncthreads();
        const int s = (int)(unsigned)sh[35];
        if (s < k) st = s + 1; else if (s > k) en = s - 1; else break;
    }
    __syncthreads();
}

// ---- reading an array in GLOBAL memory from one thread -----------------------------------------------------------
// One thread walking a global array pays an L2 round trip per element; the sequential sum below therefore stages the
// residuals through a shared-memory window with plain copy loops (independent loads: they pipeline).
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
// The reference's (trimmed) sums over the selected residuals: ub (jly_goicp.cpp:302-306) and,
// when want_lb, lb with the translation radius gt (:308-315).
// win: 2 * kSsWin floats of shared memory when m is in global memory, nullptr when m itself is in shared memory.
__device__ inline void ss_sum_selected(const float* m, int inlier_num, float gt, bool want_lb, float& ub_out, float& lb_out, float* win = nullptr)
{
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
