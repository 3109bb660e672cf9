// Device building blocks of the NMS subsystem: 64-bit composite sort keys, CTA bitonic sort, merge-path
// merge passes, and the greedy suppression of one category segment by a warp or by a whole CTA.
//
// Algorithm (replaces torchvision's nms/batched_nms as reached from python/src/utils.py:96-119):
//   1. every box gets the unique key  (segment | descending-score bits | index)  -> an ascending integer
//      sort yields, per segment (category), torch's *stable* descending score order;
//   2. a segment is swept in that order in chunks; a candidate dies iff one of the boxes ALREADY KEPT in its
//      segment suppresses it (exactly the greedy rule), so each chunk is (a) tested against the kept list in
//      parallel and (b) resolved internally with warp ballots / a small bit matrix.  Work is O(kept x boxes)
//      instead of the O(boxes^2) bitmask of the classic GPU NMS, and no boxes^2 mask is ever stored;
//   3. kept boxes are re-keyed by (descending score | index) and sorted once more for the output order.
#pragma once
#include "common.cuh"

namespace det {

constexpr uint64_t kSentinelKey = ~0ull;
constexpr int kSegBits = 15;  // categories / levels must be < 32767

template <int IDX_BITS>
struct KeyLayout {
    static constexpr int kScoreShift = IDX_BITS;
    static constexpr int kSegShift = IDX_BITS + 32;
    static_assert(IDX_BITS + 32 + kSegBits <= 64, "key does not fit");
    __device__ __forceinline__ static uint64_t make(uint32_t seg, float score, uint32_t idx) {
        return ((uint64_t)seg << kSegShift) | ((uint64_t)score_desc_key(score) << kScoreShift) | (uint64_t)idx;
    }
    __device__ __forceinline__ static uint32_t idx(uint64_t k) { return (uint32_t)(k & ((1ull << IDX_BITS) - 1)); }
    __device__ __forceinline__ static uint32_t seg(uint64_t k) { return (uint32_t)(k >> kSegShift); }
    // drop the segment field: (descending score | index), the output order key
    __device__ __forceinline__ static uint64_t strip_seg(uint64_t k) { return k & ((1ull << kSegShift) - 1); }
};

// ---- CTA-wide bitonic sort of n (power of two) keys in shared memory, ascending ------------------
// Register-resident network: thread t owns the E consecutive keys [t*E, t*E+E).  Compare-exchange distance j
//   j < E        : inside the thread (registers only)
//   j < 32*E     : partner is lane ^ (j/E) of the same warp (64-bit shuffles, no barrier)
//   j >= 32*E    : partner lives in another warp -> one round trip through shared memory
// so a 1024-key sort by 256 threads needs 6 shared-memory stages instead of 55 barrier-separated ones.
template <int E>
__device__ __forceinline__ void bitonic_ce_pair(uint64_t& lo, uint64_t& hi, bool up) {
    const bool sw = (lo > hi) == up;
    const uint64_t a = sw ? hi : lo, b = sw ? lo : hi;
    lo = a;
    hi = b;
}

template <int T, int E>
__device__ void cta_bitonic_sort_reg(uint64_t* keys, int n) {
    constexpr int LOG_E = (E == 1) ? 0 : (E == 2) ? 1 : (E == 4) ? 2 : (E == 8) ? 3 : 4;
    constexpr int LOG_MAX = LOG_E + 8 + 1;  // T <= 512
    const int t = threadIdx.x;
    const int base = t * E;
    uint64_t v[E];
#pragma unroll
    for (int e = 0; e < E; ++e) v[e] = (base + e < n) ? keys[base + e] : kSentinelKey;
    for (int a = 1; (1 << a) <= n; ++a) {
        const int k = 1 << a;
#pragma unroll
        for (int b = LOG_MAX - 1; b >= 0; --b) {
            if (b >= a) continue;
            const int j = 1 << b;
            if (b < LOG_E) {  // in registers
#pragma unroll
                for (int e = 0; e < E; ++e)
                    if ((e & j) == 0) bitonic_ce_pair<E>(v[e], v[e | j], ((base + e) & k) == 0);
            } else if (b < LOG_E + 5) {  // inside the warp
                const int lane_xor = j >> LOG_E;
                const bool lower = (t & lane_xor) == 0;
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const uint64_t other = __shfl_xor_sync(0xffffffffu, v[e], lane_xor);
                    const bool up = ((base + e) & k) == 0;
                    const bool keep_min = (up == lower);
                    const bool other_smaller = other < v[e];
                    v[e] = (keep_min == other_smaller) ? other : v[e];
                }
            } else {  // across warps
#pragma unroll
                for (int e = 0; e < E; ++e) keys[base + e] = v[e];
                __syncthreads();
                const bool lower = (base & j) == 0;
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const uint64_t other = keys[(base + e) ^ j];
                    const bool up = ((base + e) & k) == 0;
                    const bool keep_min = (up == lower);
                    const bool other_smaller = other < v[e];
                    v[e] = (keep_min == other_smaller) ? other : v[e];
                }
                __syncthreads();
            }
        }
    }
#pragma unroll
    for (int e = 0; e < E; ++e) keys[base + e] = v[e];
    __syncthreads();
}

// keys[0..n) ascending; n a power of two, n <= 16*T; the array must hold at least max(n, T) entries
template <int T>
__device__ __forceinline__ void cta_bitonic_sort(uint64_t* keys, int n) {
    if (n <= T) cta_bitonic_sort_reg<T, 1>(keys, n);
    else if (n == 2 * T) cta_bitonic_sort_reg<T, 2>(keys, n);
    else if (n == 4 * T) cta_bitonic_sort_reg<T, 4>(keys, n);
    else if (n == 8 * T) cta_bitonic_sort_reg<T, 8>(keys, n);
    else cta_bitonic_sort_reg<T, 16>(keys, n);
}

__device__ __forceinline__ int next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

// ---- warp-collective test of every lane's own box against the kept list klist[k0, k1) ------------------------------
// The kept boxes are fetched 32 at a time, one per lane (a single round of memory latency: kept index, then box + area),
// and handed round with shuffles -- the inner loop never waits for memory, which matters when the arrays live in HBM/L2
// (large path) and the list is hundreds of boxes long.  Returns the lane's `alive` flag after the tests.
template <typename KT, bool NONAN>
__device__ __forceinline__ bool alive_after_kept(const float4* sbox, const float* sarea, const KT* klist, int k0, int k1,
                                                 const float4 mb, float ma, bool alive, float thr_f) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    for (int kb = k0; kb < k1; kb += 32) {
        if (!__any_sync(FULL, alive)) break;
        float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
        float a = 0.f;
        if (kb + lane < k1) {
            const int kp = (int)klist[kb + lane];
            b = sbox[kp];
            a = sarea[kp];
        }
        const int cnt = min(32, k1 - kb);
        for (int j = 0; j < cnt; ++j) {
            if ((j & 7) == 0 && !__any_sync(FULL, alive)) break;
            const float4 ob = make_float4(__shfl_sync(FULL, b.x, j), __shfl_sync(FULL, b.y, j), __shfl_sync(FULL, b.z, j),
                                          __shfl_sync(FULL, b.w, j));
            const float oa = __shfl_sync(FULL, a, j);
            if (alive && nms_suppresses<NONAN>(ob, oa, mb, ma, thr_f)) alive = false;
        }
    }
    return alive;
}

// ---- greedy suppression of one segment [s,e) by ONE WARP -------------------------------------------
// sbox/sarea: boxes in sorted order; state: 0 candidate, 1 ignored, 2 kept; klist[s..s+nk): kept positions.
// Per 32-box chunk: (a) every lane tests its box against the kept list, (b) every surviving lane builds, with
// independent IoU tests, the 32-bit row "which later survivors do I suppress", (c) the greedy order is resolved on
// those bit rows alone (ffs / shfl / and-not), (d) kept lanes append themselves with a popc rank.
template <typename KT, bool NONAN>
__device__ int warp_segment_nms(const float4* sbox, const float* sarea, uint8_t* state, KT* klist, int s, int e,
                                float thr_f, int max_keep) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    int nk = 0;
    for (int base = s; base < e && nk < max_keep; base += 32) {
        const int p = base + lane;
        const bool act = (p < e) && (state[p] == 0);
        float4 mb = make_float4(0.f, 0.f, 0.f, 0.f);
        float ma = 0.f;
        if (act) {
            mb = sbox[p];
            ma = sarea[p];
        }
        const bool alive = alive_after_kept<KT, NONAN>(sbox, sarea, klist, s, s + nk, mb, ma, act, thr_f);  // (a)
        const unsigned am = __ballot_sync(FULL, alive);
        // (b) every unordered pair of the chunk is tested once: lane l meets lane (l+d)&31 for d = 1..16.  If the
        //     partner is later in the order, lane l is the suppressor (row bit); if the rotation wrapped, the partner
        //     is the suppressor and lane l only records "killed by" (col bit).
        unsigned row = 0, col = 0;
        const int nchunk = min(32, e - base);
        if (nchunk > 1) {
#pragma unroll 4
            for (int d = 1; d <= 16; ++d) {
                if (d >= nchunk) break;
                const int j = (lane + d) & 31;
                if (!alive || !((am >> j) & 1u) || (d == 16 && lane >= 16)) continue;
                const float4 bj = sbox[base + j];
                const float aj = sarea[base + j];
                if (j > lane) {
                    if (nms_suppresses<NONAN>(mb, ma, bj, aj, thr_f)) row |= 1u << j;
                } else {
                    if (nms_suppresses<NONAN>(bj, aj, mb, ma, thr_f)) col |= 1u << j;
                }
            }
        }
        // (c) only lanes that suppress some survivor of the chunk can change anything: visit those, in order
        unsigned keepmask = am;
        unsigned nz = (__ballot_sync(FULL, row != 0u) | __reduce_or_sync(FULL, col)) & am;
        while (nz) {
            const int l = __ffs(nz) - 1;
            nz &= nz - 1;
            const unsigned killed = __shfl_sync(FULL, row, l) | __ballot_sync(FULL, (col >> l) & 1u);
            if ((keepmask >> l) & 1u) keepmask &= ~killed;
        }
        const int rank = nk + __popc(keepmask & lt_mask);  // (d)
        if (((keepmask >> lane) & 1u) && rank < max_keep) {
            klist[s + rank] = (KT)p;
            state[p] = 2;
        }
        nk = min(nk + __popc(keepmask), max_keep);
        __syncwarp();
    }
    return nk;
}

// ---- greedy suppression of one segment [s,e) by a WHOLE CTA, in chunks of T boxes ---------------------
// blockDim.x must be a multiple of T.  With G = blockDim.x / T > 1 the extra thread groups are helpers: every group
// looks at the same T candidates, group g tests them against the g-th slice of the kept list (the verdicts meet in
// `deadmask`) and builds the bit rows of the survivor words w2 = g (mod G) -- the two O(chunk x list) parts of a
// chunk are spread over all warps of the CTA; only the order resolution (c) is serial (warp 0).
// scratch in shared memory: rowbits[T * T/32], amask[T/32], deadmask[T/32], s_nk[1].
template <int T, typename KT, bool NONAN>
__device__ int cta_segment_nms(const float4* sbox, const float* sarea, uint8_t* state, KT* klist, int s, int e,
                               float thr_f, int max_keep, uint32_t* rowbits, uint32_t* amask, uint32_t* deadmask,
                               int* s_nk) {
    constexpr int W = T / 32;
    const unsigned FULL = 0xffffffffu;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int G = (int)blockDim.x / T;
    const int gidx = tid / T, ctid = tid - gidx * T, cw = ctid >> 5;
    int nk = 0;
    DET_ACC_BEGIN();
    for (int base = s; base < e && nk < max_keep; base += T) {
        const int p = base + ctid;
        const bool act = (p < e) && (state[p] == 0);
        float4 mb = make_float4(0.f, 0.f, 0.f, 0.f);
        float ma = 0.f;
        if (act) {
            mb = sbox[p];
            ma = sarea[p];
        }
        DET_ACC(0);
        // (a) against everything already kept in this segment (this group's slice of the list)
        bool alive;
        if (G > 1) {
            if (tid < W) deadmask[tid] = 0u;
            __syncthreads();
            const int k0 = s + (int)((int64_t)nk * gidx / G), k1 = s + (int)((int64_t)nk * (gidx + 1) / G);
            const bool mine = alive_after_kept<KT, NONAN>(sbox, sarea, klist, k0, k1, mb, ma, act, thr_f);
            const unsigned dead = __ballot_sync(FULL, act && !mine);
            if (lane == 0 && dead) atomicOr(&deadmask[cw], dead);
            __syncthreads();
            alive = act && !((deadmask[cw] >> lane) & 1u);
        } else {
            alive = alive_after_kept<KT, NONAN>(sbox, sarea, klist, s, s + nk, mb, ma, act, thr_f);
        }
        const unsigned bal = __ballot_sync(FULL, alive);
        if (gidx == 0 && lane == 0) amask[cw] = bal;
        __syncthreads();
        DET_ACC(1);
        // (b) bit row of this candidate as suppressor of the later survivors of the chunk
        if (alive) {
            for (int w2 = cw; w2 < W; ++w2) {
                if (w2 % G != gidx) continue;
                unsigned cand = amask[w2];
                if (w2 == cw) cand &= ~((2u << lane) - 1u);
                unsigned bits = 0;
                while (cand) {
                    const int b = __ffs(cand) - 1;
                    cand &= cand - 1;
                    const int q = base + w2 * 32 + b;
                    if (nms_suppresses<NONAN>(mb, ma, sbox[q], sarea[q], thr_f)) bits |= 1u << b;
                }
                rowbits[ctid * W + w2] = bits;
            }
        }
        __syncthreads();
        DET_ACC(2);
        // (c) sequential resolution by warp 0, one 32-candidate word at a time: lane w owns survivor word w; inside
        //     the current word the greedy order is resolved on its diagonal bit rows, then the kept rows are
        //     OR-ed into a removal mask for the later words (independent shared-memory loads).
        if (wid == 0) {
            unsigned word = (lane < W) ? amask[lane] : 0u;
            int nkl = nk;
            for (int f = 0; f < W && nkl < max_keep; ++f) {
                const unsigned rem = __shfl_sync(FULL, word, f);
                if (!rem) continue;
                const unsigned diag = rowbits[(f * 32 + lane) * W + f];  // stale for dead lanes: only read if alive
                // only survivors whose row reaches another survivor of the word can change anything: visit those, in
                // order (with few suppressions -- the long, high-survival segments -- that is a handful of steps, not 32)
                unsigned keepmask = rem;
                unsigned nz = __ballot_sync(FULL, ((rem >> lane) & 1u) && (diag & rem) != 0u);
                while (nz) {
                    const int b = __ffs(nz) - 1;
                    nz &= nz - 1;
                    const unsigned row = __shfl_sync(FULL, diag, b);
                    if ((keepmask >> b) & 1u) keepmask &= ~row;
                }
                if (nkl + __popc(keepmask) > max_keep) {  // keep only the first (max_keep - nkl) of them
                    unsigned km = keepmask, trimmed = 0;
                    for (int t = nkl; t < max_keep; ++t) {
                        trimmed |= km & (0u - km);
                        km &= km - 1;
                    }
                    keepmask = trimmed;
                }
                const int rank = nkl + __popc(keepmask & ((1u << lane) - 1u));
                if ((keepmask >> lane) & 1u) {
                    klist[s + rank] = (KT)(base + f * 32 + lane);
                    state[base + f * 32 + lane] = 2;
                }
                nkl += __popc(keepmask);
                if (lane < W && lane > f) {
                    unsigned kill = 0;
                    for (unsigned km = keepmask; km; km &= km - 1) kill |= rowbits[(f * 32 + __ffs(km) - 1) * W + lane];
                    word &= ~kill;
                }
            }
            if (lane == 0) *s_nk = nkl;
        }
        __syncthreads();
        DET_ACC(3);
        nk = *s_nk;
    }
    return nk;
}

}  // namespace det
