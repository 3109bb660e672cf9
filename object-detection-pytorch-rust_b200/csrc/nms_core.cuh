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
template <int T>
__device__ __forceinline__ void cta_bitonic_sort(uint64_t* keys, int n) {
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < (n >> 1); t += T) {
                const int i = 2 * t - (t & (j - 1));
                const uint64_t a = keys[i], b = keys[i + j];
                const bool up = (i & k) == 0;
                if ((a > b) == up) {
                    keys[i] = b;
                    keys[i + j] = a;
                }
            }
            __syncthreads();
        }
    }
}

__device__ __forceinline__ int next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

// ---- greedy suppression of one segment [s,e) by ONE WARP -------------------------------------------
// sbox/sarea: boxes in sorted order; state: 0 candidate, 1 ignored, 2 kept; klist[s..s+nk): kept positions.
template <typename KT>
__device__ int warp_segment_nms(const float4* sbox, const float* sarea, uint8_t* state, KT* klist, int s, int e,
                                float thr_f, int max_keep) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
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
        bool alive = act;
        for (int k = 0; k < nk; ++k) {
            if (!__any_sync(FULL, alive)) break;
            const int kp = (int)klist[s + k];
            const float4 kb = sbox[kp];
            const float ka = sarea[kp];
            if (alive && nms_suppresses(kb, ka, mb, ma, thr_f)) alive = false;
        }
        unsigned m = __ballot_sync(FULL, alive);
        while (m) {
            const int l = __ffs(m) - 1;
            float4 kb;
            kb.x = __shfl_sync(FULL, mb.x, l);
            kb.y = __shfl_sync(FULL, mb.y, l);
            kb.z = __shfl_sync(FULL, mb.z, l);
            kb.w = __shfl_sync(FULL, mb.w, l);
            const float ka = __shfl_sync(FULL, ma, l);
            if (lane == l) {
                klist[s + nk] = (KT)p;
                state[p] = 2;
            }
            ++nk;
            if (nk >= max_keep) break;
            if (alive && lane > l && nms_suppresses(kb, ka, mb, ma, thr_f)) alive = false;
            m = __ballot_sync(FULL, alive) & ~((2u << l) - 1u);
        }
        __syncwarp();
    }
    return nk;
}

// ---- greedy suppression of one segment [s,e) by a WHOLE CTA of T threads ----------------------------
// scratch: rowbits[T * T/32], amask[T/32], s_nk[1] in shared memory.
template <int T, typename KT>
__device__ int cta_segment_nms(const float4* sbox, const float* sarea, uint8_t* state, KT* klist, int s, int e,
                               float thr_f, int max_keep, uint32_t* rowbits, uint32_t* amask, int* s_nk) {
    constexpr int W = T / 32;
    const unsigned FULL = 0xffffffffu;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    int nk = 0;
    for (int base = s; base < e && nk < max_keep; base += T) {
        const int p = base + tid;
        const bool act = (p < e) && (state[p] == 0);
        float4 mb = make_float4(0.f, 0.f, 0.f, 0.f);
        float ma = 0.f;
        if (act) {
            mb = sbox[p];
            ma = sarea[p];
        }
        bool alive = act;
        // (a) against everything already kept in this segment
        for (int k = 0; k < nk; ++k) {
            if (!__any_sync(FULL, alive)) break;
            const int kp = (int)klist[s + k];
            const float4 kb = sbox[kp];
            const float ka = sarea[kp];
            if (alive && nms_suppresses(kb, ka, mb, ma, thr_f)) alive = false;
        }
        const unsigned bal = __ballot_sync(FULL, alive);
        if (lane == 0) amask[wid] = bal;
        __syncthreads();
        // (b) bit row of this candidate as suppressor of the later survivors of the chunk
        if (alive) {
            for (int w2 = wid; w2 < W; ++w2) {
                unsigned cand = amask[w2];
                if (w2 == wid) cand &= ~((2u << lane) - 1u);
                unsigned bits = 0;
                while (cand) {
                    const int b = __ffs(cand) - 1;
                    cand &= cand - 1;
                    const int q = base + w2 * 32 + b;
                    if (nms_suppresses(mb, ma, sbox[q], sarea[q], thr_f)) bits |= 1u << b;
                }
                rowbits[tid * W + w2] = bits;
            }
        }
        __syncthreads();
        // (c) sequential resolution by warp 0: lane w owns survivor word w
        if (wid == 0) {
            unsigned word = (lane < W) ? amask[lane] : 0u;
            int nkl = nk;
            while (true) {
                const unsigned has = __ballot_sync(FULL, word != 0u);
                if (!has) break;
                const int f = __ffs(has) - 1;
                const unsigned wf = __shfl_sync(FULL, word, f);
                const int b = __ffs(wf) - 1;
                const int c = f * 32 + b;
                if (lane == 0) {
                    klist[s + nkl] = (KT)(base + c);
                    state[base + c] = 2;
                }
                ++nkl;
                if (lane == f) word &= ~(1u << b);
                if (nkl >= max_keep) break;
                if (lane < W && lane >= f) word &= ~rowbits[c * W + lane];
            }
            if (lane == 0) *s_nk = nkl;
        }
        __syncthreads();
        nk = *s_nk;
    }
    return nk;
}

}  // namespace det
