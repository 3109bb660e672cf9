// Small path of the batched NMS: one CTA per image, everything (keys, sorted boxes, kept lists) in shared memory.
// Shared by nms.cu (generic boxes from HBM) and yolo.cu (candidates produced in shared memory by the decode).
#pragma once
#include "nms_core.cuh"

namespace det {

constexpr int kSmallThreads = 256;
constexpr int kSmallIdxBits = 12;
constexpr int kWarpSegMax = 64;  // segments up to this length are swept by a single warp

template <int CAP>
struct SmallSmem {
    uint64_t keys[CAP];
    float4 sbox[CAP];
    float sarea[CAP];
    uint16_t klist[CAP];
    uint16_t seg_s[CAP];
    uint16_t seg_e[CAP];
    uint8_t state[CAP];
    uint32_t rowbits[kSmallThreads * (kSmallThreads / 32)];
    uint32_t amask[kSmallThreads / 32];
    float red_max[kSmallThreads / 32];
    float red_min[kSmallThreads / 32];
    int red_flag[kSmallThreads / 32];
    int nseg_small, nseg_large, nk_scratch, nkept, bad_cat;
    float span;
    int fast;
};

// candidate source for boxes that live in HBM as (boxes, scores, categories) rows of one image
struct GlobalCandidates {
    const float4* boxes;
    const float* scores;
    const int64_t* cats;  // may be null: single category
    __device__ __forceinline__ float4 box(int i) const { return boxes[i]; }
    __device__ __forceinline__ float score(int i) const { return scores[i]; }
    __device__ __forceinline__ int64_t cat(int i) const { return cats ? cats[i] : 0; }
};

// Greedy category-partitioned NMS of `cnt` (<= CAP) candidates by one CTA of kSmallThreads threads.
// Returns (block-uniform) the number of kept candidates, or -1 if a category is outside [0, 32767).
// On return sm.keys[0 .. kept) hold (descending-score bits | candidate index) in output order.
template <int CAP, typename Src>
__device__ int small_nms_body(SmallSmem<CAP>& sm, const Src& src, int cnt, float thr_f, int mode, int max_out) {
    using KL = KeyLayout<kSmallIdxBits>;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr int W = kSmallThreads / 32;
    // reference CPU rule: boxes.numel() <= 4000 -> coordinate-offset trick (torchvision/ops/boxes.py batched_nms)
    const bool trick = (mode == DET_NMS_AUTO) ? (cnt <= 1000) : (mode == DET_NMS_OFFSET_TRICK);
    if (tid == 0) {
        sm.nseg_small = 0;
        sm.nseg_large = 0;
        sm.nkept = 0;
        sm.bad_cat = 0;
        sm.span = 0.f;
        sm.fast = 1;
    }
    __syncthreads();
    // ---- phase 0: coordinate statistics for the offset trick
    if (trick) {
        float mx = -INFINITY, mn = INFINITY;
        int fin = 1, maxcat = 0;
        for (int i = tid; i < cnt; i += kSmallThreads) {
            const float4 b = src.box(i);
            mx = max_nan(mx, max_nan(max_nan(b.x, b.y), max_nan(b.z, b.w)));
            mn = min_nan(mn, min_nan(min_nan(b.x, b.y), min_nan(b.z, b.w)));
            fin &= (int)(isfinite(b.x) && isfinite(b.y) && isfinite(b.z) && isfinite(b.w));
            maxcat = max(maxcat, (int)src.cat(i));
        }
        for (int o = 16; o > 0; o >>= 1) {
            mx = max_nan(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            mn = min_nan(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            fin &= __shfl_xor_sync(0xffffffffu, fin, o);
            maxcat = max(maxcat, __shfl_xor_sync(0xffffffffu, maxcat, o));
        }
        if (lane == 0) {
            sm.red_max[wid] = mx;
            sm.red_min[wid] = mn;
            sm.red_flag[wid] = fin | (maxcat << 1);
        }
        __syncthreads();
        if (tid == 0) {
            float gmx = sm.red_max[0], gmn = sm.red_min[0];
            int gfin = sm.red_flag[0] & 1, gcat = sm.red_flag[0] >> 1;
            for (int w = 1; w < W; ++w) {
                gmx = max_nan(gmx, sm.red_max[w]);
                gmn = min_nan(gmn, sm.red_min[w]);
                gfin &= sm.red_flag[w] & 1;
                gcat = max(gcat, sm.red_flag[w] >> 1);
            }
            const float span = gmx + 1.0f;  // max_coordinate + torch.tensor(1).to(boxes)
            sm.span = span;
            // categories can be swept independently iff shifted boxes of different categories cannot intersect:
            // all coordinates finite and > -1, shifted coordinates finite, non-negative threshold
            const float far = gmx + (float)gcat * span;
            sm.fast = (gfin && gmn > -1.0f && thr_f >= 0.0f && isfinite(far)) ? 1 : 0;
        }
        __syncthreads();
    }
    const float span = sm.span;
    const bool by_cat = !trick || sm.fast;
    // ---- phase 1+2: keys, sort
    const int npad = next_pow2(max(cnt, 2));
    for (int i = tid; i < npad; i += kSmallThreads) {
        uint64_t k = kSentinelKey;
        if (i < cnt) {
            const int64_t c = src.cat(i);
            if (c < 0 || c >= (1 << kSegBits) - 1) sm.bad_cat = 1;
            k = KL::make(by_cat ? (uint32_t)(c & ((1 << kSegBits) - 1)) : 0u, src.score(i), (uint32_t)i);
        }
        sm.keys[i] = k;
    }
    __syncthreads();
    cta_bitonic_sort<kSmallThreads>(sm.keys, npad);
    // ---- phase 3: boxes in sorted order (+ offset), segment discovery
    for (int p = tid; p < cnt; p += kSmallThreads) {
        const uint64_t k = sm.keys[p];
        const int i = (int)KL::idx(k);
        float4 b = src.box(i);
        if (trick) {
            const float off = (float)src.cat(i) * span;  // idxs.to(boxes) * (max_coordinate + 1)
            b.x += off; b.y += off; b.z += off; b.w += off;
        }
        sm.sbox[p] = b;
        sm.sarea[p] = box_area(b);
        sm.state[p] = 0;
        const uint32_t sg = KL::seg(k);
        if (p == 0 || KL::seg(sm.keys[p - 1]) != sg) {
            int lo = p + 1, hi = cnt;  // segment end = first position with a larger segment field
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (KL::seg(sm.keys[mid]) > sg) hi = mid; else lo = mid + 1;
            }
            if (lo - p <= kWarpSegMax) {
                const int slot = atomicAdd(&sm.nseg_small, 1);
                sm.seg_s[slot] = (uint16_t)p;
                sm.seg_e[slot] = (uint16_t)lo;
            } else {
                const int slot = atomicAdd(&sm.nseg_large, 1);
                sm.seg_s[CAP - 1 - slot] = (uint16_t)p;
                sm.seg_e[CAP - 1 - slot] = (uint16_t)lo;
            }
        }
    }
    __syncthreads();
    // ---- phase 4: greedy suppression. short segments: one warp each, in parallel; long ones: whole CTA
    const int nsmall = sm.nseg_small, nlarge = sm.nseg_large;
    int mykept = 0;
    for (int sidx = wid; sidx < nsmall; sidx += W)
        mykept += warp_segment_nms<uint16_t>(sm.sbox, sm.sarea, sm.state, sm.klist, (int)sm.seg_s[sidx],
                                             (int)sm.seg_e[sidx], thr_f, max_out);
    if (lane == 0 && mykept) atomicAdd(&sm.nkept, mykept);
    __syncthreads();
    for (int sidx = 0; sidx < nlarge; ++sidx) {
        const int nk = cta_segment_nms<kSmallThreads, uint16_t>(sm.sbox, sm.sarea, sm.state, sm.klist,
                                                                (int)sm.seg_s[CAP - 1 - sidx],
                                                                (int)sm.seg_e[CAP - 1 - sidx], thr_f, max_out,
                                                                sm.rowbits, sm.amask, &sm.nk_scratch);
        if (tid == 0) sm.nkept += nk;
        __syncthreads();
    }
    const int kept = sm.nkept;
    // ---- phase 5: output order = (descending score, index) over the kept candidates
    for (int p = tid; p < npad; p += kSmallThreads) {
        uint64_t k = kSentinelKey;
        if (p < cnt && sm.state[p] == 2) k = KL::strip_seg(sm.keys[p]);
        sm.keys[p] = k;
    }
    __syncthreads();
    cta_bitonic_sort<kSmallThreads>(sm.keys, npad);
    return sm.bad_cat ? -1 : kept;
}

}  // namespace det
