// NMS over per-image candidate LISTS (box, score, class, id) of at most CAP entries: one CTA per image, built on
// small_nms_body.  Two users:
//   * det_dense_detect (dense_decode.cu): the lists are the positions of a dense head that passed the score threshold;
//   * det_nms_batched, large path with max_out << boxes (nms_large.cuh): the lists are the best-scored boxes of each
//     image, chosen by large_topk_select_kernel ("external tier": statistics of the full image come with the list, and
//     an image whose list does not yield max_out survivors is handed on to the full path through todo[]).
#pragma once
#include "nms_core.cuh"
#include "nms_small.cuh"

namespace det {

constexpr int kCountStride = 32;  // one 128-byte line per image counter: same-line atomics serialise in L2

// T = 256 threads (several CTAs per SM) for large batches; 512 when there are no more images than SMs, where the
// latency of the single CTA that owns an image is all that counts.
template <int CAP, int T>
struct DetectSmem {
    SmallSmem<CAP, T> nms;
    uint16_t perm[CAP];  // i-th candidate handed to the NMS -> slot in the image's list
    uint32_t hist[256];
    FullStats full;
    uint32_t sel_prefix;
    int sel_want, sub_count;
};

struct ListCandidates {
    const float4* boxes;
    const float* scores;
    const int32_t* cls;
    const uint16_t* perm;  // null: identity
    __device__ __forceinline__ int slot(int i) const { return perm ? (int)perm[i] : i; }
    __device__ __forceinline__ float4 box(int i) const { return boxes[slot(i)]; }
    __device__ __forceinline__ float score(int i) const { return scores[slot(i)]; }
    __device__ __forceinline__ int64_t cat(int i) const { return (int64_t)cls[slot(i)]; }
};

template <int CAP, int T>
__global__ void __launch_bounds__(T, 512 / T)
dense_detect_nms_kernel(const int32_t* __restrict__ cand_count, const float4* __restrict__ cand_box,
                        const float* __restrict__ cand_score, const int32_t* __restrict__ cand_cls,
                        const int32_t* __restrict__ cand_id, int cand_cap, float thr_f, int mode, int64_t max_det,
                        int64_t* __restrict__ det_idx, float4* __restrict__ det_boxes, float* __restrict__ det_scores,
                        int64_t* __restrict__ det_classes, int32_t* __restrict__ det_count,
                        int32_t* __restrict__ overflow_flag, const FullStats* __restrict__ ext_full = nullptr,
                        int32_t* __restrict__ todo = nullptr) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    DetectSmem<CAP, T>& sm = *reinterpret_cast<DetectSmem<CAP, T>*>(smem_raw);
    using KL = KeyLayout<kSmallIdxBits>;
    const int img = blockIdx.x, tid = threadIdx.x;
    pdl_wait();  // no-op unless launched with launch_pdl (det_dense_detect): the candidate lists are complete from here
    DET_MARK(0);
    const int cnt = cand_count[(int64_t)img * kCountStride];
    if (ext_full) {  // external tier: a negative count means "no usable list", the full path takes the image
        if (tid == 0) todo[img] = (cnt < 0 || cnt > cand_cap) ? 1 : 0;
        if (cnt < 0 || cnt > cand_cap) return;
    }
    if (cnt > cand_cap) {  // the list is incomplete: report, do not guess
        if (tid == 0) {
            det_count[img] = -1;
            if (overflow_flag) atomicExch(overflow_flag, 1);
        }
        return;
    }
    const int64_t base = (int64_t)img * cand_cap;
    const int cap_out = (int)min(max_det, (int64_t)CAP);
    const int lane = tid & 31, wid = tid >> 5;
    ListCandidates src{cand_box + base, cand_score + base, cand_cls + base, nullptr};
    int kept = 0;
    // Tier cut: only the first max_det detections are wanted and a box can only be suppressed by a better-scored
    // one, so the NMS first runs on the ~1.25 * max_det best candidates (every score above a 16-bit radix cut, ties
    // with the cut included).  max_det survivors there are the answer; otherwise everything is swept.  The
    // reference's branch rule and the offset trick's span are defined on ALL candidates: FullStats carries them.
    const int want_sub = cap_out + (cap_out >> 2) + 32;
    for (int tier = (cnt >= want_sub + (want_sub >> 1)) ? 0 : 1; tier < 2; ++tier) {
        int m = cnt;
        const FullStats* fs = ext_full ? ext_full + img : nullptr;
        if (tier == 0) {
            float mx = -INFINITY, mn = INFINITY;
            int fin = 1, maxcat = 0;
            uint32_t* skey = reinterpret_cast<uint32_t*>(sm.nms.sarea);
            for (int i = tid; i < cnt; i += T) {
                const float4 b = cand_box[base + i];
                mx = max_nan(mx, max_nan(max_nan(b.x, b.y), max_nan(b.z, b.w)));
                mn = min_nan(mn, min_nan(min_nan(b.x, b.y), min_nan(b.z, b.w)));
                fin &= (int)(isfinite(b.x) && isfinite(b.y) && isfinite(b.z) && isfinite(b.w));
                maxcat = max(maxcat, cand_cls[base + i]);
                skey[i] = score_desc_key(cand_score[base + i]);
            }
            for (int o = 16; o > 0; o >>= 1) {
                mx = max_nan(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                mn = min_nan(mn, __shfl_xor_sync(0xffffffffu, mn, o));
                fin &= __shfl_xor_sync(0xffffffffu, fin, o);
                maxcat = max(maxcat, __shfl_xor_sync(0xffffffffu, maxcat, o));
            }
            if (lane == 0) {
                sm.nms.red_max[wid] = mx;
                sm.nms.red_min[wid] = mn;
                sm.nms.red_flag[wid] = fin | (maxcat << 1);
            }
            if (tid == 0) {
                sm.sel_prefix = 0u;
                sm.sel_want = want_sub;
                sm.sub_count = 0;
            }
            __syncthreads();
            if (tid == 0) {
                float gmx = sm.nms.red_max[0], gmn = sm.nms.red_min[0];
                int gfin = sm.nms.red_flag[0] & 1, gcat = sm.nms.red_flag[0] >> 1;
                for (int w = 1; w < T / 32; ++w) {
                    gmx = max_nan(gmx, sm.nms.red_max[w]);
                    gmn = min_nan(gmn, sm.nms.red_min[w]);
                    gfin &= sm.nms.red_flag[w] & 1;
                    gcat = max(gcat, sm.nms.red_flag[w] >> 1);
                }
                sm.full.count = cnt; sm.full.mx = gmx; sm.full.mn = gmn; sm.full.fin = gfin; sm.full.maxcat = gcat;
            }
            // radix select on the two upper bytes of the descending-score keys: the bin of the want_sub-th best
            for (int shift = 24; shift >= 16; shift -= 8) {
                for (int b = tid; b < 256; b += T) sm.hist[b] = 0u;
                __syncthreads();
                const uint32_t prefix = sm.sel_prefix, himask = (shift == 24) ? 0u : 0xff000000u;
                for (int i = tid; i < cnt; i += T)
                    if ((skey[i] & himask) == prefix) atomicAdd(&sm.hist[(skey[i] >> shift) & 255u], 1u);
                __syncthreads();
                if (wid == 0) {
                    uint32_t c8[8], tot = 0;
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        c8[q] = sm.hist[lane * 8 + q];
                        tot += c8[q];
                    }
                    uint32_t incl = tot;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
                        incl += (lane >= o) ? up : 0u;
                    }
                    const uint32_t want = (uint32_t)sm.sel_want, before = incl - tot;
                    if (before < want && want <= incl) {  // exactly one lane (want <= number of matching keys)
                        uint32_t run = before;
                        int q = 0;
                        for (; q < 7 && run + c8[q] < want; ++q) run += c8[q];
                        sm.sel_prefix = prefix | ((uint32_t)(lane * 8 + q) << shift);
                        sm.sel_want = (int)(want - run);
                    }
                }
                __syncthreads();
            }
            const uint32_t cut = sm.sel_prefix | 0xffffu;  // every key of the chosen 16-bit bin and all better ones
            for (int i0 = 0; i0 < cnt; i0 += T) {
                const int i = i0 + tid;
                const bool in = i < cnt && skey[i] <= cut;
                const unsigned bal = __ballot_sync(0xffffffffu, in);
                int pos = 0;
                if (lane == 0 && bal) pos = atomicAdd(&sm.sub_count, __popc(bal));
                pos = __shfl_sync(0xffffffffu, pos, 0);
                if (in) sm.perm[pos + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)i;
            }
            __syncthreads();
            m = sm.sub_count;
            if (!ext_full) fs = &sm.full;  // external tier: the statistics of the whole image came with the list
            src.perm = sm.perm;
        } else {
            src.perm = nullptr;
        }
        // The list was filled in arrival order, the oracle's candidates come in row order (torch.nonzero).  The order
        // only matters where two scores tie, so the first attempt uses the list as it is and looks for ties: inside a
        // category segment (small_nms_body flags them) they may have changed the outcome and the image is redone with
        // the slots in row order (rare); ties of different segments only meet in the final list, where each run of
        // equal scores is put into row order by the thread that owns its first entry.
        const int npad = next_pow2(max(m, 2));
        for (bool presort = false;; presort = true) {
            if (presort) {
                for (int i = tid; i < npad; i += T) {
                    const int slot = (i < m) ? src.slot(i) : 0;
                    sm.nms.keys[i] = (i < m) ? (((uint64_t)(uint32_t)cand_id[base + slot] << kSmallIdxBits) | (uint64_t)slot)
                                             : kSentinelKey;
                }
                __syncthreads();
                cta_bitonic_sort<T>(sm.nms.keys, npad);
                for (int i = tid; i < m; i += T) sm.perm[i] = (uint16_t)(sm.nms.keys[i] & ((1u << kSmallIdxBits) - 1u));
                __syncthreads();
                src.perm = sm.perm;
            }
            DET_MARK(1);
            kept = small_nms_body<CAP, T>(sm.nms, src, m, thr_f, mode, cap_out, fs);
            if (presort) break;
            int redo = sm.nms.tie;
            if (!redo) {
                for (int j = tid; j + 1 < kept; j += T) {
                    const uint64_t sc = sm.nms.keys[j] >> KL::kScoreShift;
                    if ((sm.nms.keys[j + 1] >> KL::kScoreShift) != sc) continue;
                    if (j > 0 && (sm.nms.keys[j - 1] >> KL::kScoreShift) == sc) continue;  // not the head of the run
                    int len = 2;
                    while (j + len < kept && (sm.nms.keys[j + len] >> KL::kScoreShift) == sc) ++len;
                    if (len > 16) {
                        redo = 1;
                        break;
                    }
                    for (int a = 0; a + 1 < len; ++a)  // selection sort by row index; only this thread writes the run
                        for (int b = a + 1; b < len; ++b) {
                            const uint64_t ka = sm.nms.keys[j + a], kb = sm.nms.keys[j + b];
                            if (cand_id[base + src.slot((int)KL::idx(kb))] < cand_id[base + src.slot((int)KL::idx(ka))]) {
                                sm.nms.keys[j + a] = kb;
                                sm.nms.keys[j + b] = ka;
                            }
                        }
                }
            }
            if (!__syncthreads_or(redo)) break;
        }
        if (kept < 0 || kept >= cap_out) break;  // bad category, or the tier already holds max_det survivors
    }
    if (ext_full && kept >= 0 && kept < cap_out && cnt < ext_full[img].count) {
        if (tid == 0) todo[img] = 1;  // the best-scored subset fell short of max_out survivors: full path
        return;
    }
    const int nout = kept < 0 ? 0 : min(kept, cap_out);
    for (int j = tid; j < nout; j += T) {
        const int slot = src.slot((int)KL::idx(sm.nms.keys[j]));
        const int64_t o = (int64_t)img * max_det + j;
        det_idx[o] = (int64_t)cand_id[base + slot];
        if (det_boxes) det_boxes[o] = cand_box[base + slot];
        if (det_scores) det_scores[o] = cand_score[base + slot];
        if (det_classes) det_classes[o] = (int64_t)cand_cls[base + slot];
    }
    if (tid == 0) det_count[img] = kept < 0 ? -1 : nout;
    DET_MARK(14);
}

}  // namespace det
