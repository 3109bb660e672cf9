// Small path of the batched NMS: one CTA per image, everything (keys, sorted boxes, kept lists) in shared memory.
// Shared by nms.cu (generic boxes from HBM) and yolo.cu (candidates produced in shared memory by the decode).
//
// Phases (T threads, up to CAP candidates):
//   0  coordinate statistics for torchvision's offset trick (only when the image takes that branch)
//   1  keys (category | descending score | index), 2  register-resident bitonic sort
//   3  gather boxes into sorted order, discover category segments with an ordered block scan
//   4a tiny segments (<= 64 boxes): every pair once -> u64 suppression rows -> sparse greedy resolution
//   4b mid segments (<= 192): one warp each; 4c long segments: whole CTA (nms_core.cuh)
//   5  re-key the kept boxes by (descending score | index) and sort again -> output order
#pragma once
#include "nms_core.cuh"

namespace det {

constexpr int kSmallIdxBits = 12;
constexpr int kTinySegMax = 64;   // full pair bit-matrix (one u64 row per box)
constexpr int kWarpSegMax = 192;  // swept by a single warp (warps run in parallel)
constexpr int kCtaChunk = 256;    // chunk width of the CTA-wide sweep

template <int CAP, int T>
struct SmallSmem {
    uint64_t keys[CAP];
    float4 sbox[CAP];
    float sarea[CAP];
    uint16_t klist[CAP];  // mid/long segments: kept positions; tiny segments: start position of the box's segment
    uint16_t seg_s[CAP];  // all segments, in order (at most one per box)
    uint16_t seg_e[CAP];
    uint16_t mid_s[CAP / kTinySegMax], mid_e[CAP / kTinySegMax];
    uint16_t big_s[CAP / kWarpSegMax + 1], big_e[CAP / kWarpSegMax + 1];
    uint8_t state[CAP];
    uint8_t tiny_m[CAP];  // length of the tiny segment a position belongs to (0: not in a tiny segment)
    union {               // the tiny phase finishes before the CTA-wide phase starts
        uint64_t rowmask[CAP];
        uint32_t rowbits[kCtaChunk * (kCtaChunk / 32)];
    };
    uint32_t amask[kCtaChunk / 32];
    uint32_t deadmask[kCtaChunk / 32];
    float red_max[T / 32];
    float red_min[T / 32];
    int red_flag[T / 32];
    int warp_heads[T / 32];
    int nseg_small, nseg_large, nk_scratch, nkept, bad_cat;
    float span;
    int fast;
    int tie;  // two candidates of one segment share their score: their order was decided by the candidate index
    // offset trick with coordinates below -1: categories joined by an intersecting pair of shifted boxes
    int nnode, nedge;
    uint16_t node_cat[64], node_lab[64];
    uint32_t edges[64];  // (lower category << 16) | higher category
};

constexpr int kCrossMax = 256;  // offset trick: most boxes that may reach into a lower category's coordinate range
constexpr int kCrossNodes = 64;

// candidate source for boxes that live in HBM as (boxes, scores, categories) rows of one image
struct GlobalCandidates {
    const float4* boxes;
    const float* scores;
    const int64_t* cats;  // may be null: single category
    __device__ __forceinline__ float4 box(int i) const { return boxes[i]; }
    __device__ __forceinline__ float score(int i) const { return scores[i]; }
    __device__ __forceinline__ int64_t cat(int i) const { return cats ? cats[i] : 0; }
};

// Statistics of a superset of the candidates handed to small_nms_body (the caller sweeps only a best-scored subset):
// the reference's branch rule (count <= 1000 -> offset trick) and the trick's span are defined on the full set.
struct FullStats {
    int count;
    float mx, mn;
    int fin, maxcat;
};

// Greedy category-partitioned NMS of `cnt` (<= CAP) candidates by one CTA of T threads.
// Returns (block-uniform) the number of kept candidates, or -1 if a category is outside [0, 32767).
// On return sm.keys[0 .. kept) hold (descending-score bits | candidate index) in output order.
template <int CAP, int T, typename Src>
__device__ int small_nms_body(SmallSmem<CAP, T>& sm, const Src& src, int cnt, float thr_f, int mode, int max_out,
                              const FullStats* full = nullptr) {
    using KL = KeyLayout<kSmallIdxBits>;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr int W = T / 32;
    // reference CPU rule: boxes.numel() <= 4000 -> coordinate-offset trick (torchvision/ops/boxes.py batched_nms)
    const bool trick = (mode == DET_NMS_AUTO) ? ((full ? full->count : cnt) <= 1000) : (mode == DET_NMS_OFFSET_TRICK);
    if (tid == 0) {
        sm.nseg_small = 0;
        sm.nseg_large = 0;
        sm.nkept = 0;
        sm.bad_cat = 0;
        sm.span = 0.f;
        sm.fast = 1;
        sm.tie = 0;
        sm.nk_scratch = 0;
        sm.nnode = 0;
        sm.nedge = 0;
    }
    if (tid < kCrossNodes) sm.edges[tid] = 0u;  // 0 is no edge (lower < higher): a reserved, unwritten slot matches nothing
    __syncthreads();
    // ---- phase 0: coordinate statistics for the offset trick
    if (trick) {
        float mx = -INFINITY, mn = INFINITY;
        int fin = 1, maxcat = 0;
        for (int i = tid; i < cnt; i += T) {
            const float4 b = src.box(i);
            const int64_t c = src.cat(i);
            sm.sbox[i] = b;  // raw copies for the cross-category check below (phase 3 overwrites them)
            sm.seg_s[i] = (uint16_t)min(max(c, (int64_t)0), (int64_t)65535);
            mx = max_nan(mx, max_nan(max_nan(b.x, b.y), max_nan(b.z, b.w)));
            mn = min_nan(mn, min_nan(min_nan(b.x, b.y), min_nan(b.z, b.w)));
            fin &= (int)(isfinite(b.x) && isfinite(b.y) && isfinite(b.z) && isfinite(b.w));
            maxcat = max(maxcat, (int)c);
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
            if (full) {
                gmx = full->mx; gmn = full->mn; gfin = full->fin; gcat = full->maxcat;
            }
            const float span = gmx + 1.0f;  // max_coordinate + torch.tensor(1).to(boxes)
            sm.span = span;
            // categories can be swept independently iff shifted boxes of different categories cannot intersect:
            // all coordinates finite, shifted coordinates finite, non-negative threshold, and either every coordinate
            // > -1 (then category c lives in [c*span - 1, (c+1)*span - 1)) or -- checked next, fast == 2 -- no actual
            // pair of boxes of different categories with a positive intersection
            const float far = gmx + (float)gcat * span;
            const bool base = gfin && thr_f >= 0.0f && isfinite(far);
            sm.fast = !base ? 0 : (gmn > -1.0f) ? 1 : (gmx > -1.0f && gcat < 65535) ? 2 : 0;
            sm.red_max[0] = far;
        }
        __syncthreads();
        if (sm.fast == 2) {
            // A box Q of category q can reach a box P of a lower category only if BOTH its x1 and y1 lie below
            // -1 (+ a margin of ~30 ulp of the largest shifted coordinate for the fp32 roundings of the shifts):
            // otherwise fl(P.x2 + p*span) <= fl(Q.x1 + q*span) for every P (or the same in y).  Such boxes (top-left
            // corner, partly outside the frame) are few: test them against every box of a lower category with the
            // shifted coordinates exactly as phase 3 computes them.  Every intersecting pair joins its two categories;
            // the connected components of that graph are closed under suppression (boxes of different components
            // have IoU 0), so each component is swept as ONE segment and the result equals the sweep over everything.
            const float span = sm.span, lim = -1.0f + sm.red_max[0] * 4e-6f;
            for (int i = tid; i < cnt; i += T) {
                const float4 b = sm.sbox[i];
                if (sm.seg_s[i] >= 1 && b.x < lim && b.y < lim) {
                    const int slot = atomicAdd(&sm.nk_scratch, 1);
                    if (slot < kCrossMax) sm.klist[slot] = (uint16_t)i;
                }
            }
            __syncthreads();
            const int na = sm.nk_scratch;
            if (na <= kCrossMax) {
                for (int w = tid; w < na * cnt; w += T) {
                    const int a = w / cnt, i = w - a * cnt;
                    const int iq = (int)sm.klist[a];
                    const int cq = (int)sm.seg_s[iq], cp = (int)sm.seg_s[i];
                    if (cp >= cq) continue;
                    const float oq = (float)cq * span, op = (float)cp * span;
                    const float4 q = sm.sbox[iq], pb = sm.sbox[i];
                    const float ww = fminf(q.z + oq, pb.z + op) - fmaxf(q.x + oq, pb.x + op);
                    const float hh = fminf(q.w + oq, pb.w + op) - fmaxf(q.y + oq, pb.y + op);
                    if (ww > 0.0f && hh > 0.0f) {
                        const uint32_t e = ((uint32_t)cp << 16) | (uint32_t)cq;
                        bool seen = false;  // the same pair of categories is usually hit many times
                        const int ne = min(*(volatile int*)&sm.nedge, kCrossNodes);
                        for (int k = 0; k < ne; ++k) seen |= (((volatile uint32_t*)sm.edges)[k] == e);
                        if (!seen) {
                            const int slot = atomicAdd(&sm.nedge, 1);
                            if (slot < kCrossNodes) sm.edges[slot] = e;
                        }
                    }
                }
            }
            __syncthreads();
            if (tid == 0) {
                int ok = (na <= kCrossMax && sm.nedge <= kCrossNodes) ? 1 : 0;
                int nn = 0;
                for (int k = 0; ok && k < sm.nedge; ++k) {  // label propagation over a handful of edges
                    const int u = (int)(sm.edges[k] >> 16), v = (int)(sm.edges[k] & 0xffffu);
                    int lu = -1, lv = -1;
                    for (int q = 0; q < nn; ++q) {
                        if (sm.node_cat[q] == u) lu = sm.node_lab[q];
                        if (sm.node_cat[q] == v) lv = sm.node_lab[q];
                    }
                    if (lu < 0) {
                        if (nn == kCrossNodes) { ok = 0; break; }
                        sm.node_cat[nn] = (uint16_t)u; sm.node_lab[nn] = (uint16_t)u; lu = u; ++nn;
                    }
                    if (lv < 0) {
                        if (nn == kCrossNodes) { ok = 0; break; }
                        sm.node_cat[nn] = (uint16_t)v; sm.node_lab[nn] = (uint16_t)v; lv = v; ++nn;
                    }
                    if (lu != lv) {
                        const int lo = min(lu, lv), hi = max(lu, lv);
                        for (int q = 0; q < nn; ++q)
                            if (sm.node_lab[q] == hi) sm.node_lab[q] = (uint16_t)lo;
                    }
                }
                sm.nnode = ok ? nn : 0;
                sm.fast = ok;
                sm.nk_scratch = 0;
            }
            __syncthreads();
        }
    }
    DET_MARK(4);
    const float span = sm.span;
    const bool by_cat = !trick || sm.fast;
    const int nnode = (trick && sm.fast) ? sm.nnode : 0;
    // ---- phase 1+2: keys, sort
    const int npad = next_pow2(max(cnt, 2));
    for (int i = tid; i < npad; i += T) {
        uint64_t k = kSentinelKey;
        if (i < cnt) {
            const int64_t c = src.cat(i);
            if (c < 0 || c >= (1 << kSegBits) - 1) sm.bad_cat = 1;
            uint32_t seg = by_cat ? (uint32_t)(c & ((1 << kSegBits) - 1)) : 0u;
            for (int q = 0; q < nnode; ++q)  // categories joined by an intersecting pair share a segment
                if ((uint32_t)sm.node_cat[q] == seg) seg = (uint32_t)sm.node_lab[q];
            k = KL::make(seg, src.score(i), (uint32_t)i);
        }
        sm.keys[i] = k;
    }
    __syncthreads();
    DET_MARK(5);
    cta_bitonic_sort<T>(sm.keys, npad);
    DET_MARK(6);
    // ---- phase 3: boxes in sorted order (+ offset); segment heads numbered in order with a block scan, so every
    //      segment learns its end from the next head (no search)
    bool clean = true;  // no NaN coordinate seen by this thread
    int heads_before = 0;
    for (int p0 = 0; p0 < cnt; p0 += T) {
        const int p = p0 + tid;
        bool head = false;
        if (p < cnt) {
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
            sm.tiny_m[p] = 0;
            sm.rowmask[p] = 0ull;
            clean &= (b.x == b.x) && (b.y == b.y) && (b.z == b.z) && (b.w == b.w);
            head = (p == 0) || (KL::seg(sm.keys[p - 1]) != KL::seg(k));
            if (!head && ((sm.keys[p - 1] ^ k) >> KL::kScoreShift) == 0) sm.tie = 1;
        }
        const unsigned hb = __ballot_sync(0xffffffffu, head);
        if (lane == 0) sm.warp_heads[wid] = __popc(hb);
        __syncthreads();
        int before = heads_before, total = 0;
        for (int w = 0; w < W; ++w) {
            const int c = sm.warp_heads[w];
            if (w < wid) before += c;
            total += c;
        }
        if (head) {
            const int slot = before + __popc(hb & ((1u << lane) - 1u));
            sm.seg_s[slot] = (uint16_t)p;
            if (slot > 0) sm.seg_e[slot - 1] = (uint16_t)p;
        }
        heads_before += total;
        __syncthreads();
    }
    const int nseg = heads_before;
    if (tid == 0 && nseg > 0) sm.seg_e[nseg - 1] = (uint16_t)cnt;
    const bool nonan = __syncthreads_and(clean ? 1 : 0) != 0;
    DET_MARK(7);
    // classify: tiny segments are flagged through tiny_m, longer ones go to the mid/big work lists
    for (int sidx = tid; sidx < nseg; sidx += T) {
        const int s0 = (int)sm.seg_s[sidx], e0 = (int)sm.seg_e[sidx], m = e0 - s0;
        if (m <= kTinySegMax) {
            for (int q = s0; q < e0; ++q) {
                sm.klist[q] = (uint16_t)s0;
                sm.tiny_m[q] = (uint8_t)m;
            }
        } else if (m <= kWarpSegMax) {
            const int slot = atomicAdd(&sm.nseg_small, 1);
            sm.mid_s[slot] = (uint16_t)s0;
            sm.mid_e[slot] = (uint16_t)e0;
        } else {
            const int slot = atomicAdd(&sm.nseg_large, 1);
            sm.big_s[slot] = (uint16_t)s0;
            sm.big_e[slot] = (uint16_t)e0;
        }
    }
    __syncthreads();
    DET_MARK(8);
    // ---- phase 4a: tiny segments (the common case of per-class NMS): every unordered pair of a segment is tested
    //      exactly once with full lane utilisation -- row r meets row (r+d) mod m for d = 1..m/2 -- and a hit sets one
    //      bit of the earlier box's u64 row; the greedy order is then resolved on the bit rows alone.
    for (int p = tid; p < cnt; p += T) {
        const int m = (int)sm.tiny_m[p];
        if (m < 2) continue;
        const int s0 = (int)sm.klist[p], r = p - s0, half = m >> 1;
        // distance-m/2 pairs (even m) are met from the lower half only
        const int dmax = (((m & 1) == 0) && r >= half) ? half - 1 : half;
        const float4 mb = sm.sbox[p];
        const float ma = sm.sarea[p];
        const float4* sb = sm.sbox + s0;
        const float* sa = sm.sarea + s0;
        unsigned long long* rows = reinterpret_cast<unsigned long long*>(sm.rowmask + s0);
#pragma unroll 4
        for (int d = 1; d <= dmax; ++d) {
            int j = r + d;
            j = (j >= m) ? j - m : j;
            const float4 ob = sb[j];
            const float oa = sa[j];
            const bool fwd = j > r;  // the earlier box of the pair is the potential suppressor
            const float4 ka = fwd ? mb : ob, kb = fwd ? ob : mb;
            const float kaa = fwd ? ma : oa, kba = fwd ? oa : ma;
            const bool hit = nonan ? nms_suppresses<true>(ka, kaa, kb, kba, thr_f)
                                   : nms_suppresses<false>(ka, kaa, kb, kba, thr_f);
            if (hit) {
                const int lo_r = fwd ? r : j, hi_r = fwd ? j : r;
                atomicOr(rows + lo_r, 1ull << hi_r);
                // the last row of a segment is always empty (nothing comes after it): its slot collects the set of
                // non-empty rows, which is all the resolution step has to visit
                if (lo_r != m - 1) atomicOr(rows + (m - 1), 1ull << lo_r);
            }
        }
    }
    __syncthreads();
    DET_MARK(9);
    // resolution: only boxes whose row is non-zero can change the survivor set, so the greedy sweep visits just
    // those (in order); the survivor mask of the segment is left in the row slot of its first box.
    int tkept = 0;
    for (int sidx = tid; sidx < nseg; sidx += T) {
        const int s0 = (int)sm.seg_s[sidx], m = (int)sm.seg_e[sidx] - s0;
        if (m > kTinySegMax) continue;
        uint64_t alive = (m >= 64) ? ~0ull : ((1ull << m) - 1ull);
        uint64_t nz = sm.rowmask[s0 + m - 1];
        while (nz) {
            const int r = __ffsll((long long)nz) - 1;
            nz &= nz - 1;
            if ((alive >> r) & 1ull) alive &= ~sm.rowmask[s0 + r];
        }
        int nk = __popcll(alive);
        while (nk > max_out) {  // keep only the first max_out survivors: drop the highest set bits
            alive &= ~(1ull << (63 - __clzll((long long)alive)));
            --nk;
        }
        sm.rowmask[s0] = alive;
        tkept += nk;
    }
    tkept = warp_sum(tkept);
    if (lane == 0 && tkept) atomicAdd(&sm.nkept, tkept);
    __syncthreads();
    DET_MARK(10);
    // ---- phase 4b/4c: mid segments one warp each (in parallel), long ones by the whole CTA
    const int nsmall = sm.nseg_small, nlarge = sm.nseg_large;
    if (nsmall) {
        int mykept = 0;
        for (int sidx = wid; sidx < nsmall; sidx += W) {
            const int s0 = (int)sm.mid_s[sidx], e0 = (int)sm.mid_e[sidx];
            mykept += nonan ? warp_segment_nms<uint16_t, true>(sm.sbox, sm.sarea, sm.state, sm.klist, s0, e0, thr_f, max_out)
                            : warp_segment_nms<uint16_t, false>(sm.sbox, sm.sarea, sm.state, sm.klist, s0, e0, thr_f, max_out);
        }
        if (lane == 0 && mykept) atomicAdd(&sm.nkept, mykept);
        __syncthreads();
    }
    if (nlarge) {  // the CTA-wide sweep reuses the row-mask storage: move the tiny segments' survivors into state[]
        for (int p = tid; p < cnt; p += T) {
            const int m = (int)sm.tiny_m[p];
            if (m) {
                const int s0 = (int)sm.klist[p];
                sm.state[p] = ((sm.rowmask[s0] >> (p - s0)) & 1ull) ? 2 : 0;
            }
        }
        __syncthreads();
        for (int p = tid; p < cnt; p += T) sm.tiny_m[p] = 0;
        __syncthreads();
    }
    for (int sidx = 0; sidx < nlarge; ++sidx) {
        const int s0 = (int)sm.big_s[sidx], e0 = (int)sm.big_e[sidx];
        const int nk = nonan ? cta_segment_nms<kCtaChunk, uint16_t, true>(sm.sbox, sm.sarea, sm.state, sm.klist, s0, e0,
                                                                          thr_f, max_out, sm.rowbits, sm.amask,
                                                                          sm.deadmask, &sm.nk_scratch)
                             : cta_segment_nms<kCtaChunk, uint16_t, false>(sm.sbox, sm.sarea, sm.state, sm.klist, s0, e0,
                                                                           thr_f, max_out, sm.rowbits, sm.amask,
                                                                           sm.deadmask, &sm.nk_scratch);
        if (tid == 0) sm.nkept += nk;
        __syncthreads();
    }
    const int kept = sm.nkept;
    DET_MARK(11);
    // ---- phase 5: output order = (descending score, index) over the kept candidates
    for (int p = tid; p < npad; p += T) {
        uint64_t k = kSentinelKey;
        if (p < cnt) {
            const int m = (int)sm.tiny_m[p];
            bool is_kept;
            if (m) {
                const int s0 = (int)sm.klist[p];
                is_kept = (sm.rowmask[s0] >> (p - s0)) & 1ull;
            } else {
                is_kept = sm.state[p] == 2;
            }
            if (is_kept) k = KL::strip_seg(sm.keys[p]);
        }
        sm.keys[p] = k;
    }
    __syncthreads();
    DET_MARK(12);
    cta_bitonic_sort<T>(sm.keys, npad);
    DET_MARK(13);
    return sm.bad_cat ? -1 : kept;
}

}  // namespace det
