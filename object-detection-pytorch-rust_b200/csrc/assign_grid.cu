// Training side for GRID anchors (subsystem 4a on the layout AnchorGenerator really produces): IoU target assignment in
// one streaming pass, plus the sampled (sparse) loss forward + backward that reads the head where it lies (NCHW).
//
// Reference behaviour reproduced (paths relative to the reference root):
//   python/src/models/modules/anchor_generators.py:158-179   anchors of a level = shifts (h, w) + cell anchors, order (h, w, a)
//   python/src/models/rpn.py:161-168                           per image: pairwise_iou(gt, anchors) -> Matcher
//   python/src/models/components/matcher.py:53-120             column max / argmax, threshold buckets, low-quality promotion
//   python/src/models/rpn.py:187-244, components/box_regression.py:128-168   losses on the sampled anchors
//
// What the grid buys (the generic kernels of assign_loss.cu brute-force every anchor against every gt box, cull by the
// bounding box of 128 CONSECUTIVE anchors -- a 170-pixel strip of one feature-map row -- and need a second pass for the
// low-quality rule):
//   * match_rowmax_kernel (gt-centric): the anchors that can overlap a gt box form a closed-form (y, x) window per level
//     and cell anchor, so the row maximum of `pairwise_iou` is evaluated on that window only, and a whole (level, cell
//     anchor) is skipped once min(area)/max(area) -- an upper bound of its IoUs -- is below the running maximum.
//   * match_grid_kernel (anchor-centric): a warp owns an 8 x 4 TILE of positions (all A anchors of a position in one
//     lane); its bounding box is compact in both directions, so about half as many (gt, tile) pairs survive the cull as
//     with row strips.  The row maxima are already final, hence the low-quality promotion `IoU == row max`
//     (matcher.py:110-120, exact fp32 equality of the SAME pair_iou evaluation) is decided in the same pass: labels and
//     matched indices are written once, and there is no per-warp reduction, no atomic on row maxima and no second pass.
//     The kernel also counts positives / ignored anchors per image and lists the positives, which lets the sampler and
//     the loss work on O(sampled) instead of O(R) data.
// Anchor coordinates are always READ from the caller's (R, 4) table (802 KB, L2-resident): every IoU is the same fp32
// expression on the same bits as in the generic kernels, the grid layout only organises the work.
#include "common.cuh"
#include "peer.cuh"
#include "assign.cuh"

namespace det {

constexpr int kGridMaxLevels = 8;
constexpr int kTileW = 8, kTileH = 4;  // positions per warp tile (8 wide x 4 tall = one lane per position)
constexpr int kGridWarps = 8;          // warps (tiles) per CTA

struct GridLevelDev {
    int h, w, stride, tiles_x;
    int first_tile;
    int64_t first_row;
};
struct GridLayoutDev {
    GridLevelDev lv[kGridMaxLevels];
    int nlev, a, tiles_total;
};

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---- K1: row maximum of pairwise_iou(gt, anchors) for every gt box, one warp per box ----------------------------------
// stats (4 int32 per image: #positives, #ignored, 2 spare) are zeroed here for match_grid_kernel, which follows on the
// same stream (saves a memset launch).
__global__ void __launch_bounds__(256)
match_rowmax_kernel(const float4* __restrict__ gt, int64_t sum_g, const float4* __restrict__ anchors, GridLayoutDev lay,
                    float* __restrict__ rowmax, int32_t* __restrict__ stats, int stats_words) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (stats && tid < stats_words) stats[tid] = 0;
    const int64_t t = tid >> 5;
    if (t >= sum_g) return;
    const int lane = threadIdx.x & 31;
    const float4 gb = gt[t];
    const float ga = box_area(gb);
    const bool finite = isfinite(gb.x) && isfinite(gb.y) && isfinite(gb.z) && isfinite(gb.w);
    const int A = lay.a, P = lay.nlev * A;  // (level, cell anchor) pairs, P <= 32: lane p looks after pair p
    float4 ab0 = make_float4(0.f, 0.f, 0.f, 0.f);
    float bound = -1.0f;
    if (lane < P) {
        ab0 = anchors[lay.lv[lane / A].first_row + lane % A];  // the pair's anchor at position (0, 0)
        const float aa = box_area(ab0);
        // IoU <= min(area) / max(area); anything doubtful (degenerate or non-finite boxes) is never pruned
        bound = (finite && ga > 0.0f && aa > 0.0f && isfinite(aa)) ? fminf(ga, aa) / fmaxf(ga, aa) : INFINITY;
    }
    float best = 0.0f;
    unsigned pending = __ballot_sync(0xffffffffu, lane < P);
    while (pending) {
        // the pending pair with the largest bound
        const float mine = ((pending >> lane) & 1u) ? bound : -1.0f;
        const float bmax = warp_max(mine);
        if (bmax * 1.0001f < best) break;  // no remaining pair can reach (let alone equal) the maximum found so far
        const int src = __ffs(__ballot_sync(0xffffffffu, mine == bmax)) - 1;
        pending &= ~(1u << src);
        const float4 a0 = make_float4(__shfl_sync(0xffffffffu, ab0.x, src), __shfl_sync(0xffffffffu, ab0.y, src),
                                      __shfl_sync(0xffffffffu, ab0.z, src), __shfl_sync(0xffffffffu, ab0.w, src));
        const GridLevelDev L = lay.lv[src / A];
        const int a = src % A;
        // anchor at (y, x) = a0 + (x, y) * stride up to rounding: it can only intersect gb if
        //   a0.x + x s < gb.z  and  a0.z + x s > gb.x  (same in y); one position of margin on each side
        int x_lo = 0, x_hi = L.w - 1, y_lo = 0, y_hi = L.h - 1;
        if (finite) {
            const float s = (float)L.stride;
            const float fx_lo = floorf((gb.x - a0.z) / s) - 1.0f, fx_hi = ceilf((gb.z - a0.x) / s) + 1.0f;
            const float fy_lo = floorf((gb.y - a0.w) / s) - 1.0f, fy_hi = ceilf((gb.w - a0.y) / s) + 1.0f;
            x_lo = (int)fminf(fmaxf(fx_lo, 0.0f), (float)L.w);
            x_hi = (int)fmaxf(fminf(fx_hi, (float)(L.w - 1)), -1.0f);
            y_lo = (int)fminf(fmaxf(fy_lo, 0.0f), (float)L.h);
            y_hi = (int)fmaxf(fminf(fy_hi, (float)(L.h - 1)), -1.0f);
        }
        const int wx = x_hi - x_lo + 1, wy = y_hi - y_lo + 1;
        float v = 0.0f;
        if (wx > 0 && wy > 0) {
            const int nwin = wx * wy;
            for (int k = lane; k < nwin; k += 32) {
                const int y = y_lo + k / wx, x = x_lo + k % wx;
                const float4 ab = anchors[L.first_row + ((int64_t)y * L.w + x) * A + a];
                v = fmaxf(v, pair_iou(gb, ga, ab, box_area(ab)));  // same expression as match_grid_kernel
            }
        }
        best = fmaxf(best, warp_max(v));
    }
    if (lane == 0) rowmax[t] = best;
}

// ---- K2: labels + matched index of every anchor, one pass -------------------------------------------------------------
struct TileBox {
    float x1, y1, x2, y2;
};
__device__ __forceinline__ bool tile_culls(const TileBox& bb, const float4 g) {
    // true iff the gt box certainly has zero intersection with every anchor of the tile (NaN: never culled)
    return g.z <= bb.x1 || g.x >= bb.x2 || g.w <= bb.y1 || g.y >= bb.y2;
}

template <int A, bool DENSE>
__global__ void __launch_bounds__(kGridWarps * 32)
match_grid_kernel(const float4* __restrict__ gt, const int32_t* __restrict__ gt_off, const float4* __restrict__ anchors,
                  int n, int imgs, int64_t r, GridLayoutDev lay, MatchRule rule, const float* __restrict__ rowmax,
                  int64_t* __restrict__ matched, int8_t* __restrict__ labels, float* __restrict__ matched_iou,
                  int32_t* __restrict__ stats, int32_t* __restrict__ pos_list, int list_cap) {
    const unsigned FULLMASK = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int T = blockIdx.x * kGridWarps + (threadIdx.x >> 5);
    if (T >= lay.tiles_total) return;  // warp-uniform
    int l = 0;
#pragma unroll
    for (int k = 1; k < kGridMaxLevels; ++k)
        if (k < lay.nlev && T >= lay.lv[k].first_tile) l = k;
    const GridLevelDev L = lay.lv[l];
    const int tl = T - L.first_tile;
    const int y = (tl / L.tiles_x) * kTileH + (lane >> 3), x = (tl % L.tiles_x) * kTileW + (lane & 7);
    const bool valid = y < L.h && x < L.w;
    const int64_t row0 = L.first_row + ((int64_t)y * L.w + x) * A;  // this lane's A anchors are consecutive rows
    float4 ab[A];
    float aa[A];
    TileBox bb{INFINITY, INFINITY, -INFINITY, -INFINITY};
#pragma unroll
    for (int a = 0; a < A; ++a) {
        ab[a] = valid ? anchors[row0 + a] : make_float4(0.f, 0.f, 0.f, 0.f);
        aa[a] = box_area(ab[a]);
        if (valid) {  // fminf / fmaxf drop NaN coordinates: a NaN anchor intersects nothing
            bb.x1 = fminf(bb.x1, ab[a].x); bb.y1 = fminf(bb.y1, ab[a].y);
            bb.x2 = fmaxf(bb.x2, ab[a].z); bb.y2 = fmaxf(bb.y2, ab[a].w);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        bb.x1 = fminf(bb.x1, __shfl_xor_sync(FULLMASK, bb.x1, o)); bb.y1 = fminf(bb.y1, __shfl_xor_sync(FULLMASK, bb.y1, o));
        bb.x2 = fmaxf(bb.x2, __shfl_xor_sync(FULLMASK, bb.x2, o)); bb.y2 = fmaxf(bb.y2, __shfl_xor_sync(FULLMASK, bb.y2, o));
    }
    const int i0 = blockIdx.y * imgs, ni = min(imgs, n - i0);
    for (int ii = 0; ii < ni; ++ii) {
        const int img = i0 + ii;
        const int g0 = gt_off[img], G = gt_off[img + 1] - g0;
        float best[A];
        int bidx[A];
        bool hit[A];
#pragma unroll
        for (int a = 0; a < A; ++a) {
            best[a] = 0.0f;  // IoUs are >= 0 and only a strictly larger one replaces the incumbent: gt 0 wins ties at 0,
            bidx[a] = 0;     // exactly like torch.max(dim=0)
            hit[a] = false;
        }
        bool promote_all = false;  // a gt whose row maximum is 0 promotes EVERY anchor (`Q == rowmax` holds everywhere)
        for (int t0 = 0; t0 < G; t0 += 32) {
            float4 mine = make_float4(0.f, 0.f, 0.f, 0.f);
            float rm_mine = 1.0f;
            bool over = false;
            if (t0 + lane < G) {
                mine = gt[g0 + t0 + lane];
                over = !tile_culls(bb, mine);
                if (rule.allow_lq) rm_mine = rowmax[g0 + t0 + lane];
            }
            if (rule.allow_lq) promote_all |= __any_sync(FULLMASK, rm_mine == 0.0f);
            unsigned todo = __ballot_sync(FULLMASK, over);
            while (todo) {  // ascending gt index: the first maximum wins
                const int src = __ffs(todo) - 1;
                todo &= todo - 1;
                const float4 gb = make_float4(__shfl_sync(FULLMASK, mine.x, src), __shfl_sync(FULLMASK, mine.y, src),
                                              __shfl_sync(FULLMASK, mine.z, src), __shfl_sync(FULLMASK, mine.w, src));
                const float rm = __shfl_sync(FULLMASK, rm_mine, src);
                const float ga = box_area(gb);
                const int t = t0 + src;
#pragma unroll
                for (int a = 0; a < A; ++a) {
                    // pairwise_iou(gt, anchors): boxes1 = gt, boxes2 = anchors (rpn.py:167)
                    const float v = valid ? pair_iou(gb, ga, ab[a], aa[a]) : 0.0f;
                    if (v > best[a]) {
                        best[a] = v;
                        bidx[a] = t;
                    }
                    hit[a] |= (v == rm);  // matcher.py:110-120; culled pairs have v = 0 and only match rm = 0 (promote_all)
                }
            }
        }
        int8_t lab[A];
        int npos = 0, nign = 0;
#pragma unroll
        for (int a = 0; a < A; ++a) {
            lab[a] = (G == 0) ? rule.lab[0]  // matcher.py:67-77: no gt -> match 0, label labels[0]
                              : ((rule.allow_lq && (hit[a] || promote_all)) ? (int8_t)1 : bucket_label(rule, best[a]));
            npos += valid && lab[a] != 0 && lab[a] != -1;
            nign += valid && lab[a] == -1;
        }
        if (DENSE && valid) {
            const int64_t o = (int64_t)img * r + row0;
#pragma unroll
            for (int a = 0; a < A; ++a) {
                matched[o + a] = bidx[a];
                labels[o + a] = lab[a];
                if (matched_iou) matched_iou[o + a] = best[a];
            }
        }
        if (stats) {
            // warp-aggregated: one atomic per warp and image for each counter that is non-zero (both are rare)
            const unsigned pm = __ballot_sync(FULLMASK, npos > 0);
            const int tot_ign = warp_sum(nign);
            if (pm) {
                int incl = npos;  // inclusive scan of the per-lane positive counts
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int up = __shfl_up_sync(FULLMASK, incl, o);
                    if (lane >= o) incl += up;
                }
                const int total = __shfl_sync(FULLMASK, incl, 31);
                int base = 0;
                if (lane == 0) base = atomicAdd(&stats[img * 4 + 0], total);
                base = __shfl_sync(FULLMASK, base, 0) + incl - npos;
                if (pos_list) {
#pragma unroll
                    for (int a = 0; a < A; ++a)
                        if (valid && lab[a] != 0 && lab[a] != -1) {
                            if (base < list_cap)
                                pos_list[(int64_t)img * list_cap + base] = (int)(row0 + a) | ((int)(uint8_t)lab[a] << 24);
                            ++base;
                        }
                }
            }
            if (lane == 0 && tot_ign) atomicAdd(&stats[img * 4 + 1], tot_ign);
        }
    }
}

// ---- sampled loss: BCE-with-logits on the sampled anchors + smooth-L1 / GIoU on the sampled positives, forward and
//      backward, O(samples) traffic.  One CTA per image walks the image's sample list (<= num_samples entries written by
//      the sampler: anchor row | label << 24).  The head is read where the convolutions left it: either flat (N,R) /
//      (N,R,4) tensors or per-level NCHW planes (objectness (N,A,H,W), deltas (N,A*4,H,W)), gradients go to the same
//      layout.  Gradient buffers are NOT swept: the caller zeroes them once (cudaMemset) or keeps them persistent and
//      passes the previous step's sample list, whose entries are reset to zero first (`clear_*`).
//      The last CTA scales the sums, writes the 8-float result and re-arms the accumulators: no memset, no epilogue op.
struct HeadLevelDev {
    const float* obj;
    const float* dlt;
    float* g_obj;
    float* g_dlt;
    int h, w;
    int64_t first_row;
};
struct HeadLayoutDev {
    HeadLevelDev lv[kGridMaxLevels];
    int nlev;  // 0: flat layout, lv[0] holds the (N,R) / (N,R,4) pointers
    int a;
};

struct HeadAddr {
    int64_t obj;        // element offset of the logit
    int64_t dlt;        // element offset of delta component 0
    int64_t dlt_step;   // element stride between the 4 delta components
    int lvl;
};

__device__ __forceinline__ HeadAddr head_addr(const HeadLayoutDev& hl, int img, int64_t r, int64_t j) {
    HeadAddr ad;
    if (hl.nlev == 0) {
        ad.lvl = 0;
        ad.obj = (int64_t)img * r + j;
        ad.dlt = ad.obj * 4;
        ad.dlt_step = 1;
        return ad;
    }
    int l = 0;
#pragma unroll
    for (int k = 1; k < kGridMaxLevels; ++k)
        if (k < hl.nlev && j >= hl.lv[k].first_row) l = k;
    const HeadLevelDev& L = hl.lv[l];
    const int64_t local = j - L.first_row;
    const int a = (int)(local % hl.a);
    const int64_t pos = local / hl.a, hw = (int64_t)L.h * L.w;
    ad.lvl = l;
    ad.obj = ((int64_t)img * hl.a + a) * hw + pos;              // (N, A, H, W)
    ad.dlt = ((int64_t)img * hl.a * 4 + a * 4) * hw + pos;      // (N, A*4, H, W): channel = a*4 + component
    ad.dlt_step = hw;
    return ad;
}

constexpr int kSampledThreads = 128;

template <bool GIOU>
__global__ void __launch_bounds__(kSampledThreads)
rpn_loss_sampled_kernel(HeadLayoutDev hl, const int32_t* __restrict__ samples, const int32_t* __restrict__ sample_count,
                        int sample_cap, const int32_t* __restrict__ clear_samples,
                        const int32_t* __restrict__ clear_count, const int64_t* __restrict__ matched,
                        const float4* __restrict__ gt, const int32_t* __restrict__ gt_off,
                        const float4* __restrict__ anchors, int64_t r, CodecW wt, float scale_clamp, float beta,
                        float scale_cls, float scale_loc, const float* __restrict__ upstream, float* __restrict__ acc,
                        int32_t* __restrict__ ticket, float* __restrict__ sums_out, int write_grads) {
    __shared__ float s_part[4][kSampledThreads / 32];
    __shared__ float s_fin[8];
    const int img = blockIdx.x;
    float gs_cls = scale_cls, gs_loc = scale_loc;
    if (upstream) {
        gs_cls *= upstream[0];
        gs_loc *= upstream[1];
    }
    if (write_grads && clear_samples) {  // persistent gradient buffers: undo the previous step's writes of this image
        const int nc = min(clear_count[img], sample_cap);
        for (int k = threadIdx.x; k < nc; k += kSampledThreads) {
            const int packed = clear_samples[(int64_t)img * sample_cap + k];
            const HeadAddr ad = head_addr(hl, img, r, packed & 0xffffff);
            const HeadLevelDev& L = hl.lv[ad.lvl];
            L.g_obj[ad.obj] = 0.0f;
            if ((int8_t)(packed >> 24) == 1) {
                if (ad.dlt_step == 1) {
                    *reinterpret_cast<float4*>(L.g_dlt + ad.dlt) = make_float4(0.f, 0.f, 0.f, 0.f);
                } else {
#pragma unroll
                    for (int c = 0; c < 4; ++c) L.g_dlt[ad.dlt + c * ad.dlt_step] = 0.0f;
                }
            }
        }
        __syncthreads();  // the same anchor may be sampled again: zero first, then write
    }
    float acc_cls = 0.f, acc_loc = 0.f;
    int npos = 0, nneg = 0;
    const int ns = min(sample_count[img], sample_cap);
    for (int k = threadIdx.x; k < ns; k += kSampledThreads) {
        const int packed = samples[(int64_t)img * sample_cap + k];
        const int64_t j = packed & 0xffffff;
        const int lab = (int)(int8_t)(packed >> 24);
        const HeadAddr ad = head_addr(hl, img, r, j);
        const HeadLevelDev& L = hl.lv[ad.lvl];
        const float x = L.obj[ad.obj];
        const float yv = (float)lab;
        // BCE with logits: (1-y)*x - log_sigmoid(x), log_sigmoid(x) = min(x,0) - log1p(exp(-|x|))  (rpn.py:233)
        const float ls = fminf(x, 0.f) - log1pf(expf(-fabsf(x)));
        acc_cls += (1.f - yv) * x - ls;
        npos += lab == 1;
        nneg += lab == 0;
        if (write_grads) L.g_obj[ad.obj] = (1.f / (1.f + expf(-x)) - yv) * gs_cls;
        if (lab == 1) {
            const float4 g = gt[gt_off[img] + matched[(int64_t)img * r + j]];
            float4 p;
            if (ad.dlt_step == 1) {
                p = *reinterpret_cast<const float4*>(L.dlt + ad.dlt);
            } else {
                p = make_float4(L.dlt[ad.dlt], L.dlt[ad.dlt + ad.dlt_step], L.dlt[ad.dlt + 2 * ad.dlt_step],
                                L.dlt[ad.dlt + 3 * ad.dlt_step]);
            }
            float4 gd;
            if (GIOU) {
                float4 dd;
                acc_loc += giou_fwd_bwd(p, anchors[j], g, wt, scale_clamp, dd);
                gd = make_float4(dd.x * gs_loc, dd.y * gs_loc, dd.z * gs_loc, dd.w * gs_loc);
            } else {
                const float4 tgt = encode_target(anchors[j], g, wt);
                float v, d;
                smooth_l1(p.x, tgt.x, beta, v, d); acc_loc += v; gd.x = d * gs_loc;
                smooth_l1(p.y, tgt.y, beta, v, d); acc_loc += v; gd.y = d * gs_loc;
                smooth_l1(p.z, tgt.z, beta, v, d); acc_loc += v; gd.z = d * gs_loc;
                smooth_l1(p.w, tgt.w, beta, v, d); acc_loc += v; gd.w = d * gs_loc;
            }
            if (write_grads) {
                if (ad.dlt_step == 1) {
                    *reinterpret_cast<float4*>(L.g_dlt + ad.dlt) = gd;
                } else {
                    L.g_dlt[ad.dlt] = gd.x;
                    L.g_dlt[ad.dlt + ad.dlt_step] = gd.y;
                    L.g_dlt[ad.dlt + 2 * ad.dlt_step] = gd.z;
                    L.g_dlt[ad.dlt + 3 * ad.dlt_step] = gd.w;
                }
            }
        }
    }
    acc_cls = warp_sum(acc_cls);
    acc_loc = warp_sum(acc_loc);
    const float fpos = warp_sum((float)npos), fneg = warp_sum((float)nneg);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) {
        s_part[0][wid] = acc_cls;
        s_part[1][wid] = acc_loc;
        s_part[2][wid] = fpos;
        s_part[3][wid] = fneg;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        float t = 0.f;
        for (int w = 0; w < kSampledThreads / 32; ++w) t += s_part[threadIdx.x][w];
        if (t != 0.f) atomicAdd(&acc[threadIdx.x], t);
    }
    // last CTA: sums_out = [cls * scale_cls, loc * scale_loc, #pos, #neg, 0...], accumulators and ticket re-armed
    if (last_cta_arrives(ticket))
        finalize_sums(acc, sums_out, s_fin, [&](int i) { return i == 0 ? scale_cls : (i == 1 ? scale_loc : 1.0f); });
}

static int fill_layout(GridLayoutDev& lay, const det_anchor_level_t* levels_host, int num_levels, int a, int64_t r) {
    if (!levels_host || num_levels < 1 || num_levels > kGridMaxLevels || a < 1 || num_levels * a > 32) {
        set_error("grid layout: 1..%d levels, levels x anchors-per-position <= 32", kGridMaxLevels);
        return DET_ERR_UNSUPPORTED;
    }
    lay.nlev = num_levels;
    lay.a = a;
    int tiles = 0;
    int64_t row = 0;
    for (int l = 0; l < num_levels; ++l) {
        const det_anchor_level_t& s = levels_host[l];
        if (s.h < 1 || s.w < 1 || s.stride < 1 || s.first_row != row) {
            set_error("grid layout: level %d must have h, w, stride >= 1 and start at row %lld", l, (long long)row);
            return DET_ERR_BAD_ARG;
        }
        GridLevelDev& d = lay.lv[l];
        d.h = s.h; d.w = s.w; d.stride = s.stride;
        d.tiles_x = (s.w + kTileW - 1) / kTileW;
        d.first_tile = tiles;
        d.first_row = row;
        tiles += d.tiles_x * ((s.h + kTileH - 1) / kTileH);
        row += (int64_t)s.h * s.w * a;
    }
    if (row != r) {
        set_error("grid layout: levels cover %lld anchors, r = %lld", (long long)row, (long long)r);
        return DET_ERR_BAD_ARG;
    }
    lay.tiles_total = tiles;
    return DET_OK;
}

}  // namespace det

using namespace det;

extern "C" {

int64_t det_match_grid_workspace_bytes(int n, int64_t sum_g) {
    (void)n;
    return ((sum_g > 0 ? sum_g : 1) * 4 + 255) / 256 * 256;  // row maxima, fp32
}

int det_match_grid(const float* gt_boxes, const int32_t* gt_offsets, int n, int64_t sum_g, const float* anchors, int64_t r,
                   const det_anchor_level_t* levels_host, int num_levels, int a, const float* thresholds_host,
                   const int32_t* labels_host, int num_thresholds, int allow_low_quality, int64_t* matched_idx,
                   int8_t* labels, float* matched_iou, int32_t* stats, int32_t* pos_list, int list_cap, void* workspace,
                   int64_t workspace_bytes, void* stream) {
    DET_CHECK_ARG(n >= 0 && r >= 0 && sum_g >= 0, "negative size");
    if (n == 0 || r == 0) return DET_OK;
    DET_CHECK_ARG(gt_offsets && anchors && matched_idx && labels, "null pointer");
    DET_CHECK_ARG(sum_g == 0 || gt_boxes, "null gt_boxes");
    DET_CHECK_ARG(n <= 65535, "n > 65535");
    DET_CHECK_ARG(!pos_list || (stats && list_cap >= 1 && r < (1 << 24)), "pos_list needs stats, list_cap >= 1, r < 2^24");
    if (!aligned16(anchors) || (gt_boxes && !aligned16(gt_boxes))) {
        set_error("gt_boxes/anchors must be 16-byte aligned");
        return DET_ERR_ALIGN;
    }
    MatchRule rule;
    int rc = fill_rule(rule, thresholds_host, labels_host, num_thresholds, allow_low_quality);
    if (rc != DET_OK) return rc;
    GridLayoutDev lay;
    rc = fill_layout(lay, levels_host, num_levels, a, r);
    if (rc != DET_OK) return rc;
    if (a != 1 && a != 3 && a != 9) {
        set_error("det_match_grid: anchors per position must be 1, 3 or 9 (use det_match_anchors otherwise)");
        return DET_ERR_UNSUPPORTED;
    }
    if (allow_low_quality && sum_g > 0 && (!workspace || workspace_bytes < (int64_t)sizeof(float) * sum_g)) {
        set_error("workspace too small: need %lld bytes", (long long)det_match_grid_workspace_bytes(n, sum_g));
        return DET_ERR_WORKSPACE;
    }
    cudaStream_t st = as_stream(stream);
    float* rowmax = static_cast<float*>(workspace);
    auto g4 = reinterpret_cast<const float4*>(gt_boxes);
    auto a4 = reinterpret_cast<const float4*>(anchors);
    const int stats_words = stats ? n * 4 : 0;
    const int64_t k1_threads = (allow_low_quality ? sum_g * 32 : 0) > stats_words ? sum_g * 32 : stats_words;
    if (k1_threads > 0) {
        match_rowmax_kernel<<<(unsigned)((k1_threads + 255) / 256), 256, 0, st>>>(
            g4, allow_low_quality ? sum_g : 0, a4, lay, rowmax, stats, stats_words);
        DET_LAUNCH_OK("match_rowmax_kernel");
    }
    // images per CTA: amortise the anchor loads and the tile box while keeping >= 8 CTAs per SM in flight
    const int bx = (lay.tiles_total + kGridWarps - 1) / kGridWarps;
    int imgs = (int)(((int64_t)bx * n) / ((int64_t)sm_count() * 8));
    imgs = imgs < 1 ? 1 : (imgs > 8 ? 8 : imgs);
    dim3 grid((unsigned)bx, (unsigned)((n + imgs - 1) / imgs));
#define DET_LAUNCH_GRID(AA)                                                                                              \
    match_grid_kernel<AA, true><<<grid, kGridWarps * 32, 0, st>>>(g4, gt_offsets, a4, n, imgs, r, lay, rule, rowmax,      \
                                                                  matched_idx, labels, matched_iou, stats, pos_list,     \
                                                                  list_cap)
    if (a == 1) DET_LAUNCH_GRID(1);
    else if (a == 3) DET_LAUNCH_GRID(3);
    else DET_LAUNCH_GRID(9);
#undef DET_LAUNCH_GRID
    DET_LAUNCH_OK("match_grid_kernel");
    return DET_OK;
}

int det_rpn_loss_sampled(const det_head_level_t* levels_host, int num_levels, int a, const float* logits_flat,
                         const float* deltas_flat, float* grad_logits_flat, float* grad_deltas_flat,
                         const int32_t* samples, const int32_t* sample_count, int sample_cap,
                         const int32_t* clear_samples, const int32_t* clear_count, const int64_t* matched_idx,
                         const float* gt_boxes, const int32_t* gt_offsets, const float* anchors, int n, int64_t r, float wx,
                         float wy, float ww, float wh, float scale_clamp, int loss_type, float smooth_l1_beta,
                         float scale_cls, float scale_loc, const float* upstream, float* accumulators, float* sums_out,
                         void* stream) {
    DET_CHECK_ARG(n >= 0 && r >= 0 && sample_cap >= 1, "bad size");
    DET_CHECK_ARG(loss_type == 0 || loss_type == 1, "loss_type must be 0 (smooth-L1) or 1 (GIoU)");
    DET_CHECK_ARG(sums_out && accumulators, "null output");
    if (n == 0 || r == 0) {
        cudaError_t e = cudaMemsetAsync(sums_out, 0, 8 * sizeof(float), as_stream(stream));
        if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync");
        return DET_OK;
    }
    DET_CHECK_ARG(samples && sample_count && matched_idx && gt_offsets && anchors, "null pointer");
    DET_CHECK_ARG(r < (1 << 24), "r must be below 2^24 (sample entries pack the anchor row in 24 bits)");
    DET_CHECK_ARG((clear_samples == nullptr) == (clear_count == nullptr), "clear_samples and clear_count go together");
    if (!aligned16(anchors) || (gt_boxes && !aligned16(gt_boxes))) {
        set_error("gt_boxes/anchors must be 16-byte aligned");
        return DET_ERR_ALIGN;
    }
    HeadLayoutDev hl;
    int write_grads = 0;
    if (num_levels == 0) {
        DET_CHECK_ARG(logits_flat && deltas_flat, "flat layout: null logits/deltas");
        DET_CHECK_ARG((grad_logits_flat == nullptr) == (grad_deltas_flat == nullptr), "both gradients or none");
        if (!aligned16(deltas_flat) || (grad_deltas_flat && !aligned16(grad_deltas_flat))) {
            set_error("deltas / grad_deltas must be 16-byte aligned");
            return DET_ERR_ALIGN;
        }
        hl.nlev = 0;
        hl.a = 1;
        hl.lv[0] = HeadLevelDev{logits_flat, deltas_flat, grad_logits_flat, grad_deltas_flat, 1, 1, 0};
        write_grads = grad_logits_flat != nullptr;
    } else {
        DET_CHECK_ARG(levels_host && num_levels >= 1 && num_levels <= kGridMaxLevels && a >= 1, "bad head levels");
        hl.nlev = num_levels;
        hl.a = a;
        int64_t row = 0;
        int with = 0;
        for (int l = 0; l < num_levels; ++l) {
            const det_head_level_t& s = levels_host[l];
            DET_CHECK_ARG(s.objectness && s.deltas && s.h >= 1 && s.w >= 1, "head level: null pointer or empty plane");
            DET_CHECK_ARG((s.grad_objectness == nullptr) == (s.grad_deltas == nullptr), "both gradients or none");
            with += s.grad_objectness != nullptr;
            hl.lv[l] = HeadLevelDev{s.objectness, s.deltas, s.grad_objectness, s.grad_deltas, s.h, s.w, row};
            row += (int64_t)s.h * s.w * a;
        }
        DET_CHECK_ARG(row == r, "head levels do not cover r anchors");
        DET_CHECK_ARG(with == 0 || with == num_levels, "gradients for all levels or none");
        write_grads = with != 0;
    }
    const CodecW wt{wx, wy, ww, wh};
    auto g4 = reinterpret_cast<const float4*>(gt_boxes);
    auto a4 = reinterpret_cast<const float4*>(anchors);
    int32_t* ticket = reinterpret_cast<int32_t*>(accumulators + 8);
    if (loss_type == 1)
        rpn_loss_sampled_kernel<true><<<n, kSampledThreads, 0, as_stream(stream)>>>(
            hl, samples, sample_count, sample_cap, clear_samples, clear_count, matched_idx, g4, gt_offsets, a4, r, wt,
            scale_clamp, smooth_l1_beta, scale_cls, scale_loc, upstream, accumulators, ticket, sums_out, write_grads);
    else
        rpn_loss_sampled_kernel<false><<<n, kSampledThreads, 0, as_stream(stream)>>>(
            hl, samples, sample_count, sample_cap, clear_samples, clear_count, matched_idx, g4, gt_offsets, a4, r, wt,
            scale_clamp, smooth_l1_beta, scale_cls, scale_loc, upstream, accumulators, ticket, sums_out, write_grads);
    DET_LAUNCH_OK("rpn_loss_sampled_kernel");
    return DET_OK;
}

}  // extern "C"
