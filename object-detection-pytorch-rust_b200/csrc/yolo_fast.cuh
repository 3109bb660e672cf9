// Fast path of the fused YOLO-grid decode + per-class NMS (det_yolo_decode_nms) for small grids:
// at most 128 predictors (s*s*b) and 32 classes, boxes clipped to the image -- BASELINE configs[0]/[1].
//
// One CTA per image, ONE WARP PER CLASS.  A class has at most 128 candidates (one per predictor), so every
// per-class step fits a warp and needs no block barrier:
//   B  scores conf x class for the warp's class, threshold, candidate statistics (count, coordinate range)
//   C  order-preserving ballot compaction of the class's candidates
//   D  rank-by-counting sort (score descending, predictor ascending) -> boxes (+ coordinate offset when the
//      image takes torchvision's offset-trick branch), areas and scores in sorted order
//   E  every unordered pair of the class tested exactly once, load-balanced over the 32 lanes; a hit sets one bit
//      of the earlier box's suppression row (shared-memory atomicOr) and flags the row as non-empty
//   F  greedy resolution on the bit rows alone, visiting only non-empty rows
// then the per-class kept lists (each already in output order) are merged pairwise with binary-search ranks
// (log2(C) levels, truncated to max_det at every level) -- no second sort -- and the detections are written.
// Images whose offset-trick branch cannot be swept per class (non-finite or <= -1 coordinates: only possible with
// NaN/Inf logits here) take a slow exact path: repeated arg-max + suppression over all candidates.
//
// Specification: oracle/ref_torch.py yolo_decode / yolo_select_nms (no reference
// implementation exists); the NMS semantics are the reference's batched_nms (python/src/utils.py:96-119).
#pragma once
#include "nms_core.cuh"

namespace det {

constexpr int kFastMaxP = 128;
constexpr int kFastMaxC = 32;
constexpr int kFastMinWarps = 8;
constexpr int kFastMaxLevels = 6;  // ceil(log2(32)) + 1 list tables
constexpr int kFastBins = 256;      // score histogram for the top-max_det tier cut

struct YoloParams {
    int n, s, b, c;
    float stride_x, stride_y, img_w, img_h, scale_clamp, score_thresh, thr_f;
    int clip, mode;
    int64_t max_det;
};

__device__ __forceinline__ float sigmoidf_ref(float x) { return 1.0f / (1.0f + expf(-x)); }

struct FastLayout {
    int hs_floats, pp, cls_stride, pcp;
    size_t bytes;
    __host__ __device__ FastLayout(int s2, int ch, int p, int c) {
        hs_floats = (s2 * ch + 3) & ~3;
        pp = (p + 3) & ~3;
        // per class: sbox float4[pp] | rows uint4[pp] | sarea float[pp] | skey u32[pp] | spred u8[pad16(pp)]
        cls_stride = pp * 16 + pp * 16 + pp * 4 + pp * 4 + ((pp + 15) & ~15);
        pcp = (p * c + 3) & ~3;
        size_t cls_bytes = (size_t)cls_stride * c;
        const size_t merge_bytes = (size_t)pcp * 17 + 16;  // two u64 key buffers + one state byte per (p, c)
        if (cls_bytes < merge_bytes) cls_bytes = merge_bytes;
        bytes = (size_t)hs_floats * 4 + (size_t)pp * 16 + ((cls_bytes + 15) & ~(size_t)15);
    }
};

template <int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB)
yolo_fast_kernel(const float* __restrict__ head, const float2* __restrict__ priors, YoloParams prm,
                 float4* __restrict__ dense_boxes, float* __restrict__ dense_conf, float* __restrict__ dense_scores,
                 int64_t* __restrict__ det_flat, float4* __restrict__ det_boxes, float* __restrict__ det_scores,
                 int32_t* __restrict__ det_count) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int S2 = prm.s * prm.s, B = prm.b, C = prm.c, ch = B * 5 + C, P = S2 * B, PC = P * C;
    const FastLayout lay(S2, ch, P, C);
    const int Pp = lay.pp;
    float* hs = reinterpret_cast<float*>(smem_raw);
    float4* pbox = reinterpret_cast<float4*>(hs + lay.hs_floats);
    unsigned char* cls_region = reinterpret_cast<unsigned char*>(pbox + Pp);
    __shared__ int s_m[kFastMaxC];
    __shared__ float s_mx[kFastMaxC], s_mn[kFastMaxC];
    __shared__ int s_fl[kFastMaxC];
    __shared__ unsigned s_nz[kFastMaxC][4];
    __shared__ int s_kc[kFastMaxC];
    __shared__ int s_start[kFastMaxLevels][kFastMaxC + 1];
    __shared__ unsigned long long s_best;
    __shared__ __align__(16) int s_hist[kFastBins];
    const unsigned FULL = 0xffffffffu;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, T = blockDim.x;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int img = blockIdx.x;
    const int K = (int)min(prm.max_det, (int64_t)PC);

    DET_MARK(0);
    // ---- A: stage this image's logits with 16-byte coalesced loads over the aligned interior of its span
    const int64_t g0 = (int64_t)img * S2 * ch, g1 = g0 + (int64_t)S2 * ch;
    const float4* head4 = reinterpret_cast<const float4*>(head);
    for (int64_t v = (g0 >> 2) + tid; v < ((g1 + 3) >> 2); v += T) {
        const int64_t base = v << 2;
        if (base >= g0 && base + 4 <= g1) {
            const float4 q = ld_stream(head4 + v);
            float* d = hs + (base - g0);
            d[0] = q.x; d[1] = q.y; d[2] = q.z; d[3] = q.w;
        } else {
            for (int e = 0; e < 4; ++e)
                if (base + e >= g0 && base + e < g1) hs[base + e - g0] = ld_stream(head + base + e);
        }
    }
    __syncthreads();
    // transcendentals in place (sigmoid for conf and the class logits) and, by a disjoint set of threads working from
    // the raw tx, ty, tw, th (which the in-place pass leaves alone), the predictor boxes -- one barrier for both
    for (int i = tid; i < S2 * ch; i += T) {
        const int cell = i / ch, k = i - cell * ch;
        if (k < B * 5) {
            const int bi = k / 5, comp = k - bi * 5;
            if (comp != 4) continue;
            const float y = sigmoidf_ref(hs[i]);
            hs[i] = y;
            if (dense_conf) dense_conf[(int64_t)img * P + cell * B + bi] = y;
        } else {
            hs[i] = sigmoidf_ref(hs[i]);
        }
    }
    for (int p = T - 1 - tid; p < P; p += T) {  // the last threads of the CTA have the fewest logits above
        const int cell = p / B, bi = p - cell * B;
        const int row = cell / prm.s, col = cell - row * prm.s;
        const float* t = hs + cell * ch + bi * 5;
        const float2 pr = priors[bi];
        const float tw = (t[2] > prm.scale_clamp) ? prm.scale_clamp : t[2];  // torch.clamp(max=): NaN stays NaN
        const float th = (t[3] > prm.scale_clamp) ? prm.scale_clamp : t[3];
        const float cx = (sigmoidf_ref(t[0]) + (float)col) * prm.stride_x;
        const float cy = (sigmoidf_ref(t[1]) + (float)row) * prm.stride_y;
        const float w = expf(tw) * pr.x, h = expf(th) * pr.y;
        float4 bx = make_float4(cx - 0.5f * w, cy - 0.5f * h, cx + 0.5f * w, cy + 0.5f * h);
        if (prm.clip) {  // torch clamp(min=0,max=W): NaN stays NaN
            bx.x = min_nan(max_nan(bx.x, 0.f), prm.img_w); bx.z = min_nan(max_nan(bx.z, 0.f), prm.img_w);
            bx.y = min_nan(max_nan(bx.y, 0.f), prm.img_h); bx.w = min_nan(max_nan(bx.w, 0.f), prm.img_h);
        }
        pbox[p] = bx;
        if (dense_boxes) dense_boxes[(int64_t)img * P + p] = bx;
    }
    for (int i = tid; i < kFastBins; i += T) s_hist[i] = 0;
    __syncthreads();
    DET_MARK(1);

    // score histogram geometry: candidates have bits(score) in (bits(thr), bits(1.0)]; scores are >= +0
    const unsigned lo_bits = prm.score_thresh > 0.0f ? __float_as_uint(prm.score_thresh) : 0u;
    const unsigned hi_bits = 0x3F800000u;
    const int hshift = hi_bits > lo_bits ? max(0, 32 - __clz(hi_bits - lo_bits) - 8) : 0;
    // ---- B: the warp's class: scores, threshold, statistics of the candidate boxes
    const int c = wid;
    const bool cls_warp = c < C;
    float sc[4];
    bool pass[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        sc[k] = 0.f;
        pass[k] = false;
    }
    if (cls_warp) {
        float mx = -INFINITY, mn = INFINITY;
        int fin = 1, nonan = 1, mcount = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int p = lane + 32 * k;
            if (p < P) {
                const int cell = p / B, bi = p - cell * B;
                sc[k] = hs[cell * ch + bi * 5 + 4] * hs[cell * ch + B * 5 + c];
                pass[k] = sc[k] > prm.score_thresh;
                if (dense_scores) dense_scores[(int64_t)img * PC + (int64_t)p * C + c] = sc[k];
                if (pass[k]) {
                    const float4 b = pbox[p];
                    mx = max_nan(mx, max_nan(max_nan(b.x, b.y), max_nan(b.z, b.w)));
                    mn = min_nan(mn, min_nan(min_nan(b.x, b.y), min_nan(b.z, b.w)));
                    fin &= (int)(isfinite(b.x) && isfinite(b.y) && isfinite(b.z) && isfinite(b.w));
                    nonan &= (int)((b.x == b.x) && (b.y == b.y) && (b.z == b.z) && (b.w == b.w));
                    ++mcount;
                    atomicAdd(&s_hist[min(kFastBins - 1, (int)((__float_as_uint(sc[k]) - lo_bits) >> hshift))], 1);
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mx = max_nan(mx, __shfl_xor_sync(FULL, mx, o));
            mn = min_nan(mn, __shfl_xor_sync(FULL, mn, o));
            fin &= __shfl_xor_sync(FULL, fin, o);
            nonan &= __shfl_xor_sync(FULL, nonan, o);
            mcount += __shfl_xor_sync(FULL, mcount, o);
        }
        if (lane == 0) {
            s_m[c] = mcount;
            s_mx[c] = mx;
            s_mn[c] = mn;
            s_fl[c] = fin | (nonan << 1);
        }
    }
    __syncthreads();
    int cnt = 0, gcat = 0, gfl = 3;
    float gmx = -INFINITY, gmn = INFINITY;
    for (int q = 0; q < C; ++q) {
        const int mq = s_m[q];
        cnt += mq;
        if (mq) gcat = q;
        gmx = max_nan(gmx, s_mx[q]);
        gmn = min_nan(gmn, s_mn[q]);
        gfl &= s_fl[q];
    }
    // reference CPU rule: boxes.numel() <= 4000 -> coordinate-offset trick (torchvision/ops/boxes.py batched_nms)
    const bool trick = (prm.mode == DET_NMS_AUTO) ? (cnt <= 1000) : (prm.mode == DET_NMS_OFFSET_TRICK);
    const float span = trick ? gmx + 1.0f : 0.0f;  // max_coordinate + torch.tensor(1).to(boxes)
    const float thr_f = prm.thr_f;
    // categories can be swept independently iff shifted boxes of different categories cannot intersect
    const float far = gmx + (float)gcat * span;
    const bool sweep_ok = (gfl & 1) && gmn > -1.0f && thr_f >= 0.0f && isfinite(far);
    const bool by_cat = !trick || sweep_ok;
    const bool nonan = (gfl & 2) != 0;
    // ---- top-max_det tier: only the first K kept detections in global score order are output, and a candidate can
    //      only be suppressed by higher-scored ones, so the NMS is first run on the candidates above a histogram cut
    //      that holds about 4K/3 of them; if that already yields K kept boxes the rest cannot matter.  Otherwise
    //      (heavy suppression) everything is redone on all candidates.  Every warp derives the cut on its own.
    unsigned tier_bits = 0u;
    {
        const int target = K + K / 3 + 8;
        if (cnt > target) {
            int h8[8];
            const int4 ha = reinterpret_cast<const int4*>(s_hist)[lane * 2], hb = reinterpret_cast<const int4*>(s_hist)[lane * 2 + 1];
            h8[0] = ha.x; h8[1] = ha.y; h8[2] = ha.z; h8[3] = ha.w; h8[4] = hb.x; h8[5] = hb.y; h8[6] = hb.z; h8[7] = hb.w;
            const int mine = h8[0] + h8[1] + h8[2] + h8[3] + h8[4] + h8[5] + h8[6] + h8[7];
            int suf = mine;  // inclusive suffix sum over lanes (high scores live in the high bins)
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_down_sync(FULL, suf, o);
                if (lane + o < 32) suf += v;
            }
            const unsigned reach = __ballot_sync(FULL, suf >= target);  // a prefix of lanes: lane 0 holds cnt > target
            const int L = 31 - __clz(reach);
            int bin = 0;
            if (lane == L) {
                int run = suf - mine;
                bin = 8 * L;
#pragma unroll
                for (int k = 7; k >= 0; --k) {
                    run += h8[k];
                    if (run >= target) {
                        bin = 8 * L + k;
                        break;
                    }
                }
            }
            bin = __shfl_sync(FULL, bin, L);
            tier_bits = bin > 0 ? lo_bits + ((unsigned)bin << hshift) : 0u;
        }
    }
    DET_MARK(2);

    uint64_t* buf_a = reinterpret_cast<uint64_t*>(cls_region);
    uint64_t* buf_b = buf_a + lay.pcp;
    const uint64_t* fin_keys = buf_a;
    int nout = 0;

    if (by_cat) {
        uint64_t mykey[4];
        int mypos[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            mykey[k] = 0ull;
            mypos[k] = -1;
        }
        for (int tier = tier_bits ? 0 : 1; tier < 2; ++tier) {
        const unsigned cut_bits = tier == 0 ? tier_bits : 0u;
        if (cls_warp) {
            unsigned char* my = cls_region + (size_t)c * lay.cls_stride;
            float4* sbox = reinterpret_cast<float4*>(my);
            uint4* rows4 = reinterpret_cast<uint4*>(my + Pp * 16);
            unsigned* rows = reinterpret_cast<unsigned*>(rows4);
            float* sarea = reinterpret_cast<float*>(my + Pp * 32);
            unsigned* skey = reinterpret_cast<unsigned*>(my + Pp * 36);
            unsigned char* spred = my + Pp * 40;
            unsigned* ckey = rows;  // unsorted candidates live in the row storage until the rows are needed
            unsigned char* cp = reinterpret_cast<unsigned char*>(rows + Pp);
            // ---- C: order-preserving compaction (predictor ascending)
            if (lane < 4) s_nz[c][lane] = 0u;
            int m = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const bool in = pass[k] && __float_as_uint(sc[k]) >= cut_bits;
                const unsigned bal = __ballot_sync(FULL, in);
                if (in) {
                    const int pos = m + __popc(bal & lt_mask);
                    ckey[pos] = __float_as_uint(sc[k]);  // scores are >= +0: the bit patterns order like the values
                    cp[pos] = (unsigned char)(lane + 32 * k);
                }
                m += __popc(bal);
            }
            __syncwarp();
            DET_MARK(5);
            // ---- D: rank by counting -> sorted order (score descending, predictor ascending)
            const float off = trick ? (float)c * span : 0.0f;  // idxs.to(boxes) * (max_coordinate + 1)
            for (int i0 = 0; i0 < m; i0 += 32) {
                const int i = i0 + lane;
                const bool has = i < m;
                const unsigned si = has ? ckey[i] : 0u;
                int rank = 0;
#pragma unroll 4
                for (int j = 0; j < m; ++j) {
                    const unsigned sj = ckey[j];
                    rank += ((sj > si) || (sj == si && j < i)) ? 1 : 0;
                }
                if (has) {
                    const int p = (int)cp[i];
                    float4 b = pbox[p];
                    if (trick) {
                        b.x += off; b.y += off; b.z += off; b.w += off;
                    }
                    sbox[rank] = b;
                    sarea[rank] = box_area(b);
                    skey[rank] = si;
                    spred[rank] = (unsigned char)p;
                }
            }
            __syncwarp();
            for (int r = lane; r < m; r += 32) rows4[r] = make_uint4(0u, 0u, 0u, 0u);
            __syncwarp();
            DET_MARK(6);
            // ---- E: all pairs once: item q -> (row r, distance d), partner (r + d) mod m
            if (m >= 2) {
                const int half = m >> 1, items = m * half;
                const unsigned inv = half > 1 ? (0xFFFFFFFFu / (unsigned)half) + 1u : 0u;
                const bool even = (m & 1) == 0;
                for (int q = lane; q < items; q += 32) {
                    const int r = half > 1 ? (int)__umulhi((unsigned)q, inv) : q;
                    const int d = q - r * half + 1;
                    if (even && d == half && r >= half) continue;  // distance-m/2 pairs are met from the lower half
                    int j = r + d;
                    j = (j >= m) ? j - m : j;
                    const int lo = min(r, j), hi = max(r, j);
                    const float4 ba = sbox[lo], bb = sbox[hi];
                    const float aa = sarea[lo], ab = sarea[hi];
                    const bool hit = nonan ? nms_suppresses<true>(ba, aa, bb, ab, thr_f)
                                           : nms_suppresses<false>(ba, aa, bb, ab, thr_f);
                    if (hit) {
                        atomicOr(rows + lo * 4 + (hi >> 5), 1u << (hi & 31));
                        atomicOr(&s_nz[c][lo >> 5], 1u << (lo & 31));
                    }
                }
            }
            __syncwarp();
            DET_MARK(7);
            // ---- F: greedy resolution on the bit rows (uniform across the warp), only non-empty rows matter
            unsigned alive[4];
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                const int bits = min(32, max(0, m - 32 * w));
                alive[w] = bits == 32 ? 0xffffffffu : ((1u << bits) - 1u);
            }
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                unsigned nzw = s_nz[c][w];
                while (nzw) {
                    const int b = __ffs(nzw) - 1;
                    nzw &= nzw - 1;
                    if ((alive[w] >> b) & 1u) {
                        const uint4 rw = rows4[w * 32 + b];
                        alive[0] &= ~rw.x; alive[1] &= ~rw.y; alive[2] &= ~rw.z; alive[3] &= ~rw.w;
                    }
                }
            }
            int kc = __popc(alive[0]) + __popc(alive[1]) + __popc(alive[2]) + __popc(alive[3]);
            while (kc > K) {  // only the first K survivors of a class can reach the output
#pragma unroll
                for (int w = 3; w >= 0; --w)
                    if (alive[w]) {
                        alive[w] &= ~(0x80000000u >> __clz(alive[w]));
                        break;
                    }
                --kc;
            }
            // ---- G: the lane's kept keys and their ranks in the class's kept list
            int before = 0;
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                const int i = lane + 32 * w;
                mypos[w] = -1;
                if ((alive[w] >> lane) & 1u) {
                    mypos[w] = before + __popc(alive[w] & lt_mask);
                    mykey[w] = ((uint64_t)(~skey[i]) << 32) | (uint64_t)((unsigned)spred[i] * (unsigned)C + (unsigned)c);
                }
                before += __popc(alive[w]);
            }
            if (lane == 0) s_kc[c] = kc;
            DET_MARK(8);
        }
        __syncthreads();  // every class is done with its scratch: the merge buffers may now overwrite it
        if (tier == 0) {
            int tk = 0;
            for (int q = 0; q < C; ++q) tk += s_kc[q];
            if (tk >= K) break;   // the first K kept detections all lie above the cut
            __syncthreads();      // heavy suppression: redo on every candidate
        }
        }
        DET_MARK(3);
        // list tables of every merge level (lengths are known without looking at the keys)
        int levels = 0;
        if (wid == 0) {
            int len = lane < C ? s_kc[lane] : 0;
            int nl = C, lev = 0;
            while (true) {
                int incl = len;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(FULL, incl, o);
                    if (lane >= o) incl += v;
                }
                if (lane < nl) s_start[lev][lane] = incl - len;
                if (lane == nl - 1) s_start[lev][nl] = incl;
                if (nl == 1) break;
                const int la = __shfl_sync(FULL, len, (2 * lane) & 31), lb = __shfl_sync(FULL, len, (2 * lane + 1) & 31);
                const int nn = (nl + 1) >> 1;
                len = (lane < nn) ? min(la + ((2 * lane + 1 < nl) ? lb : 0), K) : 0;
                nl = nn;
                ++lev;
            }
        }
        {
            int nl = C;
            while (nl > 1) {
                nl = (nl + 1) >> 1;
                ++levels;
            }
        }
        // own class offset: exclusive prefix of s_kc
        int off_c = 0;
        for (int q = 0; q < c && q < C; ++q) off_c += s_kc[q];
#pragma unroll
        for (int w = 0; w < 4; ++w)
            if (mypos[w] >= 0) buf_a[off_c + mypos[w]] = mykey[w];
        __syncthreads();
        DET_MARK(9);
        // ---- pairwise merges: position = index in own list + lower_bound in the partner list; keys are distinct
        uint64_t* src = buf_a;
        uint64_t* dst = buf_b;
        int nl = C;
        for (int lev = 0; lev < levels; ++lev) {
            const int* st = s_start[lev];
            const int* nst = s_start[lev + 1];
            const int total = st[nl];
            for (int e = tid; e < total; e += T) {
                int lo = 0, hi = nl;
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (st[mid] <= e) lo = mid; else hi = mid;
                }
                const int l = lo, partner = l ^ 1;
                const uint64_t key = src[e];
                int pos = e - st[l];
                if (partner < nl) {
                    int a = st[partner], b = st[partner + 1];
                    const int a0 = a;
                    while (a < b) {
                        const int mid = (a + b) >> 1;
                        if (src[mid] < key) a = mid + 1; else b = mid;
                    }
                    pos += a - a0;
                }
                if (pos < K) dst[nst[l >> 1] + pos] = key;
            }
            __syncthreads();
            uint64_t* t = src;
            src = dst;
            dst = t;
            nl = (nl + 1) >> 1;
        }
        fin_keys = src;
        nout = min(s_start[levels][1], K);
    } else {
        // ---- slow exact path (offset trick with non-finite / <= -1 coordinates): global greedy by repeated arg-max
        unsigned char* st = reinterpret_cast<unsigned char*>(buf_b + lay.pcp);
        for (int f = tid; f < PC; f += T) {
            const int p = f / C, q = f - p * C, cell = p / B, bi = p - cell * B;
            const float s = hs[cell * ch + bi * 5 + 4] * hs[cell * ch + B * 5 + q];
            st[f] = (s > prm.score_thresh) ? 1 : 0;
        }
        __syncthreads();
        int k = 0;
        while (k < K) {
            if (tid == 0) s_best = ~0ull;
            __syncthreads();
            unsigned long long best = ~0ull;
            for (int f = tid; f < PC; f += T) {
                if (st[f] != 1) continue;
                const int p = f / C, q = f - p * C, cell = p / B, bi = p - cell * B;
                const float s = hs[cell * ch + bi * 5 + 4] * hs[cell * ch + B * 5 + q];
                const unsigned long long key = ((unsigned long long)(~__float_as_uint(s)) << 32) | (unsigned)f;
                best = key < best ? key : best;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const unsigned long long v = __shfl_xor_sync(FULL, best, o);
                best = v < best ? v : best;
            }
            if (lane == 0 && best != ~0ull) atomicMin(&s_best, best);
            __syncthreads();
            const unsigned long long bk = s_best;
            if (bk == ~0ull) break;
            const int fs = (int)(unsigned)bk;
            const int ps = fs / C, qs = fs - ps * C;
            float4 kb = pbox[ps];
            const float koff = (float)qs * span;
            kb.x += koff; kb.y += koff; kb.z += koff; kb.w += koff;
            const float ka = box_area(kb);
            for (int f = tid; f < PC; f += T) {
                if (st[f] != 1 || f == fs) continue;
                const int p = f / C, q = f - p * C;
                float4 b = pbox[p];
                const float o = (float)q * span;
                b.x += o; b.y += o; b.z += o; b.w += o;
                if (nms_suppresses<false>(kb, ka, b, box_area(b), thr_f)) st[f] = 0;
            }
            if (tid == 0) {
                buf_a[k] = bk;
                st[fs] = 2;
            }
            ++k;
            __syncthreads();
        }
        __syncthreads();
        fin_keys = buf_a;
        nout = k;
    }
    DET_MARK(4);
    // ---- detections, by descending score
    for (int j = tid; j < nout; j += T) {
        const uint64_t key = fin_keys[j];
        const unsigned flat = (unsigned)key;
        const int64_t o = (int64_t)img * prm.max_det + j;
        det_flat[o] = (int64_t)flat;
        if (det_boxes) det_boxes[o] = pbox[flat / (unsigned)C];
        if (det_scores) det_scores[o] = __uint_as_float(~(unsigned)(key >> 32));
    }
    if (tid == 0) det_count[img] = nout;
    DET_MARK(15);
}

}  // namespace det
