// Training side (subsystem 4): IoU target assignment, on-device fg/bg subsampling, and the fused loss
// forward + backward kernels.
//
// Reference behaviour reproduced (paths relative to the reference root):
//   python/src/models/rpn.py:161-168        per image: pairwise_iou(gt, anchors) -> Matcher
//   python/src/models/components/matcher.py:53-120   column max/argmax, threshold buckets, low-quality promotion
//   python/src/utils.py:34-76 + rpn.py:108-130       subsample_labels / _subsample_labels (counts; RNG differs)
//   python/src/models/rpn.py:187-244 + components/box_regression.py:128-168   losses (BCE-with-logits + L1/smooth-L1)
// The (G,R) IoU matrix, the (N,R,4) matched-gt tensor and the (N,R,4) target-delta tensor of the reference are
// never materialised: IoUs are recomputed from the box tables, targets are encoded on the fly for positives only.
#include "common.cuh"

namespace det {

constexpr int kMaxThresholds = 8;
struct MatchRule {
    float thr[kMaxThresholds];
    int8_t lab[kMaxThresholds + 1];
    int nthr;
    int allow_lq;
};

__device__ __forceinline__ int8_t bucket_label(const MatchRule& r, float v) {
    int k = 0;
#pragma unroll
    for (int i = 0; i < kMaxThresholds; ++i)
        if (i < r.nthr && v >= r.thr[i]) k = i + 1;  // thresholds ascending: bucket [thr[k-1], thr[k])
    return r.lab[k];
}

// IoUs are >= 0, so their bit patterns order like unsigned integers
__device__ __forceinline__ void atomic_max_nonneg(float* addr, float v) {
    atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

constexpr int kMatchThreads = 256;
constexpr int kMatchPerThread = 4;
constexpr int kGtChunk = 512;

// pass 1: per anchor column max / argmax over the image's gt boxes, threshold label; per gt row max (atomics)
__global__ void __launch_bounds__(kMatchThreads)
match_pass1_kernel(const float4* __restrict__ gt, const int32_t* __restrict__ gt_off, const float4* __restrict__ anchors,
                   int64_t r, MatchRule rule, int64_t* __restrict__ matched, int8_t* __restrict__ labels,
                   float* __restrict__ matched_iou, float* __restrict__ rowmax) {
    __shared__ float4 s_gt[kGtChunk];
    __shared__ float s_area[kGtChunk];
    __shared__ float s_rmax[kGtChunk];
    const int img = blockIdx.y;
    const int g0 = gt_off[img], g1 = gt_off[img + 1];
    const int G = g1 - g0;
    const int64_t base = (int64_t)blockIdx.x * (kMatchThreads * kMatchPerThread);
    float4 ab[kMatchPerThread];
    float aa[kMatchPerThread], best[kMatchPerThread];
    int bidx[kMatchPerThread];
#pragma unroll
    for (int k = 0; k < kMatchPerThread; ++k) {
        const int64_t j = base + k * kMatchThreads + threadIdx.x;
        ab[k] = (j < r) ? anchors[j] : make_float4(0.f, 0.f, 0.f, 0.f);
        aa[k] = box_area(ab[k]);
        best[k] = -1.0f;  // any IoU (>= 0) beats it, so the first gt wins ties like torch.max(dim=0)
        bidx[k] = 0;
    }
    for (int c0 = 0; c0 < G; c0 += kGtChunk) {
        const int cn = min(kGtChunk, G - c0);
        __syncthreads();
        for (int t = threadIdx.x; t < cn; t += kMatchThreads) {
            const float4 b = gt[g0 + c0 + t];
            s_gt[t] = b;
            s_area[t] = box_area(b);
            s_rmax[t] = 0.0f;
        }
        __syncthreads();
        for (int t = 0; t < cn; ++t) {
            const float4 gb = s_gt[t];
            const float ga = s_area[t];
            float rm = 0.0f;
#pragma unroll
            for (int k = 0; k < kMatchPerThread; ++k) {
                const int64_t j = base + k * kMatchThreads + threadIdx.x;
                // pairwise_iou(gt, anchors): boxes1 = gt, boxes2 = anchors (rpn.py:167)
                const float v = (j < r) ? pair_iou(gb, ga, ab[k], aa[k]) : 0.0f;
                if (v > best[k]) {
                    best[k] = v;
                    bidx[k] = c0 + t;
                }
                rm = fmaxf(rm, v);
            }
            if (rule.allow_lq && rm > s_rmax[t]) atomic_max_nonneg(&s_rmax[t], rm);
        }
        __syncthreads();
        if (rule.allow_lq)
            for (int t = threadIdx.x; t < cn; t += kMatchThreads)
                if (s_rmax[t] > 0.0f) atomic_max_nonneg(&rowmax[g0 + c0 + t], s_rmax[t]);
    }
#pragma unroll
    for (int k = 0; k < kMatchPerThread; ++k) {
        const int64_t j = base + k * kMatchThreads + threadIdx.x;
        if (j >= r) continue;
        const int64_t o = (int64_t)img * r + j;
        if (G == 0) {  // matcher.py:67-77: no gt -> match 0, label labels[0]
            matched[o] = 0;
            labels[o] = rule.lab[0];
            if (matched_iou) matched_iou[o] = 0.0f;
        } else {
            matched[o] = bidx[k];
            labels[o] = bucket_label(rule, best[k]);
            if (matched_iou) matched_iou[o] = best[k];
        }
    }
}

// pass 2 (low-quality promotion, matcher.py:96-120): label 1 wherever IoU(gt, anchor) == max over anchors for that gt
__global__ void __launch_bounds__(kMatchThreads)
match_pass2_kernel(const float4* __restrict__ gt, const int32_t* __restrict__ gt_off, const float4* __restrict__ anchors,
                   int64_t r, const float* __restrict__ rowmax, int8_t* __restrict__ labels) {
    __shared__ float4 s_gt[kGtChunk];
    __shared__ float s_area[kGtChunk];
    __shared__ float s_rmax[kGtChunk];
    const int img = blockIdx.y;
    const int g0 = gt_off[img], g1 = gt_off[img + 1];
    const int G = g1 - g0;
    if (G == 0) return;
    const int64_t base = (int64_t)blockIdx.x * (kMatchThreads * kMatchPerThread);
    float4 ab[kMatchPerThread];
    float aa[kMatchPerThread];
    bool hit[kMatchPerThread];
#pragma unroll
    for (int k = 0; k < kMatchPerThread; ++k) {
        const int64_t j = base + k * kMatchThreads + threadIdx.x;
        ab[k] = (j < r) ? anchors[j] : make_float4(0.f, 0.f, 0.f, 0.f);
        aa[k] = box_area(ab[k]);
        hit[k] = false;
    }
    for (int c0 = 0; c0 < G; c0 += kGtChunk) {
        const int cn = min(kGtChunk, G - c0);
        __syncthreads();
        for (int t = threadIdx.x; t < cn; t += kMatchThreads) {
            const float4 b = gt[g0 + c0 + t];
            s_gt[t] = b;
            s_area[t] = box_area(b);
            s_rmax[t] = rowmax[g0 + c0 + t];
        }
        __syncthreads();
        for (int t = 0; t < cn; ++t) {
            const float4 gb = s_gt[t];
            const float ga = s_area[t], rm = s_rmax[t];
#pragma unroll
            for (int k = 0; k < kMatchPerThread; ++k) hit[k] |= (pair_iou(gb, ga, ab[k], aa[k]) == rm);
        }
    }
#pragma unroll
    for (int k = 0; k < kMatchPerThread; ++k) {
        const int64_t j = base + k * kMatchThreads + threadIdx.x;
        if (j < r && hit[k]) labels[(int64_t)img * r + j] = 1;
    }
}

// ---- Matcher on a materialised (g, r) quality matrix ------------------------------------------------------------
__global__ void __launch_bounds__(256)
quality_pass1_kernel(const float* __restrict__ q, int64_t g, int64_t r, MatchRule rule, int64_t* __restrict__ matched,
                     int8_t* __restrict__ labels, float* __restrict__ rowmax, int32_t* __restrict__ negative_flag) {
    const int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int lane = threadIdx.x & 31;
    float best = 0.f;
    int bi = 0;
    bool bad = false;
    for (int64_t i = 0; i < g; ++i) {
        const float v = (j < r) ? q[i * r + j] : 0.0f;
        bad |= !(v >= 0.0f);
        // torch.max(dim=0): first maximum wins, NaN propagates (excluded by the >= 0 assertion)
        if (i == 0 || v > best) {
            best = v;
            bi = (int)i;
        }
        if (rule.allow_lq) {
            float m = (j < r) ? v : 0.0f;
            for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            if (lane == 0 && m > 0.0f) atomic_max_nonneg(&rowmax[i], m);
        }
    }
    if (bad && negative_flag) *negative_flag = 1;
    if (j < r) {
        matched[j] = bi;
        labels[j] = bucket_label(rule, best);
    }
}

__global__ void __launch_bounds__(256)
quality_pass2_kernel(const float* __restrict__ q, int64_t g, int64_t r, const float* __restrict__ rowmax,
                     int8_t* __restrict__ labels) {
    const int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (j >= r) return;
    bool hit = false;
    for (int64_t i = 0; i < g; ++i) hit |= (q[i * r + j] == rowmax[i]);
    if (hit) labels[j] = 1;
}

// ---- uniform random fg/bg subsample, one CTA per image -----------------------------------------------------------
__device__ __forceinline__ uint32_t sample_key(uint64_t seed, uint32_t img, uint32_t j) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ull * (((uint64_t)img << 32) | j);  // splitmix64 finaliser
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (uint32_t)(z >> 32);
}

constexpr int kSampleThreads = 1024;
constexpr int kTieCap = 256;

struct SampleSmem {
    int hist[2][256];
    int cnt[2];
    unsigned prefix[2];  // selected high bits of the k-th smallest key so far
    int remaining[2];    // rank of the threshold key inside the current bucket
    int ties[2][kTieCap];
    int ntie[2];
};

// cls: 0 = positive (label == 1 here: anything that is neither -1 nor bg), 1 = background (label == 0)
__global__ void __launch_bounds__(kSampleThreads)
subsample_kernel(int8_t* __restrict__ labels, int64_t r, int num_samples, float positive_fraction, uint64_t seed) {
    __shared__ SampleSmem sm;
    const int img = blockIdx.x, tid = threadIdx.x;
    int8_t* lab = labels + (int64_t)img * r;
    if (tid < 2) {
        sm.cnt[tid] = 0;
        sm.ntie[tid] = 0;
    }
    __syncthreads();
    int c0 = 0, c1 = 0;
    for (int64_t j = tid; j < r; j += kSampleThreads) {
        const int8_t l = lab[j];
        c0 += (l != -1 && l != 0);
        c1 += (l == 0);
    }
    c0 = warp_sum(c0);
    c1 = warp_sum(c1);
    if ((tid & 31) == 0) {
        if (c0) atomicAdd(&sm.cnt[0], c0);
        if (c1) atomicAdd(&sm.cnt[1], c1);
    }
    __syncthreads();
    const int npos = sm.cnt[0], nneg = sm.cnt[1];
    // utils.py:64-69: num_pos = min(#pos, int(S*f)); num_neg = min(#neg, S - num_pos)
    const int want_pos = min(npos, (int)((float)num_samples * positive_fraction));
    const int want_neg = min(nneg, num_samples - want_pos);
    const int want[2] = {want_pos, want_neg};
    const int have[2] = {npos, nneg};
    const bool need[2] = {want_pos < npos, want_neg < nneg};  // otherwise the whole class is kept
    if (!need[0] && !need[1]) return;
    // radix select (4 x 8 bits) of the want[c]-th smallest key of each class that needs thinning
    if (tid < 2) {
        sm.prefix[tid] = 0;
        sm.remaining[tid] = want[tid];  // we look for the key with rank want[c] (1-based) when want[c] > 0
    }
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        __syncthreads();
        for (int t = tid; t < 512; t += kSampleThreads) (&sm.hist[0][0])[t] = 0;
        __syncthreads();
        const unsigned pre0 = sm.prefix[0], pre1 = sm.prefix[1];
        const unsigned himask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
        for (int64_t j = tid; j < r; j += kSampleThreads) {
            const int8_t l = lab[j];
            if (l == -1) continue;
            const int c = (l == 0) ? 1 : 0;
            if (!need[c] || want[c] == 0) continue;
            const uint32_t key = sample_key(seed, (uint32_t)img, (uint32_t)j);
            if ((key & himask) == ((c ? pre1 : pre0) & himask)) atomicAdd(&sm.hist[c][(key >> shift) & 255], 1);
        }
        __syncthreads();
        if (tid < 2 && need[tid] && want[tid] > 0) {
            int rem = sm.remaining[tid], b = 0;
            while (b < 255 && rem > sm.hist[tid][b]) {
                rem -= sm.hist[tid][b];
                ++b;
            }
            sm.remaining[tid] = rem;  // rank inside bucket b
            sm.prefix[tid] |= (unsigned)b << shift;
        }
    }
    __syncthreads();
    // keys < T are kept; among keys == T the `remaining` lowest indices are kept
    const unsigned T[2] = {sm.prefix[0], sm.prefix[1]};
    for (int64_t j = tid; j < r; j += kSampleThreads) {
        const int8_t l = lab[j];
        if (l == -1) continue;
        const int c = (l == 0) ? 1 : 0;
        if (!need[c] || want[c] == 0) continue;
        if (sample_key(seed, (uint32_t)img, (uint32_t)j) == T[c]) {
            const int slot = atomicAdd(&sm.ntie[c], 1);
            if (slot < kTieCap) sm.ties[c][slot] = (int)j;
        }
    }
    __syncthreads();
    for (int64_t j = tid; j < r; j += kSampleThreads) {
        const int8_t l = lab[j];
        if (l == -1) continue;
        const int c = (l == 0) ? 1 : 0;
        if (!need[c]) continue;
        bool keep = false;
        if (want[c] > 0) {
            const uint32_t key = sample_key(seed, (uint32_t)img, (uint32_t)j);
            if (key < T[c]) keep = true;
            else if (key == T[c]) {
                int rank = 0;
                const int nt = min(sm.ntie[c], kTieCap);
                for (int t = 0; t < nt; ++t) rank += (sm.ties[c][t] < (int)j);
                keep = rank < sm.remaining[c];
            }
        }
        if (!keep) lab[j] = -1;
    }
    (void)have;
}

// ---- fused RPN loss forward + backward ---------------------------------------------------------------------------
struct CodecW {
    float wx, wy, ww, wh;
};

__device__ __forceinline__ float4 encode_target(const float4 s, const float4 t, const CodecW wt) {
    // Box2BoxTransform.get_deltas, box_regression.py:53-69
    const float sw = s.z - s.x, sh = s.w - s.y;
    const float scx = s.x + 0.5f * sw, scy = s.y + 0.5f * sh;
    const float tw = t.z - t.x, th = t.w - t.y;
    const float tcx = t.x + 0.5f * tw, tcy = t.y + 0.5f * th;
    return make_float4(wt.wx * (tcx - scx) / sw, wt.wy * (tcy - scy) / sh, wt.ww * logf(tw / sw),
                       wt.wh * logf(th / sh));
}

// smooth-L1 value and derivative wrt the prediction (fvcore semantics: beta < 1e-5 -> pure L1)
__device__ __forceinline__ void smooth_l1(float pred, float tgt, float beta, float& val, float& grad) {
    const float d = pred - tgt, n = fabsf(d);
    const float sgn = (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f);
    if (beta < 1e-5f) {
        val = n;
        grad = sgn;
    } else if (n < beta) {
        val = 0.5f * n * n / beta;
        grad = d / beta;
    } else {
        val = n - 0.5f * beta;
        grad = sgn;
    }
}

constexpr int kLossThreads = 256;

__global__ void __launch_bounds__(kLossThreads)
rpn_loss_kernel(const float* __restrict__ logits, const float4* __restrict__ deltas, const int8_t* __restrict__ labels,
                const int64_t* __restrict__ matched, const float4* __restrict__ gt, const int32_t* __restrict__ gt_off,
                const float4* __restrict__ anchors, int64_t total, int64_t r, CodecW wt, float beta, float gs_cls,
                float gs_loc, const float* __restrict__ upstream, float* __restrict__ sums,
                float* __restrict__ grad_logits, float4* __restrict__ grad_deltas) {
    __shared__ float s_part[4][kLossThreads / 32];
    if (upstream) {
        gs_cls *= upstream[0];
        gs_loc *= upstream[1];
    }
    float acc_cls = 0.f, acc_loc = 0.f;
    int npos = 0, nneg = 0;
    const int64_t nquads = (total + 3) >> 2;
    for (int64_t qd = (int64_t)blockIdx.x * kLossThreads + threadIdx.x; qd < nquads;
         qd += (int64_t)gridDim.x * kLossThreads) {
        const int64_t e0 = qd << 2;
        const bool full = e0 + 4 <= total;
        int8_t lab[4];
        if (full) {
            const uint32_t pk = __ldcs(reinterpret_cast<const unsigned int*>(labels + e0));
            lab[0] = (int8_t)(pk & 255); lab[1] = (int8_t)((pk >> 8) & 255);
            lab[2] = (int8_t)((pk >> 16) & 255); lab[3] = (int8_t)(pk >> 24);
        } else {
            for (int k = 0; k < 4; ++k) lab[k] = (e0 + k < total) ? labels[e0 + k] : (int8_t)-1;
        }
        const bool any_valid = (lab[0] >= 0) | (lab[1] >= 0) | (lab[2] >= 0) | (lab[3] >= 0);
        float gl[4] = {0.f, 0.f, 0.f, 0.f};
        if (any_valid) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (lab[k] < 0) continue;
                const int64_t e = e0 + k;
                const float x = logits[e];
                const float y = (float)lab[k];
                // BCE with logits: (1-y)*x - log_sigmoid(x), log_sigmoid(x) = min(x,0) - log1p(exp(-|x|))
                const float ls = fminf(x, 0.f) - log1pf(expf(-fabsf(x)));
                acc_cls += (1.f - y) * x - ls;
                gl[k] = (1.f / (1.f + expf(-x)) - y) * gs_cls;
                npos += lab[k] == 1;
                nneg += lab[k] == 0;
            }
        }
        if (grad_logits) {
            if (full) {
                st_stream(reinterpret_cast<float4*>(grad_logits + e0), make_float4(gl[0], gl[1], gl[2], gl[3]));
            } else {
                for (int k = 0; k < 4; ++k)
                    if (e0 + k < total) grad_logits[e0 + k] = gl[k];
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int64_t e = e0 + k;
            if (e >= total) break;
            float4 gd = make_float4(0.f, 0.f, 0.f, 0.f);
            if (lab[k] == 1) {
                const int64_t img = e / r, j = e - img * r;
                const float4 g = gt[gt_off[img] + matched[e]];
                const float4 tgt = encode_target(anchors[j], g, wt);
                const float4 p = deltas[e];
                float v, d;
                smooth_l1(p.x, tgt.x, beta, v, d); acc_loc += v; gd.x = d * gs_loc;
                smooth_l1(p.y, tgt.y, beta, v, d); acc_loc += v; gd.y = d * gs_loc;
                smooth_l1(p.z, tgt.z, beta, v, d); acc_loc += v; gd.z = d * gs_loc;
                smooth_l1(p.w, tgt.w, beta, v, d); acc_loc += v; gd.w = d * gs_loc;
            }
            if (grad_deltas) st_stream(grad_deltas + e, gd);
        }
    }
    // warp-shuffle + shared-memory reduction, one atomic per CTA and quantity
    acc_cls = warp_sum(acc_cls);
    acc_loc = warp_sum(acc_loc);
    float fpos = warp_sum((float)npos), fneg = warp_sum((float)nneg);
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
        for (int w = 0; w < kLossThreads / 32; ++w) t += s_part[threadIdx.x][w];
        if (t != 0.f) atomicAdd(&sums[threadIdx.x], t);
    }
}

// ---- fused YOLO-grid loss forward + backward: one thread per (image, cell) -----------------------------------------
struct YoloLossParams {
    int n, s, b, c;
    float stride_x, stride_y, lambda_coord, lambda_noobj, grad_scale;
};

__global__ void __launch_bounds__(128)
yolo_loss_kernel(const float* __restrict__ head, const int8_t* __restrict__ labels, const int64_t* __restrict__ matched,
                 const float4* __restrict__ gt, const int64_t* __restrict__ gt_cls, const int32_t* __restrict__ gt_off,
                 const float2* __restrict__ priors, YoloLossParams prm, const float* __restrict__ upstream,
                 float* __restrict__ sums, float* __restrict__ grad_head) {
    __shared__ float s_part[5][4];
    const float up_loc = upstream ? upstream[0] : 1.f, up_obj = upstream ? upstream[1] : 1.f,
                up_cls = upstream ? upstream[2] : 1.f;
    const int S2 = prm.s * prm.s, B = prm.b, C = prm.c, ch = B * 5 + C;
    const int64_t cells = (int64_t)prm.n * S2;
    const int64_t ci = (int64_t)blockIdx.x * 128 + threadIdx.x;
    float a_loc = 0.f, a_obj = 0.f, a_cls = 0.f, a_pos = 0.f, a_neg = 0.f;
    if (ci < cells) {
        const int img = (int)(ci / S2), cell = (int)(ci - (int64_t)img * S2);
        const int row = cell / prm.s, col = cell - row * prm.s;
        const float* t = head + ci * ch;
        float* g = grad_head ? grad_head + ci * ch : nullptr;
        if (g)
            for (int k = 0; k < C; ++k) g[B * 5 + k] = 0.f;
        for (int bi = 0; bi < B; ++bi) {
            const int64_t p = (int64_t)img * S2 * B + (int64_t)cell * B + bi;
            const int8_t lab = labels[p];
            float gx = 0.f, gy = 0.f, gw = 0.f, gh = 0.f, gc = 0.f;
            const float tc = t[bi * 5 + 4];
            if (lab >= 0) {
                const float y = (float)lab;
                const float wgt = (lab == 1) ? 1.0f : prm.lambda_noobj;
                const float ls = fminf(tc, 0.f) - log1pf(expf(-fabsf(tc)));
                a_obj += wgt * ((1.f - y) * tc - ls);
                gc = wgt * (1.f / (1.f + expf(-tc)) - y) * prm.grad_scale * up_obj;
                a_pos += lab == 1;
                a_neg += lab == 0;
            }
            if (lab == 1) {
                const int64_t gi = gt_off[img] + matched[p];
                const float4 gb = gt[gi];
                const float2 pr = priors[bi];
                const float bw = gb.z - gb.x, bh = gb.w - gb.y;
                const float xs = (gb.x + 0.5f * bw) / prm.stride_x - (float)col;
                const float ys = (gb.y + 0.5f * bh) / prm.stride_y - (float)row;
                const float tws = logf(bw / pr.x), ths = logf(bh / pr.y);
                const float sx = 1.f / (1.f + expf(-t[bi * 5 + 0])), sy = 1.f / (1.f + expf(-t[bi * 5 + 1]));
                const float dx = sx - xs, dy = sy - ys, dw = t[bi * 5 + 2] - tws, dh = t[bi * 5 + 3] - ths;
                a_loc += dx * dx + dy * dy + dw * dw + dh * dh;
                const float k2 = 2.0f * prm.lambda_coord * prm.grad_scale * up_loc;
                gx = k2 * dx * sx * (1.f - sx);
                gy = k2 * dy * sy * (1.f - sy);
                gw = k2 * dw;
                gh = k2 * dh;
                const int64_t cls = gt_cls[gi];
                for (int k = 0; k < C; ++k) {
                    const float x = t[B * 5 + k];
                    const float y = (k == cls) ? 1.f : 0.f;
                    const float ls = fminf(x, 0.f) - log1pf(expf(-fabsf(x)));
                    a_cls += (1.f - y) * x - ls;
                    if (g) g[B * 5 + k] += (1.f / (1.f + expf(-x)) - y) * prm.grad_scale * up_cls;
                }
            }
            if (g) {
                g[bi * 5 + 0] = gx; g[bi * 5 + 1] = gy; g[bi * 5 + 2] = gw; g[bi * 5 + 3] = gh; g[bi * 5 + 4] = gc;
            }
        }
    }
    float v[5] = {a_loc, a_obj, a_cls, a_pos, a_neg};
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        v[k] = warp_sum(v[k]);
        if (lane == 0) s_part[k][wid] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < 5) {
        const float t = s_part[threadIdx.x][0] + s_part[threadIdx.x][1] + s_part[threadIdx.x][2] + s_part[threadIdx.x][3];
        if (t != 0.f) atomicAdd(&sums[threadIdx.x], t);
    }
}

static int fill_rule(MatchRule& rule, const float* thresholds_host, const int32_t* labels_host, int num_thresholds,
                     int allow_low_quality) {
    if (num_thresholds < 0 || num_thresholds > kMaxThresholds || !labels_host || (num_thresholds && !thresholds_host)) {
        set_error("bad matcher rule (at most %d thresholds)", kMaxThresholds);
        return DET_ERR_BAD_ARG;
    }
    rule.nthr = num_thresholds;
    rule.allow_lq = allow_low_quality ? 1 : 0;
    for (int i = 0; i < kMaxThresholds; ++i) rule.thr[i] = i < num_thresholds ? thresholds_host[i] : INFINITY;
    for (int i = 0; i <= kMaxThresholds; ++i) rule.lab[i] = i <= num_thresholds ? (int8_t)labels_host[i] : (int8_t)0;
    for (int i = 0; i + 1 < num_thresholds; ++i)
        if (!(thresholds_host[i] <= thresholds_host[i + 1])) {
            set_error("thresholds must be ascending");
            return DET_ERR_BAD_ARG;
        }
    return DET_OK;
}

}  // namespace det

using namespace det;

extern "C" {

int64_t det_match_workspace_bytes(int n, int64_t r, int64_t sum_g) {
    (void)n;
    (void)r;
    return ((sum_g > 0 ? sum_g : 1) * 4 + 255) / 256 * 256;
}

int det_match_anchors(const float* gt_boxes, const int32_t* gt_offsets, int n, int64_t sum_g, const float* anchors,
                      int64_t r, const float* thresholds_host, const int32_t* labels_host, int num_thresholds,
                      int allow_low_quality, int64_t* matched_idx, int8_t* labels, float* matched_iou,
                      void* workspace, int64_t workspace_bytes, void* stream) {
    DET_CHECK_ARG(n >= 0 && r >= 0 && sum_g >= 0, "negative size");
    if (n == 0 || r == 0) return DET_OK;
    DET_CHECK_ARG(gt_offsets && anchors && matched_idx && labels, "null pointer");
    DET_CHECK_ARG(sum_g == 0 || gt_boxes, "null gt_boxes");
    DET_CHECK_ARG(n <= 65535, "n > 65535");
    if (!aligned16(anchors) || (gt_boxes && !aligned16(gt_boxes))) {
        set_error("gt_boxes/anchors must be 16-byte aligned");
        return DET_ERR_ALIGN;
    }
    MatchRule rule;
    int rc = fill_rule(rule, thresholds_host, labels_host, num_thresholds, allow_low_quality);
    if (rc != DET_OK) return rc;
    if (allow_low_quality && sum_g > 0 && (!workspace || workspace_bytes < (int64_t)sizeof(float) * sum_g)) {
        set_error("workspace too small");
        return DET_ERR_WORKSPACE;
    }
    cudaStream_t st = as_stream(stream);
    float* rowmax = static_cast<float*>(workspace);
    if (allow_low_quality && sum_g > 0) {
        cudaError_t e = cudaMemsetAsync(rowmax, 0, sizeof(float) * (size_t)sum_g, st);
        if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync");
    }
    dim3 grid((unsigned)((r + kMatchThreads * kMatchPerThread - 1) / (kMatchThreads * kMatchPerThread)), (unsigned)n);
    auto g4 = reinterpret_cast<const float4*>(gt_boxes);
    auto a4 = reinterpret_cast<const float4*>(anchors);
    match_pass1_kernel<<<grid, kMatchThreads, 0, st>>>(g4, gt_offsets, a4, r, rule, matched_idx, labels, matched_iou, rowmax);
    DET_LAUNCH_OK("match_pass1_kernel");
    if (allow_low_quality && sum_g > 0) {
        match_pass2_kernel<<<grid, kMatchThreads, 0, st>>>(g4, gt_offsets, a4, r, rowmax, labels);
        DET_LAUNCH_OK("match_pass2_kernel");
    }
    return DET_OK;
}

int det_match_quality(const float* quality, int64_t g, int64_t r, const float* thresholds_host,
                      const int32_t* labels_host, int num_thresholds, int allow_low_quality, int64_t* matched_idx,
                      int8_t* labels, int32_t* negative_flag, void* workspace, int64_t workspace_bytes, void* stream) {
    DET_CHECK_ARG(g >= 0 && r >= 0, "negative size");
    if (r == 0) return DET_OK;
    DET_CHECK_ARG(matched_idx && labels && (quality || g == 0), "null pointer");
    MatchRule rule;
    int rc = fill_rule(rule, thresholds_host, labels_host, num_thresholds, allow_low_quality);
    if (rc != DET_OK) return rc;
    cudaStream_t st = as_stream(stream);
    if (g == 0) {  // matcher.py:67-77
        cudaError_t e = cudaMemsetAsync(matched_idx, 0, sizeof(int64_t) * (size_t)r, st);
        if (e == cudaSuccess) e = cudaMemsetAsync(labels, (int)(uint8_t)rule.lab[0], (size_t)r, st);
        if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync");
        return DET_OK;
    }
    float* rowmax = static_cast<float*>(workspace);
    if (allow_low_quality) {
        if (!workspace || workspace_bytes < (int64_t)sizeof(float) * g) {
            set_error("workspace too small: need %lld bytes", (long long)(sizeof(float) * g));
            return DET_ERR_WORKSPACE;
        }
        cudaError_t e = cudaMemsetAsync(rowmax, 0, sizeof(float) * (size_t)g, st);
        if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync");
    }
    const unsigned blocks = (unsigned)((r + 255) / 256);
    quality_pass1_kernel<<<blocks, 256, 0, st>>>(quality, g, r, rule, matched_idx, labels, rowmax, negative_flag);
    DET_LAUNCH_OK("quality_pass1_kernel");
    if (allow_low_quality) {
        quality_pass2_kernel<<<blocks, 256, 0, st>>>(quality, g, r, rowmax, labels);
        DET_LAUNCH_OK("quality_pass2_kernel");
    }
    return DET_OK;
}

int det_subsample_labels(int8_t* labels, int n, int64_t r, int num_samples, float positive_fraction, uint64_t seed,
                         void* stream) {
    DET_CHECK_ARG(n >= 0 && r >= 0 && num_samples >= 0, "negative size");
    DET_CHECK_ARG(r < (1ll << 31), "r too large");
    if (n == 0 || r == 0) return DET_OK;
    DET_CHECK_ARG(labels, "null pointer");
    subsample_kernel<<<n, kSampleThreads, 0, as_stream(stream)>>>(labels, r, num_samples, positive_fraction, seed);
    DET_LAUNCH_OK("subsample_kernel");
    return DET_OK;
}

int det_rpn_loss(const float* logits, const float* deltas, const int8_t* labels, const int64_t* matched_idx,
                 const float* gt_boxes, const int32_t* gt_offsets, const float* anchors, int n, int64_t r, float wx,
                 float wy, float ww, float wh, float scale_clamp, int loss_type, float smooth_l1_beta,
                 float grad_scale_cls, float grad_scale_loc, const float* upstream, float* sums, float* grad_logits,
                 float* grad_deltas, void* stream) {
    (void)scale_clamp;
    DET_CHECK_ARG(n >= 0 && r >= 0, "negative size");
    if (loss_type != 0) {
        set_error("loss_type %d (GIoU) is not implemented in this round; use smooth_l1", loss_type);
        return DET_ERR_UNSUPPORTED;
    }
    if (n == 0 || r == 0) return DET_OK;
    DET_CHECK_ARG(logits && deltas && labels && matched_idx && gt_offsets && anchors && sums, "null pointer");
    if (!aligned16(deltas) || !aligned16(anchors) || !aligned16(logits) || (gt_boxes && !aligned16(gt_boxes)) ||
        (grad_logits && !aligned16(grad_logits)) || (grad_deltas && !aligned16(grad_deltas)) ||
        (reinterpret_cast<uintptr_t>(labels) & 3u)) {
        set_error("tensors must be 16-byte aligned (labels 4-byte)");
        return DET_ERR_ALIGN;
    }
    const int64_t total = (int64_t)n * r;
    const int64_t nquads = (total + 3) / 4;
    int64_t blocks = (nquads + kLossThreads - 1) / kLossThreads;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    rpn_loss_kernel<<<(unsigned)blocks, kLossThreads, 0, as_stream(stream)>>>(
        logits, reinterpret_cast<const float4*>(deltas), labels, matched_idx, reinterpret_cast<const float4*>(gt_boxes),
        gt_offsets, reinterpret_cast<const float4*>(anchors), total, r, CodecW{wx, wy, ww, wh}, smooth_l1_beta,
        grad_scale_cls, grad_scale_loc, upstream, sums, grad_logits, reinterpret_cast<float4*>(grad_deltas));
    DET_LAUNCH_OK("rpn_loss_kernel");
    return DET_OK;
}

int det_yolo_loss(const float* head, const int8_t* labels, const int64_t* matched_idx, const float* gt_boxes,
                  const int64_t* gt_classes, const int32_t* gt_offsets, int n, int s, int b, int c, int img_h,
                  int img_w, const float* priors, float lambda_coord, float lambda_noobj, float grad_scale,
                  const float* upstream, float* sums, float* grad_head, void* stream) {
    DET_CHECK_ARG(n >= 0 && s >= 1 && b >= 1 && c >= 0, "bad size");
    if (n == 0) return DET_OK;
    DET_CHECK_ARG(head && labels && matched_idx && gt_offsets && priors && sums, "null pointer");
    if (gt_boxes && !aligned16(gt_boxes)) {
        set_error("gt_boxes must be 16-byte aligned");
        return DET_ERR_ALIGN;
    }
    YoloLossParams prm;
    prm.n = n; prm.s = s; prm.b = b; prm.c = c;
    prm.stride_x = (float)((double)img_w / (double)s);
    prm.stride_y = (float)((double)img_h / (double)s);
    prm.lambda_coord = lambda_coord; prm.lambda_noobj = lambda_noobj; prm.grad_scale = grad_scale;
    const int64_t cells = (int64_t)n * s * s;
    yolo_loss_kernel<<<(unsigned)((cells + 127) / 128), 128, 0, as_stream(stream)>>>(
        head, labels, matched_idx, reinterpret_cast<const float4*>(gt_boxes), gt_classes, gt_offsets,
        reinterpret_cast<const float2*>(priors), prm, upstream, sums, grad_head);
    DET_LAUNCH_OK("yolo_loss_kernel");
    return DET_OK;
}

}  // extern "C"
