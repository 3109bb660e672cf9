// Dense anchor head decode for ALL pyramid levels in one persistent launch (subsystem 1, BASELINE configs[3]).
//
// head of level l: (N, A*(5+C), Hl, Wl) fp32, the conv's own NCHW layout.  Output: boxes (N,R,4), best score (N,R),
// best class (N,R) int64 with R = sum_l Hl*Wl*A in (level, h, w, a) order.  340 bytes are read per anchor and 28
// written, so the kernel is a pure HBM stream; it is organised around the copy engine rather than around loads:
//
//   * work unit = (level, image, tile of 256 consecutive positions); a CTA (one per SM, persistent) walks its units
//     and, inside a unit, the A anchors -- each (unit, anchor) "step" needs the [5+C planes x 128 positions] slab of
//     that anchor, which is contiguous runs of 1 KB, one per plane;
//   * a dedicated producer warp fetches slabs with 1-D bulk async copies (cp.async.bulk global -> shared, one per
//     plane, completion counted in bytes on a "full" mbarrier) as soon as the consumers release a stage through its
//     "empty" mbarrier: no registers and no issue slots of the compute warps, 87 KB per SM (13 MB per GPU) in flight;
//   * compute: one thread per position reads its column of the slab from shared memory (conflict-free), keeps two
//     interleaved (max, arg-max) chains over the class planes, decodes the box, and parks the result in a shared
//     output tile; after the last anchor of the unit the tile -- 256*A consecutive output rows -- is written with
//     fully coalesced 16/4/8-byte stores.
//
// Specification: oracle/ref_torch.py dense_decode (own; SURVEY.md section 8 row a15 -- no reference implementation).
#include "common.cuh"
#include <stdlib.h>

namespace det {

constexpr int kTile = 256;       // positions per slab == consumer threads per CTA
constexpr int kMaxStages = 4;
constexpr int kMaxLevels = 8;

struct DenseLevelDev {
    const float* head;
    const float2* anchors_wh;
    int w, hw, tiles, unit_begin;
    float stride;
    int64_t out_offset;
};

struct DenseArgs {
    DenseLevelDev lv[kMaxLevels];
    int num_levels, n, a, c, total_units, stages;
    float scale_clamp;
    int64_t out_img_stride;
    float4* boxes_out;
    float* score_out;
    int64_t* class_out;
};

// 1 / (1 + e^-x): __frcp_rn is the correctly rounded reciprocal, i.e. bit-identical to the IEEE quotient 1.0f / d
__device__ __forceinline__ float sigmoidf_dd(float x) { return __frcp_rn(1.0f + expf(-x)); }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            "  .reg .pred p;\n"
            "  mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "  selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
// 1-D bulk async copy global -> shared (SASS UBLKCP); bytes and both addresses are multiples of 16
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// torch.max(dim) order on (value, class): a NaN beats everything, the first maximum / first NaN wins
__device__ __forceinline__ bool argmax_takes_dd(float v, int iv, float best, int ib) {
    const bool vn = v != v, bn = best != best;
    if (vn || bn) return vn && (!bn || iv < ib);
    return v > best || (v == best && iv < ib);
}

struct StepInfo {
    int level, img, p0, valid, ai;
};

__device__ __forceinline__ StepInfo decode_step(const DenseArgs& g, int s) {
    StepInfo si;
    const int u = blockIdx.x + (s / g.a) * gridDim.x;
    si.ai = s - (s / g.a) * g.a;
    int l = 0;
#pragma unroll
    for (int q = 1; q < kMaxLevels; ++q)
        if (q < g.num_levels && u >= g.lv[q].unit_begin) l = q;
    const int local = u - g.lv[l].unit_begin;
    si.level = l;
    si.img = local / g.lv[l].tiles;
    si.p0 = (local - si.img * g.lv[l].tiles) * kTile;
    si.valid = min(kTile, g.lv[l].hw - si.p0);
    return si;
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kTile) : "memory"); }

// kTile consumer threads (one per position of a slab) + one producer warp that only feeds the copy engine
__global__ void __launch_bounds__(kTile + 32, 1) dense_decode_tma_kernel(const __grid_constant__ DenseArgs g) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int nch = 5 + g.c, A = g.a, NST = g.stages;
    float* slab = reinterpret_cast<float*>(smem_raw);                             // [NST][nch][kTile]
    float4* o_box = reinterpret_cast<float4*>(slab + (size_t)NST * nch * kTile);  // [kTile * A]
    float* o_score = reinterpret_cast<float*>(o_box + kTile * A);
    int* o_cls = reinterpret_cast<int*>(o_score + kTile * A);
    __shared__ __align__(8) unsigned long long s_full[kMaxStages], s_empty[kMaxStages];
    __shared__ StepInfo s_info[kMaxStages];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int my_units = ((int)blockIdx.x < g.total_units) ? (g.total_units - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int nsteps = my_units * A;
    if (tid == 0) {
        for (int i = 0; i < NST; ++i) {
            mbar_init(smem_u32(&s_full[i]), 1);            // the producer's arrive.expect_tx
            mbar_init(smem_u32(&s_empty[i]), kTile / 32);  // one arrive per consumer warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (wid == kTile / 32) {
        // ---- producer warp: one elected lane streams the slabs, running up to NST steps ahead of the consumers
        if (lane == 0) {
            for (int s = 0; s < nsteps; ++s) {
                const int st = s % NST;
                if (s >= NST) mbar_wait(smem_u32(&s_empty[st]), (uint32_t)(((s / NST) - 1) & 1));
                const StepInfo si = decode_step(g, s);
                s_info[st] = si;
                const DenseLevelDev& L = g.lv[si.level];
                const uint32_t bar = smem_u32(&s_full[st]);
                const uint32_t bytes = (uint32_t)si.valid * 4u;
                mbar_expect_tx(bar, bytes * (uint32_t)nch);
                const float* src = L.head + ((int64_t)(si.img * A + si.ai) * nch) * L.hw + si.p0;
                const uint32_t dst = smem_u32(slab + (size_t)st * nch * kTile);
                for (int k = 0; k < nch; ++k) bulk_g2s(dst + (uint32_t)k * kTile * 4u, src + (int64_t)k * L.hw, bytes, bar);
            }
        }
        return;
    }

    // ---- consumers
    for (int s = 0; s < nsteps; ++s) {
        const int st = s % NST;
        mbar_wait(smem_u32(&s_full[st]), (uint32_t)((s / NST) & 1));
        const StepInfo si = s_info[st];  // written by the producer before it armed the barrier
        const DenseLevelDev& L = g.lv[si.level];
        float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f, t4 = 0.f, ba = -INFINITY, bb = -INFINITY;
        int ia = 0, ib = 1;
        if (tid < si.valid) {
            const float* col = slab + (size_t)st * nch * kTile + tid;
            t0 = col[0]; t1 = col[kTile]; t2 = col[2 * kTile]; t3 = col[3 * kTile]; t4 = col[4 * kTile];
            const float* cc = col + 5 * kTile;
            // two interleaved (running max, first arg-max) chains, 8 shared-memory loads in flight.  The running max
            // is NaN-propagating (one FMNMX); `q > max` is false once the max is NaN, so a NaN column is finished by
            // the exact rescan below (torch.max: the first NaN wins).
            const int C = g.c;
            int k = 0;
            for (; k + 8 <= C; k += 8) {
                float q[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) q[i] = cc[(k + i) * kTile];
#pragma unroll
                for (int i = 0; i < 8; i += 2) {
                    ia = (q[i] > ba) ? k + i : ia;
                    ba = max_nan(ba, q[i]);
                    ib = (q[i + 1] > bb) ? k + i + 1 : ib;
                    bb = max_nan(bb, q[i + 1]);
                }
            }
            for (; k < C; ++k) {
                const float q = cc[k * kTile];
                if (k & 1) {
                    ib = (q > bb) ? k : ib;
                    bb = max_nan(bb, q);
                } else {
                    ia = (q > ba) ? k : ia;
                    ba = max_nan(ba, q);
                }
            }
            if (ba != ba || bb != bb) {  // rare: first NaN of the column
                for (int j = 0; j < C; ++j) {
                    const float q = cc[j * kTile];
                    if (q != q) {
                        ba = q;
                        ia = j;
                        break;
                    }
                }
                bb = -INFINITY;
                ib = C;
            }
        }
        // the slab is in registers: hand the stage back to the producer before the transcendental epilogue
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&s_empty[st]));
        if (tid < si.valid) {
            if (g.c > 1 && argmax_takes_dd(bb, ib, ba, ia)) {
                ba = bb;
                ia = ib;
            }
            const int pos = si.p0 + tid;
            const int row = pos / L.w, colx = pos - row * L.w;
            const float cx = (sigmoidf_dd(t0) + (float)colx) * L.stride;
            const float cy = (sigmoidf_dd(t1) + (float)row) * L.stride;
            const float tw = (t2 > g.scale_clamp) ? g.scale_clamp : t2;  // torch.clamp(max=): NaN stays NaN
            const float th = (t3 > g.scale_clamp) ? g.scale_clamp : t3;
            const float2 awh = L.anchors_wh[si.ai];
            const float bw = expf(tw) * awh.x, bh = expf(th) * awh.y;
            const int e = tid * A + si.ai;
            o_box[e] = make_float4(cx - 0.5f * bw, cy - 0.5f * bh, cx + 0.5f * bw, cy + 0.5f * bh);
            o_score[e] = sigmoidf_dd(t4) * (g.c > 0 ? sigmoidf_dd(ba) : 1.0f);
            o_cls[e] = g.c > 0 ? ia : 0;
        }
        if (si.ai == A - 1) {  // unit complete: kTile*A consecutive output rows, coalesced
            consumer_sync();
            const int64_t obase = (int64_t)si.img * g.out_img_stride + L.out_offset + (int64_t)si.p0 * A;
            const int ne = si.valid * A;
            for (int e = tid; e < ne; e += kTile) {
                st_stream(g.boxes_out + obase + e, o_box[e]);
                st_stream(g.score_out + obase + e, o_score[e]);
                g.class_out[obase + e] = (int64_t)o_cls[e];
            }
            consumer_sync();
        }
    }
}

// ---- flat streaming variant: one thread per (level, image, anchor, 4 consecutive positions) walks ALL 5+C planes of
// its anchor with 16-byte loads, 8 in flight, and keeps 4 independent (running max, first arg-max) chains in
// registers.  Consecutive threads read consecutive 16 bytes of the same plane (full 128-byte lines), the grid covers
// every level in one launch, and 2048 resident threads per SM keep ~260 KB per SM in flight.
constexpr int kFlatThreads = 256;

struct FlatArgs {
    DenseLevelDev lv[kMaxLevels];
    long long thread_begin[kMaxLevels + 1];  // first flat thread of each level
    int num_levels, n, a, c;
    float scale_clamp;
    int64_t out_img_stride;
    float4* boxes_out;
    float* score_out;
    int64_t* class_out;
};

__global__ void __launch_bounds__(kFlatThreads) dense_decode_flat_kernel(const __grid_constant__ FlatArgs g) {
    const long long gt = (long long)blockIdx.x * kFlatThreads + threadIdx.x;
    if (gt >= g.thread_begin[g.num_levels]) return;
    int l = 0;
#pragma unroll
    for (int q = 1; q < kMaxLevels; ++q)
        if (q < g.num_levels && gt >= g.thread_begin[q]) l = q;
    const DenseLevelDev& L = g.lv[l];
    const int groups = L.hw >> 2, A = g.a, C = g.c, nch = 5 + C;
    const long long local = gt - g.thread_begin[l];
    const int grp = (int)(local % groups);
    const int ia = (int)(local / groups);  // image * A + anchor
    const int img = ia / A, ai = ia - img * A;
    const int plane4 = groups;  // float4 stride between planes
    const float4* pl = reinterpret_cast<const float4*>(L.head + (int64_t)ia * nch * L.hw) + grp;
    float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    int bidx[4] = {0, 0, 0, 0};
    const float4* pk = pl + (int64_t)5 * plane4;
    int k = 0;
    for (; k + 8 <= C; k += 8) {
        float4 q[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) q[i] = ld_stream(pk + (int64_t)(k + i) * plane4);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            bidx[0] = (q[i].x > best[0]) ? k + i : bidx[0]; best[0] = max_nan(best[0], q[i].x);
            bidx[1] = (q[i].y > best[1]) ? k + i : bidx[1]; best[1] = max_nan(best[1], q[i].y);
            bidx[2] = (q[i].z > best[2]) ? k + i : bidx[2]; best[2] = max_nan(best[2], q[i].z);
            bidx[3] = (q[i].w > best[3]) ? k + i : bidx[3]; best[3] = max_nan(best[3], q[i].w);
        }
    }
    for (; k < C; ++k) {
        const float4 q = ld_stream(pk + (int64_t)k * plane4);
        bidx[0] = (q.x > best[0]) ? k : bidx[0]; best[0] = max_nan(best[0], q.x);
        bidx[1] = (q.y > best[1]) ? k : bidx[1]; best[1] = max_nan(best[1], q.y);
        bidx[2] = (q.z > best[2]) ? k : bidx[2]; best[2] = max_nan(best[2], q.z);
        bidx[3] = (q.w > best[3]) ? k : bidx[3]; best[3] = max_nan(best[3], q.w);
    }
    // the running max is NaN-propagating and `q > max` is false once it is NaN: a NaN column is finished by an exact
    // rescan (torch.max: the first NaN wins)
    if (best[0] != best[0] || best[1] != best[1] || best[2] != best[2] || best[3] != best[3]) {
        bool found[4] = {false, false, false, false};
        for (int j = 0; j < C; ++j) {
            const float4 q4 = pk[(int64_t)j * plane4];
            const float q[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
            for (int v = 0; v < 4; ++v)
                if (q[v] != q[v] && !found[v]) {
                    found[v] = true;
                    bidx[v] = j;
                }
        }
    }
    float t[5][4];  // box logits last: they would only occupy registers during the class loop
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        const float4 q = ld_stream(pl + (int64_t)j * plane4);
        t[j][0] = q.x; t[j][1] = q.y; t[j][2] = q.z; t[j][3] = q.w;
    }
    const float2 awh = L.anchors_wh[ai];
    const int64_t obase = (int64_t)img * g.out_img_stride + L.out_offset;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
        const int pos = grp * 4 + v;
        const int row = pos / L.w, colx = pos - row * L.w;
        const float cx = (sigmoidf_dd(t[0][v]) + (float)colx) * L.stride;
        const float cy = (sigmoidf_dd(t[1][v]) + (float)row) * L.stride;
        const float tw = (t[2][v] > g.scale_clamp) ? g.scale_clamp : t[2][v];  // torch.clamp(max=): NaN stays NaN
        const float th = (t[3][v] > g.scale_clamp) ? g.scale_clamp : t[3][v];
        const float bw = expf(tw) * awh.x, bh = expf(th) * awh.y;
        const int64_t o = obase + (int64_t)pos * A + ai;
        st_stream(g.boxes_out + o, make_float4(cx - 0.5f * bw, cy - 0.5f * bh, cx + 0.5f * bw, cy + 0.5f * bh));
        st_stream(g.score_out + o, sigmoidf_dd(t[4][v]) * (C > 0 ? sigmoidf_dd(best[v]) : 1.0f));
        g.class_out[o] = (int64_t)(C > 0 ? bidx[v] : 0);
    }
}

static size_t dense_smem_bytes(int nch, int a, int stages) {
    return (size_t)stages * nch * kTile * 4 + (size_t)kTile * a * 24;
}

}  // namespace det

using namespace det;

extern "C" {

int det_dense_decode_level(const float* head, int n, int a, int c, int h, int w, int stride, const float* anchors_wh,
                           float scale_clamp, float* boxes_out, float* score_out, int64_t* class_out,
                           int64_t out_img_stride, int64_t out_offset, void* stream);

int det_dense_decode(const det_dense_level_t* levels_host, int num_levels, int n, int a, int c, float scale_clamp,
                     float* boxes_out, float* score_out, int64_t* class_out, int64_t out_img_stride, void* stream) {
    DET_CHECK_ARG(num_levels >= 0 && n >= 0 && a >= 1 && c >= 0, "bad size");
    if (num_levels == 0 || n == 0) return DET_OK;
    DET_CHECK_ARG(levels_host && boxes_out && score_out && class_out, "null pointer");
    const int nch = 5 + c;
    // the bulk-copy pipeline needs 16-byte granules: every level with hw % 4 == 0 and an aligned base pointer
    bool tma_ok = num_levels <= kMaxLevels && aligned16(boxes_out);
    int64_t units = 0;
    for (int l = 0; l < num_levels && tma_ok; ++l) {
        const det_dense_level_t& L = levels_host[l];
        DET_CHECK_ARG(L.head && L.anchors_wh && L.h >= 0 && L.w >= 0 && L.out_offset >= 0, "bad level");
        const int64_t hw = (int64_t)L.h * L.w;
        if (hw % 4 != 0 || !aligned16(L.head) || hw >= (1ll << 30)) tma_ok = false;
        units += (int64_t)n * ((hw + kTile - 1) / kTile);
    }
    if (units >= (1ll << 31) || units == 0) tma_ok = false;
    if (!tma_ok) {
        for (int l = 0; l < num_levels; ++l) {
            const det_dense_level_t& L = levels_host[l];
            const int rc = det_dense_decode_level(L.head, n, a, c, L.h, L.w, L.stride, L.anchors_wh, scale_clamp, boxes_out,
                                                  score_out, class_out, out_img_stride, L.out_offset, stream);
            if (rc != DET_OK) return rc;
        }
        return DET_OK;
    }
    // DET_DENSE_PIPELINE=1 selects the bulk-async-copy pipeline (kept for comparison; the flat LDG stream is faster, see
    // profiles/README.md)
    static const bool use_pipeline = [] {
        const char* e = getenv("DET_DENSE_PIPELINE");
        return e && e[0] == '1';
    }();
    int stages = kMaxStages;
    while (stages > 2 && dense_smem_bytes(nch, a, stages) > 200 * 1024) --stages;
    if (!use_pipeline || dense_smem_bytes(nch, a, stages) > 200 * 1024) {
        FlatArgs f;
        f.num_levels = num_levels; f.n = n; f.a = a; f.c = c; f.scale_clamp = scale_clamp; f.out_img_stride = out_img_stride;
        f.boxes_out = reinterpret_cast<float4*>(boxes_out); f.score_out = score_out; f.class_out = class_out;
        long long tb = 0;
        for (int l = 0; l < kMaxLevels; ++l) {
            DenseLevelDev& D = f.lv[l];
            f.thread_begin[l] = tb;
            if (l < num_levels) {
                const det_dense_level_t& L = levels_host[l];
                D.head = L.head; D.anchors_wh = reinterpret_cast<const float2*>(L.anchors_wh);
                D.w = L.w > 0 ? L.w : 1; D.hw = L.h * L.w; D.tiles = 0; D.unit_begin = 0;
                D.stride = (float)L.stride; D.out_offset = L.out_offset;
                DET_CHECK_ARG(out_img_stride >= L.out_offset + (int64_t)D.hw * a, "output slot out of range");
                tb += (long long)n * a * (D.hw / 4);
            } else {
                D.head = nullptr; D.anchors_wh = nullptr; D.w = 1; D.hw = 4; D.tiles = 0; D.unit_begin = 0;
                D.stride = 0.f; D.out_offset = 0;
            }
        }
        f.thread_begin[kMaxLevels] = tb;
        for (int l = num_levels; l <= kMaxLevels; ++l) f.thread_begin[l] = tb;
        if (tb == 0) return DET_OK;
        const long long blocks = (tb + kFlatThreads - 1) / kFlatThreads;
        DET_CHECK_ARG(blocks < (1ll << 31), "too many positions");
        dense_decode_flat_kernel<<<(unsigned)blocks, kFlatThreads, 0, as_stream(stream)>>>(f);
        DET_LAUNCH_OK("dense_decode_flat_kernel");
        return DET_OK;
    }
    DenseArgs g;
    g.num_levels = num_levels; g.n = n; g.a = a; g.c = c; g.stages = stages;
    g.scale_clamp = scale_clamp; g.out_img_stride = out_img_stride;
    g.boxes_out = reinterpret_cast<float4*>(boxes_out); g.score_out = score_out; g.class_out = class_out;
    int ub = 0;
    for (int l = 0; l < kMaxLevels; ++l) {
        DenseLevelDev& D = g.lv[l];
        if (l < num_levels) {
            const det_dense_level_t& L = levels_host[l];
            D.head = L.head; D.anchors_wh = reinterpret_cast<const float2*>(L.anchors_wh);
            D.w = L.w > 0 ? L.w : 1; D.hw = L.h * L.w; D.tiles = (D.hw + kTile - 1) / kTile; D.unit_begin = ub;
            D.stride = (float)L.stride; D.out_offset = L.out_offset;
            DET_CHECK_ARG(out_img_stride >= L.out_offset + (int64_t)D.hw * a, "output slot out of range");
            ub += n * D.tiles;
        } else {
            D.head = nullptr; D.anchors_wh = nullptr; D.w = 1; D.hw = 0; D.tiles = 1; D.unit_begin = 0x7fffffff;
            D.stride = 0.f; D.out_offset = 0;
        }
    }
    g.total_units = ub;
    const size_t smem = dense_smem_bytes(nch, a, stages);
    cudaError_t e = cudaFuncSetAttribute(dense_decode_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(dense_decode_tma_kernel)");
    const int grid = ub < sm_count() ? ub : sm_count();
    dense_decode_tma_kernel<<<grid, kTile + 32, smem, as_stream(stream)>>>(g);
    DET_LAUNCH_OK("dense_decode_tma_kernel");
    return DET_OK;
}

}  // extern "C"
