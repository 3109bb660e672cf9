// Batched category-partitioned NMS for a whole batch of images (subsystem 3).
//
// Replaces batched_nms (python/src/utils.py:96-119) -> torchvision.ops.batched_nms / nms, which the reference
// calls once per image from a Python loop (python/src/models/utils.py:74-95).  Here the batch is one call:
//   small path  (m_max <= 4096): ONE launch, one CTA per image, everything in shared memory;
//   large path  (m_max <= 131071): global merge sort + persistent segment kernels (see nms_core.cuh).
// Per image the reference's CPU branch rule is reproduced: count <= 1000 boxes (numel <= 4000) -> torchvision's
// coordinate-offset trick, evaluated in fp32 exactly as  box + float(category) * (max_coordinate + 1) ; otherwise
// per-category NMS on the raw boxes.  Output order: descending score, ties by lower index (stable).
#include "nms_core.cuh"
#include "nms_small.cuh"
#include "nms_large.cuh"

namespace det {

constexpr int kSmallThreads = 256;

template <int CAP>
__global__ void __launch_bounds__(kSmallThreads)
nms_small_kernel(const float4* __restrict__ boxes, const float* __restrict__ scores,
                 const int64_t* __restrict__ cats, const int32_t* __restrict__ counts, int64_t m_max, float thr_f,
                 int mode, int64_t max_out, int64_t* __restrict__ keep, int32_t* __restrict__ keep_counts) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SmallSmem<CAP, kSmallThreads>& sm = *reinterpret_cast<SmallSmem<CAP, kSmallThreads>*>(smem_raw);
    using KL = KeyLayout<kSmallIdxBits>;
    const int img = blockIdx.x;
    int cnt = counts ? counts[img] : (int)m_max;
    cnt = max(0, min(cnt, (int)m_max));
    const int cap_out = (int)min(max_out, (int64_t)CAP);
    const GlobalCandidates src{boxes + (int64_t)img * m_max, scores + (int64_t)img * m_max,
                               cats ? cats + (int64_t)img * m_max : nullptr};
    const int kept = small_nms_body<CAP, kSmallThreads>(sm, src, cnt, thr_f, mode, cap_out);
    const int nout = kept < 0 ? 0 : min(kept, cap_out);
    for (int j = threadIdx.x; j < nout; j += kSmallThreads)
        keep[(int64_t)img * max_out + j] = (int64_t)KL::idx(sm.keys[j]);
    if (threadIdx.x == 0) keep_counts[img] = kept < 0 ? -1 : nout;
}

template <int CAP>
static int launch_small(const float* boxes, const float* scores, const int64_t* cats, const int32_t* counts, int n,
                        int64_t m_max, float thr_f, int mode, int64_t max_out, int64_t* keep, int32_t* keep_counts,
                        cudaStream_t st) {
    const size_t smem = sizeof(SmallSmem<CAP, kSmallThreads>);
    cudaError_t e = cudaFuncSetAttribute(nms_small_kernel<CAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(nms_small_kernel)");
    nms_small_kernel<CAP><<<n, kSmallThreads, smem, st>>>(reinterpret_cast<const float4*>(boxes), scores, cats,
                                                           counts, m_max, thr_f, mode, max_out, keep, keep_counts);
    DET_LAUNCH_OK("nms_small_kernel");
    return DET_OK;
}

}  // namespace det

using namespace det;

extern "C" {

int64_t det_nms_workspace_bytes(int n, int64_t m_max) {
    if (n <= 0 || m_max <= 4096) return 256;  // small path needs none; keep a non-zero size for allocators
    return LargeLayout(n, m_max).total;
}

int det_nms_batched(const float* boxes, const float* scores, const int64_t* categories, const int32_t* counts, int n,
                    int64_t m_max, double iou_threshold, int mode, int64_t max_out, int64_t* keep,
                    int32_t* keep_counts, void* workspace, int64_t workspace_bytes, void* stream) {
    DET_CHECK_ARG(n >= 0 && m_max >= 0 && max_out >= 0, "negative size");
    DET_CHECK_ARG(mode >= DET_NMS_AUTO && mode <= DET_NMS_OFFSET_TRICK, "unknown mode");
    if (n == 0) return DET_OK;
    DET_CHECK_ARG(keep_counts && (keep || max_out == 0), "null output");
    cudaStream_t st = as_stream(stream);
    if (m_max == 0 || max_out == 0) {
        cudaError_t e = cudaMemsetAsync(keep_counts, 0, sizeof(int32_t) * (size_t)n, st);
        if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync");
        return DET_OK;
    }
    DET_CHECK_ARG(boxes && scores, "null input");
    if (!aligned16(boxes)) {
        set_error("boxes must be 16-byte aligned");
        return DET_ERR_ALIGN;
    }
    if (m_max > (1 << kLargeIdxBits) - 1) {
        set_error("m_max %lld exceeds the limit 131071", (long long)m_max);
        return DET_ERR_UNSUPPORTED;
    }
    const float thr_f = float_threshold_below(iou_threshold);
    if (m_max <= 1024)
        return launch_small<1024>(boxes, scores, categories, counts, n, m_max, thr_f, mode, max_out, keep, keep_counts, st);
    if (m_max <= 2048)
        return launch_small<2048>(boxes, scores, categories, counts, n, m_max, thr_f, mode, max_out, keep, keep_counts, st);
    if (m_max <= 4096)
        return launch_small<4096>(boxes, scores, categories, counts, n, m_max, thr_f, mode, max_out, keep, keep_counts, st);
    LargeLayout lay(n, m_max);
    if (!workspace || workspace_bytes < lay.total) {
        set_error("workspace too small: need %lld bytes", (long long)lay.total);
        return DET_ERR_WORKSPACE;
    }
    if (!aligned16(workspace)) {
        set_error("workspace must be 16-byte aligned");
        return DET_ERR_ALIGN;
    }
    return large_nms_run(lay, workspace, boxes, scores, categories, counts, thr_f, mode, max_out, keep, keep_counts, st);
}

}  // extern "C"
