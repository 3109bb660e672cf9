// Shared device/host helpers for libdet_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "../../include/det_b200.h"

namespace det {

// ---- host-side error plumbing -------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define DET_CHECK_ARG(cond, msg)                       \
    do {                                               \
        if (!(cond)) {                                 \
            det::set_error("bad argument: %s", msg);   \
            return DET_ERR_BAD_ARG;                    \
        }                                              \
    } while (0)

#define DET_LAUNCH_OK(what)                                            \
    do {                                                               \
        cudaError_t e__ = cudaGetLastError();                          \
        if (e__ != cudaSuccess) return det::cuda_fail(e__, what);      \
    } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
int sm_count();

// largest float f with (double)f <= thr: then for every float v, ((double)v > thr) == (v > f).
// torchvision's CPU NMS compares the fp32 IoU against a *double* threshold.
static inline float float_threshold_below(double thr) {
    float f = (float)thr;
    if ((double)f > thr) f = nextafterf(f, -INFINITY);
    return f;
}

// ---- device helpers -----------------------------------------------------------------------------
#ifdef __CUDACC__

// std::max / std::min NaN behaviour (what torchvision's CPU NMS evaluates): (a<b)?b:a and (b<a)?b:a
__device__ __forceinline__ float max_std(float a, float b) { return (a < b) ? b : a; }
__device__ __forceinline__ float min_std(float a, float b) { return (b < a) ? b : a; }

// torch.minimum / torch.maximum / clamp NaN behaviour (propagate NaN): single FMNMX.NAN each
__device__ __forceinline__ float max_nan(float a, float b) {
    float r;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ float min_nan(float a, float b) {
    float r;
    asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}

// streaming (evict-first) vector accesses for write-once / read-once data
__device__ __forceinline__ void st_stream(float4* p, float4 v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream(float* p, float v) { __stcs(p, v); }
__device__ __forceinline__ float4 ld_stream(const float4* p) { return __ldcs(p); }
__device__ __forceinline__ float ld_stream(const float* p) { return __ldcs(p); }

__device__ __forceinline__ float box_area(const float4 b) { return (b.z - b.x) * (b.w - b.y); }

// 32-bit key that sorts ASCENDING into torch's descending-score order:
// NaN first (all NaNs tie), then +inf ... -inf, -0 == +0 tie.
__device__ __forceinline__ uint32_t score_desc_key(float s) {
    if (s != s) return 0u;
    uint32_t u = __float_as_uint(s);
    if (u == 0x80000000u) u = 0u;
    uint32_t asc = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    return ~asc;
}

// exact greedy-NMS predicate of torchvision's CPU kernel: does kept box `a` suppress candidate `b`?
// thr_f = float_threshold_below(double threshold).
__device__ __forceinline__ bool nms_suppresses(const float4 a, float area_a, const float4 b, float area_b,
                                               float thr_f) {
    float xx1 = max_std(a.x, b.x), yy1 = max_std(a.y, b.y);
    float xx2 = min_std(a.z, b.z), yy2 = min_std(a.w, b.w);
    float w = max_std(0.0f, xx2 - xx1), h = max_std(0.0f, yy2 - yy1);
    float inter = w * h;
    float ovr = inter / (area_a + area_b - inter);
    return ovr > thr_f;
}

// reference pairwise IoU of one pair (boxes.py:173-214), NaN-propagating min/max like torch
__device__ __forceinline__ float pair_intersection(const float4 a, const float4 b) {
    float w = max_nan(min_nan(a.z, b.z) - max_nan(a.x, b.x), 0.0f);
    float h = max_nan(min_nan(a.w, b.w) - max_nan(a.y, b.y), 0.0f);
    return w * h;
}
__device__ __forceinline__ float pair_iou(const float4 a, float area_a, const float4 b, float area_b) {
    float inter = pair_intersection(a, b);
    return (inter > 0.0f) ? inter / (area_a + area_b - inter) : 0.0f;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

#endif  // __CUDACC__
}  // namespace det
