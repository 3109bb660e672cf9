// Device side of the peer-memory exchange (see peer.cu for the protocol): shared by the stand-alone kernels of peer.cu and
// by the loss kernels that publish their sums from their last CTA (assign_loss.cu).
#pragma once
#include "common.cuh"

namespace det {

constexpr int kPeerRecordWords = 16;
constexpr int kPeerStampWord = 15;
constexpr int kPeerMaxWidth = 12;
constexpr int kPeerMaxWorld = 32;

__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ void peer_publish(const float* __restrict__ sums, int width, int rank, int world,
                                             float* const* __restrict__ peers, int slot, uint32_t stamp) {
    const int r = threadIdx.x;
    if (r >= world) return;
    float* rec = peers[r] + ((int64_t)slot * world + rank) * kPeerRecordWords;
    for (int i = 0; i < width; ++i) rec[i] = sums[i];
    __threadfence_system();
    *reinterpret_cast<volatile uint32_t*>(rec + kPeerStampWord) = stamp;
}

__device__ __forceinline__ void peer_collect(float* __restrict__ out, int width, int world, const float* __restrict__ local,
                                             int slot, uint32_t stamp, long long timeout_ns,
                                             int32_t* __restrict__ error_flag) {
    __shared__ float s_val[kPeerMaxWorld][kPeerMaxWidth];
    __shared__ int s_bad;
    const int r = threadIdx.x;
    if (r == 0) s_bad = 0;
    __syncwarp();
    if (r < world) {
        const float* rec = local + ((int64_t)slot * world + r) * kPeerRecordWords;
        const volatile uint32_t* flag = reinterpret_cast<const volatile uint32_t*>(rec + kPeerStampWord);
        const unsigned long long t0 = global_timer_ns();
        bool ok = true;
        while (*flag != stamp) {
            if ((long long)(global_timer_ns() - t0) > timeout_ns) {
                ok = false;
                break;
            }
            __nanosleep(200);
        }
        __threadfence_system();
        for (int i = 0; i < width; ++i) s_val[r][i] = ok ? *reinterpret_cast<const volatile float*>(rec + i) : NAN;
        if (!ok) s_bad = 1;
    }
    __syncwarp();
    if (r < width) {
        float acc = 0.0f;
        for (int q = 0; q < world; ++q) acc += s_val[q][r];  // rank order: the same bits on every rank
        out[r] = acc;
    }
    if (r == 0 && s_bad && error_flag) *error_flag = 1;
}

// launch parameters of a fused publish + collect (mirrors det_peer_ctx_t; enabled == 0: nothing to do)
struct PeerCtxDev {
    float* const* peers;
    float* out;
    int32_t* error_flag;
    long long timeout_ns;
    int width, rank, world, slots;
    uint32_t stamp, lag;
    uint32_t* stamp_counter;  // non-null: the step stamp is ++(*stamp_counter) instead of `stamp` (CUDA-graph replays)
    int enabled;
};

// Called by EVERY CTA of a kernel after its atomicAdd()s into the accumulators `acc`: the CTA that arrives last (all of
// its threads get `true`) sees the final values in L2.  `ticket` is a device int32 that is zero before the first launch;
// the last CTA re-arms it.
__device__ __forceinline__ bool last_cta_arrives(int32_t* ticket) {
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const int t = atomicAdd(ticket, 1);
        s_last = (t == (int)(gridDim.x * gridDim.y) - 1) ? 1 : 0;
        if (s_last) *ticket = 0;
    }
    __syncthreads();
    if (s_last) __threadfence();
    return s_last != 0;
}

// Last CTA only, every thread calls it (blockDim.x >= 32): out[i] = acc[i] * scale(i) for i < 8, accumulators re-zeroed --
// so the launch needs neither a memset before nor a scaling op after it.  The finished vector is also left in
// `s_fin` (shared, 8 floats) for a peer exchange that may follow.
template <typename ScaleFn>
__device__ __forceinline__ void finalize_sums(float* __restrict__ acc, float* __restrict__ out, float* s_fin, ScaleFn scale) {
    if (threadIdx.x < 8) {
        const float v = __ldcg(acc + threadIdx.x) * scale((int)threadIdx.x);
        out[threadIdx.x] = v;
        s_fin[threadIdx.x] = v;
        acc[threadIdx.x] = 0.0f;
    }
    __syncthreads();
}

// Last CTA only: warp 0 publishes the finished vector `s_vals` (shared memory) as this step's record into every peer's
// buffer over NVLink and collects step `stamp - lag` (compute + collective in one kernel).
__device__ __forceinline__ void peer_exchange_warp0(const PeerCtxDev& pc, const float* s_vals) {
    if (threadIdx.x >= 32) return;
    uint32_t stamp = pc.stamp;
    if (pc.stamp_counter) {  // a replayed graph cannot change its kernel arguments: the step lives on the device
        if (threadIdx.x == 0) {
            stamp = *pc.stamp_counter + 1u;
            *pc.stamp_counter = stamp;
        }
        stamp = __shfl_sync(0xffffffffu, stamp, 0);
    }
    __syncwarp();
    peer_publish(s_vals, pc.width, pc.rank, pc.world, pc.peers, (int)(stamp % (uint32_t)pc.slots), stamp);
    __syncwarp();
    if (stamp > pc.lag)
        peer_collect(pc.out, pc.width, pc.world, pc.peers[pc.rank], (int)((stamp - pc.lag) % (uint32_t)pc.slots),
                     stamp - pc.lag, pc.timeout_ns, pc.error_flag);
}

}  // namespace det
