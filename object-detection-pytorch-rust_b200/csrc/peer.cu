// The training path's one collective -- the SUM of the 8-float loss/count vector over the ranks (SURVEY.md section 8e;
// the reference logs these sums per step, python/src/models/rpn.py:216-220, 238-241) -- over NVLink PEER MEMORY
// instead of an NCCL launch.
//
// Every rank owns a symmetric buffer (torch.distributed._symmetric_memory: VMM allocation mapped into every peer)
// of `slots` x `world` records of 16 words.  publish: one warp stores the rank's vector into record [slot][rank] of
// EVERY peer's buffer (plain P2P stores through NVLink / NVSwitch), fences system-wide, then stores the step stamp
// into the record's last word.  collect: lane r of one warp spins on record [slot][r] of the LOCAL buffer until its
// stamp is the expected step, then the records are summed in rank order (bitwise identical on every rank).
// Two ~2 us launches replace a ~55 us NCCL all-reduce launch, which is what limited the scaling of the 81 us
// grid-head training step (profiles/README.md).  The collect is issued one step late, so it normally finds every
// stamp in place; a spin that outlasts `timeout_ns` (a dead peer) writes NaNs, raises *error_flag and returns instead
// of hanging the GPU.  Slot reuse is safe for slots >= 4 when every rank issues publish(t), collect(t-1) in stream
// order: a rank can only run ahead of a peer by collecting, which needs that peer's publish.
#include "peer.cuh"

namespace det {

__global__ void __launch_bounds__(32)
peer_sums_publish_kernel(const float* __restrict__ sums, int width, int rank, int world, float* const* __restrict__ peers,
                         int slot, uint32_t stamp) {
    peer_publish(sums, width, rank, world, peers, slot, stamp);
}

__global__ void __launch_bounds__(32)
peer_sums_collect_kernel(float* __restrict__ out, int width, int world, const float* __restrict__ local, int slot,
                         uint32_t stamp, long long timeout_ns, int32_t* __restrict__ error_flag) {
    peer_collect(out, width, world, local, slot, stamp, timeout_ns, error_flag);
}

// publish step `stamp`, then collect step `stamp - lag` (one launch per training step)
__global__ void __launch_bounds__(32)
peer_sums_exchange_kernel(const float* __restrict__ sums, float* __restrict__ out, int width, int rank, int world,
                          float* const* __restrict__ peers, int slots, uint32_t stamp, uint32_t lag, long long timeout_ns,
                          int32_t* __restrict__ error_flag) {
    peer_publish(sums, width, rank, world, peers, (int)(stamp % (uint32_t)slots), stamp);
    __syncwarp();
    if (stamp > lag)
        peer_collect(out, width, world, peers[rank], (int)((stamp - lag) % (uint32_t)slots), stamp - lag, timeout_ns, error_flag);
}

// the same with the step stamp kept on the device (++*stamp_counter): a captured CUDA graph cannot change its kernel
// arguments from replay to replay, so the step number has to live in memory
__global__ void __launch_bounds__(32)
peer_sums_exchange_dev_kernel(const float* __restrict__ sums, float* __restrict__ out, int width, int rank, int world,
                              float* const* __restrict__ peers, int slots, uint32_t* __restrict__ stamp_counter, uint32_t lag,
                              long long timeout_ns, int32_t* __restrict__ error_flag) {
    uint32_t stamp = 0;
    if (threadIdx.x == 0) {
        stamp = *stamp_counter + 1u;
        *stamp_counter = stamp;
    }
    stamp = __shfl_sync(0xffffffffu, stamp, 0);
    peer_publish(sums, width, rank, world, peers, (int)(stamp % (uint32_t)slots), stamp);
    __syncwarp();
    if (stamp > lag)
        peer_collect(out, width, world, peers[rank], (int)((stamp - lag) % (uint32_t)slots), stamp - lag, timeout_ns, error_flag);
}

}  // namespace det

using namespace det;

extern "C" {

int det_peer_sums_publish(const float* sums, int width, int rank, int world, const void* peers_dev, int slots, int slot,
                          uint32_t stamp, void* stream) {
    DET_CHECK_ARG(sums && peers_dev, "null pointer");
    DET_CHECK_ARG(width >= 1 && width <= kPeerMaxWidth, "width must be in [1, 12]");
    DET_CHECK_ARG(world >= 1 && world <= kPeerMaxWorld && rank >= 0 && rank < world, "bad rank / world (<= 32)");
    DET_CHECK_ARG(slots >= 1 && slot >= 0 && slot < slots, "bad slot");
    peer_sums_publish_kernel<<<1, 32, 0, as_stream(stream)>>>(sums, width, rank, world,
                                                              reinterpret_cast<float* const*>(peers_dev), slot, stamp);
    DET_LAUNCH_OK("peer_sums_publish_kernel");
    return DET_OK;
}

int det_peer_sums_collect(float* out, int width, int world, const float* local_buf, int slots, int slot, uint32_t stamp,
                          int64_t timeout_ns, int32_t* error_flag, void* stream) {
    DET_CHECK_ARG(out && local_buf, "null pointer");
    DET_CHECK_ARG(width >= 1 && width <= kPeerMaxWidth, "width must be in [1, 12]");
    DET_CHECK_ARG(world >= 1 && world <= kPeerMaxWorld, "bad world (<= 32)");
    DET_CHECK_ARG(slots >= 1 && slot >= 0 && slot < slots && timeout_ns > 0, "bad slot / timeout");
    peer_sums_collect_kernel<<<1, 32, 0, as_stream(stream)>>>(out, width, world, local_buf, slot, stamp, timeout_ns,
                                                              error_flag);
    DET_LAUNCH_OK("peer_sums_collect_kernel");
    return DET_OK;
}

int det_peer_sums_exchange(const float* sums, float* out, int width, int rank, int world, const void* peers_dev, int slots,
                           uint32_t stamp, uint32_t lag, int64_t timeout_ns, int32_t* error_flag, void* stream) {
    DET_CHECK_ARG(sums && out && peers_dev, "null pointer");
    DET_CHECK_ARG(width >= 1 && width <= kPeerMaxWidth, "width must be in [1, 12]");
    DET_CHECK_ARG(world >= 1 && world <= kPeerMaxWorld && rank >= 0 && rank < world, "bad rank / world (<= 32)");
    DET_CHECK_ARG(slots >= 4 && lag >= 1 && (int)lag <= slots - 4 + 1 && timeout_ns > 0, "slots >= 4, 1 <= lag <= slots - 3");
    peer_sums_exchange_kernel<<<1, 32, 0, as_stream(stream)>>>(sums, out, width, rank, world,
                                                               reinterpret_cast<float* const*>(peers_dev), slots, stamp, lag,
                                                               timeout_ns, error_flag);
    DET_LAUNCH_OK("peer_sums_exchange_kernel");
    return DET_OK;
}

int det_peer_sums_exchange_dev(const float* sums, float* out, int width, int rank, int world, const void* peers_dev,
                               int slots, uint32_t* stamp_counter, uint32_t lag, int64_t timeout_ns, int32_t* error_flag,
                               void* stream) {
    DET_CHECK_ARG(sums && out && peers_dev && stamp_counter, "null pointer");
    DET_CHECK_ARG(width >= 1 && width <= kPeerMaxWidth, "width must be in [1, 12]");
    DET_CHECK_ARG(world >= 1 && world <= kPeerMaxWorld && rank >= 0 && rank < world, "bad rank / world (<= 32)");
    DET_CHECK_ARG(slots >= 4 && lag >= 1 && (int)lag <= slots - 4 + 1 && timeout_ns > 0, "slots >= 4, 1 <= lag <= slots - 3");
    peer_sums_exchange_dev_kernel<<<1, 32, 0, as_stream(stream)>>>(sums, out, width, rank, world,
                                                                   reinterpret_cast<float* const*>(peers_dev), slots,
                                                                   stamp_counter, lag, timeout_ns, error_flag);
    DET_LAUNCH_OK("peer_sums_exchange_dev_kernel");
    return DET_OK;
}

}  // extern "C"
