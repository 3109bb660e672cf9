"""Data-parallel plumbing: the path shards by image; the only collective is one SUM all-reduce of
the 8-float loss/count vector per training step (NCCL over NVLink on GPUs, gloo in the CPU tests).  Inference
(decode + NMS) needs no collective."""
import os
from typing import Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: str = None) -> Tuple[int, int, int]:
    """(rank, world_size, local_rank) from torchrun's environment; initialises the process group when world_size > 1."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend=backend, **kw)
    return rank, world, local


def shard_range(num_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous batch shard [lo, hi) of rank `rank` (remainder spread over the first ranks)."""
    base, rem = divmod(num_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_sums_(sums: torch.Tensor) -> torch.Tensor:
    """In-place SUM all-reduce of the loss/count vector over all ranks (no-op for a single process).
    The local sums must already be scaled by the GLOBAL normaliser (pass num_images_global to the loss)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    return sums


class _Done:
    """Stand-in work handle for the single-process case."""

    def wait(self):
        return True


def allreduce_sums_async(sums: torch.Tensor):
    """Same reduction, not waited for: returns a handle whose ``wait()`` makes the reduced values visible to the
    current stream.  The loss scalars are only logged, so a training loop can wait one step late and keep the
    collective's latency off the critical path (gradients of this path stay local to the rank)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist.all_reduce(sums, op=dist.ReduceOp.SUM, async_op=True)
    return _Done()


class SumsReducer:
    """Logging-cadence reduction of the loss/count vector: every step's local sums are accumulated on the device and
    the accumulated vector is all-reduced (asynchronously) once every ``every`` steps.  The loss scalars are only
    reported, never fed back into the gradients of this path, so a short training step does not have to pay a
    collective's launch latency each iteration.  ``every=1`` is the per-step all-reduce."""

    def __init__(self, every: int = 16):
        self.every = max(1, int(every))
        self._acc = None
        self._steps = 0
        self._work = None
        self.last = None     # the most recent reduced vector (valid after the stream has passed its all-reduce)
        self.last_steps = 0  # number of steps it covers

    def add(self, sums: torch.Tensor):
        """Account for one step; returns the reduced vector when this call completed a window, else None."""
        if self._acc is None:
            self._acc = torch.zeros_like(sums)
        self._acc += sums.detach()
        self._steps += 1
        if self._steps < self.every:
            return None
        if self._work is not None:
            self._work.wait()
        out, self._acc = self._acc, None
        self._work = allreduce_sums_async(out)
        self.last, self.last_steps, self._steps = out, self.every, 0
        return out

    def flush(self):
        """Wait for the outstanding all-reduce (end of training / before reading ``last`` on the host)."""
        if self._work is not None:
            self._work.wait()
            self._work = None
        return self.last


class PeerSums:
    """SUM of the loss/count vector over the ranks through NVLink peer memory instead of an NCCL launch
    (csrc/peer.cu: det_peer_sums_publish / det_peer_sums_collect).  ``publish(sums)`` stores this rank's vector into
    every peer's symmetric buffer; ``collect()`` returns the world's sum of the PREVIOUS publish (one step late, so
    the peers' stores have normally arrived and the tiny kernel does not spin).  Two ~2 us launches per step instead
    of ~55 us of NCCL launch: what a 81 us training step needs to scale.  Single process: the same kernels on a
    plain buffer (world 1).  One instance serves ONE stream of steps: every rank must issue the same sequence of
    publish / collect / exchange / fused launches on a single stream (the slots and the fused path's arrival counter are
    not shared between concurrent launches)."""

    SLOTS = 8
    RECORD = 16

    def __init__(self, device, width: int = 8, timeout_s: float = 20.0, graph_safe: bool = False):
        """graph_safe=True keeps the step counter on the device (exchange() -> det_peer_sums_exchange_dev, and the fused
        path det_yolo_loss_peer, increment it themselves), so the training step can be captured in a CUDA graph and
        replayed -- every rank must replay the same number of times."""
        from . import _native as N
        self._device_stamps = bool(graph_safe)
        self._N = N
        self.device = torch.device(device)
        self.width = int(width)
        self.timeout_ns = int(timeout_s * 1e9)
        multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        self.world = dist.get_world_size() if multi else 1
        self.rank = dist.get_rank() if multi else 0
        numel = self.SLOTS * self.world * self.RECORD
        if multi:
            import torch.distributed._symmetric_memory as symm
            self.buf = symm.empty(numel, dtype=torch.float32, device=self.device)
            self.buf.zero_()
            try:
                self._hdl = symm.rendezvous(self.buf, dist.group.WORLD)
            except Exception:  # noqa: BLE001  (older torch: the group has to be enabled first)
                symm.enable_symm_mem_for_group(dist.group.WORLD.group_name)
                self._hdl = symm.rendezvous(self.buf, dist.group.WORLD)
            ptrs = [int(p) for p in self._hdl.buffer_ptrs]
            torch.cuda.synchronize(self.device)
            dist.barrier()  # every buffer is zeroed before anybody publishes into it
        else:
            self.buf = torch.zeros(numel, dtype=torch.float32, device=self.device)
            ptrs = [self.buf.data_ptr()]
        self.peers = torch.tensor(ptrs, dtype=torch.int64, device=self.device)
        self.error = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.step = 0          # stamp of the latest publish
        self._collected = 0    # stamp of the latest collect
        self._out = [torch.zeros(self.width, dtype=torch.float32, device=self.device) for _ in range(2)]

    def publish(self, sums: torch.Tensor) -> int:
        N = self._N
        assert sums.is_cuda and sums.dtype == torch.float32 and sums.numel() >= self.width
        self.step += 1
        with torch.cuda.device(self.device):
            N.call("det_peer_sums_publish", N.ptr(sums), self.width, self.rank, self.world, N.ptr(self.peers), self.SLOTS,
                   self.step % self.SLOTS, self.step & 0xffffffff, N.stream())
        return self.step

    def collect(self, step: int = None) -> torch.Tensor:
        """World sum of publish number `step` (default: the oldest one not collected yet).  Asynchronous: the returned
        tensor is valid in stream order; it is overwritten two collects later."""
        N = self._N
        step = self._collected + 1 if step is None else int(step)
        assert 0 < step <= self.step and self.step - step < self.SLOTS - 3, "collect within 4 steps of the publish"
        out = self._out[step & 1]
        with torch.cuda.device(self.device):
            N.call("det_peer_sums_collect", N.ptr(out), self.width, self.world, N.ptr(self.buf), self.SLOTS,
                   step % self.SLOTS, step & 0xffffffff, self.timeout_ns, N.ptr(self.error), N.stream())
        self._collected = step
        return out

    def exchange(self, sums: torch.Tensor):
        """One launch per step: publish this step's vector and collect the previous step's world sum (None on the
        first step).  The returned tensor is valid in stream order and overwritten two steps later."""
        N = self._N
        assert sums.is_cuda and sums.dtype == torch.float32 and sums.numel() >= self.width
        assert self._collected == self.step, "do not mix exchange() with publish() / collect()"
        if self._device_stamps:
            # graph-safe form: the step stamp is a device counter the kernel increments itself, `out` is one fixed buffer
            # that every step overwrites with the previous step's world sum (zeros until the second step)
            if not hasattr(self, "_stamp_dev"):
                self._stamp_dev = torch.zeros(1, dtype=torch.int32, device=self.device)
            out = self._out[0]
            with torch.cuda.device(self.device):
                N.call("det_peer_sums_exchange_dev", N.ptr(sums), N.ptr(out), self.width, self.rank, self.world,
                       N.ptr(self.peers), self.SLOTS, N.ptr(self._stamp_dev), 1, self.timeout_ns, N.ptr(self.error), N.stream())
            return out
        self.step += 1
        out = self._out[self.step & 1]
        with torch.cuda.device(self.device):
            N.call("det_peer_sums_exchange", N.ptr(sums), N.ptr(out), self.width, self.rank, self.world, N.ptr(self.peers),
                   self.SLOTS, self.step & 0xffffffff, 1, self.timeout_ns, N.ptr(self.error), N.stream())
        self._collected = self.step
        return out if self.step > 1 else None

    def fused_ctx(self):
        """Context for a kernel that publishes from its own last CTA (det_yolo_loss_peer): advances the step like
        exchange() and returns (det_peer_ctx_t, out) -- `out` receives the previous step's world sum of the RAW sums
        vector (None on the first step), valid in stream order after that kernel.
        With graph_safe=True (constructor) the step counter lives on the device and `out` is a single buffer that every
        step overwrites (zeros after the first step)."""
        N = self._N
        assert self._collected == self.step, "do not mix the fused path with publish() / collect()"
        if not hasattr(self, "_stamp_dev"):
            self._stamp_dev = torch.zeros(1, dtype=torch.int32, device=self.device)
        ctx = N.PeerCtx()
        ctx.peers_dev = self.peers.data_ptr()
        ctx.error_flag, ctx.done_counter = self.error.data_ptr(), None  # the ticket lives in the loss accumulators
        ctx.timeout_ns = self.timeout_ns
        ctx.width, ctx.rank, ctx.world, ctx.slots = self.width, self.rank, self.world, self.SLOTS
        ctx.lag = 1
        if self._device_stamps:
            out = self._out[0]
            ctx.out, ctx.stamp, ctx.stamp_counter = out.data_ptr(), 0, self._stamp_dev.data_ptr()
            return ctx, out
        assert not torch.cuda.is_current_stream_capturing(), "capture needs PeerSums(..., graph_safe=True)"
        self.step += 1
        out = self._out[self.step & 1]
        ctx.out, ctx.stamp, ctx.stamp_counter = out.data_ptr(), self.step & 0xffffffff, None
        self._collected = self.step
        return ctx, (out if self.step > 1 else None)

    def flush(self) -> torch.Tensor:
        """World sum of the latest published vector (end of training / before logging the last step)."""
        N = self._N
        step = self.step
        if self._device_stamps and hasattr(self, "_stamp_dev"):  # the step counter lives on the device (graph replays)
            step = int(self._stamp_dev.item())
        if step == 0:
            return torch.zeros(self.width, dtype=torch.float32, device=self.device)
        out = self._out[1] if self._device_stamps else self._out[(step + 1) & 1]
        with torch.cuda.device(self.device):
            N.call("det_peer_sums_collect", N.ptr(out), self.width, self.world, N.ptr(self.buf), self.SLOTS,
                   step % self.SLOTS, step & 0xffffffff, self.timeout_ns, N.ptr(self.error), N.stream())
        return out

    def check(self):
        """Host-side: raise if a collect ever timed out waiting for a peer (synchronises)."""
        if int(self.error.item()) != 0:
            raise RuntimeError("PeerSums: a peer did not publish within the timeout")


def global_num_images(local_n: int, device=None) -> int:
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        t = torch.tensor([local_n], dtype=torch.int64, device=device)
        dist.all_reduce(t)
        return int(t.item())
    return local_n
