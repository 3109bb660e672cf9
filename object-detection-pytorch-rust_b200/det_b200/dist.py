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


def global_num_images(local_n: int, device=None) -> int:
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        t = torch.tensor([local_n], dtype=torch.int64, device=device)
        dist.all_reduce(t)
        return int(t.item())
    return local_n
