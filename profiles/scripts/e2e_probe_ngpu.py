"""Where does the end-to-end step go at N GPUs?  Run with torchrun (one rank per GPU) or plain python (N = 1):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        profiles/scripts/e2e_probe_ngpu.py --out gpurun_out/e2e_probe_nN.json

All ranks run every section at the same time (barrier before each), so the numbers are what a rank gets while its
peers load the same host: (1) raw pinned-copy bandwidth per direction and both directions at once, at the step's
sizes (1 505 280 B up, 1 844 224 B down, 8 copies in flight) and at 64 MiB; (2) the same with the host buffers
allocated after the rank pinned itself to its own slice of the cores (first touch), with one portable registered arena,
and with write-combined memory for the upload; (3) YoloHostPipeline steps per second under the same variants.
Rank 0 writes one JSON with per-rank and aggregate figures and the box's topology."""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "object-detection-pytorch-rust_b200"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import det_b200 as det  # noqa: E402

H2D_STEP, D2H_STEP, BIG = 1505280, 1844224, 64 << 20
rt = ctypes.CDLL("libcudart.so.12")
rt.cudaHostAlloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t, ctypes.c_uint]
rt.cudaFreeHost.argtypes = [ctypes.c_void_p]
rt.cudaHostRegister.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint]
rt.cudaHostUnregister.argtypes = [ctypes.c_void_p]
rt.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]


def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=20).stdout.strip()
    except Exception as e:  # noqa: BLE001
        return f"{type(e).__name__}: {e}"


class HostBuf:
    """nbytes of page-locked host memory: kind = 'pinned' (cudaHostAlloc default), 'wc' (write-combined), 'registered'
    (malloc + first touch by this thread + cudaHostRegisterPortable)."""

    def __init__(self, nbytes, kind):
        self.kind, self.nbytes = kind, nbytes
        p = ctypes.c_void_p()
        if kind == "registered":
            self.raw = (ctypes.c_uint8 * (nbytes + 4096))()
            base = (ctypes.addressof(self.raw) + 4095) & ~4095
            ctypes.memset(base, 1, nbytes)  # first touch on this rank's cores
            assert rt.cudaHostRegister(base, nbytes, 1) == 0  # cudaHostRegisterPortable
            self.ptr = base
        else:
            assert rt.cudaHostAlloc(ctypes.byref(p), nbytes, 4 if kind == "wc" else 0) == 0
            self.ptr = p.value
            ctypes.memset(self.ptr, 1, nbytes)

    def free(self):
        if self.kind == "registered":
            rt.cudaHostUnregister(self.ptr)
        else:
            rt.cudaFreeHost(self.ptr)


def copy_rate(dev, world, nbytes_up, nbytes_down, kind_up="pinned", kind_down="pinned", inflight=8, target_ms=150.0):
    """GB/s of this rank: `inflight` copies per direction queued round-robin on `inflight` streams per direction."""
    ups = [HostBuf(nbytes_up, kind_up) for _ in range(inflight)] if nbytes_up else []
    downs = [HostBuf(nbytes_down, kind_down) for _ in range(inflight)] if nbytes_down else []
    d_up = [torch.empty(nbytes_up, dtype=torch.uint8, device=dev) for _ in ups]
    d_down = [torch.empty(nbytes_down, dtype=torch.uint8, device=dev) for _ in downs]
    s_up = [torch.cuda.Stream() for _ in ups]
    s_down = [torch.cuda.Stream() for _ in downs]

    def run(iters):
        for i in range(iters):
            k = i % inflight
            if ups:
                rt.cudaMemcpyAsync(d_up[k].data_ptr(), ups[k].ptr, nbytes_up, 1, s_up[k].cuda_stream)
            if downs:
                rt.cudaMemcpyAsync(downs[k].ptr, d_down[k].data_ptr(), nbytes_down, 2, s_down[k].cuda_stream)

    run(2 * inflight)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run(4 * inflight)
    torch.cuda.synchronize()
    per = (time.perf_counter() - t0) / (4 * inflight)
    iters = max(4 * inflight, int(target_ms * 1e-3 / max(per, 1e-6)))
    if world > 1:
        t = torch.tensor([iters], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        iters = int(t.item())
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run(iters)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    for b in ups + downs:
        b.free()
    return {"up_gbs": nbytes_up * iters / dt / 1e9, "down_gbs": nbytes_down * iters / dt / 1e9, "us_per_pair": dt / iters * 1e6}


def pipeline_rate(dev, world, depth=8, steps=3000):
    yh = det.YoloGridHead(7, 2, 20, (448, 448))
    pipe = det.YoloHostPipeline(yh, 256, 0.25, 0.5, 300, depth=depth, device=dev, index_dtype=torch.int32)
    src = torch.randn(depth, 256, 7, 7, 30)
    for s in range(depth):
        pipe.input(s).copy_(src[s])

    def run(k):
        for i in range(k):
            slot = i % depth
            if i >= depth:
                int(pipe.wait(slot)["count"][0])
            pipe.launch(slot)
        for s in range(depth):
            pipe.wait(s)

    run(200)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run(steps)
    dt = time.perf_counter() - t0
    return {"us_per_step": dt / steps * 1e6, "images_per_s": 256 * steps / dt}


def gather(world, obj):
    if world == 1:
        return [obj]
    out = [None] * world
    dist.all_gather_object(out, obj)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/e2e_probe.json")
    args = ap.parse_args()
    rank, world, local = det.dist.init_from_env()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    torch.empty(1, device=dev)
    ncpu = os.cpu_count() or 1
    full_aff = sorted(os.sched_getaffinity(0))
    res = {"n_gpus": world, "cpu_count": ncpu, "affinity_default": [full_aff[0], full_aff[-1], len(full_aff)]}
    if rank == 0:
        res["topology"] = {"nvidia_smi_topo": sh("nvidia-smi topo -m"), "lscpu": sh("lscpu | grep -i -E 'model name|numa|socket|^cpu\\(s\\)|thread'"),
                           "gpu_numa": sh("for d in /sys/bus/pci/devices/*; do c=$(cat $d/class); if [ \"$c\" = 0x030200 ]; then echo $(basename $d) numa=$(cat $d/numa_node) link=$(cat $d/current_link_speed 2>/dev/null) x$(cat $d/current_link_width 2>/dev/null); fi; done"),
                           "meminfo": sh("grep -E 'MemTotal|HugePages_Total' /proc/meminfo")}
    sections = {}

    def section(name, fn):
        if world > 1:
            dist.barrier()
        mine = fn()
        allr = gather(world, mine)
        agg = {}
        for k in allr[0]:
            vals = [r[k] for r in allr]
            agg[k] = {"sum": sum(vals), "min": min(vals), "max": max(vals)}
        sections[name] = {"per_rank": allr, "aggregate": agg}

    for tag, up, down in (("step_sizes", H2D_STEP, D2H_STEP), ("64MiB", BIG, BIG)):
        section(f"{tag}:h2d_only", lambda: copy_rate(dev, world, up, 0))
        section(f"{tag}:d2h_only", lambda: copy_rate(dev, world, 0, down))
        section(f"{tag}:both", lambda: copy_rate(dev, world, up, down))
    section("pipeline:default", lambda: pipeline_rate(dev, world))
    # variant 1: every rank on its own slice of the cores, host buffers allocated (first touched) after that
    per = max(1, len(full_aff) // world)
    mine = full_aff[rank * per:(rank + 1) * per] or full_aff
    os.sched_setaffinity(0, mine)
    section("step_sizes:both:own_cores", lambda: copy_rate(dev, world, H2D_STEP, D2H_STEP))
    section("step_sizes:both:own_cores+registered_portable", lambda: copy_rate(dev, world, H2D_STEP, D2H_STEP, "registered", "registered"))
    section("step_sizes:both:own_cores+wc_upload", lambda: copy_rate(dev, world, H2D_STEP, D2H_STEP, "wc", "pinned"))
    section("pipeline:own_cores", lambda: pipeline_rate(dev, world))
    os.sched_setaffinity(0, full_aff)
    res["sections"] = sections
    if rank == 0:
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        json.dump(res, open(args.out, "w"), indent=1)
        for k, v in sections.items():
            a = v["aggregate"]
            print(k, {kk: round(vv["sum"], 1) for kk, vv in a.items() if kk.endswith("gbs") or kk == "images_per_s"},
                  {kk: (round(vv["min"], 1), round(vv["max"], 1)) for kk, vv in a.items() if kk.startswith("us_")})
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
