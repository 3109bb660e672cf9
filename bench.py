#!/usr/bin/env python
"""bench.py -- throughput of the post-backbone detection hot path on B200 (see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--no-extras]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Primary workload (BASELINE.json configs[1]): YOLO-style 7x7x(2*5+20) head, batch 256 per GPU, 448x448, score
threshold 0.25, per-class NMS IoU 0.5, top-300 detections per image.  A step = one batch through the fused
decode+NMS kernel.  Weak scaling: every rank owns its own batches, no collective on the inference path.

Prints ONE JSON line (rank 0).  `value` is device-timed: inputs resident in HBM and rotating over a pool larger than L2,
one CUDA graph = one pass over the pool (one kernel launch per step, on a single stream unless --lanes > 1), and the timed
region is a whole number of graph replays covering at least --steps steps and at least --min-ms milliseconds, so
`--steps 20` exercises exactly the launch path `--steps 20000` does (`timing.timed_steps` is the number actually timed,
`ms_per_step` their mean).  `e2e` goes through the public Python API with pinned host buffers, H2D + D2H inside the timed
region (same rule: at least --steps steps and --min-ms ms per segment, median of five segments after 0.3 s of warm-up).  `roofline` is the
fused kernel's algorithmic bytes / measured launch time against MEASURED_PEAKS.json, `roofline.others` the fractions of
the HBM-bound kernels of the path.  `cpu_baseline` is the CPU oracle (torch CPU ops + C greedy NMS) on a bounded sample.
`extras` carries the detail: the training step (assignment + loss fwd/bwd, BASELINE configs[2]), the dense head
(configs[3]) and pairwise IoU (configs[4]) with their rooflines; `train` (printed last, mirrored in `e2e`) is the compact
per-N training summary, with `peer_equals_nccl` = the NVLink peer-memory all-reduce checked against NCCL on the real
ranks.  The e2e pipeline moves the detection indices as int32 (same values as detect()'s int64 `flat`);
`e2e.int64_indices` is the int64 wire format.
"""
import argparse
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "object-detection-pytorch-rust_b200"))

import torch  # noqa: E402

L2_BYTES = 126 * 1024 * 1024
IMG = (448, 448)
S, B, C = 7, 2, 20
BATCH = 256
SCORE_THR, IOU_THR, MAX_DET = 0.25, 0.5, 300
WORKLOAD = "yolo7x7x30_b256_decode+perclass_nms"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def load_traffic(key):
    """per-launch DRAM bytes of the dominant kernel from the committed ncu capture (profiles/traffic.json), or None."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(p)).get(key)
    except Exception:  # noqa: BLE001
        return None


class ClockSampler:
    """Samples SM clock + throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.sm_max = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:  # noqa: BLE001
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.002)

    def start(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons),
                "samples": len(s)}


def barrier(world):
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()


def max_over_ranks(ms, world, dev):
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t.item())
    return ms


def make_heads(pool, seed, pinned=False):
    g = torch.Generator().manual_seed(seed)
    t = torch.randn(pool, BATCH, S, S, B * 5 + C, generator=g)
    return t.pin_memory() if pinned else t


# ----------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle (the reference has no YOLO head; for this workload the oracle is the CPU statement).  The path is
# independent per image, so the CPU arm gets every host core: a pool of single-threaded worker processes, each running
# the oracle's per-image loop on its slice of the batch (one process per core beats intra-op threads on 98-box images).
# ----------------------------------------------------------------------------------------------------------------
def cpu_step(O, head, priors):
    boxes, conf, scores = O.yolo_decode(head, B, C, IMG, priors)
    kept = 0
    for i in range(head.shape[0]):
        f, _, _, _ = O.yolo_select_nms(boxes[i], scores[i], SCORE_THR, IOU_THR, MAX_DET)
        kept += f.numel()
    return kept


_W = {}


def _cpu_worker_init():
    torch.set_num_threads(1)
    from oracle import ref_torch as O
    import det_b200
    _W["O"], _W["priors"] = O, det_b200.YoloGridHead(S, B, C, IMG).priors


def _cpu_worker_run(head):
    return cpu_step(_W["O"], head, _W["priors"])


class CpuPool:
    """All host cores on the oracle: `cores` single-threaded processes, the batch split evenly between them."""

    def __init__(self, cores=None):
        import multiprocessing as mp
        self.cores = cores or (os.cpu_count() or 1)
        self.pool = mp.get_context("spawn").Pool(self.cores, initializer=_cpu_worker_init)
        self.pool.map(_cpu_worker_run, [torch.randn(2, S, S, B * 5 + C) for _ in range(self.cores)])  # import + warm

    def step(self, head):
        chunks = [c for c in torch.chunk(head, self.cores) if c.shape[0]]
        return sum(self.pool.map(_cpu_worker_run, chunks))

    def close(self):
        self.pool.close()
        self.pool.join()


def cpu_images_for_budget(pool, budget_s, steps):
    head = torch.randn(8 * pool.cores, S, S, B * 5 + C, generator=torch.Generator().manual_seed(1))
    pool.step(head)
    t0 = time.perf_counter()
    pool.step(head)
    per_img = (time.perf_counter() - t0) / head.shape[0]
    n = int(budget_s / max(steps, 1) / max(per_img, 1e-7))
    return max(pool.cores, min(BATCH, n))


def run_reference_arm(args, rank, world):
    """--impl reference: the CPU implementation of the same path on all host cores (rank 0 only)."""
    if rank != 0:
        return
    pool = CpuPool()
    n_img = cpu_images_for_budget(pool, 120.0, args.steps + args.warmup)
    heads = torch.randn(4, n_img, S, S, B * 5 + C, generator=torch.Generator().manual_seed(1))
    for w in range(args.warmup):
        pool.step(heads[w % 4])
    t0 = time.perf_counter()
    for k in range(args.steps):
        pool.step(heads[k % 4])
    dt = time.perf_counter() - t0
    pool.close()
    val = n_img * args.steps / dt
    sample = (f"{n_img} of {BATCH} images per step, split over {pool.cores} single-threaded worker processes "
              "(oracle: torch CPU decode + C greedy NMS, per-image loop)")
    print(json.dumps({
        "impl": "reference", "metric": "images/sec decode+NMS", "value": val, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(),
        "timing": {"images_per_step": n_img},
        "cpu_baseline": {"value": val, "unit": "images/s", "cores": pool.cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ----------------------------------------------------------------------------------------------------------------
# extras: training step (configs[2]) and dense head (configs[3])
# ----------------------------------------------------------------------------------------------------------------
def time_region(fn, iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def time_graph(fns, iters, warm=2):
    """ms per call of the callables `fns` (one step each, e.g. one per input set of a pool) replayed from ONE CUDA graph:
    launch overhead of the Python/ctypes layer is outside the timed region, the kernels and their order are not."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):  # lazy initialisation (function attributes, allocations) before the capture
        for f in fns:
            f()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for f in fns:
            f()
    for _ in range(warm):
        graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (iters * len(fns))


def synth_gt(n, seed, dev):
    g = torch.Generator().manual_seed(seed)
    counts = torch.randint(1, 17, (n,), generator=g)
    tot = int(counts.sum())
    xy = torch.rand(tot, 2, generator=g) * 0.8 * 448
    wh = torch.rand(tot, 2, generator=g) * 0.2 * 448 + 1
    boxes = torch.cat([xy, (xy + wh).clamp(max=448.0)], 1)
    cls = torch.randint(0, 20, (tot,), generator=g)
    off = torch.zeros(n + 1, dtype=torch.int32)
    off[1:] = counts.cumsum(0).to(torch.int32)
    return boxes.to(dev), cls.to(dev), off.to(dev), tot


def extras_train(det, dev, world, peak, quick):
    out = {}
    n = 1024
    # (i) YOLO grid head: assignment (Matcher on prior boxes) + fused loc/obj/cls loss fwd+bwd, batch 1024 per GPU
    yh = det.YoloGridHead(S, B, C, IMG)
    tr = det.YoloGridTrainer(yh)
    gtb, gtc, off, tot = synth_gt(n, 2, dev)
    pool = 8 if quick else 24
    heads = [torch.randn(n, S, S, B * 5 + C, device=dev) for _ in range(pool)]
    state = {"i": 0}

    def step_grid():
        h = heads[state["i"] % pool]
        state["i"] += 1
        asg = tr.assign_packed(gtb, off, n)
        res = tr.loss(h, asg, gtc, with_grads=True)
        # the 8-float all-reduce of step i is waited for at step i+1: the loss scalars are only logged
        prev = state.get("pending")
        state["pending"] = det.dist.allreduce_sums_async(res["sums"])
        if prev is not None:
            prev.wait()

    ms = time_region(step_grid, 30 if quick else 200)
    if state.get("pending") is not None:
        state["pending"].wait()
    red = det.dist.SumsReducer(every=16)

    def step_grid_cadence():
        h = heads[state["i"] % pool]
        state["i"] += 1
        asg = tr.assign_packed(gtb, off, n)
        res = tr.loss(h, asg, gtc, with_grads=True)
        red.add(res["sums"])

    ms16 = time_region(step_grid_cadence, 32 if quick else 208)
    red.flush()
    out["train_grid_b1024"] = {"workload": "yolo7x7x30 assign+loss fwd/bwd, batch 1024/GPU", "ms_per_step": ms,
                               "images_per_s_per_gpu": n / ms * 1e3,
                               "collective": "8-float all-reduce every step, waited for one step late",
                               "ms_per_step_reduce_every_16": ms16, "images_per_s_per_gpu_reduce_every_16": n / ms16 * 1e3}
    # the same all-reduce every step through NVLink peer memory (det_b200.dist.PeerSums, csrc/peer.cu): one tiny launch
    # per step publishes this step's sums into every peer's symmetric buffer and collects the previous step's
    try:
        barrier(world)  # the exchange kernels wait for their peers on the device: start the sections together
        ps = det.dist.PeerSums(dev)

        def step_grid_peer():
            h = heads[state["i"] % pool]
            state["i"] += 1
            asg = tr.assign_packed(gtb, off, n)
            res = tr.loss(h, asg, gtc, with_grads=True)
            ps.exchange(res["sums"])

        ms_p = time_region(step_grid_peer, 30 if quick else 200)
        ps.flush()
        ps.check()
        barrier(world)
        ps2 = det.dist.PeerSums(dev)

        def step_grid_fused():
            h = heads[state["i"] % pool]
            state["i"] += 1
            asg = tr.assign_packed(gtb, off, n)
            tr.loss(h, asg, gtc, with_grads=True, peer=ps2)  # the loss kernel's last CTA does the exchange

        ms_f = time_region(step_grid_fused, 30 if quick else 200)
        ps2.flush()
        ps2.check()
        out["train_grid_b1024"].update({"ms_per_step_fused_peer_exchange": ms_f,
                                        "images_per_s_per_gpu_fused_peer_exchange": n / ms_f * 1e3,
                                        "fused_peer_exchange": "det_yolo_loss_peer: the loss kernel's last CTA publishes the "
                                                               "sums to every rank and collects the previous step's world sum"})
        # the whole step (assignment, loss fwd+bwd with the fused all-reduce, scaling) replayed from a CUDA graph: the
        # 4 launches of an 80 us step are launch-bound from Python; the step counter of the exchange lives on the device
        barrier(world)
        ps3 = det.dist.PeerSums(dev, graph_safe=True)

        def graph_step(h):
            tr.loss(h, tr.assign_packed(gtb, off, n), gtc, with_grads=True, peer=ps3)

        ms_g = time_graph([(lambda h=h_: graph_step(h)) for h_ in heads], 5 if quick else 20)
        ps3.flush()
        ps3.check()
        out["train_grid_b1024"].update({"ms_per_step_graph_fused_peer_exchange": ms_g,
                                        "images_per_s_per_gpu_graph_fused_peer_exchange": n / ms_g * 1e3,
                                        "graph": f"{pool} steps per CUDA graph replay, all-reduce every step inside the loss kernel"})
        out["train_grid_b1024"].update({"ms_per_step_peer_exchange": ms_p, "images_per_s_per_gpu_peer_exchange": n / ms_p * 1e3,
                                        "peer_exchange": "det_peer_sums_exchange: P2P stores into every rank's symmetric "
                                                         "buffer + step stamps, no NCCL launch"})
    except Exception as e:  # noqa: BLE001
        out["train_grid_b1024"]["peer_exchange_error"] = f"{type(e).__name__}: {e}"
    # (ii) reference-native RPN form: R = 50127 anchors (FPN-18 @ 448), Matcher([0.3,0.7]), 256 samples, L1 + BCE, from the
    #      head convolutions' own NCHW outputs.  The step = det_assign_sampled (row maxima -> gt-centric positive search ->
    #      sampler that labels the negatives it draws) + det_rpn_loss_sampled (forward + backward into persistent NCHW
    #      gradient buffers); the dense-label form (det_match_grid -> det_subsample_labels_grid -> loss) is timed beside it.
    strides = (4, 8, 16, 32, 64)
    rpn = det.RegionProposalNetwork(list(strides))
    hw = [(448 // s, 448 // s) for s in strides]
    anchors = torch.cat(rpn.anchor_generator.grid_anchors(hw, dev), 0)
    grid = rpn.anchor_generator.grid_layout(hw)
    R = anchors.shape[0]
    nb = 256 if quick else 1024
    gtb2, _, off2, tot2 = synth_gt(nb, 3, dev)
    obj = [torch.randn(nb, 3, h, w, device=dev) for h, w in hw]
    dlt = [torch.randn(nb, 12, h, w, device=dev) * 0.5 for h, w in hw]
    g_obj, g_dlt = [torch.zeros_like(o) for o in obj], [torch.zeros_like(d) for d in dlt]
    st = {"seed": 0, "prev": None}
    m = rpn.anchor_matcher

    def assign_only():
        matched, labels, stats = m.match_packed(gtb2, off2, nb, anchors, grid=grid, with_stats=True)
        return matched, labels, stats

    def sample_only(matched, labels, stats):
        st["seed"] += 1
        _, samples, counts = det.subsample_labels_(labels, 256, 0.5, st["seed"], stats=stats, return_samples=True)
        return det.Assignment(labels, matched, gtb2, off2, samples, counts)

    def loss_only(asg):
        sums = rpn._run_sampled(anchors, obj, dlt, asg, nb * world, None, (g_obj, g_dlt), st["prev"])
        st["prev"] = (asg.samples, asg.sample_count)
        return sums

    def assign_lazy():
        # what forward(training) runs: det_assign_sampled -- only the sampled anchors are labelled (rpn.py:132-185 reduced
        # to what `losses` reads), nothing of size R is written
        st["seed"] += 1
        return rpn.assign_sampled(anchors, gtb2, off2, nb, grid, seed=st["seed"])

    def step_with(assign_fn):
        def step():
            sums = loss_only(assign_fn())
            prev = st.get("pending")
            st["pending"] = det.dist.allreduce_sums_async(sums)  # waited for one step late: the scalars are only logged
            if prev is not None:
                prev.wait()
        return step

    def drain():
        if st.get("pending") is not None:
            st["pending"].wait()
            st["pending"] = None

    it = 5 if quick else 20
    ms_step_dense = time_region(step_with(lambda: sample_only(*assign_only())), it)
    drain()
    ms_step = time_region(step_with(assign_lazy), it)
    drain()
    ms_assign_lazy = time_region(assign_lazy, it)
    ms_match = time_region(lambda: st.__setitem__("mls", assign_only()), it)
    asg0 = sample_only(*st["mls"])
    ms_loss = time_region(lambda: loss_only(asg0), it)
    # the same step replayed from a CUDA graph (two steps per graph: the persistent gradient buffers' clear lists ping-pong)
    ms_graph = None
    try:
        keep = []

        def graph_step():
            asg = assign_lazy()
            keep.append(asg)
            loss_only(asg)

        ms_graph = time_graph([graph_step, graph_step], 3 if quick else 10)
    except Exception as e:  # noqa: BLE001
        out["train_rpn_graph_error"] = f"{type(e).__name__}: {e}"[:200]
    # ... and with the step's all-reduce inside the graph: det_peer_sums_exchange_dev (NVLink peer memory, step stamp on the
    # device) publishes this step's sums to every rank and collects the previous step's world sum -- a per-step collective
    # without an NCCL launch, which is what lets the replayed step scale
    ms_graph_peer = None
    try:
        barrier(world)
        psr = det.dist.PeerSums(dev, graph_safe=True)

        def graph_step_peer():
            asg = assign_lazy()
            keep.append(asg)
            psr.exchange(loss_only(asg))

        ms_graph_peer = time_graph([graph_step_peer, graph_step_peer], 3 if quick else 10)
        psr.flush()
        psr.check()
        barrier(world)
    except Exception as e:  # noqa: BLE001
        out["train_rpn_graph_peer_error"] = f"{type(e).__name__}: {e}"[:200]
    # the dense-output forms for their rooflines: assignment writes 9R bytes per image (labels + matched), the dense
    # fused loss writes both gradient tensors in full
    ms_assign_generic = time_region(lambda: m.match_packed(gtb2, off2, nb, anchors), it)
    logits = torch.randn(nb, R, device=dev)
    deltas = torch.randn(nb, R, 4, device=dev) * 0.5
    gl, gd = torch.empty_like(logits), torch.empty_like(deltas)
    ms_loss_dense = time_region(lambda: rpn._run_loss(anchors, logits, deltas, asg0, nb * world, None, gl, gd), it)
    del logits, deltas, gl, gd
    loss_bytes = nb * (49 * R) + 16 * tot2 + 20  # SURVEY 8d formula: every input read once, both gradients written
    # what the dense fused kernel has to move: labels (1 B) read and both gradient tensors (4 + 16 B) written for every
    # anchor; logits only where label >= 0, deltas / matched index / gt only for positives (<= 256 sampled per image)
    loss_moved = nb * (21 * R) + nb * 256 * 4 + nb * 128 * (16 + 8 + 16)
    assign_bytes = nb * 9 * R + 16 * tot2
    out["train_rpn_r50127"] = {
        "workload": f"RPN form R=50127, batch {nb}/GPU, NCHW head tensors: sample-list assignment (det_assign_sampled) + "
                    "sampled loss fwd+bwd (persistent NCHW gradient buffers) + 8-float NCCL all-reduce every step (waited "
                    "one step late) -- the launches of RegionProposalNetwork.forward(training)",
        "ms_step": ms_step, "ms_assign_sampled": ms_assign_lazy, "ms_assign": ms_match, "ms_loss_fwd_bwd": ms_loss,
        "ms_step_graph_no_collective": ms_graph,
        "ms_step_graph_peer_allreduce": ms_graph_peer,
        "images_per_s_total_graph_peer_allreduce": (world * nb / ms_graph_peer * 1e3) if ms_graph_peer else None,
        "ms_step_dense_labels": ms_step_dense,
        "dense_labels_note": "ms_step_dense_labels / ms_assign: the same step with det_match_grid writing labels + matched "
                             "index of ALL anchors (the reference signature of label_and_sample_anchors) + subsample",
        "batch": nb,
        "images_per_s_per_gpu": nb / ms_step * 1e3, "images_per_s_total": world * nb / ms_step * 1e3,
        "ms_assign_generic_anchors": ms_assign_generic, "ms_loss_dense_fwd_bwd": ms_loss_dense,
        "loss_roofline": {"bound": "hbm", "achieved": loss_moved / ms_loss_dense / 1e6, "peak": peak, "unit": "GB/s",
                          "frac": loss_moved / ms_loss_dense / 1e6 / peak, "algorithmic_bytes": loss_moved,
                          "kernel": "rpn_loss_kernel (dense gradients, det_rpn_loss)",
                          "formula": "N*(21R) + sampled rows: label read + grad_logits/grad_deltas written everywhere, "
                                     "logits/deltas/targets only where sampled",
                          "survey_formula_bytes": loss_bytes, "survey_formula_gbs": loss_bytes / ms_loss_dense / 1e6,
                          "note": "the SURVEY 8d figure (49R+16G+20) counts reads the fused kernel skips, so it exceeds "
                                  "the HBM peak; frac uses the bytes actually required.  The training step itself uses "
                                  "the sampled kernel (ms_loss_fwd_bwd), which moves O(samples) bytes"},
        "assign_roofline": {"bound": "hbm", "achieved": assign_bytes / ms_match / 1e6, "peak": peak, "unit": "GB/s",
                            "frac": assign_bytes / ms_match / 1e6 / peak, "algorithmic_bytes": assign_bytes,
                            "kernel": "match_rowmax_kernel + match_grid_kernel (det_match_grid)",
                            "formula": "N*9R + 16G: labels int8 + matched int64 written, gt read (SURVEY 8d)"},
    }
    return out


def extras_dense(det, dev, peak, quick):
    C80 = 80
    strides = [8, 16, 32]
    wh = [[[10, 13], [16, 30], [33, 23]], [[30, 61], [62, 45], [59, 119]], [[116, 90], [156, 198], [373, 326]]]
    dh = det.DenseAnchorHead(strides, wh, C80)
    R = 25200
    out = {}
    for n in ((8,) if quick else (32, 256)):
        g = torch.Generator(device=dev).manual_seed(3)
        pool = 2 if (quick or n > 64) else 4  # 297 MB (n=32) / 2.4 GB (n=256) per set: every set is larger than L2
        heads = []
        for _ in range(pool):
            hs = [torch.randn(n, 3 * (5 + C80), 640 // s, 640 // s, device=dev, generator=g) for s in strides]
            for h in hs:
                h.view(n, 3, 5 + C80, h.shape[2], h.shape[3])[:, :, 4] -= 4.0  # objectness bias (SURVEY Cfg4)
            heads.append(hs)
        outbuf = (torch.empty((n, R, 4), device=dev), torch.empty((n, R), device=dev),
                  torch.empty((n, R), dtype=torch.int64, device=dev))
        st = {"i": 0}

        def step_decode():
            dh.decode(heads[st["i"] % pool], out=outbuf)
            st["i"] += 1

        ms_d = time_region(step_decode, 5 if quick else 20)
        dec_bytes = n * 9273600
        entry = {"workload": f"dense head 3x(80^2+40^2+20^2) anchors x 80 classes, batch {n}: det_dense_decode (one launch)",
                 "ms_decode": ms_d,
                 "decode_roofline": {"bound": "hbm", "achieved": dec_bytes / ms_d / 1e6, "peak": peak, "unit": "GB/s",
                                     "frac": dec_bytes / ms_d / 1e6 / peak, "algorithmic_bytes": dec_bytes,
                                     "formula": "N * (4*R*(5+C) + 28*R), R=25200, C=80 (SURVEY 8d)"}}
        if n <= 32:
            boxes, scores, classes = outbuf
            keep, cnt = det.nms_images(boxes, scores, classes, None, 0.5, 1000)

            def step_nms():
                det.nms_images(boxes, scores, classes, None, 0.5, 1000)

            ms_n = time_graph([step_nms], 3 if quick else 20)  # ~20 launches per call: replayed from a CUDA graph
            k_tot = int(cnt.sum())
            nms_bytes = n * 28 * R + 8 * k_tot + 8 * n
            entry.update({"ms_nms": ms_n, "images_per_s_decode_plus_nms": n / (ms_d + ms_n) * 1e3,
                          "nms_25200_boxes": {"ms_per_image": ms_n / n, "kept_per_image": k_tot / n,
                                              "algorithmic_bytes": nms_bytes, "achieved_gbs": nms_bytes / ms_n / 1e6,
                                              "note": "all 25200 boxes of every image handed to det_nms_batched, 80 "
                                                      "categories, top-1000 kept: the top-k tier sweeps the ~1300 "
                                                      "best-scored boxes first (exact); latency-bound, not HBM-bound"}})
        # the detector's inference path (configs[3]): decode -> score threshold -> per-class NMS -> top 300, fused
        # (det_dense_detect: streaming select kernel + one-CTA-per-image NMS kernel; no dense output)
        thr, max_det, cap = 0.1, 300, 2048
        res = {}
        for gate in (False, True):
            ws = det.DenseDetectWorkspace(n, cap, dev)  # counters zeroed once, left zero by every call
            r = dh.detect_thresholded(heads[0], thr, 0.5, max_det=max_det, cand_cap=cap, gate=gate, check=True, workspace=ws)
            fns = [(lambda hs=hs_: dh.detect_thresholded(hs, thr, 0.5, max_det=max_det, cand_cap=cap, gate=gate,
                                                         check=False, out=r, workspace=ws)) for hs_ in heads]
            ms = time_graph(fns, 3 if quick else 10)
            res[gate] = (ms, float(r["count"].float().mean()))
        k_mean = res[False][1]
        det_bytes = n * (4 * R * (5 + C80) + (36 * k_mean + 4))
        entry["detect_thresholded"] = {
            "workload": f"decode + score>{thr} + per-class NMS (IoU 0.5) + top {max_det}, batch {n}: det_dense_detect "
                        "(2 launches per batch)",
            "kept_per_image": k_mean,
            "ms_streaming": res[False][0], "images_per_s_streaming": n / res[False][0] * 1e3,
            "roofline": {"bound": "hbm", "achieved": det_bytes / res[False][0] / 1e6, "peak": peak, "unit": "GB/s",
                         "frac": det_bytes / res[False][0] / 1e6 / peak, "algorithmic_bytes": det_bytes,
                         "formula": "N * (4*R*(5+C) read + (36*K + 4) written), R=25200, C=80, K = kept per image; "
                                    "both kernels (select + NMS) inside the timed region, gate off (whole head streamed)",
                         "traffic": load_traffic(f"dense_select_kernel_n{n}"),
                         "note": "the peak is the driver's COPY figure (half reads, half writes); this path is 99.9 % "
                                 "reads, so a fraction slightly above 1 is a read-only stream without write turnarounds, "
                                 "not an error"},
            "ms_gated": res[True][0], "images_per_s_gated": n / res[True][0] * 1e3,
            "gated_note": "objectness plane first; class/box planes of 4-position groups that cannot pass are never "
                          "read (exact: score <= sigmoid(obj)); fewer bytes than the formula, so no roofline claim"}
        out[f"dense_head_25200x80_b{n}"] = entry
        del heads, outbuf
        torch.cuda.empty_cache()
    return out


def extras_iou(det, dev, peak, quick):
    """pairwise IoU (reference structures/boxes.py:193) at M = 20 000 x 20 000 (BASELINE configs[4]: the largest materialised
    matrix of the sweep, 1.6 GB written)."""
    m = 4000 if quick else 20000
    g = torch.Generator().manual_seed(4)
    xy = torch.rand(2 * m, 2, generator=g) * 0.8 * 1024
    wh = torch.rand(2 * m, 2, generator=g) * 0.2 * 1024 + 1
    bx = torch.cat([xy, xy + wh], 1).to(dev)
    b1, b2 = det.Boxes(bx[:m]), det.Boxes(bx[m:])
    det.pairwise_iou(b1, b2)
    ms = time_region(lambda: det.pairwise_iou(b1, b2), 3 if quick else 10, warm=2)  # allocates its (M,M) output per call
    nbytes = 16 * (m + m) + 4 * m * m
    return {"pairwise_iou": {"workload": f"pairwise_iou {m} x {m} boxes, matrix materialised", "ms": ms,
                             "roofline": {"bound": "hbm", "achieved": nbytes / ms / 1e6, "peak": peak, "unit": "GB/s",
                                          "frac": nbytes / ms / 1e6 / peak, "algorithmic_bytes": nbytes,
                                          "formula": "16(N+M) + 4NM (SURVEY 8d)",
                                          "traffic": load_traffic(f"pairwise_overlap_kernel_m{m}")},
                             "pair_ious_per_s": m * m / ms * 1e3}}


def extras_proposals(det, dev, peak, quick):
    """RPN inference path of the reference (rpn.py:299-328 -> models/utils.py:9-109): NCHW head -> decode of all pyramid levels
    (det_rpn_decode) -> find_top_rpn_proposals for the whole batch (det_rpn_proposals), reference defaults in eval mode
    (pre_nms_topk 12000 per level, post_nms_topk 2000, NMS 0.7), FPN-18 at 448x448: R = 50127 anchors."""
    n = 16 if quick else 64
    strides = [4, 8, 16, 32, 64]
    rpn = det.RegionProposalNetwork(strides)
    g = torch.Generator(device=dev).manual_seed(6)
    obj = [torch.randn(n, 3, 448 // s, 448 // s, device=dev, generator=g) for s in strides]
    dlt = [torch.randn(n, 12, 448 // s, 448 // s, device=dev, generator=g) * 0.4 for s in strides]
    sizes = torch.tensor([[448, 448]] * n, dtype=torch.int32, device=dev)

    def run():
        logits, boxes, level_sizes = rpn.decode_heads(obj, dlt)
        return det.rpn_proposals_batched(boxes, logits, level_sizes, sizes, 0.7, 12000, 2000, 0.0)

    def run_train_defaults():  # the reference's training-mode defaults: pre 2000 per level, post 1000
        logits, boxes, level_sizes = rpn.decode_heads(obj, dlt)
        return det.rpn_proposals_batched(boxes, logits, level_sizes, sizes, 0.7, 2000, 1000, 0.0)

    out = run()
    ms = time_graph([run], 3 if quick else 10)
    run_train_defaults()
    ms_t = time_graph([run_train_defaults], 3 if quick else 10)
    ms_dec = time_graph([lambda: rpn.decode_heads(obj, dlt)], 5 if quick else 20)
    # the other regime: clustered logits (a coarse random field: one value per 8x8 block of positions, shared by the cell
    # anchors, plus a little noise) and small deltas -- neighbouring anchors score alike and overlap, as the outputs of a
    # trained RPN around objects do, so the NMS suppresses most of the top candidates, the tier cut falls short and the
    # full segments are swept.  Random logits (above) are the low-suppression end.
    obj_c, dlt_c = [], []
    for o, d_, s_ in zip(obj, dlt, strides):
        hh = 448 // s_
        coarse = torch.randn(n, 1, (hh + 7) // 8, (hh + 7) // 8, device=dev, generator=g)
        field = torch.nn.functional.interpolate(coarse, size=(hh, hh), mode="nearest")
        obj_c.append((field.expand(n, 3, hh, hh) * 2.0 + 0.05 * torch.randn(n, 3, hh, hh, device=dev, generator=g)).contiguous())
        dlt_c.append(d_ * 0.1)

    def run_clustered():
        logits, boxes, level_sizes = rpn.decode_heads(obj_c, dlt_c)
        return det.rpn_proposals_batched(boxes, logits, level_sizes, sizes, 0.7, 2000, 1000, 0.0)

    out_c = run_clustered()
    ms_c = time_graph([run_clustered], 3 if quick else 10)
    R = 50127
    dec_bytes = n * 40 * R
    return {"rpn_proposals_r50127": {
        "workload": f"RPN head (NCHW, 5 levels, R=50127) -> decode + find_top_rpn_proposals (pre 12000/level, post 2000, NMS 0.7), "
                    f"batch {n}, replayed from a CUDA graph",
        "ms": ms, "images_per_s": n / ms * 1e3, "kept_per_image": float(out[2].float().mean()),
        "ms_pre2000_post1000": ms_t, "images_per_s_pre2000_post1000": n / ms_t * 1e3,
        "ms_pre2000_post1000_clustered": ms_c, "images_per_s_pre2000_post1000_clustered": n / ms_c * 1e3,
        "kept_per_image_clustered": float(out_c[2].float().mean()),
        "clustered_note": "logits = coarse random field (8x8 blocks) + noise, small deltas: the high-suppression regime of a "
                          "trained RPN; the random-logit lines are the low-suppression end",
        "ms_decode": ms_dec,
        "decode_roofline": {"bound": "hbm", "achieved": dec_bytes / ms_dec / 1e6, "peak": peak, "unit": "GB/s",
                            "frac": dec_bytes / ms_dec / 1e6 / peak, "algorithmic_bytes": dec_bytes,
                            "formula": "N * 40R: 4R logits + 16R deltas read, 4R logits + 16R boxes written, anchors "
                                       "synthesised (SURVEY 8d counts 36R without the logit copy)"}}}


def _r(x, nd=4):
    return None if x is None else round(float(x), nd)


def train_summary(extras, world):
    """Compact training-side numbers (BASELINE metric "loss fwd/bwd at 1/2/4/8 B200"): printed LAST in the JSON line and
    mirrored into e2e so that the driver's retained fields carry them.  Weak scaling: 1024 images per GPU per step."""
    out = {"scaling": "weak", "images_per_gpu_per_step": 1024, "n_gpus": world}
    r = extras.get("train_rpn_r50127")
    if r:
        out["rpn_r50127"] = {"ms_step": _r(r.get("ms_step")), "ms_step_dense_labels": _r(r.get("ms_step_dense_labels")),
                             "ms_assign_sampled": _r(r.get("ms_assign_sampled")), "ms_assign": _r(r.get("ms_assign")),
                             "ms_loss": _r(r.get("ms_loss_fwd_bwd")), "ms_graph": _r(r.get("ms_step_graph_no_collective")),
                             "ms_graph_peer": _r(r.get("ms_step_graph_peer_allreduce")),
                             "img_s_total": _r(r.get("images_per_s_total"), 0),
                             "img_s_total_graph_peer": _r(r.get("images_per_s_total_graph_peer_allreduce"), 0),
                             "assign_frac": _r(r["assign_roofline"]["frac"], 3), "loss_frac": _r(r["loss_roofline"]["frac"], 3),
                             "batch": r.get("batch")}
    g = extras.get("train_grid_b1024")
    if g:
        out["grid7x7x30"] = {"ms_nccl": _r(g.get("ms_per_step")), "ms_peer": _r(g.get("ms_per_step_peer_exchange")),
                            "ms_fused_peer": _r(g.get("ms_per_step_fused_peer_exchange")),
                            "ms_graph_fused_peer": _r(g.get("ms_per_step_graph_fused_peer_exchange"))}
        best = min(v for v in out["grid7x7x30"].values() if v)
        out["grid7x7x30"]["img_s_total_best"] = round(world * 1024 / best * 1e3)
    return out


def others_summary(extras):
    """roofline fractions of the HBM-bound kernels of the path (each measured in extras with its byte formula)."""
    out = {}
    for n in (32, 256):
        e = extras.get(f"dense_head_25200x80_b{n}")
        if e:
            out[f"dense_detect_b{n}"] = _r(e["detect_thresholded"]["roofline"]["frac"], 3)
            out[f"dense_decode_b{n}"] = _r(e["decode_roofline"]["frac"], 3)
    r = extras.get("train_rpn_r50127")
    if r:
        out["rpn_loss_fwd_bwd"] = _r(r["loss_roofline"]["frac"], 3)
        out["rpn_assign"] = _r(r["assign_roofline"]["frac"], 3)
    p = extras.get("pairwise_iou")
    if p:
        out["pairwise_iou_20k"] = _r(p["roofline"]["frac"], 3)
    q = extras.get("rpn_proposals_r50127")
    if q:
        out["rpn_decode_b64"] = _r(q["decode_roofline"]["frac"], 3)
    return out


# ----------------------------------------------------------------------------------------------------------------
def bench_config():
    """The workload description both arms print (identical keys and values: the driver compares them)."""
    return {"workload": WORKLOAD, "batch_per_gpu_per_step": BATCH, "grid": S, "boxes": B, "classes": C,
            "image": list(IMG), "score_thresh": SCORE_THR, "iou_thresh": IOU_THR, "max_det": MAX_DET}


def all_ranks_max_int(v, world, dev):
    if world > 1:
        t = torch.tensor([int(v)], dtype=torch.int64, device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return int(t.item())
    return int(v)


def peer_equals_nccl(det, dev, world):
    """PeerSums (NVLink peer-memory SUM, csrc/peer.cu) against NCCL's all_reduce of the same vectors, on the real ranks:
    three exchanged steps + the flush.  Equal means bitwise or within 1e-6 relative (the summation order may differ)."""
    ps = det.dist.PeerSums(dev)
    rank = ps.rank
    ok = True
    vecs = [torch.arange(8, dtype=torch.float32, device=dev) * (0.37 + step) + (rank + 1) * 1.25 + step for step in range(4)]
    want = []
    for v in vecs:
        w = v.clone()
        det.dist.allreduce_sums_(w)
        want.append(w)
    got = []
    for v in vecs:
        prev = ps.exchange(v)
        if prev is not None:
            got.append(prev.clone())
    got.append(ps.flush().clone())
    ps.check()
    # the graph-safe form (step stamp on the device): `out` holds the previous step's world sum from the second step on
    psg = det.dist.PeerSums(dev, graph_safe=True)
    for k, v in enumerate(vecs):
        prev = psg.exchange(v)
        if k > 0:
            got.append(prev.clone())
            want.append(want[k - 1])
    got.append(psg.flush().clone())
    want.append(want[len(vecs) - 1])
    psg.check()
    for g_, w_ in zip(got, want):
        ok = ok and bool(torch.equal(g_, w_) or torch.allclose(g_, w_, rtol=1e-6, atol=0.0))
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
    if world > 1:
        torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MIN)
    return bool(int(flag.item()))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--quick", action="store_true", help="smaller extras / CPU sample (CI)")
    ap.add_argument("--lanes", type=int, default=1,
                    help="streams the independent steps rotate over inside the graph (1 = the caller's single stream)")
    ap.add_argument("--min-ms", type=float, default=60.0, help="lower bound of every timed region (whole graph replays)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import det_b200 as det
    det._native.lib()  # fail loudly if the CUDA library is missing
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (there is no CPU fallback)")
    rank, world, local = det.dist.init_from_env()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    peak, peak_src = load_peaks()

    yh = det.YoloGridHead(S, B, C, IMG)
    img_bytes = S * S * (B * 5 + C) * 4
    pool = L2_BYTES // (BATCH * img_bytes) + 8  # rotating pool of input batches larger than L2
    heads = make_heads(pool, 1 + rank).to(dev)
    n_out = 4
    outs = [yh.detect(heads[i], SCORE_THR, IOU_THR, max_det=MAX_DET) for i in range(n_out)]
    torch.cuda.synchronize()

    # ONE CUDA graph = one pass over the pool: `pool` launches of the fused kernel, one per step, each on its own input
    # batch (the pool is larger than L2), outputs rotating over 4 buffers.  The timed region is a whole number of
    # replays of that graph, however few --steps asks for: the launch path is the same for --steps 20 and --steps 20000.
    n_lanes = max(1, min(4, args.lanes))
    lanes = [torch.cuda.Stream() for _ in range(n_lanes - 1)]
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for i in range(3):
            yh.detect(heads[i], SCORE_THR, IOU_THR, max_det=MAX_DET, out=outs[i % n_out])
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        cap = torch.cuda.current_stream()
        for st_ in lanes:
            st_.wait_stream(cap)
        for i in range(pool):
            lane = i % n_lanes
            if lane == 0:
                yh.detect(heads[i], SCORE_THR, IOU_THR, max_det=MAX_DET, out=outs[i % n_out])
            else:
                with torch.cuda.stream(lanes[lane - 1]):
                    yh.detect(heads[i], SCORE_THR, IOU_THR, max_det=MAX_DET, out=outs[i % n_out])
        for st_ in lanes:
            cap.wait_stream(st_)

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(max(2, -(-args.warmup // pool))):
        graph.replay()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(2):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    est_replay_ms = max(e0.elapsed_time(e1) / 2, 1e-3)
    replays = max(-(-args.steps // pool), int(math.ceil(args.min_ms / est_replay_ms)))
    replays = all_ranks_max_int(replays, world, dev)
    timed_steps = replays * pool

    sampler = ClockSampler(local)
    barrier(world)
    sampler.start()
    e0.record()
    for _ in range(replays):
        graph.replay()
    e1.record()
    barrier(world)
    clocks = sampler.stop()
    ms_total = max_over_ranks(e0.elapsed_time(e1), world, dev)
    ms_step = ms_total / timed_steps
    value = world * BATCH * timed_steps / (ms_total * 1e-3)

    # algorithmic bytes of one launch: logits read + kept detections written (fused: no decode hand-off)
    kept = sum(int(o["count"].sum()) for o in outs) / len(outs)
    algo_bytes = BATCH * img_bytes + kept * (8 + 16 + 4) + BATCH * 4
    achieved = algo_bytes / (ms_step * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": load_traffic("yolo_fast_kernel"), "kernel": "yolo_fast_kernel<640,2,7,2,20>",
                "algorithmic_bytes_per_launch": algo_bytes, "peak_source": peak_src,
                "note": "3.7 MB per launch = 0.6 us at the HBM peak: a 256-image launch is bound by launch latency and "
                        "instruction issue of the per-image NMS, not by HBM; the HBM-bound kernels of this path "
                        "(dense-head decode+NMS, loss fwd+bwd, assignment, pairwise IoU) are in `others`"}

    # end to end through the public API (det.YoloHostPipeline) with HOST buffers: every step uploads its batch from
    # pinned host memory, runs the fused kernel and downloads the detections; H2D / kernel / D2H of consecutive
    # steps overlap on the pipeline's per-slot streams.  The timed region ends when the last result is on the host.
    depth = int(os.environ.get("DET_E2E_DEPTH", "8"))  # pipeline slots (streams / graphs in flight)
    host_src = make_heads(depth, 100 + rank)

    def make_pipe(index_dtype):
        p_ = det.YoloHostPipeline(yh, BATCH, SCORE_THR, IOU_THR, MAX_DET, depth=depth, device=dev, index_dtype=index_dtype)
        for s_ in range(depth):
            p_.input(s_).copy_(host_src[s_])
        return p_

    # detection indices travel as int32 (same values as the int64 `flat` of detect(); 4 of 28 bytes per detection less
    # on the PCIe download, which is what bounds this number); the int64 wire format is timed once below for reference
    pipe = make_pipe(torch.int32)

    def e2e_run(k):
        seen = 0
        for i in range(k):
            slot = i % depth
            if i >= depth:
                seen += int(pipe.wait(slot)["count"][0])  # the result of step i - depth is read on the host
            pipe.launch(slot)
        for i in range(max(0, k - depth), k):
            seen += int(pipe.wait(i % depth)["count"][0])
        return seen

    # warm-up: the host side of a fresh box (pinned pages first touched by the copy engines, CPU clocks, the launch path)
    # needs a few hundred ms to settle -- 96 steps were not enough on one of the boxes (85 -> 68 -> 52 us per step over the
    # three segments); run until 0.3 s have passed
    t_w = time.perf_counter()
    while time.perf_counter() - t_w < 0.3:
        e2e_run(8 * depth)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e2e_run(8 * depth)
    est_step_ms = max((time.perf_counter() - t0) * 1e3 / (8 * depth), 1e-3)
    e2e_steps = all_ranks_max_int(max(args.steps, int(math.ceil(args.min_ms / est_step_ms))), world, dev)
    # five back-to-back segments, the median is reported (host-side jitter on a shared box shows up as slow segments)
    n_seg = 5
    seg_ms, seg_wall = [], []
    for _ in range(n_seg):
        barrier(world)
        t0 = time.perf_counter()
        e0.record()
        e2e_run(e2e_steps)
        e1.record()
        barrier(world)
        seg_wall.append((time.perf_counter() - t0) * 1e3)
        seg_ms.append(max_over_ranks(max(e0.elapsed_time(e1), 0.0), world, dev))
    order = sorted(range(n_seg), key=lambda i: seg_ms[i])
    e2e_ms, e2e_wall_ms = seg_ms[order[n_seg // 2]], seg_wall[order[n_seg // 2]]
    e2e = {"value": world * BATCH * e2e_steps / (e2e_ms * 1e-3), "unit": "images/s",
           "h2d_bytes_per_step": pipe.h2d_bytes, "d2h_bytes_per_step": pipe.d2h_bytes, "steps": e2e_steps,
           "ms_per_step": e2e_ms / e2e_steps, "wall_ms_per_step": e2e_wall_ms / e2e_steps,
           "segments_ms_per_step": [m / e2e_steps for m in seg_ms],
           "api": f"det_b200.YoloHostPipeline(depth={depth}, index_dtype=int32): pinned host in -> H2D -> "
                  "det_yolo_decode_nms_i32 -> D2H -> pinned host out, one CUDA graph per slot, slots on separate streams"}
    pipe = make_pipe(torch.int64)
    e2e_run(8 * depth)
    m64 = []
    for _ in range(3):
        barrier(world)
        e0.record()
        e2e_run(e2e_steps)
        e1.record()
        barrier(world)
        m64.append(max_over_ranks(max(e0.elapsed_time(e1), 0.0), world, dev))
    ms64 = sorted(m64)[1]
    e2e["int64_indices"] = {"value": world * BATCH * e2e_steps / (ms64 * 1e-3), "d2h_bytes_per_step": pipe.d2h_bytes,
                            "ms_per_step": ms64 / e2e_steps}
    del pipe

    extras, train, others = {}, {}, {}
    if not args.no_extras:
        try:
            train["peer_equals_nccl"] = peer_equals_nccl(det, dev, world)
        except Exception as e:  # noqa: BLE001
            train["peer_equals_nccl"] = False
            train["peer_error"] = f"{type(e).__name__}: {e}"[:160]
        try:
            extras.update(extras_train(det, dev, world, peak, args.quick))
            if world == 1 or rank == 0:
                extras.update(extras_dense(det, dev, peak, args.quick))
                extras.update(extras_iou(det, dev, peak, args.quick))
                extras.update(extras_proposals(det, dev, peak, args.quick))
        except Exception as e:  # noqa: BLE001
            extras["error"] = f"{type(e).__name__}: {e}"
        barrier(world)
        train.update(train_summary(extras, world))
        others = others_summary(extras)
        e2e["peer_equals_nccl"] = train["peer_equals_nccl"]
        e2e["train"] = train
    roofline["others"] = others

    cpu_baseline = None
    if rank == 0:
        cpu_pool = CpuPool()
        sample = make_heads(1, 1)[0]
        cpu_pool.step(sample)
        t0 = time.perf_counter()
        cpu_pool.step(sample)
        t_batch = time.perf_counter() - t0
        reps = max(1, min(600, int((3.0 if args.quick else 12.0) / max(t_batch, 1e-3))))
        t0 = time.perf_counter()
        for _ in range(reps):
            cpu_pool.step(sample)
        dt = time.perf_counter() - t0
        cpu_pool.close()
        cpu_baseline = {"value": reps * BATCH / dt, "unit": "images/s", "cores": cpu_pool.cores, "kind": "port",
                        "sample": f"{reps} passes over one {BATCH}-image batch of the same synthetic distribution "
                                  f"({dt:.1f} s), each split over {cpu_pool.cores} single-threaded worker processes: oracle "
                                  "torch-CPU decode + C greedy NMS, per-image loop"}
        lane_txt = ("all on the caller's single stream" if n_lanes == 1 else
                    f"consecutive (independent) batches rotate over {n_lanes} streams inside the graph")
        line = {
            "metric": "images/sec decode+NMS", "value": value, "unit": "images/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(),
            "timing": {"timed_steps": timed_steps, "timed_ms": ms_total, "graph_replays": replays,
                       "steps_per_replay": pool, "lanes": n_lanes,
                       "launch": f"CUDA graph replay: one graph = {pool} launches (one kernel per step), {lane_txt}; the "
                                 f"timed region is {replays} whole replays (>= --steps and >= {args.min_ms:.0f} ms)",
                       "l2": f"inputs rotate over a pool of {pool} batches = {pool * BATCH * img_bytes / 2**20:.0f} MiB "
                             "> 126 MiB L2"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": timed_steps, "roofline": roofline,
            "cpu_baseline": cpu_baseline, "extras": extras, "train": train,
        }
        print(json.dumps(line))
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
