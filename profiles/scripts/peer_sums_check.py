"""torchrun driver: det_b200.dist.PeerSums (NVLink peer-memory exchange, csrc/peer.cu) vs NCCL all_reduce of the same
8-float vectors -- equality of the sums on every rank, and the cost of each per training step of the grid head."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "object-detection-pytorch-rust_b200"))
import torch
import torch.distributed as dist
import det_b200 as det
from bench import synth_gt, time_region

rank, world, local = det.dist.init_from_env()
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
ps = det.dist.PeerSums(dev)
g = torch.Generator(device=dev).manual_seed(100 + rank)
ok = True
vs = []
for step in range(1, 41):
    v = torch.randn(8, device=dev, generator=g)
    vs.append(v)
    ps.publish(v)
    if step >= 2:
        got = ps.collect().clone()
        want = vs[step - 2].clone()
        dist.all_reduce(want)
        ok &= bool(torch.allclose(got, want, rtol=1e-6, atol=1e-6))
got = ps.collect().clone(); want = vs[-1].clone(); dist.all_reduce(want)
ok &= bool(torch.allclose(got, want, rtol=1e-6, atol=1e-6))
ps.check()
px = det.dist.PeerSums(dev)
prev = None
for step in range(30):
    v = torch.randn(8, device=dev, generator=g)
    got = px.exchange(v)
    if prev is not None:
        want = prev.clone(); dist.all_reduce(want)
        ok &= bool(torch.allclose(got, want, rtol=1e-6, atol=1e-6))
    prev = v
want = prev.clone(); dist.all_reduce(want)
ok &= bool(torch.allclose(px.flush(), want, rtol=1e-6, atol=1e-6))
px.check()
# identical bits on every rank
gathered = [torch.empty_like(got) for _ in range(world)]
dist.all_gather(gathered, got)
ok &= all(torch.equal(gathered[0], t) for t in gathered)

yh = det.YoloGridHead(7, 2, 20, (448, 448))
tr = det.YoloGridTrainer(yh)
n = 1024
gtb, gtc, off, tot = synth_gt(n, 2, dev)
heads = [torch.randn(n, 7, 7, 30, device=dev) for _ in range(24)]
st = {"i": 0}

def step_nccl():
    h = heads[st["i"] % 24]; st["i"] += 1
    res = tr.loss(h, tr.assign_packed(gtb, off, n), gtc, with_grads=True)
    prev = st.get("pending")
    st["pending"] = det.dist.allreduce_sums_async(res["sums"])
    if prev is not None:
        prev.wait()

def step_peer():
    h = heads[st["i"] % 24]; st["i"] += 1
    res = tr.loss(h, tr.assign_packed(gtb, off, n), gtc, with_grads=True)
    ps2.exchange(res["sums"])

ps3 = det.dist.PeerSums(dev)

def step_fused():
    h = heads[st["i"] % 24]; st["i"] += 1
    tr.loss(h, tr.assign_packed(gtb, off, n), gtc, with_grads=True, peer=ps3)

def step_none():
    h = heads[st["i"] % 24]; st["i"] += 1
    tr.loss(h, tr.assign_packed(gtb, off, n), gtc, with_grads=True)

ms_nccl = time_region(step_nccl, 300)
if st.get("pending") is not None: st["pending"].wait()
dist.barrier()
ps2 = det.dist.PeerSums(dev)
ms_peer = time_region(step_peer, 300)
ps2.flush(); ps2.check()
dist.barrier()
# fused: numerically the same as an NCCL all-reduce of the previous step's sums
prev_local = None
for i in range(5):
    res = tr.loss(heads[i], tr.assign_packed(gtb, off, n), gtc, with_grads=True, peer=ps3)
    if prev_local is not None:
        want = prev_local.clone(); dist.all_reduce(want)
        ok &= bool(torch.allclose(res["world_sums_prev"], want, rtol=1e-5, atol=1e-5))
    prev_local = res["sums"].clone()
dist.barrier()
ms_fused = time_region(step_fused, 300)
ps3.flush(); ps3.check()
dist.barrier()
ms_none = time_region(step_none, 300)
t = torch.tensor([ms_nccl, ms_peer, ms_none, ms_fused], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"world": world, "sums_equal_nccl_and_bitwise_equal_across_ranks": ok,
                      "ms_step_nccl_allreduce": round(float(t[0]), 5), "ms_step_peer_exchange": round(float(t[1]), 5),
                      "ms_step_fused_in_loss_kernel": round(float(t[3]), 5), "ms_step_no_collective": round(float(t[2]), 5)}), flush=True)
dist.destroy_process_group()
