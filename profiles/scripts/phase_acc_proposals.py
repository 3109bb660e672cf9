"""Accumulated phase cycles (block 0, -DDET_DEBUG_PHASES build) of cta_segment_nms inside large_cta_segments_kernel on the RPN
proposal path: 0 = chunk load, 1 = (a) kept-list tests, 2 = (b) bit rows, 3 = (c) resolve, 4.. = see the kernel."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
PKG = os.path.join(ROOT, "object-detection-pytorch-rust_b200")
sys.path.insert(0, PKG)
import importlib.util
spec = importlib.util.spec_from_file_location("b", os.path.join(PKG, "build.py"))
b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
dbg = b.build(debug_phases=True)
import det_b200._native as N
N._LIB_PATH = dbg
import torch
import det_b200 as det
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
PRE = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
POST = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
dev = torch.device("cuda", 0)
strides = [4, 8, 16, 32, 64]
rpn = det.RegionProposalNetwork(strides)
g = torch.Generator().manual_seed(6)
obj = [torch.randn(n, 3, 448 // s, 448 // s, generator=g).to(dev) for s in strides]
dlt = [(torch.randn(n, 12, 448 // s, 448 // s, generator=g) * 0.4).to(dev) for s in strides]
sizes = torch.tensor([[448, 448]] * n, dtype=torch.int32, device=dev)
def run():
    logits, boxes, level_sizes = rpn.decode_heads(obj, dlt)
    return det.rpn_proposals_batched(boxes, logits, level_sizes, sizes, 0.7, PRE, POST, 0.0)
for _ in range(3): run()
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 16)()
fn = N.lib().det_debug_read_acc_proposals
fn.argtypes = [ctypes.c_void_p, ctypes.c_int]
assert fn(buf, 1) == 0
run(); torch.cuda.synchronize()
assert fn(buf, 1) == 0
print(f"n={n} pre={PRE} post={POST}: accumulated cycles of block 0:", [int(x) for x in buf[:8]])
blk = (ctypes.c_longlong * (64 * 16))()
assert N.lib().det_debug_read_phase_blocks_proposals(blk) == 0
print("first 64 CTAs of the LAST large_cta_segments_kernel launch (cycles, segments, boxes):")
print([(blk[b * 16 + 1] - blk[b * 16 + 0], blk[b * 16 + 2], blk[b * 16 + 3]) for b in range(64)])
