"""Write-only bandwidth with our own streaming-store kernel (debug build): zeros vs distinct values vs 5 % non-zeros."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
PKG = os.path.join(ROOT, "object-detection-pytorch-rust_b200")
sys.path.insert(0, PKG)
import importlib.util
spec = importlib.util.spec_from_file_location("b", os.path.join(PKG, "build.py"))
b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
dbg = b.build(debug_phases=True)
import torch
lib = ctypes.CDLL(dbg)
lib.det_debug_write_probe.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
n = 400_000_000
x = torch.empty(n, dtype=torch.float32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for blocks in (148 * 8, 148 * 32, 148 * 128):
    for mode, name in ((0, "zeros"), (1, "distinct non-zero"), (2, "5% non-zero")):
        for _ in range(3): lib.det_debug_write_probe(x.data_ptr(), n, mode, blocks, st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): lib.det_debug_write_probe(x.data_ptr(), n, mode, blocks, st)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"blocks {blocks:6d} {name:18s}: {ms*1e3:7.1f} us  {4*n/ms/1e6:7.0f} GB/s written", flush=True)
