"""H2D bandwidth from torch-pinned memory vs write-combined pinned memory (cudaHostAllocWriteCombined): python tools/wc_probe.py"""
import ctypes, time
import numpy as np, torch
from cuda import cudart
dev = torch.device("cuda", 0)
n = 328 << 20
dst = torch.empty(n, dtype=torch.uint8, device=dev)
def bw(src, label, with_d2h=False):
    s = torch.cuda.Stream(dev); s2 = torch.cuda.Stream(dev)
    out_h = torch.empty(134 << 20, dtype=torch.uint8).pin_memory(); out_d = torch.empty(134 << 20, dtype=torch.uint8, device=dev)
    for _ in range(2):
        with torch.cuda.stream(s): dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(s):
        e0.record(s)
        for _ in range(8): dst.copy_(src, non_blocking=True)
        e1.record(s)
    if with_d2h:
        with torch.cuda.stream(s2):
            for _ in range(16): out_h.copy_(out_d, non_blocking=True)
    torch.cuda.synchronize()
    print(label, "with concurrent D2H" if with_d2h else "alone", round(8 * n / (e0.elapsed_time(e1) * 1e-3) / 1e9, 2), "GB/s", "pinned" if src.is_pinned() else "NOT pinned")
a = torch.empty(n, dtype=torch.uint8).pin_memory(); a.fill_(1)
err, ptr = cudart.cudaHostAlloc(n, cudart.cudaHostAllocWriteCombined)
assert err == cudart.cudaError_t.cudaSuccess, err
buf = (ctypes.c_uint8 * n).from_address(ptr)
w = torch.frombuffer(buf, dtype=torch.uint8); w.fill_(1)
for d2h in (False, True):
    bw(a, "torch pin_memory   ", d2h)
    bw(w, "write-combined host", d2h)
