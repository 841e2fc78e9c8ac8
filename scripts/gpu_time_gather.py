"""Time smcdet_gather at bench size (148 tiles x 10 000 particles x D = 10) with a multinomial-like index."""
import sys
import torch
sys.path.insert(0, ".")
from smcdet_b200 import _lib as L
dev = torch.device("cuda", 0)
T, N, D = 148, 10000, 10
g = torch.Generator(device=dev).manual_seed(0)
idx = torch.sort(torch.randint(0, N, (T, N), device=dev, generator=g), dim=1)[0]
c, l, f = torch.rand(T, N, device=dev), torch.rand(T, N, D, 2, device=dev), torch.rand(T, N, D, device=dev)
co, lo, fo = torch.empty_like(c), torch.empty_like(l), torch.empty_like(f)
ts = []
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    L.check(L.lib().smcdet_gather(L.ptr(idx, torch.int64), L.ptr(c), L.ptr(l), L.ptr(f), L.ptr(co), L.ptr(lo), L.ptr(fo), None,
                                  T, N, D, L.stream_for(c)))
    e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
ms = sorted(ts)[len(ts) // 2]
assert torch.equal(lo, torch.gather(l, 1, idx.view(T, N, 1, 1).expand(-1, -1, D, 2)))
print(f"gather {ms:.4f} ms -> {2 * T * N * (12 * D + 4) / ms / 1e6:.0f} GB/s (read + write of {12 * D + 4}-byte records)")
