"""pinned host -> device copy bandwidth of the box (what bounds bench.py's e2e leg: every compressed byte crosses PCIe once)"""
import time, torch
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for _ in range(2): d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(8): d.copy_(h, non_blocking=True)
b.record(); torch.cuda.synchronize()
print("H2D pinned: %.1f GB/s" % (8 * n / (a.elapsed_time(b) * 1e-3) / 1e9))
a.record()
for _ in range(8): h.copy_(d, non_blocking=True)
b.record(); torch.cuda.synchronize()
print("D2H pinned: %.1f GB/s" % (8 * n / (a.elapsed_time(b) * 1e-3) / 1e9))
