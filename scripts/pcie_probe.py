"""Page-locked copy bandwidth of the box (what bounds the end-to-end rate of trace-heavy configs): python scripts/pcie_probe.py"""
import json, time, torch
out = {}
for mb in (8, 32, 128):
    n = mb << 20
    d = torch.empty(n, dtype=torch.uint8, device="cuda"); h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    for name, src, dst in (("d2h", d, h), ("h2d", h, d)):
        for _ in range(3): dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): dst.copy_(src, non_blocking=True)
        e1.record(); torch.cuda.synchronize()
        out["%s_%dMB_GBs" % (name, mb)] = round(10 * n / (e0.elapsed_time(e1) * 1e-3) / 1e9, 1)
# both directions at once on two streams
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
n = 32 << 20
d1 = torch.empty(n, dtype=torch.uint8, device="cuda"); h1 = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d2 = torch.empty(n, dtype=torch.uint8, device="cuda"); h2 = torch.empty(n, dtype=torch.uint8, pin_memory=True)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10):
    with torch.cuda.stream(s1): h1.copy_(d1, non_blocking=True)
    with torch.cuda.stream(s2): d2.copy_(h2, non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
out["bidirectional_32MB_each_GBs"] = round(10 * n / dt / 1e9, 1)
print(json.dumps(out))
