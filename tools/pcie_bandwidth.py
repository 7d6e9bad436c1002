import torch, time
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, (a, b) in {"h2d": (d, h), "d2h": (h, d)}.items():
    a.copy_(b, non_blocking=True); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        a.copy_(b, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    print(name, "pinned 1 GiB:", round(5 * n / (e0.elapsed_time(e1) / 1e3) / 1e9, 1), "GB/s")
