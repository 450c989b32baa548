import torch, time
for mb in (1, 64):
    n = mb << 20
    h = torch.empty(n, dtype=torch.uint8, pin_memory=True); d = torch.empty(n, dtype=torch.uint8, device="cuda")
    for name, f in (("H2D", lambda: d.copy_(h, non_blocking=True)), ("D2H", lambda: h.copy_(d, non_blocking=True))):
        f(); torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(50): f()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t) / 50
        print(f"{name} {mb} MiB: {n/dt/1e9:.1f} GB/s ({dt*1e6:.0f} us)")
