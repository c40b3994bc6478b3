import torch, time
for mb in (0.5, 2.3, 8):
    n = int(mb * 1e6)
    h = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device="cuda")
    for name, f in (("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(d, non_blocking=True))):
        for _ in range(3): f()
        torch.cuda.synchronize(); t = time.perf_counter()
        for _ in range(20): f()
        torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 20
        print(f"{name} {mb} MB: {dt*1e6:.1f} us = {n/dt/1e9:.1f} GB/s")
