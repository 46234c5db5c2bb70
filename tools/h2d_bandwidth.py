import torch, time
n = 4 << 30
h = torch.empty(n, dtype=torch.uint8, pin_memory=True); h.fill_(7)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for chunk in (n, 1 << 30, 256 << 20):
    best = 0
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for a in range(0, n, chunk):
            d[a:a + chunk].copy_(h[a:a + chunk], non_blocking=True)
        e1.record(); torch.cuda.synchronize()
        best = max(best, n / (e0.elapsed_time(e1) / 1e3) / 1e9)
    print("H2D pinned, chunk %d MB: %.2f GB/s" % (chunk >> 20, best))
