import torch, time
dev = torch.device("cuda", 0)
for mb in (0.39, 0.79, 1.57, 3.15, 12.6, 100):
    nb = int(mb * 1e6) // 4
    h = torch.rand(nb).pin_memory(); d = torch.empty(nb, device=dev)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(5): d.copy_(h, non_blocking=True)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(2000 if mb < 4 else 50): d.copy_(h, non_blocking=True)
        e1.record()
    e1.synchronize()
    us = e0.elapsed_time(e1) / (2000 if mb < 4 else 50) * 1e3
    print(f"H2D {mb:7.2f} MB: {us:8.1f} us  {nb * 4 / us / 1e3:6.1f} GB/s")
