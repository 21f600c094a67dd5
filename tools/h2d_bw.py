import time, torch
n = 256*480*640*3
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device='cuda')
for sz in (n, n//4, n//8):
    for _ in range(3): d[:sz].copy_(h[:sz], non_blocking=True)
    torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(10): d[:sz].copy_(h[:sz], non_blocking=True)
    torch.cuda.synchronize(); dt=(time.perf_counter()-t)/10
    print(f"H2D {sz/1e6:.0f} MB: {sz/dt/1e9:.1f} GB/s  ({dt*1e3:.2f} ms)")
