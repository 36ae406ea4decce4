import torch, time
x = torch.empty(307200*3, dtype=torch.float32).pin_memory()
y = torch.empty(307200*3, dtype=torch.float32).pin_memory()
d = torch.empty_like(x, device='cuda'); e = torch.empty_like(x, device='cuda')
s = torch.cuda.Stream()
for size in (307200*3, 307200*3*8):
    a = torch.empty(size, dtype=torch.float32).pin_memory(); b = torch.empty(size, dtype=torch.float32, device='cuda')
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(50): b.copy_(a, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"H2D {size*4/1e6:.1f} MB: {size*4*50/dt/1e9:.1f} GB/s, {dt/50*1e3:.3f} ms each")
    t0 = time.perf_counter()
    for _ in range(50): a.copy_(b, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"D2H {size*4/1e6:.1f} MB: {size*4*50/dt/1e9:.1f} GB/s")
