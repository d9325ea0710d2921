import time, torch
x = torch.randn(64, 5, 128, 128).pin_memory(); print("pinned", x.is_pinned())
d = torch.empty_like(x, device="cuda")
for name, src in (("pinned", x), ("pageable", torch.randn(64, 5, 128, 128))):
    for _ in range(2): d.copy_(src, non_blocking=True); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5): d.copy_(src, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    print(name, "H2D 21MB ms", dt * 1e3, "GB/s", x.numel() * 4 / dt / 1e9)
t0 = time.perf_counter()
for _ in range(5): y = x.to("cuda", non_blocking=True)
torch.cuda.synchronize(); print("to() ms", (time.perf_counter() - t0) / 5 * 1e3)
