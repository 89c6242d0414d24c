import torch, time
x = torch.empty(1<<30, dtype=torch.uint8, device='cuda'); h = torch.empty(1<<30, dtype=torch.uint8).pin_memory()
for _ in range(2):
    torch.cuda.synchronize(); t=time.perf_counter(); h.copy_(x, non_blocking=True); torch.cuda.synchronize(); d2h=time.perf_counter()-t
    t=time.perf_counter(); x.copy_(h, non_blocking=True); torch.cuda.synchronize(); h2d=time.perf_counter()-t
print("D2H GB/s", 1.0737/d2h, "H2D GB/s", 1.0737/h2d)
