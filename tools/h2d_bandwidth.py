import torch, time
x = torch.empty(1<<30, dtype=torch.uint8).pin_memory()
y = torch.empty(1<<30, dtype=torch.uint8, device="cuda")
for _ in range(2): y.copy_(x, non_blocking=True)
torch.cuda.synchronize(); t=time.perf_counter()
for _ in range(4): y.copy_(x, non_blocking=True)
torch.cuda.synchronize(); print("H2D pinned GB/s", 4*(1<<30)/1e9/(time.perf_counter()-t))
