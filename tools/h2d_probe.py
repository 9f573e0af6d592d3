import torch, time
x = torch.empty(369_000_000//4, dtype=torch.float32).pin_memory()
d = torch.empty_like(x, device='cuda')
for n in (1, 4, 16):
    chunks_h = x.chunk(n); chunks_d = d.chunk(n)
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for h, dd in zip(chunks_h, chunks_d): dd.copy_(h, non_blocking=True)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(n, "chunks: %.2f ms  %.1f GB/s" % (dt*1e3, x.numel()*4/dt/1e9))
