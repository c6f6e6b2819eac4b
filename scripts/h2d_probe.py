import torch, time
x = torch.empty(133*1024*1024//4, dtype=torch.float32).pin_memory()
d = torch.empty_like(x, device="cuda")
s = torch.cuda.Stream()
for n in (1, 2, 4):
    chunks = x.chunk(n); dch = d.chunk(n); streams = [torch.cuda.Stream() for _ in range(n)]
    torch.cuda.synchronize(); t = time.time()
    for _ in range(10):
        for c, dc, st in zip(chunks, dch, streams):
            with torch.cuda.stream(st): dc.copy_(c, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.time() - t) / 10
    print(n, "streams:", round(x.numel()*4/dt/1e9, 1), "GB/s", round(dt*1e3, 2), "ms")
