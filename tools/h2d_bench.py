import torch, time
for mb in (16.8, 25.2):
    n = int(mb * 1e6 / 4)
    pin = torch.empty(n).pin_memory(); page = torch.empty(n); dev = torch.empty(n, device="cuda")
    for name, src in (("pinned", pin), ("pageable", page)):
        for _ in range(3): dev.copy_(src, non_blocking=True)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(10): dev.copy_(src, non_blocking=True)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
        print("%s %.1f MB: %.3f ms  %.1f GB/s" % (name, mb, dt * 1e3, mb / 1e3 / dt))
x = torch.rand(8, 3, 512, 512).pin_memory(); idx = torch.tensor([0, 1])
t0 = time.perf_counter()
for _ in range(10): y = x.index_select(1, idx)
print("cpu index_select %.3f ms" % ((time.perf_counter() - t0) / 10 * 1e3), torch.get_num_threads())
out = torch.empty(8, 2, 512, 512).pin_memory()
t0 = time.perf_counter()
for _ in range(10): torch.index_select(x, 1, idx, out=out)
print("cpu index_select into pinned %.3f ms" % ((time.perf_counter() - t0) / 10 * 1e3))
