"""Host<->device copy bandwidth of this box with pinned buffers (explains the e2e number of bench.py)."""
import torch, time
dev = torch.device("cuda:0")
def bw(nbytes, direction, streams=1, reps=10):
    n = nbytes // 4
    hs = [torch.empty(n // streams, dtype=torch.float32).pin_memory() for _ in range(streams)]
    ds = [torch.empty(n // streams, dtype=torch.float32, device=dev) for _ in range(streams)]
    ss = [torch.cuda.Stream() for _ in range(streams)]
    def go():
        for h, d, s in zip(hs, ds, ss):
            with torch.cuda.stream(s):
                (d.copy_(h, non_blocking=True) if direction == "h2d" else h.copy_(d, non_blocking=True))
    go(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): go()
    torch.cuda.synchronize()
    return nbytes * reps / (time.perf_counter() - t0) / 1e9
for mb in (8, 64, 256):
    for d in ("h2d", "d2h"):
        for st in (1, 2, 4):
            print("%4d MB %s streams=%d: %.1f GB/s" % (mb, d, st, bw(mb << 20, d, st)))
# both directions at once
h1 = torch.empty(16 << 20, dtype=torch.float32).pin_memory(); d1 = torch.empty(16 << 20, dtype=torch.float32, device=dev)
h2 = torch.empty(16 << 20, dtype=torch.float32).pin_memory(); d2 = torch.empty(16 << 20, dtype=torch.float32, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10):
    with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("duplex 64 MB each way: %.1f GB/s per direction" % (64 * (1 << 20) * 10 / dt / 1e9))
