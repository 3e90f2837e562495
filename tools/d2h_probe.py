import torch, time
n = 1 << 30
d = [torch.empty(n, dtype=torch.uint8, device="cuda") for _ in range(2)]
h = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(2)]
def run(streams, parts):
    ss = [torch.cuda.Stream() for _ in range(streams)]
    torch.cuda.synchronize(); t = time.perf_counter()
    for rep in range(4):
        for i in range(2):
            for p in range(parts):
                lo, hi = p * n // parts, (p + 1) * n // parts
                with torch.cuda.stream(ss[(i * parts + p) % streams]):
                    h[i][lo:hi].copy_(d[i][lo:hi], non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    return 8 * n / dt / 1e9
for streams, parts in ((1, 1), (2, 1), (2, 2), (4, 4), (1, 8)):
    print("streams %d parts %d: %.1f GB/s" % (streams, parts, run(streams, parts)))
