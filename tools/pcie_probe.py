"""D2H bandwidth from pinned memory: one stream vs the same bytes split over k streams (diagnostic)."""
import torch, time
dev = torch.device("cuda:0")
n = 105_578_496
src = torch.empty(n, dtype=torch.uint8, device=dev).random_()
dst = torch.empty(n, dtype=torch.uint8, pin_memory=True)
def run(k, reps=20):
    streams = [torch.cuda.Stream() for _ in range(k)]
    chunk = (n + k - 1) // k
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        for i, s in enumerate(streams):
            with torch.cuda.stream(s):
                lo, hi = i * chunk, min(n, (i + 1) * chunk)
                dst[lo:hi].copy_(src[lo:hi], non_blocking=True)
        torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    return n / dt / 1e9, dt * 1e3
for k in (1, 2, 4, 8):
    print(k, "streams: %.1f GB/s, %.3f ms" % run(k))
h = torch.empty(3_932_160, dtype=torch.uint8, pin_memory=True); d = torch.empty_like(h, device=dev)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(50): d.copy_(h, non_blocking=True); torch.cuda.synchronize()
print("H2D 3.9 MB: %.3f ms" % ((time.perf_counter() - t0) / 50 * 1e3))
