"""Does a device -> pinned host copy go faster as two concurrent halves (two copy engines) than as one?  (PCIe ceiling of the e2e legs)"""
import time

import torch

n = 1 << 31  # 2 GiB
src = torch.empty(n, dtype=torch.uint8, device="cuda")
dst = torch.empty(n, dtype=torch.uint8).pin_memory()
s = [torch.cuda.Stream() for _ in range(4)]


def run(k, chunk=None):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if chunk is None:
        step = n // k
        for i in range(k):
            with torch.cuda.stream(s[i]):
                dst[i * step:(i + 1) * step].copy_(src[i * step:(i + 1) * step], non_blocking=True)
    else:
        i = 0
        for a in range(0, n, chunk):
            with torch.cuda.stream(s[i % k]):
                dst[a:a + chunk].copy_(src[a:a + chunk], non_blocking=True)
            i += 1
    torch.cuda.synchronize()
    return n / (time.perf_counter() - t0) / 1e9


for k in (1, 2, 4, 1, 2):
    print(f"{k} stream(s), contiguous parts: {max(run(k) for _ in range(3)):.1f} GB/s")
for k in (1, 2):
    print(f"{k} stream(s), 128 MiB chunks round robin: {max(run(k, 128 << 20) for _ in range(3)):.1f} GB/s")
