"""Host-to-device rate of a 1.44 GB pinned buffer: one stream, two streams (halves), piece sizes."""
import time, torch
n = 1_440_000_000
src = torch.empty(n, dtype=torch.uint8).pin_memory()
dst = torch.empty(n, dtype=torch.uint8, device="cuda")
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
def one():
    dst.copy_(src, non_blocking=True)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def two():
    h = n // 2
    with torch.cuda.stream(s1):
        dst[:h].copy_(src[:h], non_blocking=True)
    with torch.cuda.stream(s2):
        dst[h:].copy_(src[h:], non_blocking=True)
    s1.synchronize(); s2.synchronize()
def pieces(mb):
    def f():
        step = mb << 20
        for a in range(0, n, step):
            dst[a:a + step].copy_(src[a:a + step], non_blocking=True)
    return f
print(f"one copy: {n / timed(one) / 1e9:.2f} GB/s")
print(f"two streams: {n / timed(two) / 1e9:.2f} GB/s")
for mb in (16, 64, 256):
    print(f"pieces of {mb} MB: {n / timed(pieces(mb)) / 1e9:.2f} GB/s")
