"""Development probe: host->device copy rates of a strided column block of a pinned (n, m) int8 matrix."""
import sys, time
import torch
n, m, mb = 10000, 100000, 25088
X = torch.empty((n, m), dtype=torch.int8, pin_memory=True)
X.random_(0, 3)
dst = torch.empty((n, mb), dtype=torch.int8, device="cuda")
flat = torch.empty(n * mb, dtype=torch.int8, pin_memory=True)
dflat = torch.empty(n * mb, dtype=torch.int8, device="cuda")
def t(fn, reps=5):
    torch.cuda.synchronize(); fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
a = t(lambda: dst.copy_(X[:, :mb], non_blocking=True))
b = t(lambda: dflat.copy_(flat, non_blocking=True))
print(f"2D strided block {n}x{mb} of pitch {m}: {n*mb/a/1e9:.1f} GB/s ; contiguous {n*mb/b/1e9:.1f} GB/s")
for w in (4096, 25000, 25088, 32768):
    d2 = torch.empty((n, w), dtype=torch.int8, device="cuda")
    c = t(lambda: d2.copy_(X[:, :w], non_blocking=True))
    print(f"  width {w}: {n*w/c/1e9:.1f} GB/s")
