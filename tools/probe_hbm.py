"""GPU probe: pure-write / pure-read / copy HBM bandwidth with torch ops (1 GiB buffers)."""
import torch
n = 1 << 28
a = torch.empty(n, device="cuda"); b = torch.empty(n, device="cuda")
def t(fn, reps=5):
  best = 1e9
  for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); e1.synchronize()
    best = min(best, e0.elapsed_time(e1) * 1e-3)
  return best
w = t(lambda: a.fill_(1.0)); print("fill  (write 1 GiB): %.1f us  %.0f GB/s" % (w * 1e6, n * 4 / w / 1e9))
r = t(lambda: a.sum());      print("sum   (read  1 GiB): %.1f us  %.0f GB/s" % (r * 1e6, n * 4 / r / 1e9))
c = t(lambda: b.copy_(a));   print("copy  (r+w  2 GiB): %.1f us  %.0f GB/s" % (c * 1e6, 2 * n * 4 / c / 1e9))
# write bandwidth with incompressible data: a 64 MiB random source (L2-resident) broadcast into the 1 GiB buffer
src = torch.randn(n // 16, device="cuda")
bw = b.view(16, n // 16)
e = t(lambda: bw.copy_(src.view(1, -1).expand(16, -1)))
print("write (random data, source L2-resident, 1 GiB): %.1f us  %.0f GB/s" % (e * 1e6, n * 4 / e / 1e9))
z = t(lambda: a.zero_()); print("zero_ (write 1 GiB): %.1f us  %.0f GB/s" % (z * 1e6, n * 4 / z / 1e9))
