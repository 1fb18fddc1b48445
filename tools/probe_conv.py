"""Times the SimpleCNN stem kernels (adn_conv_stem_fwd / _bwd) with CUDA events; prints achieved GB/s and GFMA/s.

  python tools/probe_conv.py [B ...]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adanet_b200 import _lib  # noqa: E402


def main():
  lib = _lib.load()
  _lib.check(lib.adn_init(), "adn_init")
  H = W = 32
  CIN, F = 3, 16
  sp = torch.cuda.current_stream().cuda_stream
  for B in [int(a) for a in sys.argv[1:]] or [256, 1024, 4096, 16384]:
    cols = (H // 2) * (W // 2) * F
    x = torch.rand((B, H, W, CIN), device="cuda")
    k = torch.randn((3, 3, CIN, F), device="cuda") * 0.27
    b = torch.zeros((F,), device="cuda")
    planes = torch.zeros((_lib.query(_lib.Q_PLANES_BYTES, B, cols) // 4,), device="cuda")
    arg = torch.zeros((B * cols // 16,), dtype=torch.int32, device="cuda")
    g = torch.randn((B, cols), device="cuda")
    dk = torch.empty_like(k)
    db = torch.empty_like(b)
    wsb = _lib.query(_lib.Q_CONV_STEM_BWD_WS, B, CIN, F)
    ws = torch.empty((wsb,), dtype=torch.uint8, device="cuda")
    flush = torch.empty((256 << 20,), dtype=torch.uint8, device="cuda")

    def fwd():
      _lib.check(lib.adn_conv_stem_fwd(x.data_ptr(), k.data_ptr(), b.data_ptr(), planes.data_ptr(), arg.data_ptr(), B, H, W,
                                       CIN, F, sp), "fwd")

    def bwd():
      _lib.check(lib.adn_conv_stem_bwd(x.data_ptr(), arg.data_ptr(), g.data_ptr(), dk.data_ptr(), db.data_ptr(), B, H, W, CIN,
                                       F, ws.data_ptr(), wsb, sp), "bwd")

    for name, fn, bytes_, fma in (("fwd", fwd, B * (H * W * CIN * 4 + cols * 8 + cols // 8 + cols // 4), B * H * W * 27 * F),
                                  ("bwd", bwd, B * (H * W * CIN * 4 + cols * 4 + cols // 4), B * cols * 27)):
      for _ in range(3):
        fn()
      ts = []
      for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
      t = float(np.median(ts))
      print("B=%6d %s: %8.1f us  %7.1f GB/s  %7.2f TFMA/s" % (B, name, t * 1e6, bytes_ / t / 1e9, fma / t / 1e12), flush=True)


if __name__ == "__main__":
  main()
