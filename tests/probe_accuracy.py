"""GPU probe (not a pytest test): error statistics of the dense paths vs fp64.

  python tests/probe_accuracy.py

For K in {128, 1024, 4096}: signed mean (bias) and rms of (got - exact) relative
to the rms magnitude of the exact result, for the SIMT fp32 path and the
tcgen05 3xTF32 path.  Used to decide accumulation strategy.
"""

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
  import torch
  import __graft_entry__ as g
  g.build()
  from adanet_b200 import _lib
  lib = _lib.load()
  _lib.check(lib.adn_init(), "adn_init")
  sp = torch.cuda.current_stream().cuda_stream
  rng = np.random.default_rng(0)
  for positive in (False, True):
    for K in (128, 1024, 4096):
      B, O = 512, 256
      x = rng.standard_normal((B, K)).astype(np.float32)
      w = (rng.standard_normal((K, O)) / np.sqrt(K)).astype(np.float32)
      if positive:     # all-positive operands: partial sums grow monotonically -> exposes truncation bias
        x, w = np.abs(x), np.abs(w)
      exact = x.astype(np.float64) @ w.astype(np.float64)
      mag = np.sqrt((exact ** 2).mean())
      xd, wd = torch.as_tensor(x).cuda(), torch.as_tensor(w).cuda()
      yd = torch.empty((B, O), device="cuda")
      for name, path in (("simt", _lib.PATH_SIMT), ("tcgen05", _lib.PATH_TCGEN05)):
        _lib.set_dense_path(path)
        nb = _lib.query(_lib.Q_DENSE_FWD_WS, B, K, O)
        ws = torch.empty((max(nb, 16),), dtype=torch.uint8, device="cuda")
        _lib.check(lib.adn_dense_fwd(xd.data_ptr(), wd.data_ptr(), None, yd.data_ptr(), B, K, O, 0, ws.data_ptr(), nb, sp),
                   "fwd")
        err = yd.cpu().numpy().astype(np.float64) - exact
        print("positive=%d K=%5d %-8s bias/mag=% .3e  rms/mag=%.3e  max/mag=%.3e" %
              (positive, K, name, err.mean() / mag, np.sqrt((err ** 2).mean()) / mag, np.abs(err).max() / mag))
      t = (torch.as_tensor(x).cuda() @ torch.as_tensor(w).cuda()).cpu().numpy().astype(np.float64) - exact
      print("positive=%d K=%5d %-8s bias/mag=% .3e  rms/mag=%.3e  max/mag=%.3e" %
            (positive, K, "cublas", t.mean() / mag, np.sqrt((t ** 2).mean()) / mag, np.abs(t).max() / mag))
      c = (x @ w).astype(np.float64) - exact
      print("positive=%d K=%5d %-8s bias/mag=% .3e  rms/mag=%.3e  max/mag=%.3e" %
            (positive, K, "numpy", c.mean() / mag, np.sqrt((c ** 2).mean()) / mag, np.abs(c).max() / mag))
  _lib.set_dense_path(_lib.PATH_AUTO)


if __name__ == "__main__":
  main()
