"""GPU probe (not a pytest test): CUDA-event timings of adn_dense_fwd / adn_dense_bwd at
BASELINE sizes, L2 flushed between launches.  Also the target of ncu captures:

  python tests/probe_dense_big.py [--reps 5] [--shapes 1024x1024,100x1024,...] [--path auto|simt]
"""

import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--reps", type=int, default=5)
  ap.add_argument("--batch", type=int, default=32768)
  ap.add_argument("--shapes", default="1024x1024,100x1024,512x512,256x256,64x64,1024x10")
  ap.add_argument("--path", default="auto")
  args = ap.parse_args()
  import torch
  import __graft_entry__ as g
  g.build()
  from adanet_b200 import _lib
  lib = _lib.load()
  _lib.check(lib.adn_init(), "adn_init")
  _lib.set_dense_path({"auto": _lib.PATH_AUTO, "simt": _lib.PATH_SIMT, "tcgen05": _lib.PATH_TCGEN05}[args.path])
  st = torch.cuda.current_stream()
  sp = st.cuda_stream
  B = args.batch
  flush = torch.empty((256 * 1024 * 1024 // 4,), device="cuda")

  def timed(fn):
    ts = []
    for i in range(args.reps + 2):
      flush.fill_(float(i))
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      e0.record(st)
      fn()
      e1.record(st)
      e1.synchronize()
      if i >= 2:
        ts.append(e0.elapsed_time(e1) * 1e-3)
    return float(np.mean(ts))

  for shp in args.shapes.split(","):
    I, O = (int(v) for v in shp.split("x"))
    x = torch.relu(torch.randn((B, I), device="cuda"))
    w = torch.randn((I, O), device="cuda") * (1.0 / np.sqrt(I))
    b = torch.zeros((O,), device="cuda")
    y = torch.empty((B, O), device="cuda")
    dz = torch.randn((B, O), device="cuda") / B
    dx = torch.empty((B, I), device="cuda")
    dw = torch.empty((I, O), device="cuda")
    db = torch.empty((O,), device="cuda")
    fb = _lib.query(_lib.Q_DENSE_FWD_WS, B, I, O)
    bb = _lib.query(_lib.Q_DENSE_BWD_WS, B, I, O)
    fws = torch.empty((max(fb, 16),), dtype=torch.uint8, device="cuda")
    bws = torch.empty((max(bb, 16),), dtype=torch.uint8, device="cuda")
    flops = 2.0 * B * I * O
    tf = timed(lambda: _lib.check(lib.adn_dense_fwd(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), B, I, O, 1,
                                                    fws.data_ptr(), fb, sp), "fwd"))
    tb = timed(lambda: _lib.check(lib.adn_dense_bwd(x.data_ptr(), w.data_ptr(), dz.data_ptr(), dx.data_ptr(), dw.data_ptr(),
                                                    db.data_ptr(), B, I, O, 1, bws.data_ptr(), bb, sp), "bwd"))
    tbw = timed(lambda: _lib.check(lib.adn_dense_bwd(x.data_ptr(), w.data_ptr(), dz.data_ptr(), None, dw.data_ptr(),
                                                     db.data_ptr(), B, I, O, 1, bws.data_ptr(), bb, sp), "bwd"))
    print("B=%d %4dx%-4d fwd %.1f us (%.1f TF/s)  bwd(dW+db+dX) %.1f us (%.1f TF/s)  bwd(dW+db) %.1f us (%.1f TF/s)  path fwd=%d bwd=%d"
          % (B, I, O, tf * 1e6, flops / tf / 1e12, tb * 1e6, 2 * flops / tb / 1e12, tbw * 1e6, flops / tbw / 1e12,
             _lib.query(_lib.Q_DENSE_FWD_PATH, B, I, O), _lib.query(_lib.Q_DENSE_BWD_PATH, B, I, O)), flush=True)


if __name__ == "__main__":
  main()
