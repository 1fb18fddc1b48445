"""GPU timing probe (not a test): every GEMM of one 100->H->H->10 candidate step on the plane pipeline.

  python tools/probe_planes.py [--H 1024] [--B 32768] [--reps 5]
Prints per-call device time (CUDA events on the launch stream, 256 MB L2 flush between calls) and the
issued-TF32 rate (3 MMAs per product) for each call."""
import argparse, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as g
g.build()
from adanet_b200 import _lib
from adanet_b200.core import engine as eng
lib = _lib.load()
_lib.check(lib.adn_init(), "init")
ap = argparse.ArgumentParser()
ap.add_argument("--H", type=int, nargs="+", default=[1024])
ap.add_argument("--B", type=int, default=32768)
ap.add_argument("--reps", type=int, default=5)
args = ap.parse_args()
st = torch.cuda.current_stream(); sp = st.cuda_stream
flush = torch.empty((256 * 1024 * 1024 // 4,), device="cuda")

def planes(r, c, scale=1.0):
  t = torch.randn((r, c), device="cuda") * scale
  pl = eng.new_planes(r, c, "cuda")
  _lib.check(lib.adn_planes_split(t.data_ptr(), r, c, pl.data_ptr(), sp), "split")
  return pl

def timeit(fn):
  ts = []
  for i in range(args.reps + 2):
    flush.fill_(float(i))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st); fn(); e1.record(st); e1.synchronize()
    if i >= 2: ts.append(e0.elapsed_time(e1) * 1e3)
  return float(np.median(ts))

B, D, C = args.B, 100, 10
for H in args.H:
  xp, h1p, h2p = planes(B, D), planes(B, H), planes(B, H)
  w1p, w2p, wop = planes(D, H, 0.1), planes(H, H, 0.03), planes(H, C, 0.03)
  dzo, dz2, dz1 = planes(B, C, 1e-4), planes(B, H, 1e-4), planes(B, H, 1e-4)
  b1 = torch.zeros((H,), device="cuda"); bo = torch.zeros((C,), device="cuda")
  logits = torch.empty((B, C), device="cuda")
  dw1, dw2, dwo = torch.empty((D, H), device="cuda"), torch.empty((H, H), device="cuda"), torch.empty((H, C), device="cuda")
  db = torch.empty((H,), device="cuda")
  nb = max(_lib.query(_lib.Q_DENSE_BWD_P_WS, B, i, o) for i, o in ((D, H), (H, H), (H, C)))
  ws = torch.empty((nb,), dtype=torch.uint8, device="cuda")
  P = lambda t: t.data_ptr()
  calls = [
      ("fwd L1  [B,%d]x[%d,%d] planes" % (D, D, H), 2.0 * B * D * H, lambda: lib.adn_dense_fwd_p(P(xp), P(w1p), P(b1), P(h1p), None, B, D, H, 1, sp)),
      ("fwd L2  [B,%d]x[%d,%d] planes" % (H, H, H), 2.0 * B * H * H, lambda: lib.adn_dense_fwd_p(P(h1p), P(w2p), P(b1), P(h2p), None, B, H, H, 1, sp)),
      ("fwd out [B,%d]x[%d,%d] dense " % (H, H, C), 2.0 * B * H * C, lambda: lib.adn_dense_fwd_p(P(h2p), P(wop), P(bo), None, P(logits), B, H, C, 0, sp)),
      ("bwd out dW[%d,%d]+dX planes" % (H, C), 4.0 * B * H * C, lambda: lib.adn_dense_bwd_p(P(h2p), P(wop), P(dzo), P(dz2), None, P(db), P(dwo), B, H, C, 1, 0, P(ws), nb, sp)),
      ("bwd out dW only", 2.0 * B * H * C, lambda: lib.adn_dense_bwd_p(P(h2p), P(wop), P(dzo), None, None, None, P(dwo), B, H, C, 1, 0, P(ws), nb, sp)),
      ("bwd L2  dW[%d,%d]+dX planes" % (H, H), 4.0 * B * H * H, lambda: lib.adn_dense_bwd_p(P(h1p), P(w2p), P(dz2), P(dz1), None, P(db), P(dw2), B, H, H, 1, 0, P(ws), nb, sp)),
      ("bwd L2  dW only", 2.0 * B * H * H, lambda: lib.adn_dense_bwd_p(P(h1p), P(w2p), P(dz2), None, None, None, P(dw2), B, H, H, 1, 0, P(ws), nb, sp)),
      ("bwd L1  dW[%d,%d]" % (D, H), 2.0 * B * D * H, lambda: lib.adn_dense_bwd_p(P(xp), P(w1p), P(dz1), None, None, None, P(dw1), B, D, H, 0, 0, P(ws), nb, sp)),
  ]
  tot = 0.0
  for name, flops, fn in calls:
    rc = fn()
    _lib.check(rc, name)
    us = timeit(fn)
    if "only" not in name: tot += us
    print("H=%4d %-34s %8.1f us  %7.1f useful TF/s  %7.1f issued TF32 TF/s" % (H, name, us, flops / us / 1e6, 3 * flops / us / 1e6))
  print("H=%4d GEMM total per step %.1f us" % (H, tot))
