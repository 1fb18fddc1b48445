"""GPU debug probe for csrc/planes.cu: error of every operand-majorness combination."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as g
g.build()
from adanet_b200 import _lib
lib = _lib.load()
_lib.check(lib.adn_init(), "init")
sp = torch.cuda.current_stream().cuda_stream

def planes(a):
  r, c = a.shape
  pl = torch.zeros((_lib.query(_lib.Q_PLANES_BYTES, r, c) // 4,), device="cuda")
  src = torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32)).cuda()
  _lib.check(lib.adn_planes_split(src.data_ptr(), r, c, pl.data_ptr(), sp), "split")
  return pl

def relerr(got, exact):
  mag = np.sqrt((exact.astype(np.float64) ** 2).mean()) + 1e-30
  return float(np.abs(got.astype(np.float64) - exact).max() / mag)

for (B, I, O) in [(128, 32, 128), (128, 128, 128), (256, 64, 256), (300, 100, 70)]:
  rng = np.random.default_rng(1)
  x = rng.standard_normal((B, I)).astype(np.float32)
  w = rng.standard_normal((I, O)).astype(np.float32)
  dz = rng.standard_normal((B, O)).astype(np.float32)
  xp, wp, dzp = planes(x), planes(w), planes(dz)
  y = torch.zeros((B, O), device="cuda")
  _lib.check(lib.adn_dense_fwd_p(xp.data_ptr(), wp.data_ptr(), None, None, y.data_ptr(), B, I, O, 0, sp), "fwd")
  nb = _lib.query(_lib.Q_DENSE_BWD_P_WS, B, I, O)
  ws = torch.empty((nb,), dtype=torch.uint8, device="cuda")
  dw = torch.zeros((I, O), device="cuda"); dx = torch.zeros((B, I), device="cuda")
  _lib.check(lib.adn_dense_bwd_p(xp.data_ptr(), wp.data_ptr(), dzp.data_ptr(), None, dx.data_ptr(), None, dw.data_ptr(),
                                 B, I, O, 0, 0, ws.data_ptr(), nb, sp), "bwd")
  torch.cuda.synchronize()
  ye = x.astype(np.float64) @ w
  yy = y.cpu().numpy()
  print("B=%d I=%d O=%d  fwd(A K,B MN) %.3e   dX(A K,B K) %.3e   dW(A MN,B MN) %.3e   |y|max %.3e" % (
      B, I, O, relerr(yy, ye), relerr(dx.cpu().numpy(), dz.astype(np.float64) @ w.T.astype(np.float64)),
      relerr(dw.cpu().numpy(), x.T.astype(np.float64) @ dz), np.abs(yy).max()))
  if relerr(yy, ye) > 1e-3 and B <= 128:
    # which entries match? try to identify a permutation: compare y to x @ w with w columns permuted
    r = yy[0, :8]; print("  y[0,:8]   ", r); print("  exact     ", ye[0, :8])
