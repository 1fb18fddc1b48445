"""GPU timing probe (not a test): the fused ensemble head (adn_ensemble_head: weighted logit sum + softmax-CE +
complexity penalty + mixture-weight gradient) at BASELINE batch sizes and at 1M rows, as achieved HBM GB/s over
its algorithmic bytes  N*C*4 (member logits) + 8 (label)  per example."""
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
st = torch.cuda.current_stream(); sp = st.cuda_stream
flush = torch.empty((256 * 1024 * 1024 // 4,), device="cuda")
C = 10
for B in (32768, 1 << 20):
  for N in (1, 2, 5):
    members = [torch.randn((B, C), device="cuda") for _ in range(N)]
    labels = torch.randint(0, C, (B,), device="cuda", dtype=torch.int64)
    w = torch.full((N,), 1.0 / N, device="cuda"); bias = torch.zeros((C,), device="cuda")
    out3 = torch.zeros((3,), device="cuda"); dw = torch.zeros((N,), device="cuda")
    nb = _lib.query(_lib.Q_HEAD_WS, B, C, N); ws = torch.empty((nb,), dtype=torch.uint8, device="cuda")
    mp = _lib.ptr_array([m.data_ptr() for m in members]); gam = _lib.f32_array([0.011] * N)
    def run():
      _lib.check(lib.adn_ensemble_head(_lib.HEAD_SOFTMAX_XENT, _lib.MIX_SCALAR, mp, N, w.data_ptr(), bias.data_ptr(), gam, 0, 2.0,
                                       labels.data_ptr(), None, out3.data_ptr(), dw.data_ptr(), None, None, None, B, C,
                                       ws.data_ptr(), nb, sp), "head")
    run()
    ts = []
    for i in range(7):
      flush.fill_(float(i))
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      e0.record(st); run(); e1.record(st); e1.synchronize()
      if i >= 2: ts.append(e0.elapsed_time(e1) * 1e-3)
    t = float(np.median(ts)); byt = B * (N * C * 4 + 8)
    print("B=%8d N=%d  %8.1f us  algorithmic %7.1f MB  %7.1f GB/s" % (B, N, t * 1e6, byt / 1e6, byt / t / 1e9))
