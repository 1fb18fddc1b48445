import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g; g.build()
from tests import parity_util as pu
from tests.parity_util import orc
from adanet_b200.core import engine as eng, search as srch
d, c, B, steps, iters = 100, 10, 256, 20, 3
x, y = orc.make_tabular(8192, d, c, seed=21)
cfgs = [(1, 48), (2, 32), (3, 24)]
ens = dict(optimizer=("sgd", 0.02), adanet_lambda=0.01, use_bias=True, mixture_weight_type="matrix")
o, _ = orc.run_adanet_strategies(lambda t, f: pu.make_specs(cfgs, d, c, t, ("sgd", 0.02))[0], x, y, B, steps, iters, orc.EnsemblerSpec(**ens), c, strategies=("all",))
for graph in (True, False):
  s = srch.AdaNetSearch(lambda t, f: pu.make_specs(cfgs, d, c, t, ("sgd", 0.02))[1], eng.EnsemblerPlanSpec(**ens), d, c, B, strategies=("all",), use_cuda_graph=graph)
  reps = s.run(srch.consecutive_batches(x, y, B), steps, iters)
  for ro, r in zip(o, reps):
    for cname, tr in ro.traces.items():
      for f in ("sub_loss", "ens_loss", "adanet_loss", "ema"):
        e = np.abs(r.traces[cname][f].astype(np.float64) - np.asarray(tr[f], np.float64))
        if np.nanmax(e) > 2e-6:
          print("graph=%s it=%d %s %s max %.3g at step %d; errs %s" % (graph, ro.iteration, cname, f, np.nanmax(e), int(np.nanargmax(e)), np.round(e * 1e6, 1)))
print("done")
