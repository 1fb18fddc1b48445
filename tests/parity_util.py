"""Shared parity helpers: drive the CUDA path through the C ABI and the CPU oracle
on the same seeded inputs.  Imported by tests/, __graft_entry__.smoke() and
bench.py only (the oracle is test infrastructure)."""

from __future__ import annotations

import os
import sys
from typing import List, Optional, Sequence

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

from oracle import adanet_oracle as orc  # noqa: E402


def dnn_dims(in_dim: int, width: int, depth: int, classes: int) -> List[int]:
  return [in_dim] + [width] * depth + [classes]


def make_specs(cfgs: Sequence[tuple], in_dim: int, classes: int, iteration: int, optimizer: tuple,
               base_seed: int = 1000, dropout=None):
  """cfgs: [(depth, width)] -> (oracle SubnetworkSpec list, engine SubnetworkPlanSpec list)
  with identical injected glorot-uniform weights (seed = 1000 + 100*t + i)."""
  from adanet_b200.core import engine as eng
  o_specs, e_specs = [], []
  names = set()
  for i, (depth, width) in enumerate(cfgs):
    dims = dnn_dims(in_dim, width, depth, classes)
    ws, bs = orc.init_mlp(dims, base_seed + 100 * iteration + i)
    name = orc.dnn_name(depth)
    if name in names:
      name = "{}_w{}".format(name, width)
    names.add(name)
    cx = float(np.sqrt(np.float32(depth)))   # simple_dnn.py:88-90
    drop = [tuple(dropout)] * depth if (dropout is not None and depth > 0) else None     # (rate, seed) after every hidden layer
    o_specs.append(orc.SubnetworkSpec(name, dims, cx, optimizer, ws=ws, bs=bs, dropout=drop))
    e_specs.append(eng.SubnetworkPlanSpec(name, dims, cx, optimizer, [w.copy() for w in ws], [b.copy() for b in bs],
                                          shared={"num_layers": depth}, dropout=drop))
  return o_specs, e_specs


def make_cnn_specs(seeds: Sequence[int], image_shape, filters: int, hidden: int, classes: int, iteration: int,
                   optimizer: tuple):
  """SimpleCNN subnetworks (customizing_adanet.ipynb SimpleCNNBuilder: conv3x3(F)+ReLU -> maxpool2 -> flatten ->
  dense(hidden)+ReLU -> dense(classes), complexity 1) differing by seed, with he_normal-scaled injected weights.
  -> (oracle SubnetworkSpec list, engine SubnetworkPlanSpec list)."""
  from adanet_b200.core import engine as eng
  h, w, cin = image_shape
  dims = [(h // 2) * (w // 2) * filters, hidden, classes]
  o_specs, e_specs = [], []
  for sd in seeds:
    rng = np.random.default_rng(7000 + 100 * iteration + sd)
    he = lambda shape, fan_in: (rng.standard_normal(shape) * np.sqrt(2.0 / fan_in)).astype(np.float32)
    ws = [he((3, 3, cin, filters), 9 * cin), he((dims[0], hidden), dims[0]), he((hidden, classes), hidden)]
    bs = [np.zeros((filters,), np.float32), np.zeros((hidden,), np.float32), np.zeros((classes,), np.float32)]
    name = "simple_cnn_s{}".format(sd)
    o_specs.append(orc.SubnetworkSpec(name, dims, 1.0, optimizer, ws=ws, bs=bs))
    e_specs.append(eng.SubnetworkPlanSpec(name, dims, 1.0, optimizer, [w_.copy() for w_ in ws], [b.copy() for b in bs],
                                          image_shape=tuple(image_shape)))
  return o_specs, e_specs


import contextlib


@contextlib.contextmanager
def oracle_noise(level: float, seed: int):
  """Temporarily perturbs the oracle's dense forward outputs and weight gradients by
  `level` relative Gaussian noise: a model of a *different but equally valid* fp32
  implementation, used to prove a parity configuration is well conditioned."""
  rng = np.random.default_rng(seed)
  f0, b0 = orc.mlp_forward, orc.mlp_backward

  def fwd(ws, bs, x, dropout=None):
    acts = [np.asarray(x, dtype=np.float32)]
    n = len(ws)
    for i in range(n):
      z = acts[-1] @ ws[i] + bs[i]
      z = z * (1 + np.float32(level) * rng.standard_normal(z.shape).astype(np.float32))
      if i < n - 1:
        z = np.maximum(z, np.float32(0))
        d = dropout[0][i] if (dropout is not None and dropout[0] is not None) else None
        if d is not None:
          keep = orc.dropout_keep_mask(d[1], i, dropout[1], z.shape[0], z.shape[1], d[0])
          z = np.where(keep, z.astype(np.float32) * np.float32(1.0 / (1.0 - float(d[0]))), np.float32(0))
      acts.append(z.astype(np.float32))
    return acts

  def bwd(ws, acts, dlogits, dropout=None):
    dws, dbs = b0(ws, acts, dlogits, dropout)
    dws = [(d * (1 + np.float32(level) * rng.standard_normal(d.shape).astype(np.float32))).astype(np.float32)
           for d in dws]
    return dws, dbs

  orc.mlp_forward, orc.mlp_backward = fwd, bwd
  try:
    yield
  finally:
    orc.mlp_forward, orc.mlp_backward = f0, b0


def rel_err(a, b) -> float:
  a = np.asarray(a, dtype=np.float64)
  b = np.asarray(b, dtype=np.float64)
  return float(np.abs(a - b).max() / max(1e-30, np.abs(b).max()))


def run_smoke():
  """One tiny AdaNet iteration step on cuda:0, checked against the oracle."""
  import torch
  from adanet_b200.core import engine as eng
  from adanet_b200.core import search as srch
  assert torch.cuda.is_available(), "smoke() needs a GPU"
  torch.cuda.set_device(0)
  B, D, C = 256, 100, 10
  x, y = orc.make_tabular(B * 4, D, C, seed=77)
  cfgs = [(1, 64), (2, 128)]
  opt = ("sgd", 0.05)
  ens_o = orc.EnsemblerSpec(optimizer=("sgd", 0.01), adanet_lambda=0.01, adanet_beta=0.001)
  ens_e = eng.EnsemblerPlanSpec(optimizer=("sgd", 0.01), adanet_lambda=0.01, adanet_beta=0.001)

  def o_space(t, frozen):
    return make_specs(cfgs, D, C, t, opt)[0]

  def e_space(t, frozen):
    return make_specs(cfgs, D, C, t, opt)[1]

  o_res, _ = orc.run_adanet(o_space, x, y, B, 4, 1, ens_o, C)
  s = srch.AdaNetSearch(e_space, ens_e, D, C, B)
  reps = s.run(srch.consecutive_batches(x, y, B), 4, 1)
  for name, tr in o_res[0].traces.items():
    got = reps[0].traces[name]
    for f in ("sub_loss", "adanet_loss", "ema"):
      want = np.asarray(tr[f], dtype=np.float64)
      err = np.abs(got[f].astype(np.float64) - want).max()
      assert err < 1e-5, "smoke parity %s/%s: max abs err %.3g" % (name, f, err)
  assert reps[0].best_index == o_res[0].best_index
  from adanet_b200 import _lib
  assert _lib.launch_count() > 0
