#!/usr/bin/env python
"""End-to-end rate of the PUBLIC `adanet.Estimator.train` call with host (pageable NumPy) minibatches.

  python tools/bench_estimator.py [--batch 32768] [--steps 60] [--layer-size 512]

examples/simple_dnn.Generator (2 candidates per iteration), one AdaNet iteration of `steps` steps; wall clock around
`train` minus the first (graph-capture) step is not separable from outside, so two runs are timed (steps and 2*steps)
and the difference gives the steady-state per-step time.  Prints one JSON line.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--batch", type=int, default=32768)
  ap.add_argument("--steps", type=int, default=60)
  ap.add_argument("--layer-size", type=int, default=512)
  a = ap.parse_args()
  import torch
  import adanet_b200 as adanet
  from adanet_b200 import graph, train
  from adanet_b200.examples import simple_dnn
  D, C, B = 100, 10, a.batch
  rng = np.random.default_rng(0)
  x = rng.standard_normal((B * 8, D)).astype(np.float32)
  y = rng.integers(0, C, B * 8)

  def input_fn(n):
    def fn():
      for i in range(n):
        o = (i % 8) * B
        yield {"x": x[o:o + B]}, y[o:o + B]
    return fn

  def run(n):
    gen = simple_dnn.Generator(feature_columns=[graph.numeric_column("x", D)], optimizer=train.GradientDescentOptimizer(0.01),
                               layer_size=a.layer_size, initial_num_layers=1, seed=1)
    est = adanet.Estimator(head=adanet.heads.MultiClassHead(C), subnetwork_generator=gen, max_iteration_steps=10 ** 9,
                           max_iterations=1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    est.train(input_fn(n), max_steps=n)
    torch.cuda.synchronize()
    return time.perf_counter() - t0

  run(5)                    # warm the library / allocator
  t1, t2 = run(a.steps), run(2 * a.steps)
  per_step = (t2 - t1) / a.steps
  print(json.dumps({"api": "adanet_b200.Estimator.train, pageable NumPy input_fn", "batch": B, "layer_size": a.layer_size,
                    "candidates": 2, "ms_per_step": per_step * 1e3, "examples_per_s": B / per_step,
                    "h2d_bytes_per_step": B * D * 4 + B * 8}))


if __name__ == "__main__":
  main()
