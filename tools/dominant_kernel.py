"""Launches the dominant kernel of the step exactly as bench.py times it (adn_dense_fwd_p, [32768,1024]x[1024,1024]
+ bias + ReLU, planes in / planes out) a few times, for an `ncu --set full -k regex:pl_gemm_kernel` capture:

  ncu --set full --clock-control none --import-source on -k regex:pl_gemm_kernel -s 2 -c 1 -o gpurun_out/prof_dominant \
      python tools/dominant_kernel.py
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as g
g.build()
from adanet_b200 import _lib
import bench
lib = _lib.load()
_lib.check(lib.adn_init(), "init")
t, flops, path = bench.measure_dominant_kernel(lib, torch, reps=3)
print("dominant kernel %.1f us, %.1f useful TF/s (%s)" % (t * 1e6, flops / t / 1e12, path))
