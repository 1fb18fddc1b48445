#!/usr/bin/env python
"""bench.py -- candidate-train examples/sec per AdaNet iteration (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one training step of EVERY candidate of the iteration on one
minibatch (subnetwork fwd+bwd+update, candidate-ensemble head, EMA).  Workload
(config.workload): BASELINE configs[2] -- 8-candidate DNN search 100->H->H->10,
H in {64..1024}, 1M x 100 synthetic tabular data, B = 32768; it fits one GPU,
so N=1 trains all 8 candidates on one B200 and N>1 shards whole candidates across
the GPUs, cost-balanced (strong scaling: total work fixed, no data-path collective;
the only exchange is the end-of-iteration loss all_gather, outside the step).

Prints ONE JSON line on rank 0.  `value` = B*K / device time (CUDA events, max
over ranks) with the dataset resident in HBM; `e2e` = same metric through
the public adanet_b200.Estimator.train call with HOST (pinned) batches, H2D of every
batch and a D2H read of every step's losses (an `after_run` hook) inside the timed region.
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WIDTHS = (64, 128, 192, 256, 384, 512, 768, 1024)
IN_DIM, CLASSES, BATCH = 100, 10, 32768
DATA_ROWS = 1_000_000
METRIC = "candidate-train examples/sec per AdaNet iteration"
REF_SAMPLE_ROWS = 4096   # rows per step of the --impl reference arm (bounded sample)


def workload_name(gpus):
  return ("configs[2]: 8-candidate DNN search 100->H->H->10, H in %s, 1Mx100 tabular synthetic, B=%d, "
          "candidates placed on %d GPU(s) cost-balanced (LPT over train FLOPs)" % (list(WIDTHS), BATCH, gpus))


def train_flops_per_example():
  return sum(6 * (IN_DIM * h + h * h + h * CLASSES) - 2 * IN_DIM * h for h in WIDTHS)


class ClockSampler:
  """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md)."""
  Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
       "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
       "clocks_event_reasons.sw_power_cap")

  def __init__(self, index=0):
    self.index, self.proc, self.lines = index, None, []

  def start(self):
    try:
      self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                    "--format=csv,noheader,nounits", "-lms", "50"],
                                   stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
      self.t = threading.Thread(target=self._read, daemon=True)
      self.t.start()
    except Exception:
      self.proc = None

  def _read(self):
    for line in self.proc.stdout:
      self.lines.append(line.strip())

  def stop(self):
    if not self.proc:
      return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
    self.proc.terminate()
    try:
      self.proc.wait(timeout=5)
    except Exception:
      self.proc.kill()
    sm, mx, reasons = [], [], set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    for l in self.lines:
      f = [v.strip() for v in l.split(",")]
      if len(f) < 9:
        continue
      try:
        sm.append(float(f[1]))
        mx.append(float(f[2]))
      except ValueError:
        continue
      for n, v in zip(names, f[5:9]):
        if v.lower().startswith("active"):
          reasons.add(n)
    return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
            "reasons": sorted(reasons), "samples": len(sm)}


def load_peaks():
  p = os.path.join(ROOT, "MEASURED_PEAKS.json")
  if os.path.exists(p):
    with open(p) as f:
      d = json.load(f)
    return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
  return 6650.0, 1590.0, 1400.0, "fallback"


def oracle_specs():
  from tests import parity_util as pu
  return [(2, h) for h in WIDTHS]


def _best_thread_count(step_fn, cores):
  """NumPy/OpenBLAS on a many-core host is often fastest well below the core count (oversubscription, NUMA):
  time one step at a few thread counts and keep the best, so the CPU arm is the strongest the port can give."""
  try:
    from threadpoolctl import threadpool_limits
  except Exception:
    return cores, None, {}
  cand = sorted({c for c in (cores, 96, 64, 48, 32, 16, 8) if 1 <= c <= cores}, reverse=True)
  timings = {}
  for c in cand:
    with threadpool_limits(limits=c):
      step_fn()                                  # warm the pools at this width
      t0 = time.perf_counter()
      step_fn()
      timings[c] = time.perf_counter() - t0
  best = min(timings, key=timings.get)
  return best, threadpool_limits, timings


def run_reference(args):
  """--impl reference: the CPU restatement of the reference's path (oracle port; the
  real reference needs TensorFlow 2.1, not installable here) on all host cores."""
  rank = int(os.environ.get("RANK", "0"))
  if rank != 0:
    return
  from tests import parity_util as pu
  from oracle import adanet_oracle as orc
  cores = os.cpu_count() or 1
  rows = REF_SAMPLE_ROWS   # a bounded sample of the B=32768 minibatch per step, so K steps finish in minutes
  x, y = orc.make_tabular(rows * 4, IN_DIM, CLASSES, seed=1234)
  o_specs, _ = pu.make_specs(oracle_specs(), IN_DIM, CLASSES, 0, ("sgd", 0.05))
  ens = orc.EnsemblerSpec(optimizer=("sgd", 0.01), adanet_lambda=0.01, adanet_beta=0.001)
  cands = orc.build_candidates(0, o_specs, [], ens, CLASSES, 0.9)
  it = 0

  def step():
    nonlocal it
    off = (it % 4) * rows
    orc.train_step(cands, [], ens, x[off:off + rows], y[off:off + rows])
    it += 1

  threads, limiter, sweep = _best_thread_count(step, cores)
  ctx = limiter(limits=threads) if limiter is not None else None
  if ctx is not None:
    ctx.__enter__()
  try:
    for _ in range(args.warmup):
      step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
      step()
    dt = time.perf_counter() - t0
  finally:
    if ctx is not None:
      ctx.__exit__(None, None, None)
  val = rows * args.steps / dt
  line = {
      "impl": "reference", "metric": METRIC, "value": val, "unit": "examples/s", "n_gpus": args.gpus,
      "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
      "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
      "config": {"workload": workload_name(args.gpus), "candidates": len(WIDTHS), "batch": BATCH},
      "cpu_baseline": {"value": val, "unit": "examples/s", "cores": threads, "kind": "port",
                       "sample": "%d steps, each a %d-row sample of the B=%d minibatch of the same 8-candidate "
                                 "workload (NumPy/OpenBLAS fp32 oracle; %d host cores, best of a thread-count sweep "
                                 "%s s/step; the TF1 reference itself is not installable: TensorFlow 2.1 absent)"
                                 % (args.steps, rows, BATCH, cores,
                                    {k: round(v, 3) for k, v in sorted(sweep.items())})},
      "e2e": {"value": val, "unit": "examples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
  }
  print(json.dumps(line), flush=True)


def cpu_baseline_sample(seconds_budget=15.0):
  """The oracle (CPU port of the reference's path) on all host cores, on the same bounded sample
  per step as the --impl reference arm."""
  from tests import parity_util as pu
  from oracle import adanet_oracle as orc
  cores = os.cpu_count() or 1
  rows = REF_SAMPLE_ROWS
  x, y = orc.make_tabular(rows * 4, IN_DIM, CLASSES, seed=1234)
  o_specs, _ = pu.make_specs(oracle_specs(), IN_DIM, CLASSES, 0, ("sgd", 0.05))
  ens = orc.EnsemblerSpec(optimizer=("sgd", 0.01), adanet_lambda=0.01, adanet_beta=0.001)
  cands = orc.build_candidates(0, o_specs, [], ens, CLASSES, 0.9)
  for i in range(2):   # warm-up
    orc.train_step(cands, [], ens, x[i * rows:(i + 1) * rows], y[i * rows:(i + 1) * rows])
  threads, limiter, sweep = _best_thread_count(lambda: orc.train_step(cands, [], ens, x[:rows], y[:rows]), cores)
  ctx = limiter(limits=threads) if limiter is not None else None
  if ctx is not None:
    ctx.__enter__()
  try:
    n, t0 = 0, time.perf_counter()
    while True:
      off = (n % 4) * rows
      orc.train_step(cands, [], ens, x[off:off + rows], y[off:off + rows])
      n += 1
      dt = time.perf_counter() - t0
      if dt > seconds_budget or n >= 200:
        break
  finally:
    if ctx is not None:
      ctx.__exit__(None, None, None)
  return {"value": rows * n / dt, "unit": "examples/s", "cores": threads, "kind": "port",
          "sample": "%d steps (%.1f s), each a %d-row sample of the B=%d minibatch of the same 8-candidate workload, "
                    "NumPy/OpenBLAS fp32 oracle, %d host cores, best thread count of a sweep %s s/step"
                    % (n, dt, rows, BATCH, cores, {k: round(v, 3) for k, v in sorted(sweep.items())})}


def measure_dominant_kernel(lib, torch, reps=20):
  """CUDA-event timing of the dominant kernel of the step -- the H=1024 hidden-layer dense forward
  [32768,1024]x[1024,1024] + bias + ReLU, planes in / planes out, exactly as the engine launches it
  (adn_dense_fwd_p) -- on the stream it is launched on, L2 flushed between launches."""
  from adanet_b200 import _lib
  from adanet_b200.core import engine as eng
  B, I, O = BATCH, 1024, 1024
  sp = torch.cuda.current_stream()
  x = torch.randn((B, I), device="cuda")
  w = torch.randn((I, O), device="cuda") * 0.03
  b = torch.zeros((O,), device="cuda")
  xp, wp, yp = eng.new_planes(B, I, "cuda"), eng.new_planes(I, O, "cuda"), eng.new_planes(B, O, "cuda")
  _lib.check(lib.adn_planes_split(x.data_ptr(), B, I, xp.data_ptr(), sp.cuda_stream), "split")
  _lib.check(lib.adn_planes_split(w.data_ptr(), I, O, wp.data_ptr(), sp.cuda_stream), "split")
  flush = torch.empty((256 * 1024 * 1024 // 4,), device="cuda")   # 256 MB > 126 MB L2
  times = []
  for i in range(reps + 3):
    flush.fill_(float(i))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(sp)
    _lib.check(lib.adn_dense_fwd_p(xp.data_ptr(), wp.data_ptr(), b.data_ptr(), yp.data_ptr(), None, B, I, O, 1,
                                   sp.cuda_stream), "adn_dense_fwd_p")
    e1.record(sp)
    e1.synchronize()
    if i >= 3:
      times.append(e0.elapsed_time(e1) * 1e-3)
  return float(np.mean(times)), 2.0 * B * I * O, "tcgen05_3xtf32_planes"


def run_ours(args):
  import torch
  import torch.distributed as dist
  world = int(os.environ.get("WORLD_SIZE", "1"))
  rank = int(os.environ.get("RANK", "0"))
  local = int(os.environ.get("LOCAL_RANK", "0"))
  torch.cuda.set_device(local)
  if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
  import __graft_entry__ as g
  g.build()
  from adanet_b200 import _lib
  from adanet_b200.core import engine as eng
  from adanet_b200.core import search as srch
  from adanet_b200.distributed import exchange as ex
  from tests import parity_util as pu
  from oracle import adanet_oracle as orc   # data + weight generation only (host side, outside timed regions)
  lib = _lib.load()
  _lib.check(lib.adn_init(), "adn_init")
  dev = torch.device("cuda", local)

  # synthetic data, generated once on the host, replicated per GPU
  x_np, y_np = orc.make_tabular(DATA_ROWS, IN_DIM, CLASSES, seed=1234)
  x_dev = torch.as_tensor(x_np).to(dev)
  y_dev = torch.as_tensor(y_np).to(dev)
  ens = eng.EnsemblerPlanSpec(optimizer=("sgd", 0.01), adanet_lambda=0.01, adanet_beta=0.001)
  space = lambda t, frozen: pu.make_specs(oracle_specs(), IN_DIM, CLASSES, t, ("sgd", 0.05))[1]

  # ---------------- value: dataset resident in HBM ----------------
  s = srch.AdaNetSearch(space, ens, IN_DIM, CLASSES, BATCH, device=dev, keep_traces=False)
  plan = s.build_iteration()
  batches = srch.consecutive_batches(x_dev, y_dev, BATCH)
  for _ in range(args.warmup):
    plan.train_step(*next(batches))
  torch.cuda.synchronize()
  if world > 1:
    dist.barrier()
  sampler = ClockSampler(local)
  if rank == 0:
    sampler.start()
  l0 = _lib.launch_count()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(args.steps):
    plan.train_step(*next(batches))
  e1.record()
  torch.cuda.synchronize()
  if world > 1:
    dist.barrier()
  secs = ex.max_over_ranks(e0.elapsed_time(e1) * 1e-3, device=dev)
  clocks = sampler.stop() if rank == 0 else None
  launches_local = plan.launches_per_step * args.steps if plan.launches_per_step else _lib.launch_count() - l0
  value = BATCH * args.steps / secs
  local_losses = plan.last_losses()
  assert np.isfinite(local_losses).all(), "non-finite loss in the timed region"
  rep = s.finish_iteration(secs)   # end-of-iteration all_gather + selection (outside the timed region)
  if args.profile:   # under ncu: only the in-HBM step loop (a number printed under a profiler is never a bench value)
    if rank == 0:
      print(json.dumps({"profile_only": True, "ms_per_step_under_profiler": secs / args.steps * 1e3,
                        "gpu_launches": int(launches_local)}), flush=True)
    if world > 1:
      dist.destroy_process_group()
    return

  # ---------------- e2e: host batches through the PUBLIC API (adanet.Estimator.train) ----------------
  e2e_steps = min(args.steps, 100)
  n_host = BATCH * 4
  x_host = torch.as_tensor(x_np[:n_host]).pin_memory()
  y_host = torch.as_tensor(y_np[:n_host]).pin_memory()
  warm = max(3, args.warmup)

  def host_batches(n):
    def fn():
      for i in range(n):
        o = (i % 4) * BATCH
        yield {"x": x_host[o:o + BATCH]}, y_host[o:o + BATCH]
    return fn

  e2e_api, e2e_note, e2e_secs, d2h = "adanet_b200.Estimator.train", None, None, 0
  try:
    if world > 1:
      # under torchrun the same measurement goes through the engine-level search API (what Estimator.train drives):
      # the Estimator path was validated against it at N=1 (equal rates) but not under NCCL in this round
      raise RuntimeError("N > 1: engine-level API")
    import adanet_b200 as adanet
    from adanet_b200 import graph, train

    class _WidthBuilder(adanet.subnetwork.Builder):
      """100 -> H -> H -> 10 with the same injected weights as the device-resident run."""

      def __init__(self, spec):
        self._spec = spec

      name = property(lambda self: self._spec.name)

      def build_subnetwork(self, features, logits_dimension, training, iteration_step, summary, previous_ensemble=None):
        h = graph.input_layer(features, [graph.numeric_column("x", IN_DIM)])
        n = len(self._spec.ws)
        for i, (w, b) in enumerate(zip(self._spec.ws, self._spec.bs)):
          h = graph.dense(h, w.shape[1], activation=graph.relu if i < n - 1 else None,
                          kernel_initializer=graph.constant_initializer(w), bias_initializer=graph.constant_initializer(b))
          if i == n - 2:
            last = h
        return adanet.Subnetwork(last_layer=last, logits=h, complexity=self._spec.complexity)

      def build_subnetwork_train_op(self, subnetwork, loss, var_list, labels, iteration_step, summary, previous_ensemble):
        return train.GradientDescentOptimizer(0.05).minimize(loss=loss, var_list=var_list)

    class _Losses:
      last = None

      def after_run(self, run_context, run_values):      # D2H read of every step's losses
        self.last = run_values.results["losses"]

    est = adanet.Estimator(
        head=adanet.heads.MultiClassHead(CLASSES),
        subnetwork_generator=adanet.subnetwork.SimpleGenerator([_WidthBuilder(sp) for sp in space(0, [])]),
        max_iteration_steps=10 ** 9, max_iterations=1,
        ensemblers=[adanet.ensemble.ComplexityRegularizedEnsembler(optimizer=train.GradientDescentOptimizer(0.01),
                                                                   adanet_lambda=0.01, adanet_beta=0.001)])
    hook = _Losses()
    est.train(host_batches(warm), steps=warm, hooks=[hook])          # builds the plan, captures the graph
    torch.cuda.synchronize()
    if world > 1:
      dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    # every step: host->device copy of its (pinned) minibatch, started under the previous step's kernels, and a
    # device->host read of its losses through the hook
    est.train(host_batches(e2e_steps), steps=e2e_steps, hooks=[hook])
    e1.record()
    torch.cuda.synchronize()
    e2e_secs = e0.elapsed_time(e1) * 1e-3
    d2h = int(hook.last.nbytes)
  except Exception as exc:      # measured below through the engine-level search API instead; the reason is reported
    e2e_api, e2e_note = "adanet_b200.core.search.AdaNetSearch.train_iteration", "Estimator path not used: %r" % (exc,)
    s2 = srch.AdaNetSearch(space, ens, IN_DIM, CLASSES, BATCH, device=dev, keep_traces=False)
    plan2 = s2.build_iteration()
    hb = srch.consecutive_batches(x_host, y_host, BATCH)
    for _ in range(warm):
      plan2.train_step(*next(hb))
      plan2.last_losses()
    torch.cuda.synchronize()
    host_losses = None

    def read_losses(plan):
      nonlocal host_losses
      host_losses = plan.last_losses()

    e2e_secs = s2.train_iteration(hb, e2e_steps, on_step=read_losses)
    torch.cuda.synchronize()
    d2h = int(host_losses.nbytes)
  if world > 1:
    dist.barrier()
  e2e_secs = ex.max_over_ranks(e2e_secs, device=dev)
  e2e = {"value": BATCH * e2e_steps / e2e_secs, "unit": "examples/s",
         "h2d_bytes_per_step": BATCH * IN_DIM * 4 + BATCH * 8, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
         "api": e2e_api}
  if e2e_note:
    e2e["note"] = e2e_note

  if rank != 0:
    if world > 1:
      dist.destroy_process_group()
    return

  # ---------------- roofline of the dominant kernel + CPU baseline (rank 0, N=1 only for cpu) ----------------
  hbm, bf16_burst, bf16_sust, which = load_peaks()
  kt, kflops, kpath = measure_dominant_kernel(lib, torch)
  traffic = None   # dram bytes per launch of that kernel from the committed ncu --set full capture
  try:
    with open(os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")) as f:
      traffic = int(json.load(f)["traffic_bytes"])
  except Exception:
    pass
  achieved = kflops / kt / 1e12
  roofline = {
      "bound": "tensor", "kernel": "adn_dense_fwd_p [32768,1024]x[1024,1024] bias+relu, planes in/out (%s)" % kpath,
      "achieved": achieved, "peak": bf16_burst, "unit": "TFLOP/s", "frac": achieved / bf16_burst,
      "peak_source": "MEASURED_PEAKS.json bf16 burst (%s); algorithmic fp32 FLOPs 2*B*in*out; the tcgen05 path "
                     "issues 3 TF32 MMAs per product (3xTF32 split for 1e-5 fp32 parity), TF32 dense peak = bf16/2"
                     % which,
      "issued_frac_of_tf32_peak": (3.0 * achieved / (bf16_burst / 2.0)) if kpath.startswith("tcgen05") else None,
      "traffic": traffic, "traffic_unit": "bytes/launch (ncu dram read+write; algorithmic 553.6 MB of split planes)",
      "launch_seconds": kt,
  }
  cpu = cpu_baseline_sample() if world == 1 else None
  line = {
      "metric": METRIC, "value": value, "unit": "examples/s", "n_gpus": world, "steps": args.steps,
      "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
      "vs_baseline": None, "dtype": "f32", "data": "synthetic",
      "config": {"workload": workload_name(world), "candidates": len(WIDTHS), "batch": BATCH,
                 "l2_policy": "inputs larger than L2: every step reads a fresh 32768x100 slice of the 400 MB "
                              "HBM-resident dataset and streams >1 GB of activations; weights stay cache-resident "
                              "as in real training",
                 "train_flops_per_example": train_flops_per_example(),
                 "candidate_examples_per_sec": value * len(WIDTHS),
                 "cuda_graph": True, "selected": rep.candidate_names[rep.best_index]},
      "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches_local),
      "roofline": roofline, "cpu_baseline": cpu,
      "useful_tflops": value * train_flops_per_example() / 1e12,
  }
  print(json.dumps(line), flush=True)
  if world > 1:
    dist.destroy_process_group()


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=200)
  ap.add_argument("--warmup", type=int, default=5)
  ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
  ap.add_argument("--profile", action="store_true", help="step loop only (for ncu captures)")
  args = ap.parse_args()
  if args.impl == "reference":
    run_reference(args)
  else:
    run_ours(args)


if __name__ == "__main__":
  main()
