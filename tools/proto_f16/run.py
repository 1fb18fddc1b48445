#!/usr/bin/env python
"""Builds and runs the fp16 hi/lo' GEMM prototype (tools/proto_f16/gemm_f16x2.cu) on a B200.

  python tools/proto_f16/run.py [M N K]

For C = A B^T with A [M, K], B [N, K] ~ N(0, 1) fp32 it prints, against an fp64 reference:
  * the error of the prototype (H + 2^-11 S, kind::f16) for chunk_kb in {0, 1, 2, 4}  (TMEM chain length)
  * the error of torch's plain fp32 matmul (TF32 off) as the yardstick
  * the kernel time / useful TFLOP/s (CUDA events), to compare with adn_dense_fwd_p on the same shape
    (tools/probe_planes.py): the expectation is ~2x the 3xTF32 rate at half the operand bytes.
Errors are reported as max |err| / max(|A| @ |B|^T) (the forward-error scale of an fp32 GEMM) and rms / rms(C).
"""
import ctypes
import os
import subprocess
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def build():
  so = os.path.join(HERE, "libproto_f16.so")
  src = os.path.join(HERE, "gemm_f16x2.cu")
  if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
                           "-Xcompiler", "-fPIC", "-shared", "-o", so, src, "-lcuda"])
  lib = ctypes.CDLL(so)
  lib.proto_gemm_f16x2.argtypes = [ctypes.c_void_p] * 5 + [ctypes.c_int] * 4 + [ctypes.c_void_p]
  lib.proto_gemm_f16x2.restype = ctypes.c_int
  return lib


def split(x):
  hi = x.half()
  lo = ((x - hi.float()) * 2048.0).half()
  return hi.contiguous(), lo.contiguous()


def main():
  M, N, K = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (4096, 1024, 1024)
  lib = build()
  torch.manual_seed(0)
  torch.backends.cuda.matmul.allow_tf32 = False
  a = torch.randn((M, K), device="cuda")
  b = torch.randn((N, K), device="cuda") / K ** 0.5
  ref = a.double() @ b.double().t()
  scale = float((a.abs().double() @ b.abs().double().t()).max())
  a_hi, a_lo = split(a)
  b_hi, b_lo = split(b)
  out = torch.empty((M, N), device="cuda")
  sp = torch.cuda.current_stream().cuda_stream

  def report(name, c):
    err = (c.double() - ref)
    print("%-28s max|err|/scale %.3e   rms(err)/rms(C) %.3e" % (name, float(err.abs().max()) / scale,
                                                               float(err.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt())))

  report("torch fp32 matmul", a @ b.t())
  for ck in (0, 1, 2, 4):
    out.zero_()
    rc = lib.proto_gemm_f16x2(a_hi.data_ptr(), a_lo.data_ptr(), b_hi.data_ptr(), b_lo.data_ptr(), out.data_ptr(), M, N, K, ck, sp)
    torch.cuda.synchronize()
    if rc != 0:
      raise SystemExit("proto_gemm_f16x2 failed: %d" % rc)
    report("fp16 hi/lo', chunk_kb=%d" % ck, out)
  # small-magnitude operand (a gradient-like tensor): with and without the power-of-two pre-scale
  for s_, tag in ((1.0, "unscaled"), (2.0 ** 15, "x 2^15")):
    g = (torch.randn((M, K), device="cuda") / 32768.0)
    g_hi, g_lo = split(g * s_)
    lib.proto_gemm_f16x2(g_hi.data_ptr(), g_lo.data_ptr(), b_hi.data_ptr(), b_lo.data_ptr(), out.data_ptr(), M, N, K, 2, sp)
    torch.cuda.synchronize()
    r = g.double() @ b.double().t()
    e = (out.double() / s_ - r)
    print("gradient-like A (%s): rms(err)/rms(C) %.3e" % (tag, float(e.pow(2).mean().sqrt() / r.pow(2).mean().sqrt())))
  # timing
  for _ in range(3):
    lib.proto_gemm_f16x2(a_hi.data_ptr(), a_lo.data_ptr(), b_hi.data_ptr(), b_lo.data_ptr(), out.data_ptr(), M, N, K, 2, sp)
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(20):
    lib.proto_gemm_f16x2(a_hi.data_ptr(), a_lo.data_ptr(), b_hi.data_ptr(), b_lo.data_ptr(), out.data_ptr(), M, N, K, 2, sp)
  e1.record()
  torch.cuda.synchronize()
  t = e0.elapsed_time(e1) * 1e-3 / 20
  print("kernel %.1f us  = %.1f useful TFLOP/s (non-persistent prototype, one 128x128 tile per CTA)" % (t * 1e6, 2.0 * M * N * K / t / 1e12))


if __name__ == "__main__":
  main()
