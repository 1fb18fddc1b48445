#!/usr/bin/env python
"""Compact view of one `ncu --set full` capture: duration, pipes, issue rate and the PC-sampling stall mix.

  python tools/ncu_stalls.py gpurun_out/x.ncu-rep [launch_index]
"""
import csv
import subprocess
import sys

WANT = ("gpu__time_duration.sum", "sm__cycles_active.avg", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors_op_write.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed")


def main():
  rep = sys.argv[1]
  idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
  out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
  rows = list(csv.reader(out.splitlines()))
  hdr, r = rows[0], rows[2 + idx]
  print("# %s  launch %d: %s" % (rep, idx, r[4][:90]))
  for h, v in zip(hdr, r):
    if h in WANT:
      print("%-84s %s" % (h, v))
  st = []
  for h, v in zip(hdr, r):
    if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h:
      try:
        st.append((float(v), h))
      except ValueError:
        pass
  tot = sum(v for v, _ in st) or 1.0
  for v, h in sorted(st, reverse=True)[:9]:
    print("pcsamp %-70s %5.1f%%" % (h.replace("smsp__pcsamp_warps_issue_stalled_", ""), 100 * v / tot))


if __name__ == "__main__":
  main()
