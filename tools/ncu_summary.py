#!/usr/bin/env python
"""Summarise ncu output for profiles/ (tracked evidence; gpurun_out/ is scratch).

  python tools/ncu_summary.py launches gpurun_out/launches.csv   > profiles/rN_launches.txt
  python tools/ncu_summary.py full gpurun_out/x.ncu-rep [idx]    > profiles/rN_kernel.txt

`launches`: per-kernel totals of the `--metrics gpu__time_duration.sum` launch list
(cold-cache, serialised: the SHARE of the step is what is comparable with bench.py).
`full`: the headline metrics of one `ncu --set full` capture (tensor pipe %, DRAM
bytes, L2, registers, smem), read with `ncu -i ... --page raw --csv`.
"""
import collections
import csv
import re
import subprocess
import sys


def launches(path):
  rows = list(csv.reader(open(path)))
  for i, r in enumerate(rows):
    if r and r[0] == "ID":
      hdr, start = r, i + 1
      break
  ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
  d = collections.OrderedDict()
  n = 0
  for r in rows[start:]:
    if len(r) <= vi:
      continue
    v = float(r[vi].replace(",", ""))
    if r[ui] == "us":
      v *= 1e3
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "")
    c = d.setdefault(name, [0, 0.0])
    c[0] += 1
    c[1] += v
    n += 1
  tot = sum(v[1] for v in d.values())
  print("# source: %s   launches: %d   total kernel time: %.1f us (ncu, cold-cache, serialised)" % (path, n, tot / 1e3))
  print("%12s %7s %7s %10s  %s" % ("total_us", "count", "share", "avg_us", "kernel"))
  for name, v in sorted(d.items(), key=lambda kv: -kv[1][1]):
    print("%12.1f %7d %6.1f%% %10.2f  %s" % (v[1] / 1e3, v[0], 100 * v[1] / tot, v[1] / v[0] / 1e3, name))


KEYS = [
    r"^gpu__time_duration\.sum$", r"^launch__grid_size$", r"^launch__block_size$", r"^launch__registers_per_thread$",
    r"^launch__shared_mem_per_block_dynamic$", r"^sm__cycles_active\.avg$",
    r"^sm__pipe_tensor_cycles_active\.avg\.pct_of_peak_sustained_(active|elapsed)$",
    r"^sm__throughput\.avg\.pct_of_peak_sustained_elapsed$", r"^sm__inst_executed_pipe_uniform\.avg\.pct",
    r"^dram__bytes_(read|write)\.sum$", r"^dram__bytes_(read|write)\.sum\.per_second$",
    r"^dram__throughput\.avg\.pct_of_peak_sustained_elapsed$", r"^lts__throughput\.avg\.pct_of_peak_sustained_elapsed$",
    r"^lts__t_sector_hit_rate\.pct$", r"^l1tex__m_xbar2l1tex_read_bytes\.sum(\.per_second)?$",
    r"^l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld\.sum$", r"^l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum$",
    r"^smsp__cycles_active\.avg$", r"^sm__warps_active\.avg\.pct_of_peak_sustained_active$",
]


def full(path, idx=0):
  out = subprocess.check_output(["ncu", "-i", path, "--page", "raw", "--csv"], text=True)
  rows = list(csv.reader(out.splitlines()))
  hdr, units = rows[0], rows[1]
  print("# source: %s  (ncu --set full --clock-control none)" % path)
  for k, vals in enumerate(rows[2:]):
    if idx is not None and k != idx:
      continue
    name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    print("## launch %d: %s" % (k, name[:150]))
    for h, u, v in zip(hdr, units, vals):
      if any(re.search(p, h) for p in KEYS):
        print("%-75s %-12s %s" % (h, u, v))


if __name__ == "__main__":
  if sys.argv[1] == "launches":
    launches(sys.argv[2])
  else:
    full(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 0)
