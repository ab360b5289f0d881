#!/usr/bin/env python
"""Markdown table of one step from an ncu launch list made with tools/profile_step.py (metrics of
tools/r02z_gpu_profile.sh): per kernel launches, time, DRAM bytes, executed FP64 flops.

    python tools/summarize_step_csv.py gpurun_out/r02z_step_cartpole.csv 2 65536
"""
import csv
import re
import sys
from collections import OrderedDict

path, steps, units = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
hdr = rows[0]
ik, im, iv, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
per = OrderedDict()
for r in rows[1:]:
    name = r[ik]
    m = re.search(r"(k_\w+)(<[^>]*(?:<[^>]*>[^>]*)*>)?", name)
    if not m:
        continue  # torch helper kernels of the workload driver
    key = m.group(1) + (m.group(2) or "")
    key = key.replace("rlmpc::", "").replace("(int)", "")
    d = per.setdefault(key, {"ids": set()})
    d["ids"].add(r[iid])
    try:
        val = float(r[iv].replace(",", ""))
    except ValueError:  # "n/a": metric not collected for this kernel
        val = 0.0
    d[r[im]] = d.get(r[im], 0.0) + val
tot = {"ms": 0.0, "rd": 0.0, "wr": 0.0, "fl": 0.0}
print("| kernel | launches | ms (ncu, serialised) | DRAM read MB | DRAM write MB | FP64 GFLOP executed | of which DMMA | TFLOP/s |")
print("|---|---|---|---|---|---|---|---|")
for k, d in per.items():
    ms = d.get("gpu__time_duration.sum", 0.0) / 1e6 / steps
    rd = d.get("dram__bytes_read.sum", 0.0) / 1e6 / steps
    wr = d.get("dram__bytes_write.sum", 0.0) / 1e6 / steps
    dmma = 512.0 * d.get("sm__inst_executed_pipe_tensor_op_dmma.sum", 0.0) / 1e9 / steps
    fl = (2.0 * d.get("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", 0.0) + d.get("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", 0.0)
          + d.get("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", 0.0)) / 1e9 / steps + dmma
    print(f"| `{k}` | {len(d['ids']) // steps} | {ms:.3f} | {rd:.1f} | {wr:.1f} | {fl:.2f} | {dmma:.2f} | {fl / ms if ms else 0:.2f} |")
    tot["ms"] += ms; tot["rd"] += rd; tot["wr"] += wr; tot["fl"] += fl
print(f"| **total** | | {tot['ms']:.3f} | {tot['rd']:.0f} | {tot['wr']:.0f} | {tot['fl']:.2f} | | |")
print()
print(f"DRAM traffic {(tot['rd'] + tot['wr']) / 1e3:.3f} GB per step = {(tot['rd'] + tot['wr']) * 1e6 / units:.0f} B per unit; "
      f"executed FP64 flops per unit {tot['fl'] * 1e9 / units:.0f}.")
