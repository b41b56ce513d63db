#!/usr/bin/env python
"""Pretty-print the JSON lines written by benchmarks/allreduce_sweep.py."""
import json
import sys

for path in sys.argv[1:]:
    print("==", path)
    for line in open(path):
        if not line.startswith("{"):
            continue
        d = json.loads(line)
        if d["kind"] == "all_reduce":
            extra = " | nccl %7.1f us %6.1f GB/s" % (d["nccl_ms"] * 1e3, d["nccl_bus_gbs"]) if "nccl_ms" in d else ""
            k = " (kernel %7.1f us %6.1f GB/s)" % (d["ours_kernel_ms"] * 1e3, d["ours_kernel_bus_gbs"]) if "ours_kernel_ms" in d else ""
            print("%-9s %11d B  ours %7.1f us %6.1f GB/s%s%s" % (d["dtype"], d["bytes"], d["ours_ms"] * 1e3, d["ours_bus_gbs"], k, extra))
        elif d["kind"] == "sync_model":
            extra = ""
            if "reference_path_nccl_ms" in d:
                extra = " | reference path over NCCL %6.0f us | one flat NCCL all-reduce %6.1f us %6.1f GB/s" % (
                    d["reference_path_nccl_ms"] * 1e3, d["nccl_flat_ms"] * 1e3, d["nccl_flat_bus_gbs"])
            k = " kernel %6.1f us = %6.1f GB/s bus;" % (d["ours_kernel_ms"] * 1e3, d["ours_kernel_bus_gbs"]) if "ours_kernel_ms" in d else ""
            print("%-8s %-8s sync_model %6.1f us = %6.1f GB/s bus;%s grads only %6.1f us%s" % (
                d["model"], d["dtype"], d["ours_ms"] * 1e3, d["ours_bus_gbs"], k, d["ours_grads_only_ms"] * 1e3, extra))
        else:
            print(d)
