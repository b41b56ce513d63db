#!/usr/bin/env python
"""Kernel-only timing of the bucketed all-reduce with W virtual ranks on ONE GPU (HBM-bound).

Drives the C ABI directly from one thread (no Python rendezvous): one plan, W pointer rows,
`iters` launches bracketed by CUDA events on the launch stream, L2 flushed between launches.

    python benchmarks/loopback_kernel.py [--world 8] [--dtype bf16] [--sizes-kb 4 64 1024 21845]
    python benchmarks/loopback_kernel.py --resnet18      # the 62-tensor gradient bucket
"""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import torch  # noqa: E402

from flashy_b200 import _native as N  # noqa: E402
from flashy_b200.engine import Engine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--world", type=int, default=8)
    ap.add_argument("--dtype", default="bf16", choices=("bf16", "fp32"))
    ap.add_argument("--sizes-kb", type=int, nargs="*", default=[4, 64, 512, 4096, 21845, 65536])
    ap.add_argument("--resnet18", action="store_true")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--algo", type=int, default=0)
    ap.add_argument("--no-flush", action="store_true")
    args = ap.parse_args()
    tdtype, fx = (torch.bfloat16, N.FX_BF16) if args.dtype == "bf16" else (torch.float32, N.FX_F32)
    eng = Engine(n_local=args.world, device=0, arena_mb=192)
    stream = torch.cuda.current_stream()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    cases = []
    if args.resnet18:
        import torchvision
        cases.append(("resnet18", [p.numel() for p in torchvision.models.resnet18(num_classes=10).parameters()]))
    else:
        for kb in args.sizes_kb:
            cases.append((f"{kb}KiB", [kb * 1024 // tdtype.itemsize]))
    for name, numels in cases:
        plan = eng.get_plan("ar", tuple(numels), fx, fx, args.algo)
        rows = [[torch.randn(n, device="cuda").to(tdtype) for n in numels] for _ in range(args.world)]
        ptrs = [[t.data_ptr() for t in row] for row in rows]
        for _ in range(3):
            eng.allreduce(plan, N.FX_AVG, ptrs, ptrs, stream)
        torch.cuda.synchronize()
        times = []
        for _ in range(args.iters):
            if not args.no_flush:
                flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            eng.allreduce(plan, N.FX_AVG, ptrs, ptrs, stream)
            e1.record(stream)
            e1.synchronize()
            times.append(e0.elapsed_time(e1))
        times.sort()
        ms = sum(times) / len(times)
        nbytes = sum(numels) * tdtype.itemsize
        W = args.world
        hbm = W * (5 + 1 / W) * nbytes if plan.info.algo != N.FX_ALGO_ONE_SHOT else W * (2 + W + 1) * nbytes
        print(json.dumps({"case": name, "world": W, "dtype": args.dtype, "bytes_per_rank": nbytes,
                          "algo": N.ALGO_NAMES[plan.info.algo], "grid_x": plan.info.grid_x,
                          "mean_us": 1e3 * ms, "min_us": 1e3 * times[0],
                          "hbm_alg_gbs": hbm / (ms * 1e-3) / 1e9}), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
