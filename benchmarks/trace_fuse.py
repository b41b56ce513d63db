#!/usr/bin/env python
"""Per-chunk timeline of the fused all-reduce kernel (k_fuse), from its own %globaltimer stamps.

    FLASHY_B200_TRACE=1 torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 benchmarks/trace_fuse.py

CTA 0 of rank 0 records when each warp role hands a chunk on (fx_fuse.cu, FZ_TR_*); this script runs
``sync_model``-sized buckets, reads the stamps of the last launch through ``fx_comm_trace_read`` and
prints where the microseconds of one call go: launch -> metadata -> pack -> flags -> reduce -> fence
-> flags -> unpack, plus the steady-state period of every role.  Benchmark infrastructure only.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch                      # noqa: E402
import torch.distributed as dist  # noqa: E402

TR_CHUNKS = 1024
TR_PACK = 16
TR_UNPACK = TR_PACK + 4 * TR_CHUNKS
TR_RED = TR_UNPACK + 4 * TR_CHUNKS
TR_SIG = TR_RED + 4 * TR_CHUNKS
TR_POLLP = TR_SIG + 2 * TR_CHUNKS
TR_POLLR = TR_POLLP + TR_CHUNKS
TR_CTA = TR_POLLR + TR_CHUNKS


def main():
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("gloo", init_method="env://")
    from flashy_b200 import _native as N
    from flashy_b200 import context as fctx
    from flashy_b200 import distrib
    import torchvision

    def read_trace():
        eng = fctx.current().engine
        words = C.c_size_t()
        N.check(N.lib.fx_comm_trace_read(eng.comm, None, 0, C.byref(words)))
        if not words.value:
            return None
        buf = (C.c_uint64 * words.value)()
        N.check(N.lib.fx_comm_trace_read(eng.comm, buf, words.value, C.byref(words)))
        return list(buf)

    cases = [("resnet18", torch.bfloat16), ("resnet50", torch.float32)]
    for name, dtype in cases:
        model = (torchvision.models.resnet18(num_classes=10) if name == "resnet18" else torchvision.models.resnet50()).to(dev).to(dtype)
        for p in model.parameters():
            p.grad = torch.randn_like(p) * 1e-2
        for _ in range(3):
            distrib.sync_model(model)
        torch.cuda.synchronize()
        read_trace()                                   # clear
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        distrib.sync_model(model)
        e1.record()
        torch.cuda.synchronize()
        t = read_trace()
        if rank != 0:
            continue
        if t is None:
            print(json.dumps({"case": name, "error": "tracing is off (set FLASHY_B200_TRACE=1)"}))
            continue
        t0, chunks = t[0], int(t[2])
        us = lambda x: None if x == 0 else round((x - t0) / 1e3, 2)    # noqa: E731
        rows = []
        for c in range(min(chunks, TR_CHUNKS)):
            rows.append({
                "c": c,
                "pack_load": us(t[TR_PACK + 4 * c]), "pack_landed": us(t[TR_PACK + 4 * c + 1]),
                "pack_stored": us(t[TR_PACK + 4 * c + 2]), "pack_signalled": us(t[TR_PACK + 4 * c + 3]),
                "all_packed": us(t[TR_POLLP + c]),
                "red_wait": us(t[TR_RED + 4 * c]), "red_go": us(t[TR_RED + 4 * c + 1]), "red_done": us(t[TR_RED + 4 * c + 2]),
                "sig_seen": us(t[TR_SIG + 2 * c]), "sig_fenced": us(t[TR_SIG + 2 * c + 1]),
                "all_reduced": us(t[TR_POLLR + c]),
                "unp_go": us(t[TR_UNPACK + 4 * c]), "unp_load": us(t[TR_UNPACK + 4 * c + 1]),
                "unp_landed": us(t[TR_UNPACK + 4 * c + 2]), "unp_stored": us(t[TR_UNPACK + 4 * c + 3]),
            })
        nbytes = sum(p.numel() for p in model.parameters()) * dtype.itemsize

        def period(key):
            vals = [r[key] for r in rows if r[key] is not None]
            return round((vals[-1] - vals[0]) / max(1, len(vals) - 1), 3) if len(vals) > 1 else None
        ctas = [(t[TR_CTA + 2 * i], t[TR_CTA + 2 * i + 1]) for i in range(512) if t[TR_CTA + 2 * i]]
        starts = sorted((a - t0) / 1e3 for a, _ in ctas)
        ends = sorted((z - t0) / 1e3 for _, z in ctas)
        durs = sorted((z - a) / 1e3 for a, z in ctas)
        q = lambda v, f: round(v[min(len(v) - 1, int(f * len(v)))], 1)      # noqa: E731
        cta = {"n": len(ctas), "start_min_med_max": [q(starts, 0), q(starts, 0.5), q(starts, 1)],
               "end_min_med_max": [q(ends, 0), q(ends, 0.5), q(ends, 1)],
               "duration_min_med_max": [q(durs, 0), q(durs, 0.5), q(durs, 1)],
               "slowest_ctas": sorted(range(len(ctas)), key=lambda i: -(ctas[i][1] - t0))[:6]}
        print(json.dumps({
            "case": f"{name}/{str(dtype).split('.')[-1]}", "ctas": cta, "world": world, "bytes": nbytes, "chunks_per_cta": chunks,
            "event_us": round(e0.elapsed_time(e1) * 1e3, 1), "meta_loaded_us": us(t[1]), "exit_us": us(t[3]),
            "period_us": {k: period(k) for k in ("pack_load", "pack_signalled", "all_packed", "red_done", "sig_fenced",
                                                 "all_reduced", "unp_stored")},
            "first": rows[:6], "last": rows[-3:],
        }))
        del model
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
