#!/usr/bin/env python
"""BASELINE configs[4]: ``distrib.all_reduce`` bandwidth sweep 4 KiB - 1 GiB, and the
``sync_model``-sized buckets of ResNet-18 / ResNet-50, next to the library path the reference
calls (torch.distributed / NCCL: one flat all-reduce, and the reference's per-tensor loop).

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 benchmarks/allreduce_sweep.py [--quick]

Timing: CUDA events on the launching stream, >= 3 warm-ups, L2 flushed (256 MiB memset) before
every timed iteration, mean over iterations, max over ranks.  busBW = 2 (W-1)/W * bytes / t.
Writes one JSON object per line to stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch                      # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--max-mb", type=int, default=1024)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--no-nccl", action="store_true")
    ap.add_argument("--only-sync", action="store_true", help="skip the flat all_reduce sweep")
    ap.add_argument("--tag", default="")
    ap.add_argument("--symm", action="store_true", help="also time torch.ops.symm_mem.{multimem,two_shot,one_shot}_all_reduce")
    ap.add_argument("--variants", default="", help="';'-separated env settings, e.g. 'FLASHY_B200_FUSE_DEPTH=4;FLASHY_B200_FUSE=0': "
                    "the sync_model cases are repeated under each (plans are rebuilt in between)")
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("cpu:gloo,cuda:nccl", init_method="env://")
    from flashy_b200 import distrib
    from flashy_b200 import _native as N
    from flashy_b200 import context as fctx
    from oracle.refdistrib import RefDistrib

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timeit(fn, iters):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        total = 0.0
        for _ in range(iters):
            flush.zero_()
            dist.barrier(device_ids=[local]) if False else None
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            total += e0.elapsed_time(e1)
        t = torch.tensor([total / iters], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)          # gloo (CPU tensor)
        return float(t[0])

    def kernel_ms(fn, iters):
        """Device time of our launches alone (CUDA events bracketing each launch on its stream),
        summed per call: what the kernels cost once the host is not the bottleneck."""
        eng = fctx.current().engine
        fn()
        torch.cuda.synchronize()
        eng.profile, eng.timings = True, []
        for _ in range(iters):
            flush.zero_()
            fn()
        torch.cuda.synchronize()
        eng.profile = False
        total = sum(e0.elapsed_time(e1) for _, e0, e1 in eng.timings)
        eng.timings = []
        t = torch.tensor([total / iters], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    def emit(**kw):
        if rank == 0:
            print(json.dumps(kw), flush=True)

    bus = lambda nbytes, ms: 2 * (world - 1) / world * nbytes / (ms * 1e-3) / 1e9   # noqa: E731

    # ------------------------------------------------------------------ torch.ops.symm_mem comparators
    symm = {"error": None}
    if not args.symm:
        symm["error"] = "not requested (--symm)"
    elif not args.no_nccl and not args.only_sync:
        try:
            import torch.distributed._symmetric_memory as symm_mem
            group = dist.group.WORLD
            pool = symm_mem.empty(min(args.max_mb, 1024) << 20, dtype=torch.uint8, device=dev)
            symm_mem.rendezvous(pool, group=group)
            symm.update(pool=pool, name=group.group_name)
        except Exception as err:      # noqa: BLE001 - comparator only
            symm["error"] = f"{type(err).__name__}: {err}"[:200]

    def symm_times(nbytes, dtype, iters):
        """ms of PyTorch's own symmetric-memory all-reduces on the same bytes, or the reason they did not run."""
        if symm["error"] is not None:
            return {} if not args.symm else {"symm_mem_error": symm["error"]}
        out = {}
        view = symm["pool"][:nbytes].view(dtype)
        ops = {"symm_multimem_ms": lambda: torch.ops.symm_mem.multimem_all_reduce_(view, "sum", symm["name"]),
               "symm_two_shot_ms": lambda: torch.ops.symm_mem.two_shot_all_reduce_(view, "sum", symm["name"])}
        if nbytes <= (1 << 20):
            ops["symm_one_shot_ms"] = lambda: torch.ops.symm_mem.one_shot_all_reduce(view, "sum", symm["name"])
        for key, fn in ops.items():
            try:
                out[key] = timeit(fn, iters)
            except Exception as err:      # noqa: BLE001
                out[key.replace("_ms", "_error")] = f"{type(err).__name__}: {err}"[:160]
        return out

    # ------------------------------------------------------------------ flat all_reduce sweep
    sizes = [4096 << k for k in range(0, 19)]
    sizes = [s for s in sizes if s <= args.max_mb << 20]
    if args.quick:
        sizes = sizes[::3]
    if args.only_sync:
        sizes = []
    for dtype in (torch.float32, torch.bfloat16):
        for nbytes in sizes:
            n = nbytes // dtype.itemsize
            x = torch.randn(n, device=dev, dtype=torch.float32).to(dtype)
            iters = args.iters if nbytes <= (64 << 20) else max(5, args.iters // 4)
            ours = timeit(lambda: distrib.all_reduce(x), iters)
            kern = kernel_ms(lambda: distrib.all_reduce(x), iters)
            row = dict(kind="all_reduce", dtype=str(dtype).split(".")[-1], bytes=nbytes, world=world,
                       ours_ms=ours, ours_bus_gbs=bus(nbytes, ours), ours_kernel_ms=kern, ours_kernel_bus_gbs=bus(nbytes, kern))
            if not args.no_nccl:
                nccl = timeit(lambda: dist.all_reduce(x), iters)
                row.update(nccl_ms=nccl, nccl_bus_gbs=bus(nbytes, nccl))
                row.update(symm_times(nbytes, dtype, iters))
            emit(**row)
            del x

    # ------------------------------------------------------------------ sync_model sized buckets
    import torchvision

    def drop_plans():
        """Collective: forget every cached plan / layout so that the next call re-plans under the current env."""
        from flashy_b200 import distrib as D
        torch.cuda.synchronize()
        dist.barrier()
        eng = fctx.current().engine
        for plan in eng.plans.values():
            plan.destroy()
        eng.plans.clear()
        eng.layouts.clear()
        eng.fast_lists.clear()
        for entry in list(D._model_cache.values()):
            entry.fast.clear()
        dist.barrier()

    variants = [v for v in args.variants.split(";") if v] if args.variants else [""]
    for variant in variants:
        saved = {}
        for kv in [x for x in variant.split(",") if x and x != "default"]:
            k, v = kv.split("=", 1)
            saved[k] = os.environ.get(k)
            os.environ[k] = v
        if args.variants:
            drop_plans()
        for name, ctor in (("resnet18", lambda: torchvision.models.resnet18(num_classes=10)),
                           ("resnet50", torchvision.models.resnet50)):
            for dtype in (torch.bfloat16, torch.float32):
                torch.manual_seed(1234)
                model = ctor().to(dev).to(dtype)
                for p in model.parameters():
                    p.grad = torch.randn_like(p) * 1e-2
                params = list(model.parameters())
                nbytes = sum(p.numel() for p in params) * dtype.itemsize
                ours = timeit(lambda: distrib.sync_model(model), args.iters)
                grads_only = timeit(lambda: distrib.sync_gradients(model.parameters()), args.iters)
                grads_list = timeit(lambda: distrib.sync_gradients(params), args.iters)
                kern = kernel_ms(lambda: distrib.sync_model(model), args.iters)
                eng = fctx.current().engine
                big = max(eng.plans.values(), key=lambda pl: pl.info.wire_bytes)
                row = dict(kind="sync_model", model=name, dtype=str(dtype).split(".")[-1], grad_bytes=nbytes, world=world,
                           variant=variant, kernel=N.KERNEL_NAMES.get(int(big.info.kernel)), chunks=int(big.info.chunks),
                           chunk_bytes=int(big.info.chunk_bytes), grid=int(big.info.grid_x),
                           ours_ms=ours, ours_grads_only_ms=grads_only, ours_grads_list_ms=grads_list, ours_bus_gbs=bus(nbytes, ours),
                           ours_kernel_ms=kern, ours_kernel_bus_gbs=bus(nbytes, kern))
                if not args.no_nccl and variant == variants[0]:
                    ref = timeit(lambda: RefDistrib.sync_model(model), max(5, args.iters // 2))
                    flat = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
                    one = timeit(lambda: dist.all_reduce(flat), args.iters)
                    row.update(reference_path_nccl_ms=ref, nccl_flat_ms=one, nccl_flat_bus_gbs=bus(nbytes, one))
                emit(**row)
                del model, params
                if args.variants:
                    drop_plans()
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    eng = fctx.current().engine
    emit(kind="info", native_launches=eng.native_launches(), mem_kind=int(eng.info.mem_kind), world=world,
         max_blocks=int(eng.info.max_blocks), nvls=bool(eng.multicast), nvls_error=eng.multicast_error,
         tag=args.tag, env={k: v for k, v in os.environ.items() if k.startswith("FLASHY_B200")})
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
