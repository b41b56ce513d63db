"""CPU baseline of the benchmark workload (TEST / BENCH INFRASTRUCTURE, never shipped).

Runs the reference's training step of ``examples/cifar`` (``/root/reference/examples/cifar/
solver.py:46-53``: forward, cross-entropy, backward, ``distrib.sync_model``, ``optim.step``,
``optim.zero_grad``) on the host cores: W processes over gloo, the gradient sync being the
reference's own call sequence as restated in ``oracle/refdistrib.py`` (one all-reduce and one
divide per tensor, two host-synchronising count checks).  Used by ``bench.py`` for the
``cpu_baseline`` object and for ``--impl reference``.
"""
from __future__ import annotations

import json
import multiprocessing as mp
import os
import socket
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, batch: int, steps: int, warmup: int, threads: int, queue):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    if str(ROOT) not in sys.path:
        sys.path.insert(0, str(ROOT))
    import torch
    import torch.distributed as dist
    import torch.nn.functional as F
    import torchvision
    from oracle.refdistrib import RefDistrib

    torch.set_num_threads(threads)
    dist.init_process_group("gloo", init_method="env://")
    torch.manual_seed(1234)                                       # identical weights on every rank
    model = torchvision.models.resnet18(num_classes=10)
    optim = torch.optim.SGD(model.parameters(), lr=1e-4)
    g = torch.Generator().manual_seed(1234 + rank)                # per-rank data
    img = torch.randn(batch, 3, 32, 32, generator=g)
    label = torch.randint(0, 10, (batch,), generator=g)
    sync_s = 0.0
    t0 = 0.0
    for step in range(warmup + steps):
        if step == warmup:
            dist.barrier()
            t0 = time.perf_counter()
            sync_s = 0.0
        loss = F.cross_entropy(model(img), label)
        loss.backward()
        s0 = time.perf_counter()
        RefDistrib.sync_model(model)
        sync_s += time.perf_counter() - s0
        optim.step()
        optim.zero_grad()
        loss.item()
    dist.barrier()
    elapsed = time.perf_counter() - t0
    stats = torch.tensor([elapsed, sync_s], dtype=torch.float64)
    dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    if rank == 0:
        queue.put({"elapsed_s": float(stats[0]), "sync_s": float(stats[1])})
    dist.destroy_process_group()


def run(world: int = 8, batch: int = 8, steps: int = 2, warmup: int = 1, cores: int | None = None) -> dict:
    """Returns samples/s of the whole W-rank CPU job plus how it was obtained."""
    cores = cores or os.cpu_count() or 1
    threads = max(1, cores // world)
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, batch, steps, warmup, threads, queue))
             for r in range(world)]
    t0 = time.perf_counter()
    for p in procs:
        p.start()
    out = queue.get(timeout=1800)
    for p in procs:
        p.join()
        assert p.exitcode == 0
    wall = time.perf_counter() - t0
    samples = world * batch * steps
    return {
        "value": samples / out["elapsed_s"],
        "unit": "samples/s",
        "cores": min(cores, threads * world),
        "kind": "port",
        "sample": (f"{steps} timed steps (+{warmup} warm-up) of the ResNet-18/CIFAR step, world {world} gloo "
                   f"processes x batch {batch}, fp32 on CPU, {threads} thread(s) per rank; "
                   f"sync_model (oracle/refdistrib.py) took {1e3 * out['sync_s'] / steps:.0f} ms/step; "
                   f"{wall:.0f} s wall including process start-up"),
        "ms_per_step": 1e3 * out["elapsed_s"] / steps,
        "sync_ms_per_step": 1e3 * out["sync_s"] / steps,
        "world": world, "batch_per_rank": batch, "steps": steps, "warmup": warmup,
    }


if __name__ == "__main__":
    print(json.dumps(run()))
