#!/usr/bin/env python
"""Benchmark of the hot path on BASELINE.json's metric: CIFAR ResNet-18 samples/sec with
``distrib.sync_model`` gradient synchronisation (reference: examples/cifar/solver.py:46-53).

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA path)
    python bench.py --impl reference --gpus N --steps K ...  # reference path on the host CPUs

Workload (``config.workload``): the 8-rank data-parallel job of BASELINE configs[1] --
torchvision resnet18(num_classes=10) in bf16, batch 64 per rank, SGD lr 1e-4, synthetic
CIFAR-shaped batches -- run on N GPUs with 8/N ranks per GPU.  On one GPU the 8 ranks are
virtual ranks of one process (one thread + one CUDA stream each); every ``sync_model`` of the
8 ranks is ONE launch of the bucketed all-reduce kernel.  Total work per step is fixed
(global batch 512), so the scaling over N is "strong".  A step per rank is:
    forward, cross_entropy, backward (replayed CUDA graph), flashy_b200.distrib.sync_model(model),
    optim.step()                    [zero_grad: the replayed backward overwrites .grad]
``value`` times K steps with the batch resident in HBM; ``e2e`` times K more steps through the
same public API with the batch copied from pinned host memory and the loss read back each step.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

WORLD = 8
BATCH = 64
METRIC = "cifar_resnet18_train_samples_per_sec"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=("native", "reference"))
    ap.add_argument("--world", type=int, default=WORLD, help="data-parallel ranks of the job")
    ap.add_argument("--batch", type=int, default=BATCH, help="samples per rank per step")
    ap.add_argument("--no-graphs", action="store_true", help="eager forward/backward instead of a CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-steps", type=int, default=2)
    ap.add_argument("--kernel-table", default="", help="write a per-kernel device-time table of the timed region (CUPTI) to this file")
    return ap.parse_args()


# =========================================================================== reference arm
def reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cpu_train
    res = cpu_train.run(world=args.world, batch=args.batch, steps=max(1, args.steps), warmup=max(0, args.warmup))
    line = {
        "impl": "reference", "metric": METRIC, "value": res["value"], "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": workload_config(args, args.gpus, cpu=True),
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": res["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, n_gpus: int, cpu: bool = False) -> dict:
    return {
        "workload": ("examples/cifar ResNet-18 (torchvision resnet18, 10 classes), distrib.sync_model gradient+buffer "
                     f"all-reduce, {args.world} data-parallel ranks x batch {args.batch}, SGD lr 1e-4"),
        "world": args.world, "ranks_per_gpu": None if cpu else args.world // n_gpus,
        "global_batch": args.world * args.batch, "image": "3x32x32",
        "grad_tensors": 62, "grad_elements": 11181642,
    }


# =========================================================================== clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            pass

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for row in out.strip().splitlines():
            cells = [c.strip() for c in row.split(",")]
            if len(cells) < 7:
                continue
            try:
                sm.append(float(cells[0]))
                mx.append(float(cells[1]))
            except ValueError:
                continue
            for name, cell in zip(names, cells[3:7]):
                if cell.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# =========================================================================== native arm
class Replica:
    """One data-parallel rank: model, optimizer, static batch, captured forward+backward."""

    def __init__(self, rank: int, args, device):
        import torch
        import torch.nn.functional as F
        import torchvision
        self.rank = rank
        self.stream = torch.cuda.Stream(device=device)
        torch.manual_seed(1234)                                       # same initial weights everywhere
        self.model = torchvision.models.resnet18(num_classes=10).to(device=device, dtype=torch.bfloat16)
        self.model = self.model.to(memory_format=torch.channels_last)
        self.optim = torch.optim.SGD(self.model.parameters(), lr=1e-4)
        g = torch.Generator().manual_seed(1234 + rank)
        n_host = 4                                                    # rotating pinned batches for the e2e leg
        self.host_img = [torch.randn(args.batch, 3, 32, 32, generator=g).to(torch.bfloat16).pin_memory() for _ in range(n_host)]
        self.host_lab = [torch.randint(0, 10, (args.batch,), generator=g).pin_memory() for _ in range(n_host)]
        self.img = self.host_img[0].to(device).contiguous(memory_format=torch.channels_last)
        self.label = self.host_lab[0].to(device)
        self.loss = torch.zeros((), device=device, dtype=torch.bfloat16)
        self.graph = None
        self.h2d_bytes = self.host_img[0].numel() * 2 + self.host_lab[0].numel() * 8
        self.d2h_bytes = 2
        self.F = F
        if not args.no_graphs:
            self._capture()

    def _fwd_bwd(self):
        loss = self.F.cross_entropy(self.model(self.img), self.label)
        loss.backward()
        return loss

    def _capture(self):
        import torch
        with torch.cuda.stream(self.stream):
            for _ in range(3):                                        # warm-up on the capture stream
                self.optim.zero_grad(set_to_none=True)
                self._fwd_bwd()
        self.stream.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        self.optim.zero_grad(set_to_none=True)
        with torch.cuda.graph(self.graph, stream=self.stream):
            loss = self._fwd_bwd()
            self.loss.copy_(loss.detach())
        self.stream.synchronize()

    def step(self, distrib, e2e: bool, it: int):
        if e2e:
            k = it % len(self.host_img)
            self.img.copy_(self.host_img[k], non_blocking=True)
            self.label.copy_(self.host_lab[k], non_blocking=True)
        if self.graph is not None:
            self.graph.replay()
        else:
            self.loss.copy_(self._fwd_bwd().detach())
        distrib.sync_model(self.model)
        self.optim.step()
        if self.graph is None:
            self.optim.zero_grad()
        if e2e:
            return self.loss.item()                                   # device -> host read of the step's result
        return None


def native_arm(args) -> None:
    import torch
    import torch.distributed as dist
    from flashy_b200 import VirtualWorld, distrib
    from flashy_b200 import context as fctx

    n_gpus = args.gpus
    proc_world = int(os.environ.get("WORLD_SIZE", "1"))
    proc_rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert proc_world == n_gpus, f"--gpus {n_gpus} needs {n_gpus} processes (torchrun), got WORLD_SIZE={proc_world}"
    assert args.world % n_gpus == 0
    n_local = args.world // n_gpus
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if proc_world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("gloo", init_method="env://")        # bootstrap + timing reduction only

    vw = VirtualWorld(n_local, device=local_rank) if n_local > 1 else None
    replicas = [Replica(proc_rank * n_local + l, args, device) for l in range(n_local)]
    torch.cuda.synchronize()

    def run_ranks(fn):
        if vw is not None:
            return vw.run(fn)
        return [fn(proc_rank, args.world)]

    def engine():
        return vw.engine if vw is not None else fctx.current().engine

    def global_barrier():
        torch.cuda.synchronize()
        if proc_world > 1:
            dist.barrier()

    def timed_region(steps: int, e2e: bool):
        """Every rank runs `steps` steps; returns max-over-ranks device time in ms."""
        def body(rank, world):
            rep = replicas[rank - proc_rank * n_local]
            with torch.cuda.stream(rep.stream):
                distrib.barrier()                                     # all ranks (threads and processes) start together
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(rep.stream)
                for it in range(steps):
                    rep.step(distrib, e2e, it)
                e1.record(rep.stream)
                rep.stream.synchronize()
                return e0.elapsed_time(e1)
        global_barrier()
        ms = max(run_ranks(body))
        global_barrier()
        if proc_world > 1:
            t = torch.tensor([ms], dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
        return ms

    # ---- warm-up (also creates the communicator and the bucket plans)
    timed_region(max(args.warmup, 3), e2e=False)
    timed_region(2, e2e=True)
    eng = engine()

    # ---- timed: K steps, inputs resident
    sampler = ClockSampler(local_rank) if proc_rank == 0 else None
    eng.profile, eng.timings = True, []
    launches0 = eng.native_launches()
    cuprof = os.environ.get("FX_BENCH_CUPROF") == "1"                 # ncu --profile-from-start off
    if cuprof:
        torch.cuda.profiler.start()
    prof = None
    if args.kernel_table:
        from torch.profiler import profile, ProfilerActivity
        prof = profile(activities=[ProfilerActivity.CUDA])
        prof.__enter__()
    ms_value = timed_region(args.steps, e2e=False)
    if prof is not None:
        torch.cuda.synchronize()
        prof.__exit__(None, None, None)
        if proc_rank == 0:
            events = [e for e in prof.key_averages() if e.device_time_total > 0]
            total = sum(e.device_time_total for e in events)
            with open(args.kernel_table, "w") as fh:
                fh.write(f"# kernels of {args.steps} timed steps (torch.profiler / CUPTI), sorted by device time; total {total:.0f} us\n")
                fh.write("share_pct,device_us,count,name\n")
                for e in sorted(events, key=lambda e: -e.device_time_total):
                    fh.write(f"{100 * e.device_time_total / total:.2f},{e.device_time_total:.1f},{e.count},{e.key[:140]}\n")
    if cuprof:
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    launches = eng.native_launches() - launches0
    eng.profile = False
    timings = list(eng.timings)
    # ---- timed: K steps end to end (H2D batch + D2H loss inside the region)
    ms_e2e = timed_region(args.steps, e2e=True)
    # The timed regions are short (tens of ms at 8 GPUs): keep the same load running until the
    # sampler has had ~0.6 s, so that the median SM clock under load rests on enough samples.
    # Collective: every rank runs the same number of extra (untimed) steps.
    t_load = ms_value + ms_e2e
    extra = 0
    while t_load < 600.0 and extra < 40:
        t_load += timed_region(args.steps, e2e=False)
        extra += 1
    clocks = sampler.stop() if sampler else None
    if clocks is not None:
        clocks["sampled_over_ms"] = t_load
    # ---- auxiliary: one rank alone on this GPU, no sync_model (the W = 1 step of the reference,
    # where the path is a no-op): what the step costs without any gradient exchange
    rep0 = replicas[0]
    torch.cuda.synchronize()
    with torch.cuda.stream(rep0.stream):
        for _ in range(3):
            rep0.graph.replay() if rep0.graph is not None else rep0._fwd_bwd()
            rep0.optim.step()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(rep0.stream)
        for _ in range(args.steps):
            rep0.graph.replay() if rep0.graph is not None else rep0._fwd_bwd()
            rep0.optim.step()
            if rep0.graph is None:
                rep0.optim.zero_grad()
        s1.record(rep0.stream)
        rep0.stream.synchronize()
    single_ms = s0.elapsed_time(s1) / args.steps

    if proc_world > 1:
        t = torch.tensor([launches], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        launches = int(t[0])

    # ---- the dominant kernel of this repo: the gradient-bucket all-reduce
    by_plan = {}
    for key, e0, e1 in timings:
        by_plan.setdefault(key, []).append(e0.elapsed_time(e1))
    grad_key = max(by_plan, key=lambda k: sum(k[1])) if by_plan else None
    roofline = None
    allreduce = None
    if grad_key is not None:
        kernel_ms = statistics.mean(by_plan[grad_key])
        payload = sum(grad_key[1]) * 2                                # bf16 bytes of one rank's bucket (N)
        W = args.world
        peaks = {}
        try:
            peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
        except OSError:
            pass
        if n_local == W:
            # all ranks on this GPU: every byte moves through HBM.  Per rank: pack 2N, reduce reads N
            # and writes N/W, gather reads N and writes N  ->  (5 + 1/W) N, times the W hosted ranks.
            alg_bytes = W * (5 + 1 / W) * payload
            peak, bound, peak_src = peaks.get("hbm_gbs", 6650.0), "hbm", ("measured" if peaks else "fallback")
        else:
            # NVLink bytes per GPU per direction: every hosted rank pulls its shard from the
            # (W - n_local) remote arenas, then the (W - n_local) remote reduced shards.
            alg_bytes = 2 * n_local * (W - n_local) / W * payload
            peak, bound, peak_src = 900.0, "nvlink", "nominal NVLink 5 per direction (measured peer copy: 770)"
        achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
        traffic = None
        try:
            traffic = json.loads((ROOT / "profiles" / "traffic.json").read_text()).get(f"n{n_gpus}")
        except (OSError, ValueError):
            pass
        roofline = {"bound": bound, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "kernel": "k_two_shot<bf16> (bucketed all-reduce of the 62 gradient tensors)",
                    "kernel_ms": kernel_ms, "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                    "launches_timed": len(by_plan[grad_key])}
        allreduce = {"payload_bytes_per_rank": payload, "kernel_ms": kernel_ms,
                     "alg_gbs": payload / (kernel_ms * 1e-3) / 1e9,
                     "bus_gbs": 2 * (W - 1) / W * payload / (kernel_ms * 1e-3) / 1e9}

    if proc_rank != 0:
        return
    samples = args.world * args.batch * args.steps
    line = {
        "metric": METRIC, "value": samples / (ms_value * 1e-3), "unit": "samples/s", "n_gpus": n_gpus,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_value / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": dict(workload_config(args, n_gpus),
                       cuda_graphs=not args.no_graphs,
                       zero_grad=("implicit: the captured backward starts from grad=None, so every replay overwrites "
                                  ".grad (same state as zero_grad(set_to_none=True) + backward)" if not args.no_graphs
                                  else "optim.zero_grad() every step"),
                       l2="not flushed: a step touches weights+grads+activations of every hosted replica "
                          "(> 126 MB L2 at 8 ranks/GPU); the all-reduce kernel streams its bucket once"),
        "e2e": {"value": samples / (ms_e2e * 1e-3), "unit": "samples/s", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": replicas[0].h2d_bytes * args.world,
                "d2h_bytes_per_step": replicas[0].d2h_bytes * args.world,
                "api": "flashy_b200.distrib.sync_model(model) per rank; pinned-host batch -> device, loss.item()"},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": roofline,
        "allreduce": allreduce,
        "aux": {"single_rank_no_sync_ms_per_step": single_ms,
                "single_rank_no_sync_samples_per_s": args.batch / (single_ms * 1e-3),
                "note": "one replica alone on one GPU without sync_model (the reference's W=1 step); "
                        "at N=8 (one rank per GPU) ms_per_step minus this is the exposed gradient-sync cost"},
    }
    if n_gpus == 1 and not args.no_cpu_baseline:
        from oracle import cpu_train
        res = cpu_train.run(world=args.world, batch=args.batch, steps=args.cpu_steps, warmup=1)
        line["cpu_baseline"] = {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")}
    print(json.dumps(line), flush=True)
    if vw is not None:
        vw.close()


def main():
    args = parse()
    if args.impl == "reference":
        reference_arm(args)
    else:
        native_arm(args)
    try:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()
    except Exception:      # noqa: BLE001
        pass


if __name__ == "__main__":
    main()
