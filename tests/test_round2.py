"""Round-2 behaviour: multi-host refusal, fused-kernel plan geometry (CPU); structural-edit cache
invalidation, multi-bucket eager plans, backward overlap, op / dtype coverage, the fast path of
``average_tensors`` and the lossless ``all_reduce`` under ``FLASHY_B200_WIRE=bf16`` (GPU, virtual
ranks).  The reference lines each case follows are cited at the case."""
import ctypes as C
import os

import pytest
import torch
from torch import nn

from tests.harness import run_ranks


# ------------------------------------------------------------------------------------------ CPU
def _multi_host_worker(rank, world):
    os.environ["FLASHY_B200_HOST_ID"] = f"node{rank}"               # pretend every rank sits on its own machine
    from flashy_b200 import distrib
    with pytest.raises(RuntimeError, match="span several hosts"):
        distrib.barrier()                                           # first use creates + connects the communicator


def test_world_spanning_hosts_is_refused_with_a_clear_error():
    """one NVSwitch domain only.  The reference works on any backend
    (flashy/distrib.py:45-47); here a multi-host world must fail at bootstrap with a message that says
    so, on every rank, instead of an unrelated fd-passing / shm_open error."""
    run_ranks(2, "tests.test_round2", "_multi_host_worker")


def _dry_plan(world, numels, dtype, algo=0):
    from flashy_b200 import _native as N
    arr = (C.c_int64 * len(numels))(*numels)
    handle = C.c_void_p()
    N.check(N.lib.fx_plan_create(None, world, arr, len(numels), dtype, dtype, algo, C.byref(handle)))
    info = N.PlanInfo()
    N.check(N.lib.fx_plan_get_info(handle, C.byref(info)))
    N.lib.fx_plan_destroy(handle)
    return info


@pytest.mark.parametrize("world", [2, 3, 4, 8, 16])
def test_fused_kernel_plan_geometry(world):
    """A float bucket reduced in its own dtype resolves to the fused TMA kernel; its chunk geometry fits
    the six staging buffers into 192 KiB of shared memory and covers the slice."""
    from flashy_b200 import _native as N
    for dtype, esize in ((N.FX_F32, 4), (N.FX_BF16, 2), (N.FX_F16, 2)):
        info = _dry_plan(world, [1 << 22, 77, 4097, 1 << 20], dtype)
        assert info.algo == N.FX_ALGO_TWO_SHOT and info.kernel == 6, (info.algo, info.kernel)   # k_fuse<NVLS=false>
        assert info.chunk_bytes % 128 == 0 and info.chunk_bytes > 0
        assert 6 * world * info.chunk_bytes <= 192 << 10
        slice_bytes = info.shard_elems // info.grid_x * esize
        assert (info.chunks - 1) * info.chunk_bytes < slice_bytes <= info.chunks * info.chunk_bytes
    small = _dry_plan(world, [100, 200], N.FX_F32)
    assert small.kernel == 1                                        # one-shot
    ints = _dry_plan(world, [1 << 22], N.FX_I64)
    assert ints.kernel == 2                                         # exact integers stay on the classic two-shot


# ------------------------------------------------------------------------------------------ GPU
def _oracle():
    from oracle import numeric
    return numeric


@pytest.mark.gpu
def test_sync_model_follows_a_replaced_head():
    """VERDICT weak #7: the reference walks the module every call (flashy/distrib.py:205-210).  Replace
    ``model.fc`` between two ``sync_model`` calls: the NEW parameters' gradients must be averaged on the
    very next call."""
    from flashy_b200 import VirtualWorld, distrib
    world = 4
    vw = VirtualWorld(world, device=0, arena_mb=64)
    try:
        def body(rank, w):
            model = nn.Sequential(nn.Linear(64, 512), nn.ReLU(), nn.Linear(512, 300)).cuda()
            model.add_module("fc", nn.Linear(300, 10).cuda())
            for p in model.parameters():
                p.grad = torch.full_like(p, float(rank + 1))
            distrib.sync_model(model)
            for p in model.parameters():
                assert torch.equal(p.grad, torch.full_like(p, (w + 1) / 2))
            model.fc = nn.Linear(300, 7).cuda()                     # structural edit
            for p in model.parameters():
                p.grad = torch.full_like(p, float(2 * rank))
            distrib.sync_model(model)
            torch.cuda.synchronize()
            got = [p.grad.cpu() for p in model.parameters()]
            return got
        res = vw.run(body)
    finally:
        vw.close()
    want = float(sum(2 * r for r in range(world))) / world
    for got in res:
        assert got[-1].shape == (7,) and got[-2].shape == (7, 300)
        for g in got:
            assert torch.equal(g, torch.full_like(g, want))


@pytest.mark.gpu
def test_eager_buckets_of_identical_shape_do_not_share_a_plan(monkeypatch):
    """ADVICE (high): with a small eager bucket cap, repeated layers become single-tensor buckets of the SAME
    shape; every bucket in flight needs its own staging regions (flashy/distrib.py:153-190 semantics)."""
    monkeypatch.setenv("FLASHY_B200_EAGER_BUCKET_MB", "1")
    from flashy_b200 import VirtualWorld, distrib
    numeric = _oracle()
    world = 4
    dims = 640                                                       # 640*640*4 B = 1.6 MB > the 1 MB cap

    torch.manual_seed(3)
    init = nn.Sequential(*[nn.Linear(dims, dims, bias=False) for _ in range(4)]).state_dict()

    def make():                                                      # (the global RNG is not thread-safe: load a fixed state)
        m = nn.Sequential(*[nn.Linear(dims, dims, bias=False) for _ in range(4)])
        m.load_state_dict(init)
        return m
    gens = [torch.Generator().manual_seed(40 + r) for r in range(world)]
    xs = [torch.randn(8, dims, generator=g) for g in gens]
    grads = []
    for r in range(world):
        m = make().cuda()
        m(xs[r].cuda()).square().mean().backward()
        grads.append([p.grad.detach().cpu() for p in m.parameters()])
    want = numeric.average_tensors(grads)[0]
    vw = VirtualWorld(world, device=0, arena_mb=128)
    try:
        def body(rank, w):
            m = make().cuda()
            for _ in range(2):                                       # twice: plans and staging parity are reused
                m.zero_grad()
                with distrib.eager_sync_model(m):
                    m(xs[rank].cuda()).square().mean().backward()
            torch.cuda.synchronize()
            return [p.grad.detach().cpu() for p in m.parameters()]
        res = vw.run(body)
    finally:
        vw.close()
    for got in res:
        for g, w_ in zip(got, want):
            assert torch.equal(g, w_)


@pytest.mark.gpu
def test_reduce_ops_and_integer_dtypes():
    """flashy/distrib.py:45-47 accepts any ReduceOp: PRODUCT / AVG / MAX / MIN, int32 and int64."""
    import torch.distributed as dist
    from flashy_b200 import VirtualWorld, distrib
    world = 4
    vw = VirtualWorld(world, device=0, arena_mb=64)
    try:
        def body(rank, w):
            out = {}
            x = torch.full((1000,), float(rank + 1), device="cuda")
            distrib.all_reduce(x, dist.ReduceOp.PRODUCT)
            out["prod"] = x.cpu()
            x = torch.arange(300000, device="cuda", dtype=torch.float32) + rank
            distrib.all_reduce(x, dist.ReduceOp.AVG)
            out["avg"] = x.cpu()
            x = torch.full((5,), rank + 1, device="cuda", dtype=torch.int32)
            distrib.all_reduce(x)
            out["i32"] = x.cpu()
            x = (torch.arange(400000, device="cuda", dtype=torch.int32) % 97) * (rank + 1)
            distrib.all_reduce(x, dist.ReduceOp.MAX)
            out["i32max"] = x.cpu()
            x = torch.full((3,), 2 ** 40 + rank, device="cuda", dtype=torch.int64)
            distrib.all_reduce(x, dist.ReduceOp.MIN)
            out["i64min"] = x.cpu()
            return out
        res = vw.run(body)
    finally:
        vw.close()
    for out in res:
        assert torch.equal(out["prod"], torch.full((1000,), 24.0))
        assert torch.equal(out["avg"], torch.arange(300000, dtype=torch.float32) + 1.5)
        assert out["i32"].tolist() == [10] * 5 and out["i32"].dtype == torch.int32
        assert torch.equal(out["i32max"], (torch.arange(400000, dtype=torch.int32) % 97) * world)
        assert out["i64min"].tolist() == [2 ** 40] * 3


@pytest.mark.gpu
def test_average_tensors_fast_path_revalidates_shapes():
    """VERDICT weak #8: repeat calls skip the key building, but a list of the same LENGTH with different
    element counts or dtypes must not reuse the cached layout."""
    from flashy_b200 import VirtualWorld, distrib
    world = 2
    vw = VirtualWorld(world, device=0, arena_mb=64)
    try:
        def body(rank, w):
            for rep in range(3):
                a = [torch.full((1000,), float(rank), device="cuda"), torch.full((70000,), float(rank), device="cuda")]
                distrib.average_tensors(a)
                assert all(torch.equal(t, torch.full_like(t, 0.5)) for t in a)
            b = [torch.full((1000,), float(rank), device="cuda"), torch.full((70001,), float(rank), device="cuda")]
            distrib.average_tensors(b)                              # same length, same first numel
            assert all(torch.equal(t, torch.full_like(t, 0.5)) for t in b)
            c = [torch.full((1000,), float(rank), device="cuda", dtype=torch.float64),
                 torch.full((70001,), float(rank), device="cuda", dtype=torch.float64)]
            distrib.average_tensors(c)                              # same numels, other dtype
            assert all(torch.equal(t, torch.full_like(t, 0.5)) for t in c)
            params = [nn.Parameter(torch.zeros(3000, device="cuda")) for _ in range(5)]
            for p in params:
                p.grad = torch.full_like(p, float(rank + 1))
            distrib.sync_gradients(params)                          # flashy/distrib.py:136-150
            assert all(torch.equal(p.grad, torch.full_like(p, 1.5)) for p in params)
            torch.cuda.synchronize()
            return True
        assert all(vw.run(body))
    finally:
        vw.close()


@pytest.mark.gpu
def test_bf16_wire_never_touches_all_reduce(monkeypatch):
    """ADVICE (low): ``FLASHY_B200_WIRE=bf16`` is for averaged gradients only; ``all_reduce`` and therefore
    ``average_metrics`` (flashy/distrib.py:50-62) must stay exact for counts above 256."""
    monkeypatch.setenv("FLASHY_B200_WIRE", "bf16")
    from flashy_b200 import VirtualWorld, distrib
    world = 2
    vw = VirtualWorld(world, device=0, arena_mb=64)
    try:
        def body(rank, w):
            x = torch.full((100000,), 1001.0 + rank, device="cuda")
            distrib.all_reduce(x)
            assert torch.equal(x, torch.full_like(x, 2003.0))       # bf16 could not represent 2003
            m = distrib.average_metrics({"n": 1001.0 + rank}, count=1000 + rank)
            want = ((1001.0 * 1000) + (1002.0 * 1001)) / 2001
            assert abs(m["n"] - want) < 1e-3
            g = [torch.full((100000,), 1001.0 + rank, device="cuda")]
            distrib.average_tensors(g)                              # the lossy wire applies here (opt-in)
            assert abs(float(g[0][0]) - 1001.5) <= 4.0
            torch.cuda.synchronize()
            return True
        assert all(vw.run(body))
    finally:
        vw.close()


@pytest.mark.gpu
def test_fused_kernel_is_what_runs_and_handles_odd_addresses():
    """The sharded float bucket goes through k_fuse (plan info), including tensors whose address is not
    16-byte aligned and element counts that leave < 16-byte tails -- bit-exact against the oracle."""
    from flashy_b200 import VirtualWorld, distrib, _native as N
    numeric = _oracle()
    world = 4
    numels = [300001, 7, 65539, 1, 128, 999999, 33]
    gens = [torch.Generator().manual_seed(70 + r) for r in range(world)]
    for dtype in (torch.float32, torch.bfloat16):
        per_rank = [[(torch.randn(n, generator=gens[r]) * 1e-2).to(dtype) for n in numels] for r in range(world)]
        want = numeric.average_tensors(per_rank)[0]
        vw = VirtualWorld(world, device=0, arena_mb=64)
        try:
            def body(rank, w):
                ts = []
                for t in per_rank[rank]:
                    base = torch.empty(t.numel() + 8, dtype=dtype, device="cuda")
                    view = base[1:1 + t.numel()]                    # element offset 1: 2- or 4-byte aligned only
                    view.copy_(t)
                    ts.append(view)
                distrib.average_tensors(ts)
                torch.cuda.synchronize()
                return [t.cpu() for t in ts]
            res = vw.run(body)
            kernels = {int(p.info.kernel) for p in vw.engine.plans.values()}
            assert 6 in kernels, kernels                            # k_fuse<NVLS=false>
            assert N.KERNEL_NAMES[6].startswith("k_fuse")
        finally:
            vw.close()
        for got in res:
            for g, w_ in zip(got, want):
                assert torch.equal(g, w_)
