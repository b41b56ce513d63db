"""Blackwell-native drop-in for ``flashy.distrib`` (reference: ``flashy/distrib.py``).

Same names, signatures, in-place semantics and error behaviour as the reference module; the
body is new.  Where the reference issues one ``torch.distributed`` collective and one divide
per tensor plus two host-synchronising count checks per ``sync_model``
(``flashy/distrib.py:78-111``), this module packs each tensor list into a bucket and issues
ONE kernel of ``libflashy_b200.so`` that reads the peers' staging arenas directly over
NVLink and fuses the ``/ world_size``.  ``torch.distributed`` is used for bootstrap
(exchanging memory handles once) and for the two non-tensor utilities ``broadcast_object``
and ``wrap``; it is never on the per-step path.  There is no CPU / gloo fallback for tensor
data: CPU tensors in a distributed collective raise.

Environment knobs (none of them changes a public signature):
``FLASHY_B200_ARENA_MB`` (512), ``FLASHY_B200_BUCKET_MB`` (64), ``FLASHY_B200_EAGER_BUCKET_MB`` (8),
``FLASHY_B200_ONE_SHOT_MAX`` (bytes, 262144), ``FLASHY_B200_SLICE_BYTES`` (8192),
``FLASHY_B200_WIRE=bf16`` (send fp32 gradients as bf16: opt-in, lossy),
``FLASHY_B200_CHECK=always|plan`` (count check every call, or only when a bucket plan is new).
"""
from __future__ import annotations

import os
import pickle
import threading
import typing as tp
from contextlib import contextmanager
from functools import wraps

import torch
from torch import distributed
from torch.nn.parallel.distributed import DistributedDataParallel
from torch.utils.data import DataLoader, Subset
from torch.utils.data.distributed import DistributedSampler

from . import _native as N
from . import context as _context
from .engine import Engine, _DTYPES, _ESIZE, _dense

__all__ = [
    "rank", "world_size", "init", "rank_zero_only", "is_rank_zero", "is_distributed", "all_reduce",
    "average_metrics", "wrap", "average_tensors", "broadcast_tensors", "broadcast_model",
    "sync_gradients", "eager_sync_gradients", "sync_model", "eager_sync_model", "loader",
    "broadcast_object", "barrier", "sync_buffers",
]


# ------------------------------------------------------------------------------------------
# rank helpers (reference: re-exported from dora.distrib at flashy/distrib.py:21, and :24-42)
# ------------------------------------------------------------------------------------------

def rank() -> int:
    return _context.current().rank


def world_size() -> int:
    return _context.current().world


def init(backend: str = "nccl") -> None:
    """Same contract as ``dora.distrib.init``: no-op if already initialised or single process;
    otherwise ``env://`` rendezvous from RANK / WORLD_SIZE / LOCAL_RANK, one device per rank."""
    if distributed.is_initialized():
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1:
        return
    proc_rank = int(os.environ["RANK"])
    local = int(os.environ.get("LOCAL_RANK", proc_rank))
    if torch.cuda.is_available():
        torch.cuda.set_device(local)
    else:
        assert backend != "nccl", "the nccl backend needs CUDA"
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    distributed.init_process_group(backend=backend, init_method="env://", world_size=world, rank=proc_rank)


def rank_zero_only(fn: tp.Callable) -> tp.Callable:
    """Decorator: run ``fn`` on rank 0 only; other ranks get ``None`` (flashy/distrib.py:24-34)."""
    @wraps(fn)
    def guarded(*args: tp.Any, **kwargs: tp.Any) -> tp.Optional[tp.Any]:
        return fn(*args, **kwargs) if is_rank_zero() else None
    return guarded


def is_rank_zero() -> bool:
    return rank() == 0


def is_distributed() -> bool:
    return world_size() > 1


# ------------------------------------------------------------------------------------------
# internals
# ------------------------------------------------------------------------------------------

class _Item(tp.NamedTuple):
    src: int        # address read by the collective
    dst: int        # address written (== src for in-place)
    numel: int      # in fx-dtype units


def _is_complex_or_float(tensor: torch.Tensor) -> bool:
    return torch.is_floating_point(tensor) or torch.is_complex(tensor)      # flashy/distrib.py:92-93


def _engine(ctx, tensors: tp.Sequence[torch.Tensor]) -> Engine:
    """The communicator of the calling rank, created (collectively) on first use."""
    cuda = [t for t in tensors if t.is_cuda]
    if not N.cuda_available():
        return ctx.engine_for(None, host_only=True)      # rendezvous fabric only; data calls will raise
    return ctx.engine_for(cuda[0].device.index if cuda else torch.cuda.current_device())


def _flat(t: torch.Tensor, device: int) -> tp.Tuple[int, int, int]:
    """(fx dtype, address, numel in fx units) of a dense CUDA tensor on this rank's device."""
    if not t.is_cuda:
        raise RuntimeError(
            "flashy_b200 moves tensor data only with its CUDA kernels: a CPU tensor was passed to a "
            "distributed collective and there is no gloo/CPU fallback on this path.")
    if t.device.index != device:
        raise RuntimeError(f"tensor is on cuda:{t.device.index} but this rank's communicator is on cuda:{device}")
    if t.dtype not in _DTYPES:
        raise RuntimeError(f"dtype {t.dtype} is not supported by the flashy_b200 collectives")
    if not _dense(t):
        raise ValueError("Tensors must be contiguous")          # c10d raises the same for NCCL
    fx, mult = _DTYPES[t.dtype]
    return fx, t.data_ptr(), t.numel() * mult


def _signature(tensors: tp.Sequence[torch.Tensor]) -> int:
    # ints and bools only: their hashes do not depend on PYTHONHASHSEED, so every process agrees
    return hash(tuple((t.element_size(), t.is_complex(), t.is_floating_point(), t.numel())
                      for t in tensors)) & (2 ** 64 - 1)


def _check_number_of_params(params: tp.Sequence[torch.Tensor]) -> None:
    """Reference flashy/distrib.py:78-89, without its device all-reduce and ``.item()`` sync:
    the counts meet in host shared memory.  Raises on EVERY rank if any rank differs."""
    ctx = _context.current()
    if ctx.world == 1 or not params:
        return
    engine = _engine(ctx, params)
    total, same = engine.host_exchange(ctx.local, len(params), _sig_cached(params))
    if total != len(params) * ctx.world:
        raise RuntimeError(f"Mismatch in number of params: ours is {len(params)}, "
                           "at least one worker has a different one.")
    if not same:
        raise RuntimeError("Mismatch in the shapes/dtypes of the tensors to synchronise: "
                           "at least one worker passed a different list.")


def _sig_cached(params: tp.Sequence[torch.Tensor]) -> int:
    return _signature(params)


def _launch_streams(engine: Engine, payloads: tp.Sequence[tp.Any], launch: tp.Callable[[int], None]):
    """Run ``launch(stream_handle)`` ordered after every hosted rank's current stream.  Returns
    an event the ranks must wait on, or None when everything already sits on one stream."""
    streams = [p["stream"] for p in payloads]
    handles = {s.cuda_stream for s in streams}
    if len(handles) == 1:
        launch(streams[0])
        return None
    side = engine.side_stream
    seen = set()
    for s in streams:
        if s.cuda_stream in seen:
            continue
        seen.add(s.cuda_stream)
        ev = torch.cuda.Event()
        ev.record(s)
        side.wait_event(ev)
    launch(side)
    done = torch.cuda.Event()
    done.record(side)
    return done


def _collective(ctx, engine: Engine, kind: str, items: tp.Sequence[_Item], dtype: int, op: int = N.FX_AVG,
                src: int = 0) -> None:
    """One bucketed collective over ``items`` (all of fx dtype ``dtype``) for this rank."""
    wire = N.FX_BF16 if (engine.wire_bf16 and dtype == N.FX_F32 and kind == "ar" and op in (N.FX_AVG, N.FX_SUM)) else dtype
    esize = _ESIZE[dtype]
    for bucket in _split(engine, items, _ESIZE[wire], esize):
        numels = tuple(it.numel for it in bucket)
        plan = engine.get_plan(kind, numels, dtype, wire)
        payload = {"in": [it.src for it in bucket], "out": [it.dst for it in bucket],
                   "stream": torch.cuda.current_stream()}

        def lead(payloads, plan=plan):
            ins = [p["in"] for p in payloads]
            outs = [p["out"] for p in payloads]
            if kind == "ar":
                return _launch_streams(engine, payloads, lambda s: engine.allreduce(plan, op, ins, outs, s))
            return _launch_streams(engine, payloads, lambda s: engine.broadcast(plan, src, outs, s))

        done = ctx.rendezvous(payload, lead)
        if done is not None:
            torch.cuda.current_stream().wait_event(done)


def _split(engine: Engine, items: tp.Sequence[_Item], wire_size: int, esize: int) -> tp.List[tp.List[_Item]]:
    cap = max(engine.bucket_cap // wire_size, 64)
    cap -= cap % 64
    out: tp.List[tp.List[_Item]] = []
    cur: tp.List[_Item] = []
    used = 0
    for it in items:
        if it.numel > cap:                      # one tensor larger than a bucket: cut it
            if cur:
                out.append(cur)
                cur, used = [], 0
            done = 0
            while done < it.numel:
                n = min(cap, it.numel - done)
                out.append([_Item(it.src + done * esize, it.dst + done * esize, n)])
                done += n
            continue
        if cur and used + it.numel > cap:
            out.append(cur)
            cur, used = [], 0
        cur.append(it)
        used += it.numel
    if cur:
        out.append(cur)
    return out


def _reduce(ctx, ins: tp.Sequence[torch.Tensor], outs: tp.Sequence[torch.Tensor], op: int) -> None:
    """Bucketed all-reduce of ``ins`` into ``outs`` (same objects for in-place), grouped by dtype
    in first-appearance order (identical on every rank because the lists are)."""
    engine = _engine(ctx, ins)
    groups: tp.Dict[int, tp.List[_Item]] = {}
    for tin, tout in zip(ins, outs):
        fx, src, numel = _flat(tin, engine.device)
        if tout is tin:
            dst = src
        else:
            fx2, dst, numel2 = _flat(tout, engine.device)
            if fx2 != fx or numel2 != numel:
                raise RuntimeError("output tensor does not match the reduced tensor")
        if numel:
            groups.setdefault(fx, []).append(_Item(src, dst, numel))
    for fx, items in groups.items():
        _collective(ctx, engine, "ar", items, fx, op)


# ------------------------------------------------------------------------------------------
# collectives of the reference surface
# ------------------------------------------------------------------------------------------

_OPS = {
    distributed.ReduceOp.SUM: N.FX_SUM, distributed.ReduceOp.AVG: N.FX_AVG,
    distributed.ReduceOp.MAX: N.FX_MAX, distributed.ReduceOp.MIN: N.FX_MIN,
    distributed.ReduceOp.PRODUCT: N.FX_PROD,
}


def all_reduce(tensor: torch.Tensor, op=distributed.ReduceOp.SUM):
    """In-place all-reduce, no-op when not distributed (flashy/distrib.py:45-47)."""
    ctx = _context.current()
    if ctx.world == 1:
        return None
    fx_op = _OPS.get(op)
    if fx_op is None:
        raise RuntimeError(f"reduce op {op} is not supported by flashy_b200")
    _reduce(ctx, [tensor], [tensor], fx_op)
    return None


def average_metrics(metrics: tp.Dict[str, float], count=1.):
    """Weighted average of a metric dict over ranks (flashy/distrib.py:50-62)."""
    ctx = _context.current()
    if ctx.world == 1:
        return metrics
    keys, values = zip(*metrics.items())
    device = "cuda" if torch.cuda.is_available() else "cpu"
    tensor = torch.tensor(list(values) + [1], device=device, dtype=torch.float32)
    tensor *= count
    all_reduce(tensor)
    averaged = (tensor[:-1] / tensor[-1]).cpu().tolist()
    return dict(zip(keys, averaged))


def wrap(model):
    """DDP comparator, kept as in the reference (flashy/distrib.py:65-75)."""
    if is_distributed():
        return DistributedDataParallel(model, device_ids=[torch.cuda.current_device()],
                                       output_device=torch.cuda.current_device())
    return model


def average_tensors(tensors: tp.Iterable[torch.Tensor]) -> None:
    """In-place mean over ranks of every float/complex tensor (flashy/distrib.py:96-111);
    other dtypes are ignored.  One bucketed kernel instead of a collective per tensor."""
    ctx = _context.current()
    if ctx.world == 1:
        return
    todo = [t for t in tensors if _is_complex_or_float(t)]
    _check_number_of_params(todo)
    if todo:
        _reduce(ctx, [t.data for t in todo], [t.data for t in todo], N.FX_AVG)


def broadcast_tensors(tensors: tp.Iterable[torch.Tensor], src: int = 0) -> None:
    """Bit copy of rank ``src``'s float/complex tensors to every rank (flashy/distrib.py:114-127)."""
    ctx = _context.current()
    if ctx.world == 1:
        return
    todo = [t for t in tensors if _is_complex_or_float(t)]
    _check_number_of_params(todo)
    if not todo:
        return
    engine = _engine(ctx, todo)
    items = []
    for t in todo:
        fx, ptr, numel = _flat(t.data, engine.device)
        if numel:
            items.append(_Item(ptr, ptr, numel * _ESIZE[fx]))     # bytes
    if items:
        _collective(ctx, engine, "bc", items, N.FX_U8, src=src)


def broadcast_model(model: torch.nn.Module, src: int = 0) -> None:
    """Parameters then buffers from ``src`` (flashy/distrib.py:130-133)."""
    broadcast_tensors(model.parameters(), src)
    broadcast_tensors(model.buffers(), src)


def sync_gradients(params: tp.Iterable[torch.Tensor]) -> None:
    """Average the existing ``.grad`` of ``params`` over ranks (flashy/distrib.py:136-150)."""
    grads = [p.grad for p in params if p.grad is not None]
    average_tensors(grads)


def _sync_buffers(model: torch.nn.Module, sync_buffers: bool, average_buffers: bool) -> None:
    if not sync_buffers:
        return
    if average_buffers:
        average_tensors(model.buffers())
    else:
        broadcast_tensors(model.buffers())


def sync_buffers(model: torch.nn.Module, average: bool = True) -> None:
    """Convenience alias (an ADDITION: the reference has no such function, only the
    ``sync_buffers=`` / ``average_buffers=`` arguments of ``sync_model``)."""
    _sync_buffers(model, True, average)


def sync_model(model: torch.nn.Module, sync_buffers: bool = True, average_buffers: bool = True) -> None:
    """Call after ``backward()``: averages gradients and (by default) float buffers over ranks
    (flashy/distrib.py:193-210).  Returns once the work is enqueued on the current stream."""
    sync_gradients(model.parameters())
    _sync_buffers(model, sync_buffers, average_buffers)


# ------------------------------------------------------------------------------------------
# eager path: start reducing bucket k while backward is still producing bucket k+1
# ------------------------------------------------------------------------------------------

class _EagerBucket:
    def __init__(self, plan, n: int, n_local: int):
        self.plan = plan
        self.rows = [[0] * n for _ in range(n_local)]
        self.left = [n] * n_local
        self.events: tp.List[tp.Any] = [None] * n_local
        self.ready = 0
        self.launched = False
        self.done = None


class _EagerSession:
    """State shared by the hosted ranks of one ``eager_sync_gradients`` context."""

    def __init__(self, engine: Engine, n_local: int, specs):
        self.engine = engine
        self.n_local = n_local
        self.lock = threading.Lock()
        self.buckets = [_EagerBucket(plan, n, n_local) for plan, n in specs]
        self.next_launch = 0

    def arrive(self, local: int, k: int, j: int, ptr: int) -> None:
        """Called from the autograd thread: gradient j of bucket k of hosted rank ``local`` exists."""
        with self.lock:
            b = self.buckets[k]
            b.rows[local][j] = ptr
            b.left[local] -= 1
            if b.left[local] == 0:
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream())
                b.events[local] = ev
                b.ready += 1
            self._launch_ready()

    def _launch_ready(self) -> None:
        # strictly in bucket order, so every process issues the same collective sequence
        while self.next_launch < len(self.buckets):
            b = self.buckets[self.next_launch]
            if b.ready < self.n_local:
                return
            side = self.engine.side_stream
            for ev in b.events:
                side.wait_event(ev)
            self.engine.allreduce_begin(b.plan, N.FX_AVG, b.rows, side)
            b.done = torch.cuda.Event()
            b.done.record(side)
            b.launched = True
            self.next_launch += 1


@contextmanager
def eager_sync_gradients(params: tp.Iterable[torch.Tensor]):
    """Context manager: gradients are reduced as they become available during the (single)
    ``backward()`` run inside it (flashy/distrib.py:153-190).  Gradients are grouped into
    buckets in expected arrival order; a bucket's reduce-scatter/all-gather is launched on a
    side stream the moment its last gradient exists, and on exit the averaged values are
    written over ``param.grad`` (``torch.div(grad, W, out=param.grad)`` in the reference)."""
    ctx = _context.current()
    if ctx.world == 1:
        yield
        return
    params = [p for p in params if p.requires_grad]
    _check_number_of_params(params)
    if not params:
        yield
        return
    engine = _engine(ctx, params)
    cap = int(os.environ.get("FLASHY_B200_EAGER_BUCKET_MB", "8")) << 20

    # ---- bucket layout: reverse registration order ~ order in which backward yields grads
    order = list(range(len(params)))[::-1]
    layout: tp.List[tp.Tuple[int, tp.List[int]]] = []           # (fx dtype, [param index])
    for i in order:
        p = params[i]
        if p.dtype not in _DTYPES or not p.is_cuda:
            _flat(p, engine.device)                              # raises with the right message
        fx, mult = _DTYPES[p.dtype]
        size = p.numel() * mult * _ESIZE[fx]
        if layout and layout[-1][0] == fx and sum(params[q].numel() * _DTYPES[params[q].dtype][1] * _ESIZE[fx]
                                                   for q in layout[-1][1]) + size <= cap:
            layout[-1][1].append(i)
        else:
            layout.append((fx, [i]))
    where = {}
    for k, (_, idxs) in enumerate(layout):
        for j, i in enumerate(idxs):
            where[i] = (k, j)

    def make_session(_payloads):
        specs = []
        for fx, idxs in layout:
            numels = tuple(params[i].numel() * _DTYPES[params[i].dtype][1] for i in idxs)
            wire = N.FX_BF16 if (engine.wire_bf16 and fx == N.FX_F32) else fx
            specs.append((engine.get_plan("ar", numels, fx, wire, N.FX_ALGO_TWO_SHOT), len(idxs)))
        return _EagerSession(engine, ctx.n_local, specs)

    session: _EagerSession = ctx.rendezvous(None, make_session)
    local = ctx.local
    waiting = set(range(len(params)))
    fired: tp.Dict[int, torch.Tensor] = {}

    def _callback(i: int, grad: torch.Tensor):
        if i not in waiting:
            raise RuntimeError(f"We got a gradient twice for parameter {params[i]}.")
        data = grad.data
        if not _dense(data):
            data = data.contiguous()
        fired[i] = data                                         # keep the storage alive until exit
        waiting.remove(i)
        k, j = where[i]
        session.arrive(local, k, j, data.data_ptr())

    hooks = [p.register_hook(lambda g, i=i: _callback(i, g)) for i, p in enumerate(params)]
    try:
        yield
    finally:
        for hook in hooks:
            hook.remove()
        _check_number_of_params([params[i] for i in sorted(waiting)])   # same leftovers everywhere
        stream = torch.cuda.current_stream()
        for k, (fx, idxs) in enumerate(layout):
            bucket = session.buckets[k]
            complete = all(i in fired for i in idxs)
            if complete:
                def lead(payloads, bucket=bucket):
                    with session.lock:
                        session._launch_ready()
                    assert bucket.launched
                    outs = [p["out"] for p in payloads]
                    streams = {p["stream"].cuda_stream for p in payloads}
                    if len(streams) == 1:
                        payloads[0]["stream"].wait_event(bucket.done)
                        engine.allreduce_finish(bucket.plan, outs, payloads[0]["stream"])
                        return None
                    engine.allreduce_finish(bucket.plan, outs, engine.side_stream)
                    done = torch.cuda.Event()
                    done.record(engine.side_stream)
                    return done
                outs = []
                for i in idxs:
                    assert params[i].grad is not None
                    outs.append(_flat(params[i].grad.data, engine.device)[1])
                done = ctx.rendezvous({"out": outs, "stream": stream}, lead)
                if done is not None:
                    stream.wait_event(done)
            else:
                # Some gradients of this bucket never arrived (unused parameters): reduce the ones
                # that did with an ordinary fused launch.  The leftovers were checked to agree.
                got = [i for i in idxs if i in fired]

                def skip(_payloads, k=k):                        # keep the launch order identical
                    with session.lock:
                        if session.next_launch == k:
                            session.next_launch += 1
                ctx.rendezvous(None, skip)
                if got:
                    for i in got:
                        assert params[i].grad is not None
                    _reduce(ctx, [fired[i] for i in got], [params[i].grad.data for i in got], N.FX_AVG)
        fired.clear()


@contextmanager
def eager_sync_model(model: torch.nn.Module, sync_buffers: bool = True, average_buffers: bool = True):
    """``sync_model`` with the gradient part overlapped with backward (flashy/distrib.py:213-224)."""
    with eager_sync_gradients(model.parameters()):
        yield
    _sync_buffers(model, sync_buffers, average_buffers)


# ------------------------------------------------------------------------------------------
# data loading / objects / barrier  (stay Python: north_star)
# ------------------------------------------------------------------------------------------

def loader(dataset, *args, shuffle=False, klass=DataLoader, **kwargs):
    """Sharded dataloader (flashy/distrib.py:227-243): ``DistributedSampler`` when
    ``shuffle=True``, a strided ``Subset`` (no duplicated samples) otherwise."""
    ctx = _context.current()
    if ctx.world == 1:
        return klass(dataset, *args, shuffle=shuffle, **kwargs)
    if shuffle:
        sampler = DistributedSampler(dataset, num_replicas=ctx.world, rank=ctx.rank)
        return klass(dataset, *args, **kwargs, sampler=sampler)
    shard = Subset(dataset, list(range(ctx.rank, len(dataset), ctx.world)))
    return klass(shard, *args, shuffle=shuffle, **kwargs)


def broadcast_object(obj: tp.Any = None, src: int = 0, device=None):
    """Share a picklable object from rank ``src`` (flashy/distrib.py:246-269).  Every rank,
    the source included, returns the unpickled copy (as the reference effectively does)."""
    ctx = _context.current()
    if ctx.world == 1:
        return obj

    def lead(payloads):
        blob = None
        for r, p in payloads:
            if r == src:
                blob = pickle.dumps(p)
        if distributed.is_initialized() and distributed.get_world_size() > 1:
            box = [blob]
            distributed.broadcast_object_list(box, src=src // ctx.n_local)
            blob = box[0]
        return blob

    return pickle.loads(ctx.rendezvous((ctx.rank, obj), lead))


def barrier() -> None:
    """All ranks wait for each other on the host (flashy/distrib.py:272-276)."""
    ctx = _context.current()
    if ctx.world == 1:
        return
    engine = ctx.engine_for(torch.cuda.current_device() if N.cuda_available() else None,
                            host_only=not N.cuda_available())
    engine.host_barrier(ctx.local)
