"""Rank identity and local rendezvous.

Production layout: one process per GPU, one rank per process (``ProcessContext``), the same
layout Dora / torchrun give the reference.  ``VirtualWorld`` additionally lets ONE process
host several *virtual* ranks on one GPU, each running on its own Python thread, so that the
multi-rank kernels can be exercised, profiled and parity-checked on a single-GPU box: the
unchanged ``flashy_b200.distrib`` calls made by the W threads meet in a local rendezvous and
leave as ONE launch whose ``gridDim.y`` spans the hosted ranks.
"""
from __future__ import annotations

import threading
import typing as tp

import torch

from .engine import Engine

_tls = threading.local()
_process_ctx: tp.Optional["ProcessContext"] = None
_process_lock = threading.Lock()


def _dist_state() -> tp.Tuple[int, int]:
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


class BaseContext:
    n_local = 1
    local = 0

    @property
    def engine(self) -> Engine:
        raise NotImplementedError

    @property
    def rank(self) -> int:
        raise NotImplementedError

    @property
    def world(self) -> int:
        raise NotImplementedError

    def rendezvous(self, payload, leader_fn):
        """Every hosted rank deposits ``payload``; ``leader_fn(list_of_payloads)`` runs once;
        its result (or exception) is handed to every hosted rank."""
        raise NotImplementedError


class ProcessContext(BaseContext):
    """The calling process is exactly one rank (``torch.distributed`` numbering)."""

    def __init__(self):
        self._engine: tp.Optional[Engine] = None
        self._lock = threading.Lock()

    @property
    def rank(self) -> int:
        return _dist_state()[0]

    @property
    def world(self) -> int:
        return _dist_state()[1]

    def engine_for(self, device: tp.Optional[int], host_only: bool = False) -> Engine:
        with self._lock:
            if self._engine is None:
                rank, world = _dist_state()
                self._engine = Engine(1, device, rank, world, host_only=host_only)
            return self._engine

    @property
    def engine(self) -> Engine:
        return self.engine_for(None)

    def cached_engine(self) -> tp.Optional[Engine]:
        return self._engine

    def rendezvous(self, payload, leader_fn):
        return leader_fn([payload])

    def reset(self) -> None:
        with self._lock:
            if self._engine is not None:
                self._engine.close()
                self._engine = None


class VirtualContext(BaseContext):
    def __init__(self, vworld: "VirtualWorld", local: int):
        self.vworld = vworld
        self.local = local
        self.n_local = vworld.n_local

    @property
    def engine(self) -> Engine:
        return self.vworld.engine

    def engine_for(self, device, host_only: bool = False) -> Engine:
        return self.vworld.engine

    def cached_engine(self) -> tp.Optional[Engine]:
        return self.vworld.engine

    @property
    def rank(self) -> int:
        return self.vworld.engine.rank0 + self.local

    @property
    def world(self) -> int:
        return self.vworld.engine.world

    def rendezvous(self, payload, leader_fn):
        return self.vworld.rendezvous(self.local, payload, leader_fn)


class VirtualWorld:
    """``n_local`` virtual ranks on one device of this process (threads + one communicator)."""

    def __init__(self, n_local: int, device: tp.Optional[int] = None, arena_mb: tp.Optional[int] = None,
                 timeout: float = 120.0):
        rank, world = _dist_state()
        self.n_local = n_local
        self.engine = Engine(n_local, device, rank, world, arena_mb=arena_mb)
        self.timeout = timeout
        self._barrier = threading.Barrier(n_local)
        self._slots: tp.List[tp.Any] = [None] * n_local
        self._result: tp.Any = None
        self._error: tp.Optional[BaseException] = None

    @property
    def world(self) -> int:
        return self.engine.world

    def rendezvous(self, local: int, payload, leader_fn):
        self._slots[local] = payload
        self._barrier.wait(self.timeout)
        if local == 0:
            try:
                self._result, self._error = leader_fn(list(self._slots)), None
            except BaseException as err:      # noqa: BLE001 - re-raised on every hosted rank
                self._result, self._error = None, err
        self._barrier.wait(self.timeout)
        if self._error is not None:
            raise self._error
        return self._result

    def run(self, fn: tp.Callable, *args, **kwargs) -> tp.List[tp.Any]:
        """Call ``fn(rank, world, *args, **kwargs)`` on one thread per hosted rank."""
        results: tp.List[tp.Any] = [None] * self.n_local
        errors: tp.List[tp.Optional[BaseException]] = [None] * self.n_local
        self._barrier.reset()
        device = self.engine.device

        def body(local: int):
            _tls.ctx = VirtualContext(self, local)
            try:
                with torch.cuda.device(device):
                    results[local] = fn(self.engine.rank0 + local, self.engine.world, *args, **kwargs)
            except BaseException as err:      # noqa: BLE001
                errors[local] = err
                self._barrier.abort()
                self.engine.abort()           # wake hosted ranks blocked in the native host fabric
            finally:
                _tls.ctx = None

        threads = [threading.Thread(target=body, args=(l,), name=f"vrank{l}") for l in range(self.n_local)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        from ._native import NativeError
        real = [e for e in errors if e is not None and not isinstance(e, (threading.BrokenBarrierError, NativeError))]
        if real:
            raise real[0]
        broken = [e for e in errors if e is not None]
        if broken:
            raise broken[0]
        return results

    def close(self) -> None:
        self.engine.close()


def current() -> BaseContext:
    ctx = getattr(_tls, "ctx", None)
    if ctx is not None:
        return ctx
    global _process_ctx
    with _process_lock:
        if _process_ctx is None:
            _process_ctx = ProcessContext()
        return _process_ctx


import atexit  # noqa: E402

atexit.register(lambda: reset_process_context())


def reset_process_context() -> None:
    global _process_ctx
    with _process_lock:
        if _process_ctx is not None:
            _process_ctx.reset()
            _process_ctx = None
