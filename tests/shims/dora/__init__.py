"""Minimal stand-in for the parts of ``dora`` (facebookresearch/dora) the reference touches:
``get_xp()`` (``flashy/solver.py:33``, ``flashy/logging.py:64``, ``flashy/loggers/*.py``) and
``hydra_main`` (``examples/*/train.py``).  TEST-ONLY.

One experiment per thread (``use_xp``), so that several virtual ranks -- threads of one process --
each own their history, exactly like the reference's one-process-per-rank layout."""
import threading
import typing as tp
from pathlib import Path

_tls = threading.local()


class Link:
    """``xp.link``: the metric history Dora stores for an experiment (``flashy/solver.py:50-52,150-153``)."""

    def __init__(self):
        self.history: tp.List[tp.Dict[str, tp.Any]] = []
        self.updates = 0

    def update_history(self, history):
        self.history = list(history)
        self.updates += 1


class XP:
    def __init__(self, folder, cfg=None, sig="deadbeef"):
        self.folder = Path(folder)
        self.folder.mkdir(parents=True, exist_ok=True)
        # a plain dict: BaseSolver stores xp.cfg in the checkpoint (solver.py:35) and modern torch.load
        # (weights_only) refuses arbitrary classes
        self.cfg = dict(vars(cfg)) if hasattr(cfg, "__dict__") else cfg
        self.sig = sig
        self.link = Link()


def use_xp(xp: tp.Optional[XP]) -> None:
    _tls.xp = xp


def get_xp() -> XP:
    xp = getattr(_tls, "xp", None)
    if xp is None:
        raise RuntimeError("no experiment is active on this thread: call dora.use_xp(XP(folder)) first")
    return xp


def hydra_main(*_args, **_kwargs):
    def deco(fn):
        fn.dora = type("DoraConfig", (), {"dir": None})()
        return fn
    return deco
