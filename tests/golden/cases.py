"""Seeded inputs shared by the golden-vector generator and the parity tests.

Everything here is deterministic CPU torch: rank r's tensors come from
``torch.Generator().manual_seed(1000 + r + 7919 * case_index)``.
The generator script also stores the inputs in the .npz, so a torch RNG change cannot
silently invalidate the fixtures: tests rebuild the inputs and first assert they equal
the stored ones.
"""
from __future__ import annotations

import typing as tp

import numpy as np
import torch

WORLDS = (2, 4, 8)

AVG_SHAPES: tp.Tuple[tp.Tuple[int, ...], ...] = ((1,), (7,), (33, 5), (1024,), (3, 3, 3), (4097,))
AVG_DTYPES = {
    "fp32": torch.float32,
    "bf16": torch.bfloat16,
    "fp16": torch.float16,
    "fp64": torch.float64,
    "c64": torch.complex64,
}
INT_SLOT = 2            # position of the int64 tensor that every collective must skip


def _gen(rank: int, case: int) -> torch.Generator:
    g = torch.Generator()
    g.manual_seed(1000 + rank + 7919 * case)
    return g


def avg_inputs(rank: int, dtype_name: str, scale: float = 1e-2) -> tp.List[torch.Tensor]:
    """Tensor list for the ``average_tensors`` / ``broadcast_tensors`` cases."""
    dtype = AVG_DTYPES[dtype_name]
    g = _gen(rank, list(AVG_DTYPES).index(dtype_name))
    out = []
    for shape in AVG_SHAPES:
        if dtype.is_complex:
            t = torch.randn(*shape, 2, generator=g, dtype=torch.float32) * scale
            t = torch.view_as_complex(t.contiguous())
        else:
            t = (torch.randn(*shape, generator=g, dtype=torch.float32) * scale).to(dtype)
        out.append(t)
    out.insert(INT_SLOT, torch.arange(5, dtype=torch.long) + 100 * rank)
    return out


def model_shapes() -> tp.List[tp.Tuple[str, tp.Tuple[int, ...]]]:
    """Parameter / buffer layout of the small conv+BN model used for sync_model cases."""
    return [
        ("0.weight", (8, 3, 3, 3)), ("0.bias", (8,)),
        ("1.weight", (8,)), ("1.bias", (8,)),
        ("3.weight", (10, 288)), ("3.bias", (10,)),
    ]


def make_model() -> torch.nn.Module:
    torch.manual_seed(1234)
    return torch.nn.Sequential(
        torch.nn.Conv2d(3, 8, 3),
        torch.nn.BatchNorm2d(8),
        torch.nn.Flatten(),
        torch.nn.Linear(8 * 6 * 6, 10),
    )


def model_local_state(rank: int):
    """Per-rank local gradients and BN buffers to inject before calling sync_model."""
    g = _gen(rank, 50)
    grads = [torch.randn(*shape, generator=g) * 1e-2 for _, shape in model_shapes()]
    running_mean = torch.randn(8, generator=g) * 0.1
    running_var = torch.rand(8, generator=g) + 0.5
    num_batches = torch.tensor(3 + rank, dtype=torch.long)
    return grads, [running_mean, running_var, num_batches]


def metrics_inputs(rank: int):
    return {"acc": 0.5 + 0.03125 * rank, "loss": 1.0 / (rank + 1), "z": float(rank)}, float(10 + rank)


def allreduce_inputs(rank: int):
    g = _gen(rank, 60)
    return [torch.randn(100, generator=g), torch.arange(3, dtype=torch.long) * (rank + 1)]


LOADER_N = 103


# ---- (de)serialisation helpers: bf16/complex have no numpy dtype ----------------------

def to_np(t: torch.Tensor) -> np.ndarray:
    t = t.detach().cpu().contiguous()
    if t.dtype == torch.bfloat16:
        return t.view(torch.int16).numpy().view(np.uint16).copy()
    if t.dtype == torch.complex64:
        return torch.view_as_real(t).numpy().copy()
    return t.numpy().copy()


def from_np(a: np.ndarray, dtype: torch.dtype) -> torch.Tensor:
    if dtype == torch.bfloat16:
        return torch.from_numpy(a.view(np.int16).copy()).view(torch.bfloat16)
    if dtype == torch.complex64:
        return torch.view_as_complex(torch.from_numpy(a.copy()))
    return torch.from_numpy(a.copy())
