"""Access to tests/golden/ref_w{W}.npz (outputs of the unmodified reference over gloo)."""
from __future__ import annotations

import functools
import hashlib
from pathlib import Path

import numpy as np
import torch

from tests.golden import cases

GOLDEN_DIR = Path(__file__).resolve().parent / "golden"


@functools.lru_cache(maxsize=None)
def _load(world: int):
    return dict(np.load(GOLDEN_DIR / f"ref_w{world}.npz"))


def get(world: int, key: str, rank: int = 0) -> np.ndarray:
    data = _load(world)
    if f"all/{key}" in data:
        return data[f"all/{key}"]
    return data[f"r{rank}/{key}"]


def check_input(world: int, key: str, rank: int, tensor: torch.Tensor) -> None:
    """Assert a rebuilt input equals the one the reference was run on (digest compare)."""
    want = _load(world)[f"r{rank}/{key}.sha"].tobytes()
    got = hashlib.sha256(cases.to_np(tensor).tobytes()).digest()
    assert got == want, f"seeded input {key} (rank {rank}) no longer matches the golden run"


def golden_tensor(world: int, key: str, dtype: torch.dtype, rank: int = 0) -> torch.Tensor:
    return cases.from_np(get(world, key, rank), dtype)


def normalised_error(out: torch.Tensor, ref: torch.Tensor, inputs) -> float:
    """max |out-ref| / (sum_r |x_r| / W): the fp32 parity metric of BASELINE.md section 5."""
    if out.numel() == 0:
        return 0.0
    if out.is_complex():
        out, ref = torch.view_as_real(out), torch.view_as_real(ref)
        inputs = [torch.view_as_real(x) for x in inputs]
    scale = sum(x.detach().double().abs() for x in inputs) / len(inputs)
    scale = scale.clamp_min(torch.finfo(torch.float32).tiny)
    return float(((out.double() - ref.double()).abs() / scale).max())


def ulp_distance_bf16(a: torch.Tensor, b: torch.Tensor) -> int:
    """Largest distance in units of bf16 ulp between two bf16 tensors (sign-magnitude order)."""
    def key(t):
        bits = t.contiguous().view(torch.int16).to(torch.int32) & 0xFFFF
        return torch.where(bits >= 0x8000, 0x8000 - bits, bits)
    if a.numel() == 0:
        return 0
    return int((key(a) - key(b)).abs().max())
