"""The C-ABI library loads without a GPU, exports every symbol include/flashy_b200.h declares,
and its host-side planner (bucket layout) behaves -- no compute calls here."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def native():
    import __graft_entry__ as entry
    from flashy_b200 import _native
    if not _native.LIB_PATH.exists():
        entry.build()
    return _native


def test_header_symbols_are_exported(native):
    header = (ROOT / "include" / "flashy_b200.h").read_text()
    declared = set(re.findall(r"\b(fx_[a-z_]+)\s*\(", header))
    declared -= {"fx_comm", "fx_plan"}
    assert declared == set(native.EXPORTS), declared ^ set(native.EXPORTS)
    raw = C.CDLL(str(native.LIB_PATH))
    for name in declared:
        assert getattr(raw, name) is not None
    assert native.lib.fx_abi_version() == int(re.search(r"#define FX_ABI_VERSION (\d+)", header).group(1))


def test_library_has_no_libcuda_link_dependency(native):
    import subprocess
    out = subprocess.run(["ldd", str(native.LIB_PATH)], capture_output=True, text=True).stdout
    assert "libcuda.so" not in out      # driver API is resolved lazily through cudart
    assert "libtorch" not in out and "libc10" not in out


def _dry_plan(native, world, numels, dtype, wire=None, algo=0):
    arr = (C.c_int64 * len(numels))(*numels)
    handle = C.c_void_p()
    native.check(native.lib.fx_plan_create(None, world, arr, len(numels), dtype, dtype if wire is None else wire,
                                           algo, C.byref(handle)))
    info = native.PlanInfo()
    native.check(native.lib.fx_plan_get_info(handle, C.byref(info)))
    offs = (C.c_int64 * len(numels))()
    native.check(native.lib.fx_plan_offsets(handle, offs))
    native.lib.fx_plan_destroy(handle)
    return info, list(offs)


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("dtype,esize", [(0, 4), (1, 2), (3, 8)])
def test_bucket_layout_invariants(native, world, dtype, esize):
    numels = [1, 7, 165, 1024, 27, 4097, 64, 2359296, 10]
    info, offs = _dry_plan(native, world, numels, dtype)
    assert info.n_tensors == len(numels) and info.world == world
    assert info.total_elems == sum(numels)
    align = 16 // esize
    end = 0
    for off, n in zip(offs, numels):
        assert off % align == 0 and off >= end          # 16-byte aligned, in order, non-overlapping
        end = off + n
    assert info.padded_elems >= end
    shards = 1 if info.algo == native.FX_ALGO_ONE_SHOT else world
    assert info.shard_elems * shards == info.padded_elems
    assert info.shard_elems % info.grid_x == 0
    assert (info.shard_elems // info.grid_x * esize) % 128 == 0     # slices start on 128-byte lines
    assert info.wire_bytes == info.padded_elems * esize
    assert 1 <= info.grid_x <= 512 and info.block == 512


def test_algorithm_selection_and_signature(native):
    small, _ = _dry_plan(native, 8, [100, 200], 0)
    big, _ = _dry_plan(native, 8, [1 << 22], 0)
    assert small.algo == native.FX_ALGO_ONE_SHOT and big.algo == native.FX_ALGO_TWO_SHOT
    forced, _ = _dry_plan(native, 8, [100, 200], 0, algo=native.FX_ALGO_TWO_SHOT)
    assert forced.algo == native.FX_ALGO_TWO_SHOT
    again, _ = _dry_plan(native, 8, [100, 200], 0)
    other, _ = _dry_plan(native, 8, [100, 201], 0)
    assert small.signature == again.signature != other.signature
    wire, _ = _dry_plan(native, 8, [1 << 20], native.FX_F32, wire=native.FX_BF16)
    assert wire.wire_bytes == wire.padded_elems * 2


def test_errors_are_reported_not_crashed(native):
    handle = C.c_void_p()
    arr = (C.c_int64 * 1)(4)
    assert native.lib.fx_plan_create(None, 99, arr, 1, 0, 0, 0, C.byref(handle)) == native.FX_ERR_INVALID
    assert b"world" in native.lib.fx_last_error()
    assert native.lib.fx_plan_create(None, 2, arr, 1, native.FX_F16, native.FX_BF16, 0, C.byref(handle)) == native.FX_ERR_UNSUPPORTED
    assert native.lib.fx_plan_create(None, 2, arr, 1, 0, 0, native.FX_ALGO_NVLS, C.byref(handle)) == native.FX_ERR_UNSUPPORTED
    comm = C.c_void_p()
    assert native.lib.fx_comm_create(4, 3, 2, 0, 1 << 20, native.FX_COMM_HOST_ONLY, C.byref(comm)) == native.FX_ERR_INVALID
    with pytest.raises(native.NativeError):
        native.check(native.FX_ERR_INVALID)


def test_host_only_comm_exchange_single_process(native):
    """All ranks in one process (virtual layout), no CUDA: the count-check fabric alone."""
    import threading
    comm = C.c_void_p()
    native.check(native.lib.fx_comm_create(4, 0, 4, -1, 0, native.FX_COMM_HOST_ONLY, C.byref(comm)))
    info = native.CommInfo()
    native.check(native.lib.fx_comm_get_info(comm, C.byref(info)))
    assert (info.world, info.n_local, info.connected, info.mem_kind) == (4, 4, 1, 0)
    sums, equal = [None] * 4, [None] * 4

    def body(l):
        for step in range(50):
            total, same = C.c_int64(), C.c_int()
            count = 3 if (step % 2 == 0 or l != 2) else 4       # rank 2 deviates on odd steps
            native.check(native.lib.fx_host_exchange(comm, l, count, 1234 + (l == 1 and step == 7),
                                                     C.byref(total), C.byref(same), 30.0))
            assert total.value == (12 if step % 2 == 0 else 13), (step, total.value)
            assert bool(same.value) == (step != 7)
        sums[l], equal[l] = total.value, same.value

    threads = [threading.Thread(target=body, args=(l,)) for l in range(4)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert sums == [13] * 4
    # a data launch on a host-only communicator must fail loudly
    assert native.lib.fx_barrier(comm, None) == native.FX_ERR_STATE
    native.lib.fx_comm_destroy(comm)


def test_bucket_splitting_is_deterministic_and_order_preserving():
    """flashy_b200.distrib.split_buckets: the host-side cut of a tensor list into launches."""
    from flashy_b200.distrib import split_buckets
    items = [(0, 10), (1, 50), (2, 45), (3, 260), (4, 5), (5, 100), (6, 1)]
    got = split_buckets(items, cap=100, esize=4)
    assert got == [
        ([0, 1], [0, 0], [10, 50]),                   # 60 <= 100, next (45) would overflow
        ([2], [0], [45]),                             # flushed before the oversized tensor
        ([3], [0], [100]), ([3], [400], [100]), ([3], [800], [60]),   # 260 cut in three, byte offsets
        ([4], [0], [5]),                              # 5 + 100 would overflow
        ([5], [0], [100]),                            # exactly one bucket's worth
        ([6], [0], [1]),
    ]
    # every element appears exactly once, in order
    flat = [(i, o, n) for idx, off, num in got for i, o, n in zip(idx, off, num)]
    assert [i for i, _, _ in flat] == sorted(i for i, _, _ in flat)
    for i, numel in items:
        assert sum(n for j, _, n in flat if j == i) == numel
    assert split_buckets([], 100, 4) == []
    assert split_buckets([(0, 100)], 100, 2) == [([0], [0], [100])]
