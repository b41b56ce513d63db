"""Benchmark / baseline plumbing that must work without a GPU: the CPU baseline's core pinning and
launcher-environment scrubbing (the reference arm hung under torch.distributed.run),
and the diagnostics entry points of the C ABI on a host-only communicator."""
import ctypes as C
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_core_partition_is_disjoint_and_even():
    from oracle import cpu_train
    sets = cpu_train.partition_cores(range(128), 8)
    assert [len(s) for s in sets] == [16] * 8
    assert sorted(c for s in sets for c in s) == list(range(128))
    assert cpu_train.partition_cores(range(8), 8) == [[i] for i in range(8)]
    assert cpu_train.partition_cores(range(4), 8) == [[] for _ in range(8)]          # fewer cores than ranks: no pinning
    odd = cpu_train.partition_cores([0, 2, 4, 6, 8, 10, 12], 2)
    assert odd == [[0, 2, 4], [6, 8, 10]]


def test_launcher_environment_is_scrubbed():
    from oracle import cpu_train
    env = {"RANK": "3", "WORLD_SIZE": "8", "LOCAL_RANK": "3", "MASTER_ADDR": "10.0.0.1", "MASTER_PORT": "1234",
           "TORCHELASTIC_USE_AGENT_STORE": "True", "TORCHELASTIC_RUN_ID": "x", "GROUP_RANK": "0", "OMP_NUM_THREADS": "1",
           "PATH": "/usr/bin", "FLASHY_B200_WIRE": "bf16"}
    cpu_train.scrub_launcher_env(env)
    assert env == {"PATH": "/usr/bin", "FLASHY_B200_WIRE": "bf16"}


def test_reference_arm_runs_under_a_launcher_environment():
    """`bench.py --impl reference` with the variables torch.distributed.run exports to its workers (the
    round-1 hang: env:// became a client of the agent's store)."""
    env = dict(os.environ, RANK="0", WORLD_SIZE="2", LOCAL_RANK="0", MASTER_ADDR="127.0.0.1", MASTER_PORT="29999",
               TORCHELASTIC_USE_AGENT_STORE="True", TORCHELASTIC_RUN_ID="none", GROUP_RANK="0", LOCAL_WORLD_SIZE="2")
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "0",
                          "--world", "2", "--batch", "2"], capture_output=True, text=True, timeout=280, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and "pinned" in line["cpu_baseline"]["sample"]
    assert "passive" in line["cpu_baseline"]["sample"]
    assert line["config"]["world"] == 2 and "ranks_per_gpu" not in line["config"]       # same config object as the native arm
    # rank 1 of the launcher prints nothing and exits 0
    env["RANK"] = "1"
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_diagnostics_entry_points_on_a_host_only_communicator():
    from flashy_b200 import _native as N
    comm = C.c_void_p()
    N.check(N.lib.fx_comm_create(1, 0, 1, -1, 0, N.FX_COMM_HOST_ONLY, C.byref(comm)))
    try:
        words = C.c_size_t(123)
        N.check(N.lib.fx_comm_trace_read(comm, None, 0, C.byref(words)))
        assert words.value == 0                                       # tracing is off (no device side)
        arenas = (C.c_void_p * 16)()
        rc = N.lib.fx_comm_get_pointers(comm, arenas, None, None, None)
        assert rc == N.FX_ERR_STATE and b"device communicator" in N.lib.fx_last_error()
    finally:
        N.lib.fx_comm_destroy(comm)
