#!/usr/bin/env python
"""Which placement of the CPU baseline (8 gloo ranks of the ResNet-18 step) is fastest AND stable on this host?
Benchmark infrastructure: informs the defaults of oracle/cpu_train.py.  Prints one JSON line per variant."""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle import cpu_train  # noqa: E402

if __name__ == "__main__":
    for name, kw in (("pinned", dict(pin=True)), ("pinned_spare1_passive", dict(pin=True, spare=1, passive=True)),
                     ("pinned_passive", dict(pin=True, passive=True)), ("unpinned", dict(pin=False)),
                     ("unpinned_passive", dict(pin=False, passive=True))):
        res = cpu_train.run(world=8, batch=64, steps=4, warmup=1, **kw)
        print(json.dumps({"variant": name, "samples_per_s": round(res["value"], 1), "ms_per_step": round(res["ms_per_step"]),
                          "sync_ms": round(res["sync_ms_per_step"]), "step_ms": [round(x) for x in res["step_ms"]]}), flush=True)
