#!/usr/bin/env python
"""Install the UNMODIFIED reference (facebookresearch/flashy, /root/reference) into baseline/_ref.

    python baseline/install_ref.py

``baseline/_ref`` is git-ignored (never committed) but travels to the GPU box with the repo
snapshot.  The reference's dependencies ``dora_search`` and ``colorlog`` are neither installed nor
in /opt/wheelhouse (no network), so the package is installed with ``--no-deps``; the solver tests
provide minimal stand-ins for the two (tests/shims/, test-only).  The source tree is read-only, so
the build runs from a copy under /tmp.  Prints one line with the outcome.
"""
import shutil
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")
DEST = ROOT / "baseline" / "_ref"


def install() -> str:
    if (DEST / "flashy" / "solver.py").exists():
        return f"already installed: {DEST}"
    if not REF.exists():
        return f"unavailable: {REF} does not exist on this machine"
    with tempfile.TemporaryDirectory() as tmp:
        src = Path(tmp) / "reference"
        shutil.copytree(REF, src, ignore=shutil.ignore_patterns(".git"))
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
               "--find-links", "/opt/wheelhouse", "--target", str(DEST), str(src)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            return "failed: " + (res.stderr.strip().splitlines() or ["pip error"])[-1]
    return f"installed {DEST} (pip --no-deps; dora_search / colorlog are not available offline)"


if __name__ == "__main__":
    print(install())
