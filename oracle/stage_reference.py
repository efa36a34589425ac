"""Stage an importable copy of the UNMODIFIED reference package under oracle/_ref/ (git-ignored, travels to the GPU box
with gpurun like a built .so) so that the plugin path — prismo.set_backend("b200") + the reference's own Simulation,
sources and monitors — can be exercised against the real CUDA library on a real GPU.  Test infrastructure only.

The reference is pure Python with a hatchling / hatch-vcs build backend that is not installed in this image (no
network), so `pip install --target oracle/_ref /root/reference` cannot build a wheel; for a pure-Python package an
install IS a copy of the package directory, which is what this does (files are byte-identical; INSTALL_RECORD.txt lists
them with their sha256).  Nothing under oracle/_ref is ever committed.

    python oracle/stage_reference.py            (also run by __graft_entry__.build() when /root/reference exists)
"""
from __future__ import annotations

import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("PRISMO_REFERENCE_TREE", "/root/reference")
DST = os.path.join(HERE, "_ref")


def stage() -> bool:
    pkg = os.path.join(SRC, "src", "prismo")
    if not os.path.isdir(pkg):
        return False
    out = os.path.join(DST, "prismo")
    if os.path.isdir(out):
        shutil.rmtree(out)
    os.makedirs(DST, exist_ok=True)
    shutil.copytree(pkg, out, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    lines = []
    for root, _, files in os.walk(out):
        for f in sorted(files):
            p = os.path.join(root, f)
            lines.append(f"{hashlib.sha256(open(p, 'rb').read()).hexdigest()}  {os.path.relpath(p, DST)}")
    with open(os.path.join(DST, "INSTALL_RECORD.txt"), "w") as fh:
        fh.write(f"# copy of {pkg} (pure-Python package; hatchling is not installed, see oracle/stage_reference.py)\n")
        fh.write("\n".join(sorted(lines)) + "\n")
    return True


if __name__ == "__main__":
    ok = stage()
    print(f"staged reference package under {DST}" if ok else f"no reference tree at {SRC}: nothing staged")
    sys.exit(0)
