"""TEST / BASELINE INFRASTRUCTURE — makes the LIVE reference travel to the GPU box.

The reference (lan496/torch-m3gnet) is pure Python, so "building" it is a copy: this recipe copies the package
directory ``/root/reference/src/torch_m3gnet`` — unmodified — into ``oracle/_ref/torch_m3gnet`` (git-ignored: the
reference's sources never enter this repository's history; NOT gpurun-ignored, so the copy ships with the snapshot).
``oracle/live_reference.py`` imports it from there when ``/root/reference`` is absent, behind the same four stand-in
modules (torch_scatter, torch_geometric, torchtyping, pymatgen are not installed).  ``bench.py --impl reference`` and the
``cpu_baseline`` leg time it (``kind: "reference"``); nothing under ``torch_m3gnet_b200/`` may import it.

  python -m oracle.build_ref        (also run by __graft_entry__.build() when /root/reference is present)
"""
from __future__ import annotations

import os
import shutil

SRC = "/root/reference/src/torch_m3gnet"
DST_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
DST = os.path.join(DST_ROOT, "torch_m3gnet")


def build(verbose: bool = False) -> bool:
    """Copy the reference package; returns True when oracle/_ref holds it afterwards."""
    if not os.path.isdir(SRC):
        return os.path.isdir(DST)
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST_ROOT, exist_ok=True)
    shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    with open(os.path.join(DST_ROOT, "README"), "w") as f:
        f.write("Unmodified copy of /root/reference/src/torch_m3gnet made by oracle/build_ref.py (git-ignored).\n")
    if verbose:
        n = sum(len(fs) for _, _, fs in os.walk(DST))
        print(f"oracle/_ref: {n} files copied from {SRC}")
    return True


if __name__ == "__main__":
    build(verbose=True)
