#!/usr/bin/env python3
"""Recipe: stage the UNMODIFIED reference package for the runs on the GPU box.

The reference is pure Python (no build step).  /root/reference exists only in the
build container, so ``__graft_entry__.build()`` runs this recipe there: it copies
``/root/reference/miniframe`` (the .py files and the bundled datasets, nothing edited)
to ``oracle/_ref/miniframe``.  ``oracle/_ref/`` is git-ignored - no reference source
enters this repository's history - but it is not gpurun-ignored, so it travels to the
GPU box, where ``bench.py --impl reference`` and ``cpu_baseline`` time the reference's
own classes (through ``oracle/ref_shim.py``) instead of the oracle's port of them.

Test infrastructure: nothing under ``miniframe_b200/`` imports it.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("MINIFRAME_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")


def stage(src=SRC, dst=DST):
    pkg = os.path.join(src, "miniframe")
    if not os.path.isdir(pkg):
        return None
    out = os.path.join(dst, "miniframe")
    if os.path.isdir(out):
        shutil.rmtree(out)
    shutil.copytree(pkg, out, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    return out


if __name__ == "__main__":
    out = stage()
    print("staged %s" % out if out else "no reference tree at %s: nothing staged" % SRC)
    sys.exit(0)
