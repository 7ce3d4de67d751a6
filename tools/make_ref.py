"""Populate baseline/_ref/ with a read-only copy of the reference's Python package:
the directory is git-ignored (never enters history) but NOT gpurun-ignored, so the unmodified reference travels to
the GPU box, where tests/test_gpu_reference.py and `bench.py --impl reference` import it through oracle/ref_shims.
Run in the build container (needs /root/reference); __graft_entry__.build() calls it."""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/lab4d"
DST = os.path.join(ROOT, "baseline", "_ref")


def make(force=False):
    if not os.path.isdir(SRC):
        return os.path.isdir(os.path.join(DST, "lab4d"))
    dst = os.path.join(DST, "lab4d")
    if os.path.isdir(dst) and not force:
        return True
    os.makedirs(DST, exist_ok=True)
    shutil.rmtree(dst, ignore_errors=True)
    shutil.copytree(SRC, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "build", "*.so"))
    return True


if __name__ == "__main__":
    print("baseline/_ref ready" if make(force="--force" in sys.argv) else "reference not available")
