"""Stage the reference tree into the git-ignored ``baseline/_ref/reart`` so that it travels to the GPU box.

    python scripts/stage_reference.py

TEST INFRASTRUCTURE ONLY.  ``/root/reference`` exists in the build container and not on the GPU box; ``gpurun``
snapshots untracked files, ``baseline/_ref/`` is in ``.gitignore`` (never committed) and not in ``.gpurunignore``.
The staged copy is what ``scripts/run_reference_dropin.py`` drives UNMODIFIED over ``reart_b200.dropin``.
Only the files the run scripts need are staged (python sources + the nao demo data and checkpoints; no assets).
"""
from __future__ import annotations

import os
import shutil
import sys

SRC = os.environ.get("REART_REFERENCE_ROOT", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DST = os.path.join(ROOT, "baseline", "_ref", "reart")


def main() -> int:
    if not os.path.isdir(os.path.join(SRC, "utils")):
        print(f"reference tree not found at {SRC}", file=sys.stderr)
        return 1
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    ignore = shutil.ignore_patterns(".git", "assets", "__pycache__", "*.gif", "*.pyc")
    shutil.copytree(SRC, DST, ignore=ignore)
    n = sum(len(f) for _, _, f in os.walk(DST))
    print(f"staged {n} files from {SRC} into {DST}")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
