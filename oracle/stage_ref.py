"""TEST / MEASUREMENT INFRASTRUCTURE — stages the UNMODIFIED Python reference for the GPU box.

The reference (helpingstar/pika-zoo) is pure Python: there is nothing to compile, but the bench must time
*it* — not only the C port — on the GPU box's own host cores, in the same run as the GPU numbers
(BASELINE.json north_star, BASELINE.md CPU-baseline plan). `/root/reference` does not exist there, while
git-ignored build artefacts in the tree do travel (like the built `.so` files). This recipe copies the
reference's Python sources (`pikazoo/**/*.py`, ~60 KB) and its sprites (`pikazoo/env/img/*.png`, 320 KB: the
renderer's GPU tests draw with them; the product never reads oracle/) byte for byte from where they lie into `oracle/_ref/`, which is listed in `.gitignore` (never enters the
history) and not in `.gpurunignore` (travels). Nothing under `oracle/_ref/` is ever imported by the product;
its only users are `oracle/ref_harness.py` (tests, golden generation) and `oracle/time_reference.py`
(bench.py's CPU legs). `STAGED.json` records the sha256 of every staged file so that a run on the box can
state exactly what it timed.

    python -m oracle.stage_ref            # called by __graft_entry__.build()
"""

from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCE = os.environ.get("PIKA_REFERENCE_SOURCE", "/root/reference")
DEST = os.path.join(HERE, "_ref")


def staged() -> bool:
    return os.path.isfile(os.path.join(DEST, "STAGED.json")) and os.path.isdir(os.path.join(DEST, "pikazoo"))


def stage(verbose: bool = False) -> bool:
    """Copy the reference's .py files into oracle/_ref/. Returns True if a staged copy exists afterwards
    (on the GPU box, where SOURCE is absent, the copy that travelled with the tree is kept)."""
    src_pkg = os.path.join(SOURCE, "pikazoo")
    if not os.path.isdir(src_pkg):
        return staged()
    if os.path.isdir(os.path.join(DEST, "pikazoo")):
        shutil.rmtree(os.path.join(DEST, "pikazoo"))
    manifest = {}
    for dirpath, _, files in os.walk(src_pkg):
        for f in sorted(files):
            if not f.endswith((".py", ".png")):  # the sprites too (320 KB): assets for tests/test_gpu_render.py
                continue
            src = os.path.join(dirpath, f)
            rel = os.path.relpath(src, SOURCE)
            dst = os.path.join(DEST, rel)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(src, dst)
            with open(dst, "rb") as fh:
                manifest[rel] = hashlib.sha256(fh.read()).hexdigest()
    for extra in ("LICENSE", "pyproject.toml"):
        if os.path.isfile(os.path.join(SOURCE, extra)):
            shutil.copyfile(os.path.join(SOURCE, extra), os.path.join(DEST, extra))
    with open(os.path.join(DEST, "STAGED.json"), "w") as fh:
        json.dump({"source": SOURCE, "what": "unmodified .py and .png files of helpingstar/pika-zoo, byte for byte",
                   "files": manifest}, fh, indent=1, sort_keys=True)
    if verbose:
        print(f"staged {len(manifest)} reference files into {DEST}")
    return True


if __name__ == "__main__":
    ok = stage(verbose=True)
    sys.exit(0 if ok else 1)
