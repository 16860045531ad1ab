"""Recipe for ``oracle/_ref/``: the UNMODIFIED reference sources, made available to the GPU box.

    python oracle/make_ref.py        # in the build container (needs /root/reference); idempotent

The reference is pure Python, so "building" it is placing its own files where they can travel: ``/root/reference``
does not exist on the GPU box, ``oracle/_ref/`` (git-ignored, not gpurun-ignored) does.  The tree is copied byte for
byte -- ``oracle/_ref/MANIFEST.json`` records every file's sha256 next to the source path so that the copy can be
checked against the original (``tests/test_oracle.py::test_ref_copy_is_unmodified``).  Nothing under ``oracle/_ref``
is ever committed, imported by the product, or edited.

Users: ``oracle/ref_runner.py`` (bench.py's ``--impl reference`` arm and ``cpu_baseline`` leg time the reference's own
``ClipLoss`` on the host cores) and ``experiments/train_step.py`` (the reference's model towers and
``train_one_epoch`` around the drop-in loss).
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

SRC = "/root/reference/src"
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
PACKAGES = ("open_clip", "open_clip_train")


def sha256(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()


def make(verbose=True):
    if not os.path.isdir(SRC):
        if verbose:
            print(f"{SRC} not present (GPU box?): keeping oracle/_ref as shipped")
        return os.path.isdir(DST)
    manifest = {}
    for pkg in PACKAGES:
        for root, dirs, files in os.walk(os.path.join(SRC, pkg)):
            dirs[:] = [d for d in dirs if d != "__pycache__"]
            for fn in files:
                if fn.endswith(".pyc"):
                    continue
                src = os.path.join(root, fn)
                rel = os.path.relpath(src, SRC)
                dst = os.path.join(DST, "src", rel)
                os.makedirs(os.path.dirname(dst), exist_ok=True)
                shutil.copyfile(src, dst)
                manifest[rel] = {"source": src, "sha256": sha256(dst)}
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    if verbose:
        print(f"oracle/_ref: {len(manifest)} reference files copied unmodified from {SRC}")
    return True


if __name__ == "__main__":
    sys.exit(0 if make() else 1)
