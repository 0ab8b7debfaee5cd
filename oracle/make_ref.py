#!/usr/bin/env python
"""Materialise the UNMODIFIED reference's own implementation of the PE path under oracle/_ref/.

    python oracle/make_ref.py            # build container only: needs /root/reference

The reference is plain Python (no build system, nothing to compile): the "build" of its path is a
byte-for-byte copy of the few source files the path and its caller loop import, taken from where
they lie under /root/reference into the git-ignored output directory oracle/_ref/ (listed in
.gitignore, NOT in .gpurunignore: like the built .so it travels to the GPU box with the snapshot
and stays out of the history — no reference source is ever committed). A manifest with the sha256
of every file is written next to them so a test can show the copies are unmodified.

What uses it (and only as the checker / the timed baseline, never as the product path):
  * bench.py --impl reference  and bench.py's cpu_baseline leg  -> kind "reference"
  * tests/test_reference_loop_gpu.py: the reference's untouched evaluate_model_link_prediction
    driving lstep_b200.LSTEP / NeighborSampler (the drop-in claim)
  * tests/golden/make_golden.py reads /root/reference directly (build container)
Imports the reference needs that are not installed offline come from oracle/shims/ (our own
restatements of torch_scatter.scatter, tgb's import, seven torch_geometric.utils functions).
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
FILES = [
    "models/LSTEP.py", "models/modules.py", "models/EdgeBank.py",
    "utils/utils.py", "utils/DataLoader.py", "utils/metrics.py", "utils/EarlyStopping.py", "utils/PositionalEncoding.py",
    "evaluate_model_utils.py", "LICENSE",
]


def sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def make(ref="/root/reference", quiet=False):
    if not os.path.isdir(ref):
        if not quiet:
            print(f"make_ref: {ref} not present (GPU box): keeping the prebuilt oracle/_ref", file=sys.stderr)
        return os.path.isdir(OUT)
    manifest = {}
    for rel in FILES:
        dst = os.path.join(OUT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(ref, rel), dst)
        manifest[rel] = sha(dst)
        assert manifest[rel] == sha(os.path.join(ref, rel))
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump({"source": ref, "sha256": manifest}, f, indent=1)
    if not quiet:
        print(f"make_ref: {len(FILES)} files -> {OUT}")
    return True


if __name__ == "__main__":
    make(os.environ.get("LSTEP_REFERENCE", "/root/reference"))
