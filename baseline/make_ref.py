#!/usr/bin/env python
"""Ship the UNMODIFIED reference to the GPU box.

`/root/reference` exists only in the build container; the GPU box receives `/root/repo`.  This script copies the
reference's own Python files for the hot path (SURVEY 8c list: modules/, utils/, experiment_modules/, options.py,
losses.py, configs/, LICENSE) byte for byte into `baseline/_ref/reference/`, and the import shims for the third-party
packages this image lacks (tests/golden/shims: kornia, timm, antialiased_cnns, pytorch_lightning, matplotlib, moviepy,
pytorch3d) into `baseline/_ref/shims/`.  `baseline/_ref/` is git-ignored (no reference source enters the history) but
not gpurun-ignored, so it travels with the snapshot.  Run by `__graft_entry__.build()` whenever /root/reference is
present; `baseline/ref_loader.py` imports from the copy.

    python baseline/make_ref.py [--src /root/reference]
"""
import argparse
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
DST = os.path.join(HERE, "_ref")
PARTS = ["modules", "utils", "experiment_modules", "options.py", "losses.py", "configs", "LICENSE"]


def main(src="/root/reference", quiet=False):
    if not os.path.isdir(src):
        if not quiet:
            print(f"{src} not present: nothing to copy (the GPU box uses the copy made in the build container)")
        return False
    ref_dst = os.path.join(DST, "reference")
    shim_dst = os.path.join(DST, "shims")
    for d in (ref_dst, shim_dst):
        if os.path.isdir(d):
            shutil.rmtree(d)
    os.makedirs(ref_dst)
    manifest = {}
    for part in PARTS:
        s = os.path.join(src, part)
        d = os.path.join(ref_dst, part)
        if os.path.isdir(s):
            shutil.copytree(s, d, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        elif os.path.isfile(s):
            shutil.copy2(s, d)
    for base, _, files in os.walk(ref_dst):
        for f in sorted(files):
            p = os.path.join(base, f)
            manifest[os.path.relpath(p, ref_dst)] = hashlib.sha256(open(p, "rb").read()).hexdigest()
    shutil.copytree(os.path.join(ROOT, "tests", "golden", "shims"), shim_dst,
                    ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": src, "files": manifest}, f, indent=1, sort_keys=True)
    if not quiet:
        print(f"copied {len(manifest)} reference files + shims into {DST}")
    return True


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    sys.exit(0 if main(ap.parse_args().src) else 1)
