"""Builds the in-tree C-ABI shared library `libb200planesweep.so` with nvcc for sm_100a.

nvcc cross-compiles without a GPU, so this runs in the build container; the resulting .so is
git-ignored but travels with the repo snapshot to the GPU box.  The sha256 of the sources is compiled INTO the
library (`b200_source_digest()`): `build()` rebuilds when it differs from the sources on disk, and `_abi.load()`
refuses a library whose digest does not match the `csrc/` next to it -- a stale binary is never called through
newer ctypes signatures.  `build_dev()` makes `libb200probe.so` from `csrc/dev/` (tcgen05 self-test and MMA-rate
probes: test / tuning tools, not part of the product library)."""
from __future__ import annotations

import ctypes
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libb200planesweep.so")
DEV_LIB = os.path.join(HERE, "libb200probe.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "-shared",
]


def sources(sub=""):
    d = os.path.join(CSRC, sub)
    return sorted(os.path.join(d, f) for f in os.listdir(d) if f.endswith(".cu"))


def source_digest(sub=""):
    h = hashlib.sha256()
    dirs = [CSRC] + ([os.path.join(CSRC, sub)] if sub else [])
    for d in dirs:
        for f in sorted(os.listdir(d)):
            if f.endswith((".cu", ".cuh", ".h")):
                h.update(f.encode())
                h.update(open(os.path.join(d, f), "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def library_digest(path=LIB):
    """Digest compiled into an existing library, or None."""
    if not os.path.exists(path):
        return None
    try:
        fn = ctypes.CDLL(path).b200_source_digest
    except (OSError, AttributeError):
        return None
    fn.restype = ctypes.c_char_p
    return fn().decode()


def _nvcc(out, srcs, digest, verbose):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + [f'-DB200_SRC_DIGEST="{digest}"'] + (["-Xptxas", "-v"] if verbose else []) + \
          ["-o", out] + srcs + ["-lcudart"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed building {os.path.basename(out)}")


def build(force: bool = False, verbose: bool = False) -> str:
    digest = source_digest()
    if not force and library_digest(LIB) == digest:
        return LIB
    _nvcc(LIB, sources(), digest, verbose)
    return LIB


def build_dev(force: bool = False, verbose: bool = False) -> str:
    stamp = DEV_LIB + ".digest"
    digest = source_digest("dev")
    if not force and os.path.exists(DEV_LIB) and os.path.exists(stamp) and open(stamp).read().strip() == digest:
        return DEV_LIB
    _nvcc(DEV_LIB, sources("dev"), digest, verbose)
    with open(stamp, "w") as f:
        f.write(digest)
    return DEV_LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_dev(force="--force" in sys.argv, verbose="-v" in sys.argv))
