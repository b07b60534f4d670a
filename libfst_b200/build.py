"""In-tree build of libfst_b200.so (nvcc, sm_100a only)."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
SO = os.environ.get("LIBFST_B200_SO") or os.path.join(HERE, "libfst_b200.so")   # override: tuning variants

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def sources():
    out = [os.path.join(ROOT, "include", "libfst_b200.h")]
    for f in sorted(os.listdir(CSRC)):
        out.append(os.path.join(CSRC, f))
    return out


def is_stale() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(s) > t for s in sources())


def build(force: bool = False, verbose: bool = False, defines=(), out: str | None = None) -> str:
    so = out or SO
    if not force and not is_stale() and out is None:
        return SO
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: libfst_b200 has no CPU build")
    cmd = [nvcc, *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-o", so, os.path.join(CSRC, "c_api.cu")]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return so


if __name__ == "__main__":
    print(build(force=True, verbose=True))
