"""In-tree build of libcrfp_b200.so (nvcc, -gencode arch=compute_100a,code=sm_100a, -lineinfo)."""
from __future__ import annotations

import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))


def build(verbose: bool = False) -> str:
    """Compile every CUDA source under csrc/ into lib/libcrfp_b200.so; returns the library path."""
    cmd = ["make", "-C", os.path.join(_HERE, "csrc"), "-j", str(os.cpu_count() or 4)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
        print(res.stderr)
    if res.returncode != 0:
        raise RuntimeError("building libcrfp_b200.so failed")
    path = os.path.join(_HERE, "lib", "libcrfp_b200.so")
    assert os.path.exists(path)
    return path
