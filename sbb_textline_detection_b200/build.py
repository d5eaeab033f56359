"""In-tree build of the CUDA library (csrc/ -> libsbb_textline.so) with nvcc for sm_100a.
No torch extension machinery: the library is a plain C-ABI shared object (static cudart)."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsbb_textline.so")
SOURCES = ["sbb_net.cu"]
DEPS = ["sbb_net.cu", "conv_gemm_tc.cuh", "kernels_aux.cuh", "epilogue.cuh", "plan.h", "ptx.cuh", "prepost.cuh",
        os.path.join("..", "..", "include", "sbb_textline.h")]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libsbb_textline.so")


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False, defines=(), out: str = LIB) -> str:
    """Compile for sm_100a (-lineinfo so ncu's source page maps to csrc/).  `defines` / `out`: experiment builds
    next to the product library (e.g. defines=("SBB_ISSUE_PROBE",), out=".../libsbb_exp.so", loaded with SBB_LIB)."""
    if not force and out == LIB and not is_stale():
        return LIB
    cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-Xcompiler", "-fPIC", "-shared", "-ldl", "-o", out] + [f"-D{d}" for d in defines] + \
          [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return out


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--define", action="append", default=[], help="extra -D for an experiment build")
    ap.add_argument("--out", default=LIB)
    a = ap.parse_args()
    print(build(force=True, verbose=True, defines=a.define, out=a.out))
