"""Build libdv_b200.so in-tree with nvcc for sm_100a (explicit -gencode; no JIT cache involved)."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libdv_b200.so")
SOURCES = ["api.cu", "ddpm_step.cu", "conv_simt.cu", "conv_tc.cu", "conv_halo.cu", "norm.cu", "attention.cu", "attn_tc.cu", "attn_temporal.cu", "xattn_fold.cu", "cond.cu", "im2col.cu", "init_conv_tc.cu", "init_conv_rs.cu", "bwd_conv.cu", "wgrad_tc.cu", "bwd_norm.cu", "bwd_attn.cu", "attn_bwd_tc.cu", "train_misc.cu", "vivit.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC", "--threads", "0"]


def _stale() -> bool:
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "dv_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.isfile(d))


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source into dalle2_video_b200/libdv_b200.so; returns the library path."""
    if not force and not _stale():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libdv_b200.so")
    cmd = [nvcc, *NVCC_FLAGS, *[os.path.join(CSRC, s) for s in SOURCES], "-o", LIB + ".tmp"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    os.replace(LIB + ".tmp", LIB)
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
