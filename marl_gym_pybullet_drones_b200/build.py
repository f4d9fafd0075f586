"""Build `libbatchdrones.so` in-tree with nvcc for sm_100a.

    python -m marl_gym_pybullet_drones_b200.build [--force]

The library is a plain C-ABI shared object (no torch, no pybind): it travels
with the source tree and is loaded through ctypes (`_native.py`).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
SOURCES = [os.path.join(CSRC, "bd_kernels.cu"), os.path.join(CSRC, "bd_tile_launch.cu"), os.path.join(CSRC, "bd_api.cu"), os.path.join(CSRC, "bd_actor.cu"),
           os.path.join(CSRC, "bd_norm.cu"), os.path.join(CSRC, "bd_ppo.cu"), os.path.join(CSRC, "bd_peer.cu")]
HEADERS = [os.path.join(CSRC, "bd_params.h"), os.path.join(CSRC, "bd_device.cuh"), os.path.join(CSRC, "bd_step_tile.cuh"),
           os.path.join(CSRC, "bd_umma.cuh"),
           os.path.join(os.path.dirname(_HERE), "include", "batch_drones.h")]
OUT = os.path.join(_HERE, "libbatchdrones.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def find_nvcc() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found; cannot build the CUDA extension")
    return cand


def up_to_date() -> bool:
    if not os.path.exists(OUT):
        return False
    t = os.path.getmtime(OUT)
    return all(os.path.getmtime(p) <= t for p in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every translation unit (in parallel, objects under csrc/_obj/) and link the shared library."""
    if not force and up_to_date():
        return OUT
    from concurrent.futures import ThreadPoolExecutor
    nvcc = find_nvcc()
    objdir = os.path.join(CSRC, "_obj")
    os.makedirs(objdir, exist_ok=True)
    compile_flags = [f for f in NVCC_FLAGS if f != "-shared"]

    def compile_one(src):
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc] + compile_flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed on " + src + ":\n" + res.stdout + res.stderr)
        if verbose:
            print(res.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as pool:
        objs = list(pool.map(compile_one, SOURCES))
    res = subprocess.run([nvcc] + NVCC_FLAGS + ["-o", OUT] + objs, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + res.stdout + res.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
