"""Build recipe for ``liblwse.so`` — the CUDA engine behind ``include/lwse.h``.

nvcc cross-compiles for sm_100a without a GPU; the library is built in-tree
(``lws_b200/liblwse.so``) so that it travels with the repository snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "liblwse.so")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-Wall",
    "-shared", "-cudart", "shared",
]


def nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def sources() -> list[str]:
    srcs = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))
    return srcs


def defines() -> list[str]:
    return []


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    deps = sources() + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    deps.append(os.path.join(HERE, "..", "include", "lwse.h"))
    deps.append(os.path.abspath(__file__))
    return os.path.getmtime(LIB) < max(os.path.getmtime(d) for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return LIB
    cmd = [nvcc(), *NVCC_FLAGS, *defines(), "-o", LIB, *sources()]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd), file=sys.stderr)
    # ptxas with the image's default CC wrapper: force the system g++ as host compiler
    env = dict(os.environ)
    r = subprocess.run(cmd + ["-ccbin", "/usr/bin/g++"], capture_output=True, text=True, env=env)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed:\n{r.stdout}\n{r.stderr}")
    if verbose:
        print(r.stderr, file=sys.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
