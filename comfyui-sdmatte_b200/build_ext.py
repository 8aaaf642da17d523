"""Build libsdmatte_b200.so in-tree with nvcc for sm_100a (no torch headers, plain C ABI).

Usage: python build_ext.py [--force]
The .so lands next to this file so it travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import fcntl
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
OUT = os.path.join(HERE, "libsdmatte_b200.so")
BUILD = os.path.join(HERE, "build")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function",
    "-I", CSRC, "-I", os.path.join(ROOT, "include"),
    "--expt-relaxed-constexpr",
]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest() -> str:
    h = hashlib.sha256()
    for f in sorted(os.listdir(CSRC)):
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(f.encode())
            h.update(fh.read())
    with open(os.path.join(ROOT, "include", "sdmatte_b200.h"), "rb") as fh:
        h.update(fh.read())
    # the flags WITHOUT the absolute include paths: the tree is copied to another root on the GPU box, and a digest that
    # changes with the checkout location would make every process there rebuild (and N ranks race on the output file)
    h.update(" ".join(f for f in FLAGS if not os.path.isabs(f)).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(BUILD, exist_ok=True)
    stamp = os.path.join(BUILD, "digest.txt")
    dig = _digest()

    def fresh() -> bool:
        return os.path.exists(OUT) and os.path.exists(stamp) and open(stamp).read() == dig

    if not force and fresh():
        return OUT
    # one builder at a time (torchrun starts N ranks at once); the others wait here and then find a fresh library
    with open(os.path.join(BUILD, ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and fresh():
            return OUT
        return _build_locked(dig, stamp, verbose)


def _build_locked(dig: str, stamp: str, verbose: bool) -> str:
    if not os.path.exists(NVCC):
        if os.path.exists(OUT):
            return OUT  # GPU box without a toolkit: use the prebuilt library shipped with the snapshot
        raise RuntimeError("nvcc not found and no prebuilt libsdmatte_b200.so")

    def compile_one(src: str) -> str:
        obj = os.path.join(BUILD, src[:-3] + ".o")
        cmd = [NVCC, *FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose or r.stderr.strip():
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 4)) as ex:
        objs = list(ex.map(compile_one, _sources()))
    tmp_out = OUT + f".tmp{os.getpid()}"
    cmd = [NVCC, "-shared", "-o", tmp_out, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp_out, OUT)  # atomic: a process that dlopens concurrently sees the old or the new file, never a partial one
    with open(stamp, "w") as fh:
        fh.write(dig)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
