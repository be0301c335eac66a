"""Build libseghiero_b200.so in-tree with nvcc for sm_100a (no torch headers involved).

Every .cu is compiled to its own object file (in parallel, only when it or a header changed), then linked."""
from __future__ import annotations

import concurrent.futures
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libseghiero_b200.so")
SOURCES = ["targets_decode.cu", "bce2.cu", "rmi3_fwd.cu", "rmi3_bwd.cu", "triplet.cu", "upsample.cu"]
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _headers_mtime() -> float:
    return max(os.path.getmtime(os.path.join(CSRC, f)) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h")))


def _compile(nvcc: str, src: str, force: bool, hdr_t: float):
    path = os.path.join(CSRC, src)
    obj = os.path.join(OBJ, src.replace(".cu", ".o"))
    if not force and os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(path), hdr_t):
        return obj, 0, ""
    cmd = [nvcc, *FLAGS, "-c", "-o", obj, path]
    res = subprocess.run(cmd, capture_output=True, text=True)
    return obj, res.returncode, " ".join(cmd) + "\n" + res.stdout + res.stderr


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu under csrc/ into one shared library. Returns its path."""
    os.makedirs(OBJ, exist_ok=True)
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    hdr_t = _headers_mtime()
    with concurrent.futures.ThreadPoolExecutor(max_workers=len(SOURCES)) as pool:
        results = list(pool.map(lambda s: _compile(nvcc, s, force, hdr_t), SOURCES))
    log = "".join(r[2] for r in results)
    objs = [r[0] for r in results]
    failed = [r for r in results if r[1] != 0]
    rebuilt = any(r[2] for r in results)
    if not failed and (rebuilt or force or not os.path.exists(LIB) or
                       any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs)):
        cmd = [nvcc, "--shared", "-o", LIB, *objs]
        res = subprocess.run(cmd, capture_output=True, text=True)
        log += " ".join(cmd) + "\n" + res.stdout + res.stderr
        if res.returncode != 0:
            failed.append((LIB, res.returncode, log))
    if log:
        with open(os.path.join(HERE, "build.log"), "w") as fh:
            fh.write(log)
    if verbose:
        print(log)
    if failed:
        raise RuntimeError("nvcc failed:\n" + log[-6000:])
    return LIB


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
