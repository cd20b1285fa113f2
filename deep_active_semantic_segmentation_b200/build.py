"""In-tree build of libdas_b200.so (sm_100a only) with plain nvcc - no torch types, no JIT cache.

    python -m deep_active_semantic_segmentation_b200.build [--force] [--verbose]

The library is linked next to this file so that it travels with the repository snapshot to the
GPU box (it is git-ignored, not gpurun-ignored).  Objects are rebuilt when a source / header is
newer than the object or the flags changed.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(CSRC, "_build")
LIB = os.path.join(PKG, "libdas_b200.so")
INCLUDE = os.path.join(ROOT, "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-I", INCLUDE, "-I", CSRC,
]

# class-count ranges of the K1/K2 template instantiations (must match DAS_DECL_RANGE in mc_api.cu)
MC_RANGES = [(2, 9), (10, 16), (17, 20), (21, 24), (25, 28), (29, 32)]


def _units():
    units = []
    for name in ("handle", "mc_api", "region", "topk", "kcenter", "gram", "accuracy", "maxsubset", "noise"):
        units.append((name, os.path.join(CSRC, name + ".cu"), []))
    for lo, hi in MC_RANGES:
        units.append((f"mc_inst_{lo}_{hi}", os.path.join(CSRC, "mc_inst.cu"), [f"-DDAS_C_LO={lo}", f"-DDAS_C_HI={hi}"]))
    return units


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found - libdas_b200.so cannot be built")
    return exe


def _deps_mtime() -> float:
    files = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))]
    files += [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE)]
    files.append(os.path.abspath(__file__))
    return max(os.path.getmtime(f) for f in files)


def _flags_tag() -> str:
    return hashlib.sha256(" ".join(NVCC_FLAGS + [str(MC_RANGES)]).encode()).hexdigest()[:12]


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    newest = _deps_mtime()
    tag_file = os.path.join(OBJ, "flags.tag")
    tag = _flags_tag()
    if not os.path.exists(tag_file) or open(tag_file).read() != tag:
        force = True

    def compile_one(unit):
        name, src, defs = unit
        obj = os.path.join(OBJ, name + ".o")
        if not force and os.path.exists(obj) and os.path.getmtime(obj) >= newest:
            return obj, False
        cmd = [nvcc] + NVCC_FLAGS + defs + ["-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {name}:\n{r.stdout}\n{r.stderr}")
        if verbose and r.stderr:
            print(r.stderr)
        return obj, True

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 4)) as ex:
        results = list(ex.map(compile_one, _units()))
    objs = [o for o, _ in results]
    rebuilt = any(changed for _, changed in results)
    stale = os.path.exists(LIB) and any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs)
    if rebuilt or stale or not os.path.exists(LIB) or force:
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs + ["-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(tag_file, "w") as f:
        f.write(tag)
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
