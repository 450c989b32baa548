"""In-tree nvcc build of the C-ABI library (sm_100a only).

``python -m viloca_b200.build`` or ``viloca_b200.build.build()``.  The resulting
``viloca_b200/lib/libviloca_b200.so`` is git-ignored but travels with the tree to the GPU box.
nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libviloca_b200.so")
SOURCES = ["vl_kernels.cu", "vl_api.cu", "vl_prepare.cu"]
HEADERS = ["vl_device.cuh", "vl_cavi.cuh", "vl_launch.h", os.path.join("..", "..", "include", "viloca_b200.h")]
# vl_kernels.cu: vl_chunk_kernel runs many dependent updates in one launch, so state written by
# one CTA is read by later CTAs of the same launch, possibly on an SM whose L1 still holds the old
# line: global loads of that file are cached in L2 only (the operand streams through TMA anyway).
# (not --split-compile: it compiles the 64 kernel instantiations of that file on all host cores, 6 min ->
# 1.5 min, but the code it produces ran 13 % slower -- hiv5k 590 -> 669 ms, amplicon 2.64 -> 2.87 s, same box;
# VL_SPLIT_COMPILE=1 turns it on for development builds)
EXTRA_FLAGS = {"vl_kernels.cu": ["-Xptxas", "-dlcm=cg"] + (["--split-compile", "0"] if os.environ.get("VL_SPLIT_COMPILE") else []),
               # host arithmetic that must match NumPy bit for bit (running mean of the qualities): no FMA contraction
               "vl_prepare.cu": ["-Xcompiler", "-ffp-contract=off"]}
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, profile: bool = False, defines=(), tag: str = "") -> str:
    """``profile=True`` compiles the per-phase cycle counters in (vl_debug_counters) and builds
    ``libviloca_b200_prof.so`` next to the product library (select it with VILOCA_B200_LIB);
    ``defines`` / ``tag`` build an experimental variant ``libviloca_b200_<tag>.so`` the same way."""
    os.makedirs(LIB_DIR, exist_ok=True)
    headers = [os.path.normpath(os.path.join(CSRC, h)) for h in HEADERS]
    if profile and not tag:
        tag = "prof"
    suffix = ("_" + tag) if tag else ""
    lib_path = os.path.join(LIB_DIR, "libviloca_b200" + suffix + ".so")
    extra_defs = ["-D" + d for d in defines]
    objs, jobs = [], []
    for src in SOURCES:
        src_path = os.path.join(CSRC, src)
        obj = os.path.join(LIB_DIR, src.replace(".cu", suffix + ".o"))
        objs.append(obj)
        if force or _stale(obj, [src_path] + headers):
            cmd = [_nvcc()] + NVCC_FLAGS + (["-DVL_RS_PROFILE"] if profile else []) + extra_defs + EXTRA_FLAGS.get(src, []) + \
                ["-c", src_path, "-o", obj]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
                print(" ".join(cmd), file=sys.stderr)
            jobs.append((cmd, subprocess.Popen(cmd)))          # the translation units compile side by side
    for cmd, proc in jobs:
        if proc.wait() != 0:
            raise subprocess.CalledProcessError(proc.returncode, cmd)
    if force or _stale(lib_path, objs):
        cmd = [_nvcc(), "-shared", "-o", lib_path] + objs + ["-lcudart_static", "-lrt", "-lpthread", "-ldl",
                                                           "-Xlinker", "--no-undefined"]   # an unresolved symbol fails here, not at load time
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        subprocess.run(cmd, check=True)
    return lib_path


if __name__ == "__main__":
    _defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    _tag = next((a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--tag=")), "")
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, profile="--profile" in sys.argv,
                defines=_defs, tag=_tag))
