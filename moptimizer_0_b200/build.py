"""In-tree build of the CUDA library (sm_100a only) and, later, the C++ host mirror.

    python -m moptimizer_0_b200.build [--force]

Produces moptimizer_0_b200/libmopt_b200.so with plain nvcc (cross-compiles without a GPU).  The
.so is git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(PKG, "_build")
LIB = os.path.join(PKG, "libmopt_b200.so")

CU_SOURCES = ["mopt_capi.cu", "mopt_store.cu", "mopt_pass_p2p.cu", "mopt_pass_dense.cu", "mopt_pass_wide.cu",
              "mopt_ingest.cu", "mopt_icp.cu"]
HEADERS = ["mopt_common.cuh", "mopt_setup.cuh", "mopt_models.cuh", "mopt_pass.cuh", "mopt_lm.cuh",
           "mopt_internal.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "-I", os.path.join(ROOT, "include"), "-I", CSRC,
]


def _mtime(p):
    return os.path.getmtime(p) if os.path.exists(p) else 0.0


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _compile(src, verbose):
    obj = os.path.join(OBJ, os.path.splitext(src)[0] + ".o")
    cmd = [_nvcc(), *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return obj, r.stderr


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    deps = [os.path.join(CSRC, h) for h in HEADERS] + [os.path.join(ROOT, "include", "mopt_capi.h"),
                                                      os.path.abspath(__file__)]
    newest_hdr = max(_mtime(d) for d in deps)
    todo = []
    for src in CU_SOURCES:
        obj = os.path.join(OBJ, os.path.splitext(src)[0] + ".o")
        if force or _mtime(obj) < max(newest_hdr, _mtime(os.path.join(CSRC, src))):
            todo.append(src)
    logs = []
    if todo:
        with cf.ThreadPoolExecutor(max_workers=min(len(todo), os.cpu_count() or 4)) as ex:
            for obj, log in ex.map(lambda s: _compile(s, verbose), todo):
                logs.append(log)
    objs = [os.path.join(OBJ, os.path.splitext(s)[0] + ".o") for s in CU_SOURCES]
    if todo or not os.path.exists(LIB) or _mtime(LIB) < max(_mtime(o) for o in objs):
        cmd = [_nvcc(), "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
               "-Xcompiler", "-fPIC", "-lcudart", "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    if verbose:
        print("\n".join(logs))
    return LIB


def build_cpp_tests(force: bool = False) -> str:
    """Host-only g++ build of the C++ API mirror tests (tests/cpp) against libmopt_b200.so."""
    build()
    tdir = os.path.join(ROOT, "tests", "cpp")
    exe = os.path.join(tdir, "run_tests")
    srcs = [os.path.join(tdir, f) for f in ("test_main.cpp", "reference_tests.cpp")]
    deps = srcs + [os.path.join(tdir, f) for f in ("mini_test.h", "fixtures.h")] + [LIB]
    for d, _, files in os.walk(os.path.join(ROOT, "include")):
        deps += [os.path.join(d, f) for f in files]
    if force or _mtime(exe) < max(_mtime(d) for d in deps):
        cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"),
               '-DTEST_DATA_PATH="tests/golden"', *srcs, "-o", exe, "-L", PKG, "-lmopt_b200",
               "-Wl,-rpath,$ORIGIN/../../moptimizer_0_b200"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"g++ failed for tests/cpp:\n{r.stdout}\n{r.stderr}")
    return exe


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
