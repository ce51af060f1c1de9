"""Builds libloghisto_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libloghisto_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-O2",
    "--fmad=true",           # FP32 fast path may contract; the exact path uses __d*_rn intrinsics only
    "-Xptxas", "-v",
    "-shared",
]


HOST_LIB = os.path.join(HERE, "libloghisto_host.so")
HOST_SRC = os.path.join(HERE, "host", "metric_system.cc")
HOST_SRCS = [HOST_SRC, os.path.join(HERE, "host", "print_benchmark.cc")]


def needs_build_host() -> bool:
    if not os.path.exists(HOST_LIB):
        return True
    deps = HOST_SRCS + [os.path.join(HERE, "host", "metric_system.h"), os.path.join(ROOT, "include", "loghisto_b200.h"), LIB]
    return _newest([d for d in deps if os.path.exists(d)]) > os.path.getmtime(HOST_LIB)


def build_host(force: bool = False) -> str:
    """C++ mirror of MetricSystem (host glue above the C ABI); links against libloghisto_b200.so."""
    build()
    if not force and not needs_build_host():
        return HOST_LIB
    gxx = shutil.which("g++") or "g++"
    cmd = [gxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden", "-Wall", "-Wextra",
           "-I", os.path.join(ROOT, "include"), "-o", HOST_LIB] + HOST_SRCS + [
           "-L", HERE, "-lloghisto_b200", "-Wl,-rpath,$ORIGIN", "-lpthread"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("g++ failed building libloghisto_host.so")
    return HOST_LIB


def sources():
    return [os.path.join(CSRC, "lh_api.cu")]


def _newest(paths):
    return max(os.path.getmtime(p) for p in paths)


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "loghisto_b200.h")]
    return _newest(deps) > os.path.getmtime(LIB)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libloghisto_b200.so (there is no CPU fallback)")
    cmd = [nvcc] + NVCC_FLAGS + ["-I", os.path.join(ROOT, "include"), "-o", LIB] + sources()
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if res.returncode != 0:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed building libloghisto_b200.so")
    if verbose:
        print(log)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
    print(build_host(force="--force" in sys.argv))
