"""In-tree build of the CUDA extension (sm_100a) and the C-ABI shared library.

  python -m gatling_b200.build [--force]

nvcc cross-compiles without a GPU.  The result (gatling_b200/libgtlc.so) is git-ignored but travels to
the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libgtlc.so")
SOURCES = ["gtlc_api.cu", "bvh8_build.cpp"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",                      # explicit fmaf only: see device_math.cuh (bit-reproducible vs the CPU oracle)
    "-Xcompiler", "-fPIC,-fopenmp,-ffp-contract=off,-fvisibility=hidden,-O3",
    "-Xptxas", "-v",
] + ["-D%s=%s" % (k, os.environ[k]) for k in ("GTLC_TRACE_MIN_BLOCKS", "GTLC_SHADE_BIG_BLOCK", "GTLC_SHADE_PHASE_MASK") if os.environ.get(k)]


def _nvcc():
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def _host_cxx():
    return "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "gtlc.h"), __file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    cmd = [_nvcc(), "-ccbin", _host_cxx(), "-shared", "-o", LIB] + NVCC_FLAGS + [os.path.join(CSRC, s) for s in SOURCES] + ["-lgomp", "-ldl"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if res.returncode != 0:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed (see gatling_b200/build.log)")
    if verbose:
        print(log)
    return LIB


FACADE_LIB = os.path.join(HERE, "libgi.so")
FACADE_TEST = os.path.join(HERE, "..", "tests", "cpp", "gi_facade_main")


def build_facade(force=False):
    """C++ facade (gtl::gi* over the C-ABI) as libgi.so + the C++ test driver.  Plain g++: no CUDA in here."""
    inc = os.path.join(HERE, "..", "include")
    src = os.path.join(HERE, "gi_cpp", "Gi.cpp")
    main = os.path.join(HERE, "..", "tests", "cpp", "gi_facade_main.cpp")
    newest = max(os.path.getmtime(x) for x in (src, main, os.path.join(inc, "gtl", "gi", "Gi.h"), os.path.join(inc, "gtlc.h"), LIB))
    if force or not os.path.exists(FACADE_LIB) or os.path.getmtime(FACADE_LIB) < newest:
        subprocess.check_call([_host_cxx(), "-std=c++17", "-O2", "-fPIC", "-shared", "-I" + inc, src, "-o", FACADE_LIB,
                               "-L" + HERE, "-lgtlc", "-Wl,-rpath,$ORIGIN"])
    if force or not os.path.exists(FACADE_TEST) or os.path.getmtime(FACADE_TEST) < max(newest, os.path.getmtime(FACADE_LIB)):
        subprocess.check_call([_host_cxx(), "-std=c++17", "-O2", "-I" + inc, main, "-o", FACADE_TEST,
                               "-L" + HERE, "-lgi", "-lgtlc", "-Wl,-rpath," + HERE])
    return FACADE_LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print("built", LIB)
    print("built", build_facade(force="--force" in sys.argv))
