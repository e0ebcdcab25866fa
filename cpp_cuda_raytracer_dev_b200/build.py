"""Build librtb.so (the C-ABI library of include/rtb.h) in-tree with nvcc for sm_100a.

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box with the
repository snapshot.  Device code is compiled with -fmad=false and the host code with
-ffp-contract=off: the path's decisions must be computed in uncontracted fp32.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "librtb.so")
DRIVER = os.path.join(HERE, "rtb_render")  # headless driver (csrc/rtb_render_main.cpp)
SOURCES = ["rtb_api.cu", "rtb_build.cu", "rtb_host.cpp"]
HEADERS = ["rtb_host.hpp", "rtb_kernels.cuh", "rtb_render.cuh", os.path.join("..", "..", "include", "rtb.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
              "-Xcompiler", "-fPIC,-O3,-ffp-contract=off,-fno-fast-math,-pthread", "-Xptxas", "-v", "-shared", "-cudart", "static"]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.exists(LIB) or not os.path.exists(DRIVER):
        return True
    t = min(os.path.getmtime(LIB), os.path.getmtime(DRIVER))
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS + ["rtb_framework.hpp", "rtb_render_main.cpp"]] + [__file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    cmd = [nvcc_path(), "-ccbin", "/usr/bin/g++"] + NVCC_FLAGS + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log = os.path.join(HERE, "build.log")
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed, see %s" % log)
    # the headless C++ driver on top of the C ABI (host code only)
    cmd = ["/usr/bin/g++", "-O2", "-std=c++17", "-Wall", os.path.join(CSRC, "rtb_render_main.cpp"), "-o", DRIVER, "-L", HERE, "-lrtb",
           "-Wl,-rpath,$ORIGIN"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    with open(log, "a") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("g++ failed for rtb_render, see %s" % log)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
