"""Builds libfemb200.so (hand-written CUDA for sm_100a + the C ABI + the C++ host mirror) in-tree.

    python go-fem_b200/build.py [--force] [--verbose]

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU box with the snapshot.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
HOST = os.path.join(HERE, "host")
OUT = os.path.join(HERE, "libfemb200.so")
OBJ = os.path.join(HERE, "build")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CUFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
           "-Xptxas", "-v", "-DFEMB200_BUILD"]
CU_SOURCES = ["api.cu", "symbolic.cu", "numeric.cu", "srows.cu", "keblocks.cu", "csr_ops.cu", "dist.cu", "mesh.cu"]
CXX_SOURCES = ["host_capi.cpp"]


def _deps():
    deps = []
    for d in (CSRC, HOST, os.path.join(HERE, "..", "include")):
        if os.path.isdir(d):
            deps += [os.path.join(d, f) for f in os.listdir(d)]
    return deps


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _run(cmd, log):
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + p.stdout)
    if p.returncode != 0:
        sys.stderr.write(p.stdout)
        raise RuntimeError("build failed: " + " ".join(cmd))
    return p.stdout


def build(force=False, verbose=False):
    deps = _deps()
    if not force and not _stale(OUT, deps):
        return OUT
    os.makedirs(OBJ, exist_ok=True)
    jobs = []
    for s in CU_SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ, s + ".o")
        if force or _stale(obj, deps):
            jobs.append(([NVCC] + ARCH + CUFLAGS + ["-c", src, "-o", obj], os.path.join(OBJ, s + ".log")))
    for s in CXX_SOURCES:
        src = os.path.join(HOST, s)
        if not os.path.exists(src):
            continue
        obj = os.path.join(OBJ, s + ".o")
        if force or _stale(obj, deps):
            jobs.append((["g++", "-O2", "-std=c++17", "-fPIC", "-Wall", "-c", src, "-o", obj], os.path.join(OBJ, s + ".log")))
    with ThreadPoolExecutor(max_workers=4) as ex:
        outs = list(ex.map(lambda j: _run(*j), jobs))
    if verbose:
        for o in outs:
            print(o)
    objs = [os.path.join(OBJ, s + ".o") for s in CU_SOURCES]
    objs += [os.path.join(OBJ, s + ".o") for s in CXX_SOURCES if os.path.exists(os.path.join(HOST, s))]
    _run([NVCC] + ARCH + ["-shared", "-o", OUT] + objs + ["-cudart", "static", "-lpthread"], os.path.join(OBJ, "link.log"))
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
