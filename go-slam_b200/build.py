"""Build libgoslam_b200.so (sm_100a) in-tree with nvcc.

The library has no torch / python dependency: plain CUDA runtime (static cudart), C-ABI
declared in include/goslam_b200.h.  Objects go to go-slam_b200/_build/, the shared object
to go-slam_b200/libgoslam_b200.so (git-ignored, travels to the GPU box with gpurun).
"""
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libgoslam_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-DGOSLAM_SM_ARCH=100",
    "-I" + os.path.join(ROOT, "include"),
    "-I" + CSRC,
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (need CUDA 12.9 for sm_100a)")


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stamp(path, deps):
    h = hashlib.sha1()
    h.update(" ".join(NVCC_FLAGS).encode())
    for p in [path] + deps:
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def build(verbose=False, force=False):
    """Compile every .cu under csrc/ for sm_100a and link the shared library."""
    os.makedirs(BUILD, exist_ok=True)
    nvcc = _nvcc()
    headers = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h")))
    headers.append(os.path.join(ROOT, "include", "goslam_b200.h"))
    objs, procs, relink = [], [], force or not os.path.exists(LIB)
    for src in sources():
        obj = os.path.join(BUILD, os.path.basename(src)[:-3] + ".o")
        stamp_file = obj + ".stamp"
        stamp = _stamp(src, headers)
        objs.append(obj)
        if (not force and os.path.exists(obj) and os.path.exists(stamp_file)
                and open(stamp_file).read() == stamp):
            continue
        cmd = [nvcc] + NVCC_FLAGS + ["-c", src, "-o", obj]
        if verbose:
            print(" ".join(cmd), flush=True)
        procs.append((subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT),
                      src, stamp_file, stamp))
        relink = True
    for p, src, stamp_file, stamp in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(out.decode(errors="replace"))
            raise RuntimeError("nvcc failed on %s" % src)
        if verbose and out:
            print(out.decode(errors="replace"))
        with open(stamp_file, "w") as f:
            f.write(stamp)
    if relink:
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(verbose=True, force="--force" in sys.argv))
