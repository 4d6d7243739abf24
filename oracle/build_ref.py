"""Build oracle/_ref/goslam_ref_kernels*.so: the reference's OWN CUDA kernels compiled for
sm_100a, as a GPU-side checker for the parity tests (TEST INFRASTRUCTURE).

Sources are read where they lie under /root/reference and patched in a temp directory — never
copied into the repo:
  * src/lib/correlation_kernels.cu, src/lib/altcorr_kernel.cu: `X.type()` -> `X.scalar_type()`
    at the AT_DISPATCH sites (torch >= 2.x API; 3 one-token edits);
  * src/lib/droid_kernels.cu: lines 1-14, 24-1116 and 1436-1541 (everything except the
    Eigen-dependent SparseBlock / schur_block / ba_cuda, Eigen being absent from the snapshot);
  * oracle/ref_binding.cu (ours) includes the slice and adds pybind wrappers.
Output only into oracle/_ref/ (git-ignored; travels to the GPU box with gpurun).
"""
import glob
import os
import re
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/src/lib"
OUT = os.path.join(HERE, "_ref")
NAME = "goslam_ref_kernels"


def existing():
    hits = glob.glob(os.path.join(OUT, NAME + "*.so"))
    return hits[0] if hits else None


def main(force=False):
    if existing() and not force:
        print("[build_ref] up to date:", existing())
        return existing()
    if not os.path.isdir(REF):
        raise SystemExit("[build_ref] %s not present (fine on the GPU box: prebuilt .so is used)" % REF)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    os.environ.setdefault("MAX_JOBS", "4")
    from torch.utils.cpp_extension import load
    tmp = tempfile.mkdtemp(prefix="goslam_ref_")
    for fn in ("correlation_kernels.cu", "altcorr_kernel.cu"):
        src = open(os.path.join(REF, fn)).read()
        src = re.sub(r"(\w+)\.type\(\)", r"\1.scalar_type()", src)
        open(os.path.join(tmp, fn), "w").write(src)
    lines = open(os.path.join(REF, "droid_kernels.cu")).read().split("\n")
    keep = lines[0:14] + lines[23:1116] + lines[1435:1541]
    open(os.path.join(tmp, "droid_kernels_slice.cuh"), "w").write("\n".join(keep) + "\n")
    shutil.copy(os.path.join(HERE, "ref_binding.cu"), os.path.join(tmp, "ref_binding.cu"))
    os.makedirs(OUT, exist_ok=True)
    build_dir = os.path.join(tmp, "build")
    os.makedirs(build_dir)
    load(name=NAME,
         sources=[os.path.join(tmp, f) for f in ("ref_binding.cu", "correlation_kernels.cu", "altcorr_kernel.cu")],
         extra_cuda_cflags=["-O3", "-gencode=arch=compute_100a,code=sm_100a",
                            '-DREF_SLICE=\\"%s\\"' % os.path.join(tmp, "droid_kernels_slice.cuh")],
         extra_cflags=["-O2"], build_directory=build_dir, verbose=False, is_python_module=False)
    so = glob.glob(os.path.join(build_dir, NAME + "*.so"))[0]
    dst = os.path.join(OUT, os.path.basename(so))
    shutil.copy(so, dst)
    shutil.rmtree(tmp, ignore_errors=True)
    print("[build_ref] built", dst)
    return dst


def load_ref():
    """import the prebuilt module (GPU box) — returns None when it was never built."""
    so = existing()
    if so is None:
        return None
    import importlib.util
    import torch  # noqa: F401  (libtorch must be loaded first)
    spec = importlib.util.spec_from_file_location(NAME, so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    main(force="--force" in sys.argv)
