"""Build an A/B variant of libgoslam_b200.so with extra nvcc flags (probes, experiment macros) next to the
shipped one:   python tools/build_variant.py probe -DGOSLAM_BA_PROBE      -> go-slam_b200/_build/variant_probe/libgoslam_b200.so
Tools load it with  use_variant("probe")  BEFORE the first goslam_b200 call.  Never used by tests or bench."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def variant_path(name):
    return os.path.join(ROOT, "go-slam_b200", "_build", "variant_" + name, "libgoslam_b200.so")


def build(name, flags):
    from goslam_b200 import build as b
    out = os.path.dirname(variant_path(name))
    os.makedirs(out, exist_ok=True)
    objs, procs = [], []
    for src in b.sources():
        obj = os.path.join(out, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        procs.append(subprocess.Popen([b._nvcc()] + b.NVCC_FLAGS + list(flags) + ["-c", src, "-o", obj]))
    for p in procs:
        if p.wait() != 0:
            raise SystemExit("nvcc failed")
    subprocess.check_call([b._nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", variant_path(name)] + objs)
    return variant_path(name)


def use_variant(name):
    from goslam_b200 import _lib, build as b
    assert _lib._LIB is None, "use_variant() must run before the library is loaded"
    b.LIB = variant_path(name)
    assert os.path.exists(b.LIB), "build it first: python tools/build_variant.py %s <flags>" % name


if __name__ == "__main__":
    print(build(sys.argv[1], sys.argv[2:]))
