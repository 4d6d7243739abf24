"""Tiny driver for ncu captures: runs each hot-path kernel a few times on config-2 shapes.
usage: python tools/profile_driver.py [build|lookup|ba|neus|all] [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    dev = torch.device("cuda:0")
    from goslam_b200 import droid_backends
    from goslam_b200.modules import CorrBlock
    sc = bench.make_window(43)
    win = bench.Window(sc, dev)
    d = win.d
    ii, jj = d["ii"], d["jj"]
    from goslam_b200.modules.corr import fmaps_to_kmajor
    km = fmaps_to_kmajor(d["fmaps"][:bench.NUM_KF])
    corr = win.build(km)                       # tiled slot pool, video-level indexed build (the bench path)
    coords, _ = droid_backends.reproject(d["poses"], d["disps"], d["intrinsics"], ii, jj, want_valid=False)
    torch.cuda.synchronize()
    for _ in range(reps):
        if what in ("build", "all"):
            corr = win.build(km)
        if what in ("lookup", "all"):
            corr(coords)
        if what in ("ba", "all"):
            d["poses"].copy_(win.poses0)
            d["disps"].copy_(win.disps0)
            droid_backends.ba(d["poses"], d["disps"], d["intrinsics"][0], d["disps_sens"], d["targets"], d["weights"],
                              d["eta"], ii, jj, 1, bench.NUM_KF, bench.BA_ITERS, 1e-4, 0.1, False)
    if what in ("neus", "all"):
        bench.RAYS = 1 << 16
        net, rays, _ = bench.make_renderer(dev, 43)
        rd = [r.to(dev) for r in rays]
        for _ in range(reps):
            net(*rd)
    torch.cuda.synchronize()
    print("done", what)


if __name__ == "__main__":
    main()
