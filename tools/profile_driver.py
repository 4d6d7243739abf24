"""Tiny driver for ncu captures: runs each hot-path kernel a few times on config-2 shapes.
usage: python tools/profile_driver.py [build|build60|lookup|ba|balarge|neus|all] [reps]"""
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
    if what == "build60":                          # north_star headline shape (640x480 -> 60x80)
        sc60 = bench.make_window(143, ht=60, wd=80)
        win60 = bench.Window(sc60, dev)
        km60 = fmaps_to_kmajor(win60.d["fmaps"][:win60.num_kf])
        for _ in range(reps):
            win60.build(km60)
    if what == "balarge":                          # config-4-sized global BA: prep, linearise, system, cluster solve, back-substitution
        from goslam_b200 import synthetic
        num_kf, ht, wd = 64, 30, 40
        s4, g4 = synthetic.make_scene(num_kf=num_kf, ht=ht, wd=wd, rgbd=True, seed=43, with_fmaps=False, buffer=num_kf + 2)
        s4["t0"], s4["t1"] = 1, num_kf
        tg, wg, eta = synthetic.make_update(s4, synthetic.true_reprojection(s4)[0], g4, noise=0.7)
        D4 = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in s4.items()}
        tg, wg, eta = tg.to(dev), wg.to(dev), eta.to(dev)
        for _ in range(reps):
            p4, d4 = D4["poses"].clone(), D4["disps"].clone()
            droid_backends.ba(p4, d4, D4["intrinsics"][0].contiguous(), D4["disps_sens"], tg, wg, eta, D4["ii"], D4["jj"], 1, num_kf,
                              2, 1e-5, 1e-2, False)
    if what == "mapping":                          # renderer backward: forward under grad + loss + backward
        from goslam_b200 import synthetic
        net, _, _ = bench.make_renderer(dev, 43)
        R = 1 << 14
        ro, rd_, zv, ds = [t.to(dev) for t in synthetic.make_rays(R, S=bench.SAMPLES, seed=47)]
        gg = torch.Generator().manual_seed(5)
        rc = torch.rand(R, 3, generator=gg).to(dev)
        depth = (0.5 + 2.5 * torch.rand(R, 1, generator=gg)).to(dev)
        for _ in range(reps):
            for prm in net.parameters():
                prm.grad = None
            with torch.enable_grad():
                o = net(ro, rd_, zv, ds)
                sl, spl = net.compute_sdf_error(sdf=o["sdf"], z_vals=o["z_vals"], gt_depth=depth)
                total = 2.0 * torch.abs(o["color"] - rc).mean() + torch.abs(o["depth"] - depth).mean() + 2.0 * (sl + spl) + 0.1 * o["gradient_error"].mean()
            total.backward()
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
