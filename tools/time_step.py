"""in-stream CUDA-event timing of each phase of the config-2 update step."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from goslam_b200 import droid_backends
from goslam_b200.modules import CorrBlock
from goslam_b200.modules.corr import fmaps_to_kmajor

dev = torch.device("cuda:0")
sc = bench.make_window(43)
win = bench.Window(sc, dev)
d = win.d
ii, jj = d["ii"], d["jj"]


def t(fn, n=50, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / n


km = fmaps_to_kmajor(d["fmaps"][:bench.NUM_KF])
corr = win.build(km)
coords, _ = droid_backends.reproject(d["poses"], d["disps"], d["intrinsics"], ii, jj, want_valid=False)


def ba(iters, motion_only=False):
    d["poses"].copy_(win.poses0); d["disps"].copy_(win.disps0)
    droid_backends.ba(d["poses"], d["disps"], d["intrinsics"][0], d["disps_sens"], d["targets"], d["weights"],
                      d["eta"], ii, jj, 1, bench.NUM_KF, iters, 1e-4, 0.1, motion_only)


print("kmajor(8 frames)      %8.1f us" % t(lambda: fmaps_to_kmajor(d["fmaps"][:bench.NUM_KF])))
print("corr build (36 edges) %8.1f us" % t(lambda: win.build(km)))
print("reproject             %8.1f us" % t(lambda: droid_backends.reproject(d["poses"], d["disps"], d["intrinsics"], ii, jj, want_valid=False)))
print("lookup (4 levels)     %8.1f us" % t(lambda: win.corr(coords)))
print("state reset (2 copies)%8.1f us" % t(lambda: (d["poses"].copy_(win.poses0), d["disps"].copy_(win.disps0))))
for it in (1, 2, 3):
    print("ba iters=%d            %8.1f us" % (it, t(lambda: ba(it))))
print("ba motion_only it=3   %8.1f us" % t(lambda: ba(3, True)))
print("frame_distance 64 prs %8.1f us" % t(lambda: droid_backends.frame_distance(d["poses"], d["disps"], d["intrinsics"][0].contiguous(), ii.repeat(2)[:64].contiguous(), jj.repeat(2)[:64].contiguous(), 0.3)))
print("full step             %8.1f us" % t(win.step))
