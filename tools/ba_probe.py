"""one config-2 BA call on a library built with -DGOSLAM_BA_PROBE (prints per-phase cycle counts)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
from build_variant import use_variant
use_variant("probe")          # python tools/build_variant.py probe -DGOSLAM_BA_PROBE
import bench
from goslam_b200 import droid_backends

dev = torch.device("cuda:0")
win = bench.Window(bench.make_window(43), dev)
d = win.d
for rep in range(3):
    d["poses"].copy_(win.poses0); d["disps"].copy_(win.disps0)
    torch.cuda.synchronize()
    print("--- call", rep, flush=True)
    droid_backends.ba(d["poses"], d["disps"], d["intrinsics"][0], d["disps_sens"], d["targets"], d["weights"],
                      d["eta"], d["ii"], d["jj"], 1, bench.NUM_KF, 3, 1e-4, 0.1, False)
    torch.cuda.synchronize()
