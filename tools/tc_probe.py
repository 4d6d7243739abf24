"""one config-2 correlation build on a library built with -DGOSLAM_TC_PROBE."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from goslam_b200.modules.corr import fmaps_to_kmajor

dev = torch.device("cuda:0")
win = bench.Window(bench.make_window(43), dev)
km = fmaps_to_kmajor(win.d["fmaps"][:bench.NUM_KF])
for rep in range(2):
    win.build(km)
    torch.cuda.synchronize()
    print("---", flush=True)
