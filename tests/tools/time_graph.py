"""proximity-edge selection: device kernel vs the reference-shaped Python loops (oracle, CPU)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from goslam_b200 import graph
from oracle import graph_oracle

dev = torch.device("cuda:0")
rng = np.random.default_rng(5)
for name, c in [("frontend keyframe (t0=t-5, window 25)", dict(t0=45, t1=25, t=50, rad=2, nms=2, thresh=16.0, maxf=48, stereo=False, n_old=40)),
                ("initialisation (12 x 12)", dict(t0=0, t1=0, t=12, rad=2, nms=2, thresh=16.0, maxf=48, stereo=False, n_old=0)),
                ("global graph (200 x 200)", dict(t0=0, t1=0, t=200, rad=2, nms=2, thresh=20.0, maxf=1600, stereo=False, n_old=400))]:
    ilen, jlen = c["t"] - c["t0"], c["t"] - c["t1"]
    dist = (rng.random(ilen * jlen) * 60).astype(np.float32)
    old = rng.integers(0, c["t"], size=(c["n_old"], 2)).astype(np.int64)
    args = (c["t0"], c["t1"], c["t"], c["rad"], c["nms"], c["thresh"], c["maxf"], c["stereo"])
    t0 = time.perf_counter()
    want = graph_oracle.proximity_edges(dist, *args, old[:, 0], old[:, 1])
    cpu_ms = 1e3 * (time.perf_counter() - t0)
    dd, io, jo = torch.from_numpy(dist).to(dev), torch.from_numpy(old[:, 0].copy()).to(dev), torch.from_numpy(old[:, 1].copy()).to(dev)
    for _ in range(3):
        ii, jj = graph.proximity_edges(dd, *args, io, jo)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        ii, jj = graph.proximity_edges(dd, *args, io, jo)
    torch.cuda.synchronize()
    gpu_ms = 1e3 * (time.perf_counter() - t0) / 20
    ok = np.array_equal(torch.stack([ii, jj], 1).cpu().numpy(), want)
    print("%-40s %4d edges  device %.3f ms (incl. the one host sync)  numpy loops on the host %.2f ms  %s" %
          (name, len(want), gpu_ms, cpu_ms, "identical" if ok else "MISMATCH"))
