"""one config-4-sized BA call on the `probe` variant (tools/build_variant.py probe -DGOSLAM_BA_PROBE):
prints the cluster solve's per-phase cycle counts."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
from build_variant import use_variant
use_variant("probe")
from goslam_b200 import droid_backends, synthetic

dev = torch.device("cuda:0")
num_kf, ht, wd = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (64, 30, 40)
sc, g = synthetic.make_scene(num_kf=num_kf, ht=ht, wd=wd, rgbd=True, seed=43, with_fmaps=False, buffer=num_kf + 2)
sc["t0"], sc["t1"] = 1, num_kf
tg, wg, eta = synthetic.make_update(sc, synthetic.true_reprojection(sc)[0], g, noise=0.7)
D = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in sc.items()}
for rep in range(2):
    p, d = D["poses"].clone(), D["disps"].clone()
    torch.cuda.synchronize()
    print("--- call", rep, flush=True)
    droid_backends.ba(p, d, D["intrinsics"][0].contiguous(), D["disps_sens"], tg.to(dev), wg.to(dev), eta.to(dev), D["ii"], D["jj"],
                      1, num_kf, 1, 1e-5, 1e-2, False)
    torch.cuda.synchronize()
