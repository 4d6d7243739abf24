import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
from goslam_b200 import neus, synthetic
from oracle import neus_oracle
dev = torch.device("cuda:0")
offs, ress, _, total = neus.hashgrid_layout()
w = synthetic.make_neus_weights(seed=7, total_grid_params=total, layout=(offs, ress))
bound = [[-2.0, 2.0]] * 3
net = neus.InstantNeuS(synthetic.NEUS_CFG, bound)
with torch.no_grad():
    net.sdf_network.encoding.encoding.params.copy_(w["grid"])
    net.sdf_network.sdf_layer.weight.copy_(w["sdf_w"])
    net.sdf_network.sdf_layer.bias.copy_(w["sdf_b"])
    net.color_network._B.copy_(w["color_B"])
    net.color_network.network.params.copy_(w["mlp"])
net = net.to(dev)
rt = torch.tensor([[-1.8, 1.9], [-2.0, 2.0], [-1.5, 2.0]])
net.update_bound(rt)
R = 256
ro, rd, zv, ds = synthetic.make_rays(R, S=72, seed=11)
out = net(ro.to(dev), rd.to(dev), zv.to(dev), ds.to(dev))
ref = neus_oracle.forward(w["grid"].half().numpy(), w["sdf_w"].numpy(), w["sdf_b"].numpy(), w["color_B"].numpy(),
                          w["mlp"].half().numpy(), np.array(bound, np.float32), rt.numpy(), 0.2, 10.0,
                          ro.numpy(), rd.numpy(), zv.numpy(), ds.numpy())
got = out["sdf"].cpu().numpy()
d = np.abs(got - ref["sdf"])
print("max abs", d.max(), "n>1e-4", (d > 1e-4).sum(), "n>1e-3", (d > 1e-3).sum(), "of", d.size)
idx = np.argsort(-d.reshape(-1))[:12]
zm = ref["z_vals"].reshape(-1)
pts = (ro.numpy()[:, None, :] + rd.numpy()[:, None, :] * ref["z_vals"][:, :, None]).reshape(-1, 3)
for i in idx:
    raw = (pts[i] - (-2.0)) / 4.0 * 2 - 1
    print(i // 72, i % 72, "got", got.reshape(-1)[i], "ref", ref["sdf"].reshape(-1)[i], "pt", pts[i], "raw", raw)
# per-level check: zero all levels but one in the SDF head to localise
for lvl in [None] + list(range(16)):
    sw = w["sdf_w"].clone()
    if lvl is not None:
        keep = torch.zeros(35, dtype=torch.bool); keep[:3] = True; keep[3 + 2 * lvl: 5 + 2 * lvl] = True
        sw[:, ~keep] = 0
    with torch.no_grad():
        net.sdf_network.sdf_layer.weight.copy_(sw.to(dev))
    o2 = net(ro.to(dev), rd.to(dev), zv.to(dev), ds.to(dev))["sdf"].cpu().numpy()
    r2 = neus_oracle.forward(w["grid"].half().numpy(), sw.numpy(), w["sdf_b"].numpy(), w["color_B"].numpy(),
                             w["mlp"].half().numpy(), np.array(bound, np.float32), rt.numpy(), 0.2, 10.0,
                             ro.numpy(), rd.numpy(), zv.numpy(), ds.numpy())["sdf"]
    dd = np.abs(o2 - r2)
    print("level", lvl, "max abs", dd.max(), "n>1e-5", (dd > 1e-5).sum())
