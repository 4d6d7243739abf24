"""diagnostic: distribution of kernel-vs-oracle errors of the composited renderer outputs per ray."""
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
R = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ro, rd, zv, ds = synthetic.make_rays(R, S=72, seed=11)
out = {k: v.cpu().numpy() for k, v in net(ro.to(dev), rd.to(dev), zv.to(dev), ds.to(dev), debug=True).items()}
args = (w["grid"].half().numpy(), w["sdf_w"].numpy(), w["sdf_b"].numpy(), w["color_B"].numpy(),
        w["mlp"].half().numpy(), np.array(bound, np.float32), rt.numpy(), 0.2, 10.0)
ref = neus_oracle.forward(*args, ro.numpy(), rd.numpy(), zv.numpy(), ds.numpy(), debug=True)
dbg = getattr(net, "last_debug", None)
for k in ("depth", "weight_sum", "normal", "depth_variance", "color"):
    scale = max(np.abs(ref[k]).max(), 1e-12)
    err = np.abs(out[k] - ref[k]).reshape(R, -1).max(1) / scale
    srt = np.sort(err)
    print("%-15s scale %.3g | median %.2e p90 %.2e p99 %.2e max %.2e | n>1e-4: %d  n>2e-4: %d  n>1e-3: %d of %d" % (
        k, scale, srt[R // 2], srt[int(.9 * R)], srt[int(.99 * R)], srt[-1], (err > 1e-4).sum(), (err > 2e-4).sum(),
        (err > 1e-3).sum(), R))
sd = np.abs(out["sdf"] - ref["sdf"])
inb = ref["sdf"] != 100
print("per-sample sdf: max abs %.3e (scale %.3g), n>1e-5 %d, n>1e-4 %d of %d" % (sd[inb].max(), np.abs(ref["sdf"][inb]).max(),
      (sd[inb] > 1e-5).sum(), (sd[inb] > 1e-4).sum(), inb.sum()))
if dbg is not None:
    a_k, g_k = dbg["alpha"].cpu().numpy(), dbg["grad"].cpu().numpy()
    a_o, g_o = ref["_alpha"], ref["_grad"].reshape(R, 72, 3)
    da = np.abs(a_k - a_o)
    dg = np.abs(g_k - g_o).max(-1)
    print("per-sample alpha: max abs %.3e n>1e-5 %d n>1e-4 %d n>1e-2 %d | grad: max abs %.3e (scale %.3g) n>1e-4 %d n>1e-2 %d" % (
        da.max(), (da > 1e-5).sum(), (da > 1e-4).sum(), (da > 1e-2).sum(), dg.max(), np.abs(g_o).max(), (dg > 1e-4).sum(), (dg > 1e-2).sum()))
    bad = np.argwhere(dg > 1e-3)[:10]
    for r, s_ in bad:
        print("  ray %d sample %d: grad kernel %s oracle %s  sdf %.5f/%.5f alpha %.5f/%.5f" % (r, s_, g_k[r, s_], g_o[r, s_],
              out["sdf"][r, s_], ref["sdf"][r, s_], a_k[r, s_], a_o[r, s_]))
    # how close do the flipped samples sit to a cell face of some level?
    metas, _ = neus_oracle.hashgrid_meta()
    zm = ref["z_vals"]
    bnd = np.array(bound, np.float32)
    for r, s_ in np.argwhere(dg > 1e-3)[:10]:
        P = (ro.numpy()[r] + rd.numpy()[r] * zm[r, s_]).astype(np.float32)
        raw = ((P - bnd[:, 0]) / (bnd[:, 1] - bnd[:, 0]) * np.float32(2) - np.float32(1)).astype(np.float32)
        x01 = ((np.clip(raw, -1, 1) + np.float32(1)) / np.float32(2)).astype(np.float32)
        best = (1e9, None)
        for l, m in enumerate(metas):
            pos = (x01.astype(np.float64) * np.float64(m["scale"]) + 0.5).astype(np.float32)
            fr = pos - np.floor(pos)
            for c in range(3):
                d_ulp = min(fr[c], 1 - fr[c]) / np.spacing(pos[c])
                if d_ulp < best[0]:
                    best = (d_ulp, (l, c, float(pos[c])))
        print("  ray %d sample %d: nearest cell face %.1f ulp away at (level, dim, pos) = %s" % (r, s_, best[0], best[1]))
    # are the normalised positions bit-identical?
    xk = dbg["pos"].cpu().numpy().reshape(-1, 3)[ref["_mask"]]
    xo = ref["_xn"]
    neq = (xk != xo)
    print("normalised positions: %d of %d components differ; max |diff| %.3e (ulp of 1.0 = 1.2e-7)" % (neq.sum(), neq.size, np.abs(xk - xo).max()))
    if neq.any():
        i = np.argwhere(neq)[0]
        print("  first differing: kernel %.9g oracle %.9g" % (xk[i[0], i[1]], xo[i[0], i[1]]))
