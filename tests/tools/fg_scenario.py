"""The factor-graph scenario of the drop-in test: one sequence of FactorGraph / DepthVideo calls, run
  * by tests/golden/make_golden.py on the REFERENCE classes (src/factor_graph.py, src/depth_video.py; CPU,
    natives stubbed by the oracle) -> tests/golden/factor_graph.npz, and
  * by tests/test_gpu_dropin.py on goslam_b200.FactorGraph / DepthVideo on the B200.
After every call the graph state is snapshotted: edge lists must match bit for bit, poses / disps / target /
weight / damping / disps_up within 1e-4 (north_star)."""
import numpy as np
import torch

from stub_update_op import update_op

NUM_KF, HT8, WD8, BUFFER = 9, 16, 24, 12


def make_inputs():
    """deterministic video content (CPU tensors): poses, disps, disps_sens, intrinsics, fmaps, nets, inps"""
    from goslam_b200 import synthetic
    sc, g = synthetic.make_scene(num_kf=NUM_KF, ht=HT8, wd=WD8, seed=43, rgbd=True, buffer=BUFFER, with_fmaps=True)
    nets = (0.5 * torch.randn(BUFFER, 128, HT8, WD8, generator=g)).half()
    inps = (0.5 * torch.randn(BUFFER, 128, HT8, WD8, generator=g)).half()
    # keep the perturbation small enough that BA stays in its basin over ~10 chained updates
    sc["poses"][1:NUM_KF, :3] += 0.005 * torch.randn(NUM_KF - 1, 3, generator=g)
    return dict(poses=sc["poses"], disps=sc["disps"], disps_sens=sc["disps_sens"], intrinsics=sc["intrinsics"],
                fmaps=sc["fmaps"], nets=nets, inps=inps)


def cfg_and_args(device):
    import types
    cfg = {"cam": {"H_out": 8 * HT8, "W_out": 8 * WD8}, "mode": "rgbd", "tracking": {"buffer": BUFFER}}
    return cfg, types.SimpleNamespace(device=device)


def fill_video(video, inputs):
    dev = video.poses.device
    for k, v in inputs.items():
        getattr(video, k)[:] = v.to(dev)
    video.counter.value = NUM_KF


def snapshot(graph, video, tag, out):
    def put(name, t):
        out["%s_%s" % (tag, name)] = t.detach().cpu().numpy().copy()
    for k in ("ii", "jj", "age", "ii_inac", "jj_inac", "ii_bad", "jj_bad"):
        put(k, getattr(graph, k))
    put("poses", video.poses[:NUM_KF])
    put("disps", video.disps[:NUM_KF])
    put("target", graph.target[0, :, ::3, ::3])
    put("weight", graph.weight[0, :, ::3, ::3])
    put("target_inac", graph.target_inac[0, :, ::3, ::3])
    put("damping", graph.damping[:NUM_KF, ::2, ::2])
    put("disps_up", video.disps_up[:NUM_KF, ::16, ::16])


def run(FactorGraph, video, device):
    """returns {tag_field: array}; tags are s00, s01, ... in call order"""
    out, step = {}, [0]

    def snap(g):
        snapshot(g, video, "s%02d" % step[0], out)
        step[0] += 1

    g = FactorGraph(video, update_op, device=device, corr_impl="volume", max_factors=18, upsample=True)
    g.add_neighborhood_factors(0, 5, r=2)                                       # s00: 14 edges among frames 0..4
    snap(g)
    g.update(1, use_inactive=True)                                              # s01
    snap(g)
    g.update(None, None, use_inactive=True)                                     # s02
    snap(g)
    g.add_proximity_factors(3, 0, rad=2, nms=1, beta=0.25, thresh=30.0, remove=True, max_t=7)   # s03
    snap(g)
    g.update(None, None, use_inactive=True)                                     # s04
    snap(g)
    g.rm_factors(g.age > 2, store=True)                                         # s05: the oldest edges become inactive
    snap(g)
    g.update(None, None, use_inactive=True)                                     # s06: inactive edges join BA
    snap(g)
    g.rm_keyframe(3)                                                            # s07
    video.counter.value -= 1
    snap(g)
    g.update(None, None, iters=3, use_inactive=True)                            # s08
    snap(g)
    # s09: duplicates (active and inactive) are filtered, the rest pushes the graph over max_factors -> age eviction
    ii = [7, 6, 7, 5, 2, 1, 6, 4, 7, 4, 5, 7, 6, 3, 7, 3]
    jj = [6, 7, 5, 7, 1, 2, 4, 6, 4, 7, 3, 3, 3, 6, 2, 7]
    g.add_factors(ii, jj, remove=True)
    snap(g)
    g.update(None, None, use_inactive=True, motion_only=True)                   # s10
    snap(g)
    g.filter_edges()                                                            # s11
    snap(g)
    g.update_fast(t0=2, t1=None, iters=2, steps=2, ba_type="loop")              # s12
    snap(g)
    g.clear_edges()                                                             # s13
    snap(g)

    # global-BA style graph: no volumes, windowed correlation, chunked update operator
    g2 = FactorGraph(video, update_op, device=device, corr_impl="alt", max_factors=80, upsample=False)
    t = video.counter.value
    ii, jj = torch.meshgrid(torch.arange(0, t), torch.arange(0, t), indexing="ij")
    keep = ((ii - jj).abs() > 0) & ((ii - jj).abs() <= 3)
    g2.add_factors(ii[keep], jj[keep], remove=True)                             # s14
    snap(g2)
    g2.update_lowmem(t0=1, t1=t, iters=2, steps=2, max_t=t, ba_type="dense")    # s15
    snap(g2)
    g2.update_lowmem(t0=2, t1=t, iters=2, steps=1, max_t=t, ba_type="loop", motion_only=True)   # s16
    snap(g2)
    g2.clear_edges()                                                            # s17
    snap(g2)
    out["n_steps"] = np.int64(step[0])
    return out
