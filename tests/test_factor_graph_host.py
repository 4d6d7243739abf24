"""CPU test of the HOST LOGIC of goslam_b200.FactorGraph (edge bookkeeping, slot pool, host mirrors, inactive-edge window,
age eviction, rm_keyframe shifting, 13-frame chunking of update_lowmem): the scenario of tests/tools/fg_scenario.py with
every kernel call replaced by the CPU oracle (TEST doubles, monkeypatched here — the product has no CPU path), compared
with the golden produced by the REFERENCE FactorGraph / DepthVideo (tests/golden/factor_graph.npz).
Edge lists must be bit-exact; the float state follows the oracle's arithmetic and agrees to 1e-4."""
import os
import sys
import types

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "tools"))

from oracle import ba_oracle, corr_oracle, geom_oracle, graph_oracle, upsample_oracle  # noqa: E402


class _CpuVideo:
    """DepthVideo stand-in on CPU tensors; geometry / BA / upsampling by the oracle"""
    takes_frame_eta = True

    def __init__(self, inputs, buffer, num_kf, ht8, wd8):
        self.ht, self.wd, self.stereo = 8 * ht8, 8 * wd8, False
        self.counter = types.SimpleNamespace(value=num_kf)
        z = lambda *s, dt=torch.float: torch.zeros(buffer, *s, dtype=dt)   # noqa: E731
        self.timestamp, self.images, self.dirty, self.red = z(), z(3, 1, 1), z(dt=torch.bool), z(dt=torch.bool)
        self.poses_gt, self.depths_gt, self.disps_up = z(4, 4), z(1, 1), z(self.ht, self.wd)
        self.poses_filtered, self.disps_filtered, self.mask_filtered, self.update_priority = z(7), z(1, 1), z(1, 1), z()
        for k, v in inputs.items():
            setattr(self, k, v.clone())

    def get_lock(self):
        import contextlib
        return contextlib.nullcontext()

    def reproject(self, ii, jj):
        c, v = geom_oracle.reproject(self.poses.numpy(), self.disps.numpy(), self.intrinsics.numpy(),
                                     torch.as_tensor(ii).numpy(), torch.as_tensor(jj).numpy())
        return torch.from_numpy(c), torch.from_numpy(v)

    def distance(self, ii, jj, beta=0.3, bidirectional=True):
        ii, jj = torch.as_tensor(ii).reshape(-1).numpy(), torch.as_tensor(jj).reshape(-1).numpy()
        f = lambda a, b: geom_oracle.frame_distance(self.poses.numpy(), self.disps.numpy(), self.intrinsics[0].numpy(), a, b, beta)   # noqa: E731
        return torch.from_numpy(0.5 * (f(ii, jj) + f(jj, ii)) if bidirectional else f(ii, jj))

    def ba(self, target, weight, eta, ii, jj, t0=1, t1=None, iters=2, lm=1e-4, ep=0.1, motion_only=False, ba_type=None,
           eta_by_frame=False):
        rp, rd, _, _, st = ba_oracle.ba(self.poses.numpy(), self.disps.numpy(), self.intrinsics[0].numpy(), self.disps_sens.numpy(),
                                        target.numpy(), weight.numpy(), eta.numpy(), ii.numpy(), jj.numpy(), t0, t1, iters, lm, ep,
                                        motion_only)
        assert list(st) == [0] * iters
        self.poses.copy_(torch.from_numpy(rp))
        self.disps.copy_(torch.from_numpy(rd).clamp(min=0.001))

    def upsample(self, ix, mask):
        self.disps_up[ix] = upsample_oracle.cvx_upsample(self.disps[ix].unsqueeze(-1), mask.float()).squeeze(-1)


class _CpuCorr:
    """CorrBlock stand-in with the slot-pool interface FactorGraph uses (from_video / cat / [mask] / __call__)"""

    def __init__(self, pyr, pool, slots):
        self.pyr, self.pool, self._slots_host = pyr, pool, slots

    @classmethod
    def from_video(cls, fmaps, ii, jj, ht, wd, rig=1, pool=None):
        slots = pool.alloc(int(ii.numel()))
        c = (ii == jj).long() if rig > 1 else torch.zeros_like(ii)
        pyr = corr_oracle.corr_build(fmaps[ii, 0], fmaps[jj, c], 4)
        return cls([p.clone() for p in pyr], pool, slots)

    def cat(self, other):
        self.pyr = [torch.cat([a, b]) for a, b in zip(self.pyr, other.pyr)]
        self._slots_host += other._slots_host
        return self

    def __getitem__(self, keep):
        keep = torch.as_tensor(keep)
        self.pool.release(s for s, k in zip(self._slots_host, keep.tolist()) if not k)
        self._slots_host = [s for s, k in zip(self._slots_host, keep.tolist()) if k]
        self.pyr = [p[keep] for p in self.pyr]
        return self

    def __call__(self, coords):
        out = corr_oracle.corr_pyramid_lookup([p.numpy() for p in self.pyr], coords[0].numpy(), 3)
        return torch.from_numpy(out)[None]


def _reproject_motion(poses, disps, intr, ii, jj, target):
    c, _ = geom_oracle.reproject(poses.numpy(), disps.numpy(), intr.numpy(), ii.numpy(), jj.numpy())
    coords = torch.from_numpy(c)
    ht, wd = disps.shape[1:]
    y, x = torch.meshgrid(torch.arange(ht).float(), torch.arange(wd).float(), indexing="ij")
    c0 = torch.stack([x, y], -1)
    motion = torch.cat([coords - c0, target.view_as(coords) - coords], dim=-1).permute(0, 1, 4, 2, 3).clamp(-64.0, 64.0)
    return coords, motion.contiguous()


class _CpuAltCorr:
    def __init__(self, fmaps, num_levels=4, radius=3):
        from goslam_b200.modules.corr import AltCorrBlock
        self.inner = AltCorrBlock.__new__(AltCorrBlock)
        AltCorrBlock.__init__(self.inner, fmaps, num_levels, radius)

    def __call__(self, coords, ii, jj):
        outs = []
        for lvl, maps in enumerate(self.inner.pyramid):
            o = corr_oracle.altcorr_forward(self.inner.pyramid[0][0, ii].float().numpy(), maps[0, jj].float().numpy(),
                                            (coords[0] / 2 ** lvl).unsqueeze(1).numpy(), 3)
            outs.append(torch.from_numpy(o[:, 0]))
        return torch.cat(outs, dim=1)[None]


@pytest.fixture(scope="module")
def run():
    import fg_scenario
    import goslam_b200.factor_graph as fg
    saved = {k: getattr(fg, k) for k in ("CorrBlock", "AltCorrBlock", "fmaps_to_kmajor")}
    saved_rm, saved_prox = fg.droid_backends.reproject_motion, fg.graph_ops.proximity_edges
    saved_km = fg.FactorGraph._kmajor_rows
    fg.CorrBlock, fg.AltCorrBlock = _CpuCorr, _CpuAltCorr
    fg.droid_backends.reproject_motion = _reproject_motion
    def prox(d, t0, t1, t, rad, nms, thresh, maxf, stereo, io, jo):
        es = graph_oracle.proximity_edges(d.numpy(), t0, t1, t, rad, nms, thresh, maxf, stereo, io.numpy(), jo.numpy())
        return torch.from_numpy(np.ascontiguousarray(es[:, 0])), torch.from_numpy(np.ascontiguousarray(es[:, 1]))
    fg.graph_ops.proximity_edges = prox
    fg.FactorGraph._kmajor_rows = lambda self, frames: (self.video.fmaps, self.video.fmaps.shape[1])   # the stand-in builds from fmaps
    try:
        video = _CpuVideo(fg_scenario.make_inputs(), fg_scenario.BUFFER, fg_scenario.NUM_KF, fg_scenario.HT8, fg_scenario.WD8)
        got = fg_scenario.run(fg.FactorGraph, video, "cpu")
    finally:
        for k, v in saved.items():
            setattr(fg, k, v)
        fg.droid_backends.reproject_motion, fg.graph_ops.proximity_edges = saved_rm, saved_prox
        fg.FactorGraph._kmajor_rows = saved_km
    return got, np.load(os.path.join(HERE, "golden", "factor_graph.npz"))


def test_edge_bookkeeping_bit_exact_on_cpu(run):
    got, want = run
    assert int(got["n_steps"]) == int(want["n_steps"]) == 18
    for step in range(18):
        for f in ("ii", "jj", "age", "ii_inac", "jj_inac", "ii_bad", "jj_bad"):
            k = "s%02d_%s" % (step, f)
            np.testing.assert_array_equal(got[k], want[k], err_msg=k)


def test_state_follows_the_reference_on_cpu(run):
    got, want = run
    for step in range(18):
        for f in ("poses", "disps", "target", "weight", "damping", "disps_up"):
            k = "s%02d_%s" % (step, f)
            a, b = got[k].astype(np.float64), want[k].astype(np.float64)
            assert a.shape == b.shape, k
            if b.size:
                assert np.abs(a - b).max() / max(np.abs(b).max(), 1e-12) < 1e-4, k
