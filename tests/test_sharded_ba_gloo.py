"""world_size-2 gloo test (CPU) of the multi-GPU BA host logic (go-slam_b200/parallel.py):
partition by source frame -> local reduced systems -> ONE all-reduce -> redundant solve ->
owned-frame back-substitution -> re-replication.  The kernels are replaced by the oracle's
phase1/phase2 behind the same backend interface; the result must equal single-process BA."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleBackend:
    def __init__(self, poses, disps, intr, sens, t0, t1):
        self.poses, self.disps, self.intr, self.sens, self.t0, self.t1 = poses, disps, intr, sens, t0, t1

    def phase1(self, targets, weights, eta_by_frame, ii, jj, motion_only):
        from oracle import ba_oracle
        self.st = ba_oracle.phase1(self.poses.numpy(), self.disps.numpy(), self.intr.numpy(), self.sens.numpy(),
                                   targets.numpy(), weights.numpy(), eta_by_frame.numpy(), ii.numpy(), jj.numpy(),
                                   self.t0, self.t1, motion_only, dtype=np.float64)
        n = self.st["Hred"].shape[0]
        return torch.from_numpy(np.concatenate([self.st["Hred"].reshape(-1), self.st["bred"]]).copy()), n

    def phase2(self, system, lm, ep, motion_only, lo, hi):
        from oracle import ba_oracle
        n = self.st["Hred"].shape[0]
        H = system[:n * n].numpy().reshape(n, n)
        b = system[n * n:].numpy()
        dx, fail = ba_oracle.solve(H, b, H, lm, ep)
        p, d = self.poses.numpy(), self.disps.numpy()
        ba_oracle.phase2(self.st, dx, p, d, self.t0, self.t1, motion_only, lo, hi)
        return torch.from_numpy(dx)


class _Adapter(OracleBackend):
    def phase1(self, *a):
        t, _ = super().phase1(*a)
        return t


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from goslam_b200 import parallel, synthetic
    from oracle import geom_oracle
    sc, g = synthetic.make_scene(num_kf=7, ht=9, wd=12, seed=5, rgbd=True, with_fmaps=False)
    coords, _ = geom_oracle.reproject(sc["poses"].numpy(), sc["disps"].numpy(), sc["intrinsics"].numpy(),
                                      sc["ii"].numpy(), sc["jj"].numpy())
    tg, wg, eta = synthetic.make_update(sc, torch.from_numpy(coords[0]), g, noise=0.8)
    num, ht, wd = sc["disps"].shape
    kx = torch.unique(torch.cat([torch.arange(1, 7), sc["ii"]]))
    eta_f = torch.zeros(num, ht, wd)
    eta_f[kx] = eta
    be = _Adapter(sc["poses"].clone(), sc["disps"].clone(), sc["intrinsics"][0], sc["disps_sens"], 1, 7)
    dx = parallel.sharded_ba(be, be.disps, tg, wg, eta_f, sc["ii"], sc["jj"], 2, 1e-4, 0.1)
    if rank == 0:
        torch.save(dict(poses=be.poses, disps=be.disps, dx=dx, tg=tg, wg=wg, eta=eta, sc=sc), out)
    # every rank must hold the same replicated state
    ref = [torch.zeros_like(be.disps) for _ in range(world)]
    dist.all_gather(ref, be.disps)
    assert all(torch.equal(r, ref[0]) for r in ref)
    dist.destroy_process_group()


def test_shard_frames_by_edges_covers_and_balances():
    sys.path.insert(0, ROOT)
    from goslam_b200 import parallel, synthetic
    ii, jj = synthetic.neighborhood_edges(0, 64, 3)
    for world in (1, 2, 4, 8):
        b = parallel.shard_frames_by_edges(ii, 64, world)
        assert b[0][0] == 0 and b[-1][1] == 64 and all(b[i][1] == b[i + 1][0] for i in range(world - 1))
        loads = [int(((ii >= lo) & (ii < hi)).sum()) for lo, hi in b]
        assert sum(loads) == ii.numel() and max(loads) <= 1.35 * ii.numel() / world + 6


@pytest.mark.timeout(300)
def test_sharded_ba_two_ranks_equals_single_process(tmp_path):
    out = str(tmp_path / "r0.pt")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out, weights_only=False)
    from oracle import ba_oracle
    sc = got["sc"]
    rp, rd, rdx, _, st = ba_oracle.ba(sc["poses"].numpy(), sc["disps"].numpy(), sc["intrinsics"][0].numpy(),
                                      sc["disps_sens"].numpy(), got["tg"].numpy(), got["wg"].numpy(), got["eta"].numpy(),
                                      sc["ii"].numpy(), sc["jj"].numpy(), 1, 7, 2, 1e-4, 0.1, False, dtype=np.float64)
    assert st.tolist() == [0, 0]
    np.testing.assert_allclose(got["poses"].numpy(), rp, rtol=0, atol=2e-6)
    np.testing.assert_allclose(got["disps"].numpy(), rd, rtol=0, atol=2e-6)
    np.testing.assert_allclose(got["dx"].numpy(), rdx, rtol=0, atol=2e-6)


def _pairs_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from goslam_b200 import parallel
    from oracle import geom_oracle
    from goslam_b200 import synthetic
    sc, _ = synthetic.make_scene(num_kf=6, ht=8, wd=10, seed=11, rgbd=True, with_fmaps=False)
    ii, jj = torch.meshgrid(torch.arange(6), torch.arange(6), indexing="ij")
    ii, jj = ii.reshape(-1)[:31], jj.reshape(-1)[:31]            # 31 pairs: uneven split over 2 ranks

    def fn(a, b):
        return torch.from_numpy(geom_oracle.frame_distance(sc["poses"].numpy(), sc["disps"].numpy(),
                                                           sc["intrinsics"][0].numpy(), a.numpy(), b.numpy(), 0.3))
    got = parallel.sharded_pairs(fn, ii, jj)
    if rank == 0:
        out.put(got.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_pair_distance_gloo_world2():
    """frame-distance pair list split over 2 ranks + one all-gather == the unsharded evaluation."""
    sys.path.insert(0, ROOT)
    from goslam_b200 import synthetic
    from oracle import geom_oracle
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29600 + os.getpid() % 200
    procs = [ctx.Process(target=_pairs_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    got = out.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sc, _ = synthetic.make_scene(num_kf=6, ht=8, wd=10, seed=11, rgbd=True, with_fmaps=False)
    ii, jj = torch.meshgrid(torch.arange(6), torch.arange(6), indexing="ij")
    ii, jj = ii.reshape(-1)[:31], jj.reshape(-1)[:31]
    want = geom_oracle.frame_distance(sc["poses"].numpy(), sc["disps"].numpy(), sc["intrinsics"][0].numpy(),
                                      ii.numpy(), jj.numpy(), 0.3)
    np.testing.assert_array_equal(got, want.astype(np.float32))
