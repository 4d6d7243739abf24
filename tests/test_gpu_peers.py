"""goslam_ba_phase1_peers / goslam_ba_phase2_peers (the split form over peer memory) on ONE GPU: a one-rank PeerLink has
nothing to map, but the whole device path runs — peer-allocated system / disps / flag buffers, the summing load of the
solve kernels, the write-through back-substitution, the epoch flags.  Must reproduce droid_backends.ba.  The 2-rank run
(results equal to the NCCL exchange bit for bit, replicas identical) is tools/run_sharded_ba.py and bench.py --gpus 2."""
import pytest
import torch
import torch.distributed as dist

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("num_kf,ht,wd,iters,lm,ep", [(8, 40, 80, 3, 1e-4, 0.1), (40, 30, 40, 2, 1e-5, 1e-2), (110, 12, 16, 1, 1e-5, 1e-2)])
def test_peer_split_form_single_rank_matches_ba(num_kf, ht, wd, iters, lm, ep):
    from goslam_b200 import droid_backends, parallel, synthetic
    dev = torch.device("cuda:0")
    own = not dist.is_initialized()
    if own:
        dist.init_process_group("gloo", store=dist.HashStore(), rank=0, world_size=1)
    try:
        sc, g = synthetic.make_scene(num_kf=num_kf, ht=ht, wd=wd, seed=31, rgbd=True, with_fmaps=False)
        tg, wg, eta = synthetic.make_update(sc, synthetic.true_reprojection(sc)[0], g, noise=0.6)
        kx = torch.unique(torch.cat([torch.arange(1, num_kf), sc["ii"]]))
        eta_f = torch.zeros(num_kf, ht, wd)
        eta_f[kx] = eta
        D = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in sc.items()}
        tg, wg, eta, eta_f = tg.to(dev), wg.to(dev), eta.to(dev), eta_f.to(dev)
        pa, da = D["poses"].clone(), D["disps"].clone()
        droid_backends.ba(pa, da, D["intrinsics"][0].contiguous(), D["disps_sens"], tg, wg, eta, D["ii"], D["jj"], 1, num_kf, iters, lm, ep, False)
        pb = D["poses"].clone()
        link = parallel.PeerLink(D["disps"], 36 * (num_kf - 1) ** 2 + 6 * (num_kf - 1))
        be = parallel.PeerBackend(pb, link, D["intrinsics"][0].contiguous(), D["disps_sens"], 1, num_kf)
        for _ in range(iters):
            dx = be.iteration(tg, wg, eta_f, D["ii"], D["jj"], lm, ep, False, 0, num_kf)
        link.wait_idle()
        torch.cuda.synchronize()
        rel = lambda a, b: ((a - b).abs().max() / b.abs().max()).item()   # noqa: E731
        assert link.epoch == iters and not bool(link.timeout.item())
        assert rel(pb, pa) < 1e-5 and rel(link.disps.tensor, da) < 1e-5
        assert torch.isfinite(dx).all()
        link.close()
    finally:
        if own:
            dist.destroy_process_group()
