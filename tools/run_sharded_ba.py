"""torchrun --nproc-per-node 2 tools/run_sharded_ba.py : one factor graph sharded over 2+ GPUs
(edges by source frame, one NCCL all-reduce of the reduced camera system per iteration) must
reproduce single-GPU droid_backends.ba."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from goslam_b200 import droid_backends, parallel, synthetic
    num_kf, ht, wd = 16, 40, 80
    sc, g = synthetic.make_scene(num_kf=num_kf, ht=ht, wd=wd, seed=9, rgbd=True, with_fmaps=False)
    coords = synthetic.true_reprojection(sc)
    tg, wg, eta = synthetic.make_update(sc, coords[0], g, noise=0.6)
    t0, t1, iters = 1, num_kf, 3
    kx = torch.unique(torch.cat([torch.arange(t0, t1), sc["ii"]]))
    eta_f = torch.zeros(num_kf, ht, wd)
    eta_f[kx] = eta
    D = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in sc.items()}
    tg, wg, eta, eta_f = tg.to(dev), wg.to(dev), eta.to(dev), eta_f.to(dev)
    # single-GPU reference on every rank
    p1, d1 = D["poses"].clone(), D["disps"].clone()
    droid_backends.ba(p1, d1, D["intrinsics"][0].contiguous(), D["disps_sens"], tg, wg, eta, D["ii"], D["jj"],
                      t0, t1, iters, 1e-4, 0.1, False)
    # sharded
    p2, d2 = D["poses"].clone(), D["disps"].clone()
    be = parallel.CudaBackend(p2, d2, D["intrinsics"][0].contiguous(), D["disps_sens"], t0, t1)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    parallel.sharded_ba(be, d2, tg, wg, eta_f, D["ii"], D["jj"], iters, 1e-4, 0.1)
    e1.record(); torch.cuda.synchronize()
    rel = lambda a, b: ((a - b).abs().max() / b.abs().max()).item()   # noqa: E731
    ok = rel(p2, p1) < 1e-5 and rel(d2, d1) < 1e-5
    print("rank %d/%d: poses rel %.2e disps rel %.2e  sharded BA %.1f us  %s" %
          (rank, world, rel(p2, p1), rel(d2, d1), 1e3 * e0.elapsed_time(e1), "OK" if ok else "MISMATCH"), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
