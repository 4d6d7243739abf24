"""world_size-2 gloo test (CPU) of the data-parallel mapping step's host logic (go-slam_b200/parallel.py: ray_slice,
mapping_loss_local, allreduce_gradients): with a stand-in differentiable renderer (plain torch, same output keys) the SUM of
the ranks' gradients must equal the gradient of the reference's single-process loss (src/mapping.py:97-128) on the whole
batch — uneven slices, rays without sensor depth on one rank only."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class ToyNet(torch.nn.Module):
    """stands in for InstantNeuS: a few parameters, the output keys the loss reads, the reference's compute_sdf_error"""
    sdf_truncation, sdf_sparse_factor = 0.16, 5

    def __init__(self):
        super().__init__()
        g = torch.Generator().manual_seed(3)
        self.a = torch.nn.Parameter(torch.randn(3, 3, generator=g) * 0.3)
        self.b = torch.nn.Parameter(torch.randn(4, generator=g) * 0.3)
        self.big = torch.nn.Parameter(torch.randn((1 << 20) + 5, generator=g) * 0.01)     # travels on its own in the all-reduce

    def forward(self, ro, rd, zv, ds):
        S = zv.shape[1]
        feat = torch.tanh(rd @ self.a)
        color = torch.sigmoid(feat + self.b[:3])
        depth = (zv * torch.softmax(zv * self.b[3], dim=1)).sum(1, keepdim=True) + self.big[:5].sum()
        sdf = (zv - depth) * (1.0 + self.a[0, 0]) + self.big[5:5 + S][None] * 3.0
        gerr = ((feat.norm(dim=1) - 1.0) ** 2).mean().reshape(1)
        return {"color": color, "depth": depth, "sdf": sdf, "z_vals": zv, "depth_variance": (zv.var(dim=1, keepdim=True) + 0.1).detach(),
                "gradient_error": gerr}

    def compute_sdf_error(self, sdf, z_vals, gt_depth):
        from goslam_b200.neus import InstantNeuS
        return InstantNeuS.compute_sdf_error(self, sdf, z_vals, gt_depth)


def _batch():
    g = torch.Generator().manual_seed(11)
    R, S = 37, 12
    ro, rd = torch.randn(R, 3, generator=g), torch.randn(R, 3, generator=g)
    zv = torch.sort(torch.rand(R, S, generator=g) * 3.0 + 0.2, dim=1)[0]
    ds = torch.rand(R, S, generator=g) * 0.1
    color = torch.rand(R, 3, generator=g)
    depth = 0.5 + 2.5 * torch.rand(R, generator=g)
    depth[[1, 4, 5, 9]] = 0.0                      # all in rank 0's slice
    return ro, rd, zv, ds, color, depth


def _reference_loss(net, out, rays_color, rays_depth):
    """src/mapping.py:97-128, weights of configs/go_slam.yaml"""
    depth = rays_depth.reshape(-1, 1)
    valid = (depth > 0).reshape(-1)
    unc = 1.0 / torch.sqrt(out["depth_variance"][valid].detach() + 1e-10)
    total = torch.abs(out["color"][valid] - rays_color[valid]).mean() * 2.0
    total = total + (torch.abs(out["depth"][valid] - depth[valid]) * unc).mean()
    sl, spl = net.compute_sdf_error(sdf=out["sdf"][valid], z_vals=out["z_vals"][valid], gt_depth=depth[valid])
    return total + (sl + spl) * 2.0 + 0.1 * out["gradient_error"].mean()


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from goslam_b200 import parallel
    ro, rd, zv, ds, color, depth = _batch()
    net = ToyNet()
    lo, hi = parallel.ray_slice(ro.shape[0])
    out = net(ro[lo:hi], rd[lo:hi], zv[lo:hi], ds[lo:hi])
    # the toy's gradient_error is a mean over RAYS (the real one over samples): its share is the slice's share of the rays
    loss = parallel.mapping_loss_local(net, out, color[lo:hi], depth[lo:hi], ro.shape[0], 2.0, 2.0, 0.1)
    loss.backward()
    parallel.allreduce_gradients(list(net.parameters()))
    total = loss.detach().clone().reshape(1)
    dist.all_reduce(total)
    if rank == 0:
        q.put((float(total), [p.grad.clone() for p in net.parameters()], (lo, hi)))
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_mapping_loss_and_gradients_match_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29650 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    total, grads, sl = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sys.path.insert(0, ROOT)
    ro, rd, zv, ds, color, depth = _batch()
    net = ToyNet()
    _reference_loss(net, net(ro, rd, zv, ds), color, depth).backward()
    want_total = float(_reference_loss(net, net(ro, rd, zv, ds), color, depth))
    assert sl == (0, 19)                                          # 37 rays over 2 ranks: 19 + 18
    assert abs(total - want_total) <= 1e-5 * abs(want_total)
    for g, p in zip(grads, net.parameters()):
        assert torch.allclose(g, p.grad, rtol=2e-4, atol=1e-7), float((g - p.grad).abs().max())
