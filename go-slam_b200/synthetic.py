"""Deterministic synthetic scenes for tests and bench: smooth inverse-depth
maps, a gently moving camera, random half-precision feature maps, DROID-style factor graphs,
update-operator-like targets / weights / damping, and ray batches for the renderer.
All generation is on the CPU with a torch.Generator (seed 43 = the reference's run.py seed),
so the GPU box and this container see bit-identical inputs.
"""
import math

import torch


def _gen(seed):
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return g


def _smooth(g, n, h, w, cells=6):
    low = torch.rand(n, 1, cells, cells + 2, generator=g)
    return torch.nn.functional.interpolate(low, size=(h, w), mode="bicubic", align_corners=True)[:, 0].clamp(0, 1)


def _qmul(a, b):
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return torch.stack([aw * bx + ax * bw + ay * bz - az * by, aw * by + ay * bw + az * bx - ax * bz,
                        aw * bz + az * bw + ax * by - ay * bx, aw * bw - ax * bx - ay * by - az * bz])


def _qrot(q, v):
    qv = q[:3]
    uv = 2 * torch.linalg.cross(qv, v)
    return v + q[3] * uv + torch.linalg.cross(qv, uv)


def make_poses(num, g, trans_sigma=0.05, rot_sigma=0.02):
    """identity for keyframe 0, then small random twists composed sequentially (world->cam)."""
    poses = torch.zeros(num, 7)
    poses[:, 6] = 1.0
    for k in range(1, num):
        xi = torch.cat([torch.randn(3, generator=g) * trans_sigma, torch.randn(3, generator=g) * rot_sigma])
        th = xi[3:].norm()
        dq = torch.cat([torch.sin(0.5 * th) / th.clamp_min(1e-9) * xi[3:], torch.cos(0.5 * th)[None]])
        q = _qmul(dq, poses[k - 1, 3:])
        t = _qrot(dq, poses[k - 1, :3]) + xi[:3]
        poses[k, :3], poses[k, 3:] = t, q / q.norm()
    return poses


def neighborhood_edges(t0, t1, r=3):
    """FactorGraph.add_neighborhood_factors (src/factor_graph.py:368-382): |i-j| <= r, i != j."""
    ii, jj = torch.meshgrid(torch.arange(t0, t1), torch.arange(t0, t1), indexing="ij")
    ii, jj = ii.reshape(-1), jj.reshape(-1)
    keep = ((ii - jj).abs() > 0) & ((ii - jj).abs() <= r)
    return ii[keep].long(), jj[keep].long()


def make_scene(num_kf=8, ht=40, wd=80, seed=43, rgbd=True, buffer=None, fx=None, radius=3,
               with_fmaps=True, stereo_edges=0):
    """A window of `num_kf` keyframes at 1/8 resolution ht x wd inside a `buffer`-long video."""
    g = _gen(seed)
    buffer = buffer or num_kf
    fx = fx if fx is not None else 0.9 * wd
    intr = torch.tensor([fx, fx, wd / 2.0 - 0.3, ht / 2.0 + 0.2])
    disps = torch.ones(buffer, ht, wd)
    disps[:num_kf] = 0.5 + 0.4 * _smooth(g, num_kf, ht, wd)
    poses = torch.zeros(buffer, 7)
    poses[:, 6] = 1.0
    poses[:num_kf] = make_poses(num_kf, g)
    intrinsics = intr[None].repeat(buffer, 1).contiguous()
    disps_sens = torch.zeros(buffer, ht, wd)
    if rgbd:
        disps_sens[:num_kf] = (disps[:num_kf] * (1 + 0.02 * torch.randn(num_kf, ht, wd, generator=g))).clamp_min(0.05)
        holes = torch.rand(num_kf, ht, wd, generator=g) < 0.1       # missing sensor readings
        disps_sens[:num_kf][holes] = 0.0
    ii, jj = neighborhood_edges(0, num_kf, radius)
    if stereo_edges:
        s = torch.arange(min(stereo_edges, num_kf))
        ii, jj = torch.cat([ii, s]), torch.cat([jj, s])
    scene = dict(poses=poses, disps=disps, intrinsics=intrinsics, disps_sens=disps_sens,
                 ii=ii, jj=jj, ht=ht, wd=wd, num_kf=num_kf, t0=1, t1=num_kf)
    if with_fmaps:
        c = 2 if stereo_edges else 1
        scene["fmaps"] = torch.randn(buffer, c, 128, ht, wd, generator=g).half()
    return scene, g


def true_reprojection(scene):
    """Where every source pixel of every edge lands in its target frame, in plain torch — the seed of the
    synthetic flow targets (setup helper for bench / tools; same model as DepthVideo.reproject:
    G_ij = G_j G_i^-1, fixed stereo baseline for i == j, pinhole projection).  Returns [1, N, ht, wd, 2]."""
    poses, disps, K = scene["poses"].float(), scene["disps"].float(), scene["intrinsics"].float()
    ii, jj = scene["ii"].long(), scene["jj"].long()
    N = ii.numel()
    _, ht, wd = disps.shape
    qi, qj = poses[ii, 3:].T, poses[jj, 3:].T                      # [4, N]
    q = _qmul(qj, qi * torch.tensor([-1.0, -1.0, -1.0, 1.0])[:, None])
    t = poses[jj, :3].T - _vrot(q, poses[ii, :3].T)                # [3, N]
    same = (ii == jj)
    q = torch.where(same[None], torch.tensor([0.0, 0.0, 0.0, 1.0])[:, None], q)
    t = torch.where(same[None], torch.tensor([-0.1, 0.0, 0.0])[:, None], t)
    v, u = torch.meshgrid(torch.arange(ht).float(), torch.arange(wd).float(), indexing="ij")
    Ki, Kj = K[ii], K[jj]                                          # [N, 4]
    X = torch.stack([(u[None] - Ki[:, 2, None, None]) / Ki[:, 0, None, None],
                     (v[None] - Ki[:, 3, None, None]) / Ki[:, 1, None, None],
                     torch.ones(N, ht, wd)], 0)                    # [3, N, ht, wd]
    X1 = _vrot(q[:, :, None, None], X) + t[:, :, None, None] * disps[ii][None]
    Z = torch.where(X1[2] < 0.1, torch.ones(()), X1[2])
    x = Kj[:, 0, None, None] * (X1[0] / Z) + Kj[:, 2, None, None]
    y = Kj[:, 1, None, None] * (X1[1] / Z) + Kj[:, 3, None, None]
    return torch.stack([x, y], -1)[None]


def _vrot(q, v):
    """rotate v [3, ...] by quaternions q [4, ...] (x, y, z, w), broadcasting."""
    qv = q[:3]
    uv = 2 * torch.stack([qv[1] * v[2] - qv[2] * v[1], qv[2] * v[0] - qv[0] * v[2], qv[0] * v[1] - qv[1] * v[0]])
    return v + q[3] * uv + torch.stack([qv[1] * uv[2] - qv[2] * uv[1], qv[2] * uv[0] - qv[0] * uv[2],
                                        qv[0] * uv[1] - qv[1] * uv[0]])


def make_update(scene, coords, g, noise=0.5, oob_frac=0.0):
    """targets / weights / eta like the update operator would emit (src/factor_graph.py:212-241):
    target = reprojection + N(0, noise px); weight ~ U(0,1); eta = 0.2*damping + 1e-7."""
    N = scene["ii"].numel()
    ht, wd = scene["ht"], scene["wd"]
    target = coords.reshape(N, ht, wd, 2) + noise * torch.randn(N, ht, wd, 2, generator=g)
    if oob_frac > 0:
        m = torch.rand(N, ht, wd, generator=g) < oob_frac
        target[m] += 200.0
    weight = torch.rand(N, ht, wd, 2, generator=g)
    t0, t1 = scene["t0"], scene["t1"]
    kx = torch.unique(torch.cat([torch.arange(t0, t1), scene["ii"]]))
    eta = 0.2 * (0.01 * torch.rand(kx.numel(), ht, wd, generator=g)) + 1e-7
    targets = target.permute(0, 3, 1, 2).contiguous()
    weights = weight.permute(0, 3, 1, 2).contiguous()
    return targets, weights, eta


def make_rays(R, S=72, seed=43, bound=((-2.0, 2.0), (-2.0, 2.0), (-2.0, 2.0)), H=512, W=512, n_uniform=24):
    """A pinhole ray batch inside `bound` with the reference's z-sampling layout
    (src/nerf_func.py:166-179 un-normalised pinhole dirs; src/render.py:99-171: N_samples
    stratified + N_surface within +-10% of the depth, sorted)."""
    g = _gen(seed)
    fx = 0.9 * W
    idx = torch.randperm(H * W, generator=g)[:R] if R <= H * W else torch.randint(0, H * W, (R,), generator=g)
    py, px = (idx // W).float(), (idx % W).float()
    dirs_cam = torch.stack([(px - W / 2) / fx, (py - H / 2) / fx, torch.ones(R)], dim=-1)
    ang = 0.3
    c2w = torch.tensor([[math.cos(ang), 0, math.sin(ang)], [0, 1, 0], [-math.sin(ang), 0, math.cos(ang)]])
    rays_d = dirs_cam @ c2w.T
    rays_o = torch.tensor([0.1, -0.2, -1.0]).expand(R, 3).contiguous()
    depth = 0.5 + 2.5 * torch.rand(R, generator=g)
    n_surface = S - n_uniform
    b = torch.tensor(bound)
    t = (b[None] - rays_o[:, :, None]) / rays_d[:, :, None]
    far = t.max(dim=2)[0].min(dim=1)[0][:, None] + 0.01
    far = far.clamp(0, float(depth.max()) * 1.2)
    near = depth[:, None] * 0.01
    tv = torch.linspace(0, 1, n_uniform)[None]
    z = near + (far - near) * tv
    mid = 0.5 * (z[:, :-1] + z[:, 1:])
    upper, lower = torch.cat([mid, z[:, -1:]], 1), torch.cat([z[:, :1], mid], 1)
    z = lower + (upper - lower) * torch.rand(n_uniform, generator=g)
    ts = torch.linspace(0, 1, n_surface)[None]
    zs = 0.9 * depth[:, None] + 0.2 * depth[:, None] * ts
    z_vals, _ = torch.sort(torch.cat([z, zs], 1), dim=1)
    sample_dist = ((far - near) / n_uniform).mean(dim=1, keepdim=True)
    dists = torch.cat([z_vals[:, 1:] - z_vals[:, :-1], sample_dist], dim=-1)
    return rays_o.contiguous(), rays_d.contiguous(), z_vals.contiguous(), dists.contiguous()


NEUS_CFG = {
    "sdf_network": {"d_in": 3, "d_out": 32},
    "color_network": {"d_in": 3, "d_feat": 31, "d_hidden": 64, "n_layers": 2},
    "variance_network": {"init_val": 0.2, "scale_factor": 10.0},
    "sdf_smooth_std": 0.005, "sdf_sparse_factor": 5, "sdf_truncation": 0.16, "sdf_random_weight": 0.04,
}


def make_neus_weights(seed=43, total_grid_params=None, trained_like=True, layout=None):
    """Random-init weights of the reference architecture.  `trained_like` gives the hash grid a
    per-level amplitude ~ 1/resolution (so d(enc)/dx is O(1) per level, as in a fitted SDF) and
    a non-zero SDF head, so that rays terminate at different depths instead of alpha == 0.
    `layout` = (offsets_in_params[17], resolutions[16]) from neus.hashgrid_layout()."""
    g = _gen(seed)
    n = total_grid_params
    grid = (torch.rand(n, generator=g) * 2 - 1)
    if trained_like and layout is not None:
        offs, ress = layout
        for l in range(len(ress)):
            grid[offs[l]:offs[l + 1]] *= 2.0 / ress[l]
    else:
        grid *= (0.05 if trained_like else 1e-4)
    sdf_w = torch.zeros(32, 35)
    sdf_w[:, :3] = torch.randn(32, 3, generator=g) * math.sqrt(2) / math.sqrt(32)
    if trained_like:
        sdf_w[:, 3:] = torch.randn(32, 32, generator=g) * 0.15
        sdf_w[0, :3] = torch.tensor([0.15, -0.25, 0.6])      # a tilted plane-ish SDF
    sdf_b = torch.zeros(32)
    if trained_like:
        sdf_b = 0.1 * torch.randn(32, generator=g)
        sdf_b[0] = -0.15
    color_B = torch.randn(3, 33, generator=g) * 25.0
    mats = []
    for fo, fi in ((64, 80), (64, 64), (16, 64)):
        lim = math.sqrt(6.0 / (fi + fo))
        mats.append((torch.rand(fo * fi, generator=g) * 2 - 1) * lim)
    mlp = torch.cat(mats)
    return dict(grid=grid, sdf_w=sdf_w, sdf_b=sdf_b, color_B=color_B, mlp=mlp)
