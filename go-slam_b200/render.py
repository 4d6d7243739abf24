"""Renderer with the reference's call signature (src/render.py:6-175): near/far from the scene
AABB, stratified + surface-guided z sampling, then ONE call of the fused marcher for the whole
ray batch (the reference splits into 10,000-ray chunks of ~60 eager kernels each, :53-59).

The z-sampling is a handful of [R, S] elementwise torch ops + one sort over 72 columns; it is
kept in torch so that, given the same torch RNG state, it reproduces the reference's samples
bit for bit (shared `perturb_rand[N_samples]` for all rays included, src/render.py:159).
"""
import torch


def sample_z(rays_o, rays_d, bound, gt_depth, n_samples, n_surface, perturb=1.0, lindisp=False):
    """Returns z_vals [R, S], dists [R, S] with S = n_samples (+ n_surface when gt_depth is given)."""
    device = rays_o.device
    n_rays = rays_o.shape[0]
    if gt_depth is None:
        n_surface = 0
        near = 0.01
    else:
        gt_depth = gt_depth.reshape(-1, 1)
        near = gt_depth.repeat(1, n_samples) * 0.01
    with torch.no_grad():
        t = (bound[None].to(device) - rays_o.detach()[:, :, None]) / rays_d.detach()[:, :, None]
        far_bb = torch.min(torch.max(t, dim=2)[0], dim=1)[0][:, None] + 0.01
    far = torch.clamp(far_bb, 0, (gt_depth * 1.2).max()) if gt_depth is not None else far_bb

    z_surface = None
    if n_surface > 0:
        valid = gt_depth > 0
        vdepth = (gt_depth * valid).repeat(1, n_surface)
        ts = torch.linspace(0, 1, steps=n_surface).float().to(device)[None, :].repeat(n_rays, 1)
        snr, sfar = (1 - 0.1) * vdepth, (1 + 0.1) * vdepth
        z_valid = snr + (sfar - snr) * ts
        z_invalid = 0.001 + (gt_depth.max() - 0.001) * ts
        z_surface = z_valid * valid + z_invalid * (1 - valid.float())

    tv = torch.linspace(0, 1, steps=n_samples, device=device)[None, :].repeat(n_rays, 1)
    if not lindisp:
        z_vals = near + (far - near) * tv
        sample_dist = ((far - near) / n_samples).mean(dim=1, keepdim=True)
    else:
        z_vals = 1.0 / (1.0 / far + (1.0 / near - 1.0 / far) * tv)
        sample_dist = 1.0 / ((1.0 / near - 1.0 / far) / n_samples).mean(dim=1, keepdim=True)
    if perturb > 0:
        mid = 0.5 * (z_vals[:, :-1] + z_vals[:, 1:])
        upper = torch.cat([mid, z_vals[:, -1:]], dim=1)
        lower = torch.cat([z_vals[:, :1], mid], dim=1)
        z_vals = lower + (upper - lower) * torch.rand(n_samples, device=device)
    if n_surface > 0:
        z_vals, _ = torch.sort(torch.cat([z_vals, z_surface.float()], dim=1), dim=1)
    dists = torch.cat([z_vals[..., 1:] - z_vals[..., :-1], sample_dist], dim=-1)
    return z_vals, dists


class Renderer(object):
    def __init__(self, cfg, args, slam, points_batch_size=1e4, ray_batch_size=5e3):
        self.ray_batch_size = int(ray_batch_size)
        self.points_batch_size = int(points_batch_size)     # kept for API parity; not needed
        r = cfg['rendering']
        self.lindisp, self.perturb = r['lindisp'], r['perturb']
        self.N_samples, self.N_surface = r['N_samples'], r['N_surface']
        self.H, self.W, self.fx, self.fy, self.cx, self.cy = slam.H, slam.W, slam.fx, slam.fy, slam.cx, slam.cy

    def eval_points(self, rays_o, rays_d, z_vals, dists, net, render_params):
        return net(rays_o, rays_d, z_vals, dists, render_params=render_params)

    def render_batch_ray(self, rays_o, rays_d, net, render_params, device='cuda:0', gt_depth=None):
        z_vals, dists = sample_z(rays_o, rays_d, net.bound, gt_depth, self.N_samples, self.N_surface,
                                 self.perturb, self.lindisp)
        return self.eval_points(rays_o, rays_d, z_vals, dists, net, render_params)
