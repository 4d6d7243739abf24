"""InstantNeuS with the reference's constructor / forward signature (src/InstantNeuS.py:219-370),
backed by the fused sm_100a ray marcher (goslam_neus_forward) instead of tiny-cuda-nn + ~60
eager torch kernels.

    net = InstantNeuS(cfg, bound, device)               # cfg = cfg['mapping']['model']
    out = net(rays_o, rays_d, z_vals, dists, render_params=None)   # dict with the 9 reference keys

Parameters keep the reference's names so `state_dict()` round-trips through the same keys the
reference checkpoint (go.ckpt 'mapping_net') holds:
    sdf_network.encoding.encoding.params   tcnn HashGrid params (fp32 master, 16 lvl x 2 feat)
    sdf_network.encoding._B                (unused by the non-directional encoding, kept)
    sdf_network.sdf_layer.{weight,bias}    nn.Linear(35, 32)
    color_network._B                       [3, 33]
    color_network.network.params           tcnn FullyFusedMLP params (fp32 master) 64x80|64x64|16x64
    variance_network.variance
The forward pass is inference-only (the renderer backward / mapping optimiser step is a
"next" row of the scope table); calling it with grad-requiring inputs raises.
"""
import ctypes
import math

import torch
import torch.nn as nn

from . import _lib
from .droid_backends import _workspace

N_LEVELS, N_FEAT = 16, 2
MLP_IN, MLP_IN_PAD, MLP_HID, MLP_OUT_PAD = 67, 80, 64, 16
MLP_PARAMS = MLP_HID * MLP_IN_PAD + MLP_HID * MLP_HID + MLP_OUT_PAD * MLP_HID


def hashgrid_layout():
    """(offsets[17] in params, resolutions[16], scales[16], total_params) from the C-ABI helper."""
    lib = _lib.load()
    off = (ctypes.c_int64 * (N_LEVELS + 1))()
    res = (ctypes.c_int * N_LEVELS)()
    sc = (ctypes.c_float * N_LEVELS)()
    total = lib.goslam_hashgrid_layout(off, res, sc)
    return list(off), list(res), list(sc), int(total)


class _TcnnParams(nn.Module):
    """stands in for a tcnn module: a flat fp32 `params` tensor (tcnn keeps fp32 masters and
    uses fp16 copies on the device, which is what we hand to the kernel)."""

    def __init__(self, n, init):
        super().__init__()
        self.params = nn.Parameter(init(n))
        self.n_output_dims = None


class Encoding(nn.Module):
    def __init__(self, n_input_dims=3, device='cuda:0', direction=False):
        super().__init__()
        if direction:
            raise NotImplementedError("directional (SH) encoding is unused by the reference model")
        self.n_input_dims = n_input_dims
        self.include_xyz = True
        _, _, _, total = hashgrid_layout()
        # tcnn initialises grid params U(-1e-4, 1e-4)
        self.encoding = _TcnnParams(total, lambda n: (torch.rand(n) * 2 - 1) * 1e-4)
        self.encoding.n_output_dims = N_LEVELS * N_FEAT
        self._B = nn.Parameter(torch.randn(n_input_dims, 3) * 25.0)
        self.n_output_dims = 3 + N_LEVELS * N_FEAT


class SDFNetwork(nn.Module):
    def __init__(self, d_in=3, d_out=32, device='cuda:0'):
        super().__init__()
        if d_out != 32:
            raise ValueError("fused marcher is specialised for d_out = 32 (reference default)")
        self.d_in, self.d_out = d_in, d_out
        self.encoding = Encoding(n_input_dims=d_in, device=device)
        self.sdf_layer = nn.Linear(self.encoding.n_output_dims, d_out)
        torch.nn.init.constant_(self.sdf_layer.bias, 0.0)
        torch.nn.init.constant_(self.sdf_layer.weight[:, 3:], 0.0)
        torch.nn.init.normal_(self.sdf_layer.weight[:, :3], mean=0.0, std=math.sqrt(2) / math.sqrt(d_out))

    def get_training_parameters(self, ignore_keys=()):
        return {'network': list(self.sdf_layer.parameters()) + [self.encoding._B],
                'volume': list(self.encoding.encoding.parameters())}


class ColorNetwork(nn.Module):
    def __init__(self, d_in=3, d_feat=31, d_hidden=64, n_layers=2, device='cuda:0'):
        super().__init__()
        if (d_feat, d_hidden, n_layers) != (31, 64, 2):
            raise ValueError("fused marcher is specialised for d_feat=31, d_hidden=64, n_layers=2")
        self._B = nn.Parameter(torch.randn(3, 33) * 25.0)
        # tcnn default init: xavier-uniform per matrix
        def init(n):
            mats = []
            for fo, fi in ((MLP_HID, MLP_IN_PAD), (MLP_HID, MLP_HID), (MLP_OUT_PAD, MLP_HID)):
                lim = math.sqrt(6.0 / (fi + fo))
                mats.append(((torch.rand(fo * fi) * 2 - 1) * lim))
            return torch.cat(mats)
        self.network = _TcnnParams(MLP_PARAMS, init)


class SingleVarianceNetwork(nn.Module):
    def __init__(self, init_val=0.2, scale_factor=10.0):
        super().__init__()
        self.scale_factor = scale_factor
        self.register_parameter('variance', nn.Parameter(torch.tensor(init_val)))

    def forward(self, x):
        return torch.ones(size=[x.shape[0], 1], device=x.device) * torch.exp(self.variance * self.scale_factor)


class InstantNeuS(nn.Module):
    def __init__(self, cfg, bound, device='cuda:0'):
        super().__init__()
        self.cfg = cfg
        self.register_buffer('bound', torch.tensor(bound).float())
        self.register_buffer('realtime_bound', torch.tensor(bound).float())
        self.device = device
        self.sdf_network = SDFNetwork(**cfg['sdf_network'], device=device)
        self.color_network = ColorNetwork(**cfg['color_network'], device=device)
        self.variance_network = SingleVarianceNetwork(**cfg['variance_network'])
        self.sdf_smooth_std = cfg.get('sdf_smooth_std')
        self.sdf_sparse_factor = cfg.get('sdf_sparse_factor')
        self.sdf_truncation = cfg.get('sdf_truncation')
        self.sdf_random_weight = cfg.get('sdf_random_weight')
        self.cos_anneal_ratio = 1.0
        self._half_cache = None

    # ---- reference helper API -------------------------------------------------------------
    def get_training_parameters(self, ignore_keys=()):
        all_params = {
            'sdf_network': list(self.sdf_network.get_training_parameters()['network']),
            'color_network': list(self.color_network.parameters()),
            'variance_network': list(self.variance_network.parameters()),
        }
        params = []
        for k, v in all_params.items():
            if k not in ignore_keys:
                params += v
        return params

    def get_volume_parameters(self):
        return list(self.sdf_network.get_training_parameters()['volume'])

    @torch.no_grad()
    def update_bound(self, bound):
        self.realtime_bound[:] = bound.float().to(self.realtime_bound.device)

    # ---- kernel-side parameter staging ------------------------------------------------------
    def refresh_device_params(self):
        """fp16 copies of the tcnn-style params (call after loading / changing weights)."""
        g = self.sdf_network.encoding.encoding.params
        m = self.color_network.network.params
        self._half_cache = (g.detach().half().contiguous(), m.detach().half().contiguous(),
                            g._version, m._version)

    def _params_struct(self):
        g = self.sdf_network.encoding.encoding.params
        m = self.color_network.network.params
        if (self._half_cache is None or self._half_cache[2] != g._version
                or self._half_cache[3] != m._version or self._half_cache[0].device != g.device):
            self.refresh_device_params()
        grid_h, mlp_h = self._half_cache[0], self._half_cache[1]
        p = _lib.NeusParams()
        keep = [grid_h, mlp_h,
                self.sdf_network.sdf_layer.weight.detach().float().contiguous(),
                self.sdf_network.sdf_layer.bias.detach().float().contiguous(),
                self.color_network._B.detach().float().contiguous()]
        p.grid = grid_h.data_ptr()
        p.mlp_w = mlp_h.data_ptr()
        p.sdf_w = keep[2].data_ptr()
        p.sdf_b = keep[3].data_ptr()
        p.color_B = keep[4].data_ptr()
        key = (self.bound._version, self.realtime_bound._version,
               self.variance_network.variance._version)
        if getattr(self, '_host_cache', None) is None or self._host_cache[0] != key:
            # one D2H sync per parameter change, not per call
            b = self.bound.detach().float().cpu().reshape(-1).tolist()
            rb = self.realtime_bound.detach().float().cpu().reshape(-1).tolist()
            inv_s = float(torch.exp(self.variance_network.variance.detach().float().cpu()
                                    * self.variance_network.scale_factor).clip(1e-6, 1e6))
            self._host_cache = (key, b, rb, inv_s)
        _, b, rb, inv_s = self._host_cache
        for i in range(6):
            p.bound[i] = b[i]
            p.rt_bound[i] = rb[i]
        p.inv_s = inv_s
        p.cos_anneal_ratio = float(self.cos_anneal_ratio)
        return p, keep, inv_s

    # ---- forward ------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, rays_o, rays_d, z_vals, dists, render_params: dict = None, debug=False):
        """debug=True (tests only): additionally keeps the per-sample NeuS alpha [R,S] and SDF normal [R,S,3]
        in `self.last_debug`."""
        if not z_vals.is_cuda:
            raise RuntimeError("InstantNeuS.forward: CUDA tensors required (no CPU fallback)")
        dev = z_vals.device
        R, S = z_vals.shape
        rays_o = rays_o.detach().float().contiguous()
        rays_d = rays_d.detach().float().contiguous()
        z_vals = z_vals.detach().float().contiguous()
        dists = dists.detach().float().contiguous()
        p, keep, inv_s = self._params_struct()
        f32 = dict(dtype=torch.float32, device=dev)
        out = {
            'color': torch.empty((R, 3), **f32), 'depth': torch.empty((R, 1), **f32),
            'depth_variance': torch.empty((R, 1), **f32), 'normal': torch.empty((R, 3), **f32),
            'weight_sum': torch.empty((R, 1), **f32), 'sdf': torch.empty((R, S), **f32),
            'z_vals': torch.empty((R, S), **f32), 'gradient_error': torch.zeros((1,), **f32),
        }
        o = _lib.NeusOut()
        o.color = out['color'].data_ptr(); o.depth = out['depth'].data_ptr()
        o.depth_variance = out['depth_variance'].data_ptr(); o.normal = out['normal'].data_ptr()
        o.weight_sum = out['weight_sum'].data_ptr(); o.sdf = out['sdf'].data_ptr()
        o.z_mid = out['z_vals'].data_ptr(); o.gradient_error = out['gradient_error'].data_ptr()
        self.last_debug = None
        if debug:
            self.last_debug = {'alpha': torch.empty((R, S), **f32), 'grad': torch.empty((R, S, 3), **f32)}
            o.alpha = self.last_debug['alpha'].data_ptr()
            o.grad = self.last_debug['grad'].data_ptr()
            self.last_debug['pos'] = torch.empty((R, S, 3), **f32)
            o.pos = self.last_debug['pos'].data_ptr()
        lib = _lib.load()
        with torch.cuda.device(dev):
            ws = _workspace(lib.goslam_neus_workspace_bytes(R, S), dev)
            rc = lib.goslam_neus_forward(ctypes.byref(p), _lib.ptr(rays_o), _lib.ptr(rays_d),
                                         _lib.ptr(z_vals), _lib.ptr(dists), R, S, ctypes.byref(o),
                                         _lib.ptr(ws), ctypes.c_size_t(ws.numel()), _lib.stream_ptr())
        _lib.check(rc, "neus_forward")
        out['sdf_variance'] = torch.full((R, 1), 1.0 / inv_s, **f32)
        return out
