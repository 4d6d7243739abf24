"""CPU oracle (TEST INFRASTRUCTURE — never imported by the product path) for the renderer BACKWARD
(SURVEY 8f-3): differentiable torch restatements of the two tiny-cuda-nn modules InstantNeuS uses
(src/InstantNeuS.py:62 `tcnn.Encoding` HashGrid, :201 `tcnn.Network` FullyFusedMLP), so that the REFERENCE's own
InstantNeuS.forward — including its `torch.autograd.grad(..., create_graph=True)` normal (:139-146) — and the losses of
Mapper.optimize_map (src/mapping.py:97-128) can be differentiated by plain autograd.

PARITY UNPINNED for tiny-cuda-nn's own backward arithmetic (no version pin, source absent: see oracle/neus_oracle.py):
what is restated is the mathematical gradient of the published forward algorithm, with fp16 ROUNDING treated as the
identity in the backward pass (straight-through), which is also what tcnn's mixed-precision backward amounts to.
Everything outside tcnn is the reference's own code (tests/golden/make_golden.py neus_grad imports it).
"""
import numpy as np
import torch

from . import neus_oracle as no


def _ste_half(x):
    """value rounded to fp16, gradient of the identity"""
    return x + (x.detach().half().to(x.dtype) - x.detach())


class TorchHashGrid(torch.nn.Module):
    """tcnn.Encoding(HashGrid, 16 levels x 2 features, T = 2^19, base 16, per-level scale 1.4473): fp32 master params,
    fp16 values at use; trilinear interpolation; differentiable w.r.t. the params AND the input (twice)."""

    def __init__(self, n_input_dims=3, encoding_config=None):
        super().__init__()
        self.metas, total = no.hashgrid_meta()
        self.n_output_dims = no.N_LEVELS * no.N_FEAT
        self.params = torch.nn.Parameter((torch.rand(total * 2) * 2 - 1) * 1e-4)

    def forward(self, x):
        table = _ste_half(self.params).view(-1, no.N_FEAT)
        outs = []
        for m in self.metas:
            pos = (x.double() * float(m["scale"]) + 0.5).to(torch.float32)          # fmaf(scale, x, 0.5)
            fl = torch.floor(pos).detach()
            fr = pos - fl
            pg = fl.to(torch.int64).numpy().astype(np.uint32)
            acc = 0
            with np.errstate(over="ignore"):
                for idx in range(8):
                    w = 1.0
                    c = []
                    for d in range(3):
                        if (idx >> d) & 1:
                            w = w * fr[:, d]
                            c.append(pg[:, d] + np.uint32(1))
                        else:
                            w = w * (1.0 - fr[:, d])
                            c.append(pg[:, d])
                    gi = torch.from_numpy(m["offset"] + no._grid_index(m, *c))
                    acc = acc + w[:, None] * table[gi]
            outs.append(acc)
        return torch.cat(outs, dim=1)


class TorchMLP(torch.nn.Module):
    """tcnn.Network(FullyFusedMLP 67 -> 64 -> 64 -> 3): input padded to 80 with ones, output padded to 16, no biases,
    ReLU, fp16 weights and activations (straight-through), fp32 accumulation."""

    def __init__(self, n_input_dims=67, n_output_dims=3, network_config=None):
        super().__init__()
        self.params = torch.nn.Parameter(torch.zeros(64 * 80 + 64 * 64 + 16 * 64))

    def forward(self, x):
        p = _ste_half(self.params)
        W1 = p[:64 * 80].view(64, 80)
        W2 = p[64 * 80:64 * 80 + 64 * 64].view(64, 64)
        W3 = p[64 * 80 + 64 * 64:].view(16, 64)
        xin = torch.cat([x.float(), torch.ones(x.shape[0], 13)], dim=1)
        h = _ste_half(torch.relu(_ste_half(xin) @ W1.t()))
        h = _ste_half(torch.relu(h @ W2.t()))
        return _ste_half(h @ W3.t())[:, :3]


def mapping_loss(net, out, rays_color, rays_depth, w_color=1.0, w_sdf=None, w_eikonal=None):
    """the loss of Mapper.optimize_map (src/mapping.py:97-128) with uncertainty_based off; `net` supplies
    compute_sdf_error (the reference's own method when `net` is the reference class)."""
    rays_depth = rays_depth.reshape(-1, 1)
    valid = (rays_depth > 0).reshape(-1)
    total = torch.abs(out["color"][valid] - rays_color[valid]).mean() * w_color
    total = total + torch.abs(out["depth"][valid] - rays_depth[valid]).mean()
    if w_sdf:
        sdf_loss, sparse_loss = net.compute_sdf_error(sdf=out["sdf"][valid], z_vals=out["z_vals"][valid], gt_depth=rays_depth[valid])
        total = total + (sdf_loss + sparse_loss) * w_sdf
    if w_eikonal:
        total = total + w_eikonal * out["gradient_error"].mean()
    return total
