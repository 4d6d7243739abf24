"""CPU: the differentiable tcnn restatements of oracle/neus_grad_oracle.py are pinned to the forward oracle
(oracle/neus_oracle.py, itself pinned to the reference golden), and the compute_sdf_error mirror to the values the
REFERENCE's method produced inside tests/golden/neus_grad.npz."""
import os

import numpy as np
import torch

from oracle import neus_grad_oracle as ngo
from oracle import neus_oracle as no


def test_torch_hashgrid_matches_numpy_oracle_and_its_input_gradient():
    torch.manual_seed(0)
    enc = ngo.TorchHashGrid()
    with torch.no_grad():
        enc.params.copy_(torch.randn(enc.params.numel()) * 0.05)
    x = torch.rand(200, 3, requires_grad=True)
    y = enc(x)
    table = enc.params.detach().half().numpy().reshape(-1, 2)
    want = no.hashgrid_encode(x.detach().numpy(), table).astype(np.float32)
    assert np.abs(y.detach().numpy() - want).max() <= 2e-3 * np.abs(want).max()       # half accumulation in the oracle
    gy = torch.randn(32)
    (gx,) = torch.autograd.grad((y * gy).sum(), x)
    want_g = no.hashgrid_input_grad(x.detach().numpy(), table, gy.numpy())
    assert np.abs(gx.numpy() - want_g).max() <= 2e-3 * np.abs(want_g).max()


def test_torch_mlp_matches_numpy_oracle():
    torch.manual_seed(1)
    mlp = ngo.TorchMLP()
    with torch.no_grad():
        mlp.params.copy_(torch.randn(mlp.params.numel()) * 0.2)
    x = torch.randn(100, 67)
    got = mlp(x).detach().numpy()
    want = no.mlp_forward(x.numpy(), mlp.params.detach().half().numpy()).astype(np.float32)
    assert np.abs(got - want).max() <= 4e-3 * max(1.0, np.abs(want).max())


def test_compute_sdf_error_mirror_matches_reference_values():
    from goslam_b200 import neus, synthetic
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "neus_grad.npz"))
    cfg = synthetic.NEUS_CFG

    class Shell:                      # the method only reads two config numbers
        sdf_truncation, sdf_sparse_factor = cfg["sdf_truncation"], cfg["sdf_sparse_factor"]
    depth = torch.from_numpy(g["rays_depth"]).reshape(-1, 1)
    valid = (depth > 0).reshape(-1)
    sdf_loss, sparse_loss = neus.InstantNeuS.compute_sdf_error(Shell, torch.from_numpy(g["out_sdf"])[valid],
                                                               torch.from_numpy(g["out_z_vals"])[valid], depth[valid])
    assert abs(float(sdf_loss) - float(g["parts"][2])) <= 1e-5 * abs(float(g["parts"][2]))
    assert abs(float(sparse_loss) - float(g["parts"][3])) <= 1e-5 * max(abs(float(g["parts"][3])), 1e-3)
