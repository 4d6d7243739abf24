"""A deterministic stand-in for DroidNet.update (src/droid_net.py:107-140) with the same signature and
output shapes, used by tests/golden/make_golden.py (driving the REFERENCE FactorGraph on the CPU) and by the
GPU drop-in test (driving goslam_b200.FactorGraph).  Every output depends on every input (correlation
features, motion features, hidden state), so an error anywhere in the graph's plumbing — edge order,
lookup channels, motion clamp, net / inp gather — shows up in the flow targets and hence in BA.
Plain float32 torch ops only: the same arithmetic on CPU and GPU."""
import torch


def update_op(net, inp, corr, motion, ii, jj):
    """net, inp [1,N,128,h,w] (half); corr [1,N,196,h,w]; motion [1,N,4,h,w]; ii, jj [N].
    Returns net' [1,N,128,h,w] (net's dtype), delta [1,N,h,w,2], weight [1,N,h,w,2], damping [1,M,h,w],
    upmask [1,M,576,h,w] with M = number of distinct source frames (GraphAgg, src/droid_net.py:51-67)."""
    c, m, n, x = corr.float(), motion.float(), net.float(), inp.float()
    lv = [c[:, :, 49 * k:49 * (k + 1)].mean(dim=2) for k in range(4)]               # [1,N,h,w] per level
    hid = n.mean(dim=2) + 0.5 * x.mean(dim=2)
    du = 0.4 * torch.tanh(0.8 * lv[0] - 0.5 * lv[2] + 0.02 * m[:, :, 2] + 0.3 * hid)
    dv = 0.4 * torch.tanh(0.8 * lv[1] - 0.5 * lv[3] - 0.02 * m[:, :, 3] - 0.3 * hid)
    delta = torch.stack([du, dv], dim=-1)
    wu = torch.sigmoid(1.5 * lv[0] + 0.01 * m[:, :, 0] + 0.5)
    wv = torch.sigmoid(1.5 * lv[1] + 0.01 * m[:, :, 1] + 0.5)
    weight = torch.stack([wu, wv], dim=-1)
    net_new = (0.9 * n + 0.1 * torch.tanh(c[:, :, :128] + x)).to(net.dtype)
    # per-source-frame aggregation, in sorted frame order like GraphAgg's unique(ii, return_inverse)
    frames, slot = torch.unique(ii, sorted=True, return_inverse=True)
    M = frames.numel()
    agg = torch.zeros((1, M) + tuple(lv[0].shape[2:]), dtype=torch.float32, device=c.device)
    cnt = torch.zeros(M, dtype=torch.float32, device=c.device)
    agg.index_add_(1, slot, wu * wv)
    cnt.index_add_(0, slot, torch.ones_like(slot, dtype=torch.float32))
    agg = agg / cnt.view(1, M, 1, 1)
    damping = 0.01 * torch.nn.functional.softplus(2.0 * agg - 0.5)
    taps = torch.sin(0.37 * torch.arange(576, dtype=torch.float32, device=c.device)).view(1, 1, 576, 1, 1)
    upmask = 2.0 * taps * (agg.unsqueeze(2) - 0.25)
    return net_new, delta, weight, damping, upmask
