"""goslam_b200.droid_net.UpdateModule at the config-2 update size (36 edges, 40x80) against the reference forward in
torch / cuDNN under autocast; per-kernel times: ncu launch list of tools/profile_update_op.py."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from goslam_b200.droid_net import UpdateModule, _nhwc_view
from goslam_b200.modules.gru import conv2d_nhwc, to_nchw, to_nhwc, to_nhwc_padded

dev = torch.device("cuda:0")
B, h, w = 36, 40, 80
torch.manual_seed(11)
m = UpdateModule().to(dev).eval()
g = torch.Generator().manual_seed(1)
net = torch.tanh(torch.randn(1, B, 128, h, w, generator=g)).half().to(dev)
inp = torch.relu(torch.randn(1, B, 128, h, w, generator=g)).half().to(dev)
corr = (0.7 * torch.randn(1, B, 196, h, w, generator=g)).half().to(dev)
flow = (4 * torch.randn(1, B, 4, h, w, generator=g)).to(dev)
ii = (torch.arange(B) // 5).to(dev)
jj = ((torch.arange(B) + 1) % 8).to(dev)


def t(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / n, 1e6 * (time.perf_counter() - t0) / n


for name, fn in [("whole forward (one library call)", lambda: m(net, inp, corr, flow, ii, jj))]:
    gpu_us, wall_us = t(fn)
    print("%-38s GPU %8.1f us   wall %8.1f us" % (name, gpu_us, wall_us))


def torch_reference_forward():
    """the reference UpdateModule.forward, op for op, in torch under autocast (cuDNN): src/droid_net.py:107-140"""
    with torch.no_grad(), torch.autocast("cuda", enabled=True):
        n_, i_, c_, f_ = [x.view(B, -1, h, w) for x in (net, inp, corr, flow)]
        c_ = m.corr_encoder(c_)
        f_ = m.flow_encoder(f_)
        gru = m.gru
        x = torch.cat([i_, c_, f_], dim=1)
        net_inp = torch.cat([n_, x], dim=1)
        glo = (torch.sigmoid(gru.w(n_)) * n_).view(B, 128, h * w).mean(dim=-1, keepdim=True).view(B, 128, 1, 1)
        z = torch.sigmoid(gru.convz(net_inp) + gru.convz_glo(glo))
        r = torch.sigmoid(gru.convr(net_inp) + gru.convr_glo(glo))
        q = torch.tanh(gru.convq(torch.cat([r * n_, x], dim=1)) + gru.convq_glo(glo))
        n2 = (1 - z) * n_ + z * q
        delta = m.delta(n2).view(1, B, -1, h, w).permute(0, 1, 3, 4, 2)[..., :2].contiguous()
        weight = m.weight(n2).view(1, B, -1, h, w).permute(0, 1, 3, 4, 2)[..., :2].contiguous()
        eta, upmask = m.agg(n2.view(1, B, 128, h, w), ii)
        return n2, delta, weight, eta, upmask


gpu_us, wall_us = t(torch_reference_forward)
print("%-38s GPU %8.1f us   wall %8.1f us" % ("torch/cuDNN autocast reference forward", gpu_us, wall_us))
