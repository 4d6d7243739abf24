"""ConvGRU at the config-2 update size (36 edges, 40x80): the tcgen05 kernel (NHWC in/out, and through the
reference's NCHW call) against the same module evaluated by torch/cuDNN under autocast (what the reference runs).
usage: python tools/time_gru.py [B h w]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F
from goslam_b200.modules.gru import ConvGRU, to_nhwc

dev = torch.device("cuda:0")
B, h, w = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (36, 40, 80)
torch.manual_seed(77)
m = ConvGRU(128, 320).to(dev)
g = torch.Generator().manual_seed(1)
net = torch.tanh(torch.randn(B, 128, h, w, generator=g)).to(dev).half()
inp, corr = [torch.relu(torch.randn(B, 128, h, w, generator=g)).to(dev).half() for _ in range(2)]
flow = torch.relu(torch.randn(B, 64, h, w, generator=g)).to(dev).half()


def t(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / n


def cudnn_gru():
    """the reference's ConvGRU.forward, op for op (src/modules/gru.py:21-39), under autocast"""
    with torch.no_grad(), torch.autocast("cuda", enabled=True):
        x = torch.cat([inp, corr, flow], dim=1)
        net_inp = torch.cat([net, x], dim=1)
        b, c, hh, ww = net.shape
        glo = torch.sigmoid(m.w(net)) * net
        glo = glo.view(b, c, hh * ww).mean(dim=-1, keepdim=True).view(b, c, 1, 1)
        z = torch.sigmoid(m.convz(net_inp) + m.convz_glo(glo))
        r = torch.sigmoid(m.convr(net_inp) + m.convr_glo(glo))
        q = torch.tanh(m.convq(torch.cat([r * net, x], dim=1)) + m.convq_glo(glo))
        return (1 - z) * net + z * q


nh = [to_nhwc(x) for x in (net, inp, corr, flow)]
flops = 2.0 * B * h * w * (3 * 9 * 448 * 128 + 128 * 128)
ours = m.forward_nhwc(*nh)
ref = cudnn_gru()
print("max |ours - cuDNN autocast| = %.3e" % (ours.permute(0, 3, 1, 2).float() - ref.float()).abs().max().item())
for name, fn in (("tcgen05 ConvGRU, NHWC in/out", lambda: m.forward_nhwc(*nh)),
                 ("tcgen05 ConvGRU, reference NCHW call (+5 layout kernels)", lambda: m(net, inp, corr, flow)),
                 ("torch / cuDNN autocast (reference path)", cudnn_gru)):
    us = t(fn)
    print("%-58s %9.1f us  %7.1f TFLOP/s" % (name, us, flops / us / 1e6))
