"""three forwards of goslam_b200.droid_net.UpdateModule at the config-2 size, for `ncu --metrics gpu__time_duration.sum`"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from goslam_b200.droid_net import UpdateModule
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
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    m(net, inp, corr, flow, ii, jj)
torch.cuda.synchronize()
print("done")
