"""The two upsampling helpers of src/droid_net.py (:9-32) on the sm_100a kernel.

`cvx_upsample(data, mask)` and `upsample_disp(disp, mask)` keep the reference's signatures;
DepthVideo.upsample (src/depth_video.py:194-196) is `cvx_upsample(disps[ix].unsqueeze(-1), mask)`.
One launch (goslam_cvx_upsample) replaces permute + softmax + unfold + mul + sum + permute.
"""
import torch

from . import _lib


def cvx_upsample(data, mask):
    """data [b, ht, wd, dim] (dim <= 4), mask [b, 9*8*8, ht, wd] (f16 or f32) -> [b, 8ht, 8wd, dim] f32."""
    if not data.is_cuda:
        raise RuntimeError("cvx_upsample: CUDA tensors required (no CPU fallback)")
    batch, ht, wd, dim = data.shape
    d = data.float().contiguous()
    m = mask.reshape(batch, 576, ht, wd).contiguous()
    if m.dtype not in (torch.float16, torch.float32):
        m = m.float()
    out = torch.empty((batch, 8 * ht, 8 * wd, dim), dtype=torch.float32, device=data.device)
    with torch.cuda.device(data.device):
        rc = _lib.load().goslam_cvx_upsample(_lib.ptr(d), _lib.ptr(m), 1 if m.dtype == torch.float16 else 0,
                                             _lib.ptr(out), batch, ht, wd, dim, _lib.stream_ptr())
    _lib.check(rc, "cvx_upsample")
    return out


def upsample_disp(disp, mask):
    batch, num, ht, wd = disp.shape
    disp = disp.reshape(batch * num, ht, wd, 1)
    mask = mask.reshape(batch * num, -1, ht, wd)
    return cvx_upsample(disp, mask).view(batch, num, 8 * ht, 8 * wd)


# ------------------------------------------------------------------------------------------------------
# The update operator: UpdateModule / GraphAgg with the reference's parameter names
# (src/droid_net.py:33-140), so the `update.*` entries of pretrained/droid.pth load with load_state_dict.
# The ConvGRU — two thirds of the operator's arithmetic — runs on the tcgen05 kernel (modules/gru.py); the
# encoders, heads and the per-frame aggregation are still torch convolutions, evaluated channels-last so that
# the GRU reads and writes their tensors in place (a channels-last [B,C,h,w] tensor IS the kernel's NHWC).
# ------------------------------------------------------------------------------------------------------
import torch.nn as nn  # noqa: E402

from .modules.gru import ConvGRU  # noqa: E402


class _Identity(nn.Module):
    """GradientClip (src/modules/clipping.py) only acts in backward; inference sees the identity.  Kept as a
    module so that Sequential indices — and therefore state_dict keys — match the reference."""

    def forward(self, x):
        return x


def _nhwc_view(x):
    """[B,C,h,w] channels-last tensor -> its [B,h,w,C] contiguous view (a copy only if it was not channels-last)"""
    v = x.permute(0, 2, 3, 1)
    return v if v.is_contiguous() else v.contiguous()


class GraphAgg(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv1 = nn.Conv2d(128, 128, 3, padding=1)
        self.conv2 = nn.Conv2d(128, 128, 3, padding=1)
        self.relu = nn.ReLU(inplace=True)
        self.eta = nn.Sequential(nn.Conv2d(128, 1, 3, padding=1), _Identity(), nn.Softplus())
        self.upmask = nn.Sequential(nn.Conv2d(128, 8 * 8 * 9, 1))

    def forward(self, net, ii):
        """net [batch, num, 128, h, w], ii [num] -> 0.01 * eta [batch, M, h, w], upmask [batch, M, 576, h, w] with one
        row per distinct source frame, in sorted frame order (src/droid_net.py:51-67; scatter_mean = index_add / count)"""
        batch, num, ch, ht, wd = net.shape
        frames, ix = torch.unique(ii, sorted=True, return_inverse=True)
        x = self.relu(self.conv1(net.reshape(batch * num, ch, ht, wd))).view(batch, num, 128, ht, wd)
        m = frames.numel()
        agg = torch.zeros((batch, m, 128, ht, wd), dtype=x.dtype, device=x.device).index_add_(1, ix, x)
        cnt = torch.zeros(m, dtype=x.dtype, device=x.device).index_add_(0, ix, torch.ones_like(ix, dtype=x.dtype))
        x = (agg / cnt.view(1, m, 1, 1, 1)).view(-1, 128, ht, wd)
        x = self.relu(self.conv2(x))
        eta = self.eta(x).view(batch, -1, ht, wd)
        upmask = self.upmask(x).view(batch, -1, 8 * 8 * 9, ht, wd)
        return 0.01 * eta, upmask


class UpdateModule(nn.Module):
    def __init__(self):
        super().__init__()
        cor_planes = 4 * (2 * 3 + 1) ** 2
        self.corr_encoder = nn.Sequential(nn.Conv2d(cor_planes, 128, 1), nn.ReLU(inplace=True),
                                          nn.Conv2d(128, 128, 3, padding=1), nn.ReLU(inplace=True))
        self.flow_encoder = nn.Sequential(nn.Conv2d(4, 128, 7, padding=3), nn.ReLU(inplace=True),
                                          nn.Conv2d(128, 64, 3, padding=1), nn.ReLU(inplace=True))
        self.weight = nn.Sequential(nn.Conv2d(128, 128, 3, padding=1), nn.ReLU(inplace=True),
                                    nn.Conv2d(128, 2, 3, padding=1), _Identity(), nn.Sigmoid())
        self.delta = nn.Sequential(nn.Conv2d(128, 128, 3, padding=1), nn.ReLU(inplace=True),
                                   nn.Conv2d(128, 2, 3, padding=1), _Identity())
        self.gru = ConvGRU(128, 128 + 128 + 64)
        self.agg = GraphAgg()

    @torch.no_grad()
    def forward(self, net, inp, corr, flow=None, ii=None, jj=None):
        """net, inp [batch, num, 128, h, w]; corr [batch, num, 196, h, w]; flow [batch, num, 4, h, w] ->
        net', delta [batch, num, h, w, 2], weight [batch, num, h, w, 2] (, eta, upmask when ii is given)
        (src/droid_net.py:107-140)."""
        batch, num, ch, ht, wd = net.shape
        if not net.is_cuda:
            raise RuntimeError("UpdateModule: CUDA tensors required (no CPU fallback)")
        if flow is None:
            flow = torch.zeros(batch, num, 4, ht, wd, device=net.device)
        cl = torch.channels_last
        with torch.autocast("cuda", enabled=True):
            def flat(t):
                return t.reshape(batch * num, -1, ht, wd).contiguous(memory_format=cl)
            c = self.corr_encoder(flat(corr))
            f = self.flow_encoder(flat(flow))
            state = self.gru.forward_nhwc(_nhwc_view(flat(net).half()), _nhwc_view(flat(inp).half()),
                                          _nhwc_view(c.half()), _nhwc_view(f.half()))
            x = state.permute(0, 3, 1, 2)                       # channels-last [B,128,h,w] view of the NHWC state
            delta = self.delta(x).view(batch, num, -1, ht, wd).permute(0, 1, 3, 4, 2)[..., :2].contiguous()
            weight = self.weight(x).view(batch, num, -1, ht, wd).permute(0, 1, 3, 4, 2)[..., :2].contiguous()
            net_out = x.contiguous().view(batch, num, -1, ht, wd)        # the graph stores `net` as [batch,num,128,h,w]
            if ii is None:
                return net_out, delta, weight
            eta, upmask = self.agg(net_out, ii.to(net.device))
            return net_out, delta, weight, eta, upmask
