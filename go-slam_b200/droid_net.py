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
