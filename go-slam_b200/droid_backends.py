"""`droid_backends` — the reference's native module name and its nine functions
(src/lib/droid.cpp:237-250), bound to libgoslam_b200.so through the C-ABI.

Same names, argument order, dtypes, in-place semantics and error behaviour as the pybind
module the reference builds from src/lib/*.cu:
  * every wrapper only checks contiguity (`TORCH_CHECK(x.is_contiguous())`, droid.cpp:84-85)
    and raises RuntimeError — plus a CUDA-device check, because there is no CPU path here;
  * `ba` mutates `poses` / `disps` in place through the caller's storage
    (src/lib/droid_kernels.cu:1389-1391,1420-1428) and returns [dx, dz];
  * the backward entry points exist and raise (inference is torch.no_grad, src/slam.py:45).

Install under the reference's import name with `goslam_b200.install()`.
"""
import ctypes

import torch

from . import _lib

_ws_cache = {}


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("goslam_b200.droid_backends: tensors must live on a CUDA device "
                               "(there is no CPU fallback)")


def _contig(**kw):
    for name, t in kw.items():
        if not t.is_contiguous():
            raise RuntimeError("%s must be contiguous" % name)


def _workspace(nbytes, device):
    """grow-only per-device scratch (borrowed for the duration of one call on the current stream)."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


def _dtype_code(t):
    if t.dtype == torch.float16:
        return 1
    if t.dtype == torch.float32:
        return 0
    raise RuntimeError("correlation volume must be float16 or float32 (got %s)" % t.dtype)


# ----------------------------------------------------------------------------- correlation
def corr_index_forward(volume, coords, radius):
    """src/lib/droid.cpp:170-178 — volume [N,h1,w1,h2,w2], coords [N,2,h1,w1] -> [corr]."""
    _contig(volume=volume, coords=coords)
    _need_cuda(volume, coords)
    N, h1, w1, h2, w2 = volume.shape
    rd = 2 * radius + 1
    corr = torch.empty((N, rd, rd, h1, w1), dtype=volume.dtype, device=volume.device)
    coords_f = coords.float()          # named: a converted copy must outlive the launch
    with torch.cuda.device(volume.device):
        rc = _lib.load().goslam_corr_index_forward(
            _lib.ptr(volume), _dtype_code(volume), _lib.ptr(coords_f), _lib.ptr(corr),
            N, h1, w1, h2, w2, int(radius), _lib.stream_ptr())
    _lib.check(rc, "corr_index_forward")
    return [corr]


def corr_index_backward(volume, coords, corr_grad, radius):
    _contig(volume=volume, coords=coords, corr_grad=corr_grad)
    raise RuntimeError("corr_index_backward: training-only entry point, not part of the "
                       "inference hot path (GOSLAM_EUNSUPPORTED)")


def altcorr_forward(fmap1, fmap2, coords, radius):
    """src/lib/droid.cpp:193-203 — fmap1 [B,H,W,C], fmap2 [B,H2,W2,C], coords [B,S,H,W,2]."""
    _contig(fmap1=fmap1, fmap2=fmap2, coords=coords)
    _need_cuda(fmap1, fmap2, coords)
    if fmap1.dtype != torch.float32 or fmap2.dtype != torch.float32:
        raise RuntimeError("altcorr_forward expects float32 feature maps (as the reference calls it, "
                           "src/modules/corr.py:125)")
    B, H, W, C = fmap1.shape
    _, H2, W2, _ = fmap2.shape
    S = coords.shape[1]
    rd = 2 * radius + 1
    corr = torch.empty((B, S, rd * rd, H, W), dtype=torch.float32, device=fmap1.device)
    with torch.cuda.device(fmap1.device):
        rc = _lib.load().goslam_altcorr_forward(
            _lib.ptr(fmap1), _lib.ptr(fmap2), _lib.ptr(coords), _lib.ptr(corr),
            B, S, H, W, H2, W2, C, int(radius), _lib.stream_ptr())
    _lib.check(rc, "altcorr_forward")
    return [corr]


def altcorr_backward(fmap1, fmap2, coords, corr_grad, radius):
    _contig(fmap1=fmap1, fmap2=fmap2, coords=coords, corr_grad=corr_grad)
    raise RuntimeError("altcorr_backward: training-only entry point, not part of the inference "
                       "hot path (GOSLAM_EUNSUPPORTED)")


# ----------------------------------------------------------------------------- geometry
def frame_distance(poses, disps, intrinsics, ii, jj, beta):
    """src/lib/droid.cpp:120-126 -> Tensor[K]."""
    _contig(poses=poses, disps=disps, intrinsics=intrinsics, ii=ii, jj=jj)
    _need_cuda(poses, disps, intrinsics, ii, jj)
    K = ii.shape[0]
    ht, wd = disps.shape[1], disps.shape[2]
    dist = torch.empty((K,), dtype=torch.float32, device=poses.device)
    with torch.cuda.device(poses.device):
        rc = _lib.load().goslam_frame_distance(
            _lib.ptr(poses), _lib.ptr(disps), _lib.ptr(intrinsics), _lib.ptr(ii), _lib.ptr(jj),
            _lib.ptr(dist), K, ht, wd, float(beta), _lib.stream_ptr())
    _lib.check(rc, "frame_distance")
    return dist


def frame_distance_bidirectional(poses, disps, intrinsics, ii, jj, beta):
    """Not in the reference module: DepthVideo.distance(bidirectional=True) (src/depth_video.py:233-245)
    = 0.5 * (frame_distance(ii, jj) + frame_distance(jj, ii)) in one launch, bit-identical to that form."""
    _contig(poses=poses, disps=disps, intrinsics=intrinsics, ii=ii, jj=jj)
    _need_cuda(poses, disps, intrinsics, ii, jj)
    K = ii.shape[0]
    ht, wd = disps.shape[1], disps.shape[2]
    dist = torch.empty((K,), dtype=torch.float32, device=poses.device)
    with torch.cuda.device(poses.device):
        rc = _lib.load().goslam_frame_distance_bidir(
            _lib.ptr(poses), _lib.ptr(disps), _lib.ptr(intrinsics), _lib.ptr(ii), _lib.ptr(jj),
            _lib.ptr(dist), K, ht, wd, float(beta), _lib.stream_ptr())
    _lib.check(rc, "frame_distance_bidir")
    return dist


def projmap(poses, disps, intrinsics, ii, jj):
    """src/lib/droid.cpp:139-144 -> [coords(N,h,w,3), valid(N,h,w,1)]."""
    _contig(poses=poses, disps=disps, intrinsics=intrinsics, ii=ii, jj=jj)
    _need_cuda(poses, disps, intrinsics, ii, jj)
    K = ii.shape[0]
    ht, wd = disps.shape[1], disps.shape[2]
    coords = torch.empty((K, ht, wd, 3), dtype=torch.float32, device=poses.device)
    valid = torch.empty((K, ht, wd, 1), dtype=torch.float32, device=poses.device)
    with torch.cuda.device(poses.device):
        rc = _lib.load().goslam_projmap(
            _lib.ptr(poses), _lib.ptr(disps), _lib.ptr(intrinsics), _lib.ptr(ii), _lib.ptr(jj),
            _lib.ptr(coords), _lib.ptr(valid), K, ht, wd, _lib.stream_ptr())
    _lib.check(rc, "projmap")
    return [coords, valid]


def iproj(poses, disps, intrinsics):
    """src/lib/droid.cpp:157-160 -> points [n,h,w,3]."""
    _contig(poses=poses, disps=disps, intrinsics=intrinsics)
    _need_cuda(poses, disps, intrinsics)
    num, ht, wd = disps.shape
    points = torch.empty((num, ht, wd, 3), dtype=torch.float32, device=disps.device)
    with torch.cuda.device(disps.device):
        rc = _lib.load().goslam_iproj(_lib.ptr(poses), _lib.ptr(disps), _lib.ptr(intrinsics),
                                      _lib.ptr(points), num, ht, wd, _lib.stream_ptr())
    _lib.check(rc, "iproj")
    return points


def depth_filter(poses, disps, intrinsics, ix, thresh):
    """src/lib/droid.cpp:220-225 -> counter [n,h,w]."""
    _contig(poses=poses, disps=disps, intrinsics=intrinsics, ix=ix, thresh=thresh)
    _need_cuda(poses, disps, intrinsics, ix, thresh)
    K = ix.shape[0]
    num, ht, wd = disps.shape
    counter = torch.empty((K, ht, wd), dtype=torch.float32, device=disps.device)
    with torch.cuda.device(disps.device):
        rc = _lib.load().goslam_depth_filter(
            _lib.ptr(poses), _lib.ptr(disps), _lib.ptr(intrinsics), _lib.ptr(ix), _lib.ptr(thresh),
            _lib.ptr(counter), K, num, ht, wd, _lib.stream_ptr())
    _lib.check(rc, "depth_filter")
    return counter


def reproject(poses, disps, intrinsics_all, ii, jj, want_valid=True):
    """Fused DepthVideo.reproject (src/depth_video.py:207-217): coords [1,N,h,w,2], valid [1,N,h,w,1].
    Not a droid_backends symbol upstream (the reference composes it from lietorch + torch ops)."""
    _contig(poses=poses, disps=disps, intrinsics_all=intrinsics_all, ii=ii, jj=jj)
    _need_cuda(poses, disps, intrinsics_all, ii, jj)
    K = ii.shape[0]
    ht, wd = disps.shape[1], disps.shape[2]
    coords = torch.empty((1, K, ht, wd, 2), dtype=torch.float32, device=poses.device)
    valid = torch.empty((1, K, ht, wd, 1), dtype=torch.float32, device=poses.device) if want_valid else None
    with torch.cuda.device(poses.device):
        rc = _lib.load().goslam_reproject(
            _lib.ptr(poses), _lib.ptr(disps), _lib.ptr(intrinsics_all), _lib.ptr(ii), _lib.ptr(jj),
            _lib.ptr(coords), _lib.ptr(valid), K, ht, wd, _lib.stream_ptr())
    _lib.check(rc, "reproject")
    return coords, valid


def reproject_motion(poses, disps, intrinsics_all, ii, jj, target):
    """DepthVideo.reproject + FactorGraph's motion features in one launch (src/factor_graph.py:202-206):
    returns coords1 [1,N,h,w,2] and motion [1,N,4,h,w] = clamp(cat([coords1 - coords0, target - coords1]), +-64)."""
    _contig(poses=poses, disps=disps, intrinsics_all=intrinsics_all, ii=ii, jj=jj, target=target)
    _need_cuda(poses, disps, intrinsics_all, ii, jj, target)
    K = ii.shape[0]
    ht, wd = disps.shape[1], disps.shape[2]
    if target.dtype != torch.float32 or target.numel() != K * ht * wd * 2:
        raise RuntimeError("reproject_motion: target must be float32 [1, N, ht, wd, 2]")
    coords = torch.empty((1, K, ht, wd, 2), dtype=torch.float32, device=poses.device)
    motion = torch.empty((1, K, 4, ht, wd), dtype=torch.float32, device=poses.device)
    with torch.cuda.device(poses.device):
        rc = _lib.load().goslam_reproject_motion(
            _lib.ptr(poses), _lib.ptr(disps), _lib.ptr(intrinsics_all), _lib.ptr(ii), _lib.ptr(jj),
            _lib.ptr(target), _lib.ptr(coords), None, _lib.ptr(motion), K, ht, wd, _lib.stream_ptr())
    _lib.check(rc, "reproject_motion")
    return coords, motion


# ----------------------------------------------------------------------------- bundle adjustment
def ba(poses, disps, intrinsics, disps_sens, targets, weights, eta, ii, jj,
       t0, t1, iterations, lm, ep, motion_only, return_status=False, eta_by_frame=False):
    """src/lib/droid.cpp:88-117.  In place on poses [num,7] / disps [num,ht,wd].

    Returns [dx (t1-t0, 6), dz].  dz is laid out [num, ht*wd] indexed by FRAME id (rows of
    frames that carry no depth variable are zero) instead of the reference's packed
    [len(unique frames), ht*wd]: producing the packed shape needs the size of a device-side
    unique(), i.e. a host sync per call, and every caller in the reference discards the
    return value (src/depth_video.py:266).  dz is None when motion_only (reference: undefined
    tensor).

    eta: [M, ht, wd] with M = |unique([t0,t1) U ii)| rows in sorted frame order (what FactorGraph passes,
    src/factor_graph.py:236-238), or one row.  A different row count leaves the state untouched and
    reports status 2 (the reference raises a broadcast error there).  eta_by_frame=True (not in the
    reference): eta is [num, ht, wd] indexed by frame id."""
    _contig(targets=targets, weights=weights, poses=poses, disps=disps, intrinsics=intrinsics,
            disps_sens=disps_sens, ii=ii, jj=jj)
    _need_cuda(poses, disps, intrinsics, disps_sens, targets, weights, ii, jj)
    if ii.dtype != torch.int64 or jj.dtype != torch.int64:
        raise RuntimeError("ii / jj must be int64")
    dev = poses.device
    N = int(ii.shape[0])
    num, ht, wd = disps.shape
    t0, t1 = int(t0), int(t1)
    P = max(t1 - t0, 0)
    lib = _lib.load()
    eta_c = None
    eta_rows = 0
    if not motion_only:
        eta_c = eta.contiguous().view(-1, ht * wd).float()
        eta_rows = int(eta_c.shape[0])
        if eta_by_frame:
            if eta_rows != num:
                raise RuntimeError("ba: eta_by_frame needs one row per frame (%d), got %d" % (num, eta_rows))
            eta_rows = -num
    dx = torch.zeros((P, 6), dtype=torch.float32, device=dev)
    dz = None if motion_only else torch.empty((num, ht * wd), dtype=torch.float32, device=dev)
    status = torch.zeros((max(int(iterations), 1),), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        nbytes = lib.goslam_ba_workspace_bytes(N, num, ht, wd, t0, t1)
        if nbytes == 0:
            raise RuntimeError("ba: invalid shapes (N=%d num=%d t0=%d t1=%d)" % (N, num, t0, t1))
        ws = _workspace(nbytes, dev)
        rc = lib.goslam_ba(
            _lib.ptr(poses), _lib.ptr(disps), _lib.ptr(intrinsics), _lib.ptr(disps_sens),
            _lib.ptr(targets), _lib.ptr(weights), _lib.ptr(eta_c), eta_rows,
            _lib.ptr(ii), _lib.ptr(jj), N, num, ht, wd, t0, t1, int(iterations),
            float(lm), float(ep), int(bool(motion_only)),
            _lib.ptr(dx), _lib.ptr(dz), _lib.ptr(status),
            _lib.ptr(ws), ctypes.c_size_t(ws.numel()), _lib.stream_ptr())
    _lib.check(rc, "ba")
    if return_status:
        return [dx, dz, status]
    return [dx, dz]
