"""`DepthVideo` — the keyframe store the tracking / backend / mapping threads share, with the reference's
constructor, attributes and method signatures (src/depth_video.py:12-269) on top of the sm_100a kernels.

What the reference composes from lietorch + ~15 eager torch kernels or two droid_backends calls is one
launch here:
    reproject(ii, jj)             -> goslam_reproject           (src/depth_video.py:207-217)
    distance(ii, jj, beta, bidir) -> goslam_frame_distance[_bidir]  (:219-255)
    ba(target, weight, eta, ...)  -> goslam_ba, in place on the shared poses / disps, then the
                                     reference's clamp_(min=0.001)  (:257-269)
    upsample(ix, mask)            -> goslam_cvx_upsample        (:194-196)
State tensors keep the reference's names, shapes and dtypes (the other processes index them directly).
There is no CPU path: the buffers live on `args.device`, which must be a CUDA device.
"""
import torch
from torch.multiprocessing import Value

from . import droid_backends
from . import lietorch
from .droid_net import cvx_upsample

_IDENTITY = (0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0)


class DepthVideo:
    takes_frame_eta = True            # ba(..., eta_by_frame=True) is understood (FactorGraph checks this)

    # (attribute, trailing shape builder, dtype, initial value) — src/depth_video.py:39-72
    @staticmethod
    def _state_table(ht, wd, c):
        s8 = (ht // 8, wd // 8)
        return [
            ("timestamp", (), torch.float, 0), ("images", (3, ht, wd), torch.float, 0),
            ("dirty", (), torch.bool, 0), ("red", (), torch.bool, 0),
            ("poses", (7,), torch.float, _IDENTITY), ("poses_gt", (4, 4), torch.float, "eye"),
            ("disps", s8, torch.float, 1), ("disps_sens", s8, torch.float, 0),
            ("depths_gt", (ht, wd), torch.float, 0), ("disps_up", (ht, wd), torch.float, 0),
            ("intrinsics", (4,), torch.float, 0),
            ("fmaps", (c, 128) + s8, torch.half, 0), ("nets", (128,) + s8, torch.half, 0),
            ("inps", (128,) + s8, torch.half, 0),
            ("poses_filtered", (7,), torch.float, _IDENTITY), ("disps_filtered", (ht, wd), torch.float, 0),
            ("mask_filtered", (ht, wd), torch.float, 0), ("update_priority", (), torch.float, 0),
        ]

    def __init__(self, cfg, args):
        self.cfg, self.args = cfg, args
        self.counter, self.ready, self.mapping = Value("i", 0), Value("i", 0), Value("i", 0)
        self.ba_lock = {"dense": Value("i", 0), "loop": Value("i", 0)}
        self.global_ba_lock = Value("i", 0)
        self.ht, self.wd = cfg["cam"]["H_out"], cfg["cam"]["W_out"]
        self.stereo = cfg["mode"] == "stereo"
        self.device = device = args.device
        if torch.device(device).type != "cuda":
            raise RuntimeError("goslam_b200.DepthVideo: a CUDA device is required (no CPU fallback)")
        self.scale_factor = 8
        buffer = cfg["tracking"]["buffer"]
        for name, tail, dtype, init in self._state_table(self.ht, self.wd, 2 if self.stereo else 1):
            t = torch.zeros((buffer,) + tuple(tail), device=device, dtype=dtype)
            if init == "eye":
                t[:] = torch.eye(4, device=device)
            elif isinstance(init, tuple):
                t[:] = torch.tensor(init, device=device)
            elif init:
                t.fill_(init)
            setattr(self, name, t if name == "images" else t.share_memory_())
        self.filtered_id = torch.tensor([-1], dtype=torch.int, device=device).share_memory_()
        self.bound = torch.zeros(1, 3, 2, device=device).share_memory_()
        self.pose_compensate = torch.tensor([_IDENTITY], device=device).share_memory_()

    # ---- locks -------------------------------------------------------------------------------
    def get_lock(self):
        return self.counter.get_lock()

    def get_ba_lock(self, ba_type):
        return self.ba_lock[ba_type].get_lock()

    def get_mapping_lock(self):
        return self.mapping.get_lock()

    # ---- item access (src/depth_video.py:85-142) ---------------------------------------------
    _ITEM_FIELDS = {0: "timestamp", 1: "images", 6: "fmaps", 7: "nets", 8: "inps"}

    def _store(self, index, item):
        if isinstance(index, int) and index >= self.counter.value:
            self.counter.value = index + 1
        elif isinstance(index, torch.Tensor) and index.max().item() > self.counter.value:
            self.counter.value = index.max().item() + 1
        for pos, name in self._ITEM_FIELDS.items():
            if pos < len(item) and (pos < 2 or item[pos] is not None):
                getattr(self, name)[index] = item[pos]
        if item[2] is not None:
            self.poses[index] = item[2]
        if item[3] is not None:
            self.disps[index] = item[3]
        if item[4] is not None:                                   # sensor depth: full-res map + 1/8 inverse depth
            self.depths_gt[index] = item[4]
            depth = item[4][..., 3::8, 3::8]
            self.disps_sens[index] = torch.where(depth > 0, 1.0 / depth, depth)
            self.disps[index] = self.disps_sens[index].clone()
        if item[5] is not None:
            self.intrinsics[index] = item[5]
        if len(item) > 9 and item[9] is not None:
            self.poses_gt[index] = item[9].to(self.poses_gt.device)

    def __setitem__(self, index, item):
        with self.get_lock():
            self._store(index, item)

    def __getitem__(self, index):
        with self.get_lock():
            if isinstance(index, int) and index > 0:              # (sic) the reference's "negative index" rule, :129
                index = self.counter.value + index
            return tuple(getattr(self, n)[index] for n in ("poses", "disps", "intrinsics", "fmaps", "nets", "inps"))

    def append(self, *item):
        with self.get_lock():
            self._store(self.counter.value, item)

    def get_bound(self):
        with self.mapping.get_lock():
            return self.bound[0]

    # ---- mapping hand-over (src/depth_video.py:151-180) --------------------------------------
    def get_mapping_item(self, index, device="cuda:0", decay=0.1):
        with self.mapping.get_lock():
            image = self.images[index].clone().permute(1, 2, 0).contiguous().to(device)
            mask = self.mask_filtered[index].clone().to(device)
            depth = 1.0 / (self.disps_filtered[index].clone().to(device) + 1e-7)
            w2c = lietorch.SE3(self.poses_filtered[index].clone()).to(device)
            c2w = (lietorch.SE3(self.pose_compensate[0].clone()).to(w2c.device) * w2c.inv()).matrix()
            gt_c2w = self.poses_gt[index].clone().to(device)
            self.update_priority[index] *= decay
            return image, depth, c2w, gt_c2w, mask

    def set_item_from_mapping(self, index, pose=None, depth=None):
        with self.get_lock():
            pass

    # ---- geometric operations ----------------------------------------------------------------
    @staticmethod
    def format_indices(ii, jj, device="cuda"):
        """to device, long, flat (src/depth_video.py:184-192)"""
        out = []
        for x in (ii, jj):
            x = x if isinstance(x, torch.Tensor) else torch.as_tensor(x)
            out.append(x.to(device=device, dtype=torch.long).reshape(-1))
        return out

    def upsample(self, ix, mask):
        up = cvx_upsample(self.disps[ix].unsqueeze(dim=-1), mask)       # [b, 8h, 8w, 1]
        self.disps_up[ix] = up.squeeze()

    def normalize(self):
        with self.get_lock():
            n = self.counter.value
            s = self.disps[:n].mean()
            self.disps[:n] /= s
            self.poses[:n, :3] *= s
            self.dirty[:n] = True

    def reproject(self, ii, jj):
        """coords [1,N,h,w,2], valid [1,N,h,w,1] of projecting frame ii's pixels into frame jj"""
        ii, jj = self.format_indices(ii, jj, self.device)
        return droid_backends.reproject(self.poses, self.disps, self.intrinsics, ii, jj)

    def distance(self, ii=None, jj=None, beta=0.3, bidirectional=True):
        """mean reprojection flow magnitude per frame pair (src/depth_video.py:219-255)"""
        matrix = ii is None
        n = self.counter.value
        if matrix:
            ii, jj = torch.meshgrid(torch.arange(n), torch.arange(n), indexing="ij")
        ii, jj = self.format_indices(ii, jj, self.device)
        intr = self.intrinsics[0]                                  # one camera per scene
        if bidirectional:
            # the reference runs both directions on a clone of poses[:counter]; reading the live buffer
            # is the same data (nothing writes poses between its two launches either)
            d = droid_backends.frame_distance_bidirectional(self.poses, self.disps, intr, ii, jj, beta)
        else:
            d = droid_backends.frame_distance(self.poses, self.disps, intr, ii, jj, beta)
        return d.reshape(n, n) if matrix else d

    def ba(self, target, weight, eta, ii, jj, t0=1, t1=None, iters=2, lm=1e-4, ep=0.1, motion_only=False,
           ba_type=None, eta_by_frame=False):
        """dense bundle adjustment, in place on poses / disps (src/depth_video.py:257-269).
        eta_by_frame (extension): eta is [buffer, h, w] indexed by frame id instead of the packed rows."""
        lock = self.get_lock() if ba_type is None else self.get_ba_lock(ba_type)
        with lock:
            if t1 is None:
                t1 = max(ii.max().item(), jj.max().item()) + 1
            droid_backends.ba(self.poses, self.disps, self.intrinsics[0], self.disps_sens, target, weight, eta,
                              ii, jj, t0, t1, iters, lm, ep, motion_only, eta_by_frame=eta_by_frame)
            self.disps.clamp_(min=0.001)
