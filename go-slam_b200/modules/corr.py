"""CorrBlock / AltCorrBlock with the reference's constructor and call signatures
(src/modules/corr.py:25-65,97-145), backed by the sm_100a kernels.

CorrBlock(fmap1, fmap2)      -> tcgen05 all-pairs build + in-epilogue 4-level pyramid
                                (goslam_corr_build; reference: torch.matmul + 3x avg_pool2d)
CorrBlock.__call__(coords)   -> ONE fused 4-level radius-3 lookup (goslam_corr_pyramid_lookup;
                                reference: 4 x corr_index_forward + torch.cat)
AltCorrBlock(fmaps)(coords, ii, jj) -> windowed correlation, all levels in one launch (goslam_altcorr_pyramid)
"""
import ctypes

import torch
import torch.nn.functional as F

from .. import _lib
from ..droid_backends import _workspace


def _ptr_array(tensors):
    arr = (ctypes.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return arr


class CorrPool:
    """Slot pool for the correlation pyramids of a factor graph.

    The reference concatenates / boolean-masks the whole pyramid on every add_factors /
    rm_factors (src/modules/corr.py:55-65 via src/factor_graph.py:114,149) — a copy of up to
    N*hw*1.33*hw*2 bytes each time.  Here the graph owns `capacity` slots per level, allocated
    once (FactorGraph.max_factors is the natural capacity); a CorrBlock built `from_video(...,
    pool=pool)` holds only a table of slot ids, so cat() joins two tables and __getitem__ filters
    one.  The build and lookup kernels follow the table on the device
    (goslam_corr_pool_build / goslam_corr_pool_lookup)."""

    ROWMAJOR, TILED = 0, 1

    def __init__(self, capacity, ht, wd, num_levels=4, device="cuda", layout="tiled"):
        """layout "tiled" (default): levels 0/1 as 4x4-element tiles (one 32-byte sector each),
        private to the build and lookup kernels — nothing else in GO-SLAM reads the pyramid
        (include/goslam_b200.h, GOSLAM_LAYOUT_TILED); "rowmajor": the reference's layout."""
        self.capacity, self.ht, self.wd, self.num_levels = int(capacity), ht, wd, num_levels
        self.layout = {"tiled": self.TILED, "rowmajor": self.ROWMAJOR}[layout]
        self.plane_elems = [self._plane_elems(i) for i in range(num_levels)]
        self.levels = [torch.empty((self.capacity, ht * wd, self.plane_elems[i]), dtype=torch.float16, device=device)
                       for i in range(num_levels)]
        self._free = list(range(self.capacity - 1, -1, -1))     # stack; slot 0 is handed out first
        self._iota = torch.arange(self.capacity, dtype=torch.int32, device=device)

    def slot_table(self, slots):
        """device int32 table for a list of slot ids (a view of a resident iota when the ids are a
        consecutive run — the common case — so that no host->device copy is needed)."""
        n = len(slots)
        if n and slots[-1] - slots[0] == n - 1 and all(slots[i + 1] == slots[i] + 1 for i in range(n - 1)):
            return self._iota[slots[0]:slots[0] + n]
        return torch.tensor(slots, dtype=torch.int32, device=self._iota.device)

    def _plane_elems(self, i):
        hl, wl = self.ht >> i, self.wd >> i
        if self.layout == self.TILED:                             # == goslam_corr_level_plane_elems
            if i < 2:
                return ((hl + 3) // 4) * ((wl + 3) // 4) * 16
            n_yb, n_xb = (self.ht + 7) // 8, (self.wd + 15) // 16
            return n_yb * ((n_xb * 8 + 15) // 16 * 16) if i == 2 else n_yb * 16
        return hl * wl

    def level_rowmajor(self, i, slots=None):
        """level i as [n, ht, wd, ht>>i, wd>>i] (a gathered, de-tiled copy; tests / debugging)."""
        lvl = self.levels[i] if slots is None else self.levels[i][slots]
        hl, wl = self.ht >> i, self.wd >> i
        n = lvl.shape[0]
        if self.layout == self.TILED and i < 2:
            h4, w4 = (hl + 3) // 4, (wl + 3) // 4
            lvl = lvl.view(n, self.ht, self.wd, h4, w4, 4, 4).permute(0, 1, 2, 3, 5, 4, 6)
            return lvl.reshape(n, self.ht, self.wd, 4 * h4, 4 * w4)[..., :hl, :wl].contiguous()
        if self.layout == self.TILED:
            n_yb, n_xb = (self.ht + 7) // 8, (self.wd + 15) // 16
            if i == 2:          # per band: 2 rows of n_xb*4 columns, padded to a multiple of 16 elements
                lvl = lvl.view(n, self.ht, self.wd, n_yb, -1)[..., :2 * n_xb * 4]
                return lvl.reshape(n, self.ht, self.wd, 2 * n_yb, n_xb * 4)[..., :hl, :wl].contiguous()
            return lvl.view(n, self.ht, self.wd, n_yb, 16)[..., :hl, :wl].contiguous()
        return lvl.view(n, self.ht, self.wd, hl, wl)

    @property
    def free_slots(self):
        return len(self._free)

    def alloc(self, n):
        if n > len(self._free):
            raise RuntimeError("CorrPool exhausted: %d slots requested, %d free of %d"
                               % (n, len(self._free), self.capacity))
        out = self._free[len(self._free) - n:][::-1]
        del self._free[len(self._free) - n:]
        return out

    def release(self, slots):
        self._free.extend(reversed(list(slots)))

    def grow(self, capacity):
        """enlarge the pool (graphs created with max_factors = -1 have no bound, e.g. PoseTrajectoryFiller's,
        src/trajectory_filler.py:63): new buffers, one copy of the old slots, slot ids stay valid."""
        capacity = int(capacity)
        if capacity <= self.capacity:
            return
        dev = self.levels[0].device
        for i, old in enumerate(self.levels):
            new = torch.empty((capacity,) + tuple(old.shape[1:]), dtype=old.dtype, device=dev)
            new[:self.capacity].copy_(old)
            self.levels[i] = new
        self._free = list(range(capacity - 1, self.capacity - 1, -1)) + self._free
        self._iota = torch.arange(capacity, dtype=torch.int32, device=dev)
        self.capacity = capacity


class CorrBlock:
    pool = None          # set for slot-pool blocks (see CorrPool); then `slots`/`_slots_host` exist

    def __init__(self, fmap1, fmap2, num_levels=4, radius=3, impl=0):
        self.num_levels = num_levels
        self.radius = radius
        if not fmap1.is_cuda:
            raise RuntimeError("CorrBlock: CUDA tensors required (no CPU fallback)")
        batch, num, dim, ht, wd = fmap1.shape
        N = batch * num
        self.ht, self.wd = ht, wd
        dev = fmap1.device
        f1 = fmap1.reshape(N, dim, ht, wd).contiguous()
        f2 = fmap2.reshape(N, dim, ht, wd).contiguous()
        lib = _lib.load()
        if f1.dtype == torch.float16:
            f2 = f2.half()             # named: a converted copy must outlive the launch
            levels = [torch.empty((N, ht, wd, ht >> i, wd >> i), dtype=torch.float16, device=dev)
                      for i in range(num_levels)]
            with torch.cuda.device(dev):
                nbytes = lib.goslam_corr_build_workspace_bytes(N, dim, ht, wd)
                ws = _workspace(nbytes, dev)
                rc = lib.goslam_corr_build(
                    _lib.ptr(f1), _lib.ptr(f2), _ptr_array(levels), num_levels, N, dim, ht, wd,
                    int(impl), _lib.ptr(ws), ctypes.c_size_t(ws.numel()), _lib.stream_ptr())
            _lib.check(rc, "corr_build")
        else:
            f1, f2 = f1.float(), f2.float()
            levels = [torch.empty((N, ht, wd, ht >> i, wd >> i), dtype=torch.float32, device=dev)
                      for i in range(num_levels)]
            with torch.cuda.device(dev):
                rc = lib.goslam_corr_build_f32(_lib.ptr(f1), _lib.ptr(f2), _ptr_array(levels),
                                               num_levels, N, dim, ht, wd, _lib.stream_ptr())
            _lib.check(rc, "corr_build_f32")
        self.corr_pyramid = levels

    @classmethod
    def from_video(cls, fmaps_kmajor, ii, jj, ht, wd, rig=1, num_levels=4, radius=3, pool=None):
        """FactorGraph.add_factors' volume for edges (ii, jj) straight from video-level K-major
        feature maps [buffer*rig, ht*wd, 128] (see `fmaps_to_kmajor`): no gathered copies.
        With `pool` the volumes are written into free slots of that CorrPool."""
        self = cls.__new__(cls)
        self.num_levels, self.radius, self.ht, self.wd = num_levels, radius, ht, wd
        dev = fmaps_kmajor.device
        N = int(ii.shape[0])
        F = int(fmaps_kmajor.shape[0])
        if pool is not None:
            if (pool.ht, pool.wd, pool.num_levels) != (ht, wd, num_levels):
                raise RuntimeError("CorrBlock.from_video: pool shape mismatch")
            self.pool = pool
            self._slots_host = pool.alloc(N)
            self.slots = pool.slot_table(self._slots_host)
            with torch.cuda.device(dev):
                rc = _lib.load().goslam_corr_pool_build(
                    _lib.ptr(fmaps_kmajor), F, int(rig), _lib.ptr(ii), _lib.ptr(jj), _lib.ptr(self.slots),
                    pool.layout, _ptr_array(pool.levels), num_levels, N, 128, ht, wd, _lib.stream_ptr())
            _lib.check(rc, "corr_pool_build")
            return self
        levels = [torch.empty((N, ht, wd, ht >> i, wd >> i), dtype=torch.float16, device=dev)
                  for i in range(num_levels)]
        with torch.cuda.device(dev):
            rc = _lib.load().goslam_corr_build_indexed(
                _lib.ptr(fmaps_kmajor), F, int(rig), _lib.ptr(ii), _lib.ptr(jj), _ptr_array(levels),
                num_levels, N, 128, ht, wd, _lib.stream_ptr())
        _lib.check(rc, "corr_build_indexed")
        self.corr_pyramid = levels
        return self

    def __call__(self, coords):
        batch, num, ht, wd, _ = coords.shape
        N = batch * num
        if self.pool is not None:
            return self._call_pooled(coords, batch, num, ht, wd)
        vol0 = self.corr_pyramid[0]
        rd = 2 * self.radius + 1
        coords = coords.reshape(N, ht, wd, 2).contiguous().float()
        out = torch.empty((batch, num, self.num_levels * rd * rd, ht, wd), dtype=vol0.dtype,
                          device=vol0.device)
        pyr = [p.contiguous() for p in self.corr_pyramid]
        with torch.cuda.device(vol0.device):
            rc = _lib.load().goslam_corr_pyramid_lookup(
                _ptr_array(pyr), 1 if vol0.dtype == torch.float16 else 0, self.num_levels,
                _lib.ptr(coords), _lib.ptr(out), N, ht, wd, vol0.shape[3], vol0.shape[4],
                int(self.radius), _lib.stream_ptr())
        _lib.check(rc, "corr_pyramid_lookup")
        return out

    def _call_pooled(self, coords, batch, num, ht, wd):
        N = batch * num
        if N != len(self._slots_host):
            raise RuntimeError("CorrBlock: %d coordinate maps for %d edges" % (N, len(self._slots_host)))
        pool = self.pool
        rd = 2 * self.radius + 1
        dev = pool.levels[0].device
        coords = coords.reshape(N, ht, wd, 2).contiguous().float()
        out = torch.empty((batch, num, self.num_levels * rd * rd, ht, wd), dtype=torch.float16, device=dev)
        with torch.cuda.device(dev):
            rc = _lib.load().goslam_corr_pool_lookup(
                _ptr_array(pool.levels), 1, self.num_levels, _lib.ptr(self.slots), pool.capacity,
                pool.layout, _lib.ptr(coords), _lib.ptr(out), N, ht, wd, pool.ht, pool.wd, int(self.radius),
                _lib.stream_ptr())
        _lib.check(rc, "corr_pool_lookup")
        return out

    def cat(self, other):
        if self.pool is not None:
            # O(edges): join the slot tables; `other` gives up its slots
            if other.pool is not self.pool:
                raise RuntimeError("CorrBlock.cat: blocks live in different pools")
            self._slots_host = self._slots_host + other._slots_host
            self.slots = torch.cat([self.slots, other.slots])
            other._slots_host, other.slots = [], other.slots[:0]
            return self
        for i in range(self.num_levels):
            self.corr_pyramid[i] = torch.cat([self.corr_pyramid[i], other.corr_pyramid[i]], dim=0)
        return self

    def __getitem__(self, index):
        if self.pool is not None:
            # O(edges): keep the selected slot ids, hand the others back to the pool.  Any index
            # form torch accepts on dim 0 works (FactorGraph passes boolean masks).
            ids = torch.arange(len(self._slots_host))[index.cpu() if torch.is_tensor(index) else index]
            keep = [int(i) for i in ids.reshape(-1).tolist()]
            kept = set(keep)
            if len(kept) != len(keep):
                raise RuntimeError("CorrBlock[index]: a pooled block cannot hold one slot twice")
            self.pool.release(s for i, s in enumerate(self._slots_host) if i not in kept)
            self._slots_host = [self._slots_host[i] for i in keep]
            self.slots = self.pool.slot_table(self._slots_host)
            return self
        for i in range(self.num_levels):
            self.corr_pyramid[i] = self.corr_pyramid[i][index]
        return self

    def free(self):
        """return every slot to the pool (FactorGraph drops `self.corr` when it clears edges)."""
        if self.pool is not None and self._slots_host:
            self.pool.release(self._slots_host)
            self._slots_host, self.slots = [], self.slots[:0]

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def gather_pyramid(self):
        """materialise [N,h,w,h>>i,w>>i] per level (tests / code that reads .corr_pyramid)."""
        if self.pool is None:
            return self.corr_pyramid
        idx = self.slots.long()
        return [self.pool.level_rowmajor(i, idx) for i in range(self.num_levels)]

    @staticmethod
    def corr(fmap1, fmap2):
        """all-pairs correlation, level 0 only ([batch, num, h, w, h, w])."""
        blk = CorrBlock(fmap1, fmap2, num_levels=1)
        batch, num, _, ht, wd = fmap1.shape
        return blk.corr_pyramid[0].view(batch, num, ht, wd, ht, wd)


def fmaps_to_kmajor(fmaps, out=None):
    """DepthVideo.fmaps [buffer, rig, 128, h, w] f16 -> K-major [buffer*rig, h*w, 128] (done once
    per inserted keyframe; `out` lets the caller convert just the new rows in place)."""
    if not fmaps.is_cuda or fmaps.dtype != torch.float16:
        raise RuntimeError("fmaps_to_kmajor: CUDA float16 tensor required")
    f = fmaps.contiguous()
    h, w = f.shape[-2], f.shape[-1]
    F = f.numel() // (128 * h * w)
    if out is None:
        out = torch.empty((F, h * w, 128), dtype=torch.float16, device=f.device)
    with torch.cuda.device(f.device):
        rc = _lib.load().goslam_fmaps_to_kmajor(_lib.ptr(f), _lib.ptr(out), F, 128, h, w, _lib.stream_ptr())
    _lib.check(rc, "fmaps_to_kmajor")
    return out


class AltCorrBlock:
    """Windowed correlation without a volume (global BA / `update_lowmem`): same constructor and call
    signature as src/modules/corr.py:97-145.  `pyramid[l]` = NHWC feature maps of level l, pre-scaled by
    1/4 and 2x2 average-pooled l times, shape [B, N, H >> l, W >> l, C]."""

    def __init__(self, fmaps, num_levels=4, radius=3):
        self.num_levels, self.radius = num_levels, radius
        b, n, c, h, w = fmaps.shape
        level = fmaps.reshape(b * n, c, h, w) * 0.25            # exact power-of-two scaling, == fmaps / 4.0
        self.pyramid = []
        for lvl in range(num_levels):
            nhwc = level.permute(0, 2, 3, 1).contiguous()
            self.pyramid.append(nhwc.view(b, n, h >> lvl, w >> lvl, c))
            if lvl + 1 < num_levels:
                level = F.avg_pool2d(level, kernel_size=2, stride=2)

    def __call__(self, coords, ii, jj):
        """coords [B, N, H, W, 2] (what FactorGraph.update_lowmem passes) or [B, N, H, W, S, 2] (S coordinate
        sets per edge, the reference's general form): [B, N, L*(2r+1)^2, H, W(, S)] in float32."""
        if not coords.is_cuda:
            raise RuntimeError("AltCorrBlock: CUDA tensors required (no CPU fallback)")
        if coords.dim() == 5:
            return self._fused(coords, ii, jj)
        return torch.stack([self._fused(coords[..., s, :], ii, jj) for s in range(coords.shape[-2])], dim=-1)

    def _fused(self, coords, ii, jj):
        """one launch for all levels, feature maps indexed per edge on the device."""
        B, N, H, W, _ = coords.shape
        if B != 1:
            raise RuntimeError("AltCorrBlock fused path expects batch 1 (as GO-SLAM uses it)")
        C = self.pyramid[0].shape[-1]
        dev = coords.device
        rd = 2 * self.radius + 1
        out = torch.empty((1, N, self.num_levels * rd * rd, H, W), dtype=torch.float32, device=dev)
        pyr = [p.contiguous() if p.dtype == torch.float16 else p.half().contiguous() for p in self.pyramid]
        ii = torch.as_tensor(ii, device=dev).long().contiguous()
        jj = torch.as_tensor(jj, device=dev).long().contiguous()
        c = coords.reshape(N, H, W, 2).float().contiguous()
        with torch.cuda.device(dev):
            rc = _lib.load().goslam_altcorr_pyramid(_ptr_array(pyr), self.num_levels, _lib.ptr(c), _lib.ptr(ii),
                                                    _lib.ptr(jj), _lib.ptr(out), N, H, W, C, int(self.radius),
                                                    _lib.stream_ptr())
        _lib.check(rc, "altcorr_pyramid")
        return out
