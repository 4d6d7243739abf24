"""`lietorch` — the SE3 subset GO-SLAM's hot path touches, as a small torch-tensor class.

thirdparty/lietorch is an empty submodule in the reference snapshot (pinned e7df8655…); the
algebra is fixed by its in-repo CUDA twins (src/lib/droid_kernels.cu:58-175,877-895) which
se3.cuh restates for the kernels and this file restates for host-side glue.  Call sites
covered: src/depth_video.py:162-164,210; src/geom/projective_ops.py:57,123-124,137,139;
src/motion_filter.py:45.  Data layout [..., 7] = (tx,ty,tz,qx,qy,qz,qw).

This is host glue (pose bookkeeping on a handful of 7-vectors); the per-pixel work of
DepthVideo.reproject goes through the fused kernel goslam_reproject instead.
"""
import torch


def _qmul(a, b):
    ax, ay, az, aw = a.unbind(-1)
    bx, by, bz, bw = b.unbind(-1)
    return torch.stack([
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by + ay * bw + az * bx - ax * bz,
        aw * bz + az * bw + ax * by - ay * bx,
        aw * bw - ax * bx - ay * by - az * bz], dim=-1)


def _qconj(q):
    return torch.cat([-q[..., :3], q[..., 3:]], dim=-1)


def _qrot(q, X):
    """R(q) X, the uv-form of src/lib/droid_kernels.cu:58-68."""
    qv, qw = q[..., :3], q[..., 3:]
    uv = 2.0 * torch.linalg.cross(qv.expand_as(X), X, dim=-1)
    return X + qw * uv + torch.linalg.cross(qv.expand_as(X), uv, dim=-1)


class SE3:
    embedded_dim = 7
    manifold_dim = 6

    def __init__(self, data):
        if isinstance(data, SE3):
            data = data.data
        self.data = data

    # ---- construction / bookkeeping
    @staticmethod
    def Identity(*batch, device=None, dtype=torch.float32, **_):
        d = torch.zeros(*batch, 7, device=device, dtype=dtype)
        d[..., 6] = 1.0
        return SE3(d)

    @property
    def shape(self):
        return self.data.shape[:-1]

    @property
    def device(self):
        return self.data.device

    @property
    def dtype(self):
        return self.data.dtype

    def __getitem__(self, idx):
        if not isinstance(idx, tuple):
            idx = (idx,)
        return SE3(self.data[idx + (slice(None),)])

    def __setitem__(self, idx, item):
        if not isinstance(idx, tuple):
            idx = (idx,)
        self.data[idx + (slice(None),)] = item.data

    def to(self, *a, **kw):
        return SE3(self.data.to(*a, **kw))

    def view(self, *dims):
        return SE3(self.data.view(*dims, 7))

    def vec(self):
        return self.data

    def detach(self):
        return SE3(self.data.detach())

    def translation(self):
        return self.data[..., :3]

    # ---- group
    def inv(self):
        t, q = self.data[..., :3], self.data[..., 3:]
        qi = _qconj(q)
        return SE3(torch.cat([-_qrot(qi, t), qi], dim=-1))

    def mul(self, other):
        t1, q1 = self.data[..., :3], self.data[..., 3:]
        t2, q2 = other.data[..., :3], other.data[..., 3:]
        return SE3(torch.cat([t1 + _qrot(q1, t2), _qmul(q1, q2)], dim=-1))

    def act(self, X):
        t, q = self.data[..., :3], self.data[..., 3:]
        if X.shape[-1] == 3:
            return _qrot(q, X) + t
        # homogeneous (X,Y,Z,d): rotate the first three, add d*t, keep d
        Y = _qrot(q, X[..., :3]) + X[..., 3:] * t
        return torch.cat([Y, X[..., 3:]], dim=-1)

    def __mul__(self, other):
        if isinstance(other, SE3):
            return self.mul(other)
        return self.act(other)

    def adjT(self, X):
        """dual adjoint on 6-covectors (src/lib/droid_kernels.cu:79-94)."""
        t, q = self.data[..., :3], self.data[..., 3:]
        qi = _qconj(q)
        a, b = X[..., :3], X[..., 3:]
        u = torch.linalg.cross(a, t.expand_as(a), dim=-1)     # (t x a) with the sign of the kernel
        ya = _qrot(qi, a)
        yb = _qrot(qi, b) + _qrot(qi, u)
        return torch.cat([ya, yb], dim=-1)

    def matrix(self):
        t, q = self.data[..., :3], self.data[..., 3:]
        eye = torch.eye(3, device=self.data.device, dtype=self.data.dtype)
        R = torch.stack([_qrot(q, eye[i].expand_as(t)) for i in range(3)], dim=-1)
        top = torch.cat([R, t[..., None]], dim=-1)
        bot = torch.zeros_like(top[..., :1, :])
        bot[..., 0, 3] = 1.0
        return torch.cat([top, bot], dim=-2)

    # ---- exp / retraction (src/lib/droid_kernels.cu:110-175,877-895)
    @staticmethod
    def exp(xi):
        tau, phi = xi[..., :3], xi[..., 3:]
        th2 = (phi * phi).sum(-1, keepdim=True)
        th = th2.sqrt()
        small = th2 < 1e-8
        th_safe = torch.where(small, torch.ones_like(th), th)
        imag = torch.where(small, 0.5 - th2 / 48.0 + th2 * th2 / 3840.0, torch.sin(0.5 * th_safe) / th_safe)
        real = torch.where(small, 1.0 - th2 / 8.0 + th2 * th2 / 384.0, torch.cos(0.5 * th_safe))
        q = torch.cat([imag * phi, real], dim=-1)
        big = th > 1e-4
        th2s = torch.where(big, th2, torch.ones_like(th2))
        a = torch.where(big, (1 - torch.cos(th)) / th2s, torch.zeros_like(th))
        b = torch.where(big, (th - torch.sin(th)) / (th_safe * th2s), torch.zeros_like(th))
        c1 = torch.linalg.cross(phi, tau, dim=-1)
        c2 = torch.linalg.cross(phi, c1, dim=-1)
        return SE3(torch.cat([tau + a * c1 + b * c2, q], dim=-1))

    def retr(self, xi):
        return SE3.exp(xi).mul(self)

    def log(self):
        """inverse of `exp` ([..., 6] = (tau, phi)); call site: PoseTrajectoryFiller, dP.log() / dt
        (src/trajectory_filler.py:53).  Rotation vector from the quaternion (atan2 form, sign-safe for
        qw < 0), then tau = V(phi)^-1 t with the same closed form `exp` inverts and the same small-angle
        switch-over (t = tau at or below theta = 1e-4 rad)."""
        t, q = self.data[..., :3], self.data[..., 3:]
        qv, qw = q[..., :3], q[..., 3:]
        n2 = (qv * qv).sum(-1, keepdim=True)
        n = n2.sqrt()
        small = n2 < 1e-12
        n_safe = torch.where(small, torch.ones_like(n), n)
        # 2*atan(n/w)/n, written with atan2 so that w <= 0 (rotations beyond pi) stays continuous
        sgn = torch.where(qw < 0, -torch.ones_like(qw), torch.ones_like(qw))
        two_atan = torch.where(small, 2.0 / qw - (2.0 / 3.0) * n2 / (qw * qw * qw),
                               2.0 * sgn * torch.atan2(n_safe, qw * sgn) / n_safe)
        phi = two_atan * qv
        th2 = (phi * phi).sum(-1, keepdim=True)
        th = th2.sqrt()
        big = th > 1e-4
        th_safe = torch.where(big, th, torch.ones_like(th))
        half = 0.5 * th_safe
        coef = (1.0 - th_safe * torch.cos(half) / (2.0 * torch.sin(half))) / (th_safe * th_safe)
        c1 = torch.linalg.cross(phi, t, dim=-1)
        c2 = torch.linalg.cross(phi, c1, dim=-1)
        # below the threshold `exp` (like expSE3, src/lib/droid_kernels.cu:160) leaves t = tau untouched
        tau = torch.where(big, t - 0.5 * c1 + coef * c2, t)
        return torch.cat([tau, phi], dim=-1)


class Sim3(SE3):
    """placeholder so `isinstance(G, Sim3)` checks in projective_ops.actp resolve (never built)."""
    embedded_dim = 8


def cat(group_list, dim):
    return SE3(torch.cat([g.data for g in group_list], dim=dim))
