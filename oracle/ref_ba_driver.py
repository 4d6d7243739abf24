"""ba_cuda (src/lib/droid_kernels.cu:1314-1434) re-driven on top of the REFERENCE'S OWN CUDA
kernels (oracle/_ref, see build_ref.py).  TEST INFRASTRUCTURE, runs on the GPU box only.

Only the Eigen-dependent host code is restated (Eigen is absent from the snapshot):
  SparseBlock::update_lhs/update_rhs (:1131-1173) -> dense float64 scatter-add
  schur_block (:1222-1311)                        -> same graph / index lists, reference
                                                     EEt6x6 / Ev6x1 kernels
  SparseBlock::solve (:1192-1213)                 -> torch.linalg.cholesky_ex in float64
Everything per-pixel runs in the reference kernels: projective_transform_kernel, accum_kernel
(through accum_cuda), EEt6x6, Ev6x1, EvT6x1, pose_retr, disp_retr.
"""
import torch


def _dense_add_blocks(A, blocks, ri, ci, P):
    for n in range(blocks.shape[0]):
        i, j = int(ri[n]), int(ci[n])
        if i >= 0 and j >= 0 and i < P and j < P:
            A[6 * i:6 * i + 6, 6 * j:6 * j + 6] += blocks[n]


def _dense_add_vecs(b, vecs, ri, P):
    for n in range(vecs.shape[0]):
        i = int(ri[n])
        if 0 <= i < P:
            b[6 * i:6 * i + 6] += vecs[n]


def ba(ref, poses, disps, intrinsics, disps_sens, targets, weights, eta, ii, jj, t0, t1, iterations,
       lm, ep, motion_only):
    """in place on poses / disps (CUDA tensors), like the reference.  Returns (dx, dz, status)."""
    dev = poses.device
    num = ii.shape[0]
    ht, wd = disps.shape[1], disps.shape[2]
    P = t1 - t0
    ts = torch.arange(t0, t1, device=dev)
    ii_exp = torch.cat([ts, ii])
    jj_exp = torch.cat([ts, jj])
    kx, kk_exp = torch.unique(ii_exp, sorted=True, return_inverse=True)
    dx = dz = None
    status = []
    for _ in range(iterations):
        Hs, vs, Eii, Eij, Cii, wi = ref.linearize(poses, disps, intrinsics, targets, weights, ii, jj)
        A = torch.zeros(6 * P, 6 * P, dtype=torch.float64)
        b = torch.zeros(6 * P, dtype=torch.float64)
        ri = (torch.cat([ii, ii, jj, jj]) - t0).cpu()
        ci = (torch.cat([ii, jj, ii, jj]) - t0).cpu()
        _dense_add_blocks(A, Hs.reshape(-1, 6, 6).cpu().double(), ri, ci, P)
        _dense_add_vecs(b, vs.reshape(-1, 6).cpu().double(), (torch.cat([ii, jj]) - t0).cpu(), P)
        if not motion_only:
            alpha = 0.05
            m = (disps_sens[kx] > 0).float().view(-1, ht * wd)
            C = ref.accum(Cii, ii, kx) + m * alpha + (1 - m) * eta.view(-1, ht * wd)
            w = ref.accum(wi, ii, kx) - m * alpha * (disps[kx] - disps_sens[kx]).view(-1, ht * wd)
            Q = 1.0 / C
            Ei = ref.accum(Eii.view(num, 6 * ht * wd), ii, ts).view(P, 6, ht * wd)
            E = torch.cat([Ei, Eij], 0)
            # schur_block
            jj_c, kk_c = jj_exp.cpu().tolist(), kk_exp.cpu().tolist()
            graph = [[] for _ in range(P)]
            index = [[] for _ in range(P)]
            for n in range(len(jj_c)):
                j, k = jj_c[n], kk_c[n]
                if t0 <= j < t1:            # the reference tests j <= t1; j == t1 would index out of range
                    graph[j - t0].append(k)
                    index[j - t0].append(n)
            ii_list, jj_list, idx = [], [], []
            for i in range(P):
                for j in range(P):
                    for k in range(len(graph[i])):
                        for l in range(len(graph[j])):
                            if graph[i][k] == graph[j][l]:
                                ii_list.append(i)
                                jj_list.append(j)
                                idx += [index[i][k], index[j][l], graph[i][k]]
            ix_cuda = torch.tensor(idx, dtype=torch.long, device=dev).view(-1, 3)
            jx_cuda = kk_exp.view(-1, 1).contiguous()
            S = ref.EEt6x6(E, Q, ix_cuda)
            v = ref.Ev6x1(E, Q, w, jx_cuda)
            SA = torch.zeros_like(A)
            sb = torch.zeros_like(b)
            _dense_add_blocks(SA, S.cpu().double(), ii_list, jj_list, P)
            _dense_add_vecs(sb, v.cpu().double(), (jj_exp - t0).cpu(), P)
            A, b = A - SA, b - sb
        L = A.clone()
        d = torch.diagonal(L)
        d += float(torch.tensor(ep, dtype=torch.float32)) + float(torch.tensor(lm, dtype=torch.float32)) * d.clone()
        c, info = torch.linalg.cholesky_ex(L)
        if int(info) == 0:
            x = torch.cholesky_solve(b[:, None], c)[:, 0]
            dx = x.view(P, 6).float().to(dev)
            status.append(0)
        else:
            dx = torch.zeros(P, 6, device=dev)
            status.append(1)
        if motion_only:
            ref.pose_retr(poses, dx, t0, t1)
        else:
            ixs = (jj_exp - t0).contiguous()
            dw = ref.EvT6x1(E, dx, ixs)
            dz = Q * (w - ref.accum(dw, ii_exp, kx))
            ref.pose_retr(poses, dx, t0, t1)
            ref.disp_retr(disps, dz.contiguous(), kx)
    return dx, dz, status, kx
