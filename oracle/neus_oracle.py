"""CPU oracle (TEST INFRASTRUCTURE — never imported by the product path) for the hash-grid
neural-surface ray marcher: InstantNeuS.forward (src/InstantNeuS.py:295-370) with tiny-cuda-nn
restated.

PARITY UNPINNED for the tiny-cuda-nn pieces: tiny-cuda-nn is a pip-from-git dependency of the
reference with no version pin (README.md:95, call sites src/InstantNeuS.py:62,77,86,192,201) and
is absent from /root/reference and from this image.  What is restated here is its published
algorithm (HashGrid encoding `grid.h`, identity-encoding padding, FullyFusedMLP):
  * level l: scale = exp2(l*log2(per_level_scale))*base_res - 1, res = ceil(scale)+1,
    params_in_level = min(next_multiple(res^3, 8), 2^19); pos = fma(scale, x, 0.5);
    dense index x + y*res + z*res^2 while the running stride <= hashmap_size, otherwise the
    coherent prime hash (x*1) ^ (y*2654435761) ^ (z*805459861); index % hashmap_size;
    features accumulated IN HALF over the 8 corners in idx order (bit d of idx = +1 along dim d):
    result[f] += (half)(weight * (float)value[f]);
  * backward w.r.t. the input (used for the SDF normal): fp32, dL/dy cast to half,
    grad_dim = scale * sum_{4 corners} w_others * (right - left) . dL/dy;
  * Network = identity encoding (input cast to half, padded to a multiple of 16 with 1.0)
    + FullyFusedMLP, weights [out,in] row-major per layer, ReLU hidden, no biases, half
    activations.  tcnn accumulates in half on tensor cores; the restatement (and the CUDA
    kernel) accumulate in fp32 — the rgb tolerance in the tests covers that.
Everything outside tcnn (masking, normalisation, Linear(35,32), NeuS alpha, compositing) is
pinned: tests/golden/neus_*.npz are produced by the reference's own InstantNeuS.forward
(imported from /root/reference, tcnn modules replaced by this restatement) via
tests/golden/make_golden.py.
"""
import math

import numpy as np

F32, F16 = np.float32, np.float16
N_LEVELS, N_FEAT, LOG2_T, BASE_RES = 16, 2, 19, 16
PER_LEVEL_SCALE = 1.447269237440378


def hashgrid_meta():
    # scale_l = exp2f(l * log2f(per_level_scale)) * base - 1 in float, with CORRECTLY ROUNDED log2f / exp2f
    # (double evaluation, one rounding to float) — what glibc's exp2f / log2f give in the library's host code.
    # numpy's float32 exp2 is 1 ulp off at levels 3, 6, 8 and 11, which moved samples that sit within an ulp
    # of a cell face of those levels into the neighbouring cell (a different, piecewise-constant, normal).
    log2b = F32(math.log2(float(F32(PER_LEVEL_SCALE))))
    metas, off = [], 0
    for l in range(N_LEVELS):
        scale = F32(F32(2.0 ** float(F32(F32(l) * log2b))) * F32(BASE_RES) - F32(1.0))
        res = int(math.ceil(float(scale))) + 1
        p = min(res ** 3, 0xFFFFFFFF // 2)
        p = (p + 7) // 8 * 8
        p = min(p, 1 << LOG2_T)
        metas.append(dict(scale=scale, res=res, offset=off, size=p))
        off += p
    return metas, off          # off = total entries (each N_FEAT params)


def _grid_index(meta, x, y, z):
    res, size = np.uint32(meta["res"]), np.uint32(meta["size"])
    x, y, z = x.astype(np.uint32), y.astype(np.uint32), z.astype(np.uint32)
    stride = np.uint64(1)
    index = np.zeros_like(x)
    dims = [x, y, z]
    d = 0
    while d < 3 and stride <= size:
        index = index + dims[d] * np.uint32(stride)
        stride = stride * np.uint64(res)
        d += 1
    if size < stride:
        index = (x * np.uint32(1)) ^ (y * np.uint32(2654435761)) ^ (z * np.uint32(805459861))
    return (index % size).astype(np.int64)


def _pos(meta, x01):
    scale = meta["scale"]
    pos = (x01.astype(np.float64) * np.float64(scale) + 0.5).astype(F32)      # fmaf
    fl = np.floor(pos)
    return fl.astype(np.int64).astype(np.uint32), (pos - fl).astype(F32), scale


def hashgrid_encode(x01, table):
    """x01 [n,3] fp32 in [0,1]; table [total_entries, 2] float16 -> enc [n,32] float16."""
    metas, _ = hashgrid_meta()
    n = x01.shape[0]
    out = np.zeros((n, N_LEVELS * N_FEAT), F16)
    with np.errstate(over="ignore"):
        for l, m in enumerate(metas):
            pg, fr, _ = _pos(m, x01)
            res = np.zeros((n, N_FEAT), F16)
            for idx in range(8):
                w = np.ones(n, F32)
                c = []
                for d in range(3):
                    if (idx >> d) & 1:
                        w = w * fr[:, d]
                        c.append(pg[:, d] + np.uint32(1))
                    else:
                        w = w * (F32(1) - fr[:, d])
                        c.append(pg[:, d])
                val = table[m["offset"] + _grid_index(m, *c)]                 # [n,2] half
                contrib = (w[:, None] * val.astype(F32)).astype(F16)
                res = (res.astype(np.float64) + contrib.astype(np.float64)).astype(F16)
            out[:, 2 * l:2 * l + 2] = res
    return out


def hashgrid_input_grad(x01, table, dLdy):
    """d(sum_k dLdy[k] * enc[k]) / d x01  -> [n,3] fp32 (dLdy rounded to half, like tcnn bwd)."""
    metas, _ = hashgrid_meta()
    n = x01.shape[0]
    g = np.zeros((n, 3), F32)
    gy = np.asarray(dLdy, F32).astype(F16).astype(F32)
    with np.errstate(over="ignore"):
        for l, m in enumerate(metas):
            pg, fr, scale = _pos(m, x01)
            for gd in range(3):
                d1, d2 = (gd + 1) % 3, (gd + 2) % 3
                acc = np.zeros(n, F32)
                for qq in range(4):
                    b1, b2 = qq & 1, (qq >> 1) & 1
                    w = F32(scale) * (fr[:, d1] if b1 else F32(1) - fr[:, d1]) * (fr[:, d2] if b2 else F32(1) - fr[:, d2])
                    cl = [None] * 3
                    cl[d1] = pg[:, d1] + np.uint32(b1)
                    cl[d2] = pg[:, d2] + np.uint32(b2)
                    cl[gd] = pg[:, gd]
                    left = table[m["offset"] + _grid_index(m, *cl)].astype(F32)
                    cl[gd] = pg[:, gd] + np.uint32(1)
                    right = table[m["offset"] + _grid_index(m, *cl)].astype(F32)
                    acc = acc + w * ((right[:, 0] - left[:, 0]) * gy[2 * l] + (right[:, 1] - left[:, 1]) * gy[2 * l + 1])
                g[:, gd] += acc
    return g


def mlp_forward(x, params):
    """x [n,67] fp32 -> [n,3] float16; params float16 flat: [64,80] | [64,64] | [16,64]."""
    n = x.shape[0]
    xin = np.ones((n, 80), F16)
    xin[:, :67] = x.astype(F16)
    p = np.asarray(params, F16)
    W1 = p[:64 * 80].reshape(64, 80).astype(F32)
    W2 = p[64 * 80:64 * 80 + 64 * 64].reshape(64, 64).astype(F32)
    W3 = p[64 * 80 + 64 * 64:].reshape(16, 64).astype(F32)
    h = np.maximum(xin.astype(F32) @ W1.T, 0).astype(F16)
    h = np.maximum(h.astype(F32) @ W2.T, 0).astype(F16)
    o = (h.astype(F32) @ W3.T).astype(F16)
    return o[:, :3]


def _sigmoid(x):
    return (1.0 / (1.0 + np.exp(-x.astype(np.float64)))).astype(F32)


def forward(grid, sdf_w, sdf_b, color_B, mlp_w, bound, rt_bound, variance, scale_factor,
            rays_o, rays_d, z_vals, dists, cos_anneal_ratio=1.0, debug=False):
    """Standalone restatement of InstantNeuS.forward.  grid: float16 [total_params];
    bound / rt_bound: [3,2].  Returns the reference's 9-key dict (numpy)."""
    table = np.asarray(grid, F16).reshape(-1, N_FEAT)
    sdf_w, sdf_b, color_B = np.asarray(sdf_w, F32), np.asarray(sdf_b, F32), np.asarray(color_B, F32)
    bound, rt = np.asarray(bound, F32), np.asarray(rt_bound, F32)
    rays_o, rays_d = np.asarray(rays_o, F32), np.asarray(rays_d, F32)
    z_vals, dists = np.asarray(z_vals, F32), np.asarray(dists, F32)
    R, S = z_vals.shape
    zm = (z_vals + dists / F32(2.0)).astype(F32)
    pts = (rays_o[:, None, :] + rays_d[:, None, :] * zm[:, :, None]).reshape(-1, 3).astype(F32)
    dirs = np.broadcast_to(rays_d[:, None, :], (R, S, 3)).reshape(-1, 3)
    mask = np.all((pts < rt[:, 1]) & (pts > rt[:, 0]), axis=1)
    if mask.sum() < 1:
        mask[:100] = True
    P = pts[mask]
    raw = ((P - bound[:, 0]) / (bound[:, 1] - bound[:, 0]) * F32(2.0) - F32(1.0)).astype(F32)
    xn = np.clip(raw, F32(-1), F32(1))
    x01 = ((xn + F32(1)) / F32(2)).astype(F32)
    enc = hashgrid_encode(x01, table)
    feat_in = np.concatenate([xn, enc.astype(F32)], axis=1)
    out = (feat_in.astype(np.float64) @ sdf_w.astype(np.float64).T + sdf_b).astype(F32)
    sdf_m, feat_m = out[:, 0], out[:, 1:]
    genc = hashgrid_input_grad(x01, table, sdf_w[0, 3:])
    passthru = ((raw >= -1) & (raw <= 1)).astype(F32)
    grad_m = ((sdf_w[0, :3][None] + F32(0.5) * genc) * passthru * (F32(2.0) / (bound[:, 1] - bound[:, 0]))[None]).astype(F32)

    n = R * S
    sdf = np.full(n, 100.0, F32)
    grad = np.zeros((n, 3), F32)
    sdf[mask] = sdf_m
    grad[mask] = grad_m
    inv_s = np.clip(np.exp(np.float64(variance) * scale_factor), 1e-6, 1e6).astype(F32)
    true_cos = (dirs * grad).sum(1).astype(F32)
    iter_cos = -(np.maximum(-true_cos * F32(0.5) + F32(0.5), 0) * F32(1.0 - cos_anneal_ratio)
                 + np.maximum(-true_cos, 0) * F32(cos_anneal_ratio))
    dflat = dists.reshape(-1)
    est_next = sdf + iter_cos * dflat / F32(2.0)
    est_prev = sdf - iter_cos * dflat / F32(2.0)
    prev_cdf = _sigmoid(est_prev * inv_s)
    next_cdf = _sigmoid(est_next * inv_s)
    alpha = np.clip((prev_cdf - next_cdf + F32(1e-5)) / (prev_cdf + F32(1e-5)), 0, 1).astype(F32)

    emb = np.sin((P.astype(np.float64) @ color_B.astype(np.float64)).astype(F32)).astype(F32)
    mlp_in = np.concatenate([emb, grad_m, feat_m], axis=1)
    x = mlp_forward(mlp_in, mlp_w)
    rgb_m = _sigmoid(x.astype(F32)).astype(F16).astype(F32)
    rgb = np.zeros((n, 3), F32)
    rgb[mask] = rgb_m

    alpha = (alpha * mask).reshape(R, S)
    T = np.cumprod(np.concatenate([np.ones((R, 1), F32), 1 - alpha + F32(1e-7)], axis=1), axis=1)[:, :-1]
    w = (alpha * T).astype(F32)
    gr = grad.reshape(R, S, 3)
    mk = mask.reshape(R, S)
    depth = (zm * w).sum(1, keepdims=True)
    gerr = ((np.linalg.norm(gr, axis=2) - 1.0) ** 2 * mk).mean()
    return {
        "color": (rgb.reshape(R, S, 3) * w[:, :, None]).sum(1).astype(F32),
        "depth": depth.astype(F32),
        "depth_variance": (((zm - depth) ** 2) * w).sum(1, keepdims=True).astype(F32),
        "normal": (gr * w[:, :, None] * mk[:, :, None]).sum(1).astype(F32),
        "weight_sum": w.sum(1, keepdims=True).astype(F32),
        "sdf_variance": np.full((R, 1), 1.0 / inv_s, F32),
        "sdf": sdf.reshape(R, S),
        "z_vals": zm,
        "gradient_error": np.array([gerr], F32),
        **({"_alpha": alpha, "_grad": grad, "_weights": w, "_xn": xn, "_mask": mask} if debug else {}),
    }
