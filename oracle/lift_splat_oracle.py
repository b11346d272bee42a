"""TEST INFRASTRUCTURE ONLY - CPU restatement (numpy) of the reference lift-splat.

This is the parity oracle for the sm_100a kernels.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` leg may import it; the product package never does (and fails loudly
without its CUDA library).

Each function cites the reference lines it restates (paths relative to the
reference root).  Arithmetic that decides an integer (voxel index, keep mask,
rank) is reproduced operation by operation in IEEE float32 so that it is
BIT-EXACT with the reference's torch-CPU execution; the floating-point sums are
done in float64 ("the reference code run in float64") and rounded
once on the fp32 store.

Pinning: ``tests/test_oracle_golden.py`` checks every function against fixtures
frozen from the *unmodified* reference (``tests/golden/make_golden.py`` runs it
through ``oracle/ref_harness.py``) and, when ``/root/reference`` is present,
against the live reference.  The reference ships no tests or golden vectors of
its own for this path, so those fixtures are the pin.

Third-party arithmetic: every op of the path is PyTorch aten (reference pins
torch 1.13.1, environment.yml:86; fixtures were produced with torch 2.11.0 CPU).
The one op that is NOT restated bit-for-bit is ``torch.inverse``
(model/bev_model.py:46,53): on CPU it is MKL LAPACK, on CUDA cuSOLVER/cuBLAS, and
the two already differ in the last bits.  ``camera_transform`` below therefore
defines its own inverse (float64 Gauss-Jordan, rounded once to float32), the
index functions take ``M, t`` as inputs, and bit-exactness of the indices is
pinned with the reference's own ``M, t``.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32
I64_MIN = np.iinfo(np.int64).min


# ----------------------------------------------------------------------------
# a1  tool/geometry.py:40-59   calculate_birds_eye_view_parameters
# ----------------------------------------------------------------------------
def bev_grid_params(x_bound, y_bound, z_bound):
    """res f32[3], start f32[3], dim i64[3].

    The reference builds python-float lists and lets ``torch.tensor`` cast them:
    float64 -> float32 for res/start, truncation toward zero for ``dtype=long``.
    """
    rows = (x_bound, y_bound, z_bound)
    res = np.array([r[2] for r in rows], dtype=np.float64).astype(F32)
    start = np.array([r[0] + r[2] / 2.0 for r in rows], dtype=np.float64).astype(F32)
    dim = np.trunc(np.array([(r[1] - r[0]) / r[2] for r in rows], dtype=np.float64)).astype(np.int64)
    return res, start, dim


def grid_offset(start, res):
    """``bev_start_pos - bev_res / 2.0`` in float32 (model/bev_model.py:85)."""
    return (start.astype(F32) - (res.astype(F32) / F32(2.0))).astype(F32)


# ----------------------------------------------------------------------------
# a3  model/bev_model.py:28-43   create_frustum
# ----------------------------------------------------------------------------
def _torch_linspace_f32(lo, hi, steps):
    """aten linspace (CPU, float): step in float32, filled symmetrically from both
    ends - first half ``lo + step*i``, second half ``hi - step*(steps-1-i)``, each a
    single fused multiply-add (the vectorised AVX2/AVX-512 kernel; emulated here by
    doing the exact product and the add in float64 and rounding once).  The frustum
    is a ``state_dict`` entry, so the kernels always read it as data; this
    restatement only has to agree with torch on the hosts the fixtures come from."""
    lo, hi = F32(lo), F32(hi)
    if steps == 1:
        return np.array([lo], F32)
    step = np.float64(F32((hi - lo) / F32(steps - 1)))
    i = np.arange(steps, dtype=np.float64)
    half = steps // 2
    up = np.float64(lo) + step * i
    down = np.float64(hi) - step * (steps - 1 - i)
    return np.where(i < half, up, down).astype(F32)


def _torch_arange_f32(lo, hi, step):
    """aten arange (float): length ceil((hi-lo)/step) in float64, values
    ``lo + i*step`` accumulated in float64, stored as float32."""
    n = int(np.ceil((float(hi) - float(lo)) / float(step)))
    return (float(lo) + np.arange(n, dtype=np.float64) * float(step)).astype(F32)


def create_frustum(d_bound, final_dim, down_sample):
    """f32[D, h, w, 3] holding (u, v, d) per frustum point."""
    H, W = final_dim
    fh, fw = H // down_sample, W // down_sample
    d = _torch_arange_f32(*d_bound)
    u = _torch_linspace_f32(0, W - 1, fw)
    v = _torch_linspace_f32(0, H - 1, fh)
    fr = np.empty((d.size, fh, fw, 3), F32)
    fr[..., 0] = u[None, None, :]
    fr[..., 1] = v[None, :, None]
    fr[..., 2] = d[:, None, None]
    return fr


# ----------------------------------------------------------------------------
# a4 (first half)  model/bev_model.py:46-47,53   inverses and R . K^-1
# ----------------------------------------------------------------------------
def _gauss_jordan_f64(a):
    """Batched inverse: Gauss-Jordan with partial (row) pivoting in float64.

    a: [..., n, n] float64.  Operation order is part of the contract with the
    CUDA kernel (``ls_camera_transform``): for column k pick the row with the
    largest |a[i,k]|, i>=k (first one on ties), swap, scale the pivot row by a
    true division, then eliminate column k from every other row with
    ``row_i = row_i - f * row_k`` (separate multiply and subtract).
    """
    a = np.array(a, dtype=np.float64)
    n = a.shape[-1]
    lead = a.shape[:-2]
    a = a.reshape(-1, n, n)
    m = np.concatenate([a, np.broadcast_to(np.eye(n), a.shape).copy()], axis=2)
    idx = np.arange(m.shape[0])
    for k in range(n):
        p = k + np.argmax(np.abs(m[:, k:, k]), axis=1)
        rk, rp = m[idx, k].copy(), m[idx, p].copy()
        m[idx, k], m[idx, p] = rp, rk
        piv = m[:, k, k].copy()
        m[:, k, :] = m[:, k, :] / piv[:, None]
        for i in range(n):
            if i == k:
                continue
            f = m[:, i, k].copy()
            m[:, i, :] = m[:, i, :] - f[:, None] * m[:, k, :]
    return m[:, :, n:].reshape(*lead, n, n)


def camera_transform(intrinsics, extrinsics):
    """(M f32[B,N,3,3], t f32[B,N,3]):  E^-1 = inverse(extr), R = E^-1[:3,:3],
    t = E^-1[:3,3], M = R . inverse(K).

    The two inverses are this module's own (see header).  The product R . K^-1 is
    the reference's small-matrix CPU matmul: unfused float32, k ascending,
    starting from +0 (aten baddbmm naive kernel).
    """
    e_inv = _gauss_jordan_f64(np.asarray(extrinsics, F32).astype(np.float64)).astype(F32)
    k_inv = _gauss_jordan_f64(np.asarray(intrinsics, F32).astype(np.float64)).astype(F32)
    rot = e_inv[..., :3, :3]
    t = np.ascontiguousarray(e_inv[..., :3, 3])
    m = np.zeros(rot.shape, F32)
    for k in range(3):
        m = (m + (rot[..., :, k, None] * k_inv[..., None, k, :]).astype(F32)).astype(F32)
    return m, t


# ----------------------------------------------------------------------------
# a4 (second half)  model/bev_model.py:49-55   frustum -> ego coordinates
# ----------------------------------------------------------------------------
def geometry(M, t, frustum):
    """geom f32[B,N,D,h,w,3] = M . (u*d, v*d, d) + t, every op rounded to float32
    in the order torch-CPU executes it: products (u*d),(v*d); row dot product as
    ((0 + m0*px) + m1*py) + m2*pz; then + t."""
    M = np.asarray(M, F32)
    t = np.asarray(t, F32)
    fr = np.asarray(frustum, F32)
    d = fr[..., 2]
    px = (fr[..., 0] * d).astype(F32)[None, None]
    py = (fr[..., 1] * d).astype(F32)[None, None]
    pz = d[None, None]
    b, n = M.shape[:2]
    out = np.empty((b, n) + fr.shape[:3] + (3,), F32)
    for i in range(3):
        mi = M[:, :, i, :, None, None, None]
        acc = (F32(0.0) + (mi[:, :, 0] * px).astype(F32)).astype(F32)
        acc = (acc + (mi[:, :, 1] * py).astype(F32)).astype(F32)
        acc = (acc + (mi[:, :, 2] * pz).astype(F32)).astype(F32)
        out[..., i] = (acc + t[:, :, i, None, None, None]).astype(F32)
    return out


# ----------------------------------------------------------------------------
# a6.1-a6.3  model/bev_model.py:85-97   voxel index, keep mask, rank, sort
# ----------------------------------------------------------------------------
def _to_long(c):
    """float32 -> int64 like ``Tensor.long()`` on x86: truncate toward zero;
    NaN / out-of-range give INT64_MIN (cvttss2si 'integer indefinite')."""
    c = np.asarray(c, F32)
    ok = np.isfinite(c) & (np.abs(c) < F32(9.2e18))
    out = np.full(c.shape, I64_MIN, np.int64)
    out[ok] = np.trunc(c[ok]).astype(np.int64)
    return out


def voxel_index(geom, start, res, dim):
    """Per sample, points flattened in (cam, d, row, col) order.

    Returns vox i64[B,Npts,3], keep bool[B,Npts], rank i64[B,Npts] (-1 where
    dropped).  ``(geom - (start - res/2)) / res`` is a float32 subtract followed by
    an IEEE float32 divide, then truncation (NOT floor).
    """
    geom = np.asarray(geom, F32)
    b = geom.shape[0]
    off = grid_offset(np.asarray(start, F32), np.asarray(res, F32))
    c = ((geom - off).astype(F32) / np.asarray(res, F32)).astype(F32)
    vox = _to_long(c.reshape(b, -1, 3))
    dim = np.asarray(dim, np.int64)
    keep = np.ones(vox.shape[:2], bool)
    for a in range(3):
        keep &= (vox[..., a] >= 0) & (vox[..., a] < dim[a])
    rank = (vox[..., 0] * (dim[1] * dim[2]) + vox[..., 1] * dim[2]) + vox[..., 2]
    rank = np.where(keep, rank, -1)
    return vox, keep, rank


def sorted_ranks(rank_b):
    """``ranks[ranks.argsort()]`` of the kept points of one sample."""
    r = rank_b[rank_b >= 0]
    return np.sort(r, kind="stable")


# ----------------------------------------------------------------------------
# a5  model/bev_model.py:64-71   depth softmax (x) feature outer product
# ----------------------------------------------------------------------------
def softmax_depth(logits, dtype=np.float64):
    z = np.asarray(logits).astype(dtype)
    z = z - z.max(axis=1, keepdims=True)
    e = np.exp(z)
    return e / e.sum(axis=1, keepdims=True)


# ----------------------------------------------------------------------------
# a6.4-a6.5  tool/geometry.py:289-305 + model/bev_model.py:101-105
#            segment sum per voxel, scatter into [B, C, X, Y]
# ----------------------------------------------------------------------------
def splat_forward(feat, logits, rank, dim, cams, acc=np.float64):
    """bev f32[B,C,X,Y], prob f32[B*N,D,h,w].

    Exact formulation of the reference's cumsum trick: every voxel receives the
    sum of prob[d,pix]*feat[:,pix] over the kept points that fall in it; voxels
    nobody hits stay 0.  Accumulated in ``acc`` and rounded once to float32.
    """
    feat = np.asarray(feat)
    bn, C, fh, fw = feat.shape
    D = logits.shape[1]
    B = bn // cams
    X, Y, Z = (int(v) for v in dim)
    prob = softmax_depth(logits, acc)
    bev = np.zeros((B, C, X, Y), F32)
    f = feat.astype(acc).reshape(B, cams, C, fh * fw)
    p = prob.reshape(B, cams, D, fh * fw)
    for b in range(B):
        r = rank[b].reshape(cams, D, fh * fw)
        accum = np.zeros((X * Y * Z, C), acc)
        for n in range(cams):
            for d in range(D):
                rr = r[n, d]
                k = rr >= 0
                if not k.any():
                    continue
                contrib = (p[b, n, d, k][:, None] * f[b, n][:, k].T)
                np.add.at(accum, rr[k], contrib)
        # Z == 1 (config/training.yaml:28): rank == gx*Y + gy
        bev[b] = accum.reshape(X, Y, Z, C)[:, :, 0, :].transpose(2, 0, 1).astype(F32)
    return bev, prob.astype(F32)


def splat_backward(feat, logits, rank, dim, cams, grad_bev, grad_prob_ext=None, acc=np.float64):
    """(grad_feat f32[B*N,C,h,w], grad_logits f32[B*N,D,h,w]).

    Autograd of the reference chain: VoxelsSumming.backward
    (tool/geometry.py:307-317) hands every kept point the gradient of its voxel;
    the outer product (bev_model.py:66) sends it to feat (sum over d) and prob
    (sum over c); softmax backward (bev_model.py:64) folds in whatever gradient
    arrives on the returned ``pred_depth`` as well.
    """
    feat = np.asarray(feat)
    bn, C, fh, fw = feat.shape
    D = logits.shape[1]
    B = bn // cams
    X, Y, Z = (int(v) for v in dim)
    prob = softmax_depth(logits, acc).reshape(B, cams, D, fh * fw)
    f = feat.astype(acc).reshape(B, cams, C, fh * fw)
    gf = np.zeros((B, cams, C, fh * fw), acc)
    gp = np.zeros((B, cams, D, fh * fw), acc)
    for b in range(B):
        g = np.asarray(grad_bev[b]).astype(acc).reshape(C, X * Y)  # Z == 1
        r = rank[b].reshape(cams, D, fh * fw)
        for n in range(cams):
            for d in range(D):
                rr = r[n, d]
                k = rr >= 0
                if not k.any():
                    continue
                gv = g[:, rr[k]]                                   # [C, kept]
                gf[b, n][:, k] += prob[b, n, d, k][None, :] * gv
                gp[b, n, d, k] = (f[b, n][:, k] * gv).sum(axis=0)
    if grad_prob_ext is not None:
        gp = gp + np.asarray(grad_prob_ext).astype(acc).reshape(gp.shape)
    gl = prob * (gp - (prob * gp).sum(axis=2, keepdims=True))
    return (gf.reshape(bn, C, fh, fw).astype(F32), gl.reshape(bn, D, fh, fw).astype(F32))


# ----------------------------------------------------------------------------
# whole path (a8  model/bev_model.py:109-117)
# ----------------------------------------------------------------------------
def lift_splat(feat, logits, intrinsics, extrinsics, frustum, x_bound, y_bound, z_bound,
               M=None, t=None):
    """Convenience: full forward from raw inputs.  Pass ``M, t`` to bypass this
    module's own inverse (e.g. with the reference's values)."""
    res, start, dim = bev_grid_params(x_bound, y_bound, z_bound)
    if M is None:
        M, t = camera_transform(intrinsics, extrinsics)
    geom = geometry(M, t, frustum)
    vox, keep, rank = voxel_index(geom, start, res, dim)
    cams = np.asarray(M).shape[1]
    bev, prob = splat_forward(feat, logits, rank, dim, cams)
    return {"M": M, "t": t, "geom": geom, "vox": vox, "keep": keep, "rank": rank,
            "bev": bev, "prob": prob, "dim": dim}
