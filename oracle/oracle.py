"""TEST INFRASTRUCTURE ONLY -- Python face of the CPU oracle for the point-set-distance hot path.

Three layers, all restating the reference (paths relative to /root/reference):

* ``C``   : ctypes bindings to ``oracle/psd_oracle.c`` (plain C, bit-faithful, fast enough for
            B=32, N=M=2048).  See that file's header for the per-function citations.
* ``np_*``: a vectorised numpy twin with an exactly-rounded fp32 FMA, used to cross-check the C
            code on small cases (it shares no code with it).
* ``torch_*``: the reference's pure-torch CPU chamfer / F-score (loss/loss_.py:66-140, fp64
            ``xx + yy - 2*bmm`` expansion).  This is BASELINE.json configs[0] ("chamfer_python") and
            the CPU arm of ``bench.py --impl reference``.  It is *not* bit-faithful near ties.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / reference legs may
import this module.  The product path (the package under ``3d-pointcloudreconstruction_b200/``)
never does, and raises if its CUDA library is missing.

Parity pin: the reference has no golden vectors; the oracle is pinned against the
reference's own CUDA extensions built unmodified into ``oracle/_ref`` (``oracle/build_ref.py``) and
run on the GPU box -- ``tests/test_reference_parity.py`` and the vectors under ``tests/golden/``.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libpsd_oracle.so")
_SRC = os.path.join(_HERE, "psd_oracle.c")


def build_c(force: bool = False) -> str:
    """gcc the C restatement in place (-ffp-contract=off: only the explicit fmaf() calls fuse)."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        subprocess.check_call(
            ["gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared", "-o", _SO, _SRC, "-lm"]
        )
    return _SO


_lib = None


def _c():
    global _lib
    if _lib is None:
        build_c()
        lib = ctypes.CDLL(_SO)
        f32p = ctypes.POINTER(ctypes.c_float)
        i32p = ctypes.POINTER(ctypes.c_int)
        i64p = ctypes.POINTER(ctypes.c_longlong)
        ci = ctypes.c_int
        lib.oracle_chamfer_forward.argtypes = [ci, ci, ci, f32p, f32p, f32p, f32p, i32p, i32p]
        lib.oracle_chamfer_forward_mt.argtypes = [ci, ci, ci, f32p, f32p, f32p, f32p, i32p, i32p, ci]
        lib.oracle_chamfer_backward.argtypes = [ci, ci, ci, f32p, f32p, f32p, f32p, f32p, f32p, i32p, i32p]
        lib.oracle_emd_forward.argtypes = [ci, ci, ci, f32p, f32p, f32p, i32p, f32p, i32p, i32p, f32p, f32p, i32p,
                                           i32p, ctypes.c_float, ci, i64p, i32p, ci]
        lib.oracle_emd_backward.argtypes = [ci, ci, f32p, f32p, f32p, f32p, i32p]
        lib.oracle_fscore_counts.argtypes = [ci, ci, ci, f32p, f32p, ctypes.c_float, i32p, i32p]
        for fn in ("oracle_chamfer_forward", "oracle_chamfer_forward_mt", "oracle_chamfer_backward",
                   "oracle_emd_forward", "oracle_emd_backward"):
            getattr(lib, fn).restype = ci
        lib.oracle_fscore_counts.restype = None
        _lib = lib
    return _lib


def _f(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _i(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))


def _as_f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a


# ---------------------------------------------------------------------------------------------
# C oracle
# ---------------------------------------------------------------------------------------------
def chamfer_forward(xyz1, xyz2, nthreads: int = 1):
    """chamfer_cuda_forward (chamfer3D.cu:136-154): returns dist1[B,N], dist2[B,M], idx1, idx2 (int32)."""
    xyz1, xyz2 = _as_f32(xyz1), _as_f32(xyz2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    dist1 = np.zeros((b, n), np.float32)
    dist2 = np.zeros((b, m), np.float32)
    idx1 = np.zeros((b, n), np.int32)
    idx2 = np.zeros((b, m), np.int32)
    if nthreads > 1:
        rc = _c().oracle_chamfer_forward_mt(b, n, m, _f(xyz1), _f(xyz2), _f(dist1), _f(dist2), _i(idx1), _i(idx2), nthreads)
    else:
        rc = _c().oracle_chamfer_forward(b, n, m, _f(xyz1), _f(xyz2), _f(dist1), _f(dist2), _i(idx1), _i(idx2))
    assert rc == 1
    return dist1, dist2, idx1, idx2


def chamfer_backward(xyz1, xyz2, graddist1, graddist2, idx1, idx2):
    """chamfer_cuda_backward (chamfer3D.cu:176-195) onto zero gradients (dist_chamfer_3D.py:62-65)."""
    xyz1, xyz2 = _as_f32(xyz1), _as_f32(xyz2)
    g1, g2 = _as_f32(graddist1), _as_f32(graddist2)
    idx1 = np.ascontiguousarray(idx1, np.int32)
    idx2 = np.ascontiguousarray(idx2, np.int32)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    gx1 = np.zeros_like(xyz1)
    gx2 = np.zeros_like(xyz2)
    rc = _c().oracle_chamfer_backward(b, n, m, _f(xyz1), _f(xyz2), _f(gx1), _f(gx2), _f(g1), _f(g2), _i(idx1), _i(idx2))
    assert rc == 1
    return gx1, gx2


def emd_forward(xyz1, xyz2, eps: float, iters: int, nthreads: int = 1, want_stats: bool = False, full_state: bool = False):
    """emdFunction.forward + emd_cuda_forward (emd_module.py:31-76, emd_cuda.cu:228-282).

    Returns (dist[B,n], assignment[B,n]) and, if asked, a stats dict (sum of bidder counts, number of
    +-1e-6 multi-winner events, number of best==better bids, per-iteration bidder counts) and/or the
    scratch state (price, assignment_inv, ...)."""
    xyz1, xyz2 = _as_f32(xyz1), _as_f32(xyz2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    # the Python asserts of emd_module.py:36-39
    assert n == m
    assert xyz1.shape[0] == xyz2.shape[0]
    assert n % 1024 == 0
    assert b <= 512
    dist = np.zeros((b, n), np.float32)
    assignment = np.full((b, n), -1, np.int32)
    assignment_inv = np.full((b, n), -1, np.int32)
    price = np.zeros((b, n), np.float32)
    bid = np.zeros((b, n), np.int32)
    bid_increments = np.zeros((b, n), np.float32)
    max_increments = np.zeros((b, n), np.float32)
    unass_idx = np.zeros((b * n,), np.int32)
    max_idx = np.zeros((b * n,), np.int32)
    stats = np.zeros(3, np.int64)
    u_hist = np.zeros((b, max(iters, 1)), np.int32)
    rc = _c().oracle_emd_forward(
        b, n, m, _f(xyz1), _f(xyz2), _f(dist), _i(assignment), _f(price), _i(assignment_inv), _i(bid),
        _f(bid_increments), _f(max_increments), _i(unass_idx), _i(max_idx), ctypes.c_float(eps), int(iters),
        stats.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong)) if want_stats else None,
        _i(u_hist) if want_stats else None, int(nthreads))
    assert rc == 1, rc
    out = [dist, assignment]
    if want_stats:
        out.append({"sum_u": int(stats[0]), "multi_winner": int(stats[1]), "best_eq_better": int(stats[2]),
                    "u_hist": u_hist[:, :iters]})
    if full_state:
        out.append({"price": price, "assignment_inv": assignment_inv, "bid": bid, "bid_increments": bid_increments,
                    "max_increments": max_increments, "max_idx": max_idx.reshape(b, n)})
    return tuple(out)


def emd_backward(xyz1, xyz2, graddist, assignment):
    """emd_cuda_backward (emd_cuda.cu:284-316): gradient w.r.t. xyz1 only (xyz2's is zeros, emd_module.py:84-87)."""
    xyz1, xyz2 = _as_f32(xyz1), _as_f32(xyz2)
    g = _as_f32(graddist)
    a = np.ascontiguousarray(assignment, np.int32)
    b, n, _ = xyz1.shape
    gx = np.zeros_like(xyz1)
    rc = _c().oracle_emd_backward(b, n, _f(xyz1), _f(xyz2), _f(gx), _f(g), _i(a))
    assert rc == 1
    return gx


def fscore_counts(dist1, dist2, thr: float = 1e-4):
    dist1, dist2 = _as_f32(dist1), _as_f32(dist2)
    b, n = dist1.shape
    m = dist2.shape[1]
    c1 = np.zeros(b, np.int32)
    c2 = np.zeros(b, np.int32)
    _c().oracle_fscore_counts(b, n, m, _f(dist1), _f(dist2), ctypes.c_float(thr), _i(c1), _i(c2))
    return c1, c2


def fscore_from_counts(cnt1, cnt2, n: int, m: int):
    """loss/loss_.py:122-140 on thresholded counts.  The reference's precision_1 is taken over
    min(P, 1) = the y->x distances (the CUDA op's dist2) and precision_2 over min(P, 2) = x->y
    (dist1) (loss_.py:106-109,131-133); returns (fscore, mean precision_1, mean precision_2)."""
    p1 = (np.asarray(cnt2, np.float32) / np.float32(m)).astype(np.float32)
    p2 = (np.asarray(cnt1, np.float32) / np.float32(n)).astype(np.float32)
    with np.errstate(invalid="ignore", divide="ignore"):
        f = (np.float32(2) * p1 * p2 / (p1 + p2)).astype(np.float32)
    f[np.isnan(f)] = 0
    return np.float32(f.mean()), np.float32(p1.mean()), np.float32(p2.mean())


# ---------------------------------------------------------------------------------------------
# numpy twin (independent restatement; exact fp32 FMA through round-to-odd in fp64)
# ---------------------------------------------------------------------------------------------
def fma32(a, b, c):
    """Correctly rounded fp32 fma(a, b, c) for fp32 arrays: the product is exact in fp64 (48 bits), the
    fp64 sum is turned into round-to-odd with TwoSum, and the final fp64->fp32 rounding is then exact."""
    a = np.asarray(a, np.float32).astype(np.float64)
    b = np.asarray(b, np.float32).astype(np.float64)
    c = np.asarray(c, np.float32).astype(np.float64)
    p = a * b
    s = p + c
    bb = s - p
    err = (p - (s - bb)) + (c - bb)  # TwoSum: p + c == s + err exactly
    fin = np.isfinite(s) & (err != 0)
    toward = np.where(err > 0, np.inf, -np.inf)
    nxt = np.nextafter(s, toward)
    s_odd = (s.view(np.uint64) & np.uint64(1)) == 1
    ro = np.where(fin & ~s_odd, nxt, s)  # of the two neighbours bracketing the exact value keep the odd one
    return ro.astype(np.float32)


def np_sqdist(t, q):
    """d = fma(dz,dz, fma(dx,dx, rn(dy*dy))), differences target - query (chamfer3D.cu:32-35)."""
    dx = (t[..., 0] - q[..., 0]).astype(np.float32)
    dy = (t[..., 1] - q[..., 1]).astype(np.float32)
    dz = (t[..., 2] - q[..., 2]).astype(np.float32)
    yy = (dy * dy).astype(np.float32)
    return fma32(dz, dz, fma32(dx, dx, yy))


def np_chamfer_forward(xyz1, xyz2):
    """Finite inputs only (np.argmin's first-occurrence rule == strict '<' scan == lowest index)."""
    xyz1, xyz2 = _as_f32(xyz1), _as_f32(xyz2)
    d = np_sqdist(xyz2[:, None, :, :], xyz1[:, :, None, :])  # [B,N,M]
    idx1 = d.argmin(2).astype(np.int32)
    idx2 = d.argmin(1).astype(np.int32)
    return d.min(2), d.min(1), idx1, idx2


def np_emd_forward(xyz1, xyz2, eps, iters):
    """Vectorised auction for small cases; lowest-index rules as in psd_oracle.c."""
    xyz1, xyz2 = _as_f32(xyz1), _as_f32(xyz2)
    b, n, _ = xyz1.shape
    eps = np.float32(eps)
    dist = np.zeros((b, n), np.float32)
    assignment = np.full((b, n), -1, np.int32)
    for i in range(b):
        ass = assignment[i]
        ass_inv = np.full(n, -1, np.int32)
        price = np.zeros(n, np.float32)
        max_inc = np.zeros(n, np.float32)
        max_idx = np.zeros(n, np.int32)
        for it in range(iters):
            last = it == iters - 1
            un = np.nonzero(ass == -1)[0]
            if un.size == 0:
                continue
            s = np_sqdist(xyz2[i][None, :, :], xyz1[i][un][:, None, :])  # [u,n]
            v = (3.0 - np.sqrt(s).astype(np.float64) - price.astype(np.float64)[None, :]).astype(np.float32)
            best_i = v.argmax(1)
            best = v[np.arange(un.size), best_i]
            v2 = v.copy()
            v2[np.arange(un.size), best_i] = -np.inf
            better = np.maximum(v2.max(1), np.float32(-1e9)).astype(np.float32)
            inc = ((best - better).astype(np.float32) + eps).astype(np.float32)
            np.maximum.at(max_inc, best_i, inc)
            mi = max_inc[best_i].astype(np.float64)
            match = (inc.astype(np.float64) - 1e-6 <= mi) & (mi <= inc.astype(np.float64) + 1e-6)
            for a in range(un.size - 1, -1, -1):
                if match[a]:
                    max_idx[best_i[a]] = un[a]
            for a in range(un.size):
                j, o = un[a], best_i[a]
                if last or max_idx[o] == j:
                    if not last and ass_inv[o] != -1:
                        ass[ass_inv[o]] = -1
                    ass_inv[o] = j
                    ass[j] = o
                    price[o] = np.float32(price[o] + inc[a])
                    max_inc[o] = np.float32(-1e9)
        sel = xyz2[i][ass]
        dist[i] = np_sqdist(xyz1[i], sel)  # CalcDist: xyz1 - xyz2 (emd_cuda.cu:221-224); squares are sign-symmetric
    return dist, assignment


# ---------------------------------------------------------------------------------------------
# the reference's pure-torch CPU chamfer ("chamfer_python", BASELINE.json configs[0])
# ---------------------------------------------------------------------------------------------
def torch_batched_pairwise_dist(a, b):
    """loss/loss_.py:66-77: fp64 xx + yy - 2*bmm(x, y^T)."""
    import torch

    x, y = a.double(), b.double()
    bs, nx, _ = x.size()
    _, ny, _ = y.size()
    xx = torch.pow(x, 2).sum(2)
    yy = torch.pow(y, 2).sum(2)
    zz = torch.bmm(x, y.transpose(2, 1))
    rx = xx.unsqueeze(1).expand(bs, ny, nx)
    ry = yy.unsqueeze(1).expand(bs, nx, ny)
    return rx.transpose(2, 1) + ry - 2 * zz


def torch_dist_chamfer(a, b):
    """loss/loss_.py:79-91 (distChamfer): dist1, dist2 (fp32), idx1, idx2 (int32)."""
    import torch

    P = torch_batched_pairwise_dist(a, b)
    m2 = torch.min(P, 2)
    m1 = torch.min(P, 1)
    return m2[0].float(), m1[0].float(), m2[1].int(), m1[1].int()


def torch_fscore(X, Y, threshold=0.0001):
    """loss/loss_.py:93-140 (batch_NN_loss + fscore): returns (fscore, mean precision_1, mean precision_2)."""
    import torch

    P = torch_batched_pairwise_dist(X, Y)
    dist1, _ = torch.min(P, 1)
    dist2, _ = torch.min(P, 2)
    p1 = torch.mean((dist1 < threshold).float(), dim=1)
    p2 = torch.mean((dist2 < threshold).float(), dim=1)
    f = 2 * p1 * p2 / (p1 + p2)
    f[torch.isnan(f)] = 0
    return torch.mean(f), torch.mean(p1), torch.mean(p2)


# ---------------------------------------------------------------------------------------------------------------------
# projection-loss min-distance terms (loss/proj_loss.py:21-40, grid_dist :46-54): dense numpy restatement in fp32
# ---------------------------------------------------------------------------------------------------------------------
def grid_dist(grid_h, grid_w):
    """proj_loss.py:46-54: cdist between all grid points, float64, [H,W,H,W]."""
    hh = np.arange(grid_h, dtype=np.float64)
    ww = np.arange(grid_w, dtype=np.float64)
    dh = hh[:, None, None, None] - hh[None, None, :, None]
    dw = ww[None, :, None, None] - ww[None, None, None, :]
    return np.sqrt(dh * dh + dw * dw)


def proj_min_dist(pred, gt, dist_mat, mode="as_written"):
    """min_dist, min_dist_inv for pred, gt [B,H,W] fp32 and dist_mat [H,W,H,W] fp32 (already incremented).
    as_written: proj_loss.py:25-41 literally (weights broadcast along the FIRST pixel pair);
    intended:   weights on the target pixel (h',w').  Products in fp32, left to right, like torch."""
    pred = np.ascontiguousarray(pred, np.float32); gt = np.ascontiguousarray(gt, np.float32)
    d = np.ascontiguousarray(dist_mat, np.float32)
    one, big = np.float32(1.0), np.float32(1e6)
    gt_th = (gt + ((one - gt) * big) * one).astype(np.float32)
    pred_mask = (pred + ((one - pred) * big) * one).astype(np.float32)
    if mode == "as_written":
        a = ((gt_th[:, :, :, None, None] * d[None]).astype(np.float32) * pred[:, :, :, None, None]).astype(np.float32)
        c = ((gt[:, :, :, None, None] * d[None]).astype(np.float32) * pred_mask[:, :, :, None, None]).astype(np.float32)
    else:
        a = ((gt_th[:, None, None, :, :] * d[None]).astype(np.float32) * pred[:, :, :, None, None]).astype(np.float32)
        c = ((pred_mask[:, None, None, :, :] * d[None]).astype(np.float32) * gt[:, :, :, None, None]).astype(np.float32)
    return a.min(axis=(3, 4)), c.min(axis=(3, 4))      # np.min propagates NaN like torch.min


# ---------------------------------------------------------------------------------------------
# ICP alignment (utils/icp.py) -- numpy restatement; the KD-tree query is replaced by the fp64 brute force it is
# equivalent to (reduced distance = sum of squared differences in axis order, first index on ties).
# Pinned by tests/golden/icp_*_ref.npz, produced by the reference's own icp() (sklearn KD-tree) in
# tests/golden/make_golden_icp.py.
# ---------------------------------------------------------------------------------------------
def icp_best_fit_transform(A, B):
    """utils/icp.py:4-46.  Like the reference, the inputs keep their dtype: icp()'s final call passes the caller's float32 A
    next to the float64 moved source, so centroid_A and AA are float32 arithmetic there (np.mean over axis 0 of a
    C-contiguous [N,3] array accumulates row after row in the array's dtype)."""
    A = np.asarray(A)
    B = np.asarray(B)
    assert A.shape == B.shape
    m = A.shape[1]
    centroid_A = np.mean(A, axis=0)
    centroid_B = np.mean(B, axis=0)
    AA = A - centroid_A
    BB = B - centroid_B
    H = np.dot(AA.T, BB)                       # :27
    U, S, Vt = np.linalg.svd(H)                # :28
    R = np.dot(Vt.T, U.T)                      # :29
    if np.linalg.det(R) < 0:                   # :32-35 reflection case
        Vt[m - 1, :] *= -1
        R = np.dot(Vt.T, U.T)
    t = centroid_B.T - np.dot(R, centroid_A.T)  # :38
    T = np.identity(m + 1)
    T[:m, :m] = R
    T[:m, m] = t
    return T, R, t


def icp_nearest_neighbor(src, dst):
    """utils/icp.py:49-65 (NearestNeighbors(n_neighbors=1).fit(dst).kneighbors(src)), as an fp64 brute force."""
    src = np.asarray(src, dtype=np.float64)
    dst = np.asarray(dst, dtype=np.float64)
    dist = np.empty(src.shape[0])
    idx = np.empty(src.shape[0], dtype=np.int64)
    for s in range(0, src.shape[0], 512):
        d = dst[None, :, :] - src[s:s + 512, None, :]
        d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
        k = np.argmin(d2, axis=1)
        idx[s:s + 512] = k
        dist[s:s + 512] = np.sqrt(d2[np.arange(d2.shape[0]), k])
    return dist, idx


def icp(A, B, init_pose=None, max_iterations=20, tolerance=0.001):
    """utils/icp.py:68-118."""
    A = np.asarray(A)
    B = np.asarray(B)
    assert A.shape == B.shape
    m = A.shape[1]
    src = np.ones((m + 1, A.shape[0]))
    dst = np.ones((m + 1, B.shape[0]))
    src[:m, :] = np.copy(A.T)
    dst[:m, :] = np.copy(B.T)
    if init_pose is not None:
        src = np.dot(init_pose, src)
    prev_error = 0
    for i in range(max_iterations):
        distances, indices = icp_nearest_neighbor(src[:m, :].T, dst[:m, :].T)      # :98
        T, _, _ = icp_best_fit_transform(src[:m, :].T, dst[:m, indices].T)         # :101
        src = np.dot(T, src)                                                        # :104
        mean_error = np.mean(distances)                                             # :107
        if np.abs(prev_error - mean_error) < tolerance:
            break
        prev_error = mean_error
    T, _, _ = icp_best_fit_transform(A, src[:m, :].T)                               # :113
    return T, distances, i


# ---------------------------------------------------------------------------------------------
# Farthest point sampling (utils/utils.py:335-360) and the Gaussian splat cont_proj (utils/projection.py:4-67, 95-106):
# numpy restatements in float32, pinned by tests/golden/fps_*_ref.npz and splat_*_ref.npz which the reference's own torch
# code produced on the CPU (tests/golden/make_golden_fps_splat.py).
# ---------------------------------------------------------------------------------------------
def farthest_point_sample(xyz, npoint, RAN=True):
    """utils/utils.py:335-360.  torch.randint(0, 1) is always 0 and torch.randint(1, 2) always 1, so the first centroid is
    index 0 (RAN) or 1 (not RAN); torch.sum over the 3 coordinates adds in axis order; torch.max returns the first maximum."""
    xyz = np.asarray(xyz, dtype=np.float32)
    B, N, C = xyz.shape
    centroids = np.zeros((B, npoint), dtype=np.int64)
    distance = np.full((B, N), 1e10, dtype=np.float32)
    farthest = np.zeros(B, dtype=np.int64) if RAN else np.ones(B, dtype=np.int64)
    rows = np.arange(B)
    for i in range(npoint):
        centroids[:, i] = farthest
        centroid = xyz[rows, farthest, :][:, None, :]
        d = xyz - centroid
        sq = d * d
        dist = (sq[..., 0] + sq[..., 1]) + sq[..., 2]
        mask = dist < distance
        distance[mask] = dist[mask]
        farthest = np.argmax(distance, axis=-1)
    return centroids


def cont_proj(pcl, grid_h, grid_w, sigma_sq=0.5):
    """utils/projection.py:4-67.  float32 rounding sequence of the torch expressions; exp is the correctly rounded float32
    exponential (torch's CPU exp and CUDA's expf are both within 1-2 ulp of it); the sum runs over the points in order."""
    pcl = np.asarray(pcl, dtype=np.float32)
    f = np.float32
    x = ((pcl[..., 0] + f(1)) * f(grid_h)) / f(2)
    y = ((pcl[..., 1] + f(1)) * f(grid_w)) / f(2)
    two_s = f(2.0 * sigma_sq)
    hh = np.arange(grid_h, dtype=np.float32)
    ww = np.arange(grid_w, dtype=np.float32)

    def kern(d):
        a = (-(d * d)) / two_s                       # float32
        return np.exp(a.astype(np.float64)).astype(np.float32)
    ex = kern(x[:, :, None] - hh[None, None, :])     # [B,N,H]
    ey = kern(y[:, :, None] - ww[None, None, :])     # [B,N,W]
    out = np.zeros((pcl.shape[0], grid_h, grid_w), dtype=np.float32)
    for p in range(pcl.shape[1]):
        out += ex[:, p, :, None] * ey[:, p, None, :]
    return out
