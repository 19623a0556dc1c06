"""Bring-up and calibration of the tensor-core chamfer NN kernel (csrc/chamfer_nn_tc.cu, variant 3).

1. dumps the raw tensor-core filter values a_k through psd_debug_tc_filter and compares them with the same quantity in
   float64: a_k = s^2 (|t_k - q|^2 - |q - c|^2) (c = the kernel's frame centre: mean of 8 evenly spaced targets, s = its
   power-of-two scale), reporting
   max |error| / S with S = (|q-c| + max|t-c|)^2 in units of u = 2^-24 -- the constant the exactness margin relies on;
2. checks dist/idx of variant 3 against the C oracle (bit-exact) on a few shapes;
3. reports the fallback rate.

    python tools/tc_calibrate.py            (GPU required)
"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import psd_b200
from oracle import oracle

pkg = psd_b200.load()
L = pkg._lib.lib
dev = torch.device("cuda:0")
vp = lambda t: ctypes.c_void_p(t.data_ptr())


def run_dump(x, y):
    b, n, _ = x.shape
    m = y.shape[1]
    ld = ((max(n, m) + 255) // 256) * 256   # the kernel dumps whole 256-column tiles
    units = b * ((n + 127) // 128 + (m + 127) // 128)
    dump = torch.full((units * 128, ld), float("nan"), device=dev)
    d1 = torch.empty(b, n, device=dev); d2 = torch.empty(b, m, device=dev)
    i1 = torch.empty(b, n, device=dev, dtype=torch.int32); i2 = torch.empty(b, m, device=dev, dtype=torch.int32)
    rc = L.psd_debug_tc_filter(vp(x), vp(y), b, n, m, vp(d1), vp(d2), vp(i1), vp(i2), vp(dump), ld, None)
    torch.cuda.synchronize()
    assert rc == 1, pkg._lib.last_error()
    return dump.cpu().numpy(), d1.cpu().numpy(), d2.cpu().numpy(), i1.cpu().numpy(), i2.cpu().numpy()


def centre(t):
    nt = t.shape[0]
    ns = min(nt, 8)
    step = nt >> 3
    ks = [s if nt < 8 else s * step for s in range(ns)]
    acc = np.zeros(3, np.float32)
    for k in ks:
        acc = (acc + t[k]).astype(np.float32)
    return (acc * np.float32(1.0 / ns)).astype(np.float32)


def calibrate(b, n, m, gen, name):
    x = gen(b, n); y = gen(b, m)
    dump, d1, d2, i1, i2 = run_dump(x.to(dev), y.to(dev))
    xn, yn = x.numpy(), y.numpy()
    worst = 0.0
    worst_budget = 0.0
    U = 2.0 ** -24
    qb1 = (n + 127) // 128
    qb2 = (m + 127) // 128
    for d, (qs, ts, qb, off) in enumerate(((xn, yn, qb1, 0), (yn, xn, qb2, b * qb1))):
        for cl in range(b):
            t = ts[cl]; q = qs[cl]
            c32 = centre(t)
            cmax = np.abs((t - c32).astype(np.float32)).max()
            e = int((np.float32(cmax).view(np.uint32) >> 23) & 0xff)
            e = min(max(e, 27), 227)
            sc = 2.0 ** (126 - e) if cmax > 0 else 1.0      # the kernel's power-of-two scale: sc*cmax in [0.5, 1)
            c = c32.astype(np.float64)
            tc = (t.astype(np.float64) - c) * sc
            qc = (q.astype(np.float64) - c) * sc
            a = (tc * tc).sum(1)[None, :] - 2.0 * qc @ tc.T           # [nq, nt]
            S = (np.sqrt((qc * qc).sum(1)) + np.sqrt((tc * tc).sum(1).max())) ** 2
            # per-pair error budget of the margin analysis (chamfer_nn_tc.cu, resolve): relative 15u (|u| + |t'|)^2 plus the
            # absolute fp16-subnormal part 2^-25 (sqrt(3) (2|u| + |t'|) + 1)
            un, tn = np.sqrt((qc * qc).sum(1))[:, None], np.sqrt((tc * tc).sum(1))[None, :]
            budget = 15 * U * (un + tn) ** 2 + 2.0 ** -25 * (1.74 * (2 * un + tn) + 1.0)
            for blk in range(qb):
                rows = dump[(off + cl * qb + blk) * 128:(off + cl * qb + blk + 1) * 128]
                q0 = blk * 128
                nq = min(128, q.shape[0] - q0)
                got = rows[:nq, :t.shape[0]].astype(np.float64)
                err = np.abs(got - a[q0:q0 + nq]) / S[q0:q0 + nq, None]
                if np.isnan(err).any():
                    print(f"  NaN in dump: dir {d} cloud {cl} block {blk}: {np.isnan(got).sum()} values")
                    return
                worst = max(worst, err.max())
                worst_budget = max(worst_budget, (np.abs(got - a[q0:q0 + nq]) / budget[q0:q0 + nq]).max())
    o1, o2, oi1, oi2 = oracle.chamfer_forward(xn, yn, nthreads=8)
    same = (np.array_equal(d1.view(np.uint32), o1.view(np.uint32)) and np.array_equal(d2.view(np.uint32), o2.view(np.uint32))
            and np.array_equal(i1, oi1) and np.array_equal(i2, oi2))
    print(f"{name:28s} B={b} N={n} M={m}: max filter error = {worst / 2 ** -24:8.2f} u*S   "
          f"max error / per-pair budget = {worst_budget:5.3f}   oracle-identical={same}", flush=True)


def parity(b, n, m, gen, name):
    x = gen(b, n); y = gen(b, m)
    st = np.zeros(2, np.int64)
    L.psd_chamfer_stats(st.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong)), 1)
    old = L.psd_chamfer_nn_variant(3)
    xd, yd = x.to(dev), y.to(dev)
    d1 = torch.empty(b, n, device=dev); d2 = torch.empty(b, m, device=dev)
    i1 = torch.empty(b, n, device=dev, dtype=torch.int32); i2 = torch.empty(b, m, device=dev, dtype=torch.int32)
    assert pkg.chamfer_3D.forward(xd, yd, d1, d2, i1, i2) == 1
    torch.cuda.synchronize()
    L.psd_chamfer_nn_variant(old)
    rc = L.psd_chamfer_stats(st.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong)), 1)
    o1, o2, oi1, oi2 = oracle.chamfer_forward(x.numpy(), y.numpy(), nthreads=16)
    d1, d2, i1, i2 = d1.cpu().numpy(), d2.cpu().numpy(), i1.cpu().numpy(), i2.cpu().numpy()
    same = (np.array_equal(d1.view(np.uint32), o1.view(np.uint32)) and np.array_equal(d2.view(np.uint32), o2.view(np.uint32))
            and np.array_equal(i1, oi1) and np.array_equal(i2, oi2))
    nbad = int((i1 != oi1).sum() + (i2 != oi2).sum())
    print(f"{name:28s} B={b} N={n} M={m}: oracle-identical={same} (idx mismatches {nbad})  stats rc={rc} "
          f"fallback queries={st[1]} ({100.0 * st[1] / (b * (n + m)):.3f} %)", flush=True)


def uniform(b, n):
    return torch.rand(b, n, 3, generator=G)


def clustered(b, n):
    cen = torch.rand(b, 16, 3, generator=G)
    which = torch.randint(0, 16, (b, n), generator=G)
    pts = cen.gather(1, which[..., None].expand(b, n, 3)) + 0.03 * torch.randn(b, n, 3, generator=G)
    return pts.clamp(0, 1).contiguous()


def offset(b, n):
    return torch.rand(b, n, 3, generator=G) + 100.0


def adversarial(kind):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_gpu_tc_hypothesis import adversarial_cloud
    rng = np.random.default_rng(5)
    return lambda b, n: torch.from_numpy(adversarial_cloud(kind, rng, b, n))


G = torch.Generator().manual_seed(11)
if __name__ == "__main__":
    calibrate(1, 128, 128, uniform, "uniform tiny")
    calibrate(2, 300, 520, uniform, "uniform ragged")
    calibrate(2, 2048, 2048, uniform, "uniform config-2 shape")
    calibrate(2, 2048, 2048, clustered, "clustered")
    calibrate(1, 1024, 2048, offset, "offset +100")
    # the adversarial distributions of tests/test_gpu_tc_hypothesis.py: the error must stay inside the per-pair budget (< 1)
    for kind in ("mixed_scales", "outlier_cluster", "split_boundary", "constant_axes", "subnormal_scaled", "planar_lattice", "two_far_clusters"):
        calibrate(1, 700, 1500, adversarial(kind), "adversarial " + kind)
    parity(32, 2048, 2048, uniform, "uniform config 2")
    parity(32, 2048, 2048, clustered, "clustered config 2")
    parity(8, 1000, 257, uniform, "ragged")
    parity(32, 1024, 1024, uniform, "config 1 shape")
