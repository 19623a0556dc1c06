"""CPU model of the tensor-core filter's decision logic (csrc/chamfer_nn_tc.cu): operands, split-fp16 filter values, chunk
minima and the margin test, in numpy -- to check the exactness argument on adversarial clouds WITHOUT a GPU.

For every query it reports whether the kernel would trust the filter (margin test passed) and, if so, whether the reference's
argmin (exact formula, lowest index) really lies in the chunk the filter picked.  A trusted query whose argmin lies elsewhere is a
VIOLATION: the kernel would return a wrong index.  The model evaluates the K=16 dot product in float64 and rounds once to fp32
(the tensor core's own accumulation error, <= 4 u S by tools/tc_calibrate.py, is covered by the margin's relative part), so a
violation here is a violation of the operand-split analysis itself.

    python tools/tc_filter_model.py [examples]      # sweeps the distributions of tests/test_gpu_tc_hypothesis.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
f32 = np.float32
# constants of resolve() in chamfer_nn_tc.cu -- keep in step with the kernel
RHO_SLACK, MARGIN_REL, MARGIN_ABS = f32(2.0e-6), f32(2.5e-6), f32(1.5e-7)


def split_h(x):
    with np.errstate(over="ignore", invalid="ignore"):
        h = x.astype(np.float16)
        l = (x - h.astype(f32)).astype(f32).astype(np.float16)
    return h.astype(f32), l.astype(f32)


def model(q, t, abs_term=True):
    """q [nq,3], t [nt,3] float32 (nt <= 2048).  Returns (trusted mask, violation mask)."""
    nt = len(t)
    ns, step = min(nt, 8), nt >> 3
    c = np.zeros(3, f32)
    for s in range(ns):
        c = (c + t[s if nt < 8 else s * step]).astype(f32)
    c = (c * f32(1.0 / ns)).astype(f32)
    cmax = np.abs((t - c).astype(f32)).max()
    e = int((f32(cmax).view(np.uint32) >> 23) & 0xff)
    bad = e < 67 or not np.isfinite(cmax)
    e = min(max(e, 27), 227)
    cs = np.uint32((253 - e) << 23).view(f32) if cmax > 0 else f32(1)
    tp = ((t - c).astype(f32) * cs).astype(f32)
    w = (tp[:, 2] * tp[:, 2] + (tp[:, 0] * tp[:, 0] + (tp[:, 1] * tp[:, 1]).astype(f32)).astype(f32)).astype(f32)
    bad = bad or not (w < 4).all()
    wmax = w.max()
    th, tl = split_h(tp)
    w1 = w.astype(np.float16).astype(f32)
    wr = (w - w1).astype(f32)
    w2 = wr.astype(np.float16).astype(f32)
    w3 = (wr - w2).astype(f32).astype(np.float16).astype(f32)
    u = ((q - c).astype(f32) * cs).astype(f32)
    qp = (u * f32(-2.0)).astype(f32)
    qh, ql = split_h(qp)
    with np.errstate(over="ignore", invalid="ignore"):
        a = (qh.astype(np.float64) @ (th + tl).astype(np.float64).T + ql.astype(np.float64) @ th.astype(np.float64).T
             + (w1.astype(np.float64) + w2 + w3)[None, :]).astype(f32)
    nch = (nt + 31) // 32
    pad = nch * 32 - nt
    if pad:
        a = np.concatenate([a, np.full((len(q), pad), f32(32768.0))], 1)
    cm = a.reshape(len(q), nch, 32).min(2)
    bc = cm.argmin(1)
    b1 = cm[np.arange(len(q)), bc]
    cm2 = cm.copy()
    cm2[np.arange(len(q)), bc] = f32(1e30)
    b2 = cm2.min(1) if nch > 1 else np.full(len(q), f32(1e30))
    qq = (u[:, 2] * u[:, 2] + (u[:, 0] * u[:, 0] + (u[:, 1] * u[:, 1]).astype(f32)).astype(f32)).astype(f32)
    qn = np.sqrt(qq)
    S = (qn + np.sqrt(wmax)) ** 2
    rho = np.sqrt(np.maximum(b1 + qq, 0) + RHO_SLACK * S)
    seff = np.minimum(S, (2 * qn + rho) ** 2)
    margin = seff * MARGIN_REL + f32(1e-36)
    if abs_term:
        margin = margin + MARGIN_ABS * (f32(0.5) + f32(2.6) * qn + f32(0.9) * rho)
    with np.errstate(invalid="ignore"):
        trusted = (not bad) & (qn < 4096) & (b2 > b1 + margin)
    d = (t[None, :, :] - q[:, None, :]).astype(f32)
    dist = (d[..., 2] * d[..., 2] + (d[..., 0] * d[..., 0] + (d[..., 1] * d[..., 1]).astype(f32)).astype(f32)).astype(f32)
    kstar = dist.argmin(1)
    return trusted, trusted & (kstar // 32 != bc)


if __name__ == "__main__":
    from test_gpu_tc_hypothesis import KINDS, adversarial_cloud
    examples = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    rng = np.random.default_rng(2024)
    tot = {True: [0, 0, 0], False: [0, 0, 0]}
    for ex in range(examples):
        kq, kt = KINDS[rng.integers(len(KINDS))], KINDS[rng.integers(len(KINDS))]
        n, m = int(rng.integers(1, 2049)), int(rng.integers(1, 2049))
        x, y = adversarial_cloud(kq, rng, 1, n)[0], adversarial_cloud(kt, rng, 1, m)[0]
        if ex % 3 == 0:
            y = adversarial_cloud(kq, rng, 1, m)[0]
        for abs_term in (False, True):
            for qs, ts in ((x, y), (y, x)):
                tr, vio = model(qs, ts, abs_term)
                tot[abs_term][0] += len(qs); tot[abs_term][1] += int(tr.sum()); tot[abs_term][2] += int(vio.sum())
                if vio.any() and abs_term:
                    print(f"VIOLATION with the absolute term: {kq}/{kt} n={n} m={m} queries {np.flatnonzero(vio)[:5]}")
    for abs_term in (False, True):
        q, tr, v = tot[abs_term]
        print(f"margin {'with' if abs_term else 'without'} the absolute term: {q} queries, {tr} trusted ({100.0 * tr / q:.2f} %), {v} violations")
    # the fallback rate that matters for speed: uniform clouds at config 2's size
    x, y = rng.random((2048, 3), dtype=f32), rng.random((2048, 3), dtype=f32)
    for abs_term in (False, True):
        tr, vio = model(x, y, abs_term)
        print(f"U[0,1)^3 N=M=2048, {'with' if abs_term else 'without'} the absolute term: exact-scan fallbacks {100.0 * (1 - tr.mean()):.3f} %, violations {int(vio.sum())}")
