"""Phase clocks of the tensor-core chamfer NN kernel (csrc/chamfer_nn_tc.cu): where do a CTA's cycles go?

Runs a few warm launches, then one instrumented launch (psd_debug_tc_prof) and prints, averaged over the CTAs,
the clock64 deltas between the stamps of consumer thread 0 and the wait totals of the MMA warp.

    python tools/tc_phase_clocks.py [B N M]
"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import psd_b200

pkg = psd_b200.load()
L = pkg._lib.lib
dev = torch.device("cuda:0")
b, n, m = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (32, 2048, 2048)
g = torch.Generator().manual_seed(3)
x = torch.rand(b, n, 3, generator=g).to(dev)
y = torch.rand(b, m, 3, generator=g).to(dev)
out = (torch.empty(b, n, device=dev), torch.empty(b, m, device=dev),
       torch.empty(b, n, device=dev, dtype=torch.int32), torch.empty(b, m, device=dev, dtype=torch.int32))
L.psd_chamfer_nn_variant(3)
prof = torch.zeros(148 * 64 + 512, dtype=torch.int64, device=dev)
L.psd_debug_tc_prof(ctypes.c_void_p(prof.data_ptr()))
for _ in range(5):
    prof.zero_()
    assert pkg.chamfer_3D.forward(x, y, *out) == 1
torch.cuda.synchronize()
L.psd_debug_tc_prof(None)
P = prof.cpu().numpy()[:148 * 64].reshape(148, 64).astype(np.float64)
act = P[:, 0] > 0
P = P[act]
print(f"B={b} N={n} M={m}: {act.sum()} CTAs")
t0 = P[:, 0]
print(f"  setup (barrier init, TMEM alloc)        {np.mean(P[:, 1] - t0):9.0f}")
print(f"  prologue: first unit's loads issued at   {np.mean(P[:, 6] - t0):9.0f}")
print(f"  prologue: raw targets landed at          {np.mean(P[:, 7] - t0):9.0f}")
print(f"  prologue: first B operand built (CTA)   {np.mean(P[:, 5] - t0):9.0f}")
print(f"  helpers: units 0 and 1 staged at        {np.mean(P[:, 2] - t0):9.0f}")
prev_scan = P[:, 2].copy()
for u in range(8):
    base = 8 + u * 6
    have = P[:, base] > 0
    if not have.any():
        break
    sc, pw, rs, st = P[have, base], P[have, base + 1], P[have, base + 2], P[have, base + 3]
    ra, rb = P[have, base + 4], P[have, base + 5]
    print(f"  unit {u} ({have.sum():3d} CTAs): scanners done at {np.mean(sc - t0[have]):8.0f} (+{np.mean(sc - prev_scan[have]):6.0f})   "
          f"helpers: partials seen {np.mean(pw - t0[have]):8.0f}, resolve +{np.mean(rs - pw):6.0f} (A {np.mean(ra - pw):5.0f} B {np.mean(rb - ra):5.0f} C {np.mean(rs - rb):5.0f}), next staged +{np.mean(st - rs):6.0f}")
    prev_scan[have] = sc
print(f"  all roles done (before deferred scans)  {np.mean(P[:, 3] - t0):9.0f}")
print(f"  deferred exact scans                    {np.mean(P[:, 4] - P[:, 3]):9.0f}")
print(f"  total                                   {np.mean(P[:, 4] - t0):9.0f}  (max {np.max(P[:, 4] - t0):.0f})")
print(f"  scanner warp 0 waiting on full barriers {np.mean(P[:, 59]):9.0f}")
print(f"  helpers waiting for parked partials     {np.mean(P[:, 60]):9.0f}")
print(f"  MMA warp: waiting for operands          {np.mean(P[:, 56]):9.0f}")
print(f"  MMA warp: waiting for empty buffers     {np.mean(P[:, 57]):9.0f}")
print(f"  MMA warp: last issue at                 {np.mean(P[:, 58] - t0):9.0f}")
print(f"  MMA warp: inside tcgen05.mma issue      {np.mean(P[:, 61]):9.0f}")
print(f"  MMA warp: inside tcgen05.commit issue   {np.mean(P[:, 62]):9.0f}")
