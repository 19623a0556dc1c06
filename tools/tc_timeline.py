"""Per-tile timeline of CTA 0 of the tensor-core chamfer NN kernel (instrumented build, csrc/chamfer_nn_tc.cu `tl`):
when was each 256-column tile issued / committed by the MMA thread, seen full / released / reduced by scanner warps 0 and 4.

    python tools/tc_timeline.py [B N M]
"""
import ctypes, os, sys
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
P = prof.cpu().numpy()
t0 = P[0]
T = P[148 * 64:].reshape(8, 64).astype(np.int64)
names = ["mma", "commit", "w0 full", "w0 rel", "w0 done", "w4 full", "w4 rel", "w4 done"]
print("tile  " + "  ".join(f"{s:>8s}" for s in names) + "   d(mma)")
prev = None
for gi in range(64):
    if T[0, gi] == 0:
        break
    row = [int(T[r, gi] - t0) if T[r, gi] else -1 for r in range(8)]
    print(f"{gi:4d}  " + "  ".join(f"{v:8d}" for v in row) + (f"   {row[0] - prev:6d}" if prev is not None else ""))
    prev = row[0]
hs = P[:64]
print("helper stamps of CTA 0 (unit: scanners done / partials seen / resolved / next staged):")
for u in range(8):
    base = 8 + u * 6
    if hs[base] == 0:
        break
    print(f"  unit {u}: " + "  ".join(f"{int(hs[base + k] - t0):7d}" for k in range(4)))
