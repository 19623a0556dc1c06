"""Per-iteration phase clocks of the auction kernel at BASELINE config 3 (instrumented build: -DPSD_EMD_PROF, loaded through
PSD_B200_LIB): cycles of compaction + count exchange, bid scan, the three cluster barriers with GetMax / Assign, for block 0.

    PSD_B200_LIB=exp/libpsd_emdprof.so python tools/emd_phase_clocks.py [B n eps iters]
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
B, n, eps, iters = (int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) >= 5 else (32, 2048, 0.005, 50)
g = torch.Generator().manual_seed(0)
x = torch.rand(B, n, 3, generator=g).to(dev)
y = torch.rand(B, n, 3, generator=g).to(dev)
d = torch.empty(B, n, device=dev)
a = torch.empty(B, n, device=dev, dtype=torch.int32)
for _ in range(3):
    pkg.emd.forward_fresh(x, y, d, a, eps, iters)
torch.cuda.synchronize()
buf = np.zeros(256 * 8, dtype=np.int64)
fn = ctypes.CDLL(os.environ["PSD_B200_LIB"]).psd_debug_emd_prof
assert fn(buf.ctypes.data_as(ctypes.c_void_p)) == 1
P = buf.reshape(256, 8)
print(" it  bidders grid   compact      bid   barrierA  getmax+B  assign+C     total")
tot = np.zeros(6)
for it in range(min(iters, 256)):
    r = P[it]
    if r[0] == 0:
        break
    nxt = P[it + 1][0] if it + 1 < 256 and P[it + 1][0] else r[5]
    ph = [r[1] - r[0], r[2] - r[1], r[3] - r[2], r[4] - r[3], r[5] - r[4], r[5] - r[0]]
    tot += ph
    if it < 12 or it % 5 == 0:
        print(f"{it:3d} {int(r[6]):8d} {int(r[7]):4d} " + " ".join(f"{int(v):9d}" for v in ph))
print("sum over iterations: " + " ".join(f"{int(v):9d}" for v in tot))
