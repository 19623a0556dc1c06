"""Read the in-kernel phase clocks of the grouped chamfer NN kernel (experimental build -DPSD_PROFILE_CLOCKS,
loaded through PSD_B200_LIB).  Prints, averaged over the compute warps of all CTAs: total cycles, start-up,
tile wait, scan (and cycles per 16-target x 128-query scan iteration), resolve, fallback; plus the SM clock
the kernel actually ran at (clock64 cycles / globaltimer ns).

    PSD_B200_LIB=$PWD/exp/libpsd_prof.so python tools/nn_phase_clocks.py [B N M]
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
pool = 6
xs = [torch.rand(b, n, 3, generator=g).to(dev) for _ in range(pool)]
ys = [torch.rand(b, m, 3, generator=g).to(dev) for _ in range(pool)]
out = (torch.empty(b, n, device=dev), torch.empty(b, m, device=dev),
       torch.empty(b, n, device=dev, dtype=torch.int32), torch.empty(b, m, device=dev, dtype=torch.int32))
L.psd_chamfer_nn_variant(2)
for s in range(40):  # back-to-back launches: the last one runs at the sustained clock
    assert pkg.chamfer_3D.forward(xs[s % pool], ys[s % pool], *out) == 1
torch.cuda.synchronize()
prof = np.zeros((148, 20, 8), np.int64)
assert L.psd_debug_read_prof(prof.ctypes.data_as(ctypes.c_void_p)) == 1
comp = prof[:, :16, :].astype(np.float64)
names = ["total", "start-up", "tile wait", "scan", "scan iterations", "resolve(+barriers)", "fallback", "globaltimer ns"]
print(f"B={b} N={n} M={m}  (mean / max over the 16 compute warps x 148 CTAs)")
for i, nm in enumerate(names):
    print(f"  {nm:20s} mean {comp[..., i].mean():12.1f}   max {comp[..., i].max():12.1f}")
it = comp[..., 4]
ok = it > 0
print(f"  cycles per scan iteration (per warp, 4 warps share a sub-partition): {(comp[..., 3][ok] / it[ok]).mean():.1f}")
print(f"  -> sub-partition cycles per iteration if 4 groups are active: {(comp[..., 3][ok] / it[ok]).mean() / 4:.1f}")
ghz = comp[..., 0] / np.maximum(comp[..., 7], 1)
print(f"  SM clock during the kernel: {ghz.mean():.3f} GHz (clock64 / globaltimer)")
per_cta_total = comp[..., 0].max(axis=1)
print(f"  slowest CTA {per_cta_total.max():.0f} cycles, fastest {per_cta_total.min():.0f}, mean {per_cta_total.mean():.0f}")
