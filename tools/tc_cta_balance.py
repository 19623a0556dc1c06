"""Per-CTA totals of the tensor-core chamfer NN kernel at config 2 (instrumented build): does a CTA whose range of units crosses a
cloud/direction boundary (second B operand build by the helpers) or has more deferred queries take longer?

    python tools/tc_cta_balance.py
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
b, n, m = 32, 2048, 2048
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
total = P[:, 4] - P[:, 0]
units = 2 * b * (n // 128)
beg = (np.arange(148) * units) // 148
end = ((np.arange(148) + 1) * units) // 148
nun = end - beg
per_group = n // 128
cross = (beg // per_group) != ((end - 1) // per_group)
first_cross = np.where(cross, ((beg // per_group) + 1) * per_group - beg, -1)   # unit index (in the CTA) of the first unit of the new group
opwait = P[:, 56]
print(f"CTAs: {148}, units per CTA {nun.min()}..{nun.max()}, crossing a group boundary: {cross.sum()}")
for sel, name in ((nun == nun.max(), "all CTAs with the larger unit count"), ((nun == nun.max()) & cross, "  crossing"), ((nun == nun.max()) & ~cross, "  not crossing")):
    print(f"{name:40s} n={sel.sum():3d}  total mean {total[sel].mean():8.0f}  max {total[sel].max():8.0f}   MMA thread waiting for operands mean {opwait[sel].mean():7.0f}")
for k in range(1, 8):
    sel = (nun == nun.max()) & (first_cross == k)
    if sel.any():
        print(f"  new group starts at unit {k}: n={sel.sum():3d}  total mean {total[sel].mean():8.0f}   operand wait {opwait[sel].mean():7.0f}")
order = np.argsort(-total)[:10]
print("slowest CTAs:", [(int(i), int(total[i]), bool(cross[i]), int(first_cross[i]), int(opwait[i])) for i in order])
