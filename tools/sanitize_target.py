"""Smallest launches of every kernel family, for compute-sanitizer (one --tool per gpurun call, VERDICT r1 missing 6):

    compute-sanitizer --tool memcheck  python tools/sanitize_target.py
    compute-sanitizer --tool racecheck python tools/sanitize_target.py [--no-tc]

Covers the tensor-core NN kernel (single- and multi-tile, fused zero fill), the FFMA NN kernel, both backward forms, the auction
kernel for cluster sizes 1/2/4/8 incl. the solo phase and the global-workspace form, the EMD gradient kernels, FPS, ICP, the
projection splat + its backward and the min-distance kernels.  Results are compared with nothing here (parity lives in tests/);
the point is a clean sanitizer log."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import psd_b200

pkg = psd_b200.load()
lib = pkg._lib.lib
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
vp = lambda t: ctypes.c_void_p(t.data_ptr())


def rnd(*s):
    return torch.rand(*s, generator=g).to(dev)


def chamfer(b, n, m, variant):
    old = lib.psd_chamfer_nn_variant(variant)
    x = rnd(b, n, 3).requires_grad_(True)
    y = rnd(b, m, 3).requires_grad_(True)
    d1, d2, i1, i2 = pkg.chamfer_3DDist()(x, y)
    (d1.mean() + d2.mean()).backward()
    loss = pkg.Loss().get_chamfer_loss(x.detach().transpose(1, 2).contiguous().transpose(1, 2).requires_grad_(True), y)
    loss.backward()
    a1 = torch.empty_like(x); a2 = torch.empty_like(y)
    assert lib.psd_chamfer_backward_ex(vp(x), vp(y), vp(a1), vp(a2), vp(d1.detach()), vp(d2.detach()), vp(i1), vp(i2), b, n, m, 0, 1, None) == 1
    torch.cuda.synchronize()
    lib.psd_chamfer_nn_variant(old)


print("chamfer ffma"); chamfer(2, 300, 520, 1)
if "--no-tc" not in sys.argv:
    print("chamfer tensor-core, single tile"); chamfer(2, 300, 520, 3)
    print("chamfer tensor-core, multi tile"); chamfer(1, 200, 2300, 3)
print("emd")
a, b_ = rnd(2, 1024, 3), rnd(2, 1024, 3)
for cluster in (1, 2, 4, 8, -2):
    dist = torch.zeros(2, 1024, device=dev); ass = torch.full((2, 1024), -1, device=dev, dtype=torch.int32)
    price = torch.zeros(2, 1024, device=dev); inv = torch.full((2, 1024), -1, device=dev, dtype=torch.int32)
    assert lib.psd_emd_forward_cluster(vp(a), vp(b_), 2, 1024, vp(dist), vp(ass), vp(price), vp(inv), 0.05, 40, cluster, None) == 1
    torch.cuda.synchronize()
ax = a.clone().requires_grad_(True)
pkg.Loss().get_emd_loss(ax, b_, eps=0.05, iters=60).backward()
dist, _ = pkg.emdModule()(ax, b_, 0.005, 8)
dist.sum().backward()
torch.cuda.synchronize()
print("fps / icp / splat / proj")
pkg.utils.farthest_point_sample(a, 32)
pkg.icp.icp_batch(a[:, :128].cpu().numpy(), b_[:, :128].cpu().numpy(), max_iterations=5, tolerance=1e-9)
p = (a[:, :200] * 1.8 - 0.9).requires_grad_(True)
img = pkg.projection.cont_proj(p, 32, 32, dev, 0.5)
img.sum().backward()
dm = torch.from_numpy(pkg.proj_loss.grid_dist(32, 32)).float()
pkg.proj_loss.get_loss_proj(img.detach().clamp(0, 1), img.detach().clamp(0, 1), dev, "bce_prob", 1.0, True, dm)
pkg.proj_loss.min_dist_terms(img.detach(), img.detach(), dm, mode="intended")
torch.cuda.synchronize()
print("sanitize target done")
