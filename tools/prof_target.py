"""Small fixed workload for ncu: 3x chamfer forward (the launch that also zero-fills the gradients), 3x backward, 1x EMD at
BASELINE configs 2/3 (+ --emd-train: 1x EMD at the training setting, --c5: one B=1 N=M=131072 multi-tile forward)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import psd_b200
pkg = psd_b200.load()
dev = torch.device("cuda:0")
if os.environ.get("PSD_NN_VARIANT"):
    pkg._lib.lib.psd_chamfer_nn_variant(int(os.environ["PSD_NN_VARIANT"]))
B, N = 32, 2048
torch.manual_seed(0)
x = torch.rand(B, N, 3).to(dev); y = torch.rand(B, N, 3).to(dev)
d1 = torch.empty(B, N, device=dev); d2 = torch.empty(B, N, device=dev)
i1 = torch.empty(B, N, device=dev, dtype=torch.int32); i2 = torch.empty(B, N, device=dev, dtype=torch.int32)
g1 = torch.rand(B, N, device=dev); g2 = torch.rand(B, N, device=dev)
gx1 = torch.zeros(B, N, 3, device=dev); gx2 = torch.zeros(B, N, 3, device=dev)
dist = torch.empty(B, N, device=dev); ass = torch.empty(B, N, device=dev, dtype=torch.int32)
import ctypes
vp = lambda t: ctypes.c_void_p(t.data_ptr())
gbuf = torch.empty(6 * B * N, device=dev)
for _ in range(3):
    assert pkg._lib.lib.psd_chamfer_forward_zero(vp(x), vp(y), B, N, N, 0, vp(d1), vp(d2), vp(i1), vp(i2), None, 0.0, None, vp(gbuf), gbuf.numel(), None) == 1
for _ in range(3):
    assert pkg._lib.lib.psd_chamfer_backward(vp(x), vp(y), vp(gbuf), vp(gbuf[3 * B * N:]), vp(g1), vp(g2), vp(i1), vp(i2), B, N, N, None) == 1
if "--c5" in sys.argv:
    xl = torch.rand(1, 131072, 3).to(dev); yl = torch.rand(1, 131072, 3).to(dev)
    out = pkg.chamfer_fscore_fused(xl, yl)
if "--no-emd" not in sys.argv:
    assert pkg.emd.forward_fresh(x, y, dist, ass, 0.005, 50) == 1
if "--emd-train" in sys.argv:   # the training setting: eps 0.05, 3000 iterations, n = 1024 (solo mode for most iterations)
    xt, yt = x[:, :1024].contiguous(), y[:, :1024].contiguous()
    dt_, at_ = torch.empty(B, 1024, device=dev), torch.empty(B, 1024, device=dev, dtype=torch.int32)
    assert pkg.emd.forward_fresh(xt, yt, dt_, at_, 0.05, 3000) == 1
torch.cuda.synchronize()
print("ok", float(d1.sum()), int(ass.sum()))
