"""Drop-in for metric/chamfer3D/dist_chamfer_3D.py: ``chamfer_3DDist()(xyz1, xyz2) -> dist1, dist2, idx1, idx2``
with autograd to both clouds.  Differences from the reference wrapper, none of them visible in results:
outputs are allocated on the device directly (the reference allocates on the CPU and copies,
dist_chamfer_3D.py:40-49), there is no global ``torch.cuda.set_device`` side effect (:50), idx outputs are
marked non-differentiable, and a failed launch raises instead of being ignored (:52)."""
import torch
from torch import nn
from torch.autograd import Function

try:
    from . import _lib, chamfer_3D
except ImportError:
    import _lib
    import chamfer_3D


class chamfer_3DFunction(Function):
    @staticmethod
    def forward(ctx, xyz1, xyz2):
        batchsize, n, dim = xyz1.size()
        assert dim == 3, "Wrong last dimension for the chamfer distance 's input! Check with .size()"
        _, m, dim = xyz2.size()
        assert dim == 3, "Wrong last dimension for the chamfer distance 's input! Check with .size()"
        device = xyz1.device
        dist1 = torch.empty(batchsize, n, device=device, dtype=torch.float32)
        dist2 = torch.empty(batchsize, m, device=device, dtype=torch.float32)
        idx1 = torch.empty(batchsize, n, device=device, dtype=torch.int32)
        idx2 = torch.empty(batchsize, m, device=device, dtype=torch.int32)
        if n == 0 or m == 0 or batchsize == 0:  # the reference leaves its zero-filled outputs untouched
            dist1.zero_(); dist2.zero_(); idx1.zero_(); idx2.zero_()
        _lib.raise_on_cuda_error(chamfer_3D.forward(xyz1, xyz2, dist1, dist2, idx1, idx2), "chamfer_3D.forward")
        ctx.save_for_backward(xyz1, xyz2, idx1, idx2)
        ctx.mark_non_differentiable(idx1, idx2)
        return dist1, dist2, idx1, idx2

    @staticmethod
    def backward(ctx, graddist1, graddist2, gradidx1, gradidx2):
        xyz1, xyz2, idx1, idx2 = ctx.saved_tensors
        graddist1 = graddist1.contiguous()
        graddist2 = graddist2.contiguous()
        # one zero-filled allocation for both gradients (a single memset instead of two)
        n1, n2 = xyz1.numel(), xyz2.numel()
        buf = torch.zeros(n1 + n2, device=xyz1.device, dtype=torch.float32)
        gradxyz1 = buf[:n1].view(xyz1.size())
        gradxyz2 = buf[n1:].view(xyz2.size())
        _lib.raise_on_cuda_error(
            chamfer_3D.backward(xyz1, xyz2, gradxyz1, gradxyz2, graddist1, graddist2, idx1, idx2), "chamfer_3D.backward")
        return gradxyz1, gradxyz2


class chamfer_3DDist(nn.Module):
    def __init__(self):
        super(chamfer_3DDist, self).__init__()

    def forward(self, input1, input2):
        input1 = input1.contiguous()
        input2 = input2.contiguous()
        return chamfer_3DFunction.apply(input1, input2)
