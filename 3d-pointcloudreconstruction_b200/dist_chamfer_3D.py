"""Drop-in for metric/chamfer3D/dist_chamfer_3D.py: ``chamfer_3DDist()(xyz1, xyz2) -> dist1, dist2, idx1, idx2``
with autograd to both clouds.  Differences from the reference wrapper, none of them visible in results:
outputs are allocated on the device directly (the reference allocates on the CPU and copies,
dist_chamfer_3D.py:40-49), there is no global ``torch.cuda.set_device`` side effect (:50), idx outputs are
marked non-differentiable, a failed launch raises instead of being ignored (:52), and a cloud that arrives as the
transposed view of a contiguous [B,3,N] tensor -- what the training loop passes, ``fake.transpose(2,1)`` at train.py:163 --
is read (and its gradient written) in place through strides instead of being copied by ``.contiguous()`` (:79-80)."""
import torch
from torch import nn
from torch.autograd import Function

try:
    from . import _lib
except ImportError:
    import _lib


def cloud_layout(t):
    """0 for a contiguous [B,N,3] cloud, 1 for the transposed view of a contiguous [B,3,N] tensor, None otherwise."""
    if t.dim() != 3 or t.shape[2] != 3:
        return None
    if t.is_contiguous():
        return 0
    b, n, _ = t.shape
    if t.stride(2) == n and t.stride(1) == 1 and (t.stride(0) == 3 * n or b == 1):
        return 1
    return None


def as_kernel_cloud(t):
    """(tensor the kernels can read in place, layout flag)."""
    lay = cloud_layout(t)
    if lay is None or t.dtype != torch.float32:
        return t.contiguous(), 0
    return t, lay


def grad_buffer_like(t, lay):
    """Uninitialised gradient with the memory layout of its cloud, seen as [B,N,3]."""
    if lay == 1:
        b, n, _ = t.shape
        return torch.empty(b, 3, n, device=t.device, dtype=torch.float32).transpose(1, 2)
    return torch.empty(t.shape, device=t.device, dtype=torch.float32)


class chamfer_3DFunction(Function):
    @staticmethod
    def forward(ctx, xyz1, xyz2):
        batchsize, n, dim = xyz1.size()
        assert dim == 3, "Wrong last dimension for the chamfer distance 's input! Check with .size()"
        _, m, dim = xyz2.size()
        assert dim == 3, "Wrong last dimension for the chamfer distance 's input! Check with .size()"
        for name, t in (("xyz1", xyz1), ("xyz2", xyz2)):
            if not t.is_cuda:
                raise RuntimeError(f"{name}: expected a CUDA tensor (this op has no CPU path)")
            if t.dtype != torch.float32:
                raise RuntimeError(f"{name}: expected dtype torch.float32, got {t.dtype}")
        if xyz2.size(0) != batchsize or xyz2.device != xyz1.device:
            raise RuntimeError("chamfer_3DFunction: xyz1 and xyz2 need the same batch size and device")
        xyz1, l1 = as_kernel_cloud(xyz1)
        xyz2, l2 = as_kernel_cloud(xyz2)
        layout = l1 | (l2 << 1)
        device = xyz1.device
        dist1 = torch.empty(batchsize, n, device=device, dtype=torch.float32)
        dist2 = torch.empty(batchsize, m, device=device, dtype=torch.float32)
        idx1 = torch.empty(batchsize, n, device=device, dtype=torch.int32)
        idx2 = torch.empty(batchsize, m, device=device, dtype=torch.int32)
        if n == 0 or m == 0 or batchsize == 0:  # the reference leaves its zero-filled outputs untouched
            dist1.zero_(); dist2.zero_(); idx1.zero_(); idx2.zero_()
        with torch.cuda.device(device):
            rc = _lib.lib.psd_chamfer_forward_ex(_lib.ptr(xyz1), _lib.ptr(xyz2), batchsize, n, m, layout, _lib.ptr(dist1),
                                                 _lib.ptr(dist2), _lib.ptr(idx1), _lib.ptr(idx2), None, 0.0, None, 0, -1,
                                                 _lib.stream_of(xyz1))
        if rc != 1:
            raise RuntimeError(f"chamfer_3D.forward failed (rc={rc}): {_lib.last_error()}")
        ctx.save_for_backward(xyz1, xyz2, idx1, idx2)
        ctx.layout = layout
        ctx.mark_non_differentiable(idx1, idx2)
        return dist1, dist2, idx1, idx2

    @staticmethod
    def backward(ctx, graddist1, graddist2, gradidx1, gradidx2):
        xyz1, xyz2, idx1, idx2 = ctx.saved_tensors
        layout = ctx.layout
        b, n, _ = xyz1.shape
        m = xyz2.shape[1]
        graddist1 = graddist1.contiguous()
        graddist2 = graddist2.contiguous()
        # no zero fill: the two-phase backward kernel stores every element once, then adds the scatter terms
        gradxyz1 = grad_buffer_like(xyz1, layout & 1)
        gradxyz2 = grad_buffer_like(xyz2, (layout >> 1) & 1)
        with torch.cuda.device(xyz1.device):
            rc = _lib.lib.psd_chamfer_backward_ex(_lib.ptr(xyz1), _lib.ptr(xyz2), _lib.ptr(gradxyz1), _lib.ptr(gradxyz2),
                                                  _lib.ptr(graddist1), _lib.ptr(graddist2), _lib.ptr(idx1), _lib.ptr(idx2),
                                                  b, n, m, layout, 1, _lib.stream_of(xyz1))
        if rc != 1:
            raise RuntimeError(f"chamfer_3D.backward failed (rc={rc}): {_lib.last_error()}")
        return gradxyz1, gradxyz2


class chamfer_3DDist(nn.Module):
    def __init__(self):
        super(chamfer_3DDist, self).__init__()

    def forward(self, input1, input2):
        # (the reference calls .contiguous() here; chamfer_3DFunction reads [B,N,3] and transposed [B,3,N] views in place
        # and copies only what it cannot address through strides)
        return chamfer_3DFunction.apply(input1, input2)
