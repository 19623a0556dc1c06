"""Drop-in for metric/emd/emd_module.py: ``emdModule()(xyz1, xyz2, eps, iters) -> dist, assignment``.

Same contract as the reference (emd_module.py:9-19): clouds [B, n, 3] of equal size normalised to [0, 1],
n a multiple of 1024, B <= 512; gradient only for xyz1 (xyz2 gets zeros); the assignment is approximate and
not guaranteed to be a bijection.  The 12 scratch tensors of the reference wrapper (:43-54) are not needed:
the persistent kernel keeps the auction state in shared memory and starts from the same initial state."""
import torch
from torch import nn
from torch.autograd import Function

try:
    from . import _lib, emd
except ImportError:
    import _lib
    import emd


class emdFunction(Function):
    @staticmethod
    def forward(ctx, xyz1, xyz2, eps, iters):
        batchsize, n, _ = xyz1.size()
        _, m, _ = xyz2.size()

        assert n == m
        assert xyz1.size()[0] == xyz2.size()[0]
        assert n % 1024 == 0
        assert batchsize <= 512

        # (the reference's .cuda() moves CUDA tensors to the CURRENT device, emd_module.py:41-42; here a CUDA input stays on
        # its own device and only host tensors are moved)
        xyz1 = xyz1.contiguous().float()
        xyz1 = xyz1 if xyz1.is_cuda else xyz1.cuda()
        xyz2 = xyz2.contiguous().float().to(xyz1.device)
        dist = torch.empty(batchsize, n, device=xyz1.device, dtype=torch.float32)
        assignment = torch.empty(batchsize, n, device=xyz1.device, dtype=torch.int32)
        rc = emd.forward_fresh(xyz1, xyz2, dist, assignment, eps, iters)
        if rc != 1:
            raise RuntimeError(f"emd.forward failed (rc={rc}): {_lib.last_error()}")
        ctx.save_for_backward(xyz1, xyz2, assignment)
        ctx.mark_non_differentiable(assignment)
        return dist, assignment

    @staticmethod
    def backward(ctx, graddist, gradidx):
        xyz1, xyz2, assignment = ctx.saved_tensors
        graddist = graddist.contiguous()
        b, n, _ = xyz1.shape
        gradxyz1 = torch.empty_like(xyz1)     # one term per address: stored by the kernel, no zero fill
        with torch.cuda.device(xyz1.device):
            rc = _lib.lib.psd_emd_backward_ex(_lib.ptr(xyz1), _lib.ptr(xyz2), _lib.ptr(gradxyz1), _lib.ptr(graddist),
                                              _lib.ptr(assignment), b, n, 1, _lib.stream_of(xyz1))
        _lib.raise_on_cuda_error(rc, "emd.backward")
        gradxyz2 = torch.zeros_like(xyz2) if ctx.needs_input_grad[1] else None   # the reference returns zeros (emd_module.py:84-87)
        return gradxyz1, gradxyz2, None, None


class emdModule(nn.Module):
    def __init__(self):
        super(emdModule, self).__init__()

    def forward(self, input1, input2, eps, iters):
        return emdFunction.apply(input1, input2, eps, iters)
