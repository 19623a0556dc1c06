"""Drop-in for loss/proj_loss.py: ``get_loss_proj(pred, gt, device, loss_type, w, min_dist_loss, dist_mat, opt)`` and
``grid_dist(grid_h, grid_w)`` with the reference's signatures and return values.

The BCE terms are the reference's torch expressions (proj_loss.py:8-19).  The min-distance branch (:21-40), which the
reference evaluates as dense [B,H,W,H,W] CPU tensors (8.6 GB at B=128, H=W=64), runs as one CUDA kernel through
psd_proj_min_dist.  ``mode="as_written"`` (default) reproduces the reference bit for bit, including its broadcast of both
images along the first pixel pair; ``mode="intended"`` evaluates the CAPNet-style masked nearest-pixel distance the code was
derived from (weights on the target pixel).  Like the reference, ``dist_mat`` is incremented in place by one per call."""
import numpy as np
import torch
import torch.nn as nn

try:
    from . import _lib
except ImportError:
    import _lib


def grid_dist(grid_h, grid_w):
    """Distance between every pair of grid points, [grid_h, grid_w, grid_h, grid_w] float64 (proj_loss.py:46-54;
    scipy's cdist on integer coordinates = sqrt of the integer squared distance in float64)."""
    hh = np.arange(grid_h, dtype=np.float64)
    ww = np.arange(grid_w, dtype=np.float64)
    dh = hh[:, None, None, None] - hh[None, None, :, None]
    dw = ww[None, :, None, None] - ww[None, None, None, :]
    return np.sqrt(dh * dh + dw * dw)


def _table_from_dist_mat(dist_mat, h, w):
    """The (|dh|,|dw|) table behind a [H,W,H,W] distance matrix: its slice at the origin pixel."""
    return dist_mat[0, 0].to(torch.float32).contiguous()


class _ForwardOnly(torch.autograd.Function):
    """Carries the kernel's values into the autograd graph of inputs that require grad, and refuses to be differentiated: the
    min-distance terms have no backward kernel (the reference only logs them, finetune.py:160-165, from detached clouds), and
    silently dropping their gradient would be worse than failing."""

    @staticmethod
    def forward(ctx, values, *inputs):
        return values.clone()

    @staticmethod
    def backward(ctx, grad):
        raise NotImplementedError("proj_loss min-distance terms are forward-only in this library (no backward kernel); "
                                  "detach them or use the reference's dense torch expression if a gradient is needed")


def min_dist_terms(pred, gt, dist_mat, mode="as_written"):
    """min_dist, min_dist_inv of proj_loss.py:25-41 for pred, gt [B,H,W]; dist_mat [H,W,H,W] AFTER its `+= 1`."""
    assert mode in ("as_written", "intended")
    if not torch.cuda.is_available():
        raise RuntimeError("proj_loss.min_dist_terms needs a CUDA device (there is no CPU fallback)")
    b, h, w = pred.shape
    dev = pred.device if pred.is_cuda else torch.device("cuda", torch.cuda.current_device())
    p = pred.detach().to(dev, torch.float32).contiguous()
    g = gt.detach().to(dev, torch.float32).contiguous()
    table = _table_from_dist_mat(dist_mat, h, w).to(dev)
    out = torch.empty(2, b, h, w, device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        rc = _lib.lib.psd_proj_min_dist(_lib.ptr(p), _lib.ptr(g), _lib.ptr(table), b, h, w, 0 if mode == "as_written" else 1,
                                        _lib.ptr(out[0]), _lib.ptr(out[1]), _lib.stream_of(p))
    _lib.raise_on_cuda_error(rc, "psd_proj_min_dist")
    a, b_ = out[0].to(pred.device), out[1].to(pred.device)   # the reference returns tensors on the inputs' device (CPU)
    if torch.is_grad_enabled() and (pred.requires_grad or gt.requires_grad):
        a, b_ = _ForwardOnly.apply(a, pred, gt), _ForwardOnly.apply(b_, pred, gt)
    return a, b_


def get_loss_proj(pred, gt, device, loss_type='bce', w=1., min_dist_loss=None, dist_mat=None, opt=None, mode="as_written"):
    loss = None
    if loss_type == 'bce':
        loss = nn.BCELoss()(gt, pred)
    if loss_type == 'weighted_bce':
        loss = nn.BCEWithLogitsLoss()(gt, pred)
    if loss_type == 'bce_prob':
        epsilon = 1e-8
        loss = -gt * torch.log(pred + epsilon) * w - (1 - gt) * torch.log(torch.abs(1 - pred - epsilon))
    min_dist = min_dist_inv = None
    if min_dist_loss is not None:
        dist_mat += 1                      # in place, like the reference (proj_loss.py:22)
        min_dist, min_dist_inv = min_dist_terms(pred, gt, dist_mat, mode=mode)
    return torch.mean(loss).cuda(), min_dist, min_dist_inv
