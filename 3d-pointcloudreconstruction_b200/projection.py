"""Drop-in for utils/projection.py's ``cont_proj(pcl, grid_h, grid_w, device, sigma_sq=0.5)`` (:4-67) and ``apply_kernel``
(:95-106).  The reference materialises a [B, N, H, W, 2] difference tensor on the CPU; here the splat is one CUDA kernel
(csrc/splat.cu, psd_cont_proj) with the reference's float32 rounding sequence and summation order, and -- because the
reference's op is plain differentiable torch code -- an autograd backward (psd_cont_proj_backward) that carries the
silhouette's gradient back to the cloud's x and y coordinates.  No CPU fallback."""
import torch

try:
    from . import _lib
except ImportError:
    import _lib


def apply_kernel(x, sigma_sq=0.5):
    """Un-normalised Gaussian of the mean-subtracted grid input (utils/projection.py:95-106)."""
    return torch.exp(-(x ** 2) / (2. * sigma_sq))


class _ContProj(torch.autograd.Function):
    @staticmethod
    def forward(ctx, p, grid_h, grid_w, sigma_sq):
        b, n, _ = p.shape
        out = torch.empty(b, grid_h, grid_w, device=p.device, dtype=torch.float32)
        with torch.cuda.device(p.device):
            rc = _lib.lib.psd_cont_proj(_lib.ptr(p), b, n, grid_h, grid_w, sigma_sq, _lib.ptr(out), _lib.stream_of(p))
        _lib.raise_on_cuda_error(rc, "psd_cont_proj")
        ctx.save_for_backward(p)
        ctx.grid = (grid_h, grid_w, sigma_sq)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (p,) = ctx.saved_tensors
        grid_h, grid_w, sigma_sq = ctx.grid
        b, n, _ = p.shape
        g = grad_out.to(torch.float32).contiguous()
        gp = torch.empty_like(p)
        with torch.cuda.device(p.device):
            rc = _lib.lib.psd_cont_proj_backward(_lib.ptr(p), _lib.ptr(g), b, n, grid_h, grid_w, sigma_sq, _lib.ptr(gp),
                                                 _lib.stream_of(p))
        if rc != 1:
            raise RuntimeError(f"psd_cont_proj_backward failed (rc={rc}): {_lib.last_error()}")
        return gp, None, None, None


def cont_proj(pcl, grid_h, grid_w, device, sigma_sq=0.5):
    """pcl [B, N, 3] in (-1, 1) -> silhouette [B, grid_h, grid_w] on `device` (the reference's callers pass 'cpu',
    utils/utils.py:232,241).  Differentiable w.r.t. pcl like the reference's torch expression."""
    if not torch.cuda.is_available():
        raise RuntimeError("cont_proj needs a CUDA device (there is no CPU fallback)")
    if pcl.dim() != 3 or pcl.shape[2] != 3:
        raise RuntimeError("cont_proj: expected pcl [B, N, 3]")
    dev = pcl.device if pcl.is_cuda else torch.device("cuda", torch.cuda.current_device())
    p = pcl.to(dev, torch.float32).contiguous()      # differentiable copies: the gradient flows back to `pcl`
    out = _ContProj.apply(p, int(grid_h), int(grid_w), float(sigma_sq))
    return out.to(device)
