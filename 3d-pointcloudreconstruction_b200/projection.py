"""Drop-in for utils/projection.py's ``cont_proj(pcl, grid_h, grid_w, device, sigma_sq=0.5)`` (:4-67) and ``apply_kernel``
(:95-106).  The reference materialises a [B, N, H, W, 2] difference tensor on the CPU; here the splat is one CUDA kernel
(csrc/splat.cu, psd_cont_proj) with the reference's float32 rounding sequence and summation order.  No CPU fallback."""
import torch

try:
    from . import _lib
except ImportError:
    import _lib


def apply_kernel(x, sigma_sq=0.5):
    """Un-normalised Gaussian of the mean-subtracted grid input (utils/projection.py:95-106)."""
    return torch.exp(-(x ** 2) / (2. * sigma_sq))


def cont_proj(pcl, grid_h, grid_w, device, sigma_sq=0.5):
    """pcl [B, N, 3] in (-1, 1) -> silhouette [B, grid_h, grid_w] on `device` (the reference's callers pass 'cpu',
    utils/utils.py:232,241)."""
    if not torch.cuda.is_available():
        raise RuntimeError("cont_proj needs a CUDA device (there is no CPU fallback)")
    dev = pcl.device if pcl.is_cuda else torch.device("cuda", torch.cuda.current_device())
    p = pcl.detach().to(dev, torch.float32).contiguous()
    b, n, _ = p.shape
    out = torch.empty(b, grid_h, grid_w, device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        rc = _lib.lib.psd_cont_proj(_lib.ptr(p), b, n, int(grid_h), int(grid_w), float(sigma_sq), _lib.ptr(out),
                                    _lib.stream_of(p))
    _lib.raise_on_cuda_error(rc, "psd_cont_proj")
    return out.to(device)
