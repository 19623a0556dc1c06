"""Drop-in for the sampling helpers of utils/utils.py: ``farthest_point_sample(xyz, npoint, RAN=True)`` (:335-360) and
``index_points(points, idx)`` (:318-333), with the reference's signatures and return types (torch.long indices on the
input's device).  The sampling loop runs as one CUDA kernel (csrc/fps.cu, psd_farthest_point_sample); there is no CPU
fallback -- CPU inputs are staged to the current CUDA device and the result is returned on the input's device."""
import torch

try:
    from . import _lib
except ImportError:
    import _lib


def farthest_point_sample(xyz, npoint, RAN=True):
    """xyz [B, N, 3] -> centroids [B, npoint] (torch.long).  The reference's 'random' start is torch.randint(0, 1) = 0 when
    RAN else torch.randint(1, 2) = 1 (utils/utils.py:347-350)."""
    if not torch.cuda.is_available():
        raise RuntimeError("farthest_point_sample needs a CUDA device (there is no CPU fallback)")
    B, N, C = xyz.shape
    assert C == 3, "the CUDA path samples 3-D clouds"
    dev = xyz.device if xyz.is_cuda else torch.device("cuda", torch.cuda.current_device())
    x = xyz.detach().to(dev, torch.float32).contiguous()
    centroids = torch.zeros(B, npoint, dtype=torch.long, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.lib.psd_farthest_point_sample(_lib.ptr(x), B, N, int(npoint), 0 if RAN else 1, _lib.ptr(centroids),
                                                _lib.stream_of(x))
    _lib.raise_on_cuda_error(rc, "psd_farthest_point_sample")
    return centroids.to(xyz.device)


def index_points(points, idx):
    """points [B, N, C], idx [B, S] (or [B, S, K]) -> [B, S, C] (utils/utils.py:318-333): a batched gather."""
    B = points.shape[0]
    view_shape = [B] + [1] * (idx.dim() - 1)
    batch_indices = torch.arange(B, dtype=torch.long, device=points.device).view(view_shape).expand_as(idx)
    return points[batch_indices, idx, :]
