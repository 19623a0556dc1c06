"""Drop-in for utils/icp.py: ``best_fit_transform(A, B)``, ``nearest_neighbor(src, dst)`` and
``icp(A, B, init_pose=None, max_iterations=20, tolerance=0.001)`` with the reference's signatures, numpy in / numpy out, plus
``icp_batch`` which aligns a whole batch of samples in one kernel launch (the reference's callers loop over the samples,
testnet.py:62-64).

The reference runs a sklearn KD-tree build + query, numpy reductions and a 3x3 SVD per iteration on the CPU; here the whole
iteration loop of every sample runs inside one CUDA kernel (csrc/icp.cu, psd_icp_batch), in fp64 like numpy.  There is no
CPU fallback."""
import ctypes

import numpy as np
import torch

try:
    from . import _lib
except ImportError:
    import _lib


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("icp needs a CUDA device (there is no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def _to_dev(x, dev):
    """[..., n, 3] numpy / torch input -> contiguous device tensor, float32 kept (exact in fp64), everything else float64."""
    t = torch.as_tensor(x)
    t = t.to(dev, torch.float32 if t.dtype == torch.float32 else torch.float64)
    return t.contiguous()


def icp_batch(A, B, init_pose=None, max_iterations=20, tolerance=0.001):
    """icp() for every sample of A, B [batch, n, 3] (numpy or torch, float32 or float64; both must share the dtype).
    Returns (T [batch,4,4] float64, distances [batch,n] float64, iterations [batch] int32) as torch tensors on the device."""
    dev = A.device if isinstance(A, torch.Tensor) and A.is_cuda else _device()
    a, b = _to_dev(A, dev), _to_dev(B, dev)
    assert a.shape == b.shape and a.dim() == 3 and a.shape[2] == 3, "A and B must both be [batch, n, 3]"
    if a.dtype != b.dtype:
        a, b = a.double(), b.double()
    batch, n, _ = a.shape
    pose = None if init_pose is None else torch.as_tensor(np.asarray(init_pose, dtype=np.float64)).to(dev).contiguous()
    if pose is not None:
        assert pose.shape == (4, 4)
    T = torch.empty(batch, 4, 4, device=dev, dtype=torch.float64)
    dist = torch.empty(batch, n, device=dev, dtype=torch.float64)
    iters = torch.zeros(batch, device=dev, dtype=torch.int32)
    with torch.cuda.device(dev):
        rc = _lib.lib.psd_icp_batch(_lib.ptr(a), _lib.ptr(b), int(a.dtype == torch.float64), batch, n,
                                    None if pose is None else _lib.ptr(pose), int(max_iterations), ctypes.c_double(tolerance),
                                    _lib.ptr(T), _lib.ptr(dist), _lib.ptr(iters), _lib.stream_of(a))
    _lib.raise_on_cuda_error(rc, "psd_icp_batch")
    return T, dist, iters


def icp(A, B, init_pose=None, max_iterations=20, tolerance=0.001):
    """utils/icp.py:68-118: (T, distances, i) for one pair of [N, 3] clouds."""
    assert A.shape == B.shape
    assert max_iterations >= 1, "the reference's icp() needs at least one iteration"
    T, dist, iters = icp_batch(torch.as_tensor(A)[None], torch.as_tensor(B)[None], init_pose, max_iterations, tolerance)
    return T[0].cpu().numpy(), dist[0].cpu().numpy(), int(iters[0].item())


def best_fit_transform(A, B):
    """utils/icp.py:4-46: (T, R, t) mapping the corresponding points A onto B (3-D, evaluated in float64)."""
    assert A.shape == B.shape and A.shape[1] == 3, "the CUDA path is 3-D only"
    dev = _device()
    a = torch.as_tensor(np.asarray(A, dtype=np.float64)).to(dev)[None].contiguous()
    b = torch.as_tensor(np.asarray(B, dtype=np.float64)).to(dev)[None].contiguous()
    T = torch.empty(1, 4, 4, device=dev, dtype=torch.float64)
    with torch.cuda.device(dev):
        rc = _lib.lib.psd_icp_batch(_lib.ptr(a), _lib.ptr(b), 1, 1, a.shape[1], None, 0, ctypes.c_double(0.0), _lib.ptr(T),
                                    None, None, _lib.stream_of(a))
    _lib.raise_on_cuda_error(rc, "psd_icp_batch")
    T = T[0].cpu().numpy()
    return T, T[:3, :3].copy(), T[:3, 3].copy()


def nearest_neighbor(src, dst):
    """utils/icp.py:49-65: Euclidean distance and dst index of the nearest neighbour of every src point."""
    assert src.shape == dst.shape
    dev = _device()
    s, d = _to_dev(src, dev)[None], _to_dev(dst, dev)[None]
    if s.dtype != d.dtype:
        s, d = s.double(), d.double()
    n = s.shape[1]
    dist = torch.empty(1, n, device=dev, dtype=torch.float64)
    idx = torch.empty(1, n, device=dev, dtype=torch.int32)
    with torch.cuda.device(dev):
        rc = _lib.lib.psd_nn_f64(_lib.ptr(s), _lib.ptr(d), int(s.dtype == torch.float64), 1, n, d.shape[1], _lib.ptr(dist),
                                 _lib.ptr(idx), _lib.stream_of(s))
    _lib.raise_on_cuda_error(rc, "psd_nn_f64")
    return dist[0].cpu().numpy(), idx[0].cpu().numpy().astype(np.int64)
