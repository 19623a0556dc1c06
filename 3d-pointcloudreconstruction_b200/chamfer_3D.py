"""Mirror of the reference's native module `chamfer_3D` (PYBIND11_MODULE in
metric/chamfer3D/chamfer_cuda.cpp:30-33): the same two functions, the same argument lists, the same
int return (1 ok / 0 CUDA error), backed by libpsd_b200.so instead of chamfer3D.cu."""
import torch

try:
    from . import _lib
except ImportError:  # imported as a top-level module (directory on sys.path, as the reference does)
    import _lib


def forward(xyz1, xyz2, dist1, dist2, idx1, idx2) -> int:
    """chamfer_forward (chamfer_cuda.cpp:17-19).  Caller-allocated outputs, written in place."""
    _lib.check_tensor("xyz1", xyz1, torch.float32, 3)
    _lib.check_tensor("xyz2", xyz2, torch.float32, 3)
    _lib.check_tensor("dist1", dist1, torch.float32)
    _lib.check_tensor("dist2", dist2, torch.float32)
    _lib.check_tensor("idx1", idx1, torch.int32)
    _lib.check_tensor("idx2", idx2, torch.int32)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    if xyz1.shape[2] != 3 or xyz2.shape[2] != 3 or xyz2.shape[0] != b:
        raise RuntimeError("chamfer_3D.forward: expected xyz1 [B,N,3] and xyz2 [B,M,3]")
    if dist1.numel() != b * n or idx1.numel() != b * n or dist2.numel() != b * m or idx2.numel() != b * m:
        raise RuntimeError("chamfer_3D.forward: output tensors have the wrong size")
    with torch.cuda.device(xyz1.device):
        return _lib.lib.psd_chamfer_forward(_lib.ptr(xyz1), _lib.ptr(xyz2), b, n, m, _lib.ptr(dist1), _lib.ptr(dist2),
                                            _lib.ptr(idx1), _lib.ptr(idx2), _lib.stream_of(xyz1))


def backward(xyz1, xyz2, gradxyz1, gradxyz2, graddist1, graddist2, idx1, idx2) -> int:
    """chamfer_backward (chamfer_cuda.cpp:22-26).  Accumulates into the caller-zeroed gradxyz1/gradxyz2."""
    _lib.check_tensor("xyz1", xyz1, torch.float32, 3)
    _lib.check_tensor("xyz2", xyz2, torch.float32, 3)
    _lib.check_tensor("gradxyz1", gradxyz1, torch.float32)
    _lib.check_tensor("gradxyz2", gradxyz2, torch.float32)
    _lib.check_tensor("graddist1", graddist1, torch.float32)
    _lib.check_tensor("graddist2", graddist2, torch.float32)
    _lib.check_tensor("idx1", idx1, torch.int32)
    _lib.check_tensor("idx2", idx2, torch.int32)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    if xyz1.shape[2] != 3 or xyz2.shape[2] != 3 or xyz2.shape[0] != b:
        raise RuntimeError("chamfer_3D.backward: expected xyz1 [B,N,3] and xyz2 [B,M,3]")
    if gradxyz1.numel() != b * n * 3 or gradxyz2.numel() != b * m * 3:
        raise RuntimeError("chamfer_3D.backward: gradient tensors have the wrong size")
    with torch.cuda.device(xyz1.device):
        return _lib.lib.psd_chamfer_backward(_lib.ptr(xyz1), _lib.ptr(xyz2), _lib.ptr(gradxyz1), _lib.ptr(gradxyz2),
                                             _lib.ptr(graddist1), _lib.ptr(graddist2), _lib.ptr(idx1), _lib.ptr(idx2),
                                             b, n, m, _lib.stream_of(xyz1))


def forward_ex(xyz1, xyz2, dist1, dist2, idx1, idx2, layout=0, sums=None, fs_thr=1e-4, fs_count=None, q_begin=0,
               q_count=-1) -> int:
    """psd_chamfer_forward_ex: fused loss sums / F-score counts, query slicing, and the layout bit mask
    (1: xyz1 is a contiguous [B,3,N] tensor, 2: xyz2 is a contiguous [B,3,M] tensor; else [B,N,3])."""
    if layout not in (0, 1, 2, 3):
        raise ValueError("layout is a bit mask: 1 = xyz1 is [B,3,N], 2 = xyz2 is [B,3,M]")
    b = xyz1.shape[0]
    n = xyz1.shape[2] if layout & 1 else xyz1.shape[1]
    m = xyz2.shape[2] if layout & 2 else xyz2.shape[1]
    if xyz1.dim() != 3 or xyz2.dim() != 3 or xyz1.shape[1 if layout & 1 else 2] != 3 or xyz2.shape[1 if layout & 2 else 2] != 3:
        raise RuntimeError("chamfer_3D.forward_ex: clouds must be [B,N,3] (or [B,3,N] where the layout mask says so)")
    for name, t, dt in (("xyz1", xyz1, torch.float32), ("xyz2", xyz2, torch.float32), ("dist1", dist1, torch.float32),
                        ("dist2", dist2, torch.float32), ("idx1", idx1, torch.int32), ("idx2", idx2, torch.int32)):
        _lib.check_tensor(name, t, dt)
    if sums is not None:
        _lib.check_tensor("sums", sums, torch.float32)
    if fs_count is not None:
        _lib.check_tensor("fs_count", fs_count, torch.int32)
    with torch.cuda.device(xyz1.device):
        return _lib.lib.psd_chamfer_forward_ex(_lib.ptr(xyz1), _lib.ptr(xyz2), b, n, m, int(layout), _lib.ptr(dist1),
                                               _lib.ptr(dist2), _lib.ptr(idx1), _lib.ptr(idx2), _lib.ptr(sums),
                                               float(fs_thr), _lib.ptr(fs_count), int(q_begin), int(q_count),
                                               _lib.stream_of(xyz1))
