"""Mirror of the reference's native module `emd` (PYBIND11_MODULE in metric/emd/emd.cpp:25-29): the same
two functions with the same argument lists and int returns (1 ok / 0 CUDA error / -1 shape violation),
backed by the persistent cluster kernel of libpsd_b200.so instead of emd_cuda.cu's 7 launches/iteration."""
import torch

try:
    from . import _lib
except ImportError:
    import _lib


def forward(xyz1, xyz2, dist, assignment, price, assignment_inv, bid, bid_increments, max_increments, unass_idx,
            unass_cnt, unass_cnt_sum, cnt_tmp, max_idx, eps, iters) -> int:
    """emd_forward (emd.cpp:12-17).  State tensors pre-initialised by the caller as emd_module.py:43-54."""
    _lib.check_tensor("xyz1", xyz1, torch.float32, 3)
    _lib.check_tensor("xyz2", xyz2, torch.float32, 3)
    _lib.check_tensor("dist", dist, torch.float32)
    _lib.check_tensor("assignment", assignment, torch.int32)
    for name, t, dt in (("price", price, torch.float32), ("assignment_inv", assignment_inv, torch.int32),
                        ("bid", bid, torch.int32), ("bid_increments", bid_increments, torch.float32),
                        ("max_increments", max_increments, torch.float32)):
        if t is not None:
            _lib.check_tensor(name, t, dt)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    with torch.cuda.device(xyz1.device):
        return _lib.lib.psd_emd_forward(_lib.ptr(xyz1), _lib.ptr(xyz2), b, n, m, _lib.ptr(dist), _lib.ptr(assignment),
                                        _lib.ptr(price), _lib.ptr(assignment_inv), _lib.ptr(bid),
                                        _lib.ptr(bid_increments), _lib.ptr(max_increments), _lib.ptr(unass_idx),
                                        _lib.ptr(unass_cnt), _lib.ptr(unass_cnt_sum), _lib.ptr(cnt_tmp),
                                        _lib.ptr(max_idx), float(eps), int(iters), _lib.stream_of(xyz1))


def forward_fresh(xyz1, xyz2, dist, assignment, eps, iters) -> int:
    """psd_emd_forward_fresh: same auction from the initial state, without the 12 scratch tensors."""
    _lib.check_tensor("xyz1", xyz1, torch.float32, 3)
    _lib.check_tensor("xyz2", xyz2, torch.float32, 3)
    _lib.check_tensor("dist", dist, torch.float32)
    _lib.check_tensor("assignment", assignment, torch.int32)
    b, n, _ = xyz1.shape
    with torch.cuda.device(xyz1.device):
        return _lib.lib.psd_emd_forward_fresh(_lib.ptr(xyz1), _lib.ptr(xyz2), b, n, _lib.ptr(dist), _lib.ptr(assignment),
                                              float(eps), int(iters), _lib.stream_of(xyz1))


def backward(xyz1, xyz2, gradxyz, graddist, idx) -> int:
    """emd_backward (emd.cpp:19-23): accumulates the xyz1 gradient into the caller-zeroed gradxyz."""
    _lib.check_tensor("xyz1", xyz1, torch.float32, 3)
    _lib.check_tensor("xyz2", xyz2, torch.float32, 3)
    _lib.check_tensor("gradxyz", gradxyz, torch.float32)
    _lib.check_tensor("graddist", graddist, torch.float32)
    _lib.check_tensor("idx", idx, torch.int32)
    b, n, _ = xyz1.shape
    with torch.cuda.device(xyz1.device):
        return _lib.lib.psd_emd_backward(_lib.ptr(xyz1), _lib.ptr(xyz2), _lib.ptr(gradxyz), _lib.ptr(graddist),
                                         _lib.ptr(idx), b, n, _lib.stream_of(xyz1))
