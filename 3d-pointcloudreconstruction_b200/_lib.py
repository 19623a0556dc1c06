"""ctypes loader for libpsd_b200.so (the C ABI declared in include/psd_b200.h).

PyTorch is only plumbing here: tensors give device pointers, the current stream and a device guard.
The library must exist -- there is deliberately no eager/CPU fallback."""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PSD_B200_LIB") or os.path.join(_HERE, "libpsd_b200.so")  # env override: A/B builds of the same ABI

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python 3d-pointcloudreconstruction_b200/build.py` "
        "(nvcc, sm_100a). This package has no CPU or eager fallback."
    )

lib = ctypes.CDLL(LIB_PATH)

_vp = ctypes.c_void_p
_ci = ctypes.c_int
_cf = ctypes.c_float

lib.psd_version.restype = _ci
lib.psd_last_error.restype = ctypes.c_char_p
lib.psd_chamfer_forward.argtypes = [_vp, _vp, _ci, _ci, _ci, _vp, _vp, _vp, _vp, _vp]
lib.psd_chamfer_forward_ex.argtypes = [_vp, _vp, _ci, _ci, _ci, _ci, _vp, _vp, _vp, _vp, _vp, _cf, _vp, _ci, _ci, _vp]
lib.psd_chamfer_backward.argtypes = [_vp] * 8 + [_ci, _ci, _ci, _vp]
lib.psd_emd_forward.argtypes = [_vp, _vp, _ci, _ci, _ci] + [_vp] * 12 + [_cf, _ci, _vp]
lib.psd_emd_forward_fresh.argtypes = [_vp, _vp, _ci, _ci, _vp, _vp, _cf, _ci, _vp]
lib.psd_emd_forward_cluster.argtypes = [_vp, _vp, _ci, _ci, _vp, _vp, _vp, _vp, _cf, _ci, _ci, _vp]
lib.psd_emd_backward.argtypes = [_vp, _vp, _vp, _vp, _vp, _ci, _ci, _vp]
lib.psd_chamfer_forward_host.argtypes = [_vp, _vp, _ci, _ci, _ci, _vp, _vp, _vp, _vp, _vp]
lib.psd_fp32_fma_peak.argtypes = [_cf, ctypes.POINTER(_cf), _vp]
lib.psd_chamfer_stats.argtypes = [ctypes.POINTER(ctypes.c_longlong), _ci]
lib.psd_chamfer_nn_variant.argtypes = [_ci]
lib.psd_chamfer_mean_loss_forward.argtypes = [_vp, _vp, _ci, _ci, _ci, _ci, _vp, _vp, _vp, _vp, _vp, _vp, _vp]
lib.psd_chamfer_mean_loss_backward.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _ci, _ci, _ci, _vp]
lib.psd_chamfer_loss_step_host.argtypes = [_vp, _vp, _ci, _ci, _ci, _vp, _vp, _vp, _vp, _vp, _vp]
lib.psd_chamfer_loss_step_host_ex.argtypes = [_vp, _vp, _ci, _ci, _ci, _vp, _vp, _vp, _vp, _vp, _ci, _ci, _vp]
lib.psd_host_step_graphs.argtypes = [_ci]
lib.psd_emd_solo_mode.argtypes = [_ci]
lib.psd_emd_grid_mode.argtypes = [_ci]
lib.psd_chamfer_tc_ctas.argtypes = [_ci]
lib.psd_proj_min_dist.argtypes = [_vp, _vp, _vp, _ci, _ci, _ci, _ci, _vp, _vp, _vp]
lib.psd_icp_batch.argtypes = [_vp, _vp, _ci, _ci, _ci, _vp, _ci, ctypes.c_double, _vp, _vp, _vp, _vp]
lib.psd_nn_f64.argtypes = [_vp, _vp, _ci, _ci, _ci, _ci, _vp, _vp, _vp]
lib.psd_farthest_point_sample.argtypes = [_vp, _ci, _ci, _ci, _ci, _vp, _vp]
lib.psd_cont_proj.argtypes = [_vp, _ci, _ci, _ci, _ci, ctypes.c_float, _vp, _vp]
lib.psd_debug_tc_prof.argtypes = [_vp]
lib.psd_debug_tc_filter.argtypes = [_vp, _vp, _ci, _ci, _ci, _vp, _vp, _vp, _vp, _vp, _ci, _vp]
for _n in ("psd_chamfer_forward", "psd_chamfer_forward_ex", "psd_chamfer_backward", "psd_emd_forward",
           "psd_emd_forward_fresh", "psd_emd_forward_cluster", "psd_emd_backward", "psd_chamfer_forward_host",
           "psd_fp32_fma_peak", "psd_chamfer_stats", "psd_chamfer_nn_variant", "psd_debug_tc_filter", "psd_debug_tc_prof", "psd_chamfer_mean_loss_forward",
           "psd_chamfer_mean_loss_backward", "psd_chamfer_loss_step_host", "psd_chamfer_loss_step_host_ex", "psd_host_step_graphs", "psd_emd_solo_mode", "psd_emd_grid_mode", "psd_chamfer_tc_ctas", "psd_proj_min_dist", "psd_icp_batch", "psd_nn_f64", "psd_farthest_point_sample", "psd_cont_proj"):
    getattr(lib, _n).restype = _ci

EXPORTS = ("psd_version", "psd_last_error", "psd_chamfer_forward", "psd_chamfer_forward_ex", "psd_chamfer_backward",
           "psd_emd_forward", "psd_emd_forward_fresh", "psd_emd_forward_cluster", "psd_emd_backward",
           "psd_chamfer_forward_host", "psd_fp32_fma_peak", "psd_chamfer_stats", "psd_chamfer_nn_variant",
           "psd_debug_tc_filter", "psd_debug_tc_prof", "psd_chamfer_mean_loss_forward", "psd_chamfer_mean_loss_backward",
           "psd_chamfer_loss_step_host", "psd_chamfer_loss_step_host_ex", "psd_host_step_graphs", "psd_emd_solo_mode", "psd_emd_grid_mode", "psd_chamfer_tc_ctas", "psd_proj_min_dist", "psd_icp_batch", "psd_nn_f64", "psd_farthest_point_sample", "psd_cont_proj")


def last_error() -> str:
    return lib.psd_last_error().decode("utf-8", "replace")


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_of(t) -> ctypes.c_void_p:
    """torch's current stream on the tensor's device (the reference launches on the legacy default stream;
    the current stream is identical behaviour there and correct everywhere else)."""
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def check_tensor(name, t, dtype, ndim=None):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor (this op has no CPU path)")
    if t.dtype != dtype:
        raise RuntimeError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError(f"{name}: expected a contiguous tensor")
    if ndim is not None and t.dim() != ndim:
        raise RuntimeError(f"{name}: expected {ndim} dimensions, got {t.dim()}")


def raise_on_cuda_error(rc: int, what: str):
    """The reference printf()s and returns 0 on a launch error and Python ignores it
    (dist_chamfer_3D.py:52); here the 0 is still returned to the caller of the native-module mirror, but
    the autograd wrappers turn it into an exception."""
    if rc == 0:
        raise RuntimeError(f"{what} failed: {last_error()}")
