// psd_pybind.cpp -- the binding a maintainer of the reference adds: the reference's two pybind modules rebuilt over the
// C ABI of libpsd_b200.so (include/psd_b200.h).  Same module names, function names, argument lists and int returns as
//   metric/chamfer3D/chamfer_cuda.cpp:17-33   (module `chamfer_3D`: forward, backward)
//   metric/emd/emd.cpp:12-29                  (module `emd`: forward, backward)
// so the reference's UNMODIFIED dist_chamfer_3D.py / emd_module.py import and run on top of it.  Compiled twice by
// build.py (torch.utils.cpp_extension.load): -DPSD_BIND_CHAMFER -> chamfer_3D.so, -DPSD_BIND_EMD -> emd.so.
// Unlike the reference's launchers the calls run on torch's current stream under a device guard, and a tensor of the
// wrong dtype / device / layout raises instead of being read as raw floats.
#include <torch/extension.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>

#include "psd_b200.h"

namespace {

void *current_stream() { return (void *)c10::cuda::getCurrentCUDAStream().stream(); }

void check(const at::Tensor &t, const char *name, at::ScalarType st) {
    TORCH_CHECK(t.is_cuda(), name, ": expected a CUDA tensor");
    TORCH_CHECK(t.scalar_type() == st, name, ": wrong dtype");
    TORCH_CHECK(t.is_contiguous(), name, ": expected a contiguous tensor");
}

#ifdef PSD_BIND_CHAMFER
int chamfer_forward(at::Tensor xyz1, at::Tensor xyz2, at::Tensor dist1, at::Tensor dist2, at::Tensor idx1, at::Tensor idx2) {
    check(xyz1, "xyz1", at::kFloat); check(xyz2, "xyz2", at::kFloat);
    check(dist1, "dist1", at::kFloat); check(dist2, "dist2", at::kFloat);
    check(idx1, "idx1", at::kInt); check(idx2, "idx2", at::kInt);
    TORCH_CHECK(xyz1.dim() == 3 && xyz2.dim() == 3 && xyz1.size(2) == 3 && xyz2.size(2) == 3 && xyz1.size(0) == xyz2.size(0),
                "chamfer_3D.forward: expected xyz1 [B,N,3] and xyz2 [B,M,3]");
    c10::cuda::CUDAGuard guard(xyz1.device());
    return psd_chamfer_forward(xyz1.data_ptr<float>(), xyz2.data_ptr<float>(), (int)xyz1.size(0), (int)xyz1.size(1),
                               (int)xyz2.size(1), dist1.data_ptr<float>(), dist2.data_ptr<float>(), idx1.data_ptr<int>(),
                               idx2.data_ptr<int>(), current_stream());
}

int chamfer_backward(at::Tensor xyz1, at::Tensor xyz2, at::Tensor gradxyz1, at::Tensor gradxyz2, at::Tensor graddist1,
                     at::Tensor graddist2, at::Tensor idx1, at::Tensor idx2) {
    check(xyz1, "xyz1", at::kFloat); check(xyz2, "xyz2", at::kFloat);
    check(gradxyz1, "gradxyz1", at::kFloat); check(gradxyz2, "gradxyz2", at::kFloat);
    check(graddist1, "graddist1", at::kFloat); check(graddist2, "graddist2", at::kFloat);
    check(idx1, "idx1", at::kInt); check(idx2, "idx2", at::kInt);
    c10::cuda::CUDAGuard guard(xyz1.device());
    return psd_chamfer_backward(xyz1.data_ptr<float>(), xyz2.data_ptr<float>(), gradxyz1.data_ptr<float>(),
                                gradxyz2.data_ptr<float>(), graddist1.data_ptr<float>(), graddist2.data_ptr<float>(),
                                idx1.data_ptr<int>(), idx2.data_ptr<int>(), (int)xyz1.size(0), (int)xyz1.size(1),
                                (int)xyz2.size(1), current_stream());
}
#endif

#ifdef PSD_BIND_EMD
int emd_forward(at::Tensor xyz1, at::Tensor xyz2, at::Tensor dist, at::Tensor assignment, at::Tensor price,
                at::Tensor assignment_inv, at::Tensor bid, at::Tensor bid_increments, at::Tensor max_increments,
                at::Tensor unass_idx, at::Tensor unass_cnt, at::Tensor unass_cnt_sum, at::Tensor cnt_tmp, at::Tensor max_idx,
                float eps, int iters) {
    check(xyz1, "xyz1", at::kFloat); check(xyz2, "xyz2", at::kFloat);
    check(dist, "dist", at::kFloat); check(assignment, "assignment", at::kInt);
    check(price, "price", at::kFloat); check(assignment_inv, "assignment_inv", at::kInt);
    check(bid, "bid", at::kInt); check(bid_increments, "bid_increments", at::kFloat);
    check(max_increments, "max_increments", at::kFloat);
    c10::cuda::CUDAGuard guard(xyz1.device());
    return psd_emd_forward(xyz1.data_ptr<float>(), xyz2.data_ptr<float>(), (int)xyz1.size(0), (int)xyz1.size(1),
                           (int)xyz2.size(1), dist.data_ptr<float>(), assignment.data_ptr<int>(), price.data_ptr<float>(),
                           assignment_inv.data_ptr<int>(), bid.data_ptr<int>(), bid_increments.data_ptr<float>(),
                           max_increments.data_ptr<float>(), unass_idx.data_ptr<int>(), unass_cnt.data_ptr<int>(),
                           unass_cnt_sum.data_ptr<int>(), cnt_tmp.data_ptr<int>(), max_idx.data_ptr<int>(), eps, iters,
                           current_stream());
}

int emd_backward(at::Tensor xyz1, at::Tensor xyz2, at::Tensor gradxyz, at::Tensor graddist, at::Tensor idx) {
    check(xyz1, "xyz1", at::kFloat); check(xyz2, "xyz2", at::kFloat);
    check(gradxyz, "gradxyz", at::kFloat); check(graddist, "graddist", at::kFloat); check(idx, "idx", at::kInt);
    c10::cuda::CUDAGuard guard(xyz1.device());
    return psd_emd_backward(xyz1.data_ptr<float>(), xyz2.data_ptr<float>(), gradxyz.data_ptr<float>(),
                            graddist.data_ptr<float>(), idx.data_ptr<int>(), (int)xyz1.size(0), (int)xyz1.size(1),
                            current_stream());
}
#endif

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
#ifdef PSD_BIND_CHAMFER
    m.def("forward", &chamfer_forward, "chamfer forward (CUDA, libpsd_b200)");
    m.def("backward", &chamfer_backward, "chamfer backward (CUDA, libpsd_b200)");
#endif
#ifdef PSD_BIND_EMD
    m.def("forward", &emd_forward, "emd forward (CUDA, libpsd_b200)");
    m.def("backward", &emd_backward, "emd backward (CUDA, libpsd_b200)");
#endif
}
