// psd_common.cuh -- shared device helpers for the sm_100a point-set-distance kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libpsd_b200 is written for sm_100a (B200) only"
#endif

namespace psd {

// Squared distance with the exact rounding sequence of the reference's SASS
// (metric/chamfer3D/chamfer3D.cu:32-35, metric/emd/emd_cuda.cu:142-146,221-224 compiled by nvcc
// without --use_fast_math): d = fma(dz,dz, fma(dx,dx, rn(dy*dy))).  Written with explicit
// intrinsics so that no compiler version can re-associate or re-contract it.
__device__ __forceinline__ float sqdist_exact(float dx, float dy, float dz) {
    return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

// Packed 2-wide FP32 FMA (SASS FFMA2, one issue slot for two FMAs; sm_100+).
__device__ __forceinline__ float2 ffma2(float a, float2 b, float2 c) {
    return __ffma2_rn(make_float2(a, a), b, c);  // ptxas folds the splat into FFMA2's scalar operand form
}

// 3-input minimum (SASS FMNMX3; sm_100+).  NaN operands are ignored like fminf.
__device__ __forceinline__ float fmin3(float a, float b, float c) {
    float d;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

__device__ __forceinline__ unsigned long long pack_key(float d, int idx) {
    return ((unsigned long long)__float_as_uint(d) << 32) | (unsigned int)idx;
}

__device__ __forceinline__ unsigned long long shfl_xor_u64(unsigned long long v, int lane_mask) {
    return __shfl_xor_sync(0xffffffffu, v, lane_mask);
}

}  // namespace psd

// host-side error plumbing (psd_capi.cu)
void psd_set_error(const char *what, cudaError_t err);
void psd_set_error_msg(const char *what);
