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

// Programmatic dependent launch (PSD_PDL, default on): the library's stream-ordered kernel pairs (NN forward -> gradient -> next
// forward) are launched with cudaLaunchAttributeProgrammaticStreamSerialization.  A kernel signals near its END that its
// successor may be launched (pdl_trigger: the successor's launch latency and set-up overlap this kernel's tail and drain) and
// waits for its predecessor's completion and memory (pdl_wait) before its FIRST global access: same results as plain stream
// order, whatever the neighbouring kernels are (a predecessor that never triggers does so implicitly when it completes).
#ifndef PSD_PDL
#define PSD_PDL 1
#endif
__device__ __forceinline__ void pdl_trigger() {
    if (PSD_PDL) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void pdl_wait() {
    if (PSD_PDL) asm volatile("griddepcontrol.wait;" ::: "memory");
}

__device__ __forceinline__ unsigned long long pack_key(float d, int idx) {
    return ((unsigned long long)__float_as_uint(d) << 32) | (unsigned int)idx;
}

__device__ __forceinline__ unsigned long long shfl_xor_u64(unsigned long long v, int lane_mask) {
    return __shfl_xor_sync(0xffffffffu, v, lane_mask);
}

}  // namespace psd

// host side of PSD_PDL: a launch whose kernel calls pdl_wait() before its first global access
template <typename... KArgs, typename... Args>
static inline cudaError_t psd_launch_pdl(bool allow, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = (PSD_PDL && allow) ? 1 : 0;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// host-side error plumbing (psd_capi.cu)
void psd_set_error(const char *what, cudaError_t err);
void psd_set_error_msg(const char *what);
