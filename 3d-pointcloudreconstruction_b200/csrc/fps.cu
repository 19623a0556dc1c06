// fps.cu -- farthest point sampling (utils/utils.py:335-360), the step that produces the hot path's multi-resolution
// ground-truth clouds (utils/datasets_sample_pcl.py:87-91: 128 and 256 centroids per sample).
//
// The reference runs `npoint` rounds of whole-tensor torch ops per call (gather, (xyz - c)**2 summed over the last axis, masked
// min-update, argmax).  Here one CTA per cloud keeps the coordinates in shared memory and every thread's running distances in
// registers; a round is: distance update, argmax over the CTA through one shuffle tree and ONE barrier (the per-warp
// candidates are double-buffered in shared memory), next centroid read from shared memory.
// Exact arithmetic of the reference (fp32, no contraction): d = ((dx*dx + dy*dy) + dz*dz) with dx = x - cx; distance starts at
// 1e10 and is replaced only where d < distance; the next centroid is the FIRST index holding the maximum (torch.max on the
// CPU) -- here a 64-bit key (distance bits, ~index) reduced with max.
#include "psd_common.cuh"

namespace psd {

constexpr int kFpsThreads = 512;
constexpr int kFpsMaxN = 16384;    // 3 fp32 arrays of n in shared memory

template <int P>   // points per thread: n <= P * 512
__global__ void __launch_bounds__(kFpsThreads, 1) fps_kernel(const float *__restrict__ xyz, int n, int npoint, int start,
                                                             long long *__restrict__ centroids) {
    extern __shared__ float smf[];
    float *sx = smf, *sy = sx + n, *sz = sy + n;
    __shared__ unsigned long long s_key[2][kFpsThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *src = xyz + (long long)blockIdx.x * n * 3;
    for (int j = tid; j < n; j += kFpsThreads) { sx[j] = src[3 * j]; sy[j] = src[3 * j + 1]; sz[j] = src[3 * j + 2]; }
    __syncthreads();
    float dist[P];
#pragma unroll
    for (int r = 0; r < P; ++r) dist[r] = 1e10f;
    long long *out = centroids + (long long)blockIdx.x * npoint;
    int far = start;
    for (int i = 0; i < npoint; ++i) {
        if (tid == 0) out[i] = far;
        const float cx = sx[far], cy = sy[far], cz = sz[far];
        unsigned long long key = 0;
#pragma unroll
        for (int r = 0; r < P; ++r) {
            const int j = tid + r * kFpsThreads;
            if (j < n) {
                const float dx = __fsub_rn(sx[j], cx), dy = __fsub_rn(sy[j], cy), dz = __fsub_rn(sz[j], cz);
                const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
                if (d < dist[r]) dist[r] = d;
                const unsigned long long k = ((unsigned long long)__float_as_uint(dist[r]) << 32) | (0xffffffffu - (unsigned)j);
                key = k > key ? k : key;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = shfl_xor_u64(key, o);
            key = other > key ? other : key;
        }
        if (lane == 0) s_key[i & 1][warp] = key;
        __syncthreads();
        unsigned long long best = s_key[i & 1][0];
#pragma unroll
        for (int w = 1; w < kFpsThreads / 32; ++w) {
            const unsigned long long k = s_key[i & 1][w];
            best = k > best ? k : best;
        }
        far = (int)(0xffffffffu - (unsigned)(best & 0xffffffffu));
    }
}

}  // namespace psd

int psd_fps_max_points() { return psd::kFpsMaxN; }

cudaError_t psd_launch_fps(const float *xyz, int b, int n, int npoint, int start, long long *centroids, cudaStream_t stream) {
    using namespace psd;
    if (b <= 0 || npoint <= 0) return cudaSuccess;
    const int p = (n + kFpsThreads - 1) / kFpsThreads;
    void (*kern)(const float *, int, int, int, long long *) =
        p <= 1 ? fps_kernel<1> : p <= 2 ? fps_kernel<2> : p <= 4 ? fps_kernel<4> : p <= 8 ? fps_kernel<8> : p <= 16 ? fps_kernel<16> : fps_kernel<32>;
    const size_t smem = sizeof(float) * (size_t)3 * n;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<b, kFpsThreads, smem, stream>>>(xyz, n, npoint, start, centroids);
    return cudaGetLastError();
}
