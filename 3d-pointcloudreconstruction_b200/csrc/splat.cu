// splat.cu -- cont_proj, the continuous orthographic projection of a cloud to a silhouette image
// (utils/projection.py:4-67 with apply_kernel :95-106), the producer of proj_loss's inputs (utils/utils.py:232,241).
//
//     x = ((p.x + 1) * grid_h) / 2,  y = ((p.y + 1) * grid_w) / 2                                   (:20-21)
//     val[b,h,w] = sum_p exp(-((x_p - h)^2) / (2 sigma^2)) * exp(-((y_p - w)^2) / (2 sigma^2))       (:61-64)
// The reference materialises the [B, N, H, W, 2] difference tensor on the CPU (1 GB at B=32, N=1024, 64x64).  The kernel is
// separable: per point one row of H and one row of W exponentials, then a rank-1 update of the image -- a small
// [H x N] x [N x W] product per sample, accumulated in point order like torch.sum over dim 1 (fp32, product rounded before the
// add).  One CTA = (sample, 64x64 output tile); 256 threads with 4x4 register tiles; the exponentials of 64 points at a time
// are staged in shared memory.
#include "psd_common.cuh"

namespace psd {

constexpr int kSplatChunk = 64;

__global__ void __launch_bounds__(256) cont_proj_kernel(const float *__restrict__ pcl, int n, int grid_h, int grid_w,
                                                        float two_sigma_sq, float *__restrict__ out) {
    __shared__ __align__(16) float s_ex[kSplatChunk][64];
    __shared__ __align__(16) float s_ey[kSplatChunk][64];
    const int tiles_w = (grid_w + 63) / 64, tiles_h = (grid_h + 63) / 64;
    const int sample = blockIdx.x / (tiles_h * tiles_w);
    const int tile = blockIdx.x - sample * tiles_h * tiles_w;
    const int h0 = (tile / tiles_w) * 64, w0 = (tile % tiles_w) * 64;
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const float *src = pcl + (long long)sample * n * 3;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int p0 = 0; p0 < n; p0 += kSplatChunk) {
        const int np = min(kSplatChunk, n - p0);
        __syncthreads();
        // 2 * 64 * 64 exponentials per chunk: thread -> (point, 16 consecutive cells of one axis)
        for (int e = tid; e < kSplatChunk * 8; e += 256) {
            const int pp = e >> 3, part = e & 7;          // part 0..3: x cells, 4..7: y cells
            const bool is_x = part < 4;
            const int c0 = (part & 3) * 16;
            float *dst = is_x ? &s_ex[pp][c0] : &s_ey[pp][c0];
            if (pp < np) {
                const float v = src[3 * (p0 + pp) + (is_x ? 0 : 1)];
                const float g = (float)(is_x ? grid_h : grid_w);
                const float pos = __fdiv_rn(__fmul_rn(__fadd_rn(v, 1.0f), g), 2.0f);
                const int base = (is_x ? h0 : w0) + c0;
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    const float d = __fsub_rn(pos, (float)(base + c));
                    dst[c] = expf(__fdiv_rn(-__fmul_rn(d, d), two_sigma_sq));
                }
            } else {
#pragma unroll
                for (int c = 0; c < 16; ++c) dst[c] = 0.f;
            }
        }
        __syncthreads();
        for (int pp = 0; pp < np; ++pp) {
            const float4 ex = *reinterpret_cast<const float4 *>(&s_ex[pp][4 * ty]);
            const float4 ey = *reinterpret_cast<const float4 *>(&s_ey[pp][4 * tx]);
            const float exv[4] = {ex.x, ex.y, ex.z, ex.w}, eyv[4] = {ey.x, ey.y, ey.z, ey.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = __fadd_rn(acc[i][j], __fmul_rn(exv[i], eyv[j]));
        }
    }
    float *o = out + (long long)sample * grid_h * grid_w;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int h = h0 + 4 * ty + i;
        if (h >= grid_h) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int w = w0 + 4 * tx + j;
            if (w < grid_w) o[(long long)h * grid_w + w] = acc[i][j];
        }
    }
}

// Backward of cont_proj (the reference's op is plain differentiable torch, utils/projection.py:4-67: autograd carries a
// gradient from the silhouette back to the cloud's x and y; z does not enter).  With ex_p[h] = exp(-(x_p-h)^2 / 2s),
// ey_p[w] likewise, x_p = ((p.x+1) H)/2 and G = d loss / d out:
//     d loss / d p.x = (H/2) sum_h (-(x_p-h)/s) ex_p[h] sum_w G[h,w] ey_p[w]
//     d loss / d p.y = (W/2) sum_w (-(y_p-w)/s) ey_p[w] sum_h G[h,w] ex_p[h],      d loss / d p.z = 0.
// One CTA = (sample, slice of its points) with the sample's G staged once in shared memory (row stride W+1: the row-owned pass
// reads columns conflict-free); one warp per point: lanes own rows in the first pass and columns in the second.
__global__ void __launch_bounds__(256) cont_proj_grad_kernel(const float *__restrict__ pcl, const float *__restrict__ gout,
                                                             int n, int grid_h, int grid_w, float sigma_sq, int pts_per_cta,
                                                             float *__restrict__ gpcl) {
    extern __shared__ __align__(16) float sm[];
    const int ld = grid_w + 1;
    float *sG = sm;                                   // [grid_h][ld]
    float *sE = sG + grid_h * ld;                     // per warp: ex[grid_h], ey[grid_w]
    const int sample = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *g = gout + (long long)sample * grid_h * grid_w;
    for (int e = tid; e < grid_h * grid_w; e += 256) {
        const int h = e / grid_w, w = e - h * grid_w;
        sG[h * ld + w] = g[e];
    }
    __syncthreads();
    float *ex = sE + warp * (grid_h + grid_w), *ey = ex + grid_h;
    const float two_sigma_sq = (float)(2.0 * (double)sigma_sq);
    const int p_begin = blockIdx.x * pts_per_cta, p_end = min(n, p_begin + pts_per_cta);
    for (int pt = p_begin + warp; pt < p_end; pt += 8) {
        const float *src = pcl + ((long long)sample * n + pt) * 3;
        const float px = __fdiv_rn(__fmul_rn(__fadd_rn(src[0], 1.0f), (float)grid_h), 2.0f);
        const float py = __fdiv_rn(__fmul_rn(__fadd_rn(src[1], 1.0f), (float)grid_w), 2.0f);
        __syncwarp();
        for (int h = lane; h < grid_h; h += 32) { const float d = px - (float)h; ex[h] = expf(__fdiv_rn(-(d * d), two_sigma_sq)); }
        for (int w = lane; w < grid_w; w += 32) { const float d = py - (float)w; ey[w] = expf(__fdiv_rn(-(d * d), two_sigma_sq)); }
        __syncwarp();
        float gx = 0.f, gy = 0.f;
        for (int h = lane; h < grid_h; h += 32) {        // rows: t[h] = sum_w G[h,w] ey[w]
            const float *row = sG + h * ld;
            float t = 0.f;
            for (int w = 0; w < grid_w; ++w) t = fmaf(row[w], ey[w], t);
            gx = fmaf(-(px - (float)h) / sigma_sq * ex[h], t, gx);
        }
        for (int w = lane; w < grid_w; w += 32) {        // columns: s[w] = sum_h G[h,w] ex[h]
            float t = 0.f;
            for (int h = 0; h < grid_h; ++h) t = fmaf(sG[h * ld + w], ex[h], t);
            gy = fmaf(-(py - (float)w) / sigma_sq * ey[w], t, gy);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { gx += __shfl_xor_sync(0xffffffffu, gx, o); gy += __shfl_xor_sync(0xffffffffu, gy, o); }
        if (lane == 0) {
            float *o = gpcl + ((long long)sample * n + pt) * 3;
            o[0] = gx * ((float)grid_h * 0.5f);
            o[1] = gy * ((float)grid_w * 0.5f);
            o[2] = 0.f;
        }
    }
}

}  // namespace psd

cudaError_t psd_launch_cont_proj(const float *pcl, int b, int n, int grid_h, int grid_w, float sigma_sq, float *out,
                                 cudaStream_t stream) {
    using namespace psd;
    if (b <= 0 || grid_h <= 0 || grid_w <= 0) return cudaSuccess;
    const int tiles = ((grid_h + 63) / 64) * ((grid_w + 63) / 64);
    // `2.*sigma_sq` is a Python double that torch casts to the tensor's float32 for the division
    const float two_sigma_sq = (float)(2.0 * (double)sigma_sq);
    cont_proj_kernel<<<b * tiles, 256, 0, stream>>>(pcl, n, grid_h, grid_w, two_sigma_sq, out);
    return cudaGetLastError();
}

// gpcl [B,N,3] = d loss / d pcl for gout [B,grid_h,grid_w] = d loss / d cont_proj(pcl); returns cudaErrorInvalidValue when the
// gradient image does not fit in shared memory (grid_h * (grid_w + 1) floats + 8 (grid_h + grid_w) <= ~50 k floats)
cudaError_t psd_launch_cont_proj_backward(const float *pcl, const float *gout, int b, int n, int grid_h, int grid_w,
                                          float sigma_sq, float *gpcl, cudaStream_t stream) {
    using namespace psd;
    if (b <= 0 || n <= 0 || grid_h <= 0 || grid_w <= 0) return cudaSuccess;
    const size_t smem = sizeof(float) * ((size_t)grid_h * (grid_w + 1) + 8 * (size_t)(grid_h + grid_w));
    if (smem > 200 * 1024 || b > 65535) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(cont_proj_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int chunks = (296 + b - 1) / b;                      // about two waves of CTAs
    if (chunks > (n + 7) / 8) chunks = (n + 7) / 8;
    if (chunks < 1) chunks = 1;
    const int pts = (n + chunks - 1) / chunks;
    chunks = (n + pts - 1) / pts;
    cont_proj_grad_kernel<<<dim3(chunks, b), 256, smem, stream>>>(pcl, gout, n, grid_h, grid_w, sigma_sq, pts, gpcl);
    return cudaGetLastError();
}
