// proj.cu -- the min-distance terms of the 2-D projection loss (loss/proj_loss.py:21-40, grid_dist :46-54).
//
// The reference materialises [B,H,W,H,W] tensors on the CPU:
//     D1[h,w,h',w']       = float32(cdist(grid, grid)) + 1                       (proj_loss.py:22, utils/utils.py:224)
//     dist_masked         = gt_th[b,h,w]  * D1 * pred[b,h,w],   gt_th     = gt   + (1-gt)  *1e6      (:37-38)
//     dist_masked_inv     = gt[b,h,w]     * D1 * pred_mask[b,h,w], pred_mask = pred + (1-pred)*1e6    (:33-34)
//     min_dist = min over (h',w'),  min_dist_inv likewise                                                (:40-41)
// As written the weights do not depend on (h',w') (both images are broadcast along the FIRST pixel pair), so
// the minimum sits at an end point of D1's range for that pixel.  mode 0 ("as written") evaluates exactly that, bit for
// bit: fl(fl(a*D1)*c) is monotone in D1, so min over (h',w') = min over {D1 = 1 (the pixel itself), D1 = max over the four
// corners}.  mode 1 ("intended", the CAPNet-style loss the code was taken from) puts the mask on the TARGET pixel,
//     min_dist[b,p]     = min_p' fl(fl(gt_th[b,p']  * D1[p,p']) * pred[b,p])
//     min_dist_inv[b,p] = min_p' fl(fl(pred_mask[b,p'] * D1[p,p']) * gt[b,p]),
// a weighted nearest-on-pixel search, here a shared-memory brute force: D1 is a (|dh|,|dw|) table of H*W floats, the
// weight image of one sample sits next to it, every thread owns one query pixel.
#include "psd_common.cuh"

namespace psd {

// table[dh*w + dw] = float32(sqrt(dh^2+dw^2)) + offset   (offset = 1 per `dist_mat += 1`), built on the host in the caller
__global__ void __launch_bounds__(256) proj_min_dist_as_written_kernel(const float *__restrict__ pred, const float *__restrict__ gt,
                                                                       const float *__restrict__ table, int b, int h, int w,
                                                                       float *__restrict__ out_min, float *__restrict__ out_inv) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)b * h * w) return;
    const int p = (int)(i % (h * w));
    const int ph = p / w, pw = p - ph * w;
    const float pr = pred[i], g = gt[i];
    const float gth = __fadd_rn(g, __fmul_rn(__fsub_rn(1.0f, g), 1e6f));     // gt + (1-gt)*1e6*1
    const float pm = __fadd_rn(pr, __fmul_rn(__fsub_rn(1.0f, pr), 1e6f));   // pred + (1-pred)*1e6*1
    const float dmin = table[0];
    const int fh = max(ph, h - 1 - ph), fw = max(pw, w - 1 - pw);
    const float dmax = table[fh * w + fw];
    const float a0 = __fmul_rn(__fmul_rn(gth, dmin), pr), a1 = __fmul_rn(__fmul_rn(gth, dmax), pr);
    const float c0 = __fmul_rn(__fmul_rn(g, dmin), pm), c1 = __fmul_rn(__fmul_rn(g, dmax), pm);
    // torch.min propagates NaN
    out_min[i] = (a0 != a0 || a1 != a1) ? __int_as_float(0x7fc00000) : fminf(a0, a1);
    out_inv[i] = (c0 != c0 || c1 != c1) ? __int_as_float(0x7fc00000) : fminf(c0, c1);
}

// one CTA = (sample, output kind, block of 256 query pixels); dynamic smem: table[h*w] + weights[h*w]
__global__ void __launch_bounds__(256) proj_min_dist_intended_kernel(const float *__restrict__ pred, const float *__restrict__ gt,
                                                                     const float *__restrict__ table, int b, int h, int w,
                                                                     float *__restrict__ out_min, float *__restrict__ out_inv) {
    extern __shared__ float sm[];
    const int hw = h * w;
    float *s_tab = sm, *s_wgt = sm + hw;
    const int qblocks = (hw + 255) / 256;
    const int sample = blockIdx.x / (2 * qblocks);
    const int rem = blockIdx.x - sample * 2 * qblocks;
    const int kind = rem / qblocks;                  // 0: min_dist (weights gt_th, factor pred), 1: min_dist_inv
    const int q = (rem - kind * qblocks) * 256 + threadIdx.x;
    const float *wsrc = (kind == 0 ? gt : pred) + (long long)sample * hw;
    const float *csrc = (kind == 0 ? pred : gt) + (long long)sample * hw;
    for (int i = threadIdx.x; i < hw; i += 256) {
        s_tab[i] = table[i];
        const float v = wsrc[i];
        s_wgt[i] = __fadd_rn(v, __fmul_rn(__fsub_rn(1.0f, v), 1e6f));   // v + (1-v)*1e6*1
    }
    __syncthreads();
    if (q >= hw) return;
    const int qh = q / w, qw = q - qh * w;
    const float c = csrc[q];
    float best = __int_as_float(0x7f800000);
    bool nan = false;
    for (int th = 0; th < h; ++th) {
        const float *trow = s_tab + abs(th - qh) * w;
        const float *wrow = s_wgt + th * w;
#pragma unroll 4
        for (int tw = 0; tw < w; ++tw) {
            const float v = __fmul_rn(__fmul_rn(wrow[tw], trow[abs(tw - qw)]), c);
            nan |= v != v;
            best = fminf(best, v);
        }
    }
    (kind == 0 ? out_min : out_inv)[(long long)sample * hw + q] = nan ? __int_as_float(0x7fc00000) : best;
}

}  // namespace psd

using namespace psd;

cudaError_t psd_launch_proj_min_dist(const float *pred, const float *gt, const float *table, int b, int h, int w, int mode,
                                     float *out_min, float *out_inv, cudaStream_t stream) {
    if (b <= 0 || h <= 0 || w <= 0) return cudaSuccess;
    const long long total = (long long)b * h * w;
    if (mode == 0) {
        proj_min_dist_as_written_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(pred, gt, table, b, h, w, out_min, out_inv);
    } else {
        const int hw = h * w;
        const size_t smem = 2 * sizeof(float) * (size_t)hw;
        if (smem > 200 * 1024) return cudaErrorInvalidValue;
        // the opt-in applies per device: set it on every launch (cheap next to this kernel; emd.cu and fps.cu do the same)
        cudaError_t e = cudaFuncSetAttribute(proj_min_dist_intended_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return e;
        const int qblocks = (hw + 255) / 256;
        proj_min_dist_intended_kernel<<<(unsigned)(b * 2 * qblocks), 256, smem, stream>>>(pred, gt, table, b, h, w, out_min, out_inv);
    }
    return cudaGetLastError();
}
