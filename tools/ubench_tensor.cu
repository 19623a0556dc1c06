// Tensor-path microbenchmarks for sm_100a that size a tensor-core NN filter (a_k = w_k - 2 q.t_k as a K=16 GEMM):
//   1. legacy mma.sync rates: m16n8k8 tf32 and m16n8k16 bf16 (HMMA), MAC/clk/SM at 4/8/16 warps per SM
//   2. the realistic mix: per 16x8 result block 1 or 2 mma.sync + 2 FMNMX3 (running min per query), B operands from LDS
//   3. tcgen05.ld (LDTM) throughput: TMEM -> registers, bytes/clk/SM at 4/8/16 warps, with and without a min consumer
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_tensor ubench_tensor.cu
// (tmem_ld_gen.cuh is generated: one asm wrapper per tcgen05.ld.32x32b.xN shape)
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#include "tmem_ld_gen.cuh"
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float min3(float a, float b, float c) {
    float r;
    asm volatile("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

// ---- 1. pure mma.sync rate: 8 independent accumulators per warp -------------------------------------------------
template <int KIND>   // 0 tf32 k8, 1 bf16 k16
__global__ void __launch_bounds__(512, 1) mma_rate(float *out, long long *cycles, int iters) {
    float c[8][4];
    uint32_t a[4], b[2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) a[j] = 0x3f800000u + threadIdx.x * 8 + j;
    b[0] = 0x3f000000u + threadIdx.x; b[1] = 0x3e800000u + threadIdx.x;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (KIND == 0) mma_tf32(c[i], a, b[0], b[1]);
                else mma_bf16(c[i], a, b[0], b[1]);
            }
    }
    __syncthreads();
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// ---- 2. filter mix: 64 queries per warp (4 m16 blocks, A resident), NB target blocks of 8 from LDS ---------------
template <int KIND>   // 0: 2 x tf32 k8 (K=16)   1: 2 x bf16 k16 (K=32)   2: 1 x bf16 k16 (K=16)   3: 1 x tf32 k8 (K=8)
__global__ void __launch_bounds__(512, 1) mma_min_mix(float *out, long long *cycles, int iters) {
    constexpr int NB = 64;    // 512 targets per pass
    __shared__ float4 sB[NB][32];
    for (int i = threadIdx.x; i < NB * 32; i += blockDim.x)
        sB[i / 32][i % 32] = make_float4(1e-3f * i, 2e-3f * i, 3e-3f * i, 4e-3f * i);
    uint32_t a[4][2][4];
    float best[8];
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int j = 0; j < 4; ++j) a[m][h][j] = 0x3f800000u + threadIdx.x * 64 + m * 8 + h * 4 + j;
#pragma unroll
    for (int i = 0; i < 8; ++i) best[i] = 1e30f;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll 8
        for (int j = 0; j < NB; ++j) {
            const float4 bv = sB[j][lane];
            const uint32_t b0 = __float_as_uint(bv.x), b1 = __float_as_uint(bv.y), b2 = __float_as_uint(bv.z),
                           b3 = __float_as_uint(bv.w);
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                float c[4] = {0.f, 0.f, 0.f, 0.f};
                if (KIND == 0) { mma_tf32(c, a[m][0], b0, b1); mma_tf32(c, a[m][1], b2, b3); }
                if (KIND == 1) { mma_bf16(c, a[m][0], b0, b1); mma_bf16(c, a[m][1], b2, b3); }
                if (KIND == 2) { mma_bf16(c, a[m][0], b0, b1); }
                if (KIND == 3) { mma_tf32(c, a[m][0], b0, b1); }
                best[2 * m] = min3(best[2 * m], c[0], c[1]);
                best[2 * m + 1] = min3(best[2 * m + 1], c[2], c[3]);
            }
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += best[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// ---- 3. tcgen05.ld throughput --------------------------------------------------------------------------------------
template <int X, bool CONSUME>
__global__ void __launch_bounds__(X == 64 ? 256 : 512, 1) tmem_ld_rate(float *out, long long *cycles, int iters) {
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&tmem_base_s);
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t base = tmem_base_s + ((uint32_t)((warp & 3) * 32) << 16);
    float best[4] = {1e30f, 1e30f, 1e30f, 1e30f};
    uint32_t r0[X], r1[X];
#pragma unroll
    for (int i = 0; i < X; ++i) { r0[i] = 0x7f000000u; r1[i] = 0x7f000000u; }
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        for (int col = 0; col < 512; col += 2 * X) {
            if constexpr (X == 32) tmem_ld_32x32b_x32(base + col, r0);
            if constexpr (X == 64) tmem_ld_32x32b_x64(base + col, r0);
            if (CONSUME) {
#pragma unroll
                for (int i = 0; i < X; i += 2)
                    best[(i >> 1) & 3] = min3(best[(i >> 1) & 3], __uint_as_float(r1[i]), __uint_as_float(r1[i + 1]));
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if constexpr (X == 32) tmem_ld_32x32b_x32(base + col + X, r1);
            if constexpr (X == 64) tmem_ld_32x32b_x64(base + col + X, r1);
            if (CONSUME) {
#pragma unroll
                for (int i = 0; i < X; i += 2)
                    best[(i >> 1) & 3] = min3(best[(i >> 1) & 3], __uint_as_float(r0[i]), __uint_as_float(r0[i + 1]));
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    float s = best[0] + best[1] + best[2] + best[3];
    if (!CONSUME) s += __uint_as_float(r0[0]) + __uint_as_float(r1[X - 1]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_s), "r"(512u));
}

static float *d_out;
static long long *d_cyc;
static long long h_cyc[148];

template <typename F>
static double run(F launch, int threads, const char *name, double units_per_cta, const char *unit) {
    launch();                         // warm-up
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    launch();
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    CK(cudaMemcpy(h_cyc, d_cyc, sizeof(h_cyc), cudaMemcpyDeviceToHost));
    double mean = 0;
    for (int i = 0; i < 148; ++i) mean += h_cyc[i];
    mean /= 148;
    printf("%-58s threads=%4d  %10.1f %s/clk/SM   (%.3f ms, %.0f cycles -> %.2f GHz)\n", name, threads,
           units_per_cta / mean, unit, ms, mean, mean / (ms * 1e6));
    fflush(stdout);
    return units_per_cta / mean;
}

int main() {
    CK(cudaMalloc(&d_out, 148 * 512 * sizeof(float)));
    CK(cudaMalloc(&d_cyc, 148 * sizeof(long long)));
    const int iters = 2000;
    for (int threads : {128, 256, 512}) {
        const double w = threads / 32;
        run([&] { mma_rate<0><<<148, threads>>>(d_out, d_cyc, iters); }, threads, "mma.sync m16n8k8 tf32", w * iters * 32.0 * 16 * 8 * 8, "MAC");
        run([&] { mma_rate<1><<<148, threads>>>(d_out, d_cyc, iters); }, threads, "mma.sync m16n8k16 bf16", w * iters * 32.0 * 16 * 8 * 16, "MAC");
    }
    const int it2 = 200;
    for (int threads : {128, 256, 512}) {
        const double w = threads / 32;
        const double pairs = w * it2 * 64.0 * 64 * 8;   // NB blocks x 64 queries x 8 targets
        run([&] { mma_min_mix<0><<<148, threads>>>(d_out, d_cyc, it2); }, threads, "filter mix: 2 x tf32 k8 + 2 FMNMX3 per 16x8 block (K=16)", pairs, "pairs");
        run([&] { mma_min_mix<1><<<148, threads>>>(d_out, d_cyc, it2); }, threads, "filter mix: 2 x bf16 k16 + 2 FMNMX3 per 16x8 block (K=32)", pairs, "pairs");
        run([&] { mma_min_mix<2><<<148, threads>>>(d_out, d_cyc, it2); }, threads, "filter mix: 1 x bf16 k16 + 2 FMNMX3 per 16x8 block (K=16)", pairs, "pairs");
        run([&] { mma_min_mix<3><<<148, threads>>>(d_out, d_cyc, it2); }, threads, "filter mix: 1 x tf32 k8 + 2 FMNMX3 per 16x8 block (K=8)", pairs, "pairs");
    }
    const int it3 = 400;
    for (int threads : {128, 256, 512}) {
        const double w = threads / 32;
        const double bytes = w * it3 * 512.0 * 32 * 4;
        run([&] { tmem_ld_rate<32, false><<<148, threads>>>(d_out, d_cyc, it3); }, threads, "tcgen05.ld 32x32b.x32, no consumer", bytes, "B");
        run([&] { tmem_ld_rate<32, true><<<148, threads>>>(d_out, d_cyc, it3); }, threads, "tcgen05.ld 32x32b.x32 + FMNMX3 per 2 values", bytes, "B");
        if (threads <= 256) {
            run([&] { tmem_ld_rate<64, false><<<148, threads>>>(d_out, d_cyc, it3); }, threads, "tcgen05.ld 32x32b.x64, no consumer", bytes, "B");
            run([&] { tmem_ld_rate<64, true><<<148, threads>>>(d_out, d_cyc, it3); }, threads, "tcgen05.ld 32x32b.x64 + FMNMX3 per 2 values", bytes, "B");
        }
    }
    printf("done\n");
    return 0;
}
