// Scanner-side pipeline microbenchmark on sm_100a (csrc/chamfer_nn_tc.cu's steady state without the helpers): one MMA thread issues
// kind::f16 M128 N256 K16 MMAs into two 256-column TMEM buffers, 8 scanner warps (lane quarter r, column half c) reduce their 128
// columns of every tile with the kernel's min32 bookkeeping.  Cycles per tile for different scanner organisations:
//   0  the kernel's order: load 64, reduce, load 64, release, reduce
//   1  two 64-column landing zones: all 128 columns of a tile are loaded up front, the buffer goes back BEFORE any reduction, and the
//      loads of the next tile are issued between the two reductions of this one (never-taken branches pin the order against ptxas)
//   2  as 1 without the scheduling fences
//   3  as 0, but the full barrier of the next tile is probed (test_wait) before the last reduction
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o ubench_scan ubench_scan.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ uint32_t try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok;
}
__device__ __forceinline__ uint32_t test_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{ .reg .pred p; mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    for (long long spins = 0; spins < 40000000LL; ++spins) if (try_wait(bar, parity)) return;
}
__device__ __forceinline__ float min3(float a, float b, float c) { float r; asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
#define R32(a) "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7]), "=r"(a[8]), "=r"(a[9]), "=r"(a[10]), "=r"(a[11]), "=r"(a[12]), "=r"(a[13]), "=r"(a[14]), "=r"(a[15]), "=r"(a[16]), "=r"(a[17]), "=r"(a[18]), "=r"(a[19]), "=r"(a[20]), "=r"(a[21]), "=r"(a[22]), "=r"(a[23]), "=r"(a[24]), "=r"(a[25]), "=r"(a[26]), "=r"(a[27]), "=r"(a[28]), "=r"(a[29]), "=r"(a[30]), "=r"(a[31])
#define RW32(a) "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]), "+r"(a[8]), "+r"(a[9]), "+r"(a[10]), "+r"(a[11]), "+r"(a[12]), "+r"(a[13]), "+r"(a[14]), "+r"(a[15]), "+r"(a[16]), "+r"(a[17]), "+r"(a[18]), "+r"(a[19]), "+r"(a[20]), "+r"(a[21]), "+r"(a[22]), "+r"(a[23]), "+r"(a[24]), "+r"(a[25]), "+r"(a[26]), "+r"(a[27]), "+r"(a[28]), "+r"(a[29]), "+r"(a[30]), "+r"(a[31])
#define LDTXT "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&a)[32]) { asm volatile(LDTXT : R32(a) : "r"(taddr) : "memory"); }
__device__ __forceinline__ void ldwait2(uint32_t (&a)[32], uint32_t (&b)[32]) { asm volatile("tcgen05.wait::ld.sync.aligned;" : RW32(a), RW32(b) :: "memory"); }
__device__ __forceinline__ void ldwait4(uint32_t (&a)[32], uint32_t (&b)[32], uint32_t (&c)[32], uint32_t (&d)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" : RW32(a), RW32(b) :: "memory");
    asm volatile("" : RW32(c), RW32(d) :: "memory");
}
__device__ __forceinline__ float min32(const uint32_t (&r)[32]) {
    float m[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        m[i] = min3(__uint_as_float(r[8 * i]), __uint_as_float(r[8 * i + 1]), __uint_as_float(r[8 * i + 2]));
        m[i] = min3(m[i], __uint_as_float(r[8 * i + 3]), __uint_as_float(r[8 * i + 4]));
        m[i] = min3(m[i], __uint_as_float(r[8 * i + 5]), __uint_as_float(r[8 * i + 6]));
    }
    float v = min3(m[0], m[1], m[2]);
    v = min3(v, m[3], __uint_as_float(r[7]));
    v = min3(v, __uint_as_float(r[15]), __uint_as_float(r[23]));
    return fminf(v, __uint_as_float(r[31]));
}

__global__ void __launch_bounds__(288, 1) scan(int org, int ntiles, long long *cycles, float *out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bars[4];
    __shared__ uint32_t tmem_s;
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar_full = smem_u32(&bars[0]), bar_empty = smem_u32(&bars[2]);
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_s)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_s;
    const long long t0 = clock64();
    if (warp == 8) {
        if (lane == 0) {
            const uint64_t dbase = ((uint64_t)(128u >> 4) << 16) | ((uint64_t)(256u >> 4) << 32) | (1ull << 46);
            const uint64_t ad = dbase | ((smem_u32(smem) & 0x3ffffu) >> 4);
            const uint64_t bd = dbase | (((smem_u32(smem) + 8192u) & 0x3ffffu) >> 4);
            const uint32_t idesc2 = (1u << 4) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
            for (int g = 0; g < ntiles; ++g) {
                const int b = g & 1;
                mbar_wait(bar_empty + 8 * b, ((g >> 1) & 1) ^ 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p; }" ::"r"(tmem + b * 256), "l"(ad), "l"(bd + (uint64_t)((g & 7) * 512)), "r"(idesc2), "r"(0u) : "memory");
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_full + 8 * b) : "memory");
            }
        }
        __syncwarp();
    } else {
        const int r = warp & 3, c = warp >> 2;
        const uint32_t tl = tmem + ((uint32_t)(r * 32) << 16) + (uint32_t)(c * 128);
        float best = 1e30f, second = 1e30f;
        int bchunk = 0;
        auto book = [&](const uint32_t (&v)[32], int cid) {
            const float m = min32(v);
            second = fminf(second, fmaxf(best, m));
            const bool lt = m < best;
            best = fminf(best, m);
            bchunk = lt ? cid : bchunk;
        };
        auto release = [&](int g) {
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_empty + 8 * (g & 1));
        };
        auto fence = [&]() { if (best != best) { out[0] = 1.f; __trap(); } };
        uint32_t a0[32], a1[32];
        if (org == 0 || org == 3) {
            uint32_t ready = 0;
            for (int g = 0; g < ntiles; ++g) {
                const uint32_t ta = tl + (g & 1) * 256;
                if (!ready) mbar_wait(bar_full + 8 * (g & 1), (g >> 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                ld32(ta, a0); ld32(ta + 32, a1); ldwait2(a0, a1);
                book(a0, g * 8 + c * 4); book(a1, g * 8 + c * 4 + 1);
                ld32(ta + 64, a0); ld32(ta + 96, a1); ldwait2(a0, a1);
                release(g);
                ready = 0;
                if (org == 3 && g + 1 < ntiles) ready = test_wait(bar_full + 8 * ((g + 1) & 1), ((g + 1) >> 1) & 1);
                book(a0, g * 8 + c * 4 + 2); book(a1, g * 8 + c * 4 + 3);
            }
        } else {
            uint32_t b0[32], b1[32];
            mbar_wait(bar_full, 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            ld32(tl, a0); ld32(tl + 32, a1); ld32(tl + 64, b0); ld32(tl + 96, b1);
            ldwait4(a0, a1, b0, b1);
            release(0);
            for (int g = 0; g < ntiles; ++g) {
                // zones a (columns 0-63) and b (64-127) hold tile g, landed; its buffer is already back with the MMA thread
                const bool nx = g + 1 < ntiles;
                const uint32_t tn = tl + ((g + 1) & 1) * 256;
                book(a0, g * 8 + c * 4); book(a1, g * 8 + c * 4 + 1);
                if (org == 1) fence();
                if (nx) {
                    mbar_wait(bar_full + 8 * ((g + 1) & 1), ((g + 1) >> 1) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    ld32(tn, a0); ld32(tn + 32, a1);
                }
                if (org == 1) fence();
                book(b0, g * 8 + c * 4 + 2); book(b1, g * 8 + c * 4 + 3);
                if (org == 1) fence();
                if (nx) {
                    ld32(tn + 64, b0); ld32(tn + 96, b1);
                    ldwait4(a0, a1, b0, b1);
                    release(g + 1);
                }
            }
        }
        out[blockIdx.x * 512 + threadIdx.x] = best + second + (float)bchunk;
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
    long long *d_cyc, h_cyc[148];
    float *d_out;
    CK(cudaMalloc(&d_cyc, sizeof(h_cyc)));
    CK(cudaMalloc(&d_out, 148 * 512 * sizeof(float)));
    CK(cudaFuncSetAttribute(scan, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    const int ntiles = 1024;
    const char *on[] = {"kernel order: ld64, reduce, ld64, release, reduce", "two 64-col zones, early release, next tile's loads between the reductions (fenced)",
                        "same without the scheduling fences", "kernel order + early probe of the next full barrier"};
    for (int org = 0; org < 4; ++org) {
        for (int rep = 0; rep < 2; ++rep) { scan<<<148, 288, 96 * 1024>>>(org, ntiles, d_cyc, d_out); CK(cudaDeviceSynchronize()); }
        CK(cudaMemcpy(h_cyc, d_cyc, sizeof(h_cyc), cudaMemcpyDeviceToHost));
        double mean = 0;
        for (int i = 0; i < 148; ++i) mean += h_cyc[i];
        printf("org %d  %-90s %8.1f cycles per 256-column tile\n", org, on[org], mean / 148 / ntiles);
        fflush(stdout);
    }
    return 0;
}
