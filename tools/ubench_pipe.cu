// MMA -> TMEM -> scanner pipeline microbenchmark on sm_100a: cycles per 128x128 tile for different consumer
// organisations, isolating protocol latency (mbarrier hops, tcgen05.commit), TMEM read latency under a concurrently
// writing tensor pipe, and the min-reduction work.  One CTA per SM, warp 16 lane 0 issues kind::f16 M128 N128 K16 MMAs
// into 4 TMEM buffers of 128 columns; warps 0..15 consume.
//   org 0: every scanner warp takes one 32-column chunk of EVERY tile (empty count 16)
//   org 1: scanner warp (r, c) takes all 4 chunks of the tiles of buffer c (empty count 4)
//   org 2: org 0, but the chunk's min work is done BEFORE the next wait while the next tile's load is already in flight
//          (two register buffers, loads of consecutive tiles overlap the compute)
//   work 0: wait + arrive only; 1: + tcgen05.ld/wait::ld; 2: + min32 bookkeeping
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o ubench_pipe ubench_pipe.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ int g_poll = 0;   // 0: try_wait (hardware suspend)   1: test_wait busy poll   2: test_wait + nanosleep 32
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    const int mode = g_poll;
    if (mode == 3) {   // one elected lane polls, the warp reconverges behind it
        if ((threadIdx.x & 31) == 0)
            for (long long spins = 0; !ok && spins < 4000000LL; ++spins)
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        __syncwarp();
        return true;
    }
    if (mode == 0) {
        for (long long spins = 0; !ok && spins < 4000000LL; ++spins)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } else {
        for (long long spins = 0; !ok && spins < 4000000LL; ++spins) {
            asm volatile("{ .reg .pred p; mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
            if (!ok && mode == 2) __nanosleep(32);
        }
    }
    return ok != 0;
}
__device__ __forceinline__ float min3(float a, float b, float c) { float r; asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&a)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7]), "=r"(a[8]), "=r"(a[9]), "=r"(a[10]), "=r"(a[11]), "=r"(a[12]), "=r"(a[13]), "=r"(a[14]), "=r"(a[15]),
                   "=r"(a[16]), "=r"(a[17]), "=r"(a[18]), "=r"(a[19]), "=r"(a[20]), "=r"(a[21]), "=r"(a[22]), "=r"(a[23]), "=r"(a[24]), "=r"(a[25]), "=r"(a[26]), "=r"(a[27]), "=r"(a[28]), "=r"(a[29]), "=r"(a[30]), "=r"(a[31])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void ldwait(uint32_t (&a)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]), "+r"(a[8]), "+r"(a[9]), "+r"(a[10]), "+r"(a[11]), "+r"(a[12]), "+r"(a[13]), "+r"(a[14]), "+r"(a[15]),
                   "+r"(a[16]), "+r"(a[17]), "+r"(a[18]), "+r"(a[19]), "+r"(a[20]), "+r"(a[21]), "+r"(a[22]), "+r"(a[23]), "+r"(a[24]), "+r"(a[25]), "+r"(a[26]), "+r"(a[27]), "+r"(a[28]), "+r"(a[29]), "+r"(a[30]), "+r"(a[31])
                 :: "memory");
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

__global__ void __launch_bounds__(544, 1) pipe(int org, int work_, int ntiles, long long *cycles, float *out) {
    const int work = work_ % 10, freerun = work_ / 10;   // freerun: the MMA thread does not wait for empty buffers
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bars[8];
    __shared__ uint32_t tmem_s;
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar_full = smem_u32(&bars[0]), bar_empty = smem_u32(&bars[4]);
    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, org == 1 ? 4 : 16); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 16) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_s)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_s;
    const long long t0 = clock64();
    if (warp == 16) {
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
            const uint64_t dbase = ((uint64_t)(128u >> 4) << 16) | ((uint64_t)(256u >> 4) << 32) | (1ull << 46);
            const uint64_t ad = dbase | ((smem_u32(smem) & 0x3ffffu) >> 4);
            uint64_t bd = dbase | (((smem_u32(smem) + 8192u) & 0x3ffffu) >> 4);
            if (org == 3) {
                const uint32_t idesc2 = (1u << 4) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
                for (int g = 0; g < ntiles / 2; ++g) {     // ntiles counts 128-column units: one N=256 MMA covers two
                    const int b = g & 1;
                    mbar_wait(bar_empty + 8 * b, ((g >> 1) & 1) ^ 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p; }" ::"r"(tmem + b * 256), "l"(ad), "l"(bd + (uint64_t)((g & 7) * 512)), "r"(idesc2), "r"(0u) : "memory");
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_full + 8 * b) : "memory");
                }
            } else
            {
            long long aw = 0, am = 0, ac = 0;
            for (int g = 0; g < ntiles; ++g) {
                const int b = g & 3;
                const long long c0 = clock64();
                if (freerun != 1) mbar_wait(bar_empty + 8 * b, ((g >> 2) & 1) ^ 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const long long c1 = clock64();
                if (freerun != 2) asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p; }" ::"r"(tmem + b * 128), "l"(ad), "l"(bd + (uint64_t)((g & 15) * 256)), "r"(idesc), "r"(0u) : "memory");
                const long long c2 = clock64();
                if (freerun == 2) mbar_arrive(bar_full + 8 * b);    // no tensor work at all: plain mbarrier arrive
                else asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_full + 8 * b) : "memory");
                const long long c3 = clock64();
                aw += c1 - c0; am += c2 - c1; ac += c3 - c2;
            }
            cycles[148 + blockIdx.x * 4 + 0] = aw; cycles[148 + blockIdx.x * 4 + 1] = am; cycles[148 + blockIdx.x * 4 + 2] = ac;
            cycles[148 + blockIdx.x * 4 + 3] = clock64() - t0;
            }
        }
        __syncwarp();
    } else {
        const int r = warp & 3, c = warp >> 2;
        const uint32_t tl = tmem + ((uint32_t)(r * 32) << 16);
        float best = 1e30f, second = 1e30f;
        int bchunk = 0;
        auto book = [&](const uint32_t (&v)[32], int cid) {
            const float m = min32(v);
            second = fminf(second, fmaxf(best, m));
            const bool lt = m < best;
            best = fminf(best, m);
            bchunk = lt ? cid : bchunk;
        };
        if (org == 0) {
            for (int g = 0; g < ntiles; ++g) {
                const int b = g & 3;
                mbar_wait(bar_full + 8 * b, (g >> 2) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                uint32_t v[32];
                if (work >= 1) { ld32(tl + b * 128 + c * 32, v); ldwait(v); }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_empty + 8 * b);
                if (work >= 2) book(v, g * 4 + c);
                else if (work == 1) best = fminf(best, __uint_as_float(v[lane]));
            }
        } else if (org == 1) {
            for (int g = c; g < ntiles; g += 4) {
                mbar_wait(bar_full + 8 * c, (g >> 2) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
                for (int j = 0; j < 4; ++j) {
                    uint32_t v[32];
                    if (work >= 1) { ld32(tl + c * 128 + j * 32, v); ldwait(v); }
                    if (j == 3) {
                        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_empty + 8 * c);
                    }
                    if (work >= 2) book(v, g * 4 + j);
                    else if (work == 1) best = fminf(best, __uint_as_float(v[lane]));
                }
            }
        } else if (org == 3) {
            for (int g = 0; g < ntiles / 2; ++g) {
                const int b = g & 1;
                mbar_wait(bar_full + 8 * b, (g >> 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
                for (int j = 0; j < 2; ++j) {
                    uint32_t v[32];
                    ld32(tl + b * 256 + c * 64 + j * 32, v); ldwait(v);
                    if (j == 1) {
                        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_empty + 8 * b);
                    }
                    book(v, g * 8 + c * 2 + j);
                }
            }
        } else {
            // software pipelined: the load of tile g+1 is in flight while tile g's chunk is reduced
            uint32_t va[32], vb[32];
            if (ntiles > 0) {
                mbar_wait(bar_full, 0);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                ld32(tl + c * 32, va);
            }
            for (int g = 0; g < ntiles; g += 2) {
                // tile g is in va (in flight); issue tile g+1 into vb, then finish va
                ldwait(va);
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_empty + 8 * (g & 3));
                if (g + 1 < ntiles) {
                    mbar_wait(bar_full + 8 * ((g + 1) & 3), ((g + 1) >> 2) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    ld32(tl + ((g + 1) & 3) * 128 + c * 32, vb);
                }
                book(va, g * 4 + c);
                if (g + 1 < ntiles) {
                    ldwait(vb);
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_empty + 8 * ((g + 1) & 3));
                    if (g + 2 < ntiles) {
                        mbar_wait(bar_full + 8 * ((g + 2) & 3), ((g + 2) >> 2) & 1);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        ld32(tl + ((g + 2) & 3) * 128 + c * 32, va);
                    }
                    book(vb, (g + 1) * 4 + c);
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
    if (warp == 16) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
    long long *d_cyc, h_cyc[148 * 5];
    float *d_out;
    CK(cudaMalloc(&d_cyc, sizeof(h_cyc)));
    CK(cudaMalloc(&d_out, 148 * 512 * sizeof(float)));
    CK(cudaFuncSetAttribute(pipe, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    const int ntiles = 1024;
    for (int poll = 0; poll < 4; poll += 3) {
    CK(cudaMemcpyToSymbol(g_poll, &poll, sizeof(int)));
    printf("---- polling mode %d (0 try_wait by all lanes, 3 try_wait by lane 0 + __syncwarp)\n", poll);
    const char *on[] = {"every warp, one chunk of every tile", "warp owns buffer c, 4 chunks per tile", "every warp, one chunk, next load in flight", "N=256 tiles, 2 buffers, 2 chunks per warp"};
    const char *wn[] = {"wait + arrive only", "+ tcgen05.ld", "+ min32 bookkeeping"};
    for (int org = 0; org < 4; ++org)
        for (int work = (org >= 2 ? 2 : 0); work < 3; ++work) {
            for (int rep = 0; rep < 2; ++rep) {
                pipe<<<148, 544, 96 * 1024>>>(org, work, ntiles, d_cyc, d_out);
                CK(cudaDeviceSynchronize());
            }
            CK(cudaMemcpy(h_cyc, d_cyc, sizeof(h_cyc), cudaMemcpyDeviceToHost));
            double mean = 0;
            for (int i = 0; i < 148; ++i) mean += h_cyc[i];
            printf("%-45s %-22s %8.1f cycles per tile", on[org], wn[work], mean / 148 / ntiles);
            if (org < 3) printf("   MMA thread per tile: wait %.0f  mma %.0f  commit %.0f  loop total %.0f", h_cyc[148] / (double)ntiles, h_cyc[149] / (double)ntiles, h_cyc[150] / (double)ntiles, h_cyc[151] / (double)ntiles);
            printf("\n");
            fflush(stdout);
        }
    for (int org = 0; org < 2; ++org) {
        for (int rep = 0; rep < 2; ++rep) { pipe<<<148, 544, 96 * 1024>>>(org, 20, ntiles, d_cyc, d_out); CK(cudaDeviceSynchronize()); }
        CK(cudaMemcpy(h_cyc, d_cyc, sizeof(h_cyc), cudaMemcpyDeviceToHost));
        double mean = 0;
        for (int i = 0; i < 148; ++i) mean += h_cyc[i];
        printf("NO tensor work: producer = plain mbarrier arrive, consumers wait + arrive (org %d): %8.1f cycles per tile   producer per tile: wait %.0f  arrive %.0f  loop total %.0f\n", org, mean / 148 / ntiles,
               h_cyc[148] / (double)ntiles, h_cyc[150] / (double)ntiles, h_cyc[151] / (double)ntiles);
    }
    }
    return 0;
}
