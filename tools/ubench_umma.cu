// tcgen05.mma issue-rate microbenchmark on sm_100a: cycles per M=128 x N x K-slice instruction as a function of the
// shared-memory operand layout (no swizzle with different core-matrix placements, SWIZZLE_32B/64B/128B), of N and of
// the operand kind (tf32 K=8, bf16 K=16).  Operand contents are irrelevant (zeros); only timing is measured.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o ubench_umma ubench_umma.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

struct Cfg {
    uint32_t layout;     // descriptor layout_type: 0 none, 6 SW32, 4 SW64, 2 SW128
    uint32_t lbo, sbo;   // bytes
    uint32_t kadv;       // byte advance of the start address for the second K-slice
    uint32_t tile_bytes; // B operand bytes per N-tile (distance between consecutive tiles)
    uint32_t n;          // MMA N
    uint32_t bf16;       // 0: kind::tf32 (K=8), 1: kind::f16 with bf16 (K=16)
    uint32_t a_tiles;    // 1: one A tile; >1: rotate A too
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, const Cfg &c) {
    return (uint64_t)((saddr & 0x3ffffu) >> 4) | ((uint64_t)(c.lbo >> 4) << 16) | ((uint64_t)(c.sbo >> 4) << 32) | (1ull << 46) |
           ((uint64_t)c.layout << 61);
}

__global__ void __launch_bounds__(128, 1) umma_rate(Cfg c, int iters, long long *cycles) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ uint32_t tmem_s;
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1u) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_s)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_s;
    if (warp == 0) {
        const uint32_t a_base = smem_u32(smem), b_base = a_base + 32 * 1024;
        const uint32_t fmt = c.bf16 ? 1u : 2u;
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((c.n >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t nbuf_mask = 512 / c.n - 1;
        const uint64_t ad0 = make_desc(a_base, c), ad1 = make_desc(a_base + c.kadv, c);
        const uint64_t bd0 = make_desc(b_base, c), bd1 = make_desc(b_base + c.kadv, c);
        const uint64_t tile16 = c.tile_bytes >> 4;     // descriptor start-address units
        const long long t0 = clock64();
        if (lane == 0) {
#pragma unroll 4
            for (int i = 0; i < iters; ++i) {
                const uint32_t d = tmem + (i & nbuf_mask) * c.n;
                const uint64_t boff = (uint64_t)(i & (c.tile_bytes > 16384 ? 3 : 7)) * tile16;   // B tiles in rotation
                if (c.bf16) {
                    asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p; }" ::"r"(d), "l"(ad0), "l"(bd0 + boff), "r"(idesc), "r"(0u) : "memory");
                    asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p; }" ::"r"(d), "l"(ad1), "l"(bd1 + boff), "r"(idesc), "r"(1u) : "memory");
                } else {
                    asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p; }" ::"r"(d), "l"(ad0), "l"(bd0 + boff), "r"(idesc), "r"(0u) : "memory");
                    asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p; }" ::"r"(d), "l"(ad1), "l"(bd1 + boff), "r"(idesc), "r"(1u) : "memory");
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        }
        __syncwarp();
        uint32_t ok = 0;
        long long spins = 0;
        while (!ok && spins < 100000000LL) {
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
            ++spins;
        }
        const long long t1 = clock64();
        if (lane == 0) cycles[blockIdx.x] = ok ? (t1 - t0) : -1;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}


// Protocol cost per tile: mode 0 = MMA only; 1 = MMA + commit (round-robin over 4 mbarriers, nobody waits);
// 2 = MMA + commit + wait for that commit (full round trip); 3 = mode 1 + tcgen05.fence::after_thread_sync per tile.
__global__ void __launch_bounds__(128, 1) umma_protocol(int mode, int iters, long long *cycles) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bars[4];
    __shared__ uint32_t tmem_s;
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[i])), "r"(1u) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_s)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_s;
    if (warp == 0) {
        Cfg c = {0, 128, 256, 0, 4096, 128, 1, 1};
        const uint32_t a_base = smem_u32(smem), b_base = a_base + 32 * 1024;
        const uint32_t idesc = (1u << 4) | (0u << 7) | (0u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);   // f16
        const uint64_t ad = make_desc(a_base, c), bd = make_desc(b_base, c);
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const uint32_t d = tmem + (i & 3) * 128;
            const uint32_t bar = smem_u32(&bars[i & 3]);
            if (mode == 3) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0) {
                asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p; }" ::"r"(d), "l"(ad), "l"(bd + (uint64_t)((i & 7) * 256)), "r"(idesc), "r"(0u) : "memory");
                if (mode >= 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
            }
            __syncwarp();
            if (mode == 2) {
                uint32_t ok = 0;
                const uint32_t parity = (i >> 2) & 1;
                while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
            }
        }
        // drain: one more commit and wait for it
        if (lane == 0) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[iters & 3])) : "memory");
        __syncwarp();
        if (mode != 2) {
            // wait until the final commit has flipped its barrier's phase: count the phases it has been through
            uint32_t ok = 0;
            const int nth = (iters + 3 - (iters & 3)) / 4 + ((mode >= 1) ? 0 : 0);
            const uint32_t parity = (mode >= 1) ? (uint32_t)(((iters >> 2)) & 1) : 0u;
            long long spins = 0;
            while (!ok && spins < 50000000LL) {
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smem_u32(&bars[iters & 3])), "r"(parity) : "memory");
                ++spins;
            }
            (void)nth;
        }
        const long long t1 = clock64();
        if (lane == 0) cycles[blockIdx.x] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}


// Which property of the per-tile pattern is slow?  One thread issues `iters` kind::f16 M128 N128 K16 MMAs:
//   rot  : 1 = rotate over the 4 TMEM buffers, 0 = always buffer 0      acc: accumulate flag      nb: B tiles in rotation
__global__ void __launch_bounds__(128, 1) umma_pattern(int rot, int acc, int nb, int iters, long long *cycles, int commit_every = 0) {
    __shared__ __align__(8) unsigned long long cbars[4];
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ uint32_t tmem_s;
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1u) : "memory");
        for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&cbars[i])), "r"(1u) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_s)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_s;
    if (warp == 0) {
        Cfg c = {0, 128, 256, 0, 4096, 128, 1, 1};
        const uint32_t a_base = smem_u32(smem), b_base = a_base + 32 * 1024;
        const uint32_t idesc = (1u << 4) | ((128u >> 3) << 17) | ((128u >> 4) << 24);   // f16
        const uint64_t ad = make_desc(a_base, c), bd = make_desc(b_base, c);
        const long long t0 = clock64();
        if (lane == 0) {
            for (int i = 0; i < iters; ++i) {
                const uint32_t d = tmem + (rot ? (i & 3) * 128 : 0);
                asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p; }" ::"r"(d), "l"(ad), "l"(bd + (uint64_t)((i & (nb - 1)) * 256)), "r"(idesc), "r"((uint32_t)acc) : "memory");
                if (commit_every > 0 && (i % commit_every) == commit_every - 1)
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&cbars[(i / commit_every) & 3])) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        }
        __syncwarp();
        uint32_t ok = 0;
        long long spins = 0;
        while (!ok && spins < 100000000LL) {
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
            ++spins;
        }
        const long long t1 = clock64();
        if (lane == 0) cycles[blockIdx.x] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}


// Handshake latencies.  mode 0: one thread: tcgen05.mma + commit, then try_wait until the commit lands (round trip).
// mode 1: mbarrier ping-pong between lane 0 of warp 0 and lane 0 of warp 1 (arrive -> the other side's try_wait returns).
// mode 2: like mode 1 but all 32 lanes of both warps poll (as the kernel's consumer warps do).
__global__ void __launch_bounds__(128, 1) handshake(int mode, int iters, long long *cycles) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bars[2];
    __shared__ uint32_t tmem_s;
    for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[i])), "r"(1u) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_s)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_s;
    const uint32_t b0 = smem_u32(&bars[0]), b1 = smem_u32(&bars[1]);
    auto wait = [&](uint32_t bar, uint32_t parity) {
        uint32_t ok = 0;
        long long spins = 0;
        while (!ok && spins < 20000000LL) {
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
            ++spins;
        }
    };
    const long long t0 = clock64();
    if (mode == 0) {
        if (warp == 0 && lane == 0) {
            Cfg c = {0, 128, 256, 0, 4096, 128, 1, 1};
            const uint32_t idesc = (1u << 4) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
            const uint64_t ad = make_desc(smem_u32(smem), c), bd = make_desc(smem_u32(smem) + 8192, c);
            for (int i = 0; i < iters; ++i) {
                asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p; }" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(0u) : "memory");
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(b0) : "memory");
                wait(b0, (uint32_t)(i & 1));
            }
        }
    } else {
        const bool poll = mode == 2 || lane == 0;
        if (warp == 0 && poll) {
            for (int i = 0; i < iters; ++i) {
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b0) : "memory");
                wait(b1, (uint32_t)(i & 1));
            }
        } else if (warp == 1 && poll) {
            for (int i = 0; i < iters; ++i) {
                wait(b0, (uint32_t)(i & 1));
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b1) : "memory");
            }
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
    long long *d_cyc, h_cyc[148];
    CK(cudaMalloc(&d_cyc, sizeof(h_cyc)));
    CK(cudaFuncSetAttribute(umma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    struct Named { const char *name; Cfg c; };
    const Named cases[] = {
        // name                                              layout lbo  sbo   kadv tile    n  bf16 a_tiles
        {"tf32 none  LBO=128 SBO=512 N=128 (kernel v1)",     {0, 128, 512, 256, 8192, 128, 0, 1}},
        {"tf32 none  LBO=128 SBO=512 N=256",                 {0, 128, 512, 256, 16384, 256, 0, 1}},
        {"tf32 none  LBO=128 SBO=256 (K-slices apart) N=128",{0, 128, 256, 4096, 8192, 128, 0, 1}},
        {"tf32 none  LBO=144 SBO=640 N=128",                 {0, 144, 640, 288, 10240, 128, 0, 1}},
        {"tf32 none  LBO=192 SBO=768 N=128",                 {0, 192, 768, 384, 12288, 128, 0, 1}},
        {"tf32 none  LBO=2048 SBO=128 (chunk-major) N=128",  {0, 2048, 128, 4096, 8192, 128, 0, 1}},
        {"tf32 SW32  SBO=256 (K-slices apart) N=128",        {6, 16, 256, 4096, 8192, 128, 0, 1}},
        {"tf32 SW64  SBO=512 N=128",                         {4, 16, 512, 32, 8192, 128, 0, 1}},
        {"tf32 SW64  SBO=512 N=256",                         {4, 16, 512, 32, 16384, 256, 0, 1}},
        {"tf32 SW128 SBO=1024 N=128",                        {2, 16, 1024, 32, 16384, 128, 0, 1}},
        {"tf32 SW128 SBO=1024 N=256",                        {2, 16, 1024, 32, 32768, 256, 0, 1}},
        {"bf16 none  LBO=128 SBO=512 N=128",                 {0, 128, 512, 256, 8192, 128, 1, 1}},
        {"bf16 SW64  SBO=512 N=128",                         {4, 16, 512, 32, 8192, 128, 1, 1}},
        {"bf16 SW128 SBO=1024 N=128",                        {2, 16, 1024, 32, 16384, 128, 1, 1}},
        {"bf16 SW128 SBO=1024 N=256",                        {2, 16, 1024, 32, 32768, 256, 1, 1}},
    };
    const int iters = 2000;
    for (const Named &nc : cases) {
        for (int rep = 0; rep < 2; ++rep) {
            umma_rate<<<148, 128, 200 * 1024>>>(nc.c, iters, d_cyc);
            CK(cudaDeviceSynchronize());
        }
        CK(cudaMemcpy(h_cyc, d_cyc, sizeof(h_cyc), cudaMemcpyDeviceToHost));
        double mean = 0;
        bool bad = false;
        for (int i = 0; i < 148; ++i) { mean += h_cyc[i]; bad |= h_cyc[i] < 0; }
        mean /= 148;
        printf("%-52s %8.1f cycles per MMA instruction (ideal %3u)%s\n", nc.name, mean / (2.0 * iters), nc.c.n / 2,
               bad ? "  [TIMEOUT]" : "");
        fflush(stdout);
    }
    CK(cudaFuncSetAttribute(umma_protocol, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    const char *mnames[] = {"MMA only (f16 N=128)", "MMA + commit per tile", "MMA + commit + wait per tile (round trip)", "MMA + commit + fence::after per tile"};
    for (int mode = 0; mode < 4; ++mode) {
        const int it = 2000;   // multiple of 4: the drain commit lands on barrier 0
        for (int rep = 0; rep < 2; ++rep) {
            umma_protocol<<<148, 128, 200 * 1024>>>(mode, it, d_cyc);
            CK(cudaDeviceSynchronize());
        }
        CK(cudaMemcpy(h_cyc, d_cyc, sizeof(h_cyc), cudaMemcpyDeviceToHost));
        double mean = 0;
        for (int i = 0; i < 148; ++i) mean += h_cyc[i];
        printf("protocol: %-45s %8.1f cycles per tile\n", mnames[mode], mean / 148 / it);
        fflush(stdout);
    }
    CK(cudaFuncSetAttribute(umma_pattern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    for (int rot = 0; rot < 2; ++rot)
        for (int acc = 0; acc < 2; ++acc)
            for (int nb = 1; nb <= 16; nb *= 16) {
                for (int rep = 0; rep < 2; ++rep) {
                    umma_pattern<<<148, 128, 200 * 1024>>>(rot, acc, nb, 2000, d_cyc);
                    CK(cudaDeviceSynchronize());
                }
                CK(cudaMemcpy(h_cyc, d_cyc, sizeof(h_cyc), cudaMemcpyDeviceToHost));
                double mean = 0;
                for (int i = 0; i < 148; ++i) mean += h_cyc[i];
                printf("pattern: rotate D=%d accumulate=%d B tiles=%2d  %8.1f cycles per MMA\n", rot, acc, nb, mean / 148 / 2000);
                fflush(stdout);
            }
    for (int ce = 1; ce <= 8; ce *= 2) {
        for (int rep = 0; rep < 2; ++rep) {
            umma_pattern<<<148, 128, 200 * 1024>>>(1, 0, 16, 2000, d_cyc, ce);
            CK(cudaDeviceSynchronize());
        }
        CK(cudaMemcpy(h_cyc, d_cyc, sizeof(h_cyc), cudaMemcpyDeviceToHost));
        double mean = 0;
        for (int i = 0; i < 148; ++i) mean += h_cyc[i];
        printf("pattern: commit (nobody waits) after every %d MMA(s): %8.1f cycles per MMA\n", ce, mean / 148 / 2000);
        fflush(stdout);
    }
    CK(cudaFuncSetAttribute(handshake, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    const char *hn[] = {"MMA + commit + wait round trip (one thread)", "mbarrier ping-pong, one lane per side (2 hops)", "mbarrier ping-pong, 32 lanes poll (2 hops)"};
    for (int mode = 0; mode < 3; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            handshake<<<148, 128, 64 * 1024>>>(mode, 2000, d_cyc);
            CK(cudaDeviceSynchronize());
        }
        CK(cudaMemcpy(h_cyc, d_cyc, sizeof(h_cyc), cudaMemcpyDeviceToHost));
        double mean = 0;
        for (int i = 0; i < 148; ++i) mean += h_cyc[i];
        printf("handshake: %-50s %8.1f cycles per iteration\n", hn[mode], mean / 148 / 2000);
        fflush(stdout);
    }
    return 0;
}
