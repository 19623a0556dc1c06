// chamfer_nn_tc.cu -- Chamfer NN forward with the FILTER on the 5th-generation tensor cores (tcgen05 / TMEM).
//
// Same contract and the same exactness scheme as chamfer.cu (reference: metric/chamfer3D/chamfer3D.cu:12-154):
// dist/idx always come from the reference's exact formula; a filter only decides where to look.  Here the filter
//     a_k = |t_k - c|^2 - 2 (q - c).(t_k - c)            ( = |t_k - q|^2 - |q - c|^2 )
// is a [128 queries] x [128 targets] x K=16 FP16 GEMM per tile: ONE tcgen05.mma.kind::f16 issued by one thread.
// fp32-class accuracy comes from (1) a per-cloud power-of-two scale s that brings max|t-c| into [0.5,1) -- exact in
// fp32, keeps every term inside the fp16 exponent range -- and (2) splitting every operand into two fp16 terms
// (hi + lo, 22 significand bits) and |t-c|^2 into three, laid out along K so that the cross terms line up:
//     A row (query) : [qh qh ql]x [qh qh ql]y [qh qh ql]z  1  1  1  0 0 0 0      q' = -2 s (q - c) = qh + ql
//     B row (target): [th tl th]x [th tl th]y [th tl th]z  w1 w2 w3 0 0 0 0      t' = s (t - c) = th + tl, |t'|^2 = w1+w2+w3
// (products of two 11-bit significands are exact in the fp32 accumulator; the dropped ql*tl terms and the split
// residues are bounded by 3*2^-22 |q'||t'|).  Measured on B200 (tools/ubench_umma.cu): kind::f16 M128 N128 K16 issues
// every 64 cycles, kind::tf32 K8 only every 96, independent of the shared-memory layout -- hence fp16 and one
// instruction per tile instead of two TF32 K-slices.  Accumulators live in TMEM (4 buffers of 128 columns).  The 16
// consumer warps read them back with tcgen05.ld.32x32b.x16 -- one thread owns one query row, so the running minimum
// needs no cross-lane traffic -- and keep per 32-target chunk (best chunk minimum, its chunk id, second best) exactly
// like the FFMA kernel: 15 FMNMX3 + 6 ALU ops per 32 pairs instead of 96 FFMA + 16 FMNMX3.
//
// CTA = one per SM, persistent: warps 0-15 consumers (warp w: TMEM lanes 32*(w%4).., buffer w/4), warp 16 issues
// the MMAs.  Per unit (128 queries of one cloud/direction): raw queries/targets arrive by cp.async one unit ahead;
// consumers build the B operand (once per cloud) and the A operand in shared memory (K-major, no swizzle: 8-row x
// 16-byte core matrices), signal the MMA warp, scan the tiles as their TMEM buffers fill (full/empty mbarriers,
// tcgen05.commit), park their partial results, stage the NEXT unit, and only then resolve the current one (merge,
// margin test, exact rescan of the best chunk from the raw targets in shared memory, fused loss-sum / F-score
// epilogue).  The few queries that fail the margin test are deferred to a per-CTA list and take an exact full scan,
// one warp per query, at the end.
#include <cuda_fp16.h>
#include "chamfer_nn.cuh"

namespace psd {
namespace tc {

constexpr int kConsWarps = 16;
constexpr int kConsThreads = kConsWarps * 32;
constexpr int kThreadsTC = kConsThreads + 32;   // + the MMA warp
constexpr int kTileN = 128;                     // targets per MMA tile = TMEM buffer width (columns)
constexpr int kBufs = 4;                        // TMEM buffers: 4 x 128 columns = all 512
constexpr int kCh = 32;                         // targets per filter chunk (one tcgen05.ld.x32)
constexpr int kMaxT = 2048;                     // targets resident in shared memory (B operand: 32 B per target)
constexpr float kPadW = 32768.0f;               // padding |t'|^2 (fp16-exact); real filter values are < 3 + 2*1e4*1.8
constexpr float kQMax = 4096.0f;                // scaled |q-c| above this sends the query to the exact scan: keeps |a| <= 3 + 3.5*kQMax < kPadW

// shared-memory carve-up (bytes)
constexpr int kOffB = 0;                                  // [kMaxT/8][2][8][16 B]
constexpr int kOffA = kOffB + kMaxT * 32;                 // [16][2][8][16 B]
constexpr int kOffRaw = kOffA + kQB * 32;                 // [2][3][kMaxT] raw target coordinates (SoA), by cloud parity
constexpr int kOffPart = kOffRaw + 2 * 3 * kMaxT * 4;     // [2][4 column groups][3][128]
constexpr int kOffSq = kOffPart + 2 * 4 * 3 * kQB * 4;    // [3][3][128] raw queries, by unit mod 3
constexpr int kFbCap = 512;                               // deferred exact-scan list (entries: unit << 8 | query)
constexpr int kOffFb = kOffSq + 3 * 3 * kQB * 4;
constexpr int kOffStat = kOffFb + kFbCap * 4;             // [16] wmax, [16] bad, [16] cmax
constexpr int kOffBar = kOffStat + 3 * kConsWarps * 4;    // 9 mbarriers
constexpr int kOffMisc = kOffBar + 16 * 8;                // tmem base, nfb, abort
constexpr int kSmemTC = kOffMisc + 64;

__device__ unsigned long long g_fallback_queries_tc = 0ull;
__device__ int g_tc_error = 0;

// ---- PTX wrappers -------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok;
}
// Bounded wait: a protocol bug must not hang the GPU.  On time-out the CTA-wide abort flag makes every later wait
// fall through; the host sees g_tc_error and reports the launch as failed.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, volatile int *abort_flag) {
    if (mbar_try_wait(bar, parity)) return;
    long long t0 = 0;
    for (unsigned spins = 1;; ++spins) {
        if (mbar_try_wait(bar, parity)) return;   // try_wait suspends the warp in hardware for a bounded time
        if ((spins & 63u) == 0u) {
            if (*abort_flag) return;
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            if (now - t0 > 400000000LL) { *abort_flag = 1; atomicExch(&g_tc_error, 1); return; }
        }
    }
}
__device__ __forceinline__ void cp_async4(void *smem_dst, const float *gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void cons_bar() { asm volatile("bar.sync 1, %0;" ::"n"(kConsThreads) : "memory"); }
// x -> (hi, lo) fp16 pair with hi + lo = x up to 2^-22 |x| (or 2^-25 absolute in the subnormal range)
__device__ __forceinline__ void split_h(float x, unsigned short &hi, unsigned short &lo) {
    const __half h = __float2half_rn(x);
    const __half l = __float2half_rn(x - __half2float(h));
    hi = __half_as_ushort(h); lo = __half_as_ushort(l);
}
__device__ __forceinline__ uint32_t pack2(unsigned short a, unsigned short b) { return (uint32_t)a | ((uint32_t)b << 16); }
// K-major, no swizzle: core matrix = 8 rows x 16 B (128 contiguous bytes, 8 fp16 per row); LBO = distance between
// the two core matrices of the K=16 slice (128 B), SBO = distance between 8-row groups (256 B).  Version 1 (sm_100).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3ffffu) >> 4) | ((uint64_t)(128u >> 4) << 16) | ((uint64_t)(256u >> 4) << 32) |
           (1ull << 46);
}
// kind::f16, D = F32, A/B = F16 (format 0), both K-major, M = 128, N = 128
constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)(kTileN >> 3) << 17) | ((128u >> 4) << 24);
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p; }"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(kIdesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// one tcgen05.ld of 16 consecutive columns of this thread's TMEM lane; completes at the next tmem_wait()
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
// wait for the outstanding tcgen05.ld; the registers are in/out operands so that no use of them can be scheduled above
__device__ __forceinline__ void tmem_wait(uint32_t (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                   "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :: "memory");
}
// minimum of 16 filter values: 7 FMNMX3 + 1 FMNMX
__device__ __forceinline__ float min16(const uint32_t (&r)[16]) {
    float a = fmin3(__uint_as_float(r[0]), __uint_as_float(r[1]), __uint_as_float(r[2]));
    float b = fmin3(__uint_as_float(r[3]), __uint_as_float(r[4]), __uint_as_float(r[5]));
    float c = fmin3(__uint_as_float(r[6]), __uint_as_float(r[7]), __uint_as_float(r[8]));
    float d = fmin3(__uint_as_float(r[9]), __uint_as_float(r[10]), __uint_as_float(r[11]));
    a = fmin3(a, __uint_as_float(r[12]), __uint_as_float(r[13]));
    b = fmin3(b, __uint_as_float(r[14]), __uint_as_float(r[15]));
    return fminf(fmin3(a, b, c), d);
}

struct Unit {
    int d, cloud, qblock;
};
__device__ __forceinline__ Unit decode_unit(const NNParams &p, int blk) {
    Unit u;
    u.d = blk >= p.blocks_dir0 ? 1 : 0;
    const int bid = u.d ? blk - p.blocks_dir0 : blk;
    const int qbn = p.dir[u.d].qblocks;
    u.cloud = bid / qbn;
    u.qblock = bid - u.cloud * qbn;
    return u;
}
__device__ __forceinline__ int group_of(const Unit &u) { return u.d * 0x40000000 + u.cloud; }

// DBG: instrumented build -- dumps every filter value to dbg[(unit*128 + row) * dbg_ld + target] when dbg != nullptr
// (calibration / bring-up) and writes phase clocks to prof (tools/tc_phase_clocks.py) when prof != nullptr.
template <bool DBG>
__global__ void __launch_bounds__(kThreadsTC, 1) chamfer_nn_tc_kernel(const NNParams p, float *dbg, int dbg_ld, long long *prof) {
    extern __shared__ __align__(128) unsigned char smem[];
    float *sraw = reinterpret_cast<float *>(smem + kOffRaw);
    float *part = reinterpret_cast<float *>(smem + kOffPart);
    float *sq = reinterpret_cast<float *>(smem + kOffSq);
    int *fb_list = reinterpret_cast<int *>(smem + kOffFb);
    float *s_wstat = reinterpret_cast<float *>(smem + kOffStat);
    int *s_bstat = reinterpret_cast<int *>(smem + kOffStat) + kConsWarps;
    float *s_cstat = reinterpret_cast<float *>(smem + kOffStat) + 2 * kConsWarps;
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(smem + kOffMisc);
    int *s_nfb = reinterpret_cast<int *>(smem + kOffMisc) + 1;
    volatile int *s_abort = reinterpret_cast<volatile int *>(smem + kOffMisc) + 2;
    const uint32_t sB_addr = smem_u32(smem + kOffB), sA_addr = smem_u32(smem + kOffA);
    const uint32_t bar0 = smem_u32(smem + kOffBar);
    const uint32_t bar_full = bar0, bar_empty = bar0 + 8 * kBufs, bar_ready = bar0 + 16 * kBufs;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x;
    const int blk_begin = (int)(((long long)blockIdx.x * p.total_blocks) / G);
    const int blk_end = (int)(((long long)(blockIdx.x + 1) * p.total_blocks) / G);
    const int nunits = blk_end - blk_begin;
    // optional phase clocks (tools/tc_phase_clocks.py): 64 slots per CTA
    long long *pf = (DBG && prof) ? prof + (long long)blockIdx.x * 64 : nullptr;
    auto stamp = [&](int slot) { if (DBG && pf && slot < 56) pf[slot] = clock64(); };
    if (tid == 0) stamp(0);

    if (tid == 0) {
        for (int i = 0; i < kBufs; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 4); }
        mbar_init(bar_ready, 1);
        *s_nfb = 0;
        *s_abort = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kConsWarps) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;
    if (tid == 0) stamp(1);

    if (warp == kConsWarps) {
        // ================================================= MMA warp
        int g = 0;
        for (int ul = 0; ul < nunits; ++ul) {
            const Unit u = decode_unit(p, blk_begin + ul);
            const int ntiles = (p.dir[u.d].nt + kTileN - 1) / kTileN;
            long long w0 = DBG ? clock64() : 0;
            mbar_wait(bar_ready, ul & 1, s_abort);
            if (DBG && pf && lane == 0) pf[56] += clock64() - w0;
            tc_fence_after();
            for (int t = 0; t < ntiles; ++t, ++g) {
                const int b = g & (kBufs - 1);
                w0 = DBG ? clock64() : 0;
                mbar_wait(bar_empty + 8 * b, ((g >> 2) & 1) ^ 1, s_abort);
                if (DBG && pf && lane == 0) pf[57] += clock64() - w0;
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t d_tmem = tmem_base + (uint32_t)(b * kTileN);
                    umma_f16(d_tmem, umma_desc(sA_addr), umma_desc(sB_addr + (uint32_t)t * (kTileN * 32)), 0u);
                    umma_commit(bar_full + 8 * b);
                }
                __syncwarp();
            }
            if (DBG && pf && lane == 0) pf[58] = clock64();
        }
    } else {
        // ================================================= consumer warps
        const int r = warp & 3, c = warp >> 2;
        const int row = r * 32 + lane;
        const uint32_t taddr0 = tmem_base + ((uint32_t)(r * 32) << 16) + (uint32_t)(c * kTileN);

        // frame of the resident B operand (identical in every thread)
        int res_group = -1, res_buf = 0;      // cloud/direction whose B operand is resident; its raw-target buffer
        float cx = 0.f, cy = 0.f, cz = 0.f, cs = 1.f, res_wmax = 0.f;   // centre, power-of-two scale, max |t'|^2
        int res_bad = 0;
        int pf_group = -1, pf_buf = 1;        // cloud/direction of the most recent raw-target prefetch; its buffer

        // cp.async prefetch of unit ul's raw queries (and raw targets if its cloud/direction is not the prefetched one)
        auto prefetch = [&](int ul) {
            const Unit u = decode_unit(p, blk_begin + ul);
            const NNDirection &D = p.dir[u.d];
            if (tid < 3 * kQB) {
                const int comp = tid >> 7, ql = tid & (kQB - 1);
                int j = D.q_begin + u.qblock * kQB + ql;
                const int q_last = D.q_begin + D.q_count - 1;
                j = j < q_last ? j : q_last;
                cp_async4(sq + ((ul % 3) * 3 + comp) * kQB + ql, D.q + (long long)u.cloud * D.q_bs + j * D.q_ps + comp * D.q_cs);
            }
            const int group = group_of(u);
            if (group != pf_group) {
                pf_buf ^= 1;
                pf_group = group;
                const float *__restrict__ tb = D.t + (long long)u.cloud * D.t_bs;
                const int nt = D.nt;
#pragma unroll
                for (int comp = 0; comp < 3; ++comp)
                    for (int k = tid; k < nt; k += kConsThreads)
                        cp_async4(sraw + (pf_buf * 3 + comp) * kMaxT + k, tb + k * D.t_ps + comp * D.t_cs);
            }
        };
        // does prefetch(ul) load raw targets?  (uniform; decided before the call to place a barrier in front of it)
        auto prefetch_needs_targets = [&](int ul) { return group_of(decode_unit(p, blk_begin + ul)) != pf_group; };

        // build the tensor-core operands of unit ul from the prefetched raw data (shared memory only)
        auto stage = [&](int ul) {
            const Unit u = decode_unit(p, blk_begin + ul);
            const NNDirection &D = p.dir[u.d];
            const int group = group_of(u);
            if (group != res_group) {
                const int nt = D.nt;
                const float *rx = sraw + (pf_buf * 3) * kMaxT, *ry = rx + kMaxT, *rz = ry + kMaxT;
                {   // centre of the filter frame: mean of up to 8 evenly spaced targets (any value is correct)
                    float sx = 0.f, sy = 0.f, sz = 0.f;
                    const int ns = nt < 8 ? nt : 8;
                    const int step = nt >> 3;
#pragma unroll
                    for (int s8 = 0; s8 < 8; ++s8) {
                        if (s8 < ns) {
                            const int k = nt < 8 ? s8 : s8 * step;
                            sx += rx[k]; sy += ry[k]; sz += rz[k];
                        }
                    }
                    const float inv = 1.0f / (float)ns;
                    cx = sx * inv; cy = sy * inv; cz = sz * inv;
                }
                // pass 1: extent of the cloud around the centre -> power-of-two scale with max|t'|_inf in [0.5, 1)
                float cmax = 0.f;
                int bad = 0;
                for (int k = tid; k < nt; k += kConsThreads) {
                    const float ax = fabsf(rx[k] - cx), ay = fabsf(ry[k] - cy), az = fabsf(rz[k] - cz);
                    bad |= !(ax < 1e18f) | !(ay < 1e18f) | !(az < 1e18f);
                    cmax = fmaxf(cmax, fmaxf(ax, fmaxf(ay, az)));
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) cmax = fmaxf(cmax, __shfl_xor_sync(0xffffffffu, cmax, o));
                if (lane == 0) s_cstat[warp] = cmax;
                cons_bar();
                cmax = 0.f;
#pragma unroll
                for (int w = 0; w < kConsWarps; ++w) cmax = fmaxf(cmax, s_cstat[w]);
                {
                    int e = (int)((__float_as_uint(cmax) >> 23) & 0xffu);   // biased exponent; 0 for cmax == 0 / subnormal
                    e = e < 27 ? 27 : (e > 227 ? 227 : e);                   // keep s within [2^-101, 2^99]
                    cs = cmax > 0.f ? __uint_as_float((uint32_t)(253 - e) << 23) : 1.0f;   // 2^(126 - e): s*cmax in [0.5, 1)
                }
                // pass 2: scaled, split B operand
                const int npad = ((nt + kTileN - 1) / kTileN) * kTileN;
                float wmax = 0.f;
                for (int k = tid; k < npad; k += kConsThreads) {
                    uint4 v0 = make_uint4(0u, 0u, 0u, 0u), v1 = v0;
                    if (k < nt) {
                        const float x = (rx[k] - cx) * cs, y = (ry[k] - cy) * cs, z = (rz[k] - cz) * cs;
                        const float w = __fmaf_rn(z, z, __fmaf_rn(x, x, y * y));
                        bad |= !(w < 4.0f);
                        wmax = fmaxf(wmax, w);
                        unsigned short xh, xl, yh, yl, zh, zl, w1, w2, w3;
                        split_h(x, xh, xl); split_h(y, yh, yl); split_h(z, zh, zl);
                        const __half hw1 = __float2half_rn(w);
                        const float wr = w - __half2float(hw1);
                        const __half hw2 = __float2half_rn(wr);
                        const __half hw3 = __float2half_rn(wr - __half2float(hw2));
                        w1 = __half_as_ushort(hw1); w2 = __half_as_ushort(hw2); w3 = __half_as_ushort(hw3);
                        v0 = make_uint4(pack2(xh, xl), pack2(xh, yh), pack2(yl, yh), pack2(zh, zl));
                        v1 = make_uint4(pack2(zh, w1), pack2(w2, w3), 0u, 0u);
                    } else {
                        v1.x = pack2(0, __half_as_ushort(__float2half_rn(kPadW)));
                    }
                    unsigned char *dst = smem + kOffB + (k >> 3) * 256 + (k & 7) * 16;
                    *reinterpret_cast<uint4 *>(dst) = v0;
                    *reinterpret_cast<uint4 *>(dst + 128) = v1;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    wmax = fmaxf(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
                    bad |= __shfl_xor_sync(0xffffffffu, bad, o);
                }
                if (lane == 0) { s_wstat[warp] = wmax; s_bstat[warp] = bad; }
            }
            if (tid < kQB) {
                const float *sqp = sq + (ul % 3) * 3 * kQB;
                const float sc = -2.0f * cs;
                const float x = (sqp[tid] - cx) * sc, y = (sqp[kQB + tid] - cy) * sc, z = (sqp[2 * kQB + tid] - cz) * sc;
                unsigned short xh, xl, yh, yl, zh, zl;
                split_h(x, xh, xl); split_h(y, yh, yl); split_h(z, zh, zl);
                const unsigned short one = 0x3c00;
                unsigned char *dst = smem + kOffA + (tid >> 3) * 256 + (tid & 7) * 16;
                *reinterpret_cast<uint4 *>(dst) = make_uint4(pack2(xh, xh), pack2(xl, yh), pack2(yh, yl), pack2(zh, zh));
                *reinterpret_cast<uint4 *>(dst + 128) = make_uint4(pack2(zl, one), pack2(one, one), 0u, 0u);
            }
            fence_async_smem();
            return group;
        };
        // after the barrier that follows stage(): pick up the statistics of a freshly staged B operand
        auto adopt = [&](int group) {
            if (group != res_group) {
                float wmax = 0.f;
                int bad = 0;
#pragma unroll
                for (int w = 0; w < kConsWarps; ++w) { wmax = fmaxf(wmax, s_wstat[w]); bad |= s_bstat[w]; }
                res_wmax = wmax; res_bad = bad; res_group = group; res_buf = pf_buf;
            }
        };

        // exact full scan for the queries on the deferred list, one warp per query.  Reference semantics incl. NaN:
        // within a 512-target tile the first element is taken unconditionally and NaN never replaces or is replaced
        // (chamfer3D.cu:36); a tile result replaces the running result only if strictly smaller (:126).
        auto run_fallbacks = [&](int nfb) {
            for (int fi = warp; fi < nfb; fi += kConsWarps) {
                const int e = fb_list[fi];
                const bool ebad = (e >> 30) & 1;
                const Unit u = decode_unit(p, blk_begin + ((e & 0x3fffffff) >> 8));
                const NNDirection &D = p.dir[u.d];
                const int nt = D.nt;
                const int j = D.q_begin + u.qblock * kQB + (e & 0xff);
                const float *__restrict__ tb = D.t + (long long)u.cloud * D.t_bs;
                const float *__restrict__ qp = D.q + (long long)u.cloud * D.q_bs + j * D.q_ps;
                const long long tps = D.t_ps, tcs = D.t_cs;
                const float x1 = __ldg(qp), y1 = __ldg(qp + D.q_cs), z1 = __ldg(qp + 2 * D.q_cs);
                const bool nan_possible = ebad || !(fabsf(x1) < 1e18f) || !(fabsf(y1) < 1e18f) || !(fabsf(z1) < 1e18f);
                unsigned long long key = ~0ull;
                for (int kb = lane; kb < nt; kb += 8 * 32) {
                    float dd[8], dts[8];
#pragma unroll
                    for (int q8 = 0; q8 < 8; ++q8) {
                        const int k = min(kb + q8 * 32, nt - 1);
                        dd[q8] = exact_d(tb, tps, tcs, k, x1, y1, z1);
                        dts[q8] = nan_possible ? exact_d(tb, tps, tcs, k & ~(kRefTile - 1), x1, y1, z1) : 0.f;
                    }
#pragma unroll
                    for (int q8 = 0; q8 < 8; ++q8) {
                        const int k = kb + q8 * 32;
                        if (k < nt && !(dd[q8] != dd[q8]) && !(dts[q8] != dts[q8])) {
                            const unsigned long long kk = pack_key(dd[q8], k);
                            key = kk < key ? kk : key;
                        }
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const unsigned long long other = shfl_xor_u64(key, o);
                    key = other < key ? other : key;
                }
                if (lane == 0) {
                    const float d0 = exact_d(tb, tps, tcs, 0, x1, y1, z1);
                    float dres;
                    int ires;
                    if (d0 != d0) { dres = d0; ires = 0; }   // tile 0 poisoned: stays NaN, index 0
                    else { dres = __uint_as_float((unsigned int)(key >> 32)); ires = (int)(key & 0xffffffffu); }
                    D.dist[(long long)u.cloud * D.nq + j] = dres;
                    D.idx[(long long)u.cloud * D.nq + j] = ires;
                    if (p.sums) atomicAdd(p.sums + u.cloud * 2 + D.slot, dres);
                    if (p.fs_count && dres < p.fs_thr) atomicAdd(p.fs_count + u.cloud * 2 + D.slot, 1);
                }
            }
            if (nfb > 0 && tid == 0) atomicAdd(&g_fallback_queries_tc, (unsigned long long)nfb);
        };

        if (nunits > 0) {
            prefetch(0);
            cp_async_wait_all();
            cons_bar();
            const int grp = stage(0);
            cons_bar();
            adopt(grp);
            if (tid == 0) mbar_arrive(bar_ready);
        }
        if (tid == 0) stamp(2);
        int g0 = 0;
        for (int ul = 0; ul < nunits; ++ul) {
            const Unit u = decode_unit(p, blk_begin + ul);
            const NNDirection &D = p.dir[u.d];
            const int nt = D.nt;
            const int ntiles = (nt + kTileN - 1) / kTileN;
            // ---------------- raw data of the next unit: in flight during the scan
            if (ul + 1 < nunits) {
                // a target prefetch reuses the raw buffer of the previous cloud/direction, which slower warps may still
                // be reading in the resolve phase of the previous unit
                if (prefetch_needs_targets(ul + 1)) cons_bar();
                prefetch(ul + 1);
            }
            // ---------------- scan: tiles whose TMEM buffer is this warp's column group; 16-column loads, one in flight
            float best = kBig, second = kBig;
            int bchunk = 0;
            for (int t = (c - g0) & (kBufs - 1); t < ntiles; t += kBufs) {
                const int gg = g0 + t;
                const long long w0 = DBG ? clock64() : 0;
                mbar_wait(bar_full + 8 * c, (gg >> 2) & 1, s_abort);
                if (DBG && pf && tid == 0) pf[59] += clock64() - w0;
                tc_fence_after();
                uint32_t ra[16], rb[16];
                tmem_ld16(taddr0, ra);
                float mprev = 0.f;
#pragma unroll
                for (int h = 0; h < kTileN / 16; ++h) {
                    float mh;
                    if ((h & 1) == 0) {
                        tmem_wait(ra);
                        tmem_ld16(taddr0 + (uint32_t)((h + 1) * 16), rb);
                        if (DBG && dbg) {
                            float *o = dbg + ((long long)(blk_begin + ul) * kQB + row) * dbg_ld + t * kTileN + h * 16;
#pragma unroll
                            for (int i = 0; i < 16; ++i) o[i] = __uint_as_float(ra[i]);
                        }
                        mh = min16(ra);
                        mprev = mh;
                    } else {
                        tmem_wait(rb);
                        if (h + 1 < kTileN / 16) tmem_ld16(taddr0 + (uint32_t)((h + 1) * 16), ra);
                        if (DBG && dbg) {
                            float *o = dbg + ((long long)(blk_begin + ul) * kQB + row) * dbg_ld + t * kTileN + h * 16;
#pragma unroll
                            for (int i = 0; i < 16; ++i) o[i] = __uint_as_float(rb[i]);
                        }
                        mh = min16(rb);
                        const float v = fminf(mprev, mh);
                        second = fminf(second, fmaxf(best, v));
                        const bool lt = v < best;
                        best = fminf(best, v);
                        bchunk = lt ? t * (kTileN / kCh) + (h >> 1) : bchunk;
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_empty + 8 * c);
            }
            g0 += ntiles;
            if (tid == 0) stamp(8 + ul * 6);
            {   // park this warp's partial results (double-buffered by unit parity)
                float *pp = part + ((ul & 1) * 4 + c) * 3 * kQB;
                pp[row] = best; pp[kQB + row] = second; reinterpret_cast<int *>(pp)[2 * kQB + row] = bchunk;
            }
            // frame of THIS unit, before staging possibly replaces it
            const float ucx = cx, ucy = cy, ucz = cz, ucs = cs, u_wmax = res_wmax;
            const int u_bad = res_bad, u_buf = res_buf;
            cp_async_wait_all();
            cons_bar();   // S1: all tiles of the unit consumed (every MMA that reads A/B has completed), partials parked,
                          //     prefetched raw data of the next unit visible
            if (tid == 0) stamp(9 + ul * 6);
            {   // deferred exact scans: flush when the next unit could overflow the list
                const int nfb = *s_nfb;
                if (nfb > kFbCap - kQB) {
                    run_fallbacks(nfb);
                    cons_bar();
                    if (tid == 0) *s_nfb = 0;
                }
            }
            int grp = res_group;
            if (ul + 1 < nunits) grp = stage(ul + 1);
            cons_bar();   // S2: operands of the next unit are in shared memory (also orders the list reset above)
            if (ul + 1 < nunits) {
                adopt(grp);
                if (tid == 0) mbar_arrive(bar_ready);
            }
            if (tid == 0) stamp(10 + ul * 6);

            // ---------------- resolve: 4 threads per query, raw targets from shared memory
            {
                const int ql = tid >> 2, sub = tid & 3;
                const int j = D.q_begin + u.qblock * kQB + ql;
                const bool live = j < D.q_begin + D.q_count;
                const float *pp = part + (ul & 1) * 4 * 3 * kQB;
                float b1 = kBig, b2 = kBig;
                int bc = 0;
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    const float v = pp[w * 3 * kQB + ql];
                    b2 = fminf(b2, fminf(pp[w * 3 * kQB + kQB + ql], fmaxf(b1, v)));
                    if (v < b1) { b1 = v; bc = reinterpret_cast<const int *>(pp)[w * 3 * kQB + 2 * kQB + ql]; }
                }
                const float *sqp = sq + (ul % 3) * 3 * kQB;
                const float x1 = sqp[ql], y1 = sqp[kQB + ql], z1 = sqp[2 * kQB + ql];
                const float ux = (x1 - ucx) * ucs, uy = (y1 - ucy) * ucs, uz = (z1 - ucz) * ucs;   // scaled frame, like the filter
                const float qq = __fmaf_rn(uz, uz, __fmaf_rn(ux, ux, uy * uy));
                // Filter error bound E = 25u*S (u = 2^-24): frame 2u, |t-c|^2 3u, operand splits 12u, tensor-core
                // accumulation 8u (measured total on B200: <= 4.1u, tools/tc_calibrate.py).  The reference's argmin
                // lies in the best chunk if second > best + 2E + 10u*S.
                //   S  = (|q-c| + max|t-c|)^2 bounds every target,
                //   S' = (2|q-c| + rho)^2 bounds the targets that can compete (within rho of the query).
                const float qn = sqrtf(qq);
                const float rr = qn + sqrtf(u_wmax);
                const float S = rr * rr;
                const float rho = sqrtf(fmaxf(b1 + qq, 0.f) + 2.4e-6f * S);
                const float r2 = 2.0f * qn + rho;
                const float Seff = fminf(S, r2 * r2);
                const float margin = __fmaf_rn(Seff, 3.7e-6f, 1e-36f);
                const bool ok = live && !u_bad && (qn < kQMax) && (b2 > b1 + margin);
                float dres = 0.f;
                bool done = false;
                if (ok) {
                    const int k0 = bc * kCh + sub * 8;
                    float dbest = 3.0e38f;
                    int ibest = 0x7fffffff;
                    if (k0 < nt) {
                        const float4 *px = reinterpret_cast<const float4 *>(sraw + (u_buf * 3) * kMaxT + k0);
                        const float4 *py = reinterpret_cast<const float4 *>(sraw + (u_buf * 3 + 1) * kMaxT + k0);
                        const float4 *pz = reinterpret_cast<const float4 *>(sraw + (u_buf * 3 + 2) * kMaxT + k0);
                        const float4 xa = px[0], xb = px[1], ya = py[0], yb = py[1], za = pz[0], zb = pz[1];
                        const float tx[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
                        const float ty[8] = {ya.x, ya.y, ya.z, ya.w, yb.x, yb.y, yb.z, yb.w};
                        const float tz[8] = {za.x, za.y, za.z, za.w, zb.x, zb.y, zb.z, zb.w};
                        dbest = sqdist_exact(tx[0] - x1, ty[0] - y1, tz[0] - z1);
                        ibest = k0;
#pragma unroll
                        for (int i = 1; i < 8; ++i) {
                            const float dv = sqdist_exact(tx[i] - x1, ty[i] - y1, tz[i] - z1);
                            if (k0 + i < nt && dv < dbest) { dbest = dv; ibest = k0 + i; }
                        }
                    }
                    // quad merge: smaller distance wins, equal distances -> lower index (sub-ranges are index-ordered)
                    const unsigned qmask = 0xFu << (lane & ~3);
#pragma unroll
                    for (int o = 1; o <= 2; o <<= 1) {
                        const float od = __shfl_xor_sync(qmask, dbest, o);
                        const int oi = __shfl_xor_sync(qmask, ibest, o);
                        if (od < dbest || (od == dbest && oi < ibest)) { dbest = od; ibest = oi; }
                    }
                    if (sub == 0) {
                        D.dist[(long long)u.cloud * D.nq + j] = dbest;
                        D.idx[(long long)u.cloud * D.nq + j] = ibest;
                        dres = dbest;
                        done = true;
                    }
                } else if (live && sub == 0) {
                    fb_list[atomicAdd(s_nfb, 1)] = (ul << 8) | ql | (u_bad ? (1 << 30) : 0);
                }
                if (p.sums != nullptr || p.fs_count != nullptr) {   // fused epilogues, one atomic per warp
                    float ws = done ? dres : 0.f;
                    int wc = (done && dres < p.fs_thr) ? 1 : 0;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        ws += __shfl_xor_sync(0xffffffffu, ws, o);
                        wc += __shfl_xor_sync(0xffffffffu, wc, o);
                    }
                    if (lane == 0) {
                        if (p.sums) atomicAdd(p.sums + u.cloud * 2 + D.slot, ws);
                        if (p.fs_count && wc) atomicAdd(p.fs_count + u.cloud * 2 + D.slot, wc);
                    }
                }
            }
            if (tid == 0) stamp(11 + ul * 6);
        }
        // ---------------- the deferred exact scans of this CTA
        cons_bar();
        if (tid == 0) stamp(3);
        run_fallbacks(*s_nfb);
        if (tid == 0) stamp(4);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kConsWarps)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

}  // namespace tc
}  // namespace psd

using namespace psd;

// Shapes the tensor-core kernel takes: both target clouds fit the resident B operand.
bool psd_nn_tc_supported(const NNParams &p) {
    return p.dir[0].nt <= tc::kMaxT && p.dir[1].nt <= tc::kMaxT;
}

cudaError_t psd_launch_nn_tc(const NNParams &p, int num_sms, cudaStream_t stream, float *dbg, int dbg_ld, long long *prof) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(tc::chamfer_nn_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kSmemTC);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(tc::chamfer_nn_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kSmemTC);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    const int grid = p.total_blocks < num_sms ? p.total_blocks : num_sms;
    if (dbg || prof) tc::chamfer_nn_tc_kernel<true><<<grid, tc::kThreadsTC, tc::kSmemTC, stream>>>(p, dbg, dbg_ld, prof);
    else tc::chamfer_nn_tc_kernel<false><<<grid, tc::kThreadsTC, tc::kSmemTC, stream>>>(p, nullptr, 0, nullptr);
    return cudaGetLastError();
}

cudaError_t psd_read_chamfer_stats_tc(unsigned long long *fallback, int *error, int reset) {
    cudaError_t e = cudaMemcpyFromSymbol(fallback, tc::g_fallback_queries_tc, sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemcpyFromSymbol(error, tc::g_tc_error, sizeof(int));
    if (e != cudaSuccess) return e;
    if (reset) {
        const unsigned long long z = 0;
        const int zi = 0;
        e = cudaMemcpyToSymbol(tc::g_fallback_queries_tc, &z, sizeof(z));
        if (e == cudaSuccess) e = cudaMemcpyToSymbol(tc::g_tc_error, &zi, sizeof(zi));
    }
    return e;
}
