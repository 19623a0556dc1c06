// chamfer_nn_grouped.cu -- Chamfer NN forward for launches that fill the GPU (>= 2 blocks of 128 queries per
// sub-partition group): same filter + exact-rescan mathematics as chamfer.cu (results bit-identical to the
// reference's NmDistanceKernel, metric/chamfer3D/chamfer3D.cu:12-134), different execution structure.
//
// Why a second structure: in the shared-block kernel all 16 warps of the CTA scan the same 128-query block and
// the latency-bound phases (tile staging, merge of the warp partials, exact rescan, exact fallback) run with the
// FMA pipes idle -- 28 % of the kernel at B=32, N=M=2048 (profiles/r1_chamfer_nn_summary.md).  Here
//   * the CTA (one per SM, persistent) is split into 4 independent GROUPS of 4 warps, one warp on each SM
//     sub-partition per group.  A group owns a whole 128-query block: its 4 warps hold the same 128 queries
//     (4 per lane), split the target chunks 4 ways, merge through 6 KB of shared memory behind a 128-thread
//     named barrier and resolve the block (margin test, exact rescan of the winning 16-target chunk, exact
//     fallback) on their own.  Groups drift out of phase, so while one group resolves, the other three keep
//     every sub-partition's dispatch port busy with filter FMAs.
//   * target tiles live in a 2-slot ring in shared memory, each slot holding the filter's SoA frame
//     (x-c, y-c, z-c, |t-c|^2) AND the raw AoS coordinates, so the exact rescan and the exact fallback read
//     shared memory instead of L2.  The first two tiles are staged by all warps at kernel start; later tiles
//     are staged in the background by a producer warp group (warps 16-19, registers handed to the compute
//     warps with setmaxnreg), hand-shaken through ready/done words.
//   * clouds larger than one tile (nt > 2048) stream their tiles through the ring; the 4 groups of a round
//     consume each tile once (mode M below), rescans then read global memory like the shared-block kernel.
#include "chamfer_nn.cuh"

namespace psd {

constexpr int kTM = 2048;                          // targets per shared-memory tile
constexpr int kGroups = 4;                         // independent groups per CTA
constexpr int kGW = 4;                             // warps per group (one per SM sub-partition)
constexpr int kCompWarps = kGroups * kGW;          // 16 compute warps
constexpr int kProdWarps = 4;                      // producer warp group (one warp per sub-partition)
constexpr int kGThreads = (kCompWarps + kProdWarps) * 32;
constexpr int kCompRegs = 112, kProdRegs = 64;     // setmaxnreg split: 4*112 + 64 = 512 = one sub-partition's registers / 32
constexpr int kTileFloats = 7 * kTM;               // SoA x,y,z,w + raw AoS xyz
constexpr int kSlots = 2;
constexpr size_t kGroupedSmem = (size_t)kSlots * kTileFloats * 4 + 3 * (size_t)kCompWarps * kQB * 4 + (size_t)kGroups * kQB * 4;

__device__ unsigned long long g_fallback_queries_grouped = 0ull;

#ifdef PSD_PROFILE_CLOCKS   // in-kernel phase clocks (experimental builds only, tools/nn_phase_clocks.py)
__device__ long long g_prof[148 * 20 * 8];
#define PROF_T(v) const long long v = clock64()
#define PROF_ADD(slot, a, b) prof[slot] += (b) - (a)
#else
#define PROF_T(v)
#define PROF_ADD(slot, a, b)
#endif

struct TileInfo {
    float cx, cy, cz;   // centre of the filter frame (a property of the cloud, identical for all its tiles)
    float wmax;         // max |t-c|^2 over the tile
    int bad;            // non-finite / huge target in the tile
    int pad[3];
};

struct Job {            // one tile to stage: targets [t0, t0+cnt) of cloud `cloud`, direction d
    int d, cloud, t0, cnt;
    int expected;       // consumer warps that will release it
};

// Every warp enumerates the CTA's tile jobs with the same arithmetic, so no job table has to be communicated.
//  mode S (nt <= kTM): one job per (direction, cloud) segment, consumed by all the CTA's blocks of that segment.
//  mode M (nt >  kTM): blocks are taken in rounds of <= 4 (one per group); every round streams all T tiles.
struct JobIter {
    int i, b1;
    int d, cloud, seg_end, nt, T, r, t;
    bool in_seg;
    __device__ void init(int b0, int b1_) { i = b0; b1 = b1_; in_seg = false; d = cloud = seg_end = nt = T = r = t = 0; }
    __device__ bool next(const NNParams &p, Job &J) {
        if (!in_seg) {
            if (i >= b1) return false;
            d = i >= p.blocks_dir0 ? 1 : 0;
            const int bid = d ? i - p.blocks_dir0 : i;
            const int qbn = p.dir[d].qblocks;
            cloud = bid / qbn;
            const int qblock = bid - cloud * qbn;
            seg_end = min(b1, i + (qbn - qblock));
            nt = p.dir[d].nt;
            T = (nt + kTM - 1) / kTM;
            r = i; t = 0; in_seg = true;
        }
        J.d = d; J.cloud = cloud;
        if (T == 1) {
            J.t0 = 0; J.cnt = nt; J.expected = (seg_end - i) * kGW;
            i = seg_end; in_seg = false;
            return true;
        }
        const int nr = min(kGroups, seg_end - r);
        J.t0 = t * kTM; J.cnt = min(kTM, nt - J.t0); J.expected = nr * kGW;
        if (++t == T) {
            t = 0; r += kGroups;
            if (r >= seg_end) { i = seg_end; in_seg = false; }
        }
        return true;
    }
};

__device__ __forceinline__ int ld_volatile_s32(const int *p) { return *reinterpret_cast<const volatile int *>(p); }
__device__ __forceinline__ void st_volatile_s32(int *p, int v) { *reinterpret_cast<volatile int *>(p) = v; }

__device__ __forceinline__ void group_barrier(int g) {   // named barriers 1..5, 128 threads each (5 = producers)
    asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "n"(kGW * 32) : "memory");
}

// Stage one tile: SoA filter frame + raw AoS copy + statistics, by `lcount` threads (whole warps).  The caller
// zeroes ti->wmax / ti->bad beforehand and barriers afterwards (statistics are merged with shared-memory atomics).
__device__ __forceinline__ void stage_tile(const NNParams &p, const Job &J, float *tile, TileInfo *ti, int ltid, int lcount) {
    constexpr int C = kChunk;
    const NNDirection &D = p.dir[J.d];
    const int nt = D.nt;
    const float *__restrict__ tb = D.t + (long long)J.cloud * D.t_bs;
    const long long tps = D.t_ps, tcs = D.t_cs;
    float *sX = tile, *sY = tile + kTM, *sZ = tile + 2 * kTM, *sW = tile + 3 * kTM, *raw = tile + 4 * kTM;
    const int cnt = J.cnt, t0 = J.t0;
    const int nchunks = (cnt + C - 1) / C;
    // centre of the filter frame: mean of up to 8 evenly spaced targets of the whole cloud.  Any value is
    // correct (results come from the exact formula); a centred frame only keeps the margin small.
    float cx, cy, cz;
    {
        float sx = 0.f, sy = 0.f, sz = 0.f;
        const int ns = nt < 8 ? nt : 8;
        const int step = nt >> 3;
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            if (s < ns) {
                const long long k = nt < 8 ? s : s * step;
                sx += __ldg(tb + k * tps);
                sy += __ldg(tb + k * tps + tcs);
                sz += __ldg(tb + k * tps + 2 * tcs);
            }
        }
        const float inv = 1.0f / (float)ns;
        cx = sx * inv; cy = sy * inv; cz = sz * inv;
    }
    float wmax = 0.f;
    int bad = 0;
    const bool vec_ok = (tps == 3) && (tcs == 1) && ((reinterpret_cast<unsigned long long>(tb) & 15ull) == 0ull);
    if (vec_ok && (cnt & 7) == 0) {
        // AoS fast path: 6 x LDG.128 = 8 whole points per thread and round (t0 is a multiple of kTM: aligned)
#pragma unroll 2
        for (int g8 = ltid; g8 * 8 < cnt; g8 += lcount) {
            const float4 *src = reinterpret_cast<const float4 *>(tb + (long long)t0 * 3) + g8 * 6;
            float4 v[6];
#pragma unroll
            for (int i = 0; i < 6; ++i) v[i] = __ldg(src + i);
            float4 *dr = reinterpret_cast<float4 *>(raw) + g8 * 6;
            float f[24];
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                dr[i] = v[i];
                f[4 * i + 0] = v[i].x; f[4 * i + 1] = v[i].y; f[4 * i + 2] = v[i].z; f[4 * i + 3] = v[i].w;
            }
            float xs[8], ys[8], zs[8], ws[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                xs[i] = f[3 * i + 0] - cx;
                ys[i] = f[3 * i + 1] - cy;
                zs[i] = f[3 * i + 2] - cz;
                ws[i] = __fmaf_rn(zs[i], zs[i], __fmaf_rn(xs[i], xs[i], ys[i] * ys[i]));
                bad |= !(ws[i] < kLimit);
                wmax = fmaxf(wmax, ws[i]);
            }
            float4 *dX = reinterpret_cast<float4 *>(sX) + g8 * 2, *dY = reinterpret_cast<float4 *>(sY) + g8 * 2;
            float4 *dZ = reinterpret_cast<float4 *>(sZ) + g8 * 2, *dW = reinterpret_cast<float4 *>(sW) + g8 * 2;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                dX[h] = make_float4(xs[4 * h], xs[4 * h + 1], xs[4 * h + 2], xs[4 * h + 3]);
                dY[h] = make_float4(ys[4 * h], ys[4 * h + 1], ys[4 * h + 2], ys[4 * h + 3]);
                dZ[h] = make_float4(zs[4 * h], zs[4 * h + 1], zs[4 * h + 2], zs[4 * h + 3]);
                dW[h] = make_float4(ws[4 * h], ws[4 * h + 1], ws[4 * h + 2], ws[4 * h + 3]);
            }
        }
        for (int k = cnt + ltid; k < nchunks * C; k += lcount) { sX[k] = 0.f; sY[k] = 0.f; sZ[k] = 0.f; sW[k] = kBig; }
    } else {
        // generic strides / ragged tile: 4 points per thread per round, loads issued before use
        for (int k0 = ltid; k0 < nchunks * C; k0 += 4 * lcount) {
            float lx[4], ly[4], lz[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int k = k0 + i * lcount;
                const float *tp = tb + (long long)(t0 + (k < cnt ? k : 0)) * tps;
                lx[i] = __ldg(tp); ly[i] = __ldg(tp + tcs); lz[i] = __ldg(tp + 2 * tcs);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int k = k0 + i * lcount;
                if (k < nchunks * C) {
                    float x = 0.f, y = 0.f, z = 0.f, w = kBig;
                    if (k < cnt) {
                        raw[3 * k] = lx[i]; raw[3 * k + 1] = ly[i]; raw[3 * k + 2] = lz[i];
                        x = lx[i] - cx; y = ly[i] - cy; z = lz[i] - cz;
                        w = __fmaf_rn(z, z, __fmaf_rn(x, x, y * y));
                        bad |= !(w < kLimit);
                        wmax = fmaxf(wmax, w);
                    }
                    sX[k] = x; sY[k] = y; sZ[k] = z; sW[k] = w;
                }
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        wmax = fmaxf(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
        bad |= __shfl_xor_sync(0xffffffffu, bad, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(reinterpret_cast<int *>(&ti->wmax), __float_as_int(wmax));  // wmax >= 0: int order == float order
        if (bad) atomicOr(&ti->bad, 1);
        if (ltid == 0) { ti->cx = cx; ti->cy = cy; ti->cz = cz; }
    }
}

__global__ void __launch_bounds__(kGThreads, 1) chamfer_nn_grouped_kernel(const NNParams p) {
    constexpr int Q = kQ, QB = kQB, C = kChunk;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *tiles = reinterpret_cast<float *>(smem_raw);                 // [kSlots][7*kTM]
    float *part_best = tiles + kSlots * kTileFloats;                    // [16 warps][QB]
    float *part_second = part_best + kCompWarps * QB;
    int *part_chunk = reinterpret_cast<int *>(part_second + kCompWarps * QB);
    int *fb_list = part_chunk + kCompWarps * QB;                        // [kGroups][QB]
    __shared__ TileInfo tinfo[kSlots];
    __shared__ int s_ready[kSlots];    // job id staged in the slot
    __shared__ int s_done[kSlots];     // consumer warps that released the staged job
    __shared__ int s_nfb[kGroups][2];  // fallback-list lengths, double-buffered by block parity

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x;
#ifdef PSD_PROFILE_CLOCKS
    long long prof[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // 0 total, 1 start-up, 2 tile wait, 3 scan, 4 scan iterations, 5 resolve, 6 fallback, 7 globaltimer ns
    long long gt0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt0));
#endif
    PROF_T(k_begin);
    const int b0 = (int)(((long long)blockIdx.x * p.total_blocks) / G);
    const int b1 = (int)(((long long)(blockIdx.x + 1) * p.total_blocks) / G);

    JobIter it;
    it.init(b0, b1);
    Job J0, J1;
    const bool has0 = it.next(p, J0);
    const bool has1 = it.next(p, J1);
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kSlots; ++s) {
            tinfo[s].cx = tinfo[s].cy = tinfo[s].cz = 0.f; tinfo[s].wmax = 0.f; tinfo[s].bad = 0;
            s_done[s] = 0;
        }
        s_ready[0] = has0 ? 0 : -1;
        s_ready[1] = has1 ? 1 : -1;
#pragma unroll
        for (int g = 0; g < kGroups; ++g) { s_nfb[g][0] = 0; s_nfb[g][1] = 0; }
    }
    __syncthreads();
    // the first two tiles are staged by everybody: warps 0-7 -> slot 0, warps 8-15 -> slot 1 (one latency period)
    if (warp < 8) {
        if (has0) stage_tile(p, J0, tiles, &tinfo[0], tid, 256);
    } else if (warp < 16) {
        if (has1) stage_tile(p, J1, tiles + kTileFloats, &tinfo[1], tid - 256, 256);
    }
    __syncthreads();

    // =============================================================== producer warps: background tile staging
    if (warp >= kCompWarps) {
#ifdef PSD_SETMAXNREG
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kProdRegs));
#endif
        const int ptid = tid - kCompWarps * 32;
        int exp_prev0 = has0 ? J0.expected : 0, exp_prev1 = has1 ? J1.expected : 0;
        int j = 2;
        Job J;
        while (it.next(p, J)) {
            const int s = j & 1;
            const int need = s ? exp_prev1 : exp_prev0;
            while (ld_volatile_s32(&s_done[s]) < need) __nanosleep(64);   // consumers of the slot's previous job
            __threadfence_block();
            group_barrier(kGroups);
            if (ptid == 0) {
                st_volatile_s32(&s_done[s], 0);   // nobody adds before the new job is published
                tinfo[s].wmax = 0.f; tinfo[s].bad = 0;
            }
            group_barrier(kGroups);
            stage_tile(p, J, tiles + s * kTileFloats, &tinfo[s], ptid, kProdWarps * 32);
            __threadfence_block();
            group_barrier(kGroups);
            if (ptid == 0) st_volatile_s32(&s_ready[s], j);
            if (s) exp_prev1 = J.expected; else exp_prev0 = J.expected;
            ++j;
        }
        return;
    }
#ifdef PSD_SETMAXNREG
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kCompRegs));
#endif

    PROF_T(k_started);
    PROF_ADD(1, k_begin, k_started);
    // =============================================================== compute groups
    const int g = warp >> 2, wg = warp & 3, gt = tid & (kGW * 32 - 1);
    float *pb = part_best + warp * QB, *ps = part_second + warp * QB;
    int *pc = part_chunk + warp * QB;
    int *my_fb = fb_list + g * QB;
    int parity = 0;
    unsigned int my_fallbacks = 0;

    int i = b0, jbase = 0, rr = 0;
    while (i < b1) {
        const int d = i >= p.blocks_dir0 ? 1 : 0;
        const NNDirection &D = p.dir[d];
        const int bid = d ? i - p.blocks_dir0 : i;
        const int qbn = D.qblocks;
        const int cloud = bid / qbn;
        const int qblock0 = bid - cloud * qbn;
        const int seg_end = min(b1, i + (qbn - qblock0));
        const int nt = D.nt;
        const int T = (nt + kTM - 1) / kTM;
        const bool single = (T == 1);
        const int round_len = single ? (seg_end - i) : kGroups;
        const float *__restrict__ tb = D.t + (long long)cloud * D.t_bs;
        const long long tps = D.t_ps, tcs = D.t_cs;
        const float *__restrict__ qbase = D.q + (long long)cloud * D.q_bs;
        const int q_last = D.q_begin + D.q_count - 1;

        for (int r = i; r < seg_end; r += round_len) {
            const int nr = min(round_len, seg_end - r);
            for (int k = r + ((g - rr) & 3); k < r + nr; k += kGroups) {
                // ------------------------------------------------------------------ one 128-query block
                const int qblock = qblock0 + (k - i);
                float qx[Q], qy[Q], qz[Q];
#pragma unroll
                for (int u = 0; u < Q; ++u) {   // raw queries first (global latency overlaps the tile wait)
                    int j = D.q_begin + qblock * QB + u * 32 + lane;
                    j = j < q_last ? j : q_last;
                    const float *qp = qbase + (long long)j * D.q_ps;
                    qx[u] = __ldg(qp); qy[u] = __ldg(qp + D.q_cs); qz[u] = __ldg(qp + 2 * D.q_cs);
                }
                float best[Q], second[Q];
                int bchunk[Q];
#pragma unroll
                for (int u = 0; u < Q; ++u) { best[u] = kBig; second[u] = kBig; bchunk[u] = 0; }
                float cx = 0.f, cy = 0.f, cz = 0.f, blk_wmax = 0.f;
                int blk_bad = 0;

                for (int t = 0; t < T; ++t) {
                    const int jj = jbase + t, s = jj & 1;
                    PROF_T(w0);
                    while (ld_volatile_s32(&s_ready[s]) != jj) __nanosleep(32);
                    __threadfence_block();
                    PROF_T(w1);
                    PROF_ADD(2, w0, w1);
                    const TileInfo ti = tinfo[s];
                    if (t == 0) {
                        cx = ti.cx; cy = ti.cy; cz = ti.cz;
#pragma unroll
                        for (int u = 0; u < Q; ++u) {   // -2 (q - c)
                            qx[u] = -2.0f * (qx[u] - cx); qy[u] = -2.0f * (qy[u] - cy); qz[u] = -2.0f * (qz[u] - cz);
                        }
                    }
                    blk_wmax = fmaxf(blk_wmax, ti.wmax);
                    blk_bad |= ti.bad;
                    const int t0 = t * kTM;
                    const int cnt = min(kTM, nt - t0);
                    const int nchunks = (cnt + C - 1) / C;
                    // software-pipelined over 4-target groups: the next group's four LDS.128 are issued before the
                    // current group's 24 FFMA2, also across chunk boundaries.
                    const float4 *T4 = reinterpret_cast<const float4 *>(tiles + s * kTileFloats);
                    constexpr int OY = kTM / 4, OZ = 2 * (kTM / 4), OW = 3 * (kTM / 4);
                    int c = wg;
                    float4 X, Y, Z, W;
                    if (c < nchunks) {
                        const float4 *g0 = T4 + c * (C / 4);
                        X = g0[0]; Y = g0[OY]; Z = g0[OZ]; W = g0[OW];
                    }
                    while (c < nchunks) {
                        float cm[Q];
                        const int cn = c + kGW;
                        const float4 *gp = T4 + c * (C / 4);
                        const float4 *gn = T4 + cn * (C / 4);
#pragma unroll
                        for (int gq = 0; gq < C / 4; ++gq) {
                            float4 Xn = X, Yn = Y, Zn = Z, Wn = W;
                            if (gq + 1 < C / 4) {
                                Xn = gp[gq + 1]; Yn = gp[gq + 1 + OY]; Zn = gp[gq + 1 + OZ]; Wn = gp[gq + 1 + OW];
                            } else if (cn < nchunks) {
                                Xn = gn[0]; Yn = gn[OY]; Zn = gn[OZ]; Wn = gn[OW];
                            }
#pragma unroll
                            for (int u = 0; u < Q; ++u) {
                                float2 a01 = ffma2(qz[u], make_float2(Z.x, Z.y), make_float2(W.x, W.y));
                                float2 a23 = ffma2(qz[u], make_float2(Z.z, Z.w), make_float2(W.z, W.w));
                                a01 = ffma2(qy[u], make_float2(Y.x, Y.y), a01);
                                a23 = ffma2(qy[u], make_float2(Y.z, Y.w), a23);
                                a01 = ffma2(qx[u], make_float2(X.x, X.y), a01);
                                a23 = ffma2(qx[u], make_float2(X.z, X.w), a23);
                                if (gq == 0) {
                                    cm[u] = fminf(fmin3(a01.x, a01.y, a23.x), a23.y);
                                } else {
                                    cm[u] = fmin3(cm[u], a01.x, a01.y);
                                    cm[u] = fmin3(cm[u], a23.x, a23.y);
                                }
                            }
                            X = Xn; Y = Yn; Z = Zn; W = Wn;
                        }
                        const int gchunk = t0 / C + c;
#pragma unroll
                        for (int u = 0; u < Q; ++u) {
                            const float v = cm[u];
                            second[u] = fminf(second[u], fmaxf(best[u], v));
                            const bool lt = v < best[u];
                            best[u] = fminf(best[u], v);
                            bchunk[u] = lt ? gchunk : bchunk[u];
                        }
                        c = cn;
#ifdef PSD_PROFILE_CLOCKS
                        prof[4] += 1;
#endif
                    }
                    PROF_T(w2);
                    PROF_ADD(3, w1, w2);
                    if (!single) {   // streamed tile: hand the slot back as soon as this warp has scanned it
                        __syncwarp();
                        if (lane == 0) { __threadfence_block(); atomicAdd(&s_done[s], 1); }
                    }
                }
                // ---- park this warp's partial results; the group merges them behind its own barrier
#pragma unroll
                for (int u = 0; u < Q; ++u) {
                    pb[u * 32 + lane] = best[u];
                    ps[u * 32 + lane] = second[u];
                    pc[u * 32 + lane] = bchunk[u];
                }
                PROF_T(r0);
                group_barrier(g);
                const int sres = jbase & 1;               // slot of the (only) tile in mode S
                if (gt == 0) s_nfb[g][parity ^ 1] = 0;   // the other parity's readers are behind this barrier
                const float *rawt = tiles + sres * kTileFloats + 4 * kTM;
                {
                    // ---- resolve: thread gt owns query gt of the block
                    const int ql = gt;
                    const int j = D.q_begin + qblock * QB + ql;
                    const bool live = j <= q_last;
                    float dres = 0.f;
                    bool done = false;
                    if (live) {
                        float m1 = kBig, m2 = kBig;
                        int bc = 0;
#pragma unroll
                        for (int w = 0; w < kGW; ++w) {
                            const float v = part_best[(g * kGW + w) * QB + ql];
                            m2 = fminf(m2, fminf(part_second[(g * kGW + w) * QB + ql], fmaxf(m1, v)));
                            if (v < m1) { m1 = v; bc = part_chunk[(g * kGW + w) * QB + ql]; }
                        }
                        const float *qp = qbase + (long long)j * D.q_ps;
                        const float x1 = __ldg(qp), y1 = __ldg(qp + D.q_cs), z1 = __ldg(qp + 2 * D.q_cs);
                        const float ux = x1 - cx, uy = y1 - cy, uz = z1 - cz;
                        const float qq = __fmaf_rn(uz, uz, __fmaf_rn(ux, ux, uy * uy));
                        // margin = 26u * min(S, S') with 15% slack (u = 2^-24; derivation in DESIGN.md):
                        //   S  = (|q-c| + max|t-c|)^2 bounds every target,
                        //   S' = (2|q-c| + rho)^2 bounds the targets that can compete (within rho of the query).
                        const float qn = sqrtf(qq);
                        const float rs = qn + sqrtf(blk_wmax);
                        const float S = rs * rs;
                        const float rho = sqrtf(fmaxf(m1 + qq, 0.f) + 2.4e-6f * S);
                        const float r2 = 2.0f * qn + rho;
                        const float Seff = fminf(S, r2 * r2);
                        const float margin = __fmaf_rn(Seff, 1.8e-6f, 1e-36f);
                        const bool ok = !blk_bad && (S < 4.0f * kLimit) && (m2 > m1 + margin);
                        if (ok) {
                            const int k0 = bc * C;
                            const int k1 = min(k0 + C, nt);
                            float dv[C];
                            if (single) {
                                // winning chunk from the raw copy in shared memory (192 B, 16-byte aligned)
                                const float4 *src = reinterpret_cast<const float4 *>(rawt + k0 * 3);
                                float f[3 * C];
#pragma unroll
                                for (int u = 0; u < 3 * C / 4; ++u) {
                                    const float4 v = src[u];
                                    f[4 * u] = v.x; f[4 * u + 1] = v.y; f[4 * u + 2] = v.z; f[4 * u + 3] = v.w;
                                }
#pragma unroll
                                for (int u = 0; u < C; ++u) dv[u] = sqdist_exact(f[3 * u] - x1, f[3 * u + 1] - y1, f[3 * u + 2] - z1);
                            } else {
#pragma unroll
                                for (int u = 0; u < C; ++u) dv[u] = exact_d(tb, tps, tcs, min(k0 + u, k1 - 1), x1, y1, z1);
                            }
                            float dbest = dv[0];
                            int ibest = k0;
#pragma unroll
                            for (int u = 1; u < C; ++u) {
                                if (k0 + u < k1 && dv[u] < dbest) { dbest = dv[u]; ibest = k0 + u; }
                            }
                            D.dist[(long long)cloud * D.nq + j] = dbest;
                            D.idx[(long long)cloud * D.nq + j] = ibest;
                            dres = dbest;
                            done = true;
                        } else {
                            my_fb[atomicAdd(&s_nfb[g][parity], 1)] = ql;
                        }
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
                            if (p.sums) atomicAdd(p.sums + cloud * 2 + D.slot, ws);
                            if (p.fs_count && wc) atomicAdd(p.fs_count + cloud * 2 + D.slot, wc);
                        }
                    }
                }
                group_barrier(g);
                PROF_T(r1);
                PROF_ADD(5, r0, r1);
                // ---- exact full scan for the flagged queries, one warp per query.  Reference semantics incl. NaN:
                // within a 512-target tile the first element is taken unconditionally and NaN never replaces or is
                // replaced (chamfer3D.cu:36); a tile result replaces the running result only if strictly smaller (:126).
                const int nfb = s_nfb[g][parity];
                for (int f = wg; f < nfb; f += kGW) {
                    const int ql = my_fb[f];
                    const int j = D.q_begin + qblock * QB + ql;
                    const float *qp = qbase + (long long)j * D.q_ps;
                    const float x1 = __ldg(qp), y1 = __ldg(qp + D.q_cs), z1 = __ldg(qp + 2 * D.q_cs);
                    const bool nan_possible = blk_bad || !(fabsf(x1) < 1e18f) || !(fabsf(y1) < 1e18f) || !(fabsf(z1) < 1e18f);
                    unsigned long long key = ~0ull;
                    for (int kb = lane; kb < nt; kb += 4 * 32) {
                        float dd[4], dts[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int kk = min(kb + u * 32, nt - 1);
                            if (single) {
                                dd[u] = sqdist_exact(rawt[3 * kk] - x1, rawt[3 * kk + 1] - y1, rawt[3 * kk + 2] - z1);
                                const int kt = kk & ~(kRefTile - 1);
                                dts[u] = nan_possible ? sqdist_exact(rawt[3 * kt] - x1, rawt[3 * kt + 1] - y1, rawt[3 * kt + 2] - z1) : 0.f;
                            } else {
                                dd[u] = exact_d(tb, tps, tcs, kk, x1, y1, z1);
                                dts[u] = nan_possible ? exact_d(tb, tps, tcs, kk & ~(kRefTile - 1), x1, y1, z1) : 0.f;
                            }
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int kk = kb + u * 32;
                            if (kk < nt && !(dd[u] != dd[u]) && !(dts[u] != dts[u])) {
                                const unsigned long long key2 = pack_key(dd[u], kk);
                                key = key2 < key ? key2 : key;
                            }
                        }
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        const unsigned long long other = shfl_xor_u64(key, o);
                        key = other < key ? other : key;
                    }
                    if (lane == 0) {
                        const float d0 = single ? sqdist_exact(rawt[0] - x1, rawt[1] - y1, rawt[2] - z1) : exact_d(tb, tps, tcs, 0, x1, y1, z1);
                        float dres;
                        int ires;
                        if (d0 != d0) { dres = d0; ires = 0; }   // tile 0 poisoned: stays NaN, index 0
                        else { dres = __uint_as_float((unsigned int)(key >> 32)); ires = (int)(key & 0xffffffffu); }
                        D.dist[(long long)cloud * D.nq + j] = dres;
                        D.idx[(long long)cloud * D.nq + j] = ires;
                        if (p.sums) atomicAdd(p.sums + cloud * 2 + D.slot, dres);
                        if (p.fs_count && dres < p.fs_thr) atomicAdd(p.fs_count + cloud * 2 + D.slot, 1);
                        ++my_fallbacks;
                    }
                }
                PROF_T(r2);
                PROF_ADD(6, r1, r2);
                if (single) {   // resident tile: released once the block no longer reads its raw copy
                    __syncwarp();
                    if (lane == 0) { __threadfence_block(); atomicAdd(&s_done[sres], 1); }
                }
                parity ^= 1;
            }
            rr += nr;
            jbase += T;
        }
        i = seg_end;
    }
    if (lane == 0 && my_fallbacks) atomicAdd(&g_fallback_queries_grouped, (unsigned long long)my_fallbacks);
#ifdef PSD_PROFILE_CLOCKS
    {
        PROF_T(k_end);
        PROF_ADD(0, k_begin, k_end);
        long long gt1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt1));
        prof[7] = gt1 - gt0;
        if (lane == 0 && blockIdx.x < 148) {
#pragma unroll
            for (int q = 0; q < 8; ++q) g_prof[(blockIdx.x * 20 + warp) * 8 + q] = prof[q];
        }
    }
#endif
}

}  // namespace psd

using namespace psd;

cudaError_t psd_launch_nn_grouped(const NNParams &p, int num_sms, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(chamfer_nn_grouped_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGroupedSmem);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    const int grid = p.total_blocks < num_sms ? p.total_blocks : num_sms;
    chamfer_nn_grouped_kernel<<<grid, kGThreads, kGroupedSmem, stream>>>(p);
    return cudaGetLastError();
}

#ifdef PSD_PROFILE_CLOCKS
extern "C" int psd_debug_read_prof(long long *host) {
    return cudaMemcpyFromSymbol(host, g_prof, sizeof(long long) * 148 * 20 * 8) == cudaSuccess;
}
#endif

cudaError_t psd_read_chamfer_stats_grouped(unsigned long long *fallback, int reset) {
    cudaError_t e = cudaMemcpyFromSymbol(fallback, g_fallback_queries_grouped, sizeof(unsigned long long));
    if (e != cudaSuccess) return e;
    if (reset) {
        const unsigned long long z = 0;
        e = cudaMemcpyToSymbol(g_fallback_queries_grouped, &z, sizeof(z));
    }
    return e;
}
