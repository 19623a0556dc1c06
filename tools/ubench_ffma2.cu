// FFMA2 operand-pattern microbenchmark (sm_100a): does the scalar-broadcast / accumulator-chain form used by the
// chamfer filter sustain one FFMA2 per 2 cycles per sub-partition?
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
constexpr int ITERS = 2048;

__device__ __forceinline__ float2 f2(float a, float2 b, float2 c) { return __ffma2_rn(make_float2(a, a), b, c); }
__device__ __forceinline__ float2 f2p(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(float* out, long long* cycles, const float4* __restrict__ tin, float seed) {
    float q[12];
    float2 acc[8];
    float2 qq[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) { q[i] = seed * (i + 1) + threadIdx.x * 1e-3f; qq[i] = make_float2(q[i], q[i]); }
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = make_float2(seed * i, seed * (i + 8));
    float4 T0 = tin[0], T1 = tin[1], T2 = tin[2], T3 = tin[3];
    float m[4] = {1e30f, 1e30f, 1e30f, 1e30f};
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
        if (MODE == 0) {  // real pattern: 4 queries x (3-deep chain) x 2 target pairs, scalar-broadcast q
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float2 a01 = f2(q[3 * i + 2], make_float2(T2.x, T2.y), make_float2(T3.x, T3.y));
                float2 a23 = f2(q[3 * i + 2], make_float2(T2.z, T2.w), make_float2(T3.z, T3.w));
                a01 = f2(q[3 * i + 1], make_float2(T1.x, T1.y), a01);
                a23 = f2(q[3 * i + 1], make_float2(T1.z, T1.w), a23);
                a01 = f2(q[3 * i], make_float2(T0.x, T0.y), a01);
                a23 = f2(q[3 * i], make_float2(T0.z, T0.w), a23);
                asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(m[i]) : "f"(a01.x), "f"(a01.y));
                asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(m[i]) : "f"(a23.x), "f"(a23.y));
            }
            T0.x += 1.0f; T1.y += 1.0f; T2.z += 1.0f; T3.w += 1.0f;  // keep the loop from being hoisted
        } else if (MODE == 1) {  // same with explicit 64-bit q pairs (no scalar broadcast)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float2 a01 = f2p(qq[3 * i + 2], make_float2(T2.x, T2.y), make_float2(T3.x, T3.y));
                float2 a23 = f2p(qq[3 * i + 2], make_float2(T2.z, T2.w), make_float2(T3.z, T3.w));
                a01 = f2p(qq[3 * i + 1], make_float2(T1.x, T1.y), a01);
                a23 = f2p(qq[3 * i + 1], make_float2(T1.z, T1.w), a23);
                a01 = f2p(qq[3 * i], make_float2(T0.x, T0.y), a01);
                a23 = f2p(qq[3 * i], make_float2(T0.z, T0.w), a23);
                asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(m[i]) : "f"(a01.x), "f"(a01.y));
                asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(m[i]) : "f"(a23.x), "f"(a23.y));
            }
            T0.x += 1.0f; T1.y += 1.0f; T2.z += 1.0f; T3.w += 1.0f;
#pragma unroll
            for (int i = 0; i < 12; ++i) qq[i].y = qq[i].x;
        } else if (MODE == 2) {  // scalar FFMA version of the same math (48 FFMA + 16 FMNMX per iteration)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float a0 = fmaf(q[3 * i + 2], T2.x, T3.x), a1 = fmaf(q[3 * i + 2], T2.y, T3.y);
                float a2 = fmaf(q[3 * i + 2], T2.z, T3.z), a3 = fmaf(q[3 * i + 2], T2.w, T3.w);
                a0 = fmaf(q[3 * i + 1], T1.x, a0); a1 = fmaf(q[3 * i + 1], T1.y, a1);
                a2 = fmaf(q[3 * i + 1], T1.z, a2); a3 = fmaf(q[3 * i + 1], T1.w, a3);
                a0 = fmaf(q[3 * i], T0.x, a0); a1 = fmaf(q[3 * i], T0.y, a1);
                a2 = fmaf(q[3 * i], T0.z, a2); a3 = fmaf(q[3 * i], T0.w, a3);
                asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(m[i]) : "f"(a0), "f"(a1));
                asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(m[i]) : "f"(a2), "f"(a3));
            }
            T0.x += 1.0f; T1.y += 1.0f; T2.z += 1.0f; T3.w += 1.0f;
        } else if (MODE == 3) {  // 8 independent accumulator chains, one shared target pair: acc = q_i * T + acc
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] = f2(q[i], make_float2(T0.x, T0.y), acc[i]);
        }
    }
    long long t1 = clock64();
    float s = m[0] + m[1] + m[2] + m[3] + T0.x + T1.y + T2.z + T3.w;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
#pragma unroll
    for (int i = 0; i < 12; ++i) s += qq[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int threads, double pipe_cycles_per_iter) {
    int grid = 148;
    float* out; long long* cyc; float4* tin;
    CK(cudaMalloc(&out, sizeof(float) * grid * 512));
    CK(cudaMalloc(&cyc, sizeof(long long) * grid));
    CK(cudaMalloc(&tin, sizeof(float4) * 4));
    CK(cudaMemset(tin, 0, sizeof(float4) * 4));
    for (int w = 0; w < 2; ++w) k<MODE><<<grid, threads>>>(out, cyc, tin, 1.0001f);
    CK(cudaDeviceSynchronize());
    long long h[148];
    CK(cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
    double avg = 0; for (int i = 0; i < grid; ++i) avg += (double)h[i]; avg /= grid;
    double warps_per_smsp = threads / 32 / 4.0;
    double cyc_per_iter_warp = avg / ITERS / warps_per_smsp;  // SMSP cycles consumed per warp-iteration
    printf("%-50s threads=%4d  cycles/warp-iter=%.2f  (FMA-pipe lower bound %.1f)  -> %.0f%% of FMA peak\n", name, threads,
           cyc_per_iter_warp, pipe_cycles_per_iter, 100.0 * pipe_cycles_per_iter / cyc_per_iter_warp);
    cudaFree(out); cudaFree(cyc); cudaFree(tin);
}

int main() {
    for (int th : {256, 512, 1024}) {
        if (th == 1024) break;  // launch bounds 512
        run<0>("filter pattern, FFMA2 scalar-broadcast q", th, 48);
        run<1>("filter pattern, FFMA2 explicit q pairs", th, 48);
        run<2>("filter pattern, scalar FFMA", th, 48);
        run<3>("8 chains acc=q_i*T+acc, FFMA2 broadcast", th, 48);
    }
    return 0;
}
