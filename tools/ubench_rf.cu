// Register-operand pattern microbenchmark for FFMA2 / FFMA on sm_100a: how many SMSP cycles does one
// FFMA2 cost when its operands (a) come from the operand-reuse cache, (b) are all distinct registers?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_rf ubench_rf.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
constexpr int ITERS = 2048;
typedef unsigned long long u64;

__device__ __forceinline__ u64 pk(float a, float b) { float2 t = make_float2(a, b); return *reinterpret_cast<u64*>(&t); }
// acc = {s,s} * T + acc   (scalar-broadcast form)
#define F2S(acc, s, T) asm volatile("{.reg .b64 ss; mov.b64 ss, {%1, %1}; fma.rn.f32x2 %0, ss, %2, %0;}" : "+l"(acc) : "f"(s), "l"(T))
#define F2P(acc, S, T) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(S), "l"(T))
#define F1(acc, s, t) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc) : "f"(s), "f"(t))

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(float* out, long long* cycles, float seed) {
    u64 acc[16];
    float s[8];
    u64 S[8];
    u64 T[4];
    float a1[32], t1[8];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = pk(seed * i, seed * (i + 3) + threadIdx.x);
#pragma unroll
    for (int i = 0; i < 8; ++i) { s[i] = seed * (i + 1) * 1e-3f + threadIdx.x * 1e-6f; S[i] = pk(s[i], s[i] + 1e-4f); t1[i] = seed * (i + 2) * 1e-3f; }
#pragma unroll
    for (int i = 0; i < 4; ++i) T[i] = pk(seed * (i + 5) * 1e-3f, seed * (i + 9) * 1e-3f);
#pragma unroll
    for (int i = 0; i < 32; ++i) a1[i] = seed * i + threadIdx.x;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
        if (MODE == 0) {          // 32 FFMA2, T shared by 8 consecutive instructions (reuse-friendly), scalar s
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int i = 0; i < 8; ++i) F2S(acc[i], s[i], T[j]);
        } else if (MODE == 1) {   // 32 FFMA2, T changes every instruction (no reuse possible), scalar s
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) F2S(acc[i], s[i], T[(i + r) & 3]);
        } else if (MODE == 2) {   // 32 FFMA2, s shared by 4 consecutive instructions, T changes
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) F2S(acc[i * 4 + j], s[i + 4 * h], T[j]);
        } else if (MODE == 3) {   // MODE 0 with packed S pairs instead of scalar broadcast
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int i = 0; i < 8; ++i) F2P(acc[i], S[i], T[j]);
        } else if (MODE == 4) {   // MODE 1 with packed S pairs
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) F2P(acc[i], S[i], T[(i + r) & 3]);
        } else if (MODE == 5) {   // scalar FFMA, 32 per iteration, t shared by 8 consecutive
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int i = 0; i < 8; ++i) F1(a1[i], s[i], t1[j]);
        } else if (MODE == 6) {   // scalar FFMA, all three operands change every instruction
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) F1(a1[i + 8 * (r & 1)], s[(i + r) & 7], t1[(i + 3 * r + 1) & 7]);
        } else if (MODE == 7) {   // MODE 0 (reuse-friendly FFMA2) + 8 FMNMX3 consuming the accumulators of the previous round
#pragma unroll
            for (int j = 0; j < 4; ++j) {
#pragma unroll
                for (int i = 0; i < 8; ++i) F2S(acc[i], s[i], T[j]);
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    float2 v = *reinterpret_cast<float2*>(&acc[8 + i + 2 * j]);
                    asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(a1[i + 2 * j]) : "f"(v.x), "f"(v.y));
                }
            }
        } else if (MODE == 8) {   // MODE 0 + 16 FMNMX (2-input) per 32 FFMA2
#pragma unroll
            for (int j = 0; j < 4; ++j) {
#pragma unroll
                for (int i = 0; i < 8; ++i) F2S(acc[i], s[i], T[j]);
#pragma unroll
                for (int i = 0; i < 4; ++i) asm volatile("min.f32 %0, %0, %1;" : "+f"(a1[i + 4 * j]) : "f"(a1[16 + i + 4 * j]));
            }
        } else if (MODE == 9) {   // MODE 0 + 32 FMNMX per 32 FFMA2 (is the shadow slot really free?)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    F2S(acc[i], s[i], T[j]);
                    asm volatile("min.f32 %0, %0, %1;" : "+f"(a1[i + 8 * (j & 1)]) : "f"(a1[16 + i + 8 * (j & 1)]));
                }
            }
        } else if (MODE == 10) {  // MODE 0 + 16 FMNMX3 per 32 FFMA2
#pragma unroll
            for (int j = 0; j < 4; ++j) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    F2S(acc[i], s[i], T[j]);
                    if (i & 1) asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(a1[i + 8 * (j & 1)]) : "f"(a1[16 + i]), "f"(a1[24 + (i >> 1)]));
                }
            }
        }
    }
    long long t1c = clock64();
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) { float2 t = *reinterpret_cast<float2*>(&acc[i]); r += t.x + t.y; }
#pragma unroll
    for (int i = 0; i < 32; ++i) r += a1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1c - t0;
}

template <int MODE>
void run(const char* name, int threads, double fma_instr_per_iter) {
    int grid = 148;
    float* out; long long* cyc;
    CK(cudaMalloc(&out, sizeof(float) * grid * 512));
    CK(cudaMalloc(&cyc, sizeof(long long) * grid));
    for (int w = 0; w < 2; ++w) k<MODE><<<grid, threads>>>(out, cyc, 1.0001f);
    CK(cudaDeviceSynchronize());
    long long h[148];
    CK(cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
    double avg = 0; for (int i = 0; i < grid; ++i) avg += (double)h[i]; avg /= grid;
    double warps_per_smsp = threads / 32 / 4.0;
    double c = avg / ITERS / warps_per_smsp;  // SMSP cycles per warp-iteration
    printf("%-72s threads=%4d  cycles/warp-iter=%7.2f  cycles per FMA instr=%.3f\n", name, threads, c, c / fma_instr_per_iter);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int th : {128, 256, 512}) {
        run<0>("FFMA2 scalar-s, T reused x8", th, 32);
        run<1>("FFMA2 scalar-s, T changes every instr", th, 32);
        run<2>("FFMA2 scalar-s reused x4, T changes", th, 32);
        run<3>("FFMA2 packed-S, T reused x8", th, 32);
        run<4>("FFMA2 packed-S, T changes every instr", th, 32);
        run<5>("FFMA scalar, t reused x8", th, 32);
        run<6>("FFMA scalar, all operands change", th, 32);
        run<7>("FFMA2 (T reused) + 8 FMNMX3 on older results", th, 32);
        run<8>("FFMA2 (T reused) + 16 FMNMX", th, 32);
        run<9>("FFMA2 (T reused) + 32 FMNMX interleaved 1:1", th, 32);
        run<10>("FFMA2 (T reused) + 16 FMNMX3 interleaved 2:1", th, 32);
    }
    return 0;
}
