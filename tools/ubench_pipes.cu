// Pipe-rate microbenchmark for sm_100a (B200): FFMA, FFMA2 (fma.rn.f32x2), FMNMX, FMNMX3,
// and the chamfer-filter instruction mix. Prints warp-instructions / clk / SM-sub-partition
// (from clock64 inside the kernel) and wall-clock TFLOP/s (from CUDA events).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_pipes ubench_pipes.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int ITERS = 4096;

template <int MODE>
__global__ void __launch_bounds__(256) pipe_kernel(float* out, long long* cycles, float seed) {
    float a[16];
    unsigned long long p[8];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = seed * (i + 1) + threadIdx.x;
#pragma unroll
    for (int i = 0; i < 8; ++i) { float2 t = make_float2(a[2*i], a[2*i+1]); p[i] = *reinterpret_cast<unsigned long long*>(&t); }
    float m0 = 1e30f, m1 = 1e30f, m2 = 1e30f, m3 = 1e30f;
    float b = seed * 0.5f, c = seed * 0.25f;
    unsigned long long bb, cc; { float2 t = make_float2(b, b); bb = *reinterpret_cast<unsigned long long*>(&t); t = make_float2(c, c); cc = *reinterpret_cast<unsigned long long*>(&t); }
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
        if (MODE == 0) {        // 16 independent FFMA
#pragma unroll
            for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
        } else if (MODE == 1) { // 8 independent FFMA2 (16 fmas)
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(bb), "l"(cc));
        } else if (MODE == 2) { // 16 FMNMX
#pragma unroll
            for (int i = 0; i < 16; ++i) asm volatile("min.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
        } else if (MODE == 3) { // 16 FMNMX3
#pragma unroll
            for (int i = 0; i < 16; ++i) asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
        } else if (MODE == 4) { // mix: 12 FFMA + 4 FMNMX  (scalar filter: 3 FFMA + 1 min per pair)
#pragma unroll
            for (int i = 0; i < 12; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
            asm volatile("min.f32 %0, %0, %1;" : "+f"(m0) : "f"(a[0]));
            asm volatile("min.f32 %0, %0, %1;" : "+f"(m1) : "f"(a[3]));
            asm volatile("min.f32 %0, %0, %1;" : "+f"(m2) : "f"(a[6]));
            asm volatile("min.f32 %0, %0, %1;" : "+f"(m3) : "f"(a[9]));
        } else if (MODE == 5) { // mix: 6 FFMA2 + 2 FMNMX3 (packed filter: 4 pairs)
#pragma unroll
            for (int i = 0; i < 6; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(bb), "l"(cc));
            asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(m0) : "f"(a[0]), "f"(a[1]));
            asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(m1) : "f"(a[2]), "f"(a[3]));
        } else if (MODE == 6) { // mix: 12 FFMA2 + 4 FMNMX3
#pragma unroll
            for (int i = 0; i < 6; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(bb), "l"(cc));
#pragma unroll
            for (int i = 0; i < 6; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(cc), "l"(bb));
            asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(m0) : "f"(a[0]), "f"(a[1]));
            asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(m1) : "f"(a[2]), "f"(a[3]));
            asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(m2) : "f"(a[4]), "f"(a[5]));
            asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(m3) : "f"(a[6]), "f"(a[7]));
        } else if (MODE == 7) { // 8 FFMA2 + 8 FMNMX (packed FMA with 2-input min)
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(bb), "l"(cc));
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("min.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
        } else if (MODE == 8) { // 8 FFMA2 + 8 FFMA: do they share the pipe?
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(bb), "l"(cc));
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
        } else if (MODE == 9) { // 8 FADD + 8 FMUL
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
#pragma unroll
            for (int i = 8; i < 16; ++i) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(c));
        } else if (MODE == 10) { // 8 FFMA + 8 setp/selp style compare-select
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
#pragma unroll
            for (int i = 8; i < 16; ++i) asm volatile("{.reg .pred q; setp.lt.f32 q, %0, %1; selp.f32 %0, %1, %0, q;}" : "+f"(a[i]) : "f"(a[i-8]));
        }
    }
    long long t1 = clock64();
    float s = m0 + m1 + m2 + m3;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) { float2 t = *reinterpret_cast<float2*>(&p[i]); s += t.x + t.y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int instr_per_iter, double flops_per_iter_thread, int ctas_per_sm, int nsm) {
    int grid = nsm * ctas_per_sm, block = 256;
    float* out; long long* cyc;
    CK(cudaMalloc(&out, sizeof(float) * grid * block));
    CK(cudaMalloc(&cyc, sizeof(long long) * grid));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int w = 0; w < 3; ++w) pipe_kernel<MODE><<<grid, block>>>(out, cyc, 1.0001f);
    CK(cudaDeviceSynchronize());
    float best_ms = 1e30f;
    for (int r = 0; r < 5; ++r) {
        CK(cudaEventRecord(e0));
        pipe_kernel<MODE><<<grid, block>>>(out, cyc, 1.0001f);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best_ms) best_ms = ms;
    }
    long long* h = (long long*)malloc(sizeof(long long) * grid);
    CK(cudaMemcpy(h, cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
    double avg = 0; for (int i = 0; i < grid; ++i) avg += (double)h[i]; avg /= grid;
    // warps per SMSP = ctas_per_sm * 8 / 4
    double warps_per_smsp = ctas_per_sm * (block / 32) / 4.0;
    double ipc_smsp = warps_per_smsp * (double)instr_per_iter * ITERS / avg;
    double tflops = flops_per_iter_thread * ITERS * (double)grid * block / (best_ms * 1e-3) / 1e12;
    double ghz = avg / (best_ms * 1e-3) / 1e9;
    printf("%-44s ctas/sm=%d  warp-instr/clk/SMSP=%.3f  cyc/iter/warp-set=%.2f  ms=%.4f  TFLOP/s=%.2f  (~%.2f GHz)\n",
           name, ctas_per_sm, ipc_smsp, avg / ITERS, best_ms, tflops, ghz);
    free(h); cudaFree(out); cudaFree(cyc);
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int nsm = prop.multiProcessorCount;
    printf("device %s, %d SMs, clock %d kHz\n", prop.name, nsm, prop.clockRate);
    for (int c : {1, 2, 4}) {
        run<0>("FFMA x16", 16, 32, c, nsm);
        run<1>("FFMA2 x8 (16 fma)", 8, 32, c, nsm);
        run<2>("FMNMX x16", 16, 0, c, nsm);
        run<3>("FMNMX3 x16", 16, 0, c, nsm);
        run<4>("12 FFMA + 4 FMNMX", 16, 24, c, nsm);
        run<5>("6 FFMA2 + 2 FMNMX3", 8, 24, c, nsm);
        run<6>("12 FFMA2 + 4 FMNMX3", 16, 48, c, nsm);
        run<7>("8 FFMA2 + 8 FMNMX", 16, 32, c, nsm);
        run<8>("8 FFMA2 + 8 FFMA", 16, 48, c, nsm);
        run<9>("8 FADD + 8 FMUL", 16, 16, c, nsm);
        run<10>("8 FFMA + 8 (FSETP+FSEL)", 24, 16, c, nsm);
    }
    return 0;
}
