// Microbenchmark: FFMA vs FFMA2 (packed f32x2) vs mixes with MUFU.RSQ / FMNMX on sm_100a.
#include <cuda_runtime.h>
#include <cstdio>
#define ITERS 4096
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float seed) {
    float a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    float2 p0 = make_float2(a0, a1), p1 = make_float2(a2, a3), p2 = make_float2(a4, a5), p3 = make_float2(a6, a7);
    float2 p4 = p0, p5 = p1, p6 = p2, p7 = p3;
    const float m = 0.999f, c = 0.001f;
    const float2 m2 = make_float2(m, m), c2 = make_float2(c, c);
#pragma unroll 1
    for (int i = 0; i < ITERS; ++i) {
        if (MODE == 0) {          // 8 independent FFMA
            a0 = fmaf(a0, m, c); a1 = fmaf(a1, m, c); a2 = fmaf(a2, m, c); a3 = fmaf(a3, m, c);
            a4 = fmaf(a4, m, c); a5 = fmaf(a5, m, c); a6 = fmaf(a6, m, c); a7 = fmaf(a7, m, c);
        } else if (MODE == 1) {   // 8 independent FFMA2 (16 FMAs)
            p0 = __ffma2_rn(p0, m2, c2); p1 = __ffma2_rn(p1, m2, c2); p2 = __ffma2_rn(p2, m2, c2); p3 = __ffma2_rn(p3, m2, c2);
            p4 = __ffma2_rn(p4, m2, c2); p5 = __ffma2_rn(p5, m2, c2); p6 = __ffma2_rn(p6, m2, c2); p7 = __ffma2_rn(p7, m2, c2);
        } else if (MODE == 2) {   // 8 FFMA2 + 2 MUFU.RSQ + 2 FMNMX
            p0 = __ffma2_rn(p0, m2, c2); p1 = __ffma2_rn(p1, m2, c2); p2 = __ffma2_rn(p2, m2, c2); p3 = __ffma2_rn(p3, m2, c2);
            p4 = __ffma2_rn(p4, m2, c2); p5 = __ffma2_rn(p5, m2, c2); p6 = __ffma2_rn(p6, m2, c2); p7 = __ffma2_rn(p7, m2, c2);
            float y0, y1;
            asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y0) : "f"(p0.x));
            asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y1) : "f"(p1.x));
            a0 = fminf(a0, y0); a1 = fminf(a1, y1);
        } else if (MODE == 3) {   // 16 FFMA + 2 MUFU + 2 FMNMX (same math as mode 2, unpacked)
            p0.x = fmaf(p0.x, m, c); p0.y = fmaf(p0.y, m, c); p1.x = fmaf(p1.x, m, c); p1.y = fmaf(p1.y, m, c);
            p2.x = fmaf(p2.x, m, c); p2.y = fmaf(p2.y, m, c); p3.x = fmaf(p3.x, m, c); p3.y = fmaf(p3.y, m, c);
            p4.x = fmaf(p4.x, m, c); p4.y = fmaf(p4.y, m, c); p5.x = fmaf(p5.x, m, c); p5.y = fmaf(p5.y, m, c);
            p6.x = fmaf(p6.x, m, c); p6.y = fmaf(p6.y, m, c); p7.x = fmaf(p7.x, m, c); p7.y = fmaf(p7.y, m, c);
            float y0, y1;
            asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y0) : "f"(p0.x));
            asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y1) : "f"(p1.x));
            a0 = fminf(a0, y0); a1 = fminf(a1, y1);
        } else if (MODE == 4) {   // 8 MUFU only
            asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(a0) : "f"(a0)); asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(a1) : "f"(a1));
            asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(a2) : "f"(a2)); asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(a3) : "f"(a3));
            asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(a4) : "f"(a4)); asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(a5) : "f"(a5));
            asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(a6) : "f"(a6)); asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(a7) : "f"(a7));
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + p0.x + p0.y + p1.x + p1.y + p2.x + p2.y + p3.x + p3.y + p4.x + p4.y + p5.x + p5.y + p6.x + p6.y + p7.x + p7.y;
}
template <int MODE> void run(const char* name, double fma_per_iter, double inst_per_iter) {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 8, 256>>>(out, 1.0f); cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) k<MODE><<<148 * 8, 256>>>(out, 1.0f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    double threads = 148.0 * 8 * 256;
    double fmas = threads * ITERS * fma_per_iter, insts = threads / 32 * ITERS * inst_per_iter;
    printf("%-40s %.3f ms  %.1f TFLOP/s (2 flop/FMA)  %.2f warp-inst/clk/SM @1.965GHz\n", name, ms, 2 * fmas / ms / 1e9,
           insts / (ms * 1e-3) / 148 / 1.965e9);
    cudaFree(out);
}
int main() {
    run<0>("8 FFMA", 8, 8);
    run<1>("8 FFMA2", 16, 8);
    run<2>("8 FFMA2 + 2 MUFU + 2 FMNMX", 16, 12);
    run<3>("16 FFMA + 2 MUFU + 2 FMNMX", 16, 20);
    run<4>("8 MUFU.RSQ", 0, 8);
    return 0;
}
