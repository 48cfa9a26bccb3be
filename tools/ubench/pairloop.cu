// Microbenchmark of K2's inner loop on sm_100a: pair-interleaved environment tile in shared memory, 2 trial atoms per thread,
// packed FP32x2.  MODE 0: the shipped form, d = e - a (12 FMA-pipe lane-operations per pair + FMNMX3);
// MODE 1: expansion d2 = |e|^2 + |a|^2 - 2 e.a with |e|^2 staged per environment atom (10 lane-operations, no running minimum).
// 128 threads, 4 CTAs per SM (register cap 128), like score_generic_kernel.
#include <cuda_runtime.h>
#include <cstdio>
#define TILE 1024
#define REPS 64
template <int MODE>
__global__ void __launch_bounds__(128, 4) k(const float4* __restrict__ env, float* out) {
    __shared__ float4 s_env[TILE];          // pair p: [2p] = x0 x1 y0 y1, [2p+1] = z0 z1 q0 q1
    __shared__ float2 s_e2[TILE / 2];
    for (int i = threadIdx.x; i < TILE; i += 128) s_env[i] = env[i];
    for (int i = threadIdx.x; i < TILE / 2; i += 128) {
        const float4 a = env[2 * i], b = env[2 * i + 1];
        s_e2[i] = make_float2(a.x * a.x + a.z * a.z + b.x * b.x, a.y * a.y + a.w * a.w + b.y * b.y);
    }
    __syncthreads();
    float2 nx[2], ny[2], nz[2], a2[2];
    for (int r = 0; r < 2; ++r) {
        const float t = 0.01f * (threadIdx.x + 128 * r);
        nx[r] = make_float2(-t, -t); ny[r] = make_float2(-0.5f * t, -0.5f * t); nz[r] = make_float2(-0.25f * t, -0.25f * t);
        const float s = t * t * (1.f + 0.25f + 0.0625f);
        a2[r] = make_float2(s, s);
        if (MODE == 1) { nx[r].x *= 2.f; nx[r].y *= 2.f; ny[r].x *= 2.f; ny[r].y *= 2.f; nz[r].x *= 2.f; nz[r].y *= 2.f; }
    }
    double E6[2] = {0, 0}, E3[2] = {0, 0}, EC[2] = {0, 0};
    float dmin[2] = {1e30f, 1e30f};
    for (int rep = 0; rep < REPS; ++rep) {
        for (int pb = 0; pb < TILE / 2; pb += 16) {
            float2 a6[2], a3[2], cq[2];
            for (int r = 0; r < 2; ++r) { a6[r] = make_float2(0.f, 0.f); a3[r] = a6[r]; cq[r] = a6[r]; }
#pragma unroll 4
            for (int p = pb; p < pb + 16; ++p) {
                const float4 exy = s_env[2 * p], ezq = s_env[2 * p + 1];
                const float2 ex = make_float2(exy.x, exy.y), ey = make_float2(exy.z, exy.w);
                const float2 ez = make_float2(ezq.x, ezq.y), eq = make_float2(ezq.z, ezq.w);
                float2 e2;
                if (MODE == 1) e2 = s_e2[p];
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    float2 d2;
                    if (MODE == 0) {
                        const float2 dx = __fadd2_rn(ex, nx[r]), dy = __fadd2_rn(ey, ny[r]), dz = __fadd2_rn(ez, nz[r]);
                        d2 = __ffma2_rn(dz, dz, __ffma2_rn(dy, dy, __fmul2_rn(dx, dx)));
                        dmin[r] = fminf(dmin[r], fminf(d2.x, d2.y));
                    } else {
                        d2 = __ffma2_rn(ex, nx[r], __ffma2_rn(ey, ny[r], __ffma2_rn(ez, nz[r], __fadd2_rn(e2, a2[r]))));
                    }
                    float2 y;
                    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y.x) : "f"(d2.x));
                    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y.y) : "f"(d2.y));
                    const float2 w = __fmul2_rn(y, y);
                    const float2 i3 = __fmul2_rn(__fmul2_rn(w, w), w);
                    a6[r] = __ffma2_rn(i3, i3, a6[r]);
                    a3[r] = __fadd2_rn(a3[r], i3);
                    cq[r] = __ffma2_rn(eq, y, cq[r]);
                }
            }
            for (int r = 0; r < 2; ++r) { E6[r] += (double)(a6[r].x + a6[r].y); E3[r] += (double)(a3[r].x + a3[r].y); EC[r] += (double)(cq[r].x + cq[r].y); }
        }
    }
    out[blockIdx.x * 128 + threadIdx.x] = (float)(E6[0] + E6[1] + E3[0] + E3[1] + EC[0] + EC[1]) + dmin[0] + dmin[1];
}
template <int MODE> void run(const char* name) {
    float4* env; float* out;
    cudaMalloc(&env, TILE * 16); cudaMalloc(&out, 148 * 16 * 128 * 4);
    float4 h[TILE];
    for (int i = 0; i < TILE; ++i) h[i] = make_float4(5.f + 0.01f * i, 6.f + 0.02f * i, 7.f + 0.015f * i, 8.f + 0.01f * i);
    cudaMemcpy(env, h, sizeof(h), cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 16, 128>>>(env, out); cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) k<MODE><<<148 * 16, 128>>>(env, out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    const double pairs = 148.0 * 16 * 256 * (double)TILE * REPS;
    printf("%-52s %.3f ms  %.3e pairs/s  = %.1f TFLOP/s at 22 flop/pair\n", name, ms, pairs / (ms * 1e-3), 22 * pairs / (ms * 1e-3) / 1e12);
    cudaFree(env); cudaFree(out);
}
int main() {
    run<0>("d = e - a (12 lane-ops + FMNMX3)");
    run<1>("d2 = e2 + a2 - 2 e.a (10 lane-ops)");
    return 0;
}
