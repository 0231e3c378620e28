// fp32_pipe.cu -- issue-rate probes for the packed FP32 pipe of sm_100a (not part of the library).
// Each kernel runs 13 or 16 warps per SM, one CTA per SM, with 8 independent dependency chains per thread, and
// reports warp-instructions per clock per SM sub-partition for one operand pattern.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 f2u(float2 v) { return *reinterpret_cast<u64*>(&v); }
__device__ __forceinline__ float2 u2f(u64 v) { return *reinterpret_cast<float2*>(&v); }
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(f2u(a)), "l"(f2u(b)), "l"(f2u(c))); return u2f(d); }
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) { u64 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2u(a)), "l"(f2u(b))); return u2f(d); }
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) { u64 d; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2u(a)), "l"(f2u(b))); return u2f(d); }

// MODE 0: a = a*b + c, b and c loop-invariant (operand reuse possible)
// MODE 1: a_i = x_i * y_i + a_i with x_i, y_i distinct registers per chain (three distinct 64-bit sources)
// MODE 2: a_i = a_i + x_i (FADD2, two sources)
// MODE 3: a_i = a_i * x_i (FMUL2, two sources)
// MODE 4: a_i = x_i * y_i + a_i ; x_i = x_i + y_i   (FFMA2 + FADD2 alternating, all distinct)
// MODE 5: scalar FFMA a_i = x_i * y_i + a_i, three distinct registers (two chains per packed chain)
// MODE 6: FFMA2 a_i = k * y_i + a_i with ONE shared k (reuse in slot A), y_i distinct
template <int MODE>
__global__ void __launch_bounds__(512, 1) probe(float* out, int iters, float s) {
    float2 a[8], x[8], y[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        a[i] = make_float2(s + i, s - i);
        x[i] = make_float2(1.0f + 1e-7f * (i + threadIdx.x), 1.0f - 1e-7f * i);
        y[i] = make_float2(1e-7f * (i + 1), -1e-7f * (i + 2));
    }
    const float2 b = make_float2(1.0000001f, 0.9999999f), c = make_float2(1e-7f, -1e-7f);
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) a[i] = ffma2(a[i], b, c);
                if (MODE == 1) a[i] = ffma2(x[i], y[i], a[i]);
                if (MODE == 2) a[i] = fadd2(a[i], x[i]);
                if (MODE == 3) a[i] = fmul2(a[i], x[i]);
                if (MODE == 4) { a[i] = ffma2(x[i], y[i], a[i]); x[i] = fadd2(x[i], y[i]); }
                if (MODE == 5) { a[i].x = fmaf(x[i].x, y[i].x, a[i].x); a[i].y = fmaf(x[i].y, y[i].y, a[i].y); }
                if (MODE == 6) a[i] = ffma2(b, y[i], a[i]);
            }
        }
    }
    float2 r = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 8; ++i) r = fadd2(r, fadd2(a[i], x[i]));
    if (r.x + r.y == 12345.678f) out[0] = r.x;
}

template <int MODE>
static void run(const char* name, int threads, int ops_per_inner, float* d, int sms, double ghz) {
    const int iters = 4096;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        probe<MODE><<<sms, threads>>>(d, iters, 0.5f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    const double warp_insts = (double)(threads / 32) * iters * 64.0 * ops_per_inner;       // per SM
    const double cycles = best * 1e-3 * ghz * 1e9;
    printf("{\"probe\": \"%s\", \"threads\": %d, \"ms\": %.4f, \"warp_inst_per_clk_per_smsp\": %.4f}\n", name, threads, best,
           warp_insts / cycles / 4.0);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double ghz = khz * 1e-6;
    float* d; cudaMalloc(&d, 64);
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_ghz_nominal\": %.3f}\n", p.name, p.multiProcessorCount, ghz);
    for (int threads : {416, 512, 256}) {
        run<0>("ffma2 a=a*b+c (b,c invariant)", threads, 1, d, p.multiProcessorCount, ghz);
        run<1>("ffma2 a=x*y+a (3 distinct)", threads, 1, d, p.multiProcessorCount, ghz);
        run<6>("ffma2 a=k*y+a (k shared)", threads, 1, d, p.multiProcessorCount, ghz);
        run<2>("fadd2 a=a+x", threads, 1, d, p.multiProcessorCount, ghz);
        run<3>("fmul2 a=a*x", threads, 1, d, p.multiProcessorCount, ghz);
        run<4>("ffma2+fadd2 alternating", threads, 2, d, p.multiProcessorCount, ghz);
        run<5>("scalar ffma x2 (3 distinct)", threads, 2, d, p.multiProcessorCount, ghz);
    }
    return 0;
}
