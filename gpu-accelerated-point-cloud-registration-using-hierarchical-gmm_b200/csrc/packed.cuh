// packed.cuh -- packed-FP32 (fma/add/mul.rn.f32x2 -> SASS FFMA2/FADD2/FMUL2) helpers shared by the flat sweep kernels
// (flat_em3.cu, flat_em5.cu): component pairs in float2 registers, the 12-operation quadratic form, and the warp
// reduce-scatter that folds PB per-point partial sums with PB+1 shuffles.
#pragma once
#include "common.cuh"

namespace hgmm {
constexpr int kChunk3 = 512;            // points staged in shared memory at a time (32 B each, duplicated)
constexpr int kPB3 = 8;                 // points per batch
constexpr float kUnder3 = 7.888609052210118e-31f;   // 2^-100

typedef unsigned long long u64;
__device__ __forceinline__ u64 f2u(float2 v) { return *reinterpret_cast<u64*>(&v); }
__device__ __forceinline__ float2 u2f(u64 v) { return *reinterpret_cast<float2*>(&v); }
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(f2u(a)), "l"(f2u(b)), "l"(f2u(c)));
    return u2f(d);
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
    u64 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2u(a)), "l"(f2u(b)));
    return u2f(d);
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
    u64 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2u(a)), "l"(f2u(b)));
    return u2f(d);
}

__device__ __forceinline__ void group_bar3(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// reduce v[0..PB) over the warp (PB = 8 or 4); lane L ends with point idx(L) in v[0]:
//   PB = 8: idx = 4*bit4 + 2*bit3 + bit2;  PB = 4: idx = 2*bit4 + bit3
template <int PB>
__device__ __forceinline__ void reduce_scatter(float* v, int lane, bool is_max);

template <>
__device__ __forceinline__ void reduce_scatter<4>(float* v, int lane, bool is_max) {
    const bool u4 = (lane & 16) != 0, u3 = (lane & 8) != 0;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float keep = u4 ? v[i + 2] : v[i], send = u4 ? v[i] : v[i + 2];
        const float o = __shfl_xor_sync(0xffffffffu, send, 16);
        v[i] = is_max ? fmaxf(keep, o) : keep + o;
    }
    {
        const float keep = u3 ? v[1] : v[0], send = u3 ? v[0] : v[1];
        const float o = __shfl_xor_sync(0xffffffffu, send, 8);
        v[0] = is_max ? fmaxf(keep, o) : keep + o;
    }
#pragma unroll
    for (int off = 4; off > 0; off >>= 1) {
        const float o = __shfl_xor_sync(0xffffffffu, v[0], off);
        v[0] = is_max ? fmaxf(v[0], o) : v[0] + o;
    }
}

template <>
__device__ __forceinline__ void reduce_scatter<8>(float* v, int lane, bool is_max) {
    const bool u4 = (lane & 16) != 0, u3 = (lane & 8) != 0, u2 = (lane & 4) != 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float keep = u4 ? v[i + 4] : v[i], send = u4 ? v[i] : v[i + 4];
        const float o = __shfl_xor_sync(0xffffffffu, send, 16);
        v[i] = is_max ? fmaxf(keep, o) : keep + o;
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float keep = u3 ? v[i + 2] : v[i], send = u3 ? v[i] : v[i + 2];
        const float o = __shfl_xor_sync(0xffffffffu, send, 8);
        v[i] = is_max ? fmaxf(keep, o) : keep + o;
    }
    {
        const float keep = u2 ? v[1] : v[0], send = u2 ? v[0] : v[1];
        const float o = __shfl_xor_sync(0xffffffffu, send, 4);
        v[0] = is_max ? fmaxf(keep, o) : keep + o;
    }
#pragma unroll
    for (int off = 2; off > 0; off >>= 1) {
        const float o = __shfl_xor_sync(0xffffffffu, v[0], off);
        v[0] = is_max ? fmaxf(v[0], o) : v[0] + o;
    }
}

struct PairParams {            // two components side by side; means stored negated so d = P + nm is one FADD2
    float2 nmx, nmy, nmz, c2;
    float2 axx, ayy, azz, axy, axz, ayz;
};

__device__ __forceinline__ float2 quad2(const PairParams& k, float2 X, float2 Y, float2 Z, float2& dx, float2& dy, float2& dz) {
    dx = fadd2(X, k.nmx);
    dy = fadd2(Y, k.nmy);
    dz = fadd2(Z, k.nmz);
    float2 t0 = fmul2(k.axz, dz);
    t0 = ffma2(k.axy, dy, t0);
    t0 = ffma2(k.axx, dx, t0);
    float2 t1 = fmul2(k.ayz, dz);
    t1 = ffma2(k.ayy, dy, t1);
    const float2 t2 = fmul2(k.azz, dz);
    float2 q = ffma2(dz, t2, k.c2);
    q = ffma2(dy, t1, q);
    q = ffma2(dx, t0, q);
    return q;
}

}  // namespace hgmm
