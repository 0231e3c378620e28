// flat_em6.cu -- fused E+M sweep of the flat mixture, one component per thread, densities staged in shared memory.
//
// Same contract as the other sweep kernels (expectationStep + maximizationStep of
// src/c++/gmm_fit/gmm_kernels.cu:278-350; e_step + m_step of src/python/gmm_waymo/src/gmm_impl.py:90-116).
// What is different, and why (profiles/microbench/fp32_pipe.cu, measured on this pool's B200):
//   * a packed FFMA2 with three distinct 64-bit sources issues every ~3.2 clk per sub-partition (register-file
//     bandwidth), a scalar FFMA with three distinct sources every ~1.1 clk -- so for this operand mix (17 of the 29
//     operations per pair have three live sources) two scalar instructions are cheaper than one packed one;
//   * one component per thread means S = Jp/32 warps per CTA (25 for J = 800) instead of ceil(S/2) = 13: the
//     busiest sub-partition carries 7 of 25 warps (89 % balance) instead of 4 of 13 (78 %), each thread needs half the
//     registers, and twice as many warps hide the MUFU / shared-memory latencies.
// Structure: the CTA owns a contiguous range of points and all J components.  Per chunk of up to 64 points
//   pass 1  q2, e = 2^(q2 - Cref) for 8 points at a time; e parked in a private shared-memory column
//           (STS.128 per 4 points); per-point partial sums: register reduce-scatter -> red[point][warp]
//   barrier; finish: warp w folds points w, w+W, ... (lane = source warp, xor-shuffle tree) -> inv[point], log-lik
//   barrier; pass 2  gamma = e * inv, ten centred moments in registers (LDS.128 per 4 points for x, y, z, e, inv)
// A chunk containing a point whose sum underflows the fixed reference is redone with exact per-point maxima.
// Every reduction has a fixed order: fits are bit-reproducible run to run.
#include "common.cuh"
#include "kernels.h"
#include "packed.cuh"

namespace hgmm {

constexpr int kChunk6 = 512;            // points staged per round: 128 groups of 4 in SoA form, 6 KB

// dynamic shared memory: float4 spts[kChunk6/4][3] | float red[CH][32] | float inv[CH] | float mval[CH] | float uflag[CH]
//                        | double wsum[32][2] | int flags[4] | float4 ebuf[CH/4][T]
__host__ __device__ inline size_t flat6_smem_bytes(int CH, int T) {
    return (size_t)(kChunk6 / 4) * 48 + (size_t)CH * 128 + (size_t)CH * 12 + 32 * 16 + 16 + (size_t)CH * T * 4;
}

struct CompParams {            // one component; mean negated so d = p + nm
    float nmx, nmy, nmz, c2, axx, ayy, azz, axy, axz, ayz;
};

__device__ __forceinline__ float quad1(const CompParams& k, float x, float y, float z) {
    const float dx = x + k.nmx, dy = y + k.nmy, dz = z + k.nmz;
    float t0 = k.axz * dz;
    t0 = fmaf(k.axy, dy, t0);
    t0 = fmaf(k.axx, dx, t0);
    float t1 = k.ayz * dz;
    t1 = fmaf(k.ayy, dy, t1);
    const float t2 = k.azz * dz;
    float q = fmaf(dz, t2, k.c2);
    q = fmaf(dy, t1, q);
    q = fmaf(dx, t0, q);
    return q;
}

__device__ __forceinline__ float f4get(const float4& v, int i) { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }

template <int MAXT>
__global__ void __launch_bounds__(MAXT, 1) em_flat6_kernel(const float* __restrict__ px, const float* __restrict__ py,
                                                           const float* __restrict__ pz, int n,
                                                           const PackedComp* __restrict__ packed,
                                                           const float* __restrict__ cref_blocks, int n_cref, int J, int Jp,
                                                           int W, int CH, float* __restrict__ partial,
                                                           double* __restrict__ rowaux, const int* __restrict__ done_flag,
                                                           float norm_eps_on) {
    if (*done_flag) return;
    constexpr int PB = 8;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int T = blockDim.x;
    float4* spts = reinterpret_cast<float4*>(smem_raw);                    // [kChunk6/4][3]  x4 | y4 | z4
    float* red = reinterpret_cast<float*>(spts + (kChunk6 / 4) * 3);       // [CH][32]
    float* inv = red + CH * 32;                                            // [CH]
    float* mval = inv + CH;                                                // [CH]  rare path: exact maxima
    float* uflag = mval + CH;                                              // [CH]  1 = the point's fixed-reference sum underflowed
    double* wsum = reinterpret_cast<double*>(uflag + CH);                  // [32][2]
    int* flags = reinterpret_cast<int*>(wsum + 64);                        // [4]
    float4* ebuf = reinterpret_cast<float4*>(flags + 4);                   // [CH/4][T]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ridx = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
    const bool rwriter = (lane & 3) == 0;
    if (tid < 4) flags[tid] = 0;

    float cref = -INFINITY;
    for (int i = 0; i < n_cref; ++i) cref = fmaxf(cref, __ldg(cref_blocks + i));
    if (!(cref > kNegBig)) cref = 0.f;

    CompParams k;
    {
        const float4* a4 = reinterpret_cast<const float4*>(packed + warp * 32 + lane);
        const float4 a0 = __ldg(a4), a1 = __ldg(a4 + 1), a2 = __ldg(a4 + 2);
        k.nmx = -a0.x; k.nmy = -a0.y; k.nmz = -a0.z;
        k.c2 = a0.w - cref;
        k.axx = a1.x; k.ayy = a1.y; k.azz = a1.z; k.axy = a1.w; k.axz = a2.x; k.ayz = a2.y;
    }
    float a[kMom];
#pragma unroll
    for (int m = 0; m < kMom; ++m) a[m] = 0.f;
    double ll = 0.0, nlive = 0.0;                         // lane 0 of every warp, for the points it finishes

    const int per = (int)(((long long)n + gridDim.x - 1) / gridDim.x);
    const int lo = min(n, (int)blockIdx.x * per), hi = min(n, lo + per);
    int parity = 0;
    float4* ecol = ebuf + tid;                            // group g of the column is ecol[g * T]

    for (int cb = lo; cb < hi; cb += kChunk6) {
        const int cn = min(kChunk6, hi - cb);
        const int cn8 = (cn + 7) & ~7;                    // staged (padded with copies of the last point) to whole batches
        __syncthreads();
        for (int i = tid; i < cn8; i += T) {
            const int src = cb + min(i, cn - 1);
            float* g = reinterpret_cast<float*>(spts + (i >> 2) * 3);
            g[i & 3] = px[src];
            g[4 + (i & 3)] = py[src];
            g[8 + (i & 3)] = pz[src];
        }
        __syncthreads();
        for (int c0 = 0; c0 < cn; c0 += CH) {
            const int ch = min(CH, cn - c0);              // valid points of this chunk
            const int ch8 = (ch + 7) & ~7;
            const float4* cpts = spts + (c0 >> 2) * 3;
            // ---------------- pass 1
            for (int b = 0; b < ch8; b += PB) {
                float q[PB];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const float4 X = cpts[((b >> 2) + h) * 3], Y = cpts[((b >> 2) + h) * 3 + 1], Z = cpts[((b >> 2) + h) * 3 + 2];
#pragma unroll
                    for (int i = 0; i < 4; ++i) q[4 * h + i] = quad1(k, f4get(X, i), f4get(Y, i), f4get(Z, i));
                }
#pragma unroll
                for (int p = 0; p < PB; ++p) q[p] = ex2f(q[p]);
                ecol[(size_t)(b >> 2) * T] = make_float4(q[0], q[1], q[2], q[3]);
                ecol[(size_t)((b >> 2) + 1) * T] = make_float4(q[4], q[5], q[6], q[7]);
                reduce_scatter<PB>(q, lane, false);
                if (rwriter) red[(b + ridx) * 32 + warp] = q[0];
            }
            __syncthreads();
            // ---------------- finish: warp w folds the partial sums of points w, w + W, ...
            for (int pt = warp; pt < ch8; pt += W) {
                float v = lane < W ? red[pt * 32 + lane] : 0.f;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0) {
                    const bool valid = pt < ch;
                    const bool under = valid && !(v >= kUnder3);
                    float iv = (valid && !under) ? __fdividef(1.0f, v) : 0.f;
                    if (valid && !under) {
                        const float lse2 = cref + lg2f(v);
                        float norm2 = lse2;
                        if (norm_eps_on != 0.f) {            // gmm_impl.py:113  log(sum exp + 1e-8)
                            const float Mx = fmaxf(lse2, kLog2Eps8);
                            norm2 = Mx + lg2f(ex2f(lse2 - Mx) + ex2f(kLog2Eps8 - Mx));
                            iv *= ex2f(lse2 - norm2);
                        }
                        ll += (double)(norm2 * kLn2);
                        nlive += 1.0;
                    }
                    inv[pt] = iv;
                    uflag[pt] = under ? 1.f : 0.f;
                    if (under) flags[parity] = 1;
                }
            }
            __syncthreads();
            const bool any_under = flags[parity] != 0;
            if (tid == 0) flags[parity ^ 1] = 0;          // next chunk's flag; its last readers passed the barrier above
            if (any_under) {
                // ---------------- rare path: exact per-point maxima for the whole chunk (CTA-uniform branch)
                for (int b = 0; b < ch8; b += PB) {
                    float q[PB];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const float4 X = cpts[((b >> 2) + h) * 3], Y = cpts[((b >> 2) + h) * 3 + 1], Z = cpts[((b >> 2) + h) * 3 + 2];
#pragma unroll
                        for (int i = 0; i < 4; ++i) q[4 * h + i] = fmaxf(quad1(k, f4get(X, i), f4get(Y, i), f4get(Z, i)), kNegBig);
                    }
                    reduce_scatter<PB>(q, lane, true);
                    if (rwriter) red[(b + ridx) * 32 + warp] = q[0];
                }
                __syncthreads();
                for (int pt = warp; pt < ch8; pt += W) {
                    float v = lane < W ? red[pt * 32 + lane] : kNegBig;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
                    if (lane == 0) mval[pt] = v;
                }
                __syncthreads();
                for (int b = 0; b < ch8; b += PB) {
                    float q[PB];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const float4 X = cpts[((b >> 2) + h) * 3], Y = cpts[((b >> 2) + h) * 3 + 1], Z = cpts[((b >> 2) + h) * 3 + 2];
                        const float4 M = *reinterpret_cast<const float4*>(mval + b + 4 * h);
#pragma unroll
                        for (int i = 0; i < 4; ++i) q[4 * h + i] = ex2f(quad1(k, f4get(X, i), f4get(Y, i), f4get(Z, i)) - f4get(M, i));
                    }
                    ecol[(size_t)(b >> 2) * T] = make_float4(q[0], q[1], q[2], q[3]);
                    ecol[(size_t)((b >> 2) + 1) * T] = make_float4(q[4], q[5], q[6], q[7]);
                    reduce_scatter<PB>(q, lane, false);
                    if (rwriter) red[(b + ridx) * 32 + warp] = q[0];
                }
                __syncthreads();
                for (int pt = warp; pt < ch8; pt += W) {
                    float v = lane < W ? red[pt * 32 + lane] : 0.f;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                    if (lane == 0) {
                        const bool valid = pt < ch;
                        const bool under = uflag[pt] != 0.f;
                        const float m = mval[pt];
                        float iv = 0.f;
                        if (valid && v > 0.f && m > kNegBig) {
                            const float lse2 = cref + m + lg2f(v);
                            float norm2 = lse2, scale = 1.0f;
                            if (norm_eps_on != 0.f) {
                                const float Mx = fmaxf(lse2, kLog2Eps8);
                                norm2 = Mx + lg2f(ex2f(lse2 - Mx) + ex2f(kLog2Eps8 - Mx));
                                scale = ex2f(lse2 - norm2);
                            }
                            iv = scale / v;
                            if (under) {                     // the fast path left only the underflowed points out
                                ll += (double)(norm2 * kLn2);
                                nlive += 1.0;
                            }
                        } else if (valid && under && norm_eps_on != 0.f) {
                            ll += (double)(kLog2Eps8 * kLn2);  // log(0 + 1e-8)
                        }
                        inv[pt] = iv;
                    }
                }
                __syncthreads();
            }
            // ---------------- pass 2: moments, four points per shared-memory round
            const int ngroups = (ch + 3) >> 2;
#pragma unroll 2
            for (int g = 0; g < ngroups; ++g) {
                const float4 X = cpts[g * 3], Y = cpts[g * 3 + 1], Z = cpts[g * 3 + 2];
                const float4 E = ecol[(size_t)g * T];
                const float4 I = *reinterpret_cast<const float4*>(inv + 4 * g);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float gam = f4get(E, i) * f4get(I, i);
                    const float dx = f4get(X, i) + k.nmx, dy = f4get(Y, i) + k.nmy, dz = f4get(Z, i) + k.nmz;
                    const float gx = gam * dx, gy = gam * dy, gz = gam * dz;
                    a[0] += gam;
                    a[1] += gx;
                    a[2] += gy;
                    a[3] += gz;
                    a[4] = fmaf(gx, dx, a[4]);
                    a[5] = fmaf(gx, dy, a[5]);
                    a[6] = fmaf(gx, dz, a[6]);
                    a[7] = fmaf(gy, dy, a[7]);
                    a[8] = fmaf(gy, dz, a[8]);
                    a[9] = fmaf(gz, dz, a[9]);
                }
            }
            parity ^= 1;
        }
    }
    // ---- partial rows: partial[row][m][Jp], row = blockIdx
    float* dst = partial + (size_t)blockIdx.x * kMom * Jp;
    {
        const int j = warp * 32 + lane;
#pragma unroll
        for (int m = 0; m < kMom; ++m) dst[(size_t)m * Jp + j] = a[m];
    }
    __syncthreads();
    if (lane == 0) {
        wsum[2 * warp] = ll;
        wsum[2 * warp + 1] = nlive;
    }
    __syncthreads();
    if (tid == 0) {
        double s0 = 0.0, s1 = 0.0;
        for (int w = 0; w < W; ++w) {
            s0 += wsum[2 * w];
            s1 += wsum[2 * w + 1];
        }
        rowaux[2 * blockIdx.x] = s0;
        rowaux[2 * blockIdx.x + 1] = s1;
    }
}

static int flat6_chunk(int T, int smem_optin) {
    int ch = 64;
    while (ch > 8 && flat6_smem_bytes(ch, T) > (size_t)smem_optin) ch -= 8;
    return ch;
}

cudaError_t launch_em_flat6(const float* x, const float* y, const float* z, int n, const FlatModel& m, const float* cref_blocks,
                            int grid, float* partial, double* rowaux, const int* done_flag, cudaStream_t s) {
    const int smem_optin = device_smem_optin();
    static DeviceOnce once;
    if (once.first()) {
        cudaFuncSetAttribute(em_flat6_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin);
        cudaFuncSetAttribute(em_flat6_kernel<640>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin);
        cudaFuncSetAttribute(em_flat6_kernel<768>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin);
        cudaFuncSetAttribute(em_flat6_kernel<896>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin);
        cudaFuncSetAttribute(em_flat6_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin);
    }
    const int W = m.Jp / 32;
    if (W < 1 || W > 32) return cudaErrorInvalidValue;
    const float eps_on = m.flavor != HGMM_FLAVOR_CPP ? 1.f : 0.f;
    const int T = W * 32;
    const int CH = flat6_chunk(T, smem_optin);
    const size_t smem = flat6_smem_bytes(CH, T);
#define HGMM_L6(MT) em_flat6_kernel<MT><<<grid, T, smem, s>>>(x, y, z, n, m.packed, cref_blocks, W, m.J, m.Jp, W, CH, partial, rowaux, done_flag, eps_on)
    if (T <= 512) HGMM_L6(512);
    else if (T <= 640) HGMM_L6(640);
    else if (T <= 768) HGMM_L6(768);
    else if (T <= 896) HGMM_L6(896);
    else HGMM_L6(1024);
#undef HGMM_L6
    return cudaGetLastError();
}

}  // namespace hgmm
