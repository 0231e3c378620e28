// flat_em5.cu -- packed-FP32 fused E+M sweep with the unnormalised densities staged in shared memory.
//
// Same contract as em_flat3_kernel (expectationStep + maximizationStep of src/c++/gmm_fit/gmm_kernels.cu:278-350;
// e_step + m_step of src/python/gmm_waymo/src/gmm_impl.py:90-116) and the same arithmetic per (point, component)
// pair -- 12 packed operations for q2 and e = 2^(q2 - Cref), 17 for the ten centred moments -- but the
// normalisation over J, which needs every warp of the CTA, is taken once per CHUNK of up to 56 points instead of
// once per batch of 8: each lane parks its pair of e values for the whole chunk in a private shared-memory column
// (conflict-free STS.64 / LDS.64, 8 B x threads x chunk points = up to 186 KB of the SM's 227 KB), so the CTA
// meets at one barrier per chunk and both phases around it are long, barrier-free FFMA2 streams.
//
//   pass 1 (per batch of 8 points): q2, e -> ebuf[point][thread]; per-point partial sums -> register
//                                   reduce-scatter -> red[chunk parity][point][warp]
//   __syncthreads
//   finish (every warp, redundantly, lanes = points): 1/sum, log-likelihood (warp 0), -> fin[warp][point]
//   pass 2 (per point): gamma = e * 1/sum, ten centred moments in float2 registers
//
// A chunk containing a point whose sum underflows the fixed reference (2^-100: further than ~11 sigma from every
// component) is redone with an exact per-point maximum (two more barriers, CTA-uniform decision).
// Partial moment rows + the fixed-order fp64 reduction are shared with flat_em2.cu / flat_em3.cu.
#include "common.cuh"
#include "kernels.h"
#include "packed.cuh"

namespace hgmm {

constexpr int kMaxWarps5 = 16;

// dynamic shared memory layout (CH = chunk points, multiple of 8; T = threads):
//   float4 spts[kChunk3][2] | float red[2][CH][16] | float redslow[2][CH][16] | float2 fin[16][CH] | float2 ebuf[CH][T]
__host__ __device__ inline size_t flat5_smem_bytes(int CH, int T) {
    return (size_t)kChunk3 * 32 + (size_t)CH * (2 * 64 + 2 * 64 + 128) + (size_t)CH * T * 8;
}

template <int MAXT>
__global__ void __launch_bounds__(MAXT, 1) em_flat5_kernel(const float* __restrict__ px, const float* __restrict__ py,
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
    float4* spts = reinterpret_cast<float4*>(smem_raw);                    // [kChunk3][2]  (x,x,y,y) (z,z,0,0)
    float* red = reinterpret_cast<float*>(spts + kChunk3 * 2);             // [2][CH][16]
    float* redslow = red + 2 * CH * 16;                                    // [2][CH][16]   rare path: maxima | sums
    float2* fin = reinterpret_cast<float2*>(redslow + 2 * CH * 16);        // [16][CH]      (inv, inv), private to each warp
    float2* ebuf = fin + 16 * CH;                                          // [CH][T]       private column per thread

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 4 * CH * 16; i += T) red[i] = 0.f;              // warp slots >= W stay zero (red and redslow)
    const int S = Jp >> 5;
    const int ridx = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
    const bool rwriter = (lane & 3) == 0;

    float cref = -INFINITY;
    for (int i = 0; i < n_cref; ++i) cref = fmaxf(cref, __ldg(cref_blocks + i));
    if (!(cref > kNegBig)) cref = 0.f;

    // ---- the lane's component pair -> registers: slots warp (low half) and warp + W (high half)
    PairParams k;
    const bool live0 = warp < S, live1 = warp + W < S;
    {
        const float4* a4 = reinterpret_cast<const float4*>(packed + (live0 ? warp * 32 + lane : 0));
        const float4* b4 = reinterpret_cast<const float4*>(packed + (live1 ? (warp + W) * 32 + lane : 0));
        const float4 a0 = __ldg(a4), a1 = __ldg(a4 + 1), a2 = __ldg(a4 + 2);
        const float4 b0 = __ldg(b4), b1 = __ldg(b4 + 1), b2 = __ldg(b4 + 2);
        k.nmx = make_float2(-a0.x, -b0.x);
        k.nmy = make_float2(-a0.y, -b0.y);
        k.nmz = make_float2(-a0.z, -b0.z);
        k.c2 = make_float2(live0 ? a0.w - cref : -INFINITY, live1 ? b0.w - cref : -INFINITY);
        k.axx = make_float2(a1.x, b1.x);
        k.ayy = make_float2(a1.y, b1.y);
        k.azz = make_float2(a1.z, b1.z);
        k.axy = make_float2(a1.w, b1.w);
        k.axz = make_float2(a2.x, b2.x);
        k.ayz = make_float2(a2.y, b2.y);
    }
    float2 a[kMom];
#pragma unroll
    for (int m = 0; m < kMom; ++m) a[m] = make_float2(0.f, 0.f);
    double ll = 0.0, nlive = 0.0;                         // accumulated by the finishing lanes of warp 0

    const int per = (int)(((long long)n + gridDim.x - 1) / gridDim.x);
    const int lo = min(n, (int)blockIdx.x * per), hi = min(n, lo + per);
    int parity = 0;
    float2* ecol = ebuf + tid;                            // element p of the column is ecol[p * T]
    float2* myfin = fin + warp * CH;

    for (int cb = lo; cb < hi; cb += kChunk3) {
        const int cn = min(kChunk3, hi - cb);
        __syncthreads();
        for (int i = tid; i < cn; i += T) {
            const float x = px[cb + i], y = py[cb + i], z = pz[cb + i];
            spts[2 * i] = make_float4(x, x, y, y);
            spts[2 * i + 1] = make_float4(z, z, 0.f, 0.f);
        }
        __syncthreads();
        for (int c0 = 0; c0 < cn; c0 += CH) {
            const int ch = min(CH, cn - c0);              // points of this chunk
            float* redp = red + parity * CH * 16;
            // ---------------- pass 1: e -> ebuf, per-point partial sums -> red[parity]
            for (int b = 0; b < ch; b += PB) {
                float sm[PB];
                float2 q[PB];
                // all shared-memory loads of the batch first: the compiler must keep an LDS behind an earlier STS it cannot
                // prove disjoint, which would serialise the eight q2 chains
#pragma unroll
                for (int p = 0; p < PB; ++p) {
                    const int ip = min(c0 + b + p, cn - 1);
                    const float4 P0 = spts[2 * ip], P1 = spts[2 * ip + 1];
                    float2 dx, dy, dz;
                    q[p] = quad2(k, make_float2(P0.x, P0.y), make_float2(P0.z, P0.w), make_float2(P1.x, P1.y), dx, dy, dz);
                }
#pragma unroll
                for (int p = 0; p < PB; ++p) {
                    const float2 e = make_float2(ex2f(q[p].x), ex2f(q[p].y));
                    ecol[(size_t)(b + p) * T] = e;
                    sm[p] = e.x + e.y;
                }
                reduce_scatter<PB>(sm, lane, false);
                if (rwriter) redp[(b + ridx) * 16 + warp] = sm[0];
            }
            __syncthreads();
            // ---------------- finish: every warp folds all the chunk's sums itself (lane = point, two rounds for CH > 32)
            unsigned under_mask = 0;                      // bit r: this lane's point of round r underflowed
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int pp = lane + 32 * r;
                if (pp < CH) {
                    const float4* r4 = reinterpret_cast<const float4*>(redp + pp * 16);
                    const float4 r0 = r4[0], r1 = r4[1], r2 = r4[2], r3 = r4[3];
                    const float v = (((r0.x + r0.y) + (r0.z + r0.w)) + ((r1.x + r1.y) + (r1.z + r1.w))) +
                                    (((r2.x + r2.y) + (r2.z + r2.w)) + ((r3.x + r3.y) + (r3.z + r3.w)));
                    const bool valid = pp < ch;
                    const bool under = valid && !(v >= kUnder3);
                    if (under) under_mask |= 1u << r;
                    float inv = (valid && !under) ? __fdividef(1.0f, v) : 0.f;
                    if (norm_eps_on != 0.f || warp == 0) {   // only the PY flavour rescales; only warp 0 keeps the log-lik
                        if (valid && !under) {
                            const float lse2 = cref + lg2f(v);
                            float norm2 = lse2;
                            if (norm_eps_on != 0.f) {        // gmm_impl.py:113  log(sum exp + 1e-8)
                                const float Mx = fmaxf(lse2, kLog2Eps8);
                                norm2 = Mx + lg2f(ex2f(lse2 - Mx) + ex2f(kLog2Eps8 - Mx));
                                inv *= ex2f(lse2 - norm2);
                            }
                            if (warp == 0) {
                                ll += (double)(norm2 * kLn2);
                                nlive += 1.0;
                            }
                        }
                    }
                    myfin[pp] = make_float2(inv, inv);
                }
            }
            const unsigned any_under = __ballot_sync(0xffffffffu, under_mask != 0);   // identical in every warp (same data, same order)
            if (any_under) {
                // ---------------- rare path: exact per-point maximum for the whole chunk
                float* rmax = redslow;
                float* rsum = redslow + CH * 16;
                for (int b = 0; b < ch; b += PB) {
                    float mx[PB];
#pragma unroll
                    for (int p = 0; p < PB; ++p) {
                        const int ip = min(c0 + b + p, cn - 1);
                        const float4 P0 = spts[2 * ip], P1 = spts[2 * ip + 1];
                        float2 dx, dy, dz;
                        const float2 q = quad2(k, make_float2(P0.x, P0.y), make_float2(P0.z, P0.w), make_float2(P1.x, P1.y), dx, dy, dz);
                        mx[p] = fmaxf(fmaxf(q.x, q.y), kNegBig);
                    }
                    reduce_scatter<PB>(mx, lane, true);
                    if (rwriter) rmax[(b + ridx) * 16 + warp] = mx[0];
                }
                __syncthreads();
                for (int b = 0; b < ch; b += PB) {
                    float sm[PB];
#pragma unroll
                    for (int p = 0; p < PB; ++p) {
                        const int ip = min(c0 + b + p, cn - 1);
                        const float4 P0 = spts[2 * ip], P1 = spts[2 * ip + 1];
                        float m = kNegBig;
                        for (int w = 0; w < W; ++w) m = fmaxf(m, rmax[(b + p) * 16 + w]);
                        float2 dx, dy, dz;
                        const float2 q = quad2(k, make_float2(P0.x, P0.y), make_float2(P0.z, P0.w), make_float2(P1.x, P1.y), dx, dy, dz);
                        const float2 e = make_float2(ex2f(q.x - m), ex2f(q.y - m));
                        ecol[(size_t)(b + p) * T] = e;
                        sm[p] = e.x + e.y;
                    }
                    reduce_scatter<PB>(sm, lane, false);
                    if (rwriter) rsum[(b + ridx) * 16 + warp] = sm[0];
                }
                __syncthreads();
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const int pp = lane + 32 * r;
                    if (pp < CH) {
                        float v = 0.f, m = kNegBig;
                        for (int w = 0; w < W; ++w) {
                            v += rsum[pp * 16 + w];
                            m = fmaxf(m, rmax[pp * 16 + w]);
                        }
                        const bool valid = pp < ch;
                        const bool under = (under_mask >> r) & 1u;
                        float inv = 0.f;
                        if (valid && v > 0.f && m > kNegBig) {
                            const float lse2 = cref + m + lg2f(v);
                            float norm2 = lse2, scale = 1.0f;
                            if (norm_eps_on != 0.f) {
                                const float Mx = fmaxf(lse2, kLog2Eps8);
                                norm2 = Mx + lg2f(ex2f(lse2 - Mx) + ex2f(kLog2Eps8 - Mx));
                                scale = ex2f(lse2 - norm2);
                            }
                            inv = scale / v;
                            if (warp == 0 && under) {         // the fast path left only the underflowed points out
                                ll += (double)(norm2 * kLn2);
                                nlive += 1.0;
                            }
                        } else if (valid && under && warp == 0 && norm_eps_on != 0.f) {
                            ll += (double)(kLog2Eps8 * kLn2);  // log(0 + 1e-8)
                        }
                        myfin[pp] = make_float2(inv, inv);
                    }
                }
            }
            __syncwarp();
            // ---------------- pass 2: moments
#pragma unroll 4
            for (int p = 0; p < ch; ++p) {
                const int ip = c0 + p;
                const float4 P0 = spts[2 * ip], P1 = spts[2 * ip + 1];
                const float2 gam = fmul2(ecol[(size_t)p * T], myfin[p]);
                const float2 dx = fadd2(make_float2(P0.x, P0.y), k.nmx);
                const float2 dy = fadd2(make_float2(P0.z, P0.w), k.nmy);
                const float2 dz = fadd2(make_float2(P1.x, P1.y), k.nmz);
                const float2 gx = fmul2(gam, dx), gy = fmul2(gam, dy), gz = fmul2(gam, dz);
                a[0] = fadd2(a[0], gam);
                a[1] = fadd2(a[1], gx);
                a[2] = fadd2(a[2], gy);
                a[3] = fadd2(a[3], gz);
                a[4] = ffma2(gx, dx, a[4]);
                a[5] = ffma2(gx, dy, a[5]);
                a[6] = ffma2(gx, dz, a[6]);
                a[7] = ffma2(gy, dy, a[7]);
                a[8] = ffma2(gy, dz, a[8]);
                a[9] = ffma2(gz, dz, a[9]);
            }
            __syncwarp();                                  // fin[warp] is rewritten by the next chunk's finishing lanes
            parity ^= 1;
        }
    }
    // ---- partial rows: partial[row][m][Jp], row = blockIdx
    float* dst = partial + (size_t)blockIdx.x * kMom * Jp;
    if (live0) {
        const int j = warp * 32 + lane;
#pragma unroll
        for (int m = 0; m < kMom; ++m) dst[(size_t)m * Jp + j] = a[m].x;
    }
    if (live1) {
        const int j = (warp + W) * 32 + lane;
#pragma unroll
        for (int m = 0; m < kMom; ++m) dst[(size_t)m * Jp + j] = a[m].y;
    }
    if (warp == 0) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            ll += __shfl_xor_sync(0xffffffffu, ll, o);
            nlive += __shfl_xor_sync(0xffffffffu, nlive, o);
        }
        if (lane == 0) {
            rowaux[2 * blockIdx.x] = ll;
            rowaux[2 * blockIdx.x + 1] = nlive;
        }
    }
}

// chunk length for T threads: as many points as fit the opt-in shared-memory limit (a multiple of 8, at most 64)
int flat5_chunk(int T, int smem_optin) {
    int ch = 64;
    while (ch > 8 && flat5_smem_bytes(ch, T) > (size_t)smem_optin) ch -= 8;
    return ch;
}

cudaError_t launch_em_flat5(const float* x, const float* y, const float* z, int n, const FlatModel& m, const float* cref_blocks,
                            int W, int grid, float* partial, double* rowaux, const int* done_flag, cudaStream_t s) {
    static int smem_optin = 0;
    if (!smem_optin) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess || smem_optin <= 0)
            smem_optin = 227 * 1024;
        cudaFuncSetAttribute(em_flat5_kernel<416>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin);
        cudaFuncSetAttribute(em_flat5_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin);
    }
    if (W < 1 || W > kMaxWarps5) return cudaErrorInvalidValue;
    const float eps_on = m.flavor == HGMM_FLAVOR_PY ? 1.f : 0.f;
    const int ncref = m.Jp / 32;
    const int T = W * 32;
    const int CH = flat5_chunk(T, smem_optin);
    const size_t smem = flat5_smem_bytes(CH, T);
    if (W <= 13)
        em_flat5_kernel<416><<<grid, T, smem, s>>>(x, y, z, n, m.packed, cref_blocks, ncref, m.J, m.Jp, W, CH, partial, rowaux, done_flag,
                                                    eps_on);
    else
        em_flat5_kernel<512><<<grid, T, smem, s>>>(x, y, z, n, m.packed, cref_blocks, ncref, m.J, m.Jp, W, CH, partial, rowaux, done_flag,
                                                    eps_on);
    return cudaGetLastError();
}

}  // namespace hgmm
