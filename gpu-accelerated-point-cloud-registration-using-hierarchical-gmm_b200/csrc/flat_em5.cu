// flat_em5.cu -- packed-FP32 fused E+M sweep with the unnormalised densities staged in shared memory.
//
// Same contract as em_flat3_kernel (expectationStep + maximizationStep of src/c++/gmm_fit/gmm_kernels.cu:278-350;
// e_step + m_step of src/python/gmm_waymo/src/gmm_impl.py:90-116) and the same arithmetic per (point, component)
// pair -- 12 packed operations for q2 and e = 2^(q2 - Cref), 17 for the ten centred moments.  Two things differ:
//
// 1. The normalisation over J, which needs every warp of the CTA, is taken once per CHUNK of up to 64 points instead
//    of once per batch of 8: each lane parks its pair of e values for the whole chunk in a private shared-memory
//    column (conflict-free STS.64 / LDS.64; 8 B x columns x chunk points, up to ~190 KB of the SM's 227 KB), so both
//    phases around the two chunk barriers are long, barrier-free FFMA2 streams.
//
//      pass 1 (batches of 8 points): q2, e -> ebuf[point][column]; per-point partial sums -> register
//                                    reduce-scatter -> red[point][pair column]
//      barrier; finish: warp w folds points w, w + W, ... (lane = pair column, xor-shuffle tree) -> inv[point]
//      barrier; pass 2 (per point): gamma = e * inv, ten centred moments in float2 registers
//
// 2. Warps are spread evenly over the SM's four sub-partitions.  P = ceil(S/2) pair columns (S = Jp/32 slots) give
//    P warps if every warp sweeps all points; at J = 800 that is 13 warps = 4+3+3+3, and the sub-partition with four
//    runs 23 % longer than the mean (profiles/microbench/fp32_pipe.cu measures exactly 13/16 of the 16-warp rate).
//    Here the P mod 4 left-over columns are each shared by 4 (or 2) warps that split the chunk's 8-point batches
//    between them, so P = 4a + 1 or 4a + 2 runs 4a + 4 warps with identical load on every sub-partition; the
//    sharing warps fold their partial moments through shared memory in a fixed order at the end.
//
// A chunk containing a point whose sum underflows the fixed reference (2^-100: further than ~11 sigma from every
// component) is redone with exact per-point maxima (CTA-uniform branch, four more barriers).
// Every reduction has a fixed order: fits are bit-reproducible.
#include "common.cuh"
#include "kernels.h"
#include "packed.cuh"

namespace hgmm {

// dynamic shared memory layout (CH = chunk points, multiple of 8; C = 32 * pair columns):
//   float4 spts[kChunk3][2] | float red[CH][16] | float2 inv[CH] | float mval[CH] | float uflag[CH] | double wsum[16][2]
//   | int flags[4] | float2 ebuf[CH][C]
__host__ __device__ inline size_t flat5_smem_bytes(int CH, int C) {
    return (size_t)kChunk3 * 32 + (size_t)CH * (64 + 8 + 4 + 4) + 16 * 16 + 16 + (size_t)CH * C * 8;
}

template <int MAXT>
__global__ void __launch_bounds__(MAXT, 1) em_flat5_kernel(const float* __restrict__ px, const float* __restrict__ py,
                                                           const float* __restrict__ pz, int n,
                                                           const PackedComp* __restrict__ packed,
                                                           const float* __restrict__ cref_blocks, int n_cref, int J, int Jp,
                                                           int P, int CH, float* __restrict__ partial,
                                                           double* __restrict__ rowaux, const int* __restrict__ done_flag,
                                                           float norm_eps_on) {
    if (*done_flag) return;
    constexpr int PB = 8;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int T = blockDim.x, W = T >> 5;
    const int C = P * 32;                                                  // ebuf columns
    float4* spts = reinterpret_cast<float4*>(smem_raw);                    // [kChunk3][2]  (x,x,y,y) (z,z,0,0)
    float* red = reinterpret_cast<float*>(spts + kChunk3 * 2);             // [CH][16]
    float2* inv = reinterpret_cast<float2*>(red + CH * 16);                // [CH]  (inv, inv)
    float* mval = reinterpret_cast<float*>(inv + CH);                      // [CH]  rare path: exact maxima
    float* uflag = mval + CH;                                              // [CH]  1 = the fixed-reference sum underflowed
    double* wsum = reinterpret_cast<double*>(uflag + CH);                  // [16][2]
    int* flags = reinterpret_cast<int*>(wsum + 32);                        // [4]
    float2* ebuf = reinterpret_cast<float2*>(flags + 4);                   // [CH][C]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int S = Jp >> 5;
    const int ridx = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
    const bool rwriter = (lane & 3) == 0;
    if (tid < 4) flags[tid] = 0;

    // ---- which pair column, and which share of every chunk's batches, this warp sweeps
    const int full = P == W ? P : (P & ~3);               // columns swept whole by one warp
    int col = warp, nsplit = 1, sidx = 0;
    if (warp >= full) {
        const int r = P - full;                           // 1 or 2 left-over columns shared by the last 4 warps
        nsplit = 4 / r;
        col = full + (warp - full) / nsplit;
        sidx = (warp - full) % nsplit;
    }

    float cref = lane < n_cref ? __ldg(cref_blocks + lane) : -INFINITY;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cref = fmaxf(cref, __shfl_xor_sync(0xffffffffu, cref, o));
    if (!(cref > kNegBig)) cref = 0.f;

    // ---- the lane's component pair -> registers: slots col (low half) and col + P (high half)
    PairParams k;
    const bool live0 = col < S, live1 = col + P < S;
    {
        const float4* a4 = reinterpret_cast<const float4*>(packed + (live0 ? col * 32 + lane : 0));
        const float4* b4 = reinterpret_cast<const float4*>(packed + (live1 ? (col + P) * 32 + lane : 0));
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
    double ll = 0.0, nlive = 0.0;                         // lane 0 of every warp, for the points it finishes

    const int per = (int)(((long long)n + gridDim.x - 1) / gridDim.x);
    const int lo = min(n, (int)blockIdx.x * per), hi = min(n, lo + per);
    int parity = 0;
    float2* ecol = ebuf + col * 32 + lane;                // element p of the column is ecol[p * C]

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
            const int ch = min(CH, cn - c0);              // valid points of this chunk
            const int nb = (ch + PB - 1) / PB;            // its 8-point batches
            const int b_lo = (sidx * nb / nsplit) * PB, b_hi = ((sidx + 1) * nb / nsplit) * PB;   // this warp's share
            const int p_hi = min(b_hi, ch);
            // ---------------- pass 1: e -> ebuf, per-point partial sums -> red
            for (int b = b_lo; b < b_hi; b += PB) {
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
                    ecol[(size_t)(b + p) * C] = e;
                    sm[p] = e.x + e.y;
                }
                reduce_scatter<PB>(sm, lane, false);
                if (rwriter) red[(b + ridx) * 16 + col] = sm[0];
            }
            __syncthreads();
            // ---------------- finish: warp w folds the partial sums of points w, w + W, ... (lane = pair column)
            for (int pt = warp; pt < nb * PB; pt += W) {
                float v = lane < P ? red[pt * 16 + lane] : 0.f;
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
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
                    inv[pt] = make_float2(iv, iv);
                    uflag[pt] = under ? 1.f : 0.f;
                    if (under) flags[parity] = 1;
                }
            }
            __syncthreads();
            const bool any_under = flags[parity] != 0;
            if (tid == 0) flags[parity ^ 1] = 0;          // the next chunk's flag; its last readers are past the barrier above
            if (any_under) {
                // ---------------- rare path: exact per-point maxima for the whole chunk (CTA-uniform branch)
                for (int b = b_lo; b < b_hi; b += PB) {
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
                    if (rwriter) red[(b + ridx) * 16 + col] = mx[0];
                }
                __syncthreads();
                for (int pt = warp; pt < nb * PB; pt += W) {
                    float v = lane < P ? red[pt * 16 + lane] : kNegBig;
#pragma unroll
                    for (int o = 8; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
                    if (lane == 0) mval[pt] = v;
                }
                __syncthreads();
                for (int b = b_lo; b < b_hi; b += PB) {
                    float sm[PB];
                    float2 q[PB];
#pragma unroll
                    for (int p = 0; p < PB; ++p) {
                        const int ip = min(c0 + b + p, cn - 1);
                        const float4 P0 = spts[2 * ip], P1 = spts[2 * ip + 1];
                        float2 dx, dy, dz;
                        q[p] = quad2(k, make_float2(P0.x, P0.y), make_float2(P0.z, P0.w), make_float2(P1.x, P1.y), dx, dy, dz);
                        const float m = mval[b + p];
                        q[p].x -= m;
                        q[p].y -= m;
                    }
#pragma unroll
                    for (int p = 0; p < PB; ++p) {
                        const float2 e = make_float2(ex2f(q[p].x), ex2f(q[p].y));
                        ecol[(size_t)(b + p) * C] = e;
                        sm[p] = e.x + e.y;
                    }
                    reduce_scatter<PB>(sm, lane, false);
                    if (rwriter) red[(b + ridx) * 16 + col] = sm[0];
                }
                __syncthreads();
                for (int pt = warp; pt < nb * PB; pt += W) {
                    float v = lane < P ? red[pt * 16 + lane] : 0.f;
#pragma unroll
                    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
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
                        inv[pt] = make_float2(iv, iv);
                    }
                }
                __syncthreads();
            }
            // ---------------- pass 2: moments of this warp's share of the chunk
#pragma unroll 4
            for (int p = b_lo; p < p_hi; ++p) {
                const int ip = c0 + p;
                const float4 P0 = spts[2 * ip], P1 = spts[2 * ip + 1];
                const float2 gam = fmul2(ecol[(size_t)p * C], inv[p]);
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
            parity ^= 1;
        }
    }
    // ---- warps sharing a column fold their partial moments in a fixed order (the e buffer is free now)
    __syncthreads();
    if (nsplit > 1 && sidx > 0) {
        float2* scratch = ebuf + ((size_t)(col - full) * 3 + (sidx - 1)) * 32 * kMom;
#pragma unroll
        for (int m = 0; m < kMom; ++m) scratch[m * 32 + lane] = a[m];
    }
    if (lane == 0) {
        wsum[2 * warp] = ll;
        wsum[2 * warp + 1] = nlive;
    }
    __syncthreads();
    if (nsplit > 1 && sidx == 0) {
        for (int s2 = 1; s2 < nsplit; ++s2) {
            const float2* scratch = ebuf + ((size_t)(col - full) * 3 + (s2 - 1)) * 32 * kMom;
#pragma unroll
            for (int m = 0; m < kMom; ++m) a[m] = fadd2(a[m], scratch[m * 32 + lane]);
        }
    }
    // ---- partial rows: partial[row][m][Jp], row = blockIdx
    if (sidx == 0) {
        float* dst = partial + (size_t)blockIdx.x * kMom * Jp;
        if (live0) {
            const int j = col * 32 + lane;
#pragma unroll
            for (int m = 0; m < kMom; ++m) dst[(size_t)m * Jp + j] = a[m].x;
        }
        if (live1) {
            const int j = (col + P) * 32 + lane;
#pragma unroll
            for (int m = 0; m < kMom; ++m) dst[(size_t)m * Jp + j] = a[m].y;
        }
    }
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

// warps for P pair columns: the P mod 4 = 1 or 2 left-over columns are shared by four warps (see the header)
int flat5_warps(int P) {
    const int r = P & 3;
    if ((r == 1 || r == 2) && (P & ~3) + 4 <= 16) return (P & ~3) + 4;
    return P;
}

// chunk length for C columns: as many points as fit the opt-in shared-memory limit (a multiple of 8, at most 64)
static int flat5_chunk(int C, int smem_optin) {
    int ch = 64;
    while (ch > 8 && flat5_smem_bytes(ch, C) > (size_t)smem_optin) ch -= 8;
    return ch;
}

cudaError_t launch_em_flat5(const float* x, const float* y, const float* z, int n, const FlatModel& m, const float* cref_blocks,
                            int P, int grid, float* partial, double* rowaux, const int* done_flag, cudaStream_t s) {
    const int smem_optin = device_smem_optin();
    static DeviceOnce once;
    if (once.first()) {
        cudaFuncSetAttribute(em_flat5_kernel<416>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin);
        cudaFuncSetAttribute(em_flat5_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin);
    }
    if (P < 1 || P > 16) return cudaErrorInvalidValue;
    const float eps_on = m.flavor != HGMM_FLAVOR_CPP ? 1.f : 0.f;
    const int ncref = m.Jp / 32;
    const int W = flat5_warps(P);
    const int T = W * 32;
    const int CH = flat5_chunk(P * 32, smem_optin);
    const size_t smem = flat5_smem_bytes(CH, P * 32);
    if (W <= 13)
        em_flat5_kernel<416><<<grid, T, smem, s>>>(x, y, z, n, m.packed, cref_blocks, ncref, m.J, m.Jp, P, CH, partial, rowaux, done_flag,
                                                    eps_on);
    else
        em_flat5_kernel<512><<<grid, T, smem, s>>>(x, y, z, n, m.packed, cref_blocks, ncref, m.J, m.Jp, P, CH, partial, rowaux, done_flag,
                                                    eps_on);
    return cudaGetLastError();
}

}  // namespace hgmm
