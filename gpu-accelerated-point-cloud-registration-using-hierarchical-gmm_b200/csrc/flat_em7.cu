// flat_em7.cu -- packed-FP32 fused E+M sweep, staged densities, barrier-free chunk pipeline.
//
// Same contract and the same arithmetic per (point, component) pair as em_flat5_kernel (expectationStep +
// maximizationStep of src/c++/gmm_fit/gmm_kernels.cu:278-350; e_step + m_step of
// src/python/gmm_waymo/src/gmm_impl.py:90-116): lane = component pair, q2 / e = 2^(q2 - Cref) in 12 packed operations,
// the ten centred moments in 17, e parked in a lane-private shared-memory column between the two passes.
// What em_flat5 spends outside the FP32 pipe (profiles/r01_em_flat5_ncu_source.txt: 19 % of the samples at the two
// chunk barriers and the single-warp finishing step between them, ~14 integer/address instructions per 64-pair step)
// is removed here:
//
// 1. No CTA-wide barrier in the steady state.  Chunks are double-buffered (e columns and per-column partial sums) and
//    every warp runs
//        wait(c) ; finish(c) ; pass 1 (c+1) ; arrive(c+1) ; pass 2 (c)
//    where arrive/wait are one mbarrier per buffer (count = warps).  A warp that has published its partial sums of
//    chunk c+1 goes straight on to the moment pass of chunk c; it only ever waits if some warp is a whole pass behind.
//    Warps drift apart, so the issue-bound pass 1 of one warp overlaps the pipe-bound pass 2 of another.
// 2. The finishing step (fold the P per-column partial sums of a point, 1/sum, log-likelihood) is done REDUNDANTLY by
//    every warp, lane = point: P conflict-free LDS + P adds per chunk instead of a second barrier.  All warps read the
//    same values and add them in the same order, so they hold bit-identical normalisers -- and agree, without
//    communicating, on whether the chunk needs the exact (max-shifted) path.  Warp 0 alone accumulates the
//    log-likelihood.
// 3. The number of pair columns P is a template parameter: every e / point address in the unrolled batches is
//    base + immediate; the staged points are padded by one batch so no index is clamped.
//
// Buffer reuse is safe by program order alone: a warp writes buffer b for chunk c+2 only after wait(c+1), i.e. after
// every warp arrived for c+1, which each warp does after its finish(c) -- the last reader of buffer b's sums; the
// shared normalisers of buffer b are re-written in finish(c+2), after wait(c+2), which every warp's arrive(c+2)
// precedes and its pass 2 (c) precedes that.
// The rare exact path (a point further than ~11 sigma from every component) uses two ordinary CTA barriers; every warp
// takes it at the same program point for the same chunk.
// Every reduction has a fixed order: fits are bit-reproducible.
#include "common.cuh"
#include "kernels.h"
#include "packed.cuh"

namespace hgmm {

constexpr int kRedPts = 32;             // points per chunk are at most 32 (lane = point in the finishing step)

// dynamic shared memory layout (CH = chunk points, SB = staged points, C = 32 * P):
//   u64 bars[2] | float red[2][16][32] | float rmax[16][32] | float rsum[16][32] | float2 inv[2][32] | float mval[32]
//   | float4 spts[SB + 8][2] | float2 ebuf[2][CH][C]
__host__ __device__ inline size_t flat7_fixed_bytes() { return 16 + 2 * 16 * kRedPts * 4 + 2 * 16 * kRedPts * 4 + 2 * kRedPts * 8 + kRedPts * 4; }
__host__ __device__ inline size_t flat7_smem_bytes(int CH, int SB, int C) {
    return flat7_fixed_bytes() + (size_t)(SB + 8) * 32 + (size_t)2 * CH * C * 8;
}

__device__ __forceinline__ void mbar_arrive7(uint64_t* bar) {
    asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// wait with a long suspend-time hint: the warps that share the left-over column are idle ~3/4 of the time and would otherwise
// spend issue slots of their sub-partition re-polling (profiles/r01_em_flat7_ncu_full.txt: SYNCS + YIELD + BRA = 10 % of the
// instructions issued); the wait still returns as soon as the phase completes
__device__ __forceinline__ void mbar_wait7(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT7_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra WAIT7_DONE;\n\t"
        "bra WAIT7_LOOP;\n\t"
        "WAIT7_DONE:\n\t"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity), "r"(20000u)
        : "memory");
}

template <int P>
__global__ void __launch_bounds__(512, 1) em_flat7_kernel(const float* __restrict__ px, const float* __restrict__ py,
                                                          const float* __restrict__ pz, int n,
                                                          const PackedComp* __restrict__ packed,
                                                          const float* __restrict__ cref_blocks, int n_cref, int Jp, int CH,
                                                          int SB, float* __restrict__ partial, double* __restrict__ rowaux,
                                                          const int* __restrict__ done_flag, float norm_eps_on) {
    constexpr int PB = 8;
    constexpr int C = P * 32;                                              // e columns
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int T = blockDim.x, W = T >> 5;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);                // [2]
    float* red = reinterpret_cast<float*>(bars + 2);                       // [2][16][32]  partial sums, column-major
    float* rmax = red + 2 * 16 * kRedPts;                                  // [16][32]     rare path: column maxima
    float* rsum = rmax + 16 * kRedPts;                                     // [16][32]     rare path: shifted sums
    float2* inv = reinterpret_cast<float2*>(rsum + 16 * kRedPts);          // [2][32]      (1/sum, 1/sum)
    float* mval = reinterpret_cast<float*>(inv + 2 * kRedPts);             // [32]         rare path: exact maxima
    float4* spts = reinterpret_cast<float4*>(mval + kRedPts);              // [SB + 8][2]  (x,x,y,y) (z,z,0,0)
    float2* ebuf = reinterpret_cast<float2*>(spts + (size_t)(SB + 8) * 2); // [2][CH][C]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int S = Jp >> 5;
    const int ridx = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
    const bool rwriter = (lane & 3) == 0;
    if (tid == 0) {
        mbar_init(&bars[0], W);
        mbar_init(&bars[1], W);
        mbar_fence_init();
    }
    const int per = (int)(((long long)n + gridDim.x - 1) / gridDim.x);
    const int lo = min(n, (int)blockIdx.x * per), hi = min(n, lo + per);
    // points of staging block sb0 -> (x,x,y,y) (z,z,0,0), plus one batch of padding (copies of the last point)
    auto stage = [&](int sb0, int cn) {
        for (int i = tid; i < cn + PB; i += T) {
            const int src = sb0 + min(i, cn - 1);
            const float x = px[src], y = py[src], z = pz[src];
            spts[2 * i] = make_float4(x, x, y, y);
            spts[2 * i + 1] = make_float4(z, z, 0.f, 0.f);
        }
    };
    // The cloud never changes between EM iterations, so the first block is staged BEFORE the programmatic-dependent-launch
    // wait: this part of the prologue overlaps the tail of the previous iteration's reduce + finalize kernel.  Everything
    // that kernel writes (done flag, Cref, packed parameters) is read after the wait.
    if (hi > lo) stage(lo, min(SB, hi - lo));
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (__ldcg(done_flag)) return;                        // L2 loads for everything the previous kernel wrote

    // ---- which pair column, and which share of every chunk's batches, this warp sweeps (as em_flat5_kernel)
    const int full = P == W ? P : (P & ~3);               // columns swept whole by one warp
    int col = warp, nsplit = 1, sidx = 0;
    if (warp >= full) {
        const int r = P - full;                           // 1 or 2 left-over columns shared by the last 4 warps
        nsplit = 4 / r;
        col = full + (warp - full) / nsplit;
        sidx = (warp - full) % nsplit;
    }

    float cref = lane < n_cref ? __ldcg(cref_blocks + lane) : -INFINITY;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cref = fmaxf(cref, __shfl_xor_sync(0xffffffffu, cref, o));
    if (!(cref > kNegBig)) cref = 0.f;

    // ---- the lane's component pair -> registers: slots col (low half) and col + P (high half)
    PairParams k;
    const bool live0 = col < S, live1 = col + P < S;
    {
        const float4* a4 = reinterpret_cast<const float4*>(packed + (live0 ? col * 32 + lane : 0));
        const float4* b4 = reinterpret_cast<const float4*>(packed + (live1 ? (col + P) * 32 + lane : 0));
        const float4 a0 = __ldcg(a4), a1 = __ldcg(a4 + 1), a2 = __ldcg(a4 + 2);
        const float4 b0 = __ldcg(b4), b1 = __ldcg(b4 + 1), b2 = __ldcg(b4 + 2);
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
    double ll = 0.0, nlive = 0.0;                         // warp 0, lane = point of the chunk

    float2* const ecol = ebuf + col * 32 + lane;          // element p of buffer b's column: ecol[(b * CH + p) * C]
    float* const redcol = red + col * kRedPts;            // this column's partial sums: redcol[b * 16 * 32 + p]
    unsigned g = 0;                                       // chunks done so far: buffer g & 1, mbarrier parity (g >> 1) & 1

    // pass 1 of the chunk starting at staged point c0 (ch valid points) into buffer b
    auto pass1 = [&](int c0, int ch, int b) {
        const int nb = (ch + PB - 1) / PB;
        const int b_lo = (sidx * nb / nsplit) * PB, b_hi = ((sidx + 1) * nb / nsplit) * PB;
        for (int bb = b_lo; bb < b_hi; bb += PB) {
            const float4* sp = spts + 2 * (c0 + bb);
            float2* eb = ecol + (size_t)(b * CH + bb) * C;
            float sm[PB];
            float2 q[PB];
            // all shared-memory loads of the batch first (an LDS cannot be hoisted over an STS the compiler cannot prove disjoint)
#pragma unroll
            for (int p = 0; p < PB; ++p) {
                const float4 P0 = sp[2 * p], P1 = sp[2 * p + 1];
                float2 dx, dy, dz;
                q[p] = quad2(k, make_float2(P0.x, P0.y), make_float2(P0.z, P0.w), make_float2(P1.x, P1.y), dx, dy, dz);
            }
#pragma unroll
            for (int p = 0; p < PB; ++p) {
                const float2 e = make_float2(ex2f(q[p].x), ex2f(q[p].y));
                eb[p * C] = e;
                sm[p] = e.x + e.y;
            }
            reduce_scatter<PB>(sm, lane, false);
            if (rwriter) redcol[b * 16 * kRedPts + bb + ridx] = sm[0];
        }
    };

    for (int sb0 = lo; sb0 < hi; sb0 += SB) {
        const int cn = min(SB, hi - sb0);
        if (sb0 != lo) {
            __syncthreads();                              // the previous block's points are no longer read
            stage(sb0, cn);
        }
        __syncthreads();                                  // points staged (and, the first time, barriers initialised)
        const int K = (cn + CH - 1) / CH;
        pass1(0, min(CH, cn), g & 1);
        __syncwarp();
        if (lane == 0) mbar_arrive7(&bars[g & 1]);
        for (int c = 0; c < K; ++c, ++g) {
            const int b = g & 1;
            const int c0 = c * CH;
            const int ch = min(CH, cn - c0);
            mbar_wait7(&bars[b], (g >> 1) & 1);
            // ---------------- finish (every warp, lane = point): fold the P column sums in a fixed order
            const bool valid = lane < ch;
            float v = 0.f;
            {
                const float* r = red + b * 16 * kRedPts + lane;
#pragma unroll
                for (int kk = 0; kk < P; ++kk) v += r[kk * kRedPts];
            }
            const bool under = valid && !(v >= kUnder3);
            float iv = (valid && !under) ? __fdividef(1.0f, v) : 0.f;
            if (valid && !under) {
                const float lse2 = cref + lg2f(v);
                float norm2 = lse2;
                if (norm_eps_on != 0.f) {                    // gmm_impl.py:113  log(sum exp + 1e-8)
                    const float Mx = fmaxf(lse2, kLog2Eps8);
                    norm2 = Mx + lg2f(ex2f(lse2 - Mx) + ex2f(kLog2Eps8 - Mx));
                    iv *= ex2f(lse2 - norm2);
                }
                if (warp == 0) {
                    ll += (double)(norm2 * kLn2);
                    nlive += 1.0;
                }
            }
            const bool any_under = __any_sync(0xffffffffu, under);
            if (any_under) {
                // ---------------- rare path: exact per-point maxima for the whole chunk (identical decision in every warp)
                const int nb = (ch + PB - 1) / PB;
                const int b_lo = (sidx * nb / nsplit) * PB, b_hi = ((sidx + 1) * nb / nsplit) * PB;
                for (int bb = b_lo; bb < b_hi; bb += PB) {
                    const float4* sp = spts + 2 * (c0 + bb);
                    float mx[PB];
#pragma unroll
                    for (int p = 0; p < PB; ++p) {
                        const float4 P0 = sp[2 * p], P1 = sp[2 * p + 1];
                        float2 dx, dy, dz;
                        const float2 q = quad2(k, make_float2(P0.x, P0.y), make_float2(P0.z, P0.w), make_float2(P1.x, P1.y), dx, dy, dz);
                        mx[p] = fmaxf(fmaxf(q.x, q.y), kNegBig);
                    }
                    reduce_scatter<PB>(mx, lane, true);
                    if (rwriter) rmax[col * kRedPts + bb + ridx] = mx[0];
                }
                __syncthreads();
                float m = kNegBig;
#pragma unroll
                for (int kk = 0; kk < P; ++kk) m = fmaxf(m, rmax[kk * kRedPts + lane]);
                mval[lane] = m;                           // every warp writes the same values
                __syncwarp();
                for (int bb = b_lo; bb < b_hi; bb += PB) {
                    const float4* sp = spts + 2 * (c0 + bb);
                    float2* eb = ecol + (size_t)(b * CH + bb) * C;
                    float sm[PB];
                    float2 q[PB];
#pragma unroll
                    for (int p = 0; p < PB; ++p) {
                        const float4 P0 = sp[2 * p], P1 = sp[2 * p + 1];
                        float2 dx, dy, dz;
                        q[p] = quad2(k, make_float2(P0.x, P0.y), make_float2(P0.z, P0.w), make_float2(P1.x, P1.y), dx, dy, dz);
                        const float mp = mval[bb + p];
                        q[p].x -= mp;
                        q[p].y -= mp;
                    }
#pragma unroll
                    for (int p = 0; p < PB; ++p) {
                        const float2 e = make_float2(ex2f(q[p].x), ex2f(q[p].y));
                        eb[p * C] = e;
                        sm[p] = e.x + e.y;
                    }
                    reduce_scatter<PB>(sm, lane, false);
                    if (rwriter) rsum[col * kRedPts + bb + ridx] = sm[0];
                }
                __syncthreads();
                float vs = 0.f;
#pragma unroll
                for (int kk = 0; kk < P; ++kk) vs += rsum[kk * kRedPts + lane];
                iv = 0.f;
                if (valid && vs > 0.f && m > kNegBig) {
                    const float lse2 = cref + m + lg2f(vs);
                    float norm2 = lse2, scale = 1.0f;
                    if (norm_eps_on != 0.f) {
                        const float Mx = fmaxf(lse2, kLog2Eps8);
                        norm2 = Mx + lg2f(ex2f(lse2 - Mx) + ex2f(kLog2Eps8 - Mx));
                        scale = ex2f(lse2 - norm2);
                    }
                    iv = scale / vs;
                    if (under && warp == 0) {             // the fast path left only the underflowed points out
                        ll += (double)(norm2 * kLn2);
                        nlive += 1.0;
                    }
                } else if (valid && under && norm_eps_on != 0.f && warp == 0) {
                    ll += (double)(kLog2Eps8 * kLn2);      // log(0 + 1e-8)
                }
            }
            inv[b * kRedPts + lane] = make_float2(iv, iv);   // every warp writes the same values
            __syncwarp();
            // ---------------- pass 1 of the next chunk, published before this chunk's moment pass
            if (c + 1 < K) {
                pass1(c0 + CH, min(CH, cn - c0 - CH), b ^ 1);
                __syncwarp();
                if (lane == 0) mbar_arrive7(&bars[b ^ 1]);
            }
            // ---------------- pass 2: moments of this warp's share of chunk c
            {
                const int nb = (ch + PB - 1) / PB;
                const int b_lo = (sidx * nb / nsplit) * PB, p_hi = min(((sidx + 1) * nb / nsplit) * PB, ch);
                const float4* sp = spts + 2 * c0;
                const float2* eb = ecol + (size_t)b * CH * C;
                const float2* ib = inv + b * kRedPts;
#pragma unroll 4
                for (int p = b_lo; p < p_hi; ++p) {
                    const float4 P0 = sp[2 * p], P1 = sp[2 * p + 1];
                    const float2 gam = fmul2(eb[(size_t)p * C], ib[p]);
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
            }
        }
    }
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");      // the reduce + finalize kernel may start launching
    // ---- warps sharing a column fold their partial moments in a fixed order (the e buffers are free now)
    __syncthreads();
    if (nsplit > 1 && sidx > 0) {
        float2* scratch = ebuf + ((size_t)(col - full) * 3 + (sidx - 1)) * 32 * kMom;
#pragma unroll
        for (int m = 0; m < kMom; ++m) scratch[m * 32 + lane] = a[m];
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
    if (warp == 0) {                                      // fixed-order fold of the 32 per-lane log-likelihood sums
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

int flat5_warps(int P);

// chunk length (multiple of 8, at most 32) and staging block (multiple of 8, at most 512) that fit the shared memory
static void flat7_shape(int C, int smem_optin, int* CH, int* SB) {
    int ch = 32;
    while (ch > 8 && flat7_smem_bytes(ch, 64, C) > (size_t)smem_optin) ch -= 8;
    long long room = (long long)smem_optin - (long long)flat7_smem_bytes(ch, 0, C);
    int sb = (int)(room / 32) / 8 * 8;
    if (sb > 512) sb = 512;
    if (sb < 8) sb = 8;
    *CH = ch;
    *SB = sb;
}

template <int P>
static cudaError_t launch7(const float* x, const float* y, const float* z, int n, const FlatModel& m, const float* cref_blocks,
                           int grid, float* partial, double* rowaux, const int* done_flag, int smem_optin, cudaStream_t s) {
    static DeviceOnce once;      // one per instantiation P
    if (once.first()) {
        cudaError_t e = cudaFuncSetAttribute(em_flat7_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin);
        if (e != cudaSuccess) return e;
    }
    int CH, SB;
    flat7_shape(P * 32, smem_optin, &CH, &SB);
    const int W = flat5_warps(P);
    const float eps_on = m.flavor != HGMM_FLAVOR_CPP ? 1.f : 0.f;
    // programmatic dependent launch: this grid may start while the previous kernel of the stream is still finishing; the
    // kernel itself waits (griddepcontrol.wait) before touching anything that kernel produces
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(W * 32);
    cfg.dynamicSmemBytes = flat7_smem_bytes(CH, SB, P * 32);
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, em_flat7_kernel<P>, x, y, z, n, m.packed, cref_blocks, m.Jp / 32, m.Jp, CH, SB, partial, rowaux,
                              done_flag, eps_on);
}

cudaError_t launch_em_flat7(const float* x, const float* y, const float* z, int n, const FlatModel& m, const float* cref_blocks,
                            int P, int grid, float* partial, double* rowaux, const int* done_flag, cudaStream_t s) {
    const int smem_optin = device_smem_optin();
#define HGMM_F7(PP) case PP: return launch7<PP>(x, y, z, n, m, cref_blocks, grid, partial, rowaux, done_flag, smem_optin, s);
    switch (P) {
        HGMM_F7(5) HGMM_F7(6) HGMM_F7(7) HGMM_F7(8) HGMM_F7(9) HGMM_F7(10) HGMM_F7(11) HGMM_F7(12)
        HGMM_F7(13) HGMM_F7(14) HGMM_F7(15) HGMM_F7(16)
        default: return cudaErrorInvalidValue;
    }
#undef HGMM_F7
}

}  // namespace hgmm
