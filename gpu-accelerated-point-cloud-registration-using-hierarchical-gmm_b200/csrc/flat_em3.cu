// flat_em3.cu -- packed-FP32 (FFMA2) fused E+M sweep of the flat mixture: the default path for J > 32.
//
// Same contract and data flow as flat_em2.cu's single-evaluation kernel (expectationStep +
// maximizationStep of src/c++/gmm_fit/gmm_kernels.cu:278-350; e_step + m_step of
// src/python/gmm_waymo/src/gmm_impl.py:90-116), re-expressed for Blackwell's packed FP32 pipe:
// a 3-register FFMA issues at half rate on sm_100, the two-wide FFMA2/FADD2/FMUL2 (PTX *.f32x2)
// restore the full 128 FMA/clk/SM.  Every lane therefore owns a PAIR of components (slots sw and
// sw + Sdiv) whose parameters, e = 2^(q - Cref) values and 10 centred moment accumulators are
// float2 registers; point coordinates are staged in shared memory already duplicated (x,x,y,y,z,z)
// so a broadcast LDS.128 pair feeds both halves.
//
// Per batch of 8 points: pass 1 (q, e, per-point partial sums) -> register reduce-scatter ->
// partial sums to shared memory -> ONE group barrier -> every warp redundantly finishes the 8
// sums (no second barrier, partial buffers double-buffered by batch parity) -> pass 2 (moments).
// Batches containing a point whose sum underflows the fixed reference take an exact max-shifted
// path.  Partial moment rows + fixed-order fp64 reduction are shared with flat_em2.cu.
#include <stdlib.h>
#include "common.cuh"
#include "kernels.h"
#include "packed.cuh"


namespace hgmm {


// grid.x CTAs; blockDim.x = 32 * G * Sdiv; G independent groups of Sdiv warps; warp sw of a group owns the
// component slots sw (low half of every pair) and sw + Sdiv (high half); lane = component inside the slot.
template <int MAXT, int MINB, int PB, int NP>
__global__ void __launch_bounds__(MAXT, MINB) em_flat3_kernel(const float* __restrict__ px, const float* __restrict__ py,
                                                              const float* __restrict__ pz, int n,
                                                              const PackedComp* __restrict__ packed,
                                                              const float* __restrict__ cref_blocks, int n_cref, int J, int Jp,
                                                              int Sdiv, int G, float* __restrict__ partial,
                                                              double* __restrict__ rowaux, const int* __restrict__ done_flag,
                                                              float norm_eps_on) {
    if (*done_flag) return;
    __shared__ __align__(16) float4 spts[kChunk3][2];            // (x,x,y,y) (z,z,0,0)
    __shared__ __align__(16) float red[2][8][PB][16];            // [batch parity][group][point][warp of group]
    __shared__ __align__(16) float2 fin[16][PB];                 // [warp][point] (inv, inv) -- private to each warp
    __shared__ __align__(16) float finmax[8][PB];                // [group][point] exact maximum (rare path)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 2 * 8 * PB * 16; i += blockDim.x) (&red[0][0][0][0])[i] = 0.f;     // slots >= Sdiv stay zero
    const int g = warp / Sdiv, sw = warp - g * Sdiv;
    const int gthreads = Sdiv * 32;
    const int S = Jp >> 5;
    const int ridx = PB == 8 ? ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1) : ((lane >> 4) & 1) * 2 + ((lane >> 3) & 1);
    const bool rwriter = PB == 8 ? (lane & 3) == 0 : (lane & 7) == 0;

    float cref = -INFINITY;
    for (int i = 0; i < n_cref; ++i) cref = fmaxf(cref, __ldg(cref_blocks + i));
    if (!(cref > kNegBig)) cref = 0.f;

    // ---- the lane's NP component pairs -> registers; pair u = slots (sw + 2u*Sdiv, sw + (2u+1)*Sdiv)
    PairParams k[NP];
    bool live0[NP], live1[NP];
#pragma unroll
    for (int u = 0; u < NP; ++u) {
        const int s0 = sw + (2 * u) * Sdiv, s1 = sw + (2 * u + 1) * Sdiv;
        live0[u] = s0 < S;
        live1[u] = s1 < S;
        const float4* a4 = reinterpret_cast<const float4*>(packed + (live0[u] ? s0 * 32 + lane : 0));
        const float4* b4 = reinterpret_cast<const float4*>(packed + (live1[u] ? s1 * 32 + lane : 0));
        const float4 a0 = __ldg(a4), a1 = __ldg(a4 + 1), a2 = __ldg(a4 + 2);
        const float4 b0 = __ldg(b4), b1 = __ldg(b4 + 1), b2 = __ldg(b4 + 2);
        k[u].nmx = make_float2(-a0.x, -b0.x);
        k[u].nmy = make_float2(-a0.y, -b0.y);
        k[u].nmz = make_float2(-a0.z, -b0.z);
        k[u].c2 = make_float2(live0[u] ? a0.w - cref : -INFINITY, live1[u] ? b0.w - cref : -INFINITY);
        k[u].axx = make_float2(a1.x, b1.x);
        k[u].ayy = make_float2(a1.y, b1.y);
        k[u].azz = make_float2(a1.z, b1.z);
        k[u].axy = make_float2(a1.w, b1.w);
        k[u].axz = make_float2(a2.x, b2.x);
        k[u].ayz = make_float2(a2.y, b2.y);
    }
    float2 a[NP][kMom];
#pragma unroll
    for (int u = 0; u < NP; ++u)
#pragma unroll
        for (int m = 0; m < kMom; ++m) a[u][m] = make_float2(0.f, 0.f);
    double ll = 0.0, nlive = 0.0;                         // accumulated by the finishing lanes of warp sw == 0

    const int per = (int)(((long long)n + gridDim.x - 1) / gridDim.x);
    const int lo = min(n, (int)blockIdx.x * per), hi = min(n, lo + per);
    int parity = 0;

    for (int cb = lo; cb < hi; cb += kChunk3) {
        const int cn = min(kChunk3, hi - cb);
        __syncthreads();
        for (int i = tid; i < cn; i += blockDim.x) {
            const float x = px[cb + i], y = py[cb + i], z = pz[cb + i];
            spts[i][0] = make_float4(x, x, y, y);
            spts[i][1] = make_float4(z, z, 0.f, 0.f);
        }
        __syncthreads();
        const int gper = (cn + G - 1) / G;
        const int gs = min(cn, g * gper), ge = min(cn, gs + gper);
        for (int b = gs; b < ge; b += PB) {
            const int np = min(PB, ge - b);
            float2 e[NP][PB];
            float sm[PB];
            // ---------------- pass 1
#pragma unroll
            for (int p = 0; p < PB; ++p) {
                const int ip = min(b + p, cn - 1);
                const float4 P0 = spts[ip][0], P1 = spts[ip][1];
                float s = 0.f;
#pragma unroll
                for (int u = 0; u < NP; ++u) {
                    float2 dx, dy, dz;
                    const float2 q = quad2(k[u], make_float2(P0.x, P0.y), make_float2(P0.z, P0.w), make_float2(P1.x, P1.y), dx, dy, dz);
                    e[u][p] = make_float2(ex2f(q.x), ex2f(q.y));
                    s += e[u][p].x + e[u][p].y;
                }
                sm[p] = s;
            }
            reduce_scatter<PB>(sm, lane, false);
            if (rwriter) red[parity][g][ridx][sw] = sm[0];
            group_bar3(1 + g, gthreads);
            // ---------------- every warp finishes the 8 sums itself (lanes 0..7), no second barrier.
            // red[..][point][0..16) is contiguous and zero beyond Sdiv: 4 LDS.128 + 15 FADD, no loop.
            bool under = false;
            if (lane < PB) {
                const float4* r4 = reinterpret_cast<const float4*>(&red[parity][g][lane][0]);
                const float4 r0 = r4[0], r1 = r4[1], r2 = r4[2], r3 = r4[3];
                const float v = (((r0.x + r0.y) + (r0.z + r0.w)) + ((r1.x + r1.y) + (r1.z + r1.w))) +
                                (((r2.x + r2.y) + (r2.z + r2.w)) + ((r3.x + r3.y) + (r3.z + r3.w)));
                const bool valid = lane < np;
                under = valid && !(v >= kUnder3);
                float inv = (valid && !under) ? __fdividef(1.0f, v) : 0.f;
                if (norm_eps_on != 0.f || sw == 0) {       // only the PY flavour rescales; only warp 0 of the group keeps the log-lik
                    if (valid && !under) {
                        const float lse2 = cref + lg2f(v);
                        float norm2 = lse2;
                        if (norm_eps_on != 0.f) {          // gmm_impl.py:113  log(sum exp + 1e-8)
                            const float Mx = fmaxf(lse2, kLog2Eps8);
                            norm2 = Mx + lg2f(ex2f(lse2 - Mx) + ex2f(kLog2Eps8 - Mx));
                            inv *= ex2f(lse2 - norm2);
                        }
                        if (sw == 0) {
                            ll += (double)(norm2 * kLn2);
                            nlive += 1.0;
                        }
                    }
                }
                fin[warp][lane] = make_float2(inv, inv);
            }
            const unsigned any_under = __ballot_sync(0xffffffffu, under);
            if (any_under) {
                // ---------------- rare path: exact per-point maximum
                float mx[PB];
#pragma unroll
                for (int p = 0; p < PB; ++p) {
                    const int ip = min(b + p, cn - 1);
                    const float4 P0 = spts[ip][0], P1 = spts[ip][1];
                    float m = kNegBig;
#pragma unroll
                    for (int u = 0; u < NP; ++u) {
                        float2 dx, dy, dz;
                        const float2 q = quad2(k[u], make_float2(P0.x, P0.y), make_float2(P0.z, P0.w), make_float2(P1.x, P1.y), dx, dy, dz);
                        e[u][p] = q;
                        m = fmaxf(m, fmaxf(q.x, q.y));
                    }
                    mx[p] = m;
                }
                reduce_scatter<PB>(mx, lane, true);
                group_bar3(1 + g, gthreads);               // everyone is done reading red[parity] (first use)
                if (rwriter) red[parity][g][ridx][sw] = mx[0];
                group_bar3(1 + g, gthreads);
                if (sw == 0 && lane < PB) {
                    float v = kNegBig;
                    for (int w = 0; w < Sdiv; ++w) v = fmaxf(v, red[parity][g][lane][w]);
                    finmax[g][lane] = v;
                }
                group_bar3(1 + g, gthreads);
#pragma unroll
                for (int p = 0; p < PB; ++p) {
                    const float m = finmax[g][p];
                    float sacc = 0.f;
#pragma unroll
                    for (int u = 0; u < NP; ++u) {
                        e[u][p] = make_float2(ex2f(e[u][p].x - m), ex2f(e[u][p].y - m));
                        sacc += e[u][p].x + e[u][p].y;
                    }
                    sm[p] = sacc;
                }
                reduce_scatter<PB>(sm, lane, false);
                if (rwriter) red[parity][g][ridx][sw] = sm[0];     // maxima were consumed before the last barrier
                group_bar3(1 + g, gthreads);
                {
                    float v = 0.f;
                    if (lane < PB) {
                        for (int w = 0; w < Sdiv; ++w) v += red[parity][g][lane][w];
                    }
                    const float m = lane < PB ? finmax[g][lane] : 0.f;
                    const bool valid = lane < np;
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
                        if (sw == 0 && under) {            // the fast path left only the underflowed points out
                            ll += (double)(norm2 * kLn2);
                            nlive += 1.0;
                        }
                    } else if (valid && under && sw == 0 && norm_eps_on != 0.f) {
                        ll += (double)(kLog2Eps8 * kLn2);   // log(0 + 1e-8)
                    }
                    if (lane < PB) fin[warp][lane] = make_float2(inv, inv);
                }
            }
            __syncwarp();
            // ---------------- pass 2: moments
#pragma unroll
            for (int p = 0; p < PB; ++p) {
                const int ip = min(b + p, cn - 1);
                const float4 P0 = spts[ip][0], P1 = spts[ip][1];
                const float2 inv2 = fin[warp][p];
                const float2 X = make_float2(P0.x, P0.y), Y = make_float2(P0.z, P0.w), Z = make_float2(P1.x, P1.y);
#pragma unroll
                for (int u = 0; u < NP; ++u) {
                    const float2 gam = fmul2(e[u][p], inv2);
                    const float2 dx = fadd2(X, k[u].nmx);
                    const float2 dy = fadd2(Y, k[u].nmy);
                    const float2 dz = fadd2(Z, k[u].nmz);
                    const float2 gx = fmul2(gam, dx), gy = fmul2(gam, dy), gz = fmul2(gam, dz);
                    a[u][0] = fadd2(a[u][0], gam);
                    a[u][1] = fadd2(a[u][1], gx);
                    a[u][2] = fadd2(a[u][2], gy);
                    a[u][3] = fadd2(a[u][3], gz);
                    a[u][4] = ffma2(gx, dx, a[u][4]);
                    a[u][5] = ffma2(gx, dy, a[u][5]);
                    a[u][6] = ffma2(gx, dz, a[u][6]);
                    a[u][7] = ffma2(gy, dy, a[u][7]);
                    a[u][8] = ffma2(gy, dz, a[u][8]);
                    a[u][9] = ffma2(gz, dz, a[u][9]);
                }
            }
            __syncwarp();                                  // fin[warp] is rewritten by the next batch's finishing lanes
            parity ^= 1;
        }
    }
    // ---- partial rows: partial[row][m][Jp], row = blockIdx * G + g
    const size_t row = (size_t)blockIdx.x * G + g;
    float* dst = partial + row * (size_t)kMom * Jp;
#pragma unroll
    for (int u = 0; u < NP; ++u) {
        if (live0[u]) {
            const int j = (sw + (2 * u) * Sdiv) * 32 + lane;
#pragma unroll
            for (int m = 0; m < kMom; ++m) dst[(size_t)m * Jp + j] = a[u][m].x;
        }
        if (live1[u]) {
            const int j = (sw + (2 * u + 1) * Sdiv) * 32 + lane;
#pragma unroll
            for (int m = 0; m < kMom; ++m) dst[(size_t)m * Jp + j] = a[u][m].y;
        }
    }
    if (sw == 0) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            ll += __shfl_xor_sync(0xffffffffu, ll, o);
            nlive += __shfl_xor_sync(0xffffffffu, nlive, o);
        }
        if (lane == 0) {
            rowaux[2 * row] = ll;
            rowaux[2 * row + 1] = nlive;
        }
    }
}

// ------------------------------------------------------------------------------------------
// em_flat4_kernel: em_flat3_kernel<.., PB=8, NP=1> with the batch loop software-pipelined.
// The group barrier of batch k+1 is split into an mbarrier ARRIVE (right after pass 1 of batch k+1 published its
// partial sums) and a WAIT placed after pass 2 of batch k, so the barrier latency and the finishing step hide behind
// FFMA2 work and the warps of a CTA drift apart instead of marching through FMA-heavy and idle phases together.
// e = 2^(q - Cref) is double-buffered in registers (two batches in flight), the partial-sum buffers are 4-deep
// (a fast warp may publish batch k+2 while a slow one still folds batch k-1), two mbarriers alternate by batch parity.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int MAXT>
__global__ void __launch_bounds__(MAXT, 1) em_flat4_kernel(const float* __restrict__ px, const float* __restrict__ py,
                                                           const float* __restrict__ pz, int n,
                                                           const PackedComp* __restrict__ packed,
                                                           const float* __restrict__ cref_blocks, int n_cref, int J, int Jp,
                                                           int Sdiv, int G, float* __restrict__ partial,
                                                           double* __restrict__ rowaux, const int* __restrict__ done_flag,
                                                           float norm_eps_on) {
    if (*done_flag) return;
    constexpr int PB = 8;
    __shared__ __align__(16) float4 spts[kChunk3][2];            // (x,x,y,y) (z,z,0,0)
    __shared__ __align__(16) float red[4][8][PB][16];            // [batch & 3][group][point][warp of group]
    __shared__ __align__(16) float redslow[8][PB][16];           // rare path scratch
    __shared__ __align__(16) float2 fin[16][PB];                 // [warp][point] (inv, inv), private to each warp
    __shared__ __align__(16) float finmax[8][PB];
    __shared__ uint64_t bars[8][2];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 4 * 8 * PB * 16; i += blockDim.x) (&red[0][0][0][0])[i] = 0.f;
    for (int i = tid; i < 8 * PB * 16; i += blockDim.x) (&redslow[0][0][0])[i] = 0.f;
    const int g = warp / Sdiv, sw = warp - g * Sdiv;
    const int gthreads = Sdiv * 32;
    const int S = Jp >> 5;
    const int ridx = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
    const bool rwriter = (lane & 3) == 0;
    if (tid < 16) mbar_init(&bars[tid >> 1][tid & 1], Sdiv);
    if (tid == 0) mbar_fence_init();

    float cref = -INFINITY;
    for (int i = 0; i < n_cref; ++i) cref = fmaxf(cref, __ldg(cref_blocks + i));
    if (!(cref > kNegBig)) cref = 0.f;

    PairParams k;
    bool live0, live1;
    {
        const int s0 = sw, s1 = sw + Sdiv;
        live0 = s0 < S;
        live1 = s1 < S;
        const float4* a4 = reinterpret_cast<const float4*>(packed + (live0 ? s0 * 32 + lane : 0));
        const float4* b4 = reinterpret_cast<const float4*>(packed + (live1 ? s1 * 32 + lane : 0));
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
    double ll = 0.0, nlive = 0.0;

    const int per = (int)(((long long)n + gridDim.x - 1) / gridDim.x);
    const int lo = min(n, (int)blockIdx.x * per), hi = min(n, lo + per);
    unsigned kb = 0;                                     // running batch counter of this group (selects buffers / parities)

    // pass 1 of the batch starting at `b`: e -> E[], partial sums -> red[kk & 3], arrive on bars[g][kk & 1]
#define HGMM_PASS1(E, b, kk)                                                                                           \
    {                                                                                                                  \
        float sm_[PB];                                                                                                 \
        _Pragma("unroll") for (int p = 0; p < PB; ++p) {                                                             \
            const int ip = min((b) + p, cn - 1);                                                                       \
            const float4 P0 = spts[ip][0], P1 = spts[ip][1];                                                          \
            float2 dx, dy, dz;                                                                                         \
            const float2 q = quad2(k, make_float2(P0.x, P0.y), make_float2(P0.z, P0.w), make_float2(P1.x, P1.y), dx, dy, dz); \
            E[p] = make_float2(ex2f(q.x), ex2f(q.y));                                                                  \
            sm_[p] = E[p].x + E[p].y;                                                                                  \
        }                                                                                                              \
        reduce_scatter<PB>(sm_, lane, false);                                                                          \
        if (rwriter) red[(kk) & 3][g][ridx][sw] = sm_[0];                                                              \
        __syncwarp();                                                                                                  \
        if (lane == 0) mbar_arrive(&bars[g][(kk) & 1]);                                                                \
    }

    for (int cb = lo; cb < hi; cb += kChunk3) {
        const int cn = min(kChunk3, hi - cb);
        __syncthreads();                                   // previous chunk fully consumed (also orders the inits above)
        for (int i = tid; i < cn; i += blockDim.x) {
            const float x = px[cb + i], y = py[cb + i], z = pz[cb + i];
            spts[i][0] = make_float4(x, x, y, y);
            spts[i][1] = make_float4(z, z, 0.f, 0.f);
        }
        __syncthreads();
        const int gper = (cn + G - 1) / G;
        const int gs = min(cn, g * gper), ge = min(cn, gs + gper);
        if (gs >= ge) continue;
        float2 e[2][PB];
        HGMM_PASS1(e[0], gs, kb)
        for (int b = gs; b < ge; b += 2 * PB) {
#pragma unroll
            for (int u = 0; u < 2; ++u) {                  // u is a compile-time constant after unrolling: e[u] stays in registers
                const int bc = b + u * PB;                 // current batch
                if (bc < ge) {
                    const int bn = bc + PB;                // next batch
                    if (bn < ge) HGMM_PASS1(e[u ^ 1], bn, kb + 1)
                    mbar_wait(&bars[g][kb & 1], (kb >> 1) & 1);
                    const int np = min(PB, ge - bc);
                    // ---------------- finishing (every warp, lanes 0..7)
                    bool under = false;
                    if (lane < PB) {
                        const float4* r4 = reinterpret_cast<const float4*>(&red[kb & 3][g][lane][0]);
                        const float4 r0 = r4[0], r1 = r4[1], r2 = r4[2], r3 = r4[3];
                        const float v = (((r0.x + r0.y) + (r0.z + r0.w)) + ((r1.x + r1.y) + (r1.z + r1.w))) +
                                        (((r2.x + r2.y) + (r2.z + r2.w)) + ((r3.x + r3.y) + (r3.z + r3.w)));
                        const bool valid = lane < np;
                        under = valid && !(v >= kUnder3);
                        float inv = (valid && !under) ? __fdividef(1.0f, v) : 0.f;
                        if (norm_eps_on != 0.f || sw == 0) {
                            if (valid && !under) {
                                const float lse2 = cref + lg2f(v);
                                float norm2 = lse2;
                                if (norm_eps_on != 0.f) {          // gmm_impl.py:113  log(sum exp + 1e-8)
                                    const float Mx = fmaxf(lse2, kLog2Eps8);
                                    norm2 = Mx + lg2f(ex2f(lse2 - Mx) + ex2f(kLog2Eps8 - Mx));
                                    inv *= ex2f(lse2 - norm2);
                                }
                                if (sw == 0) {
                                    ll += (double)(norm2 * kLn2);
                                    nlive += 1.0;
                                }
                            }
                        }
                        fin[warp][lane] = make_float2(inv, inv);
                    }
                    const unsigned any_under = __ballot_sync(0xffffffffu, under);
                    if (any_under) {
                        // ---------------- rare path: exact per-point maximum, own scratch + classic group barriers
                        float mx[PB], sm2[PB];
#pragma unroll
                        for (int p = 0; p < PB; ++p) {
                            const int ip = min(bc + p, cn - 1);
                            const float4 P0 = spts[ip][0], P1 = spts[ip][1];
                            float2 dx, dy, dz;
                            const float2 q = quad2(k, make_float2(P0.x, P0.y), make_float2(P0.z, P0.w), make_float2(P1.x, P1.y), dx, dy, dz);
                            e[u][p] = q;
                            mx[p] = fmaxf(fmaxf(q.x, q.y), kNegBig);
                        }
                        reduce_scatter<PB>(mx, lane, true);
                        group_bar3(1 + g, gthreads);               // scratch free (previous rare batch fully consumed)
                        if (rwriter) redslow[g][ridx][sw] = mx[0];
                        group_bar3(1 + g, gthreads);
                        if (sw == 0 && lane < PB) {
                            float v = kNegBig;
                            for (int w = 0; w < Sdiv; ++w) v = fmaxf(v, redslow[g][lane][w]);
                            finmax[g][lane] = v;
                        }
                        group_bar3(1 + g, gthreads);
#pragma unroll
                        for (int p = 0; p < PB; ++p) {
                            const float m = finmax[g][p];
                            e[u][p] = make_float2(ex2f(e[u][p].x - m), ex2f(e[u][p].y - m));
                            sm2[p] = e[u][p].x + e[u][p].y;
                        }
                        reduce_scatter<PB>(sm2, lane, false);
                        if (rwriter) redslow[g][ridx][sw] = sm2[0];
                        group_bar3(1 + g, gthreads);
                        {
                            float v = 0.f;
                            if (lane < PB) {
                                for (int w = 0; w < Sdiv; ++w) v += redslow[g][lane][w];
                            }
                            const float m = lane < PB ? finmax[g][lane] : 0.f;
                            const bool valid = lane < np;
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
                                if (sw == 0 && under) {
                                    ll += (double)(norm2 * kLn2);
                                    nlive += 1.0;
                                }
                            } else if (valid && under && sw == 0 && norm_eps_on != 0.f) {
                                ll += (double)(kLog2Eps8 * kLn2);
                            }
                            if (lane < PB) fin[warp][lane] = make_float2(inv, inv);
                        }
                    }
                    __syncwarp();
                    // ---------------- pass 2: moments of the current batch
#pragma unroll
                    for (int p = 0; p < PB; ++p) {
                        const int ip = min(bc + p, cn - 1);
                        const float4 P0 = spts[ip][0], P1 = spts[ip][1];
                        const float2 inv2 = fin[warp][p];
                        const float2 gam = fmul2(e[u][p], inv2);
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
                    __syncwarp();
                    ++kb;
                }
            }
        }
    }
#undef HGMM_PASS1
    const size_t row = (size_t)blockIdx.x * G + g;
    float* dst = partial + row * (size_t)kMom * Jp;
    if (live0) {
        const int j = sw * 32 + lane;
#pragma unroll
        for (int m = 0; m < kMom; ++m) dst[(size_t)m * Jp + j] = a[m].x;
    }
    if (live1) {
        const int j = (sw + Sdiv) * 32 + lane;
#pragma unroll
        for (int m = 0; m < kMom; ++m) dst[(size_t)m * Jp + j] = a[m].y;
    }
    if (sw == 0) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            ll += __shfl_xor_sync(0xffffffffu, ll, o);
            nlive += __shfl_xor_sync(0xffffffffu, nlive, o);
        }
        if (lane == 0) {
            rowaux[2 * row] = ll;
            rowaux[2 * row + 1] = nlive;
        }
    }
}

// ------------------------------------------------------------------------------------------
// packed-FP32 peak probe (mode 2 of hgmm_measure_fp32_peak): 8 independent FFMA2 chains, register operands
__global__ void __launch_bounds__(256) ffma2_peak_kernel(float* out, int iters, float seed, float bv, float cv) {
    float2 a0 = make_float2(seed, seed + 1), a1 = make_float2(seed + 2, seed + 3), a2 = make_float2(seed + 4, seed + 5),
           a3 = make_float2(seed + 6, seed + 7), a4 = make_float2(seed + 8, seed + 9), a5 = make_float2(seed + 10, seed + 11),
           a6 = make_float2(seed + 12, seed + 13), a7 = make_float2(seed + 14, seed + 15);
    const float2 b = make_float2(bv, bv * 1.0000002f), c = make_float2(cv, cv * 0.5f);
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            a0 = ffma2(a0, b, c); a1 = ffma2(a1, b, c); a2 = ffma2(a2, b, c); a3 = ffma2(a3, b, c);
            a4 = ffma2(a4, b, c); a5 = ffma2(a5, b, c); a6 = ffma2(a6, b, c); a7 = ffma2(a7, b, c);
        }
    }
    const float2 r = fadd2(fadd2(fadd2(a0, a1), fadd2(a2, a3)), fadd2(fadd2(a4, a5), fadd2(a6, a7)));
    if (r.x + r.y == 12345.678f) out[0] = r.x;
}

cudaError_t launch_ffma2_peak(float* out, int blocks, int iters, cudaStream_t s) {
    ffma2_peak_kernel<<<blocks, 256, 0, s>>>(out, iters, 0.5f, 1.0000001f, 1e-7f);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
template <typename K>
static int occ_blocks(K kern, int threads) {
    int occ = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, 0) != cudaSuccess || occ < 1) occ = 1;
    return occ;
}

// (Sdiv, G, W, grid, big) for the packed kernel; requires S = Jp/32 >= 2
void flat3_plan(int n, int Jp, int num_sms, int one_cta_per_sm, int* W, int* Sdiv, int* G, int* grid, int* big) {
    if (one_cta_per_sm == 3 && Jp / 32 >= 17 && (Jp / 32 + 3) / 4 <= 7) {          // two teams of ceil(S/4) warps, 4 components per lane
        const int S4 = Jp / 32;
        const int sd = (S4 + 3) / 4;
        *W = 2 * sd; *Sdiv = sd; *G = 2; *big = 2;
        *grid = num_sms;
        return;
    }
    const int S = Jp / 32;
    const int sdiv = (S + 1) / 2;                   // warps per group: each warp owns slots sw and sw + sdiv
    if (one_cta_per_sm == 7 && S >= 8 && S <= 32) {              // one component per thread, staged chunks (flat_em6.cu)
        int ctas = num_sms;
        if ((long long)ctas * 16 > n) ctas = (n + 15) / 16;
        if (ctas < 1) ctas = 1;
        *W = S; *Sdiv = S; *G = 1; *grid = ctas; *big = 6;
        return;
    }
    // shared-memory staged chunks with two CTA barriers per chunk (flat_em5.cu), kept selectable for A/B
    // shared-memory staged chunks with a barrier-free chunk pipeline (flat_em7.cu), one CTA per SM: the default from 9
    // pair columns up (J > 512; 10-iteration fit of configs[1]: 0.582 ms vs 0.637 for flat_em5, 0.700 for the per-batch
    // barrier kernel; J = 1024: 0.635 / 0.721 / 0.719; J = 640: 0.498 / 0.535 / 0.580 -- profiles/r01_flat_kernel_variants.json)
    // flat_em8.cu on top of it -- the same pipeline, the moment pass about one origin per chunk over a cell-sorted cloud (ten
    // FFMA2 per pair instead of seventeen packed operations), optionally the density pass in Cholesky form (9 instead of 12) and
    // a staggered pass order -- executes 20-30 % fewer FP32-pipe instructions and is NOT faster (configs[1]: 57.3 / 55.5 us per
    // sweep against em_flat7's 57.0; J = 1024: 63.3 / 62.5 against 61.4 -- profiles/r02_flat_sweep_ab.txt): the sweep is bound by
    // the latency of each warp's own instruction stream at four 128-register warps per scheduler, not by a pipe.
    // It stays selectable: tile_points = 9 em_flat8, 10 + Cholesky-form densities, 11 / 12 those two staggered;
    // HGMM_FLAT_SWEEP=8 / 9 make 9 / 10 the default from 9 pair columns up.
    if ((one_cta_per_sm >= 8 && one_cta_per_sm <= 12 && sdiv >= 5 && sdiv <= 16) ||
        (one_cta_per_sm == 0 && sdiv >= 9 && sdiv <= 16)) {
        static const int env_sweep = getenv("HGMM_FLAT_SWEEP") ? atoi(getenv("HGMM_FLAT_SWEEP")) : 7;
        int ctas = num_sms;
        if ((long long)ctas * 16 > n) ctas = (n + 15) / 16;
        if (ctas < 1) ctas = 1;
        *W = sdiv; *Sdiv = sdiv; *G = 1; *grid = ctas;
        if (one_cta_per_sm == 0) *big = (env_sweep == 8 || env_sweep == 9) ? env_sweep : 7;
        else *big = one_cta_per_sm == 8 ? 7 : one_cta_per_sm - 1;
        return;
    }
    if (one_cta_per_sm == 6 && sdiv >= 5 && sdiv <= 16) {
        int ctas = num_sms;
        if ((long long)ctas * 16 > n) ctas = (n + 15) / 16;
        if (ctas < 1) ctas = 1;
        *W = sdiv; *Sdiv = sdiv; *G = 1; *grid = ctas; *big = 5;
        return;
    }
    int g = 1;
    if (sdiv <= 4) g = 8 / sdiv;                    // small mixtures: several independent groups per CTA
    const int w = sdiv * g;
    // >= 8 warps per CTA: one register-rich CTA per SM (126 registers, no spills) measured faster on B200 than two
    // 72-register CTAs (71.5 vs 77.6 us on configs[1]); smaller CTAs co-reside 2-4 per SM.  one_cta_per_sm == 1 forces
    // the register-rich build, == 2 the two-CTA build (profiling switches).
    *big = (one_cta_per_sm == 1 || (one_cta_per_sm == 0 && w >= 8) || w > 13) ? 1 : 0;
    if (one_cta_per_sm == 4 && w >= 8 && w <= 13) *big = 3;          // software-pipelined build (em_flat4_kernel)
    if (one_cta_per_sm == 5 && w <= 13) *big = 4;                    // <=13 warps: register cap 152 instead of 128
    int occ = *big >= 3 ? 1 : *big ? occ_blocks(em_flat3_kernel<512, 1, 8, 1>, w * 32) : occ_blocks(em_flat3_kernel<416, 2, 4, 1>, w * 32);
    if (occ > 4) occ = 4;
    if (w >= 8 && occ > 2) occ = 2;
    int ctas = occ * num_sms;
    const long long min_pts = 2LL * kPB3 * g;
    if ((long long)ctas * min_pts > n) ctas = (int)((n + min_pts - 1) / min_pts);
    if (ctas < 1) ctas = 1;
    *W = w; *Sdiv = sdiv; *G = g; *grid = ctas;
}

cudaError_t launch_em_flat3(const float* x, const float* y, const float* z, int n, const FlatModel& m, const float* cref_blocks,
                            int W, int Sdiv, int G, int grid, int big, float* partial, double* rowaux, const int* done_flag,
                            cudaStream_t s) {
    const float eps_on = m.flavor != HGMM_FLAVOR_CPP ? 1.f : 0.f;
    const int ncref = m.Jp / 32;
    if (big >= 8 && big <= 11)      // 8: em_flat8, 9: + Cholesky-form densities, 10 / 11: the same with staggered warps
        return launch_em_flat8(x, y, z, n, m, cref_blocks, W, grid, (big & 1), big >= 10, partial, rowaux, done_flag, s);
    if (big == 7) return launch_em_flat7(x, y, z, n, m, cref_blocks, W, grid, partial, rowaux, done_flag, s);
    if (big == 6) return launch_em_flat6(x, y, z, n, m, cref_blocks, grid, partial, rowaux, done_flag, s);
    if (big == 5) return launch_em_flat5(x, y, z, n, m, cref_blocks, W, grid, partial, rowaux, done_flag, s);
    if (big == 4)
        em_flat3_kernel<416, 1, 8, 1><<<grid, W * 32, 0, s>>>(x, y, z, n, m.packed, cref_blocks, ncref, m.J, m.Jp, Sdiv, G, partial, rowaux,
                                                            done_flag, eps_on);
    else if (big == 3)
        em_flat4_kernel<416><<<grid, W * 32, 0, s>>>(x, y, z, n, m.packed, cref_blocks, ncref, m.J, m.Jp, Sdiv, G, partial, rowaux, done_flag,
                                                     eps_on);
    else if (big == 2)
        em_flat3_kernel<448, 1, 4, 2><<<grid, W * 32, 0, s>>>(x, y, z, n, m.packed, cref_blocks, ncref, m.J, m.Jp, Sdiv, G, partial, rowaux,
                                                            done_flag, eps_on);
    else if (big)
        em_flat3_kernel<512, 1, 8, 1><<<grid, W * 32, 0, s>>>(x, y, z, n, m.packed, cref_blocks, ncref, m.J, m.Jp, Sdiv, G, partial, rowaux,
                                                       done_flag, eps_on);
    else
        em_flat3_kernel<416, 2, 4, 1><<<grid, W * 32, 0, s>>>(x, y, z, n, m.packed, cref_blocks, ncref, m.J, m.Jp, Sdiv, G, partial, rowaux,
                                                       done_flag, eps_on);
    return cudaGetLastError();
}

}  // namespace hgmm
