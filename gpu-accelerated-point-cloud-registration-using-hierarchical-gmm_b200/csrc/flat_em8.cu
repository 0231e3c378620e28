// flat_em8.cu -- packed-FP32 fused E+M sweep over a CELL-SORTED cloud: the moment pass is ten FFMA2 per (point, pair).
//
// Same contract as em_flat7_kernel (expectationStep + maximizationStep of src/c++/gmm_fit/gmm_kernels.cu:278-350; e_step +
// m_step of src/python/gmm_waymo/src/gmm_impl.py:90-116) and the same chunk pipeline (wait ; finish ; pass 1 of the next
// chunk ; arrive ; pass 2 of this chunk, one mbarrier per buffer, no CTA barrier in the steady state).
//
// What changes is the moment pass, 59 % of em_flat7's FP32-pipe cycles (17 of its 29 packed operations per pair: gamma,
// d = x - m again, gamma*d, four adds, six FMAs).  Here the ten sums of a CHUNK (<= 32 points) are taken about one origin o,
// a point of the chunk, instead of about each component's mean:
//     psi_p = inv_p * (1, u, v, w, uu, uv, uw, vv, vw, ww),  (u, v, w) = x_p - o          once per POINT, in the finishing step
//     s_m  += e_pj * psi_p[m]                                                              ten FMAs per (point, component)
// psi does not depend on the component, so the finishing step (lane = point, already redundant per warp) builds it and parks
// it in shared memory.  The two halves of an FFMA2 are two consecutive POINTS of one component -- (e_pj, e_p+1,j) times
// (psi_p[m], psi_p+1[m]) -- so psi is stored once (five broadcast LDS.128 per point PAIR, not per point: shared-memory
// wavefronts, not the FP32 pipe, bounded the first version of this kernel) and the densities of a pair of points and the
// lane's two components travel as one STS.128 / LDS.128.  After the chunk every lane folds the even/odd halves and moves its
// two components' sums from o to the component means, delta = m - o,
//     M1 = S1 - delta S0,   M2_ab = S2_ab - delta_a S1_b - delta_b M1_a           15 FFMA2 + 3 FADD2 per chunk
// and adds them to the same centred accumulators em_flat7 keeps: ~11 packed operations per pair instead of 17, and the
// partial rows, the reduce / exchange / finalize kernels are untouched.
//
// The move cancels (|delta| / sigma)^2 leading digits of the chunk's fp32 sums.  That is why the cloud is sorted by a 16^3
// Morton cell grid first (cloud_sort.cu, once per hgmm_set_points): the 32 points of a chunk are then neighbours, the
// components with any weight there have |delta| of a few sigma, and configs[1] after 10 iterations sits at 1.1e-5 of the
// float64 oracle where em_flat7's arithmetic sits at 5.5e-6 and the same scheme on the unsorted cloud at 6.6e-5 (float32
// emulation of both, profiles/r02_flat8_numerics.txt).  Sorting changes the summation order, nothing else; the sort is
// stable, so fits stay bit-reproducible.
//
// CHOL = true additionally evaluates the density pass in Cholesky form about the chunk's origin: -A = L L^T,
//     r = L^T u + b,  b = L^T (o - m) once per chunk,   e = 2^-(r1^2 + r2^2 + r3^2 + (Cref - c2))       9 FFMA2 per pair
// instead of the 12 of d^T A d (the points are staged relative to their chunk's origin; Cref - c2 >= 0 is the fourth square).
#include <stdlib.h>
#include "common.cuh"
#include "kernels.h"
#include "packed.cuh"

namespace hgmm {

constexpr int kRed8 = 32;               // points per chunk are at most 32 (lane = point in the finishing step)

// -DHGMM_FLAT8_PROF (make prof -> build/libhgmm_prof.so, never the shipped library): every warp accumulates clock64() per phase
// -- 0 barrier wait, 1 finish + psi, 2 density pass, 3 moment pass + re-centring, 4 prologue, 5 epilogue -- read back by
// profiles/probe_flat8_phases.py through hgmm_debug_flat8_prof
#ifdef HGMM_FLAT8_PROF
__device__ unsigned long long g_f8prof[148 * 16 * 8];
#define F8T(i)                          \
    {                                   \
        const long long t_ = clock64(); \
        f8p[i] += t_ - f8t;             \
        f8t = t_;                       \
    }
#else
#define F8T(i)
#endif
constexpr int kMaxChunks8 = 64;         // chunks per staging block (SB <= 512, CH >= 8)
constexpr unsigned kLightSleepNs = 0;   // sleep of the left-over column's warps before they poll a chunk barrier (0 = poll at once)

// dynamic shared memory layout (CH = chunk points, SB = staged points, C = 32 * P):
//   u64 bars[2] | float red[2][16][32] | float mval[32] | float4 orig[64] | float4 psi[2][16][5] | float4 spts[SB + 8][2]
//   | float4 ebuf[2][CH / 2][C]
// the rare path's column maxima / shifted sums live in the OTHER buffer's half of red (free while every warp is between
// finish(c) and pass 1 of chunk c + 1)
__host__ __device__ inline size_t flat8_fixed_bytes() {
    return 16 + 2 * 16 * kRed8 * 4 + kRed8 * 4 + kMaxChunks8 * 16 + 2 * (kRed8 / 2) * 5 * 16;
}
__host__ __device__ inline size_t flat8_smem_bytes(int CH, int SB, int C) {
    return flat8_fixed_bytes() + (size_t)(SB + 8) * 32 + (size_t)2 * (CH / 2) * C * 16;
}

__device__ __forceinline__ void mbar_arrive8(uint64_t* bar) {
    asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait8(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT8_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra WAIT8_DONE;\n\t"
        "bra WAIT8_LOOP;\n\t"
        "WAIT8_DONE:\n\t"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity), "r"(20000u)
        : "memory");
}

// two components side by side.  CHOL = false: PairParams of packed.cuh (d^T A d).  CHOL = true: the upper factor of -A.
struct CholParams {
    float2 nmx, nmy, nmz, k4;                  // -mean, Cref - c2 (>= 0; +inf for a dead component)
    float2 l11, l21, l31, l22, l32, l33;       // r1 = l11 u + l21 v + l31 w + b1, r2 = l22 v + l32 w + b2, r3 = l33 w + b3
};

// 1/sqrt(t) for t > 0: MUFU.RSQ seed (2^-22) and one float64 Newton step (2^-43) -- no DSQRT / DDIV sequences in the prologue
__device__ __forceinline__ double rsqrt_nr(double t) {
    const double y = (double)rsqrtf((float)t);
    return y * (1.5 - 0.5 * t * y * y);
}
// -A (packed: diagonal, doubled off-diagonals) = L L^T in float64, a non-positive pivot zeroes its row (flat direction)
__device__ __forceinline__ void chol_of_minus_a(float axx, float ayy, float azz, float axy, float axz, float ayz, float* l) {
    const double mxx = -(double)axx, myy = -(double)ayy, mzz = -(double)azz;
    const double mxy = -0.5 * (double)axy, mxz = -0.5 * (double)axz, myz = -0.5 * (double)ayz;
    const double i11 = mxx > 1e-300 ? rsqrt_nr(mxx) : 0.0, l11 = mxx * i11;
    const double l21 = mxy * i11, l31 = mxz * i11;
    double t = myy - l21 * l21;
    const double i22 = t > 1e-300 ? rsqrt_nr(t) : 0.0, l22 = t * i22;
    const double l32 = (myz - l31 * l21) * i22;
    t = mzz - l31 * l31 - l32 * l32;
    const double l33 = t > 1e-300 ? t * rsqrt_nr(t) : 0.0;
    l[0] = (float)l11; l[1] = (float)l21; l[2] = (float)l31; l[3] = (float)l22; l[4] = (float)l32; l[5] = (float)l33;
}

template <int P, bool CHOL>
__global__ void __launch_bounds__(512, 1) em_flat8_kernel(const float* __restrict__ px, const float* __restrict__ py,
                                                          const float* __restrict__ pz, int n,
                                                          const PackedComp* __restrict__ packed,
                                                          const float* __restrict__ cref_blocks, int n_cref, int Jp, int CH,
                                                          int SB, float* __restrict__ partial, double* __restrict__ rowaux,
                                                          const int* __restrict__ done_flag, float norm_eps_on, int stagger,
                                                          unsigned light_sleep_ns) {
    constexpr int PB = 8;
    constexpr int C = P * 32;                                              // e columns (one float4 per lane and point pair)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int T = blockDim.x, W = T >> 5;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);                // [2]
    float* red = reinterpret_cast<float*>(bars + 2);                       // [2][16][32]  partial sums, column-major
    float* mval = red + 2 * 16 * kRed8;                                    // [32]         rare path: exact maxima
    float4* orig = reinterpret_cast<float4*>(mval + kRed8);                // [64]         origin of every chunk of the block
    float4* psi = orig + kMaxChunks8;                                      // [2][16][5]   (psi_p[m], psi_p+1[m]) pairs
    float4* spts = psi + 2 * (kRed8 / 2) * 5;                              // [SB + 8][2]  (x,x,y,y) (z,z,0,0)
    float4* ebuf = spts + (size_t)(SB + 8) * 2;                            // [2][CH/2][C] (eA_p, eA_p+1, eB_p, eB_p+1)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#ifdef HGMM_FLAT8_PROF
    long long f8p[6] = {0, 0, 0, 0, 0, 0}, f8t = clock64();
#endif
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
    // points of staging block sb0 -> (x,x,y,y) (z,z,0,0), plus one batch of padding (copies of the last point); the origin of
    // chunk c is its middle point; CHOL stages the points relative to it
    auto stage = [&](int sb0, int cn) {
        for (int i = tid; i < cn + PB; i += T) {
            const int ii = min(i, cn - 1);
            const int src = sb0 + ii;
            float x = px[src], y = py[src], z = pz[src];
            if (CHOL) {
                const int c0 = (ii / CH) * CH;
                const int osrc = sb0 + c0 + (min(CH, cn - c0) >> 1);
                x -= px[osrc]; y -= py[osrc]; z -= pz[osrc];
            }
            spts[2 * i] = make_float4(x, x, y, y);
            spts[2 * i + 1] = make_float4(z, z, 0.f, 0.f);
        }
        for (int c = tid; c * CH < cn; c += T) {
            const int osrc = sb0 + c * CH + (min(CH, cn - c * CH) >> 1);
            orig[c] = make_float4(px[osrc], py[osrc], pz[osrc], 0.f);
        }
    };
    // the cloud never changes between EM iterations: the first block is staged BEFORE the programmatic-dependent-launch
    // wait and overlaps the tail of the previous iteration's reduce + finalize kernel
    if (hi > lo) stage(lo, min(SB, hi - lo));
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (__ldcg(done_flag)) return;

    // ---- which pair column, and which share of every chunk's batches, this warp sweeps (as em_flat7_kernel)
    const int full = P == W ? P : (P & ~3);
    int col = warp, nsplit = 1, sidx = 0;
    if (warp >= full) {
        const int r = P - full;
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
    CholParams kc;
    const bool live0 = col < S, live1 = col + P < S;
    {
        const float4* a4 = reinterpret_cast<const float4*>(packed + (live0 ? col * 32 + lane : 0));
        const float4* b4 = reinterpret_cast<const float4*>(packed + (live1 ? (col + P) * 32 + lane : 0));
        const float4 a0 = __ldcg(a4), a1 = __ldcg(a4 + 1), a2 = __ldcg(a4 + 2);
        const float4 b0 = __ldcg(b4), b1 = __ldcg(b4 + 1), b2 = __ldcg(b4 + 2);
        if (CHOL) {
            float la[6], lb[6];
            chol_of_minus_a(a1.x, a1.y, a1.z, a1.w, a2.x, a2.y, la);
            chol_of_minus_a(b1.x, b1.y, b1.z, b1.w, b2.x, b2.y, lb);
            kc.nmx = make_float2(-a0.x, -b0.x);
            kc.nmy = make_float2(-a0.y, -b0.y);
            kc.nmz = make_float2(-a0.z, -b0.z);
            kc.k4 = make_float2(live0 && a0.w > -INFINITY ? fmaxf(cref - a0.w, 0.f) : INFINITY,
                                live1 && b0.w > -INFINITY ? fmaxf(cref - b0.w, 0.f) : INFINITY);
            kc.l11 = make_float2(la[0], lb[0]); kc.l21 = make_float2(la[1], lb[1]); kc.l31 = make_float2(la[2], lb[2]);
            kc.l22 = make_float2(la[3], lb[3]); kc.l32 = make_float2(la[4], lb[4]); kc.l33 = make_float2(la[5], lb[5]);
        } else {
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
    }
    const float2 nmx = CHOL ? kc.nmx : k.nmx, nmy = CHOL ? kc.nmy : k.nmy, nmz = CHOL ? kc.nmz : k.nmz;
    float2 a[kMom];
#pragma unroll
    for (int m = 0; m < kMom; ++m) a[m] = make_float2(0.f, 0.f);
    double ll = 0.0, nlive = 0.0;                         // warp 0, lane = point of the chunk

    float4* const ecol = ebuf + col * 32 + lane;          // pair pp of buffer b's column: ecol[(b * (CH/2) + pp) * C]
    float* const redcol = red + col * kRed8;              // this column's partial sums: redcol[b * 16 * 32 + p]
    const int CH2 = CH >> 1;
    unsigned g = 0;                                       // chunks done so far: buffer g & 1, mbarrier parity (g >> 1) & 1
    const bool late = stagger != 0 && ((warp >> 2) & 1) != 0;

    // log2 density (minus Cref) of the lane's pair at a staged point
    float2 bq1 = make_float2(0.f, 0.f), bq2 = bq1, bq3 = bq1;          // CHOL: b of the chunk being evaluated
    auto chunk_bias = [&](int cidx) {
        if (CHOL) {
            const float4 Oc = orig[cidx];
            const float2 ndx = fadd2(nmx, make_float2(Oc.x, Oc.x)), ndy = fadd2(nmy, make_float2(Oc.y, Oc.y)),
                         ndz = fadd2(nmz, make_float2(Oc.z, Oc.z));
            bq1 = ffma2(kc.l11, ndx, ffma2(kc.l21, ndy, fmul2(kc.l31, ndz)));
            bq2 = ffma2(kc.l22, ndy, fmul2(kc.l32, ndz));
            bq3 = fmul2(kc.l33, ndz);
        }
    };
    auto logdens = [&](const float4& P0, const float4& P1) -> float2 {
        if (CHOL) {
            const float2 U = make_float2(P0.x, P0.y), V = make_float2(P0.z, P0.w), Wz = make_float2(P1.x, P1.y);
            const float2 r1 = ffma2(kc.l11, U, ffma2(kc.l21, V, ffma2(kc.l31, Wz, bq1)));
            const float2 r2 = ffma2(kc.l22, V, ffma2(kc.l32, Wz, bq2));
            const float2 r3 = ffma2(kc.l33, Wz, bq3);
            const float2 s = ffma2(r3, r3, ffma2(r2, r2, ffma2(r1, r1, kc.k4)));
            return make_float2(-s.x, -s.y);
        } else {
            float2 dx, dy, dz;
            return quad2(k, make_float2(P0.x, P0.y), make_float2(P0.z, P0.w), make_float2(P1.x, P1.y), dx, dy, dz);
        }
    };

    // pass 1 of the chunk starting at staged point c0 (ch valid points) into buffer b
    auto pass1 = [&](int c0, int ch, int b) {
        const int nb = (ch + PB - 1) / PB;
        const int b_lo = (sidx * nb / nsplit) * PB, b_hi = ((sidx + 1) * nb / nsplit) * PB;
        chunk_bias(c0 / CH);
        for (int bb = b_lo; bb < b_hi; bb += PB) {
            const float4* sp = spts + 2 * (c0 + bb);
            float4* eb = ecol + (size_t)(b * CH2 + (bb >> 1)) * C;
            float sm[PB];
            float2 q[PB];
            // all shared-memory loads of the batch first (an LDS cannot be hoisted over an STS the compiler cannot prove disjoint)
#pragma unroll
            for (int p = 0; p < PB; ++p) q[p] = logdens(sp[2 * p], sp[2 * p + 1]);
#pragma unroll
            for (int p = 0; p < PB; p += 2) {
                const float2 e0 = make_float2(ex2f(q[p].x), ex2f(q[p].y));
                const float2 e1 = make_float2(ex2f(q[p + 1].x), ex2f(q[p + 1].y));
                eb[(p >> 1) * C] = make_float4(e0.x, e1.x, e0.y, e1.y);
                sm[p] = e0.x + e0.y;
                sm[p + 1] = e1.x + e1.y;
            }
            reduce_scatter<PB>(sm, lane, false);
            if (rwriter) redcol[b * 16 * kRed8 + bb + ridx] = sm[0];
        }
    };

    for (int sb0 = lo; sb0 < hi; sb0 += SB) {
        const int cn = min(SB, hi - sb0);
        if (sb0 != lo) {
            __syncthreads();                              // the previous block's points are no longer read
            stage(sb0, cn);
        }
        __syncthreads();                                  // points staged (and, the first time, barriers initialised)
        F8T(4)
        const int K = (cn + CH - 1) / CH;
        pass1(0, min(CH, cn), g & 1);
        __syncwarp();
        if (lane == 0) mbar_arrive8(&bars[g & 1]);
        F8T(2)
        for (int c = 0; c < K; ++c, ++g) {
            const int b = g & 1;
            const int c0 = c * CH;
            const int ch = min(CH, cn - c0);
            // the warps that share the left-over column idle ~3/4 of a chunk: they sleep through most of it instead of polling
            // (a poll is four instructions every ~32 cycles -- an eighth of the scheduler's issue slots)
            if (nsplit > 1 && light_sleep_ns != 0u) __nanosleep(light_sleep_ns);
            mbar_wait8(&bars[b], (g >> 1) & 1);
            F8T(0)
            // ---------------- finish (every warp, lane = point): fold the P column sums in a fixed order
            const bool valid = lane < ch;
            float v = 0.f;
            {
                const float* r = red + b * 16 * kRed8 + lane;
#pragma unroll
                for (int kk = 0; kk < P; ++kk) v += r[kk * kRed8];
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
                float* rmax = red + (b ^ 1) * 16 * kRed8;   // the other buffer's sums were consumed by finish(c - 1) of every warp
                const int nb = (ch + PB - 1) / PB;
                const int b_lo = (sidx * nb / nsplit) * PB, b_hi = ((sidx + 1) * nb / nsplit) * PB;
                chunk_bias(c);
                __syncthreads();
                for (int bb = b_lo; bb < b_hi; bb += PB) {
                    const float4* sp = spts + 2 * (c0 + bb);
                    float mx[PB];
#pragma unroll
                    for (int p = 0; p < PB; ++p) {
                        const float2 q = logdens(sp[2 * p], sp[2 * p + 1]);
                        mx[p] = fmaxf(fmaxf(q.x, q.y), kNegBig);
                    }
                    reduce_scatter<PB>(mx, lane, true);
                    if (rwriter) rmax[col * kRed8 + bb + ridx] = mx[0];
                }
                __syncthreads();
                float m = kNegBig;
#pragma unroll
                for (int kk = 0; kk < P; ++kk) m = fmaxf(m, rmax[kk * kRed8 + lane]);
                mval[lane] = m;                           // every warp writes the same values
                __syncthreads();                          // the maxima are read by everyone before the sums overwrite them
                for (int bb = b_lo; bb < b_hi; bb += PB) {
                    const float4* sp = spts + 2 * (c0 + bb);
                    float4* eb = ecol + (size_t)(b * CH2 + (bb >> 1)) * C;
                    float sm[PB];
                    float2 q[PB];
#pragma unroll
                    for (int p = 0; p < PB; ++p) {
                        q[p] = logdens(sp[2 * p], sp[2 * p + 1]);
                        const float mp = mval[bb + p];
                        q[p].x -= mp;
                        q[p].y -= mp;
                    }
#pragma unroll
                    for (int p = 0; p < PB; p += 2) {
                        const float2 e0 = make_float2(ex2f(q[p].x), ex2f(q[p].y));
                        const float2 e1 = make_float2(ex2f(q[p + 1].x), ex2f(q[p + 1].y));
                        eb[(p >> 1) * C] = make_float4(e0.x, e1.x, e0.y, e1.y);
                        sm[p] = e0.x + e0.y;
                        sm[p + 1] = e1.x + e1.y;
                    }
                    reduce_scatter<PB>(sm, lane, false);
                    if (rwriter) rmax[col * kRed8 + bb + ridx] = sm[0];
                }
                __syncthreads();
                float vs = 0.f;
#pragma unroll
                for (int kk = 0; kk < P; ++kk) vs += rmax[kk * kRed8 + lane];
                __syncthreads();                          // ... and the sums before pass 1 of the next chunk reuses the buffer
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
            // ---------------- psi of this lane's point about the chunk's origin (every warp writes the same values):
            //                  psi[b][p >> 1][h] = (psi_p[2h], psi_p+1[2h], psi_p[2h+1], psi_p+1[2h+1])
            const float4 O = orig[c];
            {
                const int sp_i = valid ? c0 + lane : c0;                                    // lanes past the chunk: iv = 0
                const float4 P0 = spts[2 * sp_i], P1 = spts[2 * sp_i + 1];
                const float u = CHOL ? P0.x : P0.x - O.x, w2 = CHOL ? P0.z : P0.z - O.y, w3 = CHOL ? P1.x : P1.x - O.z;
                const float iu = iv * u, iw2 = iv * w2, iw3 = iv * w3;
                float* dst = reinterpret_cast<float*>(psi + (b * (kRed8 / 2) + (lane >> 1)) * 5) + (lane & 1);
                dst[0] = iv;        dst[2] = iu;           // float4 0: m = 0, 1
                dst[4] = iw2;       dst[6] = iw3;          // float4 1: m = 2, 3
                dst[8] = iu * u;    dst[10] = iu * w2;     // float4 2: xx, xy
                dst[12] = iu * w3;  dst[14] = iw2 * w2;    // float4 3: xz, yy
                dst[16] = iw2 * w3; dst[18] = iw3 * w3;    // float4 4: yz, zz
            }
            __syncwarp();
            F8T(1)
            // ---------------- pass 2: ten FFMA2 per component and point PAIR of this warp's share of chunk c
            auto pass2 = [&]() {
                const int nb = (ch + PB - 1) / PB;
                const int pp_lo = (sidx * nb / nsplit) * (PB / 2);
                const int pp_hi = (min(((sidx + 1) * nb / nsplit) * PB, ch) + 1) >> 1;      // an odd tail pairs with psi = 0
                const float4* eb = ecol + (size_t)b * CH2 * C;
                const float4* ps = psi + b * (kRed8 / 2) * 5;
                float2 sA[kMom], sB[kMom];
#pragma unroll
                for (int m = 0; m < kMom; ++m) sA[m] = sB[m] = make_float2(0.f, 0.f);
#pragma unroll 2
                for (int pp = pp_lo; pp < pp_hi; ++pp) {
                    const float4 e = eb[(size_t)pp * C];
                    const float2 eA = make_float2(e.x, e.y), eB = make_float2(e.z, e.w);
#pragma unroll
                    for (int h = 0; h < 5; ++h) {
                        const float4 s = ps[5 * pp + h];
                        sA[2 * h] = ffma2(eA, make_float2(s.x, s.y), sA[2 * h]);
                        sB[2 * h] = ffma2(eB, make_float2(s.x, s.y), sB[2 * h]);
                        sA[2 * h + 1] = ffma2(eA, make_float2(s.z, s.w), sA[2 * h + 1]);
                        sB[2 * h + 1] = ffma2(eB, make_float2(s.z, s.w), sB[2 * h + 1]);
                    }
                }
                float2 sc[kMom];                           // even + odd points, the lane's two components side by side again
#pragma unroll
                for (int m = 0; m < kMom; ++m) sc[m] = make_float2(sA[m].x + sA[m].y, sB[m].x + sB[m].y);
                // from the chunk's origin to the pair's means: nd = o - m = -delta
                const float2 ndx = fadd2(nmx, make_float2(O.x, O.x));
                const float2 ndy = fadd2(nmy, make_float2(O.y, O.y));
                const float2 ndz = fadd2(nmz, make_float2(O.z, O.z));
                const float2 m1x = ffma2(ndx, sc[0], sc[1]);
                const float2 m1y = ffma2(ndy, sc[0], sc[2]);
                const float2 m1z = ffma2(ndz, sc[0], sc[3]);
                a[0] = fadd2(a[0], sc[0]);
                a[1] = fadd2(a[1], m1x);
                a[2] = fadd2(a[2], m1y);
                a[3] = fadd2(a[3], m1z);
                a[4] = fadd2(a[4], ffma2(ndx, m1x, ffma2(ndx, sc[1], sc[4])));      // xx
                a[5] = fadd2(a[5], ffma2(ndy, m1x, ffma2(ndx, sc[2], sc[5])));      // xy
                a[6] = fadd2(a[6], ffma2(ndz, m1x, ffma2(ndx, sc[3], sc[6])));      // xz
                a[7] = fadd2(a[7], ffma2(ndy, m1y, ffma2(ndy, sc[2], sc[7])));      // yy
                a[8] = fadd2(a[8], ffma2(ndz, m1y, ffma2(ndy, sc[3], sc[8])));      // yz
                a[9] = fadd2(a[9], ffma2(ndz, m1z, ffma2(ndz, sc[3], sc[9])));      // zz
            };
            // ---------------- pass 1 of the next chunk is published before this chunk's moment pass -- except, with `stagger`,
            //                  in every other warp of a scheduler (warps w and w + 4 share one): those take the moment pass first,
            //                  so the FFMA2-dense pass of one warp overlaps the MUFU / shuffle tails of its neighbours' pass 1
            if (!late) {
                if (c + 1 < K) {
                    pass1(c0 + CH, min(CH, cn - c0 - CH), b ^ 1);
                    __syncwarp();
                    if (lane == 0) mbar_arrive8(&bars[b ^ 1]);
                }
                F8T(2)
                pass2();
                F8T(3)
            } else {
                pass2();
                F8T(3)
                if (c + 1 < K) {
                    pass1(c0 + CH, min(CH, cn - c0 - CH), b ^ 1);
                    __syncwarp();
                    if (lane == 0) mbar_arrive8(&bars[b ^ 1]);
                }
                F8T(2)
            }
        }
    }
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");      // the reduce + finalize kernel may start launching
    // ---- warps sharing a column fold their partial moments in a fixed order (the e buffers are free now)
    __syncthreads();
    float2* const scratch0 = reinterpret_cast<float2*>(ebuf);
    if (nsplit > 1 && sidx > 0) {
        float2* scratch = scratch0 + ((size_t)(col - full) * 3 + (sidx - 1)) * 32 * kMom;
#pragma unroll
        for (int m = 0; m < kMom; ++m) scratch[m * 32 + lane] = a[m];
    }
    __syncthreads();
    if (nsplit > 1 && sidx == 0) {
        for (int s2 = 1; s2 < nsplit; ++s2) {
            const float2* scratch = scratch0 + ((size_t)(col - full) * 3 + (s2 - 1)) * 32 * kMom;
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
#ifdef HGMM_FLAT8_PROF
    F8T(5)
    if (lane == 0 && blockIdx.x < 148 && warp < 16)
        for (int i = 0; i < 6; ++i) g_f8prof[(blockIdx.x * 16 + warp) * 8 + i] = (unsigned long long)f8p[i];
#endif
}

int flat5_warps(int P);

// chunk length (multiple of 8, at most 32) and staging block (multiple of 8, at most 512) that fit the shared memory
static void flat8_shape(int C, int smem_optin, int* CH, int* SB) {
    int ch = 32;
    while (ch > 8 && flat8_smem_bytes(ch, 64, C) > (size_t)smem_optin) ch -= 8;
    long long room = (long long)smem_optin - (long long)flat8_smem_bytes(ch, 0, C);
    int sb = (int)(room / 32) / 8 * 8;
    if (sb > 512) sb = 512;
    if (sb < 8) sb = 8;
    *CH = ch;
    *SB = sb;
}

template <int P, bool CHOL>
static cudaError_t launch8(const float* x, const float* y, const float* z, int n, const FlatModel& m, const float* cref_blocks,
                           int grid, float* partial, double* rowaux, const int* done_flag, int smem_optin, int stagger, cudaStream_t s) {
    static DeviceOnce once;      // one per instantiation
    if (once.first()) {
        cudaError_t e = cudaFuncSetAttribute(em_flat8_kernel<P, CHOL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin);
        if (e != cudaSuccess) return e;
    }
    int CH, SB;
    flat8_shape(P * 32, smem_optin, &CH, &SB);
    const int W = flat5_warps(P);
    const float eps_on = m.flavor != HGMM_FLAVOR_CPP ? 1.f : 0.f;
    const char* sl = getenv("HGMM_FLAT8_SLEEP_NS");      // A/B switch (read per launch); default: see launch_em_flat8
    const unsigned light_sleep_ns = sl ? (unsigned)atoi(sl) : kLightSleepNs;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(W * 32);
    cfg.dynamicSmemBytes = flat8_smem_bytes(CH, SB, P * 32);
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, em_flat8_kernel<P, CHOL>, x, y, z, n, m.packed, cref_blocks, m.Jp / 32, m.Jp, CH, SB, partial,
                              rowaux, done_flag, eps_on, stagger, light_sleep_ns);
}

cudaError_t launch_em_flat8(const float* x, const float* y, const float* z, int n, const FlatModel& m, const float* cref_blocks,
                            int P, int grid, int chol, int stagger, float* partial, double* rowaux, const int* done_flag,
                            cudaStream_t s) {
    const int smem_optin = device_smem_optin();
#define HGMM_F8(PP)                                                                                                            \
    case PP:                                                                                                                   \
        return chol ? launch8<PP, true>(x, y, z, n, m, cref_blocks, grid, partial, rowaux, done_flag, smem_optin, stagger, s)  \
                    : launch8<PP, false>(x, y, z, n, m, cref_blocks, grid, partial, rowaux, done_flag, smem_optin, stagger, s);
    switch (P) {
        HGMM_F8(5) HGMM_F8(6) HGMM_F8(7) HGMM_F8(8) HGMM_F8(9) HGMM_F8(10) HGMM_F8(11) HGMM_F8(12)
        HGMM_F8(13) HGMM_F8(14) HGMM_F8(15) HGMM_F8(16)
        default: return cudaErrorInvalidValue;
    }
#undef HGMM_F8
}

}  // namespace hgmm

#ifdef HGMM_FLAT8_PROF
extern "C" int hgmm_debug_flat8_prof(unsigned long long* out) {      // 148 x 16 x 8 counters of the last launch on the current device
    return (int)cudaMemcpyFromSymbol(out, hgmm::g_f8prof, sizeof(unsigned long long) * 148 * 16 * 8);
}
#endif
