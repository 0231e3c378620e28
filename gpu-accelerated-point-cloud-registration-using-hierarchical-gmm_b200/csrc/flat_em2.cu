// flat_em2.cu -- the default fused E+M sweep of the flat mixture ("single-evaluation" kernel), sm_100a.
//
// Same contract as flat_em.cu's em_flat_kernel (expectationStep + maximizationStep of
// src/c++/gmm_fit/gmm_kernels.cu:278-350; e_step + m_step of src/python/gmm_waymo/src/gmm_impl.py:90-116)
// but every (point, component) density is evaluated ONCE:
//
//   thread <-> component for the whole sweep (JT components per lane; their parameters and the 10
//   centred moment accumulators live in registers); points stream through in batches of PB,
//   broadcast from shared memory.  Densities are taken relative to a fixed reference
//   Cref = max_j c2_j (an upper bound of every log2 density because the quadratic form is <= 0), so
//   e = 2^(q - Cref) needs no running maximum: pass 1 computes e and the per-point sum, ONE
//   cross-warp reduction (register reduce-scatter -> shared memory -> finishing lanes) yields
//   1/sum, pass 2 accumulates gamma = e / sum.  A batch in which some point's sum falls below
//   2^-100 (a point > ~11 sigma from everything) is redone with an exact per-point maximum.
//   Points are split evenly over CTAs (any range length): no tile quantisation.  Each CTA group
//   writes its partial moments as plain coalesced fp32 rows; flat_reduce_kernel sums the rows in a
//   fixed order in fp64 (deterministic, atomic-free).
#include "common.cuh"
#include "kernels.h"

namespace hgmm {

constexpr int kChunkPts = 512;          // points staged in shared memory at a time
constexpr int kPB = 8;                  // points per batch
constexpr float kUnderflow2 = 7.888609052210118e-31f;   // 2^-100

__device__ __forceinline__ void group_bar(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// sums (or maxes) v[0..8) over the 32 lanes; on return lane L holds in v[0] the result of point
// idx(L) = 4*bit4(L) + 2*bit3(L) + bit2(L)  (9 SHFL + 9 ops instead of 40 + 40)
__device__ __forceinline__ void warp_reduce_scatter8(float* v, int lane, bool is_max) {
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

// grid.x CTAs; blockDim.x = 32 * W, W = G * Sdiv warps: G independent groups of Sdiv warps, warp sw of a
// group owns component slots sw + Sdiv*c (c < JT), lane = component inside the 32-wide slot.
template <int JT, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) em_flat2_kernel(const float* __restrict__ px, const float* __restrict__ py,
                                                              const float* __restrict__ pz, int n,
                                                              const PackedComp* __restrict__ packed,
                                                              const float* __restrict__ cref_blocks, int n_cref, int J, int Jp,
                                                              int Sdiv, int G, float* __restrict__ partial,
                                                              double* __restrict__ rowaux, const int* __restrict__ done_flag,
                                                              float norm_eps_on) {
    if (*done_flag) return;
    constexpr int PB = kPB;
    __shared__ __align__(16) float4 spts[kChunkPts];
    __shared__ __align__(16) float red[8][PB][16];        // [group][point][warp of group]  partial sums / maxima
    __shared__ __align__(16) float fin[8][3][PB];         // [group][0: inverse sum | 1: exact maximum | 2: underflowed][point]
    __shared__ int s_flag[8];                             // [group] some point of the batch underflowed

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = warp / Sdiv, sw = warp - g * Sdiv;
    const int gthreads = Sdiv * 32;
    const int S = Jp >> 5;
    const int ridx = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);   // point this lane ends up owning

    float cref = -INFINITY;
    for (int i = 0; i < n_cref; ++i) cref = fmaxf(cref, __ldg(cref_blocks + i));
    if (!(cref > kNegBig)) cref = 0.f;                    // every component dead: any reference works

    // ---- parameters of this lane's components -> registers (c2 made relative to the reference)
    float4 p0[JT], p1[JT];
    float2 p2[JT];
    bool live[JT];
#pragma unroll
    for (int c = 0; c < JT; ++c) {
        const int slot = sw + Sdiv * c;
        live[c] = slot < S;
        const int j = live[c] ? slot * 32 + lane : 0;
        const float4* s4 = reinterpret_cast<const float4*>(packed + j);
        p0[c] = __ldg(s4);
        p1[c] = __ldg(s4 + 1);
        const float4 t = __ldg(s4 + 2);
        p2[c] = make_float2(t.x, t.y);
        p0[c].w = live[c] ? p0[c].w - cref : -INFINITY;
    }
    float a[JT][kMom];
#pragma unroll
    for (int c = 0; c < JT; ++c)
#pragma unroll
        for (int k = 0; k < kMom; ++k) a[c][k] = 0.f;
    double ll = 0.0, nlive = 0.0;                         // meaningful on the finishing lanes only

    // ---- this CTA's contiguous range of points, evenly split
    const int per = (int)(((long long)n + gridDim.x - 1) / gridDim.x);
    const int lo = min(n, (int)blockIdx.x * per), hi = min(n, lo + per);

    for (int cb = lo; cb < hi; cb += kChunkPts) {
        const int cn = min(kChunkPts, hi - cb);
        __syncthreads();                                   // previous chunk fully consumed by every group
        for (int i = tid; i < cn; i += blockDim.x) spts[i] = make_float4(px[cb + i], py[cb + i], pz[cb + i], 0.f);
        __syncthreads();
        const int gper = (cn + G - 1) / G;                 // group g takes an even share of the chunk
        const int gs = min(cn, g * gper), ge = min(cn, gs + gper);
        for (int b = gs; b < ge; b += PB) {
            const int np = min(PB, ge - b);
            float e[JT][PB];
            float sm[PB];
            // ---------------- pass 1: e = 2^(q - Cref), per-point sums
#pragma unroll
            for (int p = 0; p < PB; ++p) {
                const float4 P = spts[min(b + p, cn - 1)];
                float s = 0.f;
#pragma unroll
                for (int c = 0; c < JT; ++c) {
                    float dx, dy, dz;
                    const float q = quad_q2(p0[c], p1[c], p2[c], P.x, P.y, P.z, dx, dy, dz);
                    e[c][p] = ex2f(q);
                    s += e[c][p];
                }
                sm[p] = s;
            }
            warp_reduce_scatter8(sm, lane, false);
            if ((lane & 3) == 0) red[g][ridx][sw] = sm[0];
            group_bar(1 + g, gthreads);
            if (sw == 0 && lane < PB) {                    // finishing lanes: one per point of the batch
                float v = 0.f;
                for (int w = 0; w < Sdiv; ++w) v += red[g][lane][w];
                const bool valid = lane < np;
                const bool under = valid && !(v >= kUnderflow2);
                float inv = 0.f;
                if (valid && !under) {
                    const float lse2 = cref + lg2f(v);
                    float norm2 = lse2, scale = 1.0f;
                    if (norm_eps_on != 0.f) {              // gmm_impl.py:113  log(sum exp + 1e-8)
                        const float Mx = fmaxf(lse2, kLog2Eps8);
                        norm2 = Mx + lg2f(ex2f(lse2 - Mx) + ex2f(kLog2Eps8 - Mx));
                        scale = ex2f(lse2 - norm2);
                    }
                    inv = scale / v;
                    ll += (double)(norm2 * kLn2);
                    nlive += 1.0;
                }
                fin[g][0][lane] = inv;
                fin[g][2][lane] = under ? 1.f : 0.f;
                const unsigned any = __ballot_sync((1u << PB) - 1u, under);
                if (lane == 0) s_flag[g] = any != 0u;
            }
            group_bar(1 + g, gthreads);
            if (s_flag[g]) {
                // ---------------- rare path: exact per-point maximum (some point is far from every component)
                float mx[PB];
#pragma unroll
                for (int p = 0; p < PB; ++p) {
                    const float4 P = spts[min(b + p, cn - 1)];
                    float m = kNegBig;
#pragma unroll
                    for (int c = 0; c < JT; ++c) {
                        float dx, dy, dz;
                        const float q = quad_q2(p0[c], p1[c], p2[c], P.x, P.y, P.z, dx, dy, dz);
                        e[c][p] = q;
                        m = fmaxf(m, q);
                    }
                    mx[p] = m;
                }
                warp_reduce_scatter8(mx, lane, true);
                if ((lane & 3) == 0) red[g][ridx][sw] = mx[0];
                group_bar(1 + g, gthreads);
                if (sw == 0 && lane < PB) {
                    float v = kNegBig;
                    for (int w = 0; w < Sdiv; ++w) v = fmaxf(v, red[g][lane][w]);
                    fin[g][1][lane] = v;
                }
                group_bar(1 + g, gthreads);
#pragma unroll
                for (int p = 0; p < PB; ++p) {
                    const float m = fin[g][1][p];
                    float s = 0.f;
#pragma unroll
                    for (int c = 0; c < JT; ++c) {
                        e[c][p] = ex2f(e[c][p] - m);
                        s += e[c][p];
                    }
                    sm[p] = s;
                }
                warp_reduce_scatter8(sm, lane, false);
                if ((lane & 3) == 0) red[g][ridx][sw] = sm[0];
                group_bar(1 + g, gthreads);
                if (sw == 0 && lane < PB) {
                    float v = 0.f;
                    for (int w = 0; w < Sdiv; ++w) v += red[g][lane][w];
                    const float m = fin[g][1][lane];
                    const bool valid = lane < np;
                    const bool was_under = valid && fin[g][2][lane] != 0.f;      // only those were left out of ll above
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
                        if (was_under) {
                            ll += (double)(norm2 * kLn2);
                            nlive += 1.0;
                        }
                    } else if (valid && was_under && norm_eps_on != 0.f) {
                        ll += (double)(kLog2Eps8 * kLn2);                         // log(0 + 1e-8)
                    }
                    fin[g][0][lane] = inv;
                }
                group_bar(1 + g, gthreads);
            }
            // ---------------- pass 2: moments
            const float4 i0 = *reinterpret_cast<const float4*>(&fin[g][0][0]);
            const float4 i1 = *reinterpret_cast<const float4*>(&fin[g][0][4]);
            const float inv[PB] = {i0.x, i0.y, i0.z, i0.w, i1.x, i1.y, i1.z, i1.w};
#pragma unroll
            for (int p = 0; p < PB; ++p) {
                const float4 P = spts[min(b + p, cn - 1)];
#pragma unroll
                for (int c = 0; c < JT; ++c) {
                    const float gam = e[c][p] * inv[p];
                    const float dx = P.x - p0[c].x, dy = P.y - p0[c].y, dz = P.z - p0[c].z;
                    const float gx = gam * dx, gy = gam * dy, gz = gam * dz;
                    a[c][0] += gam;
                    a[c][1] += gx;
                    a[c][2] += gy;
                    a[c][3] += gz;
                    a[c][4] = fmaf(gx, dx, a[c][4]);
                    a[c][5] = fmaf(gx, dy, a[c][5]);
                    a[c][6] = fmaf(gx, dz, a[c][6]);
                    a[c][7] = fmaf(gy, dy, a[c][7]);
                    a[c][8] = fmaf(gy, dz, a[c][8]);
                    a[c][9] = fmaf(gz, dz, a[c][9]);
                }
            }
            // the next batch's first barrier orders these fin[] reads before the finishing lanes rewrite them
        }
    }
    // ---- partial rows: partial[row][m][Jp], row = blockIdx * G + g  (coalesced over lanes)
    const size_t row = (size_t)blockIdx.x * G + g;
    float* dst = partial + row * (size_t)kMom * Jp;
#pragma unroll
    for (int c = 0; c < JT; ++c) {
        if (live[c]) {
            const int j = (sw + Sdiv * c) * 32 + lane;
#pragma unroll
            for (int k = 0; k < kMom; ++k) dst[(size_t)k * Jp + j] = a[c][k];
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

// block (slot, moment): sums one moment of 32 components over all partial rows in a fixed order (fp64);
// 8 warps split the rows.  block (0,0) also folds the per-row log-likelihood / live-point count.
__global__ void __launch_bounds__(256) flat_reduce_kernel(const float* __restrict__ partial, const double* __restrict__ rowaux,
                                                          int rows, int J, int Jp, double* __restrict__ acc,
                                                          const int* __restrict__ done_flag) {
    if (*done_flag) return;
    __shared__ double sm[8][33];
    const int tid = threadIdx.x, lane = tid & 31, rg = tid >> 5;
    const int j = blockIdx.x * 32 + lane;
    const int k = blockIdx.y;
    const float* src = partial + (size_t)k * Jp + j;
    const size_t stride = (size_t)kMom * Jp;
    double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;
    int r = rg;
    for (; r + 24 < rows; r += 32) {
        v0 += (double)src[(size_t)r * stride];
        v1 += (double)src[(size_t)(r + 8) * stride];
        v2 += (double)src[(size_t)(r + 16) * stride];
        v3 += (double)src[(size_t)(r + 24) * stride];
    }
    for (; r < rows; r += 8) v0 += (double)src[(size_t)r * stride];
    sm[rg][lane] = (v0 + v1) + (v2 + v3);
    __syncthreads();
    if (rg == 0 && j < J) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += sm[w][lane];
        acc[kAccHdr + (size_t)j * kMom + k] = t;
    }
    if (blockIdx.x == 0 && blockIdx.y == 0 && rg == 1) {
        double l = 0.0, c = 0.0;
        for (int q = lane; q < rows; q += 32) {
            l += rowaux[2 * q];
            c += rowaux[2 * q + 1];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            l += __shfl_xor_sync(0xffffffffu, l, o);
            c += __shfl_xor_sync(0xffffffffu, c, o);
        }
        if (lane == 0) {
            acc[0] = l;
            acc[1] = c;
        }
    }
}

// ------------------------------------------------------------------------------------------
template <typename K>
static int occupancy_grid(K kern, int threads, int num_sms) {
    int occ = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, 0) != cudaSuccess || occ < 1) occ = 1;
    if (occ > 2) occ = 2;
    return occ * num_sms;
}

// picks (JT, W, Sdiv, G) and the grid for a problem
void flat2_plan(int n, int Jp, int num_sms, int one_cta_per_sm, int* JT, int* W, int* Sdiv, int* G, int* grid, int* big) {
    const int S = Jp / 32;
    int jt, sdiv, g;
    if (S >= 17) {              // many slots: 2 components per lane, one group spanning the CTA (S <= 32 -> W <= 16)
        jt = 2;
        sdiv = (S + jt - 1) / jt;
        g = 1;
    } else if (S >= 5) {        // one component per lane, one group
        jt = 1;
        sdiv = S;
        g = 1;
    } else {                    // few slots: several independent groups per CTA
        jt = 1;
        sdiv = S;
        g = 8 / S;
    }
    const int w = sdiv * g;
    int ctas;
    *big = (w > 13 || one_cta_per_sm) ? 1 : 0;
    if (!*big)
        ctas = jt == 1 ? occupancy_grid(em_flat2_kernel<1, 416, 2>, w * 32, num_sms) : occupancy_grid(em_flat2_kernel<2, 416, 2>, w * 32, num_sms);
    else
        ctas = jt == 1 ? occupancy_grid(em_flat2_kernel<1, 512, 1>, w * 32, num_sms) : occupancy_grid(em_flat2_kernel<2, 512, 1>, w * 32, num_sms);
    const long long min_pts = 2LL * kPB * g;        // never so many CTAs that a group sees fewer than ~2 batches
    if ((long long)ctas * min_pts > n) ctas = (int)((n + min_pts - 1) / min_pts);
    if (ctas < 1) ctas = 1;
    *JT = jt; *Sdiv = sdiv; *G = g; *W = w; *grid = ctas;
}

cudaError_t launch_em_flat2(const float* x, const float* y, const float* z, int n, const FlatModel& m, const float* cref_blocks,
                            int JT, int W, int Sdiv, int G, int grid, int big, float* partial, double* rowaux,
                            const int* done_flag, cudaStream_t s) {
    const float eps_on = m.flavor != HGMM_FLAVOR_CPP ? 1.f : 0.f;
    const int ncref = m.Jp / 32;
    const int th = W * 32;
#define HGMM_LAUNCH2(JTV, MAXT, MINB)                                                                                       \
    em_flat2_kernel<JTV, MAXT, MINB><<<grid, th, 0, s>>>(x, y, z, n, m.packed, cref_blocks, ncref, m.J, m.Jp, Sdiv, G, partial, \
                                                         rowaux, done_flag, eps_on)
    if (!big) {
        if (JT == 1) HGMM_LAUNCH2(1, 416, 2); else HGMM_LAUNCH2(2, 416, 2);
    } else {
        if (JT == 1) HGMM_LAUNCH2(1, 512, 1); else HGMM_LAUNCH2(2, 512, 1);
    }
#undef HGMM_LAUNCH2
    return cudaGetLastError();
}

cudaError_t launch_flat_reduce(const float* partial, const double* rowaux, int rows, const FlatModel& m, double* acc,
                               const int* done_flag, cudaStream_t s) {
    flat_reduce_kernel<<<dim3(m.Jp / 32, kMom), 256, 0, s>>>(partial, rowaux, rows, m.J, m.Jp, acc, done_flag);
    return cudaGetLastError();
}

}  // namespace hgmm
