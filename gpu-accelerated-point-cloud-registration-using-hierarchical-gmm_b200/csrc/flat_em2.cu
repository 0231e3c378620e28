// flat_em2.cu -- second-generation fused E+M sweep of the flat mixture (the default path), sm_100a.
//
// Same contract as flat_em.cu's em_flat_kernel (expectationStep + maximizationStep of
// src/c++/gmm_fit/gmm_kernels.cu:278-350; e_step + m_step of src/python/gmm_waymo/src/gmm_impl.py:90-116)
// but every (point, component) density is evaluated ONCE:
//
//   thread <-> component for the whole sweep (JT components per lane, parameters and the 10 centred
//   moment accumulators live in registers); points stream through in batches of PB, broadcast from
//   shared memory.  For a batch: pass 1 computes q (log2 density) and the per-point maximum, pass 2
//   turns q into e = 2^(q-max) and the per-point sum, pass 3 accumulates gamma = e / sum.  The two
//   per-point reductions go warp-shuffle -> shared memory -> one finishing warp; no N x J matrix and
//   no density recomputation.  Points are split evenly over CTAs (any range length), so there is no
//   tile quantisation.  Each CTA group writes its partial moments as plain coalesced fp32 rows; a
//   second kernel sums the rows in a fixed order in fp64 (deterministic, atomic-free).
#include "common.cuh"
#include "kernels.h"

namespace hgmm {

constexpr int kChunkPts = 512;          // points staged in shared memory at a time

__device__ __forceinline__ void group_bar(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// grid.x CTAs; blockDim.x = 32 * W, W = G * Sdiv warps: G independent groups of Sdiv warps, warp sw of a
// group owns component slots sw + Sdiv*c (c < JT), lane = component inside the 32-wide slot.
template <int JT, int PB, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) em_flat2_kernel(const float* __restrict__ px, const float* __restrict__ py,
                                                       const float* __restrict__ pz, int n,
                                                       const PackedComp* __restrict__ packed, int J, int Jp, int Sdiv, int G,
                                                       float* __restrict__ partial, double* __restrict__ rowaux,
                                                       const int* __restrict__ done_flag, float norm_eps_on) {
    if (*done_flag) return;
    __shared__ __align__(16) float4 spts[kChunkPts];
    __shared__ __align__(16) float red[16][PB][16];       // [group][point][warp of group]  (partials)
    __shared__ __align__(16) float fin[16][PB];            // [group][point]                (finished value)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = warp / Sdiv, sw = warp - g * Sdiv;
    const int gthreads = Sdiv * 32;
    const int S = Jp >> 5;

    // ---- parameters of this lane's components -> registers
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
        if (!live[c]) p0[c].w = -INFINITY;
    }
    float a[JT][kMom];
#pragma unroll
    for (int c = 0; c < JT; ++c)
#pragma unroll
        for (int k = 0; k < kMom; ++k) a[c][k] = 0.f;
    double ll = 0.0, nlive = 0.0;

    // ---- this CTA's contiguous range of points, evenly split
    const int per = (int)(((long long)n + gridDim.x - 1) / gridDim.x);
    const int lo = min(n, (int)blockIdx.x * per), hi = min(n, lo + per);

    for (int cb = lo; cb < hi; cb += kChunkPts) {
        const int cn = min(kChunkPts, hi - cb);
        __syncthreads();                                   // previous chunk fully consumed by every group
        for (int i = tid; i < cn; i += blockDim.x) spts[i] = make_float4(px[cb + i], py[cb + i], pz[cb + i], 0.f);
        __syncthreads();
        // group g takes an even share of the chunk
        const int gper = (cn + G - 1) / G;
        const int gs = min(cn, g * gper), ge = min(cn, gs + gper);
        for (int b = gs; b < ge; b += PB) {
            const int np = min(PB, ge - b);
            float q[JT][PB];
            float mx[PB];
            // ---------------- pass 1: q and per-point max
#pragma unroll
            for (int p = 0; p < PB; ++p) {
                const float4 P = spts[min(b + p, cn - 1)];
                float m = kNegBig;
#pragma unroll
                for (int c = 0; c < JT; ++c) {
                    float dx, dy, dz;
                    q[c][p] = quad_q2(p0[c], p1[c], p2[c], P.x, P.y, P.z, dx, dy, dz);
                    m = fmaxf(m, q[c][p]);
                }
                mx[p] = m;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                for (int p = 0; p < PB; ++p) mx[p] = fmaxf(mx[p], __shfl_xor_sync(0xffffffffu, mx[p], o));
            if (Sdiv > 1) {
                if (lane < PB) {
                    float v = mx[0];
#pragma unroll
                    for (int p = 1; p < PB; ++p) v = (lane == p) ? mx[p] : v;
                    red[g][lane][sw] = v;
                }
                group_bar(1 + g, gthreads);
                if (sw == 0 && lane < PB) {
                    float v = kNegBig;
                    for (int w = 0; w < Sdiv; ++w) v = fmaxf(v, red[g][lane][w]);
                    fin[g][lane] = v;
                }
                group_bar(1 + g, gthreads);
#pragma unroll
                for (int p = 0; p < PB; ++p) mx[p] = fin[g][p];
            }
            // ---------------- pass 2: e = 2^(q - max), per-point sum
            float sm[PB];
#pragma unroll
            for (int p = 0; p < PB; ++p) {
                float s = 0.f;
#pragma unroll
                for (int c = 0; c < JT; ++c) {
                    const float e = ex2f(q[c][p] - mx[p]);
                    q[c][p] = e;
                    s += e;
                }
                sm[p] = s;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                for (int p = 0; p < PB; ++p) sm[p] += __shfl_xor_sync(0xffffffffu, sm[p], o);
            if (Sdiv > 1) {
                if (lane < PB) {
                    float v = sm[0];
#pragma unroll
                    for (int p = 1; p < PB; ++p) v = (lane == p) ? sm[p] : v;
                    red[g][lane][sw] = v;
                }
                group_bar(1 + g, gthreads);
                if (sw == 0 && lane < PB) {
                    float v = 0.f;
                    for (int w = 0; w < Sdiv; ++w) v += red[g][lane][w];
                    fin[g][lane] = v;
                }
                group_bar(1 + g, gthreads);
#pragma unroll
                for (int p = 0; p < PB; ++p) sm[p] = fin[g][p];
            }
            // ---------------- normaliser, log-likelihood
            float inv[PB];
#pragma unroll
            for (int p = 0; p < PB; ++p) {
                const float lse2 = mx[p] + lg2f(sm[p]);                   // -inf / NaN-free: mx finite, sm >= 0
                float norm2 = lse2;
                float scale = 1.0f;
                if (norm_eps_on != 0.f) {                                 // gmm_impl.py:113  log(sum exp + 1e-8)
                    const float Mx = fmaxf(lse2, kLog2Eps8);
                    norm2 = Mx + lg2f(ex2f(lse2 - Mx) + ex2f(kLog2Eps8 - Mx));
                    scale = ex2f(lse2 - norm2);                           // gamma = 2^(q - norm2) = (e / sum) * scale
                }
                const bool ok = (p < np) && (sm[p] > 0.f) && (mx[p] > kNegBig);
                inv[p] = ok ? scale / sm[p] : 0.f;
                if (sw == 0 && lane == 0 && p < np) {
                    const bool fin_ok = norm2 > kNegBig;
                    ll += fin_ok ? (double)(norm2 * kLn2) : 0.0;
                    nlive += ok ? 1.0 : 0.0;
                }
            }
            // ---------------- pass 3: moments
#pragma unroll
            for (int p = 0; p < PB; ++p) {
                const float4 P = spts[min(b + p, cn - 1)];
#pragma unroll
                for (int c = 0; c < JT; ++c) {
                    const float gam = q[c][p] * inv[p];
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
    if (sw == 0 && lane == 0) {
        rowaux[2 * row] = ll;
        rowaux[2 * row + 1] = nlive;
    }
}

// sum the partial rows in a fixed order (fp64) -> acc[kAccHdr + j*10 + m]; block 0 also folds ll / live count
__global__ void __launch_bounds__(256) flat_reduce_kernel(const float* __restrict__ partial, const double* __restrict__ rowaux,
                                                          int rows, int J, int Jp, double* __restrict__ acc,
                                                          const int* __restrict__ done_flag) {
    if (*done_flag) return;
    __shared__ double sm[8][kMom][33];
    const int tid = threadIdx.x, lane = tid & 31, rg = tid >> 5;
    const int j = blockIdx.x * 32 + lane;
    double v[kMom];
#pragma unroll
    for (int k = 0; k < kMom; ++k) v[k] = 0.0;
    for (int r = rg; r < rows; r += 8) {
        const float* src = partial + (size_t)r * kMom * Jp + j;
#pragma unroll
        for (int k = 0; k < kMom; ++k) v[k] += (double)src[(size_t)k * Jp];
    }
#pragma unroll
    for (int k = 0; k < kMom; ++k) sm[rg][k][lane] = v[k];
    __syncthreads();
    if (rg == 0 && j < J) {
#pragma unroll
        for (int k = 0; k < kMom; ++k) {
            double t = 0.0;
            for (int w = 0; w < 8; ++w) t += sm[w][k][lane];
            acc[kAccHdr + (size_t)j * kMom + k] = t;
        }
    }
    if (blockIdx.x == 0 && rg == 1) {
        double l = 0.0, c = 0.0;
        for (int r = lane; r < rows; r += 32) {
            l += rowaux[2 * r];
            c += rowaux[2 * r + 1];
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

// picks (JT, W, Sdiv, G), the point-batch size PB and the grid for a problem; pb_request: 4, 8 or 0 = auto
void flat2_plan(int n, int Jp, int num_sms, int pb_request, int* JT, int* W, int* Sdiv, int* G, int* PB, int* grid) {
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
    int pb = (pb_request == 4 || pb_request == 8) ? pb_request : 8;
    if (pb == 4 && w > 13) pb = 8;
    int ctas;
    if (pb == 4)
        ctas = jt == 1 ? occupancy_grid(em_flat2_kernel<1, 4, 416, 2>, w * 32, num_sms) : occupancy_grid(em_flat2_kernel<2, 4, 416, 2>, w * 32, num_sms);
    else
        ctas = jt == 1 ? occupancy_grid(em_flat2_kernel<1, 8, 512, 1>, w * 32, num_sms) : occupancy_grid(em_flat2_kernel<2, 8, 512, 1>, w * 32, num_sms);
    // never so many CTAs that a group sees fewer than ~2 batches
    const long long min_pts = 2LL * pb * g;
    if ((long long)ctas * min_pts > n) ctas = (int)((n + min_pts - 1) / min_pts);
    if (ctas < 1) ctas = 1;
    *JT = jt; *Sdiv = sdiv; *G = g; *W = w; *PB = pb; *grid = ctas;
}

template <int JT>
static cudaError_t launch_em_flat2_t(const float* x, const float* y, const float* z, int n, const FlatModel& m, int W, int Sdiv,
                                     int G, int grid, int PB, float* partial, double* rowaux, const int* done_flag, cudaStream_t s) {
    const float eps_on = m.flavor == HGMM_FLAVOR_PY ? 1.f : 0.f;
    if (PB == 4)
        em_flat2_kernel<JT, 4, 416, 2><<<grid, W * 32, 0, s>>>(x, y, z, n, m.packed, m.J, m.Jp, Sdiv, G, partial, rowaux, done_flag, eps_on);
    else
        em_flat2_kernel<JT, 8, 512, 1><<<grid, W * 32, 0, s>>>(x, y, z, n, m.packed, m.J, m.Jp, Sdiv, G, partial, rowaux, done_flag, eps_on);
    return cudaGetLastError();
}

cudaError_t launch_em_flat2(const float* x, const float* y, const float* z, int n, const FlatModel& m, int JT, int W, int Sdiv,
                            int G, int grid, int PB, float* partial, double* rowaux, const int* done_flag, cudaStream_t s) {
    switch (JT) {
        case 1: return launch_em_flat2_t<1>(x, y, z, n, m, W, Sdiv, G, grid, PB, partial, rowaux, done_flag, s);
        default: return launch_em_flat2_t<2>(x, y, z, n, m, W, Sdiv, G, grid, PB, partial, rowaux, done_flag, s);
    }
}

cudaError_t launch_flat_reduce(const float* partial, const double* rowaux, int rows, const FlatModel& m, double* acc,
                               const int* done_flag, cudaStream_t s) {
    flat_reduce_kernel<<<m.Jp / 32, 256, 0, s>>>(partial, rowaux, rows, m.J, m.Jp, acc, done_flag);
    return cudaGetLastError();
}

}  // namespace hgmm
