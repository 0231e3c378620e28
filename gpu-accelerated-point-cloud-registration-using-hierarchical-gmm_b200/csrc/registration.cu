// registration.cu -- tree registration: descent E-step + single-CTA SE(3) solves, sm_100a.
//
// Replaces (paths relative to the reference checkout):
//   gmmTreeRegESTep (pure-Python per-point loop)      src/python/hgmm/hgmm_gpu.py:550-577
//   GMMTree.maximization_step (eigh + lstsq on host)  src/python/hgmm/hgmm_gpu.py:729-752
//   twist_mul / twist_trans / skew                    src/python/hgmm/hgmm_gpu.py:620-664
//   svd + Procrustes use (host, ICP only)             src/c++/common/svd3.h:355-401, icp/icp_kernel.cu:697-731
//   GMMRegistration::pointCloudRegisterGPU            src/c++/gmm_registration/gmm_reg.cu:54-56 (empty stub)
//   kernCopyPositionsToVBO / kernCopyVelocitiesToVBO  src/c++/gmm_registration/gmm_reg_kernels.cu:3-41
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace hgmm {

constexpr int kRegMom = 10;     // M0, M1(3), raw M2 (xx xy xz yy yz zz) -- fp64, uncentred

// ------------------------------------------------------------------------------------------
// E-step: one thread per target point, greedy root->leaf descent
// ------------------------------------------------------------------------------------------
constexpr int kTopNodes = 72;       // levels 0 and 1 (8 + 64 nodes) are folded per CTA without atomics

// E-step: one thread per target point, greedy root->leaf descent.
// Every point visits level 0 and most visit level 1, i.e. 40k points would land on 8 + 64 addresses.  Those two
// levels are therefore NOT accumulated with atomics: each thread parks (node, gamma) for its level-0 and level-1 visit
// in shared memory, and after the descent warp w sums the parked entries of "its" nodes (node w of level 0, nodes
// 8+8w..8+8w+7 of level 1) in fp64 in a fixed order, one fp64 global atomic per (node, moment, CTA).  Deeper levels
// (>= 512 nodes) go straight to fp64 global atomics.
__global__ void __launch_bounds__(256) reg_estep_kernel(const float* __restrict__ tx, const float* __restrict__ ty,
                                                        const float* __restrict__ tz, int n, const double* __restrict__ Rt,
                                                        const PackedComp* __restrict__ packed, const float* __restrict__ cplx,
                                                        int L, float lambda_c, double* __restrict__ racc, int want_m2,
                                                        const int* __restrict__ ctrl) {
    if (ctrl[0]) return;
    __shared__ float s_x[256], s_y[256], s_z[256];
    __shared__ float s_g[2][256];
    __shared__ short s_node[2][256];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int i = blockIdx.x * blockDim.x + tid;
    s_node[0][tid] = -1;
    s_node[1][tid] = -1;
    if (i < n) {
        float x, y, z;
        {
            const double a = tx[i], b = ty[i], c = tz[i];      // t_target = target R^T + t  (hgmm_gpu.py:757,613-614)
            x = (float)(Rt[0] * a + Rt[1] * b + Rt[2] * c + Rt[9]);
            y = (float)(Rt[3] * a + Rt[4] * b + Rt[5] * c + Rt[10]);
            z = (float)(Rt[6] * a + Rt[7] * b + Rt[8] * c + Rt[11]);
        }
        s_x[tid] = x;
        s_y[tid] = y;
        s_z[tid] = z;
        int j0 = 0;                                             // child(-1) = 0
        for (int l = 0; l < L; ++l) {
            const float4* c4 = reinterpret_cast<const float4*>(packed + j0);
            float q[8];
            float m = kNegBig;
            int best = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float4 p0 = __ldg(c4 + 3 * k), p1 = __ldg(c4 + 3 * k + 1);
                const float4 p2 = __ldg(c4 + 3 * k + 2);
                float dx, dy, dz;
                q[k] = quad_q2(p0, p1, make_float2(p2.x, p2.y), x, y, z, dx, dy, dz);
                if (q[k] > m) { m = q[k]; best = k; }
            }
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) s += ex2f(q[k] - m);
            const float lse2 = m + lg2f(s);
            const bool alive = lse2 > kLog2Eps15;               // den > eps else gamma = zeros (:563-567)
            const int sid = j0 + (alive ? best : 0);
            if (cplx[sid] <= lambda_c) break;                   // :572-573, before accumulating
            const float gam = alive ? 1.0f / s : 0.f;           // gamma of the arg-max child
            if (gam >= 1e-15f) {                                // accumulate() guard (:457-459)
                if (l < 2) {
                    s_node[l][tid] = (short)sid;
                    s_g[l][tid] = gam;
                } else {
                    double* A = racc + (size_t)sid * kRegMom;
                    const double g = gam, X = x, Y = y, Z = z;
                    atomicAdd(A + 0, g);
                    atomicAdd(A + 1, g * X);
                    atomicAdd(A + 2, g * Y);
                    atomicAdd(A + 3, g * Z);
                    if (want_m2) {
                        atomicAdd(A + 4, g * X * X);
                        atomicAdd(A + 5, g * X * Y);
                        atomicAdd(A + 6, g * X * Z);
                        atomicAdd(A + 7, g * Y * Y);
                        atomicAdd(A + 8, g * Y * Z);
                        atomicAdd(A + 9, g * Z * Z);
                    }
                }
            }
            j0 = (sid + 1) * 8;
        }
    }
    __syncthreads();
    // ---- fold levels 0 and 1: warp w owns node w (level 0) and nodes 8 + 8w .. 8 + 8w + 7 (level 1)
    const int nm = want_m2 ? kRegMom : 4;
    for (int slot = 0; slot < 9; ++slot) {
        const int lvl = slot == 0 ? 0 : 1;
        const int node = slot == 0 ? warp : 8 + 8 * warp + (slot - 1);
        double v[kRegMom];
#pragma unroll
        for (int k = 0; k < kRegMom; ++k) v[k] = 0.0;
        for (int t = lane; t < 256; t += 32) {
            if (s_node[lvl][t] == node) {
                const double g = s_g[lvl][t], X = s_x[t], Y = s_y[t], Z = s_z[t];
                v[0] += g; v[1] += g * X; v[2] += g * Y; v[3] += g * Z;
                if (want_m2) {
                    v[4] += g * X * X; v[5] += g * X * Y; v[6] += g * X * Z; v[7] += g * Y * Y; v[8] += g * Y * Z; v[9] += g * Z * Z;
                }
            }
        }
        const unsigned any = __ballot_sync(0xffffffffu, v[0] != 0.0);
        if (any == 0u) continue;                               // warp-uniform
        for (int k = 0; k < nm; ++k) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
        }
        if (lane == 0) {
            for (int k = 0; k < nm; ++k)
                if (v[k] != 0.0) atomicAdd(racc + (size_t)node * kRegMom + k, v[k]);
        }
    }
}

// ------------------------------------------------------------------------------------------
// E-step, second generation (DRAFT: compiled, not yet run on a GPU -- selected only by HGMM_REG_ESTEP2=1).
// reg_estep_kernel is a latency-bound single wave (profiles/r01_reg_estep_ncu_full.txt: issue-active 9 %, top stalls
// short_scoreboard / barrier / long_scoreboard): every thread walks its own point and reads the 8 children of its own parent,
// so from level 1 on a warp-wide LDG.128 touches up to 32 different lines (24 such loads per level), and the fold of the two
// top levels scans all 256 parked entries nine times per warp.
// Here 8 lanes share a point, lane = child: one 384-byte contiguous block of 8 children per point and level (3 coalesced
// LDG.128 per lane instead of 24 scattered ones), the arg-max / sum over the 8 children are three xor-shuffles inside the
// 8-lane group, and 8x more warps are in flight for the same cloud.  Levels 0 and 1 (72 nodes) accumulate with shared-memory
// fp64 atomics (one flush of the non-zero entries per CTA at the end), deeper levels with global fp64 atomics as before.
// Semantics are those of reg_estep_kernel / gmmTreeRegESTep (hgmm_gpu.py:550-577): first-maximum arg-max, gamma = 0 when the
// eight densities sum below 1e-15, stop BEFORE accumulating at a node whose complexity is <= lambda_c, skip gamma < 1e-15.
// The eight exponentials are summed by a shuffle tree instead of left to right (last-bit differences in gamma).
__global__ void __launch_bounds__(256) reg_estep2_kernel(const float* __restrict__ tx, const float* __restrict__ ty,
                                                         const float* __restrict__ tz, int n, const double* __restrict__ Rt,
                                                         const PackedComp* __restrict__ packed, const float* __restrict__ cplx,
                                                         int L, float lambda_c, double* __restrict__ racc, int want_m2,
                                                         const int* __restrict__ ctrl) {
    if (ctrl[0]) return;
    __shared__ double s_acc[kTopNodes][kRegMom];
    const int tid = threadIdx.x, lane = tid & 31;
    const int grp = lane >> 3, c = lane & 7;                    // point slot inside the warp, child
    const unsigned gmask = 0xffu << (grp * 8);
    for (int k = tid; k < kTopNodes * kRegMom; k += blockDim.x) (&s_acc[0][0])[k] = 0.0;
    __syncthreads();
    const int nm = want_m2 ? kRegMom : 4;
    const int warps_total = gridDim.x * (blockDim.x >> 5);
    const int gwarp = blockIdx.x * (blockDim.x >> 5) + (tid >> 5);
    for (int base = gwarp * 4; base < n; base += warps_total * 4) {       // warp-uniform trip count
        const int i = base + grp;
        bool active = i < n;
        float x = 0.f, y = 0.f, z = 0.f;
        if (active) {
            const double a = tx[i], b = ty[i], cc = tz[i];      // t_target = target R^T + t  (hgmm_gpu.py:757,613-614)
            x = (float)(Rt[0] * a + Rt[1] * b + Rt[2] * cc + Rt[9]);
            y = (float)(Rt[3] * a + Rt[4] * b + Rt[5] * cc + Rt[10]);
            z = (float)(Rt[6] * a + Rt[7] * b + Rt[8] * cc + Rt[11]);
        }
        int j0 = 0;                                             // child(-1) = 0
        for (int l = 0; l < L; ++l) {                           // every lane runs all L levels: the shuffles stay convergent
            float q = kNegBig;
            if (active) {
                const float4* c4 = reinterpret_cast<const float4*>(packed + j0 + c);
                const float4 p0 = __ldg(c4), p1 = __ldg(c4 + 1), p2 = __ldg(c4 + 2);
                float dx, dy, dz;
                q = quad_q2(p0, p1, make_float2(p2.x, p2.y), x, y, z, dx, dy, dz);
            }
            float m = q;
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            m = fmaxf(m, kNegBig);
            const unsigned hit = __ballot_sync(0xffffffffu, q == m) & gmask;      // first maximum of the group
            const int best = hit ? (__ffs(hit) - 1 - grp * 8) : 0;
            float sE = ex2f(q - m);
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) sE += __shfl_xor_sync(0xffffffffu, sE, o);
            const float lse2 = m + lg2f(sE);
            const bool alive = lse2 > kLog2Eps15;               // den > eps else gamma = zeros (:563-567)
            const int sid = j0 + (alive ? best : 0);
            if (active && __ldg(cplx + sid) <= lambda_c) active = false;          // :572-573, before accumulating
            const float gam = alive ? 1.0f / sE : 0.f;          // gamma of the arg-max child
            if (active && c == 0 && gam >= 1e-15f) {            // the group's leader accumulates (accumulate() guard :457-459)
                const double g = gam, X = x, Y = y, Z = z;
                double v[kRegMom] = {g, g * X, g * Y, g * Z, 0, 0, 0, 0, 0, 0};
                if (want_m2) {
                    v[4] = g * X * X; v[5] = g * X * Y; v[6] = g * X * Z; v[7] = g * Y * Y; v[8] = g * Y * Z; v[9] = g * Z * Z;
                }
                double* A = sid < kTopNodes ? &s_acc[sid][0] : racc + (size_t)sid * kRegMom;
#pragma unroll
                for (int k = 0; k < kRegMom; ++k)
                    if (k < nm) atomicAdd(A + k, v[k]);
            }
            j0 = (sid + 1) * 8;
        }
    }
    __syncthreads();
    for (int k = tid; k < kTopNodes * kRegMom; k += blockDim.x) {
        const double v = (&s_acc[0][0])[k];
        if (v != 0.0) atomicAdd(racc + k, v);                   // racc is [node][kRegMom]: same flat index
    }
}

// ------------------------------------------------------------------------------------------
// E-step, third generation (default).  What the ncu capture of reg_estep_kernel shows (profiles/r01_reg_estep_ncu_full.txt):
// 1830 warp instructions per 32 points at ~24 stall cycles each, most of them in the fold of the two top levels (every warp
// scans all 256 parked entries nine times in fp64) and at its barrier; only 8.5 warps per SM exist for a 40k-point cloud.
// Here the descent only RECORDS (node, gamma) per level in registers; the accumulation is a second, warp-cooperative phase:
// per level the warp walks its DISTINCT nodes (range scans are spatially coherent: 1-3 per warp), reduces the members'
// (gamma, gamma x [, gamma x x^T]) with a masked fp64 butterfly and the leader lane adds the result once -- to a small
// shared-memory table for the 72 nodes of the two top levels (flushed once per CTA), to L2 for deeper ones.  No fold pass,
// no barrier before the end, 128-thread CTAs (several per SM, one wave).  Semantics are those of reg_estep_kernel /
// gmmTreeRegESTep (hgmm_gpu.py:550-577): first-maximum arg-max, gamma = 0 when the eight densities sum below 1e-15, stop
// BEFORE accumulating at a node whose complexity is <= lambda_c, skip gamma < 1e-15.
// ------------------------------------------------------------------------------------------
constexpr int kRegMaxL = 6;
template <int NM>
__device__ __forceinline__ void reg_accumulate_level(int key, float gam, float x, float y, float z, int lane, double (*s_acc)[kRegMom],
                                                     double* __restrict__ racc) {
    unsigned todo = __ballot_sync(0xffffffffu, key >= 0);
    while (todo) {                                              // warp-uniform: one round per distinct node of the warp
        const int leader = __ffs(todo) - 1;
        const int k = __shfl_sync(0xffffffffu, key, leader);
        const bool mine = key == k;
        const double g = mine ? (double)gam : 0.0, X = x, Y = y, Z = z;
        double v[NM];
        v[0] = g; v[1] = g * X; v[2] = g * Y; v[3] = g * Z;
        if (NM > 4) { v[4] = g * X * X; v[5] = g * X * Y; v[6] = g * X * Z; v[7] = g * Y * Y; v[8] = g * Y * Z; v[9] = g * Z * Z; }
#pragma unroll
        for (int i = 0; i < NM; ++i) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
        }
        if (lane == leader) {                                   // two branches, not one generic pointer: the global adds are then
            if (k < kTopNodes) {                                // native RED.ADD.F64 instead of a generic-address CAS loop
#pragma unroll
                for (int i = 0; i < NM; ++i) atomicAdd(&s_acc[k][i], v[i]);
            } else {
                double* A = racc + (size_t)k * kRegMom;
#pragma unroll
                for (int i = 0; i < NM; ++i) atomicAdd(A + i, v[i]);
            }
        }
        todo &= ~__ballot_sync(0xffffffffu, mine);
    }
}

__global__ void __launch_bounds__(128) reg_estep3_kernel(const float* __restrict__ tx, const float* __restrict__ ty,
                                                         const float* __restrict__ tz, int n, const double* __restrict__ Rt,
                                                         const PackedComp* __restrict__ packed, const float* __restrict__ cplx,
                                                         int L, float lambda_c, double* __restrict__ racc, int want_m2,
                                                         const int* __restrict__ ctrl) {
    // programmatic dependent launch: the grid is scheduled while the previous solve drains; everything that kernel wrote
    // (the transform, the stop flag, the zeroed moments) is read after the wait
    __shared__ double s_acc[kTopNodes][kRegMom];
    const int tid = threadIdx.x, lane = tid & 31;
    for (int k = tid; k < kTopNodes * kRegMom; k += blockDim.x) (&s_acc[0][0])[k] = 0.0;
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (__ldcg(ctrl)) return;
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + tid;
    float x = 0.f, y = 0.f, z = 0.f;
    int node[kRegMaxL];
    float gam[kRegMaxL];
#pragma unroll
    for (int l = 0; l < kRegMaxL; ++l) { node[l] = -1; gam[l] = 0.f; }
    if (i < n) {
        const double a = tx[i], b = ty[i], c = tz[i];           // t_target = target R^T + t  (hgmm_gpu.py:757,613-614)
        x = (float)(__ldcg(Rt + 0) * a + __ldcg(Rt + 1) * b + __ldcg(Rt + 2) * c + __ldcg(Rt + 9));
        y = (float)(__ldcg(Rt + 3) * a + __ldcg(Rt + 4) * b + __ldcg(Rt + 5) * c + __ldcg(Rt + 10));
        z = (float)(__ldcg(Rt + 6) * a + __ldcg(Rt + 7) * b + __ldcg(Rt + 8) * c + __ldcg(Rt + 11));
        int j0 = 0;                                             // child(-1) = 0
#pragma unroll
        for (int l = 0; l < kRegMaxL; ++l) {
            if (l >= L) break;
            const float4* c4 = reinterpret_cast<const float4*>(packed + j0);
            float q[8];
            float m = kNegBig;
            int best = 0;
#pragma unroll
            float cx_best = 0.f, cx0 = 0.f;                     // complexity rides in PackedComp::pad0 (tree_cplx_kernel): no
            for (int k = 0; k < 8; ++k) {                       // second, dependent load per level
                const float4 p0 = __ldg(c4 + 3 * k), p1 = __ldg(c4 + 3 * k + 1);
                const float4 p2 = __ldg(c4 + 3 * k + 2);
                float dx, dy, dz;
                q[k] = quad_q2(p0, p1, make_float2(p2.x, p2.y), x, y, z, dx, dy, dz);
                if (k == 0) cx0 = p2.z;
                if (q[k] > m) { m = q[k]; best = k; cx_best = p2.z; }
            }
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) s += ex2f(q[k] - m);
            const float lse2 = m + lg2f(s);
            const bool alive = lse2 > kLog2Eps15;               // den > eps else gamma = zeros (:563-567)
            const int sid = j0 + (alive ? best : 0);
            if ((alive ? cx_best : cx0) <= lambda_c) break;     // :572-573, before accumulating
            const float g = alive ? 1.0f / s : 0.f;             // gamma of the arg-max child
            if (g >= 1e-15f) {                                  // accumulate() guard (:457-459)
                node[l] = sid;
                gam[l] = g;
            }
            j0 = (sid + 1) * 8;
        }
    }
#pragma unroll
    for (int l = 0; l < kRegMaxL; ++l) {
        if (l >= L) break;                                      // L is warp-uniform: every lane runs the same rounds
        if (want_m2) reg_accumulate_level<kRegMom>(node[l], gam[l], x, y, z, lane, s_acc, racc);
        else reg_accumulate_level<4>(node[l], gam[l], x, y, z, lane, s_acc, racc);
    }
    __syncthreads();
    for (int k = tid; k < kTopNodes * kRegMom; k += blockDim.x) {
        const double v = (&s_acc[0][0])[k];
        if (v != 0.0) atomicAdd(racc + k, v);                   // racc is [node][kRegMom]: same flat index
    }
}

// ------------------------------------------------------------------------------------------
// M-step solves.  One CTA; per-thread partial sums over nodes, fixed-order tree reduction.
// ------------------------------------------------------------------------------------------
constexpr int kSys = 28;        // H (21 upper-triangular) + g (6) + c (1)
constexpr int kPro = 18;        // W, sum w s (3), sum w mu (3), sum w s mu^T (9), sum w |s|^2, sum w |mu|^2

// fixed-order block sum of NV doubles per thread: xor-shuffle inside each warp, then thread 0 adds the warp totals
template <int NV>
__device__ void block_reduce(double* v, double* sm /*[nwarps][NV]*/, int tid, int nthreads) {
    const int lane = tid & 31, warp = tid >> 5, nw = nthreads >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    }
    if (lane == 0)
        for (int k = 0; k < NV; ++k) sm[warp * NV + k] = v[k];
    __syncthreads();
    if (tid == 0) {
        for (int k = 0; k < NV; ++k) {
            double t = 0.0;
            for (int w = 0; w < nw; ++w) t += sm[w * NV + k];
            v[k] = t;
            sm[k] = t;        // the totals also stay in sm[0, NV): column k has been read, later columns are untouched
        }
    }
}

__device__ void rodrigues(const double* w, double* R) {      // twist_trans (hgmm_gpu.py:646-664)
    const double th = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    if (th == 0.0) {
        R[0] = R[4] = R[8] = 1.0;
        R[1] = R[2] = R[3] = R[5] = R[6] = R[7] = 0.0;
        return;
    }
    const double ith = 1.0 / th;
    const double nx = w[0] * ith, ny = w[1] * ith, nz = w[2] * ith;
    double s, c;
    sincos(th, &s, &c);
    const double oc = 1.0 - c;
    R[0] = c + oc * nx * nx;      R[1] = oc * nx * ny - s * nz; R[2] = oc * nx * nz + s * ny;
    R[3] = oc * ny * nx + s * nz; R[4] = c + oc * ny * ny;      R[5] = oc * ny * nz - s * nx;
    R[6] = oc * nz * nx - s * ny; R[7] = oc * nz * ny + s * nx; R[8] = c + oc * nz * nz;
}

// (R,t) <- (dR R, dR t + dt)   (twist_mul, hgmm_gpu.py:634-644)
__device__ void compose(const double* dR, const double* dt, double* Rt) {
    double Rn[9], tn[3];
    for (int a = 0; a < 3; ++a) {
        for (int b = 0; b < 3; ++b) Rn[3 * a + b] = dR[3 * a] * Rt[b] + dR[3 * a + 1] * Rt[3 + b] + dR[3 * a + 2] * Rt[6 + b];
        tn[a] = dR[3 * a] * Rt[9] + dR[3 * a + 1] * Rt[10] + dR[3 * a + 2] * Rt[11] + dt[a];
    }
    for (int k = 0; k < 9; ++k) Rt[k] = Rn[k];
    for (int k = 0; k < 3; ++k) Rt[9 + k] = tn[k];
}

// in-place Cholesky solve of the 6x6 SPD system; returns false when not positive definite.
// The solve runs on ONE thread between two kernels of every registration iteration, i.e. on the iteration's critical path: the
// 6 square roots and 27 divisions of the textbook form (each a ~30-instruction fp64 sequence) are replaced by 6 reciprocal square
// roots (rsqrt: MUFU.RSQ64H + Newton steps, 1 ulp) and multiplications.
__device__ bool chol6_solve(double* H /*[36]*/, double* g /*[6] in, x out*/) {
    double invd[6];
    for (int j = 0; j < 6; ++j) {
        double d = H[7 * j];
        for (int k = 0; k < j; ++k) d -= H[6 * j + k] * H[6 * j + k];
        if (!(d > 0.0)) return false;
        const double r = rsqrt(d);
        invd[j] = r;
        H[7 * j] = d * r;
        for (int i = j + 1; i < 6; ++i) {
            double v = H[6 * i + j];
            for (int k = 0; k < j; ++k) v -= H[6 * i + k] * H[6 * j + k];
            H[6 * i + j] = v * r;
        }
    }
    for (int i = 0; i < 6; ++i) {
        double v = g[i];
        for (int k = 0; k < i; ++k) v -= H[6 * i + k] * g[k];
        g[i] = v * invd[i];
    }
    for (int i = 5; i >= 0; --i) {
        double v = g[i];
        for (int k = i + 1; k < 6; ++k) v -= H[6 * k + i] * g[k];
        g[i] = v * invd[i];
    }
    return true;
}

// Jacobi eigen-analysis of the symmetric 3x3 S (in place -> diagonal), V accumulates rotations.
// Same structure as svd3.h jacobiEigenanlysis (:199-241) but exact Givens angles in fp64 and
// sweeps until the off-diagonal mass is negligible instead of 4 approximate sweeps.
__device__ void jacobi3(double S[3][3], double V[3][3]) {
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) V[a][b] = (a == b) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 16; ++sweep) {
        const double off = S[0][1] * S[0][1] + S[0][2] * S[0][2] + S[1][2] * S[1][2];
        const double dia = S[0][0] * S[0][0] + S[1][1] * S[1][1] + S[2][2] * S[2][2];
        if (off <= 1e-32 * dia || off == 0.0) break;
        for (int pq = 0; pq < 3; ++pq) {
            const int p = (pq == 2) ? 0 : pq, q = (pq == 0) ? 1 : 2;     // (0,1) (1,2) (0,2) as svd3.h
            if (S[p][q] == 0.0) continue;
            const double theta = (S[q][q] - S[p][p]) / (2.0 * S[p][q]);
            const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
            const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
            for (int k = 0; k < 3; ++k) {       // S <- S G
                const double a = S[k][p], b = S[k][q];
                S[k][p] = c * a - s * b;
                S[k][q] = s * a + c * b;
            }
            for (int k = 0; k < 3; ++k) {       // S <- G^T S
                const double a = S[p][k], b = S[q][k];
                S[p][k] = c * a - s * b;
                S[q][k] = s * a + c * b;
            }
            for (int k = 0; k < 3; ++k) {
                const double a = V[k][p], b = V[k][q];
                V[k][p] = c * a - s * b;
                V[k][q] = s * a + c * b;
            }
        }
    }
}

// SVD of A (3x3) = U diag(sig) V^T following svd3.h:355-401: V from A^T A, B = A V, sort columns by
// norm (descending), U from Gram-Schmidt QR of B.
__device__ void svd3(const double A[3][3], double U[3][3], double sig[3], double V[3][3]) {
    double S[3][3];
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) S[a][b] = A[0][a] * A[0][b] + A[1][a] * A[1][b] + A[2][a] * A[2][b];
    jacobi3(S, V);
    double B[3][3];
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) B[a][b] = A[a][0] * V[0][b] + A[a][1] * V[1][b] + A[a][2] * V[2][b];
    double nrm[3];
    for (int b = 0; b < 3; ++b) nrm[b] = B[0][b] * B[0][b] + B[1][b] * B[1][b] + B[2][b] * B[2][b];
    for (int pass = 0; pass < 3; ++pass) {      // sortSingularValues (svd3.h:243-270): (0,1) (0,2) (1,2)
        const int ci = (pass == 2) ? 1 : 0, cj = (pass == 0) ? 1 : 2;
        if (nrm[ci] < nrm[cj]) {
            for (int k = 0; k < 3; ++k) {
                double t = B[k][ci]; B[k][ci] = B[k][cj]; B[k][cj] = -t;     // negated swap keeps det V = +1
                t = V[k][ci]; V[k][ci] = V[k][cj]; V[k][cj] = -t;
            }
            const double t = nrm[ci]; nrm[ci] = nrm[cj]; nrm[cj] = t;
        }
    }
    // QR by modified Gram-Schmidt; degenerate columns completed to a right-handed frame
    for (int b = 0; b < 3; ++b) {
        double v[3] = {B[0][b], B[1][b], B[2][b]};
        for (int p = 0; p < b; ++p) {
            const double d = U[0][p] * v[0] + U[1][p] * v[1] + U[2][p] * v[2];
            for (int k = 0; k < 3; ++k) v[k] -= d * U[k][p];
        }
        double n = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
        sig[b] = n;
        if (n > 1e-300 && (b == 0 || n > 1e-14 * sig[0])) {
            for (int k = 0; k < 3; ++k) U[k][b] = v[k] / n;
        } else {
            sig[b] = 0.0;
            if (b == 2) {       // cross product of the first two
                U[0][2] = U[1][0] * U[2][1] - U[2][0] * U[1][1];
                U[1][2] = U[2][0] * U[0][1] - U[0][0] * U[2][1];
                U[2][2] = U[0][0] * U[1][1] - U[1][0] * U[0][1];
            } else {            // any unit vector orthogonal to the previous columns
                double e[3] = {0, 0, 0};
                int ax = 0;
                if (b == 1) {
                    const double ax0 = fabs(U[0][0]), ax1 = fabs(U[1][0]), ax2 = fabs(U[2][0]);
                    ax = (ax0 <= ax1 && ax0 <= ax2) ? 0 : (ax1 <= ax2 ? 1 : 2);
                }
                e[ax] = 1.0;
                for (int p = 0; p < b; ++p) {
                    const double d = U[0][p] * e[0] + U[1][p] * e[1] + U[2][p] * e[2];
                    for (int k = 0; k < 3; ++k) e[k] -= d * U[k][p];
                }
                n = sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
                for (int k = 0; k < 3; ++k) U[k][b] = e[k] / n;
            }
        }
    }
}

__device__ double det3(const double M[3][3]) {
    return M[0][0] * (M[1][1] * M[2][2] - M[1][2] * M[2][1]) - M[0][1] * (M[1][0] * M[2][2] - M[1][2] * M[2][0]) +
           M[0][2] * (M[1][0] * M[2][1] - M[1][1] * M[2][0]);
}

// ---- the two solves from their reduced sums (thread 0 of the solve kernels) ----
// `v` is the block's SHARED-memory copy of the totals (block_reduce leaves it in sm[0, NV)).  Passing the kernel's per-thread
// array instead was miscompiled by nvcc 12.9 for sm_100a once these bodies became functions shared by two kernels: after
// inlining, H[] was given the stack slots of the still-live v[] (the dumped system showed H's symmetric fill inside v) and
// every twist solve failed as "not positive definite".
// v: H (21 upper-triangular) | g (6) | c.  x = H^-1 g (Cholesky), q = c - g^T x (= lstsq's residual, hgmm_gpu.py:746),
// (R,t) <- twist_mul(x, R, t); stopping rule |q - q_prev| < tol (hgmm_gpu.py:765)
__device__ __forceinline__ void twist_finish(const double* v, double* Rt, double* q_hist, double* qstate, int* ctrl, float tol) {
    double H[36], g[6], g0[6];
    int k = 0;
    for (int a = 0; a < 6; ++a)
        for (int b = a; b < 6; ++b) { H[6 * a + b] = v[k]; H[6 * b + a] = v[k]; ++k; }
    for (int a = 0; a < 6; ++a) g[a] = g0[a] = v[21 + a];
    double q;
    if (chol6_solve(H, g)) {
        q = v[27];
        for (int a = 0; a < 6; ++a) q -= g0[a] * g[a];          // residual sum of squares = c - g^T x
        double dR[9];
        rodrigues(g, dR);
        compose(dR, g + 3, Rt);
    } else {
        q = nan("");
        ctrl[2] = 1;
        ctrl[0] = 1;
        qstate[3] = v[0];                 // diagnostics for the host's error message: the failed system itself
        for (int i = 0; i < kSys; ++i) q_hist[8 + i] = v[i];
    }
    const int it = ctrl[1];
    q_hist[it] = q;
    if (qstate[2] != 0.0 && fabs(q - qstate[0]) < (double)tol) ctrl[0] = 1;   // hgmm_gpu.py:765
    qstate[0] = q;
    qstate[1] = q;
    qstate[2] = 1.0;
    ctrl[1] = it + 1;
}

// v: W | sum w s (3) | sum w mu (3) | sum w s mu^T (9) | sum w |s|^2 | sum w |mu|^2  -> weighted Procrustes (3x3 SVD)
__device__ __forceinline__ void procrustes_finish(const double* v, double* Rt, double* q_hist, double* qstate, int* ctrl, float tol) {
    double q = nan("");
    if (v[0] > 0.0) {
        const double W = v[0];
        const double sb[3] = {v[1] / W, v[2] / W, v[3] / W}, mb[3] = {v[4] / W, v[5] / W, v[6] / W};
        double Hm[3][3];
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) Hm[a][b] = v[7 + 3 * a + b] - W * sb[a] * mb[b];    // sum w (s-sb)(m-mb)^T
        double U[3][3], sig[3], V[3][3];
        svd3(Hm, U, sig, V);
        // dR = V diag(1,1,d) U^T with d = det(V U^T)  (reflection fix the reference lacks, icp_kernel.cu:718-729)
        const double d = (det3(V) * det3(U) < 0.0) ? -1.0 : 1.0;
        double dR[9];
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) dR[3 * a + b] = V[a][0] * U[b][0] + V[a][1] * U[b][1] + d * V[a][2] * U[b][2];
        double dt[3];
        for (int a = 0; a < 3; ++a) dt[a] = mb[a] - (dR[3 * a] * sb[0] + dR[3 * a + 1] * sb[1] + dR[3 * a + 2] * sb[2]);
        // q = sum w |dR (s - sb) - (m - mb)|^2 = Sss + Smm - 2 tr(dR Hm)
        double tr = 0.0;
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) tr += dR[3 * a + b] * Hm[b][a];
        const double Sss = v[16] - W * (sb[0] * sb[0] + sb[1] * sb[1] + sb[2] * sb[2]);
        const double Smm = v[17] - W * (mb[0] * mb[0] + mb[1] * mb[1] + mb[2] * mb[2]);
        q = Sss + Smm - 2.0 * tr;
        compose(dR, dt, Rt);
    } else {
        ctrl[2] = 1;
        ctrl[0] = 1;
    }
    const int it = ctrl[1];
    q_hist[it] = q;
    if (qstate[2] != 0.0 && fabs(q - qstate[0]) < (double)tol) ctrl[0] = 1;
    qstate[0] = q;
    qstate[1] = q;
    qstate[2] = 1.0;
    ctrl[1] = it + 1;
}

// ---- multi-GPU: the ranks' (M0, M1) of node i, exchanged over peer memory INSIDE the solve kernel (no ncclAllReduce, no third
// launch).  The thread that owns node i stores this rank's four sums into every peer's region as self-validating 16-byte cells
// (fp64 as two (half, tag) words, tag = base + iteration + 1), polls its own region for the peers' cells of the same node, and adds
// the R contributions in rank order (own from registers): every rank ends with bit-identical sums, hence bit-identical transforms
// and the same stopping decision.  Two parities: a rank's next push needs every peer's current one.
// The cells are self-validating (every 8-byte word carries the tag), so the stores need no ordering among themselves: relaxed
// system-scope stores can be posted back to back, where st.volatile keeps a thread's stores in program order.
__device__ __forceinline__ void reg_cell_put(uint4* cell, double v, uint32_t tag) {
    const unsigned long long u = (unsigned long long)__double_as_longlong(v);
    asm volatile("st.relaxed.sys.global.v2.u32 [%0], {%1, %2};" ::"l"(cell), "r"((uint32_t)u), "r"(tag) : "memory");
    asm volatile("st.relaxed.sys.global.v2.u32 [%0], {%1, %2};" ::"l"(reinterpret_cast<char*>(cell) + 8), "r"((uint32_t)(u >> 32)), "r"(tag)
                 : "memory");
}
// push this rank's four sums of `node` into every peer's region (all of a thread's nodes are pushed before it waits for any)
__device__ __forceinline__ void reg_push_node(const RegXchgView& xv, int node, int par, uint32_t tag, const double* A) {
    const int R = xv.nranks, me = xv.rank;
    const size_t half = (size_t)kXchgMaxRanks * kRegXchgNodes * 4;
    for (int r = 0; r < R; ++r) {
        if (r == me) continue;
        uint4* dst = xv.data[r] + (size_t)par * half + ((size_t)me * kRegXchgNodes + node) * 4;
#pragma unroll
        for (int k = 0; k < 4; ++k) reg_cell_put(dst + k, A[k], tag);
    }
}
__device__ __forceinline__ bool reg_gather_node(const RegXchgView& xv, int node, int par, uint32_t tag, double* A /*[4] in: own, out: sum*/,
                                                int* ctrl) {
    const int R = xv.nranks, me = xv.rank;
    const size_t half = (size_t)kXchgMaxRanks * kRegXchgNodes * 4;
    double S[4] = {0.0, 0.0, 0.0, 0.0};
    unsigned long long t0 = 0ull;
    for (int r = 0; r < R; ++r) {
        if (r == me) {
#pragma unroll
            for (int k = 0; k < 4; ++k) S[k] += A[k];
            continue;
        }
        const uint4* src = xv.data[me] + (size_t)par * half + ((size_t)r * kRegXchgNodes + node) * 4;
        unsigned spins = 0;
        for (;;) {
            uint4 c[4];
#pragma unroll
            for (int k = 0; k < 4; ++k)
                asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(c[k].x), "=r"(c[k].y), "=r"(c[k].z), "=r"(c[k].w) : "l"(src + k) : "memory");
            bool all = true;
#pragma unroll
            for (int k = 0; k < 4; ++k) all = all && c[k].y == tag && c[k].w == tag;
            if (all) {
#pragma unroll
                for (int k = 0; k < 4; ++k) S[k] += __longlong_as_double((long long)(((unsigned long long)c[k].z << 32) | c[k].x));
                break;
            }
            if ((++spins & 255u) == 0u) {
                unsigned long long now;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                if (t0 == 0ull) t0 = now;
                int bad;
                asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(bad) : "l"(ctrl + 7) : "memory");
                if (bad || now - t0 > xv.timeout_ns) {
                    atomicExch(ctrl + 7, 1);
                    return false;
                }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) A[k] = S[k];
    return true;
}

// ctrl: [0] done, [1] iterations, [2] numeric failure flag, [7] a wait on a peer timed out.
// qstate: [0] previous q, [1] last q, [2] has-previous flag
__global__ void __launch_bounds__(512) reg_solve_kernel(TreeModel t, double* __restrict__ racc, int zero_after, int solver,
                                                        double* __restrict__ Rt, double* __restrict__ q_hist,
                                                        double* __restrict__ qstate, int* __restrict__ ctrl, float tol, RegXchgView xv) {
    asm volatile("griddepcontrol.wait;" ::: "memory");        // PDL-chained with the E-step (all of its atomics have landed)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (ctrl[0]) return;
    __shared__ double sm[16 * kSys];
    const int tid = threadIdx.x, nth = blockDim.x;
    const double f32eps = 1.1920928955078125e-07;            // np.finfo(np.float32).eps  (hgmm_gpu.py:741)
    const bool xchg = xv.nranks > 1;
    const int x_it = ctrl[1], x_par = x_it & 1;
    const uint32_t x_tag = xv.base + (uint32_t)x_it + 1u;
    if (xchg)                                                 // one NVLink hop for the whole tree: push everything, then collect
        for (int i = tid; i < t.nt; i += nth) reg_push_node(xv, i, x_par, x_tag, racc + (size_t)i * kRegMom);
    if (solver == HGMM_SOLVER_TWIST_LSTSQ) {
        double v[kSys];
        for (int k = 0; k < kSys; ++k) v[k] = 0.0;
        for (int i = tid; i < t.nt; i += nth) {
            double* A = racc + (size_t)i * kRegMom;
            double G4[4] = {A[0], A[1], A[2], A[3]};
            if (zero_after) { A[0] = 0.0; A[1] = 0.0; A[2] = 0.0; A[3] = 0.0; }
            if (xchg && !reg_gather_node(xv, i, x_par, x_tag, G4, ctrl)) break;
            const double M0 = G4[0], S1x = G4[1], S1y = G4[2], S1z = G4[3];
            if (M0 < f32eps) continue;
            const float* c = t.cov + 9 * i;
            Sym3 s{c[0], 0.5 * ((double)c[1] + c[3]), 0.5 * ((double)c[2] + c[6]), c[4], 0.5 * ((double)c[5] + c[7]), c[8]};
            const double det = sym3_det(s);
            if (!(det > 0.0)) continue;                     // the reference would emit inf/NaN rows here
            const Sym3 ad = sym3_adj(s);
            const double r = 1.0 / det;
            const double P[3][3] = {{ad.xx * r, ad.xy * r, ad.xz * r}, {ad.xy * r, ad.yy * r, ad.yz * r}, {ad.xz * r, ad.yz * r, ad.zz * r}};
            const double sx = S1x / M0, sy = S1y / M0, sz = S1z / M0;
            const double rr[3] = {t.mu[3 * i] - sx, t.mu[3 * i + 1] - sy, t.mu[3 * i + 2] - sz};
            // J = [ -[s]x | I ]  (3x6);  rows of J^T are the 6 unknowns
            const double Jm[3][6] = {{0, sz, -sy, 1, 0, 0}, {-sz, 0, sx, 0, 1, 0}, {sy, -sx, 0, 0, 0, 1}};
            double PJ[3][6];
            for (int a = 0; a < 3; ++a)
                for (int b = 0; b < 6; ++b) PJ[a][b] = P[a][0] * Jm[0][b] + P[a][1] * Jm[1][b] + P[a][2] * Jm[2][b];
            int k = 0;
            for (int a = 0; a < 6; ++a)
                for (int b = a; b < 6; ++b) v[k++] += M0 * (Jm[0][a] * PJ[0][b] + Jm[1][a] * PJ[1][b] + Jm[2][a] * PJ[2][b]);
            const double Pr[3] = {P[0][0] * rr[0] + P[0][1] * rr[1] + P[0][2] * rr[2], P[1][0] * rr[0] + P[1][1] * rr[1] + P[1][2] * rr[2],
                                  P[2][0] * rr[0] + P[2][1] * rr[1] + P[2][2] * rr[2]};
            for (int a = 0; a < 6; ++a) v[21 + a] += M0 * (Jm[0][a] * Pr[0] + Jm[1][a] * Pr[1] + Jm[2][a] * Pr[2]);
            v[27] += M0 * (rr[0] * Pr[0] + rr[1] * Pr[1] + rr[2] * Pr[2]);
        }
        block_reduce<kSys>(v, sm, tid, nth);
        if (tid == 0) twist_finish(sm, Rt, q_hist, qstate, ctrl, tol);
    } else {
        double v[kPro];
        for (int k = 0; k < kPro; ++k) v[k] = 0.0;
        for (int i = tid; i < t.nt; i += nth) {
            double* A = racc + (size_t)i * kRegMom;
            double G4[4] = {A[0], A[1], A[2], A[3]};
            if (zero_after) { A[0] = 0.0; A[1] = 0.0; A[2] = 0.0; A[3] = 0.0; }
            if (xchg && !reg_gather_node(xv, i, x_par, x_tag, G4, ctrl)) break;
            const double w = G4[0], S1x = G4[1], S1y = G4[2], S1z = G4[3];
            if (w < f32eps) continue;
            const double s[3] = {S1x / w, S1y / w, S1z / w};
            const double m[3] = {t.mu[3 * i], t.mu[3 * i + 1], t.mu[3 * i + 2]};
            v[0] += w;
            for (int a = 0; a < 3; ++a) {
                v[1 + a] += w * s[a];
                v[4 + a] += w * m[a];
                for (int b = 0; b < 3; ++b) v[7 + 3 * a + b] += w * s[a] * m[b];
            }
            v[16] += w * (s[0] * s[0] + s[1] * s[1] + s[2] * s[2]);
            v[17] += w * (m[0] * m[0] + m[1] * m[1] + m[2] * m[2]);
        }
        block_reduce<kPro>(v, sm, tid, nth);
        if (tid == 0) procrustes_finish(sm, Rt, q_hist, qstate, ctrl, tol);
    }
}

// ------------------------------------------------------------------------------------------
// Flat-mixture registration (BASELINE configs[3] / the north star's "E-step as point-to-mixture correspondence feeding a
// weighted-Procrustes solve"; the reference's own entry point for it, GMMRegistration::pointCloudRegisterGPU, is an empty
// stub -- src/c++/gmm_registration/gmm_reg.cu:54-56).  Per iteration: the target under the current (R,t) -> the SAME fused
// sweep the flat fit uses (responsibilities over all J components, centred moments) -> fixed-order fp64 reduction -> this
// kernel.  From the centred moments of component j (S0 = sum gamma, S1 = sum gamma (y - mu_j)): mass w_j = S0, centroid of
// the mass s_j = mu_j + S1/S0; Procrustes aligns {s_j} to {mu_j} with weights w_j, the twist solver minimises
// sum_j w_j ||J_j x - (mu_j - s_j)||^2 in the Sigma_j^-1 metric (same normal equations as the tree's, full covariances only).
// ------------------------------------------------------------------------------------------
__global__ void transform_soa_kernel(const float* __restrict__ tx, const float* __restrict__ ty, const float* __restrict__ tz, int n,
                                     const double* __restrict__ Rt, float* __restrict__ ox, float* __restrict__ oy,
                                     float* __restrict__ oz, const int* __restrict__ ctrl) {
    if (ctrl[0]) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double a = tx[i], b = ty[i], c = tz[i];
    ox[i] = (float)(Rt[0] * a + Rt[1] * b + Rt[2] * c + Rt[9]);
    oy[i] = (float)(Rt[3] * a + Rt[4] * b + Rt[5] * c + Rt[10]);
    oz[i] = (float)(Rt[6] * a + Rt[7] * b + Rt[8] * c + Rt[11]);
}

__global__ void __launch_bounds__(512) reg_flat_solve_kernel(FlatModel m, const double* __restrict__ acc, int solver,
                                                             double* __restrict__ Rt, double* __restrict__ q_hist,
                                                             double* __restrict__ qstate, int* __restrict__ ctrl, float tol) {
    if (ctrl[0]) return;
    __shared__ double sm[16 * kSys];
    const int tid = threadIdx.x, nth = blockDim.x;
    const double f32eps = 1.1920928955078125e-07;
    if (solver == HGMM_SOLVER_TWIST_LSTSQ) {
        double v[kSys];
        for (int k = 0; k < kSys; ++k) v[k] = 0.0;
        for (int j = tid; j < m.J; j += nth) {
            const double* A = acc + kAccHdr + (size_t)j * kMom;
            const double M0 = A[0];
            if (M0 < f32eps) continue;
            const float* c = m.covs + 9 * (size_t)j;
            Sym3 sg{c[0], 0.5 * ((double)c[1] + c[3]), 0.5 * ((double)c[2] + c[6]), c[4], 0.5 * ((double)c[5] + c[7]), c[8]};
            const double det = sym3_det(sg);
            if (!(det > 0.0)) continue;
            const Sym3 ad = sym3_adj(sg);
            const double r = 1.0 / det;
            const double P[3][3] = {{ad.xx * r, ad.xy * r, ad.xz * r}, {ad.xy * r, ad.yy * r, ad.yz * r}, {ad.xz * r, ad.yz * r, ad.zz * r}};
            const double mx = m.means[3 * j], my = m.means[3 * j + 1], mz = m.means[3 * j + 2];
            const double rr[3] = {-A[1] / M0, -A[2] / M0, -A[3] / M0};              // mu - s = -S1/S0
            const double sx = mx - rr[0], sy = my - rr[1], sz = mz - rr[2];
            const double Jm[3][6] = {{0, sz, -sy, 1, 0, 0}, {-sz, 0, sx, 0, 1, 0}, {sy, -sx, 0, 0, 0, 1}};
            double PJ[3][6];
            for (int a = 0; a < 3; ++a)
                for (int b = 0; b < 6; ++b) PJ[a][b] = P[a][0] * Jm[0][b] + P[a][1] * Jm[1][b] + P[a][2] * Jm[2][b];
            int k = 0;
            for (int a = 0; a < 6; ++a)
                for (int b = a; b < 6; ++b) v[k++] += M0 * (Jm[0][a] * PJ[0][b] + Jm[1][a] * PJ[1][b] + Jm[2][a] * PJ[2][b]);
            const double Pr[3] = {P[0][0] * rr[0] + P[0][1] * rr[1] + P[0][2] * rr[2], P[1][0] * rr[0] + P[1][1] * rr[1] + P[1][2] * rr[2],
                                  P[2][0] * rr[0] + P[2][1] * rr[1] + P[2][2] * rr[2]};
            for (int a = 0; a < 6; ++a) v[21 + a] += M0 * (Jm[0][a] * Pr[0] + Jm[1][a] * Pr[1] + Jm[2][a] * Pr[2]);
            v[27] += M0 * (rr[0] * Pr[0] + rr[1] * Pr[1] + rr[2] * Pr[2]);
        }
        block_reduce<kSys>(v, sm, tid, nth);
        if (tid == 0) twist_finish(sm, Rt, q_hist, qstate, ctrl, tol);
    } else {
        double v[kPro];
        for (int k = 0; k < kPro; ++k) v[k] = 0.0;
        for (int j = tid; j < m.J; j += nth) {
            const double* A = acc + kAccHdr + (size_t)j * kMom;
            const double w = A[0];
            if (w < f32eps) continue;
            const double mu[3] = {m.means[3 * j], m.means[3 * j + 1], m.means[3 * j + 2]};
            const double sc[3] = {mu[0] + A[1] / w, mu[1] + A[2] / w, mu[2] + A[3] / w};
            v[0] += w;
            for (int a = 0; a < 3; ++a) {
                v[1 + a] += w * sc[a];
                v[4 + a] += w * mu[a];
                for (int b = 0; b < 3; ++b) v[7 + 3 * a + b] += w * sc[a] * mu[b];
            }
            v[16] += w * (sc[0] * sc[0] + sc[1] * sc[1] + sc[2] * sc[2]);
            v[17] += w * (mu[0] * mu[0] + mu[1] * mu[1] + mu[2] * mu[2]);
        }
        block_reduce<kPro>(v, sm, tid, nth);
        if (tid == 0) procrustes_finish(sm, Rt, q_hist, qstate, ctrl, tol);
    }
}

__global__ void zero_doubles_kernel(double* p, size_t n, const int* ctrl) {
    if (ctrl && ctrl[0]) return;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = 0.0;
}

// pos = (-x,-y,-z)/scene_scale, 1 ; colour = rgb + 0.3, 1   (gmm_kernels.cu:71-93)
__global__ void fill_vbo_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ z,
                                int64_t n, int64_t offset, float* __restrict__ vbo_pos, float* __restrict__ vbo_col,
                                float c_scale, float r, float g, float b) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t o = 4 * (offset + i);
    if (vbo_pos) {
        reinterpret_cast<float4*>(vbo_pos)[offset + i] = make_float4(x[i] * c_scale, y[i] * c_scale, z[i] * c_scale, 1.0f);
    }
    if (vbo_col) {
        vbo_col[o + 0] = r + 0.3f;
        vbo_col[o + 1] = g + 0.3f;
        vbo_col[o + 2] = b + 0.3f;
        vbo_col[o + 3] = 1.0f;
    }
}

// ------------------------------------------------------------------------------------------
cudaError_t launch_reg_estep(const float* tx, const float* ty, const float* tz, int n, const double* Rt, const TreeModel& t,
                             float lambda_c, double* racc, int want_m2, const int* ctrl, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    // HGMM_REG_ESTEP = 1 / 2: the first two generations (A/B switch); default: reg_estep3_kernel
    static const int gen = getenv("HGMM_REG_ESTEP") ? atoi(getenv("HGMM_REG_ESTEP")) : 3;
    if (gen >= 3 && t.L <= kRegMaxL) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((n + 127) / 128);
        cfg.blockDim = dim3(128);
        cfg.stream = s;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        return cudaLaunchKernelEx(&cfg, reg_estep3_kernel, tx, ty, tz, n, Rt, (const PackedComp*)t.packed, (const float*)t.cplx, t.L, lambda_c,
                                  racc, want_m2, ctrl);
    }
    const bool v2 = gen == 2;
    if (v2) {
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int need = (n + 31) / 32;                          // CTAs if every warp took exactly one group of 4 points
        const int grid = need < sms * 8 ? need : sms * 8;
        reg_estep2_kernel<<<grid, 256, 0, s>>>(tx, ty, tz, n, Rt, t.packed, t.cplx, t.L, lambda_c, racc, want_m2, ctrl);
        return cudaGetLastError();
    }
    reg_estep_kernel<<<(n + 255) / 256, 256, 0, s>>>(tx, ty, tz, n, Rt, t.packed, t.cplx, t.L, lambda_c, racc, want_m2, ctrl);
    return cudaGetLastError();
}

cudaError_t launch_reg_solve(const TreeModel& t, double* racc, int zero_after, int solver, double* Rt, double* q_hist,
                             double* qstate, int* ctrl, float tol, const RegXchgView* xvp, cudaStream_t s) {
    RegXchgView xv = {};
    xv.nranks = 1;
    if (xvp) xv = *xvp;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(1);
    cfg.blockDim = dim3(512);
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, reg_solve_kernel, t, racc, zero_after, solver, Rt, q_hist, qstate, ctrl, tol, xv);
}

void launch_transform_soa(const float* tx, const float* ty, const float* tz, int n, const double* Rt, float* ox, float* oy, float* oz,
                          const int* ctrl, cudaStream_t s) {
    if (n <= 0) return;
    transform_soa_kernel<<<(n + 255) / 256, 256, 0, s>>>(tx, ty, tz, n, Rt, ox, oy, oz, ctrl);
}

cudaError_t launch_reg_flat_solve(const FlatModel& m, const double* acc, int solver, double* Rt, double* q_hist, double* qstate,
                                  int* ctrl, float tol, cudaStream_t s) {
    reg_flat_solve_kernel<<<1, 512, 0, s>>>(m, acc, solver, Rt, q_hist, qstate, ctrl, tol);
    return cudaGetLastError();
}

void launch_zero_doubles(double* p, size_t n, const int* ctrl, cudaStream_t s) {
    if (n == 0) return;
    zero_doubles_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(p, n, ctrl);
}

void launch_fill_vbo(const float* x, const float* y, const float* z, int64_t n, int64_t offset, float* vbo_pos, float* vbo_col,
                     float scene_scale, float r, float g, float b, cudaStream_t s) {
    if (n <= 0) return;
    fill_vbo_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(x, y, z, n, offset, vbo_pos, vbo_col, -1.0f / scene_scale, r, g, b);
}

}  // namespace hgmm
