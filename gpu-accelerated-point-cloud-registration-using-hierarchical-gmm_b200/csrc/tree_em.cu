// tree_em.cu -- per-level EM of the 8-ary hierarchical mixture and the device partition, sm_100a.
//
// Replaces (paths relative to the reference checkout):
//   gmmTreeEStepKernel / accumulateDevice / gaussianPdfKernel   src/python/hgmm/hgmm_gpu.py:284-358,387-411
//   gmmtreeMStepKernel / mlEstimator                            src/python/hgmm/hgmm_gpu.py:247-277,421-426
//   (guards follow the CPU file, the semantic authority: hgmm_cupy_cpu_working.py:99-119,162-191)
//   parentIdx <- currentIdx hand-off                            src/python/hgmm/hgmm_gpu.py:540
//
// Design.  The reference leaves points in place, gathers 8 children per
// thread by parent index and issues up to 104 contended fp32 global atomics per point.  Here the
// cloud is physically re-ordered at every level hand-off so that the points of one parent are
// contiguous (stable 8-way split = a radix partition on the 3 new key bits), a parent segment is
// cut into chunks of <= chunk_points points, and ONE WARP owns a chunk: the 8 children's packed
// parameters sit in 384 B of shared memory, each lane walks its points keeping all 8 x 10 centred
// moments in registers, a butterfly reduce-scatter folds the warp, and 80 fp64 atomics per chunk
// land in the level's moment block.
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace hgmm {

// ------------------------------------------------------------------------------------------
// complexity(cov) = lambda_min / trace  (hgmm_gpu.py:78-82), closed-form symmetric 3x3 eigenvalues
// ------------------------------------------------------------------------------------------
__device__ double sym3_complexity(const Sym3& s) {
    const double tr = s.xx + s.yy + s.zz;
    const double p1 = s.xy * s.xy + s.xz * s.xz + s.yz * s.yz;
    double lmin;
    if (p1 == 0.0) {
        lmin = fmin(s.xx, fmin(s.yy, s.zz));
    } else {
        const double q = tr / 3.0;
        const double a = s.xx - q, b = s.yy - q, c = s.zz - q;
        const double p2 = a * a + b * b + c * c + 2.0 * p1;
        const double p = sqrt(p2 / 6.0);
        const double ip = 1.0 / p;
        Sym3 B{a * ip, s.xy * ip, s.xz * ip, b * ip, s.yz * ip, c * ip};
        double r = 0.5 * sym3_det(B);
        r = fmin(1.0, fmax(-1.0, r));
        const double phi = acos(r) / 3.0;
        lmin = q + 2.0 * p * cos(phi + 2.0943951023931953);   // + 2 pi / 3
    }
    return lmin / tr;
}

__global__ void tree_init_kernel(TreeModel t, const float* __restrict__ init_means, float sig2) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= t.nt) return;
    // hgmm_gpu.py:487-490: pi = 1/8, mu = points[idxs[i]], cov = sig2 * I
    const float mx = init_means[3 * j], my = init_means[3 * j + 1], mz = init_means[3 * j + 2];
    t.pi[j] = 0.125f;
    t.mu[3 * j] = mx; t.mu[3 * j + 1] = my; t.mu[3 * j + 2] = mz;
    float* c = t.cov + 9 * j;
    c[0] = c[4] = c[8] = sig2;
    c[1] = c[2] = c[3] = c[5] = c[6] = c[7] = 0.f;
    Sym3 s{sig2, 0, 0, sig2, 0, sig2};
    t.cplx[j] = (float)(1.0 / 3.0);
    t.packed[j] = pack_full(log(0.125), mx, my, mz, s, false, 1e-15);
}

// pack + complexity for an externally supplied tree (hgmm_tree_set_model)
__global__ void tree_pack_all_kernel(TreeModel t) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= t.nt) return;
    const float* c = t.cov + 9 * j;
    Sym3 s{c[0], 0.5 * ((double)c[1] + c[3]), 0.5 * ((double)c[2] + c[6]), c[4], 0.5 * ((double)c[5] + c[7]), c[8]};
    const double w = t.pi[j];
    const float cx = (float)sym3_complexity(s);
    t.cplx[j] = cx;
    PackedComp pc = pack_full(w > 0.0 ? log(w) : -INFINITY, t.mu[3 * j], t.mu[3 * j + 1], t.mu[3 * j + 2], s, false, 1e-15);
    pc.pad0 = cx;                     // the registration descent reads the complexity with the parameters (registration.cu)
    t.packed[j] = pc;
}

// ------------------------------------------------------------------------------------------
// E-step: one warp per chunk
// ------------------------------------------------------------------------------------------
template <int N>
__device__ __forceinline__ void reduce_scatter_step(float* v, int offset, bool upper) {
#pragma unroll
    for (int i = 0; i < N / 2; ++i) {
        const float keep = upper ? v[i + N / 2] : v[i];
        const float send = upper ? v[i] : v[i + N / 2];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, offset);
    }
}

__global__ void __launch_bounds__(256) tree_estep_kernel(const float* __restrict__ px, const float* __restrict__ py,
                                                         const float* __restrict__ pz,
                                                         const int* __restrict__ chunk_parent,
                                                         const int* __restrict__ chunk_start,
                                                         const int* __restrict__ chunk_len,
                                                         const int* __restrict__ n_chunks_dev,
                                                         const PackedComp* __restrict__ packed_level,
                                                         double* __restrict__ acc, uint8_t* __restrict__ slot,
                                                         const int* __restrict__ done_flag) {
    if (*done_flag) return;
    __shared__ __align__(16) float sch[8][96];
    __shared__ float s_part[8][8 * kMom];     // per-warp chunk sums (80 floats)
    __shared__ int s_parent[8];
    __shared__ double s_ll[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int chunk = blockIdx.x * 8 + warp;
    const int n_chunks = *n_chunks_dev;
    double ll = 0.0;
    int my_parent = -1;                       // -1: this warp has no chunk
    if (chunk < n_chunks) {
        const int p = chunk_parent[chunk];
        const int start = chunk_start[chunk];
        const int len = chunk_len[chunk];
        {
            const float* src = reinterpret_cast<const float*>(packed_level + 8 * (size_t)p);
            sch[warp][lane] = src[lane];
            sch[warp][lane + 32] = src[lane + 32];
            sch[warp][lane + 64] = src[lane + 64];
        }
        __syncwarp();
        const float4* c4 = reinterpret_cast<const float4*>(&sch[warp][0]);
        float v[8 * kMom];
#pragma unroll
        for (int i = 0; i < 8 * kMom; ++i) v[i] = 0.f;

        for (int r = lane; r < len; r += 32) {
            const int i = start + r;
            const float x = px[i], y = py[i], z = pz[i];
            float q[8];
            float m = kNegBig;
            int best = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float4 p0 = c4[3 * k], p1 = c4[3 * k + 1];
                const float2 p2 = *reinterpret_cast<const float2*>(c4 + 3 * k + 2);
                float dx, dy, dz;
                q[k] = quad_q2(p0, p1, p2, x, y, z, dx, dy, dz);
                if (q[k] > m) { m = q[k]; best = k; }        // first maximum, like np.argmax
            }
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) s += ex2f(q[k] - m);
            const float lse2 = m + lg2f(s);
            // hgmm_cupy_cpu_working.py:174-178: gamma = gamma/den if den > eps else zeros
            const bool alive = lse2 > kLog2Eps15;
            slot[i] = (uint8_t)(alive ? best : 0);
            ll += (double)(kLn2 * fmaxf(alive ? lse2 : kLog2Eps15, kLog2Eps15));
            if (alive) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    float gam = ex2f(q[k] - lse2);
                    gam = (gam < 1e-15f) ? 0.f : gam;         // accumulate() skips gamma < eps (:100-101)
                    const float4 p0 = c4[3 * k];
                    const float dx = x - p0.x, dy = y - p0.y, dz = z - p0.z;
                    const float gx = gam * dx, gy = gam * dy, gz = gam * dz;
                    float* a = v + k * kMom;
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
        }
        // butterfly reduce-scatter 80 -> 5 values per lane, then pair all-reduce
        reduce_scatter_step<80>(v, 16, (lane & 16) != 0);
        reduce_scatter_step<40>(v, 8, (lane & 8) != 0);
        reduce_scatter_step<20>(v, 4, (lane & 4) != 0);
        reduce_scatter_step<10>(v, 2, (lane & 2) != 0);
#pragma unroll
        for (int i = 0; i < 5; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i], 1);
        my_parent = p;
        if ((lane & 1) == 0) {
            const int idx0 = ((lane & 16) ? 40 : 0) + ((lane & 8) ? 20 : 0) + ((lane & 4) ? 10 : 0) + ((lane & 2) ? 5 : 0);
#pragma unroll
            for (int i = 0; i < 5; ++i) s_part[warp][idx0 + i] = v[i];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ll += __shfl_xor_sync(0xffffffffu, ll, o);
    }
    if (lane == 0) {
        s_ll[warp] = ll;
        s_parent[warp] = my_parent;
    }
    __syncthreads();
    // Fold the CTA's 8 chunk sums: runs of consecutive warps with the same parent are added in fp32 (<= 8 terms) and
    // leave as ONE set of <= 80 fp64 atomics -- at the top levels (few parents, thousands of chunks) this cuts the
    // same-address atomic traffic 8x.  Thread t < 80 owns moment slot t.
    if (tid < 8 * kMom) {
        int run_parent = -1;
        float run = 0.f;
        for (int w = 0; w < 8; ++w) {
            const int pw = s_parent[w];
            if (pw != run_parent) {
                if (run_parent >= 0 && run != 0.f) atomicAdd(acc + kAccHdr + (size_t)run_parent * (8 * kMom) + tid, (double)run);
                run_parent = pw;
                run = 0.f;
            }
            if (pw >= 0) run += s_part[w][tid];
        }
        if (run_parent >= 0 && run != 0.f) atomicAdd(acc + kAccHdr + (size_t)run_parent * (8 * kMom) + tid, (double)run);
    }
    if (tid == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += s_ll[w];
        if (t != 0.0) atomicAdd(acc, t);
    }
}

// ------------------------------------------------------------------------------------------
// E-step, packed-FP32 version (default): one warp per chunk, lane = (child pair cp = lane/8, point lane pl = lane%8).
// Each lane keeps ONE pair of children (FFMA2/FADD2/FMUL2 operands, 10 float2 moment accumulators, ~70 registers
// instead of ~210), the 8-way normalisation is two xor-shuffles (8, 16) across the four lanes that share a point,
// the fold over the eight point lanes is three xor-shuffles (1, 2, 4).  Same semantics as tree_estep_kernel.
// ------------------------------------------------------------------------------------------
typedef unsigned long long u64t;
__device__ __forceinline__ float2 t_ffma2(float2 a, float2 b, float2 c) {
    u64t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(*reinterpret_cast<u64t*>(&a)), "l"(*reinterpret_cast<u64t*>(&b)),
        "l"(*reinterpret_cast<u64t*>(&c)));
    return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 t_fadd2(float2 a, float2 b) {
    u64t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(*reinterpret_cast<u64t*>(&a)), "l"(*reinterpret_cast<u64t*>(&b)));
    return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 t_fmul2(float2 a, float2 b) {
    u64t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(*reinterpret_cast<u64t*>(&a)), "l"(*reinterpret_cast<u64t*>(&b)));
    return *reinterpret_cast<float2*>(&d);
}

// E-step of the eight chunks [8 * group, 8 * group + 8) by one CTA of 8 warps (the body shared by tree_estep2_kernel and the
// persistent tree_level_kernel); ends with the CTA-level fold into acc, callers separate consecutive groups with a barrier
__device__ __forceinline__ void tree_estep2_group(const float* __restrict__ px, const float* __restrict__ py,
                                                  const float* __restrict__ pz, const int* __restrict__ chunk_parent,
                                                  const int* __restrict__ chunk_start, const int* __restrict__ chunk_len,
                                                  int n_chunks, const PackedComp* __restrict__ packed_level,
                                                  double* __restrict__ acc, uint8_t* __restrict__ slot, int group,
                                                  float (*s_part)[8 * kMom], int* s_parent, double* s_ll) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cp = lane >> 3, pl = lane & 7;
    const int chunk = group * 8 + warp;
    double ll = 0.0;
    int my_parent = -1;
    if (chunk < n_chunks) {
        const int p = chunk_parent[chunk];
        const int start = chunk_start[chunk];
        const int len = chunk_len[chunk];
        // the lane's pair of children (2cp, 2cp+1) of parent p
        float2 nmx, nmy, nmz, c2, axx, ayy, azz, axy, axz, ayz;
        {
            const float4* a4 = reinterpret_cast<const float4*>(packed_level + 8 * (size_t)p + 2 * cp);
            const float4 a0 = __ldcg(a4), a1 = __ldcg(a4 + 1), a2 = __ldcg(a4 + 2);
            const float4 b0 = __ldcg(a4 + 3), b1 = __ldcg(a4 + 4), b2 = __ldcg(a4 + 5);
            nmx = make_float2(-a0.x, -b0.x); nmy = make_float2(-a0.y, -b0.y); nmz = make_float2(-a0.z, -b0.z);
            c2 = make_float2(a0.w, b0.w);
            axx = make_float2(a1.x, b1.x); ayy = make_float2(a1.y, b1.y); azz = make_float2(a1.z, b1.z);
            axy = make_float2(a1.w, b1.w); axz = make_float2(a2.x, b2.x); ayz = make_float2(a2.y, b2.y);
        }
        float2 a[kMom];
#pragma unroll
        for (int m = 0; m < kMom; ++m) a[m] = make_float2(0.f, 0.f);

        for (int rb = 0; rb < len; rb += 8) {                 // warp-uniform trip count: the shuffles need every lane
            const int r = rb + pl;
            const bool valid = r < len;
            const int i = start + (valid ? r : len - 1);
            const float x = px[i], y = py[i], z = pz[i];
            const float2 dx = t_fadd2(make_float2(x, x), nmx), dy = t_fadd2(make_float2(y, y), nmy), dz = t_fadd2(make_float2(z, z), nmz);
            float2 t0 = t_fmul2(axz, dz);
            t0 = t_ffma2(axy, dy, t0);
            t0 = t_ffma2(axx, dx, t0);
            float2 t1 = t_fmul2(ayz, dz);
            t1 = t_ffma2(ayy, dy, t1);
            const float2 t2 = t_fmul2(azz, dz);
            float2 q = t_ffma2(dz, t2, c2);
            q = t_ffma2(dy, t1, q);
            q = t_ffma2(dx, t0, q);
            float m = fmaxf(fmaxf(q.x, q.y), kNegBig);
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 8));
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 16));
            const float2 e = make_float2(ex2f(q.x - m), ex2f(q.y - m));
            float s = e.x + e.y;
            s += __shfl_xor_sync(0xffffffffu, s, 8);
            s += __shfl_xor_sync(0xffffffffu, s, 16);
            int best = (q.x == m) ? 2 * cp : ((q.y == m) ? 2 * cp + 1 : 99);      // first maximum, like np.argmax
            best = min(best, __shfl_xor_sync(0xffffffffu, best, 8));
            best = min(best, __shfl_xor_sync(0xffffffffu, best, 16));
            const float lse2 = m + lg2f(s);
            // hgmm_cupy_cpu_working.py:174-178: gamma = gamma/den if den > eps else zeros
            const bool alive = valid && (lse2 > kLog2Eps15);
            if (valid && cp == 0) {
                slot[i] = (uint8_t)((alive && best < 8) ? best : 0);
                ll += (double)(kLn2 * fmaxf(alive ? lse2 : kLog2Eps15, kLog2Eps15));
            }
            const float inv = alive ? __fdividef(1.0f, s) : 0.f;
            float2 gam = make_float2(e.x * inv, e.y * inv);
            gam.x = (gam.x < 1e-15f) ? 0.f : gam.x;            // accumulate() skips gamma < eps (:100-101)
            gam.y = (gam.y < 1e-15f) ? 0.f : gam.y;
            const float2 gx = t_fmul2(gam, dx), gy = t_fmul2(gam, dy), gz = t_fmul2(gam, dz);
            a[0] = t_fadd2(a[0], gam);
            a[1] = t_fadd2(a[1], gx);
            a[2] = t_fadd2(a[2], gy);
            a[3] = t_fadd2(a[3], gz);
            a[4] = t_ffma2(gx, dx, a[4]);
            a[5] = t_ffma2(gx, dy, a[5]);
            a[6] = t_ffma2(gx, dz, a[6]);
            a[7] = t_ffma2(gy, dy, a[7]);
            a[8] = t_ffma2(gy, dz, a[8]);
            a[9] = t_ffma2(gz, dz, a[9]);
        }
        // fold the eight point lanes
#pragma unroll
        for (int off = 1; off < 8; off <<= 1) {
#pragma unroll
            for (int m = 0; m < kMom; ++m) {
                a[m].x += __shfl_xor_sync(0xffffffffu, a[m].x, off);
                a[m].y += __shfl_xor_sync(0xffffffffu, a[m].y, off);
            }
        }
        my_parent = p;
        if (pl == 0) {
#pragma unroll
            for (int m = 0; m < kMom; ++m) {
                s_part[warp][(2 * cp) * kMom + m] = a[m].x;
                s_part[warp][(2 * cp + 1) * kMom + m] = a[m].y;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ll += __shfl_xor_sync(0xffffffffu, ll, o);
    }
    if (lane == 0) {
        s_ll[warp] = ll;
        s_parent[warp] = my_parent;
    }
    __syncthreads();
    if (tid < 8 * kMom) {                                   // same CTA-level fold as tree_estep_kernel
        int run_parent = -1;
        float run = 0.f;
        for (int w = 0; w < 8; ++w) {
            const int pw = s_parent[w];
            if (pw != run_parent) {
                if (run_parent >= 0 && run != 0.f) atomicAdd(acc + kAccHdr + (size_t)run_parent * (8 * kMom) + tid, (double)run);
                run_parent = pw;
                run = 0.f;
            }
            if (pw >= 0) run += s_part[w][tid];
        }
        if (run_parent >= 0 && run != 0.f) atomicAdd(acc + kAccHdr + (size_t)run_parent * (8 * kMom) + tid, (double)run);
    }
    if (tid == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += s_ll[w];
        if (t != 0.0) atomicAdd(acc, t);
    }
}

__global__ void __launch_bounds__(256, 3) tree_estep2_kernel(const float* __restrict__ px, const float* __restrict__ py,
                                                             const float* __restrict__ pz,
                                                             const int* __restrict__ chunk_parent,
                                                             const int* __restrict__ chunk_start,
                                                             const int* __restrict__ chunk_len,
                                                             const int* __restrict__ n_chunks_dev,
                                                             const PackedComp* __restrict__ packed_level,
                                                             double* __restrict__ acc, uint8_t* __restrict__ slot,
                                                             const int* __restrict__ done_flag) {
    // programmatic dependent launch (the build is launch-bound: ~3 us of work per EM iteration): this grid may be scheduled
    // while the M-step of the previous iteration drains; everything that kernel wrote is read after the wait, from L2
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (__ldcg(done_flag)) return;
    __shared__ float s_part[8][8 * kMom];
    __shared__ int s_parent[8];
    __shared__ double s_ll[8];
    tree_estep2_group(px, py, pz, chunk_parent, chunk_start, chunk_len, __ldcg(n_chunks_dev), packed_level, acc, slot, blockIdx.x, s_part,
                      s_parent, s_ll);
}

// ------------------------------------------------------------------------------------------
// M-step: one thread per node of the level
// ------------------------------------------------------------------------------------------
// done_at[it] is written by iteration it-1 only.  merge_converge != 0 (fast log-likelihood mode): thread 0 of block 0
// also applies the |q - prevQ| < ls rule to acc[0] (complete: the E-step has finished) and arms done_at[it+1].
__device__ void tree_mstep_node(const TreeModel& t, int lb, int local, double* __restrict__ acc, double n_total, float ld);

__global__ void tree_mstep_kernel(TreeModel t, int lb, int count, double* __restrict__ acc, double n_total, float ld,
                                  int* __restrict__ ctrl, int* __restrict__ done_at, int it, int merge_converge,
                                  double* __restrict__ qstate, float ls, int max_iters, volatile int* prog) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const bool done = __ldcg(done_at + it) != 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (done) {
            done_at[it + 1] = 1;
            if (prog && merge_converge) prog[0] = it + 1;
        } else if (merge_converge) {
            const double q = __ldcg(acc);                     // L2 loads: written by the previous kernels of the PDL chain
            acc[0] = 0.0;
            const int n_it = __ldcg(ctrl + 1) + 1;
            ctrl[1] = n_it;
            qstate[1] = q;
            const bool conv = fabs(q - __ldcg(qstate)) < (double)ls || n_it >= max_iters;
            qstate[0] = q;
            ctrl[0] = conv ? 1 : 0;
            done_at[it + 1] = conv ? 1 : 0;
            if (prog) {                                   // host-mapped progress words: [1] converged, then [0] iterations retired
                prog[1] = conv ? 1 : 0;
                __threadfence_system();
                prog[0] = it + 1;
            }
        }
    }
    if (done) return;
    const int local = blockIdx.x * blockDim.x + threadIdx.x;
    if (local >= count) return;
    tree_mstep_node(t, lb, local, acc, n_total, ld);
}

// centred moments A of node `local` of the level -> parameters (mlEstimator, hgmm_cupy_cpu_working.py:109-119: blank node if
// M0 < ld), written to the model and returned packed
__device__ __forceinline__ PackedComp tree_mstep_apply(const TreeModel& t, int lb, int local, const double* A, double n_total, float ld) {
    const int j = lb + local;
    const double M0 = A[0];
    PackedComp pc;
    if (M0 >= (double)ld) {
        const double r = 1.0 / M0;
        const double dx = A[1] * r, dy = A[2] * r, dz = A[3] * r;
        Sym3 s{A[4] * r - dx * dx, A[5] * r - dx * dy, A[6] * r - dx * dz, A[7] * r - dy * dy, A[8] * r - dy * dz,
               A[9] * r - dz * dz};
        const float fx = (float)((double)t.mu[3 * j] + dx), fy = (float)((double)t.mu[3 * j + 1] + dy),
                    fz = (float)((double)t.mu[3 * j + 2] + dz);
        const double w = M0 / n_total;
        t.pi[j] = (float)w;
        t.mu[3 * j] = fx; t.mu[3 * j + 1] = fy; t.mu[3 * j + 2] = fz;
        float* c = t.cov + 9 * j;
        c[0] = (float)s.xx; c[1] = c[3] = (float)s.xy; c[2] = c[6] = (float)s.xz;
        c[4] = (float)s.yy; c[5] = c[7] = (float)s.yz; c[8] = (float)s.zz;
        pc = pack_full(log(w), fx, fy, fz, s, false, 1e-15);
    } else {
        t.pi[j] = 0.f;
        t.mu[3 * j] = t.mu[3 * j + 1] = t.mu[3 * j + 2] = 0.f;
        float* c = t.cov + 9 * j;
        c[0] = c[4] = c[8] = 1.f;
        c[1] = c[2] = c[3] = c[5] = c[6] = c[7] = 0.f;
        pc = pack_full(-INFINITY, 0, 0, 0, Sym3{1, 0, 0, 1, 0, 1}, false, 1e-15);
    }
    t.packed[j] = pc;
    return pc;
}

// moments of node `local` of the level (in acc) -> parameters; moments cleared
__device__ void tree_mstep_node(const TreeModel& t, int lb, int local, double* __restrict__ acc, double n_total, float ld) {
    double* Ag = acc + kAccHdr + (size_t)local * kMom;
    double A[kMom];
#pragma unroll
    for (int k = 0; k < kMom; ++k) A[k] = __ldcg(Ag + k);
    tree_mstep_apply(t, lb, local, A, n_total, ld);
#pragma unroll
    for (int k = 0; k < kMom; ++k) Ag[k] = 0.0;
}

#include "tree_level.cuh"

// |q - prevQ| < ls with prevQ = 0 at level start (hgmm_gpu.py:520,533-535).  qstate: [0] prevQ, [1] last q.
__global__ void tree_converge_kernel(double* __restrict__ acc, int* __restrict__ ctrl, int* __restrict__ done_at, int it,
                                     double* __restrict__ qstate, float ls, int max_iters, volatile int* prog) {
    if (done_at[it]) {
        done_at[it + 1] = 1;
        if (prog) prog[0] = it + 1;
        return;
    }
    const double q = acc[0];
    acc[0] = 0.0;
    const int n_it = ctrl[1] + 1;
    ctrl[1] = n_it;
    qstate[1] = q;
    const bool conv = fabs(q - qstate[0]) < (double)ls || n_it >= max_iters;
    qstate[0] = q;
    ctrl[0] = conv ? 1 : 0;
    done_at[it + 1] = conv ? 1 : 0;
    if (prog) {
        prog[1] = conv ? 1 : 0;
        __threadfence_system();
        prog[0] = it + 1;
    }
}

__global__ void tree_zero_ll_kernel(double* __restrict__ acc, const int* __restrict__ done_flag) {
    if (*done_flag) return;
    acc[0] = 0.0;
}

// complexity of every node (registration pruning); run once after the build instead of inside every M-step
__global__ void tree_cplx_kernel(TreeModel t) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= t.nt) return;
    const float* c = t.cov + 9 * j;
    Sym3 s{c[0], 0.5 * ((double)c[1] + c[3]), 0.5 * ((double)c[2] + c[6]), c[4], 0.5 * ((double)c[5] + c[7]), c[8]};
    const float cx = (float)sym3_complexity(s);
    t.cplx[j] = cx;
    t.packed[j].pad0 = cx;            // the registration descent reads the complexity with the parameters (registration.cu)
}

// adaptive build: which nodes of a finished level are terminal (include/hgmm.h, hgmm_tree_config.prune_*)
__global__ void tree_prune_kernel(TreeModel t, int lb, int count, double n_total, float lambda_c, float min_points,
                                  uint8_t* __restrict__ term) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count) return;
    const int g = lb + j;
    const double w = t.pi[g];
    bool terminal = !(w > 0.0) || (min_points > 0.f && w * n_total < (double)min_points);
    if (!terminal && lambda_c > 0.f) {
        const float* c = t.cov + 9 * (size_t)g;
        Sym3 s{c[0], 0.5 * ((double)c[1] + c[3]), 0.5 * ((double)c[2] + c[6]), c[4], 0.5 * ((double)c[5] + c[7]), c[8]};
        terminal = sym3_complexity(s) <= (double)lambda_c;
    }
    term[j] = terminal ? 1 : 0;
}

// current[perm[i]] = level base + 8 * parent + slot  (hgmm_gpu.py:411)
__global__ void tree_current_kernel(const int* __restrict__ perm, const int* __restrict__ pnode,
                                    const uint8_t* __restrict__ slot, int n, int lb, int64_t* __restrict__ current) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) current[perm[i]] = (int64_t)lb + 8 * (int64_t)pnode[i] + slot[i];
}

__global__ void iota_kernel(int* p, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = i;
}

// ------------------------------------------------------------------------------------------
// partition: stable 8-way split of every parent segment by `slot`
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) part_count_kernel(const uint8_t* __restrict__ slot, int n,
                                                          uint16_t* __restrict__ group_off, uint32_t* __restrict__ tile_cnt) {
    __shared__ uint32_t cnts[32][8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int i = blockIdx.x * 1024 + tid;
    const int s = (i < n) ? (int)slot[i] : 255;
    uint32_t mine = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const uint32_t b = __ballot_sync(0xffffffffu, s == k);
        if (lane == k) mine = __popc(b);
    }
    if (lane < 8) cnts[warp][lane] = mine;
    __syncthreads();
    if (tid < 8) {
        uint32_t run = 0;
        for (int w = 0; w < 32; ++w) {
            const uint32_t c = cnts[w][tid];
            group_off[((size_t)blockIdx.x * 32 + w) * 8 + tid] = (uint16_t)run;
            run += c;
        }
        tile_cnt[(size_t)blockIdx.x * 8 + tid] = run;
    }
}

// exclusive scan of tile_cnt over tiles for each of the 8 slots; warp k owns slot k
__global__ void __launch_bounds__(256) part_scan_kernel(const uint32_t* __restrict__ tile_cnt, int n_tiles,
                                                        uint32_t* __restrict__ tile_off) {
    const int lane = threadIdx.x & 31, k = threadIdx.x >> 5;
    uint32_t run = 0;
    for (int b = 0; b < n_tiles; b += 32) {
        const int t = b + lane;
        const uint32_t c = (t < n_tiles) ? tile_cnt[(size_t)t * 8 + k] : 0u;
        uint32_t inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += u;
        }
        if (t < n_tiles) tile_off[(size_t)t * 8 + k] = run + inc - c;
        run += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) tile_off[(size_t)n_tiles * 8 + k] = run;
}

// prefix counts G_k(pos) at every segment start (one thread per (parent, slot))
__global__ void part_parents_kernel(const uint8_t* __restrict__ slot, int n, int n_tiles, const int* __restrict__ seg_start,
                                    int n_parents, const uint16_t* __restrict__ group_off,
                                    const uint32_t* __restrict__ tile_off, uint32_t* __restrict__ seg_base) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (n_parents + 1) * 8) return;
    const int p = idx >> 3, k = idx & 7;
    const int pos = seg_start[p];
    uint32_t G;
    if (pos >= n) {
        G = tile_off[(size_t)n_tiles * 8 + k];
    } else {
        const int tile = pos >> 10, grp = pos >> 5, r = pos & 31;
        G = tile_off[(size_t)tile * 8 + k] + group_off[(size_t)grp * 8 + k];
        const uint8_t* sp = slot + (size_t)grp * 32;
        for (int u = 0; u < r; ++u) G += (sp[u] == k) ? 1u : 0u;
    }
    seg_base[idx] = G;
}

__global__ void part_children_kernel(const int* __restrict__ seg_start, int n_parents, int n,
                                     const uint32_t* __restrict__ seg_base, int* __restrict__ new_seg_start,
                                     int* __restrict__ chunk_cnt, int chunk_points) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_parents) return;
    int start = seg_start[p];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int c = (int)(seg_base[(size_t)(p + 1) * 8 + k] - seg_base[(size_t)p * 8 + k]);
        new_seg_start[8 * p + k] = start;
        chunk_cnt[8 * p + k] = (c + chunk_points - 1) / chunk_points;
        start += c;
    }
    if (p == n_parents - 1) new_seg_start[8 * n_parents] = n;
}

__global__ void part_scatter_kernel(const float* __restrict__ sx, const float* __restrict__ sy, const float* __restrict__ sz,
                                    const int* __restrict__ sperm, const int* __restrict__ spnode,
                                    const uint8_t* __restrict__ slot, int n, const uint16_t* __restrict__ group_off,
                                    const uint32_t* __restrict__ tile_off, const uint32_t* __restrict__ seg_base,
                                    const int* __restrict__ new_seg_start, float* __restrict__ dx, float* __restrict__ dy,
                                    float* __restrict__ dz, int* __restrict__ dperm, int* __restrict__ dpnode) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int s = (i < n) ? (int)slot[i] : 255;
    const uint32_t same = __match_any_sync(0xffffffffu, s);
    if (i >= n) return;
    const uint32_t rank_in_group = __popc(same & ((1u << lane) - 1u));
    const int p = spnode[i];
    const uint32_t G = tile_off[(size_t)(i >> 10) * 8 + s] + group_off[(size_t)(i >> 5) * 8 + s] + rank_in_group;
    const int dest = new_seg_start[8 * p + s] + (int)(G - seg_base[(size_t)p * 8 + s]);
    dx[dest] = sx[i];
    dy[dest] = sy[i];
    dz[dest] = sz[i];
    dperm[dest] = sperm[i];
    dpnode[dest] = 8 * p + s;
}

// exclusive scan of chunk_cnt (single CTA), total -> chunk_off[count] and *n_chunks_dev
__global__ void __launch_bounds__(1024) chunk_scan_kernel(const int* __restrict__ chunk_cnt, int count,
                                                          int* __restrict__ chunk_off, int* __restrict__ n_chunks_dev) {
    __shared__ int warp_tot[32];
    __shared__ int carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int b = 0; b < count; b += 1024) {
        const int i = b + tid;
        const int c = (i < count) ? chunk_cnt[i] : 0;
        int inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += u;
        }
        if (lane == 31) warp_tot[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            int w = warp_tot[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += u;
            }
            warp_tot[lane] = w;       // inclusive over warps
        }
        __syncthreads();
        const int base = carry + (warp > 0 ? warp_tot[warp - 1] : 0);
        if (i < count) chunk_off[i] = base + inc - c;
        __syncthreads();
        if (tid == 0) carry += warp_tot[31];
        __syncthreads();
    }
    if (tid == 0) {
        chunk_off[count] = carry;
        *n_chunks_dev = carry;
    }
}

__global__ void chunk_fill_kernel(const int* __restrict__ new_seg_start, const int* __restrict__ chunk_cnt,
                                  const int* __restrict__ chunk_off, int count, int chunk_points,
                                  int* __restrict__ chunk_parent, int* __restrict__ chunk_start, int* __restrict__ chunk_len) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= count) return;
    const int s0 = new_seg_start[c], s1 = new_seg_start[c + 1];
    const int o = chunk_off[c], k = chunk_cnt[c];
    for (int q = 0; q < k; ++q) {
        const int st = s0 + q * chunk_points;
        chunk_parent[o + q] = c;
        chunk_start[o + q] = st;
        chunk_len[o + q] = min(chunk_points, s1 - st);
    }
}

__global__ void root_chunks_kernel(int n, int chunk_points, int* __restrict__ chunk_parent, int* __restrict__ chunk_start,
                                   int* __restrict__ chunk_len, int* __restrict__ n_chunks_dev, int* __restrict__ seg_start) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int total = (n + chunk_points - 1) / chunk_points;
    if (c == 0) {
        *n_chunks_dev = total;
        seg_start[0] = 0;
        seg_start[1] = n;
    }
    if (c >= total) return;
    chunk_parent[c] = 0;
    chunk_start[c] = c * chunk_points;
    chunk_len[c] = min(chunk_points, n - c * chunk_points);
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
static inline int level_base(int l) {        // 8(8^l - 1)/7
    int64_t p = 1;
    for (int i = 0; i < l; ++i) p *= 8;
    return (int)(8 * (p - 1) / 7);
}
static inline int level_count(int l) {       // 8^(l+1)
    int64_t p = 8;
    for (int i = 0; i < l; ++i) p *= 8;
    return (int)p;
}

void launch_tree_init(const TreeModel& t, const float* init_means, float sig2, cudaStream_t s) {
    tree_init_kernel<<<(t.nt + 127) / 128, 128, 0, s>>>(t, init_means, sig2);
}
void launch_tree_pack_all(const TreeModel& t, cudaStream_t s) {
    tree_pack_all_kernel<<<(t.nt + 127) / 128, 128, 0, s>>>(t);
}

cudaError_t launch_tree_estep(const TreeWork& w, const TreeModel& t, int level, double* acc, int n_chunks_bound,
                              const int* n_chunks_dev, const int* done_flag, int scalar_variant, cudaStream_t s) {
    if (n_chunks_bound <= 0) return cudaSuccess;
    const int grid = (n_chunks_bound + 7) / 8;
    if (!scalar_variant) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(256);
        cfg.stream = s;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        const PackedComp* pl = t.packed + level_base(level);
        return cudaLaunchKernelEx(&cfg, tree_estep2_kernel, (const float*)w.x, (const float*)w.y, (const float*)w.z,
                                  (const int*)w.chunk_parent, (const int*)w.chunk_start, (const int*)w.chunk_len, n_chunks_dev, pl, acc,
                                  w.slot, done_flag);
    }
    tree_estep_kernel<<<grid, 256, 0, s>>>(w.x, w.y, w.z, w.chunk_parent, w.chunk_start, w.chunk_len, n_chunks_dev,
                                           t.packed + level_base(level), acc, w.slot, done_flag);
    return cudaGetLastError();
}

void launch_tree_mstep(const TreeModel& t, int level, double* acc, double n_total, float ld, int* ctrl, int* done_at, int it,
                       int merge_converge, double* qstate, float ls, int max_iters, int* prog, cudaStream_t s) {
    const int cnt = level_count(level);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((cnt + 127) / 128);
    cfg.blockDim = dim3(128);
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, tree_mstep_kernel, t, level_base(level), cnt, acc, n_total, ld, ctrl, done_at, it, merge_converge, qstate,
                       ls, max_iters, (volatile int*)prog);
}

// shared-memory plan of tree_level_kernel for a rank holding n points: the warps' parameter blocks, the fold stage, the chunk
// descriptors, and as many resident points as the rest holds
void tree_level_plan(int n, int chunk_points, int level, int num_sms, int smem_optin, int* pt_cap, int* chunk_cap, int* stage_cap,
                     size_t* smem_bytes) {
    const int per_cta = (n + num_sms - 1) / num_sms + chunk_points + 64;          // balanced by points, whole chunks
    const long long parents = level == 0 ? 1 : (long long)level_count(level - 1);
    const int sc = 96;                                                            // parked fold results (320 B each)
    long long cc = per_cta / chunk_points + (parents + num_sms - 1) / num_sms + 16;      // chunk descriptors kept on chip
    if (cc > 4096) cc = 4096;
    const long long fixed = (long long)(kTlThreads / 32) * (320 + 1280) + (long long)sc * 324 + cc * 12;     // + the transpose buffer
    const long long room = (long long)smem_optin - fixed - 4096;                  // static shared memory of the kernel + slack
    int pc = per_cta;
    if ((long long)pc * 12 > room) pc = 0;                                        // does not fit: stream the points from L2 / HBM
    *pt_cap = pc;
    *chunk_cap = (int)cc;
    *stage_cap = sc;
    *smem_bytes = (size_t)fixed + (size_t)pc * 12;
    // test switch (read per call, tests/test_gpu_parity.py): force the kernel's overflow paths, which configs of ordinary size
    // never reach -- bit 0: points streamed from L2 instead of shared memory; bit 1: chunk descriptors read from global memory;
    // bit 2: a two-slot fold stage (every further fold goes straight to L2 atomics)
    if (const char* e = getenv("HGMM_TREE_FORCE")) {
        const int f = atoi(e);
        if (f & 1) { *smem_bytes -= (size_t)*pt_cap * 12; *pt_cap = 0; }
        if (f & 2) *chunk_cap = 1;
        if (f & 4) *stage_cap = 2;
    }
}

// one cooperative launch for the whole EM loop of a level (tree_level.cuh)
cudaError_t launch_tree_level(const TreeWork& w, const TreeModel& t, int level, int n, double* acc, size_t acc_stride,
                              const int* n_chunks_dev, double n_total, float ld, float ls, int max_iters, int* ctrl, double* qstate,
                              unsigned* gbar, int chunk_points, const TreeXchgHost& xh, int num_sms, long long* prof,
                              const uint8_t* term, cudaStream_t s) {
    TreeLevelArgs a;
    a.prof = prof;
    a.term = term;
    a.px = w.x; a.py = w.y; a.pz = w.z;
    a.chunk_parent = w.chunk_parent; a.chunk_start = w.chunk_start; a.chunk_len = w.chunk_len; a.n_chunks_dev = n_chunks_dev;
    a.slot = w.slot;
    a.t = t;
    a.lb = level_base(level); a.cnt = level_count(level); a.n = n;
    a.acc = acc; a.acc_stride = acc_stride;
    a.n_total = n_total; a.ld = ld; a.ls = ls; a.max_iters = max_iters;
    a.ctrl = ctrl; a.qstate = qstate; a.gbar = gbar;
    for (int r = 0; r < kXchgMaxRanks; ++r) {
        a.x.pk[r] = reinterpret_cast<unsigned long long*>(xh.pk[r]);
        a.x.fin[r] = reinterpret_cast<unsigned long long*>(xh.fin[r]);
        a.x.mom[r] = reinterpret_cast<uint4*>(xh.mom[r]);
        a.x.ll[r] = reinterpret_cast<uint4*>(xh.ll[r]);
    }
    a.x.rank = xh.rank; a.x.nranks = xh.nranks; a.x.base = xh.base; a.x.mom_cap = xh.mom_cap; a.x.timeout_ns = xh.timeout_ns;
    const int smem_optin = device_smem_optin();
    size_t smem = 0;
    tree_level_plan(n, chunk_points, level, num_sms, smem_optin, &a.pt_cap, &a.chunk_cap, &a.stage_cap, &smem);
    static DeviceOnce once;
    if (once.first()) {
        cudaError_t e = cudaFuncSetAttribute(tree_level_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin - 4096);
        if (e != cudaSuccess) return e;
    }
    int occ = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, tree_level_kernel, kTlThreads, smem);
    if (e != cudaSuccess) return e;
    if (occ < 1) return cudaErrorNotSupported;
    void* args[] = {(void*)&a};
    return cudaLaunchCooperativeKernel((const void*)tree_level_kernel, dim3(num_sms), dim3(kTlThreads), args, smem, s);
}

void launch_tree_prune(const TreeModel& t, int level, double n_total, float lambda_c, float min_points, uint8_t* term, cudaStream_t s) {
    const int cnt = level_count(level);
    tree_prune_kernel<<<(cnt + 127) / 128, 128, 0, s>>>(t, level_base(level), cnt, n_total, lambda_c, min_points, term);
}

void launch_tree_converge(double* acc, int* ctrl, int* done_at, int it, double* qstate, float ls, int max_iters, int* prog,
                          cudaStream_t s) {
    tree_converge_kernel<<<1, 1, 0, s>>>(acc, ctrl, done_at, it, qstate, ls, max_iters, (volatile int*)prog);
}
void launch_tree_zero_ll(double* acc, const int* done_flag, cudaStream_t s) { tree_zero_ll_kernel<<<1, 1, 0, s>>>(acc, done_flag); }
void launch_tree_cplx(const TreeModel& t, cudaStream_t s) { tree_cplx_kernel<<<(t.nt + 127) / 128, 128, 0, s>>>(t); }

void launch_tree_current(const TreeWork& w, int n, int level, int64_t* current, cudaStream_t s) {
    if (n <= 0) return;
    tree_current_kernel<<<(n + 255) / 256, 256, 0, s>>>(w.perm, w.pnode, w.slot, n, level_base(level), current);
}
void launch_iota(int* p, int n, cudaStream_t s) {
    if (n > 0) iota_kernel<<<(n + 255) / 256, 256, 0, s>>>(p, n);
}

cudaError_t launch_root_chunks(TreeWork& w, int n, const PartitionScratch& ps, int chunk_points, int* n_chunks_dev,
                               cudaStream_t s) {
    const int total = (n + chunk_points - 1) / chunk_points;
    root_chunks_kernel<<<(total + 255) / 256 > 0 ? (total + 255) / 256 : 1, 256, 0, s>>>(
        n, chunk_points, w.chunk_parent, w.chunk_start, w.chunk_len, n_chunks_dev, ps.seg_start);
    return cudaGetLastError();
}

cudaError_t launch_partition(const TreeWork& src, TreeWork& dst, int n, int n_parents, PartitionScratch& ps, int chunk_points,
                             int* n_chunks_dev, cudaStream_t s) {
    const int n_tiles = (n + 1023) / 1024;
    part_count_kernel<<<n_tiles, 1024, 0, s>>>(src.slot, n, ps.group_off, ps.tile_cnt);
    part_scan_kernel<<<1, 256, 0, s>>>(ps.tile_cnt, n_tiles, ps.tile_off);
    part_parents_kernel<<<((n_parents + 1) * 8 + 255) / 256, 256, 0, s>>>(src.slot, n, n_tiles, ps.seg_start, n_parents,
                                                                        ps.group_off, ps.tile_off, ps.seg_base);
    part_children_kernel<<<(n_parents + 127) / 128, 128, 0, s>>>(ps.seg_start, n_parents, n, ps.seg_base, ps.new_seg_start,
                                                                 ps.chunk_cnt, chunk_points);
    part_scatter_kernel<<<(n + 255) / 256, 256, 0, s>>>(src.x, src.y, src.z, src.perm, src.pnode, src.slot, n, ps.group_off,
                                                        ps.tile_off, ps.seg_base, ps.new_seg_start, dst.x, dst.y, dst.z,
                                                        dst.perm, dst.pnode);
    const int count = 8 * n_parents;
    chunk_scan_kernel<<<1, 1024, 0, s>>>(ps.chunk_cnt, count, ps.chunk_off, n_chunks_dev);
    chunk_fill_kernel<<<(count + 127) / 128, 128, 0, s>>>(ps.new_seg_start, ps.chunk_cnt, ps.chunk_off, count, chunk_points,
                                                          dst.chunk_parent, dst.chunk_start, dst.chunk_len);
    // the split segments are the next level's parent segments
    int* t = ps.seg_start;
    ps.seg_start = ps.new_seg_start;
    ps.new_seg_start = t;
    return cudaGetLastError();
}

}  // namespace hgmm
