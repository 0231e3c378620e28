// tree_level.cuh -- the WHOLE EM loop of one tree level in one persistent cooperative kernel, single- and multi-GPU.
// Included by tree_em.cu (inside namespace hgmm, after the packed-FP32 helpers and mstep_params): one translation unit, no -rdc.
//
// Replaces, per level of buildGMMTree (src/python/hgmm/hgmm_gpu.py:519-540): the reference's launch sequence
//   gmmTreeEStepKernel -> gmmtreeMStepKernel -> logLikelihoodValue -> host test |q - prevQ| < ls   (one host sync per iteration)
// and this library's own two-kernel iteration (tree_estep2_kernel + tree_mstep_kernel, ~17 us per iteration for ~3 us of
// arithmetic, profiles/r01_launches_tree_L4_100k.txt) -- and, with several ranks, the ncclAllReduce between them.
//
// One CTA per SM (cooperative launch), kTlThreads threads.  What stays ON CHIP for the whole level:
//   * the CTA's slice of the permuted cloud (a contiguous run of whole chunks, balanced by points) lives in SHARED MEMORY --
//     the cloud is read from HBM once per level, not once per EM iteration (up to ~2.7 M points per GPU fit);
//   * the chunk descriptors of the slice;  * the warps' parameter blocks and a stage of parked fold results.
// One EM iteration:
//   E   every warp walks a contiguous run of chunks, lane = point: the 8 children of the chunk's parent sit in the warp's
//       shared-memory block (one broadcast LDS.128 feeds two packed FFMA2 operands), all 8 x 10 centred moments stay in
//       registers across consecutive chunks of one parent; a butterfly reduce-scatter folds the warp when the parent changes
//       and parks the 80 sums in a stage slot; the CTA adds runs of one parent and sends <= 80 fp64 atomics per (parent, CTA)
//       into acc[it & 1].
//   B   ONE grid barrier (monotonic arrival counter).
//   X   node j of the level is owned by exactly one thread of the grid (and, with R ranks, one rank: 32-node slices dealt
//       round-robin).  Non-owners PUSH their 10 partial sums into the owner's window (reduce-scatter over NVLink); the owner
//       adds the ranks' contributions in rank order, runs the M-step (mlEstimator, hgmm_cupy_cpu_working.py:109-119) and
//       PUBLISHES the node's packed parameters into every rank's window (all-gather).  No second barrier: the parameters
//       travel as self-validating 8-byte cells (fp32 payload, 32-bit tag = level base + iteration), and the next E-step
//       simply polls the cells of the children it needs -- dataflow instead of a barrier or a flag.
//   The stopping rule |q - prevQ| < ls (hgmm_gpu.py:533-535) is evaluated by every CTA of every rank from the same R
//   log-likelihood cells in the same order: identical decisions everywhere, no host round trip, no collective call.
// Buffering: acc and the moment / log-likelihood cells alternate by iteration parity (a rank can be at most one iteration
// ahead of another: its next push needs their previous publish); parameter cells are single-buffered (an owner publishes
// iteration it+1 only after every rank has finished E-step it).  On convergence the owners also publish (pi, mu, Sigma) so
// every rank leaves the level with the full replicated model.
// Every spin has a deadline (TreeXchg::timeout_ns, %globaltimer): on expiry ctrl[7] is set, every CTA of the grid leaves, and
// the host reports HGMM_ERR_NCCL instead of hanging.

constexpr int kTlThreads = 384;
constexpr int kPkWords = 10;          // mx my mz c2 axx ayy azz axy axz ayz  (the first ten floats of PackedComp)
constexpr int kFinWords = 10;         // pi, mu(3), Sigma (xx xy xz yy yz zz)

struct TreeXchg {
    unsigned long long* pk[kXchgMaxRanks];      // [cnt_cap * kPkWords]   8-byte cells of rank r's window (own = local pointer)
    unsigned long long* fin[kXchgMaxRanks];     // [cnt_cap * kFinWords]
    uint4* mom[kXchgMaxRanks];                  // [2][mom_cap]           16-byte cells (one fp64 each)
    uint4* ll[kXchgMaxRanks];                   // [2][kXchgMaxRanks]
    int rank, nranks;
    uint32_t base;                              // tag of the level's initial parameters; iteration it publishes base + it + 1
    unsigned long long mom_cap;                 // cells per parity
    unsigned long long timeout_ns;
};

struct TreeLevelArgs {
    const float *px, *py, *pz;                  // permuted cloud of this rank
    const int *chunk_parent, *chunk_start, *chunk_len, *n_chunks_dev;
    uint8_t* slot;
    const uint8_t* term;                        // adaptive build: [parents of this level] 1 = terminal (its points sit out), or null
    TreeModel t;
    int lb, cnt, n;
    double* acc;                                // [2][acc_stride], zero on entry
    unsigned long long acc_stride;
    double n_total;
    float ld, ls;
    int max_iters;
    int* ctrl;                                  // [0] converged, [1] iterations, [7] abort
    double* qstate;                             // [0] prevQ, [1] last q
    unsigned* gbar;                             // grid barrier arrival counter, zero on entry
    int pt_cap, chunk_cap, stage_cap;           // shared-memory capacities: points, chunk descriptors, parked fold results
    long long* prof;                            // optional [gridDim.x][8] SM-clock totals per phase (HGMM_TREE_PROF=1), else null
    TreeXchg x;
};

__device__ __forceinline__ unsigned long long tl_timer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void cell8_put(unsigned long long* cell, float v, uint32_t tag) {
    const unsigned long long u = ((unsigned long long)tag << 32) | (unsigned long long)__float_as_uint(v);
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(cell), "l"(u) : "memory");
}
__device__ __forceinline__ uint4 cell8_ld2(const unsigned long long* cell) {       // two adjacent cells: (v0, tag0, v1, tag1)
    uint4 c;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(c.x), "=r"(c.y), "=r"(c.z), "=r"(c.w) : "l"(cell) : "memory");
    return c;
}
__device__ __forceinline__ void cell16_put(uint4* cell, double v, uint32_t tag) {
    const unsigned long long u = (unsigned long long)__double_as_longlong(v);
    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(cell), "r"((uint32_t)u), "r"(tag) : "memory");
    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(reinterpret_cast<char*>(cell) + 8), "r"((uint32_t)(u >> 32)), "r"(tag)
                 : "memory");
}
__device__ __forceinline__ uint4 cell16_ld(const uint4* cell) {
    uint4 c;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(c.x), "=r"(c.y), "=r"(c.z), "=r"(c.w) : "l"(cell) : "memory");
    return c;
}
__device__ __forceinline__ int tl_abort(const int* ctrl) {
    int v;
    asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(ctrl + 7) : "memory");
    return v;
}
// spin bookkeeping: true when the caller must give up (deadline passed or another CTA aborted)
__device__ __forceinline__ bool tl_spin_fail(unsigned& spins, unsigned long long& t0, const TreeLevelArgs& a) {
    if ((++spins & 255u) != 0u) return false;
    if (tl_abort(a.ctrl)) return true;
    const unsigned long long now = tl_timer();
    if (t0 == 0ull) t0 = now;
    if (now - t0 > a.x.timeout_ns) {
        atomicExch(a.ctrl + 7, 1);
        return true;
    }
    return false;
}

// `count` fp64 values of consecutive 16-byte cells, all loaded together and re-polled until every word carries `tag`
template <int COUNT>
__device__ __forceinline__ bool cell16_get(const uint4* src, uint32_t tag, double* v, const TreeLevelArgs& a) {
    unsigned spins = 0;
    unsigned long long t0 = 0ull;
    for (;;) {
        uint4 c[COUNT];
#pragma unroll
        for (int k = 0; k < COUNT; ++k) c[k] = cell16_ld(src + k);
        bool all = true;
#pragma unroll
        for (int k = 0; k < COUNT; ++k) all = all && c[k].y == tag && c[k].w == tag;
        if (all) {
#pragma unroll
            for (int k = 0; k < COUNT; ++k) v[k] = __longlong_as_double((long long)(((unsigned long long)c[k].z << 32) | c[k].x));
            return true;
        }
        if (tl_spin_fail(spins, t0, a)) return false;
    }
}
// WORDS (even) fp32 values of consecutive 8-byte cells
template <int WORDS>
__device__ __forceinline__ bool cell8_get(const unsigned long long* src, uint32_t tag, float* v, const TreeLevelArgs& a) {
    unsigned spins = 0;
    unsigned long long t0 = 0ull;
    for (;;) {
        uint4 c[WORDS / 2];
#pragma unroll
        for (int k = 0; k < WORDS / 2; ++k) c[k] = cell8_ld2(src + 2 * k);
        bool all = true;
#pragma unroll
        for (int k = 0; k < WORDS / 2; ++k) all = all && c[k].y == tag && c[k].w == tag;
        if (all) {
#pragma unroll
            for (int k = 0; k < WORDS / 2; ++k) {
                v[2 * k] = __uint_as_float(c[k].x);
                v[2 * k + 1] = __uint_as_float(c[k].z);
            }
            return true;
        }
        if (tl_spin_fail(spins, t0, a)) return false;
    }
}

// grid barrier: arrival counter that only grows (zeroed by the host before the launch); false when the grid is aborting
__device__ __forceinline__ bool tl_grid_sync(const TreeLevelArgs& a, unsigned target, int* s_nslots) {
    __syncthreads();
    __shared__ int s_ok;
    if (threadIdx.x == 0) {
        *s_nslots = 0;                                     // the fold stage has been consumed: free for the next E-step
        __threadfence();                                   // this CTA's atomics / stores are visible before it counts as arrived
        atomicAdd(a.gbar, 1u);
        int ok = 1;
        unsigned spins = 0;
        for (;;) {
            unsigned v;
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(a.gbar) : "memory");
            if (v >= target) break;
            if ((++spins & 63u) == 0u && tl_abort(a.ctrl)) { ok = 0; break; }
        }
        if (tl_abort(a.ctrl)) ok = 0;                      // ctrl[7] is the single source of truth: the decision is CTA-uniform
        s_ok = ok;
    }
    __syncthreads();
    return s_ok != 0;
}

// first i with starts[i] >= key, searched by a whole warp: 32 probes per round trip (4 rounds for a million chunks instead of
// 20 dependent loads by one thread -- this sits on the critical path of every level's prologue)
__device__ __forceinline__ int tl_lower_bound_warp(const int* __restrict__ starts, int n, long long key, int lane) {
    int lo = 0, hi = n;                                   // answer in [lo, hi]
    while (hi - lo > 0) {
        const int span = hi - lo;
        const int step = (span + 32) / 33;                // probes at lo + (lane + 1) * step - 1, clipped
        const int pos = min(hi - 1, lo + (lane + 1) * step - 1);
        const bool below = (long long)__ldg(starts + pos) < key;
        const unsigned b = __ballot_sync(0xffffffffu, below);
        const int nb = __popc(b);                         // probes are monotone: the first nb are below the key
        if (nb == 32) {
            lo = min(hi, lo + 32 * step);
        } else {
            const int first_ge = min(hi - 1, lo + (nb + 1) * step - 1);
            hi = first_ge;
            if (nb > 0) lo = min(hi, lo + nb * step);
        }
        if (step == 1) {                                  // probes were consecutive: resolved
            if (nb < 32) { lo = hi; }
        }
    }
    return lo;
}

// ---- E-step building blocks: lane = point, the 8 children of the chunk's parent live in the warp's shared-memory block ----
// wp: float [4 child pairs][10 words][2 children] (means negated): one broadcast LDS.128 feeds two packed operands.
// The 80 cells of parent p's children are polled by the whole warp (lane l: cells 2l, 2l+1; lanes 0-7 also 64+2l, 65+2l).
__device__ __forceinline__ bool tl_load_params(const unsigned long long* cells, uint32_t tag, float* wp, int lane, const TreeLevelArgs& a) {
    unsigned spins = 0;
    unsigned long long t0 = 0ull;
    for (;;) {
        const uint4 c0 = cell8_ld2(cells + 2 * lane);
        const uint4 c1 = lane < 8 ? cell8_ld2(cells + 64 + 2 * lane) : make_uint4(0u, tag, 0u, tag);
        const bool ok = c0.y == tag && c0.w == tag && c1.y == tag && c1.w == tag;
        if (__all_sync(0xffffffffu, ok)) {
            __syncwarp();                                   // every lane is done with the previous parent's block
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                if (h >= 2 && lane >= 8) break;
                const int c = (h < 2 ? 2 * lane : 64 + 2 * lane) + (h & 1);
                const uint32_t bits = h == 0 ? c0.x : h == 1 ? c0.z : h == 2 ? c1.x : c1.z;
                const int k = c / kPkWords, w = c - k * kPkWords;
                const float v = __uint_as_float(bits);
                wp[(k >> 1) * 20 + w * 2 + (k & 1)] = w < 3 ? -v : v;
            }
            __syncwarp();
            return true;
        }
        const bool fail = tl_spin_fail(spins, t0, a);
        if (__any_sync(0xffffffffu, fail)) return false;
    }
}

template <int HALF>
__device__ __forceinline__ void tl_rs_step(float2* v, int off, bool upper) {      // butterfly reduce-scatter, float2 units
#pragma unroll
    for (int i = 0; i < HALF; ++i) {
        const float2 keep = upper ? v[i + HALF] : v[i];
        const float2 send = upper ? v[i] : v[i + HALF];
        v[i].x = keep.x + __shfl_xor_sync(0xffffffffu, send.x, off);
        v[i].y = keep.y + __shfl_xor_sync(0xffffffffu, send.y, off);
    }
}

// fold the warp's 8 x 10 moments of parent p (40 float2 per lane) and park the sums in a stage slot of the CTA (or, when the
// stage is full -- thousands of tiny segments per CTA -- add them straight into L2)
__device__ __forceinline__ void tl_fold_stage(float2* v, int p, int lane, float* stage, int* stage_parent, int stage_cap,
                                              int* s_nslots, double* accp) {
    tl_rs_step<20>(v, 16, (lane & 16) != 0);
    tl_rs_step<10>(v, 8, (lane & 8) != 0);
    tl_rs_step<5>(v, 4, (lane & 4) != 0);
#pragma unroll
    for (int off = 2; off > 0; off >>= 1) {
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            v[i].x += __shfl_xor_sync(0xffffffffu, v[i].x, off);
            v[i].y += __shfl_xor_sync(0xffffffffu, v[i].y, off);
        }
    }
    int slot = 0;
    if (lane == 0) slot = atomicAdd(s_nslots, 1);
    slot = __shfl_sync(0xffffffffu, slot, 0);
    if ((lane & 3) == 0) {
        const int base2 = ((lane & 16) ? 20 : 0) + ((lane & 8) ? 10 : 0) + ((lane & 4) ? 5 : 0);      // float2 index = pair * 10 + moment
        const int cp = base2 / 10, m0 = base2 - cp * 10;
        if (slot < stage_cap) {
            float* row = stage + (size_t)slot * (8 * kMom) + (2 * cp) * kMom + m0;
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                row[i] = v[i].x;
                row[kMom + i] = v[i].y;
            }
            if (lane == 0) stage_parent[slot] = p;
        } else {
            double* row = accp + kAccHdr + (size_t)p * (8 * kMom) + (2 * cp) * kMom + m0;
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                if (v[i].x != 0.f) atomicAdd(row + i, (double)v[i].x);
                if (v[i].y != 0.f) atomicAdd(row + kMom + i, (double)v[i].y);
            }
        }
    }
}

__global__ void __launch_bounds__(kTlThreads, 1) tree_level_kernel(const TreeLevelArgs a) {
    extern __shared__ __align__(16) unsigned char tl_smem[];
    constexpr int W = kTlThreads / 32;
    float* wpar = reinterpret_cast<float*>(tl_smem);                      // [W][80], 16-byte aligned rows
    float* pkst_all = wpar + W * 80;                                      // [W][320]: parameter transpose of the publishing warps
    float* stage = pkst_all + W * 32 * kPkWords;                          // [stage_cap][80]
    int* stage_parent = reinterpret_cast<int*>(stage + (size_t)a.stage_cap * 80);
    int* sc_parent = stage_parent + a.stage_cap;
    int* sc_start = sc_parent + a.chunk_cap;
    int* sc_len = sc_start + a.chunk_cap;
    float* sx = reinterpret_cast<float*>(sc_len + a.chunk_cap);
    float* sy = sx + a.pt_cap;
    float* sz = sy + a.pt_cap;
    __shared__ double s_ll[W];
    __shared__ int s_i[8];
    __shared__ double s_q;
    __shared__ int s_nslots;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x, b = blockIdx.x;
    const int R = a.x.nranks, me = a.x.rank;
    const int owned_cap = ((a.cnt + 31) / 32 + R - 1) / R * 32;           // nodes a rank can own (multiple of 32)
    if (tl_abort(a.ctrl)) return;                                         // an earlier level of this build already failed

    // ---------------- prologue: this CTA's run of chunks (balanced by points), its points and descriptors -> shared memory
    const int n_chunks = __ldcg(a.n_chunks_dev);
    if (warp < 2) {                                       // warps 0 and 1 search the two ends of this CTA's run concurrently
        const long long key = (long long)a.n * (b + warp) / G;
        const int c = (warp == 1 && b == G - 1) ? n_chunks : tl_lower_bound_warp(a.chunk_start, n_chunks, key, lane);
        if (lane == 0) {
            s_i[warp] = c;
            s_i[2 + warp] = c < n_chunks ? __ldg(a.chunk_start + c) : a.n;
            if (warp == 0) s_nslots = 0;
        }
    }
    __syncthreads();
    const int c0 = s_i[0], c1 = s_i[1], pt0 = s_i[2], pt1 = s_i[3];
    const int nc = c1 - c0;
    const bool resident = (pt1 - pt0) <= a.pt_cap;
    if (resident)
        for (int i = tid; i < pt1 - pt0; i += kTlThreads) {
            sx[i] = a.px[pt0 + i];
            sy[i] = a.py[pt0 + i];
            sz[i] = a.pz[pt0 + i];
        }
    for (int i = tid; i < nc && i < a.chunk_cap; i += kTlThreads) {
        sc_parent[i] = a.chunk_parent[c0 + i];
        sc_start[i] = a.chunk_start[c0 + i];
        sc_len[i] = a.chunk_len[c0 + i];
    }
    // seed this rank's own parameter cells from the (replicated) model: tag = base
    {
        const float* pkf = reinterpret_cast<const float*>(a.t.packed + a.lb);
        for (int idx = b * kTlThreads + tid; idx < a.cnt * kPkWords; idx += G * kTlThreads) {
            const int j = idx / kPkWords, w = idx - j * kPkWords;
            cell8_put(a.x.pk[me] + idx, pkf[(size_t)j * 12 + w], a.x.base);
        }
    }
    __syncthreads();
    // every warp walks a CONTIGUOUS run of the CTA's chunks: consecutive chunks of one parent stay in one set of registers
    const int per_warp = (nc + W - 1) / W;
    const int cw0 = warp * per_warp, cw1 = min(nc, cw0 + per_warp);

    double prev_q = 0.0;                                                  // hgmm_gpu.py:520
    bool aborted = false;
    __shared__ long long s_pf[8];                                         // thread 0: cycles in E | fold+flush | barrier | exchange+M | verdict
    if (tid == 0) {
        for (int k = 0; k < 7; ++k) s_pf[k] = 0;
        s_pf[7] = clock64();
    }
#define TL_STAMP(k) do { if (a.prof && tid == 0) { const long long now_ = clock64(); s_pf[k] += now_ - s_pf[7]; s_pf[7] = now_; } } while (0)
    for (int it = 0; it < a.max_iters; ++it) {
        const int par = it & 1;
        const uint32_t tag_in = a.x.base + (uint32_t)it, tag_out = tag_in + 1u;
        double* accp = a.acc + (size_t)par * a.acc_stride;

        // ======================= E-step over this warp's chunks =======================
        double ll = 0.0;
        {
            float* wp = wpar + warp * 80;
            const float4* wp4 = reinterpret_cast<const float4*>(wp);
            int cur_p = -1;
            float2 acc2[4 * kMom];
            for (int ci = cw0; ci < cw1; ++ci) {
                int p, start, len;
                if (ci < a.chunk_cap) {
                    p = sc_parent[ci]; start = sc_start[ci]; len = sc_len[ci];
                } else {
                    p = __ldg(a.chunk_parent + c0 + ci); start = __ldg(a.chunk_start + c0 + ci); len = __ldg(a.chunk_len + c0 + ci);
                }
                if (a.term && a.term[p]) {                            // adaptive build: a terminal parent's points take no part --
                    for (int r = lane; r < len; r += 32) a.slot[start + r] = 0;      // child 0, no moments, no log-likelihood term
                    continue;
                }
                if (p != cur_p) {
                    if (cur_p >= 0) tl_fold_stage(acc2, cur_p, lane, stage, stage_parent, a.stage_cap, &s_nslots, accp);
                    cur_p = p;
                    if (!tl_load_params(a.x.pk[me] + (size_t)8 * p * kPkWords, tag_in, wp, lane, a)) aborted = true;
#pragma unroll
                    for (int m = 0; m < 4 * kMom; ++m) acc2[m] = make_float2(0.f, 0.f);
                }
                for (int rb = 0; rb < len; rb += 32) {
                    const int r = rb + lane;
                    const bool valid = r < len;
                    const int i = start + (valid ? r : len - 1);
                    float x, y, z;
                    if (resident) {
                        x = sx[i - pt0]; y = sy[i - pt0]; z = sz[i - pt0];
                    } else {
                        x = a.px[i]; y = a.py[i]; z = a.pz[i];
                    }
                    const float2 xx = make_float2(x, x), yy = make_float2(y, y), zz = make_float2(z, z);
                    float2 e[4];
                    float m = kNegBig;
#pragma unroll
                    for (int cp = 0; cp < 4; ++cp) {                      // q2 of the 8 children, two per packed operation
                        const float4 A0 = wp4[cp * 5], A1 = wp4[cp * 5 + 1], A2 = wp4[cp * 5 + 2], A3 = wp4[cp * 5 + 3], A4 = wp4[cp * 5 + 4];
                        const float2 dx = t_fadd2(xx, make_float2(A0.x, A0.y)), dy = t_fadd2(yy, make_float2(A0.z, A0.w)),
                                     dz = t_fadd2(zz, make_float2(A1.x, A1.y));
                        float2 t0 = t_fmul2(make_float2(A4.x, A4.y), dz);                 // axz dz
                        t0 = t_ffma2(make_float2(A3.z, A3.w), dy, t0);                    // + axy dy
                        t0 = t_ffma2(make_float2(A2.x, A2.y), dx, t0);                    // + axx dx
                        float2 t1 = t_fmul2(make_float2(A4.z, A4.w), dz);                 // ayz dz
                        t1 = t_ffma2(make_float2(A2.z, A2.w), dy, t1);                    // + ayy dy
                        const float2 t2 = t_fmul2(make_float2(A3.x, A3.y), dz);           // azz dz
                        float2 q = t_ffma2(dz, t2, make_float2(A1.z, A1.w));              // c2 + ...
                        q = t_ffma2(dy, t1, q);
                        q = t_ffma2(dx, t0, q);
                        e[cp] = q;
                        m = fmaxf(m, fmaxf(q.x, q.y));
                    }
                    int best = 0;                                         // first maximum, like np.argmax (0 when every child is dead)
#pragma unroll
                    for (int cp = 3; cp >= 0; --cp) {
                        if (e[cp].y == m) best = 2 * cp + 1;
                        if (e[cp].x == m) best = 2 * cp;
                    }
                    float s = 0.f;
#pragma unroll
                    for (int cp = 0; cp < 4; ++cp) {
                        e[cp] = make_float2(ex2f(e[cp].x - m), ex2f(e[cp].y - m));
                        s += e[cp].x + e[cp].y;
                    }
                    const float lse2 = m + lg2f(s);
                    // hgmm_cupy_cpu_working.py:174-178: gamma = gamma/den if den > eps else zeros
                    const bool alive = valid && (lse2 > kLog2Eps15);
                    if (valid) {
                        a.slot[i] = (uint8_t)(alive ? best : 0);
                        ll += (double)(kLn2 * fmaxf(alive ? lse2 : kLog2Eps15, kLog2Eps15));
                    }
                    const float inv = alive ? __fdividef(1.0f, s) : 0.f;
#pragma unroll
                    for (int cp = 0; cp < 4; ++cp) {                      // centred moments of the 8 children
                        const float4 A0 = wp4[cp * 5], A1 = wp4[cp * 5 + 1];
                        const float2 dx = t_fadd2(xx, make_float2(A0.x, A0.y)), dy = t_fadd2(yy, make_float2(A0.z, A0.w)),
                                     dz = t_fadd2(zz, make_float2(A1.x, A1.y));
                        // accumulate() skips gamma < eps = 1e-15 (:100-101).  Not re-tested per child here: such a term is below
                        // 2^-49 of the point's own mass, i.e. below float32 resolution of every sum it could enter (the dead-point
                        // rule above, which DOES change sums, is kept exactly)
                        const float2 gam = t_fmul2(e[cp], make_float2(inv, inv));
                        const float2 gx = t_fmul2(gam, dx), gy = t_fmul2(gam, dy), gz = t_fmul2(gam, dz);
                        float2* A = acc2 + cp * kMom;
                        A[0] = t_fadd2(A[0], gam);
                        A[1] = t_fadd2(A[1], gx);
                        A[2] = t_fadd2(A[2], gy);
                        A[3] = t_fadd2(A[3], gz);
                        A[4] = t_ffma2(gx, dx, A[4]);
                        A[5] = t_ffma2(gx, dy, A[5]);
                        A[6] = t_ffma2(gx, dz, A[6]);
                        A[7] = t_ffma2(gy, dy, A[7]);
                        A[8] = t_ffma2(gy, dz, A[8]);
                        A[9] = t_ffma2(gz, dz, A[9]);
                    }
                }
            }
            if (cur_p >= 0) tl_fold_stage(acc2, cur_p, lane, stage, stage_parent, a.stage_cap, &s_nslots, accp);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ll += __shfl_xor_sync(0xffffffffu, ll, o);
        if (lane == 0) s_ll[warp] = ll;
        __syncthreads();
        TL_STAMP(0);
        // the parked sums -> fp64 atomics in L2.  Thread k of a group owns moment slot k and walks the stage slots in order,
        // adding runs of one parent before they leave: at the top levels (one parent, a slot per warp) a CTA sends 80 atomics.
        {
            const int ns = min(s_nslots, a.stage_cap);
            const int groups = ns > 24 ? kTlThreads / (8 * kMom) : 1;
            const int g = tid / (8 * kMom), k = tid - g * (8 * kMom);
            if (g < groups) {
                int run_p = -1;
                float run = 0.f;
                for (int sidx = g; sidx < ns; sidx += groups) {
                    const int ps = stage_parent[sidx];
                    if (ps != run_p) {
                        if (run_p >= 0 && run != 0.f) atomicAdd(accp + kAccHdr + (size_t)run_p * (8 * kMom) + k, (double)run);
                        run_p = ps;
                        run = 0.f;
                    }
                    run += stage[(size_t)sidx * (8 * kMom) + k];
                }
                if (run_p >= 0 && run != 0.f) atomicAdd(accp + kAccHdr + (size_t)run_p * (8 * kMom) + k, (double)run);
            }
        }
        // the level's log-likelihood of this iteration: one of THREE slots (read by every CTA after the barrier, zeroed by CTA 0
        // two iterations ahead: its last readers are then behind a barrier, its next writers behind another)
        double* llslot = a.qstate + 4 + (it % 3);
        if (tid == 0) {
            double t = 0.0;
            for (int w = 0; w < W; ++w) t += s_ll[w];
            if (t != 0.0) atomicAdd(llslot, t);
        }
        // ======================= one grid barrier: every local moment has landed =======================
        TL_STAMP(1);
        if (!tl_grid_sync(a, (unsigned)(it + 1) * (unsigned)G, &s_nslots)) { aborted = true; break; }
        TL_STAMP(2);

        // ======================= exchange + M-step + publish =======================
        if (b == 0 && tid == 0) {
            a.qstate[4 + (it + 2) % 3] = 0.0;
            if (R > 1) {                                                  // this rank's log-likelihood -> every rank (own included)
                const double ql = __ldcg(llslot);
                for (int r = 0; r < R; ++r) cell16_put(a.x.ll[r] + (size_t)par * kXchgMaxRanks + me, ql, tag_out);
            }
        }
        // pass 1: sums of nodes owned elsewhere -> the owner's window (no waiting in this pass).  One cell per thread, consecutive
        // threads on consecutive cells: a warp's store is 512 contiguous bytes on the NVLink, not 32 scattered 8-byte writes
        if (R > 1) {
            const int total = a.cnt * kMom;
            for (int idx = b * kTlThreads + tid; idx < total; idx += G * kTlThreads) {
                const int j = idx / kMom, k = idx - j * kMom;
                const int owner = (j >> 5) % R;
                if (owner == me) continue;
                double* Ag = accp + kAccHdr + idx;
                uint4* dst = a.x.mom[owner] + (size_t)par * a.x.mom_cap + ((size_t)me * owned_cap + (size_t)((j >> 5) / R) * 32 + (j & 31)) * kMom + k;
                cell16_put(dst, __ldcg(Ag), tag_out);
                *Ag = 0.0;
            }
        }
        // pass 2: owned nodes, one warp per 32-node slice (slices are dealt round-robin to the ranks, a rank's slices round-robin to
        // the CTAs): add the ranks' contributions in rank order, M-step, publish to every rank through a shared-memory
        // transpose so that the 320 parameter cells of the slice leave as contiguous 256-byte stores
        {
            float* pkst = pkst_all + warp * (32 * kPkWords);
            const int n_slices = (a.cnt + 31) >> 5;
            for (int os = warp * G + b; ; os += W * G) {                  // os: index among this rank's slices
                const int sl = me + R * os;
                if (sl >= n_slices) break;
                const int j = sl * 32 + lane;
                const bool live = j < a.cnt;
                double A[kMom];
#pragma unroll
                for (int k = 0; k < kMom; ++k) A[k] = 0.0;
                if (live) {
                    double* Ag = accp + kAccHdr + (size_t)j * kMom;
                    double own[kMom];
#pragma unroll
                    for (int k = 0; k < kMom; ++k) { own[k] = __ldcg(Ag + k); Ag[k] = 0.0; }
                    const uint4* src0 = a.x.mom[me] + (size_t)par * a.x.mom_cap + ((size_t)os * 32 + lane) * kMom;
                    for (int r = 0; r < R; ++r) {
                        if (r == me) {
#pragma unroll
                            for (int k = 0; k < kMom; ++k) A[k] += own[k];
                        } else {
                            double v[kMom];
                            if (!cell16_get<kMom>(src0 + (size_t)r * owned_cap * kMom, tag_out, v, a)) { aborted = true; break; }
#pragma unroll
                            for (int k = 0; k < kMom; ++k) A[k] += v[k];
                        }
                    }
                }
                if (__any_sync(0xffffffffu, aborted)) { aborted = true; break; }
                __syncwarp();
                if (live) {
                    const PackedComp pc = tree_mstep_apply(a.t, a.lb, j, A, a.n_total, a.ld);     // writes t.pi / t.mu / t.cov / t.packed
                    pkst[lane * kPkWords + 0] = pc.mx; pkst[lane * kPkWords + 1] = pc.my; pkst[lane * kPkWords + 2] = pc.mz;
                    pkst[lane * kPkWords + 3] = pc.c2; pkst[lane * kPkWords + 4] = pc.axx; pkst[lane * kPkWords + 5] = pc.ayy;
                    pkst[lane * kPkWords + 6] = pc.azz; pkst[lane * kPkWords + 7] = pc.axy; pkst[lane * kPkWords + 8] = pc.axz;
                    pkst[lane * kPkWords + 9] = pc.ayz;
                }
                __syncwarp();
                const int ncell = min(32, a.cnt - sl * 32) * kPkWords;
                for (int r = 0; r < R; ++r) {
                    unsigned long long* dst = a.x.pk[r] + (size_t)sl * 32 * kPkWords;
                    for (int c = lane; c < ncell; c += 32) cell8_put(dst + c, pkst[c], tag_out);
                }
                __syncwarp();
            }
        }
        TL_STAMP(3);
        // stopping rule.  One rank: every thread reads the slot itself (same value everywhere, no barrier, no broadcast).
        // Several ranks: the same R cells, summed in the same order, by thread 0 of every CTA of every rank.
        double q;
        if (R == 1) {
            q = __ldcg(llslot);
        } else {
            if (tid == 0) {
                double qq = 0.0;
                bool ok = true;
                for (int r = 0; r < R && ok; ++r) {
                    double v[1];
                    ok = cell16_get<1>(a.x.ll[me] + (size_t)par * kXchgMaxRanks + r, tag_out, v, a);
                    qq += v[0];
                }
                s_q = qq;
                s_i[5] = (ok && !tl_abort(a.ctrl)) ? 1 : 0;
            }
            __syncthreads();
            if (!s_i[5]) { aborted = true; break; }          // CTA-uniform (a thread that failed a poll has set ctrl[7])
            q = s_q;
        }
        TL_STAMP(4);
        const bool conv = fabs(q - prev_q) < (double)a.ls || it + 1 >= a.max_iters;
        prev_q = q;
        if (!conv) continue;

        // ======================= converged: leave the level with the full replicated model =======================
        if (R > 1) {
            // owners publish (pi, mu, Sigma) to the peers -- with the SAME slice -> warp mapping as the M-step above, so every lane
            // re-reads what it wrote itself.  (Any other mapping reads t.pi / t.mu / t.cov of a slice another CTA may still be
            // computing: the peers then install the previous iteration's parameters for it -- seen at 8 ranks as a 1e-2 error
            // below the root level while the root, owned by the reporting rank, was exact.)
            const int n_slices = (a.cnt + 31) >> 5;
            for (int os = warp * G + b; ; os += W * G) {
                const int sl = me + R * os;
                if (sl >= n_slices) break;
                const int j = sl * 32 + lane;
                if (j >= a.cnt) continue;
                const int g = a.lb + j;
                const float* c = a.t.cov + 9 * (size_t)g;
                const float f[kFinWords] = {a.t.pi[g], a.t.mu[3 * g], a.t.mu[3 * g + 1], a.t.mu[3 * g + 2], c[0], c[1], c[2], c[4], c[5], c[8]};
                for (int r = 0; r < R; ++r) {
                    if (r == me) continue;
                    unsigned long long* dst = a.x.fin[r] + (size_t)j * kFinWords;
#pragma unroll
                    for (int w = 0; w < kFinWords; ++w) cell8_put(dst + w, f[w], tag_out);
                }
            }
            for (int j = tid * G + b; j < a.cnt; j += kTlThreads * G) {      // everyone installs the nodes owned elsewhere
                if ((j >> 5) % R == me) continue;
                const int g = a.lb + j;
                float f[kFinWords], w[kPkWords];
                if (!cell8_get<kFinWords>(a.x.fin[me] + (size_t)j * kFinWords, tag_out, f, a) ||
                    !cell8_get<kPkWords>(a.x.pk[me] + (size_t)j * kPkWords, tag_out, w, a)) { aborted = true; break; }
                a.t.pi[g] = f[0];
                a.t.mu[3 * g] = f[1]; a.t.mu[3 * g + 1] = f[2]; a.t.mu[3 * g + 2] = f[3];
                float* c = a.t.cov + 9 * (size_t)g;
                c[0] = f[4]; c[1] = c[3] = f[5]; c[2] = c[6] = f[6]; c[4] = f[7]; c[5] = c[7] = f[8]; c[8] = f[9];
                PackedComp pc;
                pc.mx = w[0]; pc.my = w[1]; pc.mz = w[2]; pc.c2 = w[3]; pc.axx = w[4]; pc.ayy = w[5]; pc.azz = w[6]; pc.axy = w[7];
                pc.axz = w[8]; pc.ayz = w[9]; pc.pad0 = 0.f; pc.pad1 = 0.f;
                a.t.packed[g] = pc;
            }
        }
        if (b == 0 && tid == 0) {
            a.ctrl[0] = 1;
            a.ctrl[1] = it + 1;
            a.qstate[0] = q;
            a.qstate[1] = q;
        }
        break;
    }
    if (aborted && tid == 0) atomicExch(a.ctrl + 7, 1);
    if (a.prof && tid == 0)
        for (int k = 0; k < 6; ++k) a.prof[(size_t)b * 8 + k] += s_pf[k];
#undef TL_STAMP
}
