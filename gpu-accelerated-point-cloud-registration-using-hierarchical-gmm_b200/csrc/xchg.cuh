// xchg.cuh -- multi-GPU flat M-step in ONE kernel: fold this rank's partial rows, exchange the J x 10 sufficient statistics
// with every peer over NVLink (stores into peer memory, NVSwitch routes them), add the ranks' contributions in rank order,
// finalize.  Replaces the three-launch sequence  flat_reduce_kernel -> ncclAllReduce(fp64) -> flat_finalize_kernel
// (the reference has no distributed code) on boxes where the ranks can map each other's memory.
// Included by flat_em.cu (inside namespace hgmm) after finalize_component / finalize_bookkeeping: one translation unit, no -rdc.
//
// Exchange window (one per rank, cudaMalloc'ed, mapped into every peer with cudaIpcOpenMemHandle), two parities so that
// epoch e+1 never overwrites what a slower rank still reads for epoch e:
//     uint4 data[2][kXchgMaxRanks][kXchgHdr + kMom * kXchgMaxJ]     one fp64 value per 16-byte cell
//                                                                   [2*cta], [2*cta+1]: that CTA's (sum log-lik, live points);
//                                                                   then moment k of component j at kXchgHdr + k*Jp + j
// A cell is two self-validating 8-byte words, (low half, epoch) and (high half, epoch): the receiver polls the cell until
// both words carry the current epoch, so no flag, no fence and no second NVLink round trip separate the data from its
// "ready" signal (the low-latency protocol NCCL uses for small messages, here fused into the M-step).  8-byte stores are
// single-copy atomic, the two words of a cell are validated independently.
// CTA b of every rank owns components [32b, 32b+32): it stores its 32 x 10 sums (+ its copy of the two scalars) into
// slot `rank` of every PEER window, then polls its OWN window for the peers' CTA b (warp r collects rank r, the 12 cells of a
// lane loaded together so the R-1 polls cost one L2 round trip, not 12 (R-1)) -- a CTA never waits for a CTA that is
// waiting for it, so no co-scheduling is needed.  Every rank adds the contributions in rank order 0..R-1 (its own from
// registers): replicas stay bit-identical.  A rank is at most one epoch ahead of any other (its next push needs their
// previous one), hence two parities suffice.
// The poll has a %globaltimer deadline (HGMM_XCHG_TIMEOUT_S, default 30 s); on expiry ctrl[7] is set, every CTA that sees it
// leaves without finalizing, and the host reports HGMM_ERR_NCCL instead of hanging.
__device__ __forceinline__ void xchg_put(uint4* cell, double v, uint32_t epoch) {
    const unsigned long long u = (unsigned long long)__double_as_longlong(v);
    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(cell), "r"((uint32_t)u), "r"(epoch) : "memory");
    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(reinterpret_cast<char*>(cell) + 8), "r"((uint32_t)(u >> 32)), "r"(epoch)
                 : "memory");
}
// one peer's contribution for this lane's component: all 12 cells are loaded together (independent loads, ONE L2 round
// trip when the data has arrived) and re-polled until every word carries the epoch.  false on timeout.
__device__ __forceinline__ bool xchg_get_row(const uint4* src, size_t Jp, int j, int b, bool hdr_lane, uint32_t epoch, double* v,
                                             unsigned long long timeout_ns, const int* ctrl) {
    unsigned long long t0 = 0ull;
    unsigned spins = 0;
    for (;;) {
        uint4 c[kMom + 2];
#pragma unroll
        for (int k = 0; k < kMom; ++k) {
            const uint4* cell = src + kXchgHdr + (size_t)k * Jp + j;
            asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(c[k].x), "=r"(c[k].y), "=r"(c[k].z), "=r"(c[k].w)
                         : "l"(cell) : "memory");
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const uint4* cell = src + 2 * b + h;
            if (hdr_lane)
                asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(c[kMom + h].x), "=r"(c[kMom + h].y),
                             "=r"(c[kMom + h].z), "=r"(c[kMom + h].w) : "l"(cell) : "memory");
            else
                c[kMom + h] = make_uint4(0u, epoch, 0u, epoch);
        }
        bool all = true;
#pragma unroll
        for (int k = 0; k < kMom + 2; ++k) all = all && c[k].y == epoch && c[k].w == epoch;
        if (all) {
#pragma unroll
            for (int k = 0; k < kMom + 2; ++k) v[k] = __longlong_as_double((long long)(((unsigned long long)c[k].z << 32) | c[k].x));
            return true;
        }
        if ((++spins & 255u) == 0u) {                          // deadline (XchgView::timeout_ns, default 30 s): a peer died or the
            unsigned long long now;                            // ranks' call sequences diverged; another CTA's verdict ends the wait too
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0ull) t0 = now;
            int bad;
            asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(bad) : "l"(ctrl + 7) : "memory");
            if (bad || now - t0 > timeout_ns) return false;
        }
    }
}

__global__ void __launch_bounds__(512) flat_reduce_exchange_finalize_kernel(FlatModel m, const float* __restrict__ partial,
                                                                            const double* __restrict__ rowaux, int rows,
                                                                            int* __restrict__ ctrl, int* __restrict__ done_at, int it,
                                                                            double* __restrict__ ll_hist, double n_total, XchgView xc) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const bool done = __ldcg(done_at + it) != 0;       // identical on every rank: the replicas are bit-identical
    if (done) {
        if (blockIdx.x == 0 && threadIdx.x == 0) done_at[it + 1] = 1;
        return;
    }
    __shared__ double sm[16][kMom][33];                 // the 16 warps' partial sums; afterwards the peers' rows
    __shared__ double s_aux[16][2];
    __shared__ int s_bad;
    static_assert(sizeof(double) * kXchgMaxRanks * (kMom + 2) * 32 <= sizeof(double) * 16 * kMom * 33, "peer rows must fit in sm");
    double(*xs)[kMom + 2][32] = reinterpret_cast<double(*)[kMom + 2][32]>(&sm[0][0][0]);     // [rank][cell][lane]
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int j = blockIdx.x * 32 + lane;
    if (tid == 0) s_bad = 0;
    {
        double v[kMom];
#pragma unroll
        for (int k = 0; k < kMom; ++k) v[k] = 0.0;
        const size_t stride = (size_t)kMom * m.Jp;
        for (int r = w; r < rows; r += 16) {
            const float* src = partial + (size_t)r * stride + j;
#pragma unroll
            for (int k = 0; k < kMom; ++k) v[k] += (double)__ldcg(src + (size_t)k * m.Jp);
        }
#pragma unroll
        for (int k = 0; k < kMom; ++k) sm[w][k][lane] = v[k];
        double l = 0.0, c = 0.0;
        for (int q = w * 32 + lane; q < rows; q += 512) {
            l += __ldcg(rowaux + 2 * q);
            c += __ldcg(rowaux + 2 * q + 1);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            l += __shfl_xor_sync(0xffffffffu, l, o);
            c += __shfl_xor_sync(0xffffffffu, c, o);
        }
        if (lane == 0) {
            s_aux[w][0] = l;
            s_aux[w][1] = c;
        }
    }
    __syncthreads();
    const int R = xc.nranks, me = xc.rank, b = blockIdx.x;
    const uint32_t epoch = xc.epoch;
    const size_t slot = kXchgHdr + (size_t)kMom * kXchgMaxJ;
    const size_t par_off = (size_t)(epoch & 1u) * kXchgMaxRanks * slot;
    double A[kMom], ll = 0.0, total = 0.0;
    if (w == 0) {
#pragma unroll
        for (int k = 0; k < kMom; ++k) {
            double t = 0.0;
#pragma unroll
            for (int q = 0; q < 16; ++q) t += sm[q][k][lane];
            A[k] = t;
        }
        for (int q = 0; q < 16; ++q) {
            ll += s_aux[q][0];
            total += s_aux[q][1];
        }
        // ---------------- push this rank's sums into every peer window: 16-byte cells, 512-byte rows per warp
        for (int r = 0; r < R; ++r) {
            if (r == me) continue;
            uint4* dst = xc.data[r] + par_off + (size_t)me * slot;
#pragma unroll
            for (int k = 0; k < kMom; ++k) xchg_put(dst + kXchgHdr + (size_t)k * m.Jp + j, A[k], epoch);
            if (lane == 0) {
                xchg_put(dst + 2 * b, ll, epoch);
                xchg_put(dst + 2 * b + 1, total, epoch);
            }
        }
    }
    __syncthreads();                                    // warp 0 has consumed sm: it now receives the peers' rows
    // ---------------- warp r collects rank r's contribution from this rank's own window (all peers polled in parallel)
    if (w < R && w != me) {
        double v[kMom + 2];
        const bool ok = xchg_get_row(xc.data[me] + par_off + (size_t)w * slot, (size_t)m.Jp, j, b, lane == 0, epoch, v, xc.timeout_ns, ctrl);
        if (!ok) {
            s_bad = 1;
            atomicExch(ctrl + 7, 1);     // the other CTAs of the grid stop waiting and skip their finalize as well
        }
#pragma unroll
        for (int k = 0; k < kMom + 2; ++k) xs[w][k][lane] = v[k];
    }
    __syncthreads();
    if (w != 0) return;
    {
        int bad;
        asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(bad) : "l"(ctrl + 7) : "memory");
        if (s_bad || bad) return;        // no CTA that has seen the failure updates its components (best effort across CTAs)
    }
    // ---------------- add the ranks' contributions in rank order (own from registers)
    {
        double S[kMom], sll = 0.0, stot = 0.0;
#pragma unroll
        for (int k = 0; k < kMom; ++k) S[k] = 0.0;
        for (int r = 0; r < R; ++r) {
            if (r == me) {
#pragma unroll
                for (int k = 0; k < kMom; ++k) S[k] += A[k];
                sll += ll;
                stot += total;
            } else {
#pragma unroll
                for (int k = 0; k < kMom; ++k) S[k] += xs[r][k][lane];
                sll += xs[r][kMom][0];
                stot += xs[r][kMom + 1][0];
            }
        }
#pragma unroll
        for (int k = 0; k < kMom; ++k) A[k] = S[k];
        ll = sll;
        total = stot;
    }
    float my_c2 = -INFINITY;
    if (j < m.J) my_c2 = finalize_component(m, j, A, total, n_total);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) my_c2 = fmaxf(my_c2, __shfl_xor_sync(0xffffffffu, my_c2, o));
    if (lane == 0) m.cref_blocks[blockIdx.x] = my_c2;
    if (blockIdx.x == 0 && lane == 0) finalize_bookkeeping(m, ll, ctrl, done_at, it, ll_hist, n_total);
}

cudaError_t launch_flat_reduce_exchange_finalize(const FlatModel& m, const float* partial, const double* rowaux, int rows, int* ctrl,
                                                 int* done_at, int it, double* ll_hist, double n_total, const XchgView& xc,
                                                 cudaStream_t s) {
    if (m.Jp > kXchgMaxJ || m.Jp / 32 > kXchgMaxCtas || xc.nranks > kXchgMaxRanks) return cudaErrorInvalidValue;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(m.Jp / 32);
    cfg.blockDim = dim3(512);
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, flat_reduce_exchange_finalize_kernel, m, partial, rowaux, rows, ctrl, done_at, it, ll_hist, n_total,
                              xc);
}

