// cloud_sort.cu -- stable counting sort of a point cloud by a 16^3 Morton cell grid.
//
// em_flat8_kernel (flat_em8.cu) accumulates every component's moments about one origin per CTA; that is accurate when a
// CTA's points are neighbours.  This file produces that order, once per hgmm_set_points, in five small launches:
//   bbox   ordered-integer atomicMax over the six extremes
//   rank   block b takes points [b*ppb, (b+1)*ppb): cell of every point, its rank among the block's EARLIER points of the
//          same cell (warps take turns in index order; __match_any_sync inside a warp), the block's 4096 cell counts
//   scan1  per cell: exclusive prefix over blocks (thread = cell, coalesced over cells)
//   scan2  exclusive prefix over the 4096 cell totals (one CTA)
//   place  dest = cell base + block offset + rank
// Equal cells keep their input order, so the result is a pure function of the input (bit-reproducible fits).  The reference
// has no counterpart: its kernels (gmm_kernels.cu:278-350) take the cloud in file order; EM sums are order-independent up to
// rounding.
#include "common.cuh"
#include "kernels.h"

namespace hgmm {

constexpr int kCellBits = 4;                      // per axis
constexpr int kCells = 1 << (3 * kCellBits);      // 4096
constexpr int kSortThreads = 256;
constexpr int kSortMaxBlocks = 1024;

__device__ __forceinline__ unsigned ord_enc(float f) {      // monotone float -> unsigned
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord_dec(unsigned e) {
    return __uint_as_float((e & 0x80000000u) ? (e & 0x7fffffffu) : ~e);
}

// box[0..2] = max of enc(x), enc(y), enc(z); box[3..5] = max of ~enc (i.e. the minima); all start at 0
__global__ void __launch_bounds__(256) sort_bbox_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                        const float* __restrict__ z, int64_t n, unsigned* __restrict__ box) {
    unsigned v[6] = {0u, 0u, 0u, 0u, 0u, 0u};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float c[3] = {x[i], y[i], z[i]};
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            if (!(fabsf(c[a]) <= 3.0e38f)) continue;                       // NaN / Inf do not stretch the grid
            const unsigned e = ord_enc(c[a]);
            v[a] = max(v[a], e);
            v[3 + a] = max(v[3 + a], ~e);
        }
    }
#pragma unroll
    for (int a = 0; a < 6; ++a) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[a] = max(v[a], __shfl_xor_sync(0xffffffffu, v[a], o));
        if ((threadIdx.x & 31) == 0 && v[a]) atomicMax(box + a, v[a]);
    }
}

__device__ __forceinline__ unsigned spread3(unsigned v) {   // 4 bits -> every third bit
    return (v & 1u) | ((v & 2u) << 2) | ((v & 4u) << 4) | ((v & 8u) << 6);
}

__device__ __forceinline__ int cell_of(float x, float y, float z, const unsigned* __restrict__ box) {
    const float lo[3] = {ord_dec(~box[3]), ord_dec(~box[4]), ord_dec(~box[5])};
    const float hi[3] = {ord_dec(box[0]), ord_dec(box[1]), ord_dec(box[2])};
    const float ext = fmaxf(fmaxf(hi[0] - lo[0], hi[1] - lo[1]), fmaxf(hi[2] - lo[2], 1e-30f));
    const float sc = (float)(1 << kCellBits) / ext;
    const float c[3] = {x, y, z};
    unsigned g[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float t = (c[a] - lo[a]) * sc;
        int q = t > 0.f ? (int)fminf(t, 1.0e6f) : 0;                       // NaN -> 0
        g[a] = (unsigned)min(q, (1 << kCellBits) - 1);
    }
    return (int)(spread3(g[0]) | (spread3(g[1]) << 1) | (spread3(g[2]) << 2));
}

// key[i] = cell << 20 | rank within (block, cell)  (ppb <= 2^20);  hist[b][cell] = the block's count
__global__ void __launch_bounds__(kSortThreads) sort_rank_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                                 const float* __restrict__ z, int64_t n, int ppb,
                                                                 const unsigned* __restrict__ box, unsigned* __restrict__ key,
                                                                 unsigned* __restrict__ hist) {
    __shared__ unsigned cnt[kCells];
    for (int c = threadIdx.x; c < kCells; c += kSortThreads) cnt[c] = 0u;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t base = (int64_t)blockIdx.x * ppb;
    for (int r = 0; r < ppb; r += kSortThreads) {
        const int64_t i = base + r + threadIdx.x;
        const bool in = r + (int)threadIdx.x < ppb && i < n;
        const int cell = in ? cell_of(x[i], y[i], z[i], box) : -1;
        for (int w = 0; w < kSortThreads / 32; ++w) {                      // warps in index order
            if (warp == w) {
                const unsigned peers = __match_any_sync(0xffffffffu, cell);
                if (in) {
                    const int leader = __ffs(peers) - 1;
                    unsigned b0 = 0u;
                    if (lane == leader) b0 = cnt[cell];
                    b0 = __shfl_sync(peers, b0, leader);
                    key[i] = ((unsigned)cell << 20) | (b0 + (unsigned)__popc(peers & ((1u << lane) - 1u)));
                    if (lane == leader) cnt[cell] = b0 + (unsigned)__popc(peers);
                }
            }
            __syncthreads();
        }
    }
    unsigned* h = hist + (size_t)blockIdx.x * kCells;
    for (int c = threadIdx.x; c < kCells; c += kSortThreads) h[c] = cnt[c];
}

// hist[b][cell] -> exclusive prefix over b; total[cell]
__global__ void __launch_bounds__(256) sort_scan1_kernel(unsigned* __restrict__ hist, int nblocks, unsigned* __restrict__ total) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= kCells) return;
    unsigned run = 0u;
    int b = 0;
    for (; b + 4 <= nblocks; b += 4) {                                     // four independent loads in flight
        unsigned* p = hist + (size_t)b * kCells + c;
        const unsigned t0 = p[0], t1 = p[kCells], t2 = p[2 * kCells], t3 = p[3 * kCells];
        p[0] = run; run += t0;
        p[kCells] = run; run += t1;
        p[2 * kCells] = run; run += t2;
        p[3 * kCells] = run; run += t3;
    }
    for (; b < nblocks; ++b) {
        unsigned* p = hist + (size_t)b * kCells + c;
        const unsigned t = *p;
        *p = run;
        run += t;
    }
    total[c] = run;
}

// total[cell] -> exclusive prefix (one CTA of 1024 threads, four cells each)
__global__ void __launch_bounds__(1024) sort_scan2_kernel(unsigned* __restrict__ total) {
    __shared__ unsigned wsum[32];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    unsigned v[4], s = 0u;
#pragma unroll
    for (int q = 0; q < 4; ++q) { v[q] = total[4 * t + q]; s += v[q]; }
    unsigned inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        unsigned w = wsum[lane], wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned u = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += u;
        }
        wsum[lane] = wi - w;
    }
    __syncthreads();
    unsigned run = wsum[warp] + inc - s;
#pragma unroll
    for (int q = 0; q < 4; ++q) { total[4 * t + q] = run; run += v[q]; }
}

__global__ void __launch_bounds__(256) sort_place_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                         const float* __restrict__ z, int64_t n, int ppb,
                                                         const unsigned* __restrict__ key, const unsigned* __restrict__ hist,
                                                         const unsigned* __restrict__ cellbase, float* __restrict__ sx,
                                                         float* __restrict__ sy, float* __restrict__ sz) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned kk = key[i];
    const unsigned cell = kk >> 20, rank = kk & 0xfffffu;
    const int64_t b = i / ppb;
    const size_t d = (size_t)cellbase[cell] + hist[(size_t)b * kCells + cell] + rank;
    sx[d] = x[i];
    sy[d] = y[i];
    sz[d] = z[i];
}

static void sort_shape(int64_t n, int* ppb, int* nblocks) {
    int64_t p = (n + kSortMaxBlocks - 1) / kSortMaxBlocks;
    p = (p + kSortThreads - 1) / kSortThreads * kSortThreads;
    if (p < kSortThreads) p = kSortThreads;
    *ppb = (int)p;
    *nblocks = (int)((n + p - 1) / p);
}

// scratch: unsigned box[8] | total[4096] | key[n] | hist[nblocks][4096]
size_t cloud_sort_scratch_bytes(int64_t n) {
    int ppb, nb;
    sort_shape(n, &ppb, &nb);
    return sizeof(unsigned) * (8 + (size_t)kCells + (size_t)n + (size_t)nb * kCells);
}

cudaError_t launch_cloud_sort(const float* x, const float* y, const float* z, int64_t n, float* sx, float* sy, float* sz,
                              void* scratch, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    int ppb, nb;
    sort_shape(n, &ppb, &nb);
    if (ppb > (1 << 20)) return cudaErrorInvalidValue;                     // > 2^30 points: the rank field would overflow
    unsigned* box = static_cast<unsigned*>(scratch);
    unsigned* total = box + 8;
    unsigned* key = total + kCells;
    unsigned* hist = key + n;
    cudaError_t e = cudaMemsetAsync(box, 0, 8 * sizeof(unsigned), s);
    if (e != cudaSuccess) return e;
    const int gb = (int)((n + 256 * 8 - 1) / (256 * 8));
    sort_bbox_kernel<<<gb < 1 ? 1 : (gb > 1184 ? 1184 : gb), 256, 0, s>>>(x, y, z, n, box);
    sort_rank_kernel<<<nb, kSortThreads, 0, s>>>(x, y, z, n, ppb, box, key, hist);
    sort_scan1_kernel<<<kCells / 256, 256, 0, s>>>(hist, nb, total);
    sort_scan2_kernel<<<1, 1024, 0, s>>>(total);
    sort_place_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(x, y, z, n, ppb, key, hist, total, sx, sy, sz);
    return cudaGetLastError();
}

}  // namespace hgmm
