// kernels.h -- host-visible launchers of libhgmm's CUDA kernels (internal; the public ABI is include/hgmm.h)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/hgmm.h"

namespace hgmm {

struct PackedComp;

// Function attributes (cudaFuncSetAttribute) and device limits are PER DEVICE: a process may own contexts on several devices
// (include/hgmm.h: hgmm_create(device)), so one-time setup is remembered per device ordinal, never in a process-wide flag.
struct DeviceOnce {
    bool done[64] = {};
    bool first() {               // true exactly once per current device
        int d = 0;
        cudaGetDevice(&d);
        d &= 63;
        if (done[d]) return false;
        done[d] = true;
        return true;
    }
};
inline int device_smem_optin() {  // opt-in shared memory per block of the current device
    static int cache[64] = {};
    int d = 0;
    cudaGetDevice(&d);
    d &= 63;
    if (!cache[d]) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, d) != cudaSuccess || v <= 0) v = 227 * 1024;
        cache[d] = v;
    }
    return cache[d];
}

constexpr int kMaxFlatJ = 1024;          // fused flat kernel keeps <= 4 components per thread in registers

struct FlatModel {
    int J, Jp;                 // components, padded to a multiple of 32
    int cov_type, flavor, sigma_bug;
    float tol;
    float* means;              // [Jp,3]
    float* covs;               // FULL [Jp,9] | DIAG [Jp,3] | SPHERICAL [Jp]
    float* weights;            // [Jp]
    float* inv_cov;            // PY flavour: 1/std, [Jp,3] | [Jp]
    PackedComp* packed;        // [Jp]
    float* cref_blocks;        // [Jp/32] max c2 of each 32-component slot (the sweep's fixed reference is their maximum)
};

// ---- multi-GPU exchange window of the flat M-step (xchg.cuh): every rank maps every peer's window
constexpr int kXchgMaxRanks = 8;
constexpr int kXchgMaxCtas = 32;                 // Jp / 32 <= 32
constexpr int kXchgMaxJ = 1024;
constexpr int kXchgHdr = 2 * kXchgMaxCtas;       // doubles: per-CTA (sum log-lik, live points)
constexpr size_t kXchgCells = (size_t)2 * kXchgMaxRanks * (kXchgHdr + (size_t)kMom * kXchgMaxJ);
constexpr size_t kXchgBytes = kXchgCells * 16;   // one fp64 value per 16-byte self-validating cell (xchg.cuh)
struct XchgView {
    uint4* data[kXchgMaxRanks];                  // window of rank r (own window: the local pointer)
    int rank, nranks;
    uint32_t epoch;                              // same on every rank; parity selects the half of the window
    unsigned long long timeout_ns;               // deadline of a wait on a peer
};

// tree-registration exchange (registration.cu: reg_solve_kernel): every rank's (M0, M1) of every node as 16-byte cells
// [2 parities][ranks][kRegXchgNodes * 4]; sized for depth <= 4 (4680 nodes), deeper trees keep ncclAllReduce
constexpr int kRegXchgNodes = 4680;
constexpr size_t kRegXchgCells = (size_t)2 * kXchgMaxRanks * kRegXchgNodes * 4;
constexpr size_t kRegXchgBytes = kRegXchgCells * 16;
struct RegXchgView {
    uint4* data[kXchgMaxRanks];                  // rank r's region as mapped here (own entry: the local pointer); all null = no exchange
    int rank, nranks;
    uint32_t base;                               // tag of iteration it is base + it + 1
    unsigned long long timeout_ns;
};

struct TreeModel {
    int L;                     // levels
    int nt;                    // total nodes 8(8^L-1)/7
    float* pi;                 // [nt]
    float* mu;                 // [nt,3]
    float* cov;                // [nt,9] row-major
    float* cplx;               // [nt] lambda_min / trace (registration pruning)
    PackedComp* packed;        // [nt]
};

// per-level work decomposition of the permuted cloud
struct TreeWork {
    float *x, *y, *z;          // points in current (permuted) order
    int* perm;                 // original index of permuted point i
    int* pnode;                // level-local index of the parent node of permuted point i
    uint8_t* slot;             // child slot (0..7) chosen by the last E-step
    int* chunk_parent;         // [n_chunks] level-local parent index
    int* chunk_start;          // [n_chunks] first permuted point
    int* chunk_len;            // [n_chunks] number of points (<= chunk_points)
};

// flat_em.cu
void launch_aos_to_soa(const float* xyz, int64_t n, float* x, float* y, float* z, cudaStream_t s);
void launch_aos_to_soa_transform(const float* xyz, int64_t n, const double* Rt, float* x, float* y, float* z, cudaStream_t s);
void launch_flat_pack(const FlatModel& m, int first, cudaStream_t s);
void launch_flat_pack_init(const FlatModel& m, int* ctrl, int* done_at, int n_done, cudaStream_t s);
void launch_flat_finalize(const FlatModel& m, const double* acc, int* ctrl, int* done_at, int it, double* ll_hist, double n_total,
                          cudaStream_t s);
void launch_flat_reduce_finalize(const FlatModel& m, const float* partial, const double* rowaux, int rows, int* ctrl, int* done_at,
                                 int it, double* ll_hist, double n_total, cudaStream_t s);
// flat_em2.cu
void flat2_plan(int n, int Jp, int num_sms, int one_cta_per_sm, int* JT, int* W, int* Sdiv, int* G, int* grid, int* big);
cudaError_t launch_em_flat2(const float* x, const float* y, const float* z, int n, const FlatModel& m, const float* cref_blocks,
                            int JT, int W, int Sdiv, int G, int grid, int big, float* partial, double* rowaux,
                            const int* done_flag, cudaStream_t s);
cudaError_t launch_flat_reduce_exchange_finalize(const FlatModel& m, const float* partial, const double* rowaux, int rows, int* ctrl,
                                                 int* done_at, int it, double* ll_hist, double n_total, const XchgView& xc,
                                                 cudaStream_t s);
cudaError_t launch_flat_reduce(const float* partial, const double* rowaux, int rows, const FlatModel& m, double* acc,
                               const int* done_flag, cudaStream_t s);
int flat_pick_tile(int n, int num_sms, int requested);
cudaError_t launch_em_flat(const float* x, const float* y, const float* z, int n, const FlatModel& m, double* acc,
                           const int* ctrl, int num_sms, int tile_points, cudaStream_t s);
cudaError_t launch_predict(const float* x, const float* y, const float* z, int n, const PackedComp* packed, int Jp,
                           int32_t* labels, int num_sms, cudaStream_t s);
cudaError_t launch_level_ll(const float* x, const float* y, const float* z, int n, const PackedComp* packed, int Jp,
                            double* acc, const int* done_flag, int num_sms, cudaStream_t s);
cudaError_t launch_ffma_peak(float* out, int blocks, int iters, int mode, cudaStream_t s);

// flat_em3.cu (packed FP32)
void flat3_plan(int n, int Jp, int num_sms, int one_cta_per_sm, int* W, int* Sdiv, int* G, int* grid, int* big);
cudaError_t launch_em_flat3(const float* x, const float* y, const float* z, int n, const FlatModel& m, const float* cref_blocks,
                            int W, int Sdiv, int G, int grid, int big, float* partial, double* rowaux, const int* done_flag,
                            cudaStream_t s);
cudaError_t launch_ffma2_peak(float* out, int blocks, int iters, cudaStream_t s);
// flat_em5.cu (packed FP32, densities staged in shared memory per chunk of points)
cudaError_t launch_em_flat5(const float* x, const float* y, const float* z, int n, const FlatModel& m, const float* cref_blocks,
                            int W, int grid, float* partial, double* rowaux, const int* done_flag, cudaStream_t s);

// flat_em7.cu (flat_em5's arithmetic, chunks double-buffered, mbarrier arrive/wait instead of CTA barriers)
cudaError_t launch_em_flat7(const float* x, const float* y, const float* z, int n, const FlatModel& m, const float* cref_blocks,
                            int P, int grid, float* partial, double* rowaux, const int* done_flag, cudaStream_t s);

// flat_em8.cu (flat_em7's density pass; the moment pass about one origin per CTA, ten FFMA2 per pair; needs a cell-sorted cloud)
cudaError_t launch_em_flat8(const float* x, const float* y, const float* z, int n, const FlatModel& m, const float* cref_blocks,
                            int P, int grid, int chol, int stagger, float* partial, double* rowaux, const int* done_flag,
                            cudaStream_t s);
// cloud_sort.cu: stable counting sort of a cloud by a 16^3 Morton cell grid (scratch: cloud_sort_scratch_bytes(n))
size_t cloud_sort_scratch_bytes(int64_t n);
cudaError_t launch_cloud_sort(const float* x, const float* y, const float* z, int64_t n, float* sx, float* sy, float* sz,
                              void* scratch, cudaStream_t s);

// flat_em6.cu (one component per thread, densities staged in shared memory)
cudaError_t launch_em_flat6(const float* x, const float* y, const float* z, int n, const FlatModel& m, const float* cref_blocks,
                            int grid, float* partial, double* rowaux, const int* done_flag, cudaStream_t s);

// tree_em.cu
void launch_tree_init(const TreeModel& t, const float* init_means, float sig2, cudaStream_t s);
void launch_tree_pack_all(const TreeModel& t, cudaStream_t s);
cudaError_t launch_tree_estep(const TreeWork& w, const TreeModel& t, int level, double* acc, int n_chunks_bound,
                              const int* n_chunks_dev, const int* done_flag, int scalar_variant, cudaStream_t s);
void launch_tree_mstep(const TreeModel& t, int level, double* acc, double n_total, float ld, int* ctrl, int* done_at, int it,
                       int merge_converge, double* qstate, float ls, int max_iters, int* prog, cudaStream_t s);
void launch_tree_converge(double* acc, int* ctrl, int* done_at, int it, double* qstate, float ls, int max_iters, int* prog,
                          cudaStream_t s);
// host view of the tree exchange windows (device pointers of every rank's regions; own entries are local allocations)
struct TreeXchgHost {
    void* pk[kXchgMaxRanks];
    void* fin[kXchgMaxRanks];
    void* mom[kXchgMaxRanks];
    void* ll[kXchgMaxRanks];
    int rank, nranks;
    uint32_t base;
    unsigned long long mom_cap;
    unsigned long long timeout_ns;
};
// byte layout of one rank's tree exchange region: parameter cells | final-model cells | moment cells [2 parities] | log-lik cells
struct TreeWinLayout {
    size_t pk_off, fin_off, mom_off, ll_off, bytes;
    unsigned long long mom_cap;                  // 16-byte cells per parity
};
inline TreeWinLayout tree_win_layout(int max_level) {
    size_t cnt = 1;
    for (int i = 0; i < max_level; ++i) cnt *= 8;                                // nodes of the deepest level
    TreeWinLayout w;
    w.pk_off = 0;
    w.fin_off = w.pk_off + cnt * 10 * 8;
    w.mom_off = w.fin_off + cnt * 10 * 8;
    w.mom_cap = (unsigned long long)(cnt + 64 * kXchgMaxRanks) * 10;             // R * owned_cap <= cnt + 32 R (+ slack)
    w.ll_off = w.mom_off + 2 * (size_t)w.mom_cap * 16;
    w.bytes = w.ll_off + 2 * (size_t)kXchgMaxRanks * 16;
    return w;
}
void tree_level_plan(int n, int chunk_points, int level, int num_sms, int smem_optin, int* pt_cap, int* chunk_cap, int* stage_cap,
                     size_t* smem_bytes);
cudaError_t launch_tree_level(const TreeWork& w, const TreeModel& t, int level, int n, double* acc, size_t acc_stride,
                              const int* n_chunks_dev, double n_total, float ld, float ls, int max_iters, int* ctrl, double* qstate,
                              unsigned* gbar, int chunk_points, const TreeXchgHost& xh, int num_sms, long long* prof,
                              const uint8_t* term, cudaStream_t s);
void launch_tree_prune(const TreeModel& t, int level, double n_total, float lambda_c, float min_points, uint8_t* term, cudaStream_t s);
void launch_tree_zero_ll(double* acc, const int* done_flag, cudaStream_t s);
void launch_tree_cplx(const TreeModel& t, cudaStream_t s);
void launch_tree_current(const TreeWork& w, int n, int level, int64_t* current, cudaStream_t s);
void launch_iota(int* p, int n, cudaStream_t s);

struct PartitionScratch {
    uint16_t* group_off;       // [n_groups,8] exclusive offset of each 32-point group inside its 1024-point tile
    uint32_t* tile_cnt;        // [n_tiles,8]
    uint32_t* tile_off;        // [n_tiles+1,8] exclusive scan (last row = totals)
    uint32_t* seg_base;        // [n_parents+1,8] prefix counts at the segment starts
    int* seg_start;            // [n_parents+1] current segments (by parent)
    int* new_seg_start;        // [8*n_parents+1] segments after the split
    int* chunk_cnt;            // [8*n_parents] chunks per new segment
    int* chunk_off;            // [8*n_parents+1] exclusive scan
};
// splits every parent segment 8 ways by `slot` (stable), producing the next level's ordering in `dst`
cudaError_t launch_partition(const TreeWork& src, TreeWork& dst, int n, int n_parents, PartitionScratch& ps, int chunk_points,
                             int* n_chunks_dev, cudaStream_t s);
// level-0 work list: one segment [0,n)
cudaError_t launch_root_chunks(TreeWork& w, int n, const PartitionScratch& ps, int chunk_points, int* n_chunks_dev,
                               cudaStream_t s);

// registration.cu
cudaError_t launch_reg_estep(const float* tx, const float* ty, const float* tz, int n, const double* Rt, const TreeModel& t,
                             float lambda_c, double* racc, int want_m2, const int* ctrl, cudaStream_t s);
cudaError_t launch_reg_solve(const TreeModel& t, double* racc, int zero_after, int solver, double* Rt, double* q_hist,
                             double* qstate, int* ctrl, float tol, const RegXchgView* xv, cudaStream_t s);
void launch_transform_soa(const float* tx, const float* ty, const float* tz, int n, const double* Rt, float* ox, float* oy, float* oz,
                          const int* ctrl, cudaStream_t s);
cudaError_t launch_reg_flat_solve(const FlatModel& m, const double* acc, int solver, double* Rt, double* q_hist, double* qstate,
                                  int* ctrl, float tol, cudaStream_t s);
void launch_zero_doubles(double* p, size_t n, const int* ctrl, cudaStream_t s);
void launch_fill_vbo(const float* x, const float* y, const float* z, int64_t n, int64_t offset, float* vbo_pos, float* vbo_col,
                     float scene_scale, float r, float g, float b, cudaStream_t s);

// l2reg.cu (L2-distance registration of two flat mixtures, float64)
cudaError_t launch_l2_cost_grad(const double* mu_s, const double* phi_s, int Js, const double* mu_t, const double* phi_t, int Jt,
                                const double* theta, double sigma, double* out, cudaStream_t s);
cudaError_t launch_l2_bfgs(const double* mu_s, const double* phi_s, int Js, const double* mu_t, const double* phi_t, int Jt,
                           double* theta, double sigma, int max_iter, double gtol, double* out, cudaStream_t s);

}  // namespace hgmm
