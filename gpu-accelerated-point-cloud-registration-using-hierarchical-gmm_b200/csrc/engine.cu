// engine.cu -- host side of libhgmm: context, buffers, EM / registration drivers, NCCL, C ABI.
//
// Host drivers replaced (paths relative to the reference checkout):
//   GMM::solve                          src/c++/gmm_fit/gmm_kernels.cu:371-504
//   train_gmm                           src/python/gmm_waymo/src/gmm_impl.py:118-145
//   buildGMMTree                        src/python/hgmm/hgmm_gpu.py:466-548
//   GMMTree.registration                src/python/hgmm/hgmm_gpu.py:754-768
// The reference syncs the device 3x and copies weights back every EM iteration and mallocs inside
// the loop (gmm_kernels.cu:304-350,455-481); here everything is enqueued on one stream, the
// stopping rules run on the device, and the host only polls a 2-int control word.
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"
#include "kernels.h"

using namespace hgmm;

// ------------------------------------------------------------------------------------------
// NCCL through dlopen (no link-time dependency; single-GPU use never touches it)
// ------------------------------------------------------------------------------------------
namespace {
typedef struct ncclComm* ncclComm_t;
struct ncclUniqueId128 { char internal[128]; };
typedef int (*fn_ncclGetUniqueId)(ncclUniqueId128*);
typedef int (*fn_ncclCommInitRank)(ncclComm_t*, int, ncclUniqueId128, int);
typedef int (*fn_ncclCommDestroy)(ncclComm_t);
typedef int (*fn_ncclAllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
typedef const char* (*fn_ncclGetErrorString)(int);
struct NcclApi {
    void* lib = nullptr;
    fn_ncclGetUniqueId GetUniqueId = nullptr;
    fn_ncclCommInitRank CommInitRank = nullptr;
    fn_ncclCommDestroy CommDestroy = nullptr;
    fn_ncclAllReduce AllReduce = nullptr;
    fn_ncclGetErrorString GetErrorString = nullptr;
    std::string err;
};
NcclApi g_nccl;
constexpr int kNcclFloat64 = 8, kNcclSum = 0;

bool nccl_load() {
    if (g_nccl.AllReduce) return true;
    const char* env = getenv("HGMM_NCCL_LIB");
    const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
        if (!nm) continue;
        g_nccl.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.lib) break;
    }
    if (!g_nccl.lib) {
        g_nccl.err = std::string("dlopen(libnccl.so.2) failed: ") + (dlerror() ? dlerror() : "?");
        return false;
    }
    g_nccl.GetUniqueId = (fn_ncclGetUniqueId)dlsym(g_nccl.lib, "ncclGetUniqueId");
    g_nccl.CommInitRank = (fn_ncclCommInitRank)dlsym(g_nccl.lib, "ncclCommInitRank");
    g_nccl.CommDestroy = (fn_ncclCommDestroy)dlsym(g_nccl.lib, "ncclCommDestroy");
    g_nccl.AllReduce = (fn_ncclAllReduce)dlsym(g_nccl.lib, "ncclAllReduce");
    g_nccl.GetErrorString = (fn_ncclGetErrorString)dlsym(g_nccl.lib, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy) {
        g_nccl.err = "libnccl is missing a required symbol";
        g_nccl.AllReduce = nullptr;
        return false;
    }
    return true;
}

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T>
    T* as() const { return reinterpret_cast<T*>(p); }
};
}  // namespace

struct hgmm_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int num_sms = 148;
    std::string err;
    int64_t launches = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double last_ms[3] = {0, 0, 0};
    bool profiling = false;
    std::vector<cudaEvent_t> pev;   // event pairs for per-launch timing

    // points (this rank's shard), SoA
    int n = 0;
    int64_t n_total = 0;
    int64_t declared_total = 0;     // hgmm_declare_total_points: > 0 replaces the all-reduce of hgmm_set_points
    DevBuf bx, by, bz, stage;
    DevBuf sx, sy, sz, sort_scratch;   // the cloud in Morton-cell order (cloud_sort.cu), built on demand for em_flat8_kernel
    bool sorted_valid = false;

    // shared small state
    DevBuf acc;          // doubles: [kAccHdr + count*kMom]
    DevBuf ctrl;         // 8 ints
    DevBuf qstate;       // 4 doubles
    DevBuf hist;         // doubles (ll / q history)
    int* h_ctrl = nullptr;      // pinned, 8 ints
    double* h_dbl = nullptr;    // pinned, 64 doubles
    int* h_prog = nullptr;      // pinned + mapped, 4 ints: the tree build's progress words, written by the M-step kernel
    int* d_prog = nullptr;      // device view of h_prog
    float* h_params = nullptr;  // pinned, kMaxFlatJ x 13 floats: the flat fit's initial (means | covs | weights) in one copy

    // flat model
    FlatModel fm{};
    bool have_flat = false;
    DevBuf f_means, f_covs, f_weights, f_invcov, f_packed, labels, done_at, partial, rowaux, cref;

    // tree model + work
    TreeModel tm{};
    bool have_tree = false;
    DevBuf t_pi, t_mu, t_cov, t_cplx, t_packed, t_init;
    DevBuf wx[2], wy[2], wz[2], wperm[2], wpnode[2], wslot[2], wcpar[2], wcstart[2], wclen[2];
    DevBuf gbar;                // persistent level kernel: grid barrier arrival counter
    DevBuf tterm;               // adaptive build: terminal flags of two consecutive levels
    DevBuf tprof;               // HGMM_TREE_PROF=1: per-CTA phase clocks of the persistent level kernel
    DevBuf twin;                // persistent level kernel, single rank: the local "exchange" region (tree_win_layout)
    int twin_levels = 0;
    uint32_t tepoch = 16;       // tag base of the next tree level (monotonic over the context's life; identical on every rank)
    int* h_lvl = nullptr;       // pinned, 8 levels x 8 ints: the control words of every level of the last tree build
    DevBuf p_group, p_tilecnt, p_tileoff, p_segbase, p_seg0, p_seg1, p_chunkcnt, p_chunkoff, nchunks, current;

    // registration
    int nt_pts = 0;
    DevBuf tx, ty, tz, racc, Rt;
    DevBuf rx, ry, rz;          // flat registration: the target under the current transform
    bool have_target = false;
    bool have_racc = false;

    // L2 registration of two flat mixtures
    DevBuf l2buf;
    int l2_js = 0, l2_jt = 0;

    // comm
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1;
    // peer-memory exchange window of the flat M-step (xchg.cuh)
    void* xwin = nullptr;                 // this rank's window (cudaMalloc)
    void* xpeer[kXchgMaxRanks] = {};      // every rank's window as mapped here (own entry = xwin)
    bool p2p_ready = false;
    uint32_t xepoch = 0;
    int xwin_tree_levels = 0;             // deepest tree the window's tree region was sized for (0: none)
    size_t xwin_reg_off = 0;              // byte offset of the tree-registration region in the window (0: none)
    double xchg_timeout_s = 30.0;         // deadline of every in-kernel wait on a peer (HGMM_XCHG_TIMEOUT_S)
};

#define CK(call)                                                                                  \
    do {                                                                                          \
        cudaError_t e__ = (call);                                                                 \
        if (e__ != cudaSuccess) {                                                                 \
            char b__[512];                                                                        \
            snprintf(b__, sizeof b__, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            ctx->err = b__;                                                                       \
            return HGMM_ERR_CUDA;                                                                 \
        }                                                                                         \
    } while (0)

#define FAIL(code, msg)      \
    do {                     \
        ctx->err = (msg);    \
        return (code);       \
    } while (0)

static int allreduce(hgmm_ctx* ctx, double* buf, size_t count) {
    if (!ctx->comm || ctx->nranks <= 1) return HGMM_OK;
    int r = g_nccl.AllReduce(buf, buf, count, kNcclFloat64, kNcclSum, ctx->comm, ctx->stream);
    if (r != 0) {
        ctx->err = std::string("ncclAllReduce: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error");
        return HGMM_ERR_NCCL;
    }
    return HGMM_OK;
}

static void p2p_release(hgmm_ctx* ctx) {
    for (int r = 0; r < kXchgMaxRanks; ++r) {
        if (ctx->xpeer[r] && ctx->xpeer[r] != ctx->xwin) cudaIpcCloseMemHandle(ctx->xpeer[r]);
        ctx->xpeer[r] = nullptr;
    }
    if (ctx->xwin) cudaFree(ctx->xwin);
    ctx->xwin = nullptr;
    ctx->p2p_ready = false;
}

static int64_t level_base_h(int l) {
    int64_t p = 1;
    for (int i = 0; i < l; ++i) p *= 8;
    return 8 * (p - 1) / 7;
}
static int64_t level_count_h(int l) {
    int64_t p = 8;
    for (int i = 0; i < l; ++i) p *= 8;
    return p;
}

extern "C" {

const char* hgmm_version(void) { return "hgmm-b200 0.1 (sm_100a)"; }

int64_t hgmm_tree_total_nodes(int32_t max_level) { return max_level < 0 ? 0 : level_base_h(max_level); }

int hgmm_create(hgmm_ctx** out, int device, void* stream) {
    if (!out) return HGMM_ERR_INVALID;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) return HGMM_ERR_CUDA;   // no silent CPU fallback
    if (device < 0 || device >= count) return HGMM_ERR_INVALID;
    hgmm_ctx* ctx = new hgmm_ctx();
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return HGMM_ERR_CUDA; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete ctx; return HGMM_ERR_CUDA; }
    ctx->num_sms = prop.multiProcessorCount;
    if (stream) {
        ctx->stream = (cudaStream_t)stream;
    } else {
        if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return HGMM_ERR_CUDA; }
        ctx->own_stream = true;
    }
    cudaEventCreate(&ctx->ev0);
    cudaEventCreate(&ctx->ev1);
    if (cudaMallocHost((void**)&ctx->h_ctrl, 8 * sizeof(int)) != cudaSuccess ||
        cudaMallocHost((void**)&ctx->h_dbl, 64 * sizeof(double)) != cudaSuccess ||
        cudaMallocHost((void**)&ctx->h_params, (size_t)kMaxFlatJ * 13 * sizeof(float)) != cudaSuccess ||
        cudaMallocHost((void**)&ctx->h_lvl, 64 * sizeof(int)) != cudaSuccess ||
        cudaHostAlloc((void**)&ctx->h_prog, 4 * sizeof(int), cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer((void**)&ctx->d_prog, ctx->h_prog, 0) != cudaSuccess ||
        ctx->ctrl.ensure(8 * sizeof(int)) != cudaSuccess ||
        ctx->qstate.ensure(8 * sizeof(double)) != cudaSuccess || ctx->nchunks.ensure(sizeof(int)) != cudaSuccess ||
        ctx->Rt.ensure(12 * sizeof(double)) != cudaSuccess) {
        hgmm_destroy(ctx);
        return HGMM_ERR_CUDA;
    }
    if (const char* e = getenv("HGMM_XCHG_TIMEOUT_S")) {
        const double v = atof(e);
        if (v > 0.0) ctx->xchg_timeout_s = v;
    }
    *out = ctx;
    return HGMM_OK;
}

int hgmm_destroy(hgmm_ctx* ctx) {
    if (!ctx) return HGMM_OK;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->comm);
    p2p_release(ctx);
    DevBuf* all[] = {&ctx->bx, &ctx->by, &ctx->bz, &ctx->stage, &ctx->acc, &ctx->ctrl, &ctx->qstate, &ctx->hist, &ctx->f_means,
                     &ctx->f_covs, &ctx->f_weights, &ctx->f_invcov, &ctx->f_packed, &ctx->labels, &ctx->done_at, &ctx->partial, &ctx->rowaux, &ctx->cref, &ctx->t_pi, &ctx->t_mu, &ctx->t_cov,
                     &ctx->t_cplx, &ctx->t_packed, &ctx->t_init, &ctx->p_group, &ctx->p_tilecnt, &ctx->p_tileoff, &ctx->p_segbase,
                     &ctx->p_seg0, &ctx->p_seg1, &ctx->p_chunkcnt, &ctx->p_chunkoff, &ctx->nchunks, &ctx->current, &ctx->tx, &ctx->ty,
                     &ctx->tz, &ctx->racc, &ctx->Rt, &ctx->l2buf, &ctx->gbar, &ctx->twin, &ctx->rx, &ctx->ry, &ctx->rz, &ctx->tprof, &ctx->tterm,
                     &ctx->sx, &ctx->sy, &ctx->sz, &ctx->sort_scratch};
    for (DevBuf* b : all) b->release();
    for (int i = 0; i < 2; ++i) {
        DevBuf* w[] = {&ctx->wx[i], &ctx->wy[i], &ctx->wz[i], &ctx->wperm[i], &ctx->wpnode[i], &ctx->wslot[i], &ctx->wcpar[i],
                       &ctx->wcstart[i], &ctx->wclen[i]};
        for (DevBuf* b : w) b->release();
    }
    if (ctx->h_ctrl) cudaFreeHost(ctx->h_ctrl);
    if (ctx->h_dbl) cudaFreeHost(ctx->h_dbl);
    if (ctx->h_params) cudaFreeHost(ctx->h_params);
    if (ctx->h_prog) cudaFreeHost(ctx->h_prog);
    if (ctx->h_lvl) cudaFreeHost(ctx->h_lvl);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    for (cudaEvent_t e : ctx->pev) cudaEventDestroy(e);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return HGMM_OK;
}

const char* hgmm_last_error(const hgmm_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
int64_t hgmm_launch_count(const hgmm_ctx* ctx) { return ctx ? ctx->launches : 0; }
int64_t hgmm_total_points(const hgmm_ctx* ctx) { return ctx ? ctx->n_total : 0; }

// host/device AoS cloud -> device SoA
static int upload_cloud(hgmm_ctx* ctx, const float* xyz, int64_t n, int mem_kind, DevBuf& X, DevBuf& Y, DevBuf& Z,
                        const double* Rt_dev) {
    if (n < 0 || n > 2000000000LL) FAIL(HGMM_ERR_INVALID, "point count out of range");
    if (n > 0 && !xyz) FAIL(HGMM_ERR_INVALID, "null point pointer");
    CK(cudaSetDevice(ctx->device));
    const size_t fb = (size_t)(n > 0 ? n : 1) * sizeof(float);
    CK(X.ensure(fb));
    CK(Y.ensure(fb));
    CK(Z.ensure(fb));
    if (n == 0) return HGMM_OK;
    const float* src = xyz;
    if (mem_kind == HGMM_MEM_HOST) {
        CK(ctx->stage.ensure(3 * fb));
        CK(cudaMemcpyAsync(ctx->stage.p, xyz, 3 * fb, cudaMemcpyHostToDevice, ctx->stream));
        src = ctx->stage.as<float>();
    } else if (mem_kind != HGMM_MEM_DEVICE) {
        FAIL(HGMM_ERR_INVALID, "mem_kind must be HGMM_MEM_HOST or HGMM_MEM_DEVICE");
    }
    if (Rt_dev) launch_aos_to_soa_transform(src, n, Rt_dev, X.as<float>(), Y.as<float>(), Z.as<float>(), ctx->stream);
    else launch_aos_to_soa(src, n, X.as<float>(), Y.as<float>(), Z.as<float>(), ctx->stream);
    ctx->launches += 1;
    CK(cudaGetLastError());
    // A caller's DEVICE buffer is read by the conversion kernel on the context's stream: wait for it, so the buffer may be freed
    // or overwritten as soon as this call returns.  HOST input: pageable memory has been staged by the runtime when
    // cudaMemcpyAsync returns; PINNED memory is read by the copy engine later and must stay untouched until the next call that
    // synchronises (any fit / predict / register) -- documented in include/hgmm.h, the Python wrapper holds a reference.
    if (mem_kind == HGMM_MEM_DEVICE) CK(cudaStreamSynchronize(ctx->stream));
    return HGMM_OK;
}

int hgmm_set_points(hgmm_ctx* ctx, const float* xyz, int64_t n, int mem_kind) {
    if (!ctx) return HGMM_ERR_INVALID;
    ctx->sorted_valid = false;
    int r = upload_cloud(ctx, xyz, n, mem_kind, ctx->bx, ctx->by, ctx->bz, nullptr);
    if (r != HGMM_OK) return r;
    ctx->n = (int)n;
    ctx->n_total = n;
    if (ctx->comm && ctx->nranks > 1 && ctx->declared_total > 0) {
        ctx->n_total = ctx->declared_total;      // the caller sharded the cloud and knows the total: no collective, no sync
    } else if (ctx->comm && ctx->nranks > 1) {
        // total point count over ranks (the tree's pi = M0 / N_total, hgmm_gpu.py:257)
        double* d = ctx->qstate.as<double>();
        ctx->h_dbl[0] = (double)n;
        CK(cudaMemcpyAsync(d, ctx->h_dbl, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        r = allreduce(ctx, d, 1);
        if (r != HGMM_OK) return r;
        CK(cudaMemcpyAsync(ctx->h_dbl, d, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        ctx->n_total = (int64_t)(ctx->h_dbl[0] + 0.5);
    }
    return HGMM_OK;
}

int hgmm_declare_total_points(hgmm_ctx* ctx, int64_t n_total) {
    if (!ctx) return HGMM_ERR_INVALID;
    if (n_total < 0) FAIL(HGMM_ERR_INVALID, "negative total");
    ctx->declared_total = n_total;
    if (n_total > 0 && ctx->nranks > 1) ctx->n_total = n_total;
    return HGMM_OK;
}

// the cloud in Morton-cell order (cloud_sort.cu), built at most once per hgmm_set_points
static int ensure_sorted_cloud(hgmm_ctx* ctx) {
    if (ctx->sorted_valid) return HGMM_OK;
    CK(ctx->sx.ensure((size_t)ctx->n * sizeof(float)));
    CK(ctx->sy.ensure((size_t)ctx->n * sizeof(float)));
    CK(ctx->sz.ensure((size_t)ctx->n * sizeof(float)));
    CK(ctx->sort_scratch.ensure(cloud_sort_scratch_bytes(ctx->n)));
    CK(launch_cloud_sort(ctx->bx.as<float>(), ctx->by.as<float>(), ctx->bz.as<float>(), ctx->n, ctx->sx.as<float>(),
                         ctx->sy.as<float>(), ctx->sz.as<float>(), ctx->sort_scratch.p, ctx->stream));
    ctx->launches += 5;
    ctx->sorted_valid = true;
    return HGMM_OK;
}

int hgmm_sorted_points(hgmm_ctx* ctx, float* out_xyz) {
    if (!ctx) return HGMM_ERR_INVALID;
    if (!out_xyz) FAIL(HGMM_ERR_INVALID, "null output");
    if (ctx->n <= 0) FAIL(HGMM_ERR_STATE, "hgmm_set_points has not been called");
    CK(cudaSetDevice(ctx->device));
    int r = ensure_sorted_cloud(ctx);
    if (r != HGMM_OK) return r;
    std::vector<float> h((size_t)ctx->n * 3);
    const size_t nb = (size_t)ctx->n * sizeof(float);
    CK(cudaMemcpyAsync(h.data(), ctx->sx.p, nb, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(h.data() + ctx->n, ctx->sy.p, nb, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(h.data() + 2 * (size_t)ctx->n, ctx->sz.p, nb, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < ctx->n; ++i) {
        out_xyz[3 * (size_t)i] = h[i];
        out_xyz[3 * (size_t)i + 1] = h[(size_t)ctx->n + i];
        out_xyz[3 * (size_t)i + 2] = h[2 * (size_t)ctx->n + i];
    }
    return HGMM_OK;
}

// ------------------------------------------------------------------------------------------
// flat fit
// ------------------------------------------------------------------------------------------
static size_t cov_elems(int cov_type) { return cov_type == HGMM_COV_FULL ? 9 : (cov_type == HGMM_COV_DIAG ? 3 : 1); }

int hgmm_fit_flat(hgmm_ctx* ctx, const hgmm_flat_config* cfg, const float* init_means, const float* init_covs,
                  const float* init_weights, float* out_means, float* out_covs, float* out_weights, float* out_inv_cov,
                  double* out_ll, int32_t* out_iters) {
    if (!ctx) return HGMM_ERR_INVALID;
    if (!cfg || !init_means || !init_covs || !init_weights) FAIL(HGMM_ERR_INVALID, "null config / init pointer");
    if (ctx->n <= 0) FAIL(HGMM_ERR_STATE, "hgmm_set_points has not been called");
    const int J = cfg->n_components;
    if (J < 1 || J > kMaxFlatJ) FAIL(HGMM_ERR_INVALID, "n_components must be in [1, 1024] (use the tree for more)");
    if (cfg->max_iter < 0 || cfg->max_iter > 1000000) FAIL(HGMM_ERR_INVALID, "bad max_iter");
    if (cfg->flavor == HGMM_FLAVOR_CPP && cfg->cov_type != HGMM_COV_FULL) FAIL(HGMM_ERR_INVALID, "CPP flavour is full-covariance");
    if (cfg->flavor == HGMM_FLAVOR_PY && cfg->cov_type == HGMM_COV_FULL) FAIL(HGMM_ERR_INVALID, "PY flavour is diag/spherical");
    if (cfg->flavor == HGMM_FLAVOR_PY_OLD && cfg->cov_type != HGMM_COV_DIAG) FAIL(HGMM_ERR_INVALID, "PY_OLD flavour is diagonal");
    if (cfg->flavor != HGMM_FLAVOR_CPP && cfg->flavor != HGMM_FLAVOR_PY && cfg->flavor != HGMM_FLAVOR_PY_OLD)
        FAIL(HGMM_ERR_INVALID, "unknown flavor");
    CK(cudaSetDevice(ctx->device));
    const int Jp = (J + 31) / 32 * 32;
    const size_t ce = cov_elems(cfg->cov_type);
    CK(ctx->f_means.ensure((size_t)Jp * 13 * sizeof(float)));     // means [Jp,3] | covs [Jp,9] | weights [Jp]: one upload
    CK(ctx->f_invcov.ensure((size_t)Jp * 3 * sizeof(float)));
    CK(ctx->f_packed.ensure((size_t)Jp * sizeof(PackedComp)));
    const size_t acc_n = kAccHdr + (size_t)Jp * kMom;
    CK(ctx->acc.ensure(acc_n * sizeof(double)));
    CK(ctx->hist.ensure((size_t)(cfg->max_iter + 1) * sizeof(double)));
    FlatModel& m = ctx->fm;
    m.J = J; m.Jp = Jp; m.cov_type = cfg->cov_type; m.flavor = cfg->flavor; m.sigma_bug = cfg->sigma_bug; m.tol = cfg->tol;
    m.means = ctx->f_means.as<float>(); m.covs = m.means + (size_t)Jp * 3; m.weights = m.means + (size_t)Jp * 12;
    m.inv_cov = ctx->f_invcov.as<float>(); m.packed = ctx->f_packed.as<PackedComp>();
    CK(ctx->cref.ensure((size_t)(Jp / 32 + 4) * sizeof(float)));
    m.cref_blocks = ctx->cref.as<float>();
    cudaStream_t s = ctx->stream;
    // initial parameters: gathered in the pinned staging block (the previous fit ended with a stream synchronise, so it is
    // free), ONE asynchronous upload; the pack kernel also clears the control words -- 2 stream operations before the first
    // sweep instead of 7
    memcpy(ctx->h_params, init_means, (size_t)J * 3 * sizeof(float));
    memcpy(ctx->h_params + (size_t)Jp * 3, init_covs, (size_t)J * ce * sizeof(float));
    memcpy(ctx->h_params + (size_t)Jp * 12, init_weights, (size_t)J * sizeof(float));
    CK(cudaMemcpyAsync(m.means, ctx->h_params, (size_t)Jp * 13 * sizeof(float), cudaMemcpyHostToDevice, s));
    CK(ctx->done_at.ensure((size_t)(cfg->max_iter + 2) * sizeof(int)));
    int* done_at = ctx->done_at.as<int>();
    const bool need_acc = cfg->reserved == 1 || cfg->reserved == 3 || (ctx->nranks > 1 && !ctx->p2p_ready);
    if (need_acc) CK(cudaMemsetAsync(ctx->acc.p, 0, acc_n * sizeof(double), s));
    launch_flat_pack_init(m, ctx->ctrl.as<int>(), done_at, cfg->max_iter + 2, s);
    ctx->launches += 1;
    // kernel variant: reserved == 1 selects the first-generation two-phase kernel (fp64 atomics), else the
    // register-resident single-evaluation kernel with deterministic partial rows
    const bool v1 = cfg->reserved == 1;
    const bool v3 = !v1 && cfg->reserved != 2 && Jp >= 64;      // packed-FP32 kernel needs at least one component pair per lane
    const int tile = flat_pick_tile(ctx->n, ctx->num_sms, cfg->tile_points);
    int JT = 1, W = 8, Sdiv = 1, G = 1, grid = 1, big = 0;
    if (v3) {
        flat3_plan(ctx->n, Jp, ctx->num_sms, (cfg->tile_points >= 1 && cfg->tile_points <= 12) ? cfg->tile_points : 0, &W, &Sdiv, &G, &grid, &big);
    } else if (!v1) {
        flat2_plan(ctx->n, Jp, ctx->num_sms, cfg->tile_points == 1, &JT, &W, &Sdiv, &G, &grid, &big);
    }
    const float *fx = ctx->bx.as<float>(), *fy = ctx->by.as<float>(), *fz = ctx->bz.as<float>();
    if (v3 && big >= 8 && big <= 11) {
        // em_flat8_kernel sums about one origin per CTA: it reads the cloud in Morton-cell order (built once per hgmm_set_points)
        int rs = ensure_sorted_cloud(ctx);
        if (rs != HGMM_OK) return rs;
        fx = ctx->sx.as<float>(); fy = ctx->sy.as<float>(); fz = ctx->sz.as<float>();
    }
    if (!v1) {
        CK(ctx->partial.ensure((size_t)grid * G * kMom * Jp * sizeof(float)));
        CK(ctx->rowaux.ensure((size_t)grid * G * 2 * sizeof(double)));
    }
    CK(cudaEventRecord(ctx->ev0, s));
    const bool prof = ctx->profiling;
    if (prof) {
        while ((int)ctx->pev.size() < 2 * cfg->max_iter) {
            cudaEvent_t e;
            CK(cudaEventCreate(&e));
            ctx->pev.push_back(e);
        }
    }
    for (int it = 0; it < cfg->max_iter; ++it) {
        if (v1) CK(cudaMemsetAsync(ctx->acc.p, 0, acc_n * sizeof(double), s));
        if (prof) CK(cudaEventRecord(ctx->pev[2 * it], s));
        if (v1) {
            CK(launch_em_flat(ctx->bx.as<float>(), ctx->by.as<float>(), ctx->bz.as<float>(), ctx->n, m, ctx->acc.as<double>(),
                              done_at + it, ctx->num_sms, tile, s));
            ctx->launches += 1;
        } else {
            if (v3)
                CK(launch_em_flat3(fx, fy, fz, ctx->n, m, m.cref_blocks, W, Sdiv, G,
                                   grid, big, ctx->partial.as<float>(), ctx->rowaux.as<double>(), done_at + it, s));
            else
                CK(launch_em_flat2(ctx->bx.as<float>(), ctx->by.as<float>(), ctx->bz.as<float>(), ctx->n, m, m.cref_blocks, JT, W, Sdiv,
                                   G, grid, big, ctx->partial.as<float>(), ctx->rowaux.as<double>(), done_at + it, s));
            if (prof) CK(cudaEventRecord(ctx->pev[2 * it + 1], s));
            if (ctx->nranks > 1 && ctx->p2p_ready && cfg->reserved != 3) {
                // multi-GPU over peer memory: reduce + NVLink exchange + finalize in one kernel (xchg.cuh)
                XchgView xv;
                for (int r = 0; r < kXchgMaxRanks; ++r) xv.data[r] = reinterpret_cast<uint4*>(ctx->xpeer[r]);
                xv.rank = ctx->rank; xv.nranks = ctx->nranks; xv.epoch = ++ctx->xepoch;
                xv.timeout_ns = (unsigned long long)(ctx->xchg_timeout_s * 1e9);
                CK(launch_flat_reduce_exchange_finalize(m, ctx->partial.as<float>(), ctx->rowaux.as<double>(), grid * G,
                                                        ctx->ctrl.as<int>(), done_at, it, ctx->hist.as<double>(),
                                                        (double)ctx->n_total, xv, s));
                ctx->launches += 2;
                continue;
            }
            if (ctx->nranks <= 1 && cfg->reserved != 3) {      // single GPU: reduce + finalize in one kernel
                launch_flat_reduce_finalize(m, ctx->partial.as<float>(), ctx->rowaux.as<double>(), grid * G, ctx->ctrl.as<int>(),
                                            done_at, it, ctx->hist.as<double>(), (double)ctx->n_total, s);
                ctx->launches += 2;
                continue;
            }
            CK(launch_flat_reduce(ctx->partial.as<float>(), ctx->rowaux.as<double>(), grid * G, m, ctx->acc.as<double>(),
                                  done_at + it, s));
            ctx->launches += 2;
        }
        if (prof && v1) CK(cudaEventRecord(ctx->pev[2 * it + 1], s));
        int r = allreduce(ctx, ctx->acc.as<double>(), kAccHdr + (size_t)J * kMom);
        if (r != HGMM_OK) return r;
        launch_flat_finalize(m, ctx->acc.as<double>(), ctx->ctrl.as<int>(), done_at, it, ctx->hist.as<double>(),
                             (double)ctx->n_total, s);
        ctx->launches += 1;
    }
    CK(cudaEventRecord(ctx->ev1, s));
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(ctx->h_ctrl, ctx->ctrl.p, 8 * sizeof(int), cudaMemcpyDeviceToHost, s));
    const bool want_model = out_means || out_covs || out_weights;      // one download into the pinned block, split on the host
    if (want_model) CK(cudaMemcpyAsync(ctx->h_params, m.means, (size_t)Jp * 13 * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (out_inv_cov && cfg->flavor != HGMM_FLAVOR_CPP)
        CK(cudaMemcpyAsync(out_inv_cov, m.inv_cov, (size_t)J * (ce == 1 ? 1 : 3) * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (out_ll && cfg->max_iter > 0)
        CK(cudaMemcpyAsync(out_ll, ctx->hist.p, (size_t)cfg->max_iter * sizeof(double), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (out_means) memcpy(out_means, ctx->h_params, (size_t)J * 3 * sizeof(float));
    if (out_covs) memcpy(out_covs, ctx->h_params + (size_t)Jp * 3, (size_t)J * ce * sizeof(float));
    if (out_weights) memcpy(out_weights, ctx->h_params + (size_t)Jp * 12, (size_t)J * sizeof(float));
    if (out_iters) *out_iters = ctx->h_ctrl[1];
    if (ctx->h_ctrl[7]) FAIL(HGMM_ERR_NCCL, "peer-memory exchange timed out: a rank is missing or the ranks' call sequences differ");
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    ctx->last_ms[0] = ms; ctx->last_ms[1] = 0; ctx->last_ms[2] = 0;
    if (prof) {
        double sum = 0.0;
        for (int it = 0; it < cfg->max_iter; ++it) {
            float k = 0.f;
            cudaEventElapsedTime(&k, ctx->pev[2 * it], ctx->pev[2 * it + 1]);
            sum += k;
        }
        ctx->last_ms[1] = sum;
        ctx->last_ms[2] = cfg->max_iter;
    }
    ctx->have_flat = true;
    return HGMM_OK;
}

int hgmm_predict_flat(hgmm_ctx* ctx, const float* xyz, int64_t n, int mem_kind, int32_t* labels) {
    if (!ctx) return HGMM_ERR_INVALID;
    if (!ctx->have_flat) FAIL(HGMM_ERR_STATE, "no flat model: call hgmm_fit_flat first");
    if (!labels) FAIL(HGMM_ERR_INVALID, "null labels");
    CK(cudaSetDevice(ctx->device));
    const float *x, *y, *z;
    int np;
    if (xyz) {
        int r = upload_cloud(ctx, xyz, n, mem_kind, ctx->tx, ctx->ty, ctx->tz, nullptr);
        if (r != HGMM_OK) return r;
        ctx->have_target = false;
        x = ctx->tx.as<float>(); y = ctx->ty.as<float>(); z = ctx->tz.as<float>();
        np = (int)n;
    } else {
        x = ctx->bx.as<float>(); y = ctx->by.as<float>(); z = ctx->bz.as<float>();
        np = ctx->n;
    }
    if (np <= 0) return HGMM_OK;
    CK(ctx->labels.ensure((size_t)np * sizeof(int32_t)));
    CK(launch_predict(x, y, z, np, ctx->fm.packed, ctx->fm.Jp, ctx->labels.as<int32_t>(), ctx->num_sms, ctx->stream));
    ctx->launches += 1;
    CK(cudaMemcpyAsync(labels, ctx->labels.p, (size_t)np * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return HGMM_OK;
}

// ------------------------------------------------------------------------------------------
// tree
// ------------------------------------------------------------------------------------------
static int ensure_tree_model(hgmm_ctx* ctx, int L) {
    if (L < 1 || L > 6) FAIL(HGMM_ERR_INVALID, "max_level must be in [1, 6]");
    const int64_t nt = level_base_h(L);
    CK(ctx->t_pi.ensure((size_t)nt * sizeof(float)));
    CK(ctx->t_mu.ensure((size_t)nt * 3 * sizeof(float)));
    CK(ctx->t_cov.ensure((size_t)nt * 9 * sizeof(float)));
    CK(ctx->t_cplx.ensure((size_t)nt * sizeof(float)));
    CK(ctx->t_packed.ensure((size_t)nt * sizeof(PackedComp)));
    TreeModel& t = ctx->tm;
    t.L = L; t.nt = (int)nt;
    t.pi = ctx->t_pi.as<float>(); t.mu = ctx->t_mu.as<float>(); t.cov = ctx->t_cov.as<float>();
    t.cplx = ctx->t_cplx.as<float>(); t.packed = ctx->t_packed.as<PackedComp>();
    return HGMM_OK;
}

int hgmm_fit_tree(hgmm_ctx* ctx, const hgmm_tree_config* cfg, const float* init_means, float* out_pi, float* out_mu,
                  float* out_cov, int64_t* out_current, int32_t* out_iters, double* out_q) {
    if (!ctx) return HGMM_ERR_INVALID;
    if (!cfg || !init_means) FAIL(HGMM_ERR_INVALID, "null config / init_means");
    if (ctx->n <= 0) FAIL(HGMM_ERR_STATE, "hgmm_set_points has not been called");
    CK(cudaSetDevice(ctx->device));
    const int L = cfg->max_level;
    int r = ensure_tree_model(ctx, L);
    if (r != HGMM_OK) return r;
    TreeModel& t = ctx->tm;
    const int n = ctx->n;
    cudaStream_t s = ctx->stream;
    int chunk = cfg->chunk_points;
    if (chunk <= 0) {
        // one warp per chunk: aim at ~32 chunk-warps per SM; tiny clouds are latency-bound, so the floor is one point per lane
        chunk = (int)(((int64_t)n / ((int64_t)ctx->num_sms * 32) + 31) / 32 * 32);
        if (chunk < 32) chunk = 32;
        if (chunk > 512) chunk = 512;
    }
    if (chunk % 32 != 0 || chunk > 65536) FAIL(HGMM_ERR_INVALID, "chunk_points must be a multiple of 32");
    const int max_iters = cfg->max_iters_per_level > 0 ? cfg->max_iters_per_level : 10000;

    // buffers
    const int64_t leaf_parents = L >= 2 ? level_count_h(L - 2) : 1;       // parents of the deepest partition's output
    const int64_t max_parents = L >= 2 ? level_count_h(L - 2) : 1;        // segments entering the last partition
    const int64_t max_newseg = 8 * max_parents;
    const int64_t max_chunks = (n + chunk - 1) / chunk + max_newseg + 8;
    (void)leaf_parents;
    for (int i = 0; i < 2; ++i) {
        CK(ctx->wx[i].ensure((size_t)n * 4)); CK(ctx->wy[i].ensure((size_t)n * 4)); CK(ctx->wz[i].ensure((size_t)n * 4));
        CK(ctx->wperm[i].ensure((size_t)n * 4)); CK(ctx->wpnode[i].ensure((size_t)n * 4)); CK(ctx->wslot[i].ensure((size_t)n + 64));
        CK(ctx->wcpar[i].ensure((size_t)max_chunks * 4)); CK(ctx->wcstart[i].ensure((size_t)max_chunks * 4));
        CK(ctx->wclen[i].ensure((size_t)max_chunks * 4));
    }
    const int n_tiles = (n + 1023) / 1024;
    CK(ctx->p_group.ensure((size_t)n_tiles * 32 * 8 * sizeof(uint16_t)));
    CK(ctx->p_tilecnt.ensure((size_t)n_tiles * 8 * sizeof(uint32_t)));
    CK(ctx->p_tileoff.ensure((size_t)(n_tiles + 1) * 8 * sizeof(uint32_t)));
    CK(ctx->p_segbase.ensure((size_t)(max_parents + 1) * 8 * sizeof(uint32_t)));
    CK(ctx->p_seg0.ensure((size_t)(max_newseg + 1) * sizeof(int)));
    CK(ctx->p_seg1.ensure((size_t)(max_newseg + 1) * sizeof(int)));
    CK(ctx->p_chunkcnt.ensure((size_t)max_newseg * sizeof(int)));
    CK(ctx->p_chunkoff.ensure((size_t)(max_newseg + 1) * sizeof(int)));
    const size_t acc_n = kAccHdr + (size_t)level_count_h(L - 1) * kMom;
    CK(ctx->acc.ensure(2 * acc_n * sizeof(double)));          // two parities for the persistent level kernel
    CK(ctx->t_init.ensure((size_t)t.nt * 3 * sizeof(float)));

    TreeWork w[2];
    for (int i = 0; i < 2; ++i) {
        w[i].x = ctx->wx[i].as<float>(); w[i].y = ctx->wy[i].as<float>(); w[i].z = ctx->wz[i].as<float>();
        w[i].perm = ctx->wperm[i].as<int>(); w[i].pnode = ctx->wpnode[i].as<int>(); w[i].slot = ctx->wslot[i].as<uint8_t>();
        w[i].chunk_parent = ctx->wcpar[i].as<int>(); w[i].chunk_start = ctx->wcstart[i].as<int>(); w[i].chunk_len = ctx->wclen[i].as<int>();
    }
    PartitionScratch ps;
    ps.group_off = ctx->p_group.as<uint16_t>(); ps.tile_cnt = ctx->p_tilecnt.as<uint32_t>(); ps.tile_off = ctx->p_tileoff.as<uint32_t>();
    ps.seg_base = ctx->p_segbase.as<uint32_t>(); ps.seg_start = ctx->p_seg0.as<int>(); ps.new_seg_start = ctx->p_seg1.as<int>();
    ps.chunk_cnt = ctx->p_chunkcnt.as<int>(); ps.chunk_off = ctx->p_chunkoff.as<int>();
    int* nchunks_dev = ctx->nchunks.as<int>();
    int* ctrl = ctx->ctrl.as<int>();
    double* acc = ctx->acc.as<double>();
    double* qstate = ctx->qstate.as<double>();

    CK(cudaEventRecord(ctx->ev0, s));
    CK(cudaMemcpyAsync(ctx->t_init.p, init_means, (size_t)t.nt * 3 * sizeof(float), cudaMemcpyHostToDevice, s));
    launch_tree_init(t, ctx->t_init.as<float>(), cfg->sig2, s);
    CK(cudaMemcpyAsync(w[0].x, ctx->bx.p, (size_t)n * 4, cudaMemcpyDeviceToDevice, s));
    CK(cudaMemcpyAsync(w[0].y, ctx->by.p, (size_t)n * 4, cudaMemcpyDeviceToDevice, s));
    CK(cudaMemcpyAsync(w[0].z, ctx->bz.p, (size_t)n * 4, cudaMemcpyDeviceToDevice, s));
    launch_iota(w[0].perm, n, s);
    CK(cudaMemsetAsync(w[0].pnode, 0, (size_t)n * 4, s));
    CK(cudaMemsetAsync(acc, 0, acc_n * sizeof(double), s));
    CK(launch_root_chunks(w[0], n, ps, chunk, nchunks_dev, s));
    ctx->launches += 3;

    int cur = 0;
    int64_t n_parents = 1;
    const int batch = 8;      // iterations enqueued between polls of the control word (no-ops once converged)
    const bool fast_ll = cfg->ll_mode == HGMM_LL_ESTEP;
    CK(ctx->done_at.ensure((size_t)(max_iters + batch + 4) * sizeof(int)));
    int* done_at = ctx->done_at.as<int>();
    // Default (fast log-likelihood mode): ONE persistent cooperative kernel per level (tree_level.cuh) -- on several ranks with
    // the reduce-scatter / all-gather of the level's statistics fused in over peer memory, no NCCL call, no host round trip.
    // The multi-kernel path below stays for HGMM_LL_LEVEL (the reference's whole-level scan), the scalar A/B kernel, ranks
    // without peer-memory windows, and HGMM_TREE_LEGACY=1.
    static const bool legacy_env = getenv("HGMM_TREE_LEGACY") && getenv("HGMM_TREE_LEGACY")[0] == '1';
    const bool persist = fast_ll && cfg->reserved == 0 && !legacy_env &&
                         (ctx->nranks <= 1 || (ctx->p2p_ready && L <= ctx->xwin_tree_levels));
    const bool adaptive = cfg->prune_lambda_c > 0.f || cfg->prune_min_points > 0.f;
    if (adaptive && !persist)
        FAIL(HGMM_ERR_INVALID, "the adaptive (pruned) build runs in the persistent level kernel: ll_mode = HGMM_LL_ESTEP, reserved = 0 "
                               "(and peer-memory windows when several ranks are attached)");
    uint8_t* term[2] = {nullptr, nullptr};
    if (adaptive) {
        CK(ctx->tterm.ensure(2 * (size_t)level_count_h(L - 1)));
        term[0] = ctx->tterm.as<uint8_t>();
        term[1] = term[0] + level_count_h(L - 1);
    }
    TreeXchgHost xh{};
    if (persist) {
        xh.rank = ctx->rank; xh.nranks = ctx->nranks;
        xh.timeout_ns = (unsigned long long)(ctx->xchg_timeout_s * 1e9);
        if (ctx->nranks <= 1) {
            xh.rank = 0; xh.nranks = 1;
            const TreeWinLayout wl = tree_win_layout(L);
            if (ctx->twin_levels < L) {
                CK(ctx->twin.ensure(wl.bytes));
                CK(cudaMemsetAsync(ctx->twin.p, 0, wl.bytes, s));
                ctx->twin_levels = L;
            }
            const TreeWinLayout cur_l = tree_win_layout(ctx->twin_levels);
            char* base = ctx->twin.as<char>();
            xh.pk[0] = base + cur_l.pk_off; xh.fin[0] = base + cur_l.fin_off; xh.mom[0] = base + cur_l.mom_off; xh.ll[0] = base + cur_l.ll_off;
            xh.mom_cap = cur_l.mom_cap;
        } else {
            const TreeWinLayout wl = tree_win_layout(ctx->xwin_tree_levels);
            for (int r = 0; r < ctx->nranks; ++r) {
                char* base = static_cast<char*>(ctx->xpeer[r]) + kXchgBytes;
                xh.pk[r] = base + wl.pk_off; xh.fin[r] = base + wl.fin_off; xh.mom[r] = base + wl.mom_off; xh.ll[r] = base + wl.ll_off;
            }
            xh.mom_cap = wl.mom_cap;
        }
        CK(ctx->gbar.ensure(4 * sizeof(unsigned)));
        static const bool prof_env = getenv("HGMM_TREE_PROF") && getenv("HGMM_TREE_PROF")[0] == '1';
        if (prof_env) {            // per-phase SM-clock totals of every CTA, dumped to stderr after the build (diagnostic switch)
            CK(ctx->tprof.ensure((size_t)ctx->num_sms * 8 * sizeof(long long)));
            CK(cudaMemsetAsync(ctx->tprof.p, 0, (size_t)ctx->num_sms * 8 * sizeof(long long), s));
        }
    }
    long long* tprof = (persist && ctx->tprof.p) ? ctx->tprof.as<long long>() : nullptr;
    for (int l = 0; l < L && persist; ++l) {
        const int64_t cnt = level_count_h(l);
        const size_t stride = kAccHdr + (size_t)cnt * kMom;
        // ctrl[7] (abort) is cleared once per build and is sticky across its levels: after a failed wait on a peer the
        // remaining level kernels return at once instead of each running into the deadline
        CK(cudaMemsetAsync(ctrl, 0, (l == 0 ? 8 : 7) * sizeof(int), s));
        CK(cudaMemsetAsync(qstate, 0, 8 * sizeof(double), s));        // [0] prevQ [1] last q, [4..6] the kernel's log-likelihood slots
        CK(cudaMemsetAsync(ctx->gbar.p, 0, 4 * sizeof(unsigned), s));
        CK(cudaMemsetAsync(acc, 0, 2 * stride * sizeof(double), s));
        xh.base = ctx->tepoch;
        ctx->tepoch += (uint32_t)max_iters + 2u;
        cudaError_t e = launch_tree_level(w[cur], t, l, n, acc, stride, nchunks_dev, (double)ctx->n_total, cfg->ld, cfg->ls, max_iters,
                                          ctrl, qstate, ctx->gbar.as<unsigned>(), chunk, xh, ctx->num_sms, tprof,
                                          (adaptive && l > 0) ? term[(l - 1) & 1] : nullptr, s);
        if (e != cudaSuccess) { ctx->err = std::string("tree level kernel launch: ") + cudaGetErrorString(e); return HGMM_ERR_CUDA; }
        ctx->launches += 1;
        if (adaptive && l < L - 1) {       // which nodes of this level are terminal: their points sit out the deeper levels
            launch_tree_prune(t, l, (double)ctx->n_total, cfg->prune_lambda_c, cfg->prune_min_points, term[l & 1], s);
            ctx->launches += 1;
        }
        // no host synchronisation between levels: the control words of every level are fetched asynchronously, read at the end
        CK(cudaMemcpyAsync(ctx->h_lvl + 8 * l, ctrl, 8 * sizeof(int), cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(ctx->h_dbl + 32 + 2 * l, qstate, 2 * sizeof(double), cudaMemcpyDeviceToHost, s));
        if (l < L - 1) {
            CK(launch_partition(w[cur], w[1 - cur], n, (int)n_parents, ps, chunk, nchunks_dev, s));
            ctx->launches += 7;
            cur = 1 - cur;
            n_parents *= 8;
        }
    }
    if (persist) {
        CK(cudaStreamSynchronize(s));
        if (tprof) {
            std::vector<long long> h((size_t)ctx->num_sms * 8);
            CK(cudaMemcpy(h.data(), tprof, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
            double tot[6] = {0, 0, 0, 0, 0, 0};
            for (int b = 0; b < ctx->num_sms; ++b)
                for (int k = 0; k < 6; ++k) tot[k] += (double)h[(size_t)b * 8 + k] / ctx->num_sms;
            fprintf(stderr, "HGMM_TREE_PROF mean SM cycles per CTA over the build: E %.0f | fold+flush %.0f | barrier %.0f | exchange+M %.0f | verdict %.0f\n",
                    tot[0], tot[1], tot[2], tot[3], tot[4]);
        }
        for (int l = 0; l < L; ++l) {
            if (ctx->h_lvl[8 * l + 7])
                FAIL(HGMM_ERR_NCCL, "tree level kernel: a wait on a peer rank timed out (rank missing, or the ranks' call sequences differ)");
        }
        for (int l = 0; l < L; ++l) {
            if (!ctx->h_lvl[8 * l]) FAIL(HGMM_ERR_CUDA, "tree level kernel ended without a verdict");
            if (out_iters) out_iters[l] = ctx->h_lvl[8 * l + 1];
            if (out_q) out_q[l] = ctx->h_dbl[32 + 2 * l + 1];
        }
    }
    for (int l = 0; l < L && !persist; ++l) {
        CK(cudaMemsetAsync(ctrl, 0, 8 * sizeof(int), s));
        CK(cudaMemsetAsync(qstate, 0, 4 * sizeof(double), s));
        CK(cudaMemsetAsync(done_at, 0, (size_t)(max_iters + batch + 4) * sizeof(int), s));
        const int64_t cnt = level_count_h(l);
        const int chunks_bound = (int)((n + chunk - 1) / chunk + (l == 0 ? 0 : cnt / 8));
        const PackedComp* level_packed = t.packed + level_base_h(l);
        bool done = false;
        int it = 0;
        auto enqueue_iteration = [&](int* prog) -> int {
            cudaError_t e = launch_tree_estep(w[cur], t, l, acc, chunks_bound, nchunks_dev, done_at + it, cfg->reserved == 1, s);
            if (e != cudaSuccess) { ctx->err = std::string("tree E-step launch: ") + cudaGetErrorString(e); return HGMM_ERR_CUDA; }
            int rr = allreduce(ctx, acc, kAccHdr + (size_t)cnt * kMom);
            if (rr != HGMM_OK) return rr;
            launch_tree_mstep(t, l, acc, (double)ctx->n_total, cfg->ld, ctrl, done_at, it, fast_ll ? 1 : 0, qstate, cfg->ls, max_iters,
                              prog, s);
            ctx->launches += 2;
            if (!fast_ll) {
                launch_tree_zero_ll(acc, done_at + it, s);
                e = launch_level_ll(ctx->bx.as<float>(), ctx->by.as<float>(), ctx->bz.as<float>(), n, level_packed, (int)cnt, acc,
                                    done_at + it, ctx->num_sms, s);
                if (e != cudaSuccess) { ctx->err = std::string("level log-likelihood launch: ") + cudaGetErrorString(e); return HGMM_ERR_CUDA; }
                rr = allreduce(ctx, acc, 1);
                if (rr != HGMM_OK) return rr;
                launch_tree_converge(acc, ctrl, done_at, it, qstate, cfg->ls, max_iters, prog, s);
                ctx->launches += 3;
            }
            ++it;
            return HGMM_OK;
        };
        if (ctx->nranks <= 1) {
            // single rank: no stream synchronise inside the level.  The kernel that evaluates the stopping rule also writes
            // (converged, iterations retired) to host-mapped memory; the host keeps at most `ahead` iterations in flight and
            // stops enqueuing when it sees the flag (iterations already enqueued are no-ops through done_at).
            volatile int* prog = ctx->h_prog;
            prog[0] = 0;
            prog[1] = 0;
            const int ahead = 6;
            // watchdog: "no new iteration retired for 120 s" (re-armed on every advance of prog[0]) or a dead stream --
            // never a bound on the level's total run time; the wait yields the core instead of spinning flat out
            auto t_progress = std::chrono::steady_clock::now();
            int seen = 0;
            while (true) {
                bool stalled = false;
                unsigned polls = 0;
                while (it - prog[0] >= ahead && !prog[1]) {
                    if (prog[0] != seen) { seen = prog[0]; t_progress = std::chrono::steady_clock::now(); }
                    if ((++polls & 1023u) == 0u) {
                        std::this_thread::yield();
                        const cudaError_t q = cudaStreamQuery(s);
                        if (q != cudaSuccess && q != cudaErrorNotReady) { stalled = true; break; }
                        if (std::chrono::steady_clock::now() - t_progress > std::chrono::seconds(120)) { stalled = true; break; }
                    }
                }
                if (stalled || prog[1] || it >= max_iters + batch) break;
                r = enqueue_iteration(ctx->d_prog);
                if (r != HGMM_OK) return r;
            }
            CK(cudaMemcpyAsync(ctx->h_ctrl, ctrl, 2 * sizeof(int), cudaMemcpyDeviceToHost, s));
            CK(cudaStreamSynchronize(s));
            done = ctx->h_ctrl[0] != 0;
            if (!done) FAIL(HGMM_ERR_CUDA, "tree build made no progress (device error?)");
        }
        while (!done) {       // several ranks: every rank must enqueue the same number of all-reduces -> fixed batches
            for (int b = 0; b < batch; ++b) {
                r = enqueue_iteration(nullptr);
                if (r != HGMM_OK) return r;
            }
            CK(cudaMemcpyAsync(ctx->h_ctrl, ctrl, 2 * sizeof(int), cudaMemcpyDeviceToHost, s));
            CK(cudaStreamSynchronize(s));
            done = ctx->h_ctrl[0] != 0;
        }
        if (out_iters) out_iters[l] = ctx->h_ctrl[1];
        if (out_q) {
            CK(cudaMemcpyAsync(ctx->h_dbl, qstate, 2 * sizeof(double), cudaMemcpyDeviceToHost, s));
            CK(cudaStreamSynchronize(s));
            out_q[l] = ctx->h_dbl[1];
        }
        if (l < L - 1) {
            CK(launch_partition(w[cur], w[1 - cur], n, (int)n_parents, ps, chunk, nchunks_dev, s));
            ctx->launches += 7;
            cur = 1 - cur;
            n_parents *= 8;
        }
    }
    launch_tree_cplx(t, s);
    ctx->launches += 1;
    if (out_current) {
        CK(ctx->current.ensure((size_t)n * sizeof(int64_t)));
        launch_tree_current(w[cur], n, L - 1, ctx->current.as<int64_t>(), s);
        ctx->launches += 1;
        CK(cudaMemcpyAsync(out_current, ctx->current.p, (size_t)n * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    }
    CK(cudaEventRecord(ctx->ev1, s));
    if (out_pi) CK(cudaMemcpyAsync(out_pi, t.pi, (size_t)t.nt * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (out_mu) CK(cudaMemcpyAsync(out_mu, t.mu, (size_t)t.nt * 3 * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (out_cov) CK(cudaMemcpyAsync(out_cov, t.cov, (size_t)t.nt * 9 * sizeof(float), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    ctx->last_ms[0] = ms; ctx->last_ms[1] = ms; ctx->last_ms[2] = 0;
    ctx->have_tree = true;
    return HGMM_OK;
}

int hgmm_tree_set_model(hgmm_ctx* ctx, int32_t max_level, const float* pi, const float* mu, const float* cov) {
    if (!ctx) return HGMM_ERR_INVALID;
    if (!pi || !mu || !cov) FAIL(HGMM_ERR_INVALID, "null model pointer");
    CK(cudaSetDevice(ctx->device));
    int r = ensure_tree_model(ctx, max_level);
    if (r != HGMM_OK) return r;
    TreeModel& t = ctx->tm;
    cudaStream_t s = ctx->stream;
    CK(cudaMemcpyAsync(t.pi, pi, (size_t)t.nt * sizeof(float), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(t.mu, mu, (size_t)t.nt * 3 * sizeof(float), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(t.cov, cov, (size_t)t.nt * 9 * sizeof(float), cudaMemcpyHostToDevice, s));
    launch_tree_pack_all(t, s);
    ctx->launches += 1;
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    ctx->have_tree = true;
    return HGMM_OK;
}

// ------------------------------------------------------------------------------------------
// registration
// ------------------------------------------------------------------------------------------
int hgmm_reg_set_target(hgmm_ctx* ctx, const float* xyz, int64_t n, int mem_kind) {
    if (!ctx) return HGMM_ERR_INVALID;
    int r = upload_cloud(ctx, xyz, n, mem_kind, ctx->tx, ctx->ty, ctx->tz, nullptr);
    if (r != HGMM_OK) return r;
    ctx->nt_pts = (int)n;
    ctx->have_target = true;
    return HGMM_OK;
}

static int upload_Rt(hgmm_ctx* ctx, const double* rot, const double* t) {
    for (int k = 0; k < 9; ++k) ctx->h_dbl[k] = rot[k];
    for (int k = 0; k < 3; ++k) ctx->h_dbl[9 + k] = t[k];
    CK(cudaMemcpyAsync(ctx->Rt.p, ctx->h_dbl, 12 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    // h_dbl is reused right away by callers only after a sync; make the copy safe
    CK(cudaStreamSynchronize(ctx->stream));
    return HGMM_OK;
}

int hgmm_reg_estep(hgmm_ctx* ctx, const double* rot, const double* t, float lambda_c, double* out_m0, double* out_m1,
                   double* out_m2) {
    if (!ctx) return HGMM_ERR_INVALID;
    if (!ctx->have_tree) FAIL(HGMM_ERR_STATE, "no tree: call hgmm_fit_tree / hgmm_tree_set_model first");
    if (!ctx->have_target) FAIL(HGMM_ERR_STATE, "no target: call hgmm_reg_set_target first");
    if (!rot || !t) FAIL(HGMM_ERR_INVALID, "null transform");
    CK(cudaSetDevice(ctx->device));
    const TreeModel& tm = ctx->tm;
    const size_t rn = (size_t)tm.nt * 10;
    CK(ctx->racc.ensure(rn * sizeof(double)));
    int r = upload_Rt(ctx, rot, t);
    if (r != HGMM_OK) return r;
    cudaStream_t s = ctx->stream;
    CK(cudaMemsetAsync(ctx->ctrl.p, 0, 8 * sizeof(int), s));
    CK(cudaMemsetAsync(ctx->racc.p, 0, rn * sizeof(double), s));
    CK(launch_reg_estep(ctx->tx.as<float>(), ctx->ty.as<float>(), ctx->tz.as<float>(), ctx->nt_pts, ctx->Rt.as<double>(), tm, lambda_c,
                        ctx->racc.as<double>(), out_m2 ? 1 : 0, ctx->ctrl.as<int>(), s));
    ctx->launches += 1;
    r = allreduce(ctx, ctx->racc.as<double>(), rn);
    if (r != HGMM_OK) return r;
    std::vector<double> h(rn);
    CK(cudaMemcpyAsync(h.data(), ctx->racc.p, rn * sizeof(double), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    for (int i = 0; i < tm.nt; ++i) {
        const double* A = h.data() + (size_t)i * 10;
        if (out_m0) out_m0[i] = A[0];
        if (out_m1) { out_m1[3 * i] = A[1]; out_m1[3 * i + 1] = A[2]; out_m1[3 * i + 2] = A[3]; }
        if (out_m2) {
            double* M = out_m2 + 9 * (size_t)i;
            M[0] = A[4]; M[1] = M[3] = A[5]; M[2] = M[6] = A[6]; M[4] = A[7]; M[5] = M[7] = A[8]; M[8] = A[9];
        }
    }
    ctx->have_racc = true;
    return HGMM_OK;
}

int hgmm_reg_mstep(hgmm_ctx* ctx, int32_t solver, double* rot, double* t, double* out_q) {
    if (!ctx) return HGMM_ERR_INVALID;
    if (!ctx->have_racc) FAIL(HGMM_ERR_STATE, "no moments: call hgmm_reg_estep first");
    if (!rot || !t) FAIL(HGMM_ERR_INVALID, "null transform");
    if (solver != HGMM_SOLVER_TWIST_LSTSQ && solver != HGMM_SOLVER_PROCRUSTES) FAIL(HGMM_ERR_INVALID, "unknown solver");
    CK(cudaSetDevice(ctx->device));
    int r = upload_Rt(ctx, rot, t);
    if (r != HGMM_OK) return r;
    cudaStream_t s = ctx->stream;
    CK(ctx->hist.ensure(64 * sizeof(double)));
    CK(cudaMemsetAsync(ctx->ctrl.p, 0, 8 * sizeof(int), s));
    CK(cudaMemsetAsync(ctx->qstate.p, 0, 4 * sizeof(double), s));
    CK(launch_reg_solve(ctx->tm, ctx->racc.as<double>(), 0, solver, ctx->Rt.as<double>(), ctx->hist.as<double>(),
                        ctx->qstate.as<double>(), ctx->ctrl.as<int>(), 0.f, nullptr, s));      // hgmm_reg_estep has already summed the ranks
    ctx->launches += 1;
    CK(cudaMemcpyAsync(ctx->h_dbl, ctx->Rt.p, 12 * sizeof(double), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(ctx->h_dbl + 16, ctx->qstate.p, 4 * sizeof(double), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(ctx->h_ctrl, ctx->ctrl.p, 4 * sizeof(int), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    for (int k = 0; k < 9; ++k) rot[k] = ctx->h_dbl[k];
    for (int k = 0; k < 3; ++k) t[k] = ctx->h_dbl[9 + k];
    if (out_q) *out_q = ctx->h_dbl[17];
    if (ctx->h_ctrl[2]) {
        std::string msg = "registration solve: singular / non-positive-definite system";
        if (solver == HGMM_SOLVER_TWIST_LSTSQ) {          // the 28 sums of the failed normal equations, for the bug report
            double hv[28];
            if (cudaMemcpy(hv, ctx->hist.as<double>() + 8, sizeof hv, cudaMemcpyDeviceToHost) == cudaSuccess) {
                msg += " v =";
                char b[40];
                for (double x : hv) { snprintf(b, sizeof b, " %.17g", x); msg += b; }
            }
        }
        FAIL(HGMM_ERR_NUMERIC, msg);
    }
    return HGMM_OK;
}

int hgmm_register_tree(hgmm_ctx* ctx, const hgmm_reg_config* cfg, double* rot, double* t, double* out_q, int32_t* out_iters,
                       double* out_q_hist) {
    if (!ctx) return HGMM_ERR_INVALID;
    if (!cfg || !rot || !t) FAIL(HGMM_ERR_INVALID, "null config / transform");
    if (!ctx->have_tree) FAIL(HGMM_ERR_STATE, "no tree: call hgmm_fit_tree / hgmm_tree_set_model first");
    if (!ctx->have_target) FAIL(HGMM_ERR_STATE, "no target: call hgmm_reg_set_target first");
    if (cfg->solver != HGMM_SOLVER_TWIST_LSTSQ && cfg->solver != HGMM_SOLVER_PROCRUSTES) FAIL(HGMM_ERR_INVALID, "unknown solver");
    if (cfg->maxiter < 1 || cfg->maxiter > 100000) FAIL(HGMM_ERR_INVALID, "bad maxiter");
    CK(cudaSetDevice(ctx->device));
    const TreeModel& tm = ctx->tm;
    const size_t rn = (size_t)tm.nt * 10;
    CK(ctx->racc.ensure(rn * sizeof(double)));
    CK(ctx->hist.ensure((size_t)(cfg->maxiter + 1) * sizeof(double)));
    int r = upload_Rt(ctx, rot, t);
    if (r != HGMM_OK) return r;
    cudaStream_t s = ctx->stream;
    int* ctrl = ctx->ctrl.as<int>();
    CK(cudaMemsetAsync(ctrl, 0, 8 * sizeof(int), s));
    CK(cudaMemsetAsync(ctx->qstate.p, 0, 4 * sizeof(double), s));
    CK(cudaMemsetAsync(ctx->racc.p, 0, rn * sizeof(double), s));
    CK(cudaEventRecord(ctx->ev0, s));
    // several ranks: the ranks' moments are all-reduced between the E-step and the solve (ncclAllReduce, 61.9 us/iteration at
    // N = 8 on the L3 bunny tree).  HGMM_REG_P2P=1 (read per call) exchanges them INSIDE the solve kernel over the peer-memory
    // windows instead (registration.cu: reg_gather_node; two PDL-chained launches per iteration, no NCCL call): bit-identical
    // results, but the all-to-all is pushed by the solve kernel's single CTA and measures 52.3 vs 49.5 us/iteration at N = 2
    // and 97.3 vs 61.9 at N = 8 (profiles/r02_multigpu_n8_test.txt), so it is not the default.
    const char* reg_p2p_env = getenv("HGMM_REG_P2P");
    const bool reg_no_p2p = !(reg_p2p_env && reg_p2p_env[0] == '1');
    RegXchgView xv = {};
    xv.nranks = 1;
    const bool fused = ctx->nranks > 1 && ctx->p2p_ready && ctx->xwin_reg_off != 0 && tm.nt <= kRegXchgNodes && !reg_no_p2p;
    if (fused) {
        for (int q = 0; q < ctx->nranks; ++q) xv.data[q] = reinterpret_cast<uint4*>(static_cast<char*>(ctx->xpeer[q]) + ctx->xwin_reg_off);
        xv.rank = ctx->rank; xv.nranks = ctx->nranks;
        xv.base = ctx->tepoch;
        ctx->tepoch += (uint32_t)cfg->maxiter + 2u;
        xv.timeout_ns = (unsigned long long)(ctx->xchg_timeout_s * 1e9);
    }
    const int batch = 4;
    int issued = 0;
    bool done = false;
    while (!done && issued < cfg->maxiter) {
        for (int b = 0; b < batch && issued < cfg->maxiter; ++b, ++issued) {
            // the solve kernel zeroes the moments it consumed, so no separate clearing launch is needed per iteration
            CK(launch_reg_estep(ctx->tx.as<float>(), ctx->ty.as<float>(), ctx->tz.as<float>(), ctx->nt_pts, ctx->Rt.as<double>(), tm,
                                cfg->lambda_c, ctx->racc.as<double>(), 0, ctrl, s));
            if (!fused) {
                r = allreduce(ctx, ctx->racc.as<double>(), rn);
                if (r != HGMM_OK) return r;
            }
            CK(launch_reg_solve(tm, ctx->racc.as<double>(), 1, cfg->solver, ctx->Rt.as<double>(), ctx->hist.as<double>(),
                                ctx->qstate.as<double>(), ctrl, cfg->tol, fused ? &xv : nullptr, s));
            ctx->launches += 2;
        }
        CK(cudaMemcpyAsync(ctx->h_ctrl, ctrl, 8 * sizeof(int), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        done = ctx->h_ctrl[0] != 0;
        if (ctx->h_ctrl[7]) FAIL(HGMM_ERR_NCCL, "registration: a wait on a peer rank timed out (rank missing, or the ranks' call sequences differ)");
    }
    CK(cudaEventRecord(ctx->ev1, s));
    const int iters = ctx->h_ctrl[1];
    CK(cudaMemcpyAsync(ctx->h_dbl, ctx->Rt.p, 12 * sizeof(double), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(ctx->h_dbl + 16, ctx->qstate.p, 2 * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (out_q_hist && iters > 0) CK(cudaMemcpyAsync(out_q_hist, ctx->hist.p, (size_t)iters * sizeof(double), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    for (int k = 0; k < 9; ++k) rot[k] = ctx->h_dbl[k];
    for (int k = 0; k < 3; ++k) t[k] = ctx->h_dbl[9 + k];
    if (out_q) *out_q = ctx->h_dbl[17];
    if (out_iters) *out_iters = iters;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    ctx->last_ms[0] = ms; ctx->last_ms[1] = ms; ctx->last_ms[2] = 0;
    ctx->have_racc = false;       // the loop's solve kernel consumed (zeroed) the moments
    if (ctx->h_ctrl[2]) FAIL(HGMM_ERR_NUMERIC, "registration solve: singular / non-positive-definite system");
    return HGMM_OK;
}

// ------------------------------------------------------------------------------------------
// flat-mixture registration: model = the context's flat mixture (hgmm_fit_flat), target = hgmm_reg_set_target
// (GMMRegistration::pointCloudRegisterGPU, src/c++/gmm_registration/gmm_reg.cu:54-56 -- an empty stub in the reference)
// ------------------------------------------------------------------------------------------
int hgmm_register_flat(hgmm_ctx* ctx, const hgmm_reg_config* cfg, double* rot, double* t, double* out_q, int32_t* out_iters,
                       double* out_q_hist) {
    if (!ctx) return HGMM_ERR_INVALID;
    if (!cfg || !rot || !t) FAIL(HGMM_ERR_INVALID, "null config / transform");
    if (!ctx->have_flat) FAIL(HGMM_ERR_STATE, "no flat model: call hgmm_fit_flat first");
    if (!ctx->have_target) FAIL(HGMM_ERR_STATE, "no target: call hgmm_reg_set_target first");
    if (cfg->solver != HGMM_SOLVER_TWIST_LSTSQ && cfg->solver != HGMM_SOLVER_PROCRUSTES) FAIL(HGMM_ERR_INVALID, "unknown solver");
    if (cfg->maxiter < 1 || cfg->maxiter > 100000) FAIL(HGMM_ERR_INVALID, "bad maxiter");
    const FlatModel& m = ctx->fm;
    if (cfg->solver == HGMM_SOLVER_TWIST_LSTSQ && m.cov_type != HGMM_COV_FULL)
        FAIL(HGMM_ERR_INVALID, "the twist solver needs full covariances; use HGMM_SOLVER_PROCRUSTES with diag / spherical mixtures");
    CK(cudaSetDevice(ctx->device));
    const int n = ctx->nt_pts;
    if (n <= 0) FAIL(HGMM_ERR_STATE, "empty target");
    const size_t fb = (size_t)n * sizeof(float);
    CK(ctx->rx.ensure(fb)); CK(ctx->ry.ensure(fb)); CK(ctx->rz.ensure(fb));
    const size_t acc_n = kAccHdr + (size_t)m.Jp * kMom;
    CK(ctx->acc.ensure(acc_n * sizeof(double)));
    CK(ctx->hist.ensure((size_t)(cfg->maxiter + 1) * sizeof(double)));
    // sweep plan for the target's size (the kernels of the fit; the PDL-chained flat_em7 stages its points before the grid
    // dependency resolves, which is only valid for a cloud that does not change between launches -- here it does)
    const bool v3 = m.Jp >= 64;
    int JT = 1, W = 8, Sdiv = 1, G = 1, grid = 1, big = 0;
    if (v3) flat3_plan(n, m.Jp, ctx->num_sms, ((m.Jp / 32 + 1) / 2 >= 9) ? 6 : 0, &W, &Sdiv, &G, &grid, &big);
    else flat2_plan(n, m.Jp, ctx->num_sms, 0, &JT, &W, &Sdiv, &G, &grid, &big);
    CK(ctx->partial.ensure((size_t)grid * G * kMom * m.Jp * sizeof(float)));
    CK(ctx->rowaux.ensure((size_t)grid * G * 2 * sizeof(double)));
    int r = upload_Rt(ctx, rot, t);
    if (r != HGMM_OK) return r;
    cudaStream_t s = ctx->stream;
    int* ctrl = ctx->ctrl.as<int>();
    CK(cudaMemsetAsync(ctrl, 0, 8 * sizeof(int), s));
    CK(cudaMemsetAsync(ctx->qstate.p, 0, 4 * sizeof(double), s));
    CK(cudaMemsetAsync(ctx->acc.p, 0, acc_n * sizeof(double), s));
    CK(cudaEventRecord(ctx->ev0, s));
    const int batch = 4;
    int issued = 0;
    bool done = false;
    while (!done && issued < cfg->maxiter) {
        for (int b = 0; b < batch && issued < cfg->maxiter; ++b, ++issued) {
            launch_transform_soa(ctx->tx.as<float>(), ctx->ty.as<float>(), ctx->tz.as<float>(), n, ctx->Rt.as<double>(),
                                 ctx->rx.as<float>(), ctx->ry.as<float>(), ctx->rz.as<float>(), ctrl, s);
            if (v3)
                CK(launch_em_flat3(ctx->rx.as<float>(), ctx->ry.as<float>(), ctx->rz.as<float>(), n, m, m.cref_blocks, W, Sdiv, G, grid, big,
                                   ctx->partial.as<float>(), ctx->rowaux.as<double>(), ctrl, s));
            else
                CK(launch_em_flat2(ctx->rx.as<float>(), ctx->ry.as<float>(), ctx->rz.as<float>(), n, m, m.cref_blocks, JT, W, Sdiv, G, grid,
                                   big, ctx->partial.as<float>(), ctx->rowaux.as<double>(), ctrl, s));
            CK(launch_flat_reduce(ctx->partial.as<float>(), ctx->rowaux.as<double>(), grid * G, m, ctx->acc.as<double>(), ctrl, s));
            r = allreduce(ctx, ctx->acc.as<double>(), kAccHdr + (size_t)m.J * kMom);
            if (r != HGMM_OK) return r;
            CK(launch_reg_flat_solve(m, ctx->acc.as<double>(), cfg->solver, ctx->Rt.as<double>(), ctx->hist.as<double>(),
                                     ctx->qstate.as<double>(), ctrl, cfg->tol, s));
            ctx->launches += 4;
        }
        CK(cudaMemcpyAsync(ctx->h_ctrl, ctrl, 4 * sizeof(int), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        done = ctx->h_ctrl[0] != 0;
    }
    CK(cudaEventRecord(ctx->ev1, s));
    const int iters = ctx->h_ctrl[1];
    CK(cudaMemcpyAsync(ctx->h_dbl, ctx->Rt.p, 12 * sizeof(double), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(ctx->h_dbl + 16, ctx->qstate.p, 2 * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (out_q_hist && iters > 0) CK(cudaMemcpyAsync(out_q_hist, ctx->hist.p, (size_t)iters * sizeof(double), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    for (int k = 0; k < 9; ++k) rot[k] = ctx->h_dbl[k];
    for (int k = 0; k < 3; ++k) t[k] = ctx->h_dbl[9 + k];
    if (out_q) *out_q = ctx->h_dbl[17];
    if (out_iters) *out_iters = iters;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    ctx->last_ms[0] = ms; ctx->last_ms[1] = ms; ctx->last_ms[2] = 0;
    if (ctx->h_ctrl[2]) FAIL(HGMM_ERR_NUMERIC, "registration solve: singular / non-positive-definite system");
    return HGMM_OK;
}

// ------------------------------------------------------------------------------------------
// L2-distance registration of two flat mixtures (gmmreg_gpu/cost_functions.py, gmmreg.py:101-107)
// device layout of l2buf (doubles): mu_s[3Js] | phi_s[Js] | mu_t[3Jt] | phi_t[Jt] | theta[8] | out[16]
// ------------------------------------------------------------------------------------------
int hgmm_l2_set_mixtures(hgmm_ctx* ctx, const double* mu_s, const double* phi_s, int32_t Js, const double* mu_t,
                         const double* phi_t, int32_t Jt) {
    if (!ctx) return HGMM_ERR_INVALID;
    if (!mu_s || !phi_s || !mu_t || !phi_t) FAIL(HGMM_ERR_INVALID, "null mixture pointer");
    // the cost / BFGS kernels are ONE CTA by design (the whole optimisation in one launch): J x J pairs per evaluation on one SM.
    // 4096 x 4096 is ~0.1 s per evaluation; the reference's use is J = 100 (gmmreg.py)
    if (Js < 1 || Jt < 1 || Js > 4096 || Jt > 4096) FAIL(HGMM_ERR_INVALID, "mixture sizes must be in [1, 4096]");
    CK(cudaSetDevice(ctx->device));
    const size_t nd = 4 * (size_t)Js + 4 * (size_t)Jt + 24;
    CK(ctx->l2buf.ensure(nd * sizeof(double)));
    double* d = ctx->l2buf.as<double>();
    cudaStream_t s = ctx->stream;
    CK(cudaMemcpyAsync(d, mu_s, 3 * (size_t)Js * sizeof(double), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d + 3 * (size_t)Js, phi_s, (size_t)Js * sizeof(double), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d + 4 * (size_t)Js, mu_t, 3 * (size_t)Jt * sizeof(double), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d + 4 * (size_t)Js + 3 * (size_t)Jt, phi_t, (size_t)Jt * sizeof(double), cudaMemcpyHostToDevice, s));
    CK(cudaStreamSynchronize(s));
    ctx->l2_js = Js;
    ctx->l2_jt = Jt;
    return HGMM_OK;
}

int hgmm_l2_cost_grad(hgmm_ctx* ctx, const double* theta, double sigma, double* out_f, double* out_grad) {
    if (!ctx) return HGMM_ERR_INVALID;
    if (ctx->l2_js <= 0) FAIL(HGMM_ERR_STATE, "no mixtures: call hgmm_l2_set_mixtures first");
    if (!theta || !(sigma > 0.0)) FAIL(HGMM_ERR_INVALID, "null theta / non-positive sigma");
    CK(cudaSetDevice(ctx->device));
    const size_t Js = ctx->l2_js, Jt = ctx->l2_jt;
    double* d = ctx->l2buf.as<double>();
    double* th = d + 4 * Js + 4 * Jt;
    double* out = th + 8;
    cudaStream_t s = ctx->stream;
    for (int k = 0; k < 7; ++k) ctx->h_dbl[k] = theta[k];
    CK(cudaMemcpyAsync(th, ctx->h_dbl, 7 * sizeof(double), cudaMemcpyHostToDevice, s));
    CK(launch_l2_cost_grad(d, d + 3 * Js, (int)Js, d + 4 * Js, d + 4 * Js + 3 * Jt, (int)Jt, th, sigma, out, s));
    ctx->launches += 1;
    CK(cudaMemcpyAsync(ctx->h_dbl + 16, out, 8 * sizeof(double), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (out_f) *out_f = ctx->h_dbl[16];
    if (out_grad) for (int k = 0; k < 7; ++k) out_grad[k] = ctx->h_dbl[17 + k];
    return HGMM_OK;
}

int hgmm_l2_optimize(hgmm_ctx* ctx, double* theta, double sigma, int32_t max_iter, double gtol, double* out_f,
                     int32_t* out_iters, int32_t* out_nfev, int32_t* out_status) {
    if (!ctx) return HGMM_ERR_INVALID;
    if (ctx->l2_js <= 0) FAIL(HGMM_ERR_STATE, "no mixtures: call hgmm_l2_set_mixtures first");
    if (!theta || !(sigma > 0.0) || max_iter < 0 || max_iter > 100000) FAIL(HGMM_ERR_INVALID, "bad theta / sigma / max_iter");
    CK(cudaSetDevice(ctx->device));
    const size_t Js = ctx->l2_js, Jt = ctx->l2_jt;
    double* d = ctx->l2buf.as<double>();
    double* th = d + 4 * Js + 4 * Jt;
    double* out = th + 8;
    cudaStream_t s = ctx->stream;
    for (int k = 0; k < 7; ++k) ctx->h_dbl[k] = theta[k];
    CK(cudaMemcpyAsync(th, ctx->h_dbl, 7 * sizeof(double), cudaMemcpyHostToDevice, s));
    CK(cudaEventRecord(ctx->ev0, s));
    CK(launch_l2_bfgs(d, d + 3 * Js, (int)Js, d + 4 * Js, d + 4 * Js + 3 * Jt, (int)Jt, th, sigma, max_iter, gtol, out, s));
    CK(cudaEventRecord(ctx->ev1, s));
    ctx->launches += 1;
    CK(cudaMemcpyAsync(ctx->h_dbl + 8, th, 7 * sizeof(double), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(ctx->h_dbl + 16, out, 11 * sizeof(double), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    for (int k = 0; k < 7; ++k) theta[k] = ctx->h_dbl[8 + k];
    if (out_f) *out_f = ctx->h_dbl[16];
    if (out_iters) *out_iters = (int32_t)ctx->h_dbl[17];
    if (out_nfev) *out_nfev = (int32_t)ctx->h_dbl[18];
    if (out_status) *out_status = (int32_t)ctx->h_dbl[19];
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    ctx->last_ms[0] = ms; ctx->last_ms[1] = ms; ctx->last_ms[2] = 0;
    return HGMM_OK;
}

// ------------------------------------------------------------------------------------------
int hgmm_fill_vbo(hgmm_ctx* ctx, float* vbo_positions, float* vbo_colors, float scene_scale, const float* rgb_points,
                  const float* rgb_target) {
    if (!ctx) return HGMM_ERR_INVALID;
    if (!(scene_scale > 0.f)) FAIL(HGMM_ERR_INVALID, "scene_scale must be positive");
    CK(cudaSetDevice(ctx->device));
    const float d1[3] = {1.f, 1.f, 1.f}, d2[3] = {1.f, 1.f, 0.f};     // gmm_kernels.cu:573-574
    const float* c1 = rgb_points ? rgb_points : d1;
    const float* c2 = rgb_target ? rgb_target : d2;
    launch_fill_vbo(ctx->bx.as<float>(), ctx->by.as<float>(), ctx->bz.as<float>(), ctx->n, 0, vbo_positions, vbo_colors, scene_scale,
                    c1[0], c1[1], c1[2], ctx->stream);
    ctx->launches += ctx->n > 0 ? 1 : 0;
    if (ctx->have_target) {
        launch_fill_vbo(ctx->tx.as<float>(), ctx->ty.as<float>(), ctx->tz.as<float>(), ctx->nt_pts, ctx->n, vbo_positions, vbo_colors,
                        scene_scale, c2[0], c2[1], c2[2], ctx->stream);
        ctx->launches += ctx->nt_pts > 0 ? 1 : 0;
    }
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));     // the reference syncs here too (gmm_kernels.cu:541)
    return HGMM_OK;
}

// ------------------------------------------------------------------------------------------
int hgmm_comm_unique_id(void* out_id128) {
    if (!out_id128) return HGMM_ERR_INVALID;
    if (!nccl_load()) return HGMM_ERR_NCCL;
    ncclUniqueId128 id;
    if (g_nccl.GetUniqueId(&id) != 0) return HGMM_ERR_NCCL;
    memcpy(out_id128, &id, 128);
    return HGMM_OK;
}

int hgmm_comm_init(hgmm_ctx* ctx, int rank, int nranks, const void* id128) {
    if (!ctx) return HGMM_ERR_INVALID;
    if (nranks < 1 || rank < 0 || rank >= nranks || !id128) FAIL(HGMM_ERR_INVALID, "bad rank / nranks / id");
    if (!nccl_load()) FAIL(HGMM_ERR_NCCL, g_nccl.err);
    CK(cudaSetDevice(ctx->device));
    ncclUniqueId128 id;
    memcpy(&id, id128, 128);
    int r = g_nccl.CommInitRank(&ctx->comm, nranks, id, rank);
    if (r != 0) {
        ctx->comm = nullptr;
        FAIL(HGMM_ERR_NCCL, std::string("ncclCommInitRank: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error"));
    }
    ctx->rank = rank;
    ctx->nranks = nranks;
    return HGMM_OK;
}

// ---- peer-memory exchange window (one process per GPU on one box; handles travel over the caller's bootstrap channel)
int hgmm_p2p_export(hgmm_ctx* ctx, void* out_handle64) {
    if (!ctx || !out_handle64) return HGMM_ERR_INVALID;
    if (ctx->nranks < 2) FAIL(HGMM_ERR_STATE, "hgmm_comm_init first");
    if (ctx->nranks > kXchgMaxRanks) FAIL(HGMM_ERR_INVALID, "peer-memory exchange supports at most 8 ranks");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    CK(cudaSetDevice(ctx->device));
    if (!ctx->xwin) {
        // flat region (xchg.cuh) followed by the tree region (tree_level.cuh), sized for trees up to HGMM_P2P_TREE_LEVELS
        // (default 5: 16 MB; 6 needs 131 MB).  Always a freshly zeroed allocation: tags / epochs only grow over the context's
        // life (never reset on re-attach), so no stale cell can ever validate.
        int lv = 5;
        if (const char* e = getenv("HGMM_P2P_TREE_LEVELS")) lv = atoi(e);
        if (lv < 0) lv = 0;
        if (lv > 6) lv = 6;
        const size_t tree_bytes = (lv > 0 ? tree_win_layout(lv).bytes : 0);
        const size_t reg_off = (kXchgBytes + tree_bytes + 255) / 256 * 256;
        const size_t bytes = reg_off + kRegXchgBytes;                   // + the tree-registration region (12 MB)
        CK(cudaMalloc(&ctx->xwin, bytes));
        CK(cudaMemset(ctx->xwin, 0, bytes));
        CK(cudaDeviceSynchronize());
        ctx->xwin_tree_levels = lv;
        ctx->xwin_reg_off = reg_off;
    }
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, ctx->xwin));
    memcpy(out_handle64, &h, 64);
    return HGMM_OK;
}

int hgmm_p2p_attach(hgmm_ctx* ctx, const void* handles, int32_t n_handles) {
    if (!ctx || !handles) return HGMM_ERR_INVALID;
    if (!ctx->xwin) FAIL(HGMM_ERR_STATE, "hgmm_p2p_export first");
    if (n_handles != ctx->nranks) FAIL(HGMM_ERR_INVALID, "one handle per rank, in rank order");
    CK(cudaSetDevice(ctx->device));
    for (int r = 0; r < ctx->nranks; ++r) {
        if (r == ctx->rank) {
            ctx->xpeer[r] = ctx->xwin;
            continue;
        }
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const char*>(handles) + 64 * (size_t)r, 64);
        void* p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            for (int q = 0; q < r; ++q)
                if (ctx->xpeer[q] && ctx->xpeer[q] != ctx->xwin) { cudaIpcCloseMemHandle(ctx->xpeer[q]); ctx->xpeer[q] = nullptr; }
            ctx->err = std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e) + " (no peer access between these devices?)";
            return HGMM_ERR_CUDA;
        }
        ctx->xpeer[r] = p;
    }
    ctx->p2p_ready = true;               // xepoch / tepoch keep counting: see hgmm_p2p_export
    return HGMM_OK;
}

int hgmm_p2p_detach(hgmm_ctx* ctx) {
    if (!ctx) return HGMM_ERR_INVALID;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    p2p_release(ctx);
    return HGMM_OK;
}

int hgmm_p2p_enabled(const hgmm_ctx* ctx) { return ctx && ctx->p2p_ready ? 1 : 0; }

int hgmm_comm_destroy(hgmm_ctx* ctx) {
    if (!ctx) return HGMM_ERR_INVALID;
    if (ctx->comm && g_nccl.CommDestroy) {
        cudaStreamSynchronize(ctx->stream);
        g_nccl.CommDestroy(ctx->comm);
    }
    cudaStreamSynchronize(ctx->stream);
    p2p_release(ctx);
    ctx->comm = nullptr;
    ctx->nranks = 1;
    ctx->rank = 0;
    return HGMM_OK;
}

// ------------------------------------------------------------------------------------------
int hgmm_measure_fp32_peak(hgmm_ctx* ctx, double* out_tflops) {
    if (!ctx || !out_tflops) return HGMM_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    CK(ctx->hist.ensure(64));
    const int blocks = ctx->num_sms * 8, iters = 4096;
    for (int mode = 0; mode < 3; ++mode) {
        double best = 0.0;
        for (int rep = 0; rep < 4; ++rep) {
            CK(cudaEventRecord(ctx->ev0, ctx->stream));
            if (mode < 2) CK(launch_ffma_peak(ctx->hist.as<float>(), blocks, iters, mode, ctx->stream));
            else CK(launch_ffma2_peak(ctx->hist.as<float>(), blocks, iters, ctx->stream));
            CK(cudaEventRecord(ctx->ev1, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            ctx->launches += 1;
            float ms = 0.f;
            cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
            const double flops = (double)blocks * 256.0 * iters * 16.0 * 8.0 * 2.0 * (mode == 2 ? 2.0 : 1.0);
            const double tf = flops / (ms * 1e-3) / 1e12;
            if (rep > 0 && tf > best) best = tf;
        }
        out_tflops[mode] = best;
    }
    return HGMM_OK;
}

int hgmm_set_profiling(hgmm_ctx* ctx, int on) {
    if (!ctx) return HGMM_ERR_INVALID;
    ctx->profiling = on != 0;
    return HGMM_OK;
}

int hgmm_last_timing(const hgmm_ctx* ctx, double* out_ms3) {
    if (!ctx || !out_ms3) return HGMM_ERR_INVALID;
    out_ms3[0] = ctx->last_ms[0];
    out_ms3[1] = ctx->last_ms[1];
    out_ms3[2] = ctx->last_ms[2];
    return HGMM_OK;
}

}  // extern "C"
