/*
 * hgmm.h -- C ABI of libhgmm, the B200 (sm_100a) hierarchical-GMM fit / register engine.
 *
 * This is the drop-in boundary for the reference's fit/register hot path.  Each entry point
 * names the reference interface it replaces (paths relative to the reference checkout):
 *
 *   hgmm_set_points        scanRegistration::initSimulation  src/c++/gmm_fit/gmm_kernels.cu:544-578
 *                          GMMRegistration::initSimulation   src/c++/gmm_registration/gmm_reg.cu:19-43
 *   hgmm_fit_flat          GMM::solve                        src/c++/gmm_fit/gmm_kernels.cu:371-504
 *                          train_gmm                         src/python/gmm_waymo/src/gmm_impl.py:118-145
 *   hgmm_predict_flat      predict                           src/python/gmm_waymo/src/gmm_impl.py:147-155
 *   hgmm_fit_tree          buildGMMTree                      src/python/hgmm/hgmm_gpu.py:466-548
 *   hgmm_tree_set_model    GMMTree._mixingCoeff/_mean/_covar src/python/hgmm/hgmm_gpu.py:685-706
 *   hgmm_reg_estep         gmmTreeRegESTep                   src/python/hgmm/hgmm_gpu.py:550-577
 *   hgmm_reg_mstep         GMMTree.maximization_step         src/python/hgmm/hgmm_gpu.py:729-752
 *   hgmm_register_tree     GMMTree.registration              src/python/hgmm/hgmm_gpu.py:754-768
 *   hgmm_register_flat     GMMRegistration::pointCloudRegisterGPU (empty stub) src/c++/gmm_registration/gmm_reg.cu:54-56
 *   hgmm_io_read_ply/pcd   readData  src/c++/main.cpp:45-79, readPointCloud  src/c++/main_reg.cpp:106-161
 *   hgmm_l2_*              RigidCostFunction / L2DistRegistration  src/python/gmmreg_gpu/cost_functions.py:29-69, gmmreg.py:62-121
 *   hgmm_fill_vbo          scanRegistration::copyBoidsToVBO  src/c++/gmm_fit/gmm_kernels.cu:532-542
 *                          GMMRegistration::copyBoidsToVBO   src/c++/gmm_registration/gmm_reg.cu:45-52
 *   hgmm_comm_*            (no reference counterpart: points sharded over ranks, one all-reduce of
 *                           the O(J) sufficient statistics per EM iteration)
 *
 * Conventions: plain pointers and sizes only; every call returns HGMM_OK (0) or a negative
 * status and never exits the process (the C++ shim in cpp/ converts status -> exit(EXIT_FAILURE)
 * to match common/utilities.cpp:16-25).  Host arrays are caller-owned.  vec3 arrays are packed
 * 3 x fp32 (12 B, the layout of glm::vec3); 3x3 matrices are 9 x fp32 row-major (symmetric
 * covariances, so glm::mat3's column-major layout is byte-identical).  A context is bound to one
 * device and one stream; calls on one context must not overlap (the reference is single-threaded).
 */
#ifndef HGMM_H_
#define HGMM_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HGMM_OK 0
#define HGMM_ERR_INVALID (-1)      /* bad argument */
#define HGMM_ERR_CUDA (-2)         /* CUDA runtime error, see hgmm_last_error */
#define HGMM_ERR_STATE (-3)        /* call order (no points / no tree yet) */
#define HGMM_ERR_NCCL (-4)         /* NCCL unavailable or failed */
#define HGMM_ERR_NUMERIC (-5)      /* singular system in the registration solve */
#define HGMM_ERR_IO (-6)           /* file cannot be opened / is not in a supported format */

#define HGMM_MEM_HOST 0
#define HGMM_MEM_DEVICE 1

/* covariance structure of the flat mixture */
#define HGMM_COV_FULL 0
#define HGMM_COV_DIAG 1
#define HGMM_COV_SPHERICAL 2

/* which reference variant's EM semantics the flat fit follows (SURVEY.md section 7 compat table) */
#define HGMM_FLAVOR_CPP 0          /* gmm_kernels.cu: full cov, pi = N_j/N, no regularisation, fixed iterations */
#define HGMM_FLAVOR_PY 1           /* gmm_waymo/src/gmm_impl.py: diag/spherical, +1e-8 / +1e-6 terms, tol on mean log-lik */
#define HGMM_FLAVOR_PY_OLD 2       /* gmmreg_gpu/gmm_impl.py (the L2 registration's fitter): diag, clip(cov, 0), pi = nk/N */

/* level log-likelihood used for the tree's convergence test */
#define HGMM_LL_LEVEL 0            /* reference: scan of all 8^(l+1) nodes of the level with the new parameters */
#define HGMM_LL_ESTEP 1            /* fast: the E-step's own 8-sibling normaliser (one iteration lag) */

#define HGMM_SOLVER_TWIST_LSTSQ 0  /* GMMTree.maximization_step (linearised twist least squares) */
#define HGMM_SOLVER_PROCRUSTES 1   /* weighted Procrustes, 3x3 Jacobi SVD (svd3.h algorithm) */

typedef struct hgmm_ctx hgmm_ctx;

typedef struct hgmm_flat_config {
    int32_t n_components;          /* J */
    int32_t cov_type;              /* HGMM_COV_* */
    int32_t flavor;                /* HGMM_FLAVOR_* */
    int32_t max_iter;
    float tol;                     /* PY flavour: stop when |d mean-log-lik| < tol; CPP flavour ignores it */
    int32_t sigma_bug;             /* CPP flavour only: reproduce gmm_kernels.cu:97-103 (d^T Sigma d) */
    int32_t tile_points;           /* 0 = auto.  64/128/256/512: points per CTA tile of the first-generation kernel (reserved = 1).
                                    * 1..12: A/B switch between builds of the packed sweep (profiles/variants_probe.py, tests):
                                    * 1 one register-rich CTA per SM, 2 two CTAs per SM, 3 two teams of 4 components per lane,
                                    * 4 per-batch mbarrier pipeline, 5 152-register build, 6 staged chunks with CTA barriers
                                    * (flat_em5.cu), 7 one component per thread (flat_em6.cu), 8 staged chunks with the
                                    * barrier-free pipeline (flat_em7.cu; what 0 selects for J > 512), 9 that pipeline with the
                                    * moment pass about one origin per chunk of the cell-sorted cloud (flat_em8.cu), 10 its
                                    * Cholesky-form density pass, 11 / 12 = 9 / 10 with a staggered pass order */
    int32_t reserved;              /* kernel variant: 0 = default (packed-FP32 sweep), 1 = first-generation two-phase kernel, 2 = scalar single-evaluation kernel, 3 = packed sweep with separate reduce / finalize launches */
} hgmm_flat_config;

typedef struct hgmm_tree_config {
    int32_t max_level;             /* L: levels of 8,64,...,8^L nodes; n_total = 8(8^L-1)/7 */
    int32_t ll_mode;               /* HGMM_LL_* */
    float ls;                      /* convergence threshold on |q - prevQ| (reference: 20 GPU / 80 CPU) */
    float ld;                      /* blank a node whose zeroth moment is < ld (reference 1e-4) */
    float sig2;                    /* initial isotropic variance (reference 0.004 GPU / 0.00034 CPU) */
    int32_t max_iters_per_level;   /* safety cap; the reference has none */
    int32_t chunk_points;          /* 0 = auto; points per warp work item (multiple of 32) */
    int32_t reserved;              /* E-step kernel: 0 = packed-FP32 (default), 1 = scalar first-generation kernel */
    /* adaptive (pruned, ragged) build -- the "adaptive scaling" of README.md:74, which the reference only applies at
     * registration time (lambda_c stop, hgmm_gpu.py:572) and through the M0 < ld blanking (hgmm_cupy_cpu_working.py:111-112).
     * After a level has converged a node becomes TERMINAL -- its 8 children stay blank and its points take no part in deeper
     * levels -- when it is blank, when complexity(Sigma) = lambda_min / trace <= prune_lambda_c (already a plane at this scale),
     * or when it holds fewer than prune_min_points points (N * pi).  0 / 0 = off (the reference's full tree).  Needs
     * ll_mode = HGMM_LL_ESTEP. */
    float prune_lambda_c;
    float prune_min_points;
} hgmm_tree_config;

typedef struct hgmm_reg_config {
    int32_t solver;                /* HGMM_SOLVER_* */
    int32_t maxiter;               /* reference default 20 */
    float tol;                     /* reference default 1e-4 on |q - q_prev| */
    float lambda_c;                /* complexity pruning threshold, reference default 0.01 */
} hgmm_reg_config;

/* ---- lifecycle ---- */
/* device: CUDA ordinal; stream: a cudaStream_t to launch on, or NULL for a context-owned stream. */
int hgmm_create(hgmm_ctx** out, int device, void* stream);
int hgmm_destroy(hgmm_ctx* ctx);
const char* hgmm_last_error(const hgmm_ctx* ctx);
const char* hgmm_version(void);
/* number of kernels this context has launched since creation (bench.py's gpu_launches evidence) */
int64_t hgmm_launch_count(const hgmm_ctx* ctx);

/* ---- data ---- */
/* The cloud the mixture is fitted to (this rank's shard when a communicator is attached).  DEVICE and pageable HOST buffers are
 * consumed before the call returns; a PINNED host buffer is read asynchronously and must stay untouched until the next call
 * that synchronises (any fit / predict / register).  HGMM_MEM_DEVICE input is read on the context's stream: the caller must have finished writing it (synchronise
 * the producing stream, or create the context on that stream) -- the Python wrapper does this for torch tensors. */
int hgmm_set_points(hgmm_ctx* ctx, const float* xyz, int64_t n, int mem_kind);
/* total number of points over all ranks (= n without a communicator) */
int64_t hgmm_total_points(const hgmm_ctx* ctx);
/* With a communicator hgmm_set_points all-reduces the shard sizes (a blocking collective).  A caller that sharded the cloud itself
 * knows the total: declaring it (> 0) makes every following hgmm_set_points local; 0 restores the all-reduce.  Every rank must
 * declare the same value. */
int hgmm_declare_total_points(hgmm_ctx* ctx, int64_t n_total);

/* ---- flat mixture ---- */
/* init_means [J,3]; init_covs: FULL [J,9] | DIAG [J,3] | SPHERICAL [J] (variances); init_weights [J].
 * Outputs (any may be NULL): out_means [J,3], out_covs (same shape rule), out_weights [J],
 * out_inv_cov (PY flavour: 1/std, [J,3] or [J]), out_ll [max_iter] (PY: mean log p per iteration;
 * CPP: sum_i log p(x_i) before each M step), out_iters (iterations actually run). All host memory. */
int hgmm_fit_flat(hgmm_ctx* ctx, const hgmm_flat_config* cfg,
                  const float* init_means, const float* init_covs, const float* init_weights,
                  float* out_means, float* out_covs, float* out_weights, float* out_inv_cov,
                  double* out_ll, int32_t* out_iters);
/* hard assignment argmax_j(log N_j(x) + log(pi_j + eps)) with the last fitted flat model;
 * xyz NULL = the context's own points; labels [n] int32 on the host. */
int hgmm_predict_flat(hgmm_ctx* ctx, const float* xyz, int64_t n, int mem_kind, int32_t* labels);

/* ---- hierarchical mixture ---- */
int64_t hgmm_tree_total_nodes(int32_t max_level);
/* init_means [n_total,3] (the reference draws points[randint]); outputs (any may be NULL):
 * out_pi [n_total], out_mu [n_total,3], out_cov [n_total,9], out_current [n] int64 node id of each
 * point at the leaf level in ORIGINAL point order, out_iters [max_level], out_q [max_level] last q. */
int hgmm_fit_tree(hgmm_ctx* ctx, const hgmm_tree_config* cfg, const float* init_means,
                  float* out_pi, float* out_mu, float* out_cov, int64_t* out_current,
                  int32_t* out_iters, double* out_q);
/* install a tree (e.g. one fitted elsewhere) as the registration model */
int hgmm_tree_set_model(hgmm_ctx* ctx, int32_t max_level, const float* pi, const float* mu, const float* cov);

/* ---- registration against the context's tree ---- */
/* target cloud (NOT transformed); kept on the device for the following calls */
int hgmm_reg_set_target(hgmm_ctx* ctx, const float* xyz, int64_t n, int mem_kind);
/* one E-step of the target transformed by (rot row-major [9], t [3]); outputs [n_total], [n_total,3],
 * [n_total,9] (out_m2 may be NULL to skip second moments) */
int hgmm_reg_estep(hgmm_ctx* ctx, const double* rot, const double* t, float lambda_c,
                   double* out_m0, double* out_m1, double* out_m2);
/* one M-step from the moments of the last hgmm_reg_estep: updates (rot, t) in place, returns q */
int hgmm_reg_mstep(hgmm_ctx* ctx, int32_t solver, double* rot, double* t, double* out_q);
/* full loop; rot/t are in/out (pass identity/zero to start); returns the FORWARD transform that
 * maps target onto the model (the reference returns its inverse, hgmm_gpu.py:768 -- the Python
 * wrapper inverts). out_q_hist [maxiter] may be NULL. */
int hgmm_register_tree(hgmm_ctx* ctx, const hgmm_reg_config* cfg, double* rot, double* t,
                       double* out_q, int32_t* out_iters, double* out_q_hist);

/* ---- registration against the context's FLAT mixture (BASELINE configs[3]: "GMM registration, J components, weighted-SVD solve") ----
 * replaces GMMRegistration::pointCloudRegisterGPU  src/c++/gmm_registration/gmm_reg.cu:54-56 (an empty stub in the reference;
 * class fields gmm_reg.h:8-20).  Model = the mixture of the last hgmm_fit_flat (fitted to the context's points), target =
 * hgmm_reg_set_target.  Each iteration: responsibilities of the transformed target over all J components (the fit's own fused
 * sweep), then HGMM_SOLVER_PROCRUSTES (weighted Procrustes between the mass centroids and the means, 3x3 Jacobi SVD; any
 * covariance type) or HGMM_SOLVER_TWIST_LSTSQ (the tree solver's linearised Mahalanobis step; full covariances).  rot/t, q,
 * iterations and the stopping rule are those of hgmm_register_tree; lambda_c is ignored. */
int hgmm_register_flat(hgmm_ctx* ctx, const hgmm_reg_config* cfg, double* rot, double* t,
                       double* out_q, int32_t* out_iters, double* out_q_hist);

/* ---- L2-distance registration of two flat mixtures (float64) ----
 * replaces RigidCostFunction.__call__ / compute_l2_dist  src/python/gmmreg_gpu/cost_functions.py:29-69,
 * GaussTransform.compute src/python/gmmreg_gpu/transforms.py:73-86, diff_rot_from_quaternion src/python/gmmreg_gpu/so.py:4-59
 * and the scipy BFGS call of L2DistRegistration.registration src/python/gmmreg_gpu/gmmreg.py:101-107.
 * mu_* [J,3], phi_* [J] (the reference passes the fitted weights x 1e3); theta = (qw,qx,qy,qz,tx,ty,tz).
 * 1 <= J <= 4096 per mixture: the evaluation is a single-CTA kernel (so that a whole BFGS run is one launch). */
int hgmm_l2_set_mixtures(hgmm_ctx* ctx, const double* mu_source, const double* phi_source, int32_t n_source,
                         const double* mu_target, const double* phi_target, int32_t n_target);
/* f(theta) and its gradient [7] exactly as the reference's cost function returns them */
int hgmm_l2_cost_grad(hgmm_ctx* ctx, const double* theta, double sigma, double* out_f, double* out_grad);
/* the whole BFGS minimisation in one kernel launch; theta is in/out; out_status: 0 = |grad|_inf <= gtol,
 * 1 = iteration limit, 2 = line search failed (the end point is still the best one found) */
int hgmm_l2_optimize(hgmm_ctx* ctx, double* theta, double sigma, int32_t max_iter, double gtol, double* out_f,
                     int32_t* out_iters, int32_t* out_nfev, int32_t* out_status);

/* ---- ingestion: the callers' step in front of the path (host only, no context, no CUDA call) ----
 * replaces readData  src/c++/main.cpp:45-79 (HGMM_PLY_VIEWER_FIT; its hard-coded similarity transforms are not applied),
 * readPointCloud  src/c++/main_reg.cpp:106-161 (HGMM_PLY_VIEWER_REG), and the Open3D read_point_cloud calls of the Python
 * drivers for `DATA ascii|binary` PCD files (src/python/hgmm/hgmm_gpu.py:813-822).  out_xyz: [capacity,3] packed float32 or
 * NULL with capacity 0; *out_n = points in the file (may exceed capacity: call again with a larger buffer). */
#define HGMM_PLY_HEADER 0          /* header-driven: `element vertex N`, x y z first, ASCII */
#define HGMM_PLY_VIEWER_FIT 1
#define HGMM_PLY_VIEWER_REG 2
int hgmm_io_read_ply(const char* path, int32_t mode, float* out_xyz, int64_t capacity, int64_t* out_n);
int hgmm_io_read_pcd(const char* path, float* out_xyz, int64_t capacity, int64_t* out_n);

/* ---- viewer glue ---- */
/* writes 4 floats per point: pos = (-x, -y, -z)/scene_scale, 1 ; col = rgb + 0.3, 1.
 * src = the context's points, then the registration target if set. vbo_* are DEVICE pointers
 * (CUDA-mapped GL buffers). */
int hgmm_fill_vbo(hgmm_ctx* ctx, float* vbo_positions, float* vbo_colors, float scene_scale,
                  const float* rgb_points, const float* rgb_target);

/* ---- multi-GPU ---- */
/* 128-byte NCCL unique id, created on rank 0 and distributed by the caller (torch.distributed) */
int hgmm_comm_unique_id(void* out_id128);
int hgmm_comm_init(hgmm_ctx* ctx, int rank, int nranks, const void* id128);
int hgmm_comm_destroy(hgmm_ctx* ctx);
/* Peer-memory exchange for the flat fit (ranks = GPUs of one box, one process each; after hgmm_comm_init).
 * hgmm_p2p_export allocates this rank's exchange window and returns its 64-byte cudaIpcMemHandle_t; the caller gathers
 * the handles of all ranks (any channel) and passes them, in rank order, to hgmm_p2p_attach.  From then on every EM
 * iteration of hgmm_fit_flat is two kernels: the sweep, and one that folds the rank's partial sums, stores them into
 * every peer's window over NVLink, adds the ranks' contributions in rank order and finalizes -- no NCCL call, no third
 * launch.  Without these calls (or if the devices cannot map each other) the NCCL all-reduce path is used. */
int hgmm_p2p_export(hgmm_ctx* ctx, void* out_handle64);
int hgmm_p2p_attach(hgmm_ctx* ctx, const void* handles, int32_t n_handles);
int hgmm_p2p_detach(hgmm_ctx* ctx);                 /* back to the NCCL path (collective decision of the caller) */
int hgmm_p2p_enabled(const hgmm_ctx* ctx);

/* ---- measurement helper ---- */
/* FP32 FMA throughput of this device in TFLOP/s, out_tflops[3]: [0] scalar FFMA with immediate
 * operands, [1] scalar FFMA with three register operands, [2] packed FFMA2 (fma.rn.f32x2) with
 * register operands -- the pipe the sweep kernels run on and the non-tensor roofline denominator
 * bench.py reports next to the HBM one */
int hgmm_measure_fp32_peak(hgmm_ctx* ctx, double* out_tflops);
/* the cloud of hgmm_set_points in the order the large-mixture sweep (J > 512) reads it: a stable
 * counting sort by a 16^3 Morton cell grid over the cloud's bounding box (csrc/cloud_sort.cu; the
 * reference's kernels, gmm_kernels.cu:278-350, take the cloud in file order -- EM sums do not depend
 * on the order beyond rounding).  out_xyz: host, n x 3 floats.  Diagnostic; builds the order if no
 * fit has needed it yet. */
int hgmm_sorted_points(hgmm_ctx* ctx, float* out_xyz);
/* device time of the last fit/registration call in ms (CUDA events on the context stream):
 * [0] whole enqueue-to-finish loop, [1] sum over the E/M sweep kernel launches alone (only when
 * profiling is on, else 0), [2] number of E/M sweep launches that were timed */
int hgmm_last_timing(const hgmm_ctx* ctx, double* out_ms3);
/* on != 0: bracket every E/M sweep launch of the following fits with its own event pair (adds a few
 * microseconds per launch; used by bench.py for the per-kernel roofline, never inside a timed step) */
int hgmm_set_profiling(hgmm_ctx* ctx, int on);

#ifdef __cplusplus
}
#endif
#endif /* HGMM_H_ */
