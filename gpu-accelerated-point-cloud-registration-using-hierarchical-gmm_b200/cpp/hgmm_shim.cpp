// hgmm_shim.cpp -- reference C++ API over the libhgmm C ABI (see hgmm_shim.h for the interfaces replaced).
// Error behaviour follows checkCUDAErrorWithLine (src/c++/common/utilities.cpp:16-25): message to stderr, exit(EXIT_FAILURE).
#include "hgmm_shim.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/hgmm.h"

static_assert(sizeof(glm::vec3) == 12, "glm::vec3 must be packed 3 x fp32");

namespace {
hgmm_ctx* g_ctx = nullptr;          // the reference keeps file-scope device state too (gmm_kernels.cu:24-30)
int g_source = 0, g_target = 0;

void check(hgmm_ctx* ctx, int rc, const char* what, int line) {
    if (rc == HGMM_OK) return;
    fprintf(stderr, "hgmm error at line %d: %s: %s\n", line, what, ctx ? hgmm_last_error(ctx) : "no context");
    exit(EXIT_FAILURE);
}
#define CHECK(ctx, call) check(ctx, (call), #call, __LINE__)

hgmm_ctx* make_ctx() {
    hgmm_ctx* c = nullptr;
    int rc = hgmm_create(&c, 0, nullptr);            // the viewer uses device 0 (main.cpp:205)
    if (rc != HGMM_OK) {
        fprintf(stderr, "hgmm_create failed (%d): no CUDA device\n", rc);
        exit(EXIT_FAILURE);
    }
    return c;
}
}  // namespace

// ------------------------------------------------------------------------------------------ GMM / scanRegistration
void scanRegistration::initSimulation(vector<glm::vec3>& source, vector<glm::vec3>& target, int /*components*/) {
    if (!g_ctx) g_ctx = make_ctx();
    g_source = (int)source.size();
    g_target = (int)target.size();
    CHECK(g_ctx, hgmm_set_points(g_ctx, reinterpret_cast<const float*>(source.data()), g_source, HGMM_MEM_HOST));
    CHECK(g_ctx, hgmm_reg_set_target(g_ctx, reinterpret_cast<const float*>(target.data()), g_target, HGMM_MEM_HOST));
}

void GMM::solveWithInit(const glm::vec3* init_mean, float sigma0_sq, glm::vec3* mean, float* weights, float* covariances,
                        int iterations, int /*N*/) {
    if (!g_ctx) {
        fprintf(stderr, "GMM::solve before scanRegistration::initSimulation\n");
        exit(EXIT_FAILURE);
    }
    vector<float> cov0((size_t)components * 9, 0.f), w0((size_t)components, 1.0f / components);
    for (int j = 0; j < components; ++j) cov0[9 * j] = cov0[9 * j + 4] = cov0[9 * j + 8] = sigma0_sq;
    hgmm_flat_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.n_components = components;
    cfg.cov_type = HGMM_COV_FULL;
    cfg.flavor = HGMM_FLAVOR_CPP;
    cfg.max_iter = iterations;
    int32_t iters = 0;
    CHECK(g_ctx, hgmm_fit_flat(g_ctx, &cfg, reinterpret_cast<const float*>(init_mean), cov0.data(), w0.data(),
                               reinterpret_cast<float*>(mean), covariances, weights, nullptr, nullptr, &iters));
}

void GMM::solve(vector<glm::vec3> points, glm::vec3* mean, float* weights, int iterations, int N) {
    // gmm_kernels.cu:374-379: mu[i] = points[rand() % sourcePoints], weights 1/J; :397-402: Sigma = I
    const int np = g_source > 0 ? g_source : (int)points.size();
    vector<glm::vec3> init((size_t)components);
    for (int i = 0; i < components; ++i) init[i] = points[rand() % np];
    solveWithInit(init.data(), 1.0f, mean, weights, nullptr, iterations, N);
}

void scanRegistration::runSimulation(vector<glm::vec3>& source, vector<glm::vec3>& /*target*/) {
    const int components = 500;                      // gmm_kernels.cu:582 (the viewer's 800 is not what runs)
    GMM g1(components);
    vector<glm::vec3> mu((size_t)components);
    vector<float> weights((size_t)components);
    g1.solve(source, mu.data(), weights.data(), 10, g_source);      // :588
}

void scanRegistration::copyBoidsToVBO(float* vbodptr_positions, float* vbodptr_velocities) {
    CHECK(g_ctx, hgmm_fill_vbo(g_ctx, vbodptr_positions, vbodptr_velocities, 0.1f, nullptr, nullptr));     // scene_scale :22
}

void scanRegistration::endSimulation() {
    if (g_ctx) hgmm_destroy(g_ctx);
    g_ctx = nullptr;
}

// ------------------------------------------------------------------------------------------ GMMRegistration
GMMRegistration::GMMRegistration(int K) {
    numComponents = K;
    numSrcPc = numTargetPc = 0;
    dev_srcPc = dev_srcTransPc = dev_targetPc = dev_srcMu = dev_targetMu = nullptr;
    dev_srcPsi = dev_targetPsi = nullptr;
    engine = nullptr;
    for (int i = 0; i < 9; ++i) rot[i] = (i % 4 == 0) ? 1.0 : 0.0;
    trans[0] = trans[1] = trans[2] = 0.0;
}

void GMMRegistration::initSimulation(int N1, glm::vec3* src_pc, int N2, glm::vec3* target_pc) {
    hgmm_ctx* c = make_ctx();
    engine = c;
    numSrcPc = N1;
    numTargetPc = N2;
    CHECK(c, hgmm_set_points(c, reinterpret_cast<const float*>(src_pc), N1, HGMM_MEM_HOST));
    CHECK(c, hgmm_reg_set_target(c, reinterpret_cast<const float*>(target_pc), N2, HGMM_MEM_HOST));
    srcHost.assign(src_pc, src_pc + N1);
    treeBuilt = false;
}

void GMMRegistration::pointCloudRegisterGPU(float /*dt*/) {
    hgmm_ctx* c = static_cast<hgmm_ctx*>(engine);
    // Default: what the class's fields describe (numComponents means dev_srcMu / weights dev_srcPsi, gmm_reg.h:8-20) -- a FLAT
    // K-component full-covariance mixture of the source (10 EM iterations, the fit driver's count, gmm_kernels.cu:588) and the
    // weighted-Procrustes registration of the target against it (BASELINE configs[3]).  HGMM_SHIM_REG=tree selects the
    // hierarchical mixture + linearised twist solve of src/python/hgmm instead.
    const char* mode = getenv("HGMM_SHIM_REG");
    if (!(mode && strcmp(mode, "tree") == 0)) {
        if (!treeBuilt) {
            const int K = numComponents;
            vector<float> mu0((size_t)K * 3), cov0((size_t)K * 9, 0.f), w0((size_t)K, 1.0f / K);
            float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
            for (const glm::vec3& p : srcHost) {
                lo[0] = fminf(lo[0], p.x); hi[0] = fmaxf(hi[0], p.x);
                lo[1] = fminf(lo[1], p.y); hi[1] = fmaxf(hi[1], p.y);
                lo[2] = fminf(lo[2], p.z); hi[2] = fmaxf(hi[2], p.z);
            }
            const float ext = fmaxf(hi[0] - lo[0], fmaxf(hi[1] - lo[1], hi[2] - lo[2]));
            const float s0 = (ext / 16) * (ext / 16);          // data-scaled start (the fitter's Sigma = I is for unit-scale data)
            unsigned s = 72u;
            for (int j = 0; j < K; ++j) {                      // seeded draw of K source points (gmm_kernels.cu:374-379 uses rand())
                s = s * 1664525u + 1013904223u;
                const glm::vec3& p = srcHost[s % (unsigned)numSrcPc];
                mu0[3 * j] = p.x; mu0[3 * j + 1] = p.y; mu0[3 * j + 2] = p.z;
                cov0[9 * j] = cov0[9 * j + 4] = cov0[9 * j + 8] = s0;
            }
            hgmm_flat_config fc;
            memset(&fc, 0, sizeof fc);
            fc.n_components = K; fc.cov_type = HGMM_COV_FULL; fc.flavor = HGMM_FLAVOR_CPP; fc.max_iter = 10;
            CHECK(c, hgmm_fit_flat(c, &fc, mu0.data(), cov0.data(), w0.data(), nullptr, nullptr, nullptr, nullptr, nullptr, nullptr));
            treeBuilt = true;
        }
        hgmm_reg_config rc;
        rc.solver = HGMM_SOLVER_PROCRUSTES; rc.maxiter = 20; rc.tol = 1e-4f; rc.lambda_c = 0.f;
        double q = 0;
        int32_t it = 0;
        CHECK(c, hgmm_register_flat(c, &rc, rot, trans, &q, &it, nullptr));
        return;
    }
    // depth from the component budget K: the deepest tree whose leaf count does not exceed K (K=100 -> 64 leaves, L=2)
    int L = 1;
    while (L < 5 && 8 * (1 << (3 * L)) <= numComponents) ++L;
    if (!treeBuilt) {
        const int64_t nt = hgmm_tree_total_nodes(L);
        vector<float> init((size_t)nt * 3);
        // the reference's draw: points[randint(nTotal)] with a fixed seed (hgmm_gpu.py:469-470); here a fixed LCG
        unsigned s = 72u;
        for (int64_t i = 0; i < nt; ++i) {
            s = s * 1664525u + 1013904223u;
            const int idx = (int)(s % (unsigned)(nt < numSrcPc ? nt : numSrcPc));
            init[3 * i] = srcHost[idx].x; init[3 * i + 1] = srcHost[idx].y; init[3 * i + 2] = srcHost[idx].z;
        }
        // initial variance: (cloud extent / 8)^2, the data-scaled analogue of the reference's hard-coded 0.004
        float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
        for (const glm::vec3& p : srcHost) {
            lo[0] = fminf(lo[0], p.x); hi[0] = fmaxf(hi[0], p.x);
            lo[1] = fminf(lo[1], p.y); hi[1] = fmaxf(hi[1], p.y);
            lo[2] = fminf(lo[2], p.z); hi[2] = fmaxf(hi[2], p.z);
        }
        const float ext = fmaxf(hi[0] - lo[0], fmaxf(hi[1] - lo[1], hi[2] - lo[2]));
        hgmm_tree_config tc;
        memset(&tc, 0, sizeof tc);
        tc.max_level = L; tc.ll_mode = HGMM_LL_ESTEP; tc.ls = 20.f; tc.ld = 1e-4f; tc.sig2 = (ext / 8) * (ext / 8);
        tc.max_iters_per_level = 200;
        CHECK(c, hgmm_fit_tree(c, &tc, init.data(), nullptr, nullptr, nullptr, nullptr, nullptr, nullptr));
        treeBuilt = true;
    }
    hgmm_reg_config rc;
    rc.solver = HGMM_SOLVER_TWIST_LSTSQ; rc.maxiter = 20; rc.tol = 1e-4f; rc.lambda_c = 0.01f;     // hgmm_gpu.py:754 defaults
    double q = 0;
    int32_t it = 0;
    CHECK(c, hgmm_register_tree(c, &rc, rot, trans, &q, &it, nullptr));
}

void GMMRegistration::copyBoidsToVBO(float* vbodptr_positions, float* vbodptr_velocities) {
    const float white[3] = {1.f, 1.f, 1.f}, red[3] = {1.f, 0.f, 0.f};      // gmm_reg_kernels.cu:25-41
    CHECK(static_cast<hgmm_ctx*>(engine),
          hgmm_fill_vbo(static_cast<hgmm_ctx*>(engine), vbodptr_positions, vbodptr_velocities, 0.1f, white, red));
}

void GMMRegistration::endSimulation() {
    if (engine) hgmm_destroy(static_cast<hgmm_ctx*>(engine));
    engine = nullptr;
}
