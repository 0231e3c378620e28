// hgmm_shim.h -- the reference viewer's C++ fit/register API, re-declared verbatim so that
// src/c++/main.cpp and src/c++/main_reg.cpp link against libhgmm instead of the reference's
// gmm_fit/ and gmm_registration/ translation units.
//
//   class GMM { GMM(int N); void solve(std::vector<glm::vec3>, glm::vec3*, float*, int, int); }   src/c++/gmm_fit/gmm.h:10-19
//   namespace scanRegistration { initSimulation, runSimulation, copyBoidsToVBO, endSimulation }   src/c++/gmm_fit/gmm_kernels.h:13-18
//   class GMMRegistration { public fields + initSimulation, pointCloudRegisterGPU, copyBoidsToVBO, endSimulation }
//                                                                                                src/c++/gmm_registration/gmm_reg.h:5-32
//
// Build with the viewer's own GLM on the include path (-I<reference>/external/include); glm::vec3 is the
// packed 3 x fp32 the C ABI expects.  Without GLM (stand-alone tests) a layout-identical stand-in is used.
#pragma once
#include <vector>

#if defined(__has_include)
#if __has_include(<glm/glm.hpp>)
#include <glm/glm.hpp>
#define HGMM_HAVE_GLM 1
#endif
#endif
#ifndef HGMM_HAVE_GLM
namespace glm {
struct vec3 {
    float x, y, z;
    vec3() : x(0), y(0), z(0) {}
    vec3(float a, float b, float c) : x(a), y(b), z(c) {}
};
}  // namespace glm
#endif

using std::vector;      // the reference headers rely on `using namespace std` (gmm_kernels.h:11)

class GMM {
    int components;

public:
    GMM(int N) { components = N; }
    // fits `components` full-covariance Gaussians to the cloud registered by scanRegistration::initSimulation
    // (first N points); writes means and weights like the reference (covariances are not returned, gmm_kernels.cu:490-498)
    void solve(vector<glm::vec3> points, glm::vec3* mean, float* weights, int iterations, int N);
    // extension: same fit with explicit initial means and the covariances returned (row-major 3x3 per component)
    void solveWithInit(const glm::vec3* init_mean, float sigma0_sq, glm::vec3* mean, float* weights, float* covariances,
                       int iterations, int N);
};

namespace scanRegistration {
void initSimulation(vector<glm::vec3>& source, vector<glm::vec3>& target, int components);
void runSimulation(vector<glm::vec3>& source, vector<glm::vec3>& target);
void copyBoidsToVBO(float* vbodptr_positions, float* vbodptr_velocities);
void endSimulation();
}  // namespace scanRegistration

class GMMRegistration {
public:
    int numComponents;
    int numSrcPc;
    int numTargetPc;

    glm::vec3* dev_srcPc;
    glm::vec3* dev_srcTransPc;
    glm::vec3* dev_targetPc;

    glm::vec3* dev_srcMu;
    glm::vec3* dev_targetMu;

    float* dev_srcPsi;
    float* dev_targetPsi;

    GMMRegistration(int K);

    void initSimulation(int N1, glm::vec3* src_pc, int N2, glm::vec3* target_pc);
    // one registration solve (the reference body is empty, gmm_reg.cu:54-56): fits the K-component mixture of the source on
    // first use (flat full-covariance; HGMM_SHIM_REG=tree: the hierarchical one), registers the target against it and
    // updates the accumulated transform
    void pointCloudRegisterGPU(float dt);
    void copyBoidsToVBO(float* vbodptr_positions, float* vbodptr_velocities);
    void endSimulation();

    // results of the last pointCloudRegisterGPU (row-major rotation, translation): target -> source frame
    double rot[9];
    double trans[3];
    void* engine;       // hgmm_ctx*
    vector<glm::vec3> srcHost;
    bool treeBuilt;
};
