// shim_demo.cpp -- drives the reference viewer's call sequence (main.cpp:212,305,126; main_reg.cpp:269,360,193)
// headlessly on a synthetic two-blob cloud and prints the fitted model; used by tests/test_shim.py on the GPU box.
#include <cmath>
#include <cstdio>
#include <vector>

#include "hgmm_shim.h"

int main() {
    std::vector<glm::vec3> src, tgt;
    unsigned s = 12345u;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((s >> 8) & 0xffff) / 65536.0f - 0.5f; };
    for (int i = 0; i < 4000; ++i) {
        const float cx = (i & 1) ? 0.5f : -0.5f;
        src.push_back(glm::vec3(cx + 0.1f * (rnd() + rnd()), 0.1f * (rnd() + rnd()), 0.1f * (rnd() + rnd())));
    }
    const float th = 0.1f;
    for (const glm::vec3& p : src) tgt.push_back(glm::vec3(cosf(th) * p.x - sinf(th) * p.y + 0.01f, sinf(th) * p.x + cosf(th) * p.y, p.z));

    scanRegistration::initSimulation(src, tgt, 2);
    GMM g(2);
    glm::vec3 init[2] = {glm::vec3(-0.3f, 0.f, 0.f), glm::vec3(0.3f, 0.f, 0.f)};
    glm::vec3 mean[2];
    float w[2], cov[18];
    g.solveWithInit(init, 0.05f, mean, w, cov, 10, (int)src.size());
    printf("FLAT %f %f %f %f %f %f %f %f\n", mean[0].x, mean[0].y, mean[0].z, mean[1].x, mean[1].y, mean[1].z, w[0], w[1]);
    scanRegistration::runSimulation(src, tgt);        // the viewer's fixed 500-component / 10-iteration fit
    scanRegistration::endSimulation();

    GMMRegistration reg(100);
    reg.initSimulation((int)src.size(), src.data(), (int)tgt.size(), tgt.data());
    reg.pointCloudRegisterGPU(0.f);
    printf("REG %f %f %f %f\n", reg.rot[0], reg.rot[1], reg.rot[3], reg.rot[4]);
    reg.endSimulation();
    return 0;
}
