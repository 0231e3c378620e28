// ref_driver_main.cu -- headless driver for the REFERENCE's own CUDA fitter (src/c++/gmm_fit/gmm_kernels.cu),
// compiled by oracle/build_ref.sh from the sources where they lie under /root/reference.
// TEST / BASELINE INFRASTRUCTURE ONLY.  Usage: ref_gmm_cuda <points.npy> <J> <iterations> <out.bin>
// Calls, like the viewer (main.cpp:212,305): scanRegistration::initSimulation -> GMM(J).solve(...).
// The reference seeds nothing (rand() with glibc's default seed 1, gmm_kernels.cu:375), so the run is
// reproducible; it prints its own "Time elapsed: %.4f" (ms, cudaEvent around the EM loop, :483-488).
// out.bin: J x 3 float32 means followed by J float32 weights.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <glm/glm.hpp>
using namespace std;
#include "gmm_fit/gmm_kernels.h"
#include "gmm_fit/gmm.h"

static vector<glm::vec3> load_npy(const char* path) {
    FILE* f = fopen(path, "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(2); }
    unsigned char hdr[10];
    if (fread(hdr, 1, 10, f) != 10 || memcmp(hdr, "\x93NUMPY", 6) != 0) { fprintf(stderr, "not an npy file\n"); exit(2); }
    const int hlen = hdr[8] | (hdr[9] << 8);
    vector<char> h(hlen + 1, 0);
    if (fread(h.data(), 1, hlen, f) != (size_t)hlen) exit(2);
    if (!strstr(h.data(), "'<f4'") || strstr(h.data(), "'fortran_order': True")) { fprintf(stderr, "need C-order float32\n"); exit(2); }
    const char* sh = strstr(h.data(), "'shape': (");
    long n = atol(sh + 10);
    vector<glm::vec3> pts((size_t)n);
    if (fread(pts.data(), sizeof(glm::vec3), (size_t)n, f) != (size_t)n) { fprintf(stderr, "short read\n"); exit(2); }
    fclose(f);
    return pts;
}

int main(int argc, char** argv) {
    if (argc < 5) { fprintf(stderr, "usage: %s points.npy J iterations out.bin\n", argv[0]); return 2; }
    vector<glm::vec3> src = load_npy(argv[1]);
    vector<glm::vec3> tgt(src.begin(), src.begin() + 1);
    const int J = atoi(argv[2]), iters = atoi(argv[3]);
    scanRegistration::initSimulation(src, tgt, J);
    GMM g(J);
    vector<glm::vec3> mu((size_t)J);
    vector<float> w((size_t)J);
    g.solve(src, mu.data(), w.data(), iters, (int)src.size());
    scanRegistration::endSimulation();
    FILE* o = fopen(argv[4], "wb");
    fwrite(mu.data(), sizeof(glm::vec3), (size_t)J, o);
    fwrite(w.data(), sizeof(float), (size_t)J, o);
    fclose(o);
    return 0;
}
