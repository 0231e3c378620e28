// common.cuh -- shared device helpers for libhgmm (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

namespace hgmm {

// ------------------------------------------------------------------------------------------
// One mixture component as the E-step consumes it: 12 floats = 3 x LDS.128 / one 48-byte TMA row.
//   q2(x) = c2 + d^T A d,  d = x - m,  A = -0.5*log2(e)*Sigma^-1  (off-diagonals pre-doubled)
// is log2( pi * N(x; m, Sigma) ), so responsibilities are ex2(q2 - lse2).  A dead component
// (blank node, det < 1e-15, pi = 0) has c2 = -inf and A = 0.
// ------------------------------------------------------------------------------------------
struct __align__(16) PackedComp {
    float mx, my, mz, c2;
    float axx, ayy, azz, axy;
    float axz, ayz, pad0, pad1;
};
static_assert(sizeof(PackedComp) == 48, "PackedComp must be 48 bytes");

constexpr int kMom = 10;                 // M0, M1(3), M2(6: xx xy xz yy yz zz), centred on the component mean
constexpr int kAccHdr = 4;               // doubles in front of the moment block: [0] sum log-lik, [1..3] spare
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kNegBig = -3.0e38f;      // finite stand-in for -inf in running maxima
constexpr float kLog2Eps15 = -49.828921423310435f;   // log2(1e-15), the hgmm files' `eps`
constexpr float kLog2Eps8 = -26.575424759098897f;    // log2(1e-8), gmm_impl.py's `eps`

__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2f(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// q2 = c2 + d^T A d with the packed symmetric form: 3 FADD + 9 FMA/FMUL.
__device__ __forceinline__ float quad_q2(const float4& p0, const float4& p1, const float2& p2,
                                         float x, float y, float z, float& dx, float& dy, float& dz) {
    dx = x - p0.x;
    dy = y - p0.y;
    dz = z - p0.z;
    float t0 = p2.x * dz;            // axz*dz
    t0 = fmaf(p1.w, dy, t0);         // + axy*dy
    t0 = fmaf(p1.x, dx, t0);         // + axx*dx
    float t1 = p2.y * dz;            // ayz*dz
    t1 = fmaf(p1.y, dy, t1);         // + ayy*dy
    float t2 = p1.z * dz;            // azz*dz
    float q = fmaf(dz, t2, p0.w);
    q = fmaf(dy, t1, q);
    q = fmaf(dx, t0, q);
    return q;
}

// ------------------------------------------------------------------------------------------
// mbarrier + 1-D bulk async copy (TMA engine, SASS UBLKCP) global -> shared
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// bytes must be a multiple of 16; src/dst 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ------------------------------------------------------------------------------------------
// small double-precision 3x3 helpers (symmetric input as xx xy xz yy yz zz)
// ------------------------------------------------------------------------------------------
struct Sym3 {
    double xx, xy, xz, yy, yz, zz;
};
__host__ __device__ __forceinline__ double sym3_det(const Sym3& s) {
    return s.xx * (s.yy * s.zz - s.yz * s.yz) - s.xy * (s.xy * s.zz - s.yz * s.xz) + s.xz * (s.xy * s.yz - s.yy * s.xz);
}
__host__ __device__ __forceinline__ Sym3 sym3_adj(const Sym3& s) {   // adjugate (= det * inverse)
    Sym3 a;
    a.xx = s.yy * s.zz - s.yz * s.yz;
    a.xy = s.xz * s.yz - s.xy * s.zz;
    a.xz = s.xy * s.yz - s.xz * s.yy;
    a.yy = s.xx * s.zz - s.xz * s.xz;
    a.yz = s.xy * s.xz - s.xx * s.yz;
    a.zz = s.xx * s.yy - s.xy * s.xy;
    return a;
}

// Pack (pi-term, mean, covariance) -> PackedComp.  logw = natural log of the mixing term to fold in
// (-inf kills the component).  If `use_sigma_as_metric` the quadratic form uses Sigma itself
// (gmm_kernels.cu:97-103 bug reproduction).  det_floor: det < det_floor => dead (hgmm files: 1e-15).
__device__ __forceinline__ PackedComp pack_full(double logw, double mx, double my, double mz, const Sym3& cov,
                                                bool use_sigma_as_metric, double det_floor) {
    PackedComp p;
    p.mx = (float)mx;
    p.my = (float)my;
    p.mz = (float)mz;
    p.pad0 = 0.f;
    p.pad1 = 0.f;
    double det = sym3_det(cov);
    bool dead = !(det >= det_floor) || !(det > 0.0) || !(logw > -1.0e300);
    if (dead) {
        p.c2 = -INFINITY;
        p.axx = p.ayy = p.azz = p.axy = p.axz = p.ayz = 0.f;
        return p;
    }
    Sym3 m;
    if (use_sigma_as_metric) {
        m = cov;
    } else {
        Sym3 a = sym3_adj(cov);
        double r = 1.0 / det;
        m.xx = a.xx * r; m.xy = a.xy * r; m.xz = a.xz * r; m.yy = a.yy * r; m.yz = a.yz * r; m.zz = a.zz * r;
    }
    const double h = -0.5 * 1.4426950408889634;
    p.axx = (float)(h * m.xx);
    p.ayy = (float)(h * m.yy);
    p.azz = (float)(h * m.zz);
    p.axy = (float)(2.0 * h * m.xy);
    p.axz = (float)(2.0 * h * m.xz);
    p.ayz = (float)(2.0 * h * m.yz);
    p.c2 = (float)(1.4426950408889634 * (logw - 0.5 * log(det) - 1.5 * 1.8378770664093453));   // 1.5*log(2 pi)
    return p;
}

}  // namespace hgmm
