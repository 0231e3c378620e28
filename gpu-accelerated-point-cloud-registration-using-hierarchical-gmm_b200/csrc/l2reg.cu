// l2reg.cu -- L2-distance registration of two flat mixtures.
//
// Replaces, on the device (paths relative to the reference checkout):
//   compute_l2_dist, RigidCostFunction.__call__   src/python/gmmreg_gpu/cost_functions.py:29-69
//   GaussTransform / _gauss_transform_direct      src/python/gmmreg_gpu/transforms.py:43-86   (J x J direct sum)
//   diff_rot_from_quaternion                      src/python/gmmreg_gpu/so.py:4-59
//   the BFGS minimisation of gmmreg.py:101-107    (scipy.optimize.minimize on the host in the reference)
//
// The mixtures are reduced to isotropic kernels of width sigma at the fitted means (the reference discards the
// covariances, gmmreg_gpu/gmm.py:57), so with T(mu) = R(q) mu + t, c_j = phi_t[j] / (2 pi sigma^2)^1.5 and
// k_ij = exp(-|T(mu_s[i]) - mu_t[j]|^2 / (2 sigma^2)):
//     f      = - sum_ij phi_s[i] c_j k_ij
//     G_i    =   phi_s[i] sum_j c_j k_ij (T(mu_s[i]) - mu_t[j]) / (2 sigma^2)
//     grad_t = sum_i G_i,   grad_q[k] = sum_ab (sum_i G_i[a] mu_s[i][b]) dR[k][a][b]
// i.e. 13 sums over the Js x Jt pairs.  Everything is float64 (the reference computes this path in float64 and BFGS
// is sensitive to the gradient's low bits); one CTA evaluates the pairs and folds the 13 sums in a fixed order.
// l2_bfgs_kernel keeps the whole minimisation in ONE launch: BFGS with a strong-Wolfe line search (Nocedal & Wright
// alg. 3.5/3.6, c1 = 1e-4, c2 = 0.9 as SciPy), every thread running the same scalar control flow and all threads
// evaluating the cost together -- the reference pays a Python call, ~2 Js NumPy passes and a host optimiser step per
// evaluation.
#include "common.cuh"
#include "kernels.h"

namespace hgmm {

constexpr int kL2Threads = 256;
constexpr int kL2Sums = 13;

struct L2Mix {
    const double* mu_s;    // [Js,3]
    const double* phi_s;   // [Js]
    const double* mu_t;    // [Jt,3]
    const double* phi_t;   // [Jt]
    int Js, Jt;
};

// transformations.quaternion_matrix (q = w,x,y,z; normalised internally)
__device__ __forceinline__ void quat_to_rot(const double* q, double R[3][3]) {
    const double n = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
    if (n < 8.881784197001252e-16) {     // 4 * DBL_EPSILON
        R[0][0] = R[1][1] = R[2][2] = 1.0;
        R[0][1] = R[0][2] = R[1][0] = R[1][2] = R[2][0] = R[2][1] = 0.0;
        return;
    }
    const double s = sqrt(2.0 / n);
    const double w = q[0] * s, x = q[1] * s, y = q[2] * s, z = q[3] * s;
    R[0][0] = 1.0 - y * y - z * z; R[0][1] = x * y - z * w;       R[0][2] = x * z + y * w;
    R[1][0] = x * y + z * w;       R[1][1] = 1.0 - x * x - z * z; R[1][2] = y * z - x * w;
    R[2][0] = x * z - y * w;       R[2][1] = y * z + x * w;       R[2][2] = 1.0 - x * x - y * y;
}

// so.py:4-59, reproduced entry by entry
__device__ void diff_rot(const double* q, const double R[3][3], double d[4][3][3]) {
    const double q2[4] = {q[0] * q[0], q[1] * q[1], q[2] * q[2], q[3] * q[3]};
    const double z = q2[0] + q2[1] + q2[2] + q2[3], z2 = z * z;
    d[0][0][0] = 4 * q[0] * (q2[2] + q2[3]) / z2;
    d[1][0][0] = 4 * q[1] * (q2[2] + q2[3]) / z2;
    d[2][0][0] = -4 * q[2] * (q2[1] + q2[0]) / z2;
    d[3][0][0] = -4 * q[3] * (q2[1] + q2[0]) / z2;
    d[0][1][1] = 4 * q[0] * (q2[1] + q2[3]) / z2;
    d[1][1][1] = -4 * q[1] * (q2[2] + q2[0]) / z2;
    d[2][1][1] = 4 * q[2] * (q2[1] + q2[3]) / z2;
    d[3][1][1] = -4 * q[3] * (q2[2] + q2[0]) / z2;
    d[0][2][2] = 4 * q[0] * (q2[1] + q2[2]) / z2;
    d[1][2][2] = -4 * q[1] * (q2[3] + q2[0]) / z2;
    d[2][2][2] = -4 * q[2] * (q2[1] + q2[2]) / z2;
    d[3][2][2] = 4 * q[3] * (q2[3] + q2[0]) / z2;
    // off-diagonals: sgn * 2 q[m] / z - 2 q[k] R[a][b] / z2, (sgn, m) per k
#define HGMM_OFF(a, b, s0, m0, s1, m1, s2, m2, s3, m3)                    \
    d[0][a][b] = (s0) * 2 * q[m0] / z - 2 * q[0] * R[a][b] / z2;          \
    d[1][a][b] = (s1) * 2 * q[m1] / z - 2 * q[1] * R[a][b] / z2;          \
    d[2][a][b] = (s2) * 2 * q[m2] / z - 2 * q[2] * R[a][b] / z2;          \
    d[3][a][b] = (s3) * 2 * q[m3] / z - 2 * q[3] * R[a][b] / z2;
    HGMM_OFF(0, 1, -1, 3, 1, 2, 1, 1, -1, 0)
    HGMM_OFF(0, 2, 1, 2, 1, 3, 1, 0, 1, 1)
    HGMM_OFF(1, 0, 1, 3, 1, 2, 1, 1, 1, 0)
    HGMM_OFF(1, 2, -1, 1, -1, 0, 1, 3, 1, 2)
    HGMM_OFF(2, 0, -1, 2, 1, 3, -1, 0, 1, 1)
    HGMM_OFF(2, 1, 1, 1, 1, 0, 1, 3, 1, 2)
#undef HGMM_OFF
}

// Cooperative evaluation by the whole CTA: every thread passes the same theta and receives the same (f, grad).
// s_red: [kL2Sums][kL2Threads] doubles of shared memory.
__device__ void l2_eval(const L2Mix& mx, const double* th, double sigma, double* s_red, double& f, double* grad) {
    const int tid = threadIdx.x;
    double R[3][3];
    quat_to_rot(th, R);
    const double two_s2 = 2.0 * sigma * sigma;
    const double zc = 1.0 / pow(2.0 * 3.141592653589793 * sigma * sigma, 1.5);
    double acc[kL2Sums];
#pragma unroll
    for (int k = 0; k < kL2Sums; ++k) acc[k] = 0.0;
    const long long npairs = (long long)mx.Js * mx.Jt;
    for (long long p = tid; p < npairs; p += kL2Threads) {
        const int i = (int)(p / mx.Jt), j = (int)(p - (long long)i * mx.Jt);
        const double sx = mx.mu_s[3 * i], sy = mx.mu_s[3 * i + 1], sz = mx.mu_s[3 * i + 2];
        const double tx = R[0][0] * sx + R[0][1] * sy + R[0][2] * sz + th[4];
        const double ty = R[1][0] * sx + R[1][1] * sy + R[1][2] * sz + th[5];
        const double tz = R[2][0] * sx + R[2][1] * sy + R[2][2] * sz + th[6];
        const double dx = tx - mx.mu_t[3 * j], dy = ty - mx.mu_t[3 * j + 1], dz = tz - mx.mu_t[3 * j + 2];
        const double w = mx.phi_s[i] * mx.phi_t[j] * zc * exp(-(dx * dx + dy * dy + dz * dz) / two_s2);
        const double vx = w * dx / two_s2, vy = w * dy / two_s2, vz = w * dz / two_s2;
        acc[0] += w;
        acc[1] += vx; acc[2] += vy; acc[3] += vz;
        acc[4] += vx * sx; acc[5] += vx * sy; acc[6] += vx * sz;
        acc[7] += vy * sx; acc[8] += vy * sy; acc[9] += vy * sz;
        acc[10] += vz * sx; acc[11] += vz * sy; acc[12] += vz * sz;
    }
    __syncthreads();                                   // previous readers of s_red are done
#pragma unroll
    for (int k = 0; k < kL2Sums; ++k) s_red[k * kL2Threads + tid] = acc[k];
    __syncthreads();
    for (int o = kL2Threads / 2; o > 0; o >>= 1) {     // fixed-order tree: deterministic
        if (tid < o) {
#pragma unroll
            for (int k = 0; k < kL2Sums; ++k) s_red[k * kL2Threads + tid] += s_red[k * kL2Threads + tid + o];
        }
        __syncthreads();
    }
    double S[kL2Sums];
#pragma unroll
    for (int k = 0; k < kL2Sums; ++k) S[k] = s_red[k * kL2Threads];
    f = -S[0];
    double d[4][3][3];
    diff_rot(th, R, d);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        double g = 0.0;
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) g += S[4 + 3 * a + b] * d[k][a][b];
        grad[k] = g;
    }
    grad[4] = S[1]; grad[5] = S[2]; grad[6] = S[3];
}

// out[0] = f, out[1..7] = grad
__global__ void __launch_bounds__(kL2Threads) l2_cost_grad_kernel(L2Mix mx, const double* __restrict__ theta, double sigma,
                                                                  double* __restrict__ out) {
    __shared__ double s_red[kL2Sums * kL2Threads];
    double th[7], g[7], f;
#pragma unroll
    for (int k = 0; k < 7; ++k) th[k] = theta[k];
    l2_eval(mx, th, sigma, s_red, f, g);
    if (threadIdx.x == 0) {
        out[0] = f;
#pragma unroll
        for (int k = 0; k < 7; ++k) out[1 + k] = g[k];
    }
}

// theta: in/out.  out: [0] f, [1] iterations, [2] evaluations, [3] status (0 converged on gtol, 1 iteration limit,
// 2 line search failed), [4..10] gradient at the end point.
__global__ void __launch_bounds__(kL2Threads) l2_bfgs_kernel(L2Mix mx, double* __restrict__ theta, double sigma, int max_iter,
                                                             double gtol, double* __restrict__ out) {
    __shared__ double s_red[kL2Sums * kL2Threads];
    __shared__ double H[7][7];
    const int tid = threadIdx.x;
    double x[7], g[7], f;
#pragma unroll
    for (int k = 0; k < 7; ++k) x[k] = theta[k];
    l2_eval(mx, x, sigma, s_red, f, g);
    int nfev = 1, iters = 0, status = 1;
    bool fresh = true;                                  // H is the identity
    if (tid < 49) H[tid / 7][tid % 7] = (tid / 7 == tid % 7) ? 1.0 : 0.0;
    __syncthreads();
    double gn2 = 0.0;
#pragma unroll
    for (int k = 0; k < 7; ++k) gn2 += g[k] * g[k];
    double f_old = f + sqrt(gn2) / 2.0;                 // SciPy's first-step heuristic
    const double c1 = 1e-4, c2 = 0.9;
    for (int it = 0; it < max_iter; ++it) {
        double gmax = 0.0;
#pragma unroll
        for (int k = 0; k < 7; ++k) gmax = fmax(gmax, fabs(g[k]));
        if (!(gmax > gtol)) { status = 0; break; }
        double p[7];
        double dphi0 = 0.0;
#pragma unroll
        for (int a = 0; a < 7; ++a) {
            double v = 0.0;
#pragma unroll
            for (int b = 0; b < 7; ++b) v -= H[a][b] * g[b];
            p[a] = v;
            dphi0 += v * g[a];
        }
        if (!(dphi0 < 0.0)) {                           // not a descent direction: restart from steepest descent
            __syncthreads();
            if (tid < 49) H[tid / 7][tid % 7] = (tid / 7 == tid % 7) ? 1.0 : 0.0;
            __syncthreads();
            dphi0 = 0.0;
#pragma unroll
            for (int a = 0; a < 7; ++a) { p[a] = -g[a]; dphi0 -= g[a] * g[a]; }
            fresh = true;
        }
        double a1 = fmin(1.0, 1.01 * 2.0 * (f - f_old) / dphi0);
        if (!(a1 > 0.0)) a1 = 1.0;
        // ---- strong-Wolfe line search along p
        double a_lo = 0.0, phi_lo = f, dphi_lo = dphi0, a_hi = 0.0, phi_hi = f;
        double a_prev = 0.0, phi_prev = f, dphi_prev = dphi0;
        double a_acc = 0.0, f_acc = f, g_acc[7];
        bool accepted = false, bracketed = false;
        double a_cur = a1;
        for (int ls = 0; ls < 12 && !accepted && !bracketed; ++ls) {
            double xt[7], gt[7], ft;
#pragma unroll
            for (int k = 0; k < 7; ++k) xt[k] = x[k] + a_cur * p[k];
            l2_eval(mx, xt, sigma, s_red, ft, gt);
            ++nfev;
            double dphi = 0.0;
#pragma unroll
            for (int k = 0; k < 7; ++k) dphi += gt[k] * p[k];
            if (!(ft <= f + c1 * a_cur * dphi0) || (ls > 0 && ft >= phi_prev)) {
                a_lo = a_prev; phi_lo = phi_prev; dphi_lo = dphi_prev; a_hi = a_cur; phi_hi = ft;
                bracketed = true;
            } else if (fabs(dphi) <= -c2 * dphi0) {
                a_acc = a_cur; f_acc = ft;
#pragma unroll
                for (int k = 0; k < 7; ++k) g_acc[k] = gt[k];
                accepted = true;
            } else if (dphi >= 0.0) {
                a_lo = a_cur; phi_lo = ft; dphi_lo = dphi; a_hi = a_prev; phi_hi = phi_prev;
                bracketed = true;
            } else {
                a_prev = a_cur; phi_prev = ft; dphi_prev = dphi;
                a_cur *= 2.0;
            }
        }
        if (bracketed) {
            for (int zi = 0; zi < 16 && !accepted; ++zi) {
                const double dd = a_hi - a_lo;
                double aq = a_lo - 0.5 * dphi_lo * dd * dd / (phi_hi - phi_lo - dphi_lo * dd);   // minimiser of the quadratic
                const double lo_b = fmin(a_lo, a_hi) + 0.1 * fabs(dd), hi_b = fmax(a_lo, a_hi) - 0.1 * fabs(dd);
                if (!(aq >= lo_b && aq <= hi_b)) aq = a_lo + 0.5 * dd;
                double xt[7], gt[7], ft;
#pragma unroll
                for (int k = 0; k < 7; ++k) xt[k] = x[k] + aq * p[k];
                l2_eval(mx, xt, sigma, s_red, ft, gt);
                ++nfev;
                double dphi = 0.0;
#pragma unroll
                for (int k = 0; k < 7; ++k) dphi += gt[k] * p[k];
                if (!(ft <= f + c1 * aq * dphi0) || ft >= phi_lo) {
                    a_hi = aq; phi_hi = ft;
                } else {
                    if (fabs(dphi) <= -c2 * dphi0) {
                        a_acc = aq; f_acc = ft;
#pragma unroll
                        for (int k = 0; k < 7; ++k) g_acc[k] = gt[k];
                        accepted = true;
                    } else {
                        if (dphi * (a_hi - a_lo) >= 0.0) { a_hi = a_lo; phi_hi = phi_lo; }
                        a_lo = aq; phi_lo = ft; dphi_lo = dphi;
                    }
                }
                if (fabs(a_hi - a_lo) < 1e-16 * fmax(1.0, fabs(a_lo))) break;
            }
        }
        if (!accepted) {
            // The reference's gradient is not the cost's derivative (so.py's diagonal entries, the halved translation part:
            // cost_functions.py:39), so the Wolfe conditions can be unsatisfiable along p.  Keep the best Armijo point the
            // search found; failing that, retry once from steepest descent; only then give up (SciPy's status 2).
            if (a_lo > 0.0 && phi_lo < f) {
                double xt[7];
#pragma unroll
                for (int k = 0; k < 7; ++k) xt[k] = x[k] + a_lo * p[k];
                l2_eval(mx, xt, sigma, s_red, f_acc, g_acc);
                ++nfev;
                a_acc = a_lo;
            } else if (!fresh) {
                __syncthreads();
                if (tid < 49) H[tid / 7][tid % 7] = (tid / 7 == tid % 7) ? 1.0 : 0.0;
                __syncthreads();
                fresh = true;
                gn2 = 0.0;
#pragma unroll
                for (int k = 0; k < 7; ++k) gn2 += g[k] * g[k];
                f_old = f + sqrt(gn2) / 2.0;
                continue;
            } else {
                status = 2;
                break;
            }
        }
        // ---- BFGS update of the inverse Hessian: H' = H - rho (s (Hy)^T + (Hy) s^T) + (rho^2 y^T H y + rho) s s^T
        double s[7], y[7], ys = 0.0;
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            s[k] = a_acc * p[k];
            y[k] = g_acc[k] - g[k];
            ys += y[k] * s[k];
        }
        const double rho = ys != 0.0 ? 1.0 / ys : 1000.0;
        double Hy[7], yHy = 0.0;
#pragma unroll
        for (int a = 0; a < 7; ++a) {
            double v = 0.0;
#pragma unroll
            for (int b = 0; b < 7; ++b) v += H[a][b] * y[b];
            Hy[a] = v;
            yHy += v * y[a];
        }
        __syncthreads();                                 // every thread has read the old H
        if (tid < 49) {
            const int a = tid / 7, b = tid % 7;
            double sa = 0, sb = 0, ha = 0, hb = 0;
#pragma unroll
            for (int k = 0; k < 7; ++k) {
                if (k == a) { sa = s[k]; ha = Hy[k]; }
                if (k == b) { sb = s[k]; hb = Hy[k]; }
            }
            H[a][b] = H[a][b] - rho * (sa * hb + ha * sb) + (rho * rho * yHy + rho) * sa * sb;
        }
        __syncthreads();
        fresh = false;
        f_old = f;
        f = f_acc;
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            x[k] += s[k];
            g[k] = g_acc[k];
        }
        iters = it + 1;
    }
    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < 7; ++k) theta[k] = x[k];
        out[0] = f;
        out[1] = (double)iters;
        out[2] = (double)nfev;
        out[3] = (double)status;
#pragma unroll
        for (int k = 0; k < 7; ++k) out[4 + k] = g[k];
    }
}

cudaError_t launch_l2_cost_grad(const double* mu_s, const double* phi_s, int Js, const double* mu_t, const double* phi_t, int Jt,
                                const double* theta, double sigma, double* out, cudaStream_t s) {
    L2Mix mx{mu_s, phi_s, mu_t, phi_t, Js, Jt};
    l2_cost_grad_kernel<<<1, kL2Threads, 0, s>>>(mx, theta, sigma, out);
    return cudaGetLastError();
}

cudaError_t launch_l2_bfgs(const double* mu_s, const double* phi_s, int Js, const double* mu_t, const double* phi_t, int Jt,
                           double* theta, double sigma, int max_iter, double gtol, double* out, cudaStream_t s) {
    L2Mix mx{mu_s, phi_s, mu_t, phi_t, Js, Jt};
    l2_bfgs_kernel<<<1, kL2Threads, 0, s>>>(mx, theta, sigma, max_iter, gtol, out);
    return cudaGetLastError();
}

}  // namespace hgmm
