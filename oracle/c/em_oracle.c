/* em_oracle.c -- plain-C (OpenMP) restatement of the flat full-covariance EM iteration.
 *
 * TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/__init__.py): used by tests/ as a second,
 * independent checker of oracle/flat_gmm.py::cpp_em_iteration and by bench.py as the timed CPU
 * baseline ("port": the reference has no CPU implementation of its full-covariance fitter).
 * Follows src/c++/gmm_fit/gmm_kernels.cu (reference checkout):
 *   :96-126   log N(x; mu, S) = -0.5 (3 log 2pi + log|S| + d^T S^-1 d)      (intended metric, not the :103 bug)
 *   :278-302  responsibilities normalised over components (max-shifted here)
 *   :156-210  pi_j = N_j / sum N_k ; mu_j = weighted mean ; S_j centred on the NEW mu_j
 * double precision; threads split the points, each with private accumulators, combined in a fixed order.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static void inv3(const double* s, double* inv, double* logdet) {
    double a = s[0], b = s[1], c = s[2], d = s[4], e = s[5], f = s[8];
    double A = d * f - e * e, B = c * e - b * f, C = b * e - c * d;
    double det = a * A + b * B + c * C;
    double r = 1.0 / det;
    inv[0] = A * r; inv[1] = B * r; inv[2] = C * r;
    inv[3] = (a * f - c * c) * r; inv[4] = (b * c - a * e) * r; inv[5] = (a * d - b * b) * r;
    *logdet = log(det);
}

void oracle_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* one EM iteration in place. x [n,3] float32; logpi [J], mu [J,3], cov [J,9] double (in/out).
 * returns sum_i log p(x_i) under the input parameters. */
double oracle_flat_em_iteration(const float* x, long n, int J, double* logpi, double* mu, double* cov) {
    double* P = (double*)malloc(sizeof(double) * (size_t)J * 7);        /* 6 inverse entries + constant */
    for (int j = 0; j < J; ++j) {
        double ld;
        inv3(cov + 9 * j, P + 7 * j, &ld);
        P[7 * j + 6] = logpi[j] - 0.5 * (3.0 * log(2.0 * M_PI) + ld);
    }
    int nth = oracle_num_threads();
    const int W = 10;
    double* part = (double*)calloc((size_t)nth * ((size_t)J * W + 1), sizeof(double));
#pragma omp parallel num_threads(nth)
    {
#ifdef _OPENMP
        int tid = omp_get_thread_num();
#else
        int tid = 0;
#endif
        double* acc = part + (size_t)tid * ((size_t)J * W + 1);
        double* q = (double*)malloc(sizeof(double) * (size_t)J);
        double ll = 0.0;
#pragma omp for schedule(static)
        for (long i = 0; i < n; ++i) {
            double X = x[3 * i], Y = x[3 * i + 1], Z = x[3 * i + 2];
            double m = -INFINITY;
            for (int j = 0; j < J; ++j) {
                const double* p = P + 7 * j;
                double dx = X - mu[3 * j], dy = Y - mu[3 * j + 1], dz = Z - mu[3 * j + 2];
                double maha = dx * (p[0] * dx + 2 * p[1] * dy + 2 * p[2] * dz) + dy * (p[3] * dy + 2 * p[4] * dz) + p[5] * dz * dz;
                q[j] = p[6] - 0.5 * maha;
                if (q[j] > m) m = q[j];
            }
            double s = 0.0;
            for (int j = 0; j < J; ++j) s += exp(q[j] - m);
            double lse = m + log(s);
            ll += lse;
            for (int j = 0; j < J; ++j) {
                double g = exp(q[j] - lse);
                double* a = acc + (size_t)j * W;
                a[0] += g; a[1] += g * X; a[2] += g * Y; a[3] += g * Z;
                a[4] += g * X * X; a[5] += g * X * Y; a[6] += g * X * Z; a[7] += g * Y * Y; a[8] += g * Y * Z; a[9] += g * Z * Z;
            }
        }
        acc[(size_t)J * W] = ll;
        free(q);
    }
    double ll = 0.0, total = 0.0;
    for (int t = 1; t < nth; ++t) {
        double* a0 = part;
        double* at = part + (size_t)t * ((size_t)J * W + 1);
        for (size_t k = 0; k < (size_t)J * W + 1; ++k) a0[k] += at[k];
    }
    ll = part[(size_t)J * W];
    for (int j = 0; j < J; ++j) total += part[(size_t)j * W];
    for (int j = 0; j < J; ++j) {
        const double* a = part + (size_t)j * W;
        double r = 1.0 / a[0];
        double mx = a[1] * r, my = a[2] * r, mz = a[3] * r;
        logpi[j] = log(a[0] / total);
        mu[3 * j] = mx; mu[3 * j + 1] = my; mu[3 * j + 2] = mz;
        double* c = cov + 9 * j;
        c[0] = a[4] * r - mx * mx; c[1] = c[3] = a[5] * r - mx * my; c[2] = c[6] = a[6] * r - mx * mz;
        c[4] = a[7] * r - my * my; c[5] = c[7] = a[8] * r - my * mz; c[8] = a[9] * r - mz * mz;
    }
    free(part);
    free(P);
    return ll;
}

/* GMM::solve (gmm_kernels.cu:371-504) with the init passed in: S = sigma0_sq I, pi = 1/J. ll [iters]. */
void oracle_flat_fit(const float* x, long n, int J, const float* mu0, double sigma0_sq, int iters, double* weights, double* mu,
                     double* cov, double* ll) {
    double* logpi = (double*)malloc(sizeof(double) * (size_t)J);
    for (int j = 0; j < J; ++j) {
        logpi[j] = -log((double)J);
        for (int k = 0; k < 3; ++k) mu[3 * j + k] = mu0[3 * j + k];
        memset(cov + 9 * j, 0, 9 * sizeof(double));
        cov[9 * j] = cov[9 * j + 4] = cov[9 * j + 8] = sigma0_sq;
    }
    for (int it = 0; it < iters; ++it) ll[it] = oracle_flat_em_iteration(x, n, J, logpi, mu, cov);
    for (int j = 0; j < J; ++j) weights[j] = exp(logpi[j]);
    free(logpi);
}
