/*
 * ssn_oracle.c -- CPU restatement of the reference SSN fixed-point solver.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under tc_gan_b200/ may import, link or
 * call this file; it exists so that tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py have an independent
 * answer to compare the CUDA path against.
 *
 * What it restates (all citations are into /root/reference/tc_gan):
 *   - transfer functions              ext/ssnode.c:21-53  (rate_to_volt, io_pow,
 *                                     io_alin, io_atanh)
 *   - one Jacobi forward-Euler sweep  ext/ssnode.c:64-67  (ODE_STEP) with
 *                                     dt/tau_E on rows [0,N) and dt/tau_I on
 *                                     rows [N,2N)            ext/ssnode.c:72-82
 *   - stopping rule                   ext/ssnode.c:84-105: first "every
 *                                     |r_new-r_old| < atol" -> 0 (result is the
 *                                     NEW state); then, power/linear only,
 *                                     "any r_new >= hard bound" -> 2; running
 *                                     out of max_iter -> 1.  The tanh variant
 *                                     has no hard-bound exit (ext/ssnode.c:153-187).
 *   - network-level rejection rule    ssnode.py:390-420 (stimuli visited last to
 *                                     first, a network stops at its first
 *                                     failing stimulus).
 *
 * Pinned against: oracle/_ref/libssnode.so (the unmodified reference C file
 * compiled by oracle/Makefile) and the reference's MATLAB golden vectors, see
 * tests/test_oracle.py.
 *
 * Written from the algorithm description above, organised as one generic
 * routine over an io_type switch with a row-blocked mat-vec; it is not a copy
 * of the reference source.
 */
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

enum { IO_POWER = 0, IO_LINEAR = 1, IO_TANH = 2 };

double oracle_rate_to_volt(double rate, double k, double n)
{
    return pow(rate / k, 1.0 / n);
}

/* f(v): 0 below threshold, k v^n up to v0, then a per-type continuation. */
double oracle_io(int io_type, double v, double r_soft, double r_hard,
                 double v0, double k, double n)
{
    if (!(v > 0.0))
        return v != v ? v : 0.0;            /* NaN propagates like pow() would */
    if (io_type == IO_POWER || v <= v0)
        return k * pow(v, n);
    if (io_type == IO_LINEAR)
        return r_soft + k * pow(v0, n - 1.0) * n * (v - v0);
    {
        const double span = r_hard - r_soft;
        return r_soft + span * tanh(n * r_soft / span * (v - v0) / v0);
    }
}

/* f'(v), the gain used by the implicit gradient (SS_grad.py:78-99). */
double oracle_io_gain(int io_type, double v, double r_soft, double r_hard,
                      double v0, double k, double n)
{
    const double vc = v > 0.0 ? v : 0.0;
    if (io_type == IO_POWER)
        return n * k * pow(vc, n - 1.0);
    if (io_type == IO_LINEAR) {
        const double vv = vc < v0 ? vc : v0;
        return n * k * pow(vv, n - 1.0);
    }
    if (vc <= v0)
        return n * k * pow(vc, n - 1.0);
    {
        const double a = (n * r_soft / v0) * (vc - v0) / (r_hard - r_soft);
        const double c = cosh(a);
        return (n * r_soft / v0) / (c * c);
    }
}

/*
 * One (network, stimulus) solve.  `r` holds the initial state on entry and the
 * final state on exit; `scratch` is 2N doubles.  Returns 0/1/2 as above and
 * stores the number of sweeps performed in *iters (may be NULL).
 */
int oracle_fixed_point(int io_type, int n_sites,
                       const double *W, const double *ext,
                       double k, double n, double *r, double *scratch,
                       double tau_E, double tau_I, double dt,
                       int max_iter, double atol,
                       double r_soft, double r_hard, int *iters)
{
    const int dim = 2 * n_sites;
    const double v0 = oracle_rate_to_volt(r_soft, k, n);
    const double step_E = dt / tau_E, step_I = dt / tau_I;
    double *cur = r, *nxt = scratch;
    int it, code = 1;

    for (it = 0; it < max_iter; ++it) {
        int moving = 0, above = 0;
        for (int i = 0; i < dim; ++i) {
            const double *row = W + (size_t)i * dim;
            double acc = 0.0;
#pragma omp simd reduction(+ : acc)
            for (int j = 0; j < dim; ++j)
                acc += row[j] * cur[j];
            const double drive = oracle_io(io_type, acc + ext[i],
                                           r_soft, r_hard, v0, k, n);
            nxt[i] = cur[i] + (drive - cur[i]) * (i < n_sites ? step_E : step_I);
        }
        for (int i = 0; i < dim; ++i) {
            if (fabs(nxt[i] - cur[i]) >= atol)
                moving = 1;
            if (nxt[i] >= r_hard)
                above = 1;
        }
        { double *t = cur; cur = nxt; nxt = t; }   /* cur is now the new state */
        if (!moving) { code = 0; ++it; break; }
        if (above && io_type != IO_TANH) { code = 2; ++it; break; }
    }
    if (cur != r)
        memcpy(r, cur, sizeof(double) * dim);
    if (iters)
        *iters = it;
    return code;
}

/*
 * Batched driver: nz networks (W[nz][2N][2N]) times nb stimuli (ext[nb][2N]),
 * every solve started from r = 0.  Stimuli are visited from the last to the
 * first and a network stops at its first failure (ssnode.py:393-408); the
 * solves it skipped keep status -1.  R[nz][nb][2N], status[nz][nb],
 * iters[nz][nb].  Runs on `threads` OpenMP threads when built with -fopenmp.
 */
void oracle_fixed_point_batch(int io_type, int nz, int nb, int n_sites,
                              const double *W, const double *ext,
                              double k, double n,
                              double tau_E, double tau_I, double dt,
                              int max_iter, double atol,
                              double r_soft, double r_hard,
                              int stop_at_first_failure, int threads,
                              double *R, int *status, int *iters)
{
    const int dim = 2 * n_sites;
    (void)threads;
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads > 0 ? threads : 1)
    for (int z = 0; z < nz; ++z) {
        double *scratch = (double *)malloc(sizeof(double) * dim);
        int failed = 0;
        for (int b = nb - 1; b >= 0; --b) {
            const size_t o = (size_t)z * nb + b;
            double *r = R + o * dim;
            memset(r, 0, sizeof(double) * dim);
            if (failed && stop_at_first_failure) {
                status[o] = -1;
                iters[o] = 0;
                continue;
            }
            status[o] = oracle_fixed_point(io_type, n_sites,
                                           W + (size_t)z * dim * dim,
                                           ext + (size_t)b * dim, k, n,
                                           r, scratch, tau_E, tau_I, dt,
                                           max_iter, atol, r_soft, r_hard,
                                           &iters[o]);
            if (status[o] != 0)
                failed = 1;
        }
        free(scratch);
    }
}
