/*
 * oracle/lpnce_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked or imported by the product path).
 *
 * CPU restatement, in double precision, of the reference's Lp-InfoNCE objective and of the
 * gradient autograd derives for it.  It follows
 *     /root/reference/losses.py:443-477   (LpSimCLRLoss.loss, p >= 1 branch)
 *     /root/reference/losses.py:506-510   (_logmeanexp, the non-"simclr compatibility" branch)
 * and the closed forms SURVEY.md section 4 (P1, P2) verified against the reference's autograd.
 *
 * Parity pin: the reference repository has no tests / golden vectors of its own, so this file is
 * pinned against OUTPUTS OF THE REFERENCE ITSELF, generated in the build container by
 * tests/golden/make_golden.py (imports /root/reference/losses.py) and committed as
 * tests/golden/lpnce_*.npz.  tests/test_oracle_vs_golden.py checks this file against them.
 *
 * Unlike the reference (which materialises B x M x d temporaries) this walks pairs with O(1)
 * extra memory, so it also serves as the checker at sizes where the reference formulation
 * does not fit in host RAM.
 *
 * Semantics (all inputs fp32 row-major with explicit leading dimensions):
 *   D_ik   = sum_c |z1[i,c] - z3[k,c]|^p          (pow=True: p-th power of the Lp norm)
 *   pos_i  = sum_c |z1[i,c] - z2[i,c]|^p
 *   if pow == 0:  D_ik, pos_i are replaced by their p-th roots (losses.py:452-454 skipped)
 *   include_pos=1 ("simclr_compatibility_mode"): lse_i = log( sum_k exp(-D_ik/tau) + exp(-pos_i/tau) )
 *   include_pos=0: lse_i = log( sum_k exp(-D_ik/tau) ) - log(M)
 *   loss_i = 2 * ( alpha * pos_i / tau + (1 - alpha) * lse_i )
 * Gradient of  L = sum_i gl[i] * loss_i   (gl == NULL means gl[i] = 1/B, i.e. L = mean loss):
 *   the softmax weights W_ik = exp(-D_ik/tau - lse'_i), w+_i = exp(-pos_i/tau - lse'_i)
 *   (lse' = lse without the -log M shift), and d|t|^p/dt = p*sign(t)*|t|^(p-1).
 */
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

static inline double abs_pow(double t, double p) {
    double a = fabs(t);
    if (p == 1.0) return a;
    if (p == 2.0) return a * a;
    if (p == 3.0) return a * a * a;
    return pow(a, p);
}

/* d/dt |t|^p ; exactly 0 at t == 0 (torch's norm backward masks zero entries). */
static inline double dabs_pow(double t, double p) {
    if (t == 0.0) return 0.0;
    double a = fabs(t), s = t > 0 ? 1.0 : -1.0;
    if (p == 1.0) return s;
    if (p == 2.0) return 2.0 * t;
    if (p == 3.0) return 3.0 * s * a * a;
    return p * s * pow(a, p - 1.0);
}

static double pair_dist(const float* a, const float* b, int d, double p) {
    double acc = 0.0;
    for (int c = 0; c < d; ++c) acc += abs_pow((double)a[c] - (double)b[c], p);
    return acc;
}

int lpnce_oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/*
 * Forward.  Outputs (any may be NULL): loss_i[B], lse[B] (the value the loss uses, i.e. including
 * the -log M shift when include_pos == 0), pos[B], scalars[3] = {mean loss, mean pos/tau, mean lse}.
 */
int lpnce_oracle_fwd(const float* z1, int ld1, const float* z2, int ld2, const float* z3, int ld3,
                     int B, int M, int d, double p, double tau, double alpha, int include_pos,
                     int use_pow, double* loss_i, double* lse, double* pos, double* scalars) {
    if (B < 0 || M < 0 || d < 0 || tau == 0.0 || p <= 0.0) return -1;
    double s_loss = 0.0, s_pos = 0.0, s_lse = 0.0;
    const double inv_p = 1.0 / p;
#pragma omp parallel for schedule(static) reduction(+ : s_loss, s_pos, s_lse)
    for (int i = 0; i < B; ++i) {
        const float* a = z1 + (size_t)i * ld1;
        double pi = pair_dist(a, z2 + (size_t)i * ld2, d, p);
        if (!use_pow) pi = pow(pi, inv_p);
        /* streaming log-sum-exp with a running maximum */
        double m = include_pos ? -pi / tau : -INFINITY, s = include_pos ? 1.0 : 0.0;
        for (int k = 0; k < M; ++k) {
            double D = pair_dist(a, z3 + (size_t)k * ld3, d, p);
            if (!use_pow) D = pow(D, inv_p);
            double x = -D / tau;
            if (x > m) { s = s * exp(m - x) + 1.0; m = x; }
            else s += exp(x - m);
        }
        double l = m + log(s);
        if (!include_pos) l -= log((double)M);
        double li = 2.0 * (alpha * pi / tau + (1.0 - alpha) * l);
        if (loss_i) loss_i[i] = li;
        if (lse) lse[i] = l;
        if (pos) pos[i] = pi;
        s_loss += li; s_pos += pi / tau; s_lse += l;
    }
    if (scalars && B > 0) { scalars[0] = s_loss / B; scalars[1] = s_pos / B; scalars[2] = s_lse / B; }
    return 0;
}

/*
 * Backward of L = sum_i gl[i]*loss_i (gl NULL -> 1/B each).  Needs lse[] and pos[] from the forward.
 * g1[B*d], g2[B*d], g3[M*d] are dense (ld = d) double outputs; any may be NULL.
 * g1 holds only the "row" contribution (z1 as anchor); if z3 aliases z1 (roll) the caller adds g3.
 */
int lpnce_oracle_bwd(const float* z1, int ld1, const float* z2, int ld2, const float* z3, int ld3,
                     int B, int M, int d, double p, double tau, double alpha, int include_pos,
                     int use_pow, const double* lse, const double* pos, const double* gl,
                     double* g1, double* g2, double* g3) {
    if (B < 0 || M < 0 || d < 0 || tau == 0.0 || p <= 0.0) return -1;
    const double shift = include_pos ? 0.0 : log((double)M);
    const double inv_p = 1.0 / p;
    /* pass 1: rows -> g1, g2 */
#pragma omp parallel for schedule(static)
    for (int i = 0; i < B; ++i) {
        const float* a = z1 + (size_t)i * ld1;
        const float* b = z2 + (size_t)i * ld2;
        const double gi = gl ? gl[i] : 1.0 / B;
        const double lsei = lse[i] + shift; /* un-shifted log-sum-exp */
        /* positive pair: dloss/dpos = 2*(alpha/tau - (1-alpha) * w+ / tau) */
        double wpos = include_pos ? exp(-pos[i] / tau - lsei) : 0.0;
        double cpos = 2.0 * gi * (alpha - (1.0 - alpha) * wpos) / tau;
        /* chain through the optional p-th root: d(S^(1/p))/dS = S^(1/p-1)/p, 0 at S == 0 */
        if (!use_pow) {
            double S = pair_dist(a, b, d, p);
            cpos *= (S > 0.0) ? pow(S, inv_p - 1.0) * inv_p : 0.0;
        }
        for (int c = 0; c < d; ++c) {
            double g = cpos * dabs_pow((double)a[c] - (double)b[c], p);
            if (g1) g1[(size_t)i * d + c] = g;
            if (g2) g2[(size_t)i * d + c] = -g;
        }
        if (!g1) continue;
        for (int k = 0; k < M; ++k) {
            const float* n = z3 + (size_t)k * ld3;
            double S = pair_dist(a, n, d, p);
            double D = use_pow ? S : pow(S, inv_p);
            double w = exp(-D / tau - lsei);
            double cneg = -2.0 * gi * (1.0 - alpha) * w / tau;
            if (!use_pow) cneg *= (S > 0.0) ? pow(S, inv_p - 1.0) * inv_p : 0.0;
            for (int c = 0; c < d; ++c)
                g1[(size_t)i * d + c] += cneg * dabs_pow((double)a[c] - (double)n[c], p);
        }
    }
    /* pass 2: columns -> g3 (recompute; parallel over k so no atomics are needed) */
    if (g3) {
#pragma omp parallel for schedule(static)
        for (int k = 0; k < M; ++k) {
            const float* n = z3 + (size_t)k * ld3;
            double* out = g3 + (size_t)k * d;
            for (int c = 0; c < d; ++c) out[c] = 0.0;
            for (int i = 0; i < B; ++i) {
                const float* a = z1 + (size_t)i * ld1;
                const double gi = gl ? gl[i] : 1.0 / B;
                double S = pair_dist(a, n, d, p);
                double D = use_pow ? S : pow(S, inv_p);
                double w = exp(-D / tau - (lse[i] + shift));
                double cneg = 2.0 * gi * (1.0 - alpha) * w / tau; /* sign flips: d/dz3 = -d/dz1 */
                if (!use_pow) cneg *= (S > 0.0) ? pow(S, inv_p - 1.0) * inv_p : 0.0;
                for (int c = 0; c < d; ++c)
                    out[c] += cneg * dabs_pow((double)a[c] - (double)n[c], p);
            }
        }
    }
    return 0;
}
