/*
 * fruits_oracle.c -- CPU restatement of the FRUITS hot path (ISS + sieves).
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle for the CUDA
 * path in fruits_b200/.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it.  The product
 * never links or calls it.
 *
 * Every function restates one numba kernel of the reference (paths relative
 * to the reference checkout) with the same operation order, so that results
 * are bit-identical where the reference is deterministic:
 *   - no FMA contraction for the real semiring (separate statements in the
 *     reference => separate roundings); compile with -ffp-contract=off.
 *   - explicit fma() for the arctic semiring (numba fastmath contracts
 *     `tmp + el*Z` into one FMA on FMA-capable hosts).
 *
 * Parity pinned: oracle/gen_golden.py runs the real reference in the build
 * container and checks this file against it (and freezes tests/golden/).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define EXPORT __attribute__((visibility("default")))

/* ------------------------------------------------------------------ */
/* fruits/cache.py:8-13  _increments(X, k)                              */
/* X, out: [n, d, t] C-order.                                           */
EXPORT void fo_increments(const double *X, double *out, int64_t n, int64_t d,
                          int64_t t, int64_t k)
{
    int64_t rows = n * d;
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < rows; r++) {
        const double *x = X + r * t;
        double *o = out + r * t;
        for (int64_t i = 0; i < t; i++)
            o[i] = (i >= k) ? x[i] - x[i - k] : 0.0;
    }
}

/* fruits/cache.py:25-31 _L1_sum / :34-40 _L2_sum on dim 0 of X[n,d,t].  */
EXPORT void fo_lsum(const double *X, double *out, int64_t n, int64_t d,
                    int64_t t, int l2)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) {
        const double *x = X + i * d * t; /* dim 0 */
        double *o = out + i * t;
        double acc = 0.0;
        for (int64_t j = 0; j < t; j++) {
            double inc = (j >= 1) ? x[j] - x[j - 1] : 0.0;
            double v = l2 ? inc * inc : fabs(inc);
            acc = acc + v;
            o[j] = acc;
        }
    }
}

/* fruits/cache.py:16-22 _coquantile: count(X[i,:] <= q*X[i,-1]).        */
EXPORT void fo_coquantile(const double *X, int64_t *out, int64_t n, int64_t t,
                          double q)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) {
        const double *x = X + i * t;
        double thr = q * x[t - 1];
        int64_t c = 0;
        for (int64_t j = 0; j < t; j++)
            c += (x[j] <= thr);
        out[i] = c;
    }
}

/* ------------------------------------------------------------------ */
/* ISS kernels.  Z: [d, t] one series; word: [p, md] int32 exponents;   */
/* alpha: [p] float32; w: [t] lookup; result: [extended, t].            */
/* tmp, aux: scratch [t].                                                */

static void mul_letters(double *tmp, const double *Z, const int32_t *el,
                        int64_t md, int64_t t)
{
    /* fruits/iss/semiring.py:143-149 (and :110-116): one pass over the
     * array per occurrence, dims ascending. */
    for (int64_t l = 0; l < md; l++) {
        int32_t occ = el[l];
        const double *z = Z + l * t;
        if (occ > 0) {
            for (int32_t c = 0; c < occ; c++)
                for (int64_t j = 0; j < t; j++)
                    tmp[j] = tmp[j] * z[j];
        } else if (occ < 0) {
            for (int32_t c = 0; c < -occ; c++)
                for (int64_t j = 0; j < t; j++)
                    tmp[j] = tmp[j] / z[j];
        }
    }
}

static void cumsum_inplace(double *a, int64_t t)
{
    double acc = 0.0;
    for (int64_t j = 0; j < t; j++) {
        acc = acc + a[j];
        a[j] = acc;
    }
}

static void roll_zero(double *a, int64_t t)
{
    /* np.roll(tmp, 1); tmp[0] = 0 */
    for (int64_t j = t - 1; j >= 1; j--)
        a[j] = a[j - 1];
    if (t > 0)
        a[0] = 0.0;
}

/* fruits/iss/semiring.py:128-158 _total_weighted_reals_single */
static void reals_total_single(const double *Z, const int32_t *word,
                               const float *alpha, const double *w,
                               int64_t p, int64_t md, int64_t t,
                               int64_t extended, double *result, double *tmp)
{
    for (int64_t j = 0; j < t; j++)
        tmp[j] = 1.0;
    for (int64_t k = 0; k < p; k++) {
        double a = (double)alpha[k];
        mul_letters(tmp, Z, word + k * md, md, t);
        for (int64_t j = 0; j < t; j++)
            tmp[j] = tmp[j] * exp(w[j] * a);
        cumsum_inplace(tmp, t);
        if (p - k <= extended) {
            double *r = result + (extended - (p - k)) * t;
            for (int64_t j = 0; j < t; j++)
                r[j] = tmp[j] * exp(-w[j] * a);
        }
        if (k < p - 1) {
            roll_zero(tmp, t);
            for (int64_t j = 0; j < t; j++)
                tmp[j] = tmp[j] * exp(-w[j] * a);
        }
    }
}

/* fruits/iss/semiring.py:93-125 _reals_single */
static void reals_single(const double *Z, const int32_t *word,
                         const float *alpha, const double *w, int64_t p,
                         int64_t md, int64_t t, int64_t extended,
                         double *result, double *tmp)
{
    for (int64_t j = 0; j < t; j++)
        tmp[j] = 1.0;
    for (int64_t k = 0; k < p; k++) {
        if (k > 0)
            roll_zero(tmp, t);
        mul_letters(tmp, Z, word + k * md, md, t);
        if (k > 0) {
            double a = (double)alpha[k - 1];
            for (int64_t j = 0; j < t; j++)
                tmp[j] = tmp[j] * exp(-w[j] * a);
        }
        if (p - k <= extended) {
            double *r = result + (extended - (p - k)) * t;
            double acc = 0.0;
            for (int64_t j = 0; j < t; j++) {
                acc = acc + tmp[j];
                r[j] = acc;
            }
        }
        if (k < p - 1) {
            double a = (double)alpha[k];
            for (int64_t j = 0; j < t; j++)
                tmp[j] = tmp[j] * exp(w[j] * a);
            cumsum_inplace(tmp, t);
        }
    }
}

/* fruits/iss/semiring.py:314-338 _total_weighted_arctic_single */
static void arctic_total_single(const double *Z, const int32_t *word,
                                const float *alpha, const double *w,
                                int64_t p, int64_t md, int64_t t,
                                int64_t extended, double *result, double *tmp)
{
    for (int64_t j = 0; j < t; j++)
        tmp[j] = 0.0;
    for (int64_t k = 0; k < p; k++) {
        double a = (double)alpha[k];
        for (int64_t l = 0; l < md; l++) {
            double el = (double)word[k * md + l];
            const double *z = Z + l * t;
            for (int64_t j = 0; j < t; j++)
                tmp[j] = fma(el, z[j], tmp[j]);
        }
        for (int64_t j = 0; j < t; j++)
            tmp[j] = fma(w[j], a, tmp[j]);
        for (int64_t j = 1; j < t; j++)
            tmp[j] = (tmp[j - 1] > tmp[j]) ? tmp[j - 1] : tmp[j];
        if (p - k <= extended) {
            double *r = result + (extended - (p - k)) * t;
            for (int64_t j = 0; j < t; j++)
                r[j] = fma(-w[j], a, tmp[j]);
        }
        if (k < p - 1) {
            for (int64_t j = 0; j < t; j++)
                tmp[j] = fma(-w[j], a, tmp[j]);
        }
    }
}

/* fruits/iss/semiring.py:282-311 _arctic_single */
static void arctic_single(const double *Z, const int32_t *word,
                          const float *alpha, const double *w, int64_t p,
                          int64_t md, int64_t t, int64_t extended,
                          double *result, double *tmp)
{
    for (int64_t j = 0; j < t; j++)
        tmp[j] = 0.0;
    for (int64_t k = 0; k < p; k++) {
        for (int64_t l = 0; l < md; l++) {
            double el = (double)word[k * md + l];
            const double *z = Z + l * t;
            for (int64_t j = 0; j < t; j++)
                tmp[j] = fma(el, z[j], tmp[j]);
        }
        if (k > 0) {
            double a = (double)alpha[k - 1];
            for (int64_t j = 0; j < t; j++)
                tmp[j] = fma(-w[j], a, tmp[j]);
        }
        if (p - k <= extended) {
            double *r = result + (extended - (p - k)) * t;
            if (t > 0)
                r[0] = tmp[0];
            for (int64_t j = 1; j < t; j++)
                r[j] = (r[j - 1] > tmp[j]) ? r[j - 1] : tmp[j];
        }
        if (k < p - 1) {
            double a = (double)alpha[k];
            for (int64_t j = 0; j < t; j++)
                tmp[j] = fma(w[j], a, tmp[j]);
            for (int64_t j = 1; j < t; j++)
                tmp[j] = (tmp[j - 1] > tmp[j]) ? tmp[j - 1] : tmp[j];
        }
    }
}

/* fruits/iss/semiring.py:503-527 _total_weighted_bayesian_single */
static void bayesian_total_single(const double *Z, const int32_t *word,
                                  const float *alpha, const double *w,
                                  int64_t p, int64_t md, int64_t t,
                                  int64_t extended, double *result, double *tmp)
{
    for (int64_t j = 0; j < t; j++)
        tmp[j] = 1.0;
    for (int64_t k = 0; k < p; k++) {
        double a = (double)alpha[k];
        mul_letters(tmp, Z, word + k * md, md, t);
        for (int64_t j = 0; j < t; j++)
            tmp[j] = tmp[j] * exp(w[j] * a);
        for (int64_t j = 1; j < t; j++)
            tmp[j] = (tmp[j - 1] > tmp[j]) ? tmp[j - 1] : tmp[j];
        if (p - k <= extended) {
            double *r = result + (extended - (p - k)) * t;
            for (int64_t j = 0; j < t; j++)
                r[j] = tmp[j] * exp(-w[j] * a);
        }
        if (k < p - 1) {
            for (int64_t j = 0; j < t; j++)
                tmp[j] = tmp[j] * exp(-w[j] * a);
        }
    }
}

/* fruits/iss/semiring.py:466-495 _bayesian_single */
static void bayesian_single(const double *Z, const int32_t *word,
                            const float *alpha, const double *w, int64_t p,
                            int64_t md, int64_t t, int64_t extended,
                            double *result, double *tmp)
{
    for (int64_t j = 0; j < t; j++)
        tmp[j] = 1.0;
    for (int64_t k = 0; k < p; k++) {
        mul_letters(tmp, Z, word + k * md, md, t);
        if (k > 0) {
            double a = (double)alpha[k - 1];
            for (int64_t j = 0; j < t; j++)
                tmp[j] = tmp[j] * exp(-w[j] * a);
        }
        if (p - k <= extended) {
            double *r = result + (extended - (p - k)) * t;
            if (t > 0)
                r[0] = tmp[0];
            for (int64_t j = 1; j < t; j++)
                r[j] = (r[j - 1] > tmp[j]) ? r[j - 1] : tmp[j];
        }
        if (k < p - 1) {
            double a = (double)alpha[k];
            for (int64_t j = 0; j < t; j++)
                tmp[j] = tmp[j] * exp(w[j] * a);
            for (int64_t j = 1; j < t; j++)
                tmp[j] = (tmp[j - 1] > tmp[j]) ? tmp[j - 1] : tmp[j];
        }
    }
}

/* fruits/iss/semiring.py:167-201 Reals._iterated_sum_fast and :354-404
 * Arctic._iterated_sum_fast: loop (prange) over series.
 *   X       [n, d, t]
 *   word    [p, md] (md <= d)
 *   lookup  [n, t]
 *   result  [n, extended, t]
 *   semiring: 0 reals, 1 arctic, 2 bayesian (:530-566)
 */
EXPORT void fo_iterated_sums(const double *X, const int32_t *word,
                             const float *alpha, const double *lookup,
                             double *result, int64_t n, int64_t d, int64_t t,
                             int64_t p, int64_t md, int64_t extended,
                             int semiring, int total)
{
#pragma omp parallel
    {
        double *tmp = (double *)malloc(sizeof(double) * (size_t)(t > 0 ? t : 1));
#pragma omp for schedule(static)
        for (int64_t i = 0; i < n; i++) {
            const double *Z = X + i * d * t;
            const double *w = lookup + i * t;
            double *r = result + i * extended * t;
            memset(r, 0, sizeof(double) * (size_t)(extended * t));
            if (semiring == 0) {
                if (total)
                    reals_total_single(Z, word, alpha, w, p, md, t, extended, r, tmp);
                else
                    reals_single(Z, word, alpha, w, p, md, t, extended, r, tmp);
            } else if (semiring == 1) {
                if (total)
                    arctic_total_single(Z, word, alpha, w, p, md, t, extended, r, tmp);
                else
                    arctic_single(Z, word, alpha, w, p, md, t, extended, r, tmp);
            } else {
                if (total)
                    bayesian_total_single(Z, word, alpha, w, p, md, t, extended, r, tmp);
                else
                    bayesian_single(Z, word, alpha, w, p, md, t, extended, r, tmp);
            }
        }
        free(tmp);
    }
}

/* ------------------------------------------------------------------ */
/* Sieve backends.  X [n, t]; cuts [n, nc] int64 (sorted, cuts[:,0]==0);*/
/* q [nq] thresholds sorted; result [n, (nc-1)*(nq-1)].                 */
/* kind: 0 NPI (increment.py:121-129), 1 MPI (:152-163),                */
/*       2 MAX (segment.py:123-140), 3 MIN (:171-188),                  */
/*       4 XPI (increment.py:184-199), 5 LPI (:217-239),                */
/*       6 CUR (segment.py:246-258; X = second-order increments)        */
EXPORT void fo_segment_sieve(const double *X, const int64_t *cuts,
                             const double *q, double *result, int64_t n,
                             int64_t t, int64_t nc, int64_t nq, int kind)
{
    int64_t nf = (nc - 1) * (nq - 1);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) {
        const double *x = X + i * t;
        for (int64_t j = 0; j < nc - 1; j++) {
            int64_t lo = cuts[i * nc + j], hi = cuts[i * nc + j + 1];
            if (lo < 0) lo = 0;
            if (hi > t) hi = t;
            for (int64_t k = 0; k < nq - 1; k++) {
                double ql = q[k], qh = q[k + 1];
                int64_t cnt = 0, longest = 0, current = 0;
                double sum = 0.0, idxsum = 0.0, sq = 0.0;
                double mx = 0.0, mn = 0.0;
                int have = 0;
                for (int64_t s = lo; s < hi; s++) {
                    double v = x[s];
                    if (ql < v && v <= qh) {
                        cnt++;
                        sum = sum + v;
                        sq = sq + v * v;
                        idxsum = idxsum + (double)(s - lo);
                        if (!have) { mx = v; mn = v; have = 1; }
                        else { if (v > mx) mx = v; if (v < mn) mn = v; }
                        current++;
                        if (current > longest) longest = current;
                    } else {
                        current = 0;
                    }
                }
                double r;
                switch (kind) {
                case 0: r = (double)cnt; break;
                case 1: r = cnt ? sum / (double)cnt : 0.0; break;
                case 2: r = have ? mx : 0.0; break;
                case 3: r = have ? mn : 0.0; break;
                case 4: r = cnt ? idxsum / (double)cnt : 0.0; break;
                case 6: r = sq; break;
                default: r = (double)longest; break;
                }
                result[i * nf + j * (nq - 1) + k] = r;
            }
        }
    }
}

/* fruits/sieving/implicit.py:114-129 PPV._transform.
 * segments==0: count(X >= q_j)/t ; segments==1: count(q_{j-1} <= X < q_j)/t
 * segments|2: :169-190 CPV._transform -- 2 * #(rising edges of the indicator,
 * zero-padded increments) / (t rounded up to even) */
EXPORT void fo_ppv(const double *X, const double *q, double *result,
                   int64_t n, int64_t t, int64_t nq, int segments)
{
    const int cpv = (segments & 2) != 0;
    segments &= 1;
    int64_t nf = segments ? nq - 1 : nq;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) {
        const double *x = X + i * t;
        for (int64_t j = 0; j < nf; j++) {
            int64_t c = 0;
            if (cpv) {
                for (int64_t s = 1; s < t; s++) {
                    int a = segments ? (q[j] <= x[s - 1] && x[s - 1] < q[j + 1]) : (x[s - 1] >= q[j]);
                    int b = segments ? (q[j] <= x[s] && x[s] < q[j + 1]) : (x[s] >= q[j]);
                    c += ((double)b - (double)a == 1.0);
                }
                result[i * nf + j] = (double)(2 * c) / (double)(t + (t & 1));
                continue;
            }
            if (segments) {
                for (int64_t s = 0; s < t; s++)
                    c += (q[j] <= x[s] && x[s] < q[j + 1]);
            } else {
                for (int64_t s = 0; s < t; s++)
                    c += (x[s] >= q[j]);
            }
            result[i * nf + j] = (double)c / (double)t;
        }
    }
}

EXPORT int fo_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* bench.py: use all host cores even when the launcher exported OMP_NUM_THREADS=1
 * (torchrun does). */
EXPORT void fo_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
