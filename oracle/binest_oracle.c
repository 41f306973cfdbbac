/*
 * binest_oracle.c — CPU restatement of the nested-sampling hot path of
 * ssmit1986/BayesianInference.  TEST INFRASTRUCTURE ONLY.
 *
 *   - Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 *     legs may load this library.  The product (bayesianinference_b200/) never does.
 *   - PARITY UNPINNED BY THE REFERENCE: the reference is pure Wolfram Language, ships no
 *     tests, golden vectors or fixtures, and cannot be executed here (no Wolfram Engine).
 *     Every pin is constructed (closed forms, quadrature, algebraic invariants, Random123
 *     known-answer vectors) — see tests/test_oracle_*.py and DESIGN.md.
 *
 * Citations: BS = BayesianInference/Kernel/BayesianStatistics.wl, BU = BayesianUtilities.wl,
 * GP = BayesianGaussianProcess.wl  (paths under /root/reference).
 *
 * The adaptive-Metropolis walk lives in the closed-source Wolfram kernel
 * (Statistics`MCMC`BuildMarkovChain, called at BS:720-727); it is restated here from
 * Haario, Saksman & Tamminen (2001) with the start state the reference passes (t = 10, supplied
 * mean/covariance, BS:715-727).
 *
 * All arithmetic is fp64 in the reference's summation order (sequential over the data, BS:492,
 * BS:581); the *_q entry points repeat the likelihood sums in __float128 to grade both the fp64
 * restatement and the GPU result.
 */
#include <math.h>
#include <quadmath.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))
#define ORC_MAX_RETRY 20
#define ORC_MAXD 16

enum { OP_GAUSSIAN_IID = 1, OP_POLYREG = 2, OP_LOGISTIC = 3, OP_GBM = 4, OP_GP_SE = 5 };
enum { PRIOR_UNIFORM = 1, PRIOR_SCALE = 2, PRIOR_NORMAL_TRUNC = 3 };

static const double LOG_2PI = 1.8378770664093454835606594728112;
static const double HALF_LOG_2PI = 0.91893853320467274178032973640562;

/* ------------------------------------------------------------------------------------------ */
/* Philox4x32-10 (Salmon et al., SC'11; Random123).  Shared counter layout with the CUDA path. */
/* ------------------------------------------------------------------------------------------ */
static inline void philox_round(uint32_t c[4], uint32_t k[2]) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0];
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1];
    const uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

ORC_API void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
    uint32_t k[2] = {key[0], key[1]};
    for (int r = 0; r < 10; ++r) {
        if (r) { k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u; }
        philox_round(c, k);
    }
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}

/* 53-bit uniform in the open interval (0,1) from two 32-bit words */
static inline double u53(uint32_t hi, uint32_t lo) {
    const uint64_t m = ((uint64_t)(hi >> 5) << 26) | (uint64_t)(lo >> 6);
    return ((double)m + 0.5) * (1.0 / 9007199254740992.0);
}

/* stream tags (counter word 3 = tag<<24 | run_id) */
enum { TAG_PRIOR = 1, TAG_START = 2, TAG_NORMAL = 3, TAG_ACCEPT = 4, TAG_EV_DEAD = 5, TAG_EV_LIVE = 6 };

/* two uniforms for (block c0, c1, c2, tag, run) */
ORC_API void orc_uniform2(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t tag,
                          uint32_t run_id, double out[2]) {
    uint32_t ctr[4] = {c0, c1, c2, (tag << 24) | (run_id & 0xFFFFFFu)};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t r[4];
    orc_philox4x32_10(ctr, key, r);
    out[0] = u53(r[0], r[1]);
    out[1] = u53(r[2], r[3]);
}

/* two standard normals (Box–Muller) for the same addressing */
ORC_API void orc_normal2(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t tag,
                         uint32_t run_id, double out[2]) {
    double u[2];
    orc_uniform2(seed, c0, c1, c2, tag, run_id, u);
    const double r = sqrt(-2.0 * log(u[0]));
    const double a = 6.283185307179586476925286766559 * u[1];
    out[0] = r * cos(a);
    out[1] = r * sin(a);
}

/* ------------------------------------------------------------------------------------------ */
/* Log-space helpers — BU:318-356                                                             */
/* ------------------------------------------------------------------------------------------ */
/* BU:337-343  logSubtract[logy, logx] = logy + Log[1 - Exp[logx - logy]] */
ORC_API double orc_logsubtract(double logy, double logx) { return logy + log(1.0 - exp(logx - logy)); }

/* BU:345-356  logAdd */
ORC_API double orc_logadd(double logy, double logx) {
    const double mx = logx > logy ? logx : logy, mn = logx > logy ? logy : logx;
    return mx + log(1.0 + exp(mn - mx));
}

/* BU:318-335  logSumExp: max + Log[Total[Exp[list - max]]] after dropping -Infinity entries */
ORC_API double orc_logsumexp(const double *v, int64_t n) {
    double mx = -INFINITY;
    int64_t cnt = 0;
    for (int64_t i = 0; i < n; ++i)
        if (isfinite(v[i])) { if (v[i] > mx) mx = v[i]; ++cnt; }
    if (!cnt) return -INFINITY;
    double s = 0.0;
    for (int64_t i = 0; i < n; ++i)
        if (isfinite(v[i])) s += exp(v[i] - mx);
    return mx + log(s);
}

/* ------------------------------------------------------------------------------------------ */
/* X values and trapezoid weights — BS:747-831                                                */
/* ------------------------------------------------------------------------------------------ */
/* BS:785-799 calculateXValues["Log"][n, nDeleted]:
 *   deleted k = 1..nDeleted: -k/n ;  live i = n..1: Log[i] - Log[n+1] - nDeleted/n            */
ORC_API void orc_xvalues_log(int64_t n, int64_t n_deleted, double *out) {
    for (int64_t k = 1; k <= n_deleted; ++k) out[k - 1] = -((double)k / (double)n);
    for (int64_t j = 0; j < n; ++j) {
        const double i = (double)(n - j);
        out[n_deleted + j] = (log(i) - log((double)n + 1.0)) - ((double)n_deleted / (double)n);
    }
}

/* Generalisation to per-sample pool sizes (batched replacement, SURVEY §7 hard part 1):
 *   deleted k: logX_k = -sum_{i<=k} 1/pool_i ; live as BS:791-797 with -nDeleted/n replaced by
 *   logX of the last deleted point.  With pool_i == n this is BS:785-799 up to rounding.       */
ORC_API void orc_xvalues_log_pool(int64_t n, int64_t n_deleted, const int64_t *pool, double *out) {
    double acc = 0.0;
    for (int64_t k = 0; k < n_deleted; ++k) { acc -= 1.0 / (double)pool[k]; out[k] = acc; }
    for (int64_t j = 0; j < n; ++j) {
        const double i = (double)(n - j);
        out[n_deleted + j] = (log(i) - log((double)n + 1.0)) + acc;
    }
}

/* BS:756-771 trapezoidWeigths["Log"]:
 *   lw_1 = log(1/2) + logSubtract[logSubtract[Log 2, lx_1], lx_2]
 *   lw_k = log(1/2) + logSubtract[lx_{k-1}, lx_{k+1}]
 *   lw_M = log(1/2) + logAdd[lx_{M-1}, lx_M]                                                  */
ORC_API void orc_trapezoid_log(const double *lx, int64_t M, double *lw) {
    const double lh = log(0.5);
    if (M < 2) { if (M == 1) lw[0] = 0.0; return; }
    for (int64_t k = 0; k < M - 1; ++k) {
        const double left = (k == 0) ? orc_logsubtract(log(2.0), lx[0]) : lx[k - 1];
        lw[k] = lh + orc_logsubtract(left, lx[k + 1]);
    }
    lw[M - 1] = lh + orc_logadd(lx[M - 2], lx[M - 1]);
}

/* BS:801-810 calculateEntropy: Sum_k Exp[w_k - logZ] * L_k - logZ, with -Infinity logL -> 0 */
ORC_API double orc_entropy(const double *crude_logw, const double *logL, int64_t M, double logZ) {
    double s = 0.0;
    for (int64_t k = 0; k < M; ++k) {
        const double L = isfinite(logL[k]) ? logL[k] : 0.0;
        s += exp(crude_logw[k] - logZ) * L;
    }
    return s - logZ;
}

/* ------------------------------------------------------------------------------------------ */
/* Priors — BS:25-64 (ignorancePrior), BS:327-427 (box constraints, logPDFFunction)           */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    int d;
    int kind[ORC_MAXD];
    double lo[ORC_MAXD], hi[ORC_MAXD];
    double p0[ORC_MAXD], p1[ORC_MAXD]; /* NORMAL_TRUNC: mean, sd */
    double logzero;
} orc_prior;

static double norm_cdf(double z) { return 0.5 * erfc(-z * 0.70710678118654752440084436210485); }

/* BS:327-336: open box lo < theta < hi */
static int in_box(const orc_prior *pr, const double *th) {
    for (int j = 0; j < pr->d; ++j)
        if (!(th[j] > pr->lo[j] && th[j] < pr->hi[j])) return 0;
    return 1;
}

static double logprior_dim(const orc_prior *pr, int j, double t) {
    switch (pr->kind[j]) {
    case PRIOR_UNIFORM: /* BS:37-39 UniformDistribution[{lo,hi}] */
        return -log(pr->hi[j] - pr->lo[j]);
    case PRIOR_SCALE: /* BS:42-48 ProbabilityDistribution[1/x, {x,lo,hi}, Method->"Normalize"] */
        return -log(t) - log(log(pr->hi[j] / pr->lo[j]));
    case PRIOR_NORMAL_TRUNC: { /* BS:51-59 TruncatedDistribution[{lo,hi}, NormalDistribution[m,s]] */
        const double m = pr->p0[j], s = pr->p1[j];
        const double z = (t - m) / s;
        const double mass = norm_cdf((pr->hi[j] - m) / s) - norm_cdf((pr->lo[j] - m) / s);
        return -0.5 * z * z - log(s) - HALF_LOG_2PI - log(mass);
    }
    }
    return pr->logzero;
}

/* BS:410-426: If[constraints[param], logPDF[param], logzero]; list prior -> ProductDistribution BS:33 */
ORC_API double orc_logprior(const orc_prior *pr, const double *th) {
    if (!in_box(pr, th)) return pr->logzero;
    double s = 0.0;
    for (int j = 0; j < pr->d; ++j) s += logprior_dim(pr, j, th[j]);
    return isfinite(s) ? s : pr->logzero;
}

/* BS:1055-1068 generateStartingPoints: n i.i.d. draws from the prior (inverse CDF; rejection
 * from the parent normal for the truncated case).  Point i uses Philox (c1 = dim, c2 = i).   */
ORC_API void orc_sample_prior(const orc_prior *pr, int64_t n, uint64_t seed, uint32_t run_id, double *out) {
    for (int64_t i = 0; i < n; ++i)
        for (int j = 0; j < pr->d; ++j) {
            double u[2];
            double v;
            orc_uniform2(seed, 0, (uint32_t)j, (uint32_t)i, TAG_PRIOR, run_id, u);
            const double lo = pr->lo[j], hi = pr->hi[j];
            if (pr->kind[j] == PRIOR_UNIFORM) v = lo + u[0] * (hi - lo);
            else if (pr->kind[j] == PRIOR_SCALE) v = lo * exp(u[0] * log(hi / lo));
            else {
                uint32_t blk = 1;
                v = NAN;
                for (;;) {
                    double z[2];
                    orc_normal2(seed, blk, (uint32_t)j, (uint32_t)i, TAG_PRIOR, run_id, z);
                    const double a = pr->p0[j] + pr->p1[j] * z[0], b = pr->p0[j] + pr->p1[j] * z[1];
                    if (a > lo && a < hi) { v = a; break; }
                    if (b > lo && b < hi) { v = b; break; }
                    ++blk;
                }
            }
            out[i * pr->d + j] = v;
        }
}

/* ------------------------------------------------------------------------------------------ */
/* Likelihood operators — BS:429-505 (i.i.d.), BS:517-595 (regression), GP:27-61,130-199      */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    int op;
    int d;            /* number of parameters */
    int64_t n;        /* rows */
    int n_in, n_out;  /* columns */
    const double *in; /* n x n_in row-major (iid: the data itself) */
    const double *out;/* n x n_out */
    int iparam[4];    /* POLYREG: degree ; LOGISTIC: n_classes ; GP: input dim */
    double logzero;
    /* derived at create time */
    double *aux;      /* GBM: a_i, b_i, const ; LOGISTIC: labels as int */
    double cst;       /* parameter-independent additive constant (GBM) */
} orc_problem;

/* operator-level constraint from DistributionParameterAssumptions (BS:439, 523): sigma > 0 */
static int op_constraints_ok(const orc_problem *p, const double *th) {
    switch (p->op) {
    case OP_GAUSSIAN_IID: return th[1] > 0.0;
    case OP_POLYREG: return th[p->iparam[0] + 1] > 0.0;
    case OP_GBM: return th[1] > 0.0;
    case OP_GP_SE: return th[0] > 0.0 && th[1] > 0.0 && th[2] > 0.0;
    default: return 1;
    }
}

/* ORC_FAST (the second build of this file, libbinest_oracle_fast.so: -O3 -march=... -ffast-math, used ONLY by the
 * timed CPU legs of bench.py): the sums may be vectorised and re-associated.  The parity build keeps the
 * reference's sequential order (BS:492, 581). */
#ifdef ORC_FAST
#define ORC_SIMD _Pragma("omp simd reduction(+ : s)")
#else
#define ORC_SIMD
#endif
#define DEFINE_SUM_OPS(NAME, T, LOGF, EXPF, SQRTF)                                                   \
    static T NAME(const orc_problem *p, const double *th) {                                         \
        T s = 0;                                                                                    \
        const int64_t n = p->n;                                                                     \
        switch (p->op) {                                                                            \
        case OP_GAUSSIAN_IID: { /* -(x-mu)^2/(2 s^2) - log s - 1/2 log 2pi, summed as BS:492 */     \
            const T mu = th[0], sg = th[1];                                                         \
            const T c = -LOGF(sg) - (T)HALF_LOG_2PI, h = 1 / (2 * sg * sg);                         \
            ORC_SIMD                                                                                \
            for (int64_t i = 0; i < n; ++i) { const T r = (T)p->in[i] - mu; s += c - r * r * h; }   \
            return s;                                                                               \
        }                                                                                           \
        case OP_POLYREG: { /* NormalDistribution[Sum_j c_j x^j, sigma], BS:581 */                   \
            const int deg = p->iparam[0];                                                           \
            const T sg = th[deg + 1];                                                               \
            const T c = -LOGF(sg) - (T)HALF_LOG_2PI, h = 1 / (2 * sg * sg);                         \
            ORC_SIMD                                                                                \
            for (int64_t i = 0; i < n; ++i) {                                                       \
                const T x = p->in[i];                                                               \
                T t = th[deg];                                                                      \
                for (int j = deg - 1; j >= 0; --j) t = t * x + (T)th[j];                            \
                const T r = (T)p->out[i] - t;                                                       \
                s += c - r * r * h;                                                                 \
            }                                                                                       \
            return s;                                                                               \
        }                                                                                           \
        case OP_LOGISTIC: { /* categorical softmax, reference class K: z_K = 0 */       \
            const int F = p->n_in, K = p->iparam[1];                                                \
            ORC_SIMD                                                                                \
            for (int64_t i = 0; i < n; ++i) {                                                       \
                T z[ORC_MAXD];                                                                      \
                T mx = 0;                                                                           \
                const int y = (int)p->out[i];                                                       \
                for (int k = 0; k < K - 1; ++k) {                                                   \
                    const double *w = th + k * (F + 1);                                             \
                    T a = w[F];                                                                     \
                    for (int f = 0; f < F; ++f) a += (T)w[f] * (T)p->in[i * F + f];                 \
                    z[k] = a;                                                                       \
                    if (a > mx) mx = a;                                                             \
                }                                                                                   \
                z[K - 1] = 0;                                                                       \
                T e = 0;                                                                            \
                for (int k = 0; k < K; ++k) e += EXPF(z[k] - mx);                                   \
                s += z[y] - (mx + LOGF(e));                                                         \
            }                                                                                       \
            return s;                                                                               \
        }                                                                                           \
        case OP_GBM: { /* GeometricBrownianMotionProcess[mu, sigma, x0] on (t_i, x_i), SURVEY 8a */ \
            const T mu = th[0], sg = th[1];                                                         \
            const T m = mu - sg * sg / 2;                                                           \
            ORC_SIMD                                                                                \
            for (int64_t i = 1; i < n; ++i) {                                                       \
                const T dt = (T)p->in[i] - (T)p->in[i - 1];                                         \
                const T r = LOGF((T)p->out[i] / (T)p->out[i - 1]);                                  \
                const T dv = r - m * dt;                                                            \
                s += -LOGF((T)p->out[i]) - LOGF(2 * (T)M_PIq * sg * sg * dt) / 2                    \
                     - dv * dv / (2 * sg * sg * dt);                                                \
            }                                                                                       \
            return s;                                                                               \
        }                                                                                           \
        }                                                                                           \
        return 0;                                                                                   \
    }

DEFINE_SUM_OPS(sum_ops_d, double, log, exp, sqrt)
DEFINE_SUM_OPS(sum_ops_q, __float128, logq, expq, sqrtq)

/* GP marginal likelihood, squared-exponential kernel + nugget:
 *   K_ij = sf^2 exp(-|x_i-x_j|^2/(2 l^2)) + delta_ij sn^2      (GP:29-43 with the SE kernel)
 *   LU with partial pivoting, logdet = Sum log|U_ii|            (GP:126-141)
 *   -1/2 (N log 2pi + logdet + r.K^-1 r), clipped to +-|logzero| (GP:181-199);
 *   singular -> logzero (GP:131-135).  theta = (sf, l, sn).                                    */
static void gp_fill_ld(const orc_problem *p, const double *th, long double *K) {
    const int64_t n = p->n;
    const int D = p->n_in;
    const long double sf2 = (long double)th[0] * th[0], il2 = 1.0L / (2.0L * th[1] * th[1]),
                      sn2 = (long double)th[2] * th[2];
    for (int64_t i = 0; i < n; ++i)
        for (int64_t j = 0; j <= i; ++j) {
            long double d2 = 0;
            for (int k = 0; k < D; ++k) {
                const long double df = (long double)p->in[i * D + k] - p->in[j * D + k];
                d2 += df * df;
            }
            long double v = sf2 * expl(-d2 * il2);
            if (i == j) v += sn2;
            K[i * n + j] = K[j * n + i] = v;
        }
}

static double gp_loglike_lu(const orc_problem *p, const double *th) {
    const int64_t n = p->n;
    const int D = p->n_in;
    double *A = (double *)malloc(sizeof(double) * n * n);
    double *r = (double *)malloc(sizeof(double) * n);
    const double sf2 = th[0] * th[0], il2 = 1.0 / (2.0 * th[1] * th[1]), sn2 = th[2] * th[2];
    for (int64_t i = 0; i < n; ++i)
        for (int64_t j = 0; j <= i; ++j) {
            double d2 = 0;
            for (int k = 0; k < D; ++k) { const double df = p->in[i * D + k] - p->in[j * D + k]; d2 += df * df; }
            double v = sf2 * exp(-d2 * il2);
            if (i == j) v += sn2;
            A[i * n + j] = A[j * n + i] = v;
        }
    for (int64_t i = 0; i < n; ++i) r[i] = p->out[i];
    double logdet = 0.0;
    int singular = 0;
    for (int64_t k = 0; k < n && !singular; ++k) {
        int64_t piv = k;
        double mx = fabs(A[k * n + k]);
        for (int64_t i = k + 1; i < n; ++i)
            if (fabs(A[i * n + k]) > mx) { mx = fabs(A[i * n + k]); piv = i; }
        if (!(mx > 0.0) || !isfinite(mx)) { singular = 1; break; }
        if (piv != k) {
            for (int64_t j = 0; j < n; ++j) { const double t = A[k * n + j]; A[k * n + j] = A[piv * n + j]; A[piv * n + j] = t; }
            const double t = r[k]; r[k] = r[piv]; r[piv] = t;
        }
        logdet += log(fabs(A[k * n + k]));
        const double inv = 1.0 / A[k * n + k];
        for (int64_t i = k + 1; i < n; ++i) {
            const double f = A[i * n + k] * inv;
            if (f != 0.0) {
                for (int64_t j = k + 1; j < n; ++j) A[i * n + j] -= f * A[k * n + j];
                r[i] -= f * r[k];
            }
        }
    }
    double res;
    if (singular) res = p->logzero;
    else {
        /* back substitution U z = r', then quad = y . z */
        for (int64_t i = n - 1; i >= 0; --i) {
            double s = r[i];
            for (int64_t j = i + 1; j < n; ++j) s -= A[i * n + j] * r[j];
            r[i] = s / A[i * n + i];
        }
        double quad = 0.0;
        for (int64_t i = 0; i < n; ++i) quad += p->out[i] * r[i];
        res = -0.5 * ((double)n * LOG_2PI + logdet + quad);
        const double lim = fabs(p->logzero);
        if (res > lim) res = lim;
        if (res < -lim) res = -lim;
        if (!isfinite(res)) res = p->logzero;
    }
    free(A); free(r);
    return res;
}

/* same quantity by long-double Cholesky (precision reference for the GPU Cholesky path) */
static long double gp_loglike_chol_ld(const orc_problem *p, const double *th, int *ok) {
    const int64_t n = p->n;
    long double *A = (long double *)malloc(sizeof(long double) * n * n);
    long double *z = (long double *)malloc(sizeof(long double) * n);
    gp_fill_ld(p, th, A);
    long double logdet = 0;
    *ok = 1;
    for (int64_t j = 0; j < n; ++j) {
        long double s = A[j * n + j];
        for (int64_t k = 0; k < j; ++k) s -= A[j * n + k] * A[j * n + k];
        if (!(s > 0)) { *ok = 0; break; }
        const long double l = sqrtl(s);
        A[j * n + j] = l;
        logdet += 2 * logl(l);
        for (int64_t i = j + 1; i < n; ++i) {
            long double t = A[i * n + j];
            for (int64_t k = 0; k < j; ++k) t -= A[i * n + k] * A[j * n + k];
            A[i * n + j] = t / l;
        }
    }
    long double res = 0;
    if (*ok) {
        long double quad = 0;
        for (int64_t i = 0; i < n; ++i) {
            long double s = p->out[i];
            for (int64_t k = 0; k < i; ++k) s -= A[i * n + k] * z[k];
            z[i] = s / A[i * n + i];
            quad += z[i] * z[i];
        }
        res = -0.5L * ((long double)n * (long double)LOG_2PI + logdet + quad);
    }
    free(A); free(z);
    return res;
}

/* predictFromGaussianProcessInternal GP:395-420 with compiledKandKappa GP:92-116, squared-exponential kernel,
 * zero mean function:  invCov = LinearSolve[K] (LU, GP:132);  k_q = (k(x_i, x*_q))_i;  kappa = k(x*,x*) + nugget;
 *   mean_q = (invCov[y]) . k_q        sd_q = Sqrt[kappa - k_q . invCov[k_q]]
 * One LU factorisation (partial pivoting, as LinearSolve does for a dense real matrix), then one solve per
 * right-hand side.  use_ld != 0: the same quantities by long-double Cholesky (precision reference).
 * Returns 0 when the matrix is singular / not positive definite (the reference Throws, GP:131-135). */
ORC_API int orc_gp_predict(const orc_problem *p, const double *th, const double *xs, int64_t Q, int use_ld,
                           double *mean, double *sd) {
    const int64_t n = p->n;
    const int D = p->n_in;
    const long double sf2 = (long double)th[0] * th[0], il2 = 1.0L / (2.0L * th[1] * th[1]),
                      sn2 = (long double)th[2] * th[2];
    int ok = 1;
    if (use_ld) {
        long double *A = (long double *)malloc(sizeof(long double) * n * n);
        long double *z = (long double *)malloc(sizeof(long double) * n);
        long double *v = (long double *)malloc(sizeof(long double) * n);
        gp_fill_ld(p, th, A);
        for (int64_t j = 0; j < n && ok; ++j) {
            long double s = A[j * n + j];
            for (int64_t k = 0; k < j; ++k) s -= A[j * n + k] * A[j * n + k];
            if (!(s > 0)) { ok = 0; break; }
            const long double l = sqrtl(s);
            A[j * n + j] = l;
            for (int64_t i = j + 1; i < n; ++i) {
                long double t = A[i * n + j];
                for (int64_t k = 0; k < j; ++k) t -= A[i * n + k] * A[j * n + k];
                A[i * n + j] = t / l;
            }
        }
        if (ok) {
            for (int64_t i = 0; i < n; ++i) {   /* z = L^-1 y */
                long double s = p->out[i];
                for (int64_t k = 0; k < i; ++k) s -= A[i * n + k] * z[k];
                z[i] = s / A[i * n + i];
            }
            for (int64_t q = 0; q < Q; ++q) {
                long double m = 0, vv = 0;
                for (int64_t i = 0; i < n; ++i) {   /* v = L^-1 k_q */
                    long double d2 = 0;
                    for (int k = 0; k < D; ++k) { const long double df = (long double)p->in[i * D + k] - xs[q * D + k]; d2 += df * df; }
                    long double s = sf2 * expl(-d2 * il2);
                    for (int64_t k = 0; k < i; ++k) s -= A[i * n + k] * v[k];
                    v[i] = s / A[i * n + i];
                    m += v[i] * z[i];
                    vv += v[i] * v[i];
                }
                mean[q] = (double)m;
                sd[q] = (double)sqrtl(fmaxl(sf2 + sn2 - vv, 0.0L));
            }
        }
        free(A); free(z); free(v);
        return ok;
    }
    double *A = (double *)malloc(sizeof(double) * n * n);
    int64_t *perm = (int64_t *)malloc(sizeof(int64_t) * n);
    double *a = (double *)malloc(sizeof(double) * n), *b = (double *)malloc(sizeof(double) * n),
           *kq = (double *)malloc(sizeof(double) * n);
    for (int64_t i = 0; i < n; ++i)
        for (int64_t j = 0; j <= i; ++j) {
            double d2 = 0;
            for (int k = 0; k < D; ++k) { const double df = p->in[i * D + k] - p->in[j * D + k]; d2 += df * df; }
            double v = (double)sf2 * exp(-d2 * (double)il2);
            if (i == j) v += (double)sn2;
            A[i * n + j] = A[j * n + i] = v;
        }
    for (int64_t i = 0; i < n; ++i) perm[i] = i;
    for (int64_t k = 0; k < n && ok; ++k) {   /* P A = L U in place */
        int64_t piv = k;
        double mx = fabs(A[k * n + k]);
        for (int64_t i = k + 1; i < n; ++i)
            if (fabs(A[i * n + k]) > mx) { mx = fabs(A[i * n + k]); piv = i; }
        if (!(mx > 0.0) || !isfinite(mx)) { ok = 0; break; }
        if (piv != k) {
            for (int64_t j = 0; j < n; ++j) { const double t = A[k * n + j]; A[k * n + j] = A[piv * n + j]; A[piv * n + j] = t; }
            const int64_t t = perm[k]; perm[k] = perm[piv]; perm[piv] = t;
        }
        const double inv = 1.0 / A[k * n + k];
        for (int64_t i = k + 1; i < n; ++i) {
            const double f = A[i * n + k] * inv;
            A[i * n + k] = f;
            if (f != 0.0)
                for (int64_t j = k + 1; j < n; ++j) A[i * n + j] -= f * A[k * n + j];
        }
    }
    if (ok) {
#define ORC_LU_SOLVE(rhs, out)                                                      \
        for (int64_t i = 0; i < n; ++i) {                                           \
            double s_ = (rhs)[perm[i]];                                             \
            for (int64_t j = 0; j < i; ++j) s_ -= A[i * n + j] * (out)[j];          \
            (out)[i] = s_;                                                          \
        }                                                                           \
        for (int64_t i = n - 1; i >= 0; --i) {                                      \
            double s_ = (out)[i];                                                   \
            for (int64_t j = i + 1; j < n; ++j) s_ -= A[i * n + j] * (out)[j];      \
            (out)[i] = s_ / A[i * n + i];                                           \
        }
        ORC_LU_SOLVE(p->out, a)   /* a = invCov[y] */
        for (int64_t q = 0; q < Q; ++q) {
            for (int64_t i = 0; i < n; ++i) {
                double d2 = 0;
                for (int k = 0; k < D; ++k) { const double df = p->in[i * D + k] - xs[q * D + k]; d2 += df * df; }
                kq[i] = (double)sf2 * exp(-d2 * (double)il2);
            }
            ORC_LU_SOLVE(kq, b)
            double m = 0.0, kk = 0.0;
            for (int64_t i = 0; i < n; ++i) { m += a[i] * kq[i]; kk += kq[i] * b[i]; }
            mean[q] = m;
            const double var = ((double)sf2 + (double)sn2) - kk;
            sd[q] = var > 0.0 ? sqrt(var) : 0.0;
        }
#undef ORC_LU_SOLVE
    }
    free(A); free(perm); free(a); free(b); free(kq);
    return ok;
}

/* predictiveDistribution for regression problems, BS:1437-1483: the mixture components are the generating
 * distribution with sample m's parameters substituted at input q (expressionToFunction, BS:1450-1462).
 * POLYREG: NormalDistribution[Sum_j c_j x^j, sigma] -> (mean, sd); LOGISTIC: class probabilities (K of them).
 * out[(m*Q + q)*C + c]; returns C (0: operator without independent variables). */
ORC_API int orc_predictive_components(const orc_problem *p, const double *theta, int64_t M, const double *xin,
                                      int64_t Q, double *out) {
    const int d = p->d;
    if (p->op == OP_POLYREG) {
        const int deg = p->iparam[0];
        for (int64_t m = 0; m < M; ++m)
            for (int64_t q = 0; q < Q; ++q) {
                const double *th = theta + m * d;
                double t = th[deg];
                for (int j = deg - 1; j >= 0; --j) t = t * xin[q] + th[j];
                const int ok = th[deg + 1] > 0.0;
                out[(m * Q + q) * 2 + 0] = ok ? t : NAN;
                out[(m * Q + q) * 2 + 1] = ok ? th[deg + 1] : NAN;
            }
        return 2;
    }
    if (p->op == OP_LOGISTIC) {
        const int K = p->iparam[1], F = p->n_in;
        for (int64_t m = 0; m < M; ++m)
            for (int64_t q = 0; q < Q; ++q) {
                const double *th = theta + m * d;
                double *o = out + (m * Q + q) * K;
                double zmax = 0.0, den = 0.0;
                for (int k = 0; k < K - 1; ++k) {
                    double z = th[k * (F + 1) + F];
                    for (int f = 0; f < F; ++f) z += th[k * (F + 1) + f] * xin[q * F + f];
                    o[k] = z;
                    if (z > zmax) zmax = z;
                }
                o[K - 1] = 0.0;
                for (int k = 0; k < K; ++k) { o[k] = exp(o[k] - zmax); den += o[k]; }
                for (int k = 0; k < K; ++k) o[k] /= den;
            }
        return K;
    }
    return 0;
}

/* BS:491-494 / BS:580-583: If[constraints[theta], Sum[...], logzero]; RuntimeErrorHandler -> logzero */
ORC_API double orc_loglike(const orc_problem *p, const orc_prior *pr, const double *th) {
    if (pr && !in_box(pr, th)) return p->logzero;
    if (!op_constraints_ok(p, th)) return p->logzero;
    double v;
    if (p->op == OP_GP_SE) v = gp_loglike_lu(p, th);
    else v = sum_ops_d(p, th);
    return isfinite(v) ? v : p->logzero;
}

/* quad-precision value of the same sum (no box guard) — grading reference */
ORC_API void orc_loglike_q(const orc_problem *p, const double *th, double *hi, double *lo) {
    __float128 v;
    if (p->op == OP_GP_SE) { int ok; v = gp_loglike_chol_ld(p, th, &ok); if (!ok) v = p->logzero; }
    else v = sum_ops_q(p, th);
    *hi = (double)v;
    *lo = (double)(v - (__float128)*hi);
}

ORC_API orc_problem *orc_problem_create(int op, int d, int64_t n, int n_in, int n_out, const double *in,
                                        const double *out, const int *iparam, double logzero) {
    orc_problem *p = (orc_problem *)calloc(1, sizeof(orc_problem));
    p->op = op; p->d = d; p->n = n; p->n_in = n_in; p->n_out = n_out;
    double *ci = (double *)malloc(sizeof(double) * (size_t)(n * (n_in > 0 ? n_in : 1)));
    double *co = (double *)malloc(sizeof(double) * (size_t)(n * (n_out > 0 ? n_out : 1)));
    if (in) memcpy(ci, in, sizeof(double) * (size_t)(n * n_in));
    if (out) memcpy(co, out, sizeof(double) * (size_t)(n * n_out));
    p->in = ci; p->out = co;
    for (int i = 0; i < 4; ++i) p->iparam[i] = iparam ? iparam[i] : 0;
    p->logzero = logzero;
    return p;
}
ORC_API void orc_problem_free(orc_problem *p) {
    if (!p) return;
    free((void *)p->in); free((void *)p->out); free(p->aux); free(p);
}
ORC_API orc_prior *orc_prior_create(int d, const int *kind, const double *lo, const double *hi,
                                    const double *p0, const double *p1, double logzero) {
    orc_prior *pr = (orc_prior *)calloc(1, sizeof(orc_prior));
    pr->d = d; pr->logzero = logzero;
    for (int j = 0; j < d; ++j) {
        pr->kind[j] = kind[j]; pr->lo[j] = lo[j]; pr->hi[j] = hi[j];
        pr->p0[j] = p0 ? p0[j] : 0.0; pr->p1[j] = p1 ? p1[j] : 1.0;
    }
    return pr;
}
ORC_API void orc_prior_free(orc_prior *pr) { free(pr); }

ORC_API void orc_loglike_batch(const orc_problem *p, const orc_prior *pr, const double *theta, int64_t P,
                               double *out, int threads) {
#pragma omp parallel for num_threads(threads > 0 ? threads : 1) schedule(dynamic, 1)
    for (int64_t i = 0; i < P; ++i) out[i] = orc_loglike(p, pr, theta + i * p->d);
}
ORC_API void orc_logprior_batch(const orc_prior *pr, const double *theta, int64_t P, double *out) {
    for (int64_t i = 0; i < P; ++i) out[i] = orc_logprior(pr, theta + i * pr->d);
}

/* ------------------------------------------------------------------------------------------ */
/* Small dense helpers (d <= 16)                                                              */
/* ------------------------------------------------------------------------------------------ */
static void mean_cov(const double *pts, int64_t n, int d, double *mean, double *cov) {
    /* Mean / Covariance (n-1 normalisation), BS:922-923, BS:989 */
    for (int j = 0; j < d; ++j) mean[j] = 0.0;
    for (int64_t i = 0; i < n; ++i) for (int j = 0; j < d; ++j) mean[j] += pts[i * d + j];
    for (int j = 0; j < d; ++j) mean[j] /= (double)n;
    for (int a = 0; a < d * d; ++a) cov[a] = 0.0;
    for (int64_t i = 0; i < n; ++i)
        for (int a = 0; a < d; ++a)
            for (int b = 0; b <= a; ++b) cov[a * d + b] += (pts[i * d + a] - mean[a]) * (pts[i * d + b] - mean[b]);
    for (int a = 0; a < d; ++a)
        for (int b = 0; b <= a; ++b) { cov[a * d + b] /= (double)(n - 1); cov[b * d + a] = cov[a * d + b]; }
}

/* proposal factor: chol( s_d * (C + eps*I) ), s_d = 2.4^2/d, eps = 1e-10 * mean(diag C) + 1e-300 */
static int proposal_chol(const double *cov, int d, double *L) {
    double tr = 0.0;
    for (int a = 0; a < d; ++a) tr += cov[a * d + a];
    const double eps = 1e-10 * (tr / d) + 1e-300;
    const double sd = 2.4 * 2.4 / (double)d;
    for (int a = 0; a < d * d; ++a) L[a] = 0.0;
    for (int j = 0; j < d; ++j) {
        double s = sd * (cov[j * d + j] + eps);
        for (int k = 0; k < j; ++k) s -= L[j * d + k] * L[j * d + k];
        if (!(s > 0.0)) return 0;
        const double l = sqrt(s);
        L[j * d + j] = l;
        for (int i = j + 1; i < d; ++i) {
            double t = sd * 0.5 * (cov[i * d + j] + cov[j * d + i]); /* symmetrizeMatrix BS:705, 716 */
            for (int k = 0; k < j; ++k) t -= L[i * d + k] * L[j * d + k];
            L[i * d + j] = t / l;
        }
    }
    return 1;
}

/* lexicographic (logL, point) order — BS:814, BS:902 SortBy[{#LogLikelihood, #Point}&] */
typedef struct { double logL; const double *pt; int d; int64_t idx; } sort_item;
static int cmp_items(const void *a, const void *b) {
    const sort_item *x = (const sort_item *)a, *y = (const sort_item *)b;
    if (x->logL < y->logL) return -1;
    if (x->logL > y->logL) return 1;
    for (int j = 0; j < x->d; ++j) {
        if (x->pt[j] < y->pt[j]) return -1;
        if (x->pt[j] > y->pt[j]) return 1;
    }
    return (x->idx > y->idx) - (x->idx < y->idx);
}

/* ------------------------------------------------------------------------------------------ */
/* The nested-sampling loop — BS:859-1040 with the walk protocol of BS:707-745                 */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    int64_t pool_size;   /* "SamplePoolSize"  BS:839 */
    int64_t batch_k;     /* points replaced per iteration; 1 = the reference scheme BS:980-1018 */
    int64_t mc_steps;    /* "MonteCarloSteps" BS:844 -> {S, S, 5S} BS:872 */
    int64_t max_iter;    /* "MaxIterations"   BS:841 */
    int64_t min_iter;    /* "MinIterations"   BS:842 */
    double term_frac;    /* "TerminationFraction" BS:845 */
    double acc_min, acc_max; /* "MinMaxAcceptanceRate" BS:848 */
    uint64_t seed;
    int64_t run_id;
    int64_t adapt_in_walk; /* 1: Haario recursion drives the proposal inside the walk (reference-like);
                              0: proposal factor frozen per iteration (what the CUDA walk does) */
    double loglmax;        /* numeric "LogLikelihoodMaximum" BS:847, 925-932; NaN = Automatic */
} orc_options;

typedef struct {
    int d;
    int64_t M, cap, n;
    int64_t n_deleted, iterations, evals;
    double *points, *logL, *logPrior, *acc;
    int64_t *pool;
    double *logX, *crude_logw;
    double crude_logZ, entropy, logLmax;
    double mean[ORC_MAXD], cov[ORC_MAXD * ORC_MAXD];
} orc_run;

static void run_reserve(orc_run *r, int64_t need) {
    if (need <= r->cap) return;
    int64_t cap = r->cap ? r->cap : 1024;
    while (cap < need) cap *= 2;
    r->points = (double *)realloc(r->points, sizeof(double) * cap * r->d);
    r->logL = (double *)realloc(r->logL, sizeof(double) * cap);
    r->logPrior = (double *)realloc(r->logPrior, sizeof(double) * cap);
    r->acc = (double *)realloc(r->acc, sizeof(double) * cap);
    r->pool = (int64_t *)realloc(r->pool, sizeof(int64_t) * cap);
    r->logX = (double *)realloc(r->logX, sizeof(double) * cap);
    r->crude_logw = (double *)realloc(r->crude_logw, sizeof(double) * cap);
    r->cap = cap;
}

/* BS:812-831 calculateWeightsCrude on (dead list already in order) + (sorted live set) */
static void crude_weights(orc_run *r, int64_t n_dead, const sort_item *live, int64_t n,
                          const double *live_logL, double *logZ, double *entropy, double *logXmin) {
    const int64_t M = n_dead + n;
    run_reserve(r, M);
    orc_xvalues_log_pool(n, n_dead, r->pool, r->logX);
    double *lw = (double *)malloc(sizeof(double) * M);
    double *LL = (double *)malloc(sizeof(double) * M);
    orc_trapezoid_log(r->logX, M, lw);
    for (int64_t k = 0; k < n_dead; ++k) LL[k] = r->logL[k];
    for (int64_t j = 0; j < n; ++j) LL[n_dead + j] = live_logL[live[j].idx];
    for (int64_t k = 0; k < M; ++k) lw[k] += LL[k];
    *logZ = orc_logsumexp(lw, M);
    *entropy = orc_entropy(lw, LL, M, *logZ);
    *logXmin = r->logX[M - 1];
    memcpy(r->crude_logw, lw, sizeof(double) * M);
    free(lw); free(LL);
}

ORC_API orc_run *orc_nested_sampling(const orc_problem *p, const orc_prior *pr, const orc_options *o,
                                     const double *start_points) {
    const int d = p->d;
    const int64_t n = o->pool_size;
    const int64_t S = o->mc_steps; /* {S, S, 5S} BS:872 */
    const uint32_t run_id = (uint32_t)o->run_id;
    orc_run *r = (orc_run *)calloc(1, sizeof(orc_run));
    r->d = d; r->n = n;
    run_reserve(r, 4 * n);

    double *lp = (double *)malloc(sizeof(double) * n * d);
    double *lL = (double *)malloc(sizeof(double) * n);
    double *lPr = (double *)malloc(sizeof(double) * n);
    double *lAcc = (double *)malloc(sizeof(double) * n);
    sort_item *ord = (sort_item *)malloc(sizeof(sort_item) * n);

    if (start_points) memcpy(lp, start_points, sizeof(double) * n * d);
    else orc_sample_prior(pr, n, o->seed, run_id, lp);
    /* BS:902-916 */
    for (int64_t i = 0; i < n; ++i) {
        lL[i] = orc_loglike(p, pr, lp + i * d);
        lPr[i] = orc_logprior(pr, lp + i * d);
        lAcc[i] = NAN; /* Missing["InitialSample"] BS:911 */
        r->evals++;
    }
    double meanEst[ORC_MAXD], covEst[ORC_MAXD * ORC_MAXD], covLive[ORC_MAXD * ORC_MAXD], mtmp[ORC_MAXD];
    mean_cov(lp, n, d, meanEst, covEst); /* BS:922-923 */

    const int64_t maxit = o->max_iter > o->min_iter ? o->max_iter : o->min_iter; /* BS:867-868 */
    const int64_t minit = o->max_iter > o->min_iter ? o->min_iter : o->max_iter;
    int64_t iteration = 1, n_dead = 0, walk_id = 0;
    double logZ = p->logzero, entropy = 0.0, logXmin = 0.0, logLmax = -INFINITY;
    const double log_frac = log(o->term_frac);

    double *wpt = (double *)malloc(sizeof(double) * o->batch_k * d);
    double *wL = (double *)malloc(sizeof(double) * o->batch_k);
    double *wPr = (double *)malloc(sizeof(double) * o->batch_k);
    double *wAcc = (double *)malloc(sizeof(double) * o->batch_k);

    for (;;) {
        /* BS:967-978; the product X_min*L_max <= Z*frac is tested in the log domain (SURVEY §7
         * hard part 2: the reference's Exp underflows for logL << -745) */
        int go = iteration <= maxit &&
                 (iteration == 1 || iteration <= minit ||
                  !(logXmin + (isnan(o->loglmax) ? logLmax : o->loglmax) <= logZ + log_frac)); /* BS:925-937 */
        if (!go) break;
        int64_t Kb = o->batch_k;
        if (Kb > maxit - iteration + 1) Kb = maxit - iteration + 1;
        if (Kb > n - 1) Kb = n - 1;

        for (int64_t i = 0; i < n; ++i) { ord[i].logL = lL[i]; ord[i].pt = lp + i * d; ord[i].d = d; ord[i].idx = i; }
        qsort(ord, n, sizeof(sort_item), cmp_items);
        const double Lstar = ord[Kb - 1].logL; /* BS:981 (K=1: Min) */
        mean_cov(lp, n, d, mtmp, covLive);
        for (int a = 0; a < d * d; ++a) covEst[a] = (covEst[a] + covLive[a]) / 2.0; /* BS:989 */

        double Lfac[ORC_MAXD * ORC_MAXD];
        int chol_ok = proposal_chol(covEst, d, Lfac);

        double msum[ORC_MAXD] = {0}, csum[ORC_MAXD * ORC_MAXD] = {0};
        for (int64_t j = 0; j < Kb; ++j, ++walk_id) {
            /* BS:993 RandomChoice of a live point (here: of the survivors, which satisfy L > L*) */
            double u[2];
            orc_uniform2(o->seed, 0, 0, (uint32_t)walk_id, TAG_START, run_id, u);
            int64_t pick = Kb + (int64_t)(u[0] * (double)(n - Kb));
            if (pick > n - 1) pick = n - 1;
            const int64_t src = ord[pick].idx;
            double x[ORC_MAXD], xn[ORC_MAXD], m[ORC_MAXD], C[ORC_MAXD * ORC_MAXD], Lw[ORC_MAXD * ORC_MAXD];
            memcpy(x, lp + src * d, sizeof(double) * d);
            double xL = lL[src], xPr = lPr[src];
            memcpy(m, meanEst, sizeof(double) * d);
            memcpy(C, covEst, sizeof(double) * d * d);
            memcpy(Lw, Lfac, sizeof(double) * d * d);
            int lw_ok = chol_ok;
            /* outer loop BS:995-1004: a walk whose acceptance rate ends outside the range is started again from a
             * fresh RandomChoice with Ceiling[factor S] steps, factor *= 1.25, keeping its own chain estimates
             * (BS:999).  The reference loops without bound; both sides stop after ORC_MAX_RETRY rounds.  Philox counter
             * word 0 is shifted by 16 * attempt so that every attempt has its own stream. */
            int64_t steps = 0, accepted = 0;
            double factor = 1.0;
            for (int attempt = 0;; ++attempt) {
                const int64_t S_k = attempt == 0 ? S : (int64_t)ceil(factor * (double)S);
                const int64_t maxS_k = 5 * S_k;
                if (attempt > 0) {
                    orc_uniform2(o->seed, (uint32_t)attempt, 0, (uint32_t)walk_id, TAG_START, run_id, u);
                    pick = Kb + (int64_t)(u[0] * (double)(n - Kb));
                    if (pick > n - 1) pick = n - 1;
                    const int64_t s2 = ord[pick].idx;
                    memcpy(x, lp + s2 * d, sizeof(double) * d);
                    xL = lL[s2]; xPr = lPr[s2];
                }
                double t = 10.0; /* startingIteration BS:715 */
                int64_t target = S_k;
                steps = 0; accepted = 0;
                for (;;) {
                    for (; steps < target; ++steps) {
                        if (o->adapt_in_walk) lw_ok = proposal_chol(C, d, Lw);
                        double z[ORC_MAXD + 1];
                        for (int b = 0; b < (d + 1) / 2; ++b)
                            orc_normal2(o->seed, (uint32_t)(b + 16 * attempt), (uint32_t)steps, (uint32_t)walk_id, TAG_NORMAL,
                                        run_id, z + 2 * b);
                        for (int a = 0; a < d; ++a) {
                            double s = x[a];
                            if (lw_ok) for (int b = 0; b <= a; ++b) s += Lw[a * d + b] * z[b];
                            xn[a] = s;
                        }
                        double ua[2];
                        orc_uniform2(o->seed, (uint32_t)(16 * attempt), (uint32_t)steps, (uint32_t)walk_id, TAG_ACCEPT, run_id, ua);
                        /* nsDensity BS:602-617: box && logL > L* (strict) ? logPrior : logzero */
                        int acc = 0;
                        double nL = p->logzero, nPr = p->logzero;
                        if (in_box(pr, xn)) {
                            nL = orc_loglike(p, pr, xn);
                            r->evals++;
                            if (nL > Lstar) {
                                nPr = orc_logprior(pr, xn);
                                if (nPr - xPr > log(ua[0])) acc = 1;
                            }
                        }
                        if (acc) { memcpy(x, xn, sizeof(double) * d); xL = nL; xPr = nPr; ++accepted; }
                        /* Haario recursion on the chain state */
                        double mo[ORC_MAXD];
                        memcpy(mo, m, sizeof(double) * d);
                        for (int a = 0; a < d; ++a) m[a] += (x[a] - m[a]) / (t + 1.0);
                        for (int a = 0; a < d; ++a)
                            for (int b = 0; b < d; ++b)
                                C[a * d + b] = (t - 1.0) / t * C[a * d + b] + (x[a] - mo[a]) * (x[b] - m[b]) / t;
                        t += 1.0;
                    }
                    const double rate = (double)accepted / (double)steps;
                    /* BS:730-736 */
                    if ((rate >= o->acc_min && rate <= o->acc_max) || steps >= maxS_k) break;
                    target += S_k;
                }
                const double rate = (double)accepted / (double)steps;
                if ((rate >= o->acc_min && rate <= o->acc_max) || attempt >= ORC_MAX_RETRY) break; /* BS:1000-1003 */
                factor *= 1.25;
            }
            memcpy(wpt + j * d, x, sizeof(double) * d);
            wL[j] = xL; wPr[j] = xPr; wAcc[j] = (double)accepted / (double)steps;
            for (int a = 0; a < d; ++a) msum[a] += m[a];
            for (int a = 0; a < d * d; ++a) csum[a] += 0.5 * (C[a] + C[(a % d) * d + a / d]);
        }
        /* BS:999: {meanEst, covEst} <- chain state (mean over the batch's walkers) */
        for (int a = 0; a < d; ++a) meanEst[a] = msum[a] / (double)Kb;
        for (int a = 0; a < d * d; ++a) covEst[a] = csum[a] / (double)Kb;

        /* kill the Kb worst in order; the j-th removed sees pool size n-j; BS:1006-1016 */
        run_reserve(r, n_dead + Kb + n);
        for (int64_t j = 0; j < Kb; ++j) {
            const int64_t src = ord[j].idx;
            memcpy(r->points + (n_dead + j) * d, lp + src * d, sizeof(double) * d);
            r->logL[n_dead + j] = lL[src]; r->logPrior[n_dead + j] = lPr[src];
            r->acc[n_dead + j] = lAcc[src]; r->pool[n_dead + j] = n - j;
        }
        for (int64_t j = 0; j < Kb; ++j) {
            const int64_t dst = ord[j].idx;
            memcpy(lp + dst * d, wpt + j * d, sizeof(double) * d);
            lL[dst] = wL[j]; lPr[dst] = wPr[j]; lAcc[dst] = wAcc[j];
        }
        n_dead += Kb;
        /* BS:1006-1020 */
        for (int64_t i = 0; i < n; ++i) { ord[i].logL = lL[i]; ord[i].pt = lp + i * d; ord[i].d = d; ord[i].idx = i; }
        qsort(ord, n, sizeof(sort_item), cmp_items);
        crude_weights(r, n_dead, ord, n, lL, &logZ, &entropy, &logXmin);
        logLmax = ord[n - 1].logL;
        iteration += Kb;
    }
    /* final sample list: dead + sorted live (BS:1026-1032) */
    for (int64_t i = 0; i < n; ++i) { ord[i].logL = lL[i]; ord[i].pt = lp + i * d; ord[i].d = d; ord[i].idx = i; }
    qsort(ord, n, sizeof(sort_item), cmp_items);
    run_reserve(r, n_dead + n);
    for (int64_t j = 0; j < n; ++j) {
        const int64_t src = ord[j].idx;
        memcpy(r->points + (n_dead + j) * d, lp + src * d, sizeof(double) * d);
        r->logL[n_dead + j] = lL[src]; r->logPrior[n_dead + j] = lPr[src];
        r->acc[n_dead + j] = lAcc[src]; r->pool[n_dead + j] = n - j;
    }
    crude_weights(r, n_dead, ord, n, lL, &logZ, &entropy, &logXmin);
    r->M = n_dead + n; r->n_deleted = n_dead; r->iterations = iteration - 1;
    r->crude_logZ = logZ; r->entropy = entropy; r->logLmax = ord[n - 1].logL;
    memcpy(r->mean, meanEst, sizeof(double) * d);
    memcpy(r->cov, covEst, sizeof(double) * d * d);
    free(lp); free(lL); free(lPr); free(lAcc); free(ord); free(wpt); free(wL); free(wPr); free(wAcc);
    return r;
}

ORC_API void orc_run_sizes(const orc_run *r, int64_t *M, int64_t *n_deleted, int64_t *iterations, int64_t *evals) {
    *M = r->M; *n_deleted = r->n_deleted; *iterations = r->iterations; *evals = r->evals;
}
ORC_API void orc_run_fetch(const orc_run *r, double *points, double *logL, double *logPrior, double *acc,
                           int64_t *pool, double *logX, double *crude_logw, double *summary /*[3]*/) {
    memcpy(points, r->points, sizeof(double) * r->M * r->d);
    memcpy(logL, r->logL, sizeof(double) * r->M);
    memcpy(logPrior, r->logPrior, sizeof(double) * r->M);
    memcpy(acc, r->acc, sizeof(double) * r->M);
    memcpy(pool, r->pool, sizeof(int64_t) * r->M);
    memcpy(logX, r->logX, sizeof(double) * r->M);
    memcpy(crude_logw, r->crude_logw, sizeof(double) * r->M);
    summary[0] = r->crude_logZ; summary[1] = r->entropy; summary[2] = r->logLmax;
}
ORC_API void orc_run_free(orc_run *r) {
    if (!r) return;
    free(r->points); free(r->logL); free(r->logPrior); free(r->acc); free(r->pool); free(r->logX);
    free(r->crude_logw); free(r);
}

/* ------------------------------------------------------------------------------------------ */
/* evidenceSampling — BS:1158-1291                                                             */
/* ------------------------------------------------------------------------------------------ */
/* Inputs: samples sorted by (logL, point) [M], per-sample pool sizes for the n_deleted dead ones,
 * pool size n of the final live set.  Outputs (all optional):
 *   z[nruns]                      logZ draws                                   BS:1228
 *   logw_mean/sd[M]               LogPosteriorWeight mean / sd over draws      BS:1245-1250
 *   slx_mean/sd[M]                SampledLogX mean / sd                        BS:1244
 *   pmean[nruns*d]                parameter means per draw                     BS:1230-1235
 *   H[nruns]                      relative entropy per draw                    BS:1263-1268
 * Draws: dead  logX_k = -cumsum Exp(rate pool_k)  (BS:1217-1224; pool_k == n in the reference),
 *        live  logX   = logX_dead_last - sorted Exp(1) draws (BS:1209-1215: Exp(1) truncated to
 *        [-min logX, inf) is the shift by -min logX).  Draw (run r, sample k) uses Philox
 *        (c0 = 0, c1 = r, c2 = k) with TAG_EV_DEAD / TAG_EV_LIVE.                              */
static int cmp_double(const void *a, const void *b) {
    const double x = *(const double *)a, y = *(const double *)b;
    return (x > y) - (x < y);
}
ORC_API void orc_evidence_sampling(int64_t M, int d, const double *points, const double *logL,
                                   const int64_t *pool, int64_t n, int64_t nruns, uint64_t seed,
                                   int sorted_draws, double *z, double *logw_mean, double *logw_sd, double *slx_mean,
                                   double *slx_sd, double *pmean, double *H) {
    const int64_t nd = M - n;
    double *lx = (double *)malloc(sizeof(double) * M);
    double *lw = (double *)malloc(sizeof(double) * M);
    double *e = (double *)malloc(sizeof(double) * n);
    double *s1 = (double *)calloc(M, sizeof(double)), *s2 = (double *)calloc(M, sizeof(double));
    double *x1 = (double *)calloc(M, sizeof(double)), *x2 = (double *)calloc(M, sizeof(double));
    for (int64_t r = 0; r < nruns; ++r) {
        double acc = 0.0, u[2];
        for (int64_t k = 0; k < nd; ++k) {
            orc_uniform2(seed, 0, (uint32_t)r, (uint32_t)k, TAG_EV_DEAD, 0, u);
            acc += log(u[0]) / (double)pool[k];
            lx[k] = acc;
        }
        if (sorted_draws) { /* literal BS:1209-1215: Sort of n i.i.d. Exp(1) draws */
            for (int64_t j = 0; j < n; ++j) {
                orc_uniform2(seed, 0, (uint32_t)r, (uint32_t)j, TAG_EV_LIVE, 0, u);
                e[j] = -log(u[0]);
            }
            qsort(e, n, sizeof(double), cmp_double);
            for (int64_t j = 0; j < n; ++j) lx[nd + j] = acc - e[j];
        } else { /* same law without the sort (Renyi representation of exponential order
                    statistics): e_(j) = Sum_{i<=j} E_i/(n-i); this is what the CUDA kernel draws */
            double c = 0.0;
            for (int64_t j = 0; j < n; ++j) {
                orc_uniform2(seed, 0, (uint32_t)r, (uint32_t)j, TAG_EV_LIVE, 0, u);
                c += log(u[0]) / (double)(n - j);
                lx[nd + j] = acc + c;
            }
        }
        orc_trapezoid_log(lx, M, lw);
        for (int64_t k = 0; k < M; ++k) lw[k] += logL[k];
        const double zr = orc_logsumexp(lw, M);
        if (z) z[r] = zr;
        double h = 0.0;
        if (pmean) for (int j = 0; j < d; ++j) pmean[r * d + j] = 0.0;
        for (int64_t k = 0; k < M; ++k) {
            const double lpw = lw[k] - zr, w = exp(lpw);
            s1[k] += lpw; s2[k] += lpw * lpw;
            x1[k] += lx[k]; x2[k] += lx[k] * lx[k];
            h += w * (isfinite(logL[k]) ? logL[k] : 0.0);
            if (pmean) for (int j = 0; j < d; ++j) pmean[r * d + j] += w * points[k * d + j];
        }
        if (H) H[r] = h - zr;
    }
    const double R = (double)nruns;
    for (int64_t k = 0; k < M; ++k) {
        /* meanAndError BS:1138-1149: Mean and (n-1) StandardDeviation */
        const double m1 = s1[k] / R, mx = x1[k] / R;
        if (logw_mean) logw_mean[k] = m1;
        if (logw_sd) { double v = (s2[k] - R * m1 * m1) / (R - 1.0); logw_sd[k] = v > 0 ? sqrt(v) : 0.0; }
        if (slx_mean) slx_mean[k] = mx;
        if (slx_sd) { double v = (x2[k] - R * mx * mx) / (R - 1.0); slx_sd[k] = v > 0 ? sqrt(v) : 0.0; }
    }
    free(lx); free(lw); free(e); free(s1); free(s2); free(x1); free(x2);
}

/* ------------------------------------------------------------------------------------------ */
/* Throughput leg for bench.py (cpu_baseline / --impl reference): the reference scheme          */
/* (one walker, S sequential density evaluations per replacement, BS:729) on T host threads,    */
/* each thread an independent chain as parallelNestedSampling's subkernels are (BS:1349-1357). */
/* Returns the number of likelihood evaluations performed.                                     */
/* ------------------------------------------------------------------------------------------ */
ORC_API int64_t orc_bench_walks(const orc_problem *p, const orc_prior *pr, const double *start_points,
                                int64_t n_start, double Lstar, int64_t replacements_per_thread, int64_t S,
                                uint64_t seed, int threads, double *sink) {
    int64_t total = 0;
    const int d = p->d;
#pragma omp parallel for num_threads(threads > 0 ? threads : 1) reduction(+ : total) schedule(static, 1)
    for (int th = 0; th < (threads > 0 ? threads : 1); ++th) {
        double mean[ORC_MAXD], cov[ORC_MAXD * ORC_MAXD], L[ORC_MAXD * ORC_MAXD];
        mean_cov(start_points, n_start, d, mean, cov);
        proposal_chol(cov, d, L);
        double acc_sink = 0.0;
        for (int64_t rep = 0; rep < replacements_per_thread; ++rep) {
            double x[ORC_MAXD], xn[ORC_MAXD];
            const int64_t src = (th * 7919 + rep * 104729) % n_start;
            memcpy(x, start_points + src * d, sizeof(double) * d);
            double xPr = orc_logprior(pr, x);
            for (int64_t s = 0; s < S; ++s) {
                double z[ORC_MAXD + 1], ua[2];
                for (int b = 0; b < (d + 1) / 2; ++b)
                    orc_normal2(seed, (uint32_t)b, (uint32_t)s, (uint32_t)rep, TAG_NORMAL, (uint32_t)th, z + 2 * b);
                for (int a = 0; a < d; ++a) {
                    double v = x[a];
                    for (int b = 0; b <= a; ++b) v += L[a * d + b] * z[b];
                    xn[a] = v;
                }
                orc_uniform2(seed, 0, (uint32_t)s, (uint32_t)rep, TAG_ACCEPT, (uint32_t)th, ua);
                /* nsDensity (BS:602-617) is If[box && logL > L*, ...]: a proposal outside the box is rejected
                 * without a likelihood evaluation, and is not counted as one */
                if (!in_box(pr, xn)) continue;
                const double nL = orc_loglike(p, NULL, xn);
                ++total;
                acc_sink += nL;
                if (nL > Lstar) {
                    const double nPr = orc_logprior(pr, xn);
                    if (nPr - xPr > log(ua[0])) { memcpy(x, xn, sizeof(double) * d); xPr = nPr; }
                }
            }
        }
        if (sink) sink[th] = acc_sink;
    }
    return total;
}

/* ------------------------------------------------------------------------------------------ */
/* createMCMCChain / iterateMCMC — BS:630-703                                                  */
/* ------------------------------------------------------------------------------------------ */
/* One adaptive-Metropolis chain on posteriorDensity = If[box, logPrior + logL, logzero] (BS:630-649), started at
 * `start` with {InitialCovariance, CovarianceLearnDelay} (BS:673-696).  The sampler itself is the closed-source
 * Statistics`MCMC`BuildMarkovChain[{"AdaptiveMetropolis","Log"}]; restated from Haario et al. (2001) like the
 * nested-sampling walk above: proposal N(x, C0) while fewer than `delay` states have been seen, then
 * N(x, s_d (C_t + eps I)); C_t by the same recursion, started at t = 1 with mean = start (=> unbiased sample
 * covariance of the visited states).  Philox: (pair, step, chain) under tags 9 / 10.
 * out: n_steps x d (state after every step); mean d, cov d x d, t_out, acc_out: final chain estimates. */
enum { TAG_MC_NORMAL = 9, TAG_MC_ACCEPT = 10 };
ORC_API int orc_mcmc_chain(const orc_problem *p, const orc_prior *pr, const double *start, const double *init_cov,
                           int64_t delay, uint64_t seed, uint32_t chain_id, int64_t n_steps, double *out,
                           double *mean, double *cov, int64_t *t_out, int64_t *acc_out) {
    const int d = p->d;
    double x[ORC_MAXD], xn[ORC_MAXD], m[ORC_MAXD], mo[ORC_MAXD], C[ORC_MAXD * ORC_MAXD], L0[ORC_MAXD * ORC_MAXD],
        L[ORC_MAXD * ORC_MAXD];
    if (delay < 2) delay = 2;
    memset(L0, 0, sizeof L0);
    for (int j = 0; j < d; ++j) {
        double s = init_cov[j * d + j];
        for (int k = 0; k < j; ++k) s -= L0[j * d + k] * L0[j * d + k];
        if (!(s > 0.0)) return 0;
        const double l = sqrt(s);
        L0[j * d + j] = l;
        for (int i = j + 1; i < d; ++i) {
            double t = 0.5 * (init_cov[i * d + j] + init_cov[j * d + i]);
            for (int k = 0; k < j; ++k) t -= L0[i * d + k] * L0[j * d + k];
            L0[i * d + j] = t / l;
        }
    }
    memcpy(x, start, sizeof(double) * d);
    memcpy(m, start, sizeof(double) * d);
    memset(C, 0, sizeof C);
    if (!in_box(pr, x)) return 0;
    double lp = orc_logprior(pr, x) + orc_loglike(p, pr, x);
    if (!(lp > 0.5 * p->logzero) || !isfinite(lp)) return 0;
    int64_t t = 1, acc = 0;
    for (int64_t k = 0; k < n_steps; ++k) {
        const double *Lw = L0;
        if (t >= delay && proposal_chol(C, d, L)) Lw = L;
        double z[ORC_MAXD + 1], ua[2];
        const uint32_t hi = (uint32_t)((uint64_t)k >> 32);
        for (int b = 0; b < (d + 1) / 2; ++b) orc_normal2(seed, (uint32_t)b, (uint32_t)k, chain_id, TAG_MC_NORMAL, hi, z + 2 * b);
        for (int a = 0; a < d; ++a) {
            double s = x[a];
            for (int b = 0; b <= a; ++b) s += Lw[a * d + b] * z[b];
            xn[a] = s;
        }
        orc_uniform2(seed, 0u, (uint32_t)k, chain_id, TAG_MC_ACCEPT, hi, ua);
        if (in_box(pr, xn)) {
            const double nPr = orc_logprior(pr, xn), nL = orc_loglike(p, pr, xn);
            if (nPr > 0.5 * p->logzero && nL > 0.5 * p->logzero && (nPr + nL) - lp > log(ua[0])) {
                memcpy(x, xn, sizeof(double) * d);
                lp = nPr + nL;
                ++acc;
            }
        }
        const double tf = (double)t;
        memcpy(mo, m, sizeof(double) * d);
        for (int a = 0; a < d; ++a) m[a] += (x[a] - m[a]) / (tf + 1.0);
        for (int a = 0; a < d; ++a)
            for (int b = 0; b < d; ++b) C[a * d + b] = (tf - 1.0) / tf * C[a * d + b] + (x[a] - mo[a]) * (x[b] - m[b]) / tf;
        t += 1;
        if (out) memcpy(out + k * d, x, sizeof(double) * d);
    }
    if (mean) memcpy(mean, m, sizeof(double) * d);
    if (cov) memcpy(cov, C, sizeof(double) * d * d);
    if (t_out) *t_out = t;
    if (acc_out) *acc_out = acc;
    return 1;
}

ORC_API int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
