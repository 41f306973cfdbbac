// operators.cuh — the fixed likelihood-operator table (per-datum device functions).
//
// Each operator replaces one symbolic builder of the reference:
//   OpGaussian / OpPolyReg / OpLogistic / OpGbm  <- logLikelihoodFunction BS:429-505 and
//   regressionLogLikelihoodFunction BS:517-595 (Sum over the data of the per-datum log-density,
//   guarded by the parameter constraints; failure -> logzero).
// Per-datum formulas: SURVEY.md §8a (WL LogLikelihood closed forms).  The parameter-independent
// additive constants are kept, because LogEvidence and the 1e-12 logL parity depend on them.
//
// Interface (lane = walker: every lane owns TW parameter vectors, rows are broadcast from smem):
//   D      number of parameters             NCOL  fp64 columns per device row
//   TW_MAX walkers register-tiled per lane
//   Row    per-theta coefficients used per datum (make_row); Coef = those of the epilogue only
//          prepare(th, ok) -> Coef   (ok = operator constraints)
//   rows<TW>(coef[TW], r, acc[TW])   accumulate one datum for TW walkers.  Written step-major (the same
//       Horner step for all TW walkers back to back) so that consecutive DFMAs share the row operand:
//       on sm_100 a DFMA with three distinct 64-bit register sources issues at 2/3 rate, with a shared
//       operand in the reuse cache at full rate (measured: scripts/dfma_patterns.cu).
//   finish(coef, acc, rows, cst) -> logL
#pragma once
#include "common.cuh"

namespace binest {

// ------------------------------------------------------------------ NormalDistribution[mu, sigma], i.i.d. data
struct OpGaussian {
    static constexpr int D = 2, NCOL = 1, TW_MAX = 8;
    struct Coef { double mu, h, lognorm; };
    struct Row { double mu; };
    using Acc = double;
    __device__ __forceinline__ static Acc acc_init() { return 0.0; }
    __device__ __forceinline__ static double acc_value(const Acc &a) { return a; }
    __device__ __forceinline__ static Row make_row(const double (&th)[D]) { return Row{th[0]}; }
    __device__ static Coef prepare(const double (&th)[D], bool &ok) {
        ok = th[1] > 0.0;  // DistributionParameterAssumptions, BS:439
        return Coef{th[0], 1.0 / (2.0 * th[1] * th[1]), -log(th[1]) - kHalfLog2Pi};
    }
    template <int TW>
    __device__ __forceinline__ static void rows(const Row (&c)[TW], const double *__restrict__ r, double (&acc)[TW]) {
        const double x = r[0];
        double e[TW];
#pragma unroll
        for (int u = 0; u < TW; ++u) e[u] = x - c[u].mu;
#pragma unroll
        for (int u = 0; u < TW; ++u) acc[u] = fma(e[u], e[u], acc[u]);
    }
    __device__ static double finish(const Coef &c, double acc, double rows, double) {
        return rows * c.lognorm - c.h * acc;
    }
};

// ------------------------------------------------------------------ NormalDistribution[Sum_j c_j x^j, sigma]
template <int DEG>
struct OpPolyReg {
    static constexpr int D = DEG + 2, NCOL = 2, TW_MAX = 8;
    struct Coef { double h, lognorm; };
    struct Row { double c[DEG + 1]; };
    using Acc = double;
    __device__ __forceinline__ static Acc acc_init() { return 0.0; }
    __device__ __forceinline__ static double acc_value(const Acc &a) { return a; }
    __device__ __forceinline__ static Row make_row(const double (&th)[D]) {
        Row c;
#pragma unroll
        for (int j = 0; j <= DEG; ++j) c.c[j] = th[j];
        return c;
    }
    __device__ static Coef prepare(const double (&th)[D], bool &ok) {
        Coef c;
        const double sg = th[DEG + 1];
        ok = sg > 0.0;  // BS:523
        c.h = 1.0 / (2.0 * sg * sg);
        c.lognorm = -log(sg) - kHalfLog2Pi;
        return c;
    }
    // 9 flop per datum-walker at DEG = 3: 3 Horner FMA, 1 subtract, 1 FMA-accumulate
    template <int TW>
    __device__ __forceinline__ static void rows(const Row (&c)[TW], const double *__restrict__ r, double (&acc)[TW]) {
        const double x = r[0], y = r[1];
        double t[TW];
#pragma unroll
        for (int u = 0; u < TW; ++u) t[u] = fma(c[u].c[DEG], x, c[u].c[DEG - 1]);
#pragma unroll
        for (int j = DEG - 2; j >= 0; --j)
#pragma unroll
            for (int u = 0; u < TW; ++u) t[u] = fma(t[u], x, c[u].c[j]);
#pragma unroll
        for (int u = 0; u < TW; ++u) t[u] = y - t[u];
#pragma unroll
        for (int u = 0; u < TW; ++u) acc[u] = fma(t[u], t[u], acc[u]);
    }
    __device__ static double finish(const Coef &c, double acc, double rows, double) {
        return rows * c.lognorm - c.h * acc;
    }
};

// ------------------------------------------------------------------ softmax classification, reference class K
// theta = K-1 blocks of (w_1..w_F, b); device row = (x_1..x_F, label, pad...) with NCOL even
template <int F, int K>
struct OpLogistic {
    static constexpr int D = (K - 1) * (F + 1), NCOL = (F + 2) & ~1, TW_MAX = 2;
    struct Coef { int unused; };
    struct Row { double w[K - 1][F + 1]; };
    __device__ __forceinline__ static Row make_row(const double (&th)[D]) {
        Row c;
#pragma unroll
        for (int k = 0; k < K - 1; ++k)
#pragma unroll
            for (int f = 0; f <= F; ++f) c.w[k][f] = th[k * (F + 1) + f];
        return c;
    }
    __device__ static Coef prepare(const double (&)[D], bool &ok) {
        ok = true;
        return Coef{0};
    }
    // log Sum_k exp z_k with z_K = 0, K <= 3, costs two exps and 1/32 of a log per datum:
    //   * after sorting, one of the K shifted terms is exactly exp(0) = 1:  s = 1 + exp(a) + exp(b), a, b <= 0;
    //   * Sum_i log s_i = log Prod_i s_i: s in (1, 3], so 32 factors (<= 3^32 = 1.9e15) are multiplied before one
    //     log is taken (relative rounding 32 * 1.1e-16 on the product, i.e. ~4e-15 absolute on a sum of ~32 terms).
    struct Acc { double lin, prod; int cnt; };
    __device__ __forceinline__ static Acc acc_init() { return Acc{0.0, 1.0, 0}; }
    __device__ __forceinline__ static double acc_value(const Acc &a) { return a.lin - log(a.prod); }
    template <int TW>
    __device__ __forceinline__ static void rows(const Row (&c)[TW], const double *__restrict__ r, Acc (&acc)[TW]) {
        static_assert(K == 2 || K == 3, "softmax operator is specialised for 2 or 3 classes");
        double z[TW][K - 1];
        const int lab = (int)r[F];
#pragma unroll
        for (int k = 0; k < K - 1; ++k)
#pragma unroll
            for (int u = 0; u < TW; ++u) z[u][k] = c[u].w[k][F];
#pragma unroll
        for (int f = 0; f < F; ++f) {
            const double x = r[f];
#pragma unroll
            for (int k = 0; k < K - 1; ++k)
#pragma unroll
                for (int u = 0; u < TW; ++u) z[u][k] = fma(c[u].w[k][f], x, z[u][k]);
        }
#pragma unroll
        for (int u = 0; u < TW; ++u) {
            double zy = 0.0;  // z_K = 0
#pragma unroll
            for (int k = 0; k < K - 1; ++k) zy = (lab == k) ? z[u][k] : zy;
            double mx, s;
            if (K == 2) {
                mx = fmax(z[u][0], 0.0);
                s = 1.0 + exp(-fabs(z[u][0]));
            } else {
                const double hi = fmax(z[u][0], z[u][K - 2]), lo = fmin(z[u][0], z[u][K - 2]);
                mx = fmax(hi, 0.0);
                s = (1.0 + exp(-fabs(hi))) + exp(lo - mx);
            }
            acc[u].lin += zy - mx;
            acc[u].prod *= s;
        }
        if (++acc[0].cnt == 32) {
#pragma unroll
            for (int u = 0; u < TW; ++u) {
                acc[u].lin -= log(acc[u].prod);
                acc[u].prod = 1.0;
            }
            acc[0].cnt = 0;
        }
    }
    __device__ static double finish(const Coef &, double acc, double, double) { return acc; }
};

// ------------------------------------------------------------------ GeometricBrownianMotionProcess[mu, sigma, x0]
// device row i = (a_i, b_i) = (r_i / sqrt(dt_i), sqrt(dt_i)), r_i = log(x_i / x_{i-1});
// cst = Sum_i(-log x_i - 1/2 log dt_i) - rows * 1/2 log 2pi  (parameter independent, fixed at upload)
struct OpGbm {
    static constexpr int D = 2, NCOL = 2, TW_MAX = 8;
    struct Coef { double h, lognorm; };
    struct Row { double negm; };
    using Acc = double;
    __device__ __forceinline__ static Acc acc_init() { return 0.0; }
    __device__ __forceinline__ static double acc_value(const Acc &a) { return a; }
    __device__ __forceinline__ static Row make_row(const double (&th)[D]) { return Row{-(th[0] - 0.5 * th[1] * th[1])}; }
    __device__ static Coef prepare(const double (&th)[D], bool &ok) {
        ok = th[1] > 0.0;
        return Coef{1.0 / (2.0 * th[1] * th[1]), -log(th[1])};
    }
    template <int TW>
    __device__ __forceinline__ static void rows(const Row (&c)[TW], const double *__restrict__ r, double (&acc)[TW]) {
        const double a = r[0], b = r[1];
        double e[TW];
#pragma unroll
        for (int u = 0; u < TW; ++u) e[u] = fma(c[u].negm, b, a);
#pragma unroll
        for (int u = 0; u < TW; ++u) acc[u] = fma(e[u], e[u], acc[u]);
    }
    __device__ static double finish(const Coef &c, double acc, double rows, double cst) {
        return rows * c.lognorm + cst - c.h * acc;
    }
};

// ------------------------------------------------------------------ GP marginal likelihood (gp.cu)
// Not a streaming reduction: gp.cu writes the finished logL where the walk expects the (single) partial sum,
// so only the epilogue interface is needed here.
struct OpGpSe {
    static constexpr int D = 3;
    struct Coef { int unused; };
    __device__ static Coef prepare(const double (&th)[D], bool &ok) {
        ok = th[0] > 0.0 && th[1] > 0.0 && th[2] > 0.0;
        return Coef{0};
    }
    __device__ static double finish(const Coef &, double acc, double, double) { return acc; }
};

// ------------------------------------------------------------------ priors (BS:25-64, BS:365-427)
struct PriorSpec {
    int d;
    int kind[BINEST_MAXD];
    double lo[BINEST_MAXD], hi[BINEST_MAXD];
    double p0[BINEST_MAXD], p1[BINEST_MAXD];
    double lognorm[BINEST_MAXD];  // per-dimension additive constant, fixed at problem creation
};

// open box lo < theta < hi (BS:327-336)
template <int D>
__device__ __forceinline__ bool in_box(const PriorSpec &pr, const double (&th)[D]) {
    bool ok = true;
#pragma unroll
    for (int j = 0; j < D; ++j) ok = ok && (th[j] > pr.lo[j]) && (th[j] < pr.hi[j]);
    return ok;
}
__device__ __forceinline__ bool in_box_dyn(const PriorSpec &pr, const double *th) {
    bool ok = true;
    for (int j = 0; j < pr.d; ++j) ok = ok && (th[j] > pr.lo[j]) && (th[j] < pr.hi[j]);
    return ok;
}
// log prior density inside the box (caller has checked the box)
__device__ __forceinline__ double logprior_dim(const PriorSpec &pr, int j, double t) {
    switch (pr.kind[j]) {
    case BINEST_PRIOR_UNIFORM: return pr.lognorm[j];             // -log(hi - lo)
    case BINEST_PRIOR_SCALE: return -log(t) + pr.lognorm[j];     // -log t - log log(hi/lo)
    default: {                                                   // truncated normal
        const double z = (t - pr.p0[j]) / pr.p1[j];
        return -0.5 * z * z + pr.lognorm[j];                     // - log s - 1/2 log 2pi - log mass
    }
    }
}
__device__ __forceinline__ double logprior_dyn(const PriorSpec &pr, const double *th, double logzero) {
    if (!in_box_dyn(pr, th)) return logzero;
    double s = 0.0;
    for (int j = 0; j < pr.d; ++j) s += logprior_dim(pr, j, th[j]);
    return isfinite(s) ? s : logzero;
}

}  // namespace binest
