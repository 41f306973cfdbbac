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
//   local(coef, acc, rows, cst) -> the sum in its shard-additive form (identity except for the polynomial operator,
//       whose moments trick is resolved with THIS GPU's data moments), and
//   RENORM > 0: renorm(acc[TW]) must be called at least every RENORM rows (softmax: running products); the row
//       loops go through sweep_rows (loglike.cuh), which does it.
//   finish_total(coef, total, rows_total, cst_total) -> logL.  Single GPU: op_finish = finish_total(local(.)).
//   Data-sharded mode: local() on every rank before the exchange, finish_total() on the rank-ordered sum after it.
#pragma once
#include "common.cuh"

namespace binest {

// parameter-independent constants of a problem's data, fixed at upload and handed to every operator epilogue:
//   c     additive constant (GBM: Sum_i(-log x_i - 1/2 log dt_i) - rows 1/2 log 2pi)
//   m[k]  data moments (polynomial regression: m[0] = Sum y, m[k] = Sum x^k, k = 1..degree)
// In the data-sharded mode these are the totals over all shards.
//   xbar, piv  pivots subtracted from the regression data at upload (polynomial regression; 0 otherwise)
struct OpCst {
    double c;
    double m[6];
    double xbar, piv;
};

template <class OP>
__device__ __forceinline__ double op_finish(const typename OP::Coef &c, double acc, double rows, const OpCst &k) {
    return OP::finish_total(c, OP::local(c, acc, rows, k), rows, k);
}

// ------------------------------------------------------------------ NormalDistribution[mu, sigma], i.i.d. data
struct OpGaussian {
    static constexpr int D = 2, NCOL = 1, TW_MAX = 8, RENORM = 0, SLOTS = 2;  // SLOTS: fp64 pipe instructions per datum
    struct Coef { double mu, h, lognorm; };
    struct Row { double mu; };
    using Acc = double;
    __device__ __forceinline__ static Acc acc_init() { return 0.0; }
    __device__ __forceinline__ static double acc_value(const Acc &a) { return a; }
    __device__ __forceinline__ static Row make_row(const double (&th)[D], const OpCst &) { return Row{th[0]}; }
    __device__ static Coef prepare(const double (&th)[D], bool &ok, const OpCst &) {
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
    __device__ static double local(const Coef &, double acc, double, const OpCst &) { return acc; }
    __device__ static double finish_total(const Coef &c, double acc, double rows, const OpCst &) {
        return rows * c.lognorm - c.h * acc;
    }
};

// ------------------------------------------------------------------ NormalDistribution[Sum_j c_j x^j, sigma]
// Per datum the reference evaluates -(y - Sum_j c_j x^j)^2 / (2 sigma^2) - log sigma - 1/2 log 2pi (BS:540-583).
// Device rows hold PIVOTED data, fixed at upload (abi_problem.cu):
//     x'_i = x_i - xbar      (xbar = the data mean of x, or 0 when the data are already centred: |mean| <= sd/2)
//     y'_i = y_i - piv       (piv = intercept of the least-squares polynomial in x')
// and the polynomial is re-expanded around xbar per walker: Sum_j c_j x^j = Sum_k c~_k x'^k (Taylor shift in
// double-double, polyreg_shift; the identity when xbar = 0).  The residual is then
//     e_i = t_i + delta,   t_i = x'_i (c~_1 + c~_2 x'_i + ...) - y'_i,   delta = c~_0 - piv
// (Horner on c~_deg..c~_1, the last FMA adds -y'), and Sum e^2 = Sum t^2 + delta (2 Sum t + N delta) with
//     Sum t = Sum_k c~_k m[k] - m[0]
// from the walker-independent moments of the pivoted data (OpCst).  The per-datum chain is deg FMAs + the
// FMA-accumulate: 4 fp64 pipe slots at degree 3 instead of 5 for the same 9 algorithmic flop of SURVEY §8d.
// Why the pivots: without them delta = c_0, and Sum t^2 ~ N c_0^2 cancels against c_0 (2 Sum t + N c_0) whenever
// |c_0| >> rms(e) (y offset 1000, sigma 0.01: 1e-6 relative on Sum e^2).  With them, a walker can only have
// N delta^2 >> Sum e^2 if the other coefficients compensate a constant over the data, and in the centred basis they
// cannot: min over them of Sum (delta + Sum_{k>=1} d_k x'^k)^2 = N delta^2 (1 - R^2), R^2 (the fraction of the
// constant explained by x', x'^2, ...) = 5/9 for a cubic on uniform x', so N delta^2 <= ~2.3 Sum e^2: at most one
// digit is lost, wherever the walker is.  tests/test_gpu_parity.py::test_polyreg_adversarial pins offsets of 50 and
// 1000 with sigma 0.05 / 0.01 and x in (100, 101) to 1e-12 against the __float128 oracle.
__device__ __forceinline__ void dd_fma_acc(double &h, double &l, double ah, double al, double x) {
    // (h, l) += (ah, al) * x in double-double
    const double ph = ah * x;
    double pl = fma(ah, x, -ph);
    pl = fma(al, x, pl);
    const double s = h + ph, bb = s - h;
    double e = (h - (s - bb)) + (ph - bb);
    e += l + pl;
    const double hn = s + e;
    l = e - (hn - s);
    h = hn;
}
// coefficients of p(x' + xbar) in powers of x' (repeated synthetic division), double-double throughout
template <int DEG>
__device__ __forceinline__ void polyreg_shift(const double *th, double xbar, double (&h)[DEG + 1], double &lo0) {
    double l[DEG + 1];
#pragma unroll
    for (int j = 0; j <= DEG; ++j) { h[j] = th[j]; l[j] = 0.0; }
#pragma unroll
    for (int i = 0; i < DEG; ++i)
#pragma unroll
        for (int j = DEG - 1; j >= i; --j) dd_fma_acc(h[j], l[j], h[j + 1], l[j + 1], xbar);
    lo0 = l[0];
}

template <int DEG>
struct OpPolyReg {
    static constexpr int D = DEG + 2, NCOL = 2, TW_MAX = 8, RENORM = 0, SLOTS = DEG + 1;
    struct Coef { double h, lognorm, c[DEG + 1]; };  // c[0] = delta, c[1..DEG] = shifted coefficients
    struct Row { double c[DEG + 1]; };               // c[0] unused per datum
    using Acc = double;
    __device__ __forceinline__ static Acc acc_init() { return 0.0; }
    __device__ __forceinline__ static double acc_value(const Acc &a) { return a; }
    __device__ __forceinline__ static Row make_row(const double (&th)[D], const OpCst &k) {
        Row c;
        if (k.xbar != 0.0) {
            double lo0;
            polyreg_shift<DEG>(th, k.xbar, c.c, lo0);
        } else {
#pragma unroll
            for (int j = 0; j <= DEG; ++j) c.c[j] = th[j];
        }
        return c;
    }
    __device__ static Coef prepare(const double (&th)[D], bool &ok, const OpCst &k) {
        Coef c;
        const double sg = th[DEG + 1];
        ok = sg > 0.0;  // BS:523
        c.h = 1.0 / (2.0 * sg * sg);
        c.lognorm = -log(sg) - kHalfLog2Pi;
        double lo0 = 0.0;
        if (k.xbar != 0.0) {
            polyreg_shift<DEG>(th, k.xbar, c.c, lo0);
        } else {
#pragma unroll
            for (int j = 0; j <= DEG; ++j) c.c[j] = th[j];
        }
        c.c[0] = (c.c[0] - k.piv) + lo0;
        return c;
    }
    template <int TW>
    __device__ __forceinline__ static void rows(const Row (&c)[TW], const double *__restrict__ r, double (&acc)[TW]) {
        const double x = r[0], y = r[1];
        double t[TW];
        if (DEG >= 2) {
#pragma unroll
            for (int u = 0; u < TW; ++u) t[u] = fma(c[u].c[DEG], x, c[u].c[DEG - 1]);
#pragma unroll
            for (int j = DEG - 2; j >= 1; --j)
#pragma unroll
                for (int u = 0; u < TW; ++u) t[u] = fma(t[u], x, c[u].c[j]);
        } else {
#pragma unroll
            for (int u = 0; u < TW; ++u) t[u] = c[u].c[1];
        }
#pragma unroll
        for (int u = 0; u < TW; ++u) t[u] = fma(t[u], x, -y);
#pragma unroll
        for (int u = 0; u < TW; ++u) acc[u] = fma(t[u], t[u], acc[u]);
    }
    // Sum e^2 over the rows the moments in k describe (this GPU's rows; a shard in the data-sharded mode)
    __device__ static double local(const Coef &c, double acc, double rows, const OpCst &k) {
        double st = -k.m[0];
#pragma unroll
        for (int j = 1; j <= DEG; ++j) st = fma(c.c[j], k.m[j], st);
        return fma(c.c[0], fma(rows, c.c[0], 2.0 * st), acc);  // Sum t^2 + delta (2 Sum t + N delta)
    }
    __device__ static double finish_total(const Coef &c, double sse, double rows, const OpCst &) {
        return rows * c.lognorm - c.h * sse;
    }
};

// ------------------------------------------------------------------ softmax classification, reference class K
// theta = K-1 blocks of (w_1..w_F, b); device row = (x_1..x_F, label, pad...) with NCOL even
//
// In-kernel exp.  libdevice exp() rebuilt its constants with ~50 integer instructions per call and carried
// slow-path branches; with two exps per datum the kernel was issue-bound at 0.47 of the fp64 pipe.  exp_bounded is
// branch-free, 14 fp64-pipe instructions: Cody-Waite reduction r = x - k ln2 (magic-number rounding, two FMAs),
// degree-10 polynomial (scripts/gen_exp_poly.py: max relative error 3.4e-16 on |r| <= ln2/2) whose coefficients
// sit in the constant bank (DFMA takes a c[][] operand: two register sources, full issue rate), then k is added to
// the exponent field.  Valid for |x| < 700 (result normal); callers route anything else through the slow path.
__constant__ double kExpPoly[11] = {
    1.0,
    1.0000000000000067,      // 0x3ff000000000001e
    0.5000000000000006,      // 0x3fe0000000000005
    0.16666666666554314,     // 0x3fc555555554b736
    0.041666666666573066,    // 0x3fa55555555520a4
    0.008333333385699212,    // 0x3f81111112ddae8b
    0.001388888893251478,    // 0x3f56c16c17f46982
    0.00019841170230570286,  // 0x3f2a01978b8b3d18
    2.480150431378554e-05,   // 0x3efa019a6611cad5
    2.764019739169484e-06,   // 0x3ec72fafebdaf273
    2.7626371065696354e-07,  // 0x3e928a2ca617b969
};
__constant__ double kExpRed[4] = {
    1.4426950408889634,         // log2(e)
    6755399441055744.0,         // 1.5 * 2^52: adding it leaves rint(x log2 e) in the low word
    -6.93147180369123816490e-01,  // -ln2, high part (trailing zeros)
    -1.90821492927058770002e-10,  // -ln2, low part
};
__device__ __forceinline__ bool exp_arg_bounded(double x) {  // |x| < 700, false for NaN/Inf
    return (__double2hiint(x) & 0x7fffffff) < 0x4085E000;
}
__device__ __forceinline__ bool exp_arg_bounded_170(double x) {  // |x| < 170, false for NaN/Inf
    return (__double2hiint(x) & 0x7fffffff) < 0x40654000;
}
// the N independent exps of one datum, step-major so that consecutive DFMAs come from different chains
template <int N>
__device__ __forceinline__ void exp_bounded(const double (&x)[N], double (&out)[N]) {
    double t[N], r[N], p[N];
    int k[N];
#pragma unroll
    for (int i = 0; i < N; ++i) t[i] = fma(x[i], kExpRed[0], kExpRed[1]);
#pragma unroll
    for (int i = 0; i < N; ++i) {
        k[i] = __double2loint(t[i]);
        t[i] -= kExpRed[1];
    }
#pragma unroll
    for (int i = 0; i < N; ++i) r[i] = fma(t[i], kExpRed[2], x[i]);
#pragma unroll
    for (int i = 0; i < N; ++i) r[i] = fma(t[i], kExpRed[3], r[i]);
#pragma unroll
    for (int i = 0; i < N; ++i) p[i] = fma(r[i], kExpPoly[10], kExpPoly[9]);
#pragma unroll
    for (int j = 8; j >= 0; --j)
#pragma unroll
        for (int i = 0; i < N; ++i) p[i] = fma(p[i], r[i], kExpPoly[j]);
#pragma unroll
    for (int i = 0; i < N; ++i) out[i] = __hiloint2double(__double2hiint(p[i]) + (k[i] << 20), __double2loint(p[i]));
}

// Table-driven variant (the one the softmax operator uses): x = (32 k + j) ln2/32 + r with |r| <= ln2/64, so
// exp(x) = 2^k * 2^(j/32) * exp(r) and a degree-5 polynomial suffices (scripts/gen_exp_poly.py 5 64: max relative error
// 1.4e-16) — 10 fp64-pipe instructions per exp instead of 14 (the operator is fp64-issue bound: 41 -> 33 per datum).
// 2^(j/32) comes from a 256-byte table read through the L1 (per-lane index, so not the constant bank).
__constant__ double kExpTabPoly[4] = {
    0.49999999998924655,   // 0x3fdffffffffd0b4b
    0.16666666666513047,   // 0x3fc5555555547d22
    0.04166691108715919,   // 0x3fa5555d88e3ae37
    0.008333368250528262,  // 0x3f811115c0cff61a
};
__constant__ double kExpTabRed[3] = {
    46.166241308446828,                  // 32 / ln2
    -6.93147180369123816490e-01 / 32.0,  // -ln2/32, high part (exact scaling of the Cody-Waite split above)
    -1.90821492927058770002e-10 / 32.0,  // -ln2/32, low part
};
__device__ const double kExp2Tab[32] = {
    1.0,  // 0x3ff0000000000000
    1.0218971486541166,  // 0x3ff059b0d3158574
    1.0442737824274138,  // 0x3ff0b5586cf9890f
    1.0671404006768237,  // 0x3ff11301d0125b51
    1.0905077326652577,  // 0x3ff172b83c7d517b
    1.1143867425958924,  // 0x3ff1d4873168b9aa
    1.1387886347566916,  // 0x3ff2387a6e756238
    1.1637248587775775,  // 0x3ff29e9df51fdee1
    1.189207115002721,  // 0x3ff306fe0a31b715
    1.215247359980469,  // 0x3ff371a7373aa9cb
    1.241857812073484,  // 0x3ff3dea64c123422
    1.2690509571917332,  // 0x3ff44e086061892d
    1.2968395546510096,  // 0x3ff4bfdad5362a27
    1.3252366431597413,  // 0x3ff5342b569d4f82
    1.3542555469368927,  // 0x3ff5ab07dd485429
    1.383909881963832,  // 0x3ff6247eb03a5585
    1.4142135623730951,  // 0x3ff6a09e667f3bcd
    1.4451808069770467,  // 0x3ff71f75e8ec5f74
    1.4768261459394993,  // 0x3ff7a11473eb0187
    1.5091644275934228,  // 0x3ff82589994cce13
    1.5422108254079407,  // 0x3ff8ace5422aa0db
    1.5759808451078865,  // 0x3ff93737b0cdc5e5
    1.6104903319492543,  // 0x3ff9c49182a3f090
    1.645755478153965,  // 0x3ffa5503b23e255d
    1.681792830507429,  // 0x3ffae89f995ad3ad
    1.718619298122478,  // 0x3ffb7f76f2fb5e47
    1.7562521603732995,  // 0x3ffc199bdd85529c
    1.7947090750031072,  // 0x3ffcb720dcef9069
    1.8340080864093424,  // 0x3ffd5818dcfba487
    1.8741676341103,  // 0x3ffdfc97337b9b5f
    1.9152065613971474,  // 0x3ffea4afa2a490da
    1.9571441241754002,  // 0x3fff50765b6e4540
};
template <int N>
__device__ __forceinline__ void exp_bounded_tab(const double (&x)[N], double (&out)[N]) {
    double t[N], r[N], p[N], s[N];
    int n[N];
#pragma unroll
    for (int i = 0; i < N; ++i) t[i] = fma(x[i], kExpTabRed[0], kExpRed[1]);
#pragma unroll
    for (int i = 0; i < N; ++i) {
        n[i] = __double2loint(t[i]);
        s[i] = __ldg(&kExp2Tab[n[i] & 31]);
        t[i] -= kExpRed[1];
    }
#pragma unroll
    for (int i = 0; i < N; ++i) r[i] = fma(t[i], kExpTabRed[1], x[i]);
#pragma unroll
    for (int i = 0; i < N; ++i) r[i] = fma(t[i], kExpTabRed[2], r[i]);
#pragma unroll
    for (int i = 0; i < N; ++i) p[i] = fma(r[i], kExpTabPoly[3], kExpTabPoly[2]);
#pragma unroll
    for (int i = 0; i < N; ++i) p[i] = fma(p[i], r[i], kExpTabPoly[1]);
#pragma unroll
    for (int i = 0; i < N; ++i) p[i] = fma(p[i], r[i], kExpTabPoly[0]);
#pragma unroll
    for (int i = 0; i < N; ++i) p[i] = fma(p[i], r[i], 1.0);
#pragma unroll
    for (int i = 0; i < N; ++i) p[i] = fma(p[i], r[i], 1.0);
#pragma unroll
    for (int i = 0; i < N; ++i) p[i] *= s[i];
#pragma unroll
    for (int i = 0; i < N; ++i)
        out[i] = __hiloint2double(__double2hiint(p[i]) + ((n[i] >> 5) << 20), __double2loint(p[i]));
}
#ifndef BINEST_EXP_TAB
#define BINEST_EXP_TAB 1
#endif

template <int F, int K>
struct OpLogistic {
    static constexpr int D = (K - 1) * (F + 1), NCOL = (F + 2) & ~1, TW_MAX = 2, SLOTS = (K - 1) * (F + 12) + 6;
    struct Coef { int unused; };
    // safe: every logit of this walker is bounded by 170 on ALL rows — |z_k| <= |b_k| + Sum_f |w_kf| max_i |x_if| with the
    // column maxima fixed at upload (OpCst::m[f]) — so the per-row range checks of the fast exp path can be skipped
    // (12 integer instructions per row and lane in a kernel that is bound by issue slots, not by the fp64 pipe)
    struct Row { double w[K - 1][F + 1]; int safe; };
    __device__ __forceinline__ static Row make_row(const double (&th)[D], const OpCst &cst) {
        Row c;
        bool safe = true;
#pragma unroll
        for (int k = 0; k < K - 1; ++k) {
            double bound = fabs(th[k * (F + 1) + F]);
#pragma unroll
            for (int f = 0; f <= F; ++f) {
                c.w[k][f] = th[k * (F + 1) + f];
                if (f < F) bound = fma(fabs(c.w[k][f]), cst.m[f], bound);
            }
            safe = safe && (bound < 169.0);  // NaN compares false
        }
        c.safe = safe ? 1 : 0;
        return c;
    }
    __device__ static Coef prepare(const double (&)[D], bool &ok, const OpCst &) {
        ok = true;
        return Coef{0};
    }
    // Per datum  log p_y = z_y - log(1 + Sum_k exp z_k)  (z_K = 0).  The exps do not depend on the label, and z_y is
    // picked with 0/1 indicators formed once per row (i_k = [label == k]), so the per-walker code has no label-dependent
    // branch or select at all (an earlier form shifted by z_y first — log p_y = -log(1 + Sum_{k != y} exp(z_k - z_y)) —
    // and paid ~25 integer/branch instructions per walker and datum for the permutation; this kernel is bound by issue
    // slots and fixed latencies, not by the fp64 pipe: profiles/r01g_ncu_loglike_c3.md).
    // Sum_i log p_i = log Prod_i e_i - log Prod_i s_i with e_i = exp z_y (one of the exps already computed, or 1 for
    // the reference class) and s_i = 1 + Sum_k exp z_k.  Both running products are renormalised to [1, 2) every
    // RENORM = 4 rows (renorm(): the exponent fields move into ONE shared integer, e2 += expo(e) - expo(s); ~3 ALU
    // instructions per datum instead of a log), two logs at the end.  With |z_k| < 170 on the fast path four factors
    // cannot leave the normal range: prod < 2 (1 + 2 e^170)^4 = 7e296, prode in (e^-680, 2 e^680).
    // Nothing of magnitude |z| is ever accumulated, so the absolute error of the sum is ~sqrt(rows) ulp of 1 whatever
    // the logits are — also for well-separated classes, where Sum z_y and Sum log s are both ~N |z| and their
    // difference is tiny (an earlier form accumulated Sum z_y and the log of the s-product separately and lost 6e-8
    // absolute at N = 1e6, |z| = 600; with perfectly separated data both products now go through identical arithmetic
    // and the result is exactly 0, as the reference's Log[e^z_y / Sum e^z] gives).  |z_k| >= 170 (NaN included) takes
    // the slow path: max-shifted libdevice exp/log, accumulated in `lin`.
    static constexpr int RENORM = 4;
    struct Acc { double lin, prod, prode; long long e2; };
    __device__ __forceinline__ static Acc acc_init() { return Acc{0.0, 1.0, 1.0, 0}; }
    __device__ __forceinline__ static double acc_value(const Acc &a) {
        return a.lin + fma((double)a.e2, 0.693147180559945309417232, log(a.prode) - log(a.prod));
    }
    template <int TW>
    __device__ __forceinline__ static void rows(const Row (&c)[TW], const double *__restrict__ r, Acc (&acc)[TW]) {
        static_assert(K == 2 || K == 3, "softmax operator is specialised for 2 or 3 classes");
        constexpr int E = K - 1;  // exps per datum
        double z[TW * E], ex[TW * E];
        const double labv = r[F];
        const double i0 = (labv == 0.0) ? 1.0 : 0.0;
        const double i1 = (K == 3 && labv == 1.0) ? 1.0 : 0.0;  // the last class is the reference class (z = 0)
        const double iK = 1.0 - i0 - i1;
#pragma unroll
        for (int k = 0; k < E; ++k)
#pragma unroll
            for (int u = 0; u < TW; ++u) z[u * E + k] = c[u].w[k][F];
#pragma unroll
        for (int f = 0; f < F; ++f) {
            const double x = r[f];
#pragma unroll
            for (int k = 0; k < E; ++k)
#pragma unroll
                for (int u = 0; u < TW; ++u) z[u * E + k] = fma(c[u].w[k][f], x, z[u * E + k]);
        }
        int safe = 1;
#pragma unroll
        for (int u = 0; u < TW; ++u) safe &= c[u].safe;
        bool fast = true;
        if (!safe) {
#pragma unroll
            for (int u = 0; u < TW; ++u)
#pragma unroll
                for (int k = 0; k < E; ++k) fast = fast && exp_arg_bounded_170(z[u * E + k]);
        }
        if (__builtin_expect(fast, 1)) {
#if BINEST_EXP_TAB
            exp_bounded_tab<TW * E>(z, ex);
#else
            exp_bounded<TW * E>(z, ex);
#endif
#pragma unroll
            for (int u = 0; u < TW; ++u) {
                double s = 1.0 + ex[u * E];
                double ey = fma(i0, ex[u * E], iK);
                if (K == 3) { s += ex[u * E + 1]; ey = fma(i1, ex[u * E + 1], ey); }
                acc[u].prod *= s;
                acc[u].prode *= ey;
            }
        } else {
#pragma unroll
            for (int u = 0; u < TW; ++u) {
                double zy = i0 * z[u * E];
                if (K == 3) zy = fma(i1, z[u * E + 1], zy);
                double mx = fmax(z[u * E], 0.0);
                if (K == 3) mx = fmax(mx, z[u * E + 1]);
                double s = exp(-mx) + exp(z[u * E] - mx);
                if (K == 3) s += exp(z[u * E + 1] - mx);
                acc[u].lin += zy - (mx + log(s));
            }
        }
    }
    // move the exponents of both running products into e2 (call at least every RENORM rows)
    template <int TW>
    __device__ __forceinline__ static void renorm(Acc (&acc)[TW]) {
#pragma unroll
        for (int u = 0; u < TW; ++u) {
            const int hp = __double2hiint(acc[u].prod), ep = (hp >> 20) - 1023;
            const int hq = __double2hiint(acc[u].prode), eq = (hq >> 20) - 1023;
            acc[u].e2 += eq - ep;
            acc[u].prod = __hiloint2double(hp - (ep << 20), __double2loint(acc[u].prod));
            acc[u].prode = __hiloint2double(hq - (eq << 20), __double2loint(acc[u].prode));
        }
    }
    __device__ static double local(const Coef &, double acc, double, const OpCst &) { return acc; }
    __device__ static double finish_total(const Coef &, double acc, double, const OpCst &) { return acc; }
};

// ------------------------------------------------------------------ GeometricBrownianMotionProcess[mu, sigma, x0]
// device row i = (a_i, b_i) = (r_i / sqrt(dt_i), sqrt(dt_i)), r_i = log(x_i / x_{i-1});
// cst = Sum_i(-log x_i - 1/2 log dt_i) - rows * 1/2 log 2pi  (parameter independent, fixed at upload)
struct OpGbm {
    static constexpr int D = 2, NCOL = 2, TW_MAX = 8, RENORM = 0, SLOTS = 2;
    struct Coef { double h, lognorm; };
    struct Row { double negm; };
    using Acc = double;
    __device__ __forceinline__ static Acc acc_init() { return 0.0; }
    __device__ __forceinline__ static double acc_value(const Acc &a) { return a; }
    __device__ __forceinline__ static Row make_row(const double (&th)[D], const OpCst &) {
        return Row{-(th[0] - 0.5 * th[1] * th[1])};
    }
    __device__ static Coef prepare(const double (&th)[D], bool &ok, const OpCst &) {
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
    __device__ static double local(const Coef &, double acc, double, const OpCst &) { return acc; }
    __device__ static double finish_total(const Coef &c, double acc, double rows, const OpCst &cst) {
        return rows * c.lognorm + cst.c - c.h * acc;
    }
};

// ------------------------------------------------------------------ GP marginal likelihood (gp.cu)
// Not a streaming reduction: gp.cu writes the finished logL where the walk expects the (single) partial sum,
// so only the epilogue interface is needed here.
struct OpGpSe {
    static constexpr int D = 3;
    struct Coef { int unused; };
    __device__ static Coef prepare(const double (&th)[D], bool &ok, const OpCst &) {
        ok = th[0] > 0.0 && th[1] > 0.0 && th[2] > 0.0;
        return Coef{0};
    }
    __device__ static double local(const Coef &, double acc, double, const OpCst &) { return acc; }
    __device__ static double finish_total(const Coef &, double acc, double, const OpCst &) { return acc; }
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
