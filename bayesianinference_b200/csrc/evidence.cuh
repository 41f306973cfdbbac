// evidence.cuh — logX / trapezoid / logsumexp evidence reductions on whole sample lists:
//   crude weights (calculateXValues BS:785-799, trapezoidWeigths["Log"] BS:756-771, calculateWeightsCrude
//   BS:812-831, logSumExp BU:318-335, calculateEntropy BS:801-810) and the Monte-Carlo error estimate
//   evidenceSampling (BS:1158-1291).
#pragma once
#include "common.cuh"

namespace binest {

// inclusive scan over the block (blockDim.x multiple of 32, <= 1024); scratch: 34 doubles.
// Returns the inclusive prefix for this thread; *total receives the block total (all threads).
__device__ __forceinline__ double block_scan_incl(double v, double *scratch, double *total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    __syncthreads();
    if (lane == 31) scratch[w] = v;
    __syncthreads();
    if (w == 0) {
        double t = lane < nw ? scratch[lane] : 0.0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double u = __shfl_up_sync(0xffffffffu, t, o);
            if (lane >= o) t += u;
        }
        scratch[lane] = t;
    }
    __syncthreads();
    const double off = w > 0 ? scratch[w - 1] : 0.0;
    *total = scratch[nw - 1];
    return v + off;
}

// trapezoid log-weight of element k (0-based) of a length-M logX sequence read through f(k)  (BS:756-771)
template <class F>
__device__ __forceinline__ double trapezoid_logw(long long k, long long M, F lx) {
    const double log_half = -0.69314718055994530941723212145818;
    if (M == 1) return 0.0;
    if (k == M - 1) return log_half + log_add(lx(M - 2), lx(M - 1));
    const double left = (k == 0) ? log_subtract(0.69314718055994530941723212145818, lx(0)) : lx(k - 1);
    return log_half + log_subtract(left, lx(k + 1));
}

// ---- crude weights of one sorted list (single CTA, grid-stride over M) ------------------------------
// logX: deleted k: -cumsum 1/pool_k ; live i = n..1: log i - log(n+1) + logX_last  (BS:785-799 generalised
// to per-sample pool sizes).  If logX_in != nullptr the deleted part is taken from it (engine state).
static __global__ void __launch_bounds__(1024)
crude_weights_kernel(long long M, long long n_live, const double *__restrict__ logL, const int *__restrict__ pool32,
                     const long long *__restrict__ pool64, double *__restrict__ logX, double *__restrict__ crude_logw,
                     double *__restrict__ summary /*[4]*/) {
    __shared__ double scratch[100];
    const int tid = threadIdx.x, nt = blockDim.x;
    const long long nd = M - n_live;
    double carry = 0.0;
    for (long long base = 0; base < nd; base += nt) {
        const long long k = base + tid;
        double v = 0.0;
        if (k < nd) v = -1.0 / (double)(pool32 ? (long long)pool32[k] : pool64[k]);
        double tot;
        const double inc = block_scan_incl(v, scratch, &tot);
        if (k < nd) logX[k] = carry + inc;
        carry += tot;
        __syncthreads();
    }
    const double log_np1 = log((double)n_live + 1.0);
    for (long long j = tid; j < n_live; j += nt) logX[nd + j] = (log((double)(n_live - j)) - log_np1) + carry;
    __syncthreads();
    LseAcc acc = lse_empty();
    double lmax = -CUDART_INF;
    for (long long k = tid; k < M; k += nt) {
        const double lw = trapezoid_logw(k, M, [&](long long i) { return logX[i]; }) + logL[k];
        crude_logw[k] = lw;
        acc = lse_merge(acc, lse_term(lw, logL[k]));
        lmax = fmax(lmax, logL[k]);
    }
    acc = block_lse(acc, scratch);
    lmax = block_max(lmax, scratch);
    if (tid == 0) {
        const double logZ = acc.m + log(acc.s0);
        summary[0] = logZ;                       // CrudeLogEvidence        BS:1186
        summary[1] = acc.s1 / acc.s0 - logZ;     // CrudeRelativeEntropy    BS:1192
        summary[2] = lmax;                       // LogLikelihoodMaximum    BS:1187
        summary[3] = logX[M - 1] + lmax;         // LogEstimatedMissingEvidence BS:1188-1191 (log domain)
    }
}

// ---- evidenceSampling: one CTA per Monte-Carlo draw r ---------------------------------------------------
// deleted: logX_k = -cumsum Exp(rate pool_k)               (BS:1217-1224; pool_k == n in the reference)
// live:    logX   = logX_last - (order statistics of n i.i.d. Exp(1))   (BS:1209-1215).  The sorted draws
//          are generated directly through the Renyi representation e_(j) = Sum_{i<=j} E_i / (n - i),
//          which has exactly the law of Sort[RandomVariate[ExponentialDistribution[1], n]].
static __global__ void __launch_bounds__(1024)
evidence_sampling_kernel(long long M, int d, long long n_live, const double *__restrict__ points,
                         const double *__restrict__ logL, const long long *__restrict__ pool, unsigned long long seed,
                         double *__restrict__ slx /*[R][M]*/, double *__restrict__ lw /*[R][M]*/,
                         double *__restrict__ z /*[R]*/, double *__restrict__ pmean /*[R][d]*/,
                         double *__restrict__ H /*[R]*/) {
    __shared__ double scratch[100];
    const int tid = threadIdx.x, nt = blockDim.x;
    const unsigned r = blockIdx.x;
    const long long nd = M - n_live;
    double *lx = slx + (size_t)r * M;
    double *lwr = lw + (size_t)r * M;
    double carry = 0.0;
    for (long long base = 0; base < nd; base += nt) {
        const long long k = base + tid;
        double v = 0.0;
        if (k < nd) {
            double u0, u1;
            rng_uniform2(seed, 0u, r, (uint32_t)k, TAG_EV_DEAD, 0u, u0, u1);
            v = log(u0) / (double)pool[k];  // = -Exp(1)/pool
        }
        double tot;
        const double inc = block_scan_incl(v, scratch, &tot);
        if (k < nd) lx[k] = carry + inc;
        carry += tot;
        __syncthreads();
    }
    const double lxD = carry;
    carry = 0.0;
    for (long long base = 0; base < n_live; base += nt) {
        const long long j = base + tid;
        double v = 0.0;
        if (j < n_live) {
            double u0, u1;
            rng_uniform2(seed, 0u, r, (uint32_t)j, TAG_EV_LIVE, 0u, u0, u1);
            v = log(u0) / (double)(n_live - j);
        }
        double tot;
        const double inc = block_scan_incl(v, scratch, &tot);
        if (j < n_live) lx[nd + j] = lxD + (carry + inc);
        carry += tot;
        __syncthreads();
    }
    __syncthreads();
    LseAcc acc = lse_empty();
    for (long long k = tid; k < M; k += nt) {
        const double t = trapezoid_logw(k, M, [&](long long i) { return lx[i]; }) + logL[k];
        lwr[k] = t;
        acc = lse_merge(acc, lse_term(t, logL[k]));
    }
    acc = block_lse(acc, scratch);
    const double zr = acc.m + log(acc.s0);  // BS:1228
    if (tid == 0) {
        z[r] = zr;
        if (H) H[r] = acc.s1 / acc.s0 - zr;  // BS:1263-1268
    }
    // posterior weights exp(lw - z) and parameter means (BS:1229-1235)
    for (int a = 0; a < d; ++a) {
        double s = 0.0;
        for (long long k = tid; k < M; k += nt) s += exp(lwr[k] - zr) * points[k * d + a];
        s = block_sum(s, scratch);
        if (tid == 0 && pmean) pmean[(size_t)r * d + a] = s;
    }
}

// per-sample mean / sd over the draws of LogPosteriorWeight = lw - z and SampledLogX (BS:1244-1250, 1138-1149)
static __global__ void evidence_moments_kernel(long long M, int R, const double *__restrict__ slx,
                                               const double *__restrict__ lw, const double *__restrict__ z,
                                               double *__restrict__ logw_mean, double *__restrict__ logw_sd,
                                               double *__restrict__ slx_mean, double *__restrict__ slx_sd) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= M) return;
    double s1 = 0, s2 = 0, x1 = 0, x2 = 0;
    for (int r = 0; r < R; ++r) {
        const double a = lw[(size_t)r * M + k] - z[r], b = slx[(size_t)r * M + k];
        s1 += a; s2 += a * a; x1 += b; x2 += b * b;
    }
    const double Rd = (double)R, m1 = s1 / Rd, mx = x1 / Rd;
    const double v1 = (s2 - Rd * m1 * m1) / (Rd - 1.0), vx = (x2 - Rd * mx * mx) / (Rd - 1.0);
    if (logw_mean) logw_mean[k] = m1;
    if (logw_sd) logw_sd[k] = v1 > 0 ? sqrt(v1) : 0.0;
    if (slx_mean) slx_mean[k] = mx;
    if (slx_sd) slx_sd[k] = vx > 0 ? sqrt(vx) : 0.0;
}

}  // namespace binest
