// loglike.cuh — the batched log-likelihood reduction  logL(theta_w) = Sum_i logpdf(theta_w; row_i)
// for P parameter vectors over N data rows (BS:492 / BS:581 evaluated for a whole batch of walkers).
//
// Mapping (B200): lane = walker.  A CTA of up to 8 warps owns up to 256 parameter vectors (their
// derived coefficients live in registers) and a contiguous slice of the data.  The slice is streamed
// through shared memory in 16 KiB tiles by the TMA engine (cp.async.bulk -> SASS UBLKCP, completion on
// an mbarrier, double buffered); every warp reads each row with one broadcast LDS and spends the
// operator's DFMA sequence on it, so the fp64 pipe — not HBM/L2 — is the bound.
// Partial sums go to partials[cta][walker] and are combined in a fixed order by the consumer
// (finalize kernel or the walk's accept kernel), so results are reproducible run to run.
#pragma once
#include "operators.cuh"

namespace binest {

constexpr int kTileBytes = 16384;
constexpr int kStages = 2;
constexpr int kMaxWarps = 8;

template <class OP>
__host__ __device__ constexpr int tile_rows() {
    return (kTileBytes / (8 * OP::NCOL)) & ~1;
}

template <class OP>
__global__ void __launch_bounds__(kMaxWarps * 32)
loglike_stream_kernel(const double *__restrict__ data, long long rows, long long rows_per_cta,
                      const double *__restrict__ theta /* SoA [D][Ps] */, int P, int Ps,
                      double *__restrict__ partials /* [gridDim.x][Ps] */) {
    constexpr int TR = tile_rows<OP>();
    constexpr int NCOL = OP::NCOL;
    __shared__ __align__(128) double tiles[kStages][TR * NCOL];
    __shared__ uint64_t full[kStages];

    const int w = blockIdx.y * blockDim.x + threadIdx.x;
    double th[OP::D];
#pragma unroll
    for (int j = 0; j < OP::D; ++j) th[j] = (w < P) ? theta[(size_t)j * Ps + w] : 1.0;
    bool ok;
    const typename OP::Coef c = OP::prepare(th, ok);

    const long long r0 = (long long)blockIdx.x * rows_per_cta;
    const long long r1 = (r0 + rows_per_cta < rows) ? r0 + rows_per_cta : rows;
    const int ntiles = (r1 > r0) ? (int)((r1 - r0 + TR - 1) / TR) : 0;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    auto issue = [&](int t) {
        const long long a = r0 + (long long)t * TR;
        const int nr = (int)((r1 - a < TR) ? (r1 - a) : TR);
        const uint32_t bytes = ((uint32_t)(nr * NCOL * 8) + 15u) & ~15u;  // buffer is padded at upload
        const int s = t % kStages;
        mbar_expect_tx(&full[s], bytes);
        bulk_g2s(&tiles[s][0], data + a * NCOL, bytes, &full[s]);
    };
    if (threadIdx.x == 0)
        for (int t = 0; t < kStages && t < ntiles; ++t) issue(t);

    double acc0 = 0.0, acc1 = 0.0;
    for (int t = 0; t < ntiles; ++t) {
        const int s = t % kStages;
        mbar_wait(&full[s], (uint32_t)((t / kStages) & 1));
        const double *__restrict__ tile = &tiles[s][0];
        const long long a = r0 + (long long)t * TR;
        const int nr = (int)((r1 - a < TR) ? (r1 - a) : TR);
        int i = 0;
#pragma unroll 4
        for (; i + 1 < nr; i += 2) {
            OP::row(c, tile + (size_t)i * NCOL, acc0);
            OP::row(c, tile + (size_t)(i + 1) * NCOL, acc1);
        }
        if (i < nr) OP::row(c, tile + (size_t)i * NCOL, acc0);
        __syncthreads();  // everyone is done with stage s before the TMA engine refills it
        if (threadIdx.x == 0 && t + kStages < ntiles) issue(t + kStages);
    }
    if (w < Ps) partials[(size_t)blockIdx.x * Ps + w] = acc0 + acc1;
}

// fixed-order combine of the per-CTA partials + operator epilogue + constraint guards
template <class OP>
__device__ __forceinline__ double loglike_combine(const double (&th)[OP::D], const double *__restrict__ partials,
                                                  int G, int Ps, int w, double rows, double cst, double logzero) {
    bool ok;
    const typename OP::Coef c = OP::prepare(th, ok);
    double s = 0.0;
    for (int g = 0; g < G; ++g) s += partials[(size_t)g * Ps + w];
    const double v = OP::finish(c, s, rows, cst);
    return (ok && isfinite(v)) ? v : logzero;  // RuntimeErrorHandler -> logzero, BS:500-503
}

template <class OP>
__global__ void loglike_finalize_kernel(const double *__restrict__ theta, int P, int Ps,
                                        const double *__restrict__ partials, int G, double rows, double cst,
                                        const __grid_constant__ PriorSpec prior, double logzero,
                                        double *__restrict__ out) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= P) return;
    double th[OP::D];
#pragma unroll
    for (int j = 0; j < OP::D; ++j) th[j] = theta[(size_t)j * Ps + w];
    double v = loglike_combine<OP>(th, partials, G, Ps, w, rows, cst, logzero);
    if (!in_box<OP::D>(prior, th)) v = logzero;  // If[constraints[theta], Sum[...], logzero] BS:491-494
    out[w] = v;
}

// log prior density for a batch (BS:410-426); theta SoA [d][Ps]
static __global__ void logprior_kernel(const double *__restrict__ theta, int P, int Ps, const __grid_constant__ PriorSpec prior,
                                double logzero, double *__restrict__ out) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= P) return;
    double th[BINEST_MAXD];
    for (int j = 0; j < prior.d; ++j) th[j] = theta[(size_t)j * Ps + w];
    out[w] = logprior_dyn(prior, th, logzero);
}

// generateStartingPoints (BS:1055-1068): i.i.d. prior draws by inverse CDF (rejection from the parent
// normal for the truncated case); out row-major [n][d].  Same Philox addressing as the oracle.
static __global__ void sample_prior_kernel(const __grid_constant__ PriorSpec prior, long long n, unsigned long long seed,
                                    unsigned run_id, double *__restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int j = 0; j < prior.d; ++j) {
        double u0, u1, v;
        rng_uniform2(seed, 0u, (uint32_t)j, (uint32_t)i, TAG_PRIOR, run_id, u0, u1);
        const double lo = prior.lo[j], hi = prior.hi[j];
        if (prior.kind[j] == BINEST_PRIOR_UNIFORM) v = lo + u0 * (hi - lo);
        else if (prior.kind[j] == BINEST_PRIOR_SCALE) v = lo * exp(u0 * log(hi / lo));
        else {
            v = 0.5 * (lo + hi);
            for (uint32_t blk = 1; blk < 100000u; ++blk) {
                double z0, z1;
                rng_normal2(seed, blk, (uint32_t)j, (uint32_t)i, TAG_PRIOR, run_id, z0, z1);
                const double a = prior.p0[j] + prior.p1[j] * z0, b = prior.p0[j] + prior.p1[j] * z1;
                if (a > lo && a < hi) { v = a; break; }
                if (b > lo && b < hi) { v = b; break; }
            }
        }
        out[i * prior.d + j] = v;
    }
}

// register-resident DFMA loop: the fp64 roofline denominator
static __global__ void __launch_bounds__(256) fp64_peak_kernel(double *out, int iters, double seed) {
    double a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3, a4 = seed + 4, a5 = seed + 5, a6 = seed + 6,
           a7 = seed + 7;
    const double m = 1.0000001, b = 1e-9 * threadIdx.x;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = fma(a0, m, b); a1 = fma(a1, m, b); a2 = fma(a2, m, b); a3 = fma(a3, m, b);
            a4 = fma(a4, m, b); a5 = fma(a5, m, b); a6 = fma(a6, m, b); a7 = fma(a7, m, b);
        }
    }
    const double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 12345.678) out[0] = s;
}

}  // namespace binest
