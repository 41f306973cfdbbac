// loglike.cuh — the batched log-likelihood reduction  logL(theta_w) = Sum_i logpdf(theta_w; row_i)
// for P parameter vectors over N data rows (BS:492 / BS:581 evaluated for a whole batch of walkers).
//
// Mapping (B200).  A CTA owns 32*TW walkers and a contiguous slice of the data rows:
//   * lane l holds the derived coefficients of TW walkers in registers (walkers l, l+32, ...);
//   * the slice is streamed through shared memory in 16 KiB tiles by the TMA engine
//     (cp.async.bulk -> SASS UBLKCP, completion on an mbarrier, double buffered);
//   * the 8 warps split the rows of each tile; a row is read with one broadcast LDS.128 and
//     feeds TW DFMA sequences, so the shared-memory return path (128 B/clk/SM = one LDS.128 per
//     4 clk per warp) stays below the fp64 pipe (first version, TW = 1, was LDS-bound at 65 %);
//   * per-warp sums are combined across warps in shared memory in a fixed order and written to
//     partials[walker][cta]; the consumer (finalize kernel / walk step) sums them in a fixed order.
// The fp64 pipe — not HBM/L2 — is the bound when P >= ~11 walkers share a tile.
#pragma once
#include "operators.cuh"
#include "xchg.cuh"

namespace binest {

constexpr int kTileBytes = 16384;
constexpr int kStages = 2;
constexpr int kWarps = 8;

template <class OP>
__host__ __device__ constexpr int tile_rows() {
    return (kTileBytes / (8 * OP::NCOL)) & ~1;
}

// The row loop of every likelihood kernel: rows first, first + stride, ... < nr of a shared-memory tile, one row at a
// time for all TW walkers of the lane — the row operands are shared by TW consecutive DFMAs (register-reuse cache),
// which keeps each DFMA at two distinct register reads; three distinct 64-bit sources issue at 2/3 rate on sm_100
// (scripts/dfma_patterns.cu).  (A register queue running the LDS two rows ahead of the arithmetic was measured slower:
// 88.7 -> 94.2 us.)  Operators with running products (RENORM > 0) are renormalised every RENORM rows.
template <class OP, int TW>
__device__ __forceinline__ void sweep_rows(const typename OP::Row (&c)[TW], const double *__restrict__ tile, int first,
                                           int stride, int nr, typename OP::Acc (&acc)[TW]) {
    constexpr int NCOL = OP::NCOL;
    if constexpr (OP::RENORM > 1) {
        int i = first;
        for (; i + (OP::RENORM - 1) * stride < nr; i += OP::RENORM * stride) {
#pragma unroll
            for (int q = 0; q < OP::RENORM; ++q) OP::template rows<TW>(c, tile + (size_t)(i + q * stride) * NCOL, acc);
            OP::template renorm<TW>(acc);
        }
        for (; i < nr; i += stride) {
            OP::template rows<TW>(c, tile + (size_t)i * NCOL, acc);
            OP::template renorm<TW>(acc);
        }
    } else {
#pragma unroll 2
        for (int i = first; i < nr; i += stride) OP::template rows<TW>(c, tile + (size_t)i * NCOL, acc);
    }
}

template <class OP, int TW>
__global__ void __launch_bounds__(kWarps * 32)
loglike_stream_kernel(const double *__restrict__ data, long long rows, long long rows_per_cta,
                      const double *__restrict__ theta /* SoA [D][Ps] */, int P, int Ps,
                      double *__restrict__ partials /* [Ps][Gs] */, int Gs, const OpCst cst) {
    constexpr int TR = tile_rows<OP>();
    constexpr int NCOL = OP::NCOL;
    __shared__ __align__(128) double tiles[kStages][TR * NCOL];
    __shared__ uint64_t full[kStages];

    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int wbase = blockIdx.y * (32 * TW);

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

    // PDL: everything above only reads the immutable data set, so it overlaps the predecessor (walk_step);
    // theta and the partials buffer belong to the predecessor / successor and are touched after the wait.
    pdl_launch_dependents();
    pdl_wait();

    // per-datum coefficients of the lane's TW walkers (loaded while the first tiles are in flight)
    typename OP::Row c[TW];
#pragma unroll
    for (int t = 0; t < TW; ++t) {
        const int w = wbase + lane + 32 * t;
        double th[OP::D];
#pragma unroll
        // ld.global.cg, not the read-only (.nc) path `const __restrict__` would select: under PDL this kernel is
        // already resident while its predecessor (walk_step) writes theta, and the non-coherent cache of the SM is
        // not invalidated by griddepcontrol.wait (seen as stale proposals at N = 1e6: stored logL != logL(point))
        for (int j = 0; j < OP::D; ++j) th[j] = (w < P) ? __ldcg(theta + (size_t)j * Ps + w) : 1.0;
        c[t] = OP::make_row(th, cst);
    }

    typename OP::Acc acc[TW];
#pragma unroll
    for (int t = 0; t < TW; ++t) acc[t] = OP::acc_init();

    for (int t = 0; t < ntiles; ++t) {
        const int s = t % kStages;
        mbar_wait(&full[s], (uint32_t)((t / kStages) & 1));
        const double *__restrict__ tile = &tiles[s][0];
        const long long a = r0 + (long long)t * TR;
        const int nr = (int)((r1 - a < TR) ? (r1 - a) : TR);
        sweep_rows<OP, TW>(c, tile, wid, kWarps, nr, acc);
        __syncthreads();  // everyone is done with stage s before the TMA engine refills it
        if (threadIdx.x == 0 && t + kStages < ntiles) issue(t + kStages);
    }

    // fixed-order cross-warp combine (reuses the tile buffer), then one partial per (walker, CTA)
    static_assert(kStages * TR * NCOL >= kWarps * 32 * TW, "tile buffer too small for the cross-warp combine");
    double *red = &tiles[0][0];  // [kWarps][32*TW]
    __syncthreads();
#pragma unroll
    for (int u = 0; u < TW; ++u) red[wid * (32 * TW) + lane + 32 * u] = OP::acc_value(acc[u]);
    __syncthreads();
    for (int k = threadIdx.x; k < 32 * TW; k += blockDim.x) {
        double sum = 0.0;
#pragma unroll
        for (int q = 0; q < kWarps; ++q) sum += red[q * (32 * TW) + k];
        const int w = wbase + k;
        if (w < Ps) partials[(size_t)w * Gs + blockIdx.x] = sum;
    }
}

// where the G partial sums of walker w live: partial (w, g) at p[w * sw + g * sg].
//   single GPU:   the per-CTA partials of loglike_stream_kernel, [Ps][Gs]            -> sw = Gs, sg = 1
//   data-sharded: the per-rank sums after the exchange (comm.cu), [world][Ps]        -> sw = 1,  sg = Ps,
//                 localized = 1: every rank has already applied OP::local with its own data constants
struct PartialView {
    const double *p;
    int G;
    long long sw, sg;
    int localized = 0;
};

// Consumer side of the in-kernel sharded exchange (xchg.cuh): the first thread of the CTA waits for the flags of the
// exchange the preceding producer kernel published, then every thread points the view at this rank's receive slot.
// xd == nullptr: nothing to do (single GPU, or the NCCL fallback whose view already points at the gathered buffer).
__device__ __forceinline__ bool resolve_exchange(PartialView &pv, const XchgDev *xd) {
    if (xd == nullptr) return true;
    __shared__ int s_ok;
    const unsigned t = xd->st->step;
    if (threadIdx.x == 0) s_ok = xchg_wait(*xd, t) ? 1 : 0;
    __syncthreads();
    pv.p = xd->buf[xd->rank] + xd->slot(t, 0);
    return s_ok != 0;
}

// fixed-order combine of the partials of one walker by one warp (all lanes get the sum)
__device__ __forceinline__ double combine_partials_warp(const PartialView &pv, int w, int lane) {
    double s = 0.0;
    // ld.global.cg: the persistent walk kernel re-reads these addresses every step while other SMs rewrite them
    // batches of 12 independent loads (one L2 round trip per batch instead of per element); same add order
    const double *base = pv.p + (size_t)w * pv.sw;
    for (int g0 = lane; g0 < pv.G; g0 += 32 * 12) {
        double v[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) v[k] = (g0 + 32 * k < pv.G) ? __ldcg(base + (size_t)(g0 + 32 * k) * pv.sg) : 0.0;
#pragma unroll
        for (int k = 0; k < 12; ++k) s += v[k];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    return s;
}

// operator epilogue + operator constraints (RuntimeErrorHandler -> logzero, BS:500-503)
template <class OP>
__device__ __forceinline__ double loglike_finish(const double (&th)[OP::D], double sum, double rows, const OpCst &cst,
                                                 double logzero, int localized = 0) {
    bool ok;
    const typename OP::Coef c = OP::prepare(th, ok, cst);
    const double v = localized ? OP::finish_total(c, sum, rows, cst) : op_finish<OP>(c, sum, rows, cst);
    return (ok && isfinite(v)) ? v : logzero;
}

// one warp per walker
template <class OP>
__global__ void loglike_finalize_kernel(const double *__restrict__ theta, int P, int Ps,
                                        PartialView pv, double rows, const OpCst cst,
                                        const __grid_constant__ PriorSpec prior, double logzero,
                                        double *__restrict__ out, const XchgDev *xd = nullptr) {
    const bool xok = resolve_exchange(pv, xd);
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= P || !xok) return;
    const double s = combine_partials_warp(pv, w, lane);
    double th[OP::D];
#pragma unroll
    for (int j = 0; j < OP::D; ++j) th[j] = theta[(size_t)j * Ps + w];
    double v = loglike_finish<OP>(th, s, rows, cst, logzero, pv.localized);
    if (!in_box<OP::D>(prior, th)) v = logzero;  // If[constraints[theta], Sum[...], logzero] BS:491-494
    if (lane == 0) out[w] = v;
}

// data-sharded mode: this rank's sum over its own slices in shard-additive form (OP::local with this rank's data
// constants), one value per walker — the payload of the exchange
template <class OP>
__global__ void shard_reduce_kernel(const PartialView pv, const double *theta /* SoA [D][Ps] */, int P, int Ps,
                                    double rows_local, const OpCst cst_local, double *send) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= Ps) return;
    double s = combine_partials_warp(pv, w, lane);
    if (w < P) {
        double th[OP::D];
#pragma unroll
        for (int j = 0; j < OP::D; ++j) th[j] = __ldcg(theta + (size_t)j * Ps + w);
        bool ok;
        const typename OP::Coef c = OP::prepare(th, ok, cst_local);
        s = OP::local(c, s, rows_local, cst_local);
    }
    if (lane == 0) send[w] = s;
}

// the same, pushing the value into every rank's receive buffer (in-kernel exchange, xchg.cuh)
template <class OP>
__global__ void shard_reduce_push_kernel(const PartialView pv, const double *theta /* SoA [D][Ps] */, int P, int Ps,
                                         double rows_local, const OpCst cst_local, const XchgDev x) {
    const unsigned t = x.st->step + 1u;
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w < Ps) {
        double s = combine_partials_warp(pv, w, lane);
        if (w < P) {
            double th[OP::D];
#pragma unroll
            for (int j = 0; j < OP::D; ++j) th[j] = __ldcg(theta + (size_t)j * Ps + w);
            bool ok;
            const typename OP::Coef c = OP::prepare(th, ok, cst_local);
            s = OP::local(c, s, rows_local, cst_local);
        }
        if (lane == 0) xchg_store(x, t, (size_t)w, s);
    }
    xchg_publish(x, t);
}

// log prior density for a batch (BS:410-426); theta SoA [d][Ps]
static __global__ void logprior_kernel(const double *__restrict__ theta, int P, int Ps,
                                       const __grid_constant__ PriorSpec prior, double logzero,
                                       double *__restrict__ out) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= P) return;
    double th[BINEST_MAXD];
    for (int j = 0; j < prior.d; ++j) th[j] = theta[(size_t)j * Ps + w];
    out[w] = logprior_dyn(prior, th, logzero);
}

// generateStartingPoints (BS:1055-1068): i.i.d. prior draws by inverse CDF (rejection from the parent
// normal for the truncated case); out row-major [n][d].  Same Philox addressing as the oracle.
static __global__ void sample_prior_kernel(const __grid_constant__ PriorSpec prior, long long n,
                                           unsigned long long seed, unsigned run_id, double *__restrict__ out) {
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
