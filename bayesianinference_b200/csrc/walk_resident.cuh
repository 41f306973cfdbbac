// walk_resident.cuh — the whole S-step constrained walk in ONE launch for data sets that fit in the shared memories
// of a thread-block cluster (C4: 16 383 increments = 256 KB over 8 CTAs; small test problems in one CTA).
//
// For such data a likelihood launch per walk step is pure launch latency (~8 us fixed cost for ~1 us of fp64 work).
// Here a cluster of CS CTAs owns WP = 32 TW walkers for the whole walk:
//   * the data rows are sharded over the CS shared memories once per launch;
//   * every step each CTA reduces ITS shard for all WP proposals: the 16 warps split the rows, lane = walker, TW walkers
//     register-tiled per lane exactly as in loglike_stream_kernel (one broadcast LDS per row feeds TW DFMA chains — with
//     TW = 1 the kernel was bound by the shared-memory return path: a GBM row is 2 DFMAs per walker against one LDS.128;
//     r01: 0.33 of the fp64 peak on C4, profiles/r01b_resident_c4.md);
//   * the first TW warps double as CHAIN warps, one per set of 32 walkers (lane = walker): they combine the 16 warp sums
//     in a fixed order, exchange the CS per-CTA sums through distributed shared memory (st.async + mbarrier
//     complete_tx: no cluster-wide barrier on the per-step path; two buffers alternate by step parity, a peer can be at
//     most one step ahead), and — redundantly and deterministically in every CTA — apply the accept rule of nsDensity
//     (BS:602-617), the Haario recursion (BS:715-727) and form the next proposal;
//   * Philox normals / log u for the next CH steps are produced by all threads at once, off the per-step path.
// Same Philox addressing and arithmetic as walk_step_kernel, so results are identical to the stepped path and to the
// oracle.
#pragma once
#include <cooperative_groups.h>

#include "walk.cuh"

namespace binest {

namespace cg = cooperative_groups;

constexpr int kResWarpsMax = 16;  // NW = 16 or 8 warps per CTA: all sweep the data; warps 0..TW-1 also run the chains.  With 8
                                  // warps two CTAs (of different clusters) share an SM: one's chain phase runs under the
                                  // other's data phase
constexpr int kResMaxChunk = 16;  // walk steps of pre-generated increments held in shared memory (upper bound)

// dynamic shared memory layout (doubles), WP = 32 TW:
//   tile | xch[2][CS][WP] | red[NW][WP] | dz[CH][D][WP] | logu[CH][WP] | mean[D][WP] | cov[D*D][WP] | row[WP] (OP::Row)
template <class OP>
__host__ __device__ inline size_t resident_smem_doubles(long long rows_per_cta, int CS, int TW, int CH, int NW) {
    const size_t WP = 32 * (size_t)TW;
    const size_t tile = ((size_t)rows_per_cta * OP::NCOL + 1) & ~(size_t)1;
    const size_t rowsz = (sizeof(typename OP::Row) * WP + 7) / 8;
    return tile + 2 * (size_t)CS * WP + (size_t)NW * WP + (size_t)CH * OP::D * WP + (size_t)CH * WP + OP::D * WP +
           OP::D * OP::D * WP + rowsz;
}

template <class OP, int TW, int NW>
__global__ void __launch_bounds__(NW * 32, NW <= 8 ? 2 : 1)
walk_resident_kernel(const __grid_constant__ RunParams prm, RunArrays A, const __grid_constant__ PriorSpec prior,
                     const double *__restrict__ data, long long rows, long long rows_per_cta, const OpCst cst, int CS,
                     int CH) {
    constexpr int D = OP::D, NCOL = OP::NCOL, WP = 32 * TW;
    extern __shared__ __align__(16) double smem[];
    const size_t tile_sz = ((size_t)rows_per_cta * NCOL + 1) & ~(size_t)1;
    double *tile = smem;
    double *xch = tile + tile_sz;                   // [2][CS][WP]
    double *red = xch + 2 * CS * WP;                // [NW][WP]
    double *s_dz = red + NW * WP;                   // [CH][D][WP]
    double *s_logu = s_dz + (size_t)CH * D * WP;    // [CH][WP]
    double *s_mean = s_logu + (size_t)CH * WP;      // [D][WP]
    double *s_cov = s_mean + D * WP;                // [D*D][WP]
    typename OP::Row *s_row = reinterpret_cast<typename OP::Row *>(s_cov + D * D * WP);

    __shared__ uint64_t xbar[2];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (CS > 1) ? (int)cluster.block_rank() : 0;
    if (threadIdx.x == 0) {
        mbar_init(&xbar[0], 1);
        mbar_init(&xbar[1], 1);
        mbar_fence_init();
    }
    const int group = blockIdx.x / CS;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, tid = threadIdx.x;
    const int K = prm.K, P = prm.R * K;
    const bool chain = wid < TW;  // chain warp of walker set `wid`

    // ---- data shard of this CTA -> shared memory (once per launch)
    const long long r0 = (long long)rank * rows_per_cta;
    const long long r1 = (r0 + rows_per_cta < rows) ? r0 + rows_per_cta : rows;
    const int nr = r1 > r0 ? (int)(r1 - r0) : 0;
    for (int e = tid; e < nr * NCOL; e += blockDim.x) tile[e] = data[r0 * NCOL + e];

    // ---- walker of this lane in a chain warp (the same in every CTA of the cluster)
    const int wl = (chain ? wid : 0) * 32 + lane;  // walker slot inside the group (chain warps only)
    const int w = group * WP + wl;
    const int run = (chain && w < P) ? w / K : 0;
    const int j = w - run * K;
    const RunState &st = A.state[run];
    const bool active = chain && (w < P) && !st.done && j < st.Kb && !(A.w_flags[w < P ? w : 0] & WF_FROZEN);
    const double Lstar = st.Lstar;

    // chain state lives in the registers of the chain warps (lane = walker); every CTA keeps an identical copy
    double x[D], xPr = 0.0, xL = 0.0;
    int steps = 0, nacc = 0;
#pragma unroll
    for (int a = 0; a < D; ++a) x[a] = 1.0;
    if (chain && w < P) {
#pragma unroll
        for (int a = 0; a < D; ++a) x[a] = A.w_theta[(size_t)w * D + a];
        xPr = A.w_logPr[w]; xL = A.w_logL[w]; steps = A.w_steps[w]; nacc = A.w_nacc[w];
#pragma unroll
        for (int a = 0; a < D; ++a) s_mean[a * WP + wl] = A.w_mean[(size_t)w * D + a];
#pragma unroll
        for (int a = 0; a < D * D; ++a) s_cov[a * WP + wl] = A.w_cov[(size_t)w * D * D + a];
    }
    // nothing to walk for this group (its runs have terminated: the host enqueues iterations ahead of reading the
    // state): every CTA of the cluster sees the same walkers and takes the same exit
    if (!__syncthreads_or(active ? 1 : 0)) return;
    if (CS > 1) cluster.sync();  // every CTA's mbarriers are initialised before a peer can signal them
    bool pre = false;
    double xn[D], nPr = 0.0;
    const int S = (int)prm.S;

    for (int s = 0; s < S; ++s) {
        // ---- (0) every CH steps: all threads pre-generate the proposal increments L z and log u
        if ((s % CH) == 0) {
            __syncthreads();
            for (int e = tid; e < CH * WP; e += blockDim.x) {
                const int sc = e / WP, ws = e - sc * WP;
                const int ww = group * WP + ws;
                if (ww < P && s + sc < S) {
                    const int rr = ww / K, jj = ww - rr * K;
                    const RunState &sr = A.state[rr];
                    const uint32_t wid_ = (uint32_t)(sr.walk_base + jj), rid_ = prm.first_run_id + rr;
                    const uint32_t stp = (uint32_t)(A.w_steps[ww] + s + sc);
                    double z[D + 1];
#pragma unroll
                    for (int b = 0; b < (D + 1) / 2; ++b)
                        rng_normal2(prm.seed, (uint32_t)(b + 16 * prm.attempt), stp, wid_, TAG_NORMAL, rid_, z[2 * b], z[2 * b + 1]);
#pragma unroll
                    for (int a = 0; a < D; ++a) {
                        double dz = 0.0;
                        if (sr.chol_ok) {
#pragma unroll
                            for (int b = 0; b <= a; ++b) dz += sr.cholL[a * D + b] * z[b];
                        }
                        s_dz[((size_t)sc * D + a) * WP + ws] = dz;
                    }
                    double u0, u1;
                    rng_uniform2(prm.seed, (uint32_t)(16 * prm.attempt), stp, wid_, TAG_ACCEPT, rid_, u0, u1);
                    s_logu[(size_t)sc * WP + ws] = log(u0);
                }
            }
            __syncthreads();
        }
        // ---- (1) chain warps: proposal, box / prior pre-check, per-datum and epilogue coefficients
        typename OP::Coef fin_c{};
        bool fin_ok = false;
        if (chain) {
            const int sc = s % CH;
#pragma unroll
            for (int a = 0; a < D; ++a) {
                // same association as walk_step_kernel: x + (L z) accumulated term by term
                xn[a] = x[a] + ((w < P) ? s_dz[((size_t)sc * D + a) * WP + wl] : 0.0);
            }
            pre = false;
            if (active && in_box<D>(prior, xn)) {
                nPr = 0.0;
#pragma unroll
                for (int a = 0; a < D; ++a) nPr += logprior_dim(prior, a, xn[a]);
                if (!isfinite(nPr)) nPr = prm.logzero;
                pre = (nPr - xPr > s_logu[(size_t)sc * WP + wl]);
            }
            s_row[wl] = OP::make_row(xn, cst);
            fin_c = OP::prepare(xn, fin_ok, cst);  // log sigma, 1/(2 sigma^2), operator constraints
        }
        __syncthreads();
        // ---- (2) all warps: this CTA's shard of the reduction for the WP proposals, TW walkers per lane
        {
            typename OP::Row c[TW];
#pragma unroll
            for (int u = 0; u < TW; ++u) c[u] = s_row[u * 32 + lane];
            typename OP::Acc acc[TW];
#pragma unroll
            for (int u = 0; u < TW; ++u) acc[u] = OP::acc_init();
            if constexpr (TW == 1 && OP::RENORM == 0) {
                // one walker per lane: four independent accumulators, or a single FMA-accumulate chain would be bound by
                // the DFMA latency instead of its issue rate
                typename OP::Acc a1[1] = {OP::acc_init()}, a2[1] = {OP::acc_init()}, a3[1] = {OP::acc_init()};
                int i = wid;
#pragma unroll 2
                for (; i + 3 * NW < nr; i += 4 * NW) {
                    OP::template rows<1>(c, tile + (size_t)i * NCOL, acc);
                    OP::template rows<1>(c, tile + (size_t)(i + NW) * NCOL, a1);
                    OP::template rows<1>(c, tile + (size_t)(i + 2 * NW) * NCOL, a2);
                    OP::template rows<1>(c, tile + (size_t)(i + 3 * NW) * NCOL, a3);
                }
                for (; i < nr; i += NW) OP::template rows<1>(c, tile + (size_t)i * NCOL, acc);
                red[wid * WP + lane] = (OP::acc_value(acc[0]) + OP::acc_value(a1[0])) + (OP::acc_value(a2[0]) + OP::acc_value(a3[0]));
            } else {
                sweep_rows<OP, TW>(c, tile, wid, NW, nr, acc);
#pragma unroll
                for (int u = 0; u < TW; ++u) red[wid * WP + u * 32 + lane] = OP::acc_value(acc[u]);
            }
        }
        __syncthreads();
        // ---- (3) chain warps: fixed-order combine inside the CTA, then all-to-all exchange of the CS partial sums through
        //          distributed shared memory.  Each peer's value arrives with st.async and signals that CTA's mbarrier
        //          (complete_tx), so no cluster-wide barrier sits on the per-step critical path; the two
        //          buffers/mbarriers alternate by step parity (a peer can be at most one step ahead).
        const int par = s & 1;
        if (chain) {
            double part = 0.0;
#pragma unroll
            for (int q = 0; q < NW; ++q) part += red[q * WP + wl];
            double sum = part;
            if (CS > 1) {
                if (tid == 0) mbar_expect_tx(&xbar[par], (uint32_t)(CS * WP * sizeof(double)));
                const uint32_t slot = smem_u32(&xch[(par * CS + rank) * WP + wl]), bar = smem_u32(&xbar[par]);
                for (int dst = 0; dst < CS; ++dst) st_async_f64(mapa_u32(slot, dst), part, mapa_u32(bar, dst));
                mbar_wait(&xbar[par], (uint32_t)((s >> 1) & 1));
                sum = 0.0;
                for (int q = 0; q < CS; ++q) sum += xch[(par * CS + q) * WP + wl];
            }
            // ---- (4) accept rule (nsDensity BS:602-617), Haario recursion (BS:715-727)
            bool acc = false;
            if (active && pre) {
                double nL = op_finish<OP>(fin_c, sum, (double)rows, cst);
                if (!(fin_ok && isfinite(nL))) nL = prm.logzero;  // RuntimeErrorHandler -> logzero, BS:500-503
                if (nL > Lstar) { acc = true; xL = nL; }
            }
            if (acc) {
#pragma unroll
                for (int a = 0; a < D; ++a) x[a] = xn[a];
                xPr = nPr;
                ++nacc;
            }
            if (active) {
                const double t = 10.0 + (double)steps;
                double dm_o[D], dm_n[D];
#pragma unroll
                for (int a = 0; a < D; ++a) {
                    const double mo = s_mean[a * WP + wl];
                    const double mn = mo + (x[a] - mo) / (t + 1.0);
                    dm_o[a] = x[a] - mo;
                    dm_n[a] = x[a] - mn;
                    s_mean[a * WP + wl] = mn;
                }
                const double f = (t - 1.0) / t;
#pragma unroll
                for (int a = 0; a < D; ++a)
#pragma unroll
                    for (int b = 0; b < D; ++b)
                        s_cov[(a * D + b) * WP + wl] = f * s_cov[(a * D + b) * WP + wl] + dm_o[a] * dm_n[b] / t;
                ++steps;
            }
        }
    }
    // ---- write the chain state back (one CTA of the cluster); freeze per BS:730-736
    if (chain && rank == 0 && active) {
#pragma unroll
        for (int a = 0; a < D; ++a) {
            A.w_theta[(size_t)w * D + a] = x[a];
            A.w_mean[(size_t)w * D + a] = s_mean[a * WP + wl];
        }
#pragma unroll
        for (int a = 0; a < D * D; ++a) A.w_cov[(size_t)w * D * D + a] = s_cov[a * WP + wl];
        A.w_logL[w] = xL;
        A.w_logPr[w] = xPr;
        A.w_nacc[w] = nacc;
        A.w_steps[w] = steps;
        int flags = 0;
        const double rate = (double)nacc / (double)steps;
        if ((rate >= prm.acc_min && rate <= prm.acc_max) || steps >= prm.maxS) {
            flags |= WF_FROZEN;
            atomicSub(A.n_unfrozen, 1);
        }
        A.w_flags[w] = flags;
    }
    if (CS > 1) cluster.sync();  // no CTA may exit while peers can still write into its shared memory
}

}  // namespace binest
