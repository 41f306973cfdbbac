// walk_resident.cuh — the whole S-step constrained walk in ONE launch for data sets that fit in the shared memories
// of a thread-block cluster (C4: 16 383 increments = 256 KB over 8 CTAs; small test problems in one CTA).
//
// For such data a likelihood launch per walk step is pure launch latency (~8 us fixed cost for ~1 us of fp64 work).
// Here a cluster of CS CTAs owns WP = 32 TW walkers for the whole walk:
//   * the data rows are sharded over the CS shared memories once per launch;
//   * every step each CTA reduces ITS shard for all WP proposals: the NW = 16 data warps split the rows, lane = walker, TW walkers
//     register-tiled per lane exactly as in loglike_stream_kernel (one broadcast LDS per row feeds TW DFMA chains — with
//     TW = 1 the kernel was bound by the shared-memory return path: a GBM row is 2 DFMAs per walker against one LDS.128;
//     r01: 0.33 of the fp64 peak on C4, profiles/r01b_resident_c4.md);
//   * TW more warps are the CHAIN warps, one per set of 32 walkers (lane = walker): they combine the NW warp sums in a
//     fixed order, exchange the CS per-CTA sums through distributed shared memory (st.async + mbarrier complete_tx: no
//     cluster-wide barrier on the per-step path; two buffers alternate by step parity, a peer can be at most one step
//     ahead), and — redundantly and deterministically in every CTA — apply the accept rule of nsDensity (BS:602-617) and
//     the Haario recursion (BS:715-727).  They work SPLIT-PHASE (as the walker warps of walk_grid.cuh): while the data
//     warps sweep step s they draw the Philox normals / log u of step s + 1 and form BOTH possible next proposals (from
//     x if the proposal in flight is rejected, from xn if accepted) with box / prior tests and coefficients, so that
//     between the arrival of the sums and the release of the data warps only combine, exchange, one FMA, a compare and a
//     select remain (r2: the chain phase was ~3900 clocks of a 24 000-clock C4 step with the chain work done by data
//     warps after the sums arrived; two named barriers per step replace three CTA-wide ones).
// Same Philox addressing and arithmetic as walk_step_kernel, so results are identical to the stepped path and to the
// oracle.
#pragma once
#include <cooperative_groups.h>

#include "walk.cuh"

namespace binest {

namespace cg = cooperative_groups;

constexpr int kResWarpsMax = 16;  // NW = 16 (or 8: BINEST_RES_NW, an experiment) DATA warps per CTA sweep the rows; TW more
                                  // warps are the CHAIN warps, one per set of 32 walkers

// dynamic shared memory layout (doubles), WP = 32 TW:
//   tile | xch[2][CS][WP] | red[NW][WP] | mean[D][WP] | cov[D*D][WP] | row[WP] (OP::Row)
template <class OP>
__host__ __device__ inline size_t resident_smem_doubles(long long rows_per_cta, int CS, int TW, int NW) {
    const size_t WP = 32 * (size_t)TW;
    const size_t tile = ((size_t)rows_per_cta * OP::NCOL + 1) & ~(size_t)1;
    const size_t rowsz = (sizeof(typename OP::Row) * WP + 7) / 8;
    return tile + 2 * (size_t)CS * WP + (size_t)NW * WP + OP::D * WP + OP::D * OP::D * WP + rowsz;
}

__device__ __forceinline__ void res_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void res_bar_arrive(int id, int nthreads) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// A candidate proposal of a chain lane with everything that does not depend on the likelihood: box / prior test, the
// Metropolis pre-test on the prior ratio, the per-datum coefficients and the epilogue coefficients.
template <class OP>
struct ResCand {
    double x[OP::D], pr;
    bool pre, ok;
    typename OP::Row row;
    typename OP::Coef fin;
};

template <class OP>
__device__ __forceinline__ void res_candidate(ResCand<OP> &c, const double (&base)[OP::D], const double (&dz)[OP::D],
                                              double basePr, double logu, bool active, bool inP, const PriorSpec &prior,
                                              const OpCst &cst, double logzero) {
    constexpr int D = OP::D;
#pragma unroll
    for (int a = 0; a < D; ++a) c.x[a] = base[a] + (inP ? dz[a] : 0.0);  // same association as walk_step_kernel: x + (L z)
    c.pre = false;
    c.pr = 0.0;
    if (active && in_box<D>(prior, c.x)) {
        double nPr = 0.0;
#pragma unroll
        for (int a = 0; a < D; ++a) nPr += logprior_dim(prior, a, c.x[a]);
        if (!isfinite(nPr)) nPr = logzero;
        c.pr = nPr;
        c.pre = (nPr - basePr > logu);
    }
    c.row = OP::make_row(c.x, cst);
    c.fin = OP::prepare(c.x, c.ok, cst);  // log sigma, 1/(2 sigma^2), operator constraints
}

template <class OP, int TW, int NW>
__global__ void __launch_bounds__((NW + TW) * 32, 1)
walk_resident_kernel(const __grid_constant__ RunParams prm, RunArrays A, const __grid_constant__ PriorSpec prior,
                     const double *__restrict__ data, long long rows, long long rows_per_cta, const OpCst cst, int CS) {
    constexpr int D = OP::D, NCOL = OP::NCOL, WP = 32 * TW, NT = (NW + TW) * 32;
    extern __shared__ __align__(16) double smem[];
    const size_t tile_sz = ((size_t)rows_per_cta * NCOL + 1) & ~(size_t)1;
    double *tile = smem;
    double *xch = tile + tile_sz;                   // [2][CS][WP]
    double *red = xch + 2 * CS * WP;                // [NW][WP]
    double *s_mean = red + NW * WP;                 // [D][WP]
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
    const bool chain = wid >= NW;  // chain warp of walker set wid - NW
    const int cset = chain ? wid - NW : 0;

    // ---- data shard of this CTA -> shared memory (once per launch)
    const long long r0 = (long long)rank * rows_per_cta;
    const long long r1 = (r0 + rows_per_cta < rows) ? r0 + rows_per_cta : rows;
    const int nr = r1 > r0 ? (int)(r1 - r0) : 0;
    for (int e = tid; e < nr * NCOL; e += blockDim.x) tile[e] = data[r0 * NCOL + e];

    // ---- walker of this lane in a chain warp (the same in every CTA of the cluster)
    const int wl = cset * 32 + lane;  // walker slot inside the group (chain warps only)
    const int w = group * WP + wl;
    const bool inP = chain && w < P;
    const int run = inP ? w / K : 0;
    const int j = w - run * K;
    const RunState &st = A.state[run];
    const bool active = inP && !st.done && j < st.Kb && !(A.w_flags[inP ? w : 0] & WF_FROZEN);
    // nothing to walk for this group (its runs have terminated: the host enqueues iterations ahead of reading the
    // state): every CTA of the cluster sees the same walkers and takes the same exit
    if (!__syncthreads_or(active ? 1 : 0)) return;
    if (CS > 1) cluster.sync();  // every CTA's mbarriers are initialised before a peer can signal them
    const int S = (int)prm.S;

    if (!chain) {
        // =============================== data warps ===============================
        // every step: wait for the proposals' per-datum coefficients, reduce this CTA's shard for the WP proposals (the NW
        // warps split the rows, lane = walker, TW walkers register-tiled per lane), hand the warp sums to the chain warps
        for (int s = 0; s < S; ++s) {
            res_bar_sync(1, NT);
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
            __threadfence_block();
            res_bar_arrive(2, NT);
        }
    } else {
        // =============================== chain warps (lane = walker; identical in every CTA of the cluster) ==========
        // Split-phase: everything of step s + 1 that does not need the likelihood of step s — the Philox draws, BOTH
        // possible proposals (from x if the proposal in flight is rejected, from xn if it is accepted) with their box /
        // prior tests and coefficients, and the Haario recursion of step s — runs while the data warps sweep.  Between the
        // arrival of the partial sums and the release of the data warps only the fixed-order combine, the DSMEM
        // exchange, one fused multiply-add, a compare and a select remain.
        const double Lstar = st.Lstar;
        double x[D], xPr = 0.0, xL = 0.0;
        int steps = 0, nacc = 0, steps0 = 0;
#pragma unroll
        for (int a = 0; a < D; ++a) x[a] = 1.0;
        if (inP) {
#pragma unroll
            for (int a = 0; a < D; ++a) x[a] = A.w_theta[(size_t)w * D + a];
            xPr = A.w_logPr[w]; xL = A.w_logL[w]; steps = A.w_steps[w]; nacc = A.w_nacc[w];
            steps0 = steps;
#pragma unroll
            for (int a = 0; a < D; ++a) s_mean[a * WP + wl] = A.w_mean[(size_t)w * D + a];
#pragma unroll
            for (int a = 0; a < D * D; ++a) s_cov[a * WP + wl] = A.w_cov[(size_t)w * D * D + a];
        }
        const uint32_t wid_ = (uint32_t)(st.walk_base + j), rid_ = prm.first_run_id + (uint32_t)run;
        // proposal increment L z and log u of walk step sc (Philox addressed by the chain's step count)
        auto draws = [&](int sc, double (&dz)[D], double &logu) {
#pragma unroll
            for (int a = 0; a < D; ++a) dz[a] = 0.0;
            logu = 0.0;
            if (!inP) return;
            const uint32_t stp = (uint32_t)(steps0 + sc);
            double z[D + 1];
#pragma unroll
            for (int b = 0; b < (D + 1) / 2; ++b)
                rng_normal2(prm.seed, (uint32_t)(b + 16 * prm.attempt), stp, wid_, TAG_NORMAL, rid_, z[2 * b], z[2 * b + 1]);
            if (st.chol_ok) {
#pragma unroll
                for (int a = 0; a < D; ++a) {
                    double v = 0.0;
#pragma unroll
                    for (int b = 0; b <= a; ++b) v += st.cholL[a * D + b] * z[b];
                    dz[a] = v;
                }
            }
            double u0, u1;
            rng_uniform2(prm.seed, (uint32_t)(16 * prm.attempt), stp, wid_, TAG_ACCEPT, rid_, u0, u1);
            logu = log(u0);
        };
        ResCand<OP> cur, c0, c1;  // proposal in flight; next proposal if it is rejected / accepted
        if (S > 0) {
            double dz[D], logu;
            draws(0, dz, logu);
            res_candidate<OP>(cur, x, dz, xPr, logu, active, inP, prior, cst, prm.logzero);
            s_row[wl] = cur.row;
            __threadfence_block();
            res_bar_arrive(1, NT);
            if (S > 1) {
                draws(1, dz, logu);
                res_candidate<OP>(c0, x, dz, xPr, logu, active, inP, prior, cst, prm.logzero);
                res_candidate<OP>(c1, cur.x, dz, cur.pr, logu, active, inP, prior, cst, prm.logzero);
            }
        }
        for (int s = 0; s < S; ++s) {
            res_bar_sync(2, NT);
            // ---- fixed-order combine inside the CTA, then all-to-all exchange of the CS partial sums through distributed
            //      shared memory.  Each peer's value arrives with st.async and signals that CTA's mbarrier (complete_tx):
            //      no cluster-wide barrier on the per-step path; the two buffers/mbarriers alternate by step parity (a
            //      peer can be at most one step ahead).
            const int par = s & 1;
            double part = 0.0;
#pragma unroll
            for (int q = 0; q < NW; ++q) part += red[q * WP + wl];
            double sum = part;
            if (CS > 1) {
                if (cset == 0 && lane == 0) mbar_expect_tx(&xbar[par], (uint32_t)(CS * WP * sizeof(double)));
                const uint32_t slot = smem_u32(&xch[(par * CS + rank) * WP + wl]), bar = smem_u32(&xbar[par]);
                for (int dst = 0; dst < CS; ++dst) st_async_f64(mapa_u32(slot, dst), part, mapa_u32(bar, dst));
                mbar_wait(&xbar[par], (uint32_t)((s >> 1) & 1));
                sum = 0.0;
                for (int q = 0; q < CS; ++q) sum += xch[(par * CS + q) * WP + wl];
            }
            // ---- accept rule (nsDensity BS:602-617)
            bool acc = false;
            double nL = 0.0;
            if (active && cur.pre) {
                nL = op_finish<OP>(cur.fin, sum, (double)rows, cst);
                if (!(cur.ok && isfinite(nL))) nL = prm.logzero;  // RuntimeErrorHandler -> logzero, BS:500-503
                acc = nL > Lstar;
            }
            if (s + 1 < S) {
                s_row[wl] = acc ? c1.row : c0.row;
                __threadfence_block();
                res_bar_arrive(1, NT);  // the data warps start on step s + 1
            }
            // ---- off the critical path: adopt, Haario recursion (BS:715-727), candidates of step s + 2
            if (acc) {
#pragma unroll
                for (int a = 0; a < D; ++a) x[a] = cur.x[a];
                xPr = cur.pr;
                xL = nL;
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
            if (s + 1 < S) {
                cur = acc ? c1 : c0;
                if (s + 2 < S) {
                    double dz[D], logu;
                    draws(s + 2, dz, logu);
                    res_candidate<OP>(c0, x, dz, xPr, logu, active, inP, prior, cst, prm.logzero);
                    res_candidate<OP>(c1, cur.x, dz, cur.pr, logu, active, inP, prior, cst, prm.logzero);
                }
            }
        }
        // ---- write the chain state back (one CTA of the cluster); freeze per BS:730-736
        if (rank == 0 && active) {
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
    }
    if (CS > 1) cluster.sync();  // no CTA may exit while peers can still write into its shared memory
}

}  // namespace binest
