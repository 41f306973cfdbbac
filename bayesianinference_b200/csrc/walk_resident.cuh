// walk_resident.cuh — the whole S-step constrained walk in ONE launch for data sets that fit in shared memory.
//
// For small data (C1: 800 B, C4: 256 KB) a likelihood launch per walk step is pure launch latency (~8 us fixed cost
// for < 1 us of fp64 work).  Here a thread-block cluster of CS CTAs owns 32 walkers (lane = walker) for the whole
// walk: the data rows are sharded over the CS shared memories once, every step each CTA reduces its shard for the
// 32 proposals (8 warps split the rows; broadcast LDS feeds the operator's DFMA sequence exactly as in
// loglike_stream_kernel), the CS partial sums are exchanged through distributed shared memory (one
// barrier.cluster per step) and summed in a fixed order, and warp 0 of every CTA — redundantly and
// deterministically — applies the accept rule of nsDensity (BS:602-617), the Haario recursion (BS:715-727) and
// forms the next proposal.  Philox normals for the next 16 steps are produced by all 256 threads at once, off the
// critical path.  Same Philox addressing and arithmetic as walk_step_kernel, so results are identical to the
// stepped path and to the oracle.
#pragma once
#include <cooperative_groups.h>

#include "walk.cuh"

namespace binest {

namespace cg = cooperative_groups;

constexpr int kResChunk = 16;  // walk steps of pre-generated increments held in shared memory
constexpr int kResWarps = 16;  // warp 0: chain logic, warps 1..15: data (few resident warps per SM, so latency is hidden by ILP)

// dynamic shared memory layout (doubles): tile | xch[2][CS][32] | red[kResWarps][32] | dz[kResChunk][D][32] |
//                                          logu[kResChunk][32] | mean[D][32] | cov[D*D][32] | row[32] (OP::Row)
template <class OP>
__host__ __device__ inline size_t resident_smem_doubles(long long rows_per_cta, int CS) {
    const size_t tile = ((size_t)rows_per_cta * OP::NCOL + 1) & ~(size_t)1;
    const size_t rowsz = (sizeof(typename OP::Row) * 32 + 7) / 8;
    return tile + 2 * (size_t)CS * 32 + kResWarps * 32 + (size_t)kResChunk * OP::D * 32 + kResChunk * 32 + OP::D * 32 +
           OP::D * OP::D * 32 + rowsz;
}

template <class OP>
__global__ void __launch_bounds__(kResWarps * 32)
walk_resident_kernel(const __grid_constant__ RunParams prm, RunArrays A, const __grid_constant__ PriorSpec prior,
                     const double *__restrict__ data, long long rows, long long rows_per_cta, const OpCst cst, int CS) {
    constexpr int D = OP::D, NCOL = OP::NCOL;
    extern __shared__ __align__(16) double smem[];
    const size_t tile_sz = ((size_t)rows_per_cta * NCOL + 1) & ~(size_t)1;
    double *tile = smem;
    double *xch = tile + tile_sz;                   // [2][CS][32]
    double *red = xch + 2 * CS * 32;                // [kResWarps][32]
    double *s_dz = red + kResWarps * 32;               // [kResChunk][D][32]
    double *s_logu = s_dz + kResChunk * D * 32;     // [kResChunk][32]
    double *s_mean = s_logu + kResChunk * 32;       // [D][32]
    double *s_cov = s_mean + D * 32;                // [D*D][32]
    typename OP::Row *s_row = reinterpret_cast<typename OP::Row *>(s_cov + D * D * 32);

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

    // ---- data shard of this CTA -> shared memory (once per launch)
    const long long r0 = (long long)rank * rows_per_cta;
    const long long r1 = (r0 + rows_per_cta < rows) ? r0 + rows_per_cta : rows;
    const int nr = r1 > r0 ? (int)(r1 - r0) : 0;
    for (int e = tid; e < nr * NCOL; e += blockDim.x) tile[e] = data[r0 * NCOL + e];

    // ---- walker of this lane (the same in every warp and every CTA of the cluster)
    const int w = group * 32 + lane;
    const int run = (w < P) ? w / K : 0;
    const int j = w - run * K;
    const RunState &st = A.state[run];
    bool active = (w < P) && !st.done && j < st.Kb && !(A.w_flags[w] & WF_FROZEN);
    const uint32_t walk_id = (uint32_t)(st.walk_base + j), run_id = prm.first_run_id + run;
    const double Lstar = st.Lstar;
    const int chol_ok = st.chol_ok;

    // chain state lives in the registers of warp 0 (lane = walker); every CTA keeps an identical copy
    double x[D], xPr = 0.0, xL = 0.0;
    int steps = 0, nacc = 0;
    if (wid == 0) {
#pragma unroll
        for (int a = 0; a < D; ++a) x[a] = (w < P) ? A.w_theta[(size_t)w * D + a] : 1.0;
        if (w < P) {
            xPr = A.w_logPr[w]; xL = A.w_logL[w]; steps = A.w_steps[w]; nacc = A.w_nacc[w];
#pragma unroll
            for (int a = 0; a < D; ++a) s_mean[a * 32 + lane] = A.w_mean[(size_t)w * D + a];
#pragma unroll
            for (int a = 0; a < D * D; ++a) s_cov[a * 32 + lane] = A.w_cov[(size_t)w * D * D + a];
        }
    }
    if (CS > 1) cluster.sync();  // every CTA's mbarriers are initialised before a peer can signal them
    bool pre = false;
    double xn[D], nPr = 0.0;
    const int S = (int)prm.S;

    for (int s = 0; s < S; ++s) {
        // ---- (0) every kResChunk steps: all threads pre-generate the proposal increments L z and log u
        if ((s % kResChunk) == 0) {
            __syncthreads();
            for (int e = tid; e < kResChunk * 32; e += blockDim.x) {
                const int sc = e >> 5, wl = e & 31;
                const int ww = group * 32 + wl;
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
                        s_dz[(sc * D + a) * 32 + wl] = dz;
                    }
                    double u0, u1;
                    rng_uniform2(prm.seed, (uint32_t)(16 * prm.attempt), stp, wid_, TAG_ACCEPT, rid_, u0, u1);
                    s_logu[sc * 32 + wl] = log(u0);
                }
            }
            __syncthreads();
        }
        // ---- (1) warp 0: proposal, box / prior pre-check, per-datum coefficients
        if (wid == 0) {
            const int sc = s % kResChunk;
#pragma unroll
            for (int a = 0; a < D; ++a) {
                // same association as walk_step_kernel: x + (L z) accumulated term by term
                xn[a] = x[a] + s_dz[(sc * D + a) * 32 + lane];
            }
            pre = false;
            if (active && in_box<D>(prior, xn)) {
                nPr = 0.0;
#pragma unroll
                for (int a = 0; a < D; ++a) nPr += logprior_dim(prior, a, xn[a]);
                if (!isfinite(nPr)) nPr = prm.logzero;
                pre = (nPr - xPr > s_logu[sc * 32 + lane]);
            }
            s_row[lane] = OP::make_row(xn, cst);
        }
        __syncthreads();
        // ---- (2) warps 1..7: this CTA's shard of the reduction for the 32 proposals; warp 0 meanwhile evaluates the
        //          epilogue coefficients of its proposal (log sigma, 1/(2 sigma^2), constraints) off the critical path
        typename OP::Coef fin_c{};
        bool fin_ok = false;
        if (wid == 0) {
            fin_c = OP::prepare(xn, fin_ok, cst);
        } else {
            typename OP::Row c[1];
            c[0] = s_row[lane];
            // four independent accumulators: with one walker per lane a single FMA-accumulate chain would be
            // bound by the DFMA latency, not by its issue rate
            typename OP::Acc a0[1] = {OP::acc_init()}, a1[1] = {OP::acc_init()}, a2[1] = {OP::acc_init()}, a3[1] = {OP::acc_init()};
            constexpr int DW = kResWarps - 1;
            int i = wid - 1;
#pragma unroll 2
            for (; i + 3 * DW < nr; i += 4 * DW) {
                OP::template rows<1>(c, tile + (size_t)i * NCOL, a0);
                OP::template rows<1>(c, tile + (size_t)(i + DW) * NCOL, a1);
                OP::template rows<1>(c, tile + (size_t)(i + 2 * DW) * NCOL, a2);
                OP::template rows<1>(c, tile + (size_t)(i + 3 * DW) * NCOL, a3);
                if constexpr (OP::RENORM > 0) {  // one row per accumulator and iteration: small data, cost irrelevant
                    OP::template renorm<1>(a0); OP::template renorm<1>(a1); OP::template renorm<1>(a2); OP::template renorm<1>(a3);
                }
            }
            for (; i < nr; i += DW) {
                OP::template rows<1>(c, tile + (size_t)i * NCOL, a0);
                if constexpr (OP::RENORM > 0) OP::template renorm<1>(a0);
            }
            red[(wid - 1) * 32 + lane] = (OP::acc_value(a0[0]) + OP::acc_value(a1[0])) + (OP::acc_value(a2[0]) + OP::acc_value(a3[0]));
        }
        __syncthreads();
        // ---- (3) warp 0: fixed-order combine inside the CTA, then all-to-all exchange of the CS partial sums through
        //          distributed shared memory.  Each peer's value arrives with st.async and signals that CTA's
        //          mbarrier (complete_tx), so no cluster-wide barrier sits on the per-step critical path; the two
        //          buffers/mbarriers alternate by step parity (a peer can be at most one step ahead).
        const int par = s & 1;
        if (wid == 0) {
            double part = 0.0;
#pragma unroll
            for (int q = 0; q < kResWarps - 1; ++q) part += red[q * 32 + lane];
            double sum = part;
            if (CS > 1) {
                if (lane == 0) mbar_expect_tx(&xbar[par], (uint32_t)(CS * 32 * sizeof(double)));
                __syncwarp();
                const uint32_t slot = smem_u32(&xch[(par * CS + rank) * 32 + lane]), bar = smem_u32(&xbar[par]);
                for (int dst = 0; dst < CS; ++dst) st_async_f64(mapa_u32(slot, dst), part, mapa_u32(bar, dst));
                mbar_wait(&xbar[par], (uint32_t)((s >> 1) & 1));
                sum = 0.0;
                for (int q = 0; q < CS; ++q) sum += xch[(par * CS + q) * 32 + lane];
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
                    const double mo = s_mean[a * 32 + lane];
                    const double mn = mo + (x[a] - mo) / (t + 1.0);
                    dm_o[a] = x[a] - mo;
                    dm_n[a] = x[a] - mn;
                    s_mean[a * 32 + lane] = mn;
                }
                const double f = (t - 1.0) / t;
#pragma unroll
                for (int a = 0; a < D; ++a)
#pragma unroll
                    for (int b = 0; b < D; ++b)
                        s_cov[(a * D + b) * 32 + lane] = f * s_cov[(a * D + b) * 32 + lane] + dm_o[a] * dm_n[b] / t;
                ++steps;
            }
        }
    }
    // ---- write the chain state back (one CTA of the cluster); freeze per BS:730-736
    if (wid == 0 && rank == 0 && active) {
#pragma unroll
        for (int a = 0; a < D; ++a) {
            A.w_theta[(size_t)w * D + a] = x[a];
            A.w_mean[(size_t)w * D + a] = s_mean[a * 32 + lane];
        }
#pragma unroll
        for (int a = 0; a < D * D; ++a) A.w_cov[(size_t)w * D * D + a] = s_cov[a * 32 + lane];
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
