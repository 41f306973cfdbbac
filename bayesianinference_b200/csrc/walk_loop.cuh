// walk_loop.cuh — the whole nested-sampling loop of nestedSamplingInternal (BS:967-1022) resident on the device, for
// small, latency-bound problems (BASELINE config C1: 100 data rows, 100 live points, one replacement per iteration).
//
// Why.  With the per-iteration kernels (run_update_kernel, then a walk launch) the host synchronises three times per
// iteration and a 200-step walk on 100 rows costs ~420 us (r01f: 2 378 replacements/s at K = 1) — launch latency,
// stream round trips and block-wide barriers inside the walk, not arithmetic.  Here ONE launch runs iterations until
// the run terminates (or a budget of iterations / the dead-list capacity is used up):
//   * one CTA per run; the data rows are loaded into shared memory once per launch;
//   * the update (sort, evidence, termination test, kill, covariance blend, starts: run_update_body, walk.cuh) is
//     executed by the whole CTA, exactly the code of run_update_kernel;
//   * the walks are WARP-PER-WALKER: warp j walks walker j of the batch (K <= warps), speculating over rejections
//     (warp_walk below): the proposals of the next 8 steps are scored at once by 8 groups of 4 lanes and the first
//     accepted one ends the round — no block barrier inside a walk.  Philox normals and log u are drawn 32 steps at a
//     time (lane = step) and handed out by shuffle.
//   * the host polls nothing inside the launch; it reads the run state once per launch.
// Same Philox addressing, accept rule and Haario recursion as walk_step_walker, so trajectories coincide with the
// other walk paths and with the oracle up to the summation order of the likelihood (pinned by
// test_every_walk_path_matches_oracle).
#pragma once
#include "walk.cuh"

namespace binest {

struct LoopCtl {
    long long iters;      // iterations executed by the launch (max over runs)
    int need_grow;        // the dead list of some run is full: the host grows it and relaunches
    int pad_;
};

// dynamic shared memory: tile[rows * NCOL (even)] doubles | s_key[n_pad] doubles | s_idx[n_pad] ints
template <class OP>
__host__ __device__ inline size_t loop_smem_bytes(long long rows, int n_pad) {
    const size_t tile = (((size_t)rows * OP::NCOL + 1) & ~(size_t)1);
    return (tile + (size_t)n_pad) * sizeof(double) + (size_t)n_pad * sizeof(int);
}

// One walker, one warp, S steps — speculative over rejections.
// A Metropolis chain only moves when a proposal is accepted; while proposals are rejected the base point stays put, so
// the proposals of the next G = 8 steps can all be formed from the current point and scored AT ONCE (8 groups of 4
// lanes, each group its own step: draws of that step, box / prior test, likelihood over the data rows split 4 ways).
// The first accepted group g* ends the round: steps s .. s+g* are committed (g* rejections and one move), later groups
// are discarded and their steps re-done from the new point in the next round — with the SAME Philox draws, because
// draws are addressed by step.  The chain is therefore exactly the sequential one (same draws, same decisions, same
// points); only the latency changes: ~1/acceptance steps per round instead of one.  The Haario recursion is applied
// step by step for the committed steps (its divisors, which depend on the step count only, are precomputed by the
// lanes in parallel and multiplied in: a last-ulp difference to x / t, far inside the trajectory tolerance).
constexpr int kSpecGroups = 8, kSpecLanes = 4;

template <class OP>
__device__ __forceinline__ void warp_walk(const RunParams &prm, const RunArrays &A, const PriorSpec &prior,
                                          const double *__restrict__ tile, int nr, double rows, const OpCst &cst, int w,
                                          int lane) {
    constexpr int D = OP::D, NZ = (D + 1) / 2, NCOL = OP::NCOL, G = kSpecGroups, GL = kSpecLanes;
    const int K = prm.K;
    const int r = w / K, j = w - r * K;
    const RunState &st = A.state[r];
    const uint32_t walk_id = (uint32_t)(st.walk_base + j), run_id = prm.first_run_id + (uint32_t)r;
    const double Lstar = st.Lstar;
    const int chol_ok = st.chol_ok;
    double L[D * (D + 1) / 2];
#pragma unroll
    for (int a = 0; a < D; ++a)
#pragma unroll
        for (int b = 0; b <= a; ++b) L[a * (a + 1) / 2 + b] = st.cholL[a * D + b];

    double x[D], mean[D];
#pragma unroll
    for (int a = 0; a < D; ++a) { x[a] = A.w_theta[(size_t)w * D + a]; mean[a] = A.w_mean[(size_t)w * D + a]; }
    // covariance entries spread over the lanes: lane e (+32 k) holds entry e
    constexpr int NC = (D * D + 31) / 32;
    double cov[NC];
#pragma unroll
    for (int e = 0; e < NC; ++e) cov[e] = (lane + 32 * e < D * D) ? A.w_cov[(size_t)w * D * D + lane + 32 * e] : 0.0;
    double xPr = A.w_logPr[w], xL = A.w_logL[w];
    const int steps0 = A.w_steps[w];
    int nacc = A.w_nacc[w];
    const int S = (int)prm.S;
    const int g = lane / GL, sub = lane - g * GL;
    const unsigned gmask = ((1u << GL) - 1u) << (g * GL);

    double zl[2 * NZ], lul = 0.0;  // this lane's draws for step (cb + lane)
    int cb = -64;                  // first step of the current chunk of draws
    int s = 0;
    while (s < S) {
        if (s + G > cb + 32) {  // the window [s, s + G) must lie inside the chunk
            cb = s;
            const uint32_t c = (uint32_t)(steps0 + cb + lane);
#pragma unroll
            for (int b = 0; b < NZ; ++b)
                rng_normal2(prm.seed, (uint32_t)(b + 16 * prm.attempt), c, walk_id, TAG_NORMAL, run_id, zl[2 * b], zl[2 * b + 1]);
            double u0, u1;
            rng_uniform2(prm.seed, (uint32_t)(16 * prm.attempt), c, walk_id, TAG_ACCEPT, run_id, u0, u1);
            lul = log(u0);
        }
        // ---- group g scores the proposal of step s + g, formed from the current point
        const int src = s + g - cb;  // < 32
        double z[D];
#pragma unroll
        for (int b = 0; b < D; ++b) z[b] = __shfl_sync(0xffffffffu, zl[b], src);
        const double logu = __shfl_sync(0xffffffffu, lul, src);
        double xn[D];
#pragma unroll
        for (int a = 0; a < D; ++a) {
            double v = x[a];  // x' = x + L z (same association as walk_step_walker)
            if (chol_ok) {
#pragma unroll
                for (int b = 0; b <= a; ++b) v += L[a * (a + 1) / 2 + b] * z[b];
            }
            xn[a] = v;
        }
        double nPr = 0.0;
        bool pre = false;
        if (s + g < S && in_box<D>(prior, xn)) {
#pragma unroll
            for (int a = 0; a < D; ++a) nPr += logprior_dim(prior, a, xn[a]);
            if (!isfinite(nPr)) nPr = prm.logzero;
            pre = nPr - xPr > logu;  // Metropolis rule on the log density
        }
        // likelihood of every group's proposal (unconditionally: the warp runs the code once either way)
        bool ok;
        const typename OP::Coef cf = OP::prepare(xn, ok, cst);
        typename OP::Row c[1] = {OP::make_row(xn, cst)};
        // four independent accumulators: a single FMA-accumulate chain over ~25 rows would be DFMA-latency bound
        typename OP::Acc a0[1] = {OP::acc_init()}, a1[1] = {OP::acc_init()}, a2[1] = {OP::acc_init()}, a3[1] = {OP::acc_init()};
        {
            int i = sub;
            for (; i + 3 * GL < nr; i += 4 * GL) {
                OP::template rows<1>(c, tile + (size_t)i * NCOL, a0);
                OP::template rows<1>(c, tile + (size_t)(i + GL) * NCOL, a1);
                OP::template rows<1>(c, tile + (size_t)(i + 2 * GL) * NCOL, a2);
                OP::template rows<1>(c, tile + (size_t)(i + 3 * GL) * NCOL, a3);
                if constexpr (OP::RENORM > 0) {
                    OP::template renorm<1>(a0); OP::template renorm<1>(a1); OP::template renorm<1>(a2); OP::template renorm<1>(a3);
                }
            }
            for (; i < nr; i += GL) {
                OP::template rows<1>(c, tile + (size_t)i * NCOL, a0);
                if constexpr (OP::RENORM > 0) OP::template renorm<1>(a0);
            }
        }
        double sum = (OP::acc_value(a0[0]) + OP::acc_value(a1[0])) + (OP::acc_value(a2[0]) + OP::acc_value(a3[0]));
        sum += __shfl_xor_sync(0xffffffffu, sum, 1);
        sum += __shfl_xor_sync(0xffffffffu, sum, 2);
        double nL = op_finish<OP>(cf, sum, rows, cst);
        if (!(ok && isfinite(nL))) nL = prm.logzero;  // RuntimeErrorHandler -> logzero, BS:500-503
        const bool accg = pre && nL > Lstar;              // nsDensity: logL > threshold, strict (BS:605)
        // ---- first accepted group ends the round
        const unsigned votes = __ballot_sync(0xffffffffu, accg && sub == 0);
        const int remaining = S - s;
        int gstar = votes ? (__ffs(votes) - 1) / GL : -1;
        const int n_adv = gstar >= 0 ? gstar + 1 : (remaining < G ? remaining : G);
        double xa[D], aPr = 0.0, aL = 0.0;
        if (gstar >= 0) {
#pragma unroll
            for (int a = 0; a < D; ++a) xa[a] = __shfl_sync(0xffffffffu, xn[a], gstar * GL);
            aPr = __shfl_sync(0xffffffffu, nPr, gstar * GL);
            aL = __shfl_sync(0xffffffffu, nL, gstar * GL);
        }
        // ---- Haario recursion for the committed steps, started at t = 10 (BS:715-727); divisors by lane k = step s + k
        const double tk = 10.0 + (double)(steps0 + s + (lane < G ? lane : 0));
        const double r1k = 1.0 / (tk + 1.0), fk = (tk - 1.0) / tk, r3k = 1.0 / tk;
        double r1[G], ff[G], r3[G];  // handed out before the recursion: no shuffle latency inside its dependent chain
#pragma unroll
        for (int k = 0; k < G; ++k) {
            r1[k] = __shfl_sync(0xffffffffu, r1k, k);
            ff[k] = __shfl_sync(0xffffffffu, fk, k);
            r3[k] = __shfl_sync(0xffffffffu, r3k, k);
        }
#pragma unroll
        for (int k = 0; k < G; ++k) {
            if (k < n_adv) {
                if (k == gstar) {
#pragma unroll
                    for (int a = 0; a < D; ++a) x[a] = xa[a];
                    xPr = aPr;
                    xL = aL;
                    ++nacc;
                }
                double dm_o[D], dm_n[D];
#pragma unroll
                for (int a = 0; a < D; ++a) {
                    const double mn = fma(x[a] - mean[a], r1[k], mean[a]);
                    dm_o[a] = x[a] - mean[a];
                    dm_n[a] = x[a] - mn;
                    mean[a] = mn;
                }
#pragma unroll
                for (int e = 0; e < NC; ++e) {
                    const int idx = lane + 32 * e;
                    if (idx < D * D) {
                        const int a = idx / D, b = idx - a * D;
                        double da = 0.0, db = 0.0;
#pragma unroll
                        for (int q = 0; q < D; ++q) { if (q == a) da = dm_o[q]; if (q == b) db = dm_n[q]; }
                        cov[e] = fma(ff[k], cov[e], da * db * r3[k]);
                    }
                }
            }
        }
        s += n_adv;
    }
    // chain state back (the update of the next iteration adopts it, BS:999, 1006-1016)
    if (lane == 0) {
#pragma unroll
        for (int a = 0; a < D; ++a) { A.w_theta[(size_t)w * D + a] = x[a]; A.w_mean[(size_t)w * D + a] = mean[a]; }
        A.w_logL[w] = xL; A.w_logPr[w] = xPr; A.w_nacc[w] = nacc; A.w_steps[w] = steps0 + S;
        A.w_flags[w] = WF_FROZEN;
    }
#pragma unroll
    for (int e = 0; e < NC; ++e)
        if (lane + 32 * e < D * D) A.w_cov[(size_t)w * D * D + lane + 32 * e] = cov[e];
    (void)gmask;
}

// ---------------------------------------------------------------------------------------------------
// K = 1 (the reference scheme: one replacement per iteration, BS:980-1018): ONE walker per CTA, so seven of the eight
// warps idled at the barrier while warp 0 walked.  Here all eight warps speculate together, two levels deep:
//   warp 0      scores the proposals of steps s .. s+7 from the current point x (as warp_walk does);
//   warp h + 1  (h = 0..5) ASSUMES proposal h of warp 0 is the first one accepted and scores the proposals of steps
//               s+h+1 .. s+h+8 from that candidate point (which it forms itself with the same arithmetic);
//   warp 7      keeps the books: the Haario recursion of the steps committed by the previous round.
// When warp 0's first acceptance g1 is known, the round that follows it has already been scored by warp g1 + 1: a
// super-round commits the steps up to the SECOND acceptance (~2/acceptance-rate steps instead of ~1/rate).  The chain is
// still exactly the sequential one: every proposal is formed from the point the sequential chain would hold, with the
// draws of its own step (Philox is addressed by step; all draws of the walk and the Haario divisors, which depend on the
// step count only, are tabulated in shared memory by all 256 threads before the first round).
constexpr int kK1MaxTableBytes = 96 * 1024;
template <class OP>
__host__ __device__ inline size_t k1_table_doubles(long long S) { return (size_t)S * (OP::D + 4); }

template <class OP>
struct K1Result {
    double xa[OP::D], aPr, aL;
    int gstar;  // first accepted group of the warp's round, -1: none
    int pad_;
};

// one speculative round by one warp: 8 groups of 4 lanes, group g scores the proposal of step start + g formed from base
template <class OP>
__device__ __forceinline__ void k1_round(const RunParams &prm, const PriorSpec &prior, const double *__restrict__ tile, int nr,
                                         double rows, const OpCst &cst, const double (&L)[OP::D * (OP::D + 1) / 2], int chol_ok,
                                         double Lstar, const double (&base)[OP::D], double basePr, int start, bool valid, int S,
                                         const double *__restrict__ tz, const double *__restrict__ tlogu, int lane,
                                         K1Result<OP> &out) {
    constexpr int D = OP::D, NCOL = OP::NCOL, G = kSpecGroups, GL = kSpecLanes;
    const int g = lane / GL, sub = lane - g * GL;
    const int sg = start + g;
    const bool live = valid && sg < S;
    const int si = sg < S ? sg : S - 1;
    double xn[D];
#pragma unroll
    for (int a = 0; a < D; ++a) {
        double v = base[a];  // x' = x + L z (same association as walk_step_walker)
        if (chol_ok) {
#pragma unroll
            for (int b = 0; b <= a; ++b) v += L[a * (a + 1) / 2 + b] * tz[(size_t)si * D + b];
        }
        xn[a] = v;
    }
    double nPr = 0.0;
    bool pre = false;
    if (live && in_box<D>(prior, xn)) {
#pragma unroll
        for (int a = 0; a < D; ++a) nPr += logprior_dim(prior, a, xn[a]);
        if (!isfinite(nPr)) nPr = prm.logzero;
        pre = nPr - basePr > tlogu[si];  // Metropolis rule on the log density
    }
    bool ok;
    const typename OP::Coef cf = OP::prepare(xn, ok, cst);
    typename OP::Row c[1] = {OP::make_row(xn, cst)};
    typename OP::Acc a0[1] = {OP::acc_init()}, a1[1] = {OP::acc_init()}, a2[1] = {OP::acc_init()}, a3[1] = {OP::acc_init()};
    {
        int i = sub;
        for (; i + 3 * GL < nr; i += 4 * GL) {
            OP::template rows<1>(c, tile + (size_t)i * NCOL, a0);
            OP::template rows<1>(c, tile + (size_t)(i + GL) * NCOL, a1);
            OP::template rows<1>(c, tile + (size_t)(i + 2 * GL) * NCOL, a2);
            OP::template rows<1>(c, tile + (size_t)(i + 3 * GL) * NCOL, a3);
            if constexpr (OP::RENORM > 0) {
                OP::template renorm<1>(a0); OP::template renorm<1>(a1); OP::template renorm<1>(a2); OP::template renorm<1>(a3);
            }
        }
        for (; i < nr; i += GL) {
            OP::template rows<1>(c, tile + (size_t)i * NCOL, a0);
            if constexpr (OP::RENORM > 0) OP::template renorm<1>(a0);
        }
    }
    double sum = (OP::acc_value(a0[0]) + OP::acc_value(a1[0])) + (OP::acc_value(a2[0]) + OP::acc_value(a3[0]));
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
    double nL = op_finish<OP>(cf, sum, rows, cst);
    if (!(ok && isfinite(nL))) nL = prm.logzero;  // RuntimeErrorHandler -> logzero, BS:500-503
    const bool accg = pre && nL > Lstar;              // nsDensity: logL > threshold, strict (BS:605)
    const unsigned votes = __ballot_sync(0xffffffffu, accg && sub == 0);
    const int gstar = votes ? (__ffs(votes) - 1) / GL : -1;
    const int src = gstar >= 0 ? gstar * GL : 0;
    double xa[D];
#pragma unroll
    for (int a = 0; a < D; ++a) xa[a] = __shfl_sync(0xffffffffu, xn[a], src);
    const double aPr = __shfl_sync(0xffffffffu, nPr, src), aL = __shfl_sync(0xffffffffu, nL, src);
    if (lane == 0) {
#pragma unroll
        for (int a = 0; a < D; ++a) out.xa[a] = xa[a];
        out.aPr = aPr; out.aL = aL; out.gstar = gstar;
    }
    (void)G;
}

// the S-step walk of the run's single walker by the whole CTA (256 threads = 8 warps)
template <class OP>
__device__ __forceinline__ void cta_walk_k1(const RunParams &prm, const RunArrays &A, const PriorSpec &prior,
                                            const double *__restrict__ tile, int nr, double rows, const OpCst &cst, int w,
                                            double *__restrict__ tab) {
    constexpr int D = OP::D, NZ = (D + 1) / 2, G = kSpecGroups;
    __shared__ K1Result<OP> s_res[8];
    __shared__ double s_x[D], s_pl[2];
    const int tid = threadIdx.x, lane = tid & 31, v = tid >> 5;
    const int S = (int)prm.S;
    const RunState &st = A.state[w];  // K = 1: walker w belongs to run w
    const uint32_t walk_id = (uint32_t)st.walk_base, run_id = prm.first_run_id + (uint32_t)w;
    const double Lstar = st.Lstar;
    const int chol_ok = st.chol_ok;
    double L[D * (D + 1) / 2];
#pragma unroll
    for (int a = 0; a < D; ++a)
#pragma unroll
        for (int b = 0; b <= a; ++b) L[a * (a + 1) / 2 + b] = st.cholL[a * D + b];
    const int steps0 = A.w_steps[w];
    double *tz = tab, *tlogu = tab + (size_t)S * D, *tr1 = tlogu + S, *tff = tr1 + S, *tr3 = tff + S;
    // ---- tables: draws and Haario divisors of every step of the walk, one step per thread
    for (int sc = tid; sc < S; sc += blockDim.x) {
        const uint32_t c = (uint32_t)(steps0 + sc);
        double z[2 * NZ];
#pragma unroll
        for (int b = 0; b < NZ; ++b)
            rng_normal2(prm.seed, (uint32_t)(b + 16 * prm.attempt), c, walk_id, TAG_NORMAL, run_id, z[2 * b], z[2 * b + 1]);
#pragma unroll
        for (int b = 0; b < D; ++b) tz[(size_t)sc * D + b] = z[b];
        double u0, u1;
        rng_uniform2(prm.seed, (uint32_t)(16 * prm.attempt), c, walk_id, TAG_ACCEPT, run_id, u0, u1);
        tlogu[sc] = log(u0);
        const double tk = 10.0 + (double)(steps0 + sc);  // Haario recursion started at t = 10 (BS:715-727)
        tr1[sc] = 1.0 / (tk + 1.0); tff[sc] = (tk - 1.0) / tk; tr3[sc] = 1.0 / tk;
    }
    if (tid < D) s_x[tid] = A.w_theta[(size_t)w * D + tid];
    if (tid == 0) { s_pl[0] = A.w_logPr[w]; s_pl[1] = A.w_logL[w]; }
    // Haario state lives in warp 0: mean in every lane, covariance entries spread over the lanes
    double mean[D];
#pragma unroll
    for (int a = 0; a < D; ++a) mean[a] = A.w_mean[(size_t)w * D + a];
    constexpr int NC = (D * D + 31) / 32;
    double cov[NC];
#pragma unroll
    for (int e = 0; e < NC; ++e) cov[e] = (lane + 32 * e < D * D) ? A.w_cov[(size_t)w * D * D + lane + 32 * e] : 0.0;
    int nacc = A.w_nacc[w];
    __syncthreads();

    // warps 0 .. kK1Eval-1 evaluate (warp 0 the round from x, warp h + 1 the round from candidate h); the last warp is the
    // BOOK-KEEPER: it runs the Haario recursion of the steps committed by round r while the others score round r + 1
    // (the recursion is a serial chain of fp64 latencies per committed step and nothing in the walk depends on it)
    constexpr int kK1Eval = 7;
    const bool keeper = v == kK1Eval;
    double hx[D], hx1[D], hx2[D];  // book-keeper: chain point before the round, after the first / second acceptance
    int hs = 0, hn = 0, hg1 = -1, hk2 = -1;
#pragma unroll
    for (int a = 0; a < D; ++a) { hx[a] = A.w_theta[(size_t)w * D + a]; hx1[a] = hx2[a] = 0.0; }  // hx follows the chain
    auto haario = [&]() {
        for (int k = 0; k < hn; ++k) {
            if (k == hg1) {
#pragma unroll
                for (int a = 0; a < D; ++a) hx[a] = hx1[a];
            }
            if (k == hk2) {
#pragma unroll
                for (int a = 0; a < D; ++a) hx[a] = hx2[a];
            }
            const double r1 = tr1[hs + k], ff = tff[hs + k], r3 = tr3[hs + k];
            double dm_o[D], dm_n[D];
#pragma unroll
            for (int a = 0; a < D; ++a) {
                const double mn = fma(hx[a] - mean[a], r1, mean[a]);
                dm_o[a] = hx[a] - mean[a];
                dm_n[a] = hx[a] - mn;
                mean[a] = mn;
            }
#pragma unroll
            for (int e = 0; e < NC; ++e) {
                const int idx = lane + 32 * e;
                if (idx < D * D) {
                    const int a = idx / D, b = idx - a * D;
                    double da = 0.0, db = 0.0;
#pragma unroll
                    for (int q = 0; q < D; ++q) { if (q == a) da = dm_o[q]; if (q == b) db = dm_n[q]; }
                    cov[e] = fma(ff, cov[e], da * db * r3);
                }
            }
        }
        hn = 0;
    };

    int s = 0;
    while (s < S) {  // s, and everything it is derived from, is identical in all threads
        if (keeper) {
            haario();  // the steps the previous round committed
        } else {
            double x[D];
#pragma unroll
            for (int a = 0; a < D; ++a) x[a] = s_x[a];
            const double xPr = s_pl[0];
            // ---- this warp's base point: the current point (warp 0) or warp 0's candidate h = v - 1
            double base[D], basePr = xPr;
            int start = s;
            bool valid = true;
#pragma unroll
            for (int a = 0; a < D; ++a) base[a] = x[a];
            if (v > 0) {
                const int sh = s + v - 1;
                start = sh + 1;
                valid = start < S;
                const int si = sh < S ? sh : S - 1;
#pragma unroll
                for (int a = 0; a < D; ++a) {
                    double t = x[a];
                    if (chol_ok) {
#pragma unroll
                        for (int b = 0; b <= a; ++b) t += L[a * (a + 1) / 2 + b] * tz[(size_t)si * D + b];
                    }
                    base[a] = t;
                }
                valid = valid && in_box<D>(prior, base);
                basePr = 0.0;
                if (valid) {
#pragma unroll
                    for (int a = 0; a < D; ++a) basePr += logprior_dim(prior, a, base[a]);
                    if (!isfinite(basePr)) basePr = prm.logzero;
                }
            }
            if (valid) k1_round<OP>(prm, prior, tile, nr, rows, cst, L, chol_ok, Lstar, base, basePr, start, valid, S, tz, tlogu, lane, s_res[v]);
            else if (lane == 0) s_res[v].gstar = -1;
        }
        __syncthreads();
        // ---- resolution (every thread, same values): first acceptance g1 of warp 0, then the round warp g1 + 1 scored
        const int g1 = s_res[0].gstar;
        const int rem = S - s;
        int n1, n2 = 0, g2 = -1;
        if (g1 < 0) {
            n1 = rem < G ? rem : G;
        } else {
            n1 = g1 + 1;
            const int rem2 = S - (s + n1);
            if (rem2 > 0 && g1 + 1 < kK1Eval) {
                g2 = s_res[g1 + 1].gstar;
                n2 = g2 >= 0 ? g2 + 1 : (rem2 < G ? rem2 : G);
            }
        }
        if (keeper) {  // what the recursion needs, before the result slots are overwritten by the next round
#pragma unroll
            for (int a = 0; a < D; ++a) {
                hx1[a] = s_res[0].xa[a];
                hx2[a] = s_res[g2 >= 0 ? g1 + 1 : 0].xa[a];
            }
            hs = s; hn = n1 + n2; hg1 = g1; hk2 = g2 >= 0 ? n1 + g2 : -1;
        }
        if (g1 >= 0) {
            const K1Result<OP> &fin = (g2 >= 0) ? s_res[g1 + 1] : s_res[0];
            double nx = 0.0, npr = 0.0, nl = 0.0;
            if (tid < D) nx = fin.xa[tid];
            if (tid == 0) { npr = fin.aPr; nl = fin.aL; }
            __syncwarp();
            // s_x / s_pl are read at the top of a round only: nobody reads them between the barrier above and the one below
            if (!keeper) {
                if (tid < D) s_x[tid] = nx;
                if (tid == 0) { s_pl[0] = npr; s_pl[1] = nl; }
            }
            nacc += (g2 >= 0) ? 2 : 1;
        }
        s += n1 + n2;
        __syncthreads();
    }
    if (keeper) haario();  // the last round's steps
    __syncthreads();
    // mean / covariance live in the book-keeper warp: hand them to the writers below
    __shared__ double s_hm[D];
    if (keeper && lane < D) {
#pragma unroll
        for (int a = 0; a < D; ++a) if (lane == a) s_hm[a] = mean[a];
    }
    if (keeper) {
#pragma unroll
        for (int e = 0; e < NC; ++e)
            if (lane + 32 * e < D * D) A.w_cov[(size_t)w * D * D + lane + 32 * e] = cov[e];
    }
    __syncthreads();
#pragma unroll
    for (int a = 0; a < D; ++a) mean[a] = s_hm[a];
    // chain state back (the update of the next iteration adopts it, BS:999, 1006-1016)
    if (tid == 0) {
#pragma unroll
        for (int a = 0; a < D; ++a) { A.w_theta[(size_t)w * D + a] = s_x[a]; A.w_mean[(size_t)w * D + a] = mean[a]; }
        A.w_logL[w] = s_pl[1]; A.w_logPr[w] = s_pl[0]; A.w_nacc[w] = nacc; A.w_steps[w] = steps0 + S;
        A.w_flags[w] = WF_FROZEN;
    }
}

// one CTA of NT threads per run: NT = 256 (K <= 8 walkers per iteration) or 1024 (K <= 32)
template <class OP, int NT>
__global__ void __launch_bounds__(NT)
ns_loop_kernel(const __grid_constant__ RunParams prm, RunArrays A, const __grid_constant__ PriorSpec prior,
               const double *__restrict__ data, long long rows, const OpCst cst, int n_pad, int first_mode,
               long long max_iters, LoopCtl *ctl, int k1_tables /* K = 1: shared-memory tables for cta_walk_k1 follow s_idx */) {
    constexpr int NCOL = OP::NCOL;
    extern __shared__ __align__(16) double lsm[];
    const size_t tile_sz = (((size_t)rows * NCOL + 1) & ~(size_t)1);
    double *tile = lsm;
    double *s_key = tile + tile_sz;
    int *s_idx = reinterpret_cast<int *>(s_key + n_pad);
    const int r = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (long long e = tid; e < rows * NCOL; e += blockDim.x) tile[e] = data[e];
    __syncthreads();
    int mode = first_mode;
    long long it = 0;
    for (; it < max_iters; ++it) {
        const RunState &st = A.state[r];
        if (!st.done && st.n_dead + prm.K + 1 > prm.cap) {  // uniform: state written before the last barrier
            if (tid == 0) atomicExch(&ctl->need_grow, 1);
            break;
        }
        const int go = run_update_body(prm, A, n_pad, mode, r, s_key, s_idx);
        mode = 0;
        __syncthreads();  // walker starts, threshold and proposal factor are visible to the walking warps
        if (!go) break;
        const int Kb = A.state[r].Kb;
        if constexpr (NT == 256) {
            if (k1_tables) {  // K = 1: the whole CTA walks the run's walker (uniform branch: barriers inside)
                double *tab = reinterpret_cast<double *>(s_idx + ((n_pad + 1) & ~1));
                if (Kb > 0) cta_walk_k1<OP>(prm, A, prior, tile, (int)rows, (double)rows, cst, r, tab);
                __syncthreads();
                continue;
            }
        }
        if (wid < Kb) warp_walk<OP>(prm, A, prior, tile, (int)rows, (double)rows, cst, r * prm.K + wid, lane);
        __syncthreads();  // chain states are in place for the insert of the next update
    }
    if (tid == 0) atomicMax((unsigned long long *)&ctl->iters, (unsigned long long)it);
}

}  // namespace binest
