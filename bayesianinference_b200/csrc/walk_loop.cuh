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

// one CTA of NT threads per run: NT = 256 (K <= 8 walkers per iteration) or 1024 (K <= 32)
template <class OP, int NT>
__global__ void __launch_bounds__(NT)
ns_loop_kernel(const __grid_constant__ RunParams prm, RunArrays A, const __grid_constant__ PriorSpec prior,
               const double *__restrict__ data, long long rows, const OpCst cst, int n_pad, int first_mode,
               long long max_iters, LoopCtl *ctl) {
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
        if (wid < Kb) warp_walk<OP>(prm, A, prior, tile, (int)rows, (double)rows, cst, r * prm.K + wid, lane);
        __syncthreads();  // chain states are in place for the insert of the next update
    }
    if (tid == 0) atomicMax((unsigned long long *)&ctl->iters, (unsigned long long)it);
}

}  // namespace binest
