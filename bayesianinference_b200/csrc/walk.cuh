// walk.cuh — device side of nestedSamplingInternal (BS:859-1040):
//   run_update_kernel   one CTA per run: insert the walkers' points, sort the live set by {logL, point}
//                       (BS:814, 902), crude evidence / entropy / termination test in the log domain
//                       (BS:967-978, 1006-1020), kill the Kb worst (pool sizes n, n-1, ...), covariance blend
//                       (BS:989), proposal Cholesky, RandomChoice of the walkers' starting points (BS:993)
//   walk_step_kernel    one thread per walker: accept/reject the proposal just scored (nsDensity BS:602-617,
//                       Metropolis rule of the "Log" chain BS:720-727), Haario recursion on the chain state,
//                       next proposal x' = x + L z from Philox normals
// The likelihood of all proposals of a step is scored by loglike_stream_kernel (loglike.cuh) in between.
#pragma once
#include "loglike.cuh"

namespace binest {

struct RunState {
    double Lstar;       // likelihood threshold of the current batch (BS:981)
    double logZ;        // crude log evidence (BS:1019)
    double entropy;     // BS:1020
    double logLmax;
    double logXlast;    // logX of the last deleted point
    double logXmin;     // logX of the best live point (BS:929, 935)
    LseAcc dead;        // finalised trapezoid terms of deleted points 1..n_dead-1
    long long iteration;   // 1-based, counts replacements (BS:885, 1021)
    long long n_dead;
    long long walk_base;   // walks started before this batch (Philox counter word 2)
    int done;
    int Kb;             // points replaced by the batch in flight
    int chol_ok;
    int pad_;
    double meanEst[BINEST_MAXD];
    double covEst[BINEST_MAXD * BINEST_MAXD];
    double cholL[BINEST_MAXD * BINEST_MAXD];
};

struct RunParams {
    int d, n, K, R, Ps;
    long long cap;         // dead capacity per run
    long long max_iter, min_iter;
    double log_term_frac;
    double acc_min, acc_max;
    long long S, maxS;
    unsigned long long seed;
    unsigned first_run_id;
    double logzero;
    double loglmax_opt;  // numeric "LogLikelihoodMaximum" (BS:925-932), NaN = use the live set's maximum
    int attempt;  // outer acceptance retry round (BS:1000-1003): offsets the Philox counter word 0 by 16 * attempt
};

struct RunArrays {
    // live set, row-major per run
    double *live_theta;   // [R][n][d]
    double *live_logL, *live_logPr, *live_acc;  // [R][n]
    // dead list
    double *dead_theta;   // [R][cap][d]
    double *dead_logL, *dead_logPr, *dead_acc, *dead_logX;  // [R][cap]
    int *dead_pool;       // [R][cap]
    int *order;           // [R][n] live slots sorted ascending by {logL, point}
    int *kill_slot;       // [R][K] live slots freed by the batch in flight
    RunState *state;      // [R]
    // walkers, SoA with stride Ps; walker w = run * K + j
    double *w_theta;      // [d][Ps] current chain position
    double *w_logL, *w_logPr;
    double *w_prop;       // [d][Ps] proposal being scored
    double *w_prop_logPr;
    double *w_mean;       // [d][Ps]   Haario running mean
    double *w_cov;        // [d*d][Ps] Haario running covariance
    int *w_flags;         // bit0: proposal passes box + prior ratio; bit1: a proposal is in flight; bit2: frozen
    int *w_nacc, *w_steps;
    int *n_unfrozen;      // walkers still stepping (acceptance-range loop, BS:730-736)
};

enum : int { WF_PRE = 1, WF_HASPROP = 2, WF_FROZEN = 4 };

// ---------------------------------------------------------------------------------------------------
// ordering {logL, point, slot}
__device__ __forceinline__ bool sample_less(double la, int ia, double lb, int ib, const double *__restrict__ theta,
                                            int d) {
    if (la < lb) return true;
    if (la > lb) return false;
    if (ia < 0 || ib < 0) return ib < 0 && ia >= 0;  // padding sorts last
    for (int j = 0; j < d; ++j) {
        const double a = theta[(size_t)ia * d + j], b = theta[(size_t)ib * d + j];
        if (a < b) return true;
        if (a > b) return false;
    }
    return ia < ib;
}

// proposal factor chol(s_d (C + eps I)), s_d = 2.4^2/d (Haario et al. 2001); row-major lower triangle
__device__ inline int proposal_chol(const double *cov, int d, double *L) {
    double tr = 0.0;
    for (int a = 0; a < d; ++a) tr += cov[a * d + a];
    const double eps = 1e-10 * (tr / d) + 1e-300;
    const double sd = 2.4 * 2.4 / (double)d;
    for (int a = 0; a < d * d; ++a) L[a] = 0.0;
    for (int j = 0; j < d; ++j) {
        double s = sd * (cov[j * d + j] + eps);
        for (int k = 0; k < j; ++k) s -= L[j * d + k] * L[j * d + k];
        if (!(s > 0.0)) return 0;
        const double l = sqrt(s);
        L[j * d + j] = l;
        for (int i = j + 1; i < d; ++i) {
            double t = sd * 0.5 * (cov[i * d + j] + cov[j * d + i]);  // symmetrizeMatrix BS:705, 716
            for (int k = 0; k < j; ++k) t -= L[i * d + k] * L[j * d + k];
            L[i * d + j] = t / l;
        }
    }
    return 1;
}

// The per-iteration update of ONE run by one CTA (any multiple of 32 threads): shared by run_update_kernel (one launch
// per iteration) and the device-resident loop (walk_loop.cuh).  s_key / s_idx: n_pad doubles / ints of shared memory.
// Returns 1 when a new batch of walks has been set up, 0 when the run is finished / flushed (nothing to walk).
__device__ __forceinline__ int run_update_body(const RunParams &prm, const RunArrays &A, int n_pad, int mode, int r,
                                               double *s_key, int *s_idx) {
    const int first_call = (mode == 1);
    __shared__ double scratch[100];
    __shared__ double scratch_m[33 * 16];
    __shared__ double s_mean[BINEST_MAXD];
    __shared__ double s_cov[BINEST_MAXD * BINEST_MAXD];

    const int tid = threadIdx.x, nt = blockDim.x;
    const int d = prm.d, n = prm.n, K = prm.K;
    RunState &st = A.state[r];
    if (st.done) return 0;
    double *lth = A.live_theta + (size_t)r * n * d;
    double *lL = A.live_logL + (size_t)r * n, *lPr = A.live_logPr + (size_t)r * n, *lAcc = A.live_acc + (size_t)r * n;
    int *order = A.order + (size_t)r * n;
    int *kill = A.kill_slot + (size_t)r * K;
    const size_t dbase = (size_t)r * prm.cap;

    // ---- 1. insert the points the walkers produced (BS:1006-1016) and adopt their chain estimates (BS:999)
    const int Kprev = first_call ? 0 : st.Kb;
    if (Kprev > 0) {
        for (int j = tid; j < Kprev; j += nt) {
            const int w = r * K + j, slot = kill[j];
            for (int a = 0; a < d; ++a) lth[(size_t)slot * d + a] = A.w_theta[(size_t)w * d + a];
            lL[slot] = A.w_logL[w];
            lPr[slot] = A.w_logPr[w];
            lAcc[slot] = (double)A.w_nacc[w] / (double)max(A.w_steps[w], 1);
        }
        if (Kprev <= 32) {
            // few walkers: thread e owns one entry of the mean / covariance estimate and adds the walkers in order —
            // no block reduction (with K = 1, the reference scheme, this is a copy)
            for (int e = tid; e < d + d * d; e += nt) {
                double v = 0.0;
                if (e < d) {
                    for (int j = 0; j < Kprev; ++j) v += A.w_mean[(size_t)(r * K + j) * d + e];
                    st.meanEst[e] = v / (double)Kprev;
                } else {
                    const int a = (e - d) / d, b = (e - d) - a * d;
                    for (int j = 0; j < Kprev; ++j)
                        v += 0.5 * (A.w_cov[(size_t)(r * K + j) * d * d + a * d + b] + A.w_cov[(size_t)(r * K + j) * d * d + b * d + a]);
                    st.covEst[a * d + b] = v / (double)Kprev;
                }
            }
        } else {
            for (int a = 0; a < d; ++a) {
                double v = 0.0;
                for (int j = tid; j < Kprev; j += nt) v += A.w_mean[(size_t)(r * K + j) * d + a];
                v = block_sum(v, scratch);
                if (tid == 0) st.meanEst[a] = v / (double)Kprev;
            }
            for (int a = 0; a < d; ++a)
                for (int b = 0; b <= a; ++b) {
                    double v = 0.0;
                    for (int j = tid; j < Kprev; j += nt)
                        v += 0.5 * (A.w_cov[(size_t)(r * K + j) * d * d + a * d + b] +
                                    A.w_cov[(size_t)(r * K + j) * d * d + b * d + a]);
                    v = block_sum(v, scratch);
                    if (tid == 0) st.covEst[a * d + b] = st.covEst[b * d + a] = v / (double)Kprev;
                }
        }
        if (tid == 0) st.walk_base += Kprev;
    }
    __syncthreads();

    // ---- 2. sort the live set ascending by {logL, point} (bitonic network in shared memory)
    for (int i = tid; i < n_pad; i += nt) {
        s_key[i] = (i < n) ? lL[i] : CUDART_INF;
        s_idx[i] = (i < n) ? i : -1;
    }
    __syncthreads();
    for (int k = 2; k <= n_pad; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < n_pad; i += nt) {
                const int p = i ^ j;
                if (p > i) {
                    const bool up = ((i & k) == 0);
                    const double ka = s_key[i], kb = s_key[p];
                    const int ia = s_idx[i], ib = s_idx[p];
                    const bool lt = sample_less(kb, ib, ka, ia, lth, d);  // partner < me
                    if (lt == up) { s_key[i] = kb; s_key[p] = ka; s_idx[i] = ib; s_idx[p] = ia; }
                }
            }
            __syncthreads();
        }
    for (int i = tid; i < n; i += nt) order[i] = s_idx[i];
    const double logLmax = s_key[n - 1];
    __syncthreads();

    // ---- 3. crude evidence with the current dead list + sorted live set (calculateWeightsCrude BS:812-831)
    const long long D = st.n_dead;
    const double lxD = st.logXlast;  // 0 when D == 0
    const long long M = D + n;
    const double log_half = log(0.5), log_np1 = log((double)n + 1.0);
    LseAcc acc = lse_empty();
    for (int j = tid; j < n; j += nt) {
        const long long k = D + j + 1;  // 1-based index in the full list
        const double lx = (log((double)(n - j)) - log_np1) + lxD;
        double left;
        if (k == 1) left = log_subtract(log(2.0), lx);
        else left = (j == 0) ? lxD : (log((double)(n - j + 1)) - log_np1) + lxD;
        double lw;
        if (k < M) lw = log_half + log_subtract(left, (log((double)(n - j - 1)) - log_np1) + lxD);
        else lw = log_half + log_add(left, lx);
        acc = lse_merge(acc, lse_term(lw + s_key[j], s_key[j]));
    }
    if (tid == 0 && D >= 1) {  // the last deleted point: right neighbour is the worst live point
        const double left = (D == 1) ? log_subtract(log(2.0), lxD) : A.dead_logX[dbase + D - 2];
        const double right = (log((double)n) - log_np1) + lxD;
        const double L = A.dead_logL[dbase + D - 1];
        acc = lse_merge(acc, lse_term(log_half + log_subtract(left, right) + L, L));
    }
    acc = block_lse2(acc, scratch, scratch_m);
    const LseAcc tot = lse_merge(st.dead, acc);
    const double logZ = tot.m + log(tot.s0);
    const double entropy = tot.s1 / tot.s0 - logZ;  // BS:801-810
    const double logXmin = lxD - log_np1;
    __syncthreads();

    // ---- 4. termination test (BS:967-978) in the log domain: X_min L_max <= Z frac
    const long long it = st.iteration;
    const bool go = it <= prm.max_iter &&
                    (it == 1 || it <= prm.min_iter ||
                     !(logXmin + (isnan(prm.loglmax_opt) ? logLmax : prm.loglmax_opt) <= logZ + prm.log_term_frac));
    if (tid == 0) {
        st.logZ = logZ; st.entropy = entropy; st.logLmax = logLmax; st.logXmin = logXmin;
    }
    if (mode == 2) {  // flush for binest_run_fetch on an unfinished run: no kill, no new batch
        if (tid == 0) st.Kb = 0;
        return 0;
    }
    if (!go) {
        if (tid == 0) { st.done = 1; st.Kb = 0; }
        return 0;
    }
    long long Kb_ll = K;
    if (Kb_ll > prm.max_iter - it + 1) Kb_ll = prm.max_iter - it + 1;
    if (Kb_ll > n - 1) Kb_ll = n - 1;
    const int Kb = (int)Kb_ll;

    // ---- 5. covariance of the live set, blend with the running estimate (BS:989), proposal factor.  The d means, then
    //         the d (d + 1) / 2 covariance entries, are reduced 16 at a time (block_sum_multi: three barriers per
    //         round instead of per entry; same reduction tree per value as block_sum)
    {
        double v[16];
        for (int a = 0; a < 16; ++a) v[a] = 0.0;
        for (int i = tid; i < n; i += nt)
            for (int a = 0; a < d; ++a) v[a] += lth[(size_t)i * d + a];
        block_sum_multi(v, d, scratch_m);
        if (tid < d) s_mean[tid] = v[tid] / (double)n;
    }
    __syncthreads();
    {
        const int npair = d * (d + 1) / 2;
        for (int p0 = 0; p0 < npair; p0 += 16) {
            const int np = min(16, npair - p0);
            double v[16];
            for (int k = 0; k < 16; ++k) v[k] = 0.0;
            for (int i = tid; i < n; i += nt) {
                int a = 0, b = 0;
                for (int k = 0; k < p0; ++k) { if (++b > a) { ++a; b = 0; } }  // pair index -> (a, b), b <= a
                for (int k = 0; k < np; ++k) {
                    v[k] += (lth[(size_t)i * d + a] - s_mean[a]) * (lth[(size_t)i * d + b] - s_mean[b]);
                    if (++b > a) { ++a; b = 0; }
                }
            }
            block_sum_multi(v, np, scratch_m);
            if (tid < np) {
                int a = 0, b = 0;
                for (int k = 0; k < p0 + tid; ++k) { if (++b > a) { ++a; b = 0; } }
                s_cov[a * d + b] = s_cov[b * d + a] = v[tid] / (double)(n - 1);
            }
        }
    }
    __syncthreads();
    if (tid == 0) {
        if (first_call) {  // BS:922-923
            for (int a = 0; a < d; ++a) st.meanEst[a] = s_mean[a];
            for (int a = 0; a < d * d; ++a) st.covEst[a] = s_cov[a];
        }
        for (int a = 0; a < d * d; ++a) st.covEst[a] = (st.covEst[a] + s_cov[a]) / 2.0;  // BS:989
        st.chol_ok = proposal_chol(st.covEst, d, st.cholL);
        st.Lstar = s_key[Kb - 1];  // BS:981 (K = 1: Min)
        st.Kb = Kb;
    }

    // ---- 6. kill the Kb worst in order; the j-th removed sees pool size n - j
    for (int j = tid; j < Kb; j += nt) {
        const int slot = s_idx[j];
        const size_t k = dbase + D + j;
        for (int a = 0; a < d; ++a) A.dead_theta[k * d + a] = lth[(size_t)slot * d + a];
        A.dead_logL[k] = lL[slot];
        A.dead_logPr[k] = lPr[slot];
        A.dead_acc[k] = lAcc[slot];
        A.dead_pool[k] = n - j;
        kill[j] = slot;
    }
    if (tid == 0) {  // logX_k = logX_{k-1} - 1/pool_k, sequential like the oracle
        double c = lxD;
        for (int j = 0; j < Kb; ++j) { c -= 1.0 / (double)(n - j); A.dead_logX[dbase + D + j] = c; }
        st.logXlast = c;
    }
    __syncthreads();
    // newly finalised trapezoid terms: 1-based k = max(1, D) .. D + Kb - 1
    LseAcc fin = lse_empty();
    for (int j = tid; j < Kb; j += nt) {
        const long long k = D + j;
        if (k >= 1 && k <= D + Kb - 1) {
            const double left = (k == 1) ? log_subtract(log(2.0), A.dead_logX[dbase]) : A.dead_logX[dbase + k - 2];
            const double right = A.dead_logX[dbase + k];
            const double L = A.dead_logL[dbase + k - 1];
            fin = lse_merge(fin, lse_term(log_half + log_subtract(left, right) + L, L));
        }
    }
    fin = block_lse2(fin, scratch, scratch_m);
    if (tid == 0) {
        st.dead = lse_merge(st.dead, fin);
        st.n_dead = D + Kb;
        st.iteration = it + Kb;
    }

    // ---- 7. start the walkers at random survivors (RandomChoice BS:993)
    for (int j = tid; j < K; j += nt) {
        const int w = r * K + j;
        if (j < Kb) {
            double u0, u1;
            rng_uniform2(prm.seed, 0u, 0u, (uint32_t)(st.walk_base + j), TAG_START, prm.first_run_id + r, u0, u1);
            int pick = Kb + (int)(u0 * (double)(n - Kb));
            if (pick > n - 1) pick = n - 1;
            const int src = s_idx[pick];
            for (int a = 0; a < d; ++a) A.w_theta[(size_t)w * d + a] = lth[(size_t)src * d + a];
            A.w_logL[w] = lL[src];
            A.w_logPr[w] = lPr[src];
            A.w_flags[w] = 0;
        } else {
            A.w_flags[w] = WF_FROZEN;
        }
        A.w_nacc[w] = 0;
        A.w_steps[w] = 0;
    }
    __syncthreads();  // st.meanEst / covEst written by thread 0 above
    for (int j = tid; j < Kb; j += nt) {
        const int w = r * K + j;
        for (int a = 0; a < d; ++a) A.w_mean[(size_t)w * d + a] = st.meanEst[a];
        for (int a = 0; a < d * d; ++a) A.w_cov[(size_t)w * d * d + a] = st.covEst[a];
    }
    if (tid == 0) atomicAdd(A.n_unfrozen, Kb);
    return 1;
}

// one CTA (1024 threads) per run.  Dynamic smem: n_pad doubles + n_pad ints.
__global__ void __launch_bounds__(1024) run_update_kernel(const __grid_constant__ RunParams prm, RunArrays A, int n_pad,
                                                          int mode /* 0 normal, 1 first call, 2 insert only */) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *s_key = reinterpret_cast<double *>(smem_raw);
    int *s_idx = reinterpret_cast<int *>(s_key + n_pad);
    run_update_body(prm, A, n_pad, mode, blockIdx.x, s_key, s_idx);
}

// ---------------------------------------------------------------------------------------------------
// One warp per walker.  Accept phase of the proposal scored by the preceding loglike_stream_kernel (all 32
// lanes sum the per-CTA partials of their walker in a fixed order), then the next proposal.  The scalar
// chain logic is executed redundantly by every lane (same values); only lane 0 stores.
// final_step: accept only.
// walk_step_walker: the step of ONE walker executed by one warp (shared by walk_step_kernel and the persistent
// walk_grid_kernel, walk_grid.cuh).
template <class OP>
__device__ __forceinline__ void walk_step_walker(const RunParams &prm, const RunArrays &A, const PriorSpec &prior,
                                                 const PartialView &pv, double rows, const OpCst &cst, int final_step,
                                                 int w, int lane) {
    constexpr int D = OP::D;
    const int K = prm.K, Ps = prm.Ps;
    if (w >= prm.R * K) return;
    const int r = w / K, j = w - r * K;
    const RunState &st = A.state[r];
    if (st.done || j >= st.Kb) return;
    int flags = A.w_flags[w];
    if (flags & WF_FROZEN) return;
    const bool lead = lane == 0;

    double x[D];
#pragma unroll
    for (int a = 0; a < D; ++a) x[a] = A.w_theta[(size_t)w * D + a];
    double xPr = A.w_logPr[w];
    int steps = A.w_steps[w];
    int nacc = A.w_nacc[w];

    if (flags & WF_HASPROP) {
        double xn[D];
#pragma unroll
        for (int a = 0; a < D; ++a) xn[a] = A.w_prop[(size_t)a * Ps + w];
        bool acc = false;
        if (flags & WF_PRE) {
            const double sum = combine_partials_warp(pv, w, lane);
            const double nL = loglike_finish<OP>(xn, sum, rows, cst, prm.logzero, pv.localized);
            if (nL > st.Lstar) {  // nsDensity: logL > threshold, strict (BS:605)
                acc = true;
                if (lead) A.w_logL[w] = nL;
            }
        }
        if (acc) {
#pragma unroll
            for (int a = 0; a < D; ++a) {
                x[a] = xn[a];
                if (lead) A.w_theta[(size_t)w * D + a] = xn[a];
            }
            xPr = A.w_prop_logPr[w];
            ++nacc;
            if (lead) { A.w_logPr[w] = xPr; A.w_nacc[w] = nacc; }
        }
        // Haario recursion on the chain state, started at t = 10 (BS:715-727); entries spread over the lanes
        const double t = 10.0 + (double)steps;
        double dm_o[D], dm_n[D];  // x - mean_old, x - mean_new
#pragma unroll
        for (int a = 0; a < D; ++a) {
            const double mo = A.w_mean[(size_t)w * D + a];
            const double mn = mo + (x[a] - mo) / (t + 1.0);
            dm_o[a] = x[a] - mo;
            dm_n[a] = x[a] - mn;
            if (lead) A.w_mean[(size_t)w * D + a] = mn;
        }
        {
            double *cov = A.w_cov + (size_t)w * D * D;
            const double f = (t - 1.0) / t;
#pragma unroll
            for (int a = 0; a < D; ++a)
#pragma unroll
                for (int b = 0; b < D; ++b)
                    if (((a * D + b) & 31) == lane) cov[a * D + b] = f * cov[a * D + b] + dm_o[a] * dm_n[b] / t;
        }
        ++steps;
        if (lead) A.w_steps[w] = steps;
        flags &= ~(WF_HASPROP | WF_PRE);
        if (steps % prm.S == 0) {  // BS:730-736: extra S-step blocks until the rate is in range or 5S steps
            const double rate = (double)nacc / (double)steps;
            if ((rate >= prm.acc_min && rate <= prm.acc_max) || steps >= prm.maxS) {
                flags |= WF_FROZEN;
                if (lead) atomicSub(A.n_unfrozen, 1);
            }
        }
    }
    if (!final_step && !(flags & WF_FROZEN)) {
        const uint32_t walk_id = (uint32_t)(st.walk_base + j), run_id = prm.first_run_id + r;
        double z[D + 1];
#pragma unroll
        for (int b = 0; b < (D + 1) / 2; ++b)
            rng_normal2(prm.seed, (uint32_t)(b + 16 * prm.attempt), (uint32_t)steps, walk_id, TAG_NORMAL, run_id, z[2 * b],
                        z[2 * b + 1]);
        double xn[D];
#pragma unroll
        for (int a = 0; a < D; ++a) {
            double s = x[a];
            if (st.chol_ok) {
#pragma unroll
                for (int b = 0; b <= a; ++b) s += st.cholL[a * D + b] * z[b];
            }
            xn[a] = s;
            if (lead) A.w_prop[(size_t)a * Ps + w] = s;
        }
        double u0, u1;
        rng_uniform2(prm.seed, (uint32_t)(16 * prm.attempt), (uint32_t)steps, walk_id, TAG_ACCEPT, run_id, u0, u1);
        int pre = 0;
        if (in_box<D>(prior, xn)) {
            double nPr = 0.0;
#pragma unroll
            for (int a = 0; a < D; ++a) nPr += logprior_dim(prior, a, xn[a]);
            if (!isfinite(nPr)) nPr = prm.logzero;
            if (lead) A.w_prop_logPr[w] = nPr;
            pre = (nPr - xPr > log(u0)) ? WF_PRE : 0;  // Metropolis rule on the log density
        }
        flags |= WF_HASPROP | pre;
    }
    if (lead) A.w_flags[w] = flags;
}

// Outer acceptance retry (BS:995-1004): a walker whose final acceptance rate is outside "MinMaxAcceptanceRate" is
// restarted from a fresh RandomChoice of the survivors (BS:993) with its own chain estimates (BS:999) and — set by
// the host for the whole round — Ceiling[1.25^attempt S] steps.  One thread per walker; counts the restarted
// walkers in n_unfrozen.
static __global__ void walk_retry_kernel(const __grid_constant__ RunParams prm, RunArrays A) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    const int K = prm.K, d = prm.d, n = prm.n;
    if (w >= prm.R * K) return;
    const int r = w / K, j = w - r * K;
    const RunState &st = A.state[r];
    if (st.done || j >= st.Kb) return;
    const int steps = A.w_steps[w];
    if (steps <= 0) return;
    const double rate = (double)A.w_nacc[w] / (double)steps;
    if (rate >= prm.acc_min && rate <= prm.acc_max) return;
    double u0, u1;
    rng_uniform2(prm.seed, (uint32_t)prm.attempt, 0u, (uint32_t)(st.walk_base + j), TAG_START, prm.first_run_id + r, u0, u1);
    int pick = st.Kb + (int)(u0 * (double)(n - st.Kb));
    if (pick > n - 1) pick = n - 1;
    const int src = A.order[(size_t)r * n + pick];
    const double *lth = A.live_theta + ((size_t)r * n + src) * d;
    for (int a = 0; a < d; ++a) A.w_theta[(size_t)w * d + a] = lth[a];
    A.w_logL[w] = A.live_logL[(size_t)r * n + src];
    A.w_logPr[w] = A.live_logPr[(size_t)r * n + src];
    A.w_nacc[w] = 0;
    A.w_steps[w] = 0;
    A.w_flags[w] = 0;
    atomicAdd(A.n_unfrozen, 1);
}

template <class OP>
__global__ void __launch_bounds__(256)
walk_step_kernel(const __grid_constant__ RunParams prm, RunArrays A, const __grid_constant__ PriorSpec prior,
                 const PartialView pv_in, double rows, const OpCst cst, int final_step, const XchgDev *xd = nullptr,
                 int first_step = 0) {
    pdl_wait();               // partials / walker state of the predecessors are complete and visible
    pdl_launch_dependents();  // let the next likelihood kernel start prefetching its data tiles
    PartialView pv = pv_in;
    // sharded modes: wait (in-kernel) for the exchange the preceding producer kernel published; the first step of a
    // block has no proposal in flight and nothing to wait for
    if (xd != nullptr && !first_step && !resolve_exchange(pv, xd)) return;
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    walk_step_walker<OP>(prm, A, prior, pv, rows, cst, final_step, w, lane);
}

}  // namespace binest
