// mcmc.cuh — posterior sampler: createMCMCChain / iterateMCMC (BS:630-703) on the GPU operators.
//
// The reference builds ONE adaptive-Metropolis chain on the unnormalised log posterior
//     posteriorDensity(theta) = If[box(theta), logPrior(theta) + logL(theta), logzero]            (BS:630-649)
// with Statistics`MCMC`BuildMarkovChain[{"AdaptiveMetropolis", "Log"}][start, density, {InitialCovariance,
// CovarianceLearnDelay}] (BS:673-696) and advances it with MarkovChainIterate (BS:703).  That sampler is closed
// source; it is restated from Haario, Saksman & Tamminen (2001) exactly as the nested-sampling walk is (walk.cuh,
// DESIGN §2): proposal N(x, C0) while fewer than `delay` states have been seen, N(x, s_d (C_t + eps I)) afterwards,
// s_d = 2.4^2/d, C_t the running sample covariance of all states so far (same recursion as the walk, started at
// t = 1 with mean = start, so that C_t is the unbiased sample covariance); accept iff
// density(x') - density(x) > log u.
//
// Here n_chains independent chains advance in lock-step (1 = the reference): every step scores all proposals with
// ONE batched likelihood launch (loglike_device), and one small kernel (a thread per chain) does accept + Haario
// update + record + next proposal, so nothing returns to the host inside binest_chain_iterate.  Philox counters:
// (pair index, step, chain id) under tags TAG_MC_NORMAL / TAG_MC_ACCEPT — shared with the oracle restatement
// (orc_mcmc_chain), which makes whole trajectories comparable.
#pragma once
#include "problem.cuh"
#include "walk.cuh"

struct binest_chain {
    binest_problem *p = nullptr;
    int d = 0, C = 0, Ps = 0;
    uint64_t seed = 0;
    long long delay = 20, step = 0;
    bool has_prop = false;
    binest::DevBuf<double> x, lp, mean, cov, chol0, prop, prop_lpr, logu, ll;
    binest::DevBuf<long long> t, nacc;
};

namespace binest {

enum : uint32_t { TAG_MC_NORMAL = 9, TAG_MC_ACCEPT = 10 };

struct ChainArrays {
    double *x, *lp, *mean, *cov;  // [C][d], [C], [C][d], [C][d*d]
    const double *chol0;          // [d*d] Cholesky factor of "InitialCovariance" (row-major lower)
    double *prop, *prop_lpr, *logu;  // SoA [d][Ps], [C], [C]
    const double *ll;             // [Ps] log-likelihood of the pending proposals (logzero outside the box)
    long long *t, *nacc;          // states seen so far (start included), accepted moves
};

// accept != 0: finish the pending proposal of every chain (scored in A.ll) and record the new state in out_row
// ([C][d], may be null); propose != 0: draw the next proposal (global step index `step`).
__global__ void __launch_bounds__(128)
mcmc_step_kernel(ChainArrays A, const __grid_constant__ PriorSpec prior, int d, int C, int Ps, uint64_t seed,
                 long long delay, double logzero, long long step, int accept, int propose, double *__restrict__ out_row) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double x[BINEST_MAXD];
    for (int a = 0; a < d; ++a) x[a] = A.x[(size_t)c * d + a];
    long long t = A.t[c];
    if (accept) {
        const double nPr = A.prop_lpr[c], nL = A.ll[c];
        // posteriorDensity BS:630-640: outside the box, or where an operator constraint fails, the density is logzero
        const bool valid = nPr > 0.5 * logzero && nL > 0.5 * logzero;
        const double lpn = valid ? nPr + nL : logzero;
        if (valid && lpn - A.lp[c] > A.logu[c]) {
            for (int a = 0; a < d; ++a) { x[a] = A.prop[(size_t)a * Ps + c]; A.x[(size_t)c * d + a] = x[a]; }
            A.lp[c] = lpn;
            A.nacc[c] += 1;
        }
        // running mean / sample covariance of the visited states (the recursion of the walk, t = states so far)
        const double tf = (double)t;
        double *m = A.mean + (size_t)c * d, *cov = A.cov + (size_t)c * d * d;
        double dm_o[BINEST_MAXD], dm_n[BINEST_MAXD];
        for (int a = 0; a < d; ++a) {
            const double mo = m[a], mn = mo + (x[a] - mo) / (tf + 1.0);
            dm_o[a] = x[a] - mo;
            dm_n[a] = x[a] - mn;
            m[a] = mn;
        }
        const double f = (tf - 1.0) / tf;
        for (int a = 0; a < d; ++a)
            for (int b = 0; b < d; ++b) cov[a * d + b] = f * cov[a * d + b] + dm_o[a] * dm_n[b] / tf;
        t += 1;
        A.t[c] = t;
        if (out_row)
            for (int a = 0; a < d; ++a) out_row[(size_t)c * d + a] = x[a];
    }
    if (propose) {
        double L[BINEST_MAXD * BINEST_MAXD];
        bool own = false;
        if (t >= delay) own = proposal_chol(A.cov + (size_t)c * d * d, d, L) != 0;
        if (!own)
            for (int a = 0; a < d * d; ++a) L[a] = A.chol0[a];
        double z[BINEST_MAXD + 1];
        const uint32_t hi = (uint32_t)((unsigned long long)step >> 32);
        for (int b = 0; b < (d + 1) / 2; ++b)
            rng_normal2(seed, (uint32_t)b, (uint32_t)step, (uint32_t)c, TAG_MC_NORMAL, hi, z[2 * b], z[2 * b + 1]);
        double xn[BINEST_MAXD];
        for (int a = 0; a < d; ++a) {
            double s = x[a];
            for (int b = 0; b <= a; ++b) s += L[a * d + b] * z[b];
            xn[a] = s;
            A.prop[(size_t)a * Ps + c] = s;
        }
        double u0, u1;
        rng_uniform2(seed, 0u, (uint32_t)step, (uint32_t)c, TAG_MC_ACCEPT, hi, u0, u1);
        A.prop_lpr[c] = logprior_dyn(prior, xn, logzero);
        A.logu[c] = log(u0);
    }
}

inline ChainArrays chain_arrays(binest_chain &c) {
    return ChainArrays{c.x.p, c.lp.p, c.mean.p, c.cov.p, c.chol0.p, c.prop.p, c.prop_lpr.p, c.logu.p, c.ll.p, c.t.p, c.nacc.p};
}

}  // namespace binest

extern "C" {

int binest_chain_create(binest_problem *p, const double *start, int64_t n_chains, const double *init_cov,
                        int64_t learn_delay, uint64_t seed, binest_chain **out) {
    using namespace binest;
    return guard([&] {
        BN_REQUIRE(p && start && init_cov && out, BINEST_ERR_TYPE, "null argument");
        BN_REQUIRE(n_chains >= 1 && n_chains <= (1 << 20), BINEST_ERR_DIMENSION, "need 1 <= n_chains <= 2^20");
        BN_REQUIRE(!p->comm, BINEST_ERR_FUNCTION, "the posterior sampler does not run on a data-sharded problem");
        BN_CUDA(cudaSetDevice(p->device));
        const int d = p->d, C = (int)n_chains, Ps = (C + 31) & ~31;
        // Cholesky factor of "InitialCovariance" (symmetrised, BS:705): the proposal until `learn_delay` states exist
        std::vector<double> L((size_t)d * d, 0.0);
        for (int j = 0; j < d; ++j) {
            double s = init_cov[j * d + j];
            for (int k = 0; k < j; ++k) s -= L[j * d + k] * L[j * d + k];
            BN_REQUIRE(s > 0.0 && std::isfinite(s), BINEST_ERR_NUMERICAL, "InitialCovariance is not positive definite");
            const double l = std::sqrt(s);
            L[j * d + j] = l;
            for (int i = j + 1; i < d; ++i) {
                double t = 0.5 * (init_cov[i * d + j] + init_cov[j * d + i]);
                for (int k = 0; k < j; ++k) t -= L[i * d + k] * L[j * d + k];
                L[i * d + j] = t / l;
            }
        }
        auto c = std::make_unique<binest_chain>();
        c->p = p; c->d = d; c->C = C; c->Ps = Ps; c->seed = seed; c->delay = std::max<int64_t>(learn_delay, 2);
        c->x.alloc((size_t)C * d); c->lp.alloc(C); c->mean.alloc((size_t)C * d); c->cov.alloc((size_t)C * d * d);
        c->chol0.alloc((size_t)d * d); c->prop.alloc((size_t)d * Ps); c->prop_lpr.alloc(Ps); c->logu.alloc(Ps);
        c->ll.alloc(Ps); c->t.alloc(C); c->nacc.alloc(C);
        cudaStream_t s = p->stream;
        // density at the starting points: must be a number (createMCMCChain needs a valid start, BS:651-660)
        std::vector<double> ll(C), lpr(C);
        upload_theta(*p, start, C, Ps);
        loglike_device(*p, p->s_theta.p, C, Ps, c->ll.p);
        logprior_kernel<<<(C + 127) / 128, 128, 0, s>>>(p->s_theta.p, C, Ps, p->prior, g_logzero, c->prop_lpr.p);
        BN_LAUNCH_CHECK();
        BN_CUDA(cudaMemcpyAsync(ll.data(), c->ll.p, sizeof(double) * C, cudaMemcpyDeviceToHost, s));
        BN_CUDA(cudaMemcpyAsync(lpr.data(), c->prop_lpr.p, sizeof(double) * C, cudaMemcpyDeviceToHost, s));
        BN_CUDA(cudaStreamSynchronize(s));
        std::vector<long long> ones(C, 1);
        for (int i = 0; i < C; ++i) {
            BN_REQUIRE(ll[i] > 0.5 * g_logzero && lpr[i] > 0.5 * g_logzero && std::isfinite(ll[i] + lpr[i]),
                       BINEST_ERR_BAD_LIKELIHOOD, "starting point outside the support of the posterior density");
            ll[i] += lpr[i];
        }
        BN_CUDA(cudaMemcpyAsync(c->x.p, start, sizeof(double) * C * d, cudaMemcpyHostToDevice, s));
        BN_CUDA(cudaMemcpyAsync(c->mean.p, start, sizeof(double) * C * d, cudaMemcpyHostToDevice, s));
        BN_CUDA(cudaMemcpyAsync(c->lp.p, ll.data(), sizeof(double) * C, cudaMemcpyHostToDevice, s));
        BN_CUDA(cudaMemcpyAsync(c->chol0.p, L.data(), sizeof(double) * d * d, cudaMemcpyHostToDevice, s));
        BN_CUDA(cudaMemcpyAsync(c->t.p, ones.data(), sizeof(long long) * C, cudaMemcpyHostToDevice, s));
        {   // lanes of the SoA proposal block beyond n_chains are scored too: keep them at a harmless value
            std::vector<double> one((size_t)d * Ps, 1.0);
            BN_CUDA(cudaMemcpyAsync(c->prop.p, one.data(), one.size() * sizeof(double), cudaMemcpyHostToDevice, s));
            BN_CUDA(cudaStreamSynchronize(s));
        }
        c->cov.zero(s);
        c->nacc.zero(s);
        BN_CUDA(cudaStreamSynchronize(s));
        *out = c.release();
    });
}

int binest_chain_iterate(binest_chain *c, int64_t n_steps, double *out) {
    using namespace binest;
    return guard([&] {
        BN_REQUIRE(c, BINEST_ERR_TYPE, "null argument");
        if (n_steps <= 0) return;
        binest_problem &p = *c->p;
        BN_CUDA(cudaSetDevice(p.device));
        cudaStream_t s = p.stream;
        const int C = c->C, d = c->d, Ps = c->Ps, grid = (C + 127) / 128;
        DevBuf<double> rec(out ? (size_t)n_steps * C * d : 0);
        ChainArrays A = chain_arrays(*c);
        if (!c->has_prop) {
            mcmc_step_kernel<<<grid, 128, 0, s>>>(A, p.prior, d, C, Ps, c->seed, c->delay, g_logzero, c->step, 0, 1, nullptr);
            BN_LAUNCH_CHECK();
            c->has_prop = true;
        }
        for (int64_t i = 0; i < n_steps; ++i) {
            loglike_device(p, c->prop.p, C, Ps, c->ll.p);  // one batched launch scores every chain's proposal
            c->step += 1;
            mcmc_step_kernel<<<grid, 128, 0, s>>>(A, p.prior, d, C, Ps, c->seed, c->delay, g_logzero, c->step, 1, 1,
                                                  out ? rec.p + (size_t)i * C * d : nullptr);
            BN_LAUNCH_CHECK();
        }
        if (out) BN_CUDA(cudaMemcpyAsync(out, rec.p, rec.n * sizeof(double), cudaMemcpyDeviceToHost, s));
        BN_CUDA(cudaStreamSynchronize(s));
    });
}

int binest_chain_state(binest_chain *c, double *x, double *logdensity, double *mean, double *cov, int64_t *t,
                       int64_t *accepted) {
    using namespace binest;
    return guard([&] {
        BN_REQUIRE(c, BINEST_ERR_TYPE, "null argument");
        BN_CUDA(cudaSetDevice(c->p->device));
        cudaStream_t s = c->p->stream;
        const size_t C = c->C, d = c->d;
        if (x) BN_CUDA(cudaMemcpyAsync(x, c->x.p, sizeof(double) * C * d, cudaMemcpyDeviceToHost, s));
        if (logdensity) BN_CUDA(cudaMemcpyAsync(logdensity, c->lp.p, sizeof(double) * C, cudaMemcpyDeviceToHost, s));
        if (mean) BN_CUDA(cudaMemcpyAsync(mean, c->mean.p, sizeof(double) * C * d, cudaMemcpyDeviceToHost, s));
        if (cov) BN_CUDA(cudaMemcpyAsync(cov, c->cov.p, sizeof(double) * C * d * d, cudaMemcpyDeviceToHost, s));
        if (t) BN_CUDA(cudaMemcpyAsync(t, c->t.p, sizeof(int64_t) * C, cudaMemcpyDeviceToHost, s));
        if (accepted) BN_CUDA(cudaMemcpyAsync(accepted, c->nacc.p, sizeof(int64_t) * C, cudaMemcpyDeviceToHost, s));
        BN_CUDA(cudaStreamSynchronize(s));
    });
}

int binest_chain_free(binest_chain *c) {
    using namespace binest;
    return guard([&] { delete c; });
}

}  // extern "C"
