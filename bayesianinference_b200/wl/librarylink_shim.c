/*
 * librarylink_shim.c — LibraryLink entry points over the C ABI of libbinest.so (include/binest.h).
 *
 * Contains NO logic: every function unpacks MArguments, calls one binest_* function and packs the result.
 * It is compiled on the user's machine, where WolframLibrary.h exists (it is not present in the build image):
 *
 *   Needs["CCompilerDriver`"];
 *   CreateLibrary[{"librarylink_shim.c"}, "binestLink",
 *     "IncludeDirectories" -> {"<repo>/include"}, "Libraries" -> {"binest"},
 *     "LibraryDirectories" -> {"<repo>/bayesianinference_b200"}]
 *
 * Handles (binest_problem*, binest_run*) cross the boundary as machine integers.  Status codes 1..6 of
 * binest.h coincide with LIBRARY_TYPE_ERROR .. LIBRARY_FUNCTION_ERROR; 7 (CUDA) and 8 (bad likelihood) are
 * returned as LIBRARY_FUNCTION_ERROR and the host package reads binest_last_error() for the message.
 *
 * Replaces, on the Wolfram side, the compiled functions built at BayesianStatistics.wl:365-595 and the body
 * of nestedSamplingInternal (BayesianStatistics.wl:859-1040); see BayesianInferenceB200.wl.
 */
#include "WolframLibrary.h"
#include "binest.h"

#include <stdint.h>
#include <string.h>

static int st(int code);
static int st_(int code) { return st(code); }
static int st(int code) { return code == 0 ? LIBRARY_NO_ERROR : (code <= 6 ? code : LIBRARY_FUNCTION_ERROR); }

DLLEXPORT mint WolframLibrary_getVersion(void) { return WolframLibraryVersion; }
DLLEXPORT int WolframLibrary_initialize(WolframLibraryData libData) { return LIBRARY_NO_ERROR; }
DLLEXPORT void WolframLibrary_uninitialize(WolframLibraryData libData) {}

/* binestInit[logzero_Real, device_Integer] -> status */
DLLEXPORT int binestInit(WolframLibraryData libData, mint Argc, MArgument *Args, MArgument Res) {
    const int rc = binest_init(MArgument_getReal(Args[0]), (int)MArgument_getInteger(Args[1]));
    MArgument_setInteger(Res, rc);
    return LIBRARY_NO_ERROR;
}

/* binestLastError[] -> "UTF8String" */
DLLEXPORT int binestLastError(WolframLibraryData libData, mint Argc, MArgument *Args, MArgument Res) {
    MArgument_setUTF8String(Res, (char *)binest_last_error());
    return LIBRARY_NO_ERROR;
}

/* binestProblemCreate[op, iparam {Integer,1}, inputs {Real,2}, outputs {Real,2}, priorKind {Integer,1},
 *                     lo {Real,1}, hi {Real,1}, p0 {Real,1}, p1 {Real,1}] -> handle (Integer) */
DLLEXPORT int binestProblemCreate(WolframLibraryData libData, mint Argc, MArgument *Args, MArgument Res) {
    const int op = (int)MArgument_getInteger(Args[0]);
    MTensor ip = MArgument_getMTensor(Args[1]), in = MArgument_getMTensor(Args[2]), out = MArgument_getMTensor(Args[3]);
    MTensor kind = MArgument_getMTensor(Args[4]), lo = MArgument_getMTensor(Args[5]), hi = MArgument_getMTensor(Args[6]);
    MTensor p0 = MArgument_getMTensor(Args[7]), p1 = MArgument_getMTensor(Args[8]);
    if (libData->MTensor_getRank(in) != 2) return LIBRARY_RANK_ERROR;
    const mint *din = libData->MTensor_getDimensions(in), *dout = libData->MTensor_getDimensions(out);
    const mint d = libData->MTensor_getFlattenedLength(kind);
    int64_t iparam[4] = {0, 0, 0, 0};
    int32_t kinds[16];
    const mint *ipd = libData->MTensor_getIntegerData(ip), *kd = libData->MTensor_getIntegerData(kind);
    mint i;
    if (d > 16) return LIBRARY_DIMENSION_ERROR;
    for (i = 0; i < 4 && i < libData->MTensor_getFlattenedLength(ip); ++i) iparam[i] = ipd[i];
    for (i = 0; i < d; ++i) kinds[i] = (int32_t)kd[i];
    const int has_out = libData->MTensor_getFlattenedLength(out) > 0;
    binest_problem *h = 0;
    const int rc = binest_problem_create(op, iparam, libData->MTensor_getRealData(in), din[0], din[1],
                                         has_out ? libData->MTensor_getRealData(out) : 0, has_out ? dout[1] : 0, d, kinds,
                                         libData->MTensor_getRealData(lo), libData->MTensor_getRealData(hi),
                                         libData->MTensor_getRealData(p0), libData->MTensor_getRealData(p1), &h);
    if (rc) return st(rc);
    MArgument_setInteger(Res, (mint)(intptr_t)h);
    return LIBRARY_NO_ERROR;
}

DLLEXPORT int binestProblemFree(WolframLibraryData libData, mint Argc, MArgument *Args, MArgument Res) {
    MArgument_setInteger(Res, binest_problem_free((binest_problem *)(intptr_t)MArgument_getInteger(Args[0])));
    return LIBRARY_NO_ERROR;
}

/* shared shape of binestLogLike / binestLogPrior: [handle, theta {Real,2}] -> {Real,1} */
static int batch_eval(WolframLibraryData libData, MArgument *Args, MArgument Res,
                      int (*fn)(binest_problem *, const double *, int64_t, double *)) {
    binest_problem *h = (binest_problem *)(intptr_t)MArgument_getInteger(Args[0]);
    MTensor th = MArgument_getMTensor(Args[1]), out;
    const mint *dims = libData->MTensor_getDimensions(th);
    mint P = dims[0];
    int rc = libData->MTensor_new(MType_Real, 1, &P, &out);
    if (rc) return rc;
    rc = fn(h, libData->MTensor_getRealData(th), P, libData->MTensor_getRealData(out));
    if (rc) { libData->MTensor_free(out); return st(rc); }
    MArgument_setMTensor(Res, out);
    return LIBRARY_NO_ERROR;
}
DLLEXPORT int binestLogLike(WolframLibraryData libData, mint Argc, MArgument *Args, MArgument Res) {
    return batch_eval(libData, Args, Res, binest_loglike);
}
DLLEXPORT int binestLogPrior(WolframLibraryData libData, mint Argc, MArgument *Args, MArgument Res) {
    return batch_eval(libData, Args, Res, binest_logprior);
}

/* binestChainCreate[problem, start {Real,2} (chains x d), initCov {Real,2}, learnDelay, seed] -> chain handle
 * (createMCMCChain, BayesianStatistics.wl:651-701) */
DLLEXPORT int binestChainCreate(WolframLibraryData libData, mint Argc, MArgument *Args, MArgument Res) {
    binest_problem *h = (binest_problem *)(intptr_t)MArgument_getInteger(Args[0]);
    MTensor st = MArgument_getMTensor(Args[1]), cv = MArgument_getMTensor(Args[2]);
    binest_chain *c = 0;
    int rc;
    if (libData->MTensor_getRank(st) != 2 || libData->MTensor_getRank(cv) != 2) return LIBRARY_RANK_ERROR;
    rc = binest_chain_create(h, libData->MTensor_getRealData(st), libData->MTensor_getDimensions(st)[0],
                             libData->MTensor_getRealData(cv), MArgument_getInteger(Args[3]),
                             (uint64_t)MArgument_getInteger(Args[4]), &c);
    if (rc) return st_(rc);
    MArgument_setInteger(Res, (mint)(intptr_t)c);
    return LIBRARY_NO_ERROR;
}

/* binestChainIterate[chain, n, chains, d] -> {Real,3} (n x chains x d): the state after every step (iterateMCMC, :703) */
DLLEXPORT int binestChainIterate(WolframLibraryData libData, mint Argc, MArgument *Args, MArgument Res) {
    binest_chain *c = (binest_chain *)(intptr_t)MArgument_getInteger(Args[0]);
    mint dims[3];
    MTensor out;
    int rc;
    dims[0] = MArgument_getInteger(Args[1]);
    dims[1] = MArgument_getInteger(Args[2]);
    dims[2] = MArgument_getInteger(Args[3]);
    rc = libData->MTensor_new(MType_Real, 3, dims, &out);
    if (rc) return rc;
    rc = binest_chain_iterate(c, dims[0], libData->MTensor_getRealData(out));
    if (rc) { libData->MTensor_free(out); return st_(rc); }
    MArgument_setMTensor(Res, out);
    return LIBRARY_NO_ERROR;
}

/* binestChainState[chain, chains, d] -> flat {x (chains d), mean (chains d), cov (chains d d), t (chains), accepted (chains)} */
DLLEXPORT int binestChainState(WolframLibraryData libData, mint Argc, MArgument *Args, MArgument Res) {
    binest_chain *c = (binest_chain *)(intptr_t)MArgument_getInteger(Args[0]);
    const mint C = MArgument_getInteger(Args[1]), d = MArgument_getInteger(Args[2]);
    mint len = 2 * C * d + C * d * d + 2 * C, i;
    MTensor out;
    double *o;
    int64_t tt[64], aa[64];
    int rc;
    if (C > 64) return LIBRARY_DIMENSION_ERROR;
    rc = libData->MTensor_new(MType_Real, 1, &len, &out);
    if (rc) return rc;
    o = libData->MTensor_getRealData(out);
    rc = binest_chain_state(c, o, 0, o + C * d, o + 2 * C * d, tt, aa);
    if (rc) { libData->MTensor_free(out); return st_(rc); }
    for (i = 0; i < C; ++i) { o[2 * C * d + C * d * d + i] = (double)tt[i]; o[2 * C * d + C * d * d + C + i] = (double)aa[i]; }
    MArgument_setMTensor(Res, out);
    return LIBRARY_NO_ERROR;
}

DLLEXPORT int binestChainFree(WolframLibraryData libData, mint Argc, MArgument *Args, MArgument Res) {
    MArgument_setInteger(Res, binest_chain_free((binest_chain *)(intptr_t)MArgument_getInteger(Args[0])));
    return LIBRARY_NO_ERROR;
}

/* binestPredictiveComponents[handle, theta {Real,2} (M x d), inputs {Real,2} (Q x F)] -> {Real,3} (M x Q x C)
 * (predictiveDistribution, BayesianStatistics.wl:1437-1483) */
DLLEXPORT int binestPredictiveComponents(WolframLibraryData libData, mint Argc, MArgument *Args, MArgument Res) {
    binest_problem *h = (binest_problem *)(intptr_t)MArgument_getInteger(Args[0]);
    MTensor th = MArgument_getMTensor(Args[1]), xs = MArgument_getMTensor(Args[2]), out;
    mint dims[3];
    int64_t C = 0;
    int rc;
    if (libData->MTensor_getRank(th) != 2 || libData->MTensor_getRank(xs) != 2) return LIBRARY_RANK_ERROR;
    rc = binest_predictive_width(h, &C);
    if (rc) return st(rc);
    if (C <= 0) return LIBRARY_FUNCTION_ERROR;
    dims[0] = libData->MTensor_getDimensions(th)[0];
    dims[1] = libData->MTensor_getDimensions(xs)[0];
    dims[2] = (mint)C;
    rc = libData->MTensor_new(MType_Real, 3, dims, &out);
    if (rc) return rc;
    rc = binest_predictive_components(h, libData->MTensor_getRealData(th), dims[0], libData->MTensor_getRealData(xs),
                                      dims[1], libData->MTensor_getRealData(out));
    if (rc) { libData->MTensor_free(out); return st(rc); }
    MArgument_setMTensor(Res, out);
    return LIBRARY_NO_ERROR;
}

/* binestGPPredict[handle, theta {Real,2} (M x 3), xstar {Real,2} (Q x D)] -> {Real,3}: {means (M x Q), sds (M x Q)}
 * (predictFromGaussianProcess, BayesianGaussianProcess.wl:332-422) */
DLLEXPORT int binestGPPredict(WolframLibraryData libData, mint Argc, MArgument *Args, MArgument Res) {
    binest_problem *h = (binest_problem *)(intptr_t)MArgument_getInteger(Args[0]);
    MTensor th = MArgument_getMTensor(Args[1]), xs = MArgument_getMTensor(Args[2]), out;
    mint dims[3];
    int rc;
    if (libData->MTensor_getRank(th) != 2 || libData->MTensor_getRank(xs) != 2) return LIBRARY_RANK_ERROR;
    dims[0] = 2;
    dims[1] = libData->MTensor_getDimensions(th)[0];
    dims[2] = libData->MTensor_getDimensions(xs)[0];
    rc = libData->MTensor_new(MType_Real, 3, dims, &out);
    if (rc) return rc;
    rc = binest_gp_predict(h, libData->MTensor_getRealData(th), dims[1], libData->MTensor_getRealData(xs), dims[2],
                           libData->MTensor_getRealData(out), libData->MTensor_getRealData(out) + dims[1] * dims[2]);
    if (rc) { libData->MTensor_free(out); return st(rc); }
    MArgument_setMTensor(Res, out);
    return LIBRARY_NO_ERROR;
}

/* binestSamplePrior[handle, n, seed, runId] -> {Real,2} (generateStartingPoints) */
DLLEXPORT int binestSamplePrior(WolframLibraryData libData, mint Argc, MArgument *Args, MArgument Res) {
    binest_problem *h = (binest_problem *)(intptr_t)MArgument_getInteger(Args[0]);
    int64_t d = 0;
    MTensor out;
    mint dims[2];
    int rc = binest_problem_dim(h, &d);
    if (rc) return st(rc);
    dims[0] = MArgument_getInteger(Args[1]);
    dims[1] = (mint)d;
    rc = libData->MTensor_new(MType_Real, 2, dims, &out);
    if (rc) return rc;
    rc = binest_sample_prior(h, dims[0], (uint64_t)MArgument_getInteger(Args[2]), MArgument_getInteger(Args[3]),
                             libData->MTensor_getRealData(out));
    if (rc) { libData->MTensor_free(out); return st(rc); }
    MArgument_setMTensor(Res, out);
    return LIBRARY_NO_ERROR;
}

/* binestRunCreate[handle, iopts {Integer,1} (pool, K, S, maxIter, minIter, seed, firstRun, nRuns),
 *                 ropts {Real,1} (termFrac, accMin, accMax, logLmax; "LogLikelihoodMaximum" -> Automatic is sent as
 *                 1e308 and stays NaN here), start {Real,3} or {}] -> run handle */
DLLEXPORT int binestRunCreate(WolframLibraryData libData, mint Argc, MArgument *Args, MArgument Res) {
    binest_problem *h = (binest_problem *)(intptr_t)MArgument_getInteger(Args[0]);
    const mint *io = libData->MTensor_getIntegerData(MArgument_getMTensor(Args[1]));
    const double *ro = libData->MTensor_getRealData(MArgument_getMTensor(Args[2]));
    MTensor sp = MArgument_getMTensor(Args[3]);
    binest_options o;
    binest_run *r = 0;
    int rc;
    binest_default_options(&o);
    o.pool_size = io[0]; o.batch_k = io[1]; o.mc_steps = io[2]; o.max_iter = io[3]; o.min_iter = io[4];
    o.seed = (uint64_t)io[5]; o.first_run_id = io[6]; o.n_runs = io[7];
    o.term_frac = ro[0]; o.acc_min = ro[1]; o.acc_max = ro[2];
    if (ro[3] < 1e307) o.loglmax = ro[3]; /* else keep NaN = Automatic (a WL packed array cannot carry NaN) */
    rc = binest_run_create(h, &o, libData->MTensor_getFlattenedLength(sp) > 0 ? libData->MTensor_getRealData(sp) : 0, &r);
    if (rc) return st(rc);
    MArgument_setInteger(Res, (mint)(intptr_t)r);
    return LIBRARY_NO_ERROR;
}

/* binestRunAdvance[run, maxBatches] -> 1 if finished */
DLLEXPORT int binestRunAdvance(WolframLibraryData libData, mint Argc, MArgument *Args, MArgument Res) {
    int32_t fin = 0;
    const int rc = binest_run_advance((binest_run *)(intptr_t)MArgument_getInteger(Args[0]), MArgument_getInteger(Args[1]), &fin);
    if (rc) return st(rc);
    MArgument_setInteger(Res, fin);
    return LIBRARY_NO_ERROR;
}

/* binestRunFetch[run, runIndex] -> {Real,2}: M x (d + 7) rows (point, logL, logPrior, acc, pool, logX, crudeLogW, pad);
 * the last row carries the summary {CrudeLogEvidence, CrudeRelativeEntropy, LogLikelihoodMaximum, LogEstimatedMissing} */
DLLEXPORT int binestRunFetch(WolframLibraryData libData, mint Argc, MArgument *Args, MArgument Res) {
    binest_run *r = (binest_run *)(intptr_t)MArgument_getInteger(Args[0]);
    const mint run = MArgument_getInteger(Args[1]);
    int64_t M = 0, nd = 0, it = 0, ev = 0, d = (int64_t)MArgument_getInteger(Args[2]);
    MTensor pts, cols, pool;
    mint dims[2];
    int rc = binest_run_fetch(r, run, 0, 0, 0, 0, 0, 0, 0, 0); /* flush an unfinished batch */
    if (rc) return st(rc);
    rc = binest_run_sizes(r, run, &M, &nd, &it, &ev);
    if (rc) return st(rc);
    dims[0] = (mint)M + 1; dims[1] = (mint)d + 6;
    rc = libData->MTensor_new(MType_Real, 2, dims, &pts);
    if (rc) return rc;
    {
        /* column blocks are fetched into temporaries and interleaved row-wise; no arithmetic */
        double *o = libData->MTensor_getRealData(pts);
        mint one = (mint)M;
        double *P, *L, *Pr, *A, *X, *W, summary[4];
        int64_t *pl;
        mint k, j;
        dims[0] = (mint)M; dims[1] = (mint)d;
        rc = libData->MTensor_new(MType_Real, 2, dims, &cols);
        if (rc) return rc;
        P = libData->MTensor_getRealData(cols);
        rc = libData->MTensor_new(MType_Integer, 1, &one, &pool);
        if (rc) return rc;
        pl = (int64_t *)libData->MTensor_getIntegerData(pool);
        L = o; /* reuse the tail of the output as scratch is not possible row-wise: use five more temporaries */
        {
            MTensor t[5];
            int q;
            for (q = 0; q < 5; ++q) { rc = libData->MTensor_new(MType_Real, 1, &one, &t[q]); if (rc) return rc; }
            L = libData->MTensor_getRealData(t[0]); Pr = libData->MTensor_getRealData(t[1]);
            A = libData->MTensor_getRealData(t[2]); X = libData->MTensor_getRealData(t[3]); W = libData->MTensor_getRealData(t[4]);
            rc = binest_run_fetch(r, run, P, L, Pr, A, pl, X, W, summary);
            if (!rc)
                for (k = 0; k < (mint)M; ++k) {
                    double *row = o + k * (d + 6);
                    for (j = 0; j < (mint)d; ++j) row[j] = P[k * d + j];
                    row[d] = L[k]; row[d + 1] = Pr[k]; row[d + 2] = A[k]; row[d + 3] = (double)pl[k];
                    row[d + 4] = X[k]; row[d + 5] = W[k];
                }
            for (q = 0; q < 5; ++q) libData->MTensor_free(t[q]);
        }
        if (!rc) {
            double *row = o + (mint)M * (d + 6);
            memset(row, 0, sizeof(double) * (size_t)(d + 6));
            row[0] = summary[0]; row[1] = summary[1]; row[2] = summary[2]; row[3] = summary[3];
            row[4] = (double)nd; row[5] = (double)it;
        }
        libData->MTensor_free(cols);
        libData->MTensor_free(pool);
    }
    if (rc) { libData->MTensor_free(pts); return st(rc); }
    MArgument_setMTensor(Res, pts);
    return LIBRARY_NO_ERROR;
}

DLLEXPORT int binestRunFree(WolframLibraryData libData, mint Argc, MArgument *Args, MArgument Res) {
    MArgument_setInteger(Res, binest_run_free((binest_run *)(intptr_t)MArgument_getInteger(Args[0])));
    return LIBRARY_NO_ERROR;
}

/* binestEvidenceSampling[points {Real,2}, logL {Real,1}, pool {Integer,1}, nLive, postRuns, seed]
 *   -> {Real,2}: rows = {z (R), H (R), pmean (R x d flattened), logwMean (M), logwSd (M), slxMean (M), slxSd (M)} packed
 *      into one flat vector; the host package partitions it. */
DLLEXPORT int binestEvidenceSampling(WolframLibraryData libData, mint Argc, MArgument *Args, MArgument Res) {
    MTensor pts = MArgument_getMTensor(Args[0]), lL = MArgument_getMTensor(Args[1]), pool = MArgument_getMTensor(Args[2]);
    const mint *dims = libData->MTensor_getDimensions(pts);
    const mint M = dims[0], d = dims[1], n = MArgument_getInteger(Args[3]), R = MArgument_getInteger(Args[4]);
    mint len = 2 * R + R * d + 4 * M;
    MTensor out;
    double *o;
    int rc = libData->MTensor_new(MType_Real, 1, &len, &out);
    if (rc) return rc;
    o = libData->MTensor_getRealData(out);
    rc = binest_evidence_sampling(M, d, libData->MTensor_getRealData(pts), libData->MTensor_getRealData(lL),
                                  (const int64_t *)libData->MTensor_getIntegerData(pool), n, R,
                                  (uint64_t)MArgument_getInteger(Args[5]), o, o + 2 * R + R * d, o + 2 * R + R * d + M,
                                  o + 2 * R + R * d + 2 * M, o + 2 * R + R * d + 3 * M, o + 2 * R, o + R);
    if (rc) { libData->MTensor_free(out); return st(rc); }
    MArgument_setMTensor(Res, out);
    return LIBRARY_NO_ERROR;
}

/* binestCrudeWeights[logL {Real,1}, pool {Integer,1}, nLive] -> flat {logX (M), crudeLogW (M), summary (4)} */
DLLEXPORT int binestCrudeWeights(WolframLibraryData libData, mint Argc, MArgument *Args, MArgument Res) {
    MTensor lL = MArgument_getMTensor(Args[0]), pool = MArgument_getMTensor(Args[1]);
    const mint M = libData->MTensor_getFlattenedLength(lL);
    mint len = 2 * M + 4;
    MTensor out;
    double *o;
    int rc = libData->MTensor_new(MType_Real, 1, &len, &out);
    if (rc) return rc;
    o = libData->MTensor_getRealData(out);
    rc = binest_crude_weights(M, libData->MTensor_getRealData(lL), (const int64_t *)libData->MTensor_getIntegerData(pool),
                              MArgument_getInteger(Args[2]), o, o + M, o + 2 * M);
    if (rc) { libData->MTensor_free(out); return st(rc); }
    MArgument_setMTensor(Res, out);
    return LIBRARY_NO_ERROR;
}

/* binestRunCombine[run, d, scheme, postRuns, seed] -> {Real,2}: combineRuns -> evidenceSampling of the group's runs on
 * the device (binest_run_combine).  (M + R + 1) x (d + 13) rows:
 *   rows 1..M      point (d), the BINEST_NCOL_F double columns in the order of binest.h, PoolSize, RunIndex
 *   rows M+1..M+R  per Monte-Carlo draw: z, H, parameter means (d)
 *   last row       summary (4), live block, M */
DLLEXPORT int binestRunCombine(WolframLibraryData libData, mint Argc, MArgument *Args, MArgument Res) {
    binest_run *r = (binest_run *)(intptr_t)MArgument_getInteger(Args[0]);
    const mint d = MArgument_getInteger(Args[1]), R = MArgument_getInteger(Args[3]);
    const mint W = d + BINEST_NCOL_F + 2;
    int64_t Mt = 0, M = 0, nlive = 0;
    MTensor out, tab, itab, pts, zz;
    mint dims[2], len, k, j;
    double *o, *T, *P, *Z, summary[4];
    int64_t *I;
    int rc = binest_run_merge_size(r, &Mt);
    if (rc) return st(rc);
    len = (mint)Mt * BINEST_NCOL_F;
    rc = libData->MTensor_new(MType_Real, 1, &len, &tab);
    if (rc) return rc;
    len = (mint)Mt * 2;
    rc = libData->MTensor_new(MType_Integer, 1, &len, &itab);
    if (rc) return rc;
    len = (mint)Mt * d;
    rc = libData->MTensor_new(MType_Real, 1, &len, &pts);
    if (rc) return rc;
    len = R * (d + 2);
    rc = libData->MTensor_new(MType_Real, 1, &len, &zz);
    if (rc) return rc;
    T = libData->MTensor_getRealData(tab); P = libData->MTensor_getRealData(pts); Z = libData->MTensor_getRealData(zz);
    I = (int64_t *)libData->MTensor_getIntegerData(itab);
    rc = binest_run_combine(r, (int32_t)MArgument_getInteger(Args[2]), R, (uint64_t)MArgument_getInteger(Args[4]), P, T, I,
                            Z, Z + 2 * R, Z + R, summary, &M, &nlive);
    if (!rc) {
        dims[0] = (mint)M + R + 1; dims[1] = W;
        rc = libData->MTensor_new(MType_Real, 2, dims, &out);
    }
    if (!rc) {
        o = libData->MTensor_getRealData(out);
        memset(o, 0, sizeof(double) * (size_t)(dims[0] * dims[1]));
        for (k = 0; k < (mint)M; ++k) { /* column blocks interleaved row-wise; no arithmetic */
            double *row = o + k * W;
            for (j = 0; j < d; ++j) row[j] = P[k * d + j];
            for (j = 0; j < BINEST_NCOL_F; ++j) row[d + j] = T[j * (mint)Mt + k];
            row[d + BINEST_NCOL_F] = (double)I[k];
            row[d + BINEST_NCOL_F + 1] = (double)I[(mint)Mt + k];
        }
        for (k = 0; k < R; ++k) {
            double *row = o + ((mint)M + k) * W;
            row[0] = Z[k]; row[1] = Z[R + k];
            for (j = 0; j < d; ++j) row[2 + j] = Z[2 * R + k * d + j];
        }
        {
            double *row = o + ((mint)M + R) * W;
            row[0] = summary[0]; row[1] = summary[1]; row[2] = summary[2]; row[3] = summary[3];
            row[4] = (double)nlive; row[5] = (double)M;
        }
    }
    libData->MTensor_free(tab); libData->MTensor_free(itab); libData->MTensor_free(pts); libData->MTensor_free(zz);
    if (rc) return st(rc);
    MArgument_setMTensor(Res, out);
    return LIBRARY_NO_ERROR;
}
