/*
 * binest.h — C ABI of the B200-native nested-sampling engine (libbinest.so).
 *
 * This is the drop-in boundary for the hot path of ssmit1986/BayesianInference: the library
 * replaces the *body* of nestedSamplingInternal (BayesianStatistics.wl:859-1040) and the
 * operators it calls, and is reached from the Wolfram Language host package only through the
 * LibraryLink shim (bayesianinference_b200/wl/librarylink_shim.c).  Plain pointers and sizes
 * only: row-major fp64 arrays, 64-bit integer counts, opaque handles, int status codes.
 * All array arguments are HOST pointers unless a name ends in _dev.
 *
 * Citations: BS = BayesianInference/Kernel/BayesianStatistics.wl, BU = BayesianUtilities.wl,
 * GP = BayesianGaussianProcess.wl (reference checkout).
 *
 * Error convention (BS:422-425, BU:47): operators never fail — a numerically impossible
 * parameter vector evaluates to `logzero`.  API misuse / CUDA failures return a non-zero
 * status; binest_last_error() gives the message (thread-local).  There is NO CPU fallback:
 * every compute entry point fails with BINEST_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef BINEST_H
#define BINEST_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BINEST_VERSION 100

/* status codes; 1..6 coincide with LibraryLink's LIBRARY_*_ERROR so the shim passes them through */
enum {
    BINEST_OK = 0,
    BINEST_ERR_TYPE = 1,
    BINEST_ERR_RANK = 2,
    BINEST_ERR_DIMENSION = 3,
    BINEST_ERR_NUMERICAL = 4,
    BINEST_ERR_MEMORY = 5,
    BINEST_ERR_FUNCTION = 6,
    BINEST_ERR_CUDA = 7,
    BINEST_ERR_BAD_LIKELIHOOD = 8 /* "Bad likelihood function", BS:917-921 */
};

/* Fixed operator table.  Replaces the symbolic builders logLikelihoodFunction (BS:429-505),
 * regressionLogLikelihoodFunction (BS:517-595) and the GP wiring (GP:161-199, 296-307). */
enum {
    BINEST_OP_GAUSSIAN_IID = 1, /* NormalDistribution[mu, sigma]; theta = (mu, sigma)                  */
    BINEST_OP_POLYREG = 2,      /* NormalDistribution[Sum_j c_j x^j, sigma]; theta = (c_0..c_deg, sigma)
                                   iparam[0] = degree (1..5)                                          */
    BINEST_OP_LOGISTIC = 3,     /* softmax, reference class K (z_K = 0); theta = K-1 blocks (w_1..w_F, b)
                                   iparam[1] = K (2..3); outputs = class index 0..K-1 as fp64          */
    BINEST_OP_GBM = 4,          /* GeometricBrownianMotionProcess[mu, sigma, x0] on (t_i, x_i); theta = (mu, sigma);
                                   TemporalData adaptor BS:511-515: inputs = times, outputs = values   */
    BINEST_OP_GP_SE = 5         /* GP marginal likelihood, squared-exponential kernel + nugget;
                                   theta = (sigma_f, ell, sigma_n)                                     */
};

/* Prior kinds per parameter: ignorancePrior BS:25-64 / logPDFFunction BS:365-427. */
enum {
    BINEST_PRIOR_UNIFORM = 1,     /* "LocationParameter": UniformDistribution[{lo, hi}]      BS:37-39 */
    BINEST_PRIOR_SCALE = 2,       /* "ScaleParameter": 1/x normalised on [lo, hi]            BS:42-48 */
    BINEST_PRIOR_NORMAL_TRUNC = 3 /* NormalDistribution[p0, p1] truncated to the box         BS:51-59 */
};

typedef struct binest_problem binest_problem;
typedef struct binest_run binest_run;

/* Options of nestedSampling (BS:837-851) + evidenceSampling (BS:833-836), flattened. */
typedef struct binest_options {
    int64_t pool_size;   /* "SamplePoolSize"        default 100    BS:839 */
    int64_t batch_k;     /* live points replaced per iteration; 1 = the reference scheme BS:980-1018 */
    int64_t mc_steps;    /* "MonteCarloSteps"       default 200    BS:844 ({S, S, 5S} BS:872) */
    int64_t max_iter;    /* "MaxIterations"         default 10000  BS:841 */
    int64_t min_iter;    /* "MinIterations"         default 100    BS:842 */
    double term_frac;    /* "TerminationFraction"   default 0.01   BS:845 */
    double acc_min;      /* "MinMaxAcceptanceRate"  default {0,1}  BS:848 */
    double acc_max;
    uint64_t seed;       /* Philox key; results depend on (seed, run id) only, not on the GPU count */
    int64_t first_run_id;/* id of the first run of this group (run-sharding across ranks/GPUs) */
    int64_t n_runs;      /* independent runs advanced in lock-step on this device ("ParallelRuns" BS:1369) */
    double loglmax;      /* "LogLikelihoodMaximum" BS:847: a number replaces the running maximum of the live set in the
                            termination estimate X_min * L_max (BS:925-932); NaN = Automatic (the default) */
} binest_options;

/* ---- library ----------------------------------------------------------------------------- */
int binest_version(void);
const char *binest_last_error(void);
/* logzero = $MachineLogZero of the host (BU:47); device < 0 keeps the current CUDA device. */
int binest_init(double logzero, int device);
int binest_device_count(int *count);
void binest_default_options(binest_options *opts);
/* Measured fp64 FMA throughput of the current device (register-resident DFMA loop on all SMs),
 * in TFLOP/s; the roofline denominator SURVEY.md §8d asks the build to measure. */
int binest_measure_fp64_peak(double *tflops, double *ms);
/* Number of kernels this library has launched since load (bench.py's gpu_launches). */
int64_t binest_launch_count(void);

/* ---- problem definition: the data-carrying half of defineInferenceProblem (BS:167-307) ---- */
/* inputs: n_rows x n_in (i.i.d. data: the data matrix itself, BS:492); outputs: n_rows x n_out or NULL.
 * Data are uploaded once and stay device-resident (the reference bakes them into the compiled
 * function, BS:488-504).  lo/hi: open parameter box (BS:327-336); prior_p0/p1 may be NULL. */
int binest_problem_create(int op_id, const int64_t *iparam /*[4]*/, const double *inputs, int64_t n_rows,
                          int64_t n_in, const double *outputs, int64_t n_out, int64_t d,
                          const int32_t *prior_kind, const double *lo, const double *hi,
                          const double *prior_p0, const double *prior_p1, binest_problem **out);
int binest_problem_free(binest_problem *p);
int binest_problem_dim(const binest_problem *p, int64_t *d);

/* "LogLikelihoodFunction" (Listable, BS:499): theta P x d -> out[P].  Box / operator constraints
 * violated -> logzero (BS:491-494, 580-583). */
int binest_loglike(binest_problem *p, const double *theta, int64_t P, double *out);
/* predictFromGaussianProcess (GP:332-422, predictFromGaussianProcessInternal GP:395-420) for a BINEST_OP_GP_SE
 * problem (squared-exponential kernel + nugget, zero mean function): for every parameter vector theta_m (the
 * samples' points, M x 3) and every prediction input x*_q (Q x n_in, row-major)
 *   mean[m*Q + q] = k_q . K^-1 y,   sd[m*Q + q] = Sqrt[kappa - k_q . K^-1 k_q],
 *   k_q = (k(x_i, x*_q))_i,  kappa = k(x*, x*) + nugget = sf^2 + sn^2  (compiledKandKappa GP:92-116)
 * — the parameters of the NormalDistribution components the reference mixes with the samples' weights (GP:357).
 * A covariance matrix that cannot be factored gives NaN for that theta. */
int binest_gp_predict(binest_problem *p, const double *theta, int64_t M, const double *xstar, int64_t Q,
                      double *mean, double *sd);
/* predictiveDistribution for regression problems (BS:1437-1483): the parameters of the mixture components
 * dist[theta_m, x_q] ("GeneratingDistribution" with sample m's parameters at input q, BS:1450-1462) for every
 * sample (theta M x d) and input (Q x n_in): out[(m*Q + q)*C + c] with
 *   BINEST_OP_POLYREG:  C = 2, (mean, sd) of NormalDistribution[Sum_j c_j x^j, sigma]
 *   BINEST_OP_LOGISTIC: C = K, class probabilities (softmax with reference class K)
 * binest_predictive_width returns C (0 for operators without independent variables, whose mixture components
 * are the samples' parameters themselves, BS:1421-1435). */
int binest_predictive_width(const binest_problem *p, int64_t *n_comp);
int binest_predictive_components(binest_problem *p, const double *theta, int64_t M, const double *inputs,
                                 int64_t Q, double *out);
/* "LogPriorPDFFunction" (BS:410-426). */
int binest_logprior(binest_problem *p, const double *theta, int64_t P, double *out);
/* generateStartingPoints (BS:1055-1068): n i.i.d. prior draws, out n x d. */
int binest_sample_prior(binest_problem *p, int64_t n, uint64_t seed, int64_t run_id, double *out);

/* ---- the engine: nestedSamplingInternal (BS:859-1040) for n_runs lock-step runs ----------- */
/* start_points: n_runs x pool_size x d, or NULL to draw them from the prior (each run its own,
 * BS:1320-1332).  Returns BINEST_ERR_BAD_LIKELIHOOD when an initial logL is not finite (BS:917-921). */
int binest_run_create(binest_problem *p, const binest_options *opts, const double *start_points,
                      binest_run **out);
/* Advance by at most max_batches iterations (each replaces batch_k points per run); <= 0: to
 * termination (BS:967-978).  *finished = 1 when every run has terminated. */
int binest_run_advance(binest_run *r, int64_t max_batches, int32_t *finished);
/* per run: total samples M (dead + live), deleted count, iterations (= replacements), and the
 * number of likelihood evaluations spent so far (whole group). */
int binest_run_sizes(binest_run *r, int64_t run, int64_t *M, int64_t *n_deleted, int64_t *iterations,
                     int64_t *evals);
/* Sorted sample list of one run (SortBy {logL, point}, BS:814): any output may be NULL.
 * points M x d; acc = NaN for initial samples (Missing["InitialSample"], BS:911); pool = pool size
 * at each sample's removal; logX, crude_logw = calculateWeightsCrude (BS:812-831);
 * summary[4] = {CrudeLogEvidence, CrudeRelativeEntropy, LogLikelihoodMaximum, LogEstimatedMissingEvidence}
 * (BS:1183-1194). */
int binest_run_fetch(binest_run *r, int64_t run, double *points, double *logL, double *logPrior,
                     double *acc, int64_t *pool, double *logX, double *crude_logw, double *summary);
/* walker-level chain estimates of the last iteration: MeanEstimate d, CovarianceEstimate d x d (BS:999) */
int binest_run_estimates(binest_run *r, int64_t run, double *mean, double *cov);
/* device time spent in the walk graphs so far (CUDA events on the run's stream), graphs launched, batches */
int binest_run_timing(binest_run *r, double *walk_ms, int64_t *walk_graphs, int64_t *batches);
/* which device implementation walks this run (measurement / tests): */
enum binest_walk_path {
    BINEST_WALK_STEPPED_GRAPH = 0,   /* CUDA graph of [walk_step, loglike_stream] x S                     */
    BINEST_WALK_CLUSTER_RESIDENT = 1,/* one launch per walk, data in the shared memories of a CTA cluster */
    BINEST_WALK_GRID_RESIDENT = 2,   /* one persistent cooperative launch per walk, data in all SMs' smem  */
    BINEST_WALK_STEPPED_SHARDED = 3, /* data-sharded: stepped with an all-gather per step                 */
    BINEST_WALK_STEPPED_GP = 4,      /* GP operator: stepped, hundreds of launches per likelihood          */
    BINEST_WALK_DEVICE_LOOP = 5      /* tiny problems: the whole nested-sampling loop in one launch, warp per walker */
};
int binest_run_path(const binest_run *r, int *path);
/* the CUDA stream (cudaStream_t) every kernel of this problem and of its runs is launched on, so that callers
 * can record their own events / order their own work on it */
int binest_problem_stream(const binest_problem *p, void **stream);
int binest_run_free(binest_run *r);

/* ---- posterior sampler: createMCMCChain / iterateMCMC (BS:630-703) ------------------------ */
/* n_chains adaptive-Metropolis chains (1 = the reference) on the unnormalised log posterior
 * posteriorDensity = If[box, logPrior + logL, logzero] (BS:630-649), advanced in lock-step: one batched
 * likelihood launch scores all proposals of a step.  start: n_chains x d; init_cov: d x d "InitialCovariance"
 * (BS:676-683, already expanded to a matrix by the host); learn_delay: "CovarianceLearnDelay" (BS:684-690,
 * default 20): the proposal covariance is init_cov until that many states exist, then 2.4^2/d times the running
 * sample covariance (Haario et al. 2001 — the Wolfram sampler is closed source).
 * BINEST_ERR_BAD_LIKELIHOOD: a starting point outside the support (createMCMCChain::start, BS:651-655). */
typedef struct binest_chain binest_chain;
int binest_chain_create(binest_problem *p, const double *start, int64_t n_chains, const double *init_cov,
                        int64_t learn_delay, uint64_t seed, binest_chain **out);
/* iterateMCMC[chain, n] (BS:703): n steps of every chain; out (may be NULL): n_steps x n_chains x d, the state
 * after each step.  Thinning / burn-in ({n, thin} forms) are slices the host takes. */
int binest_chain_iterate(binest_chain *c, int64_t n_steps, double *out);
/* chain["StateData"] = {x, t, mean, cov} and the accepted-move count behind chain["AcceptanceRate"];
 * any output may be NULL.  x, mean: n_chains x d; cov: n_chains x d x d; t = states seen (start included). */
int binest_chain_state(binest_chain *c, double *x, double *logdensity, double *mean, double *cov, int64_t *t,
                       int64_t *accepted);
int binest_chain_free(binest_chain *c);

/* ---- evidenceSampling (BS:1158-1291) on a sorted sample list ------------------------------ */
/* pool: per-sample pool sizes (first M - n_live entries used); outputs may be NULL:
 *  z[post_runs]; logw_mean/sd[M] (LogPosteriorWeight); slx_mean/sd[M] (SampledLogX);
 *  pmean[post_runs x d] (parameter means per draw); H[post_runs] (relative entropy per draw). */
int binest_evidence_sampling(int64_t M, int64_t d, const double *points, const double *logL,
                             const int64_t *pool, int64_t n_live, int64_t post_runs, uint64_t seed,
                             double *z, double *logw_mean, double *logw_sd, double *slx_mean,
                             double *slx_sd, double *pmean, double *H);
/* calculateWeightsCrude + logSumExp + calculateEntropy (BS:812-831, BU:318-335, BS:801-810). */
int binest_crude_weights(int64_t M, const double *logL, const int64_t *pool, int64_t n_live,
                         double *logX, double *crude_logw, double *summary /*[4]*/);

/* ---- combineRuns (BS:1293-1315) on the device ------------------------------------------------ */
/* Join of R runs' sample lists, DeleteDuplicatesBy Point (the first in Join order kept), SortBy {logL, point}
 * (BS:1293-1297), plus what the re-weighting of the merged list needs: the summed pool size at every sample's
 * likelihood level (the runs' pool sizes are step functions of the level) and the index of the run a sample came from.
 * Inputs: the runs concatenated in Join order — sizes[R]; points Mtot x d; logL, logPrior (may be NULL), acc (may be
 * NULL) [Mtot]; pool [Mtot] = each run's own pool size at its samples; run_id [Mtot] or NULL (= position in the Join).
 * Every run must be sorted by {logL, point}, the order binest_run_fetch returns (BINEST_ERR_DIMENSION otherwise).
 * Outputs (caller allocates Mtot entries, *M_out <= Mtot are written; any may be NULL).  *live_block = length of the
 * tail whose summed pool sizes are M, M-1, .., 1 counted from its start — the part that acts as a final live set. */
int binest_merge_runs(int64_t R, const int64_t *sizes, int64_t d, const double *points, const double *logL,
                      const double *logPrior, const double *acc, const int64_t *pool, const int64_t *run_id,
                      double *points_out, double *logL_out, double *logPrior_out, double *acc_out, int64_t *pool_out,
                      int64_t *run_out, int64_t *M_out, int64_t *live_block);
/* The whole of combineRuns -> evidenceSampling (BS:1293-1315, 1158-1291) in one call: the merge above, the X sequence
 * of the merged list (scheme), calculateWeightsCrude, the Monte-Carlo evidence error, and the final
 * SortBy[-CrudePosteriorWeight] (BS:1241, stable) applied to every column.
 *   BINEST_MERGE_REFERENCE: pool size n_tot = Total[SamplePoolSize] for the first M - n_tot samples, then n_tot..1
 *                           (BS:1307-1309 feeding BS:785-799);
 *   BINEST_MERGE_POOLSIZES: the summed pool sizes to the end, live block as returned by binest_merge_runs.
 * table_out: BINEST_NCOL_F columns of stride Mtot (column c at table_out + c*Mtot), itable_out: PoolSize, RunIndex
 * (stride Mtot); points_out M x d; z, H [post_runs]; pmean [post_runs x d]; summary[4] as binest_crude_weights. */
enum { BINEST_MERGE_REFERENCE = 0, BINEST_MERGE_POOLSIZES = 1 };
enum {
    BINEST_COL_LOGL = 0, BINEST_COL_LOGPRIOR, BINEST_COL_ACC, BINEST_COL_LOGX, BINEST_COL_X, BINEST_COL_CRUDE_LOGW,
    BINEST_COL_CRUDE_W, BINEST_COL_SLX_MEAN, BINEST_COL_SLX_SD, BINEST_COL_LOGW_MEAN, BINEST_COL_LOGW_SD, BINEST_NCOL_F
};
int binest_combine_runs(int64_t R, const int64_t *sizes, int64_t d, const double *points, const double *logL,
                        const double *logPrior, const double *acc, const int64_t *pool, const int64_t *run_id,
                        int32_t scheme, int64_t n_tot, int64_t post_runs, uint64_t seed, double *points_out,
                        double *table_out, int64_t *itable_out, double *z, double *pmean, double *H, double *summary,
                        int64_t *M_out, int64_t *n_live_out);

/* The same two operations on a run group's own device state (no per-run fetch, no host round trip of the inputs):
 * the Join is the runs of the group in order, each as binest_run_fetch would return it, run_id = first_run_id + index.
 * binest_run_merge_size gives the allocation bound (total samples before duplicate removal).  binest_run_combine uses
 * n_tot = n_runs * pool_size.  A pre-merged list (its summed PoolSize and RunIndex columns) can itself be an input
 * of binest_merge_runs / binest_combine_runs: merging per-GPU merges equals merging all runs at once. */
int binest_run_merge_size(binest_run *r, int64_t *M_total);
int binest_run_merge(binest_run *r, double *points_out, double *logL_out, double *logPrior_out, double *acc_out,
                     int64_t *pool_out, int64_t *run_out, int64_t *M_out, int64_t *live_block);
int binest_run_combine(binest_run *r, int32_t scheme, int64_t post_runs, uint64_t seed, double *points_out,
                       double *table_out, int64_t *itable_out, double *z, double *pmean, double *H, double *summary,
                       int64_t *M_out, int64_t *n_live_out);

/* Device-resident variants for multi-GPU jobs (one process per GPU): the per-GPU merge stays in device memory, the
 * host language gathers the merges with its collective library (NCCL all-gather over NVLink: the one exchange step of
 * parallelNestedSampling, BS:1349-1363) and hands the gathered lists back without a host round trip.
 * A packed table has one row per sample, d + 5 doubles: point[d], logL, logPrior, acc, pool size, run id.
 * binest_run_merge_dev writes the merge of the group's runs to table_dev (caller-allocated, binest_run_merge_size
 * rows); binest_combine_runs_dev takes R such lists concatenated in rank order (sizes[R] rows each). */
int binest_run_merge_dev(binest_run *r, double *table_dev, int64_t *M_out, int64_t *live_block);
int binest_combine_runs_dev(int64_t R, const int64_t *sizes, int64_t d, const double *table_dev, int32_t scheme,
                            int64_t n_tot, int64_t post_runs, uint64_t seed, double *points_out, double *table_out,
                            int64_t *itable_out, double *z, double *pmean, double *H, double *summary, int64_t *M_out,
                            int64_t *n_live_out);

/* ---- data-sharded mode ("very large N"): rows split across the GPUs of one box ------------
 * One process per GPU.  Rank 0 makes an id, the host passes it to every rank (any channel), every rank creates
 * its communicator (collective), defines its problem from ITS rows only (GBM: shards overlap by one point, the
 * increment between them belongs to the later shard) and declares it a shard (collective).  From then on
 * binest_loglike / binest_run_* are collectives: all ranks must make the same calls with the same theta / options /
 * seed / first_run_id.  Every rank walks the same chains (same Philox counters); after each likelihood launch the
 * per-walker shard sums are exchanged — pushed by the reducing kernel itself into every peer's receive buffer over
 * peer-mapped memory (CUDA IPC, NVLink), 8 P bytes to each peer, the whole S-step walk being one CUDA graph; or
 * ncclAllGather on the fallback — and added in rank order, so every rank takes bit-identical accept/reject decisions
 * and returns the same samples.  The reference has no such mode (its only
 * parallelism is independent runs, BS:1349-1357); logL values equal the unsharded ones up to summation order.
 * Not available for BINEST_OP_GP_SE (replicas only). */
#define BINEST_COMM_ID_BYTES 128
typedef struct binest_comm binest_comm;
int binest_comm_unique_id(uint8_t *id /*[BINEST_COMM_ID_BYTES]*/);
int binest_comm_create(int rank, int world, const uint8_t *id, binest_comm **out);
int binest_comm_info(const binest_comm *c, int *rank, int *world);
int binest_comm_free(binest_comm *c);
int binest_problem_shard(binest_problem *p, binest_comm *c);
/* Batch-sharded mode (BINEST_OP_GP_SE only): the data are replicated, every theta batch
 * (binest_loglike, the proposals of a GP walk) is split into `world` contiguous slices, each rank fills and factors
 * the covariance matrices of its slice (GP:130-141, 181-199), and the finished log-likelihoods are exchanged
 * (8 * ceil(P / world) bytes to each peer).  Collective like the data-sharded mode: same calls on every rank. */
int binest_problem_shard_batch(binest_problem *p, binest_comm *c);
/* Exchange statistics: exchanges issued, payload bytes this rank stored into its peers' buffers, and peer_path = 1
 * when the exchange runs inside our kernels over peer-mapped memory (CUDA IPC over NVLink; the default), 0 on the
 * ncclAllGather fallback (BINEST_XCHG=nccl or no peer access). */
int binest_comm_stats(const binest_comm *c, int64_t *exchanges, int64_t *bytes_pushed, int *peer_path);

/* ---- measurement helper (bench.py): inputs resident in HBM, CUDA events on the launching stream ---- */
/* Scores P prior draws `reps` times after `warmup`; flush_l2 != 0 writes 256 MiB between repetitions.
 * ms_kernel: mean duration of the streaming likelihood kernel; ms_total: including the finalize kernel. */
int binest_bench_loglike(binest_problem *p, int64_t P, int64_t reps, int64_t warmup, int flush_l2,
                         double *ms_kernel, double *ms_total);

#ifdef __cplusplus
}
#endif
#endif /* BINEST_H */
