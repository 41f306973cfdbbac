// engine.cu — host orchestration of nestedSamplingInternal (BS:859-1040) for a group of lock-step runs,
// plus the evidence entry points (BS:812-831, 1158-1291).  One process drives one GPU; runs are sharded
// across GPUs by the caller (first_run_id / n_runs) and merged on the host (combineRuns BS:1293-1315).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <vector>

#include "evidence.cuh"
#include "problem.cuh"
#include "walk.cuh"
#include "walk_grid.cuh"
#include "walk_loop.cuh"
#include "run_view.cuh"
#include "walk_resident.cuh"

namespace binest {
int guard(const std::function<void()> &f);
void upload_theta(binest_problem &p, const double *theta, int64_t P, int Ps);
void gp_loglike_device_strided(binest_problem &p, const double *theta_dev, int P, int Ps, double *out_dev,
                               int out_stride, bool check_box);
}  // namespace binest

using namespace binest;

constexpr int kMaxRetry = 20;  // outer acceptance-retry rounds (1.25^20 = 87 x S steps in the last one)

struct binest_run {
    binest_problem *prob = nullptr;
    binest_options opt{};
    RunParams prm{};
    RunArrays A{};
    int n_pad = 0;
    StreamGeom geom{};
    bool resident = false;      // whole walk in one launch, data in (distributed) shared memory
    int res_cs = 1;             // cluster size of the resident kernel
    int res_tw = 1, res_ch = 16; // walkers per lane, pre-generated steps
    int res_nw = 16;            // warps per CTA (8: two CTAs per SM)
    long long res_rpc = 0;      // data rows per CTA of the cluster
    size_t res_smem = 0;
    bool loop = false;          // whole nested-sampling loop in one launch per advance (walk_loop.cuh)
    int loop_nt = 256;
    bool loop_k1 = false;       // K = 1: CTA-wide two-level speculative walk
    size_t loop_smem = 0;
    LoopCtl *h_ctl = nullptr;   // pinned
    DevBuf<LoopCtl> ctl;
    bool grid = false;          // whole walk in one persistent cooperative launch, data resident in all SMs' smem
    int grid_G = 0, grid_Gs = 0, grid_tw = 1, grid_passes = 0, grid_passesA = 0;
    long long grid_rpc = 0;
    size_t grid_smem = 0;
    unsigned *h_abort = nullptr;  // pinned
    bool first = true;
    bool finished = false;
    int64_t evals = 0;
    int64_t batches = 0;
    int64_t dead_upper = 0;  // host-side upper bound of n_dead per run
    cudaStream_t stream = nullptr;
    cudaGraphExec_t walk_graph = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::vector<cudaEvent_t> chunk_ev;  // event pairs around the walks of a chunk of iterations
    int64_t graph_exchanges = 0, graph_bytes = 0, graph_launches = 0;  // per launch of the sharded walk graph
    double walk_ms = 0.0;   // device time spent in walk graphs (CUDA events on the run's stream)
    int64_t walk_graphs = 0;
    RunState *h_state = nullptr;  // pinned mirror
    int *h_unfrozen = nullptr;    // pinned
    DevBuf<double> live_theta, live_logL, live_logPr, live_acc;
    DevBuf<double> dead_theta, dead_logL, dead_logPr, dead_acc, dead_logX;
    DevBuf<int> dead_pool, order, kill_slot, w_flags, w_nacc, w_steps, n_unfrozen;
    DevBuf<RunState> state;
    DevBuf<GridSync> gsync;
    DevBuf<double> w_theta, w_logL, w_logPr, w_prop, w_prop_logPr, w_mean, w_cov, partials;

    ~binest_run() {
        if (walk_graph) cudaGraphExecDestroy(walk_graph);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        for (cudaEvent_t e : chunk_ev) cudaEventDestroy(e);
        if (h_state) cudaFreeHost(h_state);
        if (h_unfrozen) cudaFreeHost(h_unfrozen);
        if (h_abort) cudaFreeHost(h_abort);
        if (h_ctl) cudaFreeHost(h_ctl);
    }
};

namespace {

void bind_arrays(binest_run &r) {
    RunArrays &A = r.A;
    A.live_theta = r.live_theta.p; A.live_logL = r.live_logL.p; A.live_logPr = r.live_logPr.p; A.live_acc = r.live_acc.p;
    A.dead_theta = r.dead_theta.p; A.dead_logL = r.dead_logL.p; A.dead_logPr = r.dead_logPr.p;
    A.dead_acc = r.dead_acc.p; A.dead_logX = r.dead_logX.p; A.dead_pool = r.dead_pool.p;
    A.order = r.order.p; A.kill_slot = r.kill_slot.p; A.state = r.state.p;
    A.w_theta = r.w_theta.p; A.w_logL = r.w_logL.p; A.w_logPr = r.w_logPr.p; A.w_prop = r.w_prop.p;
    A.w_prop_logPr = r.w_prop_logPr.p; A.w_mean = r.w_mean.p; A.w_cov = r.w_cov.p;
    A.w_flags = r.w_flags.p; A.w_nacc = r.w_nacc.p; A.w_steps = r.w_steps.p; A.n_unfrozen = r.n_unfrozen.p;
}

template <class T>
void grow(DevBuf<T> &buf, int R, int64_t old_cap, int64_t new_cap, int width, cudaStream_t s) {
    DevBuf<T> nb((size_t)R * new_cap * width);
    for (int r = 0; r < R; ++r)
        BN_CUDA(cudaMemcpyAsync(nb.p + (size_t)r * new_cap * width, buf.p + (size_t)r * old_cap * width,
                                sizeof(T) * old_cap * width, cudaMemcpyDeviceToDevice, s));
    BN_CUDA(cudaStreamSynchronize(s));
    buf = std::move(nb);
}

void ensure_dead_capacity(binest_run &r, int64_t need) {
    if (need <= r.prm.cap) return;
    int64_t cap = r.prm.cap;
    while (cap < need) cap *= 2;
    const int R = r.prm.R, d = r.prm.d;
    grow(r.dead_theta, R, r.prm.cap, cap, d, r.stream);
    grow(r.dead_logL, R, r.prm.cap, cap, 1, r.stream);
    grow(r.dead_logPr, R, r.prm.cap, cap, 1, r.stream);
    grow(r.dead_acc, R, r.prm.cap, cap, 1, r.stream);
    grow(r.dead_logX, R, r.prm.cap, cap, 1, r.stream);
    grow(r.dead_pool, R, r.prm.cap, cap, 1, r.stream);
    r.prm.cap = cap;
    bind_arrays(r);
}

// Geometry of the persistent grid-resident walk (walk_grid.cuh): G = SMs x resident CTAs, every CTA keeps
// rows/G data rows in shared memory; walkers are tiled TW per lane and split into two alternating sets.
// Returns false when the data do not fit (the stepped graph is used instead).
// a walker warp keeps at most kGridMaxOwn walkers per set (G >= num_sms CTAs)
inline bool G_own_too_many(int num_sms, int PA, int P) {
    const int a = std::min(PA, P), b = P - a;
    return (a + num_sms - 1) / num_sms > kGridMaxOwn || (b + num_sms - 1) / num_sms > kGridMaxOwn;
}

// Geometry of the cluster-resident walk (walk_resident.cuh): TW walkers per lane (a cluster owns 32 TW walkers), CS CTAs
// per cluster sharing the data rows, CH pre-generated steps of increments.  Every (TW, CS) that fits is priced with a
// small model of one walk step and the cheapest wins:
//   data phase  rows_per_cta x max(4, TW SLOTS / 2) clocks per SM — one broadcast LDS.128 per row (512 B at 128 B/clk)
//               against TW SLOTS warp-wide DFMAs at 64 lanes/clk; TW = 1 is bound by the shared-memory return path;
//   per step    ~5000 clocks of chain logic, barriers and the DSMEM exchange (measured: C4, r2g); CTAs that end up on
//               the same SM multiply the data phase (r2k: TW = 2 / CS = 4 put 256 CTAs on 148 SMs, 3.5 ms per C4 walk
//               against 2.5 ms for TW = 4 / CS = 4 with 128);
//   waves       all clusters must be co-resident or the walk runs in several waves: a cluster of 8 fits only twice into
//               a GPC and one GPC of a B200 is short of SMs — 15 clusters of 8, not 16 (ncu r2g: 16 clusters of 8 ran
//               as two waves, 5.1 ms instead of 2.5 ms per C4 walk) — cudaOccupancyMaxActiveClusters tells.
// BINEST_RES_TW / BINEST_RES_CS force a tile width / cluster size (experiments).
constexpr double kResChainClocks = 2000.0;

template <class OP>
bool plan_resident(binest_run &r, int P) {
    binest_problem &p = *r.prob;
    const size_t budget = 200 * 1024;
    static const int tw_force = [] { const char *e = std::getenv("BINEST_RES_TW"); return e ? std::atoi(e) : 0; }();
    static const int cs_force = [] { const char *e = std::getenv("BINEST_RES_CS"); return e ? std::atoi(e) : 0; }();
    static const int nw_force = [] { const char *e = std::getenv("BINEST_RES_NW"); return e ? std::atoi(e) : 0; }();
    struct Plan { int tw = 0, cs = 0, nw = 0, ctas = 0; double cost = 1e300; long long rpc = 0; size_t smem = 0; } best;
    for (int nw = kResWarpsMax; nw >= 8; nw >>= 1) {
        // 8 data warps per CTA (two CTAs per SM) stay an experiment (BINEST_RES_NW=8): measured slower than 16 wherever
        // the data phase matters (profiles/r02_resident_sweep.md)
        if (nw_force > 0 ? nw != nw_force : nw != kResWarpsMax) continue;
        for (int tw = OP::TW_MAX; tw >= 1; tw >>= 1) {
            if (tw_force > 0 && tw != tw_force && tw_force <= OP::TW_MAX) continue;
            if ((nw + tw) * 32 > 640) continue;  // 65536 registers / ~100 per thread
            const int groups = (P + 32 * tw - 1) / (32 * tw);
            if (tw > 1 && groups * 32 * tw >= 2 * P) continue;  // more than half of the tile would be padding
            for (int cs = 1; cs <= 16; cs <<= 1) {
                if (cs_force > 0 && cs != cs_force) continue;
                if (cs > 1 && p.rows / cs < 4 * nw) break;  // shards thinner than a few rows per warp
                Plan pl;
                pl.tw = tw; pl.cs = cs; pl.nw = nw; pl.ctas = groups * cs;
                pl.rpc = ((p.rows + cs - 1) / cs + 1) & ~1LL;
                pl.smem = resident_smem_doubles<OP>(pl.rpc, cs, tw, nw) * sizeof(double);
                if (pl.smem > budget) continue;
                int max_clusters = 0, per_sm_fit = 1;
                dispatch_tw<OP>(tw, [&](auto twc) {
                    constexpr int TW = decltype(twc)::value;
                    auto query = [&](auto kern) {
                        BN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
                        if (cs > 8 && cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
                            cudaGetLastError();
                            return;
                        }
                        cudaLaunchConfig_t cfg{};
                        cudaLaunchAttribute attr[1];
                        cfg.gridDim = dim3(pl.ctas);
                        cfg.blockDim = dim3((nw + tw) * 32);
                        cfg.dynamicSmemBytes = pl.smem;
                        attr[0].id = cudaLaunchAttributeClusterDimension;
                        attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
                        cfg.attrs = attr;
                        cfg.numAttrs = 1;
                        if (cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg) != cudaSuccess) {
                            cudaGetLastError();
                            max_clusters = cs > 8 ? 0 : p.num_sms / cs;
                        }
                        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_fit, kern, (nw + tw) * 32, pl.smem) != cudaSuccess) {
                            cudaGetLastError();
                            per_sm_fit = 1;
                        }
                    };
                    if (nw == 8) query(walk_resident_kernel<OP, TW, 8>);
                    else query(walk_resident_kernel<OP, TW, 16>);
                });
                if (max_clusters < 1) continue;
                const int waves = (groups + max_clusters - 1) / max_clusters;
                // Cost per walk step, fitted to a sweep of (NW, TW, CS) on C4 with 64 and 8 runs per GPU
                // (profiles/r02_resident_sweep.md): a row costs 1.15 clocks of shared-memory traffic plus 0.435 clocks per
                // DFMA slot of the TW walkers of a lane; what stays between two data phases (barriers, combine, DSMEM
                // exchange, accept rule) ~2000 clocks, ~1500 more across a 16-CTA cluster.  CTAs sharing an SM share its
                // fp64 pipe; they did not overlap one's chain phase with the other's data phase as hoped.
                const int per_sm = waves == 1 ? std::min(std::max(per_sm_fit, 1), (pl.ctas + p.num_sms - 1) / p.num_sms) : std::max(per_sm_fit, 1);
                const double data_clk = (double)pl.rpc * (1.15 + 0.435 * tw * OP::SLOTS);
                const double chain_clk = kResChainClocks + (cs > 8 ? 1500.0 : 0.0);
                const double step_clk = (per_sm * data_clk + chain_clk) * (per_sm > 1 ? 1.4 : 1.0);
                pl.cost = waves * step_clk * (1.0 + 0.01 / tw);  // ties: the wider tile
                if (std::getenv("BINEST_PLAN_DEBUG"))
                    std::fprintf(stderr, "resident plan: nw %d tw %d cs %d groups %d ctas %d max_clusters %d per_sm %d waves %d smem %zu cost %.0f\n",
                                 nw, tw, cs, groups, pl.ctas, max_clusters, per_sm, waves, pl.smem, pl.cost);
                if (pl.cost < best.cost * 0.999 || (pl.cost <= best.cost * 1.001 && pl.ctas > best.ctas)) best = pl;
            }
        }
    }
    if (best.tw == 0) return false;
    r.resident = true;
    r.res_tw = best.tw; r.res_cs = best.cs; r.res_rpc = best.rpc; r.res_smem = best.smem; r.res_nw = best.nw;
    dispatch_tw<OP>(best.tw, [&](auto twc) {
        constexpr int TW = decltype(twc)::value;
        auto prep = [&](auto kern) {
            BN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)r.res_smem));
            if (best.cs > 8) BN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        };
        if (best.nw == 8) prep(walk_resident_kernel<OP, TW, 8>);
        else prep(walk_resident_kernel<OP, TW, 16>);
    });
    if (std::getenv("BINEST_PLAN_DEBUG"))
        std::fprintf(stderr, "resident plan chosen: nw %d tw %d cs %d ctas %d smem %zu\n", best.nw, best.tw, best.cs, best.ctas, best.smem);
    return true;
}

// Device-resident loop (walk_loop.cuh): data rows and the sort buffers fit one CTA's shared memory, at most 32 walkers
// per iteration (a warp each), default acceptance range (the acceptance-retry protocol BS:730-736, 995-1004 needs the
// host between blocks of steps).
template <class OP>
bool plan_loop(binest_run &r) {
    binest_problem &p = *r.prob;
    const RunParams &q = r.prm;
    if (q.acc_min > 0.0 || q.acc_max < 1.0) return false;
    if (q.K > 32 || p.rows * OP::NCOL > 4096 || r.n_pad > 4096) return false;
    if (q.K > 8 && OP::D > 3) return false;  // the 1024-thread instantiation has 64 registers per thread
    r.loop_nt = q.K <= 8 ? 256 : 1024;
    r.loop_smem = loop_smem_bytes<OP>(p.rows, r.n_pad);
    // K = 1 (the reference scheme): the whole CTA walks the run's single walker, two speculation levels deep
    // (cta_walk_k1); its draw / divisor tables follow the sort buffers in shared memory
    r.loop_k1 = q.K == 1 && r.loop_nt == 256 && std::getenv("BINEST_NO_K1") == nullptr &&
                k1_table_doubles<OP>(q.S) * sizeof(double) <= (size_t)kK1MaxTableBytes;
    if (r.loop_k1) r.loop_smem = ((r.loop_smem + 15) & ~(size_t)15) + k1_table_doubles<OP>(q.S) * sizeof(double);
    if (r.loop_nt == 256)
        BN_CUDA(cudaFuncSetAttribute(ns_loop_kernel<OP, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)r.loop_smem));
    else
        BN_CUDA(cudaFuncSetAttribute(ns_loop_kernel<OP, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)r.loop_smem));
    BN_CUDA(cudaHostAlloc((void **)&r.h_ctl, sizeof(LoopCtl), cudaHostAllocDefault));
    r.ctl.alloc(1);
    r.loop = true;
    return true;
}

// One launch = iterations until every run of the group has terminated, `budget` iterations are done, or a dead list is
// full.  Returns the number of iterations executed (max over the runs).
template <class OP>
long long launch_loop(binest_run &r, long long budget) {
    binest_problem &p = *r.prob;
    BN_CUDA(cudaMemsetAsync(r.ctl.p, 0, sizeof(LoopCtl), r.stream));
    BN_CUDA(cudaMemsetAsync(r.n_unfrozen.p, 0, sizeof(int), r.stream));
    const double *data = p.data.p;
    const int mode = r.first ? 1 : 0;
    BN_CUDA(cudaEventRecord(r.ev0, r.stream));
    if (r.loop_nt == 256)
        ns_loop_kernel<OP, 256><<<r.prm.R, 256, r.loop_smem, r.stream>>>(r.prm, r.A, p.prior, data, p.rows, p.cst, r.n_pad, mode, budget, r.ctl.p, r.loop_k1 ? 1 : 0);
    else
        ns_loop_kernel<OP, 1024><<<r.prm.R, 1024, r.loop_smem, r.stream>>>(r.prm, r.A, p.prior, data, p.rows, p.cst, r.n_pad, mode, budget, r.ctl.p, 0);
    BN_LAUNCH_CHECK();
    BN_CUDA(cudaEventRecord(r.ev1, r.stream));
    r.first = false;
    BN_CUDA(cudaMemcpyAsync(r.h_ctl, r.ctl.p, sizeof(LoopCtl), cudaMemcpyDeviceToHost, r.stream));
    BN_CUDA(cudaMemcpyAsync(r.h_state, r.state.p, sizeof(RunState) * r.prm.R, cudaMemcpyDeviceToHost, r.stream));
    BN_CUDA(cudaStreamSynchronize(r.stream));
    float ms = 0;
    BN_CUDA(cudaEventElapsedTime(&ms, r.ev0, r.ev1));
    r.walk_ms += ms;
    r.walk_graphs += 1;
    return r.h_ctl->iters;
}

template <class OP>
bool plan_grid_walk(binest_run &r, int P) {
    binest_problem &p = *r.prob;
    int coop = 0;
    BN_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, p.device));
    if (!coop) return false;
    const char *e = std::getenv("BINEST_GRID_SETS");
    const int nsets = (e && std::atoi(e) == 1) ? 1 : 2;
    const int lanesets = (P + 31) / 32;
    int tw = 1;
    while (tw * 2 <= OP::TW_MAX && tw * 2 * nsets <= lanesets) tw <<= 1;
    const int passes = (P + 32 * tw - 1) / (32 * tw);
    const int passesA = nsets == 2 ? (passes + 1) / 2 : passes;
    bool ok = false;
    if (G_own_too_many(p.num_sms, 32 * tw * passesA, P)) return false;
    dispatch_tw<OP>(tw, [&](auto twc) {
        constexpr int TW = decltype(twc)::value;
        const int G = p.num_sms;  // one CTA per SM
        long long rpc = ((p.rows + G - 1) / G + 1) & ~1LL;
        rpc = std::max<long long>(rpc, 2);
        const size_t smem = grid_smem_bytes<OP, TW>(rpc);
        if (smem > 212u * 1024u) return;
        BN_CUDA(cudaFuncSetAttribute(walk_grid_kernel<OP, TW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int nb = 0;
        BN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, walk_grid_kernel<OP, TW>, kGridThreads, smem));
        if (nb < 1) return;
        r.grid_G = G;
        r.grid_Gs = (G + 3) & ~3;
        r.grid_rpc = rpc;
        r.grid_smem = smem;
        ok = true;
    });
    if (!ok) return false;
    r.grid = true;
    r.grid_tw = tw;
    r.grid_passes = passes;
    r.grid_passesA = passesA;
    r.partials.alloc((size_t)r.grid_Gs * r.prm.Ps);
    r.gsync.alloc(1);
    BN_CUDA(cudaHostAlloc((void **)&r.h_abort, sizeof(unsigned), cudaHostAllocDefault));
    *r.h_abort = 0;
    return true;
}

template <class OP>
void launch_grid_walk(binest_run &r, const RunParams &q) {
    binest_problem &p = *r.prob;
    long long *dbg = nullptr;
    if (std::getenv("BINEST_GRID_TRACE")) {
        BN_CUDA(cudaMalloc((void **)&dbg, sizeof(long long) * 2 * 2 * 16 * 2 * 4));
        BN_CUDA(cudaMemset(dbg, 0, sizeof(long long) * 2 * 2 * 16 * 2 * 4));
    }
    BN_CUDA(cudaMemsetAsync(r.gsync.p, 0, sizeof(GridSync), r.stream));
    dispatch_tw<OP>(r.grid_tw, [&](auto twc) {
        constexpr int TW = decltype(twc)::value;
        cudaLaunchConfig_t cfg{};
        cudaLaunchAttribute attr[1];
        cfg.gridDim = dim3(r.grid_G);
        cfg.blockDim = dim3(kGridThreads);
        cfg.dynamicSmemBytes = r.grid_smem;
        cfg.stream = r.stream;
        attr[0].id = cudaLaunchAttributeCooperative;  // all CTAs co-resident, or the launch fails
        attr[0].val.cooperative = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        const double *data = p.data.p;
        long long rows = p.rows, rpc = r.grid_rpc;
        OpCst cst = p.cst;
        double *partials = r.partials.p;
        int Gs = r.grid_Gs, passes = r.grid_passes, passesA = r.grid_passesA;
        GridSync *gs = r.gsync.p;
        BN_CUDA(cudaLaunchKernelEx(&cfg, walk_grid_kernel<OP, TW>, q, r.A, p.prior, data, rows, rpc, cst, partials, Gs,
                                   passes, passesA, gs, dbg));
        BN_LAUNCH_CHECK();
    });
    if (dbg) {  // BINEST_GRID_TRACE: print the timeline of CTAs 0 and G/2 (cycles relative to the first event)
        std::vector<long long> h(2 * 2 * 16 * 2 * 4);
        BN_CUDA(cudaMemcpyAsync(h.data(), dbg, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost, r.stream));
        BN_CUDA(cudaStreamSynchronize(r.stream));
        for (int c = 0; c < 2; ++c) {
            long long t0 = 0;
            for (long long v : std::vector<long long>(h.begin() + c * 256, h.begin() + (c + 1) * 256))
                if (v && (!t0 || v < t0)) t0 = v;
            for (int role = 0; role < 2; ++role)
                for (int s = 0; s < 16; ++s)
                    for (int X = 0; X < 2; ++X) {
                        const long long *e = &h[((((size_t)c * 2 + role) * 16 + s) * 2 + X) * 4];
                        std::fprintf(stderr, "trace cta%d %s s=%2d X=%d  %9.2f %9.2f %9.2f %9.2f us\n", c,
                                     role ? "data  " : "walker", s, X, (e[0] - t0) / 1965.0, (e[1] - t0) / 1965.0,
                                     (e[2] - t0) / 1965.0, (e[3] - t0) / 1965.0);
                    }
        }
        cudaFree(dbg);
    }
    BN_CUDA(cudaMemcpyAsync(r.h_abort, &r.gsync.p->abort, sizeof(unsigned), cudaMemcpyDeviceToHost, r.stream));
}

// [walk_step, loglike_stream (, shard exchange)] x S + the final accept, enqueued on the run's stream (directly, or
// under stream capture).  Data-sharded: every rank walks the same chains (same Philox counters) on its own rows; after
// each likelihood launch the per-rank sums are exchanged (problem.cuh: shard_exchange — in-kernel pushes over
// peer-mapped memory, or ncclAllGather on the fallback) and combined in rank order by the next walk_step.
void stepped_walk(binest_run &r, const RunParams &q) {
    binest_problem &p = *r.prob;
    const int P = q.R * q.K;
    const dim3 sgrid((P * 32 + 255) / 256), sblock(256);
    dispatch_op(p, [&](auto op) {
        using OP = decltype(op);
        PartialView pv = p.comm ? PartialView{p.sh_recv.p, p.comm->world, 1, q.Ps, 1}
                                : PartialView{r.partials.p, r.geom.G, r.geom.Gs, 1};
        const XchgDev *xd = (p.comm && p.comm->peer) ? p.comm->xd_dev : nullptr;
        for (int step = 0; step <= q.S; ++step) {
            walk_step_kernel<OP><<<sgrid, sblock, 0, r.stream>>>(q, r.A, p.prior, pv, p.rows_eff(), p.cst_eff(),
                                                                step == q.S ? 1 : 0, xd, step == 0 ? 1 : 0);
            BN_LAUNCH_CHECK();
            if (step < q.S) {
                launch_loglike<OP>(p, r.w_prop.p, P, q.Ps, r.partials.p, r.geom, r.stream);
                if (p.comm) pv = shard_exchange<OP>(p, r.partials.p, r.w_prop.p, P, q.Ps, r.geom, r.stream);
            }
        }
    });
}

// the S-step walk as one CUDA graph: [walk_step, loglike_stream] x S, then the final accept
void build_walk_graph(binest_run &r) {
    binest_problem &p = *r.prob;
    const int P = r.prm.R * r.prm.K;
    const int S = (int)r.prm.S;
    const cudaStream_t s = r.stream;
    if (p.op == BINEST_OP_GP_SE) {  // a GP step is hundreds of launches: stepped directly, see walk_block()
        r.geom = StreamGeom{1, 1, 1, 4, 0};
        r.partials.alloc((size_t)r.geom.Gs * r.prm.Ps);
        return;
    }
    dispatch_op(p, [&](auto op) {
        using OP = decltype(op);
        r.geom = stream_geom<OP>(p, P);
        r.partials.alloc((size_t)r.geom.Gs * r.prm.Ps);
        // data-sharded on the peer path: the whole S-step walk incl. the in-kernel exchanges is ONE graph
        // (3 S + 1 nodes); on the NCCL fallback it is stepped directly, see walk_block()
        if (p.comm) {
            if (!p.comm->peer || std::getenv("BINEST_NO_SHARD_GRAPH")) return;
            BN_REQUIRE(r.prm.Ps <= kXchgSlotDoubles, BINEST_ERR_DIMENSION, "too many walkers for the sharded exchange buffer");
            const int64_t ex0 = p.comm->exchanges, by0 = p.comm->bytes_pushed, l0 = g_launches.load();
            BN_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
            stepped_walk(r, r.prm);
            cudaGraph_t g;
            BN_CUDA(cudaStreamEndCapture(s, &g));
            BN_CUDA(cudaGraphInstantiate(&r.walk_graph, g, 0));
            cudaGraphDestroy(g);
            r.graph_exchanges = p.comm->exchanges - ex0;
            r.graph_bytes = p.comm->bytes_pushed - by0;
            r.graph_launches = g_launches.load() - l0;
            p.comm->exchanges = ex0; p.comm->bytes_pushed = by0; g_launches.store(l0);  // capture enqueued nothing
            return;
        }
        // tiny, latency-bound problems: the whole nested-sampling loop stays on the device (walk_loop.cuh)
        if (std::getenv("BINEST_NO_LOOP") == nullptr && plan_loop<OP>(r)) return;
        // small data: the resident cluster kernel replaces the per-step graph (walk_resident.cuh)
        if (std::getenv("BINEST_NO_RESIDENT") == nullptr && plan_resident<OP>(r, P)) return;
        // data that fits the shared memories of all SMs: one persistent launch per walk (walk_grid.cuh)
        if (std::getenv("BINEST_NO_GRID") == nullptr && plan_grid_walk<OP>(r, P)) return;
        const dim3 sgrid((P * 32 + 255) / 256), sblock(256);  // one warp per walker
        BN_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        // every node after the first is a programmatic dependent of its predecessor (PDL): the likelihood
        // kernel prefetches its first data tiles while walk_step runs, walk_step is resident when the
        // likelihood kernel drains
        const bool pdl = std::getenv("BINEST_NO_PDL") == nullptr;
        for (int step = 0; step <= S; ++step) {
            PdlConfig lc(sgrid, sblock, s, pdl && step > 0);
            const PartialView pv{r.partials.p, r.geom.G, r.geom.Gs, 1};
            int fin = step == S ? 1 : 0;
            double rows = (double)p.rows;
            OpCst cst = p.cst;
            const XchgDev *xd = nullptr;
            int first = step == 0 ? 1 : 0;
            BN_CUDA(cudaLaunchKernelEx(&lc.cfg, walk_step_kernel<OP>, r.prm, r.A, p.prior, pv, rows, cst, fin, xd, first));
            if (step < S) launch_loglike<OP>(p, r.w_prop.p, P, r.prm.Ps, r.partials.p, r.geom, s, false, pdl);
        }
        cudaGraph_t g;
        BN_CUDA(cudaStreamEndCapture(s, &g));
        BN_CUDA(cudaGraphInstantiate(&r.walk_graph, g, 0));
        cudaGraphDestroy(g);
    });
}

// one block of q.S walk steps for every active walker.  q = r.prm, or its copy for an outer acceptance-retry round
// (attempt > 0: more steps, shifted Philox counters); the pre-built CUDA graph only serves q.S == r.prm.S.
void walk_block(binest_run &r, const RunParams &q) {
    binest_problem &p = *r.prob;
    if (r.resident) {
        dispatch_op(p, [&](auto op) {
            using OP = decltype(op);
            const int groups = (q.R * q.K + 32 * r.res_tw - 1) / (32 * r.res_tw);
            cudaLaunchConfig_t cfg{};
            cudaLaunchAttribute attr[1];
            cfg.gridDim = dim3(groups * r.res_cs);
            cfg.blockDim = dim3((r.res_nw + r.res_tw) * 32);
            cfg.dynamicSmemBytes = r.res_smem;
            cfg.stream = r.stream;
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = r.res_cs;
            attr[0].val.clusterDim.y = 1;
            attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = r.res_cs > 1 ? 1 : 0;
            const double *data = p.data.p;
            long long rows = p.rows, rpc = r.res_rpc;
            OpCst cst = p.cst;
            int cs = r.res_cs;
            dispatch_tw<OP>(r.res_tw, [&](auto twc) {
                constexpr int TW = decltype(twc)::value;
                if (r.res_nw == 8)
                    BN_CUDA(cudaLaunchKernelEx(&cfg, walk_resident_kernel<OP, TW, 8>, q, r.A, p.prior, data, rows, rpc, cst, cs));
                else
                    BN_CUDA(cudaLaunchKernelEx(&cfg, walk_resident_kernel<OP, TW, 16>, q, r.A, p.prior, data, rows, rpc, cst, cs));
            });
            BN_LAUNCH_CHECK();
        });
        return;
    }
    if (r.grid) {
        dispatch_op(p, [&](auto op) { launch_grid_walk<decltype(op)>(r, q); });
        return;
    }
    const int P = q.R * q.K;
    const dim3 sgrid((P * 32 + 255) / 256), sblock(256);
    if (p.op == BINEST_OP_GP_SE) {
        for (int step = 0; step <= q.S; ++step) {
            walk_step_kernel<OpGpSe><<<sgrid, sblock, 0, r.stream>>>(q, r.A, p.prior, PartialView{r.partials.p, r.geom.G, r.geom.Gs, 1},
                                                                     (double)p.rows, p.cst, step == q.S ? 1 : 0);
            BN_LAUNCH_CHECK();
            if (step < q.S) gp_loglike_device_strided(p, r.w_prop.p, P, q.Ps, r.partials.p, r.geom.Gs, false);
        }
        return;
    }
    if (q.attempt == 0 && q.S == r.prm.S && r.walk_graph) {
        BN_CUDA(cudaGraphLaunch(r.walk_graph, r.stream));
        if (p.comm) {
            p.comm->exchanges += r.graph_exchanges;
            p.comm->bytes_pushed += r.graph_bytes;
            count_launch((int)r.graph_launches);
        } else {
            count_launch(2 * (int)q.S + 1);
        }
        return;
    }
    // stepped directly: the retry rounds of the graph path, and the data-sharded mode on the NCCL fallback
    stepped_walk(r, q);
}

void launch_update(binest_run &r, bool insert_only = false) {
    BN_CUDA(cudaMemsetAsync(r.n_unfrozen.p, 0, sizeof(int), r.stream));
    const size_t smem = (size_t)r.n_pad * (sizeof(double) + sizeof(int));
    const int mode = insert_only ? 2 : (r.first ? 1 : 0);
    run_update_kernel<<<r.prm.R, 1024, smem, r.stream>>>(r.prm, r.A, r.n_pad, mode);
    BN_LAUNCH_CHECK();
    if (!insert_only) r.first = false;
}

bool all_done(const binest_run &r) {
    for (int i = 0; i < r.prm.R; ++i)
        if (!r.h_state[i].done) return false;
    return true;
}

void fetch_state(binest_run &r) {
    BN_CUDA(cudaMemcpyAsync(r.h_state, r.state.p, sizeof(RunState) * r.prm.R, cudaMemcpyDeviceToHost, r.stream));
    BN_CUDA(cudaMemcpyAsync(r.h_unfrozen, r.n_unfrozen.p, sizeof(int), cudaMemcpyDeviceToHost, r.stream));
    BN_CUDA(cudaStreamSynchronize(r.stream));
}

}  // namespace

extern "C" {

int binest_run_create(binest_problem *p, const binest_options *o, const double *start_points, binest_run **out) {
    return guard([&] {
        BN_REQUIRE(p && o && out, BINEST_ERR_TYPE, "null argument");
        // the live set is sorted in one CTA's shared memory (12 bytes per padded slot): 16384 slots = 192 KB
        BN_REQUIRE(o->pool_size >= 2 && o->pool_size <= 16384, BINEST_ERR_DIMENSION, "2 <= SamplePoolSize <= 16384");
        BN_REQUIRE(o->batch_k >= 1 && o->batch_k < o->pool_size, BINEST_ERR_DIMENSION, "1 <= batch_k < SamplePoolSize");
        BN_REQUIRE(o->mc_steps >= 1 && o->mc_steps <= 100000, BINEST_ERR_DIMENSION, "MonteCarloSteps out of range");
        BN_REQUIRE(o->n_runs >= 1 && o->n_runs <= 4096, BINEST_ERR_DIMENSION, "n_runs out of range");
        BN_REQUIRE(o->term_frac > 0.0, BINEST_ERR_NUMERICAL, "TerminationFraction must be positive");
        BN_CUDA(cudaSetDevice(p->device));
        std::unique_ptr<binest_run> r(new binest_run());
        r->prob = p;
        r->opt = *o;
        r->stream = p->stream;
        RunParams &q = r->prm;
        q.d = p->d; q.n = (int)o->pool_size; q.K = (int)o->batch_k; q.R = (int)o->n_runs;
        q.Ps = (q.R * q.K + 31) & ~31;
        q.max_iter = std::max(o->max_iter, o->min_iter);  // BS:867-868
        q.min_iter = std::min(o->max_iter, o->min_iter);
        q.log_term_frac = std::log(o->term_frac);
        q.acc_min = o->acc_min; q.acc_max = o->acc_max;
        q.S = o->mc_steps; q.maxS = 5 * o->mc_steps;  // BS:872
        q.seed = o->seed; q.first_run_id = (unsigned)o->first_run_id;
        q.logzero = g_logzero;
        q.loglmax_opt = o->loglmax;
        q.cap = std::max<int64_t>(4096, 16 * (int64_t)q.n);
        r->n_pad = 1;
        while (r->n_pad < q.n) r->n_pad <<= 1;
        const size_t Rn = (size_t)q.R * q.n, Ps = q.Ps, d = q.d;
        r->live_theta.alloc(Rn * d); r->live_logL.alloc(Rn); r->live_logPr.alloc(Rn); r->live_acc.alloc(Rn);
        r->dead_theta.alloc((size_t)q.R * q.cap * d); r->dead_logL.alloc((size_t)q.R * q.cap);
        r->dead_logPr.alloc((size_t)q.R * q.cap); r->dead_acc.alloc((size_t)q.R * q.cap);
        r->dead_logX.alloc((size_t)q.R * q.cap); r->dead_pool.alloc((size_t)q.R * q.cap);
        r->order.alloc(Rn); r->kill_slot.alloc((size_t)q.R * q.K);
        r->state.alloc(q.R); r->n_unfrozen.alloc(1);
        r->w_theta.alloc(d * Ps); r->w_logL.alloc(Ps); r->w_logPr.alloc(Ps); r->w_prop.alloc(d * Ps);
        r->w_prop_logPr.alloc(Ps); r->w_mean.alloc(d * Ps); r->w_cov.alloc(d * d * Ps);
        r->w_flags.alloc(Ps); r->w_nacc.alloc(Ps); r->w_steps.alloc(Ps);
        r->w_prop.zero(r->stream); r->w_theta.zero(r->stream); r->w_flags.zero(r->stream);
        BN_CUDA(cudaMallocHost(&r->h_state, sizeof(RunState) * q.R));
        BN_CUDA(cudaMallocHost(&r->h_unfrozen, sizeof(int)));
        BN_CUDA(cudaEventCreate(&r->ev0));
        BN_CUDA(cudaEventCreate(&r->ev1));
        bind_arrays(*r);

        // starting points: supplied, or i.i.d. prior draws per run (BS:1099-1114, 1320-1332)
        if (start_points) {
            BN_CUDA(cudaMemcpyAsync(r->live_theta.p, start_points, sizeof(double) * Rn * d, cudaMemcpyHostToDevice,
                                    r->stream));
        } else {
            for (int i = 0; i < q.R; ++i) {
                sample_prior_kernel<<<(q.n + 127) / 128, 128, 0, r->stream>>>(p->prior, q.n, q.seed, q.first_run_id + i,
                                                                             r->live_theta.p + (size_t)i * q.n * d);
                BN_LAUNCH_CHECK();
            }
        }
        // initial logL / log prior of all start points (BS:902-916): transpose to SoA, evaluate as one batch
        {
            std::vector<double> h(Rn * d);
            BN_CUDA(cudaMemcpyAsync(h.data(), r->live_theta.p, sizeof(double) * Rn * d, cudaMemcpyDeviceToHost, r->stream));
            BN_CUDA(cudaStreamSynchronize(r->stream));
            const int P = (int)Rn, Pst = (P + 31) & ~31;
            upload_theta(*p, h.data(), P, Pst);
            loglike_device(*p, p->s_theta.p, P, Pst, r->live_logL.p);
            logprior_kernel<<<(P + 127) / 128, 128, 0, r->stream>>>(p->s_theta.p, P, Pst, p->prior, g_logzero,
                                                                   r->live_logPr.p);
            BN_LAUNCH_CHECK();
            std::vector<double> hl(Rn), nan(Rn, std::nan(""));
            BN_CUDA(cudaMemcpyAsync(hl.data(), r->live_logL.p, sizeof(double) * Rn, cudaMemcpyDeviceToHost, r->stream));
            BN_CUDA(cudaMemcpyAsync(r->live_acc.p, nan.data(), sizeof(double) * Rn, cudaMemcpyHostToDevice, r->stream));
            BN_CUDA(cudaStreamSynchronize(r->stream));
            for (double v : hl)  // BS:917-921
                BN_REQUIRE(std::isfinite(v), BINEST_ERR_BAD_LIKELIHOOD, "Bad likelihood function");
            r->evals += (int64_t)Rn;
        }
        std::vector<RunState> init(q.R);
        for (auto &s : init) {
            std::memset(&s, 0, sizeof(RunState));
            s.dead.m = -INFINITY;
            s.iteration = 1;
            s.logZ = g_logzero;
        }
        BN_CUDA(cudaMemcpyAsync(r->state.p, init.data(), sizeof(RunState) * q.R, cudaMemcpyHostToDevice, r->stream));
        BN_CUDA(cudaStreamSynchronize(r->stream));
        {   // the opt-in limit is per function, not per run: never lower it under a run that is still alive
            static std::atomic<int> lim{0};
            int want = std::max(64 * 1024, r->n_pad * 12), cur = lim.load();
            while (want > cur && !lim.compare_exchange_weak(cur, want)) {}
            BN_CUDA(cudaFuncSetAttribute(run_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, lim.load()));
        }
        build_walk_graph(*r);
        *out = r.release();
    });
}

int binest_run_advance(binest_run *r, int64_t max_batches, int32_t *finished) {
    return guard([&] {
        BN_REQUIRE(r, BINEST_ERR_TYPE, "null run");
        BN_CUDA(cudaSetDevice(r->prob->device));
        const RunParams &q = r->prm;
        const bool acc_loop = q.acc_min > 0.0 || q.acc_max < 1.0;
        int64_t done_batches = 0;
        // device-resident loop: no host round trip per iteration (walk_loop.cuh)
        while (r->loop && !r->finished && (max_batches <= 0 || done_batches < max_batches)) {
            ensure_dead_capacity(*r, r->dead_upper + q.K + 2);
            const long long budget = max_batches > 0 ? (long long)(max_batches - done_batches) : (1LL << 60);
            long long it0 = 0;
            for (int i = 0; i < q.R; ++i) it0 += r->first ? 1 : r->h_state[i].iteration;
            long long iters = 0;
            dispatch_op(*r->prob, [&](auto op) { iters = launch_loop<decltype(op)>(*r, budget); });
            long long it1 = 0, dmax = 0;
            for (int i = 0; i < q.R; ++i) { it1 += r->h_state[i].iteration; dmax = std::max<long long>(dmax, r->h_state[i].n_dead); }
            r->evals += (int64_t)q.S * (it1 - it0);
            r->dead_upper = dmax;
            done_batches += iters;
            r->batches += iters;
            if (all_done(*r)) { r->finished = true; break; }
            if (r->h_ctl->need_grow) ensure_dead_capacity(*r, 2 * q.cap);
        }
        // resident / grid walk paths without the acceptance protocol: iterations are enqueued kChunk at a time and the
        // run state is read once per chunk (the update and walk kernels of a terminated run return at once), instead
        // of three host round trips per iteration
        constexpr int kChunk = 8;
        while (!r->loop && !acc_loop && (r->resident || r->grid) && !r->finished &&
               (max_batches <= 0 ? true : max_batches - done_batches >= 2)) {
            const int chunk = (int)std::min<int64_t>(kChunk, max_batches <= 0 ? kChunk : max_batches - done_batches);
            ensure_dead_capacity(*r, r->dead_upper + (int64_t)chunk * q.K + 1);
            long long it0 = 0;
            for (int i = 0; i < q.R; ++i) it0 += r->first ? 1 : r->h_state[i].iteration;
            while ((int)r->chunk_ev.size() < 2 * kChunk) {
                cudaEvent_t e;
                BN_CUDA(cudaEventCreate(&e));
                r->chunk_ev.push_back(e);
            }
            for (int c = 0; c < chunk; ++c) {
                launch_update(*r);
                BN_CUDA(cudaEventRecord(r->chunk_ev[2 * c], r->stream));
                walk_block(*r, q);
                BN_CUDA(cudaEventRecord(r->chunk_ev[2 * c + 1], r->stream));
            }
            fetch_state(*r);
            BN_REQUIRE(!(r->grid && *r->h_abort), BINEST_ERR_CUDA, "walk_grid_kernel: grid barrier timed out (walk aborted)");
            long long it1 = 0, dmax = 0;
            for (int i = 0; i < q.R; ++i) { it1 += r->h_state[i].iteration; dmax = std::max<long long>(dmax, r->h_state[i].n_dead); }
            const long long reps = it1 - it0;                      // replacements of the chunk, all runs
            const long long per_iter = (long long)q.R * q.K;
            // walks after the last run terminated were no-ops: keep them out of the per-walk timing statistics
            const int walked = (int)std::min<long long>(chunk, (reps + per_iter - 1) / std::max<long long>(per_iter, 1));
            for (int c = 0; c < chunk; ++c) {
                float ms = 0;
                BN_CUDA(cudaEventElapsedTime(&ms, r->chunk_ev[2 * c], r->chunk_ev[2 * c + 1]));
                if (c < std::max(walked, 1)) { r->walk_ms += ms; r->walk_graphs += 1; }
            }
            r->evals += (int64_t)q.S * reps;
            r->dead_upper = dmax;
            done_batches += chunk;
            r->batches += chunk;
            if (all_done(*r)) { r->finished = true; break; }
        }
        while (!r->loop && !r->finished && (max_batches <= 0 || done_batches < max_batches)) {
            ensure_dead_capacity(*r, r->dead_upper + q.K + 1);
            launch_update(*r);
            fetch_state(*r);
            if (all_done(*r)) { r->finished = true; break; }
            int64_t active = 0;
            for (int i = 0; i < q.R; ++i) active += r->h_state[i].done ? 0 : r->h_state[i].Kb;
            r->dead_upper += q.K;
            // S steps; then extra S-step blocks while some walker's acceptance is out of range (BS:730-736)
            auto walk_until_frozen = [&](const RunParams &qq, int64_t unfrozen0) {
                int blocks = 0;
                do {
                    BN_CUDA(cudaEventRecord(r->ev0, r->stream));
                    walk_block(*r, qq);
                    BN_CUDA(cudaEventRecord(r->ev1, r->stream));
                    BN_CUDA(cudaEventSynchronize(r->ev1));
                    float ms = 0;
                    BN_CUDA(cudaEventElapsedTime(&ms, r->ev0, r->ev1));
                    r->walk_ms += ms;
                    r->walk_graphs += 1;
                    BN_REQUIRE(!(r->grid && *r->h_abort), BINEST_ERR_CUDA,
                               "walk_grid_kernel: grid barrier timed out (walk aborted)");
                    if (r->prob->comm) comm_check_abort(*r->prob->comm, r->stream);
                    if (r->prob->comm_batch) comm_check_abort(*r->prob->comm_batch, r->stream);
                    r->evals += (int64_t)qq.S * (int64_t)(blocks == 0 ? unfrozen0 : *r->h_unfrozen);
                    ++blocks;
                    if (!acc_loop) break;
                    BN_CUDA(cudaMemcpyAsync(r->h_unfrozen, r->n_unfrozen.p, sizeof(int), cudaMemcpyDeviceToHost, r->stream));
                    BN_CUDA(cudaStreamSynchronize(r->stream));
                } while (*r->h_unfrozen > 0 && blocks < 5);
            };
            walk_until_frozen(q, active);
            // outer retry (BS:995-1004): walkers still outside the range start again from a fresh live point with
            // Ceiling[1.25^k S] steps; the reference loops without bound, here at most kMaxRetry rounds
            if (acc_loop) {
                double factor = 1.0;
                for (int attempt = 1; attempt <= kMaxRetry; ++attempt) {
                    factor *= 1.25;
                    RunParams qq = q;
                    qq.attempt = attempt;
                    qq.S = (long long)std::ceil(factor * (double)q.S);
                    qq.maxS = 5 * qq.S;
                    const int P = q.R * q.K;
                    BN_CUDA(cudaMemsetAsync(r->n_unfrozen.p, 0, sizeof(int), r->stream));
                    walk_retry_kernel<<<(P + 127) / 128, 128, 0, r->stream>>>(qq, r->A);
                    BN_LAUNCH_CHECK();
                    BN_CUDA(cudaMemcpyAsync(r->h_unfrozen, r->n_unfrozen.p, sizeof(int), cudaMemcpyDeviceToHost, r->stream));
                    BN_CUDA(cudaStreamSynchronize(r->stream));
                    if (*r->h_unfrozen <= 0) break;
                    walk_until_frozen(qq, *r->h_unfrozen);
                }
            }
            ++done_batches;
            ++r->batches;
        }
        if (finished) *finished = r->finished ? 1 : 0;
    });
}

int binest_run_sizes(binest_run *r, int64_t run, int64_t *M, int64_t *n_deleted, int64_t *iterations, int64_t *evals) {
    return guard([&] {
        BN_REQUIRE(r && run >= 0 && run < r->prm.R, BINEST_ERR_DIMENSION, "run index out of range");
        BN_CUDA(cudaSetDevice(r->prob->device));
        fetch_state(*r);
        const RunState &s = r->h_state[run];
        if (M) *M = s.n_dead + r->prm.n;
        if (n_deleted) *n_deleted = s.n_dead;
        if (iterations) *iterations = s.iteration - 1;
        if (evals) *evals = r->evals;
    });
}

// Assemble the sorted sample list of one run: deleted points in order of removal + the live set sorted
// by {logL, point}.  If the run has not terminated, the batch in flight is first inserted (one update).
int binest_run_fetch(binest_run *r, int64_t run, double *points, double *logL, double *logPrior, double *acc,
                     int64_t *pool, double *logX, double *crude_logw, double *summary) {
    return guard([&] {
        BN_REQUIRE(r && run >= 0 && run < r->prm.R, BINEST_ERR_DIMENSION, "run index out of range");
        BN_CUDA(cudaSetDevice(r->prob->device));
        BN_REQUIRE(!r->first, BINEST_ERR_FUNCTION, "binest_run_fetch: advance the run first");
        if (!r->finished) launch_update(*r, true);  // insert the batch in flight and re-sort; no new kill
        fetch_state(*r);
        if (!points && !logL && !logPrior && !acc && !pool && !logX && !crude_logw && !summary) return;  // flush only
        const RunParams &q = r->prm;
        const RunState &s = r->h_state[run];
        const int64_t D = s.n_dead, n = q.n, M = D + n;
        const int d = q.d;
        std::vector<double> h_pts((size_t)M * d), h_L(M), h_Pr(M), h_acc(M);
        std::vector<int> h_pool(M), h_order(n);
        std::vector<double> l_th((size_t)n * d), l_L(n), l_Pr(n), l_acc(n);
        cudaStream_t st = r->stream;
        const size_t db = (size_t)run * q.cap, lb = (size_t)run * n;
        auto d2h = [&](void *dst, const void *src, size_t bytes) {
            if (bytes) BN_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st));
        };
        d2h(h_pts.data(), r->dead_theta.p + db * d, sizeof(double) * D * d);
        d2h(h_L.data(), r->dead_logL.p + db, sizeof(double) * D);
        d2h(h_Pr.data(), r->dead_logPr.p + db, sizeof(double) * D);
        d2h(h_acc.data(), r->dead_acc.p + db, sizeof(double) * D);
        d2h(h_pool.data(), r->dead_pool.p + db, sizeof(int) * D);
        d2h(h_order.data(), r->order.p + lb, sizeof(int) * n);
        d2h(l_th.data(), r->live_theta.p + lb * d, sizeof(double) * n * d);
        d2h(l_L.data(), r->live_logL.p + lb, sizeof(double) * n);
        d2h(l_Pr.data(), r->live_logPr.p + lb, sizeof(double) * n);
        d2h(l_acc.data(), r->live_acc.p + lb, sizeof(double) * n);
        BN_CUDA(cudaStreamSynchronize(st));
        for (int64_t j = 0; j < n; ++j) {
            const int src = h_order[j];
            std::memcpy(&h_pts[(size_t)(D + j) * d], &l_th[(size_t)src * d], sizeof(double) * d);
            h_L[D + j] = l_L[src]; h_Pr[D + j] = l_Pr[src]; h_acc[D + j] = l_acc[src];
            h_pool[D + j] = (int)(n - j);
        }
        if (points) std::memcpy(points, h_pts.data(), sizeof(double) * M * d);
        if (logL) std::memcpy(logL, h_L.data(), sizeof(double) * M);
        if (logPrior) std::memcpy(logPrior, h_Pr.data(), sizeof(double) * M);
        if (acc) std::memcpy(acc, h_acc.data(), sizeof(double) * M);
        if (pool) for (int64_t k = 0; k < M; ++k) pool[k] = h_pool[k];
        if (logX || crude_logw || summary) {
            DevBuf<double> dL(M), dX(M), dW(M), dS(4);
            DevBuf<int> dP(M);
            BN_CUDA(cudaMemcpyAsync(dL.p, h_L.data(), sizeof(double) * M, cudaMemcpyHostToDevice, st));
            BN_CUDA(cudaMemcpyAsync(dP.p, h_pool.data(), sizeof(int) * M, cudaMemcpyHostToDevice, st));
            crude_weights_kernel<<<1, 1024, 0, st>>>(M, n, dL.p, dP.p, nullptr, dX.p, dW.p, dS.p);
            BN_LAUNCH_CHECK();
            if (logX) d2h(logX, dX.p, sizeof(double) * M);
            if (crude_logw) d2h(crude_logw, dW.p, sizeof(double) * M);
            if (summary) d2h(summary, dS.p, sizeof(double) * 4);
            BN_CUDA(cudaStreamSynchronize(st));
        }
    });
}

int binest_run_estimates(binest_run *r, int64_t run, double *mean, double *cov) {
    return guard([&] {
        BN_REQUIRE(r && run >= 0 && run < r->prm.R, BINEST_ERR_DIMENSION, "run index out of range");
        BN_CUDA(cudaSetDevice(r->prob->device));
        fetch_state(*r);
        const RunState &s = r->h_state[run];
        const int d = r->prm.d;
        if (mean) std::memcpy(mean, s.meanEst, sizeof(double) * d);
        if (cov) std::memcpy(cov, s.covEst, sizeof(double) * d * d);
    });
}

// device time of the walk graphs so far (CUDA events on the run's stream), number of graphs, kernel launches/graph
int binest_run_timing(binest_run *r, double *walk_ms, int64_t *walk_graphs, int64_t *batches) {
    return guard([&] {
        BN_REQUIRE(r, BINEST_ERR_TYPE, "null run");
        if (walk_ms) *walk_ms = r->walk_ms;
        if (walk_graphs) *walk_graphs = r->walk_graphs;
        if (batches) *batches = r->batches;
    });
}

int binest_run_path(const binest_run *r, int *path) {
    return guard([&] {
        BN_REQUIRE(r && path, BINEST_ERR_TYPE, "null run");
        if (r->prob->op == BINEST_OP_GP_SE) *path = BINEST_WALK_STEPPED_GP;
        else if (r->prob->comm) *path = BINEST_WALK_STEPPED_SHARDED;
        else if (r->loop) *path = BINEST_WALK_DEVICE_LOOP;
        else if (r->resident) *path = BINEST_WALK_CLUSTER_RESIDENT;
        else if (r->grid) *path = BINEST_WALK_GRID_RESIDENT;
        else *path = BINEST_WALK_STEPPED_GRAPH;
    });
}

int binest_problem_stream(const binest_problem *p, void **stream) {
    return guard([&] {
        BN_REQUIRE(p && stream, BINEST_ERR_TYPE, "null problem");
        *stream = (void *)p->stream;
    });
}

// device state of a run group for csrc/merge.cu (run_view.cuh)
extern "C++" {
namespace binest {
void run_view(binest_run *r, RunView &v) {
    BN_REQUIRE(r, BINEST_ERR_TYPE, "null run");
    BN_CUDA(cudaSetDevice(r->prob->device));
    BN_REQUIRE(!r->first, BINEST_ERR_FUNCTION, "advance the run first");
    if (!r->finished) launch_update(*r, true);  // insert the batch in flight and re-sort; no new kill
    fetch_state(*r);
    BN_CUDA(cudaStreamSynchronize(r->stream));
    const RunParams &q = r->prm;
    v.device = r->prob->device;
    v.dev = RunViewDev{q.R, q.n, q.d, q.cap, (long long)q.first_run_id,
                       r->dead_theta.p, r->dead_logL.p, r->dead_logPr.p, r->dead_acc.p, r->dead_pool.p,
                       r->live_theta.p, r->live_logL.p, r->live_logPr.p, r->live_acc.p, r->order.p};
    v.n_dead.resize(q.R);
    for (int c = 0; c < q.R; ++c) v.n_dead[c] = r->h_state[c].n_dead;
}
}  // namespace binest
}  // extern "C++"

int binest_run_free(binest_run *r) {
    return guard([&] {
        if (r) { cudaSetDevice(r->prob->device); cudaStreamSynchronize(r->stream); }
        delete r;
    });
}

int binest_crude_weights(int64_t M, const double *logL, const int64_t *pool, int64_t n_live, double *logX,
                         double *crude_logw, double *summary) {
    return guard([&] {
        BN_REQUIRE(logL && pool && M >= 2 && n_live >= 1 && n_live <= M, BINEST_ERR_DIMENSION, "bad sample list");
        DevBuf<double> dL(M), dX(M), dW(M), dS(4);
        DevBuf<long long> dP(M);
        BN_CUDA(cudaMemcpy(dL.p, logL, sizeof(double) * M, cudaMemcpyHostToDevice));
        BN_CUDA(cudaMemcpy(dP.p, pool, sizeof(int64_t) * M, cudaMemcpyHostToDevice));
        crude_weights_kernel<<<1, 1024>>>(M, n_live, dL.p, nullptr, dP.p, dX.p, dW.p, dS.p);
        BN_LAUNCH_CHECK();
        if (logX) BN_CUDA(cudaMemcpy(logX, dX.p, sizeof(double) * M, cudaMemcpyDeviceToHost));
        if (crude_logw) BN_CUDA(cudaMemcpy(crude_logw, dW.p, sizeof(double) * M, cudaMemcpyDeviceToHost));
        if (summary) BN_CUDA(cudaMemcpy(summary, dS.p, sizeof(double) * 4, cudaMemcpyDeviceToHost));
    });
}

int binest_evidence_sampling(int64_t M, int64_t d, const double *points, const double *logL, const int64_t *pool,
                             int64_t n_live, int64_t post_runs, uint64_t seed, double *z, double *logw_mean,
                             double *logw_sd, double *slx_mean, double *slx_sd, double *pmean, double *H) {
    return guard([&] {
        BN_REQUIRE(points && logL && pool && M >= 2 && n_live >= 1 && n_live <= M && d >= 1, BINEST_ERR_DIMENSION,
                   "bad sample list");
        BN_REQUIRE(post_runs >= 2 && post_runs <= 65535, BINEST_ERR_DIMENSION, "2 <= PostProcessSamplingRuns <= 65535");
        const int R = (int)post_runs;
        // the two R x M work arrays (214 MB each for a merged C4 run) are kept between calls: returning them to the
        // pool and mapping them again stalled one call in three by ~0.5 s (bench r2h)
        static thread_local DevBuf<double> dSlx, dLw;
        if (dSlx.n < (size_t)R * M) { dSlx.alloc((size_t)R * M); dLw.alloc((size_t)R * M); }
        DevBuf<double> dPts((size_t)M * d), dL(M), dZ(R), dPm((size_t)R * d), dH(R);
        DevBuf<long long> dP(M);
        DevBuf<double> o1(M), o2(M), o3(M), o4(M);
        BN_CUDA(cudaMemcpy(dPts.p, points, sizeof(double) * M * d, cudaMemcpyHostToDevice));
        BN_CUDA(cudaMemcpy(dL.p, logL, sizeof(double) * M, cudaMemcpyHostToDevice));
        BN_CUDA(cudaMemcpy(dP.p, pool, sizeof(int64_t) * M, cudaMemcpyHostToDevice));
        evidence_sampling_kernel<<<R, 1024>>>(M, (int)d, n_live, dPts.p, dL.p, dP.p, seed, dSlx.p, dLw.p, dZ.p, dPm.p, dH.p);
        BN_LAUNCH_CHECK();
        evidence_moments_kernel<<<(unsigned)((M + 255) / 256), 256>>>(M, R, dSlx.p, dLw.p, dZ.p, o1.p, o2.p, o3.p, o4.p);
        BN_LAUNCH_CHECK();
        auto back = [&](double *dst, const double *src, size_t cnt) {
            if (dst) BN_CUDA(cudaMemcpy(dst, src, sizeof(double) * cnt, cudaMemcpyDeviceToHost));
        };
        back(z, dZ.p, R); back(logw_mean, o1.p, M); back(logw_sd, o2.p, M); back(slx_mean, o3.p, M);
        back(slx_sd, o4.p, M); back(pmean, dPm.p, (size_t)R * d); back(H, dH.p, R);
    });
}

// ---- measurement helpers for bench.py: inputs resident in HBM, CUDA events on the launching stream ----
// Evaluate the likelihood of P prior draws `reps` times (after `warmup`), optionally flushing L2 between
// repetitions; ms_kernel = average duration of loglike_stream_kernel alone, ms_total incl. the finalize kernel.
int binest_bench_loglike(binest_problem *p, int64_t P, int64_t reps, int64_t warmup, int flush_l2, double *ms_kernel,
                         double *ms_total) {
    return guard([&] {
        BN_REQUIRE(p && P > 0 && reps > 0, BINEST_ERR_DIMENSION, "bad arguments");
        BN_CUDA(cudaSetDevice(p->device));
        const int Ps = (int)((P + 31) & ~31LL);
        DevBuf<double> rows((size_t)P * p->d), soa((size_t)p->d * Ps), out(Ps), flush;
        sample_prior_kernel<<<(unsigned)((P + 127) / 128), 128, 0, p->stream>>>(p->prior, P, 900, 0, rows.p);
        BN_LAUNCH_CHECK();
        std::vector<double> h((size_t)P * p->d), hs((size_t)p->d * Ps, 1.0);
        BN_CUDA(cudaMemcpyAsync(h.data(), rows.p, sizeof(double) * h.size(), cudaMemcpyDeviceToHost, p->stream));
        BN_CUDA(cudaStreamSynchronize(p->stream));
        for (int64_t i = 0; i < P; ++i)
            for (int j = 0; j < p->d; ++j) hs[(size_t)j * Ps + i] = h[i * p->d + j];
        BN_CUDA(cudaMemcpyAsync(soa.p, hs.data(), sizeof(double) * hs.size(), cudaMemcpyHostToDevice, p->stream));
        const size_t flush_n = (size_t)256 << 20;  // 256 MiB > 126 MB L2
        if (flush_l2) flush.alloc(flush_n / sizeof(double));
        cudaEvent_t e0, e1, e2;
        BN_CUDA(cudaEventCreate(&e0)); BN_CUDA(cudaEventCreate(&e1)); BN_CUDA(cudaEventCreate(&e2));
        double tk = 0.0, tt = 0.0;
        if (p->op == BINEST_OP_GP_SE) {  // whole fill + Cholesky pipeline per repetition (ms_kernel == ms_total)
            for (int64_t it = 0; it < warmup + reps; ++it) {
                if (flush_l2) BN_CUDA(cudaMemsetAsync(flush.p, it & 0xff, flush_n, p->stream));
                BN_CUDA(cudaEventRecord(e0, p->stream));
                gp_loglike_device_strided(*p, soa.p, (int)P, Ps, out.p, 1, true);
                BN_CUDA(cudaEventRecord(e2, p->stream));
                BN_CUDA(cudaEventSynchronize(e2));
                float b = 0;
                BN_CUDA(cudaEventElapsedTime(&b, e0, e2));
                if (it >= warmup) { tk += b; tt += b; }
            }
        } else
        dispatch_op(*p, [&](auto op) {
            using OP = decltype(op);
            const StreamGeom g = stream_geom<OP>(*p, (int)P);
            DevBuf<double> partials((size_t)g.Gs * Ps);
            for (int64_t it = 0; it < warmup + reps; ++it) {
                if (flush_l2) BN_CUDA(cudaMemsetAsync(flush.p, it & 0xff, flush_n, p->stream));
                BN_CUDA(cudaEventRecord(e0, p->stream));
                launch_loglike<OP>(*p, soa.p, (int)P, Ps, partials.p, g, p->stream);
                BN_CUDA(cudaEventRecord(e1, p->stream));
                loglike_finalize_kernel<OP><<<(unsigned)((P * 32 + 255) / 256), 256, 0, p->stream>>>(
                    soa.p, (int)P, Ps, PartialView{partials.p, g.G, g.Gs, 1}, (double)p->rows, p->cst, p->prior, g_logzero, out.p);
                BN_LAUNCH_CHECK();
                BN_CUDA(cudaEventRecord(e2, p->stream));
                BN_CUDA(cudaEventSynchronize(e2));
                float a = 0, b = 0;
                BN_CUDA(cudaEventElapsedTime(&a, e0, e1));
                BN_CUDA(cudaEventElapsedTime(&b, e0, e2));
                if (it >= warmup) { tk += a; tt += b; }
            }
        });
        cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
        if (ms_kernel) *ms_kernel = tk / (double)reps;
        if (ms_total) *ms_total = tt / (double)reps;
    });
}

}  // extern "C"

// posterior sampler (createMCMCChain / iterateMCMC, BS:630-703): kernels + C ABI
#include "mcmc.cuh"
