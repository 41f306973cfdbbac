// walk_grid.cuh — the whole S-step constrained walk (nsMCMC BS:707-745 for every walker of the batch) in ONE
// persistent launch over all SMs, for data sets that fit in the shared memories of the whole GPU
// (148 SMs x 227 KB = 33 MB; C2's 1e6 rows x 16 B = 16 MB is 54 KB per CTA at 2 CTAs/SM).
//
// Why.  The stepped path launches [walk_step, loglike_stream] per walk step: ~9 us of every ~98 us C2 step were
// launch latency, kernel drain/fill and the serial walk_step kernel (ncu/bench: 0.71 of the fp64 peak for the
// isolated kernel, 0.64 inside the graph).  Here nothing is launched inside a walk and the data rows never leave
// shared memory:
//   * grid = (resident CTAs per SM) x 148, launched cooperatively (co-residency guaranteed); CTA g loads its row
//     slice ONCE with the TMA engine (cp.async.bulk) and keeps it for all S steps;
//   * warps 0..7 ("data warps") reduce the slice for the proposals of the step exactly like loglike_stream_kernel
//     (lane = walker, TW walkers register-tiled per lane, rows broadcast from shared memory, step-major DFMA
//     order); the warp sums are combined in shared memory in a fixed order and published as partials[walker][g];
//   * warp 8 ("walker warp") of CTA g owns walkers g, g+G, ...: it sums the G partials of its walker in a fixed
//     order and applies the accept rule of nsDensity, the Haario recursion and the next Philox proposal, in the
//     split-phase form of GridWalker below (everything that does not need the likelihood is precomputed);
//   * the walkers are split into two sets A/B that alternate: while the walker warps process set A's step (two
//     grid-wide dependencies: partials -> chain logic -> proposals), the data warps already reduce set B.  The
//     grid barriers are split-phase (release-add on a monotone counter in L2, acquire-poll with back-off), so
//     their latency and the chain logic are off the fp64 critical path whenever both sets are non-empty.
// Every poll loop is bounded (~2 s): on expiry an abort flag is raised, all waits fall through and the host
// reports BINEST_ERR_CUDA instead of hanging the device.
#pragma once
#include "walk.cuh"

namespace binest {

constexpr int kGridGroupWarps = 8;                            // data warps per walker set
constexpr int kGridThreads = (2 * kGridGroupWarps + 2) * 32;  // two data groups + two walker warps

// Every counter on its own 256-byte line: polls of one counter do not queue behind the arrivals of another in the
// same L2 sector (all four in one 32-byte sector made every arrival wait behind ~6 polls/ns of hot-spot traffic).
struct GridSync {
    unsigned ctr[4][64];  // [X]: props_ready of set X (arrivals of walker warps: next proposals published)
                          // [2 + X]: partials_ready of set X (arrivals of CTAs: partial sums published)
    unsigned abort;
    unsigned pad_[63];
};
__device__ __forceinline__ unsigned *gs_props(GridSync *gs, int X) { return &gs->ctr[X][0]; }
__device__ __forceinline__ unsigned *gs_partials(GridSync *gs, int X) { return &gs->ctr[2 + X][0]; }

__device__ __forceinline__ void red_release_add_u32(unsigned *p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void fence_acquire_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

// one thread polls until *ctr >= target (monotone counter); false after an abort / time-out.  Relaxed polls with
// exponential back-off (every ld.acquire would also invalidate the SM's L1), one acquire fence on success.
__device__ __forceinline__ bool grid_wait(const unsigned *ctr, unsigned target, GridSync *gs, unsigned max_ns) {
    if (ld_relaxed_u32(ctr) >= target) { fence_acquire_gpu(); return true; }
    const long long t0 = clock64();
    unsigned ns = 64;
    for (unsigned it = 1;; ++it) {
        __nanosleep(ns);
        if (ns < max_ns) ns <<= 1;
        if (ld_relaxed_u32(ctr) >= target) { fence_acquire_gpu(); return true; }
        if ((it & 63u) == 0u) {
            if (ld_relaxed_u32(&gs->abort)) return false;
            if (clock64() - t0 > 4000000000LL) {  // ~2 s at 1.9 GHz
                atomicExch(&gs->abort, 1u);
                return false;
            }
        }
    }
}

// timeline trace: CTAs 0 and G/2, first 16 steps; slot [cta][role][s][X][4]
__device__ __forceinline__ void grid_trace(long long *dbg, int g, int G, int role, int s, int X, int ev) {
    if (dbg == nullptr || s >= 16 || (g != 0 && g != G / 2)) return;
    dbg[((((g ? 1 : 0) * 2 + role) * 16 + s) * 2 + X) * 4 + ev] = clock64();
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---------------------------------------------------------------------------------------------------
// Split-phase walker (one per own walker of the walker warp, in shared memory).  walk_step_walker does, per step,
// accept -> Haario -> Philox -> proposal -> prior, all after the partial sums arrive; inside the persistent kernel
// that whole chain (~40 us next to 16 fp64-bound data warps) sat on the grid-wide critical path.  Here everything
// that does not depend on the likelihood is done BEFORE the partial sums are awaited: the Philox normals / log u of
// the next step, BOTH possible next proposals (from x if the current proposal is rejected, from xn if accepted) with
// their box / prior tests, and the epilogue coefficients of the current proposal.  After the wait only the partial
// sum, one fused multiply-add, a compare, a select and D stores remain.  Same arithmetic, same Philox counters and
// the same operation order per value as walk_step_walker, so the trajectories coincide.
constexpr int kGridMaxOwn = 2;   // own walkers per walker warp and set (P <= 2 * 2 * G)
constexpr int kGridDrawChunk = 32;  // walk steps whose Philox draws are generated together, one step per lane

template <class OP>
struct GridWalker {
    static constexpr int D = OP::D, NZ = (OP::D + 1) / 2;
    double x[D], xn[D], cand[2][D], mean[D], cov[D * D];
    double z[kGridDrawChunk][2 * NZ], logu[kGridDrawChunk];
    double xPr, xL, nPr, nL, candPr[2];
    typename OP::Coef coef, candCoef[2];
    int ok, pre, candPre[2], candOk[2], steps, nacc, active, hasprop, sel, w, ndraw;
    unsigned walk_id, run_id;
};

template <class OP>
__device__ __forceinline__ void gw_init(GridWalker<OP> &q, const RunParams &prm, const RunArrays &A, int w, int lane) {
    constexpr int D = OP::D;
    const int K = prm.K;
    const int r = w / K, j = w - r * K;
    const RunState &st = A.state[r];
    const bool active = !st.done && j < st.Kb && !(A.w_flags[w] & WF_FROZEN);
    if (lane < D) {
        q.x[lane] = A.w_theta[(size_t)w * D + lane];
        q.mean[lane] = A.w_mean[(size_t)w * D + lane];
    }
    for (int e = lane; e < D * D; e += 32) q.cov[e] = A.w_cov[(size_t)w * D * D + e];
    if (lane == 0) {
        q.xPr = A.w_logPr[w]; q.xL = A.w_logL[w]; q.steps = A.w_steps[w]; q.nacc = A.w_nacc[w];
        q.active = active ? 1 : 0; q.hasprop = 0; q.sel = 0; q.w = w; q.pre = 0; q.ok = 0; q.nPr = 0.0; q.nL = 0.0;
        q.ndraw = 0;
        q.walk_id = (unsigned)(st.walk_base + j); q.run_id = prm.first_run_id + (unsigned)r;
    }
    __syncwarp();
}

// The draws of the next step (generated kGridDrawChunk steps at a time, lane = step) and both candidate
// proposals: lanes with (lane & 1) == 0 form cand[0] from x (current proposal rejected), the others cand[1] from
// xn (accepted), each with its box / prior test and epilogue coefficients.  The proposal generated while the chain
// has taken `steps` steps uses Philox counter word `steps`, as in walk_step_walker.
template <class OP>
__device__ __forceinline__ void gw_pre(GridWalker<OP> &q, const RunParams &prm, const RunArrays &A, const PriorSpec &prior,
                                       const OpCst &cst, int lane) {
    constexpr int D = OP::D, NZ = (D + 1) / 2;
    if (!q.active) return;
    const RunState &st = A.state[q.w / prm.K];
    const int slot = q.ndraw % kGridDrawChunk;
    if (slot == 0) {
        const uint32_t c = (uint32_t)(q.steps + (q.hasprop ? 1 : 0) + lane);
#pragma unroll
        for (int b = 0; b < NZ; ++b) {
            double za, zb;
            rng_normal2(prm.seed, (uint32_t)(b + 16 * prm.attempt), c, q.walk_id, TAG_NORMAL, q.run_id, za, zb);
            q.z[lane][2 * b] = za;
            q.z[lane][2 * b + 1] = zb;
        }
        double u0, u1;
        rng_uniform2(prm.seed, (uint32_t)(16 * prm.attempt), c, q.walk_id, TAG_ACCEPT, q.run_id, u0, u1);
        q.logu[lane] = log(u0);
        __syncwarp();
    }
    const int v = lane & 1;
    const double logu = q.logu[slot];
    const double basePr = v ? q.nPr : q.xPr;
    double cd[D];
#pragma unroll
    for (int a = 0; a < D; ++a) {
        double s = v ? q.xn[a] : q.x[a];
        if (st.chol_ok) {
#pragma unroll
            for (int b = 0; b <= a; ++b) s += st.cholL[a * D + b] * q.z[slot][b];
        }
        cd[a] = s;
    }
    double cpr = 0.0;
    int cpre = 0;
    if (in_box<D>(prior, cd)) {
        double nPr = 0.0;
#pragma unroll
        for (int a = 0; a < D; ++a) nPr += logprior_dim(prior, a, cd[a]);
        if (!isfinite(nPr)) nPr = prm.logzero;
        cpr = nPr;
        cpre = (nPr - basePr > logu) ? 1 : 0;  // Metropolis rule on the log density
    }
    bool ok = false;
    const typename OP::Coef cf = OP::prepare(cd, ok, cst);
    __syncwarp();
    if (lane < 2) {
#pragma unroll
        for (int a = 0; a < D; ++a) q.cand[v][a] = cd[a];
        q.candPr[v] = cpr; q.candPre[v] = cpre; q.candCoef[v] = cf; q.candOk[v] = ok ? 1 : 0;
    }
    if (lane == 0) q.ndraw += 1;
    __syncwarp();
}

// after the partial sums of the current proposal arrived: decide, publish the next proposal
template <class OP>
__device__ __forceinline__ void gw_post(GridWalker<OP> &q, const RunParams &prm, const RunArrays &A, const PartialView &pv,
                                        double rows, const OpCst &cst, bool publish, int lane) {
    constexpr int D = OP::D;
    if (!q.active) return;
    int sel = 0;
    double nL = 0.0;
    if (q.hasprop && q.pre) {
        const double sum = combine_partials_warp(pv, q.w, lane);
        nL = op_finish<OP>(q.coef, sum, rows, cst);
        if (!(q.ok && isfinite(nL))) nL = prm.logzero;      // RuntimeErrorHandler -> logzero, BS:500-503
        if (nL > A.state[q.w / prm.K].Lstar) sel = 1;       // nsDensity: logL > threshold, strict (BS:605)
    }
    if (publish && lane < D) __stcg(A.w_prop + (size_t)lane * prm.Ps + q.w, q.cand[sel][lane]);
    if (lane == 0) { q.sel = sel; q.nL = nL; }
    __syncwarp();
}

// bookkeeping after the proposal is out: accept, Haario recursion (BS:715-727), adopt the published proposal
template <class OP>
__device__ __forceinline__ void gw_update(GridWalker<OP> &q, int lane) {
    constexpr int D = OP::D;
    if (!q.active) return;
    const int sel = q.sel, hasprop = q.hasprop, steps = q.steps;
    double x[D], nx[D];
#pragma unroll
    for (int a = 0; a < D; ++a) {
        x[a] = (hasprop && sel) ? q.xn[a] : q.x[a];
        nx[a] = q.cand[sel][a];
    }
    const double nPr = q.candPr[sel];
    const int pre = q.candPre[sel], cok = q.candOk[sel];
    const typename OP::Coef cf = q.candCoef[sel];
    double mo[D];
#pragma unroll
    for (int a = 0; a < D; ++a) mo[a] = q.mean[a];
    double covv[(D * D + 31) / 32];
#pragma unroll
    for (int e = 0; e < (D * D + 31) / 32; ++e) covv[e] = (lane + 32 * e < D * D) ? q.cov[lane + 32 * e] : 0.0;
    const double accPr = q.nPr, accL = q.nL;
    __syncwarp();
    if (hasprop) {
        if (sel && lane == 0) { q.xPr = accPr; q.xL = accL; q.nacc += 1; }
        const double t = 10.0 + (double)steps;
        double dm_o[D], dm_n[D];
#pragma unroll
        for (int a = 0; a < D; ++a) {
            const double mn = mo[a] + (x[a] - mo[a]) / (t + 1.0);
            dm_o[a] = x[a] - mo[a];
            dm_n[a] = x[a] - mn;
            if (lane == a) { q.mean[a] = mn; q.x[a] = x[a]; }
        }
        const double f = (t - 1.0) / t;
#pragma unroll
        for (int e = 0; e < (D * D + 31) / 32; ++e) {
            const int idx = lane + 32 * e;
            if (idx < D * D) {
                const int a = idx / D, b = idx - a * D;
                double da = 0.0, db = 0.0;
#pragma unroll
                for (int k = 0; k < D; ++k) { if (k == a) da = dm_o[k]; if (k == b) db = dm_n[k]; }
                q.cov[idx] = f * covv[e] + da * db / t;
            }
        }
        if (lane == 0) q.steps = steps + 1;
    }
    if (lane < D) q.xn[lane] = nx[lane < D ? lane : 0];
    if (lane == 0) { q.nPr = nPr; q.pre = pre; q.hasprop = 1; q.coef = cf; q.ok = cok; }
    __syncwarp();
}

// end of the S-step block: chain state back to global memory; freeze per BS:730-736
template <class OP>
__device__ __forceinline__ void gw_final(GridWalker<OP> &q, const RunParams &prm, const RunArrays &A, int lane) {
    constexpr int D = OP::D;
    if (!q.active) return;
    const int w = q.w;
    if (lane < D) {
        A.w_theta[(size_t)w * D + lane] = q.x[lane];
        A.w_mean[(size_t)w * D + lane] = q.mean[lane];
    }
    for (int e = lane; e < D * D; e += 32) A.w_cov[(size_t)w * D * D + e] = q.cov[e];
    if (lane == 0) {
        A.w_logL[w] = q.xL; A.w_logPr[w] = q.xPr; A.w_nacc[w] = q.nacc; A.w_steps[w] = q.steps;
        int flags = 0;
        if (q.steps % prm.S == 0) {
            const double rate = (double)q.nacc / (double)q.steps;
            if ((rate >= prm.acc_min && rate <= prm.acc_max) || q.steps >= prm.maxS) {
                flags |= WF_FROZEN;
                atomicSub(A.n_unfrozen, 1);
            }
        }
        A.w_flags[w] = flags;
    }
}

// dynamic shared memory (doubles): tile[rows_per_cta * NCOL (even)] | red[2 groups][kGridGroupWarps][32 * TW]
template <class OP, int TW>
__host__ __device__ inline size_t grid_smem_bytes(long long rows_per_cta) {
    const size_t tile = (((size_t)rows_per_cta * OP::NCOL + 1) & ~(size_t)1);
    return (tile + (size_t)2 * kGridGroupWarps * 32 * TW) * sizeof(double);
}

// One CTA per SM: warps [0, 8) = data group of set A, [8, 16) = data group of set B, warp 16 / 17 = walker warp of
// set A / B.  Both groups sweep ALL rows of the CTA's resident slice, each for its own walkers, and run
// independently of each other: while one group sits in its per-step overhead (grid wait, proposal loads, cross-warp
// combine, publish) the other has the whole fp64 pipe.  (History, all measured on C2: two alternating sets inside
// 8-warp CTAs at 2 CTAs/SM fell into a schedule where the co-resident CTAs took turns — each alone on the SM, its
// overhead never hidden, 87 us/step; 16 warps sweeping one set at a time: no waits but 95 us/step.)
// passesA: walker passes (of 32*TW walkers each) in set A; the remaining passes form set B (may be empty).
template <class OP, int TW>
__global__ void __launch_bounds__(kGridThreads, 1)
walk_grid_kernel(const __grid_constant__ RunParams prm, RunArrays A, const __grid_constant__ PriorSpec prior,
                 const double *__restrict__ data, long long rows, long long rows_per_cta, const OpCst cst,
                 double *__restrict__ partials /* [Ps][Gs] */, int Gs, int passes, int passesA,
                 GridSync *gs, long long *dbg /* timeline trace (BINEST_GRID_TRACE), or nullptr */) {
    constexpr int NCOL = OP::NCOL, D = OP::D, WP = 32 * TW, GW = kGridGroupWarps;
    extern __shared__ __align__(128) double smem[];
    const size_t tile_sz = (((size_t)rows_per_cta * NCOL + 1) & ~(size_t)1);
    double *tile = smem;
    __shared__ uint64_t full;
    __shared__ GridWalker<OP> slots[2][kGridMaxOwn];

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int G = gridDim.x, g = blockIdx.x;
    const int P = prm.R * prm.K, Ps = prm.Ps, S = (int)prm.S;
    const int PA = min(passesA * WP, P);  // walkers [0, PA) are set A, [PA, P) set B

    const long long r0 = (long long)g * rows_per_cta;
    const long long r1 = (r0 + rows_per_cta < rows) ? r0 + rows_per_cta : rows;
    const int nr = r1 > r0 ? (int)(r1 - r0) : 0;

    // nothing to walk (every run of the group has terminated: the host enqueues iterations ahead of reading the state):
    // the run states are constant during the launch, so all CTAs take the same exit and no barrier is ever armed
    {
        int live = 0;
        for (int rr = tid; rr < prm.R; rr += blockDim.x) live |= (!A.state[rr].done && A.state[rr].Kb > 0) ? 1 : 0;
        if (!__syncthreads_or(live)) return;
    }
    // ---- the row slice of this CTA -> shared memory, once (TMA bulk copies, one mbarrier phase)
    if (tid == 0) {
        mbar_init(&full, 1);
        mbar_fence_init();
        const uint32_t total = ((uint32_t)(nr * NCOL * 8) + 15u) & ~15u;  // buffer is padded at upload
        if (total) {
            mbar_expect_tx(&full, total);
            for (uint32_t off = 0; off < total; off += 32768u) {
                const uint32_t b = (total - off < 32768u) ? total - off : 32768u;
                bulk_g2s(reinterpret_cast<char *>(tile) + off, reinterpret_cast<const char *>(data + r0 * NCOL) + off, b,
                         &full);
            }
        }
    }
    __syncthreads();

    if (wid >= 2 * GW) {
        // =============================== walker warp of set X ===============================
        const int X = wid - 2 * GW;
        const int wlo = X ? PA : 0, whi = X ? P : PA;
        if (wlo >= whi) return;  // empty set: nobody waits on its counters
        const PartialView pv{partials, G, (long long)Gs, 1};
        int w0 = g;
        while (w0 < wlo) w0 += G;
        const int nown = w0 < whi ? (whi - w0 + G - 1) / G : 0;
        for (int k = 0; k < nown; ++k) {
            gw_init<OP>(slots[X][k], prm, A, w0 + k * G, lane);
            gw_pre<OP>(slots[X][k], prm, A, prior, cst, lane);
        }
        for (int s = 0; s <= S; ++s) {
            if (lane == 0) grid_trace(dbg, g, G, 0, s, X, 0);
            if (s > 0) {
                if (lane == 0) grid_wait(gs_partials(gs, X), (unsigned)G * (unsigned)s, gs, 512u);
                __syncwarp();
            }
            if (lane == 0) grid_trace(dbg, g, G, 0, s, X, 1);
            for (int k = 0; k < nown; ++k) gw_post<OP>(slots[X][k], prm, A, pv, (double)rows, cst, s < S, lane);
            if (s < S && lane == 0) {
                __threadfence();
                red_release_add_u32(gs_props(gs, X), 1u);
            }
            if (lane == 0) grid_trace(dbg, g, G, 0, s, X, 2);
            // off the critical path: bookkeeping, then the draws and candidates of the next step
            for (int k = 0; k < nown; ++k) {
                gw_update<OP>(slots[X][k], lane);
                if (s < S) gw_pre<OP>(slots[X][k], prm, A, prior, cst, lane);
                else gw_final<OP>(slots[X][k], prm, A, lane);
            }
            if (lane == 0) grid_trace(dbg, g, G, 0, s, X, 3);
        }
        return;
    }

    // =============================== data group of set X ===============================
    // The 8 warps of the group split the rows of the slice, combine their sums in shared memory in a fixed order and
    // publish partials[walker][g].  (Barrier-free variants — one partial per warp, or the last warp combining — were
    // slower: the walker then sums 8x more partials, or the kernel needs too many registers.)
    const int X = wid / GW, gwid = wid - X * GW, gtid = tid - X * GW * 32;
    const int pass0 = X ? passesA : 0, pass1 = X ? passes : passesA;
    if (pass0 >= pass1) return;
    double *red = smem + tile_sz + (size_t)X * GW * WP;  // [GW][WP]
    const int bar_id = 1 + X;
    if (nr > 0) mbar_wait(&full, 0);
    for (int s = 0; s < S; ++s) {
        if (gtid == 0) {
            grid_trace(dbg, g, G, 1, s, X, 0);
            grid_wait(gs_props(gs, X), (unsigned)G * (unsigned)(s + 1), gs, 256u);
            grid_trace(dbg, g, G, 1, s, X, 1);
        }
        named_bar_sync(bar_id, GW * 32);
        for (int pass = pass0; pass < pass1; ++pass) {
            const int wbase = pass * WP;
            typename OP::Row c[TW];
#pragma unroll
            for (int t = 0; t < TW; ++t) {
                const int w = wbase + lane + 32 * t;
                double th[D];
#pragma unroll
                for (int j = 0; j < D; ++j) th[j] = (w < P) ? __ldcg(A.w_prop + (size_t)j * Ps + w) : 1.0;
                c[t] = OP::make_row(th, cst);
            }
            typename OP::Acc acc[TW];
#pragma unroll
            for (int t = 0; t < TW; ++t) acc[t] = OP::acc_init();
            sweep_rows<OP, TW>(c, tile, gwid, GW, nr, acc);
#pragma unroll
            for (int u = 0; u < TW; ++u) red[gwid * WP + lane + 32 * u] = OP::acc_value(acc[u]);
            named_bar_sync(bar_id, GW * 32);
            for (int k = gtid; k < WP; k += GW * 32) {
                double sum = 0.0;
#pragma unroll
                for (int q = 0; q < GW; ++q) sum += red[q * WP + k];
                const int w = wbase + k;
                if (w < Ps) __stcg(partials + (size_t)w * Gs + g, sum);
            }
            named_bar_sync(bar_id, GW * 32);  // red is free again; all partial stores are issued
        }
        if (gtid == 0) {
            grid_trace(dbg, g, G, 1, s, X, 2);
            __threadfence();
            red_release_add_u32(gs_partials(gs, X), 1u);
            grid_trace(dbg, g, G, 1, s, X, 3);
        }
    }
}

}  // namespace binest
