// problem.cuh — host-side problem object: operator id + device-resident data + prior.
// This is the data-carrying half of defineInferenceProblem (BS:167-307): where the reference bakes
// the data matrix into a compiled function (BS:488-504, 576-593), the data are uploaded once here.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <type_traits>
#include <vector>

#include "loglike.cuh"

#include "xchg.cuh"

// sharded modes: one communicator per process, created by binest_comm_create (comm.cu)
struct binest_comm {
    void *nccl = nullptr;  // ncclComm_t: set-up traffic (IPC handles, data constants) and the fallback exchange
    int rank = 0, world = 1, device = 0;
    // in-kernel exchange over peer-mapped memory (xchg.cuh); peer == false: host-issued ncclAllGather per exchange
    bool peer = false;
    binest::XchgDev xd{};
    double *xbuf_local = nullptr;
    unsigned *xflag_local = nullptr;
    binest::XchgState *xst = nullptr;
    unsigned *h_abort = nullptr;  // pinned mirror of xst->abort
    binest::XchgDev *xd_dev = nullptr;  // device copy of xd (consumer kernels take a pointer: nullptr = unsharded)
    int64_t exchanges = 0;        // exchanges issued so far (host count; NVLink bytes = 8 * count * (world - 1) each)
    int64_t bytes_pushed = 0;     // payload bytes this rank has stored into peers' buffers
};

struct binest_problem {
    int op = 0;
    int64_t iparam[4] = {0, 0, 0, 0};
    int d = 0;
    int64_t rows = 0;      // device rows (GBM: increments = n_rows - 1)
    int ncol = 0;          // fp64 columns per device row
    binest::OpCst cst{};   // parameter-independent constants of the data (additive constant, moments)
    binest::DevBuf<double> data;
    binest::PriorSpec prior{};
    int device = 0;
    int num_sms = 148;
    cudaStream_t stream = nullptr;
    // GP operator keeps the raw inputs/outputs (gp.cu)
    int64_t gp_n = 0, gp_dim = 0;
    binest::DevBuf<double> gp_x, gp_y;
    // scratch reused by the batched entry points
    binest::DevBuf<double> s_theta, s_partials, s_out;
    // data-sharded mode: this problem holds rows [shard of the data]; logL = Sum over ranks of the shard sums
    binest_comm *comm = nullptr;
    binest_comm *comm_batch = nullptr;  // batch-sharded mode (GP): the theta batch is split across the ranks
    double rows_total = 0.0;
    binest::OpCst cst_total{};                  // over all shards (operator epilogues need the global values)
    binest::DevBuf<double> sh_send, sh_recv;    // [Ps], [world][Ps]
    double rows_eff() const { return comm ? rows_total : (double)rows; }
    const binest::OpCst &cst_eff() const { return comm ? cst_total : cst; }

    ~binest_problem() {
        if (stream) cudaStreamDestroy(stream);
    }
};

namespace binest {

// launch geometry of the streaming kernel: TW walkers per lane, G data slices (CTAs in x), pgroups walker
// groups (CTAs in y); every slice is a contiguous, even-sized row range
struct StreamGeom {
    int tw, pgroups, G, Gs;
    long long rows_per_cta;
};

template <class OP, class F>
inline void dispatch_tw(int tw, F &&f);

template <class OP>
inline StreamGeom stream_geom(const binest_problem &p, int P) {
    StreamGeom g;
    const int lanesets = (P + 31) / 32;
    g.tw = 1;
    static const int tw_cap = [] { const char *e = getenv("BINEST_TW_MAX"); return e ? atoi(e) : 8; }();  // experiments
    while (g.tw < OP::TW_MAX && g.tw < lanesets && 2 * g.tw <= tw_cap) g.tw <<= 1;
    g.pgroups = (P + 32 * g.tw - 1) / (32 * g.tw);
    // exactly one wave: grid = SMs x resident CTAs of this instantiation, all slices the same size
    // (a 592-CTA grid at 3 resident CTAs/SM ran 1.33 waves and left the fp64 pipe 22 % idle in the tail)
    int ctas_per_sm = 1;
    dispatch_tw<OP>(g.tw, [&](auto twc) {
        constexpr int TW = decltype(twc)::value;
        BN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, loglike_stream_kernel<OP, TW>, kWarps * 32, 0));
    });
    ctas_per_sm = std::max(1, ctas_per_sm);
    const long long gmax = std::max(1, (p.num_sms * ctas_per_sm) / g.pgroups);
    long long rpc = (p.rows + gmax - 1) / gmax;
    rpc = std::max<long long>(rpc, 8 * kWarps);
    rpc = (rpc + 1) & ~1LL;
    g.rows_per_cta = rpc;
    g.G = (int)std::max<long long>(1, (p.rows + rpc - 1) / rpc);
    g.Gs = (g.G + 3) & ~3;
    return g;
}

// invoke f(OP{}) for the operator type of the problem
template <class F>
inline void dispatch_op(const binest_problem &p, F &&f) {
    switch (p.op) {
    case BINEST_OP_GAUSSIAN_IID: f(OpGaussian{}); return;
    case BINEST_OP_GBM: f(OpGbm{}); return;
    case BINEST_OP_POLYREG:
        switch (p.iparam[0]) {
        case 1: f(OpPolyReg<1>{}); return;
        case 2: f(OpPolyReg<2>{}); return;
        case 3: f(OpPolyReg<3>{}); return;
        case 4: f(OpPolyReg<4>{}); return;
        case 5: f(OpPolyReg<5>{}); return;
        }
        break;
    case BINEST_OP_LOGISTIC: {
        const int F_ = (int)p.iparam[2], K_ = (int)p.iparam[1];
        if (K_ == 3 && F_ == 4) { f(OpLogistic<4, 3>{}); return; }
        if (K_ == 2 && F_ == 4) { f(OpLogistic<4, 2>{}); return; }
        if (K_ == 3 && F_ == 2) { f(OpLogistic<2, 3>{}); return; }
        if (K_ == 2 && F_ == 2) { f(OpLogistic<2, 2>{}); return; }
        if (K_ == 2 && F_ == 1) { f(OpLogistic<1, 2>{}); return; }
        break;
    }
    }
    throw Error(BINEST_ERR_FUNCTION, "operator/shape not in the fixed operator table");
}

// invoke f(integral_constant<int, TW>) for tw in {1, 2, 4, 8} and tw <= OP::TW_MAX
template <class OP, class F>
inline void dispatch_tw(int tw, F &&f) {
    if (tw == 1) { f(std::integral_constant<int, 1>{}); return; }
    if constexpr (OP::TW_MAX >= 2) if (tw == 2) { f(std::integral_constant<int, 2>{}); return; }
    if constexpr (OP::TW_MAX >= 4) if (tw == 4) { f(std::integral_constant<int, 4>{}); return; }
    if constexpr (OP::TW_MAX >= 8) if (tw == 8) { f(std::integral_constant<int, 8>{}); return; }
    throw Error(BINEST_ERR_FUNCTION, "bad walker tiling");
}

// theta SoA [d][Ps] on the device -> partials[Ps][Gs]; runs on stream s (also used under graph capture)
// launch configuration with the programmatic-dependent-launch attribute (see pdl_wait in common.cuh)
struct PdlConfig {
    cudaLaunchConfig_t cfg{};
    cudaLaunchAttribute attr[1];
    PdlConfig(dim3 grid, dim3 block, cudaStream_t s, bool pdl) {
        cfg.gridDim = grid;
        cfg.blockDim = block;
        cfg.dynamicSmemBytes = 0;
        cfg.stream = s;
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = pdl ? 1 : 0;
    }
};

template <class OP>
inline void launch_loglike(binest_problem &p, const double *theta_dev, int P, int Ps, double *partials_dev,
                           const StreamGeom &g, cudaStream_t s, bool check = true, bool pdl = false) {
    dim3 grid(g.G, g.pgroups), block(kWarps * 32);
    dispatch_tw<OP>(g.tw, [&](auto twc) {
        constexpr int TW = decltype(twc)::value;
        PdlConfig lc(grid, block, s, pdl);
        const double *data = p.data.p;
        long long rows = p.rows, rpc = g.rows_per_cta;
        int Gs = g.Gs;
        OpCst cst = p.cst;  // this GPU's data constants (pivots of the polynomial operator)
        BN_CUDA(cudaLaunchKernelEx(&lc.cfg, loglike_stream_kernel<OP, TW>, data, rows, rpc, theta_dev, P, Ps,
                                   partials_dev, Gs, cst));
    });
    if (check) BN_LAUNCH_CHECK();
}

void gp_loglike_device(binest_problem &p, const double *theta_dev, int P, int Ps, double *out_dev, bool check_box);
void gp_predict_device(binest_problem &p, const double *theta_dev, int P, int Ps, const double *xs_dev, int Q,
                       double *mean_dev, double *sd_dev);

// comm.cu: all-gather of `count` doubles per rank on stream s (NCCL over NVLink)
void comm_allgather_f64(binest_comm &c, const double *send, double *recv, size_t count, cudaStream_t s);

// Data-sharded exchange after a likelihood launch: reduce this rank's per-CTA partials to one sum per walker in a
// fixed order, all-gather the P sums of every rank, and hand the consumer a view over [world][Ps].  Every rank then
// adds the same `world` numbers in the same (rank) order, so accept/reject decisions are bit-identical everywhere —
// which an all-reduce would not guarantee across algorithms.  Traffic: 8 P bytes per rank per step.
// true when a consumer of the in-kernel exchange timed out (a peer died or left the lock step); host side, after a sync
inline void comm_check_abort(binest_comm &c, cudaStream_t s) {
    if (!c.peer) return;
    BN_CUDA(cudaMemcpyAsync(c.h_abort, &c.xst->abort, sizeof(unsigned), cudaMemcpyDeviceToHost, s));
    BN_CUDA(cudaStreamSynchronize(s));
    BN_REQUIRE(*c.h_abort == 0, BINEST_ERR_CUDA, "sharded exchange timed out: a peer rank did not publish its values");
}

// Data-sharded exchange after a likelihood launch: reduce this rank's per-CTA partials to one value per walker in
// shard-additive form (OP::local with this rank's data constants) and make the P values of every rank visible on every
// rank.  Peer path (default): shard_reduce_push_kernel stores them straight into all peers' receive buffers and raises
// the flags (xchg.cuh); the consumer kernel gets the XchgDev and waits in-kernel.  Fallback (BINEST_XCHG=nccl or no
// peer access): ncclAllGather on the stream.  Every rank then adds the same `world` numbers in rank order, so
// accept/reject decisions are bit-identical everywhere.  Traffic: 8 P bytes to each peer per step.
template <class OP>
inline PartialView shard_exchange(binest_problem &p, const double *partials, const double *theta_dev, int P, int Ps,
                                  const StreamGeom &g, cudaStream_t s) {
    binest_comm &c = *p.comm;
    c.exchanges += 1;
    c.bytes_pushed += (int64_t)8 * Ps * (c.world - 1);
    if (c.peer) {
        BN_REQUIRE(Ps <= kXchgSlotDoubles, BINEST_ERR_DIMENSION, "too many walkers for the sharded exchange buffer");
        shard_reduce_push_kernel<OP><<<(Ps * 32 + 255) / 256, 256, 0, s>>>(PartialView{partials, g.G, g.Gs, 1}, theta_dev, P,
                                                                          Ps, (double)p.rows, p.cst, c.xd);
        BN_LAUNCH_CHECK();
        return PartialView{nullptr, c.world, 1, kXchgSlotDoubles, 1};  // resolved in-kernel from the XchgDev
    }
    if (p.sh_send.n < (size_t)Ps) p.sh_send.alloc(Ps);
    if (p.sh_recv.n < (size_t)Ps * c.world) p.sh_recv.alloc((size_t)Ps * c.world);
    shard_reduce_kernel<OP><<<(Ps * 32 + 255) / 256, 256, 0, s>>>(PartialView{partials, g.G, g.Gs, 1}, theta_dev, P, Ps,
                                                                 (double)p.rows, p.cst, p.sh_send.p);
    BN_LAUNCH_CHECK();
    comm_allgather_f64(c, p.sh_send.p, p.sh_recv.p, (size_t)Ps, s);
    return PartialView{p.sh_recv.p, c.world, 1, Ps, 1};
}

// full batched evaluation: theta_dev SoA [d][Ps] -> out_dev[P]
inline void loglike_device(binest_problem &p, const double *theta_dev, int P, int Ps, double *out_dev) {
    if (p.op == BINEST_OP_GP_SE) {
        gp_loglike_device(p, theta_dev, P, Ps, out_dev, true);
        return;
    }
    dispatch_op(p, [&](auto op) {
        using OP = decltype(op);
        const StreamGeom g = stream_geom<OP>(p, P);
        const size_t need = (size_t)g.Gs * Ps;
        if (p.s_partials.n < need) p.s_partials.alloc(need);
        launch_loglike<OP>(p, theta_dev, P, Ps, p.s_partials.p, g, p.stream);
        PartialView pv{p.s_partials.p, g.G, g.Gs, 1};
        if (p.comm) pv = shard_exchange<OP>(p, p.s_partials.p, theta_dev, P, Ps, g, p.stream);
        const XchgDev *xd = (p.comm && p.comm->peer) ? p.comm->xd_dev : nullptr;
        loglike_finalize_kernel<OP><<<(P * 32 + 255) / 256, 256, 0, p.stream>>>(
            theta_dev, P, Ps, pv, p.rows_eff(), p.cst_eff(), p.prior, g_logzero, out_dev, xd);
        BN_LAUNCH_CHECK();
    });
}

}  // namespace binest
