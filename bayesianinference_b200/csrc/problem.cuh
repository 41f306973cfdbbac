// problem.cuh — host-side problem object: operator id + device-resident data + prior.
// This is the data-carrying half of defineInferenceProblem (BS:167-307): where the reference bakes
// the data matrix into a compiled function (BS:488-504, 576-593), the data are uploaded once here.
#pragma once
#include <cmath>
#include <vector>

#include "loglike.cuh"

struct binest_problem {
    int op = 0;
    int64_t iparam[4] = {0, 0, 0, 0};
    int d = 0;
    int64_t rows = 0;      // device rows (GBM: increments = n_rows - 1)
    int ncol = 0;          // fp64 columns per device row
    double cst = 0.0;      // parameter-independent additive constant
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

    ~binest_problem() {
        if (stream) cudaStreamDestroy(stream);
    }
};

namespace binest {

// choose the data split of the streaming kernel: G CTAs in x, each a contiguous even-sized row slice
struct StreamGeom {
    int nwarps, pgroups, G;
    long long rows_per_cta;
};
inline StreamGeom stream_geom(const binest_problem &p, int P) {
    StreamGeom g;
    g.nwarps = std::min(kMaxWarps, std::max(1, (P + 31) / 32));
    g.pgroups = (P + g.nwarps * 32 - 1) / (g.nwarps * 32);
    const int ctas_per_sm = 4;
    long long gmax = std::max(1, (p.num_sms * ctas_per_sm) / g.pgroups);
    long long rpc = (p.rows + gmax - 1) / gmax;
    rpc = std::max<long long>(rpc, 64);
    rpc = (rpc + 1) & ~1LL;
    g.rows_per_cta = rpc;
    g.G = (int)std::max<long long>(1, (p.rows + rpc - 1) / rpc);
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

// theta SoA [d][Ps] on the device -> out_dev[P]; runs on p.stream
template <class OP>
inline void launch_loglike(binest_problem &p, const double *theta_dev, int P, int Ps, double *partials_dev,
                           const StreamGeom &g) {
    dim3 grid(g.G, g.pgroups), block(g.nwarps * 32);
    loglike_stream_kernel<OP><<<grid, block, 0, p.stream>>>(p.data.p, p.rows, g.rows_per_cta, theta_dev, P, Ps,
                                                           partials_dev);
    BN_LAUNCH_CHECK();
}

void gp_loglike_device(binest_problem &p, const double *theta_dev, int P, int Ps, double *out_dev, bool check_box);

// full batched evaluation: theta_dev SoA [d][Ps] -> out_dev[P]
inline void loglike_device(binest_problem &p, const double *theta_dev, int P, int Ps, double *out_dev) {
    if (p.op == BINEST_OP_GP_SE) {
        gp_loglike_device(p, theta_dev, P, Ps, out_dev, true);
        return;
    }
    const StreamGeom g = stream_geom(p, P);
    const size_t need = (size_t)g.G * Ps;
    if (p.s_partials.n < need) p.s_partials.alloc(need);
    dispatch_op(p, [&](auto op) {
        using OP = decltype(op);
        launch_loglike<OP>(p, theta_dev, P, Ps, p.s_partials.p, g);
        loglike_finalize_kernel<OP><<<(P + 127) / 128, 128, 0, p.stream>>>(
            theta_dev, P, Ps, p.s_partials.p, g.G, (double)p.rows, p.cst, p.prior, g_logzero, out_dev);
        BN_LAUNCH_CHECK();
    });
}

}  // namespace binest
