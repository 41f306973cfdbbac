// predictive.cuh — component parameters of predictiveDistribution for regression problems (BS:1437-1483).
//
// The reference builds, for every input x_q, MixtureDistribution[weights, {dist[theta_m, x_q]}_m] where dist is the
// "GeneratingDistribution" with the parameters of sample m substituted (expressionToFunction, BS:1450-1462).  The
// numerical content is the M x Q table of component parameters; this kernel fills it for the operators that have
// independent variables:
//   polynomial regression   NormalDistribution[Sum_j c_j x^j, sigma]         -> (mean, sd)            C = 2
//   softmax classification  class probabilities p_k = exp z_k / Sum exp z    -> (p_1 .. p_K)          C = K
// One thread per (sample, input) pair, inputs fastest: theta is a warp-wide broadcast, the output is written
// contiguously.  Bound: HBM write, 8 C bytes per pair.
#pragma once
#include "common.cuh"

namespace binest {

constexpr int kPredMaxD = 64;

__global__ void __launch_bounds__(256)
predictive_kernel(int op, int deg, int K, int F, const double *__restrict__ theta /* SoA [d][Ms] */, long long M,
                  int Ms, const double *__restrict__ xin /* [Q][F] */, long long Q, double *__restrict__ out) {
    const long long total = M * Q;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long m = idx / Q, q = idx - m * Q;
        if (op == BINEST_OP_POLYREG) {
            const double x = xin[q];
            double t = theta[(size_t)deg * Ms + m];
            for (int j = deg - 1; j >= 0; --j) t = fma(t, x, theta[(size_t)j * Ms + m]);  // Horner, as the likelihood
            const double sg = theta[(size_t)(deg + 1) * Ms + m];
            const bool ok = sg > 0.0;  // BS:523
            const double nan = __longlong_as_double(0x7ff8000000000000LL);
            reinterpret_cast<double2 *>(out)[idx] = make_double2(ok ? t : nan, ok ? sg : nan);
        } else {  // BINEST_OP_LOGISTIC: z_k = b_k + w_k . x (k < K), z_K = 0
            double z[kPredMaxD];
            double zmax = 0.0;
            for (int k = 0; k < K - 1; ++k) {
                double s = theta[(size_t)(k * (F + 1) + F) * Ms + m];
                for (int f = 0; f < F; ++f) s = fma(theta[(size_t)(k * (F + 1) + f) * Ms + m], xin[q * F + f], s);
                z[k] = s;
                zmax = fmax(zmax, s);
            }
            double den = exp(-zmax);
            for (int k = 0; k < K - 1; ++k) { z[k] = exp(z[k] - zmax); den += z[k]; }
            double *o = out + idx * K;
            for (int k = 0; k < K - 1; ++k) o[k] = z[k] / den;
            o[K - 1] = exp(-zmax) / den;
        }
    }
}

}  // namespace binest
