// gp.cu — GP marginal likelihood operator (placeholder until the batched Cholesky lands).
#include "problem.cuh"
namespace binest {
void gp_loglike_device(binest_problem &, const double *, int, int, double *, bool) {
    throw Error(BINEST_ERR_FUNCTION, "GP operator not built yet");
}
}  // namespace binest
