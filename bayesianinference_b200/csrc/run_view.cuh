// run_view.cuh — read-only view of a run group's device state for csrc/merge.cu (the run object itself is private to
// engine.cu).
#pragma once
#include <vector>

#include "common.cuh"

namespace binest {

struct RunViewDev {
    int R, n, d;
    long long cap;            // dead-list capacity per run
    long long first_run_id;
    const double *dead_theta, *dead_logL, *dead_logPr, *dead_acc;  // [R][cap][d], [R][cap]
    const int *dead_pool;                                          // [R][cap]
    const double *live_theta, *live_logL, *live_logPr, *live_acc;  // [R][n][d], [R][n]
    const int *order;                                              // [R][n] live slots ascending by {logL, point}
};

struct RunView {
    RunViewDev dev;
    int device = 0;
    std::vector<long long> n_dead;  // per run
};

// flushes the batch in flight (as binest_run_fetch does), synchronises the run's stream and fills the view
void run_view(binest_run *r, RunView &v);

}  // namespace binest
