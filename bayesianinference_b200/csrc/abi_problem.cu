// abi_problem.cu — library state, problem definition and the batched operator entry points of the
// C ABI (include/binest.h).  No CPU fallback: every compute call needs a CUDA device.
#include <algorithm>
#include <cstring>
#include <functional>
#include <memory>

#include "predictive.cuh"
#include "problem.cuh"

namespace binest {
std::atomic<int64_t> g_launches{0};
double g_logzero = -1.7976931348623157e308;
thread_local std::string t_last_error;

int guard(const std::function<void()> &f) {
    try {
        f();
        return BINEST_OK;
    } catch (const Error &e) {
        t_last_error = e.what();
        return e.code;
    } catch (const std::bad_alloc &) {
        t_last_error = "host allocation failed";
        return BINEST_ERR_MEMORY;
    } catch (const std::exception &e) {
        t_last_error = e.what();
        return BINEST_ERR_FUNCTION;
    }
}

// device row layout from the caller's column blocks: row i = (inputs[i][0..n_in), outputs[i], 0-padding to ncol).
// The raw blocks are copied host->device as they are (one DMA each, straight from the caller's — possibly pinned —
// buffers) and interleaved here, instead of being repacked on the host into a pageable staging vector.
// n_classes > 0: outputs are class labels, anything but an integer in [0, n_classes) raises *bad.
// in_shift / out_shift: pivots subtracted from the (single) input column / the output (polynomial regression).
__global__ void pack_rows_kernel(const double *__restrict__ in, int n_in, const double *__restrict__ out, long long rows,
                                 int ncol, int n_classes, double in_shift, double out_shift,
                                 double *__restrict__ data, int *__restrict__ bad) {
    const long long total = rows * ncol;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long i = e / ncol;
        const int c = (int)(e - i * ncol);
        double v = 0.0;
        if (c < n_in) v = in[i * n_in + c] - in_shift;
        else if (c == n_in && out) {
            v = out[i];
            if (n_classes > 0 && !(v >= 0.0 && v < (double)n_classes && v == floor(v))) atomicOr(bad, 1);
            v -= out_shift;
        }
        data[e] = v;
    }
}

// max_i |x_if| per input column of the packed rows (softmax operator: bound on the logits, operators.cuh OpLogistic::Row)
__global__ void __launch_bounds__(256)
colmax_kernel(const double *__restrict__ data, long long rows, int ncol, int n_in, double *__restrict__ out /* [gridDim.x][6] */) {
    double mx[6] = {0, 0, 0, 0, 0, 0};
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += (long long)gridDim.x * blockDim.x)
        for (int f = 0; f < n_in && f < 6; ++f) mx[f] = fmax(mx[f], fabs(data[i * ncol + f]));
    __shared__ double sh[8][6];
    for (int f = 0; f < 6; ++f) {
        const double v = warp_max(mx[f]);
        if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5][f] = v;
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        double v = 0.0;
        for (int w = 0; w < 8; ++w) v = fmax(v, sh[w][threadIdx.x]);
        out[(size_t)blockIdx.x * 6 + threadIdx.x] = v;
    }
}

// Pivot selection for the polynomial operator (operators.cuh), plain fp64 block sums — the pivots only have to be
// near the data, not exact.  With u = x - xbar:  out[b][k] = Sum u^k (k = 0..2 deg), out[b][11 + k] = Sum y u^k
// (k = 0..deg): the normal equations of the least-squares polynomial in u.  Called first with xbar = 0, deg = 1
// (Sum x, Sum x^2 -> mean and spread of x), then with the chosen xbar.
__global__ void __launch_bounds__(256)
polyreg_pivot_kernel(const double *__restrict__ x, const double *__restrict__ y, long long rows, int deg, double xbar,
                     double *__restrict__ out /* [gridDim.x][17] */) {
    double s[17];
    for (int k = 0; k < 17; ++k) s[k] = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += (long long)gridDim.x * blockDim.x) {
        const double u = x[i] - xbar, yv = y[i];
        double up = 1.0;
        for (int k = 0; k <= 2 * deg; ++k) {
            s[k] += up;
            if (k <= deg) s[11 + k] += yv * up;
            up *= u;
        }
    }
    __shared__ double sh[8][17];
    for (int k = 0; k < 17; ++k) {
        const double v = warp_sum(s[k]);
        if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < 17) {
        double v = 0.0;
        for (int w = 0; w < 8; ++w) v += sh[w][threadIdx.x];
        out[(size_t)blockIdx.x * 17 + threadIdx.x] = v;
    }
}

// data moments of the polynomial-regression epilogue (operators.cuh), taken over the PIVOTED device rows
// (x', y') exactly as the row loop reads them: m[0] = Sum y', m[k] = Sum x'^k, k = 1..deg.
// Sum t enters Sum e^2 multiplied by 2 delta: it has to be good to ~1e-15 relative.  Every thread accumulates in
// double-double (TwoSum), a block combines its 256 pairs the same way and writes (hi, lo) per moment; the host
// adds the few hundred block results in long double.
__device__ __forceinline__ void dd_add(double &hi, double &lo, double v) {
    const double s = hi + v, bb = s - hi;
    lo += (hi - (s - bb)) + (v - bb);
    hi = s;
}
__global__ void __launch_bounds__(256)
polyreg_moments_kernel(const double *__restrict__ data /* rows x (x', y') */, long long rows, int deg,
                       double *__restrict__ out /* [gridDim.x][6][2] */) {
    double hi[6] = {0, 0, 0, 0, 0, 0}, lo[6] = {0, 0, 0, 0, 0, 0};
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += (long long)gridDim.x * blockDim.x) {
        const double2 r = reinterpret_cast<const double2 *>(data)[i];
        const double xv = r.x;
        dd_add(hi[0], lo[0], r.y);
        double xp = xv;
        for (int k = 1; k <= deg; ++k) { dd_add(hi[k], lo[k], xp); xp *= xv; }
    }
    __shared__ double sh[256][12];
    for (int k = 0; k < 6; ++k) { sh[threadIdx.x][2 * k] = hi[k]; sh[threadIdx.x][2 * k + 1] = lo[k]; }
    __syncthreads();
    if (threadIdx.x < 6) {
        const int k = threadIdx.x;
        double h = 0.0, l = 0.0;
        for (int t = 0; t < 256; ++t) { dd_add(h, l, sh[t][2 * k]); l += sh[t][2 * k + 1]; }
        out[((size_t)blockIdx.x * 6 + k) * 2] = h;
        out[((size_t)blockIdx.x * 6 + k) * 2 + 1] = l;
    }
}

static double norm_cdf(double z) { return 0.5 * std::erfc(-z * 0.70710678118654752440084436210485); }

static void fill_prior(PriorSpec &pr, int d, const int32_t *kind, const double *lo, const double *hi,
                       const double *p0, const double *p1) {
    pr.d = d;
    for (int j = 0; j < d; ++j) {
        pr.kind[j] = kind[j];
        pr.lo[j] = lo[j];
        pr.hi[j] = hi[j];
        pr.p0[j] = p0 ? p0[j] : 0.0;
        pr.p1[j] = p1 ? p1[j] : 1.0;
        BN_REQUIRE(lo[j] < hi[j], BINEST_ERR_DIMENSION, "parameter box needs lo < hi");
        switch (kind[j]) {
        case BINEST_PRIOR_UNIFORM:  // BS:37-39
            BN_REQUIRE(std::isfinite(lo[j]) && std::isfinite(hi[j]), BINEST_ERR_NUMERICAL,
                       "LocationParameter prior needs a finite box");
            pr.lognorm[j] = -std::log(hi[j] - lo[j]);
            break;
        case BINEST_PRIOR_SCALE:  // BS:42-48
            BN_REQUIRE(lo[j] > 0.0 && std::isfinite(hi[j]), BINEST_ERR_NUMERICAL,
                       "ScaleParameter prior needs 0 < lo < hi < inf");
            pr.lognorm[j] = -std::log(std::log(hi[j] / lo[j]));
            break;
        case BINEST_PRIOR_NORMAL_TRUNC: {  // BS:51-59
            BN_REQUIRE(pr.p1[j] > 0.0, BINEST_ERR_NUMERICAL, "truncated normal prior needs sd > 0");
            const double mass = norm_cdf((hi[j] - pr.p0[j]) / pr.p1[j]) - norm_cdf((lo[j] - pr.p0[j]) / pr.p1[j]);
            BN_REQUIRE(mass > 0.0, BINEST_ERR_NUMERICAL, "prior has no mass inside the parameter box");
            pr.lognorm[j] = -std::log(pr.p1[j]) - kHalfLog2Pi - std::log(mass);
            break;
        }
        default: throw Error(BINEST_ERR_TYPE, "unknown prior kind");
        }
    }
}

// row-major P x d host -> SoA [d][Ps] on the device (scratch of the problem)
void upload_theta(binest_problem &p, const double *theta, int64_t P, int Ps) {
    std::vector<double> soa((size_t)p.d * Ps, 1.0);
    for (int64_t i = 0; i < P; ++i)
        for (int j = 0; j < p.d; ++j) soa[(size_t)j * Ps + i] = theta[i * p.d + j];
    if (p.s_theta.n < soa.size()) p.s_theta.alloc(soa.size());
    if (p.s_out.n < (size_t)Ps) p.s_out.alloc(Ps);
    BN_CUDA(cudaMemcpyAsync(p.s_theta.p, soa.data(), soa.size() * sizeof(double), cudaMemcpyHostToDevice, p.stream));
    BN_CUDA(cudaStreamSynchronize(p.stream));  // soa is a stack-lifetime buffer
}

// Pivots of the polynomial operator (operators.cuh): xbar = mean of x unless the data are already centred
// (|mean| <= sd/2 -> 0, the rows stay as uploaded), piv = intercept of the least-squares polynomial of degree deg in
// x - xbar.  Two small reductions over the raw columns; the (deg+1) x (deg+1) normal equations are solved on the
// host in long double, in the scaled variable (x - xbar)/sd.  Any finite pivot gives the same likelihood
// algebraically — a poor one only costs digits — so nothing here needs to be exact.
static void choose_polyreg_pivots(binest_problem &p, const double *x_dev, const double *y_dev) {
    const int deg = (int)p.iparam[0], blocks = 2 * p.num_sms;
    DevBuf<double> part((size_t)blocks * 17);
    std::vector<double> h((size_t)blocks * 17);
    auto reduce = [&](int dg, double xbar, long double (&tot)[17]) {
        polyreg_pivot_kernel<<<blocks, 256, 0, p.stream>>>(x_dev, y_dev, p.rows, dg, xbar, part.p);
        BN_LAUNCH_CHECK();
        BN_CUDA(cudaMemcpyAsync(h.data(), part.p, h.size() * sizeof(double), cudaMemcpyDeviceToHost, p.stream));
        BN_CUDA(cudaStreamSynchronize(p.stream));
        for (int k = 0; k < 17; ++k) tot[k] = 0.0L;
        for (int b = 0; b < blocks; ++b)
            for (int k = 0; k < 17; ++k) tot[k] += h[(size_t)b * 17 + k];
    };
    long double t[17];
    const long double n = (long double)p.rows;
    reduce(1, 0.0, t);
    const long double mean = t[1] / n;
    const long double var = std::max<long double>(t[2] / n - mean * mean, 0.0L);
    double xbar = (std::fabs((double)mean) > 0.5 * std::sqrt((double)var)) ? (double)mean : 0.0;
    if (!std::isfinite(xbar)) xbar = 0.0;
    reduce(deg, xbar, t);
    double piv = (double)(t[11] / n);  // fallback: mean of y
    {
        const long double sd = sqrtl(std::max<long double>(t[2] / n, 0.0L));  // rms of u = x - xbar
        if (sd > 0 && std::isfinite((double)sd)) {
            long double A[6][7];
            const int m = deg + 1;
            for (int i = 0; i < m; ++i) {
                for (int j = 0; j < m; ++j) A[i][j] = t[i + j] / n / powl(sd, i + j);
                A[i][m] = t[11 + i] / n / powl(sd, i);
            }
            bool ok = true;
            for (int c = 0; c < m && ok; ++c) {  // Gaussian elimination, partial pivoting
                int piv_r = c;
                for (int r2 = c + 1; r2 < m; ++r2)
                    if (fabsl(A[r2][c]) > fabsl(A[piv_r][c])) piv_r = r2;
                if (!(fabsl(A[piv_r][c]) > 1e-14L)) { ok = false; break; }
                if (piv_r != c)
                    for (int j = 0; j <= m; ++j) std::swap(A[c][j], A[piv_r][j]);
                for (int r2 = 0; r2 < m; ++r2) {
                    if (r2 == c) continue;
                    const long double f = A[r2][c] / A[c][c];
                    for (int j = c; j <= m; ++j) A[r2][j] -= f * A[c][j];
                }
            }
            if (ok) {
                const double a0 = (double)(A[0][m] / A[0][0]);  // the intercept does not depend on the scaling of u
                if (std::isfinite(a0)) piv = a0;
            }
        }
    }
    if (!std::isfinite(piv)) piv = 0.0;
    p.cst.xbar = xbar;
    p.cst.piv = piv;
}
}  // namespace binest

using namespace binest;

extern "C" {

int binest_version(void) { return BINEST_VERSION; }
const char *binest_last_error(void) { return t_last_error.c_str(); }
int64_t binest_launch_count(void) { return g_launches.load(); }

int binest_device_count(int *count) {
    return guard([&] { BN_CUDA(cudaGetDeviceCount(count)); });
}

int binest_init(double logzero, int device) {
    return guard([&] {
        BN_REQUIRE(logzero < 0.0 && std::isfinite(logzero), BINEST_ERR_NUMERICAL, "logzero must be finite and negative");
        int n = 0;
        BN_CUDA(cudaGetDeviceCount(&n));
        BN_REQUIRE(n > 0, BINEST_ERR_CUDA, "no CUDA device: libbinest has no CPU fallback");
        if (device >= 0) BN_CUDA(cudaSetDevice(device));
        int dev = 0;
        BN_CUDA(cudaGetDevice(&dev));
        cudaDeviceProp prop;
        BN_CUDA(cudaGetDeviceProperties(&prop, dev));
        BN_REQUIRE(prop.major == 10, BINEST_ERR_CUDA,
                   std::string("libbinest is built for sm_100a only; found ") + prop.name);
        dev_pool_configure(dev);
        g_logzero = logzero;
    });
}

void binest_default_options(binest_options *o) {
    if (!o) return;
    o->pool_size = 100;   // BS:839
    o->batch_k = 1;
    o->mc_steps = 200;    // BS:844
    o->max_iter = 10000;  // BS:841
    o->min_iter = 100;    // BS:842
    o->term_frac = 0.01;  // BS:845
    o->acc_min = 0.0;     // BS:848
    o->acc_max = 1.0;
    o->seed = 1;
    o->first_run_id = 0;
    o->n_runs = 1;
    o->loglmax = std::nan("");  // Automatic, BS:847
}

int binest_measure_fp64_peak(double *tflops, double *ms_out) {
    return guard([&] {
        int dev = 0, sms = 0;
        BN_CUDA(cudaGetDevice(&dev));
        BN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        DevBuf<double> out(1);
        const int iters = 4096, blocks = sms * 8, threads = 256;
        cudaEvent_t e0, e1;
        BN_CUDA(cudaEventCreate(&e0));
        BN_CUDA(cudaEventCreate(&e1));
        float best = 1e30f;
        for (int rep = 0; rep < 5; ++rep) {
            BN_CUDA(cudaEventRecord(e0));
            fp64_peak_kernel<<<blocks, threads>>>(out.p, iters, 1.0 + rep);
            BN_LAUNCH_CHECK();
            BN_CUDA(cudaEventRecord(e1));
            BN_CUDA(cudaEventSynchronize(e1));
            float ms = 0;
            BN_CUDA(cudaEventElapsedTime(&ms, e0, e1));
            if (rep > 0 && ms < best) best = ms;
        }
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        const double flops = 2.0 * 64.0 * (double)iters * (double)blocks * threads;
        if (tflops) *tflops = flops / (best * 1e-3) / 1e12;
        if (ms_out) *ms_out = best;
    });
}

int binest_problem_create(int op_id, const int64_t *iparam, const double *inputs, int64_t n_rows, int64_t n_in,
                          const double *outputs, int64_t n_out, int64_t d, const int32_t *prior_kind,
                          const double *lo, const double *hi, const double *prior_p0, const double *prior_p1,
                          binest_problem **out) {
    return guard([&] {
        BN_REQUIRE(out, BINEST_ERR_TYPE, "null output handle");
        BN_REQUIRE(inputs && n_rows > 0 && n_in > 0, BINEST_ERR_DIMENSION, "empty data");
        BN_REQUIRE(d > 0 && d <= BINEST_MAXD, BINEST_ERR_DIMENSION, "1 <= d <= 16 parameters supported");
        BN_REQUIRE(prior_kind && lo && hi, BINEST_ERR_TYPE, "prior / parameter box missing");
        int ndev = 0;
        BN_CUDA(cudaGetDeviceCount(&ndev));
        BN_REQUIRE(ndev > 0, BINEST_ERR_CUDA, "no CUDA device: libbinest has no CPU fallback");
        std::unique_ptr<binest_problem> p(new binest_problem());
        p->op = op_id;
        for (int i = 0; i < 4; ++i) p->iparam[i] = iparam ? iparam[i] : 0;
        p->d = (int)d;
        BN_CUDA(cudaGetDevice(&p->device));
        BN_CUDA(cudaDeviceGetAttribute(&p->num_sms, cudaDevAttrMultiProcessorCount, p->device));
        BN_CUDA(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
        fill_prior(p->prior, (int)d, prior_kind, lo, hi, prior_p0, prior_p1);

        std::vector<double> host;  // only the GBM adaptor pre-processes on the host (long double increments)
        auto need_out = [&] { BN_REQUIRE(outputs && n_out == 1, BINEST_ERR_DIMENSION, "operator needs one output column"); };
        int n_classes = 0;
        bool pack = false;
        switch (op_id) {
        case BINEST_OP_GAUSSIAN_IID:
            BN_REQUIRE(n_in == 1 && d == 2, BINEST_ERR_DIMENSION, "Gaussian i.i.d.: 1 data column, theta = (mu, sigma)");
            p->rows = n_rows; p->ncol = 1;
            pack = true;
            break;
        case BINEST_OP_POLYREG: {
            need_out();
            const int64_t deg = p->iparam[0];
            BN_REQUIRE(n_in == 1 && deg >= 1 && deg <= 5 && d == deg + 2, BINEST_ERR_DIMENSION,
                       "polynomial regression: 1 input column, degree 1..5, theta = (c_0..c_deg, sigma)");
            p->rows = n_rows; p->ncol = 2;
            pack = true;
            break;
        }
        case BINEST_OP_LOGISTIC: {
            need_out();
            const int64_t K = p->iparam[1];
            p->iparam[2] = n_in;
            BN_REQUIRE(K >= 2 && d == (K - 1) * (n_in + 1), BINEST_ERR_DIMENSION,
                       "logistic: theta = (K-1) blocks of (w_1..w_F, b)");
            p->rows = n_rows; p->ncol = (int)((n_in + 2) & ~1LL);
            n_classes = (int)K;
            pack = true;
            break;
        }
        case BINEST_OP_GBM: {  // TemporalData adaptor BS:511-515: inputs = times, outputs = values
            need_out();
            BN_REQUIRE(n_in == 1 && d == 2 && n_rows >= 2, BINEST_ERR_DIMENSION,
                       "GBM: times column, values column, theta = (mu, sigma)");
            p->rows = n_rows - 1; p->ncol = 2;
            host.resize((size_t)p->rows * 2);
            long double cst = 0.0L;
            for (int64_t i = 1; i < n_rows; ++i) {
                const long double dt = (long double)inputs[i] - (long double)inputs[i - 1];
                BN_REQUIRE(dt > 0 && outputs[i] > 0 && outputs[i - 1] > 0, BINEST_ERR_NUMERICAL,
                           "GBM: times must increase and values must be positive");
                const long double r = logl((long double)outputs[i] / (long double)outputs[i - 1]);
                const long double sq = sqrtl(dt);
                host[2 * (i - 1)] = (double)(r / sq);
                host[2 * (i - 1) + 1] = (double)sq;
                cst += -logl((long double)outputs[i]) - 0.5L * logl(dt);
            }
            cst -= (long double)p->rows * (long double)kHalfLog2Pi;
            p->cst.c = (double)cst;
            break;
        }
        case BINEST_OP_GP_SE: {
            need_out();
            BN_REQUIRE(d == 3, BINEST_ERR_DIMENSION, "GP (SE kernel): theta = (sigma_f, ell, sigma_n)");
            p->gp_n = n_rows; p->gp_dim = n_in;
            p->gp_x.alloc((size_t)n_rows * n_in);
            p->gp_y.alloc((size_t)n_rows);
            BN_CUDA(cudaMemcpy(p->gp_x.p, inputs, sizeof(double) * n_rows * n_in, cudaMemcpyHostToDevice));
            BN_CUDA(cudaMemcpy(p->gp_y.p, outputs, sizeof(double) * n_rows, cudaMemcpyHostToDevice));
            p->rows = n_rows; p->ncol = (int)n_in;
            break;
        }
        default: throw Error(BINEST_ERR_FUNCTION, "unknown operator id");
        }
        if (pack || !host.empty()) {
            dispatch_op(*p, [](auto) {});  // reject shapes outside the fixed table before uploading
            const size_t cells = (size_t)p->rows * p->ncol;
            p->data.alloc(cells + 4);  // slack so 16-byte-rounded bulk copies stay in bounds
            BN_CUDA(cudaMemsetAsync(p->data.p + cells, 0, 4 * sizeof(double), p->stream));
            if (!host.empty()) {
                BN_CUDA(cudaMemcpyAsync(p->data.p, host.data(), cells * sizeof(double), cudaMemcpyHostToDevice, p->stream));
                BN_CUDA(cudaStreamSynchronize(p->stream));
            } else if (p->ncol == n_in && !outputs) {  // rows are already in device layout
                BN_CUDA(cudaMemcpyAsync(p->data.p, inputs, cells * sizeof(double), cudaMemcpyHostToDevice, p->stream));
                BN_CUDA(cudaStreamSynchronize(p->stream));
            } else {
                DevBuf<double> raw_in((size_t)n_rows * n_in), raw_out(outputs ? (size_t)n_rows : 0);
                DevBuf<int> bad(1);
                bad.zero(p->stream);
                BN_CUDA(cudaMemcpyAsync(raw_in.p, inputs, raw_in.n * sizeof(double), cudaMemcpyHostToDevice, p->stream));
                if (outputs)
                    BN_CUDA(cudaMemcpyAsync(raw_out.p, outputs, raw_out.n * sizeof(double), cudaMemcpyHostToDevice, p->stream));
                const unsigned grid = (unsigned)std::min<size_t>((cells + 255) / 256, (size_t)p->num_sms * 16);
                const bool poly = p->op == BINEST_OP_POLYREG;
                if (poly) choose_polyreg_pivots(*p, raw_in.p, raw_out.p);
                pack_rows_kernel<<<grid, 256, 0, p->stream>>>(raw_in.p, (int)n_in, raw_out.p, p->rows, p->ncol, n_classes,
                                                            p->cst.xbar, p->cst.piv, p->data.p, bad.p);
                BN_LAUNCH_CHECK();
                // softmax: column maxima of |x| -> OpCst::m[f] (NaN / Inf inputs give an infinite bound: always checked)
                const int cm_blocks = p->op == BINEST_OP_LOGISTIC ? 2 * p->num_sms : 0;
                DevBuf<double> cm((size_t)cm_blocks * 6);
                std::vector<double> h_cm((size_t)cm_blocks * 6);
                if (cm_blocks) {
                    colmax_kernel<<<cm_blocks, 256, 0, p->stream>>>(p->data.p, p->rows, p->ncol, (int)n_in, cm.p);
                    BN_LAUNCH_CHECK();
                    BN_CUDA(cudaMemcpyAsync(h_cm.data(), cm.p, h_cm.size() * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
                }
                const int mom_blocks = poly ? 2 * p->num_sms : 0;
                DevBuf<double> mom((size_t)mom_blocks * 12);
                std::vector<double> h_mom((size_t)mom_blocks * 12);
                if (mom_blocks) {
                    polyreg_moments_kernel<<<mom_blocks, 256, 0, p->stream>>>(p->data.p, p->rows, (int)p->iparam[0], mom.p);
                    BN_LAUNCH_CHECK();
                    BN_CUDA(cudaMemcpyAsync(h_mom.data(), mom.p, h_mom.size() * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
                }
                int h_bad = 0;
                BN_CUDA(cudaMemcpyAsync(&h_bad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, p->stream));
                BN_CUDA(cudaStreamSynchronize(p->stream));
                BN_REQUIRE(!h_bad, BINEST_ERR_NUMERICAL, "logistic: class labels must be integers 0..K-1");
                for (int f = 0; f < 6 && cm_blocks; ++f) {
                    double m = 0.0;
                    bool nan = false;
                    for (int b = 0; b < cm_blocks; ++b) {
                        const double v = h_cm[(size_t)b * 6 + f];
                        nan = nan || !(v == v);
                        m = std::max(m, v);
                    }
                    p->cst.m[f] = nan ? INFINITY : m;
                }
                for (int k = 0; k < 6 && mom_blocks; ++k) {
                    long double m = 0.0L;
                    for (int b = 0; b < mom_blocks; ++b)
                        m += (long double)h_mom[((size_t)b * 6 + k) * 2] + (long double)h_mom[((size_t)b * 6 + k) * 2 + 1];
                    p->cst.m[k] = (double)m;
                }
            }
        }
        *out = p.release();
    });
}

int binest_problem_free(binest_problem *p) {
    return guard([&] { delete p; });
}

int binest_problem_dim(const binest_problem *p, int64_t *d) {
    return guard([&] {
        BN_REQUIRE(p && d, BINEST_ERR_TYPE, "null argument");
        *d = p->d;
    });
}

int binest_loglike(binest_problem *p, const double *theta, int64_t P, double *out) {
    return guard([&] {
        BN_REQUIRE(p && theta && out, BINEST_ERR_TYPE, "null argument");
        if (P <= 0) return;  // empty list in, empty list out (Listable)
        BN_CUDA(cudaSetDevice(p->device));
        const int Ps = (int)((P + 31) & ~31LL);
        upload_theta(*p, theta, P, Ps);
        loglike_device(*p, p->s_theta.p, (int)P, Ps, p->s_out.p);
        BN_CUDA(cudaMemcpyAsync(out, p->s_out.p, sizeof(double) * P, cudaMemcpyDeviceToHost, p->stream));
        BN_CUDA(cudaStreamSynchronize(p->stream));
        if (p->comm) comm_check_abort(*p->comm, p->stream);
        if (p->comm_batch) comm_check_abort(*p->comm_batch, p->stream);
    });
}

int binest_gp_predict(binest_problem *p, const double *theta, int64_t M, const double *xstar, int64_t Q, double *mean,
                      double *sd) {
    return guard([&] {
        BN_REQUIRE(p && theta && xstar && mean && sd, BINEST_ERR_TYPE, "null argument");
        BN_REQUIRE(p->op == BINEST_OP_GP_SE, BINEST_ERR_FUNCTION, "binest_gp_predict needs a Gaussian-process problem");
        BN_REQUIRE(M < (1LL << 30) && Q < (1LL << 24), BINEST_ERR_DIMENSION, "too many samples or prediction points");
        if (M <= 0 || Q <= 0) return;
        BN_CUDA(cudaSetDevice(p->device));
        const int Ps = (int)((M + 31) & ~31LL);
        upload_theta(*p, theta, M, Ps);
        DevBuf<double> xs((size_t)Q * p->gp_dim), dm((size_t)M * Q), ds((size_t)M * Q);
        BN_CUDA(cudaMemcpyAsync(xs.p, xstar, xs.n * sizeof(double), cudaMemcpyHostToDevice, p->stream));
        gp_predict_device(*p, p->s_theta.p, (int)M, Ps, xs.p, (int)Q, dm.p, ds.p);
        BN_CUDA(cudaMemcpyAsync(mean, dm.p, dm.n * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
        BN_CUDA(cudaMemcpyAsync(sd, ds.p, ds.n * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
        BN_CUDA(cudaStreamSynchronize(p->stream));
    });
}

int binest_predictive_width(const binest_problem *p, int64_t *n_comp) {
    return guard([&] {
        BN_REQUIRE(p && n_comp, BINEST_ERR_TYPE, "null argument");
        *n_comp = p->op == BINEST_OP_POLYREG ? 2 : p->op == BINEST_OP_LOGISTIC ? p->iparam[1] : 0;
    });
}

int binest_predictive_components(binest_problem *p, const double *theta, int64_t M, const double *inputs, int64_t Q,
                                 double *out) {
    return guard([&] {
        BN_REQUIRE(p && theta && inputs && out, BINEST_ERR_TYPE, "null argument");
        BN_REQUIRE(p->op == BINEST_OP_POLYREG || p->op == BINEST_OP_LOGISTIC, BINEST_ERR_FUNCTION,
                   "predictive components need an operator with independent variables (polynomial regression, softmax)");
        if (M <= 0 || Q <= 0) return;
        const int K = (int)p->iparam[1], F = p->op == BINEST_OP_LOGISTIC ? (int)p->iparam[2] : 1;
        BN_REQUIRE(p->op != BINEST_OP_LOGISTIC || K <= kPredMaxD, BINEST_ERR_DIMENSION, "too many classes");
        const int64_t C = p->op == BINEST_OP_POLYREG ? 2 : K;
        BN_CUDA(cudaSetDevice(p->device));
        const int Ps = (int)((M + 31) & ~31LL);
        upload_theta(*p, theta, M, Ps);
        DevBuf<double> xs((size_t)Q * F), o((size_t)M * Q * C);
        BN_CUDA(cudaMemcpyAsync(xs.p, inputs, xs.n * sizeof(double), cudaMemcpyHostToDevice, p->stream));
        const long long total = (long long)M * Q;
        const unsigned grid = (unsigned)std::min<long long>((total + 255) / 256, (long long)p->num_sms * 8);
        predictive_kernel<<<grid, 256, 0, p->stream>>>(p->op, (int)p->iparam[0], K, F, p->s_theta.p, M, Ps, xs.p, Q, o.p);
        BN_LAUNCH_CHECK();
        BN_CUDA(cudaMemcpyAsync(out, o.p, o.n * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
        BN_CUDA(cudaStreamSynchronize(p->stream));
    });
}

int binest_logprior(binest_problem *p, const double *theta, int64_t P, double *out) {
    return guard([&] {
        BN_REQUIRE(p && theta && out, BINEST_ERR_TYPE, "null argument");
        if (P <= 0) return;
        BN_CUDA(cudaSetDevice(p->device));
        const int Ps = (int)((P + 31) & ~31LL);
        upload_theta(*p, theta, P, Ps);
        logprior_kernel<<<(unsigned)((P + 127) / 128), 128, 0, p->stream>>>(p->s_theta.p, (int)P, Ps, p->prior,
                                                                          g_logzero, p->s_out.p);
        BN_LAUNCH_CHECK();
        BN_CUDA(cudaMemcpyAsync(out, p->s_out.p, sizeof(double) * P, cudaMemcpyDeviceToHost, p->stream));
        BN_CUDA(cudaStreamSynchronize(p->stream));
    });
}

int binest_sample_prior(binest_problem *p, int64_t n, uint64_t seed, int64_t run_id, double *out) {
    return guard([&] {
        BN_REQUIRE(p && out, BINEST_ERR_TYPE, "null argument");
        if (n <= 0) return;
        BN_CUDA(cudaSetDevice(p->device));
        DevBuf<double> buf((size_t)n * p->d);
        sample_prior_kernel<<<(unsigned)((n + 127) / 128), 128, 0, p->stream>>>(p->prior, n, seed, (unsigned)run_id,
                                                                               buf.p);
        BN_LAUNCH_CHECK();
        BN_CUDA(cudaMemcpyAsync(out, buf.p, sizeof(double) * n * p->d, cudaMemcpyDeviceToHost, p->stream));
        BN_CUDA(cudaStreamSynchronize(p->stream));
    });
}

}  // extern "C"
