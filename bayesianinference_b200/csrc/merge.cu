// merge.cu — combineRuns (BS:1293-1315) and the post-processing of the merged list (evidenceSampling BS:1158-1291) on
// the device.
//
// The reference joins the runs' sample lists, deletes duplicate points, re-sorts by {logL, point} and hands the list to
// evidenceSampling, which weights it as ONE run and finally sorts it by posterior weight (BS:1241).  For config C4 as
// stated (64 runs x 512 live points, 2.7e5 samples) that host work was a third of a parallelNestedSampling call on one
// GPU and most of it on eight.  Here it is a handful of kernels:
//   * the runs arrive sorted (the engine's fetch order), so the merged position of a sample is the number of samples
//     before it in the total order {logL, point, Join position}: one binary search per (sample, run) — no sort;
//   * duplicates (a walk without an accepted move returns a copy of a live point) are adjacent in that order, the first in
//     Join order in front: DeleteDuplicatesBy[Point] is a comparison with the predecessor and a stream compaction;
//   * the summed pool size at a sample's likelihood level is an exclusive prefix sum of the runs' pool-size steps;
//   * the final SortBy[-CrudePosteriorWeight] sorts chunks of 2048 keys in shared memory (bitonic) and merges the sorted
//     chunks with the same rank-by-binary-search kernel.
// Index work is exact, so the merged list equals the host merge of api._merge_samples sample for sample (tests).
#include <chrono>
#include <cstdlib>
#include <functional>
#include <vector>

#include "evidence.cuh"
#include "run_view.cuh"

namespace binest {
int guard(const std::function<void()> &f);
}

using namespace binest;

namespace {

constexpr int kSortChunk = 2048;  // keys per CTA of the shared-memory bitonic sort

// total order of SortBy[{#LogLikelihood, #Point}&] on the joined list; fully equal rows keep their Join order
struct CmpSample {
    const double *L;
    const double *pts;
    int d;
    __device__ __forceinline__ bool less(long long e, long long i) const {
        const double le = L[e], li = L[i];
        if (le < li) return true;
        if (le > li) return false;
        for (int a = 0; a < d; ++a) {
            const double pe = pts[e * d + a], pi = pts[i * d + a];
            if (pe < pi) return true;
            if (pe > pi) return false;
        }
        return e < i;
    }
};

// (key ascending, original index ascending): a stable sort
struct CmpKey {
    const unsigned long long *u;
    const unsigned *idx;
    __device__ __forceinline__ bool less(long long e, long long i) const {
        const unsigned long long ue = u[e], ui = u[i];
        return ue < ui || (ue == ui && idx[e] < idx[i]);
    }
};

// chunk (run) of every element: the last c with offs[c] <= i
__global__ void chunk_id_kernel(long long M, int R, const long long *__restrict__ offs, int *__restrict__ cid) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    int lo = 0, hi = R;  // invariant: offs[lo] <= i < offs[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (offs[mid] <= i) lo = mid; else hi = mid;
    }
    cid[i] = lo;
}

template <class C>
__global__ void check_sorted_kernel(long long M, const int *__restrict__ cid, C cmp, int *__restrict__ bad) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i + 1 >= M) return;
    if (cid[i] == cid[i + 1] && !cmp.less(i, i + 1)) *bad = 1;
}

// merged position = number of elements of all chunks that precede element i in the total order
template <class C>
__global__ void __launch_bounds__(256) rank_kernel(long long M, int R, const long long *__restrict__ offs, C cmp,
                                                    long long *__restrict__ rank) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    long long r = 0;
    for (int c = 0; c < R; ++c) {
        const long long o = offs[c];
        long long lo = o, hi = offs[c + 1];
        while (lo < hi) {
            const long long mid = (lo + hi) >> 1;
            if (cmp.less(mid, i)) lo = mid + 1; else hi = mid;
        }
        r += lo - o;
    }
    rank[i] = r;
}

__global__ void scatter_perm_kernel(long long M, const long long *__restrict__ rank, long long *__restrict__ perm) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < M) perm[rank[i]] = i;
}

// per merged position s: the pool-size step its sample contributes (its run's pool size after it minus at it; the last
// sample of a run takes the run's pool to zero), whether it opens a new likelihood level, whether it survives
// DeleteDuplicatesBy[Point]
__global__ void merge_flags_kernel(long long M, int d, const long long *__restrict__ perm, const int *__restrict__ cid,
                                   const long long *__restrict__ offs, const double *__restrict__ L,
                                   const double *__restrict__ pts, const long long *__restrict__ pool,
                                   long long *__restrict__ delta, unsigned char *__restrict__ newlev,
                                   unsigned char *__restrict__ keep) {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= M) return;
    const long long i = perm[s];
    const int c = cid[i];
    delta[s] = (i + 1 < offs[c + 1]) ? pool[i + 1] - pool[i] : -pool[i];
    bool nl = true, kp = true;
    if (s > 0) {
        const long long j = perm[s - 1];
        nl = L[i] != L[j];
        bool same = true;
        for (int a = 0; a < d; ++a) same = same && (pts[i * d + a] == pts[j * d + a]);
        kp = !same;
    }
    newlev[s] = nl ? 1 : 0;
    keep[s] = kp ? 1 : 0;
}

// single CTA: pool[s] = base + (sum of the steps of all samples STRICTLY below the level of s), outpos[s] = number of
// kept samples before s; *n_keep = kept samples in total
__global__ void __launch_bounds__(1024)
merge_scan_kernel(long long M, long long base, const long long *__restrict__ delta, const unsigned char *__restrict__ newlev,
                  const unsigned char *__restrict__ keep, long long *__restrict__ pool_s, long long *__restrict__ outpos,
                  long long *__restrict__ n_keep) {
    __shared__ long long s_sum[1024], s_keep[1024], s_lev[1024];
    __shared__ int s_has[1024];
    const int t = threadIdx.x, nt = blockDim.x;
    const long long per = (M + nt - 1) / nt;
    const long long a = (long long)t * per, b = (a + per < M) ? a + per : M;
    long long sum = 0, kp = 0;
    for (long long s = a; s < b; ++s) { sum += delta[s]; kp += keep[s]; }
    s_sum[t] = sum; s_keep[t] = kp;
    __syncthreads();
    if (t == 0) {  // exclusive prefixes over the 1024 segments
        long long x = 0, y = 0;
        for (int q = 0; q < nt; ++q) {
            const long long u = s_sum[q], v = s_keep[q];
            s_sum[q] = x; s_keep[q] = y;
            x += u; y += v;
        }
        *n_keep = y;
    }
    __syncthreads();
    // level value at the end of each segment: the exclusive prefix at the last level start inside it
    long long e = s_sum[t], lev = 0;
    int has = 0;
    for (long long s = a; s < b; ++s) {
        if (newlev[s]) { lev = e; has = 1; }
        e += delta[s];
    }
    s_lev[t] = lev; s_has[t] = has;
    __syncthreads();
    // carry-in: the level value of the nearest earlier segment that contains a level start (s = 0 always starts one)
    long long carry = 0;
    for (int q = t - 1; q >= 0; --q)
        if (s_has[q]) { carry = s_lev[q]; break; }
    e = s_sum[t];
    long long o = s_keep[t];
    lev = carry;
    for (long long s = a; s < b; ++s) {
        if (newlev[s]) lev = e;
        pool_s[s] = base + lev;
        outpos[s] = o;
        e += delta[s];
        o += keep[s];
    }
}

// stream compaction of the kept samples into the merged list; last_bad = the last merged index whose pool size differs
// from that of a final live set (M - k)
__global__ void merge_compact_kernel(long long M, int d, const long long *__restrict__ perm, const int *__restrict__ cid,
                                     const unsigned char *__restrict__ keep, const long long *__restrict__ outpos,
                                     const long long *__restrict__ pool_s, const long long *__restrict__ n_keep,
                                     const double *__restrict__ pts, const double *__restrict__ L,
                                     const double *__restrict__ lp, const double *__restrict__ acc,
                                     const long long *__restrict__ rid_in, double *__restrict__ o_pts,
                                     double *__restrict__ o_L, double *__restrict__ o_lp, double *__restrict__ o_acc,
                                     long long *__restrict__ o_pool, long long *__restrict__ o_rid,
                                     long long *__restrict__ last_bad) {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= M || !keep[s]) return;
    const long long i = perm[s], o = outpos[s];
    for (int a = 0; a < d; ++a) o_pts[o * d + a] = pts[i * d + a];
    o_L[o] = L[i];
    o_lp[o] = lp ? lp[i] : 0.0;
    o_acc[o] = acc ? acc[i] : 0.0;
    o_pool[o] = pool_s[s];
    o_rid[o] = rid_in ? rid_in[i] : (long long)cid[i];
    if (pool_s[s] != *n_keep - o) atomicMax(last_bad, o);
}

// the literal X sequence of the reference (BS:1307-1309 -> BS:785-799): n_tot for the deleted part, n_tot..1 for the tail
__global__ void reference_pool_kernel(long long M, long long n_tot, long long *__restrict__ pool) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < M) pool[k] = (k < M - n_tot) ? n_tot : M - k;
}

// CrudeLogPosteriorWeight - CrudeLogEvidence, its exponential, X = exp(logX) (BS:1236-1237); sort key of BS:1241
__global__ void normalise_weights_kernel(long long M, const double *__restrict__ summary, double *__restrict__ clw,
                                         double *__restrict__ cw, const double *__restrict__ logX, double *__restrict__ X) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= M) return;
    const double w = clw[k] - summary[0];
    clw[k] = w;
    cw[k] = exp(w);
    X[k] = exp(logX[k]);
}

// order-preserving map double -> uint64 (ascending); NaN sorts last, the padding after it
__device__ __forceinline__ unsigned long long ordered_bits(double x) {
    if (x != x) return 0xFFFFFFFFFFFFFFFEull;
    const long long b = __double_as_longlong(x + 0.0);  // -0.0 and 0.0 are one key
    return b < 0 ? ~(unsigned long long)b : ((unsigned long long)b | 0x8000000000000000ull);
}

// stable sort of chunks of kSortChunk keys, DESCENDING in w: bitonic network on (ordered_bits(-w), index) in shared memory
__global__ void __launch_bounds__(1024)
chunk_sort_desc_kernel(long long M, const double *__restrict__ w, unsigned long long *__restrict__ ukey,
                       unsigned *__restrict__ uidx) {
    __shared__ unsigned long long sk[kSortChunk];
    __shared__ unsigned si[kSortChunk];
    const long long base = (long long)blockIdx.x * kSortChunk;
    for (int e = threadIdx.x; e < kSortChunk; e += blockDim.x) {
        const long long g = base + e;
        sk[e] = g < M ? ordered_bits(-w[g]) : 0xFFFFFFFFFFFFFFFFull;
        si[e] = (unsigned)(g < M ? g : 0xFFFFFFFFu);
    }
    __syncthreads();
    for (int k = 2; k <= kSortChunk; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int e = threadIdx.x; e < kSortChunk; e += blockDim.x) {
                const int p = e ^ j;
                if (p > e) {
                    const bool up = (e & k) == 0;
                    const unsigned long long a = sk[e], b = sk[p];
                    const unsigned ia = si[e], ib = si[p];
                    const bool gt = a > b || (a == b && ia > ib);
                    if (gt == up) { sk[e] = b; sk[p] = a; si[e] = ib; si[p] = ia; }
                }
            }
            __syncthreads();
        }
    }
    for (int e = threadIdx.x; e < kSortChunk; e += blockDim.x) {
        const long long g = base + e;
        if (g < M) { ukey[g] = sk[e]; uidx[g] = si[e]; }
    }
}

__global__ void scatter_order_kernel(long long M, const long long *__restrict__ rank, const unsigned *__restrict__ uidx,
                                     long long *__restrict__ order) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < M) order[rank[j]] = (long long)uidx[j];
}

// out[c][k] = col_c[order[k]] for the double columns, likewise points and the two integer columns
struct GatherCols {
    const double *col[BINEST_NCOL_F];
};
__global__ void gather_table_kernel(long long M, long long stride, int d, const long long *__restrict__ order,
                                    GatherCols in, const double *__restrict__ pts, const long long *__restrict__ pool,
                                    const long long *__restrict__ rid, double *__restrict__ table,
                                    double *__restrict__ o_pts, long long *__restrict__ itable) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= M) return;
    const long long i = order ? order[k] : k;
#pragma unroll
    for (int c = 0; c < BINEST_NCOL_F; ++c) table[(size_t)c * stride + k] = in.col[c] ? in.col[c][i] : CUDART_NAN;
    for (int a = 0; a < d; ++a) o_pts[k * d + a] = pts[i * d + a];
    itable[k] = pool[i];
    itable[stride + k] = rid[i];
}

inline unsigned nblk(long long n, int t = 256) { return (unsigned)((n + t - 1) / t); }

// BINEST_MERGE_DEBUG=1: wall-clock phases of a call on stderr (synchronises the device at every mark)
struct PhaseClock {
    bool on = std::getenv("BINEST_MERGE_DEBUG") != nullptr;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    void mark(const char *what) {
        if (!on) return;
        cudaDeviceSynchronize();
        const auto t1 = std::chrono::steady_clock::now();
        std::fprintf(stderr, "  merge phase %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

// the runs' sample lists joined in run order, on the device
struct Joined {
    long long Mt = 0, base = 0;  // samples; sum of the runs' pool sizes at their first samples
    int R = 0, d = 0;
    bool has_lp = false, has_acc = false, has_rid = false;
    DevBuf<double> pts, L, lp, acc;
    DevBuf<long long> pool, rid, offs;
};

// device-resident merged list
struct Merged {
    long long M = 0;         // samples after DeleteDuplicatesBy
    long long last_bad = -1; // last index whose summed pool size is not that of a final live set
    DevBuf<double> pts, L, lp, acc;
    DevBuf<long long> pool, rid;
};

void join_from_host(int64_t R, const int64_t *sizes, int64_t d, const double *points, const double *logL,
                    const double *logPrior, const double *acc, const int64_t *pool, const int64_t *run_id, Joined &j) {
    BN_REQUIRE(R >= 1 && sizes && d >= 1 && d <= BINEST_MAXD && points && logL && pool, BINEST_ERR_DIMENSION, "bad run lists");
    std::vector<long long> offs((size_t)R + 1, 0);
    for (int64_t c = 0; c < R; ++c) {
        BN_REQUIRE(sizes[c] >= 0, BINEST_ERR_DIMENSION, "negative run size");
        offs[c + 1] = offs[c] + sizes[c];
        if (sizes[c] > 0) j.base += pool[offs[c]];
    }
    const long long Mt = offs[R];
    BN_REQUIRE(Mt >= 1 && Mt < (1LL << 31), BINEST_ERR_DIMENSION, "1 <= total samples < 2^31");
    j.Mt = Mt; j.R = (int)R; j.d = (int)d;
    j.has_lp = logPrior != nullptr; j.has_acc = acc != nullptr; j.has_rid = run_id != nullptr;
    j.pts.alloc((size_t)Mt * d); j.L.alloc(Mt); j.pool.alloc(Mt); j.offs.alloc(R + 1);
    if (j.has_lp) j.lp.alloc(Mt);
    if (j.has_acc) j.acc.alloc(Mt);
    if (j.has_rid) j.rid.alloc(Mt);
    BN_CUDA(cudaMemcpy(j.pts.p, points, sizeof(double) * Mt * d, cudaMemcpyHostToDevice));
    BN_CUDA(cudaMemcpy(j.L.p, logL, sizeof(double) * Mt, cudaMemcpyHostToDevice));
    if (logPrior) BN_CUDA(cudaMemcpy(j.lp.p, logPrior, sizeof(double) * Mt, cudaMemcpyHostToDevice));
    if (acc) BN_CUDA(cudaMemcpy(j.acc.p, acc, sizeof(double) * Mt, cudaMemcpyHostToDevice));
    BN_CUDA(cudaMemcpy(j.pool.p, pool, sizeof(int64_t) * Mt, cudaMemcpyHostToDevice));
    if (run_id) BN_CUDA(cudaMemcpy(j.rid.p, run_id, sizeof(int64_t) * Mt, cudaMemcpyHostToDevice));
    BN_CUDA(cudaMemcpy(j.offs.p, offs.data(), sizeof(long long) * (R + 1), cudaMemcpyHostToDevice));
}

// sample e of run c straight from the engine's device state: the dead list, then the live set in sorted order with pool
// sizes n..1 — exactly the list binest_run_fetch assembles on the host
__global__ void join_runs_kernel(RunViewDev v, long long Mt, const long long *__restrict__ offs,
                                 const long long *__restrict__ n_dead, double *__restrict__ pts, double *__restrict__ L,
                                 double *__restrict__ lp, double *__restrict__ acc, long long *__restrict__ pool,
                                 long long *__restrict__ rid) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Mt) return;
    int lo = 0, hi = v.R;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (offs[mid] <= i) lo = mid; else hi = mid;
    }
    const int c = lo, d = v.d;
    const long long e = i - offs[c], D = n_dead[c];
    if (e < D) {
        const size_t k = (size_t)c * v.cap + e;
        for (int a = 0; a < d; ++a) pts[i * d + a] = v.dead_theta[k * d + a];
        L[i] = v.dead_logL[k]; lp[i] = v.dead_logPr[k]; acc[i] = v.dead_acc[k]; pool[i] = v.dead_pool[k];
    } else {
        const long long jj = e - D;
        const size_t k = (size_t)c * v.n + v.order[(size_t)c * v.n + jj];
        for (int a = 0; a < d; ++a) pts[i * d + a] = v.live_theta[k * d + a];
        L[i] = v.live_logL[k]; lp[i] = v.live_logPr[k]; acc[i] = v.live_acc[k]; pool[i] = v.n - jj;
    }
    rid[i] = v.first_run_id + c;
}

// packed device tables of the *_dev entry points: row k = (point[d], logL, logPrior, acc, pool, run id), all as doubles
__global__ void pack_merged_kernel(long long M, int d, const double *__restrict__ pts, const double *__restrict__ L,
                                   const double *__restrict__ lp, const double *__restrict__ acc,
                                   const long long *__restrict__ pool, const long long *__restrict__ rid,
                                   double *__restrict__ out) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= M) return;
    double *row = out + (size_t)k * (d + 5);
    for (int a = 0; a < d; ++a) row[a] = pts[k * d + a];
    row[d] = L[k]; row[d + 1] = lp[k]; row[d + 2] = acc[k]; row[d + 3] = (double)pool[k]; row[d + 4] = (double)rid[k];
}
__global__ void unpack_joined_kernel(long long M, int d, const double *__restrict__ in, double *__restrict__ pts,
                                     double *__restrict__ L, double *__restrict__ lp, double *__restrict__ acc,
                                     long long *__restrict__ pool, long long *__restrict__ rid) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= M) return;
    const double *row = in + (size_t)k * (d + 5);
    for (int a = 0; a < d; ++a) pts[k * d + a] = row[a];
    L[k] = row[d]; lp[k] = row[d + 1]; acc[k] = row[d + 2]; pool[k] = (long long)row[d + 3]; rid[k] = (long long)row[d + 4];
}

void join_from_dev(int64_t R, const int64_t *sizes, int64_t d, const double *table_dev, Joined &j) {
    BN_REQUIRE(R >= 1 && sizes && d >= 1 && d <= BINEST_MAXD && table_dev, BINEST_ERR_DIMENSION, "bad run lists");
    std::vector<long long> offs((size_t)R + 1, 0);
    for (int64_t c = 0; c < R; ++c) {
        BN_REQUIRE(sizes[c] >= 0, BINEST_ERR_DIMENSION, "negative run size");
        offs[c + 1] = offs[c] + sizes[c];
    }
    const long long Mt = offs[R];
    BN_REQUIRE(Mt >= 1 && Mt < (1LL << 31), BINEST_ERR_DIMENSION, "1 <= total samples < 2^31");
    for (int64_t c = 0; c < R; ++c)
        if (sizes[c] > 0) {  // pool size at the first sample of every list
            double v = 0.0;
            BN_CUDA(cudaMemcpy(&v, table_dev + (size_t)offs[c] * (d + 5) + d + 3, sizeof(double), cudaMemcpyDeviceToHost));
            j.base += (long long)v;
        }
    j.Mt = Mt; j.R = (int)R; j.d = (int)d;
    j.has_lp = j.has_acc = j.has_rid = true;
    j.pts.alloc((size_t)Mt * d); j.L.alloc(Mt); j.lp.alloc(Mt); j.acc.alloc(Mt); j.pool.alloc(Mt); j.rid.alloc(Mt);
    j.offs.alloc(R + 1);
    BN_CUDA(cudaMemcpy(j.offs.p, offs.data(), sizeof(long long) * (R + 1), cudaMemcpyHostToDevice));
    unpack_joined_kernel<<<nblk(Mt), 256>>>(Mt, (int)d, table_dev, j.pts.p, j.L.p, j.lp.p, j.acc.p, j.pool.p, j.rid.p);
    BN_LAUNCH_CHECK();
}

void join_from_run(binest_run *r, Joined &j) {
    RunView v;
    run_view(r, v);  // flushes the batch in flight, reads the run states
    BN_CUDA(cudaSetDevice(v.device));
    std::vector<long long> offs((size_t)v.dev.R + 1, 0);
    for (int c = 0; c < v.dev.R; ++c) offs[c + 1] = offs[c] + v.n_dead[c] + v.dev.n;
    const long long Mt = offs[v.dev.R];
    BN_REQUIRE(Mt >= 1 && Mt < (1LL << 31), BINEST_ERR_DIMENSION, "1 <= total samples < 2^31");
    j.Mt = Mt; j.R = v.dev.R; j.d = v.dev.d;
    j.base = (long long)v.dev.R * v.dev.n;  // every run starts with a full pool
    j.has_lp = j.has_acc = j.has_rid = true;
    j.pts.alloc((size_t)Mt * j.d); j.L.alloc(Mt); j.lp.alloc(Mt); j.acc.alloc(Mt); j.pool.alloc(Mt); j.rid.alloc(Mt);
    j.offs.alloc(j.R + 1);
    DevBuf<long long> nd(j.R);
    BN_CUDA(cudaMemcpy(j.offs.p, offs.data(), sizeof(long long) * (j.R + 1), cudaMemcpyHostToDevice));
    BN_CUDA(cudaMemcpy(nd.p, v.n_dead.data(), sizeof(long long) * j.R, cudaMemcpyHostToDevice));
    join_runs_kernel<<<nblk(Mt), 256>>>(v.dev, Mt, j.offs.p, nd.p, j.pts.p, j.L.p, j.lp.p, j.acc.p, j.pool.p, j.rid.p);
    BN_LAUNCH_CHECK();
    BN_CUDA(cudaDeviceSynchronize());  // nd is released here
}

void device_merge(Joined &j, Merged &m) {
    const long long Mt = j.Mt;
    const int R = j.R, d = j.d;
    PhaseClock pc;
    DevBuf<long long> dRank(Mt), dPerm(Mt), dDelta(Mt), dPoolS(Mt), dOut(Mt), dCnt(2);
    DevBuf<int> dCid(Mt), dBad(1);
    DevBuf<unsigned char> dNew(Mt), dKeep(Mt);
    BN_CUDA(cudaMemset(dBad.p, 0, sizeof(int)));
    const long long init_cnt[2] = {0, -1};
    BN_CUDA(cudaMemcpy(dCnt.p, init_cnt, sizeof(init_cnt), cudaMemcpyHostToDevice));
    pc.mark("alloc");
    const CmpSample cmp{j.L.p, j.pts.p, d};
    chunk_id_kernel<<<nblk(Mt), 256>>>(Mt, R, j.offs.p, dCid.p);
    BN_LAUNCH_CHECK();
    check_sorted_kernel<<<nblk(Mt), 256>>>(Mt, dCid.p, cmp, dBad.p);
    BN_LAUNCH_CHECK();
    int bad = 0;  // an unsorted run would make the ranks below collide (perm would not be a permutation): stop here
    BN_CUDA(cudaMemcpy(&bad, dBad.p, sizeof(int), cudaMemcpyDeviceToHost));
    BN_REQUIRE(!bad, BINEST_ERR_DIMENSION, "every run's samples must be sorted by {logL, point} (the order binest_run_fetch returns)");
    pc.mark("chunk ids + sorted check");
    rank_kernel<<<nblk(Mt), 256>>>(Mt, R, j.offs.p, cmp, dRank.p);
    BN_LAUNCH_CHECK();
    pc.mark("rank");
    scatter_perm_kernel<<<nblk(Mt), 256>>>(Mt, dRank.p, dPerm.p);
    BN_LAUNCH_CHECK();
    merge_flags_kernel<<<nblk(Mt), 256>>>(Mt, d, dPerm.p, dCid.p, j.offs.p, j.L.p, j.pts.p, j.pool.p, dDelta.p, dNew.p, dKeep.p);
    BN_LAUNCH_CHECK();
    pc.mark("scatter + flags");
    merge_scan_kernel<<<1, 1024>>>(Mt, j.base, dDelta.p, dNew.p, dKeep.p, dPoolS.p, dOut.p, dCnt.p);
    BN_LAUNCH_CHECK();
    pc.mark("scan");
    m.pts.alloc((size_t)Mt * d); m.L.alloc(Mt); m.lp.alloc(Mt); m.acc.alloc(Mt); m.pool.alloc(Mt); m.rid.alloc(Mt);
    merge_compact_kernel<<<nblk(Mt), 256>>>(Mt, d, dPerm.p, dCid.p, dKeep.p, dOut.p, dPoolS.p, dCnt.p, j.pts.p, j.L.p,
                                            j.has_lp ? j.lp.p : nullptr, j.has_acc ? j.acc.p : nullptr,
                                            j.has_rid ? j.rid.p : nullptr, m.pts.p, m.L.p, m.lp.p, m.acc.p, m.pool.p,
                                            m.rid.p, dCnt.p + 1);
    BN_LAUNCH_CHECK();
    long long cnt[2];
    BN_CUDA(cudaMemcpy(cnt, dCnt.p, sizeof(cnt), cudaMemcpyDeviceToHost));
    m.M = cnt[0];
    m.last_bad = cnt[1];
    pc.mark("compact + counts");
}

// stable descending argsort of w[0..M) on the device: order[k] = index of the k-th largest
void device_argsort_desc(long long M, const double *w_dev, long long *order_dev) {
    const int nch = (int)((M + kSortChunk - 1) / kSortChunk);
    DevBuf<unsigned long long> uk(M);
    DevBuf<unsigned> ui(M);
    DevBuf<long long> offs(nch + 1), rank(M);
    std::vector<long long> h((size_t)nch + 1);
    for (int c = 0; c <= nch; ++c) h[c] = std::min<long long>((long long)c * kSortChunk, M);
    BN_CUDA(cudaMemcpy(offs.p, h.data(), sizeof(long long) * h.size(), cudaMemcpyHostToDevice));
    chunk_sort_desc_kernel<<<nch, 1024>>>(M, w_dev, uk.p, ui.p);
    BN_LAUNCH_CHECK();
    rank_kernel<<<nblk(M), 256>>>(M, nch, offs.p, CmpKey{uk.p, ui.p}, rank.p);
    BN_LAUNCH_CHECK();
    scatter_order_kernel<<<nblk(M), 256>>>(M, rank.p, ui.p, order_dev);
    BN_LAUNCH_CHECK();
}

void merged_to_host(const Joined &j, const Merged &m, double *points_out, double *logL_out, double *logPrior_out,
                    double *acc_out, int64_t *pool_out, int64_t *run_out, int64_t *M_out, int64_t *live_block) {
    auto back = [&](void *dst, const void *src, size_t bytes) {
        if (dst && bytes) BN_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
    };
    back(points_out, m.pts.p, sizeof(double) * m.M * j.d);
    back(logL_out, m.L.p, sizeof(double) * m.M);
    if (j.has_lp) back(logPrior_out, m.lp.p, sizeof(double) * m.M);
    if (j.has_acc) back(acc_out, m.acc.p, sizeof(double) * m.M);
    back(pool_out, m.pool.p, sizeof(int64_t) * m.M);
    back(run_out, m.rid.p, sizeof(int64_t) * m.M);
    if (M_out) *M_out = m.M;
    if (live_block) *live_block = m.M - (m.last_bad + 1);
}

// X sequence, crude weights, evidenceSampling, sort by weight, table to the host (stride Mt = the caller's allocation)
void device_post(const Joined &j, Merged &m, int32_t scheme, int64_t n_tot, int64_t post_runs, uint64_t seed,
                 double *points_out, double *table_out, int64_t *itable_out, double *z, double *pmean, double *H,
                 double *summary, int64_t *M_out, int64_t *n_live_out) {
    BN_REQUIRE(scheme == BINEST_MERGE_REFERENCE || scheme == BINEST_MERGE_POOLSIZES, BINEST_ERR_TYPE, "unknown merge scheme");
    BN_REQUIRE(post_runs >= 2 && post_runs <= 65535, BINEST_ERR_DIMENSION, "2 <= PostProcessSamplingRuns <= 65535");
    BN_REQUIRE(points_out && table_out && itable_out, BINEST_ERR_DIMENSION, "missing output tables");
    PhaseClock pc;
    const long long M = m.M, Mt = j.Mt;
    const int d = j.d;
    BN_REQUIRE(M >= 2, BINEST_ERR_DIMENSION, "bad sample list");
    long long n_live;
    if (scheme == BINEST_MERGE_REFERENCE) {
        BN_REQUIRE(n_tot >= 1 && n_tot <= M, BINEST_ERR_DIMENSION, "Total[SamplePoolSize] exceeds the merged list");
        reference_pool_kernel<<<nblk(M), 256>>>(M, n_tot, m.pool.p);
        BN_LAUNCH_CHECK();
        n_live = n_tot;
    } else {
        n_live = std::max<long long>(M - (m.last_bad + 1), 1);
    }
    const int Rr = (int)post_runs;
    static thread_local DevBuf<double> dSlx, dLw;  // R x M work arrays kept between calls (see binest_evidence_sampling)
    if (dSlx.n < (size_t)Rr * M) { dSlx.alloc((size_t)Rr * M); dLw.alloc((size_t)Rr * M); }
    DevBuf<double> logX(M), clw(M), cw(M), X(M), dS(4), dZ(Rr), dPm((size_t)Rr * d), dH(Rr), o1(M), o2(M), o3(M), o4(M);
    pc.mark("alloc work arrays");
    crude_weights_kernel<<<1, 1024>>>(M, n_live, m.L.p, nullptr, m.pool.p, logX.p, clw.p, dS.p);
    BN_LAUNCH_CHECK();
    pc.mark("crude weights");
    evidence_sampling_kernel<<<Rr, 1024>>>(M, d, n_live, m.pts.p, m.L.p, m.pool.p, seed, dSlx.p, dLw.p, dZ.p, dPm.p, dH.p);
    BN_LAUNCH_CHECK();
    evidence_moments_kernel<<<nblk(M), 256>>>(M, Rr, dSlx.p, dLw.p, dZ.p, o1.p, o2.p, o3.p, o4.p);
    BN_LAUNCH_CHECK();
    pc.mark("evidence sampling");
    normalise_weights_kernel<<<nblk(M), 256>>>(M, dS.p, clw.p, cw.p, logX.p, X.p);
    BN_LAUNCH_CHECK();
    DevBuf<long long> order(M), itab((size_t)2 * Mt);
    device_argsort_desc(M, clw.p, order.p);
    pc.mark("sort by weight");
    DevBuf<double> tab((size_t)BINEST_NCOL_F * Mt), opts((size_t)M * d);
    GatherCols gc{};
    gc.col[BINEST_COL_LOGL] = m.L.p;
    gc.col[BINEST_COL_LOGPRIOR] = j.has_lp ? m.lp.p : nullptr;
    gc.col[BINEST_COL_ACC] = j.has_acc ? m.acc.p : nullptr;
    gc.col[BINEST_COL_LOGX] = logX.p;
    gc.col[BINEST_COL_X] = X.p;
    gc.col[BINEST_COL_CRUDE_LOGW] = clw.p;
    gc.col[BINEST_COL_CRUDE_W] = cw.p;
    gc.col[BINEST_COL_SLX_MEAN] = o3.p;
    gc.col[BINEST_COL_SLX_SD] = o4.p;
    gc.col[BINEST_COL_LOGW_MEAN] = o1.p;
    gc.col[BINEST_COL_LOGW_SD] = o2.p;
    gather_table_kernel<<<nblk(M), 256>>>(M, Mt, d, order.p, gc, m.pts.p, m.pool.p, m.rid.p, tab.p, opts.p, itab.p);
    BN_LAUNCH_CHECK();
    BN_CUDA(cudaMemcpy(points_out, opts.p, sizeof(double) * M * d, cudaMemcpyDeviceToHost));
    for (int c = 0; c < BINEST_NCOL_F; ++c)  // only the M filled entries of every column travel
        BN_CUDA(cudaMemcpy(table_out + (size_t)c * Mt, tab.p + (size_t)c * Mt, sizeof(double) * M, cudaMemcpyDeviceToHost));
    for (int c = 0; c < 2; ++c)
        BN_CUDA(cudaMemcpy(itable_out + (size_t)c * Mt, itab.p + (size_t)c * Mt, sizeof(int64_t) * M, cudaMemcpyDeviceToHost));
    auto back = [&](double *dst, const double *src, size_t cnt) {
        if (dst) BN_CUDA(cudaMemcpy(dst, src, sizeof(double) * cnt, cudaMemcpyDeviceToHost));
    };
    back(z, dZ.p, Rr); back(pmean, dPm.p, (size_t)Rr * d); back(H, dH.p, Rr); back(summary, dS.p, 4);
    pc.mark("gather + D2H");
    if (M_out) *M_out = M;
    if (n_live_out) *n_live_out = n_live;
}

}  // namespace

int binest_merge_runs(int64_t R, const int64_t *sizes, int64_t d, const double *points, const double *logL,
                      const double *logPrior, const double *acc, const int64_t *pool, const int64_t *run_id,
                      double *points_out, double *logL_out, double *logPrior_out, double *acc_out, int64_t *pool_out,
                      int64_t *run_out, int64_t *M_out, int64_t *live_block) {
    return guard([&] {
        Joined j;
        Merged m;
        join_from_host(R, sizes, d, points, logL, logPrior, acc, pool, run_id, j);
        device_merge(j, m);
        merged_to_host(j, m, points_out, logL_out, logPrior_out, acc_out, pool_out, run_out, M_out, live_block);
    });
}

int binest_combine_runs(int64_t R, const int64_t *sizes, int64_t d, const double *points, const double *logL,
                        const double *logPrior, const double *acc, const int64_t *pool, const int64_t *run_id,
                        int32_t scheme, int64_t n_tot, int64_t post_runs, uint64_t seed, double *points_out,
                        double *table_out, int64_t *itable_out, double *z, double *pmean, double *H, double *summary,
                        int64_t *M_out, int64_t *n_live_out) {
    return guard([&] {
        Joined j;
        Merged m;
        join_from_host(R, sizes, d, points, logL, logPrior, acc, pool, run_id, j);
        device_merge(j, m);
        device_post(j, m, scheme, n_tot, post_runs, seed, points_out, table_out, itable_out, z, pmean, H, summary, M_out,
                    n_live_out);
    });
}

int binest_run_merge_size(binest_run *r, int64_t *M_total) {
    return guard([&] {
        BN_REQUIRE(r && M_total, BINEST_ERR_TYPE, "null argument");
        RunView v;
        run_view(r, v);
        long long Mt = 0;
        for (int c = 0; c < v.dev.R; ++c) Mt += v.n_dead[c] + v.dev.n;
        *M_total = Mt;
    });
}

int binest_run_merge(binest_run *r, double *points_out, double *logL_out, double *logPrior_out, double *acc_out,
                     int64_t *pool_out, int64_t *run_out, int64_t *M_out, int64_t *live_block) {
    return guard([&] {
        BN_REQUIRE(r, BINEST_ERR_TYPE, "null run");
        Joined j;
        Merged m;
        join_from_run(r, j);
        device_merge(j, m);
        merged_to_host(j, m, points_out, logL_out, logPrior_out, acc_out, pool_out, run_out, M_out, live_block);
    });
}

int binest_run_combine(binest_run *r, int32_t scheme, int64_t post_runs, uint64_t seed, double *points_out,
                       double *table_out, int64_t *itable_out, double *z, double *pmean, double *H, double *summary,
                       int64_t *M_out, int64_t *n_live_out) {
    return guard([&] {
        BN_REQUIRE(r, BINEST_ERR_TYPE, "null run");
        Joined j;
        Merged m;
        join_from_run(r, j);
        device_merge(j, m);
        device_post(j, m, scheme, j.base, post_runs, seed, points_out, table_out, itable_out, z, pmean, H, summary, M_out,
                    n_live_out);
    });
}

int binest_run_merge_dev(binest_run *r, double *table_dev, int64_t *M_out, int64_t *live_block) {
    return guard([&] {
        BN_REQUIRE(r && table_dev, BINEST_ERR_TYPE, "null argument");
        Joined j;
        Merged m;
        join_from_run(r, j);
        device_merge(j, m);
        pack_merged_kernel<<<nblk(m.M), 256>>>(m.M, j.d, m.pts.p, m.L.p, m.lp.p, m.acc.p, m.pool.p, m.rid.p, table_dev);
        BN_LAUNCH_CHECK();
        BN_CUDA(cudaDeviceSynchronize());
        if (M_out) *M_out = m.M;
        if (live_block) *live_block = m.M - (m.last_bad + 1);
    });
}

int binest_combine_runs_dev(int64_t R, const int64_t *sizes, int64_t d, const double *table_dev, int32_t scheme,
                            int64_t n_tot, int64_t post_runs, uint64_t seed, double *points_out, double *table_out,
                            int64_t *itable_out, double *z, double *pmean, double *H, double *summary, int64_t *M_out,
                            int64_t *n_live_out) {
    return guard([&] {
        Joined j;
        Merged m;
        join_from_dev(R, sizes, d, table_dev, j);
        device_merge(j, m);
        device_post(j, m, scheme, n_tot, post_runs, seed, points_out, table_out, itable_out, z, pmean, H, summary, M_out,
                    n_live_out);
    });
}
