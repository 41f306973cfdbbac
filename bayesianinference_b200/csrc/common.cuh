// common.cuh — shared device/host helpers for libbinest (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include <atomic>
#include <cstdio>
#include <stdexcept>
#include <string>

#include "../../include/binest.h"

#define BINEST_MAXD 16

namespace binest {

constexpr double kLog2Pi = 1.8378770664093454835606594728112;
constexpr double kHalfLog2Pi = 0.91893853320467274178032973640562;
constexpr double kTwoPi = 6.283185307179586476925286766559;

// ------------------------------------------------------------------ host-side error plumbing
struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

#define BN_CUDA(expr)                                                                                    \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess)                                                                           \
            throw ::binest::Error(BINEST_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));  \
    } while (0)

#define BN_REQUIRE(cond, code, msg)                                 \
    do {                                                            \
        if (!(cond)) throw ::binest::Error((code), (msg));          \
    } while (0)

extern std::atomic<int64_t> g_launches;
extern double g_logzero;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }
#define BN_LAUNCH_CHECK()          \
    do {                           \
        ::binest::count_launch();  \
        BN_CUDA(cudaGetLastError()); \
    } while (0)

// Device memory comes from the device's stream-ordered pool (cudaMallocAsync) with a release threshold, so that
// the create/free churn of short-lived problems and runs (one per user call) recycles pool memory instead of
// mapping/unmapping it (cudaMalloc + cudaFree cost ~0.1-1 ms each; a run owns ~30 buffers).  Semantics stay those
// of cudaMalloc/cudaFree: the pointer is usable on every stream at return, and a free waits for the device.
inline void dev_pool_configure(int dev) {
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        uint64_t keep = 2ull << 30;  // freed memory retained for reuse (bytes); the rest returns to the driver
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
}
inline cudaError_t dev_alloc(void **p, size_t bytes) {
    cudaError_t e = cudaMallocAsync(p, bytes, (cudaStream_t)0);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize((cudaStream_t)0);
}
inline void dev_free(void *p) {
    cudaDeviceSynchronize();
    cudaFreeAsync(p, (cudaStream_t)0);
}

template <class T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    explicit DevBuf(size_t count) { alloc(count); }
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    DevBuf(DevBuf &&o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    DevBuf &operator=(DevBuf &&o) noexcept {
        if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
        return *this;
    }
    ~DevBuf() { release(); }
    void alloc(size_t count) {
        release();
        n = count;
        if (count) BN_CUDA(dev_alloc((void **)&p, count * sizeof(T)));
    }
    void release() {
        if (p) dev_free(p);
        p = nullptr;
        n = 0;
    }
    void zero(cudaStream_t s = 0) { if (n) BN_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s)); }
};

// ------------------------------------------------------------------ Philox4x32-10 (Random123)
// Same counter layout as oracle/binest_oracle.c: ctr = (c0, c1, c2, tag<<24 | run_id), key = seed.
enum : uint32_t { TAG_PRIOR = 1, TAG_START = 2, TAG_NORMAL = 3, TAG_ACCEPT = 4, TAG_EV_DEAD = 5, TAG_EV_LIVE = 6 };

__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                       uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
#ifdef __CUDA_ARCH__
        const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
#else
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t h0 = (uint32_t)(p0 >> 32), l0 = (uint32_t)p0, h1 = (uint32_t)(p1 >> 32), l1 = (uint32_t)p1;
#endif
        const uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__host__ __device__ __forceinline__ double u53(uint32_t hi, uint32_t lo) {
    const uint64_t m = ((uint64_t)(hi >> 5) << 26) | (uint64_t)(lo >> 6);
    return ((double)m + 0.5) * (1.0 / 9007199254740992.0);
}

__host__ __device__ __forceinline__ void rng_uniform2(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2,
                                                      uint32_t tag, uint32_t run_id, double &a, double &b) {
    uint32_t r[4];
    philox4x32_10(c0, c1, c2, (tag << 24) | (run_id & 0xFFFFFFu), (uint32_t)seed, (uint32_t)(seed >> 32), r);
    a = u53(r[0], r[1]);
    b = u53(r[2], r[3]);
}

__device__ __forceinline__ void rng_normal2(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t tag,
                                            uint32_t run_id, double &a, double &b) {
    double u0, u1;
    rng_uniform2(seed, c0, c1, c2, tag, run_id, u0, u1);
    const double r = sqrt(-2.0 * log(u0));
    double s, c;
    sincos(kTwoPi * u1, &s, &c);
    a = r * c;
    b = r * s;
}

// ------------------------------------------------------------------ log-space helpers (BU:318-356)
__device__ __forceinline__ double log_subtract(double logy, double logx) {  // BU:337-343
    return logy + log(1.0 - exp(logx - logy));
}
__device__ __forceinline__ double log_add(double logy, double logx) {  // BU:345-356
    const double mx = fmax(logx, logy), mn = fmin(logx, logy);
    return mx + log(1.0 + exp(mn - mx));
}

// running logsumexp state: value = m + log(s0); s1 carries Sum exp(t - m) * L for the entropy (BS:801-810)
struct LseAcc {
    double m, s0, s1;
};
__device__ __forceinline__ LseAcc lse_empty() { return LseAcc{-CUDART_INF, 0.0, 0.0}; }
__device__ __forceinline__ LseAcc lse_term(double t, double L) {
    if (!(t > -CUDART_INF) || !isfinite(t)) return lse_empty();  // Select[NumericQ] BU:333
    return LseAcc{t, 1.0, isfinite(L) ? L : 0.0};
}
__device__ __forceinline__ LseAcc lse_merge(const LseAcc &a, const LseAcc &b) {
    if (a.s0 == 0.0) return b;
    if (b.s0 == 0.0) return a;
    const double m = fmax(a.m, b.m);
    const double ea = exp(a.m - m), eb = exp(b.m - m);
    return LseAcc{m, a.s0 * ea + b.s0 * eb, a.s1 * ea + b.s1 * eb};
}
__device__ __forceinline__ LseAcc lse_shfl_down(const LseAcc &a, int off) {
    return LseAcc{__shfl_down_sync(0xffffffffu, a.m, off), __shfl_down_sync(0xffffffffu, a.s0, off),
                  __shfl_down_sync(0xffffffffu, a.s1, off)};
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
    return v;
}

// block-wide sum broadcast to all threads; scratch must hold 33 doubles; blockDim.x multiple of 32
__device__ __forceinline__ double block_sum(double v, double *scratch) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[w] = v;
    __syncthreads();
    if (w == 0) {
        double t = lane < nw ? scratch[lane] : 0.0;
        t = warp_sum(t);
        if (lane == 0) scratch[32] = t;
    }
    __syncthreads();
    return scratch[32];
}
// block-wide sums of nv <= 16 values at once, broadcast to all threads (v[k] <- total); scratch: 33 * 16 doubles.
// Per value the reduction tree is block_sum's (warp shuffles, then one warp over the warp sums): identical results.
__device__ __forceinline__ void block_sum_multi(double (&v)[16], int nv, double *scratch) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < 16; ++k)
        if (k < nv) v[k] = warp_sum(v[k]);
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 16; ++k)
            if (k < nv) scratch[k * 33 + w] = v[k];
    }
    __syncthreads();
    for (int k = w; k < nv; k += nw) {
        double t = lane < nw ? scratch[k * 33 + lane] : 0.0;
        t = warp_sum(t);
        if (lane == 0) scratch[k * 33 + 32] = t;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k)
        if (k < nv) v[k] = scratch[k * 33 + 32];
}
__device__ __forceinline__ double block_max(double v, double *scratch) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) scratch[w] = v;
    __syncthreads();
    if (w == 0) {
        double t = lane < nw ? scratch[lane] : -CUDART_INF;
        t = warp_max(t);
        if (lane == 0) scratch[32] = t;
    }
    __syncthreads();
    return scratch[32];
}
// block-wide LseAcc combine in two passes — block maximum of the partial maxima, ONE exp per thread to rescale its
// partial sums, then two block sums — instead of a tree of pairwise merges with two exps each (10 dependent merge
// levels: the evidence reductions were a third of the per-iteration update of a small run).  scratch: 33 doubles,
// scratch_m: 33 * 16 doubles.
__device__ __forceinline__ LseAcc block_lse2(const LseAcc &a, double *scratch, double *scratch_m) {
    const double m = block_max(a.s0 == 0.0 ? -CUDART_INF : a.m, scratch);
    if (!(m > -CUDART_INF)) return lse_empty();
    const double sc = a.s0 == 0.0 ? 0.0 : exp(a.m - m);
    double v[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] = 0.0;
    v[0] = a.s0 * sc;
    v[1] = a.s1 * sc;
    block_sum_multi(v, 2, scratch_m);
    return LseAcc{m, v[0], v[1]};
}

// block-wide LseAcc merge broadcast to all threads; scratch: 3*33 doubles
__device__ __forceinline__ LseAcc block_lse(LseAcc a, double *scratch) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a = lse_merge(a, lse_shfl_down(a, o));
    __syncthreads();
    if (lane == 0) { scratch[w] = a.m; scratch[33 + w] = a.s0; scratch[66 + w] = a.s1; }
    __syncthreads();
    if (w == 0) {
        LseAcc t = lane < nw ? LseAcc{scratch[lane], scratch[33 + lane], scratch[66 + lane]} : lse_empty();
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t = lse_merge(t, lse_shfl_down(t, o));
        if (lane == 0) { scratch[32] = t.m; scratch[65] = t.s0; scratch[98] = t.s1; }
    }
    __syncthreads();
    return LseAcc{scratch[32], scratch[65], scratch[98]};
}

// ------------------------------------------------------------------ mbarrier + 1-D bulk copy (TMA engine; SASS: UBLKCP)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// distributed shared memory: address of the same smem offset in CTA `rank` of the cluster, and a remote 8-byte store
// that signals the remote CTA's mbarrier with complete_tx (sm_90+)
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_async_f64(uint32_t remote_addr, double v, uint32_t remote_mbar) {
    asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(remote_addr),
                 "l"(__double_as_longlong(v)), "r"(remote_mbar)
                 : "memory");
}

// programmatic dependent launch (PDL): the next kernel of the walk graph may start its prologue while this one
// drains; it must call pdl_wait() before touching anything its predecessor wrote.  Both are no-ops for kernels
// launched without the programmatic-stream-serialization attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// global -> shared::cta bulk async copy, completion signalled on an mbarrier (bytes multiple of 16, 16B aligned)
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

}  // namespace binest
