// gp.cu — GP marginal-likelihood operator (squared-exponential kernel + nugget), batched over parameter vectors.
//
// Replaces, for theta = (sigma_f, ell, sigma_n):
//   covarianceMatrix / compiledCovarianceMatrix  GP:27-61   K_ij = sf^2 exp(-|x_i-x_j|^2/(2 l^2)) + delta_ij sn^2
//   matrixInverseAndDet                          GP:130-141 (LU in the reference; K is SPD, so Cholesky here)
//   gaussianProcessLogLikelihood[]               GP:181-199 -1/2 (N log 2pi + logdet + r.K^-1 r), clipped to
//                                                           +-|logzero|; factorisation failure -> logzero
// Layout: B matrices of order Np = roundup(N, 128), column-major, lower triangle, one after the other in HBM
// (134 MB each at N = 4096; 256 of them = 34 GB of the 180 GB).  Padding rows/columns carry the identity.
//
// Blocked right-looking Cholesky, panel width NB = 128, all B matrices advanced together:
//   gp_fill_kernel    one 128x128 tile per CTA, elementwise exp                                (fp64 pipe)
//   gp_potf2_kernel   diagonal block in shared memory: unblocked Cholesky, L11^-1, z_k = L11^-1 y_k,
//                     logdet and quadratic-form accumulation (the forward solve is fused into the sweep)
//   gp_trsm_kernel    L21 = A21 L11^-T as a 128x128x128 GEMM per row tile (DMMA), y_rest -= L21 z_k
//   gp_syrk_kernel    trailing update C -= L21 L21^T: the dense contraction, FP64 tensor-core MMA
//                     (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4), 128x128 tile per CTA, K streamed in chunks of
//                     32 through shared memory with cp.async double buffering
#include <algorithm>
#include <vector>

#include "problem.cuh"

namespace binest {
namespace {

constexpr int NB = 128;        // panel width / tile edge
constexpr int KC = 32;         // K chunk of the GEMM kernels
constexpr int LDS_ = NB + 4;   // smem leading dimension: half-warp fragment loads hit 16 distinct 8-byte banks

struct GpBatch {
    double *A;        // [B][Np*Np] column-major
    double *y;        // [B][Np] working right-hand side (consumed by the fused forward solve)
    double *z;        // [B][NB]  z_k of the current panel
    double *linvT;    // [B][NB*NB] (k, n) -> L11^-1[n][k]
    double *logdet;   // [B]
    double *quad;     // [B]
    int *fail;        // [B]
    int Np, N, B;
};

__device__ __forceinline__ void dmma_8x8x4(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N_>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N_) : "memory"); }

// ---------------------------------------------------------------------------------------------------------
// covariance fill: tile (ti, tj), ti >= tj, of matrix b.  theta SoA [3][Ps].
__global__ void __launch_bounds__(256)
gp_fill_kernel(GpBatch g, const double *__restrict__ x, int dim, const double *__restrict__ yin,
               const double *__restrict__ theta, int Ps, int b0) {
    const int b = blockIdx.y;
    // decode the lower-triangular tile index
    int t = blockIdx.x, ti = 0;
    while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
    const int tj = t - ti * (ti + 1) / 2;
    const double sf = theta[0 * (size_t)Ps + b0 + b], ell = theta[1 * (size_t)Ps + b0 + b], sn = theta[2 * (size_t)Ps + b0 + b];
    const double sf2 = sf * sf, il2 = 1.0 / (2.0 * ell * ell), sn2 = sn * sn;
    double *A = g.A + (size_t)b * g.Np * g.Np;
    extern __shared__ double sx[];  // [2][NB][dim]
    double *xi = sx, *xj = sx + NB * dim;
    for (int k = threadIdx.x; k < NB * dim; k += blockDim.x) {
        const int r = k / dim, c = k - r * dim;
        const int gi = ti * NB + r, gj = tj * NB + r;
        xi[k] = gi < g.N ? x[(size_t)gi * dim + c] : 0.0;
        xj[k] = gj < g.N ? x[(size_t)gj * dim + c] : 0.0;
    }
    if (tj == 0 && ti * NB + threadIdx.x < g.Np && threadIdx.x < NB) {
        const int gi = ti * NB + threadIdx.x;
        g.y[(size_t)b * g.Np + gi] = gi < g.N ? yin[gi] : 0.0;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { g.logdet[b] = 0.0; g.quad[b] = 0.0; g.fail[b] = 0; }
    __syncthreads();
    for (int e = threadIdx.x; e < NB * NB; e += blockDim.x) {
        const int c = e / NB, r = e - c * NB;  // column-major inside the tile: consecutive threads -> consecutive rows
        const int gi = ti * NB + r, gj = tj * NB + c;
        if (gi < gj) continue;
        double v;
        if (gi >= g.N || gj >= g.N) v = (gi == gj) ? 1.0 : 0.0;
        else {
            double d2 = 0.0;
            for (int k = 0; k < dim; ++k) { const double df = xi[r * dim + k] - xj[c * dim + k]; d2 = fma(df, df, d2); }
            v = sf2 * exp(-d2 * il2);
            if (gi == gj) v += sn2;
        }
        A[(size_t)gj * g.Np + gi] = v;
    }
}

// ---------------------------------------------------------------------------------------------------------
// diagonal block: Cholesky, inverse, z_k, logdet/quad.  One CTA (256 threads) per matrix; smem NB x (NB+1).
__global__ void __launch_bounds__(256) gp_potf2_kernel(GpBatch g, int k0) {
    extern __shared__ double s[];  // [NB][NB+1] row-major: s[i*(NB+1)+j]
    constexpr int LD = NB + 1;
    __shared__ double s_z[NB];
    __shared__ int s_fail;
    const int b = blockIdx.x, tid = threadIdx.x;
    if (g.fail[b]) return;
    double *A = g.A + (size_t)b * g.Np * g.Np;
    for (int e = tid; e < NB * NB; e += blockDim.x) {
        const int c = e / NB, r = e - c * NB;
        s[r * LD + c] = (r >= c) ? A[(size_t)(k0 + c) * g.Np + k0 + r] : 0.0;
    }
    if (tid == 0) s_fail = 0;
    __syncthreads();
    // left-looking (Crout) column sweep: two threads per row split the dot product L[i][0:j] . L[j][0:j] by the
    // parity of k and keep two accumulators each; one barrier pair per column, no index arithmetic in the loop
    {
        const int i = tid >> 1, half = tid & 1;
        for (int j = 0; j < NB; ++j) {
            double v = 0.0;
            if (i >= j) {
                double a0 = 0.0, a1 = 0.0;
                const double *ri = s + i * LD, *rj = s + j * LD;
                int k = half;
                for (; k + 2 < j; k += 4) {
                    a0 = fma(ri[k], rj[k], a0);
                    a1 = fma(ri[k + 2], rj[k + 2], a1);
                }
                if (k < j) a0 = fma(ri[k], rj[k], a0);
                v = a0 + a1;
            }
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v = s[i * LD + j] - v;  // only meaningful for i >= j
            if (i == j && half == 0) {
                if (!(v > 0.0) || !isfinite(v)) s_fail = 1;  // not positive definite -> logzero (GP:131-135)
                s[j * LD + j] = sqrt(v);
            }
            __syncthreads();
            if (s_fail) break;
            if (i > j && half == 0) s[i * LD + j] = v / s[j * LD + j];
            __syncthreads();
        }
    }
    __syncthreads();
    if (s_fail) {
        if (tid == 0) g.fail[b] = 1;
        return;
    }
    // write L11 back
    for (int e = tid; e < NB * NB; e += blockDim.x) {
        const int c = e / NB, r = e - c * NB;
        if (r >= c) A[(size_t)(k0 + c) * g.Np + k0 + r] = s[r * LD + c];
    }
    // X = L11^-1, one column per thread (forward substitution), written transposed for the GEMM's B operand
    // The strict upper triangle of s is free: thread c keeps column c of X below the diagonal in row c of it,
    // X[k][c] (k > c) at s[c*LD + k]; X[c][c] = 1 / l_cc.
    __syncthreads();
    double *linvT = g.linvT + (size_t)b * NB * NB;
    if (tid < NB) {
        const int c = tid;
        const double xcc = 1.0 / s[c * LD + c];
        for (int i = c + 1; i < NB; ++i) {
            double a0 = s[i * LD + c] * xcc, a1 = 0.0, a2 = 0.0, a3 = 0.0;
            const double *ri = s + i * LD, *rc = s + c * LD;
            int k = c + 1;
            for (; k + 3 < i; k += 4) {
                a0 = fma(ri[k], rc[k], a0);
                a1 = fma(ri[k + 1], rc[k + 1], a1);
                a2 = fma(ri[k + 2], rc[k + 2], a2);
                a3 = fma(ri[k + 3], rc[k + 3], a3);
            }
            for (; k < i; ++k) a0 = fma(ri[k], rc[k], a0);
            s[c * LD + i] = -((a0 + a1) + (a2 + a3)) / ri[i];
        }
    }
    __syncthreads();
    for (int e = tid; e < NB * NB; e += blockDim.x) {
        const int k = e / NB, n = e - k * NB;  // (k, n) -> X[n][k]
        linvT[e] = (n > k) ? s[k * LD + n] : (n == k ? 1.0 / s[k * LD + k] : 0.0);
    }
    // z_k = L11^-1 y_k ; logdet += 2 sum log l_jj ; quad += z.z
    const double *y = g.y + (size_t)b * g.Np + k0;
    if (tid < NB) {
        double acc = y[tid] / s[tid * LD + tid];
        for (int c = 0; c < tid; ++c) acc = fma(s[c * LD + tid], y[c], acc);
        s_z[tid] = acc;
        g.z[(size_t)b * NB + tid] = acc;
    }
    __syncthreads();
    if (tid < 32) {
        double ld = 0.0, q = 0.0;
        for (int i = tid; i < NB; i += 32) { ld += log(s[i * LD + i]); q = fma(s_z[i], s_z[i], q); }
        ld = warp_sum(ld);
        q = warp_sum(q);
        if (tid == 0) { g.logdet[b] += 2.0 * ld; g.quad[b] += q; }
    }
}

// ---------------------------------------------------------------------------------------------------------
// shared GEMM core: acc[4][4][2] (32x32 per warp, 16 warps -> 128x128) += sign * A(128 x KC) B(128 x KC)^T
// sA, sB: [KC][LDS_] (k-major).  Fragment maps of mma.m8n8k4.f64: a[row = lane/4][k = lane%4],
// b[k = lane%4][col = lane/4], c[row = lane/4][col = 2*(lane%4) + {0,1}].
template <int LDB>
__device__ __forceinline__ void gemm_chunk(const double *__restrict__ sA, const double *__restrict__ sB, int wm, int wn,
                                           int lane, double (&acc)[4][4][2], bool negate) {
    const int r = lane >> 2, q = lane & 3;
#pragma unroll
    for (int kk = 0; kk < KC; kk += 4) {
        double a[4], bb[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const double v = sA[(kk + q) * LDS_ + wm * 32 + i * 8 + r];
            a[i] = negate ? -v : v;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) bb[j] = sB[(kk + q) * LDB + wn * 32 + j * 8 + r];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) dmma_8x8x4(acc[i][j][0], acc[i][j][1], a[i], bb[j]);
    }
}

// load a ROWS x KC column-major panel chunk (columns contiguous in global) into smem [KC][LDD] with cp.async
template <int ROWS = NB, int LDD = LDS_>
__device__ __forceinline__ void load_chunk_async(double *sdst, const double *__restrict__ gsrc, size_t ld) {
    for (int e = threadIdx.x; e < KC * (ROWS / 2); e += blockDim.x) {
        const int k = e / (ROWS / 2), m2 = e - k * (ROWS / 2);
        cp_async16(sdst + k * LDD + 2 * m2, gsrc + (size_t)k * ld + 2 * m2);
    }
}

// trailing update: C(I,J) -= P_I P_J^T on 128 x 64 tiles (I: 128-row blocks, J: 64-column blocks) that touch the
// lower triangle of the trailing matrix (starts at k0 + NB).  8 warps (4 x 2), 32 x 32 per warp; two CTAs per SM so
// that one CTA's C-tile load/store overlaps the other's DMMA stream (a single 128 x 128 CTA per SM left the tensor
// pipe idle 38 % of the time while C moved).
constexpr int NBH = NB / 2;
constexpr int LDSH_ = NBH + 4;
__global__ void __launch_bounds__(256, 2) gp_syrk_kernel(GpBatch g, int k0) {
    extern __shared__ __align__(16) double sm[];  // 2 stages x (A chunk [KC][LDS_] + B chunk [KC][LDSH_])
    const int b = blockIdx.y;
    if (g.fail[b]) return;
    int t = blockIdx.x, ti = 0;
    while ((ti + 1) * (ti + 2) <= t) ++ti;
    const int tj = t - ti * (ti + 1);  // 0 .. 2*ti + 1
    const int base = k0 + NB;
    const int i0 = base + ti * NB, j0 = base + tj * NBH;
    double *A = g.A + (size_t)b * g.Np * g.Np;
    const size_t ld = g.Np;
    const double *PA = A + (size_t)k0 * ld + i0, *PB = A + (size_t)k0 * ld + j0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wm = warp >> 1, wn = warp & 1;
    const int r = lane >> 2, q = lane & 3;
    constexpr int STAGE = KC * LDS_ + KC * LDSH_;
    double *sA[2] = {sm, sm + STAGE}, *sB[2] = {sm + KC * LDS_, sm + STAGE + KC * LDS_};
    load_chunk_async<NB, LDS_>(sA[0], PA, ld);
    load_chunk_async<NBH, LDSH_>(sB[0], PB, ld);
    cp_async_commit();

    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int gi = i0 + wm * 32 + i * 8 + r, gj = j0 + wn * 32 + j * 8 + 2 * q + h;
                acc[i][j][h] = (gi >= gj) ? A[(size_t)gj * ld + gi] : 0.0;
            }
    constexpr int NCH = NB / KC;
#pragma unroll 1
    for (int c = 0; c < NCH; ++c) {
        if (c + 1 < NCH) {
            load_chunk_async<NB, LDS_>(sA[(c + 1) & 1], PA + (size_t)(c + 1) * KC * ld, ld);
            load_chunk_async<NBH, LDSH_>(sB[(c + 1) & 1], PB + (size_t)(c + 1) * KC * ld, ld);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        gemm_chunk<LDSH_>(sA[c & 1], sB[c & 1], wm, wn, lane, acc, true);
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int gi = i0 + wm * 32 + i * 8 + r, gj = j0 + wn * 32 + j * 8 + 2 * q + h;
                if (gi >= gj) A[(size_t)gj * ld + gi] = acc[i][j][h];
            }
}

// panel solve: L21(tile) = A21(tile) L11^-T as a GEMM against the transposed inverse; then y_tile -= L21 z_k
__global__ void __launch_bounds__(512) gp_trsm_kernel(GpBatch g, int k0) {
    extern __shared__ __align__(16) double sm[];  // A tile [NB][LDS_] + 2 x B chunk [KC][LDS_] + z[NB]
    const int b = blockIdx.y;
    if (g.fail[b]) return;
    const int i0 = k0 + NB + blockIdx.x * NB;
    double *A = g.A + (size_t)b * g.Np * g.Np;
    const size_t ld = g.Np;
    double *sAfull = sm;                       // [NB (k)][LDS_]
    double *sB[2] = {sm + NB * LDS_, sm + NB * LDS_ + KC * LDS_};
    double *sz = sm + NB * LDS_ + 2 * KC * LDS_;
    const double *linvT = g.linvT + (size_t)b * NB * NB;  // (k, n) at k*NB + n : already k-major
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wm = warp >> 2, wn = warp & 3;
    const int r = lane >> 2, q = lane & 3;
    // whole A21 tile (all 128 k-columns) first: it is overwritten in place afterwards
    for (int c = 0; c < NB / KC; ++c) load_chunk_async(sAfull + c * KC * LDS_, A + (size_t)(k0 + c * KC) * ld + i0, ld);
    load_chunk_async(sB[0], linvT, NB);
    cp_async_commit();
    if (threadIdx.x < NB) sz[threadIdx.x] = g.z[(size_t)b * NB + threadIdx.x];
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    constexpr int NCH = NB / KC;
#pragma unroll 1
    for (int c = 0; c < NCH; ++c) {
        if (c + 1 < NCH) {
            load_chunk_async(sB[(c + 1) & 1], linvT + (size_t)(c + 1) * KC * NB, NB);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        gemm_chunk<LDS_>(sAfull + c * KC * LDS_, sB[c & 1], wm, wn, lane, acc, false);
        __syncthreads();
    }
    // results: write L21 in place and keep a copy in smem ([n][m] over the A tile buffer) for the y update
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int m = wm * 32 + i * 8 + r, n = wn * 32 + j * 8 + 2 * q + h;
                A[(size_t)(k0 + n) * ld + i0 + m] = acc[i][j][h];
                sAfull[n * LDS_ + m] = acc[i][j][h];
            }
    __syncthreads();
    if (threadIdx.x < NB) {
        const int m = threadIdx.x;
        double s = 0.0;
        for (int n = 0; n < NB; ++n) s = fma(sAfull[n * LDS_ + m], sz[n], s);
        g.y[(size_t)b * g.Np + i0 + m] -= s;
    }
}

__global__ void gp_finish_kernel(GpBatch g, const double *__restrict__ theta, int Ps, int b0,
                                 const __grid_constant__ PriorSpec prior, double logzero, int check_box,
                                 double *__restrict__ out, int out_stride) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= g.B) return;
    double th[3];
    for (int j = 0; j < 3; ++j) th[j] = theta[(size_t)j * Ps + b0 + b];
    double v = -0.5 * ((double)g.N * kLog2Pi + g.logdet[b] + g.quad[b]);
    const double lim = fabs(logzero);
    v = fmin(fmax(v, -lim), lim);  // Clip[..., +-|logzero|] GP:190-197
    const bool ok = th[0] > 0.0 && th[1] > 0.0 && th[2] > 0.0;
    if (g.fail[b] || !ok || !isfinite(v)) v = logzero;
    if (check_box && !in_box<3>(prior, th)) v = logzero;
    out[(size_t)(b0 + b) * out_stride] = v;
}

}  // namespace

// theta_dev SoA [3][Ps]; out_dev[(w) * out_stride]
void gp_loglike_device_strided(binest_problem &p, const double *theta_dev, int P, int Ps, double *out_dev,
                               int out_stride, bool check_box) {
    const int N = (int)p.gp_n, Np = (N + NB - 1) / NB * NB, T = Np / NB;
    size_t free_b = 0, total_b = 0;
    BN_CUDA(cudaMemGetInfo(&free_b, &total_b));
    const size_t per = (size_t)Np * Np * sizeof(double);
    // workspace: reuse across calls; sized for as many matrices as fit in 60 % of the free memory
    static thread_local DevBuf<double> wsA, wsY, wsZ, wsL, wsLd, wsQ;
    static thread_local DevBuf<int> wsF;
    static thread_local int ws_cap = 0;
    static thread_local int ws_np = 0;
    int want = std::min<int>(P, 512);
    if (ws_cap < want || ws_np != Np) {
        wsA.release();
        BN_CUDA(cudaMemGetInfo(&free_b, &total_b));
        int fit = (int)std::min<size_t>((size_t)want, (size_t)(0.6 * (double)free_b) / per);
        BN_REQUIRE(fit >= 1, BINEST_ERR_MEMORY, "not enough device memory for one GP covariance matrix");
        wsA.alloc((size_t)fit * Np * Np);
        wsY.alloc((size_t)fit * Np); wsZ.alloc((size_t)fit * NB); wsL.alloc((size_t)fit * NB * NB);
        wsLd.alloc(fit); wsQ.alloc(fit); wsF.alloc(fit);
        ws_cap = fit; ws_np = Np;
    }
    const size_t smem_potf2 = (size_t)NB * (NB + 1) * sizeof(double);
    const size_t smem_syrk = (size_t)2 * (KC * LDS_ + KC * LDSH_) * sizeof(double);
    const size_t smem_trsm = ((size_t)NB * LDS_ + 2 * KC * LDS_ + NB) * sizeof(double);
    BN_CUDA(cudaFuncSetAttribute(gp_potf2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_potf2));
    BN_CUDA(cudaFuncSetAttribute(gp_syrk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_syrk));
    BN_CUDA(cudaFuncSetAttribute(gp_trsm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_trsm));
    cudaStream_t s = p.stream;
    for (int b0 = 0; b0 < P; b0 += ws_cap) {
        const int B = std::min(ws_cap, P - b0);
        GpBatch g{wsA.p, wsY.p, wsZ.p, wsL.p, wsLd.p, wsQ.p, wsF.p, Np, N, B};
        gp_fill_kernel<<<dim3(T * (T + 1) / 2, B), 256, 2 * NB * p.gp_dim * sizeof(double), s>>>(
            g, p.gp_x.p, (int)p.gp_dim, p.gp_y.p, theta_dev, Ps, b0);
        BN_LAUNCH_CHECK();
        for (int k = 0; k < T; ++k) {
            const int k0 = k * NB, rest = T - k - 1;
            gp_potf2_kernel<<<B, 256, smem_potf2, s>>>(g, k0);
            BN_LAUNCH_CHECK();
            if (rest > 0) {
                gp_trsm_kernel<<<dim3(rest, B), 512, smem_trsm, s>>>(g, k0);
                BN_LAUNCH_CHECK();
                gp_syrk_kernel<<<dim3(rest * (rest + 1), B), 256, smem_syrk, s>>>(g, k0);
                BN_LAUNCH_CHECK();
            }
        }
        gp_finish_kernel<<<(B + 127) / 128, 128, 0, s>>>(g, theta_dev, Ps, b0, p.prior, g_logzero, check_box ? 1 : 0,
                                                         out_dev, out_stride);
        BN_LAUNCH_CHECK();
    }
}

void gp_loglike_device(binest_problem &p, const double *theta_dev, int P, int Ps, double *out_dev, bool check_box) {
    gp_loglike_device_strided(p, theta_dev, P, Ps, out_dev, 1, check_box);
}

}  // namespace binest
