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
//
// predictFromGaussianProcess (GP:332-422, compiledKandKappa GP:92-116) rides on the same sweep: the Q prediction inputs
// are appended as Qp = roundup(Q, 128) extra ROWS below every matrix (leading dimension ld = Np + Qp), filled with
// the cross-covariances k(x*_q, x_j).  The panel solve then turns those rows into V = (L^-1 K*)^T tile by tile, its
// fused right-hand-side update accumulates -V z = -k*.K^-1 y (the predictive mean) in y[Np + q], and the row sums of
// squares of V give k*.K^-1 k* for the predictive variance — no factor is ever re-read and no back substitution runs.
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "problem.cuh"

namespace binest {
namespace {

constexpr int NB = 128;        // panel width / tile edge
constexpr int KC = 32;         // K chunk of the GEMM kernels
constexpr int LDS_ = NB + 4;   // smem leading dimension: half-warp fragment loads hit 16 distinct 8-byte banks
constexpr int kGpBlockPanels = 16;
constexpr int kGpStreams = 4;      // sub-batches of a chunk swept on separate streams (BINEST_GP_STREAMS overrides, 1..8)  // panels per group of the two-level blocking (BINEST_GP_BLOCK overrides, 1..8)

struct GpBatch {
    double *A;        // [B][ld*Np] column-major: Np columns of ld = Np + Qp rows (Qp = 0 for the likelihood)
    double *y;        // [B][ld] working right-hand side (consumed by the fused forward solve); rows >= Np: -mean
    double *z;        // [B][NB]  z_k of the current panel
    double *linvT;    // [B][NB*NB] (k, n) -> L11^-1[n][k]
    double *logdet;   // [B]
    double *quad;     // [B]
    int *fail;        // [B]
    int Np, N, B;
    int ld;           // rows per column
    int Q;            // prediction inputs (rows Np .. Np + Q - 1 are live, the rest of the Qp block is zero)
    const double *xs; // [Q][dim] prediction inputs (device) or nullptr
    double *v2;       // [B][ld] row sums of squares of the solved prediction rows (k*.K^-1 k*), or nullptr
    __host__ __device__ size_t mat() const { return (size_t)ld * Np; }
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
    // decode the tile index: lower-triangular tiles of the square part first, then the Qp x Np prediction rows
    const int T = g.Np / NB, tri = T * (T + 1) / 2;
    int t = blockIdx.x, ti = 0, tj;
    if (t < tri) {
        while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
        tj = t - ti * (ti + 1) / 2;
    } else {
        ti = T + (t - tri) / T;
        tj = (t - tri) % T;
    }
    const bool pred = ti >= T;
    const double sf = theta[0 * (size_t)Ps + b0 + b], ell = theta[1 * (size_t)Ps + b0 + b], sn = theta[2 * (size_t)Ps + b0 + b];
    const double sf2 = sf * sf, il2 = 1.0 / (2.0 * ell * ell), sn2 = sn * sn;
    double *A = g.A + (size_t)b * g.mat();
    extern __shared__ double sx[];  // [2][NB][dim]
    double *xi = sx, *xj = sx + NB * dim;
    for (int k = threadIdx.x; k < NB * dim; k += blockDim.x) {
        const int r = k / dim, c = k - r * dim;
        const int gi = ti * NB + r, gj = tj * NB + r;
        if (pred) xi[k] = gi - g.Np < g.Q ? g.xs[(size_t)(gi - g.Np) * dim + c] : 0.0;
        else xi[k] = gi < g.N ? x[(size_t)gi * dim + c] : 0.0;
        xj[k] = gj < g.N ? x[(size_t)gj * dim + c] : 0.0;
    }
    if (tj == 0 && threadIdx.x < NB) {
        const int gi = ti * NB + threadIdx.x;  // < ld
        g.y[(size_t)b * g.ld + gi] = gi < g.N ? yin[gi] : 0.0;
        if (g.v2) g.v2[(size_t)b * g.ld + gi] = 0.0;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { g.logdet[b] = 0.0; g.quad[b] = 0.0; g.fail[b] = 0; }
    __syncthreads();
    for (int e = threadIdx.x; e < NB * NB; e += blockDim.x) {
        const int c = e / NB, r = e - c * NB;  // column-major inside the tile: consecutive threads -> consecutive rows
        const int gi = ti * NB + r, gj = tj * NB + c;
        if (gi < gj) continue;
        double v;
        const bool live_row = pred ? (gi - g.Np < g.Q) : (gi < g.N);
        if (!live_row || gj >= g.N) v = (gi == gj) ? 1.0 : 0.0;
        else {
            double d2 = 0.0;
            for (int k = 0; k < dim; ++k) { const double df = xi[r * dim + k] - xj[c * dim + k]; d2 = fma(df, df, d2); }
            // branch-free exp (operators.cuh: Cody-Waite + degree-10 polynomial, 3.4e-16 relative) instead of libdevice's
            // (~50 integer instructions of constant set-up per call): the fill was 3.5 % of a B = 256 sweep
            // (profiles/r02_gp_launches.md).  Below -700 the true value is < 1e-304 sigma_f^2: taken as 0.
            const double arg = -d2 * il2;
            double ev[1] = {0.0};
            if (arg > -700.0) {
                const double ax[1] = {arg};
                exp_bounded<1>(ax, ev);
            }
            v = sf2 * ev[0];   // prediction rows: the cross-covariance k(x*, x_j), no nugget (GP:103-109)
            if (gi == gj) v += sn2;
        }
        A[(size_t)gj * g.ld + gi] = v;
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
    double *A = g.A + (size_t)b * g.mat();
    for (int e = tid; e < NB * NB; e += blockDim.x) {
        const int c = e / NB, r = e - c * NB;
        s[r * LD + c] = (r >= c) ? A[(size_t)(k0 + c) * g.ld + k0 + r] : 0.0;
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
        if (r >= c) A[(size_t)(k0 + c) * g.ld + k0 + r] = s[r * LD + c];
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
    const double *y = g.y + (size_t)b * g.ld + k0;
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
// diagonal block, register-tiled (the default).  The column sweep of gp_potf2_kernel is a chain of long dot
// products (2 threads per row) and its inverse a chain of 128 dependent substitutions per thread: 0.25 ms per panel,
// 22 % of a B = 32 sweep (profiles/r01f_gp_launches.md).  Here both are right-looking outer-product sweeps over a
// register-resident tile: thread (tx, ty) of a 16 x 16 grid owns the elements (i, c) = (ty + 16 a, tx + 16 b),
// a, b = 0..7 — cyclic, so the shrinking trailing matrix stays balanced — and every step is one barrier:
//   factor, step j:  the warp that owns column j scales it (diagonal broadcast by shuffle, rsqrt) and publishes it in
//                    shared memory (column j of Ls); after the barrier everybody applies the rank-1 update to the
//                    elements it owns right of column j (<= 36 independent DFMAs per thread);
//   invert, step k:  R starts as I; the owners of row k scale it by 1/l_kk (that is row k of X = L^-1) and publish it;
//                    after the barrier rows i > k subtract l_ik x_k,: .
// tid = 16 tx + ty: the owners of a column sit in one warp (shuffle), the owners of a row are spread over all warps.
__global__ void __launch_bounds__(256, 1) gp_potf2_reg_kernel(GpBatch g, int k0) {
    extern __shared__ __align__(16) double sm2[];
    double *Ls = sm2;                 // [NB (column j)][NB (row i)]: L, column-major
    double *rowbuf = Ls + NB * NB;    // [2][NB] row k of X, double buffered
    double *dg = rowbuf + 2 * NB;     // [NB] l_jj
    double *zp = dg + NB;             // [NB][17] partial sums of z = X y
    __shared__ double s_z[NB];
    __shared__ int s_fail[2];  // alternating: a fast warp may already be flagging step j + 1 while a slow one reads step j
    const int b = blockIdx.x, tid = threadIdx.x, tx = tid >> 4, ty = tid & 15;
    if (g.fail[b]) return;
    double *A = g.A + (size_t)b * g.mat();
    const size_t ld = g.ld;
    double acc[8][8];
#pragma unroll
    for (int bb = 0; bb < 8; ++bb)
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            const int i = ty + 16 * a, c = tx + 16 * bb;
            acc[a][bb] = (i >= c) ? A[(size_t)(k0 + c) * ld + k0 + i] : 0.0;
        }
    if (tid == 0) s_fail[0] = s_fail[1] = 0;
    __syncthreads();

    // ---- factor
    bool failed = false;
#pragma unroll
    for (int jb = 0; jb < 8; ++jb) {
        if (failed) break;
#pragma unroll 1
        for (int jt = 0; jt < 16; ++jt) {
            const int j = 16 * jb + jt;
            if ((tx >> 1) == (jt >> 1)) {  // the warp holding column j
                double d = acc[jb][jb];
                d = __shfl_sync(0xffffffffu, d, ((jt & 1) << 4) | jt);  // a_jj lives at tx == ty == jt
                const bool bad = !(d > 0.0) || !isfinite(d);
                const double rinv = rsqrt(d);
                if (tx == jt) {
#pragma unroll
                    for (int a = 0; a < 8; ++a) {
                        const int i = ty + 16 * a;
                        double l = 0.0;
                        if (a >= jb && i >= j) {
                            l = acc[a][jb] * rinv;   // i == j: d / sqrt(d) = sqrt(d)
                            acc[a][jb] = l;
                        }
                        Ls[j * NB + i] = l;
                    }
                    if (ty == jt) {
                        dg[j] = d * rinv;
                        if (bad) s_fail[j & 1] = 1;  // not positive definite -> logzero (GP:131-135)
                    }
                }
            }
            __syncthreads();
            if (s_fail[j & 1]) { failed = true; break; }
            double lr[8], lc[8];
#pragma unroll
            for (int a = jb; a < 8; ++a) lr[a] = Ls[j * NB + ty + 16 * a];
#pragma unroll
            for (int bb = jb; bb < 8; ++bb) lc[bb] = Ls[j * NB + tx + 16 * bb];
#pragma unroll
            for (int bb = jb; bb < 8; ++bb) {
                if (bb > jb || tx > jt) {  // columns right of j only
#pragma unroll
                    for (int a = bb; a < 8; ++a) acc[a][bb] = fma(-lr[a], lc[bb], acc[a][bb]);
                }
            }
        }
    }
    if (failed) {
        if (tid == 0) g.fail[b] = 1;
        return;
    }
    // L11 back to the matrix
#pragma unroll
    for (int bb = 0; bb < 8; ++bb)
#pragma unroll
        for (int a = bb; a < 8; ++a) {
            const int i = ty + 16 * a, c = tx + 16 * bb;
            if (i >= c) A[(size_t)(k0 + c) * ld + k0 + i] = acc[a][bb];
        }

    // ---- invert: X = L^-1 (lower triangular), R = I to start with
#pragma unroll
    for (int bb = 0; bb < 8; ++bb)
#pragma unroll
        for (int a = 0; a < 8; ++a) acc[a][bb] = (a == bb && tx == ty) ? 1.0 : 0.0;
#pragma unroll
    for (int kb = 0; kb < 8; ++kb) {
#pragma unroll 1
        for (int kt = 0; kt < 16; ++kt) {
            const int k = 16 * kb + kt;
            double *rb = rowbuf + (k & 1) * NB;
            if (ty == kt) {  // owners of row k: x_k,c = r_k,c / l_kk for c <= k
                const double dinv = 1.0 / dg[k];
#pragma unroll
                for (int bb = 0; bb <= kb; ++bb) {
                    const int c = tx + 16 * bb;
                    const double x = (c <= k) ? acc[kb][bb] * dinv : 0.0;
                    acc[kb][bb] = x;
                    rb[c] = x;
                }
            }
            __syncthreads();
            double lr[8], xc[8];
#pragma unroll
            for (int a = kb; a < 8; ++a) lr[a] = Ls[k * NB + ty + 16 * a];
#pragma unroll
            for (int bb = 0; bb <= kb; ++bb) xc[bb] = rb[tx + 16 * bb];
#pragma unroll
            for (int a = kb; a < 8; ++a) {
                if (a > kb || ty > kt) {  // rows below k only
#pragma unroll
                    for (int bb = 0; bb <= kb; ++bb) acc[a][bb] = fma(-lr[a], xc[bb], acc[a][bb]);
                }
            }
        }
    }
    // transposed inverse for the panel GEMM: linvT[(k, n)] = X[n][k]
    double *linvT = g.linvT + (size_t)b * NB * NB;
#pragma unroll
    for (int bb = 0; bb < 8; ++bb)
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            const int i = ty + 16 * a, c = tx + 16 * bb;
            linvT[c * NB + i] = (i >= c) ? acc[a][bb] : 0.0;
        }
    // z_k = X y_k ; logdet += 2 sum log l_jj ; quad += z.z
    {
        const double *y = g.y + (size_t)b * ld + k0;
        double yc[8];
#pragma unroll
        for (int bb = 0; bb < 8; ++bb) yc[bb] = y[tx + 16 * bb];
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            double s = 0.0;
#pragma unroll
            for (int bb = 0; bb <= a; ++bb) s = fma(acc[a][bb], yc[bb], s);  // acc is 0 above the diagonal
            zp[(ty + 16 * a) * 17 + tx] = s;
        }
    }
    __syncthreads();
    if (tid < NB) {
        double s = 0.0;
#pragma unroll
        for (int t = 0; t < 16; ++t) s += zp[tid * 17 + t];
        s_z[tid] = s;
        g.z[(size_t)b * NB + tid] = s;
    }
    __syncthreads();
    if (tid < 32) {
        double ldt = 0.0, q = 0.0;
        for (int i = tid; i < NB; i += 32) { ldt += log(dg[i]); q = fma(s_z[i], s_z[i], q); }
        ldt = warp_sum(ldt);
        q = warp_sum(q);
        if (tid == 0) { g.logdet[b] += 2.0 * ldt; g.quad[b] += q; }
    }
}

// ---------------------------------------------------------------------------------------------------------
// shared GEMM core: acc[4][4][2] (32x32 per warp, 16 warps -> 128x128) += sign * A(128 x KC) B(128 x KC)^T
// sA, sB: [KC][LDS_] (k-major).  Fragment maps of mma.m8n8k4.f64: a[row = lane/4][k = lane%4],
// b[k = lane%4][col = lane/4], c[row = lane/4][col = 2*(lane%4) + {0,1}].
template <int LDB, int KCH = KC>
__device__ __forceinline__ void gemm_chunk(const double *__restrict__ sA, const double *__restrict__ sB, int wm, int wn,
                                           int lane, double (&acc)[4][4][2], bool negate) {
    const int r = lane >> 2, q = lane & 3;
#pragma unroll
    for (int kk = 0; kk < KCH; kk += 4) {
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
// Two-level blocking.  Panels are factored in groups of `nblk`: inside a group the finished panels only update the
// column blocks of their own group (ncol64 != 0: the trapezoid rows >= base, columns [base, base + 64 ncol64), see the
// binary schedule in gp_sweep), and when the group is done ONE pass with K = kw = nblk * 128 updates everything to the
// right of it — the trailing matrix crosses HBM once per group instead of once per panel (the DMMA work is unchanged).
__global__ void __launch_bounds__(256, 2) gp_syrk_kernel(GpBatch g, int k0, int kw, int base, int ncol64) {
    extern __shared__ __align__(16) double sm[];  // 2 stages x (A chunk [KC][LDS_] + B chunk [KC][LDSH_])
    const int b = blockIdx.y;
    if (g.fail[b]) return;
    const int rest = (g.Np - base) / NB;  // row blocks of the square part at and below `base`
    int t = blockIdx.x, ti = 0, tj;
    if (ncol64 > 0) {  // narrow: ti over all row blocks (prediction rows included), tj < ncol64
        ti = t / ncol64;
        tj = t - ti * ncol64;
        if (tj * NBH > ti * NB + NB - 1) return;  // tile entirely above the diagonal
    } else if (t < rest * (rest + 1)) {  // lower-triangular tiles of the square trailing part ...
        while ((ti + 1) * (ti + 2) <= t) ++ti;
        tj = t - ti * (ti + 1);  // 0 .. 2*ti + 1
    } else {                     // ... then the prediction rows (every 64-column block of the trailing part)
        t -= rest * (rest + 1);
        ti = rest + t / (2 * rest);
        tj = t % (2 * rest);
    }
    const int i0 = base + ti * NB, j0 = base + tj * NBH;
    double *A = g.A + (size_t)b * g.mat();
    const size_t ld = g.ld;
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
    const int NCH = kw / KC;
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

// ---------------------------------------------------------------------------------------------------------
// trailing update, TMA-fed — an EXPERIMENT kept for the record (BINEST_GP_SYRK=1 selects it; gp_syrk_kernel above is the
// default).  Measured on B200, N = 4096, B = 256 (r2l): 242.1 ms per batch against 222.2 ms with the cp.async kernel —
// the panel chunks are rows of 1 KB / 512 B, so a chunk is 32 separate bulk copies whose fixed cost the TMA engine does
// not amortise, while 256 threads issuing 16-byte cp.async spread the same traffic over all LSUs; and the tensor pipe
// was never starved by the two barriers per chunk (83.6 % active, the rest is the C-tile prologue/epilogue).  Same
// tiles, same DMMA core and the same arithmetic order as gp_syrk_kernel; what changes is how the panel chunks reach
// shared memory and how the warps synchronise:
//   * every K-row of a chunk (128 resp. 64 consecutive doubles of a panel column) is ONE bulk copy of the TMA engine
//     (cp.async.bulk -> SASS UBLKCP) into the padded row of the stage, issued by one thread, completion counted on the
//     stage's `full` mbarrier — instead of 12 cp.async per thread and chunk;
//   * a ring of 4 stages of 16 K-rows (same 100 KB as the two 32-row stages before, so still two CTAs per SM), filled
//     three chunks ahead;
//   * no CTA-wide barrier in the K loop: a warp waits on `full[s]`, multiplies, and arrives on `empty[s]`; only the
//     issuing thread waits for `empty` — of the chunk BEFORE the one it has just finished — before it refills that
//     stage.  (gp_syrk_kernel paid two __syncthreads per 32 rows of K: 64 lock-step points per tile.)
// Every mbarrier wait is bounded (~2^30 polls -> trap): a protocol error ends the kernel with an error, not a hang.
constexpr int WS_KC = 16, WS_NS = 4;
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait_bounded(uint64_t *bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
    for (unsigned it = 0;; ++it) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(a), "r"(parity)
            : "memory");
        if (ok) return;
        if (it > (1u << 30)) __trap();
    }
}

__global__ void __launch_bounds__(256, 2) gp_syrk_ws_kernel(GpBatch g, int k0, int kw, int base, int ncol64) {
    extern __shared__ __align__(128) double sm[];  // WS_NS stages x (A chunk [WS_KC][LDS_] + B chunk [WS_KC][LDSH_])
    __shared__ uint64_t full[WS_NS], empty[WS_NS];
    const int b = blockIdx.y;
    if (g.fail[b]) return;
    const int rest = (g.Np - base) / NB;
    int t = blockIdx.x, ti = 0, tj;
    if (ncol64 > 0) {
        ti = t / ncol64;
        tj = t - ti * ncol64;
        if (tj * NBH > ti * NB + NB - 1) return;
    } else if (t < rest * (rest + 1)) {
        while ((ti + 1) * (ti + 2) <= t) ++ti;
        tj = t - ti * (ti + 1);
    } else {
        t -= rest * (rest + 1);
        ti = rest + t / (2 * rest);
        tj = t % (2 * rest);
    }
    const int i0 = base + ti * NB, j0 = base + tj * NBH;
    double *A = g.A + (size_t)b * g.mat();
    const size_t ld = g.ld;
    const double *PA = A + (size_t)k0 * ld + i0, *PB = A + (size_t)k0 * ld + j0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wm = warp >> 1, wn = warp & 1;
    const int r = lane >> 2, q = lane & 3;
    constexpr int STAGE = WS_KC * LDS_ + WS_KC * LDSH_;
    constexpr uint32_t kStageBytes = WS_KC * (NB + NBH) * sizeof(double);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < WS_NS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 8); }
        mbar_fence_init();
    }
    __syncthreads();
    const int NCH = kw / WS_KC;
    auto issue = [&](int c) {  // one thread: the WS_KC rows of chunk c of both panels -> stage c % WS_NS
        const int s = c % WS_NS;
        double *sA = sm + (size_t)s * STAGE, *sB = sA + WS_KC * LDS_;
        mbar_expect_tx(&full[s], kStageBytes);
#pragma unroll 4
        for (int k = 0; k < WS_KC; ++k) {
            bulk_g2s(sA + k * LDS_, PA + (size_t)(c * WS_KC + k) * ld, NB * sizeof(double), &full[s]);
            bulk_g2s(sB + k * LDSH_, PB + (size_t)(c * WS_KC + k) * ld, NBH * sizeof(double), &full[s]);
        }
    };
    if (threadIdx.x == 0)
        for (int c = 0; c < WS_NS && c < NCH; ++c) issue(c);

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
#pragma unroll 1
    for (int c = 0; c < NCH; ++c) {
        // refill the stage of the chunk before this one (every warp is done with it, or about to be)
        if (threadIdx.x == 0 && c >= 1 && c - 1 + WS_NS < NCH) {
            mbar_wait_bounded(&empty[(c - 1) % WS_NS], (uint32_t)(((c - 1) / WS_NS) & 1));
            issue(c - 1 + WS_NS);
        }
        const int s = c % WS_NS;
        mbar_wait_bounded(&full[s], (uint32_t)((c / WS_NS) & 1));
        const double *sA = sm + (size_t)s * STAGE, *sB = sA + WS_KC * LDS_;
        gemm_chunk<LDSH_, WS_KC>(sA, sB, wm, wn, lane, acc, true);
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
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
    double *A = g.A + (size_t)b * g.mat();
    const size_t ld = g.ld;
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
    // L11^-1 is lower triangular: column n of the product only needs k <= n.  Each warp owns two 16-column halves, h
    // and 7 - h (h = wn), so that every warp skips the same number of all-zero K chunks: half h needs chunks
    // 0 .. h / 2, i.e. 5 half-chunks per warp instead of 8 (the panel solves were 11 % of a B = 256 sweep,
    // profiles/r02_gp_launches.md).
    const int cb0 = wn * 16, cb1 = (7 - wn) * 16;
    double acc[4][4][2];  // [i: 8-row block][j: 0,1 -> columns cb0 + 8 j; 2,3 -> columns cb1 + 8 (j - 2)]
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
        const bool use0 = c <= wn / 2, use1 = c <= (7 - wn) / 2;
        const double *sAc = sAfull + c * KC * LDS_, *sBc = sB[c & 1];
        if (use0 || use1) {
#pragma unroll
            for (int kk = 0; kk < KC; kk += 4) {
                double a[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = sAc[(kk + q) * LDS_ + wm * 32 + i * 8 + r];
                if (use0) {
                    const double b0 = sBc[(kk + q) * LDS_ + cb0 + r], b1 = sBc[(kk + q) * LDS_ + cb0 + 8 + r];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        dmma_8x8x4(acc[i][0][0], acc[i][0][1], a[i], b0);
                        dmma_8x8x4(acc[i][1][0], acc[i][1][1], a[i], b1);
                    }
                }
                if (use1) {
                    const double b0 = sBc[(kk + q) * LDS_ + cb1 + r], b1 = sBc[(kk + q) * LDS_ + cb1 + 8 + r];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        dmma_8x8x4(acc[i][2][0], acc[i][2][1], a[i], b0);
                        dmma_8x8x4(acc[i][3][0], acc[i][3][1], a[i], b1);
                    }
                }
            }
        }
        __syncthreads();
    }
    // results: write L21 in place and keep a copy in smem ([n][m] over the A tile buffer) for the y update
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int m = wm * 32 + i * 8 + r, n = (j < 2 ? cb0 + j * 8 : cb1 + (j - 2) * 8) + 2 * q + h;
                A[(size_t)(k0 + n) * ld + i0 + m] = acc[i][j][h];
                sAfull[n * LDS_ + m] = acc[i][j][h];
            }
    __syncthreads();
    if (threadIdx.x < NB) {
        const int m = threadIdx.x;
        double s = 0.0, ss = 0.0;
        for (int n = 0; n < NB; ++n) {
            const double l = sAfull[n * LDS_ + m];
            s = fma(l, sz[n], s);
            ss = fma(l, l, ss);
        }
        g.y[(size_t)b * g.ld + i0 + m] -= s;
        if (i0 >= g.Np) g.v2[(size_t)b * g.ld + i0 + m] += ss;  // prediction rows: this panel's share of |L^-1 k*|^2
    }
}

// predictive mean and standard deviation (GP:401-418): mean = k*.K^-1 y, sd = Sqrt[kappa - k*.K^-1 k*] with
// kappa = k(x*, x*) + nugget = sf^2 + sn^2 (GP:110-113).  A covariance that did not factor gives NaN (the reference
// Throws out of matrixInverseAndDet, GP:131-135).  Rounding can push the radicand of a point that coincides with a
// noise-free datum a few ulp below zero, where the reference would return a complex number: clamped to 0 here.
__global__ void gp_predict_finish_kernel(GpBatch g, const double *__restrict__ theta, int Ps, int b0, int64_t Qtot,
                                         double *__restrict__ mean, double *__restrict__ sd) {
    const int b = blockIdx.y, q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= g.Q) return;
    const double sf = theta[0 * (size_t)Ps + b0 + b], sn = theta[2 * (size_t)Ps + b0 + b], ell = theta[1 * (size_t)Ps + b0 + b];
    const size_t o = (size_t)(b0 + b) * Qtot + q;
    if (g.fail[b] || !(sf > 0.0 && ell > 0.0 && sn > 0.0)) {
        mean[o] = sd[o] = __longlong_as_double(0x7ff8000000000000LL);
        return;
    }
    mean[o] = -g.y[(size_t)b * g.ld + g.Np + q];
    const double var = fma(sf, sf, sn * sn) - g.v2[(size_t)b * g.ld + g.Np + q];
    sd[o] = sqrt(fmax(var, 0.0));
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

namespace {

// workspace shared by the likelihood and the prediction entry points: reused across calls, sized for as many
// matrices as fit in 60 % of the free memory
struct GpWorkspace {
    DevBuf<double> A, Y, Z, L, Ld, Qd, V2;
    DevBuf<int> F;
    int cap = 0, np = 0, ld = 0;
    void ensure(int want, int Np, int ld_) {
        if (cap >= want && np == Np && ld == ld_) return;
        A.release();
        size_t free_b = 0, total_b = 0;
        BN_CUDA(cudaMemGetInfo(&free_b, &total_b));
        const size_t per = (size_t)ld_ * Np * sizeof(double);
        const int fit = (int)std::min<size_t>((size_t)want, (size_t)(0.6 * (double)free_b) / per);
        BN_REQUIRE(fit >= 1, BINEST_ERR_MEMORY, "not enough device memory for one GP covariance matrix");
        A.alloc((size_t)fit * ld_ * Np);
        Y.alloc((size_t)fit * ld_); V2.alloc((size_t)fit * ld_); Z.alloc((size_t)fit * NB); L.alloc((size_t)fit * NB * NB);
        Ld.alloc(fit); Qd.alloc(fit); F.alloc(fit);
        cap = fit; np = Np; ld = ld_;
    }
};
thread_local GpWorkspace g_ws;

// auxiliary streams of the batch split (below): created once per host thread and device
struct GpStreams {
    int device = -1;
    std::vector<cudaStream_t> aux;
    std::vector<cudaEvent_t> done;
    std::vector<std::vector<cudaEvent_t>> panel;  // [sub-batch][panel]: diagonal block + panel solve finished
    cudaEvent_t fork = nullptr;
    void ensure_panels(int nsub, int T) {
        if ((int)panel.size() < nsub) panel.resize(nsub);
        for (int h = 0; h < nsub; ++h)
            while ((int)panel[h].size() < T) {
                cudaEvent_t ev;
                BN_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
                panel[h].push_back(ev);
            }
    }
    void ensure(int dev, int n) {
        if (device != dev) { aux.clear(); done.clear(); panel.clear(); fork = nullptr; device = dev; }  // (a thread serves one device)
        if (!fork) BN_CUDA(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming));
        while ((int)aux.size() < n) {
            cudaStream_t st; cudaEvent_t ev;
            BN_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
            BN_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            aux.push_back(st); done.push_back(ev);
        }
    }
};
thread_local GpStreams g_streams;

// blocked Cholesky sweep (with the fused forward solve) of matrices [h0, h0 + Bh) of the chunk on stream s
// wait_ev / rec_ev (per panel, may be null): staggering of the sub-batches — this sweep starts panel k only after the
// previous sub-batch has finished ITS diagonal block and panel solve of k (wait_ev[k]) and announces its own (rec_ev[k]),
// so that the latency-bound small kernels of one sub-batch run under the trailing updates of the others instead of
// all four sub-batches marching in phase.
void gp_sweep(const GpBatch &gc, int h0, int Bh, int nblk, cudaStream_t s, size_t smem_potf2, size_t smem_trsm,
              size_t smem_syrk, cudaEvent_t *wait_ev = nullptr, cudaEvent_t *rec_ev = nullptr) {
    GpBatch g = gc;
    g.A += (size_t)h0 * gc.mat(); g.y += (size_t)h0 * gc.ld; g.z += (size_t)h0 * NB; g.linvT += (size_t)h0 * NB * NB;
    g.logdet += h0; g.quad += h0; g.fail += h0; g.B = Bh;
    if (g.v2) g.v2 += (size_t)h0 * gc.ld;
    const int T = g.Np / NB, Tq = (g.ld - g.Np) / NB, rows_all = T + Tq, B = Bh;
    static const bool potf2_reg = [] { const char *e = getenv("BINEST_GP_POTF2"); return !(e && atoi(e) == 1); }();  // 1: column sweep
    static const bool syrk_ws = [] { const char *e = getenv("BINEST_GP_SYRK"); return e && atoi(e) == 1; }();  // 1: TMA-fed variant
    static const bool left_looking = [] { const char *e = getenv("BINEST_GP_LEFT"); return !(e && atoi(e) == 0); }();
    const size_t smem_ws = (size_t)WS_NS * (WS_KC * LDS_ + WS_KC * LDSH_) * sizeof(double);
    auto launch_syrk = [&](dim3 grid, int k0, int kw, int base, int ncol64) {
        if (syrk_ws) gp_syrk_ws_kernel<<<grid, 256, smem_ws, s>>>(g, k0, kw, base, ncol64);
        else gp_syrk_kernel<<<grid, 256, smem_syrk, s>>>(g, k0, kw, base, ncol64);
        BN_LAUNCH_CHECK();
    };
    const size_t smem_potf2_reg = (size_t)(NB * NB + 2 * NB + NB + NB * 17) * sizeof(double);
    for (int kb = 0; kb < T; kb += nblk) {
        const int kend = std::min(kb + nblk, T);  // panels [kb, kend) form one group
        for (int k = kb; k < kend; ++k) {
            const int k0 = k * NB, below = rows_all - k - 1;
            // in-group updates, left-looking (the default): right before panel k is factored, its column block receives
            // ALL earlier panels of the group in ONE pass with K = 128 (k - kb) — every column block of a group is loaded
            // once, flop-weighted mean K 640 (binary schedule below: up to three passes per block, mean K 384)
            if (left_looking && k > kb) {
                launch_syrk(dim3((rows_all - k) * 2, B), kb * NB, (k - kb) * NB, k0, 2);
            }
            if (wait_ev) BN_CUDA(cudaStreamWaitEvent(s, wait_ev[k], 0));
            if (potf2_reg) gp_potf2_reg_kernel<<<B, 256, smem_potf2_reg, s>>>(g, k0);
            else gp_potf2_kernel<<<B, 256, smem_potf2, s>>>(g, k0);
            BN_LAUNCH_CHECK();
            if (below > 0) {
                gp_trsm_kernel<<<dim3(below, B), 512, smem_trsm, s>>>(g, k0);
                BN_LAUNCH_CHECK();
            }
            if (rec_ev) BN_CUDA(cudaEventRecord(rec_ev[k], s));
            // in-group updates, binary schedule (BINEST_GP_LEFT=0): with o panels of the group done, the last
            // w = lowbit(o) of them update the next w column blocks (all rows below) in one K = 128 w pass.  Every column
            // block has then received all earlier panels of its group when its turn comes (the o's that reach it are the
            // prefixes of its binary offset).
            const int o = k - kb + 1;
            if (!left_looking && o < kend - kb) {
                const int w = o & -o, cols = std::min(w, kend - (k + 1));
                launch_syrk(dim3(below * 2 * cols, B), (k + 1 - w) * NB, w * NB, k0 + NB, 2 * cols);
            }
        }
        const int rest = T - kend;
        if (rest > 0) {
            launch_syrk(dim3(rest * (rest + 1) + Tq * 2 * rest, B), kb * NB, (kend - kb) * NB, kend * NB, 0);
        }
    }
}

// fill + sweep of one chunk of matrices.  The chunk is split into `nsplit` sub-batches that run their sweeps on
// separate streams: the diagonal-block kernel is a latency-bound serial column sweep (one CTA per matrix, ~0.25 ms per
// panel, 22 % of a B = 32 sweep when it runs alone), and with the sub-batches drifting out of phase it executes under
// another sub-batch's trailing update instead of leaving the SMs idle.
void gp_factor_chunk(binest_problem &p, const GpBatch &g, const double *theta_dev, int Ps, int b0) {
    const int T = g.Np / NB, Tq = (g.ld - g.Np) / NB, B = g.B;
    const size_t smem_potf2 = (size_t)NB * (NB + 1) * sizeof(double);
    const size_t smem_syrk = (size_t)2 * (KC * LDS_ + KC * LDSH_) * sizeof(double);
    const size_t smem_trsm = ((size_t)NB * LDS_ + 2 * KC * LDS_ + NB) * sizeof(double);
    BN_CUDA(cudaFuncSetAttribute(gp_potf2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_potf2));
    BN_CUDA(cudaFuncSetAttribute(gp_potf2_reg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)((NB * NB + 2 * NB + NB + NB * 17) * sizeof(double))));
    BN_CUDA(cudaFuncSetAttribute(gp_syrk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_syrk));
    BN_CUDA(cudaFuncSetAttribute(gp_syrk_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)((size_t)WS_NS * (WS_KC * LDS_ + WS_KC * LDSH_) * sizeof(double))));
    BN_CUDA(cudaFuncSetAttribute(gp_trsm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_trsm));
    cudaStream_t s = p.stream;
    gp_fill_kernel<<<dim3(T * (T + 1) / 2 + Tq * T, B), 256, 2 * NB * p.gp_dim * sizeof(double), s>>>(
        g, p.gp_x.p, (int)p.gp_dim, p.gp_y.p, theta_dev, Ps, b0);
    BN_LAUNCH_CHECK();
    static const int nblk_env = [] { const char *e = getenv("BINEST_GP_BLOCK"); return e ? atoi(e) : 0; }();
    static const int nsplit_env = [] { const char *e = getenv("BINEST_GP_STREAMS"); return e ? atoi(e) : 0; }();
    const int nblk = std::max(1, std::min(nblk_env > 0 ? nblk_env : kGpBlockPanels, 32));
    int nsplit = std::max(1, std::min(nsplit_env > 0 ? nsplit_env : kGpStreams, 8));
    nsplit = std::min(nsplit, std::max(1, B / 4));  // tiny batches stay whole
    if (nsplit == 1) {
        gp_sweep(g, 0, B, nblk, s, smem_potf2, smem_trsm, smem_syrk);
        return;
    }
    GpStreams &st = g_streams;
    st.ensure(p.device, nsplit - 1);
    // opt-in (BINEST_GP_STAGGER=1): measured 219.6 ms with, 218.2 ms without at N = 4096, B = 256 (r2q) — the sweep is bound
    // by the trailing updates themselves, not by small kernels left exposed
    static const bool stagger = [] { const char *e = getenv("BINEST_GP_STAGGER"); return e && atoi(e) == 1; }();
    if (stagger) st.ensure_panels(nsplit, T);
    BN_CUDA(cudaEventRecord(st.fork, s));
    for (int h = 0; h < nsplit; ++h) {
        const int h0 = (int)((long long)B * h / nsplit), h1 = (int)((long long)B * (h + 1) / nsplit);
        cudaStream_t sh = h == 0 ? s : st.aux[h - 1];
        if (h > 0) BN_CUDA(cudaStreamWaitEvent(sh, st.fork, 0));
        gp_sweep(g, h0, h1 - h0, nblk, sh, smem_potf2, smem_trsm, smem_syrk,
                 (stagger && h > 0) ? st.panel[h - 1].data() : nullptr, (stagger && h + 1 < nsplit) ? st.panel[h].data() : nullptr);
        if (h > 0) {
            BN_CUDA(cudaEventRecord(st.done[h - 1], sh));
            BN_CUDA(cudaStreamWaitEvent(s, st.done[h - 1], 0));
        }
    }
}

}  // namespace

// out[k * stride] = in[k], k < total (NCCL fallback of the batch-sharded gather)
static __global__ void gp_scatter_kernel(const double *__restrict__ in, int total, double *__restrict__ out, int stride) {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < total; k += gridDim.x * blockDim.x) out[(size_t)k * stride] = in[k];
}

// matrices [lo, hi) of the batch: fill + factor + finish, results at out_dev[(w) * out_stride]
static void gp_loglike_range(binest_problem &p, const double *theta_dev, int lo, int hi, int Ps, double *out_dev,
                             int out_stride, bool check_box) {
    const int N = (int)p.gp_n, Np = (N + NB - 1) / NB * NB;
    if (hi <= lo) return;
    GpWorkspace &ws = g_ws;
    ws.ensure(std::min<int>(hi - lo, 512), Np, Np);
    cudaStream_t s = p.stream;
    for (int b0 = lo; b0 < hi; b0 += ws.cap) {
        const int B = std::min(ws.cap, hi - b0);
        GpBatch g{ws.A.p, ws.Y.p, ws.Z.p, ws.L.p, ws.Ld.p, ws.Qd.p, ws.F.p, Np, N, B, Np, 0, nullptr, nullptr};
        gp_factor_chunk(p, g, theta_dev, Ps, b0);
        gp_finish_kernel<<<(B + 127) / 128, 128, 0, s>>>(g, theta_dev, Ps, b0, p.prior, g_logzero, check_box ? 1 : 0,
                                                         out_dev, out_stride);
        BN_LAUNCH_CHECK();
    }
}

// theta_dev SoA [3][Ps]; out_dev[(w) * out_stride].
// Batch-sharded mode (binest_problem_shard_batch): the P parameter vectors are split into `world`
// contiguous slices of cnt = ceil(P / world); this rank fills and factors only its own (the data are replicated), and
// the cnt finished log-likelihoods of every rank are exchanged — pushed into all peers' receive buffers by
// xchg_push_kernel and scattered by xchg_gather_kernel (xchg.cuh; 8 cnt bytes to each peer), or ncclAllGather on the
// fallback.  All ranks end with the same P values, so a GP walk takes identical decisions everywhere.
void gp_loglike_device_strided(binest_problem &p, const double *theta_dev, int P, int Ps, double *out_dev,
                               int out_stride, bool check_box) {
    binest_comm *c = p.comm_batch;
    if (c == nullptr || c->world == 1) {
        gp_loglike_range(p, theta_dev, 0, P, Ps, out_dev, out_stride, check_box);
        return;
    }
    const int W = c->world, cnt = (P + W - 1) / W;
    const int lo = std::min(P, c->rank * cnt), hi = std::min(P, lo + cnt);
    cudaStream_t s = p.stream;
    if (p.sh_send.n < (size_t)cnt) p.sh_send.alloc(cnt);
    BN_CUDA(cudaMemsetAsync(p.sh_send.p, 0, sizeof(double) * cnt, s));
    gp_loglike_range(p, theta_dev, lo, hi, Ps, p.sh_send.p - lo, 1, check_box);
    c->exchanges += 1;
    c->bytes_pushed += (int64_t)8 * cnt * (W - 1);
    if (c->peer) {
        BN_REQUIRE(cnt <= kXchgSlotDoubles, BINEST_ERR_DIMENSION, "batch too large for the sharded exchange buffer");
        xchg_push_kernel<<<std::max(1, std::min(32, (cnt + 255) / 256)), 256, 0, s>>>(c->xd, p.sh_send.p, cnt);
        BN_LAUNCH_CHECK();
        xchg_gather_kernel<<<std::max(1, std::min(32, (P + 255) / 256)), 256, 0, s>>>(c->xd, cnt, P, out_dev, out_stride);
        BN_LAUNCH_CHECK();
    } else {
        if (p.sh_recv.n < (size_t)cnt * W) p.sh_recv.alloc((size_t)cnt * W);
        comm_allgather_f64(*c, p.sh_send.p, p.sh_recv.p, (size_t)cnt, s);
        gp_scatter_kernel<<<std::max(1, std::min(32, (P + 255) / 256)), 256, 0, s>>>(p.sh_recv.p, P, out_dev, out_stride);
        BN_LAUNCH_CHECK();
    }
}

void gp_loglike_device(binest_problem &p, const double *theta_dev, int P, int Ps, double *out_dev, bool check_box) {
    gp_loglike_device_strided(p, theta_dev, P, Ps, out_dev, 1, check_box);
}

// predictFromGaussianProcess (GP:332-422): theta_dev SoA [3][Ps], xs_dev [Q][dim]; mean_dev / sd_dev [P][Q]
void gp_predict_device(binest_problem &p, const double *theta_dev, int P, int Ps, const double *xs_dev, int Q,
                       double *mean_dev, double *sd_dev) {
    const int N = (int)p.gp_n, Np = (N + NB - 1) / NB * NB, Qp = (Q + NB - 1) / NB * NB;
    GpWorkspace &ws = g_ws;
    ws.ensure(std::min<int>(P, 512), Np, Np + Qp);
    cudaStream_t s = p.stream;
    for (int b0 = 0; b0 < P; b0 += ws.cap) {
        const int B = std::min(ws.cap, P - b0);
        GpBatch g{ws.A.p, ws.Y.p, ws.Z.p, ws.L.p, ws.Ld.p, ws.Qd.p, ws.F.p, Np, N, B, Np + Qp, Q, xs_dev, ws.V2.p};
        gp_factor_chunk(p, g, theta_dev, Ps, b0);
        gp_predict_finish_kernel<<<dim3((Q + 127) / 128, B), 128, 0, s>>>(g, theta_dev, Ps, b0, (int64_t)Q, mean_dev, sd_dev);
        BN_LAUNCH_CHECK();
    }
}

}  // namespace binest
