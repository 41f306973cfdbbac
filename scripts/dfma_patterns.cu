// Micro-benchmarks of the fp64 pipe on sm_100a: how operand patterns and occupancy change DFMA throughput.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/dfma_patterns scripts/dfma_patterns.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int PATTERN, int CHAINS>
__global__ void __launch_bounds__(256) k(double *out, const double *in, int iters) {
    double a[CHAINS], c[CHAINS], e[CHAINS];
    for (int i = 0; i < CHAINS; ++i) { a[i] = in[i] + threadIdx.x; c[i] = in[8 + i]; e[i] = in[16 + i]; }
    double x = in[31], y = in[30];
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        if (PATTERN == 0) {  // a = a*m + b, m,b shared
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int i = 0; i < CHAINS; ++i) a[i] = fma(a[i], x, y);
        } else if (PATTERN == 1) {  // Horner-like: a_i = a_i * x + c_i (x shared, c_i distinct)
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int i = 0; i < CHAINS; ++i) a[i] = fma(a[i], x, c[i]);
        } else if (PATTERN == 2) {  // three distinct registers per instruction
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int i = 0; i < CHAINS; ++i) a[i] = fma(a[i], e[i], c[i]);
        } else if (PATTERN == 3) {  // poly3 datum: 3 Horner + sub + square-accumulate per chain (5 ops)
#pragma unroll
            for (int i = 0; i < CHAINS; ++i) {
                double t = fma(c[i], x, e[i]);
                t = fma(t, x, c[(i + 1) % CHAINS]);
                t = fma(t, x, e[(i + 1) % CHAINS]);
                const double r = y - t;
                a[i] = fma(r, r, a[i]);
            }
            x += 1e-9;
        }
    }
    double s = 0;
    for (int i = 0; i < CHAINS; ++i) s += a[i];
    if (s == 1.2345) out[0] = s;
}

template <int PATTERN, int CHAINS>
void run(const char *name, int ctas_per_sm, int threads, double *out, double *in, int ops_per_iter) {
    const int iters = 4096;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        k<PATTERN, CHAINS><<<148 * ctas_per_sm, threads>>>(out, in, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    const double ops = (double)ops_per_iter * iters * 148.0 * ctas_per_sm * threads;
    printf("%-28s chains=%d warps/SM=%2d : %7.2f Gop/s-lane => %5.1f %% of 148*64*1.965GHz\n", name, CHAINS,
           ctas_per_sm * threads / 32, ops / best / 1e6, 100.0 * ops / (best * 1e-3) / (148.0 * 64 * 1.965e9));
}

int main() {
    double *out, *in;
    cudaMalloc(&out, 8); cudaMalloc(&in, 32 * 8);
    double h[32]; for (int i = 0; i < 32; ++i) h[i] = 1.0 + 1e-7 * i;
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    for (int w : {1, 2, 3, 4, 6, 8}) {
        run<0, 8>("const m,b", w, 256, out, in, 32);
        run<1, 8>("horner x shared", w, 256, out, in, 32);
        run<2, 8>("3 distinct", w, 256, out, in, 32);
        run<3, 8>("poly3 datum (5 ops)", w, 256, out, in, 40);
        run<3, 4>("poly3 datum (5 ops)", w, 256, out, in, 20);
    }
    return 0;
}
