// xchg.cuh — in-kernel all-gather of small fp64 vectors between the GPUs of one box, over peer-mapped memory
// (CUDA IPC over NVLink 5 / NVSwitch), used by the two sharded modes:
//   data-sharded   every rank reduces its rows for the same P proposals; the P per-rank sums are exchanged after
//                  every likelihood launch (8 P bytes to each of the W - 1 peers per step);
//   batch-sharded  (GP) every rank factors its slice of the theta batch; the finished log-likelihoods are exchanged.
// Both are latency-bound exchanges of a few KiB.  A host-issued ncclAllGather costs a collective launch per walk
// step and cannot live inside a CUDA graph node of ours; here the producing kernel itself PUSHES its values into
// every peer's receive buffer with plain P2P stores and raises a per-(peer, rank) flag, and the consuming kernel
// polls its OWN flags (local memory) — no host involvement, no extra launch, graph-capturable.
//
// Protocol, exchange number t = 1, 2, ... (st->step, advanced by the producer on every rank in lock step):
//   producer kernel   every thread stores its values to slot [t & 1][rank] of EVERY rank's buffer (its own included),
//                     fence.sys; the last CTA to finish (device-scope ticket) writes flag[rank] = t on every rank
//                     (st.release.sys) and advances st->step;
//   consumer kernel   (next kernel on the stream) one thread per CTA polls flag[r] >= t for all r (ld.acquire.sys,
//                     bounded: ~4 s, then an abort flag is raised and the host reports BINEST_ERR_CUDA), reads slot
//                     [t & 1][r] of its own buffer.
// Two slots suffice: a rank can only write exchange t + 2 after it has consumed t + 1, which needs every peer's flag
// t + 1, which a peer raises after it has finished consuming t (stream order on that peer).
// Every rank adds the W values in rank order, so all ranks take bit-identical decisions.
#pragma once
#include "common.cuh"

namespace binest {

constexpr int kXchgMaxWorld = 8;
constexpr int kXchgSlotDoubles = 1 << 16;  // capacity per rank and parity (512 KiB): P <= 65536 walkers

struct XchgState {
    unsigned step;      // exchanges completed by this rank's producers
    unsigned ticket;    // CTAs of the running producer that have finished
    unsigned abort;     // a consumer timed out
    unsigned pad_;
};

struct XchgDev {
    double *buf[kXchgMaxWorld];      // buf[r]: rank r's receive buffer [2][world][kXchgSlotDoubles] (peer-mapped)
    unsigned *flag[kXchgMaxWorld];   // flag[r]: rank r's flags [world], each on its own 128-byte line
    XchgState *st;                   // local
    int rank, world;
    __host__ __device__ size_t slot(unsigned t, int r) const {
        return ((size_t)(t & 1u) * world + r) * kXchgSlotDoubles;
    }
};
constexpr int kXchgFlagStride = 32;  // unsigned per flag line

__device__ __forceinline__ void st_release_sys_u32(unsigned *p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// producer side, called by ALL threads of the grid after they have stored their values with xchg_store()
__device__ __forceinline__ void xchg_store(const XchgDev &x, unsigned t, size_t idx, double v) {
    const size_t o = x.slot(t, x.rank) + idx;
    for (int r = 0; r < x.world; ++r) x.buf[r][o] = v;
}
__device__ __forceinline__ void xchg_publish(const XchgDev &x, unsigned t) {
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned total = gridDim.x * gridDim.y * gridDim.z;
        if (atomicAdd(&x.st->ticket, 1u) == total - 1u) {  // every other CTA's stores are fenced before its ticket
            __threadfence_system();
            x.st->ticket = 0u;
            x.st->step = t;
            for (int r = 0; r < x.world; ++r) st_release_sys_u32(x.flag[r] + (size_t)x.rank * kXchgFlagStride, t);
        }
    }
}

// consumer side: one thread polls this rank's own flags; false after a time-out / abort
__device__ __forceinline__ bool xchg_wait(const XchgDev &x, unsigned t) {
    const unsigned *fl = x.flag[x.rank];
    const long long t0 = clock64();
    for (int r = 0; r < x.world; ++r) {
        unsigned ns = 32;
        while (ld_acquire_sys_u32(fl + (size_t)r * kXchgFlagStride) < t) {
            if (*(volatile unsigned *)&x.st->abort) return false;
            if (clock64() - t0 > 8000000000LL) {  // ~4 s at 1.9 GHz
                atomicExch(&x.st->abort, 1u);
                return false;
            }
            __nanosleep(ns);
            if (ns < 1024) ns <<= 1;
        }
    }
    return true;
}

// standalone producer: push send[0..count) (batch-sharded GP: this rank's finished log-likelihoods)
static __global__ void xchg_push_kernel(const XchgDev x, const double *__restrict__ send, int count) {
    const unsigned t = x.st->step + 1u;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) xchg_store(x, t, i, send[i]);
    xchg_publish(x, t);
}

// standalone consumer: out[(r * count + i) * stride] = value i of rank r, for r * count + i < total
static __global__ void xchg_gather_kernel(const XchgDev x, int count, int total, double *__restrict__ out, int stride) {
    __shared__ int ok;
    const unsigned t = x.st->step;
    if (threadIdx.x == 0) ok = xchg_wait(x, t) ? 1 : 0;
    __syncthreads();
    if (!ok) return;
    const double *buf = x.buf[x.rank];
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < total; k += gridDim.x * blockDim.x) {
        const int r = k / count, i = k - r * count;
        out[(size_t)k * stride] = __ldcg(buf + x.slot(t, r) + i);
    }
}

}  // namespace binest
