// comm.cu — the sharded modes: data-sharded (rows split across the GPUs of one box, every rank evaluates
// the same proposals on its own rows, the per-walker partial sums are exchanged once per likelihood launch) and
// batch-sharded (GP: the theta batch is split, the finished log-likelihoods are exchanged).  This is the only place the
// library talks to another GPU; run-sharded jobs (parallelNestedSampling, BS:1349-1357) never come here.
//
// The per-step exchange runs INSIDE our kernels over peer-mapped memory (xchg.cuh): binest_comm_create allocates a
// receive buffer and a flag line per rank with cudaMalloc, passes their CUDA IPC handles around, and every rank maps
// every peer's buffer (cudaIpcOpenMemHandle -> P2P over NVLink 5 / NVSwitch).  NCCL carries the set-up traffic (IPC
// handles, data constants) and remains as the fallback exchange (BINEST_XCHG=nccl, or when peer mapping fails): bound
// at run time (dlopen of libnccl.so.2 — the copy torch has already loaded when the host is Python, the system one
// otherwise), so libbinest.so has no link-time dependency on it and single-GPU users never load it.
#include <dlfcn.h>
#include <nccl.h>

#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <vector>

#include "problem.cuh"

namespace binest {
int guard(const std::function<void()> &f);

namespace {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi &nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
            api.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) return;
        auto sym = [&](const char *n) { return dlsym(api.handle, n); };
        api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
        api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
        api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
        api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    });
    BN_REQUIRE(api.handle && api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather,
               BINEST_ERR_FUNCTION, "data-sharded mode needs NCCL (libnccl.so.2 not found)");
    return api;
}

void nccl_check(ncclResult_t r, const char *what) {
    if (r == ncclSuccess) return;
    const char *msg = nccl().GetErrorString ? nccl().GetErrorString(r) : "?";
    throw Error(BINEST_ERR_CUDA, std::string(what) + ": " + msg);
}

}  // namespace

// Peer-mapped exchange buffers (xchg.cuh).  Collective.  Any failure on any rank (no P2P between the devices, IPC not
// permitted in the container, BINEST_XCHG=nccl) leaves EVERY rank on the NCCL fallback: the outcome is agreed with one
// more all-gather, so the ranks never disagree about the path.
void setup_peer_exchange(binest_comm &c) {
    struct Blob { cudaIpcMemHandle_t buf, flag; int ok; int pad[3]; };
    static_assert(sizeof(Blob) % 8 == 0, "blob travels as doubles");
    const int W = c.world;
    BN_REQUIRE(W <= kXchgMaxWorld, BINEST_ERR_DIMENSION, "sharded modes support up to 8 ranks (one box)");
    const char *e = std::getenv("BINEST_XCHG");
    bool want = !(e && std::strcmp(e, "nccl") == 0);
    Blob mine;
    std::memset(&mine, 0, sizeof(mine));
    const size_t buf_bytes = sizeof(double) * 2 * (size_t)W * kXchgSlotDoubles;
    const size_t flag_bytes = sizeof(unsigned) * (size_t)W * kXchgFlagStride;
    if (want) {
        want = cudaMalloc((void **)&c.xbuf_local, buf_bytes) == cudaSuccess &&
               cudaMalloc((void **)&c.xflag_local, flag_bytes) == cudaSuccess &&
               cudaMalloc((void **)&c.xst, sizeof(XchgState)) == cudaSuccess &&
               cudaMalloc((void **)&c.xd_dev, sizeof(XchgDev)) == cudaSuccess &&
               cudaMemset(c.xbuf_local, 0, buf_bytes) == cudaSuccess &&
               cudaMemset(c.xflag_local, 0, flag_bytes) == cudaSuccess &&
               cudaMemset(c.xst, 0, sizeof(XchgState)) == cudaSuccess &&
               cudaIpcGetMemHandle(&mine.buf, c.xbuf_local) == cudaSuccess &&
               cudaIpcGetMemHandle(&mine.flag, c.xflag_local) == cudaSuccess;
        cudaGetLastError();
    }
    mine.ok = want ? 1 : 0;
    auto gather = [&](const Blob &in, std::vector<Blob> &all) {
        constexpr size_t ND = sizeof(Blob) / sizeof(double);
        DevBuf<double> send(ND), recv(ND * W);
        BN_CUDA(cudaMemcpy(send.p, &in, sizeof(Blob), cudaMemcpyHostToDevice));
        nccl_check(nccl().AllGather(send.p, recv.p, ND, ncclFloat64, (ncclComm_t)c.nccl, (cudaStream_t)0), "ncclAllGather");
        all.resize(W);
        BN_CUDA(cudaMemcpy(all.data(), recv.p, sizeof(Blob) * W, cudaMemcpyDeviceToHost));
    };
    std::vector<Blob> all;
    gather(mine, all);
    bool ok = true;
    for (int r = 0; r < W; ++r) ok = ok && all[r].ok;
    XchgDev xd;
    std::memset(&xd, 0, sizeof(xd));
    xd.rank = c.rank; xd.world = W; xd.st = c.xst;
    for (int r = 0; r < W && ok; ++r) {
        if (r == c.rank) { xd.buf[r] = c.xbuf_local; xd.flag[r] = c.xflag_local; continue; }
        void *pb = nullptr, *pf = nullptr;
        ok = cudaIpcOpenMemHandle(&pb, all[r].buf, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess &&
             cudaIpcOpenMemHandle(&pf, all[r].flag, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
        cudaGetLastError();
        xd.buf[r] = (double *)pb; xd.flag[r] = (unsigned *)pf;
    }
    // second round: every rank reports whether ITS mappings succeeded; the peer path is used only if all did.  The
    // collective also orders every rank's memset before anybody's first push.
    Blob st2;
    std::memset(&st2, 0, sizeof(st2));
    st2.ok = ok ? 1 : 0;
    gather(st2, all);
    bool all_ok = true;
    for (int r = 0; r < W; ++r) all_ok = all_ok && all[r].ok;
    c.xd = xd;
    c.peer = all_ok;
    if (all_ok) {
        BN_CUDA(cudaMemcpy(c.xd_dev, &xd, sizeof(xd), cudaMemcpyHostToDevice));
        BN_CUDA(cudaHostAlloc((void **)&c.h_abort, sizeof(unsigned), cudaHostAllocDefault));
        *c.h_abort = 0;
    }
}

void comm_allgather_f64(binest_comm &c, const double *send, double *recv, size_t count, cudaStream_t s) {
    nccl_check(nccl().AllGather(send, recv, count, ncclFloat64, (ncclComm_t)c.nccl, s), "ncclAllGather");
    count_launch();
}

}  // namespace binest

using namespace binest;

extern "C" {

int binest_comm_unique_id(uint8_t *id /*[BINEST_COMM_ID_BYTES]*/) {
    return guard([&] {
        BN_REQUIRE(id, BINEST_ERR_TYPE, "null argument");
        static_assert(sizeof(ncclUniqueId) <= BINEST_COMM_ID_BYTES, "ncclUniqueId does not fit the ABI buffer");
        ncclUniqueId u;
        nccl_check(nccl().GetUniqueId(&u), "ncclGetUniqueId");
        std::memset(id, 0, BINEST_COMM_ID_BYTES);
        std::memcpy(id, &u, sizeof(u));
    });
}

int binest_comm_create(int rank, int world, const uint8_t *id, binest_comm **out) {
    return guard([&] {
        BN_REQUIRE(id && out, BINEST_ERR_TYPE, "null argument");
        BN_REQUIRE(world >= 1 && rank >= 0 && rank < world, BINEST_ERR_DIMENSION, "0 <= rank < world");
        std::unique_ptr<binest_comm> c(new binest_comm());
        c->rank = rank;
        c->world = world;
        BN_CUDA(cudaGetDevice(&c->device));
        ncclUniqueId u;
        std::memcpy(&u, id, sizeof(u));
        ncclComm_t comm = nullptr;
        nccl_check(nccl().CommInitRank(&comm, world, u, rank), "ncclCommInitRank");
        c->nccl = comm;
        setup_peer_exchange(*c);
        *out = c.release();
    });
}

int binest_comm_free(binest_comm *c) {
    return guard([&] {
        if (!c) return;
        cudaSetDevice(c->device);
        cudaDeviceSynchronize();
        if (c->nccl && c->peer) {  // nobody unmaps while a peer may still push: one last collective as a barrier
            DevBuf<double> a(1), b((size_t)c->world);
            a.zero();
            nccl().AllGather(a.p, b.p, 1, ncclFloat64, (ncclComm_t)c->nccl, (cudaStream_t)0);
            cudaDeviceSynchronize();
        }
        for (int r = 0; r < c->world && c->peer; ++r) {
            if (r == c->rank) continue;
            if (c->xd.buf[r]) cudaIpcCloseMemHandle(c->xd.buf[r]);
            if (c->xd.flag[r]) cudaIpcCloseMemHandle(c->xd.flag[r]);
        }
        if (c->xbuf_local) cudaFree(c->xbuf_local);
        if (c->xflag_local) cudaFree(c->xflag_local);
        if (c->xst) cudaFree(c->xst);
        if (c->xd_dev) cudaFree(c->xd_dev);
        if (c->h_abort) cudaFreeHost(c->h_abort);
        if (c->nccl) nccl().CommDestroy((ncclComm_t)c->nccl);
        delete c;
    });
}

// traffic counters of the sharded exchange (bench.py): exchanges issued, payload bytes this rank pushed to its peers,
// and whether the in-kernel peer path (1) or the NCCL fallback (0) is in use
int binest_comm_stats(const binest_comm *c, int64_t *exchanges, int64_t *bytes_pushed, int *peer_path) {
    return guard([&] {
        BN_REQUIRE(c, BINEST_ERR_TYPE, "null argument");
        if (exchanges) *exchanges = c->exchanges;
        if (bytes_pushed) *bytes_pushed = c->bytes_pushed;
        if (peer_path) *peer_path = c->peer ? 1 : 0;
    });
}

// Collective over the communicator: declares that `p` holds one shard of the data rows.  The operator epilogues
// (rows * log-normalisation, the GBM constant) need the totals over all shards: exchanged here, once.
int binest_problem_shard(binest_problem *p, binest_comm *c) {
    return guard([&] {
        BN_REQUIRE(p && c, BINEST_ERR_TYPE, "null argument");
        BN_REQUIRE(p->op != BINEST_OP_GP_SE, BINEST_ERR_FUNCTION,
                   "the GP operator does not shard by rows (replicas only); shard the theta batch instead");
        BN_REQUIRE(p->device == c->device, BINEST_ERR_CUDA, "problem and communicator live on different devices");
        BN_CUDA(cudaSetDevice(p->device));
        constexpr int NV = 8;  // rows, additive constant, 6 data moments
        DevBuf<double> send(NV), recv((size_t)NV * c->world);
        const double h[NV] = {(double)p->rows, p->cst.c, p->cst.m[0], p->cst.m[1], p->cst.m[2], p->cst.m[3], p->cst.m[4],
                              p->cst.m[5]};
        BN_CUDA(cudaMemcpyAsync(send.p, h, sizeof(h), cudaMemcpyHostToDevice, p->stream));
        comm_allgather_f64(*c, send.p, recv.p, NV, p->stream);
        std::vector<double> all((size_t)NV * c->world);
        BN_CUDA(cudaMemcpyAsync(all.data(), recv.p, sizeof(double) * all.size(), cudaMemcpyDeviceToHost, p->stream));
        BN_CUDA(cudaStreamSynchronize(p->stream));
        long double tot[NV] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int r = 0; r < c->world; ++r)
            for (int v = 0; v < NV; ++v) tot[v] += all[(size_t)NV * r + v];
        p->rows_total = (double)tot[0];
        // Only the row count and the additive constant are totals.  The polynomial operator's moments and pivots stay
        // per shard: every rank resolves its own Sum e^2 with them (OP::local) before the exchange.
        p->cst_total = p->cst;
        p->cst_total.c = (double)tot[1];
        p->comm = c;
    });
}

// Collective: from now on the GP problem `p` (data replicated on every rank) evaluates every theta batch in slices,
// one per rank (gp.cu: gp_loglike_device_strided).  binest_loglike / binest_run_* become collectives: all ranks must
// make the same calls with the same theta / options / seed.
int binest_problem_shard_batch(binest_problem *p, binest_comm *c) {
    return guard([&] {
        BN_REQUIRE(p && c, BINEST_ERR_TYPE, "null argument");
        BN_REQUIRE(p->op == BINEST_OP_GP_SE, BINEST_ERR_FUNCTION,
                   "batch sharding is the GP operator's mode; the streaming operators shard by rows (binest_problem_shard)");
        BN_REQUIRE(p->device == c->device, BINEST_ERR_CUDA, "problem and communicator live on different devices");
        BN_REQUIRE(!p->comm, BINEST_ERR_FUNCTION, "problem is already data-sharded");
        p->comm_batch = c;
    });
}

int binest_comm_info(const binest_comm *c, int *rank, int *world) {
    return guard([&] {
        BN_REQUIRE(c, BINEST_ERR_TYPE, "null argument");
        if (rank) *rank = c->rank;
        if (world) *world = c->world;
    });
}

}  // extern "C"
