// comm.cu — data-sharded mode (third row): the data rows are split across the GPUs of one box, every
// rank evaluates the same proposals on its own rows, and the per-walker partial sums are exchanged once per
// likelihood launch.  This is the only place the library talks to another GPU; run-sharded jobs
// (parallelNestedSampling, BS:1349-1357) never come here.
//
// NCCL is bound at run time (dlopen of libnccl.so.2 — the copy torch has already loaded when the host is Python, the
// system one otherwise), so libbinest.so itself has no link-time dependency on it and single-GPU users never load it.
// The exchange is an all-gather of 8 P bytes per rank followed by a fixed-order sum on every rank
// (problem.cuh: shard_exchange): latency-bound, NVLink/NVSwitch carries it in one hop.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <functional>
#include <memory>
#include <mutex>

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
        *out = c.release();
    });
}

int binest_comm_free(binest_comm *c) {
    return guard([&] {
        if (!c) return;
        if (c->nccl) {
            cudaSetDevice(c->device);
            cudaDeviceSynchronize();
            nccl().CommDestroy((ncclComm_t)c->nccl);
        }
        delete c;
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

int binest_comm_info(const binest_comm *c, int *rank, int *world) {
    return guard([&] {
        BN_REQUIRE(c, BINEST_ERR_TYPE, "null argument");
        if (rank) *rank = c->rank;
        if (world) *world = c->world;
    });
}

}  // extern "C"
