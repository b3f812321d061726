// Multi-GPU plumbing of the mapping path: one NCCL communicator per context (one process per GPU).
//
// M = sum_i w_i h_i h_i^T, j = sum_i w_i V_i h_i and H0 are sums over independent visibilities
// (frank/statistical_models.py:200-218 already accumulates them chunk by chunk), so every rank maps its own slice and
// the partial results are combined by ONE grouped all-reduce issued on the library's stream right behind the scaling
// kernel -- no host synchronisation between the Gram kernel and the collective.  The range of q (for
// _check_uv_range, statistical_models.py:512-535) and the call's status bits travel in a second, 4-double max-reduce,
// so that every rank returns the same status.
//
// NCCL is loaded with dlopen("libnccl.so.2") at fb_comm_init: a process that already carries NCCL (PyTorch bundles one)
// shares that copy, a plain C host gets the system library, and a single-GPU user never needs it.
#include "fb_common.cuh"

#include <dlfcn.h>
#include <nccl.h>

namespace {

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId *);
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    const char *(*GetErrorString)(ncclResult_t);
    void *handle = nullptr;
};

NcclApi g_nccl;

int load_nccl(std::string &err)
{
    if (g_nccl.handle) return 0;
    void *h = nullptr;
    for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
        h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) { err = std::string("fb_comm: cannot load NCCL: ") + dlerror(); return -70; }
#define FB_SYM(field, sym)                                                           \
    g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(h, sym));          \
    if (!g_nccl.field) { err = std::string("fb_comm: NCCL symbol missing: ") + sym; return -71; }
    FB_SYM(GetUniqueId, "ncclGetUniqueId")
    FB_SYM(CommInitRank, "ncclCommInitRank")
    FB_SYM(CommDestroy, "ncclCommDestroy")
    FB_SYM(AllReduce, "ncclAllReduce")
    FB_SYM(AllGather, "ncclAllGather")
    FB_SYM(GroupStart, "ncclGroupStart")
    FB_SYM(GroupEnd, "ncclGroupEnd")
    FB_SYM(GetErrorString, "ncclGetErrorString")
#undef FB_SYM
    g_nccl.handle = h;
    return 0;
}

#define FB_NCCL(call)                                                                          \
    do {                                                                                       \
        ncclResult_t r__ = (call);                                                             \
        if (r__ != ncclSuccess) {                                                              \
            ctx->err = std::string(#call) + " failed: " + g_nccl.GetErrorString(r__);          \
            return -72;                                                                        \
        }                                                                                      \
    } while (0)

// result = {H0, qmin, qmax, status} -> pack = {-qmin, qmax, range flag, table flag}: one max-reduce combines all four
__global__ void k_pack_result(const double *__restrict__ result, double *__restrict__ pack)
{
    const int st = (int)result[3];
    pack[0] = -result[1]; pack[1] = result[2];
    pack[2] = (st & FB_ST_QRANGE) ? 1.0 : 0.0;
    pack[3] = (st & FB_ST_TABLE) ? 1.0 : 0.0;
}

__global__ void k_unpack_result(const double *__restrict__ pack, double *__restrict__ result, double *__restrict__ dev_H0)
{
    result[1] = -pack[0]; result[2] = pack[1];
    const int st = (pack[2] > 0.0 ? FB_ST_QRANGE : 0) | (pack[3] > 0.0 ? FB_ST_TABLE : 0);
    result[3] = (double)st;
    if (dev_H0 && !st) dev_H0[0] = result[0];
}

int ensure_pack(fb_ctx *ctx, size_t doubles)
{
    if (doubles > ctx->pack_cap) {
        FB_CUDA(cudaDeviceSynchronize());
        if (ctx->d_pack) FB_CUDA(cudaFree(ctx->d_pack));
        ctx->d_pack = nullptr;
        FB_CUDA(cudaMalloc(&ctx->d_pack, sizeof(double) * doubles));
        ctx->pack_cap = doubles;
    }
    return 0;
}

}  // namespace

// Sum (M, j, H0) and max-reduce (range of q, status) over the communicator, in place, on stream `st`.
int fb_comm_allreduce_map(fb_ctx *ctx, cudaStream_t st, int nchan, double *dev_M, double *dev_j, double *dev_H0)
{
    ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
    const size_t N = ctx->N;
    int rc = ensure_pack(ctx, 8);
    if (rc) return rc;
    k_pack_result<<<1, 1, 0, st>>>(ctx->d_result, ctx->d_pack);
    FB_CUDA(cudaGetLastError());
    FB_NCCL(g_nccl.GroupStart());
    FB_NCCL(g_nccl.AllReduce(dev_M, dev_M, (size_t)nchan * N * N, ncclDouble, ncclSum, comm, st));
    FB_NCCL(g_nccl.AllReduce(dev_j, dev_j, (size_t)nchan * N, ncclDouble, ncclSum, comm, st));
    FB_NCCL(g_nccl.AllReduce(ctx->d_result, ctx->d_result, 1, ncclDouble, ncclSum, comm, st));
    FB_NCCL(g_nccl.AllReduce(ctx->d_pack, ctx->d_pack, 4, ncclDouble, ncclMax, comm, st));
    FB_NCCL(g_nccl.GroupEnd());
    k_unpack_result<<<1, 1, 0, st>>>(ctx->d_pack, ctx->d_result, dev_H0);
    FB_CUDA(cudaGetLastError());
    return 0;
}

extern "C" {

int fb_comm_unique_id(void *out128)
{
    std::string err;
    if (!out128 || load_nccl(err)) return -70;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    return g_nccl.GetUniqueId((ncclUniqueId *)out128) == ncclSuccess ? 0 : -72;
}

int fb_comm_init(fb_ctx *ctx, int nranks, int rank, const void *id128)
{
    if (!ctx) return -1;
    if (nranks < 1 || rank < 0 || rank >= nranks || !id128) FB_FAIL(-73, "fb_comm_init: bad arguments");
    FB_CUDA(cudaSetDevice(ctx->device));
    int rc = load_nccl(ctx->err);
    if (rc) return rc;
    fb_comm_destroy(ctx);
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclComm_t comm = nullptr;
    FB_NCCL(g_nccl.CommInitRank(&comm, nranks, id, rank));
    ctx->nccl_comm = comm;
    ctx->comm_rank = rank;
    ctx->comm_size = nranks;
    return 0;
}

int fb_comm_destroy(fb_ctx *ctx)
{
    if (!ctx) return 0;
    if (ctx->nccl_comm && g_nccl.handle) {
        cudaSetDevice(ctx->device);
        cudaDeviceSynchronize();
        g_nccl.CommDestroy((ncclComm_t)ctx->nccl_comm);
    }
    ctx->nccl_comm = nullptr;
    ctx->comm_rank = 0;
    ctx->comm_size = 1;
    return 0;
}

int fb_comm_info(fb_ctx *ctx, int *rank, int *nranks)
{
    if (!ctx) return -1;
    if (rank) *rank = ctx->comm_rank;
    if (nranks) *nranks = ctx->comm_size;
    return ctx->nccl_comm ? 1 : 0;
}

// All-gather `count` doubles per rank (host buffers; staged through the device): recv holds nranks * count doubles in
// rank order.  Used to collect the results of a hyper-parameter sweep sharded by grid point.
int fb_comm_allgather(fb_ctx *ctx, const double *host_send, int64_t count, double *host_recv)
{
    if (!ctx) return -1;
    if (!host_send || !host_recv || count < 1) FB_FAIL(-73, "fb_comm_allgather: bad arguments");
    FB_CUDA(cudaSetDevice(ctx->device));
    if (!ctx->nccl_comm) {                                  // single rank: a copy
        memcpy(host_recv, host_send, sizeof(double) * count);
        return 0;
    }
    const size_t R = ctx->comm_size;
    int rc = ensure_pack(ctx, (size_t)count * (R + 1));
    if (rc) return rc;
    double *d_send = ctx->d_pack, *d_recv = ctx->d_pack + count;
    FB_CUDA(cudaMemcpyAsync(d_send, host_send, sizeof(double) * count, cudaMemcpyHostToDevice, ctx->stream));
    FB_NCCL(g_nccl.AllGather(d_send, d_recv, (size_t)count, ncclDouble, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    FB_CUDA(cudaMemcpyAsync(host_recv, d_recv, sizeof(double) * count * R, cudaMemcpyDeviceToHost, ctx->stream));
    FB_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// Sum-all-reduce of a device buffer on the library stream (asynchronous): the packed mapping outputs of callers that
// keep M, j, H0 in one allocation.
int fb_comm_allreduce_sum_dev(fb_ctx *ctx, double *dev_buf, int64_t count)
{
    if (!ctx) return -1;
    if (!dev_buf || count < 1) FB_FAIL(-73, "fb_comm_allreduce_sum_dev: bad arguments");
    if (!ctx->nccl_comm) return 0;
    FB_CUDA(cudaSetDevice(ctx->device));
    FB_NCCL(g_nccl.AllReduce(dev_buf, dev_buf, (size_t)count, ncclDouble, ncclSum, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    return 0;
}

}  // extern "C"
