// Multi-GPU plumbing: one process per GPU, columns split like the reference's
// np.array_split(sorted keys, size)[rank] (src/rt/mass_matrix.py:50-51).
//
// Two exchanges exist on the path, both over NCCL (NVLink 5 / NVSwitch):
//   * halo: packed face traces of the columns on a partition boundary, sent to
//     the ranks whose inflow faces read them (replaces the VecScatter inside
//     PETSc's parallel MatMult);
//   * all-reduce of the k+2 Gram-Schmidt scalars per Arnoldi step (replaces
//     VecMDot/VecNorm's MPI_Allreduce).
// Mass, volume and scattering terms are column-local and need no exchange.
//
// NCCL is resolved with dlopen at run time so that librtb200.so has no link
// dependency on a particular libnccl; inside a torch process this picks up the
// copy torch already loaded.
#include "rtb200_internal.cuh"

#include <dlfcn.h>
#include <nccl.h>

#include <cstring>

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                              cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi g_api;
static bool g_api_tried = false;

static NcclApi* nccl_api() {
    if (g_api.lib) return &g_api;
    if (g_api_tried) return nullptr;
    g_api_tried = true;
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) return nullptr;
#define SYM(field, name)                                         \
    *(void**)(&g_api.field) = dlsym(lib, name);                  \
    if (!g_api.field) return nullptr;
    SYM(GetUniqueId, "ncclGetUniqueId")
    SYM(CommInitRank, "ncclCommInitRank")
    SYM(CommDestroy, "ncclCommDestroy")
    SYM(Send, "ncclSend")
    SYM(Recv, "ncclRecv")
    SYM(AllReduce, "ncclAllReduce")
    SYM(GroupStart, "ncclGroupStart")
    SYM(GroupEnd, "ncclGroupEnd")
    SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    g_api.lib = lib;
    return &g_api;
}

static int nccl_fail(rtb200_ctx* ctx, NcclApi* api, ncclResult_t r, const char* what) {
    return rt_fail(ctx, RTB200_ENCCL, std::string("NCCL error: ") + api->GetErrorString(r) + " in " + what);
}

#define RT_NCCL(ctx, api, call)                                        \
    do {                                                               \
        ncclResult_t r__ = (call);                                     \
        if (r__ != ncclSuccess) return nccl_fail(ctx, api, r__, #call);\
    } while (0)

extern "C" int rtb200_nccl_unique_id(char* id128) {
    if (!id128) return RTB200_EINVAL;
    NcclApi* api = nccl_api();
    if (!api) return RTB200_ENCCL;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    if (api->GetUniqueId(&id) != ncclSuccess) return RTB200_ENCCL;
    std::memcpy(id128, &id, 128);
    return RTB200_OK;
}

extern "C" int rtb200_comm_init(rtb200_ctx* ctx, const char* id128, int rank, int nranks) {
    if (!ctx) return RTB200_EINVAL;
    if (!id128 || nranks < 1 || rank < 0 || rank >= nranks)
        return rt_fail(ctx, RTB200_EINVAL, "comm_init: bad rank / size");
    NcclApi* api = nccl_api();
    if (!api) return rt_fail(ctx, RTB200_ENCCL, "libnccl.so.2 could not be loaded");
    RT_CUDA(ctx, cudaSetDevice(ctx->device));
    rt_comm_destroy(ctx);
    ncclUniqueId id;
    std::memcpy(&id, id128, 128);
    ncclComm_t comm = nullptr;
    RT_NCCL(ctx, api, api->CommInitRank(&comm, nranks, id, rank));
    ctx->nccl = api;
    ctx->nccl_comm = (void*)comm;
    ctx->rank = rank;
    ctx->nranks = nranks;
    return RTB200_OK;
}

void rt_comm_destroy(rtb200_ctx* ctx) {
    if (ctx->nccl && ctx->nccl_comm) ctx->nccl->CommDestroy((ncclComm_t)ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
    ctx->nranks = 1;
    ctx->rank = 0;
}

int rt_allreduce_sum(rtb200_ctx* ctx, double* buf_dev, int count, cudaStream_t st) {
    if (ctx->nranks <= 1) return RTB200_OK;
    NcclApi* api = ctx->nccl;
    if (!api || !ctx->nccl_comm) return rt_fail(ctx, RTB200_ESTATE, "communicator not initialised");
    RT_NCCL(ctx, api, api->AllReduce(buf_dev, buf_dev, (size_t)count, ncclDouble, ncclSum,
                                     (ncclComm_t)ctx->nccl_comm, st));
    return RTB200_OK;
}

// gather the doubles other ranks need into one contiguous send buffer
__global__ void k_pack(const double* __restrict__ x, const int64_t* __restrict__ src,
                       double* __restrict__ out, int64_t n) {
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += step)
        out[e] = x[src[e]];
}

extern "C" int rtb200_halo_exchange(rtb200_ctx* ctx, const double* x_slab_dev, void* stream) {
    if (!ctx) return RTB200_EINVAL;
    if (!ctx->have_mesh) return rt_fail(ctx, RTB200_ESTATE, "no mesh set");
    if (ctx->nranks <= 1 || ctx->peer_rank.empty()) return RTB200_OK;
    NcclApi* api = ctx->nccl;
    if (!api || !ctx->nccl_comm) return rt_fail(ctx, RTB200_ESTATE, "communicator not initialised");
    cudaStream_t st = (cudaStream_t)stream;
    if (ctx->n_send > 0) {
        int blocks = (int)std::min<int64_t>((ctx->n_send + 255) / 256, (int64_t)ctx->sm_count * 2);
        k_pack<<<blocks, 256, 0, st>>>(x_slab_dev, ctx->send_src_dev, ctx->send_dev, ctx->n_send);
        ctx->launches++;
        RT_CUDA(ctx, cudaGetLastError());
    }
    if (ctx->ev_packed && st == ctx->comm_stream) RT_CUDA(ctx, cudaEventRecord(ctx->ev_packed, st));
    ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
    RT_NCCL(ctx, api, api->GroupStart());
    ncclResult_t bad = ncclSuccess;
    for (size_t p = 0; p < ctx->peer_rank.size() && bad == ncclSuccess; ++p) {
        const int64_t ns = ctx->peer_send0[p + 1] - ctx->peer_send0[p];
        const int64_t nr = ctx->peer_recv0[p + 1] - ctx->peer_recv0[p];
        if (ns > 0)
            bad = api->Send(ctx->send_dev + ctx->peer_send0[p], (size_t)ns, ncclDouble,
                            ctx->peer_rank[p], comm, st);
        if (nr > 0 && bad == ncclSuccess)
            bad = api->Recv(ctx->ghost_dev + ctx->peer_recv0[p], (size_t)nr, ncclDouble,
                            ctx->peer_rank[p], comm, st);
    }
    if (bad != ncclSuccess) {
        api->GroupEnd();                       // never leave the group open
        return nccl_fail(ctx, api, bad, "ncclSend / ncclRecv (halo exchange)");
    }
    RT_NCCL(ctx, api, api->GroupEnd());
    return RTB200_OK;
}
