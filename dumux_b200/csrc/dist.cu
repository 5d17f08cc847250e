// dist.cu -- slab-decomposed multi-GPU plumbing: NCCL halo exchange and all-reduce.
//
// Replaces the MPI traffic DuMux gets from dune-istl's overlapping-Schwarz classes:
//   BlockPreconditioner::apply -> copyOwnerToAll            -> halo_exchange(): pack owner planes, ncclSend/ncclRecv
//                                                              grouped with both neighbours, unpack into overlap planes
//   OverlappingSchwarzScalarProduct (MPI_Allreduce sum)     -> allreduce_sum on the packed scalars of the fused dots
//   comm.max(shift) (nonlinear/newtonsolver.hh:1142-1143)   -> allreduce_max
// NCCL is loaded with dlopen so a single-GPU process never needs it; inside a torch process the already-loaded
// libnccl.so.2 is reused.
#include <dlfcn.h>

#include "common.cuh"

namespace dmx {

namespace {
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclInt32 = 2, ncclFloat64 = 8 };
enum { ncclSum = 0, ncclProd = 1, ncclMax = 2, ncclMin = 3 };

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;

bool load_nccl()
{
    if (g_nccl.handle) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* nm : names) {
        h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) return false;
    g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(h, "ncclCommInitRank");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(h, "ncclCommDestroy");
    g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(h, "ncclAllReduce");
    g_nccl.Send = (decltype(g_nccl.Send))dlsym(h, "ncclSend");
    g_nccl.Recv = (decltype(g_nccl.Recv))dlsym(h, "ncclRecv");
    g_nccl.GroupStart = (decltype(g_nccl.GroupStart))dlsym(h, "ncclGroupStart");
    g_nccl.GroupEnd = (decltype(g_nccl.GroupEnd))dlsym(h, "ncclGroupEnd");
    g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(h, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.Send || !g_nccl.Recv || !g_nccl.GroupStart ||
        !g_nccl.GroupEnd)
        return false;
    g_nccl.handle = h;
    return true;
}

#define DMX_NCCL(call)                                                                                              \
    do {                                                                                                            \
        ncclResult_t r__ = (call);                                                                                  \
        if (r__ != 0)                                                                                               \
            return fail(ctx, DMX_ERR_NCCL, std::string(#call) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r__) : "nccl error")); \
    } while (0)

// gather / scatter one plane (index `layer` along the split axis) of a block vector
__global__ void __launch_bounds__(256) plane_pack_kernel(int nx, int ny, int nz, int b, int axis, int layer, const double* v, double* buf)
{
    const int nc[3] = {nx, ny, nz};
    int m = 1;
    for (int d = 0; d < 3; ++d) if (d != axis) m *= nc[d];
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m * b) return;
    const int f = t / b, e = t % b;
    int c[3];
    if (axis == 0) { c[0] = layer; c[1] = f % ny; c[2] = f / ny; }
    else if (axis == 1) { c[0] = f % nx; c[1] = layer; c[2] = f / nx; }
    else { c[0] = f % nx; c[1] = f / nx; c[2] = layer; }
    const size_t I = c[0] + (size_t)nx * (c[1] + (size_t)ny * c[2]);
    buf[t] = v[I * b + e];
}
__global__ void __launch_bounds__(256) plane_unpack_kernel(int nx, int ny, int nz, int b, int axis, int layer, double* v, const double* buf)
{
    const int nc[3] = {nx, ny, nz};
    int m = 1;
    for (int d = 0; d < 3; ++d) if (d != axis) m *= nc[d];
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m * b) return;
    const int f = t / b, e = t % b;
    int c[3];
    if (axis == 0) { c[0] = layer; c[1] = f % ny; c[2] = f / ny; }
    else if (axis == 1) { c[0] = f % nx; c[1] = layer; c[2] = f / nx; }
    else { c[0] = f % nx; c[1] = f / nx; c[2] = layer; }
    const size_t I = c[0] + (size_t)nx * (c[1] + (size_t)ny * c[2]);
    v[I * b + e] = buf[t];
}
} // namespace

int nccl_get_unique_id(void* out)
{
    if (!load_nccl()) return DMX_ERR_NCCL;
    ncclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != 0) return DMX_ERR_NCCL;
    memcpy(out, &id, sizeof(id));
    return 0;
}

int nccl_init(dmx_ctx* ctx, const void* uid)
{
    if (!load_nccl()) return fail(ctx, DMX_ERR_NCCL, "cannot load libnccl.so.2");
    if (!uid) return fail(ctx, DMX_ERR_USAGE, "distributed context needs an ncclUniqueId");
    ncclUniqueId id;
    memcpy(&id, uid, sizeof(id));
    ncclComm_t comm = nullptr;
    DMX_NCCL(g_nccl.CommInitRank(&comm, ctx->nranks, id, ctx->rank));
    ctx->nccl_comm = comm;
    return 0;
}
int nccl_destroy(dmx_ctx* ctx)
{
    if (ctx->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
    return 0;
}

// copyOwnerToAll for a slab decomposition with overlap 1: my first/last OWNED plane goes to the neighbour's
// overlap plane; my overlap planes are overwritten with the neighbours' owned planes.
int halo_exchange(dmx_ctx* ctx, double* v)
{
    if (ctx->nranks == 1) return 0;
    ProfScope ps(ctx, DMX_K_HALO);
    ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
    const int sa = ctx->split_axis;
    const int nx = ctx->nc[0], ny = ctx->nc[1], nz = ctx->nc[2], b = ctx->b;
    const int m = (ctx->n / ctx->nc[sa]) * b;
    const int grid = (m + 255) / 256;
    const bool hasLo = ctx->off[sa] > 0;
    const bool hasHi = ctx->off[sa] + ctx->nc[sa] < ctx->gcells[sa];
    if (hasLo) { plane_pack_kernel<<<grid, 256, 0, ctx->stream>>>(nx, ny, nz, b, sa, ctx->own_begin, v, ctx->d_send_lo); DMX_CHECK_LAUNCH(); }
    if (hasHi) { plane_pack_kernel<<<grid, 256, 0, ctx->stream>>>(nx, ny, nz, b, sa, ctx->own_end - 1, v, ctx->d_send_hi); DMX_CHECK_LAUNCH(); }
    DMX_NCCL(g_nccl.GroupStart());
    if (hasLo) {
        DMX_NCCL(g_nccl.Send(ctx->d_send_lo, m, ncclFloat64, ctx->rank - 1, comm, ctx->stream));
        DMX_NCCL(g_nccl.Recv(ctx->d_recv_lo, m, ncclFloat64, ctx->rank - 1, comm, ctx->stream));
    }
    if (hasHi) {
        DMX_NCCL(g_nccl.Send(ctx->d_send_hi, m, ncclFloat64, ctx->rank + 1, comm, ctx->stream));
        DMX_NCCL(g_nccl.Recv(ctx->d_recv_hi, m, ncclFloat64, ctx->rank + 1, comm, ctx->stream));
    }
    DMX_NCCL(g_nccl.GroupEnd());
    if (hasLo) { plane_unpack_kernel<<<grid, 256, 0, ctx->stream>>>(nx, ny, nz, b, sa, ctx->own_begin - 1, v, ctx->d_recv_lo); DMX_CHECK_LAUNCH(); }
    if (hasHi) { plane_unpack_kernel<<<grid, 256, 0, ctx->stream>>>(nx, ny, nz, b, sa, ctx->own_end, v, ctx->d_recv_hi); DMX_CHECK_LAUNCH(); }
    return 0;
}

int allreduce_sum(dmx_ctx* ctx, double* d_buf, int count)
{
    if (ctx->nranks == 1) return 0;
    DMX_NCCL(g_nccl.AllReduce(d_buf, d_buf, count, ncclFloat64, ncclSum, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    return 0;
}
int allreduce_max(dmx_ctx* ctx, double* d_buf, int count)
{
    if (ctx->nranks == 1) return 0;
    DMX_NCCL(g_nccl.AllReduce(d_buf, d_buf, count, ncclFloat64, ncclMax, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    return 0;
}
int allreduce_min_int(dmx_ctx* ctx, int* d_buf, int count)
{
    if (ctx->nranks == 1) return 0;
    DMX_NCCL(g_nccl.AllReduce(d_buf, d_buf, count, ncclInt32, ncclMin, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    return 0;
}

} // namespace dmx
