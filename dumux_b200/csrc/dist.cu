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
    ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
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
    g_nccl.AllGather = (decltype(g_nccl.AllGather))dlsym(h, "ncclAllGather");
    g_nccl.Send = (decltype(g_nccl.Send))dlsym(h, "ncclSend");
    g_nccl.Recv = (decltype(g_nccl.Recv))dlsym(h, "ncclRecv");
    g_nccl.GroupStart = (decltype(g_nccl.GroupStart))dlsym(h, "ncclGroupStart");
    g_nccl.GroupEnd = (decltype(g_nccl.GroupEnd))dlsym(h, "ncclGroupEnd");
    g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(h, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.AllGather || !g_nccl.Send || !g_nccl.Recv || !g_nccl.GroupStart ||
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

// gather / scatter the copyOwnerToAll regions of all neighbours in one launch: region q = box lo[q] + [0, size[q]) of the local
// grid, packed x fastest at buf + off[q]
struct HaloRegions {
    int nreg;
    int lo[26][3], size[26][3];
    long long off[27];         // in doubles; off[nreg] = total
};
template <bool PACK>
__global__ void __launch_bounds__(256) halo_pack_kernel(HaloRegions R, int nx, int ny, int b, double* v, double* buf)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= R.off[R.nreg]) return;
    int q = 0;
    while (t >= R.off[q + 1]) ++q;
    long long f = (t - R.off[q]) / b;
    const int e = (int)((t - R.off[q]) % b);
    const int i = R.lo[q][0] + (int)(f % R.size[q][0]);
    f /= R.size[q][0];
    const int j = R.lo[q][1] + (int)(f % R.size[q][1]);
    const int k = R.lo[q][2] + (int)(f / R.size[q][1]);
    const size_t I = i + (size_t)nx * (j + (size_t)ny * k);
    if (PACK) buf[t] = v[I * b + e];
    else v[I * b + e] = buf[t];
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

// copyOwnerToAll for a block decomposition with overlap 1 (BlockPreconditioner::apply / pre): every overlap
// cell receives the value of the rank that owns it -- up to 26 neighbours (faces, edges, corners), ONE grouped NCCL exchange.
// Regions that are contiguous in the vector (the planes of a slab decomposition along the last axis) are sent from and received
// into the vector directly; the others go through one pack and one unpack launch.
int halo_exchange(dmx_ctx* ctx, double* v)
{
    if (ctx->nranks == 1 || ctx->halo_nb.empty()) return 0;
    ProfScope ps(ctx, DMX_K_HALO);
    ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
    const int nx = ctx->nc[0], ny = ctx->nc[1], b = ctx->b;
    HaloRegions S, R;
    S.nreg = R.nreg = 0;
    long long so = 0;
    for (const auto& nb : ctx->halo_nb) {
        if (nb.contiguous) continue;
        for (int a = 0; a < 3; ++a) { S.lo[S.nreg][a] = nb.slo[a]; S.size[S.nreg][a] = nb.size[a]; R.lo[R.nreg][a] = nb.rlo[a]; R.size[R.nreg][a] = nb.size[a]; }
        S.off[S.nreg] = R.off[R.nreg] = so;
        so += nb.count;
        ++S.nreg; ++R.nreg;
    }
    S.off[S.nreg] = R.off[R.nreg] = so;
    const int grid = (int)((so + 255) / 256);
    if (so > 0) { halo_pack_kernel<true><<<grid, 256, 0, ctx->stream>>>(S, nx, ny, b, v, ctx->d_send); DMX_CHECK_LAUNCH(); }
    auto at = [&](const int* lo) { return ((size_t)lo[0] + (size_t)nx * (lo[1] + (size_t)ny * lo[2])) * b; };
    DMX_NCCL(g_nccl.GroupStart());
    long long po = 0;
    for (const auto& nb : ctx->halo_nb) {
        const double* sp = nb.contiguous ? v + at(nb.slo) : ctx->d_send + po;
        double* rp = nb.contiguous ? v + at(nb.rlo) : ctx->d_recv + po;
        if (!nb.contiguous) po += nb.count;
        DMX_NCCL(g_nccl.Send(sp, (size_t)nb.count, ncclFloat64, nb.rank, comm, ctx->stream));
        DMX_NCCL(g_nccl.Recv(rp, (size_t)nb.count, ncclFloat64, nb.rank, comm, ctx->stream));
    }
    DMX_NCCL(g_nccl.GroupEnd());
    if (so > 0) { halo_pack_kernel<false><<<grid, 256, 0, ctx->stream>>>(R, nx, ny, b, v, ctx->d_recv); DMX_CHECK_LAUNCH(); }
    return 0;
}

// Global sums of the scalar products (OverlappingSchwarzScalarProduct: MPI_Allreduce).  The order in which an all-reduce adds
// the per-rank partial sums is the library's choice and changes the last bits of every dot product, which BiCGSTAB amplifies
// into iteration counts that wander by several per cent.  So the partial sums are ALL-GATHERED (count <= 8 doubles per rank)
// and every rank adds them in rank order: deterministic, identical on all ranks, and the order the CPU multi-rank oracle uses
// (oracle/dist_oracle.py ThreadComm.allreduce) -- iteration counts can be compared with == for any number of ranks.
__global__ void ordered_sum_kernel(int nranks, int count, const double* __restrict__ gathered, double* __restrict__ out)
{
    const int q = threadIdx.x;
    if (q >= count) return;
    double acc = gathered[q];
    for (int r = 1; r < nranks; ++r) acc = acc + gathered[(size_t)r * count + q];
    out[q] = acc;
}
int allreduce_sum(dmx_ctx* ctx, double* d_buf, int count)
{
    if (ctx->nranks == 1) return 0;
    if (count > 8) return fail(ctx, DMX_ERR_USAGE, "allreduce_sum: at most 8 scalars");
    if (!ctx->d_gather) DMX_CUDA(cudaMalloc((void**)&ctx->d_gather, (size_t)ctx->nranks * 8 * sizeof(double)));
    DMX_NCCL(g_nccl.AllGather(d_buf, ctx->d_gather, count, ncclFloat64, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    ordered_sum_kernel<<<1, 32, 0, ctx->stream>>>(ctx->nranks, count, ctx->d_gather, d_buf);
    DMX_CHECK_LAUNCH();
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

int allreduce_max_int(dmx_ctx* ctx, int* d_buf, int count)
{
    if (ctx->nranks == 1) return 0;
    DMX_NCCL(g_nccl.AllReduce(d_buf, d_buf, count, ncclInt32, ncclMax, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    return 0;
}

// The failure word ctx->d_flag (non-finite residual, singular diagonal block) agreed over all ranks before any of them acts on
// it: what `comm.min(succeeded)` does in FVAssembler::assemble_ (assembly/fvassembler.hh:504-509) and
// NewtonSolver::solveLinearSystem (nonlinear/newtonsolver.hh:510-523) -- every rank returns the same status, so a
// NumericalProblem on one rank becomes a dt-halving retry on all of them instead of a hang in the next collective.
int agree_flag(dmx_ctx* ctx, int* flag_out)
{
    if (int rc = allreduce_max_int(ctx, ctx->d_flag, 1)) return rc;
    DMX_CUDA(cudaMemcpyAsync(ctx->h_flag, ctx->d_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    DMX_CUDA(cudaStreamSynchronize(ctx->stream));
    *flag_out = *ctx->h_flag;
    return 0;
}

} // namespace dmx
