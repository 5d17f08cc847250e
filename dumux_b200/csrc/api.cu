// api.cu -- the C ABI (include/dumux_b200.h): context, grid/pattern set-up, data movement, Newton driver.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

using namespace dmx;

namespace {

int alloc_vectors(dmx_ctx* ctx)
{
    const size_t len = (size_t)ctx->n * ctx->b;
    for (int v = 0; v < DMX_NUM_VECS; ++v) {
        if (ctx->d_vec[v]) cudaFree(ctx->d_vec[v]);
        DMX_CUDA(cudaMalloc((void**)&ctx->d_vec[v], len * sizeof(double)));
        DMX_CUDA(cudaMemset(ctx->d_vec[v], 0, len * sizeof(double)));
    }
    double** w[] = {&ctx->d_rt, &ctx->d_p, &ctx->d_v, &ctx->d_t, &ctx->d_y, &ctx->d_z};
    for (double** p : w) {
        if (*p) cudaFree(*p);
        DMX_CUDA(cudaMalloc((void**)p, len * sizeof(double)));
        DMX_CUDA(cudaMemset(*p, 0, len * sizeof(double)));
    }
    if (ctx->d_J) cudaFree(ctx->d_J);
    if (ctx->d_ilu) { cudaFree(ctx->d_ilu); ctx->d_ilu = nullptr; }
    if (ctx->d_dinv) { cudaFree(ctx->d_dinv); ctx->d_dinv = nullptr; }
    if (ctx->d_gm) { cudaFree(ctx->d_gm); ctx->d_gm = nullptr; ctx->gm_vectors = 0; }      // sized by the old vector length
    if (ctx->d_xold) { cudaFree(ctx->d_xold); ctx->d_xold = nullptr; }
    ctx->ilu_valid = false;
    ctx->jac_diagonal = false;
    DMX_CUDA(cudaMalloc((void**)&ctx->d_J, (size_t)ctx->nnzb * ctx->b * ctx->b * sizeof(double)));
    DMX_CUDA(cudaMemset(ctx->d_J, 0, (size_t)ctx->nnzb * ctx->b * ctx->b * sizeof(double)));
    return 0;
}

int upload_pattern(dmx_ctx* ctx)
{
    if (ctx->d_rowptr) cudaFree(ctx->d_rowptr);
    if (ctx->d_colidx) cudaFree(ctx->d_colidx);
    DMX_CUDA(cudaMalloc((void**)&ctx->d_rowptr, ctx->h_rowptr.size() * sizeof(int)));
    DMX_CUDA(cudaMalloc((void**)&ctx->d_colidx, ctx->h_colidx.size() * sizeof(int)));
    DMX_CUDA(cudaMemcpy(ctx->d_rowptr, ctx->h_rowptr.data(), ctx->h_rowptr.size() * sizeof(int), cudaMemcpyHostToDevice));
    DMX_CUDA(cudaMemcpy(ctx->d_colidx, ctx->h_colidx.data(), ctx->h_colidx.size() * sizeof(int), cudaMemcpyHostToDevice));
    return build_diag(ctx);
}

// Jacobian pattern of the CCTpfa scheme on the local box: (I,I) and (I,J) for all face neighbours, columns ascending
// (assembly/jacobianpattern.hh:27-52, discretization/cellcentered/connectivitymap.hh:66-115)
void build_grid_pattern(dmx_ctx* ctx)
{
    const int nx = ctx->nc[0], ny = ctx->nc[1], nz = ctx->nc[2];
    const int n = ctx->n;
    ctx->h_rowptr.assign(n + 1, 0);
    ctx->h_colidx.clear();
    ctx->h_colidx.reserve((size_t)n * 7);
    const int dim = ctx->dim;
    for (int k = 0; k < nz; ++k)
        for (int j = 0; j < ny; ++j)
            for (int i = 0; i < nx; ++i) {
                const int I = i + nx * (j + ny * k);
                if (dim > 2 && k > 0) ctx->h_colidx.push_back(I - nx * ny);
                if (dim > 1 && j > 0) ctx->h_colidx.push_back(I - nx);
                if (i > 0) ctx->h_colidx.push_back(I - 1);
                ctx->h_colidx.push_back(I);
                if (i + 1 < nx) ctx->h_colidx.push_back(I + 1);
                if (dim > 1 && j + 1 < ny) ctx->h_colidx.push_back(I + nx);
                if (dim > 2 && k + 1 < nz) ctx->h_colidx.push_back(I + nx * ny);
                ctx->h_rowptr[I + 1] = (int)ctx->h_colidx.size();
            }
    ctx->nnzb = (long long)ctx->h_colidx.size();
}

int setup_geometry(dmx_ctx* ctx)
{
    // per-axis arrays: width, |(x_f - x_c)/|x_f - x_c|^2| for the low and the high face of each cell
    // (YaspGrid: x_i = origin + i*h; AxisAlignedCubeGeometry::center = 0.5*(lower+upper);
    //  computeTpfaTransmissibility: d = x_f - x_c; d /= |d|^2; t*extrusion*(d.n)) [dune-grid semantics]
    size_t total = 0;
    for (int a = 0; a < 3; ++a) total += 3 * (size_t)ctx->nc[a];
    std::vector<double> h(total);
    size_t o = 0;
    size_t offs[3][3];
    for (int a = 0; a < 3; ++a) {
        const int m = ctx->nc[a];
        offs[a][0] = o; offs[a][1] = o + m; offs[a][2] = o + 2 * (size_t)m;
        for (int i = 0; i < m; ++i) {
            const double x0 = ctx->xn[a][i], x1 = ctx->xn[a][i + 1];
            const double c = 0.5 * (x0 + x1);
            h[o + i] = x1 - x0;
            double dlo = 0.5 * (x0 + x0) - c;
            dlo /= dlo * dlo;
            double dhi = 0.5 * (x1 + x1) - c;
            dhi /= dhi * dhi;
            h[o + m + i] = dlo * -1.0;
            h[o + 2 * (size_t)m + i] = dhi * 1.0;
        }
        o += 3 * (size_t)m;
    }
    if (ctx->d_geom) cudaFree(ctx->d_geom);
    DMX_CUDA(cudaMalloc((void**)&ctx->d_geom, total * sizeof(double)));
    DMX_CUDA(cudaMemcpy(ctx->d_geom, h.data(), total * sizeof(double), cudaMemcpyHostToDevice));
    for (int a = 0; a < 3; ++a) {
        ctx->d_width[a] = ctx->d_geom + offs[a][0];
        ctx->d_gflo[a] = ctx->d_geom + offs[a][1];
        ctx->d_gfhi[a] = ctx->d_geom + offs[a][2];
    }
    return 0;
}

int finish_grid(dmx_ctx* ctx, int model, int dim, const int* cells, const std::vector<double> (&gx)[3])
{
    if (model != DMX_MODEL_1P && model != DMX_MODEL_2P && model != DMX_MODEL_TRACER) return fail(ctx, DMX_ERR_USAGE, "unknown model");
    if (dim < 1 || dim > 3) return fail(ctx, DMX_ERR_USAGE, "dim must be 1..3");
    DMX_CUDA(cudaSetDevice(ctx->device));
    ctx->model = model;
    ctx->b = (model == DMX_MODEL_2P) ? 2 : 1;
    ctx->dim = dim;
    for (int a = 0; a < 3; ++a) { ctx->gcells[a] = (a < dim) ? cells[a] : 1; ctx->nc[a] = ctx->gcells[a]; ctx->off[a] = 0; }
    // Block decomposition with overlap 1 (Grid.Partitioning "px py pz" -> Dune::Yasp::FixedSizePartitioning, Grid.Overlap 1;
    // io/grid/gridmanager_yasp.hh:129,194-203).  Default: slabs along the last axis ("1 .. P").  Rank -> torus coordinate with x
    // fastest and, per axis, n/P cells for the first P - n%P processes and one more for the rest (dune-grid torus.hh
    // Torus::rank_to_coord / Torus::partition) [DUNE-ext].
    for (int a = 0; a < 3; ++a) {
        if (!ctx->explicit_box) { ctx->part[a] = 1; ctx->pcoord[a] = 0; }
        ctx->own_lo[a] = 0; ctx->own_hi[a] = ctx->gcells[a];
    }
    if (ctx->nranks > 1 && ctx->explicit_box) {
        // level context of a block-decomposed AMG hierarchy: part / pcoord are the parent's, the box is handed in
        for (int a = 0; a < 3; ++a) {
            ctx->off[a] = ctx->xb_off[a]; ctx->nc[a] = ctx->xb_nc[a]; ctx->own_lo[a] = ctx->xb_own_lo[a]; ctx->own_hi[a] = ctx->xb_own_hi[a];
        }
    } else if (ctx->nranks > 1) {
        if (ctx->part_req[0] > 0) {
            for (int a = 0; a < 3; ++a) ctx->part[a] = ctx->part_req[a];
            for (int a = dim; a < 3; ++a)
                if (ctx->part[a] != 1) return fail(ctx, DMX_ERR_USAGE, "Grid.Partitioning: more than one rank along an axis the grid does not have");
        } else ctx->part[dim - 1] = ctx->nranks;
        if (ctx->part[0] * ctx->part[1] * ctx->part[2] != ctx->nranks)
            return fail(ctx, DMX_ERR_USAGE, "Grid.Partitioning: the product of the per-axis rank counts must equal the number of ranks");
        int r = ctx->rank;
        for (int a = 0; a < 3; ++a) { ctx->pcoord[a] = r % ctx->part[a]; r /= ctx->part[a]; }
        for (int a = 0; a < 3; ++a) {
            const int N = ctx->gcells[a], P = ctx->part[a], c = ctx->pcoord[a];
            if (N < P) return fail(ctx, DMX_ERR_USAGE, "fewer cell layers than ranks along a partitioned axis");
            const int m = N / P, rem = N % P;
            const int b0 = (c < P - rem) ? c * m : (P - rem) * m + (c - (P - rem)) * (m + 1);
            const int b1 = b0 + ((c < P - rem) ? m : m + 1);
            const int lo = std::max(0, b0 - 1), hi = std::min(N, b1 + 1);
            ctx->off[a] = lo;
            ctx->nc[a] = hi - lo;
            ctx->own_lo[a] = b0 - lo;
            ctx->own_hi[a] = b1 - lo;
        }
    }
    for (int a = 0; a < 3; ++a) {
        ctx->xn[a].assign(gx[a].begin() + ctx->off[a], gx[a].begin() + ctx->off[a] + ctx->nc[a] + 1);
    }
    ctx->n = ctx->nc[0] * ctx->nc[1] * ctx->nc[2];
    ctx->has_grid = true;
    ctx->prepared = false;
    amg_free(ctx);
    ctx->amg_dirty = true;
    ctx->h_K.assign(ctx->n, 1e-10);
    ctx->h_phi.assign(ctx->n, 0.4);
    ctx->h_region.assign(ctx->n, 0);
    for (int s = 0; s < 6; ++s) { ctx->h_bc_type[s].clear(); ctx->h_bc_val[s].clear(); }
    // processor boundaries: outer faces of overlap layers carry no scvf (tpfa/fvgridgeometry.hh:272-320)
    for (int a = 0; a < 3 && ctx->nranks > 1; ++a) {
        int nf = 1;
        for (int d = 0; d < 3; ++d) if (d != a) nf *= ctx->nc[d];
        if (ctx->off[a] > 0) { ctx->h_bc_type[2 * a].assign(nf, DMX_BC_NONE); ctx->h_bc_val[2 * a].assign((size_t)nf * ctx->b, 0.0); }
        if (ctx->off[a] + ctx->nc[a] < ctx->gcells[a]) { ctx->h_bc_type[2 * a + 1].assign(nf, DMX_BC_NONE); ctx->h_bc_val[2 * a + 1].assign((size_t)nf * ctx->b, 0.0); }
    }
    build_grid_pattern(ctx);
    if (int rc = setup_geometry(ctx)) return rc;
    auto up = [&](auto** d, const auto& h) -> int {
        if (*d) cudaFree(*d);
        DMX_CUDA(cudaMalloc((void**)d, h.size() * sizeof(h[0])));
        DMX_CUDA(cudaMemcpy(*d, h.data(), h.size() * sizeof(h[0]), cudaMemcpyHostToDevice));
        return 0;
    };
    if (int rc = up(&ctx->d_K, ctx->h_K)) return rc;
    if (int rc = up(&ctx->d_phi, ctx->h_phi)) return rc;
    if (int rc = up(&ctx->d_region, ctx->h_region)) return rc;
    if (ctx->d_q) { cudaFree(ctx->d_q); ctx->d_q = nullptr; }
    for (int a = 0; a < 3; ++a) if (ctx->d_tij[a]) { cudaFree(ctx->d_tij[a]); ctx->d_tij[a] = nullptr; }
    for (int a = 0; a < 3; ++a) if (ctx->d_Kaxis[a]) { cudaFree(ctx->d_Kaxis[a]); ctx->d_Kaxis[a] = nullptr; }
    if (ctx->d_law_rec) { cudaFree(ctx->d_law_rec); ctx->d_law_rec = nullptr; }
    if (ctx->d_disp) { cudaFree(ctx->d_disp); ctx->d_disp = nullptr; }
    if (int rc = upload_pattern(ctx)) return rc;
    if (int rc = alloc_vectors(ctx)) return rc;
    // structured-grid ILU sweeps (DMX_ILU_GENERIC=1 keeps the level-scheduled generic kernels, for A/B runs)
    sk_free(ctx);
    {
        const char* env = getenv("DMX_ILU_GENERIC");
        if (!(env && env[0] == '1'))
            if (int rc = sk_setup(ctx)) return rc;
    }
    // owner mask (distributed): a cell is owner where it is interior (parallelhelpers.hh:485-497)
    if (ctx->d_owner) { cudaFree(ctx->d_owner); ctx->d_owner = nullptr; }
    if (ctx->d_send) { cudaFree(ctx->d_send); ctx->d_send = nullptr; }
    if (ctx->d_recv) { cudaFree(ctx->d_recv); ctx->d_recv = nullptr; }
    ctx->halo_nb.clear();
    ctx->halo_total = 0;
    if (ctx->nranks > 1) {
        std::vector<unsigned char> own(ctx->n, 0);
        size_t I = 0;
        for (int k = 0; k < ctx->nc[2]; ++k)
            for (int j = 0; j < ctx->nc[1]; ++j)
                for (int i = 0; i < ctx->nc[0]; ++i, ++I)
                    own[I] = (i >= ctx->own_lo[0] && i < ctx->own_hi[0] && j >= ctx->own_lo[1] && j < ctx->own_hi[1] && k >= ctx->own_lo[2] &&
                              k < ctx->own_hi[2]) ? 1 : 0;
        if (int rc = up(&ctx->d_owner, own)) return rc;
        // copyOwnerToAll neighbour table: direction (dx,dy,dz) in {-1,0,1}^3 \ 0, z slowest.  Along an axis with d = -1 my first
        // owned layer goes out and my low overlap layer comes in, d = +1 likewise at the high end, d = 0: the owned range.
        for (int dz = -1; dz <= 1; ++dz)
            for (int dy = -1; dy <= 1; ++dy)
                for (int dx = -1; dx <= 1; ++dx) {
                    const int d[3] = {dx, dy, dz};
                    if (!dx && !dy && !dz) continue;
                    bool ok = true;
                    for (int a = 0; a < 3; ++a) { const int c = ctx->pcoord[a] + d[a]; if (c < 0 || c >= ctx->part[a]) ok = false; }
                    if (!ok) continue;
                    dmx_ctx::HaloNb nb;
                    nb.rank = (ctx->pcoord[0] + dx) + ctx->part[0] * ((ctx->pcoord[1] + dy) + ctx->part[1] * (ctx->pcoord[2] + dz));
                    long long cnt = ctx->b;
                    for (int a = 0; a < 3; ++a) {
                        if (d[a] < 0) { nb.slo[a] = ctx->own_lo[a]; nb.rlo[a] = ctx->own_lo[a] - 1; nb.size[a] = 1; }
                        else if (d[a] > 0) { nb.slo[a] = ctx->own_hi[a] - 1; nb.rlo[a] = ctx->own_hi[a]; nb.size[a] = 1; }
                        else { nb.slo[a] = nb.rlo[a] = ctx->own_lo[a]; nb.size[a] = ctx->own_hi[a] - ctx->own_lo[a]; }
                        cnt *= nb.size[a];
                    }
                    nb.count = cnt;
                    nb.off = ctx->halo_total;
                    // a region that spans the full local extent of every faster axis is one contiguous range of the vector
                    nb.contiguous = nb.size[0] == ctx->nc[0] && (nb.size[1] == ctx->nc[1] || nb.size[2] == 1);
                    ctx->halo_total += cnt;
                    ctx->halo_nb.push_back(nb);
                }
        if (ctx->halo_total > 0) {
            DMX_CUDA(cudaMalloc((void**)&ctx->d_send, (size_t)ctx->halo_total * sizeof(double)));
            DMX_CUDA(cudaMalloc((void**)&ctx->d_recv, (size_t)ctx->halo_total * sizeof(double)));
        }
    }
    return 0;
}

double now_ms()
{
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

} // namespace

extern "C" {

const char* dmx_version(void) { return "dumux_b200 0.1 (sm_100a)"; }

void dmx_default_options(dmx_options* o)
{
    o->enable_gravity = 1; o->gravity = 9.81; o->upwind_weight = 1.0; o->fd_method = 1; o->base_eps = 1e-10;
    o->privar_magnitude[0] = o->privar_magnitude[1] = -1.0; o->stationary = 0; o->dt = 1.0; o->extrusion = 1.0;
}
void dmx_default_amg_params(dmx_amg_params* p)
{
    p->pre_steps = 2; p->post_steps = 2; p->prolongation_damping = 1.6; p->smoother = DMX_PRECOND_SSOR;
    p->coarsest_cells = 8; p->coarsest_steps = 8; p->max_levels = 15;
    p->smoother_iterations = 1; p->smoother_relaxation = 1.0;
}
void dmx_default_newton_params(dmx_newton_params* p)
{
    p->max_relative_shift = 1e-8; p->min_steps = 2; p->max_steps = 18; p->lin_reduction = 1e-6; p->lin_maxit = 250;
    p->preconditioner = DMX_PRECOND_ILU0;
    p->use_line_search = 0; p->line_search_min_relaxation = 0.125;
    p->enable_shift_criterion = 1; p->enable_residual_criterion = 0; p->enable_absolute_residual_criterion = 0;
    p->satisfy_residual_and_shift = 0; p->residual_reduction = 1e-5; p->max_absolute_residual = 1e-5;
}

static int alloc_ctx_scratch(dmx_ctx* ctx)
{
    cudaMalloc((void**)&ctx->d_partials, 3 * 1184 * sizeof(double));
    cudaMalloc((void**)&ctx->d_scalars, 8 * sizeof(double));
    cudaMallocHost((void**)&ctx->h_scalars, 8 * sizeof(double));
    cudaMalloc((void**)&ctx->d_flag, sizeof(int));
    cudaMallocHost((void**)&ctx->h_flag, sizeof(int));
    cudaMalloc((void**)&ctx->d_barrier, sizeof(unsigned int));
    for (auto& e : ctx->ev) cudaEventCreate(&e);
    return cudaGetLastError() == cudaSuccess ? 0 : DMX_ERR_CUDA;
}

int dmx_create_distributed(dmx_ctx** out, int device, const void* uid, int rank, int nranks)
{
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) return DMX_ERR_CUDA;   // no CPU fallback
    if (device < 0 || device >= count) return DMX_ERR_USAGE;
    dmx_ctx* ctx = new dmx_ctx;
    ctx->device = device;
    ctx->rank = rank;
    ctx->nranks = nranks;
    dmx_default_options(&ctx->opt);
    dmx_default_amg_params(&ctx->amg_prm);
    if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return DMX_ERR_CUDA; }
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    ctx->num_sms = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return DMX_ERR_CUDA; }
    if (alloc_ctx_scratch(ctx)) { delete ctx; return DMX_ERR_CUDA; }
    if (nranks > 1) {
        if (int rc = nccl_init(ctx, uid)) { delete ctx; return rc; }
    }
    *out = ctx;
    return 0;
}
int dmx_create(dmx_ctx** out, int device) { return dmx_create_distributed(out, device, nullptr, 0, 1); }
int dmx_get_nccl_unique_id(void* out128) { return nccl_get_unique_id(out128); }

int dmx_destroy(dmx_ctx* ctx)
{
    if (!ctx) return 0;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    amg_free(ctx);
    sk_free(ctx);
    if (ctx->nccl_comm && ctx->owns_stream) nccl_destroy(ctx);
    void* ptrs[] = {ctx->d_geom, ctx->d_K, ctx->d_phi, ctx->d_q, ctx->d_region, ctx->d_tij[0], ctx->d_tij[1], ctx->d_tij[2], ctx->d_laws,
                    ctx->d_tab_buf, ctx->d_rowptr, ctx->d_colidx, ctx->d_diag, ctx->d_J, ctx->d_ilu, ctx->d_rt, ctx->d_p, ctx->d_v, ctx->d_t,
                    ctx->d_y, ctx->d_z, ctx->d_dinv, ctx->d_gm, ctx->d_vf, ctx->d_color_rows, ctx->d_xold, ctx->d_lrows, ctx->d_urows, ctx->d_lptr, ctx->d_uptr, ctx->d_barrier, ctx->d_partials,
                    ctx->d_scalars, ctx->d_flag, ctx->d_owner, ctx->d_send, ctx->d_recv, ctx->d_gather, ctx->d_law_rec,
                    ctx->d_Kaxis[0], ctx->d_Kaxis[1], ctx->d_Kaxis[2], ctx->d_disp};
    for (void* p : ptrs) if (p) cudaFree(p);
    for (int v = 0; v < DMX_NUM_VECS; ++v) if (ctx->d_vec[v]) cudaFree(ctx->d_vec[v]);
    for (int s = 0; s < 6; ++s) {
        void* b[] = {ctx->d_bc_type[s], ctx->d_bc_neumann[s], ctx->d_bc_p[s], ctx->d_bc_up[s], ctx->d_bc_rho[s]};
        for (void* p : b) if (p) cudaFree(p);
    }
    if (ctx->h_scalars) cudaFreeHost(ctx->h_scalars);
    if (ctx->h_flag) cudaFreeHost(ctx->h_flag);
    for (auto& e : ctx->ev) if (e) cudaEventDestroy(e);
    for (auto& r : ctx->prof_pending) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    for (auto& e : ctx->prof_pool) cudaEventDestroy(e);
    if (ctx->stream && ctx->owns_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return 0;
}
const char* dmx_last_error(const dmx_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int dmx_grid_structured(dmx_ctx* ctx, int model, int dim, const int* cells, const double* lower, const double* upper)
{
    std::vector<double> gx[3];
    for (int a = 0; a < 3; ++a) {
        const int m = (a < dim) ? cells[a] : 1;
        gx[a].resize(m + 1);
        if (a < dim) {
            const double h = (upper[a] - lower[a]) / cells[a];
            for (int i = 0; i <= m; ++i) gx[a][i] = lower[a] + i * h;
        } else { gx[a][0] = 0.0; gx[a][1] = 1.0; }
    }
    return finish_grid(ctx, model, dim, cells, gx);
}
int dmx_grid_tensor(dmx_ctx* ctx, int model, int dim, const int* cells, const double* x, const double* y, const double* z)
{
    const double* src[3] = {x, y, z};
    std::vector<double> gx[3];
    for (int a = 0; a < 3; ++a) {
        const int m = (a < dim) ? cells[a] : 1;
        gx[a].resize(m + 1);
        if (a < dim) for (int i = 0; i <= m; ++i) gx[a][i] = src[a][i];
        else { gx[a][0] = 0.0; gx[a][1] = 1.0; }
    }
    return finish_grid(ctx, model, dim, cells, gx);
}
int dmx_local_box(const dmx_ctx* ctx, int* cells, int* offset, int* owned_begin, int* owned_end)
{
    for (int a = 0; a < 3; ++a) { cells[a] = ctx->nc[a]; offset[a] = ctx->off[a]; }
    // owned range along the last grid axis (the slab axis of the default partitioning)
    *owned_begin = ctx->own_lo[ctx->dim > 0 ? ctx->dim - 1 : 0];
    *owned_end = ctx->own_hi[ctx->dim > 0 ? ctx->dim - 1 : 0];
    return 0;
}
int dmx_set_partitioning(dmx_ctx* ctx, const int* ranks_per_axis)
{
    if (!ranks_per_axis) { ctx->part_req[0] = ctx->part_req[1] = ctx->part_req[2] = 0; return 0; }
    for (int a = 0; a < 3; ++a) {
        if (ranks_per_axis[a] < 1) return fail(ctx, DMX_ERR_USAGE, "set_partitioning: ranks per axis must be >= 1");
        ctx->part_req[a] = ranks_per_axis[a];
    }
    return 0;
}
int dmx_local_box3(const dmx_ctx* ctx, int* cells, int* offset, int* owned_begin, int* owned_end, int* ranks_per_axis, int* coord)
{
    for (int a = 0; a < 3; ++a) {
        cells[a] = ctx->nc[a]; offset[a] = ctx->off[a]; owned_begin[a] = ctx->own_lo[a]; owned_end[a] = ctx->own_hi[a];
        if (ranks_per_axis) ranks_per_axis[a] = ctx->part[a];
        if (coord) coord[a] = ctx->pcoord[a];
    }
    return 0;
}
int dmx_num_cells(const dmx_ctx* ctx) { return ctx->n; }
int dmx_num_eq(const dmx_ctx* ctx) { return ctx->b; }
long long dmx_nnz_blocks(const dmx_ctx* ctx) { return ctx->nnzb; }
int dmx_pattern(const dmx_ctx* ctx, int* rowptr, int* colidx)
{
    std::memcpy(rowptr, ctx->h_rowptr.data(), ctx->h_rowptr.size() * sizeof(int));
    std::memcpy(colidx, ctx->h_colidx.data(), ctx->h_colidx.size() * sizeof(int));
    return 0;
}
int dmx_bcrs_pattern(dmx_ctx* ctx, int n, int b, const int* rowptr, const int* colidx)
{
    if (b != 1 && b != 2) return fail(ctx, DMX_ERR_USAGE, "block size must be 1 or 2");
    DMX_CUDA(cudaSetDevice(ctx->device));
    sk_free(ctx);
    amg_free(ctx);
    ctx->amg_dirty = true;
    ctx->has_grid = false;
    ctx->model = 0;
    ctx->n = n;
    ctx->b = b;
    ctx->h_rowptr.assign(rowptr, rowptr + n + 1);
    ctx->h_colidx.assign(colidx, colidx + rowptr[n]);
    ctx->nnzb = rowptr[n];
    for (int i = 0; i < n; ++i)
        for (int k = rowptr[i] + 1; k < rowptr[i + 1]; ++k)
            if (colidx[k] <= colidx[k - 1]) return fail(ctx, DMX_ERR_USAGE, "column indices must be strictly ascending per row");
    if (int rc = upload_pattern(ctx)) return rc;
    return alloc_vectors(ctx);
}

int dmx_set_options(dmx_ctx* ctx, const dmx_options* o)
{
    const bool structural = (o->fd_method != ctx->opt.fd_method) || (o->extrusion != ctx->opt.extrusion);
    ctx->opt = *o;
    if (o->fd_method != 1 && o->fd_method != 0 && o->fd_method != -1 && o->fd_method != 5 && o->fd_method != DMX_DIFF_ANALYTIC)
        return fail(ctx, DMX_ERR_USAGE, "Assembly.NumericDifferenceMethod must be 1, 0, -1 or 5 (or DMX_DIFF_ANALYTIC)");
    if (structural) ctx->prepared = false;
    return 0;
}
int dmx_set_cell_fields(dmx_ctx* ctx, const double* K, const double* phi, const int* region)
{
    if (!ctx->has_grid) return fail(ctx, DMX_ERR_USAGE, "set grid first");
    DMX_CUDA(cudaSetDevice(ctx->device));
    const int n = ctx->n;
    if (K) { ctx->h_K.assign(K, K + n); DMX_CUDA(cudaMemcpy(ctx->d_K, K, n * sizeof(double), cudaMemcpyHostToDevice)); }
    if (phi) { ctx->h_phi.assign(phi, phi + n); DMX_CUDA(cudaMemcpy(ctx->d_phi, phi, n * sizeof(double), cudaMemcpyHostToDevice)); }
    if (region) {
        for (int i = 0; i < n; ++i)
            if (region[i] < 0 || region[i] >= DMX_MAX_REGIONS) return fail(ctx, DMX_ERR_USAGE, "region id out of range");
        ctx->h_region.assign(region, region + n);
        DMX_CUDA(cudaMemcpy(ctx->d_region, region, n * sizeof(int), cudaMemcpyHostToDevice));
    }
    ctx->prepared = false;
    return 0;
}
int dmx_set_permeability_diagonal(dmx_ctx* ctx, const double* kx, const double* ky, const double* kz)
{
    if (!ctx->has_grid) return fail(ctx, DMX_ERR_USAGE, "set grid first");
    DMX_CUDA(cudaSetDevice(ctx->device));
    const double* k[3] = {kx, ky, kz};
    const size_t n = (size_t)ctx->n;
    bool any = false;
    for (int a = 0; a < ctx->dim; ++a) any = any || k[a];
    for (int a = 0; a < 3; ++a)
        if (ctx->d_Kaxis[a]) { cudaFree(ctx->d_Kaxis[a]); ctx->d_Kaxis[a] = nullptr; }
    ctx->prepared = false;
    if (!any) return 0;
    for (int a = 0; a < ctx->dim; ++a) {
        if (!k[a]) return fail(ctx, DMX_ERR_USAGE, "permeability_diagonal: one array per grid axis");
        DMX_CUDA(cudaMalloc((void**)&ctx->d_Kaxis[a], n * sizeof(double)));
        DMX_CUDA(cudaMemcpy(ctx->d_Kaxis[a], k[a], n * sizeof(double), cudaMemcpyHostToDevice));
    }
    // the scalar field keeps the entry along the gravity axis (what the gravity terms of the vertical faces use)
    const int va = ctx->dim - 1;
    ctx->h_K.assign(k[va], k[va] + n);
    DMX_CUDA(cudaMemcpy(ctx->d_K, k[va], n * sizeof(double), cudaMemcpyHostToDevice));
    return 0;
}
int dmx_set_source(dmx_ctx* ctx, const double* q)
{
    if (!ctx->has_grid) return fail(ctx, DMX_ERR_USAGE, "set grid first");
    DMX_CUDA(cudaSetDevice(ctx->device));
    const size_t len = (size_t)ctx->n * ctx->b;
    if (!ctx->d_q) DMX_CUDA(cudaMalloc((void**)&ctx->d_q, len * sizeof(double)));
    DMX_CUDA(cudaMemcpy(ctx->d_q, q, len * sizeof(double), cudaMemcpyHostToDevice));
    return 0;
}
int dmx_set_material(dmx_ctx* ctx, int region, int law, const double* params, double swr, double snr, int regularize, const double* reg)
{
    if (region < 0 || region >= DMX_MAX_REGIONS) return fail(ctx, DMX_ERR_USAGE, "region id out of range");
    if ((int)ctx->laws.size() <= region) ctx->laws.resize(region + 1, MaterialLaw{});
    MaterialLaw& l = ctx->laws[region];
    std::memset(&l, 0, sizeof(l));
    l.kind = law; l.regularized = regularize ? 1 : 0; l.swr = swr; l.snr = snr;
    l.pcLowSwe = 0.01; l.pcHighSwe = 0.99; l.krnLowSwe = 0.1; l.krwHighSwe = 0.9;
    if (law == DMX_LAW_BROOKSCOREY) {
        l.pcEntry = params[0]; l.lambda = params[1];
        if (reg) l.pcLowSwe = reg[0];
    } else if (law == DMX_LAW_VANGENUCHTEN) {
        l.alpha = params[0]; l.n = params[1]; l.m = 1.0 - 1.0 / l.n; l.l = params[2];
        if (reg) { l.pcLowSwe = reg[0]; l.pcHighSwe = reg[1]; l.krnLowSwe = reg[2]; l.krwHighSwe = reg[3]; }
    } else return fail(ctx, DMX_ERR_USAGE, "unknown material law");
    ctx->prepared = false;
    return 0;
}
int dmx_set_wetting_phase(dmx_ctx* ctx, int region, int phase)
{
    if (region < 0 || region >= DMX_MAX_REGIONS) return fail(ctx, DMX_ERR_USAGE, "region id out of range");
    if (phase != 0 && phase != 1) return fail(ctx, DMX_ERR_USAGE, "wetting phase must be 0 or 1");
    if ((int)ctx->laws.size() <= region) return fail(ctx, DMX_ERR_USAGE, "set_wetting_phase: call dmx_set_material for the region first");
    ctx->laws[region].wetting = phase;
    ctx->prepared = false;
    return 0;
}
int dmx_set_fluids(dmx_ctx* ctx, const double* density, const double* viscosity)
{
    const int nph = (ctx->model == DMX_MODEL_2P) ? 2 : 1;
    for (int i = 0; i < nph; ++i) { ctx->rho[i] = density[i]; ctx->mu[i] = viscosity[i]; }
    ctx->tabulated = false;
    ctx->prepared = false;
    return 0;
}
int dmx_set_fluid_table(dmx_ctx* ctx, int nT, int nP, double Tmin, double Tmax, const double* pmin, const double* pmax,
                        const double* density, const double* viscosity, double temperature)
{
    if (ctx->model != DMX_MODEL_1P) return fail(ctx, DMX_ERR_USAGE, "tabulated fluid is supported for the 1p model");
    DMX_CUDA(cudaSetDevice(ctx->device));
    ctx->h_tab_pmin.assign(pmin, pmin + nT);
    ctx->h_tab_pmax.assign(pmax, pmax + nT);
    ctx->h_tab_rho.assign(density, density + (size_t)nT * nP);
    ctx->h_tab_mu.assign(viscosity, viscosity + (size_t)nT * nP);
    ctx->h_table = FluidTable{nT, nP, Tmin, Tmax, temperature, ctx->h_tab_pmin.data(), ctx->h_tab_pmax.data(), ctx->h_tab_rho.data(),
                              ctx->h_tab_mu.data()};
    const size_t total = 2 * (size_t)nT + 2 * (size_t)nT * nP;
    if (ctx->d_tab_buf) cudaFree(ctx->d_tab_buf);
    DMX_CUDA(cudaMalloc((void**)&ctx->d_tab_buf, total * sizeof(double)));
    double* d = ctx->d_tab_buf;
    DMX_CUDA(cudaMemcpy(d, pmin, nT * sizeof(double), cudaMemcpyHostToDevice));
    DMX_CUDA(cudaMemcpy(d + nT, pmax, nT * sizeof(double), cudaMemcpyHostToDevice));
    DMX_CUDA(cudaMemcpy(d + 2 * nT, density, (size_t)nT * nP * sizeof(double), cudaMemcpyHostToDevice));
    DMX_CUDA(cudaMemcpy(d + 2 * nT + (size_t)nT * nP, viscosity, (size_t)nT * nP * sizeof(double), cudaMemcpyHostToDevice));
    ctx->d_table = FluidTable{nT, nP, Tmin, Tmax, temperature, d, d + nT, d + 2 * nT, d + 2 * nT + (size_t)nT * nP};
    ctx->tabulated = true;
    ctx->prepared = false;
    return 0;
}
int dmx_side_faces(const dmx_ctx* ctx, int side)
{
    const int a = side / 2;
    int nf = 1;
    for (int d = 0; d < 3; ++d) if (d != a) nf *= ctx->nc[d];
    return nf;
}
int dmx_set_boundary(dmx_ctx* ctx, int side, const int* type, const double* values)
{
    if (!ctx->has_grid) return fail(ctx, DMX_ERR_USAGE, "set grid first");
    if (side < 0 || side >= 2 * ctx->dim) return fail(ctx, DMX_ERR_USAGE, "side out of range");
    // a processor boundary keeps its DMX_BC_NONE marking
    const int sa = side / 2;
    if (ctx->nranks > 1 && (((side & 1) == 0 && ctx->off[sa] > 0) || ((side & 1) == 1 && ctx->off[sa] + ctx->nc[sa] < ctx->gcells[sa])))
        return 0;
    const int nf = dmx_side_faces(ctx, side);
    ctx->h_bc_type[side].assign(type, type + nf);
    ctx->h_bc_val[side].assign(values, values + (size_t)nf * ctx->b);
    ctx->prepared = false;
    return 0;
}

// ---- vectors ----
static int vec_ok(dmx_ctx* ctx, int v)
{
    if (v < 0 || v >= DMX_NUM_VECS || !ctx->d_vec[v]) return fail(ctx, DMX_ERR_USAGE, "bad vector id or no pattern set");
    return 0;
}
int dmx_vec_upload(dmx_ctx* ctx, int vec, const double* host)
{
    if (int rc = vec_ok(ctx, vec)) return rc;
    DMX_CUDA(cudaMemcpyAsync(ctx->d_vec[vec], host, (size_t)ctx->n * ctx->b * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    DMX_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}
int dmx_vec_download(dmx_ctx* ctx, int vec, double* host)
{
    if (int rc = vec_ok(ctx, vec)) return rc;
    DMX_CUDA(cudaMemcpyAsync(host, ctx->d_vec[vec], (size_t)ctx->n * ctx->b * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    DMX_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}
int dmx_vec_copy(dmx_ctx* ctx, int dst, int src)
{
    if (int rc = vec_ok(ctx, dst)) return rc;
    if (int rc = vec_ok(ctx, src)) return rc;
    DMX_CUDA(cudaMemcpyAsync(ctx->d_vec[dst], ctx->d_vec[src], (size_t)ctx->n * ctx->b * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    return 0;
}
int dmx_jacobian_upload(dmx_ctx* ctx, const double* values)
{
    ctx->jac_diagonal = false;
    if (!ctx->d_J) return fail(ctx, DMX_ERR_USAGE, "no pattern set");
    DMX_CUDA(cudaMemcpyAsync(ctx->d_J, values, (size_t)ctx->nnzb * ctx->b * ctx->b * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    DMX_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}
int dmx_jacobian_download(dmx_ctx* ctx, double* values)
{
    if (!ctx->d_J) return fail(ctx, DMX_ERR_USAGE, "no pattern set");
    DMX_CUDA(cudaMemcpyAsync(values, ctx->d_J, (size_t)ctx->nnzb * ctx->b * ctx->b * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    DMX_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}
void* dmx_vec_device_ptr(dmx_ctx* ctx, int vec) { return (vec >= 0 && vec < DMX_NUM_VECS) ? ctx->d_vec[vec] : nullptr; }
void* dmx_jacobian_device_ptr(dmx_ctx* ctx) { ctx->jac_diagonal = false; return ctx->d_J; }   /* the caller may write through it */
int dmx_synchronize(dmx_ctx* ctx)
{
    DMX_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}
int dmx_kernel_launch_count(const dmx_ctx* ctx, long long* launches)
{
    long long total = ctx->launches;
    for (const dmx_ctx* c : ctx->children) total += c->launches;
    *launches = total;
    return 0;
}

int dmx_profile(dmx_ctx* ctx, int enable)
{
    DMX_CUDA(cudaSetDevice(ctx->device));
    DMX_CUDA(cudaStreamSynchronize(ctx->stream));
    prof_drain(ctx);
    if (enable) {
        for (int k = 0; k < DMX_NUM_KCLASS; ++k) { ctx->prof_ms[k] = 0.0; ctx->prof_n[k] = 0; }
    }
    amg_level_profile(ctx, 0, nullptr, nullptr, enable != 0);
    ctx->prof_on = enable != 0;
    return 0;
}
int dmx_amg_level_profile(dmx_ctx* ctx, int level, double* ms_total, long long* cycles)
{
    DMX_CUDA(cudaSetDevice(ctx->device));
    DMX_CUDA(cudaStreamSynchronize(ctx->stream));
    *ms_total = 0.0;
    *cycles = 0;
    if (amg_level_profile(ctx, level, ms_total, cycles, false)) return fail(ctx, DMX_ERR_USAGE, "no such AMG level");
    return 0;
}
int dmx_profile_read(dmx_ctx* ctx, int kclass, double* ms_total, long long* units)
{
    if (kclass < 0 || kclass >= DMX_NUM_KCLASS) return fail(ctx, DMX_ERR_USAGE, "unknown kernel class");
    DMX_CUDA(cudaSetDevice(ctx->device));
    DMX_CUDA(cudaStreamSynchronize(ctx->stream));
    prof_drain(ctx);
    *ms_total = ctx->prof_ms[kclass];
    *units = ctx->prof_n[kclass];
    return 0;
}

// ---- hot path ----
int dmx_volume_flux(dmx_ctx* ctx, double* out)
{
    if (!ctx->has_grid || ctx->model != DMX_MODEL_1P) return fail(ctx, DMX_ERR_USAGE, "volume_flux: needs a 1p ctx with a grid");
    DMX_CUDA(cudaSetDevice(ctx->device));
    const size_t len = (size_t)ctx->n * 2 * ctx->dim;
    double* d_out = nullptr;
    DMX_CUDA(cudaMalloc((void**)&d_out, len * sizeof(double)));
    int rc = launch_volume_flux(ctx, d_out);
    if (!rc) {
        cudaError_t e = cudaMemcpyAsync(out, d_out, len * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) rc = fail(ctx, DMX_ERR_CUDA, cudaGetErrorString(e));
    }
    cudaFree(d_out);
    return rc;
}
int dmx_set_volume_flux(dmx_ctx* ctx, const double* vf)
{
    if (!ctx->has_grid || ctx->model != DMX_MODEL_TRACER) return fail(ctx, DMX_ERR_USAGE, "set_volume_flux: needs a tracer ctx with a grid");
    DMX_CUDA(cudaSetDevice(ctx->device));
    const size_t len = (size_t)ctx->n * 2 * ctx->dim;
    if (!ctx->d_vf) DMX_CUDA(cudaMalloc((void**)&ctx->d_vf, len * sizeof(double)));
    DMX_CUDA(cudaMemcpyAsync(ctx->d_vf, vf, len * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    DMX_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}
int dmx_set_tracer_dispersion(dmx_ctx* ctx, const double* disp)
{
    if (!ctx->has_grid || ctx->model != DMX_MODEL_TRACER) return fail(ctx, DMX_ERR_USAGE, "set_tracer_dispersion: needs a tracer context with a grid");
    DMX_CUDA(cudaSetDevice(ctx->device));
    if (ctx->d_disp) { cudaFree(ctx->d_disp); ctx->d_disp = nullptr; }
    if (!disp) return 0;
    const size_t len = (size_t)ctx->n * 2 * ctx->dim;
    DMX_CUDA(cudaMalloc((void**)&ctx->d_disp, len * sizeof(double)));
    DMX_CUDA(cudaMemcpy(ctx->d_disp, disp, len * sizeof(double), cudaMemcpyHostToDevice));
    return 0;
}
int dmx_set_tracer_diffusion(dmx_ctx* ctx, double D, double tortuosity)
{
    if (ctx->model != DMX_MODEL_TRACER) return fail(ctx, DMX_ERR_USAGE, "set_tracer_diffusion: not a tracer context");
    ctx->tracer_D = D;
    ctx->tracer_tau = tortuosity;
    return 0;
}
int dmx_set_tracer(dmx_ctx* ctx, int implicit)
{
    if (ctx->model != DMX_MODEL_TRACER) return fail(ctx, DMX_ERR_USAGE, "set_tracer: not a tracer ctx");
    ctx->tracer_implicit = implicit ? 1 : 0;
    return 0;
}
int dmx_assemble(dmx_ctx* ctx, int with_jacobian)
{
    if (!ctx->has_grid) return fail(ctx, DMX_ERR_USAGE, "assemble: no grid");
    DMX_CUDA(cudaSetDevice(ctx->device));
    DMX_CUDA(cudaMemsetAsync(ctx->d_flag, 0, sizeof(int), ctx->stream));
    if (int rc = launch_assemble(ctx, with_jacobian != 0)) return rc;
    // comm().min(succeeded) (assembly/fvassembler.hh:504-509): all ranks see the same flag
    int bad = 0;
    if (int rc = agree_flag(ctx, &bad)) return rc;
    if (bad) { ctx->err = "assemble: non-finite residual"; return DMX_STATUS_NONFINITE; }
    return 0;
}
int dmx_assemble_host(dmx_ctx* ctx, const double* cur, const double* prev, double* residual, double* jacobian)
{
    int rc;
    if ((rc = dmx_vec_upload(ctx, DMX_VEC_CUR, cur))) return rc;
    if (prev && (rc = dmx_vec_upload(ctx, DMX_VEC_PREV, prev))) return rc;
    if ((rc = dmx_assemble(ctx, jacobian != nullptr))) return rc;
    if (residual && (rc = dmx_vec_download(ctx, DMX_VEC_RESIDUAL, residual))) return rc;
    if (jacobian && (rc = dmx_jacobian_download(ctx, jacobian))) return rc;
    return 0;
}
int dmx_set_linear_solver(dmx_ctx* ctx, int solver, int restart)
{
    if (solver != DMX_SOLVER_BICGSTAB && solver != DMX_SOLVER_RESTARTED_GMRES && solver != DMX_SOLVER_CG) return fail(ctx, DMX_ERR_USAGE, "unknown linear solver");
    ctx->linear_solver = solver;
    ctx->gmres_restart = restart > 0 ? restart : 10;
    return 0;
}
int dmx_linear_solve(dmx_ctx* ctx, double reduction, int maxit, int preconditioner, int* iterations, double* achieved)
{
    if (!ctx->d_J) return fail(ctx, DMX_ERR_USAGE, "linear_solve: no pattern");
    DMX_CUDA(cudaSetDevice(ctx->device));
    return linear_solve(ctx, reduction, maxit, preconditioner, iterations, achieved);
}
int dmx_linear_solve_host(dmx_ctx* ctx, const double* values, double* x, const double* b, double reduction, int maxit, int preconditioner,
                          int* iterations, double* achieved)
{
    int rc;
    if ((rc = dmx_jacobian_upload(ctx, values))) return rc;
    if ((rc = dmx_vec_upload(ctx, DMX_VEC_DELTA, x))) return rc;
    if ((rc = dmx_vec_upload(ctx, DMX_VEC_RESIDUAL, b))) return rc;
    const int st = dmx_linear_solve(ctx, reduction, maxit, preconditioner, iterations, achieved);
    if (st < 0) return st;
    if ((rc = dmx_vec_download(ctx, DMX_VEC_DELTA, x))) return rc;
    return st;
}
int dmx_norm2(dmx_ctx* ctx, int vec, double* out)
{
    if (int rc = vec_ok(ctx, vec)) return rc;
    double s;
    if (int rc = dot(ctx, ctx->d_vec[vec], ctx->d_vec[vec], &s)) return rc;
    *out = std::sqrt(s);
    return 0;
}
int dmx_dot(dmx_ctx* ctx, int a, int b, double* out)
{
    if (int rc = vec_ok(ctx, a)) return rc;
    if (int rc = vec_ok(ctx, b)) return rc;
    return dot(ctx, ctx->d_vec[a], ctx->d_vec[b], out);
}
int dmx_newton_update(dmx_ctx* ctx, double* shift) { return newton_update(ctx, 1.0, shift); }
int dmx_advance_timestep(dmx_ctx* ctx) { return dmx_vec_copy(ctx, DMX_VEC_PREV, DMX_VEC_CUR); }
int dmx_reset_timestep(dmx_ctx* ctx) { return dmx_vec_copy(ctx, DMX_VEC_CUR, DMX_VEC_PREV); }

int dmx_newton_step(dmx_ctx* ctx, const dmx_newton_params* prm, int* linear_iterations, double* shift, float* ms_assemble,
                    float* ms_solve, float* ms_update)
{
    int rc;
    DMX_CUDA(cudaSetDevice(ctx->device));
    DMX_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
    // uLastIter = current iterate (newtonsolver.hh:985,1000)
    if ((rc = dmx_vec_copy(ctx, DMX_VEC_ULAST, DMX_VEC_CUR))) return rc;
    if ((rc = dmx_assemble(ctx, 1))) return rc;
    DMX_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
    DMX_CUDA(cudaMemsetAsync(ctx->d_vec[DMX_VEC_DELTA], 0, (size_t)ctx->n * ctx->b * sizeof(double), ctx->stream));   // deltaU = 0 (:1032)
    double red = 0;
    rc = linear_solve(ctx, prm->lin_reduction, prm->lin_maxit, prm->preconditioner, linear_iterations, &red);
    if (rc) return rc;
    DMX_CUDA(cudaEventRecord(ctx->ev[2], ctx->stream));
    if ((rc = newton_update(ctx, 1.0, shift))) return rc;
    DMX_CUDA(cudaEventRecord(ctx->ev[3], ctx->stream));
    DMX_CUDA(cudaEventSynchronize(ctx->ev[3]));
    float a = 0, s = 0, u = 0;
    cudaEventElapsedTime(&a, ctx->ev[0], ctx->ev[1]);
    cudaEventElapsedTime(&s, ctx->ev[1], ctx->ev[2]);
    cudaEventElapsedTime(&u, ctx->ev[2], ctx->ev[3]);
    if (ms_assemble) *ms_assemble = a;
    if (ms_solve) *ms_solve = s;
    if (ms_update) *ms_update = u;
    return 0;
}

int dmx_newton_step_host(dmx_ctx* ctx, double* u, const dmx_newton_params* prm, int* linear_iterations, double* shift)
{
    int rc;
    if ((rc = dmx_vec_upload(ctx, DMX_VEC_CUR, u))) return rc;
    const int st = dmx_newton_step(ctx, prm, linear_iterations, shift, nullptr, nullptr, nullptr);
    if (st) return st;
    return dmx_vec_download(ctx, DMX_VEC_CUR, u);
}
int dmx_timer_start(dmx_ctx* ctx)
{
    DMX_CUDA(cudaSetDevice(ctx->device));
    DMX_CUDA(cudaEventRecord(ctx->ev[4], ctx->stream));
    return 0;
}
int dmx_timer_stop(dmx_ctx* ctx, float* ms)
{
    DMX_CUDA(cudaSetDevice(ctx->device));
    DMX_CUDA(cudaEventRecord(ctx->ev[5], ctx->stream));
    DMX_CUDA(cudaEventSynchronize(ctx->ev[5]));
    DMX_CUDA(cudaEventElapsedTime(ms, ctx->ev[4], ctx->ev[5]));
    return 0;
}

// NewtonSolver::solveImpl_ (newtonsolver.hh:976-1072) with newtonProceed (:428-446), newtonConverged (:657-666, shift
// criterion), on the device-resident CUR/PREV.  Returns 0 if converged, DMX_STATUS_* otherwise.
// NewtonSolver::solveImpl_ (newtonsolver.hh:976-1072) with newtonBeginStep (:448-461), newtonUpdate (:543-566),
// lineSearchUpdate_ (:1154-1178), computeResidualReduction_ (:869-881), newtonConverged (:657-701), newtonProceed (:428-446)
int dmx_newton_solve(dmx_ctx* ctx, const dmx_newton_params* prm, dmx_newton_report* rep)
{
    std::memset(rep, 0, sizeof(*rep));
    const bool shiftCrit = prm->enable_shift_criterion != 0;
    const bool absResCrit = prm->enable_absolute_residual_criterion != 0;
    const bool resCrit = prm->enable_residual_criterion != 0 || absResCrit;
    if (!shiftCrit && !resCrit) return fail(ctx, DMX_ERR_USAGE, "Newton: at least one of the shift / residual criteria has to be enabled");
    const bool lineSearch = prm->use_line_search != 0;
    const bool needResidual = lineSearch || resCrit;
    int numSteps = 0;
    double shift = 0.0, lastShift = 0.0, reduction = 1.0, lastReduction = 1.0, residualNorm = 0.0, initialResidual = 0.0;
    bool converged = false;
    auto proceed = [&]() {
        if (numSteps < prm->min_steps) return true;
        else if (converged) return false;
        else if (numSteps >= prm->max_steps) return shiftCrit ? shift * 4.0 < lastShift : reduction * 4.0 < lastReduction;
        return true;
    };
    auto isConverged = [&]() {
        const bool resOk = absResCrit ? residualNorm <= prm->max_absolute_residual : reduction <= prm->residual_reduction;
        if (shiftCrit && !resCrit) return shift <= prm->max_relative_shift;
        if (!shiftCrit && resCrit) return resOk;
        if (prm->satisfy_residual_and_shift) return shift <= prm->max_relative_shift && resOk;
        return shift <= prm->max_relative_shift || resOk;
    };
    if (!needResidual) {
        // default path: one fused step per iteration
        while (proceed()) {
            lastShift = shift;
            int its = 0;
            float a = 0, s = 0, u = 0;
            const int rc = dmx_newton_step(ctx, prm, &its, &shift, &a, &s, &u);
            if (numSteps < 64) rep->linear_iterations[numSteps] = its;
            rep->linear_iterations_total += its;
            if (rc) { rep->newton_iterations = numSteps; rep->converged = 0; return rc; }
            rep->t_assemble += a * 1e-3; rep->t_solve += s * 1e-3; rep->t_update += u * 1e-3;
            if (numSteps < 64) { rep->shifts[numSteps] = shift; rep->relaxation[numSteps] = 1.0; }
            ++numSteps;
            converged = isConverged();
        }
    } else {
        DMX_CUDA(cudaSetDevice(ctx->device));
        const size_t bytes = (size_t)ctx->n * ctx->b * sizeof(double);
        while (proceed()) {
            lastShift = shift;
            lastReduction = numSteps == 0 ? 1.0 : reduction;
            int rc, its = 0;
            if ((rc = dmx_vec_copy(ctx, DMX_VEC_ULAST, DMX_VEC_CUR))) return rc;
            if ((rc = dmx_assemble(ctx, 1))) { rep->newton_iterations = numSteps; return rc; }
            if (numSteps == 0 && (rc = dmx_norm2(ctx, DMX_VEC_RESIDUAL, &initialResidual))) return rc;     // solveLinearSystem :495
            DMX_CUDA(cudaMemsetAsync(ctx->d_vec[DMX_VEC_DELTA], 0, bytes, ctx->stream));
            double red = 0;
            rc = linear_solve(ctx, prm->lin_reduction, prm->lin_maxit, prm->preconditioner, &its, &red);
            if (numSteps < 64) rep->linear_iterations[numSteps] = its;
            rep->linear_iterations_total += its;
            if (rc) { rep->newton_iterations = numSteps; rep->converged = 0; return rc; }
            double lambda = 1.0;
            while (true) {
                // uCurrentIter = uLastIter; axpy(-lambda, deltaU); residual and its norm at the trial point
                if ((rc = newton_update(ctx, lambda, &shift))) return rc;
                if ((rc = dmx_assemble(ctx, 0))) { rep->newton_iterations = numSteps; return rc; }
                if ((rc = dmx_norm2(ctx, DMX_VEC_RESIDUAL, &residualNorm))) return rc;
                reduction = residualNorm / initialResidual;
                if (!lineSearch || reduction < lastReduction || lambda <= prm->line_search_min_relaxation) break;
                lambda *= 0.5;
            }
            if (numSteps < 64) { rep->shifts[numSteps] = shift; rep->relaxation[numSteps] = lambda; }
            ++numSteps;
            converged = isConverged();
        }
    }
    rep->newton_iterations = numSteps;
    rep->converged = converged ? 1 : 0;
    rep->last_shift = shift;
    rep->last_reduction = reduction;
    rep->last_residual_norm = residualNorm;
    return converged ? 0 : DMX_STATUS_NOT_CONVERGED;
}
int dmx_newton_solve_host(dmx_ctx* ctx, double* u, const double* prev, const dmx_newton_params* prm, dmx_newton_report* rep)
{
    int rc;
    if ((rc = dmx_vec_upload(ctx, DMX_VEC_CUR, u))) return rc;
    if (prev && (rc = dmx_vec_upload(ctx, DMX_VEC_PREV, prev))) return rc;
    const int st = dmx_newton_solve(ctx, prm, rep);
    if (st < 0) return st;
    if ((rc = dmx_vec_download(ctx, DMX_VEC_CUR, u))) return rc;
    return st;
}

// ---- kernel-level entry points ----
int dmx_spmv(dmx_ctx* ctx, int x_vec, int y_vec)
{
    if (int rc = vec_ok(ctx, x_vec)) return rc;
    if (int rc = vec_ok(ctx, y_vec)) return rc;
    return launch_spmv(ctx, ctx->d_vec[x_vec], ctx->d_vec[y_vec]);
}
int dmx_ilu0_factor(dmx_ctx* ctx) { return ilu0_factor(ctx); }
int dmx_ilu0_apply(dmx_ctx* ctx, int d_vec, int v_vec)
{
    if (!ctx->ilu_valid) return fail(ctx, DMX_ERR_USAGE, "ilu0_apply before ilu0_factor");
    if (int rc = vec_ok(ctx, d_vec)) return rc;
    if (int rc = vec_ok(ctx, v_vec)) return rc;
    return ilu0_apply(ctx, ctx->d_vec[d_vec], ctx->d_vec[v_vec]);
}
int dmx_num_output_fields(const dmx_ctx* ctx) { return ctx->model == DMX_MODEL_2P ? 10 : 1; }
int dmx_output_fields(dmx_ctx* ctx, double* out)
{
    if (!ctx->has_grid) return fail(ctx, DMX_ERR_USAGE, "output_fields: set grid first");
    DMX_CUDA(cudaSetDevice(ctx->device));
    const size_t len = (size_t)dmx_num_output_fields(ctx) * ctx->n;
    double* d = nullptr;
    DMX_CUDA(cudaMalloc((void**)&d, len * sizeof(double)));
    int rc = launch_output_fields(ctx, d);
    if (!rc) {
        cudaError_t e = cudaMemcpyAsync(out, d, len * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) rc = fail(ctx, DMX_ERR_CUDA, std::string("output_fields: ") + cudaGetErrorString(e));
    }
    cudaFree(d);
    return rc;
}
int dmx_ssor_apply(dmx_ctx* ctx, int d_vec, int v_vec)
{
    if (!ctx->d_J) return fail(ctx, DMX_ERR_USAGE, "ssor_apply: no pattern");
    if (int rc = vec_ok(ctx, d_vec)) return rc;
    if (int rc = vec_ok(ctx, v_vec)) return rc;
    return ssor_apply(ctx, ctx->d_vec[d_vec], ctx->d_vec[v_vec]);
}
int dmx_set_amg_params(dmx_ctx* ctx, const dmx_amg_params* p)
{
    if (p->pre_steps < 0 || p->post_steps < 0 || p->pre_steps + p->post_steps < 1) return fail(ctx, DMX_ERR_USAGE, "AMG: preSteps + postSteps must be >= 1");
    const bool parmt = p->smoother == DMX_PRECOND_PARMT_JAC || p->smoother == DMX_PRECOND_PARMT_SOR || p->smoother == DMX_PRECOND_PARMT_SSOR;
    if (p->smoother != DMX_PRECOND_SSOR && p->smoother != DMX_PRECOND_ILU0 && !parmt)
        return fail(ctx, DMX_ERR_USAGE, "AMG smoother must be DMX_PRECOND_SSOR, DMX_PRECOND_ILU0 or DMX_PRECOND_PARMT_*");
    if (p->smoother_iterations < 1) return fail(ctx, DMX_ERR_USAGE, "AMG: smoother_iterations must be >= 1");
    if (!parmt && (p->smoother_iterations != 1 || p->smoother_relaxation != 1.0))
        return fail(ctx, DMX_ERR_USAGE, "AMG: the ssor / ilu smoothers run one iteration with relaxation 1 (factorised sweeps)");
    if (p->coarsest_cells < 1 || p->coarsest_steps < 1 || p->max_levels < 1) return fail(ctx, DMX_ERR_USAGE, "AMG: coarsest_cells, coarsest_steps, max_levels must be >= 1");
    if (p->coarsest_cells != ctx->amg_prm.coarsest_cells || p->max_levels != ctx->amg_prm.max_levels) ctx->amg_dirty = true;
    ctx->amg_prm = *p;
    return 0;
}
int dmx_amg_levels(dmx_ctx* ctx) { return amg_num_levels(ctx); }
int dmx_amg_level_cells(dmx_ctx* ctx, int level, int* cells)
{
    dmx_ctx* c = amg_level_ctx(ctx, level);
    if (!c) return fail(ctx, DMX_ERR_USAGE, "no such AMG level");
    for (int a = 0; a < 3; ++a) cells[a] = c->nc[a];
    return 0;
}
long long dmx_amg_level_nnz_blocks(dmx_ctx* ctx, int level)
{
    dmx_ctx* c = amg_level_ctx(ctx, level);
    return c ? c->nnzb : -1;
}
int dmx_amg_level_matrix(dmx_ctx* ctx, int level, double* values)
{
    dmx_ctx* c = amg_level_ctx(ctx, level);
    if (!c) return fail(ctx, DMX_ERR_USAGE, "no such AMG level");
    DMX_CUDA(cudaSetDevice(ctx->device));
    DMX_CUDA(cudaMemcpyAsync(values, c->d_J, (size_t)c->nnzb * c->b * c->b * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    DMX_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}
int dmx_set_preconditioner_params(dmx_ctx* ctx, int iterations, double relaxation)
{
    if (iterations < 1) return fail(ctx, DMX_ERR_USAGE, "LinearSolver.PreconditionerIterations must be >= 1");
    ctx->precond_iterations = iterations;
    ctx->precond_relaxation = relaxation;
    return 0;
}
int dmx_precond_apply(dmx_ctx* ctx, int preconditioner, int d_vec, int v_vec)
{
    if (!ctx->d_J) return fail(ctx, DMX_ERR_USAGE, "precond_apply: no pattern");
    if (int rc = vec_ok(ctx, d_vec)) return rc;
    if (int rc = vec_ok(ctx, v_vec)) return rc;
    DMX_CUDA(cudaSetDevice(ctx->device));
    if (int rc = precond_setup(ctx, preconditioner)) return rc;
    return precond_apply_local(ctx, preconditioner, ctx->d_vec[d_vec], ctx->d_vec[v_vec]);
}
int dmx_ilu0_download(dmx_ctx* ctx, double* values)
{
    if (!ctx->ilu_valid) return fail(ctx, DMX_ERR_USAGE, "no ILU factorisation");
    if (!ctx->ilu_bcrs_valid) {
        // structured path: the factors live in the sweep streams; rebuild the BCRS view from (J, Dinv) -- J must be unchanged
        if (!ctx->d_ilu) DMX_CUDA(cudaMalloc((void**)&ctx->d_ilu, (size_t)ctx->nnzb * ctx->b * ctx->b * sizeof(double)));
        if (int rc = sk_export_bcrs(ctx, ctx->d_ilu)) return rc;
        ctx->ilu_bcrs_valid = true;
    }
    DMX_CUDA(cudaMemcpyAsync(values, ctx->d_ilu, (size_t)ctx->nnzb * ctx->b * ctx->b * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    DMX_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}
int dmx_halo_exchange(dmx_ctx* ctx, int vec)
{
    if (int rc = vec_ok(ctx, vec)) return rc;
    if (ctx->nranks == 1) return 0;
    return halo_exchange(ctx, ctx->d_vec[vec]);
}

int dmx_debug_sweep_trace(dmx_ctx* ctx, long long* out6144) { return sk_trace_read(ctx, out6144); }

int dmx_time_kernel(dmx_ctx* ctx, int which, int reps, float* ms_avg)
{
    DMX_CUDA(cudaSetDevice(ctx->device));
    if (reps < 1) reps = 1;
    int rc = 0;
    if (which == 2 && !ctx->ilu_valid) return fail(ctx, DMX_ERR_USAGE, "time ILU apply: factor first");
    if (which == 0 && (rc = prepare(ctx))) return rc;
    DMX_CUDA(cudaStreamSynchronize(ctx->stream));
    DMX_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
    for (int i = 0; i < reps && !rc; ++i) {
        switch (which) {
            case 0: rc = launch_assemble(ctx, true); break;
            case 1: rc = launch_spmv(ctx, ctx->d_vec[DMX_VEC_WORK0], ctx->d_vec[DMX_VEC_WORK1]); break;
            case 2: rc = ilu0_apply(ctx, ctx->d_vec[DMX_VEC_WORK0], ctx->d_vec[DMX_VEC_WORK1]); break;
            case 3: rc = ilu0_factor(ctx); break;
            default: return fail(ctx, DMX_ERR_USAGE, "unknown kernel id");
        }
    }
    if (rc) return rc;
    DMX_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
    DMX_CUDA(cudaEventSynchronize(ctx->ev[1]));
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
    *ms_avg = ms / reps;
    (void)now_ms;
    return 0;
}

} // extern "C"

namespace dmx {

int make_child_ctx(dmx_ctx* parent, const int* gcells, const int* off, const int* nc, const int* own_lo, const int* own_hi, dmx_ctx** out)
{
    dmx_ctx* ctx = parent;       // for the error macros
    dmx_ctx* c = new dmx_ctx;
    c->device = parent->device;
    c->num_sms = parent->num_sms;
    c->stream = parent->stream;
    c->owns_stream = false;
    c->prof_parent = parent;
    if (parent->nranks > 1 && off) {
        c->rank = parent->rank; c->nranks = parent->nranks; c->nccl_comm = parent->nccl_comm;
        c->explicit_box = true;
        for (int a = 0; a < 3; ++a) {
            c->part[a] = parent->part[a]; c->pcoord[a] = parent->pcoord[a];
            c->xb_off[a] = off[a]; c->xb_nc[a] = nc[a]; c->xb_own_lo[a] = own_lo[a]; c->xb_own_hi[a] = own_hi[a];
        }
    }
    dmx_default_options(&c->opt);
    dmx_default_amg_params(&c->amg_prm);
    if (alloc_ctx_scratch(c)) { dmx_destroy(c); return fail(ctx, DMX_ERR_CUDA, "AMG: cannot allocate a level context"); }
    const double lower[3] = {0.0, 0.0, 0.0}, upper[3] = {1.0, 1.0, 1.0};
    const int rc = dmx_grid_structured(c, parent->b == 2 ? DMX_MODEL_2P : DMX_MODEL_1P, parent->dim, gcells, lower, upper);
    if (rc) {
        parent->err = "AMG level context: " + c->err;
        dmx_destroy(c);
        return rc;
    }
    parent->children.push_back(c);
    *out = c;
    return 0;
}
void destroy_child_ctx(dmx_ctx* child)
{
    if (!child) return;
    if (child->prof_parent) {
        auto& v = child->prof_parent->children;
        v.erase(std::remove(v.begin(), v.end(), child), v.end());
    }
    dmx_destroy(child);
}

} // namespace dmx
