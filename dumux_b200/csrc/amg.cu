// amg.cu -- aggregation AMG preconditioner on the structured grid hierarchy (AMGBiCGSTABIstlSolver / AMGCGIstlSolver,
// dumux/linear/istlsolvers.hh:716-757: Dune::Amg::AMG built by Dune::AMGCreator with dune-istl's default parameters).
//
// What dune-istl's AMG does [DUNE-ext, paamg/amg.hh, parameters.hh, solverfactory defaults]: aggregates of strongly connected
// unknowns (default aggregate size 4..8, "isotropic" with diameter 2), piecewise-constant prolongation, Galerkin coarse
// matrices P^T A P (blocks summed), coarse-grid correction damped by prolongationDampingFactor = 1.6, a V-cycle (gamma 1)
// with preSteps = postSteps = 2 smoothing steps of SeqSSOR (1 iteration, relaxation 1), where a smoothing step is
// "update = 0; smoother.apply(update, defect); lhs += update; defect -= A update", and a direct solve on the coarsest level.
// Its aggregation is a graph heuristic that cannot be restated from memory, so this is NOT a bit-for-bit port of dune's
// hierarchy.  It is the same cycle on the hierarchy the heuristic degenerates to on a 7-point structured box: aggregates
// = 2 x 2 x 2 cell blocks (8 unknowns, the upper end of dune's default range).  With those the coarse pattern stays the
// 7-point stencil of the coarse box, so EVERY level runs the structured kernels of this library (stencil SpMV, tile-wavefront
// sweeps); the hierarchy ends at <= coarsest_cells cells, where coarsest_steps smoothing steps stand in for the direct solve.
// The smoother is SeqSSOR in its factorised form M = (D + L) D^-1 (D + U) -- one forward and one backward block Gauss-Seidel
// sweep started from zero, which is what a smoothing step applies -- executed by the ILU sweep kernels with Dinv_i = A_ii^-1
// (L~ = L D^-1, U = D + U); SeqILU(0) is the alternative (dune: smoother "ilu").  oracle/amg_oracle.py restates this cycle
// operation by operation (same summation orders), so device and oracle agree bit for bit.
#include <algorithm>

#include "common.cuh"

namespace dmx {

struct AmgLevel {
    dmx_ctx* c = nullptr;        // level 0: the parent context; coarser levels: child contexts sharing its stream
    double *x = nullptr, *r = nullptr, *u = nullptr, *t = nullptr;
};

// Block-decomposed context (Grid.Partitioning, overlap 1): the hierarchy is the GLOBAL one and every level is decomposed like the
// grid.  Aggregates never cross a processor boundary (as in dune-istl's parallel AMG): along every axis the OWNED range
// [flo, fhi) of the local fine box is cut into pairs starting at its first cell (a last single cell if the length is odd); the
// overlap cell below / above belongs to the neighbour's last / first aggregate = the coarse overlap cell clo - 1 / chi.  Single
// domain: the owned range is the box, aggregate = index >> 1.  The cycle then contains the parallel pieces of dune's
// overlapping AMG: smoother = BlockPreconditioner (local sweeps, copyOwnerToAll), operator = local mv + project; restriction
// and prolongation need no exchange (children of an owned aggregate are owned; corrections are consistent on the overlap).
struct AggMap { int flo[3], fhi[3], clo[3]; };
__device__ __forceinline__ int agg1(int i, int flo, int fhi, int clo)
{
    if (i < flo) return clo - 1;
    if (i >= fhi) return clo + ((fhi - flo + 1) >> 1);
    return clo + ((i - flo) >> 1);
}
__device__ __forceinline__ void children1(int I, int flo, int fhi, int clo, int& c0, int& c1)
{
    const int chi = clo + ((fhi - flo + 1) >> 1);
    if (I < clo) { c0 = flo - 1; c1 = flo; }
    else if (I >= chi) { c0 = fhi; c1 = fhi + 1; }
    else { c0 = flo + 2 * (I - clo); c1 = min(c0 + 2, fhi); }
}
struct AmgState {
    std::vector<AmgLevel> levels;
    double* own[3] = {nullptr, nullptr, nullptr};      // r, u, t of level 0
    const double* rhs = nullptr;                       // right-hand side of the cycle being applied (level 0)
    // per-level device time of the cycle (exclusive of the coarser levels), recorded while dmx_profile is on
    struct LevelRec { int level; cudaEvent_t e[4]; };
    std::vector<LevelRec> pending;
    std::vector<double> level_ms;
    std::vector<long long> level_n;
};

// ---------------------------------------------------------------------------------------------
// Galerkin product for 2x2x2 box aggregates and piecewise-constant prolongation: coarse block (I,J) = sum of the fine blocks
// (i,j) with i in aggregate I, j in aggregate J.  One thread per coarse row.  Canonical summation order (restated by
// oracle.cpp orc_amg_galerkin): children in lexicographic order (x fastest), per child its entries in column order
// -z,-y,-x,diag,+x,+y,+z, every coarse slot accumulated from 0 in that visiting order.
// ---------------------------------------------------------------------------------------------
template <int B>
__global__ void __launch_bounds__(128) amg_galerkin_kernel(int fx, int fy, int fz, int cx, int cy, int cz, int dim, AggMap M,
                                                           const int* __restrict__ f_rowptr, const double* __restrict__ fA,
                                                           const int* __restrict__ c_rowptr, double* __restrict__ cA)
{
    constexpr int BB = B * B;
    const int Ic = blockIdx.x * blockDim.x + threadIdx.x;
    if (Ic >= cx * cy * cz) return;
    const int I = Ic % cx, J = (Ic / cx) % cy, K = Ic / (cx * cy);
    double acc[7][BB];
#pragma unroll
    for (int s = 0; s < 7; ++s)
#pragma unroll
        for (int q = 0; q < BB; ++q) acc[s][q] = 0.0;
    auto add = [&](int slot, const double* blk) {
#pragma unroll
        for (int s = 0; s < 7; ++s)
            if (s == slot) {
#pragma unroll
                for (int q = 0; q < BB; ++q) acc[s][q] += blk[q];
            }
    };
    int i0, i1, j0, j1, k0, k1;
    children1(I, M.flo[0], M.fhi[0], M.clo[0], i0, i1);
    children1(J, M.flo[1], M.fhi[1], M.clo[1], j0, j1);
    children1(K, M.flo[2], M.fhi[2], M.clo[2], k0, k1);
    for (int k = k0; k < k1; ++k)
        for (int j = j0; j < j1; ++j)
            for (int i = i0; i < i1; ++i) {
                const size_t row = (size_t)i + (size_t)fx * (j + (size_t)fy * k);
                const double* p = fA + (size_t)f_rowptr[row] * BB;
                if (dim > 2 && k > 0) { add(agg1(k - 1, M.flo[2], M.fhi[2], M.clo[2]) == K ? 3 : 0, p); p += BB; }
                if (dim > 1 && j > 0) { add(agg1(j - 1, M.flo[1], M.fhi[1], M.clo[1]) == J ? 3 : 1, p); p += BB; }
                if (i > 0) { add(agg1(i - 1, M.flo[0], M.fhi[0], M.clo[0]) == I ? 3 : 2, p); p += BB; }
                add(3, p); p += BB;
                if (i + 1 < fx) { add(agg1(i + 1, M.flo[0], M.fhi[0], M.clo[0]) == I ? 3 : 4, p); p += BB; }
                if (dim > 1 && j + 1 < fy) { add(agg1(j + 1, M.flo[1], M.fhi[1], M.clo[1]) == J ? 3 : 5, p); p += BB; }
                if (dim > 2 && k + 1 < fz) { add(agg1(k + 1, M.flo[2], M.fhi[2], M.clo[2]) == K ? 3 : 6, p); p += BB; }
            }
    double* out = cA + (size_t)c_rowptr[Ic] * BB;
    const bool ex[7] = {dim > 2 && K > 0, dim > 1 && J > 0, I > 0, true, I + 1 < cx, dim > 1 && J + 1 < cy, dim > 2 && K + 1 < cz};
#pragma unroll
    for (int s = 0; s < 7; ++s)
        if (ex[s]) {
#pragma unroll
            for (int q = 0; q < BB; ++q) out[q] = acc[s][q];
            out += BB;
        }
}

// restriction with P^T: r_c[I] = sum of the children's entries, lexicographic child order, from 0 (coarse overlap cells: their
// local children are fine overlap cells, whose defect is projected to zero)
template <int B>
__global__ void __launch_bounds__(256) amg_restrict_kernel(int fx, int fy, int fz, int cx, int cy, int cz, AggMap M, const double* __restrict__ rf,
                                                           double* __restrict__ rc)
{
    const int Ic = blockIdx.x * blockDim.x + threadIdx.x;
    if (Ic >= cx * cy * cz) return;
    const int I = Ic % cx, J = (Ic / cx) % cy, K = Ic / (cx * cy);
    double s[B];
#pragma unroll
    for (int e = 0; e < B; ++e) s[e] = 0.0;
    int i0, i1, j0, j1, k0, k1;
    children1(I, M.flo[0], M.fhi[0], M.clo[0], i0, i1);
    children1(J, M.flo[1], M.fhi[1], M.clo[1], j0, j1);
    children1(K, M.flo[2], M.fhi[2], M.clo[2], k0, k1);
    for (int k = k0; k < k1; ++k)
        for (int j = j0; j < j1; ++j)
            for (int i = i0; i < i1; ++i) {
                const size_t row = (size_t)i + (size_t)fx * (j + (size_t)fy * k);
#pragma unroll
                for (int e = 0; e < B; ++e) s[e] += rf[row * B + e];
            }
#pragma unroll
    for (int e = 0; e < B; ++e) rc[(size_t)Ic * B + e] = s[e];
}

// prolongation of the coarse correction, damped: u_i = damp * x_c[aggregate(i)];  x_i += u_i (x_i = u_i if `first`)
template <int B>
__global__ void __launch_bounds__(256) amg_prolong_kernel(int fx, int fy, int fz, int cx, int cy, AggMap M, double damp, const double* __restrict__ xc,
                                                          double* __restrict__ u, double* __restrict__ x, int first)
{
    const size_t row = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= (size_t)fx * fy * fz) return;
    const int i = (int)(row % fx), j = (int)((row / fx) % fy), k = (int)(row / ((size_t)fx * fy));
    const size_t Ic = (size_t)agg1(i, M.flo[0], M.fhi[0], M.clo[0]) +
                      (size_t)cx * (agg1(j, M.flo[1], M.fhi[1], M.clo[1]) + (size_t)cy * agg1(k, M.flo[2], M.fhi[2], M.clo[2]));
#pragma unroll
    for (int e = 0; e < B; ++e) {
        const double v = damp * xc[Ic * B + e];
        u[row * B + e] = v;
        x[row * B + e] = first ? v : x[row * B + e] + v;
    }
}

// after a smoothing step / the coarse-grid correction: lhs += update (lhs = update if `first`), defect -= A update (t = A update)
__global__ void __launch_bounds__(256) amg_update_kernel(size_t len, const double* __restrict__ u, const double* __restrict__ t, double* x,
                                                         double* r, int first, int with_x, int with_r)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (size_t)gridDim.x * blockDim.x) {
        if (with_x) x[i] = first ? u[i] : x[i] + u[i];
        if (with_r) r[i] -= t[i];
    }
}

static AmgState* amg_of(dmx_ctx* ctx) { return static_cast<AmgState*>(ctx->amg); }

void amg_free(dmx_ctx* ctx)
{
    AmgState* st = amg_of(ctx);
    if (!st) return;
    for (size_t l = 1; l < st->levels.size(); ++l) destroy_child_ctx(st->levels[l].c);
    for (auto& r : st->pending) for (cudaEvent_t e : r.e) if (e) cudaEventDestroy(e);
    for (double* p : st->own) if (p) cudaFree(p);
    delete st;
    ctx->amg = nullptr;
}

static AggMap agg_map(const dmx_ctx* f, const dmx_ctx* c)
{
    AggMap M;
    for (int a = 0; a < 3; ++a) { M.flo[a] = f->own_lo[a]; M.fhi[a] = f->own_hi[a]; M.clo[a] = c->own_lo[a]; }
    return M;
}

// hierarchy of boxes: pair the cells of every axis (per owned range) until <= coarsest_cells GLOBAL cells remain, no axis can be
// coarsened any further (every rank owns one cell along the partitioned axes) or max_levels is reached.  Every rank computes the
// owned sizes of all torus coordinates, so the boxes of a level need no communication.
static int amg_build(dmx_ctx* ctx)
{
    amg_free(ctx);
    if (!ctx->has_grid) return fail(ctx, DMX_ERR_USAGE, "AMG: needs a structured grid (dmx_grid_structured / dmx_grid_tensor)");
    AmgState* st = new AmgState;
    ctx->amg = st;
    const size_t len = (size_t)ctx->n * ctx->b;
    for (double*& p : st->own) DMX_CUDA(cudaMalloc((void**)&p, len * sizeof(double)));
    AmgLevel l0;
    l0.c = ctx; l0.r = st->own[0]; l0.u = st->own[1]; l0.t = st->own[2];
    st->levels.push_back(l0);
    const bool dist = ctx->nranks > 1;
    std::vector<int> sizes[3];                      // owned cells of every torus coordinate, per axis
    for (int a = 0; a < 3; ++a) {
        const int N = ctx->gcells[a], P = dist ? ctx->part[a] : 1;
        const int m = N / P, rem = N % P;
        for (int c = 0; c < P; ++c) sizes[a].push_back(c < P - rem ? m : m + 1);        // Yasp's partitioning, see finish_grid
    }
    int g[3] = {ctx->gcells[0], ctx->gcells[1], ctx->gcells[2]};
    while ((int)st->levels.size() < ctx->amg_prm.max_levels) {
        if ((long long)g[0] * g[1] * g[2] <= ctx->amg_prm.coarsest_cells) break;
        int cg[3], off[3], nc[3], olo[3], ohi[3];
        bool can = false;
        for (int a = 0; a < 3; ++a) {
            if (a < ctx->dim)
                for (int& s : sizes[a]) s = (s + 1) / 2;
            cg[a] = 0;
            int b0 = 0;
            const int mine = dist ? ctx->pcoord[a] : 0;
            for (int c = 0; c < (int)sizes[a].size(); ++c) {
                if (c == mine) b0 = cg[a];
                cg[a] += sizes[a][c];
            }
            const int b1 = b0 + sizes[a][mine];
            const int lo = std::max(0, b0 - 1), hi = std::min(cg[a], b1 + 1);
            off[a] = lo; nc[a] = hi - lo; olo[a] = b0 - lo; ohi[a] = b1 - lo;
            if (cg[a] != g[a]) can = true;
        }
        if (!can) break;
        dmx_ctx* child = nullptr;
        if (int rc = dist ? make_child_ctx(ctx, cg, off, nc, olo, ohi, &child) : make_child_ctx(ctx, cg, nullptr, nullptr, nullptr, nullptr, &child))
            return rc;
        for (int a = 0; a < 3; ++a) g[a] = cg[a];
        AmgLevel lv;
        lv.c = child;
        lv.x = child->d_vec[DMX_VEC_DELTA]; lv.r = child->d_vec[DMX_VEC_RESIDUAL];
        lv.u = child->d_vec[DMX_VEC_WORK0]; lv.t = child->d_vec[DMX_VEC_WORK1];
        st->levels.push_back(lv);
    }
    return 0;
}

int amg_num_levels(dmx_ctx* ctx) { return amg_of(ctx) ? (int)amg_of(ctx)->levels.size() : 0; }
dmx_ctx* amg_level_ctx(dmx_ctx* ctx, int level)
{
    AmgState* st = amg_of(ctx);
    return (st && level >= 0 && level < (int)st->levels.size()) ? st->levels[level].c : nullptr;
}

// Galerkin coarse matrices from the current Jacobian, then the smoother of every level
int amg_setup(dmx_ctx* ctx)
{
    AmgState* st = amg_of(ctx);
    if (!st || ctx->amg_dirty) {
        if (int rc = amg_build(ctx)) return rc;
        ctx->amg_dirty = false;
        st = amg_of(ctx);
    }
    const int smoother = ctx->amg_prm.smoother;
    const bool parmt = smoother == DMX_PRECOND_PARMT_JAC || smoother == DMX_PRECOND_PARMT_SOR || smoother == DMX_PRECOND_PARMT_SSOR;
    if (smoother != DMX_PRECOND_SSOR && smoother != DMX_PRECOND_ILU0 && !parmt)
        return fail(ctx, DMX_ERR_USAGE, "AMG smoother must be DMX_PRECOND_SSOR, DMX_PRECOND_ILU0 or DMX_PRECOND_PARMT_*");
    for (size_t l = 0; l < st->levels.size(); ++l) {
        dmx_ctx* c = st->levels[l].c;
        if (l > 0) {
            dmx_ctx* f = st->levels[l - 1].c;
            ProfScope ps(ctx, DMX_K_AMG);
            const int grid = (c->n + 127) / 128;
            const AggMap M = agg_map(f, c);
            if (ctx->b == 2)
                amg_galerkin_kernel<2><<<grid, 128, 0, ctx->stream>>>(f->nc[0], f->nc[1], f->nc[2], c->nc[0], c->nc[1], c->nc[2], ctx->dim, M,
                                                                      f->d_rowptr, f->d_J, c->d_rowptr, c->d_J);
            else
                amg_galerkin_kernel<1><<<grid, 128, 0, ctx->stream>>>(f->nc[0], f->nc[1], f->nc[2], c->nc[0], c->nc[1], c->nc[2], ctx->dim, M,
                                                                      f->d_rowptr, f->d_J, c->d_rowptr, c->d_J);
            DMX_CHECK_LAUNCH();
            c->jac_diagonal = false;
        }
        if (parmt) {
            // Dumux::ParMTJac / ParMTSOR / ParMTSSOR as the smoother: no factorisation, the colour sets of the level's pattern
            if (int rc = precond_setup(c, smoother)) {
                if (c != ctx) ctx->err = c->err;
                return rc;
            }
            continue;
        }
        // smoother set-up: the ILU machinery with Dinv = A_ii^-1 (factorised SSOR) or the ILU(0) recurrence
        c->ssor_factorised = (smoother == DMX_PRECOND_SSOR);
        const int rc = ilu0_factor(c);
        if (rc) {
            if (c != ctx) ctx->err = c->err;
            return rc;
        }
    }
    return 0;
}

// `rin`: where the defect of this step is read from (L.r, or the caller's right-hand side in the first step of level 0, which
// saves copying it); the updated defect always goes to L.r
static int amg_smooth_step(dmx_ctx* ctx, AmgLevel& L, bool first, bool need_defect, const double* rin = nullptr)
{
    dmx_ctx* c = L.c;
    const size_t len = (size_t)c->n * c->b;
    if (!rin) rin = L.r;
    const int sm = ctx->amg_prm.smoother;
    if (sm == DMX_PRECOND_SSOR || sm == DMX_PRECOND_ILU0) {
        if (int rc = ilu0_apply(c, rin, L.u)) return rc;                      // update = M^-1 defect (from update = 0)
    } else {
        if (int rc = parmt_apply_prm(c, sm, rin, L.u, ctx->amg_prm.smoother_iterations, ctx->amg_prm.smoother_relaxation)) return rc;
    }
    if (c->nranks > 1)
        if (int rc = halo_exchange(c, L.u)) return rc;                        // BlockPreconditioner::apply: copyOwnerToAll
    if (need_defect && spmv_update_supported(c))
        return launch_spmv_update(c, L.u, rin, L.r, L.x, true, first);        // lhs += update; defect -= A update, one pass
    if (rin != L.r) DMX_CUDA(cudaMemcpyAsync(L.r, rin, len * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    if (need_defect)
        if (int rc = launch_spmv(c, L.u, L.t)) return rc;                     // A update (block-decomposed: projected)
    ProfScope ps(ctx, DMX_K_AMG);
    const int grid = (int)std::min<size_t>((len + 255) / 256, 148 * 8);
    amg_update_kernel<<<grid, 256, 0, ctx->stream>>>(len, L.u, L.t, L.x, L.r, first ? 1 : 0, 1, need_defect ? 1 : 0);
    DMX_CHECK_LAUNCH();
    return 0;
}

static int amg_cycle_level(dmx_ctx* ctx, AmgState* st, int l, cudaEvent_t* ev);
static int amg_cycle(dmx_ctx* ctx, AmgState* st, int l)
{
    if (!ctx->prof_on) return amg_cycle_level(ctx, st, l, nullptr);
    AmgState::LevelRec rec;
    rec.level = l;
    for (cudaEvent_t& e : rec.e) { e = nullptr; cudaEventCreate(&e); }
    cudaEventRecord(rec.e[0], ctx->stream);
    const int rc = amg_cycle_level(ctx, st, l, rec.e);
    cudaEventRecord(rec.e[3], ctx->stream);
    st->pending.push_back(rec);
    return rc;
}
static int amg_cycle_level(dmx_ctx* ctx, AmgState* st, int l, cudaEvent_t* ev)
{
    AmgLevel& L = st->levels[l];
    dmx_ctx* c = L.c;
    const auto& prm = ctx->amg_prm;
    // level 0: the first smoothing step reads the caller's right-hand side in place
    const double* rhs0 = (l == 0) ? st->rhs : nullptr;
    if (l + 1 == (int)st->levels.size()) {
        // coarsest level: coarsest_steps smoothing steps instead of dune's direct solve
        const int ns = std::max(1, prm.coarsest_steps);
        for (int s = 0; s < ns; ++s)
            if (int rc = amg_smooth_step(ctx, L, s == 0, s + 1 < ns, s == 0 ? rhs0 : nullptr)) return rc;
        if (ev) { cudaEventRecord(ev[1], ctx->stream); cudaEventRecord(ev[2], ctx->stream); }
        return 0;
    }
    AmgLevel& C = st->levels[l + 1];
    dmx_ctx* cc = C.c;
    for (int s = 0; s < prm.pre_steps; ++s)
        if (int rc = amg_smooth_step(ctx, L, s == 0, true, s == 0 ? rhs0 : nullptr)) return rc;
    if (rhs0 && prm.pre_steps == 0)
        DMX_CUDA(cudaMemcpyAsync(L.r, rhs0, (size_t)c->n * c->b * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    {
        ProfScope ps(ctx, DMX_K_AMG);
        const int grid = (cc->n + 255) / 256;
        const AggMap M = agg_map(c, cc);
        if (ctx->b == 2) amg_restrict_kernel<2><<<grid, 256, 0, ctx->stream>>>(c->nc[0], c->nc[1], c->nc[2], cc->nc[0], cc->nc[1], cc->nc[2], M, L.r, C.r);
        else amg_restrict_kernel<1><<<grid, 256, 0, ctx->stream>>>(c->nc[0], c->nc[1], c->nc[2], cc->nc[0], cc->nc[1], cc->nc[2], M, L.r, C.r);
        DMX_CHECK_LAUNCH();
    }
    if (ev) cudaEventRecord(ev[1], ctx->stream);
    if (int rc = amg_cycle(ctx, st, l + 1)) return rc;
    if (ev) cudaEventRecord(ev[2], ctx->stream);
    {
        ProfScope ps(ctx, DMX_K_AMG);
        const int grid = (c->n + 255) / 256;
        const int first = prm.pre_steps == 0 ? 1 : 0;
        const AggMap M = agg_map(c, cc);
        if (ctx->b == 2)
            amg_prolong_kernel<2><<<grid, 256, 0, ctx->stream>>>(c->nc[0], c->nc[1], c->nc[2], cc->nc[0], cc->nc[1], M, prm.prolongation_damping, C.x, L.u, L.x, first);
        else
            amg_prolong_kernel<1><<<grid, 256, 0, ctx->stream>>>(c->nc[0], c->nc[1], c->nc[2], cc->nc[0], cc->nc[1], M, prm.prolongation_damping, C.x, L.u, L.x, first);
        DMX_CHECK_LAUNCH();
    }
    if (prm.post_steps > 0) {
        if (spmv_update_supported(c)) {
            if (int rc = launch_spmv_update(c, L.u, L.r, L.r, L.x, false, false)) return rc;      // defect -= A (coarse-grid correction)
        } else {
            if (int rc = launch_spmv(c, L.u, L.t)) return rc;
            ProfScope ps(ctx, DMX_K_AMG);
            const size_t len = (size_t)c->n * c->b;
            const int grid = (int)std::min<size_t>((len + 255) / 256, 148 * 8);
            amg_update_kernel<<<grid, 256, 0, ctx->stream>>>(len, L.u, L.t, L.x, L.r, 0, 0, 1);
            DMX_CHECK_LAUNCH();
        }
    }
    for (int s = 0; s < prm.post_steps; ++s)
        if (int rc = amg_smooth_step(ctx, L, false, s + 1 < prm.post_steps)) return rc;
    return 0;
}

// accumulated device milliseconds one level spent in the cycles timed so far (exclusive of the coarser levels) and the number of
// cycles; call after a synchronisation.  Reset by dmx_profile(ctx, 1).
int amg_level_profile(dmx_ctx* ctx, int level, double* ms, long long* n, bool reset)
{
    AmgState* st = amg_of(ctx);
    if (!st) return 0;
    st->level_ms.resize(st->levels.size(), 0.0);
    st->level_n.resize(st->levels.size(), 0);
    for (auto& r : st->pending) {
        float a = 0.f, b = 0.f;
        if (cudaEventElapsedTime(&a, r.e[0], r.e[1]) == cudaSuccess && cudaEventElapsedTime(&b, r.e[2], r.e[3]) == cudaSuccess &&
            r.level < (int)st->level_ms.size()) {
            st->level_ms[r.level] += a + b;
            st->level_n[r.level]++;
        }
        for (cudaEvent_t e : r.e) cudaEventDestroy(e);
    }
    st->pending.clear();
    if (reset) { std::fill(st->level_ms.begin(), st->level_ms.end(), 0.0); std::fill(st->level_n.begin(), st->level_n.end(), 0); return 0; }
    if (level < 0 || level >= (int)st->levels.size()) return DMX_ERR_USAGE;
    if (ms) *ms = st->level_ms[level];
    if (n) *n = st->level_n[level];
    return 0;
}

// v = AMG(J)(d): one V-cycle from v = 0
int amg_apply(dmx_ctx* ctx, const double* d, double* v)
{
    AmgState* st = amg_of(ctx);
    if (!st) return fail(ctx, DMX_ERR_USAGE, "AMG apply before set-up");
    AmgLevel& L0 = st->levels[0];
    const size_t len = (size_t)ctx->n * ctx->b;
    (void)len;
    st->rhs = d;
    L0.x = v;
    return amg_cycle(ctx, st, 0);
}

} // namespace dmx
