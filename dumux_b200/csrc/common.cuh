// common.cuh -- context object and helpers shared by the translation units of libdumux_b200.so
#pragma once
#include <cuda_runtime.h>

#include <cstdio>
#include <string>
#include <vector>

#include "../../include/dumux_b200.h"
#include "physics.cuh"

#define DMX_NUM_VECS 7
#define DMX_MAX_REGIONS 8
#define DMX_NUM_KCLASS 8
#define DMX_ERR_CUDA (-1)
#define DMX_ERR_USAGE (-2)
#define DMX_ERR_NCCL (-3)

namespace dmx {

// Everything the assembly kernels need, passed by value (fits the 4 KB kernel-parameter space).
struct AsmParams {
    int model, b, dim;
    int nc[3];
    int n;
    // options
    int enable_gravity, fd_method, stationary, nd;
    double gravity, upwind_weight, base_eps, mag[2], dt, extrusion;
    // geometry, per axis a: width[a][i], gf_lo[a][i] (>0), gf_hi[a][i] (>0): |(x_f - x_c)/|x_f - x_c|^2|
    const double* width[3];
    const double* gf_lo[3];
    const double* gf_hi[3];
    // cell fields
    const double* K;          // scalar permeability; with a diagonal tensor: its entry along the gravity axis
    const double* Kaxis[3];   // diagonal tensor: K_aa per axis, null = scalar K
    const double* phi;
    const int* region;
    const double* q;          // may be null
    const double* tij[3];     // transmissibility of the + face of each cell along axis a
    const double* vf;         // tracer: frozen volume fluxes [n][2*dim]
    const double* disp;       // tracer: n.D.n of the mechanical dispersion tensor at every face [n][2*dim], null = off
    int tracer_implicit;
    double tracer_D, tracer_tau;          // Fick's law: binary diffusion coefficient, constant tortuosity
    // fluids
    double rho[2], mu[2];
    double rmu[2], rdt;       // correctly rounded 1/mu, 1/dt (host IEEE division) for div_by
    int zchunk;               // layers per CTA of the tile kernel
    int tabulated;
    FluidTable table;
    const MaterialLaw* laws;
    int nlaws;
    // boundary data per side: type[nf], neumann[nf*b], Dirichlet state p[nf*2], up[nf*2], rho[nf*2]
    const int* bc_type[6];
    const double* bc_neumann[6];
    const double* bc_p[6];
    const double* bc_up[6];
    const double* bc_rho[6];
    // state
    const double* cur;
    const double* prev;
    // outputs
    const int* rowptr;
    double* residual;
    double* jac;
    int* flag_nonfinite;
};

} // namespace dmx

struct dmx_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    long long launches = 0;
    int num_sms = 148;

    // distributed
    int rank = 0, nranks = 1;
    void* nccl_comm = nullptr;
    // Grid.Partitioning (io/grid/gridmanager_yasp.hh:194-203): ranks per axis, 0 = not set (slabs along the last axis);
    // pcoord = this rank's position in the process torus (x fastest); own_lo/own_hi = owned index range per axis (LOCAL indices)
    int part_req[3] = {0, 0, 0};
    int part[3] = {1, 1, 1}, pcoord[3] = {0, 0, 0};
    int own_lo[3] = {0, 0, 0}, own_hi[3] = {1, 1, 1};
    // copyOwnerToAll neighbours (up to 26 directions): owned cells that lie in the neighbour's overlap go out, my overlap
    // cells the neighbour owns come in.  Regions are boxes in local indices; buffers are packed in neighbour order.
    struct HaloNb { int rank; int slo[3], rlo[3], size[3]; long long off, count; bool contiguous; };
    std::vector<HaloNb> halo_nb;
    long long halo_total = 0;           // doubles per exchange (all neighbours)

    // grid
    int model = 0, b = 0, dim = 0;
    int gcells[3] = {1, 1, 1};
    int nc[3] = {1, 1, 1}, off[3] = {0, 0, 0};
    int n = 0;
    long long nnzb = 0;
    bool has_grid = false;
    std::vector<double> xn[3];
    double* d_geom = nullptr;
    const double *d_width[3] = {}, *d_gflo[3] = {}, *d_gfhi[3] = {};

    dmx_options opt;
    bool prepared = false;

    // cell fields
    std::vector<double> h_K, h_phi;
    std::vector<int> h_region;
    double *d_K = nullptr, *d_phi = nullptr, *d_q = nullptr;
    double* d_Kaxis[3] = {nullptr, nullptr, nullptr};      // diagonal permeability tensor (dmx_set_permeability_diagonal), null = scalar
    int* d_region = nullptr;
    double* d_tij[3] = {nullptr, nullptr, nullptr};
    double* d_vf = nullptr;             // tracer: frozen volume fluxes [n][2*dim]
    double* d_disp = nullptr;           // tracer: dispersion-tensor entries at the faces [n][2*dim] (dmx_set_tracer_dispersion)
    double* d_law_rec = nullptr;        // DiffMethod::analytic (2p): per-cell material-law record [6][n], allocated on first use
    int tracer_implicit = 0;
    double tracer_D = 0.0, tracer_tau = 0.5;

    std::vector<dmx::MaterialLaw> laws;
    dmx::MaterialLaw* d_laws = nullptr;
    double rho[2] = {1000.0, 1460.0}, mu[2] = {1e-3, 5.7e-4};
    bool tabulated = false;
    dmx::FluidTable h_table{};          // host copy of table (host pointers)
    std::vector<double> h_tab_pmin, h_tab_pmax, h_tab_rho, h_tab_mu;
    dmx::FluidTable d_table{};          // device pointers
    double* d_tab_buf = nullptr;

    // boundary
    std::vector<int> h_bc_type[6];
    std::vector<double> h_bc_val[6];
    int* d_bc_type[6] = {};
    double *d_bc_neumann[6] = {}, *d_bc_p[6] = {}, *d_bc_up[6] = {}, *d_bc_rho[6] = {};

    // pattern + matrix
    std::vector<int> h_rowptr, h_colidx;
    int *d_rowptr = nullptr, *d_colidx = nullptr, *d_diag = nullptr;
    double *d_J = nullptr, *d_ilu = nullptr;
    bool ilu_valid = false;
    bool jac_diagonal = false;        // every off-diagonal block of d_J is exactly zero (explicit tracer assembly): ILU0 = D^-1
    bool ilu_bcrs_valid = false;      // d_ilu holds the factorised BCRS values (generic path; structured path: on download only)

    // vectors
    double* d_vec[DMX_NUM_VECS] = {};
    double *d_rt = nullptr, *d_p = nullptr, *d_v = nullptr, *d_t = nullptr, *d_y = nullptr, *d_z = nullptr, *d_dinv = nullptr;

    // ILU0 level schedule (rows sorted by level; level_ptr on host)
    int *d_lrows = nullptr, *d_urows = nullptr;
    std::vector<int> l_ptr, u_ptr;
    int *d_lptr = nullptr, *d_uptr = nullptr;
    unsigned int* d_barrier = nullptr;
    int linear_solver = DMX_SOLVER_BICGSTAB, gmres_restart = 10;     // dmx_set_linear_solver
    int precond_iterations = 1;       // LinearSolver.PreconditionerIterations (ParMT* smoothers)
    double precond_relaxation = 1.0;  // LinearSolver.PreconditionerRelaxation
    int* d_color_rows = nullptr;      // ParMTSOR/SSOR: rows sorted by colour (computeColorsForMatrixSweep_)
    std::vector<int> color_ptr;       // host: first row of every colour in d_color_rows (+ end)
    double* d_xold = nullptr;         // ParMTJac: previous iterate
    double* d_gm = nullptr;           // GMRes basis: restart+1 vectors, w, defect
    int gm_vectors = 0;
    void* skew = nullptr;             // SkewState of ilu_structured.cu (structured-grid ILU sweeps), null: generic kernels
    int sk_tile = 16;                 // tile edge of the sweep kernels serving this context (16 or 8, see ilu_structured.cu)
    bool ssor_factorised = false;     // the "ILU" machinery holds SeqSSOR in factorised form: Dinv_i = A_ii^-1 instead of the ILU(0) recurrence
    void* amg = nullptr;              // AmgState of amg.cu
    dmx_amg_params amg_prm;
    bool amg_dirty = true;            // hierarchy has to be (re)built: new grid or new parameters
    bool owns_stream = true;          // false: a level context of an AMG hierarchy running on its parent's stream (and communicator)
    // level contexts of a block-decomposed hierarchy get their box handed in instead of Yasp's partitioning formula
    bool explicit_box = false;
    int xb_off[3] = {0, 0, 0}, xb_nc[3] = {1, 1, 1}, xb_own_lo[3] = {0, 0, 0}, xb_own_hi[3] = {1, 1, 1};
    dmx_ctx* prof_parent = nullptr;   // kernel-class timers of a level context are booked on the parent
    std::vector<dmx_ctx*> children;   // level contexts (for the launch count)

    // reductions
    double* d_partials = nullptr;
    double* d_scalars = nullptr;
    double* h_scalars = nullptr;     // pinned
    int* d_flag = nullptr;
    int* h_flag = nullptr;           // pinned
    unsigned char* d_owner = nullptr;  // 1 = owner (null in single-GPU mode)

    // halo (distributed)
    double *d_send = nullptr, *d_recv = nullptr;
    double* d_gather = nullptr;      // all-gathered partial sums of a scalar product [nranks][<= 8]

    cudaEvent_t ev[6] = {};

    // per-kernel-class device timers (dmx_profile_*): CUDA-event pairs recorded on the ctx stream around the launches
    bool prof_on = false;
    struct ProfRec { int cls; cudaEvent_t a, b; };
    std::vector<ProfRec> prof_pending;
    std::vector<cudaEvent_t> prof_pool;
    double prof_ms[DMX_NUM_KCLASS] = {};
    long long prof_n[DMX_NUM_KCLASS] = {};
};

namespace dmx {

inline int fail(dmx_ctx* c, int code, const std::string& msg)
{
    if (c) c->err = msg;
    return code;
}

#define DMX_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess)                                                                     \
            return dmx::fail(ctx, DMX_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
    } while (0)

#define DMX_CHECK_LAUNCH()                                                                          \
    do {                                                                                            \
        ctx->launches++;                                                                            \
        cudaError_t e__ = cudaGetLastError();                                                       \
        if (e__ != cudaSuccess)                                                                     \
            return dmx::fail(ctx, DMX_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(e__)); \
    } while (0)

// Times everything launched on the ctx stream during its lifetime as one unit of kernel class `cls` (no-op unless
// dmx_profile(ctx, 1)).  Events are resolved by prof_drain() at a point where the stream is known to be idle.
struct ProfScope {
    dmx_ctx* c;
    int cls;
    cudaEvent_t a = nullptr, b = nullptr;
    static cudaEvent_t get(dmx_ctx* c)
    {
        cudaEvent_t e = nullptr;
        if (!c->prof_pool.empty()) { e = c->prof_pool.back(); c->prof_pool.pop_back(); }
        else cudaEventCreate(&e);
        return e;
    }
    ProfScope(dmx_ctx* ctx, int k) : c(ctx->prof_parent ? ctx->prof_parent : ctx), cls(k)
    {
        if (!c->prof_on) return;
        a = get(c); b = get(c);
        cudaEventRecord(a, c->stream);
    }
    ~ProfScope()
    {
        if (!a) return;
        cudaEventRecord(b, c->stream);
        c->prof_pending.push_back({cls, a, b});
    }
};
// call only right after a stream synchronisation
inline void prof_drain(dmx_ctx* c)
{
    for (auto& r : c->prof_pending) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) { c->prof_ms[r.cls] += ms; c->prof_n[r.cls]++; }
        c->prof_pool.push_back(r.a);
        c->prof_pool.push_back(r.b);
    }
    c->prof_pending.clear();
}

#ifdef __CUDACC__
// FieldMatrix::invert for 1x1 / 2x2 blocks (dune-common densematrix.hh, the 2x2 closed form); false: singular
template <int B>
__device__ __forceinline__ bool invert_block(double* A)
{
    if (B == 1) {
        if (A[0] == 0.0) return false;
        A[0] = 1.0 / A[0];
        return true;
    } else {
        double detinv = A[0] * A[3] - A[1] * A[2];
        if (detinv == 0.0 || detinv != detinv) return false;
        detinv = 1.0 / detinv;
        const double temp = A[0];
        A[0] = A[3] * detinv;
        A[1] = -A[1] * detinv;
        A[2] = -A[2] * detinv;
        A[3] = temp * detinv;
        return true;
    }
}

// grid-wide barrier for the persistent level loops (all CTAs co-resident: cooperative launch)
__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int& epoch)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        epoch += gridDim.x;
        __threadfence();
        atomicAdd(counter, 1u);
        while (*((volatile unsigned int*)counter) < epoch) { }
        __threadfence();
    }
    __syncthreads();
}

#endif

// implemented in assembly.cu
int prepare(dmx_ctx* ctx);
int launch_assemble(dmx_ctx* ctx, bool with_jacobian);
int launch_volume_flux(dmx_ctx* ctx, double* d_out);
int launch_output_fields(dmx_ctx* ctx, double* d_out);
// implemented in linalg.cu
int build_diag(dmx_ctx* ctx);
int build_level_schedule(dmx_ctx* ctx);
int launch_spmv(dmx_ctx* ctx, const double* x, double* y);
int launch_spmv_local(dmx_ctx* ctx, const double* x, double* y);     // without the owner projection
bool spmv_update_supported(const dmx_ctx* ctx);
int launch_spmv_update(dmx_ctx* ctx, const double* u, const double* r_in, double* r_out, double* x, bool with_x, bool first);
int ilu0_factor(dmx_ctx* ctx);
int ilu0_factor_bcrs(dmx_ctx* ctx);
int ilu0_apply(dmx_ctx* ctx, const double* d, double* v);
int ssor_apply(dmx_ctx* ctx, const double* d, double* v);
int block_jacobi_setup(dmx_ctx* ctx);
int block_jacobi_apply(dmx_ctx* ctx, const double* d, double* v);
int parmt_apply(dmx_ctx* ctx, int precond, const double* d, double* v);
int parmt_apply_prm(dmx_ctx* ctx, int precond, const double* d, double* v, int iterations, double relaxation);
int precond_apply_local(dmx_ctx* ctx, int precond, const double* d, double* v);
int precond_setup(dmx_ctx* ctx, int precond);
int dot(dmx_ctx* ctx, const double* a, const double* b, double* out);
int bicgstab(dmx_ctx* ctx, double reduction, int maxit, int precond, int* iterations, double* achieved);
int gmres(dmx_ctx* ctx, double reduction, int maxit, int restart, int precond, int* iterations, double* achieved);
int linear_solve(dmx_ctx* ctx, double reduction, int maxit, int precond, int* iterations, double* achieved);
int newton_update(dmx_ctx* ctx, double lambda, double* shift);
// implemented in ilu_structured.cu
int sk_setup(dmx_ctx* ctx);
void sk_free(dmx_ctx* ctx);
int sk_factor(dmx_ctx* ctx);
int sk_export_bcrs(dmx_ctx* ctx, double* out);
int sk_apply(dmx_ctx* ctx, const double* d, double* v);
int sk_trace_read(dmx_ctx* ctx, long long* out);
// implemented in amg.cu
void amg_free(dmx_ctx* ctx);
int amg_setup(dmx_ctx* ctx);
int amg_apply(dmx_ctx* ctx, const double* d, double* v);
int amg_num_levels(dmx_ctx* ctx);
dmx_ctx* amg_level_ctx(dmx_ctx* ctx, int level);
int amg_level_profile(dmx_ctx* ctx, int level, double* ms, long long* n, bool reset);
// implemented in api.cu: level contexts of a hierarchy (own grid, matrix and vectors; the parent's device and stream)
// `gcells`: global cells of the level; distributed parent: off / nc = local box (overlap included), own_lo / own_hi = owned range
// in local indices (all per axis); single domain: pass nullptr for the four
int make_child_ctx(dmx_ctx* parent, const int* gcells, const int* off, const int* nc, const int* own_lo, const int* own_hi, dmx_ctx** out);
void destroy_child_ctx(dmx_ctx* child);
// implemented in dist.cu
int nccl_init(dmx_ctx* ctx, const void* uid);
int nccl_get_unique_id(void* out);
int nccl_destroy(dmx_ctx* ctx);
int halo_exchange(dmx_ctx* ctx, double* v);
int allreduce_sum(dmx_ctx* ctx, double* d_buf, int count);
int allreduce_max(dmx_ctx* ctx, double* d_buf, int count);
int allreduce_min_int(dmx_ctx* ctx, int* d_buf, int count);
int allreduce_max_int(dmx_ctx* ctx, int* d_buf, int count);
int agree_flag(dmx_ctx* ctx, int* flag_out);

} // namespace dmx
