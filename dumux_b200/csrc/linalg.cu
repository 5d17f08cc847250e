// linalg.cu -- BCRS block SpMV, block ILU(0), fused BLAS-1 and the BiCGSTAB driver.
//
// Replaces what DuMux delegates to dune-istl through IstlIterativeLinearSolver
// (dumux/linear/istlsolvers.hh:273,457-464,535-568; alias ILUBiCGSTABIstlSolver :636-642):
//   Dune::BCRSMatrix::mv            -> bcrs_spmv_kernel<b>
//   Dune::SeqILU(n=0,w=1)           -> ilu0_factor_kernel<b> / ilu0_lower_kernel<b> / ilu0_upper_kernel<b>
//                                      (level-scheduled: rows of one dependency level run in parallel, per-row
//                                      operation order identical to ILU::blockILU0Decomposition / blockILUBacksolve)
//   Dune::BiCGSTABSolver::apply     -> bicgstab() below, same operation sequence and stopping rules
//   SeqScalarProduct / OverlappingSchwarzScalarProduct -> dot kernels (fixed-shape, deterministic; owner-masked)
// Compiled with -fmad=false so per-row sums match the sequential CPU arithmetic bit for bit.
#include <algorithm>
#include <cfloat>
#include <cmath>

#include "common.cuh"

namespace dmx {

static constexpr int RED_BLOCKS = 1184;   // 148 SMs x 8
static constexpr int RED_THREADS = 256;

// ---------------------------------------------------------------------------------------------
// SpMV: one thread per scalar row (b threads per block row); per-row sum in column order from 0
// ---------------------------------------------------------------------------------------------
template <int B>
__global__ void __launch_bounds__(256) bcrs_spmv_kernel(int n, const int* __restrict__ rowptr, const int* __restrict__ colidx,
                                                        const double* __restrict__ A, const double* __restrict__ x,
                                                        double* __restrict__ y, const unsigned char* __restrict__ owner)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int row = (int)(t / B);
    const int e = (int)(t % B);
    if (row >= n) return;
    double acc = 0.0;
    const int k0 = rowptr[row], k1 = rowptr[row + 1];
    for (int k = k0; k < k1; ++k) {
        const int c = colidx[k];
        if (B == 2) {
            const double2 a = *reinterpret_cast<const double2*>(A + ((size_t)k * 4 + e * 2));
            const double2 xv = *reinterpret_cast<const double2*>(x + (size_t)c * 2);
            acc += a.x * xv.x;
            acc += a.y * xv.y;
        } else {
            acc += A[k] * x[c];
        }
    }
    if (owner && !owner[row]) acc = 0.0;   // OverlappingSchwarzOperator: project() zeroes non-owner rows
    y[(size_t)row * B + e] = acc;
}

// Structured-grid variant: same BCRS values, same per-row summation order (columns ascending = -z,-y,-x,diag,+x,+y,+z among
// the existing neighbours), but the column indices come from the grid instead of colidx -- 28 of the 288 bytes per 2x2 block
// row are not read.  Launched as (x-blocks, ny, nz): no integer division per thread.
template <int B>
__global__ void __launch_bounds__(128) stencil_spmv_kernel(int nx, int ny, int nz, int dim, const int* __restrict__ rowptr,
                                                           const double* __restrict__ A, const double* __restrict__ x,
                                                           double* __restrict__ y, const unsigned char* __restrict__ owner)
{
    const int tx = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = tx / B, e = tx % B;
    const int j = blockIdx.y, k = blockIdx.z;
    if (i >= nx) return;
    const int row = i + nx * (j + ny * k);
    const int sx = 1, sy = nx, sz = nx * ny;
    int kpos = rowptr[row];
    double acc = 0.0;
    auto term = [&](int c) {
        if (B == 2) {
            const double2 a = *reinterpret_cast<const double2*>(A + ((size_t)kpos * 4 + e * 2));
            const double2 xv = *reinterpret_cast<const double2*>(x + (size_t)c * 2);
            acc += a.x * xv.x;
            acc += a.y * xv.y;
        } else {
            acc += A[kpos] * x[c];
        }
        ++kpos;
    };
    if (dim > 2 && k > 0) term(row - sz);
    if (dim > 1 && j > 0) term(row - sy);
    if (i > 0) term(row - sx);
    term(row);
    if (i + 1 < nx) term(row + sx);
    if (dim > 1 && j + 1 < ny) term(row + sy);
    if (dim > 2 && k + 1 < nz) term(row + sz);
    if (owner && !owner[row]) acc = 0.0;   // OverlappingSchwarzOperator: project() zeroes non-owner rows
    y[(size_t)row * B + e] = acc;
}

// The defect update of a multigrid smoothing step fused into the operator application: t = A u (row sum in column order from 0,
// projected to the owner rows in a block-decomposed context), r_out = r_in - t, and lhs += u (lhs = u if `first`) -- what
// "lhs += update; defect -= A update" of dune-istl's AMG does [DUNE-ext paamg/amg.hh], without writing and re-reading A u.
// Every output entry is produced by the thread that computed the row sum; u is only read.  Same operations, same bits.
template <int B>
__global__ void __launch_bounds__(128) stencil_spmv_update_kernel(int nx, int ny, int nz, int dim, const int* __restrict__ rowptr,
                                                                  const double* __restrict__ A, const double* __restrict__ u,
                                                                  const double* r_in, double* r_out, double* x, int with_x, int first,
                                                                  const unsigned char* __restrict__ owner)
{
    const int tx = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = tx / B, e = tx % B;
    const int j = blockIdx.y, k = blockIdx.z;
    if (i >= nx) return;
    const int row = i + nx * (j + ny * k);
    const int sx = 1, sy = nx, sz = nx * ny;
    int kpos = rowptr[row];
    double acc = 0.0;
    auto term = [&](int c) {
        if (B == 2) {
            const double2 a = *reinterpret_cast<const double2*>(A + ((size_t)kpos * 4 + e * 2));
            const double2 xv = *reinterpret_cast<const double2*>(u + (size_t)c * 2);
            acc += a.x * xv.x;
            acc += a.y * xv.y;
        } else {
            acc += A[kpos] * u[c];
        }
        ++kpos;
    };
    if (dim > 2 && k > 0) term(row - sz);
    if (dim > 1 && j > 0) term(row - sy);
    if (i > 0) term(row - sx);
    term(row);
    if (i + 1 < nx) term(row + sx);
    if (dim > 1 && j + 1 < ny) term(row + sy);
    if (dim > 2 && k + 1 < nz) term(row + sz);
    if (owner && !owner[row]) acc = 0.0;
    const size_t o = (size_t)row * B + e;
    r_out[o] = r_in[o] - acc;
    if (with_x) {
        const double uv = u[o];
        x[o] = first ? uv : x[o] + uv;
    }
}

// BCRSMatrix::mv for a Jacobian whose off-diagonal blocks are all exactly zero (explicit tracer step): only the diagonal
// block contributes to the row sum (the skipped terms are 0 * x_j = 0), so six of the seven blocks per row are not read
template <int B>
__global__ void __launch_bounds__(256) diag_spmv_kernel(size_t len, const int* __restrict__ diag, const double* __restrict__ A,
                                                        const double* __restrict__ x, double* __restrict__ y,
                                                        const unsigned char* __restrict__ owner)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= len) return;
    const size_t row = t / B;
    const int e = (int)(t % B);
    const size_t kpos = (size_t)diag[row];
    double acc = 0.0;
#pragma unroll
    for (int c = 0; c < B; ++c) acc += A[kpos * B * B + e * B + c] * x[row * B + c];
    if (owner && !owner[row]) acc = 0.0;
    y[t] = acc;
}

static int launch_spmv_owner(dmx_ctx* ctx, const double* x, double* y, const unsigned char* owner)
{
    ProfScope ps(ctx, DMX_K_SPMV);
    if (ctx->jac_diagonal && ctx->d_diag) {
        const size_t len = (size_t)ctx->n * ctx->b;
        const unsigned grid = (unsigned)((len + 255) / 256);
        if (ctx->b == 2) diag_spmv_kernel<2><<<grid, 256, 0, ctx->stream>>>(len, ctx->d_diag, ctx->d_J, x, y, owner);
        else diag_spmv_kernel<1><<<grid, 256, 0, ctx->stream>>>(len, ctx->d_diag, ctx->d_J, x, y, owner);
        DMX_CHECK_LAUNCH();
        return 0;
    }
    if (ctx->has_grid && ctx->nc[1] <= 65535 && ctx->nc[2] <= 65535) {
        const int bs = 128;
        const dim3 grid((unsigned)((ctx->nc[0] * ctx->b + bs - 1) / bs), (unsigned)ctx->nc[1], (unsigned)ctx->nc[2]);
        if (ctx->b == 2)
            stencil_spmv_kernel<2><<<grid, bs, 0, ctx->stream>>>(ctx->nc[0], ctx->nc[1], ctx->nc[2], ctx->dim, ctx->d_rowptr, ctx->d_J, x, y, owner);
        else
            stencil_spmv_kernel<1><<<grid, bs, 0, ctx->stream>>>(ctx->nc[0], ctx->nc[1], ctx->nc[2], ctx->dim, ctx->d_rowptr, ctx->d_J, x, y, owner);
        DMX_CHECK_LAUNCH();
        return 0;
    }
    const long long threads = (long long)ctx->n * ctx->b;
    const int bs = 256;
    const int grid = (int)((threads + bs - 1) / bs);
    if (ctx->b == 2)
        bcrs_spmv_kernel<2><<<grid, bs, 0, ctx->stream>>>(ctx->n, ctx->d_rowptr, ctx->d_colidx, ctx->d_J, x, y, owner);
    else
        bcrs_spmv_kernel<1><<<grid, bs, 0, ctx->stream>>>(ctx->n, ctx->d_rowptr, ctx->d_colidx, ctx->d_J, x, y, owner);
    DMX_CHECK_LAUNCH();
    return 0;
}
// r_out = r_in - A u (projected), optionally lhs (+)= u, in one pass; false: this context has no structured SpMV (caller falls back)
bool spmv_update_supported(const dmx_ctx* ctx) { return ctx->has_grid && ctx->nc[1] <= 65535 && ctx->nc[2] <= 65535 && !ctx->jac_diagonal; }
int launch_spmv_update(dmx_ctx* ctx, const double* u, const double* r_in, double* r_out, double* x, bool with_x, bool first)
{
    ProfScope ps(ctx, DMX_K_SPMV);
    const int bs = 128;
    const dim3 grid((unsigned)((ctx->nc[0] * ctx->b + bs - 1) / bs), (unsigned)ctx->nc[1], (unsigned)ctx->nc[2]);
    if (ctx->b == 2)
        stencil_spmv_update_kernel<2><<<grid, bs, 0, ctx->stream>>>(ctx->nc[0], ctx->nc[1], ctx->nc[2], ctx->dim, ctx->d_rowptr, ctx->d_J, u, r_in, r_out, x,
                                                                    with_x ? 1 : 0, first ? 1 : 0, ctx->d_owner);
    else
        stencil_spmv_update_kernel<1><<<grid, bs, 0, ctx->stream>>>(ctx->nc[0], ctx->nc[1], ctx->nc[2], ctx->dim, ctx->d_rowptr, ctx->d_J, u, r_in, r_out, x,
                                                                    with_x ? 1 : 0, first ? 1 : 0, ctx->d_owner);
    DMX_CHECK_LAUNCH();
    return 0;
}
// OverlappingSchwarzOperator::apply: local A.mv, then project (non-owner rows zero)
int launch_spmv(dmx_ctx* ctx, const double* x, double* y) { return launch_spmv_owner(ctx, x, y, ctx->d_owner); }
// the plain local product (inside a sequential preconditioner)
int launch_spmv_local(dmx_ctx* ctx, const double* x, double* y) { return launch_spmv_owner(ctx, x, y, nullptr); }

// ---------------------------------------------------------------------------------------------
// reductions: fixed grid, per-block partials, single-block final pass in fixed order -> deterministic
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
    return v;
}
template <bool MAX>
__device__ __forceinline__ double block_reduce(double v)
{
    __shared__ double sh[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = MAX ? warp_max(v) : warp_sum(v);
    __syncthreads();
    if (lane == 0) sh[wid] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    v = (threadIdx.x < nw) ? sh[threadIdx.x] : 0.0;
    if (wid == 0) v = MAX ? warp_max(v) : warp_sum(v);
    return v;
}

// up to 3 simultaneous dot products: out[q] = sum_i a_q[i]*b_q[i] over owner entries
template <int NQ>
__global__ void __launch_bounds__(RED_THREADS) dot_kernel(size_t len, int b, const double* a0, const double* b0, const double* a1,
                                                          const double* b1, const double* a2, const double* b2,
                                                          const unsigned char* owner, double* partials)
{
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (size_t)gridDim.x * blockDim.x) {
        if (owner && !owner[i / b]) continue;
        s0 += a0[i] * b0[i];
        if (NQ > 1) s1 += a1[i] * b1[i];
        if (NQ > 2) s2 += a2[i] * b2[i];
    }
    s0 = block_reduce<false>(s0);
    if (NQ > 1) s1 = block_reduce<false>(s1);
    if (NQ > 2) s2 = block_reduce<false>(s2);
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = s0;
        if (NQ > 1) partials[gridDim.x + blockIdx.x] = s1;
        if (NQ > 2) partials[2 * gridDim.x + blockIdx.x] = s2;
    }
}
template <bool MAX>
__global__ void __launch_bounds__(RED_THREADS) final_reduce_kernel(int nblocks, int nq, const double* partials, double* out)
{
    for (int q = 0; q < nq; ++q) {
        double s = 0.0;
        for (int i = threadIdx.x; i < nblocks; i += blockDim.x) s = MAX ? fmax(s, partials[q * nblocks + i]) : s + partials[q * nblocks + i];
        s = block_reduce<MAX>(s);
        if (threadIdx.x == 0) out[q] = s;
        __syncthreads();
    }
}

// x += alpha*y ; r += -alpha*v ; partial ||r||^2   (the two axpy + norm of a BiCGSTAB half step)
__global__ void __launch_bounds__(RED_THREADS) axpy2_norm_kernel(size_t len, int b, double alpha, const double* y, const double* v,
                                                                 double* x, double* r, const unsigned char* owner, double* partials)
{
    double s = 0.0;
    const double malpha = -alpha;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (size_t)gridDim.x * blockDim.x) {
        x[i] += alpha * y[i];
        const double ri = r[i] + malpha * v[i];
        r[i] = ri;
        if (!owner || owner[i / b]) s += ri * ri;
    }
    s = block_reduce<false>(s);
    if (threadIdx.x == 0) partials[blockIdx.x] = s;
}
// first half step of BiCGSTAB: r += -alpha*v ; partial ||r||^2.  The matching x += alpha*y is deferred to the second half
// step (x is not needed in between), which saves one read and one write of x per iteration.
// Device-resident Krylov scalars (ctx->d_scalars): [0..2] reduction results, [3] rho, [4] alpha, [5] omega, [6] h.  alpha and
// omega are derived on the device right after their dot products (derive_kernel), so the update kernels can be queued
// without a host round trip; the host reads all of them together with the residual norm, once per half step.
enum { SC_RED = 0, SC_RHO = 3, SC_ALPHA = 4, SC_OMEGA = 5, SC_H = 6 };
__global__ void derive_kernel(double* sc, int op)
{
    if (op == 0) { sc[SC_H] = sc[0]; sc[SC_ALPHA] = sc[SC_RHO] / sc[0]; }       // alpha = rho / <rt, v>
    else if (op == 1) sc[SC_OMEGA] = sc[0] / sc[1];                               // omega = <t, r> / <t, t>
    else if (op == 2) sc[SC_RHO] = sc[1];                                         // rho of the next iteration (fused update)
    else sc[SC_RHO] = sc[0];                                                      // rho from the stand-alone dot
}
__global__ void __launch_bounds__(RED_THREADS) axpy_r_norm_kernel(size_t len, int b, const double* sc, const double* v, double* r,
                                                                  const unsigned char* owner, double* partials)
{
    double s = 0.0;
    const double malpha = -sc[SC_ALPHA];
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (size_t)gridDim.x * blockDim.x) {
        const double ri = r[i] + malpha * v[i];
        r[i] = ri;
        if (!owner || owner[i / b]) s += ri * ri;
    }
    s = block_reduce<false>(s);
    if (threadIdx.x == 0) partials[blockIdx.x] = s;
}
// second half step: x = (x + alpha*y) + omega*z ; r += -omega*t ; partial ||r||^2 and partial <rt, r> (the next iteration's rho)
__global__ void __launch_bounds__(RED_THREADS) axpy3_norm_dot_kernel(size_t len, int b, const double* sc, const double* y,
                                                                     const double* z, const double* t, const double* rt, double* x,
                                                                     double* r, const unsigned char* owner, double* partials)
{
    double s = 0.0, d = 0.0;
    const double alpha = sc[SC_ALPHA], omega = sc[SC_OMEGA];
    const double momega = -omega;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (size_t)gridDim.x * blockDim.x) {
        double xi = x[i];
        xi += alpha * y[i];
        xi += omega * z[i];
        x[i] = xi;
        const double ri = r[i] + momega * t[i];
        r[i] = ri;
        if (!owner || owner[i / b]) { s += ri * ri; d += rt[i] * ri; }
    }
    s = block_reduce<false>(s);
    d = block_reduce<false>(d);
    if (threadIdx.x == 0) { partials[blockIdx.x] = s; partials[gridDim.x + blockIdx.x] = d; }
}
// x += alpha*y (only when the solver stops after a first half step)
__global__ void __launch_bounds__(256) axpy_kernel(size_t len, double alpha, const double* y, double* x)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (size_t)gridDim.x * blockDim.x) x[i] += alpha * y[i];
}
// p = ((p + (-omega)*v) * beta) + r
__global__ void __launch_bounds__(256) p_update_kernel(size_t len, double beta, double omega, const double* r, const double* v, double* p)
{
    const double momega = -omega;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (size_t)gridDim.x * blockDim.x) {
        double pi = p[i] + momega * v[i];
        pi *= beta;
        pi += r[i];
        p[i] = pi;
    }
}
// r = b - A x given t = A x ; rt = r ; partial ||r||^2
__global__ void __launch_bounds__(RED_THREADS) residual_init_kernel(size_t len, int b, const double* rhs, const double* Ax, double* r,
                                                                    double* rt, const unsigned char* owner, double* partials)
{
    double s = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (size_t)gridDim.x * blockDim.x) {
        // OverlappingSchwarzOperator::applyscaleadd(-1, x, r) projects r: non-owner entries are zero
        const double ri = (!owner || owner[i / b]) ? rhs[i] - Ax[i] : 0.0;
        r[i] = ri;
        rt[i] = ri;
        s += ri * ri;
    }
    s = block_reduce<false>(s);
    if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

// u = uLast + (-lambda)*delta ; shift = max_i |u - uLast| / max(1, |u + uLast|/2)   (newtonsolver.hh:111-129,556-557,1163)
__global__ void __launch_bounds__(RED_THREADS) newton_update_kernel(size_t len, int b, double mlambda, const double* uLast, const double* delta, double* u,
                                                                    const unsigned char* owner, double* partials)
{
    double s = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (size_t)gridDim.x * blockDim.x) {
        const double ul = uLast[i];
        const double un = ul + mlambda * delta[i];
        u[i] = un;
        const double sh = fabs(un - ul) / fmax(1.0, fabs(un + ul) * 0.5);
        if (!owner || owner[i / b]) s = fmax(s, sh);
    }
    s = block_reduce<true>(s);
    if (threadIdx.x == 0) partials[blockIdx.x] = s;
}
static int reduce_to_host(dmx_ctx* ctx, int nq, bool max, double* out)
{
    if (max) final_reduce_kernel<true><<<1, RED_THREADS, 0, ctx->stream>>>(RED_BLOCKS, nq, ctx->d_partials, ctx->d_scalars);
    else final_reduce_kernel<false><<<1, RED_THREADS, 0, ctx->stream>>>(RED_BLOCKS, nq, ctx->d_partials, ctx->d_scalars);
    DMX_CHECK_LAUNCH();
    if (ctx->nranks > 1) {
        if (int rc = max ? allreduce_max(ctx, ctx->d_scalars, nq) : allreduce_sum(ctx, ctx->d_scalars, nq)) return rc;
    }
    DMX_CUDA(cudaMemcpyAsync(ctx->h_scalars, ctx->d_scalars, nq * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    DMX_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->prof_pending.size() > 512) prof_drain(ctx);
    for (int q = 0; q < nq; ++q) out[q] = ctx->h_scalars[q];
    return 0;
}

// reduction of the per-block partials into d_scalars[0..nq) (+ all-reduce), then a derived scalar; no host round trip
static int reduce_on_device(dmx_ctx* ctx, int nq, int derive_op)
{
    final_reduce_kernel<false><<<1, RED_THREADS, 0, ctx->stream>>>(RED_BLOCKS, nq, ctx->d_partials, ctx->d_scalars);
    DMX_CHECK_LAUNCH();
    if (ctx->nranks > 1) {
        if (int rc = allreduce_sum(ctx, ctx->d_scalars, nq)) return rc;
    }
    if (derive_op >= 0) {
        derive_kernel<<<1, 1, 0, ctx->stream>>>(ctx->d_scalars, derive_op);
        DMX_CHECK_LAUNCH();
    }
    return 0;
}
static int read_scalars(dmx_ctx* ctx, double* out8)
{
    DMX_CUDA(cudaMemcpyAsync(ctx->h_scalars, ctx->d_scalars, 8 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    DMX_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->prof_pending.size() > 512) prof_drain(ctx);
    for (int q = 0; q < 8; ++q) out8[q] = ctx->h_scalars[q];
    return 0;
}

int dot(dmx_ctx* ctx, const double* a, const double* b, double* out)
{
    const size_t len = (size_t)ctx->n * ctx->b;
    ProfScope ps(ctx, DMX_K_BLAS1);
    dot_kernel<1><<<RED_BLOCKS, RED_THREADS, 0, ctx->stream>>>(len, ctx->b, a, b, nullptr, nullptr, nullptr, nullptr, ctx->d_owner, ctx->d_partials);
    DMX_CHECK_LAUNCH();
    return reduce_to_host(ctx, 1, false, out);
}
int newton_update(dmx_ctx* ctx, double lambda, double* shift)
{
    const size_t len = (size_t)ctx->n * ctx->b;
    ProfScope ps(ctx, DMX_K_BLAS1);
    newton_update_kernel<<<RED_BLOCKS, RED_THREADS, 0, ctx->stream>>>(len, ctx->b, -lambda, ctx->d_vec[DMX_VEC_ULAST], ctx->d_vec[DMX_VEC_DELTA],
                                                                      ctx->d_vec[DMX_VEC_CUR], ctx->d_owner, ctx->d_partials);
    DMX_CHECK_LAUNCH();
    return reduce_to_host(ctx, 1, true, shift);
}

// ---------------------------------------------------------------------------------------------
// block ILU(0), level scheduled.  Level of row i (lower) = 1 + max level of rows j<i in its pattern; for the
// 7-point stencil that is the hyperplane i+j+k.  Rows of one level are independent.
// ---------------------------------------------------------------------------------------------
// position of the diagonal block of every row (uploaded with the pattern); the level schedule itself is built lazily, only when
// the generic (non-structured) ILU kernels are actually used
int build_diag(dmx_ctx* ctx)
{
    const int n = ctx->n;
    const std::vector<int>& rp = ctx->h_rowptr;
    const std::vector<int>& ci = ctx->h_colidx;
    std::vector<int> diag(n, -1);
    for (int i = 0; i < n; ++i) {
        for (int k = rp[i]; k < rp[i + 1]; ++k)
            if (ci[k] == i) { diag[i] = k; break; }
        if (diag[i] < 0) return fail(ctx, DMX_ERR_USAGE, "pattern without diagonal entry");
    }
    if (ctx->d_diag) cudaFree(ctx->d_diag);
    ctx->d_diag = nullptr;
    DMX_CUDA(cudaMalloc((void**)&ctx->d_diag, diag.size() * sizeof(int)));
    DMX_CUDA(cudaMemcpy(ctx->d_diag, diag.data(), diag.size() * sizeof(int), cudaMemcpyHostToDevice));
    ctx->l_ptr.clear();
    ctx->u_ptr.clear();
    ctx->color_ptr.clear();
    ctx->ilu_valid = false;
    ctx->ilu_bcrs_valid = false;
    return 0;
}

int build_level_schedule(dmx_ctx* ctx)
{
    const int n = ctx->n;
    const std::vector<int>& rp = ctx->h_rowptr;
    const std::vector<int>& ci = ctx->h_colidx;
    std::vector<int> lev(n, 0);
    int nl = 0;
    for (int i = 0; i < n; ++i) {
        int l = 0;
        for (int k = rp[i]; k < rp[i + 1]; ++k) {
            const int j = ci[k];
            if (j < i) l = std::max(l, lev[j] + 1);
        }
        lev[i] = l;
        nl = std::max(nl, l + 1);
    }
    auto bucket = [&](const std::vector<int>& level, int nlev, std::vector<int>& ptr, std::vector<int>& rows) {
        ptr.assign(nlev + 1, 0);
        for (int i = 0; i < n; ++i) ptr[level[i] + 1]++;
        for (int l = 0; l < nlev; ++l) ptr[l + 1] += ptr[l];
        rows.resize(n);
        std::vector<int> cur(ptr.begin(), ptr.end() - 1);
        for (int i = 0; i < n; ++i) rows[cur[level[i]]++] = i;
    };
    std::vector<int> lrows, urows;
    bucket(lev, nl, ctx->l_ptr, lrows);
    // upper: dependencies on rows j > i
    std::vector<int> ulev(n, 0);
    int nu = 0;
    for (int i = n - 1; i >= 0; --i) {
        int l = 0;
        for (int k = rp[i]; k < rp[i + 1]; ++k)
            if (ci[k] > i) l = std::max(l, ulev[ci[k]] + 1);
        ulev[i] = l;
        nu = std::max(nu, l + 1);
    }
    bucket(ulev, nu, ctx->u_ptr, urows);
    auto up = [&](int** d, const std::vector<int>& h) -> int {
        if (*d) cudaFree(*d);
        DMX_CUDA(cudaMalloc((void**)d, h.size() * sizeof(int)));
        DMX_CUDA(cudaMemcpy(*d, h.data(), h.size() * sizeof(int), cudaMemcpyHostToDevice));
        return 0;
    };
    if (int rc = up(&ctx->d_lrows, lrows)) return rc;
    if (int rc = up(&ctx->d_urows, urows)) return rc;
    if (int rc = up(&ctx->d_lptr, ctx->l_ptr)) return rc;
    if (int rc = up(&ctx->d_uptr, ctx->u_ptr)) return rc;
    return 0;
}

template <int B>
__device__ __forceinline__ void factor_row(int i, const int* rowptr, const int* colidx, const int* diag, double* A, int* flag)
{
    constexpr int BB = B * B;
    const int iend = rowptr[i + 1];
    const int kd = diag[i];
    for (int kij = rowptr[i]; kij < kd; ++kij) {
        const int j = colidx[kij];
        const int kjj = diag[j];
        // A_ij <- A_ij * A_jj^-1   (rightmultiply: C[r][c] = sum_k A[r][k]*B[k][c], sum started from 0)
        double L[BB], D[BB];
#pragma unroll
        for (int q = 0; q < BB; ++q) { L[q] = A[(size_t)kij * BB + q]; D[q] = __ldcg(A + (size_t)kjj * BB + q); }
        double C[BB];
#pragma unroll
        for (int r = 0; r < B; ++r)
#pragma unroll
            for (int c = 0; c < B; ++c) {
                double s = 0.0;
#pragma unroll
                for (int k = 0; k < B; ++k) s += L[r * B + k] * D[k * B + c];
                C[r * B + c] = s;
            }
#pragma unroll
        for (int q = 0; q < BB; ++q) A[(size_t)kij * BB + q] = C[q];
        // A_ik -= A_ij * A_jk for k > j present in both rows
        int ik = kij + 1, jk = kjj + 1;
        const int jend = rowptr[j + 1];
        while (ik < iend && jk < jend) {
            const int ci = colidx[ik], cj = colidx[jk];
            if (ci == cj) {
                double* T = A + (size_t)ik * BB;
                const double* U = A + (size_t)jk * BB;
#pragma unroll
                for (int r = 0; r < B; ++r)
#pragma unroll
                    for (int c = 0; c < B; ++c) {
                        double t = T[r * B + c];
#pragma unroll
                        for (int k = 0; k < B; ++k) t -= C[r * B + k] * __ldcg(U + k * B + c);
                        T[r * B + c] = t;
                    }
                ++ik; ++jk;
            } else if (ci < cj) ++ik;
            else ++jk;
        }
    }
    if (!invert_block<B>(A + (size_t)kd * BB)) atomicOr(flag, 1);
}

template <int B>
__global__ void __launch_bounds__(128) ilu0_factor_kernel(int nlev, const int* lptr, const int* lrows, const int* rowptr, const int* colidx,
                                                          const int* diag, double* A, int* flag, unsigned int* barrier)
{
    unsigned int epoch = 0;
    for (int l = 0; l < nlev; ++l) {
        const int r0 = lptr[l], r1 = lptr[l + 1];
        for (int q = r0 + blockIdx.x * blockDim.x + threadIdx.x; q < r1; q += gridDim.x * blockDim.x)
            factor_row<B>(lrows[q], rowptr, colidx, diag, A, flag);
        if (l + 1 < nlev) grid_barrier(barrier, epoch);
    }
}

// v_i = d_i - sum_{j<i} A_ij v_j
template <int B>
__global__ void __launch_bounds__(128) ilu0_lower_kernel(int nlev, const int* lptr, const int* lrows, const int* rowptr, const int* colidx,
                                                         const int* diag, const double* A, const double* d, double* v, unsigned int* barrier)
{
    constexpr int BB = B * B;
    unsigned int epoch = 0;
    for (int l = 0; l < nlev; ++l) {
        const int r0 = lptr[l], r1 = lptr[l + 1];
        for (int q = r0 + blockIdx.x * blockDim.x + threadIdx.x; q < r1; q += gridDim.x * blockDim.x) {
            const int i = lrows[q];
            double rhs[B];
#pragma unroll
            for (int e = 0; e < B; ++e) rhs[e] = d[(size_t)i * B + e];
            const int kd = diag[i];
            for (int k = rowptr[i]; k < kd; ++k) {
                const int c = colidx[k];
                const double* vc = v + (size_t)c * B;
                double xv[B];
#pragma unroll
                for (int e = 0; e < B; ++e) xv[e] = __ldcg(vc + e);
#pragma unroll
                for (int r = 0; r < B; ++r)
#pragma unroll
                    for (int cc = 0; cc < B; ++cc) rhs[r] -= A[(size_t)k * BB + r * B + cc] * xv[cc];
            }
#pragma unroll
            for (int e = 0; e < B; ++e) __stcg(v + (size_t)i * B + e, rhs[e]);
        }
        if (l + 1 < nlev) grid_barrier(barrier, epoch);
    }
}
// v_i = A_ii^-1 (v_i - sum_{j>i} A_ij v_j)
template <int B>
__global__ void __launch_bounds__(128) ilu0_upper_kernel(int nlev, const int* uptr, const int* urows, const int* rowptr, const int* colidx,
                                                         const int* diag, const double* A, double* v, unsigned int* barrier)
{
    constexpr int BB = B * B;
    unsigned int epoch = 0;
    for (int l = 0; l < nlev; ++l) {
        const int r0 = uptr[l], r1 = uptr[l + 1];
        for (int q = r0 + blockIdx.x * blockDim.x + threadIdx.x; q < r1; q += gridDim.x * blockDim.x) {
            const int i = urows[q];
            double rhs[B];
#pragma unroll
            for (int e = 0; e < B; ++e) rhs[e] = __ldcg(v + (size_t)i * B + e);
            const int kd = diag[i];
            const int kend = rowptr[i + 1];
            for (int k = kd + 1; k < kend; ++k) {
                const int c = colidx[k];
                double xv[B];
#pragma unroll
                for (int e = 0; e < B; ++e) xv[e] = __ldcg(v + (size_t)c * B + e);
#pragma unroll
                for (int r = 0; r < B; ++r)
#pragma unroll
                    for (int cc = 0; cc < B; ++cc) rhs[r] -= A[(size_t)k * BB + r * B + cc] * xv[cc];
            }
            double out[B];
#pragma unroll
            for (int r = 0; r < B; ++r) {
                double s = 0.0;
#pragma unroll
                for (int cc = 0; cc < B; ++cc) s += A[(size_t)kd * BB + r * B + cc] * rhs[cc];
                out[r] = s;
            }
#pragma unroll
            for (int e = 0; e < B; ++e) __stcg(v + (size_t)i * B + e, out[e]);
        }
        if (l + 1 < nlev) grid_barrier(barrier, epoch);
    }
}


// ---------------------------------------------------------------------------------------------
// Dune::SeqSSOR(1 iteration, w = 1) (what SSORCGIstlSolver / SSORBiCGSTABIstlSolver use, dumux/linear/istlsolvers.hh:686-714):
// one forward and one backward block Gauss-Seidel sweep (dune-istl gsetc.hh bsorf / bsorb).  Per row, columns ascending INCLUDING
// the diagonal: rhs = d_i - sum_j A_ij x_j (newest values), v = A_ii^-1 rhs (FieldMatrix::solve), x_i += v.  Level-scheduled like
// the generic ILU0 sweeps: rows of one level are independent, lower (upper) neighbours belong to earlier levels and upper
// (lower) ones to later levels, so "newest value" is well defined with one grid barrier per level.
// ---------------------------------------------------------------------------------------------
template <int B>
__device__ __forceinline__ void gs_row(int i, const int* rowptr, const int* colidx, const int* diag, const double* A, const double* d, double* x)
{
    constexpr int BB = B * B;
    double rhs[B];
#pragma unroll
    for (int e = 0; e < B; ++e) rhs[e] = d[(size_t)i * B + e];
    const int kend = rowptr[i + 1];
    for (int k = rowptr[i]; k < kend; ++k) {
        const int c = colidx[k];
        double xv[B];
#pragma unroll
        for (int e = 0; e < B; ++e) xv[e] = __ldcg(x + (size_t)c * B + e);
#pragma unroll
        for (int r = 0; r < B; ++r)
#pragma unroll
            for (int cc = 0; cc < B; ++cc) rhs[r] -= A[(size_t)k * BB + r * B + cc] * xv[cc];
    }
    const double* D = A + (size_t)diag[i] * BB;
    double v[B];
    if (B == 1) v[0] = rhs[0] / D[0];
    else {
        double detinv = D[0] * D[3] - D[1] * D[2];
        detinv = 1 / detinv;
        v[0] = detinv * (D[3] * rhs[0] - D[1] * rhs[B - 1]);
        v[B - 1] = detinv * (D[0] * rhs[B - 1] - D[2] * rhs[0]);
    }
#pragma unroll
    for (int e = 0; e < B; ++e) __stcg(x + (size_t)i * B + e, __ldcg(x + (size_t)i * B + e) + 1.0 * v[e]);
}
template <int B>
__global__ void __launch_bounds__(128) gs_sweep_kernel(int nlev, const int* lptr, const int* lrows, const int* rowptr, const int* colidx,
                                                       const int* diag, const double* A, const double* d, double* x, unsigned int* barrier)
{
    unsigned int epoch = 0;
    for (int l = 0; l < nlev; ++l) {
        const int r0 = lptr[l], r1 = lptr[l + 1];
        for (int q = r0 + blockIdx.x * blockDim.x + threadIdx.x; q < r1; q += gridDim.x * blockDim.x)
            gs_row<B>(lrows[q], rowptr, colidx, diag, A, d, x);
        if (l + 1 < nlev) grid_barrier(barrier, epoch);
    }
}

static int coop_grid(dmx_ctx* ctx, const void* kernel, int threads)
{
    int perSm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kernel, threads, 0);
    if (perSm < 1) perSm = 1;
    if (perSm > 4) perSm = 4;
    return perSm * ctx->num_sms;
}

template <class... Args>
static int coop_launch(dmx_ctx* ctx, void (*kernel)(Args...), int threads, Args... args)
{
    const int grid = coop_grid(ctx, (const void*)kernel, threads);
    DMX_CUDA(cudaMemsetAsync(ctx->d_barrier, 0, sizeof(unsigned int), ctx->stream));
    void* params[] = {(void*)&args...};
    DMX_CUDA(cudaLaunchCooperativeKernel((const void*)kernel, dim3(grid), dim3(threads), params, 0, ctx->stream));
    ctx->launches++;
    return 0;
}

// SeqSSOR in factorised form on a generic pattern, in the layout of the ILU sweeps: L~_ij = A_ij A_jj^-1 (j < i), A_ii^-1 on the
// diagonal, U_ij = A_ij (j > i) -- what ilu_skew_kernel builds on a structured box with Dinv_i = A_ii^-1 (same per-block arithmetic)
template <int B>
__global__ void __launch_bounds__(128) ssor_factor_bcrs_kernel(int n, const int* rowptr, const int* colidx, const int* diag, const double* A,
                                                              double* out, int* flag)
{
    constexpr int BB = B * B;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) {
        const int j = colidx[k];
        double blk[BB];
#pragma unroll
        for (int q = 0; q < BB; ++q) blk[q] = A[(size_t)k * BB + q];
        if (j == i) {
            if (!invert_block<B>(blk)) atomicOr(flag, 1);
        } else if (j < i) {
            double D[BB], C[BB];
#pragma unroll
            for (int q = 0; q < BB; ++q) D[q] = A[(size_t)diag[j] * BB + q];
            invert_block<B>(D);
#pragma unroll
            for (int r = 0; r < B; ++r)
#pragma unroll
                for (int c = 0; c < B; ++c) {
                    double s = 0.0;
#pragma unroll
                    for (int kk = 0; kk < B; ++kk) s += blk[r * B + kk] * D[kk * B + c];
                    C[r * B + c] = s;
                }
#pragma unroll
            for (int q = 0; q < BB; ++q) blk[q] = C[q];
        }
#pragma unroll
        for (int q = 0; q < BB; ++q) out[(size_t)k * BB + q] = blk[q];
    }
}

// generic pattern: level-scheduled factorisation of a copy of J (ctx->d_ilu); sets ctx->d_flag on a singular block
int ilu0_factor_bcrs(dmx_ctx* ctx)
{
    const size_t bytes = (size_t)ctx->nnzb * ctx->b * ctx->b * sizeof(double);
    if (!ctx->d_ilu) DMX_CUDA(cudaMalloc((void**)&ctx->d_ilu, bytes));
    if (ctx->ssor_factorised) {
        DMX_CUDA(cudaMemsetAsync(ctx->d_flag, 0, sizeof(int), ctx->stream));
        const int grid = (ctx->n + 127) / 128;
        if (ctx->b == 2) ssor_factor_bcrs_kernel<2><<<grid, 128, 0, ctx->stream>>>(ctx->n, ctx->d_rowptr, ctx->d_colidx, ctx->d_diag, ctx->d_J, ctx->d_ilu, ctx->d_flag);
        else ssor_factor_bcrs_kernel<1><<<grid, 128, 0, ctx->stream>>>(ctx->n, ctx->d_rowptr, ctx->d_colidx, ctx->d_diag, ctx->d_J, ctx->d_ilu, ctx->d_flag);
        DMX_CHECK_LAUNCH();
        return 0;
    }
    DMX_CUDA(cudaMemcpyAsync(ctx->d_ilu, ctx->d_J, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    DMX_CUDA(cudaMemsetAsync(ctx->d_flag, 0, sizeof(int), ctx->stream));
    if (ctx->l_ptr.empty()) { if (int rc = build_level_schedule(ctx)) return rc; }
    const int nlev = (int)ctx->l_ptr.size() - 1;
    if (ctx->b == 2)
        return coop_launch(ctx, ilu0_factor_kernel<2>, 128, nlev, (const int*)ctx->d_lptr, (const int*)ctx->d_lrows, (const int*)ctx->d_rowptr,
                           (const int*)ctx->d_colidx, (const int*)ctx->d_diag, ctx->d_ilu, ctx->d_flag, ctx->d_barrier);
    return coop_launch(ctx, ilu0_factor_kernel<1>, 128, nlev, (const int*)ctx->d_lptr, (const int*)ctx->d_lrows, (const int*)ctx->d_rowptr,
                       (const int*)ctx->d_colidx, (const int*)ctx->d_diag, ctx->d_ilu, ctx->d_flag, ctx->d_barrier);
}

int ilu0_factor(dmx_ctx* ctx)
{
    ProfScope ps(ctx, DMX_K_ILU_FACTOR);
    ctx->ilu_valid = false;
    ctx->ilu_bcrs_valid = false;
    if (ctx->skew) {
        // structured box: diagonal recurrence + sweep streams straight from J (ilu_structured.cu), no factorised BCRS copy
        if (int rc = sk_factor(ctx)) return rc;
    } else {
        if (int rc = ilu0_factor_bcrs(ctx)) return rc;
    }
    // agreed over all ranks (the reference: the solver construction throws on one rank, newtonsolver.hh:510-523 comm.min)
    int bad = 0;
    if (int rc = agree_flag(ctx, &bad)) return rc;
    if (bad) { ctx->err = "ILU0: singular diagonal block"; return DMX_STATUS_BREAKDOWN; }
    ctx->ilu_valid = true;
    ctx->ilu_bcrs_valid = !ctx->skew;
    return 0;
}

int ilu0_apply(dmx_ctx* ctx, const double* d, double* v)
{
    ProfScope ps(ctx, DMX_K_ILU_APPLY);
    if (ctx->skew) return sk_apply(ctx, d, v);
    if (ctx->l_ptr.empty()) { if (int rc0 = build_level_schedule(ctx)) return rc0; }
    const int nl = (int)ctx->l_ptr.size() - 1, nu = (int)ctx->u_ptr.size() - 1;
    int rc;
    if (ctx->b == 2) {
        rc = coop_launch(ctx, ilu0_lower_kernel<2>, 128, nl, (const int*)ctx->d_lptr, (const int*)ctx->d_lrows, (const int*)ctx->d_rowptr,
                         (const int*)ctx->d_colidx, (const int*)ctx->d_diag, (const double*)ctx->d_ilu, d, v, ctx->d_barrier);
        if (rc) return rc;
        rc = coop_launch(ctx, ilu0_upper_kernel<2>, 128, nu, (const int*)ctx->d_uptr, (const int*)ctx->d_urows, (const int*)ctx->d_rowptr,
                         (const int*)ctx->d_colidx, (const int*)ctx->d_diag, (const double*)ctx->d_ilu, v, ctx->d_barrier);
    } else {
        rc = coop_launch(ctx, ilu0_lower_kernel<1>, 128, nl, (const int*)ctx->d_lptr, (const int*)ctx->d_lrows, (const int*)ctx->d_rowptr,
                         (const int*)ctx->d_colidx, (const int*)ctx->d_diag, (const double*)ctx->d_ilu, d, v, ctx->d_barrier);
        if (rc) return rc;
        rc = coop_launch(ctx, ilu0_upper_kernel<1>, 128, nu, (const int*)ctx->d_uptr, (const int*)ctx->d_urows, (const int*)ctx->d_rowptr,
                         (const int*)ctx->d_colidx, (const int*)ctx->d_diag, (const double*)ctx->d_ilu, v, ctx->d_barrier);
    }
    return rc;
}

// v = SeqSSOR(J)(d), starting from v = 0 (the Krylov solvers clear the correction before Preconditioner::apply)
int ssor_apply(dmx_ctx* ctx, const double* d, double* v)
{
    ProfScope ps(ctx, DMX_K_ILU_APPLY);
    if (ctx->l_ptr.empty()) { if (int rc0 = build_level_schedule(ctx)) return rc0; }
    const size_t len = (size_t)ctx->n * ctx->b;
    DMX_CUDA(cudaMemsetAsync(v, 0, len * sizeof(double), ctx->stream));
    const int nl = (int)ctx->l_ptr.size() - 1, nu = (int)ctx->u_ptr.size() - 1;
    int rc;
    if (ctx->b == 2) {
        rc = coop_launch(ctx, gs_sweep_kernel<2>, 128, nl, (const int*)ctx->d_lptr, (const int*)ctx->d_lrows, (const int*)ctx->d_rowptr,
                         (const int*)ctx->d_colidx, (const int*)ctx->d_diag, (const double*)ctx->d_J, d, v, ctx->d_barrier);
        if (rc) return rc;
        rc = coop_launch(ctx, gs_sweep_kernel<2>, 128, nu, (const int*)ctx->d_uptr, (const int*)ctx->d_urows, (const int*)ctx->d_rowptr,
                         (const int*)ctx->d_colidx, (const int*)ctx->d_diag, (const double*)ctx->d_J, d, v, ctx->d_barrier);
    } else {
        rc = coop_launch(ctx, gs_sweep_kernel<1>, 128, nl, (const int*)ctx->d_lptr, (const int*)ctx->d_lrows, (const int*)ctx->d_rowptr,
                         (const int*)ctx->d_colidx, (const int*)ctx->d_diag, (const double*)ctx->d_J, d, v, ctx->d_barrier);
        if (rc) return rc;
        rc = coop_launch(ctx, gs_sweep_kernel<1>, 128, nu, (const int*)ctx->d_uptr, (const int*)ctx->d_urows, (const int*)ctx->d_rowptr,
                         (const int*)ctx->d_colidx, (const int*)ctx->d_diag, (const double*)ctx->d_J, d, v, ctx->d_barrier);
    }
    return rc;
}

// ---------------------------------------------------------------------------------------------
// block-Jacobi alternative: v = D^-1 d
// ---------------------------------------------------------------------------------------------
template <int B>
__global__ void __launch_bounds__(256) jacobi_setup_kernel(int n, const int* diag, const double* A, double* dinv, int* flag)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double D[B * B];
#pragma unroll
    for (int q = 0; q < B * B; ++q) D[q] = A[(size_t)diag[i] * B * B + q];
    if (!invert_block<B>(D)) atomicOr(flag, 1);
#pragma unroll
    for (int q = 0; q < B * B; ++q) dinv[(size_t)i * B * B + q] = D[q];
}
template <int B>
__global__ void __launch_bounds__(256) jacobi_apply_kernel(int n, const double* dinv, const double* d, double* v)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
#pragma unroll
    for (int r = 0; r < B; ++r) {
        double s = 0.0;
#pragma unroll
        for (int c = 0; c < B; ++c) s += dinv[(size_t)i * B * B + r * B + c] * d[(size_t)i * B + c];
        v[(size_t)i * B + r] = s;
    }
}
int block_jacobi_setup(dmx_ctx* ctx)
{
    if (!ctx->d_dinv) DMX_CUDA(cudaMalloc((void**)&ctx->d_dinv, (size_t)ctx->n * ctx->b * ctx->b * sizeof(double)));
    DMX_CUDA(cudaMemsetAsync(ctx->d_flag, 0, sizeof(int), ctx->stream));
    const int grid = (ctx->n + 255) / 256;
    if (ctx->b == 2) jacobi_setup_kernel<2><<<grid, 256, 0, ctx->stream>>>(ctx->n, ctx->d_diag, ctx->d_J, ctx->d_dinv, ctx->d_flag);
    else jacobi_setup_kernel<1><<<grid, 256, 0, ctx->stream>>>(ctx->n, ctx->d_diag, ctx->d_J, ctx->d_dinv, ctx->d_flag);
    DMX_CHECK_LAUNCH();
    int bad = 0;
    if (int rc = agree_flag(ctx, &bad)) return rc;
    if (bad) { ctx->err = "block-Jacobi: singular diagonal block"; return DMX_STATUS_BREAKDOWN; }
    return 0;
}
int block_jacobi_apply(dmx_ctx* ctx, const double* d, double* v)
{
    ProfScope ps(ctx, DMX_K_JACOBI);
    const int grid = (ctx->n + 255) / 256;
    if (ctx->b == 2) jacobi_apply_kernel<2><<<grid, 256, 0, ctx->stream>>>(ctx->n, ctx->d_dinv, d, v);
    else jacobi_apply_kernel<1><<<grid, 256, 0, ctx->stream>>>(ctx->n, ctx->d_dinv, d, v);
    DMX_CHECK_LAUNCH();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Dumux::ParMTJac / ParMTSOR / ParMTSSOR (dumux/linear/preconditioners.hh:330-400, 442-620), DuMux's own multi-threaded
// smoothers ("par_mt_jac", "par_mt_sor", "par_mt_ssor").  Rows are coloured like computeColorsForMatrixSweep_ (:408-440: in
// index order, smallest colour no matrix neighbour has -- the checkerboard on the 7-point pattern); a sweep visits the
// colours in order (backward: descending), all rows of a colour in parallel: rhs = d_i - sum_j A_ij x_j over ALL columns
// ascending (diagonal included), v = A_ii^-1 rhs (FieldMatrix::solve), x_i += w v.  ParMTJac uses the previous iterate for
// every row.  One thread per block row; no wavefront, so these are plain HBM-streaming kernels.
// ---------------------------------------------------------------------------------------------
template <int B>
__global__ void __launch_bounds__(256) parmt_row_kernel(int nrows, const int* __restrict__ rows, const int* __restrict__ rowptr,
                                                        const int* __restrict__ colidx, const int* __restrict__ diag,
                                                        const double* __restrict__ A, const double* __restrict__ d, const double* xin,
                                                        double* xout, double w)
{
    constexpr int BB = B * B;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nrows) return;
    const int i = rows ? rows[q] : q;
    double rhs[B];
#pragma unroll
    for (int e = 0; e < B; ++e) rhs[e] = d[(size_t)i * B + e];
    const int kend = rowptr[i + 1];
    for (int k = rowptr[i]; k < kend; ++k) {
        const int c = colidx[k];
        double xv[B];
#pragma unroll
        for (int e = 0; e < B; ++e) xv[e] = xin[(size_t)c * B + e];
#pragma unroll
        for (int r = 0; r < B; ++r)
#pragma unroll
            for (int cc = 0; cc < B; ++cc) rhs[r] -= A[(size_t)k * BB + r * B + cc] * xv[cc];
    }
    const double* D = A + (size_t)diag[i] * BB;
    double v[B];
    if (B == 1) v[0] = rhs[0] / D[0];
    else {
        double detinv = D[0] * D[3] - D[1] * D[2];
        detinv = 1 / detinv;
        v[0] = detinv * (D[3] * rhs[0] - D[1] * rhs[B - 1]);
        v[B - 1] = detinv * (D[0] * rhs[B - 1] - D[2] * rhs[0]);
    }
#pragma unroll
    for (int e = 0; e < B; ++e) xout[(size_t)i * B + e] = xin[(size_t)i * B + e] + w * v[e];
}

static int parmt_build_colors(dmx_ctx* ctx)
{
    const int n = ctx->n;
    const std::vector<int>& rp = ctx->h_rowptr;
    const std::vector<int>& ci = ctx->h_colidx;
    std::vector<int> colors(n, -1), nb;
    std::vector<char> used;
    int ncol = 0;
    for (int i = 0; i < n; ++i) {
        nb.clear();
        for (int k = rp[i]; k < rp[i + 1]; ++k) nb.push_back(colors[ci[k]]);
        const int m = (int)nb.size();
        used.assign(m, 0);
        for (int q = 0; q < m; ++q)
            if (nb[q] >= 0 && nb[q] < m) used[nb[q]] = 1;
        int c = m;
        for (int q = 0; q < m; ++q)
            if (!used[q]) { c = q; break; }
        colors[i] = c;
        ncol = std::max(ncol, c + 1);
    }
    ctx->color_ptr.assign(ncol + 1, 0);
    for (int i = 0; i < n; ++i) ctx->color_ptr[colors[i] + 1]++;
    for (int c = 0; c < ncol; ++c) ctx->color_ptr[c + 1] += ctx->color_ptr[c];
    std::vector<int> rows(n), cur(ctx->color_ptr.begin(), ctx->color_ptr.end() - 1);
    for (int i = 0; i < n; ++i) rows[cur[colors[i]]++] = i;
    if (ctx->d_color_rows) cudaFree(ctx->d_color_rows);
    ctx->d_color_rows = nullptr;
    DMX_CUDA(cudaMalloc((void**)&ctx->d_color_rows, (size_t)n * sizeof(int)));
    DMX_CUDA(cudaMemcpy(ctx->d_color_rows, rows.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice));
    return 0;
}

// v = ParMT{Jac,SOR,SSOR}(J)(d) from v = 0
int parmt_apply(dmx_ctx* ctx, int precond, const double* d, double* v)
{
    return parmt_apply_prm(ctx, precond, d, v, ctx->precond_iterations, ctx->precond_relaxation);
}
// ... with the iteration count and relaxation factor handed in (as an AMG smoother: SmootherArgs of the hierarchy)
int parmt_apply_prm(dmx_ctx* ctx, int precond, const double* d, double* v, int iterations, double relaxation)
{
    ProfScope ps(ctx, DMX_K_JACOBI);
    const size_t len = (size_t)ctx->n * ctx->b;
    DMX_CUDA(cudaMemsetAsync(v, 0, len * sizeof(double), ctx->stream));
    const double w = relaxation;
    auto launch = [&](int nrows, const int* rows, const double* xin, double* xout) -> int {
        if (nrows <= 0) return 0;
        const int grid = (nrows + 255) / 256;
        if (ctx->b == 2)
            parmt_row_kernel<2><<<grid, 256, 0, ctx->stream>>>(nrows, rows, ctx->d_rowptr, ctx->d_colidx, ctx->d_diag, ctx->d_J, d, xin, xout, w);
        else
            parmt_row_kernel<1><<<grid, 256, 0, ctx->stream>>>(nrows, rows, ctx->d_rowptr, ctx->d_colidx, ctx->d_diag, ctx->d_J, d, xin, xout, w);
        DMX_CHECK_LAUNCH();
        return 0;
    };
    if (precond == DMX_PRECOND_PARMT_JAC) {
        if (!ctx->d_xold) DMX_CUDA(cudaMalloc((void**)&ctx->d_xold, len * sizeof(double)));
        for (int it = 0; it < iterations; ++it) {
            DMX_CUDA(cudaMemcpyAsync(ctx->d_xold, v, len * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
            if (int rc = launch(ctx->n, nullptr, ctx->d_xold, v)) return rc;
        }
        return 0;
    }
    const int ncol = (int)ctx->color_ptr.size() - 1;
    auto sweep = [&](bool forward) -> int {
        for (int cc = 0; cc < ncol; ++cc) {
            const int c = forward ? cc : ncol - 1 - cc;
            if (int rc = launch(ctx->color_ptr[c + 1] - ctx->color_ptr[c], ctx->d_color_rows + ctx->color_ptr[c], v, v)) return rc;
        }
        return 0;
    };
    for (int it = 0; it < iterations; ++it) {
        if (int rc = sweep(true)) return rc;
        if (precond == DMX_PRECOND_PARMT_SSOR)
            if (int rc = sweep(false)) return rc;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// BiCGSTAB: Dune::BiCGSTABSolver::apply restated.  x = DELTA (initial guess as given),
// b = RESIDUAL (not modified).  In a distributed ctx: operator = local SpMV + project, preconditioner = local
// ILU0 + copyOwnerToAll, scalar product = owner-masked dot + all-reduce (OverlappingSchwarz*, BlockPreconditioner).
// ---------------------------------------------------------------------------------------------
// fresh preconditioner per solve (istlsolvers.hh:457-463); SeqSSOR has no set-up
int precond_setup(dmx_ctx* ctx, int precond)
{
    if (precond == DMX_PRECOND_ILU0) { ctx->ssor_factorised = false; return ilu0_factor(ctx); }
    if (precond == DMX_PRECOND_AMG) return amg_setup(ctx);
    if (precond == DMX_PRECOND_SSOR) return 0;
    if (precond == DMX_PRECOND_BLOCKJACOBI) return block_jacobi_setup(ctx);
    if (precond == DMX_PRECOND_PARMT_JAC) return 0;
    if (precond == DMX_PRECOND_PARMT_SOR || precond == DMX_PRECOND_PARMT_SSOR) {
        if (ctx->color_ptr.empty()) return parmt_build_colors(ctx);
        return 0;
    }
    return fail(ctx, DMX_ERR_USAGE, "unknown preconditioner");
}
// the sequential preconditioner on this rank's (interior + overlap) matrix
int precond_apply_local(dmx_ctx* ctx, int precond, const double* d, double* v)
{
    switch (precond) {
        case DMX_PRECOND_ILU0: return ilu0_apply(ctx, d, v);
        case DMX_PRECOND_AMG: return amg_apply(ctx, d, v);
        case DMX_PRECOND_SSOR: return ssor_apply(ctx, d, v);
        case DMX_PRECOND_BLOCKJACOBI: return block_jacobi_apply(ctx, d, v);
        case DMX_PRECOND_PARMT_JAC:
        case DMX_PRECOND_PARMT_SOR:
        case DMX_PRECOND_PARMT_SSOR: return parmt_apply(ctx, precond, d, v);
    }
    return fail(ctx, DMX_ERR_USAGE, "unknown preconditioner");
}
// BlockPreconditioner::apply: the sequential preconditioner, then copyOwnerToAll
static int precond_apply(dmx_ctx* ctx, int precond, const double* d, double* v)
{
    if (int rc = precond_apply_local(ctx, precond, d, v)) return rc;
    // the block-decomposed AMG cycle already returns a correction that is consistent on the overlap (amg.cu)
    if (ctx->nranks > 1 && precond != DMX_PRECOND_AMG) return halo_exchange(ctx, v);
    return 0;
}

int bicgstab(dmx_ctx* ctx, double reduction, int maxit, int precond, int* iterations, double* achieved)
{
    const size_t len = (size_t)ctx->n * ctx->b;
    double* x = ctx->d_vec[DMX_VEC_DELTA];
    const double* rhs = ctx->d_vec[DMX_VEC_RESIDUAL];
    // r lives in WORK0 so that RESIDUAL stays intact for the caller (Newton's residual norm, tests)
    double* r = ctx->d_vec[DMX_VEC_WORK0];
    double *rt = ctx->d_rt, *p = ctx->d_p, *v = ctx->d_v, *t = ctx->d_t, *y = ctx->d_y;
    int rc;
    *iterations = 0;
    *achieved = 1.0;

    // fresh preconditioner per call (istlsolvers.hh:457-463)
    if ((rc = precond_setup(ctx, precond))) return rc;

    if (ctx->nranks > 1 && (rc = halo_exchange(ctx, x))) return rc;       // BlockPreconditioner::pre: copyOwnerToAll(x)
    if ((rc = launch_spmv(ctx, x, t))) return rc;
    {
        ProfScope ps(ctx, DMX_K_BLAS1);
        residual_init_kernel<<<RED_BLOCKS, RED_THREADS, 0, ctx->stream>>>(len, ctx->b, rhs, t, r, rt, ctx->d_owner, ctx->d_partials);
        DMX_CHECK_LAUNCH();
    }
    double s[3];
    if ((rc = reduce_to_host(ctx, 1, false, s))) return rc;
    const double norm0 = std::sqrt(s[0]);
    double norm = norm0;
    if (!(norm0 == norm0) || std::isinf(norm0)) return DMX_STATUS_NONFINITE;
    auto converged = [&](double nrm) { return nrm < reduction * norm0 || nrm < 1e-30; };
    if (converged(norm0)) { *achieved = norm0 > 0 ? 1.0 : 0.0; return 0; }
    DMX_CUDA(cudaMemsetAsync(p, 0, len * sizeof(double), ctx->stream));
    DMX_CUDA(cudaMemsetAsync(v, 0, len * sizeof(double), ctx->stream));

    double rho = 1, alpha = 1, omega = 1, rho_new = 0, h, beta;
    const double EPSILON = 1e-80;
    double it;
    int status = DMX_STATUS_NOT_CONVERGED;
    double* z = ctx->d_z;       // M^-1 r of the second half step (y keeps M^-1 p until x is updated)
    bool have_rho = false;      // <rt, r> already produced by the previous iteration's fused update
    auto flush_x = [&]() -> int {      // x += alpha*y that the first half step deferred
        ProfScope ps(ctx, DMX_K_BLAS1);
        axpy_kernel<<<RED_BLOCKS, 256, 0, ctx->stream>>>(len, alpha, y, x);
        DMX_CHECK_LAUNCH();
        return 0;
    };
    double sc[8];
    for (it = 0.5; it < maxit; it += .5) {
        if (!have_rho) {
            // <rt, r> of the first iteration: stand-alone dot, kept on the device as rho and read back
            {
                ProfScope ps(ctx, DMX_K_BLAS1);
                dot_kernel<1><<<RED_BLOCKS, RED_THREADS, 0, ctx->stream>>>(len, ctx->b, rt, r, nullptr, nullptr, nullptr, nullptr, ctx->d_owner,
                                                                            ctx->d_partials);
                DMX_CHECK_LAUNCH();
            }
            if ((rc = reduce_on_device(ctx, 1, 3))) return rc;
            if ((rc = read_scalars(ctx, sc))) return rc;
            rho_new = sc[SC_RHO];
        }
        if (std::fabs(rho) <= EPSILON || std::fabs(omega) <= EPSILON) { status = DMX_STATUS_BREAKDOWN; break; }
        if (it < 1)
            DMX_CUDA(cudaMemcpyAsync(p, r, len * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
        else {
            beta = (rho_new / rho) * (alpha / omega);
            ProfScope ps(ctx, DMX_K_BLAS1);
            p_update_kernel<<<RED_BLOCKS, 256, 0, ctx->stream>>>(len, beta, omega, r, v, p);
            DMX_CHECK_LAUNCH();
        }
        if ((rc = precond_apply(ctx, precond, p, y))) return rc;
        if ((rc = launch_spmv(ctx, y, v))) return rc;
        {
            // h = <rt, v>, alpha = rho / h on the device; r -= alpha v and ||r||^2 queued behind it; ONE host read for both
            ProfScope ps(ctx, DMX_K_BLAS1);
            dot_kernel<1><<<RED_BLOCKS, RED_THREADS, 0, ctx->stream>>>(len, ctx->b, rt, v, nullptr, nullptr, nullptr, nullptr, ctx->d_owner,
                                                                        ctx->d_partials);
            DMX_CHECK_LAUNCH();
            if ((rc = reduce_on_device(ctx, 1, 0))) return rc;
            axpy_r_norm_kernel<<<RED_BLOCKS, RED_THREADS, 0, ctx->stream>>>(len, ctx->b, ctx->d_scalars, v, r, ctx->d_owner, ctx->d_partials);
            DMX_CHECK_LAUNCH();
            if ((rc = reduce_on_device(ctx, 1, -1))) return rc;
        }
        if ((rc = read_scalars(ctx, sc))) return rc;
        h = sc[SC_H];
        if (std::fabs(h) < EPSILON) { status = DMX_STATUS_BREAKDOWN; break; }
        alpha = sc[SC_ALPHA];
        norm = std::sqrt(sc[0]);
        if (!(norm == norm) || std::isinf(norm)) { if ((rc = flush_x())) return rc; status = DMX_STATUS_NONFINITE; break; }
        if (converged(norm)) { if ((rc = flush_x())) return rc; status = 0; break; }
        it += .5;
        if ((rc = precond_apply(ctx, precond, r, z))) return rc;
        if ((rc = launch_spmv(ctx, z, t))) return rc;
        {
            // omega = <t, r> / <t, t> on the device; x, r update with ||r||^2 and the next rho queued behind it
            ProfScope ps(ctx, DMX_K_BLAS1);
            dot_kernel<2><<<RED_BLOCKS, RED_THREADS, 0, ctx->stream>>>(len, ctx->b, t, r, t, t, nullptr, nullptr, ctx->d_owner, ctx->d_partials);
            DMX_CHECK_LAUNCH();
            if ((rc = reduce_on_device(ctx, 2, 1))) return rc;
            axpy3_norm_dot_kernel<<<RED_BLOCKS, RED_THREADS, 0, ctx->stream>>>(len, ctx->b, ctx->d_scalars, y, z, t, rt, x, r, ctx->d_owner,
                                                                                 ctx->d_partials);
            DMX_CHECK_LAUNCH();
            if ((rc = reduce_on_device(ctx, 2, 2))) return rc;
        }
        if ((rc = read_scalars(ctx, sc))) return rc;
        omega = sc[SC_OMEGA];
        rho = rho_new;
        rho_new = sc[SC_RHO];
        have_rho = true;
        norm = std::sqrt(sc[0]);
        if (!(norm == norm) || std::isinf(norm)) { status = DMX_STATUS_NONFINITE; break; }
        if (converged(norm)) { status = 0; break; }
    }
    *iterations = (int)std::ceil(std::min(it, (double)maxit));
    *achieved = norm0 > 0 ? norm / norm0 : 0.0;
    if (status == DMX_STATUS_BREAKDOWN) ctx->err = "BiCGSTAB breakdown (rho/omega/h ~ 0)";
    if (status == DMX_STATUS_NOT_CONVERGED) ctx->err = "BiCGSTAB: maximum iterations reached";
    return status;
}

// ---------------------------------------------------------------------------------------------
// Restarted GMRes: Dune::RestartedGMResSolver::apply restated (what ILURestartedGMResIstlSolver runs,
// dumux/linear/istlsolvers.hh:660-667; the reference's 2p test uses it, test/porousmediumflow/2p/incompressible/main.cc:134).
// LEFT-preconditioned GMRes(m): the monitored norm is ||M^-1 (b - A x)||, Arnoldi with modified Gram-Schmidt, Givens rotations
// on the host (an (m+1) x m Hessenberg matrix), update by back substitution, restart from the recomputed defect.
// x = DELTA, b = RESIDUAL (not modified).  The basis lives in ctx->d_gm: m+1 basis vectors, w, and the defect.
// Every Gram-Schmidt coefficient is needed on the host for the QR update, so each dot is one host read.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gm_sub_kernel(size_t len, int b, const double* rhs, const double* Ax, double* out,
                                                     const unsigned char* __restrict__ owner)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (size_t)gridDim.x * blockDim.x) {
        const bool own = !owner || owner[i / b];
        out[i] = own ? rhs[i] - Ax[i] : 0.0;
    }
}
// x += (-h[0]) * y with the coefficient taken from device memory (the Gram-Schmidt chain runs without host round trips)
__global__ void __launch_bounds__(256) gm_axpy_neg_dev_kernel(size_t len, const double* __restrict__ h, const double* y, double* x)
{
    const double mh = -h[0];
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (size_t)gridDim.x * blockDim.x) x[i] += mh * y[i];
}
__global__ void __launch_bounds__(256) gm_scale_kernel(size_t len, double f, const double* src, double* dst)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i] * f;
}
static void gm_generate_rotation(double dx, double dy, double& cs, double& sn)
{
    const double eps = 1e-15;
    const double ndx = std::fabs(dx), ndy = std::fabs(dy);
    const double nmax = std::max(ndx, ndy), nmin = std::min(ndx, ndy);
    const double temp = nmin / nmax;
    if (ndy < eps) { cs = 1.0; sn = 0.0; }
    else if (ndx < eps) { cs = 0.0; sn = 1.0; }
    else if (ndy > ndx) { cs = 1.0 / std::sqrt(1.0 + temp * temp) * temp; sn = 1.0 / std::sqrt(1.0 + temp * temp) * dx * dy / ndx / ndy; }
    else { cs = 1.0 / std::sqrt(1.0 + temp * temp); sn = 1.0 / std::sqrt(1.0 + temp * temp) * dy / dx; }
}
static void gm_apply_rotation(double& dx, double& dy, double cs, double sn)
{
    const double temp = cs * dx + sn * dy;
    dy = -sn * dx + cs * dy;
    dx = temp;
}

int gmres(dmx_ctx* ctx, double reduction, int maxit, int restart, int precond, int* iterations, double* achieved)
{
    const size_t len = (size_t)ctx->n * ctx->b;
    const int m = restart > 0 ? restart : 10;
    if (!ctx->d_gm || ctx->gm_vectors < m + 3) {
        if (ctx->d_gm) cudaFree(ctx->d_gm);
        ctx->d_gm = nullptr;
        DMX_CUDA(cudaMalloc((void**)&ctx->d_gm, ((size_t)(m + 3) * len + (size_t)(m + 2)) * sizeof(double)));
        ctx->gm_vectors = m + 3;
    }
    double* d_h = ctx->d_gm + (size_t)ctx->gm_vectors * len;      // Gram-Schmidt coefficients of one Arnoldi step + ||w||^2
    std::vector<double> h_h(m + 2);
    auto V = [&](int k) { return ctx->d_gm + (size_t)k * len; };
    double* w = V(m + 1);
    double* bdef = V(m + 2);
    double* x = ctx->d_vec[DMX_VEC_DELTA];
    const double* rhs = ctx->d_vec[DMX_VEC_RESIDUAL];
    double* t = ctx->d_t;
    int rc;
    *iterations = 0;
    *achieved = 1.0;
    if ((rc = precond_setup(ctx, precond))) return rc;

    auto axpy = [&](double alpha, const double* y, double* xx) -> int {
        ProfScope ps(ctx, DMX_K_BLAS1);
        axpy_kernel<<<RED_BLOCKS, 256, 0, ctx->stream>>>(len, alpha, y, xx);
        DMX_CHECK_LAUNCH();
        return 0;
    };
    auto defect = [&](double* nrm) -> int {          // bdef = b - A x ; v[0] = M^-1 bdef ; ||v[0]||
        int r;
        if ((r = launch_spmv(ctx, x, t))) return r;
        {
            ProfScope ps(ctx, DMX_K_BLAS1);
            gm_sub_kernel<<<RED_BLOCKS, 256, 0, ctx->stream>>>(len, ctx->b, rhs, t, bdef, ctx->d_owner);
            DMX_CHECK_LAUNCH();
        }
        if ((r = precond_apply(ctx, precond, bdef, V(0)))) return r;
        double s2;
        if ((r = dot(ctx, V(0), V(0), &s2))) return r;
        *nrm = std::sqrt(s2);
        return 0;
    };
    // slab-decomposed context: the same overlapping-Schwarz pieces as bicgstab() -- operator = local SpMV + project (launch_spmv),
    // preconditioner = local ILU0 + copyOwnerToAll (precond_apply), scalar product = owner-masked dot + all-reduce (dot_kernel,
    // final reduction + allreduce_sum below); the basis vectors stay consistent on the overlap because every one of them is a
    // linear combination of preconditioner outputs
    if (ctx->nranks > 1 && (rc = halo_exchange(ctx, x))) return rc;       // BlockPreconditioner::pre: copyOwnerToAll(x)
    double norm;
    if ((rc = defect(&norm))) return rc;
    const double norm0 = norm;
    if (!(norm0 == norm0) || std::isinf(norm0)) return DMX_STATUS_NONFINITE;
    auto conv = [&](double nrm) { return nrm < reduction * norm0 || nrm < 1e-30; };
    if (conv(norm0)) { *achieved = norm0 > 0 ? 1.0 : 0.0; return 0; }

    const double EPSILON = 1e-80;
    std::vector<double> s(m + 1), sn(m), cs(m), H((size_t)(m + 1) * (m + 1), 0.0);
    auto Hm = [&](int r, int c) -> double& { return H[(size_t)r * (m + 1) + c]; };
    int j = 1;
    bool converged = false;
    int status = DMX_STATUS_NOT_CONVERGED;
    while (j <= maxit && !converged) {
        int i = 0;
        {
            ProfScope ps(ctx, DMX_K_BLAS1);
            gm_scale_kernel<<<RED_BLOCKS, 256, 0, ctx->stream>>>(len, norm == 0.0 ? 0.0 : 1.0 / norm, V(0), V(0));
            DMX_CHECK_LAUNCH();
        }
        s[0] = norm;
        for (i = 1; i < m + 1; ++i) s[i] = 0.0;
        for (i = 0; i < m && j <= maxit && !converged; ++i, ++j) {
            if ((rc = launch_spmv(ctx, V(i), V(i + 1)))) return rc;
            if ((rc = precond_apply(ctx, precond, V(i + 1), w))) return rc;
            {
                // modified Gram-Schmidt, chained on the device: h_k = <v_k, w> is reduced into d_h[k] and consumed by the next
                // update from there; the host reads all coefficients of the step (and ||w||^2) once
                ProfScope ps(ctx, DMX_K_BLAS1);
                for (int k = 0; k < i + 1; ++k) {
                    dot_kernel<1><<<RED_BLOCKS, RED_THREADS, 0, ctx->stream>>>(len, ctx->b, V(k), w, nullptr, nullptr, nullptr, nullptr,
                                                                                ctx->d_owner, ctx->d_partials);
                    DMX_CHECK_LAUNCH();
                    final_reduce_kernel<false><<<1, RED_THREADS, 0, ctx->stream>>>(RED_BLOCKS, 1, ctx->d_partials, d_h + k);
                    DMX_CHECK_LAUNCH();
                    if (ctx->nranks > 1 && (rc = allreduce_sum(ctx, d_h + k, 1))) return rc;
                    gm_axpy_neg_dev_kernel<<<RED_BLOCKS, 256, 0, ctx->stream>>>(len, d_h + k, V(k), w);
                    DMX_CHECK_LAUNCH();
                }
                dot_kernel<1><<<RED_BLOCKS, RED_THREADS, 0, ctx->stream>>>(len, ctx->b, w, w, nullptr, nullptr, nullptr, nullptr, ctx->d_owner,
                                                                            ctx->d_partials);
                DMX_CHECK_LAUNCH();
                final_reduce_kernel<false><<<1, RED_THREADS, 0, ctx->stream>>>(RED_BLOCKS, 1, ctx->d_partials, d_h + i + 1);
                DMX_CHECK_LAUNCH();
                if (ctx->nranks > 1 && (rc = allreduce_sum(ctx, d_h + i + 1, 1))) return rc;
            }
            DMX_CUDA(cudaMemcpyAsync(h_h.data(), d_h, (size_t)(i + 2) * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
            DMX_CUDA(cudaStreamSynchronize(ctx->stream));
            if (ctx->prof_pending.size() > 512) prof_drain(ctx);
            for (int k = 0; k < i + 1; ++k) Hm(k, i) = h_h[k];
            Hm(i + 1, i) = std::sqrt(h_h[i + 1]);
            *iterations = j;
            if (!(Hm(i + 1, i) == Hm(i + 1, i)) || std::isinf(Hm(i + 1, i))) return DMX_STATUS_NONFINITE;
            if (std::fabs(Hm(i + 1, i)) < EPSILON) {
                *achieved = norm / norm0;
                ctx->err = "GMRes breakdown (|w| == 0)";
                return DMX_STATUS_BREAKDOWN;
            }
            {
                ProfScope ps(ctx, DMX_K_BLAS1);
                gm_scale_kernel<<<RED_BLOCKS, 256, 0, ctx->stream>>>(len, norm == 0.0 ? 0.0 : 1.0 / Hm(i + 1, i), w, V(i + 1));
                DMX_CHECK_LAUNCH();
            }
            for (int k = 0; k < i; ++k) gm_apply_rotation(Hm(k, i), Hm(k + 1, i), cs[k], sn[k]);
            gm_generate_rotation(Hm(i, i), Hm(i + 1, i), cs[i], sn[i]);
            gm_apply_rotation(Hm(i, i), Hm(i + 1, i), cs[i], sn[i]);
            gm_apply_rotation(s[i], s[i + 1], cs[i], sn[i]);
            norm = std::fabs(s[i + 1]);
            if (conv(norm)) { converged = true; status = 0; }
        }
        // update(): y from the triangular system, x += sum_a y_a v_a accumulated from the last basis vector to the first
        DMX_CUDA(cudaMemsetAsync(w, 0, len * sizeof(double), ctx->stream));
        {
            std::vector<double> y(s);
            for (int a = i - 1; a >= 0; --a) {
                double r = s[a];
                for (int c = a + 1; c < i; ++c) r -= Hm(a, c) * y[c];
                y[a] = (r == 0.0) ? 0.0 : r / Hm(a, a);
                if ((rc = axpy(y[a], V(a), w))) return rc;
            }
        }
        if ((rc = axpy(1.0, w, x))) return rc;
        if (!converged && j < maxit && (rc = defect(&norm))) return rc;
    }
    *achieved = norm0 > 0 ? norm / norm0 : 0.0;
    if (status == DMX_STATUS_NOT_CONVERGED) ctx->err = "GMRes: maximum iterations reached";
    return status;
}

// ---------------------------------------------------------------------------------------------
// Conjugate gradients: Dune::CGSolver::apply restated (SSORCGIstlSolver, dumux/linear/istlsolvers.hh:701-714 -- the linear solver
// of the reference's 1p incompressible test).  x = DELTA, b = RESIDUAL (not modified); the defect lives in WORK0.
// ---------------------------------------------------------------------------------------------
int cg(dmx_ctx* ctx, double reduction, int maxit, int precond, int* iterations, double* achieved)
{
    // distributed ctx: the overlapping-Schwarz pieces as in bicgstab() (operator projects, dots are owner-masked + all-reduced,
    // the preconditioner is followed by copyOwnerToAll)
    const size_t len = (size_t)ctx->n * ctx->b;
    double* x = ctx->d_vec[DMX_VEC_DELTA];
    const double* rhs = ctx->d_vec[DMX_VEC_RESIDUAL];
    double* b = ctx->d_vec[DMX_VEC_WORK0];
    double *p = ctx->d_p, *q = ctx->d_v;
    int rc;
    *iterations = 0;
    *achieved = 1.0;
    if ((rc = precond_setup(ctx, precond))) return rc;
    auto axpy = [&](double alpha, const double* y, double* xx) -> int {
        ProfScope ps(ctx, DMX_K_BLAS1);
        axpy_kernel<<<RED_BLOCKS, 256, 0, ctx->stream>>>(len, alpha, y, xx);
        DMX_CHECK_LAUNCH();
        return 0;
    };
    if (ctx->nranks > 1 && (rc = halo_exchange(ctx, x))) return rc;       // BlockPreconditioner::pre: copyOwnerToAll(x)
    if ((rc = launch_spmv(ctx, x, q))) return rc;
    {
        ProfScope ps(ctx, DMX_K_BLAS1);
        gm_sub_kernel<<<RED_BLOCKS, 256, 0, ctx->stream>>>(len, ctx->b, rhs, q, b, ctx->d_owner);
        DMX_CHECK_LAUNCH();
    }
    double s2;
    if ((rc = dot(ctx, b, b, &s2))) return rc;
    double def = std::sqrt(s2);
    const double def0 = def;
    if (!(def0 == def0) || std::isinf(def0)) return DMX_STATUS_NONFINITE;
    auto conv = [&](double nrm) { return nrm < reduction * def0 || nrm < 1e-30; };
    if (conv(def0)) { *achieved = def0 > 0 ? 1.0 : 0.0; return 0; }
    if ((rc = precond_apply(ctx, precond, b, p))) return rc;
    double rholast, rho, lambda, alpha, beta;
    if ((rc = dot(ctx, p, b, &rholast))) return rc;
    int status = DMX_STATUS_NOT_CONVERGED, i = 1;
    for (; i <= maxit; ++i) {
        if ((rc = launch_spmv(ctx, p, q))) return rc;
        if ((rc = dot(ctx, p, q, &alpha))) return rc;
        lambda = rholast / alpha;
        if ((rc = axpy(lambda, p, x))) return rc;
        if ((rc = axpy(-lambda, q, b))) return rc;
        if ((rc = dot(ctx, b, b, &s2))) return rc;
        def = std::sqrt(s2);
        *iterations = i;
        if (!(def == def) || std::isinf(def)) { status = DMX_STATUS_NONFINITE; break; }
        if (conv(def)) { status = 0; break; }
        if ((rc = precond_apply(ctx, precond, b, q))) return rc;
        if ((rc = dot(ctx, q, b, &rho))) return rc;
        beta = rho / rholast;
        {
            // p = p*beta + q (p *= beta; p += q)
            ProfScope ps(ctx, DMX_K_BLAS1);
            gm_scale_kernel<<<RED_BLOCKS, 256, 0, ctx->stream>>>(len, beta, p, p);
            DMX_CHECK_LAUNCH();
        }
        if ((rc = axpy(1.0, q, p))) return rc;
        rholast = rho;
    }
    if (i > maxit) *iterations = maxit;
    *achieved = def0 > 0 ? def / def0 : 0.0;
    if (status == DMX_STATUS_NOT_CONVERGED) ctx->err = "CG: maximum iterations reached";
    return status;
}

// the solver selected with dmx_set_linear_solver (what NewtonSolver::solveLinearSystem calls)
int linear_solve(dmx_ctx* ctx, double reduction, int maxit, int precond, int* iterations, double* achieved)
{
    if (ctx->linear_solver == DMX_SOLVER_CG) return cg(ctx, reduction, maxit, precond, iterations, achieved);
    if (ctx->linear_solver == DMX_SOLVER_RESTARTED_GMRES) return gmres(ctx, reduction, maxit, ctx->gmres_restart, precond, iterations, achieved);
    return bicgstab(ctx, reduction, maxit, precond, iterations, achieved);
}

} // namespace dmx
