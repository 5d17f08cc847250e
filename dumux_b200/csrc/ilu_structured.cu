// ilu_structured.cu -- block ILU(0) triangular sweeps for the 7-point CCTpfa pattern on a structured box.
//
// Replaces Dune::SeqILU::apply -> ILU::blockILUBacksolve (dune-istl, called through
// dumux/linear/istlsolvers.hh:535-568) for matrices whose pattern comes from dmx_grid_structured.  The generic
// level-scheduled kernels in linalg.cu pay one grid-wide barrier per hyperplane i+j+k (3N-2 levels per sweep, ~7 us
// each at 256^3); here the same hyperplane order is executed WITHOUT grid barriers:
//
//   * the (i,j) plane is cut into 16x16 tiles, one CTA per tile, one thread per grid line (i,j); the CTA marches along
//     k and thread (a,b) handles layer k = s - a - b at step s, so inside a tile the wavefront costs one named barrier
//     among the 256 compute threads per step and neighbour values travel through shared memory;
//   * tiles depend only on their -x / -y neighbours (lower sweep; +x / +y for the upper sweep), which must be 16 steps
//     ahead; the boundary values travel as self-validating tagged words (see ilu_sweep_kernel) instead of behind a grid
//     barrier; tiles are handed out by an atomic ticket in dependency order, so waiting never deadlocks;
//   * factors AND the right-hand side travel in one tile-skewed stream per sweep, stream[tile][step][factors | rhs][thread]:
//     what a CTA needs at step s is ONE contiguous chunk (28 KB lower, 36 KB upper for 2x2 blocks), fetched with a single
//     cp.async.bulk (TMA, 1-D) into a shared-memory ring by a dedicated PRODUCER thread (warp 8) that re-arms a slot as soon
//     as the 256 compute threads release it (full/empty mbarriers).  Measured on B200 (scripts/probes/stream_probe.cu): with
//     the copy issued by a compute thread after the step's barrier an SM gets ONE copy per ~850-1100 cycles (26-34 B/clk at
//     28 KB), with a producer thread the same ring never makes the consumer wait -- what a tile that runs alone (fill/drain
//     of the tile wavefront) needs.  HBM sees long sequential streams, no index arrays at all;
//   * a SYNC warp (warp 9) does everything that talks to other tiles (polling and staging the upstream boundary values),
//     so none of that latency is on the compute threads' critical path.
//
// Per-row arithmetic (operation order, no FMA contraction) is that of blockILUBacksolve: columns ascending
// (-z,-y,-x | +x,+y,+z), y -= A x per block (FieldMatrix::mmv), v = Dinv * rhs last (FieldMatrix::mv, sum from 0), so
// results are bit-identical to the generic kernels and to the CPU oracle.
#include <algorithm>
#include <cstdint>
#include <cstdlib>

#include "common.cuh"

// This file is compiled TWICE into the library (csrc/Makefile): with 16 x 16 tiles (namespace sk16) and with 8 x 8 tiles
// (-DSK_TILE=8, namespace sk8; 64 compute threads, four CTAs per SM).  The dmx::sk_* entry points at the end of the 16 x 16
// build pick the variant per context.  Measured on B200 (AMG hierarchy of 256^3, 2x2 blocks, ms per V-cycle and level with
// 16 x 16 / 8 x 8 tiles): 128^3 2.57 / 2.61, 64^3 1.14 / 1.11, 32^3 0.56 / 0.54, 8^3 0.24 / 0.21, 4^3 0.21 / 0.18; 1x1 blocks at
// 256^3: 0.88 / 1.20 ms per application.  Small boxes are bound by steps x step latency (mbarrier hand-offs, the step barrier, the
// L2 round trip of the tile-to-tile words), which the smaller tile shortens only marginally -- so 8 x 8 serves boxes up to 64.
#ifndef SK_TILE
#define SK_TILE 16       // edge of the square (i,j) tile of a CTA (power of two, <= 16)
#endif
#if SK_TILE == 16
#define SK_VNS sk16
#else
#define SK_VNS sk8
#endif

namespace dmx {
namespace SK_VNS {

constexpr int SK_TI = SK_TILE, SK_TJ = SK_TILE, SK_THREADS = SK_TI * SK_TJ;
#ifndef SK_CTAS_PER_SM
#define SK_CTAS_PER_SM ((SK_TILE == 16) ? 1 : 4)
#endif
constexpr int SK_CTAS = SK_CTAS_PER_SM;     // resident CTAs per SM the sweep kernel is built for
constexpr int SK_C = 8;                      // unroll factor of the step loop (and granularity of the developer timeline)
static_assert(SK_TI + SK_TJ <= 32 && (SK_TI & (SK_TI - 1)) == 0, "one sync-warp lane per halo value of a step");

struct SkewGrid {
    int nx, ny, nz, ntx, nty, ntiles, NS;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(phase)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// barrier among the 256 compute threads only (the producer warp does not take part)
__device__ __forceinline__ void compute_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(SK_THREADS) : "memory"); }
__device__ __forceinline__ unsigned long long ld_relaxed(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// number of factor components (doubles) per cell: lower L_z,L_y,L_x ; upper U_x,U_y,U_z,Dinv
template <int B, bool UPPER>
struct SkewLayout {
    static constexpr int NBLK = UPPER ? 4 : 3;
    static constexpr int NC = NBLK * B * B;
    static constexpr int STAGE_DOUBLES = NC * SK_THREADS;                    // factor part of one step
    static constexpr int VEC_DOUBLES = B * SK_THREADS;                       // one step of a skewed vector
    // the upper stream carries the lower sweep's result next to the factors; the lower sweep reads its right-hand side
    // straight from the natural-layout vector (prefetched into registers), so its stream is factors only
    static constexpr int STEP_DOUBLES = STAGE_DOUBLES + VEC_DOUBLES;         // stream stride per step: [factors | right-hand side]
    // depth of the shared-memory ring (steps in flight): 7 x 28 KB lower, 5 x 36 KB upper for 2x2 blocks
    static constexpr int S = (B == 2) ? (UPPER ? 5 : 7) : 16;
};

// position of component (blk,r,c) of thread t inside one step chunk: 2x2 blocks are stored as double2 rows
template <int B>
__host__ __device__ __forceinline__ int sk_pos(int blk, int r, int c, int t)
{
    if (B == 2) return ((blk * 2 + r) * SK_THREADS + t) * 2 + c;
    return blk * SK_THREADS + t;
}

// ------------------------------------------------------------------------------------------------------------
// Structured block ILU(0) factorisation (ILU::blockILU0Decomposition, dune-istl ilu.hh, restated in oracle/oracle.cpp).
// On the 7-point pattern no product L_ij U_jk with k != i falls inside the pattern (boxes with >= 3 cells per axis, see
// sk_supported), so the elimination only ever updates the DIAGONAL block:
//     L_ij = A_ij Dinv_j                       (rightmultiply, j in {-z,-y,-x} in ascending column order)
//     D_i  = A_ii - sum_j L_ij A_ji            (per j: T[r][c] -= L[r][k] A_ji[k][c], k ascending)
//     Dinv_i = D_i^-1                          (FieldMatrix::invert)
// and U_ij = A_ij.  The recurrence runs through Dinv only, so it is split in two kernels:
//   ilu_diag_kernel   the recurrence over the nx+ny+nz-2 hyperplanes x+y+z = l (cooperative launch, one grid barrier per
//                     hyperplane, rows addressed from the grid instead of through a level schedule; reads J, writes n blocks);
//   ilu_skew_kernel   embarrassingly parallel: forms L_ij = A_ij Dinv_j and writes L, U and Dinv straight into the two
//                     tile-skewed streams of the sweeps -- no factorised BCRS copy of the Jacobian is made at all.
// Same operation sequence per block as factor_row (linalg.cu) -> bit-identical factors.
// ------------------------------------------------------------------------------------------------------------
// C = L * D (FieldMatrix::rightmultiply: sums start from 0)
template <int B>
__device__ __forceinline__ void block_rmul(const double* L, const double* D, double* C)
{
#pragma unroll
    for (int r = 0; r < B; ++r)
#pragma unroll
        for (int c = 0; c < B; ++c) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < B; ++k) s += L[r * B + k] * D[k * B + c];
            C[r * B + c] = s;
        }
}
template <int B>
__device__ __forceinline__ void load_block(const double* p, double* out)
{
    if (B == 2) {
        const double2 lo = *reinterpret_cast<const double2*>(p), hi = *reinterpret_cast<const double2*>(p + 2);
        out[0] = lo.x; out[B - 1] = lo.y; out[B * B - 2] = hi.x; out[B * B - 1] = hi.y;
    } else out[0] = p[0];
}
template <int B>
__device__ __forceinline__ void load_block_cg(const double* p, double* out)
{
    if (B == 2) {
        const double2 lo = __ldcg(reinterpret_cast<const double2*>(p)), hi = __ldcg(reinterpret_cast<const double2*>(p + 2));
        out[0] = lo.x; out[B - 1] = lo.y; out[B * B - 2] = hi.x; out[B * B - 1] = hi.y;
    } else out[0] = __ldcg(p);
}

template <int B>
__global__ void __launch_bounds__(256) ilu_diag_kernel(SkewGrid g, const int* __restrict__ diag, const double* __restrict__ A,
                                                        double* Dinv, int* flag, unsigned int* barrier)
{
    constexpr int BB = B * B;
    unsigned int epoch = 0;
    const int nlev = g.nx + g.ny + g.nz - 2;
    const int nthreads = gridDim.x * blockDim.x, tid = blockIdx.x * blockDim.x + threadIdx.x;
    const size_t sy = (size_t)g.nx, sz = (size_t)g.nx * g.ny;
    for (int l = 0; l < nlev; ++l) {
        const int zmin = max(0, l - (g.nx - 1) - (g.ny - 1)), zmax = min(g.nz - 1, l);
        const int cand = (zmax - zmin + 1) * g.ny;
        for (int q = tid; q < cand; q += nthreads) {
            const int zz = q / g.ny;
            const int z = zmin + zz, y = q - zz * g.ny, x = l - y - z;
            if (x < 0 || x >= g.nx) continue;
            const size_t I = (size_t)x + sy * y + sz * z;
            const int kd = __ldg(diag + I);
            const bool ex[3] = {z > 0, y > 0, x > 0};
            const size_t J[3] = {I - sz, I - sy, I - 1};
            // position of A_ji in row j behind its diagonal: the upper neighbours of j with a smaller column than i
            const int up[3] = {1 + (x + 1 < g.nx ? 1 : 0) + (y + 1 < g.ny ? 1 : 0), 1 + (x + 1 < g.nx ? 1 : 0), 1};
            double D[BB], Lb[3][BB], Ub[3][BB], Dj[3][BB];
            int p = kd - (int)ex[0] - (int)ex[1] - (int)ex[2];
            load_block<B>(A + (size_t)kd * BB, D);
#pragma unroll
            for (int s = 0; s < 3; ++s)
                if (ex[s]) {
                    load_block<B>(A + (size_t)p * BB, Lb[s]);
                    load_block<B>(A + ((size_t)__ldg(diag + J[s]) + up[s]) * BB, Ub[s]);
                    load_block_cg<B>(Dinv + J[s] * BB, Dj[s]);
                    ++p;
                }
#pragma unroll
            for (int s = 0; s < 3; ++s)
                if (ex[s]) {
                    double C[BB];
                    block_rmul<B>(Lb[s], Dj[s], C);
#pragma unroll
                    for (int r = 0; r < B; ++r)
#pragma unroll
                        for (int c = 0; c < B; ++c) {
                            double t = D[r * B + c];
#pragma unroll
                            for (int k = 0; k < B; ++k) t -= C[r * B + k] * Ub[s][k * B + c];
                            D[r * B + c] = t;
                        }
                }
            if (!invert_block<B>(D)) atomicOr(flag, 1);
            if (B == 2) {
                __stcg(reinterpret_cast<double2*>(Dinv + I * BB), make_double2(D[0], D[B - 1]));
                __stcg(reinterpret_cast<double2*>(Dinv + I * BB + 2), make_double2(D[BB - 2], D[BB - 1]));
            } else __stcg(Dinv + I, D[0]);
        }
        if (l + 1 < nlev) grid_barrier(barrier, epoch);
    }
}

// Jacobian with all off-diagonal blocks exactly zero (explicit tracer step): every L_ij is 0 * Dinv_j = 0 and the diagonal is
// never updated, so the recurrence degenerates to Dinv_i = A_ii^-1 for all rows at once -- same bits, no hyperplane loop.
template <int B>
__global__ void __launch_bounds__(256) ilu_diag_only_kernel(int n, const int* __restrict__ diag, const double* __restrict__ A, double* Dinv,
                                                             int* flag)
{
    constexpr int BB = B * B;
    const int I = blockIdx.x * blockDim.x + threadIdx.x;
    if (I >= n) return;
    double D[BB];
    load_block<B>(A + (size_t)diag[I] * BB, D);
    if (!invert_block<B>(D)) atomicOr(flag, 1);
    for (int e = 0; e < BB; ++e) Dinv[(size_t)I * BB + e] = D[e];
}

// ------------------------------------------------------------------------------------------------------------
// (J, Dinv) -> tile-skewed L and U streams.  One thread per (tile, step, thread slot).
// Mirrored thread coordinates: slot t = a + 16*b; lower: il = a, jl = b, k = s - a - b;
//                              upper: il = 15-a, jl = 15-b, k = nz-1 - (s - a - b).
// ------------------------------------------------------------------------------------------------------------
template <int B>
__global__ void __launch_bounds__(SK_THREADS) ilu_skew_kernel(SkewGrid g, const int* __restrict__ diag, const double* __restrict__ A,
                                                               const double* __restrict__ Dinv, double* __restrict__ Lsk,
                                                               double* __restrict__ Usk)
{
    constexpr int BB = B * B;
    const int t = threadIdx.x;
    const int a = t & (SK_TI - 1), b = t / SK_TI;
    const int s = blockIdx.x % g.NS;
    const int tile = blockIdx.x / g.NS;
    const int ti = tile % g.ntx, tj = tile / g.ntx;
    const size_t sy = (size_t)g.nx, sz = (size_t)g.nx * g.ny;
    {
        // lower
        const int i = ti * SK_TI + a, j = tj * SK_TJ + b, k = s - a - b;
        double* dst = Lsk + ((size_t)tile * g.NS + s) * SkewLayout<B, false>::STEP_DOUBLES;
        const bool valid = i < g.nx && j < g.ny && k >= 0 && k < g.nz;
        double blk[3][BB];
#pragma unroll
        for (int q = 0; q < 3; ++q)
#pragma unroll
            for (int e = 0; e < BB; ++e) blk[q][e] = 0.0;
        if (valid) {
            const size_t I = (size_t)i + sy * j + sz * k;
            const bool ex[3] = {k > 0, j > 0, i > 0};
            const size_t J[3] = {I - sz, I - sy, I - 1};
            int p = diag[I] - (int)ex[0] - (int)ex[1] - (int)ex[2];
#pragma unroll
            for (int q = 0; q < 3; ++q)
                if (ex[q]) {
                    double Lb[BB], Dj[BB];
                    load_block<B>(A + (size_t)p * BB, Lb);
                    load_block<B>(Dinv + J[q] * BB, Dj);
                    block_rmul<B>(Lb, Dj, blk[q]);
                    ++p;
                }
        }
#pragma unroll
        for (int q = 0; q < 3; ++q)
#pragma unroll
            for (int r = 0; r < B; ++r)
#pragma unroll
                for (int c = 0; c < B; ++c) dst[sk_pos<B>(q, r, c, t)] = blk[q][r * B + c];
    }
    {
        // upper
        const int i = ti * SK_TI + (SK_TI - 1 - a), j = tj * SK_TJ + (SK_TJ - 1 - b), k = g.nz - 1 - (s - a - b);
        double* dst = Usk + ((size_t)tile * g.NS + s) * SkewLayout<B, true>::STEP_DOUBLES;
        const bool valid = i < g.nx && j < g.ny && k >= 0 && k < g.nz;
        double blk[4][BB];
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int e = 0; e < BB; ++e) blk[q][e] = 0.0;
        if (valid) {
            const size_t I = (size_t)i + sy * j + sz * k;
            int p = diag[I];
            load_block<B>(Dinv + I * BB, blk[3]);
            ++p;
            if (i + 1 < g.nx) { load_block<B>(A + (size_t)p * BB, blk[0]); ++p; }
            if (j + 1 < g.ny) { load_block<B>(A + (size_t)p * BB, blk[1]); ++p; }
            if (k + 1 < g.nz) { load_block<B>(A + (size_t)p * BB, blk[2]); ++p; }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int r = 0; r < B; ++r)
#pragma unroll
                for (int c = 0; c < B; ++c) dst[sk_pos<B>(q, r, c, t)] = blk[q][r * B + c];
    }
}

// factorised BCRS values (what the generic ilu0_factor_kernel leaves in ctx->d_ilu) from (J, Dinv): for dmx_ilu0_download only
template <int B>
__global__ void __launch_bounds__(256) ilu_export_kernel(int n, const int* __restrict__ rowptr, const int* __restrict__ colidx,
                                                          const double* __restrict__ A, const double* __restrict__ Dinv, double* out)
{
    constexpr int BB = B * B;
    const int I = blockIdx.x * blockDim.x + threadIdx.x;
    if (I >= n) return;
    for (int k = rowptr[I]; k < rowptr[I + 1]; ++k) {
        const int j = colidx[k];
        double blk[BB], r[BB];
        load_block<B>(A + (size_t)k * BB, blk);
        if (j < I) {
            double Dj[BB];
            load_block<B>(Dinv + (size_t)j * BB, Dj);
            block_rmul<B>(blk, Dj, r);
        } else if (j == I) load_block<B>(Dinv + (size_t)I * BB, r);
        else
            for (int e = 0; e < BB; ++e) r[e] = blk[e];
        for (int e = 0; e < BB; ++e) out[(size_t)k * BB + e] = r[e];
    }
}

// ------------------------------------------------------------------------------------------------------------
// vectors in the tile-skewed layout (LOWER indexing): [(tile*NS + s)][t][e], thread slot t = a + 16 b, cell (i0+a, j0+b,
// k = s-a-b).  vec_skew writes the right-hand side into the vector slots of the LOWER stream; the upper sweep walks the same
// ordering backwards: step s_up = NS-1-s, lane 255-t.
// ------------------------------------------------------------------------------------------------------------
template <int B>
__global__ void __launch_bounds__(SK_THREADS) vec_skew_kernel(SkewGrid g, const double* __restrict__ x, double* __restrict__ xsk)
{
    const int t = threadIdx.x;
    const int a = t & (SK_TI - 1), b = t / SK_TI;
    const int s = blockIdx.x % g.NS;
    const int tile = blockIdx.x / g.NS;
    const int i = (tile % g.ntx) * SK_TI + a, j = (tile / g.ntx) * SK_TJ + b, k = s - a - b;
    const bool valid = i < g.nx && j < g.ny && k >= 0 && k < g.nz;
    double val[B];
#pragma unroll
    for (int e = 0; e < B; ++e) val[e] = 0.0;
    if (valid) {
        const size_t I = (size_t)i + (size_t)g.nx * (j + (size_t)g.ny * k);
        if (B == 2) {
            const double2 w = *reinterpret_cast<const double2*>(x + I * 2);
            val[0] = w.x; val[B - 1] = w.y;
        } else val[0] = x[I];
    }
    using LY = SkewLayout<B, false>;
    double* dst = xsk + (size_t)blockIdx.x * LY::STEP_DOUBLES + LY::STAGE_DOUBLES + (size_t)t * B;
    if (B == 2) *reinterpret_cast<double2*>(dst) = make_double2(val[0], val[B - 1]);
    else dst[0] = val[0];
}
// ------------------------------------------------------------------------------------------------------------
// One triangular sweep.  LOWER: L^-1 rhs (unit lower); the right-hand side sits in the vector slots of the lower stream
// (vec_skew_kernel), the result goes into the vector slots of the UPPER stream (`out`).  UPPER: U^-1 rhs, result scattered
// into the natural-layout vector (`out`; 16-byte stores -- the two halves of a 32-byte sector are written one step apart and
// merge in L2, which replaces a separate un-skew pass).
//
// Tile-to-tile halo exchange is flag-in-data: the two edge lines of a tile publish every value as 64-bit words that carry a
// 32-bit tag next to half a double (strong 8-byte stores are single-copy atomic, so a word is valid as soon as its tag
// matches; the tag changes with every sweep); the sync warp of the downstream tile polls exactly the words it needs, two steps
// per batch.  No progress counters, no release fences, and the tile-to-tile lag is the geometric minimum of 16 steps.
// Measured and rejected on the way (B200, 256^3, ms per ILU application): progress words + st.release published by thread 0
// every 8 steps with the TMA ring issued by thread 0: 1.92; + producer thread: 1.81; + sync warp (polling, halo staging,
// publication off the compute threads): 1.68; 4-step chunks: slower (the release fence cannot keep up); lower sweep gathering
// its right-hand side from the natural layout: slower than the skew pass; flag-in-data with dependent polls: 2.79; with
// batched polls: 1.39.
// ------------------------------------------------------------------------------------------------------------
#ifndef SK_RING
#define SK_RING 8
#endif
#ifndef SK_BATCH
#define SK_BATCH 2       // measured at 256^3: 2 -> 1.39 ms per apply, 4 -> 1.42, 8 -> 1.57
#endif
// halo ring depth in steps / steps whose halo words the sync warp requests together.  1x1 blocks move four times less data per
// step, their sweeps are pure latency: a deeper ring with 4-step batches measured 0.924 -> 0.869 ms per application at 256^3
template <int B> struct SkHalo {
    static constexpr int R = (B == 1) ? 2 * SK_RING : SK_RING;
    static constexpr int NB = (B == 1) ? 2 * SK_BATCH : SK_BATCH;
};
template <int B, bool UPPER>
__global__ void __launch_bounds__(SK_THREADS + 64, SK_CTAS) ilu_sweep_kernel(SkewGrid g, const double* __restrict__ stream, double* out,
                                                                        unsigned long long* ll, unsigned int tag,
                                                                           const int* __restrict__ order, unsigned long long* ticket_ctr,
                                                                           unsigned long long ticket_base, long long* trace)
{
    using LY = SkewLayout<B, UPPER>;
    using LYU = SkewLayout<B, true>;
    constexpr int S = LY::S;
    constexpr int SK_R = SkHalo<B>::R, SK_NB = SkHalo<B>::NB;
    // LOWER: results go to the vector slots of the upper stream (step stride STEP_U, offset FAC_U)
    constexpr size_t OUT_STEP = (size_t)LYU::STEP_DOUBLES;
    constexpr size_t OUT_OFF = (size_t)LYU::STAGE_DOUBLES;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* stages = reinterpret_cast<double*>(smem_raw);                              // [S][factors | rhs]
    double* sv = stages + (size_t)S * LY::STEP_DOUBLES;                                // [2][TJ+1][TI+1][B]
    double* hring = sv + 2 * (SK_TJ + 1) * (SK_TI + 1) * B;                            // [SK_R][x halo TJ | y halo TI][B]
    constexpr int HSTEP_DOUBLES = (SK_TI + SK_TJ) * B;
    uint64_t* full = reinterpret_cast<uint64_t*>(hring + SK_R * HSTEP_DOUBLES);        // [S] data landed
    uint64_t* empty = full + S;                                                        // [S] slot released by the compute threads
    uint64_t* hready = empty + S;                                                      // [SK_R] halo of a step staged by the sync warp
    uint64_t* hfree = hready + SK_R;                                                   // [SK_R] halo slot released by the compute threads
    __shared__ int s_tile;

    const int t = threadIdx.x;
    const int a = t & (SK_TI - 1), b = t / SK_TI;          // mirrored coordinates for UPPER
    const int tl = UPPER ? SK_THREADS - 1 - t : t;      // lane in LOWER indexing (storage)
    if (t == 0) {
        const unsigned long long ticket = atomicAdd(ticket_ctr, 1ull) - ticket_base;
        s_tile = order[(int)ticket];
        for (int q = 0; q < S; ++q) { mbar_init(&full[q], 1); mbar_init(&empty[q], 1); }
        for (int q = 0; q < SK_R; ++q) { mbar_init(&hready[q], 1); mbar_init(&hfree[q], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int tile = s_tile;
    const int ti = tile % g.ntx, tj = tile / g.ntx;
    const int NS = g.NS;
    const double* stream_tile = stream + (size_t)tile * NS * LY::STEP_DOUBLES;
    // results of step s (this sweep's step index) go to "lower step" sl = UPPER ? NS-1-s : s; the LOWER sweep stores them where
    // the upper sweep will fetch them: upper step NS-1-sl
    auto out_step_index = [&](int sl) { return NS - 1 - sl; };
    double* out_tile = out + (size_t)tile * NS * OUT_STEP + OUT_OFF;
    constexpr int HSHIFT = SK_TI - 1;       // == SK_TJ - 1
    static_assert(SK_TI == SK_TJ, "square tiles");
    const int tix = UPPER ? ti + 1 : ti - 1, tjy = UPPER ? tj + 1 : tj - 1;
    const bool tilex = tix >= 0 && tix < g.ntx, tiley = tjy >= 0 && tjy < g.nty;

    // halo words of (tile, step, direction, edge cell): 2*B 64-bit words {tag : 32 | half of a double : 32}
    auto ll_words = [&](int tl_, int s_, int dir, int e) {
        return ll + ((((size_t)tl_ * NS + s_) * 2 + dir) * SK_TI + e) * (2 * B);
    };
    if (t >= SK_THREADS + 32) {
        // ---- sync warp: stages the upstream tiles' boundary values step by step, polling the self-validating words the
        //      upstream edge threads wrote (no progress counters, no fences: every 8-byte word carries its own tag) ----
        const int lane = t - (SK_THREADS + 32);
        const int dir = lane / SK_TI, hl = lane % SK_TI;        // lanes 0-15: x halo (row hl), 16-31: y halo (column hl)
        const int htile = dir == 0 ? tix + g.ntx * tj : ti + g.ntx * tjy;
        const int other = dir == 0 ? tj * SK_TJ + (UPPER ? SK_TJ - 1 - hl : hl) : ti * SK_TI + (UPPER ? SK_TI - 1 - hl : hl);
        const bool hlane = lane < SK_TI + SK_TJ;               // lanes beyond the two edges idle (tiles smaller than 16 x 16)
        const bool hvalid = hlane && (dir == 0 ? tilex : tiley) && other < (dir == 0 ? g.ny : g.nx);
        // Batches of SK_NB steps: all words of a batch are requested at once (one L2 round trip per batch instead of 2*B
        // dependent round trips per step); a word whose tag does not match yet is polled again when its step is due.
        for (int s0 = 0; s0 < NS; s0 += SK_NB) {
            unsigned long long wq[SK_NB][2 * B];
            bool okq[SK_NB];
#pragma unroll
            for (int k = 0; k < SK_NB; ++k) {
                const int s = s0 + k;
                const int kk = s - hl;
                okq[k] = s < NS && hvalid && kk >= 0 && kk < g.nz;
                if (okq[k]) {
                    const unsigned long long* w = ll_words(htile, s + HSHIFT, dir, hl);
#pragma unroll
                    for (int e = 0; e < 2 * B; ++e) wq[k][e] = ld_relaxed(w + e);
                }
            }
#pragma unroll
            for (int k = 0; k < SK_NB; ++k) {
                const int s = s0 + k;
                if (s >= NS) break;       // uniform
                const int q = s % SK_R;
                double val[B];
#pragma unroll
                for (int e = 0; e < B; ++e) val[e] = 0.0;
                if (okq[k]) {
                    const unsigned long long* w = ll_words(htile, s + HSHIFT, dir, hl);
#pragma unroll
                    for (int e = 0; e < 2 * B; ++e)
                        while ((unsigned int)(wq[k][e] >> 32) != tag) wq[k][e] = ld_relaxed(w + e);
#pragma unroll
                    for (int e = 0; e < B; ++e)
                        val[e] = __longlong_as_double((long long)((wq[k][2 * e + 1] << 32) | (wq[k][2 * e] & 0xffffffffull)));
                }
                if (s >= SK_R) {
                    if (lane == 0) mbar_wait(&hfree[q], (uint32_t)(((s / SK_R) - 1) & 1));
                }
                __syncwarp();
#pragma unroll
                for (int e = 0; e < B; ++e)
                    if (hlane) hring[q * HSTEP_DOUBLES + lane * B + e] = val[e];
                __syncwarp();
                if (lane == 0) mbar_arrive(&hready[q]);
            }
        }
        return;
    }
    if (t >= SK_THREADS) {
        // ---- producer: one bulk copy per step, as far ahead as the ring allows ----
        if (t == SK_THREADS) {
            constexpr uint32_t BYTES = LY::STEP_DOUBLES * sizeof(double);
            for (int s = 0; s < NS; ++s) {
                const int q = s % S;
                if (s >= S) mbar_wait(&empty[q], (uint32_t)(((s / S) - 1) & 1));
                mbar_expect_tx(&full[q], BYTES);
                bulk_g2s(stages + (size_t)q * LY::STEP_DOUBLES, stream_tile + (size_t)s * LY::STEP_DOUBLES, BYTES, &full[q]);
            }
        }
        return;
    }

    // ---- compute threads: actual cell line of this thread ----
    const int i = ti * SK_TI + (UPPER ? SK_TI - 1 - a : a);
    const int j = tj * SK_TJ + (UPPER ? SK_TJ - 1 - b : b);
    const bool line = i < g.nx && j < g.ny;
    const bool depx = UPPER ? (i + 1 < g.nx) : (i > 0);          // a -x (+x) neighbour cell exists
    const bool depy = UPPER ? (j + 1 < g.ny) : (j > 0);

    double vprev[B];
#pragma unroll
    for (int e = 0; e < B; ++e) vprev[e] = 0.0;
    // natural-layout index of this thread's cell at wavefront distance kk (layer kk for the lower sweep, nz-1-kk for the upper)
    auto cell_index = [&](int kk) { return (size_t)i + (size_t)g.nx * ((size_t)j + (size_t)g.ny * (size_t)(UPPER ? g.nz - 1 - kk : kk)); };
    long long* tr = nullptr;       // optional timeline (developer diagnostic): tiles ticketed 0 and ntiles/2
    if (trace && t == 0) {
        if (tile == order[0]) tr = trace;
        else if (tile == order[g.ntiles / 2]) tr = trace + 24 * 64;
    }
#define SK_STAMP(slot) do { if (tr && s0 / SK_C < 64) tr[(s0 / SK_C) * 24 + (slot)] = clock64(); } while (0)
    for (int s0 = 0; s0 < NS; s0 += SK_C) {
        // ---- chunk head: the sync warp has staged the upstream tiles' boundary values of this chunk ----
        SK_STAMP(0);
        SK_STAMP(1);
        SK_STAMP(2);

#pragma unroll
        for (int c = 0; c < SK_C; ++c) {
            const int s = s0 + c;
            if (s < NS) {       // uniform
                const int stage = s % S;
                const int hq = s % SK_R;
                // only the threads on the two upstream edges consume staged halo values; the per-step barrier below keeps the
                // others from running ahead of the ring
                if (a == 0 || b == 0) mbar_wait(&hready[hq], (uint32_t)((s / SK_R) & 1));
                const double* hx = hring + hq * HSTEP_DOUBLES;
                const double* hy = hx + SK_TJ * B;
                mbar_wait(&full[stage], (uint32_t)((s / S) & 1));
                SK_STAMP(3 + 2 * c);
                const double* f = stages + (size_t)stage * LY::STEP_DOUBLES;
                const int kk = s - a - b;
                const bool active = line && kk >= 0 && kk < g.nz;
                const int rb = (s + 1) & 1, wb = s & 1;       // buffer written at step s-1 / written now
                double* svw = sv + ((wb * (SK_TJ + 1) + (b + 1)) * (SK_TI + 1) + (a + 1)) * B;
                if (active) {
                    double r[B];
#pragma unroll
                    for (int e = 0; e < B; ++e) r[e] = f[LY::STAGE_DOUBLES + tl * B + e];
                    double xv[B], yv[B];
                    const double* xs = (a == 0) ? hx + b * B : sv + ((rb * (SK_TJ + 1) + (b + 1)) * (SK_TI + 1) + a) * B;
                    const double* ys = (b == 0) ? hy + a * B : sv + ((rb * (SK_TJ + 1) + b) * (SK_TI + 1) + (a + 1)) * B;
#pragma unroll
                    for (int e = 0; e < B; ++e) { xv[e] = xs[e]; yv[e] = ys[e]; }
                    double blk[LY::NBLK][B * B];
                    if (B == 2) {
#pragma unroll
                        for (int q = 0; q < LY::NBLK; ++q)
#pragma unroll
                            for (int rr = 0; rr < 2; ++rr) {
                                const double2 w = reinterpret_cast<const double2*>(f)[(q * 2 + rr) * SK_THREADS + t];
                                blk[q][rr * 2] = w.x;
                                blk[q][rr * 2 + 1] = w.y;
                            }
                    } else {
#pragma unroll
                        for (int q = 0; q < LY::NBLK; ++q) blk[q][0] = f[q * SK_THREADS + t];
                    }
                    const bool depz = kk > 0;
                    if (!UPPER) {
                        // columns ascending: -z, -y, -x   (rhs -= A_ij v_j, FieldMatrix::mmv order)
                        if (depz) {
#pragma unroll
                            for (int rr = 0; rr < B; ++rr)
#pragma unroll
                                for (int cc = 0; cc < B; ++cc) r[rr] -= blk[0][rr * B + cc] * vprev[cc];
                        }
                        if (depy) {
#pragma unroll
                            for (int rr = 0; rr < B; ++rr)
#pragma unroll
                                for (int cc = 0; cc < B; ++cc) r[rr] -= blk[1][rr * B + cc] * yv[cc];
                        }
                        if (depx) {
#pragma unroll
                            for (int rr = 0; rr < B; ++rr)
#pragma unroll
                                for (int cc = 0; cc < B; ++cc) r[rr] -= blk[2][rr * B + cc] * xv[cc];
                        }
                    } else {
                        // columns ascending: +x, +y, +z, then v = Dinv * rhs (sum from 0)
                        if (depx) {
#pragma unroll
                            for (int rr = 0; rr < B; ++rr)
#pragma unroll
                                for (int cc = 0; cc < B; ++cc) r[rr] -= blk[0][rr * B + cc] * xv[cc];
                        }
                        if (depy) {
#pragma unroll
                            for (int rr = 0; rr < B; ++rr)
#pragma unroll
                                for (int cc = 0; cc < B; ++cc) r[rr] -= blk[1][rr * B + cc] * yv[cc];
                        }
                        if (depz) {
#pragma unroll
                            for (int rr = 0; rr < B; ++rr)
#pragma unroll
                                for (int cc = 0; cc < B; ++cc) r[rr] -= blk[2][rr * B + cc] * vprev[cc];
                        }
                        double o[B];
#pragma unroll
                        for (int rr = 0; rr < B; ++rr) {
                            double acc = 0.0;
#pragma unroll
                            for (int cc = 0; cc < B; ++cc) acc += blk[3][rr * B + cc] * r[cc];
                            o[rr] = acc;
                        }
#pragma unroll
                        for (int e = 0; e < B; ++e) r[e] = o[e];
                    }
#pragma unroll
                    for (int e = 0; e < B; ++e) { vprev[e] = r[e]; svw[e] = r[e]; }
                    const int sl = UPPER ? NS - 1 - s : s;
                    if (!UPPER) {
                        double* dstp = out_tile + (size_t)out_step_index(sl) * OUT_STEP + (size_t)tl * B;
                        if (B == 2) __stcg(reinterpret_cast<double2*>(dstp), make_double2(r[0], r[B - 1]));
                        else __stcg(dstp, r[0]);
                    }
                    if (a == SK_TI - 1 || b == SK_TJ - 1) {
                        // boundary values for the downstream tiles: each 64-bit word = {tag, half a double}
                        unsigned long long wv[2 * B];
#pragma unroll
                        for (int e = 0; e < B; ++e) {
                            const unsigned long long bits = (unsigned long long)__double_as_longlong(r[e]);
                            wv[2 * e] = ((unsigned long long)tag << 32) | (bits & 0xffffffffull);
                            wv[2 * e + 1] = ((unsigned long long)tag << 32) | (bits >> 32);
                        }
                        if (a == SK_TI - 1) {
                            unsigned long long* w = ll_words(tile, s, 0, b);
#pragma unroll
                            for (int e = 0; e < 2 * B; ++e) st_relaxed(w + e, wv[e]);       // strong 8-byte stores: single-copy atomic, race-free against the relaxed polls
                        }
                        if (b == SK_TJ - 1) {
                            unsigned long long* w = ll_words(tile, s, 1, a);
#pragma unroll
                            for (int e = 0; e < 2 * B; ++e) st_relaxed(w + e, wv[e]);       // strong 8-byte stores: single-copy atomic, race-free against the relaxed polls
                        }
                    }
                    if (UPPER) {
                        double* dn = out + cell_index(kk) * B;
                        if (B == 2) *reinterpret_cast<double2*>(dn) = make_double2(r[0], r[B - 1]);
                        else dn[0] = r[0];
                    }
                }
                compute_barrier();
                SK_STAMP(4 + 2 * c);
                if (t == 0) {
                    mbar_arrive(&empty[stage]);
                    mbar_arrive(&hfree[hq]);
                }
            }
        }
        SK_STAMP(19);
    }
#undef SK_STAMP
}

// ------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------
struct SkewState {
    SkewGrid g{};
    int b = 0;
    double *Lsk = nullptr, *Usk = nullptr;
    double* Dinv = nullptr;                // inverted diagonal blocks of the ILU(0) factorisation, natural layout [n][b*b]
    int diag_grid = 0;                     // co-resident CTAs of ilu_diag_kernel
    int *order_lo = nullptr, *order_up = nullptr;
    unsigned long long* ctl = nullptr;       // [0],[1]: tile tickets of the lower / upper sweep
    unsigned long long seq_lo = 0, seq_up = 0;
    long long* trace = nullptr;            // 2 kernels x 2 tiles x 64 chunks x 24 stamps (DMX_SK_TRACE=1)
    unsigned long long* ll = nullptr;      // flag-in-data halo words [tile][step][dir 2][edge 16][2*b]
    unsigned int ll_seq = 0;               // tag counter of the flag-in-data sweeps
    bool diag_only = false;                // block-diagonal Jacobian: the factorisation is Dinv alone, no streams were built
};

// ILU(0) application for a block-diagonal Jacobian (explicit tracer step): every L and U block is exactly zero, so the lower
// sweep returns its right-hand side and the upper sweep v_i = Dinv_i * y_i (sum from 0) -- the same bits as the sweeps,
// without streaming two sets of zero factors
template <int B>
__global__ void __launch_bounds__(256) dinv_apply_kernel(size_t n, const double* __restrict__ Dinv, const double* __restrict__ d,
                                                         double* __restrict__ v)
{
    const size_t I = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (I >= n) return;
    double r[B], o[B];
#pragma unroll
    for (int e = 0; e < B; ++e) r[e] = d[I * B + e];
#pragma unroll
    for (int rr = 0; rr < B; ++rr) {
        double acc = 0.0;
#pragma unroll
        for (int cc = 0; cc < B; ++cc) acc += Dinv[I * B * B + rr * B + cc] * r[cc];
        o[rr] = acc;
    }
#pragma unroll
    for (int e = 0; e < B; ++e) v[I * B + e] = o[e];
}

template <int B, bool UPPER>
static int sweep_launch_ll(dmx_ctx* ctx, SkewState* st, const double* stream, double* out, unsigned int tag, const int* order,
                           unsigned long long* tick, unsigned long long base, long long* trace)
{
    using LY = SkewLayout<B, UPPER>;
    const SkewGrid& g = st->g;
    auto kern = ilu_sweep_kernel<B, UPPER>;
    constexpr int SK_R = SkHalo<B>::R;
    const size_t smem = ((size_t)LY::S * LY::STEP_DOUBLES + 2 * (SK_TJ + 1) * (SK_TI + 1) * B + SK_R * (SK_TI + SK_TJ) * B) * sizeof(double) +
                        (2 * LY::S + 2 * SK_R) * sizeof(uint64_t);
    DMX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<g.ntiles, SK_THREADS + 64, smem, ctx->stream>>>(g, stream, out, st->ll, tag, order, tick, base, trace);
    DMX_CHECK_LAUNCH();
    return 0;
}

bool sk_supported(const dmx_ctx* ctx)
{
    if (!ctx->has_grid) return false;
    for (int a = 0; a < ctx->dim; ++a)
        if (ctx->nc[a] < 3) return false;      // smaller boxes alias stencil columns (fill inside the pattern): generic path
    return true;
}

void sk_free(dmx_ctx* ctx)
{
    SkewState* st = static_cast<SkewState*>(ctx->skew);
    if (!st) return;
    cudaFree(st->Lsk); cudaFree(st->Usk); cudaFree(st->Dinv); cudaFree(st->ll); cudaFree(st->order_lo); cudaFree(st->order_up); cudaFree(st->ctl);
    if (st->trace) cudaFree(st->trace);
    delete st;
    ctx->skew = nullptr;
}

int sk_setup(dmx_ctx* ctx)
{
    sk_free(ctx);
    if (!sk_supported(ctx)) return 0;
    SkewState* st = new SkewState;
    ctx->skew = st;
    SkewGrid& g = st->g;
    g.nx = ctx->nc[0]; g.ny = ctx->nc[1]; g.nz = ctx->nc[2];
    g.ntx = (g.nx + SK_TI - 1) / SK_TI; g.nty = (g.ny + SK_TJ - 1) / SK_TJ;
    g.ntiles = g.ntx * g.nty;
    g.NS = g.nz + SK_TI + SK_TJ - 2;
    st->b = ctx->b;
    const int BB = ctx->b * ctx->b;
    const size_t slots = (size_t)g.ntiles * g.NS * SK_THREADS;
    // streams [factors | vector] per step: the vector slots of Lsk hold the right-hand side, those of Usk the lower sweep's result
    DMX_CUDA(cudaMalloc((void**)&st->Lsk, slots * (3 * BB + ctx->b) * sizeof(double)));
    DMX_CUDA(cudaMalloc((void**)&st->Usk, slots * (4 * BB + ctx->b) * sizeof(double)));
    DMX_CUDA(cudaMalloc((void**)&st->Dinv, (size_t)ctx->n * BB * sizeof(double)));
    DMX_CUDA(cudaMemsetAsync(st->Lsk, 0, slots * (3 * BB + ctx->b) * sizeof(double), ctx->stream));
    DMX_CUDA(cudaMemsetAsync(st->Usk, 0, slots * (4 * BB + ctx->b) * sizeof(double), ctx->stream));
    {
        const size_t words = (size_t)g.ntiles * g.NS * 2 * SK_TI * 2 * ctx->b;
        DMX_CUDA(cudaMalloc((void**)&st->ll, words * sizeof(unsigned long long)));
        DMX_CUDA(cudaMemsetAsync(st->ll, 0, words * sizeof(unsigned long long), ctx->stream));
    }
    std::vector<int> lo(g.ntiles), up(g.ntiles);
    for (int q = 0; q < g.ntiles; ++q) lo[q] = up[q] = q;
    auto key = [&](int q) { return (q % g.ntx) + (q / g.ntx); };
    std::stable_sort(lo.begin(), lo.end(), [&](int x, int y) { return key(x) < key(y); });
    std::stable_sort(up.begin(), up.end(), [&](int x, int y) { return key(x) > key(y); });
    DMX_CUDA(cudaMalloc((void**)&st->order_lo, g.ntiles * sizeof(int)));
    DMX_CUDA(cudaMalloc((void**)&st->order_up, g.ntiles * sizeof(int)));
    DMX_CUDA(cudaMemcpy(st->order_lo, lo.data(), g.ntiles * sizeof(int), cudaMemcpyHostToDevice));
    DMX_CUDA(cudaMemcpy(st->order_up, up.data(), g.ntiles * sizeof(int), cudaMemcpyHostToDevice));
    DMX_CUDA(cudaMalloc((void**)&st->ctl, 2 * sizeof(unsigned long long)));
    DMX_CUDA(cudaMemset(st->ctl, 0, 2 * sizeof(unsigned long long)));
    {
        const char* env = getenv("DMX_SK_TRACE");
        if (env && env[0] == '1') {
            DMX_CUDA(cudaMalloc((void**)&st->trace, 2 * 2 * 64 * 24 * sizeof(long long)));
            DMX_CUDA(cudaMemset(st->trace, 0, 2 * 2 * 64 * 24 * sizeof(long long)));
        }
    }
    return 0;
}

// ILU(0) factorisation of ctx->d_J on the structured pattern: diagonal recurrence, then the two sweep streams
template <int B>
static int sk_factor_t(dmx_ctx* ctx, SkewState* st)
{
    const SkewGrid& g = st->g;
    auto kern = ilu_diag_kernel<B>;
    if (!st->diag_grid) {
        int perSm = 0;
        DMX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kern, 256, 0));
        st->diag_grid = std::max(1, std::min(perSm, 2)) * ctx->num_sms;
    }
    DMX_CUDA(cudaMemsetAsync(ctx->d_flag, 0, sizeof(int), ctx->stream));
    if (ctx->jac_diagonal || ctx->ssor_factorised) {
        // Dinv_i = A_ii^-1: exact ILU(0) of a block-diagonal matrix, or SeqSSOR in factorised form (L~ = L D^-1, U = D + U)
        ilu_diag_only_kernel<B><<<(unsigned)((ctx->n + 255) / 256), 256, 0, ctx->stream>>>(ctx->n, ctx->d_diag, ctx->d_J, st->Dinv, ctx->d_flag);
        DMX_CHECK_LAUNCH();
        st->diag_only = ctx->jac_diagonal && !ctx->ssor_factorised;
        if (st->diag_only) return 0;              // applied by dinv_apply_kernel, no streams needed
        const unsigned grid = (unsigned)((size_t)g.ntiles * g.NS);
        ilu_skew_kernel<B><<<grid, SK_THREADS, 0, ctx->stream>>>(g, ctx->d_diag, ctx->d_J, st->Dinv, st->Lsk, st->Usk);
        DMX_CHECK_LAUNCH();
        return 0;
    }
    st->diag_only = false;
    DMX_CUDA(cudaMemsetAsync(ctx->d_barrier, 0, sizeof(unsigned int), ctx->stream));
    SkewGrid gg = g;
    const int* diag = ctx->d_diag;
    const double* A = ctx->d_J;
    double* Dinv = st->Dinv;
    int* flag = ctx->d_flag;
    unsigned int* barrier = ctx->d_barrier;
    void* params[] = {(void*)&gg, (void*)&diag, (void*)&A, (void*)&Dinv, (void*)&flag, (void*)&barrier};
    DMX_CUDA(cudaLaunchCooperativeKernel((const void*)kern, dim3(st->diag_grid), dim3(256), params, 0, ctx->stream));
    ctx->launches++;
    const unsigned grid = (unsigned)((size_t)g.ntiles * g.NS);
    ilu_skew_kernel<B><<<grid, SK_THREADS, 0, ctx->stream>>>(g, ctx->d_diag, ctx->d_J, st->Dinv, st->Lsk, st->Usk);
    DMX_CHECK_LAUNCH();
    return 0;
}

int sk_factor(dmx_ctx* ctx)
{
    SkewState* st = static_cast<SkewState*>(ctx->skew);
    return ctx->b == 2 ? sk_factor_t<2>(ctx, st) : sk_factor_t<1>(ctx, st);
}

// factorised BCRS values into `out` (device, nnzb*b*b doubles), the layout of the generic path
int sk_export_bcrs(dmx_ctx* ctx, double* out)
{
    SkewState* st = static_cast<SkewState*>(ctx->skew);
    const unsigned grid = (unsigned)((ctx->n + 255) / 256);
    if (ctx->b == 2) ilu_export_kernel<2><<<grid, 256, 0, ctx->stream>>>(ctx->n, ctx->d_rowptr, ctx->d_colidx, ctx->d_J, st->Dinv, out);
    else ilu_export_kernel<1><<<grid, 256, 0, ctx->stream>>>(ctx->n, ctx->d_rowptr, ctx->d_colidx, ctx->d_J, st->Dinv, out);
    DMX_CHECK_LAUNCH();
    return 0;
}

template <int B>
static int sk_apply_t(dmx_ctx* ctx, SkewState* st, const double* d, double* v)
{
    const SkewGrid& g = st->g;
    if (st->diag_only) {
        dinv_apply_kernel<B><<<(unsigned)((ctx->n + 255) / 256), 256, 0, ctx->stream>>>((size_t)ctx->n, st->Dinv, d, v);
        DMX_CHECK_LAUNCH();
        return 0;
    }
    unsigned long long* tick_lo = st->ctl;
    unsigned long long* tick_up = st->ctl + 1;
    // The device ticket counters advance by ntiles per sweep; the host mirrors (seq_lo / seq_up) advance only once the sweep
    // that consumes the tickets has been enqueued.  If a launch is refused the counters are re-zeroed on both sides, so a
    // recoverable launch error can never leave a later sweep with a ticket base the device does not share.
    const unsigned long long base_lo = st->seq_lo * (unsigned long long)g.ntiles;
    const unsigned long long base_up = st->seq_up * (unsigned long long)g.ntiles;
    // tags of the halo words: never 0 (the buffer starts zeroed), different for every sweep
    const unsigned int tag_lo = 2 * st->ll_seq + 1, tag_up = 2 * st->ll_seq + 2;
    st->ll_seq = (st->ll_seq + 1) % 0x7ffffff0u;
    auto resync = [&](int rc) {
        cudaMemsetAsync(st->ctl, 0, 2 * sizeof(unsigned long long), ctx->stream);
        st->seq_lo = st->seq_up = 0;
        return rc;
    };
    vec_skew_kernel<B><<<(unsigned)((size_t)g.ntiles * g.NS), SK_THREADS, 0, ctx->stream>>>(g, d, st->Lsk);
    {
        ctx->launches++;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return resync(fail(ctx, DMX_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(e)));
    }
    if (int rc = sweep_launch_ll<B, false>(ctx, st, st->Lsk, st->Usk, tag_lo, st->order_lo, tick_lo, base_lo, st->trace)) return resync(rc);
    st->seq_lo++;
    if (int rc = sweep_launch_ll<B, true>(ctx, st, st->Usk, v, tag_up, st->order_up, tick_up, base_up, st->trace ? st->trace + 2 * 64 * 24 : nullptr))
        return resync(rc);
    st->seq_up++;
    return 0;
}

int sk_apply(dmx_ctx* ctx, const double* d, double* v)
{
    SkewState* st = static_cast<SkewState*>(ctx->skew);
    return ctx->b == 2 ? sk_apply_t<2>(ctx, st, d, v) : sk_apply_t<1>(ctx, st, d, v);
}

int sk_trace_read(dmx_ctx* ctx, long long* out)
{
    SkewState* st = static_cast<SkewState*>(ctx->skew);
    if (!st || !st->trace) return fail(ctx, DMX_ERR_USAGE, "no sweep trace (set DMX_SK_TRACE=1 before creating the grid)");
    DMX_CUDA(cudaStreamSynchronize(ctx->stream));
    DMX_CUDA(cudaMemcpy(out, st->trace, 2 * 2 * 64 * 24 * sizeof(long long), cudaMemcpyDeviceToHost));
    return 0;
}

} // namespace SK_VNS

#if SK_TILE == 16
namespace sk8 {
void sk_free(dmx_ctx* ctx);
int sk_setup(dmx_ctx* ctx);
int sk_factor(dmx_ctx* ctx);
int sk_export_bcrs(dmx_ctx* ctx, double* out);
int sk_apply(dmx_ctx* ctx, const double* d, double* v);
int sk_trace_read(dmx_ctx* ctx, long long* out);
}
// Tile edge of a context: 8 x 8 tiles for in-plane boxes up to DMX_SK8_MAX_EDGE (default 64) cells per axis, 16 x 16 otherwise.
// Same arithmetic, same bits.
static int sk_choose_tile(const dmx_ctx* ctx)
{
    static int max_edge = -1;
    if (max_edge < 0) {
        const char* env = getenv("DMX_SK8_MAX_EDGE");
        max_edge = env ? atoi(env) : 64;
    }
    return std::max(ctx->nc[0], ctx->nc[1]) <= max_edge ? 8 : 16;
}
void sk_free(dmx_ctx* ctx) { if (ctx->sk_tile == 8) sk8::sk_free(ctx); else sk16::sk_free(ctx); }
int sk_setup(dmx_ctx* ctx)
{
    sk_free(ctx);
    ctx->sk_tile = sk_choose_tile(ctx);
    return ctx->sk_tile == 8 ? sk8::sk_setup(ctx) : sk16::sk_setup(ctx);
}
int sk_factor(dmx_ctx* ctx) { return ctx->sk_tile == 8 ? sk8::sk_factor(ctx) : sk16::sk_factor(ctx); }
int sk_export_bcrs(dmx_ctx* ctx, double* out) { return ctx->sk_tile == 8 ? sk8::sk_export_bcrs(ctx, out) : sk16::sk_export_bcrs(ctx, out); }
int sk_apply(dmx_ctx* ctx, const double* d, double* v) { return ctx->sk_tile == 8 ? sk8::sk_apply(ctx, d, v) : sk16::sk_apply(ctx, d, v); }
int sk_trace_read(dmx_ctx* ctx, long long* out) { return ctx->sk_tile == 8 ? sk8::sk_trace_read(ctx, out) : sk16::sk_trace_read(ctx, out); }
#endif

} // namespace dmx
