// assembly.cu -- CCTpfa residual + numerically differentiated Jacobian, written straight into the BCRS values.
//
// Replaces FVAssembler::assembleJacobianAndResidual (dumux/assembly/fvassembler.hh:179-207,462-510) and the
// per-element CCLocalAssembler<numeric,implicit>::assembleJacobianAndResidualImpl
// (dumux/assembly/cclocalassembler.hh:164-348) with a ROW-GATHER formulation: one thread owns block row I and
// produces res[I], A[I][I] and all A[I][J].  A[I][J] = d(flux_I over face IJ)/d(u_J) with eps(u_J) is exactly what
// element J's column pass scatters into A[I][J] in the reference (:324-333), so no colouring and no atomics.
//
// Kernels:
//   transmissibility_kernel  setup: t_ij of every +face (flux/cctpfa/darcyslaw.hh:218-259,
//                            discretization/cellcentered/tpfa/computetransmissibility.hh:69-80)
//   volvars_kernel           per-cell secondary variables at the base and FD-deflected states
//                            (porousmediumflow/2p/volumevariables.hh:75-190, 1p/volumevariables.hh:65-125), so each
//                            material-law pow is evaluated once per cell instead of once per stencil visit
//   assemble_kernel          storage + source + 2*dim TPFA fluxes at base and deflected states, FD quotients
//                            (common/numericdifferentiation.hh:67-123), BCRS write
//
// Floating-point contract: compiled with -fmad=false; every expression is written in the operation order of
// the reference (see the op-order notes inline) so results are reproducible against the CPU oracle.
#include <algorithm>
#include <cfloat>
#include <cstdlib>

#include "common.cuh"

namespace dmx {

// ------------------------------------------------------------------------------------------------
// setup: transmissibilities of the +faces.  tij = A * ti*tj/(ti+tj) (0 if ti*tj <= 0)
// ------------------------------------------------------------------------------------------------
// permeability entering a face with normal e_a: K_aa of a diagonal tensor, else the scalar
__device__ __forceinline__ double perm_axis(const AsmParams& P, int a, size_t C, double Kscalar)
{
    return P.Kaxis[a] ? __ldg(P.Kaxis[a] + C) : Kscalar;
}

__global__ void transmissibility_kernel(AsmParams P, double* t0, double* t1, double* t2)
{
    const int I = blockIdx.x * blockDim.x + threadIdx.x;
    if (I >= P.n) return;
    const int i = I % P.nc[0];
    const int j = (I / P.nc[0]) % P.nc[1];
    const int k = I / (P.nc[0] * P.nc[1]);
    const int c[3] = {i, j, k};
    const int stride[3] = {1, P.nc[0], P.nc[0] * P.nc[1]};
    double* out[3] = {t0, t1, t2};
    const double KI = P.K[I];
    for (int a = 0; a < P.dim; ++a) {
        double tij = 0.0;
        if (c[a] + 1 < P.nc[a]) {
            double area = 1.0;
            for (int d = 0; d < P.dim; ++d)
                if (d != a) area *= P.width[d][c[d]];
            const double ti = perm_axis(P, a, I, KI) * P.extrusion * P.gf_hi[a][c[a]];
            const double tj = perm_axis(P, a, I + stride[a], P.K[I + stride[a]]) * P.extrusion * P.gf_lo[a][c[a] + 1];
            if (ti * tj <= 0.0) tij = 0;
            else tij = area * (ti * tj) / (ti + tj);
        }
        out[a][I] = tij;
    }
}

// FD step and deflected value: numericepsilon.hh:47-51, numericdifferentiation.hh:36-41,67-123
__device__ __forceinline__ double fd_eps(const AsmParams& P, double x, int pv)
{
    return P.mag[pv] > 0.0 ? P.base_eps * P.mag[pv] : P.base_eps * (fabs(x) + 1.0);
}
__device__ __forceinline__ double fd_deflect(int method, int k, double x0, double eps)
{
    if (method == 1) return x0 + eps;
    if (method == -1) return x0 - eps;
    if (k == 0) return x0 + eps;
    if (k == 1) return x0 - eps;
    if (k == 2) return x0 - 2.0 * eps;
    return x0 + 2.0 * eps;
}
// per-cell state entering a flux evaluation
template <int NPH>
struct CellState {
    double p[NPH];
    double up[NPH];    // rho*mobility, the upwinded quantity (immiscible/localresidual.hh:113-114)
    double rho[NPH];
};

// geometry / solution-independent data of one face seen from cell I
template <int NPH>
struct FaceData {
    double tij, area, tJ, alphaI, alphaJ;
    bool interior, grav;
    double c1[NPH], c2[NPH];   // gravity terms for constant density
};

// storage term: immiscible/localresidual.hh:64-83: porosity*density*saturation per phase
template <int MODEL, int NPH>
__device__ __forceinline__ void storage_term(double phiE, const CellState<NPH>& s, double Sn, double* st)
{
    if constexpr (MODEL == DMX_MODEL_2P) {
        st[0] = phiE * s.rho[0] * (1 - Sn);
        st[NPH - 1] = phiE * s.rho[NPH - 1] * Sn;
    } else {
        st[0] = phiE * s.rho[0] * 1.0;
    }
}

// ================================================================================================
// Fused tile kernel (the production path).
//
// One CTA owns a 32 x 8 column of cells and marches along the last axis over `zchunk` layers.  Per layer it
//   1. evaluates the secondary variables (TwoPVolumeVariables::update / OnePVolumeVariables::update) of the next layer's
//      tile + one-cell halo at the base and FD-deflected states ONCE into a three-layer shared-memory ring
//      (material-law pows share their log, see law_eval3), together with K and the reciprocal FD denominators;
//   2. assembles the rows of its 256 cells from shared memory: storage, source, 2*dim TPFA fluxes at the base state,
//      at the own deflections (diagonal block) and at each neighbour's deflections (off-diagonal blocks), FD quotients
//      through div_by (correctly rounded, shared reciprocal), 32-byte block stores into the BCRS values.
// HBM traffic per cell: cur, prev, K, phi, region, 3 transmissibilities, rowptr in; residual + 7 blocks out -- no
// secondary-variable records in global memory.  Arithmetic and operation order follow the reference's local assembler
// (bit-identical to the CPU oracle).
// ================================================================================================
constexpr int AT_TX = 32, AT_TY = 8, AT_THREADS = AT_TX * AT_TY;
constexpr int AT_HX = AT_TX + 2, AT_HY = AT_TY + 2, AT_HC = AT_HX * AT_HY;

// shared-memory record of one cell, SoA: field f of halo slot h lives at rec[f * AT_HC + h]
template <int MODEL, bool TABLE, int NDR>
struct RecLayout {
    static constexpr int NB = (MODEL == DMX_MODEL_2P) ? 2 : 1;
    static constexpr int NS = (MODEL == DMX_MODEL_2P) ? 3 : (TABLE ? 2 : 0);   // 2p: pc, rho_w*mob_w, rho_n*mob_n; 1p table: rho, rho*mob
    static constexpr int U = 0;               // primary variables
    static constexpr int EPS = NB;            // FD step per primary variable (SIGNED for one-sided differences, see fd_step)
    static constexpr int RD = 2 * NB;         // reciprocal FD denominators per primary variable
    static constexpr int KF = 3 * NB;         // permeability
    static constexpr int ST = 3 * NB + 1;     // states: base, then one per deflection of the state-changing primary variable
    static constexpr int NF = ST + NS * (1 + NDR);
};

// Step as stored per cell.  One-sided differences (ND == 1) keep the SIGNED step (backward: -eps): x0 + (-eps) == x0 - eps and
// (f1 - f0)/(-eps) == (f0 - f1)/eps bit for bit, so neither the deflection nor the quotient needs to look at the method again.
template <int ND>
__device__ __forceinline__ double fd_step(const AsmParams& P, double x, int pv)
{
    const double eps = fd_eps(P, x, pv);
    return (ND == 1 && P.fd_method < 0) ? -eps : eps;
}
template <int ND>
__device__ __forceinline__ double fd_defl(const AsmParams& P, int k, double x0, double step)
{
    if (ND == 1) return x0 + step;
    return fd_deflect(P.fd_method, k, x0, step);
}
template <int ND>
__device__ __forceinline__ double fd_den(double eps)
{
    if (ND == 1) return eps;              // delta = 0 + eps
    if (ND == 2) return eps + eps;        // delta = (0 + eps) + eps
    return 12.0 * eps;
}
// fd_quotient with the division by the step replaced by div_by (same bits)
template <int ND>
__device__ __forceinline__ double fd_quotient_r(double f0, const double* f, double eps, double rden)
{
    double d;
    if (ND == 1) {
        d = f[0];
        d -= f0;
    } else if (ND == 2) {
        d = f[0];
        d -= f[1];
    } else {
        d = f[0];
        d -= f[1];
        d *= 8.0;
        d += f[2];
        d -= f[3];
    }
    return div_by(d, fd_den<ND>(eps), rden);
}

// secondary variables of cell C into its shared-memory record
template <int MODEL, bool TABLE, int ND, bool JAC>
__device__ __forceinline__ void fill_cell(const AsmParams& P, const MaterialLaw* slaws, size_t C, double* rec)
{
    using L = RecLayout<MODEL, TABLE, JAC ? ND : 0>;
    rec[L::KF * AT_HC] = __ldg(P.K + C);
    if constexpr (MODEL == DMX_MODEL_2P) {
        const double2 u = __ldg(reinterpret_cast<const double2*>(P.cur) + C);
        rec[(L::U + 0) * AT_HC] = u.x;
        rec[(L::U + 1) * AT_HC] = u.y;
        // p0s1 formulation with a per-region wetting phase w (2p/volumevariables.hh:87-96,132-152): the law is evaluated at the
        // saturation of phase w, krw belongs to phase w, and p1 = p0 + pc (w = 0) or p0 - pc (w = 1; stored as -pc: x - y == x + (-y))
        const MaterialLaw& law = slaws[__ldg(P.region + C)];
        const bool w1 = law.wetting != 0;
        Law3 c = law_eval3_call(&law, w1 ? u.y : 1 - u.y);
        rec[(L::ST + 0) * AT_HC] = w1 ? -c.pc : c.pc;
        rec[(L::ST + 1) * AT_HC] = P.rho[0] * div_by(w1 ? c.krn : c.krw, P.mu[0], P.rmu[0]);
        rec[(L::ST + 2) * AT_HC] = P.rho[1] * div_by(w1 ? c.krw : c.krn, P.mu[1], P.rmu[1]);
        if constexpr (JAC) {
            const double epsP = fd_step<ND>(P, u.x, 0), epsS = fd_step<ND>(P, u.y, 1);
            rec[(L::EPS + 0) * AT_HC] = epsP;
            rec[(L::EPS + 1) * AT_HC] = epsS;
            rec[(L::RD + 0) * AT_HC] = 1.0 / fd_den<ND>(epsP);
            rec[(L::RD + 1) * AT_HC] = 1.0 / fd_den<ND>(epsS);
#pragma unroll
            for (int k = 0; k < ND; ++k) {
                const double Snk = fd_defl<ND>(P, k, u.y, epsS);
                c = law_eval3_call(&law, w1 ? Snk : 1 - Snk);
                rec[(L::ST + 3 * (1 + k) + 0) * AT_HC] = w1 ? -c.pc : c.pc;
                rec[(L::ST + 3 * (1 + k) + 1) * AT_HC] = P.rho[0] * div_by(w1 ? c.krn : c.krw, P.mu[0], P.rmu[0]);
                rec[(L::ST + 3 * (1 + k) + 2) * AT_HC] = P.rho[1] * div_by(w1 ? c.krw : c.krn, P.mu[1], P.rmu[1]);
            }
        }
    } else {
        const double p = __ldg(P.cur + C);
        rec[L::U * AT_HC] = p;
        double eps = 0.0;
        if constexpr (JAC) {
            eps = fd_step<ND>(P, p, 0);
            rec[L::EPS * AT_HC] = eps;
            rec[L::RD * AT_HC] = 1.0 / fd_den<ND>(eps);
        }
        if constexpr (TABLE) {
            double rho, mu;
            table_interp2(P.table, p, &rho, &mu);
            rec[(L::ST + 0) * AT_HC] = rho;
            rec[(L::ST + 1) * AT_HC] = rho * (1.0 / mu);
            if constexpr (JAC) {
#pragma unroll
                for (int k = 0; k < ND; ++k) {
                    table_interp2(P.table, fd_defl<ND>(P, k, p, eps), &rho, &mu);
                    rec[(L::ST + 2 * (1 + k) + 0) * AT_HC] = rho;
                    rec[(L::ST + 2 * (1 + k) + 1) * AT_HC] = rho * (1.0 / mu);
                }
            }
        }
    }
}

// state of a cell at the base point (pv < 0) or at deflection (pv, k), from its shared-memory record
template <int MODEL, bool TABLE, int NDR, int NPH>
__device__ __forceinline__ void tile_state(const AsmParams& P, const double* rec, int pv, int k, const double* uC, const double* epsC,
                                           CellState<NPH>& s, double* Sn_out)
{
    using L = RecLayout<MODEL, TABLE, NDR>;
    if constexpr (MODEL == DMX_MODEL_2P) {
        double pw = uC[0], Sn = uC[1];
        int st = 0;
        if (pv == 0) pw = fd_defl<NDR>(P, k, pw, epsC[0]);
        if (pv == 1) { Sn = fd_defl<NDR>(P, k, Sn, epsC[1]); st = 1 + k; }
        const double pc = rec[(L::ST + 3 * st + 0) * AT_HC];
        s.p[0] = pw;
        s.p[NPH - 1] = pw + pc;
        s.up[0] = rec[(L::ST + 3 * st + 1) * AT_HC];
        s.up[NPH - 1] = rec[(L::ST + 3 * st + 2) * AT_HC];
        s.rho[0] = P.rho[0];
        s.rho[NPH - 1] = P.rho[1];
        *Sn_out = Sn;
    } else {
        double p = uC[0];
        if (pv == 0) p = fd_defl<NDR>(P, k, p, epsC[0]);
        s.p[0] = p;
        if constexpr (TABLE) {
            const int st = (pv == 0) ? 1 + k : 0;
            s.rho[0] = rec[(L::ST + 2 * st + 0) * AT_HC];
            s.up[0] = rec[(L::ST + 2 * st + 1) * AT_HC];
        } else {
            s.rho[0] = P.rho[0];
            s.up[0] = P.rho[0] * (1.0 / P.mu[0]);
        }
        *Sn_out = 0.0;
    }
}

// as face_flux, the division by the neighbour's half transmissibility through its reciprocal (TABLE only)
template <int NPH, bool TABLE>
__device__ __forceinline__ void face_flux_r(const FaceData<NPH>& F, double rtJ, const CellState<NPH>& sI, const CellState<NPH>& sJ,
                                            bool fullUpwind, double w, double* out)
{
#pragma unroll
    for (int ph = 0; ph < NPH; ++ph) {
        double f = F.tij * (sI.p[ph] - sJ.p[ph]);
        if (F.grav) {
            if constexpr (TABLE) {
                const double rho = F.interior ? (sI.rho[ph] + sJ.rho[ph]) * 0.5 : sJ.rho[ph];
                f = f + rho * F.area * F.alphaI;
                if (F.interior) f -= div_by(rho * F.tij, F.tJ, rtJ) * (F.alphaI - F.alphaJ);
            } else {
                f = f + F.c1[ph];
                if (F.interior) f -= F.c2[ph];
            }
        }
        double mult;
        if (fullUpwind) mult = signbit(f) ? sJ.up[ph] : sI.up[ph];
        else if (signbit(f)) mult = w * sJ.up[ph] + (1.0 - w) * sI.up[ph];
        else mult = w * sI.up[ph] + (1.0 - w) * sJ.up[ph];
        out[ph] = f * mult;
    }
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <int NB>
__device__ __forceinline__ void store_block(double* dst, const double (&blk)[NB][NB])
{
    if constexpr (NB == 2) {
        asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(dst), "d"(blk[0][0]), "d"(blk[0][1]), "d"(blk[1][0]), "d"(blk[1][1])
                     : "memory");
    } else
        dst[0] = blk[0][0];
}

template <int MODEL, bool TABLE, int ND, bool JAC, int DIM>
__global__ void __launch_bounds__(AT_THREADS, (ND == 1) ? 2 : 1) assemble_tile_kernel(const AsmParams P)
{
    constexpr int NDR = JAC ? ND : 0;
    using L = RecLayout<MODEL, TABLE, NDR>;
    constexpr int NB = L::NB, NPH = NB;
    constexpr int NDE = JAC ? ND : 1;       // array extents (unused when !JAC)
    extern __shared__ __align__(16) double sm[];
    MaterialLaw* slaws = reinterpret_cast<MaterialLaw*>(sm + 3 * L::NF * AT_HC);

    const int t = threadIdx.x, li = t & (AT_TX - 1), lj = t / AT_TX;
    const int nx = P.nc[0], ny = P.nc[1], nz = P.nc[2];
    constexpr int dim = DIM, va = DIM - 1;
    const int ntx = (nx + AT_TX - 1) / AT_TX;
    const int i0 = ((int)blockIdx.x % ntx) * AT_TX, j0 = ((int)blockIdx.x / ntx) * AT_TY;
    const int kz0 = (int)blockIdx.y * P.zchunk, kz1 = min(nz, kz0 + P.zchunk);
    const int i = i0 + li, j = j0 + lj;
    const bool own = i < nx && j < ny;
    const bool fullUpwind = (P.upwind_weight == 1.0);
    const double w = P.upwind_weight;
    const double extr = P.extrusion;

    if constexpr (MODEL == DMX_MODEL_2P) {
        const int words = P.nlaws * (int)(sizeof(MaterialLaw) / sizeof(double));
        for (int q = t; q < words; q += AT_THREADS) reinterpret_cast<double*>(slaws)[q] = reinterpret_cast<const double*>(P.laws)[q];
        __syncthreads();
    }

    auto slot_of = [&](int k) { return sm + ((k + 3) % 3) * (L::NF * AT_HC); };
    auto fill = [&](int k, bool core_only) {
        if (k < 0 || k >= nz) return;
        double* dst = slot_of(k);
        for (int h = t; h < AT_HC; h += AT_THREADS) {
            const int hi = h % AT_HX, hj = h / AT_HX;
            const bool xh = (hi == 0 || hi == AT_HX - 1), yh = (hj == 0 || hj == AT_HY - 1);
            if ((xh && yh) || (core_only && (xh || yh))) continue;
            const int ci = i0 - 1 + hi, cj = j0 - 1 + hj;
            if (ci < 0 || ci >= nx || cj < 0 || cj >= ny) continue;
            fill_cell<MODEL, TABLE, ND, JAC>(P, slaws, (size_t)ci + (size_t)nx * (cj + (size_t)ny * k), dst + h);
        }
    };

    // in-plane geometry of this thread's column
    const int h0 = (lj + 1) * AT_HX + (li + 1);
    const double wx = own ? P.width[0][i] : 1.0;
    const double wy = (own && dim > 1) ? P.width[1][j] : 1.0;
    bool exxy[4];
    exxy[0] = i > 0; exxy[1] = i + 1 < nx; exxy[2] = (dim > 1) && j > 0; exxy[3] = (dim > 1) && j + 1 < ny;

    fill(kz0 - 1, true);
    fill(kz0, false);
    double tz_below = 0.0;
    if (own && dim == 3 && kz0 > 0) tz_below = P.tij[2][(size_t)i + (size_t)nx * (j + (size_t)ny * (kz0 - 1))];

    for (int k = kz0; k < kz1; ++k) {
        // global data of this thread's own row, requested before the secondary-variable phase so that the DRAM latency
        // overlaps it; the next layer's lines are pulled into L2 one iteration ahead
        const size_t I = (size_t)i + (size_t)nx * (j + (size_t)ny * k);
        double g_tij[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0}, g_phi = 0.0, g_prev = 0.0;
        int g_rowStart = 0;
        if (own) {
            if (exxy[0]) g_tij[0] = __ldg(P.tij[0] + I - 1);
            if (exxy[1]) g_tij[1] = __ldg(P.tij[0] + I);
            if (exxy[2]) g_tij[2] = __ldg(P.tij[1] + I - nx);
            if (exxy[3]) g_tij[3] = __ldg(P.tij[1] + I);
            if (dim > 2 && k + 1 < nz) g_tij[5] = __ldg(P.tij[2] + I);
            g_tij[4] = tz_below;
            if constexpr (JAC) g_rowStart = __ldg(P.rowptr + I);
            if (!P.stationary) {
                g_phi = __ldg(P.phi + I);
                if constexpr (MODEL == DMX_MODEL_2P) g_prev = __ldg(P.prev + I * NB + NB - 1);
                else g_prev = __ldg(P.prev + I);
            }
            if (k + 1 < kz1) {
                const size_t In = I + (size_t)nx * ny;
                prefetch_l2(P.tij[0] + In);
                if (dim > 1) prefetch_l2(P.tij[1] + In);
                if (dim > 2) prefetch_l2(P.tij[2] + In);
                if constexpr (JAC) prefetch_l2(P.rowptr + In);
                if (!P.stationary) { prefetch_l2(P.phi + In); prefetch_l2(P.prev + In * NB); }
            }
        }
        if (own && k + 2 < nz && k + 2 <= kz1) {
            const size_t I2 = I + 2 * (size_t)nx * ny;
            prefetch_l2(P.cur + I2 * NB);
            prefetch_l2(P.K + I2);
            if constexpr (MODEL == DMX_MODEL_2P) prefetch_l2(P.region + I2);
        }
        fill(k + 1, k + 1 == kz1);
        __syncthreads();
        if (own) {
            const int ci[3] = {i, j, k};
            const double* recI = slot_of(k) + h0;
            const double* recN[6] = {recI - 1, recI + 1, recI - AT_HX, recI + AT_HX, slot_of(k - 1) + h0, slot_of(k + 1) + h0};

            double uI[NB], epsI[NB], rdI[NB];
#pragma unroll
            for (int e = 0; e < NB; ++e) {
                uI[e] = recI[(L::U + e) * AT_HC];
                epsI[e] = JAC ? recI[(L::EPS + e) * AT_HC] : 0.0;
                rdI[e] = JAC ? recI[(L::RD + e) * AT_HC] : 0.0;
            }
            CellState<NPH> sI0, sId[NB][NDE];
            double SnI0, SnId[NB][NDE];
            tile_state<MODEL, TABLE, NDR, NPH>(P, recI, -1, 0, uI, epsI, sI0, &SnI0);
            if constexpr (JAC) {
#pragma unroll
                for (int pv = 0; pv < NB; ++pv)
#pragma unroll
                    for (int kk = 0; kk < ND; ++kk) tile_state<MODEL, TABLE, NDR, NPH>(P, recI, pv, kk, uI, epsI, sId[pv][kk], &SnId[pv][kk]);
            }

            const double KI = recI[L::KF * AT_HC];
            const double wz = (dim > 2) ? P.width[2][k] : 1.0;
            // volume = ((1*w0)*w1)*w2 (AxisAlignedCubeGeometry::volume)
            double vol = 1.0;
            vol *= wx;
            if (dim > 1) vol *= wy;
            if (dim > 2) vol *= wz;

            double R0[NB], Rd[NB][NDE][NB];
#pragma unroll
            for (int e = 0; e < NB; ++e) {
                double source = P.q ? P.q[I * NB + e] : 0.0;      // fvlocalresidual.hh:319-333
                source *= vol * extr;
                double r = 0.0;
                r -= source;
                R0[e] = r;
#pragma unroll
                for (int pv = 0; pv < NB; ++pv)
#pragma unroll
                    for (int kk = 0; kk < NDE; ++kk) Rd[pv][kk][e] = r;
            }

            bool ex[6];
            ex[0] = exxy[0]; ex[1] = exxy[1]; ex[2] = exxy[2]; ex[3] = exxy[3];
            ex[4] = (dim > 2) && k > 0;
            ex[5] = (dim > 2) && k + 1 < nz;
            int pos[6], posDiag = 0;
            if constexpr (JAC) {
                const int rowStart = g_rowStart;
                posDiag = rowStart + (ex[4] ? 1 : 0) + (ex[2] ? 1 : 0) + (ex[0] ? 1 : 0);
                pos[4] = rowStart;
                pos[2] = rowStart + (ex[4] ? 1 : 0);
                pos[0] = pos[2] + (ex[2] ? 1 : 0);
                pos[1] = posDiag + 1;
                pos[3] = pos[1] + (ex[1] ? 1 : 0);
                pos[5] = pos[3] + (ex[3] ? 1 : 0);
            }
#pragma unroll
            for (int s = 0; s < 6; ++s) {
                const int a = s >> 1;
                if (a >= dim) continue;
                const bool hi = (s & 1);
                // face area = product of the widths of the other axes, ascending axis order
                double area = 1.0;
                if (a != 0) area *= wx;
                if (a != 1 && dim > 1) area *= wy;
                if (a != 2 && dim > 2) area *= wz;
                FaceData<NPH> F;
                F.area = area;
                F.grav = P.enable_gravity && (a == va);
                const double ng = hi ? -P.gravity : P.gravity;       // n.g with g = -gravity*e_va
                F.alphaI = KI * ng * extr;                           // vtmv(n,K,g)*extrusion (common/math.hh:908-913)
                F.alphaJ = 0.0;
                F.tJ = 1.0;
                double rtJ = 1.0;
                if (ex[s]) {
                    const double* recJ = recN[s];
                    const int cj = hi ? ci[a] + 1 : ci[a] - 1;
                    F.interior = true;
                    F.tij = g_tij[s];
                    if (F.grav) {
                        const double KJ = recJ[L::KF * AT_HC];
                        F.tJ = KJ * extr * (hi ? P.gf_lo[a][cj] : P.gf_hi[a][cj]);
                        rtJ = 1.0 / F.tJ;
                        F.alphaJ = KJ * ng * extr;
                        if (!TABLE) {
#pragma unroll
                            for (int ph = 0; ph < NPH; ++ph) {
                                const double rho = (P.rho[ph] + P.rho[ph]) * 0.5;
                                F.c1[ph] = rho * area * F.alphaI;
                                F.c2[ph] = div_by(rho * F.tij, F.tJ, rtJ) * (F.alphaI - F.alphaJ);
                            }
                        }
                    }
                    double uJ[NB], epsJ[NB];
#pragma unroll
                    for (int e = 0; e < NB; ++e) { uJ[e] = recJ[(L::U + e) * AT_HC]; epsJ[e] = JAC ? recJ[(L::EPS + e) * AT_HC] : 0.0; }
                    CellState<NPH> sJ0;
                    double SnJ;
                    tile_state<MODEL, TABLE, NDR, NPH>(P, recJ, -1, 0, uJ, epsJ, sJ0, &SnJ);
                    double F0[NPH];
                    face_flux_r<NPH, TABLE>(F, rtJ, sI0, sJ0, fullUpwind, w, F0);
#pragma unroll
                    for (int e = 0; e < NB; ++e) R0[e] += F0[e];
                    if constexpr (JAC) {
#pragma unroll
                        for (int pv = 0; pv < NB; ++pv)
#pragma unroll
                            for (int kk = 0; kk < ND; ++kk) {
                                double Fk[NPH];
                                face_flux_r<NPH, TABLE>(F, rtJ, sId[pv][kk], sJ0, fullUpwind, w, Fk);
#pragma unroll
                                for (int e = 0; e < NB; ++e) Rd[pv][kk][e] += Fk[e];
                            }
                        // A[I][J][e][pv] = FD quotient of the face flux w.r.t. u_J[pv] (cclocalassembler.hh:211-218,254-258,331-332)
                        double blk[NB][NB];
#pragma unroll
                        for (int pv = 0; pv < NB; ++pv) {
                            double Fd[ND][NPH];
#pragma unroll
                            for (int kk = 0; kk < ND; ++kk) {
                                CellState<NPH> sJd;
                                double dummy;
                                tile_state<MODEL, TABLE, NDR, NPH>(P, recJ, pv, kk, uJ, epsJ, sJd, &dummy);
                                face_flux_r<NPH, TABLE>(F, rtJ, sI0, sJd, fullUpwind, w, Fd[kk]);
                            }
                            const double rdJ = recJ[(L::RD + pv) * AT_HC];
#pragma unroll
                            for (int e = 0; e < NB; ++e) {
                                double fk[ND];
#pragma unroll
                                for (int kk = 0; kk < ND; ++kk) fk[kk] = Fd[kk][e];
                                blk[e][pv] = fd_quotient_r<ND>(F0[e], fk, epsJ[pv], rdJ);
                            }
                        }
                        store_block<NB>(P.jac + (size_t)pos[s] * (NB * NB), blk);
                    }
                } else {
                    // boundary face: cclocalresidual.hh:64-105
                    int f;   // face index within the side, lower remaining axis fastest
                    if (a == 0) f = ci[1] + ny * ci[2];
                    else if (a == 1) f = ci[0] + nx * ci[2];
                    else f = ci[0] + nx * ci[1];
                    const int type = P.bc_type[s] ? P.bc_type[s][f] : DMX_BC_NEUMANN;
                    if (type == DMX_BC_DIRICHLET) {
                        F.interior = false;
                        const double ti = perm_axis(P, a, I, KI) * extr * (hi ? P.gf_hi[a][ci[a]] : P.gf_lo[a][ci[a]]);
                        F.tij = area * ti;
                        CellState<NPH> sD;
#pragma unroll
                        for (int ph = 0; ph < NPH; ++ph) {
                            sD.p[ph] = P.bc_p[s][(size_t)f * 2 + ph];
                            sD.up[ph] = P.bc_up[s][(size_t)f * 2 + ph];
                            sD.rho[ph] = P.bc_rho[s][(size_t)f * 2 + ph];
                            F.c1[ph] = sD.rho[ph] * area * F.alphaI;
                            F.c2[ph] = 0.0;
                        }
                        double F0[NPH];
                        face_flux_r<NPH, TABLE>(F, rtJ, sI0, sD, fullUpwind, w, F0);
#pragma unroll
                        for (int e = 0; e < NB; ++e) R0[e] += F0[e];
                        if constexpr (JAC) {
#pragma unroll
                            for (int pv = 0; pv < NB; ++pv)
#pragma unroll
                                for (int kk = 0; kk < ND; ++kk) {
                                    double Fk[NPH];
                                    face_flux_r<NPH, TABLE>(F, rtJ, sId[pv][kk], sD, fullUpwind, w, Fk);
#pragma unroll
                                    for (int e = 0; e < NB; ++e) Rd[pv][kk][e] += Fk[e];
                                }
                        }
                    } else if (type == DMX_BC_NEUMANN) {
#pragma unroll
                        for (int e = 0; e < NB; ++e) {
                            double nf = P.bc_neumann[s] ? P.bc_neumann[s][(size_t)f * NB + e] : 0.0;
                            nf *= area * extr;
                            R0[e] += nf;
#pragma unroll
                            for (int pv = 0; pv < NB; ++pv)
#pragma unroll
                                for (int kk = 0; kk < NDE; ++kk) Rd[pv][kk][e] += nf;
                        }
                    }
                    // DMX_BC_NONE: outer face of an overlap cell, no scvf exists (tpfa/fvgridgeometry.hh:272-320)
                }
            }
            tz_below = g_tij[5];

            // storage: fvlocalresidual.hh:274-304: ((S(cur)*extr - S(prev)*extr)*V)/dt, added after flux+source
            if (!P.stationary) {
                const double phiI = g_phi;
                const double phiE = 1.0 - (1.0 - phiI);     // porosity = 1 - inert volume fraction
                double prevSt[NB];
                {
                    CellState<NPH> sP;
                    double SnP = 0.0;
                    if constexpr (MODEL == DMX_MODEL_2P) {
                        SnP = g_prev;
                        sP.rho[0] = P.rho[0];
                        sP.rho[NPH - 1] = P.rho[1];
                    } else {
                        sP.rho[0] = TABLE ? table_interp(P.table, P.table.rho, g_prev) : P.rho[0];
                    }
                    storage_term<MODEL, NPH>(phiE, sP, SnP, prevSt);
#pragma unroll
                    for (int e = 0; e < NB; ++e) prevSt[e] *= extr;
                }
                auto addStorage = [&](const CellState<NPH>& s, double Sn, double* acc) {
                    double st[NB];
                    storage_term<MODEL, NPH>(phiE, s, Sn, st);
#pragma unroll
                    for (int e = 0; e < NB; ++e) {
                        st[e] *= extr;
                        st[e] -= prevSt[e];
                        st[e] *= vol;
                        st[e] = div_by(st[e], P.dt, P.rdt);
                        acc[e] += st[e];
                    }
                };
                addStorage(sI0, SnI0, R0);
                if constexpr (JAC) {
#pragma unroll
                    for (int pv = 0; pv < NB; ++pv)
#pragma unroll
                        for (int kk = 0; kk < ND; ++kk) addStorage(sId[pv][kk], SnId[pv][kk], Rd[pv][kk]);
                }
            }

            bool finite = true;
#pragma unroll
            for (int e = 0; e < NB; ++e) finite = finite && (fabs(R0[e]) <= DBL_MAX);
            if constexpr (NB == 2) reinterpret_cast<double2*>(P.residual)[I] = make_double2(R0[0], R0[NB - 1]);
            else P.residual[I] = R0[0];
            if (!finite) atomicOr(P.flag_nonfinite, 1);

            if constexpr (JAC) {
                double blk[NB][NB];
#pragma unroll
                for (int pv = 0; pv < NB; ++pv)
#pragma unroll
                    for (int e = 0; e < NB; ++e) {
                        double fk[ND];
#pragma unroll
                        for (int kk = 0; kk < ND; ++kk) fk[kk] = Rd[pv][kk][e];
                        blk[e][pv] = fd_quotient_r<ND>(R0[e], fk, epsI[pv], rdI[pv]);
                    }
                store_block<NB>(P.jac + (size_t)posDiag * (NB * NB), blk);
            }
        }
        __syncthreads();
    }
}

// ================================================================================================
// DiffMethod::analytic for the incompressible 1p model: CCLocalAssembler<analytic, implicit> (assembly/cclocalassembler.hh:
// 490-600) + OnePIncompressibleLocalResidual::addFluxDerivatives / addCCDirichletFluxDerivatives
// (porousmediumflow/1p/incompressiblelocalresidual.hh:76-123,204-221): A[I][I] += tij*up over interior and Dirichlet faces
// (-x,+x,-y,+y,-z,+z), A[I][J] -= tij*up, up = density/viscosity; no storage derivative (:51-60), Neumann faces contribute
// nothing.  One thread per row, rows written whole; the residual comes from the JAC = false instantiation of the tile kernel.
// HBM-bound stream: 3 transmissibilities + rowptr in, 7 entries out per row.
// ================================================================================================
template <int DIM>
__global__ void __launch_bounds__(256) onep_analytic_jacobian_kernel(const AsmParams P, double up)
{
    const size_t I = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (I >= (size_t)P.n) return;
    const int nx = P.nc[0], ny = P.nc[1];
    const int ci[3] = {(int)(I % nx), (int)((I / nx) % ny), (int)(I / ((size_t)nx * ny))};
    const size_t stride[3] = {1, (size_t)nx, (size_t)nx * ny};
    bool ex[6];
    int pos[6];
#pragma unroll
    for (int s = 0; s < 6; ++s) ex[s] = (s >> 1) < DIM && ((s & 1) ? ci[s >> 1] + 1 < P.nc[s >> 1] : ci[s >> 1] > 0);
    const int rowStart = P.rowptr[I];
    const int posDiag = rowStart + (ex[4] ? 1 : 0) + (ex[2] ? 1 : 0) + (ex[0] ? 1 : 0);
    pos[4] = rowStart;
    pos[2] = rowStart + (ex[4] ? 1 : 0);
    pos[0] = pos[2] + (ex[2] ? 1 : 0);
    pos[1] = posDiag + 1;
    pos[3] = pos[1] + (ex[1] ? 1 : 0);
    pos[5] = pos[3] + (ex[3] ? 1 : 0);
    double diag = 0.0;
#pragma unroll
    for (int s = 0; s < 2 * DIM; ++s) {
        const int a = s >> 1;
        const bool hi = (s & 1);
        if (ex[s]) {
            const double tij = hi ? P.tij[a][I] : P.tij[a][I - stride[a]];
            const double deriv = tij * up;
            diag += deriv;
            double off = 0.0;
            off -= deriv;
            P.jac[pos[s]] = off;
        } else {
            int f;
            if (a == 0) f = ci[1] + ny * ci[2];
            else if (a == 1) f = ci[0] + nx * ci[2];
            else f = ci[0] + nx * ci[1];
            const int type = P.bc_type[s] ? P.bc_type[s][f] : DMX_BC_NEUMANN;
            if (type == DMX_BC_DIRICHLET) {
                double area = 1.0;
                if (a != 0) area *= P.width[0][ci[0]];
                if (a != 1 && DIM > 1) area *= P.width[1][ci[1]];
                if (a != 2 && DIM > 2) area *= P.width[2][ci[2]];
                const double ti = perm_axis(P, a, I, P.K[I]) * P.extrusion * (hi ? P.gf_hi[a][ci[a]] : P.gf_lo[a][ci[a]]);
                diag += (area * ti) * up;
            }
        }
    }
    P.jac[posDiag] = diag;
}

// ================================================================================================
// DiffMethod::analytic for the incompressible 2p model (p0-s1, phase 0 wetting): CCLocalAssembler<analytic, implicit>
// (assembly/cclocalassembler.hh:490-600) + TwoPIncompressibleLocalResidual (porousmediumflow/2p/incompressiblelocalresidual.hh:
// 80-101 storage derivatives, :137-234 TPFA flux derivatives, :420-481 Dirichlet faces; Neumann faces contribute nothing).
// One thread per block row, rows written whole, blocks [eq][priVar] with priVars (p_w, S_n); the residual comes from the
// JAC = false instantiation of the tile kernel.  Same operation sequence as the oracle (bit-identical).
// Two kernels: twop_analytic_record_kernel evaluates the material law of every cell ONCE -- pc, the two mobilities and the three
// regularised derivatives -dkrw/dSw, -dkrn/dSw, -dpc/dSw, six pow-heavy evaluations -- into a structure-of-arrays scratch
// (6 doubles per cell); the row kernel then reads the records of its cell and its six neighbours (L2 hits) instead of
// re-evaluating them seven times per row.
// ================================================================================================
struct TwoPState {
    double p[2], mob[2], Sw, K;
    int region;
    double dKrw_dSn, dKrn_dSn, dpc_dSn;
};
__global__ void __launch_bounds__(256) twop_analytic_record_kernel(const AsmParams P, double* __restrict__ rec)
{
    const size_t C = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t n = (size_t)P.n;
    if (C >= n) return;
    const double Sn = P.cur[C * 2 + 1];
    const MaterialLaw& law = P.laws[P.region[C]];
    const double Sw = 1 - Sn;
    rec[0 * n + C] = law_pc(law, Sw);
    rec[1 * n + C] = law_krw(law, Sw) / P.mu[0];
    rec[2 * n + C] = law_krn(law, Sw) / P.mu[1];
    rec[3 * n + C] = -1.0 * law_dkrw_dsw(law, Sw);
    rec[4 * n + C] = -1.0 * law_dkrn_dsw(law, Sw);
    rec[5 * n + C] = -1.0 * law_dpc_dsw(law, Sw);
}
__device__ __forceinline__ TwoPState twop_state_rec(const AsmParams& P, const double* __restrict__ rec, size_t C)
{
    TwoPState s;
    const size_t n = (size_t)P.n;
    const double2 u = reinterpret_cast<const double2*>(P.cur)[C];
    s.region = 0;
    s.K = P.K[C];
    s.Sw = 1 - u.y;
    const double pc = rec[0 * n + C];
    s.p[0] = u.x;
    s.p[1] = u.x + pc;
    s.mob[0] = rec[1 * n + C];
    s.mob[1] = rec[2 * n + C];
    s.dKrw_dSn = rec[3 * n + C];
    s.dKrn_dSn = rec[4 * n + C];
    s.dpc_dSn = rec[5 * n + C];
    return s;
}
template <int DIM>
__global__ void __launch_bounds__(128) twop_analytic_jacobian_kernel(const AsmParams P, const double* __restrict__ rec)
{
    const size_t I = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (I >= (size_t)P.n) return;
    const int nx = P.nc[0], ny = P.nc[1];
    const int ci[3] = {(int)(I % nx), (int)((I / nx) % ny), (int)(I / ((size_t)nx * ny))};
    const size_t stride[3] = {1, (size_t)nx, (size_t)nx * ny};
    bool ex[6];
    int pos[6];
#pragma unroll
    for (int s = 0; s < 6; ++s) ex[s] = (s >> 1) < DIM && ((s & 1) ? ci[s >> 1] + 1 < P.nc[s >> 1] : ci[s >> 1] > 0);
    const int rowStart = P.rowptr[I];
    const int posDiag = rowStart + (ex[4] ? 1 : 0) + (ex[2] ? 1 : 0) + (ex[0] ? 1 : 0);
    pos[4] = rowStart;
    pos[2] = rowStart + (ex[4] ? 1 : 0);
    pos[0] = pos[2] + (ex[2] ? 1 : 0);
    pos[1] = posDiag + 1;
    pos[3] = pos[1] + (ex[1] ? 1 : 0);
    pos[5] = pos[3] + (ex[3] ? 1 : 0);

    const TwoPState sI = twop_state_rec(P, rec, I);
    const double w = P.upwind_weight, extr = P.extrusion;
    const double rho_w = P.rho[0], rho_n = P.rho[1];
    const double rhow_muw = rho_w / P.mu[0], rhon_mun = rho_n / P.mu[1];
    const double dKrw_dSn_inside = sI.dKrw_dSn;
    const double dKrn_dSn_inside = sI.dKrn_dSn;
    const double dpc_dSn_inside = sI.dpc_dSn;
    double wdt[3] = {P.width[0][ci[0]], DIM > 1 ? P.width[1][ci[1]] : 1.0, DIM > 2 ? P.width[2][ci[2]] : 1.0};
    double AII[4] = {0.0, 0.0, 0.0, 0.0};
    if (!P.stationary) {
        double vol = 1.0;
        vol *= wdt[0];
        if (DIM > 1) vol *= wdt[1];
        if (DIM > 2) vol *= wdt[2];
        const double phiI = P.phi[I];
        const double porosity = 1.0 - (1.0 - phiI);
        const double poreVolume = vol * porosity;
        AII[1] -= poreVolume * rho_w / P.dt;
        AII[3] += poreVolume * rho_n / P.dt;
    }
#pragma unroll
    for (int s = 0; s < 2 * DIM; ++s) {
        const int a = s >> 1;
        const bool hi = (s & 1);
        double area = 1.0;
        if (a != 0) area *= wdt[0];
        if (a != 1 && DIM > 1) area *= wdt[1];
        if (a != 2 && DIM > 2) area *= wdt[2];
        const bool grav = P.enable_gravity && (a == DIM - 1);
        const double ng = hi ? -P.gravity : P.gravity;
        const double alphaI = sI.K * ng * extr;
        double pJ[2], upJ[2], tij, flux[2];
        double dKrw_dSn_outside = 0.0, dKrn_dSn_outside = 0.0, dpc_dSn_outside = 0.0;
        bool boundary = false;
        if (ex[s]) {
            const size_t J = hi ? I + stride[a] : I - stride[a];
            const int cj = hi ? ci[a] + 1 : ci[a] - 1;
            const TwoPState sJ = twop_state_rec(P, rec, J);
            tij = hi ? P.tij[a][I] : P.tij[a][J];
            pJ[0] = sJ.p[0]; pJ[1] = sJ.p[1];
            upJ[0] = rho_w * sJ.mob[0]; upJ[1] = rho_n * sJ.mob[1];
            const double tJ = sJ.K * extr * (hi ? P.gf_lo[a][cj] : P.gf_hi[a][cj]);
            const double alphaJ = sJ.K * ng * extr;
#pragma unroll
            for (int ph = 0; ph < 2; ++ph) {
                double f = tij * (sI.p[ph] - pJ[ph]);
                if (grav) {
                    const double rho = (P.rho[ph] + P.rho[ph]) * 0.5;
                    f = f + rho * area * alphaI;
                    f -= rho * tij / tJ * (alphaI - alphaJ);
                }
                flux[ph] = f;
            }
            dKrw_dSn_outside = sJ.dKrw_dSn;
            dKrn_dSn_outside = sJ.dKrn_dSn;
            dpc_dSn_outside = sJ.dpc_dSn;
        } else {
            int f_;
            if (a == 0) f_ = ci[1] + ny * ci[2];
            else if (a == 1) f_ = ci[0] + nx * ci[2];
            else f_ = ci[0] + nx * ci[1];
            const int type = P.bc_type[s] ? P.bc_type[s][f_] : DMX_BC_NEUMANN;
            if (type != DMX_BC_DIRICHLET) continue;
            boundary = true;
            const double ti = perm_axis(P, a, I, sI.K) * extr * (hi ? P.gf_hi[a][ci[a]] : P.gf_lo[a][ci[a]]);
            tij = area * ti;
#pragma unroll
            for (int ph = 0; ph < 2; ++ph) {
                pJ[ph] = P.bc_p[s][(size_t)f_ * 2 + ph];
                upJ[ph] = P.bc_up[s][(size_t)f_ * 2 + ph];
                double f = tij * (sI.p[ph] - pJ[ph]);
                if (grav) f = f + P.bc_rho[s][(size_t)f_ * 2 + ph] * area * alphaI;
                flux[ph] = f;
            }
        }
        const double flux_w = flux[0], flux_n = flux[1];
        const double insideWeight_w = signbit(flux_w) ? (1.0 - w) : w;
        const double outsideWeight_w = 1.0 - insideWeight_w;
        const double insideWeight_n = signbit(flux_n) ? (1.0 - w) : w;
        const double outsideWeight_n = 1.0 - insideWeight_n;
        const double up_w = (rho_w * sI.mob[0]) * insideWeight_w + upJ[0] * outsideWeight_w;
        const double up_n = (rho_n * sI.mob[1]) * insideWeight_n + upJ[1] * outsideWeight_n;
        if (boundary) {
            AII[0] += tij * up_w;
            AII[1] += rhow_muw * flux_w * dKrw_dSn_inside * insideWeight_w;
            AII[2] += tij * up_n;
            AII[3] += rhon_mun * flux_n * dKrn_dSn_inside * insideWeight_n;
            AII[3] += tij * dpc_dSn_inside * up_n;
            continue;
        }
        const double rho_mu_flux_w = rhow_muw * flux_w, rho_mu_flux_n = rhon_mun * flux_n;
        const double tij_up_w = tij * up_w, tij_up_n = tij * up_n;
        double AIJ[4] = {0.0, 0.0, 0.0, 0.0};
        AII[0] += tij_up_w;
        AIJ[0] -= tij_up_w;
        AII[1] += rho_mu_flux_w * dKrw_dSn_inside * insideWeight_w;
        AIJ[1] += rho_mu_flux_w * dKrw_dSn_outside * outsideWeight_w;
        AII[2] += tij_up_n;
        AIJ[2] -= tij_up_n;
        AII[3] += rho_mu_flux_n * dKrn_dSn_inside * insideWeight_n;
        AIJ[3] += rho_mu_flux_n * dKrn_dSn_outside * outsideWeight_n;
        AII[3] += tij_up_n * dpc_dSn_inside;
        AIJ[3] -= tij_up_n * dpc_dSn_outside;
        double* dst = P.jac + (size_t)pos[s] * 4;
        dst[0] = AIJ[0]; dst[1] = AIJ[1]; dst[2] = AIJ[2]; dst[3] = AIJ[3];
    }
    double* dst = P.jac + (size_t)posDiag * 4;
    dst[0] = AII[0]; dst[1] = AII[1]; dst[2] = AII[2]; dst[3] = AII[3];
}

// ================================================================================================
// Tracer transport on a frozen velocity field (BASELINE config 5, examples/1ptracer)
// ================================================================================================
// Volume fluxes over all scvfs from the 1p pressure field in CUR: examples/1ptracer/main.cc:162-199
// (fluxVars.advectiveFlux(0, upwindTerm = mobility), flux/cctpfa/darcyslaw.hh:154-213); Neumann faces stay 0.
template <bool TABLE, int DIM>
__global__ void __launch_bounds__(256) volume_flux_kernel(const AsmParams P, double* __restrict__ out)
{
    const size_t I = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (I >= (size_t)P.n) return;
    const int nx = P.nc[0], ny = P.nc[1];
    const int ci[3] = {(int)(I % nx), (int)((I / nx) % ny), (int)(I / ((size_t)nx * ny))};
    const size_t stride[3] = {1, (size_t)nx, (size_t)nx * ny};
    const double w = P.upwind_weight, extr = P.extrusion;
    auto fluid = [&](double p, double* rho, double* mob) {
        double mu = P.mu[0];
        *rho = P.rho[0];
        if constexpr (TABLE) table_interp2(P.table, p, rho, &mu);
        *mob = 1.0 / mu;                                            // 1p/volumevariables.hh:178
    };
    const double pI = P.cur[I], KI = P.K[I];
    double rhoI, mobI;
    fluid(pI, &rhoI, &mobI);
#pragma unroll
    for (int s = 0; s < 2 * DIM; ++s) {
        const int a = s >> 1;
        const bool hi = (s & 1);
        double area = 1.0;
        for (int d = 0; d < DIM; ++d)
            if (d != a) area *= P.width[d][ci[d]];
        const bool grav = P.enable_gravity && (a == DIM - 1);
        const double ng = hi ? -P.gravity : P.gravity;
        const double alphaI = KI * ng * extr;
        double result = 0.0;
        const bool exists = hi ? (ci[a] + 1 < P.nc[a]) : (ci[a] > 0);
        if (exists) {
            const size_t J = hi ? I + stride[a] : I - stride[a];
            const int cj = hi ? ci[a] + 1 : ci[a] - 1;
            const double pJ = P.cur[J], KJ = P.K[J];
            double rhoJ, mobJ;
            fluid(pJ, &rhoJ, &mobJ);
            const double tij = hi ? P.tij[a][I] : P.tij[a][J];
            double f = tij * (pI - pJ);
            if (grav) {
                const double rho = (rhoI + rhoJ) * 0.5;
                f = f + rho * area * alphaI;
                const double tJ = KJ * extr * (hi ? P.gf_lo[a][cj] : P.gf_hi[a][cj]);
                const double alphaJ = KJ * ng * extr;
                f -= rho * tij / tJ * (alphaI - alphaJ);
            }
            double mult;
            if (signbit(f)) mult = w * mobJ + (1.0 - w) * mobI;
            else mult = w * mobI + (1.0 - w) * mobJ;
            result = f * mult;
        } else {
            int fidx;
            if (a == 0) fidx = ci[1] + ny * ci[2];
            else if (a == 1) fidx = ci[0] + nx * ci[2];
            else fidx = ci[0] + nx * ci[1];
            const int type = P.bc_type[s] ? P.bc_type[s][fidx] : DMX_BC_NEUMANN;
            if (type == DMX_BC_DIRICHLET) {
                const double pD = P.bc_p[s][(size_t)fidx * 2];
                double rhoD, mobD;
                fluid(pD, &rhoD, &mobD);
                const double ti = perm_axis(P, a, I, KI) * extr * (hi ? P.gf_hi[a][ci[a]] : P.gf_lo[a][ci[a]]);
                const double tij = area * ti;
                double f = tij * (pI - pD);
                if (grav) f = f + rhoD * area * alphaI;
                double mult;
                if (signbit(f)) mult = w * mobD + (1.0 - w) * mobI;
                else mult = w * mobI + (1.0 - w) * mobD;
                result = f * mult;
            }
        }
        out[I * (2 * DIM) + s] = result;
    }
}

// TracerLocalResidual (porousmediumflow/tracer/localresidual.hh:74-107 storage, :119-186 advective flux over
// StationaryVelocityField, flux/stationaryvelocityfield.hh:55; mass fractions, one component, D = 0) assembled as
//   explicit: CCLocalAssembler<analytic, implicit=false> (assembly/cclocalassembler.hh:607-675): fluxes at PREV, the
//             Jacobian is the storage derivative on the diagonal (localresidual.hh:193-214)
//   implicit: CCLocalAssembler<analytic, implicit=true> (:490-600): fluxes at CUR, addFluxDerivatives (:237-296); the
//             outflow Neumann term contributes volumeFlux*rho*extrusion to the diagonal (the reference has no analytic
//             Robin derivative, fvlocalresidual.hh:455-464).
// One thread per row; every entry of the row is written (zeros where the explicit scheme has none).
template <int DIM, bool JAC>
__global__ void __launch_bounds__(256) tracer_assemble_kernel(const AsmParams P)
{
    const size_t I = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (I >= (size_t)P.n) return;
    const int nx = P.nc[0], ny = P.nc[1];
    const int ci[3] = {(int)(I % nx), (int)((I / nx) % ny), (int)(I / ((size_t)nx * ny))};
    const size_t stride[3] = {1, (size_t)nx, (size_t)nx * ny};
    const double w = P.upwind_weight, extr = P.extrusion, rho = P.rho[0];
    const bool implicit = P.tracer_implicit != 0;
    const double* __restrict__ X = implicit ? P.cur : P.prev;
    const double XI = X[I];

    double vol = 1.0;
    for (int a = 0; a < DIM; ++a) vol *= P.width[a][ci[a]];
    double res = 0.0;
    {
        double source = P.q ? P.q[I] : 0.0;
        source *= vol * extr;
        res -= source;
    }
    const double phiInert = 1.0 - P.phi[I];
    const double porosity = 1.0 - phiInert;
    const double saturation = fmax(1e-8, 1.0);

    bool ex[6] = {false, false, false, false, false, false};
#pragma unroll
    for (int s = 0; s < 2 * DIM; ++s) ex[s] = (s & 1) ? (ci[s >> 1] + 1 < P.nc[s >> 1]) : (ci[s >> 1] > 0);
    int pos[6], posDiag = 0;
    if constexpr (JAC) {
        const int rowStart = P.rowptr[I];
        posDiag = rowStart + (ex[4] ? 1 : 0) + (ex[2] ? 1 : 0) + (ex[0] ? 1 : 0);
        pos[4] = rowStart;
        pos[2] = rowStart + (ex[4] ? 1 : 0);
        pos[0] = pos[2] + (ex[2] ? 1 : 0);
        pos[1] = posDiag + 1;
        pos[3] = pos[1] + (ex[1] ? 1 : 0);
        pos[5] = pos[3] + (ex[3] ? 1 : 0);
    }
    double diag = 0.0;
    if constexpr (JAC) {
        const double d_storage = vol * porosity * rho * saturation / P.dt;
        diag += d_storage;
    }
#pragma unroll
    for (int s = 0; s < 2 * DIM; ++s) {
        const int a = s >> 1;
        const bool hi = (s & 1);
        const double vflux = P.vf[I * (2 * DIM) + s];
        if (ex[s]) {
            const size_t J = hi ? I + stride[a] : I - stride[a];
            const double upIn = rho * XI, upOut = rho * X[J];
            double mult;
            if (signbit(vflux)) mult = w * upOut + (1.0 - w) * upIn;
            else mult = w * upIn + (1.0 - w) * upOut;
            // Fick's law, TPFA (flux/cctpfa/fickslaw.hh) with D_eff = porosity * S * tau * D (S = 1): harmonic transmissibility of
            // the two half-cell values, flux = rho_avg * tij * (X_I - X_J); D = 0 gives tij = 0
            const int cj = hi ? ci[a] + 1 : ci[a] - 1;
            const double porosityJ = 1.0 - (1.0 - P.phi[J]);
            const double DeffI = porosity * 1.0 * P.tracer_tau * P.tracer_D;
            const double DeffJ = porosityJ * 1.0 * P.tracer_tau * P.tracer_D;
            double areaF = 1.0;
            for (int d = 0; d < DIM; ++d)
                if (d != a) areaF *= P.width[d][ci[d]];
            const double ti = DeffI * extr * (hi ? P.gf_hi[a][ci[a]] : P.gf_lo[a][ci[a]]);
            const double tj = DeffJ * extr * (hi ? P.gf_lo[a][cj] : P.gf_hi[a][cj]);
            double dTij;
            if (ti * tj <= 0.0) dTij = 0;
            else dTij = areaF * (ti * tj) / (ti + tj);
            const double rhoAvg = 0.5 * (rho + rho);
            double flux = 0.0;
            flux += vflux * mult;
            flux += rhoAvg * dTij * (XI - X[J]);
            if (P.disp) {
                // mechanical dispersion (flux/cctpfa/dispersionflux.hh:93-104,172-213): the tensor is given at the face, D_i = D_j, and
                // only its normal entry enters the TPFA transmissibility; no derivative in the Jacobian, as in the reference's
                // TracerLocalResidual::addFluxDerivatives (tracer/localresidual.hh:237-291)
                const double Dd = P.disp[I * (2 * DIM) + s];
                const double mi = Dd * extr * (hi ? P.gf_hi[a][ci[a]] : P.gf_lo[a][ci[a]]);
                const double mj = Dd * extr * (hi ? P.gf_lo[a][cj] : P.gf_hi[a][cj]);
                double mTij;
                if (mi * mj <= 0.0) mTij = 0;
                else mTij = areaF * (mi * mj) / (mi + mj);
                flux += rhoAvg * mTij * (XI - X[J]);
            }
            res += flux;
            if constexpr (JAC) {
                double offdiag = 0.0;
                if (implicit) {
                    const double insideWeight = signbit(vflux) ? (1.0 - w) : w;
                    const double outsideWeight = 1.0 - insideWeight;
                    const double diffDeriv = rhoAvg * dTij;
                    diag += (vflux * rho * insideWeight + diffDeriv);
                    offdiag += (vflux * rho * outsideWeight - diffDeriv);
                }
                P.jac[pos[s]] = offdiag;
            }
        } else {
            int fidx;
            if (a == 0) fidx = ci[1] + ny * ci[2];
            else if (a == 1) fidx = ci[0] + nx * ci[2];
            else fidx = ci[0] + nx * ci[1];
            const int type = P.bc_type[s] ? P.bc_type[s][fidx] : DMX_BC_NEUMANN;
            if (type == DMX_BC_NONE) continue;
            double area = 1.0;
            for (int d = 0; d < DIM; ++d)
                if (d != a) area *= P.width[d][ci[d]];
            double nf;
            if (type == DMX_BC_OUTFLOW) {
                nf = vflux * XI * rho / area;
                if (JAC && implicit) diag += vflux * rho * extr;
            } else
                nf = P.bc_neumann[s] ? P.bc_neumann[s][fidx] : 0.0;
            nf *= area * extr;
            res += nf;
        }
    }
    {
        // fvlocalresidual.hh:274-304
        double prevStorage = porosity * rho * P.prev[I] * saturation;
        double storage = porosity * rho * P.cur[I] * saturation;
        prevStorage *= extr;
        storage *= extr;
        storage -= prevStorage;
        storage *= vol;
        storage /= P.dt;
        double st = 0.0;
        st += storage;
        res += st;
    }
    P.residual[I] = res;
    if (!(fabs(res) <= DBL_MAX)) atomicOr(P.flag_nonfinite, 1);
    if constexpr (JAC) P.jac[posDiag] = 0.0 + diag;
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static void fill_params(dmx_ctx* ctx, AsmParams& P)
{
    P.model = ctx->model; P.b = ctx->b; P.dim = ctx->dim;
    for (int a = 0; a < 3; ++a) {
        P.nc[a] = ctx->nc[a];
        P.width[a] = ctx->d_width[a]; P.gf_lo[a] = ctx->d_gflo[a]; P.gf_hi[a] = ctx->d_gfhi[a];
        P.tij[a] = ctx->d_tij[a];
    }
    P.n = ctx->n;
    const dmx_options& o = ctx->opt;
    P.enable_gravity = o.enable_gravity; P.fd_method = o.fd_method; P.stationary = o.stationary;
    P.nd = (o.fd_method == 5) ? 4 : (o.fd_method == 0 ? 2 : 1);
    P.gravity = o.gravity; P.upwind_weight = o.upwind_weight; P.base_eps = o.base_eps;
    P.mag[0] = o.privar_magnitude[0]; P.mag[1] = o.privar_magnitude[1];
    P.dt = o.dt; P.extrusion = o.extrusion;
    for (int a = 0; a < 3; ++a) P.Kaxis[a] = ctx->d_Kaxis[a];
    P.K = ctx->d_K; P.phi = ctx->d_phi; P.region = ctx->d_region; P.q = ctx->d_q;
    for (int i = 0; i < 2; ++i) { P.rho[i] = ctx->rho[i]; P.mu[i] = ctx->mu[i]; P.rmu[i] = 1.0 / ctx->mu[i]; }
    P.rdt = 1.0 / o.dt;
    P.nlaws = (int)ctx->laws.size();
    {
        // layers per CTA of the tile kernel: about ten waves of two CTAs per SM (measured at 256^3: 2.53 ms with 24 layers, 2.68 with
        // 37, 3.0 with 128 -- the tail of the last wave costs more than the 2/zchunk halo layers), at least 16 layers
        const int tiles = ((ctx->nc[0] + AT_TX - 1) / AT_TX) * ((ctx->nc[1] + AT_TY - 1) / AT_TY);
        const int want = std::max(1, (20 * ctx->num_sms + tiles - 1) / tiles);
        int zc = (ctx->nc[2] + want - 1) / want;
        zc = std::max(zc, std::min(16, ctx->nc[2]));
        static const int zc_env = [] { const char* e = getenv("DMX_ZCHUNK"); return e ? atoi(e) : 0; }();   // tuning override
        P.zchunk = zc_env > 0 ? std::min(zc_env, ctx->nc[2]) : zc;
    }
    P.tabulated = ctx->tabulated ? 1 : 0;
    P.table = ctx->d_table;
    P.laws = ctx->d_laws;
    for (int s = 0; s < 6; ++s) {
        P.bc_type[s] = ctx->d_bc_type[s]; P.bc_neumann[s] = ctx->d_bc_neumann[s];
        P.bc_p[s] = ctx->d_bc_p[s]; P.bc_up[s] = ctx->d_bc_up[s]; P.bc_rho[s] = ctx->d_bc_rho[s];
    }
    P.cur = ctx->d_vec[DMX_VEC_CUR]; P.prev = ctx->d_vec[DMX_VEC_PREV];
    P.vf = ctx->d_vf; P.disp = ctx->d_disp; P.tracer_implicit = ctx->tracer_implicit;
    P.tracer_D = ctx->tracer_D; P.tracer_tau = ctx->tracer_tau;
    P.rowptr = ctx->d_rowptr; P.residual = ctx->d_vec[DMX_VEC_RESIDUAL]; P.jac = ctx->d_J;
    P.flag_nonfinite = ctx->d_flag;
}

// host evaluation of the Dirichlet "outside" volume variables with the INSIDE cell's spatial parameters
// (discretization/cellcentered/tpfa/elementvolumevariables.hh:318-346)
static void dirichlet_state(const dmx_ctx* ctx, int cell, const double* pv, double* p, double* up, double* rho)
{
    if (ctx->model == DMX_MODEL_2P) {
        const MaterialLaw& law = ctx->laws[ctx->h_region[cell]];
        const double Sn = pv[1];
        const int w = law.wetting ? 1 : 0, nw = 1 - w;
        const double sw = w ? Sn : 1 - Sn;                     // saturation of the wetting phase
        const double pc = law_pc(law, sw);
        p[0] = pv[0];
        p[1] = w ? pv[0] - pc : pv[0] + pc;
        up[w] = ctx->rho[w] * (law_krw(law, sw) / ctx->mu[w]);
        up[nw] = ctx->rho[nw] * (law_krn(law, sw) / ctx->mu[nw]);
        rho[0] = ctx->rho[0];
        rho[1] = ctx->rho[1];
    } else {
        p[0] = pv[0]; p[1] = 0.0;
        double r = ctx->rho[0], m = ctx->mu[0];
        if (ctx->tabulated) {
            r = table_interp(ctx->h_table, ctx->h_table.rho, pv[0]);
            m = table_interp(ctx->h_table, ctx->h_table.mu, pv[0]);
        }
        rho[0] = r; rho[1] = 0.0;
        up[0] = r * (1.0 / m); up[1] = 0.0;
    }
}

static int side_faces(const dmx_ctx* ctx, int side)
{
    const int a = side / 2;
    int nf = 1;
    for (int d = 0; d < 3; ++d)
        if (d != a) nf *= ctx->nc[d];
    return nf;
}
static int side_face_cell(const dmx_ctx* ctx, int side, int f)
{
    const int a = side / 2;
    int c[3] = {0, 0, 0};
    if (a == 0) { c[1] = f % ctx->nc[1]; c[2] = f / ctx->nc[1]; }
    else if (a == 1) { c[0] = f % ctx->nc[0]; c[2] = f / ctx->nc[0]; }
    else { c[0] = f % ctx->nc[0]; c[1] = f / ctx->nc[0]; }
    c[a] = (side & 1) ? ctx->nc[a] - 1 : 0;
    return c[0] + ctx->nc[0] * (c[1] + ctx->nc[1] * c[2]);
}

template <class T>
static int upload(dmx_ctx* ctx, T** dptr, const std::vector<T>& h)
{
    if (*dptr) { cudaFree(*dptr); *dptr = nullptr; }
    if (h.empty()) return 0;
    DMX_CUDA(cudaMalloc((void**)dptr, h.size() * sizeof(T)));
    DMX_CUDA(cudaMemcpyAsync(*dptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}

// One-time (per parameter change) device set-up: laws, boundary states, transmissibilities, record buffer.
int prepare(dmx_ctx* ctx)
{
    if (ctx->prepared) return 0;
    if (!ctx->has_grid) return fail(ctx, DMX_ERR_USAGE, "prepare: no grid set");
    DMX_CUDA(cudaSetDevice(ctx->device));
    // material laws
    if (ctx->model == DMX_MODEL_2P && ctx->laws.empty()) return fail(ctx, DMX_ERR_USAGE, "2p model without material law");
    if (ctx->model == DMX_MODEL_2P) {
        // every region id a cell refers to needs its dmx_set_material call (spatialParams.fluidMatrixInteraction of that cell)
        int maxRegion = 0;
        for (int r : ctx->h_region) maxRegion = std::max(maxRegion, r);
        if (maxRegion >= (int)ctx->laws.size())
            return fail(ctx, DMX_ERR_USAGE, "a cell refers to material region " + std::to_string(maxRegion) + " but only " +
                                                std::to_string(ctx->laws.size()) + " regions were given to dmx_set_material");
        std::vector<char> used(ctx->laws.size(), 0);
        for (int r : ctx->h_region) used[r] = 1;
        for (size_t r = 0; r < ctx->laws.size(); ++r)
            if (used[r] && ctx->laws[r].kind == DMX_LAW_BROOKSCOREY && ctx->laws[r].lambda == 0.0)      // value-initialised slot
                return fail(ctx, DMX_ERR_USAGE, "material region " + std::to_string(r) + " is used by cells but was never set");
    }
    for (auto& l : ctx->laws) law_init(l);
    if (int rc = upload(ctx, &ctx->d_laws, ctx->laws)) return rc;
    // boundary data
    for (int s = 0; s < 2 * ctx->dim; ++s) {
        const int nf = side_faces(ctx, s);
        std::vector<int> type(nf, DMX_BC_NEUMANN);
        std::vector<double> neu((size_t)nf * ctx->b, 0.0), p((size_t)nf * 2, 0.0), up((size_t)nf * 2, 0.0), rho((size_t)nf * 2, 0.0);
        if (!ctx->h_bc_type[s].empty()) {
            type = ctx->h_bc_type[s];
            for (int f = 0; f < nf; ++f) {
                const double* v = &ctx->h_bc_val[s][(size_t)f * ctx->b];
                if (type[f] == DMX_BC_DIRICHLET && ctx->model == DMX_MODEL_TRACER)
                    return fail(ctx, DMX_ERR_USAGE, "tracer: Dirichlet boundaries are not supported (use DMX_BC_OUTFLOW / Neumann)");
                if (type[f] == DMX_BC_DIRICHLET) {
                    double pv[2] = {v[0], ctx->b > 1 ? v[1] : 0.0};
                    dirichlet_state(ctx, side_face_cell(ctx, s, f), pv, &p[(size_t)f * 2], &up[(size_t)f * 2], &rho[(size_t)f * 2]);
                } else if (type[f] == DMX_BC_NEUMANN)
                    for (int e = 0; e < ctx->b; ++e) neu[(size_t)f * ctx->b + e] = v[e];
            }
        }
        if (int rc = upload(ctx, &ctx->d_bc_type[s], type)) return rc;
        if (int rc = upload(ctx, &ctx->d_bc_neumann[s], neu)) return rc;
        if (int rc = upload(ctx, &ctx->d_bc_p[s], p)) return rc;
        if (int rc = upload(ctx, &ctx->d_bc_up[s], up)) return rc;
        if (int rc = upload(ctx, &ctx->d_bc_rho[s], rho)) return rc;
    }
    // transmissibilities
    for (int a = 0; a < 3; ++a)
        if (!ctx->d_tij[a]) DMX_CUDA(cudaMalloc((void**)&ctx->d_tij[a], (size_t)ctx->n * sizeof(double)));
    AsmParams P;
    fill_params(ctx, P);
    const int threads = 256;
    transmissibility_kernel<<<(ctx->n + threads - 1) / threads, threads, 0, ctx->stream>>>(P, ctx->d_tij[0], ctx->d_tij[1], ctx->d_tij[2]);
    DMX_CHECK_LAUNCH();
    ctx->prepared = true;
    return 0;
}

template <int MODEL, bool TABLE, int ND, bool JAC, int DIM>
static int launch_tile_dim(dmx_ctx* ctx, const AsmParams& P)
{
    using L = RecLayout<MODEL, TABLE, JAC ? ND : 0>;
    const size_t smem = 3 * (size_t)L::NF * AT_HC * sizeof(double) + (MODEL == DMX_MODEL_2P ? DMX_MAX_REGIONS * sizeof(MaterialLaw) : 0);
    auto kern = assemble_tile_kernel<MODEL, TABLE, ND, JAC, DIM>;
    DMX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int ntx = (ctx->nc[0] + AT_TX - 1) / AT_TX, nty = (ctx->nc[1] + AT_TY - 1) / AT_TY;
    const dim3 grid((unsigned)(ntx * nty), (unsigned)((ctx->nc[2] + P.zchunk - 1) / P.zchunk));
    ProfScope ps__(ctx, DMX_K_ASSEMBLY);
    kern<<<grid, AT_THREADS, smem, ctx->stream>>>(P);
    DMX_CHECK_LAUNCH();
    return 0;
}

template <int MODEL, bool TABLE, int ND, bool JAC>
static int launch_tile_inst(dmx_ctx* ctx, const AsmParams& P)
{
    if (P.dim == 3) return launch_tile_dim<MODEL, TABLE, ND, JAC, 3>(ctx, P);
    if (P.dim == 2) return launch_tile_dim<MODEL, TABLE, ND, JAC, 2>(ctx, P);
    return launch_tile_dim<MODEL, TABLE, ND, JAC, 1>(ctx, P);
}

template <int MODEL, bool TABLE>
static int launch_tile(dmx_ctx* ctx, const AsmParams& P, bool with_jac)
{
    if (P.nlaws > DMX_MAX_REGIONS) return fail(ctx, DMX_ERR_USAGE, "too many material-law regions");
    if (!with_jac) return launch_tile_inst<MODEL, TABLE, 1, false>(ctx, P);
    if (P.nd == 1) return launch_tile_inst<MODEL, TABLE, 1, true>(ctx, P);
    if (P.nd == 2) return launch_tile_inst<MODEL, TABLE, 2, true>(ctx, P);
    return launch_tile_inst<MODEL, TABLE, 4, true>(ctx, P);
}

static int launch_impl(dmx_ctx* ctx, bool with_jac, bool volvars_only)
{
    if (int rc = prepare(ctx)) return rc;
    if (!ctx->opt.stationary && ctx->opt.dt <= 0.0) return fail(ctx, DMX_ERR_USAGE, "assemble: dt must be > 0");
    AsmParams P;
    fill_params(ctx, P);
    (void)volvars_only;
    if (ctx->model == DMX_MODEL_TRACER) {
        if (!ctx->d_vf) return fail(ctx, DMX_ERR_USAGE, "tracer: no volume fluxes set (dmx_set_volume_flux)");
        if (ctx->opt.stationary) return fail(ctx, DMX_ERR_USAGE, "tracer: instationary only");
        const unsigned grid = (unsigned)((ctx->n + 255) / 256);
        ProfScope ps__(ctx, DMX_K_ASSEMBLY);
#define DMX_TRACER(D)                                                                                         \
        do {                                                                                                  \
            if (with_jac) tracer_assemble_kernel<D, true><<<grid, 256, 0, ctx->stream>>>(P);                  \
            else tracer_assemble_kernel<D, false><<<grid, 256, 0, ctx->stream>>>(P);                          \
        } while (0)
        if (ctx->dim == 3) DMX_TRACER(3);
        else if (ctx->dim == 2) DMX_TRACER(2);
        else DMX_TRACER(1);
#undef DMX_TRACER
        DMX_CHECK_LAUNCH();
        return 0;
    }
    if (ctx->opt.fd_method == DMX_DIFF_ANALYTIC) {
        if (ctx->tabulated) return fail(ctx, DMX_ERR_USAGE, "DiffMethod::analytic is available for incompressible fluids only");
        const unsigned grid = (unsigned)((ctx->n + 127) / 128);
        if (ctx->model == DMX_MODEL_2P) {
            for (const auto& l : ctx->laws)
                if (l.wetting != 0)
                    return fail(ctx, DMX_ERR_USAGE, "DiffMethod::analytic (2p/incompressiblelocalresidual.hh) assumes that phase 0 is the wetting phase");
            if (int rc = launch_tile<DMX_MODEL_2P, false>(ctx, P, false)) return rc;      // residual
            if (!with_jac) return 0;
            ProfScope ps__(ctx, DMX_K_ASSEMBLY);
            if (!ctx->d_law_rec) DMX_CUDA(cudaMalloc((void**)&ctx->d_law_rec, (size_t)ctx->n * 6 * sizeof(double)));
            twop_analytic_record_kernel<<<(unsigned)((ctx->n + 255) / 256), 256, 0, ctx->stream>>>(P, ctx->d_law_rec);
            DMX_CHECK_LAUNCH();
            if (ctx->dim == 3) twop_analytic_jacobian_kernel<3><<<grid, 128, 0, ctx->stream>>>(P, ctx->d_law_rec);
            else if (ctx->dim == 2) twop_analytic_jacobian_kernel<2><<<grid, 128, 0, ctx->stream>>>(P, ctx->d_law_rec);
            else twop_analytic_jacobian_kernel<1><<<grid, 128, 0, ctx->stream>>>(P, ctx->d_law_rec);
            DMX_CHECK_LAUNCH();
            return 0;
        }
        if (int rc = launch_tile<DMX_MODEL_1P, false>(ctx, P, false)) return rc;          // residual
        if (!with_jac) return 0;
        const double up = ctx->rho[0] / ctx->mu[0];               // volVars.density() / volVars.viscosity()
        ProfScope ps__(ctx, DMX_K_ASSEMBLY);
        if (ctx->dim == 3) onep_analytic_jacobian_kernel<3><<<grid, 128, 0, ctx->stream>>>(P, up);
        else if (ctx->dim == 2) onep_analytic_jacobian_kernel<2><<<grid, 128, 0, ctx->stream>>>(P, up);
        else onep_analytic_jacobian_kernel<1><<<grid, 128, 0, ctx->stream>>>(P, up);
        DMX_CHECK_LAUNCH();
        return 0;
    }
    if (ctx->model == DMX_MODEL_2P) return launch_tile<DMX_MODEL_2P, false>(ctx, P, with_jac);
    if (ctx->tabulated) return launch_tile<DMX_MODEL_1P, true>(ctx, P, with_jac);
    return launch_tile<DMX_MODEL_1P, false>(ctx, P, with_jac);
}

int launch_volume_flux(dmx_ctx* ctx, double* d_out)
{
    if (int rc = prepare(ctx)) return rc;
    AsmParams P;
    fill_params(ctx, P);
    const unsigned grid = (unsigned)((ctx->n + 255) / 256);
#define DMX_VF(T, D) volume_flux_kernel<T, D><<<grid, 256, 0, ctx->stream>>>(P, d_out)
    if (ctx->tabulated) { if (ctx->dim == 3) DMX_VF(true, 3); else if (ctx->dim == 2) DMX_VF(true, 2); else DMX_VF(true, 1); }
    else { if (ctx->dim == 3) DMX_VF(false, 3); else if (ctx->dim == 2) DMX_VF(false, 2); else DMX_VF(false, 1); }
#undef DMX_VF
    DMX_CHECK_LAUNCH();
    return 0;
}

// ================================================================================================
// Output fields of the device-resident solution (what VtkOutputModule asks the volume variables for):
// TwoPIOFields (porousmediumflow/2p/iofields.hh:31-50): per phase S, p, rho, mobility, then pc and porosity;
// OnePIOFields (1p/iofields.hh:30-33): p.  out[field][cell] (SoA), computed from CUR with the same secondary-variable
// evaluation as the assembly (TwoPVolumeVariables::update incl. the per-region wetting phase).
// ================================================================================================
__global__ void __launch_bounds__(256) output_fields_2p_kernel(const AsmParams P, double* __restrict__ out)
{
    const size_t I = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t n = (size_t)P.n;
    if (I >= n) return;
    const double2 u = reinterpret_cast<const double2*>(P.cur)[I];
    const MaterialLaw& law = P.laws[P.region[I]];
    const int w = law.wetting ? 1 : 0, nw = 1 - w;
    double S[2], p[2], mob[2];
    S[1] = u.y;
    S[0] = 1 - u.y;
    const double pc = law_pc(law, S[w]);
    p[0] = u.x;
    p[1] = w ? u.x - pc : u.x + pc;
    mob[w] = law_krw(law, S[w]) / P.mu[w];
    mob[nw] = law_krn(law, S[w]) / P.mu[nw];
#pragma unroll
    for (int ph = 0; ph < 2; ++ph) {
        out[(size_t)(4 * ph + 0) * n + I] = S[ph];
        out[(size_t)(4 * ph + 1) * n + I] = p[ph];
        out[(size_t)(4 * ph + 2) * n + I] = P.rho[ph];
        out[(size_t)(4 * ph + 3) * n + I] = mob[ph];
    }
    out[(size_t)8 * n + I] = pc;
    out[(size_t)9 * n + I] = 1.0 - (1.0 - P.phi[I]);
}

int launch_output_fields(dmx_ctx* ctx, double* d_out)
{
    if (int rc = prepare(ctx)) return rc;
    AsmParams P;
    fill_params(ctx, P);
    if (ctx->model == DMX_MODEL_2P) {
        output_fields_2p_kernel<<<(unsigned)((ctx->n + 255) / 256), 256, 0, ctx->stream>>>(P, d_out);
        DMX_CHECK_LAUNCH();
        return 0;
    }
    DMX_CUDA(cudaMemcpyAsync(d_out, ctx->d_vec[DMX_VEC_CUR], (size_t)ctx->n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    return 0;
}

int launch_assemble(dmx_ctx* ctx, bool with_jacobian)
{
    const int rc = launch_impl(ctx, with_jacobian, false);
    // explicit tracer steps assemble the storage derivative only (tracer/localresidual.hh:193-214): off-diagonal blocks are +0.0
    if (with_jacobian) ctx->jac_diagonal = (rc == 0 && ctx->model == DMX_MODEL_TRACER && !ctx->tracer_implicit);
    return rc;
}

} // namespace dmx
