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
#include <cfloat>

#include "common.cuh"

namespace dmx {

// ------------------------------------------------------------------------------------------------
// setup: transmissibilities of the +faces.  tij = A * ti*tj/(ti+tj) (0 if ti*tj <= 0)
// ------------------------------------------------------------------------------------------------
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
            const double ti = KI * P.extrusion * P.gf_hi[a][c[a]];
            const double tj = P.K[I + stride[a]] * P.extrusion * P.gf_lo[a][c[a] + 1];
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
// f0: undeflected value, f[k]: deflected values
template <int ND>
__device__ __forceinline__ double fd_quotient(int method, double f0, const double* f, double eps)
{
    if (ND == 1) {
        double d, delta = 0.0;
        if (method >= 0) { delta += eps; d = f[0]; d -= f0; }
        else { delta += eps; d = f0; d -= f[0]; }
        return d / delta;
    } else if (ND == 2) {
        double delta = 0.0;
        delta += eps;
        double d = f[0];
        delta += eps;
        d -= f[1];
        return d / delta;
    } else {
        double d = f[0];
        d -= f[1];
        d *= 8.0;
        d += f[2];
        d -= f[3];
        return d / (12.0 * eps);
    }
}

// ------------------------------------------------------------------------------------------------
// secondary variables at base + deflected states, SoA records rec[r*n + I]
//   2p:         r = 0: pc, 1: rho_w*mob_w, 2: rho_n*mob_n;  then per deflection k of S_n: 3+3k .. 5+3k
//   1p (table): r = 0: rho, 1: rho*mob;                     then per deflection k of p:   2+2k .. 3+2k
// ------------------------------------------------------------------------------------------------
template <int MODEL, int ND>
__global__ void __launch_bounds__(256) volvars_kernel(AsmParams P)
{
    const int I = blockIdx.x * blockDim.x + threadIdx.x;
    if (I >= P.n) return;
    const size_t n = (size_t)P.n;
    if constexpr (MODEL == DMX_MODEL_2P) {
        const MaterialLaw& law = P.laws[P.region[I]];
        const double Sn = P.cur[2 * (size_t)I + 1];
        {
            const double sw = 1 - Sn;
            P.rec[0 * n + I] = law_pc(law, sw);
            P.rec[1 * n + I] = P.rho[0] * (law_krw(law, sw) / P.mu[0]);
            P.rec[2 * n + I] = P.rho[1] * (law_krn(law, sw) / P.mu[1]);
        }
        const double eps = fd_eps(P, Sn, 1);
#pragma unroll
        for (int k = 0; k < ND; ++k) {
            const double Snk = fd_deflect(P.fd_method, k, Sn, eps);
            const double sw = 1 - Snk;
            P.rec[(3 + 3 * k + 0) * n + I] = law_pc(law, sw);
            P.rec[(3 + 3 * k + 1) * n + I] = P.rho[0] * (law_krw(law, sw) / P.mu[0]);
            P.rec[(3 + 3 * k + 2) * n + I] = P.rho[1] * (law_krn(law, sw) / P.mu[1]);
        }
    } else {
        const double p = P.cur[I];
        {
            const double rho = table_interp(P.table, P.table.rho, p);
            const double mu = table_interp(P.table, P.table.mu, p);
            P.rec[0 * n + I] = rho;
            P.rec[1 * n + I] = rho * (1.0 / mu);
        }
        const double eps = fd_eps(P, p, 0);
#pragma unroll
        for (int k = 0; k < ND; ++k) {
            const double pk = fd_deflect(P.fd_method, k, p, eps);
            const double rho = table_interp(P.table, P.table.rho, pk);
            const double mu = table_interp(P.table, P.table.mu, pk);
            P.rec[(2 + 2 * k + 0) * n + I] = rho;
            P.rec[(2 + 2 * k + 1) * n + I] = rho * (1.0 / mu);
        }
    }
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

// TPFA Darcy flux with gravity + upwinding for all phases: flux/cctpfa/darcyslaw.hh:154-213, flux/upwindscheme.hh:36-54
//   f = tij*(pI - pJ) + (rho*A)*alphaI;  interior: f -= ((rho*tij)/tJ)*(alphaI - alphaJ);  flux = f * upwind(rho*mob)
template <int NPH, bool TABLE>
__device__ __forceinline__ void face_flux(const FaceData<NPH>& F, const CellState<NPH>& sI, const CellState<NPH>& sJ,
                                          bool fullUpwind, double w, double* out)
{
#pragma unroll
    for (int ph = 0; ph < NPH; ++ph) {
        double f = F.tij * (sI.p[ph] - sJ.p[ph]);
        if (F.grav) {
            if constexpr (TABLE) {
                const double rho = F.interior ? (sI.rho[ph] + sJ.rho[ph]) * 0.5 : sJ.rho[ph];
                f = f + rho * F.area * F.alphaI;
                if (F.interior) f -= rho * F.tij / F.tJ * (F.alphaI - F.alphaJ);
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

// Builds the state of cell C (own or neighbour) at the base point or at deflection (pv,k).
// pv < 0: base.  Records come from volvars_kernel.
template <int MODEL, bool TABLE, int NPH>
__device__ __forceinline__ void load_state(const AsmParams& P, int C, int pv, int k, const double* uC, const double* epsC,
                                           CellState<NPH>& s, double* Sn_out)
{
    const size_t n = (size_t)P.n;
    if constexpr (MODEL == DMX_MODEL_2P) {
        double pw = uC[0], Sn = uC[1];
        int r = 0;
        if (pv == 0) pw = fd_deflect(P.fd_method, k, pw, epsC[0]);
        if (pv == 1) { Sn = fd_deflect(P.fd_method, k, Sn, epsC[1]); r = 3 + 3 * k; }
        const double pc = P.rec[(r + 0) * n + C];
        s.p[0] = pw;
        s.p[NPH - 1] = pw + pc;
        s.up[0] = P.rec[(r + 1) * n + C];
        s.up[NPH - 1] = P.rec[(r + 2) * n + C];
        s.rho[0] = P.rho[0];
        s.rho[NPH - 1] = P.rho[1];
        *Sn_out = Sn;
    } else {
        double p = uC[0];
        if (pv == 0) p = fd_deflect(P.fd_method, k, p, epsC[0]);
        s.p[0] = p;
        if constexpr (TABLE) {
            const int r = (pv == 0) ? 2 + 2 * k : 0;
            s.rho[0] = P.rec[(r + 0) * n + C];
            s.up[0] = P.rec[(r + 1) * n + C];
        } else {
            s.rho[0] = P.rho[0];
            s.up[0] = P.rho[0] * (1.0 / P.mu[0]);
        }
        *Sn_out = 0.0;
    }
}

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

template <int MODEL, bool TABLE, int ND>
__global__ void __launch_bounds__(128) assemble_kernel(AsmParams P, int with_jac)
{
    constexpr int NB = (MODEL == DMX_MODEL_2P) ? 2 : 1;   // block size = numEq = phases
    constexpr int NPH = NB;
    const int I = blockIdx.x * blockDim.x + threadIdx.x;
    if (I >= P.n) return;
    const int nx = P.nc[0], ny = P.nc[1], nz = P.nc[2];
    const int ci[3] = {I % nx, (I / nx) % ny, I / (nx * ny)};
    const int stride[3] = {1, nx, nx * ny};
    const int dim = P.dim;
    const int va = dim - 1;
    const bool fullUpwind = (P.upwind_weight == 1.0);
    const double w = P.upwind_weight;
    const double extr = P.extrusion;
    (void)nz;

    // own primary variables, FD steps, states
    double uI[NB], epsI[NB];
#pragma unroll
    for (int e = 0; e < NB; ++e) { uI[e] = P.cur[(size_t)I * NB + e]; epsI[e] = fd_eps(P, uI[e], e); }
    CellState<NPH> sI0, sId[NB][ND];
    double SnI0, SnId[NB][ND];
    load_state<MODEL, TABLE, NPH>(P, I, -1, 0, uI, epsI, sI0, &SnI0);
#pragma unroll
    for (int pv = 0; pv < NB; ++pv)
#pragma unroll
        for (int k = 0; k < ND; ++k) load_state<MODEL, TABLE, NPH>(P, I, pv, k, uI, epsI, sId[pv][k], &SnId[pv][k]);

    const double KI = P.K[I];
    // volume = ((1*w0)*w1)*w2 (AxisAlignedCubeGeometry::volume)
    double vol = 1.0;
    for (int a = 0; a < dim; ++a) vol *= P.width[a][ci[a]];

    // accumulators: residual at the base state and at each own deflection
    double R0[NB], Rd[NB][ND][NB];
#pragma unroll
    for (int e = 0; e < NB; ++e) {
        double source = P.q ? P.q[(size_t)I * NB + e] : 0.0;      // fvlocalresidual.hh:319-333
        source *= vol * extr;
        double r = 0.0;
        r -= source;
        R0[e] = r;
#pragma unroll
        for (int pv = 0; pv < NB; ++pv)
#pragma unroll
            for (int k = 0; k < ND; ++k) Rd[pv][k][e] = r;
    }

    // BCRS positions: columns ascending = -z,-y,-x,diag,+x,+y,+z among the existing neighbours
    bool ex[6];
#pragma unroll
    for (int s = 0; s < 6; ++s) {
        const int a = s >> 1;
        ex[s] = (a < dim) && ((s & 1) ? (ci[a] + 1 < P.nc[a]) : (ci[a] > 0));
    }
    const int rowStart = P.rowptr[I];
    const int posDiag = rowStart + (ex[4] ? 1 : 0) + (ex[2] ? 1 : 0) + (ex[0] ? 1 : 0);
    int pos[6];
    pos[4] = rowStart;
    pos[2] = rowStart + (ex[4] ? 1 : 0);
    pos[0] = pos[2] + (ex[2] ? 1 : 0);
    pos[1] = posDiag + 1;
    pos[3] = pos[1] + (ex[1] ? 1 : 0);
    pos[5] = pos[3] + (ex[3] ? 1 : 0);

#pragma unroll
    for (int s = 0; s < 6; ++s) {
        const int a = s >> 1;
        if (a >= dim) continue;
        const bool hi = (s & 1);
        // face area = product of the widths of the other axes, ascending axis order
        double area = 1.0;
        for (int d = 0; d < dim; ++d)
            if (d != a) area *= P.width[d][ci[d]];
        FaceData<NPH> F;
        F.area = area;
        F.grav = P.enable_gravity && (a == va);
        const double ng = hi ? -P.gravity : P.gravity;       // n.g with g = -gravity*e_va
        F.alphaI = KI * ng * extr;                           // vtmv(n,K,g)*extrusion (common/math.hh:908-913)
        F.alphaJ = 0.0;
        F.tJ = 1.0;
        if (ex[s]) {
            const int J = hi ? I + stride[a] : I - stride[a];
            const int cj = hi ? ci[a] + 1 : ci[a] - 1;
            F.interior = true;
            F.tij = hi ? P.tij[a][I] : P.tij[a][J];
            const double KJ = P.K[J];
            if (F.grav) {
                F.tJ = KJ * extr * (hi ? P.gf_lo[a][cj] : P.gf_hi[a][cj]);
                F.alphaJ = KJ * ng * extr;
                if (!TABLE) {
#pragma unroll
                    for (int ph = 0; ph < NPH; ++ph) {
                        const double rho = (P.rho[ph] + P.rho[ph]) * 0.5;
                        F.c1[ph] = rho * area * F.alphaI;
                        F.c2[ph] = rho * F.tij / F.tJ * (F.alphaI - F.alphaJ);
                    }
                }
            }
            double uJ[NB], epsJ[NB];
#pragma unroll
            for (int e = 0; e < NB; ++e) { uJ[e] = P.cur[(size_t)J * NB + e]; epsJ[e] = fd_eps(P, uJ[e], e); }
            CellState<NPH> sJ0;
            double SnJ;
            load_state<MODEL, TABLE, NPH>(P, J, -1, 0, uJ, epsJ, sJ0, &SnJ);
            double F0[NPH];
            face_flux<NPH, TABLE>(F, sI0, sJ0, fullUpwind, w, F0);
#pragma unroll
            for (int e = 0; e < NB; ++e) R0[e] += F0[e];
#pragma unroll
            for (int pv = 0; pv < NB; ++pv)
#pragma unroll
                for (int k = 0; k < ND; ++k) {
                    double Fk[NPH];
                    face_flux<NPH, TABLE>(F, sId[pv][k], sJ0, fullUpwind, w, Fk);
#pragma unroll
                    for (int e = 0; e < NB; ++e) Rd[pv][k][e] += Fk[e];
                }
            if (with_jac) {
                // A[I][J][e][pv] = FD quotient of the face flux w.r.t. u_J[pv] (cclocalassembler.hh:211-218,254-258,331-332)
                double blk[NB][NB];
#pragma unroll
                for (int pv = 0; pv < NB; ++pv) {
                    double Fd[ND][NPH];
#pragma unroll
                    for (int k = 0; k < ND; ++k) {
                        CellState<NPH> sJd;
                        double dummy;
                        load_state<MODEL, TABLE, NPH>(P, J, pv, k, uJ, epsJ, sJd, &dummy);
                        face_flux<NPH, TABLE>(F, sI0, sJd, fullUpwind, w, Fd[k]);
                    }
#pragma unroll
                    for (int e = 0; e < NB; ++e) {
                        double fk[ND];
#pragma unroll
                        for (int k = 0; k < ND; ++k) fk[k] = Fd[k][e];
                        // reference accumulates the neighbour flux into a zeroed vector first (0 + f)
                        blk[e][pv] = fd_quotient<ND>(P.fd_method, F0[e], fk, epsJ[pv]);
                    }
                }
                double* dst = P.jac + (size_t)pos[s] * (NB * NB);
                if (NB == 2) {
                    reinterpret_cast<double2*>(dst)[0] = make_double2(blk[0][0], blk[0][NB - 1]);
                    reinterpret_cast<double2*>(dst)[1] = make_double2(blk[NB - 1][0], blk[NB - 1][NB - 1]);
                } else
                    dst[0] = blk[0][0];
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
                const double ti = KI * extr * (hi ? P.gf_hi[a][ci[a]] : P.gf_lo[a][ci[a]]);
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
                face_flux<NPH, TABLE>(F, sI0, sD, fullUpwind, w, F0);
#pragma unroll
                for (int e = 0; e < NB; ++e) R0[e] += F0[e];
#pragma unroll
                for (int pv = 0; pv < NB; ++pv)
#pragma unroll
                    for (int k = 0; k < ND; ++k) {
                        double Fk[NPH];
                        face_flux<NPH, TABLE>(F, sId[pv][k], sD, fullUpwind, w, Fk);
#pragma unroll
                        for (int e = 0; e < NB; ++e) Rd[pv][k][e] += Fk[e];
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
                        for (int k = 0; k < ND; ++k) Rd[pv][k][e] += nf;
                }
            }
            // DMX_BC_NONE: outer face of an overlap cell, no scvf exists (tpfa/fvgridgeometry.hh:272-320)
        }
    }

    // storage: fvlocalresidual.hh:274-304: ((S(cur)*extr - S(prev)*extr)*V)/dt, added after flux+source
    if (!P.stationary) {
        const double phiI = P.phi[I];
        const double phiE = 1.0 - (1.0 - phiI);     // porosity = 1 - inert volume fraction
        double prevSt[NB];
        {
            CellState<NPH> sP;
            double SnP = 0.0;
            if constexpr (MODEL == DMX_MODEL_2P) {
                SnP = P.prev[(size_t)I * NB + NB - 1];
                sP.rho[0] = P.rho[0];
                sP.rho[NPH - 1] = P.rho[1];
            } else {
                sP.rho[0] = TABLE ? table_interp(P.table, P.table.rho, P.prev[I]) : P.rho[0];
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
                st[e] /= P.dt;
                acc[e] += st[e];
            }
        };
        addStorage(sI0, SnI0, R0);
#pragma unroll
        for (int pv = 0; pv < NB; ++pv)
#pragma unroll
            for (int k = 0; k < ND; ++k) addStorage(sId[pv][k], SnId[pv][k], Rd[pv][k]);
    }

    bool finite = true;
#pragma unroll
    for (int e = 0; e < NB; ++e) {
        P.residual[(size_t)I * NB + e] = R0[e];
        finite = finite && (fabs(R0[e]) <= DBL_MAX);
    }
    if (!finite) atomicOr(P.flag_nonfinite, 1);

    if (with_jac) {
        double blk[NB][NB];
#pragma unroll
        for (int pv = 0; pv < NB; ++pv)
#pragma unroll
            for (int e = 0; e < NB; ++e) {
                double fk[ND];
#pragma unroll
                for (int k = 0; k < ND; ++k) fk[k] = Rd[pv][k][e];
                blk[e][pv] = fd_quotient<ND>(P.fd_method, R0[e], fk, epsI[pv]);
            }
        double* dst = P.jac + (size_t)posDiag * (NB * NB);
        if (NB == 2) {
            reinterpret_cast<double2*>(dst)[0] = make_double2(blk[0][0], blk[0][NB - 1]);
            reinterpret_cast<double2*>(dst)[1] = make_double2(blk[NB - 1][0], blk[NB - 1][NB - 1]);
        } else
            dst[0] = blk[0][0];
    }
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
    P.K = ctx->d_K; P.phi = ctx->d_phi; P.region = ctx->d_region; P.q = ctx->d_q;
    for (int i = 0; i < 2; ++i) { P.rho[i] = ctx->rho[i]; P.mu[i] = ctx->mu[i]; }
    P.tabulated = ctx->tabulated ? 1 : 0;
    P.table = ctx->d_table;
    P.laws = ctx->d_laws;
    for (int s = 0; s < 6; ++s) {
        P.bc_type[s] = ctx->d_bc_type[s]; P.bc_neumann[s] = ctx->d_bc_neumann[s];
        P.bc_p[s] = ctx->d_bc_p[s]; P.bc_up[s] = ctx->d_bc_up[s]; P.bc_rho[s] = ctx->d_bc_rho[s];
    }
    P.cur = ctx->d_vec[DMX_VEC_CUR]; P.prev = ctx->d_vec[DMX_VEC_PREV];
    P.rec = ctx->d_rec; P.nrec = ctx->nrec;
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
        const double sw = 1 - Sn;
        const double pc = law_pc(law, sw);
        p[0] = pv[0];
        p[1] = pv[0] + pc;
        up[0] = ctx->rho[0] * (law_krw(law, sw) / ctx->mu[0]);
        up[1] = ctx->rho[1] * (law_krn(law, sw) / ctx->mu[1]);
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
    // records
    const int nd = (ctx->opt.fd_method == 5) ? 4 : (ctx->opt.fd_method == 0 ? 2 : 1);
    int nrec = 0;
    if (ctx->model == DMX_MODEL_2P) nrec = 3 * (1 + nd);
    else if (ctx->tabulated) nrec = 2 * (1 + nd);
    if (nrec != ctx->nrec) {
        if (ctx->d_rec) { cudaFree(ctx->d_rec); ctx->d_rec = nullptr; }
        if (nrec) DMX_CUDA(cudaMalloc((void**)&ctx->d_rec, (size_t)nrec * ctx->n * sizeof(double)));
        ctx->nrec = nrec;
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

template <int MODEL, bool TABLE>
static int launch_nd(dmx_ctx* ctx, const AsmParams& P, bool with_jac, bool volvars_only)
{
    const int n = ctx->n;
    const int vt = 256, at = 128;
    const bool needRec = (MODEL == DMX_MODEL_2P) || TABLE;
#define DMX_LAUNCH_ND(ND)                                                                                    \
    do {                                                                                                     \
        if (needRec) {                                                                                       \
            ProfScope ps__(ctx, DMX_K_VOLVARS);                                                              \
            volvars_kernel<MODEL, ND><<<(n + vt - 1) / vt, vt, 0, ctx->stream>>>(P);                         \
            DMX_CHECK_LAUNCH();                                                                              \
        }                                                                                                    \
        if (!volvars_only) {                                                                                 \
            ProfScope ps__(ctx, DMX_K_ASSEMBLY);                                                             \
            assemble_kernel<MODEL, TABLE, ND><<<(n + at - 1) / at, at, 0, ctx->stream>>>(P, with_jac ? 1 : 0); \
            DMX_CHECK_LAUNCH();                                                                              \
        }                                                                                                    \
    } while (0)
    if (P.nd == 1) DMX_LAUNCH_ND(1);
    else if (P.nd == 2) DMX_LAUNCH_ND(2);
    else DMX_LAUNCH_ND(4);
#undef DMX_LAUNCH_ND
    return 0;
}

static int launch_impl(dmx_ctx* ctx, bool with_jac, bool volvars_only)
{
    if (int rc = prepare(ctx)) return rc;
    if (!ctx->opt.stationary && ctx->opt.dt <= 0.0) return fail(ctx, DMX_ERR_USAGE, "assemble: dt must be > 0");
    AsmParams P;
    fill_params(ctx, P);
    if (ctx->model == DMX_MODEL_2P) return launch_nd<DMX_MODEL_2P, false>(ctx, P, with_jac, volvars_only);
    if (ctx->tabulated) return launch_nd<DMX_MODEL_1P, true>(ctx, P, with_jac, volvars_only);
    return launch_nd<DMX_MODEL_1P, false>(ctx, P, with_jac, volvars_only);
}

int launch_assemble(dmx_ctx* ctx, bool with_jacobian) { return launch_impl(ctx, with_jacobian, false); }
int launch_volvars_only(dmx_ctx* ctx) { return launch_impl(ctx, false, true); }

} // namespace dmx
