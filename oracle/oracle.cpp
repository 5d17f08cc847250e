/*
 * oracle/oracle.cpp -- CPU restatement of DuMux's CCTpfa Newton-step hot path.
 * TEST INFRASTRUCTURE ONLY: loaded by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs.  The product (dumux_b200/csrc) never includes or links this file.
 *
 * Every function cites the reference file:line it restates (paths relative to the DuMux tree).
 * dune-istl / dune-grid arithmetic is restated from DUNE 2.10 (not vendored in the reference tree):
 * "parity unpinned" at tight tolerance for those parts, see oracle.h.
 *
 * Floating-point contract (the "canonical operation sequence" the CUDA kernels must reproduce):
 * compiled with -ffp-contract=off; source order of operations below is the order executed; the only
 * fused operations are the explicit fma calls inside orc_det_pow.
 */
#include "oracle.h"
#include "det_math.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

typedef double (*pow_fn)(double, double);
double std_pow_(double x, double y) { return std::pow(x, y); }
double det_pow_(double x, double y) { return orc_det_pow(x, y); }

inline double clamp01(double v) { return v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v); }

// ---------------------------------------------------------------------------------------------
// 2-point cubic spline with prescribed end slopes: dumux/common/spline.hh:263-347 (Spline<Scalar,2>),
// dumux/common/splinecommon_.hh:419-436 (makeFullSystem_), :444-474 (natural part), :478-501 (eval_),
// :504-523 (evalDerivative_).  The 2x2 moment system is solved with Dune::FieldMatrix<2,2>::solve
// (dune-common densematrix.hh, explicit Cramer branch for n==2) [DUNE-ext].
// ---------------------------------------------------------------------------------------------
struct Spline2 {
    double x0 = 0, x1 = 1, y0 = 0, y1 = 0, M0 = 0, M1 = 0;
    void set(double x0_, double x1_, double y0_, double y1_, double m0, double m1)
    {
        x0 = x0_; x1 = x1_; y0 = y0_; y1 = y1_;
        const double h = x1 - x0;
        // makeNaturalSystem_: M[0][0]=2, M[n][n]=2; makeFullSystem_: M[0][1]=1, M[n][n-1]=1
        const double m00 = 2, m01 = 1, m10 = 1, m11 = 2;
        const double d0 = 6 / h * ((y1 - y0) / h - m0);
        const double d1 = 6 / h * (m1 - (y1 - y0) / h);
        double detinv = m00 * m11 - m01 * m10;
        detinv = 1.0 / detinv;
        M0 = detinv * (m11 * d0 - m01 * d1);
        M1 = detinv * (m00 * d1 - m10 * d0);
    }
    double eval(double x) const
    {
        const double h = x1 - x0;
        const double xi = x - x0;
        const double xi1 = x1 - x;
        const double A = (y1 - y0) / h - h / 6 * (M1 - M0);
        const double B = y0 - M0 * (h * h) / 6;
        return M0 * xi1 * xi1 * xi1 / (6 * h) + M1 * xi * xi * xi / (6 * h) + A * xi + B;
    }
    double evalDerivative(double x) const
    {
        const double h = x1 - x0;
        const double xi = x - x0;
        const double xi1 = x1 - x;
        const double A = (y1 - y0) / h - h / 6 * (M1 - M0);
        return -M0 * xi1 * xi1 / (2 * h) + M1 * xi * xi / (2 * h) + A;
    }
};

// ---------------------------------------------------------------------------------------------
// Material law = TwoPMaterialLaw<BaseLaw, Regularization, TwoPEffToAbsDefaultPolicy>
//   wrapper:        dumux/material/fluidmatrixinteractions/2p/materiallaw.hh:104-243
//   eff<->abs:      .../2p/efftoabsdefaultpolicy.hh:106-152
//   Brooks-Corey:   .../2p/brookscorey.hh:100-108 (pc), :165-174 (dpc_dswe), :215-223 (krw), :266-276 (krn),
//                   regularisation :365-477, initPcParameters_ :481-489
//   van Genuchten:  .../2p/vangenuchten.hh pc/krw/krn/derivatives, regularisation (pc spline on (pcHighSwe,1),
//                   krw spline on [krwHighSwe,1), krn spline on (0,krnLowSwe]), initPcParameters_/initKrParameters_
// ---------------------------------------------------------------------------------------------
struct Law {
    int kind = ORC_LAW_BROOKSCOREY;
    bool reg = true;
    double swr = 0, snr = 0;
    // BC
    double pe = 0, lambda = 2;
    // VG
    double alpha = 0, n = 2, m = 0.5, l = 0.5;
    // regularisation thresholds
    double pcLowSwe = 0.01, pcHighSwe = 0.99, krnLowSwe = 0.1, krwHighSwe = 0.9;
    // derived
    double pcLowSwePcValue = 0, pcHighSwePcValue = 0, pcDerivativeLowSw = 0, pcDerivativeHighSwEnd = 0,
           pcDerivativeHighSweThreshold = 0;
    Spline2 pcSpline, krwSpline, krnSpline;
    pow_fn pw = det_pow_;
    int wetting = 0;               // wetting phase index (spatialParams.wettingPhase, 2p/volumevariables.hh:132)

    // efftoabsdefaultpolicy.hh:106-152
    double swToSwe(double sw) const { return (sw - swr) / (1.0 - swr - snr); }
    double sweToSw(double swe) const { return swe * (1.0 - swr - snr) + swr; }
    double dswe_dsw() const { return 1.0 / (1.0 - swr - snr); }
    double dsw_dswe() const { return 1.0 - swr - snr; }

    // ---- base laws (unregularised, effective saturation) ----
    double base_pc(double swe) const
    {
        swe = clamp01(swe);
        if (kind == ORC_LAW_BROOKSCOREY)
            return pe * pw(swe, -1.0 / lambda);                                  // brookscorey.hh:107
        return pw(pw(swe, -1.0 / m) - 1, 1.0 / n) / alpha;                        // vangenuchten.hh pc
    }
    double base_dpc_dswe(double swe) const
    {
        swe = clamp01(swe);
        if (kind == ORC_LAW_BROOKSCOREY)
            return -pe / lambda * pw(swe, -1.0 / lambda - 1.0);                  // brookscorey.hh:173
        const double powSwe = pw(swe, -1 / m);                                    // vangenuchten.hh:179-181
        return -1.0 / alpha * pw(powSwe - 1, 1.0 / n - 1) / n * powSwe / swe / m;
    }
    double base_krw(double swe) const
    {
        swe = clamp01(swe);
        if (kind == ORC_LAW_BROOKSCOREY)
            return pw(swe, 2.0 / lambda + 3.0);                                  // brookscorey.hh:222
        const double r = 1.0 - pw(1.0 - pw(swe, 1.0 / m), m);                     // vangenuchten.hh krw
        return pw(swe, l) * r * r;
    }
    double base_dkrw_dswe(double swe) const
    {
        swe = clamp01(swe);
        if (kind == ORC_LAW_BROOKSCOREY)
            return (2.0 / lambda + 3.0) * pw(swe, 2.0 / lambda + 2.0);           // brookscorey.hh:247
        const double x = 1.0 - pw(swe, 1.0 / m);
        const double xToM = pw(x, m);
        return (1.0 - xToM) * pw(swe, l - 1) * ((1.0 - xToM) * l + 2 * xToM * (1.0 - x) / x);
    }
    double base_krn(double swe) const
    {
        swe = clamp01(swe);
        if (kind == ORC_LAW_BROOKSCOREY) {
            const double exponent = 2.0 / lambda + 1.0;                          // brookscorey.hh:273-275
            const double sne = 1.0 - swe;
            return sne * sne * (1.0 - pw(swe, exponent));
        }
        return pw(1 - swe, l) * pw(1 - pw(swe, 1.0 / m), 2 * m);                  // vangenuchten.hh krn
    }
    double base_dkrn_dswe(double swe) const
    {
        swe = clamp01(swe);
        if (kind == ORC_LAW_BROOKSCOREY) {
            const double lambdaInv = 1.0 / lambda;                               // brookscorey.hh:302-304
            const double swePow = pw(swe, 2 * lambdaInv);
            return 2.0 * (swe - 1.0) * (1.0 + (0.5 + lambdaInv) * swePow - (1.5 + lambdaInv) * swePow * swe);
        }
        const double sne = 1.0 - swe;
        const double x = 1.0 - pw(swe, 1.0 / m);
        return -pw(sne, l - 1.0) * pw(x, 2 * m - 1.0) * (l * x + 2.0 * sne / swe * (1.0 - x));
    }

    // unregularised absolute-saturation versions used by the regularisation init (materiallaw.hh, <false>)
    double pc_noreg(double sw) const { return base_pc(swToSwe(sw)); }
    double dpc_dsw_noreg(double sw) const { return base_dpc_dswe(swToSwe(sw)) * dswe_dsw(); }
    double krw_noreg(double sw) const { return base_krw(swToSwe(sw)); }
    double dkrw_dsw_noreg(double sw) const { return base_dkrw_dswe(swToSwe(sw)) * dswe_dsw(); }
    double krn_noreg(double sw) const { return base_krn(swToSwe(sw)); }
    double dkrn_dsw_noreg(double sw) const { return base_dkrn_dswe(swToSwe(sw)) * dswe_dsw(); }

    void init()
    {
        if (!reg) return;
        if (kind == ORC_LAW_BROOKSCOREY) {
            // brookscorey.hh:481-489
            const double lowSw = sweToSw(pcLowSwe);
            const double highSw = sweToSw(1.0);
            const double dsw = dsw_dswe();
            pcDerivativeLowSw = dpc_dsw_noreg(lowSw) * dsw;
            pcDerivativeHighSwEnd = dpc_dsw_noreg(highSw) * dsw;
            pcLowSwePcValue = pc_noreg(lowSw);
        } else {
            // vangenuchten.hh initPcParameters_
            {
                const double lowSw = sweToSw(pcLowSwe);
                const double highSw = sweToSw(pcHighSwe);
                const double dsw = dsw_dswe();
                pcDerivativeLowSw = dpc_dsw_noreg(lowSw) * dsw;
                pcDerivativeHighSweThreshold = dpc_dsw_noreg(highSw) * dsw;
                pcDerivativeHighSwEnd = 2.0 * (0.0 - pc_noreg(highSw)) / (1.0 - pcHighSwe);
                pcLowSwePcValue = pc_noreg(lowSw);
                pcHighSwePcValue = pc_noreg(highSw);
                if (pcHighSwe < 1.0)
                    pcSpline.set(pcHighSwe, 1.0, pcHighSwePcValue, 0, pcDerivativeHighSweThreshold, pcDerivativeHighSwEnd);
            }
            // vangenuchten.hh initKrParameters_
            {
                const double lowSw = sweToSw(krnLowSwe);
                const double highSw = sweToSw(krwHighSwe);
                const double dsw = dsw_dswe();
                const double krwHighSw = krw_noreg(highSw);
                const double dkrwHighSw = dkrw_dsw_noreg(highSw) * dsw;
                const double krnLowSw = krn_noreg(lowSw);
                const double dkrnLowSw = dkrn_dsw_noreg(lowSw) * dsw;
                if (krwHighSwe < 1.0) krwSpline.set(krwHighSwe, 1.0, krwHighSw, 1.0, dkrwHighSw, 0.0);
                if (krnLowSwe > 0.0) krnSpline.set(0.0, krnLowSwe, 1.0, krnLowSw, 0.0, dkrnLowSw);
            }
        }
    }

    // materiallaw.hh:104-118 + brookscorey.hh:365-380 / vangenuchten.hh regularised pc
    double pc(double sw) const
    {
        const double swe = swToSwe(sw);
        if (reg) {
            if (kind == ORC_LAW_BROOKSCOREY) {
                if (swe <= pcLowSwe) return pcLowSwePcValue + pcDerivativeLowSw * (swe - pcLowSwe);
                else if (swe >= 1.0) return pcDerivativeHighSwEnd * (swe - 1.0) + pe;
            } else {
                if (swe <= pcLowSwe) return pcLowSwePcValue + pcDerivativeLowSw * (swe - pcLowSwe);
                else if (swe >= 1.0) return pcDerivativeHighSwEnd * (swe - 1.0);
                else if (swe > pcHighSwe) return pcSpline.eval(swe);
            }
        }
        return base_pc(swe);
    }
    double dpc_dsw(double sw) const
    {
        const double swe = swToSwe(sw);
        if (reg) {
            if (swe <= pcLowSwe) return pcDerivativeLowSw * dswe_dsw();
            else if (swe >= 1.0) return pcDerivativeHighSwEnd * dswe_dsw();
            else if (kind == ORC_LAW_VANGENUCHTEN && swe > pcHighSwe) return pcSpline.evalDerivative(swe) * dswe_dsw();
        }
        return base_dpc_dswe(swe) * dswe_dsw();
    }
    // materiallaw.hh:176-188 + regularised krw
    double krw(double sw) const
    {
        const double swe = swToSwe(sw);
        if (reg) {
            if (swe <= 0.0) return 0.0;
            else if (swe >= 1.0) return 1.0;
            else if (kind == ORC_LAW_VANGENUCHTEN && swe >= krwHighSwe) return krwSpline.eval(swe);
        }
        return base_krw(swe);
    }
    double dkrw_dsw(double sw) const
    {
        const double swe = swToSwe(sw);
        if (reg) {
            if (swe <= 0.0) return 0.0;
            else if (swe >= 1.0) return 0.0;
            else if (kind == ORC_LAW_VANGENUCHTEN && swe >= krwHighSwe) return krwSpline.evalDerivative(swe) * dswe_dsw();
        }
        return base_dkrw_dswe(swe) * dswe_dsw();
    }
    // materiallaw.hh:210-222 + regularised krn
    double krn(double sw) const
    {
        const double swe = swToSwe(sw);
        if (reg) {
            if (swe <= 0.0) return 1.0;
            else if (swe >= 1.0) return 0.0;
            else if (kind == ORC_LAW_VANGENUCHTEN && swe <= krnLowSwe) return krnSpline.eval(swe);
        }
        return base_krn(swe);
    }
    double dkrn_dsw(double sw) const
    {
        const double swe = swToSwe(sw);
        if (reg) {
            if (swe <= 0.0) return 0.0;
            else if (swe >= 1.0) return 0.0;
            else if (kind == ORC_LAW_VANGENUCHTEN && swe <= krnLowSwe) return krnSpline.evalDerivative(swe) * dswe_dsw();
        }
        return base_dkrn_dswe(swe) * dswe_dsw();
    }
};

// ---------------------------------------------------------------------------------------------
// Fluids.  Incompressible constants: SimpleH2O rho=1000, mu=1e-3 (dumux/material/components/simpleh2o.hh:265-312),
// Trichloroethene rho=1460, mu=5.7e-4 (trichloroethene.hh:148-172).  Tabulated liquid: bilinear (T,p) lookup
// dumux/material/components/tabulatedcomponent.hh:1166-1203 (interpolateTP_), table layout values[iT + iP*nT].
// ---------------------------------------------------------------------------------------------
struct Fluids {
    double rho[2] = {1000.0, 1460.0};
    double mu[2] = {1e-3, 5.7e-4};
    bool tabulated = false;
    int nT = 0, nP = 0;
    double Tmin = 0, Tmax = 0, T = 293.15;
    std::vector<double> pmin, pmax, rhoTab, muTab;

    double interp(const std::vector<double>& values, double p) const
    {
        double alphaT = (nT - 1) * (T - Tmin) / (Tmax - Tmin);                       // tempIdx
        if (alphaT < 0 - 1e-7 * nT || alphaT >= nT - 1 + 1e-7 * nT) return std::numeric_limits<double>::quiet_NaN();
        const int iT = std::min(std::max((int)alphaT, 0), nT - 2);
        alphaT -= iT;
        double alphaP1 = (nP - 1) * (p - pmin[iT]) / (pmax[iT] - pmin[iT]);           // pressIdx(p, iT)
        double alphaP2 = (nP - 1) * (p - pmin[iT + 1]) / (pmax[iT + 1] - pmin[iT + 1]);
        const int iP1 = std::min(std::max((int)alphaP1, 0), nP - 2);
        const int iP2 = std::min(std::max((int)alphaP2, 0), nP - 2);
        alphaP1 -= iP1;
        alphaP2 -= iP2;
        return values[(iT) + (iP1)*nT] * (1 - alphaT) * (1 - alphaP1) + values[(iT) + (iP1 + 1) * nT] * (1 - alphaT) * (alphaP1)
             + values[(iT + 1) + (iP2)*nT] * (alphaT) * (1 - alphaP2) + values[(iT + 1) + (iP2 + 1) * nT] * (alphaT) * (alphaP2);
    }
    double density(int phase, double p) const { return tabulated ? interp(rhoTab, p) : rho[phase]; }
    double viscosity(int phase, double p) const { return tabulated ? interp(muTab, p) : mu[phase]; }
};

// ---------------------------------------------------------------------------------------------
// Secondary variables.
//   2p: dumux/porousmediumflow/2p/volumevariables.hh:75-102 (update), :118-190 (completeFluidState, p0s1,
//       wetting phase = phase 0); 1p: dumux/porousmediumflow/1p/volumevariables.hh:65-125.
//   porosity = 1 - inertVolumeFraction, inertVolumeFraction = 1 - porosity(x)
//       (dumux/common/fvporousmediumspatialparams.hh:112-118, material/solidstates/inertsolidstate.hh:55-61).
// ---------------------------------------------------------------------------------------------
struct VolVars {
    double p[2], S[2], rho[2], mu[2], mob[2], pc, porosity, K, extr;
    double Kd[3];      // permeability entering a face with normal e_a: K_aa of a diagonal tensor, else the scalar K
};

} // namespace

struct orc_problem {
    int model = ORC_MODEL_1P, b = 1, dim = 2;
    int nc[3] = {1, 1, 1};
    std::vector<double> xn[3];                 // node coordinates per axis
    int n = 0;
    orc_options opt;
    std::vector<double> K, phi, q;
    std::vector<double> Kdiag[3];              // diagonal permeability tensor per axis (empty: scalar K)
    std::vector<int> region;
    std::vector<Law> laws;
    Fluids fluids;
    std::vector<int> bcType[6];
    std::vector<double> bcVal[6];
    std::vector<int> rowptr, colidx;
    int linearSolver = ORC_SOLVER_BICGSTAB, gmresRestart = 10;   // orc_set_linear_solver
    double tracerD = 0.0, tracerTau = 0.5;     // binary diffusion coefficient of the tracer and SpatialParams.Tortuosity
    std::vector<double> tracerDisp;            // mechanical dispersion: n.D.n of the dispersion tensor at every face [n][2*dim]
    bool volumeFluxMode = false;               // upwind term = mobility only (examples/1ptracer/main.cc:170)

    // ---- grid geometry: YaspGrid equidistant / tensor coordinates, AxisAlignedCubeGeometry [DUNE-ext] ----
    int idx(int i, int j, int k) const { return i + nc[0] * (j + nc[1] * k); }
    void ijk(int c, int* o) const { o[0] = c % nc[0]; o[1] = (c / nc[0]) % nc[1]; o[2] = c / (nc[0] * nc[1]); }
    double center(int a, int i) const { return 0.5 * (xn[a][i] + xn[a][i + 1]); }
    double width(int a, int i) const { return xn[a][i + 1] - xn[a][i]; }
    double volume(const int* c) const
    {
        double vol = 1.0;
        for (int a = 0; a < dim; ++a) vol *= width(a, c[a]);
        return vol;
    }
    double faceArea(int axis, const int* c) const
    {
        double vol = 1.0;
        for (int a = 0; a < dim; ++a)
            if (a != axis) vol *= width(a, c[a]);
        return vol;
    }
    int sideFaces(int side) const
    {
        const int a = side / 2;
        int nf = 1;
        for (int d = 0; d < 3; ++d)
            if (d != a) nf *= nc[d];
        return nf;
    }
    // face index within its side: lower remaining axis fastest
    int sideFaceIndex(int side, const int* c) const
    {
        const int a = side / 2;
        if (a == 0) return c[1] + nc[1] * c[2];
        if (a == 1) return c[0] + nc[0] * c[2];
        return c[0] + nc[0] * c[1];
    }
    int neighbor(const int* c, int side) const
    {
        const int a = side / 2;
        if (a >= dim) return -1;
        int d[3] = {c[0], c[1], c[2]};
        d[a] += (side & 1) ? 1 : -1;
        if (d[a] < 0 || d[a] >= nc[a]) return -1;
        return idx(d[0], d[1], d[2]);
    }

    // gravity vector: dumux/common/fvspatialparams.hh:48-52 (g[dimWorld-1] = -9.81 if enabled)
    void gravityVec(double* g) const
    {
        g[0] = g[1] = g[2] = 0.0;
        if (opt.enable_gravity) g[dim - 1] = -opt.gravity;
    }

    // ---- volume variables ----
    void updateVolVars(VolVars& v, const double* priVars, int cell) const
    {
        const double phiInert = 1.0 - phi[cell];
        v.porosity = 1.0 - phiInert;
        v.K = K[cell];
        for (int a = 0; a < 3; ++a) v.Kd[a] = Kdiag[a].empty() ? K[cell] : Kdiag[a][cell];
        v.extr = opt.extrusion;
        if (model == ORC_MODEL_1P) {
            v.S[0] = 1.0; v.S[1] = 0.0;
            v.p[0] = priVars[0]; v.p[1] = 0.0;
            v.rho[0] = fluids.density(0, v.p[0]);
            v.mu[0] = fluids.viscosity(0, v.p[0]);
            v.mob[0] = 1.0 / v.mu[0];                                   // 1p/volumevariables.hh:178
            v.rho[1] = v.mu[1] = v.mob[1] = 0.0; v.pc = 0.0;
            return;
        }
        // p0s1 formulation (2p/volumevariables.hh:133-152): priVars = (p of phase 0, S of phase 1); the law is evaluated at
        // the saturation of the WETTING phase; p1 = p0 + pc if phase 0 wets, p0 - pc if phase 1 wets
        const Law& law = laws[region[cell]];
        const int w = law.wetting, nw = 1 - w;
        const double Sn = priVars[1];
        v.p[0] = priVars[0];
        v.S[1] = Sn;
        v.S[0] = 1 - Sn;
        v.pc = law.pc(v.S[w]);
        v.p[1] = (w == 1) ? priVars[0] - v.pc : priVars[0] + v.pc;
        for (int ph = 0; ph < 2; ++ph) {
            v.mu[ph] = fluids.viscosity(ph, v.p[ph]);
            v.rho[ph] = fluids.density(ph, v.p[ph]);
        }
        v.mob[w] = law.krw(v.S[w]) / v.mu[w];                           // :90-96
        v.mob[nw] = law.krn(v.S[w]) / v.mu[nw];
    }

    // ---- TPFA transmissibility: dumux/discretization/cellcentered/tpfa/computetransmissibility.hh:69-80 ----
    // scvf of cell c on `side`; scv = cell sc (either c itself or its neighbour across that face).
    double computeTpfaTransmissibility(const int* c, int side, const int* sc, double t, double extr) const
    {
        const int a = side / 2;
        double ip[3] = {0, 0, 0}, ctr[3] = {0, 0, 0}, nrm[3] = {0, 0, 0};
        for (int d = 0; d < dim; ++d) {
            ip[d] = (d == a) ? ((side & 1) ? xn[a][c[a] + 1] : xn[a][c[a]]) : center(d, c[d]);
            // AxisAlignedCubeGeometry::center of the (degenerate) face box: 0.5*(x+x) = x exactly
            if (d == a) ip[d] = 0.5 * (ip[d] + ip[d]);
            ctr[d] = center(d, sc[d]);
        }
        nrm[a] = (side & 1) ? 1.0 : -1.0;
        double dv[3];
        double n2 = 0.0;
        for (int d = 0; d < dim; ++d) { dv[d] = ip[d] - ctr[d]; n2 += dv[d] * dv[d]; }
        double dot = 0.0;
        for (int d = 0; d < dim; ++d) { dv[d] /= n2; dot += dv[d] * nrm[d]; }
        return t * extr * dot;
    }

    // ---- advective TPFA flux of one phase over the face `side` of `cI`: flux/cctpfa/darcyslaw.hh:154-213,
    //      transmissibility :218-259, upwinding flux/upwindscheme.hh:36-54,
    //      phase mass flux porousmediumflow/immiscible/localresidual.hh:98-127 ----
    void computeFlux(double* flux, const int* cI, int side, const VolVars& in, const VolVars& out, bool boundary, const int* cJ,
                     double* darcy = nullptr) const
    {
        const int a = side / 2;
        const double area = faceArea(a, cI);
        // vtmv(n, K, n) = K_aa for the axis-aligned normal of this face (diagonal tensor or scalar)
        const double ti = computeTpfaTransmissibility(cI, side, cI, in.Kd[a], in.extr);
        double tij, tj = 0.0;
        if (boundary)
            tij = area * ti;
        else {
            tj = -1.0 * computeTpfaTransmissibility(cI, side, cJ, out.Kd[a], out.extr);
            if (ti * tj <= 0.0) tij = 0;
            else tij = area * (ti * tj) / (ti + tj);
        }
        double g[3];
        gravityVec(g);
        double nrm[3] = {0, 0, 0};
        nrm[a] = (side & 1) ? 1.0 : -1.0;
        const int nph = (model == ORC_MODEL_1P) ? 1 : 2;
        for (int ph = 0; ph < nph; ++ph) {
            double f;
            if (opt.enable_gravity) {
                const double rho = boundary ? out.rho[ph] : (in.rho[ph] + out.rho[ph]) * 0.5;
                const double pInside = in.p[ph];
                const double pOutside = out.p[ph];
                double ng = 0.0;
                for (int d = 0; d < dim; ++d) ng += nrm[d] * g[d];
                const double alpha_inside = in.Kd[a] * ng * in.extr;  // vtmv(n,K,g)*extr, dumux/common/math.hh:908-913
                f = tij * (pInside - pOutside) + rho * area * alpha_inside;
                if (!boundary) {
                    const double outsideTi = tj;                       // same expression as in calculateTransmissibility
                    const double alpha_outside = out.Kd[a] * ng * out.extr;
                    f -= rho * tij / outsideTi * (alpha_inside - alpha_outside);
                }
            } else {
                f = tij * (in.p[ph] - out.p[ph]);
            }
            if (darcy) darcy[ph] = f;        // AdvectionType::flux (darcyslaw.hh:154-213): what the analytic derivatives upwind with
            // upwindscheme.hh:43-53, upwind term = density*mobility (immiscible/localresidual.hh:113-114)
            const double w = opt.upwind_weight;
            const double upIn = volumeFluxMode ? in.mob[ph] : in.rho[ph] * in.mob[ph];
            const double upOut = volumeFluxMode ? out.mob[ph] : out.rho[ph] * out.mob[ph];
            double mult;
            if (std::signbit(f)) mult = w * upOut + (1.0 - w) * upIn;
            else mult = w * upIn + (1.0 - w) * upOut;
            flux[ph] = f * mult;
        }
    }

    // Dirichlet "outside" vol vars are built with the INSIDE cell's spatial parameters:
    // dumux/discretization/cellcentered/tpfa/elementvolumevariables.hh:318-346
    // Flux over one scvf of cell I incl. boundary dispatch: dumux/assembly/cclocalresidual.hh:64-105
    // vvI: (possibly deflected) vol vars of I; vvJ: vol vars of the neighbour (ignored on boundaries)
    bool evalFlux(double* flux, int I, const int* cI, int side, const VolVars& vvI, const VolVars* vvJ) const
    {
        const int a = side / 2;
        flux[0] = flux[1] = 0.0;
        if (a >= dim) return false;
        const int J = neighbor(cI, side);
        if (J >= 0) {
            int cJ[3];
            ijk(J, cJ);
            double f[2];
            computeFlux(f, cI, side, vvI, *vvJ, false, cJ);
            for (int e = 0; e < b; ++e) flux[e] += f[e];
            return true;
        }
        const int fidx = sideFaceIndex(side, cI);
        const int type = bcType[side].empty() ? ORC_BC_NEUMANN : bcType[side][fidx];
        if (type == ORC_BC_NONE) return false;   // processor boundary of an overlap cell: no scvf (tpfa/fvgridgeometry.hh:272-320)
        if (type == ORC_BC_DIRICHLET) {
            VolVars bv;
            double pv[2] = {0, 0};
            for (int e = 0; e < b; ++e) pv[e] = bcVal[side][(size_t)fidx * b + e];
            updateVolVars(bv, pv, I);
            double f[2];
            computeFlux(f, cI, side, vvI, bv, true, nullptr);
            for (int e = 0; e < b; ++e) flux[e] += f[e];
        } else {
            // Neumann: cclocalresidual.hh:89-98: neumannFluxes *= area*extrusion
            const double area = faceArea(a, cI);
            for (int e = 0; e < b; ++e) {
                double nf = bcVal[side].empty() ? 0.0 : bcVal[side][(size_t)fidx * b + e];
                nf *= area * vvI.extr;
                flux[e] += nf;
            }
        }
        return true;
    }

    // storage term: porousmediumflow/immiscible/localresidual.hh:64-83
    void computeStorage(double* s, const VolVars& v) const
    {
        const int nph = (model == ORC_MODEL_1P) ? 1 : 2;
        for (int ph = 0; ph < nph; ++ph) s[ph] = v.porosity * v.rho[ph] * v.S[ph];
    }

    // complete local residual of element I: dumux/assembly/fvlocalassemblerbase.hh:108-135,
    // flux+source dumux/assembly/fvlocalresidual.hh:150-169 (source first :319-333, then all scvfs in
    // intersection order -x,+x,-y,+y,-z,+z), storage :274-304 added afterwards.
    void evalLocalResidual(double* res, int I, const int* cI, const VolVars& vvI, const VolVars* nb /*[6]*/, const VolVars* prevI) const
    {
        const double vol = volume(cI);
        for (int e = 0; e < b; ++e) {
            double source = q.empty() ? 0.0 : q[(size_t)I * b + e];
            source *= vol * vvI.extr;
            res[e] = 0.0;
            res[e] -= source;
        }
        for (int side = 0; side < 2 * dim; ++side) {
            double f[2];
            if (evalFlux(f, I, cI, side, vvI, &nb[side]))
                for (int e = 0; e < b; ++e) res[e] += f[e];
        }
        if (!opt.stationary) {
            double prevStorage[2], storage[2];
            computeStorage(prevStorage, *prevI);
            computeStorage(storage, vvI);
            for (int e = 0; e < b; ++e) {
                prevStorage[e] *= prevI->extr;
                storage[e] *= vvI.extr;
                storage[e] -= prevStorage[e];
                storage[e] *= vol;
                storage[e] /= opt.dt;
                double st = 0.0;
                st += storage[e];
                res[e] += st;
            }
        }
    }

    // advectionTij of the flux-variables cache (flux/cctpfa/darcyslaw.hh:218-259): what computeFlux uses, factored out for the
    // analytic flux derivatives
    double advectionTij(const int* cI, int side, double KI, double extrI, bool boundary, const int* cJ, double KJ, double extrJ) const
    {
        const double area = faceArea(side / 2, cI);
        const double ti = computeTpfaTransmissibility(cI, side, cI, KI, extrI);
        if (boundary) return area * ti;
        const double tj = -1.0 * computeTpfaTransmissibility(cI, side, cJ, KJ, extrJ);
        if (ti * tj <= 0.0) return 0;
        return area * (ti * tj) / (ti + tj);
    }

    // ---------------------------------------------------------------------------------------------
    // DiffMethod::analytic for the incompressible 1p model: CCLocalAssembler<analytic, implicit>
    // (dumux/assembly/cclocalassembler.hh:490-600) with OnePIncompressibleLocalResidual
    // (dumux/porousmediumflow/1p/incompressiblelocalresidual.hh:51-60 no storage derivative, :76-123 flux derivatives
    // A[I][I] += tij*up, A[I][J] -= tij*up with up = density/viscosity, :204-221 Dirichlet faces A[I][I] += tij*up;
    // Neumann faces: addRobinFluxDerivatives is empty).  What test_1p_incompressible_tpfa runs by default.
    // ---------------------------------------------------------------------------------------------
    void assembleElementAnalytic1p(int I, const double* cur, const double* prev, double* residual, double* jac) const
    {
        int cI[3];
        ijk(I, cI);
        VolVars vvI, prevVV, nb[6];
        updateVolVars(vvI, cur + (size_t)I * b, I);
        if (!opt.stationary) updateVolVars(prevVV, prev + (size_t)I * b, I);
        int nbIdx[6], nbC[6][3];
        for (int side = 0; side < 6; ++side) {
            nbIdx[side] = (side < 2 * dim) ? neighbor(cI, side) : -1;
            if (nbIdx[side] >= 0) {
                ijk(nbIdx[side], nbC[side]);
                updateVolVars(nb[side], cur + (size_t)nbIdx[side] * b, nbIdx[side]);
            }
        }
        double orig0[2];
        evalLocalResidual(orig0, I, cI, vvI, nb, &prevVV);
        if (residual) residual[I] = orig0[0];
        if (!jac) return;
        const double up = vvI.rho[0] / vvI.mu[0];
        const int kd = findEntry(I, I);
        for (int side = 0; side < 2 * dim; ++side) {
            const int J = nbIdx[side];
            if (J >= 0) {
                const double deriv = advectionTij(cI, side, vvI.Kd[side / 2], vvI.extr, false, nbC[side], nb[side].Kd[side / 2], nb[side].extr) * up;
                jac[kd] += deriv;
                jac[findEntry(I, J)] -= deriv;
            } else {
                const int fidx = sideFaceIndex(side, cI);
                const int type = bcType[side].empty() ? ORC_BC_NEUMANN : bcType[side][fidx];
                if (type == ORC_BC_DIRICHLET) {
                    const double deriv = advectionTij(cI, side, vvI.Kd[side / 2], vvI.extr, true, nullptr, 0.0, 0.0) * up;
                    jac[kd] += deriv;
                }
            }
        }
    }

    // ---------------------------------------------------------------------------------------------
    // DiffMethod::analytic for the incompressible 2p model (p0-s1, phase 0 wetting): CCLocalAssembler<analytic, implicit>
    // (dumux/assembly/cclocalassembler.hh:490-600) with TwoPIncompressibleLocalResidual
    // (dumux/porousmediumflow/2p/incompressiblelocalresidual.hh:80-101 storage, :137-234 TPFA flux derivatives,
    // :420-481 Dirichlet faces; Neumann faces contribute nothing).  Blocks are [eq][priVar], priVars = (p_w, S_n).
    // What test_2p_incompressible_tpfa_analytic runs.
    // ---------------------------------------------------------------------------------------------
    void assembleElementAnalytic2p(int I, const double* cur, const double* prev, double* residual, double* jac) const
    {
        int cI[3];
        ijk(I, cI);
        VolVars vvI, prevVV, nb[6];
        updateVolVars(vvI, cur + (size_t)I * b, I);
        if (!opt.stationary) updateVolVars(prevVV, prev + (size_t)I * b, I);
        int nbIdx[6], nbC[6][3];
        for (int side = 0; side < 6; ++side) {
            nbIdx[side] = (side < 2 * dim) ? neighbor(cI, side) : -1;
            if (nbIdx[side] >= 0) {
                ijk(nbIdx[side], nbC[side]);
                updateVolVars(nb[side], cur + (size_t)nbIdx[side] * b, nbIdx[side]);
            }
        }
        double orig0[2];
        evalLocalResidual(orig0, I, cI, vvI, nb, &prevVV);
        if (residual)
            for (int e = 0; e < 2; ++e) residual[(size_t)I * 2 + e] = orig0[e];
        if (!jac) return;
        double* AII = jac + (size_t)findEntry(I, I) * 4;
        if (!opt.stationary) {
            // :96-101 (Extrusion::volume of the default NoExtrusion: the scv volume)
            const double poreVolume = volume(cI) * vvI.porosity;
            AII[0 * 2 + 1] -= poreVolume * vvI.rho[0] / opt.dt;
            AII[1 * 2 + 1] += poreVolume * vvI.rho[1] / opt.dt;
        }
        const Law& lawI = laws[region[I]];
        const double w = opt.upwind_weight;
        const double rho_w = vvI.rho[0], rho_n = vvI.rho[1];
        const double rhow_muw = rho_w / vvI.mu[0], rhon_mun = rho_n / vvI.mu[1];
        const double insideSw = vvI.S[0];
        const double dKrw_dSn_inside = -1.0 * lawI.dkrw_dsw(insideSw);
        const double dKrn_dSn_inside = -1.0 * lawI.dkrn_dsw(insideSw);
        const double dpc_dSn_inside = -1.0 * lawI.dpc_dsw(insideSw);
        for (int side = 0; side < 2 * dim; ++side) {
            const int J = nbIdx[side];
            VolVars bv;
            const VolVars* out = nullptr;
            bool boundary = false;
            if (J >= 0) out = &nb[side];
            else {
                const int fidx = sideFaceIndex(side, cI);
                const int type = bcType[side].empty() ? ORC_BC_NEUMANN : bcType[side][fidx];
                if (type != ORC_BC_DIRICHLET) continue;
                double pv[2] = {bcVal[side][(size_t)fidx * 2], bcVal[side][(size_t)fidx * 2 + 1]};
                updateVolVars(bv, pv, I);
                out = &bv;
                boundary = true;
            }
            double fl[2], darcy[2];
            computeFlux(fl, cI, side, vvI, *out, boundary, boundary ? nullptr : nbC[side], darcy);
            const double flux_w = darcy[0], flux_n = darcy[1];
            const double insideWeight_w = std::signbit(flux_w) ? (1.0 - w) : w;
            const double outsideWeight_w = 1.0 - insideWeight_w;
            const double insideWeight_n = std::signbit(flux_n) ? (1.0 - w) : w;
            const double outsideWeight_n = 1.0 - insideWeight_n;
            const double rhowKrw_muw_inside = rho_w * vvI.mob[0], rhonKrn_mun_inside = rho_n * vvI.mob[1];
            const double rhowKrw_muw_outside = rho_w * out->mob[0], rhonKrn_mun_outside = rho_n * out->mob[1];
            const double tij = boundary ? advectionTij(cI, side, vvI.Kd[side / 2], vvI.extr, true, nullptr, 0.0, 0.0)
                                        : advectionTij(cI, side, vvI.Kd[side / 2], vvI.extr, false, nbC[side], out->Kd[side / 2], out->extr);
            const double up_w = rhowKrw_muw_inside * insideWeight_w + rhowKrw_muw_outside * outsideWeight_w;
            const double up_n = rhonKrn_mun_inside * insideWeight_n + rhonKrn_mun_outside * outsideWeight_n;
            if (boundary) {
                AII[0] += tij * up_w;
                AII[1] += rhow_muw * flux_w * dKrw_dSn_inside * insideWeight_w;
                AII[2] += tij * up_n;
                AII[3] += rhon_mun * flux_n * dKrn_dSn_inside * insideWeight_n;
                AII[3] += tij * dpc_dSn_inside * up_n;
                continue;
            }
            const Law& lawJ = laws[region[J]];
            const double outsideSw = out->S[0];
            const double dKrw_dSn_outside = -1.0 * lawJ.dkrw_dsw(outsideSw);
            const double dKrn_dSn_outside = -1.0 * lawJ.dkrn_dsw(outsideSw);
            const double dpc_dSn_outside = -1.0 * lawJ.dpc_dsw(outsideSw);
            const double rho_mu_flux_w = rhow_muw * flux_w, rho_mu_flux_n = rhon_mun * flux_n;
            const double tij_up_w = tij * up_w, tij_up_n = tij * up_n;
            double* AIJ = jac + (size_t)findEntry(I, J) * 4;
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
        }
    }

    // FD epsilon: dumux/assembly/numericepsilon.hh:47-51, dumux/common/numericdifferentiation.hh:36-41
    double fdEps(double priVar, int pvIdx) const
    {
        return opt.privar_magnitude[pvIdx] > 0.0 ? opt.base_eps * opt.privar_magnitude[pvIdx]
                                                 : opt.base_eps * (std::fabs(priVar) + 1.0);
    }

    void buildPattern()
    {
        // dumux/assembly/jacobianpattern.hh:27-52 + cellcentered/connectivitymap.hh:66-115: (I,I) and (J,I) for all
        // face neighbours; BCRS columns ascending (Dune::MatrixIndexSet::exportIdx) [DUNE-ext]
        rowptr.assign(n + 1, 0);
        colidx.clear();
        for (int I = 0; I < n; ++I) {
            int c[3];
            ijk(I, c);
            int cols[7], nn = 0;
            cols[nn++] = I;
            for (int side = 0; side < 2 * dim; ++side) {
                const int J = neighbor(c, side);
                if (J >= 0) cols[nn++] = J;
            }
            std::sort(cols, cols + nn);
            for (int k = 0; k < nn; ++k) colidx.push_back(cols[k]);
            rowptr[I + 1] = (int)colidx.size();
        }
    }
    int findEntry(int row, int col) const
    {
        for (int k = rowptr[row]; k < rowptr[row + 1]; ++k)
            if (colidx[k] == col) return k;
        return -1;
    }

    // ---------------------------------------------------------------------------------------------
    // Column-wise numeric-differentiation assembly of one element:
    // dumux/assembly/cclocalassembler.hh:164-348 (assembleJacobianAndResidualImpl), FD formula
    // dumux/common/numericdifferentiation.hh:67-123.
    // ---------------------------------------------------------------------------------------------
    void assembleElement(int I, const double* cur, const double* prev, double* residual, double* jac) const
    {
        int cI[3];
        ijk(I, cI);
        VolVars vvI, prevVV, nb[6];
        int nbIdx[6];
        int nbC[6][3];
        updateVolVars(vvI, cur + (size_t)I * b, I);
        if (!opt.stationary) updateVolVars(prevVV, prev + (size_t)I * b, I);
        for (int side = 0; side < 6; ++side) {
            nbIdx[side] = (side < 2 * dim) ? neighbor(cI, side) : -1;
            if (nbIdx[side] >= 0) {
                ijk(nbIdx[side], nbC[side]);
                updateVolVars(nb[side], cur + (size_t)nbIdx[side] * b, nbIdx[side]);
            }
        }
        // origResiduals[0] (:191) and neighbour fluxes in the undeflected state (:211-218)
        double orig0[2];
        evalLocalResidual(orig0, I, cI, vvI, nb, &prevVV);
        double origN[6][2];
        for (int side = 0; side < 2 * dim; ++side) {
            origN[side][0] = origN[side][1] = 0.0;
            if (nbIdx[side] < 0) continue;
            // flux from J's side over the shared face: J's face is the opposite side
            double f[2];
            evalFlux(f, nbIdx[side], nbC[side], side ^ 1, nb[side], &vvI);
            for (int e = 0; e < b; ++e) origN[side][e] += f[e];
        }
        if (residual)
            for (int e = 0; e < b; ++e) residual[(size_t)I * b + e] = orig0[e];
        if (!jac) return;

        for (int pv = 0; pv < b; ++pv) {
            const double x0 = cur[(size_t)I * b + pv];
            const double eps = fdEps(x0, pv);
            // evalResiduals(priVar) (:241-261)
            auto evalResiduals = [&](double priVar, double* r0, double rN[6][2]) {
                double pvs[2] = {cur[(size_t)I * b], b > 1 ? cur[(size_t)I * b + 1] : 0.0};
                pvs[pv] = priVar;
                VolVars defl;
                updateVolVars(defl, pvs, I);
                evalLocalResidual(r0, I, cI, defl, nb, &prevVV);
                for (int side = 0; side < 2 * dim; ++side) {
                    rN[side][0] = rN[side][1] = 0.0;
                    if (nbIdx[side] < 0) continue;
                    double f[2];
                    evalFlux(f, nbIdx[side], nbC[side], side ^ 1, nb[side], &defl);
                    for (int e = 0; e < b; ++e) rN[side][e] += f[e];
                }
            };
            double d0[2] = {0, 0}, dN[6][2];
            for (int s = 0; s < 6; ++s) dN[s][0] = dN[s][1] = 0.0;
            const int method = opt.fd_method;
            auto forEach = [&](auto&& fn) {
                for (int e = 0; e < b; ++e) fn(d0[e], -1, e);
                for (int side = 0; side < 2 * dim; ++side)
                    if (nbIdx[side] >= 0)
                        for (int e = 0; e < b; ++e) fn(dN[side][e], side, e);
            };
            if (method == 5) {
                double a0[2], aN[6][2], t0[2], tN[6][2];
                evalResiduals(x0 + eps, a0, aN);
                evalResiduals(x0 - eps, t0, tN);
                forEach([&](double& d, int s, int e) { d = (s < 0 ? a0[e] : aN[s][e]); d -= (s < 0 ? t0[e] : tN[s][e]); d *= 8.0; });
                evalResiduals(x0 - 2.0 * eps, t0, tN);
                forEach([&](double& d, int s, int e) { d += (s < 0 ? t0[e] : tN[s][e]); });
                evalResiduals(x0 + 2.0 * eps, t0, tN);
                forEach([&](double& d, int s, int e) { d -= (s < 0 ? t0[e] : tN[s][e]); d /= 12.0 * eps; });
            } else {
                double delta = 0.0;
                if (method >= 0) {
                    delta += eps;
                    double a0[2], aN[6][2];
                    evalResiduals(x0 + eps, a0, aN);
                    forEach([&](double& d, int s, int e) { d = (s < 0 ? a0[e] : aN[s][e]); });
                } else
                    forEach([&](double& d, int s, int e) { d = (s < 0 ? orig0[e] : origN[s][e]); });
                if (method <= 0) {
                    delta += eps;
                    double t0[2], tN[6][2];
                    evalResiduals(x0 - eps, t0, tN);
                    forEach([&](double& d, int s, int e) { d -= (s < 0 ? t0[e] : tN[s][e]); });
                } else
                    forEach([&](double& d, int s, int e) { d -= (s < 0 ? orig0[e] : origN[s][e]); });
                forEach([&](double& d, int, int) { d /= delta; });
            }
            // scatter (:324-333): A[I][I][eq][pv] += d0[eq]; A[J][I][eq][pv] += dN[j][eq]
            const int kd = findEntry(I, I);
            for (int e = 0; e < b; ++e) jac[((size_t)kd * b + e) * b + pv] += d0[e];
            for (int side = 0; side < 2 * dim; ++side) {
                if (nbIdx[side] < 0) continue;
                const int k = findEntry(nbIdx[side], I);
                for (int e = 0; e < b; ++e) jac[((size_t)k * b + e) * b + pv] += dN[side][e];
            }
        }
    }
};

// =================================================================================================
// dune-istl 2.10 restatement [DUNE-ext]
// =================================================================================================
namespace {

// FieldMatrix<double,b,b>::invert() [DUNE-ext, dune-common densematrix.hh]: n==1: 1/a; n==2: explicit
// (detinv = 1/(a00*a11 - a01*a10); swap diag, negate off-diag, scale).
bool invertBlock(double* A, int b)
{
    if (b == 1) {
        if (A[0] == 0.0) return false;
        A[0] = 1.0 / A[0];
        return true;
    }
    double detinv = A[0] * A[3] - A[1] * A[2];
    if (detinv == 0.0 || !(detinv == detinv)) return false;
    detinv = 1.0 / detinv;
    const double temp = A[0];
    A[0] = A[3] * detinv;
    A[1] = -A[1] * detinv;
    A[2] = -A[2] * detinv;
    A[3] = temp * detinv;
    return true;
}
// C = A*B (rightmultiply semantic: A <- A*B) with Dune's loop order: C[i][j] = sum_k A[i][k]*B[k][j], sum from 0
inline void rightMultiply(double* A, const double* B, int b)
{
    double C[4];
    for (int i = 0; i < b; ++i)
        for (int j = 0; j < b; ++j) {
            double s = 0.0;
            for (int k = 0; k < b; ++k) s += A[i * b + k] * B[k * b + j];
            C[i * b + j] = s;
        }
    for (int i = 0; i < b * b; ++i) A[i] = C[i];
}
// y -= A x  (FieldMatrix::mmv)
__attribute__((always_inline)) inline void mmv(const double* A, const double* x, double* y, int b)
{
    for (int i = 0; i < b; ++i)
        for (int j = 0; j < b; ++j) y[i] -= A[i * b + j] * x[j];
}
// y += A x (umv)
__attribute__((always_inline)) inline void umv(const double* A, const double* x, double* y, int b)
{
    for (int i = 0; i < b; ++i)
        for (int j = 0; j < b; ++j) y[i] += A[i * b + j] * x[j];
}

// ILU::blockILU0Decomposition (dune-istl ilu.hh) [DUNE-ext]
int ilu0Factor(int n, int b, const int* rowptr, const int* colidx, double* A)
{
    const int bb = b * b;
    for (int i = 0; i < n; ++i) {
        int kdiag = -1;
        for (int kij = rowptr[i]; kij < rowptr[i + 1]; ++kij) {
            const int j = colidx[kij];
            if (j >= i) { if (j == i) kdiag = kij; break; }
            // find diagonal of row j
            int kjj = -1;
            for (int k = rowptr[j]; k < rowptr[j + 1]; ++k) if (colidx[k] == j) { kjj = k; break; }
            rightMultiply(A + (size_t)kij * bb, A + (size_t)kjj * bb, b);      // A_ij <- A_ij * A_jj^{-1}
            // for all k>j present in both rows: A_ik -= A_ij * A_jk
            int ik = kij + 1, jk = kjj + 1;
            const int iend = rowptr[i + 1], jend = rowptr[j + 1];
            while (ik < iend && jk < jend) {
                if (colidx[ik] == colidx[jk]) {
                    // FieldMatrix mmm: C -= A*B
                    double* C = A + (size_t)ik * bb;
                    const double* L = A + (size_t)kij * bb;
                    const double* U = A + (size_t)jk * bb;
                    for (int r = 0; r < b; ++r)
                        for (int c = 0; c < b; ++c)
                            for (int k = 0; k < b; ++k) C[r * b + c] -= L[r * b + k] * U[k * b + c];
                    ++ik; ++jk;
                } else if (colidx[ik] < colidx[jk]) ++ik;
                else ++jk;
            }
        }
        if (kdiag < 0) return 1;
        if (!invertBlock(A + (size_t)kdiag * bb, b)) return 1;
    }
    return 0;
}

// ILU::blockILUBacksolve [DUNE-ext]
template <int b>
void ilu0ApplyT(int n, const int* rowptr, const int* colidx, const double* A, double* v, const double* d)
{
    constexpr int bb = b * b;
    for (int i = 0; i < n; ++i) {
        double rhs[2];
        for (int e = 0; e < b; ++e) rhs[e] = d[(size_t)i * b + e];
        for (int k = rowptr[i]; k < rowptr[i + 1] && colidx[k] < i; ++k) mmv(A + (size_t)k * bb, v + (size_t)colidx[k] * b, rhs, b);
        for (int e = 0; e < b; ++e) v[(size_t)i * b + e] = rhs[e];
    }
    for (int i = n - 1; i >= 0; --i) {
        double rhs[2];
        for (int e = 0; e < b; ++e) rhs[e] = v[(size_t)i * b + e];
        int kd = -1;
        for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) {
            if (colidx[k] == i) kd = k;
            else if (colidx[k] > i) mmv(A + (size_t)k * bb, v + (size_t)colidx[k] * b, rhs, b);
        }
        double out[2] = {0, 0};
        umv(A + (size_t)kd * bb, rhs, out, b);      // v_i = A_ii^{-1} * rhs (FieldMatrix::mv, sum from 0)
        for (int e = 0; e < b; ++e) v[(size_t)i * b + e] = out[e];
    }
}

void ilu0Apply(int n, int b, const int* rowptr, const int* colidx, const double* A, double* v, const double* d)
{
    if (b == 1) ilu0ApplyT<1>(n, rowptr, colidx, A, v, d);
    else ilu0ApplyT<2>(n, rowptr, colidx, A, v, d);
}

// BCRSMatrix::mv: y = A x, per row sum from 0 in column order
template <int b>
void spmvT(int n, const int* rowptr, const int* colidx, const double* A, const double* x, double* y)
{
    constexpr int bb = b * b;
    for (int i = 0; i < n; ++i) {
        double acc[2] = {0, 0};
        for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) umv(A + (size_t)k * bb, x + (size_t)colidx[k] * b, acc, b);
        for (int e = 0; e < b; ++e) y[(size_t)i * b + e] = acc[e];
    }
}
void spmv(int n, int b, const int* rowptr, const int* colidx, const double* A, const double* x, double* y)
{
    if (b == 1) spmvT<1>(n, rowptr, colidx, A, x, y);
    else spmvT<2>(n, rowptr, colidx, A, x, y);
}
double dot(size_t n, const double* a, const double* c)
{
    double s = 0.0;
    for (size_t i = 0; i < n; ++i) s += a[i] * c[i];
    return s;
}

// Dune::BiCGSTABSolver::apply (dune-istl solvers.hh) [DUNE-ext], SURVEY Appendix A.
// `prec(v,d)`: v = M^-1 d;  `op(x,y)`: y = A x;  `sp(a,b)`: scalar product.
template <class Op, class Prec, class Dot>
int bicgstab(size_t N, Op&& op, Prec&& prec, Dot&& sp, double* x, const double* rhs, double reduction, int maxit,
             int* iterations, double* achieved)
{
    std::vector<double> r(rhs, rhs + N), rt(N), p(N, 0.0), v(N, 0.0), t(N), y(N), tmp(N);
    // r = b - A x
    op(x, tmp.data());
    for (size_t i = 0; i < N; ++i) r[i] -= tmp[i];
    rt = r;
    double norm0 = std::sqrt(sp(r.data(), r.data()));
    double norm = norm0;
    auto converged = [&](double nrm) { return nrm < reduction * norm0 || nrm < 1e-30; };
    *iterations = 0;
    *achieved = 1.0;
    if (!(norm0 == norm0) || std::isinf(norm0)) return 3;
    if (converged(norm0)) { *achieved = (norm0 > 0 ? norm / norm0 : 0.0); return 0; }
    double rho = 1, alpha = 1, omega = 1, rho_new, h, beta;
    const double EPSILON = 1e-80;
    double it;
    int status = 1;
    for (it = 0.5; it < maxit; it += .5) {
        rho_new = sp(rt.data(), r.data());
        if (std::fabs(rho) <= EPSILON || std::fabs(omega) <= EPSILON) { status = 2; break; }
        if (it < 1)
            p = r;
        else {
            beta = (rho_new / rho) * (alpha / omega);
            for (size_t i = 0; i < N; ++i) p[i] += -omega * v[i];   // p.axpy(-omega,v)
            for (size_t i = 0; i < N; ++i) p[i] *= beta;
            for (size_t i = 0; i < N; ++i) p[i] += r[i];
        }
        std::fill(y.begin(), y.end(), 0.0);
        prec(y.data(), p.data());
        op(y.data(), v.data());
        h = sp(rt.data(), v.data());
        if (std::fabs(h) < EPSILON) { status = 2; break; }
        alpha = rho_new / h;
        for (size_t i = 0; i < N; ++i) x[i] += alpha * y[i];
        for (size_t i = 0; i < N; ++i) r[i] += -alpha * v[i];
        norm = std::sqrt(sp(r.data(), r.data()));
        if (!(norm == norm) || std::isinf(norm)) { status = 3; break; }
        if (converged(norm)) { status = 0; break; }
        it += .5;
        std::fill(y.begin(), y.end(), 0.0);
        prec(y.data(), r.data());
        op(y.data(), t.data());
        omega = sp(t.data(), r.data()) / sp(t.data(), t.data());
        for (size_t i = 0; i < N; ++i) x[i] += omega * y[i];
        for (size_t i = 0; i < N; ++i) r[i] += -omega * t[i];
        rho = rho_new;
        norm = std::sqrt(sp(r.data(), r.data()));
        if (!(norm == norm) || std::isinf(norm)) { status = 3; break; }
        if (converged(norm)) { status = 0; break; }
    }
    *iterations = (int)std::ceil(std::min(it, (double)maxit));
    *achieved = norm0 > 0 ? norm / norm0 : 0.0;
    return status;
}

// Dune::SeqSSOR(n=1, w) [DUNE-ext, restated from dune-istl preconditioners.hh / gsetc.hh]: one bsorf + one bsorb per apply.
// Per row (forward: ascending rows, backward: descending), columns ascending INCLUDING the diagonal:
//   rhs = d_i - sum_j A_ij x_j (newest values) ; v = A_ii^-1 rhs (FieldMatrix::solve: 1x1 division, 2x2 closed form) ; x_i += w v
// Restated for w = 1 (LinearSolver.PreconditionerRelaxation default), where the two nested relaxation factors of gsetc.hh coincide.
inline void blockSolve(const double* A, const double* rhs, double* x, int b)
{
    if (b == 1) { x[0] = rhs[0] / A[0]; return; }
    double detinv = A[0] * A[3] - A[1] * A[2];
    detinv = 1 / detinv;
    x[0] = detinv * (A[3] * rhs[0] - A[1] * rhs[1]);
    x[1] = detinv * (A[0] * rhs[1] - A[2] * rhs[0]);
}
void ssorApply(int n, int b, const int* rowptr, const int* colidx, const double* A, double* x, const double* d)
{
    const int bb = b * b;
    auto row = [&](int i) {
        double rhs[2], v[2];
        for (int e = 0; e < b; ++e) rhs[e] = d[(size_t)i * b + e];
        int kd = -1;
        for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) {
            if (colidx[k] == i) kd = k;
            mmv(A + (size_t)k * bb, x + (size_t)colidx[k] * b, rhs, b);
        }
        blockSolve(A + (size_t)kd * bb, rhs, v, b);
        for (int e = 0; e < b; ++e) x[(size_t)i * b + e] += 1.0 * v[e];
    };
    for (int i = 0; i < n; ++i) row(i);
    for (int i = n - 1; i >= 0; --i) row(i);
}

// Dumux::ParMTJac / ParMTSOR / ParMTSSOR (dumux/linear/preconditioners.hh:330-400, 442-620): DuMux's own multi-threaded smoothers.
// Colours: computeColorsForMatrixSweep_ (:408-440) -- rows in index order take the smallest colour none of their matrix
// neighbours has (Detail::smallestAvailableColor, dumux/assembly/coloring.hh:195-215; the row's own, still unset colour -1 is
// in the list); the 7-point stencil in lexicographic order gets the two checkerboard colours.  Sweeps: parallelBlockSOR_
// (:442-476) colour by colour (backward: colours descending), per row rhs = b_i - sum_j A_ij update_j over ALL columns in
// ascending order (diagonal included, newest values), v = A_ii^-1 rhs (bsorf/bsorb at block level 0: FieldMatrix::solve),
// update_i += w v.  ParMTJac::apply (:364-393): the same row formula against the PREVIOUS iterate, all rows at once.
void parmtColors(int n, const int* rowptr, const int* colidx, std::vector<int>& colors, int& ncolors)
{
    colors.assign(n, -1);
    ncolors = 0;
    std::vector<int> nb;
    std::vector<char> used;
    for (int i = 0; i < n; ++i) {
        nb.clear();
        for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) nb.push_back(colors[colidx[k]]);
        const int m = (int)nb.size();
        used.assign(m, 0);
        for (int q = 0; q < m; ++q)
            if (nb[q] >= 0 && nb[q] < m) used[nb[q]] = 1;
        int c = m;
        for (int q = 0; q < m; ++q)
            if (!used[q]) { c = q; break; }
        colors[i] = c;
        ncolors = std::max(ncolors, c + 1);
    }
}
// kind 0: ParMTJac, 1: ParMTSOR (forward sweeps), 2: ParMTSSOR (forward + backward per iteration)
void parmtApply(int kind, int n, int b, const int* rowptr, const int* colidx, const double* A, double* update, const double* defect,
                int iterations, double w)
{
    const int bb = b * b;
    auto row = [&](int i, const double* xin) {
        double rhs[2], v[2];
        for (int e = 0; e < b; ++e) rhs[e] = defect[(size_t)i * b + e];
        int kd = -1;
        for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) {
            if (colidx[k] == i) kd = k;
            mmv(A + (size_t)k * bb, xin + (size_t)colidx[k] * b, rhs, b);
        }
        blockSolve(A + (size_t)kd * bb, rhs, v, b);
        for (int e = 0; e < b; ++e) update[(size_t)i * b + e] += w * v[e];
    };
    if (kind == 0) {
        std::vector<double> xOld(update, update + (size_t)n * b);
        for (int it = 0; it < iterations; ++it) {
            if (it > 0) std::copy(update, update + (size_t)n * b, xOld.begin());
            for (int i = 0; i < n; ++i) row(i, xOld.data());
        }
        return;
    }
    std::vector<int> colors;
    int nc = 0;
    parmtColors(n, rowptr, colidx, colors, nc);
    auto sweep = [&](bool forward) {
        for (int cc = 0; cc < nc; ++cc) {
            const int c = forward ? cc : nc - 1 - cc;
            for (int i = 0; i < n; ++i)
                if (colors[i] == c) row(i, update);
        }
    };
    for (int it = 0; it < iterations; ++it) {
        sweep(true);
        if (kind == 2) sweep(false);
    }
}

// Dune::CGSolver::apply (dune-istl solvers.hh) [DUNE-ext]: preconditioned conjugate gradients, the defect norm is that of b - A x
template <class Op, class Prec, class Dot>
int cgSolve(size_t N, Op&& op, Prec&& prec, Dot&& sp, double* x, const double* rhs, double reduction, int maxit, int* iterations,
            double* achieved)
{
    std::vector<double> b(rhs, rhs + N), p(N, 0.0), q(N, 0.0);
    op(x, q.data());
    for (size_t i = 0; i < N; ++i) b[i] -= q[i];                 // applyscaleadd(-1, x, b)
    double def = std::sqrt(sp(b.data(), b.data()));
    const double def0 = def;
    *iterations = 0;
    *achieved = 1.0;
    if (!(def0 == def0) || std::isinf(def0)) return 3;
    auto conv = [&](double nrm) { return nrm < reduction * def0 || nrm < 1e-30; };
    if (conv(def0)) { *achieved = def0 > 0 ? 1.0 : 0.0; return 0; }
    std::fill(p.begin(), p.end(), 0.0);
    prec(p.data(), b.data());
    double rholast = sp(p.data(), b.data()), rho, lambda, alpha, beta;
    int status = 1, i = 1;
    for (; i <= maxit; ++i) {
        op(p.data(), q.data());
        alpha = sp(p.data(), q.data());
        lambda = rholast / alpha;
        for (size_t k = 0; k < N; ++k) x[k] += lambda * p[k];
        for (size_t k = 0; k < N; ++k) b[k] += -lambda * q[k];
        def = std::sqrt(sp(b.data(), b.data()));
        *iterations = i;
        if (!(def == def) || std::isinf(def)) { status = 3; break; }
        if (conv(def)) { status = 0; break; }
        std::fill(q.begin(), q.end(), 0.0);
        prec(q.data(), b.data());
        rho = sp(q.data(), b.data());
        beta = rho / rholast;
        for (size_t k = 0; k < N; ++k) p[k] *= beta;
        for (size_t k = 0; k < N; ++k) p[k] += q[k];
        rholast = rho;
    }
    if (i > maxit) *iterations = maxit;
    *achieved = def0 > 0 ? def / def0 : 0.0;
    return status;
}

// Dune::RestartedGMResSolver::apply (dune-istl solvers.hh) [DUNE-ext, restated from the published algorithm]: LEFT-preconditioned
// GMRes(m) -- the monitored norm is that of the PRECONDITIONED defect M^-1(b - A x) --, Arnoldi with modified Gram-Schmidt,
// Givens rotations (generatePlaneRotation / applyPlaneRotation), update() by back substitution with the correction added on
// the fly from the last basis vector to the first, restart from the recomputed defect.  Iterations count operator applications.
inline void gmresGenerateRotation(double dx, double dy, double& cs, double& sn)
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
inline void gmresApplyRotation(double& dx, double& dy, double cs, double sn)
{
    const double temp = cs * dx + sn * dy;
    dy = -sn * dx + cs * dy;
    dx = temp;
}
template <class Op, class Prec, class Dot>
int restartedGmres(size_t N, Op&& op, Prec&& prec, Dot&& sp, double* x, const double* rhs, double reduction, int maxit, int restart,
                   int* iterations, double* achieved)
{
    const int m = restart;
    const double EPSILON = 1e-80;
    std::vector<double> s(m + 1), sn(m), cs(m), b(rhs, rhs + N), w(N), tmp(N);
    std::vector<std::vector<double>> H(m + 1, std::vector<double>(m + 1, 0.0)), v(m + 1, std::vector<double>(N, 0.0));
    auto defect = [&]() {                                   // b = rhs - A x ; v[0] = M^-1 b
        op(x, tmp.data());
        for (size_t i = 0; i < N; ++i) b[i] = rhs[i] - tmp[i];
        std::fill(v[0].begin(), v[0].end(), 0.0);
        prec(v[0].data(), b.data());
        return std::sqrt(sp(v[0].data(), v[0].data()));
    };
    double norm = defect();
    const double norm0 = norm;
    *iterations = 0;
    *achieved = 1.0;
    if (!(norm0 == norm0) || std::isinf(norm0)) return 3;
    auto conv = [&](double nrm) { return nrm < reduction * norm0 || nrm < 1e-30; };
    if (conv(norm0)) { *achieved = norm0 > 0 ? 1.0 : 0.0; return 0; }
    int j = 1;
    bool converged = false;
    int status = 1;
    while (j <= maxit && !converged) {
        int i = 0;
        {
            const double f = (norm == 0.0) ? 0.0 : 1.0 / norm;
            for (size_t q = 0; q < N; ++q) v[0][q] *= f;
        }
        s[0] = norm;
        for (i = 1; i < m + 1; ++i) s[i] = 0.0;
        for (i = 0; i < m && j <= maxit && !converged; ++i, ++j) {
            std::fill(w.begin(), w.end(), 0.0);
            op(v[i].data(), v[i + 1].data());
            prec(w.data(), v[i + 1].data());
            for (int k = 0; k < i + 1; ++k) {
                H[k][i] = sp(v[k].data(), w.data());
                const double mh = -H[k][i];
                for (size_t q = 0; q < N; ++q) w[q] += mh * v[k][q];
            }
            H[i + 1][i] = std::sqrt(sp(w.data(), w.data()));
            if (!(H[i + 1][i] == H[i + 1][i]) || std::isinf(H[i + 1][i])) { *iterations = j; return 3; }
            if (std::fabs(H[i + 1][i]) < EPSILON) { *iterations = j; *achieved = norm / norm0; return 2; }
            {
                const double f = (norm == 0.0) ? 0.0 : 1.0 / H[i + 1][i];
                for (size_t q = 0; q < N; ++q) v[i + 1][q] = w[q] * f;
            }
            for (int k = 0; k < i; ++k) gmresApplyRotation(H[k][i], H[k + 1][i], cs[k], sn[k]);
            gmresGenerateRotation(H[i][i], H[i + 1][i], cs[i], sn[i]);
            gmresApplyRotation(H[i][i], H[i + 1][i], cs[i], sn[i]);
            gmresApplyRotation(s[i], s[i + 1], cs[i], sn[i]);
            norm = std::fabs(s[i + 1]);
            *iterations = j;
            if (conv(norm)) { converged = true; status = 0; }
        }
        // update(w, i, H, s, v): back substitution, x += sum_a y_a v_a accumulated from a = i-1 down to 0
        std::fill(w.begin(), w.end(), 0.0);
        {
            std::vector<double> y(s);
            for (int a = i - 1; a >= 0; --a) {
                double r = s[a];
                for (int c = a + 1; c < i; ++c) r -= H[a][c] * y[c];
                y[a] = (r == 0.0) ? 0.0 : r / H[a][a];
                for (size_t q = 0; q < N; ++q) w[q] += y[a] * v[a][q];
            }
        }
        for (size_t q = 0; q < N; ++q) x[q] += w[q];
        if (!converged && j < maxit) norm = defect();
    }
    *achieved = norm0 > 0 ? norm / norm0 : 0.0;
    return status;
}

double nowSec()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

} // namespace

// =================================================================================================
// C API
// =================================================================================================
extern "C" {

void orc_default_options(orc_options* o)
{
    o->enable_gravity = 1;
    o->gravity = 9.81;
    o->upwind_weight = 1.0;
    o->fd_method = 1;
    o->base_eps = 1e-10;
    o->privar_magnitude[0] = o->privar_magnitude[1] = -1.0;
    o->stationary = 0;
    o->dt = 1.0;
    o->extrusion = 1.0;
    o->use_std_pow = 0;
    o->num_threads = 0;
}

static orc_problem* createCommon(int model, int dim, const int* cells)
{
    orc_problem* p = new orc_problem;
    p->model = model;
    p->b = (model == ORC_MODEL_2P) ? 2 : 1;
    p->dim = dim;
    for (int a = 0; a < 3; ++a) p->nc[a] = (a < dim) ? cells[a] : 1;
    p->n = p->nc[0] * p->nc[1] * p->nc[2];
    orc_default_options(&p->opt);
    p->K.assign(p->n, 1e-10);
    p->phi.assign(p->n, 0.4);
    p->region.assign(p->n, 0);
    p->laws.resize(1);
    return p;
}

orc_problem* orc_create(int model, int dim, const int* cells, const double* lower, const double* upper)
{
    orc_problem* p = createCommon(model, dim, cells);
    // YaspGrid EquidistantOffsetCoordinates: x_i = origin + i*h, h = (upper-lower)/cells [DUNE-ext]
    for (int a = 0; a < 3; ++a) {
        p->xn[a].resize(p->nc[a] + 1);
        if (a < dim) {
            const double h = (upper[a] - lower[a]) / cells[a];
            for (int i = 0; i <= p->nc[a]; ++i) p->xn[a][i] = lower[a] + i * h;
        } else { p->xn[a][0] = 0.0; p->xn[a][1] = 1.0; }
    }
    p->buildPattern();
    return p;
}

orc_problem* orc_create_tensor(int model, int dim, const int* cells, const double* x, const double* y, const double* z)
{
    orc_problem* p = createCommon(model, dim, cells);
    const double* src[3] = {x, y, z};
    for (int a = 0; a < 3; ++a) {
        p->xn[a].resize(p->nc[a] + 1);
        if (a < dim) for (int i = 0; i <= p->nc[a]; ++i) p->xn[a][i] = src[a][i];
        else { p->xn[a][0] = 0.0; p->xn[a][1] = 1.0; }
    }
    p->buildPattern();
    return p;
}

void orc_destroy(orc_problem* p) { delete p; }
int orc_num_cells(const orc_problem* p) { return p->n; }
int orc_num_eq(const orc_problem* p) { return p->b; }

void orc_set_options(orc_problem* p, const orc_options* o)
{
    p->opt = *o;
    for (auto& l : p->laws) { l.pw = o->use_std_pow ? std_pow_ : det_pow_; l.init(); }
}
void orc_set_cell_fields(orc_problem* p, const double* K, const double* phi, const int* region)
{
    if (K) p->K.assign(K, K + p->n);
    if (phi) p->phi.assign(phi, phi + p->n);
    if (region) p->region.assign(region, region + p->n);
}
// diagonal permeability tensor (SpatialParams::permeability returning a FieldMatrix without off-diagonal entries): one array per
// grid axis, all null = back to the scalar field
void orc_set_permeability_diagonal(orc_problem* p, const double* kx, const double* ky, const double* kz)
{
    const double* k[3] = {kx, ky, kz};
    for (int a = 0; a < 3; ++a) {
        p->Kdiag[a].clear();
        if (k[a] && a < p->dim) p->Kdiag[a].assign(k[a], k[a] + p->n);
    }
    if (k[p->dim - 1]) p->K.assign(k[p->dim - 1], k[p->dim - 1] + p->n);
}
void orc_set_source(orc_problem* p, const double* q) { p->q.assign(q, q + (size_t)p->n * p->b); }

void orc_set_material(orc_problem* p, int region, int law, const double* params, double swr, double snr,
                      int regularize, const double* reg)
{
    if ((int)p->laws.size() <= region) p->laws.resize(region + 1);
    Law& l = p->laws[region];
    l.kind = law;
    l.swr = swr; l.snr = snr;
    l.reg = regularize != 0;
    l.pw = p->opt.use_std_pow ? std_pow_ : det_pow_;
    if (law == ORC_LAW_BROOKSCOREY) {
        l.pe = params[0]; l.lambda = params[1];
        l.pcLowSwe = reg ? reg[0] : 0.01;
    } else {
        l.alpha = params[0]; l.n = params[1]; l.m = 1.0 - 1.0 / l.n;   // vangenuchten.hh Params::setN
        l.l = params[2];
        if (reg) { l.pcLowSwe = reg[0]; l.pcHighSwe = reg[1]; l.krnLowSwe = reg[2]; l.krwHighSwe = reg[3]; }
    }
    l.init();
}
void orc_set_wetting_phase(orc_problem* p, int region, int phase)
{
    if ((int)p->laws.size() <= region) p->laws.resize(region + 1);
    p->laws[region].wetting = phase ? 1 : 0;
}
void orc_set_fluids(orc_problem* p, const double* rho, const double* mu)
{
    p->fluids.tabulated = false;
    for (int i = 0; i < (p->model == ORC_MODEL_2P ? 2 : 1); ++i) { p->fluids.rho[i] = rho[i]; p->fluids.mu[i] = mu[i]; }
}
void orc_set_fluid_table(orc_problem* p, int nT, int nP, double Tmin, double Tmax, const double* pmin, const double* pmax,
                         const double* rho, const double* mu, double temperature);

int orc_side_faces(const orc_problem* p, int side) { return p->sideFaces(side); }
void orc_set_boundary(orc_problem* p, int side, const int* type, const double* values)
{
    const int nf = p->sideFaces(side);
    p->bcType[side].assign(type, type + nf);
    p->bcVal[side].assign(values, values + (size_t)nf * p->b);
}
void orc_side_face_center(const orc_problem* p, int side, int f, double* out)
{
    const int a = side / 2;
    int c[3] = {0, 0, 0};
    if (a == 0) { c[1] = f % p->nc[1]; c[2] = f / p->nc[1]; c[0] = (side & 1) ? p->nc[0] - 1 : 0; }
    else if (a == 1) { c[0] = f % p->nc[0]; c[2] = f / p->nc[0]; c[1] = (side & 1) ? p->nc[1] - 1 : 0; }
    else { c[0] = f % p->nc[0]; c[1] = f / p->nc[0]; c[2] = (side & 1) ? p->nc[2] - 1 : 0; }
    for (int d = 0; d < 3; ++d) out[d] = (d < p->dim) ? p->center(d, c[d]) : 0.0;
    out[a] = (side & 1) ? p->xn[a][c[a] + 1] : p->xn[a][c[a]];
}
void orc_cell_center(const orc_problem* p, int cell, double* out)
{
    int c[3];
    p->ijk(cell, c);
    for (int d = 0; d < 3; ++d) out[d] = (d < p->dim) ? p->center(d, c[d]) : 0.0;
}

int orc_pattern_nnz(const orc_problem* p) { return (int)p->colidx.size(); }
void orc_pattern(const orc_problem* p, int* rowptr, int* colidx)
{
    std::memcpy(rowptr, p->rowptr.data(), sizeof(int) * (p->n + 1));
    std::memcpy(colidx, p->colidx.data(), sizeof(int) * p->colidx.size());
}

// global assembly: dumux/assembly/fvassembler.hh:179-207 (reset J and r, loop all elements :462-510)
void orc_assemble(orc_problem* p, const double* cur, const double* prev, double* residual, double* jac)
{
    const size_t nnz = p->colidx.size();
    if (jac) std::memset(jac, 0, sizeof(double) * nnz * p->b * p->b);
    if (residual) std::memset(residual, 0, sizeof(double) * (size_t)p->n * p->b);
#ifdef _OPENMP
    const int nt = p->opt.num_threads > 0 ? p->opt.num_threads : omp_get_max_threads();
#pragma omp parallel for schedule(static) num_threads(nt)
#endif
    for (int I = 0; I < p->n; ++I) {
        if (p->opt.fd_method == ORC_DIFF_ANALYTIC && p->model == ORC_MODEL_2P) p->assembleElementAnalytic2p(I, cur, prev, residual, jac);
        else if (p->opt.fd_method == ORC_DIFF_ANALYTIC) p->assembleElementAnalytic1p(I, cur, prev, residual, jac);
        else p->assembleElement(I, cur, prev, residual, jac);
    }
}

void orc_volvars(orc_problem* p, const double* cur, double* out)
{
    for (int I = 0; I < p->n; ++I) {
        VolVars v;
        p->updateVolVars(v, cur + (size_t)I * p->b, I);
        double* o = out + (size_t)I * 12;
        o[0] = v.S[0]; o[1] = v.S[1]; o[2] = v.p[0]; o[3] = v.p[1]; o[4] = v.rho[0]; o[5] = v.rho[1];
        o[6] = v.mob[0]; o[7] = v.mob[1]; o[8] = v.pc; o[9] = v.porosity; o[10] = v.K; o[11] = 0.0;
    }
}

int orc_ilu0_factor(int n, int b, const int* rowptr, const int* colidx, const double* values, double* ilu)
{
    std::memcpy(ilu, values, sizeof(double) * (size_t)rowptr[n] * b * b);
    return ilu0Factor(n, b, rowptr, colidx, ilu);
}
void orc_ilu0_apply(int n, int b, const int* rowptr, const int* colidx, const double* ilu, double* v, const double* d)
{
    ilu0Apply(n, b, rowptr, colidx, ilu, v, d);
}
void orc_spmv(int n, int b, const int* rowptr, const int* colidx, const double* values, const double* x, double* y)
{
    spmv(n, b, rowptr, colidx, values, x, y);
}
double orc_norm2(int n, const double* v) { return std::sqrt(dot((size_t)n, v, v)); }
double orc_dot(int n, const double* a, const double* b) { return dot((size_t)n, a, b); }

// dumux/nonlinear/newtonsolver.hh:111-129
double orc_max_relative_shift(int n, const double* u1, const double* u2)
{
    double s = 0.0;
    for (int i = 0; i < n; ++i) {
        const double v = std::fabs(u1[i] - u2[i]) / std::max(1.0, std::fabs(u1[i] + u2[i]) * 0.5);
        s = std::max(s, v);
    }
    return s;
}

// IstlIterativeLinearSolver::solve -> SeqILU(n=0,w=1) + BiCGSTABSolver, built fresh per call
// (dumux/linear/istlsolvers.hh:457-464,535-568,636-642)
int orc_ilu0_bicgstab(int n, int b, const int* rowptr, const int* colidx, const double* values,
                      double* x, const double* rhs, double reduction, int maxit, int* iterations, double* achieved)
{
    const size_t N = (size_t)n * b;
    std::vector<double> ilu((size_t)rowptr[n] * b * b);
    std::memcpy(ilu.data(), values, sizeof(double) * ilu.size());
    if (ilu0Factor(n, b, rowptr, colidx, ilu.data())) return 2;
    return bicgstab(
        N, [&](const double* in, double* out) { spmv(n, b, rowptr, colidx, values, in, out); },
        [&](double* v, const double* d) { ilu0Apply(n, b, rowptr, colidx, ilu.data(), v, d); },
        [&](const double* a, const double* c) { return dot(N, a, c); }, x, rhs, reduction, maxit, iterations, achieved);
}

// The same solver on P MPI ranks, the way DuMux runs it on a decomposed grid (linear/linearsolvertraits.hh:79-91 +
// istlsolvers.hh:550-566): dune-istl's OverlappingSchwarzOperator (local mv, then project = zero the non-owner entries),
// OverlappingSchwarzScalarProduct (owner-masked dot + global sum) and BlockPreconditioner (local SeqILU, then copyOwnerToAll)
// around the SAME BiCGSTABSolver::apply.  The ranks are OpenMP threads of this process (no MPI in this image): every thread runs
// bicgstab() on its local box, collectives are a shared slot per rank and a barrier, the global sum adds the per-rank values in
// rank order.  copyOwnerToAll is a copy list per rank: local cell `dst` takes the value of cell `src_idx` of rank `src_rank`
// (its owner).  Used by bench.py's CPU arms: one native thread per rank, no interpreter in the loop.
// x[r] / rhs[r]: local vectors of rank r (rhs may hold the incomplete residual rows of the overlap cells: they are projected out).
int orc_schwarz_ilu0_bicgstab(int nranks, int b, const int* n, const int* const* rowptr, const int* const* colidx,
                              const double* const* values, double* const* x, const double* const* rhs,
                              const unsigned char* const* owner, const int* ncopy, const int* const* copy_dst,
                              const int* const* copy_src_rank, const int* const* copy_src_idx, double reduction, int maxit,
                              int* iterations, double* achieved, double* seconds)
{
    std::vector<double> red(nranks, 0.0);
    std::vector<const double*> vec(nranks, nullptr);
    std::vector<int> status(nranks, 0), its(nranks, 0), bad(nranks, 0);
    std::vector<double> ach(nranks, 1.0), secs(nranks, 0.0);
#pragma omp parallel num_threads(nranks)
    {
        const int r = omp_get_thread_num();
        const int nr = n[r];
        const size_t N = (size_t)nr * b;
        const unsigned char* own = owner[r];
        std::vector<double> ilu((size_t)rowptr[r][nr] * b * b);
        std::memcpy(ilu.data(), values[r], sizeof(double) * ilu.size());
        std::vector<double> rhsP(rhs[r], rhs[r] + N);
        for (int i = 0; i < nr; ++i)
            if (!own[i]) for (int e = 0; e < b; ++e) rhsP[(size_t)i * b + e] = 0.0;      // applyscaleadd(-1, x, r) projects r
#pragma omp barrier
        const double t0 = omp_get_wtime();
        bad[r] = ilu0Factor(nr, b, rowptr[r], colidx[r], ilu.data());
#pragma omp barrier
        int anyBad = 0;
        for (int q = 0; q < nranks; ++q) anyBad |= bad[q];
        if (anyBad) {
            status[r] = 2;
        } else {
            auto op = [&](const double* in, double* out) {
                spmv(nr, b, rowptr[r], colidx[r], values[r], in, out);
                for (int i = 0; i < nr; ++i)
                    if (!own[i]) for (int e = 0; e < b; ++e) out[(size_t)i * b + e] = 0.0;
            };
            auto prec = [&](double* v, const double* d) {
                ilu0Apply(nr, b, rowptr[r], colidx[r], ilu.data(), v, d);
                vec[r] = v;
#pragma omp barrier
                for (int k = 0; k < ncopy[r]; ++k) {
                    const double* src = vec[copy_src_rank[r][k]] + (size_t)copy_src_idx[r][k] * b;
                    double* dst = v + (size_t)copy_dst[r][k] * b;
                    for (int e = 0; e < b; ++e) dst[e] = src[e];
                }
#pragma omp barrier
            };
            auto sp = [&](const double* a, const double* c) {
                double s = 0.0;
                for (int i = 0; i < nr; ++i)
                    if (own[i]) for (int e = 0; e < b; ++e) s += a[(size_t)i * b + e] * c[(size_t)i * b + e];
                red[r] = s;
#pragma omp barrier
                double total = red[0];
                for (int q = 1; q < nranks; ++q) total = total + red[q];
#pragma omp barrier
                return total;
            };
            status[r] = bicgstab(N, op, prec, sp, x[r], rhsP.data(), reduction, maxit, &its[r], &ach[r]);
        }
        secs[r] = omp_get_wtime() - t0;
    }
    *iterations = its[0];
    *achieved = ach[0];
    double smax = 0.0;
    for (int q = 0; q < nranks; ++q) smax = std::max(smax, secs[q]);
    if (seconds) *seconds = smax;
    return status[0];
}

// ILURestartedGMResIstlSolver (dumux/linear/istlsolvers.hh:660-667): SeqILU(0) + Dune::RestartedGMResSolver,
// restart = LinearSolver.GMResRestart (default 10, linearsolverparameters.hh:115-116,138)
int orc_ilu0_gmres(int n, int b, const int* rowptr, const int* colidx, const double* values, double* x, const double* rhs,
                   double reduction, int maxit, int restart, int* iterations, double* achieved)
{
    const size_t N = (size_t)n * b;
    std::vector<double> ilu((size_t)rowptr[n] * b * b);
    std::memcpy(ilu.data(), values, sizeof(double) * ilu.size());
    if (ilu0Factor(n, b, rowptr, colidx, ilu.data())) return 2;
    return restartedGmres(
        N, [&](const double* in, double* out) { spmv(n, b, rowptr, colidx, values, in, out); },
        [&](double* v, const double* d) { ilu0Apply(n, b, rowptr, colidx, ilu.data(), v, d); },
        [&](const double* a, const double* c) { return dot(N, a, c); }, x, rhs, reduction, maxit, restart, iterations, achieved);
}
// SSORCGIstlSolver / SSORBiCGSTABIstlSolver (dumux/linear/istlsolvers.hh:686-714): Dune::SeqSSOR(1 iteration, w = 1) with
// Dune::CGSolver (kind 0) or Dune::BiCGSTABSolver (kind 1).  SSORCG is the linear solver of the reference's 1p incompressible test.
int orc_ssor_solve(int n, int b, const int* rowptr, const int* colidx, const double* values, double* x, const double* rhs, int krylov,
                   double reduction, int maxit, int* iterations, double* achieved)
{
    const size_t N = (size_t)n * b;
    auto op = [&](const double* in, double* out) { spmv(n, b, rowptr, colidx, values, in, out); };
    auto prec = [&](double* v, const double* d) { ssorApply(n, b, rowptr, colidx, values, v, d); };
    auto sp = [&](const double* a, const double* c) { return dot(N, a, c); };
    if (krylov == 0) return cgSolve(N, op, prec, sp, x, rhs, reduction, maxit, iterations, achieved);
    return bicgstab(N, op, prec, sp, x, rhs, reduction, maxit, iterations, achieved);
}
void orc_ssor_apply(int n, int b, const int* rowptr, const int* colidx, const double* values, double* v, const double* d)
{
    std::fill(v, v + (size_t)n * b, 0.0);
    ssorApply(n, b, rowptr, colidx, values, v, d);
}
// the "par_mt_jac" / "par_mt_sor" / "par_mt_ssor" preconditioners (preconditioners.hh:400,540,618) under CG (krylov 0) or BiCGSTAB
int orc_parmt_solve(int kind, int iterations, double relaxation, int n, int b, const int* rowptr, const int* colidx, const double* values,
                    double* x, const double* rhs, int krylov, double reduction, int maxit, int* its, double* achieved)
{
    const size_t N = (size_t)n * b;
    auto op = [&](const double* in, double* out) { spmv(n, b, rowptr, colidx, values, in, out); };
    auto prec = [&](double* v, const double* d) { parmtApply(kind, n, b, rowptr, colidx, values, v, d, iterations, relaxation); };
    auto sp = [&](const double* a, const double* c) { return dot(N, a, c); };
    if (krylov == 0) return cgSolve(N, op, prec, sp, x, rhs, reduction, maxit, its, achieved);
    return bicgstab(N, op, prec, sp, x, rhs, reduction, maxit, its, achieved);
}
void orc_parmt_apply(int kind, int iterations, double relaxation, int n, int b, const int* rowptr, const int* colidx, const double* values,
                     double* v, const double* d)
{
    std::fill(v, v + (size_t)n * b, 0.0);
    parmtApply(kind, n, b, rowptr, colidx, values, v, d, iterations, relaxation);
}
int orc_parmt_colors(int n, const int* rowptr, const int* colidx, int* colors)
{
    std::vector<int> c;
    int nc = 0;
    parmtColors(n, rowptr, colidx, c, nc);
    std::copy(c.begin(), c.end(), colors);
    return nc;
}
// ---- aggregation AMG on the structured hierarchy (oracle/amg_oracle.py; device: dumux_b200/csrc/amg.cu) ----
// Galerkin product P^T A P for box aggregates (2 cells per axis) and piecewise-constant prolongation on the 7-point pattern:
// coarse block (I,J) = sum of the fine blocks (i,j), i in I, j in J.  Canonical summation order: children of a coarse cell in
// lexicographic order (x fastest), per child its entries in column order, every coarse slot accumulated from 0 in that
// visiting order.
// Aggregates never cross a processor boundary (as in dune-istl's parallel AMG): along every axis the OWNED range
// [fown_lo, fown_hi) of the local fine box is cut into pairs starting at its first cell (a last single cell if the length is
// odd); the overlap cell below / above belongs to the neighbour's last / first aggregate, which is the coarse overlap cell
// cown_lo - 1 / cown_hi.  Single domain: the owned range is the box, i.e. aggregate = index >> 1.  Rows of coarse overlap
// cells are partial sums (only the children inside the local fine box), like the incomplete rows of fine overlap cells.
static inline int amg_agg1(int i, int flo, int fhi, int clo)
{
    if (i < flo) return clo - 1;
    if (i >= fhi) return clo + ((fhi - flo + 1) >> 1);
    return clo + ((i - flo) >> 1);
}
static inline void amg_children1(int I, int flo, int fhi, int clo, int& c0, int& c1)
{
    const int chi = clo + ((fhi - flo + 1) >> 1);
    if (I < clo) { c0 = flo - 1; c1 = flo; }
    else if (I >= chi) { c0 = fhi; c1 = fhi + 1; }
    else { c0 = flo + 2 * (I - clo); c1 = c0 + 2 < fhi ? c0 + 2 : fhi; }
}
void orc_amg_galerkin(int b, int dim, const int* fcells, const int* fown_lo, const int* fown_hi, const int* f_rowptr, const double* fA,
                      const int* ccells, const int* cown_lo, const int* c_rowptr, double* cA)
{
    const int bb = b * b;
    const int fx = fcells[0], fy = fcells[1], fz = fcells[2], cx = ccells[0], cy = ccells[1], cz = ccells[2];
    auto agg = [&](int a, int i) { return amg_agg1(i, fown_lo[a], fown_hi[a], cown_lo[a]); };
    for (int K = 0; K < cz; ++K)
        for (int J = 0; J < cy; ++J)
            for (int I = 0; I < cx; ++I) {
                const int Ic = I + cx * (J + cy * K);
                double acc[7][4] = {};
                auto add = [&](int slot, const double* blk) { for (int q = 0; q < bb; ++q) acc[slot][q] += blk[q]; };
                int i0, i1, j0, j1, k0, k1;
                amg_children1(I, fown_lo[0], fown_hi[0], cown_lo[0], i0, i1);
                amg_children1(J, fown_lo[1], fown_hi[1], cown_lo[1], j0, j1);
                amg_children1(K, fown_lo[2], fown_hi[2], cown_lo[2], k0, k1);
                for (int k = k0; k < k1; ++k)
                    for (int j = j0; j < j1; ++j)
                        for (int i = i0; i < i1; ++i) {
                            const size_t row = (size_t)i + (size_t)fx * (j + (size_t)fy * k);
                            const double* p = fA + (size_t)f_rowptr[row] * bb;
                            if (dim > 2 && k > 0) { add(agg(2, k - 1) == K ? 3 : 0, p); p += bb; }
                            if (dim > 1 && j > 0) { add(agg(1, j - 1) == J ? 3 : 1, p); p += bb; }
                            if (i > 0) { add(agg(0, i - 1) == I ? 3 : 2, p); p += bb; }
                            add(3, p); p += bb;
                            if (i + 1 < fx) { add(agg(0, i + 1) == I ? 3 : 4, p); p += bb; }
                            if (dim > 1 && j + 1 < fy) { add(agg(1, j + 1) == J ? 3 : 5, p); p += bb; }
                            if (dim > 2 && k + 1 < fz) { add(agg(2, k + 1) == K ? 3 : 6, p); p += bb; }
                        }
                double* out = cA + (size_t)c_rowptr[Ic] * bb;
                const bool ex[7] = {dim > 2 && K > 0, dim > 1 && J > 0, I > 0, true, I + 1 < cx, dim > 1 && J + 1 < cy, dim > 2 && K + 1 < cz};
                for (int s = 0; s < 7; ++s)
                    if (ex[s]) {
                        for (int q = 0; q < bb; ++q) out[q] = acc[s][q];
                        out += bb;
                    }
            }
}
// SeqSSOR (one forward + one backward block Gauss-Seidel sweep from zero) in factorised form M = (D + L) D^-1 (D + U), stored in
// the layout blockILUBacksolve consumes: L~_ij = A_ij A_jj^-1 (rightmultiply with the inverted diagonal block), A_ii^-1 on the
// diagonal, U_ij = A_ij.  Returns 1 if a diagonal block is singular.
int orc_ssor_factor(int n, int b, const int* rowptr, const int* colidx, const double* A, double* out)
{
    const int bb = b * b;
    int bad = 0;
    std::vector<double> dinv((size_t)n * bb);
    for (int i = 0; i < n; ++i) {
        int kd = -1;
        for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) if (colidx[k] == i) kd = k;
        if (kd < 0) return 1;
        for (int q = 0; q < bb; ++q) dinv[(size_t)i * bb + q] = A[(size_t)kd * bb + q];
        if (!invertBlock(&dinv[(size_t)i * bb], b)) bad = 1;
    }
    for (int i = 0; i < n; ++i)
        for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) {
            const int j = colidx[k];
            double* o = out + (size_t)k * bb;
            for (int q = 0; q < bb; ++q) o[q] = A[(size_t)k * bb + q];
            if (j == i) for (int q = 0; q < bb; ++q) o[q] = dinv[(size_t)i * bb + q];
            else if (j < i) rightMultiply(o, &dinv[(size_t)j * bb], b);
        }
    return bad;
}
// Mechanical dispersion (EnableCompositionalDispersion, flux/cctpfa/dispersionflux.hh:66-113): the dispersion tensor is given at
// the scvfs (Scheidegger: D = (aL - aT) v v^T / |v| + aT |v| I from the STATIONARY velocity field, scheidegger.hh:152-176), so it
// is a sampled array like the volume fluxes: for every cell and side the entry n.D.n, which is all a TPFA transmissibility of an
// axis-aligned face sees.  null = off.
void orc_set_tracer_dispersion(orc_problem* p, const double* disp)
{
    if (disp) p->tracerDisp.assign(disp, disp + (size_t)p->n * 2 * p->dim);
    else p->tracerDisp.clear();
}
void orc_set_tracer_diffusion(orc_problem* p, double D, double tortuosity)
{
    p->tracerD = D;
    p->tracerTau = tortuosity;
}
void orc_set_linear_solver(orc_problem* p, int kind, int restart)
{
    p->linearSolver = kind;
    p->gmresRestart = restart > 0 ? restart : 10;
}
// the solver selected with orc_set_linear_solver (what NewtonSolver::solveLinearSystem calls)
static int orcLinearSolve(orc_problem* p, const double* values, double* x, const double* rhs, double reduction, int maxit, int* iterations,
                          double* achieved)
{
    if (p->linearSolver == ORC_SOLVER_GMRES)
        return orc_ilu0_gmres(p->n, p->b, p->rowptr.data(), p->colidx.data(), values, x, rhs, reduction, maxit, p->gmresRestart, iterations,
                              achieved);
    return orc_ilu0_bicgstab(p->n, p->b, p->rowptr.data(), p->colidx.data(), values, x, rhs, reduction, maxit, iterations, achieved);
}

// Newton loop at fixed dt: dumux/nonlinear/newtonsolver.hh:976-1072 (solveImpl_), newtonProceed :428-446,
// newtonUpdate :543-557, shift :1138-1144, newtonConverged :657-666 (shift criterion only, defaults :1213-1247)
int orc_newton_solve(orc_problem* p, double* u, const double* prev, double lin_reduction, int lin_maxit,
                     double max_rel_shift, int min_steps, int max_steps, orc_newton_report* rep)
{
    const int n = p->n, b = p->b;
    const size_t N = (size_t)n * b;
    const size_t nnz = p->colidx.size();
    std::vector<double> J(nnz * b * b), r(N), delta(N), uLast(u, u + N);
    int numSteps = 0;
    double shift = 0.0, lastShift = 0.0;
    bool converged = false;
    std::memset(rep, 0, sizeof(*rep));
    auto proceed = [&]() {
        if (numSteps < min_steps) return true;
        else if (converged) return false;
        else if (numSteps >= max_steps) return shift * 4.0 < lastShift;
        return true;
    };
    while (proceed()) {
        lastShift = shift;
        if (numSteps > 0) uLast.assign(u, u + N);
        double t0 = nowSec();
        orc_assemble(p, u, prev, r.data(), J.data());
        double t1 = nowSec();
        for (size_t i = 0; i < N; ++i)
            if (!(r[i] == r[i]) || std::isinf(r[i])) return 3;
        std::fill(delta.begin(), delta.end(), 0.0);
        int its = 0;
        double red = 0;
        const int st = orcLinearSolve(p, J.data(), delta.data(), r.data(), lin_reduction, lin_maxit, &its, &red);
        double t2 = nowSec();
        if (numSteps < 64) rep->linear_iterations[numSteps] = its;
        rep->linear_iterations_total += its;
        if (st != 0) { rep->newton_iterations = numSteps; rep->converged = 0; return st; }
        for (size_t i = 0; i < N; ++i) u[i] = uLast[i] + (-1.0) * delta[i];     // uCurrentIter = uLastIter; axpy(-1, deltaU)
        shift = orc_max_relative_shift((int)N, u, uLast.data());
        double t3 = nowSec();
        rep->t_assemble += t1 - t0; rep->t_solve += t2 - t1; rep->t_update += t3 - t2;
        if (numSteps < 64) rep->shifts[numSteps] = shift;
        ++numSteps;
        converged = shift <= max_rel_shift;
    }
    rep->newton_iterations = numSteps;
    rep->converged = converged ? 1 : 0;
    rep->last_shift = shift;
    return converged ? 0 : 1;
}

void orc_default_newton_options(orc_newton_options* o)
{
    o->use_line_search = 0; o->line_search_min_relaxation = 0.125;
    o->enable_shift_criterion = 1; o->enable_residual_criterion = 0; o->enable_absolute_residual_criterion = 0;
    o->satisfy_residual_and_shift = 0; o->residual_reduction = 1e-5; o->max_absolute_residual = 1e-5;
}

// NewtonSolver::solveImpl_ with the non-default update strategies and criteria: newtonBeginStep (newtonsolver.hh:448-461),
// solveLinearSystem's initial residual (:495), newtonUpdate (:543-566), lineSearchUpdate_ (:1154-1178),
// computeResidualReduction_ (:869-881), newtonConverged (:657-701), newtonProceed (:428-446)
int orc_newton_solve_ex(orc_problem* p, double* u, const double* prev, double lin_reduction, int lin_maxit,
                        double max_rel_shift, int min_steps, int max_steps, const orc_newton_options* opt,
                        orc_newton_report* rep, double* relaxation, double* reduction_out)
{
    const int n = p->n, b = p->b;
    const size_t N = (size_t)n * b;
    const size_t nnz = p->colidx.size();
    std::vector<double> J(nnz * b * b), r(N), delta(N), uLast(u, u + N);
    const bool shiftCrit = opt->enable_shift_criterion != 0;
    const bool absResCrit = opt->enable_absolute_residual_criterion != 0;
    const bool resCrit = opt->enable_residual_criterion != 0 || absResCrit;
    const bool lineSearch = opt->use_line_search != 0;
    int numSteps = 0;
    double shift = 0.0, lastShift = 0.0, reduction = 1.0, lastReduction = 1.0, residualNorm = 0.0, initialResidual = 0.0;
    bool converged = false;
    std::memset(rep, 0, sizeof(*rep));
    auto proceed = [&]() {
        if (numSteps < min_steps) return true;
        else if (converged) return false;
        else if (numSteps >= max_steps) return shiftCrit ? shift * 4.0 < lastShift : reduction * 4.0 < lastReduction;
        return true;
    };
    auto isConverged = [&]() {
        const bool resOk = absResCrit ? residualNorm <= opt->max_absolute_residual : reduction <= opt->residual_reduction;
        if (shiftCrit && !resCrit) return shift <= max_rel_shift;
        if (!shiftCrit && resCrit) return resOk;
        if (opt->satisfy_residual_and_shift) return shift <= max_rel_shift && resOk;
        return shift <= max_rel_shift || resOk;
    };
    while (proceed()) {
        lastShift = shift;
        lastReduction = numSteps == 0 ? 1.0 : reduction;
        uLast.assign(u, u + N);
        orc_assemble(p, u, prev, r.data(), J.data());
        for (size_t i = 0; i < N; ++i)
            if (!(r[i] == r[i]) || std::isinf(r[i])) return 3;
        if (numSteps == 0) initialResidual = std::sqrt(dot(N, r.data(), r.data()));
        std::fill(delta.begin(), delta.end(), 0.0);
        int its = 0;
        double red = 0;
        const int st = orcLinearSolve(p, J.data(), delta.data(), r.data(), lin_reduction, lin_maxit, &its, &red);
        if (numSteps < 64) rep->linear_iterations[numSteps] = its;
        rep->linear_iterations_total += its;
        if (st != 0) { rep->newton_iterations = numSteps; rep->converged = 0; return st; }
        double lambda = 1.0;
        while (true) {
            for (size_t i = 0; i < N; ++i) u[i] = uLast[i] + (-lambda) * delta[i];     // axpy(-lambda, deltaU, uCurrentIter)
            if (lineSearch || resCrit) {
                orc_assemble(p, u, prev, r.data(), nullptr);
                residualNorm = std::sqrt(dot(N, r.data(), r.data()));
                reduction = residualNorm / initialResidual;
            }
            if (!lineSearch || reduction < lastReduction || lambda <= opt->line_search_min_relaxation) break;
            lambda *= 0.5;
        }
        if (shiftCrit) shift = orc_max_relative_shift((int)N, u, uLast.data());
        if (numSteps < 64) { rep->shifts[numSteps] = shift; if (relaxation) relaxation[numSteps] = lambda; }
        ++numSteps;
        converged = isConverged();
    }
    rep->newton_iterations = numSteps;
    rep->converged = converged ? 1 : 0;
    rep->last_shift = shift;
    if (reduction_out) *reduction_out = reduction;
    return converged ? 0 : 1;
}

// test/porousmediumflow/2p/incompressible/main.cc:126-163 with dumux/common/timeloop.hh:239-252,320-332,385-411
// and NewtonSolver::solve(vars,timeLoop) retry logic newtonsolver.hh:309-355, suggestTimeStepSize :784-798
int orc_run_timeloop(orc_problem* p, double* u, double t_end, double dt_initial, double max_dt,
                     int* newton_its, double* dts, int max_steps_out)
{
    const size_t N = (size_t)p->n * p->b;
    std::vector<double> uOld(u, u + N);
    double time = 0.0, tStart = 0.0;
    const double baseEps = 1e-10;
    auto finished = [&]() { return (t_end - time) < baseEps * (time - tStart); };
    auto maxTimeStepSize = [&]() { return finished() ? 0.0 : std::min(max_dt, std::max(0.0, t_end - time)); };
    double dt = std::min(dt_initial, maxTimeStepSize());
    int step = 0;
    const int targetSteps = 10;
    do {
        int numSteps = 0;
        bool ok = false;
        for (int i = 0; i <= 10; ++i) {
            p->opt.dt = dt;
            orc_newton_report rep;
            const int st = orc_newton_solve(p, u, uOld.data(), 1e-6, 250, 1e-8, 2, 18, &rep);
            numSteps = rep.newton_iterations;
            if (st == 0) { ok = true; break; }
            if (i < 10) {
                std::copy(uOld.begin(), uOld.end(), u);
                dt = std::min(dt * 0.5, maxTimeStepSize());
            }
        }
        if (!ok) return -1;
        uOld.assign(u, u + N);
        if (step < max_steps_out) { if (newton_its) newton_its[step] = numSteps; if (dts) dts[step] = dt; }
        ++step;
        time += dt;
        dt = std::min(dt, maxTimeStepSize());
        double suggested;
        if (numSteps > targetSteps) {
            const double percent = double(numSteps - targetSteps) / targetSteps;
            suggested = dt / (1.0 + percent);
        } else {
            const double percent = double(targetSteps - numSteps) / targetSteps;
            suggested = dt * (1.0 + percent / 1.2);
        }
        dt = std::min(suggested, maxTimeStepSize());
    } while (!finished());
    return step;
}

// ------------------------------------------------------------------------------------------------------------
// Tracer transport on a frozen velocity field (BASELINE config 5, examples/1ptracer).
// ------------------------------------------------------------------------------------------------------------
// Volume fluxes over all scvfs from a 1p pressure solution: examples/1ptracer/main.cc:162-199
// (fluxVars.advectiveFlux(0, upwindTerm = mobility)); Neumann boundary faces are skipped (stay 0).
// vf[I*2*dim + side]: flux through face `side` of cell I seen from I (each scvf has its own entry in the reference).
void orc_volume_flux(orc_problem* p, const double* pressure, double* vf)
{
    const int dim = p->dim;
    p->volumeFluxMode = true;
    for (int I = 0; I < p->n; ++I) {
        int cI[3];
        p->ijk(I, cI);
        VolVars vI;
        p->updateVolVars(vI, &pressure[I], I);
        for (int side = 0; side < 2 * dim; ++side) {
            double out = 0.0;
            const int J = p->neighbor(cI, side);
            double f[2] = {0.0, 0.0};
            if (J >= 0) {
                int cJ[3];
                p->ijk(J, cJ);
                VolVars vJ;
                p->updateVolVars(vJ, &pressure[J], J);
                p->computeFlux(f, cI, side, vI, vJ, false, cJ);
                out = f[0];
            } else {
                const int fidx = p->sideFaceIndex(side, cI);
                const int type = p->bcType[side].empty() ? ORC_BC_NEUMANN : p->bcType[side][fidx];
                if (type == ORC_BC_DIRICHLET) {
                    VolVars bv;
                    double pv[2] = {p->bcVal[side][(size_t)fidx], 0.0};
                    p->updateVolVars(bv, pv, I);
                    p->computeFlux(f, cI, side, vI, bv, true, nullptr);
                    out = f[0];
                }
            }
            vf[(size_t)I * 2 * dim + side] = out;
        }
    }
    p->volumeFluxMode = false;
}

// TracerLocalResidual (porousmediumflow/tracer/localresidual.hh:74-107 storage, :119-186 advective flux with
// StationaryVelocityField, flux/stationaryvelocityfield.hh:55; mass fractions, one component, D = 0) assembled by
//   implicit = 0: CCLocalAssembler<analytic, implicit=false> (assembly/cclocalassembler.hh:607-675): fluxes at prevSol,
//                 Jacobian = storage derivative on the diagonal (localresidual.hh:193-214)
//   implicit = 1: CCLocalAssembler<analytic, implicit=true> (:490-600): fluxes at curSol, addFluxDerivatives (:237-296).
//                 The reference has no analytic derivative for the solution-dependent outflow Neumann term
//                 (fvlocalresidual.hh:455-464 throws); here it is volumeFlux*rho*extrusion on the diagonal (OUR extension).
// Boundary faces: ORC_BC_NEUMANN with a fixed flux, ORC_BC_OUTFLOW = volumeFlux*X*rho/area (problem_tracer.hh:92-115).
void orc_tracer_assemble(orc_problem* p, const double* vf, const double* cur, const double* prev, int implicit, double rho,
                         double* residual, double* jac)
{
    const int dim = p->dim;
    const double w = p->opt.upwind_weight;
    const double extr = p->opt.extrusion;
    const double* X = implicit ? cur : prev;
    if (jac) std::fill(jac, jac + p->colidx.size(), 0.0);
    for (int I = 0; I < p->n; ++I) {
        int cI[3];
        p->ijk(I, cI);
        const double vol = p->volume(cI);
        double res = 0.0;
        {
            double source = p->q.empty() ? 0.0 : p->q[I];
            source *= vol * extr;
            res -= source;
        }
        const double phiInert = 1.0 - p->phi[I];
        const double porosity = 1.0 - phiInert;
        const double saturation = std::max(1e-8, 1.0);
        // position of the diagonal / neighbours in row I
        auto pos = [&](int col) {
            for (int k = p->rowptr[I]; k < p->rowptr[I + 1]; ++k)
                if (p->colidx[k] == col) return k;
            return -1;
        };
        double diag = 0.0;
        if (jac) {
            const double d_storage = vol * porosity * rho * saturation / p->opt.dt;
            diag += d_storage;
        }
        for (int side = 0; side < 2 * dim; ++side) {
            const int a = side / 2;
            const double vflux = vf[(size_t)I * 2 * dim + side];
            const int J = p->neighbor(cI, side);
            if (J >= 0) {
                const double upIn = rho * X[I], upOut = rho * X[J];
                double mult;
                if (std::signbit(vflux)) mult = w * upOut + (1.0 - w) * upIn;
                else mult = w * upIn + (1.0 - w) * upOut;
                // Fick's law, TPFA (flux/cctpfa/fickslaw.hh:120-175 flux_, :177-230 calculateTransmissibility) with
                // DiffusivityConstantTortuosity (diffusivityconstanttortuosity.hh:55-63): D_eff = porosity * S * tau * D, S = 1;
                // mass-averaged reference system with mass fractions: flux = rho_avg * tij * (X_I - X_J)
                int cJ[3];
                p->ijk(J, cJ);
                const double porosityJ = 1.0 - (1.0 - p->phi[J]);
                const double DeffI = porosity * 1.0 * p->tracerTau * p->tracerD;
                const double DeffJ = porosityJ * 1.0 * p->tracerTau * p->tracerD;
                const double dTij = p->advectionTij(cI, side, DeffI, extr, false, cJ, DeffJ, extr);
                const double rhoAvg = 0.5 * (rho + rho);
                double flux = 0.0;
                flux += vflux * mult;
                flux += rhoAvg * dTij * (X[I] - X[J]);
                if (!p->tracerDisp.empty()) {
                    // dispersionflux.hh:93-104 + calculateTransmissibility_ :172-213 with D_i = D_j = the tensor at the face; the
                    // reference's analytic Jacobian (localresidual.hh:237-291) has NO dispersion derivative, so neither has this one
                    const double Dd = p->tracerDisp[(size_t)I * 2 * dim + side];
                    const double mTij = p->advectionTij(cI, side, Dd, extr, false, cJ, Dd, extr);
                    flux += rhoAvg * mTij * (X[I] - X[J]);
                }
                res += flux;
                if (jac && implicit) {
                    const double insideWeight = std::signbit(vflux) ? (1.0 - w) : w;
                    const double outsideWeight = 1.0 - insideWeight;
                    const double diffDeriv = rhoAvg * dTij;           // tracer/localresidual.hh:268-291
                    diag += (vflux * rho * insideWeight + diffDeriv);
                    jac[pos(J)] += (vflux * rho * outsideWeight - diffDeriv);
                }
            } else {
                const int fidx = p->sideFaceIndex(side, cI);
                const int type = p->bcType[side].empty() ? ORC_BC_NEUMANN : p->bcType[side][fidx];
                if (type == ORC_BC_NONE) continue;
                const double area = p->faceArea(a, cI);
                double nf;
                if (type == ORC_BC_OUTFLOW) {
                    nf = vflux * X[I] * rho / area;
                    if (jac && implicit) diag += vflux * rho * extr;
                } else
                    nf = p->bcVal[side].empty() ? 0.0 : p->bcVal[side][fidx];
                nf *= area * extr;
                res += nf;
            }
        }
        {
            // fvlocalresidual.hh:274-304
            double prevStorage = porosity * rho * prev[I] * saturation;
            double storage = porosity * rho * cur[I] * saturation;
            prevStorage *= extr;
            storage *= extr;
            storage -= prevStorage;
            storage *= vol;
            storage /= p->opt.dt;
            double st = 0.0;
            st += storage;
            res += st;
        }
        if (residual) residual[I] = res;
        if (jac) jac[pos(I)] += diag;
    }
}

double orc_law_eval(orc_problem* p, int region, int which, double sw)
{
    const Law& l = p->laws[region];
    switch (which) {
        case 0: return l.pc(sw);
        case 1: return l.krw(sw);
        case 2: return l.krn(sw);
        case 3: return l.dpc_dsw(sw);
        case 4: return l.dkrw_dsw(sw);
        case 5: return l.dkrn_dsw(sw);
    }
    return 0.0;
}
double orc_pow(double x, double y) { return orc_det_pow(x, y); }

void orc_set_fluid_table(orc_problem* p, int nT, int nP, double Tmin, double Tmax, const double* pmin, const double* pmax,
                         const double* rho, const double* mu, double temperature)
{
    Fluids& f = p->fluids;
    f.tabulated = true;
    f.nT = nT; f.nP = nP; f.Tmin = Tmin; f.Tmax = Tmax; f.T = temperature;
    f.pmin.assign(pmin, pmin + nT);
    f.pmax.assign(pmax, pmax + nT);
    f.rhoTab.assign(rho, rho + (size_t)nT * nP);
    f.muTab.assign(mu, mu + (size_t)nT * nP);
}

} // extern "C"
