// physics.cuh -- constitutive relations of the 1p / 2p immiscible models, host + device.
//
// Mirrors (same names, same argument meaning):
//   TwoPMaterialLaw wrapper        dumux/material/fluidmatrixinteractions/2p/materiallaw.hh:104-243
//   TwoPEffToAbsDefaultPolicy      .../2p/efftoabsdefaultpolicy.hh:106-152
//   BrooksCorey (+Regularization)  .../2p/brookscorey.hh:100-108,165-174,215-223,266-276,365-477,481-489
//   VanGenuchten (+Regularization) .../2p/vangenuchten.hh (pc, krw, krn, derivatives, spline regularisation)
//   Spline<Scalar,2>               dumux/common/spline.hh:263-347, splinecommon_.hh:419-436,478-523
//   TabulatedComponent lookup      dumux/material/components/tabulatedcomponent.hh:1166-1203
// All arithmetic is plain IEEE in source order (device code is compiled with -fmad=false), pow is dmx::det_pow.
#pragma once
#include "det_math.cuh"

namespace dmx {

enum { LAW_BROOKSCOREY = 0, LAW_VANGENUCHTEN = 1 };

struct Spline2 {
    double x0, x1, y0, y1, M0, M1;
};

DMX_HD void spline_set(Spline2& s, double x0, double x1, double y0, double y1, double m0, double m1)
{
    s.x0 = x0; s.x1 = x1; s.y0 = y0; s.y1 = y1;
    const double h = x1 - x0;
    // moment system [[2,1],[1,2]] M = d, solved with the explicit 2x2 formula of Dune::FieldMatrix::solve
    const double m00 = 2, m01 = 1, m10 = 1, m11 = 2;
    const double d0 = 6 / h * ((y1 - y0) / h - m0);
    const double d1 = 6 / h * (m1 - (y1 - y0) / h);
    double detinv = m00 * m11 - m01 * m10;
    detinv = 1.0 / detinv;
    s.M0 = detinv * (m11 * d0 - m01 * d1);
    s.M1 = detinv * (m00 * d1 - m10 * d0);
}
DMX_HD double spline_eval(const Spline2& s, double x)
{
    const double h = s.x1 - s.x0;
    const double xi = x - s.x0;
    const double xi1 = s.x1 - x;
    const double A = (s.y1 - s.y0) / h - h / 6 * (s.M1 - s.M0);
    const double B = s.y0 - s.M0 * (h * h) / 6;
    return s.M0 * xi1 * xi1 * xi1 / (6 * h) + s.M1 * xi * xi * xi / (6 * h) + A * xi + B;
}
DMX_HD double spline_eval_derivative(const Spline2& s, double x)
{
    const double h = s.x1 - s.x0;
    const double xi = x - s.x0;
    const double xi1 = s.x1 - x;
    const double A = (s.y1 - s.y0) / h - h / 6 * (s.M1 - s.M0);
    return -s.M0 * xi1 * xi1 / (2 * h) + s.M1 * xi * xi / (2 * h) + A;
}

// One pc-kr-Sw relation with its effective-to-absolute policy and regularisation ("fluidMatrixInteraction").
struct MaterialLaw {
    int kind;
    int regularized;
    int wetting;                   // wetting phase index (spatialParams.wettingPhase, 2p/volumevariables.hh:132): 0 or 1
    int pad_;
    double swr, snr;
    double pcEntry, lambda;        // Brooks-Corey
    double alpha, n, m, l;         // van Genuchten
    double pcLowSwe, pcHighSwe, krnLowSwe, krwHighSwe;
    double pcLowSwePcValue, pcHighSwePcValue, pcDerivativeLowSw, pcDerivativeHighSwEnd, pcDerivativeHighSweThreshold;
    Spline2 pcSpline, krwSpline, krnSpline;
    double sdenom, rsdenom;        // 1 - swr - snr and its correctly rounded reciprocal (law_init; used by law_eval3)
};

DMX_HD double clamp01(double v) { return v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v); }
DMX_HD double swToSwe(const MaterialLaw& p, double sw) { return (sw - p.swr) / (1.0 - p.swr - p.snr); }
DMX_HD double sweToSw(const MaterialLaw& p, double swe) { return swe * (1.0 - p.swr - p.snr) + p.swr; }
DMX_HD double dswe_dsw(const MaterialLaw& p) { return 1.0 / (1.0 - p.swr - p.snr); }
DMX_HD double dsw_dswe(const MaterialLaw& p) { return 1.0 - p.swr - p.snr; }

// ---- base laws on the effective saturation ----
DMX_HD double base_pc(const MaterialLaw& p, double swe)
{
    swe = clamp01(swe);
    if (p.kind == LAW_BROOKSCOREY) return p.pcEntry * det_pow(swe, -1.0 / p.lambda);
    return det_pow(det_pow(swe, -1.0 / p.m) - 1, 1.0 / p.n) / p.alpha;
}
DMX_HD double base_dpc_dswe(const MaterialLaw& p, double swe)
{
    swe = clamp01(swe);
    if (p.kind == LAW_BROOKSCOREY) return -p.pcEntry / p.lambda * det_pow(swe, -1.0 / p.lambda - 1.0);
    const double powSwe = det_pow(swe, -1 / p.m);
    return -1.0 / p.alpha * det_pow(powSwe - 1, 1.0 / p.n - 1) / p.n * powSwe / swe / p.m;
}
DMX_HD double base_krw(const MaterialLaw& p, double swe)
{
    swe = clamp01(swe);
    if (p.kind == LAW_BROOKSCOREY) return det_pow(swe, 2.0 / p.lambda + 3.0);
    const double r = 1.0 - det_pow(1.0 - det_pow(swe, 1.0 / p.m), p.m);
    return det_pow(swe, p.l) * r * r;
}
DMX_HD double base_dkrw_dswe(const MaterialLaw& p, double swe)
{
    swe = clamp01(swe);
    if (p.kind == LAW_BROOKSCOREY) return (2.0 / p.lambda + 3.0) * det_pow(swe, 2.0 / p.lambda + 2.0);
    const double x = 1.0 - det_pow(swe, 1.0 / p.m);
    const double xToM = det_pow(x, p.m);
    return (1.0 - xToM) * det_pow(swe, p.l - 1) * ((1.0 - xToM) * p.l + 2 * xToM * (1.0 - x) / x);
}
DMX_HD double base_krn(const MaterialLaw& p, double swe)
{
    swe = clamp01(swe);
    if (p.kind == LAW_BROOKSCOREY) {
        const double exponent = 2.0 / p.lambda + 1.0;
        const double sne = 1.0 - swe;
        return sne * sne * (1.0 - det_pow(swe, exponent));
    }
    return det_pow(1 - swe, p.l) * det_pow(1 - det_pow(swe, 1.0 / p.m), 2 * p.m);
}
DMX_HD double base_dkrn_dswe(const MaterialLaw& p, double swe)
{
    swe = clamp01(swe);
    if (p.kind == LAW_BROOKSCOREY) {
        const double lambdaInv = 1.0 / p.lambda;
        const double swePow = det_pow(swe, 2 * lambdaInv);
        return 2.0 * (swe - 1.0) * (1.0 + (0.5 + lambdaInv) * swePow - (1.5 + lambdaInv) * swePow * swe);
    }
    const double sne = 1.0 - swe;
    const double x = 1.0 - det_pow(swe, 1.0 / p.m);
    return -det_pow(sne, p.l - 1.0) * det_pow(x, 2 * p.m - 1.0) * (p.l * x + 2.0 * sne / swe * (1.0 - x));
}

// ---- regularised laws on the absolute saturation (what VolumeVariables::update calls) ----
DMX_HD double law_pc(const MaterialLaw& p, double sw)
{
    const double swe = swToSwe(p, sw);
    if (p.regularized) {
        if (swe <= p.pcLowSwe) return p.pcLowSwePcValue + p.pcDerivativeLowSw * (swe - p.pcLowSwe);
        if (p.kind == LAW_BROOKSCOREY) {
            if (swe >= 1.0) return p.pcDerivativeHighSwEnd * (swe - 1.0) + p.pcEntry;
        } else {
            if (swe >= 1.0) return p.pcDerivativeHighSwEnd * (swe - 1.0);
            else if (swe > p.pcHighSwe) return spline_eval(p.pcSpline, swe);
        }
    }
    return base_pc(p, swe);
}
DMX_HD double law_krw(const MaterialLaw& p, double sw)
{
    const double swe = swToSwe(p, sw);
    if (p.regularized) {
        if (swe <= 0.0) return 0.0;
        else if (swe >= 1.0) return 1.0;
        else if (p.kind == LAW_VANGENUCHTEN && swe >= p.krwHighSwe) return spline_eval(p.krwSpline, swe);
    }
    return base_krw(p, swe);
}
DMX_HD double law_krn(const MaterialLaw& p, double sw)
{
    const double swe = swToSwe(p, sw);
    if (p.regularized) {
        if (swe <= 0.0) return 1.0;
        else if (swe >= 1.0) return 0.0;
        else if (p.kind == LAW_VANGENUCHTEN && swe <= p.krnLowSwe) return spline_eval(p.krnSpline, swe);
    }
    return base_krn(p, swe);
}

// regularised derivatives w.r.t. the absolute wetting saturation (materiallaw.hh dpc_dsw / dkrw_dsw / dkrn_dsw with the
// regularisation of brookscorey.hh / vangenuchten.hh): what the analytic 2p Jacobian (2p/incompressiblelocalresidual.hh) calls
DMX_HD double law_dpc_dsw(const MaterialLaw& p, double sw)
{
    const double swe = swToSwe(p, sw);
    if (p.regularized) {
        if (swe <= p.pcLowSwe) return p.pcDerivativeLowSw * dswe_dsw(p);
        else if (swe >= 1.0) return p.pcDerivativeHighSwEnd * dswe_dsw(p);
        else if (p.kind == LAW_VANGENUCHTEN && swe > p.pcHighSwe) return spline_eval_derivative(p.pcSpline, swe) * dswe_dsw(p);
    }
    return base_dpc_dswe(p, swe) * dswe_dsw(p);
}
DMX_HD double law_dkrw_dsw(const MaterialLaw& p, double sw)
{
    const double swe = swToSwe(p, sw);
    if (p.regularized) {
        if (swe <= 0.0) return 0.0;
        else if (swe >= 1.0) return 0.0;
        else if (p.kind == LAW_VANGENUCHTEN && swe >= p.krwHighSwe) return spline_eval_derivative(p.krwSpline, swe) * dswe_dsw(p);
    }
    return base_dkrw_dswe(p, swe) * dswe_dsw(p);
}
DMX_HD double law_dkrn_dsw(const MaterialLaw& p, double sw)
{
    const double swe = swToSwe(p, sw);
    if (p.regularized) {
        if (swe <= 0.0) return 0.0;
        else if (swe >= 1.0) return 0.0;
        else if (p.kind == LAW_VANGENUCHTEN && swe <= p.krnLowSwe) return spline_eval_derivative(p.krnSpline, swe) * dswe_dsw(p);
    }
    return base_dkrn_dswe(p, swe) * dswe_dsw(p);
}

// derived regularisation constants: brookscorey.hh:481-489; vangenuchten.hh initPcParameters_/initKrParameters_
inline void law_init(MaterialLaw& p)
{
    p.sdenom = 1.0 - p.swr - p.snr;
    p.rsdenom = 1.0 / p.sdenom;
    if (!p.regularized) return;
    const double dsw = dsw_dswe(p);
    auto dpc_dsw_noreg = [&](double sw) { return base_dpc_dswe(p, swToSwe(p, sw)) * dswe_dsw(p); };
    auto pc_noreg = [&](double sw) { return base_pc(p, swToSwe(p, sw)); };
    if (p.kind == LAW_BROOKSCOREY) {
        const double lowSw = sweToSw(p, p.pcLowSwe);
        const double highSw = sweToSw(p, 1.0);
        p.pcDerivativeLowSw = dpc_dsw_noreg(lowSw) * dsw;
        p.pcDerivativeHighSwEnd = dpc_dsw_noreg(highSw) * dsw;
        p.pcLowSwePcValue = pc_noreg(lowSw);
        return;
    }
    {
        const double lowSw = sweToSw(p, p.pcLowSwe);
        const double highSw = sweToSw(p, p.pcHighSwe);
        p.pcDerivativeLowSw = dpc_dsw_noreg(lowSw) * dsw;
        p.pcDerivativeHighSweThreshold = dpc_dsw_noreg(highSw) * dsw;
        p.pcDerivativeHighSwEnd = 2.0 * (0.0 - pc_noreg(highSw)) / (1.0 - p.pcHighSwe);
        p.pcLowSwePcValue = pc_noreg(lowSw);
        p.pcHighSwePcValue = pc_noreg(highSw);
        if (p.pcHighSwe < 1.0)
            spline_set(p.pcSpline, p.pcHighSwe, 1.0, p.pcHighSwePcValue, 0, p.pcDerivativeHighSweThreshold, p.pcDerivativeHighSwEnd);
    }
    {
        const double lowSw = sweToSw(p, p.krnLowSwe);
        const double highSw = sweToSw(p, p.krwHighSwe);
        const double krwHighSw = base_krw(p, swToSwe(p, highSw));
        const double dkrwHighSw = base_dkrw_dswe(p, swToSwe(p, highSw)) * dswe_dsw(p) * dsw;
        const double krnLowSw = base_krn(p, swToSwe(p, lowSw));
        const double dkrnLowSw = base_dkrn_dswe(p, swToSwe(p, lowSw)) * dswe_dsw(p) * dsw;
        if (p.krwHighSwe < 1.0) spline_set(p.krwSpline, p.krwHighSwe, 1.0, krwHighSw, 1.0, dkrwHighSw, 0.0);
        if (p.krnLowSwe > 0.0) spline_set(p.krnSpline, 0.0, p.krnLowSwe, 1.0, krnLowSw, 0.0, dkrnLowSw);
    }
}

// pc, krw and krn of one saturation in one go (what TwoPVolumeVariables::completeFluidState needs,
// porousmediumflow/2p/volumevariables.hh:141-190).  Returns exactly the bits of law_pc / law_krw / law_krn: the
// same operations in the same order, but powers of a common base share one det_log2 (PowBase) and the division by
// 1 - swr - snr uses the precomputed reciprocal (div_by is correctly rounded).
struct Law3 {
    double pc, krw, krn;
};
DMX_HD void law_eval3(const MaterialLaw& p, double sw, double* pc, double* krw, double* krn);
// out-of-line entry for the kernels (results in registers)
DMX_HD_NOINLINE Law3 law_eval3_call(const MaterialLaw* p, double sw)
{
    Law3 r;
    law_eval3(*p, sw, &r.pc, &r.krw, &r.krn);
    return r;
}
DMX_HD void law_eval3(const MaterialLaw& p, double sw, double* pc, double* krw, double* krn)
{
    const double swe = div_by(sw - p.swr, p.sdenom, p.rsdenom);
    const bool reg = p.regularized != 0;
    const bool mid = !reg || (swe > 0.0 && swe < 1.0);
    if (!mid) {
        // end points: the kr curves are constants, pc is linear (or, for an exotic pcLowSwe < 0, the plain law)
        *pc = law_pc(p, sw);
        *krw = swe <= 0.0 ? 0.0 : 1.0;
        *krn = swe <= 0.0 ? 1.0 : 0.0;
        return;
    }
    const double c = clamp01(swe);
    const PowBase B(c);
    if (p.kind == LAW_BROOKSCOREY) {
        if (reg && swe <= p.pcLowSwe) *pc = p.pcLowSwePcValue + p.pcDerivativeLowSw * (swe - p.pcLowSwe);
        else if (reg && swe >= 1.0) *pc = p.pcDerivativeHighSwEnd * (swe - 1.0) + p.pcEntry;
        else *pc = p.pcEntry * B.pow(-1.0 / p.lambda);
        *krw = B.pow(2.0 / p.lambda + 3.0);
        const double exponent = 2.0 / p.lambda + 1.0;
        const double sne = 1.0 - c;
        *krn = sne * sne * (1.0 - B.pow(exponent));
        return;
    }
    if (reg && swe <= p.pcLowSwe) *pc = p.pcLowSwePcValue + p.pcDerivativeLowSw * (swe - p.pcLowSwe);
    else if (reg && swe >= 1.0) *pc = p.pcDerivativeHighSwEnd * (swe - 1.0);
    else if (reg && swe > p.pcHighSwe) *pc = spline_eval(p.pcSpline, swe);
    else *pc = det_pow(B.pow(-1.0 / p.m) - 1, 1.0 / p.n) / p.alpha;
    const bool krwSpl = reg && swe >= p.krwHighSwe;
    const bool krnSpl = reg && swe <= p.krnLowSwe;
    if (krwSpl && krnSpl) {
        *krw = spline_eval(p.krwSpline, swe);
        *krn = spline_eval(p.krnSpline, swe);
        return;
    }
    const PowBase BX(1.0 - B.pow(1.0 / p.m));
    if (krwSpl) *krw = spline_eval(p.krwSpline, swe);
    else {
        const double r = 1.0 - BX.pow(p.m);
        *krw = B.pow(p.l) * r * r;
    }
    if (krnSpl) *krn = spline_eval(p.krnSpline, swe);
    else *krn = det_pow(1 - c, p.l) * BX.pow(2 * p.m);
}

// ---- fluids ----
struct FluidTable {
    int nT, nP;
    double Tmin, Tmax, T;
    const double* pmin;   // [nT]
    const double* pmax;   // [nT]
    const double* rho;    // values[iT + iP*nT]
    const double* mu;
};

DMX_HD double table_interp(const FluidTable& t, const double* values, double p)
{
    double alphaT = (t.nT - 1) * (t.T - t.Tmin) / (t.Tmax - t.Tmin);
    if (alphaT < 0 - 1e-7 * t.nT || alphaT >= t.nT - 1 + 1e-7 * t.nT) return u2d(0x7ff8000000000000ull);
    int iT = (int)alphaT;
    iT = iT < 0 ? 0 : (iT > t.nT - 2 ? t.nT - 2 : iT);
    alphaT -= iT;
    double alphaP1 = (t.nP - 1) * (p - t.pmin[iT]) / (t.pmax[iT] - t.pmin[iT]);
    double alphaP2 = (t.nP - 1) * (p - t.pmin[iT + 1]) / (t.pmax[iT + 1] - t.pmin[iT + 1]);
    int iP1 = (int)alphaP1;
    iP1 = iP1 < 0 ? 0 : (iP1 > t.nP - 2 ? t.nP - 2 : iP1);
    int iP2 = (int)alphaP2;
    iP2 = iP2 < 0 ? 0 : (iP2 > t.nP - 2 ? t.nP - 2 : iP2);
    alphaP1 -= iP1;
    alphaP2 -= iP2;
    return values[(iT) + (iP1)*t.nT] * (1 - alphaT) * (1 - alphaP1) + values[(iT) + (iP1 + 1) * t.nT] * (1 - alphaT) * (alphaP1)
         + values[(iT + 1) + (iP2)*t.nT] * (alphaT) * (1 - alphaP2) + values[(iT + 1) + (iP2 + 1) * t.nT] * (alphaT) * (alphaP2);
}

// density and viscosity of one pressure in one go: same operations as two table_interp calls, shared index search
DMX_HD void table_interp2(const FluidTable& t, double p, double* rho, double* mu)
{
    double alphaT = (t.nT - 1) * (t.T - t.Tmin) / (t.Tmax - t.Tmin);
    if (alphaT < 0 - 1e-7 * t.nT || alphaT >= t.nT - 1 + 1e-7 * t.nT) {
        *rho = *mu = u2d(0x7ff8000000000000ull);
        return;
    }
    int iT = (int)alphaT;
    iT = iT < 0 ? 0 : (iT > t.nT - 2 ? t.nT - 2 : iT);
    alphaT -= iT;
    double alphaP1 = (t.nP - 1) * (p - t.pmin[iT]) / (t.pmax[iT] - t.pmin[iT]);
    double alphaP2 = (t.nP - 1) * (p - t.pmin[iT + 1]) / (t.pmax[iT + 1] - t.pmin[iT + 1]);
    int iP1 = (int)alphaP1;
    iP1 = iP1 < 0 ? 0 : (iP1 > t.nP - 2 ? t.nP - 2 : iP1);
    int iP2 = (int)alphaP2;
    iP2 = iP2 < 0 ? 0 : (iP2 > t.nP - 2 ? t.nP - 2 : iP2);
    alphaP1 -= iP1;
    alphaP2 -= iP2;
    const double* v = t.rho;
    *rho = v[(iT) + (iP1)*t.nT] * (1 - alphaT) * (1 - alphaP1) + v[(iT) + (iP1 + 1) * t.nT] * (1 - alphaT) * (alphaP1)
         + v[(iT + 1) + (iP2)*t.nT] * (alphaT) * (1 - alphaP2) + v[(iT + 1) + (iP2 + 1) * t.nT] * (alphaT) * (alphaP2);
    v = t.mu;
    *mu = v[(iT) + (iP1)*t.nT] * (1 - alphaT) * (1 - alphaP1) + v[(iT) + (iP1 + 1) * t.nT] * (1 - alphaT) * (alphaP1)
        + v[(iT + 1) + (iP2)*t.nT] * (alphaT) * (1 - alphaP2) + v[(iT + 1) + (iP2 + 1) * t.nT] * (alphaT) * (alphaP2);
}

} // namespace dmx
