// Host-side bit-exactness check of the shared-log material-law evaluation (dumux_b200/csrc/physics.cuh: law_eval3,
// table_interp2, div_by, PowBase) against the one-curve-at-a-time functions and the oracle's det_pow.
// Built and run by tests/test_host_physics.py (g++ -ffp-contract=off -mfma, like the oracle).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "../../dumux_b200/csrc/physics.cuh"
#include "../../oracle/det_math.h"

using namespace dmx;

static bool same(double a, double b) { return d2u(a) == d2u(b) || (a != a && b != b); }

static MaterialLaw make_law(int kind, int reg, double swr, double snr)
{
    MaterialLaw p;
    memset(&p, 0, sizeof(p));
    p.kind = kind; p.regularized = reg; p.swr = swr; p.snr = snr;
    p.pcEntry = 1234.5; p.lambda = 2.3;
    p.alpha = 0.0037; p.n = 4.7; p.m = 1.0 - 1.0 / p.n; p.l = 0.5;
    p.pcLowSwe = 0.01; p.pcHighSwe = 0.99; p.krnLowSwe = 0.1; p.krwHighSwe = 0.9;
    law_init(p);
    return p;
}

int main()
{
    std::mt19937_64 rng(7);
    std::uniform_real_distribution<double> U(0.0, 1.0);
    long bad = 0, n = 0;
    // det_pow / PowBase vs the oracle's det_pow
    for (int i = 0; i < 2000000; ++i) {
        double x = U(rng), y = (U(rng) - 0.5) * 12.0;
        if (i % 97 == 0) x = 1.0;
        if (i % 101 == 0) x = 0.0;
        if (i % 103 == 0) y = 0.0;
        if (i % 107 == 0) x = x * 1e-310;
        if (i % 109 == 0) x = 1.0 + x * 1e-12;
        const double ref = orc_det_pow(x, y);
        const PowBase B(x);
        if (!same(det_pow(x, y), ref) || !same(B.pow(y), ref) || !same(B.pow(-y), orc_det_pow(x, -y))) ++bad;
        ++n;
    }
    // div_by vs IEEE division
    for (int i = 0; i < 4000000; ++i) {
        const double b = (i & 1) ? 1e-10 * (U(rng) * 1e5 + 1.0) : U(rng) * 3.0 + 1e-4;
        const double a = (U(rng) - 0.5) * std::exp((U(rng) - 0.5) * 60.0);
        if (!same(div_by(a, b, 1.0 / b), a / b) && !(a == 0.0)) ++bad;
        ++n;
    }
    // law_eval3 vs law_pc / law_krw / law_krn
    for (int kind = 0; kind < 2; ++kind)
        for (int reg = 0; reg < 2; ++reg)
            for (int v = 0; v < 2; ++v) {
                const MaterialLaw p = make_law(kind, reg, v ? 0.18 : 0.05, v ? 0.1 : 0.0);
                for (int i = 0; i < 400000; ++i) {
                    double sw;
                    switch (i % 8) {
                    case 0: sw = 1.0 - 1e-10 * (1 + i % 5); break;       // FD-deflected fully saturated cell
                    case 1: sw = 1.0; break;
                    case 2: sw = p.swr + U(rng) * 0.02; break;
                    case 3: sw = 1.0 - p.snr - U(rng) * 0.02; break;
                    case 4: sw = -0.1 + 1.3 * U(rng); break;
                    default: sw = U(rng);
                    }
                    if (!reg && (sw <= p.swr || sw >= 1.0 - p.snr)) sw = p.swr + 0.5 * (1.0 - p.snr - p.swr);
                    double pc, krw, krn;
                    law_eval3(p, sw, &pc, &krw, &krn);
                    if (!same(pc, law_pc(p, sw)) || !same(krw, law_krw(p, sw)) || !same(krn, law_krn(p, sw))) {
                        if (bad < 10)
                            printf("law mismatch kind %d reg %d sw %.17g: %a %a | %a %a | %a %a\n", kind, reg, sw, pc, law_pc(p, sw), krw,
                                   law_krw(p, sw), krn, law_krn(p, sw));
                        ++bad;
                    }
                    ++n;
                }
            }
    // table_interp2 vs table_interp
    {
        const int nT = 10, nP = 200;
        std::vector<double> pmin(nT, 1e4), pmax(nT, 1e6), rho(nT * nP), mu(nT * nP);
        for (auto& r : rho) r = 990.0 + 20.0 * U(rng);
        for (auto& m : mu) m = 1e-3 * (0.8 + 0.4 * U(rng));
        FluidTable t{nT, nP, 273.15, 294.15, 293.15, pmin.data(), pmax.data(), rho.data(), mu.data()};
        for (int i = 0; i < 400000; ++i) {
            const double p = 1e4 + U(rng) * 1.2e6 - 1e5;
            double r, m;
            table_interp2(t, p, &r, &m);
            if (!same(r, table_interp(t, t.rho, p)) || !same(m, table_interp(t, t.mu, p))) ++bad;
            ++n;
        }
    }
    printf("checked %ld cases, %ld mismatches\n", n, bad);
    return bad ? 1 : 0;
}
