/*
 * oracle/det_math.h -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 *
 * Deterministic pow() built from IEEE-754 +,-,*,/ and correctly-rounded fma only.
 *
 * Why it exists: the reference differentiates its residual numerically with
 * eps = 1e-10*(|x|+1) (dumux/common/numericdifferentiation.hh:36-41,
 * dumux/assembly/numericepsilon.hh:47-51).  A last-ulp difference between glibc's
 * pow and CUDA's pow inside the Brooks-Corey / van Genuchten laws
 * (dumux/material/fluidmatrixinteractions/2p/brookscorey.hh:100-108,215-223,266-276;
 * vangenuchten.hh) is amplified by 1/eps = 1e10, so GPU-vs-CPU Jacobian parity at 1e-10
 * needs a pow that executes the same IEEE operation sequence on both sides.
 * This file is the CPU statement of that sequence (the "spec"); the CUDA product carries
 * its own statement of the same sequence.  std::pow is available as an oracle switch so the
 * deviation from a glibc build of the reference can be quantified.
 *
 * Algorithm (all steps exact or one correctly-rounded op; no libm):
 *   x = 2^e * m, m in [0.75,1.5);  s = (m-1)/(m+1) as double-double (division residual by fma)
 *   log(m) = 2s + s^3 * P(s^2)       (atanh series, 13 tail terms, |s| <= 0.2)
 *   log2(x) = e + log(m)/ln2         (double-double product with a hi/lo split of 1/ln2)
 *   z = y*log2(x) as double-double;  n = round(z);  r = z-n in [-0.5,0.5]
 *   2^r = sum_{k<=14} (ln2^k/k!) r^k (Horner in fma);  result = 2^r * 2^n (exponent bits)
 * Accuracy: <= 1 ulp against glibc on (0,1] x [-8,8] (tests/test_det_math.py).
 */
#ifndef ORACLE_DET_MATH_H
#define ORACLE_DET_MATH_H

#include <stdint.h>
#include <string.h>

static inline double orc_fma(double a, double b, double c) { return __builtin_fma(a, b, c); }
static inline uint64_t orc_d2u(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
static inline double orc_u2d(uint64_t u) { double x; memcpy(&x, &u, 8); return x; }

static inline double orc_det_pow(double x, double y)
{
    if (y == 0.0) return 1.0;
    if (x == 1.0) return 1.0;
    if (x != x || y != y) return x + y;
    if (x < 0.0) return orc_u2d(0x7ff8000000000000ull);
    if (x == 0.0) return y > 0.0 ? 0.0 : orc_u2d(0x7ff0000000000000ull);
    if (x == orc_u2d(0x7ff0000000000000ull)) return y > 0.0 ? x : 0.0;

    uint64_t bits = orc_d2u(x);
    int e = (int)((bits >> 52) & 0x7ff);
    if (e == 0) { /* subnormal: renormalise exactly */
        x = x * 18014398509481984.0; /* 2^54 */
        bits = orc_d2u(x);
        e = (int)((bits >> 52) & 0x7ff) - 54;
    }
    e -= 1023;
    double m = orc_u2d((bits & 0x000fffffffffffffull) | 0x3ff0000000000000ull);
    if (m >= 1.5) { m = m * 0.5; e += 1; }

    const double a = m - 1.0;            /* exact */
    const double b = m + 1.0;            /* may round */
    const double b_lo = m - (b - 1.0);   /* exact rounding error of b */
    const double s_hi = a / b;
    double res = orc_fma(-s_hi, b, a);
    res = orc_fma(-s_hi, b_lo, res);
    const double s_lo = res / b;
    const double s2 = s_hi * s_hi;

    double P = 0x1.2f684bda12f68p-4;          /* 2/27 */
    P = orc_fma(P, s2, 0x1.47ae147ae147bp-4);  /* 2/25 */
    P = orc_fma(P, s2, 0x1.642c8590b2164p-4);  /* 2/23 */
    P = orc_fma(P, s2, 0x1.8618618618618p-4);  /* 2/21 */
    P = orc_fma(P, s2, 0x1.af286bca1af28p-4);  /* 2/19 */
    P = orc_fma(P, s2, 0x1.e1e1e1e1e1e1ep-4);  /* 2/17 */
    P = orc_fma(P, s2, 0x1.1111111111111p-3);  /* 2/15 */
    P = orc_fma(P, s2, 0x1.3b13b13b13b14p-3);  /* 2/13 */
    P = orc_fma(P, s2, 0x1.745d1745d1746p-3);  /* 2/11 */
    P = orc_fma(P, s2, 0x1.c71c71c71c71cp-3);  /* 2/9  */
    P = orc_fma(P, s2, 0x1.2492492492492p-2);  /* 2/7  */
    P = orc_fma(P, s2, 0x1.999999999999ap-2);  /* 2/5  */
    P = orc_fma(P, s2, 0x1.5555555555555p-1);  /* 2/3  */
    const double tail = (s_hi * s2) * P;

    const double lh = 2.0 * s_hi;
    const double ll = orc_fma(2.0, s_lo, tail);
    const double th = lh + ll;
    const double tl = ll - (th - lh);

    const double INVLN2_HI = 0x1.71547652b82fep+0;
    const double INVLN2_LO = 0x1.777d0ffda0d24p-56;
    const double ph = th * INVLN2_HI;
    double pl = orc_fma(th, INVLN2_HI, -ph);
    pl = orc_fma(th, INVLN2_LO, pl);
    pl = orc_fma(tl, INVLN2_HI, pl);

    const double ed = (double)e;
    const double Lh = ed + ph;
    double Ll = (ed - Lh) + ph;
    Ll = Ll + pl;

    const double zh = y * Lh;
    double zl = orc_fma(y, Lh, -zh);
    zl = orc_fma(y, Ll, zl);

    if (zh >= 1024.0) return orc_u2d(0x7ff0000000000000ull);
    if (zh <= -1100.0) return 0.0;

    const long long n = (long long)(zh + (zh >= 0.0 ? 0.5 : -0.5));
    const double r = (zh - (double)n) + zl;

    double Q = 0x1.314964d5878a9p-44;
    Q = orc_fma(Q, r, 0x1.816193166d0f9p-40);
    Q = orc_fma(Q, r, 0x1.c3bd650fc2986p-36);
    Q = orc_fma(Q, r, 0x1.e8cac7351bb25p-32);
    Q = orc_fma(Q, r, 0x1.e4cf5158b8ecap-28);
    Q = orc_fma(Q, r, 0x1.b5253d395e7c4p-24);
    Q = orc_fma(Q, r, 0x1.62c0223a5c824p-20);
    Q = orc_fma(Q, r, 0x1.ffcbfc588b0c7p-17);
    Q = orc_fma(Q, r, 0x1.430912f86c787p-13);
    Q = orc_fma(Q, r, 0x1.5d87fe78a6731p-10);
    Q = orc_fma(Q, r, 0x1.3b2ab6fba4e77p-7);
    Q = orc_fma(Q, r, 0x1.c6b08d704a0c0p-5);
    Q = orc_fma(Q, r, 0x1.ebfbdff82c58fp-3);
    Q = orc_fma(Q, r, 0x1.62e42fefa39efp-1);
    Q = orc_fma(Q, r, 1.0);

    if (n >= -1022 && n <= 1023)
        return Q * orc_u2d((uint64_t)(n + 1023) << 52);
    if (n > 1023)
        return (Q * 0x1p1023) * orc_u2d((uint64_t)(n - 1023 + 1023) << 52);
    /* n < -1022: two-step scaling into the subnormal range */
    {
        long long n2 = n + 1022;
        if (n2 < -1022) n2 = -1022;
        return (Q * 0x1p-1022) * orc_u2d((uint64_t)(n2 + 1023) << 52);
    }
}

#endif
