// det_math.cuh -- deterministic pow() for the material laws, host + device.
//
// The FD Jacobian (eps = 1e-10*(|x|+1), dumux/common/numericdifferentiation.hh:36-41) amplifies last-ulp
// differences in pow by 1e10, so the kernels cannot use CUDA's libdevice pow (2-ulp, different bits from
// glibc).  dmx_pow executes one fixed sequence of IEEE +,-,*,/ and correctly rounded fma operations, so
// it returns the same bits on sm_100a and on any IEEE host (<= 1 ulp from glibc pow over the range the
// Brooks-Corey / van Genuchten laws use).  Compile device code with -fmad=false: the only fused
// operations are the explicit fma() calls below.
//
//   x = 2^e * m, m in [0.75,1.5);  s = (m-1)/(m+1) in double-double (division residual via fma)
//   log(m) = 2s + s^3 P(s^2) (atanh series);  log2(x) = e + log(m)/ln2 in double-double
//   z = y*log2(x);  n = round(z), r = z - n;  2^r by a degree-14 Taylor polynomial;  scale by 2^n.
#pragma once
#include <cstdint>
#include <cstring>
#include <cmath>

#ifdef __CUDACC__
#define DMX_HD __host__ __device__ __forceinline__
// out-of-line building blocks (one copy per translation unit): keeps the assembly kernels' code inside the instruction cache
#define DMX_HD_NOINLINE static __host__ __device__ __noinline__
#else
#define DMX_HD inline
#define DMX_HD_NOINLINE static inline
#endif

namespace dmx {

DMX_HD double u2d(uint64_t u)
{
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)u);
#else
    double x; memcpy(&x, &u, 8); return x;
#endif
}
DMX_HD uint64_t d2u(double x)
{
#ifdef __CUDA_ARCH__
    return (uint64_t)__double_as_longlong(x);
#else
    uint64_t u; memcpy(&u, &x, 8); return u;
#endif
}
DMX_HD double fma_rn(double a, double b, double c)
{
#ifdef __CUDA_ARCH__
    return __fma_rn(a, b, c);
#else
    return __builtin_fma(a, b, c);
#endif
}

// Polynomial coefficients (atanh series for log, Taylor series of 2^r).  On the device they live in the constant bank so
// that each Horner step is ONE DFMA with a c[][] operand (64-bit immediates cost two extra uniform moves per step).
#define DMX_DET_LOG_COEFFS {0x1.2f684bda12f68p-4, 0x1.47ae147ae147bp-4, 0x1.642c8590b2164p-4, 0x1.8618618618618p-4, 0x1.af286bca1af28p-4, 0x1.e1e1e1e1e1e1ep-4, 0x1.1111111111111p-3, 0x1.3b13b13b13b14p-3, 0x1.745d1745d1746p-3, 0x1.c71c71c71c71cp-3, 0x1.2492492492492p-2, 0x1.999999999999ap-2, 0x1.5555555555555p-1}
#define DMX_DET_EXP_COEFFS {0x1.314964d5878a9p-44, 0x1.816193166d0f9p-40, 0x1.c3bd650fc2986p-36, 0x1.e8cac7351bb25p-32, 0x1.e4cf5158b8ecap-28, 0x1.b5253d395e7c4p-24, 0x1.62c0223a5c824p-20, 0x1.ffcbfc588b0c7p-17, 0x1.430912f86c787p-13, 0x1.5d87fe78a6731p-10, 0x1.3b2ab6fba4e77p-7, 0x1.c6b08d704a0c0p-5, 0x1.ebfbdff82c58fp-3, 0x1.62e42fefa39efp-1}
#ifdef __CUDACC__
static __constant__ double det_log_c_dev[13] = DMX_DET_LOG_COEFFS;
static __constant__ double det_exp_c_dev[14] = DMX_DET_EXP_COEFFS;
#endif
static const double det_log_c_host[13] = DMX_DET_LOG_COEFFS;
static const double det_exp_c_host[14] = DMX_DET_EXP_COEFFS;
#ifdef __CUDA_ARCH__
#define DMX_DET_LOG_C det_log_c_dev
#define DMX_DET_EXP_C det_exp_c_dev
#else
#define DMX_DET_LOG_C det_log_c_host
#define DMX_DET_EXP_C det_exp_c_host
#endif

// log2(x) in double-double for finite x > 0: the first half of det_pow.  Several powers of the same base (the three
// Brooks-Corey curves, the nested van Genuchten terms) share ONE det_log2 -- same operation sequence per power, so
// det_pow_from_log2(det_log2(x), x, y) returns exactly the bits of det_pow(x, y).
struct DetLog2 {
    double Lh, Ll;
};

DMX_HD_NOINLINE DetLog2 det_log2(double x)
{
    uint64_t bits = d2u(x);
    int e = (int)((bits >> 52) & 0x7ff);
    if (e == 0) {
        x = x * 18014398509481984.0;
        bits = d2u(x);
        e = (int)((bits >> 52) & 0x7ff) - 54;
    }
    e -= 1023;
    double m = u2d((bits & 0x000fffffffffffffull) | 0x3ff0000000000000ull);
    if (m >= 1.5) { m = m * 0.5; e += 1; }

    const double a = m - 1.0;
    const double b = m + 1.0;
    const double b_lo = m - (b - 1.0);
    const double s_hi = a / b;
    double res = fma_rn(-s_hi, b, a);
    res = fma_rn(-s_hi, b_lo, res);
    const double s_lo = res / b;
    const double s2 = s_hi * s_hi;

    double P = DMX_DET_LOG_C[0];
#pragma unroll
    for (int q = 1; q < 13; ++q) P = fma_rn(P, s2, DMX_DET_LOG_C[q]);
    const double tail = (s_hi * s2) * P;

    const double lh = 2.0 * s_hi;
    const double ll = fma_rn(2.0, s_lo, tail);
    const double th = lh + ll;
    const double tl = ll - (th - lh);

    const double INVLN2_HI = 0x1.71547652b82fep+0;
    const double INVLN2_LO = 0x1.777d0ffda0d24p-56;
    const double ph = th * INVLN2_HI;
    double pl = fma_rn(th, INVLN2_HI, -ph);
    pl = fma_rn(th, INVLN2_LO, pl);
    pl = fma_rn(tl, INVLN2_HI, pl);

    const double ed = (double)e;
    DetLog2 L;
    L.Lh = ed + ph;
    double Ll = (ed - L.Lh) + ph;
    L.Ll = Ll + pl;
    return L;
}

// true when det_pow(x, y) takes the log/exp path for every y != 0 (finite x > 0, x != 1)
DMX_HD bool det_pow_regular(double x) { return x > 0.0 && x != 1.0 && x < u2d(0x7ff0000000000000ull); }

// 2^(y*L): the second half of det_pow
DMX_HD_NOINLINE double det_exp2_scaled(DetLog2 L, double y)
{
    const double zh = y * L.Lh;
    double zl = fma_rn(y, L.Lh, -zh);
    zl = fma_rn(y, L.Ll, zl);

    if (zh >= 1024.0) return u2d(0x7ff0000000000000ull);
    if (zh <= -1100.0) return 0.0;

    const long long n = (long long)(zh + (zh >= 0.0 ? 0.5 : -0.5));
    const double r = (zh - (double)n) + zl;

    double Q = DMX_DET_EXP_C[0];
#pragma unroll
    for (int q = 1; q < 14; ++q) Q = fma_rn(Q, r, DMX_DET_EXP_C[q]);
    Q = fma_rn(Q, r, 1.0);

    if (n >= -1022 && n <= 1023) return Q * u2d((uint64_t)(n + 1023) << 52);
    if (n > 1023) return (Q * 0x1p1023) * u2d((uint64_t)(n - 1023 + 1023) << 52);
    long long n2 = n + 1022;
    if (n2 < -1022) n2 = -1022;
    return (Q * 0x1p-1022) * u2d((uint64_t)(n2 + 1023) << 52);
}

// special cases of pow that never reach the log/exp path
DMX_HD double det_pow_special(double x, double y)
{
    if (y == 0.0) return 1.0;
    if (x == 1.0) return 1.0;
    if (x != x || y != y) return x + y;
    if (x < 0.0) return u2d(0x7ff8000000000000ull);
    if (x == 0.0) return y > 0.0 ? 0.0 : u2d(0x7ff0000000000000ull);
    return y > 0.0 ? x : 0.0;       // x == +inf
}

DMX_HD double det_pow(double x, double y)
{
    if (y == 0.0 || y != y || !det_pow_regular(x)) return det_pow_special(x, y);
    return det_exp2_scaled(det_log2(x), y);
}

// A base whose log2 is evaluated lazily at most once: pw.pow(y) == det_pow(x, y) bit for bit.
struct PowBase {
    double x;
    DetLog2 L;
    bool regular;
    DMX_HD explicit PowBase(double x_) : x(x_), regular(det_pow_regular(x_))
    {
        if (regular) L = det_log2(x_);
        else { L.Lh = 0.0; L.Ll = 0.0; }
    }
    DMX_HD double pow(double y) const
    {
        if (y == 0.0 || y != y || !regular) return det_pow_special(x, y);
        return det_exp2_scaled(L, y);
    }
};

/* Correctly rounded a/b from the correctly rounded reciprocal y = RN(1/b) (Markstein): q0 = RN(a*y), two residual
   corrections with exact FMA residuals.  Bit-identical to the IEEE division a/b for normal-range operands; used where
   many quotients share one divisor (FD steps, viscosities, dt). */
DMX_HD double div_by(double a, double b, double y)
{
    double q = a * y;
    double r = fma_rn(-b, q, a);
    q = fma_rn(r, y, q);
    r = fma_rn(-b, q, a);
    return fma_rn(r, y, q);
}

} // namespace dmx
