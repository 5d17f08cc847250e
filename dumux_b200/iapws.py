"""Host-side set-up of the tabulated liquid water used by the compressible 1p configuration (BASELINE config 2).

Restates what the reference's problem constructor does once at start-up,
`Components::TabulatedComponent<Components::H2O<Scalar>>::init(273.15, 294.15, 10, 1e4, 1e6, 200)`
(test/porousmediumflow/1p/compressible/instationary/problem.hh:46-47):

  * table layout, temperature / pressure sampling, vapour-pressure dependent pressure range per temperature:
    dumux/material/components/tabulatedcomponent.hh:138-200, 235-253, 330-345
  * H2O::liquidDensity / liquidViscosity incl. the low-pressure regularisation: components/h2o.hh:601-641, 722-728, 838-844
  * IAPWS-IF97 region 1 (dimensionless Gibbs energy derivative): components/iapws/region1.hh:69-87, 170-185, 277-332
  * IAPWS-IF97 region 4 saturation pressure: components/iapws/region4.hh:52-73
  * IAPWS 2008 viscosity: components/iapws/common.hh:88-134

The result is a plain dict that travels over the C ABI (`dmx_set_fluid_table`) and into the oracle
(`orc_set_fluid_table`); the device kernels only ever see the table (tabulatedcomponent.hh:1166-1203 lookup).
The coefficient tables below are the published IAPWS-IF97 / IAPWS-2008 constants.
"""
from __future__ import annotations

import math

import numpy as np

R_GAS = 8.314472                      # dumux/material/constants.hh:32
MOLAR_MASS = 18.01518e-3              # iapws/common.hh:47
RS = R_GAS / MOLAR_MASS
T_CRIT = 647.096
T_TRIPLE = 273.16
RHO_CRIT = 322.0

_N1 = (0.14632971213167, -0.84548187169114, -0.37563603672040e1, 0.33855169168385e1, -0.95791963387872, 0.15772038513228,
       -0.16616417199501e-1, 0.81214629983568e-3, 0.28319080123804e-3, -0.60706301565874e-3, -0.18990068218419e-1,
       -0.32529748770505e-1, -0.21841717175414e-1, -0.52838357969930e-4, -0.47184321073267e-3, -0.30001780793026e-3,
       0.47661393906987e-4, -0.44141845330846e-5, -0.72694996297594e-15, -0.31679644845054e-4, -0.28270797985312e-5,
       -0.85205128120103e-9, -0.22425281908000e-5, -0.65171222895601e-6, -0.14341729937924e-12, -0.40516996860117e-6,
       -0.12734301741641e-8, -0.17424871230634e-9, -0.68762131295531e-18, 0.14478307828521e-19, 0.26335781662795e-22,
       -0.11947622640071e-22, 0.18228094581404e-23, -0.93537087292458e-25)
_I1 = (0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 3, 3, 3, 4, 4, 4, 5, 8, 8, 21, 23, 29, 30, 31, 32)
_J1 = (-2, -1, 0, 1, 2, 3, 4, 5, -9, -7, -1, 0, 1, 3, -3, 0, 1, 3, 17, -4, 0, 6, -5, -2, 10, -8, -11, -6, -29, -31, -38, -39, -40, -41)

_N4 = (0.11670521452767e4, -0.72421316703206e6, -0.17073846940092e2, 0.12020824702470e5, -0.32325550322333e7,
       0.14915108613530e2, -0.48232657361591e4, 0.40511340542057e6, -0.23855557567849, 0.65017534844798e3)

_HIJ = ((5.20094e-1, 2.22531e-1, -2.81378e-1, 1.61913e-1, -3.25372e-2, 0.0, 0.0),
        (8.50895e-2, 9.99115e-1, -9.06851e-1, 2.57399e-1, 0.0, 0.0, 0.0),
        (-1.08374, 1.88797, -7.72479e-1, 0.0, 0.0, 0.0, 0.0),
        (-2.89555e-1, 1.26613, -4.89837e-1, 0.0, 6.98452e-2, 0.0, -4.35673e-3),
        (0.0, 0.0, -2.57040e-1, 0.0, 0.0, 8.72102e-3, 0.0),
        (0.0, 1.20573e-1, 0.0, 0.0, 0.0, 0.0, -5.93264e-4))
_H0 = (1.67752, 2.20462, 0.6366564, -0.241605)


def region1_dgamma_dpi(temperature: float, pressure: float) -> float:
    """iapws/region1.hh:170-185"""
    tau = 1386.0 / temperature
    pi = pressure / 16.53e6
    result = 0.0
    for n, i, j in zip(_N1, _I1, _J1):
        result += -n * i * math.pow(7.1 - pi, i - 1) * math.pow(tau - 1.222, j)
    return result


def volume_region1(temperature: float, pressure: float) -> float:
    """h2o.hh:838-844"""
    return (pressure / 16.53e6) * region1_dgamma_dpi(temperature, pressure) * RS * temperature / pressure


def saturation_pressure(temperature: float) -> float:
    """iapws/region4.hh:52-73"""
    n = _N4
    sigma = temperature + n[8] / (temperature - n[9])
    A = (sigma + n[0]) * sigma + n[1]
    B = (n[2] * sigma + n[3]) * sigma + n[4]
    Cc = (n[5] * sigma + n[6]) * sigma + n[7]
    tmp = 2 * Cc / (math.sqrt(B * B - 4 * A * Cc) - B)
    tmp *= tmp
    tmp *= tmp
    return 1e6 * tmp


def vapor_pressure(temperature: float) -> float:
    """h2o.hh:118-126"""
    T = min(temperature, T_CRIT)
    T = max(T, T_TRIPLE)
    return saturation_pressure(T)


def liquid_density(temperature: float, pressure: float) -> float:
    """h2o.hh:601-641 (with the straight-line extrapolation below the vapour pressure)"""
    pv = vapor_pressure(temperature)
    if pressure < pv:
        eps = pv * 1e-8
        v0 = volume_region1(temperature, pv)
        v1 = volume_region1(temperature, pv + eps)
        dv_dp = (v1 - v0) / eps
        drho_dp = -1 / (v0 * v0) * dv_dp
        return 1.0 / v0 + (pressure - pv) * drho_dp
    return 1 / volume_region1(temperature, pressure)


def viscosity(temperature: float, rho: float) -> float:
    """iapws/common.hh:88-134"""
    rho_bar = rho / RHO_CRIT
    t_bar = temperature / T_CRIT
    tmp3 = 1.0
    mu_bar = 0.0
    for i in range(6):
        tmp = 0.0
        tmp2 = 1.0
        for j in range(7):
            tmp += _HIJ[i][j] * tmp2
            tmp2 *= (rho_bar - 1)
        mu_bar += tmp3 * tmp
        tmp3 *= 1.0 / t_bar - 1
    mu_bar *= rho_bar
    mu_bar = math.exp(mu_bar)
    mu_bar *= 100 * math.sqrt(t_bar)
    tmp, tmp2 = 0.0, 1.0
    for i in range(4):
        tmp += _H0[i] / tmp2
        tmp2 *= t_bar
    mu_bar /= tmp
    return 1e-6 * mu_bar


def liquid_viscosity(temperature: float, pressure: float) -> float:
    """h2o.hh:722-728"""
    return viscosity(temperature, liquid_density(temperature, pressure))


def tabulated_h2o(temp_min=273.15, temp_max=294.15, n_temp=10, press_min=1.0e4, press_max=1.0e6, n_press=200,
                  temperature=293.15) -> dict:
    """TabulatedComponent<H2O>::init (useVaporPressure = true): values[iT + iP*nT], per-temperature pressure range
    [max(pressMin, pv/1.1), max(pressMax, pv*1.1)] (tabulatedcomponent.hh:235-253), samples
    T = iT*(Tmax-Tmin)/(nT-1) + Tmin, p = iP*(pMax-pMin)/(nP-1) + pMin (:330-345)."""
    pmin = np.empty(n_temp)
    pmax = np.empty(n_temp)
    rho = np.empty(n_temp * n_press)
    mu = np.empty(n_temp * n_press)
    for it in range(n_temp):
        T = it * (temp_max - temp_min) / (n_temp - 1) + temp_min
        pv = vapor_pressure(T)
        pmin[it] = max(press_min, pv / 1.1)
        pmax[it] = max(press_max, pv * 1.1)
        for ip in range(n_press):
            p = ip * (pmax[it] - pmin[it]) / (n_press - 1) + pmin[it]
            rho[it + ip * n_temp] = liquid_density(T, p)
            mu[it + ip * n_temp] = liquid_viscosity(T, p)
    return {"nT": n_temp, "nP": n_press, "Tmin": temp_min, "Tmax": temp_max, "pmin": pmin, "pmax": pmax, "rho": rho, "mu": mu,
            "T": temperature}
